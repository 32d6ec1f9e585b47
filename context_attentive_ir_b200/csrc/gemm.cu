// Generic fp32 GEMM  C[M,N] = act(A[M,K] W[N,K]^T + bias[N])  on CUDA cores.
//
// A rows are either dense or gathered on the fly from an embedding table by token id
// (optionally a window of `win` consecutive tokens concatenated along K: the im2col of a
// valid Conv1d), so the [tokens, E] embedded tensor of the reference
// (neuroir/modules/embeddings.py:243-252) is never materialised in HBM.
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, float4 global loads.
#include "common.cuh"

namespace cair {

constexpr int BM = 64, BN = 64, BK = 16, PADT = 4;

// scalar element for the non-vectorisable case (K or E not a multiple of 4): plain / gathered / pooled providers
__device__ __forceinline__ float a_load1(const GemmA& a, int64_t r, int kk) {
  if (a.table) {
    const int64_t seq = r / a.T;
    const int t = (int)(r - seq * a.T);
    const int seg = kk / a.E;
    const int pos = t + seg - a.pad;
    if (pos < 0 || pos >= a.L) return 0.f;
    return a.table[checked_id(a.ids[seq * a.L + pos], a.V, a.err) * a.E + (kk - seg * a.E)];
  }
  if (a.dwin == 1) {
    const int64_t seq = r / a.L;
    const int seg = kk / a.E;
    const int pos = (int)(r - seq * a.L) + seg - a.pad;
    if (pos < 0 || pos >= a.L) return 0.f;
    return a.dense[(seq * a.L + pos) * a.lda + (kk - seg * a.E)];
  }
  if (a.dwin == 2) {
    const int64_t seq = r / a.L;
    const int p = (int)(r - seq * a.L);
    const int y = p / a.Wm, x = p - y * a.Wm;
    const int seg = kk / a.E;
    const int yy = y + seg / 3 - 1, xx = x + seg % 3 - 1;
    if (yy < 0 || yy >= a.L / a.Wm || xx < 0 || xx >= a.Wm) return 0.f;
    return a.dense[(seq * a.L + yy * a.Wm + xx) * a.lda + (kk - seg * a.E)];
  }
  if (!a.pool) return a.dense[r * a.lda + kk];
  const int64_t seq = r / a.T;
  const float* base = a.dense + (seq * a.L + (r - seq * a.T)) * a.lda + kk;
  float m = base[0];
  for (int k = 1; k < a.win; ++k) m = fmaxf(m, base[k * a.lda]);
  return m;
}

template <bool VEC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmA a, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ c,
                                                       int64_t ldc, int64_t M, int N, int K, int act) {
  __shared__ __align__(16) float As[BK][BM + PADT];
  __shared__ __align__(16) float Bs[BK][BN + PADT];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t arow = m0 + lrow;
  const int brow = n0 + lrow;
  for (int k0 = 0; k0 < K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int kk = k0 + lk;
    if (arow < M) {
      if (VEC) {
        if (kk < K) {
          float4 v = gemm_a_load4(a, arow, kk);
          av[0] = v.x, av[1] = v.y, av[2] = v.z, av[3] = v.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (kk + u < K) av[u] = a_load1(a, arow, kk + u);
      }
    }
    if (brow < N) {
      if (VEC) {
        if (kk < K) {
          float4 v = *reinterpret_cast<const float4*>(w + (int64_t)brow * K + kk);
          bv[0] = v.x, bv[1] = v.y, bv[2] = v.z, bv[3] = v.w;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (kk + u < K) bv[u] = w[(int64_t)brow * K + kk + u];
      }
    }
    __syncthreads();  // previous tile fully consumed
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      As[lk + u][lrow] = av[u];
      Bs[lk + u][lrow] = bv[u];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      float v = acc[i][j] + (bias ? bias[col] : 0.f);
      if (act == ACT_TANH) v = tanhf(v);
      if (act == ACT_RELU) v = fmaxf(v, 0.f);
      c[r * ldc + col] = v;
    }
  }
}

int32_t gemm_f32(const GemmA& a, const float* w, const float* bias, float* c, int64_t ldc, int64_t M, int N,
                 int K, Act act, cudaStream_t s) {
  if (M <= 0 || N <= 0 || K <= 0) return CAIR_OK;
  if ((a.table || a.dwin) && K != a.win * a.E) return fail(CAIR_ERR_BAD_ARG, "gemm_f32: K != win*E");
  bool vec = (K % 4 == 0) && ((uintptr_t)w % 16 == 0);
  if (a.table)
    vec = vec && (a.E % 4 == 0) && ((uintptr_t)a.table % 16 == 0);
  else if (a.dwin)
    vec = vec && (a.E % 4 == 0) && (a.lda % 4 == 0) && ((uintptr_t)a.dense % 16 == 0);
  else
    vec = vec && (a.lda % 4 == 0) && ((uintptr_t)a.dense % 16 == 0);
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  if (vec)
    CAIR_LAUNCH(gemm_f32_kernel<true>, grid, 256, 0, s, a, w, bias, c, ldc, M, N, K, (int)act);
  else
    CAIR_LAUNCH(gemm_f32_kernel<false>, grid, 256, 0, s, a, w, bias, c, ldc, M, N, K, (int)act);
  return CAIR_OK;
}

}  // namespace cair
