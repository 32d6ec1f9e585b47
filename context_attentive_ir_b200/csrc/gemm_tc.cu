// Generic GEMM  C[M,N] = act(A[M,K] W[N,K]^T + bias[N])  on the 5th-gen tensor cores (tcgen05, bf16x3 split
// precision, fp32 accumulate in TMEM).  Same A-row providers as gemm.cu: dense rows, embedding rows gathered by
// token id (optionally a window of `win` consecutive tokens = the im2col of a valid Conv1d) or max-pooled over a
// time window on load - none of those tensors is ever materialised.
//
// One CTA computes a 128-row x NT-column tile (NT <= 256): the weights are pre-packed ONCE (gemm_tc_pack) as bf16
// hi/lo operand images per (column tile, 64-wide K chunk) and streamed with cp.async.bulk; the A chunk is loaded
// fp32 by 128 loader threads (thread <-> row, 16 independent 128-bit loads in flight), split into hi/lo bf16 and
// written in operand-image order; one elected lane issues 4 k-steps x 3 passes of tcgen05.mma per chunk; the
// epilogue reads TMEM (thread <-> row), adds bias, applies the activation and stores 128-byte row segments.
// Ring of 2 stages: loads of chunk k+1 overlap the MMAs of chunk k.
#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int GT_BM = 128;        // rows per CTA
constexpr int GT_BK = 64;         // K chunk
constexpr int GT_STAGES = 2;
constexpr int GT_THREADS = 192;   // warps 0-3 loaders + epilogue, warp 4 W producer, warp 5 MMA issuer
constexpr uint32_t GT_APLANE = GT_BM * 16;
constexpr uint32_t GT_AIMG = (GT_BK / 8) * GT_APLANE;  // one (hi|lo) A chunk image: 16 KB

int g_gemm_impl = 1;  // 1: tcgen05 where the shape allows, 0: always the fp32 CUDA-core kernel

// W image: [column tile][K chunk][hi|lo][plane kc][row n < NT][8 x bf16]
__global__ void gemm_tc_pack_kernel(const float* __restrict__ w, int N, int K, int NT, int nkc, uint8_t* __restrict__ img) {
  const int ct = blockIdx.y, kcnk = blockIdx.x;
  const size_t half = (size_t)(GT_BK / 8) * NT * 16;
  uint8_t* out = img + ((size_t)ct * nkc + kcnk) * 2 * half;
  for (int u = threadIdx.x; u < (GT_BK / 8) * NT; u += blockDim.x) {
    const int kc = u / NT, n = u - kc * NT;
    const int col = ct * NT + n;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int k = kcnk * GT_BK + kc * 8 + 2 * e + z;
        v[z] = (col < N && k < K) ? w[(size_t)col * K + k] : 0.f;
      }
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[0], h0, l0);
      split_bf16(v[1], h1, l1);
      hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t off = ((size_t)kc * NT + n) * 16;
    *reinterpret_cast<uint4*>(out + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out + half + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

int32_t gemm_tc_pack(Owned& own, const float* w, int N, int K, GemmTcW* out, cudaStream_t s) {
  out->N = N, out->K = K;
  out->nct = (N + 255) / 256;                                  // column tiles
  out->NT = (((N + out->nct - 1) / out->nct) + 15) & ~15;      // columns per tile, multiple of 16, <= 256
  out->nkc = (K + GT_BK - 1) / GT_BK;
  const size_t bytes = (size_t)out->nct * out->nkc * 2 * (GT_BK / 8) * out->NT * 16;
  CAIR_CUDA(own.alloc(&out->img, bytes));
  CAIR_LAUNCH(gemm_tc_pack_kernel, dim3(out->nkc, out->nct), 256, 0, s, w, N, K, out->NT, out->nkc, out->img);
  return CAIR_OK;
}

__device__ __forceinline__ void gt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// smem: A ring [stages][hi|lo] | W ring [stages][hi|lo]
__global__ void __launch_bounds__(GT_THREADS, 1)
    gemm_tc_kernel(GemmA a, const uint8_t* __restrict__ wimg, const float* __restrict__ bias, float* __restrict__ c,
                   int64_t ldc, int64_t M, int N, int K, int NT, int nkc, int act, uint32_t tcols) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t a_full[GT_STAGES], w_full[GT_STAGES], empty[GT_STAGES], acc_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * GT_BM;
  const int ct = blockIdx.y;
  const uint32_t w_plane = (uint32_t)NT * 16, w_half = (GT_BK / 8) * w_plane;
  uint8_t* a_ring = smraw;
  uint8_t* w_ring = a_ring + GT_STAGES * 2 * GT_AIMG;

  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 160) {
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&a_full[s], 4);   // one arrive per loader warp
      mbar_init(&w_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  if (warp == 4) {
    // ---- W producer ----
    if (lane == 0) {
      const uint8_t* src = wimg + (size_t)ct * nkc * 2 * w_half;
      for (int kc = 0; kc < nkc; ++kc) {
        const int s = kc % GT_STAGES;
        mbar_wait_relaxed(&empty[s], ((kc / GT_STAGES) & 1) ^ 1);
        const uint32_t bytes = 2 * w_half;
        mbar_arrive_expect_tx(&w_full[s], bytes);
        uint8_t* dst = w_ring + (size_t)s * bytes;
        for (uint32_t o = 0; o < bytes; o += 32768) bulk_g2s(dst + o, src + (size_t)kc * bytes + o, min(32768u, bytes - o), &w_full[s]);
      }
    }
  } else if (warp == 5) {
    // ---- MMA issuer ----
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, NT);
    const uint64_t ad0 = smem_desc(smem_u32(a_ring), GT_APLANE, 128);
    const uint64_t wd0 = smem_desc(smem_u32(w_ring), w_plane, 128);
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc % GT_STAGES;
      const uint32_t ph = (kc / GT_STAGES) & 1;
      mbar_wait(&a_full[s], ph);
      mbar_wait(&w_full[s], ph);
      tc_fence_after();
      const uint64_t ah = ad0 + (uint64_t)((uint32_t)s * 2 * GT_AIMG >> 4), al = ah + (uint64_t)(GT_AIMG >> 4);
      const uint64_t wh = wd0 + (uint64_t)((uint32_t)s * 2 * w_half >> 4), wl = wh + (uint64_t)(w_half >> 4);
#pragma unroll
      for (int ks = 0; ks < GT_BK / 16; ++ks) {
        const uint64_t ao = (uint64_t)(ks * ((2 * GT_APLANE) >> 4)), wo = (uint64_t)(ks * ((2 * w_plane) >> 4));
        mma_bf16_ss_w(tbase, ah + ao, wh + wo, idesc, (uint32_t)((kc | ks) != 0), issue);
        mma_bf16_ss_w(tbase, al + ao, wh + wo, idesc, 1, issue);
        mma_bf16_ss_w(tbase, ah + ao, wl + wo, idesc, 1, issue);
      }
      mma_commit_w(&empty[s], issue);
    }
    mma_commit_w(&acc_full, issue);
  } else {
    // ---- A loaders (thread <-> row), then epilogue ----
    const int64_t r = m0 + tid;
    const bool rvalid = r < M;
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc % GT_STAGES;
      mbar_wait_relaxed(&empty[s], ((kc / GT_STAGES) & 1) ^ 1);
      float4 v[GT_BK / 4];
#pragma unroll
      for (int i = 0; i < GT_BK / 4; ++i) {
        const int kk = kc * GT_BK + i * 4;
        v[i] = (rvalid && kk < K) ? gemm_a_load4(a, r, kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint8_t* ah = a_ring + (size_t)s * 2 * GT_AIMG;
#pragma unroll
      for (int pl = 0; pl < GT_BK / 8; ++pl) {
        const float x[8] = {v[2 * pl].x, v[2 * pl].y, v[2 * pl].z, v[2 * pl].w,
                            v[2 * pl + 1].x, v[2 * pl + 1].y, v[2 * pl + 1].z, v[2 * pl + 1].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(x[2 * e], h0, l0);
          split_bf16(x[2 * e + 1], h1, l1);
          hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const size_t off = (size_t)pl * GT_APLANE + (size_t)tid * 16;
        *reinterpret_cast<uint4*>(ah + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(ah + GT_AIMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) gt_arrive(&a_full[s]);
    }
    // ---- epilogue ----
    mbar_wait_relaxed(&acc_full, 0);
    tc_fence_after();
    const int n0 = ct * NT;
    for (int c0 = 0; c0 < NT; c0 += 32) {
      float v[32];
      tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (rvalid) {
        float* crow = c + r * ldc + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          if (c0 + j < NT && col < N) {
            float x = v[j] + (bias ? bias[col] : 0.f);
            if (act == ACT_TANH) x = tanhf(x);
            if (act == ACT_RELU) x = fmaxf(x, 0.f);
            crow[j] = x;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, tcols);
}

bool gemm_tc_usable(const GemmA& a, int K) {
  if (!g_gemm_impl) return false;
  if (K % 4) return false;
  if (a.table) return (a.E % 4 == 0) && ((uintptr_t)a.table % 16 == 0);
  if (a.dwin) return (a.E % 4 == 0) && (a.lda % 4 == 0) && ((uintptr_t)a.dense % 16 == 0);
  return (a.lda % 4 == 0) && ((uintptr_t)a.dense % 16 == 0);
}

int32_t gemm_tc(const GemmA& a, const GemmTcW& w, const float* bias, float* c, int64_t ldc, int64_t M, Act act,
                cudaStream_t s) {
  if (M <= 0) return CAIR_OK;
  if ((a.table || a.dwin) && w.K != a.win * a.E) return fail(CAIR_ERR_BAD_ARG, "gemm_tc: K != win*E");
  const size_t smem = (size_t)GT_STAGES * 2 * GT_AIMG + (size_t)GT_STAGES * 2 * (GT_BK / 8) * w.NT * 16;
  uint32_t tcols = 32;
  while ((int)tcols < ((w.NT + 31) & ~31)) tcols <<= 1;
  CAIR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + GT_BM - 1) / GT_BM), (unsigned)w.nct);
  CAIR_LAUNCH(gemm_tc_kernel, grid, GT_THREADS, smem, s, a, w.img, bias, c, ldc, M, w.N, w.K, w.NT, w.nkc, (int)act, tcols);
  return CAIR_OK;
}

}  // namespace cair
