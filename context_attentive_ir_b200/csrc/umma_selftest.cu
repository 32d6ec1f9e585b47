// On-device self test of the tcgen05 conventions in umma.cuh (descriptor fields, canonical no-swizzle
// K-major layout, row-shifted start addresses, TMEM lane/column mapping, split-bf16 accumulation):
//   D[m][n] = sum_k A[m + shift][k] * B[n][k],   m < 128, n < N, one CTA, one accumulator tile.
// Driven by tests/test_parity_gpu.py against a float64 product.
#include "common.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D, int N, int K, int shift, int RA,
                                                            int split, uint32_t tcols) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = K / 8;
  const uint32_t a_plane = RA * 16, b_plane = N * 16;
  uint8_t* a_hi = sm;
  uint8_t* a_lo = a_hi + (size_t)KC * a_plane;
  uint8_t* b_hi = a_lo + (size_t)KC * a_plane;
  uint8_t* b_lo = b_hi + (size_t)KC * b_plane;

  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int idx = tid; idx < RA * K; idx += 128) {
    int r = idx / K, k = idx - r * K;
    float v = (r < 128 + shift) ? A[(size_t)r * K + k] : 0.f;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    size_t off = (size_t)(k >> 3) * a_plane + (size_t)r * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(a_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(a_lo + off) = lo;
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    int r = idx / K, k = idx - r * K;
    __nv_bfloat16 hi, lo;
    split_bf16(B[(size_t)r * K + k], hi, lo);
    size_t off = (size_t)(k >> 3) * b_plane + (size_t)r * 16 + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(b_hi + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(b_lo + off) = lo;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (split == 2) {
    // A operand from tensor memory: thread <-> row; columns [acol, acol + K/2) hold bf16 pairs (k, k+1)
    const uint32_t acol = (uint32_t)((N + 31) & ~31);
    const int m = warp * 32 + lane;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t pk[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const __nv_bfloat16 lo = __float2bfloat16_rn(A[(size_t)(m + shift) * K + k0 + 2 * e]);
        const __nv_bfloat16 hi = __float2bfloat16_rn(A[(size_t)(m + shift) * K + k0 + 2 * e + 1]);
        pk[e] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
      }
      tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + acol + (uint32_t)(k0 / 2), pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      const uint32_t issue = elect_one();
      const uint32_t idesc = idesc_bf16_f32(128, N);
      for (int ks = 0; ks < K / 16; ++ks)
        mma_bf16_ts_w(tbase, tbase + acol + (uint32_t)(ks * 8), smem_desc(smem_u32(b_hi) + (2 * ks) * b_plane, b_plane, 128),
                      idesc, (uint32_t)(ks != 0), issue);
      mma_commit_w(&bar, issue);
    }
  } else if (tid == 0) {
    const uint32_t idesc = idesc_bf16_f32(128, N);
    bool acc = false;
    for (int pass = 0; pass < (split ? 3 : 1); ++pass) {
      const uint8_t* ap = (pass == 1) ? a_lo : a_hi;
      const uint8_t* bp = (pass == 2) ? b_lo : b_hi;
      for (int ks = 0; ks < K / 16; ++ks) {
        uint64_t ad = smem_desc(smem_u32(ap) + (2 * ks) * a_plane + shift * 16, a_plane, 128);
        uint64_t bd = smem_desc(smem_u32(bp) + (2 * ks) * b_plane, b_plane, 128);
        mma_bf16_ss(tbase, ad, bd, idesc, acc);
        acc = true;
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    const int m = warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c0 + c < N) D[(size_t)m * N + c0 + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, tcols);
}

}  // namespace cair

extern "C" int32_t cair_umma_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t shift,
                                      int32_t split, void* stream) {
  using namespace cair;
  if (!A || !B || !D || N < 16 || N > 256 || N % 16 || K < 16 || K % 16 || shift < 0 || shift > 64)
    return fail(CAIR_ERR_BAD_ARG, "umma_selftest: need 16<=N<=256, N%%16==0, K%%16==0, 0<=shift<=64");
  int RA = (128 + shift + 7) & ~7;
  size_t smem = (size_t)(K / 8) * 16 * (RA + N) * 2;
  if (smem > 220 * 1024) return fail(CAIR_ERR_BAD_ARG, "umma_selftest: operands do not fit in shared memory");
  uint32_t tcols = 32;
  while ((int)tcols < N) tcols <<= 1;
  // tmem_ld32 reads 32-column groups: keep the whole last group inside the allocation
  while ((int)tcols < ((N + 31) & ~31)) tcols <<= 1;
  if (split == 2)  // room for the A operand columns (K/2) behind the accumulator
    while ((int)tcols < ((N + 31) & ~31) + K / 2) tcols <<= 1;
  CAIR_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(umma_selftest_kernel, 1, 128, smem, (cudaStream_t)stream, A, B, D, N, K, shift, RA, split, tcols);
  return CAIR_OK;
}

// ---- issue/throughput microbenchmark: `reps` back-to-back passes of K/16 MMAs (M=128, N) on zeroed operands,
// issued concurrently by `nwarps` warps (each into its own TMEM columns) ----
namespace cair {
using namespace umma;
__global__ void __launch_bounds__(160) umma_bench_kernel(int N, int K, int reps, int uniform, int nwarps, uint32_t tcols,
                                                         int shift, long long* __restrict__ cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int KC = K / 8;
  const uint32_t a_plane = 144 * 16, b_plane = N * 16;
  for (int i = tid; i < (int)((size_t)KC * (a_plane + b_plane) / 16); i += 160) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t idesc = idesc_bf16_f32(128, N);
  const uint32_t a0 = smem_u32(sm), b0 = a0 + KC * a_plane;
  if (warp >= 1 && warp <= nwarps) {
    const uint32_t tbase = tmem_slot + (uint32_t)(warp - 1) * N;
    long long t0 = clock64();
    if (uniform) {
      const uint32_t issue = elect_one();
      const uint64_t ad = smem_desc(a0 + (uint32_t)(shift & 15) * 16, a_plane, 128), bd = smem_desc(b0, b_plane, 128);
      if (uniform & 2) {
        const uint32_t acol = tmem_slot + 256;  // operand columns (contents irrelevant for timing)
        for (int r = 0; r < reps; ++r)
          for (int ks = 0; ks < K / 16; ++ks)
            mma_bf16_ts_w(tbase, acol + (uint32_t)(ks * 8), bd + (uint64_t)(ks * ((2 * b_plane) >> 4)), idesc, 1, issue);
      } else {
        const int cperiod = shift >> 8;
        int since = 0;
        for (int r = 0; r < reps; ++r)
          for (int ks = 0; ks < K / 16; ++ks) {
            mma_bf16_ss_w(tbase, ad + (uint64_t)(ks * ((2 * a_plane) >> 4)), bd + (uint64_t)(ks * ((2 * b_plane) >> 4)), idesc, 1, issue);
            if (cperiod && ++since == cperiod) {
              since = 0;
              mma_commit_w(&bar[3], issue);   // nobody waits on it: cost of the commit itself in the MMA stream
            }
          }
      }
      long long t1 = clock64();
      mma_commit_w(&bar[warp - 1], issue);
      mbar_wait(&bar[warp - 1], 0);
      long long t2 = clock64();
      if (issue) cycles[2 * (warp - 1)] = t1 - t0, cycles[2 * (warp - 1) + 1] = t2 - t0;
    } else if ((tid & 31) == 0) {
      for (int r = 0; r < reps; ++r)
        for (int ks = 0; ks < K / 16; ++ks)
          mma_bf16_ss(tbase, smem_desc(a0 + (2 * ks) * a_plane, a_plane, 128), smem_desc(b0 + (2 * ks) * b_plane, b_plane, 128), idesc, true);
      long long t1 = clock64();
      mma_commit(&bar[warp - 1]);
      mbar_wait(&bar[warp - 1], 0);
      long long t2 = clock64();
      cycles[2 * (warp - 1)] = t1 - t0, cycles[2 * (warp - 1) + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, tcols);
}
}  // namespace cair

extern "C" CAIR_API int32_t cair_umma_bench(int32_t N, int32_t K, int32_t reps, int32_t uniform, long long* cycles,
                                            void* stream) {
  using namespace cair;
  // uniform: bit 0 = warp-uniform issue loop, bits 4..7 = number of concurrently issuing warps (default 1),
  // bits 8..11 = row shift of the A descriptor (the row-shifted conv taps of the Match-Tensor kernel)
  int nwarps = (uniform >> 4) & 15;
  const int shift = (uniform >> 8) & 15;
  const int cperiod = (uniform >> 12) & 255;   // bits 12..19: tcgen05.commit (to a spare mbarrier) after every cperiod MMAs, 0 = never
  if (nwarps < 1) nwarps = 1;
  if (nwarps > 4 || nwarps * N > 512) return fail(CAIR_ERR_BAD_ARG, "umma_bench: too many warps / columns");
  size_t smem = (size_t)(K / 8) * 16 * (144 + N);
  uint32_t tcols = 32;
  while ((int)tcols < nwarps * N) tcols <<= 1;
  if (uniform & 2) tcols = 512;
  CAIR_CUDA(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(umma_bench_kernel, 1, 160, smem, (cudaStream_t)stream, N, K, reps, uniform & 3, nwarps, tcols, shift | (cperiod << 8), cycles);
  return CAIR_OK;
}
