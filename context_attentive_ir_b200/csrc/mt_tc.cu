// Match-Tensor interaction on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
// Same algebra as mt.cu (merged 3x7 stencil, factorised over the rank-1 product channels):
//
//   conv[j, (i,f)] = sum_{bt<7} sum_{c<C} cd[j+bt-3, c] * T[i, bt, c, f]
//
// as a GEMM with M = document positions j (128 rows per tile), N = 96 columns = IPT query positions x
// FP filters, K = C (padded to CP, a multiple of 16) per tap; the 7 taps accumulate into the same TMEM
// tile and read ONE staged copy of the document rows through row-shifted shared-memory descriptors
// (no-swizzle K-major layout, see umma.cuh) - no im2col, no [BN,C+1,Lq,Ld] tensor.
// Precision: bf16x3 split (hi*hi + lo*hi + hi*lo, fp32 accumulate) - operands are fp32-accurate to
// ~2^-17, which the 1e-3 parity bar needs after two max-pools; plain bf16 does not hold it.
// One CTA per (query, column tile): the B operand (T, 7*CP x 96, hi+lo = 172 KB at C=50) is bulk-copied
// (cp.async.bulk -> mbarrier) into shared memory ONCE and reused by all candidate documents of the query;
// per (doc, row tile): stage A (fp32 -> hi/lo bf16), 84 MMAs issued by one thread, tcgen05.ld epilogue:
// exact-match taps + bias + ReLU + 1x1 conv + running max, folded into a per-pair max buffer with
// float atomics; a last tiny kernel applies the output Linear.
#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int TC_NROWS = 96;     // N of the MMA (columns of the accumulator tile)
constexpr int TC_TCOLS = 128;    // TMEM columns allocated (power of two >= TC_NROWS)
constexpr int TC_RA = 136;       // staged A rows: 128 + 6 halo rows, rounded up to 8
constexpr int TC_THREADS = 128;

size_t mt_tc_image_bytes(int CP) { return (size_t)2 * 7 * (CP / 8) * TC_NROWS * 16; }

// B image for (query qi, column tile nt): [hi|lo][plane = bt*CP/8 + c/8][row n = il*FP + f][8 x bf16]
__global__ void __launch_bounds__(256) mt_tc_build_t_kernel(const float* __restrict__ cq, MtPack p, int Lq, int CP,
                                                            int IPT, int ntiles, uint8_t* __restrict__ img) {
  const int nt = blockIdx.x, qi = blockIdx.y;
  const int C = p.C, FP = p.FP, FPP = p.FPP;
  const int planes = 7 * (CP / 8);
  const size_t half = (size_t)planes * TC_NROWS * 16;
  uint8_t* out = img + ((size_t)qi * ntiles + nt) * 2 * half;
  const int total = planes * TC_NROWS * 8;  // one element per (plane, row, e)
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int e = idx & 7, n = (idx >> 3) % TC_NROWS, pl = idx / (8 * TC_NROWS);
    const int bt = pl / (CP / 8), c = (pl - bt * (CP / 8)) * 8 + e;
    const int il = n / FP, f = n - il * FP, i = nt * IPT + il;
    float v = 0.f;
    if (il < IPT && i < Lq && c < C) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        int ii = i + a - 1;
        if (ii >= 0 && ii < Lq)
          v = fmaf(p.w7[(((size_t)a * 7 + bt) * C + c) * FPP + f], cq[((size_t)qi * Lq + ii) * C + c], v);
      }
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const size_t off = ((size_t)pl * TC_NROWS + n) * 16 + e * 2;
    *reinterpret_cast<__nv_bfloat16*>(out + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(out + half + off) = lo;
  }
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// dynamic smem: B image | A image (hi, lo) | wem[21*FPP] bias[FPP] w1[M*FPP] b1[32] red[32*4] | dids[Ld+6+128] qids[Lq]
template <int NF>
__global__ void __launch_bounds__(TC_THREADS, 1)
    mt_tc_interact_kernel(const float* __restrict__ cd, const uint8_t* __restrict__ timg, MtPack p,
                          const int64_t* __restrict__ q, const int64_t* __restrict__ d, int N, int Lq, int Ld, int CP,
                          int ntiles, int64_t pair_begin, int64_t pair_count, int64_t q_begin,
                          float* __restrict__ maxbuf) {
  constexpr int FP = 3 * NF, FPP = (FP + 3) & ~3, IPT = TC_NROWS / FP;
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t bar_b, bar_mma;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C, M = p.M;
  const int KC = CP / 8, planes = 7 * KC;
  const uint32_t b_plane = TC_NROWS * 16, a_plane = TC_RA * 16;
  const size_t b_half = (size_t)planes * b_plane, a_half = (size_t)KC * a_plane;
  uint8_t* b_img = smraw;
  uint8_t* a_img = b_img + 2 * b_half;
  float* wem = reinterpret_cast<float*>(a_img + 2 * a_half);
  float* bias = wem + 21 * FPP;
  float* w1 = bias + FPP;
  float* b1 = w1 + MT_TC_MAXM * FPP;
  float* red = b1 + MT_TC_MAXM;
  int* dids = reinterpret_cast<int*>(red + MT_TC_MAXM * 4);
  const int nmt = (Ld + 127) / 128;
  const int dlen_pad = nmt * 128 + 6;
  int* qids = dids + dlen_pad;

  const int nt = blockIdx.x % ntiles;
  const int64_t ql = blockIdx.x / ntiles;     // query local to the slice
  const int64_t b = q_begin + ql;             // global query
  // candidate docs of this query inside the pair slice
  int64_t p_lo = b * N, p_hi = b * N + N;
  if (p_lo < pair_begin) p_lo = pair_begin;
  if (p_hi > pair_begin + pair_count) p_hi = pair_begin + pair_count;

  if (warp == 0) tmem_alloc(&tmem_slot, TC_TCOLS);
  if (tid == 0) {
    mbar_init(&bar_b, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {  // B image: bulk copy, completion counted in bytes on bar_b
    const uint8_t* src = timg + ((size_t)ql * ntiles + nt) * 2 * b_half;
    const uint32_t total = (uint32_t)(2 * b_half);
    mbar_arrive_expect_tx(&bar_b, total);
    const uint32_t chunk = 32768;
    for (uint32_t o = 0; o < total; o += chunk) bulk_g2s(b_img + o, src + o, min(chunk, total - o), &bar_b);
  }
  for (int i = tid; i < 21 * FPP; i += TC_THREADS) wem[i] = p.wem[i];
  for (int i = tid; i < FPP; i += TC_THREADS) bias[i] = p.bias[i];
  for (int i = tid; i < M * FPP; i += TC_THREADS) w1[i] = p.w1[i];
  for (int i = tid; i < M; i += TC_THREADS) b1[i] = p.b1[i];
  for (int i = tid; i < Lq; i += TC_THREADS) qids[i] = (int)q[b * Lq + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  const uint32_t idesc = idesc_bf16_f32(128, TC_NROWS);
  uint32_t mma_phase = 0;
  bool b_ready = false;

  for (int64_t pg = p_lo; pg < p_hi; ++pg) {
    const int64_t pl = pg - pair_begin;
    for (int i = tid; i < dlen_pad; i += TC_THREADS) {
      int j = i - 3;
      dids[i] = (j >= 0 && j < Ld) ? (int)d[pg * Ld + j] : -1;
    }
    float mx[MT_TC_MAXM];
#pragma unroll
    for (int m = 0; m < MT_TC_MAXM; ++m) mx[m] = -INFINITY;
    const float* cdp = cd + (size_t)pl * Ld * C;

    for (int mt = 0; mt < nmt; ++mt) {
      // ---- stage A: rows r <-> doc position j = mt*128 + r - 3, fp32 -> (hi, lo) bf16, 16-byte units ----
      for (int u = tid; u < TC_RA * KC; u += TC_THREADS) {
        const int kc = u / TC_RA, r = u - kc * TC_RA;
        const int j = mt * 128 + r - 3;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int c = kc * 8 + e;
          v[e] = (j >= 0 && j < Ld && c < C) ? cdp[(size_t)j * C + c] : 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[2 * e], h0, l0);
          split_bf16(v[2 * e + 1], h1, l1);
          hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        const size_t off = (size_t)kc * a_plane + (size_t)r * 16;
        *reinterpret_cast<uint4*>(a_img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_img + a_half + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_proxy_async();
      __syncthreads();
      // ---- MMA: 7 taps x CP/16 k-steps x 3 split-precision passes into one TMEM tile ----
      if (tid == 0) {
        if (!b_ready) mbar_wait(&bar_b, 0);
        tc_fence_after();
        const uint32_t a0 = smem_u32(a_img), b0 = smem_u32(b_img);
        bool acc = false;
        for (int bt = 0; bt < 7; ++bt)
          for (int ks = 0; ks < CP / 16; ++ks) {
            const uint32_t ao = (uint32_t)(2 * ks) * a_plane + (uint32_t)bt * 16;
            const uint32_t bo = (uint32_t)(bt * KC + 2 * ks) * b_plane;
            const uint64_t a_hi = smem_desc(a0 + ao, a_plane, 128), a_lo = smem_desc(a0 + (uint32_t)a_half + ao, a_plane, 128);
            const uint64_t b_hi = smem_desc(b0 + bo, b_plane, 128), b_lo = smem_desc(b0 + (uint32_t)b_half + bo, b_plane, 128);
            mma_bf16_ss(tbase, a_hi, b_hi, idesc, acc);
            mma_bf16_ss(tbase, a_lo, b_hi, idesc, true);
            mma_bf16_ss(tbase, a_hi, b_lo, idesc, true);
            acc = true;
          }
        mma_commit(&bar_mma);
      }
      b_ready = true;
      mbar_wait(&bar_mma, mma_phase);
      mma_phase ^= 1;
      tc_fence_after();
      // ---- epilogue: thread <-> doc position j; columns (il, f) ----
      float acc[TC_NROWS];
#pragma unroll
      for (int c0 = 0; c0 < TC_NROWS; c0 += 32) tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, acc + c0);
      tmem_ld_wait();
      const int j = mt * 128 + tid;
      if (j < Ld) {
#pragma unroll
        for (int il = 0; il < IPT; ++il) {
          const int i = nt * IPT + il;
          if (i < Lq) {
            float y[FPP];
#pragma unroll
            for (int f = 0; f < FPP; ++f) y[f] = (f < FP) ? acc[il * FP + f] + bias[f] : 0.f;
            // exact-match channel: alpha * W7[f, C, a, bt] wherever q[i+a-1] == d[j+bt-3] (PAD==PAD counts)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const int ii = i + a - 1;
              if (ii < 0 || ii >= Lq) continue;
              const int qi = qids[ii];
#pragma unroll
              for (int bt = 0; bt < 7; ++bt) {
                if (dids[j + bt] == qi) {
                  const float* we = wem + (a * 7 + bt) * FPP;
#pragma unroll
                  for (int f = 0; f < FP; ++f) y[f] += we[f];
                }
              }
            }
#pragma unroll
            for (int f = 0; f < FPP; ++f) y[f] = fmaxf(y[f], 0.f);
#pragma unroll
            for (int m = 0; m < MT_TC_MAXM; ++m) {
              if (m < M) {
                float z = b1[m];
                const float4* wr = reinterpret_cast<const float4*>(w1 + m * FPP);
#pragma unroll
                for (int f4 = 0; f4 < FPP / 4; ++f4) {
                  float4 w4 = wr[f4];
                  z = fmaf(w4.x, y[4 * f4 + 0], z);
                  z = fmaf(w4.y, y[4 * f4 + 1], z);
                  z = fmaf(w4.z, y[4 * f4 + 2], z);
                  z = fmaf(w4.w, y[4 * f4 + 3], z);
                }
                mx[m] = fmaxf(mx[m], z);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncthreads();  // TMEM tile and A image are free again
    }
    // ---- fold this CTA's (i-range, all j) maxima into the pair's max buffer ----
#pragma unroll
    for (int m = 0; m < MT_TC_MAXM; ++m) {
      float v = warp_max(mx[m]);
      if (lane == 0) red[m * 4 + warp] = v;
    }
    __syncthreads();
    if (tid < M) {
      float v = fmaxf(fmaxf(red[tid * 4], red[tid * 4 + 1]), fmaxf(red[tid * 4 + 2], red[tid * 4 + 3]));
      atomic_max_float(maxbuf + pl * MT_TC_MAXM + tid, v);
    }
    __syncthreads();
  }
  if (!b_ready && tid == 0) mbar_wait(&bar_b, 0);  // never leave a bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TC_TCOLS);
}

// score[p] = wo . max + bo   (mtensor.py:128-130)
__global__ void mt_tc_score_kernel(const float* __restrict__ maxbuf, MtPack p, int64_t pair_begin, int64_t pair_count,
                                   float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t pl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pl >= pair_count) return;
  float v = (lane < p.M) ? maxbuf[pl * MT_TC_MAXM + lane] * p.wo[lane] : 0.f;
  v = warp_sum(v);
  if (lane == 0) scores[pair_begin + pl] = v + p.wo[p.M];
}

bool mt_tc_supported(const MtPack& p, int Lq, int Ld) {
  if (p.nf != 4 && p.nf != 6) return false;
  if (p.M > MT_TC_MAXM) return false;
  const int CP = (p.C + 15) & ~15;
  const int nmt = (Ld + 127) / 128;
  size_t smem = mt_tc_image_bytes(CP) + (size_t)2 * (CP / 8) * TC_RA * 16 +
                (size_t)(21 * p.FPP + p.FPP + MT_TC_MAXM * p.FPP + MT_TC_MAXM + MT_TC_MAXM * 4) * sizeof(float) +
                (size_t)(nmt * 128 + 6 + Lq) * sizeof(int);
  return smem <= 226 * 1024;
}

void mt_tc_workspace(const MtPack& p, int64_t nq, int64_t pc, int Lq, size_t* img_bytes, size_t* max_floats) {
  const int CP = (p.C + 15) & ~15;
  const int IPT = TC_NROWS / p.FP, ntiles = (Lq + IPT - 1) / IPT;
  *img_bytes = (size_t)nq * ntiles * mt_tc_image_bytes(CP);
  *max_floats = (size_t)pc * MT_TC_MAXM;
}

int32_t mt_tc_interact(const MtPack& p, const float* cq, const float* cd, uint8_t* timg, float* maxbuf,
                       const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pair_begin,
                       int64_t pair_count, int64_t q_begin, int64_t nq, float* scores, cudaStream_t s) {
  if (pair_count <= 0) return CAIR_OK;
  const int CP = (p.C + 15) & ~15;
  const int IPT = TC_NROWS / p.FP, ntiles = (Lq + IPT - 1) / IPT;
  const int nmt = (Ld + 127) / 128;
  prof_mark("build_T", s);
  CAIR_LAUNCH(mt_tc_build_t_kernel, dim3(ntiles, (unsigned)nq), 256, 0, s, cq, p, Lq, CP, IPT, ntiles, timg);
  const int64_t nmax = pair_count * MT_TC_MAXM;
  CAIR_LAUNCH(fill_kernel, (unsigned)((nmax + 255) / 256), 256, 0, s, maxbuf, nmax, -INFINITY);
  size_t smem = mt_tc_image_bytes(CP) + (size_t)2 * (CP / 8) * TC_RA * 16 +
                (size_t)(21 * p.FPP + p.FPP + MT_TC_MAXM * p.FPP + MT_TC_MAXM + MT_TC_MAXM * 4) * sizeof(float) +
                (size_t)(nmt * 128 + 6 + Lq) * sizeof(int);
  prof_mark("interact", s);
  const unsigned grid = (unsigned)(nq * ntiles);
  if (p.nf == 6) {
    CAIR_CUDA(cudaFuncSetAttribute(mt_tc_interact_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAIR_LAUNCH(mt_tc_interact_kernel<6>, grid, TC_THREADS, smem, s, cd, timg, p, q, d, N, Lq, Ld, CP, ntiles,
                pair_begin, pair_count, q_begin, maxbuf);
  } else {
    CAIR_CUDA(cudaFuncSetAttribute(mt_tc_interact_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAIR_LAUNCH(mt_tc_interact_kernel<4>, grid, TC_THREADS, smem, s, cd, timg, p, q, d, N, Lq, Ld, CP, ntiles,
                pair_begin, pair_count, q_begin, maxbuf);
  }
  prof_mark("score", s);
  CAIR_LAUNCH(mt_tc_score_kernel, (unsigned)((pair_count + 7) / 8), 256, 0, s, maxbuf, p, pair_begin, pair_count,
              scores);
  return CAIR_OK;
}

}  // namespace cair
