// Match-Tensor interaction on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
// Same algebra as mt.cu (merged 3x7 stencil, factorised over the rank-1 product channels):
//
//   conv[j, (i,f)] = sum_{bt<7} sum_{c<C} cd[j+bt-3, c] * T[i, bt, c, f]
//
// as a GEMM with M = document positions j (128 rows per tile, up to 2 tiles), N = 96 columns = IPT query
// positions x FP filters per column tile, K = C (padded to CP, a multiple of 16) per tap; the 7 taps
// accumulate into the same TMEM tile and read ONE staged copy of the document rows through row-shifted
// shared-memory descriptors (no-swizzle K-major layout, see umma.cuh): no im2col, no [BN,C+1,Lq,Ld] tensor.
// Precision: bf16x3 split (hi*hi + lo*hi + hi*lo, fp32 accumulate): operands are accurate to ~2^-17,
// which the 1e-3 parity bar needs after two max-pools; plain bf16 does not hold it.
//
// Persistent, warp-specialised kernel, one (query, doc) pair at a time per CTA (320 threads):
//   warp 8      producer: streams the B operand (T of the pair's query, one 24 KB hi+lo slab per tap and
//               column tile, pre-packed in shared-memory image order) with cp.async.bulk into a ring of
//               stages guarded by full/empty mbarriers.  T is shared by the N docs of a query -> L2 hits.
//   warp 9      MMA issuer (one lane): per column tile 7 taps x CP/16 k-steps x 3 passes x 2 row tiles,
//               tcgen05.commit releases the B stage / publishes the accumulator stage.
//   warps 0-7   two epilogue warpgroups (one per 128-row tile): stage A (fp32 -> hi/lo bf16) once per
//               pair, then per column tile tcgen05.ld -> exact-match taps + bias + ReLU + 1x1 conv +
//               running max in registers; accumulators are double-buffered in TMEM (4 x 96 columns) so
//               the epilogue of tile t overlaps the MMAs of tile t+1.  Last: max over rows, Linear, ONE
//               4-byte score store per pair.
#include <cstring>
#include <vector>

#include <cstdlib>

#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int TC_NROWS = 96;      // N of the MMA (columns of one accumulator tile)
constexpr int TC_TCOLS = 512;     // TMEM columns allocated: 2 stages x 2 row tiles x 96 (384) -> next power of two
constexpr int TC_MAXRA = 264;     // staged A rows: 2*128 + 6 halo rows, rounded up to 8
constexpr int TC_EPI_WARPS = 8;  // 3 per SM sub-partition (= TMEM lane quarter)
constexpr int TC_NI = 2;          // query positions per epilogue work unit (2 = shared weight loads; measured slower: 0.344 vs 0.295 ms)
constexpr int TC_THREADS = (TC_EPI_WARPS + 4) * 32;   // epilogue warps + B producer + MMA issuer 0 + A producer + MMA issuer 1
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;

static inline int tc_cp(int C) { return (C + 15) & ~15; }   // channel stride of the packed stencil weights (w7t) and N of the projection GEMM

// K layout of the interaction GEMM.  Channels [0, Cm) (Cm a multiple of 16) are the "main" channels: Cm/16 k-steps per
// tap over KCm = Cm/8 operand planes.  The remaining r = C - Cm channels (1 <= r <= 8) would cost a whole extra k-step
// per tap if padded to 16; they are packed ACROSS taps into one extra "tail" plane instead: a 16-byte unit holds tpc
// taps x rr channel slots (rr = 8 / tpc >= r) - the unit of image row r' carries the tail channels of the document
// positions r'-3 .. r'-3+tpc-1.  One tail k-step covers 2 tpc taps: its two K chunks are the SAME plane at row shifts 0
// and tpc (descriptor LBO = tpc rows), so the 7 taps cost ntail = ceil(ceil(7/tpc)/2) k-steps instead of 7.
// C = 50 (hyparam default): 7 x 3 + 1 = 22 k-steps per column tile instead of 28.
struct TcK {
  int Cm, KCm, nkm, r, tpc, rr, ntail, KA;
};
__host__ __device__ inline TcK tc_k(int C) {
  TcK k;
  k.r = C & 15;
  if (k.r == 0 || k.r > 8 || C < 16) {
    k.Cm = (C + 15) & ~15, k.r = 0, k.tpc = 0, k.rr = 0, k.ntail = 0;
  } else {
    k.Cm = C & ~15;
    k.tpc = k.r <= 2 ? 4 : k.r <= 4 ? 2 : 1;
    k.rr = 8 / k.tpc;
    k.ntail = ((7 + k.tpc - 1) / k.tpc + 1) / 2;
  }
  k.KCm = k.Cm / 8, k.nkm = k.Cm / 16, k.KA = k.KCm + (k.r ? 1 : 0);
  return k;
}
__host__ __device__ inline uint32_t tc_slab_main(const TcK& k) { return 2u * k.KCm * TC_NROWS * 16; }      // hi + lo, one tap
__host__ __device__ inline uint32_t tc_slab_tail(const TcK& k) { return 2u * 2 * k.ntail * TC_NROWS * 16; }  // hi + lo, all taps
__host__ __device__ inline uint32_t tc_stage_bytes(const TcK& k) {
  return tc_slab_main(k) > tc_slab_tail(k) ? tc_slab_main(k) : tc_slab_tail(k);
}
__host__ __device__ inline uint32_t tc_timg_per_tile(const TcK& k) { return 7u * tc_slab_main(k) + tc_slab_tail(k); }
static inline size_t tc_a_bytes(const TcK& k) { return (size_t)2 * k.KA * TC_MAXRA * 16; }
constexpr int TC_PADTAB = 8 * 8 * 24;   // floats of MtEpiConst::padtab, copied to shared memory by the kernel
static inline size_t tc_misc_bytes(const MtPack& p, int Lq) {
  (void)p;
  return (size_t)(MT_TC_MAXM * TC_EPI_WARPS + 24 * MT_TC_MAXM + TC_PADTAB) * sizeof(float) + (size_t)(TC_MAXRA + 8 + Lq) * sizeof(int);
}
static inline int tc_stages(const MtPack& p, int Lq) {
  const TcK k = tc_k(p.C);
  size_t fixed = tc_a_bytes(k) + tc_misc_bytes(p, Lq) + 1024;
  int s = (int)((226 * 1024 - fixed) / tc_stage_bytes(k));
  return s > 8 ? 8 : s;
}

// B slabs for (query qi, column tile nt): 7 main slabs (one per tap bt) [hi|lo][plane c/8 < KCm][row n = il*FP + f][8 x bf16]
// followed by the tail slab [hi|lo][chunk ch < 2 ntail][row n][8 x bf16] (element e of chunk ch: tap ch*tpc + e/rr,
// channel Cm + e%rr).  T = sum_a W7[f,c,a,bt] * cq[i+a-1,c], split to hi/lo.
// One CTA per (query, column tile): the IPT + 2 query rows the tile needs are staged once in shared memory (zero rows
// outside the query, zero pad channels); a thread owns one (tap, channel plane, filter) combination, keeps its 3 x 8
// stencil weights in registers and walks the IPT query positions (6 LDS.128 + 24 FMA + split + 2 x 16-byte stores per unit;
// the first version re-read weights and query rows from global memory for every unit and was instruction-bound).
constexpr int BT_THREADS = 256;
__global__ void __launch_bounds__(BT_THREADS) mt_tc_build_t_kernel(const float* __restrict__ cq, MtPack p, int Lq, int IPT,
                                                                   int ntiles, uint8_t* __restrict__ img) {
  extern __shared__ __align__(16) float bt_cq[];   // [IPT + 2][CPW]
  const TcK k = tc_k(p.C);
  const int nt = blockIdx.x, qi = blockIdx.y;
  const int C = p.C, FP = p.FP, CPW = (C + 15) & ~15;
  const int i0 = nt * IPT;
  for (int idx = threadIdx.x; idx < (IPT + 2) * CPW; idx += BT_THREADS) {
    const int r = idx / CPW, c = idx - r * CPW;
    const int ii = i0 - 1 + r;
    bt_cq[idx] = (ii >= 0 && ii < Lq && c < C) ? cq[((size_t)qi * Lq + ii) * C + c] : 0.f;
  }
  __syncthreads();
  uint8_t* tile_out = img + ((size_t)qi * ntiles + nt) * tc_timg_per_tile(k);
  // ---- main slabs: work item = (tap bt, plane kc, filter f) ----
  const size_t half = (size_t)k.KCm * TC_NROWS * 16;
  const int nwork = 7 * k.KCm * FP;
  for (int wi = threadIdx.x; wi < nwork; wi += BT_THREADS) {
    const int bt = wi / (k.KCm * FP), rem = wi - bt * (k.KCm * FP);
    const int kc = rem / FP, f = rem - kc * FP;
    float w[3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float4* w4 = reinterpret_cast<const float4*>(p.w7t + (((size_t)a * 7 + bt) * FP + f) * CPW + kc * 8);
      const float4 wa = w4[0], wb = w4[1];
      w[a][0] = wa.x, w[a][1] = wa.y, w[a][2] = wa.z, w[a][3] = wa.w;
      w[a][4] = wb.x, w[a][5] = wb.y, w[a][6] = wb.z, w[a][7] = wb.w;
    }
    uint8_t* out = tile_out + (size_t)bt * tc_slab_main(k);
    for (int il = 0; il < IPT; ++il) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (i0 + il < Lq) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {   // staged row index of query position i0 + il + a - 1 is il + a
          const float4* c4 = reinterpret_cast<const float4*>(bt_cq + (size_t)(il + a) * CPW + kc * 8);
          const float4 ca = c4[0], cb = c4[1];
          v[0] = fmaf(w[a][0], ca.x, v[0]), v[1] = fmaf(w[a][1], ca.y, v[1]);
          v[2] = fmaf(w[a][2], ca.z, v[2]), v[3] = fmaf(w[a][3], ca.w, v[3]);
          v[4] = fmaf(w[a][4], cb.x, v[4]), v[5] = fmaf(w[a][5], cb.y, v[5]);
          v[6] = fmaf(w[a][6], cb.z, v[6]), v[7] = fmaf(w[a][7], cb.w, v[7]);
        }
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
      const size_t off = ((size_t)kc * TC_NROWS + il * FP + f) * 16;
      *reinterpret_cast<uint4*>(out + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(out + half + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  // rows n >= IPT*FP of every plane (MMA N padding) are zero
  const int npad = TC_NROWS - IPT * FP;
  for (int idx = threadIdx.x; idx < 7 * k.KCm * npad; idx += BT_THREADS) {
    const int bt = idx / (k.KCm * npad), rem = idx - bt * (k.KCm * npad);
    const int kc = rem / npad, n = IPT * FP + (rem - kc * npad);
    uint8_t* out = tile_out + (size_t)bt * tc_slab_main(k) + ((size_t)kc * TC_NROWS + n) * 16;
    *reinterpret_cast<uint4*>(out) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(out + half) = make_uint4(0, 0, 0, 0);
  }
  // ---- tail slab: one thread per 16-byte unit ----
  if (k.ntail) {
    const int planes = 2 * k.ntail;
    const size_t thalf = (size_t)planes * TC_NROWS * 16;
    uint8_t* out = tile_out + (size_t)7 * tc_slab_main(k);
    for (int u = threadIdx.x; u < planes * TC_NROWS; u += BT_THREADS) {
      const int kc = u / TC_NROWS, n = u - kc * TC_NROWS;
      const int il = n / FP, f = n - il * FP;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (il < IPT && i0 + il < Lq) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int bt = kc * k.tpc + e / k.rr, ce = e % k.rr;
            if (bt < 7 && ce < k.r)
              v[e] = fmaf(p.w7t[(((size_t)a * 7 + bt) * FP + f) * CPW + k.Cm + ce], bt_cq[(size_t)(il + a) * CPW + k.Cm + ce], v[e]);
          }
        }
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
      const size_t off = ((size_t)kc * TC_NROWS + n) * 16;
      *reinterpret_cast<uint4*>(out + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(out + thalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// A images: per pair [hi|lo][plane < KA][row r <-> doc position r-3][8 x bf16]; planes < KCm hold 8 main channels each,
// plane KCm (if any) is the tail plane (see TcK); halo / pad rows and pad channels are 0.
// One thread per 16-byte unit; the interaction kernel then loads a whole image with one bulk copy.
__global__ void __launch_bounds__(256) mt_tc_image_kernel(const float* __restrict__ cd, int C, int Ld, int RA,
                                                           int64_t pair_count, uint8_t* __restrict__ aimg) {
  const TcK k = tc_k(C);
  const int64_t pl = blockIdx.x;
  const float* cdp = cd + (size_t)pl * Ld * C;
  const size_t half = (size_t)k.KA * RA * 16;
  uint8_t* out = aimg + (size_t)pl * 2 * half;
  for (int u = threadIdx.x; u < RA * k.KA; u += blockDim.x) {
    const int kc = u / RA, r = u - kc * RA;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int j, c;
      bool ok;
      if (kc < k.KCm) {
        j = r - 3, c = kc * 8 + e, ok = c < C;
      } else {
        j = r - 3 + e / k.rr, c = k.Cm + e % k.rr, ok = (e % k.rr) < k.r;
      }
      v[e] = (ok && j >= 0 && j < Ld) ? cdp[(size_t)j * C + c] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_bf16x2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
    const size_t off = ((size_t)kc * RA + r) * 16;
    *reinterpret_cast<uint4*>(out + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(out + half + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- fused document channel projection + A image (mtensor.py:108 feeding the interaction GEMM) ----
// cd = enc_d Wd^T + bd is itself a GEMM (M = document positions, K = Hd, N = C): persistent CTAs walk the
// (pair, 128-row tile) items, keep the hi/lo image of Wd resident in shared memory, stage each tile of encoder
// rows as a hi/lo bf16 operand image, run KP/16 x 3 tcgen05.mma and write the result straight from TMEM into the
// interaction kernel's A image (bias added, split to hi/lo) - the fp32 [pairs, Ld, C] tensor never exists.
constexpr int PJ_THREADS = 256;
// Plane stride 128 rows + one 16-byte pad and a 64-byte skew of the lo half: the 8 lanes of one 128-bit store phase
// (4 planes x hi/lo of one row) then hit 8 distinct 16-byte bank groups.
constexpr uint32_t PJ_APLANE = 128 * 16 + 16;
constexpr uint32_t PJ_LOSKEW = 64;

// Wd image: [hi|lo][plane k/8][row n < CP][8 x bf16], zero beyond (C, Hd)
__global__ void mt_tc_pack_wd_kernel(const float* __restrict__ w, int C, int Hd, int CP, int KP, uint8_t* __restrict__ img) {
  const size_t half = (size_t)(KP / 8) * CP * 16;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < (KP / 8) * CP; u += gridDim.x * blockDim.x) {
    const int kc = u / CP, n = u - kc * CP;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v[2];
#pragma unroll
      for (int z = 0; z < 2; ++z) {
        const int k = kc * 8 + 2 * e + z;
        v[z] = (n < C && k < Hd) ? w[(size_t)n * Hd + k] : 0.f;
      }
      split_bf16x2(v[0], v[1], hi[e], lo[e]);
    }
    const size_t off = ((size_t)kc * CP + n) * 16;
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + half + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// smem: A image [hi|lo][KP/8][128 rows (+pad)][16 B] | W image [hi|lo][KP/8][CP rows][16 B] | bias[CP]
__global__ void __launch_bounds__(PJ_THREADS, 2)
    mt_tc_proj_image_kernel(const float* __restrict__ enc, int Hd, int KP, const uint8_t* __restrict__ wimg,
                            const float* __restrict__ bd, int C, int CP, int Ld, int RA, int ntile, int64_t nitems,
                            uint32_t tcols, uint8_t* __restrict__ aimg) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t w_full, acc_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_half = (uint32_t)(KP / 8) * PJ_APLANE + PJ_LOSKEW;
  const uint32_t w_plane = (uint32_t)CP * 16, w_half = (uint32_t)(KP / 8) * w_plane;
  uint8_t* a_img = smraw;
  uint8_t* w_img = smraw + 2 * a_half;
  float* bias_s = reinterpret_cast<float*>(w_img + 2 * w_half);
  const TcK k = tc_k(C);
  const int KC = k.KCm;   // main planes of the output image; plane KC is the tail plane (k.r > 0)
  const size_t o_half = (size_t)k.KA * RA * 16;

  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 32) {
    mbar_init(&w_full, 1);
    mbar_init(&acc_full, 1);
    fence_mbar_init();
  }
  for (int c = tid; c < CP; c += PJ_THREADS) bias_s[c] = c < C ? bd[c] : 0.f;  // pad channels: W rows 0, bias 0
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t bytes = 2 * w_half;
    mbar_arrive_expect_tx(&w_full, bytes);
    for (uint32_t o = 0; o < bytes; o += 32768) bulk_g2s(w_img + o, wimg + o, min(32768u, bytes - o), &w_full);
  }
  const uint32_t issue = elect_one();
  const uint32_t idesc = idesc_bf16_f32(128, CP);
  const uint64_t ah = smem_desc(smem_u32(a_img), PJ_APLANE, 128), al = ah + (uint64_t)(a_half >> 4);
  const uint64_t wh = smem_desc(smem_u32(w_img), w_plane, 128), wl = wh + (uint64_t)(w_half >> 4);
  const int half = lane & 1;
  const int ncb = (KP + 127) / 128;  // 128-channel column blocks: one warp load = one 512-byte row segment
  const int row0 = warp * 16;        // warp <-> 16 rows of the tile
  uint32_t phase = 0;
  bool w_ready = false;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t pl = item / ntile;
    const int mt = (int)(item - pl * ntile);
    const float* src = enc + ((size_t)pl * Ld + (size_t)mt * 128) * Hd;
    const int nrow = min(128, Ld - mt * 128);
    // ---- stage the encoder rows: a lane pair owns one 8-channel unit of the row.  Rows >= nrow only feed accumulator
    //      rows the epilogue never reads, so they are loaded from a clamped (valid) address and left as they are. ----
    for (int cb = 0; cb < ncb; ++cb) {
      const int k = cb * 128 + lane * 4, kc = k >> 3;
      const float* colp = src + min(k, Hd - 4);
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        v[i] = ldg_stream(reinterpret_cast<const float4*>(colp + (size_t)min(row0 + i, nrow - 1) * Hd));
      uint8_t* dst = a_img + (half ? a_half : 0u) + (uint32_t)kc * PJ_APLANE + (uint32_t)row0 * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        uint32_t hw0, lw0, hw1, lw1;
        split_bf16x2(v[i].x, v[i].y, hw0, lw0);
        split_bf16x2(v[i].z, v[i].w, hw1, lw1);
        // even lane keeps the hi unit, odd lane the lo unit: swap the two words the partner needs
        const uint32_t r0 = __shfl_xor_sync(0xffffffffu, half ? hw0 : lw0, 1);
        const uint32_t r1 = __shfl_xor_sync(0xffffffffu, half ? hw1 : lw1, 1);
        uint4 unit = half ? make_uint4(r0, r1, lw0, lw1) : make_uint4(hw0, hw1, r0, r1);
        if (k >= Hd) unit = make_uint4(0u, 0u, 0u, 0u);  // K padding (Hd < KP)
        if (k < KP) *reinterpret_cast<uint4*>(dst + i * 16) = unit;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (!w_ready) mbar_wait(&w_full, 0), w_ready = true;
      tc_fence_after();
      for (int ks = 0; ks < KP / 16; ++ks) {
        const uint64_t ao = (uint64_t)(ks * ((2 * PJ_APLANE) >> 4)), wo = (uint64_t)(ks * ((2 * w_plane) >> 4));
        mma_bf16_ss_w(tbase, ah + ao, wh + wo, idesc, (uint32_t)(ks != 0), issue);
        mma_bf16_ss_w(tbase, al + ao, wh + wo, idesc, 1, issue);
        mma_bf16_ss_w(tbase, ah + ao, wl + wo, idesc, 1, issue);
      }
      mma_commit_w(&acc_full, issue);
    }
    // ---- halo / tail rows of the image are zero ----
    uint8_t* out = aimg + (size_t)pl * 2 * o_half;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    if (mt == 0 && tid < 3)
      for (int kc = 0; kc < KC; ++kc) {
        *reinterpret_cast<uint4*>(out + ((size_t)kc * RA + tid) * 16) = zero;
        *reinterpret_cast<uint4*>(out + o_half + ((size_t)kc * RA + tid) * 16) = zero;
      }
    if (mt == ntile - 1)
      for (int r = Ld + 3 + tid; r < RA; r += PJ_THREADS)
        for (int kc = 0; kc < KC; ++kc) {
          *reinterpret_cast<uint4*>(out + ((size_t)kc * RA + r) * 16) = zero;
          *reinterpret_cast<uint4*>(out + o_half + ((size_t)kc * RA + r) * 16) = zero;
        }
    // tail plane: every (row, tap slot) whose document position is outside [0, Ld) is zero; each byte of the plane has
    // exactly one writer (these threads for the invalid slots, the epilogue thread of position j for the valid ones)
    if (k.r && (mt == 0 || mt == ntile - 1)) {
      uint8_t* tp = out + (size_t)KC * RA * 16;
      for (int idx = tid; idx < RA * k.tpc; idx += PJ_THREADS) {
        const int r = idx / k.tpc, te = idx - r * k.tpc, pos = r - 3 + te;
        if ((pos < 0 && mt == 0) || (pos >= Ld && mt == ntile - 1)) {
          uint16_t* ph = reinterpret_cast<uint16_t*>(tp + (size_t)r * 16) + te * k.rr;
          uint16_t* plo = reinterpret_cast<uint16_t*>(tp + o_half + (size_t)r * 16) + te * k.rr;
          for (int ce = 0; ce < k.rr; ++ce) ph[ce] = 0, plo[ce] = 0;
        }
      }
    }
    mbar_wait_relaxed(&acc_full, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: lane <-> row (TMEM lane quarter warp % 4), warps 0-3 / 4-7 take alternate 16-column chunks;
    //      + bias, split, 16-byte units of the A image ----
    const int rl = (warp & 3) * 32 + lane, j = mt * 128 + rl;
    for (int c0 = (warp >> 2) * 16; c0 < CP; c0 += 32) {
      float v[16];
      tmem_ld16(tbase + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (j < Ld) {
        if (c0 < k.Cm) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t hi[4], lo[4];
            const float4 ba = *reinterpret_cast<const float4*>(bias_s + c0 + g * 8);
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + g * 8 + 4);
            split_bf16x2(v[g * 8 + 0] + ba.x, v[g * 8 + 1] + ba.y, hi[0], lo[0]);
            split_bf16x2(v[g * 8 + 2] + ba.z, v[g * 8 + 3] + ba.w, hi[1], lo[1]);
            split_bf16x2(v[g * 8 + 4] + bb.x, v[g * 8 + 5] + bb.y, hi[2], lo[2]);
            split_bf16x2(v[g * 8 + 6] + bb.z, v[g * 8 + 7] + bb.w, hi[3], lo[3]);
            const size_t off = ((size_t)(c0 / 8 + g) * RA + j + 3) * 16;
            *reinterpret_cast<uint4*>(out + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(out + o_half + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        } else {
          // tail channels Cm .. Cm+r-1 of position j: slot te of the units of image rows j+3-te, te < tpc
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = 2 * e < k.r ? v[2 * e] + bias_s[c0 + 2 * e] : 0.f;
            const float x1 = 2 * e + 1 < k.r ? v[2 * e + 1] + bias_s[c0 + 2 * e + 1] : 0.f;
            split_bf16x2(x0, x1, hi[e], lo[e]);
          }
          uint8_t* tp = out + (size_t)KC * RA * 16;
          for (int te = 0; te < k.tpc; ++te) {
            uint8_t* dh = tp + (size_t)(j + 3 - te) * 16 + te * k.rr * 2;
            uint8_t* dl = dh + o_half;
            if (k.rr == 2) {
              *reinterpret_cast<uint32_t*>(dh) = hi[0];
              *reinterpret_cast<uint32_t*>(dl) = lo[0];
            } else if (k.rr == 4) {
              *reinterpret_cast<uint2*>(dh) = make_uint2(hi[0], hi[1]);
              *reinterpret_cast<uint2*>(dl) = make_uint2(lo[0], lo[1]);
            } else {
              *reinterpret_cast<uint4*>(dh) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(dl) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM accumulator and the A image are reusable
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, tcols);
}

// acc.{x,y} += w.{x,y} * y  as ONE packed instruction (fma.rn.f32x2, SASS FFMA2 with a broadcast scalar operand)
__device__ __forceinline__ void ffma2(float2& acc, float2 w, float y) {
  unsigned long long a, c, yy;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(w.x), "f"(w.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(yy));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Optional role timing (dbg != nullptr, CTA 0 only): cycles spent in each wait / phase, see cair_mt_debug_timing.
#define TC_T0() long long t0_ = dbg ? clock64() : 0
#define TC_ACC(slot) do { if (dbg && blockIdx.x == 0 && lane == 0) dbg[slot] += clock64() - t0_; } while (0)

// TMEM -> registers, FP (= 12 or 18) consecutive fp32 columns of this thread's lane
template <int FP>
__device__ __forceinline__ void tmem_ld_fp(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  if constexpr (FP == 18) {
    tmem_ld16(taddr, v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[16]), "=r"(r[17]) : "r"(taddr + 16) : "memory");
  } else {
    static_assert(FP == 12, "FP must be 12 or 18");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11])
                 : "r"(taddr + 8)
                 : "memory");
  }
}

// y[f] = relu(conv accumulator + bias + exact-match taps) of (query position i, this thread's document row)
// exact-match channel: alpha * W7[f, C, a, bt] wherever q[i+a-1] == d[j+bt-3] (PAD==PAD counts); dj = dids + jrow
template <int FP>
__device__ __forceinline__ void mt_epi_prep(float* y, int i, int Lq, const int* dj, const int* qids, const MtEpiConst& ec,
                                            const float* padtab) {
#pragma unroll
  for (int f = 0; f < FP; ++f) y[f] += ec.bias[f];
  // Out-of-range query positions are -2, out-of-range document positions -1: they never match.
  int qv[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int ii = i + a - 1;
    qv[a] = (ii >= 0 && ii < Lq) ? qids[ii] : -2;
  }
  // PAD == PAD counts as a match (mtensor.py:156), and on real data most of the [Lq, Ld] plane is PAD x PAD (documents
  // average 63 of 200 positions, queries 4 of 20).  Those matches are separable - tap (a, bt) matches iff query tap a
  // AND document tap bt are PAD - and the PAD taps of a row are a contiguous run [lo, hi), so their contribution is a
  // difference of two pre-summed table rows: 2 x FP shared-memory loads for ANY pad pattern, the rows at the border of the
  // padding included (with a tap loop those rows made every (i, warp) unit of a padded pair take 21 x FP predicated adds).
  // Real-token matches (a few cells per pair) keep the tap loop.
  // Common case first (full-length batches, real x real cells): 21 predicate-accumulating compares and one branch.
  bool hit = false;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bt = 0; bt < 7; ++bt) hit |= dj[bt] == qv[a];
  if (hit) {
    int zm = 0, qz = 0;
    bool real = false;
#pragma unroll
    for (int bt = 0; bt < 7; ++bt) zm |= (dj[bt] == 0 ? 1 : 0) << bt;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      qz |= (qv[a] == 0 ? 1 : 0) << a;
#pragma unroll
      for (int bt = 0; bt < 7; ++bt) real |= qv[a] > 0 && dj[bt] == qv[a];
    }
    bool padgen = false;
    if (zm != 0 && qz != 0) {
      const int lo = __ffs(zm) - 1, hi = 32 - __clz(zm);
      if ((zm >> lo) == (1 << (hi - lo)) - 1 && qz != 5) {
        const float* th = padtab + (qz * 8 + hi) * 24;
        const float* tl = padtab + (qz * 8 + lo) * 24;
#pragma unroll
        for (int f = 0; f < FP; ++f) y[f] += th[f] - tl[f];
      } else {
        padgen = true;   // PAD tokens inside a sequence (not produced by batchify): the general tap loop handles them
      }
    }
    if (real || padgen) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int bt = 0; bt < 7; ++bt) {
          if (dj[bt] == qv[a] && (qv[a] > 0 || padgen)) {
#pragma unroll
            for (int f = 0; f < FP; ++f) y[f] += ec.wem[a * 7 + bt][f];
          }
        }
      }
    }
  }
#pragma unroll
  for (int f = 0; f < FP; ++f) y[f] = fmaxf(y[f], 0.f);
}

// One epilogue work unit: NI (1 or 2) consecutive query positions of one document row; M <= 20 outputs.
// 1x1 conv: f outer, output-channel pairs inner -> independent accumulators; packed fp32x2 FMAs (FFMA2) with y[f] as the
// broadcast scalar operand; the weights come from shared memory as broadcast LDS.128 (w1t[f][m], m contiguous), shared
// by the NI positions.  Outputs m >= M see zero weights and are ignored at the end.
// ARG (training forward): also track, per output channel, the cell i * Ld + jrow of this thread's running maximum (strict
// comparison: cells are visited in ascending linear order per thread, so the first maximum is kept, as torch.max does).
template <int NF, int NI, bool ARG = false>
__device__ __forceinline__ void mt_epi_unit(uint32_t tacc, int i0, int Lq, bool row_ok, const int* djp, const int* qids,
                                            const MtEpiConst& ec, const float* w1t, float* mx, int* ai = nullptr, int jrow = 0,
                                            int Ld = 0) {
  constexpr int FP = 3 * NF;
  float y[NI][FP];
#pragma unroll
  for (int n = 0; n < NI; ++n) tmem_ld_fp<FP>(tacc + n * FP, y[n]);
  tmem_ld_wait();
  if (!row_ok || i0 >= Lq) return;
  int dj[7];
#pragma unroll
  for (int bt = 0; bt < 7; ++bt) dj[bt] = djp[bt];
#pragma unroll
  for (int n = 0; n < NI; ++n) mt_epi_prep<FP>(y[n], i0 + n, Lq, dj, qids, ec, w1t + 24 * MT_TC_MAXM);
  float2 zz[NI][10];
#pragma unroll
  for (int n = 0; n < NI; ++n)
#pragma unroll
    for (int m2 = 0; m2 < 10; ++m2) zz[n][m2] = make_float2(ec.b1[2 * m2], ec.b1[2 * m2 + 1]);
#pragma unroll
  for (int f = 0; f < FP; ++f) {
#pragma unroll
    for (int m2 = 0; m2 < 10; ++m2) {
      // weights straight from the constant bank (LDCU.128 -> uniform registers -> FFMA2 operand): no shared-memory
      // traffic - the MMA operand fetches need that bandwidth
      const float2 w2 = make_float2(ec.w1t[f][2 * m2], ec.w1t[f][2 * m2 + 1]);
#pragma unroll
      for (int n = 0; n < NI; ++n) ffma2(zz[n][m2], w2, y[n][f]);
    }
  }
#pragma unroll
  for (int n = 0; n < NI; ++n) {
    if (i0 + n < Lq) {
#pragma unroll
      for (int m2 = 0; m2 < 10; ++m2) {
        if constexpr (ARG) {
          const int cell = (i0 + n) * Ld + jrow;
          if (zz[n][m2].x > mx[2 * m2]) mx[2 * m2] = zz[n][m2].x, ai[2 * m2] = cell;
          if (zz[n][m2].y > mx[2 * m2 + 1]) mx[2 * m2 + 1] = zz[n][m2].y, ai[2 * m2 + 1] = cell;
        } else {
          mx[2 * m2] = fmaxf(mx[2 * m2], zz[n][m2].x);
          mx[2 * m2 + 1] = fmaxf(mx[2 * m2 + 1], zz[n][m2].y);
        }
      }
    }
  }
}

// General-M (20 < M <= MT_TC_MAXM) unit, one query position, scalar FMAs.
template <int NF, bool ARG = false>
__device__ __forceinline__ void mt_epi_unit_wide(uint32_t tacc, int i, int Lq, bool row_ok, const int* djp, const int* qids,
                                                 const MtEpiConst& ec, const float* w1t, float* mx, int* ai = nullptr, int jrow = 0,
                                                 int Ld = 0) {
  constexpr int FP = 3 * NF;
  float y[FP];
  tmem_ld_fp<FP>(tacc, y);
  tmem_ld_wait();
  if (!row_ok || i >= Lq) return;
  int dj[7];
#pragma unroll
  for (int bt = 0; bt < 7; ++bt) dj[bt] = djp[bt];
  mt_epi_prep<FP>(y, i, Lq, dj, qids, ec, w1t + 24 * MT_TC_MAXM);
  float z[MT_TC_MAXM];
#pragma unroll
  for (int m = 0; m < MT_TC_MAXM; ++m) z[m] = ec.b1[m];
#pragma unroll
  for (int f = 0; f < FP; ++f) {
    const float4* wr = reinterpret_cast<const float4*>(w1t + f * MT_TC_MAXM);
#pragma unroll
    for (int m4 = 0; m4 < MT_TC_MAXM / 4; ++m4) {
      const float4 w4 = wr[m4];
      z[4 * m4 + 0] = fmaf(w4.x, y[f], z[4 * m4 + 0]);
      z[4 * m4 + 1] = fmaf(w4.y, y[f], z[4 * m4 + 1]);
      z[4 * m4 + 2] = fmaf(w4.z, y[f], z[4 * m4 + 2]);
      z[4 * m4 + 3] = fmaf(w4.w, y[f], z[4 * m4 + 3]);
    }
  }
#pragma unroll
  for (int m = 0; m < MT_TC_MAXM; ++m) {
    if constexpr (ARG) {
      if (z[m] > mx[m]) mx[m] = z[m], ai[m] = i * Ld + jrow;
    } else {
      mx[m] = fmaxf(mx[m], z[m]);
    }
  }
}

template <int NF, bool ARG = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
    mt_tc_interact_kernel(const uint8_t* __restrict__ aimg, const uint8_t* __restrict__ timg, MtPack p,
                          const __grid_constant__ MtEpiConst ec, const int64_t* __restrict__ q,
                          const int64_t* __restrict__ d, int N, int Lq, int Ld, int ntiles, int nstages,
                          int64_t pair_begin, int64_t pair_count, int64_t q_begin, float* __restrict__ scores,
                          long long* __restrict__ dbg, float* __restrict__ pooled, int* __restrict__ argidx) {
  constexpr int FP = 3 * NF, FPP = (FP + 3) & ~3, IPT = TC_NROWS / FP;
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t full_b[8], empty_b[8], acc_full[2], acc_empty[2], a_full, a_empty;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = p.M;
  const TcK k = tc_k(p.C);
  const int nit = 7 + (k.ntail ? 1 : 0);            // ring items per column tile: 7 taps (+ the tail slab)
  const int nmt = (Ld + 127) / 128;                 // 1 or 2 row tiles
  const int RA = (nmt * 128 + 6 + 7) & ~7;          // staged rows
  const uint32_t b_plane = TC_NROWS * 16, a_plane = (uint32_t)RA * 16;
  const uint32_t b_half = (uint32_t)k.KCm * b_plane, a_half = (uint32_t)k.KA * a_plane;
  const uint32_t slab = tc_slab_main(k), slab_t = tc_slab_tail(k), stage = tc_stage_bytes(k);
  uint8_t* a_img = smraw;
  uint8_t* b_ring = a_img + (size_t)2 * k.KA * TC_MAXRA * 16;  // host side: tc_a_bytes()
  // epilogue weights (exact-match taps, bias, 1x1 conv) come from the constant bank (kernel parameter `ec`):
  // FFMA takes them as immediate c[][] operands, no shared-memory loads in the hot loop
  float* red = reinterpret_cast<float*>(b_ring + (size_t)nstages * stage);  // [8 warps][32]
  float* w1t = red + MT_TC_MAXM * TC_EPI_WARPS;               // [24][32] 1x1 conv weights, transposed (m contiguous)
  float* padtab = w1t + 24 * MT_TC_MAXM;                      // [8][8][24] PAD x PAD exact-match sums (mt_epi_prep)
  int* dids = reinterpret_cast<int*>(padtab + TC_PADTAB);
  int* qids = dids + TC_MAXRA + 8;

  if (warp == 0) tmem_alloc(&tmem_slot, TC_TCOLS);
  if (tid == TC_EPI_WARPS * 32) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], nmt);    // one commit per MMA issuer (= per row tile)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], nmt);
      mbar_init(&acc_empty[s], TC_EPI_WARPS);   // one arrive per epilogue warp
    }
    mbar_init(&a_full, 1);
    mbar_init(&a_empty, nmt);
    fence_mbar_init();
  }
  for (int i = tid; i < 24 * MT_TC_MAXM; i += TC_THREADS) w1t[i] = ec.w1t[i / MT_TC_MAXM][i % MT_TC_MAXM];
  for (int i = tid; i < TC_PADTAB; i += TC_THREADS) padtab[i] = ec.padtab[i / 192][(i / 24) % 8][i % 24];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  // Warp roles: the scheduler favours the highest warp id of an SM sub-partition (wid % 4), so the two latency-
  // critical single-lane roles are the LAST warps of their sub-partitions: warp 12 = producer, warp 13 = MMA issuer.
  if (warp == TC_EPI_WARPS) {
    // ================= producer: B slabs through the ring =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t pl = blockIdx.x; pl < pair_count; pl += gridDim.x) {
        const int64_t ql = (pair_begin + pl) / N - q_begin;
        const uint8_t* src = timg + (size_t)ql * ntiles * tc_timg_per_tile(k);
        for (int nt = 0; nt < ntiles; ++nt) {
          for (int item = 0; item < nit; ++item, ++it) {
            const int s = it % nstages;
            const uint32_t ph = (it / nstages) & 1;
            const uint32_t bytes = item < 7 ? slab : slab_t;
            { TC_T0(); mbar_wait_relaxed(&empty_b[s], ph ^ 1); TC_ACC(0); }
            mbar_arrive_expect_tx(&full_b[s], bytes);
            bulk_g2s(b_ring + (size_t)s * stage, src, bytes, &full_b[s]);
            src += bytes;
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 2) {
    // ================= A producer: one bulk copy of the pair's document image, as soon as the previous pair's MMAs retired ======
    if (lane == 0) {
      uint32_t pair_it = 0;
      for (int64_t pl = blockIdx.x; pl < pair_count; pl += gridDim.x, ++pair_it) {
        mbar_wait_relaxed(&a_empty, (pair_it & 1) ^ 1);
        const uint32_t bytes = 2 * a_half;
        const uint8_t* src = aimg + (size_t)pl * bytes;
        mbar_arrive_expect_tx(&a_full, bytes);
        for (uint32_t o = 0; o < bytes; o += 16896) bulk_g2s(a_img + o, src + o, min(16896u, bytes - o), &a_full);
      }
    }
  } else if (warp == TC_EPI_WARPS + 1 || warp == TC_EPI_WARPS + 3) {
    // ================= MMA issuers (whole warp runs the uniform control flow, one elected lane issues) =========
    // One issuer warp per 128-row tile: a single warp sustains only one MMA per ~80 cycles through this loop (the
    // uniform-datapath descriptor arithmetic is on its critical path; tools/umma_bench.py), the tensor pipe takes an
    // N = 96 MMA every 56 - two independent issue streams keep it fed.  Both read the same B stage and document
    // image, each commits to the shared barriers (count = number of row tiles).
    const int mi = warp == TC_EPI_WARPS + 1 ? 0 : 1;
    if (mi < nmt) {
      const uint32_t issue = elect_one();
      const uint32_t idesc = idesc_bf16_f32(128, TC_NROWS);
      const uint32_t a0 = smem_u32(a_img), b0 = smem_u32(b_ring);
      const uint64_t ad_hi = smem_desc(a0, a_plane, 128), ad_lo = smem_desc(a0 + a_half, a_plane, 128);
      const uint32_t a_ks = (2 * a_plane) >> 4, b_ks = (2 * b_plane) >> 4, b_lo_off = b_half >> 4;
      const int nks = k.nkm;
      // tail k-steps: A = the tail plane at row shifts 2 s tpc (+ tpc for the second K chunk: LBO = tpc rows)
      const uint32_t at0 = a0 + (uint32_t)k.KCm * a_plane;
      const uint32_t bt_lo_off = (uint32_t)(2 * k.ntail) * b_plane >> 4;
      uint32_t it = 0, tile = 0, pair_it = 0;
      const long long tk0 = dbg ? clock64() : 0;
      for (int64_t pl = blockIdx.x; pl < pair_count; pl += gridDim.x, ++pair_it) {
        if (dbg && blockIdx.x == 0 && lane == 0 && mi == 0) dbg[7] = clock64() - tk0, dbg[8] = pair_it;
        { TC_T0(); mbar_wait(&a_full, pair_it & 1); TC_ACC(1); }
        tc_fence_after();
        for (int nt = 0; nt < ntiles; ++nt, ++tile) {
          const int as = tile & 1;
          { TC_T0(); mbar_wait(&acc_empty[as], ((tile >> 1) & 1) ^ 1); TC_ACC(2); }
          tc_fence_after();
          for (int bt = 0; bt < nit; ++bt, ++it) {
            const int s = it % nstages;
            { TC_T0(); mbar_wait(&full_b[s], (it / nstages) & 1); TC_ACC(3); }
            tc_fence_after();
            const uint64_t bd_hi = smem_desc(b0 + (uint32_t)s * stage, b_plane, 128);
            const uint32_t accf = bt != 0;
            if (bt < 7) {
              // 32-bit arithmetic on the descriptor low words (the high words are loop-invariant)
              const uint32_t dhi = (uint32_t)(ad_hi >> 32), bhi = (uint32_t)(bd_hi >> 32);
              const uint32_t bl0 = (uint32_t)bd_hi;
              {
                const int mt = mi;
                {
                  const uint32_t tacc = tbase + (uint32_t)(as * 2 + mt) * TC_NROWS;
                  const uint32_t ah0 = (uint32_t)ad_hi + (uint32_t)(mt * 128 + bt), al0 = (uint32_t)ad_lo + (uint32_t)(mt * 128 + bt);
#pragma unroll 4
                  for (int ks = 0; ks < nks; ++ks) {
                    const uint32_t ahk = ah0 + (uint32_t)ks * a_ks, alk = al0 + (uint32_t)ks * a_ks;
                    const uint32_t bhk = bl0 + (uint32_t)ks * b_ks, blk = bhk + b_lo_off;
                    mma_bf16_ss_w32(tacc, ahk, dhi, bhk, bhi, idesc, accf | (uint32_t)(ks != 0), issue);
                    mma_bf16_ss_w32(tacc, alk, dhi, bhk, bhi, idesc, 1, issue);
                    mma_bf16_ss_w32(tacc, ahk, dhi, blk, bhi, idesc, 1, issue);
                  }
                }
              }
            } else {
              {
                const int mt = mi;
                {
                  const uint32_t tacc = tbase + (uint32_t)(as * 2 + mt) * TC_NROWS;
                  for (int ts = 0; ts < k.ntail; ++ts) {
                    const uint32_t arow = at0 + (uint32_t)(mt * 128 + 2 * ts * k.tpc) * 16;
                    const uint64_t ahk = smem_desc(arow, (uint32_t)k.tpc * 16, 128);
                    const uint64_t alk = smem_desc(arow + a_half, (uint32_t)k.tpc * 16, 128);
                    const uint64_t bhk = bd_hi + (uint64_t)(ts * b_ks), blk = bhk + (uint64_t)bt_lo_off;
                    mma_bf16_ss_w(tacc, ahk, bhk, idesc, 1, issue);
                    mma_bf16_ss_w(tacc, alk, bhk, idesc, 1, issue);
                    mma_bf16_ss_w(tacc, ahk, blk, idesc, 1, issue);
                  }
                }
              }
            }
            mma_commit_w(&empty_b[s], issue);        // slab free once these MMAs retire
          }
          mma_commit_w(&acc_full[as], issue);        // accumulator stage complete
        }
        mma_commit_w(&a_empty, issue);               // document image free once this pair's MMAs retire
      }
    }
  } else {
    // ================= epilogue warpgroups (+ A staging) =================
    const int et = tid;                           // 0..TC_EPI_THREADS-1
    const int lane_base = (warp & 3) * 32;        // TMEM lane quarter this warp may read
    const int wq = warp >> 2;                     // this warp's index among the TC_EPI_WARPS/4 warps of its quarter
    const int nun = nmt * IPT;                    // work units (row tile mt, query position il) of one column tile
    uint32_t tile = 0;
    for (int64_t pl = blockIdx.x; pl < pair_count; pl += gridDim.x) {
      const int64_t pg = pair_begin + pl;
      const int64_t b = pg / N;
      TC_T0();
      // ---- token ids of the pair for the exact-match channel ----
      for (int i = et; i < TC_MAXRA + 8; i += TC_EPI_THREADS) {
        int j = i - 3;
        dids[i] = (j >= 0 && j < Ld) ? (int)d[pg * Ld + j] : -1;
      }
      for (int i = et; i < Lq; i += TC_EPI_THREADS) qids[i] = (int)q[b * Lq + i];
      named_bar_sync(1, TC_EPI_THREADS);          // dids/qids visible to all epilogue threads
      if (warp == 0) TC_ACC(5);

      float mx[MT_TC_MAXM];
      int ai[ARG ? MT_TC_MAXM : 1];
#pragma unroll
      for (int m = 0; m < MT_TC_MAXM; ++m) mx[m] = -INFINITY;
      if constexpr (ARG) {
#pragma unroll
        for (int m = 0; m < MT_TC_MAXM; ++m) ai[m] = 0x7fffffff;
      }

      for (int nt = 0; nt < ntiles; ++nt, ++tile) {
        const int as = tile & 1;
        { TC_T0(); mbar_wait_relaxed(&acc_full[as], (tile >> 1) & 1); if (warp == 0) TC_ACC(4); }
        tc_fence_after();
        const long long te0_ = dbg ? clock64() : 0;
        {
          // Work unit = (row tile mt, TC_NI query positions from il0); units are dealt round-robin to the warps of a quarter.
          constexpr int G = (IPT + TC_NI - 1) / TC_NI;
#pragma unroll 1
          for (int un = wq; un < nmt * G; un += TC_EPI_WARPS / 4) {
            if (dbg && dbg[15]) continue;   // timing experiment (tools/mt_timing.py): MMA stream without the epilogue math
            const int mt = un % nmt, il0 = TC_NI * (un / nmt);
            const int jrow = mt * 128 + lane_base + lane;   // doc position of this thread
            const uint32_t tacc = tbase + ((uint32_t)lane_base << 16) + (uint32_t)(as * 2 + mt) * TC_NROWS + il0 * FP;
            const int i0 = nt * IPT + il0;
            if (M <= 20 && TC_NI == 2 && il0 + 1 < IPT)
              mt_epi_unit<NF, TC_NI, ARG>(tacc, i0, Lq, jrow < Ld, dids + jrow, qids, ec, w1t, mx, ai, jrow, Ld);
            else if (M <= 20)
              mt_epi_unit<NF, 1, ARG>(tacc, i0, Lq, jrow < Ld, dids + jrow, qids, ec, w1t, mx, ai, jrow, Ld);
            else {
              for (int n = 0; n < TC_NI && il0 + n < IPT; ++n)
                mt_epi_unit_wide<NF, ARG>(tacc + n * FP, i0 + n, Lq, jrow < Ld, dids + jrow, qids, ec, w1t, mx, ai, jrow, Ld);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (dbg && blockIdx.x == 0 && warp == 0 && lane == 0) dbg[6] += clock64() - te0_;
        if (lane == 0) mbar_arrive(&acc_empty[as]);
      }
      // ---- max over all rows of the pair, Linear(M -> 1), one store ----
      int* redi = dids;   // (ARG) the ids of this pair are no longer needed; 8 warps x 32 slots fit in TC_MAXRA + 8 ints
      if constexpr (ARG) named_bar_sync(1, TC_EPI_THREADS);   // every warp is done with dids
#pragma unroll
      for (int m = 0; m < MT_TC_MAXM; ++m) {
        if constexpr (ARG) {
          float v = mx[m];
          int ix = ai[m];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
            if (ov > v || (ov == v && oi < ix)) v = ov, ix = oi;
          }
          if (lane == 0) red[warp * MT_TC_MAXM + m] = v, redi[warp * MT_TC_MAXM + m] = ix;
        } else {
          float v = warp_max(mx[m]);
          if (lane == 0) red[warp * MT_TC_MAXM + m] = v;
        }
      }
      named_bar_sync(1, TC_EPI_THREADS);
      if (warp == 0) {
        float v = 0.f;
        if (lane < M) {
          float best = red[lane];
          int bi = ARG ? redi[lane] : 0;
#pragma unroll
          for (int w = 1; w < TC_EPI_WARPS; ++w) {
            const float o = red[w * MT_TC_MAXM + lane];
            if constexpr (ARG) {
              const int oi = redi[w * MT_TC_MAXM + lane];
              if (o > best || (o == best && oi < bi)) best = o, bi = oi;
            } else {
              best = fmaxf(best, o);
            }
          }
          if constexpr (ARG) pooled[pl * M + lane] = best, argidx[pl * M + lane] = bi;
          v = best * p.wo[lane];
        }
        v = warp_sum(v);
        if (lane == 0) scores[pg] = v + p.wo[M];
      }
      named_bar_sync(1, TC_EPI_THREADS);          // red / dids / A image reusable
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, TC_TCOLS);
}

long long* g_mt_dbg = nullptr;  // device buffer of 16 counters when role timing is on (tools/mt_timing.py)

bool mt_tc_supported(const MtPack& p, int Lq, int Ld) {
  if (p.nf != 4 && p.nf != 6) return false;
  if (p.M > MT_TC_MAXM || Ld > 250 || Lq > 4096) return false;
  return tc_stages(p, Lq) >= 2;
}

static inline int tc_ra(int Ld) { return (((Ld + 127) / 128) * 128 + 6 + 7) & ~7; }

void mt_tc_workspace(const MtPack& p, int64_t nq, int64_t pc, int Lq, int Ld, size_t* timg_bytes, size_t* aimg_bytes) {
  const TcK k = tc_k(p.C);
  const int IPT = TC_NROWS / p.FP, ntiles = (Lq + IPT - 1) / IPT;
  *timg_bytes = (size_t)nq * ntiles * tc_timg_per_tile(k);
  *aimg_bytes = (size_t)pc * 2 * k.KA * tc_ra(Ld) * 16;
}

// Host copy of the epilogue weights (exact-match taps, bias, 1x1 conv) for the constant-bank kernel parameter.
int32_t mt_epi_const(const MtPack& p, MtEpiConst* out, cudaStream_t s) {
  memset(out, 0, sizeof(*out));
  if (p.FPP > 24 || p.M > MT_TC_MAXM) return CAIR_OK;  // tensor-core path unsupported anyway
  std::vector<float> wem((size_t)21 * p.FPP), bias(p.FPP), w1((size_t)p.M * p.FPP), b1(p.M);
  CAIR_CUDA(cudaMemcpyAsync(wem.data(), p.wem, wem.size() * 4, cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaMemcpyAsync(bias.data(), p.bias, bias.size() * 4, cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaMemcpyAsync(w1.data(), p.w1, w1.size() * 4, cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaMemcpyAsync(b1.data(), p.b1, b1.size() * 4, cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaStreamSynchronize(s));
  for (int t = 0; t < 21; ++t)
    for (int f = 0; f < p.FPP; ++f) {
      out->wem[t][f] = wem[(size_t)t * p.FPP + f];
    }
  for (int qz = 1; qz < 8; ++qz)
    for (int k = 1; k < 8; ++k)
      for (int f = 0; f < p.FPP; ++f) {
        float acc = 0.f;
        for (int a = 0; a < 3; ++a)
          if ((qz >> a) & 1)
            for (int bt = 0; bt < k; ++bt) acc += wem[(size_t)(a * 7 + bt) * p.FPP + f];
        out->padtab[qz][k][f] = acc;
      }
  for (int f = 0; f < p.FPP; ++f) out->bias[f] = bias[f];
  for (int m = 0; m < p.M; ++m) {
    out->b1[m] = b1[m];
    for (int f = 0; f < p.FPP; ++f) out->w1t[f][m] = w1[(size_t)m * p.FPP + f];
  }
  return CAIR_OK;
}

int32_t mt_tc_build_t(const MtPack& p, const float* cq, uint8_t* timg, int Lq, int64_t nq, cudaStream_t s) {
  if (nq <= 0) return CAIR_OK;
  const TcK k = tc_k(p.C);
  const int IPT = TC_NROWS / p.FP, ntiles = (Lq + IPT - 1) / IPT;
  const size_t smem = (size_t)(IPT + 2) * tc_cp(p.C) * sizeof(float);
  (void)k;
  CAIR_LAUNCH(mt_tc_build_t_kernel, dim3(ntiles, (unsigned)nq), BT_THREADS, smem, s, cq, p, Lq, IPT, ntiles, timg);
  return CAIR_OK;
}

int32_t mt_tc_doc_image(const MtPack& p, const float* cd, uint8_t* aimg, int Ld, int64_t pair_count, cudaStream_t s) {
  if (pair_count <= 0) return CAIR_OK;
  prof_mark("doc_image", s);
  CAIR_LAUNCH(mt_tc_image_kernel, (unsigned)pair_count, 256, 0, s, cd, p.C, Ld, tc_ra(Ld), pair_count, aimg);
  return CAIR_OK;
}

// Fused projection + image (see mt_tc_proj_image_kernel).  wd_img from mt_tc_pack_wd.
static inline int pj_kp(int Hd) { return (Hd + 15) & ~15; }
static inline size_t pj_smem(int C, int Hd) {
  return (size_t)2 * ((pj_kp(Hd) / 8) * PJ_APLANE + PJ_LOSKEW) + (size_t)pj_kp(Hd) * 4 * tc_cp(C) + (size_t)tc_cp(C) * 4;
}
bool mt_tc_proj_supported(int C, int Hd) {
  return Hd % 4 == 0 && tc_cp(C) <= 256 && pj_smem(C, Hd) <= 110 * 1024;  // two CTAs per SM
}
int32_t mt_tc_pack_wd(Owned& own, const float* wd, int C, int Hd, uint8_t** img, cudaStream_t s) {
  const int CP = tc_cp(C), KP = pj_kp(Hd);
  CAIR_CUDA(own.alloc(img, (size_t)2 * (KP / 8) * CP * 16));
  CAIR_LAUNCH(mt_tc_pack_wd_kernel, 8, 256, 0, s, wd, C, Hd, CP, KP, *img);
  return CAIR_OK;
}
int32_t mt_tc_proj_image(const MtPack& p, const float* enc_d, int Hd, const uint8_t* wd_img, const float* bd,
                         uint8_t* aimg, int Ld, int64_t pair_count, cudaStream_t s) {
  if (pair_count <= 0) return CAIR_OK;
  if ((uintptr_t)enc_d % 16) return fail(CAIR_ERR_BAD_ARG, "match_tensor: encoder output not 16-byte aligned");
  const int CP = tc_cp(p.C), KP = pj_kp(Hd), ntile = (Ld + 127) / 128;
  const int64_t nitems = pair_count * ntile;
  const size_t smem = pj_smem(p.C, Hd);
  uint32_t tcols = 32;
  while ((int)tcols < CP) tcols <<= 1;
  CAIR_CUDA(cudaFuncSetAttribute(mt_tc_proj_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)(nitems < 2 * kSMs ? nitems : 2 * kSMs);
  CAIR_LAUNCH(mt_tc_proj_image_kernel, grid, PJ_THREADS, smem, s, enc_d, Hd, KP, wd_img, bd, p.C, CP, Ld, tc_ra(Ld), ntile,
              nitems, tcols, aimg);
  return CAIR_OK;
}

int32_t mt_tc_interact(const MtPack& p, const MtEpiConst& ec, const uint8_t* timg, const uint8_t* aimg, const int64_t* q,
                       const int64_t* d, int N, int Lq, int Ld, int64_t pair_begin, int64_t pair_count, int64_t q_begin,
                       int64_t nq, float* scores, cudaStream_t s, int max_ctas, float* pooled, int* argidx) {
  (void)nq;
  if (pair_count <= 0) return CAIR_OK;
  const TcK k = tc_k(p.C);
  const int IPT = TC_NROWS / p.FP, ntiles = (Lq + IPT - 1) / IPT;
  const int nstages = tc_stages(p, Lq);
  const size_t smem = tc_a_bytes(k) + (size_t)nstages * tc_stage_bytes(k) + tc_misc_bytes(p, Lq);
  prof_mark("interact", s);
  unsigned grid = (unsigned)(pair_count < kSMs ? pair_count : kSMs);
  if (max_ctas > 0 && grid > (unsigned)max_ctas) grid = (unsigned)max_ctas;
  if (pooled && argidx) {
    // training forward (train.cu): pooled features and arg-max cells of every (pair, output channel); local pair indexing
    if (p.nf == 6) {
      CAIR_CUDA(cudaFuncSetAttribute(mt_tc_interact_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CAIR_LAUNCH((mt_tc_interact_kernel<6, true>), grid, TC_THREADS, smem, s, aimg, timg, p, ec, q, d, N, Lq, Ld, ntiles, nstages,
                  pair_begin, pair_count, q_begin, scores, (long long*)nullptr, pooled, argidx);
    } else {
      CAIR_CUDA(cudaFuncSetAttribute(mt_tc_interact_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CAIR_LAUNCH((mt_tc_interact_kernel<4, true>), grid, TC_THREADS, smem, s, aimg, timg, p, ec, q, d, N, Lq, Ld, ntiles, nstages,
                  pair_begin, pair_count, q_begin, scores, (long long*)nullptr, pooled, argidx);
    }
    return CAIR_OK;
  }
  if (p.nf == 6) {
    CAIR_CUDA(cudaFuncSetAttribute(mt_tc_interact_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAIR_LAUNCH(mt_tc_interact_kernel<6>, grid, TC_THREADS, smem, s, aimg, timg, p, ec, q, d, N, Lq, Ld, ntiles,
                nstages, pair_begin, pair_count, q_begin, scores, g_mt_dbg, (float*)nullptr, (int*)nullptr);
  } else {
    CAIR_CUDA(cudaFuncSetAttribute(mt_tc_interact_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAIR_LAUNCH(mt_tc_interact_kernel<4>, grid, TC_THREADS, smem, s, aimg, timg, p, ec, q, d, N, Lq, Ld, ntiles,
                nstages, pair_begin, pair_count, q_begin, scores, g_mt_dbg, (float*)nullptr, (int*)nullptr);
  }
  return CAIR_OK;
}

}  // namespace cair
