// Generic GEMM  C[M,N] = act(A[M,K] W[N,K]^T + bias[N])  on the 5th-gen tensor cores (tcgen05, bf16x3 split
// precision, fp32 accumulate in TMEM).  Same A-row providers as gemm.cu: dense rows, embedding rows gathered by
// token id (optionally a window of `win` consecutive tokens = the im2col of a valid Conv1d) or max-pooled over a
// time window on load - none of those tensors is ever materialised.
//
// One CTA computes a 128-row x NT-column tile (NT <= 256): the weights are pre-packed ONCE (gemm_tc_pack) as bf16
// hi/lo operand images per (column tile, 64-wide K chunk) and streamed with cp.async.bulk; the A chunk is loaded
// fp32 by 256 loader threads (two per row, 8 independent 128-bit loads in flight each), split into hi/lo bf16 and
// written in operand-image order; one elected lane issues 4 k-steps x 3 passes of tcgen05.mma per chunk; the
// epilogue reads TMEM (thread <-> row), transposes the tile through shared memory, adds bias, applies the
// activation and stores coalesced 128-byte row segments.
// Ring of 2-4 stages (as many as fit): the loads of the next chunks overlap the MMAs of chunk k.
#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int GT_BM = 128;        // rows per CTA
constexpr int GT_BK = 64;         // K chunk
constexpr int GT_MAXSTAGES = 4;   // ring depth: as many (A chunk + W chunk) stages as fit in shared memory, 2..4
constexpr int GT_LWARPS = 8;      // loader / epilogue warps: two threads per row (32 K elements each)
constexpr int GT_THREADS = (GT_LWARPS + 2) * 32;   // + W producer warp + MMA issuer warp
// A operand planes are padded by one 16-byte unit: with 16 lanes of a half warp writing the 8 planes of ONE row (coalesced
// loader below), an unpadded plane stride (2048 B = a multiple of 128 B) would put all of them on the same banks
constexpr uint32_t GT_APLANE = GT_BM * 16 + 16;
constexpr uint32_t GT_AIMG = (GT_BK / 8) * GT_APLANE;  // one (hi|lo) A chunk image: 16.1 KB

int g_gemm_impl = 1;  // 1: tcgen05 where the shape allows, 0: always the fp32 CUDA-core kernel
int g_gemm_dbg = 0;   // timing experiments of the persistent kernel (results are wrong): 1 no stores, 2 no W traffic, 4 short epilogue

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

// Re-packs changed weights into an image of the same shape (training: the parameters move every step).
int32_t gemm_tc_repack(const float* w, const GemmTcW& tw, cudaStream_t s) {
  if (!tw.img) return CAIR_OK;
  CAIR_LAUNCH(gemm_tc_pack_kernel, dim3(tw.nkc, tw.nct), 256, 0, s, w, tw.N, tw.K, tw.NT, tw.nkc, tw.img);
  return CAIR_OK;
}

__device__ __forceinline__ void gt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// smem: A ring [stages][hi|lo] | W ring [stages][hi|lo]
__global__ void __launch_bounds__(GT_THREADS, 1)
    gemm_tc_kernel(GemmA a, const uint8_t* __restrict__ wimg, const float* __restrict__ bias, float* __restrict__ c,
                   int64_t ldc, int64_t M, int N, int K, int NT, int nkc, int act, uint32_t tcols, int nst) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t a_full[GT_MAXSTAGES], w_full[GT_MAXSTAGES], empty[GT_MAXSTAGES], acc_full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * GT_BM;
  const int ct = blockIdx.y;
  const uint32_t w_plane = (uint32_t)NT * 16, w_half = (GT_BK / 8) * w_plane;
  uint8_t* a_ring = smraw;
  uint8_t* w_ring = a_ring + (size_t)nst * 2 * GT_AIMG;

  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == GT_LWARPS * 32) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&a_full[s], GT_LWARPS);   // one arrive per loader warp
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

  if (warp == GT_LWARPS) {
    // ---- W producer ----
    if (lane == 0) {
      const uint8_t* src = wimg + (size_t)ct * nkc * 2 * w_half;
      for (int kc = 0; kc < nkc; ++kc) {
        const int s = kc % nst;
        mbar_wait_relaxed(&empty[s], ((kc / nst) & 1) ^ 1);
        const uint32_t bytes = 2 * w_half;
        mbar_arrive_expect_tx(&w_full[s], bytes);
        uint8_t* dst = w_ring + (size_t)s * bytes;
        for (uint32_t o = 0; o < bytes; o += 32768) bulk_g2s(dst + o, src + (size_t)kc * bytes + o, min(32768u, bytes - o), &w_full[s]);
      }
    }
  } else if (warp == GT_LWARPS + 1) {
    // ---- MMA issuer ----
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, NT);
    const uint64_t ad0 = smem_desc(smem_u32(a_ring), GT_APLANE, 128);
    const uint64_t wd0 = smem_desc(smem_u32(w_ring), w_plane, 128);
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc % nst;
      const uint32_t ph = (kc / nst) & 1;
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
    // ---- A loaders, then epilogue ----
    // Coalesced: lanes 0-15 of a warp read the 16 float4 slices (256 contiguous bytes) of ONE row of the 64-wide K chunk,
    // lanes 16-31 those of the next row, so a load instruction touches 4 cache lines instead of 32 (one 16-byte slice
    // from each of 32 different rows: the L1 wavefront rate, 2048 per chunk, used to set the pace).  Warp w owns rows
    // 16 w .. 16 w + 15 of the tile, iteration i -> row 16 w + 2 i + lane / 16; 8 independent loads in flight per thread.
    const int j = lane & 15, rsub = lane >> 4;
    constexpr int NV = 8;
    int64_t gseq[NV];
    int gt0[NV];
    bool rv[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int64_t r = m0 + warp * 16 + 2 * i + rsub;
      rv[i] = r < M;
      gseq[i] = a.table ? (rv[i] ? r / a.T : 0) : r;
      gt0[i] = a.table ? (int)((rv[i] ? r : 0) - gseq[i] * a.T) - a.pad : 0;
    }
    bool bad_id = false;
    // Software pipeline over the K chunks: the token ids of chunk kc+2 are requested and the row slices of chunk kc+1
    // loaded BEFORE chunk kc is converted and written, so the two dependent global-memory latencies of a gathered chunk
    // (id -> row) overlap the conversion / barrier wait of the previous chunks.
    int64_t idraw[NV];      // raw token ids of the chunk after next (consumed one iteration after they were requested)
    const float* src[NV];   // row slices of the next chunk
    float4 vn[NV];          // row slices of the next chunk to convert
    // (segment, offset inside the embedding row) of this thread's slice, advanced by 64 columns per chunk (no divisions)
    const int E_ = a.table ? a.E : 1;
    int seg_c = a.table ? (4 * j) / E_ : 0, rem_c = a.table ? (4 * j) % E_ : 0;   // chunk whose ids are requested next
    int seg_p = seg_c, rem_p = rem_c;                                              // chunk whose pointers are formed next
    auto advance = [&](int& seg, int& rem) {
      rem += GT_BK;
      while (rem >= E_) rem -= E_, ++seg;
    };
    auto request_ids = [&](int kc) {   // independent loads, nothing consumes them in this iteration
      const int kk = kc * GT_BK + 4 * j;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int pos = gt0[i] + seg_c;
        const bool ok = rv[i] && kk < K && pos >= 0 && pos < a.L;
        idraw[i] = a.ids[ok ? gseq[i] * a.L + pos : 0];
      }
      advance(seg_c, rem_c);
    };
    auto form_ptrs = [&](int kc) {     // ids of chunk kc (requested an iteration ago) -> source pointers
      const int kk = kc * GT_BK + 4 * j;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int pos = gt0[i] + seg_p;
        const bool ok = rv[i] && kk < K && pos >= 0 && pos < a.L;
        int64_t id = idraw[i];
        const bool inr = id >= 0 && id < a.V;
        bad_id |= ok && !inr;
        id = inr ? id : 0;
        src[i] = ok ? a.table + id * a.E + rem_p : nullptr;
      }
      advance(seg_p, rem_p);
    };
    auto load_rows = [&](int kc) {
      const int kk = kc * GT_BK + 4 * j;
      if (a.table) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          vn[i] = src[i] ? *reinterpret_cast<const float4*>(src[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          vn[i] = (rv[i] && kk < K) ? gemm_a_load4(a, gseq[i], kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (a.table) {
      request_ids(0);
      form_ptrs(0);
    }
    load_rows(0);
    if (a.table && nkc > 1) request_ids(1);
    // operand image position of this thread's 8-byte half unit: plane j/2, row, bytes (j & 1) * 8
    const uint32_t aoff = (uint32_t)(j >> 1) * GT_APLANE + (uint32_t)(warp * 16 + rsub) * 16 + (uint32_t)(j & 1) * 8;
    for (int kc = 0; kc < nkc; ++kc) {
      const int s = kc % nst;
      float4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = vn[i];
      if (kc + 1 < nkc) {
        if (a.table) form_ptrs(kc + 1);
        load_rows(kc + 1);
      }
      if (a.table && kc + 2 < nkc) request_ids(kc + 2);
      mbar_wait_relaxed(&empty[s], ((kc / nst) & 1) ^ 1);
      uint8_t* ah = a_ring + (size_t)s * 2 * GT_AIMG + aoff;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v[i].x, v[i].y, h0, l0);
        split_bf16x2(v[i].z, v[i].w, h1, l1);
        *reinterpret_cast<uint2*>(ah + (size_t)i * 32) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(ah + GT_AIMG + (size_t)i * 32) = make_uint2(l0, l1);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) gt_arrive(&a_full[s]);
    }
    if (bad_id && a.err) atomicOr(a.err, ERRF_BAD_TOKEN);
    // ---- epilogue ----
    // TMEM hands every thread 32 consecutive columns of ITS row; storing those directly makes each store instruction
    // touch 32 different rows (32 sectors).  The tile goes through shared memory instead (the A ring is free once all
    // MMAs have retired; row stride 33 floats = conflict-free both ways, and a warp only re-reads the 32 rows its own
    // lanes wrote, so a __syncwarp suffices): every store instruction then writes one 128-byte row segment.
    mbar_wait_relaxed(&acc_full, 0);
    tc_fence_after();
    const int n0 = ct * NT;
    const int qt = warp & 3;   // TMEM lane quarter of this warp; warps 0-3 / 4-7 take alternate 32-column chunks
    float* stage = reinterpret_cast<float*>(a_ring) + (size_t)warp * 32 * 33;
    for (int c0 = (warp >> 2) * 32; c0 < NT; c0 += 64) {
      float v[32];
      tmem_ld32(tbase + ((uint32_t)(qt * 32) << 16) + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j];
      __syncwarp();
      const int col = n0 + c0 + lane;
      const bool cvalid = c0 + lane < NT && col < N;
      const float bcol = (bias && cvalid) ? bias[col] : 0.f;
      const int64_t rbase = m0 + qt * 32;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) {
        float x = stage[rr * 33 + lane] + bcol;
        if (act == ACT_TANH) x = tanhf(x);
        if (act == ACT_RELU) x = fmaxf(x, 0.f);
        if (cvalid && rbase + rr < M) c[(rbase + rr) * ldc + col] = x;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, tcols);
}


// ---- persistent, fully pipelined version (default) -------------------------------------------------------------
// The kernel above runs one tile per CTA and one CTA per SM: while a tile is in its epilogue (128 x NT fp32 through shared
// memory to HBM) nothing is being loaded, and every tile pays the id -> row latency chain of its first chunk again.  Here a
// CTA walks over tiles (column tile fastest, so the row tile's A rows are re-read from L2 by neighbouring SMs at the same
// time): the loader warps run ONE software pipeline over the flattened (tile, K chunk) sequence, the accumulator is double
// buffered in TMEM, and four dedicated warps drain accumulator k while the MMAs of tile k+1 run.
constexpr int G2_EWARPS = 4;
// A operand image for weights with several column tiles (N > 256, or N = 300 cut in two): the persistent kernel would gather and
// convert the same A rows once per column tile; instead one pass writes every (row tile, K chunk) stage in the exact shared-memory
// layout (hi | lo images with the padded planes), and the GEMM's loader warps become plain cp.async.bulk producers.
size_t gemm_tc_aimg_bytes(int64_t M, int K) {
  return (size_t)((M + GT_BM - 1) / GT_BM) * (size_t)((K + GT_BK - 1) / GT_BK) * 2 * GT_AIMG;
}
__global__ void __launch_bounds__(256) gemm_tc_aimg_kernel(GemmA a, int64_t M, int K, int nkc, uint8_t* __restrict__ img) {
  const int kc = blockIdx.x, tid = threadIdx.x;
  const int64_t mt = blockIdx.y;
  const int row = tid >> 1, half = tid & 1;
  const int64_t r = mt * GT_BM + row;
  uint8_t* blk = img + ((size_t)mt * nkc + kc) * 2 * GT_AIMG;
#pragma unroll
  for (int u = 0; u < 4; ++u) {   // 16-byte unit u of this thread = K elements [8 (4 half + u), +8) of the chunk = plane 4 half + u
    const int pl = 4 * half + u;
    const int kk = kc * GT_BK + 8 * pl;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (r < M && kk < K) v0 = gemm_a_load4(a, r, kk);
    if (r < M && kk + 4 < K) v1 = gemm_a_load4(a, r, kk + 4);
    uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
    split_bf16x2(v0.x, v0.y, h0, l0);
    split_bf16x2(v0.z, v0.w, h1, l1);
    split_bf16x2(v1.x, v1.y, h2, l2);
    split_bf16x2(v1.z, v1.w, h3, l3);
    const size_t off = (size_t)pl * GT_APLANE + (size_t)row * 16;
    *reinterpret_cast<uint4*>(blk + off) = make_uint4(h0, h1, h2, h3);
    *reinterpret_cast<uint4*>(blk + GT_AIMG + off) = make_uint4(l0, l1, l2, l3);
  }
  // the 16 padding bytes at the end of every plane are never read by the MMA (rows 0..127 only)
}

constexpr int G2_THREADS = (GT_LWARPS + G2_EWARPS + 2) * 32;   // 8 loaders, 4 epilogue, W producer, MMA issuer
constexpr int G2_ESTAGE = 32 * 33;                               // floats per epilogue warp

__global__ void __launch_bounds__(G2_THREADS, 1)
    gemm_tc2_kernel(GemmA a, const uint8_t* __restrict__ wimg, const float* __restrict__ bias, float* __restrict__ c,
                    int64_t ldc, int64_t M, int N, int K, int NT, int nkc, int nct, int act, uint32_t tcols1, int nst,
                    int ntiles, int dbg, const float* __restrict__ dot_w, const float* __restrict__ dot_b,
                    float* __restrict__ dot_out, const uint8_t* __restrict__ aimg) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t a_full[GT_MAXSTAGES], w_full[GT_MAXSTAGES], empty[GT_MAXSTAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t w_plane = (uint32_t)NT * 16, w_half = (GT_BK / 8) * w_plane;
  uint8_t* a_ring = smraw;
  uint8_t* w_ring = a_ring + (size_t)nst * 2 * GT_AIMG;
  float* estage = reinterpret_cast<float*>(w_ring + (size_t)nst * 2 * w_half);
  const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) tmem_alloc(&tmem_slot, 2 * tcols1);
  if (tid == GT_LWARPS * 32) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(&a_full[s], GT_LWARPS);
      mbar_init(&w_full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], G2_EWARPS);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;

  if (warp == GT_LWARPS + G2_EWARPS) {
    // ---- W producer ----
    if (lane == 0) {
      int g = 0;
      for (int kt = 0; kt < my_tiles; ++kt) {
        const int t = (int)blockIdx.x + kt * (int)gridDim.x, ct = t % nct;
        const uint8_t* src = wimg + (size_t)ct * nkc * 2 * w_half;
        for (int kc = 0; kc < nkc; ++kc, ++g) {
          const int s = g % nst;
          mbar_wait_relaxed(&empty[s], ((g / nst) & 1) ^ 1);
          const uint32_t bytes = 2 * w_half;
          if ((dbg & 2) && g >= nst) {   // timing experiment: no W traffic (stale operands)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&w_full[s])) : "memory");
            continue;
          }
          mbar_arrive_expect_tx(&w_full[s], bytes);
          uint8_t* dst = w_ring + (size_t)s * bytes;
          for (uint32_t o = 0; o < bytes; o += 32768) bulk_g2s(dst + o, src + (size_t)kc * bytes + o, min(32768u, bytes - o), &w_full[s]);
        }
      }
    }
  } else if (warp == GT_LWARPS + G2_EWARPS + 1) {
    // ---- MMA issuer ----
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, NT);
    const uint64_t ad0 = smem_desc(smem_u32(a_ring), GT_APLANE, 128);
    const uint64_t wd0 = smem_desc(smem_u32(w_ring), w_plane, 128);
    int g = 0;
    for (int kt = 0; kt < my_tiles; ++kt) {
      const int buf = kt & 1;
      mbar_wait(&acc_empty[buf], ((kt >> 1) & 1) ^ 1);   // the epilogue of tile kt-2 has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tbase + (uint32_t)buf * tcols1;
      for (int kc = 0; kc < nkc; ++kc, ++g) {
        const int s = g % nst;
        const uint32_t ph = (g / nst) & 1;
        mbar_wait(&a_full[s], ph);
        mbar_wait(&w_full[s], ph);
        tc_fence_after();
        const uint64_t ah = ad0 + (uint64_t)((uint32_t)s * 2 * GT_AIMG >> 4), al = ah + (uint64_t)(GT_AIMG >> 4);
        const uint64_t wh = wd0 + (uint64_t)((uint32_t)s * 2 * w_half >> 4), wl = wh + (uint64_t)(w_half >> 4);
#pragma unroll
        for (int ks = 0; ks < GT_BK / 16; ++ks) {
          const uint64_t ao = (uint64_t)(ks * ((2 * GT_APLANE) >> 4)), wo = (uint64_t)(ks * ((2 * w_plane) >> 4));
          mma_bf16_ss_w(tacc, ah + ao, wh + wo, idesc, (uint32_t)((kc | ks) != 0), issue);
          mma_bf16_ss_w(tacc, al + ao, wh + wo, idesc, 1, issue);
          mma_bf16_ss_w(tacc, ah + ao, wl + wo, idesc, 1, issue);
        }
        mma_commit_w(&empty[s], issue);
      }
      mma_commit_w(&acc_full[buf], issue);
    }
  } else if (warp >= GT_LWARPS) {
    // ---- epilogue warps: TMEM -> shared-memory transpose -> coalesced 128-byte row segments ----
    const int qt = warp & 3;   // TMEM lane quarter (warps 8..11 -> 0..3)
    float* stage = estage + (size_t)(warp - GT_LWARPS) * G2_ESTAGE;
    const bool vec_ok = (ldc & 3) == 0 && ((uintptr_t)c & 15) == 0 && (NT & 3) == 0 && (!bias || ((uintptr_t)bias & 15) == 0);
    for (int kt = 0; kt < my_tiles; ++kt) {
      const int buf = kt & 1;
      const int t = (int)blockIdx.x + kt * (int)gridDim.x, mt = t / nct, ct = t - mt * nct;
      const int64_t rbase = (int64_t)mt * GT_BM + qt * 32;
      const int n0 = ct * NT;
      mbar_wait_relaxed(&acc_full[buf], (kt >> 1) & 1);
      tc_fence_after();
      if (dot_out) {
        // row-dot epilogue (one column tile): out[r] = dot_w . act(row r + bias) + dot_b; the [M, N] product is never stored
        float dot = 0.f;
        for (int c0 = 0; c0 < NT; c0 += 32) {
          float v[32];
          tmem_ld32(tbase + ((uint32_t)(qt * 32) << 16) + (uint32_t)buf * tcols1 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int col = c0 + jj;
            if (col < N) {   // uniform
              float x = v[jj] + (bias ? __ldg(bias + col) : 0.f);
              if (act == ACT_TANH) x = tanhf(x);
              if (act == ACT_RELU) x = fmaxf(x, 0.f);
              dot = fmaf(__ldg(dot_w + col), x, dot);
            }
          }
        }
        if (rbase + lane < M) dot_out[rbase + lane] = dot + (dot_b ? __ldg(dot_b) : 0.f);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) gt_arrive(&acc_empty[buf]);
        continue;
      }
      for (int c0 = 0; c0 < NT; c0 += 32) {
        float v[32];
        tmem_ld32(tbase + ((uint32_t)(qt * 32) << 16) + (uint32_t)buf * tcols1 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) stage[lane * 33 + jj] = v[jj];
        __syncwarp();
        if (vec_ok && n0 + c0 + 32 <= N && c0 + 32 <= NT) {
          // four rows x 128 bytes per store instruction: lanes 8 g .. 8 g + 7 write row 4 k + g as float4 (a quarter of the
          // store instructions of the one-row-per-instruction form; the stride-33 staging stays conflict-free)
          const int g4 = lane >> 3, cq = (lane & 7) * 4;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) b4 = *reinterpret_cast<const float4*>(bias + n0 + c0 + cq);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = 4 * k + g4;
            const float* sp = stage + rr * 33 + cq;
            float4 x = make_float4(sp[0] + b4.x, sp[1] + b4.y, sp[2] + b4.z, sp[3] + b4.w);
            if (act == ACT_TANH) x = make_float4(tanhf(x.x), tanhf(x.y), tanhf(x.z), tanhf(x.w));
            if (act == ACT_RELU) x = make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
            if (rbase + rr < M && !(dbg & 1)) *reinterpret_cast<float4*>(c + (rbase + rr) * ldc + n0 + c0 + cq) = x;
          }
        } else {
          const int col = n0 + c0 + lane;
          const bool cvalid = c0 + lane < NT && col < N;
          const float bcol = (bias && cvalid) ? bias[col] : 0.f;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            float x = stage[rr * 33 + lane] + bcol;
            if (act == ACT_TANH) x = tanhf(x);
            if (act == ACT_RELU) x = fmaxf(x, 0.f);
            if (cvalid && rbase + rr < M && !(dbg & 1)) c[(rbase + rr) * ldc + col] = x;   // dbg 1: timing experiment without the stores
          }
        }
        __syncwarp();
        if (dbg & 4) break;   // timing experiment: drain one chunk only
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) gt_arrive(&acc_empty[buf]);
    }
  } else if (aimg) {
    // ---- A producers: the stages were written by gemm_tc_aimg_kernel; each loader warp bulk-copies its eighth of a stage ----
    if (lane == 0) {
      constexpr uint32_t PART = 2 * GT_AIMG / GT_LWARPS;
      static_assert(PART % 16 == 0 && PART * GT_LWARPS == 2 * GT_AIMG, "stage must split into 16-byte multiples");
      int g = 0;
      for (int kt = 0; kt < my_tiles; ++kt) {
        const int t = (int)blockIdx.x + kt * (int)gridDim.x;
        const uint8_t* src = aimg + (size_t)(t / nct) * nkc * 2 * GT_AIMG + (size_t)warp * PART;
        for (int kc = 0; kc < nkc; ++kc, ++g) {
          const int s = g % nst;
          mbar_wait_relaxed(&empty[s], ((g / nst) & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[s], PART);
          bulk_g2s(a_ring + (size_t)s * 2 * GT_AIMG + (size_t)warp * PART, src + (size_t)kc * 2 * GT_AIMG, PART, &a_full[s]);
        }
      }
    }
  } else {
    // ---- A loaders: one software pipeline over the flattened (tile, chunk) sequence ----
    // ids of chunk g+2 requested, rows of chunk g+1 loaded, chunk g converted and written (see the kernel above for the
    // lane <-> (row, 16-byte slice) mapping).  Each of the three stages has its own (tile, chunk) cursor, so the
    // pipeline does not drain at tile boundaries.
    const int j = lane & 15, rsub = lane >> 4;
    constexpr int NV = 8;
    const int total = my_tiles * nkc;
    const bool gathered = a.table != nullptr;
    const int E_ = gathered ? a.E : 1;
    const int seg0 = gathered ? (4 * j) / E_ : 0, rem0 = gathered ? (4 * j) % E_ : 0;
    // row set of the tile the FIRST pipeline stage is in (ids stage when gathered, load stage when dense)
    int rs_seq[NV], rs_t0[NV];
    unsigned rs_valid = 0;
    auto set_rows = [&](int kt) {
      const int t = (int)blockIdx.x + kt * (int)gridDim.x;
      const int64_t m0 = (int64_t)(t / nct) * GT_BM;
      rs_valid = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int64_t r = m0 + warp * 16 + 2 * i + rsub;
        const bool v = r < M;
        rs_valid |= (v ? 1u : 0u) << i;
        if (gathered) {
          const int64_t sq = v ? r / a.T : 0;
          rs_seq[i] = (int)sq;
          rs_t0[i] = (int)((v ? r : 0) - sq * a.T) - a.pad;
        } else {
          rs_seq[i] = (int)r;   // dense providers index rows directly (M < 2^31)
          rs_t0[i] = 0;
        }
      }
    };
    bool bad_id = false;
    int64_t idraw[NV];
    unsigned ok_ids = 0, ok_ptr = 0;   // per-slice validity of the chunk in the ids stage / handed on to the pointer stage
    int src[NV];   // element offset of the slice in the table (-1: zero slice); V * E < 2^31 is checked by the launcher
    float4 vn[NV];
    int c_kt = 0, c_kc = 0, c_seg = seg0, c_rem = rem0;   // ids stage cursor (gathered)
    int p_kc = 0, p_rem = rem0, p_kt = 0;                 // pointer / load stage cursor
    int rows_kt = -1;
    auto request_ids = [&]() {
      if (rows_kt != c_kt) set_rows(c_kt), rows_kt = c_kt;
      const int kk = c_kc * GT_BK + 4 * j;
      ok_ids = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int pos = rs_t0[i] + c_seg;
        const bool ok = ((rs_valid >> i) & 1) && kk < K && pos >= 0 && pos < a.L;
        ok_ids |= (ok ? 1u : 0u) << i;
        idraw[i] = a.ids[ok ? (int64_t)rs_seq[i] * a.L + pos : 0];
      }
      if (++c_kc == nkc) {
        c_kc = 0, ++c_kt, c_seg = seg0, c_rem = rem0;
      } else {
        c_rem += GT_BK;
        while (c_rem >= E_) c_rem -= E_, ++c_seg;
      }
    };
    auto form_ptrs = [&]() {   // ids requested one call of request_ids ago -> source pointers
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const bool ok = (ok_ptr >> i) & 1;
        int64_t id = idraw[i];
        const bool inr = id >= 0 && id < a.V;
        bad_id |= ok && !inr;
        id = inr ? id : 0;
        src[i] = ok ? (int)id * a.E + p_rem : -1;
      }
    };
    auto load_rows = [&]() {
      if (gathered) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          vn[i] = src[i] >= 0 ? *reinterpret_cast<const float4*>(a.table + src[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        if (rows_kt != p_kt) set_rows(p_kt), rows_kt = p_kt;
        const int kk = p_kc * GT_BK + 4 * j;
#pragma unroll
        for (int i = 0; i < NV; ++i)
          vn[i] = (((rs_valid >> i) & 1) && kk < K) ? gemm_a_load4(a, (int64_t)rs_seq[i], kk) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++p_kc == nkc) {
        p_kc = 0, ++p_kt, p_rem = rem0;
      } else if (gathered) {
        p_rem += GT_BK;
        while (p_rem >= E_) p_rem -= E_;
      }
    };
    if (total > 0) {
      if (gathered) {
        request_ids();
        ok_ptr = ok_ids;
        form_ptrs();
      }
      load_rows();
      if (gathered && total > 1) request_ids();
    }
    const uint32_t aoff = (uint32_t)(j >> 1) * GT_APLANE + (uint32_t)(warp * 16 + rsub) * 16 + (uint32_t)(j & 1) * 8;
    for (int g = 0; g < total; ++g) {
      const int s = g % nst;
      float4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = vn[i];
      if (g + 1 < total) {
        if (gathered) {
          ok_ptr = ok_ids;
          form_ptrs();
        }
        load_rows();
      }
      if (gathered && g + 2 < total) request_ids();
      mbar_wait_relaxed(&empty[s], ((g / nst) & 1) ^ 1);
      uint8_t* ah = a_ring + (size_t)s * 2 * GT_AIMG + aoff;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v[i].x, v[i].y, h0, l0);
        split_bf16x2(v[i].z, v[i].w, h1, l1);
        *reinterpret_cast<uint2*>(ah + (size_t)i * 32) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(ah + GT_AIMG + (size_t)i * 32) = make_uint2(l0, l1);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) gt_arrive(&a_full[s]);
    }
    if (bad_id && a.err) atomicOr(a.err, ERRF_BAD_TOKEN);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 2 * tcols1);
}

bool gemm_tc_usable(const GemmA& a, int K) {
  if (!g_gemm_impl) return false;
  if (K % 4) return false;
  if (a.table) return (a.E % 4 == 0) && ((uintptr_t)a.table % 16 == 0);
  if (a.dwin) return (a.E % 4 == 0) && (a.lda % 4 == 0) && ((uintptr_t)a.dense % 16 == 0);
  return (a.lda % 4 == 0) && ((uintptr_t)a.dense % 16 == 0);
}

static bool gemm_tc2_shape_ok(const GemmA& a, int64_t M) {
  return g_gemm_impl == 1 && M < ((int64_t)1 << 31) && (!a.table || (int64_t)a.V * a.E < ((int64_t)1 << 31));
}
// out[r] = dot_w . act(A[r] W^T + bias) + dot_b without storing the [M, N] product (N <= 256: one column tile)
bool gemm_tc_rowdot_usable(const GemmA& a, const GemmTcW& w, int64_t M) {
  return w.img && w.nct == 1 && M >= 128 && gemm_tc_usable(a, w.K) && gemm_tc2_shape_ok(a, M);
}
static int32_t gemm_tc2_launch(const GemmA& a, const GemmTcW& w, const float* bias, float* c, int64_t ldc, int64_t M, Act act,
                               const float* dot_w, const float* dot_b, float* dot_out, cudaStream_t s, bool* done,
                               uint8_t* aimg_scratch = nullptr) {
  *done = false;
  const size_t stage_bytes = (size_t)2 * GT_AIMG + (size_t)2 * (GT_BK / 8) * w.NT * 16;
  const size_t ebytes = (size_t)G2_EWARPS * G2_ESTAGE * sizeof(float);
  int nst = (int)((220 * 1024 - ebytes) / stage_bytes);
  nst = nst > GT_MAXSTAGES ? GT_MAXSTAGES : nst;
  if (nst < 2) return CAIR_OK;
  uint32_t tcols1 = 32;
  while ((int)tcols1 < ((w.NT + 31) & ~31)) tcols1 <<= 1;
  const int64_t ntiles = ((M + GT_BM - 1) / GT_BM) * w.nct;
  if (ntiles >= ((int64_t)1 << 31)) return CAIR_OK;
  const size_t smem = (size_t)nst * stage_bytes + ebytes;
  CAIR_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)(ntiles < kSMs ? ntiles : kSMs);
  const uint8_t* aimg = nullptr;
  // several column tiles share the A rows: convert them once - unless the provider is a window of several tokens, whose image
  // would be `win` times the size of the rows it is made of (DUET conv_d1 at N = 500: 11.5 ms with the image, 9.1 ms without)
  if (aimg_scratch && (w.nct >= 2 || dot_out) && !((a.table || a.dwin) && a.win > 1) && !(g_gemm_dbg & 8)) {
    CAIR_LAUNCH(gemm_tc_aimg_kernel, dim3((unsigned)w.nkc, (unsigned)((M + GT_BM - 1) / GT_BM)), 256, 0, s, a, M, w.K, w.nkc, aimg_scratch);
    aimg = aimg_scratch;
  }
  CAIR_LAUNCH(gemm_tc2_kernel, grid, G2_THREADS, smem, s, a, w.img, bias, c, ldc, M, w.N, w.K, w.NT, w.nkc, w.nct, (int)act, tcols1, nst,
              (int)ntiles, g_gemm_dbg, dot_w, dot_b, dot_out, aimg);
  *done = true;
  return CAIR_OK;
}
int32_t gemm_tc_rowdot(const GemmA& a, const GemmTcW& w, const float* bias, Act act, const float* dot_w, const float* dot_b,
                       float* out, int64_t M, cudaStream_t s, uint8_t* aimg_scratch) {
  if ((a.table || a.dwin) && w.K != a.win * a.E) return fail(CAIR_ERR_BAD_ARG, "gemm_tc: K != win*E");
  bool done = false;
  CAIR_TRY(gemm_tc2_launch(a, w, bias, nullptr, 0, M, act, dot_w, dot_b, out, s, &done, aimg_scratch));
  return done ? CAIR_OK : fail(CAIR_ERR_UNSUPPORTED, "gemm_tc_rowdot: shape not supported (check gemm_tc_rowdot_usable)");
}

int32_t gemm_tc(const GemmA& a, const GemmTcW& w, const float* bias, float* c, int64_t ldc, int64_t M, Act act,
                cudaStream_t s, uint8_t* aimg_scratch) {
  if (M <= 0) return CAIR_OK;
  if ((a.table || a.dwin) && w.K != a.win * a.E) return fail(CAIR_ERR_BAD_ARG, "gemm_tc: K != win*E");
  if (gemm_tc2_shape_ok(a, M)) {
    bool done = false;
    CAIR_TRY(gemm_tc2_launch(a, w, bias, c, ldc, M, act, nullptr, nullptr, nullptr, s, &done, aimg_scratch));
    if (done) return CAIR_OK;
  }
  const size_t stage_bytes = (size_t)2 * GT_AIMG + (size_t)2 * (GT_BK / 8) * w.NT * 16;
  int nst = (int)((220 * 1024) / stage_bytes);
  nst = nst > GT_MAXSTAGES ? GT_MAXSTAGES : nst;
  if (nst > w.nkc) nst = w.nkc < 2 ? 2 : w.nkc;
  size_t smem = (size_t)nst * stage_bytes;
  if (smem < (size_t)GT_LWARPS * 32 * 33 * 4 + 1024) smem = (size_t)GT_LWARPS * 32 * 33 * 4 + 1024;   // epilogue staging lives in the A ring
  uint32_t tcols = 32;
  while ((int)tcols < ((w.NT + 31) & ~31)) tcols <<= 1;
  CAIR_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((M + GT_BM - 1) / GT_BM), (unsigned)w.nct);
  CAIR_LAUNCH(gemm_tc_kernel, grid, GT_THREADS, smem, s, a, w.img, bias, c, ldc, M, w.N, w.K, w.NT, w.nkc, (int)act, tcols, nst);
  return CAIR_OK;
}

}  // namespace cair
