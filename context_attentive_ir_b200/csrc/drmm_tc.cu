// DRMM (neuroir/rankers/drmm.py:29-84, GatingNetwork :87-98) with the 20 x 200 cosines of a pair on the tensor cores
// and ONE pass over HBM.
//
// The fp32 kernels (drmm.cu) read every document row twice (norm pass + L2 re-stream) and spend their time issuing
// fp32 FMAs.  Here one CTA per pair gathers the document rows once (128 tokens per tile, 64-wide K chunks, coalesced
// 256-byte row segments, software-pipelined), converts them to bf16 hi/lo operand images on the fly while accumulating
// the row norms, and the cosines are  D[token, qrow] = d_raw[token, :] . qn[qrow, :]  on tcgen05 (bf16x3 split precision,
// fp32 accumulate in TMEM; qn = the query rows normalised exactly as the fp32 kernels do), scaled by 1 / |d| in the epilogue.
// The bins of numpy.histogram are discontinuous, so a cosine that lands within DT_TOL of a bin edge (-1, -.5, 0, .5, 1)
// is NOT trusted: that cell is recomputed by a warp with the fp32 kernels' arithmetic - the same normalisation (lane-strided
// partial sums + butterfly) and the same sequential fp32 FMA chain over k - so every histogram is identical to
// drmm_kernel's.  Zero rows (PAD / OOV tokens) give an exact 0 in the reference (x / max(|x|, eps) = 0) and go to bin 2
// without a recompute.  About 1-2 cells of 4000 per pair are near an edge for random embeddings; exact-match cells
// (cos ~ 1) always are.
#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int DT_BM = 128;                      // document tokens per tile (MMA M)
constexpr int DT_BK = 64;                       // K chunk
constexpr int DT_LW = 8;                        // loader / epilogue warps
constexpr int DT_THREADS = (DT_LW + 1) * 32;    // + MMA issuer warp
constexpr int DT_NQ = 32;                       // query rows padded to the MMA N
constexpr int DT_MAXLD = 256;                   // two tiles
constexpr int DT_NST = 2;                       // A ring stages (two CTAs per SM fit)
constexpr uint32_t DT_APLANE = DT_BM * 16 + 16; // padded plane: the half-warp-per-row stores are conflict-free (see gemm_tc.cu)
constexpr uint32_t DT_AIMG = (DT_BK / 8) * DT_APLANE;
constexpr uint32_t DT_QPLANE = DT_NQ * 16;
constexpr float DT_TOL = 3e-5f;                 // bf16x3 error bound is ~2^-16 sum|q_k d_k| <= 1.6e-5; typical 5e-7

int g_drmm_impl = 1;   // 1: tcgen05 kernel where the shape allows, 0: fp32 kernels
long long* g_drmm_dbg = nullptr;   // optional phase clocks of CTAs 0 and 1000, warp 0 (tools/drmm_timing.py)
#define DT_STAMP(k) do { if (dbg && lane == 0 && warp == 0 && (blockIdx.x == 0 || blockIdx.x == 1000)) dbg[(blockIdx.x ? 16 : 0) + (k)] = clock64(); } while (0)

__device__ __forceinline__ int dt_bin(float c) {
  // numpy.histogram(bins=[-1,-.5,0,.5,1,1]): half-open bins, last bin closed ({1.0}), outside dropped
  if (!(c >= -1.0f) || c > 1.0f) return -1;
  if (c == 1.0f) return 4;
  if (c < -0.5f) return 0;
  if (c < 0.0f) return 1;
  if (c < 0.5f) return 2;
  return 3;
}
__device__ __forceinline__ void dt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dt_bar_epi() { asm volatile("bar.sync 1, %0;" ::"n"(DT_LW * 32) : "memory"); }

// One cell with the arithmetic of drmm_kernel (drmm.cu): query row normalised from scalar lane-strided partial sums,
// document row from float4 lane-strided partial sums, both through the xor butterfly, inv = 1 / max(sqrt(ss), 1e-8),
// elements scaled one by one, then ONE sequential fp32 FMA chain over k (x, y, z, w of ascending float4).  Warp-collective;
// sq / sd: per-warp scratch of ES floats each.
__device__ float dt_exact_cos(const float* __restrict__ table, int E, int64_t qid, int64_t did, int lane, float* sq, float* sd) {
  const float* qs = table + qid * E;
  float ss = 0.f;
  for (int e = lane; e < E; e += 32) {
    float v = qs[e];
    sq[e] = v;
    ss += v * v;
  }
  ss = warp_sum(ss);
  const float invq = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
  for (int e = lane; e < E; e += 32) sq[e] *= invq;
  const float4* s4 = reinterpret_cast<const float4*>(table + did * E);
  float sd2 = 0.f;
  for (int e4 = lane; e4 < E / 4; e4 += 32) {
    float4 v = ldg_stream(s4 + e4);
    reinterpret_cast<float4*>(sd)[e4] = v;
    sd2 += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  sd2 = warp_sum(sd2);
  const float invd = 1.0f / fmaxf(sqrtf(sd2), 1e-8f);
  for (int e = lane; e < E; e += 32) sd[e] *= invd;
  __syncwarp();
  float acc = 0.f;
  if (lane == 0) {
#pragma unroll 5
    for (int k4 = 0; k4 < E / 4; ++k4) {
      const float4 qv = reinterpret_cast<const float4*>(sq)[k4];
      const float4 dv = reinterpret_cast<const float4*>(sd)[k4];
      acc = fmaf(qv.x, dv.x, acc);
      acc = fmaf(qv.y, dv.y, acc);
      acc = fmaf(qv.z, dv.z, acc);
      acc = fmaf(qv.w, dv.w, acc);
    }
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, acc, 0);
}

// ---- per-query operands, built once per query (shared by its N documents) ----------------------------------------
// record of query b in the workspace: [Q operand image: hi|lo x KP/8 planes x 32 rows x 16 B][gate logits: 32 floats]
// [token ids: 32 ints][zero-row flags: 32 ints]
__host__ __device__ inline size_t dt_qrec_bytes(int KP) { return (size_t)2 * (KP / 8) * DT_QPLANE + 3 * DT_NQ * 4; }

__global__ void __launch_bounds__(256) drmm_tc_qprep_kernel(const float* __restrict__ table, int V, int E, int KP,
                                                            const int64_t* __restrict__ q, int Lq, int64_t q_begin,
                                                            const float* __restrict__ wg, const float* __restrict__ bg,
                                                            uint8_t* __restrict__ qrec, int* err) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = q_begin + blockIdx.x;
  uint8_t* rec = qrec + (size_t)blockIdx.x * dt_qrec_bytes(KP);
  const uint32_t qhalf = (uint32_t)(KP / 8) * DT_QPLANE;
  float* gate = reinterpret_cast<float*>(rec + 2 * qhalf);
  int* qids = reinterpret_cast<int*>(gate + DT_NQ);
  int* qzero = qids + DT_NQ;
  // gather, gate logit on the raw row, normalise by max(||x||, eps) - the arithmetic of drmm_kernel - then the normalised row as
  // bf16 hi/lo in operand-image order; rows >= Lq and the K padding are zero
  for (int i = warp; i < DT_NQ; i += 8) {
    float inv = 0.f;
    const float* src = table;
    if (i < Lq) {
      const int64_t id = checked_id(q[b * Lq + i], V, err);
      src = table + id * E;
      float ss = 0.f, gl = 0.f;
      for (int e = lane; e < E; e += 32) {
        float v = src[e];
        ss += v * v;
        gl += v * wg[e];
      }
      ss = warp_sum(ss);
      gl = warp_sum(gl);
      inv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
      if (lane == 0) gate[i] = gl + bg[0], qids[i] = (int)id, qzero[i] = ss == 0.f;
    } else if (lane == 0) {
      gate[i] = 0.f, qids[i] = 0, qzero[i] = 1;
    }
    for (int e = lane; e < KP; e += 32) {
      const float x = (i < Lq && e < E) ? src[e] * inv : 0.f;
      __nv_bfloat16 hi, lo;
      split_bf16(x, hi, lo);
      const size_t off = (size_t)(e >> 3) * DT_QPLANE + (size_t)i * 16 + (e & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(rec + off) = hi;
      *reinterpret_cast<__nv_bfloat16*>(rec + qhalf + off) = lo;
    }
  }
}

// smem: query record (operand image + gate / ids / zero flags, one bulk copy) | A ring [DT_NST][hi|lo] (after the MMAs:
// flag list + scratch)
__global__ void __launch_bounds__(DT_THREADS, 2)
    drmm_tc_kernel(const float* __restrict__ table, int V, int E, int KP, const uint8_t* __restrict__ qrec, int64_t q_begin,
                   const int64_t* __restrict__ d, int N, int Lq, int Ld, int64_t pair_begin,
                   const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ w1,
                   const float* __restrict__ b1, const float* __restrict__ wo, const float* __restrict__ bo,
                   float* __restrict__ scores, int32_t* __restrict__ hist_out, int* err, long long* __restrict__ dbg) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ __align__(8) uint64_t a_full[DT_NST], empty[DT_NST], acc_full[2], q_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float invd[DT_MAXLD];
  __shared__ int dids[DT_MAXLD], nflag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t p = pair_begin + blockIdx.x;
  const int64_t b = p / N;
  const int planes = KP / 8, nkc = KP / DT_BK, ntile = (Ld + DT_BM - 1) / DT_BM;
  const uint32_t qhalf = (uint32_t)planes * DT_QPLANE;
  const uint32_t qbytes = (uint32_t)dt_qrec_bytes(KP);
  uint8_t* qimg = smraw;
  const float* gate = reinterpret_cast<const float*>(qimg + 2 * qhalf);
  const int* qids = reinterpret_cast<const int*>(gate + DT_NQ);
  const int* qzero = qids + DT_NQ;
  uint8_t* a_ring = qimg + ((qbytes + 127) & ~127u);

  DT_STAMP(0);
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == DT_LW * 32) {
    for (int s = 0; s < DT_NST; ++s) {
      mbar_init(&a_full[s], DT_LW);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    mbar_init(&q_full, 1);
    fence_mbar_init();
    nflag = 0;
    // this query's operands (built by drmm_tc_qprep_kernel): one bulk copy, overlapped with the first document rows
    mbar_arrive_expect_tx(&q_full, qbytes);
    const uint8_t* src = qrec + (size_t)(b - q_begin) * qbytes;
    for (uint32_t o = 0; o < qbytes; o += 16384) bulk_g2s(qimg + o, src + o, min(16384u, qbytes - o), &q_full);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_slot;
  DT_STAMP(1);

  if (warp == DT_LW) {
    // ===================== MMA issuer =====================
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, DT_NQ);
    const uint64_t ad0 = smem_desc(smem_u32(a_ring), DT_APLANE, 128);
    const uint64_t qd0 = smem_desc(smem_u32(qimg), DT_QPLANE, 128);
    mbar_wait(&q_full, 0);
    int g = 0;
    for (int t = 0; t < ntile; ++t) {
      for (int kc = 0; kc < nkc; ++kc, ++g) {
        const int s = g % DT_NST;
        mbar_wait(&a_full[s], (uint32_t)(g / DT_NST) & 1);
        tc_fence_after();
        const uint64_t ah = ad0 + (uint64_t)((uint32_t)s * 2 * DT_AIMG >> 4), al = ah + (uint64_t)(DT_AIMG >> 4);
        const uint64_t qh = qd0 + (uint64_t)((uint32_t)kc * (DT_BK / 8) * DT_QPLANE >> 4), ql = qh + (uint64_t)(qhalf >> 4);
#pragma unroll
        for (int ks = 0; ks < DT_BK / 16; ++ks) {
          const uint64_t ao = (uint64_t)(ks * ((2 * DT_APLANE) >> 4)), qo = (uint64_t)(ks * ((2 * DT_QPLANE) >> 4));
          mma_bf16_ss_w(tbase + (uint32_t)t * DT_NQ, ah + ao, qh + qo, idesc, (uint32_t)((kc | ks) != 0), issue);
          mma_bf16_ss_w(tbase + (uint32_t)t * DT_NQ, al + ao, qh + qo, idesc, 1, issue);
          mma_bf16_ss_w(tbase + (uint32_t)t * DT_NQ, ah + ao, ql + qo, idesc, 1, issue);
        }
        mma_commit_w(&empty[s], issue);
      }
      mma_commit_w(&acc_full[t], issue);
    }
  } else {
    // ===================== document-row loaders, then epilogue =====================
    // lanes 0-15 read the 16 float4 slices (256 contiguous bytes) of one row of the K chunk, lanes 16-31 the next row;
    // warp w owns rows 16 w .. 16 w + 15 of the tile, iteration i -> row 16 w + 2 i + lane / 16.
    // Two register buffers in ping-pong: while chunk g is converted, chunks g+1 AND g+2 are in flight (a gather like this
    // is bound by the bytes in flight per SM: 2 CTAs x 256 threads x 16 x 16 B = 128 KB).
    const int j = lane & 15, rsub = lane >> 4;
    constexpr int NV = 8;
    const int total = ntile * nkc;
    int rowoff[2][NV];   // element offset of the row in the table for both tiles (-1: beyond Ld)
    float ssq[NV];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int r = t * DT_BM + warp * 16 + 2 * i + rsub;
        const bool ok = t < ntile && r < Ld;
        const int64_t id = ok ? checked_id(d[p * Ld + r], V, err) : 0;
        rowoff[t][i] = ok ? (int)(id * E) : -1;
        if (ok && j == 0) dids[r] = (int)id;
      }
#pragma unroll
    for (int i = 0; i < NV; ++i) ssq[i] = 0.f;
    const uint32_t aoff = (uint32_t)(j >> 1) * DT_APLANE + (uint32_t)(warp * 16 + rsub) * 16 + (uint32_t)(j & 1) * 8;
    auto load_rows = [&](float4* buf, int g) {
      const int t = g >= nkc ? 1 : 0, kk = (g - t * nkc) * DT_BK + 4 * j;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int ro = t ? rowoff[1][i] : rowoff[0][i];
        buf[i] = (ro >= 0 && kk < E) ? ldg_stream(reinterpret_cast<const float4*>(table + ro + kk)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto process = [&](const float4* v, int g) {
      const int t = g >= nkc ? 1 : 0, kc = g - t * nkc, s = g % DT_NST;
#pragma unroll
      for (int i = 0; i < NV; ++i) ssq[i] += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      if (kc == nkc - 1) {
        // row norms of the finished tile: 16 lanes hold the partial sums of one row
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          float ss = ssq[i];
          ss += __shfl_xor_sync(0xffffffffu, ss, 1);
          ss += __shfl_xor_sync(0xffffffffu, ss, 2);
          ss += __shfl_xor_sync(0xffffffffu, ss, 4);
          ss += __shfl_xor_sync(0xffffffffu, ss, 8);
          const int r = t * DT_BM + warp * 16 + 2 * i + rsub;
          if (j == 0 && r < DT_MAXLD) invd[r] = ss > 0.f ? 1.0f / fmaxf(sqrtf(ss), 1e-8f) : 0.f;   // 0 marks an all-zero row
          ssq[i] = 0.f;
        }
      }
      mbar_wait_relaxed(&empty[s], ((uint32_t)(g / DT_NST) & 1) ^ 1);
      uint8_t* ah = a_ring + (size_t)s * 2 * DT_AIMG + aoff;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint32_t h0, l0, h1, l1;
        split_bf16x2(v[i].x, v[i].y, h0, l0);
        split_bf16x2(v[i].z, v[i].w, h1, l1);
        *reinterpret_cast<uint2*>(ah + (size_t)i * 32) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(ah + DT_AIMG + (size_t)i * 32) = make_uint2(l0, l1);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) dt_arrive(&a_full[s]);
    };
    float4 va[NV], vb[NV];
    load_rows(va, 0);
    if (total > 1) load_rows(vb, 1);
    DT_STAMP(2);
    for (int g = 0; g < total; g += 2) {
      process(va, g);
      if (g == 0) DT_STAMP(3);
      if (g + 2 < total) load_rows(va, g + 2);
      if (g + 1 < total) {
        process(vb, g + 1);
        if (g + 3 < total) load_rows(vb, g + 3);
      }
      if (g == 4) DT_STAMP(4);
    }
    // ---- epilogue: all MMAs retired (the A ring becomes the flag list + scratch), row norms visible ----
    DT_STAMP(5);
    mbar_wait_relaxed(&acc_full[ntile - 1], 0);
    tc_fence_after();
    dt_bar_epi();
    DT_STAMP(6);
    mbar_wait_relaxed(&q_full, 0);                                  // gate / ids / zero flags of the query record
    int* flist = reinterpret_cast<int*>(a_ring);                    // up to 32 x 256 cells
    float* scratch = reinterpret_cast<float*>(a_ring + 32 * 1024);  // 8 warps x 2 rows of KP floats
    int (*whist)[DT_NQ * 5] = reinterpret_cast<int (*)[DT_NQ * 5]>(a_ring + 32 * 1024 + DT_LW * 2 * KP * 4);   // per-warp counts
    const int tile = warp >> 2, qt = warp & 3;
    if (tile < ntile) {
      float acc[DT_NQ];
      tmem_ld32(tbase + ((uint32_t)(qt * 32) << 16) + (uint32_t)tile * DT_NQ, acc);
      tmem_ld_wait();
      DT_STAMP(10);
      const int r = tile * DT_BM + qt * 32 + lane;
      const bool rvalid = r < Ld;
      const float iv = rvalid ? invd[r] : 0.f;
      int* wh = whist[warp];   // this warp's private counts: no atomics, no contention
#pragma unroll
      for (int i = 0; i < DT_NQ; ++i) {
        if (i < Lq) {   // uniform
          const bool zero = iv == 0.f || qzero[i];
          const float c = zero ? 0.f : acc[i] * iv;
          const float ac = fabsf(c);
          const bool near = !zero && (ac < DT_TOL || fabsf(ac - 0.5f) < DT_TOL || fabsf(ac - 1.0f) < DT_TOL);
          const int bin = !rvalid ? -1 : (near ? -2 : dt_bin(c));
          int mine = 0;   // lane k < 5 keeps the count of bin k
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int cnt = __popc(__ballot_sync(0xffffffffu, bin == k));
            mine = lane == k ? cnt : mine;
          }
          if (lane < 5) wh[i * 5 + lane] = mine;
          if (bin == -2) flist[atomicAdd(&nflag, 1)] = (i << 16) | r;
        }
      }
    }
    else {
      for (int i = lane; i < DT_NQ * 5; i += 32) whist[warp][i] = 0;
    }
    DT_STAMP(11);
    tc_fence_before();
    dt_bar_epi();
    DT_STAMP(7);
    // ---- cells near a bin edge: the fp32 kernels' exact arithmetic ----
    const int nf = nflag;
    if (dbg && tid == 0 && (blockIdx.x == 0 || blockIdx.x == 1000)) dbg[(blockIdx.x ? 16 : 0) + 12] = nf;
    float* sq = scratch + (size_t)warp * 2 * KP;
    for (int f = warp; f < nf; f += DT_LW) {
      const int cell = flist[f], i = cell >> 16, r = cell & 0xffff;
      const float c = dt_exact_cos(table, E, qids[i], dids[r], lane, sq, sq + KP);
      const int bin = dt_bin(c);
      if (lane == 0 && bin >= 0) atomicAdd(&whist[0][i * 5 + bin], 1);
    }
    dt_bar_epi();
    DT_STAMP(8);
    // ---- totals, softmax gate over ALL Lq positions, ffnn(5->1->1), weighted sum, output (identical to drmm_kernel) ----
    if (warp == 0) {
      int hrow[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        int c = 0;
        if (lane < Lq)
          for (int w = 0; w < DT_LW; ++w) c += whist[w][lane * 5 + k];
        hrow[k] = c;
      }
      float gt = (lane < Lq) ? gate[lane] : -INFINITY;
      float mx = warp_max(gt);
      float ex = (lane < Lq) ? __expf(gt - mx) : 0.f;
      float den = warp_sum(ex);
      float f = 0.f;
      if (lane < Lq) {
        float f0 = b0[0];
#pragma unroll
        for (int k = 0; k < 5; ++k) f0 = fmaf(w0[k], (float)hrow[k], f0);
        f = (w1[0] * f0 + b1[0]) * (ex / den);
        if (hist_out)
#pragma unroll
          for (int k = 0; k < 5; ++k) hist_out[p * Lq * 5 + lane * 5 + k] = hrow[k];
      }
      f = warp_sum(f);
      if (lane == 0) scores[p] = wo[0] * f + bo[0];
    }
  }
  tc_fence_before();
  __syncthreads();
  DT_STAMP(9);
  if (warp == 0) tmem_dealloc(tbase, 64);
}

static size_t dt_smem_bytes(int KP) { return ((dt_qrec_bytes(KP) + 127) & ~(size_t)127) + (size_t)DT_NST * 2 * DT_AIMG; }

bool drmm_tc_usable(int E, int Lq, int Ld, const float* table, int64_t vocab) {
  if (!g_drmm_impl) return false;
  const int KP = (E + DT_BK - 1) / DT_BK * DT_BK;
  return (E & 3) == 0 && Lq <= DT_NQ && Ld <= DT_MAXLD && Ld >= 1 && ((uintptr_t)table & 15) == 0 && dt_smem_bytes(KP) <= 110 * 1024 &&
         vocab * (int64_t)E < (int64_t)1 << 31 &&
         (size_t)32 * 1024 + (size_t)DT_LW * 2 * KP * 4 + (size_t)DT_LW * DT_NQ * 5 * 4 <= (size_t)DT_NST * 2 * DT_AIMG;
}

size_t drmm_tc_workspace_bytes(int E, int64_t nq) {
  const int KP = (E + DT_BK - 1) / DT_BK * DT_BK;
  return (size_t)nq * dt_qrec_bytes(KP) + 256;
}

int32_t drmm_tc_forward(const cair_drmm_weights& w, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                        int64_t pair_begin, int64_t pair_count, float* scores, int32_t* hist_out, uint8_t* qrec, int* err,
                        cudaStream_t s) {
  const int E = w.emsize, KP = (E + DT_BK - 1) / DT_BK * DT_BK;
  const int64_t qb = pair_begin / N, nq = (pair_begin + pair_count - 1) / N - qb + 1;
  CAIR_LAUNCH(drmm_tc_qprep_kernel, (unsigned)nq, 256, 0, s, w.table, w.vocab, E, KP, q, Lq, qb, w.gating.w, w.gating.b, qrec, err);
  const size_t smem = dt_smem_bytes(KP);
  CAIR_CUDA(cudaFuncSetAttribute(drmm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(drmm_tc_kernel, (unsigned)pair_count, DT_THREADS, smem, s, w.table, w.vocab, E, KP, qrec, qb, d, N, Lq, Ld, pair_begin,
              w.ffnn0.w, w.ffnn0.b, w.ffnn1.w, w.ffnn1.b, w.output.w, w.output.b, scores, hist_out, err, g_drmm_dbg);
  return CAIR_OK;
}

}  // namespace cair
