// (Bi)LSTM / (Bi)GRU recurrence on the 5th-gen tensor cores, gate rows split over a thread-block cluster.
// RNNEncoder.forward semantics (neuroir/encoders/rnn_encoder.py:36-53,62-141): packed sequences without sorting
// (sequence s runs exactly len[s] steps, the reverse direction starts at its own last token, bank rows t >= len[s]
// are zeros), per-direction hidden size h = hidden_size // 2, torch gate orders (LSTM i,f,g,o; GRU r,z,n).
//
// Work split.  A cluster of CS = ceil(h / 32) CTAs owns `spc` sequences of one direction for all of their steps.
// CTA `rank` owns the 32 hidden units [32 rank, 32 rank + 32): their 4 x 32 = 128 gate rows are the M side of ONE
// tcgen05 tile per step,  G^T[128 x N] = W[128 x K] . Z^T,  Z = [x_t | 1 | h_{t-1}]  (N = sequences, padded to 16),
// with the weights resident in shared memory for all steps (bulk-copied once), fp32 accumulators in TMEM, and
// bf16x3 split precision (hi*hi + lo*hi + hi*lo) so the 200-step recurrence stays at fp32 accuracy.
// Compared with one CTA holding all 4h gate rows (round 1: 2 row tiles, 24 recurrent MMAs per step at h = 64), every SM
// issues only the MMAs of its own 128 rows (6 CS per step: 12 at h = 64) and runs only its own 32 units' cell updates
// (the MUFU-bound part); the price is the exchange of h: every epilogue warp writes the bf16 hi/lo operand rows of its
// cells (8-byte units assembled by two warp shuffles, 256 contiguous bytes per warp and column block) into the
// next-step operand buffer of its OWN CTA and forwards that chunk to every other CTA of the cluster with one bulk
// copy shared -> distributed shared memory (cp.async.bulk.shared::cluster.shared::cta, SASS UBLKCP.S.S) whose bytes
// are counted on an mbarrier of the destination (complete_tx); the MMA warp consumes the blocks in arrival order (own
// block first).  Measured alternatives (profiles/r02_rnn_exchange_modes.txt): st.async per 8-byte unit (1280
// transactions per step: the block arrives ~1000 cycles late), remote st.shared::cluster + mbarrier.arrive.release.cluster
// (the cluster-scope release alone costs ~1500 cycles per step), one bulk copy per TMEM quarter or one per CTA and
// destination issued by the MMA thread (no faster: the transfer, not the issue, is the cost; chunks sent as each warp
// finishes overlap it with the other warps' cell updates).
// The hop costs 400-850 cycles per step (pair dependent), which is why at h = 64 / 1280 x 2 sequences (cfg2: a pure
// latency chain of 200 steps) the single-CTA round-1 kernel (lstm_tc.cu) is still faster and stays the AUTO choice
// there; this kernel carries GRU, h > 64 (stock Match-Tensor 70 / direction, CARS 128) and in >= 48.
// GRU uses the same tile: per unit the four rows are r, z, n_x (input part of the candidate, W_in x + b_in) and
// n_h (hidden part, W_hn h + b_hn), so  n = tanh(n_x + r * n_h)  needs no second GEMM.
// Two input modes:
//   FUSED (in <= 47): x_t rows (embedding rows by token id from a pre-split bf16 hi/lo table, or dense fp32 rows) are
//     gathered into a 4-slot operand ring by four gather warps; the x part of step t+1 is issued right behind the h part
//     of step t into the other TMEM accumulator, the bias rides in K slot `in` (constant-1 column): no pre-gate tensor.
//   PRE (any in): pre-gates  P[n L, dirs 4h] = X W_ih'^T + b'  come from one tcgen05 GEMM (gemm_tc.cu) whose weight rows
//     are permuted to [dir][unit][4] and pre-scaled, so a cell reads its four pre-gates as ONE 16-byte load, prefetched a
//     step ahead.
// Every weight row is pre-scaled by -log2(e) (-2 log2(e) for the tanh rows): accumulators are exp2 arguments.
// Warp roles: warps 0-19 epilogue (TMEM quarter = warp % 4 -> 8 units, column blocks of 8 sequences round-robin over
// warp / 4; all four gates of a unit land in one thread via tcgen05.ld.16x256b on permuted rows), warp 20 MMA issuer,
// warps 21-23 gather (FUSED).
#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int RT_XP = 48;              // K slots of the fused x part (in + bias column <= 48)
constexpr int RT_UPC = 32;             // hidden units per CTA
constexpr int RT_MAXCS = 4;            // cluster size limit: h <= 128 per direction
constexpr int RT_EPI_WARPS = 20;       // 5 per SM sub-partition: at 33..40 sequences per cluster every warp owns exactly one column block
constexpr int RT_CW = RT_EPI_WARPS / 4; // column-block lanes
constexpr int RT_XS = 3;               // x-operand ring slots = gather warps
constexpr int RT_MMA_WARP = 20;
constexpr int RT_THREADS_FUSED = 24 * 32;
constexpr int RT_THREADS_PRE = 21 * 32;
constexpr int RT_MAXN = 128;           // sequences per cluster (MMA N)
constexpr uint32_t RT_APLANE = 128 * 16;
constexpr float RT_LOG2E = 1.4426950408889634f;

// warp 20 (sub-partition 0) issues the MMAs; the gather warps 21, 22, 23 sit on the other three sub-partitions
__device__ __forceinline__ int rt_gather_slot(int warp) { return warp >= 21 && warp < 21 + RT_XS ? warp - 21 : -1; }

bool rnn_tc_supported(int in, int h) { return in >= 1 && h >= 1 && h <= RT_UPC * RT_MAXCS; }

// ---- weight images --------------------------------------------------------------------------------------------
// Source row / scale of tile row `row` of CTA `rank`; part 0 = input weights + bias, part 1 = recurrent weights.
struct RtRow {
  int grow;       // row of w_ih / w_hh (-1: zero row)
  float scale;
  int bias_mode;  // 0: b_ih + b_hh, 1: b_ih only, 2: b_hh only
};
__device__ __forceinline__ RtRow rt_row(int gru, int h, int u, int type, int part) {
  RtRow r;
  r.scale = (type >= 2 && (gru || type == 2)) ? -2.0f * RT_LOG2E : -RT_LOG2E;
  if (!gru) {
    r.grow = type * h + u, r.bias_mode = 0;
    return r;
  }
  if (type < 2) {
    r.grow = type * h + u, r.bias_mode = 0;
  } else if (type == 2) {   // n_x: input part only
    r.grow = part == 0 ? 2 * h + u : -1, r.bias_mode = 1;
  } else {                  // n_h: hidden part only (+ b_hn)
    r.grow = part == 1 ? 2 * h + u : -1, r.bias_mode = 2;
  }
  return r;
}
__device__ __forceinline__ float rt_bias(const RtRow& r, int gru, int h, int u, int type, const float* b_ih, const float* b_hh) {
  const int brow = gru ? (type < 2 ? type * h + u : 2 * h + u) : type * h + u;
  const float bi = b_ih ? b_ih[brow] : 0.f, bh = b_hh ? b_hh[brow] : 0.f;
  return r.bias_mode == 0 ? bi + bh : r.bias_mode == 1 ? bi : bh;
}

// image of one (dir, rank): [hi|lo][plane][128 rows][8 x bf16]; planes = (fused ? 6 : 0) + 4 cs;
// tile row = 32 q + 8 type + j  <->  unit u = 32 rank + 8 q + j.
__global__ void rnn_tc_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                   const float* __restrict__ b_ih, const float* __restrict__ b_hh, int in, int h, int gru,
                                   int fused, int cs, uint8_t* __restrict__ img) {
  const int xpl = fused ? RT_XP / 8 : 0, planes = xpl + 4 * cs;
  const int total = cs * planes * 128 * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7, row = (idx >> 3) & 127, pl = (idx >> 10) % planes, rank = idx / (planes * 1024);
    const int type = (row >> 3) & 3, u = rank * 32 + (row >> 5) * 8 + (row & 7);
    float v = 0.f;
    if (u < h) {
      if (pl < xpl) {
        const int k = pl * 8 + e;
        const RtRow r = rt_row(gru, h, u, type, 0);
        if (k < in) v = r.grow >= 0 ? w_ih[(size_t)r.grow * in + k] * r.scale : 0.f;
        else if (k == in) v = rt_bias(r, gru, h, u, type, b_ih, b_hh) * r.scale;
      } else {
        const int k = (pl - xpl) * 8 + e;
        const RtRow r = rt_row(gru, h, u, type, 1);
        if (k < h && r.grow >= 0) v = w_hh[(size_t)r.grow * h + k] * r.scale;
      }
    }
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const size_t half = (size_t)planes * RT_APLANE;
    uint8_t* base = img + (size_t)rank * 2 * half;
    const size_t off = (size_t)pl * RT_APLANE + (size_t)row * 16 + e * 2;
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + half + off) = lo;
  }
}

// PRE mode: weights / bias of the pre-gate GEMM, output column c = dir 4h + 4 u + type, pre-scaled like the image.
__global__ void rnn_tc_pack_pre_kernel(const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                                       const float* __restrict__ b_hh, int in, int h, int gru, float* __restrict__ wp,
                                       float* __restrict__ bp) {
  const int64_t total = (int64_t)4 * h * in;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx / in), k = (int)(idx - (int64_t)c * in);
    const int u = c >> 2, type = c & 3;
    const RtRow r = rt_row(gru, h, u, type, 0);
    wp[idx] = r.grow >= 0 ? w_ih[(size_t)r.grow * in + k] * r.scale : 0.f;
    if (k == 0) bp[c] = rt_bias(r, gru, h, u, type, b_ih, b_hh) * r.scale;
  }
}

int32_t rnn_tc_pack(Owned& own, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, int in, int h, int rnn_type,
                    RnnTcPack* out, cudaStream_t s) {
  if (!fwd || !fwd->w_ih || !fwd->w_hh) return fail(CAIR_ERR_BAD_ARG, "rnn_tc: null weights");
  if (!rnn_tc_supported(in, h)) return fail(CAIR_ERR_UNSUPPORTED, "rnn_tc: in=%d h=%d not supported", in, h);
  const int dirs = rev ? 2 : 1;
  out->in = in, out->h = h, out->dirs = dirs, out->gru = rnn_type == CAIR_RNN_GRU ? 1 : 0;
  out->cs = (h + RT_UPC - 1) / RT_UPC;
  out->fused = in < RT_XP ? 1 : 0;
  out->planes = (out->fused ? RT_XP / 8 : 0) + 4 * out->cs;
  out->img_bytes = (size_t)2 * out->planes * RT_APLANE;
  CAIR_CUDA(own.alloc(&out->wimg, (size_t)dirs * out->cs * out->img_bytes));
  for (int d = 0; d < dirs; ++d) {
    const cair_lstm_dir* w = d ? rev : fwd;
    if (!w->w_ih || !w->w_hh) return fail(CAIR_ERR_BAD_ARG, "rnn_tc: null weights");
    CAIR_LAUNCH(rnn_tc_pack_kernel, 96, 256, 0, s, w->w_ih, w->w_hh, w->b_ih, w->b_hh, in, h, out->gru, out->fused, out->cs,
                out->wimg + (size_t)d * out->cs * out->img_bytes);
  }
  if (!out->fused) {
    const int PW = dirs * 4 * h;
    CAIR_CUDA(own.alloc(&out->wp, (size_t)PW * in));
    CAIR_CUDA(own.alloc(&out->bp, (size_t)PW));
    for (int d = 0; d < dirs; ++d) {
      const cair_lstm_dir* w = d ? rev : fwd;
      CAIR_LAUNCH(rnn_tc_pack_pre_kernel, 256, 256, 0, s, w->w_ih, w->b_ih, w->b_hh, in, h, out->gru,
                  out->wp + (size_t)d * 4 * h * in, out->bp + (size_t)d * 4 * h);
    }
    if ((in & 3) == 0) CAIR_TRY(gemm_tc_pack(own, out->wp, PW, in, &out->wp_tc, s));
  }
  return CAIR_OK;
}

// x-operand rows of a gathered table: row v = [hi: 48 bf16][lo: 48 bf16] (192 B), K slot `in` = 1 (bias column).
__global__ void rnn_tc_pack_table_kernel(const float* __restrict__ table, int V, int in, uint8_t* __restrict__ img) {
  const int64_t total = (int64_t)V * RT_XP;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / RT_XP;
    const int k = (int)(idx - v * RT_XP);
    const float x = k < in ? table[v * in + k] : (k == in ? 1.0f : 0.f);
    __nv_bfloat16 hi, lo;
    split_bf16(x, hi, lo);
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(img + v * (2 * RT_XP * 2));
    row[k] = hi;
    row[RT_XP + k] = lo;
  }
}
int32_t rnn_tc_pack_table(Owned& own, const float* table, int V, int in, uint8_t** img, cudaStream_t s) {
  if (in >= RT_XP) return fail(CAIR_ERR_UNSUPPORTED, "rnn_tc: pre-split table needs in < %d", RT_XP);
  CAIR_CUDA(own.alloc(img, (size_t)V * 2 * RT_XP * 2));
  CAIR_LAUNCH(rnn_tc_pack_table_kernel, 1184, 256, 0, s, table, V, in, *img);
  return CAIR_OK;
}

// ---- device helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rt_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void rt_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t rt_mapa(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
// Bulk copy own shared memory -> shared memory of another CTA of the cluster (async proxy, SASS UBLKCP); the bytes are
// counted on an mbarrier of the DESTINATION CTA (complete_tx): data and signal travel together, the producer never
// waits for the round trip and needs no cluster-scope release fence.
__device__ __forceinline__ void rt_bulk_s2c(uint32_t rdst, uint32_t lsrc, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst),
               "r"(lsrc), "r"(bytes), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void rt_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded waits: a protocol error traps (the launch fails with an error) instead of hanging the GPU.
constexpr uint32_t RT_SPIN_LIMIT = 1u << 26;
__device__ __forceinline__ void rt_wait(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity))
    if (++n > RT_SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void rt_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, n = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
        : "memory");
    if (!ok && ++n > (1u << 22)) __trap();
  }
}
// tcgen05.ld.16x256b.x1: 16 TMEM lanes x 8 columns per warp; thread (t0 = lane % 4, t1 = lane / 4) receives
// r0,r1 = (lane t1, cols 2 t0, 2 t0 + 1), r2,r3 = (lane t1 + 8, same cols).
__device__ __forceinline__ void rt_tmem_ld_16x256b(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
// 2^t, t clamped from above at 42 (products of three (1 + e) terms stay below 2^126; sigmoid / tanh are saturated to
// fp32 rounding long before).  ex2.approx: 2^-22 relative error; large negative t flushes to 0.
__device__ __forceinline__ float rt_ex2(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fminf(t, 42.0f)));
  return r;
}
__device__ __forceinline__ float rt_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// LSTM cell with 5 exponentials and 2 reciprocals.  Inputs are exp2 arguments: ti = -log2e a_i, tf = -log2e a_f,
// tg = -2 log2e a_g, to = -log2e a_o.   sigmoid(a) = 1/(1+ea), tanh(b) = (1-eb)/(1+eb), ea = e^-a, eb = e^-2b:
//   c' = [c (1+ei)(1+eg) + (1-eg)(1+ef)] / [(1+ef)(1+ei)(1+eg)],   h' = (1-ec) / [(1+eo)(1+ec)]
__device__ __forceinline__ void rt_lstm_cell(float ti, float tf, float tg, float to, float c, float& c_new, float& h_new) {
  const float ei = rt_ex2(ti), ef = rt_ex2(tf), eg = rt_ex2(tg), eo = rt_ex2(to);
  const float pi = 1.0f + ei, pf = 1.0f + ef, pg = 1.0f + eg;
  const float pig = pi * pg;
  const float num = fmaf(c, pig, (1.0f - eg) * pf);
  c_new = num * rt_rcp(pig * pf);
  const float ec = rt_ex2(-2.0f * RT_LOG2E * c_new);
  h_new = (1.0f - ec) * rt_rcp((1.0f + eo) * (1.0f + ec));
}
// GRU cell with 3 exponentials and 2 reciprocals: tr = -log2e a_r, tz = -log2e a_z, tnx / tnh = -2 log2e (n_x / n_h).
//   r = 1/(1+er),  en = e^-2(n_x + r n_h),  h' = (1-z) n + z h = [ez (1-en) + h (1+en)] / [(1+ez)(1+en)]
__device__ __forceinline__ float rt_gru_cell(float tr, float tz, float tnx, float tnh, float hprev) {
  const float er = rt_ex2(tr), ez = rt_ex2(tz);
  const float r = rt_rcp(1.0f + er);
  const float en = rt_ex2(fmaf(r, tnh, tnx));
  const float pn = 1.0f + en;
  return fmaf(ez, 1.0f - en, hprev * pn) * rt_rcp((1.0f + ez) * pn);
}

long long* g_rnn_dbg = nullptr;
int g_rnn_impl = RNN_IMPL_AUTO;
#define RT_T0() long long t0_ = a.dbg ? clock64() : 0
#define RT_ACC(slot) do { if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) a.dbg[slot] += clock64() - t0_; } while (0)

struct RnnTcArgs {
  GemmA x;                 // FUSED: x rows (gathered table rows or dense rows)
  const uint8_t* ximg;     // FUSED, gathered: pre-split table (rnn_tc_pack_table) or nullptr
  const float* pre;        // PRE: pre-gates [n L, dirs 4h], column dir 4h + 4u + type
  const uint8_t* wimg;     // [dirs][cs] images
  const int64_t* len;
  int n, L, in, h, dirs, cs, spc, npad, planes;
  float *out, *h_n, *c_n;
  int* err;
  long long* dbg;
};

// smem: W image (2 planes APLANE) | h operand [2 parities][cs blocks][hi|lo][4 planes][npad][16 B] | FUSED: x ring
// [RT_XS][hi|lo][6 planes][npad][16 B].   TMEM: [2 step parities][npad] fp32 columns.
template <bool GRU, bool FUSED, int NB>
__global__ void __launch_bounds__(FUSED ? RT_THREADS_FUSED : RT_THREADS_PRE, 1) rnn_tc_kernel(const RnnTcArgs a) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ __align__(8) uint64_t bar_w, bar_acc[2], bar_h[2][RT_MAXCS], x_full[RT_XS], x_empty[RT_XS];
  __shared__ uint32_t tmem_slot;
  __shared__ int slen[RT_MAXN];
  __shared__ int smaxlen;
  constexpr int NTHREADS = FUSED ? RT_THREADS_FUSED : RT_THREADS_PRE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = a.cs, npad = a.npad, spc = a.spc, h = a.h, L = a.L;
  const int rank = cs > 1 ? (int)rt_cluster_rank() : 0;
  const int dir = blockIdx.y, s0 = (blockIdx.x / cs) * spc;
  const int Hout = a.dirs * h;
  const uint32_t bplane = (uint32_t)npad * 16;           // one K chunk (8 elements) of all operand rows
  const uint32_t himg = 4 * bplane;                      // one (hi|lo) image of one 32-unit block
  const uint32_t hpar = (uint32_t)cs * 2 * himg;         // one parity
  const uint32_t ximg_b = (RT_XP / 8) * bplane;          // one (hi|lo) x image
  const uint32_t wbytes = (uint32_t)2 * a.planes * RT_APLANE;
  uint8_t* w_img = smraw;
  uint8_t* h_img = w_img + wbytes;
  uint8_t* x_img = h_img + 2 * hpar;
  const uint32_t tcols = 2 * npad <= 32 ? 32 : 2 * npad <= 64 ? 64 : 2 * npad <= 128 ? 128 : 256;

  if (warp == 0) tmem_alloc(&tmem_slot, tcols);
  if (tid == 32) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_acc[i], 1);
      // own block: the 16 epilogue warps arrive; blocks of the other CTAs: armed by the MMA thread, completed by their bytes
      for (int b = 0; b < RT_MAXCS; ++b) mbar_init(&bar_h[i][b], b == rank ? RT_EPI_WARPS : 1);
    }
    for (int i = 0; i < RT_XS; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < RT_MAXN; i += NTHREADS) {
    int s = s0 + i, l = 0;
    if (s < a.n && i < spc) {
      int64_t ll = a.len[s];
      if (ll < 1 || ll > L) {
        atomicOr(a.err, ERRF_BAD_LENGTH);
        ll = ll < 1 ? 1 : L;
      }
      l = (int)ll;
    }
    slen[i] = l;
  }
  // zero the operand buffers (h_0 = 0; K padding and unused rows stay zero for ever)
  {
    const uint32_t zbytes = 2 * hpar + (FUSED ? 2 * RT_XS * ximg_b : 0);
    for (uint32_t i = tid; i < zbytes / 16; i += NTHREADS) reinterpret_cast<uint4*>(h_img)[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    int m = 0;
    for (int s = 0; s < spc; ++s) m = max(m, slen[s]);
    smaxlen = m;
    mbar_arrive_expect_tx(&bar_w, wbytes);
    const uint8_t* src = a.wimg + ((size_t)dir * cs + rank) * wbytes;
    for (uint32_t o = 0; o < wbytes; o += 4 * RT_APLANE) bulk_g2s(w_img + o, src + o, min(4 * RT_APLANE, wbytes - o), &bar_w);
  }
  // zero the pad rows of the memory bank (this CTA's units of this direction)
  {
    const int u0 = rank * RT_UPC, nu = min(RT_UPC, h - u0);
    for (int s = 0; s < spc && nu > 0; ++s) {
      if (s0 + s >= a.n) break;
      const int npadr = (L - slen[s]) * nu;
      float* o = a.out + ((size_t)(s0 + s) * L + slen[s]) * Hout + dir * h + u0;
      for (int i = tid; i < npadr; i += NTHREADS) o[(size_t)(i / nu) * Hout + (i % nu)] = 0.f;
    }
  }
  __syncthreads();
  if (cs > 1) rt_cluster_sync();   // every CTA's barriers and zeroed buffers exist before any remote write / arrive
  const int maxlen = smaxlen;      // identical in all CTAs of the cluster (same sequences)
  const uint32_t tbase = tmem_slot;

  if (FUSED && rt_gather_slot(warp) >= 0) {
    // ===================== x gather: embedding rows (or dense rows) -> hi/lo bf16 ring =====================
    const GemmA& x = a.x;
    const int in = a.in;
    const bool vec = (in & 3) == 0 && (x.table ? ((x.E & 3) == 0) : ((x.lda & 3) == 0));
    const int slot = rt_gather_slot(warp);
    for (int step = slot; step < maxlen; step += RT_XS) {
      { RT_T0(); rt_wait_relaxed(&x_empty[slot], ((step / RT_XS) & 1) ^ 1); RT_ACC(7); }
      uint8_t* xs = x_img + (size_t)slot * 2 * ximg_b;
      for (int row = lane; row < spc; row += 32) {
        const int myl = slen[row];
        const bool active = step < myl;
        const int t = dir ? myl - 1 - step : step;
        const int64_t r = (int64_t)(s0 + row) * L + (active ? t : 0);
        uint8_t* xh = xs + (size_t)row * 16;
        if (a.ximg) {
          uint4 u[2 * (RT_XP / 8)];
          if (active) {
            const uint4* srow = reinterpret_cast<const uint4*>(a.ximg + checked_id(x.ids[r], x.V, x.err) * (2 * RT_XP * 2));
#pragma unroll
            for (int i = 0; i < 2 * (RT_XP / 8); ++i) u[i] = __ldg(srow + i);
          } else {
#pragma unroll
            for (int i = 0; i < 2 * (RT_XP / 8); ++i) u[i] = make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int pl = 0; pl < RT_XP / 8; ++pl) {
            if (pl * 8 <= in) {
              *reinterpret_cast<uint4*>(xh + (size_t)pl * bplane) = u[pl];
              *reinterpret_cast<uint4*>(xh + ximg_b + (size_t)pl * bplane) = u[RT_XP / 8 + pl];
            }
          }
          continue;
        }
        const float* src = nullptr;
        if (active) src = x.table ? x.table + checked_id(x.ids[r], x.V, x.err) * x.E : x.dense + r * x.lda;
        float v[RT_XP];
#pragma unroll
        for (int k4 = 0; k4 < RT_XP / 4; ++k4) {
          float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active && k4 * 4 < in) {
            if (vec) {
              f4 = *reinterpret_cast<const float4*>(src + k4 * 4);
            } else {
              f4.x = src[k4 * 4];
              if (k4 * 4 + 1 < in) f4.y = src[k4 * 4 + 1];
              if (k4 * 4 + 2 < in) f4.z = src[k4 * 4 + 2];
              if (k4 * 4 + 3 < in) f4.w = src[k4 * 4 + 3];
            }
          }
          if (k4 == (in >> 2)) {  // bias column: constant 1 at K slot `in`
            const int e = in & 3;
            f4.x = e == 0 ? 1.0f : f4.x, f4.y = e == 1 ? 1.0f : f4.y, f4.z = e == 2 ? 1.0f : f4.z, f4.w = e == 3 ? 1.0f : f4.w;
          }
          v[4 * k4] = f4.x, v[4 * k4 + 1] = f4.y, v[4 * k4 + 2] = f4.z, v[4 * k4 + 3] = f4.w;
        }
#pragma unroll
        for (int pl = 0; pl < RT_XP / 8; ++pl) {
          if (pl * 8 <= in) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16x2(v[pl * 8 + 2 * e], v[pl * 8 + 2 * e + 1], hi[e], lo[e]);
            *reinterpret_cast<uint4*>(xh + (size_t)pl * bplane) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(xh + ximg_b + (size_t)pl * bplane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) rt_arrive_local(&x_full[slot]);
    }
  } else if (warp == RT_MMA_WARP) {
    // ===================== MMA issuer (uniform control flow, one elected lane issues) =====================
    rt_wait(&bar_w, 0);
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, npad);
    const uint64_t wd0 = smem_desc(smem_u32(w_img), RT_APLANE, 128);
    // h operand: [blk][plane][8-row group][hi|lo][8 rows][16 B] - K chunks 2 bplane apart, 8-row groups 256 B apart
    const uint64_t hd0 = smem_desc(smem_u32(h_img), 2 * bplane, 256);
    const uint32_t hhi_h = (uint32_t)(hd0 >> 32);
    const uint64_t xd0 = smem_desc(smem_u32(x_img), bplane, 128);
    const uint32_t whi = (uint32_t)(wd0 >> 32), hhi = (uint32_t)(xd0 >> 32);
    const uint32_t wlo0 = (uint32_t)wd0, hlo0 = (uint32_t)hd0, xlo0 = (uint32_t)xd0;
    const uint32_t whalf = ((uint32_t)a.planes * RT_APLANE) >> 4;   // hi -> lo weight image
    const int xpl = FUSED ? RT_XP / 8 : 0;
    const int nxk = FUSED ? (a.in + 1 + 15) / 16 : 0;                // x k-steps in use (incl. the bias column)
    // x part of one step into accumulator `par` (fresh)
    auto issue_x = [&](int par, int slot) {
      const uint32_t tacc = tbase + (uint32_t)par * npad;
      const uint32_t xb = xlo0 + (((uint32_t)slot * 2 * ximg_b) >> 4);
      uint32_t acc = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t wp = wlo0 + (pass == 1 ? whalf : 0);          // weights: hi, lo, hi
        const uint32_t bp = xb + (pass == 2 ? (ximg_b >> 4) : 0);    // activations: hi, hi, lo
#pragma unroll
        for (int ks = 0; ks < RT_XP / 16; ++ks) {
          if (ks < nxk) {
            mma_bf16_ss_w32(tacc, wp + (uint32_t)(2 * ks) * (RT_APLANE >> 4), whi, bp + (uint32_t)(2 * ks) * (bplane >> 4), hhi,
                            idesc, acc, issue);
            acc = 1;
          }
        }
      }
    };
    long long ts_m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // bytes one source CTA delivers per step: (column blocks in use) x 4 quarter warps x 256 B
    const uint32_t blk_bytes = (uint32_t)((spc + 7) / 8) * 1024u;
    if (maxlen > 0 && lane == 0)   // h_0 = 0 is already in operand buffer 0: complete the first phase of the remote blocks
      for (int b = 0; b < cs; ++b)
        if (b != rank) rt_arrive_local(&bar_h[0][b]);
    if (FUSED && maxlen > 0) {
      { RT_T0(); rt_wait(&x_full[0], 0); RT_ACC(0); }
      tc_fence_after();
      issue_x(0, 0);
    }
    for (int step = 0; step < maxlen; ++step) {
      const int par = step & 1;
      const uint32_t ph = (uint32_t)(step >> 1) & 1;
      const uint32_t tacc = tbase + (uint32_t)par * npad;
      uint32_t acc = FUSED ? 1u : 0u;
      if (step + 1 < maxlen && lane == 0)   // arm the barriers of the NEXT step's h blocks (their bytes may already be landing)
        for (int b = 0; b < cs; ++b)
          if (b != rank) mbar_arrive_expect_tx(&bar_h[par ^ 1][b], blk_bytes);
      // h part: the 32-unit blocks of h_{step-1} in arrival order - own block first (its wait also guarantees that this
      // CTA's epilogue has finished reading the accumulator the next x part / PRE-mode MMA overwrites).
      if (step == 100 || step == 101) ts_m[(step - 100) * 4] = clock64();
      for (int bi = 0; bi < cs; ++bi) {
        int blk = rank + bi;
        blk = blk >= cs ? blk - cs : blk;
        // own block: generic-proxy stores fenced by the writers; other blocks: async-proxy bulk copies (complete_tx)
        { RT_T0(); rt_wait(&bar_h[par][blk], ph); RT_ACC(1); }
        if ((step == 100 || step == 101) && bi < 2) ts_m[(step - 100) * 4 + 1 + bi] = clock64();
        tc_fence_after();
        RT_T0();
        const uint32_t wb = wlo0 + (uint32_t)(xpl + 4 * blk) * (RT_APLANE >> 4);
        const uint32_t hb = hlo0 + (((uint32_t)par * hpar + (uint32_t)blk * 8 * bplane) >> 4);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t wp = wb + (pass == 1 ? whalf : 0);
          const uint32_t bp = hb + (pass == 2 ? (128u >> 4) : 0);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            mma_bf16_ss_w32(tacc, wp + (uint32_t)(2 * ks) * (RT_APLANE >> 4), whi, bp + (uint32_t)(2 * ks) * (2 * bplane >> 4), hhi_h,
                            idesc, acc, issue);
            acc = 1;
          }
        }
        RT_ACC(2);
      }
      mma_commit_w(&bar_acc[par], issue);
      if (step == 100 || step == 101) ts_m[(step - 100) * 4 + 3] = clock64();
      if (FUSED) {
        mma_commit_w(&x_empty[step % RT_XS], issue);
        if (step + 1 < maxlen) {
          const int nslot = (step + 1) % RT_XS;
          { RT_T0(); rt_wait(&x_full[nslot], ((step + 1) / RT_XS) & 1); RT_ACC(0); }
          tc_fence_after();
          issue_x(par ^ 1, nslot);
        }
      }
    }
    if (a.dbg && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && maxlen > 101)
      for (int i = 0; i < 8; ++i) a.dbg[16 + i] = ts_m[i];
  } else if (warp < RT_EPI_WARPS) {
    // ===================== epilogue warps =====================
    // warp -> (TMEM lane quarter q = 8 units, column-block lane cw); lane -> (t0 = lane % 4, j = lane / 4):
    // unit u = 32 rank + 8 q + j; block k of this warp = columns 8 (cw + 4k) ..+7, cells c = 0,1: sequence 8 b + 2 t0 + c
    const int q = warp & 3, cw = warp >> 2;
    const int t0i = lane & 3, j = lane >> 2;
    const int u = rank * RT_UPC + q * 8 + j;
    const bool uvalid = u < h;
    const int b0 = j & 1, b1 = (j >> 1) & 1, b2 = j >> 2;   // butterfly roles (see below)
    int lk[2 * NB];
    int ooff[2 * NB];      // bank offset of this step's h (relative to sequence s0), advanced by +-Hout per step
    float cst[2 * NB], hst[2 * NB];
    float4 pg[2 * NB];     // PRE: pre-gates of the NEXT step (prefetched)
    const int ostep = dir ? -Hout : Hout;
    const size_t PW = (size_t)a.dirs * 4 * h;
    float* const obase = a.out + (size_t)s0 * L * Hout + dir * h + u;
    const float* const pbase = a.pre ? a.pre + (size_t)s0 * L * PW + (size_t)dir * 4 * h + 4 * u : nullptr;
#pragma unroll
    for (int k = 0; k < NB; ++k)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int sq = 8 * (cw + RT_CW * k) + 2 * t0i + c;
        const int l = sq < RT_MAXN ? slen[sq] : 0;
        lk[2 * k + c] = l;
        cst[2 * k + c] = 0.f, hst[2 * k + c] = 0.f;
        ooff[2 * k + c] = (sq * L + (dir ? max(l - 1, 0) : 0)) * Hout;
        pg[2 * k + c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    auto load_pre = [&](int step) {
      if (!FUSED) {
#pragma unroll
        for (int k = 0; k < NB; ++k)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int sq = 8 * (cw + RT_CW * k) + 2 * t0i + c;
            const int l = lk[2 * k + c];
            if (uvalid && step < l) {
              const int t = dir ? l - 1 - step : step;
              pg[2 * k + c] = __ldg(reinterpret_cast<const float4*>(pbase + ((size_t)sq * L + t) * PW));
            }
          }
      }
    };
    // destination windows of the h operand buffer in every CTA of the cluster
    // lane i < cs - 1 forwards this warp's chunks to CTA (rank + 1 + i) % cs: that CTA's operand buffer and its
    // completion barrier of block `rank` (parity 0)
    const int fdst = (rank + 1 + lane) % cs;
    const uint32_t fwd_h = (cs > 1 && lane < cs - 1) ? rt_mapa(smem_u32(h_img), (uint32_t)fdst) : 0;
    const uint32_t fwd_b = (cs > 1 && lane < cs - 1) ? rt_mapa(smem_u32(&bar_h[0][rank]), (uint32_t)fdst) : 0;
    // block chunk (rank, plane q, group blk8) = 256 contiguous bytes: [hi|lo][8 rows][16 B]
    const uint32_t hoff_warp = (uint32_t)rank * 8 * bplane + (uint32_t)q * 2 * bplane;
    const uint32_t hoff_lane = (uint32_t)b1 * 128 + (uint32_t)(2 * t0i + b0) * 16 + (uint32_t)b2 * 8;
    const uint32_t tq = tbase + ((uint32_t)(q * 32) << 16);
    if (maxlen > 0) {
      if (lane == 0) rt_arrive_local(&bar_h[0][rank]);   // h_0 = 0 is already in operand buffer 0
      load_pre(0);
    }
    long long ts_e[6] = {0, 0, 0, 0, 0, 0};
    for (int step = 0; step < maxlen; ++step) {
      const int par = step & 1;
      float4 pgc[2 * NB];
      if (!FUSED) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) pgc[i] = pg[i];
        if (step + 1 < maxlen) load_pre(step + 1);
      }
      { RT_T0(); rt_wait(&bar_acc[par], (uint32_t)(step >> 1) & 1); if (warp == 0) RT_ACC(3); }
      if (step == 100) ts_e[0] = clock64();
      tc_fence_after();
      RT_T0();
      float hv[2 * NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const int blk8 = cw + RT_CW * k;
        hv[2 * k] = 0.f, hv[2 * k + 1] = 0.f;
        if (blk8 * 8 < spc) {   // warp-uniform
          float ga[4], gb[4];   // ga: gate types 0 (cells 0,1) and 1; gb: types 2 and 3
          const uint32_t ta = tq + (uint32_t)(par * npad + blk8 * 8);
          rt_tmem_ld_16x256b(ta, ga);
          rt_tmem_ld_16x256b(ta + (16u << 16), gb);
          tmem_ld_wait();
          tc_fence_before();
          if (step == 100 && k == 0) ts_e[1] = clock64();
          uint32_t w[2];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int i = 2 * k + c;
            float g0 = ga[c], g1 = ga[2 + c], g2 = gb[c], g3 = gb[2 + c];
            if (!FUSED) g0 += pgc[i].x, g1 += pgc[i].y, g2 += pgc[i].z, g3 += pgc[i].w;
            const bool act = step < lk[i];
            float hn_v;
            if (GRU) {
              hn_v = rt_gru_cell(g0, g1, g2, g3, hst[i]);
            } else {
              float cn;
              rt_lstm_cell(g0, g1, g2, g3, cst[i], cn, hn_v);
              cst[i] = act ? cn : cst[i];
            }
            hst[i] = act ? hn_v : hst[i];
            hv[i] = hn_v;
            uint32_t hi, lo;
            split_bf16_alu(hst[i], hi, lo);
            w[c] = (hi & 0xffffu) | (lo << 16);
          }
          // butterfly over the 8 lanes that share t0 (j = b0 + 2 b1 + 4 b2): after two exchanges this lane holds, for
          // cell b0, the hi (b1 = 0) or lo (b1 = 1) halves of units 4 b2 .. 4 b2 + 3 = one dense 8-byte store
          const uint32_t rA = __shfl_xor_sync(0xffffffffu, b0 ? w[0] : w[1], 4);
          const uint32_t ulo = b0 ? rA : w[0], uhi = b0 ? w[1] : rA;     // units (j & ~1), (j | 1) of cell b0
          const uint32_t HH = __byte_perm(ulo, uhi, 0x5410), LL = __byte_perm(ulo, uhi, 0x7632);
          const uint32_t rB = __shfl_xor_sync(0xffffffffu, b1 ? HH : LL, 8);
          const uint32_t w0 = b1 ? rB : HH, w1 = b1 ? LL : rB;
          const uint32_t off = (uint32_t)(par ^ 1) * hpar + hoff_warp + (uint32_t)blk8 * 256 + hoff_lane;
          *reinterpret_cast<uint2*>(h_img + off) = make_uint2(w0, w1);
        }
      }
      // hand the new h over: generic-proxy stores -> async proxy (the MMAs of this CTA and its bulk copies to the others)
      if (step == 100) ts_e[3] = clock64();
      fence_proxy_async();
      __syncwarp();
      if (step == 100) ts_e[4] = clock64();
      if (step + 1 < maxlen) {
        if (lane < cs - 1) {
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            const int blk8 = cw + RT_CW * k;
            if (blk8 * 8 < spc) {
              const uint32_t off = (uint32_t)(par ^ 1) * hpar + hoff_warp + (uint32_t)blk8 * 256;
              rt_bulk_s2c(fwd_h + off, smem_u32(h_img) + off, 256u, fwd_b + (uint32_t)((par ^ 1) * RT_MAXCS * 8));
            }
          }
        }
        if (step == 100) ts_e[5] = clock64();
        if (lane == 0) rt_arrive_local(&bar_h[par ^ 1][rank]);
      }
      if (step == 100) ts_e[2] = clock64();
      if (warp == 0) RT_ACC(4);
      // memory bank (fp32), off the critical path: 8 consecutive units x 8 sequences per warp store
#pragma unroll
      for (int i = 0; i < 2 * NB; ++i) {
        if (uvalid && step < lk[i]) obase[ooff[i]] = hv[i];
        ooff[i] += ostep;
      }
      if (warp == 0) RT_ACC(5);
    }
    if (a.dbg && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && maxlen > 101)
      for (int i = 0; i < 6; ++i) a.dbg[24 + warp * 6 + i] = ts_e[i];
    if (uvalid && (a.h_n || a.c_n)) {
#pragma unroll
      for (int k = 0; k < NB; ++k)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int sq = 8 * (cw + RT_CW * k) + 2 * t0i + c;
          if (sq >= spc || s0 + sq >= a.n) continue;
          if (a.h_n) a.h_n[((size_t)dir * a.n + s0 + sq) * h + u] = hst[2 * k + c];
          if (a.c_n && !GRU) a.c_n[((size_t)dir * a.n + s0 + sq) * h + u] = cst[2 * k + c];
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (cs > 1) rt_cluster_sync();   // no CTA exits while a peer may still write into its shared memory
  if (warp == 0) tmem_dealloc(tbase, tcols);
}

// ---- host side ------------------------------------------------------------------------------------------------
static size_t rt_smem_bytes(const RnnTcPack& p, int npad) {
  return (size_t)2 * p.planes * RT_APLANE + (size_t)2 * p.cs * 2 * 4 * npad * 16 +
         (p.fused ? (size_t)RT_XS * 2 * (RT_XP / 8) * npad * 16 : 0);
}
static int rt_max_npad(const RnnTcPack& p) {
  int best = 16;
  for (int np = 16; np <= RT_MAXN; np += 16)
    if (rt_smem_bytes(p, np) <= 220 * 1024) best = np;
  return best;
}
int g_rnn_spc_min = 8;   // process-wide floor of the sequences per cluster (tuning knob)
int g_rnn_spc_force = 0; // > 0: use exactly this many sequences per cluster (tools)

// Sequences per cluster.  Measured per-step cost (profiles/r02_cars_spc_sweep.txt, r02_rnn_exchange_modes.txt): fixed
// hand-over latencies (~1200 cycles) + per "round" of column blocks (20 epilogue warps = 5 blocks of 8 sequences per
// round) the MMAs and cell updates (~800 cycles) and the DSMEM exchange (~900 cycles per destination); more than two
// rounds per step spill the per-cell state.  A wave holds as many clusters as the hardware co-schedules
// (cudaOccupancyMaxActiveClusters; 148 / cs at best).
template <bool GRU, bool FUSED, int NB>
static int rt_max_clusters(int cs, size_t smem) {
  static int cache[RT_MAXCS + 1] = {0, 0, 0, 0, 0};
  if (cache[cs] > 0) return cache[cs];
  auto kern = rnn_tc_kernel<GRU, FUSED, NB>;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * (kSMs / cs)), 1);
    cfg.blockDim = dim3(FUSED ? RT_THREADS_FUSED : RT_THREADS_PRE);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    (void)cudaGetLastError();
  }
  if (n <= 0) n = kSMs / cs;
  return cache[cs] = n;
}

RnnTcPlan rnn_tc_plan(const RnnTcPack& p, int n, int min_spc) {
  RnnTcPlan pl;
  const int maxn = rt_max_npad(p);
  const int maxcl = p.fused ? rt_max_clusters<false, true, 1>(p.cs, rt_smem_bytes(p, 48))
                            : rt_max_clusters<false, false, 1>(p.cs, rt_smem_bytes(p, 48));
  if (min_spc < g_rnn_spc_min) min_spc = g_rnn_spc_min;
  if (min_spc > maxn) min_spc = maxn;
  double best = 1e30;
  int best_spc = min_spc;
  for (int spc = min_spc; spc <= maxn; ++spc) {
    const int rounds = ((spc + 7) / 8 + RT_CW - 1) / RT_CW;
    const int64_t groups = (int64_t)((n + spc - 1) / spc) * p.dirs;
    const int64_t waves = (groups + maxcl - 1) / maxcl;
    const double step = 1200.0 + rounds * (800.0 + 900.0 * (p.cs - 1)) * (rounds > 2 ? 1.5 : 1.0);
    const double cost = (double)waves * step;
    if (cost < best * 0.999) best = cost, best_spc = spc;
  }
  if (g_rnn_spc_force > 0) best_spc = std::min(g_rnn_spc_force, maxn);
  pl.spc = best_spc;
  pl.npad = (best_spc + 15) / 16 * 16;
  pl.groups = (n + best_spc - 1) / best_spc;
  pl.ctas = pl.groups * p.dirs * p.cs;
  pl.nb = ((best_spc + 7) / 8 + RT_CW - 1) / RT_CW;
  return pl;
}

// pre-gates [n*L, dirs*4h] and, when the pre-gate weights span several column tiles, the A operand image of the pre-gate GEMM
// (the embedding rows gathered / converted once instead of once per column tile)
static bool rt_pre_uses_aimg(const RnnTcPack& p) { return !p.fused && p.wp_tc.img && p.wp_tc.nct >= 2; }
size_t rnn_tc_workspace_floats(const RnnTcPack& p, int64_t n, int L) {
  if (p.fused) return 0;
  size_t f = (size_t)n * L * p.dirs * 4 * p.h;
  if (rt_pre_uses_aimg(p)) f += gemm_tc_aimg_bytes(n * L, p.in) / sizeof(float) + 64;
  return f;
}

template <bool GRU, bool FUSED, int NB>
static int32_t rt_launch(const RnnTcArgs& a, const RnnTcPlan& pl, int dirs, size_t smem, cudaStream_t s) {
  auto kern = rnn_tc_kernel<GRU, FUSED, NB>;
  CAIR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pl.groups * a.cs), (unsigned)dirs);
  cfg.blockDim = dim3(FUSED ? RT_THREADS_FUSED : RT_THREADS_PRE);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)a.cs, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail(CAIR_ERR_CUDA, "rnn_tc launch (cs=%d, npad=%d): %s", a.cs, a.npad, cudaGetErrorString(e));
  return CAIR_OK;
}
template <bool GRU, bool FUSED>
static int32_t rt_launch_nb(const RnnTcArgs& a, const RnnTcPlan& pl, int dirs, size_t smem, cudaStream_t s) {
  switch (pl.nb) {
    case 1: return rt_launch<GRU, FUSED, 1>(a, pl, dirs, smem, s);
    case 2: return rt_launch<GRU, FUSED, 2>(a, pl, dirs, smem, s);
    case 3: return rt_launch<GRU, FUSED, 3>(a, pl, dirs, smem, s);
    default: return rt_launch<GRU, FUSED, 4>(a, pl, dirs, smem, s);
  }
}

int32_t rnn_tc_run(const RnnTcPack& p, const GemmA& x, const int64_t* len, int n, int L, float* out, float* h_n, float* c_n,
                   float* ws_pre, int* err, cudaStream_t s, const char* rec_name, const uint8_t* ximg, int min_spc) {
  if (n <= 0) return CAIR_OK;
  const RnnTcPlan pl = rnn_tc_plan(p, n, min_spc);
  if (!p.fused) {
    if (!ws_pre) return fail(CAIR_ERR_WORKSPACE, "rnn_tc: pre-gate workspace missing");
    const int PW = p.dirs * 4 * p.h;
    uint8_t* aimg = nullptr;
    if (rt_pre_uses_aimg(p)) aimg = reinterpret_cast<uint8_t*>(((uintptr_t)(ws_pre + (size_t)n * L * PW) + 127) & ~(uintptr_t)127);
    CAIR_TRY(gemm_auto(x, p.wp, p.wp_tc, p.bp, ws_pre, PW, (int64_t)n * L, PW, p.in, ACT_NONE, s, aimg));
  }
  if (rec_name) prof_mark(rec_name, s);
  RnnTcArgs a;
  a.x = x, a.ximg = (p.fused && x.table) ? ximg : nullptr, a.pre = p.fused ? nullptr : ws_pre, a.wimg = p.wimg, a.len = len;
  a.n = n, a.L = L, a.in = p.in, a.h = p.h, a.dirs = p.dirs, a.cs = p.cs, a.spc = pl.spc, a.npad = pl.npad, a.planes = p.planes;
  a.out = out, a.h_n = h_n, a.c_n = c_n, a.err = err, a.dbg = g_rnn_dbg;

  const size_t smem = rt_smem_bytes(p, pl.npad);
  if (p.gru) return p.fused ? rt_launch_nb<true, true>(a, pl, p.dirs, smem, s) : rt_launch_nb<true, false>(a, pl, p.dirs, smem, s);
  return p.fused ? rt_launch_nb<false, true>(a, pl, p.dirs, smem, s) : rt_launch_nb<false, false>(a, pl, p.dirs, smem, s);
}

}  // namespace cair
