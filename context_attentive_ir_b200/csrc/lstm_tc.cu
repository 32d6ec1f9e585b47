// (Bi)LSTM recurrence on the 5th-gen tensor cores: one fused tcgen05 tile per time step.
// RNNEncoder.forward semantics (neuroir/encoders/rnn_encoder.py:62-141), same contract as lstm.cu.
//
// Per step the gate pre-activations of NSEQ=32 sequences are ONE GEMM  G^T[4h x 32] = W[4h x K] . Z^T,
// Z = [x_t | h_{t-1}] (K = 48 + 64), i.e. the input projection and the recurrent projection are fused:
// no pre-gate tensor is ever written to HBM.  The big, constant operand (the weights, up to 2 x 128 gate
// rows) is the MMA's M side and stays resident in shared memory for all steps (bulk-copied once); the
// small per-step operand (32 sequences) is the N side, so a step costs 2 x 7 x 3 MMAs of 16 cycles
// instead of 128-cycle ones.  bf16x3 split precision (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM)
// keeps ~fp32 accuracy through the 200-step recurrence (plain bf16/tf32 does not hold the 1e-3 bar).
// Gate rows are permuted so that, inside every 32-lane TMEM quarter, rows 8*type + j hold gate `type` (i,f,g,o) of
// unit j: a pair of tcgen05.ld.16x256b then hands each thread ALL FOUR gates of one unit for 4 sequences, so the
// whole cell update (5 exponentials + 2 reciprocals per element, c/h state in registers) runs without any
// inter-thread exchange.
// Warp roles (672 threads): warps 0-15 epilogue (TMEM -> gates -> c/h update; h written as next step's bf16 hi/lo
// operand, then the fp32 memory bank), warp 16: MMA issuer - the x part of step t+1 is issued right behind the h
// part of step t into the other TMEM accumulator, so only the recurrent half of the GEMM sits on the critical path,
// warps 17-20: gather of x (embedding rows by token id, or dense rows) into a 4-slot operand ring, one warp per
// slot, so the id -> row -> convert latency chain of a step has four step times to complete.
#include "models.cuh"
#include "umma.cuh"

namespace cair {

using namespace umma;

constexpr int LT_XP = 48;                 // K slots of the x part (in <= 48)
constexpr int LT_HP = 64;                 // K slots of the h part (h <= 64)
constexpr int LT_K = LT_XP + LT_HP;       // 112
constexpr int LT_PLANES = LT_K / 8;       // 14
constexpr int LT_NSEQ = 32;               // sequences per CTA = N of the MMA
constexpr int LT_EPI_WARPS = 16;          // epilogue warps (4 per SM sub-partition)
constexpr int LT_XS = 4;                   // x-operand ring slots = gather warps (warp g fills slot g for steps = g mod 4)
// Auxiliary warps are placed so that no SM sub-partition (warp % 4) carries both the MMA issuer and a gather warp on
// top of its four epilogue warps: warp 17 = issuer, warps 18, 19, 22, 23 = gather slots 0..3; warps 16, 20, 21 idle.
constexpr int LT_WARPS = 24;
constexpr int LT_MMA_WARP = 17;
constexpr int LT_THREADS = LT_WARPS * 32;
__device__ __forceinline__ int lt_gather_slot(int warp) {
  return warp == 18 ? 0 : warp == 19 ? 1 : warp == 22 ? 2 : warp == 23 ? 3 : -1;
}
constexpr uint32_t LT_APLANE = 128 * 16;  // weight image: 128 rows per plane
constexpr uint32_t LT_BPLANE = LT_NSEQ * 16;
constexpr uint32_t LT_AIMG = LT_PLANES * LT_APLANE;  // one (row tile, hi|lo) image: 28672 B
constexpr uint32_t LT_BIMG = LT_PLANES * LT_BPLANE;  // one (hi|lo) image: 7168 B

// in < LT_XP: K slot `in` of the x part carries a constant 1 whose weight column is the bias (folded into the GEMM)
bool lstm_tc_supported(int in, int h) { return in >= 1 && in < LT_XP && h >= 1 && h <= LT_HP; }

// weight image [dir][row tile][hi|lo][plane][row][8 x bf16]; row = 32*q + 8*type + j  <->  gate row type*h + (32*tile + 8*q + j).
// Column `in` holds b_ih + b_hh (the x operand carries a constant 1 there), and every row is pre-scaled by -log2(e)
// (-2 log2(e) for the cell gate g): the accumulator is directly the exp2 argument of e^-a / e^-2g in lt_cell.
__global__ void lstm_tc_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, int in, int h,
                                    uint8_t* __restrict__ img) {
  const int total = 2 * LT_PLANES * 128 * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7, row = (idx >> 3) & 127, pl = (idx >> 10) % LT_PLANES, mt = idx / (LT_PLANES * 1024);
    const int k = pl * 8 + e, type = (row >> 3) & 3, u = mt * 32 + (row >> 5) * 8 + (row & 7);
    float v = 0.f;
    if (u < h) {
      const int grow = type * h + u;
      if (k < LT_XP) {
        if (k < in) v = w_ih[(size_t)grow * in + k];
        else if (k == in) v = (b_ih ? b_ih[grow] : 0.f) + (b_hh ? b_hh[grow] : 0.f);
      } else if (k - LT_XP < h) {
        v = w_hh[(size_t)grow * h + (k - LT_XP)];
      }
    }
    v *= (type == 2) ? -2.0f * 1.4426950408889634f : -1.4426950408889634f;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    const size_t off = (size_t)pl * LT_APLANE + (size_t)row * 16 + e * 2;
    uint8_t* base = img + (size_t)mt * 2 * LT_AIMG;
    *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(base + LT_AIMG + off) = lo;
  }
}

int32_t lstm_tc_pack(Owned& own, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, int in, int h, LstmTcPack* out,
                     cudaStream_t s) {
  const int dirs = rev ? 2 : 1;
  out->in = in, out->h = h, out->dirs = dirs;
  CAIR_CUDA(own.alloc(&out->wimg, (size_t)dirs * 4 * LT_AIMG));
  for (int d = 0; d < dirs; ++d) {
    const cair_lstm_dir* w = d ? rev : fwd;
    CAIR_LAUNCH(lstm_tc_pack_kernel, 64, 256, 0, s, w->w_ih, w->w_hh, w->b_ih, w->b_hh, in, h, out->wimg + (size_t)d * 4 * LT_AIMG);
  }
  return CAIR_OK;
}

// Re-packs changed weights into an existing image (training: the parameters move every step).
int32_t lstm_tc_repack(const LstmTcPack& p, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, cudaStream_t s) {
  for (int d = 0; d < p.dirs; ++d) {
    const cair_lstm_dir* w = d ? rev : fwd;
    CAIR_LAUNCH(lstm_tc_pack_kernel, 64, 256, 0, s, w->w_ih, w->w_hh, w->b_ih, w->b_hh, p.in, p.h, p.wimg + (size_t)d * 4 * LT_AIMG);
  }
  return CAIR_OK;
}

// x-operand rows of a gathered table: row v = [hi: LT_XP bf16][lo: LT_XP bf16] (192 B), K slot `in` = 1 (bias column).
__global__ void lstm_tc_pack_table_kernel(const float* __restrict__ table, int V, int in, uint8_t* __restrict__ img) {
  const int64_t total = (int64_t)V * LT_XP;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = idx / LT_XP;
    const int k = (int)(idx - v * LT_XP);
    const float x = k < in ? table[v * in + k] : (k == in ? 1.0f : 0.f);
    __nv_bfloat16 hi, lo;
    split_bf16(x, hi, lo);
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(img + v * (2 * LT_XP * 2));
    row[k] = hi;
    row[LT_XP + k] = lo;
  }
}

int32_t lstm_tc_pack_table(Owned& own, const float* table, int V, int in, uint8_t** img, cudaStream_t s) {
  CAIR_CUDA(own.alloc(img, (size_t)V * 2 * LT_XP * 2));
  CAIR_LAUNCH(lstm_tc_pack_table_kernel, 1184, 256, 0, s, table, V, in, *img);
  return CAIR_OK;
}

__device__ __forceinline__ void lt_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void lt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// tcgen05.ld.16x256b.x2: 16 TMEM lanes x 16 columns per warp; thread (t0 = lane % 4, t1 = lane / 4) receives
// r0,r1 = (lane t1, cols 2t0, 2t0+1), r2,r3 = (lane t1+8, same cols), r4..r7 = the same for cols + 8.
__device__ __forceinline__ void lt_tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// 2^t with t clamped from above at 42: products of three (1 + e) terms stay below 2^126, and sigmoid / tanh are
// saturated to fp32 rounding long before.  ex2.approx: 2^-22 relative error; large negative t flushes to 0.
__device__ __forceinline__ float lt_ex2(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fminf(t, 42.0f)));
  return r;
}
__device__ __forceinline__ float lt_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// One LSTM cell update with 5 exponentials and 2 reciprocals (instead of 5 + 5).  Inputs are the exp2 arguments the
// pre-scaled weights produce: ti = -log2e a_i, tf = -log2e a_f, tg = -2 log2e a_g, to = -log2e a_o.
//   sigmoid(a) = 1/(1+ea),  tanh(b) = (1-eb)/(1+eb)  with  ea = e^-a, eb = e^-2b
//   c' = sigmoid(f) c + sigmoid(i) tanh(g) = [c (1+ei)(1+eg) + (1-eg)(1+ef)] / [(1+ef)(1+ei)(1+eg)]
//   h' = sigmoid(o) tanh(c')              = (1-ec) / [(1+eo)(1+ec)]
__device__ __forceinline__ void lt_cell(float ti, float tf, float tg, float to, float c, float& c_new, float& h_new) {
  const float ei = lt_ex2(ti), ef = lt_ex2(tf), eg = lt_ex2(tg), eo = lt_ex2(to);
  const float pi = 1.0f + ei, pf = 1.0f + ef, pg = 1.0f + eg;
  const float pig = pi * pg;
  const float num = fmaf(c, pig, (1.0f - eg) * pf);
  c_new = num * lt_rcp(pig * pf);
  const float ec = lt_ex2(-2.0f * 1.4426950408889634f * c_new);
  h_new = (1.0f - ec) * lt_rcp((1.0f + eo) * (1.0f + ec));
}

// The same cell update that also returns the four gate activations (training: the BPTT kernel wants sigma(i), sigma(f),
// tanh(g), sigma(o) and c per step).  No extra MUFU work: every activation is a product of terms already at hand and one of the
// two reciprocals (1/(1+ei) = (1+eg)(1+ef) r1, ...).
__device__ __forceinline__ void lt_cell_save(float ti, float tf, float tg, float to, float c, float& c_new, float& h_new,
                                             float& gi, float& gf, float& gg, float& go) {
  const float ei = lt_ex2(ti), ef = lt_ex2(tf), eg = lt_ex2(tg), eo = lt_ex2(to);
  const float pi = 1.0f + ei, pf = 1.0f + ef, pg = 1.0f + eg;
  const float pig = pi * pg, mg = 1.0f - eg;
  const float r1 = lt_rcp(pig * pf);
  c_new = fmaf(c, pig, mg * pf) * r1;
  gi = pg * pf * r1;
  gf = pig * r1;
  gg = mg * pi * pf * r1;
  const float ec = lt_ex2(-2.0f * 1.4426950408889634f * c_new);
  const float pc = 1.0f + ec;
  const float r2 = lt_rcp((1.0f + eo) * pc);
  h_new = (1.0f - ec) * r2;
  go = pc * r2;
}

// Optional role timing (dbg != nullptr; CTA (0,0), lane 0 of the role's first warp), see tools/lstm_timing.py
#define LT_T0() long long t0_ = dbg ? clock64() : 0
#define LT_ACC(slot) do { if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) dbg[slot] += clock64() - t0_; } while (0)
long long* g_lstm_dbg = nullptr;

constexpr uint32_t LT_XIMG = (LT_XP / 8) * LT_BPLANE;     // one (hi|lo) x image: 6 planes, 3072 B
constexpr uint32_t LT_HIMG = (LT_HP / 8) * LT_BPLANE;     // one (hi|lo) h image: 8 planes, 4096 B
constexpr uint32_t LT_TCOLS = 4 * LT_NSEQ;                 // TMEM columns: 2 parities x 2 row tiles x 32

// smem: W image (4 x LT_AIMG) | h operand [2 parities][hi|lo] | x operand ring [LT_XS][hi|lo]
// TMEM: [2 step parities][2 row tiles][32 sequences] fp32 columns
// SAVE (training forward): additionally writes the gate activations [n*L, dirs*4h] (i, f, g, o blocks per direction) and the
// cell states [n*L, dirs*h] of every valid step - the inputs of the BPTT kernel (train.cu).
template <bool SAVE>
__global__ void __launch_bounds__(LT_THREADS, 1)
    lstm_tc_kernel(GemmA x, const uint8_t* __restrict__ wimg_all, const float* __restrict__ bias_all,
                   const int64_t* __restrict__ len, int n, int L, int in, int h, int dirs, uint32_t ks_mask, int spc,
                   float* __restrict__ out, float* __restrict__ h_n, float* __restrict__ c_n, int* err,
                   long long* __restrict__ dbg, const uint8_t* __restrict__ ximg, float* __restrict__ gates_out,
                   float* __restrict__ cseq_out) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t bar_w, bar_h[2], bar_acc[2], x_full[LT_XS], x_empty[LT_XS];
  __shared__ uint32_t tmem_slot;
  __shared__ int slen[LT_NSEQ];
  __shared__ int smaxlen;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y, s0 = blockIdx.x * spc;  // spc <= LT_NSEQ sequences per CTA; MMA columns beyond stay empty
  const int nmt = (h + 31) / 32;  // row tiles in use (1 or 2)
  if (dbg && tid == 0 && blockIdx.x == 1 && blockIdx.y == 0) {  // whole-kernel cycles / ns of the un-instrumented CTA (1,0)
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    dbg[8] -= clock64(), dbg[9] -= (long long)ns;
  }
  uint8_t* w_img = smraw;
  uint8_t* h_img = w_img + 4 * LT_AIMG;
  uint8_t* x_img = h_img + 4 * LT_HIMG;
  (void)bias_all;  // folded into the weight image
  const int Hout = dirs * h;

  if (warp == 0) tmem_alloc(&tmem_slot, LT_TCOLS);
  if (tid == 32) {
    mbar_init(&bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_h[i], LT_EPI_WARPS / 2);  // epilogue warps of row tile i: their 32 units of h_t are in place
      mbar_init(&bar_acc[i], 1);               // accumulator of row tile i is complete
    }
    for (int i = 0; i < LT_XS; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (tid < LT_NSEQ) {
    int s = s0 + tid, l = 0;
    if (s < n && tid < spc) {
      int64_t ll = len[s];
      if (ll < 1 || ll > L) {
        atomicOr(err, ERRF_BAD_LENGTH);
        ll = ll < 1 ? 1 : L;
      }
      l = (int)ll;
    }
    slen[tid] = l;
  }
  // zero the operand buffers (h_0 = 0, K padding stays zero for ever)
  for (int i = tid; i < (int)((4 * LT_HIMG + 2 * LT_XS * LT_XIMG) / 16); i += LT_THREADS)
    reinterpret_cast<uint4*>(h_img)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    int m = 0;
    for (int s = 0; s < LT_NSEQ; ++s) m = max(m, slen[s]);
    smaxlen = m;
    // weights of this direction: one bulk copy, resident for all steps
    const uint32_t bytes = (uint32_t)(nmt * 2) * LT_AIMG;
    mbar_arrive_expect_tx(&bar_w, bytes);
    const uint8_t* src = wimg_all + (size_t)dir * 4 * LT_AIMG;
    for (uint32_t o = 0; o < bytes; o += LT_AIMG) bulk_g2s(w_img + o, src + o, LT_AIMG, &bar_w);
  }
  // The pad rows of the memory bank (this direction's half) are zeroed by the three otherwise idle warps WHILE the recurrence
  // runs (nothing in this kernel reads the bank; with the dataset's real lengths - documents average 63 of 200 positions -
  // an up-front fill by all threads cost ~0.02 ms before the first step).
  __syncthreads();
  const int maxlen = smaxlen;
  const uint32_t tbase = tmem_slot;
  if (warp == 16 || warp == 20 || warp == 21) {
    const int wi = (warp == 16 ? 0 : warp == 20 ? 1 : 2) * 32 + lane;
    const bool v4 = (h & 3) == 0 && (Hout & 3) == 0 && ((uintptr_t)out & 15) == 0;
    for (int s = 0; s < spc; ++s) {
      if (s0 + s >= n) break;
      float* o = out + ((size_t)(s0 + s) * L + slen[s]) * Hout + dir * h;
      if (v4) {
        const int h4 = h >> 2, npad4 = (L - slen[s]) * h4;
        for (int i = wi; i < npad4; i += 96)
          *reinterpret_cast<float4*>(o + (size_t)(i / h4) * Hout + (i % h4) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const int npad = (L - slen[s]) * h;
        for (int i = wi; i < npad; i += 96) o[(size_t)(i / h) * Hout + (i % h)] = 0.f;
      }
    }
  }

  if (lt_gather_slot(warp) >= 0) {
    // ===================== x gather: embedding rows (or dense rows) -> hi/lo bf16 ring =====================
    // One warp per ring slot: each has LT_XS step times to cover its id -> row -> convert latency chain.
    const int myl = slen[lane];   // lane <-> sequence row
    const bool vec = (in & 3) == 0 && (x.table ? ((x.E & 3) == 0) : ((x.lda & 3) == 0));
    const int slot = lt_gather_slot(warp);
    for (int step = slot; step < maxlen; step += LT_XS) {
      { LT_T0(); mbar_wait_relaxed(&x_empty[slot], ((step / LT_XS) & 1) ^ 1); LT_ACC(7); }
      const bool active = step < myl;
      const int t = dir ? myl - 1 - step : step;
      const int64_t r = (int64_t)(s0 + lane) * L + (active ? t : 0);
      if (ximg) {
        // pre-split table: 12 x 16-byte units per token, no conversion (keeps the gather off the issue slots the
        // epilogue warps of this SM sub-partition need)
        uint4 u[2 * (LT_XP / 8)];
        if (active) {
          const uint4* srow = reinterpret_cast<const uint4*>(ximg + checked_id(x.ids[r], x.V, x.err) * (2 * LT_XP * 2));
#pragma unroll
          for (int i = 0; i < 2 * (LT_XP / 8); ++i) u[i] = __ldg(srow + i);
        } else {
#pragma unroll
          for (int i = 0; i < 2 * (LT_XP / 8); ++i) u[i] = make_uint4(0, 0, 0, 0);
        }
        uint8_t* xh = x_img + (size_t)slot * 2 * LT_XIMG + (size_t)lane * 16;
#pragma unroll
        for (int pl = 0; pl < LT_XP / 8; ++pl) {
          if (pl * 8 <= in) {
            *reinterpret_cast<uint4*>(xh + (size_t)pl * LT_BPLANE) = u[pl];
            *reinterpret_cast<uint4*>(xh + LT_XIMG + (size_t)pl * LT_BPLANE) = u[LT_XP / 8 + pl];
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) lt_arrive(&x_full[slot]);
        continue;
      }
      const float* src = nullptr;
      if (active) src = x.table ? x.table + checked_id(x.ids[r], x.V, x.err) * x.E : x.dense + r * x.lda;
      float v[LT_XP];
#pragma unroll
      for (int k4 = 0; k4 < LT_XP / 4; ++k4) {
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

      uint8_t* xh = x_img + (size_t)slot * 2 * LT_XIMG;
#pragma unroll
      for (int pl = 0; pl < LT_XP / 8; ++pl) {
        if (pl * 8 <= in) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(v[pl * 8 + 2 * e], h0, l0);
            split_bf16(v[pl * 8 + 2 * e + 1], h1, l1);
            hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          }
          const size_t off = (size_t)pl * LT_BPLANE + (size_t)lane * 16;
          *reinterpret_cast<uint4*>(xh + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(xh + LT_XIMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) lt_arrive(&x_full[slot]);
    }
  } else if (warp == LT_MMA_WARP) {
    // ===================== MMA issuer (uniform control flow, one elected lane issues) =====================
    mbar_wait(&bar_w, 0);
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, LT_NSEQ);
    const uint64_t wd0 = smem_desc(smem_u32(w_img), LT_APLANE, 128);
    const uint64_t hd0 = smem_desc(smem_u32(h_img), LT_BPLANE, 128);
    const uint64_t xd0 = smem_desc(smem_u32(x_img), LT_BPLANE, 128);
    // k-steps [ks_lo, ks_hi) of one step's GEMM into the accumulators of parity `par`; `fresh`: first MMA overwrites
    auto issue_part = [&](int par, uint64_t bdesc0, uint32_t bimg, int ks_lo, int ks_hi, int ks_sub, bool fresh,
                          int mt_lo, int mt_hi) {
      // (row tile outermost: interleaving the two accumulators between consecutive MMAs measured slower)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < nmt && mt >= mt_lo && mt < mt_hi) {
          const uint64_t wdm = wd0 + (uint64_t)((uint32_t)mt * 2 * LT_AIMG >> 4);
          const uint32_t tacc = tbase + (uint32_t)(par * 2 + mt) * LT_NSEQ;
          uint32_t acc = fresh ? 0u : 1u;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t wp = wdm + (pass == 1 ? (LT_AIMG >> 4) : 0);   // weights: hi, lo, hi
            const uint64_t bp = bdesc0 + (pass == 2 ? (bimg >> 4) : 0);   // activations: hi, hi, lo
#pragma unroll
            for (int ks = 0; ks < LT_K / 16; ++ks) {
              if (ks >= ks_lo && ks < ks_hi && ((ks_mask >> ks) & 1)) {
                mma_bf16_ss_w(tacc, wp + (uint64_t)((2 * ks) * (LT_APLANE >> 4)),
                              bp + (uint64_t)((2 * (ks - ks_sub)) * (LT_BPLANE >> 4)), idesc, acc, issue);
                acc = 1;
              }
            }
          }
        }
      }
    };
    long long ts_m[4] = {0, 0, 0, 0};
    if (maxlen > 0) {
      { LT_T0(); mbar_wait(&x_full[0], 0); LT_ACC(0); }
      tc_fence_after();
      issue_part(0, xd0, LT_XIMG, 0, LT_XP / 16, 0, true, 0, 2);
    }
    for (int step = 0; step < maxlen; ++step) {
      const int par = step & 1, slot = step % LT_XS;
      // The h part is software-pipelined against the previous step's epilogue.  K half kh of h_{step-1} (units
      // 32 kh .. 32 kh + 31) is produced by the epilogue warps of row tile kh, which signal bar_h[kh]: the K-half-0
      // MMAs of BOTH row tiles go out as soon as the tile-0 warps are done (the tile-1 warps are still in their cell
      // update); after bar_h[1] only the K-half-1 MMAs remain, tile 0 first with its own commit, so the tile-0 warps
      // start this step's epilogue while the tile-1 MMAs still run.  Both waits together also guarantee that the
      // epilogue of step-1 has finished reading accumulator par^1 (re-used by the x part below).
      const uint64_t hdp = hd0 + (uint64_t)((uint32_t)par * 2 * LT_HIMG >> 4);
      constexpr int KH0 = LT_XP / 16, KH1 = LT_XP / 16 + LT_HP / 32, KH2 = LT_K / 16;
      { LT_T0(); mbar_wait(&bar_h[0], par); LT_ACC(1); }
      if (step == 100 || step == 101) ts_m[(step - 100) * 2] = clock64();
      tc_fence_after();
      LT_T0();
      issue_part(par, hdp, LT_HIMG, KH0, KH1, KH0, false, 0, 2);
      mbar_wait(&bar_h[1], par);
      tc_fence_after();
      issue_part(par, hdp, LT_HIMG, KH1, KH2, KH0, false, 0, 1);
      mma_commit_w(&bar_acc[0], issue);
      issue_part(par, hdp, LT_HIMG, KH1, KH2, KH0, false, 1, 2);
      mma_commit_w(&bar_acc[1], issue);
      mma_commit_w(&x_empty[slot], issue);
      if (step == 100 || step == 101) ts_m[(step - 100) * 2 + 1] = clock64();
      LT_ACC(2);
      if (step + 1 < maxlen) {
        // x part of the next step, behind the h part in the tensor pipe: runs while the epilogue works
        const int nslot = (step + 1) % LT_XS;
        { LT_T0(); mbar_wait(&x_full[nslot], ((step + 1) / LT_XS) & 1); LT_ACC(0); }
        tc_fence_after();
        issue_part(par ^ 1, xd0 + (uint64_t)((uint32_t)nslot * 2 * LT_XIMG >> 4), LT_XIMG, 0, LT_XP / 16, 0, true, 0, 2);
      }
    }
    if (dbg && lane == 0 && blockIdx.x == 1 && blockIdx.y == 0 && maxlen > 101)
      for (int i = 0; i < 4; ++i) dbg[16 + i] = ts_m[i];
  } else if (warp < LT_EPI_WARPS) {
    // ===================== epilogue warps =====================
    // warp -> (TMEM lane quarter q, row tile mt, half of the 32 sequences); lane -> (t0 = lane % 4, j = lane / 4):
    // unit u = 32 mt + 8 q + j, sequences sl[c] = 16 shalf + {2 t0, 2 t0 + 1, 8 + 2 t0, 9 + 2 t0}
    const int q = warp & 3, mt = (warp >> 2) & 1, shalf = warp >> 3;
    const int t0i = lane & 3, j = lane >> 2;
    const int u = mt * 32 + q * 8 + j;
    const bool uvalid = u < h && mt < nmt;
    // column groups of 8 sequences beyond spc are empty: skip their cells (warp-uniform)
    const bool do01 = shalf * 16 < spc, do23 = shalf * 16 + 8 < spc;
    int sl[4], lk[4];
    float cst[4], hst[4];
    float* optr[4];   // memory-bank position of this step's h for each cell, advanced by one time step per step
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      sl[c] = shalf * 16 + (c >> 1) * 8 + 2 * t0i + (c & 1);
      lk[c] = slen[sl[c]];
      cst[c] = 0.f, hst[c] = 0.f;
      optr[c] = out + ((size_t)(s0 + sl[c]) * L + (dir ? max(lk[c] - 1, 0) : 0)) * Hout + dir * h + u;
    }
    int rowi[4] = {0, 0, 0, 0};   // SAVE: (sequence, time) row of this step for each cell
    if (SAVE) {
#pragma unroll
      for (int c = 0; c < 4; ++c) rowi[c] = (s0 + sl[c]) * L + (dir ? max(lk[c] - 1, 0) : 0);
    }
    const ptrdiff_t ostep = dir ? -(ptrdiff_t)Hout : (ptrdiff_t)Hout;
    const uint32_t tq = tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * LT_NSEQ + shalf * 16);
    const uint32_t hoff = (uint32_t)(u >> 3) * LT_BPLANE + (uint32_t)(u & 7) * 2;
    long long ts_e[4] = {0, 0, 0, 0};
    if (lane == 0 && maxlen > 0) lt_arrive(&bar_h[mt]);   // h_0 = 0 is already in operand buffer 0
    for (int step = 0; step < maxlen; ++step) {
      const int par = step & 1;
      { LT_T0(); mbar_wait(&bar_acc[mt], par); if (warp == 0) LT_ACC(3); }
      if (step == 100) ts_e[0] = clock64();
      tc_fence_after();
      LT_T0();
      float hv[4] = {0.f, 0.f, 0.f, 0.f};
      if (mt < nmt && do01) {
        float ga[8], gb[8];  // ga: gates i (0,1,4,5) and f (2,3,6,7); gb: g and o
        const uint32_t ta = tq + (uint32_t)(par * 2 * LT_NSEQ);
        lt_tmem_ld_16x256b_x2(ta, ga);
        lt_tmem_ld_16x256b_x2(ta + (16u << 16), gb);
        tmem_ld_wait();
        if (step == 100) ts_e[1] = clock64();
        tc_fence_before();
        uint8_t* hn = h_img + (size_t)(par ^ 1) * 2 * LT_HIMG + hoff;  // next step's h operand
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c >= 2 && !do23) continue;
          const int r = (c >> 1) * 4 + (c & 1);
          float cn, hn_v;
          const bool act = step < lk[c];
          if (SAVE) {
            float gi, gf, gg, go;
            lt_cell_save(ga[r], ga[r + 2], gb[r], gb[r + 2], cst[c], cn, hn_v, gi, gf, gg, go);
            if (act && uvalid) {
              float* g = gates_out + (size_t)rowi[c] * (4 * Hout) + dir * 4 * h + u;
              g[0] = gi, g[h] = gf, g[2 * h] = gg, g[3 * h] = go;
              cseq_out[(size_t)rowi[c] * Hout + dir * h + u] = cn;
            }
            rowi[c] += dir ? -1 : 1;
          } else {
            lt_cell(ga[r], ga[r + 2], gb[r], gb[r + 2], cst[c], cn, hn_v);
          }
          cst[c] = act ? cn : cst[c];
          hst[c] = act ? hn_v : hst[c];
          hv[c] = hn_v;
          uint32_t hi, lo;
          split_bf16_alu(hst[c], hi, lo);
          *reinterpret_cast<uint16_t*>(hn + sl[c] * 16) = (uint16_t)hi;
          *reinterpret_cast<uint16_t*>(hn + LT_HIMG + sl[c] * 16) = (uint16_t)lo;
        }
      }
      if (step == 100) ts_e[2] = clock64();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && step + 1 < maxlen) lt_arrive(&bar_h[mt]);
      if (step == 100) ts_e[3] = clock64();
      if (warp == 0) LT_ACC(4);
      // memory bank (fp32), off the critical path: 8 consecutive units x 4 sequences per warp store
      if (uvalid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (step < lk[c]) *optr[c] = hv[c];
          optr[c] += ostep;
        }
      }
      if (warp == 0) LT_ACC(5);
    }
    if (dbg && lane == 0 && blockIdx.x == 1 && blockIdx.y == 0 && maxlen > 101)
      for (int i = 0; i < 4; ++i) dbg[24 + warp * 4 + i] = ts_e[i];
    if (uvalid && (h_n || c_n)) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int s = sl[c];
        if (s0 + s >= n || s >= spc) continue;
        if (h_n) h_n[((size_t)dir * n + s0 + s) * h + u] = hst[c];
        if (c_n) c_n[((size_t)dir * n + s0 + s) * h + u] = cst[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && tid == 0 && blockIdx.x == 1 && blockIdx.y == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
    dbg[8] += clock64(), dbg[9] += (long long)ns;
  }
  if (warp == 0) tmem_dealloc(tbase, LT_TCOLS);
}

// Sequences per CTA: the per-step cost is latency + the cell updates of one SM (MUFU / issue bound), not the MMAs
// (an N = 32 MMA costs the same as a narrower one), so use the fewest sequences per CTA that still fit one wave.
int g_lstm_spc_min = 8;   // process-wide floor of the sequences per CTA (tuning knob, tools/pipe_tune.py)
int lstm_tc_seqs_per_cta(int n, int dirs, int min_spc) {
  if (min_spc < g_lstm_spc_min) min_spc = g_lstm_spc_min;
  for (int c = min_spc; c < LT_NSEQ; c += 8)
    if ((int64_t)((n + c - 1) / c) * dirs <= kSMs) return c;
  return LT_NSEQ;
}
int lstm_tc_ctas(int n, int dirs, int min_spc) {
  const int spc = lstm_tc_seqs_per_cta(n, dirs, min_spc);
  return ((n + spc - 1) / spc) * dirs;
}

int32_t lstm_tc_run(const LstmTcPack& p, const float* bias, const GemmA& x, const int64_t* len, int n, int L,
                    float* out, float* h_n, float* c_n, int* err, cudaStream_t s, const char* rec_name, const uint8_t* ximg,
                    int min_spc, float* gates_out, float* cseq_out) {
  if (n <= 0) return CAIR_OK;
  if (rec_name) prof_mark(rec_name, s);
  uint32_t ks_mask = 0;
  for (int ks = 0; ks < LT_K / 16; ++ks) {
    const int k0 = 16 * ks, k1 = k0 + 16;
    const bool x_part = k0 < p.in + 1;  // + the bias column
    const bool h_part = k1 > LT_XP && k0 < LT_XP + p.h;
    if (x_part || h_part) ks_mask |= 1u << ks;
  }
  const size_t smem = (size_t)4 * LT_AIMG + 4 * LT_HIMG + 2 * LT_XS * LT_XIMG;
  const bool save = gates_out != nullptr && cseq_out != nullptr;
  if (save && (int64_t)n * L >= ((int64_t)1 << 31)) return fail(CAIR_ERR_UNSUPPORTED, "lstm_tc: too many rows for the training outputs");
  if (save)
    CAIR_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    CAIR_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // Sequences per CTA: the per-step cost is latency + the cell updates of one SM (MUFU / issue bound), not the MMAs
  // (an N = 32 MMA costs the same as a narrower one), so use the fewest sequences per CTA that still fit one wave.
  const int spc = lstm_tc_seqs_per_cta(n, p.dirs, min_spc);
  dim3 grid((n + spc - 1) / spc, p.dirs);
  if (save)
    CAIR_LAUNCH(lstm_tc_kernel<true>, grid, LT_THREADS, smem, s, x, p.wimg, bias, len, n, L, p.in, p.h, p.dirs, ks_mask, spc, out,
                h_n, c_n, err, g_lstm_dbg, x.table ? ximg : nullptr, gates_out, cseq_out);
  else
    CAIR_LAUNCH(lstm_tc_kernel<false>, grid, LT_THREADS, smem, s, x, p.wimg, bias, len, n, L, p.in, p.h, p.dirs, ks_mask, spc, out,
                h_n, c_n, err, g_lstm_dbg, x.table ? ximg : nullptr, (float*)nullptr, (float*)nullptr);
  return CAIR_OK;
}

}  // namespace cair
