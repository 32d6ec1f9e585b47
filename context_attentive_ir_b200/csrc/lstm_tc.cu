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
// Gate rows are permuted so that TMEM lane quarter q of row tile m holds gate type q (i,f,g,o) of units
// 32m..32m+31: the activation is warp-uniform and bias is a per-thread scalar.
// Warp roles (576 threads): warps 0-15 epilogue (phase 1: tcgen05.ld + sigmoid/tanh -> smem; phase 2: c/h
// update in registers, h written as next step's bf16 hi/lo operand + fp32 memory bank), warp 16: MMA issuer,
// warp 17: gather of x (embedding rows by token id, or dense rows) into a 4-slot operand ring, running up
// to 3 steps ahead so the global-load latency never sits on the recurrence's critical path.
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
constexpr int LT_THREADS = (LT_EPI_WARPS + 2) * 32;
constexpr uint32_t LT_APLANE = 128 * 16;  // weight image: 128 rows per plane
constexpr uint32_t LT_BPLANE = LT_NSEQ * 16;
constexpr uint32_t LT_AIMG = LT_PLANES * LT_APLANE;  // one (row tile, hi|lo) image: 28672 B
constexpr uint32_t LT_BIMG = LT_PLANES * LT_BPLANE;  // one (hi|lo) image: 7168 B

bool lstm_tc_supported(int in, int h) { return in >= 1 && in <= LT_XP && h >= 1 && h <= LT_HP; }

// weight image [dir][row tile][hi|lo][plane][row][8 x bf16]; row = type*32 + l  <->  gate row type*h + (32*tile + l)
__global__ void lstm_tc_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, int in, int h,
                                    uint8_t* __restrict__ img) {
  const int total = 2 * LT_PLANES * 128 * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e = idx & 7, row = (idx >> 3) & 127, pl = (idx >> 10) % LT_PLANES, mt = idx / (LT_PLANES * 1024);
    const int k = pl * 8 + e, type = row >> 5, u = mt * 32 + (row & 31);
    float v = 0.f;
    if (u < h) {
      const int grow = type * h + u;
      if (k < LT_XP) {
        if (k < in) v = w_ih[(size_t)grow * in + k];
      } else if (k - LT_XP < h) {
        v = w_hh[(size_t)grow * h + (k - LT_XP)];
      }
    }
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
    CAIR_LAUNCH(lstm_tc_pack_kernel, 64, 256, 0, s, w->w_ih, w->w_hh, in, h, out->wimg + (size_t)d * 4 * LT_AIMG);
  }
  return CAIR_OK;
}

__device__ __forceinline__ void lt_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void lt_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// fast, accurate-enough activations: ex2.approx (2^-22 rel) + rcp.approx (1 ulp).  No clamping needed:
// exp2 saturates to inf/0 and 1/(1+inf) = 0.  tanh(x) = 2*sigmoid(2x) - 1 shares the same two MUFU ops.
__device__ __forceinline__ float lt_sigmoid(float x) { return __fdividef(1.0f, 1.0f + exp2f(-1.4426950408889634f * x)); }
__device__ __forceinline__ float lt_tanh(float x) { return fmaf(2.0f, lt_sigmoid(2.0f * x), -1.0f); }

// Optional role timing (dbg != nullptr; CTA (0,0), lane 0 of the role's first warp), see tools/lstm_timing.py
#define LT_T0() long long t0_ = dbg ? clock64() : 0
#define LT_ACC(slot) do { if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0) dbg[slot] += clock64() - t0_; } while (0)
long long* g_lstm_dbg = nullptr;

constexpr int LT_XS = 4;                                   // x-operand ring slots (gather runs up to 3 steps ahead)
constexpr uint32_t LT_XIMG = (LT_XP / 8) * LT_BPLANE;     // one (hi|lo) x image: 6 planes, 3072 B
constexpr uint32_t LT_HIMG = (LT_HP / 8) * LT_BPLANE;     // one (hi|lo) h image: 8 planes, 4096 B

// smem: W image (4 x LT_AIMG) | h operand [2 parities][hi|lo] | x operand ring [LT_XS][hi|lo] | gsm [4][32][64] f32
__global__ void __launch_bounds__(LT_THREADS, 1)
    lstm_tc_kernel(GemmA x, const uint8_t* __restrict__ wimg_all, const float* __restrict__ bias_all,
                   const int64_t* __restrict__ len, int n, int L, int in, int h, int dirs, uint32_t ks_mask,
                   float* __restrict__ out, float* __restrict__ h_n, float* __restrict__ c_n, int* err,
                   long long* __restrict__ dbg) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ uint64_t bar_w, bar_h, bar_acc, x_full[LT_XS], x_empty[LT_XS];
  __shared__ uint32_t tmem_slot;
  __shared__ int slen[LT_NSEQ];
  __shared__ int smaxlen;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dir = blockIdx.y, s0 = blockIdx.x * LT_NSEQ;
  const int nmt = (h + 31) / 32;  // row tiles in use (1 or 2)
  uint8_t* w_img = smraw;
  uint8_t* h_img = w_img + 4 * LT_AIMG;
  uint8_t* x_img = h_img + 4 * LT_HIMG;
  float* gsm = reinterpret_cast<float*>(x_img + 2 * LT_XS * LT_XIMG);
  const float* bias = bias_all + (size_t)dir * 4 * h;
  const int Hout = dirs * h;

  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  if (tid == 32) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_h, LT_EPI_WARPS);  // epilogue warps: h_t written as next step's operand
    mbar_init(&bar_acc, 1);
    for (int i = 0; i < LT_XS; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (tid < LT_NSEQ) {
    int s = s0 + tid, l = 0;
    if (s < n) {
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
  // zero the pad rows of the memory bank (this direction's half)
  for (int s = 0; s < LT_NSEQ; ++s) {
    if (s0 + s >= n) break;
    const int npad = (L - slen[s]) * h;
    float* o = out + ((size_t)(s0 + s) * L + slen[s]) * Hout + dir * h;
    for (int i = tid; i < npad; i += LT_THREADS) o[(size_t)(i / h) * Hout + (i % h)] = 0.f;
  }
  __syncthreads();
  const int maxlen = smaxlen;
  const uint32_t tbase = tmem_slot;

  if (warp == LT_EPI_WARPS + 1) {
    // ===================== x gather: embedding rows (or dense rows) -> hi/lo bf16 ring, runs ahead =====================
    const int myl = slen[lane];   // lane <-> sequence row
    const bool vec = (in & 3) == 0 && (x.table ? ((x.E & 3) == 0) : ((x.lda & 3) == 0));
    for (int step = 0; step < maxlen; ++step) {
      const int slot = step % LT_XS;
      { LT_T0(); mbar_wait_relaxed(&x_empty[slot], ((step / LT_XS) & 1) ^ 1); LT_ACC(7); }
      const bool active = step < myl;
      const int t = dir ? myl - 1 - step : step;
      const int64_t r = (int64_t)(s0 + lane) * L + (active ? t : 0);
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
        v[4 * k4] = f4.x, v[4 * k4 + 1] = f4.y, v[4 * k4 + 2] = f4.z, v[4 * k4 + 3] = f4.w;
      }
      uint8_t* xh = x_img + (size_t)slot * 2 * LT_XIMG;
#pragma unroll
      for (int pl = 0; pl < LT_XP / 8; ++pl) {
        if (pl * 8 < in) {
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
  } else if (warp == LT_EPI_WARPS) {
    // ===================== MMA issuer (uniform control flow, one elected lane issues) =====================
    mbar_wait(&bar_w, 0);
    const uint32_t issue = elect_one();
    const uint32_t idesc = idesc_bf16_f32(128, LT_NSEQ);
    const uint64_t wd0 = smem_desc(smem_u32(w_img), LT_APLANE, 128);
    const uint64_t hd0 = smem_desc(smem_u32(h_img), LT_BPLANE, 128);
    const uint64_t xd0 = smem_desc(smem_u32(x_img), LT_BPLANE, 128);
    for (int step = 0; step < maxlen; ++step) {
      const int par = step & 1, slot = step % LT_XS;
      { LT_T0(); mbar_wait(&x_full[slot], (step / LT_XS) & 1); LT_ACC(0); }
      { LT_T0(); mbar_wait(&bar_h, par); LT_ACC(1); }
      tc_fence_after();
      LT_T0();
      const uint64_t hdp = hd0 + (uint64_t)((uint32_t)par * 2 * LT_HIMG >> 4);
      const uint64_t xdp = xd0 + (uint64_t)((uint32_t)slot * 2 * LT_XIMG >> 4);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < nmt) {
          const uint64_t wdm = wd0 + (uint64_t)((uint32_t)mt * 2 * LT_AIMG >> 4);
          const uint32_t tacc = tbase + (uint32_t)mt * LT_NSEQ;
          uint32_t acc = 0;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t wp = wdm + (pass == 1 ? (LT_AIMG >> 4) : 0);   // weights: hi, lo, hi
            const uint64_t xp = xdp + (pass == 2 ? (LT_XIMG >> 4) : 0);   // activations: hi, hi, lo
            const uint64_t hp = hdp + (pass == 2 ? (LT_HIMG >> 4) : 0);
#pragma unroll
            for (int ks = 0; ks < LT_K / 16; ++ks) {
              if ((ks_mask >> ks) & 1) {
                const uint64_t bd = (ks < LT_XP / 16) ? xp + (uint64_t)((2 * ks) * (LT_BPLANE >> 4))
                                                      : hp + (uint64_t)((2 * (ks - LT_XP / 16)) * (LT_BPLANE >> 4));
                mma_bf16_ss_w(tacc, wp + (uint64_t)((2 * ks) * (LT_APLANE >> 4)), bd, idesc, acc, issue);
                acc = 1;
              }
            }
          }
        }
      }
      mma_commit_w(&bar_acc, issue);
      mma_commit_w(&x_empty[slot], issue);
      LT_ACC(2);
    }
  } else {
    // ===================== epilogue warps =====================
    // phase 1: warp -> (row tile mt, gate type = TMEM lane quarter, half of the 32 sequences); lane -> unit
    const int type = warp & 3, mt = (warp >> 2) & 1, shalf = warp >> 3;
    const int u1 = mt * 32 + lane;
    const float bias1 = (u1 < h) ? bias[type * h + u1] : 0.f;
    const float pre = (type == 2) ? 2.0f : 1.0f;     // tanh(x) = 2*sigmoid(2x) - 1 for the cell gate
    const float post_a = (type == 2) ? 2.0f : 1.0f, post_b = (type == 2) ? -1.0f : 0.0f;
    // phase 2: thread -> unit u2, sequences s = sg + 8k (k < 4)
    const int u2 = tid & 63, sg = tid >> 6;
    float cst[4], hst[4];
    int lk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cst[k] = 0.f, hst[k] = 0.f, lk[k] = slen[sg + 8 * k];
    if (lane == 0 && maxlen > 0) lt_arrive(&bar_h);   // h_0 = 0 is already in operand buffer 0
    for (int step = 0; step < maxlen; ++step) {
      const int par = step & 1;
      { LT_T0(); mbar_wait(&bar_acc, par); if (warp == 0) LT_ACC(3); }
      tc_fence_after();
      LT_T0();
      // ---- phase 1: activation of one gate row for 16 sequences ----
      if (mt < nmt) {
        float v[16];
        tmem_ld16(tbase + ((uint32_t)(type * 32) << 16) + (uint32_t)(mt * LT_NSEQ + shalf * 16), v);
        tmem_ld_wait();
        if (u1 < h) {
#pragma unroll
          for (int s = 0; s < 16; ++s) {
            const float sgm = lt_sigmoid(pre * (v[s] + bias1));
            gsm[(type * 32 + shalf * 16 + s) * 64 + u1] = fmaf(post_a, sgm, post_b);
          }
        }
      }
      tc_fence_before();
      if (warp == 0) LT_ACC(4);
      lt_named_bar(1, LT_EPI_WARPS * 32);
      if (warp == 0) LT_ACC(5);
      // ---- phase 2: state update, branch-free so the 4 items interleave ----
      uint8_t* hn = h_img + (size_t)(par ^ 1) * 2 * LT_HIMG;  // next step's h operand
      if (u2 < h) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int s = sg + 8 * k;
          const bool act = step < lk[k];
          const float ig = gsm[(0 * 32 + s) * 64 + u2], fg = gsm[(1 * 32 + s) * 64 + u2];
          const float gg = gsm[(2 * 32 + s) * 64 + u2], og = gsm[(3 * 32 + s) * 64 + u2];
          const float c = fmaf(fg, cst[k], ig * gg);
          const float hv = og * lt_tanh(c);
          cst[k] = act ? c : cst[k];
          hst[k] = act ? hv : hst[k];
          if (act) {
            const int t = dir ? lk[k] - 1 - step : step;
            out[((size_t)(s0 + s) * L + t) * Hout + dir * h + u2] = hv;
          }
          __nv_bfloat16 hi, lo;
          split_bf16(hst[k], hi, lo);
          const size_t off = (size_t)(u2 >> 3) * LT_BPLANE + (size_t)s * 16 + (u2 & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(hn + off) = hi;
          *reinterpret_cast<__nv_bfloat16*>(hn + LT_HIMG + off) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (warp == 0) LT_ACC(6);
      if (lane == 0 && step + 1 < maxlen) lt_arrive(&bar_h);
    }
    if (u2 < h && (h_n || c_n)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int s = sg + 8 * k;
        if (s0 + s >= n) continue;
        if (h_n) h_n[((size_t)dir * n + s0 + s) * h + u2] = hst[k];
        if (c_n) c_n[((size_t)dir * n + s0 + s) * h + u2] = cst[k];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

int32_t lstm_tc_run(const LstmTcPack& p, const float* bias, const GemmA& x, const int64_t* len, int n, int L,
                    float* out, float* h_n, float* c_n, int* err, cudaStream_t s, const char* rec_name) {
  if (n <= 0) return CAIR_OK;
  if (rec_name) prof_mark(rec_name, s);
  uint32_t ks_mask = 0;
  for (int ks = 0; ks < LT_K / 16; ++ks) {
    const int k0 = 16 * ks, k1 = k0 + 16;
    const bool x_part = k0 < p.in;
    const bool h_part = k1 > LT_XP && k0 < LT_XP + p.h;
    if (x_part || h_part) ks_mask |= 1u << ks;
  }
  const size_t smem = (size_t)4 * LT_AIMG + 4 * LT_HIMG + 2 * LT_XS * LT_XIMG + (size_t)4 * 32 * 64 * sizeof(float);
  CAIR_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n + LT_NSEQ - 1) / LT_NSEQ, p.dirs);
  CAIR_LAUNCH(lstm_tc_kernel, grid, LT_THREADS, smem, s, x, p.wimg, bias, len, n, L, p.in, p.h, p.dirs, ks_mask, out,
              h_n, c_n, err, g_lstm_dbg);
  return CAIR_OK;
}

}  // namespace cair
