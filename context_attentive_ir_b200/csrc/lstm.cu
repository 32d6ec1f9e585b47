// RNNEncoder.forward for one (bi)LSTM or (bi)GRU layer (neuroir/encoders/rnn_encoder.py:62-141), fp32 CUDA cores.
// This is the general-shape path (any hidden size); lstm_tc.cu is the tensor-core path for LSTM with h <= 64.
//
// Two phases:
//   1. pre-gates  P[n*L, dirs*G] = X W_ih^T + bias   (G = 4h LSTM / 3h GRU) - one GEMM over all time steps
//      (X optionally gathered from an embedding table inside the GEMM; tcgen05 GEMM for large projections);
//   2. a persistent recurrence kernel: one CTA owns a tile of TS sequences of one direction for all of their
//      steps; W_hh^T lives in shared memory (or is streamed from L2 when it does not fit), the state never leaves
//      the SM, only h_t is stored to the memory bank.
// Packed-sequence semantics without sorting: sequence s runs exactly len[s] steps, the reverse direction starts
// at its own last token, bank rows t >= len[s] are written as zeros (pad_packed_sequence + rnn_encoder.py:135-139).
// Gate orders are torch's: LSTM i,f,g,o; GRU r,z,n with  n = tanh(W_in x + b_in + r * (W_hn h + b_hn)).
#include "common.cuh"

namespace cair {

constexpr int REC_THREADS = 256;

__global__ void lstm_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                 const float* __restrict__ b_ih, const float* __restrict__ b_hh, int in, int h, int gates,
                                 float* __restrict__ o_ih, float* __restrict__ o_bias, float* __restrict__ o_hh_t,
                                 float* __restrict__ o_bhn) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int G = gates * h;
  if (i < (int64_t)G * in) o_ih[i] = w_ih[i];
  if (i < G) {
    // LSTM: both biases fold into the pre-gates.  GRU: b_hn must stay inside r * (W_hn h + b_hn).
    const bool gru_n = gates == 3 && i >= 2 * h;
    o_bias[i] = gru_n ? b_ih[i] : b_ih[i] + b_hh[i];
    if (gru_n) o_bhn[i - 2 * h] = b_hh[i];
  }
  if (i < (int64_t)G * h) {
    int r = (int)(i / h), k = (int)(i % h);
    o_hh_t[(int64_t)k * G + r] = w_hh[i];
  }
}

int32_t lstm_pack(Owned& own, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, int in, int h, LstmPack* out,
                  cudaStream_t s, int rnn_type) {
  if (!fwd || !fwd->w_ih || !fwd->w_hh || !fwd->b_ih || !fwd->b_hh) return fail(CAIR_ERR_BAD_ARG, "rnn: null weights");
  if (in <= 0 || h <= 0) return fail(CAIR_ERR_BAD_ARG, "rnn: bad sizes");
  const int gates = rnn_type == CAIR_RNN_GRU ? 3 : 4;
  int dirs = rev ? 2 : 1, G = gates * h;
  out->in = in, out->h = h, out->dirs = dirs, out->gates = gates;
  CAIR_CUDA(own.alloc(&out->w_ih, (size_t)dirs * G * in));
  CAIR_CUDA(own.alloc(&out->bias, (size_t)dirs * G));
  CAIR_CUDA(own.alloc(&out->w_hh_t, (size_t)dirs * G * h));
  CAIR_CUDA(own.alloc(&out->b_hn, (size_t)dirs * h));
  for (int d = 0; d < dirs; ++d) {
    const cair_lstm_dir* w = d ? rev : fwd;
    if (!w->w_ih || !w->w_hh || !w->b_ih || !w->b_hh) return fail(CAIR_ERR_BAD_ARG, "rnn: null weights");
    int64_t n = (int64_t)G * (in > h ? in : h);
    CAIR_LAUNCH(lstm_pack_kernel, (unsigned)((n + 255) / 256), 256, 0, s, w->w_ih, w->w_hh, w->b_ih, w->b_hh, in, h,
                gates, out->w_ih + (size_t)d * G * in, out->bias + (size_t)d * G, out->w_hh_t + (size_t)d * G * h,
                out->b_hn + (size_t)d * h);
  }
  if (gates == 4 && (size_t)h * G * sizeof(float) > 200 * 1024) {  // stepwise path (few sequences, W_hh larger than smem)
    CAIR_CUDA(own.alloc(&out->w_hh, (size_t)dirs * G * h));
    for (int d = 0; d < dirs; ++d)
      CAIR_CUDA(cudaMemcpyAsync(out->w_hh + (size_t)d * G * h, (d ? rev : fwd)->w_hh, (size_t)G * h * sizeof(float),
                                cudaMemcpyDeviceToDevice, s));
  }
  if ((size_t)dirs * G * in >= 64 * 1024)  // big input projections (CARS: 1024 x 300) go to the tensor cores
    CAIR_TRY(gemm_tc_pack(own, out->w_ih, dirs * G, in, &out->w_ih_tc, s));
  return CAIR_OK;
}

// Few sequences with a recurrent matrix that does not fit in shared memory (CARS session encoders: 32 sessions, h = 512):
// one persistent CTA per 8 sequences would stream 4 MB of W_hh from L2 per step on 4 SMs.  Instead every step is a
// [n, h] x [h, 4h] GEMM over the whole GPU + a cell-update kernel (2 launches per step and direction).
static inline bool lstm_stepwise(const LstmPack& p, int64_t n) { return p.gates == 4 && p.w_hh != nullptr && n <= 64; }
size_t lstm_workspace_floats(const LstmPack& p, int64_t n, int L) {
  size_t f = (size_t)n * L * p.dirs * p.gates * p.h;
  if (lstm_stepwise(p, n)) f += (size_t)p.dirs * n * (p.gates * p.h + 2 * p.h);   // gate scratch + h and c state
  return f;
}

// One step of the stepwise path: gates = pre[:, t] + hprev W_hh^T (tmp), LSTM cell, state and memory bank update.
__global__ void lstm_cell_step_kernel(const float* __restrict__ tmp, const float* __restrict__ pre,
                                      const int64_t* __restrict__ len, int n, int L, int h, int dirs, int dir, int step,
                                      float* __restrict__ hprev, float* __restrict__ cst, float* __restrict__ out,
                                      float* __restrict__ h_n, float* __restrict__ c_n, int* err, float* __restrict__ c_seq) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n * h) return;
  const int s = (int)(i / h), u = (int)(i - (int64_t)s * h);
  const int G = 4 * h, PG = dirs * G, Hout = dirs * h;
  int64_t ll = len[s];
  if (ll < 1 || ll > L) {
    if (step == 0 && u == 0) atomicOr(err, ERRF_BAD_LENGTH);
    ll = ll < 1 ? 1 : L;
  }
  const int l = (int)ll;
  if (step < l) {
    const int t = dir ? l - 1 - step : step;
    const float* g = tmp + (size_t)s * G;
    const float* pg = pre + ((size_t)s * L + t) * PG + (size_t)dir * G;
    const float ig = sigmoid_f(g[u] + pg[u]), fg = sigmoid_f(g[h + u] + pg[h + u]);
    const float gg = tanhf(g[2 * h + u] + pg[2 * h + u]), og = sigmoid_f(g[3 * h + u] + pg[3 * h + u]);
    const float c = fg * cst[i] + ig * gg;
    const float hv = og * tanhf(c);
    cst[i] = c;
    hprev[i] = hv;
    out[((size_t)s * L + t) * Hout + dir * h + u] = hv;
    if (c_seq) c_seq[((size_t)s * L + t) * Hout + dir * h + u] = c;
  }
  if (step == L - 1) {
    if (h_n) h_n[((size_t)dir * n + s) * h + u] = hprev[i];
    if (c_n) c_n[((size_t)dir * n + s) * h + u] = cst[i];
  }
}

// smem: [W_hh^T: h*G floats if WSMEM] [hprev: TS*hp] [c (LSTM) | pre_n (GRU): TS*h] [gates: TS*G] ; hp = h rounded up to 4
// TS = sequences per CTA.
template <bool WSMEM, bool GRU, int TS>
__global__ void __launch_bounds__(REC_THREADS) rnn_rec_kernel(const float* __restrict__ pre, const float* __restrict__ w_hh_t,
                                                              const float* __restrict__ b_hn_all,
                                                              const int64_t* __restrict__ len, int n, int L, int h, int dirs,
                                                              float* __restrict__ out, float* __restrict__ h_n,
                                                              float* __restrict__ c_n, int* err, float* __restrict__ c_seq) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NG = GRU ? 3 : 4;
  const int G = NG * h, hp = (h + 3) & ~3;
  const int dir = blockIdx.y;
  const int s0 = blockIdx.x * TS;
  const int tid = threadIdx.x;
  float* wsm = smem;
  float* hprev = smem + (WSMEM ? (size_t)h * G : 0);
  float* cst = hprev + TS * hp;   // LSTM: cell state; GRU: the n-gate pre-activation W_in x + b_in of this step
  float* gates = cst + TS * h;
  __shared__ int slen[TS];
  __shared__ int smaxlen;

  const float* wt = w_hh_t + (size_t)dir * h * G;
  const float* b_hn = b_hn_all + (size_t)dir * h;
  if (WSMEM)
    for (int i = tid; i < h * G; i += REC_THREADS) wsm[i] = wt[i];
  const float* W = WSMEM ? wsm : wt;
  for (int i = tid; i < TS * hp; i += REC_THREADS) hprev[i] = 0.f;
  for (int i = tid; i < TS * h; i += REC_THREADS) cst[i] = 0.f;
  if (tid < TS) {
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
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int s = 0; s < TS; ++s) m = max(m, slen[s]);
    smaxlen = m;
  }
  // zero the pad rows of the memory bank (this direction's half)
  const int Hout = dirs * h;
  for (int s = 0; s < TS; ++s) {
    if (s0 + s >= n) break;
    int npad = (L - slen[s]) * h;
    float* o = out + ((size_t)(s0 + s) * L + slen[s]) * Hout + dir * h;
    for (int i = tid; i < npad; i += REC_THREADS) o[(size_t)(i / h) * Hout + (i % h)] = 0.f;
  }
  __syncthreads();
  const int maxlen = smaxlen;
  const int PG = dirs * G;  // pre-gate row width

  for (int step = 0; step < maxlen; ++step) {
    // gate rows: r = tid, tid+256, ...
    for (int r = tid; r < G; r += REC_THREADS) {
      const bool n_row = GRU && r >= 2 * h;   // GRU candidate gate: keep W_in x + b_in apart from r * (W_hn h + b_hn)
      float acc[TS];
#pragma unroll
      for (int s = 0; s < TS; ++s) {
        int l = slen[s];
        float v = 0.f;
        if (step < l) {
          int t = dir ? l - 1 - step : step;
          v = pre[((size_t)(s0 + s) * L + t) * PG + dir * G + r];
        }
        if (n_row) {
          cst[s * h + (r - 2 * h)] = v;
          v = b_hn[r - 2 * h];
        }
        acc[s] = v;
      }
      int k = 0;
#pragma unroll 4
      for (; k + 4 <= h; k += 4) {  // partially unrolled: 16 independent weight loads in flight when W is streamed from L2
        float w0 = W[(size_t)(k + 0) * G + r], w1 = W[(size_t)(k + 1) * G + r];
        float w2 = W[(size_t)(k + 2) * G + r], w3 = W[(size_t)(k + 3) * G + r];
#pragma unroll
        for (int s = 0; s < TS; ++s) {
          float4 hv = *reinterpret_cast<const float4*>(&hprev[s * hp + k]);
          acc[s] = fmaf(w0, hv.x, acc[s]);
          acc[s] = fmaf(w1, hv.y, acc[s]);
          acc[s] = fmaf(w2, hv.z, acc[s]);
          acc[s] = fmaf(w3, hv.w, acc[s]);
        }
      }
      for (; k < h; ++k) {
        float w0 = W[(size_t)k * G + r];
#pragma unroll
        for (int s = 0; s < TS; ++s) acc[s] = fmaf(w0, hprev[s * hp + k], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < TS; ++s) gates[s * G + r] = acc[s];
    }
    __syncthreads();
    for (int i = tid; i < TS * h; i += REC_THREADS) {
      int s = i / h, u = i - s * h;
      int l = slen[s];
      if (step < l) {
        const float* g = gates + s * G;
        float hv;
        if (GRU) {
          const float rg = sigmoid_f(g[u]), zg = sigmoid_f(g[h + u]);
          const float ng = tanhf(cst[i] + rg * g[2 * h + u]);
          hv = (1.0f - zg) * ng + zg * hprev[s * hp + u];
        } else {
          float ig = sigmoid_f(g[u]), fg = sigmoid_f(g[h + u]);
          float gg = tanhf(g[2 * h + u]), og = sigmoid_f(g[3 * h + u]);
          float c = fg * cst[i] + ig * gg;
          hv = og * tanhf(c);
          cst[i] = c;
          if (c_seq) c_seq[((size_t)(s0 + s) * L + (dir ? l - 1 - step : step)) * Hout + dir * h + u] = c;
        }
        hprev[s * hp + u] = hv;
        int t = dir ? l - 1 - step : step;
        out[((size_t)(s0 + s) * L + t) * Hout + dir * h + u] = hv;
      }
    }
    __syncthreads();
  }
  if (h_n || c_n)
    for (int i = tid; i < TS * h; i += REC_THREADS) {
      int s = i / h, u = i - s * h;
      if (s0 + s >= n) continue;
      if (h_n) h_n[((size_t)dir * n + s0 + s) * h + u] = hprev[s * hp + u];
      if (c_n && !GRU) c_n[((size_t)dir * n + s0 + s) * h + u] = cst[i];
    }
}

// ---- LSTM recurrence split over a 2-CTA thread-block cluster (W_hh between 200 and 400 KB, e.g. CARS h = 128/dir) ----
// W_hh^T (h x 4h fp32) does not fit one SM's shared memory, and streaming it from L2 on every step makes the
// recurrence L2-bandwidth bound (560 CTAs x 200 steps x 256 KB at cfg4).  Here the two CTAs of a cluster each keep the
// gate columns of HALF the hidden units resident (h x 4 (h/2) floats = 128 KB), compute the gates and the cell update of
// their units for the same TS sequences, and write the new h values into BOTH CTAs' next-step buffers (distributed
// shared memory); one cluster barrier per step.  smem: W half | hprev [2 parities][TS][h] | c [TS][hh] | gates [TS][4hh].
constexpr int REC2_TS = 16;   // sequences per cluster: amortises the per-step fixed costs (pre-gate loads, barriers) over more work

__device__ __forceinline__ uint32_t rec2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rec2_st_peer(float* local_ptr, uint32_t peer_rank, float v) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(rec2_smem_u32(local_ptr)), "r"(peer_rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
__device__ __forceinline__ void rec2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(REC_THREADS, 1)
    lstm_rec2_kernel(const float* __restrict__ pre, const float* __restrict__ w_hh_t, const int64_t* __restrict__ len, int n,
                     int L, int h, int dirs, float* __restrict__ out, float* __restrict__ h_n, float* __restrict__ c_n,
                     int* err) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TS = REC2_TS;
  const int hh = h >> 1, G = 4 * h, GH = 4 * hh;     // GH = gate rows of this CTA
  const int half = blockIdx.x & 1;                    // rank in the cluster = which half of the hidden units
  const int s0 = (blockIdx.x >> 1) * TS;
  const int dir = blockIdx.y;
  const int tid = threadIdx.x;
  float* wsm = smem;                                  // [h][GH]: row k, local gate row r = type*hh + ul
  float* hprev = wsm + (size_t)h * GH;                // [2][TS][h]
  float* cst = hprev + 2 * TS * h;                    // [TS][hh]
  float* gates = cst + TS * hh;                       // [TS][GH]
  __shared__ int slen[TS];
  __shared__ int smaxlen;

  const float* wt = w_hh_t + (size_t)dir * h * G;
  for (int i = tid; i < h * GH; i += REC_THREADS) {
    const int k = i / GH, r = i - k * GH;
    const int type = r / hh, ul = r - type * hh;
    wsm[i] = wt[(size_t)k * G + type * h + half * hh + ul];
  }
  for (int i = tid; i < 2 * TS * h; i += REC_THREADS) hprev[i] = 0.f;
  for (int i = tid; i < TS * hh; i += REC_THREADS) cst[i] = 0.f;
  if (tid < TS) {
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
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int s = 0; s < TS; ++s) m = max(m, slen[s]);
    smaxlen = m;
  }
  // zero the pad rows of the memory bank (this CTA's units of this direction)
  const int Hout = dirs * h;
  for (int s = 0; s < TS; ++s) {
    if (s0 + s >= n) break;
    const int npad = (L - slen[s]) * hh;
    float* o = out + ((size_t)(s0 + s) * L + slen[s]) * Hout + dir * h + half * hh;
    for (int i = tid; i < npad; i += REC_THREADS) o[(size_t)(i / hh) * Hout + (i % hh)] = 0.f;
  }
  __syncthreads();
  rec2_cluster_sync();   // both CTAs have zeroed their state before anyone writes remotely
  const int maxlen = smaxlen;   // identical in both CTAs (same sequences)
  const int PG = dirs * G;

  // pre-gates of this thread's gate row (GH == REC_THREADS: one row per thread) are prefetched one step ahead, so their
  // L2 / HBM latency overlaps the cell update and the cluster barrier of the previous step
  const bool one_row = GH == REC_THREADS;
  float nxt[TS];
  auto load_pre = [&](int step, int r, float* v) {
    const int type = r / hh, ul = r - type * hh;
    const int grow = type * h + half * hh + ul;
#pragma unroll
    for (int s = 0; s < TS; ++s) {
      const int l = slen[s];
      float x = 0.f;
      if (step < l) {
        const int t = dir ? l - 1 - step : step;
        x = pre[((size_t)(s0 + s) * L + t) * PG + dir * G + grow];
      }
      v[s] = x;
    }
  };
  if (one_row && maxlen > 0) load_pre(0, tid, nxt);
  for (int step = 0; step < maxlen; ++step) {
    const float* hp = hprev + (size_t)(step & 1) * TS * h;
    float* hn = hprev + (size_t)((step & 1) ^ 1) * TS * h;
    for (int r = tid; r < GH; r += REC_THREADS) {
      float acc[TS];
      if (one_row) {
#pragma unroll
        for (int s = 0; s < TS; ++s) acc[s] = nxt[s];
        if (step + 1 < maxlen) load_pre(step + 1, r, nxt);
      } else {
        load_pre(step, r, acc);
      }
#pragma unroll 4
      for (int k = 0; k < h; k += 4) {
        const float w0 = wsm[(size_t)(k + 0) * GH + r], w1 = wsm[(size_t)(k + 1) * GH + r];
        const float w2 = wsm[(size_t)(k + 2) * GH + r], w3 = wsm[(size_t)(k + 3) * GH + r];
#pragma unroll
        for (int s = 0; s < TS; ++s) {
          const float4 hv = *reinterpret_cast<const float4*>(&hp[s * h + k]);
          acc[s] = fmaf(w0, hv.x, acc[s]);
          acc[s] = fmaf(w1, hv.y, acc[s]);
          acc[s] = fmaf(w2, hv.z, acc[s]);
          acc[s] = fmaf(w3, hv.w, acc[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < TS; ++s) gates[s * GH + r] = acc[s];
    }
    __syncthreads();
    for (int i = tid; i < TS * hh; i += REC_THREADS) {
      const int s = i / hh, ul = i - s * hh;
      const int l = slen[s];
      const int u = half * hh + ul;
      float hv = hp[s * h + u];   // inactive sequences carry their state forward
      if (step < l) {
        const float* g = gates + s * GH;
        const float ig = sigmoid_f(g[ul]), fg = sigmoid_f(g[hh + ul]);
        const float gg = tanhf(g[2 * hh + ul]), og = sigmoid_f(g[3 * hh + ul]);
        const float c = fg * cst[i] + ig * gg;
        hv = og * tanhf(c);
        cst[i] = c;
        const int t = dir ? l - 1 - step : step;
        out[((size_t)(s0 + s) * L + t) * Hout + dir * h + u] = hv;
      }
      hn[s * h + u] = hv;
      rec2_st_peer(&hn[s * h + u], (uint32_t)(half ^ 1), hv);
    }
    rec2_cluster_sync();   // next step's h is complete in both CTAs; also orders the reuse of `gates` and of the other parity
  }
  if (h_n || c_n) {
    const float* hp = hprev + (size_t)(maxlen & 1) * TS * h;
    for (int i = tid; i < TS * hh; i += REC_THREADS) {
      const int s = i / hh, ul = i - s * hh;
      if (s0 + s >= n) continue;
      const int u = half * hh + ul;
      if (h_n) h_n[((size_t)dir * n + s0 + s) * h + u] = hp[s * h + u];
      if (c_n) c_n[((size_t)dir * n + s0 + s) * h + u] = cst[i];
    }
  }
  rec2_cluster_sync();   // no CTA exits while its peer may still write into its shared memory
}

static int32_t launch_rec2(const LstmPack& p, const float* pre, const int64_t* len, int n, int L, float* out, float* h_n,
                           float* c_n, int* err, cudaStream_t s) {
  const int h = p.h, hh = h / 2;
  const size_t smem = ((size_t)h * 4 * hh + 2 * REC2_TS * h + REC2_TS * hh + REC2_TS * 4 * hh) * sizeof(float);
  CAIR_CUDA(cudaFuncSetAttribute(lstm_rec2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * ((n + REC2_TS - 1) / REC2_TS), p.dirs);
  cfg.blockDim = dim3(REC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_rec2_kernel, pre, (const float*)p.w_hh_t, len, n, L, h, p.dirs, out, h_n, c_n, err);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) return fail(CAIR_ERR_CUDA, "lstm_rec2 launch: %s", cudaGetErrorString(e));
  return CAIR_OK;
}
static inline bool lstm_rec2_usable(const LstmPack& p) {
  const size_t wbytes = (size_t)p.h * 4 * p.h * sizeof(float);
  return p.gates == 4 && (p.h % 8) == 0 && wbytes > 200 * 1024 &&
         ((size_t)p.h * 2 * p.h + 2 * REC2_TS * p.h + REC2_TS * (p.h / 2) + REC2_TS * 2 * p.h) * sizeof(float) <= 220 * 1024;
}

template <bool WSMEM, bool GRU, int TS>
static int32_t launch_rec(const LstmPack& p, const float* pre, const int64_t* len, int n, int L, float* out, float* h_n,
                          float* c_n, int* err, size_t smem, cudaStream_t s, float* c_seq) {
  if (smem > 40 * 1024)  // static smem counts against the 48 KB default limit too
    CAIR_CUDA(cudaFuncSetAttribute(rnn_rec_kernel<WSMEM, GRU, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n + TS - 1) / TS, p.dirs);
  CAIR_LAUNCH((rnn_rec_kernel<WSMEM, GRU, TS>), grid, REC_THREADS, smem, s, pre, p.w_hh_t, p.b_hn, len, n, L, p.h, p.dirs, out,
              h_n, c_n, err, c_seq);
  return CAIR_OK;
}

int32_t lstm_run(const LstmPack& p, const GemmA& x, const int64_t* len, int n, int L, float* out, float* h_n,
                 float* c_n, float* ws_pre, int* err, cudaStream_t s, const char* rec_name, float* c_seq) {
  if (n <= 0) return CAIR_OK;
  if (c_seq && p.gates != 4) return fail(CAIR_ERR_UNSUPPORTED, "rnn: a cell-state sequence exists for LSTM only");
  const int G = p.gates * p.h, PG = p.dirs * G;
  CAIR_TRY(gemm_auto(x, p.w_ih, p.w_ih_tc, p.bias, ws_pre, PG, (int64_t)n * L, PG, p.in, ACT_NONE, s));
  if (rec_name) prof_mark(rec_name, s);
  if (lstm_stepwise(p, n)) {
    const int h = p.h;
    float* scratch = ws_pre + (size_t)n * L * PG;
    float* tmp = scratch;                                   // [dirs][n][G]
    float* hprev = tmp + (size_t)p.dirs * n * G;            // [dirs][n][h]
    float* cst = hprev + (size_t)p.dirs * n * h;            // [dirs][n][h]
    CAIR_CUDA(cudaMemsetAsync(hprev, 0, (size_t)p.dirs * n * 2 * h * sizeof(float), s));
    CAIR_CUDA(cudaMemsetAsync(out, 0, (size_t)n * L * p.dirs * h * sizeof(float), s));   // rows t >= len stay zero
    const unsigned blocks = (unsigned)(((int64_t)n * h + 255) / 256);
    for (int step = 0; step < L; ++step)
      for (int d = 0; d < p.dirs; ++d) {
        float* hp_d = hprev + (size_t)d * n * h;
        float* tmp_d = tmp + (size_t)d * n * G;
        CAIR_TRY(gemm_f32(gemm_dense(hp_d, h), p.w_hh + (size_t)d * G * h, nullptr, tmp_d, G, n, G, h, ACT_NONE, s));
        CAIR_LAUNCH(lstm_cell_step_kernel, blocks, 256, 0, s, tmp_d, ws_pre, len, n, L, h, p.dirs, d, step, hp_d,
                    cst + (size_t)d * n * h, out, h_n, c_n, err, c_seq);
      }
    return CAIR_OK;
  }
  const int hp = (p.h + 3) & ~3;
  const size_t wbytes = (size_t)p.h * G * sizeof(float);
  const size_t state8 = (size_t)(8 * hp + 8 * p.h + 8 * G) * sizeof(float);
  const size_t state16 = 2 * state8;
  const bool wsmem = wbytes + state8 <= 200 * 1024;
  const bool gru = p.gates == 3;
  if (wsmem)
    return gru ? launch_rec<true, true, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8 + wbytes, s, c_seq)
               : launch_rec<true, false, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8 + wbytes, s, c_seq);
  if (lstm_rec2_usable(p) && !c_seq) return launch_rec2(p, ws_pre, len, n, L, out, h_n, c_n, err, s);
  (void)state16;  // TS = 16 measured slower (fewer CTAs in flight); the streamed path keeps 8 sequences per CTA
  return gru ? launch_rec<false, true, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8, s, c_seq)
             : launch_rec<false, false, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8, s, c_seq);
}

}  // namespace cair
