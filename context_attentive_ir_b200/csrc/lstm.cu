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
  if ((size_t)dirs * G * in >= 64 * 1024)  // big input projections (CARS: 1024 x 300) go to the tensor cores
    CAIR_TRY(gemm_tc_pack(own, out->w_ih, dirs * G, in, &out->w_ih_tc, s));
  return CAIR_OK;
}

size_t lstm_workspace_floats(const LstmPack& p, int64_t n, int L) { return (size_t)n * L * p.dirs * p.gates * p.h; }

// smem: [W_hh^T: h*G floats if WSMEM] [hprev: TS*hp] [c (LSTM) | pre_n (GRU): TS*h] [gates: TS*G] ; hp = h rounded up to 4
// TS = sequences per CTA.
template <bool WSMEM, bool GRU, int TS>
__global__ void __launch_bounds__(REC_THREADS) rnn_rec_kernel(const float* __restrict__ pre, const float* __restrict__ w_hh_t,
                                                              const float* __restrict__ b_hn_all,
                                                              const int64_t* __restrict__ len, int n, int L, int h, int dirs,
                                                              float* __restrict__ out, float* __restrict__ h_n,
                                                              float* __restrict__ c_n, int* err) {
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

template <bool WSMEM, bool GRU, int TS>
static int32_t launch_rec(const LstmPack& p, const float* pre, const int64_t* len, int n, int L, float* out, float* h_n,
                          float* c_n, int* err, size_t smem, cudaStream_t s) {
  if (smem > 40 * 1024)  // static smem counts against the 48 KB default limit too
    CAIR_CUDA(cudaFuncSetAttribute(rnn_rec_kernel<WSMEM, GRU, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n + TS - 1) / TS, p.dirs);
  CAIR_LAUNCH((rnn_rec_kernel<WSMEM, GRU, TS>), grid, REC_THREADS, smem, s, pre, p.w_hh_t, p.b_hn, len, n, L, p.h, p.dirs, out,
              h_n, c_n, err);
  return CAIR_OK;
}

int32_t lstm_run(const LstmPack& p, const GemmA& x, const int64_t* len, int n, int L, float* out, float* h_n,
                 float* c_n, float* ws_pre, int* err, cudaStream_t s, const char* rec_name) {
  if (n <= 0) return CAIR_OK;
  const int G = p.gates * p.h, PG = p.dirs * G;
  CAIR_TRY(gemm_auto(x, p.w_ih, p.w_ih_tc, p.bias, ws_pre, PG, (int64_t)n * L, PG, p.in, ACT_NONE, s));
  if (rec_name) prof_mark(rec_name, s);
  const int hp = (p.h + 3) & ~3;
  const size_t wbytes = (size_t)p.h * G * sizeof(float);
  const size_t state8 = (size_t)(8 * hp + 8 * p.h + 8 * G) * sizeof(float);
  const size_t state16 = 2 * state8;
  const bool wsmem = wbytes + state8 <= 200 * 1024;
  const bool gru = p.gates == 3;
  if (wsmem)
    return gru ? launch_rec<true, true, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8 + wbytes, s)
               : launch_rec<true, false, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8 + wbytes, s);
  (void)state16;  // TS = 16 measured slower (fewer CTAs in flight); the streamed path keeps 8 sequences per CTA
  return gru ? launch_rec<false, true, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8, s)
             : launch_rec<false, false, 8>(p, ws_pre, len, n, L, out, h_n, c_n, err, state8, s);
}

}  // namespace cair
