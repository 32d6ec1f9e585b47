// MNSRF ranking path (neuroir/multitask/mnsrf.py:61-162: encode + rank_document).
//   queries [B,S,Lq] -> embedder -> (Bi)RNN encoder -> max over time (pad rows of the bank are zero and take part, :78 /
//   :235-237) = memory_bank [B,S,Hq]; unidirectional session RNN over the S pooled queries (:91-102) = session_bank
//   [B,S,Hs]; per query i: tanh(projection([q_i ; i == 0 ? 0 : session_bank_i])) (:143-150; note that session_bank_i
//   already contains query i - reproduced, not fixed) dotted with the max-pooled encodings of its N documents (:127-131,
//   :155-156) -> scores [B,S,N].
// Reuses the recurrence engines of the rankers (tcgen05 rnn_tc.cu where the shape allows, else lstm.cu; embedding rows
// gathered inside the pre-gate GEMM).  Sessions are independent: [session_begin, +session_count) selects a shard.
#include "models.cuh"

namespace cair {

// pooled[n, :] = max_t bank[n, t, :]   (all L positions: pad rows are zeros)
__global__ void maxpool_time_kernel(const float* __restrict__ bank, int L, int H, float* __restrict__ pooled) {
  const int64_t n = blockIdx.x;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    const float* p = bank + (size_t)n * L * H + u;
    float m = p[0];
    for (int t = 1; t < L; ++t) m = fmaxf(m, p[(size_t)t * H]);
    pooled[n * H + u] = m;
  }
}

// comb[r, :] = [pq[r, :] ; s == 0 ? 0 : sess[r, :]]   (row r = (b, s))
__global__ void mnsrf_combine_kernel(const float* __restrict__ pq, const float* __restrict__ sess, int S, int Hq, int Hs,
                                     float* __restrict__ comb) {
  const int64_t r = blockIdx.x;
  const int s = (int)(r % S);
  for (int k = threadIdx.x; k < Hq + Hs; k += blockDim.x)
    comb[r * (Hq + Hs) + k] = k < Hq ? pq[r * Hq + k] : (s == 0 ? 0.f : sess[r * Hs + (k - Hq)]);
}

// scores[r, n] = proj[r, :] . pd[r * N + n, :]   (one warp per (r, n))
__global__ void mnsrf_dot_kernel(const float* __restrict__ proj, const float* __restrict__ pd, int64_t rows, int N, int H,
                                 float* __restrict__ scores) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= rows * N) return;
  const float* a = proj + (w / N) * H;
  const float* b = pd + w * H;
  float acc = 0.f;
  for (int k = lane; k < H; k += 32) acc = fmaf(a[k], b[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) scores[w] = acc;
}

__global__ void mnsrf_fill_len_kernel(int64_t* len, int n, int64_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = v;
}

struct MnsrfState {
  int device = 0, V = 0, E = 0, Hq = 0, Hd = 0, Hs = 0, rnn = 0;
  Owned own;
  float* table = nullptr;
  LstmPack enc_q{}, enc_d{}, sess{};
  RnnTcPack rt_q{}, rt_d{};
  float *pw = nullptr, *pb = nullptr;   // projection.linear [Hd, Hq+Hs] + bias
  int* d_err = nullptr;
};

static int32_t mnsrf_encode_pool(const MnsrfState& st, const LstmPack& lp, const RnnTcPack& rt, const int64_t* ids,
                                 const int64_t* len, int64_t n, int L, int H, float* pre, float* enc, float* pooled,
                                 cudaStream_t s) {
  if (g_rnn_impl >= RNN_IMPL_AUTO && rt.wimg)
    CAIR_TRY(rnn_tc_run(rt, gemm_gather(st.table, st.V, st.E, ids, 1, 1, 1, st.d_err), len, (int)n, L, enc, nullptr, nullptr, pre,
                        st.d_err, s, nullptr));
  else
    CAIR_TRY(lstm_run(lp, gemm_gather(st.table, st.V, st.E, ids, 1, 1, 1, st.d_err), len, (int)n, L, enc, nullptr, nullptr, pre,
                      st.d_err, s, nullptr));
  CAIR_LAUNCH(maxpool_time_kernel, (unsigned)n, 128, 0, s, enc, L, H, pooled);
  return CAIR_OK;
}

struct MnsrfWs {
  float *pre_q, *enc_q, *pq, *pre_d, *enc_d, *pd, *pre_s, *Qs, *Qc, *comb, *proj;
  int64_t* slen;
};
static void mnsrf_carve(const MnsrfState& st, Arena& ws, int S, int N, int Lq, int Ld, int sc, MnsrfWs* o) {
  const int64_t nrows = (int64_t)sc * S, ndocs = nrows * N;
  size_t pq = lstm_workspace_floats(st.enc_q, nrows, Lq), pd = lstm_workspace_floats(st.enc_d, ndocs, Ld);
  if (st.rt_q.wimg) pq = std::max(pq, rnn_tc_workspace_floats(st.rt_q, nrows, Lq));
  if (st.rt_d.wimg) pd = std::max(pd, rnn_tc_workspace_floats(st.rt_d, ndocs, Ld));
  o->pre_q = ws.take<float>(pq), o->enc_q = ws.take<float>((size_t)nrows * Lq * st.Hq), o->pq = ws.take<float>((size_t)nrows * st.Hq);
  o->pre_d = ws.take<float>(pd), o->enc_d = ws.take<float>((size_t)ndocs * Ld * st.Hd), o->pd = ws.take<float>((size_t)ndocs * st.Hd);
  o->pre_s = ws.take<float>(lstm_workspace_floats(st.sess, sc, S));
  o->Qs = ws.take<float>((size_t)nrows * st.Hs), o->Qc = ws.take<float>((size_t)nrows * st.Hs);
  o->comb = ws.take<float>((size_t)nrows * (st.Hq + st.Hs)), o->proj = ws.take<float>((size_t)nrows * st.Hd);
  o->slen = ws.take<int64_t>((size_t)sc);
}


// ---- attention-free suggestion decoder of MNSRF / M_MATCH_TENSOR (mnsrf.py:258-300, mmtensor.py:258-300) ----------
// decode(): RNNDecoder without attention (decoders/rnn_decoder.py:44-70; the state object is updated in place by every
// call, decoders/decoder.py:118-155) + generator + arg-max, greedy, from BOS.  Row conventions are the reference's: the
// initial states are concatenated query-index-major (torch.cat(hidden_states[:-1], dim=1): row i = s*B + b) while the
// predictions are viewed batch-major ([B, S-1, max_len]: row i = b*(S-1) + s) - reproduced, not fixed.
__global__ void sd_gather_state_kernel(const float* __restrict__ sess_h, const float* __restrict__ sess_c, int B, int S, int H,
                                       float* __restrict__ h, float* __restrict__ c) {
  const int i = blockIdx.x;   // decode row
  const int s1 = i / B, b1 = i - s1 * B;
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    h[(size_t)i * H + k] = sess_h[((size_t)b1 * S + s1) * H + k];
    c[(size_t)i * H + k] = sess_c[((size_t)b1 * S + s1) * H + k];
  }
}
__global__ void sd_cell_kernel(const float* __restrict__ gx, const float* __restrict__ gh, int R, int H, float* __restrict__ h,
                               float* __restrict__ c) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)R * H) return;
  const int r = (int)(idx / H), u = (int)(idx - (int64_t)r * H);
  const float* a = gx + (size_t)r * 4 * H;
  const float* b = gh + (size_t)r * 4 * H;
  const float ig = sigmoid_f(a[u] + b[u]), fg = sigmoid_f(a[H + u] + b[H + u]);
  const float gg = tanhf(a[2 * H + u] + b[2 * H + u]), og = sigmoid_f(a[3 * H + u] + b[3 * H + u]);
  const float cn = fg * c[idx] + ig * gg;
  c[idx] = cn;
  h[idx] = og * tanhf(cn);
}
// first maximum over the target vocabulary (torch.max), prediction store, next input through the target -> source id map
__global__ void __launch_bounds__(256) sd_argmax_kernel(const float* __restrict__ logits, int Vt, const int64_t* __restrict__ tgt2src,
                                                        int64_t* __restrict__ pred, int max_len, int t, int64_t* __restrict__ next) {
  __shared__ float bv[256];
  __shared__ int bi[256];
  const int i = blockIdx.x, tid = threadIdx.x;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int v = tid; v < Vt; v += 256) {
    const float x = logits[(size_t)i * Vt + v];
    if (x > best) best = x, arg = v;
  }
  bv[tid] = best, bi[tid] = arg;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o && (bv[tid + o] > bv[tid] || (bv[tid + o] == bv[tid] && bi[tid + o] < bi[tid]))) bv[tid] = bv[tid + o], bi[tid] = bi[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    const int a = bi[0] == 0x7fffffff ? 0 : bi[0];
    pred[(size_t)i * max_len + t] = a;
    next[i] = tgt2src[a];
  }
}
__global__ void sd_fill_kernel(int64_t* p, int n, int64_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct SessDecState {
  int device = 0, V = 0, E = 0, Hin = 0, Hs = 0, Vt = 0;
  bool has_session = false;
  Owned own;
  const float* table = nullptr;   // live pointer (the embedder's table)
  LstmPack sess{};
  float *w_ih = nullptr, *w_hh = nullptr, *bias = nullptr, *gen_w = nullptr, *gen_b = nullptr;
  int* d_err = nullptr;
};
struct SessDecWs {
  float *pre_s, *h, *c, *gx, *gh, *logits;
  int64_t *slen, *tok;
};
static void sessdec_carve(const SessDecState& st, Arena& ws, int B, int S, SessDecWs* o) {
  const size_t R = (size_t)B * (S > 1 ? S - 1 : 1), H = st.Hs;
  o->pre_s = ws.take<float>(st.has_session ? lstm_workspace_floats(st.sess, B, S) : 0);
  o->h = ws.take<float>(R * H), o->c = ws.take<float>(R * H), o->gx = ws.take<float>(R * 4 * H), o->gh = ws.take<float>(R * 4 * H);
  o->logits = ws.take<float>(R * st.Vt);
  o->slen = ws.take<int64_t>((size_t)B), o->tok = ws.take<int64_t>(R);
}

}  // namespace cair

using namespace cair;

struct cair_sessdec {
  SessDecState st;
};

struct cair_mnsrf {
  MnsrfState st;
};

namespace {
struct DevGuard2 {
  int prev = -1;
  explicit DevGuard2(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DevGuard2() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace

extern "C" {

int32_t cair_mnsrf_create(const cair_mnsrf_weights* w, int32_t device, cair_mnsrf** out) {
  if (!w || !out || !w->table || !w->projection.w || !w->projection.b || !w->session.w_ih)
    return fail(CAIR_ERR_BAD_ARG, "mnsrf_create: null argument");
  if (w->rnn_type != CAIR_RNN_LSTM && w->rnn_type != CAIR_RNN_GRU) return fail(CAIR_ERR_BAD_ARG, "mnsrf_create: bad rnn_type");
  DevGuard2 g(device);
  cair_mnsrf* h = new cair_mnsrf();
  MnsrfState& st = h->st;
  st.device = device, st.V = w->vocab, st.E = w->emsize, st.Hq = w->nhid_query, st.Hd = w->nhid_document, st.Hs = w->nhid_session;
  st.rnn = w->rnn_type;
  const int dirs = w->bidirectional ? 2 : 1;
  cudaStream_t s = 0;
  auto body = [&]() -> int32_t {
    if (st.Hq % dirs || st.Hd % dirs) return fail(CAIR_ERR_BAD_SHAPE, "mnsrf_create: hidden sizes must divide by the directions");
    CAIR_TRY(dev_copy(st.own, w->table, (size_t)st.V * st.E, &st.table, s));
    CAIR_TRY(lstm_pack(st.own, &w->query_fwd, dirs == 2 ? &w->query_rev : nullptr, st.E, st.Hq / dirs, &st.enc_q, s, w->rnn_type));
    CAIR_TRY(lstm_pack(st.own, &w->doc_fwd, dirs == 2 ? &w->doc_rev : nullptr, st.E, st.Hd / dirs, &st.enc_d, s, w->rnn_type));
    if (rnn_tc_supported(st.E, st.Hq / dirs))
      CAIR_TRY(rnn_tc_pack(st.own, &w->query_fwd, dirs == 2 ? &w->query_rev : nullptr, st.E, st.Hq / dirs, w->rnn_type, &st.rt_q, s));
    if (rnn_tc_supported(st.E, st.Hd / dirs))
      CAIR_TRY(rnn_tc_pack(st.own, &w->doc_fwd, dirs == 2 ? &w->doc_rev : nullptr, st.E, st.Hd / dirs, w->rnn_type, &st.rt_d, s));
    CAIR_TRY(lstm_pack(st.own, &w->session, nullptr, st.Hq, st.Hs, &st.sess, s, w->rnn_type));
    CAIR_TRY(dev_copy(st.own, w->projection.w, (size_t)st.Hd * (st.Hq + st.Hs), &st.pw, s));
    CAIR_TRY(dev_copy(st.own, w->projection.b, (size_t)st.Hd, &st.pb, s));
    CAIR_CUDA(st.own.alloc(&st.d_err, 1));
    CAIR_CUDA(cudaMemsetAsync(st.d_err, 0, sizeof(int), s));
    CAIR_CUDA(cudaStreamSynchronize(s));
    return CAIR_OK;
  };
  const int32_t rc = body();
  if (rc != CAIR_OK) {
    st.own.release();
    delete h;
    *out = nullptr;
    return rc;
  }
  *out = h;
  return CAIR_OK;
}

int32_t cair_mnsrf_destroy(cair_mnsrf* h) {
  if (!h) return CAIR_OK;
  DevGuard2 g(h->st.device);
  cudaDeviceSynchronize();
  h->st.own.release();
  delete h;
  return CAIR_OK;
}

int32_t cair_mnsrf_workspace_bytes(cair_mnsrf* h, int32_t B, int32_t S, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes) {
  if (!h || !bytes || B <= 0 || S <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_ARG, "mnsrf_workspace_bytes: bad argument");
  Arena a(nullptr, 0);
  MnsrfWs o;
  mnsrf_carve(h->st, a, S, N, Lq, Ld, B, &o);
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_mnsrf_forward(cair_mnsrf* h, const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen, int32_t B,
                           int32_t S, int32_t N, int32_t Lq, int32_t Ld, int32_t session_begin, int32_t session_count, float* scores,
                           float* memory_bank, float* session_bank, float* session_cell, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (!h || !q || !qlen || !d || !dlen || !scores || !workspace) return fail(CAIR_ERR_BAD_ARG, "mnsrf_forward: null argument");
  if (session_begin < 0 || session_count < 0 || session_begin + session_count > B) return fail(CAIR_ERR_BAD_ARG, "mnsrf_forward: session slice outside B");
  if ((uintptr_t)workspace % 256) return fail(CAIR_ERR_WORKSPACE, "mnsrf_forward: workspace must be 256-byte aligned");
  if (session_count == 0) return CAIR_OK;
  const MnsrfState& st = h->st;
  DevGuard2 g(st.device);
  cudaStream_t s = (cudaStream_t)stream;
  const int sc = session_count;
  const int64_t r0 = (int64_t)session_begin * S, nrows = (int64_t)sc * S, ndocs = nrows * N;
  Arena ws(workspace, workspace_bytes);
  MnsrfWs o;
  mnsrf_carve(st, ws, S, N, Lq, Ld, sc, &o);
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "mnsrf_forward: workspace too small");
  // encode (:61-80) and session encoding (:84-102)
  CAIR_TRY(mnsrf_encode_pool(st, st.enc_q, st.rt_q, q + r0 * Lq, qlen + r0, nrows, Lq, st.Hq, o.pre_q, o.enc_q, o.pq, s));
  CAIR_LAUNCH(mnsrf_fill_len_kernel, (sc + 255) / 256, 256, 0, s, o.slen, sc, (int64_t)S);
  CAIR_TRY(lstm_run(st.sess, gemm_dense(o.pq, st.Hq), o.slen, sc, S, o.Qs, nullptr, nullptr, o.pre_s, st.d_err, s, nullptr,
                    st.rnn == CAIR_RNN_LSTM ? o.Qc : nullptr));
  // rank_document (:116-162)
  CAIR_TRY(mnsrf_encode_pool(st, st.enc_d, st.rt_d, d + r0 * N * Ld, dlen + r0 * N, ndocs, Ld, st.Hd, o.pre_d, o.enc_d, o.pd, s));
  CAIR_LAUNCH(mnsrf_combine_kernel, (unsigned)nrows, 128, 0, s, o.pq, o.Qs, S, st.Hq, st.Hs, o.comb);
  CAIR_TRY(gemm_f32(gemm_dense(o.comb, st.Hq + st.Hs), st.pw, st.pb, o.proj, st.Hd, nrows, st.Hd, st.Hq + st.Hs, ACT_TANH, s));
  CAIR_LAUNCH(mnsrf_dot_kernel, (unsigned)((ndocs * 32 + 255) / 256), 256, 0, s, o.proj, o.pd, nrows, N, st.Hd, scores + r0 * N);
  if (memory_bank) CAIR_CUDA(cudaMemcpyAsync(memory_bank + r0 * st.Hq, o.pq, (size_t)nrows * st.Hq * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (session_bank) CAIR_CUDA(cudaMemcpyAsync(session_bank + r0 * st.Hs, o.Qs, (size_t)nrows * st.Hs * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (session_cell && st.rnn == CAIR_RNN_LSTM)
    CAIR_CUDA(cudaMemcpyAsync(session_cell + r0 * st.Hs, o.Qc, (size_t)nrows * st.Hs * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return CAIR_OK;
}

int32_t cair_sessdec_create(const cair_sessdec_weights* w, int32_t device, cair_sessdec** out) {
  if (!w || !out || !w->table || !w->dec_rnn.w_ih || !w->dec_rnn.w_hh || !w->dec_rnn.b_ih || !w->dec_rnn.b_hh || !w->generator.w ||
      !w->generator.b)
    return fail(CAIR_ERR_BAD_ARG, "sessdec_create: null argument");
  DevGuard2 g(device);
  cair_sessdec* h = new cair_sessdec();
  SessDecState& st = h->st;
  st.device = device, st.V = w->vocab, st.E = w->emsize, st.Hin = w->nhid_in, st.Hs = w->nhid_session, st.Vt = w->tgt_vocab;
  st.table = w->table;
  cudaStream_t s = 0;
  auto body = [&]() -> int32_t {
    const int H = st.Hs;
    if (w->session.w_ih) {
      CAIR_TRY(lstm_pack(st.own, &w->session, nullptr, st.Hin, H, &st.sess, s, CAIR_RNN_LSTM));
      st.has_session = true;
    }
    CAIR_TRY(dev_copy(st.own, w->dec_rnn.w_ih, (size_t)4 * H * st.E, &st.w_ih, s));
    CAIR_TRY(dev_copy(st.own, w->dec_rnn.w_hh, (size_t)4 * H * H, &st.w_hh, s));
    CAIR_CUDA(st.own.alloc(&st.bias, (size_t)4 * H));
    std::vector<float> bi(4 * H), bh(4 * H);
    CAIR_CUDA(cudaMemcpy(bi.data(), w->dec_rnn.b_ih, (size_t)4 * H * sizeof(float), cudaMemcpyDeviceToHost));
    CAIR_CUDA(cudaMemcpy(bh.data(), w->dec_rnn.b_hh, (size_t)4 * H * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4 * H; ++i) bi[i] += bh[i];
    CAIR_CUDA(cudaMemcpy(st.bias, bi.data(), (size_t)4 * H * sizeof(float), cudaMemcpyHostToDevice));
    CAIR_TRY(dev_copy(st.own, w->generator.w, (size_t)st.Vt * H, &st.gen_w, s));
    CAIR_TRY(dev_copy(st.own, w->generator.b, (size_t)st.Vt, &st.gen_b, s));
    CAIR_CUDA(st.own.alloc(&st.d_err, 1));
    CAIR_CUDA(cudaMemsetAsync(st.d_err, 0, sizeof(int), s));
    CAIR_CUDA(cudaStreamSynchronize(s));
    return CAIR_OK;
  };
  const int32_t rc = body();
  if (rc != CAIR_OK) {
    st.own.release();
    delete h;
    *out = nullptr;
    return rc;
  }
  *out = h;
  return CAIR_OK;
}

int32_t cair_sessdec_destroy(cair_sessdec* h) {
  if (!h) return CAIR_OK;
  DevGuard2 g(h->st.device);
  cudaDeviceSynchronize();
  h->st.own.release();
  delete h;
  return CAIR_OK;
}

int32_t cair_sessdec_workspace_bytes(cair_sessdec* h, int32_t B, int32_t S, size_t* bytes) {
  if (!h || !bytes || B <= 0 || S <= 0) return fail(CAIR_ERR_BAD_ARG, "sessdec_workspace_bytes: bad argument");
  Arena a(nullptr, 0);
  SessDecWs o;
  sessdec_carve(h->st, a, B, S, &o);
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_sessdec_states(cair_sessdec* h, const float* pooled, int32_t B, int32_t S, float* sess_h, float* sess_c, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (!h || !pooled || !sess_h || !sess_c || !workspace) return fail(CAIR_ERR_BAD_ARG, "sessdec_states: null argument");
  const SessDecState& st = h->st;
  if (!st.has_session) return fail(CAIR_ERR_BAD_ARG, "sessdec_states: created without session-encoder weights");
  DevGuard2 g(st.device);
  cudaStream_t s = (cudaStream_t)stream;
  Arena ws(workspace, workspace_bytes);
  SessDecWs o;
  sessdec_carve(st, ws, B, S, &o);
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "sessdec_states: workspace too small");
  CAIR_LAUNCH(mnsrf_fill_len_kernel, (B + 255) / 256, 256, 0, s, o.slen, B, (int64_t)S);
  return lstm_run(st.sess, gemm_dense(pooled, st.Hin), o.slen, B, S, sess_h, nullptr, nullptr, o.pre_s, st.d_err, s, nullptr, sess_c);
}

int32_t cair_sessdec_decode(cair_sessdec* h, const float* sess_h, const float* sess_c, int32_t B, int32_t S, int32_t max_len,
                            const int64_t* tgt2src, int64_t bos_id, int64_t* predictions, void* workspace, size_t workspace_bytes,
                            void* stream) {
  if (!h || !sess_h || !sess_c || !tgt2src || !predictions || !workspace) return fail(CAIR_ERR_BAD_ARG, "sessdec_decode: null argument");
  if (S < 2 || max_len < 1) return CAIR_OK;
  const SessDecState& st = h->st;
  DevGuard2 g(st.device);
  cudaStream_t s = (cudaStream_t)stream;
  Arena ws(workspace, workspace_bytes);
  SessDecWs o;
  sessdec_carve(st, ws, B, S, &o);
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "sessdec_decode: workspace too small");
  const int R = B * (S - 1), H = st.Hs;
  CAIR_LAUNCH(sd_gather_state_kernel, (unsigned)R, 128, 0, s, sess_h, sess_c, B, S, H, o.h, o.c);
  CAIR_LAUNCH(sd_fill_kernel, (R + 255) / 256, 256, 0, s, o.tok, R, bos_id);
  for (int t = 0; t < max_len; ++t) {
    CAIR_TRY(gemm_f32(gemm_gather(st.table, st.V, st.E, o.tok, 1, 1, 1, st.d_err), st.w_ih, st.bias, o.gx, 4 * H, R, 4 * H, st.E, ACT_NONE, s));
    CAIR_TRY(gemm_f32(gemm_dense(o.h, H), st.w_hh, nullptr, o.gh, 4 * H, R, 4 * H, H, ACT_NONE, s));
    CAIR_LAUNCH(sd_cell_kernel, (unsigned)(((int64_t)R * H + 255) / 256), 256, 0, s, o.gx, o.gh, R, H, o.h, o.c);
    CAIR_TRY(gemm_f32(gemm_dense(o.h, H), st.gen_w, st.gen_b, o.logits, st.Vt, R, st.Vt, H, ACT_NONE, s));
    CAIR_LAUNCH(sd_argmax_kernel, (unsigned)R, 256, 0, s, o.logits, st.Vt, tgt2src, predictions, max_len, t, o.tok);
  }
  return CAIR_OK;
}

/* out[n, :] = max_t (x[n, t, :] W^T + b): the max-pooled projected queries M_MATCH_TENSOR.encode feeds its session encoder
 * (mmtensor.py:86-92); scratch: n*L*C floats */
int32_t cair_linear_maxpool(const float* x, const float* w, const float* b, int32_t n, int32_t L, int32_t H, int32_t Cc, float* out,
                            float* scratch, void* stream) {
  if (!x || !w || !out || !scratch || n <= 0 || L <= 0) return fail(CAIR_ERR_BAD_ARG, "linear_maxpool: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  CAIR_TRY(gemm_f32(gemm_dense(x, H), w, b, scratch, Cc, (int64_t)n * L, Cc, H, ACT_NONE, s));
  CAIR_LAUNCH(maxpool_time_kernel, (unsigned)n, 128, 0, s, scratch, L, Cc, out);
  return CAIR_OK;
}

int32_t cair_mnsrf_poll_error(cair_mnsrf* h, void* stream) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  DevGuard2 g(h->st.device);
  int flags = 0;
  CAIR_CUDA(cudaMemcpyAsync(&flags, h->st.d_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CAIR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (flags) CAIR_CUDA(cudaMemsetAsync(h->st.d_err, 0, sizeof(int), (cudaStream_t)stream));
  if (flags & ERRF_BAD_TOKEN) return fail(CAIR_ERR_BAD_ARG, "token id outside [0, vocab)");
  if (flags & ERRF_BAD_LENGTH) return fail(CAIR_ERR_BAD_ARG, "sequence length outside [1, L]");
  return CAIR_OK;
}

}  // extern "C"
