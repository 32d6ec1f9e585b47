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

}  // namespace cair

using namespace cair;

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
