// C ABI of libcair.so (include/cair.h): handles, weight repacking, forward orchestration.
#include <cstdarg>
#include <cstring>

#include "models.cuh"

namespace cair {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
thread_local Profiler* g_prof = nullptr;

int32_t fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

}  // namespace cair

using namespace cair;

struct cair_handle {
  int model = 0;
  int device = 0;
  Owned own;
  int* d_err = nullptr;  // device error flags (ERRF_*)
  Profiler prof;
  // weights: struct copies hold the caller's device pointers only where create() documents a copy
  EsmState esm;
  MtState mt;
  DrmmState drmm;
  DuetState duet;
  DssmState dssm;
  ArciState arci;
  ArciiState arcii;
  CdssmState cdssm;
  CarsState cars;
  // host-path staging (cair_ranker_forward_host)
  void* stage_dev = nullptr;
  size_t stage_dev_bytes = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // host path as one CUDA graph (H2D + kernels + D2H), re-captured when the call signature changes
  struct HostKey {
    const void *q = nullptr, *ql = nullptr, *d = nullptr, *dl = nullptr, *scores = nullptr, *stage = nullptr, *ws = nullptr;
    int B = 0, N = 0, Lq = 0, Ld = 0, impl = -1;
    bool operator==(const HostKey& o) const {
      return q == o.q && ql == o.ql && d == o.d && dl == o.dl && scores == o.scores && stage == o.stage && ws == o.ws &&
             B == o.B && N == o.N && Lq == o.Lq && Ld == o.Ld && impl == o.impl;
    }
  } host_key;
  int host_key_hits = 0;
  cudaGraphExec_t host_graph = nullptr;
  cudaStream_t host_stream = nullptr;
  cudaEvent_t host_ev = nullptr;
  int* host_err = nullptr;  // pinned
  // pipelined host path (cair_ranker_submit_host / cair_ranker_wait_host): CAIR_PIPE_SLOTS staging slots + workspaces
  struct PipeSlot {
    void* stage = nullptr;
    size_t stage_bytes = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    cudaEvent_t ev_in = nullptr, ev_enc = nullptr, ev_done = nullptr, ev_int = nullptr;   // ev_int: interaction kernels done
    int* err = nullptr;  // pinned
    int* d_err = nullptr;  // this slot's own device error word: up to three batches are in flight on different streams
    bool busy = false;
    bool tail_pending = false;   // encoder enqueued, interaction not yet (cross-batch software pipeline)
    int B = 0, N = 0, Lq = 0, Ld = 0;
    float* scores_host = nullptr;
    cudaEvent_t tr[8] = {};   // optional trace (timing events): enc begin/end, part 1 begin/end, part 2 begin/end, finish
  } pipe[3];
  bool pipe_trace = false;
  int pipe_last = -1, pipe_last2 = -1;   // slots of the two most recently submitted batches
  float pipe_frac = 0.5f;                // share of a batch's pairs scored under the NEXT batch's document encoder
  int pipe_spc = 32;                     // document-encoder sequences per CTA in the pipeline (80 CTAs at cfg2: 68 SMs stay free)
  cudaStream_t hi_stream = nullptr, lo_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  // optional: all-gather of the scores over NVLink peer memory between the kernels and the D2H copy of the host entry points
  struct Gather {
    bool on = false;
    int rank = 0, world = 1;
    int64_t count = 0;
    uint32_t seq = 0;
    std::vector<uint64_t> recv[2], flags;
  } gather;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int32_t new_handle(int model, int device, cair_handle** out) {
  if (!out) return fail(CAIR_ERR_BAD_ARG, "null handle out-pointer");
  int ndev = 0;
  CAIR_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(CAIR_ERR_BAD_ARG, "device %d out of range (%d visible)", device, ndev);
  cair_handle* h = new cair_handle();
  h->model = model, h->device = device;
  *out = h;
  return CAIR_OK;
}

template <typename T>
int32_t copy_weights(Owned& own, const T* src, size_t n, T** dst, cudaStream_t s) {
  if (!src) return fail(CAIR_ERR_BAD_ARG, "null weight pointer");
  CAIR_CUDA(own.alloc(dst, n));
  CAIR_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, s));
  return CAIR_OK;
}

int32_t finish_create(cair_handle* h, int32_t rc, cair_handle** out) {
  if (rc == CAIR_OK) {
    cudaError_t e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) rc = fail(CAIR_ERR_CUDA, "create: %s", cudaGetErrorString(e));
  }
  if (rc != CAIR_OK) {
    h->own.release();
    delete h;
    *out = nullptr;
  }
  return rc;
}

int32_t init_err_flag(cair_handle* h) {
  CAIR_CUDA(h->own.alloc(&h->d_err, 1));
  CAIR_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), 0));
  return CAIR_OK;
}

}  // namespace

extern "C" {

int32_t cair_version(void) { return CAIR_VERSION; }
const char* cair_last_error(void) { return g_last_error.c_str(); }
int64_t cair_launch_count(void) { return g_launches.load(); }

int32_t cair_destroy(cair_handle* h) {
  if (!h) return CAIR_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();   // batches submitted through the pipelined entry points may still be in flight
  h->own.release();
  h->prof.release();
  if (h->mt.side) cudaStreamDestroy(h->mt.side);
  if (h->mt.ev_fork) cudaEventDestroy(h->mt.ev_fork);
  if (h->mt.ev_join) cudaEventDestroy(h->mt.ev_join);
  if (h->cars.side) cudaStreamDestroy(h->cars.side);
  if (h->cars.ev_fork) cudaEventDestroy(h->cars.ev_fork);
  if (h->cars.ev_join) cudaEventDestroy(h->cars.ev_join);
  if (h->duet.side) cudaStreamDestroy(h->duet.side);
  if (h->duet.ev_fork) cudaEventDestroy(h->duet.ev_fork);
  if (h->duet.ev_join) cudaEventDestroy(h->duet.ev_join);
  if (h->stage_dev) cudaFree(h->stage_dev);
  if (h->ws) cudaFree(h->ws);
  if (h->host_graph) cudaGraphExecDestroy(h->host_graph);
  if (h->host_stream) cudaStreamDestroy(h->host_stream);
  if (h->host_ev) cudaEventDestroy(h->host_ev);
  if (h->host_err) cudaFreeHost(h->host_err);
  for (auto& p : h->pipe) {
    if (p.stage) cudaFree(p.stage);
    if (p.ws) cudaFree(p.ws);
    if (p.ev_in) cudaEventDestroy(p.ev_in);
    if (p.ev_enc) cudaEventDestroy(p.ev_enc);
    if (p.ev_done) cudaEventDestroy(p.ev_done);
    if (p.ev_int) cudaEventDestroy(p.ev_int);
    for (auto& e : p.tr)
      if (e) cudaEventDestroy(e);
    if (p.err) cudaFreeHost(p.err);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->hi_stream) cudaStreamDestroy(h->hi_stream);
  if (h->lo_stream) cudaStreamDestroy(h->lo_stream);
  delete h;
  return CAIR_OK;
}

int32_t cair_poll_error(cair_handle* h, void* stream) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  DeviceGuard g(h->device);
  int flags = 0;
  CAIR_CUDA(cudaMemcpyAsync(&flags, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CAIR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (flags) {
    CAIR_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), (cudaStream_t)stream));
    if (flags & ERRF_BAD_TOKEN) return fail(CAIR_ERR_BAD_ARG, "token id outside [0, vocab)");
    return fail(CAIR_ERR_BAD_ARG, "sequence length outside [1, padded length]");
  }
  return CAIR_OK;
}

// ---- kernel-level entry points -----------------------------------------------------------------
int32_t cair_embed_gather(const float* table, int32_t V, int32_t E, const int64_t* ids, int64_t T, float* out,
                          void* stream) {
  if (!table || !ids || !out || V <= 0 || E <= 0 || T < 0) return fail(CAIR_ERR_BAD_ARG, "embed_gather: bad argument");
  return embed_gather(table, V, E, ids, T, out, nullptr, (cudaStream_t)stream);
}

int32_t cair_rnn_forward(int32_t rnn_type, const float* x, const int64_t* len, int32_t n, int32_t L, int32_t in, int32_t h,
                         const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out, float* h_n, float* c_n,
                         void* stream) {
  if (!x || !len || !out || !fwd || n < 0 || L <= 0) return fail(CAIR_ERR_BAD_ARG, "rnn_forward: bad argument");
  if (rnn_type != CAIR_RNN_LSTM && rnn_type != CAIR_RNN_GRU) return fail(CAIR_ERR_BAD_ARG, "rnn_forward: rnn_type must be LSTM or GRU");
  // unit-test entry point: packs the weights and allocates its scratch on every call
  cudaStream_t s = (cudaStream_t)stream;
  Owned own;
  LstmPack p;
  int32_t rc = lstm_pack(own, fwd, rev, in, h, &p, s, rnn_type);
  float* pre = nullptr;
  int* err = nullptr;
  // engine: cluster-split tcgen05 kernel when the shape fits (h <= 128), else / on request the round-1 tcgen05 kernel
  // (LSTM, in < 48, h <= 64) or the fp32 kernels
  const bool rt = rnn_tc_supported(in, h) &&
                  (g_rnn_impl == RNN_IMPL_CLUSTER || (g_rnn_impl == RNN_IMPL_AUTO && !rnn_prefers_r1(rnn_type, in, h)));
  const bool tc = !rt && g_rnn_impl != RNN_IMPL_FP32 && rnn_type == CAIR_RNN_LSTM && lstm_tc_supported(in, h);
  RnnTcPack rp;
  LstmTcPack tp;
  if (rc == CAIR_OK && rt) rc = rnn_tc_pack(own, fwd, rev, in, h, rnn_type, &rp, s);
  if (rc == CAIR_OK && tc) rc = lstm_tc_pack(own, fwd, rev, in, h, &tp, s);
  const size_t pre_floats = rt ? rnn_tc_workspace_floats(rp, n, L) : tc ? 0 : lstm_workspace_floats(p, n, L);
  if (rc == CAIR_OK && pre_floats && own.alloc(&pre, pre_floats) != cudaSuccess) rc = fail(CAIR_ERR_CUDA, "rnn_forward: out of memory");
  if (rc == CAIR_OK && own.alloc(&err, 1) != cudaSuccess) rc = fail(CAIR_ERR_CUDA, "rnn_forward: out of memory");
  if (rc == CAIR_OK) {
    cudaMemsetAsync(err, 0, sizeof(int), s);
    if (rt)
      rc = rnn_tc_run(rp, gemm_dense(x, in), len, n, L, out, h_n, c_n, pre, err, s, "lstm_recurrence");
    else if (tc)
      rc = lstm_tc_run(tp, p.bias, gemm_dense(x, in), len, n, L, out, h_n, c_n, err, s, "lstm_recurrence");
    else
      rc = lstm_run(p, gemm_dense(x, in), len, n, L, out, h_n, c_n, pre, err, s);
  }
  int flags = 0;
  if (rc == CAIR_OK) cudaMemcpyAsync(&flags, err, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  own.release();
  if (rc == CAIR_OK && e != cudaSuccess) rc = fail(CAIR_ERR_CUDA, "rnn_forward: %s", cudaGetErrorString(e));
  if (rc == CAIR_OK && flags) rc = fail(CAIR_ERR_BAD_ARG, "rnn_forward: length outside [1, L]");
  return rc;
}

int32_t cair_lstm_forward(const float* x, const int64_t* len, int32_t n, int32_t L, int32_t in, int32_t h,
                          const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out, float* h_n, float* c_n,
                          void* stream) {
  return cair_rnn_forward(CAIR_RNN_LSTM, x, len, n, L, in, h, fwd, rev, out, h_n, c_n, stream);
}

int32_t cair_set_rnn_impl(int32_t impl) {
  if (impl < RNN_IMPL_FP32 || impl > RNN_IMPL_CLUSTER)
    return fail(CAIR_ERR_BAD_ARG, "set_rnn_impl: 0 (fp32 CUDA cores), 1 (single-CTA tcgen05 kernel), 2 (auto) or 3 (cluster-split tcgen05 kernel)");
  g_rnn_impl = impl;
  return CAIR_OK;
}

// debugging / tuning aids (not in the public header)
extern "C" __attribute__((visibility("default"))) int32_t cair_drmm_debug_timing(long long* dev_counters) {
  cair::g_drmm_dbg = dev_counters;
  return CAIR_OK;
}
extern "C" __attribute__((visibility("default"))) int32_t cair_rnn_debug_timing(long long* dev_counters) {
  cair::g_rnn_dbg = dev_counters;
  return CAIR_OK;
}
extern "C" __attribute__((visibility("default"))) int32_t cair_rnn_set_seqs_per_cluster(int32_t min_spc, int32_t force_spc) {
  cair::g_rnn_spc_min = min_spc > 0 ? min_spc : 8;
  cair::g_rnn_spc_force = force_spc > 0 ? force_spc : 0;
  return CAIR_OK;
}

// ---- create ------------------------------------------------------------------------------------
int32_t cair_esm_create(const cair_esm_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table || w->vocab <= 0 || w->emsize <= 0) return fail(CAIR_ERR_BAD_ARG, "esm_create: bad weights");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_ESM, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  h->esm.V = w->vocab, h->esm.E = w->emsize;
  if (rc == CAIR_OK) rc = copy_weights(h->own, w->table, (size_t)w->vocab * w->emsize, &h->esm.table, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_mt_create(const cair_mt_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table) return fail(CAIR_ERR_BAD_ARG, "mt_create: bad weights");
  if (w->rnn_type != CAIR_RNN_LSTM && w->rnn_type != CAIR_RNN_GRU) return fail(CAIR_ERR_UNSUPPORTED, "mt_create: rnn_type must be LSTM or GRU");
  int dirs = w->bidirectional ? 2 : 1;
  if (w->nhid_query % dirs || w->nhid_doc % dirs) return fail(CAIR_ERR_BAD_SHAPE, "mt_create: hidden size not divisible by directions");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_MT, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = mt_create_state(h->own, *w, &h->mt, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_mt_add_encoder_layer(cair_handle* h, int32_t side, const cair_lstm_dir* fwd, const cair_lstm_dir* rev) {
  if (!h || h->model != CAIR_MODEL_MT) return fail(CAIR_ERR_BAD_ARG, "mt_add_encoder_layer: not a match-tensor handle");
  DeviceGuard g(h->device);
  CAIR_TRY(mt_add_encoder_layer(h->own, &h->mt, side, fwd, rev, 0));
  if (cudaStreamSynchronize(0) != cudaSuccess) return fail(CAIR_ERR_CUDA, "mt_add_encoder_layer: %s", cudaGetErrorString(cudaGetLastError()));
  return CAIR_OK;
}

int32_t cair_mt_set_debug(cair_handle* h, float* enc_q, float* enc_d) {
  if (!h || h->model != CAIR_MODEL_MT) return fail(CAIR_ERR_BAD_ARG, "mt_set_debug: not a match-tensor handle");
  h->mt.dbg_enc_q = enc_q, h->mt.dbg_enc_d = enc_d;
  return CAIR_OK;
}

int32_t cair_set_gemm_impl(int32_t impl) {
  if (impl < 0 || impl > 2) return fail(CAIR_ERR_BAD_ARG, "set_gemm_impl: 0 (fp32 CUDA cores), 1 (tcgen05, persistent) or 2 (tcgen05, one tile per CTA)");
  g_gemm_impl = impl;
  return CAIR_OK;
}

/* timing experiments only (not in the header): see g_gemm_dbg */
CAIR_API int32_t cair_debug_gemm(int32_t bits) {
  g_gemm_dbg = bits;
  return CAIR_OK;
}

int32_t cair_allgather_scores(const float* send, int64_t count, const uint64_t* peer_recv, const uint64_t* peer_flags,
                              int32_t rank, int32_t world, uint32_t seq, void* stream) {
  return allgather_scores_p2p(send, count, peer_recv, peer_flags, rank, world, seq, (cudaStream_t)stream);
}

int32_t cair_set_drmm_impl(int32_t impl) {
  if (impl != 0 && impl != 1) return fail(CAIR_ERR_BAD_ARG, "set_drmm_impl: 0 (fp32 CUDA cores) or 1 (tcgen05 cosines)");
  g_drmm_impl = impl;
  return CAIR_OK;
}

int32_t cair_mt_set_impl(cair_handle* h, int32_t impl) {
  if (!h || h->model != CAIR_MODEL_MT) return fail(CAIR_ERR_BAD_ARG, "mt_set_impl: not a match-tensor handle");
  if (impl < MT_IMPL_FP32 || impl > MT_IMPL_TC_SPLIT) return fail(CAIR_ERR_BAD_ARG, "mt_set_impl: impl must be 0 (fp32), 1 (tcgen05) or 2 (tcgen05, unfused projection)");
  h->mt.impl = impl;
  return CAIR_OK;
}

// debugging aid (not in the public header): role-timing counters of the tcgen05 interaction kernel
extern "C" __attribute__((visibility("default"))) int32_t cair_mt_debug_timing(long long* dev_counters) {
  cair::g_mt_dbg = dev_counters;
  return CAIR_OK;
}

// Tuning knob (tools/pipe_tune.py): smallest number of sequences per CTA the tcgen05 LSTM may choose (8, 16, 24 or 32).
extern "C" __attribute__((visibility("default"))) int32_t cair_lstm_set_min_seqs_per_cta(int32_t spc) {
  if (spc != 8 && spc != 16 && spc != 24 && spc != 32) return fail(CAIR_ERR_BAD_ARG, "seqs per CTA must be 8, 16, 24 or 32");
  cair::g_lstm_spc_min = spc;
  return CAIR_OK;
}

extern "C" __attribute__((visibility("default"))) int32_t cair_lstm_debug_timing(long long* dev_counters) {
  cair::g_lstm_dbg = dev_counters;
  return CAIR_OK;
}

int32_t cair_drmm_create(const cair_drmm_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table) return fail(CAIR_ERR_BAD_ARG, "drmm_create: bad weights");
  if (w->nbins != 5) return fail(CAIR_ERR_UNSUPPORTED, "drmm_create: nbins must be 5 (neuroir/hyparam.py:78-81)");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_DRMM, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  DrmmState& st = h->drmm;
  st.w = *w;
  float* t = nullptr;
  if (rc == CAIR_OK) rc = copy_weights(h->own, w->table, (size_t)w->vocab * w->emsize, &t, 0);
  st.w.table = t;
  auto cp = [&](const float* src, size_t n, const float** dst) {
    float* p = nullptr;
    if (rc == CAIR_OK) rc = copy_weights(h->own, src, n, &p, 0);
    *dst = p;
  };
  cp(w->gating.w, w->emsize, &st.w.gating.w);
  cp(w->gating.b, 1, &st.w.gating.b);
  cp(w->ffnn0.w, 5, &st.w.ffnn0.w);
  cp(w->ffnn0.b, 1, &st.w.ffnn0.b);
  cp(w->ffnn1.w, 1, &st.w.ffnn1.w);
  cp(w->ffnn1.b, 1, &st.w.ffnn1.b);
  cp(w->output.w, 1, &st.w.output.w);
  cp(w->output.b, 1, &st.w.output.b);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_drmm_set_debug(cair_handle* h, int32_t* hist) {
  if (!h || h->model != CAIR_MODEL_DRMM) return fail(CAIR_ERR_BAD_ARG, "drmm_set_debug: not a DRMM handle");
  h->drmm.dbg_hist = hist;
  return CAIR_OK;
}

int32_t cair_duet_create(const cair_duet_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table) return fail(CAIR_ERR_BAD_ARG, "duet_create: bad weights");
  if (w->local_filter_size != 1 || w->dist_filter_size != 3)
    return fail(CAIR_ERR_UNSUPPORTED, "duet_create: only local_filter_size=1, dist_filter_size=3 (the shapes duet.py:144 admits)");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_DUET, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = duet_create_state(h->own, *w, &h->duet, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_dssm_create(const cair_dssm_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table || w->nhid <= 0 || w->nout <= 0) return fail(CAIR_ERR_BAD_ARG, "dssm_create: bad weights");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_DSSM, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = dssm_create_state(h->own, *w, &h->dssm, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_cdssm_create(const cair_cdssm_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table || w->nhid <= 0 || w->nout <= 0) return fail(CAIR_ERR_BAD_ARG, "cdssm_create: bad weights");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_CDSSM, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = cdssm_create_state(h->own, *w, &h->cdssm, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_arci_create(const cair_arci_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table) return fail(CAIR_ERR_BAD_ARG, "arci_create: bad weights");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_ARCI, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = arci_create_state(h->own, *w, &h->arci, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_arcii_create(const cair_arcii_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table) return fail(CAIR_ERR_BAD_ARG, "arcii_create: bad weights");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_ARCII, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = arcii_create_state(h->own, *w, &h->arcii, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

int32_t cair_cars_create(const cair_cars_weights* w, int32_t device, cair_handle** out) {
  if (!w || !w->table) return fail(CAIR_ERR_BAD_ARG, "cars_create: bad weights");
  cair_handle* h = nullptr;
  CAIR_TRY(new_handle(CAIR_MODEL_CARS, device, &h));
  DeviceGuard g(device);
  int32_t rc = init_err_flag(h);
  if (rc == CAIR_OK) rc = cars_create_state(h->own, *w, &h->cars, 0);
  rc = finish_create(h, rc, out);
  if (rc == CAIR_OK) *out = h;
  return rc;
}

// ---- stand-alone rankers: workspace + forward ---------------------------------------------------
static int32_t check_ranker_args(cair_handle* h, int32_t B, int32_t N, int32_t Lq, int32_t Ld) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  if (h->model == CAIR_MODEL_CARS) return fail(CAIR_ERR_BAD_ARG, "CARS handles use cair_cars_forward");
  if (B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_SHAPE, "B, N, Lq, Ld must be positive");
  return CAIR_OK;
}

static int32_t ranker_run(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                          const int64_t* dlen, int B, int N, int Lq, int Ld, int64_t pb, int64_t pc, float* scores,
                          Arena& ws, cudaStream_t s, bool dry) {
  switch (h->model) {
    case CAIR_MODEL_ESM:
      if (dry) return CAIR_OK;
      return esm_forward(h->esm.table, h->esm.V, h->esm.E, q, d, N, Lq, Ld, pb, pc, scores, h->d_err, s);
    case CAIR_MODEL_DRMM:
      return drmm_forward(h->drmm.w, q, d, N, Lq, Ld, pb, pc, scores, h->drmm.dbg_hist, ws, h->d_err, s, dry);
    case CAIR_MODEL_MT:
      return mt_forward(h->mt, q, qlen, d, dlen, B, N, Lq, Ld, pb, pc, scores, ws, h->d_err, s, dry);
    case CAIR_MODEL_DSSM:
      return dssm_forward(h->dssm, q, d, N, Lq, Ld, pb, pc, scores, ws, h->d_err, s, dry);
    case CAIR_MODEL_CDSSM:
      return cdssm_forward(h->cdssm, q, d, N, Lq, Ld, pb, pc, scores, ws, h->d_err, s, dry);
    case CAIR_MODEL_ARCI:
      return arci_forward(h->arci, q, d, N, Lq, Ld, pb, pc, scores, ws, h->d_err, s, dry);
    case CAIR_MODEL_ARCII:
      return arcii_forward(h->arcii, q, d, N, Lq, Ld, pb, pc, scores, ws, h->d_err, s, dry);
    case CAIR_MODEL_DUET:
      return duet_forward(h->duet, q, d, B, N, Lq, Ld, pb, pc, scores, ws, h->d_err, s, dry);
  }
  return fail(CAIR_ERR_BAD_ARG, "unknown model");
}

int32_t cair_ranker_workspace_bytes(cair_handle* h, int32_t B, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes) {
  CAIR_TRY(check_ranker_args(h, B, N, Lq, Ld));
  if (!bytes) return fail(CAIR_ERR_BAD_ARG, "null bytes");
  Arena a(nullptr, 0);
  CAIR_TRY(ranker_run(h, nullptr, nullptr, nullptr, nullptr, B, N, Lq, Ld, 0, (int64_t)B * N, nullptr, a, 0, true));
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_ranker_forward(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                            const int64_t* dlen, int32_t B, int32_t N, int32_t Lq, int32_t Ld, int64_t pair_begin,
                            int64_t pair_count, float* scores, void* workspace, size_t workspace_bytes,
                            void* stream) {
  CAIR_TRY(check_ranker_args(h, B, N, Lq, Ld));
  if (!q || !qlen || !d || !dlen || !scores) return fail(CAIR_ERR_BAD_ARG, "ranker_forward: null tensor");
  if (pair_begin < 0 || pair_count < 0 || pair_begin + pair_count > (int64_t)B * N)
    return fail(CAIR_ERR_BAD_ARG, "ranker_forward: pair slice [%lld, +%lld) outside B*N", (long long)pair_begin, (long long)pair_count);
  if (((uintptr_t)workspace & 255) != 0) return fail(CAIR_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  DeviceGuard g(h->device);
  Arena probe(nullptr, 0);
  CAIR_TRY(ranker_run(h, nullptr, nullptr, nullptr, nullptr, B, N, Lq, Ld, pair_begin, pair_count, nullptr, probe, 0, true));
  if (probe.off > workspace_bytes || (probe.off > 0 && !workspace))
    return fail(CAIR_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", probe.off, workspace_bytes);
  Arena ws(workspace, workspace_bytes);
  h->prof.reset();
  g_prof = &h->prof;
  prof_mark("begin", (cudaStream_t)stream);
  int32_t rc = ranker_run(h, q, qlen, d, dlen, B, N, Lq, Ld, pair_begin, pair_count, scores, ws, (cudaStream_t)stream, false);
  prof_mark("end", (cudaStream_t)stream);
  g_prof = nullptr;
  return rc;
}

int32_t cair_profile_enable(cair_handle* h, int32_t on) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  h->prof.on = on != 0;
  h->prof.reset();
  return CAIR_OK;
}

int32_t cair_profile_read(cair_handle* h, char* names, size_t names_bytes, float* ms, int32_t capacity, int32_t* count) {
  if (!h || !ms || !count) return fail(CAIR_ERR_BAD_ARG, "profile_read: bad argument");
  DeviceGuard g(h->device);
  Profiler& p = h->prof;
  int n = p.cursor > 0 ? p.cursor - 1 : 0;
  if (n > capacity) n = capacity;
  std::string joined;
  for (int i = 0; i < n; ++i) {
    CAIR_CUDA(cudaEventSynchronize(p.ev[i + 1]));
    CAIR_CUDA(cudaEventElapsedTime(&ms[i], p.ev[i], p.ev[i + 1]));
    joined += p.names[i];
    joined += (i + 1 < n) ? "," : "";
  }
  if (names && names_bytes > 0) {
    strncpy(names, joined.c_str(), names_bytes - 1);
    names[names_bytes - 1] = 0;
  }
  *count = n;
  return CAIR_OK;
}

int32_t cair_ranker_forward_host(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                                 const int64_t* dlen, int32_t B, int32_t N, int32_t Lq, int32_t Ld, float* scores,
                                 void* stream) {
  CAIR_TRY(check_ranker_args(h, B, N, Lq, Ld));
  if (!q || !qlen || !d || !dlen || !scores) return fail(CAIR_ERR_BAD_ARG, "ranker_forward_host: null tensor");
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nq = (size_t)B * Lq, nd = (size_t)B * N * Ld, nb = (size_t)B, nbn = (size_t)B * N;
  const size_t in_bytes = (nq + nb + nd + nbn) * sizeof(int64_t);
  const size_t need = align_up(in_bytes) + align_up(nbn * sizeof(float));
  if (need > h->stage_dev_bytes) {
    if (h->stage_dev) CAIR_CUDA(cudaFree(h->stage_dev));
    h->stage_dev = nullptr, h->stage_dev_bytes = 0;
    CAIR_CUDA(cudaMalloc(&h->stage_dev, need));
    h->stage_dev_bytes = need;
  }
  size_t wsb = 0;
  CAIR_TRY(cair_ranker_workspace_bytes(h, B, N, Lq, Ld, &wsb));
  if (wsb > h->ws_bytes) {
    if (h->ws) CAIR_CUDA(cudaFree(h->ws));
    h->ws = nullptr, h->ws_bytes = 0;
    CAIR_CUDA(cudaMalloc(&h->ws, wsb));
    h->ws_bytes = wsb;
  }
  int64_t* dq = (int64_t*)h->stage_dev;
  int64_t* dql = dq + nq;
  int64_t* dd = dql + nb;
  int64_t* ddl = dd + nd;
  float* ds = (float*)((char*)h->stage_dev + align_up(in_bytes));
  if (!h->host_stream) {
    CAIR_CUDA(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    CAIR_CUDA(cudaEventCreateWithFlags(&h->host_ev, cudaEventDisableTiming));
    CAIR_CUDA(cudaHostAlloc((void**)&h->host_err, sizeof(int), cudaHostAllocDefault));
  }
  // The whole call runs on the handle's own stream, ordered after the work already queued on `stream`
  // (the legacy default stream cannot be captured); the call returns after a host synchronisation.
  cudaStream_t hs = h->host_stream;
  CAIR_CUDA(cudaEventRecord(h->host_ev, s));
  CAIR_CUDA(cudaStreamWaitEvent(hs, h->host_ev, 0));
  auto enqueue = [&]() -> int32_t {
    CAIR_CUDA(cudaMemcpyAsync(dq, q, nq * sizeof(int64_t), cudaMemcpyHostToDevice, hs));
    CAIR_CUDA(cudaMemcpyAsync(dql, qlen, nb * sizeof(int64_t), cudaMemcpyHostToDevice, hs));
    CAIR_CUDA(cudaMemcpyAsync(dd, d, nd * sizeof(int64_t), cudaMemcpyHostToDevice, hs));
    CAIR_CUDA(cudaMemcpyAsync(ddl, dlen, nbn * sizeof(int64_t), cudaMemcpyHostToDevice, hs));
    CAIR_TRY(cair_ranker_forward(h, dq, dql, dd, ddl, B, N, Lq, Ld, 0, (int64_t)nbn, ds, h->ws, h->ws_bytes, hs));
    CAIR_CUDA(cudaMemcpyAsync(scores, ds, nbn * sizeof(float), cudaMemcpyDeviceToHost, hs));
    CAIR_CUDA(cudaMemcpyAsync(h->host_err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, hs));
    return CAIR_OK;
  };
  cair_handle::HostKey key;
  key.q = q, key.ql = qlen, key.d = d, key.dl = dlen, key.scores = scores, key.stage = h->stage_dev, key.ws = h->ws;
  key.B = B, key.N = N, key.Lq = Lq, key.Ld = Ld, key.impl = h->mt.impl * 2 + g_gemm_impl;
  const bool same = key == h->host_key;
  if (same && h->host_graph && !h->prof.on) {
    CAIR_CUDA(cudaGraphLaunch(h->host_graph, hs));
  } else if (same && !h->prof.on && h->host_key_hits >= 1) {
    // second call with this signature: capture H2D + kernels + D2H once, replay from now on
    if (h->host_graph) cudaGraphExecDestroy(h->host_graph);
    h->host_graph = nullptr;
    cudaGraph_t graph = nullptr;
    CAIR_CUDA(cudaStreamBeginCapture(hs, cudaStreamCaptureModeThreadLocal));
    int32_t rc = enqueue();
    cudaError_t ce = cudaStreamEndCapture(hs, &graph);
    if (rc != CAIR_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (ce != cudaSuccess) return fail(CAIR_ERR_CUDA, "forward_host: graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&h->host_graph, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return fail(CAIR_ERR_CUDA, "forward_host: graph instantiate failed: %s", cudaGetErrorString(ce));
    CAIR_CUDA(cudaGraphLaunch(h->host_graph, hs));
  } else {
    if (!same) {
      if (h->host_graph) cudaGraphExecDestroy(h->host_graph);
      h->host_graph = nullptr;
      h->host_key = key;
      h->host_key_hits = 0;
    }
    h->host_key_hits++;
    CAIR_TRY(enqueue());
  }
  CAIR_CUDA(cudaStreamSynchronize(hs));
  if (*h->host_err) {
    const int flags = *h->host_err;
    CAIR_CUDA(cudaMemsetAsync(h->d_err, 0, sizeof(int), hs));
    CAIR_CUDA(cudaStreamSynchronize(hs));
    if (flags & ERRF_BAD_TOKEN) return fail(CAIR_ERR_BAD_ARG, "token id outside [0, vocab)");
    return fail(CAIR_ERR_BAD_ARG, "sequence length outside [1, padded length]");
  }
  return CAIR_OK;
}

namespace {

struct PipeView {
  int64_t *dq, *dql, *dd, *ddl;
  float* ds;
};
PipeView pipe_view(const cair_handle::PipeSlot& p) {
  const size_t nq = (size_t)p.B * p.Lq, nd = (size_t)p.B * p.N * p.Ld, nb = (size_t)p.B, nbn = (size_t)p.B * p.N;
  PipeView v;
  v.dq = (int64_t*)p.stage;
  v.dql = v.dq + nq;
  v.dd = v.dql + nb;
  v.ddl = v.dd + nd;
  v.ds = (float*)((char*)p.stage + align_up((nq + nb + nd + nbn) * sizeof(int64_t)));
  return v;
}

int32_t pipe_mark(cair_handle* h, int sl, int i, cudaStream_t st) {
  if (!h->pipe_trace) return CAIR_OK;
  cair_handle::PipeSlot& p = h->pipe[sl];
  if (!p.tr[i]) CAIR_CUDA(cudaEventCreate(&p.tr[i]));
  CAIR_CUDA(cudaEventRecord(p.tr[i], st));
  return CAIR_OK;
}

// Interaction phase of the batch in slot `sl` over the pair sub-range [ib, ib+ic) on stream `st`.
int32_t pipe_interact(cair_handle* h, int sl, int64_t ib, int64_t ic, int max_ctas, bool join, cudaStream_t st) {
  cair_handle::PipeSlot& p = h->pipe[sl];
  const PipeView v = pipe_view(p);
  Arena ws(p.ws, p.ws_bytes);
  MtPhase ph;
  ph.phase = MT_INTERACT, ph.ib = ib, ph.ic = ic, ph.max_ctas = max_ctas, ph.join = join;
  return mt_forward(h->mt, v.dq, v.dql, v.dd, v.ddl, p.B, p.N, p.Lq, p.Ld, 0, (int64_t)p.B * p.N, v.ds, ws, p.d_err, st,
                    false, ph);
}

// Scores + error flag of slot `sl` back to the host on stream `st`, completion event.
int32_t pipe_finish(cair_handle* h, int sl, cudaStream_t st) {
  cair_handle::PipeSlot& p = h->pipe[sl];
  const PipeView v = pipe_view(p);
  CAIR_CUDA(cudaEventRecord(p.ev_int, st));   // the machine is free again: the next encoder need not wait for the copies
  if (h->gather.on) {
    // doc-parallel serving: every rank's slice into every rank's receive buffer (one kernel over NVLink peer memory), then
    // ALL world * B * N scores to this rank's host buffer.  Two receive buffers alternate; a peer overwrites buffer k & 1
    // only after this rank's gather k-1 ran, which is stream-ordered behind the D2H copy of result k-2.
    cair_handle::Gather& g = h->gather;
    if ((int64_t)p.B * p.N != g.count) return fail(CAIR_ERR_BAD_SHAPE, "ranker host path: batch has %lld pairs, the gather was set up for %lld", (long long)p.B * p.N, (long long)g.count);
    const uint32_t seq = ++g.seq;
    const int par = (int)(seq & 1);
    CAIR_TRY(allgather_scores_p2p(v.ds, g.count, g.recv[par].data(), g.flags.data(), g.rank, g.world, seq, st));
    CAIR_CUDA(cudaMemcpyAsync(p.scores_host, reinterpret_cast<const float*>(g.recv[par][g.rank]), (size_t)g.world * g.count * sizeof(float),
                              cudaMemcpyDeviceToHost, st));
  } else {
    CAIR_CUDA(cudaMemcpyAsync(p.scores_host, v.ds, (size_t)p.B * p.N * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  int* flag = p.tail_pending ? p.d_err : h->d_err;   // pipelined batches carry their own flag; the plain form uses the handle's
  CAIR_CUDA(cudaMemcpyAsync(p.err, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CAIR_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  CAIR_CUDA(cudaEventRecord(p.ev_done, st));
  p.tail_pending = false;
  return CAIR_OK;
}

}  // namespace

int32_t cair_ranker_submit_host(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                                const int64_t* dlen, int32_t B, int32_t N, int32_t Lq, int32_t Ld, float* scores,
                                int32_t slot, void* stream) {
  CAIR_TRY(check_ranker_args(h, B, N, Lq, Ld));
  if (!q || !qlen || !d || !dlen || !scores) return fail(CAIR_ERR_BAD_ARG, "ranker_submit_host: null tensor");
  if (slot < 0 || slot > 2) return fail(CAIR_ERR_BAD_ARG, "ranker_submit_host: slot must be 0, 1 or 2");
  DeviceGuard g(h->device);
  cair_handle::PipeSlot& p = h->pipe[slot];
  if (p.busy) return fail(CAIR_ERR_BAD_ARG, "ranker_submit_host: slot %d submitted again before its wait", slot);
  const size_t nq = (size_t)B * Lq, nd = (size_t)B * N * Ld, nb = (size_t)B, nbn = (size_t)B * N;
  const size_t in_bytes = (nq + nb + nd + nbn) * sizeof(int64_t);
  const size_t need = align_up(in_bytes) + align_up(nbn * sizeof(float));
  if (!h->host_stream) {
    CAIR_CUDA(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    CAIR_CUDA(cudaEventCreateWithFlags(&h->host_ev, cudaEventDisableTiming));
    CAIR_CUDA(cudaHostAlloc((void**)&h->host_err, sizeof(int), cudaHostAllocDefault));
  }
  if (!h->copy_stream) {
    int lo = 0, hi = 0;
    CAIR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // numerically lowest = highest priority
    CAIR_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CAIR_CUDA(cudaStreamCreateWithPriority(&h->hi_stream, cudaStreamNonBlocking, hi));
    CAIR_CUDA(cudaStreamCreateWithPriority(&h->lo_stream, cudaStreamNonBlocking, lo));
  }
  if (!p.ev_in) {
    CAIR_CUDA(cudaEventCreateWithFlags(&p.ev_in, cudaEventDisableTiming));
    CAIR_CUDA(cudaEventCreateWithFlags(&p.ev_enc, cudaEventDisableTiming));
    CAIR_CUDA(cudaEventCreateWithFlags(&p.ev_done, cudaEventDisableTiming));
    CAIR_CUDA(cudaEventCreateWithFlags(&p.ev_int, cudaEventDisableTiming));
    CAIR_CUDA(cudaHostAlloc((void**)&p.err, sizeof(int), cudaHostAllocDefault));
    *p.err = 0;
    CAIR_CUDA(h->own.alloc(&p.d_err, 1));
    CAIR_CUDA(cudaMemset(p.d_err, 0, sizeof(int)));
  }
  size_t wsb = 0;
  CAIR_TRY(cair_ranker_workspace_bytes(h, B, N, Lq, Ld, &wsb));
  if (need > p.stage_bytes || wsb > p.ws_bytes) {   // (re)allocation synchronises; steady state never gets here
    CAIR_CUDA(cudaDeviceSynchronize());
    if (need > p.stage_bytes) {
      if (p.stage) CAIR_CUDA(cudaFree(p.stage));
      p.stage = nullptr, p.stage_bytes = 0;
      CAIR_CUDA(cudaMalloc(&p.stage, need));
      p.stage_bytes = need;
    }
    if (wsb > p.ws_bytes) {
      if (p.ws) CAIR_CUDA(cudaFree(p.ws));
      p.ws = nullptr, p.ws_bytes = 0;
      CAIR_CUDA(cudaMalloc(&p.ws, wsb));
      p.ws_bytes = wsb;
    }
  }
  p.B = B, p.N = N, p.Lq = Lq, p.Ld = Ld, p.scores_host = scores;
  const PipeView v = pipe_view(p);
  cudaStream_t us = (cudaStream_t)stream;
  // everything below is ordered after the work already queued on the caller's stream
  CAIR_CUDA(cudaEventRecord(h->host_ev, us));
  cudaStream_t cs = h->copy_stream;
  CAIR_CUDA(cudaStreamWaitEvent(cs, h->host_ev, 0));
  CAIR_CUDA(cudaMemcpyAsync(v.dq, q, nq * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
  CAIR_CUDA(cudaMemcpyAsync(v.dql, qlen, nb * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
  CAIR_CUDA(cudaMemcpyAsync(v.dd, d, nd * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
  CAIR_CUDA(cudaMemcpyAsync(v.ddl, dlen, nbn * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
  CAIR_CUDA(cudaEventRecord(p.ev_in, cs));

  const bool pipelined = h->model == CAIR_MODEL_MT && mt_can_pipeline(h->mt, Lq, Ld) && !h->prof.on;
  if (!pipelined) {
    // plain form: kernels + scores back on one compute stream, behind the previous batch
    cudaStream_t hs = us ? us : h->host_stream;
    CAIR_CUDA(cudaStreamWaitEvent(hs, p.ev_in, 0));
    CAIR_TRY(cair_ranker_forward(h, v.dq, v.dql, v.dd, v.ddl, B, N, Lq, Ld, 0, (int64_t)nbn, v.ds, p.ws, p.ws_bytes, hs));
    CAIR_TRY(pipe_finish(h, slot, hs));
    p.busy = true;
    return CAIR_OK;
  }
  // ---- cross-batch software pipeline (Match-Tensor, tcgen05 path) ----
  // The document encoder is a 200-step dependency chain on ~108 of the 148 SMs; the interaction kernel of the PREVIOUS
  // batch has no such chain.  Per submit:  lo: interaction part 1 of the previous batch on the SMs the encoder leaves
  // free  ||  hi: this batch's encoder + projection;  then lo: interaction part 2 of the previous batch on all SMs,
  // its scores back to the host.  This batch's own interaction is enqueued by the next submit (or by its wait).
  cudaStream_t H = h->hi_stream, L = h->lo_stream;
  const int prev = (h->pipe_last >= 0 && h->pipe[h->pipe_last].tail_pending) ? h->pipe_last : -1;
  int64_t c1 = 0;
  int free_sms = 0;
  if (prev >= 0) {
    cair_handle::PipeSlot& pp = h->pipe[prev];
    const int64_t pcp = (int64_t)pp.B * pp.N;
    free_sms = kSMs - (mt_doc_uses_cluster_kernel(h->mt) ? rnn_tc_plan(h->mt.rt_d, (int)nbn, h->pipe_spc).ctas
                                                         : lstm_tc_ctas((int)nbn, h->mt.tc_d.dirs, h->pipe_spc));
    if (free_sms >= 8) c1 = (int64_t)((double)pcp * h->pipe_frac);
    CAIR_CUDA(cudaStreamWaitEvent(L, pp.ev_enc, 0));
    CAIR_TRY(pipe_mark(h, prev, 2, L));
    if (c1 > 0) CAIR_TRY(pipe_interact(h, prev, 0, c1, free_sms, true, L));   // joins the previous batch's query side
    CAIR_TRY(pipe_mark(h, prev, 3, L));
  }
  CAIR_CUDA(cudaStreamWaitEvent(H, p.ev_in, 0));
  // the encoder must not start while the batch before the previous one still holds the machine with its part 2
  if (h->pipe_last2 >= 0 && h->pipe_last2 != slot && h->pipe[h->pipe_last2].busy)
    CAIR_CUDA(cudaStreamWaitEvent(H, h->pipe[h->pipe_last2].ev_int, 0));
  CAIR_TRY(pipe_mark(h, slot, 0, H));
  {
    Arena ws(p.ws, p.ws_bytes);
    MtPhase ph;
    ph.phase = MT_ENCODE;
    ph.doc_min_spc = h->pipe_spc;
    CAIR_TRY(mt_forward(h->mt, v.dq, v.dql, v.dd, v.ddl, B, N, Lq, Ld, 0, (int64_t)nbn, v.ds, ws, p.d_err, H, false, ph));
  }
  CAIR_TRY(pipe_mark(h, slot, 1, H));
  CAIR_CUDA(cudaEventRecord(p.ev_enc, H));
  if (prev >= 0) {
    cair_handle::PipeSlot& pp = h->pipe[prev];
    const int64_t pcp = (int64_t)pp.B * pp.N;
    CAIR_CUDA(cudaStreamWaitEvent(L, p.ev_enc, 0));   // part 2 gets the whole machine: after this batch's encoder
    CAIR_TRY(pipe_mark(h, prev, 4, L));
    CAIR_TRY(pipe_interact(h, prev, c1, pcp - c1, 0, c1 == 0, L));
    CAIR_TRY(pipe_mark(h, prev, 5, L));
    CAIR_TRY(pipe_finish(h, prev, L));
    CAIR_TRY(pipe_mark(h, prev, 6, L));
  }
  p.tail_pending = true;
  p.busy = true;
  h->pipe_last2 = h->pipe_last;
  h->pipe_last = slot;
  return CAIR_OK;
}

int32_t cair_ranker_wait_host(cair_handle* h, int32_t slot) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  if (slot < 0 || slot > 2) return fail(CAIR_ERR_BAD_ARG, "ranker_wait_host: slot must be 0, 1 or 2");
  DeviceGuard g(h->device);
  cair_handle::PipeSlot& p = h->pipe[slot];
  if (!p.busy) return fail(CAIR_ERR_BAD_ARG, "ranker_wait_host: nothing submitted on slot %d", slot);
  if (p.tail_pending) {
    // no later submit took care of this batch's interaction: run it now on the whole machine
    cudaStream_t L = h->lo_stream;
    CAIR_CUDA(cudaStreamWaitEvent(L, p.ev_enc, 0));
    CAIR_TRY(pipe_interact(h, slot, 0, (int64_t)p.B * p.N, 0, true, L));
    CAIR_TRY(pipe_finish(h, slot, L));
  }
  CAIR_CUDA(cudaEventSynchronize(p.ev_done));
  p.busy = false;
  const int flags = *p.err;
  *p.err = 0;
  if (flags & ERRF_BAD_TOKEN) return fail(CAIR_ERR_BAD_ARG, "token id outside [0, vocab)");
  if (flags) return fail(CAIR_ERR_BAD_ARG, "sequence length outside [1, padded length]");
  return CAIR_OK;
}

// Debug: with on != 0 the pipelined submits record timing events; cair_ranker_pipeline_trace returns, for a slot whose
// wait has returned, the times in ms relative to the begin of its encoder: encoder end, part 1 begin / end,
// part 2 begin / end, finish (scores on the host).
extern "C" __attribute__((visibility("default"))) int32_t cair_ranker_pipeline_trace(cair_handle* h, int32_t on, int32_t slot,
                                                                                      float* ms6) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  h->pipe_trace = on != 0;
  if (!ms6 || slot < 0 || slot > 2) return CAIR_OK;
  cair_handle::PipeSlot& p = h->pipe[slot];
  for (int i = 1; i <= 6; ++i) {
    ms6[i - 1] = -1.f;
    if (p.tr[0] && p.tr[i] && cudaEventElapsedTime(&ms6[i - 1], p.tr[0], p.tr[i]) != cudaSuccess) {
      ms6[i - 1] = -1.f;
      cudaGetLastError();
    }
  }
  return CAIR_OK;
}

int32_t cair_ranker_set_gather(cair_handle* h, const uint64_t* peer_recv0, const uint64_t* peer_recv1, const uint64_t* peer_flags,
                               int32_t rank, int32_t world, int64_t count) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  cair_handle::Gather& g = h->gather;
  if (world <= 1 || !peer_recv0) {   // switch off
    g.on = false;
    return CAIR_OK;
  }
  if (!peer_recv1 || !peer_flags || rank < 0 || rank >= world || world > 16 || count <= 0)
    return fail(CAIR_ERR_BAD_ARG, "ranker_set_gather: bad argument");
  for (int k = 0; k < 3; ++k)
    if (h->pipe[k].busy) return fail(CAIR_ERR_BAD_ARG, "ranker_set_gather: batches in flight");
  g.recv[0].assign(peer_recv0, peer_recv0 + world);
  g.recv[1].assign(peer_recv1, peer_recv1 + world);
  g.flags.assign(peer_flags, peer_flags + world);
  g.rank = rank, g.world = world, g.count = count, g.seq = 0, g.on = true;
  return CAIR_OK;
}

int32_t cair_ranker_set_pipeline_split(cair_handle* h, float frac) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "null handle");
  if (!(frac >= 0.f && frac <= 0.9f)) return fail(CAIR_ERR_BAD_ARG, "pipeline split must be in [0, 0.9]");
  h->pipe_frac = frac;
  return CAIR_OK;
}

// ---- CARS ----------------------------------------------------------------------------------------
int32_t cair_cars_workspace_bytes(cair_handle* h, int32_t B, int32_t S, int32_t N, int32_t Lq, int32_t Ld,
                                  size_t* bytes) {
  if (!h || h->model != CAIR_MODEL_CARS || !bytes) return fail(CAIR_ERR_BAD_ARG, "cars_workspace_bytes: bad argument");
  if (B <= 0 || S <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_SHAPE, "B, S, N, Lq, Ld must be positive");
  Arena a(nullptr, 0);
  CarsIO io{};
  CAIR_TRY(cars_forward(h->cars, io, B, S, N, Lq, Ld, 0, B, a, h->d_err, 0, true));
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_cars_forward_ex(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                             const int64_t* dlen, const float* labels, int32_t B, int32_t S, int32_t N, int32_t Lq,
                             int32_t Ld, int32_t session_begin, int32_t session_count, float* scores,
                             const cair_cars_outputs* outs, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || h->model != CAIR_MODEL_CARS) return fail(CAIR_ERR_BAD_ARG, "cars_forward: not a CARS handle");
  if (B <= 0 || S <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_SHAPE, "B, S, N, Lq, Ld must be positive");
  if (!q || !qlen || !d || !dlen || !labels || !scores) return fail(CAIR_ERR_BAD_ARG, "cars_forward: null tensor");
  if (session_begin < 0 || session_count < 0 || session_begin + session_count > B)
    return fail(CAIR_ERR_BAD_ARG, "cars_forward: session slice outside B");
  if (((uintptr_t)workspace & 255) != 0) return fail(CAIR_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  DeviceGuard g(h->device);
  CarsIO io{q, qlen, d, dlen, labels, scores, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (outs) {
    io.pooled_q = outs->pooled_q, io.pooled_d = outs->pooled_d, io.clicks = outs->clicks;
    io.sess_q_attn = outs->sess_q_attn, io.sess_d_attn = outs->sess_d_attn;
    io.enc_q = outs->enc_q, io.sess_h = outs->sess_h, io.sess_c = outs->sess_c;
  }
  Arena probe(nullptr, 0);
  CAIR_TRY(cars_forward(h->cars, io, B, S, N, Lq, Ld, session_begin, session_count, probe, h->d_err, 0, true));
  if (probe.off > workspace_bytes || (probe.off > 0 && !workspace))
    return fail(CAIR_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", probe.off, workspace_bytes);
  Arena ws(workspace, workspace_bytes);
  h->prof.reset();
  g_prof = &h->prof;
  prof_mark("begin", (cudaStream_t)stream);
  int32_t rc = cars_forward(h->cars, io, B, S, N, Lq, Ld, session_begin, session_count, ws, h->d_err, (cudaStream_t)stream, false);
  prof_mark("end", (cudaStream_t)stream);
  g_prof = nullptr;
  return rc;
}

int32_t cair_cars_forward(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                          const int64_t* dlen, const float* labels, int32_t B, int32_t S, int32_t N, int32_t Lq,
                          int32_t Ld, int32_t session_begin, int32_t session_count, float* scores, float* pooled_q,
                          float* pooled_d, float* clicks, float* sess_q_attn, float* sess_d_attn, void* workspace,
                          size_t workspace_bytes, void* stream) {
  cair_cars_outputs o{pooled_q, pooled_d, clicks, sess_q_attn, sess_d_attn, nullptr, nullptr, nullptr};
  return cair_cars_forward_ex(h, q, qlen, d, dlen, labels, B, S, N, Lq, Ld, session_begin, session_count, scores, &o, workspace,
                              workspace_bytes, stream);
}

int32_t cair_cars_set_decoder(cair_handle* h, const cair_cars_decoder_weights* w) {
  if (!h || h->model != CAIR_MODEL_CARS || !w) return fail(CAIR_ERR_BAD_ARG, "cars_set_decoder: not a CARS handle");
  DeviceGuard g(h->device);
  CAIR_TRY(cars_set_decoder(h->own, &h->cars, *w, 0));
  CAIR_CUDA(cudaStreamSynchronize(0));
  return CAIR_OK;
}

int32_t cair_cars_decode_workspace_bytes(cair_handle* h, int32_t B, int32_t S, int32_t Lq, size_t* bytes) {
  if (!h || h->model != CAIR_MODEL_CARS || !bytes) return fail(CAIR_ERR_BAD_ARG, "cars_decode_workspace_bytes: bad argument");
  if (!h->cars.dec.ready) return fail(CAIR_ERR_BAD_ARG, "cars_decode: no decoder weights (cair_cars_set_decoder)");
  if (B <= 0 || S <= 0 || Lq <= 0) return fail(CAIR_ERR_BAD_SHAPE, "B, S, Lq must be positive");
  *bytes = cars_decode_workspace_bytes(h->cars, B, S, Lq);
  return CAIR_OK;
}

int32_t cair_cars_decode(cair_handle* h, const float* enc_q, const int64_t* qlen, const float* sess_h, const float* sess_c,
                         const float* sess_q_attn, const float* sess_d_attn, int32_t B, int32_t S, int32_t Lq,
                         int32_t max_len, const int64_t* tgt2src, int64_t bos_id, int64_t* predictions, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!h || h->model != CAIR_MODEL_CARS) return fail(CAIR_ERR_BAD_ARG, "cars_decode: not a CARS handle");
  if (!enc_q || !qlen || !sess_h || !sess_c || !sess_q_attn || !sess_d_attn || !tgt2src || !predictions)
    return fail(CAIR_ERR_BAD_ARG, "cars_decode: null tensor");
  if (B <= 0 || S <= 0 || Lq <= 0 || max_len < 0) return fail(CAIR_ERR_BAD_SHAPE, "B, S, Lq must be positive");
  if (((uintptr_t)workspace & 255) != 0) return fail(CAIR_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  DeviceGuard g(h->device);
  h->prof.reset();
  g_prof = &h->prof;
  int32_t rc = cars_decode(h->cars, enc_q, qlen, sess_h, sess_c, sess_q_attn, sess_d_attn, B, S, Lq, max_len, tgt2src, bos_id,
                           predictions, workspace, workspace_bytes, h->d_err, (cudaStream_t)stream);
  prof_mark("end", (cudaStream_t)stream);
  g_prof = nullptr;
  return rc;
}

}  // extern "C"
