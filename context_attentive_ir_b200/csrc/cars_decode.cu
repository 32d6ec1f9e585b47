// CARS query-suggestion decoder: greedy decode of the next query for every (session, query) row
// (neuroir/multitask/cars.py:706-791 decode; :605-657 the decoder-side modules; decoders/rnn_decoder.py:19-90 RNNDecoder
// step; modules/global_attention.py:121-211 GlobalAttention 'general'; decoders/decoder.py:118-155: the decoder state
// object is updated in place by every call, so the LSTM state carries from token to token).
// Row conventions are the reference's, including its mixed orderings (reproduced, not fixed): the initial states are
// concatenated query-index-major (torch.cat(hidden_states[:-1], dim=1), cars.py:440-453: row i = s*B + b) while the
// memory banks, their lengths, the session summaries and the predictions are batch-major (row i = b*(S-1) + s,
// cars.py:724-731,747-757,786).
// One decode step = a handful of small GEMMs (rows = B*(S-1), at most a few hundred: latency-sized, fp32 CUDA cores)
// around three small kernels: LSTM cell, attention over the <= Lq memory rows, arg-max + target->source id map.
#include "models.cuh"

namespace cair {

__global__ void dec_add_kernel(const float* a, const float* b, float* o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

int32_t cars_set_decoder(Owned& own, CarsState* st, const cair_cars_decoder_weights& w, cudaStream_t s) {
  CarsDecoder& d = st->dec;
  const int Hs = st->Hsq + st->Hsd, Hdec = w.nhid_decoder, Vt = w.tgt_vocab, E = st->E;
  if (Hdec <= 0 || Vt <= 0) return fail(CAIR_ERR_BAD_ARG, "cars_set_decoder: bad sizes");
  if (!w.transform_hid.w || !w.transform_hid.b || !w.transform_cell.w || !w.transform_cell.b || !w.rnn.w_ih || !w.rnn.w_hh ||
      !w.rnn.b_ih || !w.rnn.b_hh || !w.attn_in.w || !w.attn_out.w || !w.dec_attn.w || !w.predictor1.w || !w.predictor2.w ||
      !w.private_session_projector2.w)
    return fail(CAIR_ERR_BAD_ARG, "cars_set_decoder: null weight pointer");
  d.Hdec = Hdec, d.Vt = Vt;
  CAIR_TRY(dev_copy(own, w.transform_hid.w, (size_t)Hdec * Hs, &d.th_w, s));
  CAIR_TRY(dev_copy(own, w.transform_hid.b, (size_t)Hdec, &d.th_b, s));
  CAIR_TRY(dev_copy(own, w.transform_cell.w, (size_t)Hdec * Hs, &d.tc_w, s));
  CAIR_TRY(dev_copy(own, w.transform_cell.b, (size_t)Hdec, &d.tc_b, s));
  CAIR_TRY(dev_copy(own, w.rnn.w_ih, (size_t)4 * Hdec * E, &d.w_ih, s));
  CAIR_TRY(dev_copy(own, w.rnn.w_hh, (size_t)4 * Hdec * Hdec, &d.w_hh, s));
  CAIR_CUDA(own.alloc(&d.bias, (size_t)4 * Hdec));
  CAIR_LAUNCH(dec_add_kernel, (4 * Hdec + 255) / 256, 256, 0, s, w.rnn.b_ih, w.rnn.b_hh, d.bias, (int64_t)4 * Hdec);
  CAIR_TRY(dev_copy(own, w.attn_in.w, (size_t)Hdec * Hdec, &d.attn_in, s));
  CAIR_TRY(dev_copy(own, w.attn_out.w, (size_t)Hdec * 2 * Hdec, &d.attn_out, s));
  CAIR_TRY(dev_copy(own, w.dec_attn.w, (size_t)Hdec * st->Hq, &d.dec_attn, s));
  CAIR_TRY(dev_copy(own, w.predictor1.w, (size_t)st->Hd * Hdec, &d.pred1, s));
  CAIR_TRY(dev_copy(own, w.predictor2.w, (size_t)Vt * st->Hd, &d.pred2, s));
  // both bias-free projectors see the same session summary (cars.py:766-768)
  const int64_t np = (int64_t)st->Hd * Hs;
  CAIR_CUDA(own.alloc(&d.sess_proj2, (size_t)np));
  CAIR_LAUNCH(dec_add_kernel, (unsigned)((np + 255) / 256), 256, 0, s, st->shared_proj, w.private_session_projector2.w,
              d.sess_proj2, np);
  d.ready = true;
  return CAIR_OK;
}

// initial-state rows (query-index-major, row i = s*B + b) and session-summary rows (batch-major, row i = b*(S-1) + s)
__global__ void dec_gather_kernel(const float* __restrict__ sess_h, const float* __restrict__ sess_c,
                                  const float* __restrict__ sqa, const float* __restrict__ sda, const int64_t* __restrict__ qlen,
                                  int B, int S, int Hsq, int Hsd, float* __restrict__ h_rows, float* __restrict__ c_rows,
                                  float* __restrict__ sum_rows, int* __restrict__ memlen, int* __restrict__ memrow) {
  const int Hs = Hsq + Hsd, R = B * (S - 1);
  const int i = blockIdx.x;
  if (i >= R) return;
  const int s1 = i / B, b1 = i - s1 * B;            // state rows
  const int b2 = i / (S - 1), s2 = i - b2 * (S - 1);   // memory / summary rows
  for (int k = threadIdx.x; k < Hs; k += blockDim.x) {
    h_rows[(size_t)i * Hs + k] = sess_h[((size_t)b1 * S + s1) * Hs + k];
    c_rows[(size_t)i * Hs + k] = sess_c[((size_t)b1 * S + s1) * Hs + k];
    sum_rows[(size_t)i * Hs + k] = k < Hsq ? sqa[((size_t)b2 * S + s2) * Hsq + k] : sda[((size_t)b2 * S + s2) * Hsd + (k - Hsq)];
  }
  if (threadIdx.x == 0) {
    memlen[i] = (int)qlen[b2 * S + s2];
    memrow[i] = b2 * S + s2;
  }
}

// torch.nn.LSTM step, gate order i,f,g,o; gates = gx (input part + both biases) + gh (hidden part); state in place
__global__ void dec_cell_kernel(const float* __restrict__ gx, const float* __restrict__ gh, int R, int H, float* __restrict__ h,
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

// GlobalAttention 'general' for one target step (global_attention.py:150-193): align_s = (W_in h) . m_s, positions
// s >= memory_len masked to -inf, softmax, context = sum_s a_s m_s; writes [context, h] for linear_out.  One CTA per row.
__global__ void __launch_bounds__(128) dec_attn_kernel(const float* __restrict__ hq, const float* __restrict__ h,
                                                       const float* __restrict__ mb, const int* __restrict__ memlen,
                                                       const int* __restrict__ memrow, int Lq, int H, float* __restrict__ cat) {
  extern __shared__ float sc[];   // [Lq]
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* m = mb + (size_t)memrow[i] * Lq * H;
  int l = memlen[i];
  l = l < 1 ? 1 : (l > Lq ? Lq : l);
  for (int s = warp; s < Lq; s += 4) {
    float a = 0.f;
    for (int k = lane; k < H; k += 32) a = fmaf(hq[(size_t)i * H + k], m[(size_t)s * H + k], a);
    a = warp_sum(a);
    if (lane == 0) sc[s] = s < l ? a : -INFINITY;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int s = 0; s < Lq; ++s) mx = fmaxf(mx, sc[s]);
  float den = 0.f;
  for (int s = 0; s < l; ++s) den += expf(sc[s] - mx);
  const float inv = 1.0f / den;
  for (int k = tid; k < H; k += 128) {
    float a = 0.f;
    for (int s = 0; s < l; ++s) a = fmaf(expf(sc[s] - mx) * inv, m[(size_t)s * H + k], a);
    cat[(size_t)i * 2 * H + k] = a;
    cat[(size_t)i * 2 * H + H + k] = h[(size_t)i * H + k];
  }
}

__global__ void dec_addrows_kernel(float* __restrict__ o, const float* __restrict__ add, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] += add[i];
}

// arg-max over the target vocabulary (first maximum, as torch.max), prediction store, next input token through the
// target -> source id map (cars.py:774-783: tgt_dict[idx] -> word -> src_dict[word])
__global__ void __launch_bounds__(256) dec_argmax_kernel(const float* __restrict__ logits, int Vt, const int64_t* __restrict__ tgt2src,
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
    if (tid < o) {
      if (bv[tid + o] > bv[tid] || (bv[tid + o] == bv[tid] && bi[tid + o] < bi[tid])) bv[tid] = bv[tid + o], bi[tid] = bi[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const int a = bi[0] == 0x7fffffff ? 0 : bi[0];
    pred[(size_t)i * max_len + t] = a;
    next[i] = tgt2src[a];
  }
}

__global__ void dec_fill_kernel(int64_t* p, int n, int64_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct DecWs {
  float *h_rows, *c_rows, *sum_rows, *h, *c, *mb, *sess, *gx, *gh, *hq, *cat, *ah, *o1, *logits;
  int *memlen, *memrow;
  int64_t* tok;
};
static void dec_carve(const CarsState& st, int B, int S, int Lq, Arena& ws, DecWs* o) {
  const CarsDecoder& d = st.dec;
  const size_t R = (size_t)B * (S - 1), Hs = st.Hsq + st.Hsd, H = d.Hdec;
  o->h_rows = ws.take<float>(R * Hs), o->c_rows = ws.take<float>(R * Hs), o->sum_rows = ws.take<float>(R * Hs);
  o->h = ws.take<float>(R * H), o->c = ws.take<float>(R * H);
  o->mb = ws.take<float>((size_t)B * S * Lq * H);
  o->sess = ws.take<float>(R * st.Hd);
  o->gx = ws.take<float>(R * 4 * H), o->gh = ws.take<float>(R * 4 * H);
  o->hq = ws.take<float>(R * H), o->cat = ws.take<float>(R * 2 * H), o->ah = ws.take<float>(R * H);
  o->o1 = ws.take<float>(R * st.Hd), o->logits = ws.take<float>(R * d.Vt);
  o->memlen = ws.take<int>(R), o->memrow = ws.take<int>(R);
  o->tok = ws.take<int64_t>(R);
}
size_t cars_decode_workspace_bytes(const CarsState& st, int B, int S, int Lq) {
  Arena probe(nullptr, 0);
  DecWs w;
  dec_carve(st, B, S, Lq, probe, &w);
  return probe.off + 256;
}

int32_t cars_decode(const CarsState& st, const float* enc_q, const int64_t* qlen, const float* sess_h, const float* sess_c,
                    const float* sess_q_attn, const float* sess_d_attn, int B, int S, int Lq, int max_len, const int64_t* tgt2src,
                    int64_t bos, int64_t* predictions, void* wsp, size_t ws_bytes, int* err, cudaStream_t s) {
  const CarsDecoder& d = st.dec;
  if (!d.ready) return fail(CAIR_ERR_BAD_ARG, "cars_decode: no decoder weights (cair_cars_set_decoder)");
  if (S < 2 || max_len < 1) return CAIR_OK;   // nothing to suggest (session_len - 1 == 0)
  const int R = B * (S - 1), Hs = st.Hsq + st.Hsd, H = d.Hdec;
  Arena ws(wsp, ws_bytes);
  DecWs w;
  dec_carve(st, B, S, Lq, ws, &w);
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "cars_decode: workspace too small");
  prof_mark("decode_setup", s);
  CAIR_LAUNCH(dec_gather_kernel, (unsigned)R, 128, 0, s, sess_h, sess_c, sess_q_attn, sess_d_attn, qlen, B, S, st.Hsq, st.Hsd,
              w.h_rows, w.c_rows, w.sum_rows, w.memlen, w.memrow);
  // decoder initial state: transform_hid / transform_cell of the concatenated session-encoder states (cars.py:440-453)
  CAIR_TRY(gemm_f32(gemm_dense(w.h_rows, Hs), d.th_w, d.th_b, w.h, H, R, H, Hs, ACT_NONE, s));
  CAIR_TRY(gemm_f32(gemm_dense(w.c_rows, Hs), d.tc_w, d.tc_b, w.c, H, R, H, Hs, ACT_NONE, s));
  // memory banks through dec_attn (cars.py:759), all B*S queries at once (the rows with s = S-1 are never read)
  CAIR_TRY(gemm_f32(gemm_dense(enc_q, st.Hq), d.dec_attn, nullptr, w.mb, H, (int64_t)B * S * Lq, H, st.Hq, ACT_NONE, s));
  // session summary added to every step's predictor input (cars.py:766-769)
  CAIR_TRY(gemm_f32(gemm_dense(w.sum_rows, Hs), d.sess_proj2, nullptr, w.sess, st.Hd, R, st.Hd, Hs, ACT_NONE, s));
  CAIR_LAUNCH(dec_fill_kernel, (R + 255) / 256, 256, 0, s, w.tok, R, bos);
  prof_mark("decode_steps", s);
  for (int t = 0; t < max_len; ++t) {
    CAIR_TRY(gemm_f32(gemm_gather(st.table, st.V, st.E, w.tok, 1, 1, 1, err), d.w_ih, d.bias, w.gx, 4 * H, R, 4 * H, st.E, ACT_NONE, s));
    CAIR_TRY(gemm_f32(gemm_dense(w.h, H), d.w_hh, nullptr, w.gh, 4 * H, R, 4 * H, H, ACT_NONE, s));
    CAIR_LAUNCH(dec_cell_kernel, (unsigned)(((int64_t)R * H + 255) / 256), 256, 0, s, w.gx, w.gh, R, H, w.h, w.c);
    CAIR_TRY(gemm_f32(gemm_dense(w.h, H), d.attn_in, nullptr, w.hq, H, R, H, H, ACT_NONE, s));
    CAIR_LAUNCH(dec_attn_kernel, (unsigned)R, 128, (size_t)Lq * sizeof(float), s, w.hq, w.h, w.mb, w.memlen, w.memrow, Lq, H, w.cat);
    CAIR_TRY(gemm_f32(gemm_dense(w.cat, 2 * H), d.attn_out, nullptr, w.ah, H, R, H, 2 * H, ACT_TANH, s));
    CAIR_TRY(gemm_f32(gemm_dense(w.ah, H), d.pred1, nullptr, w.o1, st.Hd, R, st.Hd, H, ACT_NONE, s));
    CAIR_LAUNCH(dec_addrows_kernel, (unsigned)(((int64_t)R * st.Hd + 255) / 256), 256, 0, s, w.o1, w.sess, (int64_t)R * st.Hd);
    CAIR_TRY(gemm_f32(gemm_dense(w.o1, st.Hd), d.pred2, nullptr, w.logits, d.Vt, R, d.Vt, st.Hd, ACT_NONE, s));
    CAIR_LAUNCH(dec_argmax_kernel, (unsigned)R, 256, 0, s, w.logits, d.Vt, tgt2src, predictions, max_len, t, w.tok);
  }
  return CAIR_OK;
}

}  // namespace cair
