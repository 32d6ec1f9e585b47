/*
 * cair_oracle.c - TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's scoring forward passes (eval mode), used only
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
 * the checker for the CUDA library.  Nothing under context_attentive_ir_b200/ may link,
 * import or call it.
 *
 * Pinned against the UNMODIFIED reference: the tests/golden npz files hold inputs, state_dicts and
 * outputs produced by importing the reference's own modules from /root/reference
 * (oracle/gen_golden.py); tests/test_oracle_golden.py checks every function below against
 * them.  The reference itself ships no golden vectors or tests; the
 * arithmetic lives in torch 2.11 (ATen CPU) and numpy 2.3 - not vendored, no lockfile - so
 * operator semantics (nn.LSTM gate order, Conv2d cross-correlation, cosine_similarity
 * eps clamp, numpy.histogram bin rule) are restated from their documentation and pinned
 * operationally by those fixtures.
 *
 * Weight structs are the ones of include/cair.h but with HOST pointers.
 * Dot products accumulate in double and round to float once; everything else is fp32.
 * Pairs / sequences are processed in parallel with OpenMP when compiled with -fopenmp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cair.h"

#define ORA_API __attribute__((visibility("default")))

static float dotf(const float* a, const float* b, int n) {
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += (double)a[k] * (double)b[k];
  return (float)s;
}

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* modules/embeddings.py:243-252 + util_class.py:42-53: out[t,:] = table[ids[t],:] */
ORA_API int cair_oracle_embed(const float* table, int V, int E, const int64_t* ids, int64_t T,
                              float* out) {
  for (int64_t t = 0; t < T; ++t) {
    if (ids[t] < 0 || ids[t] >= V) return CAIR_ERR_BAD_ARG;
    memcpy(out + t * E, table + ids[t] * (int64_t)E, sizeof(float) * E);
  }
  return CAIR_OK;
}

/* nn.Linear over rows: y[r,:] = x[r,:] W^T + b   (W [out,in]) */
static void linear_rows(const float* x, int64_t rows, int in, const cair_linear* l, int out,
                        float* y) {
  for (int64_t r = 0; r < rows; ++r)
    for (int o = 0; o < out; ++o)
      y[r * out + o] = dotf(x + r * in, l->w + (int64_t)o * in, in) + (l->b ? l->b[o] : 0.0f);
}

/* One LSTM cell step, torch.nn.LSTM semantics (gate rows i,f,g,o; both biases added). */
static void lstm_step(const float* x, int in, int h, const cair_lstm_dir* w, float* hs, float* cs,
                      float* gates) {
  for (int r = 0; r < 4 * h; ++r)
    gates[r] = dotf(x, w->w_ih + (int64_t)r * in, in) + w->b_ih[r] +
               dotf(hs, w->w_hh + (int64_t)r * h, h) + w->b_hh[r];
  for (int u = 0; u < h; ++u) {
    float ig = sigmoidf_(gates[u]), fg = sigmoidf_(gates[h + u]);
    float gg = tanhf(gates[2 * h + u]);
    cs[u] = fg * cs[u] + ig * gg;
  }
  for (int u = 0; u < h; ++u) hs[u] = sigmoidf_(gates[3 * h + u]) * tanhf(cs[u]);
}

/* One GRU cell step, torch.nn.GRU semantics (gate rows r,z,n; n = tanh(W_in x + b_in + r * (W_hn h + b_hn))). */
static void gru_step(const float* x, int in, int h, const cair_lstm_dir* w, float* hs, float* gates) {
  for (int u = 0; u < h; ++u) {
    float ar = dotf(x, w->w_ih + (int64_t)u * in, in) + w->b_ih[u] + dotf(hs, w->w_hh + (int64_t)u * h, h) + w->b_hh[u];
    float az = dotf(x, w->w_ih + (int64_t)(h + u) * in, in) + w->b_ih[h + u] + dotf(hs, w->w_hh + (int64_t)(h + u) * h, h) + w->b_hh[h + u];
    float xn = dotf(x, w->w_ih + (int64_t)(2 * h + u) * in, in) + w->b_ih[2 * h + u];
    float hn = dotf(hs, w->w_hh + (int64_t)(2 * h + u) * h, h) + w->b_hh[2 * h + u];
    float r = sigmoidf_(ar), z = sigmoidf_(az);
    float nn = tanhf(xn + r * hn);
    gates[u] = (1.0f - z) * nn + z * hs[u];
  }
  memcpy(hs, gates, sizeof(float) * h);
}

/* RNNEncoder.forward with lengths (encoders/rnn_encoder.py:62-141), one LSTM layer.
 * sort/pack/unpack/unsort is a per-sequence no-op: each sequence runs over its own
 * len steps, the reverse direction starts at its own last token, outputs at t >= len are
 * zero (the zero-pad of :135-139 and pad_packed_sequence). */
static int oracle_rnn(int rnn_type, const float* x, const int64_t* len, int n, int L, int in, int h,
                      const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out, float* h_n, float* c_n);

ORA_API int cair_oracle_lstm(const float* x, const int64_t* len, int n, int L, int in, int h,
                             const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out,
                             float* h_n, float* c_n) {
  return oracle_rnn(CAIR_RNN_LSTM, x, len, n, L, in, h, fwd, rev, out, h_n, c_n);
}

/* Same with the cell type as an argument (getattr(nn, rnn_type), encoders/rnn_encoder.py:45-53); c_n is left
 * untouched for GRU. */
ORA_API int cair_oracle_rnn(int rnn_type, const float* x, const int64_t* len, int n, int L, int in, int h,
                            const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out,
                            float* h_n, float* c_n) {
  if (rnn_type != CAIR_RNN_LSTM && rnn_type != CAIR_RNN_GRU) return CAIR_ERR_UNSUPPORTED;
  return oracle_rnn(rnn_type, x, len, n, L, in, h, fwd, rev, out, h_n, rnn_type == CAIR_RNN_GRU ? NULL : c_n);
}

static int oracle_rnn(int rnn_type, const float* x, const int64_t* len, int n, int L, int in, int h,
                      const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out, float* h_n, float* c_n) {
  int dirs = rev ? 2 : 1;
  for (int s = 0; s < n; ++s)
    if (len[s] < 1 || len[s] > L) return CAIR_ERR_BAD_ARG;
  memset(out, 0, sizeof(float) * (size_t)n * L * dirs * h);
#pragma omp parallel for schedule(dynamic, 4)
  for (int s = 0; s < n; ++s) {
    float* hs = (float*)malloc(sizeof(float) * h * 6);
    float *cs = hs + h, *gates = hs + 2 * h;
    int T = (int)len[s];
    for (int dir = 0; dir < dirs; ++dir) {
      const cair_lstm_dir* w = dir ? rev : fwd;
      memset(hs, 0, sizeof(float) * 2 * h);
      for (int k = 0; k < T; ++k) {
        int t = dir ? T - 1 - k : k;
        if (rnn_type == CAIR_RNN_GRU)
          gru_step(x + ((int64_t)s * L + t) * in, in, h, w, hs, gates);
        else
          lstm_step(x + ((int64_t)s * L + t) * in, in, h, w, hs, cs, gates);
        memcpy(out + ((int64_t)s * L + t) * dirs * h + dir * h, hs, sizeof(float) * h);
      }
      if (h_n) memcpy(h_n + ((int64_t)dir * n + s) * h, hs, sizeof(float) * h);
      if (c_n) memcpy(c_n + ((int64_t)dir * n + s) * h, cs, sizeof(float) * h);
    }
    free(hs);
  }
  return CAIR_OK;
}

/* torch>=2 F.cosine_similarity: normalise each vector by max(||x||, eps) first, then dot. */
static float cosine_norm_first(const float* a, const float* b, int n, float eps) {
  float na = sqrtf(dotf(a, a, n)), nb = sqrtf(dotf(b, b, n));
  if (na < eps) na = eps;
  if (nb < eps) nb = eps;
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += (double)(a[k] / na) * (double)(b[k] / nb);
  return (float)s;
}

static int check_ids(const int64_t* ids, int64_t n, int V) {
  for (int64_t i = 0; i < n; ++i)
    if (ids[i] < 0 || ids[i] >= V) return 0;
  return 1;
}

/* ---- ESM (rankers/esm.py:34-44): mean over the PADDED length, then cosine ---------------- */
ORA_API int cair_oracle_esm(const cair_esm_weights* w, const int64_t* q, const int64_t* qlen,
                            const int64_t* d, const int64_t* dlen, int B, int N, int Lq, int Ld,
                            float* scores) {
  (void)qlen;
  (void)dlen;
  int E = w->emsize;
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, (int64_t)B * N * Ld, w->vocab))
    return CAIR_ERR_BAD_ARG;
#pragma omp parallel for
  for (int b = 0; b < B; ++b) {
    float* vq = (float*)calloc(2 * (size_t)E, sizeof(float));
    float* vd = vq + E;
    for (int t = 0; t < Lq; ++t)
      for (int k = 0; k < E; ++k) vq[k] += w->table[q[b * Lq + t] * E + k];
    for (int k = 0; k < E; ++k) vq[k] /= (float)Lq;
    for (int n = 0; n < N; ++n) {
      memset(vd, 0, sizeof(float) * E);
      const int64_t* dd = d + ((int64_t)b * N + n) * Ld;
      for (int t = 0; t < Ld; ++t)
        for (int k = 0; k < E; ++k) vd[k] += w->table[dd[t] * E + k];
      for (int k = 0; k < E; ++k) vd[k] /= (float)Ld;
      scores[b * N + n] = cosine_norm_first(vq, vd, E, 1e-8f);
    }
    free(vq);
  }
  return CAIR_OK;
}

/* ---- Match-Tensor (rankers/mtensor.py:62-131; exact match :144-158) ----------------------- */
/* Stacked encoders (encoders/rnn_encoder.py:45-53, :92-113): layer k >= 1 reads the memory bank of layer k-1 (the dropout between
 * them is the identity in eval mode; use_last = True keeps the top bank only).  xq / xd: [nextra][fwd, rev] weights of the
 * query / document encoder's layers 1..nextra (rev ignored when unidirectional). */
static int oracle_rnn_stack(int rnn_type, float** bank, const int64_t* len, int n, int L, int H, int dirs,
                            const cair_lstm_dir* extra, int nextra) {
  for (int k = 0; k < nextra; ++k) {
    float* next = (float*)malloc(sizeof(float) * (size_t)n * L * H);
    int rc = oracle_rnn(rnn_type, *bank, len, n, L, H, H / dirs, &extra[2 * k], dirs == 2 ? &extra[2 * k + 1] : NULL, next, NULL, NULL);
    free(*bank);
    *bank = next;
    if (rc != CAIR_OK) return rc;
  }
  return CAIR_OK;
}

ORA_API int cair_oracle_mt_stacked(const cair_mt_weights* w, const cair_lstm_dir* xq, const cair_lstm_dir* xd, int nextra,
                                   const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen, int B, int N,
                                   int Lq, int Ld, float* scores, float* enc_q_out, float* enc_d_out);
ORA_API int cair_oracle_mt(const cair_mt_weights* w, const int64_t* q, const int64_t* qlen,
                           const int64_t* d, const int64_t* dlen, int B, int N, int Lq, int Ld,
                           float* scores, float* enc_q_out, float* enc_d_out) {
  return cair_oracle_mt_stacked(w, NULL, NULL, 0, q, qlen, d, dlen, B, N, Lq, Ld, scores, enc_q_out, enc_d_out);
}

ORA_API int cair_oracle_mt_stacked(const cair_mt_weights* w, const cair_lstm_dir* xq, const cair_lstm_dir* xd, int nextra,
                                   const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen, int B, int N,
                                   int Lq, int Ld, float* scores, float* enc_q_out, float* enc_d_out) {
  if (w->rnn_type != CAIR_RNN_LSTM && w->rnn_type != CAIR_RNN_GRU) return CAIR_ERR_UNSUPPORTED;
  int E = w->emsize, F = w->featsize, C = w->nchannels, nf = w->nfilters,
      M = w->match_filter_size;
  int dirs = w->bidirectional ? 2 : 1;
  int Hq = w->nhid_query, Hd = w->nhid_doc, hq = Hq / dirs, hd = Hd / dirs;
  int64_t BN = (int64_t)B * N;
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, BN * Ld, w->vocab))
    return CAIR_ERR_BAD_ARG;
  /* :77-90 embedding + linear projection (pads project to the bias) */
  float* eq = (float*)malloc(sizeof(float) * B * Lq * E);
  float* ed = (float*)malloc(sizeof(float) * BN * Ld * E);
  cair_oracle_embed(w->table, w->vocab, E, q, (int64_t)B * Lq, eq);
  cair_oracle_embed(w->table, w->vocab, E, d, BN * Ld, ed);
  float* pq = (float*)malloc(sizeof(float) * B * Lq * F);
  float* pd = (float*)malloc(sizeof(float) * BN * Ld * F);
  linear_rows(eq, (int64_t)B * Lq, E, &w->linear_projection, F, pq);
  linear_rows(ed, BN * Ld, E, &w->linear_projection, F, pd);
  free(eq);
  free(ed);
  /* :93-94 separate query / document encoders */
  float* hq_ = (float*)malloc(sizeof(float) * B * Lq * Hq);
  float* hd_ = (float*)malloc(sizeof(float) * BN * Ld * Hd);
  int rc = oracle_rnn(w->rnn_type, pq, qlen, B, Lq, F, hq, &w->query_fwd,
                      dirs == 2 ? &w->query_rev : NULL, hq_, NULL, NULL);
  if (rc == CAIR_OK)
    rc = oracle_rnn(w->rnn_type, pd, dlen, (int)BN, Ld, F, hd, &w->doc_fwd,
                    dirs == 2 ? &w->doc_rev : NULL, hd_, NULL, NULL);
  free(pq);
  free(pd);
  if (rc == CAIR_OK) rc = oracle_rnn_stack(w->rnn_type, &hq_, qlen, B, Lq, Hq, dirs, xq, nextra);
  if (rc == CAIR_OK) rc = oracle_rnn_stack(w->rnn_type, &hd_, dlen, (int)BN, Ld, Hd, dirs, xd, nextra);
  if (rc != CAIR_OK) {
    free(hq_);
    free(hd_);
    return rc;
  }
  if (enc_q_out) memcpy(enc_q_out, hq_, sizeof(float) * B * Lq * Hq);
  if (enc_d_out) memcpy(enc_d_out, hd_, sizeof(float) * BN * Ld * Hd);
  /* :99,108 channel projections (bias also at pad positions) */
  float* cq = (float*)malloc(sizeof(float) * B * Lq * C);
  float* cd = (float*)malloc(sizeof(float) * BN * Ld * C);
  linear_rows(hq_, (int64_t)B * Lq, Hq, &w->query_projection, C, cq);
  linear_rows(hd_, BN * Ld, Hd, &w->document_projection, C, cd);
  free(hq_);
  free(hd_);
  const float alpha = w->alpha[0];
  const cair_linear* convs[3] = {&w->conv1, &w->conv2, &w->conv3};
  const int C1 = C + 1;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < BN; ++p) {
    int b = (int)(p / N);
    /* :100-119 match tensor [C+1, Lq, Ld] */
    float* mt = (float*)malloc(sizeof(float) * (size_t)C1 * Lq * Ld);
    float* y = (float*)malloc(sizeof(float) * (size_t)3 * nf * Lq * Ld);
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < Lq; ++i)
        for (int j = 0; j < Ld; ++j)
          mt[((size_t)c * Lq + i) * Ld + j] = cq[((size_t)b * Lq + i) * C + c] * cd[(p * Ld + j) * C + c];
    for (int i = 0; i < Lq; ++i)
      for (int j = 0; j < Ld; ++j)
        mt[((size_t)C * Lq + i) * Ld + j] = (q[b * Lq + i] == d[p * Ld + j]) ? alpha : 0.0f;
    /* :122-125 three same-padded convs (cross-correlation) + ReLU.  Same (c, a, bb) summation order
     * per output cell as a naive per-cell loop, arranged with j innermost so the compiler vectorises. */
    double* accb = (double*)malloc(sizeof(double) * (size_t)Lq * Ld);
    for (int k = 0; k < 3; ++k) {
      int kw = 3 + 2 * k, pw = 1 + k;
      for (int f = 0; f < nf; ++f) {
        for (int x = 0; x < Lq * Ld; ++x) accb[x] = convs[k]->b[f];
        for (int c = 0; c < C1; ++c)
          for (int a = 0; a < 3; ++a)
            for (int bb = 0; bb < kw; ++bb) {
              const double wv = convs[k]->w[(((size_t)f * C1 + c) * 3 + a) * kw + bb];
              int jlo = pw - bb > 0 ? pw - bb : 0;
              int jhi = Ld + pw - bb < Ld ? Ld + pw - bb : Ld;
              for (int i = 0; i < Lq; ++i) {
                int ii = i + a - 1;
                if (ii < 0 || ii >= Lq) continue;
                const float* src = mt + ((size_t)c * Lq + ii) * Ld + (bb - pw);
                double* dst = accb + (size_t)i * Ld;
                for (int j = jlo; j < jhi; ++j) dst[j] += wv * (double)src[j];
              }
            }
        float* yo = y + (size_t)(k * nf + f) * Lq * Ld;
        for (int x = 0; x < Lq * Ld; ++x) {
          float v = (float)accb[x];
          yo[x] = v > 0.0f ? v : 0.0f;
        }
      }
    }
    free(accb);
    /* :126-130 1x1 conv (no activation), max over Ld then Lq (pads included), Linear(M,1) */
    double sc = w->output.b[0];
    for (int m = 0; m < M; ++m) {
      float best = -INFINITY;
      for (int i = 0; i < Lq; ++i)
        for (int j = 0; j < Ld; ++j) {
          double s = w->conv.b[m];
          for (int f = 0; f < 3 * nf; ++f)
            s += (double)w->conv.w[m * 3 * nf + f] * (double)y[((size_t)f * Lq + i) * Ld + j];
          if ((float)s > best) best = (float)s;
        }
      sc += (double)w->output.w[m] * (double)best;
    }
    scores[p] = (float)sc;
    free(mt);
    free(y);
  }
  free(cq);
  free(cd);
  return CAIR_OK;
}

/* ---- DRMM (rankers/drmm.py:29-84, gating :87-98) ------------------------------------------
 * numpy.histogram(x, bins=[-1,-.5,0,.5,1,1]): half-open bins [e_k, e_k+1), last bin closed,
 * values outside [-1, 1] dropped.  With the duplicated last edge: bin 3 = [.5, 1), bin 4 = {1}. */
static int drmm_bin(float c) {
  if (!(c >= -1.0f) || c > 1.0f) return -1;
  if (c == 1.0f) return 4;
  if (c < -0.5f) return 0;
  if (c < 0.0f) return 1;
  if (c < 0.5f) return 2;
  return 3;
}

ORA_API int cair_oracle_drmm(const cair_drmm_weights* w, const int64_t* q, const int64_t* qlen,
                             const int64_t* d, const int64_t* dlen, int B, int N, int Lq, int Ld,
                             float* scores, int32_t* hist_out, float* cos_out) {
  (void)qlen;
  (void)dlen;
  if (w->nbins != 5) return CAIR_ERR_UNSUPPORTED;
  int E = w->emsize;
  int64_t BN = (int64_t)B * N;
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, BN * Ld, w->vocab))
    return CAIR_ERR_BAD_ARG;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < BN; ++p) {
    int b = (int)(p / N);
    /* :45-51 gating softmax over ALL Lq positions (pads included) */
    float* g = (float*)malloc(sizeof(float) * Lq);
    float mx = -INFINITY, den = 0.0f;
    for (int i = 0; i < Lq; ++i) {
      g[i] = dotf(w->table + q[b * Lq + i] * E, w->gating.w, E) + w->gating.b[0];
      if (g[i] > mx) mx = g[i];
    }
    for (int i = 0; i < Lq; ++i) {
      g[i] = expf(g[i] - mx);
      den += g[i];
    }
    double acc = 0.0;
    for (int i = 0; i < Lq; ++i) {
      int32_t h5[5] = {0, 0, 0, 0, 0};
      const float* xq = w->table + q[b * Lq + i] * E;
      for (int j = 0; j < Ld; ++j) {
        float c = cosine_norm_first(xq, w->table + d[p * Ld + j] * E, E, 1e-8f);
        if (cos_out) cos_out[(p * Lq + i) * Ld + j] = c;
        int k = drmm_bin(c);
        if (k >= 0) h5[k]++;
      }
      if (hist_out) memcpy(hist_out + (p * Lq + i) * 5, h5, sizeof(h5));
      /* :26,80 ffnn = Linear(5,1) then Linear(1,1), no non-linearity */
      float f0 = w->ffnn0.b[0];
      for (int k = 0; k < 5; ++k) f0 += w->ffnn0.w[k] * (float)h5[k];
      float f1 = w->ffnn1.w[0] * f0 + w->ffnn1.b[0];
      acc += (double)f1 * (double)(g[i] / den);
    }
    scores[p] = w->output.w[0] * (float)acc + w->output.b[0];
    free(g);
  }
  return CAIR_OK;
}

/* ---- DUET (rankers/duet.py:77-121 local, :148-208 distributed, :58 sum) -------------------- */
ORA_API int cair_oracle_duet(const cair_duet_weights* w, const int64_t* q, const int64_t* qlen,
                             const int64_t* d, const int64_t* dlen, int B, int N, int Lq, int Ld,
                             float* scores, float* local_out) {
  (void)qlen;
  (void)dlen;
  int E = w->emsize, nf = w->nfilters, ks = w->dist_filter_size, pool = w->pool_size;
  /* shape algebra the reference asserts through its layer sizes (duet.py:69,73,144) */
  if (Lq != w->max_query_len || Ld != w->max_doc_len || w->local_filter_size != 1 || ks != 3)
    return CAIR_ERR_BAD_SHAPE;
  int64_t BN = (int64_t)B * N;
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, BN * Ld, w->vocab))
    return CAIR_ERR_BAD_ARG;
  int Tq = Lq - ks + 1, Td = Ld - ks + 1, Tp = Td - pool + 1; /* Tp == Ld - pool - 1 */
  /* query side: conv_q -> tanh -> global max -> fc1 -> tanh */
  float* rq = (float*)malloc(sizeof(float) * B * nf);
  for (int b = 0; b < B; ++b) {
    float* mq = (float*)malloc(sizeof(float) * nf);
    for (int f = 0; f < nf; ++f) {
      float best = -INFINITY;
      for (int t = 0; t < Tq; ++t) {
        double s = w->conv_q.b[f];
        for (int k = 0; k < ks; ++k) {
          const float* x = w->table + q[b * Lq + t + k] * E;
          for (int e = 0; e < E; ++e) s += (double)w->conv_q.w[((size_t)f * E + e) * ks + k] * x[e];
        }
        float v = tanhf((float)s);
        if (v > best) best = v;
      }
      mq[f] = best;
    }
    for (int f = 0; f < nf; ++f)
      rq[b * nf + f] = tanhf(dotf(mq, w->dist_fc1.w + (size_t)f * nf, nf) + w->dist_fc1.b[f]);
    free(mq);
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < BN; ++p) {
    int b = (int)(p / N);
    const int64_t* dd = d + p * Ld;
    /* local model: X[j,i] = (d_j == q_i); conv1d(Ld -> nf, k=1) over length Lq */
    float* m1 = (float*)malloc(sizeof(float) * nf * 2);
    float* m2 = m1 + nf;
    for (int f = 0; f < nf; ++f) {
      double s1 = w->local_fc1.b[0];
      for (int i = 0; i < Lq; ++i) {
        double s = w->local_conv1d.b[f];
        for (int j = 0; j < Ld; ++j)
          if (dd[j] == q[b * Lq + i]) s += w->local_conv1d.w[(size_t)f * Ld + j];
        s1 += (double)w->local_fc1.w[i] * (double)tanhf((float)s);
      }
      m1[f] = tanhf((float)s1);
    }
    for (int f = 0; f < nf; ++f)
      m2[f] = tanhf(dotf(m1, w->local_fc2.w + (size_t)f * nf, nf) + w->local_fc2.b[f]);
    float local = tanhf(dotf(m2, w->local_fc3.w, nf) + w->local_fc3.b[0]);
    if (local_out) local_out[p] = local;
    /* distributed model */
    float* cd = (float*)malloc(sizeof(float) * (size_t)nf * Td);
    float* pl = (float*)malloc(sizeof(float) * (size_t)nf * Tp);
    for (int f = 0; f < nf; ++f)
      for (int t = 0; t < Td; ++t) {
        double s = w->conv_d1.b[f];
        for (int k = 0; k < ks; ++k) {
          const float* x = w->table + dd[t + k] * E;
          for (int e = 0; e < E; ++e) s += (double)w->conv_d1.w[((size_t)f * E + e) * ks + k] * x[e];
        }
        cd[(size_t)f * Td + t] = tanhf((float)s);
      }
    for (int f = 0; f < nf; ++f)
      for (int t = 0; t < Tp; ++t) {
        float best = cd[(size_t)f * Td + t];
        for (int k = 1; k < pool; ++k)
          if (cd[(size_t)f * Td + t + k] > best) best = cd[(size_t)f * Td + t + k];
        pl[(size_t)f * Tp + t] = best;
      }
    for (int f = 0; f < nf; ++f) {
      double s2 = w->dist_fc2.b[0];
      for (int t = 0; t < Tp; ++t) {
        double s = w->conv_d2.b[f];
        for (int g = 0; g < nf; ++g) s += (double)w->conv_d2.w[(size_t)f * nf + g] * pl[(size_t)g * Tp + t];
        float rd = tanhf((float)s);
        s2 += (double)w->dist_fc2.w[t] * (double)(rq[b * nf + f] * rd);
      }
      m1[f] = tanhf((float)s2);
    }
    for (int f = 0; f < nf; ++f)
      m2[f] = tanhf(dotf(m1, w->dist_fc3.w + (size_t)f * nf, nf) + w->dist_fc3.b[f]);
    float dist = tanhf(dotf(m2, w->dist_fc4.w, nf) + w->dist_fc4.b[0]);
    scores[p] = local + dist;
    free(m1);
    free(cd);
    free(pl);
  }
  free(rq);
  return CAIR_OK;
}

/* ---- CARS ranking path ------------------------------------------------------------------- */
/* apply_pooling 'attn' (multitask/cars.py:671-691): score_t = l3(tanh(l0 x_t)), t >= len -> -inf
 * (utils/misc.py:65-74), softmax over the padded length, weighted sum. */
static void attn_pool(const float* enc, int L, int H, int len, const cair_attn_mlp* a, float* out) {
  float* sc = (float*)malloc(sizeof(float) * (L + H));
  float* hid = sc + L;
  float mx = -INFINITY;
  for (int t = 0; t < L; ++t) {
    if (t >= len) {
      sc[t] = -INFINITY;
      continue;
    }
    for (int o = 0; o < H; ++o)
      hid[o] = tanhf(dotf(enc + (size_t)t * H, a->l0.w + (size_t)o * H, H) + a->l0.b[o]);
    sc[t] = dotf(hid, a->l3.w, H) + a->l3.b[0];
    if (sc[t] > mx) mx = sc[t];
  }
  float den = 0.0f;
  for (int t = 0; t < L; ++t) {
    sc[t] = (t < len) ? expf(sc[t] - mx) : 0.0f;
    den += sc[t];
  }
  for (int o = 0; o < H; ++o) {
    double s = 0.0;
    for (int t = 0; t < L; ++t) s += (double)enc[(size_t)t * H + o] * (double)(sc[t] / den);
    out[o] = (float)s;
  }
  free(sc);
}

/* embed + BiLSTM + attention pooling over n sequences (cars.py:193-225 / :227-260) */
static int cars_encode(const cair_cars_weights* w, const int64_t* ids, const int64_t* len, int n,
                       int L, int H, const cair_lstm_dir* fwd, const cair_lstm_dir* rev,
                       const cair_attn_mlp* attn, float* pooled, float* bank_out) {
  int E = w->emsize;
  float* x = (float*)malloc(sizeof(float) * (size_t)n * L * E);
  float* enc = (float*)malloc(sizeof(float) * (size_t)n * L * H);
  int rc = cair_oracle_embed(w->table, w->vocab, E, ids, (int64_t)n * L, x);
  if (rc == CAIR_OK) rc = cair_oracle_lstm(x, len, n, L, E, H / 2, fwd, rev, enc, NULL, NULL);
  if (rc == CAIR_OK) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int s = 0; s < n; ++s)
      attn_pool(enc + (size_t)s * L * H, L, H, (int)len[s], attn, pooled + (size_t)s * H);
  }
  if (rc == CAIR_OK && bank_out) memcpy(bank_out, enc, sizeof(float) * (size_t)n * L * H);
  free(x);
  free(enc);
  return rc;
}

/* softmax(scores[0..n)) then out = sum_k w_k * states[k] */
static void softmax_mix(float* sc, int n, const float* states, int stride, int dim, float* out) {
  float mx = -INFINITY, den = 0.0f;
  for (int k = 0; k < n; ++k)
    if (sc[k] > mx) mx = sc[k];
  for (int k = 0; k < n; ++k) {
    sc[k] = expf(sc[k] - mx);
    den += sc[k];
  }
  for (int o = 0; o < dim; ++o) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += (double)states[(size_t)k * stride + o] * (double)(sc[k] / den);
    out[o] = (float)s;
  }
}

/* Maxout (modules/maxout.py:70-84): affine to out*pool, view(out, pool), max over pool. */
static void maxout_layer(const float* x, int in, const cair_linear* l, int out, int pool, float* y) {
  for (int o = 0; o < out; ++o) {
    float best = -INFINITY;
    for (int p = 0; p < pool; ++p) {
      int r = o * pool + p;
      float v = dotf(x, l->w + (size_t)r * in, in) + l->b[r];
      if (v > best) best = v;
    }
    y[o] = best;
  }
}

ORA_API int cair_oracle_cars_ex(const cair_cars_weights* w, const int64_t* q, const int64_t* qlen,
                                const int64_t* d, const int64_t* dlen, const float* labels, int B,
                                int S, int N, int Lq, int Ld, float* scores, float* pooled_q_out,
                                float* pooled_d_out, float* clicks_out, float* sess_q_attn_out,
                                float* sess_d_attn_out, float* enc_q_out, float* sess_h_out, float* sess_c_out);

ORA_API int cair_oracle_cars(const cair_cars_weights* w, const int64_t* q, const int64_t* qlen,
                             const int64_t* d, const int64_t* dlen, const float* labels, int B,
                             int S, int N, int Lq, int Ld, float* scores, float* pooled_q_out,
                             float* pooled_d_out, float* clicks_out, float* sess_q_attn_out,
                             float* sess_d_attn_out) {
  return cair_oracle_cars_ex(w, q, qlen, d, dlen, labels, B, S, N, Lq, Ld, scores, pooled_q_out, pooled_d_out, clicks_out,
                             sess_q_attn_out, sess_d_attn_out, NULL, NULL, NULL);
}

/* Same with the decoder-side outputs: enc_q [B*S,Lq,Hq] query memory banks (cars.py:214-225), sess_h / sess_c
 * [B,S,Hsq+Hsd] the (h, c) of both session encoders after every query, query part first (cars.py:391-411). */
ORA_API int cair_oracle_cars_ex(const cair_cars_weights* w, const int64_t* q, const int64_t* qlen,
                                const int64_t* d, const int64_t* dlen, const float* labels, int B,
                                int S, int N, int Lq, int Ld, float* scores, float* pooled_q_out,
                                float* pooled_d_out, float* clicks_out, float* sess_q_attn_out,
                                float* sess_d_attn_out, float* enc_q_out, float* sess_h_out, float* sess_c_out) {
  int Hq = w->nhid_query, Hd = w->nhid_document, Hsq = w->nhid_session_query,
      Hsd = w->nhid_session_document;
  int BS = B * S;
  if (!check_ids(q, (int64_t)BS * Lq, w->vocab) || !check_ids(d, (int64_t)BS * N * Ld, w->vocab))
    return CAIR_ERR_BAD_ARG;
  float* pq = (float*)malloc(sizeof(float) * (size_t)BS * Hq);
  float* pd = (float*)malloc(sizeof(float) * (size_t)BS * N * Hd);
  float* clk = (float*)malloc(sizeof(float) * (size_t)BS * Hd);
  int rc = cars_encode(w, q, qlen, BS, Lq, Hq, &w->query_fwd, &w->query_rev, &w->q_attn, pq, enc_q_out);
  if (rc == CAIR_OK)
    rc = cars_encode(w, d, dlen, BS * N, Ld, Hd, &w->doc_fwd, &w->doc_rev, &w->d_attn, pd, NULL);
  if (rc != CAIR_OK) {
    free(pq);
    free(pd);
    free(clk);
    return rc;
  }
  /* encode_clicks (cars.py:262-304), including the batch-global mask width */
  int m = 0;
  for (int r = 0; r < BS; ++r) {
    int k = 0;
    for (int n = 0; n < N; ++n) k += labels[r * N + n] != 0.0f;
    if (k > m) m = k;
  }
  for (int r = 0; r < BS; ++r) {
    int* order = (int*)malloc(sizeof(int) * N);
    float* sc = (float*)malloc(sizeof(float) * (N + Hd) + sizeof(float) * (size_t)N * Hd);
    float* hid = sc + N;
    float* sorted = hid + Hd;
    int k = 0;
    for (int n = 0; n < N; ++n) {
      order[n] = n;
      k += labels[r * N + n] != 0.0f;
    }
    /* stable descending sort by label (torch CPU sort keeps ties in index order) */
    for (int a = 1; a < N; ++a) {
      int v = order[a], c = a - 1;
      while (c >= 0 && labels[r * N + order[c]] < labels[r * N + v]) {
        order[c + 1] = order[c];
        --c;
      }
      order[c + 1] = v;
    }
    for (int n = 0; n < N; ++n)
      memcpy(sorted + (size_t)n * Hd, pd + ((size_t)r * N + order[n]) * Hd, sizeof(float) * Hd);
    for (int n = 0; n < N; ++n) {
      int keep = (n < m) ? (n < k) : 1;
      if (!keep) {
        sc[n] = -INFINITY;
        continue;
      }
      for (int o = 0; o < Hd; ++o)
        hid[o] = tanhf(dotf(sorted + (size_t)n * Hd, w->click_attn.l0.w + (size_t)o * Hd, Hd) +
                       w->click_attn.l0.b[o]);
      sc[n] = dotf(hid, w->click_attn.l3.w, Hd) + w->click_attn.l3.b[0];
    }
    softmax_mix(sc, N, sorted, Hd, Hd, clk + (size_t)r * Hd);
    free(order);
    free(sc);
  }
  if (pooled_q_out) memcpy(pooled_q_out, pq, sizeof(float) * (size_t)BS * Hq);
  if (pooled_d_out) memcpy(pooled_d_out, pd, sizeof(float) * (size_t)BS * N * Hd);
  if (clicks_out) memcpy(clicks_out, clk, sizeof(float) * (size_t)BS * Hd);
  /* encode_session (cars.py:306-458) + rank (:460-520), per session b */
  const int rd0 = w->rank_dims[0], rd1 = w->rank_dims[1], rd2 = w->rank_dims[2], pool = w->rank_pool;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int Hs = Hsq + Hsd;
    float* Q = (float*)calloc((size_t)(S + 1) * Hsq, sizeof(float)); /* state 0 = zeros */
    float* D = (float*)calloc((size_t)(S + 1) * Hsd, sizeof(float));
    float* hq = (float*)calloc(2 * (size_t)Hsq + 4 * Hsq, sizeof(float));
    float* cqs = hq + Hsq;
    float* gq = cqs + Hsq;
    float* hdn = (float*)calloc(2 * (size_t)Hsd + 4 * Hsd, sizeof(float));
    float* cds = hdn + Hsd;
    float* gd = cds + Hsd;
    float* tmp = (float*)malloc(sizeof(float) * ((size_t)(S + 1) + Hq + Hd + Hs + Hd + 4 * Hd + rd0 + rd1 + rd2 + Hsq + Hsd));
    float* att = tmp;
    float* proj = att + (S + 1);
    float* sess = proj + (Hq > Hd ? Hq : Hd);
    float* qr = sess + Hs;
    float* feat = qr + Hd;
    float* y0 = feat + 4 * Hd;
    float* y1 = y0 + rd0;
    float* y2 = y1 + rd1;
    float* hidb = y2 + rd2;
    for (int s = 0; s < S; ++s) {
      const float* cur = pq + ((size_t)b * S + s) * Hq;
      int ns = s + 1;
      /* attention over past query-session states with the current query (:349-354) */
      for (int k = 0; k < ns; ++k) {
        for (int o = 0; o < Hq; ++o)
          proj[o] = dotf(Q + (size_t)k * Hsq, w->session_query_attn.w + (size_t)o * Hsq, Hsq) +
                    w->session_query_attn.b[o];
        att[k] = dotf(proj, cur, Hq);
      }
      softmax_mix(att, ns, Q, Hsq, Hsq, sess);
      /* same for doc-session states, still keyed by the QUERY (:359-364) */
      for (int k = 0; k < ns; ++k) {
        for (int o = 0; o < Hd; ++o)
          proj[o] = dotf(D + (size_t)k * Hsd, w->session_doc_attn.w + (size_t)o * Hsd, Hsd) +
                    w->session_doc_attn.b[o];
        att[k] = dotf(proj, cur, Hq);
      }
      softmax_mix(att, ns, D, Hsd, Hsd, sess + Hsq);
      /* rank (:460-520) */
      for (int o = 0; o < Hd; ++o)
        qr[o] = dotf(cur, w->q_projection.w + (size_t)o * Hq, Hq) + w->q_projection.b[o] +
                (dotf(sess, w->shared_session_projector.w + (size_t)o * Hs, Hs) +
                 dotf(sess, w->private_session_projector1.w + (size_t)o * Hs, Hs));
      for (int n = 0; n < N; ++n) {
        const float* dv = pd + (((size_t)b * S + s) * N + n) * Hd;
        for (int o = 0; o < Hd; ++o) {
          feat[o] = qr[o];
          feat[Hd + o] = dv[o];
          feat[2 * Hd + o] = fabsf(qr[o] - dv[o]);
          feat[3 * Hd + o] = qr[o] * dv[o];
        }
        maxout_layer(feat, 4 * Hd, &w->ranknet[0], rd0, pool, y0);
        maxout_layer(y0, rd0, &w->ranknet[1], rd1, pool, y1);
        maxout_layer(y1, rd1, &w->ranknet[2], rd2, pool, y2);
        scores[((size_t)b * S + s) * N + n] = y2[0];
      }
      /* step both session LSTMs (:378-383, :400-405): single-step, carried (h, c) */
      lstm_step(cur, Hq, Hsq, &w->session_query, hq, cqs, gq);
      memcpy(Q + (size_t)(s + 1) * Hsq, hq, sizeof(float) * Hsq);
      lstm_step(clk + ((size_t)b * S + s) * Hd, Hd, Hsd, &w->session_doc, hdn, cds, gd);
      memcpy(D + (size_t)(s + 1) * Hsd, hdn, sizeof(float) * Hsd);
      if (sess_h_out) {
        memcpy(sess_h_out + ((size_t)b * S + s) * Hs, hq, sizeof(float) * Hsq);
        memcpy(sess_h_out + ((size_t)b * S + s) * Hs + Hsq, hdn, sizeof(float) * Hsd);
      }
      if (sess_c_out) {
        memcpy(sess_c_out + ((size_t)b * S + s) * Hs, cqs, sizeof(float) * Hsq);
        memcpy(sess_c_out + ((size_t)b * S + s) * Hs + Hsq, cds, sizeof(float) * Hsd);
      }
      /* inner attention over states 1..s+1 (:385-389, :407-411) - decoder-side outputs */
      if (sess_q_attn_out) {
        for (int k = 0; k < ns; ++k) {
          const float* st = Q + (size_t)(k + 1) * Hsq;
          for (int o = 0; o < Hsq; ++o)
            hidb[o] = tanhf(dotf(st, w->session_query_inner_attn.l0.w + (size_t)o * Hsq, Hsq) +
                            w->session_query_inner_attn.l0.b[o]);
          att[k] = dotf(hidb, w->session_query_inner_attn.l3.w, Hsq) + w->session_query_inner_attn.l3.b[0];
        }
        softmax_mix(att, ns, Q + Hsq, Hsq, Hsq, sess_q_attn_out + ((size_t)b * S + s) * Hsq);
      }
      if (sess_d_attn_out) {
        for (int k = 0; k < ns; ++k) {
          const float* st = D + (size_t)(k + 1) * Hsd;
          for (int o = 0; o < Hsd; ++o)
            hidb[o] = tanhf(dotf(st, w->session_doc_inner_attn.l0.w + (size_t)o * Hsd, Hsd) +
                            w->session_doc_inner_attn.l0.b[o]);
          att[k] = dotf(hidb, w->session_doc_inner_attn.l3.w, Hsd) + w->session_doc_inner_attn.l3.b[0];
        }
        softmax_mix(att, ns, D + Hsd, Hsd, Hsd, sess_d_attn_out + ((size_t)b * S + s) * Hsd);
      }
    }
    free(Q);
    free(D);
    free(hq);
    free(hdn);
    free(tmp);
  }
  free(pq);
  free(pd);
  free(clk);
  return CAIR_OK;
}

/* CARS.decode (multitask/cars.py:706-791): greedy decode of max_len tokens for every (b, s < S-1) row.
 * decoders/rnn_decoder.py:19-90: one nn.LSTM step from the carried state (decoders/decoder.py:118-155 updates the
 * state object in place on every call), GlobalAttention 'general' (modules/global_attention.py:121-211: align = (W_in h) m,
 * masked beyond the memory length, softmax, context, tanh(W_out [context; h])), then token_prob_predictor1, the session
 * summary (shared_session_projector + private_session_projector2), token_prob_predictor2, softmax, arg-max (cars.py:761-775),
 * and the target-id -> source-id map for the next input (:780-783).  Row orderings as in the reference: initial states
 * row i = s*B + b (torch.cat(hidden_states[:-1], dim=1), :440-453); memory banks, lengths, session summaries and
 * predictions row i = b*(S-1) + s (:724-731, :747-757, :786). */
ORA_API int cair_oracle_cars_decode(const cair_cars_weights* w, const cair_cars_decoder_weights* dw, const float* enc_q,
                                    const int64_t* qlen, const float* sess_h, const float* sess_c, const float* sqa,
                                    const float* sda, int B, int S, int Lq, int max_len, const int64_t* tgt2src,
                                    int64_t bos, int64_t* predictions) {
  const int Hq = w->nhid_query, Hd = w->nhid_document, Hsq = w->nhid_session_query, Hsd = w->nhid_session_document;
  const int Hs = Hsq + Hsd, H = dw->nhid_decoder, Vt = dw->tgt_vocab, E = w->emsize;
  const int R = B * (S - 1);
  if (S < 2) return CAIR_OK;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < R; ++i) {
    const int s1 = i / B, b1 = i % B;            /* state rows */
    const int b2 = i / (S - 1), s2 = i % (S - 1); /* memory / summary rows */
    float* buf = (float*)malloc(sizeof(float) * ((size_t)2 * H + (size_t)Lq * H + Hs + Hd + 4 * H + H + Lq + 2 * H + H + Hd + Vt));
    float *hs = buf, *cs = hs + H, *mb = cs + H, *sum = mb + (size_t)Lq * H, *sess = sum + Hs, *gates = sess + Hd,
          *hq = gates + 4 * H, *al = hq + H, *cat = al + Lq, *ah = cat + 2 * H, *o1 = ah + H, *logit = o1 + Hd;
    const float* h0 = sess_h + ((size_t)b1 * S + s1) * Hs;
    const float* c0 = sess_c + ((size_t)b1 * S + s1) * Hs;
    for (int o = 0; o < H; ++o) {
      hs[o] = dotf(h0, dw->transform_hid.w + (size_t)o * Hs, Hs) + dw->transform_hid.b[o];
      cs[o] = dotf(c0, dw->transform_cell.w + (size_t)o * Hs, Hs) + dw->transform_cell.b[o];
    }
    const float* bank = enc_q + ((size_t)b2 * S + s2) * Lq * Hq;
    for (int t = 0; t < Lq; ++t)
      for (int o = 0; o < H; ++o) mb[(size_t)t * H + o] = dotf(bank + (size_t)t * Hq, dw->dec_attn.w + (size_t)o * Hq, Hq);
    int ml = (int)qlen[b2 * S + s2];
    for (int k = 0; k < Hs; ++k) sum[k] = k < Hsq ? sqa[((size_t)b2 * S + s2) * Hsq + k] : sda[((size_t)b2 * S + s2) * Hsd + (k - Hsq)];
    for (int o = 0; o < Hd; ++o)
      sess[o] = dotf(sum, w->shared_session_projector.w + (size_t)o * Hs, Hs) +
                dotf(sum, dw->private_session_projector2.w + (size_t)o * Hs, Hs);
    int64_t tok = bos;
    for (int t = 0; t < max_len; ++t) {
      lstm_step(w->table + (size_t)tok * E, E, H, &dw->rnn, hs, cs, gates);
      for (int o = 0; o < H; ++o) hq[o] = dotf(hs, dw->attn_in.w + (size_t)o * H, H);
      float mx = -INFINITY, den = 0.0f;
      for (int p = 0; p < ml; ++p) {
        al[p] = dotf(hq, mb + (size_t)p * H, H);
        if (al[p] > mx) mx = al[p];
      }
      for (int p = 0; p < ml; ++p) {
        al[p] = expf(al[p] - mx);
        den += al[p];
      }
      for (int o = 0; o < H; ++o) {
        double a = 0.0;
        for (int p = 0; p < ml; ++p) a += (double)(al[p] / den) * (double)mb[(size_t)p * H + o];
        cat[o] = (float)a;
        cat[H + o] = hs[o];
      }
      for (int o = 0; o < H; ++o) ah[o] = tanhf(dotf(cat, dw->attn_out.w + (size_t)o * 2 * H, 2 * H));
      for (int o = 0; o < Hd; ++o) o1[o] = dotf(ah, dw->predictor1.w + (size_t)o * H, H) + sess[o];
      int best = 0;
      for (int v = 0; v < Vt; ++v) {
        logit[v] = dotf(o1, dw->predictor2.w + (size_t)v * Hd, Hd);
        if (logit[v] > logit[best]) best = v;
      }
      predictions[(size_t)i * max_len + t] = best;
      tok = tgt2src[best];
    }
    free(buf);
  }
  return CAIR_OK;
}

/* Ranker.predict post-op (models/ranker.py:257-258): softmax over the N candidates. */
ORA_API int cair_oracle_softmax(const float* scores, int B, int N, float* out) {
  for (int b = 0; b < B; ++b) {
    float mx = -INFINITY, den = 0.0f;
    for (int n = 0; n < N; ++n)
      if (scores[b * N + n] > mx) mx = scores[b * N + n];
    for (int n = 0; n < N; ++n) {
      out[b * N + n] = expf(scores[b * N + n] - mx);
      den += out[b * N + n];
    }
    for (int n = 0; n < N; ++n) out[b * N + n] /= den;
  }
  return CAIR_OK;
}

/* ---- DSSM (rankers/dssm.py:33-63): max-pool over ALL positions (zero PAD rows take part), two-layer tanh MLP per
 * side, cosine ---------------------------------------------------------------------------------------------- */
static void mlp2_tanh(const float* x, int in, const cair_linear* l0, int hid, const cair_linear* l1, int out, float* y) {
  float* h = (float*)malloc(sizeof(float) * hid);
  for (int o = 0; o < hid; ++o) h[o] = tanhf(dotf(x, l0->w + (size_t)o * in, in) + l0->b[o]);
  for (int o = 0; o < out; ++o) y[o] = tanhf(dotf(h, l1->w + (size_t)o * hid, hid) + l1->b[o]);
  free(h);
}

ORA_API int cair_oracle_dssm(const cair_dssm_weights* w, const int64_t* q, const int64_t* qlen, const int64_t* d,
                             const int64_t* dlen, int B, int N, int Lq, int Ld, float* scores) {
  (void)qlen;
  (void)dlen;
  int E = w->emsize, H = w->nhid, O = w->nout;
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, (int64_t)B * N * Ld, w->vocab)) return CAIR_ERR_BAD_ARG;
#pragma omp parallel for
  for (int b = 0; b < B; ++b) {
    float* buf = (float*)malloc(sizeof(float) * (2 * (size_t)E + 2 * O));
    float *pq = buf, *pd = buf + E, *rq = pd + E, *rd = rq + O;
    for (int k = 0; k < E; ++k) {
      float m = -INFINITY;
      for (int t = 0; t < Lq; ++t) m = fmaxf(m, w->table[q[b * Lq + t] * E + k]);
      pq[k] = m;
    }
    mlp2_tanh(pq, E, &w->query_mlp0, H, &w->query_mlp2, O, rq);
    for (int n = 0; n < N; ++n) {
      const int64_t* dd = d + ((int64_t)b * N + n) * Ld;
      for (int k = 0; k < E; ++k) {
        float m = -INFINITY;
        for (int t = 0; t < Ld; ++t) m = fmaxf(m, w->table[dd[t] * E + k]);
        pd[k] = m;
      }
      mlp2_tanh(pd, E, &w->doc_mlp0, H, &w->doc_mlp2, O, rd);
      scores[b * N + n] = cosine_norm_first(rq, rd, O, 1e-8f);
    }
    free(buf);
  }
  return CAIR_OK;
}

/* ---- CDSSM (rankers/cdssm.py:32-77): window-3 interleave, Conv1d(3E -> L, k=3) (valid), tanh, Linear(L,O), tanh,
 * max over positions, cosine ----------------------------------------------------------------------------------- */
static void cdssm_side(const float* table, int E, const int64_t* ids, int L, const cair_linear* conv, int Lh,
                       const cair_linear* sem, int O, float* rep) {
  int T1 = L - 2, T2 = T1 - 2;  /* interleaved length, conv output length */
  float* hid = (float*)malloc(sizeof(float) * Lh);
  for (int o = 0; o < O; ++o) rep[o] = -INFINITY;
  for (int t = 0; t < T2; ++t) {
    for (int f = 0; f < Lh; ++f) {
      double s = conv->b[f];
      for (int k = 0; k < 3; ++k)       /* conv tap over the interleaved sequence */
        for (int wi = 0; wi < 3; ++wi) { /* position inside the window-3 interleave */
          const float* x = table + ids[t + k + wi] * E;
          const float* wr = conv->w + ((size_t)f * 3 * E + (size_t)wi * E) * 3 + k;
          for (int e = 0; e < E; ++e) s += (double)wr[(size_t)e * 3] * (double)x[e];
        }
      hid[f] = tanhf((float)s);
    }
    for (int o = 0; o < O; ++o) {
      float v = tanhf(dotf(hid, sem->w + (size_t)o * Lh, Lh) + sem->b[o]);
      if (v > rep[o]) rep[o] = v;
    }
  }
  free(hid);
}

ORA_API int cair_oracle_cdssm(const cair_cdssm_weights* w, const int64_t* q, const int64_t* qlen, const int64_t* d,
                              const int64_t* dlen, int B, int N, int Lq, int Ld, float* scores) {
  (void)qlen;
  (void)dlen;
  int E = w->emsize, H = w->nhid, O = w->nout;
  if (Lq < 5 || Ld < 5) return CAIR_ERR_BAD_SHAPE;  /* interleave needs >= 3 tokens, the k=3 conv 2 more */
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, (int64_t)B * N * Ld, w->vocab)) return CAIR_ERR_BAD_ARG;
#pragma omp parallel for
  for (int b = 0; b < B; ++b) {
    float* rq = (float*)malloc(sizeof(float) * 2 * O);
    float* rd = rq + O;
    cdssm_side(w->table, E, q + (size_t)b * Lq, Lq, &w->query_conv, H, &w->query_sem, O, rq);
    for (int n = 0; n < N; ++n) {
      cdssm_side(w->table, E, d + ((size_t)b * N + n) * Ld, Ld, &w->doc_conv, H, &w->doc_sem, O, rd);
      scores[b * N + n] = cosine_norm_first(rq, rd, O, 1e-8f);
    }
    free(rq);
  }
  return CAIR_OK;
}

/* ---- ARC-I (rankers/arci.py:60-105) -----------------------------------------------------------------------
 * per side: [Conv1d(same padding k/2) -> ReLU -> MaxPool1d(pool)] x nlayers over the embedded tokens, flatten(1) of the
 * [C, L'] maps, concat(query, doc), Linear(inp, inp/2) -> Linear(inp/2, 1) (no non-linearity between them). */
static int arc_conv_stack(const float* x0, int L, int Cin, int nl, const int32_t* filters, const int32_t* kernel,
                          const int32_t* pool, const cair_linear* convs, float** out, int* Lout) {
  float* cur = (float*)malloc(sizeof(float) * (size_t)L * Cin);  /* [L][C] */
  memcpy(cur, x0, sizeof(float) * (size_t)L * Cin);
  for (int l = 0; l < nl; ++l) {
    int F = filters[l], k = kernel[l], pad = k / 2, P = pool[l], Lp = L / P;
    float* y = (float*)malloc(sizeof(float) * (size_t)L * F);
    for (int t = 0; t < L; ++t)
      for (int f = 0; f < F; ++f) {
        double s = convs[l].b[f];
        for (int kk = 0; kk < k; ++kk) {
          int tt = t + kk - pad;
          if (tt < 0 || tt >= L) continue;
          for (int c = 0; c < Cin; ++c) s += (double)convs[l].w[((size_t)f * Cin + c) * k + kk] * (double)cur[(size_t)tt * Cin + c];
        }
        y[(size_t)t * F + f] = (float)s > 0.f ? (float)s : 0.f;
      }
    float* z = (float*)malloc(sizeof(float) * (size_t)(Lp > 0 ? Lp : 1) * F);
    for (int t = 0; t < Lp; ++t)
      for (int f = 0; f < F; ++f) {
        float m = -INFINITY;
        for (int u = 0; u < P; ++u) m = fmaxf(m, y[(size_t)(t * P + u) * F + f]);
        z[(size_t)t * F + f] = m;
      }
    free(cur);
    free(y);
    cur = z, L = Lp, Cin = F;
    if (L <= 0) {
      free(cur);
      return CAIR_ERR_BAD_SHAPE;
    }
  }
  *out = cur, *Lout = L;
  return CAIR_OK;
}

ORA_API int cair_oracle_arci(const cair_arci_weights* w, const int64_t* q, const int64_t* qlen, const int64_t* d,
                             const int64_t* dlen, int B, int N, int Lq, int Ld, float* scores) {
  (void)qlen;
  (void)dlen;
  if (Lq != w->max_query_len || Ld != w->max_doc_len) return CAIR_ERR_BAD_SHAPE; /* mlp in-features are baked in (:46-57) */
  int E = w->emsize, nl = w->nlayers, Fl = w->filters[nl - 1];
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, (int64_t)B * N * Ld, w->vocab)) return CAIR_ERR_BAD_ARG;
  int rc_all = CAIR_OK;
#pragma omp parallel for
  for (int b = 0; b < B; ++b) {
    float* xq = (float*)malloc(sizeof(float) * (size_t)Lq * E);
    cair_oracle_embed(w->table, w->vocab, E, q + (size_t)b * Lq, Lq, xq);
    float* fq = NULL;
    int Lqo = 0;
    int rc = arc_conv_stack(xq, Lq, E, nl, w->filters, w->kernel, w->pool, w->qconv, &fq, &Lqo);
    free(xq);
    if (rc != CAIR_OK) {
      rc_all = rc;
      continue;
    }
    for (int n = 0; n < N; ++n) {
      float* xd = (float*)malloc(sizeof(float) * (size_t)Ld * E);
      cair_oracle_embed(w->table, w->vocab, E, d + ((size_t)b * N + n) * Ld, Ld, xd);
      float* fd = NULL;
      int Ldo = 0;
      rc = arc_conv_stack(xd, Ld, E, nl, w->filters, w->kernel, w->pool, w->dconv, &fd, &Ldo);
      free(xd);
      if (rc != CAIR_OK) {
        rc_all = rc;
        continue;
      }
      int inp = Fl * (Lqo + Ldo), hid = inp / 2;
      /* flatten(1) of [C, L'] maps: index c*L' + l; query block first */
      float* com = (float*)malloc(sizeof(float) * (size_t)inp);
      for (int c = 0; c < Fl; ++c) {
        for (int l = 0; l < Lqo; ++l) com[c * Lqo + l] = fq[(size_t)l * Fl + c];
        for (int l = 0; l < Ldo; ++l) com[Fl * Lqo + c * Ldo + l] = fd[(size_t)l * Fl + c];
      }
      double sc = w->mlp1.b[0];
      for (int o = 0; o < hid; ++o)
        sc += (double)w->mlp1.w[o] * (double)(dotf(com, w->mlp0.w + (size_t)o * inp, inp) + w->mlp0.b[o]);
      scores[b * N + n] = (float)sc;
      free(com);
      free(fd);
    }
    free(fq);
  }
  return rc_all;
}

/* ---- ARC-II (rankers/arcii.py:58-111) --------------------------------------------------------------------
 * Conv1d(same) on both sides (no activation), comb[f,j,i] = cd[f,j] + cq[f,i] (a SUM, :99), MaxPool2d(2,2), then
 * [Conv2d(3x3, same) -> ReLU -> MaxPool2d(2,2)] x nlayers over the (doc, query) map, flatten(1), two Linear layers. */
ORA_API int cair_oracle_arcii(const cair_arcii_weights* w, const int64_t* q, const int64_t* qlen, const int64_t* d,
                              const int64_t* dlen, int B, int N, int Lq, int Ld, float* scores) {
  (void)qlen;
  (void)dlen;
  if (Lq != w->max_query_len || Ld != w->max_doc_len) return CAIR_ERR_BAD_SHAPE;
  int E = w->emsize, F1 = w->filters_1d, k1 = w->kernel_1d, pad1 = k1 / 2, nl = w->nlayers2d;
  if (!check_ids(q, (int64_t)B * Lq, w->vocab) || !check_ids(d, (int64_t)B * N * Ld, w->vocab)) return CAIR_ERR_BAD_ARG;
  int64_t BN = (int64_t)B * N;
  int rc_all = CAIR_OK;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t p = 0; p < BN; ++p) {
    int b = (int)(p / N);
    /* 1-D convs: cq [Lq][F1], cd [Ld][F1] */
    float* cq = (float*)malloc(sizeof(float) * (size_t)(Lq + Ld) * F1);
    float* cd = cq + (size_t)Lq * F1;
    for (int side = 0; side < 2; ++side) {
      const int64_t* ids = side ? d + p * Ld : q + (size_t)b * Lq;
      int L = side ? Ld : Lq;
      const cair_linear* cv = side ? &w->conv_doc : &w->conv_query;
      float* o = side ? cd : cq;
      for (int t = 0; t < L; ++t)
        for (int f = 0; f < F1; ++f) {
          double s = cv->b[f];
          for (int kk = 0; kk < k1; ++kk) {
            int tt = t + kk - pad1;
            if (tt < 0 || tt >= L) continue;
            const float* x = w->table + ids[tt] * E;
            for (int e = 0; e < E; ++e) s += (double)cv->w[((size_t)f * E + e) * k1 + kk] * (double)x[e];
          }
          o[(size_t)t * F1 + f] = (float)s;
        }
    }
    /* sum + first 2x2 max-pool: map [H][W][C], H = Ld/2 (docs), W = Lq/2 (query) */
    int H = Ld / 2, W = Lq / 2, C = F1;
    float* cur = (float*)malloc(sizeof(float) * (size_t)H * W * C);
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < C; ++c) {
          float m = -INFINITY;
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
              m = fmaxf(m, cd[(size_t)(2 * y + dy) * F1 + c] + cq[(size_t)(2 * x + dx) * F1 + c]);
          cur[((size_t)y * W + x) * C + c] = m;
        }
    free(cq);
    int bad = 0;
    for (int l = 0; l < nl && !bad; ++l) {
      int F = w->filters_2d[l];
      float* yv = (float*)malloc(sizeof(float) * (size_t)H * W * F);
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          for (int f = 0; f < F; ++f) {
            double s = w->conv2d[l].b[f];
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx) {
                int yy = y + ky - 1, xx = x + kx - 1;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                const float* src = cur + ((size_t)yy * W + xx) * C;
                const float* wr = w->conv2d[l].w + (size_t)f * C * 9 + ky * 3 + kx;
                for (int c = 0; c < C; ++c) s += (double)wr[(size_t)c * 9] * (double)src[c];
              }
            yv[((size_t)y * W + x) * F + f] = (float)s > 0.f ? (float)s : 0.f;
          }
      int H2 = H / 2, W2 = W / 2;
      if (H2 <= 0 || W2 <= 0) {
        bad = 1;
        free(yv);
        break;
      }
      float* z = (float*)malloc(sizeof(float) * (size_t)H2 * W2 * F);
      for (int y = 0; y < H2; ++y)
        for (int x = 0; x < W2; ++x)
          for (int f = 0; f < F; ++f) {
            float m = -INFINITY;
            for (int dy = 0; dy < 2; ++dy)
              for (int dx = 0; dx < 2; ++dx) m = fmaxf(m, yv[((size_t)(2 * y + dy) * W + 2 * x + dx) * F + f]);
            z[((size_t)y * W2 + x) * F + f] = m;
          }
      free(yv);
      free(cur);
      cur = z, H = H2, W = W2, C = F;
    }
    if (bad) {
      rc_all = CAIR_ERR_BAD_SHAPE;
      free(cur);
      continue;
    }
    int inp = C * H * W, hid = inp / 2;
    float* com = (float*)malloc(sizeof(float) * (size_t)inp); /* flatten(1) of [C, H, W] */
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) com[((size_t)c * H + y) * W + x] = cur[((size_t)y * W + x) * C + c];
    double sc = w->mlp1.b[0];
    for (int o = 0; o < hid; ++o)
      sc += (double)w->mlp1.w[o] * (double)(dotf(com, w->mlp0.w + (size_t)o * inp, inp) + w->mlp0.b[o]);
    scores[p] = (float)sc;
    free(com);
    free(cur);
  }
  return rc_all;
}

/* Number of OpenMP threads of the following calls (bench.py's CPU-baseline legs: torchrun exports OMP_NUM_THREADS=1,
 * and libgomp reads the environment only once at load time). */
#ifdef _OPENMP
#include <omp.h>
#endif
__attribute__((visibility("default"))) int cair_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}
