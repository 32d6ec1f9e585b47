// CARS ranking path (neuroir/multitask/cars.py): encode :193-225, encode_document :227-260,
// apply_pooling :671-691, encode_clicks :262-304, encode_session :306-458, rank :460-520,
// Maxout (neuroir/modules/maxout.py:70-84).
//
// The reference walks the session with a Python loop.  The session LSTM states depend only on
// the pooled queries / click vectors, never on the scores, so here
//   1. queries and documents are encoded (gathered pre-gate GEMM + persistent BiLSTM) and
//      attention-pooled;
//   2. click vectors are built for all (b,s) rows in parallel (stable label sort + the
//      batch-global mask width);
//   3. both session LSTMs run as ordinary length-S sequences over the pooled vectors;
//   4. session attention + rank head (Maxout) run for all (b,s) in parallel.
#include "models.cuh"

namespace cair {

static int32_t attn_copy(Owned& own, const cair_attn_mlp& a, int H, AttnPack* p, cudaStream_t s) {
  p->H = H;
  CAIR_TRY(dev_copy(own, a.l0.w, (size_t)H * H, &p->w0, s));
  CAIR_TRY(dev_copy(own, a.l0.b, (size_t)H, &p->b0, s));
  CAIR_TRY(dev_copy(own, a.l3.w, (size_t)H, &p->w3, s));
  CAIR_TRY(dev_copy(own, a.l3.b, 1, &p->b3, s));
  CAIR_TRY(gemm_tc_pack(own, p->w0, H, H, &p->w0_tc, s));
  return CAIR_OK;
}

__global__ void add_kernel(const float* a, const float* b, float* o, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

int32_t cars_create_state(Owned& own, const cair_cars_weights& w, CarsState* st, cudaStream_t s) {
  st->V = w.vocab, st->E = w.emsize, st->Hq = w.nhid_query, st->Hd = w.nhid_document;
  st->Hsq = w.nhid_session_query, st->Hsd = w.nhid_session_document;
  if (st->Hq != st->Hd) return fail(CAIR_ERR_BAD_SHAPE, "cars: nhid_query must equal nhid_document (cars.py:359-361 bmm)");
  if (st->Hq % 2) return fail(CAIR_ERR_BAD_SHAPE, "cars: bidirectional hidden size must be even");
  if (w.rank_pool != 2) return fail(CAIR_ERR_UNSUPPORTED, "cars: maxout pool size must be 2");
  for (int i = 0; i < 3; ++i) st->rd[i] = w.rank_dims[i];
  if (st->rd[2] != 1) return fail(CAIR_ERR_UNSUPPORTED, "cars: last maxout layer must have 1 output");
  st->pool = w.rank_pool;
  const int Hs = st->Hsq + st->Hsd;
  CAIR_TRY(dev_copy(own, w.table, (size_t)w.vocab * w.emsize, &st->table, s));
  CAIR_TRY(lstm_pack(own, &w.query_fwd, &w.query_rev, w.emsize, st->Hq / 2, &st->enc_q, s));
  CAIR_TRY(lstm_pack(own, &w.doc_fwd, &w.doc_rev, w.emsize, st->Hd / 2, &st->enc_d, s));
  if (rnn_tc_supported(w.emsize, st->Hq / 2)) CAIR_TRY(rnn_tc_pack(own, &w.query_fwd, &w.query_rev, w.emsize, st->Hq / 2, CAIR_RNN_LSTM, &st->rt_q, s));
  if (rnn_tc_supported(w.emsize, st->Hd / 2)) CAIR_TRY(rnn_tc_pack(own, &w.doc_fwd, &w.doc_rev, w.emsize, st->Hd / 2, CAIR_RNN_LSTM, &st->rt_d, s));
  CAIR_TRY(lstm_pack(own, &w.session_query, nullptr, st->Hq, st->Hsq, &st->sess_q, s));
  CAIR_TRY(lstm_pack(own, &w.session_doc, nullptr, st->Hd, st->Hsd, &st->sess_d, s));
  CAIR_TRY(attn_copy(own, w.q_attn, st->Hq, &st->q_attn, s));
  CAIR_TRY(attn_copy(own, w.d_attn, st->Hd, &st->d_attn, s));
  CAIR_TRY(attn_copy(own, w.click_attn, st->Hd, &st->click_attn, s));
  CAIR_TRY(attn_copy(own, w.session_query_inner_attn, st->Hsq, &st->sq_inner, s));
  CAIR_TRY(attn_copy(own, w.session_doc_inner_attn, st->Hsd, &st->sd_inner, s));
  CAIR_TRY(dev_copy(own, w.session_query_attn.w, (size_t)st->Hq * st->Hsq, &st->sqa_w, s));
  CAIR_TRY(dev_copy(own, w.session_query_attn.b, (size_t)st->Hq, &st->sqa_b, s));
  CAIR_TRY(dev_copy(own, w.session_doc_attn.w, (size_t)st->Hd * st->Hsd, &st->sda_w, s));
  CAIR_TRY(dev_copy(own, w.session_doc_attn.b, (size_t)st->Hd, &st->sda_b, s));
  CAIR_TRY(dev_copy(own, w.q_projection.w, (size_t)st->Hd * st->Hq, &st->qp_w, s));
  CAIR_TRY(dev_copy(own, w.q_projection.b, (size_t)st->Hd, &st->qp_b, s));
  if (!w.shared_session_projector.w || !w.private_session_projector1.w) return fail(CAIR_ERR_BAD_ARG, "cars: null projector");
  // both bias-free projectors see the same input (cars.py:506-512): W_shared x + W_priv x = (W_shared + W_priv) x
  const int64_t np = (int64_t)st->Hd * Hs;
  CAIR_CUDA(own.alloc(&st->sess_proj, (size_t)np));
  CAIR_TRY(dev_copy(own, w.shared_session_projector.w, (size_t)np, &st->shared_proj, s));
  CAIR_LAUNCH(add_kernel, (unsigned)((np + 255) / 256), 256, 0, s, w.shared_session_projector.w,
              w.private_session_projector1.w, st->sess_proj, np);
  int in = 4 * st->Hd;
  for (int i = 0; i < 3; ++i) {
    CAIR_TRY(dev_copy(own, w.ranknet[i].w, (size_t)st->rd[i] * 2 * in, &st->rk_w[i], s));
    CAIR_TRY(dev_copy(own, w.ranknet[i].b, (size_t)st->rd[i] * 2, &st->rk_b[i], s));
    if (i < 2 && in % 4 == 0) CAIR_TRY(gemm_tc_pack(own, st->rk_w[i], st->rd[i] * 2, in, &st->rk_tc[i], s));
    in = st->rd[i];
  }
  CAIR_CUDA(cudaStreamCreateWithFlags(&st->side, cudaStreamNonBlocking));
  CAIR_CUDA(cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming));
  CAIR_CUDA(cudaEventCreateWithFlags(&st->ev_join, cudaEventDisableTiming));
  return CAIR_OK;
}

// ---- attention pooling (cars.py:671-691): one CTA per sequence ----
// hid [n*L, H] = tanh(l0(enc)) precomputed by the GEMM; score_t = w3 . hid_t + b3, masked softmax, mix.
// scores (optional): w3 . hid_t + b3 already computed by the GEMM's row-dot epilogue (hid is then not materialised)
__global__ void __launch_bounds__(256) attn_pool_kernel(const float* __restrict__ enc, const float* __restrict__ hid,
                                                        const int64_t* __restrict__ len, int L, int H,
                                                        const float* __restrict__ w3, const float* __restrict__ b3,
                                                        float* __restrict__ pooled, const float* __restrict__ scores) {
  extern __shared__ float sc[];  // [L]
  __shared__ float red[8];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int l = (int)len[s];
  l = l < 1 ? 1 : (l > L ? L : l);
  for (int t = warp; t < L; t += 8) {
    float v = -INFINITY;
    if (t < l && scores) {
      v = scores[(size_t)s * L + t];
    } else if (t < l) {
      const float* hrow = hid + ((size_t)s * L + t) * H;
      float a = 0.f;
      for (int k = lane; k < H; k += 32) a = fmaf(w3[k], hrow[k], a);
      v = warp_sum(a) + b3[0];
    }
    if (lane == 0) sc[t] = v;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = tid; t < L; t += 256) m = fmaxf(m, sc[t]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float e = 0.f;
  for (int t = tid; t < L; t += 256) {
    float v = (t < l) ? __expf(sc[t] - m) : 0.f;
    sc[t] = v;
    e += v;
  }
  e = warp_sum(e);
  if (lane == 0) red[warp] = e;
  __syncthreads();
  float den = 0.f;
  for (int i = 0; i < 8; ++i) den += red[i];
  const float inv = 1.0f / den;
  for (int o = tid; o < H; o += 256) {
    float a = 0.f;
    for (int t = 0; t < l; ++t) a = fmaf(enc[((size_t)s * L + t) * H + o], sc[t] * inv, a);
    pooled[(size_t)s * H + o] = a;
  }
}

// row dot: out[r] = w . x[r, :] + b ; one warp per row
__global__ void rowdot_kernel(const float* __restrict__ x, int64_t rows, int H, const float* __restrict__ w,
                              const float* __restrict__ b, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float a = 0.f;
  for (int k = lane; k < H; k += 32) a = fmaf(w[k], x[r * H + k], a);
  a = warp_sum(a);
  if (lane == 0) out[r] = a + b[0];
}

// batch-global click-mask width m = max_r #nonzero(labels[r,:]) over ALL B*S rows (cars.py:285-292)
__global__ void click_width_kernel(const float* __restrict__ labels, int rows, int N, int* __restrict__ m) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int k = 0;
  for (int n = 0; n < N; ++n) k += labels[(size_t)r * N + n] != 0.0f;
  atomicMax(m, k);
}

// encode_clicks (cars.py:262-304): one CTA (128 threads) per (b,s) row of this rank's slice.
__global__ void __launch_bounds__(128) clicks_kernel(const float* __restrict__ pd, const float* __restrict__ att,
                                                     const float* __restrict__ labels, int N, int Hd,
                                                     const int* __restrict__ mwidth, int64_t row_begin,
                                                     float* __restrict__ clicks) {
  extern __shared__ float smc[];  // w[N] | order[N] (int)
  float* w = smc;
  int* order = reinterpret_cast<int*>(smc + N);
  const int64_t rl = blockIdx.x, rg = row_begin + rl;  // local / global row
  const float* lab = labels + rg * N;
  if (threadIdx.x == 0) {
    const int m = *mwidth;
    int k = 0;
    for (int n = 0; n < N; ++n) {
      order[n] = n;
      k += lab[n] != 0.0f;
    }
    // stable descending sort by label: ties keep index order (torch CPU sort behaviour, App. B4)
    for (int a = 1; a < N; ++a) {
      int v = order[a], c = a - 1;
      while (c >= 0 && lab[order[c]] < lab[v]) {
        order[c + 1] = order[c];
        --c;
      }
      order[c + 1] = v;
    }
    float mx = -INFINITY;
    for (int n = 0; n < N; ++n) {
      bool keep = (n < m) ? (n < k) : true;
      w[n] = keep ? att[rl * N + order[n]] : -INFINITY;
      mx = fmaxf(mx, w[n]);
    }
    float den = 0.f;
    for (int n = 0; n < N; ++n) {
      w[n] = (w[n] == -INFINITY) ? 0.f : __expf(w[n] - mx);
      den += w[n];
    }
    for (int n = 0; n < N; ++n) w[n] /= den;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < Hd; o += blockDim.x) {
    float a = 0.f;
    for (int n = 0; n < N; ++n) a = fmaf(pd[(rl * N + order[n]) * Hd + o], w[n], a);
    clicks[rl * Hd + o] = a;
  }
}

// block-cooperative y[o] = W[o,:] . x (+ b[o]) for o < out; W row-major [out, in]; one warp per row
__device__ void gemv_rows(const float* __restrict__ W, const float* __restrict__ b, const float* x, int in, int out,
                          float* y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = warp; o < out; o += nw) {
    const float* wr = W + (size_t)o * in;
    float a = 0.f;
    for (int k = lane; k < in; k += 32) a = fmaf(wr[k], x[k], a);
    a = warp_sum(a);
    if (lane == 0) y[o] = a + (b ? b[o] : 0.f);
  }
}

// softmax over att[0..ns) in place (ns <= 32), executed by warp 0
__device__ void softmax_small(float* att, int ns) {
  if (threadIdx.x < 32) {
    float v = threadIdx.x < ns ? att[threadIdx.x] : -INFINITY;
    float m = warp_max(v);
    float e = threadIdx.x < ns ? __expf(v - m) : 0.f;
    float den = warp_sum(e);
    if (threadIdx.x < ns) att[threadIdx.x] = e / den;
  }
}

// Session attention + rank head for one (b,s) (cars.py:346-373, :460-520).
// Qs/Ds: session LSTM outputs [nb, S, Hs*]; state k of the reference's list is zeros for k=0 and
// row k-1 for k>=1.  smem: cur[Hq] u[Hmax] att[32] sess[Hs] qr[Hd] feat[N*4Hd] y0[N*rd0] y1[N*rd1]
__global__ void __launch_bounds__(256) cars_rank_kernel(CarsState st, const float* __restrict__ pq,
                                                        const float* __restrict__ pd, const float* __restrict__ Qs,
                                                        const float* __restrict__ Ds, int S, int N,
                                                        int64_t row_begin, float* __restrict__ scores,
                                                        float* __restrict__ feat_out) {
  extern __shared__ float smr[];
  const int Hq = st.Hq, Hd = st.Hd, Hsq = st.Hsq, Hsd = st.Hsd, Hs = Hsq + Hsd;
  const int Hmax = max(max(Hsq, Hsd), Hd);  // u doubles as the session-projection scratch
  float* cur = smr;
  float* u = cur + Hq;
  float* att = u + Hmax;
  float* sess = att + 32;
  float* qr = sess + Hs;
  float* feat = qr + Hd;
  float* y0 = feat + (size_t)N * 4 * Hd;
  float* y1 = y0 + (size_t)N * st.rd[0];
  __shared__ float bdot_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t rl = blockIdx.x;           // local (b,s) row
  const int s = (int)((row_begin + rl) % S);
  const int64_t bl = rl / S;               // local session (slices are whole sessions)
  const int ns = s + 1;
  for (int k = tid; k < Hq; k += 256) cur[k] = pq[rl * Hq + k];
  __syncthreads();
  for (int side = 0; side < 2; ++side) {
    const float* W = side ? st.sda_w : st.sqa_w;   // [Hq, Hs*]
    const float* bb = side ? st.sda_b : st.sqa_b;
    const float* states = side ? Ds : Qs;
    const int Hx = side ? Hsd : Hsq;
    // att_k = (W state_k + b) . cur  =  state_k . (W^T cur) + b . cur
    for (int k = tid; k < Hx; k += 256) {
      float a = 0.f;
      for (int o = 0; o < Hq; ++o) a = fmaf(W[(size_t)o * Hx + k], cur[o], a);
      u[k] = a;
    }
    if (warp == 0) {
      float a = 0.f;
      for (int o = lane; o < Hq; o += 32) a = fmaf(bb[o], cur[o], a);
      a = warp_sum(a);
      if (lane == 0) bdot_s = a;
    }
    __syncthreads();
    for (int k = warp; k < ns; k += 8) {
      float a = 0.f;
      if (k > 0) {
        const float* st_k = states + ((size_t)bl * S + (k - 1)) * Hx;
        for (int e = lane; e < Hx; e += 32) a = fmaf(st_k[e], u[e], a);
        a = warp_sum(a);
      }
      if (lane == 0) att[k] = a + bdot_s;
    }
    __syncthreads();
    softmax_small(att, ns);
    __syncthreads();
    for (int e = tid; e < Hx; e += 256) {
      float a = 0.f;
      for (int k = 1; k < ns; ++k) a = fmaf(states[((size_t)bl * S + (k - 1)) * Hx + e], att[k], a);
      sess[(side ? Hsq : 0) + e] = a;
    }
    __syncthreads();
  }
  // q' = q_projection(cur) + (W_shared + W_priv1) [sq; sd]
  gemv_rows(st.qp_w, st.qp_b, cur, Hq, Hd, qr);
  __syncthreads();
  gemv_rows(st.sess_proj, nullptr, sess, Hs, Hd, u);
  __syncthreads();
  for (int o = tid; o < Hd; o += 256) qr[o] += u[o];
  __syncthreads();
  for (int i = tid; i < N * Hd; i += 256) {
    int n = i / Hd, o = i - n * Hd;
    float dv = pd[(rl * N + n) * Hd + o], qv = qr[o];
    float* f = feat + (size_t)n * 4 * Hd;
    f[o] = qv, f[Hd + o] = dv, f[2 * Hd + o] = fabsf(qv - dv), f[3 * Hd + o] = qv * dv;
  }
  __syncthreads();
  if (feat_out) {   // the Maxout layers run as GEMMs over all B*S*N candidate rows (cars_forward)
    float* fo = feat_out + (size_t)rl * N * 4 * Hd;
    for (int i = tid; i < N * 4 * Hd; i += 256) fo[i] = feat[i];
    return;
  }
  // Maxout layers: out[o] = max(row 2o, row 2o+1); each warp owns output o for all N docs
  const float* xin = feat;
  float* yout = y0;
  int in = 4 * Hd;
  for (int layer = 0; layer < 3; ++layer) {
    const int out = st.rd[layer];
    const float* W = st.rk_w[layer];
    const float* bb = st.rk_b[layer];
    for (int o = warp; o < out; o += 8) {
      for (int n0 = 0; n0 < N; n0 += 8) {
        float a0[8], a1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a0[j] = 0.f, a1[j] = 0.f;
        const float* w0 = W + (size_t)(2 * o) * in;
        const float* w1 = w0 + in;
        for (int k = lane; k < in; k += 32) {
          float wa = w0[k], wb = w1[k];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (n0 + j < N) {
              float xv = xin[(size_t)(n0 + j) * in + k];
              a0[j] = fmaf(wa, xv, a0[j]);
              a1[j] = fmaf(wb, xv, a1[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (n0 + j < N) {
            float v0 = warp_sum(a0[j]) + bb[2 * o], v1 = warp_sum(a1[j]) + bb[2 * o + 1];
            if (lane == 0) {
              float v = fmaxf(v0, v1);
              if (layer == 2)
                scores[(row_begin + rl) * N + n0 + j] = v;
              else
                yout[(size_t)(n0 + j) * out + o] = v;
            }
          }
        }
      }
    }
    __syncthreads();
    xin = yout;
    in = out;
    yout = y1;
  }
}

// Maxout pooling between the rank-net GEMMs (modules/maxout.py:70-84): x[r, o] = max(y[r, 2o], y[r, 2o+1])
__global__ void maxout_pool_kernel(const float* __restrict__ y, int64_t rows, int out, float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * out) return;
  const float2 v = *reinterpret_cast<const float2*>(y + 2 * i);
  x[i] = fmaxf(v.x, v.y);
}
// last Maxout layer (1 output, pool 2): score[r] = max(w0 . x + b0, w1 . x + b1); one warp per row
__global__ void maxout_final_kernel(const float* __restrict__ x, int64_t rows, int in, const float* __restrict__ w,
                                    const float* __restrict__ b, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float a0 = 0.f, a1 = 0.f;
  for (int k = lane; k < in; k += 32) {
    const float xv = x[r * in + k];
    a0 = fmaf(w[k], xv, a0);
    a1 = fmaf(w[in + k], xv, a1);
  }
  a0 = warp_sum(a0), a1 = warp_sum(a1);
  if (lane == 0) scores[r] = fmaxf(a0 + b[0], a1 + b[1]);
}

// inner attention over the session states 1..s+1 (cars.py:385-389, :407-411): one CTA per (b,s)
// sc [nb*S] = l3 . tanh(l0 state) + b precomputed per state; out[b,s] = sum_k softmax_k(sc[b,0..s]) state_k
__global__ void inner_attn_kernel(const float* __restrict__ states, const float* __restrict__ sc, int S, int H,
                                  float* __restrict__ out) {
  __shared__ float att[32];
  const int64_t rl = blockIdx.x;
  const int s = (int)(rl % S);
  const int64_t bl = rl / S;
  const int ns = s + 1;
  if (threadIdx.x < ns) att[threadIdx.x] = sc[bl * S + threadIdx.x];
  __syncthreads();
  softmax_small(att, ns);
  __syncthreads();
  for (int e = threadIdx.x; e < H; e += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < ns; ++k) a = fmaf(states[((size_t)bl * S + k) * H + e], att[k], a);
    out[rl * H + e] = a;
  }
}

__global__ void fill_len_kernel(int64_t* len, int n, int64_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len[i] = v;
}

static int32_t encode_pool(const CarsState& st, const LstmPack& lp, const RnnTcPack& rt, const AttnPack& ap, const int64_t* ids,
                           const int64_t* len, int64_t n, int L, float* pre, size_t pre_floats, float* enc, float* hid,
                           float* pooled, int* err, cudaStream_t s, const char* rec_name, bool marks = true) {
  const int H = ap.H;
  if (g_rnn_impl >= RNN_IMPL_AUTO && rt.wimg)   // tcgen05 recurrence (pre-gates from the gathered tcgen05 GEMM)
    CAIR_TRY(rnn_tc_run(rt, gemm_gather(st.table, st.V, st.E, ids, 1, 1, 1, err), len, (int)n, L, enc, nullptr, nullptr, pre,
                        err, s, rec_name));
  else
    CAIR_TRY(lstm_run(lp, gemm_gather(st.table, st.V, st.E, ids, 1, 1, 1, err), len, (int)n, L, enc, nullptr, nullptr,
                      pre, err, s, rec_name));
  if (marks) prof_mark("attention_pool", s);
  if (gemm_tc_rowdot_usable(gemm_dense(enc, H), ap.w0_tc, n * L)) {
    // tanh(l0(enc)) . w3 + b3 inside the GEMM epilogue: the [n*L, H] hidden tensor is never written (hid holds the scores)
    // the pre-gate workspace is free once the recurrence has run: it holds the A operand image of this GEMM (the memory bank
    // split into bf16 hi / lo once by gemm_tc_aimg_kernel, then streamed by bulk copies instead of the register loader)
    uint8_t* aimg = reinterpret_cast<uint8_t*>(((uintptr_t)pre + 127) & ~(uintptr_t)127);
    if (!pre || pre_floats * sizeof(float) < gemm_tc_aimg_bytes(n * L, H) + 128) aimg = nullptr;
    CAIR_TRY(gemm_tc_rowdot(gemm_dense(enc, H), ap.w0_tc, ap.b0, ACT_TANH, ap.w3, ap.b3, hid, n * L, s, aimg));
    CAIR_LAUNCH(attn_pool_kernel, (unsigned)n, 256, (size_t)L * sizeof(float), s, enc, nullptr, len, L, H, ap.w3, ap.b3, pooled, hid);
    return CAIR_OK;
  }
  CAIR_TRY(gemm_auto(gemm_dense(enc, H), ap.w0, ap.w0_tc, ap.b0, hid, H, n * L, H, H, ACT_TANH, s));
  CAIR_LAUNCH(attn_pool_kernel, (unsigned)n, 256, (size_t)L * sizeof(float), s, enc, hid, len, L, H, ap.w3, ap.b3,
              pooled, (const float*)nullptr);
  return CAIR_OK;
}

int32_t cars_forward(const CarsState& st, const CarsIO& io, int B, int S, int N, int Lq, int Ld, int sb, int sc,
                     Arena& ws, int* err, cudaStream_t s, bool dry) {
  if (S > 31) return fail(CAIR_ERR_UNSUPPORTED, "cars: session length %d > 31", S);
  const int Hq = st.Hq, Hd = st.Hd, Hsq = st.Hsq, Hsd = st.Hsd;
  const int64_t nrows = (int64_t)sc * S, ndocs = nrows * N;
  const int64_t r0 = (int64_t)sb * S;
  // workspace
  // pre-gate workspaces: the larger of what the two recurrence engines ask for (the tcgen05 path adds the A operand image of its
  // pre-gate GEMM, rnn_tc_workspace_floats)
  auto pre_floats = [](const LstmPack& lp, const RnnTcPack& rt, int64_t n, int L) {
    const size_t a = lstm_workspace_floats(lp, n, L), b = rt.wimg ? rnn_tc_workspace_floats(rt, n, L) : 0;
    return a > b ? a : b;
  };
  const size_t pre_q_floats = pre_floats(st.enc_q, st.rt_q, nrows, Lq), pre_d_floats = pre_floats(st.enc_d, st.rt_d, ndocs, Ld);
  float* pre_q = ws.take<float>(pre_q_floats);
  float* enc_q = ws.take<float>((size_t)nrows * Lq * Hq);
  float* hid_q = ws.take<float>((size_t)nrows * Lq * Hq);
  float* pq = ws.take<float>((size_t)nrows * Hq);
  float* pre_d = ws.take<float>(pre_d_floats);
  float* enc_d = ws.take<float>((size_t)ndocs * Ld * Hd);
  float* hid_d = ws.take<float>((size_t)ndocs * Ld * Hd);
  float* pd = ws.take<float>((size_t)ndocs * Hd);
  float* hid_c = ws.take<float>((size_t)ndocs * Hd);
  float* att_c = ws.take<float>((size_t)ndocs);
  float* clk = ws.take<float>((size_t)nrows * Hd);
  int* mwidth = ws.take<int>(1);
  int64_t* slen = ws.take<int64_t>((size_t)sc);
  float* pre_sq = ws.take<float>(lstm_workspace_floats(st.sess_q, sc, S));
  float* Qs = ws.take<float>((size_t)nrows * Hsq);
  float* Qc = ws.take<float>((size_t)nrows * Hsq);   // cell states per step (decoder-side output; reserved unconditionally so that
  float* Dc = ws.take<float>((size_t)nrows * Hsd);   // the workspace size does not depend on the requested outputs)
  float* pre_sd = ws.take<float>(lstm_workspace_floats(st.sess_d, sc, S));
  float* Ds = ws.take<float>((size_t)nrows * Hsd);
  const int Hsmax = Hsq > Hsd ? Hsq : Hsd;
  float* hid_s = ws.take<float>((size_t)nrows * Hsmax);
  float* sc_s = ws.take<float>((size_t)nrows);
  // rank head as GEMMs over the candidate rows
  float* rk_feat = ws.take<float>((size_t)ndocs * 4 * Hd);
  float* rk_y = ws.take<float>((size_t)ndocs * 2 * (st.rd[0] > st.rd[1] ? st.rd[0] : st.rd[1]));
  float* rk_x0 = ws.take<float>((size_t)ndocs * st.rd[0]);
  float* rk_x1 = ws.take<float>((size_t)ndocs * st.rd[1]);
  if (dry || sc <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "cars: workspace too small");

  // 1. encode + attention pooling.  The query chain (encoder, pooling, query-session LSTM) only needs the queries: it is
  // forked onto the handle's side stream and runs under the document chain; joined before the rank head.
  CAIR_LAUNCH(fill_len_kernel, (sc + 255) / 256, 256, 0, s, slen, sc, (int64_t)S);
  cudaStream_t sq = st.side ? st.side : s;
  if (st.side) {
    CAIR_CUDA(cudaEventRecord(st.ev_fork, s));
    CAIR_CUDA(cudaStreamWaitEvent(st.side, st.ev_fork, 0));
  } else {
    prof_mark("query_pregates", s);
  }
  CAIR_TRY(encode_pool(st, st.enc_q, st.rt_q, st.q_attn, io.q + r0 * Lq, io.qlen + r0, nrows, Lq, pre_q, pre_q_floats, enc_q, hid_q, pq, err, sq,
                       st.side ? nullptr : "query_recurrence", !st.side));
  // 3a. session LSTM over the pooled queries (zero initial state, S steps)
  CAIR_TRY(lstm_run(st.sess_q, gemm_dense(pq, Hq), slen, sc, S, Qs, nullptr, nullptr, pre_sq, err, sq, st.side ? nullptr : "lstm_recurrence",
                    io.sess_c ? Qc : nullptr));
  if (st.side) CAIR_CUDA(cudaEventRecord(st.ev_join, st.side));
  prof_mark("doc_pregates", s);
  CAIR_TRY(encode_pool(st, st.enc_d, st.rt_d, st.d_attn, io.d + r0 * N * Ld, io.dlen + r0 * N, ndocs, Ld, pre_d, pre_d_floats, enc_d, hid_d, pd, err, s, "doc_recurrence"));
  // 2. click vectors
  prof_mark("clicks", s);
  CAIR_CUDA(cudaMemsetAsync(mwidth, 0, sizeof(int), s));
  CAIR_LAUNCH(click_width_kernel, (B * S + 255) / 256, 256, 0, s, io.labels, B * S, N, mwidth);
  CAIR_TRY(gemm_auto(gemm_dense(pd, Hd), st.click_attn.w0, st.click_attn.w0_tc, st.click_attn.b0, hid_c, Hd, ndocs, Hd, Hd,
                     ACT_TANH, s));
  CAIR_LAUNCH(rowdot_kernel, (unsigned)((ndocs + 7) / 8), 256, 0, s, hid_c, ndocs, Hd, st.click_attn.w3, st.click_attn.b3, att_c);
  CAIR_LAUNCH(clicks_kernel, (unsigned)nrows, 128, (size_t)N * 8, s, pd, att_c, io.labels, N, Hd, mwidth, r0, clk);
  // 3b. session LSTM over the click vectors
  prof_mark("session_encoders", s);
  CAIR_TRY(lstm_run(st.sess_d, gemm_dense(clk, Hd), slen, sc, S, Ds, nullptr, nullptr, pre_sd, err, s, "lstm_recurrence",
                    io.sess_c ? Dc : nullptr));
  if (st.side) {
    prof_mark("join_query_side", s);
    CAIR_CUDA(cudaStreamWaitEvent(s, st.ev_join, 0));
  }
  // 4. session attention + rank head
  prof_mark("rank_head", s);
  {
    const int Hs = Hsq + Hsd;
    const int Hu = Hsmax > Hd ? Hsmax : Hd;
    size_t smem = ((size_t)Hq + Hu + 32 + Hs + Hd + (size_t)N * 4 * Hd + (size_t)N * st.rd[0] + (size_t)N * st.rd[1]) * sizeof(float);
    if (smem > 220 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "cars: N=%d x Hd=%d does not fit the rank kernel", N, Hd);
    CAIR_CUDA(cudaFuncSetAttribute(cars_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CarsState stc = st;
    const bool as_gemm = st.rk_tc[0].img && st.rk_tc[1].img && ndocs >= 128;
    CAIR_LAUNCH(cars_rank_kernel, (unsigned)nrows, 256, smem, s, stc, pq, pd, Qs, Ds, S, N, r0, io.scores,
                as_gemm ? rk_feat : (float*)nullptr);
    if (as_gemm) {
      // Maxout(4Hd -> rd0 -> rd1 -> 1, pool 2): two tensor-core GEMMs over all B*S*N rows instead of every (b,s) CTA
      // streaming the 2 MB of layer-0 weights from L2
      const int o0 = st.rd[0], o1 = st.rd[1];
      CAIR_TRY(gemm_auto(gemm_dense(rk_feat, 4 * Hd), st.rk_w[0], st.rk_tc[0], st.rk_b[0], rk_y, 2 * o0, ndocs, 2 * o0, 4 * Hd, ACT_NONE, s));
      CAIR_LAUNCH(maxout_pool_kernel, (unsigned)((ndocs * o0 + 255) / 256), 256, 0, s, rk_y, ndocs, o0, rk_x0);
      CAIR_TRY(gemm_auto(gemm_dense(rk_x0, o0), st.rk_w[1], st.rk_tc[1], st.rk_b[1], rk_y, 2 * o1, ndocs, 2 * o1, o0, ACT_NONE, s));
      CAIR_LAUNCH(maxout_pool_kernel, (unsigned)((ndocs * o1 + 255) / 256), 256, 0, s, rk_y, ndocs, o1, rk_x1);
      CAIR_LAUNCH(maxout_final_kernel, (unsigned)((ndocs + 7) / 8), 256, 0, s, rk_x1, ndocs, o1, st.rk_w[2], st.rk_b[2], io.scores + r0 * N);
    }
  }
  // decoder-side outputs: query memory banks, session-encoder states after every query
  if (io.enc_q)
    CAIR_CUDA(cudaMemcpyAsync(io.enc_q + r0 * Lq * Hq, enc_q, (size_t)nrows * Lq * Hq * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (io.sess_h) {
    CAIR_CUDA(cudaMemcpy2DAsync(io.sess_h + r0 * (Hsq + Hsd), (size_t)(Hsq + Hsd) * 4, Qs, (size_t)Hsq * 4, (size_t)Hsq * 4, nrows, cudaMemcpyDeviceToDevice, s));
    CAIR_CUDA(cudaMemcpy2DAsync(io.sess_h + r0 * (Hsq + Hsd) + Hsq, (size_t)(Hsq + Hsd) * 4, Ds, (size_t)Hsd * 4, (size_t)Hsd * 4, nrows, cudaMemcpyDeviceToDevice, s));
  }
  if (io.sess_c) {
    CAIR_CUDA(cudaMemcpy2DAsync(io.sess_c + r0 * (Hsq + Hsd), (size_t)(Hsq + Hsd) * 4, Qc, (size_t)Hsq * 4, (size_t)Hsq * 4, nrows, cudaMemcpyDeviceToDevice, s));
    CAIR_CUDA(cudaMemcpy2DAsync(io.sess_c + r0 * (Hsq + Hsd) + Hsq, (size_t)(Hsq + Hsd) * 4, Dc, (size_t)Hsd * 4, (size_t)Hsd * 4, nrows, cudaMemcpyDeviceToDevice, s));
  }
  // optional stage outputs (decoder-side session summaries included)
  if (io.pooled_q) CAIR_CUDA(cudaMemcpyAsync(io.pooled_q + r0 * Hq, pq, (size_t)nrows * Hq * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (io.pooled_d) CAIR_CUDA(cudaMemcpyAsync(io.pooled_d + r0 * N * Hd, pd, (size_t)ndocs * Hd * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (io.clicks) CAIR_CUDA(cudaMemcpyAsync(io.clicks + r0 * Hd, clk, (size_t)nrows * Hd * sizeof(float), cudaMemcpyDeviceToDevice, s));
  for (int side = 0; side < 2; ++side) {
    float* outp = side ? io.sess_d_attn : io.sess_q_attn;
    if (!outp) continue;
    const AttnPack& ap = side ? st.sd_inner : st.sq_inner;
    const float* states = side ? Ds : Qs;
    const int H = ap.H;
    CAIR_TRY(gemm_f32(gemm_dense(states, H), ap.w0, ap.b0, hid_s, H, nrows, H, H, ACT_TANH, s));
    CAIR_LAUNCH(rowdot_kernel, (unsigned)((nrows + 7) / 8), 256, 0, s, hid_s, nrows, H, ap.w3, ap.b3, sc_s);
    CAIR_LAUNCH(inner_attn_kernel, (unsigned)nrows, 256, 0, s, states, sc_s, S, H, outp + r0 * H);
  }
  return CAIR_OK;
}

}  // namespace cair
