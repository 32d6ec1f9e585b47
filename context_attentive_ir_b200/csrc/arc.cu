// ARC-I (neuroir/rankers/arci.py:60-105) and ARC-II (neuroir/rankers/arcii.py:58-111).
//
// Every convolution is a GEMM whose A rows are built on the fly by the generic providers (common.cuh): layer 0
// gathers a same-padded window of embedding rows by token id (no [n, L, E] tensor, no im2col buffer), deeper
// layers read same-padded 1-D / 3x3 windows of the previous NHWC activation.  Activations stay position-major
// ([n, L, C] / [n, H, W, C]); the reference's channel-major flatten(1) is absorbed into a one-off column
// permutation of the first MLP layer at create.  ARC-I's concat(query, doc) -> Linear splits into a per-query and a
// per-doc GEMM.  ARC-II's broadcast sum followed by MaxPool2d(2,2) is separable: max_{2x2}(d_j + q_i) =
// max(d_2y, d_2y+1) + max(q_2x, q_2x+1), so the [BN, F, Ld, Lq] tensor is never formed.
#include "models.cuh"

namespace cair {

// conv weights [F][C][k...] -> [F][tap*C + c]   (tap = k index for Conv1d, ky*3+kx for Conv2d)
__global__ void arc_pack_conv_kernel(const float* __restrict__ w, int F, int C, int taps, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)F * C * taps) return;
  const int t = (int)(i % taps), c = (int)((i / taps) % C), f = (int)(i / ((int64_t)taps * C));
  out[((size_t)f * taps + t) * C + c] = w[i];
}
// MLP columns: reference index c*S + s (channel-major flatten) -> s*C + c, inside the column block [col0, col0 + C*S)
__global__ void arc_pack_mlp_kernel(const float* __restrict__ w, int rows, int in_total, int col0, int C, int S,
                                    float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * C * S) return;
  const int s = (int)(i % S), c = (int)((i / S) % C), o = (int)(i / ((int64_t)S * C));
  out[(size_t)o * C * S + (size_t)s * C + c] = w[(size_t)o * in_total + col0 + (size_t)c * S + s];
}

__global__ void maxpool1d_kernel(const float* __restrict__ x, int64_t n, int L, int C, int P, float* __restrict__ y) {
  const int Lp = L / P;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * Lp * C) return;
  const int c = (int)(i % C), t = (int)((i / C) % Lp);
  const int64_t s = i / ((int64_t)C * Lp);
  float m = -INFINITY;
  for (int u = 0; u < P; ++u) m = fmaxf(m, x[(s * L + t * P + u) * C + c]);
  y[i] = m;
}
__global__ void maxpool2d_kernel(const float* __restrict__ x, int64_t n, int H, int W, int C, float* __restrict__ y) {
  const int H2 = H / 2, W2 = W / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * H2 * W2 * C) return;
  const int c = (int)(i % C), xx = (int)((i / C) % W2), yy = (int)((i / ((int64_t)C * W2)) % H2);
  const int64_t s = i / ((int64_t)C * W2 * H2);
  const float* b = x + ((s * H + 2 * yy) * W + 2 * xx) * C + c;
  y[i] = fmaxf(fmaxf(b[0], b[C]), fmaxf(b[(size_t)W * C], b[(size_t)W * C + C]));
}
// comb[p, y, x, c] = max(cd[p, 2y, c], cd[p, 2y+1, c]) + max(cq[b, 2x, c], cq[b, 2x+1, c])   (arcii.py:97-101)
__global__ void arcii_comb_kernel(const float* __restrict__ cq, const float* __restrict__ cd, int N, int Lq, int Ld, int C,
                                  int64_t pair_begin, int64_t pair_count, int64_t q_begin, float* __restrict__ out) {
  const int H = Ld / 2, W = Lq / 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pair_count * H * W * C) return;
  const int c = (int)(i % C), x = (int)((i / C) % W), y = (int)((i / ((int64_t)C * W)) % H);
  const int64_t pl = i / ((int64_t)C * W * H);
  const int64_t ql = (pair_begin + pl) / N - q_begin;
  const float* dq = cq + (ql * Lq + 2 * x) * C + c;
  const float* dd = cd + (pl * Ld + 2 * y) * C + c;
  out[i] = fmaxf(dd[0], dd[C]) + fmaxf(dq[0], dq[C]);
}
// ARC-I head: score[p] = b1 + sum_o w1[o] * (hq[b][o] + hd[p][o])   (mlp.0 bias already inside hq); one warp per pair
__global__ void arci_score_kernel(const float* __restrict__ hq, const float* __restrict__ hd, const float* __restrict__ w1,
                                  const float* __restrict__ b1, int hid, int N, int64_t pair_begin, int64_t pair_count,
                                  int64_t q_begin, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t pl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pl >= pair_count) return;
  const float* a = hq + ((pair_begin + pl) / N - q_begin) * hid;
  const float* b = hd + pl * hid;
  float s = 0.f;
  for (int o = lane; o < hid; o += 32) s = fmaf(w1[o], a[o] + b[o], s);
  s = warp_sum(s);
  if (lane == 0) scores[pair_begin + pl] = s + b1[0];
}
__global__ void arc_rowdot_kernel(const float* __restrict__ x, int64_t rows, int H, const float* __restrict__ w,
                                  const float* __restrict__ b, int64_t out_offset, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float a = 0.f;
  for (int k = lane; k < H; k += 32) a = fmaf(w[k], x[r * H + k], a);
  a = warp_sum(a);
  if (lane == 0) out[out_offset + r] = a + b[0];
}

static int32_t pack_conv(Owned& own, const cair_linear& l, int F, int C, int taps, float** w, float** b, GemmTcW* tc,
                         cudaStream_t s) {
  if (!l.w || !l.b) return fail(CAIR_ERR_BAD_ARG, "arc: null conv weights");
  const int64_t n = (int64_t)F * C * taps;
  CAIR_CUDA(own.alloc(w, (size_t)n));
  CAIR_LAUNCH(arc_pack_conv_kernel, (unsigned)((n + 255) / 256), 256, 0, s, l.w, F, C, taps, *w);
  CAIR_TRY(dev_copy(own, l.b, (size_t)F, b, s));
  if (n >= 32 * 1024) CAIR_TRY(gemm_tc_pack(own, *w, F, C * taps, tc, s));
  return CAIR_OK;
}

// ---------------------------------------------------------------- ARC-I
int32_t arci_create_state(Owned& own, const cair_arci_weights& w, ArciState* st, cudaStream_t s) {
  if (w.nlayers < 1 || w.nlayers > CAIR_ARC_MAX_LAYERS) return fail(CAIR_ERR_UNSUPPORTED, "arci: 1..%d conv layers", CAIR_ARC_MAX_LAYERS);
  st->V = w.vocab, st->E = w.emsize, st->nl = w.nlayers, st->Lq = w.max_query_len, st->Ld = w.max_doc_len;
  CAIR_TRY(dev_copy(own, w.table, (size_t)w.vocab * w.emsize, &st->table, s));
  int C = w.emsize, lq = st->Lq, ld = st->Ld;
  for (int i = 0; i < w.nlayers; ++i) {
    st->F[i] = w.filters[i], st->k[i] = w.kernel[i], st->P[i] = w.pool[i];
    if (st->k[i] % 2 == 0 || st->P[i] < 1) return fail(CAIR_ERR_UNSUPPORTED, "arci: even kernel size / bad pool size");
    CAIR_TRY(pack_conv(own, w.qconv[i], st->F[i], C, st->k[i], &st->qw[i], &st->qb[i], &st->qtc[i], s));
    CAIR_TRY(pack_conv(own, w.dconv[i], st->F[i], C, st->k[i], &st->dw[i], &st->db[i], &st->dtc[i], s));
    C = st->F[i], lq /= st->P[i], ld /= st->P[i];
    if (lq <= 0 || ld <= 0) return fail(CAIR_ERR_BAD_SHAPE, "arci: sequence pooled away (arci.py:46-48)");
  }
  st->Lqo = lq, st->Ldo = ld;
  const int Fl = C, inp = Fl * (lq + ld), hid = inp / 2;
  st->hid = hid;
  if (!w.mlp0.w || !w.mlp0.b || !w.mlp1.w || !w.mlp1.b) return fail(CAIR_ERR_BAD_ARG, "arci: null mlp weights");
  CAIR_CUDA(own.alloc(&st->mq, (size_t)hid * Fl * lq));
  CAIR_CUDA(own.alloc(&st->md, (size_t)hid * Fl * ld));
  CAIR_LAUNCH(arc_pack_mlp_kernel, (unsigned)(((int64_t)hid * Fl * lq + 255) / 256), 256, 0, s, w.mlp0.w, hid, inp, 0, Fl, lq, st->mq);
  CAIR_LAUNCH(arc_pack_mlp_kernel, (unsigned)(((int64_t)hid * Fl * ld + 255) / 256), 256, 0, s, w.mlp0.w, hid, inp, Fl * lq, Fl, ld, st->md);
  CAIR_TRY(dev_copy(own, w.mlp0.b, (size_t)hid, &st->b0, s));
  CAIR_TRY(dev_copy(own, w.mlp1.w, (size_t)hid, &st->w1, s));
  CAIR_TRY(dev_copy(own, w.mlp1.b, 1, &st->b1, s));
  CAIR_TRY(gemm_tc_pack(own, st->md, hid, Fl * ld, &st->mdtc, s));
  return CAIR_OK;
}

// conv stack of one side; returns the final [n, Lout, Fl] activation in *out (workspace carved from ws)
static int32_t arci_side(const ArciState& st, bool doc, const int64_t* ids, int64_t n, int L, Arena& ws, float** out, int* err,
                         cudaStream_t s, bool dry) {
  const float* cur = nullptr;
  int C = st.E;
  for (int i = 0; i < st.nl; ++i) {
    float* y = ws.take<float>((size_t)n * L * st.F[i]);
    float* z = ws.take<float>((size_t)n * (L / st.P[i]) * st.F[i]);
    if (!dry && n > 0) {
      const float* w = doc ? st.dw[i] : st.qw[i];
      const float* b = doc ? st.db[i] : st.qb[i];
      const GemmTcW& tc = doc ? st.dtc[i] : st.qtc[i];
      GemmA a = (i == 0) ? gemm_gather(st.table, st.V, st.E, ids, st.k[i], L, L, err, st.k[i] / 2)
                         : gemm_window1d(cur, C, st.k[i], L);
      CAIR_TRY(gemm_auto(a, w, tc, b, y, st.F[i], n * L, st.F[i], st.k[i] * C, ACT_RELU, s));
      const int64_t tot = n * (L / st.P[i]) * st.F[i];
      CAIR_LAUNCH(maxpool1d_kernel, (unsigned)((tot + 255) / 256), 256, 0, s, y, n, L, st.F[i], st.P[i], z);
    }
    cur = z, C = st.F[i], L /= st.P[i];
    *out = z;
  }
  return CAIR_OK;
}

int32_t arci_forward(const ArciState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb, int64_t pc,
                     float* scores, Arena& ws, int* err, cudaStream_t s, bool dry) {
  if (Lq != st.Lq || Ld != st.Ld)
    return fail(CAIR_ERR_BAD_SHAPE, "arci: batch padded to (%d,%d) but the model was built for (%d,%d)", Lq, Ld, st.Lq, st.Ld);
  const int64_t qb = pc > 0 ? pb / N : 0, nq = pc > 0 ? (pb + pc - 1) / N - qb + 1 : 0;
  float *fq = nullptr, *fd = nullptr;
  CAIR_TRY(arci_side(st, false, q ? q + qb * Lq : nullptr, nq, Lq, ws, &fq, err, s, dry || pc <= 0));
  CAIR_TRY(arci_side(st, true, d ? d + pb * Ld : nullptr, pc, Ld, ws, &fd, err, s, dry || pc <= 0));
  float* hq = ws.take<float>((size_t)nq * st.hid);
  float* hd = ws.take<float>((size_t)pc * st.hid);
  if (dry || pc <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "arci: workspace too small");
  const int Fl = st.F[st.nl - 1];
  GemmTcW none;
  CAIR_TRY(gemm_auto(gemm_dense(fq, (int64_t)Fl * st.Lqo), st.mq, none, st.b0, hq, st.hid, nq, st.hid, Fl * st.Lqo, ACT_NONE, s));
  CAIR_TRY(gemm_auto(gemm_dense(fd, (int64_t)Fl * st.Ldo), st.md, st.mdtc, nullptr, hd, st.hid, pc, st.hid, Fl * st.Ldo, ACT_NONE, s));
  CAIR_LAUNCH(arci_score_kernel, (unsigned)((pc + 7) / 8), 256, 0, s, hq, hd, st.w1, st.b1, st.hid, N, pb, pc, qb, scores);
  return CAIR_OK;
}

// ---------------------------------------------------------------- ARC-II
int32_t arcii_create_state(Owned& own, const cair_arcii_weights& w, ArciiState* st, cudaStream_t s) {
  if (w.nlayers2d < 1 || w.nlayers2d > CAIR_ARC_MAX_LAYERS) return fail(CAIR_ERR_UNSUPPORTED, "arcii: 1..%d conv2d layers", CAIR_ARC_MAX_LAYERS);
  if (w.kernel_1d % 2 == 0) return fail(CAIR_ERR_UNSUPPORTED, "arcii: even kernel_size_1d");
  st->V = w.vocab, st->E = w.emsize, st->F1 = w.filters_1d, st->k1 = w.kernel_1d, st->nl = w.nlayers2d;
  st->Lq = w.max_query_len, st->Ld = w.max_doc_len;
  CAIR_TRY(dev_copy(own, w.table, (size_t)w.vocab * w.emsize, &st->table, s));
  CAIR_TRY(pack_conv(own, w.conv_query, st->F1, st->E, st->k1, &st->cqw, &st->cqb, &st->cqtc, s));
  CAIR_TRY(pack_conv(own, w.conv_doc, st->F1, st->E, st->k1, &st->cdw, &st->cdb, &st->cdtc, s));
  int C = st->F1, H = st->Ld / 2, W = st->Lq / 2;
  if (H <= 0 || W <= 0) return fail(CAIR_ERR_BAD_SHAPE, "arcii: map pooled away");
  for (int i = 0; i < st->nl; ++i) {
    st->F2[i] = w.filters_2d[i];
    CAIR_TRY(pack_conv(own, w.conv2d[i], st->F2[i], C, 9, &st->w2[i], &st->b2[i], &st->tc2[i], s));
    C = st->F2[i], H /= 2, W /= 2;
    if (H <= 0 || W <= 0) return fail(CAIR_ERR_BAD_SHAPE, "arcii: map pooled away (arcii.py:50-52)");
  }
  st->Hf = H, st->Wf = W, st->Cf = C;
  const int inp = C * H * W, hid = inp / 2;
  st->hid = hid;
  if (!w.mlp0.w || !w.mlp0.b || !w.mlp1.w || !w.mlp1.b) return fail(CAIR_ERR_BAD_ARG, "arcii: null mlp weights");
  CAIR_CUDA(own.alloc(&st->m0, (size_t)hid * inp));
  CAIR_LAUNCH(arc_pack_mlp_kernel, (unsigned)(((int64_t)hid * inp + 255) / 256), 256, 0, s, w.mlp0.w, hid, inp, 0, C, H * W, st->m0);
  CAIR_TRY(dev_copy(own, w.mlp0.b, (size_t)hid, &st->b0, s));
  CAIR_TRY(dev_copy(own, w.mlp1.w, (size_t)hid, &st->w1, s));
  CAIR_TRY(dev_copy(own, w.mlp1.b, 1, &st->b1, s));
  CAIR_TRY(gemm_tc_pack(own, st->m0, hid, inp, &st->m0tc, s));
  return CAIR_OK;
}

int32_t arcii_forward(const ArciiState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb, int64_t pc,
                      float* scores, Arena& ws, int* err, cudaStream_t s, bool dry) {
  if (Lq != st.Lq || Ld != st.Ld)
    return fail(CAIR_ERR_BAD_SHAPE, "arcii: batch padded to (%d,%d) but the model was built for (%d,%d)", Lq, Ld, st.Lq, st.Ld);
  const int64_t qb = pc > 0 ? pb / N : 0, nq = pc > 0 ? (pb + pc - 1) / N - qb + 1 : 0;
  const bool run = !dry && pc > 0;
  float* cq = ws.take<float>((size_t)nq * Lq * st.F1);
  float* cd = ws.take<float>((size_t)pc * Ld * st.F1);
  int C = st.F1, H = Ld / 2, W = Lq / 2;
  float* cur = ws.take<float>((size_t)pc * H * W * C);
  if (run) {
    if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "arcii: workspace too small");
    CAIR_TRY(gemm_auto(gemm_gather(st.table, st.V, st.E, q + qb * Lq, st.k1, Lq, Lq, err, st.k1 / 2), st.cqw, st.cqtc, st.cqb, cq,
                       st.F1, nq * Lq, st.F1, st.k1 * st.E, ACT_NONE, s));
    CAIR_TRY(gemm_auto(gemm_gather(st.table, st.V, st.E, d + pb * Ld, st.k1, Ld, Ld, err, st.k1 / 2), st.cdw, st.cdtc, st.cdb, cd,
                       st.F1, pc * Ld, st.F1, st.k1 * st.E, ACT_NONE, s));
    const int64_t tot = pc * H * W * C;
    CAIR_LAUNCH(arcii_comb_kernel, (unsigned)((tot + 255) / 256), 256, 0, s, cq, cd, N, Lq, Ld, C, pb, pc, qb, cur);
  }
  for (int i = 0; i < st.nl; ++i) {
    float* y = ws.take<float>((size_t)pc * H * W * st.F2[i]);
    float* z = ws.take<float>((size_t)pc * (H / 2) * (W / 2) * st.F2[i]);
    if (run) {
      if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "arcii: workspace too small");
      CAIR_TRY(gemm_auto(gemm_window2d(cur, C, H, W), st.w2[i], st.tc2[i], st.b2[i], y, st.F2[i], pc * H * W, st.F2[i], 9 * C,
                         ACT_RELU, s));
      const int64_t tot = pc * (H / 2) * (W / 2) * st.F2[i];
      CAIR_LAUNCH(maxpool2d_kernel, (unsigned)((tot + 255) / 256), 256, 0, s, y, pc, H, W, st.F2[i], z);
    }
    cur = z, C = st.F2[i], H /= 2, W /= 2;
  }
  float* h0 = ws.take<float>((size_t)pc * st.hid);
  if (!run) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "arcii: workspace too small");
  const int inp = C * H * W;
  CAIR_TRY(gemm_auto(gemm_dense(cur, inp), st.m0, st.m0tc, st.b0, h0, st.hid, pc, st.hid, inp, ACT_NONE, s));
  CAIR_LAUNCH(arc_rowdot_kernel, (unsigned)((pc + 7) / 8), 256, 0, s, h0, pc, st.hid, st.w1, st.b1, pb, scores);
  return CAIR_OK;
}

}  // namespace cair
