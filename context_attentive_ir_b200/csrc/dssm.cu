// DSSM (neuroir/rankers/dssm.py:33-63) and CDSSM (neuroir/rankers/cdssm.py:42-77).
//
// DSSM: max-pool of the gathered embedding rows over ALL positions (zero PAD rows take part) in a
// streaming gather kernel, the two tanh layers as batched GEMMs over sequences, cosine per pair.
// CDSSM: the window-3 interleave followed by a k=3 Conv1d is one valid convolution over FIVE consecutive tokens with
// merged weights W5[f, s, e] = sum_{k+w=s} W[f, w*E+e, k]; it runs as a GEMM whose A rows are five embedding rows
// gathered by token id inside the GEMM (no [n, L-2, 3E] interleaved tensor), then tanh, Linear, tanh, max over
// positions, cosine.
#include "models.cuh"

namespace cair {

// ---- max-pool of table rows over the L tokens of each sequence: one CTA per sequence ----
__global__ void __launch_bounds__(256) maxpool_rows_kernel(const float* __restrict__ table, int V, int E,
                                                           const int64_t* __restrict__ ids, int L,
                                                           float* __restrict__ out, int* err) {
  extern __shared__ __align__(16) float part[];  // [ngroups][E]
  const int64_t s = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t* sid = ids + s * L;
  if ((E & 3) == 0 && E / 4 <= 256) {
    const int E4 = E >> 2, ngroups = 256 / E4;
    const int g = tid / E4, c = tid - g * E4;
    if (g < ngroups) {
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      for (int t = g; t < L; t += ngroups) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(table + checked_id(sid[t], V, err) * E) + c);
        m.x = fmaxf(m.x, v.x), m.y = fmaxf(m.y, v.y), m.z = fmaxf(m.z, v.z), m.w = fmaxf(m.w, v.w);
      }
      reinterpret_cast<float4*>(part + (size_t)g * E)[c] = m;
    }
    __syncthreads();
    for (int e = tid; e < E; e += 256) {
      float m = -INFINITY;
      for (int g2 = 0; g2 < ngroups; ++g2) m = fmaxf(m, part[(size_t)g2 * E + e]);
      out[s * E + e] = m;
    }
  } else {
    for (int e = tid; e < E; e += 256) {
      float m = -INFINITY;
      for (int t = 0; t < L; ++t) m = fmaxf(m, table[checked_id(sid[t], V, err) * E + e]);
      out[s * E + e] = m;
    }
  }
}

// column max over time: out[s, f] = max_t x[s, t, f]
__global__ void colmax_t_kernel(const float* __restrict__ x, int T, int nf, float* __restrict__ out) {
  const int64_t s = blockIdx.x;
  for (int f = threadIdx.x; f < nf; f += blockDim.x) {
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, x[((size_t)s * T + t) * nf + f]);
    out[(size_t)s * nf + f] = m;
  }
}

// torch>=2 cosine_similarity(query_rep, doc_rep): normalise (clamped at eps) first, then dot; one warp per pair
__global__ void cosine_pairs_kernel(const float* __restrict__ rq, const float* __restrict__ rd, int O, int N,
                                    int64_t pair_begin, int64_t pair_count, int64_t q_begin,
                                    float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t pl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pl >= pair_count) return;
  const float* a = rq + ((pair_begin + pl) / N - q_begin) * O;
  const float* b = rd + pl * O;
  float sa = 0.f, sb = 0.f;
  for (int k = lane; k < O; k += 32) sa = fmaf(a[k], a[k], sa), sb = fmaf(b[k], b[k], sb);
  const float na = fmaxf(sqrtf(warp_sum(sa)), 1e-8f), nb = fmaxf(sqrtf(warp_sum(sb)), 1e-8f);
  float dot = 0.f;
  for (int k = lane; k < O; k += 32) dot = fmaf(a[k] / na, b[k] / nb, dot);
  dot = warp_sum(dot);
  if (lane == 0) scores[pair_begin + pl] = dot;
}

static size_t pool_smem(int E) { return ((E & 3) == 0 && E / 4 <= 256) ? (size_t)(256 / (E / 4)) * E * sizeof(float) : 0; }

int32_t dssm_create_state(Owned& own, const cair_dssm_weights& w, DssmState* st, cudaStream_t s) {
  st->V = w.vocab, st->E = w.emsize, st->H = w.nhid, st->O = w.nout;
  CAIR_TRY(dev_copy(own, w.table, (size_t)w.vocab * w.emsize, &st->table, s));
  const cair_linear* src[4] = {&w.query_mlp0, &w.query_mlp2, &w.doc_mlp0, &w.doc_mlp2};
  for (int i = 0; i < 4; ++i) {
    const int in = (i & 1) ? w.nhid : w.emsize, out = (i & 1) ? w.nout : w.nhid;
    if (!src[i]->w || !src[i]->b) return fail(CAIR_ERR_BAD_ARG, "dssm_create: null weight pointer");
    CAIR_TRY(dev_copy(own, src[i]->w, (size_t)out * in, &st->w[i], s));
    CAIR_TRY(dev_copy(own, src[i]->b, (size_t)out, &st->b[i], s));
  }
  return CAIR_OK;
}

int32_t dssm_forward(const DssmState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb,
                     int64_t pc, float* scores, Arena& ws, int* err, cudaStream_t s, bool dry) {
  const int64_t qb = pc > 0 ? pb / N : 0, nq = pc > 0 ? (pb + pc - 1) / N - qb + 1 : 0;
  float* pq = ws.take<float>((size_t)nq * st.E);
  float* pd = ws.take<float>((size_t)pc * st.E);
  float* hq = ws.take<float>((size_t)nq * st.H);
  float* hd = ws.take<float>((size_t)pc * st.H);
  float* rq = ws.take<float>((size_t)nq * st.O);
  float* rd = ws.take<float>((size_t)pc * st.O);
  if (dry || pc <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "dssm: workspace too small");
  const size_t sm = pool_smem(st.E);
  if (sm > 48 * 1024) CAIR_CUDA(cudaFuncSetAttribute(maxpool_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  CAIR_LAUNCH(maxpool_rows_kernel, (unsigned)nq, 256, sm, s, st.table, st.V, st.E, q + qb * Lq, Lq, pq, err);
  CAIR_LAUNCH(maxpool_rows_kernel, (unsigned)pc, 256, sm, s, st.table, st.V, st.E, d + pb * Ld, Ld, pd, err);
  CAIR_TRY(gemm_f32(gemm_dense(pq, st.E), st.w[0], st.b[0], hq, st.H, nq, st.H, st.E, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(hq, st.H), st.w[1], st.b[1], rq, st.O, nq, st.O, st.H, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(pd, st.E), st.w[2], st.b[2], hd, st.H, pc, st.H, st.E, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(hd, st.H), st.w[3], st.b[3], rd, st.O, pc, st.O, st.H, ACT_TANH, s));
  CAIR_LAUNCH(cosine_pairs_kernel, (unsigned)((pc + 7) / 8), 256, 0, s, rq, rd, st.O, N, pb, pc, qb, scores);
  return CAIR_OK;
}

// W5[f][s*E + e] = sum_{k + w = s} conv.weight[f][w*E + e][k]      (conv.weight is [L, 3E, 3])
__global__ void cdssm_merge_kernel(const float* __restrict__ w, int Lh, int E, float* __restrict__ w5) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Lh * 5 * E) return;
  const int e = (int)(i % E), sft = (int)((i / E) % 5), f = (int)(i / ((int64_t)5 * E));
  float v = 0.f;
  for (int k = 0; k < 3; ++k) {
    const int wi = sft - k;
    if (wi >= 0 && wi < 3) v += w[((size_t)f * 3 * E + (size_t)wi * E + e) * 3 + k];
  }
  w5[i] = v;
}

void cdssm_merge_launch(const float* w, int H, int E, float* w5, cudaStream_t s) {
  const int64_t n5 = (int64_t)H * 5 * E;
  cdssm_merge_kernel<<<(unsigned)((n5 + 255) / 256), 256, 0, s>>>(w, H, E, w5);
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

int32_t cdssm_create_state(Owned& own, const cair_cdssm_weights& w, CdssmState* st, cudaStream_t s) {
  st->V = w.vocab, st->E = w.emsize, st->H = w.nhid, st->O = w.nout;
  CAIR_TRY(dev_copy(own, w.table, (size_t)w.vocab * w.emsize, &st->table, s));
  const cair_linear* conv[2] = {&w.query_conv, &w.doc_conv};
  const cair_linear* sem[2] = {&w.query_sem, &w.doc_sem};
  for (int i = 0; i < 2; ++i) {
    if (!conv[i]->w || !conv[i]->b || !sem[i]->w || !sem[i]->b) return fail(CAIR_ERR_BAD_ARG, "cdssm_create: null weight pointer");
    const int64_t n5 = (int64_t)w.nhid * 5 * w.emsize;
    CAIR_CUDA(own.alloc(&st->w5[i], (size_t)n5));
    CAIR_LAUNCH(cdssm_merge_kernel, (unsigned)((n5 + 255) / 256), 256, 0, s, conv[i]->w, w.nhid, w.emsize, st->w5[i]);
    CAIR_TRY(dev_copy(own, conv[i]->b, (size_t)w.nhid, &st->b5[i], s));
    CAIR_TRY(dev_copy(own, sem[i]->w, (size_t)w.nout * w.nhid, &st->ws[i], s));
    CAIR_TRY(dev_copy(own, sem[i]->b, (size_t)w.nout, &st->bs[i], s));
  }
  return CAIR_OK;
}

int32_t cdssm_forward(const CdssmState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb,
                      int64_t pc, float* scores, Arena& ws, int* err, cudaStream_t s, bool dry) {
  if (Lq < 5 || Ld < 5) return fail(CAIR_ERR_BAD_SHAPE, "cdssm: sequences shorter than 5 tokens (cdssm.py:32-40 + k=3 conv)");
  const int64_t qb = pc > 0 ? pb / N : 0, nq = pc > 0 ? (pb + pc - 1) / N - qb + 1 : 0;
  const int Tq = Lq - 4, Td = Ld - 4;
  float* hq = ws.take<float>((size_t)nq * Tq * st.H);
  float* hd = ws.take<float>((size_t)pc * Td * st.H);
  float* sq = ws.take<float>((size_t)nq * Tq * st.O);
  float* sd = ws.take<float>((size_t)pc * Td * st.O);
  float* rq = ws.take<float>((size_t)nq * st.O);
  float* rd = ws.take<float>((size_t)pc * st.O);
  if (dry || pc <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "cdssm: workspace too small");
  CAIR_TRY(gemm_f32(gemm_gather(st.table, st.V, st.E, q + qb * Lq, 5, Lq, Tq, err), st.w5[0], st.b5[0], hq, st.H, nq * Tq,
                    st.H, 5 * st.E, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(hq, st.H), st.ws[0], st.bs[0], sq, st.O, nq * Tq, st.O, st.H, ACT_TANH, s));
  CAIR_LAUNCH(colmax_t_kernel, (unsigned)nq, 128, 0, s, sq, Tq, st.O, rq);
  CAIR_TRY(gemm_f32(gemm_gather(st.table, st.V, st.E, d + pb * Ld, 5, Ld, Td, err), st.w5[1], st.b5[1], hd, st.H, pc * Td,
                    st.H, 5 * st.E, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(hd, st.H), st.ws[1], st.bs[1], sd, st.O, pc * Td, st.O, st.H, ACT_TANH, s));
  CAIR_LAUNCH(colmax_t_kernel, (unsigned)pc, 128, 0, s, sd, Td, st.O, rd);
  CAIR_LAUNCH(cosine_pairs_kernel, (unsigned)((pc + 7) / 8), 256, 0, s, rq, rd, st.O, N, pb, pc, qb, scores);
  return CAIR_OK;
}

}  // namespace cair
