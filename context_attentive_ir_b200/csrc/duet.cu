// DUET (neuroir/rankers/duet.py): LocalModel :65-121, DistributedModel :127-208, sum :58.
//
// Local model: the binary overlap matrix X[j,i] = (d_j == q_i) is never built; per pair the
// matching doc positions of every query term are compacted (ballot) and the k=1 Conv1d over
// the Ld "channels" becomes a sum of the matching columns of its (transposed) weight.
// Distributed model: the two k=3 Conv1d are GEMMs whose A rows are windows of three embedding
// rows gathered inside the GEMM (no [BN,Ld,E] tensor, no im2col buffer); max_pool1d(5,1) is
// fused into the A load of the 1x1 conv GEMM; Hadamard product with the query vector + fc2
// over time + tanh is one reduction kernel; the 300x300 layers are batched GEMMs over pairs.
#include "models.cuh"

namespace cair {

__global__ void duet_pack_kernel(const float* __restrict__ lconv, int nf, int Ld, float* __restrict__ lconv_t,
                                 const float* __restrict__ cq, const float* __restrict__ cd1, int E,
                                 float* __restrict__ cq_o, float* __restrict__ cd1_o) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)nf * Ld) {  // [nf][Ld][1] -> [Ld][nf]
    int f = (int)(i / Ld), j = (int)(i % Ld);
    lconv_t[(size_t)j * nf + f] = lconv[i];
  }
  if (i < (int64_t)nf * E * 3) {  // Conv1d weight [nf][E][3] -> [nf][3][E]  (K index = tap*E + e)
    int k = (int)(i % 3), e = (int)((i / 3) % E), f = (int)(i / (3 * E));
    size_t o = ((size_t)f * 3 + k) * E + e;
    cq_o[o] = cq[i];
    cd1_o[o] = cd1[i];
  }
}

int32_t duet_create_state(Owned& own, const cair_duet_weights& w, DuetState* st, cudaStream_t s) {
  const int nf = w.nfilters, E = w.emsize, Lq = w.max_query_len, Ld = w.max_doc_len;
  if (Ld - w.pool_size - 1 <= 0 || Lq < 3) return fail(CAIR_ERR_BAD_SHAPE, "duet_create: max_doc_len/max_query_len too small");
  st->V = w.vocab, st->E = E, st->nf = nf, st->pool = w.pool_size, st->Lq = Lq, st->Ld = Ld;
  const cair_linear* all[] = {&w.local_conv1d, &w.local_fc1, &w.local_fc2, &w.local_fc3, &w.conv_q, &w.conv_d1,
                              &w.conv_d2,      &w.dist_fc1,  &w.dist_fc2,  &w.dist_fc3,  &w.dist_fc4};
  for (const cair_linear* l : all)
    if (!l->w || !l->b) return fail(CAIR_ERR_BAD_ARG, "duet_create: null weight pointer");
  CAIR_TRY(dev_copy(own, w.table, (size_t)w.vocab * E, &st->table, s));
  CAIR_CUDA(own.alloc(&st->lconv_t, (size_t)Ld * nf));
  CAIR_CUDA(own.alloc(&st->cq_w, (size_t)nf * 3 * E));
  CAIR_CUDA(own.alloc(&st->cd1_w, (size_t)nf * 3 * E));
  int64_t total = (int64_t)nf * (3 * E > Ld ? 3 * E : Ld);
  CAIR_LAUNCH(duet_pack_kernel, (unsigned)((total + 255) / 256), 256, 0, s, w.local_conv1d.w, nf, Ld, st->lconv_t,
              w.conv_q.w, w.conv_d1.w, E, st->cq_w, st->cd1_w);
  CAIR_TRY(gemm_tc_pack(own, st->cq_w, nf, 3 * E, &st->cq_tc, s));
  CAIR_TRY(gemm_tc_pack(own, st->cd1_w, nf, 3 * E, &st->cd1_tc, s));
  CAIR_TRY(dev_copy(own, w.local_conv1d.b, (size_t)nf, &st->lconv_b, s));
  CAIR_TRY(dev_copy(own, w.local_fc1.w, (size_t)Lq, &st->lfc1_w, s));
  CAIR_TRY(dev_copy(own, w.local_fc1.b, 1, &st->lfc1_b, s));
  CAIR_TRY(dev_copy(own, w.local_fc2.w, (size_t)nf * nf, &st->lfc2_w, s));
  CAIR_TRY(dev_copy(own, w.local_fc2.b, (size_t)nf, &st->lfc2_b, s));
  CAIR_TRY(dev_copy(own, w.local_fc3.w, (size_t)nf, &st->lfc3_w, s));
  CAIR_TRY(dev_copy(own, w.local_fc3.b, 1, &st->lfc3_b, s));
  CAIR_TRY(dev_copy(own, w.conv_q.b, (size_t)nf, &st->cq_b, s));
  CAIR_TRY(dev_copy(own, w.conv_d1.b, (size_t)nf, &st->cd1_b, s));
  CAIR_TRY(dev_copy(own, w.conv_d2.w, (size_t)nf * nf, &st->cd2_w, s));
  CAIR_TRY(dev_copy(own, w.conv_d2.b, (size_t)nf, &st->cd2_b, s));
  CAIR_TRY(gemm_tc_pack(own, st->cd2_w, nf, nf, &st->cd2_tc, s));
  CAIR_TRY(dev_copy(own, w.dist_fc1.w, (size_t)nf * nf, &st->fc1_w, s));
  CAIR_TRY(dev_copy(own, w.dist_fc1.b, (size_t)nf, &st->fc1_b, s));
  CAIR_TRY(dev_copy(own, w.dist_fc2.w, (size_t)(Ld - w.pool_size - 1), &st->fc2_w, s));
  CAIR_TRY(dev_copy(own, w.dist_fc2.b, 1, &st->fc2_b, s));
  CAIR_TRY(dev_copy(own, w.dist_fc3.w, (size_t)nf * nf, &st->fc3_w, s));
  CAIR_TRY(dev_copy(own, w.dist_fc3.b, (size_t)nf, &st->fc3_b, s));
  CAIR_TRY(dev_copy(own, w.dist_fc4.w, (size_t)nf, &st->fc4_w, s));
  CAIR_TRY(dev_copy(own, w.dist_fc4.b, 1, &st->fc4_b, s));
  CAIR_CUDA(cudaStreamCreateWithFlags(&st->side, cudaStreamNonBlocking));
  CAIR_CUDA(cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming));
  CAIR_CUDA(cudaEventCreateWithFlags(&st->ev_join, cudaEventDisableTiming));
  return CAIR_OK;
}

// ---- local model, stage 1: m1[p, f] = tanh(fc1(tanh(conv1d(X)))[f])  (duet.py:93-118) ----
// smem: qids[Lq] | dids[Ld] | cnt[Lq] | list[Lq][Ld]
__global__ void __launch_bounds__(256) duet_local_kernel(const int64_t* __restrict__ q, const int64_t* __restrict__ d,
                                                         int N, int Lq, int Ld, int nf, int64_t pair_begin,
                                                         const float* __restrict__ wt, const float* __restrict__ wb,
                                                         const float* __restrict__ fc1w, const float* __restrict__ fc1b,
                                                         float* __restrict__ m1) {
  extern __shared__ int smi[];
  int* qids = smi;
  int* dids = qids + Lq;
  int* cnt = dids + Ld;
  int* list = cnt + Lq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t pl = blockIdx.x, p = pair_begin + pl, b = p / N;
  for (int i = tid; i < Lq; i += 256) qids[i] = (int)q[b * Lq + i];
  for (int j = tid; j < Ld; j += 256) dids[j] = (int)d[p * Ld + j];
  __syncthreads();
  for (int i = warp; i < Lq; i += 8) {  // ordered compaction of the matching doc positions of query term i
    int n = 0;
    const int qi = qids[i];
    for (int j0 = 0; j0 < Ld; j0 += 32) {
      int j = j0 + lane;
      bool hit = j < Ld && dids[j] == qi;
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) list[i * Ld + n + __popc(m & ((1u << lane) - 1))] = j;
      n += __popc(m);
    }
    if (lane == 0) cnt[i] = n;
  }
  __syncthreads();
  for (int f = tid; f < nf; f += 256) {
    float s1 = fc1b[0];
    const float bf = wb[f];
    for (int i = 0; i < Lq; ++i) {
      float s = bf;
      const int n = cnt[i];
      for (int k = 0; k < n; ++k) s += wt[(size_t)list[i * Ld + k] * nf + f];
      s1 = fmaf(fc1w[i], tanhf(s), s1);
    }
    m1[pl * nf + f] = tanhf(s1);
  }
}

// max_pool1d(win, stride 1) over time (duet.py:168): out[b, t, f] = max_{k<win} x[b, t+k, f], one thread per 4 channels.
// Materialising the pooled tensor (76 MB at cfg5 N=10) costs ~30 us; pooling inside the GEMM's A loader multiplied its
// loads by `win` and made conv_d2 the slowest kernel of the model.
__global__ void timepool4_kernel(const float4* __restrict__ x, int T, int Tp, int win, int nf4, int64_t total,
                                 float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % nf4);
    const int64_t bt = i / nf4;
    const int t = (int)(bt % Tp);
    const int64_t b = bt / Tp;
    const float4* src = x + ((size_t)b * T + t) * nf4 + f;
    float4 m = src[0];
    for (int k = 1; k < win; ++k) {
      const float4 v = src[(size_t)k * nf4];
      m.x = fmaxf(m.x, v.x), m.y = fmaxf(m.y, v.y), m.z = fmaxf(m.z, v.z), m.w = fmaxf(m.w, v.w);
    }
    out[i] = m;
  }
}

// column max over time: out[b, f] = max_t x[b, t, f]   (max_pool1d over the whole length, duet.py:178)
__global__ void colmax_kernel(const float* __restrict__ x, int T, int nf, float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int f = threadIdx.x; f < nf; f += blockDim.x) {
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, x[((size_t)b * T + t) * nf + f]);
    out[(size_t)b * nf + f] = m;
  }
}

// m1[p, f] = tanh(fc2_b + sum_t fc2_w[t] * rq[b, f] * rd[p, t, f])   (duet.py:185-201)
__global__ void __launch_bounds__(128) duet_hadamard_kernel(const float* __restrict__ rd, const float* __restrict__ rq,
                                                            const float* __restrict__ fc2w,
                                                            const float* __restrict__ fc2b, int N, int Tp, int nf,
                                                            int64_t pair_begin, int64_t q_begin,
                                                            float* __restrict__ m1) {
  const int64_t pl = blockIdx.x, b = (pair_begin + pl) / N - q_begin;
  const int f = blockIdx.y * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const float rqf = rq[b * nf + f];
  const float* r = rd + (size_t)pl * Tp * nf + f;
  float s = fc2b[0];
  for (int t = 0; t < Tp; ++t) s = fmaf(fc2w[t], rqf * r[(size_t)t * nf], s);
  m1[pl * nf + f] = tanhf(s);
}

// score[p] = tanh(fc3_l . m2l[p] + b) + tanh(fc4 . m2d[p] + b)   (duet.py:120, :204, :58); one warp per pair
__global__ void duet_final_kernel(const float* __restrict__ m2l, const float* __restrict__ m2d,
                                  const float* __restrict__ w3, const float* __restrict__ b3,
                                  const float* __restrict__ w4, const float* __restrict__ b4, int nf,
                                  int64_t pair_begin, int64_t pc, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t pl = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pl >= pc) return;
  float a = 0.f, c = 0.f;
  for (int f = lane; f < nf; f += 32) {
    a = fmaf(w3[f], m2l[pl * nf + f], a);
    c = fmaf(w4[f], m2d[pl * nf + f], c);
  }
  a = warp_sum(a), c = warp_sum(c);
  if (lane == 0) scores[pair_begin + pl] = tanhf(a + b3[0]) + tanhf(c + b4[0]);
}

int32_t duet_forward(const DuetState& st, const int64_t* q, const int64_t* d, int B, int N, int Lq, int Ld,
                     int64_t pb, int64_t pc, float* scores, Arena& ws, int* err, cudaStream_t s, bool dry) {
  (void)B;
  // fc1 / fc2 / conv1d sizes are baked into the weights (duet.py:69,73,144): batches must be force-padded
  if (Lq != st.Lq || Ld != st.Ld)
    return fail(CAIR_ERR_BAD_SHAPE, "duet: batch padded to (%d,%d) but the model was built for (%d,%d)", Lq, Ld, st.Lq, st.Ld);
  const int nf = st.nf, E = st.E, Tq = Lq - 2, Td = Ld - 2, Tp = Td - st.pool + 1;
  const int64_t qb = pc > 0 ? pb / N : 0;
  const int64_t nq = pc > 0 ? (pb + pc - 1) / N - qb + 1 : 0;
  float* m1l = ws.take<float>((size_t)pc * nf);
  float* m2l = ws.take<float>((size_t)pc * nf);
  float* cqv = ws.take<float>((size_t)nq * Tq * nf);
  float* mq = ws.take<float>((size_t)nq * nf);
  float* rq = ws.take<float>((size_t)nq * nf);
  float* cdv = ws.take<float>((size_t)pc * Td * nf);
  float* rd = ws.take<float>((size_t)pc * Tp * nf);
  float* pooled = ws.take<float>((nf & 3) ? 0 : (size_t)pc * Tp * nf);
  float* m1d = ws.take<float>((size_t)pc * nf);
  float* m2d = ws.take<float>((size_t)pc * nf);
  // A operand image of the pooled-row GEMM, whose weights span two column tiles (nf = 300): its rows are converted once instead
  // of once per column tile.  (conv_d1 keeps the in-kernel gather: the image of its 3-token windows would be three times the rows.)
  uint8_t* aimg = nullptr;
  if (st.cd2_tc.img && st.cd2_tc.nct >= 2) {
    aimg = ws.take<uint8_t>(gemm_tc_aimg_bytes(pc * Tp, nf) + 128);
    aimg = reinterpret_cast<uint8_t*>(((uintptr_t)aimg + 127) & ~(uintptr_t)127);
  }
  if (dry || pc <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "duet: workspace too small");
  // local model
  size_t smem = (size_t)(Lq + Ld + Lq + Lq * Ld) * sizeof(int);
  if (smem > 200 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "duet: Lq*Ld too large for the local model kernel");
  if (smem > 48 * 1024)
    CAIR_CUDA(cudaFuncSetAttribute(duet_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // The local model (ids only) and the query side of the distributed model are small, latency-sized kernels that do not
  // depend on the document convolutions: forked onto the handle's side stream (events keep the caller's stream semantics
  // and stay graph-capturable), joined before the head.
  cudaStream_t sq = st.side ? st.side : s;
  if (st.side) {
    CAIR_CUDA(cudaEventRecord(st.ev_fork, s));
    CAIR_CUDA(cudaStreamWaitEvent(st.side, st.ev_fork, 0));
  } else {
    prof_mark("local_model", s);
  }
  CAIR_LAUNCH(duet_local_kernel, (unsigned)pc, 256, smem, sq, q, d, N, Lq, Ld, nf, pb, st.lconv_t, st.lconv_b,
              st.lfc1_w, st.lfc1_b, m1l);
  CAIR_TRY(gemm_f32(gemm_dense(m1l, nf), st.lfc2_w, st.lfc2_b, m2l, nf, pc, nf, nf, ACT_TANH, sq));
  // distributed model, query side
  if (!st.side) prof_mark("query_conv", s);
  CAIR_TRY(gemm_auto(gemm_gather(st.table, st.V, E, q + qb * Lq, 3, Lq, Tq, err), st.cq_w, st.cq_tc, st.cq_b, cqv, nf,
                     nq * Tq, nf, 3 * E, ACT_TANH, sq));
  CAIR_LAUNCH(colmax_kernel, (unsigned)nq, 256, 0, sq, cqv, Tq, nf, mq);
  CAIR_TRY(gemm_f32(gemm_dense(mq, nf), st.fc1_w, st.fc1_b, rq, nf, nq, nf, nf, ACT_TANH, sq));
  if (st.side) CAIR_CUDA(cudaEventRecord(st.ev_join, st.side));
  // distributed model, document side
  prof_mark("conv_d1", s);
  CAIR_TRY(gemm_auto(gemm_gather(st.table, st.V, E, d + pb * Ld, 3, Ld, Td, err), st.cd1_w, st.cd1_tc, st.cd1_b, cdv, nf,
                     pc * Td, nf, 3 * E, ACT_TANH, s));
  prof_mark("pool_conv_d2", s);
  if ((nf & 3) == 0) {
    const int64_t total = (int64_t)pc * Tp * (nf / 4);
    CAIR_LAUNCH(timepool4_kernel, (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, s,
                reinterpret_cast<const float4*>(cdv), Td, Tp, st.pool, nf / 4, total, reinterpret_cast<float4*>(pooled));
    CAIR_TRY(gemm_auto(gemm_dense(pooled, nf), st.cd2_w, st.cd2_tc, st.cd2_b, rd, nf, pc * Tp, nf, nf, ACT_TANH, s, aimg));
  } else {
    CAIR_TRY(gemm_auto(gemm_pooled(cdv, nf, st.pool, Td, Tp), st.cd2_w, st.cd2_tc, st.cd2_b, rd, nf, pc * Tp, nf, nf,
                       ACT_TANH, s));
  }
  if (st.side) {
    prof_mark("join_query_side", s);
    CAIR_CUDA(cudaStreamWaitEvent(s, st.ev_join, 0));
  }
  prof_mark("head", s);
  CAIR_LAUNCH(duet_hadamard_kernel, dim3((unsigned)pc, (nf + 127) / 128), 128, 0, s, rd, rq, st.fc2_w, st.fc2_b, N, Tp,
              nf, pb, qb, m1d);
  CAIR_TRY(gemm_f32(gemm_dense(m1d, nf), st.fc3_w, st.fc3_b, m2d, nf, pc, nf, nf, ACT_TANH, s));
  CAIR_LAUNCH(duet_final_kernel, (unsigned)((pc + 7) / 8), 256, 0, s, m2l, m2d, st.lfc3_w, st.lfc3_b, st.fc4_w,
              st.fc4_b, nf, pb, pc, scores);
  return CAIR_OK;
}

}  // namespace cair
