// Match-Tensor interaction (neuroir/rankers/mtensor.py:99-131, exact match :144-158).
//
// The reference materialises MT[BN, C+1, Lq, Ld] (three torch.stack broadcasts + product),
// runs three same-padded convs (3x3, 3x5, 3x7), ReLU, a 1x1 conv, two max-pools and a Linear.
// Here the three convs are merged into one zero-extended 3x7 stencil W7[f,c,a,bt] and, because
// the first C channels of MT are rank-1 in (i,j) (MT[c,i,j] = cq[i,c]*cd[j,c]), factorised:
//
//   conv[f,i,j] = bias[f] + sum_{bt,c} cd[j+bt-3, c] * T[i,bt,c,f]  +  exact-match taps
//   T[i,bt,c,f] = sum_{a : 0<=i+a-1<Lq} W7[f,c,a,bt] * cq[i+a-1, c]        (per QUERY, shared by its N docs)
//
// (zero padding of the product at the (Lq,Ld) borders == zero rows of cq / cd outside the
// borders; pad positions inside the borders keep the projection bias).
// One CTA per (query, doc) pair: cd is staged transposed in shared memory with a 3-row halo,
// T slices stream from L2, thread j owns output column j: conv + exact match + bias + ReLU +
// 1x1 conv + running max over (i,j) in registers; one 4-byte score store per pair.
// The [BN,C+1,Lq,Ld] tensor never exists.
#include "models.cuh"

namespace cair {

constexpr int MT_THREADS = 256;
constexpr int MT_MAXF = 24;   // 3*nfilters
constexpr int MT_MAXM = 32;   // match_filter_size


__global__ void mt_pack_kernel(const float* __restrict__ c1, const float* __restrict__ c2,
                               const float* __restrict__ c3, const float* __restrict__ cb1,
                               const float* __restrict__ cb2, const float* __restrict__ cb3,
                               const float* __restrict__ alpha, const float* __restrict__ conv_w,
                               const float* __restrict__ conv_b, const float* __restrict__ out_w,
                               const float* __restrict__ out_b, MtPack p) {
  const int C = p.C, nf = p.nf, FPP = p.FPP, C1 = C + 1;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 3 * 7 * C1 * FPP;
  if (idx < total) {
    int f = idx % FPP, c = (idx / FPP) % C1, bt = (idx / (FPP * C1)) % 7, a = idx / (FPP * C1 * 7);
    float v = 0.f;
    if (f < p.FP) {
      int k = f / nf, ff = f - k * nf;  // conv k has kernel width 3+2k, padding 1+k
      int kw = 3 + 2 * k, bb = bt - (2 - k);
      const float* w = k == 0 ? c1 : (k == 1 ? c2 : c3);
      if (bb >= 0 && bb < kw) v = w[(((size_t)ff * C1 + c) * 3 + a) * kw + bb];
    }
    if (c < C) {
      p.w7[(((size_t)a * 7 + bt) * C + c) * FPP + f] = v;
      if (f < p.FP) p.w7t[(((size_t)a * 7 + bt) * p.FP + f) * ((C + 15) & ~15) + c] = v;
    }
    else
      p.wem[((size_t)a * 7 + bt) * FPP + f] = v * alpha[0];
  }
  if (idx < FPP) {
    float v = 0.f;
    if (idx < p.FP) {
      int k = idx / nf, ff = idx - k * nf;
      v = (k == 0 ? cb1 : (k == 1 ? cb2 : cb3))[ff];
    }
    p.bias[idx] = v;
  }
  if (idx < p.M * FPP) {
    int m = idx / FPP, f = idx % FPP;
    p.w1[idx] = f < p.FP ? conv_w[(size_t)m * p.FP + f] : 0.f;
  }
  if (idx < p.M) {
    p.b1[idx] = conv_b[idx];
    p.wo[idx] = out_w[idx];
  }
  if (idx == 0) p.wo[p.M] = out_b[0];
}

int32_t mt_pack(Owned& own, const cair_mt_weights& w, MtPack* p, cudaStream_t s) {
  p->C = w.nchannels, p->nf = w.nfilters, p->FP = 3 * w.nfilters, p->M = w.match_filter_size;
  p->FPP = (p->FP + 3) & ~3;
  if (p->FP > MT_MAXF) return fail(CAIR_ERR_UNSUPPORTED, "match_tensor: nfilters %d > %d", w.nfilters, MT_MAXF / 3);
  if (p->M > MT_MAXM) return fail(CAIR_ERR_UNSUPPORTED, "match_tensor: match_filter_size %d > %d", p->M, MT_MAXM);
  CAIR_CUDA(own.alloc(&p->w7, (size_t)21 * p->C * p->FPP));
  {
    const size_t n7t = (size_t)21 * p->FP * ((p->C + 15) & ~15);
    CAIR_CUDA(own.alloc(&p->w7t, n7t));
    CAIR_CUDA(cudaMemsetAsync(p->w7t, 0, n7t * sizeof(float), s));
  }
  CAIR_CUDA(own.alloc(&p->wem, (size_t)21 * p->FPP));
  CAIR_CUDA(own.alloc(&p->bias, (size_t)p->FPP));
  CAIR_CUDA(own.alloc(&p->w1, (size_t)p->M * p->FPP));
  CAIR_CUDA(own.alloc(&p->b1, (size_t)p->M));
  CAIR_CUDA(own.alloc(&p->wo, (size_t)p->M + 1));
  int total = 21 * (p->C + 1) * p->FPP;
  if (total < p->M * p->FPP) total = p->M * p->FPP;
  CAIR_LAUNCH(mt_pack_kernel, (total + 255) / 256, 256, 0, s, w.conv1.w, w.conv2.w, w.conv3.w, w.conv1.b, w.conv2.b,
              w.conv3.b, w.alpha, w.conv.w, w.conv.b, w.output.w, w.output.b, *p);
  return CAIR_OK;
}

// T[qi][i][bt][c][f] for the queries [q_begin, q_begin+nq): grid (Lq, nq)
__global__ void __launch_bounds__(256) mt_build_t_kernel(const float* __restrict__ cq, MtPack p, int Lq,
                                                         float* __restrict__ T) {
  const int i = blockIdx.x, qi = blockIdx.y;
  const int C = p.C, FPP = p.FPP;
  const int n = 7 * C * FPP;
  float* out = T + ((size_t)qi * Lq + i) * n;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    int c = (idx / FPP) % C;
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int ii = i + a - 1;
      if (ii >= 0 && ii < Lq) v = fmaf(p.w7[(size_t)a * n + idx], cq[((size_t)qi * Lq + ii) * C + c], v);
    }
    out[idx] = v;
  }
}

// smem (floats): Dt[C][LdP] | Tsl[7*C*FPP] | wem[21*FPP] | bias[FPP] | w1[M*FPP] | b1[M] | red[M*8] | ids
// ARG (training, train.cu): also returns the pooled features [pairs, M] and, per (pair, m), the cell i*Ld + j of the maximum
// - the first one in row-major order, which is where torch's two max reductions (mtensor.py:128-129) route the gradient.
template <bool ARG>
__global__ void __launch_bounds__(MT_THREADS) mt_interact_kernel(const float* __restrict__ cd,
                                                                 const float* __restrict__ T, MtPack p,
                                                                 const int64_t* __restrict__ q,
                                                                 const int64_t* __restrict__ d, int N, int Lq, int Ld,
                                                                 int64_t pair_begin, int64_t q_begin,
                                                                 float* __restrict__ scores, float* __restrict__ pooled,
                                                                 int* __restrict__ argidx) {
  extern __shared__ __align__(16) float sm[];
  const int C = p.C, FPP = p.FPP, FP4 = FPP / 4, M = p.M;
  const int LdP = Ld + 6;
  float* Dt = sm;
  float* Tsl = Dt + (((size_t)C * LdP + 3) & ~(size_t)3);  // keep float4 alignment
  float* wem = Tsl + 7 * C * FPP;
  float* bias = wem + 21 * FPP;
  float* w1 = bias + FPP;
  float* b1 = w1 + M * FPP;
  float* red = b1 + MT_MAXM;
  int* dids = reinterpret_cast<int*>(red + MT_MAXM * (MT_THREADS / 32));
  int* qids = dids + LdP;
  const int tid = threadIdx.x;
  const int64_t pl = blockIdx.x;            // pair index local to this rank's slice
  const int64_t p_glob = pair_begin + pl;
  const int64_t b = p_glob / N;
  const int64_t ql = b - q_begin;           // query index local to the slice

  for (int i = tid; i < C * LdP; i += MT_THREADS) Dt[i] = 0.f;
  for (int i = tid; i < LdP; i += MT_THREADS) {
    int j = i - 3;
    dids[i] = (j >= 0 && j < Ld) ? (int)d[p_glob * Ld + j] : -1;
  }
  for (int i = tid; i < Lq; i += MT_THREADS) qids[i] = (int)q[b * Lq + i];
  for (int i = tid; i < 21 * FPP; i += MT_THREADS) wem[i] = p.wem[i];
  for (int i = tid; i < FPP; i += MT_THREADS) bias[i] = p.bias[i];
  for (int i = tid; i < M * FPP; i += MT_THREADS) w1[i] = p.w1[i];
  for (int i = tid; i < M; i += MT_THREADS) b1[i] = p.b1[i];
  __syncthreads();
  // stage cd[pair] transposed with a 3-row zero halo
  const float* cdp = cd + (size_t)pl * Ld * C;
  for (int i = tid; i < Ld * C; i += MT_THREADS) {
    int j = i / C, c = i - j * C;
    Dt[(size_t)c * LdP + j + 3] = cdp[i];
  }
  float mx[MT_MAXM];
  int ax[ARG ? MT_MAXM : 1];
#pragma unroll
  for (int m = 0; m < MT_MAXM; ++m) mx[m] = -INFINITY;
  if (ARG) {
#pragma unroll
    for (int m = 0; m < MT_MAXM; ++m) ax[ARG ? m : 0] = 0x7fffffff;
  }

  const float* Tq = T + (size_t)ql * Lq * 7 * C * FPP;
  for (int i = 0; i < Lq; ++i) {
    __syncthreads();  // previous slice consumed (and Dt staged, first iteration)
    {
      const float4* src = reinterpret_cast<const float4*>(Tq + (size_t)i * 7 * C * FPP);
      float4* dst = reinterpret_cast<float4*>(Tsl);
      for (int k = tid; k < 7 * C * FP4; k += MT_THREADS) dst[k] = src[k];
    }
    __syncthreads();
    for (int j = tid; j < Ld; j += MT_THREADS) {
      float acc[MT_MAXF];
#pragma unroll
      for (int f = 0; f < MT_MAXF; ++f) acc[f] = (f < FPP) ? bias[f] : 0.f;
      for (int bt = 0; bt < 7; ++bt) {
        const float* dcol = Dt + j + bt;
        const float4* tb = reinterpret_cast<const float4*>(Tsl + (size_t)bt * C * FPP);
        for (int c = 0; c < C; ++c) {
          float dv = dcol[(size_t)c * LdP];
#pragma unroll
          for (int f4 = 0; f4 < MT_MAXF / 4; ++f4) {
            if (f4 < FP4) {
              float4 t4 = tb[c * FP4 + f4];
              acc[4 * f4 + 0] = fmaf(dv, t4.x, acc[4 * f4 + 0]);
              acc[4 * f4 + 1] = fmaf(dv, t4.y, acc[4 * f4 + 1]);
              acc[4 * f4 + 2] = fmaf(dv, t4.z, acc[4 * f4 + 2]);
              acc[4 * f4 + 3] = fmaf(dv, t4.w, acc[4 * f4 + 3]);
            }
          }
        }
      }
      // exact-match channel: alpha * W7[f, C, a, bt] wherever q[i+a-1] == d[j+bt-3] (PAD==PAD counts)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        int ii = i + a - 1;
        if (ii < 0 || ii >= Lq) continue;
        int qi = qids[ii];
#pragma unroll
        for (int bt = 0; bt < 7; ++bt) {
          if (dids[j + bt] == qi) {
            const float* we = wem + (a * 7 + bt) * FPP;
#pragma unroll
            for (int f = 0; f < MT_MAXF; ++f)
              if (f < FPP) acc[f] += we[f];
          }
        }
      }
#pragma unroll
      for (int f = 0; f < MT_MAXF; ++f) acc[f] = fmaxf(acc[f], 0.f);
#pragma unroll
      for (int m = 0; m < MT_MAXM; ++m) {
        if (m < M) {  // uniform branch; mx[] stays statically indexed (registers)
          float z = b1[m];
          const float4* wr = reinterpret_cast<const float4*>(w1 + m * FPP);
#pragma unroll
          for (int f4 = 0; f4 < MT_MAXF / 4; ++f4) {
            if (f4 < FP4) {
              float4 w4 = wr[f4];
              z = fmaf(w4.x, acc[4 * f4 + 0], z);
              z = fmaf(w4.y, acc[4 * f4 + 1], z);
              z = fmaf(w4.z, acc[4 * f4 + 2], z);
              z = fmaf(w4.w, acc[4 * f4 + 3], z);
            }
          }
          if (ARG) {
            if (z > mx[m]) mx[m] = z, ax[ARG ? m : 0] = i * Ld + j;   // strict: keeps this thread's first maximum
          } else {
            mx[m] = fmaxf(mx[m], z);
          }
        }
      }
    }
  }
  if (ARG) {
    // block arg-max: larger value wins, equal values -> smaller cell index
    int* redi = reinterpret_cast<int*>(Tsl);   // the T slice is no longer needed
    __syncthreads();
#pragma unroll
    for (int m = 0; m < MT_MAXM; ++m) {
      if (m < M) {
        float v = mx[m];
        int ix = ax[ARG ? m : 0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
          if (ov > v || (ov == v && oi < ix)) v = ov, ix = oi;
        }
        if ((tid & 31) == 0) red[m * (MT_THREADS / 32) + (tid >> 5)] = v, redi[m * (MT_THREADS / 32) + (tid >> 5)] = ix;
      }
    }
    __syncthreads();
    if (tid < 32) {
      float sv = 0.f;
      if (tid < M) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        for (int w = 0; w < MT_THREADS / 32; ++w) {
          const float v = red[tid * (MT_THREADS / 32) + w];
          const int ix = redi[tid * (MT_THREADS / 32) + w];
          if (v > best || (v == best && ix < bi)) best = v, bi = ix;
        }
        pooled[pl * M + tid] = best;
        argidx[pl * M + tid] = bi;
        sv = best * p.wo[tid];
      }
      sv = warp_sum(sv);
      if (tid == 0) scores[p_glob] = sv + p.wo[M];
    }
    return;
  }
  // block max over j, then score = wo . max + bo
#pragma unroll
  for (int m = 0; m < MT_MAXM; ++m) {
    float v = warp_max(mx[m]);
    if ((tid & 31) == 0) red[m * (MT_THREADS / 32) + (tid >> 5)] = v;
  }
  __syncthreads();
  if (tid < 32) {
    float v = 0.f;
    if (tid < M) {
      float best = -INFINITY;
      for (int w = 0; w < MT_THREADS / 32; ++w) best = fmaxf(best, red[tid * (MT_THREADS / 32) + w]);
      v = best * p.wo[tid];
    }
    v = warp_sum(v);
    if (tid == 0) scores[p_glob] = v + p.wo[M];
  }
}

size_t mt_t_floats(const MtPack& p, int64_t nq, int Lq) { return (size_t)nq * Lq * 7 * p.C * p.FPP; }

int32_t mt_interact(const MtPack& p, const float* cq, const float* cd, float* T, const int64_t* q, const int64_t* d,
                    int N, int Lq, int Ld, int64_t pair_begin, int64_t pair_count, int64_t q_begin, int64_t nq,
                    float* scores, cudaStream_t s) {
  if (pair_count <= 0) return CAIR_OK;
  prof_mark("build_T", s);
  CAIR_LAUNCH(mt_build_t_kernel, dim3(Lq, (unsigned)nq), 256, 0, s, cq, p, Lq, T);
  size_t smem = ((((size_t)p.C * (Ld + 6) + 3) & ~(size_t)3) + 7 * p.C * p.FPP + 21 * p.FPP + p.FPP + p.M * p.FPP + MT_MAXM +
                 MT_MAXM * (MT_THREADS / 32)) * sizeof(float) + (size_t)(Ld + 6 + Lq) * sizeof(int);
  if (smem > 220 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "match_tensor: Ld=%d x C=%d does not fit in shared memory", Ld, p.C);
  CAIR_CUDA(cudaFuncSetAttribute(mt_interact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_mark("interact", s);
  CAIR_LAUNCH(mt_interact_kernel<false>, (unsigned)pair_count, MT_THREADS, smem, s, cd, T, p, q, d, N, Lq, Ld, pair_begin,
              q_begin, scores, (float*)nullptr, (int*)nullptr);
  return CAIR_OK;
}

// training forward (train.cu): all pairs, with the pooled features and arg-max cells
int32_t mt_interact_train(const MtPack& p, const float* cq, const float* cd, float* T, const int64_t* q, const int64_t* d, int N,
                          int Lq, int Ld, int64_t pairs, int64_t nq, float* scores, float* pooled, int* argidx, cudaStream_t s) {
  if (pairs <= 0) return CAIR_OK;
  CAIR_LAUNCH(mt_build_t_kernel, dim3(Lq, (unsigned)nq), 256, 0, s, cq, p, Lq, T);
  size_t smem = ((((size_t)p.C * (Ld + 6) + 3) & ~(size_t)3) + 7 * p.C * p.FPP + 21 * p.FPP + p.FPP + p.M * p.FPP + MT_MAXM +
                 MT_MAXM * (MT_THREADS / 32)) * sizeof(float) + (size_t)(Ld + 6 + Lq) * sizeof(int);
  if (smem > 220 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "match_tensor: Ld=%d x C=%d does not fit in shared memory", Ld, p.C);
  CAIR_CUDA(cudaFuncSetAttribute(mt_interact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(mt_interact_kernel<true>, (unsigned)pairs, MT_THREADS, smem, s, cd, T, p, q, d, N, Lq, Ld, (int64_t)0, (int64_t)0,
              scores, pooled, argidx);
  return CAIR_OK;
}

// weights changed in place (training): rebuild the packed copies in the buffers mt_pack allocated
int32_t mt_repack(const cair_mt_weights& w, MtPack* p, cudaStream_t s) {
  const size_t n7t = (size_t)21 * p->FP * ((p->C + 15) & ~15);
  CAIR_CUDA(cudaMemsetAsync(p->w7t, 0, n7t * sizeof(float), s));
  int total = 21 * (p->C + 1) * p->FPP;
  if (total < p->M * p->FPP) total = p->M * p->FPP;
  CAIR_LAUNCH(mt_pack_kernel, (total + 255) / 256, 256, 0, s, w.conv1.w, w.conv2.w, w.conv3.w, w.conv1.b, w.conv2.b,
              w.conv3.b, w.alpha, w.conv.w, w.conv.b, w.output.w, w.output.b, *p);
  return CAIR_OK;
}


// ---- whole-model orchestration (rankers/mtensor.py:62-131) -----------------------------------------
int32_t mt_create_state(Owned& own, const cair_mt_weights& w, MtState* st, cudaStream_t s) {
  const int dirs = w.bidirectional ? 2 : 1;
  st->V = w.vocab, st->E = w.emsize, st->F = w.featsize, st->Hq = w.nhid_query, st->Hd = w.nhid_doc, st->C = w.nchannels;
  st->rnn_type = w.rnn_type, st->dirs = dirs;
  if (!w.linear_projection.w || !w.linear_projection.b || !w.query_projection.w || !w.query_projection.b ||
      !w.document_projection.w || !w.document_projection.b || !w.alpha || !w.conv1.w || !w.conv2.w || !w.conv3.w ||
      !w.conv1.b || !w.conv2.b || !w.conv3.b || !w.conv.w || !w.conv.b || !w.output.w || !w.output.b)
    return fail(CAIR_ERR_BAD_ARG, "mt_create: null weight pointer");
  // eval-mode fold of embedding + linear_projection (:77-90): folded[v] = table[v] Wp^T + bp; folded[PAD] = bp
  CAIR_CUDA(own.alloc(&st->folded, (size_t)w.vocab * w.featsize));
  CAIR_TRY(gemm_f32(gemm_dense(w.table, w.emsize), w.linear_projection.w, w.linear_projection.b, st->folded,
                    w.featsize, w.vocab, w.featsize, w.emsize, ACT_NONE, s));
  CAIR_TRY(lstm_pack(own, &w.query_fwd, dirs == 2 ? &w.query_rev : nullptr, w.featsize, w.nhid_query / dirs, &st->enc_q, s, w.rnn_type));
  CAIR_TRY(lstm_pack(own, &w.doc_fwd, dirs == 2 ? &w.doc_rev : nullptr, w.featsize, w.nhid_doc / dirs, &st->enc_d, s, w.rnn_type));
  // tcgen05 recurrence (rnn_tc.cu): LSTM and GRU, h <= 128 per direction; round-1 kernel (lstm_tc.cu) kept for A/B runs
  if (rnn_tc_supported(w.featsize, w.nhid_query / dirs))
    CAIR_TRY(rnn_tc_pack(own, &w.query_fwd, dirs == 2 ? &w.query_rev : nullptr, w.featsize, w.nhid_query / dirs, w.rnn_type, &st->rt_q, s));
  if (rnn_tc_supported(w.featsize, w.nhid_doc / dirs))
    CAIR_TRY(rnn_tc_pack(own, &w.doc_fwd, dirs == 2 ? &w.doc_rev : nullptr, w.featsize, w.nhid_doc / dirs, w.rnn_type, &st->rt_d, s));
  const bool lstm = w.rnn_type == CAIR_RNN_LSTM;
  if (lstm && lstm_tc_supported(w.featsize, w.nhid_query / dirs))
    CAIR_TRY(lstm_tc_pack(own, &w.query_fwd, dirs == 2 ? &w.query_rev : nullptr, w.featsize, w.nhid_query / dirs, &st->tc_q, s));
  if (lstm && lstm_tc_supported(w.featsize, w.nhid_doc / dirs))
    CAIR_TRY(lstm_tc_pack(own, &w.doc_fwd, dirs == 2 ? &w.doc_rev : nullptr, w.featsize, w.nhid_doc / dirs, &st->tc_d, s));
  if ((st->rt_q.wimg && st->rt_q.fused) || (st->rt_d.wimg && st->rt_d.fused) || st->tc_q.wimg || st->tc_d.wimg)
    CAIR_TRY(rnn_tc_pack_table(own, st->folded, w.vocab, w.featsize, &st->folded_img, s));   // same row format for both kernels
  CAIR_TRY(dev_copy(own, w.query_projection.w, (size_t)w.nchannels * w.nhid_query, &st->wq, s));
  CAIR_TRY(dev_copy(own, w.query_projection.b, (size_t)w.nchannels, &st->bq, s));
  CAIR_TRY(dev_copy(own, w.document_projection.w, (size_t)w.nchannels * w.nhid_doc, &st->wd, s));
  CAIR_TRY(dev_copy(own, w.document_projection.b, (size_t)w.nchannels, &st->bd, s));
  CAIR_TRY(mt_pack(own, w, &st->pack, s));
  if (mt_tc_proj_supported(w.nchannels, w.nhid_doc)) CAIR_TRY(mt_tc_pack_wd(own, st->wd, w.nchannels, w.nhid_doc, &st->wd_img, s));
  CAIR_TRY(mt_epi_const(st->pack, &st->epi, s));
  CAIR_CUDA(cudaStreamCreateWithFlags(&st->side, cudaStreamNonBlocking));
  CAIR_CUDA(cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming));
  CAIR_CUDA(cudaEventCreateWithFlags(&st->ev_join, cudaEventDisableTiming));
  return CAIR_OK;
}

int32_t mt_add_encoder_layer(Owned& own, MtState* st, int side, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, cudaStream_t s) {
  if (side < 0 || side > 1 || !fwd) return fail(CAIR_ERR_BAD_ARG, "mt_add_encoder_layer: side must be 0 (query) or 1 (document)");
  if ((st->dirs == 2) != (rev != nullptr)) return fail(CAIR_ERR_BAD_ARG, "mt_add_encoder_layer: directions differ from layer 0");
  if (st->nextra[side] >= MtState::MAX_EXTRA) return fail(CAIR_ERR_UNSUPPORTED, "mt_add_encoder_layer: at most 4 stacked layers");
  const int H = side ? st->Hd : st->Hq, h = H / st->dirs, i = st->nextra[side];
  CAIR_TRY(lstm_pack(own, fwd, rev, H, h, &st->xl[side][i], s, st->rnn_type));
  if (rnn_tc_supported(H, h)) CAIR_TRY(rnn_tc_pack(own, fwd, rev, H, h, st->rnn_type, &st->xrt[side][i], s));
  st->nextra[side] = i + 1;
  return CAIR_OK;
}

// Layers 1.. of a stacked encoder over the dense bank of the layer below (the dropout between layers is the identity in
// eval mode; pad rows of a bank are zero and are never read by the packed-sequence recurrence).  Returns the top bank.
static int32_t mt_extra_layers(const MtState& st, int side, float*& bank, float*& other, float* pre, const int64_t* len, int n,
                               int L, int* err, cudaStream_t s, const char* tag) {
  const int H = side ? st.Hd : st.Hq;
  for (int i = 0; i < st.nextra[side]; ++i) {
    const RnnTcPack& rt = st.xrt[side][i];
    if (rt.wimg && st.impl != MT_IMPL_FP32 && g_rnn_impl != RNN_IMPL_FP32)
      CAIR_TRY(rnn_tc_run(rt, gemm_dense(bank, H), len, n, L, other, nullptr, nullptr, pre, err, s, tag));
    else
      CAIR_TRY(lstm_run(st.xl[side][i], gemm_dense(bank, H), len, n, L, other, nullptr, nullptr, pre, err, s, tag));
    float* t = bank;
    bank = other, other = t;
  }
  return CAIR_OK;
}
static size_t mt_extra_ws_floats(const MtState& st, int side, int64_t n, int L) {
  size_t m = 0;
  for (int i = 0; i < st.nextra[side]; ++i) {
    const RnnTcPack& rt = st.xrt[side][i];
    const size_t f = (rt.wimg && st.impl != MT_IMPL_FP32 && g_rnn_impl != RNN_IMPL_FP32) ? rnn_tc_workspace_floats(rt, n, L)
                                                                                         : lstm_workspace_floats(st.xl[side][i], n, L);
    m = f > m ? f : m;
  }
  return m;
}

// engine of one encoder under the process-wide g_rnn_impl switch (see common.cuh)
static bool mt_uses_cluster_kernel(const RnnTcPack& rt, const LstmTcPack& r1) {
  if (!rt.wimg) return false;
  if (g_rnn_impl == RNN_IMPL_CLUSTER) return true;
  return g_rnn_impl == RNN_IMPL_AUTO && !(r1.wimg && rnn_prefers_r1(rt.gru ? CAIR_RNN_GRU : CAIR_RNN_LSTM, rt.in, rt.h));
}
bool mt_doc_uses_cluster_kernel(const MtState& st) { return st.impl != MT_IMPL_FP32 && mt_uses_cluster_kernel(st.rt_d, st.tc_d); }

int32_t mt_forward(const MtState& st, const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen,
                   int B, int N, int Lq, int Ld, int64_t pb, int64_t pc, float* scores, Arena& ws, int* err,
                   cudaStream_t s, bool dry, MtPhase ph) {
  (void)B;
  // queries touched by the pair slice [pb, pb+pc)
  const int64_t qb = pc > 0 ? pb / N : 0;
  const int64_t nq = pc > 0 ? (pb + pc - 1) / N - qb + 1 : 0;
  const bool rt_q = st.impl != MT_IMPL_FP32 && mt_uses_cluster_kernel(st.rt_q, st.tc_q);
  const bool rt_d = st.impl != MT_IMPL_FP32 && mt_uses_cluster_kernel(st.rt_d, st.tc_d);
  const bool tc_q = !rt_q && st.impl != MT_IMPL_FP32 && g_rnn_impl != RNN_IMPL_FP32 && st.tc_q.wimg != nullptr;
  const bool tc_d = !rt_d && st.impl != MT_IMPL_FP32 && g_rnn_impl != RNN_IMPL_FP32 && st.tc_d.wimg != nullptr;
  float* pre_q = ws.take<float>(rt_q ? rnn_tc_workspace_floats(st.rt_q, nq, Lq) : tc_q ? 0 : lstm_workspace_floats(st.enc_q, nq, Lq));
  float* enc_q = ws.take<float>((size_t)nq * Lq * st.Hq);
  float* pre_d = ws.take<float>(rt_d ? rnn_tc_workspace_floats(st.rt_d, pc, Ld) : tc_d ? 0 : lstm_workspace_floats(st.enc_d, pc, Ld));
  float* enc_d = ws.take<float>((size_t)pc * Ld * st.Hd);
  float* cq = ws.take<float>((size_t)nq * Lq * st.C);
  float* cd = ws.take<float>((size_t)pc * Ld * st.C);
  float* enc_q2 = st.nextra[0] ? ws.take<float>((size_t)nq * Lq * st.Hq) : nullptr;
  float* enc_d2 = st.nextra[1] ? ws.take<float>((size_t)pc * Ld * st.Hd) : nullptr;
  float* pre_xq = st.nextra[0] ? ws.take<float>(mt_extra_ws_floats(st, 0, nq, Lq)) : nullptr;
  float* pre_xd = st.nextra[1] ? ws.take<float>(mt_extra_ws_floats(st, 1, pc, Ld)) : nullptr;
  const bool use_tc = st.impl != MT_IMPL_FP32 && mt_tc_supported(st.pack, Lq, Ld);
  float* T = nullptr;
  uint8_t* timg = nullptr;
  uint8_t* aimg = nullptr;
  if (use_tc) {
    size_t timg_bytes, aimg_bytes;
    mt_tc_workspace(st.pack, nq, pc, Lq, Ld, &timg_bytes, &aimg_bytes);
    timg = ws.take<uint8_t>(timg_bytes);
    aimg = ws.take<uint8_t>(aimg_bytes);
  } else {
    T = ws.take<float>(mt_t_floats(st.pack, nq, Lq));
  }
  if (dry || pc <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "match_tensor: workspace too small");
  if (ph.phase != MT_ALL && !(use_tc && st.wd_img && st.impl == MT_IMPL_TC))
    return fail(CAIR_ERR_UNSUPPORTED, "match_tensor: phased forward needs the tcgen05 path");
  if (ph.phase == MT_INTERACT) {
    const int64_t ic = ph.ic < 0 ? pc - ph.ib : ph.ic;
    if (ph.ib < 0 || ic < 0 || ph.ib + ic > pc) return fail(CAIR_ERR_BAD_ARG, "match_tensor: bad interaction sub-range");
    if (ph.join && st.side) CAIR_CUDA(cudaStreamWaitEvent(s, st.ev_join, 0));
    size_t tb, ab;
    mt_tc_workspace(st.pack, nq, pc, Lq, Ld, &tb, &ab);
    return mt_tc_interact(st.pack, st.epi, timg, aimg + (size_t)ph.ib * (ab / (size_t)pc), q, d, N, Lq, Ld, pb + ph.ib, ic,
                          qb, nq, scores, s, ph.max_ctas);
  }
  const int64_t* qs = q + qb * Lq;
  const int64_t* ds = d + pb * Ld;
  // The query side (encoder, channel projection, T operand) depends only on the queries: it runs on the handle's
  // side stream, concurrently with the document encoder (which occupies ~80 of the 148 SMs), and joins before
  // the interaction kernel.  Fork/join with events keeps the caller's stream semantics (and is graph-capturable).
  cudaStream_t sq = st.side ? st.side : s;
  if (st.side) {
    CAIR_CUDA(cudaEventRecord(st.ev_fork, s));
    CAIR_CUDA(cudaStreamWaitEvent(st.side, st.ev_fork, 0));
  }
  // ---- query side: embedding + projection (folded table) -> BiLSTM (:77-94) -> channel projection (:99) ----
  if (rt_q)
    CAIR_TRY(rnn_tc_run(st.rt_q, gemm_gather(st.folded, st.V, st.F, qs, 1, 1, 1, err), qlen + qb, (int)nq, Lq, enc_q, nullptr,
                        nullptr, pre_q, err, sq, st.side ? nullptr : "query_recurrence", st.folded_img));
  else if (tc_q)
    CAIR_TRY(lstm_tc_run(st.tc_q, st.enc_q.bias, gemm_gather(st.folded, st.V, st.F, qs, 1, 1, 1, err), qlen + qb, (int)nq,
                         Lq, enc_q, nullptr, nullptr, err, sq, st.side ? nullptr : "query_recurrence", st.folded_img));
  else
    CAIR_TRY(lstm_run(st.enc_q, gemm_gather(st.folded, st.V, st.F, qs, 1, 1, 1, err), qlen + qb, (int)nq, Lq, enc_q,
                      nullptr, nullptr, pre_q, err, sq, st.side ? nullptr : "query_recurrence"));
  CAIR_TRY(mt_extra_layers(st, 0, enc_q, enc_q2, pre_xq, qlen + qb, (int)nq, Lq, err, sq, st.side ? nullptr : "query_recurrence_upper"));
  if (st.dbg_enc_q)
    CAIR_CUDA(cudaMemcpyAsync(st.dbg_enc_q + (size_t)qb * Lq * st.Hq, enc_q, (size_t)nq * Lq * st.Hq * sizeof(float),
                              cudaMemcpyDeviceToDevice, sq));
  CAIR_TRY(gemm_f32(gemm_dense(enc_q, st.Hq), st.wq, st.bq, cq, st.C, nq * Lq, st.C, st.Hq, ACT_NONE, sq));
  if (use_tc) CAIR_TRY(mt_tc_build_t(st.pack, cq, timg, Lq, nq, sq));
  if (st.side) CAIR_CUDA(cudaEventRecord(st.ev_join, st.side));
  // ---- document side ----
  if (rt_d)
    CAIR_TRY(rnn_tc_run(st.rt_d, gemm_gather(st.folded, st.V, st.F, ds, 1, 1, 1, err), dlen + pb, (int)pc, Ld, enc_d, nullptr,
                        nullptr, pre_d, err, s, "doc_recurrence", st.folded_img, ph.doc_min_spc));
  else if (tc_d)
    CAIR_TRY(lstm_tc_run(st.tc_d, st.enc_d.bias, gemm_gather(st.folded, st.V, st.F, ds, 1, 1, 1, err), dlen + pb, (int)pc,
                         Ld, enc_d, nullptr, nullptr, err, s, "doc_recurrence", st.folded_img, ph.doc_min_spc));
  else
    CAIR_TRY(lstm_run(st.enc_d, gemm_gather(st.folded, st.V, st.F, ds, 1, 1, 1, err), dlen + pb, (int)pc, Ld, enc_d,
                      nullptr, nullptr, pre_d, err, s, "doc_recurrence"));
  CAIR_TRY(mt_extra_layers(st, 1, enc_d, enc_d2, pre_xd, dlen + pb, (int)pc, Ld, err, s, "doc_recurrence_upper"));
  if (st.dbg_enc_d)
    CAIR_CUDA(cudaMemcpyAsync(st.dbg_enc_d + (size_t)pb * Ld * st.Hd, enc_d, (size_t)pc * Ld * st.Hd * sizeof(float),
                              cudaMemcpyDeviceToDevice, s));
  // channel projection (:108): the bias also lands on pad positions (zero memory-bank rows)
  prof_mark("doc_projection", s);
  if (use_tc) {
    if (st.wd_img && st.impl == MT_IMPL_TC) {
      // tcgen05 projection written straight into the interaction kernel's operand image
      CAIR_TRY(mt_tc_proj_image(st.pack, enc_d, st.Hd, st.wd_img, st.bd, aimg, Ld, pc, s));
      if (ph.phase == MT_ENCODE) return CAIR_OK;   // the interaction phase joins the query side
    } else {
      CAIR_TRY(gemm_f32(gemm_dense(enc_d, st.Hd), st.wd, st.bd, cd, st.C, pc * Ld, st.C, st.Hd, ACT_NONE, s));
      CAIR_TRY(mt_tc_doc_image(st.pack, cd, aimg, Ld, pc, s));
    }
    prof_mark("join_query_side", s);
    if (st.side) CAIR_CUDA(cudaStreamWaitEvent(s, st.ev_join, 0));
    return mt_tc_interact(st.pack, st.epi, timg, aimg, q, d, N, Lq, Ld, pb, pc, qb, nq, scores, s);
  }
  CAIR_TRY(gemm_f32(gemm_dense(enc_d, st.Hd), st.wd, st.bd, cd, st.C, pc * Ld, st.C, st.Hd, ACT_NONE, s));
  if (st.side) CAIR_CUDA(cudaStreamWaitEvent(s, st.ev_join, 0));
  return mt_interact(st.pack, cq, cd, T, q, d, N, Lq, Ld, pb, pc, qb, nq, scores, s);
}

bool mt_can_pipeline(const MtState& st, int Lq, int Ld) {
  return st.impl == MT_IMPL_TC && st.wd_img != nullptr && st.side != nullptr && mt_tc_supported(st.pack, Lq, Ld);
}

}  // namespace cair
