// Training step of Match-Tensor: train-mode forward with saved activations + hand-written
// backward, so that the reference's Ranker.update (neuroir/models/ranker.py:192-230: forward, criterion, loss.backward(),
// clip_grad_norm, optimizer.step()) runs on libcair kernels.  The loss, the clipping and the optimizer stay in the
// unchanged wrapper (they are torch calls on [B,N] scores / on the parameter list); what is here is everything between the
// token ids and the scores, forward and backward:
//   ids -> embedding rows * dropout mask (mtensor.py:77-84) -> linear_projection (:88-90) -> (Bi)LSTM encoders (:93-94,
//   rnn_encoder.py:62-141, packed-sequence semantics) -> channel projections (:99,:108) -> match tensor + exact-match
//   channel (:113-120) -> conv1/2/3 + ReLU + 1x1 conv + two max-pools + Linear (:123-131).
// fp32 throughout (CUDA cores): the backward is a "next" row, correctness first.  Algebra:
//  * the two max-pools route the gradient of a pair to at most match_filter_size cells (i*, j*) of the [Lq, Ld] plane
//    (first maximum in row-major order, as ATen's max reduction breaks ties), so the backward of the whole interaction
//    stack is SPARSE: per (pair, m) the conv outputs of that one cell are recomputed from cq / cd / ids and the
//    gradient is scattered to the 3 x 7 x (C+1) taps around it.  The forward only has to remember the arg-max cells.
//  * BPTT: a reverse-time recurrence kernel turns d(bank) into the pre-activation gate gradients in place of the saved
//    gate activations; dW_hh, dW_ih, db, dx are then plain GEMMs over all (sequence, step) rows.
//  * dropout masks are a counter-based hash of (seed, element index), recomputed in the backward (never stored).
#include "models.cuh"

namespace cair {

// ------------------------------------------------------------------------------------------------
// dropout: keep-scale of element idx (0 or 1/(1-p)); splitmix64 of (seed, idx)
__device__ __forceinline__ float drop_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  if (p <= 0.f) return 1.f;
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  return u >= p ? inv_keep : 0.f;
}

__global__ void dropout_mask_kernel(uint64_t seed, float p, int64_t n, float* __restrict__ out) {
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = drop_scale(seed, (uint64_t)i, p, inv);
}

// x[r, :] = table[ids[r], :] * mask(row0 + r, :)
__global__ void embed_drop_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids, int V, int E, int64_t rows,
                                  int64_t row0, float p, uint64_t seed, float* __restrict__ out, int* err) {
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  const int64_t total = rows * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int e = (int)(i - r * E);
    const int64_t id = checked_id(ids[r], V, err);
    out[i] = table[id * E + e] * drop_scale(seed, (uint64_t)((row0 + r) * E + e), p, inv);
  }
}

__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// out[c * ldo + r] = in[r * cols + c]  (small weight matrices)
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out, int64_t ldo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * cols) {
    const int r = i / cols, c = i - r * cols;
    out[(int64_t)c * ldo + r] = in[i];
  }
}

// ------------------------------------------------------------------------------------------------
// C[M,N] += sum_r A[r, :M]^T B[r', :N]; r' = r + bshift inside the same block of L rows (zero outside): the weight
// gradients.  64x64 tile per CTA over a chunk of rows, atomicAdd into C.
constexpr int TN_T = 64, TN_K = 16;
__global__ void __launch_bounds__(256) gemm_tn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                      int64_t ldb, int bshift, int L, float* __restrict__ C, int64_t ldc,
                                                      int64_t R, int M, int N, int64_t rows_per_cta) {
  __shared__ __align__(16) float As[TN_K][TN_T + 4];
  __shared__ __align__(16) float Bs[TN_K][TN_T + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TN_T, n0 = blockIdx.y * TN_T;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < R ? r_begin + rows_per_cta : R;
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += TN_K) {
    const int64_t r = r0 + lr;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < r_end) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (m0 + lc + u < M) av[u] = A[r * lda + m0 + lc + u];
      bool ok = true;
      int64_t rb = r;
      if (bshift != 0) {
        const int t = (int)(r % L) + bshift;
        ok = t >= 0 && t < L;
        rb = r + bshift;
      }
      if (ok) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (n0 + lc + u < N) bv[u] = B[rb * ldb + n0 + lc + u];
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      As[lr][lc + u] = av[u];
      Bs[lr][lc + u] = bv[u];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TN_K; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N && acc[i][j] != 0.f) atomicAdd(&C[(int64_t)m * ldc + n], acc[i][j]);
    }
  }
}

static int32_t gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int bshift, int L, float* C, int64_t ldc,
                       int64_t R, int M, int N, cudaStream_t s) {
  if (R <= 0 || M <= 0 || N <= 0) return CAIR_OK;
  const int mt = (M + TN_T - 1) / TN_T, nt = (N + TN_T - 1) / TN_T;
  int64_t z = (4 * kSMs + mt * nt - 1) / (mt * nt);
  int64_t rpc = (R + z - 1) / z;
  rpc = (rpc + TN_K - 1) / TN_K * TN_K;
  if (rpc < 256) rpc = 256;
  z = (R + rpc - 1) / rpc;
  CAIR_LAUNCH(gemm_tn_kernel, dim3(mt, nt, (unsigned)z), 256, 0, s, A, lda, B, ldb, bshift, L, C, ldc, R, M, N, rpc);
  return CAIR_OK;
}

// out[c] += sum_r A[r, c]  (bias gradients); out2 (optional) receives the same sums (b_ih and b_hh share a gradient)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ A, int64_t lda, int64_t R, int N,
                                                     int64_t rows_per_cta, float* __restrict__ out, float* __restrict__ out2) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < R ? r_begin + rows_per_cta : R;
  float acc = 0.f;
  if (c < N)
    for (int64_t r = r_begin + ry; r < r_end; r += 8) acc += A[r * lda + c];
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < N) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][cx];
    if (v != 0.f) {
      atomicAdd(&out[c], v);
      if (out2) atomicAdd(&out2[c], v);
    }
  }
}
static int32_t colsum(const float* A, int64_t lda, int64_t R, int N, float* out, float* out2, cudaStream_t s) {
  if (R <= 0 || N <= 0 || !out) return CAIR_OK;
  const int ct = (N + 31) / 32;
  int64_t y = (2 * kSMs + ct - 1) / ct;
  int64_t rpc = (R + y - 1) / y;
  if (rpc < 64) rpc = 64;
  y = (R + rpc - 1) / rpc;
  CAIR_LAUNCH(colsum_kernel, dim3(ct, (unsigned)y), 256, 0, s, A, lda, R, N, rpc, out, out2);
  return CAIR_OK;
}

// ------------------------------------------------------------------------------------------------
// LSTM forward for training: as rnn_rec_kernel (lstm.cu), and the gate ACTIVATIONS (i, f, g, o) replace the pre-gates in
// `gates` [n*L, dirs*4h]; c_seq / out [n*L, dirs*h].  TS sequences of one direction per CTA.
constexpr int TR_TS = 8;
template <bool WSMEM>
__global__ void __launch_bounds__(256) lstm_train_fwd_kernel(float* __restrict__ gates, const float* __restrict__ w_hh_t,
                                                             const int64_t* __restrict__ len, int n, int L, int h, int dirs,
                                                             float* __restrict__ out, float* __restrict__ c_seq, int* err) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TS = TR_TS;
  const int G = 4 * h, hp = (h + 3) & ~3, PG = dirs * G, Hout = dirs * h;
  const int dir = blockIdx.y, s0 = blockIdx.x * TS, tid = threadIdx.x;
  float* wsm = smem;
  float* hprev = smem + (WSMEM ? (size_t)h * G : 0);
  float* cst = hprev + TS * hp;
  float* gs = cst + TS * h;
  __shared__ int slen[TS];
  __shared__ int smaxlen;
  const float* wt = w_hh_t + (size_t)dir * h * G;
  if (WSMEM)
    for (int i = tid; i < h * G; i += 256) wsm[i] = wt[i];
  const float* W = WSMEM ? wsm : wt;
  for (int i = tid; i < TS * hp; i += 256) hprev[i] = 0.f;
  for (int i = tid; i < TS * h; i += 256) cst[i] = 0.f;
  if (tid < TS) {
    int s = s0 + tid, l = 0;
    if (s < n) {
      int64_t ll = len[s];
      if (ll < 1 || ll > L) {
        atomicOr(err, ERRF_BAD_LENGTH);
        ll = ll < 1 ? 1 : L;
      }
      l = (int)ll;
    }
    slen[tid] = l;
  }
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int s = 0; s < TS; ++s) m = max(m, slen[s]);
    smaxlen = m;
  }
  for (int s = 0; s < TS; ++s) {   // pad rows of the memory bank and of c_seq: zeros
    if (s0 + s >= n) break;
    const int npad = (L - slen[s]) * h;
    const size_t base = ((size_t)(s0 + s) * L + slen[s]) * Hout + dir * h;
    for (int i = tid; i < npad; i += 256) {
      const size_t o = base + (size_t)(i / h) * Hout + (i % h);
      out[o] = 0.f;
      c_seq[o] = 0.f;
    }
  }
  __syncthreads();
  const int maxlen = smaxlen;
  for (int step = 0; step < maxlen; ++step) {
    for (int r = tid; r < G; r += 256) {
      float acc[TS];
#pragma unroll
      for (int s = 0; s < TS; ++s) {
        const int l = slen[s];
        acc[s] = step < l ? gates[((size_t)(s0 + s) * L + (dir ? l - 1 - step : step)) * PG + dir * G + r] : 0.f;
      }
      int k = 0;
#pragma unroll 2
      for (; k + 4 <= h; k += 4) {   // four weights + one 128-bit state load per sequence per 4 k: 12 loads per 32 FMAs
        const float w0 = W[(size_t)(k + 0) * G + r], w1 = W[(size_t)(k + 1) * G + r];
        const float w2 = W[(size_t)(k + 2) * G + r], w3 = W[(size_t)(k + 3) * G + r];
#pragma unroll
        for (int s = 0; s < TS; ++s) {
          const float4 hv = *reinterpret_cast<const float4*>(&hprev[s * hp + k]);
          acc[s] = fmaf(w0, hv.x, acc[s]);
          acc[s] = fmaf(w1, hv.y, acc[s]);
          acc[s] = fmaf(w2, hv.z, acc[s]);
          acc[s] = fmaf(w3, hv.w, acc[s]);
        }
      }
      for (; k < h; ++k) {
        const float w0 = W[(size_t)k * G + r];
#pragma unroll
        for (int s = 0; s < TS; ++s) acc[s] = fmaf(w0, hprev[s * hp + k], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < TS; ++s) gs[s * G + r] = acc[s];
    }
    __syncthreads();
    for (int i = tid; i < TS * h; i += 256) {
      const int s = i / h, u = i - s * h, l = slen[s];
      if (step < l) {
        const float* g = gs + s * G;
        const float ig = 1.f / (1.f + expf(-g[u])), fg = 1.f / (1.f + expf(-g[h + u]));
        const float gg = tanhf(g[2 * h + u]), og = 1.f / (1.f + expf(-g[3 * h + u]));
        const float c = fg * cst[i] + ig * gg;
        const float hv = og * tanhf(c);
        cst[i] = c;
        hprev[s * hp + u] = hv;
        const size_t row = (size_t)(s0 + s) * L + (dir ? l - 1 - step : step);
        float* gp = gates + row * PG + dir * G;
        gp[u] = ig, gp[h + u] = fg, gp[2 * h + u] = gg, gp[3 * h + u] = og;
        out[row * Hout + dir * h + u] = hv;
        c_seq[row * Hout + dir * h + u] = c;
      }
    }
    __syncthreads();
  }
}

// BPTT: `gates` holds the activations on entry and the pre-activation gradients dL/da (zero at t >= len) on exit.
// denc [n*L, dirs*h]: gradient of the memory bank.  w_hh [dirs][4h][h] (torch layout).
template <bool WSMEM, int TS>
__global__ void __launch_bounds__(256) lstm_train_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_seq,
                                                             const float* __restrict__ denc, const float* __restrict__ w_hh,
                                                             const int64_t* __restrict__ len, int n, int L, int h, int dirs) {
  extern __shared__ __align__(16) float smem[];
  const int G = 4 * h, PG = dirs * G, Hout = dirs * h;
  const int dir = blockIdx.y, s0 = blockIdx.x * TS, tid = threadIdx.x;
  float* wsm = smem;                                         // [G][h]
  float* da = smem + (WSMEM ? (size_t)G * h : 0);            // [G][TS]
  float* dh = da + (size_t)G * TS;                           // [TS][h]
  float* dc = dh + TS * h;                                   // [TS][h]
  float* part = dc + TS * h;                                 // [4][TS][h] partial sums of the W_hh^T product
  __shared__ int slen[TS];
  __shared__ int smaxlen;
  const float* wg = w_hh + (size_t)dir * G * h;
  if (WSMEM)
    for (int i = tid; i < G * h; i += 256) wsm[i] = wg[i];
  const float* W = WSMEM ? wsm : wg;
  for (int i = tid; i < TS * h; i += 256) dh[i] = 0.f, dc[i] = 0.f;
  if (tid < TS) {
    int s = s0 + tid, l = 0;
    if (s < n) {
      int64_t ll = len[s];
      ll = ll < 1 ? 1 : (ll > L ? L : ll);
      l = (int)ll;
    }
    slen[tid] = l;
  }
  __syncthreads();
  if (tid == 0) {
    int m = 0;
    for (int s = 0; s < TS; ++s) m = max(m, slen[s]);
    smaxlen = m;
  }
  for (int s = 0; s < TS; ++s) {   // pad rows: zero gate gradients (the weight-gradient GEMMs sum over all rows)
    if (s0 + s >= n) break;
    const int npad = (L - slen[s]) * G;
    float* base = gates + ((size_t)(s0 + s) * L + slen[s]) * PG + dir * G;
    for (int i = tid; i < npad; i += 256) base[(size_t)(i / G) * PG + (i % G)] = 0.f;
  }
  __syncthreads();
  const int maxlen = smaxlen;
  for (int step = maxlen - 1; step >= 0; --step) {
    for (int i = tid; i < TS * h; i += 256) {
      const int s = i / h, u = i - s * h, l = slen[s];
      float ai = 0.f, af = 0.f, ag = 0.f, ao = 0.f;
      if (step < l) {
        const int t = dir ? l - 1 - step : step;
        const size_t row = (size_t)(s0 + s) * L + t;
        float* gp = gates + row * PG + dir * G;
        const float gi = gp[u], gf = gp[h + u], gg = gp[2 * h + u], go = gp[3 * h + u];
        const float c = c_seq[row * Hout + dir * h + u];
        const float cprev = step > 0 ? c_seq[(dir ? row + 1 : row - 1) * Hout + dir * h + u] : 0.f;
        const float dht = denc[row * Hout + dir * h + u] + dh[i];
        const float tc = tanhf(c);
        const float dct = dc[i] + dht * go * (1.f - tc * tc);
        ai = dct * gg * gi * (1.f - gi);
        af = dct * cprev * gf * (1.f - gf);
        ag = dct * gi * (1.f - gg * gg);
        ao = dht * tc * go * (1.f - go);
        dc[i] = dct * gf;
        gp[u] = ai, gp[h + u] = af, gp[2 * h + u] = ag, gp[3 * h + u] = ao;
      }
      da[(size_t)u * TS + s] = ai;
      da[(size_t)(h + u) * TS + s] = af;
      da[(size_t)(2 * h + u) * TS + s] = ag;
      da[(size_t)(3 * h + u) * TS + s] = ao;
    }
    __syncthreads();
    // dh[s][k] = sum_r da[r][s] W[r][k]: thread (k, group grp of gate rows) for all TS sequences
    for (int idx = tid; idx < 4 * h; idx += 256) {
      const int grp = idx / h, k = idx - grp * h;
      float acc[TS];
#pragma unroll
      for (int s = 0; s < TS; ++s) acc[s] = 0.f;
      const int r_lo = grp * h, r_hi = r_lo + h;   // one gate type per thread group
      for (int r = r_lo; r < r_hi; ++r) {
        const float w0 = W[(size_t)r * h + k];
#pragma unroll
        for (int s4 = 0; s4 < TS / 4; ++s4) {
          const float4 d0 = *reinterpret_cast<const float4*>(&da[(size_t)r * TS + 4 * s4]);
          acc[4 * s4] = fmaf(w0, d0.x, acc[4 * s4]), acc[4 * s4 + 1] = fmaf(w0, d0.y, acc[4 * s4 + 1]);
          acc[4 * s4 + 2] = fmaf(w0, d0.z, acc[4 * s4 + 2]), acc[4 * s4 + 3] = fmaf(w0, d0.w, acc[4 * s4 + 3]);
        }
      }
#pragma unroll
      for (int s = 0; s < TS; ++s) part[((size_t)grp * TS + s) * h + k] = acc[s];
    }
    __syncthreads();
    for (int i = tid; i < TS * h; i += 256)
      dh[i] = part[i] + part[(size_t)TS * h + i] + part[(size_t)2 * TS * h + i] + part[(size_t)3 * TS * h + i];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// merged stencil for the sparse interaction backward: W7[(a*7+bt)*C1 + c][FPP] (f fastest, raw weights, no alpha)
__global__ void mt_train_pack_kernel(const float* __restrict__ c1, const float* __restrict__ c2, const float* __restrict__ c3,
                                     int C1, int nf, int FP, int FPP, float* __restrict__ w7) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 21 * C1 * FPP) return;
  const int f = idx % FPP, c = (idx / FPP) % C1, bt = (idx / (FPP * C1)) % 7, a = idx / (FPP * C1 * 7);
  float v = 0.f;
  if (f < FP) {
    const int k = f / nf, ff = f - k * nf, kw = 3 + 2 * k, bb = bt - (2 - k);
    const float* w = k == 0 ? c1 : (k == 1 ? c2 : c3);
    if (bb >= 0 && bb < kw) v = w[(((size_t)ff * C1 + c) * 3 + a) * kw + bb];
  }
  w7[idx] = v;
}

struct MtGradPtrs {
  float *conv1_w, *conv2_w, *conv3_w, *conv1_b, *conv2_b, *conv3_b, *conv_w, *conv_b, *out_w, *out_b, *alpha;
};
constexpr int TRB_MAXT = 6;    // taps per thread: ceil(21 * 65 / 256)
constexpr int TRB_MAXF = 24;

// smem: dW7 [21*C1*FPP] | y [FPP] | red [8][FPP] | dw1 [M*FP] | db1 [M] | dwo [M] | dbias [FP] | misc [4] | qids [Lq] | dids [Ld]
__global__ void __launch_bounds__(256) mt_train_interact_bwd_kernel(
    const float* __restrict__ cq, const float* __restrict__ cd, const int64_t* __restrict__ q, const int64_t* __restrict__ d,
    const float* __restrict__ dscores, const float* __restrict__ pooled, const int* __restrict__ argidx,
    const float* __restrict__ w7, const float* __restrict__ cbias1, const float* __restrict__ cbias2,
    const float* __restrict__ cbias3, const float* __restrict__ w1, const float* __restrict__ wo,
    const float* __restrict__ alpha_p, int N, int Lq, int Ld, int C, int nf, int M, int64_t pairs, float* __restrict__ dcq,
    float* __restrict__ dcd, MtGradPtrs g, int w7_smem) {
  extern __shared__ __align__(16) float sm[];
  const int C1 = C + 1, FP = 3 * nf, FPP = (FP + 3) & ~3, NT = 21 * C1;
  float* dW7 = sm;
  float* w7s = dW7 + (size_t)NT * FPP;                                  // the stencil itself, when it fits beside its gradient
  float* ysm = w7s + (w7_smem ? (size_t)NT * FPP : 0);
  float* red = ysm + FPP;
  float* dw1 = red + 8 * FPP;
  float* db1 = dw1 + M * FP;
  float* dwo = db1 + M;
  float* dbias = dwo + M;
  float* misc = dbias + FP;     // [0] d alpha, [1] d output bias
  int* qids = reinterpret_cast<int*>(misc + 4);
  int* dids = qids + Lq;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float alpha = alpha_p[0];
  for (int i = tid; i < NT * FPP; i += 256) dW7[i] = 0.f;
  for (int i = tid; i < FPP + 8 * FPP + M * FP + 2 * M + FP + 4; i += 256) ysm[i] = 0.f;
  if (w7_smem) {
    for (int i = tid; i < NT * FPP; i += 256) w7s[i] = w7[i];
    w7 = w7s;   // every (pair, m) unit reads all 21 (C+1) FPP weights twice: from shared memory instead of L1 / L2
  }
  __syncthreads();
  float dalpha_loc = 0.f;
  for (int64_t p = blockIdx.x; p < pairs; p += gridDim.x) {
    const int64_t b = p / N;
    __syncthreads();
    for (int i = tid; i < Lq; i += 256) qids[i] = (int)q[b * Lq + i];
    for (int i = tid; i < Ld; i += 256) dids[i] = (int)d[p * Ld + i];
    const float ds = dscores[p];
    if (tid < M) dwo[tid] += ds * pooled[p * M + tid];
    if (tid == 0) misc[1] += ds;
    __syncthreads();
    for (int m = 0; m < M; ++m) {
      const int cell = argidx[p * M + m];
      const int is = cell / Ld, js = cell - is * Ld;
      const float dz = ds * wo[m];
      float mtv[TRB_MAXT], cqv[TRB_MAXT], cdv[TRB_MAXT];
      float acc[TRB_MAXF];
#pragma unroll
      for (int f = 0; f < TRB_MAXF; ++f) acc[f] = 0.f;
#pragma unroll
      for (int k = 0; k < TRB_MAXT; ++k) {
        const int tap = tid + k * 256;
        mtv[k] = 0.f, cqv[k] = 0.f, cdv[k] = 0.f;
        if (tap < NT) {
          const int c = tap % C1, ab = tap / C1, a = ab / 7, bt = ab - a * 7;
          const int ii = is + a - 1, jj = js + bt - 3;
          if (ii >= 0 && ii < Lq && jj >= 0 && jj < Ld) {
            if (c < C) {
              cqv[k] = cq[((size_t)b * Lq + ii) * C + c];
              cdv[k] = cd[((size_t)p * Ld + jj) * C + c];
              mtv[k] = cqv[k] * cdv[k];
            } else {
              cqv[k] = qids[ii] == dids[jj] ? 1.f : 0.f;   // the match indicator
              mtv[k] = alpha * cqv[k];
            }
            const float* wr = w7 + (size_t)tap * FPP;
#pragma unroll
            for (int f = 0; f < TRB_MAXF; ++f)
              if (f < FPP) acc[f] = fmaf(wr[f], mtv[k], acc[f]);
          }
        }
      }
#pragma unroll
      for (int f = 0; f < TRB_MAXF; ++f) {
        if (f < FPP) {
          const float v = warp_sum(acc[f]);
          if (lane == 0) red[warp * FPP + f] = v;
        }
      }
      __syncthreads();
      if (tid < FP) {
        const int k = tid / nf, ff = tid - k * nf;
        float y = (k == 0 ? cbias1 : (k == 1 ? cbias2 : cbias3))[ff];
#pragma unroll
        for (int w = 0; w < 8; ++w) y += red[w * FPP + tid];
        const float g1 = y > 0.f ? w1[m * FP + tid] * dz : 0.f;
        ysm[tid] = g1;
        dw1[m * FP + tid] += dz * fmaxf(y, 0.f);
        dbias[tid] += g1;
      } else if (tid >= FP && tid < FPP) {
        ysm[tid] = 0.f;
      }
      if (tid == 255) db1[m] += dz;
      __syncthreads();
      float g1[TRB_MAXF];
#pragma unroll
      for (int f = 0; f < TRB_MAXF; ++f) g1[f] = f < FPP ? ysm[f] : 0.f;
#pragma unroll
      for (int k = 0; k < TRB_MAXT; ++k) {
        const int tap = tid + k * 256;
        if (tap < NT && (mtv[k] != 0.f || cqv[k] != 0.f || cdv[k] != 0.f)) {
          const int c = tap % C1, ab = tap / C1, a = ab / 7, bt = ab - a * 7;
          const int ii = is + a - 1, jj = js + bt - 3;
          const float* wr = w7 + (size_t)tap * FPP;
          float* dwr = dW7 + (size_t)tap * FPP;
          float dmt = 0.f;
#pragma unroll
          for (int f = 0; f < TRB_MAXF; ++f) {
            if (f < FPP) {
              dwr[f] = fmaf(g1[f], mtv[k], dwr[f]);
              dmt = fmaf(wr[f], g1[f], dmt);
            }
          }
          if (c < C) {
            if (dmt != 0.f) {
              atomicAdd(&dcq[((size_t)b * Lq + ii) * C + c], dmt * cdv[k]);
              atomicAdd(&dcd[((size_t)p * Ld + jj) * C + c], dmt * cqv[k]);
            }
          } else {
            dalpha_loc += dmt * cqv[k];
          }
        }
      }
    }
  }
  // flush the CTA's accumulators
  dalpha_loc = warp_sum(dalpha_loc);
  if (lane == 0 && dalpha_loc != 0.f) atomicAdd(&misc[0], dalpha_loc);
  __syncthreads();
  for (int i = tid; i < NT * FPP; i += 256) {
    const float v = dW7[i];
    if (v == 0.f) continue;
    const int f = i % FPP, tap = i / FPP;
    if (f >= FP) continue;
    const int c = tap % C1, ab = tap / C1, a = ab / 7, bt = ab - a * 7;
    const int k = f / nf, ff = f - k * nf, kw = 3 + 2 * k, bb = bt - (2 - k);
    if (bb < 0 || bb >= kw) continue;
    float* gw = k == 0 ? g.conv1_w : (k == 1 ? g.conv2_w : g.conv3_w);
    atomicAdd(&gw[(((size_t)ff * C1 + c) * 3 + a) * kw + bb], v);
  }
  for (int i = tid; i < M * FP; i += 256)
    if (dw1[i] != 0.f) atomicAdd(&g.conv_w[i], dw1[i]);
  for (int i = tid; i < M; i += 256) {
    atomicAdd(&g.conv_b[i], db1[i]);
    atomicAdd(&g.out_w[i], dwo[i]);
  }
  for (int i = tid; i < FP; i += 256) {
    const int k = i / nf, ff = i - k * nf;
    atomicAdd(&(k == 0 ? g.conv1_b : (k == 1 ? g.conv2_b : g.conv3_b))[ff], dbias[i]);
  }
  if (tid == 0) {
    atomicAdd(g.alpha, misc[0]);
    atomicAdd(g.out_b, misc[1]);
  }
}

// d table[id[r], e] += mask(r, e) * sum_f df[r, f] Wp[f, e]   (PAD rows receive no gradient: nn.Embedding padding_idx)
__global__ void __launch_bounds__(256) embed_grad_kernel(const float* __restrict__ df, const float* __restrict__ wp,
                                                         const int64_t* __restrict__ ids, int V, int E, int F, int64_t rows,
                                                         int64_t row0, float p, uint64_t seed, float* __restrict__ dtable) {
  extern __shared__ __align__(16) float sm[];
  float* dfr = sm;   // [8][F]
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  const int tid = threadIdx.x;
  for (int64_t r0 = (int64_t)blockIdx.x * 8; r0 < rows; r0 += (int64_t)gridDim.x * 8) {
    __syncthreads();
    for (int i = tid; i < 8 * F; i += 256) {
      const int64_t r = r0 + i / F;
      dfr[i] = r < rows ? df[r * F + (i % F)] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < E; e += 256) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      for (int f = 0; f < F; ++f) {
        const float w = wp[(size_t)f * E + e];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, dfr[k * F + f], acc[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int64_t r = r0 + k;
        if (r >= rows) break;
        const int64_t id = ids[r];
        if (id <= 0 || id >= V) continue;   // PAD (0) and invalid ids
        const float v = acc[k] * drop_scale(seed, (uint64_t)((row0 + r) * E + e), p, inv);
        if (v != 0.f) atomicAdd(&dtable[id * E + e], v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct MtTrainer {
  int device = 0;
  cair_mt_weights w{};    // LIVE parameter pointers (read at every step)
  Owned own;
  MtPack pack{};          // forward interaction weights (re-packed every step)
  float* w7 = nullptr;    // merged stencil for the backward
  float *whht_q = nullptr, *whht_d = nullptr;   // [dirs][h][4h]
  float *bias_q = nullptr, *bias_d = nullptr;   // [dirs][4h] b_ih + b_hh
  float *wiht_q = nullptr, *wiht_d = nullptr;   // [F][dirs*4h]
  float *wqt = nullptr, *wdt = nullptr;         // [Hq][C], [Hd][C]
  int dirs = 1, hq = 0, hd = 0;
  int tc_forward = 1;     // forward interaction on the tcgen05 kernel (with arg-max) where the shape allows; 0: fp32 kernel
  // bf16x3 operand images of the dense layers for the tcgen05 GEMM (re-packed from the live parameters every step; used when
  // tc_forward is on and the operand layout allows, else the fp32 GEMM): linear_projection, w_ih per encoder and direction,
  // the channel projections, and w_ih^T (gate gradients -> d featsize rows in the backward)
  GemmTcW tw_proj{}, tw_ih_q[2]{}, tw_ih_d[2]{}, tw_cq{}, tw_cd{}, tw_wiht_q{}, tw_wiht_d{};
  // weight images of the tcgen05 recurrence (lstm_tc.cu, gate-saving instantiation) where the encoder fits it (in < 48,
  // h <= 64 per direction): the forward recurrence then runs fused with its input projection, no pre-gate GEMM
  LstmTcPack tcp_q{}, tcp_d{};
};

struct MtTrainWs {
  int* err;
  uint8_t *timg, *aimg;   // operand images of the tensor-core interaction kernel (forward)
  float *xq, *xd, *fq, *fd, *gq, *gd, *cseq_q, *cseq_d, *enc_q, *enc_d, *cq, *cd, *T, *pooled;
  int* argidx;
  float *dcq, *dcd, *denc_q, *denc_d, *dfq, *dfd;
};

static void mt_train_layout(const MtTrainer& t, Arena& a, int B, int N, int Lq, int Ld, MtTrainWs* o) {
  const size_t Rq = (size_t)B * Lq, Rd = (size_t)B * N * Ld, P = (size_t)B * N;
  const cair_mt_weights& w = t.w;
  o->err = a.take<int>(64);
  o->xq = a.take<float>(Rq * w.emsize), o->xd = a.take<float>(Rd * w.emsize);
  o->fq = a.take<float>(Rq * w.featsize), o->fd = a.take<float>(Rd * w.featsize);
  o->gq = a.take<float>(Rq * t.dirs * 4 * t.hq), o->gd = a.take<float>(Rd * t.dirs * 4 * t.hd);
  o->cseq_q = a.take<float>(Rq * w.nhid_query), o->cseq_d = a.take<float>(Rd * w.nhid_doc);
  o->enc_q = a.take<float>(Rq * w.nhid_query), o->enc_d = a.take<float>(Rd * w.nhid_doc);
  o->cq = a.take<float>(Rq * w.nchannels), o->cd = a.take<float>(Rd * w.nchannels);
  o->T = a.take<float>(mt_t_floats(t.pack, B, Lq));
  o->timg = o->aimg = nullptr;
  if (mt_tc_supported(t.pack, Lq, Ld)) {
    size_t tb = 0, ab = 0;
    mt_tc_workspace(t.pack, B, (int64_t)P, Lq, Ld, &tb, &ab);
    o->timg = a.take<uint8_t>(tb), o->aimg = a.take<uint8_t>(ab);
  }
  o->pooled = a.take<float>(P * w.match_filter_size);
  o->argidx = a.take<int>(P * w.match_filter_size);
  o->dcq = a.take<float>(Rq * w.nchannels), o->dcd = a.take<float>(Rd * w.nchannels);
  o->denc_q = a.take<float>(Rq * w.nhid_query), o->denc_d = a.take<float>(Rd * w.nhid_doc);
  o->dfq = a.take<float>(Rq * w.featsize), o->dfd = a.take<float>(Rd * w.featsize);
}

static int32_t lstm_train_fwd(float* gates, const float* whht, const int64_t* len, int n, int L, int h, int dirs, float* out,
                              float* c_seq, int* err, cudaStream_t s) {
  const int G = 4 * h, hp = (h + 3) & ~3;
  const size_t wbytes = (size_t)h * G * sizeof(float);
  const size_t rest = ((size_t)TR_TS * hp + TR_TS * h + TR_TS * G) * sizeof(float);
  const bool wsmem = wbytes + rest <= 200 * 1024;
  const size_t smem = (wsmem ? wbytes : 0) + rest;
  dim3 grid((n + TR_TS - 1) / TR_TS, dirs);
  if (wsmem) {
    CAIR_CUDA(cudaFuncSetAttribute(lstm_train_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAIR_LAUNCH(lstm_train_fwd_kernel<true>, grid, 256, smem, s, gates, whht, len, n, L, h, dirs, out, c_seq, err);
  } else {
    CAIR_CUDA(cudaFuncSetAttribute(lstm_train_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAIR_LAUNCH(lstm_train_fwd_kernel<false>, grid, 256, smem, s, gates, whht, len, n, L, h, dirs, out, c_seq, err);
  }
  return CAIR_OK;
}

static int32_t lstm_train_bwd(float* gates, const float* c_seq, const float* denc, const float* w_hh_fwd, const float* w_hh_rev,
                              float* whh_scratch, const int64_t* len, int n, int L, int h, int dirs, cudaStream_t s) {
  // the kernel wants [dirs][4h][h] contiguous: the two directions are separate parameters, so stage them side by side
  const int G = 4 * h;
  CAIR_CUDA(cudaMemcpyAsync(whh_scratch, w_hh_fwd, (size_t)G * h * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (dirs == 2)
    CAIR_CUDA(cudaMemcpyAsync(whh_scratch + (size_t)G * h, w_hh_rev, (size_t)G * h * sizeof(float), cudaMemcpyDeviceToDevice, s));
  const size_t wbytes = (size_t)G * h * sizeof(float);
  // Sequences per CTA: the fewest of 8 / 12 / 16 whose grid fits ONE wave of two resident CTAs per SM (1280 documents x 2
  // directions at 8 per CTA are 320 CTAs = a full wave plus a 24-CTA tail that costs a second full pass)
  int ts = 8;
  while (ts < 16 && (int64_t)((n + ts - 1) / ts) * dirs > 2 * kSMs) ts += 4;
  const size_t rest = ((size_t)G * ts + 2 * ts * h + 4 * ts * h) * sizeof(float);
  const bool wsmem = wbytes + rest <= 200 * 1024;
  const size_t smem = (wsmem ? wbytes : 0) + rest;
  dim3 grid((n + ts - 1) / ts, dirs);
#define CAIR_BWD_LAUNCH(WS, TSV)                                                                                           \
  do {                                                                                                                     \
    CAIR_CUDA(cudaFuncSetAttribute(lstm_train_bwd_kernel<WS, TSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    CAIR_LAUNCH((lstm_train_bwd_kernel<WS, TSV>), grid, 256, smem, s, gates, c_seq, denc, whh_scratch, len, n, L, h, dirs);  \
  } while (0)
  if (wsmem) {
    if (ts == 8) CAIR_BWD_LAUNCH(true, 8);
    else if (ts == 12) CAIR_BWD_LAUNCH(true, 12);
    else CAIR_BWD_LAUNCH(true, 16);
  } else {
    if (ts == 8) CAIR_BWD_LAUNCH(false, 8);
    else if (ts == 12) CAIR_BWD_LAUNCH(false, 12);
    else CAIR_BWD_LAUNCH(false, 16);
  }
#undef CAIR_BWD_LAUNCH
  return CAIR_OK;
}

static int32_t refresh_lstm(const cair_lstm_dir& fwd, const cair_lstm_dir& rev, int dirs, int in, int h, float* whht, float* bias,
                            float* wiht, cudaStream_t s) {
  const int G = 4 * h;
  for (int dd = 0; dd < dirs; ++dd) {
    const cair_lstm_dir& w = dd ? rev : fwd;
    CAIR_LAUNCH(transpose_kernel, (G * h + 255) / 256, 256, 0, s, w.w_hh, G, h, whht + (size_t)dd * h * G, (int64_t)G);
    CAIR_LAUNCH(add_vec_kernel, (G + 255) / 256, 256, 0, s, w.b_ih, w.b_hh, G, bias + (size_t)dd * G);
    // wiht[f][dd*G + g] = w_ih[g][f]
    CAIR_LAUNCH(transpose_kernel, (G * in + 255) / 256, 256, 0, s, w.w_ih, G, in, wiht + (size_t)dd * G, (int64_t)dirs * G);
  }
  return CAIR_OK;
}

int32_t mt_repack(const cair_mt_weights& w, MtPack* p, cudaStream_t s);   // mt.cu
int32_t mt_interact_train(const MtPack& p, const float* cq, const float* cd, float* T, const int64_t* q, const int64_t* d, int N,
                          int Lq, int Ld, int64_t pairs, int64_t nq, float* scores, float* pooled, int* argidx,
                          cudaStream_t s);   // mt.cu


// ------------------------------------------------------------------------------------------------
// DRMM training step (neuroir/rankers/drmm.py:29-84 in train mode).  The histogram is computed with numpy in the reference
// (no gradient flows through the cosines), so the differentiable part is tiny: gating softmax over the (dropped) query
// embeddings, ffnn (5 -> 1 -> 1), output layer.  Forward: the dropped embedding rows of this batch are materialised as a
// per-batch "virtual table" (query rows first, then document rows; row r of it = token r of the batch) and the ordinary
// DRMM kernels run on it with ids 0, 1, 2, ... - so train-mode dropout reaches the cosines exactly as in the reference,
// and with p = 0 the arithmetic is the eval path's.  The histograms are kept for the backward.
__global__ void iota_kernel(int64_t* p, int64_t n, int64_t base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = base + i;
}

// one CTA per query: its N documents share the term gate
__global__ void __launch_bounds__(128) drmm_train_bwd_kernel(const float* __restrict__ vtable, const int64_t* __restrict__ q,
                                                             const int32_t* __restrict__ hist, const float* __restrict__ dscores,
                                                             const float* __restrict__ wg, const float* __restrict__ bg,
                                                             const float* __restrict__ w0, const float* __restrict__ b0,
                                                             const float* __restrict__ w1, const float* __restrict__ b1,
                                                             const float* __restrict__ wo, int V, int E, int N, int Lq, float p,
                                                             uint64_t seed, float* g_wg, float* g_bg, float* g_w0, float* g_b0,
                                                             float* g_w1, float* g_b1, float* g_wo, float* g_bo, float* g_table) {
  __shared__ float gate[32], dg[32], da[32], acc[16];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  if (tid < 16) acc[tid] = 0.f;
  // gate = softmax_i(wg . xq_i + bg) over all Lq positions (drmm.py:95-98)
  for (int i = warp; i < Lq; i += 4) {
    const float* x = vtable + ((size_t)b * Lq + i) * E;
    float a = 0.f;
    for (int e = lane; e < E; e += 32) a = fmaf(wg[e], x[e], a);
    a = warp_sum(a);
    if (lane == 0) gate[i] = a + bg[0];
  }
  __syncthreads();
  if (warp == 0) {
    float v = lane < Lq ? gate[lane] : -INFINITY;
    const float mx = warp_max(v);
    const float ex = lane < Lq ? expf(v - mx) : 0.f;
    const float den = warp_sum(ex);
    if (lane < Lq) gate[lane] = ex / den, dg[lane] = 0.f;
  }
  __syncthreads();
  // per document: score = wo * sum_i f_i g_i + bo, f_i = w1 (w0 . hist_i + b0) + b1
  if (warp == 0) {
    float a_wo = 0.f, a_bo = 0.f, a_w1 = 0.f, a_b1 = 0.f, a_b0 = 0.f, a_w0[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, dgl = 0.f;
    for (int n = 0; n < N; ++n) {
      const int64_t pr = (int64_t)b * N + n;
      const float ds = dscores[pr];
      float z = 0.f, f = 0.f, h5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      if (lane < Lq) {
        z = b0[0];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          h5[k] = (float)hist[(pr * Lq + lane) * 5 + k];
          z = fmaf(w0[k], h5[k], z);
        }
        f = w1[0] * z + b1[0];
      }
      const float S = warp_sum(lane < Lq ? f * gate[lane] : 0.f);
      a_wo += ds * S, a_bo += ds;
      const float dS = ds * wo[0];
      if (lane < Lq) {
        const float df = dS * gate[lane];
        dgl += dS * f;
        a_w1 += df * z, a_b1 += df;
        const float dz = df * w1[0];
        a_b0 += dz;
#pragma unroll
        for (int k = 0; k < 5; ++k) a_w0[k] += dz * h5[k];
      }
    }
    a_w1 = warp_sum(a_w1), a_b1 = warp_sum(a_b1), a_b0 = warp_sum(a_b0);
#pragma unroll
    for (int k = 0; k < 5; ++k) a_w0[k] = warp_sum(a_w0[k]);
    // softmax backward
    const float gd = warp_sum(lane < Lq ? gate[lane] * dgl : 0.f);
    if (lane < Lq) da[lane] = gate[lane] * (dgl - gd);
    const float sda = warp_sum(lane < Lq ? gate[lane] * (dgl - gd) : 0.f);
    if (lane == 0) {
      atomicAdd(g_wo, a_wo), atomicAdd(g_bo, a_bo), atomicAdd(g_w1, a_w1), atomicAdd(g_b1, a_b1), atomicAdd(g_b0, a_b0);
#pragma unroll
      for (int k = 0; k < 5; ++k) atomicAdd(&g_w0[k], a_w0[k]);
      atomicAdd(g_bg, sda);
    }
  }
  __syncthreads();
  // d wg = sum_i da_i xq_i ;  d xq_i = da_i wg  -> mask -> table rows (PAD skipped)
  for (int e = tid; e < E; e += 128) {
    float a = 0.f;
    for (int i = 0; i < Lq; ++i) a = fmaf(da[i], vtable[((size_t)b * Lq + i) * E + e], a);
    if (a != 0.f) atomicAdd(&g_wg[e], a);
    if (g_table) {
      const float w = wg[e];
      for (int i = 0; i < Lq; ++i) {
        const int64_t id = q[(size_t)b * Lq + i];
        if (id <= 0 || id >= V) continue;
        const float v = da[i] * w * drop_scale(seed, (uint64_t)(((size_t)b * Lq + i) * E + e), p, inv);
        if (v != 0.f) atomicAdd(&g_table[id * E + e], v);
      }
    }
  }
}

struct DrmmTrainWs {
  int* err;
  float* vtable;
  int64_t *vq, *vd;
  int32_t* hist;
};
static void drmm_train_layout(Arena& a, int E, int B, int N, int Lq, int Ld, DrmmTrainWs* o) {
  const size_t Rq = (size_t)B * Lq, Rd = (size_t)B * N * Ld;
  o->err = a.take<int>(64);
  o->vtable = a.take<float>((Rq + Rd) * E);
  o->vq = a.take<int64_t>(Rq), o->vd = a.take<int64_t>(Rd);
  o->hist = a.take<int32_t>((size_t)B * N * Lq * 5);
}

// ------------------------------------------------------------------------------------------------
// DSSM (dssm.py:33-63) in train mode: x[r, e] = max_t table[id_t, e] * mask(t, e) with the arg-max position kept for the
// backward; the two Linear + Tanh layers per side are small dense GEMMs; cosine as in ESM.
struct DssmTrainWs {
  float *x, *h1, *y, *nrm, *dy, *dh1, *dx;   // rows: B queries then B*N documents
  int* arg;                                    // [R, E] arg-max token position
  float *w0t_q, *w2t_q, *w0t_d, *w2t_d;        // transposed weights for the input gradients
  int* err;
};
static void dssm_train_layout(Arena& a, int E, int H, int O, int B, int N, DssmTrainWs* o) {
  const size_t R = (size_t)B + (size_t)B * N;
  o->x = a.take<float>(R * E), o->h1 = a.take<float>(R * H), o->y = a.take<float>(R * O), o->nrm = a.take<float>(R);
  o->dy = a.take<float>(R * O), o->dh1 = a.take<float>(R * H), o->dx = a.take<float>(R * E);
  o->arg = a.take<int>(R * E);
  o->w0t_q = a.take<float>((size_t)E * H), o->w2t_q = a.take<float>((size_t)H * O);
  o->w0t_d = a.take<float>((size_t)E * H), o->w2t_d = a.take<float>((size_t)H * O);
  o->err = a.take<int>(4);
}

// one CTA per row; mask element index = (token row of the batch: queries first) * E + e, as embed_drop_kernel
__global__ void __launch_bounds__(256) dssm_train_pool_kernel(const float* __restrict__ table, const int64_t* __restrict__ q,
                                                              const int64_t* __restrict__ d, int V, int E, int B, int Lq, int Ld,
                                                              float p, uint64_t seed, float* __restrict__ x, int* __restrict__ arg,
                                                              int* err) {
  const int64_t r = blockIdx.x;
  const bool isq = r < B;
  const int L = isq ? Lq : Ld;
  const int64_t* ids = isq ? q + r * Lq : d + (r - B) * Ld;
  const int64_t tok0 = isq ? r * Lq : (int64_t)B * Lq + (r - B) * Ld;
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  for (int e = threadIdx.x; e < E; e += 256) {
    float best = -INFINITY;
    int bt = 0;
    for (int t = 0; t < L; ++t) {
      const int64_t id = checked_id(ids[t], V, err);
      const float v = table[id * E + e] * drop_scale(seed, (uint64_t)((tok0 + t) * E + e), p, inv);
      if (v > best) best = v, bt = t;
    }
    x[r * E + e] = best;
    arg[r * E + e] = bt;
  }
}

__global__ void row_norm_kernel(const float* __restrict__ y, int O, int64_t R, float* __restrict__ nrm) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  float ss = 0.f;
  for (int e = lane; e < O; e += 32) ss += y[r * O + e] * y[r * O + e];
  ss = warp_sum(ss);
  if (lane == 0) nrm[r] = fmaxf(sqrtf(ss), 1e-8f);
}

__global__ void cos_score_kernel(const float* __restrict__ y, const float* __restrict__ nrm, int O, int B, int N,
                                 float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= (int64_t)B * N) return;
  const int64_t b = p / N;
  const float* a = y + b * O;
  const float* c = y + ((int64_t)B + p) * O;
  const float na = nrm[b], nc = nrm[B + p];
  float dot = 0.f;
  for (int e = lane; e < O; e += 32) dot += (a[e] / na) * (c[e] / nc);
  dot = warp_sum(dot);
  if (lane == 0) scores[p] = dot;
}

// d(pre-activation of the last Tanh) for every row: the cosine's derivative (see esm_train_bwd_kernel) times 1 - y^2
__global__ void __launch_bounds__(128) dssm_train_cos_bwd_kernel(const float* __restrict__ y, const float* __restrict__ nrm,
                                                                 const float* __restrict__ scores, const float* __restrict__ dscores,
                                                                 int O, int B, int N, float* __restrict__ dy) {
  const int b = blockIdx.x;
  const float* a = y + (size_t)b * O;
  const float na = nrm[b];
  for (int e = threadIdx.x; e < O; e += 128) {
    const float ah = a[e] / na;
    float da = 0.f;
    for (int n = 0; n < N; ++n) {
      const int64_t p = (int64_t)b * N + n;
      const float nc = nrm[B + p], s = scores[p], g = dscores[p];
      const float cv = y[((size_t)B + p) * O + e], ch = cv / nc;
      da += g * (na <= 1e-8f ? ch : (ch - s * ah)) / na;
      dy[((size_t)B + p) * O + e] = g * (nc <= 1e-8f ? ah : (ah - s * ch)) / nc * (1.f - cv * cv);
    }
    dy[(size_t)b * O + e] = da * (1.f - a[e] * a[e]);
  }
}

__global__ void tanh_bwd_kernel(float* __restrict__ dh, const float* __restrict__ h, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dh[i] *= 1.f - h[i] * h[i];
}

// d table[id of the arg-max token, e] += dx[r, e] * mask   (PAD: no gradient)
__global__ void __launch_bounds__(256) dssm_train_scatter_kernel(const float* __restrict__ dx, const int* __restrict__ arg,
                                                                 const int64_t* __restrict__ q, const int64_t* __restrict__ d, int V,
                                                                 int E, int B, int Lq, int Ld, float p, uint64_t seed,
                                                                 float* __restrict__ dtable) {
  const int64_t r = blockIdx.x;
  const bool isq = r < B;
  const int64_t* ids = isq ? q + r * Lq : d + (r - B) * Ld;
  const int64_t tok0 = isq ? r * Lq : (int64_t)B * Lq + (r - B) * Ld;
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  for (int e = threadIdx.x; e < E; e += 256) {
    const int t = arg[r * E + e];
    const int64_t id = ids[t];
    if (id <= 0 || id >= V) continue;
    const float m = drop_scale(seed, (uint64_t)((tok0 + t) * E + e), p, inv);
    if (m != 0.f) atomicAdd(dtable + id * E + e, dx[r * E + e] * m);
  }
}

// ------------------------------------------------------------------------------------------------
// CDSSM (cdssm.py:42-77) in train mode.  The interleave + Conv1d(k = 3) pair is one linear map over 5-token windows (the merged
// weight W5 of the scoring path, cdssm_merge_kernel), so with one ROW PER TOKEN (window r = the 5E contiguous floats starting at
// token row r of the dropped embedding matrix; rows t >= L - 4 of a sequence are not windows of it: never pooled, zero
// gradient) every layer is a dense GEMM forward and backward:
//   h1 = tanh(X5 W5^T + b) [R, H],  s2 = tanh(h1 Ws^T + bs) [R, O],  y = max over the valid t (arg-max kept),  cosine;
//   dz2 = scatter of dy (1 - y^2) to the arg-max rows,  dWs = dz2^T h1,  dh1 = (dz2 Ws) (1 - h1^2),  dW5 = dh1^T X5,
//   dx[r] = sum_j dh1[r - j] W5[:, j]  = one GEMM over the 5H floats ending at row r of dh1 (4 zero rows in front),
//   table gradient = dx * mask scattered to the non-PAD token ids, conv gradient = W5 gradient un-merged.
void cdssm_merge_launch(const float* w, int H, int E, float* w5, cudaStream_t s);   // dssm.cu

struct CdssmTrainWs {
  float *x, *h1, *s2, *y, *nrm, *dy, *dz2, *dh1, *dx;   // token rows: B*Lq query rows, then B*N*Ld document rows (+ padding rows)
  int* arg;                                               // [B + B*N, O] arg-max token row
  float *w5[2], *dw5[2], *wst[2], *w5r[2];                // merged conv weight [H,5E], its gradient, Ws^T [H,O], reordered [E,5H]
  int* err;
};
static void cdssm_train_layout(Arena& a, int E, int H, int O, int B, int N, int Lq, int Ld, CdssmTrainWs* o) {
  const size_t Rt = (size_t)B * Lq + (size_t)B * N * Ld, Rs = (size_t)B + (size_t)B * N;
  o->x = a.take<float>((Rt + 4) * E);
  o->h1 = a.take<float>(Rt * H), o->s2 = a.take<float>(Rt * O);
  o->y = a.take<float>(Rs * O), o->nrm = a.take<float>(Rs), o->dy = a.take<float>(Rs * O);
  o->dz2 = a.take<float>(Rt * O), o->dh1 = a.take<float>((Rt + 4) * H), o->dx = a.take<float>(Rt * E);
  o->arg = a.take<int>(Rs * O);
  for (int i = 0; i < 2; ++i) {
    o->w5[i] = a.take<float>((size_t)H * 5 * E), o->dw5[i] = a.take<float>((size_t)H * 5 * E);
    o->wst[i] = a.take<float>((size_t)H * O), o->w5r[i] = a.take<float>((size_t)E * 5 * H);
  }
  o->err = a.take<int>(4);
}

// y[s, o] = max over the valid windows t < L - 4 of s2[(row0 + s*L + t), o]; arg = that token row (first maximum)
__global__ void cdssm_train_max_kernel(const float* __restrict__ s2, int O, int L, int64_t row0, int64_t nseq, float* __restrict__ y,
                                       int* __restrict__ arg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nseq * O) return;
  const int64_t s = i / O;
  const int o = (int)(i - s * O);
  float best = -INFINITY;
  int bt = 0;
  for (int t = 0; t < L - 4; ++t) {
    const float v = s2[(row0 + s * L + t) * O + o];
    if (v > best) best = v, bt = t;
  }
  y[i] = best;
  arg[i] = (int)(row0 + s * L + bt);
}

__global__ void cdssm_train_route_kernel(const float* __restrict__ dy, const int* __restrict__ arg, int O, int64_t n,
                                         float* __restrict__ dz2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dz2[(int64_t)arg[i] * O + (i % O)] = dy[i];
}

// w5r[e][jj*H + h] = w5[h][(4 - jj)*E + e]: the weight of "dx[r] = [dh1[r-4] .. dh1[r]] . w5r^T"
__global__ void cdssm_train_reorder_kernel(const float* __restrict__ w5, int H, int E, float* __restrict__ w5r) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)E * 5 * H) return;
  const int h = (int)(i % H), jj = (int)((i / H) % 5), e = (int)(i / ((int64_t)5 * H));
  w5r[i] = w5[(size_t)h * 5 * E + (size_t)(4 - jj) * E + e];
}

// d conv.weight[f][wi*E + e][k] += d W5[f][(wi + k)*E + e]   (the transpose of cdssm_merge_kernel)
__global__ void cdssm_train_unmerge_kernel(const float* __restrict__ dw5, int H, int E, float* __restrict__ dw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)H * 3 * E * 3) return;
  const int k = (int)(i % 3);
  const int64_t c = i / 3;
  const int e = (int)(c % E), wi = (int)((c / E) % 3), f = (int)(c / ((int64_t)3 * E));
  dw[i] += dw5[(size_t)f * 5 * E + (size_t)(wi + k) * E + e];
}

// d table[ids[r], e] += dx[r, e] * mask(row0 + r, e)   (PAD: no gradient)
__global__ void embed_scatter_kernel(const float* __restrict__ dx, const int64_t* __restrict__ ids, int V, int E, int64_t rows,
                                     int64_t row0, float p, uint64_t seed, float* __restrict__ dtable) {
  const float inv = p < 1.f ? 1.f / (1.f - p) : 0.f;
  const int64_t total = rows * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int e = (int)(i - r * E);
    const int64_t id = ids[r];
    if (id <= 0 || id >= V) continue;
    const float m = drop_scale(seed, (uint64_t)((row0 + r) * E + e), p, inv);
    const float g = dx[(row0 + r) * E + e] * m;
    if (g != 0.f) atomicAdd(dtable + id * E + e, g);
  }
}

}  // namespace cair

using namespace cair;

struct cair_mt_trainer {
  MtTrainer t;
  float* whh_scratch = nullptr;
};

namespace {
struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
}  // namespace

extern "C" {

int32_t cair_dropout_mask(uint64_t seed, float p, int64_t n, float* out, void* stream) {
  if (!out || n < 0 || p < 0.f || p > 1.f) return fail(CAIR_ERR_BAD_ARG, "dropout_mask: bad argument");
  if (n == 0) return CAIR_OK;
  CAIR_LAUNCH(dropout_mask_kernel, 592, 256, 0, (cudaStream_t)stream, seed, p, n, out);
  return CAIR_OK;
}

int32_t cair_mt_train_create(const cair_mt_weights* w, int32_t device, cair_mt_trainer** out) {
  if (!w || !out) return fail(CAIR_ERR_BAD_ARG, "mt_train_create: null argument");
  if (w->rnn_type != CAIR_RNN_LSTM) return fail(CAIR_ERR_UNSUPPORTED, "mt_train: only LSTM encoders have a backward pass");
  DevGuard g(device);
  cair_mt_trainer* h = new cair_mt_trainer();
  MtTrainer& t = h->t;
  t.device = device, t.w = *w, t.dirs = w->bidirectional ? 2 : 1;
  t.hq = w->nhid_query / t.dirs, t.hd = w->nhid_doc / t.dirs;
  cudaStream_t s = 0;
  auto body = [&]() -> int32_t {
    CAIR_TRY(mt_pack(t.own, *w, &t.pack, s));
    const int C1 = w->nchannels + 1, FPP = t.pack.FPP;
    if (21 * C1 > TRB_MAXT * 256) return fail(CAIR_ERR_UNSUPPORTED, "mt_train: nchannels %d too large", w->nchannels);
    CAIR_CUDA(t.own.alloc(&t.w7, (size_t)21 * C1 * FPP));
    CAIR_CUDA(t.own.alloc(&t.whht_q, (size_t)t.dirs * t.hq * 4 * t.hq));
    CAIR_CUDA(t.own.alloc(&t.whht_d, (size_t)t.dirs * t.hd * 4 * t.hd));
    CAIR_CUDA(t.own.alloc(&t.bias_q, (size_t)t.dirs * 4 * t.hq));
    CAIR_CUDA(t.own.alloc(&t.bias_d, (size_t)t.dirs * 4 * t.hd));
    CAIR_CUDA(t.own.alloc(&t.wiht_q, (size_t)w->featsize * t.dirs * 4 * t.hq));
    CAIR_CUDA(t.own.alloc(&t.wiht_d, (size_t)w->featsize * t.dirs * 4 * t.hd));
    CAIR_CUDA(t.own.alloc(&t.wqt, (size_t)w->nhid_query * w->nchannels));
    CAIR_CUDA(t.own.alloc(&t.wdt, (size_t)w->nhid_doc * w->nchannels));
    CAIR_TRY(gemm_tc_pack(t.own, w->linear_projection.w, w->featsize, w->emsize, &t.tw_proj, s));
    for (int dd = 0; dd < t.dirs; ++dd) {
      CAIR_TRY(gemm_tc_pack(t.own, (dd ? w->query_rev : w->query_fwd).w_ih, 4 * t.hq, w->featsize, &t.tw_ih_q[dd], s));
      CAIR_TRY(gemm_tc_pack(t.own, (dd ? w->doc_rev : w->doc_fwd).w_ih, 4 * t.hd, w->featsize, &t.tw_ih_d[dd], s));
    }
    CAIR_TRY(gemm_tc_pack(t.own, w->query_projection.w, w->nchannels, w->nhid_query, &t.tw_cq, s));
    CAIR_TRY(gemm_tc_pack(t.own, w->document_projection.w, w->nchannels, w->nhid_doc, &t.tw_cd, s));
    CAIR_TRY(gemm_tc_pack(t.own, t.wiht_q, w->featsize, t.dirs * 4 * t.hq, &t.tw_wiht_q, s));   // contents refreshed per step
    CAIR_TRY(gemm_tc_pack(t.own, t.wiht_d, w->featsize, t.dirs * 4 * t.hd, &t.tw_wiht_d, s));
    if (lstm_tc_supported(w->featsize, t.hq))
      CAIR_TRY(lstm_tc_pack(t.own, &w->query_fwd, t.dirs == 2 ? &w->query_rev : nullptr, w->featsize, t.hq, &t.tcp_q, s));
    if (lstm_tc_supported(w->featsize, t.hd))
      CAIR_TRY(lstm_tc_pack(t.own, &w->doc_fwd, t.dirs == 2 ? &w->doc_rev : nullptr, w->featsize, t.hd, &t.tcp_d, s));
    const int hm = t.hq > t.hd ? t.hq : t.hd;
    CAIR_CUDA(t.own.alloc(&h->whh_scratch, (size_t)t.dirs * 4 * hm * hm));
    CAIR_CUDA(cudaStreamSynchronize(s));
    return CAIR_OK;
  };
  const int32_t rc = body();
  if (rc != CAIR_OK) {
    t.own.release();
    delete h;
    *out = nullptr;
    return rc;
  }
  *out = h;
  return CAIR_OK;
}

int32_t cair_mt_train_set_impl(cair_mt_trainer* h, int32_t tc_forward) {
  if (!h) return fail(CAIR_ERR_BAD_ARG, "mt_train_set_impl: null trainer");
  h->t.tc_forward = tc_forward ? 1 : 0;
  return CAIR_OK;
}

int32_t cair_mt_train_destroy(cair_mt_trainer* h) {
  if (!h) return CAIR_OK;
  DevGuard g(h->t.device);
  cudaDeviceSynchronize();
  h->t.own.release();
  delete h;
  return CAIR_OK;
}

int32_t cair_mt_train_workspace_bytes(cair_mt_trainer* h, int32_t B, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes) {
  if (!h || !bytes || B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_ARG, "mt_train_workspace_bytes: bad argument");
  Arena a(nullptr, 0);
  MtTrainWs o;
  mt_train_layout(h->t, a, B, N, Lq, Ld, &o);
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_mt_train_forward(cair_mt_trainer* h, const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen,
                              int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, float* scores, void* ws,
                              size_t ws_bytes, void* stream) {
  if (!h || !q || !qlen || !d || !dlen || !scores || !ws) return fail(CAIR_ERR_BAD_ARG, "mt_train_forward: null argument");
  if (p_drop < 0.f || p_drop >= 1.f) return fail(CAIR_ERR_BAD_ARG, "mt_train_forward: dropout must be in [0, 1)");
  if ((uintptr_t)ws % 256) return fail(CAIR_ERR_WORKSPACE, "mt_train_forward: workspace must be 256-byte aligned");
  MtTrainer& t = h->t;
  DevGuard g(t.device);
  cudaStream_t s = (cudaStream_t)stream;
  const cair_mt_weights& w = t.w;
  Arena a(ws, ws_bytes);
  MtTrainWs o;
  mt_train_layout(t, a, B, N, Lq, Ld, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "mt_train_forward: workspace too small (%zu < %zu)", ws_bytes, a.off);
  const int64_t Rq = (int64_t)B * Lq, Rd = (int64_t)B * N * Ld, P = (int64_t)B * N;
  const int E = w.emsize, F = w.featsize, C = w.nchannels, Hq = w.nhid_query, Hd = w.nhid_doc;
  CAIR_CUDA(cudaMemsetAsync(o.err, 0, 256, s));
  // weights change every step: refresh the repacked copies from the live parameters
  CAIR_TRY(mt_repack(w, &t.pack, s));
  CAIR_LAUNCH(mt_train_pack_kernel, (21 * (C + 1) * t.pack.FPP + 255) / 256, 256, 0, s, w.conv1.w, w.conv2.w, w.conv3.w, C + 1,
              w.nfilters, t.pack.FP, t.pack.FPP, t.w7);
  CAIR_TRY(refresh_lstm(w.query_fwd, w.query_rev, t.dirs, F, t.hq, t.whht_q, t.bias_q, t.wiht_q, s));
  CAIR_TRY(refresh_lstm(w.doc_fwd, w.doc_rev, t.dirs, F, t.hd, t.whht_d, t.bias_d, t.wiht_d, s));
  CAIR_LAUNCH(transpose_kernel, (C * Hq + 255) / 256, 256, 0, s, w.query_projection.w, C, Hq, t.wqt, (int64_t)C);
  CAIR_LAUNCH(transpose_kernel, (C * Hd + 255) / 256, 256, 0, s, w.document_projection.w, C, Hd, t.wdt, (int64_t)C);
  const GemmTcW none{};
  const bool tcg = t.tc_forward != 0;
  if (tcg) {
    CAIR_TRY(gemm_tc_repack(w.linear_projection.w, t.tw_proj, s));
    for (int dd = 0; dd < t.dirs; ++dd) {
      CAIR_TRY(gemm_tc_repack((dd ? w.query_rev : w.query_fwd).w_ih, t.tw_ih_q[dd], s));
      CAIR_TRY(gemm_tc_repack((dd ? w.doc_rev : w.doc_fwd).w_ih, t.tw_ih_d[dd], s));
    }
    CAIR_TRY(gemm_tc_repack(w.query_projection.w, t.tw_cq, s));
    CAIR_TRY(gemm_tc_repack(w.document_projection.w, t.tw_cd, s));
    CAIR_TRY(gemm_tc_repack(t.wiht_q, t.tw_wiht_q, s));
    CAIR_TRY(gemm_tc_repack(t.wiht_d, t.tw_wiht_d, s));
  }
  // embedding + dropout (mtensor.py:77-84), linear_projection (:88-90)
  CAIR_LAUNCH(embed_drop_kernel, 1184, 256, 0, s, w.table, q, w.vocab, E, Rq, (int64_t)0, p_drop, seed, o.xq, o.err);
  CAIR_LAUNCH(embed_drop_kernel, 1184, 256, 0, s, w.table, d, w.vocab, E, Rd, Rq, p_drop, seed, o.xd, o.err);
  CAIR_TRY(gemm_auto(gemm_dense(o.xq, E), w.linear_projection.w, tcg ? t.tw_proj : none, w.linear_projection.b, o.fq, F, Rq, F, E, ACT_NONE, s));
  CAIR_TRY(gemm_auto(gemm_dense(o.xd, E), w.linear_projection.w, tcg ? t.tw_proj : none, w.linear_projection.b, o.fd, F, Rd, F, E, ACT_NONE, s));
  // encoders (:93-94): the tcgen05 recurrence with its fused input projection where the encoder fits it (it leaves the gate
  // activations and cell states the BPTT kernel reads), else pre-gate GEMMs + the fp32 recurrence
  const bool rq = tcg && t.tcp_q.wimg != nullptr && g_rnn_impl != RNN_IMPL_FP32;
  const bool rd = tcg && t.tcp_d.wimg != nullptr && g_rnn_impl != RNN_IMPL_FP32;
  if (rq) CAIR_TRY(lstm_tc_repack(t.tcp_q, &w.query_fwd, &w.query_rev, s));
  if (rd) CAIR_TRY(lstm_tc_repack(t.tcp_d, &w.doc_fwd, &w.doc_rev, s));
  for (int dd = 0; dd < t.dirs; ++dd) {
    const int Gq = 4 * t.hq, Gd = 4 * t.hd;
    if (!rq)
      CAIR_TRY(gemm_auto(gemm_dense(o.fq, F), (dd ? w.query_rev : w.query_fwd).w_ih, tcg ? t.tw_ih_q[dd] : none, t.bias_q + (size_t)dd * Gq,
                         o.gq + (size_t)dd * Gq, (int64_t)t.dirs * Gq, Rq, Gq, F, ACT_NONE, s));
    if (!rd)
      CAIR_TRY(gemm_auto(gemm_dense(o.fd, F), (dd ? w.doc_rev : w.doc_fwd).w_ih, tcg ? t.tw_ih_d[dd] : none, t.bias_d + (size_t)dd * Gd,
                         o.gd + (size_t)dd * Gd, (int64_t)t.dirs * Gd, Rd, Gd, F, ACT_NONE, s));
  }
  if (rq)
    CAIR_TRY(lstm_tc_run(t.tcp_q, nullptr, gemm_dense(o.fq, F), qlen, B, Lq, o.enc_q, nullptr, nullptr, o.err, s, nullptr, nullptr, 8,
                         o.gq, o.cseq_q));
  else
    CAIR_TRY(lstm_train_fwd(o.gq, t.whht_q, qlen, B, Lq, t.hq, t.dirs, o.enc_q, o.cseq_q, o.err, s));
  if (rd)
    CAIR_TRY(lstm_tc_run(t.tcp_d, nullptr, gemm_dense(o.fd, F), dlen, (int)P, Ld, o.enc_d, nullptr, nullptr, o.err, s, nullptr, nullptr, 8,
                         o.gd, o.cseq_d));
  else
    CAIR_TRY(lstm_train_fwd(o.gd, t.whht_d, dlen, (int)P, Ld, t.hd, t.dirs, o.enc_d, o.cseq_d, o.err, s));
  // channel projections (:99, :108)
  CAIR_TRY(gemm_auto(gemm_dense(o.enc_q, Hq), w.query_projection.w, tcg ? t.tw_cq : none, w.query_projection.b, o.cq, C, Rq, C, Hq, ACT_NONE, s));
  CAIR_TRY(gemm_auto(gemm_dense(o.enc_d, Hd), w.document_projection.w, tcg ? t.tw_cd : none, w.document_projection.b, o.cd, C, Rd, C, Hd, ACT_NONE, s));
  // interaction (:113-131) with the arg-max cells of the two max-pools: the tensor-core kernel of the scoring path (its
  // arg-max instantiation) where the shape allows, else the fp32 kernel.  The backward recomputes the winning cells in fp32.
  if (t.tc_forward && o.timg) {
    MtEpiConst ec;
    CAIR_TRY(mt_epi_const(t.pack, &ec, s));   // epilogue constants of the CURRENT weights (synchronises s)
    CAIR_TRY(mt_tc_build_t(t.pack, o.cq, o.timg, Lq, B, s));
    CAIR_TRY(mt_tc_doc_image(t.pack, o.cd, o.aimg, Ld, P, s));
    return mt_tc_interact(t.pack, ec, o.timg, o.aimg, q, d, N, Lq, Ld, 0, P, 0, B, scores, s, 0, o.pooled, o.argidx);
  }
  return mt_interact_train(t.pack, o.cq, o.cd, o.T, q, d, N, Lq, Ld, P, B, scores, o.pooled, o.argidx, s);
}

static float* gp(const float* p) { return const_cast<float*>(p); }

int32_t cair_mt_train_backward(cair_mt_trainer* h, const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen,
                               int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, const float* dscores,
                               const cair_mt_weights* grads, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !q || !qlen || !d || !dlen || !dscores || !grads || !ws) return fail(CAIR_ERR_BAD_ARG, "mt_train_backward: null argument");
  MtTrainer& t = h->t;
  DevGuard g(t.device);
  cudaStream_t s = (cudaStream_t)stream;
  const cair_mt_weights& w = t.w;
  const cair_mt_weights& G = *grads;
  if (!G.linear_projection.w || !G.linear_projection.b || !G.query_projection.w || !G.query_projection.b ||
      !G.document_projection.w || !G.document_projection.b || !G.alpha || !G.conv1.w || !G.conv2.w || !G.conv3.w || !G.conv1.b ||
      !G.conv2.b || !G.conv3.b || !G.conv.w || !G.conv.b || !G.output.w || !G.output.b || !G.query_fwd.w_ih || !G.doc_fwd.w_ih)
    return fail(CAIR_ERR_BAD_ARG, "mt_train_backward: null gradient pointer (only `table` may be NULL: fixed embeddings)");
  Arena a(ws, ws_bytes);
  MtTrainWs o;
  mt_train_layout(t, a, B, N, Lq, Ld, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "mt_train_backward: workspace too small");
  const int64_t Rq = (int64_t)B * Lq, Rd = (int64_t)B * N * Ld, P = (int64_t)B * N;
  const int E = w.emsize, F = w.featsize, C = w.nchannels, Hq = w.nhid_query, Hd = w.nhid_doc, M = w.match_filter_size;
  // ---- interaction stack (sparse) ----
  CAIR_CUDA(cudaMemsetAsync(o.dcq, 0, (size_t)Rq * C * sizeof(float), s));
  CAIR_CUDA(cudaMemsetAsync(o.dcd, 0, (size_t)Rd * C * sizeof(float), s));
  {
    const int C1 = C + 1, FP = t.pack.FP, FPP = t.pack.FPP;
    const size_t rest = ((size_t)FPP + 8 * FPP + (size_t)M * FP + 2 * M + FP + 4) * sizeof(float) + (size_t)(Lq + Ld) * sizeof(int);
    const size_t w7b = (size_t)21 * C1 * FPP * sizeof(float);
    const int w7_smem = 2 * w7b + rest <= 220 * 1024;
    const size_t smem = (w7_smem ? 2 : 1) * w7b + rest;
    if (smem > 220 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "mt_train_backward: stencil does not fit in shared memory");
    CAIR_CUDA(cudaFuncSetAttribute(mt_train_interact_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MtGradPtrs gpz{gp(G.conv1.w), gp(G.conv2.w), gp(G.conv3.w), gp(G.conv1.b), gp(G.conv2.b), gp(G.conv3.b),
                   gp(G.conv.w),  gp(G.conv.b),  gp(G.output.w), gp(G.output.b), gp(G.alpha)};
    const unsigned grid = (unsigned)(P < (w7_smem ? 1 : 2) * kSMs ? P : (w7_smem ? 1 : 2) * kSMs);
    CAIR_LAUNCH(mt_train_interact_bwd_kernel, grid, 256, smem, s, o.cq, o.cd, q, d, dscores, o.pooled, o.argidx, t.w7, w.conv1.b,
                w.conv2.b, w.conv3.b, w.conv.w, w.output.w, w.alpha, N, Lq, Ld, C, w.nfilters, M, P, o.dcq, o.dcd, gpz, w7_smem);
  }
  // ---- channel projections ----
  CAIR_TRY(gemm_tn(o.dcq, C, o.enc_q, Hq, 0, Lq, gp(G.query_projection.w), Hq, Rq, C, Hq, s));
  CAIR_TRY(colsum(o.dcq, C, Rq, C, gp(G.query_projection.b), nullptr, s));
  CAIR_TRY(gemm_f32(gemm_dense(o.dcq, C), t.wqt, nullptr, o.denc_q, Hq, Rq, Hq, C, ACT_NONE, s));
  CAIR_TRY(gemm_tn(o.dcd, C, o.enc_d, Hd, 0, Ld, gp(G.document_projection.w), Hd, Rd, C, Hd, s));
  CAIR_TRY(colsum(o.dcd, C, Rd, C, gp(G.document_projection.b), nullptr, s));
  CAIR_TRY(gemm_f32(gemm_dense(o.dcd, C), t.wdt, nullptr, o.denc_d, Hd, Rd, Hd, C, ACT_NONE, s));
  // ---- encoders: BPTT, then the weight gradients as GEMMs over all (sequence, step) rows ----
  struct Enc {
    float *gates, *cseq, *denc, *enc, *f, *df, *wiht;
    const GemmTcW* tw_wiht;
    const int64_t* len;
    int n, L, h;
    int64_t R;
    const cair_lstm_dir *wf, *wr, *gf, *gr;
  } encs[2] = {{o.gq, o.cseq_q, o.denc_q, o.enc_q, o.fq, o.dfq, t.wiht_q, &t.tw_wiht_q, qlen, B, Lq, t.hq, Rq, &w.query_fwd, &w.query_rev, &G.query_fwd, &G.query_rev},
               {o.gd, o.cseq_d, o.denc_d, o.enc_d, o.fd, o.dfd, t.wiht_d, &t.tw_wiht_d, dlen, (int)P, Ld, t.hd, Rd, &w.doc_fwd, &w.doc_rev, &G.doc_fwd, &G.doc_rev}};
  for (const Enc& e : encs) {
    const int Gh = 4 * e.h, PG = t.dirs * Gh, Hout = t.dirs * e.h;
    CAIR_TRY(lstm_train_bwd(e.gates, e.cseq, e.denc, e.wf->w_hh, e.wr->w_hh, h->whh_scratch, e.len, e.n, e.L, e.h, t.dirs, s));
    for (int dd = 0; dd < t.dirs; ++dd) {
      const cair_lstm_dir* gw = dd ? e.gr : e.gf;
      if (!gw->w_ih || !gw->w_hh || !gw->b_ih || !gw->b_hh) return fail(CAIR_ERR_BAD_ARG, "mt_train_backward: null LSTM gradient pointer");
      const float* dg = e.gates + (size_t)dd * Gh;
      // h_{t-1} of the forward direction is the bank row before, of the reverse direction the row after
      CAIR_TRY(gemm_tn(dg, PG, e.enc + (size_t)dd * e.h, Hout, dd ? 1 : -1, e.L, gp(gw->w_hh), e.h, e.R, Gh, e.h, s));
      CAIR_TRY(gemm_tn(dg, PG, e.f, F, 0, e.L, gp(gw->w_ih), F, e.R, Gh, F, s));
      CAIR_TRY(colsum(dg, PG, e.R, Gh, gp(gw->b_ih), gp(gw->b_hh), s));
    }
    {
      const GemmTcW none{};
      CAIR_TRY(gemm_auto(gemm_dense(e.gates, PG), e.wiht, t.tc_forward ? *e.tw_wiht : none, nullptr, e.df, F, e.R, F, PG, ACT_NONE, s));
    }
  }
  // ---- linear_projection and the embedding table ----
  CAIR_TRY(gemm_tn(o.dfq, F, o.xq, E, 0, Lq, gp(G.linear_projection.w), E, Rq, F, E, s));
  CAIR_TRY(gemm_tn(o.dfd, F, o.xd, E, 0, Ld, gp(G.linear_projection.w), E, Rd, F, E, s));
  CAIR_TRY(colsum(o.dfq, F, Rq, F, gp(G.linear_projection.b), nullptr, s));
  CAIR_TRY(colsum(o.dfd, F, Rd, F, gp(G.linear_projection.b), nullptr, s));
  if (G.table) {
    const size_t smem = (size_t)8 * F * sizeof(float);
    CAIR_LAUNCH(embed_grad_kernel, 1184, 256, smem, s, o.dfq, w.linear_projection.w, q, w.vocab, E, F, Rq, (int64_t)0, p_drop, seed,
                gp(G.table));
    CAIR_LAUNCH(embed_grad_kernel, 1184, 256, smem, s, o.dfd, w.linear_projection.w, d, w.vocab, E, F, Rd, Rq, p_drop, seed,
                gp(G.table));
  }
  return CAIR_OK;
}

int32_t cair_drmm_train_workspace_bytes(int32_t emsize, int32_t B, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes) {
  if (!bytes || emsize <= 0 || B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_ARG, "drmm_train_workspace_bytes: bad argument");
  Arena a(nullptr, 0);
  DrmmTrainWs o;
  drmm_train_layout(a, emsize, B, N, Lq, Ld, &o);
  Arena fwd(nullptr, 0);
  cair_drmm_weights w{};
  w.emsize = emsize;
  CAIR_TRY(drmm_forward(w, nullptr, nullptr, N, Lq, Ld, 0, (int64_t)B * N, nullptr, nullptr, fwd, nullptr, 0, true));
  *bytes = align_up(a.off) + align_up(fwd.off) + 512;
  return CAIR_OK;
}

int32_t cair_drmm_train_forward(const cair_drmm_weights* w, const int64_t* q, const int64_t* d, int32_t B, int32_t N, int32_t Lq,
                                int32_t Ld, float p_drop, uint64_t seed, float* scores, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !q || !d || !scores || !ws || !w->table) return fail(CAIR_ERR_BAD_ARG, "drmm_train_forward: null argument");
  if (p_drop < 0.f || p_drop >= 1.f) return fail(CAIR_ERR_BAD_ARG, "drmm_train_forward: dropout must be in [0, 1)");
  if ((uintptr_t)ws % 256) return fail(CAIR_ERR_WORKSPACE, "drmm_train_forward: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  DrmmTrainWs o;
  drmm_train_layout(a, w->emsize, B, N, Lq, Ld, &o);
  a.off = align_up(a.off);
  const int64_t Rq = (int64_t)B * Lq, Rd = (int64_t)B * N * Ld;
  if (Rq + Rd >= ((int64_t)1 << 31)) return fail(CAIR_ERR_UNSUPPORTED, "drmm_train_forward: batch too large");
  {
    Arena probe = a;
    cair_drmm_weights wp = *w;
    CAIR_TRY(drmm_forward(wp, nullptr, nullptr, N, Lq, Ld, 0, (int64_t)B * N, nullptr, nullptr, probe, nullptr, 0, true));
    if (!probe.ok()) return fail(CAIR_ERR_WORKSPACE, "drmm_train_forward: workspace too small");
  }
  CAIR_CUDA(cudaMemsetAsync(o.err, 0, 256, s));
  CAIR_LAUNCH(embed_drop_kernel, 1184, 256, 0, s, w->table, q, w->vocab, w->emsize, Rq, (int64_t)0, p_drop, seed, o.vtable, o.err);
  CAIR_LAUNCH(embed_drop_kernel, 1184, 256, 0, s, w->table, d, w->vocab, w->emsize, Rd, Rq, p_drop, seed,
              o.vtable + (size_t)Rq * w->emsize, o.err);
  CAIR_LAUNCH(iota_kernel, (unsigned)((Rq + 255) / 256), 256, 0, s, o.vq, Rq, (int64_t)0);
  CAIR_LAUNCH(iota_kernel, (unsigned)((Rd + 255) / 256), 256, 0, s, o.vd, Rd, Rq);
  cair_drmm_weights wv = *w;
  wv.table = o.vtable;
  wv.vocab = (int32_t)(Rq + Rd);
  return drmm_forward(wv, o.vq, o.vd, N, Lq, Ld, 0, (int64_t)B * N, scores, o.hist, a, o.err, s, false);
}

int32_t cair_drmm_train_backward(const cair_drmm_weights* w, const cair_drmm_weights* grads, const int64_t* q, int32_t B, int32_t N,
                                 int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, const float* dscores, void* ws, size_t ws_bytes,
                                 void* stream) {
  if (!w || !grads || !q || !dscores || !ws) return fail(CAIR_ERR_BAD_ARG, "drmm_train_backward: null argument");
  const cair_drmm_weights& G = *grads;
  if (!G.gating.w || !G.gating.b || !G.ffnn0.w || !G.ffnn0.b || !G.ffnn1.w || !G.ffnn1.b || !G.output.w || !G.output.b)
    return fail(CAIR_ERR_BAD_ARG, "drmm_train_backward: null gradient pointer (only `table` may be NULL)");
  if (Lq > 32) return fail(CAIR_ERR_UNSUPPORTED, "drmm_train_backward: max_query_len > 32");
  Arena a(ws, ws_bytes);
  DrmmTrainWs o;
  drmm_train_layout(a, w->emsize, B, N, Lq, Ld, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "drmm_train_backward: workspace too small");
  CAIR_LAUNCH(drmm_train_bwd_kernel, (unsigned)B, 128, 0, (cudaStream_t)stream, o.vtable, q, o.hist, dscores, w->gating.w, w->gating.b,
              w->ffnn0.w, w->ffnn0.b, w->ffnn1.w, w->ffnn1.b, w->output.w, w->vocab, w->emsize, N, Lq, p_drop, seed,
              gp(G.gating.w), gp(G.gating.b), gp(G.ffnn0.w), gp(G.ffnn0.b), gp(G.ffnn1.w), gp(G.ffnn1.b), gp(G.output.w),
              gp(G.output.b), gp(G.table));
  return CAIR_OK;
}

/* device error word of the last forward (bad token ids / lengths): synchronises the stream */
int32_t cair_mt_train_poll_error(cair_mt_trainer* h, void* ws, void* stream) {
  if (!h || !ws) return fail(CAIR_ERR_BAD_ARG, "mt_train_poll_error: null argument");
  DevGuard g(h->t.device);
  int flags = 0;
  CAIR_CUDA(cudaMemcpyAsync(&flags, ws, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CAIR_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (flags & ERRF_BAD_TOKEN) return fail(CAIR_ERR_BAD_ARG, "token id outside [0, vocab)");
  if (flags & ERRF_BAD_LENGTH) return fail(CAIR_ERR_BAD_ARG, "sequence length outside [1, L]");
  return CAIR_OK;
}

int32_t cair_dssm_train_workspace_bytes(int32_t emsize, int32_t nhid, int32_t nout, int32_t B, int32_t N, size_t* bytes) {
  if (!bytes || emsize <= 0 || nhid <= 0 || nout <= 0 || B <= 0 || N <= 0) return fail(CAIR_ERR_BAD_ARG, "dssm_train_workspace_bytes: bad argument");
  Arena a(nullptr, 0);
  DssmTrainWs o;
  dssm_train_layout(a, emsize, nhid, nout, B, N, &o);
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_dssm_train_forward(const cair_dssm_weights* w, const int64_t* q, const int64_t* d, int32_t B, int32_t N, int32_t Lq,
                                int32_t Ld, float p_drop, uint64_t seed, float* scores, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !q || !d || !scores || !ws || !w->table || !w->query_mlp0.w || !w->query_mlp0.b || !w->query_mlp2.w || !w->query_mlp2.b ||
      !w->doc_mlp0.w || !w->doc_mlp0.b || !w->doc_mlp2.w || !w->doc_mlp2.b)
    return fail(CAIR_ERR_BAD_ARG, "dssm_train_forward: null argument");
  if (B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_ARG, "dssm_train_forward: bad shape");
  if (p_drop < 0.f || p_drop >= 1.f) return fail(CAIR_ERR_BAD_ARG, "dssm_train_forward: dropout must be in [0, 1)");
  if ((uintptr_t)ws % 256) return fail(CAIR_ERR_WORKSPACE, "dssm_train_forward: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = w->emsize, H = w->nhid, O = w->nout;
  Arena a(ws, ws_bytes);
  DssmTrainWs o;
  dssm_train_layout(a, E, H, O, B, N, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "dssm_train_forward: workspace too small");
  const int64_t P = (int64_t)B * N, R = B + P;
  CAIR_CUDA(cudaMemsetAsync(o.err, 0, 4 * sizeof(int), s));
  CAIR_LAUNCH(dssm_train_pool_kernel, (unsigned)R, 256, 0, s, w->table, q, d, w->vocab, E, B, Lq, Ld, p_drop, seed, o.x, o.arg, o.err);
  // query_mlp on rows [0, B), doc_mlp on rows [B, R)  (dssm.py:58-59)
  CAIR_TRY(gemm_f32(gemm_dense(o.x, E), w->query_mlp0.w, w->query_mlp0.b, o.h1, H, B, H, E, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(o.h1, H), w->query_mlp2.w, w->query_mlp2.b, o.y, O, B, O, H, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(o.x + (size_t)B * E, E), w->doc_mlp0.w, w->doc_mlp0.b, o.h1 + (size_t)B * H, H, P, H, E, ACT_TANH, s));
  CAIR_TRY(gemm_f32(gemm_dense(o.h1 + (size_t)B * H, H), w->doc_mlp2.w, w->doc_mlp2.b, o.y + (size_t)B * O, O, P, O, H, ACT_TANH, s));
  CAIR_LAUNCH(row_norm_kernel, (unsigned)((R + 7) / 8), 256, 0, s, o.y, O, R, o.nrm);
  CAIR_LAUNCH(cos_score_kernel, (unsigned)((P + 7) / 8), 256, 0, s, o.y, o.nrm, O, B, N, scores);
  return CAIR_OK;
}

int32_t cair_dssm_train_backward(const cair_dssm_weights* w, const cair_dssm_weights* grads, const int64_t* q, const int64_t* d,
                                 int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, const float* scores,
                                 const float* dscores, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !grads || !q || !d || !scores || !dscores || !ws) return fail(CAIR_ERR_BAD_ARG, "dssm_train_backward: null argument");
  const cair_dssm_weights& G = *grads;
  if (!G.query_mlp0.w || !G.query_mlp0.b || !G.query_mlp2.w || !G.query_mlp2.b || !G.doc_mlp0.w || !G.doc_mlp0.b || !G.doc_mlp2.w ||
      !G.doc_mlp2.b)
    return fail(CAIR_ERR_BAD_ARG, "dssm_train_backward: null gradient pointer (only `table` may be NULL: fixed embeddings)");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = w->emsize, H = w->nhid, O = w->nout;
  Arena a(ws, ws_bytes);
  DssmTrainWs o;
  dssm_train_layout(a, E, H, O, B, N, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "dssm_train_backward: workspace too small");
  int flags = 0;
  CAIR_CUDA(cudaMemcpyAsync(&flags, o.err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaStreamSynchronize(s));
  if (flags) return fail(CAIR_ERR_BAD_ARG, "dssm_train: token id outside [0, vocab)");
  const int64_t P = (int64_t)B * N, R = B + P;
  CAIR_LAUNCH(dssm_train_cos_bwd_kernel, (unsigned)B, 128, 0, s, o.y, o.nrm, scores, dscores, O, B, N, o.dy);
  const cair_linear* l0[2] = {&w->query_mlp0, &w->doc_mlp0};
  const cair_linear* l2[2] = {&w->query_mlp2, &w->doc_mlp2};
  const cair_linear* g0[2] = {&G.query_mlp0, &G.doc_mlp0};
  const cair_linear* g2[2] = {&G.query_mlp2, &G.doc_mlp2};
  float* w0t[2] = {o.w0t_q, o.w0t_d};
  float* w2t[2] = {o.w2t_q, o.w2t_d};
  for (int side = 0; side < 2; ++side) {
    const int64_t r0 = side ? B : 0, rows = side ? P : B;
    const float* x = o.x + (size_t)r0 * E;
    const float* h1 = o.h1 + (size_t)r0 * H;
    float* dy = o.dy + (size_t)r0 * O;
    float* dh1 = o.dh1 + (size_t)r0 * H;
    float* dx = o.dx + (size_t)r0 * E;
    // second Linear: db2, dW2 = dy^T h1, dh1 = dy W2
    CAIR_TRY(colsum(dy, O, rows, O, gp(g2[side]->b), nullptr, s));
    CAIR_TRY(gemm_tn(dy, O, h1, H, 0, 1, gp(g2[side]->w), H, rows, O, H, s));
    CAIR_LAUNCH(transpose_kernel, (O * H + 255) / 256, 256, 0, s, l2[side]->w, O, H, w2t[side], (int64_t)O);
    CAIR_TRY(gemm_f32(gemm_dense(dy, O), w2t[side], nullptr, dh1, H, rows, H, O, ACT_NONE, s));
    CAIR_LAUNCH(tanh_bwd_kernel, 296, 256, 0, s, dh1, h1, rows * H);
    // first Linear: db0, dW0 = dh1^T x, dx = dh1 W0
    CAIR_TRY(colsum(dh1, H, rows, H, gp(g0[side]->b), nullptr, s));
    CAIR_TRY(gemm_tn(dh1, H, x, E, 0, 1, gp(g0[side]->w), E, rows, H, E, s));
    if (G.table) {
      CAIR_LAUNCH(transpose_kernel, (H * E + 255) / 256, 256, 0, s, l0[side]->w, H, E, w0t[side], (int64_t)H);
      CAIR_TRY(gemm_f32(gemm_dense(dh1, H), w0t[side], nullptr, dx, E, rows, E, H, ACT_NONE, s));
    }
  }
  if (G.table)
    CAIR_LAUNCH(dssm_train_scatter_kernel, (unsigned)R, 256, 0, s, o.dx, o.arg, q, d, w->vocab, E, B, Lq, Ld, p_drop, seed, gp(G.table));
  return CAIR_OK;
}


int32_t cair_cdssm_train_workspace_bytes(int32_t emsize, int32_t nhid, int32_t nout, int32_t B, int32_t N, int32_t Lq, int32_t Ld,
                                         size_t* bytes) {
  if (!bytes || emsize <= 0 || nhid <= 0 || nout <= 0 || B <= 0 || N <= 0 || Lq < 5 || Ld < 5)
    return fail(CAIR_ERR_BAD_ARG, "cdssm_train_workspace_bytes: bad argument (sequences need at least 5 tokens)");
  Arena a(nullptr, 0);
  CdssmTrainWs o;
  cdssm_train_layout(a, emsize, nhid, nout, B, N, Lq, Ld, &o);
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_cdssm_train_forward(const cair_cdssm_weights* w, const int64_t* q, const int64_t* d, int32_t B, int32_t N, int32_t Lq,
                                 int32_t Ld, float p_drop, uint64_t seed, float* scores, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !q || !d || !scores || !ws || !w->table || !w->query_conv.w || !w->query_conv.b || !w->query_sem.w || !w->query_sem.b ||
      !w->doc_conv.w || !w->doc_conv.b || !w->doc_sem.w || !w->doc_sem.b)
    return fail(CAIR_ERR_BAD_ARG, "cdssm_train_forward: null argument");
  if (B <= 0 || N <= 0 || Lq < 5 || Ld < 5) return fail(CAIR_ERR_BAD_SHAPE, "cdssm_train_forward: sequences shorter than 5 tokens");
  if (p_drop < 0.f || p_drop >= 1.f) return fail(CAIR_ERR_BAD_ARG, "cdssm_train_forward: dropout must be in [0, 1)");
  if ((uintptr_t)ws % 256) return fail(CAIR_ERR_WORKSPACE, "cdssm_train_forward: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = w->emsize, H = w->nhid, O = w->nout;
  Arena a(ws, ws_bytes);
  CdssmTrainWs o;
  cdssm_train_layout(a, E, H, O, B, N, Lq, Ld, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "cdssm_train_forward: workspace too small");
  const int64_t Rq = (int64_t)B * Lq, Rd = (int64_t)B * N * Ld, Rt = Rq + Rd, P = (int64_t)B * N;
  if (Rt + 4 >= ((int64_t)1 << 31)) return fail(CAIR_ERR_UNSUPPORTED, "cdssm_train_forward: too many token rows");
  CAIR_CUDA(cudaMemsetAsync(o.err, 0, 4 * sizeof(int), s));
  CAIR_CUDA(cudaMemsetAsync(o.x + (size_t)Rt * E, 0, (size_t)4 * E * sizeof(float), s));   // the last windows read 4 rows past the end
  CAIR_LAUNCH(embed_drop_kernel, 1184, 256, 0, s, w->table, q, w->vocab, E, Rq, (int64_t)0, p_drop, seed, o.x, o.err);
  CAIR_LAUNCH(embed_drop_kernel, 1184, 256, 0, s, w->table, d, w->vocab, E, Rd, Rq, p_drop, seed, o.x + (size_t)Rq * E, o.err);
  const cair_linear* conv[2] = {&w->query_conv, &w->doc_conv};
  const cair_linear* sem[2] = {&w->query_sem, &w->doc_sem};
  for (int side = 0; side < 2; ++side) {
    const int64_t r0 = side ? Rq : 0, rows = side ? Rd : Rq, nseq = side ? P : B;
    const int L = side ? Ld : Lq;
    cdssm_merge_launch(conv[side]->w, H, E, o.w5[side], s);
    CAIR_TRY(gemm_f32(gemm_dense(o.x + (size_t)r0 * E, E), o.w5[side], conv[side]->b, o.h1 + (size_t)r0 * H, H, rows, H, 5 * E, ACT_TANH, s));
    CAIR_TRY(gemm_f32(gemm_dense(o.h1 + (size_t)r0 * H, H), sem[side]->w, sem[side]->b, o.s2 + (size_t)r0 * O, O, rows, O, H, ACT_TANH, s));
    CAIR_LAUNCH(cdssm_train_max_kernel, (unsigned)((nseq * O + 255) / 256), 256, 0, s, o.s2, O, L, r0, nseq, o.y + (size_t)(side ? B : 0) * O,
                o.arg + (size_t)(side ? B : 0) * O);
  }
  const int64_t Rs = B + P;
  CAIR_LAUNCH(row_norm_kernel, (unsigned)((Rs + 7) / 8), 256, 0, s, o.y, O, Rs, o.nrm);
  CAIR_LAUNCH(cos_score_kernel, (unsigned)((P + 7) / 8), 256, 0, s, o.y, o.nrm, O, B, N, scores);
  return CAIR_OK;
}

int32_t cair_cdssm_train_backward(const cair_cdssm_weights* w, const cair_cdssm_weights* grads, const int64_t* q, const int64_t* d,
                                  int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, const float* scores,
                                  const float* dscores, void* ws, size_t ws_bytes, void* stream) {
  if (!w || !grads || !q || !d || !scores || !dscores || !ws) return fail(CAIR_ERR_BAD_ARG, "cdssm_train_backward: null argument");
  const cair_cdssm_weights& G = *grads;
  if (!G.query_conv.w || !G.query_conv.b || !G.query_sem.w || !G.query_sem.b || !G.doc_conv.w || !G.doc_conv.b || !G.doc_sem.w ||
      !G.doc_sem.b)
    return fail(CAIR_ERR_BAD_ARG, "cdssm_train_backward: null gradient pointer (only `table` may be NULL: fixed embeddings)");
  cudaStream_t s = (cudaStream_t)stream;
  const int E = w->emsize, H = w->nhid, O = w->nout;
  Arena a(ws, ws_bytes);
  CdssmTrainWs o;
  cdssm_train_layout(a, E, H, O, B, N, Lq, Ld, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "cdssm_train_backward: workspace too small");
  int flags = 0;
  CAIR_CUDA(cudaMemcpyAsync(&flags, o.err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaStreamSynchronize(s));
  if (flags) return fail(CAIR_ERR_BAD_ARG, "cdssm_train: token id outside [0, vocab)");
  const int64_t Rq = (int64_t)B * Lq, Rd = (int64_t)B * N * Ld, Rt = Rq + Rd, P = (int64_t)B * N, Rs = B + P;
  CAIR_LAUNCH(dssm_train_cos_bwd_kernel, (unsigned)B, 128, 0, s, o.y, o.nrm, scores, dscores, O, B, N, o.dy);
  CAIR_CUDA(cudaMemsetAsync(o.dz2, 0, (size_t)Rt * O * sizeof(float), s));
  CAIR_LAUNCH(cdssm_train_route_kernel, (unsigned)((Rs * O + 255) / 256), 256, 0, s, o.dy, o.arg, O, Rs * O, o.dz2);
  CAIR_CUDA(cudaMemsetAsync(o.dh1, 0, (size_t)4 * H * sizeof(float), s));   // the 4 rows in front of the first token
  float* dh1 = o.dh1 + (size_t)4 * H;
  const cair_linear* sem[2] = {&w->query_sem, &w->doc_sem};
  const cair_linear* gconv[2] = {&G.query_conv, &G.doc_conv};
  const cair_linear* gsem[2] = {&G.query_sem, &G.doc_sem};
  for (int side = 0; side < 2; ++side) {
    const int64_t r0 = side ? Rq : 0, rows = side ? Rd : Rq;
    const float* x = o.x + (size_t)r0 * E;
    const float* h1 = o.h1 + (size_t)r0 * H;
    const float* dz2 = o.dz2 + (size_t)r0 * O;
    float* dh = dh1 + (size_t)r0 * H;
    // sem layer
    CAIR_TRY(colsum(dz2, O, rows, O, gp(gsem[side]->b), nullptr, s));
    CAIR_TRY(gemm_tn(dz2, O, h1, H, 0, 1, gp(gsem[side]->w), H, rows, O, H, s));
    CAIR_LAUNCH(transpose_kernel, (O * H + 255) / 256, 256, 0, s, sem[side]->w, O, H, o.wst[side], (int64_t)O);
    CAIR_TRY(gemm_f32(gemm_dense(dz2, O), o.wst[side], nullptr, dh, H, rows, H, O, ACT_NONE, s));
    CAIR_LAUNCH(tanh_bwd_kernel, 296, 256, 0, s, dh, h1, rows * H);
    // conv layer (merged 5-token form): bias, W5 gradient, un-merged into conv.weight
    CAIR_TRY(colsum(dh, H, rows, H, gp(gconv[side]->b), nullptr, s));
    CAIR_CUDA(cudaMemsetAsync(o.dw5[side], 0, (size_t)H * 5 * E * sizeof(float), s));
    CAIR_TRY(gemm_tn(dh, H, x, E, 0, 1, o.dw5[side], (int64_t)5 * E, rows, H, 5 * E, s));
    CAIR_LAUNCH(cdssm_train_unmerge_kernel, (unsigned)(((int64_t)H * 9 * E + 255) / 256), 256, 0, s, o.dw5[side], H, E, gp(gconv[side]->w));
    if (G.table) {
      // a window never crosses a sequence end, and the rows of the other side carry the other weights: the 4 rows in front of
      // this side's first token must read as zero - they do for the queries (zeroed above); for the documents they are the last
      // 4 query rows, which are never valid windows (t >= Lq - 4), so their dh1 is exactly zero
      CAIR_LAUNCH(cdssm_train_reorder_kernel, (unsigned)(((int64_t)E * 5 * H + 255) / 256), 256, 0, s, o.w5[side], H, E, o.w5r[side]);
      CAIR_TRY(gemm_f32(gemm_dense(dh - (size_t)4 * H, H), o.w5r[side], nullptr, o.dx + (size_t)r0 * E, E, rows, E, 5 * H, ACT_NONE, s));
    }
  }
  if (G.table) {
    CAIR_LAUNCH(embed_scatter_kernel, 1184, 256, 0, s, o.dx, q, w->vocab, E, Rq, (int64_t)0, p_drop, seed, gp(G.table));
    CAIR_LAUNCH(embed_scatter_kernel, 1184, 256, 0, s, o.dx, d, w->vocab, E, Rd, Rq, p_drop, seed, gp(G.table));
  }
  return CAIR_OK;
}


}  // extern "C"
