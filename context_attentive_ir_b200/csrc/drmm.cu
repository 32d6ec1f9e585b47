// DRMM (neuroir/rankers/drmm.py:29-84, GatingNetwork :87-98), one CTA per (query, doc) pair.
//
// Reference materialises two [BN,Lq,Ld,E] broadcasts, takes cosine_similarity, copies to the
// host and runs numpy.histogram per row.  Here: query rows are gathered + normalised into
// shared memory once, doc rows are gathered in chunks of 32 tokens, normalised in place, the
// [Lq x 32] cosine tile is a register-tiled fp32 product (fp32 FMA on purpose: the bins are
// discontinuous, so the cosine must be fp32-exact to land in the reference's bin), bin counts
// are warp-aggregated into shared memory and the gating softmax + 5->1->1 FFN + sum + output
// layer run in the epilogue.  One 4-byte store per pair.
#include "models.cuh"

namespace cair {

constexpr int DR_THREADS = 256;
constexpr int DR_CHUNK = 32;   // doc tokens per chunk == warp width (lane <-> token in the dot phase)
constexpr int DR_MAXLQ = 32;   // 8 warps x up to 4 query rows each

__device__ __forceinline__ int drmm_bin(float c) {
  // numpy.histogram(bins=[-1,-.5,0,.5,1,1]): half-open bins, last bin closed ({1.0}), outside dropped
  if (!(c >= -1.0f) || c > 1.0f) return -1;
  if (c == 1.0f) return 4;
  if (c < -0.5f) return 0;
  if (c < 0.0f) return 1;
  if (c < 0.5f) return 2;
  return 3;
}

// smem: qn[Lq][ES] | dn[32][ES] | gate[Lq] | hist[Lq][5]     ES = row stride (floats), ES/4 odd
__global__ void __launch_bounds__(DR_THREADS) drmm_kernel(const float* __restrict__ table, int V, int E, int ES,
                                                          const int64_t* __restrict__ q,
                                                          const int64_t* __restrict__ d, int N, int Lq, int Ld,
                                                          int64_t pair_begin, const float* __restrict__ wg,
                                                          const float* __restrict__ bg, const float* __restrict__ w0,
                                                          const float* __restrict__ b0, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, const float* __restrict__ wo,
                                                          const float* __restrict__ bo, float* __restrict__ scores,
                                                          int32_t* __restrict__ hist_out, int* err) {
  extern __shared__ __align__(16) float sm[];
  float* qn = sm;
  float* dn = qn + (size_t)Lq * ES;
  float* gate = dn + (size_t)DR_CHUNK * ES;
  int* hist = reinterpret_cast<int*>(gate + DR_MAXLQ);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t p = pair_begin + blockIdx.x;
  const int64_t b = p / N;
  const bool vec = (E & 3) == 0;

  for (int i = tid; i < Lq * 5; i += DR_THREADS) hist[i] = 0;
  // ---- query rows: gather, gate logit on the raw row, normalise by max(||x||, eps) ----
  for (int i = warp; i < Lq; i += DR_THREADS / 32) {
    int64_t id = checked_id(q[b * Lq + i], V, err);
    const float* src = table + id * E;
    float ss = 0.f, gl = 0.f;
    for (int e = lane; e < E; e += 32) {
      float v = src[e];
      qn[(size_t)i * ES + e] = v;
      ss += v * v;
      gl += v * wg[e];
    }
    ss = warp_sum(ss);
    gl = warp_sum(gl);
    float inv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
    for (int e = lane; e < E; e += 32) qn[(size_t)i * ES + e] *= inv;
    for (int e = E + lane; e < ES; e += 32) qn[(size_t)i * ES + e] = 0.f;
    if (lane == 0) gate[i] = gl + bg[0];
  }
  __syncthreads();

  // rows of the cosine tile owned by this warp: i = warp, warp+8, warp+16, warp+24
  const int NW = DR_THREADS / 32;
  for (int j0 = 0; j0 < Ld; j0 += DR_CHUNK) {
    const int nj = min(DR_CHUNK, Ld - j0);
    // ---- gather + normalise up to 32 doc rows (4 rows per warp, loads of all 4 in flight) ----
    for (int jj = warp; jj < DR_CHUNK; jj += NW) {
      float* dst = dn + (size_t)jj * ES;
      if (jj < nj) {
        int64_t id = checked_id(d[p * Ld + j0 + jj], V, err);
        float ss = 0.f;
        if (vec) {
          const float4* s4 = reinterpret_cast<const float4*>(table + id * E);
          for (int e4 = lane; e4 < E / 4; e4 += 32) {
            float4 v = ldg_stream(s4 + e4);
            reinterpret_cast<float4*>(dst)[e4] = v;
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          }
        } else {
          for (int e = lane; e < E; e += 32) {
            float v = table[id * E + e];
            dst[e] = v;
            ss += v * v;
          }
        }
        ss = warp_sum(ss);
        float inv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
        for (int e = lane; e < E; e += 32) dst[e] *= inv;
        for (int e = E + lane; e < ES; e += 32) dst[e] = 0.f;
      } else {
        for (int e = lane; e < ES; e += 32) dst[e] = 0.f;
      }
    }
    __syncthreads();
    // ---- cosine tile: lane <-> doc token, warp <-> up to 4 query rows ----
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* drow = reinterpret_cast<const float4*>(dn + (size_t)lane * ES);
    for (int k4 = 0; k4 < ES / 4; ++k4) {
      float4 dv = drow[k4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int i = warp + u * NW;
        if (i < Lq) {
          float4 qv = reinterpret_cast<const float4*>(qn + (size_t)i * ES)[k4];
          acc[u] = fmaf(qv.x, dv.x, acc[u]);
          acc[u] = fmaf(qv.y, dv.y, acc[u]);
          acc[u] = fmaf(qv.z, dv.z, acc[u]);
          acc[u] = fmaf(qv.w, dv.w, acc[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int i = warp + u * NW;
      if (i >= Lq) continue;
      int bin = (lane < nj) ? drmm_bin(acc[u]) : -1;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        unsigned m = __ballot_sync(0xffffffffu, bin == k);
        if (lane == 0 && m) hist[i * 5 + k] += __popc(m);  // row i is owned by this warp only
      }
    }
    __syncthreads();
  }
  // ---- epilogue: softmax gate over ALL Lq positions, ffnn(5->1->1), weighted sum, output ----
  if (warp == 0) {
    float g = (lane < Lq) ? gate[lane] : -INFINITY;
    float mx = warp_max(g);
    float ex = (lane < Lq) ? __expf(g - mx) : 0.f;
    float den = warp_sum(ex);
    float f = 0.f;
    if (lane < Lq) {
      float f0 = b0[0];
#pragma unroll
      for (int k = 0; k < 5; ++k) f0 = fmaf(w0[k], (float)hist[lane * 5 + k], f0);
      f = (w1[0] * f0 + b1[0]) * (ex / den);
    }
    f = warp_sum(f);
    if (lane == 0) scores[p] = wo[0] * f + bo[0];
  }
  if (hist_out)
    for (int i = tid; i < Lq * 5; i += DR_THREADS) hist_out[p * Lq * 5 + i] = hist[i];
}

// acc.{x,y} = fma(w.{x,y}, y, acc.{x,y}) as one packed instruction (fma.rn.f32x2: two independent IEEE FMAs)
__device__ __forceinline__ void ffma2_d(float2& acc, float2 w, float y) {
  unsigned long long a, c, yy;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(w.x), "f"(w.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc.x), "f"(acc.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(yy));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c));
}

// ---- register-tiled variant (E % 4 == 0, Lq <= 20, Ld <= 200) --------------------------------------------------
// drmm_kernel is bound by shared-memory wavefronts: one thread = one document token x 4 query rows needs 5 LDS.128
// per 16 FMAs.  Here a thread owns a 4 x 4 tile (4 query rows x 4 tokens: 8 LDS.128 per 64 FMAs), which needs all
// Ld tokens of the pair side by side; whole normalised rows of all tokens do not fit in shared memory, so the K
// dimension is chunked instead: pass 1 computes the row norms (the rows come from HBM once), pass 2 streams the rows
// again in 64-float K chunks (from L2), normalises them on the way into shared memory and accumulates.
// Every cell is still the SAME sequential fp32 FMA chain over k (k ascending, x, y, z, w), on the same normalised
// values (v * inv, inv from the same lane-strided partial sums + butterfly), as in drmm_kernel: the bins are identical.
constexpr int D2_TOK = 200;        // tokens per pair held side by side (Ld <= D2_TOK)
constexpr int D2_TG = D2_TOK / 4;  // token groups: thread tg owns tokens tg, tg + 50, tg + 100, tg + 150
constexpr int D2_QG = 5;           // query-row groups of 4 (Lq <= 20)
constexpr int D2_KC = 8;           // float4 per K chunk
constexpr int D2_RS = D2_KC + 1;   // row stride of the chunk tile in float4 (odd: conflict-free LDS.128 across tokens)

// smem: qn[20][ES] | qn2[10][ES][2] (row pairs interleaved) | dt[D2_TOK][D2_RS] float4 | inv[D2_TOK] | ids[D2_TOK] (int) |
//       gate[32] | hist[20*5] (int)
__global__ void __launch_bounds__(DR_THREADS, 2)
    drmm2_kernel(const float* __restrict__ table, int V, int E, int ES, const int64_t* __restrict__ q,
                 const int64_t* __restrict__ d, int N, int Lq, int Ld, int64_t pair_begin, const float* __restrict__ wg,
                 const float* __restrict__ bg, const float* __restrict__ w0, const float* __restrict__ b0,
                 const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ wo,
                 const float* __restrict__ bo, float* __restrict__ scores, int32_t* __restrict__ hist_out, int* err) {
  extern __shared__ __align__(16) float sm[];
  float* qn = sm;
  float* qn2 = qn + (size_t)20 * ES;
  float4* dt = reinterpret_cast<float4*>(qn2 + (size_t)20 * ES);
  float* inv = reinterpret_cast<float*>(dt + (size_t)D2_TOK * D2_RS);
  int* ids = reinterpret_cast<int*>(inv + D2_TOK);
  float* gate = reinterpret_cast<float*>(ids + D2_TOK);
  int* hist = reinterpret_cast<int*>(gate + DR_MAXLQ);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t p = pair_begin + blockIdx.x;
  const int64_t b = p / N;
  const int E4 = E >> 2, ES4 = ES >> 2;

  for (int i = tid; i < 20 * 5; i += DR_THREADS) hist[i] = 0;
  // ---- query rows: gather, gate logit on the raw row, normalise by max(||x||, eps) (identical to drmm_kernel) ----
  for (int i = warp; i < 20; i += DR_THREADS / 32) {
    if (i < Lq) {
      int64_t id = checked_id(q[b * Lq + i], V, err);
      const float* src = table + id * E;
      float ss = 0.f, gl = 0.f;
      for (int e = lane; e < E; e += 32) {
        float v = src[e];
        qn[(size_t)i * ES + e] = v;
        ss += v * v;
        gl += v * wg[e];
      }
      ss = warp_sum(ss);
      gl = warp_sum(gl);
      float iv = 1.0f / fmaxf(sqrtf(ss), 1e-8f);
      for (int e = lane; e < E; e += 32) qn[(size_t)i * ES + e] *= iv;
      for (int e = E + lane; e < ES; e += 32) qn[(size_t)i * ES + e] = 0.f;
      if (lane == 0) gate[i] = gl + bg[0];
    } else {
      for (int e = lane; e < ES; e += 32) qn[(size_t)i * ES + e] = 0.f;   // rows beyond Lq: zero (their cells are never counted)
    }
  }
  // ---- pass 1: token ids and row norms (same lane-strided partial sums + butterfly as drmm_kernel) ----
  // 5 rows per warp iteration: all their loads are issued before the first reduction (the gather is latency-bound)
  constexpr int P1 = 5;
  for (int j0 = warp * P1; j0 < D2_TOK; j0 += (DR_THREADS / 32) * P1) {
    int64_t id[P1];
    float4 v[P1][3];
#pragma unroll
    for (int r = 0; r < P1; ++r) id[r] = (j0 + r < Ld) ? checked_id(d[p * Ld + j0 + r], V, err) : 0;
#pragma unroll
    for (int r = 0; r < P1; ++r) {
      const float4* s4 = reinterpret_cast<const float4*>(table + id[r] * E);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int e4 = lane + 32 * c;
        v[r][c] = (j0 + r < Ld && e4 < E4) ? ldg_stream(s4 + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < P1; ++r) {
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (lane + 32 * c < E4) ss += v[r][c].x * v[r][c].x + v[r][c].y * v[r][c].y + v[r][c].z * v[r][c].z + v[r][c].w * v[r][c].w;
      for (int e4 = lane + 96; e4 < E4; e4 += 32) {   // rows longer than 384 floats
        const float4 x = ldg_stream(reinterpret_cast<const float4*>(table + id[r] * E) + e4);
        if (j0 + r < Ld) ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      }
      ss = warp_sum(ss);
      if (lane == 0 && j0 + r < D2_TOK) {
        inv[j0 + r] = (j0 + r < Ld) ? 1.0f / fmaxf(sqrtf(ss), 1e-8f) : 0.f;
        ids[j0 + r] = (int)id[r];
      }
    }
  }
  __syncthreads();

  for (int idx = tid; idx < 10 * ES; idx += DR_THREADS) {   // interleave the normalised query rows in pairs
    const int r = idx / ES, k = idx - r * ES;
    qn2[(size_t)idx * 2] = qn[(size_t)(2 * r) * ES + k];
    qn2[(size_t)idx * 2 + 1] = qn[(size_t)(2 * r + 1) * ES + k];
  }
  __syncthreads();
  // ---- pass 2: K chunks ----
  const int tg = tid % D2_TG, qg = tid / D2_TG;   // qg == 5 for the last 6 threads: loaders only
  float2 acc2[2][4];   // [row pair][token]: .x = row 4qg + 2up, .y = row 4qg + 2up + 1
#pragma unroll
  for (int up = 0; up < 2; ++up)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc2[up][t] = make_float2(0.f, 0.f);
  // Software pipeline through registers: the gather of chunk c+1 is issued before the dot product of chunk c, so its L2
  // latency is hidden behind the FMAs; all loads of a thread are issued before its first store.  A thread's P2 gather
  // slots (token, float4 column) are the same in every chunk: row pointers, tile offsets and 1/norm are computed once.
  constexpr int P2 = (D2_TOK * D2_KC + DR_THREADS - 1) / DR_THREADS;   // 7
  const float4* gsrc[P2];
  int goff[P2];
  float ginv[P2];
#pragma unroll
  for (int it = 0; it < P2; ++it) {
    const int idx = tid + it * DR_THREADS;
    const int j = idx / D2_KC, f4 = idx - j * D2_KC;
    const bool ok = idx < D2_TOK * D2_KC && j < Ld;
    gsrc[it] = ok ? reinterpret_cast<const float4*>(table + (size_t)ids[j] * E) + f4 : nullptr;
    goff[it] = idx < D2_TOK * D2_KC ? j * D2_RS + f4 : -1;
    ginv[it] = ok ? inv[j] : 0.f;
  }
  float4 lv[P2];
  auto gather = [&](int k0) {
#pragma unroll
    for (int it = 0; it < P2; ++it) {
      const int f4 = (tid + it * DR_THREADS) % D2_KC;
      lv[it] = (gsrc[it] && k0 + f4 < E4) ? gsrc[it][k0] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  gather(0);
  // the thread's two query-row pairs, interleaved in shared memory as (row 2r, row 2r+1) per k: one packed FFMA2 updates
  // the same cell of both rows (fma.rn.f32x2 = two independent IEEE FMAs: every cell keeps its sequential chain over k)
  const float2* qp = reinterpret_cast<const float2*>(qn2) + (size_t)(2 * qg) * ES;
  for (int k0 = 0; k0 < ES4; k0 += D2_KC) {
    const int nk = min(D2_KC, ES4 - k0);
    // normalised chunk of every token: dt[token][f4] = table[id][k0 + f4] * inv (zeros beyond E, zero rows beyond Ld)
#pragma unroll
    for (int it = 0; it < P2; ++it) {
      if (goff[it] >= 0) {
        const float iv = ginv[it];
        float4 v = lv[it];
        v.x *= iv, v.y *= iv, v.z *= iv, v.w *= iv;
        dt[goff[it]] = v;
      }
    }
    __syncthreads();
    if (k0 + D2_KC < ES4) gather(k0 + D2_KC);
    if (qg < D2_QG) {
#pragma unroll 2
      for (int f4 = 0; f4 < nk; ++f4) {
        float4 dv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) dv[t] = dt[(size_t)(tg + D2_TG * t) * D2_RS + f4];
        const int k = (k0 + f4) * 4;
#pragma unroll
        for (int up = 0; up < 2; ++up) {
          // (row 4qg+2up, row 4qg+2up+1) at k .. k+3: two 16-byte loads
          const float4 qa = *reinterpret_cast<const float4*>(qp + (size_t)up * ES + k);
          const float4 qb = *reinterpret_cast<const float4*>(qp + (size_t)up * ES + k + 2);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            ffma2_d(acc2[up][t], make_float2(qa.x, qa.y), dv[t].x);
            ffma2_d(acc2[up][t], make_float2(qa.z, qa.w), dv[t].y);
            ffma2_d(acc2[up][t], make_float2(qb.x, qb.y), dv[t].z);
            ffma2_d(acc2[up][t], make_float2(qb.z, qb.w), dv[t].w);
          }
        }
      }
    }
    __syncthreads();
  }
  float acc[4][4];
#pragma unroll
  for (int up = 0; up < 2; ++up)
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[2 * up][t] = acc2[up][t].x, acc[2 * up + 1][t] = acc2[up][t].y;
  // ---- histograms ----
  if (qg < D2_QG) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = 4 * qg + u;
      if (i >= Lq) continue;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = tg + D2_TG * t;
        const int bin = j < Ld ? drmm_bin(acc[u][t]) : -1;
        if (bin >= 0) atomicAdd(&hist[i * 5 + bin], 1);
      }
    }
  }
  __syncthreads();
  // ---- epilogue: softmax gate over ALL Lq positions, ffnn(5->1->1), weighted sum, output (identical to drmm_kernel) ----
  if (warp == 0) {
    float g = (lane < Lq) ? gate[lane] : -INFINITY;
    float mx = warp_max(g);
    float ex = (lane < Lq) ? __expf(g - mx) : 0.f;
    float den = warp_sum(ex);
    float f = 0.f;
    if (lane < Lq) {
      float f0 = b0[0];
#pragma unroll
      for (int k = 0; k < 5; ++k) f0 = fmaf(w0[k], (float)hist[lane * 5 + k], f0);
      f = (w1[0] * f0 + b1[0]) * (ex / den);
    }
    f = warp_sum(f);
    if (lane == 0) scores[p] = wo[0] * f + bo[0];
  }
  if (hist_out)
    for (int i = tid; i < Lq * 5; i += DR_THREADS) hist_out[p * Lq * 5 + i] = hist[i];
}

int32_t drmm_forward(const cair_drmm_weights& w, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                     int64_t pair_begin, int64_t pair_count, float* scores, int32_t* hist_out, Arena& ws, int* err,
                     cudaStream_t s, bool dry) {
  const int E = w.emsize;
  const int64_t nq = pair_count > 0 ? (pair_begin + pair_count - 1) / N - pair_begin / N + 1 : 0;
  uint8_t* qrec = ws.take<uint8_t>(drmm_tc_workspace_bytes(E, nq));   // reserved whatever the engine: the size must not depend on it
  if (dry || pair_count <= 0) return CAIR_OK;
  if (!ws.ok()) return fail(CAIR_ERR_WORKSPACE, "drmm: workspace too small");
  if (Lq > DR_MAXLQ) return fail(CAIR_ERR_UNSUPPORTED, "drmm: max_query_len %d > %d", Lq, DR_MAXLQ);
  if (drmm_tc_usable(E, Lq, Ld, w.table, w.vocab))   // tensor-core cosines, one HBM pass, exact recompute at the bin edges
    return drmm_tc_forward(w, q, d, N, Lq, Ld, pair_begin, pair_count, scores, hist_out, qrec, err, s);
  int ES = (E + 3) & ~3;
  if (((ES / 4) & 1) == 0) ES += 4;  // odd number of 16-byte groups per row: conflict-free LDS.128
  if ((E & 3) == 0 && Lq <= 4 * D2_QG && Ld <= D2_TOK && ((uintptr_t)w.table & 15) == 0) {
    const size_t smem2 = (size_t)40 * ES * sizeof(float) + (size_t)D2_TOK * D2_RS * sizeof(float4) +
                         (size_t)D2_TOK * (sizeof(float) + sizeof(int)) + DR_MAXLQ * sizeof(float) + 20 * 5 * sizeof(int);
    CAIR_CUDA(cudaFuncSetAttribute(drmm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    CAIR_LAUNCH(drmm2_kernel, (unsigned)pair_count, DR_THREADS, smem2, s, w.table, w.vocab, E, ES, q, d, N, Lq, Ld,
                pair_begin, w.gating.w, w.gating.b, w.ffnn0.w, w.ffnn0.b, w.ffnn1.w, w.ffnn1.b, w.output.w,
                w.output.b, scores, hist_out, err);
    return CAIR_OK;
  }
  size_t smem = ((size_t)(Lq + DR_CHUNK) * ES + DR_MAXLQ) * sizeof(float) + (size_t)Lq * 5 * sizeof(int);
  if (smem > 220 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "drmm: emsize %d too large", E);
  CAIR_CUDA(cudaFuncSetAttribute(drmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(drmm_kernel, (unsigned)pair_count, DR_THREADS, smem, s, w.table, w.vocab, E, ES, q, d, N, Lq, Ld,
              pair_begin, w.gating.w, w.gating.b, w.ffnn0.w, w.ffnn0.b, w.ffnn1.w, w.ffnn1.b, w.output.w,
              w.output.b, scores, hist_out, err);
  return CAIR_OK;
}

}  // namespace cair
