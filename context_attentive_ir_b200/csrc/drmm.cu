// DRMM (neuroir/rankers/drmm.py:29-84, GatingNetwork :87-98), one CTA per (query, doc) pair.
//
// Reference materialises two [BN,Lq,Ld,E] broadcasts, takes cosine_similarity, copies to the
// host and runs numpy.histogram per row.  Here: query rows are gathered + normalised into
// shared memory once, doc rows are gathered in chunks of 32 tokens, normalised in place, the
// [Lq x 32] cosine tile is a register-tiled fp32 product (fp32 FMA on purpose: the bins are
// discontinuous, so the cosine must be fp32-exact to land in the reference's bin), bin counts
// are warp-aggregated into shared memory and the gating softmax + 5->1->1 FFN + sum + output
// layer run in the epilogue.  One 4-byte store per pair.
#include "common.cuh"

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

int32_t drmm_forward(const cair_drmm_weights& w, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                     int64_t pair_begin, int64_t pair_count, float* scores, int32_t* hist_out, int* err,
                     cudaStream_t s) {
  if (pair_count <= 0) return CAIR_OK;
  if (Lq > DR_MAXLQ) return fail(CAIR_ERR_UNSUPPORTED, "drmm: max_query_len %d > %d", Lq, DR_MAXLQ);
  const int E = w.emsize;
  int ES = (E + 3) & ~3;
  if (((ES / 4) & 1) == 0) ES += 4;  // odd number of 16-byte groups per row: conflict-free LDS.128
  size_t smem = ((size_t)(Lq + DR_CHUNK) * ES + DR_MAXLQ) * sizeof(float) + (size_t)Lq * 5 * sizeof(int);
  if (smem > 220 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "drmm: emsize %d too large", E);
  CAIR_CUDA(cudaFuncSetAttribute(drmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(drmm_kernel, (unsigned)pair_count, DR_THREADS, smem, s, w.table, w.vocab, E, ES, q, d, N, Lq, Ld,
              pair_begin, w.gating.w, w.gating.b, w.ffnn0.w, w.ffnn0.b, w.ffnn1.w, w.ffnn1.b, w.output.w,
              w.output.b, scores, hist_out, err);
  return CAIR_OK;
}

}  // namespace cair
