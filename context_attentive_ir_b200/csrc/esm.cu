// Embedding gather (neuroir/modules/embeddings.py:243-252) and the ESM ranker
// (neuroir/rankers/esm.py:19-45): mean-pool over the PADDED length, cosine.  Pure HBM-bound
// gathers: 128-bit L1-bypassing loads, several rows in flight per thread, one score store per pair.
#include "common.cuh"

namespace cair {

__global__ void __launch_bounds__(256) embed_gather_kernel(const float* __restrict__ table, int V, int E,
                                                           const int64_t* __restrict__ ids, int64_t T,
                                                           float* __restrict__ out, int* err) {
  // one warp per token row; float4 when E % 4 == 0
  const int lane = threadIdx.x & 31;
  int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  int64_t id = checked_id(ids[t], V, err);
  const float* src = table + id * E;
  float* dst = out + t * E;
  if ((E & 3) == 0 && ((uintptr_t)table & 15) == 0 && ((uintptr_t)out & 15) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < E / 4; i += 32) d4[i] = ldg_stream(s4 + i);
  } else {
    for (int i = lane; i < E; i += 32) dst[i] = src[i];
  }
}

int32_t embed_gather(const float* table, int V, int E, const int64_t* ids, int64_t T, float* out, int* err,
                     cudaStream_t s) {
  if (T <= 0) return CAIR_OK;
  CAIR_LAUNCH(embed_gather_kernel, (unsigned)((T + 7) / 8), 256, 0, s, table, V, E, ids, T, out, err);
  return CAIR_OK;
}

// Sum of the table rows of `L` tokens into smem acc[E] (block-cooperative).
// Threads are laid out as (token group g, float4 column c): each thread streams rows
// t = g, g+ngroups, ... with 4 independent 128-bit loads in flight.
template <int THREADS>
__device__ void pooled_sum(const float* __restrict__ table, int V, int E, const int64_t* __restrict__ ids, int L,
                           float* part /* [ngroups][E] smem */, float* acc /* [E] smem */, int* err) {
  const int tid = threadIdx.x;
  if ((E & 3) == 0) {
    const int E4 = E >> 2;
    const int ngroups = THREADS / E4 > 0 ? THREADS / E4 : 1;
    if (E4 <= THREADS) {
      const int g = tid / E4, c = tid - g * E4;
      if (g < ngroups) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        int t = g;
        for (; t + 3 * ngroups < L; t += 4 * ngroups) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            int64_t id = checked_id(ids[t + u * ngroups], V, err);
            v[u] = ldg_stream(reinterpret_cast<const float4*>(table + id * E) + c);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) s.x += v[u].x, s.y += v[u].y, s.z += v[u].z, s.w += v[u].w;
        }
        for (; t < L; t += ngroups) {
          int64_t id = checked_id(ids[t], V, err);
          float4 v = ldg_stream(reinterpret_cast<const float4*>(table + id * E) + c);
          s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
        }
        reinterpret_cast<float4*>(part + (size_t)g * E)[c] = s;
      }
      __syncthreads();
      for (int e = tid; e < E; e += THREADS) {
        float s = 0.f;
        for (int g2 = 0; g2 < ngroups; ++g2) s += part[(size_t)g2 * E + e];
        acc[e] = s;
      }
      __syncthreads();
      return;
    }
  }
  // generic fallback (E not a multiple of 4, or very wide rows)
  for (int e = tid; e < E; e += THREADS) {
    float s = 0.f;
    for (int t = 0; t < L; ++t) s += table[checked_id(ids[t], V, err) * E + e];
    acc[e] = s;
  }
  __syncthreads();
}

template <int THREADS>
__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < THREADS / 32; ++i) s += red[i];
  return s;
}

constexpr int ESM_THREADS = 256;

// One CTA per (query, doc) pair.  smem: part[ngroups*E] | vq[E] | vd[E]
__global__ void __launch_bounds__(ESM_THREADS) esm_kernel(const float* __restrict__ table, int V, int E,
                                                          const int64_t* __restrict__ q,
                                                          const int64_t* __restrict__ d, int N, int Lq, int Ld,
                                                          int64_t pair_begin, float* __restrict__ scores,
                                                          int ngroups_alloc, int* err) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[ESM_THREADS / 32];
  float* part = sm;
  float* vq = part + (size_t)ngroups_alloc * E;
  float* vd = vq + E;
  const int64_t p = pair_begin + blockIdx.x;
  const int64_t b = p / N;
  pooled_sum<ESM_THREADS>(table, V, E, q + b * Lq, Lq, part, vq, err);
  pooled_sum<ESM_THREADS>(table, V, E, d + p * Ld, Ld, part, vd, err);
  // mean over the padded length, then torch>=2 cosine: normalise (clamped at eps) first, then dot
  float sq = 0.f, sd = 0.f;
  const float iq = 1.0f / (float)Lq, id = 1.0f / (float)Ld;
  for (int e = threadIdx.x; e < E; e += ESM_THREADS) {
    float a = vq[e] * iq, c = vd[e] * id;
    sq += a * a, sd += c * c;
  }
  float nq = fmaxf(sqrtf(block_sum<ESM_THREADS>(sq, red)), 1e-8f);
  float nd = fmaxf(sqrtf(block_sum<ESM_THREADS>(sd, red)), 1e-8f);
  float dot = 0.f;
  for (int e = threadIdx.x; e < E; e += ESM_THREADS) dot += (vq[e] * iq / nq) * (vd[e] * id / nd);
  dot = block_sum<ESM_THREADS>(dot, red);
  if (threadIdx.x == 0) scores[p] = dot;
}

int32_t esm_forward(const float* table, int V, int E, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                    int64_t pair_begin, int64_t pair_count, float* scores, int* err, cudaStream_t s) {
  if (pair_count <= 0) return CAIR_OK;
  int ngroups = 1;
  if ((E & 3) == 0 && E / 4 <= ESM_THREADS) ngroups = ESM_THREADS / (E / 4);
  size_t smem = ((size_t)ngroups * E + 2 * (size_t)E) * sizeof(float);
  if (smem > 200 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "esm: emsize %d too large", E);
  if (smem > 48 * 1024)
    CAIR_CUDA(cudaFuncSetAttribute(esm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(esm_kernel, (unsigned)pair_count, ESM_THREADS, smem, s, table, V, E, q, d, N, Lq, Ld, pair_begin,
              scores, ngroups, err);
  return CAIR_OK;
}

// ---- training step of ESM (esm.py:19-45 under models/ranker.py:192-230) --------------------
// The only parameter is the embedding table.  forward keeps the mean vectors and their norms; backward is the cosine's
// derivative per pair, reduced over the N documents of a query, then a scatter-add of dv / L into the rows of the non-PAD
// tokens (nn.Embedding padding_idx: the PAD row receives no gradient).
struct EsmTrainWs {
  float* v;     // [B + B*N, E] mean over the padded length (queries first)
  float* nrm;   // [B + B*N]    max(|v|, 1e-8)
  float* dv;    // [B + B*N, E] d loss / d v
  int* err;
};
static void esm_train_layout(Arena& a, int E, int B, int N, EsmTrainWs* o) {
  const size_t R = (size_t)B + (size_t)B * N;
  o->v = a.take<float>(R * E);
  o->nrm = a.take<float>(R);
  o->dv = a.take<float>(R * E);
  o->err = a.take<int>(4);
}

// one CTA per row (query b, or document p = r - B)
__global__ void __launch_bounds__(ESM_THREADS) esm_train_rows_kernel(const float* __restrict__ table, int V, int E,
                                                                     const int64_t* __restrict__ q, const int64_t* __restrict__ d,
                                                                     int B, int Lq, int Ld, float* __restrict__ v,
                                                                     float* __restrict__ nrm, int ngroups_alloc, int* err) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[ESM_THREADS / 32];
  float* part = sm;
  float* acc = part + (size_t)ngroups_alloc * E;
  const int64_t r = blockIdx.x;
  const bool isq = r < B;
  const int L = isq ? Lq : Ld;
  pooled_sum<ESM_THREADS>(table, V, E, isq ? q + r * Lq : d + (r - B) * Ld, L, part, acc, err);
  const float inv = 1.0f / (float)L;
  float ss = 0.f;
  for (int e = threadIdx.x; e < E; e += ESM_THREADS) {
    const float a = acc[e] * inv;
    v[r * E + e] = a;
    ss += a * a;
  }
  ss = block_sum<ESM_THREADS>(ss, red);
  if (threadIdx.x == 0) nrm[r] = fmaxf(sqrtf(ss), 1e-8f);
}

// one warp per pair: normalise first, then dot (the order of the eval kernel and of torch >= 2)
__global__ void __launch_bounds__(256) esm_train_score_kernel(const float* __restrict__ v, const float* __restrict__ nrm, int E, int B,
                                                              int N, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= (int64_t)B * N) return;
  const int64_t b = p / N;
  const float* a = v + b * E;
  const float* c = v + ((int64_t)B + p) * E;
  const float na = nrm[b], nc = nrm[B + p];
  float dot = 0.f;
  for (int e = lane; e < E; e += 32) dot += (a[e] / na) * (c[e] / nc);
  dot = warp_sum(dot);
  if (lane == 0) scores[p] = dot;
}

// one CTA per query: dv of the query (sum over its documents) and of its N documents.
// s = a_hat . c_hat, a_hat = a / max(|a|, eps):  ds/da = (c_hat - s a_hat) / |a|  (|a| >= eps; below, the clamp is constant)
__global__ void __launch_bounds__(256) esm_train_bwd_kernel(const float* __restrict__ v, const float* __restrict__ nrm,
                                                            const float* __restrict__ scores, const float* __restrict__ dscores, int E,
                                                            int B, int N, float* __restrict__ dv) {
  const int b = blockIdx.x;
  const float* a = v + (size_t)b * E;
  const float na = nrm[b];
  const bool aclamped = na <= 1e-8f;
  for (int e = threadIdx.x; e < E; e += 256) {
    const float ah = a[e] / na;
    float da = 0.f;
    for (int n = 0; n < N; ++n) {
      const int64_t p = (int64_t)b * N + n;
      const float nc = nrm[B + p], s = scores[p], g = dscores[p];
      const float ch = v[((size_t)B + p) * E + e] / nc;
      da += g * (aclamped ? ch : (ch - s * ah)) / na;
      dv[((size_t)B + p) * E + e] = g * (nc <= 1e-8f ? ah : (ah - s * ch)) / nc;
    }
    dv[(size_t)b * E + e] = da;
  }
}

// one CTA per row: d table[id] += dv[r] / L for the non-PAD tokens of the row
__global__ void __launch_bounds__(256) esm_train_scatter_kernel(const float* __restrict__ dv, const int64_t* __restrict__ q,
                                                                const int64_t* __restrict__ d, int V, int E, int B, int Lq, int Ld,
                                                                float* __restrict__ dtable) {
  const int64_t r = blockIdx.x;
  const bool isq = r < B;
  const int L = isq ? Lq : Ld;
  const int64_t* ids = isq ? q + r * Lq : d + (r - B) * Ld;
  const float inv = 1.0f / (float)L;
  for (int t = 0; t < L; ++t) {
    const int64_t id = ids[t];
    if (id <= 0 || id >= V) continue;   // PAD (0): no gradient; out-of-range ids were flagged by the forward
    for (int e = threadIdx.x; e < E; e += 256) atomicAdd(dtable + id * E + e, dv[r * E + e] * inv);
  }
}

}  // namespace cair

using namespace cair;
extern "C" {

int32_t cair_esm_train_workspace_bytes(int32_t emsize, int32_t B, int32_t N, size_t* bytes) {
  if (!bytes || emsize <= 0 || B <= 0 || N <= 0) return fail(CAIR_ERR_BAD_ARG, "esm_train_workspace_bytes: bad argument");
  Arena a(nullptr, 0);
  EsmTrainWs o;
  esm_train_layout(a, emsize, B, N, &o);
  *bytes = align_up(a.off) + 256;
  return CAIR_OK;
}

int32_t cair_esm_train_forward(const float* table, int32_t V, int32_t E, const int64_t* q, const int64_t* d, int32_t B, int32_t N,
                               int32_t Lq, int32_t Ld, float* scores, void* ws, size_t ws_bytes, void* stream) {
  if (!table || !q || !d || !scores || !ws || V <= 0 || E <= 0 || B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0)
    return fail(CAIR_ERR_BAD_ARG, "esm_train_forward: bad argument");
  if ((uintptr_t)ws % 256) return fail(CAIR_ERR_WORKSPACE, "esm_train_forward: workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  EsmTrainWs o;
  esm_train_layout(a, E, B, N, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "esm_train_forward: workspace too small");
  int ngroups = 1;
  if ((E & 3) == 0 && E / 4 <= ESM_THREADS) ngroups = ESM_THREADS / (E / 4);
  const size_t smem = ((size_t)ngroups * E + (size_t)E) * sizeof(float);
  if (smem > 200 * 1024) return fail(CAIR_ERR_UNSUPPORTED, "esm_train_forward: emsize %d too large", E);
  if (smem > 48 * 1024)
    CAIR_CUDA(cudaFuncSetAttribute(esm_train_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_CUDA(cudaMemsetAsync(o.err, 0, 4 * sizeof(int), s));
  const int64_t R = (int64_t)B + (int64_t)B * N;
  CAIR_LAUNCH(esm_train_rows_kernel, (unsigned)R, ESM_THREADS, smem, s, table, V, E, q, d, B, Lq, Ld, o.v, o.nrm, ngroups, o.err);
  CAIR_LAUNCH(esm_train_score_kernel, (unsigned)(((int64_t)B * N + 7) / 8), 256, 0, s, o.v, o.nrm, E, B, N, scores);
  return CAIR_OK;
}

int32_t cair_esm_train_backward(int32_t V, int32_t E, const int64_t* q, const int64_t* d, int32_t B, int32_t N, int32_t Lq,
                                int32_t Ld, const float* scores, const float* dscores, float* dtable, void* ws, size_t ws_bytes,
                                void* stream) {
  if (!q || !d || !scores || !dscores || !ws || V <= 0 || E <= 0 || B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0)
    return fail(CAIR_ERR_BAD_ARG, "esm_train_backward: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  EsmTrainWs o;
  esm_train_layout(a, E, B, N, &o);
  if (!a.ok()) return fail(CAIR_ERR_WORKSPACE, "esm_train_backward: workspace too small");
  int flags = 0;
  CAIR_CUDA(cudaMemcpyAsync(&flags, o.err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CAIR_CUDA(cudaStreamSynchronize(s));
  if (flags) return fail(CAIR_ERR_BAD_ARG, "esm_train: token id outside [0, vocab)");
  if (!dtable) return CAIR_OK;   // fixed embeddings: ESM has nothing else to train
  CAIR_LAUNCH(esm_train_bwd_kernel, (unsigned)B, 256, 0, s, o.v, o.nrm, scores, dscores, E, B, N, o.dv);
  const int64_t R = (int64_t)B + (int64_t)B * N;
  CAIR_LAUNCH(esm_train_scatter_kernel, (unsigned)R, 256, 0, s, o.dv, q, d, V, E, B, Lq, Ld, dtable);
  return CAIR_OK;
}

}  // extern "C"
