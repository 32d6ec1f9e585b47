// Ranking metrics of the reference's evaluation loops, on the device: removes the per-batch
// scores.cpu() + numpy argsort + Python metric loops of main/ranker.py:257-264 / main/multitask.py:286-293.
//   scores -> softmax over the N candidates (models/ranker.py:257-258) -> predictions = argsort(-probs)
//   -> average precision (eval/ltorank.py:4-26), reciprocal rank (:104-123), precision@1/3/5 (:29-47).
// One CTA per query row.  Instead of sorting, every candidate gets its rank = number of candidates that precede it in
// the stable descending order (higher probability, or equal probability and lower index; numpy's default
// argsort is unstable, so the reference's order inside an exact tie is implementation-defined); the metrics are then sums over the relevant candidates.
// Label semantics are the reference's: MAP / MRR count labels == 1, precision@k counts NON-ZERO labels.
#include "common.cuh"

namespace cair {

constexpr int RM_THREADS = 128;

__device__ __forceinline__ float rm_block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < RM_THREADS / 32; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ double rm_block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int w = 1; w < RM_THREADS / 32; ++w) r += red[w];   // fixed order: deterministic
  __syncthreads();
  return r;
}

// smem: p[N] float | rank[N] int | lab[N] int
__global__ void __launch_bounds__(RM_THREADS) rank_metrics_kernel(const float* __restrict__ scores,
                                                                   const int64_t* __restrict__ labels, int N,
                                                                   int apply_softmax, double* __restrict__ per_row) {
  extern __shared__ __align__(16) unsigned char rm_smem[];
  __shared__ float redf[RM_THREADS / 32];
  __shared__ double redd[RM_THREADS / 32];
  float* p = reinterpret_cast<float*>(rm_smem);
  int* rank = reinterpret_cast<int*>(p + N);
  int* lab = rank + N;
  const int tid = threadIdx.x;
  const int64_t row = blockIdx.x;
  const float* x = scores + row * N;
  float mx = -INFINITY;
  for (int i = tid; i < N; i += RM_THREADS) {
    const float v = x[i];
    p[i] = v;
    const int64_t l = labels[row * N + i];
    lab[i] = l == 1 ? 1 : (l != 0 ? 2 : 0);   // 1: relevant for MAP / MRR and precision; 2: counts for precision@k only
    mx = fmaxf(mx, v);
  }
  __syncthreads();
  if (apply_softmax) {
    mx = rm_block_max(mx, redf);
    float s = 0.f;
    for (int i = tid; i < N; i += RM_THREADS) {
      const float e = expf(p[i] - mx);
      p[i] = e;
      s += e;
    }
    const float tot = (float)rm_block_sum((double)s, redd);
    for (int i = tid; i < N; i += RM_THREADS) p[i] = p[i] / tot;
    __syncthreads();
  }
  for (int i = tid; i < N; i += RM_THREADS) {
    const float pi = p[i];
    int r = 0;
    for (int j = 0; j < N; ++j) {
      const float pj = p[j];
      r += (pj > pi) || (pj == pi && j < i);
    }
    rank[i] = r;
  }
  __syncthreads();
  // sums over the relevant candidates
  double ap = 0.0, nrel = 0.0, p1 = 0.0, p3 = 0.0, p5 = 0.0;
  int best = N;   // rank of the first relevant candidate
  for (int i = tid; i < N; i += RM_THREADS) {
    const int li = lab[i], ri = rank[i];
    if (li) {
      p1 += ri < 1, p3 += ri < 3, p5 += ri < 5;
    }
    if (li == 1) {
      int cnt = 0;   // relevant candidates at or before this one
      for (int j = 0; j < N; ++j) cnt += (lab[j] == 1) && (rank[j] <= ri);
      ap += (double)cnt / (double)(ri + 1);
      nrel += 1.0;
      best = min(best, ri);
    }
  }
  ap = rm_block_sum(ap, redd);
  nrel = rm_block_sum(nrel, redd);
  p1 = rm_block_sum(p1, redd);
  p3 = rm_block_sum(p3, redd);
  p5 = rm_block_sum(p5, redd);
  const float bestf = -rm_block_max(-(float)best, redf);
  if (tid == 0) {
    double* o = per_row + row * 5;
    o[0] = ap / nrel;                                  // no relevant document: NaN (the reference raises)
    o[1] = bestf < (float)N ? 1.0 / ((double)bestf + 1.0) : 0.0;
    o[2] = p1, o[3] = p3 / 3.0, o[4] = p5 / 5.0;
  }
}

// batch means in a fixed order (one CTA): mean[c] = sum_rows per_row[r][c] / B
__global__ void __launch_bounds__(RM_THREADS) rank_metrics_mean_kernel(const double* __restrict__ per_row, int B,
                                                                        double* __restrict__ mean) {
  __shared__ double redd[RM_THREADS / 32];
  for (int c = 0; c < 5; ++c) {
    double s = 0.0;
    for (int r = threadIdx.x; r < B; r += RM_THREADS) s += per_row[(size_t)r * 5 + c];
    s = rm_block_sum(s, redd);
    if (threadIdx.x == 0) mean[c] = s / (double)B;
  }
}

}  // namespace cair

extern "C" CAIR_API int32_t cair_rank_metrics(const float* scores, const int64_t* labels, int32_t B, int32_t N,
                                              int32_t apply_softmax, double* per_row, double* batch_mean, void* stream) {
  using namespace cair;
  if (!scores || !labels || !per_row) return fail(CAIR_ERR_BAD_ARG, "rank_metrics: null tensor");
  if (B <= 0 || N <= 0) return fail(CAIR_ERR_BAD_SHAPE, "rank_metrics: B and N must be positive");
  if (N < 5) return fail(CAIR_ERR_BAD_SHAPE, "rank_metrics: precision@5 needs at least 5 candidates (eval/ltorank.py:41)");
  if (N > 4096) return fail(CAIR_ERR_UNSUPPORTED, "rank_metrics: at most 4096 candidates per query");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)N * 12;
  if (smem > 32 * 1024) CAIR_CUDA(cudaFuncSetAttribute(rank_metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CAIR_LAUNCH(rank_metrics_kernel, (unsigned)B, RM_THREADS, smem, s, scores, labels, N, apply_softmax, per_row);
  if (batch_mean) CAIR_LAUNCH(rank_metrics_mean_kernel, 1, RM_THREADS, 0, s, per_row, B, batch_mean);
  return CAIR_OK;
}

// ---- batchify on the device -------------------------------------------------------
// The reference pads every example on the host with per-example copy_ loops (inputters/ranker/vector.py:39-90) and
// ships padded int64 tensors.  Here the host ships the RAGGED batch (concatenated int32 token ids + int64 offsets,
// 4 bytes per real token) and one kernel writes the padded int64 [B,Lq] / [B,N,Ld] id tensors and the length tensors
// the scoring entry points take: PAD (0) beyond each length, lengths = token counts.
namespace cair {
__global__ void batchify_kernel(const int32_t* __restrict__ tokens, const int64_t* __restrict__ offsets, int64_t nseq, int L,
                                int64_t* __restrict__ ids, int64_t* __restrict__ lens, int* err) {
  const int64_t total = nseq * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / L;
    const int l = (int)(i - s * L);
    const int64_t o0 = offsets[s], n = offsets[s + 1] - o0;
    if (l == 0) {
      lens[s] = n;
      if (n < 1 || n > L) atomicOr(err, ERRF_BAD_LENGTH);   // the reference's copy_ raises on a longer example
    }
    ids[i] = (l < n) ? (int64_t)tokens[o0 + l] : 0;
  }
}
}  // namespace cair

extern "C" CAIR_API int32_t cair_batchify_ranker(const int32_t* q_tokens, const int64_t* q_offsets, const int32_t* d_tokens,
                                                 const int64_t* d_offsets, int32_t B, int32_t N, int32_t Lq, int32_t Ld,
                                                 int64_t* q, int64_t* qlen, int64_t* d, int64_t* dlen, int32_t* err_flag,
                                                 void* stream) {
  using namespace cair;
  if (!q_tokens || !q_offsets || !d_tokens || !d_offsets || !q || !qlen || !d || !dlen || !err_flag)
    return fail(CAIR_ERR_BAD_ARG, "batchify_ranker: null tensor");
  if (B <= 0 || N <= 0 || Lq <= 0 || Ld <= 0) return fail(CAIR_ERR_BAD_SHAPE, "batchify_ranker: B, N, Lq, Ld must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t tq = (int64_t)B * Lq, td = (int64_t)B * N * Ld;
  const auto blocks = [](int64_t t) { return (unsigned)((t + 255) / 256 < 148 * 16 ? (t + 255) / 256 : 148 * 16); };
  CAIR_LAUNCH(batchify_kernel, blocks(tq), 256, 0, s, q_tokens, q_offsets, (int64_t)B, Lq, q, qlen, err_flag);
  CAIR_LAUNCH(batchify_kernel, blocks(td), 256, 0, s, d_tokens, d_offsets, (int64_t)B * N, Ld, d, dlen, err_flag);
  return CAIR_OK;
}
