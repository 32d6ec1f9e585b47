// Shared helpers of libcair.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cair.h"

namespace cair {

extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

int32_t fail(int32_t code, const char* fmt, ...);

#define CAIR_CUDA(expr)                                                                    \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return ::cair::fail(CAIR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,       \
                          cudaGetErrorString(e_));                                         \
  } while (0)

#define CAIR_TRY(expr)                 \
  do {                                 \
    int32_t rc_ = (expr);              \
    if (rc_ != CAIR_OK) return rc_;    \
  } while (0)

// Every kernel launch of the library goes through this: counts it (cair_launch_count) and
// surfaces launch-configuration errors immediately.
#define CAIR_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                            \
    ::cair::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    cudaError_t e_ = cudaGetLastError();                                                   \
    if (e_ != cudaSuccess)                                                                 \
      return ::cair::fail(CAIR_ERR_CUDA, "%s:%d launch %s: %s", __FILE__, __LINE__,        \
                          #kernel, cudaGetErrorString(e_));                                \
  } while (0)

constexpr int kSMs = 148;  // B200

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace (forward never allocates).
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
  template <typename T>
  T* take(size_t n) {
    size_t o = align_up(off);
    off = o + n * sizeof(T);
    return (T*)(base ? base + o : nullptr);
  }
  bool ok() const { return off <= cap; }
};

// Device buffers owned by a handle.
struct Owned {
  std::vector<void*> ptrs;
  template <typename T>
  cudaError_t alloc(T** out, size_t n) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T) > 0 ? n * sizeof(T) : 16);
    if (e == cudaSuccess) ptrs.push_back(p);
    *out = (T*)p;
    return e;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
  }
};

enum { ERRF_BAD_TOKEN = 1, ERRF_BAD_LENGTH = 2 };

// Optional stage timing (cair_profile_*): CUDA events recorded on the launching stream between the
// stages of a forward; interval i runs from mark i to mark i+1.  Off by default (no events recorded).
struct Profiler {
  bool on = false;
  int cursor = 0;
  std::vector<std::string> names;
  std::vector<cudaEvent_t> ev;
  void reset() { cursor = 0; }
  void mark(const char* name, cudaStream_t s) {
    if (!on) return;
    if (cursor == (int)ev.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      ev.push_back(e);
      names.push_back(name);
    }
    names[cursor] = name;
    cudaEventRecord(ev[cursor++], s);
  }
  void release() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear();
    names.clear();
  }
};
extern thread_local Profiler* g_prof;
inline void prof_mark(const char* name, cudaStream_t s) {
  if (g_prof) g_prof->mark(name, s);
}

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// Token id -> table row, flagging ids outside [0, V) (the reference raises IndexError there,
// modules/embeddings.py:165 nn.Embedding) and substituting the PAD row.
__device__ __forceinline__ int64_t checked_id(int64_t id, int V, int* err) {
  if (id < 0 || id >= V) {
    if (err) atomicOr(err, ERRF_BAD_TOKEN);
    return 0;
  }
  return id;
}

// 128-bit streaming load that does not allocate in L1 (gathered rows are touched once).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace cair

// ---- internal cross-file API -----------------------------------------------------------------
namespace cair {

enum Act { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2 };

// A-row providers of the generic GEMM  C[M,N] = act(A[M,K] W[N,K]^T + bias).
struct GemmA {
  const float* dense;   // row r at dense + r*lda (when table == nullptr)
  int64_t lda;
  const float* table;   // gathered: row r = concat_{k<win} table[ids[seq*L + t + k - pad]], r = seq*T + t
  const int64_t* ids;
  int V, E, win, L, T;  // K must equal win*E
  int* err;
  int pool;             // dense only: row r = seq*T + t holds max_{k<win} dense[(seq*L + t + k)*lda + :]
                        // (max_pool1d(win, stride 1) over the time axis fused into the A load)
  int pad;              // gathered / windowed: left zero padding in positions ("same" convolutions)
  int dwin;             // dense windows (E = channels): 1 = 1-D window of `win` positions (im2col of a Conv1d over
                        // [seq, L, E]), 2 = 3x3 window over a [seq, L/Wm, Wm, E] map (im2col of a Conv2d, pad 1)
  int Wm;               // map width for dwin == 2
};
inline GemmA gemm_dense(const float* a, int64_t lda) { return GemmA{a, lda, nullptr, nullptr, 0, 0, 0, 0, 0, nullptr, 0, 0, 0, 0}; }
inline GemmA gemm_gather(const float* table, int V, int E, const int64_t* ids, int win, int L, int T, int* err,
                         int pad = 0) {
  return GemmA{nullptr, 0, table, ids, V, E, win, L, T, err, 0, pad, 0, 0};
}
inline GemmA gemm_pooled(const float* a, int64_t lda, int win, int L, int T) {
  return GemmA{a, lda, nullptr, nullptr, 0, 0, win, L, T, nullptr, 1, 0, 0, 0};
}
// same-padded Conv1d over a dense [seq, L, C] tensor: row (seq, t) = concat_{k<win} x[seq, t + k - win/2, :]
inline GemmA gemm_window1d(const float* a, int C, int win, int L) {
  return GemmA{a, C, nullptr, nullptr, 0, C, win, L, L, nullptr, 0, win / 2, 1, 0};
}
// 3x3 same-padded Conv2d over a dense NHWC map [seq, H, W, C]: row (seq, y, x) = concat_{ky,kx} x[seq, y+ky-1, x+kx-1, :]
inline GemmA gemm_window2d(const float* a, int C, int H, int W) {
  return GemmA{a, C, nullptr, nullptr, 0, C, 9, H * W, H * W, nullptr, 0, 1, 2, W};
}

// One 4-float slice A[r][kk..kk+3] (kk % 4 == 0, E % 4 == 0) for every provider; zeros outside the padding.
__device__ __forceinline__ float4 gemm_a_load4(const GemmA& a, int64_t r, int kk) {
  if (a.table) {
    const int64_t seq = r / a.T;
    const int t = (int)(r - seq * a.T);
    const int seg = kk / a.E;
    const int pos = t + seg - a.pad;
    if (pos < 0 || pos >= a.L) return make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t id = checked_id(a.ids[seq * a.L + pos], a.V, a.err);
    return *reinterpret_cast<const float4*>(a.table + id * a.E + (kk - seg * a.E));
  }
  if (a.dwin == 1) {
    const int64_t seq = r / a.L;
    const int t = (int)(r - seq * a.L);
    const int seg = kk / a.E;
    const int pos = t + seg - a.pad;
    if (pos < 0 || pos >= a.L) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(a.dense + (seq * a.L + pos) * a.lda + (kk - seg * a.E));
  }
  if (a.dwin == 2) {
    const int64_t seq = r / a.L;
    const int p = (int)(r - seq * a.L);
    const int y = p / a.Wm, x = p - y * a.Wm;
    const int seg = kk / a.E;
    const int yy = y + seg / 3 - 1, xx = x + seg % 3 - 1;
    if (yy < 0 || yy >= a.L / a.Wm || xx < 0 || xx >= a.Wm) return make_float4(0.f, 0.f, 0.f, 0.f);
    return *reinterpret_cast<const float4*>(a.dense + (seq * a.L + yy * a.Wm + xx) * a.lda + (kk - seg * a.E));
  }
  if (!a.pool) return *reinterpret_cast<const float4*>(a.dense + r * a.lda + kk);
  const int64_t seq = r / a.T;
  const float* base = a.dense + (seq * a.L + (r - seq * a.T)) * a.lda + kk;
  float4 m = *reinterpret_cast<const float4*>(base);
  for (int k = 1; k < a.win; ++k) {
    const float4 v = *reinterpret_cast<const float4*>(base + k * a.lda);
    m.x = fmaxf(m.x, v.x), m.y = fmaxf(m.y, v.y), m.z = fmaxf(m.z, v.z), m.w = fmaxf(m.w, v.w);
  }
  return m;
}
int32_t gemm_f32(const GemmA& a, const float* w, const float* bias, float* c, int64_t ldc, int64_t M,
                 int N, int K, Act act, cudaStream_t s);

// tcgen05 GEMM (gemm_tc.cu): weights pre-packed as bf16 hi/lo operand images per (column tile, K chunk).
struct GemmTcW {
  int N = 0, K = 0, NT = 0, nct = 0, nkc = 0;
  uint8_t* img = nullptr;
};
extern int g_gemm_dbg;
extern int g_gemm_impl;  // 1 (default): tcgen05 GEMM where usable, 0: fp32 CUDA-core GEMM everywhere
int32_t gemm_tc_pack(Owned& own, const float* w, int N, int K, GemmTcW* out, cudaStream_t s);
int32_t gemm_tc_repack(const float* w, const GemmTcW& tw, cudaStream_t s);
bool gemm_tc_usable(const GemmA& a, int K);
// aimg_scratch (optional, gemm_tc_aimg_bytes(M, K) bytes, 128-byte aligned): when the weights span several column tiles the A
// rows are gathered / converted ONCE into this operand image and the GEMM streams it with bulk copies
size_t gemm_tc_aimg_bytes(int64_t M, int K);
int32_t gemm_tc(const GemmA& a, const GemmTcW& w, const float* bias, float* c, int64_t ldc, int64_t M, Act act,
                cudaStream_t s, uint8_t* aimg_scratch = nullptr);
// out[r] = dot_w . act(A[r] W^T + bias) + dot_b with the [M, N] product kept in TMEM (attention-MLP scores)
bool gemm_tc_rowdot_usable(const GemmA& a, const GemmTcW& w, int64_t M);
int32_t gemm_tc_rowdot(const GemmA& a, const GemmTcW& w, const float* bias, Act act, const float* dot_w, const float* dot_b,
                       float* out, int64_t M, cudaStream_t s, uint8_t* aimg_scratch = nullptr);
// tensor-core GEMM when a packed image exists and the A provider is 128-bit loadable, else the fp32 kernel
inline int32_t gemm_auto(const GemmA& a, const float* w, const GemmTcW& tw, const float* bias, float* c, int64_t ldc,
                         int64_t M, int N, int K, Act act, cudaStream_t s, uint8_t* aimg_scratch = nullptr) {
  if (tw.img && M >= 128 && gemm_tc_usable(a, K)) return gemm_tc(a, tw, bias, c, ldc, M, act, s, aimg_scratch);
  return gemm_f32(a, w, bias, c, ldc, M, N, K, act, s);
}

// LSTM weights repacked for the recurrent kernel.
struct LstmPack {
  int in, h, dirs;
  int gates = 4;    // 4: LSTM (i,f,g,o), 3: GRU (r,z,n)
  float* b_hn = nullptr;  // [dirs][h] GRU only: hidden bias of the candidate gate (stays inside r * (...))
  float* w_ih;    // [dirs*4h, in]  (fwd rows then rev rows): the pre-gate GEMM's W
  float* bias;    // [dirs*4h]      b_ih + b_hh
  float* w_hh_t;  // [dirs][h][4h]  k-major recurrent weights
  float* w_hh = nullptr;  // [dirs][4h][h] torch layout, kept only when W_hh does not fit in shared memory (stepwise path)
  GemmTcW w_ih_tc;  // tensor-core image of w_ih (pre-gate GEMM)
};
int32_t lstm_pack(Owned& own, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, int in, int h, LstmPack* out,
                  cudaStream_t s, int rnn_type = CAIR_RNN_LSTM);
size_t lstm_workspace_floats(const LstmPack& p, int64_t n, int L);
// pre-gates GEMM (optionally gathering rows from `table`) + recurrence; out [n,L,dirs*h], zeros at t>=len.
// c_seq (optional, LSTM): the cell state of every step, same layout as `out` (needed by the CARS decoder, which starts
// from the session encoders' (h, c) after each query - neuroir/multitask/cars.py:385-411)
int32_t lstm_run(const LstmPack& p, const GemmA& x, const int64_t* len, int n, int L, float* out, float* h_n,
                 float* c_n, float* ws_pre, int* err, cudaStream_t s, const char* rec_name = "lstm_recurrence",
                 float* c_seq = nullptr);

// tcgen05 LSTM (lstm_tc.cu): fused input + recurrent projection per step, weights resident in smem.
struct LstmTcPack {
  int in = 0, h = 0, dirs = 0;
  uint8_t* wimg = nullptr;  // [dirs][2 row tiles][hi|lo] bf16 operand images
};
extern long long* g_lstm_dbg;  // optional role-timing counters (debug)
bool lstm_tc_supported(int in, int h);
extern int g_lstm_spc_min;
int lstm_tc_seqs_per_cta(int n, int dirs, int min_spc = 8);
int lstm_tc_ctas(int n, int dirs, int min_spc = 8);   // grid size lstm_tc_run uses for n sequences
int32_t lstm_tc_pack(Owned& own, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, int in, int h, LstmTcPack* out,
                     cudaStream_t s);
// bias: [dirs][4h] = b_ih + b_hh (LstmPack::bias)
// ximg (optional, gathered x only): the table pre-split into x-operand rows by lstm_tc_pack_table - the gather warps
// then copy 16-byte units instead of converting fp32 rows on every step.
int32_t lstm_tc_run(const LstmTcPack& p, const float* bias, const GemmA& x, const int64_t* len, int n, int L,
                    float* out, float* h_n, float* c_n, int* err, cudaStream_t s, const char* rec_name,
                    const uint8_t* ximg = nullptr, int min_spc = 8,   // min_spc: fewest sequences per CTA to consider
                    float* gates_out = nullptr, float* cseq_out = nullptr);   // training: gate activations [n*L, dirs*4h], c [n*L, dirs*h]
int32_t lstm_tc_repack(const LstmTcPack& p, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, cudaStream_t s);
// table [V, in] fp32 -> image [V][hi|lo][48 x bf16] (constant 1 in K slot `in` = the bias column, zero padding)
int32_t lstm_tc_pack_table(Owned& own, const float* table, int V, int in, uint8_t** img, cudaStream_t s);

// all-gather of scores over NVLink peer memory (p2p.cu)
int32_t allgather_scores_p2p(const float* send, int64_t count, const uint64_t* peer_recv, const uint64_t* peer_flags, int rank,
                             int world, uint32_t seq, cudaStream_t s);

// tcgen05 recurrence, gate rows split over a thread-block cluster (rnn_tc.cu): LSTM and GRU, h <= 128 per direction.
struct RnnTcPack {
  int in = 0, h = 0, dirs = 0, cs = 0, gru = 0, fused = 0, planes = 0;
  size_t img_bytes = 0;      // one (dir, rank) weight image
  uint8_t* wimg = nullptr;   // [dirs][cs] images
  float *wp = nullptr, *bp = nullptr;   // PRE mode (in >= 48): permuted + pre-scaled W_ih [dirs 4h, in] and bias
  GemmTcW wp_tc;
};
struct RnnTcPlan {
  int spc = 0, npad = 0, groups = 0, ctas = 0, nb = 0;
};
enum { RNN_IMPL_FP32 = 0, RNN_IMPL_TC_R1 = 1, RNN_IMPL_AUTO = 2, RNN_IMPL_CLUSTER = 3 };
// recurrence engine: 2 (default) = per shape the faster tcgen05 kernel (single-CTA lstm_tc.cu for LSTM, in < 48,
// 32 < h <= 64; cluster-split rnn_tc.cu otherwise), 3 = rnn_tc.cu wherever it applies, 1 = lstm_tc.cu where it applies,
// 0 = fp32 CUDA cores
extern int g_rnn_impl;
inline bool rnn_prefers_r1(int rnn_type, int in, int h) { return rnn_type == CAIR_RNN_LSTM && in < 48 && h > 32 && h <= 64; }
extern long long* g_rnn_dbg;
extern int g_rnn_spc_min, g_rnn_spc_force;
bool rnn_tc_supported(int in, int h);
int32_t rnn_tc_pack(Owned& own, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, int in, int h, int rnn_type,
                    RnnTcPack* out, cudaStream_t s);
int32_t rnn_tc_pack_table(Owned& own, const float* table, int V, int in, uint8_t** img, cudaStream_t s);
RnnTcPlan rnn_tc_plan(const RnnTcPack& p, int n, int min_spc = 8);
size_t rnn_tc_workspace_floats(const RnnTcPack& p, int64_t n, int L);   // PRE mode pre-gates
// x: dense rows or gathered table rows; ws_pre: rnn_tc_workspace_floats floats (PRE mode); ximg: pre-split table (FUSED, gathered)
int32_t rnn_tc_run(const RnnTcPack& p, const GemmA& x, const int64_t* len, int n, int L, float* out, float* h_n, float* c_n,
                   float* ws_pre, int* err, cudaStream_t s, const char* rec_name, const uint8_t* ximg = nullptr, int min_spc = 8);

}  // namespace cair
