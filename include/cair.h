/*
 * cair.h - C ABI of libcair.so: the B200-native (sm_100a) scoring hot path of
 * wasiahmad/context_attentive_ir.
 *
 * The reference has no FFI boundary: its "plugin API" is the Python
 * call  network(queries, que_len, documents, doc_len) -> FloatTensor[B,N]
 * (neuroir/models/ranker.py:213,257; neuroir/models/multitask.py:264-269).  Every
 * cair_<model>_forward below is what a binding for that call would bind; the reference
 * forward it replaces is cited next to it.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types, no exceptions.
 *  - every function returns an int32 status (CAIR_OK == 0, < 0 error class); the text of
 *    the last error of the calling thread is cair_last_error().
 *  - ids/lengths are int64, PAD=0 (neuroir/inputters/constants.py:1-4,
 *    neuroir/inputters/ranker/vector.py:53-69); scores are fp32; row-major, contiguous.
 *  - *_forward: all pointers are DEVICE pointers owned by the caller; the call enqueues
 *    kernels on `stream` (a cudaStream_t passed as void*), never synchronises and never
 *    allocates: scratch comes from the caller-provided workspace
 *    (size from cair_<model>_workspace_bytes).
 *  - *_forward_host: all pointers are HOST pointers (pinned memory recommended); the call
 *    copies ids/lengths host->device, runs the same kernels, copies the scores back and
 *    returns after a host synchronisation.  The work runs on a stream owned by the handle,
 *    ordered after everything already queued on `stream`; from the second call with the same
 *    buffers and shapes on, the whole sequence is replayed as ONE captured CUDA graph.
 *    Staging buffers live in the handle.
 *  - weights structs hold DEVICE pointers in the torch state_dict layouts (SURVEY.md
 *    App. D); cair_<model>_create repacks/converts them once into the handle; the caller
 *    may free or modify its tensors afterwards (call create again to pick up new values).
 *  - a handle is bound to one device; calls on one handle are serialised by the caller;
 *    distinct handles may be used concurrently from different threads.
 *
 * The same weight structs (with HOST pointers) are consumed by the CPU oracle in
 * oracle/cair_oracle.c, which is test infrastructure and not part of this library.
 */
#ifndef CAIR_H_
#define CAIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAIR_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define CAIR_API __attribute__((visibility("default")))
#else
#define CAIR_API
#endif

enum {
  CAIR_OK = 0,
  CAIR_ERR_BAD_ARG = -1,     /* null pointer, negative size                              */
  CAIR_ERR_BAD_SHAPE = -2,   /* shape the reference would assert on (e.g. mtensor.py:71) */
  CAIR_ERR_UNSUPPORTED = -3, /* configuration outside what the kernels implement         */
  CAIR_ERR_CUDA = -4,        /* CUDA runtime error, text in cair_last_error()            */
  CAIR_ERR_WORKSPACE = -5    /* workspace too small / misaligned                         */
};

enum { CAIR_RNN_LSTM = 0, CAIR_RNN_GRU = 1 };

typedef struct cair_handle cair_handle; /* opaque */

CAIR_API int32_t cair_version(void);
CAIR_API const char* cair_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
CAIR_API int64_t cair_launch_count(void);
CAIR_API int32_t cair_destroy(cair_handle* h);
/* Kernels never trap on bad input: token ids outside [0, vocab) (nn.Embedding raises IndexError,
 * neuroir/modules/embeddings.py:165) and lengths outside [1, L] (pack_padded_sequence raises,
 * neuroir/encoders/rnn_encoder.py:73) set a device flag instead.  This call synchronises `stream`,
 * returns CAIR_ERR_BAD_ARG if the flag was set since the last poll, and clears it.
 * The *_forward_host entry points poll before returning. */
CAIR_API int32_t cair_poll_error(cair_handle* h, void* stream);

/* Stage timing for bench.py's roofline: when enabled, *_forward records CUDA events on the launching
 * stream between its stages (a few events per call, no synchronisation).  cair_profile_read waits
 * for the last recorded call and returns the stage durations in ms and their comma-joined names. */
CAIR_API int32_t cair_profile_enable(cair_handle* h, int32_t on);
CAIR_API int32_t cair_profile_read(cair_handle* h, char* names, size_t names_bytes, float* ms,
                                   int32_t capacity, int32_t* count);

/* ---- shared weight fragments ------------------------------------------------------------- */

/* nn.Linear: w [out,in], b [out] or NULL (bias=False). */
typedef struct {
  const float* w;
  const float* b;
} cair_linear;

/* One direction of one nn.LSTM layer, torch layout, gate row order i,f,g,o
 * (neuroir/encoders/rnn_encoder.py:45-53 -> torch.nn.LSTM): w_ih [4h,in], w_hh [4h,h], b_* [4h]. */
typedef struct {
  const float* w_ih;
  const float* w_hh;
  const float* b_ih;
  const float* b_hh;
} cair_lstm_dir;

/* nn.Sequential(Linear(H,H), Tanh, Dropout, Linear(H,1)) - state_dict keys "<name>.0.*", "<name>.3.*"
 * (neuroir/multitask/cars.py:41-46). */
typedef struct {
  cair_linear l0;
  cair_linear l3;
} cair_attn_mlp;

/* ---- kernel-level entry points (unit tests, reuse) --------------------------------------- */

/* out[t,:] = table[ids[t],:]   (neuroir/modules/embeddings.py:243-252 + util_class.py:42-53).
 * ids [T] int64, table [V,E] fp32, out [T,E] fp32. */
CAIR_API int32_t cair_embed_gather(const float* table, int32_t V, int32_t E, const int64_t* ids, int64_t T,
                          float* out, void* stream);

/* RNNEncoder.forward with lengths (neuroir/encoders/rnn_encoder.py:62-141), one layer LSTM:
 * x [n,L,in], len [n] int64 -> out [n,L,dirs*h], zeros at t >= len; reverse direction starts at
 * each sequence's own last token.  rev may be NULL (unidirectional).
 * h_n/c_n [dirs,n,h] optional (NULL to skip). */
CAIR_API int32_t cair_lstm_forward(const float* x, const int64_t* len, int32_t n, int32_t L, int32_t in,
                          int32_t h, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out,
                          float* h_n, float* c_n, void* stream);

/* Same for either cell type (rnn_type = CAIR_RNN_LSTM | CAIR_RNN_GRU; GRU weights [3h,*], gate rows r,z,n, and
 * n = tanh(W_in x + b_in + r * (W_hn h + b_hn)); c_n is not written for GRU).
 * Replaces getattr(nn, rnn_type) at neuroir/encoders/rnn_encoder.py:45-53 (called :100-102). */
CAIR_API int32_t cair_rnn_forward(int32_t rnn_type, const float* x, const int64_t* len, int32_t n, int32_t L,
                          int32_t in, int32_t h, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, float* out,
                          float* h_n, float* c_n, void* stream);

/* Process-wide recurrence engine (A/B runs and on-device cross-checks): 2 = auto (default): per shape the faster
 * tcgen05 kernel - the single-CTA kernel for LSTM with in < 48 and 32 < h <= 64 (a pure latency chain at cfg2),
 * the cluster-split kernel (LSTM and GRU, h <= 128 per direction, any input size) otherwise; 3 = cluster-split kernel
 * wherever it applies; 1 = single-CTA kernel where it applies; 0 = fp32 CUDA-core kernels.  Shapes an engine does not
 * cover fall through to the next one. */
CAIR_API int32_t cair_set_rnn_impl(int32_t impl);

/* On-device self test of the tcgen05 operand/descriptor conventions (csrc/umma.cuh):
 * D[m][n] = sum_k A[m+shift][k] * B[n][k] for m < 128; A [(128+shift),K], B [N,K], D [128,N] fp32 device
 * pointers; split=0: plain bf16 operands, split=1: bf16x3 split precision (~fp32 accuracy). */
CAIR_API int32_t cair_umma_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K,
                                    int32_t shift, int32_t split, void* stream);

/* Microbenchmark of the tcgen05 issue path: `reps` x K/16 back-to-back MMAs (M=128, N, bf16) by one CTA;
 * cycles[0] = SM cycles to issue them, cycles[1] = cycles until they have all retired (device int64[2]).
 * uniform=1: whole-warp uniform issue loop (as the product kernels), 0: single-thread issue. */
CAIR_API int32_t cair_umma_bench(int32_t N, int32_t K, int32_t reps, int32_t uniform, long long* cycles,
                                 void* stream);

/* ---- ESM (neuroir/rankers/esm.py:19-45) --------------------------------------------------- */
typedef struct {
  int32_t vocab, emsize;
  const float* table; /* word_embeddings.make_embedding.emb_luts.0.weight [V,E] */
} cair_esm_weights;

CAIR_API int32_t cair_esm_create(const cair_esm_weights* w, int32_t device, cair_handle** out);

/* ---- Match-Tensor (neuroir/rankers/mtensor.py:27-60 ctor, :62-131 forward, :134-158 exact match) */
typedef struct {
  int32_t vocab, emsize, featsize, nhid_query, nhid_doc, nchannels, nfilters, match_filter_size;
  int32_t rnn_type, bidirectional;
  const float* table;              /* [V,E] */
  cair_linear linear_projection;   /* [F,E] */
  cair_lstm_dir query_fwd, query_rev, doc_fwd, doc_rev; /* {query,document}_encoder.rnns.0.* */
  cair_linear query_projection;    /* [C,Hq] */
  cair_linear document_projection; /* [C,Hd] */
  const float* alpha;              /* exact_match_channel.alpha [1] */
  cair_linear conv1, conv2, conv3; /* w [nf,C+1,3,{3,5,7}], b [nf] */
  cair_linear conv;                /* w [mfs,3nf,1,1], b [mfs] */
  cair_linear output;              /* w [1,mfs], b [1] */
} cair_mt_weights;

CAIR_API int32_t cair_mt_create(const cair_mt_weights* w, int32_t device, cair_handle** out);
/* Stacked encoders (neuroir/encoders/rnn_encoder.py:45-53 builds `nlayers` single-layer RNNs, :92-113 feeds each the bank of
 * the one below; Match-Tensor keeps use_last = True, so only the top bank is consumed): appends layer `rnns.<k>` (k = 1, 2, 3 in
 * call order) to the query (side 0) or document (side 1) encoder.  Its input width is the side's hidden size; `rev` is NULL for a
 * unidirectional encoder.  Call after cair_mt_create and before the first forward; the weights are packed (copied) here. */
CAIR_API int32_t cair_mt_add_encoder_layer(cair_handle* h, int32_t side, const cair_lstm_dir* fwd, const cair_lstm_dir* rev);
/* Stage outputs for parity tests (any may be NULL): encoder memory banks
 * enc_q [B,Lq,Hq], enc_d [B*N,Ld,Hd] as RNNEncoder returns them (mtensor.py:93-94). */
CAIR_API int32_t cair_mt_set_debug(cair_handle* h, float* enc_q, float* enc_d);

/* Process-wide GEMM engine for the generic projections / convolutions (DUET, CARS, DSSM, channel projections):
 * 1 = tcgen05 bf16x3 tensor-core GEMM where the operand layout allows (default: persistent CTAs, one loader pipeline across
 * tiles, double-buffered TMEM accumulator drained by dedicated warps), 2 = the same arithmetic with one tile per CTA (the
 * round-1 kernel, kept for A/B runs), 0 = fp32 CUDA-core GEMM. */
CAIR_API int32_t cair_set_gemm_impl(int32_t impl);

/* Interaction kernel selection: 1 = tcgen05 bf16x3 split-precision tensor-core kernel (default when the
 * configuration fits: nfilters in {4,6}, nchannels <= 64, match_filter_size <= 32), 0 = fp32 CUDA-core
 * kernel (always available; the on-device cross-check of the tensor-core path), 2 = as 1 but with the
 * document channel projection on the fp32 GEMM followed by a separate image kernel (A/B check of the fused
 * tcgen05 projection that impl 1 uses). */
CAIR_API int32_t cair_mt_set_impl(cair_handle* h, int32_t impl);

/* ---- DRMM (neuroir/rankers/drmm.py:13-27 ctor, :29-84 forward, :87-98 gating) -------------- */
typedef struct {
  int32_t vocab, emsize, nbins; /* nbins must be 5 (neuroir/hyparam.py:78-81) */
  const float* table;
  cair_linear gating; /* gating_network.weight [1,E] */
  cair_linear ffnn0;  /* ffnn.0 [1,5] */
  cair_linear ffnn1;  /* ffnn.1 [1,1] */
  cair_linear output; /* output [1,1] */
} cair_drmm_weights;

/* Process-wide DRMM engine (A/B runs, on-device cross-check): 1 = tcgen05 bf16x3 cosines with an exact fp32 recompute of
 * every cell near a bin edge (default; histograms identical to the fp32 kernels'), 0 = fp32 CUDA-core kernels. */
CAIR_API int32_t cair_set_drmm_impl(int32_t impl);
CAIR_API int32_t cair_drmm_create(const cair_drmm_weights* w, int32_t device, cair_handle** out);
/* Optional parity output: hist [B*N,Lq,5] int32 (numpy.histogram counts, drmm.py:71-75). */
CAIR_API int32_t cair_drmm_set_debug(cair_handle* h, int32_t* hist);

/* ---- DUET (neuroir/rankers/duet.py:28-59, LocalModel :65-121, DistributedModel :127-208) --- */
typedef struct {
  int32_t vocab, emsize, nfilters, local_filter_size, dist_filter_size, pool_size;
  int32_t max_query_len, max_doc_len;
  const float* table;
  cair_linear local_conv1d; /* [nf,Ld,1] */
  cair_linear local_fc1;    /* [1,Lq]    */
  cair_linear local_fc2;    /* [nf,nf]   */
  cair_linear local_fc3;    /* [1,nf]    */
  cair_linear conv_q;       /* [nf,E,3]  */
  cair_linear conv_d1;      /* [nf,E,3]  */
  cair_linear conv_d2;      /* [nf,nf,1] */
  cair_linear dist_fc1;     /* [nf,nf]   */
  cair_linear dist_fc2;     /* [1,Ld-pool-1] */
  cair_linear dist_fc3;     /* [nf,nf]   */
  cair_linear dist_fc4;     /* [1,nf]    */
} cair_duet_weights;

CAIR_API int32_t cair_duet_create(const cair_duet_weights* w, int32_t device, cair_handle** out);

/* ---- DSSM (neuroir/rankers/dssm.py:10-31 ctor, :33-63 forward) ------------------------------------------- */
typedef struct {
  int32_t vocab, emsize, nhid, nout;
  const float* table;
  cair_linear query_mlp0, query_mlp2; /* query_mlp.0 [nhid,E], query_mlp.2 [nout,nhid] */
  cair_linear doc_mlp0, doc_mlp2;     /* doc_mlp.0, doc_mlp.2 */
} cair_dssm_weights;
CAIR_API int32_t cair_dssm_create(const cair_dssm_weights* w, int32_t device, cair_handle** out);

/* ---- CDSSM (neuroir/rankers/cdssm.py:10-30 ctor, :42-77 forward) ------------------------------------------ */
typedef struct {
  int32_t vocab, emsize, nhid, nout;
  const float* table;
  cair_linear query_conv, query_sem; /* Conv1d [nhid, 3E, 3], Linear [nout, nhid] */
  cair_linear doc_conv, doc_sem;
} cair_cdssm_weights;
CAIR_API int32_t cair_cdssm_create(const cair_cdssm_weights* w, int32_t device, cair_handle** out);

/* ---- ARC-I (neuroir/rankers/arci.py:10-58 ctor, :60-105 forward) ------------------------------------------- */
#define CAIR_ARC_MAX_LAYERS 4
typedef struct {
  int32_t vocab, emsize, nlayers;                 /* nlayers = len(filters_1d) <= CAIR_ARC_MAX_LAYERS */
  int32_t filters[CAIR_ARC_MAX_LAYERS], kernel[CAIR_ARC_MAX_LAYERS], pool[CAIR_ARC_MAX_LAYERS];
  int32_t max_query_len, max_doc_len;
  const float* table;
  cair_linear qconv[CAIR_ARC_MAX_LAYERS];         /* query_conv1d_layers.<i>.0 : [F_i, C_i, k_i] */
  cair_linear dconv[CAIR_ARC_MAX_LAYERS];         /* doc_conv1d_layers.<i>.0 */
  cair_linear mlp0, mlp1;                         /* mlp.0 [inp/2, inp], mlp.1 [1, inp/2] */
} cair_arci_weights;
CAIR_API int32_t cair_arci_create(const cair_arci_weights* w, int32_t device, cair_handle** out);

/* ---- ARC-II (neuroir/rankers/arcii.py:10-56 ctor, :58-111 forward); 2-D kernels 3x3, max-pools 2x2 ---------- */
typedef struct {
  int32_t vocab, emsize, filters_1d, kernel_1d, nlayers2d;
  int32_t filters_2d[CAIR_ARC_MAX_LAYERS];
  int32_t max_query_len, max_doc_len;
  const float* table;
  cair_linear conv_query, conv_doc;               /* [F1, E, k1] */
  cair_linear conv2d[CAIR_ARC_MAX_LAYERS];        /* conv2d_layers.<i>.0 : [F2_i, C_i, 3, 3] */
  cair_linear mlp0, mlp1;
} cair_arcii_weights;
CAIR_API int32_t cair_arcii_create(const cair_arcii_weights* w, int32_t device, cair_handle** out);

/* ---- forward for the stand-alone rankers ---------------------------------------------
 * network(queries, que_len, documents, doc_len) (neuroir/models/ranker.py:213,257):
 * q [B,Lq], qlen [B], d [B,N,Ld], dlen [B,N] int64 -> scores [B,N] fp32 (no softmax).
 * pair_begin/pair_count select the contiguous slice of the flattened pairs p=b*N+n this rank
 * scores (doc-parallel sharding); scores is still indexed [B,N] and only
 * the slice is written.  Use 0, B*N for everything. */
CAIR_API int32_t cair_ranker_workspace_bytes(cair_handle* h, int32_t B, int32_t N, int32_t Lq, int32_t Ld,
                                    size_t* bytes);
CAIR_API int32_t cair_ranker_forward(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                            const int64_t* dlen, int32_t B, int32_t N, int32_t Lq, int32_t Ld,
                            int64_t pair_begin, int64_t pair_count, float* scores, void* workspace,
                            size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_ranker_forward_host(cair_handle* h, const int64_t* q, const int64_t* qlen,
                                 const int64_t* d, const int64_t* dlen, int32_t B, int32_t N,
                                 int32_t Lq, int32_t Ld, float* scores, void* stream);
/* Pipelined form of cair_ranker_forward_host for a serving loop (what the reference does per batch in
 * main/ranker.py:252-264: Ranker.predict = .cuda(non_blocking) of the ids, forward, scores back to the host).
 * submit enqueues the H2D copies of the ids (copy stream), the scoring kernels and the D2H copy of the scores,
 * ordered after the work already queued on `stream`, and returns without waiting.  `slot` (0, 1 or 2) selects
 * one of three device staging areas + workspaces, so up to three batches are in flight: the copies of batch k+1
 * overlap the kernels of batch k, and for Match-Tensor (tcgen05 path) the batches are software-pipelined on the
 * device as well - the interaction kernel of batch k runs partly on the SMs that the 200-step document-encoder
 * recurrence of batch k+1 leaves idle (internal high / low priority streams; cair_ranker_set_pipeline_split
 * sets the share of a batch's pairs scored there, default 0.5).  wait blocks until that slot's scores are in
 * `scores` and reports bad token ids / lengths like the synchronous call (with several batches in flight a bad
 * id is reported by the first wait after it was detected).  The host buffers of a slot must stay valid and
 * unmodified until its wait returns; a slot is re-submitted only after its wait. */
CAIR_API int32_t cair_ranker_submit_host(cair_handle* h, const int64_t* q, const int64_t* qlen,
                                const int64_t* d, const int64_t* dlen, int32_t B, int32_t N,
                                int32_t Lq, int32_t Ld, float* scores, int32_t slot, void* stream);
CAIR_API int32_t cair_ranker_wait_host(cair_handle* h, int32_t slot);
CAIR_API int32_t cair_ranker_set_pipeline_split(cair_handle* h, float frac);

/* ---- the one collective of the doc-parallel path (SURVEY.md sections 8b, 8e) ------------------------
 * All-gather of fp32 scores over NVLink peer memory, replacing the gather of nn.DataParallel
 * (neuroir/models/ranker.py:341-346; the softmax / loss / MAP of :87,:258 need all N candidates of a query).
 * send [count] is this rank's slice; peer_recv[r] / peer_flags[r] (HOST arrays of `world` device pointers) are rank r's
 * receive buffer [world*count floats] and flag array [world uint32, zero before the first call] as mapped into THIS
 * process (torch symmetric memory or CUDA IPC); seq must increase by one per call on all ranks.  On return of the
 * enqueued kernel slot r of this rank's own receive buffer holds rank r's scores for every r.  Consecutive calls must
 * alternate between two receive buffers (the peers write the next call's slices while this rank may still read the
 * previous result).  One launch on `stream`, nothing is synchronised; a peer that never arrives traps the launch. */
CAIR_API int32_t cair_allgather_scores(const float* send, int64_t count, const uint64_t* peer_recv,
                              const uint64_t* peer_flags, int32_t rank, int32_t world, uint32_t seq, void* stream);
/* The same gather INSIDE the host entry points (cair_ranker_submit_host / wait_host / the plain submit form): after this
 * call every batch's scores are all-gathered between the kernels and the D2H copy, and the host buffer passed to
 * cair_ranker_submit_host receives ALL world * B * N scores (rank-major), so the doc-parallel serving loop keeps its
 * cross-batch software pipeline at N > 1.  peer_recv0 / peer_recv1: the two alternating receive buffers (each rank's
 * [world * count] floats, as mapped into this process), peer_flags as above (zeroed, used by this handle only);
 * count = B * N of every batch.  All ranks must submit the same number of batches.  world <= 1 or NULL switches it off. */
CAIR_API int32_t cair_ranker_set_gather(cair_handle* h, const uint64_t* peer_recv0, const uint64_t* peer_recv1,
                               const uint64_t* peer_flags, int32_t rank, int32_t world, int64_t count);

/* ---- ranking metrics of the evaluation loops, on the device ----------
 * Replaces, per batch, `scores.cpu()` + `np.argsort(-scores)` + MAP / MRR / precision_at_k(1,3,5)
 * (main/ranker.py:257-264, main/multitask.py:286-293; neuroir/eval/ltorank.py:4-26, 29-47, 104-123)
 * and, when apply_softmax != 0, the `f.softmax(scores, dim=-1)` of Ranker.predict (models/ranker.py:257-258).
 * scores [B,N] fp32 and labels [B,N] int64 are DEVICE pointers.  per_row [B,5] (device, float64) receives
 * average precision, reciprocal rank, precision@1, @3, @5 of every query; batch_mean [5] (device, float64,
 * may be NULL) their means = the values the reference feeds to its AverageMeters.  Ties are ranked in index
 * order (numpy's default argsort is unstable: the reference's order inside a tie is implementation-defined).  A row without a relevant document gives NaN (the reference
 * divides by zero there).  N < 5 is CAIR_ERR_BAD_SHAPE (ltorank.py:41 asserts).  Enqueued on `stream`. */
CAIR_API int32_t cair_rank_metrics(const float* scores, const int64_t* labels, int32_t B, int32_t N,
                          int32_t apply_softmax, double* per_row, double* batch_mean, void* stream);

/* ---- batchify on the device ----------------------------------------
 * Replaces the host-side padding loops of batchify (neuroir/inputters/ranker/vector.py:39-90): the batch arrives
 * RAGGED - q_tokens / d_tokens are the concatenated int32 token ids of the B queries / B*N documents, q_offsets
 * [B+1] / d_offsets [B*N+1] their int64 start offsets (all DEVICE pointers; 4 bytes per real token cross the bus
 * instead of 8 bytes per padded position) - and leaves as the padded int64 tensors the scoring entry points
 * take: q [B,Lq], qlen [B], d [B,N,Ld], dlen [B,N], PAD = 0 beyond each length.  Lq / Ld are the batch maxima
 * (or max_query_len / max_doc_len under force_pad, vector.py:18-21), chosen by the caller.  A sequence that is
 * empty or longer than its padded length ORs 2 (bad length) into *err_flag (device int32; the reference's
 * copy_ raises there).  The ranking-side tensors of the multitask batch (inputters/multitask/vector.py:82-149:
 * source_words [B,S,Lq], document_words [B,S,N,Ld] and their lengths) are the same call with B*S queries.
 * Enqueued on `stream`, nothing is synchronised. */
CAIR_API int32_t cair_batchify_ranker(const int32_t* q_tokens, const int64_t* q_offsets, const int32_t* d_tokens,
                             const int64_t* d_offsets, int32_t B, int32_t N, int32_t Lq, int32_t Ld,
                             int64_t* q, int64_t* qlen, int64_t* d, int64_t* dlen, int32_t* err_flag,
                             void* stream);

/* ---- CARS ranking path (neuroir/multitask/cars.py:193-304 encode*, :306-458 encode_session,
 *      :460-540 rank/rank_document, :671-691 apply_pooling; neuroir/modules/maxout.py:70-84) --- */
typedef struct {
  int32_t vocab, emsize, nhid_query, nhid_document, nhid_session_query, nhid_session_document;
  int32_t rank_dims[3];  /* Maxout output_dims, reference fixes [256,128,1] (cars.py:128-131) */
  int32_t rank_pool;     /* 2 */
  const float* table;    /* embedder.word_embeddings.make_embedding.emb_luts.0.weight */
  cair_lstm_dir query_fwd, query_rev, doc_fwd, doc_rev; /* {query,document}_encoder.encoder.rnns.0.* */
  cair_attn_mlp q_attn, d_attn, click_attn;
  cair_attn_mlp session_query_inner_attn, session_doc_inner_attn;
  cair_lstm_dir session_query, session_doc; /* session_{query,doc}_encoder.encoder.rnns.0.* */
  cair_linear session_query_attn, session_doc_attn;         /* [Hq,Hsq], [Hd,Hsd] */
  cair_linear shared_session_projector, private_session_projector1; /* [Hd,Hsq+Hsd], no bias */
  cair_linear q_projection;                                 /* [Hd,Hq] */
  cair_linear ranknet[3];                                   /* ranknet._linear_layers.{0,1,2} */
} cair_cars_weights;

CAIR_API int32_t cair_cars_create(const cair_cars_weights* w, int32_t device, cair_handle** out);
CAIR_API int32_t cair_cars_workspace_bytes(cair_handle* h, int32_t B, int32_t S, int32_t N, int32_t Lq,
                                  int32_t Ld, size_t* bytes);
/* encode + rank_document (multitask.py:264-269):
 * q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N] int64, labels [B,S,N] fp32
 * -> scores [B,S,N]; optional outputs (NULL to skip): pooled_q [B,S,Hq], pooled_d [B,S,N,Hd],
 * clicks [B,S,Hd], sess_q_attn [B,S,Hsq], sess_d_attn [B,S,Hsd].
 * session_begin/session_count select the sessions this rank computes (sessions are independent
 * except for the batch-global click-mask width, which is always taken over ALL B*S label rows -
 * SURVEY.md section 8e / App. B4); all tensors are indexed [B,...] and only the slice is written.
 * workspace_bytes is sized for all B sessions by cair_cars_workspace_bytes. */
CAIR_API int32_t cair_cars_forward(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                          const int64_t* dlen, const float* labels, int32_t B, int32_t S, int32_t N,
                          int32_t Lq, int32_t Ld, int32_t session_begin, int32_t session_count,
                          float* scores, float* pooled_q, float* pooled_d, float* clicks,
                          float* sess_q_attn, float* sess_d_attn, void* workspace,
                          size_t workspace_bytes, void* stream);


/* Optional outputs of the CARS forward beyond the scores (any pointer may be NULL): the stage outputs of
 * cair_cars_forward plus what the suggestion decoder starts from - enc_q [B*S,Lq,Hq] the query memory banks
 * (`encoded_source` of CARS.encode, cars.py:214-225), sess_h / sess_c [B,S,Hsq+Hsd] the (h, c) of the session query
 * and session document encoders after every query, query part first (cars.py:391-411). */
typedef struct {
  float *pooled_q, *pooled_d, *clicks, *sess_q_attn, *sess_d_attn;
  float *enc_q, *sess_h, *sess_c;
} cair_cars_outputs;
CAIR_API int32_t cair_cars_forward_ex(cair_handle* h, const int64_t* q, const int64_t* qlen, const int64_t* d,
                             const int64_t* dlen, const float* labels, int32_t B, int32_t S, int32_t N,
                             int32_t Lq, int32_t Ld, int32_t session_begin, int32_t session_count,
                             float* scores, const cair_cars_outputs* outs, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---- CARS query-suggestion decoder ------------------------------------
 * Weights of the decoder-side modules (neuroir/multitask/cars.py:605-657): transform_{hid,cell}.linear
 * [Hdec,Hsq+Hsd] + bias; decoder.decoder.rnn (nn.LSTM: weight_ih_l0 [4Hdec,E], weight_hh_l0 [4Hdec,Hdec], biases);
 * decoder.decoder.attn.linear_in [Hdec,Hdec] / linear_out [Hdec,2Hdec] (GlobalAttention 'general', no biases);
 * dec_attn [Hdec,Hq]; token_prob_predictor1 [Hd,Hdec]; token_prob_predictor2 [Vt,Hd];
 * private_session_projector2.linear [Hd,Hsq+Hsd] (all without bias).  Copied into the handle. */
typedef struct {
  int32_t nhid_decoder, tgt_vocab;
  cair_linear transform_hid, transform_cell;
  cair_lstm_dir rnn;
  cair_linear attn_in, attn_out, dec_attn, predictor1, predictor2, private_session_projector2;
} cair_cars_decoder_weights;
CAIR_API int32_t cair_cars_set_decoder(cair_handle* h, const cair_cars_decoder_weights* w);
CAIR_API int32_t cair_cars_decode_workspace_bytes(cair_handle* h, int32_t B, int32_t S, int32_t Lq, size_t* bytes);
/* CARS.decode (cars.py:706-791; called by Multitask.predict, models/multitask.py:281-292): greedy decode of max_len
 * tokens of the next query for every (session b, query s < S-1) row, starting from BOS.  Inputs are the outputs of
 * cair_cars_forward_ex (device pointers) and qlen [B,S]; tgt2src [Vt] int64 maps a target-vocabulary id to the
 * source-vocabulary id of the same word (the tgt_dict[idx] -> src_dict[word] round trip of cars.py:780-783).
 * predictions [B,S-1,max_len] int64 (target-vocabulary ids).  The reference's row orderings are reproduced as they
 * are (initial states query-index-major, memory banks / session summaries / predictions batch-major).
 * Enqueued on `stream`; nothing is synchronised. */
CAIR_API int32_t cair_cars_decode(cair_handle* h, const float* enc_q, const int64_t* qlen, const float* sess_h,
                         const float* sess_c, const float* sess_q_attn, const float* sess_d_attn, int32_t B,
                         int32_t S, int32_t Lq, int32_t max_len, const int64_t* tgt2src, int64_t bos_id,
                         int64_t* predictions, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training step of Match-Tensor ---------------------------------------
 * Replaces, inside Ranker.update (neuroir/models/ranker.py:192-230), the train-mode `self.network(...)` call (:213) and
 * the part of `loss.backward()` (:219) below the scores: criterion (:214, :53-65, :80-89), clip_grad_norm (:222) and
 * optimizer.step() (:226) stay torch calls of the unchanged wrapper on the [B,N] scores / the parameter list.
 * A trainer holds the LIVE device pointers of the parameters (cair_mt_weights; read again at every step, so in-place
 * optimizer updates are seen; LSTM encoders only) and small repacked copies it refreshes itself.
 * forward: train-mode MatchTensor.forward (rankers/mtensor.py:62-131) with emb_drop (:84, :89; mask = counter-based
 *   hash of (seed, element), scale 1/(1-p); p = 0 gives the eval arithmetic) -> scores [B,N]; the workspace keeps the
 *   saved activations (gate activations, cell states, memory banks, channel projections, arg-max cells of the max-pools).
 * backward: dscores [B,N] -> gradients ACCUMULATED (+=) into the buffers of `grads`, a cair_mt_weights whose pointers
 *   address zero-initialised (or running) gradient buffers of the parameters' shapes; grads->table may be NULL
 *   (--fix_embeddings, models/ranker.py:160-162).  Must follow the forward of the same batch with the same workspace,
 *   p_drop and seed.  Row 0 (PAD) of the table receives no gradient (nn.Embedding padding_idx).
 * Both enqueue on `stream` and synchronise nothing.  cair_mt_train_poll_error synchronises and reports token ids /
 * lengths out of range seen by the last forward.  cair_dropout_mask writes the keep-scale (0 or 1/(1-p)) of elements
 * [0, n) of that hash: element index = row * emsize + e with the B*Lq query rows first, then the B*N*Ld document rows. */
typedef struct cair_mt_trainer cair_mt_trainer;
CAIR_API int32_t cair_mt_train_create(const cair_mt_weights* params, int32_t device, cair_mt_trainer** out);
CAIR_API int32_t cair_mt_train_destroy(cair_mt_trainer* t);
/* engines: 1 (default) = the tcgen05 interaction kernel of the scoring path in its arg-max instantiation and the tcgen05
 * bf16x3 GEMM for the dense layers (linear_projection, pre-gates, channel projections, gate gradients -> feature rows) where
 * the shapes allow, 0 = fp32 CUDA-core kernels throughout (A/B runs; the recurrences, the sparse interaction backward and the
 * weight-gradient GEMMs are the same either way) */
CAIR_API int32_t cair_mt_train_set_impl(cair_mt_trainer* t, int32_t tc_forward);
CAIR_API int32_t cair_mt_train_workspace_bytes(cair_mt_trainer* t, int32_t B, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes);
CAIR_API int32_t cair_mt_train_forward(cair_mt_trainer* t, const int64_t* q, const int64_t* qlen, const int64_t* d,
                              const int64_t* dlen, int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed,
                              float* scores, void* workspace, size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_mt_train_backward(cair_mt_trainer* t, const int64_t* q, const int64_t* qlen, const int64_t* d,
                               const int64_t* dlen, int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed,
                               const float* dscores, const cair_mt_weights* grads, void* workspace, size_t workspace_bytes,
                               void* stream);
CAIR_API int32_t cair_mt_train_poll_error(cair_mt_trainer* t, void* workspace, void* stream);
CAIR_API int32_t cair_dropout_mask(uint64_t seed, float p, int64_t n, float* out, void* stream);

/* Training step of DRMM (neuroir/rankers/drmm.py:29-84 in train mode under Ranker.update, models/ranker.py:192-230).  The
 * reference computes the histograms with numpy, so no gradient flows through the cosines: the differentiable part is the
 * term gate (softmax over the dropped query embeddings), ffnn and output.  Stateless: `w` holds the live parameter pointers.
 * forward: train-mode scores [B,N]; emb_drop (same hash and element order as above: B*Lq query rows, then B*N*Ld document
 * rows) is applied to the rows the cosines see, exactly as drmm.py:47,57 do; the workspace keeps the dropped rows and the
 * histograms.  backward: dscores -> gradients accumulated into `grads` (a cair_drmm_weights of gradient buffers; table may
 * be NULL = fixed embeddings; only query tokens reach the table, through the gate; PAD row untouched). */
CAIR_API int32_t cair_drmm_train_workspace_bytes(int32_t emsize, int32_t B, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes);
CAIR_API int32_t cair_drmm_train_forward(const cair_drmm_weights* w, const int64_t* q, const int64_t* d, int32_t B, int32_t N,
                                int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, float* scores, void* workspace,
                                size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_drmm_train_backward(const cair_drmm_weights* w, const cair_drmm_weights* grads, const int64_t* q, int32_t B,
                                 int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, const float* dscores,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Training step of ESM (neuroir/rankers/esm.py:19-45 under Ranker.update, models/ranker.py:192-230).  The embedding table is
 * the model's only parameter.  forward: train-mode scores [B,N] (ESM has no dropout: the same arithmetic as the scoring path,
 * normalise first, then dot) and, in the workspace, the mean vectors and norms.  backward: dscores [B,N] -> dtable [V,E]
 * ACCUMULATED (atomic adds; zero it first): the cosine's derivative per pair, summed over the N documents of a query, divided
 * by the padded length and scattered to the rows of the non-PAD tokens (row 0 untouched, nn.Embedding padding_idx).
 * dtable NULL = fixed embeddings (nothing to do).  `scores` in backward are the forward's. */
CAIR_API int32_t cair_esm_train_workspace_bytes(int32_t emsize, int32_t B, int32_t N, size_t* bytes);
CAIR_API int32_t cair_esm_train_forward(const float* table, int32_t vocab, int32_t emsize, const int64_t* q, const int64_t* d, int32_t B,
                               int32_t N, int32_t Lq, int32_t Ld, float* scores, void* workspace, size_t workspace_bytes,
                               void* stream);
CAIR_API int32_t cair_esm_train_backward(int32_t vocab, int32_t emsize, const int64_t* q, const int64_t* d, int32_t B, int32_t N,
                                int32_t Lq, int32_t Ld, const float* scores, const float* dscores, float* dtable, void* workspace,
                                size_t workspace_bytes, void* stream);

/* Training step of DSSM (neuroir/rankers/dssm.py:33-63 under Ranker.update, models/ranker.py:192-230).  Stateless: `w` holds
 * the live parameter pointers.  forward: embedding rows x emb_drop mask (the hash and element order of cair_dropout_mask: B*Lq
 * query token rows, then B*N*Ld document token rows) -> max over the padded length (arg-max position kept) -> Linear, Tanh,
 * Linear, Tanh per side -> cosine; train-mode scores [B,N].  backward: dscores -> gradients ACCUMULATED into `grads` (a
 * cair_dssm_weights of zeroed gradient buffers; table may be NULL = fixed embeddings; the PAD row receives nothing).
 * `scores` in backward are the forward's. */
CAIR_API int32_t cair_dssm_train_workspace_bytes(int32_t emsize, int32_t nhid, int32_t nout, int32_t B, int32_t N, size_t* bytes);
CAIR_API int32_t cair_dssm_train_forward(const cair_dssm_weights* w, const int64_t* q, const int64_t* d, int32_t B, int32_t N,
                                int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, float* scores, void* workspace,
                                size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_dssm_train_backward(const cair_dssm_weights* w, const cair_dssm_weights* grads, const int64_t* q, const int64_t* d,
                                 int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, const float* scores,
                                 const float* dscores, void* workspace, size_t workspace_bytes, void* stream);

/* Training step of CDSSM (neuroir/rankers/cdssm.py:42-77 under Ranker.update).  As cair_dssm_train_*: stateless, live parameter
 * pointers in `w`, gradients ACCUMULATED into zeroed buffers laid out as a cair_cdssm_weights (conv weights in the reference's
 * Conv1d layout [nhid, 3E, 3]; table may be NULL), the emb_drop mask of cair_dropout_mask.  The interleave + Conv1d pair runs
 * as one linear map over 5-token windows, every layer a dense GEMM forward and backward; Lq, Ld >= 5. */
CAIR_API int32_t cair_cdssm_train_workspace_bytes(int32_t emsize, int32_t nhid, int32_t nout, int32_t B, int32_t N, int32_t Lq,
                                         int32_t Ld, size_t* bytes);
CAIR_API int32_t cair_cdssm_train_forward(const cair_cdssm_weights* w, const int64_t* q, const int64_t* d, int32_t B, int32_t N,
                                 int32_t Lq, int32_t Ld, float p_drop, uint64_t seed, float* scores, void* workspace,
                                 size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_cdssm_train_backward(const cair_cdssm_weights* w, const cair_cdssm_weights* grads, const int64_t* q,
                                  const int64_t* d, int32_t B, int32_t N, int32_t Lq, int32_t Ld, float p_drop, uint64_t seed,
                                  const float* scores, const float* dscores, void* workspace, size_t workspace_bytes, void* stream);

/* ---- MNSRF ranking path ---------------------------------------------------
 * Replaces MNSRF.encode + MNSRF.rank_document (neuroir/multitask/mnsrf.py:61-162) as Multitask.predict calls them
 * (neuroir/models/multitask.py:270-276).  Weights are copied into the handle: table = embedder.word_embeddings...weight
 * [V,E]; {query,document}_encoder.encoder.rnns.0.* ((Bi)LSTM or (Bi)GRU); session = session_query_encoder.encoder.rnns.0.*
 * (unidirectional, input nhid_query, hidden nhid_session); projection = projection.linear [Hd, Hq+Hs] + bias.
 * forward: q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N] int64 -> scores [B,S,N] (no softmax) for the sessions
 * [session_begin, session_begin+session_count); optional outputs (NULL to skip): memory_bank [B,S,Hq] (max-pooled query
 * encodings), session_bank [B,S,Hs], session_cell [B,S,Hs] (LSTM only) - what MNSRF.encode returns / the decoder starts from.
 * M_MATCH_TENSOR.rank_document (multitask/mmtensor.py:127-189) needs no entry point of its own: its scores do not depend
 * on the session, it is cair_ranker_forward on a Match-Tensor handle with B*S queries (see multitask.py M_MATCH_TENSOR). */
typedef struct {
  int32_t vocab, emsize, nhid_query, nhid_document, nhid_session, rnn_type, bidirectional;
  const float* table;
  cair_lstm_dir query_fwd, query_rev, doc_fwd, doc_rev, session;
  cair_linear projection;
} cair_mnsrf_weights;
typedef struct cair_mnsrf cair_mnsrf;
CAIR_API int32_t cair_mnsrf_create(const cair_mnsrf_weights* w, int32_t device, cair_mnsrf** out);
CAIR_API int32_t cair_mnsrf_destroy(cair_mnsrf* h);
CAIR_API int32_t cair_mnsrf_workspace_bytes(cair_mnsrf* h, int32_t B, int32_t S, int32_t N, int32_t Lq, int32_t Ld, size_t* bytes);
CAIR_API int32_t cair_mnsrf_forward(cair_mnsrf* h, const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen,
                           int32_t B, int32_t S, int32_t N, int32_t Lq, int32_t Ld, int32_t session_begin, int32_t session_count,
                           float* scores, float* memory_bank, float* session_bank, float* session_cell, void* workspace,
                           size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_mnsrf_poll_error(cair_mnsrf* h, void* stream);

/* Suggestion decoder of MNSRF / M_MATCH_TENSOR (multitask/mnsrf.py:258-300, mmtensor.py:258-300; called by
 * Multitask.predict, models/multitask.py:281-292): RNNDecoder without attention + generator, greedy from BOS.
 * table = the embedder's table [V,E] (LIVE pointer, not copied); session = session_query_encoder.encoder.rnns.0.* or all
 * NULL when the states come from cair_mnsrf_forward; dec_rnn = decoder.decoder.rnn.* ([4Hs,E], [4Hs,Hs]); generator [Vt,Hs]+b.
 * cair_sessdec_states: session LSTM over pooled [B,S,nhid_in] -> sess_h, sess_c [B,S,Hs] (M_MATCH_TENSOR.encode, :94-116).
 * cair_sessdec_decode: sess_h / sess_c [B,S,Hs] -> predictions [B,S-1,max_len] int64 (target ids); tgt2src [Vt] int64 as in
 * cair_cars_decode.  The reference's row orderings are reproduced (initial states query-index-major, predictions
 * batch-major).  cair_linear_maxpool: out[n,:] = max_t(x[n,t,:] W^T + b) (mmtensor.py:86-92), scratch n*L*C floats. */
typedef struct {
  int32_t vocab, emsize, nhid_in, nhid_session, tgt_vocab;
  const float* table;
  cair_lstm_dir session, dec_rnn;
  cair_linear generator;
} cair_sessdec_weights;
typedef struct cair_sessdec cair_sessdec;
CAIR_API int32_t cair_sessdec_create(const cair_sessdec_weights* w, int32_t device, cair_sessdec** out);
CAIR_API int32_t cair_sessdec_destroy(cair_sessdec* h);
CAIR_API int32_t cair_sessdec_workspace_bytes(cair_sessdec* h, int32_t B, int32_t S, size_t* bytes);
CAIR_API int32_t cair_sessdec_states(cair_sessdec* h, const float* pooled, int32_t B, int32_t S, float* sess_h, float* sess_c,
                            void* workspace, size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_sessdec_decode(cair_sessdec* h, const float* sess_h, const float* sess_c, int32_t B, int32_t S,
                            int32_t max_len, const int64_t* tgt2src, int64_t bos_id, int64_t* predictions, void* workspace,
                            size_t workspace_bytes, void* stream);
CAIR_API int32_t cair_linear_maxpool(const float* x, const float* w, const float* b, int32_t n, int32_t L, int32_t H, int32_t C,
                            float* out, float* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAIR_H_ */
