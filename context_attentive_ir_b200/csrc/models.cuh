// Per-model state held by a cair_handle and the internal forward entry points.
#pragma once
#include "common.cuh"

namespace cair {

enum { CAIR_MODEL_ESM = 1, CAIR_MODEL_MT = 2, CAIR_MODEL_DRMM = 3, CAIR_MODEL_DUET = 4, CAIR_MODEL_CARS = 5,
       CAIR_MODEL_DSSM = 6, CAIR_MODEL_CDSSM = 7, CAIR_MODEL_ARCI = 8, CAIR_MODEL_ARCII = 9 };

int32_t embed_gather(const float* table, int V, int E, const int64_t* ids, int64_t T, float* out, int* err,
                     cudaStream_t s);

// ---- ESM ----
struct EsmState {
  int V = 0, E = 0;
  float* table = nullptr;
};
int32_t esm_forward(const float* table, int V, int E, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                    int64_t pair_begin, int64_t pair_count, float* scores, int* err, cudaStream_t s);

// ---- DRMM ----
struct DrmmState {
  cair_drmm_weights w{};  // pointers into handle-owned copies
  int32_t* dbg_hist = nullptr;
};
int32_t drmm_forward(const cair_drmm_weights& w, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                     int64_t pair_begin, int64_t pair_count, float* scores, int32_t* hist_out, Arena& ws, int* err,
                     cudaStream_t s, bool dry);

extern long long* g_drmm_dbg;
extern int g_drmm_impl;   // 1 (default): tcgen05 cosines + exact recompute at the bin edges, 0: fp32 CUDA-core kernels
bool drmm_tc_usable(int E, int Lq, int Ld, const float* table, int64_t vocab);
size_t drmm_tc_workspace_bytes(int E, int64_t nq);   // per-query operand records
int32_t drmm_tc_forward(const cair_drmm_weights& w, const int64_t* q, const int64_t* d, int N, int Lq, int Ld,
                        int64_t pair_begin, int64_t pair_count, float* scores, int32_t* hist_out, uint8_t* qrec, int* err,
                        cudaStream_t s);

// ---- Match-Tensor ----
struct MtPack {
  int C, nf, FP, FPP, M;
  float* w7;    // [3][7][C][FPP]   merged conv weights of the C product channels, f fastest
  float* w7t;   // [3][7][FP][CP]   same, channel fastest and zero-padded to CP = ceil16(C) (tensor-core T builder)
  float* wem;   // [3][7][FPP]      alpha * merged weights of the exact-match channel
  float* bias;  // [FPP]
  float* w1;    // [M][FPP]         1x1 conv
  float* b1;    // [M]
  float* wo;    // [M] + bo at [M]
};
constexpr int MT_TC_MAXM = 32;  // match_filter_size bound of the tcgen05 interaction kernel
// Epilogue weights of the tcgen05 interaction kernel, passed BY VALUE as a __grid_constant__ kernel
// parameter so the hot loop reads them as constant-bank operands (host copy made once at create).
struct MtEpiConst {
  float wem[21][24];        // alpha * W7[f, C, a, bt], index a*7+bt
  float bias[24];           // merged conv bias
  float w1t[24][MT_TC_MAXM]; // 1x1 conv, transposed: [f][m] (m contiguous: one 128-bit constant load = 4 output channels)
  float b1[MT_TC_MAXM];
  // PAD x PAD exact matches (PAD == PAD counts, mtensor.py:156): prefix sums over the document taps, already summed over the
  // query taps of a PAD run: padtab[qz][k][f] = sum_{a in qz} sum_{bt < k} wem[a*7+bt][f] (qz = bit mask of the PAD query taps)
  float padtab[8][8][24];
};
int32_t mt_pack(Owned& own, const cair_mt_weights& w, MtPack* p, cudaStream_t s);
size_t mt_t_floats(const MtPack& p, int64_t nq, int Lq);
bool mt_tc_supported(const MtPack& p, int Lq, int Ld);
void mt_tc_workspace(const MtPack& p, int64_t nq, int64_t pc, int Lq, int Ld, size_t* timg_bytes, size_t* aimg_bytes);
int32_t mt_epi_const(const MtPack& p, MtEpiConst* out, cudaStream_t s);  // synchronises s
int32_t mt_tc_build_t(const MtPack& p, const float* cq, uint8_t* timg, int Lq, int64_t nq, cudaStream_t s);
int32_t mt_tc_doc_image(const MtPack& p, const float* cd, uint8_t* aimg, int Ld, int64_t pair_count, cudaStream_t s);
bool mt_tc_proj_supported(int C, int Hd);
int32_t mt_tc_pack_wd(Owned& own, const float* wd, int C, int Hd, uint8_t** img, cudaStream_t s);
int32_t mt_tc_proj_image(const MtPack& p, const float* enc_d, int Hd, const uint8_t* wd_img, const float* bd,
                         uint8_t* aimg, int Ld, int64_t pair_count, cudaStream_t s);
int32_t mt_tc_interact(const MtPack& p, const MtEpiConst& ec, const uint8_t* timg, const uint8_t* aimg,
                       const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pair_begin,
                       int64_t pair_count, int64_t q_begin, int64_t nq, float* scores, cudaStream_t s,
                       int max_ctas = 0, float* pooled = nullptr, int* argidx = nullptr);

extern long long* g_mt_dbg;  // optional role-timing counters of the tcgen05 interaction kernel (debug)
enum { MT_IMPL_FP32 = 0, MT_IMPL_TC = 1, MT_IMPL_TC_SPLIT = 2 };  // 2: tcgen05 interaction, fp32 doc projection + image kernel
struct MtState {
  int V = 0, E = 0, F = 0, Hq = 0, Hd = 0, C = 0;
  int impl = MT_IMPL_TC;  // interaction kernel: tcgen05 bf16x3 (default) or the fp32 CUDA-core kernel
  float* folded = nullptr;  // [V, F] = table W_p^T + b_p  (eval-mode fold of mtensor.py:77-90)
  LstmPack enc_q{}, enc_d{};
  LstmTcPack tc_q{}, tc_d{};  // round-1 tensor-core encoders (LSTM, h <= 64, in < 48; kept for A/B runs)
  RnnTcPack rt_q{}, rt_d{};   // cluster-split tensor-core encoders (LSTM + GRU, h <= 128)
  uint8_t* folded_img = nullptr;  // folded table pre-split into hi/lo bf16 x-operand rows (lstm_tc_pack_table)
  float *wq = nullptr, *bq = nullptr, *wd = nullptr, *bd = nullptr;  // channel projections
  uint8_t* wd_img = nullptr;  // hi/lo operand image of the doc projection (fused projection + A image kernel)
  MtPack pack{};
  MtEpiConst epi{};
  cudaStream_t side = nullptr;            // query-side work runs here, forked/joined with events
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  float *dbg_enc_q = nullptr, *dbg_enc_d = nullptr;
  // stacked encoder layers 1.. (rnn_encoder.py:45-53,92-113; use_last: only the top layer's bank is consumed): dense input of
  // width H, h = H / dirs per direction.  Side 0 = query encoder, 1 = document encoder.
  static constexpr int MAX_EXTRA = 3;
  int nextra[2] = {0, 0};
  LstmPack xl[2][MAX_EXTRA]{};
  RnnTcPack xrt[2][MAX_EXTRA]{};
  int rnn_type = 0, dirs = 1;
};
int32_t mt_add_encoder_layer(Owned& own, MtState* st, int side, const cair_lstm_dir* fwd, const cair_lstm_dir* rev, cudaStream_t s);
int32_t mt_create_state(Owned& own, const cair_mt_weights& w, MtState* st, cudaStream_t s);
// Phases of one forward, for the cross-batch software pipeline of cair_ranker_submit_host: MT_ALL = everything on `s`;
// MT_ENCODE = query side (forked stream) + document encoder + channel projection / operand image; MT_INTERACT = the
// interaction kernel over the sub-range [ib, ib+ic) of the pair slice, on at most max_ctas CTAs (0 = all SMs), after
// joining the query side when `join` is set.  The workspace layout is identical in all phases.
enum { MT_ALL = 0, MT_ENCODE = 1, MT_INTERACT = 2 };
struct MtPhase {
  int phase = MT_ALL;
  int64_t ib = 0, ic = -1;
  int max_ctas = 0;
  bool join = true;
  int doc_min_spc = 8;   // MT_ENCODE: fewest sequences per CTA of the document encoder (32 in the serving pipeline: fewer,
                         // slightly slower encoder CTAs leave more SMs to the previous batch's interaction)
};
bool mt_can_pipeline(const MtState& st, int Lq, int Ld);
bool mt_doc_uses_cluster_kernel(const MtState& st);
int32_t mt_forward(const MtState& st, const int64_t* q, const int64_t* qlen, const int64_t* d, const int64_t* dlen,
                   int B, int N, int Lq, int Ld, int64_t pb, int64_t pc, float* scores, Arena& ws, int* err,
                   cudaStream_t s, bool dry, MtPhase ph = MtPhase());

// ---- DSSM / CDSSM ----
struct DssmState {
  int V = 0, E = 0, H = 0, O = 0;
  float* table = nullptr;
  float *w[4] = {nullptr, nullptr, nullptr, nullptr}, *b[4] = {nullptr, nullptr, nullptr, nullptr};  // q0, q2, d0, d2
};
int32_t dssm_create_state(Owned& own, const cair_dssm_weights& w, DssmState* st, cudaStream_t s);
int32_t dssm_forward(const DssmState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb,
                     int64_t pc, float* scores, Arena& ws, int* err, cudaStream_t s, bool dry);
struct CdssmState {
  int V = 0, E = 0, H = 0, O = 0;
  float* table = nullptr;
  float *w5[2] = {nullptr, nullptr}, *b5[2] = {nullptr, nullptr};  // merged 5-token conv weights [H, 5E] (query, doc)
  float *ws[2] = {nullptr, nullptr}, *bs[2] = {nullptr, nullptr};  // sem Linear [O, H]
};
int32_t cdssm_create_state(Owned& own, const cair_cdssm_weights& w, CdssmState* st, cudaStream_t s);
int32_t cdssm_forward(const CdssmState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb,
                      int64_t pc, float* scores, Arena& ws, int* err, cudaStream_t s, bool dry);

// ---- ARC-I / ARC-II ----
struct ArciState {
  int V = 0, E = 0, nl = 0, Lq = 0, Ld = 0, Lqo = 0, Ldo = 0, hid = 0;
  int F[CAIR_ARC_MAX_LAYERS] = {0}, k[CAIR_ARC_MAX_LAYERS] = {0}, P[CAIR_ARC_MAX_LAYERS] = {0};
  float* table = nullptr;
  float *qw[CAIR_ARC_MAX_LAYERS] = {nullptr}, *qb[CAIR_ARC_MAX_LAYERS] = {nullptr};  // [F][k*C] tap-major
  float *dw[CAIR_ARC_MAX_LAYERS] = {nullptr}, *db[CAIR_ARC_MAX_LAYERS] = {nullptr};
  GemmTcW qtc[CAIR_ARC_MAX_LAYERS], dtc[CAIR_ARC_MAX_LAYERS], mdtc;
  float *mq = nullptr, *md = nullptr, *b0 = nullptr, *w1 = nullptr, *b1 = nullptr;  // mlp.0 split (query | doc), columns permuted
};
int32_t arci_create_state(Owned& own, const cair_arci_weights& w, ArciState* st, cudaStream_t s);
int32_t arci_forward(const ArciState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb, int64_t pc,
                     float* scores, Arena& ws, int* err, cudaStream_t s, bool dry);
struct ArciiState {
  int V = 0, E = 0, F1 = 0, k1 = 0, nl = 0, Lq = 0, Ld = 0, Hf = 0, Wf = 0, Cf = 0, hid = 0;
  int F2[CAIR_ARC_MAX_LAYERS] = {0};
  float* table = nullptr;
  float *cqw = nullptr, *cqb = nullptr, *cdw = nullptr, *cdb = nullptr;
  float *w2[CAIR_ARC_MAX_LAYERS] = {nullptr}, *b2[CAIR_ARC_MAX_LAYERS] = {nullptr};  // [F][(ky*3+kx)*C + c]
  GemmTcW cqtc, cdtc, tc2[CAIR_ARC_MAX_LAYERS], m0tc;
  float *m0 = nullptr, *b0 = nullptr, *w1 = nullptr, *b1 = nullptr;
};
int32_t arcii_create_state(Owned& own, const cair_arcii_weights& w, ArciiState* st, cudaStream_t s);
int32_t arcii_forward(const ArciiState& st, const int64_t* q, const int64_t* d, int N, int Lq, int Ld, int64_t pb, int64_t pc,
                      float* scores, Arena& ws, int* err, cudaStream_t s, bool dry);

// ---- DUET ----
struct DuetState {
  int V = 0, E = 0, nf = 0, pool = 0, Lq = 0, Ld = 0;
  float* table = nullptr;
  float *lconv_t = nullptr, *lconv_b = nullptr;  // [Ld][nf] transposed local conv1d weight, [nf]
  float *lfc1_w = nullptr, *lfc1_b = nullptr, *lfc2_w = nullptr, *lfc2_b = nullptr, *lfc3_w = nullptr, *lfc3_b = nullptr;
  float *cq_w = nullptr, *cq_b = nullptr, *cd1_w = nullptr, *cd1_b = nullptr;  // [nf][3*E] (tap-major K)
  float *cd2_w = nullptr, *cd2_b = nullptr;
  GemmTcW cq_tc, cd1_tc, cd2_tc;  // tensor-core images of the three convolutions
  float *fc1_w = nullptr, *fc1_b = nullptr, *fc2_w = nullptr, *fc2_b = nullptr, *fc3_w = nullptr, *fc3_b = nullptr,
        *fc4_w = nullptr, *fc4_b = nullptr;
  // the local model and the query side depend on the queries only: they run on a side stream, concurrently with the document convolutions
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
int32_t duet_create_state(Owned& own, const cair_duet_weights& w, DuetState* st, cudaStream_t s);
int32_t duet_forward(const DuetState& st, const int64_t* q, const int64_t* d, int B, int N, int Lq, int Ld,
                     int64_t pb, int64_t pc, float* scores, Arena& ws, int* err, cudaStream_t s, bool dry);

// ---- CARS ----
struct AttnPack {
  int H = 0;
  float *w0 = nullptr, *b0 = nullptr, *w3 = nullptr, *b3 = nullptr;
  GemmTcW w0_tc;
};
// suggestion decoder (cars.py:605-657 weights, :706-791 greedy decode)
struct CarsDecoder {
  bool ready = false;
  int Hdec = 0, Vt = 0;
  float *th_w = nullptr, *th_b = nullptr, *tc_w = nullptr, *tc_b = nullptr;   // transform_hid / transform_cell [Hdec, Hs]
  float *w_ih = nullptr, *w_hh = nullptr, *bias = nullptr;                     // decoder LSTM [4Hdec,E], [4Hdec,Hdec], b_ih+b_hh
  float *attn_in = nullptr, *attn_out = nullptr;                               // GlobalAttention 'general' [Hdec,Hdec], [Hdec,2Hdec]
  float *dec_attn = nullptr;                                                   // [Hdec,Hq]
  float *pred1 = nullptr, *pred2 = nullptr;                                    // [Hd,Hdec], [Vt,Hd]
  float *sess_proj2 = nullptr;                                                 // shared_session_projector + private_session_projector2 [Hd,Hs]
};
struct CarsState {
  int V = 0, E = 0, Hq = 0, Hd = 0, Hsq = 0, Hsd = 0;
  int rd[3] = {0, 0, 0}, pool = 2;
  float* table = nullptr;
  LstmPack enc_q{}, enc_d{}, sess_q{}, sess_d{};
  RnnTcPack rt_q{}, rt_d{};   // tcgen05 recurrence of the query / document BiLSTM (h <= 128 per direction)
  AttnPack q_attn, d_attn, click_attn, sq_inner, sd_inner;
  float *sqa_w = nullptr, *sqa_b = nullptr, *sda_w = nullptr, *sda_b = nullptr;  // session_{query,doc}_attn
  float *qp_w = nullptr, *qp_b = nullptr;  // q_projection
  float* sess_proj = nullptr;              // shared_session_projector + private_session_projector1 [Hd, Hsq+Hsd]
  float *rk_w[3] = {nullptr, nullptr, nullptr}, *rk_b[3] = {nullptr, nullptr, nullptr};
  GemmTcW rk_tc[2];                        // tensor-core images of the first two Maxout layers (rows = B*S*N candidates)
  float* shared_proj = nullptr;            // shared_session_projector alone (the decoder adds private_session_projector2 to it)
  CarsDecoder dec;
  // the query chain (encode, pool, query-session LSTM) depends on the queries only: forked onto a side stream, it runs under
  // the document chain
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};
struct CarsIO {
  const int64_t *q, *qlen, *d, *dlen;
  const float* labels;
  float *scores, *pooled_q, *pooled_d, *clicks, *sess_q_attn, *sess_d_attn;
  // decoder-side outputs (optional): query memory banks [B*S,Lq,Hq]; (h, c) of the two session encoders after every
  // query, concatenated [B,S,Hsq+Hsd] (query part first, cars.py:391-411)
  float *enc_q = nullptr, *sess_h = nullptr, *sess_c = nullptr;
};
int32_t cars_create_state(Owned& own, const cair_cars_weights& w, CarsState* st, cudaStream_t s);
int32_t cars_set_decoder(Owned& own, CarsState* st, const cair_cars_decoder_weights& w, cudaStream_t s);
size_t cars_decode_workspace_bytes(const CarsState& st, int B, int S, int Lq);
// greedy decode of the next-query suggestion for the rows (b, s < S-1); predictions [B, S-1, max_len] int64
int32_t cars_decode(const CarsState& st, const float* enc_q, const int64_t* qlen, const float* sess_h, const float* sess_c,
                    const float* sess_q_attn, const float* sess_d_attn, int B, int S, int Lq, int max_len, const int64_t* tgt2src,
                    int64_t bos, int64_t* predictions, void* ws, size_t ws_bytes, int* err, cudaStream_t s);
int32_t cars_forward(const CarsState& st, const CarsIO& io, int B, int S, int N, int Lq, int Ld, int sb, int sc,
                     Arena& ws, int* err, cudaStream_t s, bool dry);

// helper shared by the model files
template <typename T>
inline int32_t dev_copy(Owned& own, const T* src, size_t n, T** dst, cudaStream_t s) {
  if (!src) return fail(CAIR_ERR_BAD_ARG, "null weight pointer");
  CAIR_CUDA(own.alloc(dst, n));
  CAIR_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, s));
  return CAIR_OK;
}

}  // namespace cair
