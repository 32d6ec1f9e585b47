"""TEST INFRASTRUCTURE - CPU restatement (numpy / plain loops) of the reference's ranking metrics as the evaluation
loops use them.  Only tests/ may import this module; it is the checker of cair_rank_metrics, never a product path.

Reference path restated (file:line relative to /root/reference):
  models/ranker.py:257-258      scores = f.softmax(network(...), dim=-1)
  main/ranker.py:257-264        predictions = np.argsort(-scores); MAP / MRR / precision_at_k(1, 3, 5) per batch
  main/multitask.py:286-293     the same for CARS (scores flattened to [B*S, N])
  eval/ltorank.py:4-26          MAP   (divides by the number of relevant docs, no zero guard - SURVEY B8)
  eval/ltorank.py:29-47         precision_at_k (counts NON-ZERO labels among the top k)
  eval/ltorank.py:104-123       MRR   (first position whose label == 1)
Ties: numpy's default argsort is not stable (numpy >= 1.25 sorts float keys with the AVX-512 / AVX2 x86-simd-sort
kernels, whose order inside a run of equal keys depends on the ISA), so the reference's order inside a tie is
implementation-defined; the restatement (and the CUDA kernel) rank tied candidates in index order.
Pinned against the reference's own functions by tests/golden/rank_metrics*.npz (oracle/gen_golden.py).
"""
import numpy as np


def softmax_f32(scores):
    """fp32 softmax over the last axis the way torch computes it: exp(x - max) / sum."""
    x = np.asarray(scores, dtype=np.float32)
    e = np.exp(x - x.max(axis=-1, keepdims=True), dtype=np.float32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)


def predictions(probs):
    """np.argsort(-scores) with the stable tie order (main/ranker.py:257)."""
    return np.argsort(-np.asarray(probs), axis=-1, kind='stable')


def per_row(pred, target):
    """[B,5] float64: average precision, reciprocal rank, precision@1/3/5 of every row (eval/ltorank.py loops)."""
    nrow, ncol = target.shape
    out = np.zeros((nrow, 5), dtype=np.float64)
    for i in range(nrow):
        ap, num_rel, rr = 0.0, 0, 0.0
        for j in range(ncol):
            if target[i, pred[i, j]] == 1:
                num_rel += 1
                ap += num_rel / (j + 1)
                if rr == 0.0:
                    rr = 1.0 / (j + 1)
        out[i, 0] = ap / num_rel if num_rel else np.nan   # the reference raises ZeroDivisionError here (B8)
        out[i, 1] = rr
        for c, k in enumerate((1, 3, 5)):
            out[i, 2 + c] = np.count_nonzero(target[i, pred[i, :k]]) / k
    return out


def rank_metrics(scores, labels, apply_softmax=True):
    """Batch means (map, mrr, prec@1, prec@3, prec@5) and the per-row values, from raw network scores."""
    p = softmax_f32(scores) if apply_softmax else np.asarray(scores, dtype=np.float32)
    pred = predictions(p)
    rows = per_row(pred, np.asarray(labels))
    return rows.mean(axis=0), rows, pred


# ---------------------------------------------------------------------------------------------------------------
# batchify, restated for flat token arrays.
#   inputters/ranker/vector.py:39-90   doc_word [B,N,max_doc_len] / que_word [B,max_que_len] zero (PAD) filled,
#                                      doc_len [B,N] / que_len [B] = token counts, max lens = batch maxima
#                                      (or args.max_doc_len / max_query_len under force_pad, :18-21)
def batchify_flat(q_tokens, q_offsets, d_tokens, d_offsets, B, N, Lq=None, Ld=None):
    """q_tokens / d_tokens: concatenated token ids; *_offsets: [B+1] / [B*N+1] starts.  Returns the four int64 tensors
    of the reference's batch dict (que_rep, que_len, doc_rep, doc_len)."""
    q_offsets = np.asarray(q_offsets, dtype=np.int64)
    d_offsets = np.asarray(d_offsets, dtype=np.int64)
    qlen = np.diff(q_offsets)
    dlen = np.diff(d_offsets).reshape(B, N)
    Lq = int(qlen.max()) if Lq is None else Lq
    Ld = int(dlen.max()) if Ld is None else Ld
    q = np.zeros((B, Lq), dtype=np.int64)
    d = np.zeros((B, N, Ld), dtype=np.int64)
    for b in range(B):
        q[b, :qlen[b]] = q_tokens[q_offsets[b]:q_offsets[b + 1]]
        for n in range(N):
            i = b * N + n
            d[b, n, :dlen[b, n]] = d_tokens[d_offsets[i]:d_offsets[i + 1]]
    return q, qlen.astype(np.int64), d, dlen.astype(np.int64)
