"""Device-side batchify: the padded id / length tensors of the reference's batch dict
(neuroir/inputters/ranker/vector.py:39-90: que_rep, que_len, doc_rep, doc_len) built on the GPU from the ragged batch.

    q, qlen, d, dlen = batchify_ranker(q_tokens, q_offsets, d_tokens, d_offsets, B, N)   # all CUDA tensors
    scores = network(q, qlen, d, dlen)
"""
import torch

from . import lib


def batchify_ranker(q_tokens, q_offsets, d_tokens, d_offsets, B, N, max_query_len=None, max_doc_len=None, check=True):
    """q_tokens / d_tokens: int32 CUDA tensors of concatenated token ids; q_offsets [B+1] / d_offsets [B*N+1]: int64
    CUDA tensors.  max_query_len / max_doc_len: padded lengths (force_pad); default = the batch maxima (one small
    device->host read of the two maxima).  Returns (q [B,Lq], qlen [B], d [B,N,Ld], dlen [B,N]) int64 CUDA tensors."""
    dev = q_tokens.device
    if not q_tokens.is_cuda:
        raise RuntimeError('batchify_ranker works on CUDA tensors (context_attentive_ir_b200 has no CPU path)')
    q_tokens, d_tokens = q_tokens.to(torch.int32).contiguous(), d_tokens.to(torch.int32).contiguous()
    q_offsets, d_offsets = q_offsets.to(torch.int64).contiguous(), d_offsets.to(torch.int64).contiguous()
    if max_query_len is None or max_doc_len is None:
        mx = torch.stack([(q_offsets[1:] - q_offsets[:-1]).max(), (d_offsets[1:] - d_offsets[:-1]).max()]).tolist()
        max_query_len = int(mx[0]) if max_query_len is None else max_query_len
        max_doc_len = int(mx[1]) if max_doc_len is None else max_doc_len
    q = torch.empty(B, max_query_len, dtype=torch.int64, device=dev)
    qlen = torch.empty(B, dtype=torch.int64, device=dev)
    d = torch.empty(B, N, max_doc_len, dtype=torch.int64, device=dev)
    dlen = torch.empty(B, N, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        lib.check(lib.load().cair_batchify_ranker(q_tokens.data_ptr(), q_offsets.data_ptr(), d_tokens.data_ptr(),
                                                  d_offsets.data_ptr(), B, N, max_query_len, max_doc_len, q.data_ptr(),
                                                  qlen.data_ptr(), d.data_ptr(), dlen.data_ptr(), err.data_ptr(),
                                                  torch.cuda.current_stream(dev).cuda_stream))
    if check and int(err.item()):
        raise lib.CairError(-1, 'batchify_ranker: a sequence is empty or longer than its padded length')
    return q, qlen, d, dlen


def batchify_sessions(q_tokens, q_offsets, d_tokens, d_offsets, B, S, N, max_query_len=None, max_doc_len=None, check=True):
    """The ranking-side tensors of the multitask batch (neuroir/inputters/multitask/vector.py:82-149: source_words
    [B,S,Lq], source_lens [B,S], document_words [B,S,N,Ld], document_lens [B,S,N]) from the ragged batch of B sessions
    of S queries with N candidates each: the same kernel over B*S queries."""
    q, qlen, d, dlen = batchify_ranker(q_tokens, q_offsets, d_tokens, d_offsets, B * S, N, max_query_len, max_doc_len, check)
    return q.view(B, S, -1), qlen.view(B, S), d.view(B, S, N, -1), dlen.view(B, S, N)
