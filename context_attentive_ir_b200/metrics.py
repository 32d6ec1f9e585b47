"""On-device ranking metrics: the per-batch body of the reference's evaluation loops
(main/ranker.py:257-264, main/multitask.py:286-293) without the device->host copy of the scores.

    m = rank_metrics(scores, labels)            # scores: raw network output [B,N] on the GPU, labels [B,N]
    m['map'], m['mrr'], m['prec@1'], m['prec@3'], m['prec@5']   # the values the reference feeds its AverageMeters
"""
import torch

from . import lib

KEYS = ('map', 'mrr', 'prec@1', 'prec@3', 'prec@5')


def rank_metrics_device(scores, labels, apply_softmax=True):
    """Returns (batch_mean[5], per_row[B,5]) float64 CUDA tensors; nothing is synchronised."""
    if not scores.is_cuda:
        raise RuntimeError('scores must be a CUDA tensor (context_attentive_ir_b200 has no CPU path)')
    scores = scores.to(torch.float32).contiguous()
    if scores.dim() == 3:  # CARS: [B,S,N] -> [B*S,N] as main/multitask.py:286-287 does
        scores = scores.reshape(-1, scores.shape[-1])
    labels = labels.to(scores.device).to(torch.int64).reshape(scores.shape).contiguous()
    B, N = scores.shape
    per_row = torch.empty(B, 5, dtype=torch.float64, device=scores.device)
    mean = torch.empty(5, dtype=torch.float64, device=scores.device)
    with torch.cuda.device(scores.device):
        lib.check(lib.load().cair_rank_metrics(scores.data_ptr(), labels.data_ptr(), B, N, 1 if apply_softmax else 0,
                                               per_row.data_ptr(), mean.data_ptr(),
                                               torch.cuda.current_stream(scores.device).cuda_stream))
    return mean, per_row


def rank_metrics(scores, labels, apply_softmax=True):
    """dict(map, mrr, prec@1, prec@3, prec@5) of python floats (one 40-byte device->host copy)."""
    mean, _ = rank_metrics_device(scores, labels, apply_softmax)
    return dict(zip(KEYS, mean.cpu().tolist()))
