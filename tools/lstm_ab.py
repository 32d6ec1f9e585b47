"""Sequences per CTA of lstm_tc on the cfg2 shape: whole-forward time (CUDA events, warm)
and the per-kernel time of the document recurrence from the handle's profiler marks."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench, helpers
from context_attentive_ir_b200 import lib
L = C.CDLL(lib.LIB_PATH)
torch.manual_seed(1013)
net = helpers.build_module(bench.CFG).to('cuda')
batches = [helpers.to_dev(bench.make_batch(1236 + i), 'cuda') for i in range(8)]


def fwd_ms(iters=40):
    with torch.no_grad():
        for b in batches[:3]:
            net(*b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            net(*batches[i % len(batches)])
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for spc in (8, 16, 24, 32):
    L.cair_lstm_set_min_seqs_per_cta(spc)
    print('min seqs/CTA %2d: forward %.4f ms' % (spc, fwd_ms()), flush=True)
L.cair_lstm_set_min_seqs_per_cta(8)
