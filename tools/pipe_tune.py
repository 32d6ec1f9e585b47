"""Sweep of the pipeline split (share of a batch's pairs scored under the next batch's document encoder) for the
submit_host / wait_host serving loop, with and without the per-step L2 flush."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import bench, helpers
from context_attentive_ir_b200 import lib
dev = torch.device('cuda', 0)
torch.manual_seed(1013)
net = helpers.build_module(bench.CFG).to(dev)
batch = bench.make_batch(1236)
hq, hql, hd, hdl = [torch.from_numpy(np.ascontiguousarray(batch[k])).pin_memory() for k in ('q', 'qlen', 'd', 'dlen')]
houts = [torch.empty(bench.B, bench.N, dtype=torch.float32).pin_memory() for _ in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cstream = torch.cuda.Stream(dev)
L = lib.load()


def loop(k, do_flush, depth=3):
    with torch.cuda.stream(cstream):
        for i in range(k):
            if i >= depth:
                net.wait_host((i - depth) % 3)
            if do_flush:
                flush.fill_(i & 0xff)
            net.submit_host(hq, hql, hd, hdl, out=houts[i % 3], slot=i % 3, device=dev, stream=cstream)
        for i in range(max(0, k - depth), k):
            net.wait_host(i % 3)


import ctypes as C
Lb = C.CDLL(lib.LIB_PATH)
loop(6, True)
for spc, fracs in ((8, (0.0, 0.33)), (32, (0.33, 0.45, 0.55, 0.65, 0.75))):
  Lb.cair_lstm_set_min_seqs_per_cta(spc)
  for do_flush in (True, False):
    for depth in (3,):
        for frac in fracs:
            lib.check(L.cair_ranker_set_pipeline_split(net._cair_handle, frac))
            loop(6, do_flush, depth)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loop(200, do_flush, depth)
            dt = time.perf_counter() - t0
            print('min seqs/CTA=%d flush=%d depth=%d frac=%.2f: %.3f ms/step  %.3f M pairs/s' % (spc, do_flush, depth, frac, dt / 200 * 1e3, bench.B * bench.N * 200 / dt / 1e6), flush=True)
