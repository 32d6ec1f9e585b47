"""Timeline of the cross-batch software pipeline (submit_host / wait_host): per step, when do the document encoder,
the two interaction parts and the score copy of one batch run, relative to the begin of its encoder?"""
import ctypes as C, os, sys, time
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
Lb = C.CDLL(lib.LIB_PATH)
Lb.cair_ranker_pipeline_trace.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.33
for i in range(4):
    net.submit_host(hq, hql, hd, hdl, out=houts[i % 3], slot=i % 3, device=dev)
    if i >= 2:
        net.wait_host((i - 2) % 3)
for i in range(2, 4):
    net.wait_host(i % 3)
lib.check(lib.load().cair_ranker_set_pipeline_split(net._cair_handle, frac))
Lb.cair_ranker_pipeline_trace(net._cair_handle, 1, -1, None)
names = ['encoder end', 'part 1 begin', 'part 1 end', 'part 2 begin', 'part 2 end', 'finish']
K = 12
t0 = time.perf_counter()
for i in range(K):
    if i >= 3:
        net.wait_host(i % 3)
        ms = (C.c_float * 6)()
        Lb.cair_ranker_pipeline_trace(net._cair_handle, 1, i % 3, ms)
        print('step %2d: ' % (i - 3) + '  '.join('%s %.3f' % (n, v) for n, v in zip(names, ms)))
    net.submit_host(hq, hql, hd, hdl, out=houts[i % 3], slot=i % 3, device=dev)
for i in range(K - 3, K):
    net.wait_host(i % 3)
print('frac %.2f: %.3f ms/step (with tracing events)' % (frac, (time.perf_counter() - t0) / K * 1e3))
