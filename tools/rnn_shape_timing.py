"""Role timing and sequences-per-cluster sweep of the cluster-split recurrence through cair_rnn_forward on one shape."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from context_attentive_ir_b200 import lib, _abi
n, Lx, inp, h = [int(x) for x in sys.argv[1:5]]
rt = int(sys.argv[5]) if len(sys.argv) > 5 else 0
G = 3 if rt else 4
L = C.CDLL(lib.LIB_PATH)
lib.load()
dev = 'cuda'
g = torch.Generator().manual_seed(5)
k = 1.0 / np.sqrt(h)
x = torch.randn(n, Lx, inp, generator=g).to(dev)
lens = torch.full((n,), Lx, dtype=torch.int64, device=dev)
def mk():
    w = [(torch.rand(G * h, inp, generator=g) * 2 - 1) * k, (torch.rand(G * h, h, generator=g) * 2 - 1) * k,
         (torch.rand(G * h, generator=g) * 2 - 1) * k, (torch.rand(G * h, generator=g) * 2 - 1) * k]
    w = [t.to(dev) for t in w]
    return w, _abi.LstmDir(*[C.cast(t.data_ptr(), _abi.f32p) for t in w])
wf, f = mk(); wr, r = mk()
out = torch.empty(n, Lx, 2 * h, device=dev)
names = ['mma wait x_full', 'mma wait bar_h (all blocks)', 'mma issue h part', 'epi(w0) wait bar_acc',
         'epi(w0) tmem ld + cells + h exchange + arrive', 'epi(w0) same + memory-bank stores', '-', 'gather wait x_empty']

def call():
    lib.check(lib.load().cair_rnn_forward(rt, x.data_ptr(), lens.data_ptr(), n, Lx, inp, h, C.byref(f), C.byref(r), out.data_ptr(),
                                          None, None, torch.cuda.current_stream().cuda_stream))

def timed(reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

L.cair_set_rnn_impl(3)
call(); torch.cuda.synchronize()
cnt = torch.zeros(192, dtype=torch.int64, device=dev)
L.cair_rnn_debug_timing(C.c_void_p(cnt.data_ptr()))
call(); torch.cuda.synchronize()
L.cair_rnn_debug_timing(None)
c = cnt.cpu().tolist()
for nm, v in zip(names, c):
    print('%-48s %12d   per step %8.0f' % (nm, v, v / Lx))
t0 = c[16]
print('step-100 timeline of CTA (0,0): mma own %d, 2nd %d, issue done %d | next begin %d, own %d, 2nd %d, done %d' % tuple(v - t0 for v in c[17:24]))
for w in range(20):
    e = c[24 + 6 * w: 30 + 6 * w]
    e = [e[0], e[1], e[3], e[4], e[5], e[2]]
    print('  epi warp %2d: acc seen %6d  first ld %6d  cells+stores %6d  fenced %6d  bulk issued %6d  arrived %6d' % ((w,) + tuple(v - t0 for v in e)))
print('whole call (pack + pre-gate GEMM + recurrence), default plan: %.3f ms' % timed())
for spc in [int(s) for s in (sys.argv[6].split(',') if len(sys.argv) > 6 else [])]:
    L.cair_rnn_set_seqs_per_cluster(8, spc)
    print('forced %3d sequences per cluster: %.3f ms' % (spc, timed()))
L.cair_rnn_set_seqs_per_cluster(8, 0)
for impl in (1, 0):
    L.cair_set_rnn_impl(impl)
    print('impl %d: %.3f ms' % (impl, timed()))
