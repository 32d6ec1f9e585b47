"""Role timing of the tcgen05 LSTM kernel (CTA (0,0)) on the cfg2 document encoder shape."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench, helpers
from context_attentive_ir_b200 import lib
L = C.CDLL(lib.LIB_PATH)
torch.manual_seed(1013)
net = helpers.build_module(bench.CFG).to('cuda')
q, ql, d, dl = helpers.to_dev(bench.make_batch(1236), 'cuda')
names = ['mma wait x_full', 'mma wait bar_h', 'mma issue h part + commit', 'epi(w0) wait bar_acc',
         'epi(w0) tmem ld + cell + h operand + arrive', 'epi(w0) same + memory-bank stores', '-', 'gather wait x_empty']


def run(cold):
    cnt = torch.zeros(96, dtype=torch.int64, device='cuda')
    if cold:
        torch.empty(256 << 20, dtype=torch.uint8, device='cuda').fill_(1)   # evict L2
    torch.cuda.synchronize()
    L.cair_lstm_debug_timing(C.c_void_p(cnt.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net(q, ql, d, dl); e1.record()
    torch.cuda.synchronize()
    L.cair_lstm_debug_timing(None)
    print('%s L2: forward %.3f ms (query encoder 20 steps + doc encoder 200 steps of CTA (0,0), cycles)'
          % ('cold' if cold else 'warm', e0.elapsed_time(e1)))
    for n, v in zip(names, cnt.cpu().tolist()):
        print('%-48s %12d   per step %8.0f' % (n, v, v / 220))
    c = cnt.cpu().tolist()
    t0 = c[16]
    print('step-100 timeline of CTA (1,0), cycles after the MMA warp saw bar_h:')
    print('  mma: issue done %d | next step bar_h seen %d, issue done %d' % (c[17] - t0, c[18] - t0, c[19] - t0))
    for w in range(16):
        e = c[24 + 4 * w: 28 + 4 * w]
        print('  epi warp %2d: acc seen %5d  tmem ld done %5d  cell+h stores done %5d  arrived %5d' % ((w,) + tuple(x - t0 for x in e)))
    print('CTA(0,0) both kernels: %d cycles in %d ns -> %.3f GHz' % (c[8], c[9], c[8] / max(c[9], 1)))


with torch.no_grad():
    for _ in range(3):
        net(q, ql, d, dl)
    run(False)
    run(True)
    run(True)
