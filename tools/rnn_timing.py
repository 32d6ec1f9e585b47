"""Role timing of the cluster-split tcgen05 recurrence (CTA (0,0)) on the cfg2 document encoder shape, and a sweep of
the sequences-per-cluster knob."""
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
names = ['mma wait x_full', 'mma wait bar_h (all blocks)', 'mma issue h part', 'epi(w0) wait bar_acc',
         'epi(w0) tmem ld + cells + h exchange + arrive', 'epi(w0) same + memory-bank stores', '-', 'gather wait x_empty']


def stage_ms(reps=20):

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        torch.empty(192 << 20, dtype=torch.uint8, device='cuda').fill_(1)
        e0.record(); net(q, ql, d, dl); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def run():
    cnt = torch.zeros(192, dtype=torch.int64, device='cuda')
    torch.cuda.synchronize()
    L.cair_rnn_debug_timing(C.c_void_p(cnt.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net(q, ql, d, dl); e1.record()
    torch.cuda.synchronize()
    L.cair_rnn_debug_timing(None)
    print('forward %.3f ms (query encoder 20 steps + doc encoder 200 steps of CTA (0,0), cycles)' % e0.elapsed_time(e1))
    c = cnt.cpu().tolist()
    for n, v in zip(names, c):
        print('%-48s %12d   per step %8.0f' % (n, v, v / 220))
    t0 = c[16]
    print('step-100 timeline of CTA (0,0), cycles after the MMA warp began waiting for h blocks:')
    print('  mma: own block seen %d, 2nd block seen %d, issue+commit done %d | next step: begin %d, own %d, 2nd %d, done %d'
          % tuple(x - t0 for x in c[17:24]))
    for w in range(20):
        e = c[24 + 6 * w: 30 + 6 * w]
        e = [e[0], e[1], e[3], e[4], e[5], e[2]]
        print('  epi warp %2d: acc seen %5d  first tmem ld done %5d  cells + local stores done %5d  fenced %5d  bulk copies issued %5d  arrived %5d'
              % ((w,) + tuple(x - t0 for x in e)))


with torch.no_grad():
    for _ in range(3):
        net(q, ql, d, dl)
    L.cair_set_rnn_impl(3)
    run()
    for impl in (3, 1):
        L.cair_set_rnn_impl(impl)
        print('rnn impl %d: forward median %.4f ms' % (impl, stage_ms()))
    L.cair_set_rnn_impl(3)
    for spc in (24, 28, 32, 35, 36, 40, 48, 64):
        L.cair_rnn_set_seqs_per_cluster(8, spc)
        print('forced %d sequences per cluster: forward median %.4f ms' % (spc, stage_ms()))
    L.cair_rnn_set_seqs_per_cluster(8, 0)
    L.cair_set_rnn_impl(2)
