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
with torch.no_grad():
    for _ in range(3):
        net(q, ql, d, dl)
    torch.cuda.synchronize()
    cnt = torch.zeros(16, dtype=torch.int64, device='cuda')
    L.cair_lstm_debug_timing(C.c_void_p(cnt.data_ptr()))
    net(q, ql, d, dl)
    torch.cuda.synchronize()
    L.cair_lstm_debug_timing(None)
names = ['mma wait x_full', 'mma wait bar_h', 'mma issue h part + commit', 'epi(w0) wait bar_acc',
         'epi(w0) tmem ld + cell + h operand + arrive', 'epi(w0) same + memory-bank stores', '-', 'gather wait x_empty']
print('(query encoder 20 steps + doc encoder 200 steps of CTA (0,0), cycles)')
for n, v in zip(names, cnt.cpu().tolist()):
    print('%-40s %12d   per step %8.0f' % (n, v, v / 220))
