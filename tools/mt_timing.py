"""Role timing of the tcgen05 Match-Tensor interaction kernel (CTA 0): where do the cycles go?"""
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
names = ['producer wait empty_b', 'mma wait a_full', 'mma wait acc_empty', 'mma wait full_b', 'epi(w0) wait acc_full',
         'epi(w0) stage A + barrier', 'epi(w0) epilogue math', 'mma loop total (to last pair start)', 'pairs-1 of CTA 0']
with torch.no_grad():
    for _ in range(3):
        net(q, ql, d, dl)
    torch.cuda.synchronize()
    for skip in (0, 1):
        cnt = torch.zeros(16, dtype=torch.int64, device='cuda')
        cnt[15] = skip
        L.cair_mt_debug_timing(C.c_void_p(cnt.data_ptr()))
        net(q, ql, d, dl)
        torch.cuda.synchronize()
        L.cair_mt_debug_timing(None)
        print('--- epilogue math %s ---' % ('SKIPPED (scores are garbage: MMA stream alone)' if skip else 'on'))
        for n, v in zip(names, cnt.cpu().tolist()):
            print('%-40s %12d' % (n, v))
