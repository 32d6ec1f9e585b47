"""Why is the interaction stage slower on realistic (mostly padded) batches?  Role counters of CTA 0 for full-length and
realistic inputs, with and without the epilogue math."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench, helpers
from context_attentive_ir_b200 import lib, synth
L = C.CDLL(lib.LIB_PATH)
torch.manual_seed(1013)
net = helpers.build_module(bench.CFG).to('cuda')
names = ['producer wait empty_b', 'mma wait a_full', 'mma wait acc_empty', 'mma wait full_b', 'epi(w0) wait acc_full',
         'epi(w0) stage A + barrier', 'epi(w0) epilogue math', 'mma loop total (to last pair start)', 'pairs-1 of CTA 0']
for label, kw in (('full lengths', dict(variable=False)), ('realistic', dict(realistic=True)), ('variable, no bos/eos', dict(variable=True))):
    b = synth.ranker_batch(1234, bench.B, bench.N, bench.LQ, bench.LD, bench.CFG['src_vocab_size'], **kw)
    q, ql, d, dl = helpers.to_dev(b, 'cuda')
    with torch.no_grad():
        for _ in range(3):
            net(q, ql, d, dl)
        torch.cuda.synchronize()
        for skip in (0, 1):
            cnt = torch.zeros(16, dtype=torch.int64, device='cuda')
            cnt[15] = skip
            L.cair_mt_debug_timing(C.c_void_p(cnt.data_ptr()))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); net(q, ql, d, dl); e1.record()
            torch.cuda.synchronize()
            L.cair_mt_debug_timing(None)
            print('--- %s, epilogue math %s: forward %.3f ms ---' % (label, 'SKIPPED' if skip else 'on', e0.elapsed_time(e1)))
            for n, v in zip(names, cnt.cpu().tolist()):
                print('    %-40s %12d' % (n, v))
