"""Phase clocks of the tcgen05 DRMM kernel (CTAs 0 and 1000, warp 0) at cfg3."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import helpers
from context_attentive_ir_b200 import lib, synth
L = C.CDLL(lib.LIB_PATH)
cfg = dict(model='drmm', emsize=300, src_vocab_size=131072, dropout_emb=0.2, nbins=5)
torch.manual_seed(1013)
net = helpers.build_module(cfg).to('cuda')
batch = synth.ranker_batch(1234, 256, 10, 20, 200, cfg['src_vocab_size'], variable=False)
t = helpers.to_dev(batch, 'cuda')
names = ['start', 'prologue done (Q image, barriers)', 'first ids + rows requested', 'chunk 0 written', 'chunk 4 written (tile 0 done)',
         'all chunks written', 'MMAs retired + barrier', 'histogram pass done', 'exact cells done', 'end', '  (tmem ld done)', '  (histogram loop done, before the barrier)']
with torch.no_grad():
    for _ in range(3):
        net(*t)
    cnt = torch.zeros(32, dtype=torch.int64, device='cuda')
    L.cair_drmm_debug_timing(C.c_void_p(cnt.data_ptr()))
    net(*t); torch.cuda.synchronize()
    L.cair_drmm_debug_timing(None)
    c = cnt.cpu().tolist()
    for base in (0, 16):
        print('CTA %d: %d cells recomputed exactly' % (1000 if base else 0, c[base + 12]))
        for i, n in enumerate(names):
            print('  %-42s %8d' % (n, c[base + i] - c[base]))
    for impl in (1, 0):
        L.cair_set_drmm_impl(impl)
        ts = []
        for i in range(10):
            torch.empty(192 << 20, dtype=torch.uint8, device='cuda').fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); net(*t); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('drmm impl %d: median %.4f ms' % (impl, sorted(ts)[5]))
