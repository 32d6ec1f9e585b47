#!/usr/bin/env python
"""Secondary bench: the other configs of BASELINE.json (ESM cfg1, DRMM cfg3, CARS cfg4, DUET cfg5) on one B200.
Device-resident inputs, CUDA events, L2 flushed between steps.  One JSON line per model (pairs/s, ms/step and, for
the HBM-bound models, algorithmic GB/s against the measured copy bandwidth)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import helpers
from context_attentive_ir_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=20)
ap.add_argument('--warmup', type=int, default=3)
ap.add_argument('--models', default='gather,esm,esm300,drmm,duet,cars')
args = ap.parse_args()
dev = 'cuda:0'
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
HBM = peaks.get('hbm_gbs', 6650.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for i in range(args.steps):
        flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.mean(ms)), float(np.min(ms))


def stages_of(net, fn, reps=3):
    """Per-stage CUDA-event times (ms) of the library's own profiler, averaged over a few flushed steps."""
    import ctypes as C
    from context_attentive_ir_b200 import lib
    L = lib.load()
    h = net.__dict__['_cair_handle']
    lib.check(L.cair_profile_enable(h, 1))
    acc = {}
    for i in range(reps):
        flush.fill_(i)
        fn()
        torch.cuda.synchronize()
        names = C.create_string_buffer(2048)
        ms = (C.c_float * 64)()
        cnt = C.c_int32()
        lib.check(L.cair_profile_read(h, names, 2048, ms, 64, C.byref(cnt)))
        for nm, v in zip(names.value.decode().split(','), list(ms)[:cnt.value]):
            acc.setdefault(nm, []).append(v)
    lib.check(L.cair_profile_enable(h, 0))
    return {k: round(float(np.mean(v)), 4) for k, v in acc.items()}


def ranker(name, cfg, B, N, Lq, Ld, bytes_per_pair=None, **kw):
    torch.manual_seed(1013)
    net = helpers.build_module(cfg).to(dev)
    batch = synth.ranker_batch(1234, B, N, Lq, Ld, cfg['src_vocab_size'], variable=False, **kw)
    t = helpers.to_dev(batch, dev)
    with torch.no_grad():
        mean, best = timeit(lambda: net(*t))
    out = dict(model=name, config=dict(B=B, N=N, Lq=Lq, Ld=Ld, E=cfg['emsize'], V=cfg['src_vocab_size']),
               pairs_per_s=B * N / (mean / 1e3), ms_per_step=mean, ms_best=best)
    if bytes_per_pair:
        gbs = bytes_per_pair * B * N / (mean / 1e3) / 1e9
        out.update(hbm_algorithmic_gbs=gbs, hbm_peak_gbs=HBM, hbm_frac=gbs / HBM, bytes_per_pair=bytes_per_pair)
    print(json.dumps(out), flush=True)


def bpp(E, Lq, Ld, N):  # int64 ids, fp32 table rows, one fp32 score
    return Ld * (8 + E * 4) + (Lq * (8 + E * 4) + 8) / N + 12

def gather_bench(V, E, T):
    import ctypes as C
    from context_attentive_ir_b200 import lib
    torch.manual_seed(3)
    table = torch.randn(V, E, device=dev)
    ids = torch.randint(4, V, (T,), device=dev, dtype=torch.int64)
    out = torch.empty(T, E, device=dev)
    L = lib.load()
    fn = lambda: lib.check(L.cair_embed_gather(table.data_ptr(), V, E, ids.data_ptr(), T, out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
    mean, best = timeit(fn)
    moved = T * (8 + 2 * E * 4)   # id + row read + row written
    print(json.dumps(dict(model='embed_gather', config=dict(V=V, E=E, tokens=T), tokens_per_s=T / (mean / 1e3), ms_per_step=mean,
                          ms_best=best, hbm_gbs=moved / (mean / 1e3) / 1e9, hbm_peak_gbs=HBM, hbm_frac=moved / (mean / 1e3) / 1e9 / HBM,
                          bytes_moved=moved)), flush=True)


for m in args.models.split(','):
    if m == 'gather':
        gather_bench(131072, 300, 512000)   # the doc tokens of cfg3 (B=256, N=10, Ld=200)
    elif m == 'esm':
        ranker('esm cfg1', dict(model='esm', emsize=64, src_vocab_size=10000), 8, 5, 10, 50, bpp(64, 10, 50, 5))
    elif m == 'esm300':
        ranker('esm E=300 (cfg3 shape)', dict(model='esm', emsize=300, src_vocab_size=131072), 256, 10, 20, 200, bpp(300, 20, 200, 10))
    elif m == 'drmm':
        ranker('drmm cfg3', dict(model='drmm', emsize=300, src_vocab_size=131072, dropout_emb=0.2, nbins=5), 256, 10, 20, 200,
               bpp(300, 20, 200, 10))
    elif m == 'duet':
        cfg = dict(model='duet', emsize=300, src_vocab_size=131072, dropout_emb=0.2, dropout=0.2, use_word=True, nfilters=300,
                   local_filter_size=1, dist_filter_size=3, pool_size=5, max_doc_len=200, max_query_len=20)
        for N in (10, 50):
            ranker('duet cfg5 N=%d' % N, cfg, 32, N, 20, 200)
    elif m == 'cars':
        cfg = dict(model='cars', emsize=300, src_vocab_size=131072, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2, rnn_type='LSTM',
                   bidirection=True, nlayers=1, nhid_query=256, nhid_document=256, nhid_click=512, nhid_session_query=512,
                   nhid_session_document=512, nhid_decoder=512, query_session_off=False, doc_session_off=False, dropout_rnn=0.2,
                   attn_type='general', mlp_nhid=150, pool_type='attn', regularize_coeff=0.1, alpha=0.1, lambda1=0.01,
                   lambda2=0.0001, turn_ranker_off=False, turn_recommender_off=False)
        torch.manual_seed(1013)
        net = helpers.build_module(cfg).to(dev)
        B, S, N, Lq, Ld = 32, 7, 10, 20, 200
        batch = synth.session_batch(1238, B, S, N, Lq, Ld, cfg['src_vocab_size'], variable=False, max_clicks=2)
        t = helpers.to_dev(batch, dev, ('q', 'qlen', 'd', 'dlen', 'label'))
        with torch.no_grad():
            mean, best = timeit(lambda: net.score(*t))
            st = stages_of(net, lambda: net.score(*t))
        print(json.dumps(dict(model='cars cfg4 (ranking path)', config=dict(B=B, S=S, N=N, Lq=Lq, Ld=Ld, E=300, H=256),
                              pairs_per_s=B * S * N / (mean / 1e3), ms_per_step=mean, ms_best=best, stages_ms=st)), flush=True)
