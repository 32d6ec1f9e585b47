"""CARS cfg4: stage times (library profiler) against the forced sequences-per-cluster of the tcgen05 recurrence."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import helpers
from context_attentive_ir_b200 import synth, lib
dev = 'cuda:0'
L = C.CDLL(lib.LIB_PATH)
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def stages(reps=3):
    h = net.__dict__['_cair_handle']
    lib.check(lib.load().cair_profile_enable(h, 1))
    acc = {}
    tot = []
    for i in range(reps):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.score(*t); e1.record()
        torch.cuda.synchronize()
        tot.append(e0.elapsed_time(e1))
        names = C.create_string_buffer(2048); ms = (C.c_float * 64)(); cnt = C.c_int32()
        lib.check(lib.load().cair_profile_read(h, names, 2048, ms, 64, C.byref(cnt)))
        seen = {}
        for nm, v in zip(names.value.decode().split(','), list(ms)[:cnt.value]):
            k = nm if nm not in seen else nm + '#2'
            seen[k] = 1
            acc.setdefault(k, []).append(v)
    return min(tot), {k: round(float(np.mean(v)), 3) for k, v in acc.items()}

with torch.no_grad():
    net.score(*t); torch.cuda.synchronize()
    for spc in [0] + [int(x) for x in sys.argv[1].split(',')]:
        L.cair_rnn_set_seqs_per_cluster(8, spc)
        net.score(*t); torch.cuda.synchronize()
        tot, st = stages()
        print('spc %3d: total %.3f ms  %s' % (spc, tot, json.dumps(st)), flush=True)
    L.cair_rnn_set_seqs_per_cluster(8, 0)
    for impl in (0,):
        L.cair_set_rnn_impl(impl)
        net.score(*t); torch.cuda.synchronize()
        tot, st = stages()
        print('rnn impl %d: total %.3f ms  %s' % (impl, tot, json.dumps(st)), flush=True)
