"""Timing experiments on the persistent tcgen05 GEMM at the CARS pre-gate shape (M = 448k gathered rows, K = 300, N = 1024)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import helpers
from context_attentive_ir_b200 import lib, synth
L = C.CDLL(lib.LIB_PATH)
cfg = dict(model='cars', emsize=300, src_vocab_size=131072, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2, rnn_type='LSTM',
           bidirection=True, nlayers=1, nhid_query=256, nhid_document=256, nhid_click=512, nhid_session_query=512,
           nhid_session_document=512, nhid_decoder=512, query_session_off=False, doc_session_off=False, dropout_rnn=0.2,
           attn_type='general', mlp_nhid=150, pool_type='attn', regularize_coeff=0.1, alpha=0.1, lambda1=0.01, lambda2=0.0001,
           turn_ranker_off=False, turn_recommender_off=True)
torch.manual_seed(1013)
net = helpers.build_module(cfg).to('cuda')
batch = synth.session_batch(1238, 32, 7, 10, 20, 200, cfg['src_vocab_size'], variable=False, max_clicks=2)
t = helpers.to_dev(batch, 'cuda', ('q', 'qlen', 'd', 'dlen', 'label'))
def run(label):
    ts = []
    with torch.no_grad():
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); net.score(*t); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    print('%-40s forward median %.3f ms' % (label, sorted(ts)[3]), flush=True)
run('baseline (persistent gemm)')
for bits, label in ((1, 'no epilogue stores'), (4, 'short epilogue (1 of 8 chunks)'), (2, 'no W traffic'), (6, 'no W traffic + short epilogue')):
    L.cair_debug_gemm(bits); run(label)
L.cair_debug_gemm(0)
L.cair_set_gemm_impl(2); run('one tile per CTA kernel (round 1)')
