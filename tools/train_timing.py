"""Match-Tensor training step at the BASELINE cfg2 shape: forward / backward / optimizer times (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import helpers
from context_attentive_ir_b200 import synth
cfg = dict(model='match_tensor', emsize=300, src_vocab_size=131072, dropout_emb=0.2, rnn_type='LSTM', bidirection=True,
           nlayers=1, dropout_rnn=0.2, featsize=40, nhid_query=128, nhid_doc=128, nchannels=50, nfilters=6, match_filter_size=20)
dev = 'cuda:0'
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
fix = len(sys.argv) > 2 and sys.argv[2] == 'fix'
torch.manual_seed(5)
net = helpers.build_module(cfg).to(dev).train()
if fix:
    net.word_embeddings.word_lut.weight.requires_grad = False
batch = synth.ranker_batch(99, B, 10, 20, 200, cfg['src_vocab_size'], variable=False)
q, ql, d, dl = helpers.to_dev(batch, dev)
labels = torch.from_numpy(batch['label']).float().to(dev)
opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], 0.05)
crit = torch.nn.BCEWithLogitsLoss()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tf = tb = to = 0.0
n = 6
for it in range(n + 2):
    ev[0].record()
    loss = crit(net(q, ql, d, dl), labels)
    ev[1].record()
    opt.zero_grad()
    loss.backward()
    ev[2].record()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 5.0)
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2]); to += ev[2].elapsed_time(ev[3])
print('B=%d N=10 (%d pairs)%s: forward %.2f ms, backward %.2f ms, clip+SGD %.2f ms => %.0f pairs/s; loss %.4f; peak mem %.2f GB' % (
    B, B * 10, ' fixed embeddings' if fix else '', tf / n, tb / n, to / n, B * 10 / ((tf + tb + to) / n) * 1e3, float(loss),
    torch.cuda.max_memory_allocated() / 2**30))
