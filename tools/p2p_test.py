"""torchrun --nproc-per-node N tools/p2p_test.py: cair_allgather_scores (NVLink peer stores) against NCCL all_gather."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from context_attentive_ir_b200.parallel import P2PScoreGather, gather_scores
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
ok = True
for per in (1280, 7, 16000):
    g = P2PScoreGather(per, dev)
    total = per * world
    for it in range(40):
        x = torch.randn(per, device=dev) + rank * 1000 + it
        a = g(x, total).clone()
        b = gather_scores(x, total)
        ok &= bool(torch.equal(a, b))
    def timeit(fn, n=300):
        for _ in range(20): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    x = torch.randn(per, device=dev)
    t_p2p = timeit(lambda: g(x, total))
    t_nccl = timeit(lambda: gather_scores(x, total))
    if rank == 0:
        print('per=%d world=%d: identical=%s  p2p %.1f us  nccl (+ pad/slice helpers) %.1f us per call' % (per, world, ok, t_p2p, t_nccl), flush=True)
# the same gather INSIDE the serving pipeline (cair_ranker_set_gather): submit_host / wait_host return all ranks' scores
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import helpers
from context_attentive_ir_b200 import synth
cfg = dict(model='match_tensor', emsize=300, src_vocab_size=5000, dropout_emb=0.2, rnn_type='LSTM', bidirection=True, nlayers=1,
           dropout_rnn=0.2, featsize=40, nhid_query=128, nhid_doc=128, nchannels=50, nfilters=6, match_filter_size=20)
torch.manual_seed(7)
net = helpers.build_module(cfg).to(dev)
B, N = 16, 10
batches = [synth.ranker_batch(100 * rank + i, B, N, 20, 200, cfg['src_vocab_size'], bos_eos=True) for i in range(7)]
pinned = [[torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in ('q', 'qlen', 'd', 'dlen')] for b in batches]
with torch.no_grad():
    ref = [gather_scores(net(*helpers.to_dev(b, dev)).reshape(-1), B * N * world).cpu() for b in batches]
P2PScoreGather(B * N, dev).attach(net)
outs = [torch.empty(world * B * N).pin_memory() for _ in range(3)]
got = []
for i in range(len(batches)):
    if i >= 3:
        net.wait_host(i % 3)
        got.append(outs[i % 3].clone())
    net.submit_host(*pinned[i], out=outs[i % 3], slot=i % 3, device=dev)
for i in range(max(0, len(batches) - 3), len(batches)):
    net.wait_host(i % 3)
    got.append(outs[i % 3].clone())
pipe_ok = all(torch.equal(a, b) for a, b in zip(got, ref))
if rank == 0:
    print('pipelined gather inside submit_host / wait_host: identical=%s (%d batches)' % (pipe_ok, len(got)), flush=True)
ok &= pipe_ok
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
