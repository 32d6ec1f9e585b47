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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
