"""world_size-2 gloo test of the doc-parallel plumbing (partition + one all-gather), with the CPU
oracle standing in for the per-rank scorer (no GPU here)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, name, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from context_attentive_ir_b200.parallel import ShardedCars, ShardedRanker, ShardedSessionRanker
    cfg, ins, sd, outs = ol.load_golden(name)
    try:
        if cfg['model'] == 'mnsrf':
            sys.path.insert(0, os.path.join(os.path.dirname(HERE), 'oracle'))
            import session_models_oracle as smo

            def score(q, ql, d, dl, b, c):
                o = smo.mnsrf_rank(sd, q.numpy(), ql.numpy(), d.numpy(), dl.numpy())[0].astype(np.float32)
                full = np.zeros_like(o)
                full[b:b + c] = o[b:b + c]
                return torch.from_numpy(full)
            sh = ShardedSessionRanker(None, score_slice=score)
            got = sh(*[torch.from_numpy(ins[k]) for k in ('q', 'qlen', 'd', 'dlen')])
        elif cfg['model'] == 'cars':
            def score(q, ql, d, dl, lab, b, c):
                o = ol.run_cars(cfg, sd, q.numpy(), ql.numpy(), d.numpy(), dl.numpy(), lab.numpy())['scores']
                full = np.zeros_like(o)
                full[b:b + c] = o[b:b + c]   # a rank only contributes its own sessions
                return torch.from_numpy(full)
            sh = ShardedCars(None, score_slice=score)
            got = sh(*[torch.from_numpy(ins[k]) for k in ('q', 'qlen', 'd', 'dlen', 'label')])
        else:
            def score(q, ql, d, dl, b, c):
                o = ol.run_ranker(cfg, sd, q.numpy(), ql.numpy(), d.numpy(), dl.numpy())['scores']
                full = np.zeros_like(o).reshape(-1)
                full[b:b + c] = o.reshape(-1)[b:b + c]
                return torch.from_numpy(full.reshape(o.shape))
            sh = ShardedRanker(None, score_slice=score)
            got = sh(*[torch.from_numpy(ins[k]) for k in ('q', 'qlen', 'd', 'dlen')])
        ret[rank] = float(np.abs(got.numpy() - outs['scores']).max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name,world', [('mt_tiny', 2), ('esm_cfg1', 2), ('cars_clicks', 2), ('esm_cfg1', 3), ('mnsrf_small', 2)])
def test_sharded_scores_match_unsharded(name, world):
    ret = mp.Manager().dict()
    port = 29500 + (os.getpid() + hash(name) + world) % 2000
    mp.spawn(_worker, args=(world, port, name, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-4, (r, ret[r])
