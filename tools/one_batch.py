"""Three forwards of the headline batch (cfg2), for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench, helpers
torch.manual_seed(1013)
net = helpers.build_module(bench.CFG).to('cuda')
args = helpers.to_dev(bench.make_batch(1236), 'cuda')
with torch.no_grad():
    for _ in range(3):
        net(*args)
torch.cuda.synchronize()
