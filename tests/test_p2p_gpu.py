"""cair_allgather_scores (include/cair.h): the score all-gather as our own kernel over NVLink peer memory, checked against
NCCL's all_gather on 2 GPUs (skipped on a single-GPU box; the sharding logic itself is covered on CPU by test_sharding_gloo)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_p2p_allgather_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs with peer access')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tools', 'p2p_test.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'identical=True' in r.stdout and 'pipelined gather inside submit_host / wait_host: identical=True' in r.stdout
