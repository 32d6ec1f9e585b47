"""TEST INFRASTRUCTURE - numpy restatement of the dropout keep-scale hash of csrc/train.cu (drop_scale): splitmix64 of
(seed, element index), keep when the top 24 bits / 2^24 >= p, scale 1/(1-p).  Used by oracle/gen_golden.py to run the
reference's train-mode forward with exactly the mask the CUDA kernels draw, and by the tests to pin cair_dropout_mask."""
import numpy as np


def drop_scale(seed, n, p):
    if p <= 0:
        return np.ones(n, dtype=np.float32)
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over='ignore'):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(u >= np.float32(p), inv, np.float32(0.0)).astype(np.float32)
