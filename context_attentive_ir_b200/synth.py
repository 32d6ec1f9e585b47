"""Seeded synthetic sessions of (query, N candidate documents).

The input layout contract is the reference's batchify() output
(neuroir/inputters/ranker/vector.py:39-90, neuroir/inputters/multitask/vector.py:82-149):
int64 ids with PAD=0, UNK=1, BOS=2, EOS=3 (neuroir/inputters/constants.py:1-4),
int64 lengths >= 1, every position >= length is PAD.

numpy-only on purpose: the golden-vector generator (oracle/gen_golden.py), the
parity tests and bench.py must all draw exactly the same inputs.
"""
import numpy as np

PAD, UNK, BOS, EOS = 0, 1, 2, 3


def _fill(rng, n, L, V, lens, lo=4, hi=None, bos_eos=False):
    hi = V if hi is None else hi
    ids = rng.integers(lo, hi, size=(n, L), dtype=np.int64)
    pos = np.arange(L)[None, :]
    ids[pos >= lens[:, None]] = PAD
    if bos_eos:
        ids[:, 0] = BOS
        ids[np.arange(n), lens - 1] = EOS
    return ids


def ranker_batch(seed, B, N, Lq, Ld, V, variable=True, bos_eos=False,
                 disjoint=False, overlap=0.0, realistic=False):
    """Returns dict(q[B,Lq], qlen[B], d[B,N,Ld], dlen[B,N], label[B,N]) of int64.

    variable=False: all lengths at max (headline throughput set).
    variable=True: qlen~U{2..Lq}, dlen~U{2..Ld}, element 0 forced to max so the
    padded dims equal Lq/Ld.
    disjoint=True: query ids in [4,V/2), doc ids in [V/2,V) (DRMM strict parity).
    overlap>0: that fraction of doc tokens is replaced by tokens of its query.
    """
    rng = np.random.default_rng(seed)
    if realistic:
        # the dataset's length statistics (reference README.md:80-83: avg query 3.84 / max 40 tokens, avg document 63.41 /
        # max 290, truncated to 200 by scripts/ranker.sh:18), plus the BOS / EOS the loader adds (inputters/ranker/utils.py:36,61)
        qlen = np.clip(rng.poisson(3.84, size=B) + 2, 3, Lq).astype(np.int64)
        dlen = np.clip(np.round(rng.lognormal(np.log(50.0), 0.65, size=B * N)) + 2, 3, Ld).astype(np.int64)
        qlen[0] = Lq
        dlen[0] = Ld
        bos_eos = True
    elif variable:
        qlen = rng.integers(2, Lq + 1, size=B, dtype=np.int64)
        dlen = rng.integers(2, Ld + 1, size=B * N, dtype=np.int64)
        qlen[0] = Lq
        dlen[0] = Ld
    else:
        qlen = np.full(B, Lq, dtype=np.int64)
        dlen = np.full(B * N, Ld, dtype=np.int64)
    if disjoint:
        q = _fill(rng, B, Lq, V, qlen, 4, V // 2, bos_eos=False)
        d = _fill(rng, B * N, Ld, V, dlen, V // 2, V, bos_eos=False)
    else:
        q = _fill(rng, B, Lq, V, qlen, bos_eos=bos_eos)
        d = _fill(rng, B * N, Ld, V, dlen, bos_eos=bos_eos)
    if overlap > 0:
        pick = rng.random((B * N, Ld)) < overlap
        src = rng.integers(0, Lq, size=(B * N, Ld))
        qrep = np.repeat(q, N, axis=0)
        cand = np.take_along_axis(qrep, src, axis=1)
        valid = (np.arange(Ld)[None, :] < dlen[:, None]) & (cand != PAD) & pick
        if bos_eos:
            valid[:, 0] = False
            valid[np.arange(B * N), dlen - 1] = False
        d = np.where(valid, cand, d)
    label = np.zeros((B, N), dtype=np.int64)
    label[np.arange(B), rng.integers(0, N, size=B)] = 1
    return dict(q=q, qlen=qlen, d=d.reshape(B, N, Ld), dlen=dlen.reshape(B, N),
                label=label)


def session_batch(seed, B, S, N, Lq, Ld, V, variable=True, max_clicks=1):
    """CARS batch: q[B,S,Lq], qlen[B,S], d[B,S,N,Ld], dlen[B,S,N], label[B,S,N] float32.

    Queries carry BOS/EOS, documents do not (neuroir/inputters/multitask/utils.py:40,56-57).
    Every (b,s) row has between 1 and max_clicks clicked documents.
    """
    rng = np.random.default_rng(seed)
    nq, nd = B * S, B * S * N
    if variable:
        qlen = rng.integers(2, Lq + 1, size=nq, dtype=np.int64)
        dlen = rng.integers(2, Ld + 1, size=nd, dtype=np.int64)
        qlen[0] = Lq
        dlen[0] = Ld
    else:
        qlen = np.full(nq, Lq, dtype=np.int64)
        dlen = np.full(nd, Ld, dtype=np.int64)
    q = _fill(rng, nq, Lq, V, qlen, bos_eos=True)
    d = _fill(rng, nd, Ld, V, dlen)
    label = np.zeros((nq, N), dtype=np.float32)
    for r in range(nq):
        k = int(rng.integers(1, max_clicks + 1))
        label[r, rng.choice(N, size=k, replace=False)] = 1.0
    return dict(q=q.reshape(B, S, Lq), qlen=qlen.reshape(B, S),
                d=d.reshape(B, S, N, Ld), dlen=dlen.reshape(B, S, N),
                label=label.reshape(B, S, N))
