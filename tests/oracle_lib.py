"""TEST INFRASTRUCTURE: ctypes loader for oracle/libcair_oracle.so (the CPU restatement) and
helpers to read the golden fixtures.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

from context_attentive_ir_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
_LIB = None


def build():
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, 'oracle', 'libcair_oracle.so')
        src = os.path.join(ROOT, 'oracle', 'cair_oracle.c')
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        _LIB = C.CDLL(path)
    return _LIB


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(bytes(z['meta']).decode())
    ins = {k[3:]: z[k] for k in z.files if k.startswith('in/')}
    sd = {k[3:]: np.ascontiguousarray(z[k], dtype=np.float32) for k in z.files if k.startswith('sd/')}
    outs = {k[4:]: z[k] for k in z.files if k.startswith('out/')}
    return meta['cfg'], ins, sd, outs


def host_getter(sd):
    def get(key):
        return sd[key].ctypes.data_as(_abi.f32p)
    return get


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_abi.i64p)


def _f32(a):
    return a.ctypes.data_as(_abi.f32p) if a is not None else None


def _check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed: %s' % (what, _abi.ERR_NAMES.get(rc, rc)))


def run_ranker(cfg, sd, q, qlen, d, dlen, want=()):
    """Oracle forward of one of the stand-alone rankers; returns dict(scores=..., +requested stage outputs)."""
    L = lib()
    model = cfg['model']
    w = _abi.PACKERS[model](cfg, host_getter(sd))
    B, Lq = q.shape
    _, N, Ld = d.shape
    q_, qp = _i64(q)
    ql_, qlp = _i64(qlen)
    d_, dp = _i64(d)
    dl_, dlp = _i64(dlen)
    scores = np.zeros((B, N), np.float32)
    out = dict(scores=scores)
    if model == 'esm':
        _check(L.cair_oracle_esm(C.byref(w), qp, qlp, dp, dlp, B, N, Lq, Ld, _f32(scores)), 'esm')
    elif model == 'match_tensor':
        eq = np.zeros((B, Lq, cfg['nhid_query']), np.float32) if 'enc_queries' in want else None
        ed = np.zeros((B * N, Ld, cfg['nhid_doc']), np.float32) if 'enc_docs' in want else None
        nx = int(cfg.get('nlayers', 1)) - 1
        get, bi = host_getter(sd), bool(cfg['bidirection'])
        extra = []
        for enc in ('query_encoder', 'document_encoder'):      # layers 1.. of the stacked encoders: [k][fwd, rev]
            arr = (_abi.LstmDir * max(2 * nx, 1))()
            for k in range(nx):
                arr[2 * k] = _abi._lstm(get, '%s.rnns.%d' % (enc, k + 1))
                if bi:
                    arr[2 * k + 1] = _abi._lstm(get, '%s.rnns.%d' % (enc, k + 1), '_reverse')
            extra.append(arr)
        _check(L.cair_oracle_mt_stacked(C.byref(w), extra[0], extra[1], nx, qp, qlp, dp, dlp, B, N, Lq, Ld, _f32(scores),
                                        _f32(eq), _f32(ed)), 'mt')
        out.update(enc_queries=eq, enc_docs=ed)
    elif model == 'drmm':
        hist = np.zeros((B * N, Lq, 5), np.int32) if 'hist' in want else None
        cos = np.zeros((B * N, Lq, Ld), np.float32) if 'cos' in want else None
        _check(L.cair_oracle_drmm(C.byref(w), qp, qlp, dp, dlp, B, N, Lq, Ld, _f32(scores),
                                  hist.ctypes.data_as(_abi.i32p) if hist is not None else None, _f32(cos)), 'drmm')
        out.update(hist=hist, cos=cos)
    elif model in ('dssm', 'cdssm', 'arci', 'arcii'):
        fn = getattr(L, 'cair_oracle_' + model)
        _check(fn(C.byref(w), qp, qlp, dp, dlp, B, N, Lq, Ld, _f32(scores)), model)
    elif model == 'duet':
        loc = np.zeros((B, N), np.float32) if 'local' in want else None
        _check(L.cair_oracle_duet(C.byref(w), qp, qlp, dp, dlp, B, N, Lq, Ld, _f32(scores), _f32(loc)), 'duet')
        out.update(local=loc)
    else:
        raise ValueError(model)
    return out


def run_cars_decode(cfg, sd, fwd, qlen, max_len, tgt2src, bos=2):
    """Oracle greedy decode from the outputs of run_cars (enc_q, sess_h, sess_c, sess_q_attn, sess_d_attn)."""
    L = lib()
    w = _abi.pack_cars(cfg, host_getter(sd))
    dw = _abi.pack_cars_decoder(cfg, host_getter(sd))
    B, S = fwd['sess_h'].shape[:2]
    Lq = fwd['enc_q'].shape[1]
    ql_, qlp = _i64(qlen)
    t2s, t2sp = _i64(tgt2src)
    pred = np.zeros((B, S - 1, max_len), np.int64)
    _check(L.cair_oracle_cars_decode(C.byref(w), C.byref(dw), _f32(fwd['enc_q']), qlp, _f32(fwd['sess_h']), _f32(fwd['sess_c']),
                                     _f32(fwd['sess_q_attn']), _f32(fwd['sess_d_attn']), B, S, Lq, max_len, t2sp, C.c_int64(bos),
                                     pred.ctypes.data_as(_abi.i64p)), 'cars_decode')
    return pred


def run_cars(cfg, sd, q, qlen, d, dlen, label):
    L = lib()
    w = _abi.pack_cars(cfg, host_getter(sd))
    B, S, Lq = q.shape
    N, Ld = d.shape[2], d.shape[3]
    q_, qp = _i64(q)
    ql_, qlp = _i64(qlen)
    d_, dp = _i64(d)
    dl_, dlp = _i64(dlen)
    lab = np.ascontiguousarray(label, dtype=np.float32)
    Hq, Hd = cfg['nhid_query'], cfg['nhid_document']
    out = dict(scores=np.zeros((B, S, N), np.float32), pooled_queries=np.zeros((B, S, Hq), np.float32),
               pooled_docs=np.zeros((B, S, N, Hd), np.float32), clicks=np.zeros((B, S, Hd), np.float32),
               sess_q_attn=np.zeros((B, S, cfg['nhid_session_query']), np.float32),
               sess_d_attn=np.zeros((B, S, cfg['nhid_session_document']), np.float32))
    hs = cfg['nhid_session_query'] + cfg['nhid_session_document']
    out.update(enc_q=np.zeros((B * S, Lq, Hq), np.float32), sess_h=np.zeros((B, S, hs), np.float32),
               sess_c=np.zeros((B, S, hs), np.float32))
    _check(L.cair_oracle_cars_ex(C.byref(w), qp, qlp, dp, dlp, _f32(lab), B, S, N, Lq, Ld, _f32(out['scores']),
                                 _f32(out['pooled_queries']), _f32(out['pooled_docs']), _f32(out['clicks']),
                                 _f32(out['sess_q_attn']), _f32(out['sess_d_attn']), _f32(out['enc_q']), _f32(out['sess_h']),
                                 _f32(out['sess_c'])), 'cars')
    return out


def run_lstm(x, lens, fwd, rev, h, rnn_type=0):
    """fwd/rev: dict(w_ih,w_hh,b_ih,b_hh) of float32 arrays (rev may be None); rnn_type 0 = LSTM, 1 = GRU."""
    L = lib()
    n, T, inp = x.shape
    x = np.ascontiguousarray(x, np.float32)
    l_, lp = _i64(lens)

    def mk(dd):
        return _abi.LstmDir(*[_f32(np.ascontiguousarray(dd[k], np.float32)) for k in ('w_ih', 'w_hh', 'b_ih', 'b_hh')])
    keep = [fwd, rev]
    f = mk(fwd)
    r = mk(rev) if rev is not None else None
    dirs = 2 if rev is not None else 1
    out = np.zeros((n, T, dirs * h), np.float32)
    hn = np.zeros((dirs, n, h), np.float32)
    cn = np.zeros((dirs, n, h), np.float32)
    _check(L.cair_oracle_rnn(rnn_type, _f32(x), lp, n, T, inp, h, C.byref(f), C.byref(r) if r is not None else None,
                             _f32(out), _f32(hn), _f32(cn)), 'rnn')
    del keep
    return out, hn, cn


def rel_err(a, ref):
    """|a-ref| / max(|ref|, 1e-2*scale), scale = max|ref| over the batch: scores can be ~0, where a
    pure relative error is meaningless."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = max(float(np.max(np.abs(ref))), 1e-12)
    return np.abs(a - ref) / np.maximum(np.abs(ref), 1e-2 * scale + 1e-30)
