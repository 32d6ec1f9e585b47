"""MNSRF and M-Match-Tensor ranking paths: the numpy oracle against the reference fixtures (CPU),
the CUDA path against both (GPU), session sharding, and the call sequence of the reference's Multitask.predict."""
import os
import sys

import numpy as np
import pytest
import torch

import helpers
import oracle_lib as ol

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import session_models_oracle as smo  # noqa: E402

DEV = 'cuda:0'
TOL = 1e-3


class _TgtDict:
    """tgt_dict[i] -> word (the reference's Vocabulary indexing)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return 'w%d' % i


class _SrcDict:
    """src_dict[word] -> id, from the fixture's target -> source map."""

    def __init__(self, tgt2src):
        self.m = {'w%d' % i: int(v) for i, v in enumerate(tgt2src)}

    def __getitem__(self, w):
        return self.m[w]


def _rel(a, ref):
    return float(ol.rel_err(np.asarray(a), np.asarray(ref)).max())


@pytest.mark.parametrize('name', ['mnsrf_tiny', 'mnsrf_small'])
def test_mnsrf_oracle_matches_reference_fixture(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    scores, mem, sess = smo.mnsrf_rank(sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'])
    assert np.abs(mem - outs['memory_bank']).max() < 2e-5
    assert np.abs(sess - outs['session_bank']).max() < 2e-5
    assert _rel(scores, outs['scores']) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['mnsrf_tiny', 'mnsrf_small', 'mnsrf_h256'])
def test_mnsrf_golden(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, DEV)
    q, ql, d, dl = helpers.to_dev(ins, DEV)
    with torch.no_grad():
        out = net.score(q, ql, d, dl, want_banks=True)
    torch.cuda.synchronize()
    net.poll_error()
    assert np.abs(out['memory_bank'].cpu().numpy() - outs['memory_bank']).max() < 5e-4
    assert np.abs(out['session_bank'].cpu().numpy() - outs['session_bank']).max() < 5e-4
    assert _rel(out['scores'].cpu().numpy(), outs['scores']) < TOL
    # session shards assemble to the full result (sessions are independent)
    B = q.shape[0]
    parts = torch.zeros_like(out['scores'])
    with torch.no_grad():
        for b in range(B):
            parts += net.score(q, ql, d, dl, session_slice=(b, 1))['scores']
    assert torch.equal(parts, out['scores'])
    # the reference's predict-time sequence (models/multitask.py:270-276)
    with torch.no_grad():
        mb, sb, states = net.encode(q, ql)
        s2 = net.rank_document(q, mb, sb, d, dl)
    assert torch.equal(s2, out['scores'])
    with torch.no_grad():
        dec = net.decode(states=states, max_len=6, src_dict=_SrcDict(ins['tgt2src']), tgt_dict=_TgtDict(len(ins['tgt2src'])),
                         batch_size=B, session_len=q.shape[1] - 1, use_cuda=True)
    assert np.array_equal(dec['predictions'].cpu().numpy(), outs['predictions'])     # greedy tokens identical to the reference


@pytest.mark.gpu
def test_mnsrf_fresh_inputs_vs_oracle():
    from context_attentive_ir_b200 import synth
    cfg, _, sd, _ = ol.load_golden('mnsrf_small')
    net = helpers.build_module(cfg, sd, DEV)
    batch = synth.session_batch(4321, 3, 4, 3, 7, 19, cfg['src_vocab_size'], max_clicks=1)
    ref, _, _ = smo.mnsrf_rank(sd, batch['q'], batch['qlen'], batch['d'], batch['dlen'])
    with torch.no_grad():
        out = net.score(*helpers.to_dev(batch, DEV))
    assert _rel(out['scores'].cpu().numpy(), ref) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize('impl', ['tc', 'fp32'])
@pytest.mark.parametrize('name', ['mmt_tiny', 'mmt_arch'])
def test_m_match_tensor_golden(name, impl):
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, DEV)
    net.__dict__['_cair_impl'] = {'fp32': 0, 'tc': 1}[impl]
    q, ql, d, dl = helpers.to_dev(ins, DEV)
    with torch.no_grad():
        mb, sb, states = net.encode(q, ql)
        s = net.rank_document(q, mb, sb, d, dl)
    torch.cuda.synchronize()
    net.poll_error()
    assert s.shape == outs['scores'].shape
    assert _rel(s.cpu().numpy(), outs['scores']) < TOL
    with torch.no_grad():
        dec = net.decode(states=states, max_len=6, src_dict=_SrcDict(ins['tgt2src']), tgt_dict=_TgtDict(len(ins['tgt2src'])),
                         batch_size=q.shape[0], session_len=q.shape[1] - 1, use_cuda=True)
    assert np.abs(net._fwd['session_bank'].cpu().numpy() - outs['session_bank']).max() < 5e-4
    assert np.array_equal(dec['predictions'].cpu().numpy(), outs['predictions'])
    # the same numbers as the stand-alone Match-Tensor oracle on the flattened (session, query) rows
    B, S, Lq = ins['q'].shape
    N, Ld = ins['d'].shape[2], ins['d'].shape[3]
    mt_cfg = dict(cfg, model='match_tensor', nhid_doc=cfg['nhid_document'])
    mt_sd = {}
    for k, v in sd.items():
        k2 = k.replace('embedder.word_embeddings', 'word_embeddings').replace('_encoder.encoder.', '_encoder.')
        mt_sd[k2] = v
    ref = ol.run_ranker(mt_cfg, mt_sd, ins['q'].reshape(B * S, Lq), ins['qlen'].reshape(-1), ins['d'].reshape(B * S, N, Ld),
                        ins['dlen'].reshape(B * S, N))['scores']
    assert _rel(s.cpu().numpy().reshape(B * S, N), ref) < TOL
