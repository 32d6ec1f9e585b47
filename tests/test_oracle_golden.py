"""Pins the CPU oracle (oracle/cair_oracle.c) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/gen_golden.py importing /root/reference)."""
import numpy as np
import pytest

import oracle_lib as ol

TOL = 1e-3  # BASELINE.json north_star: within 1e-3 relative fp32


def _max_rel(a, ref):
    return float(ol.rel_err(a, ref).max())


@pytest.mark.parametrize('name', ['esm_cfg1', 'esm_e300'])
def test_esm(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'])
    assert _max_rel(o['scores'], outs['scores']) < 1e-5


@pytest.mark.parametrize('name', ['mt_tiny', 'mt_fullpad', 'mt_stock', 'mt_cfg2arch', 'mt_gru', 'mt_2layer', 'mt_3layer_gru'])
def test_match_tensor(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'], want=('enc_queries', 'enc_docs'))
    assert np.abs(o['enc_queries'] - outs['enc_queries']).max() < 1e-5
    assert np.abs(o['enc_docs'] - outs['enc_docs']).max() < 1e-5
    # pad positions of the memory bank are exactly zero (rnn_encoder.py:135-139)
    dl = ins['dlen'].reshape(-1)
    for s in range(len(dl)):
        assert not o['enc_docs'][s, dl[s]:].any()
    assert _max_rel(o['scores'], outs['scores']) < 1e-4


def test_drmm_strict():
    cfg, ins, sd, outs = ol.load_golden('drmm_strict')
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'], want=('hist', 'cos'))
    assert np.abs(o['cos'] - outs['cos']).max() < 1e-6
    assert (o['hist'] == outs['hist']).all()
    assert _max_rel(o['scores'], outs['scores']) < 1e-5


def test_drmm_overlap_excluding_bin_edge_cells():
    """Exact-match cosines land within a few ulp of the bin edge 1.0 and are rounding-chaotic in the
    reference itself: compare histograms only over cells away from the edges."""
    cfg, ins, sd, outs = ol.load_golden('drmm_overlap')
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'], want=('hist', 'cos'))
    ref_cos = outs['cos']
    assert np.abs(o['cos'] - ref_cos).max() < 1e-6
    edges = np.array([-1.0, -0.5, 0.0, 0.5, 1.0], np.float32)
    near = (np.abs(ref_cos[..., None] - edges) < 4e-7).any(-1) & (ref_cos != 0.0)
    n_excluded = int(near.sum())
    assert 0 < n_excluded < ref_cos.size // 10

    def hist_of(cos, keep):
        h = np.zeros(cos.shape[:2] + (5,), np.int64)
        for p in range(cos.shape[0]):
            for i in range(cos.shape[1]):
                h[p, i] = np.histogram(cos[p, i][keep[p, i]], bins=[-1.0, -0.5, 0, 0.5, 1.0, 1.0])[0]
        return h
    assert (hist_of(o['cos'], ~near) == hist_of(ref_cos, ~near)).all()
    # rows without any near-edge cell must match the reference histogram exactly
    clean = ~near.any(-1)
    assert (o['hist'][clean] == outs['hist'][clean]).all()


@pytest.mark.parametrize('name', ['duet_tiny', 'duet_e300'])
def test_duet(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'], want=('local',))
    assert np.abs(o['local'] - outs['local']).max() < 1e-5
    assert _max_rel(o['scores'], outs['scores']) < 1e-4


def test_duet_rejects_unpadded_shapes():
    cfg, ins, sd, outs = ol.load_golden('duet_tiny')
    with pytest.raises(RuntimeError, match='BAD_SHAPE'):
        ol.run_ranker(cfg, sd, ins['q'][:, :-1], ins['qlen'], ins['d'], ins['dlen'])


@pytest.mark.parametrize('name', ['cars_tiny', 'cars_clicks', 'cars_zeroclick', 'cars_mid', 'cars_h256'])
def test_cars(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    o = ol.run_cars(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'], ins['label'])
    for k in ('pooled_queries', 'pooled_docs', 'clicks', 'sess_q_attn', 'sess_d_attn'):
        assert np.abs(o[k] - outs[k]).max() < 2e-5, k
    assert _max_rel(o['scores'], outs['scores']) < 1e-4
    # decoder side: memory banks, and the greedy decode (cars.py:706-791) token for token
    assert np.abs(o['enc_q'] - outs['encoded_source']).max() < 2e-5
    pred = ol.run_cars_decode(cfg, sd, o, ins['qlen'], outs['predictions'].shape[-1], ins['tgt2src'])
    assert np.array_equal(pred, outs['predictions'])


def test_cars_click_mask_width_matters_in_the_h256_fixture():
    """cars_h256 mixes rows of one and of three clicks, so the batch-global mask width m = 3 leaks unclicked documents into
    the one-click rows: scoring a one-click session alone (m = 1) must give different later-query scores."""
    cfg, ins, sd, outs = ol.load_golden('cars_h256')
    clicks = ins['label'].sum(-1)
    assert clicks.max() == 3 and clicks.min() == 1
    b = int(np.argmin(clicks.max(axis=1)))
    if clicks[b].max() < 3:
        alone = ol.run_cars(cfg, sd, ins['q'][b:b + 1], ins['qlen'][b:b + 1], ins['d'][b:b + 1], ins['dlen'][b:b + 1], ins['label'][b:b + 1])
        assert np.abs(alone['scores'][0, 1:] - outs['scores'][b, 1:]).max() > 1e-6


@pytest.mark.parametrize('rnn', [0, 1])
def test_oracle_rnn_entry_matches_torch(rnn):
    """cair_oracle_rnn (LSTM and GRU) against torch.nn.LSTM / GRU on packed sequences - the operator the reference calls
    (encoders/rnn_encoder.py:45-53, 72-74, 100-113); torch is the reference's arithmetic dependency."""
    import torch
    rng = np.random.default_rng(3)
    n, L, inp, h, G = 5, 7, 6, 9, (4, 3)[rnn]
    x = rng.standard_normal((n, L, inp)).astype(np.float32)
    lens = np.array([7, 3, 5, 1, 6], np.int64)
    mk = lambda: dict(w_ih=(rng.standard_normal((G * h, inp)) * .3).astype(np.float32), w_hh=(rng.standard_normal((G * h, h)) * .3).astype(np.float32),
                      b_ih=(rng.standard_normal(G * h) * .1).astype(np.float32), b_hh=(rng.standard_normal(G * h) * .1).astype(np.float32))
    f, r = mk(), mk()
    out, hn, cn = ol.run_lstm(x, lens, f, r, h, rnn)
    m = (torch.nn.LSTM, torch.nn.GRU)[rnn](inp, h, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for sfx, w in (('', f), ('_reverse', r)):
            for k in ('w_ih', 'w_hh', 'b_ih', 'b_hh'):
                getattr(m, k.replace('w_', 'weight_').replace('b_', 'bias_') + '_l0' + sfx).copy_(torch.from_numpy(w[k]))
        pk = torch.nn.utils.rnn.pack_padded_sequence(torch.from_numpy(x), lens, batch_first=True, enforce_sorted=False)
        o, st = m(pk)
        o, _ = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=L)
    assert np.abs(o.numpy() - out).max() < 1e-5
    assert np.abs((st[0] if rnn == 0 else st).numpy() - hn).max() < 1e-5


def test_bad_token_id_is_an_error():
    cfg, ins, sd, outs = ol.load_golden('esm_cfg1')
    q = ins['q'].copy()
    q[0, 0] = cfg['src_vocab_size']
    with pytest.raises(RuntimeError, match='BAD_ARG'):
        ol.run_ranker(cfg, sd, q, ins['qlen'], ins['d'], ins['dlen'])


@pytest.mark.parametrize('name', ['dssm_tiny', 'dssm_e300', 'cdssm_tiny', 'cdssm_e300'])
def test_dssm_cdssm(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'])
    assert _max_rel(o['scores'], outs['scores']) < 1e-4


def test_cdssm_rejects_too_short_sequences():
    cfg, ins, sd, outs = ol.load_golden('cdssm_tiny')
    with pytest.raises(RuntimeError, match='BAD_SHAPE'):
        ol.run_ranker(cfg, sd, ins['q'][:, :4], ins['qlen'], ins['d'], ins['dlen'])


@pytest.mark.parametrize('name', ['arci_tiny', 'arci_mid', 'arcii_tiny', 'arcii_mid'])
def test_arc(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'])
    assert _max_rel(o['scores'], outs['scores']) < 1e-4
