"""Ranking metrics: oracle vs the reference's own eval/ltorank.py (golden fixtures), CUDA vs oracle."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import rank_metrics_oracle as rmo  # noqa: E402

FIXTURES = ['rank_metrics_n10', 'rank_metrics_ties', 'rank_metrics_n100']


@pytest.mark.parametrize('name', FIXTURES)
def test_oracle_matches_reference_functions(name):
    _, ins, _, outs = ol.load_golden(name)
    mean, rows, pred = rmo.rank_metrics(ins['scores'], ins['labels'])
    # same ranking as the reference's np.argsort(-softmax) up to the (implementation-defined) order inside exact ties
    probs = outs['probs']
    assert np.array_equal(np.take_along_axis(probs, pred, 1), np.take_along_axis(probs, outs['predictions'], 1))
    if 'ties' not in name:
        assert np.array_equal(pred, outs['predictions'])
    ref = np.array([outs['map'], outs['mrr'], outs['p1'], outs['p3'], outs['p5']], dtype=np.float64)
    assert np.abs(mean - ref).max() < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('name', FIXTURES)
def test_cuda_matches_reference_golden(name):
    from context_attentive_ir_b200.metrics import rank_metrics, rank_metrics_device
    _, ins, _, outs = ol.load_golden(name)
    s = torch.from_numpy(ins['scores']).cuda()
    lab = torch.from_numpy(ins['labels']).cuda()
    m = rank_metrics(s, lab)
    ref = dict(zip(('map', 'mrr', 'prec@1', 'prec@3', 'prec@5'), (outs['map'], outs['mrr'], outs['p1'], outs['p3'], outs['p5'])))
    for k in ref:
        assert abs(m[k] - float(ref[k])) < 1e-12, (k, m[k], ref[k])
    _, rows = rank_metrics_device(s, lab)
    _, orows, _ = rmo.rank_metrics(ins['scores'], ins['labels'])
    assert np.abs(rows.cpu().numpy() - orows).max() < 1e-12


@pytest.mark.gpu
def test_cuda_vs_oracle_large_and_edge_cases():
    from context_attentive_ir_b200 import lib
    from context_attentive_ir_b200.metrics import rank_metrics_device
    rng = np.random.RandomState(5)
    for B, N, nrel in [(1280, 10, 1), (37, 500, 7), (3, 4096, 20), (5, 5, 5)]:
        scores = (rng.randn(B, N) * 4).astype(np.float32)
        scores[:, N // 2] = scores[:, 0]                       # exact ties
        labels = np.zeros((B, N), dtype=np.int64)
        for b in range(B):
            labels[b, rng.choice(N, size=min(nrel, N), replace=False)] = 1
        labels[0, 1] = 2                                       # non-binary label: precision counts it, MAP / MRR do not
        mean, rows = rank_metrics_device(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda())
        omean, orows, _ = rmo.rank_metrics(scores, labels)
        assert np.abs(rows.cpu().numpy() - orows).max() < 1e-12, (B, N)
        assert np.abs(mean.cpu().numpy() - omean).max() < 1e-12, (B, N)
    # no softmax: rank the given values as they are; a row without a relevant document is NaN, not a crash
    scores = rng.randn(4, 8).astype(np.float32)
    labels = np.zeros((4, 8), dtype=np.int64)
    labels[1:, 3] = 1
    _, rows = rank_metrics_device(torch.from_numpy(scores).cuda(), torch.from_numpy(labels).cuda(), apply_softmax=False)
    _, orows, _ = rmo.rank_metrics(scores, labels, apply_softmax=False)
    r = rows.cpu().numpy()
    assert np.isnan(r[0, 0]) and np.isnan(orows[0, 0])
    assert np.abs(r[1:] - orows[1:]).max() < 1e-12 and np.array_equal(r[0, 1:], orows[0, 1:])
    # fewer than 5 candidates: the reference asserts in precision_at_k(…, 5)
    with pytest.raises(lib.CairError):
        rank_metrics_device(torch.zeros(2, 4).cuda(), torch.zeros(2, 4, dtype=torch.int64).cuda())


# ---- batchify ----------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['batchify_small', 'batchify_cfg2'])
def test_batchify_oracle_matches_reference(name):
    cfg, ins, _, outs = ol.load_golden(name)
    q, ql, d, dl = rmo.batchify_flat(ins['q_tokens'], ins['q_offsets'], ins['d_tokens'], ins['d_offsets'], cfg['B'], cfg['N'])
    for a, k in ((q, 'q'), (ql, 'qlen'), (d, 'd'), (dl, 'dlen')):
        assert a.dtype == np.int64 and np.array_equal(a, outs[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['batchify_small', 'batchify_cfg2'])
def test_batchify_cuda_matches_reference(name):
    from context_attentive_ir_b200 import lib
    from context_attentive_ir_b200.inputters import batchify_ranker
    cfg, ins, _, outs = ol.load_golden(name)
    t = {k: torch.from_numpy(v).cuda() for k, v in ins.items()}
    got = batchify_ranker(t['q_tokens'], t['q_offsets'], t['d_tokens'], t['d_offsets'], cfg['B'], cfg['N'])
    for a, k in zip(got, ('q', 'qlen', 'd', 'dlen')):
        assert a.dtype == torch.int64 and np.array_equal(a.cpu().numpy(), outs[k]), k
    # force_pad: longer padded lengths keep the same content, PAD beyond
    Lq, Ld = outs['q'].shape[1] + 3, outs['d'].shape[2] + 5
    q2, ql2, d2, dl2 = batchify_ranker(t['q_tokens'], t['q_offsets'], t['d_tokens'], t['d_offsets'], cfg['B'], cfg['N'], Lq, Ld)
    oq, oql, od, odl = rmo.batchify_flat(ins['q_tokens'], ins['q_offsets'], ins['d_tokens'], ins['d_offsets'], cfg['B'], cfg['N'], Lq, Ld)
    assert np.array_equal(q2.cpu().numpy(), oq) and np.array_equal(d2.cpu().numpy(), od)
    assert np.array_equal(ql2.cpu().numpy(), oql) and np.array_equal(dl2.cpu().numpy(), odl)
    # an example longer than the padded length is reported (the reference's copy_ raises)
    with pytest.raises(lib.CairError):
        batchify_ranker(t['q_tokens'], t['q_offsets'], t['d_tokens'], t['d_offsets'], cfg['B'], cfg['N'], outs['q'].shape[1],
                        outs['d'].shape[2] - 1)


@pytest.mark.gpu
def test_batchify_sessions_cuda_vs_oracle():
    from context_attentive_ir_b200.inputters import batchify_sessions
    rng = np.random.RandomState(3)
    B, S, N = 3, 4, 5
    ql = rng.randint(1, 12, size=B * S)
    dl = rng.randint(1, 40, size=B * S * N)
    qt = rng.randint(4, 999, size=ql.sum()).astype(np.int32)
    dtok = rng.randint(4, 999, size=dl.sum()).astype(np.int32)
    qo = np.concatenate([[0], np.cumsum(ql)]).astype(np.int64)
    do = np.concatenate([[0], np.cumsum(dl)]).astype(np.int64)
    got = batchify_sessions(*[torch.from_numpy(x).cuda() for x in (qt, qo, dtok, do)], B, S, N)
    oq, oql, od, odl = rmo.batchify_flat(qt, qo, dtok, do, B * S, N)
    assert np.array_equal(got[0].cpu().numpy(), oq.reshape(B, S, -1)) and np.array_equal(got[1].cpu().numpy(), oql.reshape(B, S))
    assert np.array_equal(got[2].cpu().numpy(), od.reshape(B, S, N, -1)) and np.array_equal(got[3].cpu().numpy(), odl.reshape(B, S, N))


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['mt_map64', 'cars_map70'])
def test_metrics_of_b200_scores_equal_metrics_of_reference_scores(name):
    """BASELINE.json's metric is "...; MAP vs ref": the B200 model scores of a >= 64-query fixture through cair_rank_metrics
    (softmax + ranking + MAP / MRR / P@1,3,5 on the device) against the reference's own evaluation of ITS scores
    (fixture out/metric_*, computed by neuroir.eval.ltorank on np.argsort(-softmax(scores)) as main/ranker.py:257-264 does)."""
    import helpers
    from context_attentive_ir_b200.metrics import rank_metrics
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, 'cuda')
    with torch.no_grad():
        if cfg['model'] == 'cars':
            s = net.score(*helpers.to_dev(ins, 'cuda', ('q', 'qlen', 'd', 'dlen', 'label')))['scores']
        else:
            s = net(*helpers.to_dev(ins, 'cuda'))
    n = s.shape[-1]
    assert s.numel() // n >= 64
    lab = torch.from_numpy(np.asarray(ins['label']).reshape(-1, n).astype(np.int64)).cuda()
    m = rank_metrics(s.reshape(-1, n).contiguous(), lab)
    ref = dict(zip(('map', 'mrr', 'prec@1', 'prec@3', 'prec@5'),
                   (outs['metric_map'], outs['metric_mrr'], outs['metric_p1'], outs['metric_p3'], outs['metric_p5'])))
    for k in ref:
        assert abs(m[k] - float(ref[k])) < 1e-12, (k, m[k], float(ref[k]))


@pytest.mark.parametrize('name', ['mt_map64', 'cars_map70'])
def test_metrics_of_oracle_scores_equal_reference_metrics(name):
    cfg, ins, sd, outs = ol.load_golden(name)
    if cfg['model'] == 'cars':
        s = ol.run_cars(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'], ins['label'])['scores']
    else:
        s = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'])['scores']
    n = s.shape[-1]
    mean, _, _ = rmo.rank_metrics(s.reshape(-1, n), np.asarray(ins['label']).reshape(-1, n).astype(np.int64))
    ref = np.array([outs['metric_map'], outs['metric_mrr'], outs['metric_p1'], outs['metric_p3'], outs['metric_p5']], dtype=np.float64)
    assert np.abs(mean - ref).max() < 1e-12
