"""Parity tests proper: the CUDA path (through the nn.Module mirrors -> C ABI of libcair.so)
against (a) the frozen outputs of the unmodified reference (tests/golden) and (b) the CPU oracle
on fresh seeded inputs, plus size-independent properties at BASELINE.json's full shapes.
Tolerance: 1e-3 relative fp32 (BASELINE.json north_star), relative error floored at 1 % of the
batch score scale (oracle_lib.rel_err)."""
import numpy as np
import pytest
import torch

from context_attentive_ir_b200 import lib, synth

import helpers
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-3
DEV = 'cuda:0'


def _max_rel(a, ref):
    return float(ol.rel_err(a, ref).max())


def _run(name, **kw):
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, DEV)
    with torch.no_grad():
        s = net(*helpers.to_dev(ins, DEV), **kw)
    torch.cuda.synchronize()
    net.poll_error()
    return cfg, ins, sd, outs, net, s.cpu().numpy()


def test_native_library_is_the_one_running():
    n0 = lib.launch_count()
    _run('esm_cfg1')
    assert lib.launch_count() > n0


@pytest.mark.parametrize('name', ['esm_cfg1', 'esm_e300'])
def test_esm_golden(name):
    *_, outs, net, s = _run(name)
    assert _max_rel(s, outs['scores']) < 1e-4


def test_tcgen05_conventions_selftest():
    """D = A[shift:shift+128] @ B^T through tcgen05.mma with the library's operand layout / descriptors."""
    rng = np.random.default_rng(0)
    L = lib.load()
    for (N, K, shift, split) in [(64, 64, 0, 0), (96, 64, 3, 1), (96, 64, 6, 1), (256, 112, 0, 1), (16, 16, 1, 0),
                                 (32, 112, 0, 2), (64, 32, 2, 2)]:  # split=2: A operand from tensor memory
        A = rng.standard_normal((128 + shift, K)).astype(np.float32)
        B = rng.standard_normal((N, K)).astype(np.float32)
        a, b = torch.from_numpy(A).to(DEV), torch.from_numpy(B).to(DEV)
        d = torch.full((128, N), float('nan'), device=DEV)
        lib.check(L.cair_umma_selftest(a.data_ptr(), b.data_ptr(), d.data_ptr(), N, K, shift, split,
                                       torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        if split == 1:
            ref = A[shift:shift + 128].astype(np.float64) @ B.astype(np.float64).T
            tol = 1e-3  # ~2^-16 relative per product, |sum| ~ sqrt(K)
        else:
            bf = lambda x: torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy()
            ref = bf(A[shift:shift + 128]) @ bf(B).T
            tol = 1e-4
        err = np.abs(d.cpu().numpy() - ref).max()
        assert err < tol, (N, K, shift, split, err)


@pytest.mark.parametrize('impl', ['tc', 'tc_split', 'fp32'])
@pytest.mark.parametrize('name', ['mt_tiny', 'mt_fullpad', 'mt_stock', 'mt_cfg2arch', 'mt_gru', 'mt_2layer', 'mt_3layer_gru'])
def test_match_tensor_golden(name, impl):
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, DEV).set_interaction_impl(impl)
    B, Lq = ins['q'].shape
    _, N, Ld = ins['d'].shape
    enc_q = torch.full((B, Lq, cfg['nhid_query']), float('nan'), device=DEV)
    enc_d = torch.full((B * N, Ld, cfg['nhid_doc']), float('nan'), device=DEV)
    with torch.no_grad():
        net(*helpers.to_dev(ins, DEV))  # creates the handle
        lib.check(lib.load().cair_mt_set_debug(net._cair_handle, enc_q.data_ptr(), enc_d.data_ptr()))
        s = net(*helpers.to_dev(ins, DEV))
        lib.check(lib.load().cair_mt_set_debug(net._cair_handle, None, None))
    torch.cuda.synchronize()
    # stage-wise: encoder memory banks (zeros at pads) then scores
    assert np.abs(enc_q.cpu().numpy() - outs['enc_queries']).max() < 2e-4
    ed = enc_d.cpu().numpy()
    assert np.abs(ed - outs['enc_docs']).max() < 2e-4
    dl = ins['dlen'].reshape(-1)
    for i in range(len(dl)):
        assert not ed[i, dl[i]:].any()
    assert _max_rel(s.cpu().numpy(), outs['scores']) < TOL
    net.poll_error()


def test_match_tensor_tc_vs_fp32_kernels_cfg2():
    """The tensor-core interaction kernel against the fp32 CUDA-core kernel on the same device inputs."""
    torch.manual_seed(11)
    net = helpers.build_module(MT_CFG2).to(DEV)
    batch = synth.ranker_batch(77, 16, 10, 20, 200, MT_CFG2['src_vocab_size'], bos_eos=True, overlap=0.05)
    args = helpers.to_dev(batch, DEV)
    with torch.no_grad():
        a = net.set_interaction_impl('tc')(*args).cpu().numpy()
        b = net.set_interaction_impl('fp32')(*args).cpu().numpy()
    assert _max_rel(a, b) < 2e-4


@pytest.fixture(params=['tcgen05-cos', 'fp32-cos'])
def drmm_engine(request):
    """DRMM with the tcgen05 cosines (+ exact recompute at the bin edges) and with the fp32 kernels."""
    lib.check(lib.load().cair_set_drmm_impl(1 if request.param == 'tcgen05-cos' else 0))
    yield request.param
    lib.check(lib.load().cair_set_drmm_impl(1))


def test_drmm_golden_strict(drmm_engine):
    cfg, ins, sd, outs = ol.load_golden('drmm_strict')
    net = helpers.build_module(cfg, sd, DEV)
    B, Lq = ins['q'].shape
    _, N, Ld = ins['d'].shape
    hist = torch.full((B * N, Lq, 5), -1, dtype=torch.int32, device=DEV)
    with torch.no_grad():
        net(*helpers.to_dev(ins, DEV))
        lib.check(lib.load().cair_drmm_set_debug(net._cair_handle, hist.data_ptr()))
        s = net(*helpers.to_dev(ins, DEV))
        lib.check(lib.load().cair_drmm_set_debug(net._cair_handle, None))
    torch.cuda.synchronize()
    assert (hist.cpu().numpy() == outs['hist']).all()   # integer work: bit-exact
    assert _max_rel(s.cpu().numpy(), outs['scores']) < 1e-4


def _drmm_hist(net, args, rows, Lq):
    hist = torch.full((rows, Lq, 5), -1, dtype=torch.int32, device=DEV)
    with torch.no_grad():
        net(*args)
        lib.check(lib.load().cair_drmm_set_debug(net._cair_handle, hist.data_ptr()))
        s = net(*args)
        lib.check(lib.load().cair_drmm_set_debug(net._cair_handle, None))
    torch.cuda.synchronize()
    return hist.cpu().numpy(), s.cpu().numpy()


@pytest.mark.parametrize('case', ['overlap_pads', 'cfg3_full_lengths', 'odd_sizes'])
def test_drmm_tensor_core_histograms_identical_to_fp32_kernels(case):
    """The tcgen05 path recomputes every cosine near a bin edge with the fp32 kernels' arithmetic, so its histograms must be
    IDENTICAL to theirs - also where the reference itself is rounding-chaotic (exact matches, cos ~ 1) and on PAD rows."""
    E, V, B, N, Lq, Ld, kw = dict(overlap_pads=(300, 3000, 24, 10, 20, 200, dict(bos_eos=True, overlap=0.15)),
                                  cfg3_full_lengths=(300, 131072, 64, 10, 20, 200, dict(variable=False)),
                                  odd_sizes=(64, 500, 5, 3, 7, 131, dict(bos_eos=True, overlap=0.3)))[case]
    cfg = dict(model='drmm', emsize=E, src_vocab_size=V, dropout_emb=0.2, nbins=5)
    torch.manual_seed(11)
    net = helpers.build_module(cfg).to(DEV)
    with torch.no_grad():
        net.word_embeddings.word_lut.weight[7].zero_()          # an OOV-style all-zero row that is not PAD
    batch = synth.ranker_batch(21, B, N, Lq, Ld, V, **kw)
    args = helpers.to_dev(batch, DEV)
    L = lib.load()
    lib.check(L.cair_set_drmm_impl(1))
    h_tc, s_tc = _drmm_hist(net, args, B * N, Lq)
    lib.check(L.cair_set_drmm_impl(0))
    h_fp, s_fp = _drmm_hist(net, args, B * N, Lq)
    lib.check(L.cair_set_drmm_impl(1))
    assert (h_tc == h_fp).all(), int((h_tc != h_fp).sum())
    assert np.array_equal(s_tc, s_fp)


def test_drmm_golden_overlap_rows_away_from_bin_edges(drmm_engine):
    cfg, ins, sd, outs = ol.load_golden('drmm_overlap')
    net = helpers.build_module(cfg, sd, DEV)
    B, Lq = ins['q'].shape
    _, N, Ld = ins['d'].shape
    hist = torch.zeros((B * N, Lq, 5), dtype=torch.int32, device=DEV)
    with torch.no_grad():
        net(*helpers.to_dev(ins, DEV))
        lib.check(lib.load().cair_drmm_set_debug(net._cair_handle, hist.data_ptr()))
        net(*helpers.to_dev(ins, DEV))
        lib.check(lib.load().cair_drmm_set_debug(net._cair_handle, None))
    torch.cuda.synchronize()
    ref_cos = outs['cos']
    edges = np.array([-1.0, -0.5, 0.0, 0.5, 1.0], np.float32)
    near = (np.abs(ref_cos[..., None] - edges) < 4e-7).any(-1) & (ref_cos != 0.0)
    clean = ~near.any(-1)
    assert clean.sum() > 0
    h = hist.cpu().numpy()
    assert (h[clean] == outs['hist'][clean]).all()
    # every row still accounts for all Ld cells except the (few) dropped > 1.0 ones
    assert (h.sum(-1) <= Ld).all() and (h.sum(-1) >= Ld - near.sum(-1)).all()


@pytest.fixture(params=['tcgen05-gemm', 'fp32-gemm'])
def gemm_engine(request):
    """Runs a test with both engines of the generic GEMM (tcgen05 bf16x3 / fp32 CUDA cores)."""
    lib.check(lib.load().cair_set_gemm_impl(1 if request.param == 'tcgen05-gemm' else 0))
    yield request.param
    lib.check(lib.load().cair_set_gemm_impl(1))


@pytest.mark.parametrize('name', ['duet_tiny', 'duet_e300'])
def test_duet_golden(name, gemm_engine):
    *_, outs, net, s = _run(name)
    assert _max_rel(s, outs['scores']) < TOL


@pytest.mark.parametrize('name', ['dssm_tiny', 'dssm_e300', 'cdssm_tiny', 'cdssm_e300'])
def test_dssm_cdssm_golden(name):
    *_, outs, net, s = _run(name)
    assert _max_rel(s, outs['scores']) < TOL


@pytest.mark.parametrize('name', ['arci_tiny', 'arci_mid', 'arcii_tiny', 'arcii_mid'])
def test_arc_golden(name, gemm_engine):
    *_, outs, net, s = _run(name)
    assert _max_rel(s, outs['scores']) < TOL


def test_duet_rejects_unpadded_batches():
    cfg, ins, sd, outs = ol.load_golden('duet_tiny')
    net = helpers.build_module(cfg, sd, DEV)
    q, ql, d, dl = helpers.to_dev(ins, DEV)
    with pytest.raises(lib.CairError, match='BAD_SHAPE'):
        net(q[:, :-1], ql, d, dl)


@pytest.mark.parametrize('name', ['cars_tiny', 'cars_clicks', 'cars_zeroclick', 'cars_mid', 'cars_h256'])
def test_cars_golden(name, gemm_engine):
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, DEV)
    args = helpers.to_dev(ins, DEV, ('q', 'qlen', 'd', 'dlen', 'label'))
    with torch.no_grad():
        out = net.score(*args, want_stages=True, want_decoder_inputs=True)
        # the reference's two-call predict sequence gives the same scores
        pooled, _, _ = net.encode(args[0], args[1])
        s2, _, attns = net.rank_document(pooled, args[2], args[3], args[4])
    torch.cuda.synchronize()
    for k in ('pooled_queries', 'pooled_docs', 'clicks', 'sess_q_attn', 'sess_d_attn'):
        assert np.abs(out[k].cpu().numpy() - outs[k]).max() < 5e-4, k
    assert np.abs(out['enc_q'].cpu().numpy() - outs['encoded_source']).max() < 5e-4
    assert _max_rel(out['scores'].cpu().numpy(), outs['scores']) < TOL
    assert torch.equal(s2, out['scores'])


class _TgtDict(list):
    """tgt_dict[idx] -> word (neuroir Vocabulary indexing by int)."""


@pytest.mark.parametrize('name', ['cars_tiny', 'cars_clicks', 'cars_zeroclick', 'cars_mid', 'cars_h256'])
def test_cars_predict_sequence_matches_reference_decode(name):
    """The body of the reference's Multitask.predict (neuroir/models/multitask.py:264-292) on the B200 module: encode ->
    rank_document -> softmax -> decode(states=..., encoded_source=..., session_attns=...).  click scores within 1e-3 of
    the reference's, suggested tokens identical to its greedy decode (fixture out/predictions, generated by the unmodified
    CARS.decode)."""
    cfg, ins, sd, outs = ol.load_golden(name)
    net = helpers.build_module(cfg, sd, DEV)
    src, qlen, docs, dlen, lab = helpers.to_dev(ins, DEV, ('q', 'qlen', 'd', 'dlen', 'label'))
    nt = cfg['tgt_vocab_size']
    tgt_dict = _TgtDict('w%d' % i for i in range(nt))
    src_dict = {'w%d' % i: int(ins['tgt2src'][i]) for i in range(nt)}
    B, S = src.shape[0], src.shape[1]
    max_len = outs['predictions'].shape[-1]
    with torch.no_grad():
        pooled_rep, encoded_source, _ = net.encode(src, qlen)
        click_scores, states, session_attns = net.rank_document(pooled_rep, docs, dlen, lab)
        click_scores = torch.softmax(click_scores, dim=-1)
        dec = net.decode(states=states, max_len=max_len, src_dict=src_dict, tgt_dict=tgt_dict, batch_size=B,
                         session_len=S - 1, use_cuda=True, encoded_source=encoded_source, source_len=qlen,
                         session_attns=session_attns)
    ref_probs = torch.softmax(torch.from_numpy(outs['scores']), dim=-1).numpy()
    assert _max_rel(click_scores.cpu().numpy(), ref_probs) < TOL
    pred = dec['predictions'].cpu().numpy()
    assert pred.shape == outs['predictions'].shape
    assert np.array_equal(pred, outs['predictions']), (pred != outs['predictions']).mean()
    # and against the oracle's own decode from the device outputs
    fwd = {k: net._fwd[k].cpu().numpy() for k in ('enc_q', 'sess_h', 'sess_c', 'sess_q_attn', 'sess_d_attn')}
    assert np.array_equal(pred, ol.run_cars_decode(cfg, sd, fwd, ins['qlen'], max_len, ins['tgt2src']))


# ---- fresh seeded inputs against the oracle, sizes the oracle finishes in seconds ---------------
def _fresh(cfg, seed, B, N, Lq, Ld, **kw):
    torch.manual_seed(seed)
    net = helpers.build_module(cfg).to(DEV)
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, cfg['src_vocab_size'], **kw)
    with torch.no_grad():
        s = net(*helpers.to_dev(batch, DEV)).cpu().numpy()
    o = ol.run_ranker(cfg, helpers.state_dict_numpy(net), batch['q'], batch['qlen'], batch['d'], batch['dlen'])
    return net, batch, s, o['scores']


MT_CFG2 = dict(model='match_tensor', emsize=300, src_vocab_size=5000, dropout_emb=0.2, rnn_type='LSTM',
               bidirection=True, nlayers=1, dropout_rnn=0.2, featsize=40, nhid_query=128, nhid_doc=128,
               nchannels=50, nfilters=6, match_filter_size=20)


def test_match_tensor_cfg2_shapes_vs_oracle():
    net, batch, s, ref = _fresh(MT_CFG2, 101, B=4, N=10, Lq=20, Ld=200, bos_eos=True, overlap=0.05)
    assert _max_rel(s, ref) < TOL


@pytest.mark.parametrize('C', [16, 17, 20, 22, 33, 50, 64])
def test_match_tensor_channel_counts_vs_oracle(C):
    """K layouts of the tcgen05 interaction GEMM: C % 16 in {1,2} / {3,4} / {5..8} tail channels packed 4 / 2 / 1
    taps per 16-byte unit, and the plain padded layout (C % 16 == 0 or > 8)."""
    cfg = dict(MT_CFG2, src_vocab_size=300, nchannels=C)
    for impl in ('tc', 'tc_split'):
        torch.manual_seed(5)
        net = helpers.build_module(cfg).to(DEV).set_interaction_impl(impl)
        batch = synth.ranker_batch(9, 2, 3, 20, 200, 300, bos_eos=True, overlap=0.05)
        with torch.no_grad():
            s = net(*helpers.to_dev(batch, DEV)).cpu().numpy()
        ref = ol.run_ranker(cfg, helpers.state_dict_numpy(net), batch['q'], batch['qlen'], batch['d'], batch['dlen'])['scores']
        assert _max_rel(s, ref) < TOL, (C, impl)


def test_match_tensor_ragged_and_minimal_lengths():
    cfg = dict(MT_CFG2, src_vocab_size=500)
    torch.manual_seed(3)
    net = helpers.build_module(cfg).to(DEV)
    batch = synth.ranker_batch(7, 3, 5, 20, 200, 500)
    batch['qlen'][1] = 1          # shortest legal query
    batch['q'][1, 1:] = 0
    batch['dlen'][2, :] = 1       # one-token documents
    batch['d'][2, :, 1:] = 0
    with torch.no_grad():
        s = net(*helpers.to_dev(batch, DEV)).cpu().numpy()
    ref = ol.run_ranker(cfg, helpers.state_dict_numpy(net), batch['q'], batch['qlen'], batch['d'], batch['dlen'])['scores']
    assert _max_rel(s, ref) < TOL


def test_drmm_cfg3_shapes_vs_oracle():
    cfg = dict(model='drmm', emsize=300, src_vocab_size=5000, dropout_emb=0.2, nbins=5)
    net, batch, s, ref = _fresh(cfg, 102, B=6, N=10, Lq=20, Ld=200, disjoint=True)
    assert _max_rel(s, ref) < TOL


def test_esm_cfg1_vs_oracle():
    cfg = dict(model='esm', emsize=64, src_vocab_size=10000)
    net, batch, s, ref = _fresh(cfg, 103, B=8, N=5, Lq=10, Ld=50)
    assert _max_rel(s, ref) < 1e-4


def test_duet_cfg5_shapes_vs_oracle(gemm_engine):
    cfg = dict(model='duet', emsize=300, src_vocab_size=3000, dropout_emb=0.2, dropout=0.2, use_word=True,
               nfilters=300, local_filter_size=1, dist_filter_size=3, pool_size=5, max_doc_len=200, max_query_len=20)
    net, batch, s, ref = _fresh(cfg, 104, B=2, N=10, Lq=20, Ld=200, overlap=0.1)
    assert _max_rel(s, ref) < TOL


# ---- properties at full BASELINE shapes (no oracle needed) ------------------------------------------
def test_match_tensor_full_cfg2_properties():
    cfg = dict(MT_CFG2, src_vocab_size=131072)
    torch.manual_seed(5)
    net = helpers.build_module(cfg).to(DEV)
    B, N, Lq, Ld = 128, 10, 20, 200
    batch = synth.ranker_batch(1236, B, N, Lq, Ld, cfg['src_vocab_size'], bos_eos=True)
    q, ql, d, dl = helpers.to_dev(batch, DEV)
    with torch.no_grad():
        s = net(q, ql, d, dl)
        # (1) permutation equivariance over the candidate docs of each query
        perm = torch.randperm(N, device=DEV)
        sp = net(q, ql, d[:, perm].contiguous(), dl[:, perm].contiguous())
        # (2) batch-composition independence: a sub-batch scores the same
        sub = net(q[:5], ql[:5], d[:5], dl[:5])
        # (3) doc-parallel slices assemble to the full result
        half = B * N // 2 + 3
        a = net(q, ql, d, dl, pair_slice=(0, half))
        b = net(q, ql, d, dl, pair_slice=(half, B * N - half))
    torch.cuda.synchronize()
    net.poll_error()
    assert torch.isfinite(s).all()
    scale = s.abs().max().item()
    assert (sp - s[:, perm]).abs().max().item() <= 1e-5 * scale
    assert (sub - s[:5]).abs().max().item() <= 1e-5 * scale
    assert torch.equal((a + b), s)
    # (4) end-to-end host entry point returns the same scores
    host = [torch.from_numpy(batch[k]).pin_memory() for k in ('q', 'qlen', 'd', 'dlen')]
    out = torch.empty(B, N, dtype=torch.float32).pin_memory()
    for _ in range(4):  # 1st call eager, 2nd captures the CUDA graph, later calls replay it
        out.zero_()
        hs = net.forward_host(*host, out=out)
        assert torch.equal(hs, s.cpu())
    # (4b) pipelined submit/wait: same scores with 1, 2 and 3 batches in flight (the cross-batch software pipeline
    # scores part of batch k under the document encoder of batch k+1), for several pipeline splits, on fresh inputs
    outs3 = [torch.zeros(B, N, dtype=torch.float32).pin_memory() for _ in range(3)]
    batch2 = synth.ranker_batch(4321, B, N, 20, 200, cfg['src_vocab_size'], variable=True)
    host2 = [torch.from_numpy(batch2[k]).pin_memory() for k in ('q', 'qlen', 'd', 'dlen')]
    with torch.no_grad():
        s2 = net(*helpers.to_dev(batch2, DEV)).cpu()
    feeds, refs = [host, host2], [s.cpu(), s2]
    for frac in (0.33, 0.0, 0.6):
        lib.check(lib.load().cair_ranker_set_pipeline_split(net._cair_handle, frac))
        for depth in (1, 2, 3):
            n_it = 7
            for it in range(n_it + depth):
                if it >= depth:
                    j = it - depth
                    net.wait_host(j % 3)
                    assert torch.equal(outs3[j % 3], refs[j % 2]), (frac, depth, j)
                    outs3[j % 3].zero_()
                if it < n_it:
                    net.submit_host(*feeds[it % 2], out=outs3[it % 3], slot=it % 3)
    lib.check(lib.load().cair_ranker_set_pipeline_split(net._cair_handle, 0.5))
    # (5) the spot-checked oracle agrees on a few pairs of the big batch
    idx = [0, 57, 127]
    ref = ol.run_ranker(cfg, helpers.state_dict_numpy(net), batch['q'][idx], batch['qlen'][idx], batch['d'][idx],
                        batch['dlen'][idx])['scores']
    assert _max_rel(s[idx].cpu().numpy(), ref) < TOL


def _ranker_properties(cfg, seed, B, N, Lq, Ld, spot=(0, 1), tol_eq=1e-5, **kw):
    """Size-independent properties at a BASELINE.json shape: candidate-permutation equivariance, batch-composition
    independence, doc-parallel slices assembling to the full result, and the oracle on a few spot queries."""
    torch.manual_seed(seed)
    net = helpers.build_module(cfg).to(DEV)
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, cfg['src_vocab_size'], **kw)
    q, ql, d, dl = helpers.to_dev(batch, DEV)
    with torch.no_grad():
        s = net(q, ql, d, dl)
        perm = torch.randperm(N, device=DEV)
        sp = net(q, ql, d[:, perm].contiguous(), dl[:, perm].contiguous())
        sub = net(q[:3], ql[:3], d[:3], dl[:3])
        cut = B * N // 2 + 3
        a = net(q, ql, d, dl, pair_slice=(0, cut))
        b = net(q, ql, d, dl, pair_slice=(cut, B * N - cut))
    torch.cuda.synchronize()
    net.poll_error()
    assert torch.isfinite(s).all()
    scale = max(s.abs().max().item(), 1e-12)
    assert (sp - s[:, perm]).abs().max().item() <= tol_eq * scale
    assert (sub - s[:3]).abs().max().item() <= tol_eq * scale
    assert (a + b - s).abs().max().item() <= tol_eq * scale
    idx = list(spot)
    ref = ol.run_ranker(cfg, helpers.state_dict_numpy(net), batch['q'][idx], batch['qlen'][idx], batch['d'][idx],
                        batch['dlen'][idx])['scores']
    return s[idx].cpu().numpy(), ref


def test_drmm_full_cfg3_properties(drmm_engine):
    """BASELINE configs[2]: B=256, Lq=20, Ld=200, N=10, E=300.  Disjoint query / document ids keep the cosines away from
    the exact-match bin edge, so the histograms - and hence the scores - are exactly permutation- and
    batch-invariant."""
    cfg = dict(model='drmm', emsize=300, src_vocab_size=131072, dropout_emb=0.2, nbins=5)
    got, ref = _ranker_properties(cfg, 1237, 256, 10, 20, 200, spot=(0, 100, 255), tol_eq=0.0, disjoint=True)
    assert _max_rel(got, ref) < TOL


@pytest.mark.parametrize('N', [10, 50])
def test_duet_full_cfg5_properties(N):
    """BASELINE configs[4]: B=32, Lq=20, Ld=200, E=300, 300 filters, candidate sweep (N = 500 runs in tools/bench_models.py)."""
    cfg = dict(model='duet', emsize=300, src_vocab_size=131072, dropout_emb=0.2, dropout=0.2, use_word=True,
               nfilters=300, local_filter_size=1, dist_filter_size=3, pool_size=5, max_doc_len=200, max_query_len=20)
    got, ref = _ranker_properties(cfg, 1239, 32, N, 20, 200, spot=(0, 31), overlap=0.1)
    assert _max_rel(got, ref) < TOL


def test_cars_full_cfg4_properties():
    """BASELINE configs[3]: CARS B=32, S=7, N=10, Lq=20, Ld=200, E=300, H=256 - session sharding assembles to the full
    result and a sub-batch of sessions scores the same (the click-mask width is batch-global, so the sub-batch keeps the
    session that carries the batch maximum of clicks); oracle on that sub-batch."""
    cfg = dict(model='cars', emsize=300, src_vocab_size=131072, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2, rnn_type='LSTM',
               bidirection=True, nlayers=1, nhid_query=256, nhid_document=256, nhid_click=512, nhid_session_query=512,
               nhid_session_document=512, nhid_decoder=512, query_session_off=False, doc_session_off=False, dropout_rnn=0.2,
               attn_type='general', mlp_nhid=150, pool_type='attn', regularize_coeff=0.1, alpha=0.1, lambda1=0.01,
               lambda2=0.0001, turn_ranker_off=False, turn_recommender_off=False)
    torch.manual_seed(11)
    net = helpers.build_module(cfg).to(DEV)
    B, S, N = 32, 7, 10
    batch = synth.session_batch(1238, B, S, N, 20, 200, cfg['src_vocab_size'], max_clicks=1)
    args = helpers.to_dev(batch, DEV, ('q', 'qlen', 'd', 'dlen', 'label'))
    with torch.no_grad():
        s = net.score(*args)['scores']
        a = net.score(*args, session_slice=(0, 13))['scores']
        b = net.score(*args, session_slice=(13, B - 13))['scores']
        sub = net.score(*[t[:2].contiguous() for t in args])['scores']
    torch.cuda.synchronize()
    assert torch.isfinite(s).all()
    scale = s.abs().max().item()
    assert (a + b - s).abs().max().item() <= 1e-5 * scale
    assert (sub - s[:2]).abs().max().item() <= 1e-5 * scale     # max_clicks=1: every row has exactly one click
    ref = ol.run_cars(cfg, helpers.state_dict_numpy(net), batch['q'][:2], batch['qlen'][:2], batch['d'][:2], batch['dlen'][:2],
                      batch['label'][:2])
    assert _max_rel(s[:2].cpu().numpy(), ref['scores']) < TOL


@pytest.mark.parametrize('V,E,T', [(1000, 64, 400), (5000, 300, 12345), (300, 50, 77)])
def test_embed_gather_entry_point(V, E, T):
    """cair_embed_gather (modules/embeddings.py:243-252): out[t] = table[ids[t]], bit-exact, PAD row included."""
    rng = np.random.default_rng(V + E)
    table = rng.standard_normal((V, E)).astype(np.float32)
    table[0] = 0
    ids = rng.integers(0, V, T).astype(np.int64)
    ids[:5] = 0
    tt, ti = torch.from_numpy(table).to(DEV), torch.from_numpy(ids).to(DEV)
    out = torch.full((T, E), float('nan'), device=DEV)
    lib.check(lib.load().cair_embed_gather(tt.data_ptr(), V, E, ti.data_ptr(), T, out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), table[ids])


def test_bad_token_id_is_reported():
    cfg, ins, sd, outs = ol.load_golden('esm_cfg1')
    net = helpers.build_module(cfg, sd, DEV)
    q, ql, d, dl = helpers.to_dev(ins, DEV)
    q = q.clone()
    q[0, 0] = cfg['src_vocab_size'] + 5
    with torch.no_grad():
        net(q, ql, d, dl)
    with pytest.raises(lib.CairError, match='BAD_ARG'):
        net.poll_error()


def test_cars_cfg4_architecture_vs_oracle(gemm_engine):
    """Stock CARS sizes (E=300, H=256, session 512) at a reduced batch: exercises the tensor-core pre-gate GEMM."""
    cfg = dict(model='cars', emsize=300, src_vocab_size=2000, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2, rnn_type='LSTM',
               bidirection=True, nlayers=1, nhid_query=256, nhid_document=256, nhid_click=512, nhid_session_query=512,
               nhid_session_document=512, nhid_decoder=512, query_session_off=False, doc_session_off=False, dropout_rnn=0.2,
               attn_type='general', mlp_nhid=150, pool_type='attn', regularize_coeff=0.1, alpha=0.1, lambda1=0.01,
               lambda2=0.0001, turn_ranker_off=False, turn_recommender_off=False)
    torch.manual_seed(7)
    net = helpers.build_module(cfg).to(DEV)
    batch = synth.session_batch(1238, 2, 3, 10, 20, 60, cfg['src_vocab_size'], max_clicks=2)
    args = helpers.to_dev(batch, DEV, ('q', 'qlen', 'd', 'dlen', 'label'))
    with torch.no_grad():
        s = net.score(*args)['scores'].cpu().numpy()
    ref = ol.run_cars(cfg, helpers.state_dict_numpy(net), batch['q'], batch['qlen'], batch['d'], batch['dlen'], batch['label'])
    assert _max_rel(s, ref['scores']) < TOL


@pytest.fixture(params=['cluster', 'r1', 'fp32'])
def rnn_engine(request):
    """Recurrence engine under test: the cluster-split tcgen05 kernel (default), the round-1 tcgen05 kernel, fp32 kernels.
    Shapes an engine does not cover fall through to the next one (cair_set_rnn_impl)."""
    L = lib.load()
    lib.check(L.cair_set_rnn_impl({'fp32': 0, 'r1': 1, 'cluster': 3}[request.param]))
    yield request.param
    lib.check(L.cair_set_rnn_impl(2))


@pytest.mark.parametrize('rnn', ['LSTM', 'GRU'])
@pytest.mark.parametrize('n,L,inp,h', [
    (37, 23, 40, 64),      # cfg2 document encoder shape: 2-CTA cluster, fused x part
    (5, 9, 17, 24),        # one CTA, odd sizes
    (300, 12, 40, 64),     # several clusters
    (61, 17, 40, 70),      # reference stock Match-Tensor (nhid_doc 140): 3-CTA cluster, partly filled last block
    (33, 20, 40, 15),      # reference stock query encoder (nhid_query 30)
    (21, 15, 96, 96),      # pre-gate mode (in >= 48), 3-CTA cluster
    (37, 11, 300, 128),    # CARS document encoder: 4-CTA cluster, tensor-core pre-gate GEMM
    (150, 9, 300, 128),    # more than 128 sequences per direction: several clusters, 8 cells per thread
    (32, 7, 256, 512),     # h > 128: stepwise fp32 path (CARS session encoders)
])
def test_lstm_entry_point_vs_oracle(n, L, inp, h, rnn, rnn_engine):
    if rnn == 'GRU' and h > 128:
        pytest.skip('GRU beyond h = 128 is not used by any model')
    if rnn_engine != 'cluster' and (n, L) not in ((37, 23), (21, 15), (37, 11), (32, 7)):
        pytest.skip('fallback engines: one shape per code path is enough')
    G = 4 if rnn == 'LSTM' else 3
    rng = np.random.default_rng(9)
    x = rng.standard_normal((n, L, inp)).astype(np.float32)
    lens = rng.integers(1, L + 1, n).astype(np.int64)
    lens[0] = L

    def mk():
        k = 1.0 / np.sqrt(h)
        return dict(w_ih=rng.uniform(-k, k, (G * h, inp)).astype(np.float32),
                    w_hh=rng.uniform(-k, k, (G * h, h)).astype(np.float32),
                    b_ih=rng.uniform(-k, k, G * h).astype(np.float32), b_hh=rng.uniform(-k, k, G * h).astype(np.float32))
    fwd, rev = mk(), mk()
    rt = 0 if rnn == 'LSTM' else 1
    ref, hn, cn = ol.run_lstm(x, lens, fwd, rev, h, rt)
    import ctypes as C
    from context_attentive_ir_b200 import _abi
    t = {k: torch.from_numpy(v).to(DEV) for k, v in dict(x=x, lens=lens).items()}
    wf = {k: torch.from_numpy(v).to(DEV) for k, v in fwd.items()}
    wr = {k: torch.from_numpy(v).to(DEV) for k, v in rev.items()}

    def sd(w):
        return _abi.LstmDir(*[C.cast(w[k].data_ptr(), _abi.f32p) for k in ('w_ih', 'w_hh', 'b_ih', 'b_hh')])
    out = torch.full((n, L, 2 * h), float('nan'), device=DEV)
    hn_d = torch.zeros(2, n, h, device=DEV)
    cn_d = torch.zeros(2, n, h, device=DEV)
    f, r = sd(wf), sd(wr)
    lib.check(lib.load().cair_rnn_forward(rt, t['x'].data_ptr(), t['lens'].data_ptr(), n, L, inp, h, C.byref(f), C.byref(r),
                                          out.data_ptr(), hn_d.data_ptr(), cn_d.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-4
    assert np.abs(hn_d.cpu().numpy() - hn).max() < 1e-4
    if rnn == 'LSTM':
        assert np.abs(cn_d.cpu().numpy() - cn).max() < 1e-4


def test_rnn_unidirectional_and_single_sequence():
    """rev = NULL (one direction) and n = 1: the smallest launch of the cluster-split kernel."""
    import ctypes as C
    from context_attentive_ir_b200 import _abi
    rng = np.random.default_rng(4)
    for n, L, inp, h in ((1, 5, 40, 64), (9, 6, 300, 128)):
        x = rng.standard_normal((n, L, inp)).astype(np.float32)
        lens = np.full(n, L, np.int64)
        k = 1.0 / np.sqrt(h)
        fwd = dict(w_ih=rng.uniform(-k, k, (4 * h, inp)).astype(np.float32), w_hh=rng.uniform(-k, k, (4 * h, h)).astype(np.float32),
                   b_ih=rng.uniform(-k, k, 4 * h).astype(np.float32), b_hh=rng.uniform(-k, k, 4 * h).astype(np.float32))
        ref, hn, cn = ol.run_lstm(x, lens, fwd, None, h)
        w = {kk: torch.from_numpy(v).to(DEV) for kk, v in fwd.items()}
        f = _abi.LstmDir(*[C.cast(w[kk].data_ptr(), _abi.f32p) for kk in ('w_ih', 'w_hh', 'b_ih', 'b_hh')])
        xd, ld = torch.from_numpy(x).to(DEV), torch.from_numpy(lens).to(DEV)
        out = torch.full((n, L, h), float('nan'), device=DEV)
        lib.check(lib.load().cair_rnn_forward(0, xd.data_ptr(), ld.data_ptr(), n, L, inp, h, C.byref(f), None, out.data_ptr(),
                                              None, None, torch.cuda.current_stream().cuda_stream))
        assert np.abs(out.cpu().numpy() - ref).max() < 1e-4
