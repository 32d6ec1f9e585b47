#!/usr/bin/env python
"""TEST INFRASTRUCTURE - generates tests/golden/*.npz by running the UNMODIFIED reference.

Imports the reference's own modules from /root/reference (read-only, never copied),
builds each network on CPU with torch.manual_seed(1013) (the reference's default seed,
main/ranker.py:49), feeds the seeded synthetic batches of
context_attentive_ir_b200/synth.py and freezes inputs, state_dict and outputs.
/root/reference does not exist on the GPU box, hence the frozen fixtures.

Out-of-tree compat shims (library drift):
  C1  rankers/drmm.py:71-75 - numpy.apply_along_axis over numpy.histogram returns a
      ragged tuple and raises under numpy >= 1.24; the shim subclass restates forward
      and replaces only those lines by a per-row numpy.histogram(...)[0].
  C2  multitask/cars.py:298 - masked_fill_ with a uint8 mask raises under torch >= 2;
      Tensor.masked_fill_ is wrapped to cast uint8 -> bool in this process only.

Run:  python oracle/gen_golden.py            (writes tests/golden/)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

from context_attentive_ir_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _apply_shims():
    orig = torch.Tensor.masked_fill_

    def masked_fill_(self, mask, value):
        if mask.dtype == torch.uint8:
            mask = mask.bool()
        return orig(self, mask, value)

    torch.Tensor.masked_fill_ = masked_fill_


def _drmm_shim_class():
    import numpy
    import torch.nn.functional as f
    from neuroir.rankers.drmm import DRMM

    class DRMMShim(DRMM):
        def forward(self, batch_queries, query_len, batch_docs, doc_len):
            batch_size = batch_queries.shape[0]
            qlen = batch_queries.shape[1]
            num_docs, dlen = batch_docs.shape[1], batch_docs.shape[2]
            embedded_queries = self.word_embeddings(batch_queries.unsqueeze(2))
            embedded_queries = self.emb_drop(embedded_queries)
            term_weights = self.gating_network(embedded_queries).unsqueeze(1).expand(
                batch_size, num_docs, qlen)
            doc_rep = batch_docs.view(batch_size * num_docs, dlen)
            embedded_docs = self.word_embeddings(doc_rep.unsqueeze(2))
            embedded_docs = self.emb_drop(embedded_docs)
            embedded_queries = torch.stack([embedded_queries] * num_docs, dim=1)
            embedded_queries = embedded_queries.contiguous().view(batch_size * num_docs, qlen, -1)
            embedded_queries = torch.stack([embedded_queries] * dlen, dim=2)
            embedded_docs = torch.stack([embedded_docs] * qlen, dim=1)
            cos_sim = f.cosine_similarity(embedded_queries, embedded_docs, 3)
            cs = cos_sim.detach().cpu().numpy()
            self._last_cos = cs
            # --- the only replaced lines (drmm.py:71-75) ---
            hist = numpy.stack([numpy.stack([numpy.histogram(cs[i, j], bins=self.bins)[0]
                                             for j in range(cs.shape[1])])
                                for i in range(cs.shape[0])])
            self._last_hist = hist
            histogram_feats = torch.from_numpy(hist).float()
            # -----------------------------------------------
            ffnn_out = self.ffnn(histogram_feats).squeeze(2)
            ffnn_out = ffnn_out.view(batch_size, num_docs, -1).contiguous()
            weighted_ffnn_out = ffnn_out * term_weights
            score = self.output(torch.sum(weighted_ffnn_out, 2, keepdim=True)).squeeze(1)
            return score.view(batch_size, num_docs)

    return DRMMShim


def _ns(**kw):
    return argparse.Namespace(**kw)


def _save(name, cfg, batch, net, outputs):
    arrs = {}
    for k, v in batch.items():
        arrs['in/' + k] = v
    for k, v in net.state_dict().items():
        arrs['sd/' + k] = v.detach().cpu().numpy()
    for k, v in outputs.items():
        arrs['out/' + k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    meta = dict(cfg=cfg, torch=torch.__version__, numpy=np.__version__,
                threads=torch.get_num_threads(), seed=1013)
    arrs['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrs)
    print('%-28s %8.1f KB  %s' % (name, os.path.getsize(path) / 1024,
                                  {k: tuple(np.shape(v)) for k, v in outputs.items()}))


def _t(batch):
    return {k: torch.from_numpy(v) for k, v in batch.items()}


# ------------------------------------------------------------------ ESM
def gen_esm(name, seed, B, N, Lq, Ld, E, V, **kw):
    from neuroir.rankers.esm import ESM
    torch.manual_seed(1013)
    net = ESM(_ns(emsize=E, src_vocab_size=V)).eval()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    with torch.no_grad():
        s = net(t['q'], t['qlen'], t['d'], t['dlen'])
    _save(name, dict(model='esm', emsize=E, src_vocab_size=V), batch, net, dict(scores=s))


# ------------------------------------------------------------------ DSSM / CDSSM
def gen_dssm(name, seed, B, N, Lq, Ld, E, V, nhid, nout, conv=False, **kw):
    if conv:
        from neuroir.rankers.cdssm import CDSSM as Net
    else:
        from neuroir.rankers.dssm import DSSM as Net
    torch.manual_seed(1013)
    cfg = dict(model='cdssm' if conv else 'dssm', emsize=E, src_vocab_size=V, dropout_emb=0.2, nhid=nhid, nout=nout)
    net = Net(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    with torch.no_grad():
        s = net(t['q'], t['qlen'], t['d'], t['dlen'])
    _save(name, cfg, batch, net, dict(scores=s))


# ------------------------------------------------------------------ ARC-I / ARC-II
def gen_arc(name, seed, B, N, Lq, Ld, E, V, two, **arch):
    if two:
        from neuroir.rankers.arcii import ARCII as Net
    else:
        from neuroir.rankers.arci import ARCI as Net
    torch.manual_seed(1013)
    cfg = dict(model='arcii' if two else 'arci', emsize=E, src_vocab_size=V, dropout_emb=0.2, max_query_len=Lq,
               max_doc_len=Ld, **arch)
    net = Net(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, bos_eos=True)
    t = _t(batch)
    with torch.no_grad():
        s = net(t['q'], t['qlen'], t['d'], t['dlen'])
    _save(name, cfg, batch, net, dict(scores=s))


# ------------------------------------------------------------------ Match-Tensor
def _ref_metrics(scores, labels):
    """MAP / MRR / P@1,3,5 of model scores exactly as the reference's evaluation loops compute them (main/ranker.py:257-264,
    main/multitask.py:286-293): softmax over the candidates, np.argsort(-scores), neuroir.eval.ltorank."""
    import torch.nn.functional as f
    from neuroir.eval.ltorank import MAP, MRR, precision_at_k
    probs = f.softmax(scores.reshape(-1, scores.shape[-1]), dim=-1).numpy()
    lab = np.asarray(labels).reshape(-1, scores.shape[-1]).astype(np.int64)
    pred = np.argsort(-probs)
    return dict(metric_map=np.float64(MAP(pred, lab)), metric_mrr=np.float64(MRR(pred, lab)),
                metric_p1=np.float64(precision_at_k(pred, lab, 1)), metric_p3=np.float64(precision_at_k(pred, lab, 3)),
                metric_p5=np.float64(precision_at_k(pred, lab, 5)))


def gen_mt(name, seed, B, N, Lq, Ld, E, V, F, Hq, Hd, C, nf, mfs, rnn_type='LSTM', metrics=False, nlayers=1, **kw):
    from neuroir.rankers.mtensor import MatchTensor
    torch.manual_seed(1013)
    cfg = dict(model='match_tensor', emsize=E, src_vocab_size=V, dropout_emb=0.2, rnn_type=rnn_type,
               bidirection=True, nlayers=nlayers, dropout_rnn=0.2, featsize=F, nhid_query=Hq,
               nhid_doc=Hd, nchannels=C, nfilters=nf, match_filter_size=mfs)
    net = MatchTensor(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    with torch.no_grad():
        s = net(t['q'], t['qlen'], t['d'], t['dlen'])
        # intermediates (mtensor.py:77-94), for stage-wise parity
        xd = net.linear_projection(net.word_embeddings(t['d'].view(B * N, Ld).unsqueeze(2)))
        xq = net.linear_projection(net.word_embeddings(t['q'].unsqueeze(2)))
        _, hq = net.query_encoder(xq, t['qlen'])
        _, hd = net.document_encoder(xd, t['dlen'].reshape(-1))
    outs = dict(scores=s, proj_docs=xd, enc_queries=hq, enc_docs=hd)
    if metrics:
        outs = dict(scores=s, **_ref_metrics(s, batch['label']))
    _save(name, cfg, batch, net, outs)


def gen_mt_train(name, seed, B, N, Lq, Ld, E, V, F, Hq, Hd, C, nf, mfs, p_drop=0.0, drop_seed=0, steps=3, lr=0.05,
                 clip=5.0, **kw):
    """Training fixtures: one train-mode forward + loss.backward() of the unmodified reference
    MatchTensor under BCEWithLogitsLoss (models/ranker.py:62-63, 213-219), then `steps` updates in the order of
    Ranker.update (:192-230: forward, criterion, zero_grad, backward, clip_grad_norm, SGD step) on the same batch.
    emb_drop (mtensor.py:33, :84, :89) is replaced by a module that applies the keep-scale mask of oracle/dropout_oracle.py
    (the hash the CUDA kernels draw from): queries first, then documents - nn.Dropout's own Philox stream cannot be
    reproduced outside torch."""
    from neuroir.rankers.mtensor import MatchTensor
    from dropout_oracle import drop_scale
    torch.manual_seed(1013)
    cfg = dict(model='match_tensor', emsize=E, src_vocab_size=V, dropout_emb=p_drop, rnn_type='LSTM',
               bidirection=True, nlayers=1, dropout_rnn=0.2, featsize=F, nhid_query=Hq,
               nhid_doc=Hd, nchannels=C, nfilters=nf, match_filter_size=mfs)
    net = MatchTensor(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).train()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    mask = torch.from_numpy(drop_scale(drop_seed, (B * Lq + B * N * Ld) * E, p_drop))

    class MaskDrop(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.calls = 0

        def forward(self, x):
            lo = 0 if self.calls % 2 == 0 else B * Lq * E
            self.calls += 1
            return x * mask[lo:lo + x.numel()].view(x.shape)
    net.emb_drop = MaskDrop()
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    labels = t['label'].float()
    crit = torch.nn.BCEWithLogitsLoss()
    scores = net(t['q'], t['qlen'], t['d'], t['dlen'])
    loss = crit(scores, labels)
    loss.backward()
    outs = dict(scores=scores.detach(), loss=loss.detach())
    for k, p in net.named_parameters():
        outs['grad/' + k] = p.grad.detach().clone()
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr, momentum=0.0, weight_decay=0.0)
    losses = []
    for _ in range(steps):
        sc = net(t['q'], t['qlen'], t['d'], t['dlen'])
        ls = crit(sc, labels)
        opt.zero_grad()
        ls.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        losses.append(float(ls))
    outs['losses'] = np.asarray(losses, dtype=np.float64)
    for k, v in net.state_dict().items():
        outs['final/' + k] = v.detach().clone()
    cfg = dict(cfg, drop_seed=drop_seed, lr=lr, clip=clip, steps=steps)
    net.load_state_dict(init)
    _save(name, cfg, batch, net, outs)


# ------------------------------------------------------------------ DRMM
def gen_drmm(name, seed, B, N, Lq, Ld, E, V, **kw):
    DRMMShim = _drmm_shim_class()
    torch.manual_seed(1013)
    cfg = dict(model='drmm', emsize=E, src_vocab_size=V, dropout_emb=0.2, nbins=5)
    net = DRMMShim(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    with torch.no_grad():
        s = net(t['q'], t['qlen'], t['d'], t['dlen'])
    _save(name, cfg, batch, net, dict(scores=s, cos=net._last_cos, hist=net._last_hist))


def gen_drmm_train(name, seed, B, N, Lq, Ld, E, V, p_drop=0.0, drop_seed=0, steps=3, lr=0.05, clip=5.0, **kw):
    """DRMM under the statement order of Ranker.update (models/ranker.py:192-230, criterion BCEWithLogitsLoss :59-60): gradients of
    one train-mode forward / backward and a 3-step SGD curve; emb_drop replaced by the oracle mask (queries, then documents)."""
    from dropout_oracle import drop_scale
    DRMMShim = _drmm_shim_class()
    torch.manual_seed(1013)
    cfg = dict(model='drmm', emsize=E, src_vocab_size=V, dropout_emb=p_drop, nbins=5)
    net = DRMMShim(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).train()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    mask = torch.from_numpy(drop_scale(drop_seed, (B * Lq + B * N * Ld) * E, p_drop))

    class MaskDrop(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.calls = 0

        def forward(self, x):
            lo = 0 if self.calls % 2 == 0 else B * Lq * E
            self.calls += 1
            return x * mask[lo:lo + x.numel()].view(x.shape)
    net.emb_drop = MaskDrop()
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    labels = t['label'].float()
    crit = torch.nn.BCEWithLogitsLoss()
    scores = net(t['q'], t['qlen'], t['d'], t['dlen'])
    loss = crit(scores, labels)
    loss.backward()
    outs = dict(scores=scores.detach(), loss=loss.detach(), hist=net._last_hist)
    for k, p in net.named_parameters():
        outs['grad/' + k] = p.grad.detach().clone()
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr, momentum=0.0, weight_decay=0.0)
    losses = []
    for _ in range(steps):
        ls = crit(net(t['q'], t['qlen'], t['d'], t['dlen']), labels)
        opt.zero_grad()
        ls.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        losses.append(float(ls.detach()))
    outs['losses'] = np.asarray(losses, dtype=np.float64)
    for k, v in net.state_dict().items():
        outs['final/' + k] = v.detach().clone()
    cfg = dict(cfg, drop_seed=drop_seed, lr=lr, clip=clip, steps=steps)
    net.load_state_dict(init)
    _save(name, cfg, batch, net, outs)


def gen_esm_train(name, seed, B, N, Lq, Ld, E, V, steps=3, lr=10.0, clip=5.0, **kw):
    """ESM under the statement order of Ranker.update (models/ranker.py:192-230, criterion BCEWithLogitsLoss :59-60): the gradient
    of the embedding table from one train-mode forward / backward of the unmodified reference, and a 3-step SGD curve."""
    from neuroir.rankers.esm import ESM
    torch.manual_seed(1013)
    cfg = dict(model='esm', emsize=E, src_vocab_size=V)
    net = ESM(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).train()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    labels = t['label'].float()
    crit = torch.nn.BCEWithLogitsLoss()
    scores = net(t['q'], t['qlen'], t['d'], t['dlen'])
    loss = crit(scores, labels)
    loss.backward()
    outs = dict(scores=scores.detach(), loss=loss.detach())
    for k, p in net.named_parameters():
        outs['grad/' + k] = p.grad.detach().clone()
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr, momentum=0.0, weight_decay=0.0)
    losses = []
    for _ in range(steps):
        ls = crit(net(t['q'], t['qlen'], t['d'], t['dlen']), labels)
        opt.zero_grad()
        ls.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        losses.append(float(ls.detach()))
    outs['losses'] = np.asarray(losses, dtype=np.float64)
    for k, v in net.state_dict().items():
        outs['final/' + k] = v.detach().clone()
    cfg = dict(cfg, drop_seed=0, lr=lr, clip=clip, steps=steps)
    net.load_state_dict(init)
    _save(name, cfg, batch, net, outs)


def gen_dssm_train(name, seed, B, N, Lq, Ld, E, V, nhid, nout, p_drop=0.0, drop_seed=0, steps=3, lr=1.0, clip=5.0, conv=False, **kw):
    """DSSM under the statement order of Ranker.update (models/ranker.py:192-230): gradients of one train-mode forward / backward
    of the unmodified reference and a 3-step SGD curve; emb_drop replaced by the oracle mask (query token rows, then documents)."""
    from neuroir.rankers.dssm import DSSM
    from neuroir.rankers.cdssm import CDSSM
    from dropout_oracle import drop_scale
    torch.manual_seed(1013)
    cfg = dict(model='cdssm' if conv else 'dssm', emsize=E, src_vocab_size=V, dropout_emb=p_drop, nhid=nhid, nout=nout)
    net = (CDSSM if conv else DSSM)(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).train()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    mask = torch.from_numpy(drop_scale(drop_seed, (B * Lq + B * N * Ld) * E, p_drop))

    class MaskDrop(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.calls = 0

        def forward(self, x):
            lo = 0 if self.calls % 2 == 0 else B * Lq * E
            self.calls += 1
            return x * mask[lo:lo + x.numel()].view(x.shape)
    net.emb_drop = MaskDrop()
    init = {k: v.detach().clone() for k, v in net.state_dict().items()}
    labels = t['label'].float()
    crit = torch.nn.BCEWithLogitsLoss()
    scores = net(t['q'], t['qlen'], t['d'], t['dlen'])
    loss = crit(scores, labels)
    loss.backward()
    outs = dict(scores=scores.detach(), loss=loss.detach())
    for k, p in net.named_parameters():
        outs['grad/' + k] = p.grad.detach().clone()
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr, momentum=0.0, weight_decay=0.0)
    losses = []
    for _ in range(steps):
        ls = crit(net(t['q'], t['qlen'], t['d'], t['dlen']), labels)
        opt.zero_grad()
        ls.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        losses.append(float(ls.detach()))
    outs['losses'] = np.asarray(losses, dtype=np.float64)
    for k, v in net.state_dict().items():
        outs['final/' + k] = v.detach().clone()
    cfg = dict(cfg, drop_seed=drop_seed, lr=lr, clip=clip, steps=steps)
    net.load_state_dict(init)
    _save(name, cfg, batch, net, outs)


# ------------------------------------------------------------------ DUET
def gen_duet(name, seed, B, N, Lq, Ld, E, V, nf, pool=5, **kw):
    from neuroir.rankers.duet import DUET
    torch.manual_seed(1013)
    cfg = dict(model='duet', emsize=E, src_vocab_size=V, dropout_emb=0.2, dropout=0.2, use_word=True,
               nfilters=nf, local_filter_size=1, dist_filter_size=3, pool_size=pool,
               max_doc_len=Ld, max_query_len=Lq)
    net = DUET(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.ranker_batch(seed, B, N, Lq, Ld, V, **kw)
    t = _t(batch)
    with torch.no_grad():
        s = net(t['q'], t['qlen'], t['d'], t['dlen'])
        loc = net.local_model(t['q'], t['d'])
    _save(name, cfg, batch, net, dict(scores=s, local=loc))


# ------------------------------------------------------------------ CARS (ranking path)
def gen_cars(name, seed, B, S, N, Lq, Ld, E, V, Hq, Hd, Hs, max_clicks=1, metrics=False, zero_click=None, **kw):
    from neuroir.multitask.cars import CARS
    torch.manual_seed(1013)
    cfg = dict(model='cars', emsize=E, src_vocab_size=V, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2,
               rnn_type='LSTM', bidirection=True, nlayers=1, nhid_query=Hq, nhid_document=Hd,
               nhid_click=Hs, nhid_session_query=Hs, nhid_session_document=Hs, nhid_decoder=Hs,
               query_session_off=False, doc_session_off=False, dropout_rnn=0.2,
               attn_type='general', mlp_nhid=150, pool_type='attn', regularize_coeff=0.1,
               alpha=0.1, lambda1=0.01, lambda2=0.0001, turn_ranker_off=False,
               turn_recommender_off=False)
    net = CARS(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.session_batch(seed, B, S, N, Lq, Ld, V, max_clicks=max_clicks, **kw)
    if zero_click is not None:   # a query nobody clicked on: its click vector is built from masked attention only (cars.py:285-304)
        batch['label'][zero_click[0], zero_click[1], :] = 0.0
    t = _t(batch)
    # dictionaries of the decode loop (cars.py:774-783: tgt_dict[idx] -> word -> src_dict[word]); the target word i maps
    # to an arbitrary fixed source id
    rng = np.random.RandomState(seed)
    tgt2src = rng.randint(4, V, size=cfg['tgt_vocab_size']).astype(np.int64)
    tgt2src[:4] = np.arange(4)
    tgt_dict = ['w%d' % i for i in range(cfg['tgt_vocab_size'])]
    src_dict = {'w%d' % i: int(tgt2src[i]) for i in range(cfg['tgt_vocab_size'])}
    max_len = 6
    with torch.no_grad():
        pooled, enc_src, _ = net.encode(t['q'], t['qlen'])
        pooled_docs = net.encode_document(t['d'], t['dlen'])
        clicks = net.encode_clicks(pooled_docs, t['label'])
        scores, states, sess_attn = net.rank_document(pooled, t['d'], t['dlen'], t['label'])
        dec = net.decode(states=states, max_len=max_len, src_dict=src_dict, tgt_dict=tgt_dict, batch_size=B,
                         session_len=S - 1, use_cuda=False, encoded_source=enc_src, source_len=t['qlen'],
                         session_attns=sess_attn)
    batch = dict(batch, tgt2src=tgt2src)
    if metrics:
        _save(name, cfg, batch, net, dict(scores=scores, predictions=dec['predictions'], **_ref_metrics(scores, batch['label'])))
        return
    _save(name, cfg, batch, net, dict(scores=scores, pooled_queries=pooled, pooled_docs=pooled_docs,
                                      clicks=clicks, sess_q_attn=sess_attn[0], sess_d_attn=sess_attn[1],
                                      encoded_source=enc_src, dec_h0=states[0], dec_c0=states[1],
                                      predictions=dec['predictions']))


# ------------------------------------------------------------------ MNSRF / M-Match-Tensor ranking paths
def gen_session_ranker(name, which, seed, B, S, N, Lq, Ld, E, V, Hq, Hd, Hs, rnn_type='LSTM', **arch):
    """encode + rank_document of the unmodified multitask models, as Multitask.predict calls them
    (models/multitask.py:270-276)."""
    if which == 'mnsrf':
        from neuroir.multitask.mnsrf import MNSRF as Net
    else:
        from neuroir.multitask.mmtensor import M_MATCH_TENSOR as Net
    torch.manual_seed(1013)
    cfg = dict(model=which, emsize=E, src_vocab_size=V, tgt_vocab_size=50, dropout_emb=0.2, dropout=0.2, rnn_type=rnn_type,
               bidirection=True, nlayers=1, nhid_query=Hq, nhid_document=Hd, nhid_session=Hs, dropout_rnn=0.2,
               regularize_coeff=0.1, **arch)
    net = Net(_ns(**{k: v for k, v in cfg.items() if k != 'model'})).eval()
    batch = synth.session_batch(seed, B, S, N, Lq, Ld, V, max_clicks=1)
    if which == 'm_match_tensor':   # some exact matches for the extra channel
        rng = np.random.default_rng(seed + 1)
        d, q = batch['d'], batch['q']
        for b in range(B):
            for s_ in range(S):
                for n in range(N):
                    L = int(batch['dlen'][b, s_, n])
                    for j in rng.choice(L, size=max(1, L // 8), replace=False):
                        d[b, s_, n, j] = q[b, s_, rng.integers(0, int(batch['qlen'][b, s_]))]
    t = _t(batch)
    # dictionaries of the decode loop (mnsrf.py:291-294: tgt_dict[idx] -> word -> src_dict[word]), as in gen_cars
    rng = np.random.RandomState(seed)
    tgt2src = rng.randint(4, V, size=cfg['tgt_vocab_size']).astype(np.int64)
    tgt2src[:4] = np.arange(4)
    tgt_dict = ['w%d' % i for i in range(cfg['tgt_vocab_size'])]
    src_dict = {'w%d' % i: int(tgt2src[i]) for i in range(cfg['tgt_vocab_size'])}
    with torch.no_grad():
        memory_bank, session_bank, states = net.encode(t['q'], t['qlen'])
        scores = net.rank_document(t['q'], memory_bank, session_bank, t['d'], t['dlen'])
        dec = net.decode(states=states, max_len=6, src_dict=src_dict, tgt_dict=tgt_dict, batch_size=B, session_len=S - 1,
                         use_cuda=False)
    outs = dict(scores=scores, session_bank=session_bank, predictions=dec['predictions'])
    if which == 'mnsrf':
        outs['memory_bank'] = memory_bank
    _save(name, cfg, dict(batch, tgt2src=tgt2src), net, outs)


# ------------------------------------------------------------------ ranking metrics (eval/ltorank.py)
def gen_rank_metrics(name, seed, B, N, max_rel=1, ties=False):
    """scores -> f.softmax (models/ranker.py:258) -> np.argsort(-scores) -> MAP / MRR / precision_at_k exactly as
    main/ranker.py:257-264 does, with the reference's own neuroir.eval.ltorank functions."""
    import torch.nn.functional as f
    from neuroir.eval.ltorank import MAP, MRR, precision_at_k
    rng = np.random.RandomState(seed)
    scores = (rng.randn(B, N) * 3).astype(np.float32)
    cand = np.arange(N)
    if ties:
        # Exact ties.  numpy's default argsort is NOT stable (since 1.25 float keys go through the AVX-512 / AVX2
        # x86-simd-sort kernels even for a dozen elements), so the reference's order inside a tie is implementation-
        # defined; the tied documents are kept non-relevant here, which makes every metric independent of that order.
        scores[:, 1] = scores[:, 0]
        scores[:, -1] = scores[:, 2]
        cand = np.arange(3, N - 1)
    labels = np.zeros((B, N), dtype=np.int64)
    for b in range(B):
        k = rng.randint(1, max_rel + 1)
        labels[b, rng.choice(cand, size=k, replace=False)] = 1
    probs = f.softmax(torch.from_numpy(scores), dim=-1).numpy()
    pred = np.argsort(-probs)
    out = dict(map=np.float64(MAP(pred, labels)), mrr=np.float64(MRR(pred, labels)),
               p1=np.float64(precision_at_k(pred, labels, 1)), p3=np.float64(precision_at_k(pred, labels, 3)),
               p5=np.float64(precision_at_k(pred, labels, 5)), predictions=pred.astype(np.int64), probs=probs)
    os.makedirs(OUT, exist_ok=True)
    meta = dict(cfg=dict(model='rank_metrics', B=B, N=N), torch=torch.__version__, numpy=np.__version__)
    arrays = {'meta': np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), 'in/scores': scores, 'in/labels': labels}
    arrays.update({'out/' + k: np.asarray(v) for k, v in out.items()})
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrays)
    print('wrote', name, 'MAP %.6f MRR %.6f' % (out['map'], out['mrr']))


# ------------------------------------------------------------------ batchify (inputters/ranker/vector.py:39-90)
def gen_batchify(name, seed, B, N, max_q, max_d, V):
    """Runs the reference's own batchify on vectorised examples (the dicts vectorize() returns, vector.py:24-36) and
    freezes the ragged inputs (flat token arrays + offsets) next to its padded outputs."""
    from neuroir.inputters.ranker.vector import batchify
    rng = np.random.RandomState(seed)
    batch, q_tok, d_tok, q_off, d_off = [], [], [], [0], [0]
    for b in range(B):
        ql = rng.randint(2, max_q + 1)
        qw = rng.randint(4, V, size=ql)
        docs = []
        for n in range(N):
            dl = rng.randint(2, max_d + 1)
            docs.append(rng.randint(4, V, size=dl))
        q_tok.append(qw)
        q_off.append(q_off[-1] + ql)
        for dw in docs:
            d_tok.append(dw)
            d_off.append(d_off[-1] + len(dw))
        batch.append({'id': 'q%d' % b, 'query_words': torch.LongTensor(qw), 'doc_words': [torch.LongTensor(x) for x in docs],
                      'label': torch.LongTensor(rng.randint(0, 2, size=N)), 'num_candidates': N,
                      'max_doc_len': max(len(x) for x in docs), 'max_query_len': ql})
    out = batchify(batch)
    os.makedirs(OUT, exist_ok=True)
    meta = dict(cfg=dict(model='batchify', B=B, N=N), torch=torch.__version__, numpy=np.__version__)
    arrays = {'meta': np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
              'in/q_tokens': np.concatenate(q_tok).astype(np.int32), 'in/q_offsets': np.asarray(q_off, np.int64),
              'in/d_tokens': np.concatenate(d_tok).astype(np.int32), 'in/d_offsets': np.asarray(d_off, np.int64),
              'out/q': out['que_rep'].numpy(), 'out/qlen': out['que_len'].numpy(), 'out/d': out['doc_rep'].numpy(),
              'out/dlen': out['doc_len'].numpy()}
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **arrays)
    print('wrote', name, tuple(out['doc_rep'].shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    _apply_shims()
    torch.set_num_threads(1)  # deterministic MKLDNN summation order
    only = sys.argv[1:]  # optional: fixture-name prefixes to (re)generate
    if only:
        g = globals()
        for fn in ('gen_esm', 'gen_mt', 'gen_mt_train', 'gen_drmm_train', 'gen_esm_train', 'gen_dssm_train', 'gen_session_ranker', 'gen_drmm', 'gen_duet', 'gen_cars', 'gen_dssm', 'gen_arc', 'gen_rank_metrics', 'gen_batchify'):
            g[fn] = (lambda f: (lambda name, *a, **k: f(name, *a, **k) if any(name.startswith(o) for o in only) else None))(g[fn])
    # BASELINE configs[0]: the reference's own CPU-runnable case (vocab cut 10k -> 1k to keep the file small)
    gen_esm('esm_cfg1', 1235, B=8, N=5, Lq=10, Ld=50, E=64, V=1000)
    gen_esm('esm_e300', 11, B=3, N=4, Lq=20, Ld=200, E=300, V=300, bos_eos=True)
    # Match-Tensor: tiny, cfg2 architecture (H=128) at reduced batch/vocab, stock 30/140 sizes
    gen_mt('mt_tiny', 21, B=3, N=4, Lq=7, Ld=23, E=32, V=100, F=8, Hq=12, Hd=20, C=10, nf=4, mfs=6,
           overlap=0.2)
    gen_mt('mt_cfg2arch', 1236, B=2, N=3, Lq=20, Ld=200, E=300, V=400, F=40, Hq=128, Hd=128, C=50, nf=6,
           mfs=20, bos_eos=True, overlap=0.1)
    gen_mt('mt_stock', 23, B=2, N=2, Lq=20, Ld=200, E=300, V=400, F=40, Hq=30, Hd=140, C=50, nf=6,
           mfs=20, bos_eos=True)
    gen_mt('mt_fullpad', 24, B=2, N=2, Lq=12, Ld=40, E=48, V=200, F=16, Hq=32, Hd=32, C=18, nf=6,
           mfs=20, variable=False)
    gen_mt('mt_gru', 25, B=2, N=3, Lq=9, Ld=31, E=32, V=150, F=12, Hq=20, Hd=28, C=10, nf=6, mfs=8, rnn_type='GRU',
           bos_eos=True, overlap=0.15)
    # stacked encoders (--nlayers, encoders/rnn_encoder.py:45-53, :92-113): 2 LSTM layers at the cfg2 hidden sizes, 3 GRU layers
    gen_mt('mt_2layer', 35, B=2, N=3, Lq=12, Ld=50, E=48, V=300, F=40, Hq=128, Hd=128, C=50, nf=6, mfs=20, nlayers=2,
           bos_eos=True, overlap=0.1)
    gen_mt('mt_3layer_gru', 36, B=2, N=3, Lq=9, Ld=31, E=32, V=150, F=12, Hq=20, Hd=28, C=10, nf=6, mfs=8, rnn_type='GRU',
           nlayers=3, bos_eos=True, overlap=0.15)
    # >= 64 queries with the reference's own MAP / MRR / P@k of the reference scores (metric parity of model scores)
    gen_mt('mt_map64', 27, B=64, N=10, Lq=8, Ld=24, E=32, V=300, F=8, Hq=16, Hd=24, C=10, nf=6, mfs=8, bos_eos=True,
           overlap=0.1, metrics=True)
    # training step (first "next" row): gradients of every parameter + a 3-step SGD loss curve
    gen_mt_train('mt_train_tiny', 28, B=3, N=4, Lq=7, Ld=23, E=32, V=100, F=8, Hq=12, Hd=20, C=10, nf=4, mfs=6, overlap=0.2)
    gen_mt_train('mt_train_drop', 29, B=4, N=3, Lq=9, Ld=31, E=32, V=150, F=12, Hq=20, Hd=28, C=10, nf=6, mfs=8, p_drop=0.3,
                 drop_seed=20261017, bos_eos=True, overlap=0.15)
    gen_mt_train('mt_train_arch', 30, B=2, N=3, Lq=20, Ld=200, E=300, V=400, F=40, Hq=128, Hd=128, C=50, nf=6, mfs=20,
                 p_drop=0.2, drop_seed=77, bos_eos=True, overlap=0.1)
    gen_esm_train('esm_train_tiny', 37, B=3, N=4, Lq=7, Ld=23, E=32, V=100, overlap=0.2)
    gen_esm_train('esm_train_e300', 38, B=4, N=5, Lq=20, Ld=200, E=300, V=300, bos_eos=True, overlap=0.1)
    gen_dssm_train('dssm_train_tiny', 39, B=3, N=4, Lq=6, Ld=19, E=24, V=120, nhid=16, nout=8)
    gen_dssm_train('dssm_train_drop', 40, B=4, N=3, Lq=12, Ld=60, E=64, V=300, nhid=48, nout=32, p_drop=0.2, drop_seed=977,
                   bos_eos=True, overlap=0.1)
    gen_dssm_train('cdssm_train_tiny', 42, B=3, N=4, Lq=7, Ld=21, E=24, V=120, nhid=16, nout=8, conv=True)
    gen_dssm_train('cdssm_train_drop', 43, B=4, N=3, Lq=12, Ld=60, E=64, V=300, nhid=48, nout=32, p_drop=0.2, drop_seed=1977,
                   bos_eos=True, overlap=0.1, conv=True)
    gen_drmm_train('drmm_train_tiny', 33, B=3, N=4, Lq=8, Ld=30, E=32, V=200, disjoint=True)
    gen_drmm_train('drmm_train_drop', 34, B=4, N=3, Lq=12, Ld=60, E=64, V=300, p_drop=0.2, drop_seed=4242, disjoint=True)
    # DRMM: strict (disjoint ids) and overlapping (bin-edge cells excluded by the test using out/cos)
    gen_drmm('drmm_strict', 1237, B=3, N=4, Lq=20, Ld=200, E=300, V=400, disjoint=True)
    gen_drmm('drmm_overlap', 32, B=2, N=3, Lq=12, Ld=60, E=64, V=300, bos_eos=True, overlap=0.1)
    # DSSM / CDSSM (stock sizes nhid=300, nout=128 at E=300; tiny variants)
    gen_dssm('dssm_tiny', 61, B=2, N=3, Lq=6, Ld=19, E=24, V=120, nhid=16, nout=8)
    gen_dssm('dssm_e300', 1240, B=3, N=4, Lq=20, Ld=200, E=300, V=400, nhid=300, nout=128, bos_eos=True)
    gen_dssm('cdssm_tiny', 62, B=2, N=3, Lq=7, Ld=21, E=24, V=120, nhid=16, nout=8, conv=True)
    gen_dssm('cdssm_e300', 1241, B=2, N=3, Lq=20, Ld=200, E=300, V=400, nhid=96, nout=64, conv=True, bos_eos=True)
    # ARC-I / ARC-II (force_pad shapes; stock layer structure at reduced widths to keep the fixtures small)
    gen_arc('arci_tiny', 71, B=2, N=3, Lq=8, Ld=30, E=16, V=100, two=False, filters_1d=[12, 8], kernel_size_1d=[3, 3],
            maxpool_size_1d=[2, 2])
    gen_arc('arci_mid', 1242, B=2, N=3, Lq=20, Ld=200, E=300, V=300, two=False, filters_1d=[32, 16], kernel_size_1d=[3, 3],
            maxpool_size_1d=[2, 2])
    gen_arc('arcii_tiny', 72, B=2, N=3, Lq=8, Ld=24, E=16, V=100, two=True, filters_1d=8, kernel_size_1d=3,
            filters_2d=[12, 8], kernel_size_2d=[[3, 3], [3, 3]], maxpool_size_2d=[[2, 2], [2, 2]])
    gen_arc('arcii_mid', 1243, B=2, N=3, Lq=10, Ld=100, E=300, V=300, two=True, filters_1d=32, kernel_size_1d=3,
            filters_2d=[48, 32], kernel_size_2d=[[3, 3], [3, 3]], maxpool_size_2d=[[2, 2], [2, 2]])
    # DUET (force_pad shapes: every batch padded to max lens, lengths still variable)
    gen_duet('duet_tiny', 41, B=2, N=3, Lq=8, Ld=30, E=24, V=120, nf=16, overlap=0.2)
    gen_duet('duet_e300', 1239, B=2, N=3, Lq=20, Ld=200, E=300, V=400, nf=64, bos_eos=True, overlap=0.1)
    # batchify of ragged examples (inputters/ranker/vector.py:39-90)
    gen_batchify('batchify_small', 91, B=5, N=3, max_q=9, max_d=31, V=500)
    gen_batchify('batchify_cfg2', 92, B=16, N=10, max_q=20, max_d=200, V=30000)
    # ranking metrics of the evaluation loops (main/ranker.py:257-264)
    gen_rank_metrics('rank_metrics_n10', 81, B=64, N=10, max_rel=1)
    gen_rank_metrics('rank_metrics_ties', 82, B=32, N=12, max_rel=3, ties=True)
    gen_rank_metrics('rank_metrics_n100', 83, B=16, N=100, max_rel=5)
    # MNSRF / M-Match-Tensor ranking paths
    gen_session_ranker('mnsrf_tiny', 'mnsrf', 53, B=3, S=4, N=5, Lq=8, Ld=30, E=32, V=200, Hq=32, Hd=32, Hs=48)
    # (rnn_type GRU is not a fixture: the reference's own session loop raises there - `if init_states:` on a tensor,
    # encoders/rnn_encoder.py:77)
    gen_session_ranker('mnsrf_small', 'mnsrf', 54, B=2, S=3, N=4, Lq=6, Ld=17, E=24, V=150, Hq=16, Hd=20, Hs=24)
    gen_session_ranker('mnsrf_h256', 'mnsrf', 55, B=2, S=5, N=6, Lq=12, Ld=60, E=64, V=300, Hq=256, Hd=256, Hs=96)
    gen_session_ranker('mmt_tiny', 'm_match_tensor', 56, B=2, S=3, N=4, Lq=7, Ld=23, E=32, V=100, Hq=12, Hd=20, Hs=24,
                       featsize=8, nchannels=10, nfilters=4, match_filter_size=6)
    gen_session_ranker('mmt_arch', 'm_match_tensor', 57, B=2, S=3, N=3, Lq=12, Ld=60, E=64, V=300, Hq=128, Hd=128, Hs=64,
                       featsize=40, nchannels=50, nfilters=6, match_filter_size=20)
    # CARS ranking path
    gen_cars('cars_tiny', 51, B=2, S=3, N=4, Lq=6, Ld=17, E=24, V=150, Hq=16, Hd=16, Hs=24, max_clicks=1)
    gen_cars('cars_zeroclick', 61, B=3, S=4, N=5, Lq=6, Ld=17, E=24, V=150, Hq=16, Hd=16, Hs=24, max_clicks=3, zero_click=(1, 1))
    gen_cars('cars_clicks', 52, B=3, S=4, N=5, Lq=8, Ld=30, E=32, V=200, Hq=32, Hd=32, Hs=48, max_clicks=3)
    gen_cars('cars_mid', 1238, B=2, S=7, N=10, Lq=20, Ld=200, E=300, V=400, Hq=64, Hd=64, Hs=96,
             max_clicks=2)
    # stock encoder sizes (256 = 128 per direction: the 4-CTA cluster recurrence) at a batch where the click-mask width m = 3 > 1 matters at scale, with a few rows of one
    # click next to rows of three
    gen_cars('cars_map70', 1245, B=10, S=7, N=10, Lq=6, Ld=16, E=24, V=200, Hq=16, Hd=16, Hs=24, max_clicks=2, metrics=True)
    gen_cars('cars_h256', 1244, B=4, S=7, N=10, Lq=20, Ld=60, E=64, V=500, Hq=256, Hd=256, Hs=128, max_clicks=3)


if __name__ == '__main__':
    main()
