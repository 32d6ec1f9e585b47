"""Drop-in check against the UNMODIFIED reference wrappers (only where /root/reference exists, i.e. in the build
container; the GPU box has no reference tree): after integration.install() the reference's own Ranker / Multitask
wrappers construct the B200 networks, and a checkpoint written by the reference's stock network loads into ours through
the wrapper's own save / load code (.mdl format, neuroir/models/ranker.py:266-327)."""
import argparse
import os
import sys
import types

import pytest
import torch

import oracle_lib as ol

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'neuroir')), reason='reference tree not present')


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if 'prettytable' not in sys.modules:   # models/ranker.py imports it for a parameter table only
        stub = types.ModuleType('prettytable')
        stub.PrettyTable = type('PrettyTable', (), {'__init__': lambda self, *a, **k: None})
        sys.modules['prettytable'] = stub
    import neuroir.models.ranker as ref_ranker
    return ref_ranker


def _args(cfg):
    ns = argparse.Namespace(**{k: v for k, v in cfg.items() if k != 'model'})
    ns.model_type = cfg['model']
    return ns


class _Dict(dict):
    """Stand-in for the reference's vocabulary object: the wrapper only calls len() on it here."""


@pytest.mark.parametrize('name', ['mt_tiny', 'drmm_overlap', 'duet_tiny', 'esm_cfg1'])
def test_reference_wrapper_builds_b200_networks_and_round_trips_mdl(name, tmp_path):
    ref_ranker = _import_reference()
    import importlib
    stock = importlib.reload(ref_ranker)                      # pristine class bindings
    cfg, ins, sd, outs = ol.load_golden(name)
    vocab = _Dict((i, i) for i in range(cfg['src_vocab_size']))
    # 1. the reference's stock network, written to a .mdl by the reference's own save()
    ref_model = stock.Ranker(_args(cfg), vocab, {k: torch.from_numpy(v) for k, v in sd.items()})
    stock_cls = type(ref_model.network)
    path = str(tmp_path / 'model.mdl')
    ref_model.save(path)
    # 2. install() rebinds the class names the wrapper imported; the same wrapper code now builds our module
    import context_attentive_ir_b200.integration as integ
    from context_attentive_ir_b200 import rankers
    done = integ.install()
    assert any(x.startswith('neuroir.models.ranker.') for x in done)
    orig_load = torch.load
    torch.load = lambda *a, **k: orig_load(*a, **{**k, 'weights_only': False})   # SURVEY App. C5 (torch >= 2.6 default)
    try:
        mine = stock.Ranker.load(path)
    finally:
        torch.load = orig_load
    assert isinstance(mine.network, rankers._Ranker) and not isinstance(mine.network, stock_cls)
    got = mine.network.state_dict()
    assert sorted(got) == sorted(sd)
    for k in sd:
        assert torch.equal(got[k], torch.from_numpy(sd[k])), k
    # 3. no silent CPU path: the wrapper's predict() on CPU tensors fails loudly instead of falling back
    ex = {'que_rep': torch.from_numpy(ins['q']), 'que_len': torch.from_numpy(ins['qlen']),
          'doc_rep': torch.from_numpy(ins['d']), 'doc_len': torch.from_numpy(ins['dlen'])}
    with pytest.raises(RuntimeError, match='CUDA'):
        mine.predict(ex)
    importlib.reload(ref_ranker)                               # leave the reference module pristine for other tests


def test_reference_multitask_wrapper_builds_b200_cars_and_keeps_decoder_keys(tmp_path):
    """CARS under the reference's Multitask wrapper: the FULL reference state_dict (ranking + suggestion-decoder keys)
    loads through the wrapper's strict load_state_dict into real parameters of the B200 module and comes back unchanged
    from the next save(); predict() reaches the B200 module's encode / rank_document / decode (which refuse CPU tensors)."""
    _import_reference()
    import importlib
    import neuroir.models.multitask as ref_mt
    ref_mt = importlib.reload(ref_mt)
    from neuroir.multitask.cars import CARS as RefCARS
    cfg, ins, sd, outs = ol.load_golden('cars_tiny')
    args = _args(cfg)
    src = _Dict((i, i) for i in range(cfg['src_vocab_size']))
    tgt = _Dict((i, i) for i in range(cfg['tgt_vocab_size']))
    torch.manual_seed(1013)
    stock = ref_mt.Multitask(argparse.Namespace(**vars(args)), src, tgt)
    assert isinstance(stock.network, RefCARS)
    full = stock.network.state_dict()
    for k, v in sd.items():                                    # the fixture's ranking weights into the stock network
        if k in full:
            full[k] = torch.from_numpy(v)
    import context_attentive_ir_b200.integration as integ
    from context_attentive_ir_b200 import multitask
    assert 'neuroir.models.multitask.CARS' in integ.install()
    mine = ref_mt.Multitask(argparse.Namespace(**vars(args)), src, tgt, dict(full))
    assert isinstance(mine.network, multitask.CARS)
    got = mine.network.state_dict()
    assert sorted(got) == sorted(full)
    decoder_keys = [k for k in full if k.startswith(multitask.DECODER_PREFIXES)]
    assert decoder_keys and all(k in dict(mine.network.named_parameters()) for k in decoder_keys)
    for k in full:
        assert torch.equal(got[k], full[k]), k
    assert callable(mine.network.decode) and callable(mine.network.encode) and callable(mine.network.rank_document)
    importlib.reload(ref_mt)


@pytest.mark.parametrize('name, ref_cls', [('mnsrf_tiny', 'MNSRF'), ('mmt_tiny', 'M_MATCH_TENSOR')])
def test_reference_multitask_wrapper_builds_b200_mnsrf_and_mmt(name, ref_cls):
    """MNSRF / M_MATCH_TENSOR under the reference's Multitask wrapper (models/multitask.py:41-58): after install() the wrapper
    constructs the B200 mirrors, and the stock network's full state_dict loads strictly and comes back unchanged."""
    _import_reference()
    import importlib
    import neuroir.models.multitask as ref_mt
    ref_mt = importlib.reload(ref_mt)
    cfg, ins, sd, outs = ol.load_golden(name)
    args = _args(cfg)
    src = _Dict((i, i) for i in range(cfg['src_vocab_size']))
    tgt = _Dict((i, i) for i in range(cfg['tgt_vocab_size']))
    stock = ref_mt.Multitask(argparse.Namespace(**vars(args)), src, tgt, {k: torch.from_numpy(v) for k, v in sd.items()})
    assert type(stock.network).__name__ == ref_cls and type(stock.network).__module__.startswith('neuroir.')
    full = stock.network.state_dict()
    import context_attentive_ir_b200.integration as integ
    from context_attentive_ir_b200 import multitask
    assert 'neuroir.models.multitask.' + ref_cls in integ.install()
    mine = ref_mt.Multitask(argparse.Namespace(**vars(args)), src, tgt, dict(full))
    assert isinstance(mine.network, getattr(multitask, ref_cls))
    got = mine.network.state_dict()
    assert sorted(got) == sorted(full)
    for k in full:
        assert torch.equal(got[k], full[k]), k
    for fn in ('encode', 'rank_document', 'decode'):     # the three calls of Multitask.predict (models/multitask.py:270-292)
        assert callable(getattr(mine.network, fn))
    importlib.reload(ref_mt)


def test_stacked_encoders_through_the_reference_config_path():
    """`--nlayers` travels through the unmodified neuroir.config (add_model_args -> get_model_args) into Ranker.__init__.
    hyparam's arch dict overrides the CLI flag (neuroir/hyparam.py:88-100 fixes nlayers = 1), so a stacked
    request is injected where get_model_args leaves it: on the Namespace.  The B200 Match-Tensor builds 1..4 layers with the
    reference's state_dict keys (`rnns.<k>.*`; scores: tests/test_parity_gpu.py mt_2layer / mt_3layer_gru) and refuses
    more loudly instead of truncating."""
    ref_ranker = _import_reference()
    import importlib
    stock = importlib.reload(ref_ranker)
    import neuroir.config as config
    parser = argparse.ArgumentParser()
    parser.register('type', 'bool', lambda x: x.lower() in ('yes', 'true', 't', '1', 'y'))   # main/ranker.py:34
    parser.add_argument('--model_type', type=str, default='dssm')                             # main/ranker.py:40
    config.add_model_args(parser)
    args = parser.parse_args(['--model_type', 'match_tensor', '--nlayers', '2'])
    margs = config.get_model_args(args)
    assert margs.nlayers == 1          # hyparam.py wins over the flag ...
    ref_keys = set(stock.Ranker(margs, _Dict((i, i) for i in range(50))).network.state_dict())
    import context_attentive_ir_b200.integration as integ
    integ.install()
    vocab = _Dict((i, i) for i in range(50))
    stock.Ranker(margs, vocab)         # ... so the stock path builds
    margs.nlayers = 2
    importlib.reload(ref_ranker)
    two_ref = set(stock.Ranker(margs, vocab).network.state_dict())
    integ.install()
    two = stock.Ranker(margs, vocab).network
    assert type(two).__module__.startswith('context_attentive_ir_b200')
    assert set(two.state_dict()) == two_ref and two_ref > ref_keys
    assert 'document_encoder.rnns.1.weight_ih_l0_reverse' in two_ref
    margs.nlayers = 5
    with pytest.raises(NotImplementedError, match='at most 4'):
        stock.Ranker(margs, vocab)
    importlib.reload(ref_ranker)
