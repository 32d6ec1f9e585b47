"""CPU-side checks of the drop-in boundary: libcair.so loads and exports every symbol that
include/cair.h declares, the ctypes mirror matches the header, the nn.Module mirrors carry the
reference's state_dict keys, and nothing silently falls back to a CPU path."""
import os
import re

import numpy as np
import pytest
import torch

import context_attentive_ir_b200 as cair
from context_attentive_ir_b200 import lib

import helpers
import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'cair.h')).read()
    return sorted(set(re.findall(r'CAIR_API\s+[\w\s\*]+?\b(cair_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 25
    L = lib.load()
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(lib.PROTOTYPES) == syms
    assert L.cair_version() == 100


def test_no_compute_without_gpu_fails_loudly():
    cfg, ins, sd, _ = ol.load_golden('esm_cfg1')
    net = helpers.build_module(cfg, sd)
    q, ql, d, dl = [torch.from_numpy(ins[k]) for k in ('q', 'qlen', 'd', 'dlen')]
    with pytest.raises(RuntimeError, match='CUDA'):
        net(q, ql, d, dl)


@pytest.mark.parametrize('name', ['esm_cfg1', 'mt_cfg2arch', 'mt_stock', 'drmm_strict', 'duet_e300', 'cars_mid', 'dssm_e300',
                                  'cdssm_tiny', 'arci_mid', 'arcii_mid'])
def test_state_dict_keys_and_shapes_match_reference(name):
    cfg, _, sd, _ = ol.load_golden(name)
    net = helpers.build_module(cfg)
    mine = net.state_dict()
    if cfg['model'] == 'cars':  # the suggestion-decoder keys are real parameters too (cars.py:605-657)
        assert any(k.startswith('decoder.decoder.rnn.') for k in sd) and 'token_prob_predictor2.weight' in sd
    assert sorted(mine) == sorted(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    # PAD row of a fresh table is zero (nn.Embedding padding_idx, modules/embeddings.py:165)
    tkey = [k for k in mine if k.endswith('emb_luts.0.weight')][0]
    assert not mine[tkey][0].any()


def test_weight_struct_layout_matches_oracle_view():
    # the same ctypes structs drive the C oracle (host pointers): a wrong field order would break test_oracle_golden
    cfg, ins, sd, outs = ol.load_golden('mt_tiny')
    o = ol.run_ranker(cfg, sd, ins['q'], ins['qlen'], ins['d'], ins['dlen'])
    assert np.abs(o['scores'] - outs['scores']).max() < 1e-5


def test_unsupported_configs_raise():
    cfg, _, _, _ = ol.load_golden('mt_tiny')
    bad = dict(cfg, nlayers=5)     # stacked encoders: 1..4 layers
    with pytest.raises(NotImplementedError):
        helpers.build_module(bad)
    cfgd, _, _, _ = ol.load_golden('duet_tiny')
    with pytest.raises(TypeError):
        helpers.build_module(dict(cfgd, use_word=False))


def test_pair_slices_cover_everything():
    from context_attentive_ir_b200.parallel import pair_slice
    for total in (1, 7, 10, 1280, 1283):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                b, c = pair_slice(r, world, total)
                got += list(range(b, b + c))
            assert got == list(range(total))
