"""Shared test helpers: build the B200 modules from a golden fixture / config dict."""
import argparse

import numpy as np
import torch

import context_attentive_ir_b200 as cair

CLASSES = {'arci': cair.ARCI, 'arcii': cair.ARCII, 'dssm': cair.DSSM, 'cdssm': cair.CDSSM, 'esm': cair.ESM, 'match_tensor': cair.MatchTensor, 'drmm': cair.DRMM, 'duet': cair.DUET,
           'cars': cair.CARS, 'mnsrf': cair.MNSRF, 'm_match_tensor': cair.M_MATCH_TENSOR}


def namespace(cfg):
    return argparse.Namespace(**{k: v for k, v in cfg.items() if k != 'model'})


def build_module(cfg, sd=None, device=None):
    net = CLASSES[cfg['model']](namespace(cfg)).eval()
    if sd is not None:
        missing, unexpected = net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()},
                                                   strict=False)
        assert not missing, missing
        assert not unexpected, unexpected
    if device is not None:
        net = net.to(device)
    return net


def state_dict_numpy(net):
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in net.state_dict().items()}


def to_dev(batch, device, keys=('q', 'qlen', 'd', 'dlen')):
    return [torch.from_numpy(np.ascontiguousarray(batch[k])).to(device) for k in keys]
