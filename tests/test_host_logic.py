"""Host-side logic of the nn.Module mirrors that needs no GPU: ownership of the native handle (copies, pickles and
DataParallel replicas must not share or double-free it), weight-change detection, argument validation of the host entry
points, pair / session partitioning."""
import copy
import ctypes as C
import pickle

import pytest
import torch

import context_attentive_ir_b200 as cair
from context_attentive_ir_b200 import parallel, rankers

import helpers
import oracle_lib as ol


class _StubLib:
    def __init__(self):
        self.destroyed = []

    def cair_destroy(self, h):
        self.destroyed.append(h.value if hasattr(h, 'value') else h)
        return 0


@pytest.fixture
def stub(monkeypatch):
    s = _StubLib()
    monkeypatch.setattr(rankers.lib, 'load', lambda: s)
    return s


def _net():
    cfg, _, sd, _ = ol.load_golden('mt_tiny')
    return helpers.build_module(cfg, sd)


def test_copies_do_not_share_the_native_handle(stub):
    net = _net()
    net.__dict__['_cair_handle'] = C.c_void_p(0x1234)
    net.__dict__['_cair_key'] = ('k',)
    for clone in (copy.copy(net), copy.deepcopy(net), pickle.loads(pickle.dumps(net)), net._replicate_for_data_parallel()):
        assert '_cair_handle' not in clone.__dict__ and '_cair_key' not in clone.__dict__
        del clone
    assert stub.destroyed == []                       # dropping the copies released nothing
    net._release()
    assert stub.destroyed == [0x1234]                 # the owner releases exactly once
    net._release()
    assert stub.destroyed == [0x1234]


def test_invalidate_and_version_bumps(stub):
    net = _net()
    net.__dict__['_cair_handle'] = C.c_void_p(0x42)
    k0 = net._state_key()
    with torch.no_grad():
        net.output.bias.add_(1.0)                     # what an optimizer step does
    assert net._state_key() != k0                     # in-place updates are seen ...
    k1 = net._state_key()
    net.output.bias.data.mul_(2.0)                    # ... writes through .data are not (torch does not version them)
    assert net._state_key() == k1
    net.invalidate()                                  # hence the explicit hook
    assert stub.destroyed == [0x42] and net.__dict__.get('_cair_handle') is None

    class Vocab:
        ind2tok = {i: 'w%d' % i for i in range(net.word_embeddings.word_lut.weight.shape[0])}

        def __len__(self):
            return len(self.ind2tok)
    v0 = net.word_embeddings.word_lut.weight._version
    net.word_embeddings.init_word_vectors(Vocab(), {'w3': torch.ones(net.word_embeddings.word_vec_size)}, fixed=True)
    assert net.word_embeddings.word_lut.weight._version > v0        # load_embeddings rebuilds the handle on next use
    assert not net.word_embeddings.word_lut.weight.requires_grad


def test_host_entry_points_validate_their_arguments():
    net = _net()
    B, N, Lq, Ld = 2, 3, 5, 7
    q = torch.zeros(B, Lq, dtype=torch.int64)
    ql = torch.ones(B, dtype=torch.int64)
    d = torch.zeros(B, N, Ld, dtype=torch.int64)
    dl = torch.ones(B, N, dtype=torch.int64)
    out = torch.zeros(B, N)
    net._host_args(q, ql, d, dl, out, need_pinned=False)
    with pytest.raises(ValueError):
        net._host_args(q.int(), ql, d, dl, out, need_pinned=False)            # int32 ids
    with pytest.raises(ValueError):
        net._host_args(q, ql, d.transpose(1, 2), dl, out, need_pinned=False)   # not contiguous
    with pytest.raises(ValueError):
        net._host_args(q, ql[:1], d, dl, out, need_pinned=False)              # wrong number of lengths
    with pytest.raises(ValueError):
        net._host_args(q, ql, d, dl, torch.zeros(B, N + 1), need_pinned=False)
    with pytest.raises(ValueError):
        net._host_args(q, ql, d, dl, out, need_pinned=True)                   # asynchronous copies need pinned memory
    net._host_args(q, ql, d, dl, torch.zeros(4 * B * N), need_pinned=False, mult=4)   # gather attached: world * B * N scores
    with pytest.raises(RuntimeError):
        net.wait_host(0)                                                       # nothing submitted, no handle
    with pytest.raises(RuntimeError):
        net(q, ql, d, dl)                                                      # CPU tensors: no CPU path exists


def test_train_mode_dispatch_without_a_gpu():
    cfg, _, sd, _ = ol.load_golden('duet_tiny')
    duet = helpers.build_module(cfg, sd).train()
    x = torch.zeros(1, 8, dtype=torch.int64)
    with pytest.raises(NotImplementedError):
        duet(x, torch.ones(1, dtype=torch.int64), torch.zeros(1, 2, 30, dtype=torch.int64), torch.ones(1, 2, dtype=torch.int64))
    for cls in (cair.CARS, cair.MNSRF):
        assert cls.forward is not torch.nn.Module.forward       # the training entry point refuses with an explanation


def test_pair_slices_partition_the_batch():
    for total in (1, 7, 1280, 1283):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                b, c = parallel.pair_slice(r, world, total)
                covered.extend(range(b, b + c))
            assert covered == list(range(total)), (total, world)
