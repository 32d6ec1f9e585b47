"""Doc-parallel sharding of the scoring path across the GPUs of one box.

One process per GPU (torchrun); weights replicated; the flattened (query, doc) pairs p = b*N + n
are cut into contiguous, equal slices; every rank scores its slice with no data-path collective;
ONE all-gather of the per-pair fp32 scores (NCCL over NVLink on GPUs, gloo in the CPU tests)
gives every rank the full [B, N] matrix for the softmax / pairwise loss / MAP that need all N
candidates of a query (neuroir/models/ranker.py:87,258).  CARS shards by session instead
(sessions are independent; multitask/cars.py has no cross-session op except the click-mask width,
which libcair always computes over the replicated labels).
"""
import torch
import torch.distributed as dist


def pair_slice(rank, world, total):
    """Contiguous slice [begin, begin+count) of `total` units for `rank`; slices are ceil(total/world)
    long (balanced even when N=10, world=8); trailing ranks may get a short or empty slice."""
    per = (total + world - 1) // world
    begin = min(rank * per, total)
    return begin, min(per, total - begin)


def gather_scores(local, total, group=None):
    """local: 1-D tensor with this rank's slice (count may differ on the last ranks).
    Returns the 1-D tensor of all `total` scores on every rank (one all_gather)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:total]
    per = (total + world - 1) // world
    send = local.new_zeros(per)
    send[:local.numel()] = local
    recv = local.new_empty(per * world)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:total]


class P2PScoreGather:
    """The same all-gather as our own kernel over NVLink peer memory (cair_allgather_scores): every rank stores its slice
    straight into every peer's receive buffer (torch symmetric memory supplies the peer mappings) and waits on flags -
    one launch, no NCCL call.  `per` = slice length of every rank (the last ranks' short slices are padded by the caller's
    layout: slot r of the result starts at r * per).  The returned tensor is a view of one of two alternating receive
    buffers: it stays valid until the call after next."""

    def __init__(self, per, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.per, self.dev = dist.get_world_size(group), dist.get_rank(group), int(per), device
        n = self.world * self.per
        self.recv = symm_mem.empty(2 * n, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(64, dtype=torch.int32, device=device)
        self.recv.zero_()
        self.flags.zero_()
        hr = symm_mem.rendezvous(self.recv, self.group)
        hf = symm_mem.rendezvous(self.flags, self.group)
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # every rank's flags are zero before anyone publishes into them
        self._keep = (hr, hf)
        self._flag_ptrs = (C.c_uint64 * self.world)(*[int(p) for p in hf.buffer_ptrs])
        self._recv_ptrs = [(C.c_uint64 * self.world)(*[int(p) + par * n * 4 for p in hr.buffer_ptrs]) for par in (0, 1)]
        self.seq = 0
        self._lib, self._check = lib.load(), lib.check

    def attach(self, net):
        """Hands this instance's buffers to the ranker's host entry points (cair_ranker_set_gather): submit_host / wait_host
        then return ALL world * per scores per batch, gathered inside the serving pipeline.  The instance must not be
        called directly afterwards (the sequence numbers are the handle's from now on)."""
        dev = self.dev if isinstance(self.dev, torch.device) else torch.device(self.dev)
        h = net._handle_for(dev)
        self._check(self._lib.cair_ranker_set_gather(h, self._recv_ptrs[0], self._recv_ptrs[1], self._flag_ptrs, self.rank,
                                                     self.world, self.per))
        self._attached = net
        net.__dict__['_cair_gather_world'] = self.world
        net.__dict__['_cair_gather'] = self          # the peer mappings must outlive every batch the handle gathers
        return self

    def __call__(self, local, total):
        assert not getattr(self, '_attached', None), 'attached to a ranker: use its submit_host / wait_host'
        assert local.is_cuda and local.dtype == torch.float32 and local.is_contiguous() and local.numel() <= self.per
        send = local
        if local.numel() < self.per:              # short trailing slice: pad to the common slot length
            send = local.new_zeros(self.per)
            send[:local.numel()] = local
        self.seq += 1
        par = self.seq & 1
        self._check(self._lib.cair_allgather_scores(send.data_ptr(), self.per, self._recv_ptrs[par], self._flag_ptrs, self.rank,
                                                    self.world, self.seq, torch.cuda.current_stream(self.dev).cuda_stream))
        n = self.world * self.per
        return self.recv[par * n:par * n + n][:total]


class ShardedRanker:
    """Wraps a ranker network: each rank scores its pair slice, one all-gather assembles [B, N].
    `score_slice(q, qlen, d, dlen, begin, count) -> [B, N] tensor with the slice filled` defaults to the
    network's own pair_slice forward; tests inject a CPU scorer to exercise the plumbing under gloo."""

    def __init__(self, network, score_slice=None, group=None):
        self.network = network
        self.group = group
        self.score_slice = score_slice or (lambda q, ql, d, dl, b, c: network(q, ql, d, dl, pair_slice=(b, c)))

    def __call__(self, q, qlen, d, dlen):
        B, N = d.shape[0], d.shape[1]
        total = B * N
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        begin, count = pair_slice(rank, world, total)
        full = self.score_slice(q, qlen, d, dlen, begin, count)
        local = full.reshape(-1)[begin:begin + count].contiguous()
        return gather_scores(local, total, self.group).reshape(B, N)


class ShardedCars:
    """Session-sharded CARS scoring: rank r scores sessions [begin, begin+count), one all-gather of scores."""

    def __init__(self, network, score_slice=None, group=None):
        self.network = network
        self.group = group
        self.score_slice = score_slice or (
            lambda q, ql, d, dl, lab, b, c: network.score(q, ql, d, dl, lab, session_slice=(b, c))['scores'])

    def __call__(self, q, qlen, d, dlen, labels):
        B, S, N = d.shape[0], d.shape[1], d.shape[2]
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        begin, count = pair_slice(rank, world, B)
        full = self.score_slice(q, qlen, d, dlen, labels, begin, count)
        local = full.reshape(B, -1)[begin:begin + count].reshape(-1).contiguous()
        per = (B + world - 1) // world
        if world == 1:
            return full
        send = local.new_zeros(per * S * N)
        send[:local.numel()] = local
        recv = local.new_empty(per * S * N * world)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        return recv[:B * S * N].reshape(B, S, N)


class ShardedSessionRanker(ShardedCars):
    """Session-sharded MNSRF / M_MATCH_TENSOR scoring (no labels on their ranking path): rank r scores sessions
    [begin, begin+count), one all-gather of scores.  `network.score(q, qlen, d, dlen, session_slice=...)` for MNSRF; the
    default for networks without a session slice (M_MATCH_TENSOR scores every (session, query) row independently) slices
    the batch itself."""

    def __init__(self, network, score_slice=None, group=None):
        def default(q, ql, d, dl, lab, b, c):
            import inspect
            if 'session_slice' in inspect.signature(network.score).parameters:
                return network.score(q, ql, d, dl, session_slice=(b, c))['scores']
            full = q.new_zeros(d.shape[:3], dtype=torch.float32)
            if c > 0:
                full[b:b + c] = network.score(q[b:b + c], ql[b:b + c], d[b:b + c], dl[b:b + c])
            return full
        super().__init__(network, score_slice=(lambda q, ql, d, dl, lab, b, c: score_slice(q, ql, d, dl, b, c)) if score_slice else default,
                         group=group)

    def __call__(self, q, qlen, d, dlen):
        return super().__call__(q, qlen, d, dlen, None)
