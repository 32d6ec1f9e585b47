"""Host-side mirror of the reference's ranker networks (neuroir/rankers/*.py).

Same constructor contract (an argparse.Namespace from neuroir.config.get_model_args plus
src_vocab_size), same call contract `network(queries, que_len, documents, doc_len) ->
FloatTensor[B, N]` (neuroir/models/ranker.py:213,257), same parameter names / shapes /
initialisers as the reference modules, so reference state_dicts load by key.
The arithmetic is not here: forward() hands device pointers to libcair.so.  Inputs must be CUDA
tensors; there is no CPU or eager-PyTorch fallback.  In eval mode a module scores through its libcair
handle; in train mode MatchTensor runs libcair's training step (train-mode forward + hand-written
backward behind a torch.autograd.Function), so the reference's Ranker.update (loss, backward, clipping,
optimizer) runs unchanged on top.  The other rankers raise in train mode.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _abi, lib

PAD = 0


class Embeddings(nn.Module):
    """Parameter container mirroring neuroir/modules/embeddings.py:124-196 (word LUT only):
    key `make_embedding.emb_luts.0.weight`, PAD row zero."""

    def __init__(self, word_vec_size, word_vocab_size, word_padding_idx=PAD, fix_word_vecs=False):
        super().__init__()
        self.word_vec_size = word_vec_size
        self.embedding_size = word_vec_size
        luts = nn.ModuleList([nn.Embedding(word_vocab_size, word_vec_size, padding_idx=word_padding_idx)])
        self.make_embedding = nn.Sequential(OrderedDict([('emb_luts', luts)]))
        if fix_word_vecs:
            self.word_lut.weight.requires_grad = False

    @property
    def word_lut(self):
        return self.make_embedding[0][0]

    @property
    def emb_luts(self):
        return self.make_embedding[0]

    def init_word_vectors(self, vocabulary, embeddings_index, fixed):
        """Same contract as embeddings.py:216-229: rows of words absent from the index become zero."""
        pretrained = torch.zeros(len(vocabulary), self.word_vec_size)
        for i in range(len(vocabulary)):
            tok = vocabulary.ind2tok[i]
            if tok in embeddings_index:
                pretrained[i] = embeddings_index[tok]
        with torch.no_grad():   # in-place copy that bumps the parameter version: the libcair handle is rebuilt on next use
            self.word_lut.weight.copy_(pretrained)
        if fixed:
            self.word_lut.weight.requires_grad = False

    def forward(self, source):  # [B, L, 1] ids -> [B, L, E]; only used by callers outside the fused path
        ids = source.squeeze(-1).contiguous()
        w = self.word_lut.weight
        if not ids.is_cuda:
            raise RuntimeError('context_attentive_ir_b200 runs on CUDA tensors only')
        out = torch.empty(ids.shape + (w.shape[1],), device=ids.device, dtype=torch.float32)
        lib.check(lib.load().cair_embed_gather(w.data_ptr(), w.shape[0], w.shape[1], ids.data_ptr(), ids.numel(),
                                               out.data_ptr(), torch.cuda.current_stream(ids.device).cuda_stream))
        return out


class RNNEncoder(nn.Module):
    """Parameter container mirroring neuroir/encoders/rnn_encoder.py:25-60 (keys `rnns.<i>.*`)."""

    def __init__(self, rnn_type, input_size, bidirectional, num_layers, hidden_size, dropout=0.0):
        super().__init__()
        dirs = 2 if bidirectional else 1
        assert hidden_size % dirs == 0
        self.rnn_type, self.bidirectional, self.nlayers = rnn_type, bidirectional, num_layers
        self.hidden_size = hidden_size // dirs
        self.rnns = nn.ModuleList()
        for i in range(num_layers):
            in_sz = input_size if i == 0 else self.hidden_size * dirs
            self.rnns.append(getattr(nn, rnn_type)(input_size=in_sz, hidden_size=self.hidden_size, num_layers=1,
                                                    bidirectional=bidirectional, batch_first=True))
        self.dropout = nn.Dropout(dropout)


def _named_tensors(module, prefix=''):
    """(name, tensor) of every parameter and buffer, including the plain-tensor parameter copies that
    nn.DataParallel's replicate() leaves in `_former_parameters` of a replica (a replica has no Parameters)."""
    for name, p in module._parameters.items():
        if p is not None:
            yield prefix + name, p
    for name, p in getattr(module, '_former_parameters', {}).items():
        if p is not None:
            yield prefix + name, p
    for name, b in module._buffers.items():
        if b is not None:
            yield prefix + name, b
    for name, child in module._modules.items():
        if child is not None:
            yield from _named_tensors(child, prefix + name + '.')


_HANDLE_KEYS = ('_cair_handle', '_cair_key', '_cair_ws', '_cair_trainer', '_cair_trainer_key', '_cair_sessdec', '_cair_sessdec_key', '_fwd', '_cair_gather_world', '_cair_gather')


def _ptr_getter(module, keep):
    sd = dict(_named_tensors(module))

    def get(key):
        t = sd[key]
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.detach().float().contiguous()
            keep.append(t)
        return C.cast(t.data_ptr(), _abi.f32p)
    return get


class _CairModule(nn.Module):
    """Owns the libcair handle; rebuilds it when a parameter changed (version / storage / device)."""
    MODEL = None

    def _cfg(self):
        raise NotImplementedError

    def _create(self, weights, device, out):
        raise NotImplementedError

    def _state_key(self):
        """Storage, version counter and device of every weight.  In-place ops (optimizer steps, `copy_`, `load_state_dict`)
        bump the version; writes through `.data` do NOT - call invalidate() after editing a parameter that way."""
        return tuple((t.data_ptr(), t._version, str(t.device)) for _, t in _named_tensors(self))

    def invalidate(self):
        """Drop the native handle so the next call re-reads (and re-packs) the weights; needed only after a weight was
        modified through `.data`, which torch's version counters cannot see."""
        self._release()
        return self

    # The native handle belongs to exactly one Python object: copies (copy.copy / deepcopy / pickle / DataParallel
    # replicas) start without one and create their own on first use.
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in _HANDLE_KEYS:
            state.pop(k, None)
        return state

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        for k in _HANDLE_KEYS:
            replica.__dict__.pop(k, None)
        return replica

    def _handle_for(self, device):
        key = (self._state_key(), device.index)
        h = self.__dict__.get('_cair_handle')
        if h is not None and self.__dict__.get('_cair_key') == key:
            return h
        self._release()
        if any(not t.is_cuda for _, t in _named_tensors(self)):
            raise RuntimeError('%s parameters must be on a CUDA device (call .cuda()); no CPU path exists'
                               % type(self).__name__)
        keep = []
        w = _abi.PACKERS[self.MODEL](self._cfg(), _ptr_getter(self, keep))
        out = C.c_void_p()
        torch.cuda.synchronize(device)  # weights may still be in flight on another stream
        lib.check(self._create(C.byref(w), device.index, C.byref(out)))
        self.__dict__['_cair_handle'] = out
        self.__dict__['_cair_key'] = key
        self.__dict__['_cair_ws'] = None
        self._on_handle_created(out)
        return out

    def _on_handle_created(self, handle):
        pass

    def _release(self):
        h = self.__dict__.get('_cair_handle')
        if h is not None:
            lib.load().cair_destroy(h)
            self.__dict__['_cair_handle'] = None
        self.__dict__.pop('_cair_gather_world', None)   # a score gather is attached to a handle: re-attach after a rebuild
        self.__dict__.pop('_cair_gather', None)

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _workspace(self, nbytes, device):
        ws = self.__dict__.get('_cair_ws')
        if ws is None or ws.numel() < nbytes or ws.device != device:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self.__dict__['_cair_ws'] = ws
        return ws

    @staticmethod
    def _ids(t, name):
        if not torch.is_tensor(t):
            t = torch.as_tensor(t)
        if not t.is_cuda:
            raise RuntimeError('%s must be a CUDA tensor (context_attentive_ir_b200 has no CPU path)' % name)
        return t.to(torch.int64).contiguous()


class _Ranker(_CairModule):
    def forward(self, batch_queries, query_len, batch_docs, doc_len, pair_slice=None):
        """scores[B, N] (fp32, no softmax).  pair_slice=(begin, count) scores only that contiguous
        slice of the flattened pairs (doc-parallel sharding); the rest of the output is left zero."""
        assert batch_queries.shape[0] == batch_docs.shape[0]  # rankers/*.py, e.g. mtensor.py:71
        if self.training:
            if pair_slice is not None:
                raise NotImplementedError('pair_slice is a scoring-path (eval) feature')
            return self._train_forward(batch_queries, query_len, batch_docs, doc_len)
        q = self._ids(batch_queries, 'batch_queries')
        d = self._ids(batch_docs, 'batch_docs')
        ql = self._ids(query_len, 'query_len').to(q.device)
        dl = self._ids(doc_len, 'doc_len').to(q.device).reshape(d.shape[0], d.shape[1])
        B, Lq = q.shape
        _, N, Ld = d.shape
        dev = q.device
        L = lib.load()
        h = self._handle_for(dev)
        nbytes = C.c_size_t()
        lib.check(L.cair_ranker_workspace_bytes(h, B, N, Lq, Ld, C.byref(nbytes)))
        ws = self._workspace(nbytes.value, dev)
        begin, count = (0, B * N) if pair_slice is None else pair_slice
        scores = torch.zeros(B, N, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib.check(L.cair_ranker_forward(h, q.data_ptr(), ql.data_ptr(), d.data_ptr(), dl.data_ptr(), B, N, Lq, Ld,
                                        begin, count, scores.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return scores

    def _train_forward(self, batch_queries, query_len, batch_docs, doc_len):
        raise NotImplementedError('%s: the libcair training step exists for MatchTensor, DRMM, ESM, DSSM and CDSSM only; score under .eval()'
                                  % type(self).__name__)

    @staticmethod
    def _host_args(q, qlen, d, dlen, out, need_pinned, mult=1):
        """The host entry points take raw pointers: insist on CPU int64 contiguous ids / lengths of the documented shapes
        and a CPU float32 [B, N] result buffer (pinned where the copy is asynchronous)."""
        for name, t in (('q', q), ('qlen', qlen), ('d', d), ('dlen', dlen)):
            if not torch.is_tensor(t) or t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous():
                raise ValueError('%s must be a contiguous CPU int64 tensor' % name)
            if need_pinned and not t.is_pinned():
                raise ValueError('%s must be in pinned host memory (asynchronous copy)' % name)
        if q.dim() != 2 or d.dim() != 3 or d.shape[0] != q.shape[0]:
            raise ValueError('q must be [B, Lq] and d [B, N, Ld]')
        B, N = d.shape[0], d.shape[1]
        if qlen.numel() != B or dlen.numel() != B * N:
            raise ValueError('qlen must hold B and dlen B*N lengths')
        if out is not None:
            if (not torch.is_tensor(out) or out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous()
                    or out.numel() != B * N * mult):
                raise ValueError('out must be a contiguous CPU float32 tensor with B*N elements (world*B*N when a score '
                                 'gather is attached to the handle)')
            if need_pinned and not out.is_pinned():
                raise ValueError('out must be in pinned host memory (asynchronous copy)')

    def forward_host(self, q, qlen, d, dlen, out=None, device=None):
        """End-to-end entry point on HOST tensors (pinned recommended): ids are copied host->device,
        scored, and the scores copied back; returns a CPU tensor.  Raises on bad token ids / lengths."""
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._host_args(q, qlen, d, dlen, out, need_pinned=False)
        B, Lq = q.shape
        _, N, Ld = d.shape
        if out is None:
            out = torch.empty(B, N, dtype=torch.float32).pin_memory()
        L = lib.load()
        h = self._handle_for(dev)
        lib.check(L.cair_ranker_forward_host(h, q.data_ptr(), qlen.data_ptr(), d.data_ptr(), dlen.data_ptr(),
                                             B, N, Lq, Ld, out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return out

    def submit_host(self, q, qlen, d, dlen, out, slot, device=None, stream=None):
        """Pipelined form of forward_host: enqueue H2D of the (pinned) id tensors, the scoring kernels and the D2H
        of the scores into `out` (pinned) without waiting; up to three batches (slots 0, 1, 2) are in flight, so the copies
        of one batch overlap the kernels of the previous one and, for Match-Tensor, the interaction kernel of batch k
        runs partly under the document encoder of batch k+1.  Call wait_host(slot) before reading `out` or re-using
        the slot."""
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if out is None:
            raise ValueError('submit_host needs a pinned float32 result buffer')
        self._host_args(q, qlen, d, dlen, out, need_pinned=True, mult=self.__dict__.get('_cair_gather_world', 1))
        B, Lq = q.shape
        _, N, Ld = d.shape
        h = self._handle_for(dev)
        lib.check(lib.load().cair_ranker_submit_host(h, q.data_ptr(), qlen.data_ptr(), d.data_ptr(), dlen.data_ptr(),
                                                     B, N, Lq, Ld, out.data_ptr(), slot,
                                                     stream.cuda_stream if stream is not None else None))

    def wait_host(self, slot):
        h = self.__dict__.get('_cair_handle')
        if h is None:
            raise RuntimeError('wait_host(%d): nothing was submitted (no native handle yet)' % slot)
        lib.check(lib.load().cair_ranker_wait_host(h, slot))

    def poll_error(self):
        h = self.__dict__.get('_cair_handle')
        if h is not None:
            lib.check(lib.load().cair_poll_error(h, torch.cuda.current_stream().cuda_stream))


class ESM(_Ranker):
    """neuroir/rankers/esm.py:11-45."""
    MODEL = 'esm'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)

    def _cfg(self):
        return dict(src_vocab_size=self.args.src_vocab_size, emsize=self.args.emsize)

    def _create(self, w, device, out):
        return lib.load().cair_esm_create(w, device, out)

    def _train_forward(self, batch_queries, query_len, batch_docs, doc_len):
        """Train mode (Ranker.update): cair_esm_train_forward / cair_esm_train_backward on the live embedding table."""
        q = self._ids(batch_queries, 'batch_queries')
        d = self._ids(batch_docs, 'batch_docs')
        table = self.word_embeddings.word_lut.weight
        if not table.is_cuda or table.dtype != torch.float32 or not table.is_contiguous():
            raise RuntimeError('ESM training needs a contiguous fp32 CUDA embedding table')
        return _EsmTrainFn.apply(q, d, table)


class _EsmTrainFn(torch.autograd.Function):
    """Train-mode ESM scores through libcair with libcair's backward into the embedding table (skipped under --fix_embeddings)."""

    @staticmethod
    def forward(ctx, q, d, table):
        dev = q.device
        L = lib.load()
        B, Lq = q.shape
        _, N, Ld = d.shape
        V, E = table.shape
        nbytes = C.c_size_t()
        lib.check(L.cair_esm_train_workspace_bytes(E, B, N, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        lib.check(L.cair_esm_train_forward(table.data_ptr(), V, E, q.data_ptr(), d.data_ptr(), B, N, Lq, Ld, scores.data_ptr(),
                                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.args = (q, d, ws, scores, table.shape, table.requires_grad)
        return scores.clone()

    @staticmethod
    def backward(ctx, dscores):
        q, d, ws, scores, shape, req = ctx.args
        dev = q.device
        B, Lq = q.shape
        _, N, Ld = d.shape
        dtable = torch.zeros(shape, dtype=torch.float32, device=dev) if req else None
        dscores = dscores.contiguous().float()
        lib.check(lib.load().cair_esm_train_backward(shape[0], shape[1], q.data_ptr(), d.data_ptr(), B, N, Lq, Ld, scores.data_ptr(),
                                                     dscores.data_ptr(), dtable.data_ptr() if req else None, ws.data_ptr(),
                                                     ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return None, None, dtable


class DSSM(_Ranker):
    """neuroir/rankers/dssm.py:7-63."""
    MODEL = 'dssm'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        self.query_mlp = nn.Sequential(nn.Linear(args.emsize, args.nhid), nn.Tanh(), nn.Linear(args.nhid, args.nout), nn.Tanh())
        self.doc_mlp = nn.Sequential(nn.Linear(args.emsize, args.nhid), nn.Tanh(), nn.Linear(args.nhid, args.nout), nn.Tanh())

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nhid=a.nhid, nout=a.nout)

    def _create(self, w, device, out):
        return lib.load().cair_dssm_create(w, device, out)

    def _train_forward(self, batch_queries, query_len, batch_docs, doc_len):
        """Train mode (Ranker.update): cair_dssm_train_forward / cair_dssm_train_backward on the live parameters."""
        q = self._ids(batch_queries, 'batch_queries')
        d = self._ids(batch_docs, 'batch_docs')
        for name, p in self.named_parameters():
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('%s training needs contiguous fp32 CUDA parameters (%s)' % (type(self).__name__, name))
        seed = self.__dict__.get('_cair_drop_seed')   # tests pin the mask; otherwise drawn from torch's host generator
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        names = [n for n, _ in self.named_parameters()]
        return _DssmTrainFn.apply(self, names, q, d, float(self.emb_drop.p), seed, *[p for _, p in self.named_parameters()])


class _DssmTrainFn(torch.autograd.Function):
    """Train-mode DSSM scores through libcair with libcair's backward for the two MLPs and, through the max-pool's arg-max
    positions and the dropout mask, the embedding table (skipped under --fix_embeddings)."""

    @staticmethod
    def forward(ctx, module, names, q, d, p_drop, seed, *params):
        dev = q.device
        L = lib.load()
        B, Lq = q.shape
        _, N, Ld = d.shape
        a = module.args
        live = dict(zip(names, params))
        kind = module.MODEL   # 'dssm' or 'cdssm': same call shapes, cair_<kind>_train_*
        w = _abi.PACKERS[kind](module._cfg(), lambda k: C.cast(live[k].data_ptr(), _abi.f32p))
        nbytes = C.c_size_t()
        if kind == 'dssm':
            lib.check(L.cair_dssm_train_workspace_bytes(a.emsize, a.nhid, a.nout, B, N, C.byref(nbytes)))
        else:
            lib.check(L.cair_cdssm_train_workspace_bytes(a.emsize, a.nhid, a.nout, B, N, Lq, Ld, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        lib.check(getattr(L, 'cair_%s_train_forward' % kind)(C.byref(w), q.data_ptr(), d.data_ptr(), B, N, Lq, Ld, p_drop, seed, scores.data_ptr(),
                                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.module, ctx.names, ctx.args = module, names, (q, d, p_drop, seed, ws, scores, params)
        ctx.shapes = [(p.shape, p.requires_grad) for p in params]
        return scores.clone()

    @staticmethod
    def backward(ctx, dscores):
        module, names = ctx.module, ctx.names
        q, d, p_drop, seed, ws, scores, params = ctx.args
        dev = q.device
        B, Lq = q.shape
        _, N, Ld = d.shape
        live = dict(zip(names, params))
        grads = {}
        for name, (shape, req) in zip(names, ctx.shapes):
            grads[name] = None if (name == _abi.TABLE_KEY and not req) else torch.zeros(shape, dtype=torch.float32, device=dev)
        kind = module.MODEL
        w = _abi.PACKERS[kind](module._cfg(), lambda k: C.cast(live[k].data_ptr(), _abi.f32p))
        gw = _abi.PACKERS[kind](module._cfg(), lambda k: C.cast(grads[k].data_ptr() if grads[k] is not None else None, _abi.f32p))
        dscores = dscores.contiguous().float()
        lib.check(getattr(lib.load(), 'cair_%s_train_backward' % kind)(C.byref(w), C.byref(gw), q.data_ptr(), d.data_ptr(), B, N, Lq, Ld, p_drop, seed,
                                                      scores.data_ptr(), dscores.data_ptr(), ws.data_ptr(), ws.numel(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
        return (None,) * 6 + tuple(grads[n] if req else None for n, (_, req) in zip(names, ctx.shapes))


class CDSSM(_Ranker):
    """neuroir/rankers/cdssm.py:7-77."""
    MODEL = 'cdssm'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.window = 3
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        self.query_conv = nn.Conv1d(self.window * args.emsize, args.nhid, 3)
        self.query_sem = nn.Linear(args.nhid, args.nout)
        self.doc_conv = nn.Conv1d(self.window * args.emsize, args.nhid, 3)
        self.doc_sem = nn.Linear(args.nhid, args.nout)

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nhid=a.nhid, nout=a.nout)

    def _create(self, w, device, out):
        return lib.load().cair_cdssm_create(w, device, out)

    _train_forward = DSSM._train_forward   # same call shape: cair_cdssm_train_forward / cair_cdssm_train_backward


class ARCI(_Ranker):
    """neuroir/rankers/arci.py:7-105."""
    MODEL = 'arci'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        nl = len(args.filters_1d)
        assert nl == len(args.kernel_size_1d) == len(args.maxpool_size_1d)
        qf, df = args.max_query_len, args.max_doc_len
        ql, dl = [], []
        for i in range(nl):
            inp = args.emsize if i == 0 else args.filters_1d[i - 1]
            for lst in (ql, dl):
                lst.append(nn.Sequential(nn.Conv1d(inp, args.filters_1d[i], args.kernel_size_1d[i],
                                                   padding=args.kernel_size_1d[i] // 2),
                                         nn.ReLU(inplace=True), nn.MaxPool1d(args.maxpool_size_1d[i])))
            df, qf = df // args.maxpool_size_1d[i], qf // args.maxpool_size_1d[i]
            assert qf != 0 and df != 0
        self.query_conv1d_layers, self.doc_conv1d_layers = nn.ModuleList(ql), nn.ModuleList(dl)
        inp = args.filters_1d[-1] * (qf + df)
        self.mlp = nn.Sequential(nn.Linear(inp, inp // 2), nn.Linear(inp // 2, 1))

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, filters_1d=list(a.filters_1d),
                    kernel_size_1d=list(a.kernel_size_1d), maxpool_size_1d=list(a.maxpool_size_1d),
                    max_query_len=a.max_query_len, max_doc_len=a.max_doc_len)

    def _create(self, w, device, out):
        return lib.load().cair_arci_create(w, device, out)


class ARCII(_Ranker):
    """neuroir/rankers/arcii.py:7-111."""
    MODEL = 'arcii'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        self.conv_query = nn.Conv1d(args.emsize, args.filters_1d, args.kernel_size_1d, padding=args.kernel_size_1d // 2)
        self.conv_doc = nn.Conv1d(args.emsize, args.filters_1d, args.kernel_size_1d, padding=args.kernel_size_1d // 2)
        self.maxpool1 = nn.MaxPool2d((2, 2))
        nl = len(args.kernel_size_2d)
        assert nl == len(args.maxpool_size_2d)
        df, qf = args.max_doc_len // 2, args.max_query_len // 2
        layers = []
        for i in range(nl):
            inp = args.filters_1d if i == 0 else args.filters_2d[i - 1]
            ks, mp = args.kernel_size_2d[i], args.maxpool_size_2d[i]
            layers.append(nn.Sequential(nn.Conv2d(inp, args.filters_2d[i], tuple(ks), padding=(ks[0] // 2, ks[1] // 2)),
                                        nn.ReLU(inplace=True), nn.MaxPool2d((mp[0], mp[1]))))
            df, qf = df // mp[0], qf // mp[1]
            assert qf != 0 and df != 0
        self.conv2d_layers = nn.ModuleList(layers)
        inp = args.filters_2d[-1] * qf * df
        self.mlp = nn.Sequential(nn.Linear(inp, inp // 2), nn.Linear(inp // 2, 1))

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, filters_1d=a.filters_1d, kernel_size_1d=a.kernel_size_1d,
                    filters_2d=list(a.filters_2d), kernel_size_2d=[list(k) for k in a.kernel_size_2d],
                    maxpool_size_2d=[list(k) for k in a.maxpool_size_2d], max_query_len=a.max_query_len,
                    max_doc_len=a.max_doc_len)

    def _create(self, w, device, out):
        return lib.load().cair_arcii_create(w, device, out)


class ExactMatchChannel(nn.Module):
    """neuroir/rankers/mtensor.py:134-142: one learnable scalar, U(0,1) init."""

    def __init__(self):
        super().__init__()
        self.alpha = nn.Parameter(torch.empty(1))
        nn.init.uniform_(self.alpha)


class MatchTensor(_Ranker):
    """neuroir/rankers/mtensor.py:24-131."""
    MODEL = 'match_tensor'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        self.linear_projection = nn.Linear(args.emsize, args.featsize)
        self.query_encoder = RNNEncoder(args.rnn_type, args.featsize, args.bidirection, args.nlayers,
                                        args.nhid_query, args.dropout_rnn)
        self.document_encoder = RNNEncoder(args.rnn_type, args.featsize, args.bidirection, args.nlayers,
                                           args.nhid_doc, args.dropout_rnn)
        self.query_projection = nn.Linear(args.nhid_query, args.nchannels)
        self.document_projection = nn.Linear(args.nhid_doc, args.nchannels)
        self.exact_match_channel = ExactMatchChannel()
        self.conv1 = nn.Conv2d(args.nchannels + 1, args.nfilters, (3, 3), padding=1)
        self.conv2 = nn.Conv2d(args.nchannels + 1, args.nfilters, (3, 5), padding=(1, 2))
        self.conv3 = nn.Conv2d(args.nchannels + 1, args.nfilters, (3, 7), padding=(1, 3))
        self.conv = nn.Conv2d(args.nfilters * 3, args.match_filter_size, (1, 1))
        self.output = nn.Linear(args.match_filter_size, 1)
        if not 1 <= args.nlayers <= 4:
            raise NotImplementedError('libcair stacks at most 4 encoder layers (neuroir/hyparam.py:88-100 uses 1)')

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, featsize=a.featsize, nhid_query=a.nhid_query,
                    nhid_doc=a.nhid_doc, nchannels=a.nchannels, nfilters=a.nfilters,
                    match_filter_size=a.match_filter_size, rnn_type=a.rnn_type, bidirection=a.bidirection)

    def _create(self, w, device, out):
        return lib.load().cair_mt_create(w, device, out)

    # ---- training: cair_mt_train_forward / cair_mt_train_backward ----
    def _trainer_for(self, device):
        """The native trainer reads the LIVE parameter storage at every step; it is rebuilt only when a parameter moved."""
        params = dict(self.named_parameters())
        for name, p in params.items():
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('MatchTensor training needs contiguous fp32 CUDA parameters (%s)' % name)
        key = (tuple(p.data_ptr() for p in params.values()), device.index)
        t = self.__dict__.get('_cair_trainer')
        if t is not None and self.__dict__.get('_cair_trainer_key') == key:
            return t
        self._release_trainer()
        w = _abi.PACKERS[self.MODEL](self._cfg(), lambda k: C.cast(params[k].data_ptr(), _abi.f32p))
        out = C.c_void_p()
        lib.check(lib.load().cair_mt_train_create(C.byref(w), device.index, C.byref(out)))
        self.__dict__['_cair_trainer'], self.__dict__['_cair_trainer_key'] = out, key
        if self.__dict__.get('_cair_train_tc') is not None:
            lib.check(lib.load().cair_mt_train_set_impl(out, int(self.__dict__['_cair_train_tc'])))
        return out

    def set_training_impl(self, tc_forward=True):
        """Training forward interaction on the tcgen05 kernel (default) or the fp32 kernel."""
        self.__dict__['_cair_train_tc'] = bool(tc_forward)
        t = self.__dict__.get('_cair_trainer')
        if t is not None:
            lib.check(lib.load().cair_mt_train_set_impl(t, int(bool(tc_forward))))
        return self

    def _release_trainer(self):
        t = self.__dict__.get('_cair_trainer')
        if t is not None:
            lib.load().cair_mt_train_destroy(t)
            self.__dict__['_cair_trainer'] = None

    def _release(self):
        super()._release()
        self._release_trainer()

    def _train_forward(self, batch_queries, query_len, batch_docs, doc_len):
        if self.args.nlayers != 1:
            raise NotImplementedError('MatchTensor: the libcair training step covers single-layer encoders; score under .eval()')
        q = self._ids(batch_queries, 'batch_queries')
        d = self._ids(batch_docs, 'batch_docs')
        ql = self._ids(query_len, 'query_len').to(q.device)
        dl = self._ids(doc_len, 'doc_len').to(q.device).reshape(d.shape[0], d.shape[1])
        p_drop = float(self.emb_drop.p)
        seed = self.__dict__.get('_cair_drop_seed')   # tests pin the mask; otherwise drawn from torch's host generator
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        names = [n for n, _ in self.named_parameters()]
        return _MtTrainFn.apply(self, names, q, ql, d, dl, p_drop, seed, *[p for _, p in self.named_parameters()])

    def set_interaction_impl(self, impl):
        """'tc' (default): tcgen05 bf16x3 tensor-core kernels; 'fp32': CUDA-core fp32 kernels;
        'tc_split': tensor-core interaction with the unfused fp32 document projection."""
        self.__dict__['_cair_impl'] = {'fp32': 0, 'tc': 1, 'tc_split': 2}[impl]
        h = self.__dict__.get('_cair_handle')
        if h is not None:
            lib.check(lib.load().cair_mt_set_impl(h, self.__dict__['_cair_impl']))
        return self

    def _on_handle_created(self, handle):
        impl = self.__dict__.get('_cair_impl')
        if impl is not None:
            lib.check(lib.load().cair_mt_set_impl(handle, impl))
        # stacked encoders (rnn_encoder.py:45-53): layers 1.. are appended to the handle, which packs (copies) them
        keep = []
        get = _ptr_getter(self, keep)
        for side, enc in enumerate(('query_encoder', 'document_encoder')):
            for k in range(1, self.args.nlayers):
                fwd = _abi._lstm(get, '%s.rnns.%d' % (enc, k))
                rev = C.byref(_abi._lstm(get, '%s.rnns.%d' % (enc, k), '_reverse')) if self.args.bidirection else None
                lib.check(lib.load().cair_mt_add_encoder_layer(handle, side, C.byref(fwd), rev))


class _MtTrainFn(torch.autograd.Function):
    """scores = MatchTensor(q, d) in train mode through libcair, with libcair's backward: gradients for every parameter
    that requires one (the embedding table is skipped when its requires_grad is False, i.e. --fix_embeddings)."""

    @staticmethod
    def forward(ctx, module, names, q, ql, d, dl, p_drop, seed, *params):
        dev = q.device
        L = lib.load()
        t = module._trainer_for(dev)
        B, Lq = q.shape
        _, N, Ld = d.shape
        nbytes = C.c_size_t()
        lib.check(L.cair_mt_train_workspace_bytes(t, B, N, Lq, Ld, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib.check(L.cair_mt_train_forward(t, q.data_ptr(), ql.data_ptr(), d.data_ptr(), dl.data_ptr(), B, N, Lq, Ld, p_drop,
                                          seed, scores.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        ctx.module, ctx.names, ctx.args = module, names, (q, ql, d, dl, p_drop, seed, ws, t)
        ctx.shapes = [(p.shape, p.requires_grad) for p in params]
        return scores

    @staticmethod
    def backward(ctx, dscores):
        module, names = ctx.module, ctx.names
        q, ql, d, dl, p_drop, seed, ws, t = ctx.args
        dev = q.device
        B, Lq = q.shape
        _, N, Ld = d.shape
        dscores = dscores.contiguous().float()
        grads, keep = {}, []
        for name, (shape, req) in zip(names, ctx.shapes):
            if name == _abi.TABLE_KEY and not req:
                grads[name] = None            # fixed embeddings: the scatter-add is skipped
            else:
                grads[name] = torch.zeros(shape, dtype=torch.float32, device=dev)

        def get(k):
            g = grads[k]
            return C.cast(g.data_ptr(), _abi.f32p) if g is not None else C.cast(None, _abi.f32p)
        gw = _abi.PACKERS[module.MODEL](module._cfg(), get)
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib.check(lib.load().cair_mt_train_backward(t, q.data_ptr(), ql.data_ptr(), d.data_ptr(), dl.data_ptr(), B, N, Lq, Ld,
                                                    p_drop, seed, dscores.data_ptr(), C.byref(gw), ws.data_ptr(), ws.numel(),
                                                    stream))
        out = [grads[n] if req else None for n, (_, req) in zip(names, ctx.shapes)]
        return (None,) * 8 + tuple(out)


class _DrmmTrainFn(torch.autograd.Function):
    """Train-mode DRMM scores through libcair (cair_drmm_train_forward) with libcair's backward for the gate / ffnn / output
    parameters and, through the gate, the embedding rows of the query tokens."""

    @staticmethod
    def forward(ctx, module, names, q, d, p_drop, seed, *params):
        dev = q.device
        L = lib.load()
        B, Lq = q.shape
        _, N, Ld = d.shape
        live = dict(zip(names, params))
        w = _abi.PACKERS['drmm'](module._cfg(), lambda k: C.cast(live[k].data_ptr(), _abi.f32p))
        nbytes = C.c_size_t()
        lib.check(L.cair_drmm_train_workspace_bytes(module.args.emsize, B, N, Lq, Ld, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        lib.check(L.cair_drmm_train_forward(C.byref(w), q.data_ptr(), d.data_ptr(), B, N, Lq, Ld, p_drop, seed, scores.data_ptr(),
                                            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        ctx.module, ctx.names, ctx.args = module, names, (q, d, p_drop, seed, ws, params)
        ctx.shapes = [(p.shape, p.requires_grad) for p in params]
        return scores

    @staticmethod
    def backward(ctx, dscores):
        module, names = ctx.module, ctx.names
        q, d, p_drop, seed, ws, params = ctx.args
        dev = q.device
        B, Lq = q.shape
        _, N, Ld = d.shape
        live = dict(zip(names, params))
        grads = {}
        for name, (shape, req) in zip(names, ctx.shapes):
            grads[name] = None if (name == _abi.TABLE_KEY and not req) else torch.zeros(shape, dtype=torch.float32, device=dev)
        w = _abi.PACKERS['drmm'](module._cfg(), lambda k: C.cast(live[k].data_ptr(), _abi.f32p))
        gw = _abi.PACKERS['drmm'](module._cfg(), lambda k: C.cast(grads[k].data_ptr() if grads[k] is not None else None, _abi.f32p))
        dscores = dscores.contiguous().float()
        lib.check(lib.load().cair_drmm_train_backward(C.byref(w), C.byref(gw), q.data_ptr(), B, N, Lq, Ld, p_drop, seed,
                                                      dscores.data_ptr(), ws.data_ptr(), ws.numel(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
        return (None,) * 6 + tuple(grads[n] if req else None for n, (_, req) in zip(names, ctx.shapes))


class GatingNetwork(nn.Module):
    """neuroir/rankers/drmm.py:87-93."""

    def __init__(self, emsize):
        super().__init__()
        self.weight = nn.Linear(emsize, 1)


class DRMM(_Ranker):
    """neuroir/rankers/drmm.py:10-84."""
    MODEL = 'drmm'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        self.nbins = args.nbins
        self.bins = [-1.0, -0.5, 0, 0.5, 1.0, 1.0]
        self.gating_network = GatingNetwork(args.emsize)
        self.ffnn = nn.Sequential(nn.Linear(self.nbins, 1), nn.Linear(1, 1))
        self.output = nn.Linear(1, 1)

    def _cfg(self):
        return dict(src_vocab_size=self.args.src_vocab_size, emsize=self.args.emsize, nbins=self.nbins)

    def _create(self, w, device, out):
        return lib.load().cair_drmm_create(w, device, out)

    def _train_forward(self, batch_queries, query_len, batch_docs, doc_len):
        """Train mode (Ranker.update): cair_drmm_train_forward / cair_drmm_train_backward on the live parameters."""
        q = self._ids(batch_queries, 'batch_queries')
        d = self._ids(batch_docs, 'batch_docs')
        for name, p in self.named_parameters():
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('DRMM training needs contiguous fp32 CUDA parameters (%s)' % name)
        seed = self.__dict__.get('_cair_drop_seed')
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        names = [n for n, _ in self.named_parameters()]
        return _DrmmTrainFn.apply(self, names, q, d, float(self.emb_drop.p), seed, *[p for _, p in self.named_parameters()])


class LocalModel(nn.Module):
    """neuroir/rankers/duet.py:62-75."""

    def __init__(self, args):
        super().__init__()
        self.conv1d = nn.Conv1d(args.max_doc_len, args.nfilters, args.local_filter_size)
        self.drop = nn.Dropout(args.dropout)
        self.fc1 = nn.Linear(args.max_query_len, 1)
        self.fc2 = nn.Linear(args.nfilters, args.nfilters)
        self.fc3 = nn.Linear(args.nfilters, 1)


class DistributedModel(nn.Module):
    """neuroir/rankers/duet.py:124-146."""

    def __init__(self, args):
        super().__init__()
        self.conv_q = nn.Conv1d(args.emsize, args.nfilters, args.dist_filter_size)
        self.conv_d1 = nn.Conv1d(args.emsize, args.nfilters, args.dist_filter_size)
        self.conv_d2 = nn.Conv1d(args.nfilters, args.nfilters, 1)
        self.pool_size = args.pool_size
        self.dropout = nn.Dropout(args.dropout)
        self.fc1 = nn.Linear(args.nfilters, args.nfilters)
        self.fc2 = nn.Linear(args.max_doc_len - args.pool_size - 1, 1)
        self.fc3 = nn.Linear(args.nfilters, args.nfilters)
        self.fc4 = nn.Linear(args.nfilters, 1)


class DUET(_Ranker):
    """neuroir/rankers/duet.py:9-59."""
    MODEL = 'duet'

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.use_word = args.use_word
        if not self.use_word:
            raise TypeError('Non-word inputs are not supported!')  # duet.py:23
        self.word_embeddings = Embeddings(args.emsize, args.src_vocab_size, PAD)
        self.emb_drop = nn.Dropout(p=args.dropout_emb)
        self.local_model = LocalModel(args)
        self.distributed_model = DistributedModel(args)

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nfilters=a.nfilters,
                    local_filter_size=a.local_filter_size, dist_filter_size=a.dist_filter_size,
                    pool_size=a.pool_size, max_query_len=a.max_query_len, max_doc_len=a.max_doc_len)

    def _create(self, w, device, out):
        return lib.load().cair_duet_create(w, device, out)


RANKERS = {'ARCI': ARCI, 'ARCII': ARCII, 'DSSM': DSSM, 'CDSSM': CDSSM, 'ESM': ESM, 'MATCH_TENSOR': MatchTensor, 'DRMM': DRMM, 'DUET': DUET}
