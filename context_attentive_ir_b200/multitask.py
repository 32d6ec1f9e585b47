"""Host-side mirror of CARS (neuroir/multitask/cars.py, layers.py): ranking path and greedy suggestion decoder.

Keeps the reference's predict-time call sequence (neuroir/models/multitask.py:264-292):
    pooled, encoded, hidden = net.encode(source_words, source_lens)
    click_scores, states, session_attns = net.rank_document(pooled, document_words, document_lens, document_label)
    out = net.decode(states=..., max_len=..., src_dict=..., tgt_dict=..., batch_size=..., session_len=..., use_cuda=...,
                     encoded_source=..., source_len=..., session_attns=...)          # {'predictions': [B, S-1, max_len]}
and the reference's parameter names (extra `.encoder.` / `embedder.` levels from layers.py; `decoder.decoder.rnn.*`,
`decoder.decoder.attn.linear_{in,out}`), so a reference state_dict loads by key.  The arithmetic is in libcair.so
(cair_cars_forward_ex, cair_cars_decode); there is no PyTorch fallback.
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _abi, lib
from .rankers import PAD, Embeddings, ExactMatchChannel, RNNEncoder, _CairModule, _Ranker

BOS = 2   # neuroir/inputters/constants.py:3


class Embedder(nn.Module):
    """neuroir/multitask/layers.py:10-27."""

    def __init__(self, emsize, src_vocab_size, dropout_emb):
        super().__init__()
        self.word_embeddings = Embeddings(emsize, src_vocab_size, PAD)
        self.output_size = emsize
        self.dropout = nn.Dropout(dropout_emb)


class Encoder(nn.Module):
    """neuroir/multitask/layers.py:30-54."""

    def __init__(self, rnn_type, input_size, bidirection, nlayers, nhid, dropout_rnn):
        super().__init__()
        self.encoder = RNNEncoder(rnn_type, input_size, bidirection, nlayers, nhid, dropout_rnn)


def _attn_mlp(h, dropout):
    return nn.Sequential(nn.Linear(h, h), nn.Tanh(), nn.Dropout(p=dropout), nn.Linear(h, 1))


def _proj(i, o, dropout, bias):
    return nn.Sequential(OrderedDict([('dropout', nn.Dropout(p=dropout)), ('linear', nn.Linear(i, o, bias=bias))]))


class Maxout(nn.Module):
    """Parameter container for neuroir/modules/maxout.py:29-68 (keys `_linear_layers.<i>.*`)."""

    def __init__(self, input_dim, num_layers, output_dims, pool_sizes):
        super().__init__()
        dims = [input_dim] + output_dims[:-1]
        self._linear_layers = nn.ModuleList([nn.Linear(i, o * p) for i, o, p in zip(dims, output_dims, pool_sizes)])
        self._output_dims, self._pool_sizes = output_dims, pool_sizes


class _GeneralAttention(nn.Module):
    """Parameters of GlobalAttention(attn_type='general') (modules/global_attention.py:59-96)."""

    def __init__(self, dim):
        super().__init__()
        self.linear_in = nn.Linear(dim, dim, bias=False)
        self.linear_out = nn.Linear(dim * 2, dim, bias=False)


class _RNNDecoder(nn.Module):
    """Parameters of RNNDecoder (decoders/decoder.py:68-104): one LSTM layer + attention."""

    def __init__(self, input_size, hidden_size, attn=True, rnn_type='LSTM'):
        super().__init__()
        self.rnn = getattr(nn, rnn_type)(input_size=input_size, hidden_size=hidden_size, num_layers=1, batch_first=True)
        if attn:   # attn_type 'none' (MNSRF, M_MATCH_TENSOR): the reference decoder has no attention parameters
            self.attn = _GeneralAttention(hidden_size)


class Decoder(nn.Module):
    """neuroir/multitask/layers.py:57-93 (adds the `decoder.` level to the keys)."""

    def __init__(self, input_size, nhid, attn=True, rnn_type='LSTM'):
        super().__init__()
        self.decoder = _RNNDecoder(input_size, nhid, attn, rnn_type)


class _DecoderStates:
    """What rank_document hands to decode() as `states`: the session encoders' (h, c) after every query; the
    transform_hid / transform_cell projections of cars.py:440-453 are applied inside cair_cars_decode."""

    def __init__(self, sess_h, sess_c):
        self.sess_h, self.sess_c = sess_h, sess_c


class CARS(_CairModule):
    """Ranking half of neuroir/multitask/cars.py (stock configuration: LSTM, bidirectional, one layer,
    attention pooling, both session encoders on - neuroir/hyparam.py:197-225)."""
    MODEL = 'cars'

    def __init__(self, args):
        super().__init__()
        self.args = args
        if args.rnn_type != 'LSTM' or not args.bidirection or args.nlayers != 1 or args.pool_type != 'attn' \
                or args.query_session_off or args.doc_session_off or args.turn_ranker_off:
            raise NotImplementedError('libcair implements the stock CARS ranking configuration only')
        self.embedder = Embedder(args.emsize, args.src_vocab_size, args.dropout_emb)
        self.query_encoder = Encoder(args.rnn_type, args.emsize, True, 1, args.nhid_query, args.dropout_rnn)
        self.document_encoder = Encoder(args.rnn_type, args.emsize, True, 1, args.nhid_document, args.dropout_rnn)
        self.q_attn = _attn_mlp(args.nhid_query, args.dropout)
        self.d_attn = _attn_mlp(args.nhid_document, args.dropout)
        self.nhid_session_query = args.nhid_session_query
        self.session_query_encoder = Encoder(args.rnn_type, args.nhid_query, False, 1, args.nhid_session_query,
                                             args.dropout_rnn)
        self.session_query_attn = nn.Linear(args.nhid_session_query, args.nhid_query)
        self.session_query_inner_attn = _attn_mlp(args.nhid_session_query, args.dropout)
        self.click_attn = _attn_mlp(args.nhid_document, args.dropout)
        self.nhid_session_document = args.nhid_session_document
        self.session_doc_encoder = Encoder(args.rnn_type, args.nhid_document, False, 1, args.nhid_session_document,
                                           args.dropout_rnn)
        self.session_doc_attn = nn.Linear(args.nhid_session_document, args.nhid_document)
        self.session_doc_inner_attn = _attn_mlp(args.nhid_session_document, args.dropout)
        srs = args.nhid_session_query + args.nhid_session_document
        self.shared_session_projector = _proj(srs, args.nhid_document, args.dropout, False)
        self.q_projection = _proj(args.nhid_query, args.nhid_document, args.dropout, True)
        self.private_session_projector1 = _proj(srs, args.nhid_document, args.dropout, False)
        self.ranknet = Maxout(args.nhid_document * 4, 3, [256, 128, 1], [2, 2, 2])
        # decoder-side modules (cars.py:605-657): parameter containers, the decode loop runs in cair_cars_decode
        self.has_decoder = not getattr(args, 'turn_recommender_off', False)
        if self.has_decoder:
            if getattr(args, 'attn_type', 'general') != 'general':
                raise NotImplementedError('libcair implements the stock CARS decoder attention (attn_type general)')
            self.private_session_projector2 = _proj(srs, args.nhid_document, args.dropout, False)
            self.transform_hid = _proj(srs, args.nhid_decoder, args.dropout, True)
            self.transform_cell = _proj(srs, args.nhid_decoder, args.dropout, True)
            self.decoder = Decoder(args.emsize, args.nhid_decoder)
            self.dec_attn = nn.Linear(args.nhid_query, args.nhid_decoder, bias=False)
            self.token_prob_predictor1 = nn.Linear(args.nhid_decoder, args.nhid_document, bias=False)
            self.token_prob_predictor2 = nn.Linear(args.nhid_document, args.tgt_vocab_size, bias=False)
        self._last = None
        self._fwd = None

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nhid_query=a.nhid_query,
                    nhid_document=a.nhid_document, nhid_session_query=a.nhid_session_query,
                    nhid_session_document=a.nhid_session_document)

    def _create(self, w, device, out):
        return lib.load().cair_cars_create(w, device, out)

    def _on_handle_created(self, handle):
        if self.has_decoder:
            keep = []
            a = self.args
            from .rankers import _ptr_getter
            w = _abi.pack_cars_decoder(dict(nhid_decoder=a.nhid_decoder, tgt_vocab_size=a.tgt_vocab_size), _ptr_getter(self, keep))
            lib.check(lib.load().cair_cars_set_decoder(handle, C.byref(w)))

    def score(self, queries, query_len, docs, doc_len, doc_labels, session_slice=None, want_stages=False, want_decoder_inputs=False):
        """One fused call: q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N], labels [B,S,N] ->
        dict(scores [B,S,N], + pooled_queries, pooled_docs, clicks, sess_q_attn, sess_d_attn if want_stages,
        + enc_q [B*S,Lq,Hq], sess_h / sess_c [B,S,Hsq+Hsd] if want_decoder_inputs)."""
        q = self._ids(queries, 'queries')
        d = self._ids(docs, 'docs')
        dev = q.device
        ql = self._ids(query_len, 'query_len').to(dev)
        dl = self._ids(doc_len, 'doc_len').to(dev)
        lab = doc_labels.to(device=dev, dtype=torch.float32).contiguous()
        B, S, Lq = q.shape
        N, Ld = d.shape[2], d.shape[3]
        a = self.args
        L = lib.load()
        h = self._handle_for(dev)
        nbytes = C.c_size_t()
        lib.check(L.cair_cars_workspace_bytes(h, B, S, N, Lq, Ld, C.byref(nbytes)))
        ws = self._workspace(nbytes.value, dev)
        out = dict(scores=torch.zeros(B, S, N, device=dev))
        if want_stages:
            out.update(pooled_queries=torch.zeros(B, S, a.nhid_query, device=dev),
                       pooled_docs=torch.zeros(B, S, N, a.nhid_document, device=dev),
                       clicks=torch.zeros(B, S, a.nhid_document, device=dev),
                       sess_q_attn=torch.zeros(B, S, a.nhid_session_query, device=dev),
                       sess_d_attn=torch.zeros(B, S, a.nhid_session_document, device=dev))

        if want_decoder_inputs:
            hs = a.nhid_session_query + a.nhid_session_document
            out.update(enc_q=torch.zeros(B * S, Lq, a.nhid_query, device=dev), sess_h=torch.zeros(B, S, hs, device=dev),
                       sess_c=torch.zeros(B, S, hs, device=dev))

        def p(k):
            return C.cast(out[k].data_ptr(), _abi.f32p) if k in out else None
        sb, sc = (0, B) if session_slice is None else session_slice
        outs = _abi.CarsOutputs(p('pooled_queries'), p('pooled_docs'), p('clicks'), p('sess_q_attn'), p('sess_d_attn'),
                                p('enc_q'), p('sess_h'), p('sess_c'))
        lib.check(L.cair_cars_forward_ex(h, q.data_ptr(), ql.data_ptr(), d.data_ptr(), dl.data_ptr(), lab.data_ptr(),
                                         B, S, N, Lq, Ld, sb, sc, out['scores'].data_ptr(), C.byref(outs),
                                         ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return out

    def forward(self, *args, **kwargs):
        raise NotImplementedError(_NO_TRAINING % 'CARS')

    # -- the reference's two-call predict sequence (models/multitask.py:264-269) ------------------
    def encode(self, queries, query_length):
        """Defers the work: the fused kernel sequence runs in rank_document, which needs the documents.
        Returns (token, token, None): the tokens stand in for pooled_rep / encoded_source and are resolved by
        rank_document() / decode() (the query memory banks are an output of the same fused forward)."""
        self._last = (queries, query_length)
        tok = ('cair-deferred', id(self))
        return tok, tok, None

    def rank_document(self, pooled_rep, document_rep, document_len, document_label):
        if self._last is None:
            raise RuntimeError('rank_document() must follow encode() (models/multitask.py:264-269)')
        queries, qlen = self._last
        self._last = None
        out = self.score(queries, qlen, document_rep, document_len, document_label, want_stages=True,
                         want_decoder_inputs=self.has_decoder)
        self._fwd = out
        states = _DecoderStates(out['sess_h'], out['sess_c']) if self.has_decoder else None
        return out['scores'], states, (out['sess_q_attn'], out['sess_d_attn'])

    def decode(self, states, max_len, src_dict, tgt_dict, batch_size, session_len, use_cuda=True, encoded_source=None,
               source_len=None, session_attns=None, **_):
        """Greedy suggestion decode (cars.py:706-791): `session_len` is the number of decoded queries per session (S - 1);
        returns {'predictions': LongTensor [batch_size, session_len, max_len]} of target-vocabulary ids."""
        if not self.has_decoder:
            raise RuntimeError('this CARS was built with turn_recommender_off: there is no decoder')
        fwd = self._fwd
        if fwd is None or not isinstance(states, _DecoderStates):
            raise RuntimeError('decode() must follow rank_document() (models/multitask.py:264-292)')
        enc_q = fwd['enc_q'] if not torch.is_tensor(encoded_source) else encoded_source.contiguous()
        dev = enc_q.device
        B, S = states.sess_h.shape[0], states.sess_h.shape[1]
        assert batch_size == B and session_len == S - 1
        Lq = enc_q.shape[1]
        qlen = self._ids(source_len, 'source_len').to(dev).reshape(B, S)
        sqa, sda = session_attns
        L = lib.load()
        h = self._handle_for(dev)
        nbytes = C.c_size_t()
        lib.check(L.cair_cars_decode_workspace_bytes(h, B, S, Lq, C.byref(nbytes)))
        ws = self.__dict__.get('_cair_dec_ws')
        if ws is None or ws.numel() < nbytes.value or ws.device != dev:
            ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
            self.__dict__['_cair_dec_ws'] = ws
        preds = torch.zeros(B, S - 1, max_len, dtype=torch.int64, device=dev)
        t2s = _tgt2src_table(self, tgt_dict, src_dict, dev, self.args.tgt_vocab_size)
        lib.check(L.cair_cars_decode(h, enc_q.data_ptr(), qlen.data_ptr(), states.sess_h.data_ptr(), states.sess_c.data_ptr(),
                                     sqa.contiguous().data_ptr(), sda.contiguous().data_ptr(), B, S, Lq, max_len, t2s.data_ptr(),
                                     BOS, preds.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return {'predictions': preds}



# decoder-side parameters of the reference CARS: real parameters of this module unless it was built with
# turn_recommender_off, in which case a checkpoint that still holds them is carried through untouched
DECODER_PREFIXES = ('decoder.', 'token_prob_predictor', 'dec_attn', 'transform_', 'private_session_projector2')


def _carry_load_state_dict(self, state_dict, strict=True, **kw):
    """nn.Module.load_state_dict for a reference checkpoint: decoder-side tensors this module has no parameter for are
    kept untouched and handed back by state_dict(), so a checkpoint loaded and saved through the reference's Multitask
    wrapper (models/multitask.py:49-58, 330-352) keeps every key."""
    own = set(nn.Module.state_dict(self).keys())
    carried = {k: v for k, v in state_dict.items() if k.startswith(DECODER_PREFIXES) and k not in own}
    mine = {k: v for k, v in state_dict.items() if k not in carried}
    out = nn.Module.load_state_dict(self, mine, strict=strict, **kw)
    self.__dict__['_carried_decoder_state'] = carried
    return out


def _carry_state_dict(self, *args, **kw):
    sd = nn.Module.state_dict(self, *args, **kw)
    prefix = kw.get('prefix', args[1] if len(args) > 1 else '')
    for k, v in self.__dict__.get('_carried_decoder_state', {}).items():
        sd[prefix + k] = v
    return sd


CARS.load_state_dict = _carry_load_state_dict
CARS.state_dict = _carry_state_dict


def ranking_state_dict(reference_state_dict):
    """Filters a reference CARS state_dict down to the ranking-path keys this module owns."""
    return {k: v for k, v in reference_state_dict.items() if not k.startswith(DECODER_PREFIXES)}


def _tgt2src_table(module, tgt_dict, src_dict, device, tgt_vocab_size):
    """Target-vocabulary id -> source-vocabulary id of the same word (the per-step tgt_dict[idx] -> src_dict[word] round trip
    of the reference decode loops) as one cached lookup table."""
    key = (id(tgt_dict), id(src_dict), len(tgt_dict), str(device))
    cache = module.__dict__.get('_tgt2src_cache')
    if cache is None or cache[0] != key:
        m = torch.tensor([int(src_dict[tgt_dict[i]]) for i in range(len(tgt_dict))], dtype=torch.int64)
        if m.numel() < tgt_vocab_size:
            m = torch.cat([m, torch.full((tgt_vocab_size - m.numel(),), 1, dtype=torch.int64)])   # UNK
        cache = (key, m.to(device))
        module.__dict__['_tgt2src_cache'] = cache
    return cache[1]


_NO_TRAINING = ('%s.forward() is the training entry point of the reference (ranking + suggestion losses, '
                'models/multitask.py:161-223): the libcair training step exists for the stand-alone MatchTensor, DRMM, ESM, DSSM and CDSSM rankers '
                'only.  Scoring and suggestion decoding run through encode() / rank_document() / decode().')


class _SessionDecoderMixin:
    """decode() of MNSRF / M_MATCH_TENSOR (mnsrf.py:258-300, mmtensor.py:258-300) on cair_sessdec_*: an attention-free LSTM
    decoder + generator, greedy from BOS.  The native decoder object follows the parameters like the scoring handle."""

    def _sessdec_for(self, device, nhid_in, with_session):
        from .rankers import _ptr_getter
        key = (self._state_key(), device.index)
        h = self.__dict__.get('_cair_sessdec')
        if h is not None and self.__dict__.get('_cair_sessdec_key') == key:
            return h
        self._release_sessdec()
        a = self.args
        keep = []
        w = _abi.pack_sessdec(dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nhid_in=nhid_in, nhid_session=a.nhid_session,
                                   tgt_vocab_size=a.tgt_vocab_size), _ptr_getter(self, keep), with_session)
        out = C.c_void_p()
        torch.cuda.synchronize(device)
        lib.check(lib.load().cair_sessdec_create(C.byref(w), device.index, C.byref(out)))
        self.__dict__['_cair_sessdec'], self.__dict__['_cair_sessdec_key'] = out, key
        return out

    def _release_sessdec(self):
        h = self.__dict__.get('_cair_sessdec')
        if h is not None:
            lib.load().cair_sessdec_destroy(h)
            self.__dict__['_cair_sessdec'] = None

    def _greedy(self, h, sess_h, sess_c, max_len, src_dict, tgt_dict):
        dev = sess_h.device
        B, S = sess_h.shape[0], sess_h.shape[1]
        L = lib.load()
        nbytes = C.c_size_t()
        lib.check(L.cair_sessdec_workspace_bytes(h, B, S, C.byref(nbytes)))
        ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        preds = torch.zeros(B, max(S - 1, 0), max_len, dtype=torch.int64, device=dev)
        t2s = _tgt2src_table(self, tgt_dict, src_dict, dev, self.args.tgt_vocab_size)
        lib.check(L.cair_sessdec_decode(h, sess_h.data_ptr(), sess_c.data_ptr(), B, S, max_len, t2s.data_ptr(), BOS, preds.data_ptr(),
                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return {'predictions': preds}


class MNSRF(_SessionDecoderMixin, _CairModule):
    """Ranking half of neuroir/multitask/mnsrf.py:10-162 (encode + rank_document); decoder-side parameters (decoder.*,
    generator.*) feed the greedy suggestion decoder (decode(), cair_sessdec_decode)."""
    MODEL = 'mnsrf'

    def __init__(self, args):
        super().__init__()
        self.args = args
        if args.nlayers != 1:
            raise NotImplementedError('libcair implements single-layer encoders')
        if args.rnn_type != 'LSTM':   # the reference's own session loop raises for GRU (`if init_states:`, rnn_encoder.py:77)
            raise NotImplementedError('MNSRF: rnn_type LSTM only (the reference cannot run its session loop with GRU either)')
        self.embedder = Embedder(args.emsize, args.src_vocab_size, args.dropout_emb)
        self.query_encoder = Encoder(args.rnn_type, args.emsize, args.bidirection, 1, args.nhid_query, args.dropout_rnn)
        self.document_encoder = Encoder(args.rnn_type, args.emsize, args.bidirection, 1, args.nhid_document, args.dropout_rnn)
        self.nhid_session = args.nhid_session
        self.session_query_encoder = Encoder(args.rnn_type, args.nhid_query, False, 1, args.nhid_session, args.dropout_rnn)
        self.decoder = Decoder(args.emsize, args.nhid_session, attn=False, rnn_type=args.rnn_type)
        self.projection = nn.Sequential(OrderedDict([('linear', nn.Linear(args.nhid_query + args.nhid_session, args.nhid_document)),
                                                     ('tanh', nn.Tanh())]))
        self.dropout = nn.Dropout(args.dropout)
        self.generator = nn.Linear(args.nhid_session, args.tgt_vocab_size)
        self._last = None

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nhid_query=a.nhid_query, nhid_document=a.nhid_document,
                    nhid_session=a.nhid_session, rnn_type=a.rnn_type, bidirection=a.bidirection)

    def _create(self, w, device, out):
        return lib.load().cair_mnsrf_create(w, device, out)

    def _release(self):
        h = self.__dict__.get('_cair_handle')
        if h is not None:
            lib.load().cair_mnsrf_destroy(h)
            self.__dict__['_cair_handle'] = None
        self._release_sessdec()

    def score(self, queries, query_len, docs, doc_len, session_slice=None, want_banks=False):
        """q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N] -> dict(scores [B,S,N] + memory_bank [B,S,Hq], session_bank,
        session_cell [B,S,Hs] if want_banks); session_slice=(begin, count) scores only those sessions."""
        q = self._ids(queries, 'queries')
        d = self._ids(docs, 'docs')
        dev = q.device
        ql = self._ids(query_len, 'query_len').to(dev)
        dl = self._ids(doc_len, 'doc_len').to(dev)
        B, S, Lq = q.shape
        N, Ld = d.shape[2], d.shape[3]
        a = self.args
        L = lib.load()
        h = self._handle_for(dev)
        nbytes = C.c_size_t()
        lib.check(L.cair_mnsrf_workspace_bytes(h, B, S, N, Lq, Ld, C.byref(nbytes)))
        ws = self._workspace(nbytes.value, dev)
        out = dict(scores=torch.zeros(B, S, N, device=dev))
        if want_banks:
            out.update(memory_bank=torch.zeros(B, S, a.nhid_query, device=dev), session_bank=torch.zeros(B, S, a.nhid_session, device=dev),
                       session_cell=torch.zeros(B, S, a.nhid_session, device=dev))

        def p(k):
            return out[k].data_ptr() if k in out else None
        sb, sc = (0, B) if session_slice is None else session_slice
        lib.check(L.cair_mnsrf_forward(h, q.data_ptr(), ql.data_ptr(), d.data_ptr(), dl.data_ptr(), B, S, N, Lq, Ld, sb, sc,
                                       out['scores'].data_ptr(), p('memory_bank'), p('session_bank'), p('session_cell'),
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return out

    def poll_error(self):
        h = self.__dict__.get('_cair_handle')
        if h is not None:
            lib.check(lib.load().cair_mnsrf_poll_error(h, torch.cuda.current_stream().cuda_stream))

    def forward(self, *args, **kwargs):
        raise NotImplementedError(_NO_TRAINING % 'MNSRF')

    # -- the reference's predict-time call sequence (models/multitask.py:270-276) --
    def encode(self, source_rep, source_len):
        """Deferred like CARS.encode: the fused forward runs in rank_document, which has the documents."""
        self._last = (source_rep, source_len)
        tok = ('cair-deferred', id(self))
        return tok, tok, None

    def rank_document(self, source_rep, memory_bank, session_bank, document_rep, document_len):
        if self._last is None:
            raise RuntimeError('rank_document() must follow encode() (models/multitask.py:270-276)')
        queries, qlen = self._last
        self._last = None
        out = self.score(queries, qlen, document_rep, document_len, want_banks=True)
        self.__dict__['_fwd'] = out
        return out['scores']

    def decode(self, states, max_len, src_dict, tgt_dict, batch_size, session_len, use_cuda=True, **_):
        """Greedy suggestion decode (mnsrf.py:258-300) from the session states of the last rank_document();
        returns {'predictions': LongTensor [batch_size, session_len, max_len]} (session_len = S - 1)."""
        fwd = self.__dict__.get('_fwd')
        if fwd is None:
            raise RuntimeError('decode() must follow rank_document() (models/multitask.py:270-292)')
        sess_h, sess_c = fwd['session_bank'], fwd['session_cell']
        assert batch_size == sess_h.shape[0] and session_len == sess_h.shape[1] - 1
        h = self._sessdec_for(sess_h.device, self.args.nhid_query, with_session=False)
        return self._greedy(h, sess_h, sess_c, max_len, src_dict, tgt_dict)


class M_MATCH_TENSOR(_SessionDecoderMixin, _Ranker):
    """Ranking half of neuroir/multitask/mmtensor.py:10-189.  rank_document never looks at the session: it is Match-Tensor on
    every (query, candidate) of every session, so it runs on a Match-Tensor handle with B*S queries; the session encoder,
    decoder and generator serve decode() (cair_linear_maxpool, cair_sessdec_states, cair_sessdec_decode)."""
    MODEL = 'm_match_tensor'

    def __init__(self, args):
        super().__init__()
        self.args = args
        if args.nlayers != 1:
            raise NotImplementedError('libcair implements single-layer encoders')
        self.embedder = Embedder(args.emsize, args.src_vocab_size, args.dropout_emb)
        self.linear_projection = nn.Linear(args.emsize, args.featsize)
        self.query_encoder = Encoder(args.rnn_type, args.featsize, args.bidirection, 1, args.nhid_query, args.dropout_rnn)
        self.document_encoder = Encoder(args.rnn_type, args.featsize, args.bidirection, 1, args.nhid_document, args.dropout_rnn)
        self.query_projection = nn.Linear(args.nhid_query, args.nchannels)
        self.document_projection = nn.Linear(args.nhid_document, args.nchannels)
        self.exact_match_channel = ExactMatchChannel()
        self.conv1 = nn.Conv2d(args.nchannels + 1, args.nfilters, (3, 3), padding=1)
        self.conv2 = nn.Conv2d(args.nchannels + 1, args.nfilters, (3, 5), padding=(1, 2))
        self.conv3 = nn.Conv2d(args.nchannels + 1, args.nfilters, (3, 7), padding=(1, 3))
        self.conv = nn.Conv2d(args.nfilters * 3, args.match_filter_size, (1, 1))
        self.output = nn.Linear(args.match_filter_size, 1)
        self.nhid_session = args.nhid_session
        self.session_query_encoder = Encoder(args.rnn_type, args.nchannels, False, 1, args.nhid_session, args.dropout_rnn)
        self.decoder = Decoder(args.emsize, args.nhid_session, attn=False, rnn_type=args.rnn_type)
        self.dropout = nn.Dropout(args.dropout)
        self.generator = nn.Linear(args.nhid_session, args.tgt_vocab_size)
        self._last = None

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, featsize=a.featsize, nhid_query=a.nhid_query,
                    nhid_document=a.nhid_document, nchannels=a.nchannels, nfilters=a.nfilters,
                    match_filter_size=a.match_filter_size, rnn_type=a.rnn_type, bidirection=a.bidirection)

    def _create(self, w, device, out):
        return lib.load().cair_mt_create(w, device, out)

    def _on_handle_created(self, handle):
        impl = self.__dict__.get('_cair_impl')   # 0: fp32 kernels, 1: tcgen05 (default), as MatchTensor.set_interaction_impl
        if impl is not None:
            lib.check(lib.load().cair_mt_set_impl(handle, impl))

    def score(self, queries, query_len, docs, doc_len):
        """q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N] -> scores [B,S,N]."""
        B, S, Lq = queries.shape
        N, Ld = docs.shape[2], docs.shape[3]
        s = self.forward(queries.reshape(B * S, Lq), query_len.reshape(B * S), docs.reshape(B * S, N, Ld),
                         doc_len.reshape(B * S, N))
        return s.view(B, S, N)

    def encode(self, source_rep, source_len):
        self._last = (source_rep, source_len)
        tok = ('cair-deferred', id(self))
        return tok, tok, None

    def rank_document(self, source_rep, projected_queries, session_bank, document_rep, document_len):
        if self._last is None:
            raise RuntimeError('rank_document() must follow encode() (models/multitask.py:270-276)')
        queries, qlen = self._last
        self._last = None
        # keep the query memory banks of this forward: decode() derives the session states from them
        B, S, Lq = queries.shape
        dev = document_rep.device
        enc_q = torch.zeros(B * S, Lq, self.args.nhid_query, device=dev)
        h = self._handle_for(dev)
        lib.check(lib.load().cair_mt_set_debug(h, enc_q.data_ptr(), None))
        try:
            scores = self.score(queries, qlen, document_rep, document_len)
        finally:
            lib.check(lib.load().cair_mt_set_debug(h, None, None))
        self.__dict__['_fwd'] = dict(enc_q=enc_q, B=B, S=S)
        return scores

    def _release(self):
        super()._release()
        self._release_sessdec()

    def decode(self, states, max_len, src_dict, tgt_dict, batch_size, session_len, use_cuda=True, **_):
        """Greedy suggestion decode (mmtensor.py:258-300): the session encoder runs over the max-pooled projected queries of the
        last rank_document() (mmtensor.py:86-116), its (h, c) after every query start the decoder."""
        fwd = self.__dict__.get('_fwd')
        if fwd is None:
            raise RuntimeError('decode() must follow rank_document() (models/multitask.py:270-292)')
        enc_q, B, S = fwd['enc_q'], fwd['B'], fwd['S']
        assert batch_size == B and session_len == S - 1
        dev = enc_q.device
        a = self.args
        L = lib.load()
        stream = torch.cuda.current_stream(dev).cuda_stream
        Lq = enc_q.shape[1]
        pooled = torch.empty(B * S, a.nchannels, device=dev)
        scratch = torch.empty(B * S * Lq * a.nchannels, device=dev)
        wq, bq = self.query_projection.weight.detach().contiguous(), self.query_projection.bias.detach().contiguous()
        lib.check(L.cair_linear_maxpool(enc_q.data_ptr(), wq.data_ptr(), bq.data_ptr(), B * S, Lq, a.nhid_query, a.nchannels,
                                        pooled.data_ptr(), scratch.data_ptr(), stream))
        h = self._sessdec_for(dev, a.nchannels, with_session=True)
        nbytes = C.c_size_t()
        lib.check(L.cair_sessdec_workspace_bytes(h, B, S, C.byref(nbytes)))
        ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        sess_h = torch.empty(B, S, a.nhid_session, device=dev)
        sess_c = torch.empty(B, S, a.nhid_session, device=dev)
        lib.check(L.cair_sessdec_states(h, pooled.data_ptr(), B, S, sess_h.data_ptr(), sess_c.data_ptr(), ws.data_ptr(), ws.numel(),
                                        stream))
        self.__dict__['_fwd'].update(session_bank=sess_h, session_cell=sess_c)
        return self._greedy(h, sess_h, sess_c, max_len, src_dict, tgt_dict)


MULTITASK = {'CARS': CARS, 'MNSRF': MNSRF, 'M_MATCH_TENSOR': M_MATCH_TENSOR}
