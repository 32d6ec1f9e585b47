"""Host-side mirror of the CARS ranking path (neuroir/multitask/cars.py, layers.py).

Keeps the reference's predict-time call sequence (neuroir/models/multitask.py:264-269):
    pooled, encoded, hidden = net.encode(source_words, source_lens)
    click_scores, states, session_attns = net.rank_document(pooled, document_words, document_lens, document_label)
and the reference's parameter names (extra `.encoder.` / `embedder.` levels from layers.py), so a
reference state_dict loads by key (strict=False: decoder-side keys are carried by the caller).
The suggestion decoder (cars.py:605-657, :706-791) is out of scope (scope table row f.3).
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _abi, lib
from .rankers import PAD, Embeddings, RNNEncoder, _CairModule


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
        self._last = None

    def _cfg(self):
        a = self.args
        return dict(src_vocab_size=a.src_vocab_size, emsize=a.emsize, nhid_query=a.nhid_query,
                    nhid_document=a.nhid_document, nhid_session_query=a.nhid_session_query,
                    nhid_session_document=a.nhid_session_document)

    def _create(self, w, device, out):
        return lib.load().cair_cars_create(w, device, out)

    def score(self, queries, query_len, docs, doc_len, doc_labels, session_slice=None, want_stages=False):
        """One fused call: q [B,S,Lq], qlen [B,S], d [B,S,N,Ld], dlen [B,S,N], labels [B,S,N] ->
        dict(scores [B,S,N], + pooled_queries, pooled_docs, clicks, sess_q_attn, sess_d_attn if want_stages)."""
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

        def p(k):
            return out[k].data_ptr() if k in out else None
        sb, sc = (0, B) if session_slice is None else session_slice
        lib.check(L.cair_cars_forward(h, q.data_ptr(), ql.data_ptr(), d.data_ptr(), dl.data_ptr(), lab.data_ptr(),
                                      B, S, N, Lq, Ld, sb, sc, out['scores'].data_ptr(), p('pooled_queries'),
                                      p('pooled_docs'), p('clicks'), p('sess_q_attn'), p('sess_d_attn'),
                                      ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        return out

    # -- the reference's two-call predict sequence (models/multitask.py:264-269) ------------------
    def encode(self, queries, query_length):
        """Defers the work: the fused kernel sequence runs in rank_document, which needs the documents.
        Returns (token, None, None); `token` stands in for pooled_rep and is only meaningful to
        rank_document (the decoder-side encoded_rep / hidden outputs are out of scope)."""
        self._last = (queries, query_length)
        return ('cair-deferred', id(self)), None, None

    def rank_document(self, pooled_rep, document_rep, document_len, document_label):
        if self._last is None:
            raise RuntimeError('rank_document() must follow encode() (models/multitask.py:264-269)')
        queries, qlen = self._last
        self._last = None
        out = self.score(queries, qlen, document_rep, document_len, document_label, want_stages=True)
        return out['scores'], None, (out['sess_q_attn'], out['sess_d_attn'])


# decoder-side parameters of the reference CARS (suggestion path, out of scope here)
DECODER_PREFIXES = ('decoder.', 'token_prob_predictor', 'dec_attn', 'transform_', 'private_session_projector2')


def _carry_load_state_dict(self, state_dict, strict=True, **kw):
    """nn.Module.load_state_dict for a reference checkpoint: the decoder-side tensors (suggestion path) are not parameters
    of this module; they are kept untouched and handed back by state_dict(), so a checkpoint loaded and saved through the
    reference's Multitask wrapper (models/multitask.py:49-58, 330-352) keeps every key."""
    carried = {k: v for k, v in state_dict.items() if k.startswith(DECODER_PREFIXES)}
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


MULTITASK = {'CARS': CARS}
