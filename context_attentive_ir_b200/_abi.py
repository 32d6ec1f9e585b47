"""ctypes mirror of include/cair.h (struct layouts, status codes, weight packing).

The structs are addressed by the reference's state_dict key names, so the
same packing code serves device pointers (torch CUDA tensors -> libcair.so) and host pointers
(numpy arrays -> the CPU oracle used by the tests).
"""
import ctypes as C

CAIR_OK = 0
ERR_NAMES = {-1: 'CAIR_ERR_BAD_ARG', -2: 'CAIR_ERR_BAD_SHAPE', -3: 'CAIR_ERR_UNSUPPORTED',
             -4: 'CAIR_ERR_CUDA', -5: 'CAIR_ERR_WORKSPACE'}
RNN_TYPES = {'LSTM': 0, 'GRU': 1}

f32p = C.POINTER(C.c_float)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


class Linear(C.Structure):
    _fields_ = [('w', f32p), ('b', f32p)]


class LstmDir(C.Structure):
    _fields_ = [('w_ih', f32p), ('w_hh', f32p), ('b_ih', f32p), ('b_hh', f32p)]


class AttnMlp(C.Structure):
    _fields_ = [('l0', Linear), ('l3', Linear)]


class EsmWeights(C.Structure):
    _fields_ = [('vocab', C.c_int32), ('emsize', C.c_int32), ('table', f32p)]


class MtWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ('vocab', 'emsize', 'featsize', 'nhid_query', 'nhid_doc', 'nchannels', 'nfilters',
                 'match_filter_size', 'rnn_type', 'bidirectional')] + [
        ('table', f32p), ('linear_projection', Linear),
        ('query_fwd', LstmDir), ('query_rev', LstmDir), ('doc_fwd', LstmDir), ('doc_rev', LstmDir),
        ('query_projection', Linear), ('document_projection', Linear), ('alpha', f32p),
        ('conv1', Linear), ('conv2', Linear), ('conv3', Linear), ('conv', Linear), ('output', Linear)]


class MnsrfWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ('vocab', 'emsize', 'nhid_query', 'nhid_document', 'nhid_session', 'rnn_type',
                                          'bidirectional')] + [
        ('table', f32p), ('query_fwd', LstmDir), ('query_rev', LstmDir), ('doc_fwd', LstmDir), ('doc_rev', LstmDir),
        ('session', LstmDir), ('projection', Linear)]


class SessDecWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ('vocab', 'emsize', 'nhid_in', 'nhid_session', 'tgt_vocab')] + [
        ('table', f32p), ('session', LstmDir), ('dec_rnn', LstmDir), ('generator', Linear)]


class DrmmWeights(C.Structure):
    _fields_ = [('vocab', C.c_int32), ('emsize', C.c_int32), ('nbins', C.c_int32), ('table', f32p),
                ('gating', Linear), ('ffnn0', Linear), ('ffnn1', Linear), ('output', Linear)]


class DuetWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ('vocab', 'emsize', 'nfilters', 'local_filter_size', 'dist_filter_size', 'pool_size',
                 'max_query_len', 'max_doc_len')] + [('table', f32p)] + [
        (k, Linear) for k in ('local_conv1d', 'local_fc1', 'local_fc2', 'local_fc3', 'conv_q',
                              'conv_d1', 'conv_d2', 'dist_fc1', 'dist_fc2', 'dist_fc3', 'dist_fc4')]


class DssmWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ('vocab', 'emsize', 'nhid', 'nout')] + [('table', f32p)] + [
        (k, Linear) for k in ('query_mlp0', 'query_mlp2', 'doc_mlp0', 'doc_mlp2')]


class CdssmWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ('vocab', 'emsize', 'nhid', 'nout')] + [('table', f32p)] + [
        (k, Linear) for k in ('query_conv', 'query_sem', 'doc_conv', 'doc_sem')]


ARC_MAX_LAYERS = 4


class ArciWeights(C.Structure):
    _fields_ = [('vocab', C.c_int32), ('emsize', C.c_int32), ('nlayers', C.c_int32),
                ('filters', C.c_int32 * ARC_MAX_LAYERS), ('kernel', C.c_int32 * ARC_MAX_LAYERS),
                ('pool', C.c_int32 * ARC_MAX_LAYERS), ('max_query_len', C.c_int32), ('max_doc_len', C.c_int32),
                ('table', f32p), ('qconv', Linear * ARC_MAX_LAYERS), ('dconv', Linear * ARC_MAX_LAYERS),
                ('mlp0', Linear), ('mlp1', Linear)]


class ArciiWeights(C.Structure):
    _fields_ = [('vocab', C.c_int32), ('emsize', C.c_int32), ('filters_1d', C.c_int32), ('kernel_1d', C.c_int32),
                ('nlayers2d', C.c_int32), ('filters_2d', C.c_int32 * ARC_MAX_LAYERS), ('max_query_len', C.c_int32),
                ('max_doc_len', C.c_int32), ('table', f32p), ('conv_query', Linear), ('conv_doc', Linear),
                ('conv2d', Linear * ARC_MAX_LAYERS), ('mlp0', Linear), ('mlp1', Linear)]


class CarsWeights(C.Structure):
    _fields_ = [(k, C.c_int32) for k in
                ('vocab', 'emsize', 'nhid_query', 'nhid_document', 'nhid_session_query',
                 'nhid_session_document')] + [
        ('rank_dims', C.c_int32 * 3), ('rank_pool', C.c_int32), ('table', f32p),
        ('query_fwd', LstmDir), ('query_rev', LstmDir), ('doc_fwd', LstmDir), ('doc_rev', LstmDir),
        ('q_attn', AttnMlp), ('d_attn', AttnMlp), ('click_attn', AttnMlp),
        ('session_query_inner_attn', AttnMlp), ('session_doc_inner_attn', AttnMlp),
        ('session_query', LstmDir), ('session_doc', LstmDir),
        ('session_query_attn', Linear), ('session_doc_attn', Linear),
        ('shared_session_projector', Linear), ('private_session_projector1', Linear),
        ('q_projection', Linear), ('ranknet', Linear * 3)]


class CarsOutputs(C.Structure):
    _fields_ = [(k, f32p) for k in ('pooled_q', 'pooled_d', 'clicks', 'sess_q_attn', 'sess_d_attn', 'enc_q', 'sess_h', 'sess_c')]


class CarsDecoderWeights(C.Structure):
    _fields_ = [('nhid_decoder', C.c_int32), ('tgt_vocab', C.c_int32), ('transform_hid', Linear), ('transform_cell', Linear),
                ('rnn', LstmDir), ('attn_in', Linear), ('attn_out', Linear), ('dec_attn', Linear), ('predictor1', Linear),
                ('predictor2', Linear), ('private_session_projector2', Linear)]


TABLE_KEY = 'word_embeddings.make_embedding.emb_luts.0.weight'


def _lin(get, name, bias=True):
    return Linear(get(name + '.weight'), get(name + '.bias') if bias else None)


def _lstm(get, prefix, suffix=''):
    return LstmDir(get('%s.weight_ih_l0%s' % (prefix, suffix)), get('%s.weight_hh_l0%s' % (prefix, suffix)),
                   get('%s.bias_ih_l0%s' % (prefix, suffix)), get('%s.bias_hh_l0%s' % (prefix, suffix)))


def _attn(get, name):
    return AttnMlp(_lin(get, name + '.0'), _lin(get, name + '.3'))


def pack_esm(cfg, get):
    return EsmWeights(cfg['src_vocab_size'], cfg['emsize'], get(TABLE_KEY))


def pack_mt(cfg, get):
    """cfg: the reference Namespace fields (rankers/mtensor.py:27-60); get(key) -> float*."""
    bi = bool(cfg['bidirection'])
    w = MtWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['featsize'], cfg['nhid_query'],
                  cfg['nhid_doc'], cfg['nchannels'], cfg['nfilters'], cfg['match_filter_size'],
                  RNN_TYPES[cfg['rnn_type']], int(bi))
    w.table = get(TABLE_KEY)
    w.linear_projection = _lin(get, 'linear_projection')
    w.query_fwd = _lstm(get, 'query_encoder.rnns.0')
    w.doc_fwd = _lstm(get, 'document_encoder.rnns.0')
    if bi:
        w.query_rev = _lstm(get, 'query_encoder.rnns.0', '_reverse')
        w.doc_rev = _lstm(get, 'document_encoder.rnns.0', '_reverse')
    w.query_projection = _lin(get, 'query_projection')
    w.document_projection = _lin(get, 'document_projection')
    w.alpha = get('exact_match_channel.alpha')
    for k in ('conv1', 'conv2', 'conv3', 'conv', 'output'):
        setattr(w, k, _lin(get, k))
    return w


def pack_mmt(cfg, get):
    """M_MATCH_TENSOR (multitask/mmtensor.py:10-48): the Match-Tensor struct from the multitask key names
    (`embedder.` / `.encoder.` levels of multitask/layers.py, `nhid_document`)."""
    def mapped(key):
        if key == TABLE_KEY:
            return get('embedder.' + key)
        for enc in ('query_encoder.', 'document_encoder.'):
            if key.startswith(enc + 'rnns.'):
                return get(enc + 'encoder.' + key[len(enc):])
        return get(key)
    return pack_mt(dict(cfg, nhid_doc=cfg['nhid_document']), mapped)


def pack_mnsrf(cfg, get):
    """MNSRF (multitask/mnsrf.py:10-57)."""
    bi = bool(cfg['bidirection'])
    w = MnsrfWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nhid_query'], cfg['nhid_document'], cfg['nhid_session'],
                     RNN_TYPES[cfg['rnn_type']], int(bi))
    w.table = get('embedder.' + TABLE_KEY)
    w.query_fwd = _lstm(get, 'query_encoder.encoder.rnns.0')
    w.doc_fwd = _lstm(get, 'document_encoder.encoder.rnns.0')
    if bi:
        w.query_rev = _lstm(get, 'query_encoder.encoder.rnns.0', '_reverse')
        w.doc_rev = _lstm(get, 'document_encoder.encoder.rnns.0', '_reverse')
    w.session = _lstm(get, 'session_query_encoder.encoder.rnns.0')
    w.projection = _lin(get, 'projection.linear')
    return w


def pack_sessdec(cfg, get, with_session):
    """Decoder-side weights of MNSRF / M_MATCH_TENSOR (multitask/mnsrf.py:33-57)."""
    w = SessDecWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nhid_in'], cfg['nhid_session'], cfg['tgt_vocab_size'])
    w.table = get('embedder.' + TABLE_KEY)
    if with_session:
        w.session = _lstm(get, 'session_query_encoder.encoder.rnns.0')
    w.dec_rnn = LstmDir(get('decoder.decoder.rnn.weight_ih_l0'), get('decoder.decoder.rnn.weight_hh_l0'),
                        get('decoder.decoder.rnn.bias_ih_l0'), get('decoder.decoder.rnn.bias_hh_l0'))
    w.generator = _lin(get, 'generator')
    return w


def pack_drmm(cfg, get):
    w = DrmmWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nbins'], get(TABLE_KEY))
    w.gating = _lin(get, 'gating_network.weight')
    w.ffnn0 = _lin(get, 'ffnn.0')
    w.ffnn1 = _lin(get, 'ffnn.1')
    w.output = _lin(get, 'output')
    return w


def pack_duet(cfg, get):
    w = DuetWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nfilters'], cfg['local_filter_size'],
                    cfg['dist_filter_size'], cfg['pool_size'], cfg['max_query_len'], cfg['max_doc_len'],
                    get(TABLE_KEY))
    w.local_conv1d = _lin(get, 'local_model.conv1d')
    w.local_fc1 = _lin(get, 'local_model.fc1')
    w.local_fc2 = _lin(get, 'local_model.fc2')
    w.local_fc3 = _lin(get, 'local_model.fc3')
    for k in ('conv_q', 'conv_d1', 'conv_d2'):
        setattr(w, k, _lin(get, 'distributed_model.' + k))
    for i in (1, 2, 3, 4):
        setattr(w, 'dist_fc%d' % i, _lin(get, 'distributed_model.fc%d' % i))
    return w


def pack_dssm(cfg, get):
    w = DssmWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nhid'], cfg['nout'], get(TABLE_KEY))
    w.query_mlp0, w.query_mlp2 = _lin(get, 'query_mlp.0'), _lin(get, 'query_mlp.2')
    w.doc_mlp0, w.doc_mlp2 = _lin(get, 'doc_mlp.0'), _lin(get, 'doc_mlp.2')
    return w


def pack_cdssm(cfg, get):
    w = CdssmWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nhid'], cfg['nout'], get(TABLE_KEY))
    w.query_conv, w.query_sem = _lin(get, 'query_conv'), _lin(get, 'query_sem')
    w.doc_conv, w.doc_sem = _lin(get, 'doc_conv'), _lin(get, 'doc_sem')
    return w


def pack_arci(cfg, get):
    nl = len(cfg['filters_1d'])
    if nl > ARC_MAX_LAYERS:
        raise NotImplementedError('ARC-I with more than %d conv layers' % ARC_MAX_LAYERS)
    w = ArciWeights(cfg['src_vocab_size'], cfg['emsize'], nl)
    for i in range(nl):
        w.filters[i], w.kernel[i], w.pool[i] = cfg['filters_1d'][i], cfg['kernel_size_1d'][i], cfg['maxpool_size_1d'][i]
        w.qconv[i] = _lin(get, 'query_conv1d_layers.%d.0' % i)
        w.dconv[i] = _lin(get, 'doc_conv1d_layers.%d.0' % i)
    w.max_query_len, w.max_doc_len, w.table = cfg['max_query_len'], cfg['max_doc_len'], get(TABLE_KEY)
    w.mlp0, w.mlp1 = _lin(get, 'mlp.0'), _lin(get, 'mlp.1')
    return w


def pack_arcii(cfg, get):
    nl = len(cfg['filters_2d'])
    if nl > ARC_MAX_LAYERS:
        raise NotImplementedError('ARC-II with more than %d conv2d layers' % ARC_MAX_LAYERS)
    if any(list(k) != [3, 3] for k in cfg['kernel_size_2d']) or any(list(k) != [2, 2] for k in cfg['maxpool_size_2d']):
        raise NotImplementedError('libcair implements the stock 3x3 kernels / 2x2 max-pools (neuroir/hyparam.py:61-76)')
    w = ArciiWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['filters_1d'], cfg['kernel_size_1d'], nl)
    for i in range(nl):
        w.filters_2d[i] = cfg['filters_2d'][i]
        w.conv2d[i] = _lin(get, 'conv2d_layers.%d.0' % i)
    w.max_query_len, w.max_doc_len, w.table = cfg['max_query_len'], cfg['max_doc_len'], get(TABLE_KEY)
    w.conv_query, w.conv_doc = _lin(get, 'conv_query'), _lin(get, 'conv_doc')
    w.mlp0, w.mlp1 = _lin(get, 'mlp.0'), _lin(get, 'mlp.1')
    return w


def pack_cars(cfg, get):
    """Stock CARS ranking path (multitask/cars.py:28-131): LSTM, bidirectional, 1 layer, attn pooling."""
    w = CarsWeights(cfg['src_vocab_size'], cfg['emsize'], cfg['nhid_query'], cfg['nhid_document'],
                    cfg['nhid_session_query'], cfg['nhid_session_document'])
    w.rank_dims = (C.c_int32 * 3)(256, 128, 1)
    w.rank_pool = 2
    w.table = get('embedder.' + TABLE_KEY)
    w.query_fwd = _lstm(get, 'query_encoder.encoder.rnns.0')
    w.query_rev = _lstm(get, 'query_encoder.encoder.rnns.0', '_reverse')
    w.doc_fwd = _lstm(get, 'document_encoder.encoder.rnns.0')
    w.doc_rev = _lstm(get, 'document_encoder.encoder.rnns.0', '_reverse')
    for k in ('q_attn', 'd_attn', 'click_attn', 'session_query_inner_attn', 'session_doc_inner_attn'):
        setattr(w, k, _attn(get, k))
    w.session_query = _lstm(get, 'session_query_encoder.encoder.rnns.0')
    w.session_doc = _lstm(get, 'session_doc_encoder.encoder.rnns.0')
    w.session_query_attn = _lin(get, 'session_query_attn')
    w.session_doc_attn = _lin(get, 'session_doc_attn')
    w.shared_session_projector = _lin(get, 'shared_session_projector.linear', bias=False)
    w.private_session_projector1 = _lin(get, 'private_session_projector1.linear', bias=False)
    w.q_projection = _lin(get, 'q_projection.linear')
    for i in range(3):
        w.ranknet[i] = _lin(get, 'ranknet._linear_layers.%d' % i)
    return w


def pack_cars_decoder(cfg, get):
    """Decoder-side modules of CARS (multitask/cars.py:605-657), attn_type 'general'."""
    w = CarsDecoderWeights(cfg['nhid_decoder'], cfg['tgt_vocab_size'])
    w.transform_hid = _lin(get, 'transform_hid.linear')
    w.transform_cell = _lin(get, 'transform_cell.linear')
    w.rnn = _lstm(get, 'decoder.decoder.rnn')
    w.attn_in = _lin(get, 'decoder.decoder.attn.linear_in', bias=False)
    w.attn_out = _lin(get, 'decoder.decoder.attn.linear_out', bias=False)
    w.dec_attn = _lin(get, 'dec_attn', bias=False)
    w.predictor1 = _lin(get, 'token_prob_predictor1', bias=False)
    w.predictor2 = _lin(get, 'token_prob_predictor2', bias=False)
    w.private_session_projector2 = _lin(get, 'private_session_projector2.linear', bias=False)
    return w


PACKERS = {'arci': pack_arci, 'arcii': pack_arcii, 'dssm': pack_dssm, 'cdssm': pack_cdssm, 'esm': pack_esm, 'match_tensor': pack_mt, 'm_match_tensor': pack_mmt, 'mnsrf': pack_mnsrf, 'drmm': pack_drmm, 'duet': pack_duet,
           'cars': pack_cars}
