"""Loader and ctypes prototypes of libcair.so (the C ABI declared in include/cair.h).

There is no CPU or PyTorch fallback: if the CUDA library is missing or a call fails, this
module raises."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcair.so')
_lib = None

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64

# name -> (restype, argtypes); every symbol include/cair.h declares
PROTOTYPES = {
    'cair_version': (i32, []),
    'cair_last_error': (C.c_char_p, []),
    'cair_launch_count': (i64, []),
    'cair_destroy': (i32, [vp]),
    'cair_poll_error': (i32, [vp, vp]),
    'cair_profile_enable': (i32, [vp, i32]),
    'cair_profile_read': (i32, [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_float), i32, C.POINTER(i32)]),
    'cair_embed_gather': (i32, [vp, i32, i32, vp, i64, vp, vp]),
    'cair_lstm_forward': (i32, [vp, vp, i32, i32, i32, i32, C.POINTER(_abi.LstmDir), C.POINTER(_abi.LstmDir),
                                vp, vp, vp, vp]),
    'cair_rnn_forward': (i32, [i32, vp, vp, i32, i32, i32, i32, C.POINTER(_abi.LstmDir), C.POINTER(_abi.LstmDir),
                               vp, vp, vp, vp]),
    'cair_set_rnn_impl': (i32, [i32]),
    'cair_umma_selftest': (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    'cair_umma_bench': (i32, [i32, i32, i32, i32, vp, vp]),
    'cair_esm_create': (i32, [C.POINTER(_abi.EsmWeights), i32, C.POINTER(vp)]),
    'cair_mt_create': (i32, [C.POINTER(_abi.MtWeights), i32, C.POINTER(vp)]),
    'cair_mt_set_debug': (i32, [vp, vp, vp]),
    'cair_mt_add_encoder_layer': (i32, [vp, i32, vp, vp]),
    'cair_mt_set_impl': (i32, [vp, i32]),
    'cair_allgather_scores': (i32, [vp, i64, vp, vp, i32, i32, C.c_uint32, vp]),
    'cair_ranker_set_gather': (i32, [vp, vp, vp, vp, i32, i32, i64]),
    'cair_set_gemm_impl': (i32, [i32]),
    'cair_set_drmm_impl': (i32, [i32]),
    'cair_drmm_create': (i32, [C.POINTER(_abi.DrmmWeights), i32, C.POINTER(vp)]),
    'cair_drmm_set_debug': (i32, [vp, vp]),
    'cair_dssm_create': (i32, [C.POINTER(_abi.DssmWeights), i32, C.POINTER(vp)]),
    'cair_cdssm_create': (i32, [C.POINTER(_abi.CdssmWeights), i32, C.POINTER(vp)]),
    'cair_arci_create': (i32, [C.POINTER(_abi.ArciWeights), i32, C.POINTER(vp)]),
    'cair_arcii_create': (i32, [C.POINTER(_abi.ArciiWeights), i32, C.POINTER(vp)]),
    'cair_duet_create': (i32, [C.POINTER(_abi.DuetWeights), i32, C.POINTER(vp)]),
    'cair_ranker_workspace_bytes': (i32, [vp, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_ranker_forward': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, vp, vp, C.c_size_t, vp]),
    'cair_ranker_forward_host': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    'cair_ranker_submit_host': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, vp]),
    'cair_ranker_wait_host': (i32, [vp, i32]),
    'cair_ranker_set_pipeline_split': (i32, [vp, C.c_float]),
    'cair_batchify_ranker': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    'cair_rank_metrics': (i32, [vp, vp, i32, i32, i32, vp, vp, vp]),
    'cair_cars_create': (i32, [C.POINTER(_abi.CarsWeights), i32, C.POINTER(vp)]),
    'cair_cars_workspace_bytes': (i32, [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_cars_forward': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32,
                                vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    'cair_cars_forward_ex': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, C.POINTER(_abi.CarsOutputs),
                                   vp, C.c_size_t, vp]),
    'cair_cars_set_decoder': (i32, [vp, C.POINTER(_abi.CarsDecoderWeights)]),
    'cair_cars_decode_workspace_bytes': (i32, [vp, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_mt_train_create': (i32, [C.POINTER(_abi.MtWeights), i32, C.POINTER(vp)]),
    'cair_mt_train_destroy': (i32, [vp]),
    'cair_mt_train_set_impl': (i32, [vp, i32]),
    'cair_mt_train_workspace_bytes': (i32, [vp, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_mt_train_forward': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, C.c_float, C.c_uint64, vp, vp, C.c_size_t, vp]),
    'cair_mt_train_backward': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, C.c_float, C.c_uint64, vp,
                                     C.POINTER(_abi.MtWeights), vp, C.c_size_t, vp]),
    'cair_mt_train_poll_error': (i32, [vp, vp, vp]),
    'cair_dropout_mask': (i32, [C.c_uint64, C.c_float, i64, vp, vp]),
    'cair_drmm_train_workspace_bytes': (i32, [i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_drmm_train_forward': (i32, [C.POINTER(_abi.DrmmWeights), vp, vp, i32, i32, i32, i32, C.c_float, C.c_uint64, vp, vp,
                                      C.c_size_t, vp]),
    'cair_drmm_train_backward': (i32, [C.POINTER(_abi.DrmmWeights), C.POINTER(_abi.DrmmWeights), vp, i32, i32, i32, i32, C.c_float,
                                       C.c_uint64, vp, vp, C.c_size_t, vp]),
    'cair_esm_train_workspace_bytes': (i32, [i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_esm_train_forward': (i32, [vp, i32, i32, vp, vp, i32, i32, i32, i32, vp, vp, C.c_size_t, vp]),
    'cair_esm_train_backward': (i32, [i32, i32, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, C.c_size_t, vp]),
    'cair_dssm_train_workspace_bytes': (i32, [i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_dssm_train_forward': (i32, [C.POINTER(_abi.DssmWeights), vp, vp, i32, i32, i32, i32, C.c_float, C.c_uint64, vp, vp,
                                      C.c_size_t, vp]),
    'cair_dssm_train_backward': (i32, [C.POINTER(_abi.DssmWeights), C.POINTER(_abi.DssmWeights), vp, vp, i32, i32, i32, i32,
                                       C.c_float, C.c_uint64, vp, vp, vp, C.c_size_t, vp]),
    'cair_cdssm_train_workspace_bytes': (i32, [i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_cdssm_train_forward': (i32, [C.POINTER(_abi.CdssmWeights), vp, vp, i32, i32, i32, i32, C.c_float, C.c_uint64, vp, vp,
                                       C.c_size_t, vp]),
    'cair_cdssm_train_backward': (i32, [C.POINTER(_abi.CdssmWeights), C.POINTER(_abi.CdssmWeights), vp, vp, i32, i32, i32, i32,
                                        C.c_float, C.c_uint64, vp, vp, vp, C.c_size_t, vp]),
    'cair_mnsrf_create': (i32, [C.POINTER(_abi.MnsrfWeights), i32, C.POINTER(vp)]),
    'cair_mnsrf_destroy': (i32, [vp]),
    'cair_mnsrf_workspace_bytes': (i32, [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_mnsrf_forward': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, C.c_size_t, vp]),
    'cair_mnsrf_poll_error': (i32, [vp, vp]),
    'cair_sessdec_create': (i32, [C.POINTER(_abi.SessDecWeights), i32, C.POINTER(vp)]),
    'cair_sessdec_destroy': (i32, [vp]),
    'cair_sessdec_workspace_bytes': (i32, [vp, i32, i32, C.POINTER(C.c_size_t)]),
    'cair_sessdec_states': (i32, [vp, vp, i32, i32, vp, vp, vp, C.c_size_t, vp]),
    'cair_sessdec_decode': (i32, [vp, vp, vp, i32, i32, i32, vp, i64, vp, vp, C.c_size_t, vp]),
    'cair_linear_maxpool': (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    'cair_cars_decode': (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, i64, vp, vp, C.c_size_t, vp]),
}


class CairError(RuntimeError):
    def __init__(self, code, text):
        super().__init__('%s: %s' % (_abi.ERR_NAMES.get(code, code), text))
        self.code = code


def load():
    """dlopen libcair.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               '(or `make -C context_attentive_ir_b200/csrc`); there is no fallback path' % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise CairError(rc, (load().cair_last_error() or b'').decode())


def launch_count():
    return int(load().cair_launch_count())
