"""ctypes binding of libr2n2_b200.so (the C ABI declared in include/r2n2_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libr2n2_b200.so')
if os.environ.get('R2N2_LIB_PATH'):      # diagnostic A/B runs against another build of the library (never the product's default)
    LIB_PATH = os.environ['R2N2_LIB_PATH']
elif os.environ.get('R2N2_WAITSTAT'):      # diagnostic build with per-warp mbarrier wait statistics (tools/waitstat.py)
    LIB_PATH = os.path.join(_HERE, 'libr2n2_b200_ws.so')

F32, BF16 = 0, 1
EPI_BIAS, EPI_LRELU, EPI_RESIDUAL, EPI_MASK = 1, 2, 4, 8
EPI_RES_PRE = 64
CONV_UP2, CONV_DOWN2 = 16, 32
WGRAD_ACCUMULATE, WGRAD_UP2 = 1, 2
ALGO_SIMT, ALGO_UMMA = 0, 1
GRU_UR, GRU_BLEND, GRU_RBWD, GRU_OBWD = 1, 2, 3, 4
ALGO_SPLIT = 2      # host-level only: three ALGO_UMMA passes over hi/lo split operands (COMPUTE='bf16x3')
ELT_SIGMOID, ELT_TANH, ELT_COMPLEMENT, ELT_MUL = 0, 1, 2, 3
ELT_SIGMOID_BWD, ELT_TANH_BWD, ELT_NEG = 4, 5, 6

_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (must mirror include/r2n2_b200.h exactly; tests/test_abi.py checks the list)
PROTOTYPES = {
    'r2n2_conv_fwd': [_p, _p, _p, _p, _p, _p] + [_i] * 9 + [_i, _i, _i, _i, _p],
    'r2n2_conv_fwd_colsum': [_p, _p, _p, _p, _p, _p] + [_i] * 9 + [_i, _i, _i, _i, _p, _p],
    'r2n2_conv_fwd_gru': [_i] + [_p] * 10 + [_i] * 10 + [_p],
    'r2n2_conv_wgrad': [_p, _p, _p] + [_i] * 9 + [_i, _i, _i, _p],
    'r2n2_bias_grad': [_p, _p, _ll, _i, _i, _i, _p],
    'r2n2_weight_transpose': [_p, _p, _i, _i, _i, _i, _p],
    'r2n2_weight_transpose_batch': [_p, _p, _p, _i, _ll, _i, _p],
    'r2n2_weight_transpose_batch_bf16': [_p, _p, _p, _i, _ll, _p],
    'r2n2_cast': [_p, _p, _ll, _i, _p],
    'r2n2_split_bf16': [_p, _p, _p, _ll, _p],
    'r2n2_nchw_to_nhwc': [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_nchw_to_rowpack': [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_maxpool2d_fwd': [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    'r2n2_maxpool2d_bwd': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_maxpool2d_bwd_fused': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_maxpool2d_fwd_code': [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    'r2n2_maxpool2d_bwd_code': [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_unpool3d_fwd': [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_unpool3d_bwd': [_p, _p, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_add': [_p, _p, _p, _ll, _i, _p],
    'r2n2_lrelu_fwd': [_p, _p, _ll, _i, _p],
    'r2n2_lrelu_bwd': [_p, _p, _p, _p, _ll, _i, _p],
    'r2n2_lrelu_bwd_colsum': [_p, _p, _p, _p, _ll, _i, _p, _i, _p],
    'r2n2_eltwise': [_i, _p, _p, _p, _ll, _i, _p],
    'r2n2_gru_gates_ur_fwd': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    'r2n2_gru_gates_out_fwd': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    'r2n2_gru_gates_out_bwd': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    'r2n2_gru_gates_r_bwd': [_p, _p, _p, _p, _p, _i, _i, _p],
    'r2n2_lstm_gates_fwd': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    'r2n2_lstm_gates_bwd': [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p],
    'r2n2_softmax_ce3d': [_p, _p, _p, _p, _p, _i, _i, _f, _i, _i, _p],
    'r2n2_adam': [_p, _p, _p, _p, _p, _ll, _ll, _f, _f, _f, _f, _f, _f, _i, _p],
    'r2n2_sgd': [_p, _p, _p, _p, _ll, _ll, _f, _f, _f, _f, _i, _p],
    'r2n2_voxel_eval': [_p, _p, _f, _p, _i, _i, _p],
    'r2n2_preprocess_views': [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p],
    'r2n2_average_precision': [_p, _p, _p, _i, _i, _p],
    'r2n2_voxel_labels': [_p, _p, _i, _i, _p],
    'r2n2_fill_zero': [_p, _ll, _p],
    'r2n2_umma_error_flag': [],
    'r2n2_set_sm_limit': [_i],
}

_lib = None


def load():
    """Load the library (building it is `__graft_entry__.build()`'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libr2n2_b200.so not found at %s: build it with `python __graft_entry__.py` '
            '(there is no CPU/PyTorch fallback for the hot path)' % LIB_PATH)
    import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)
    lib = C.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.r2n2_version.restype = C.c_char_p
    lib.r2n2_last_error.restype = C.c_char_p
    _lib = lib
    return lib


class R2N2Error(RuntimeError):
    pass


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise R2N2Error('%s failed (%d): %s' % (name, rc, lib.r2n2_last_error().decode()))


def version():
    return load().r2n2_version().decode()
