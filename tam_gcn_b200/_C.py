"""ctypes binding of libtamgcn.so (the C-ABI declared in include/tamgcn.h).

There is no fallback: if the shared library is missing, or a call is made without a CUDA device,
this module raises.  Build the library with `python -m tam_gcn_b200.build` (or
`__graft_entry__.build()`).
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TAMGCN_LIB') or os.path.join(PKG, 'lib', 'libtamgcn.so')   # TAMGCN_LIB: A/B timing of kernel variants

F32, BF16 = 0, 1
RES_NONE, RES_IDENTITY, RES_AFFINE = 0, 1, 2

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f32 = C.c_float
f64 = C.c_double


class Operand(C.Structure):
    _fields_ = [('p', vp), ('q', vp), ('a', vp), ('b', vp), ('c', vp), ('p_nstride', i64), ('q_nstride', i64),
                ('relu', C.c_int32), ('reserved', C.c_int32)]


class ConvGeom(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ('N', 'Cin', 'Cout', 'T', 'To', 'V', 'k', 'stride', 'dil', 'pad')]


class BnDesc(C.Structure):
    _fields_ = [('sum', vp), ('sumsq', vp), ('gamma', vp), ('beta', vp), ('rmean', vp), ('rvar', vp), ('nbt', vp),
                ('scale', vp), ('shift', vp), ('mean', vp), ('invstd', vp), ('C', C.c_int32), ('reserved', C.c_int32)]


class BnBwdDesc(C.Structure):
    _fields_ = [('s1', vp), ('s2', vp), ('gamma', vp), ('mean', vp), ('invstd', vp), ('A', vp), ('B', vp), ('Cc', vp),
                ('dgamma', vp), ('dbeta', vp), ('C', C.c_int32), ('reserved', C.c_int32)]


OP = C.POINTER(Operand)
GEOM = C.POINTER(ConvGeom)

# name -> argtypes; must list every entry point of include/tamgcn.h (tests/test_cabi.py checks this)
SIGNATURES = {
    'tamgcn_conv_pack_weights': [vp, i32, i32, i32, vp, vp, vp],
    'tamgcn_conv_pack_weights_batched': [vp, i32, vp],
    'tamgcn_conv_needs_pack': [i32, i32, i32, i32, i32, i32],
    'tamgcn_conv_fwd': [GEOM, i32, OP, vp, vp, vp, vp, i64, vp, vp, i32, vp],
    'tamgcn_conv_dgrad': [GEOM, i32, OP, vp, vp, vp, i64, vp, i64, vp, f32, OP, vp, vp, vp],
    'tamgcn_conv_wgrad': [GEOM, i32, OP, OP, vp, vp, vp],
    'tamgcn_mean_t': [i32, vp, i64, i32, i32, i32, i32, vp, vp],
    'tamgcn_ctrgc_fwd': [i32, vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp],
    'tamgcn_ctrgc_bwd': [i32, OP, vp, i64, i32, i32, i32, i32, i32, i32, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp,
                         vp, vp, vp, vp, vp],
    'tamgcn_bn_finalize': [i32, C.POINTER(BnDesc), f64, f32, f32, i32, vp],
    'tamgcn_bn_bwd_coef': [i32, C.POINTER(BnBwdDesc), f64, i32, vp],
    'tamgcn_gcn_epilogue_fwd': [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, i32, vp, i64, vp, vp, vp, vp],
    'tamgcn_gcn_epilogue_bwd': [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    'tamgcn_gcn_mid_bwd': [i32, i32, i32, i32, vp, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp],
    'tamgcn_coef_diff': [i32, vp, vp, vp, vp, vp, vp],
    'tamgcn_tcn_epilogue_fwd': [i32, i32, i32, i32, vp, i64, vp, vp, i32, vp, i64, vp, vp, i32, vp, vp],
    'tamgcn_tcn_epilogue_bwd': [i32, i32, i32, i32, vp, vp, i32, vp, i64, vp, i64, vp, vp, vp, vp, vp],
    'tamgcn_maxpool_fwd': [i32, i32, i32, i32, i32, i32, i32, OP, vp, i64, vp, vp, vp],
    'tamgcn_maxpool_bwd': [i32, i32, i32, i32, i32, i32, i32, OP, OP, vp, i64, vp, vp, vp],
    'tamgcn_graph_agg_fwd': [i32, i32, i32, i32, i32, i32, vp, i64, vp, vp, i64, vp, vp, vp],
    'tamgcn_graph_agg_bwd': [i32, i32, i32, i32, i32, i32, OP, vp, i64, vp, vp, i64, vp, vp],
    'tamgcn_data_bn_fwd': [i32, vp, C.POINTER(i64), i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp,
                           vp, vp],
    'tamgcn_data_bn_bwd': [i32, vp, vp, C.POINTER(i64), i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp],
    'tamgcn_pool_fc_fwd': [i32, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    'tamgcn_pool_fc_bwd': [i32, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    'tamgcn_transpose_act': [vp, vp, i32, i32, i32, vp, vp],
    'tamgcn_softmax_ce_fwd': [vp, vp, i32, i32, vp, vp, vp],
    'tamgcn_softmax_ce_bwd': [vp, vp, i32, i32, vp, vp],
    'tamgcn_sgd_step': [vp, vp, vp, i64, vp, f32, f32, i32, f32, vp],
    'tamgcn_feeder_nucla': [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp],
}

_lib = None


def lib():
    """Load libtamgcn.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libtamgcn.so not found at %s — run `python -m tam_gcn_b200.build`; '
                               'there is no CPU or PyTorch fallback for the CUDA kernels' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.tamgcn_version.restype = i32
        l.tamgcn_last_error.restype = C.c_char_p
        l.tamgcn_launch_count.restype = i64
        l.tamgcn_set_wgrad_sm_share.restype = i32
        l.tamgcn_set_wgrad_sm_share.argtypes = [i32]
        l.tamgcn_set_main_sm_share.restype = i32
        l.tamgcn_set_main_sm_share.argtypes = [i32]
        l.tamgcn_conv_pack_bytes.restype = i64
        l.tamgcn_conv_pack_bytes.argtypes = [i32, i32, i32, i32]
        for name, args in SIGNATURES.items():
            f = getattr(l, name)
            f.argtypes = args
            f.restype = i32
        _lib = l
    return _lib


def check(rc, name):
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, rc, lib().tamgcn_last_error().decode()))


def launch_count():
    return int(lib().tamgcn_launch_count())
