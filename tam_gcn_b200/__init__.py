"""tam_gcn_b200 — B200-native (sm_100a) CTR-GCN / ST-GCN skeleton-graph-convolution hot path.

Drop-in modules with the reference's names, signatures and state_dict layout:

    tam_gcn_b200.ctrgcn : TemporalConv, MultiScale_TemporalConv, CTRGC, unit_tcn, unit_gcn, TCN_GCN_unit, Model
    tam_gcn_b200.stgcn  : ConvTemporalGraphical, st_gcn, Model

backed by hand-written CUDA kernels behind a C-ABI (include/tamgcn.h, tam_gcn_b200/csrc).  There is no
CPU fallback, no Triton and no multi-backend dispatch: without libtamgcn.so and a CUDA device the
modules raise.
"""
import contextlib
import os

import torch

_ACT_DTYPE = {'f32': torch.float32, 'fp32': torch.float32, 'float32': torch.float32,
              'bf16': torch.bfloat16, 'bfloat16': torch.bfloat16}
_act_dtype = _ACT_DTYPE[os.environ.get('TAMGCN_ACT_DTYPE', 'f32').lower()]


def get_act_dtype():
    """Storage dtype of activations inside `Model` (parameters, BN statistics and accumulation stay fp32)."""
    return _act_dtype


def set_act_dtype(dtype):
    global _act_dtype
    if isinstance(dtype, str):
        dtype = _ACT_DTYPE[dtype.lower()]
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError('activation dtype must be float32 or bfloat16')
    _act_dtype = dtype


@contextlib.contextmanager
def act_dtype(dtype):
    old = get_act_dtype()
    set_act_dtype(dtype)
    try:
        yield
    finally:
        set_act_dtype(old)


def patch_reference(ctrgcn_module=None, stgcn_module=None):
    """Rebind the layer classes inside the reference's `models.ctrgcn` / `models.stgcn` modules to the
    B200-native ones, so that the reference's own `Model` classes (and therefore main.py, processor/,
    ensemble/, tools/) run unchanged on the new kernels.  The reference looks these names up as module
    globals at construction time (models/ctrgcn.py:207,269-270,280,305-314; models/stgcn.py:75,141-150)."""
    from . import ctrgcn as C, stgcn as S
    if ctrgcn_module is not None:
        for name in ('TemporalConv', 'MultiScale_TemporalConv', 'CTRGC', 'unit_tcn', 'unit_gcn', 'TCN_GCN_unit'):
            setattr(ctrgcn_module, name, getattr(C, name))
    if stgcn_module is not None:
        for name in ('ConvTemporalGraphical', 'st_gcn'):
            setattr(stgcn_module, name, getattr(S, name))
