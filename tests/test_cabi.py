"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/tamgcn.h declares
(no compute calls: there is no GPU in the CPU test tier)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'tamgcn.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tamgcn_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert 'tamgcn_ctrgc_fwd' in syms and 'tamgcn_conv_wgrad' in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for s in declared_symbols():
        assert hasattr(lib, s), 'libtamgcn.so does not export %s' % s


def test_binding_covers_header(lib_built):
    from tam_gcn_b200 import _C
    bound = set(_C.SIGNATURES) | {'tamgcn_version', 'tamgcn_last_error', 'tamgcn_launch_count', 'tamgcn_conv_pack_bytes',
                                     'tamgcn_set_wgrad_sm_share', 'tamgcn_set_main_sm_share'}
    assert bound == set(declared_symbols())
    l = _C.lib()
    assert l.tamgcn_version() >= 100
    assert _C.launch_count() == 0


def test_argument_validation_needs_no_gpu(lib_built):
    """Bad arguments are rejected on the host before any CUDA call, with a message."""
    from tam_gcn_b200 import _C
    l = _C.lib()
    g = _C.ConvGeom()
    g.N, g.Cin, g.Cout, g.T, g.To, g.V, g.k, g.stride, g.dil, g.pad = 1, 4, 4, 8, 7, 20, 1, 1, 1, 0   # To wrong
    op = _C.Operand()
    rc = l.tamgcn_conv_fwd(ctypes.byref(g), 0, ctypes.byref(op), None, None, None, None, 0, None, None, 0, None)
    assert rc < 0 and b'inconsistent' in l.tamgcn_last_error()
    rc = l.tamgcn_ctrgc_fwd(0, None, 0, 1, 64, 8, 17, 3, 8, None, None, 0, None, None, None, None, None, 0, None, None, None)
    assert rc < 0 and b'V=17' in l.tamgcn_last_error()


def test_built_for_sm_100a(lib_built):
    out = subprocess.run(['cuobjdump', '-lelf', lib_built], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_modules_refuse_cpu_tensors():
    """No CPU fallback: the modules raise instead of silently computing elsewhere."""
    import torch
    import tam_gcn_b200.ctrgcn as C
    m = C.unit_tcn(4, 4, kernel_size=1)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.zeros(1, 4, 3, 20))
