"""Host-side orchestration (tam_gcn_b200/functional.py) against the golden fixtures, on CPU.

The CUDA entry points are replaced by their pure-torch emulations (tests/emu_ops.py), so what is under
test here is everything that is NOT a kernel: parameter packing, BatchNorm coefficient algebra, the
hand-derived backward passes and the routing of gradients / running statistics to the reference's
parameter names.  The kernels themselves are checked against the same emulations on the GPU box.
"""
import pytest
import torch

import emu_ops
import helpers as H

MODULE_CASES = [n for n, c in H.CASES.items() if not c['kind'].endswith('_model')]
MODEL_CASES = [n for n, c in H.CASES.items() if c['kind'].endswith('_model')]


@pytest.mark.parametrize('name', MODULE_CASES)
def test_module_matches_golden(name, monkeypatch):
    emu_ops.install(monkeypatch)
    res = H.run_case(name, 'cpu')
    rep = []
    fails = H.compare(name, res, H.load_fixture(name), tol_y=2e-5, tol_dx=2e-4, tol_g=5e-4, tol_buf=1e-5, report=rep)
    print(rep[0])
    assert not fails, '\n'.join(fails)


@pytest.mark.parametrize('name', MODEL_CASES)
def test_model_matches_golden(name, monkeypatch):
    emu_ops.install(monkeypatch)
    res = H.run_case(name, 'cpu')
    fx = H.load_fixture(name)
    rep = []
    # end-to-end gradients carry the reference's own fp32 noise (2e-3 on dx): loose bounds
    fails = H.compare(name, res, fx, tol_y=1e-4, tol_dx=2e-2, tol_g=2e-2, tol_buf=1e-4, report=rep)
    print(rep[0])
    assert (res['y'].argmax(1) == fx['y'].argmax(1)).all()
    assert not fails, '\n'.join(fails)
