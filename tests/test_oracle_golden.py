"""The CPU oracle (oracle/gcn_oracle.py) replayed against the golden fixtures dumped from the UNMODIFIED
reference (oracle/make_golden.py; /root/reference is not needed here).  fp64 oracle vs fp32-stored
reference outputs: agreement to fp32 storage precision pins the oracle."""
import pytest
import torch

import helpers as H
from oracle import gcn_oracle as O


@pytest.mark.parametrize('name', list(H.CASES))
def test_oracle_matches_reference_fixture(name):
    case = H.CASES[name]
    built = H.build_case(case)
    fx = H.load_fixture(name)
    chk = float(sum(v.double().abs().sum() for v in built['state'].values()
                    if torch.is_tensor(v) and v.is_floating_point()))
    assert abs(chk - fx['state_checksum']) <= 1e-9 * abs(chk), 'seeded state drifted from the one the fixture used'
    assert torch.equal(built['x'], fx['x']) and torch.equal(built['cot'], fx['cot'])
    D = torch.float64
    p = O.clone_state(built['state'], D, requires_grad=True)
    x = built['x'].to(D).requires_grad_(True)
    extra = {k: v.to(D).requires_grad_(True) for k, v in built['extra'].items()}
    y = H.oracle_forward(case, x, p, extra)
    y.backward(built['cot'].to(D))
    assert O.rel_err(y, fx['y']) < 5e-7
    assert O.rel_err(x.grad, fx['dx']) < 5e-7
    pre = '' if case['kind'].endswith('_model') else 'm.'
    gscale = max(float(v.norm()) for v in fx['grads'].values())
    for k, gref in fx['grads'].items():
        go = extra[k[2:]].grad if k.startswith('__') else p[pre + k].grad
        if float(gref.norm()) > 1e-6 * gscale:
            assert O.rel_err(go, gref) < 5e-6, k
    for k, b in fx['buffers'].items():
        if b.is_floating_point():
            assert O.rel_err(p[pre + k], b) < 5e-7, k
        else:
            assert int(p[pre + k]) == int(b), k
