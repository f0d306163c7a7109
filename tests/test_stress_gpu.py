"""Seeded random-shape sweep of the kernels whose tiling depends on the shape in several ways at once (partial time
tiles, ragged channel tiles, plane alignment, stride-2 phases): tconv9 forward / data gradient / weight gradient, the
warp-MMA graph aggregation and the warp-MMA CTRGC forward, each against the pure-torch emulation of the same op."""
import random

import pytest
import torch

import emu_ops as E
from test_kernels_gpu import _ctrgc_inputs, _dev, coef, gen, make_operand, real_opnd, rel, rnd, tol

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _shapes(seed, n):
    r = random.Random(seed)
    out = []
    for _ in range(n):
        Cin = r.choice([64, 80, 96, 128, 144, 256])
        Cout = r.choice([64, 72, 96, 128, 160, 256])
        k, s = r.choice([(9, 1), (9, 1), (9, 2), (7, 1), (5, 1), (3, 1)])
        T = r.randint(9, 70)
        N = r.randint(1, 4)
        out.append((N, Cin, Cout, T, k, s))
    return out


@pytest.mark.parametrize('shape', _shapes(1, 14))
def test_tconv9_random_shapes(shape):
    _dev()
    from tam_gcn_b200 import ops
    N, Cin, Cout, T, k, s = shape
    V, p = 25, (k - 1) // 2
    To = (T + 2 * p - (k - 1) - 1) // s + 1
    g = gen(100 + T + Cin)
    W = rnd(g, Cout, Cin, k, scale=(Cin * k) ** -0.5)
    b = rnd(g, Cout)
    wp = ops.conv_pack_weights(W, Cout, Cin, k)
    x = make_operand(g, N, Cin, T, V, BF, 'affine_relu', wide=True)
    dy = make_operand(g, N, Cout, To, V, BF, 'two')
    mask = E.Opnd(rnd(g, N, Cin, T, V, dt=BF), a=coef(g, Cin, True), c=coef(g, Cin))
    res = []
    for o, xo, dyo, mo in ((ops, real_opnd(x), real_opnd(dy), real_opnd(mask)), (E, x, dy, mask)):
        y = torch.zeros(N, Cout, To, V, device='cuda', dtype=BF)
        st = torch.zeros(2, Cout, device='cuda', dtype=torch.float64)
        o.conv_fwd(xo, W, b, y, k, s, 1, p, stats=(st[0], st[1]), wpack=wp[0])
        dx = torch.zeros(N, Cin, T, V, device='cuda', dtype=BF)
        sd = torch.zeros(2, Cin, device='cuda', dtype=torch.float64)
        o.conv_dgrad(dyo, W, dx, k, s, 1, p, mask=mo, stats=(sd[0], sd[1]), wpack=wp[1])
        dW = torch.zeros(Cout, Cin * k, device='cuda')
        db = torch.zeros(Cout, device='cuda')
        o.conv_wgrad(dyo, xo, dW, db, k, s, 1, p)
        res.append((y, st, dx, sd, dW, db))
    for nm, a_, b_ in zip(('y', 'stats', 'dx', 'dstats', 'dW', 'db'), res[0], res[1]):
        assert rel(a_, b_) < tol(BF), (nm, shape)


@pytest.mark.parametrize('seed', range(8))
def test_graph_agg_random_shapes(seed):
    _dev()
    from tam_gcn_b200 import ops
    r = random.Random(seed)
    V = r.choice([20, 25])
    K, C, T, N = r.randint(1, 3), r.randint(3, 40), r.randint(1, 80), r.randint(1, 5)
    g = gen(200 + seed)
    y = rnd(g, N, K * C, T, V, dt=BF)
    A = rnd(g, K, V, V, scale=0.3)
    go = make_operand(g, N, C, T, V, BF, r.choice(['plain', 'two', 'affine_relu']))
    res = []
    for o, conv in ((ops, real_opnd), (E, lambda t: t)):
        out = torch.zeros(N, C, T, V, device='cuda', dtype=BF)
        st = torch.zeros(2, C, device='cuda', dtype=torch.float64)
        o.graph_agg_fwd(y, A, out, stats=(st[0], st[1]))
        dy = torch.zeros_like(y)
        dA = torch.zeros_like(A)
        o.graph_agg_bwd(conv(go), y, A, dy, dA)
        res.append((out, st, dy, dA))
    for nm, a_, b_ in zip(('out', 'stats', 'dy', 'dA'), res[0], res[1]):
        assert rel(a_, b_) < max(tol(BF), 1e-4), (nm, V, K, C, T, N)


@pytest.mark.parametrize('seed', range(8))
def test_ctrgc_large_r_random_shapes(seed):
    """R = 16 / 24 / 32, T <= 64: the warp-MMA forward and the (staged) warp-MMA backward."""
    _dev()
    from tam_gcn_b200 import ops
    r = random.Random(seed)
    V = r.choice([20, 25])
    R = r.choice([16, 24, 32])
    Cout, T, K = r.choice([16, 40, 64, 100, 256]), r.randint(3, 64), r.randint(1, 3)
    g = gen(300 + seed)
    x3, x12, W4, b4, PA, alpha = _ctrgc_inputs(g, BF, V, Cout, R, T, K)
    go = make_operand(g, 3, Cout, T, V, BF, 'two')
    res = []
    for o, conv in ((ops, real_opnd), (E, lambda t: t)):
        y = torch.zeros(3, Cout, T, V, device='cuda', dtype=BF)
        st = torch.zeros(2, Cout, device='cuda', dtype=torch.float64)
        o.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
        dx3 = torch.zeros(3, K * Cout, T, V, device='cuda', dtype=BF)
        dx12 = torch.zeros_like(x12)
        dW4, db4, dPA, dal = torch.zeros_like(W4), torch.zeros_like(b4), torch.zeros_like(PA), torch.zeros(1, device='cuda')
        o.ctrgc_bwd(conv(go), x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, dx3, dx12[:, :K * R], dx12[:, K * R:],
                    dW4, db4, dPA, dal)
        res.append((y, st, dx3, dx12, dW4, db4, dPA))
    for nm, a_, b_ in zip(('y', 'stats', 'dx3', 'dx12', 'dW4', 'db4', 'dPA'), res[0], res[1]):
        assert rel(a_, b_) < max(tol(BF), 1e-4), (nm, V, R, Cout, T, K)
