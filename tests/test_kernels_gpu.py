"""Every CUDA entry point of libtamgcn.so against its pure-torch emulation (tests/emu_ops.py), called through
the C-ABI (ctypes) on the same seeded inputs.  fp32 storage: <= 2e-5 relative (fp32 accumulation-order
noise); bf16 storage: the emulation rounds at the same points, tolerance 1.5e-2 relative (bf16 eps = 3.9e-3).
"""
import itertools

import pytest
import torch

import emu_ops as E

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16]


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def tol(dt):
    return 2e-5 if dt == torch.float32 else 1.5e-2


def rnd(g, *shape, dt=torch.float32, scale=1.0):
    return (torch.randn(*shape, generator=g, device='cuda') * scale).to(dt)


def gen(seed):
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    return g


def coef(g, C, pos=False):
    c = torch.randn(C, generator=g, device='cuda') * 0.5
    return (c.abs() + 0.5) if pos else c


def make_operand(g, N, C, T, V, dt, mode, wide=False):
    """mode: plain | affine_relu | two | slice-of-wider (wide=True wraps p in a channel slice of a wider tensor)."""
    def act():
        if wide:
            big = rnd(g, N, C + 5, T, V, dt=dt)
            return big[:, 3:3 + C]
        return rnd(g, N, C, T, V, dt=dt)
    if mode == 'plain':
        p = act()
        return E.Opnd(p)
    if mode == 'affine_relu':
        return E.Opnd(act(), a=coef(g, C, True), c=coef(g, C), relu=True)
    if mode == 'two':
        return E.Opnd(act(), act(), a=coef(g, C), b=coef(g, C), c=coef(g, C))
    raise KeyError(mode)


def real_opnd(o):
    from tam_gcn_b200.ops import Opnd
    return Opnd(o.p, o.q, o.a, o.b, o.c, o.relu)


CONV_GEOMS = [  # (Cin, Cout, T, k, s, d)
    (64, 64, 12, 1, 1, 1), (3, 70, 9, 1, 1, 1), (16, 16, 13, 5, 1, 2), (16, 24, 13, 5, 2, 1), (32, 48, 12, 1, 2, 1),
    (24, 24, 17, 9, 2, 1), (8, 8, 11, 3, 1, 4), (96, 10, 1, 1, 1, 1), (72, 272, 52, 1, 1, 1), (300, 40, 7, 3, 1, 1),
    # the MS-TCN branch shapes served by the warp-MMA kernels (tconv_mma.cu)
    (16, 16, 52, 5, 1, 1), (32, 32, 26, 5, 2, 2), (64, 64, 13, 5, 1, 2), (32, 32, 9, 3, 1, 1), (16, 16, 7, 5, 2, 1),
    # the ST-GCN temporal convolutions served (at V = 25) by the V-padded tcgen05 kernel (tconv9.cu)
    (64, 64, 40, 9, 1, 1), (128, 64, 35, 9, 1, 1), (64, 128, 75, 5, 1, 1), (256, 256, 19, 9, 1, 1), (64, 64, 300, 9, 1, 1),
    # ... and the stride-2 layers (l5, l8)
    (64, 64, 40, 9, 2, 1), (128, 128, 75, 9, 2, 1), (64, 128, 31, 9, 2, 1),
    # 1x1 graph convolutions on planes of odd length (T*V % 4 != 0 at V = 25): weight gradient through tconv9's k = 1 tiles
    (64, 192, 75, 1, 1, 1), (256, 96, 30, 1, 1, 1),
    # ragged channel counts / odd lengths through the same kernels
    (80, 96, 23, 9, 1, 1), (64, 64, 41, 9, 2, 1), (96, 80, 33, 7, 1, 1)]


def _pad(k, d):
    return (k + (k - 1) * (d - 1) - 1) // 2


@pytest.mark.parametrize('dt,geom,V,mode', [(dt, gm, V, md) for dt in DT for gm in CONV_GEOMS for V, md in
                                            [(20, 'plain'), (25, 'affine_relu'), (20, 'two'), (25, 'plain')]])
def test_conv_fwd(dt, geom, V, mode):
    _dev()
    from tam_gcn_b200 import ops
    Cin, Cout, T, k, s, d = geom
    N, p = 3, _pad(k, d)
    To = (T + 2 * p - d * (k - 1) - 1) // s + 1
    g = gen(1)
    x = make_operand(g, N, Cin, T, V, dt, mode, wide=(mode != 'plain'))
    W = rnd(g, Cout, Cin, k, scale=(Cin * k) ** -0.5)
    b = rnd(g, Cout)
    c0 = Cout // 3
    outs = []
    wp = ops.conv_pack_weights(W, Cout, Cin, k) if dt == torch.bfloat16 else (None, None)
    for fn, xo in ((ops.conv_fwd, real_opnd(x)), (E.conv_fwd, x)):
        big = torch.zeros(N, Cout + 4, To, V, device='cuda', dtype=dt)
        y = big[:, 2:2 + Cout]
        st = torch.zeros(2, Cout - c0, device='cuda', dtype=torch.float64)
        fn(xo, W, b, y, k, s, d, p, stats=(st[0], st[1]), stat_c0=c0, wpack=wp[0])
        outs.append((big, st))
    assert rel(outs[0][0], outs[1][0]) < tol(dt)
    assert rel(outs[0][1], outs[1][1]) < tol(dt)


@pytest.mark.parametrize('dt,geom,V,opts', [(dt, gm, V, o) for dt in DT for gm in CONV_GEOMS for V, o in
                                            [(20, 'plain'), (25, 'mask'), (20, 'addend')]])
def test_conv_dgrad(dt, geom, V, opts):
    _dev()
    from tam_gcn_b200 import ops
    Cin, Cout, T, k, s, d = geom
    N, p = 3, _pad(k, d)
    To = (T + 2 * p - d * (k - 1) - 1) // s + 1
    g = gen(2)
    dy = make_operand(g, N, Cout, To, V, dt, 'two' if opts != 'plain' else 'plain')
    W = rnd(g, Cout, Cin, k, scale=(Cout * k) ** -0.5)
    addend = rnd(g, N, Cin, T, V, dt=dt) if opts == 'addend' else None
    bcast = rnd(g, N, Cin, 1, V) if opts == 'addend' else None
    mask = E.Opnd(rnd(g, N, Cin, T, V, dt=dt), a=coef(g, Cin, True), c=coef(g, Cin)) if opts == 'mask' else None
    outs = []
    wp = ops.conv_pack_weights(W, Cout, Cin, k) if dt == torch.bfloat16 else (None, None)
    for fn, conv in ((ops.conv_dgrad, real_opnd), (E.conv_dgrad, lambda o: o)):
        dx = torch.zeros(N, Cin, T, V, device='cuda', dtype=dt)
        st = torch.zeros(2, Cin, device='cuda', dtype=torch.float64)
        fn(conv(dy), W, dx, k, s, d, p, addend=addend, bcast=bcast, bcast_scale=0.25,
           mask=conv(mask) if mask is not None else None, stats=(st[0], st[1]) if mask is not None else None,
           wpack=wp[1])
        outs.append((dx, st))
    assert rel(outs[0][0], outs[1][0]) < tol(dt)
    if mask is not None:
        assert rel(outs[0][1], outs[1][1]) < tol(dt)


@pytest.mark.parametrize('Cin,Cout,V', [(64, 48, 20), (256, 192, 20), (3, 16, 20), (128, 96, 25), (70, 50, 25)])
def test_conv_small_l(Cin, Cout, V):
    """fp32 1x1 convs over a handful of positions per sample (CTRGC conv1 / conv2 on the T-mean): dedicated kernels."""
    _dev()
    from tam_gcn_b200 import ops
    N = 5
    g = gen(11)
    x = rnd(g, N, Cin, 1, V)
    dy = rnd(g, N, Cout, 1, V)
    W = rnd(g, Cout, Cin, scale=Cin ** -0.5)
    b = rnd(g, Cout)
    outs = []
    for fwd, dgrad, conv in ((ops.conv_fwd, ops.conv_dgrad, real_opnd), (E.conv_fwd, E.conv_dgrad, lambda o: o)):
        y = torch.zeros(N, Cout, 1, V, device='cuda')
        dx = torch.zeros(N, Cin, 1, V, device='cuda')
        fwd(conv(E.Opnd(x)), W, b, y)
        dgrad(conv(E.Opnd(dy)), W, dx)
        outs.append((y, dx))
    assert rel(outs[0][0], outs[1][0]) < 1e-5
    assert rel(outs[0][1], outs[1][1]) < 1e-5


@pytest.mark.parametrize('dt,geom,V', [(dt, gm, V) for dt in DT for gm in CONV_GEOMS for V in (20, 25)])
def test_conv_wgrad(dt, geom, V):
    _dev()
    from tam_gcn_b200 import ops
    Cin, Cout, T, k, s, d = geom
    N, p = 5, _pad(k, d)
    To = (T + 2 * p - d * (k - 1) - 1) // s + 1
    g = gen(3)
    dy = make_operand(g, N, Cout, To, V, dt, 'two')
    x = make_operand(g, N, Cin, T, V, dt, 'affine_relu', wide=True)
    outs = []
    for fn, conv in ((ops.conv_wgrad, real_opnd), (E.conv_wgrad, lambda o: o)):
        dW = torch.zeros(Cout, Cin * k, device='cuda')
        db = torch.zeros(Cout, device='cuda')
        fn(conv(dy), conv(x), dW, db, k, s, d, p)
        outs.append((dW, db))
    assert rel(outs[0][0], outs[1][0]) < tol(dt)
    assert rel(outs[0][1], outs[1][1]) < tol(dt)


@pytest.mark.parametrize('dt,V', itertools.product(DT, (20, 25)))
def test_mean_t(dt, V):
    _dev()
    from tam_gcn_b200 import ops
    g = gen(4)
    N, C, T = 3, 19, 13
    x = rnd(g, N, C + 2, T, V, dt=dt)[:, 1:1 + C]
    m0 = torch.empty(N, C, 1, V, device='cuda')
    m1 = torch.empty_like(m0)
    ops.mean_t(x, m0)
    E.mean_t(x, m1)
    assert rel(m0, m1) < 1e-6


CTRGC_CFG = [(20, 64, 8, 12, 3), (20, 24, 8, 52, 3), (25, 64, 16, 10, 3), (25, 40, 32, 7, 1), (20, 256, 32, 13, 3),
             (25, 64, 8, 64, 3), (20, 10, 8, 70, 3), (20, 128, 16, 26, 3), (20, 6, 8, 5, 1), (25, 7, 8, 33, 2),
             (20, 64, 8, 52, 3), (20, 8, 8, 40, 3), (20, 12, 8, 33, 2), (20, 4, 8, 64, 1), (20, 256, 8, 34, 3),
             # V = 25, R = 8, 32 < T <= 64, T % 8 == 0: ctrgc_tc4.cu
             (25, 128, 8, 48, 3), (25, 8, 8, 40, 2), (25, 64, 8, 56, 1), (25, 4, 8, 64, 3),
             # V = 25, R > 16: the backward's staged (LEAN) variant
             (25, 256, 32, 16, 3), (25, 48, 24, 9, 2)]


def _ctrgc_inputs(g, dt, V, Cout, R, T, K, N=3):
    x3 = rnd(g, N, K * Cout + 3, T, V, dt=dt)[:, :K * Cout]
    x12 = rnd(g, N, 2 * K * R, 1, V)
    W4 = rnd(g, K, Cout, R, scale=R ** -0.5)
    b4 = rnd(g, K, Cout, scale=0.1)
    PA = rnd(g, K, V, V, scale=0.3)
    alpha = torch.full((1,), 0.7, device='cuda')
    return x3, x12, W4, b4, PA, alpha


@pytest.mark.parametrize('dt,cfg', itertools.product(DT, CTRGC_CFG))
def test_ctrgc_fwd(dt, cfg):
    _dev()
    from tam_gcn_b200 import ops
    V, Cout, R, T, K = cfg
    x3, x12, W4, b4, PA, alpha = _ctrgc_inputs(gen(5), dt, V, Cout, R, T, K)
    outs = []
    for fn in (ops.ctrgc_fwd, E.ctrgc_fwd):
        y = torch.zeros(3, Cout, T, V, device='cuda', dtype=dt)
        st = torch.zeros(2, Cout, device='cuda', dtype=torch.float64)
        fn(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
        outs.append((y, st))
    assert rel(outs[0][0], outs[1][0]) < tol(dt)
    assert rel(outs[0][1], outs[1][1]) < tol(dt)


@pytest.mark.parametrize('dt,cfg', itertools.product(DT, CTRGC_CFG))
def test_ctrgc_bwd(dt, cfg):
    _dev()
    from tam_gcn_b200 import ops
    V, Cout, R, T, K = cfg
    g = gen(6)
    x3, x12, W4, b4, PA, alpha = _ctrgc_inputs(g, dt, V, Cout, R, T, K)
    go = make_operand(g, 3, Cout, T, V, dt, 'two')
    outs = []
    for fn, conv in ((ops.ctrgc_bwd, real_opnd), (E.ctrgc_bwd, lambda o: o)):
        dx3 = torch.zeros(3, K * Cout, T, V, device='cuda', dtype=dt)
        dx12 = torch.zeros_like(x12)
        dW4, db4, dPA, dal = torch.zeros_like(W4), torch.zeros_like(b4), torch.zeros_like(PA), torch.zeros(1, device='cuda')
        fn(conv(go), x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, dx3, dx12[:, :K * R], dx12[:, K * R:], dW4,
           db4, dPA, dal)
        outs.append((dx3, dx12, dW4, db4, dPA, dal))
    names = ('dx3', 'dx12', 'dW4', 'db4', 'dPA', 'dalpha')
    for nm, a, b in zip(names, outs[0], outs[1]):
        t = tol(dt) if nm == 'dx3' else max(tol(dt), 1e-4)
        if nm == 'dalpha' and dt == torch.bfloat16:
            t = 0.15      # one cancellation-heavy scalar: the bf16 rounding of the cotangent operand dominates it
        assert rel(a, b) < t, nm


@pytest.mark.parametrize('train', [True, False])
def test_bn_coefficients(train):
    _dev()
    from tam_gcn_b200 import ops
    g = gen(7)
    Cs = [16, 40, 7]
    count = 1234.0

    def mk():
        gg = gen(70)
        ds = []
        for Cn in Cs:
            mean = torch.randn(Cn, generator=gg, device='cuda', dtype=torch.float64)
            var = torch.rand(Cn, generator=gg, device='cuda', dtype=torch.float64) + 0.1
            ds.append(dict(sum=mean * count, sumsq=(var + mean * mean) * count, gamma=coef(gg, Cn), beta=coef(gg, Cn),
                           rmean=coef(gg, Cn), rvar=coef(gg, Cn, True), nbt=torch.zeros((), dtype=torch.long, device='cuda'),
                           scale=torch.empty(Cn, device='cuda'), shift=torch.empty(Cn, device='cuda'),
                           mean=torch.empty(Cn, device='cuda'), invstd=torch.empty(Cn, device='cuda')))
        return ds
    a, b = mk(), mk()
    ops.bn_finalize(a, count, 0.1, 1e-5, train)
    E.bn_finalize(b, count, 0.1, 1e-5, train)
    for da, db in zip(a, b):
        for k in ('scale', 'shift', 'mean', 'invstd', 'rmean', 'rvar'):
            assert rel(da[k], db[k]) < 1e-6, k
        assert int(da['nbt']) == int(db['nbt']) == (1 if train else 0)

    def mkb(src):
        gg = gen(71)
        ds = []
        for d in src:
            Cn = d['scale'].numel()
            ds.append(dict(s1=torch.randn(Cn, generator=gg, device='cuda', dtype=torch.float64) * 30,
                           s2=torch.randn(Cn, generator=gg, device='cuda', dtype=torch.float64) * 30, gamma=d['gamma'],
                           mean=d['mean'], invstd=d['invstd'], A=torch.empty(Cn, device='cuda'),
                           B=torch.empty(Cn, device='cuda'), Cc=torch.empty(Cn, device='cuda'),
                           dgamma=torch.empty(Cn, device='cuda'), dbeta=torch.empty(Cn, device='cuda')))
        return ds
    a2, b2 = mkb(a), mkb(b)
    ops.bn_bwd_coef(a2, count, train)
    E.bn_bwd_coef(b2, count, train)
    for da, db in zip(a2, b2):
        for k in ('A', 'B', 'Cc', 'dgamma', 'dbeta'):
            assert rel(da[k], db[k]) < 1e-6, k


@pytest.mark.parametrize('dt,res_mode,TV', [(dt, rm, tv) for dt in DT for rm in (0, 1, 2) for tv in ((13, 20), (4, 25))])
def test_gcn_epilogues(dt, res_mode, TV):
    _dev()
    from tam_gcn_b200 import ops
    T, V = TV
    N, C = 5, 24
    g = gen(8)
    y0, z = rnd(g, N, C, T, V, dt=dt), rnd(g, N, C, T, V, dt=dt)
    sg, hg, so, ho, sr, hr = (coef(g, C) for _ in range(6))
    r = rnd(g, N, C + 4, T, V, dt=dt)[:, 2:2 + C] if res_mode else None
    o0, o1 = torch.empty_like(y0), torch.empty_like(y0)
    ops.gcn_epilogue_fwd(y0, sg, hg, z, so, ho, res_mode, r, sr if res_mode == 2 else None, hr if res_mode == 2 else None, o0)
    E.gcn_epilogue_fwd(y0, sg, hg, z, so, ho, res_mode, r, sr, hr, o1)
    assert rel(o0, o1) < tol(dt)
    gr = rnd(g, N, C, T, V, dt=dt)
    outs = []
    for fn in (ops.gcn_epilogue_bwd, E.gcn_epilogue_bwd):
        G, DZ = torch.empty_like(y0), torch.empty_like(y0)
        st = torch.zeros(2, C, device='cuda', dtype=torch.float64)
        fn(gr, o1, z, so, ho, G, DZ, st[0], st[1])
        outs.append((G, DZ, st))
    for a, b in zip(outs[0], outs[1]):
        assert rel(a, b) < tol(dt)
    DD = rnd(g, N, C, T, V, dt=dt)
    outs = []
    for fn in (ops.gcn_mid_bwd, E.gcn_mid_bwd):
        G = rnd(gen(80), N, C, T, V, dt=dt)
        G0 = G.clone()
        big = torch.zeros(N, C + 4, T, V, device='cuda', dtype=dt)
        dr = big[:, 1:1 + C] if res_mode else None
        st = torch.zeros(4, C, device='cuda', dtype=torch.float64)
        extra = rnd(gen(81), N, C, T, V, dt=dt) if res_mode == 1 else None     # second cotangent of an identity residual
        fn(G, DD, dr, y0, r if res_mode == 2 else None, st[0], st[1], st[2] if res_mode == 2 else None,
           st[3] if res_mode == 2 else None, extra=extra)
        outs.append((G, big, st, G0))
    assert rel(outs[0][0], outs[1][0]) < tol(dt)
    if res_mode:
        assert rel(outs[0][1], outs[1][1]) < tol(dt)
    assert rel(outs[0][2], outs[1][2]) < tol(dt)


@pytest.mark.parametrize('dt,res_mode,relu', [(dt, rm, rl) for dt in DT for rm in (0, 1, 2) for rl in (True, False)])
def test_tcn_epilogues(dt, res_mode, relu):
    _dev()
    from tam_gcn_b200 import ops
    N, C, T, V = 5, 24, 7, 25
    g = gen(9)
    u = rnd(g, N, C + 3, T, V, dt=dt)[:, 1:1 + C]
    su, hu, sr, hr = (coef(g, C) for _ in range(4))
    r = rnd(g, N, C, T, V, dt=dt) if res_mode else None
    o0 = torch.empty(N, C, T, V, device='cuda', dtype=dt)
    o1 = torch.empty_like(o0)
    ops.tcn_epilogue_fwd(u, su, hu, res_mode, r, sr if res_mode == 2 else None, hr if res_mode == 2 else None, relu, o0)
    E.tcn_epilogue_fwd(u, su, hu, res_mode, r, sr, hr, relu, o1)
    assert rel(o0, o1) < tol(dt)
    gr = rnd(g, N, C, T, V, dt=dt)
    outs = []
    for fn in (ops.tcn_epilogue_bwd, E.tcn_epilogue_bwd):
        G = torch.zeros_like(o0) if relu else None
        st = torch.zeros(3, C, device='cuda', dtype=torch.float64)
        fn(gr, o1 if relu else None, relu, u, r if res_mode == 2 else None, G, st[0], st[1], st[2] if res_mode == 2 else None)
        outs.append((G, st))
    if relu:
        assert rel(outs[0][0], outs[1][0]) < tol(dt)
    assert rel(outs[0][1], outs[1][1]) < tol(dt)


@pytest.mark.parametrize('dt,stride,T,V', [(dt, s, T, V) for dt in DT for s in (1, 2) for T, V in ((12, 20), (13, 25), (1, 20))])
def test_maxpool(dt, stride, T, V):
    _dev()
    from tam_gcn_b200 import ops
    N, C = 4, 10
    To = (T + 2 - 3) // stride + 1
    g = gen(10)
    x = make_operand(g, N, C, T, V, dt, 'affine_relu', wide=True)
    outs = []
    for fn, conv in ((ops.maxpool_fwd, real_opnd), (E.maxpool_fwd, lambda o: o)):
        y = torch.zeros(N, C, To, V, device='cuda', dtype=dt)
        st = torch.zeros(2, C, device='cuda', dtype=torch.float64)
        fn(conv(x), y, stride, stats=(st[0], st[1]))
        outs.append((y, st))
    assert rel(outs[0][0], outs[1][0]) < tol(dt) and rel(outs[0][1], outs[1][1]) < tol(dt)
    dy = make_operand(g, N, C, To, V, dt, 'two')
    outs = []
    for fn, conv in ((ops.maxpool_bwd, real_opnd), (E.maxpool_bwd, lambda o: o)):
        dh = torch.zeros(N, C, T, V, device='cuda', dtype=dt)
        st = torch.zeros(2, C, device='cuda', dtype=torch.float64)
        fn(conv(dy), conv(x), dh, stride, stats=(st[0], st[1]))
        outs.append((dh, st))
    assert rel(outs[0][0], outs[1][0]) < tol(dt) and rel(outs[0][1], outs[1][1]) < tol(dt)


@pytest.mark.parametrize('dt,V,T,K,mode', [(dt, V, T, K, m) for dt in DT
                                           for V, T, K, m in ((25, 30, 3, 'two'), (20, 9, 3, 'two'), (25, 75, 3, 'affine_relu'),
                                                              (25, 300, 3, 'plain'), (20, 52, 2, 'two'), (25, 17, 1, 'two'),
                                                              (25, 16, 3, 'wide'))])
def test_graph_agg(dt, V, T, K, mode):
    """75 / 17 rows of 25 joints: planes start at every 2-byte alignment; 'wide': operands are channel slices."""
    _dev()
    from tam_gcn_b200 import ops
    N, C = 5, 12
    g = gen(11)
    wide = mode == 'wide'
    y = rnd(g, N, K * C + (7 if wide else 0), T, V, dt=dt)[:, :K * C]
    A = rnd(g, K, V, V, scale=0.3)
    outs = []
    for fn in (ops.graph_agg_fwd, E.graph_agg_fwd):
        o = torch.zeros(N, C + (3 if wide else 0), T, V, device='cuda', dtype=dt)[:, :C]
        st = torch.zeros(2, C, device='cuda', dtype=torch.float64)
        fn(y, A, o, stats=(st[0], st[1]))
        outs.append((o, st))
    assert rel(outs[0][0], outs[1][0]) < tol(dt) and rel(outs[0][1], outs[1][1]) < tol(dt)
    go = make_operand(g, N, C, T, V, dt, 'two' if wide else mode, wide=wide)
    outs = []
    for fn, conv in ((ops.graph_agg_bwd, real_opnd), (E.graph_agg_bwd, lambda o: o)):
        dy = torch.zeros(N, K * C + (7 if wide else 0), T, V, device='cuda', dtype=dt)[:, :K * C]
        dA = torch.zeros_like(A)
        fn(conv(go), y, A, dy, dA)
        outs.append((dy, dA))
    assert rel(outs[0][0], outs[1][0]) < tol(dt)
    assert rel(outs[0][1], outs[1][1]) < max(tol(dt), 1e-4)


def test_errors_are_loud():
    _dev()
    from tam_gcn_b200 import ops
    x = torch.zeros(2, 4, 6, 17, device='cuda')
    with pytest.raises(RuntimeError, match='V=17'):
        ops.ctrgc_fwd(x, torch.zeros(2, 8, 1, 17, device='cuda'), torch.zeros(2, 8, 1, 17, device='cuda'),
                      torch.zeros(1, 4, 8, device='cuda'), torch.zeros(1, 4, device='cuda'),
                      torch.zeros(1, 17, 17, device='cuda'), torch.ones(1, device='cuda'), torch.zeros_like(x))
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.mean_t(torch.zeros(1, 2, 3, 20), torch.zeros(1, 2, 1, 20))


def test_batched_weight_pack_matches_per_matrix_pack():
    """tamgcn_conv_pack_weights_batched (one launch for a whole model) writes the same tiles as one
    tamgcn_conv_pack_weights call per matrix, and PackCache serves them without re-packing once fresh."""
    dev = _dev()
    import torch.nn as nn
    from tam_gcn_b200 import _C, ops
    torch.manual_seed(3)
    shapes = [(64, 3, 1), (128, 64, 5), (256, 128, 1), (40, 24, 9), (64, 64, 3)]
    model = nn.ParameterList([nn.Parameter(torch.randn(co, ci * k, device=dev)) for co, ci, k in shapes])
    cache = ops.PackCache(model)
    ref = []
    for W, (co, ci, k) in zip(model, shapes):
        ref.append(ops.conv_pack_weights(W, co, ci, k))
        cache.get(W, co, ci, k, 1, 0)
    assert len(cache.entries) == len(shapes)
    with torch.no_grad():
        for W in model:
            W.mul_(-0.5)
    ref = [ops.conv_pack_weights(W, co, ci, k) for W, (co, ci, k) in zip(model, shapes)]
    cache.repack_all()
    n0 = _C.launch_count()
    for W, (co, ci, k), (rf, rd) in zip(model, shapes, ref):
        wf, wd = cache.get(W, co, ci, k, 1, 0)
        assert torch.equal(wf, rf) and torch.equal(wd, rd)
    assert _C.launch_count() == n0                   # fresh: no per-matrix launch
    cache.end_step()
    cache.get(model[0], *shapes[0], 1, 0)
    assert _C.launch_count() == n0 + 1               # not fresh: the matrix handed out is re-packed
    # a matrix that is not parameter memory is packed but never cached
    tmp = torch.randn(32, 16, device=dev)
    cache.get(tmp, 32, 16, 1, 1, 0)
    assert len(cache.entries) == len(shapes)
