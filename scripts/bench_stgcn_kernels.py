"""ST-GCN kernels alone at the NTU batch-16 training shapes (N*M = 32 sequences): graph aggregation fwd / bwd and the
9x1 temporal convolution fwd / dgrad / wgrad.  Prints microseconds per launch and the algorithmic GB/s / TFLOP/s."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tam_gcn_b200 import ops
from tam_gcn_b200.ops import Opnd

dev = 'cuda'
g = torch.Generator(device=dev).manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else 'all'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
SHAPES = [(64, 300), (128, 150), (256, 75)]
K, V = 3, 25


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def rnd(*s, dt=torch.bfloat16):
    return torch.randn(*s, device=dev, generator=g).to(dt)


for C, T in SHAPES:
    if what in ('all', 'agg'):
        y = rnd(N, K * C, T, V)
        A = torch.rand(K, V, V, device=dev, generator=g) * 0.2
        out = torch.empty(N, C, T, V, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(2, C, device=dev, dtype=torch.float64)
        go = Opnd(rnd(N, C, T, V), rnd(N, C, T, V), a=torch.rand(C, device=dev), b=torch.rand(C, device=dev), c=torch.rand(C, device=dev))
        dy = torch.empty_like(y)
        dA = torch.zeros_like(A)
        tf = timeit(lambda: ops.graph_agg_fwd(y, A, out, stats=(st[0], st[1])))
        tb = timeit(lambda: ops.graph_agg_bwd(go, y, A, dy, dA))
        bf = 2 * (y.numel() + out.numel())
        bb = 2 * (2 * y.numel() + 2 * out.numel())
        print('agg  C=%3d T=%3d  fwd %7.1f us %6.0f GB/s | bwd(dy+dA) %7.1f us %6.0f GB/s' % (C, T, tf, bf / tf / 1e3, tb, bb / tb / 1e3))
    if what in ('all', 'conv'):
        k, pad = 9, 4
        x = rnd(N, C, T, V)
        W = torch.randn(C, C * k, device=dev, generator=g) * (C * k) ** -0.5
        b = torch.zeros(C, device=dev)
        yo = torch.empty(N, C, T, V, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(2, C, device=dev, dtype=torch.float64)
        wf, wd = ops.conv_pack_weights(W, C, C, k, 1, V)
        xin = Opnd(x, a=torch.rand(C, device=dev), c=torch.rand(C, device=dev), relu=True)
        gy = Opnd(rnd(N, C, T, V), rnd(N, C, T, V), a=torch.rand(C, device=dev), b=torch.rand(C, device=dev), c=torch.rand(C, device=dev))
        dx = torch.empty_like(x)
        dW = torch.zeros_like(W)
        db = torch.zeros(C, device=dev)
        tf = timeit(lambda: ops.conv_fwd(xin, W, b, yo, k, 1, 1, pad, stats=(st[0], st[1]), wpack=wf))
        td = timeit(lambda: ops.conv_dgrad(gy, W, dx, k, 1, 1, pad, wpack=wd))
        tw = timeit(lambda: ops.conv_wgrad(gy, xin, dW, db, k, 1, 1, pad))
        fl = 2.0 * N * T * V * C * C * k
        print('conv C=%3d T=%3d  fwd %7.1f us %5.1f TF/s | dgrad %7.1f us %5.1f TF/s | wgrad %7.1f us %5.1f TF/s' %
              (C, T, tf, fl / tf / 1e6, td, fl / td / 1e6, tw, fl / tw / 1e6))

if what in ('all', 'gcn'):
    # the 1x1 graph convolution Cin -> K*Cout of every layer (models/stgcn.py:47-53)
    for Cin, Cout, T in ((3, 64, 300), (64, 64, 300), (64, 128, 300), (128, 128, 150), (128, 256, 150), (256, 256, 75)):
        KC = K * Cout
        x = rnd(N, Cin, T, V)
        W = torch.randn(KC, Cin, device=dev, generator=g) * Cin ** -0.5
        b = torch.zeros(KC, device=dev)
        yo = torch.empty(N, KC, T, V, device=dev, dtype=torch.bfloat16)
        wf, wd = ops.conv_pack_weights(W, KC, Cin, 1, 1, V)
        gy = rnd(N, KC, T, V)
        dx = torch.empty_like(x)
        dW = torch.zeros_like(W)
        db = torch.zeros(KC, device=dev)
        tf = timeit(lambda: ops.conv_fwd(x, W, b, yo, 1, 1, 1, 0, wpack=wf))
        td = timeit(lambda: ops.conv_dgrad(gy, W, dx, 1, 1, 1, 0, wpack=wd))
        tw = timeit(lambda: ops.conv_wgrad(gy, x, dW, db, 1, 1, 1, 0))
        by = 2.0 * (x.numel() + yo.numel())
        print('gcn  %3d->%3d T=%3d  fwd %7.1f us %5.0f GB/s | dgrad %7.1f us %5.0f GB/s | wgrad %7.1f us %5.0f GB/s' %
              (Cin, KC, T, tf, by / tf / 1e3, td, by / td / 1e3, tw, by / tw / 1e3))
