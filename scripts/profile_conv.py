"""One conv (fwd, dgrad, wgrad) at a CTR-GCN layer shape, CUDA-event timed; target for ncu --set full."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tam_gcn_b200 import ops
from tam_gcn_b200.ops import Opnd

dt = torch.bfloat16
N, Cin, Cout, T, V, k = 64, 64, 192, 52, 20, 1
if len(sys.argv) > 1:
    N, Cin, Cout, T, V, k = [int(a) for a in sys.argv[1:7]]
pad = (k - 1) // 2
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(N, Cin, T, V, device='cuda', generator=g).to(dt)
a = torch.rand(Cin, device='cuda', generator=g) + 0.5
c = torch.randn(Cin, device='cuda', generator=g) * 0.1
W = torch.randn(Cout, Cin * k, device='cuda', generator=g) * (Cin * k) ** -0.5
b = torch.zeros(Cout, device='cuda')
y = torch.empty(N, Cout, T, V, device='cuda', dtype=dt)
dy = torch.randn(N, Cout, T, V, device='cuda', generator=g).to(dt)
dx = torch.empty_like(x)
st = torch.zeros(2, Cout, device='cuda', dtype=torch.float64)
dW = torch.zeros_like(W)
db = torch.zeros(Cout, device='cuda')
wf, wd = ops.conv_pack_weights(W, Cout, Cin, k)
xo = Opnd(x, a=a, c=c, relu=True)


def run():
    ops.conv_fwd(xo, W, b, y, k, 1, 1, pad, stats=(st[0], st[1]), wpack=wf)
    ops.conv_dgrad(Opnd(dy, y, a=b + 1, b=b, c=b), W, dx, k, 1, 1, pad, wpack=wd)
    ops.conv_wgrad(Opnd(dy, y, a=b + 1, b=b, c=b), xo, dW, db, k, 1, 1, pad)


for _ in range(3):
    run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
ops.conv_fwd(xo, W, b, y, k, 1, 1, pad, stats=(st[0], st[1]), wpack=wf)
ev[1].record()
ops.conv_dgrad(Opnd(dy, y, a=b + 1, b=b, c=b), W, dx, k, 1, 1, pad, wpack=wd)
ev[2].record()
ops.conv_wgrad(Opnd(dy, y, a=b + 1, b=b, c=b), xo, dW, db, k, 1, 1, pad)
ev[3].record()
torch.cuda.synchronize()
byt = 2 * N * T * V * (Cin + Cout)
print('shape N%d %d->%d T%d V%d k%d | fwd %.1f us (%.0f GB/s)  dgrad %.1f us  wgrad %.1f us' % (
    N, Cin, Cout, T, V, k, 1e3 * ev[0].elapsed_time(ev[1]), byt / ev[0].elapsed_time(ev[1]) / 1e6,
    1e3 * ev[1].elapsed_time(ev[2]), 1e3 * ev[2].elapsed_time(ev[3])))
