"""The fused CTRGC forward / backward kernels alone at an HBM-resident size, for `ncu --set full`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tam_gcn_b200 import ops
from tam_gcn_b200.ops import Opnd

dtype = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == 'bf16') else torch.float32
N, Cout, T, V, K, R = 1024, 64, 64, 25, 3, 8
if len(sys.argv) > 2 and sys.argv[2] == 'ucla':
    N, Cout, T, V, K, R = 2048, 64, 52, 20, 3, 8
if len(sys.argv) > 2 and sys.argv[2] == 'ucla64':
    N, Cout, T, V, K, R = 64, 64, 52, 20, 3, 8
if len(sys.argv) > 2 and sys.argv[2] == 'l6_64':
    N, Cout, T, V, K, R = 64, 128, 26, 20, 3, 16
if len(sys.argv) > 2 and sys.argv[2] == 'l9_64':
    N, Cout, T, V, K, R = 64, 256, 13, 20, 3, 32
if len(sys.argv) > 2 and sys.argv[2] == 'ntu_l6':
    N, Cout, T, V, K, R = 512, 128, 32, 25, 3, 16
if len(sys.argv) > 2 and sys.argv[2] == 'ntu_l8':
    N, Cout, T, V, K, R = 512, 256, 32, 25, 3, 16
if len(sys.argv) > 2 and sys.argv[2] == 'ntu_l9':
    N, Cout, T, V, K, R = 512, 256, 16, 25, 3, 32
if len(sys.argv) > 2 and sys.argv[2] == 'ntu_l9_64':
    N, Cout, T, V, K, R = 64, 256, 16, 25, 3, 32
if len(sys.argv) > 2 and sys.argv[2] == 'ntu2048':
    N, Cout, T, V, K, R = 2048, 64, 64, 25, 3, 8
dev = 'cuda'
g = torch.Generator(device='cuda').manual_seed(0)
x3 = torch.randn(N, K * Cout, T, V, device=dev, generator=g).to(dtype)
x12 = torch.randn(N, 2 * K * R, 1, V, device=dev, generator=g)
W4 = torch.randn(K, Cout, R, device=dev, generator=g) * R ** -0.5
b4 = torch.zeros(K, Cout, device=dev)
PA = torch.rand(K, V, V, device=dev, generator=g) * 0.2
alpha = torch.full((1,), 0.7, device=dev)
y = torch.empty(N, Cout, T, V, device=dev, dtype=dtype)
st = torch.zeros(2, Cout, device=dev, dtype=torch.float64)
gr = torch.randn(N, Cout, T, V, device=dev, generator=g).to(dtype)
dx3 = torch.empty_like(x3)
dx12 = torch.zeros_like(x12)
dW4, db4, dPA, dal = torch.zeros_like(W4), torch.zeros_like(b4), torch.zeros_like(PA), torch.zeros(1, device=dev)
fwd_only = len(sys.argv) > 3 and sys.argv[3] == 'fwd'
for it in range(3):
    ops.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
    ops.ctrgc_bwd(Opnd(gr), x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, dx3, dx12[:, :K * R],
                  dx12[:, K * R:], dW4, db4, dPA, dal)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
ops.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
e[1].record()
ops.ctrgc_bwd(Opnd(gr), x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, dx3, dx12[:, :K * R], dx12[:, K * R:],
              dW4, db4, dPA, dal)
e[2].record()
torch.cuda.synchronize()
s = 2 if dtype == torch.bfloat16 else 4
bf = s * N * T * V * Cout * (K + 1) + K * 8 * N * R * V
bb = s * N * T * V * Cout * (1 + 2 * K) + 2 * K * 8 * N * R * V
tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
print('ctrgc_fwd %.3f ms  %.1f GB/s | ctrgc_bwd %.3f ms  %.1f GB/s' % (tf, bf / tf / 1e6, tb, bb / tb / 1e6))
