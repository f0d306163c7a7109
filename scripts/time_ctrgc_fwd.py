"""CUDA-event timing of the fused CTRGC forward alone (HBM-resident sizes): python scripts/time_ctrgc_fwd.py [ucla|ntu|ucla64] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tam_gcn_b200 import ops

SHAPES = {'ucla': (2048, 64, 52, 20, 3, 8), 'ntu': (1024, 64, 64, 25, 3, 8), 'ucla64': (64, 64, 52, 20, 3, 8),
          'ucla_l6': (2048, 128, 26, 20, 3, 16), 'ucla_l9': (2048, 256, 13, 20, 3, 32)}
name = sys.argv[1] if len(sys.argv) > 1 else 'ucla'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N, Cout, T, V, K, R = SHAPES[name]
dev, dtype = 'cuda', torch.bfloat16
g = torch.Generator(device='cuda').manual_seed(0)
x3 = torch.randn(N, K * Cout, T, V, device=dev, generator=g).to(dtype)
x12 = torch.randn(N, 2 * K * R, 1, V, device=dev, generator=g)
W4 = torch.randn(K, Cout, R, device=dev, generator=g) * R ** -0.5
b4 = torch.zeros(K, Cout, device=dev)
PA = torch.rand(K, V, V, device=dev, generator=g) * 0.2
alpha = torch.full((1,), 0.7, device=dev)
y = torch.empty(N, Cout, T, V, device=dev, dtype=dtype)
st = torch.zeros(2, Cout, device=dev, dtype=torch.float64)
run = lambda: ops.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
alg = 2 * N * T * V * Cout * (K + 1) + K * 8 * N * R * V + K * 4 * (Cout * R + Cout + V * V)
print('%s dbg=%s ctrgc_fwd %.1f us  %.0f GB/s (%.1f%% of 6547.5)' % (name, os.environ.get('TAMGCN_CTC_DBG', '0'), ms * 1e3,
                                                                  alg / ms / 1e6, alg / ms / 1e6 / 65.475))
