import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tam_gcn_b200 import ops
N, Cout, T, V, K, R = 512, 256, 16, 25, 3, 32
dev, dtype = 'cuda', torch.bfloat16
g = torch.Generator(device='cuda').manual_seed(0)
x3 = torch.randn(N, K * Cout, T, V, device=dev, generator=g).to(dtype)
x12 = torch.randn(N, 2 * K * R, 1, V, device=dev, generator=g)
W4 = torch.randn(K, Cout, R, device=dev, generator=g) * R ** -0.5
b4 = torch.zeros(K, Cout, device=dev)
PA = torch.rand(K, V, V, device=dev, generator=g) * 0.2
alpha = torch.full((1,), 0.7, device=dev)
y = torch.empty(N, Cout, T, V, device=dev, dtype=dtype)
for _ in range(3):
    ops.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y)
torch.cuda.synchronize()
