import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tam_gcn_b200
from tam_gcn_b200 import ctrgcn
tam_gcn_b200.set_act_dtype(torch.bfloat16)
torch.manual_seed(0)
m = ctrgcn.Model(num_class=60, num_point=25, num_person=2, graph='graph.ntu_rgb_d.Graph', graph_args=dict(labeling_mode='spatial')).cuda().train()
with torch.no_grad():
    for _ in range(2):
        m((torch.randn(16, 3, 64, 25, 2, device='cuda') * 0.5).clamp_(-1, 1))
m.eval()
x = (torch.randn(256, 3, 64, 25, 2, device='cuda') * 0.5).clamp_(-1, 1)
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
