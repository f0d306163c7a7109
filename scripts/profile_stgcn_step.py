"""One eager ST-GCN (NTU RGB+D graph, T=300, V=25, M=2) training step between cudaProfilerStart/Stop (ncu launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tam_gcn_b200
from tam_gcn_b200 import stgcn, engine

tam_gcn_b200.set_act_dtype(torch.bfloat16)
torch.manual_seed(0)
sm = stgcn.Model(in_channels=3, num_class=60, num_point=25, num_person=1, graph='graph.ntu_rgb_d.Graph',
                 graph_args=dict(labeling_mode='spatial'), edge_importance_weighting=True).cuda().train()
tr = engine.Trainer(sm, use_graph=False)
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = (torch.randn(bs, 3, 300, 25, 2, device='cuda') * 0.5).clamp_(-1, 1)
y = torch.randint(0, 60, (bs,), device='cuda')
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
