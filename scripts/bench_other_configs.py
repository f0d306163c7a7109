"""Informational measurements of the other BASELINE.json configs (not the bench.py headline):
   cfg4  CTR-GCN NTU-60 shape (V=25, T=64, M=2) eval-mode inference throughput, batch sweep (CUDA-graph Predictor)
   cfg3  ST-GCN on the NTU RGB+D graph (V=25, T=300, M=2, 60 classes) training step (CUDA-graph Trainer)
   python scripts/bench_other_configs.py [max_batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tam_gcn_b200
from tam_gcn_b200 import ctrgcn, stgcn, engine

tam_gcn_b200.set_act_dtype(torch.bfloat16)
dev = 'cuda'
max_b = int(sys.argv[1]) if len(sys.argv) > 1 else 1024


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
m = ctrgcn.Model(num_class=60, num_point=25, num_person=2, graph='graph.ntu_rgb_d.Graph', graph_args=dict(labeling_mode='spatial'))
m = m.to(dev).train()
with torch.no_grad():                                   # calibrate the BatchNorm running statistics before eval
    for _ in range(3):
        m((torch.randn(16, 3, 64, 25, 2, device=dev) * 0.5).clamp_(-1, 1))
m.eval()
b = 256
while b <= max_b:
    x = (torch.randn(b, 3, 64, 25, 2, device=dev) * 0.5).clamp_(-1, 1)
    pred = engine.Predictor(m)
    ms = timed(lambda: pred(x), 5)
    print('cfg4 CTR-GCN NTU-60 shape inference bf16: batch %5d  %8.2f ms  %9.0f samples/s  (%d launches)' % (
        b, ms, b / ms * 1e3, pred.captured_launches), flush=True)
    del pred, x
    torch.cuda.empty_cache()
    b *= 2

torch.manual_seed(0)
# num_person=1 with M=2 data: the reference's data_bn is built for V*C features (SURVEY 8c caveat 6)
sm = stgcn.Model(in_channels=3, num_class=60, num_point=25, num_person=1, graph='graph.ntu_rgb_d.Graph',
                 graph_args=dict(labeling_mode='spatial'), edge_importance_weighting=True)
sm = sm.to(dev).train()
tr = engine.Trainer(sm, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
for bs in (16, 32):
    x = (torch.randn(bs, 3, 300, 25, 2, device=dev) * 0.5).clamp_(-1, 1)
    y = torch.randint(0, 60, (bs,), device=dev)
    tr.graph = None
    ms = timed(lambda: tr.step(x, y), 5)
    print('cfg3 ST-GCN NTU RGB+D training step bf16: batch %3d  %8.2f ms  %8.0f samples/s' % (bs, ms, bs / ms * 1e3), flush=True)
