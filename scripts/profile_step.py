"""One eager (un-graphed) training step (workload argv[3], default ucla_train) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum ...   (launch list of a step)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import tam_gcn_b200
from tam_gcn_b200 import ctrgcn, engine

dtype = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == 'bf16') else torch.float32
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
tam_gcn_b200.set_act_dtype(dtype)
torch.manual_seed(0)
wname = sys.argv[3] if len(sys.argv) > 3 else 'ucla_train'
w = bench.WORKLOADS[wname]
model = bench.build_model(w, 'cuda').train()
tr = engine.Trainer(model, use_graph=False, side_stream=False)
x, y = bench.synthetic_batch(w, batch, 1, 'cuda')
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
loss = tr.step(x, y)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('loss', float(loss))
