"""World-size-2 data-parallel logic of tam_gcn_b200.engine on CPU (gloo): the flat-bucket gradient average and the
parameter broadcast, with a plain torch model standing in for the CUDA modules (the engine is model-agnostic)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tam_gcn_b200 import engine
    torch.manual_seed(100 + rank)                       # different initial weights per rank: broadcast must fix that
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    tr = engine.Trainer(model, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, use_graph=False, fused=False)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 6, generator=g)
    y = torch.randint(0, 3, (8,), generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]      # batch sharded over ranks
    for _ in range(3):
        tr.step(xs, ys)
    out[rank] = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_single_process():
    world, port = 2, 29000 + os.getpid() % 2000
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.equal(out[0], out[1]), 'ranks diverged'
    # single-process reference on the full batch, starting from rank 0's weights
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 6, generator=g)
    y = torch.randint(0, 3, (8,), generator=g)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(out[0], ref, atol=1e-6, rtol=1e-5)
