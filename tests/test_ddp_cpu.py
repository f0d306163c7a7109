"""World-size-2 data-parallel logic of tam_gcn_b200.engine on CPU (gloo): parameter broadcast through the flat buffer,
in-place all-reduce of the flat gradient buffer (1/world folded into the optimiser), sharded batch.

Two models stand in for the GPU run: a plain torch MLP (the engine is model-agnostic) and the B200-native CTR-GCN
with its CUDA entry points replaced by their pure-torch emulations (tests/emu_ops.py) — so the direct-to-bucket
gradient path of tam_gcn_b200.functional is what gets averaged.  The N-rank result must equal a single process that
averages the per-rank losses (per-rank BatchNorm statistics, the reference's nn.DataParallel semantics,
processor/io.py:85-87)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _mlp():
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))


def _gcn_block_model(seed):
    """A small but complete network on the B200-native layers: data_bn -> two TCN_GCN_units -> pooled classifier."""
    import tam_gcn_b200.ctrgcn as C
    from tam_gcn_b200 import functional as Fn
    from tam_gcn_b200.graph import ucla

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            A = ucla.Graph().A
            self.num_point = 20
            self.data_bn = torch.nn.BatchNorm1d(3 * 20)
            self.l1 = C.TCN_GCN_unit(3, 16, A, residual=False)
            self.l2 = C.TCN_GCN_unit(16, 16, A)
            self.l3 = C.TCN_GCN_unit(16, 32, A, stride=2)
            self.l4 = C.TCN_GCN_unit(32, 32, A)
            self.fc = torch.nn.Linear(32, 4)

        def forward(self, x):
            x = Fn.DataBnFn.apply(x, self.data_bn, 20, False, torch.float32, self.data_bn.weight, self.data_bn.bias)
            x = self.l4(self.l3(self.l2(self.l1(x))))
            return Fn.PoolFcFn.apply(x, 1, self.fc.weight, self.fc.bias)

    torch.manual_seed(seed)
    m = Net()
    with torch.no_grad():                                   # wake up the paths the reference init leaves dead
        for k, p in m.named_parameters():
            if k.endswith('alpha'):
                p.fill_(0.7)
            elif k.endswith('offset_conv.0.weight'):
                p.normal_(0, 0.05)
            elif k.endswith('gcn1.bn.weight'):
                p.fill_(1.0)
    return m


def _data(kind):
    g = torch.Generator().manual_seed(7)
    if kind == 'mlp':
        return torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
    return torch.randn(4, 3, 8, 20, 1, generator=g), torch.randint(0, 4, (4,), generator=g)


def _worker(rank, world, port, kind, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import emu_ops
    emu_ops.install()
    from tam_gcn_b200 import engine
    model = _mlp() if kind == 'mlp' else _gcn_block_model(100 + rank)   # different weights per rank: broadcast must fix that
    if kind == 'mlp':
        torch.manual_seed(100 + rank)
        model = _mlp()
    tr = engine.Trainer(model, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, use_graph=False)
    x, y = _data(kind)
    n = x.shape[0] // world
    xs, ys = x[rank * n:(rank + 1) * n], y[rank * n:(rank + 1) * n]      # batch sharded over ranks
    for _ in range(3):
        tr.step(xs, ys)
    out[rank] = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    dist.destroy_process_group()


def _run(kind):
    world, port = 2, 29000 + os.getpid() % 2000 + (7 if kind == 'gcn' else 0)
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, kind, out), nprocs=world, join=True)
    assert torch.equal(out[0], out[1]), 'ranks diverged'
    return out[0]


def test_two_rank_data_parallel_matches_single_process():
    got = _run('mlp')
    torch.manual_seed(100)
    model = _mlp()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    x, y = _data('mlp')
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(got, ref, atol=1e-6, rtol=1e-5)


def test_two_rank_gcn_gradients_match_single_process(monkeypatch):
    """The B200-native layers (emulated kernels) under the 2-rank Trainer == one process that averages the two
    per-rank losses with per-rank BatchNorm statistics, through torch autograd + torch.optim.SGD."""
    got = _run('gcn')
    sys.path.insert(0, HERE)
    import emu_ops
    emu_ops.install(monkeypatch)
    model = _gcn_block_model(100)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    x, y = _data('gcn')
    for _ in range(3):
        opt.zero_grad()
        loss = 0.5 * (torch.nn.functional.cross_entropy(model(x[:2]), y[:2]) +
                      torch.nn.functional.cross_entropy(model(x[2:]), y[2:]))
        loss.backward()
        opt.step()
    ref = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    err = float((got - ref).norm() / ref.norm())
    assert err < 1e-5, err
