"""Data-parallel training on 2 real GPUs over NCCL (skipped with fewer than 2 devices; run with `gpurun --gpus 2`).

Each rank steps `engine.Trainer` on its shard of the batch (CUDA graph, weight gradients on the side stream, upper-layer
all-reduce launched from inside backward).  Checks:
  * after a step with lr = 0 the all-reduced gradient buffer (x 1/world) equals the single-process average of the
    per-shard gradients computed through plain autograd on one GPU (per-rank BatchNorm statistics — the reference's
    nn.DataParallel semantics, processor/io.py:85-87);
  * after further steps all ranks hold bit-identical parameters, and they match a single-process run that averages
    the per-shard gradients."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, use_graph, overlap, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import helpers as H
    import tam_gcn_b200
    from tam_gcn_b200 import engine
    from oracle import gcn_oracle as O
    tam_gcn_b200.set_act_dtype(torch.float32)
    m = H.fresh_ctrgcn(10 + rank).to(dev).train()                      # different weights per rank: broadcast fixes it
    tr = engine.Trainer(m, lr=0.0, use_graph=use_graph, overlap_allreduce=overlap)
    x = O.synthetic_skeletons(8 * world, 52, 20, 1, C=3, seed=4)
    y = torch.randint(0, 10, (8 * world,), generator=torch.Generator().manual_seed(4))
    xs, ys = x[rank * 8:(rank + 1) * 8].to(dev), y[rank * 8:(rank + 1) * 8].to(dev)
    tr.step(xs, ys)                                                    # lr = 0: parameters unchanged, G = sum of gradients
    torch.cuda.synchronize()
    G1 = (tr.store.G / world).cpu().clone()
    P0 = tr.store.P.cpu().clone()
    tr.set_lr(0.01)
    for _ in range(2):
        tr.step(xs, ys)
    torch.cuda.synchronize()
    res = dict(G1=G1, P0=P0, P=tr.store.P.cpu().clone(), split=tr._split, launches=tr.captured_launches)
    if rank == 0:
        # single-process reference on this GPU: same initial parameters (rank 0's), plain autograd per shard
        ref = H.fresh_ctrgcn(10).to(dev).train()
        from tam_gcn_b200.params import ParamStore
        st = ParamStore(ref)                                            # same flat layout, for comparison only
        ref.zero_grad(set_to_none=True)
        loss = 0
        for r in range(world):
            loss = loss + torch.nn.functional.cross_entropy(ref(x[r * 8:(r + 1) * 8].to(dev)), y[r * 8:(r + 1) * 8].to(dev))
        (loss / world).backward()
        Gref = torch.zeros_like(st.G)
        for _, p, o in st.order:
            if p.grad is not None:
                Gref[o:o + p.numel()] = p.grad.reshape(-1)
        res['Gref'] = Gref.cpu()
    out[rank] = res
    dist.barrier()
    tr.graph = None
    dist.destroy_process_group()


@pytest.mark.parametrize('use_graph,overlap', [(False, True), (True, True), (True, False)])
def test_two_gpu_nccl_gradients_and_parameters(use_graph, overlap):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    import torch.multiprocessing as mp
    world, port = 2, 29500 + os.getpid() % 1000 + (3 if use_graph else 0) + (5 if overlap else 0)
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, use_graph, overlap, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert torch.equal(a['P0'], b['P0']), 'broadcast failed'
    assert torch.equal(a['G1'], b['G1']), 'all-reduced gradients differ between ranks'
    assert torch.equal(a['P'], b['P']), 'ranks diverged'
    if overlap:
        assert a['split'] is not None and 0 < a['split'] < a['G1'].numel()
    Gref = a['Gref']
    mask = Gref.abs() > 0
    e = float((a['G1'] - Gref)[mask].double().norm() / Gref[mask].double().norm())
    print('2-GPU averaged gradient vs single-process: rel err %.2e (graph=%s overlap=%s, %d launches)' %
          (e, use_graph, overlap, a['launches']))
    assert e < 1e-2            # two fp32 summation orders of the end-to-end gradient: the reference itself sits 2e-3 from fp64; 3x that, rounded up
    assert not torch.equal(a['P'], a['P0'])


def test_data_parallel_two_gpus():
    """nn.DataParallel — the reference's only multi-GPU mode (processor/io.py:85-87): one process, two devices, replica
    threads.  Forward + backward in fp32 and bf16 must equal the same model run shard by shard on one device (per-replica
    BatchNorm statistics), which needs every per-device launch cache of libtamgcn.so to be keyed by device."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    sys.path.insert(0, HERE)
    import helpers as H
    import tam_gcn_b200
    from oracle import gcn_oracle as O
    x = O.synthetic_skeletons(8, 52, 20, 1, C=3, seed=6)
    y = torch.randint(0, 10, (8,), generator=torch.Generator().manual_seed(6))
    for act, tol in ((torch.float32, 1e-2), (torch.bfloat16, 0.5)):
        with tam_gcn_b200.act_dtype(act):
            m = H.fresh_ctrgcn(3).to('cuda:0').train()
            dp = torch.nn.DataParallel(m, device_ids=[0, 1])
            out = dp(x.to('cuda:0'))
            assert out.shape == (8, 10)
            torch.nn.functional.cross_entropy(out.float(), y.to('cuda:0')).backward()
            g_dp = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).double().cpu()
            y_dp = out.detach().double().cpu()
            ref = H.fresh_ctrgcn(3).to('cuda:0').train()
            o = torch.cat([ref(x[:4].to('cuda:0')), ref(x[4:].to('cuda:0'))])
            torch.nn.functional.cross_entropy(o.float(), y.to('cuda:0')).backward()
            g_ref = torch.cat([p.grad.reshape(-1) for p in ref.parameters()]).double().cpu()
            e_y = float((y_dp - o.detach().double().cpu()).norm() / o.detach().double().norm())
            e_g = float((g_dp - g_ref).norm() / g_ref.norm())
            print('DataParallel %s: logits %.2e grads %.2e' % (act, e_y, e_g))
            assert e_y < (1e-5 if act == torch.float32 else 3e-2) and e_g < tol
