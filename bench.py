#!/usr/bin/env python
"""bench.py — CTR-GCN (NW-UCLA shape) training-step throughput on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype bf16|f32] [--batch 64]

Headline metric (BASELINE.json): CTR-GCN fwd+bwd samples/s, NW-UCLA shape (C=3, T=52, V=20, M=1, 10 classes),
batch 64 per GPU, bf16 activations (fp32 master weights / accumulation / BN statistics), data parallel over N GPUs.
One step = forward + cross-entropy + backward + (N>1: gradient all-reduce) + SGD-nesterov update, replayed as one
CUDA graph.  Prints ONE JSON line (rank 0).

  value       device-resident inputs, CUDA-event timed per step, L2 flushed between steps (outside the events)
  e2e         the same step through tam_gcn_b200.engine.Trainer.step_from_host: pinned host batch -> H2D -> step
              -> D2H loss read every step (host wall clock between synchronize())
  roofline    the fused CTRGC forward kernel at an HBM-resident size (CUDA events on the launching stream)
  cpu_baseline  the CPU oracle (functional restatement of the reference, same ATen CPU kernels) on the host cores

`--impl reference` times that CPU path alone (the reference has no GPU kernels of its own to run).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

UCLA = dict(num_class=10, num_point=20, num_person=1, graph='graph.ucla.Graph', graph_args=dict(labeling_mode='spatial'))
SHAPE = dict(C=3, T=52, V=20, M=1)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--batch', type=int, default=64, help='samples per GPU')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    return ap.parse_args()


def synthetic_batch(n, seed, device='cpu'):
    """randn*0.5 clipped to [-1,1] (mimics the feeder's min-max output, feeder/feeder_nucla_gcn.py:103-105)."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, SHAPE['C'], SHAPE['T'], SHAPE['V'], SHAPE['M'], generator=g) * 0.5).clamp_(-1, 1)
    y = torch.randint(0, UCLA['num_class'], (n,), generator=g)
    return x.to(device), y.to(device)


def perturb_(named_params, seed=0):
    """Reference init leaves alpha=0, offset conv=0, unit_gcn.bn.weight=1e-6: dead paths would make
    the timed work unrepresentative of a trained model, so give them trained-like magnitudes."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in named_params:
            if k.endswith('gcn1.alpha'):
                p.fill_(0.7)
            elif k.endswith('offset_conv.0.weight'):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif k.endswith('gcn1.bn.weight'):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))


# ------------------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference) — the cpu_baseline leg and the --impl reference arm
# ------------------------------------------------------------------------------------------------------------
def cpu_train_steps(batch, steps, warmup, threads):
    from oracle import gcn_oracle as O          # bench.py's cpu legs are the one place allowed to execute oracle/
    from tam_gcn_b200.graph import ucla
    import torch.nn.functional as F
    torch.set_num_threads(threads)
    A = ucla.Graph().A
    p = O.clone_state(O.make_ctrgcn_state(A, UCLA['num_class'], 1, seed=0), torch.float32, requires_grad=True)
    params = [v for v in p.values() if v.requires_grad]
    mom = [torch.zeros_like(v) for v in params]
    x, y = synthetic_batch(batch, 0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.ctrgcn_forward(x, p, UCLA['num_point'], train=True)
        loss = F.cross_entropy(out, y)
        grads = torch.autograd.grad(loss, params)
        with torch.no_grad():                    # SGD nesterov, lr 0.1, wd 1e-4 (config/nucla/gcn.yaml:29-41)
            for v, g, m in zip(params, grads, mom):
                g = g.add(v, alpha=1e-4)
                m.mul_(0.9).add_(g)
                v.add_(g.add(m, alpha=0.9), alpha=-0.1)
        float(loss.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    batch = args.batch
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    times = cpu_train_steps(batch, steps, warmup, cores)
    ms = 1e3 * sum(times) / len(times)
    val = batch / (ms / 1e3)
    sample = '%d fwd+bwd+SGD steps at batch %d after %d warm-up, fp32, %d threads' % (steps, batch, warmup, cores)
    line = dict(metric='ctrgcn_nucla_train_samples_per_s', value=val, unit='samples/s', n_gpus=args.gpus, steps=steps,
                warmup=warmup, ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                data='synthetic', impl='reference',
                config=dict(workload='CTR-GCN NW-UCLA training fwd+bwd, batch %d (reference CPU path)' % batch,
                            batch_per_gpu=batch, T=52, V=20, M=1, num_class=10),
                cpu_baseline=dict(value=val, unit='samples/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=val, unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------------
# roofline of the fused CTRGC forward kernel at an HBM-resident size
# ------------------------------------------------------------------------------------------------------------
def ctrgc_roofline(dtype, peak_gbs, peak_src):
    from tam_gcn_b200 import ops
    dev = torch.device('cuda')
    # the headline (NW-UCLA) l2-l4 block shape at an HBM-resident batch: N'=2048, Cout=64, T=52, V=20, K=3, R=8
    N, Cout, T, V, K, R = 2048, 64, 52, 20, 3, 8
    s = 2 if dtype == torch.bfloat16 else 4
    g = torch.Generator(device='cuda').manual_seed(0)
    x3 = torch.randn(N, K * Cout, T, V, device=dev, generator=g).to(dtype)
    x12 = torch.randn(N, 2 * K * R, 1, V, device=dev, generator=g)
    W4 = torch.randn(K, Cout, R, device=dev, generator=g) * R ** -0.5
    b4 = torch.zeros(K, Cout, device=dev)
    PA = torch.rand(K, V, V, device=dev, generator=g) * 0.2
    alpha = torch.full((1,), 0.7, device=dev)
    y = torch.empty(N, Cout, T, V, device=dev, dtype=dtype)
    st = torch.zeros(2, Cout, device=dev, dtype=torch.float64)
    # algorithmic bytes per launch: read K x3 planes once, write y once, read fp32 x1/x2, weights
    alg = s * N * T * V * Cout * (K + 1) + K * 8 * N * R * V + K * 4 * (Cout * R + Cout + V * V)
    run = lambda: ops.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    iters = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    ach = alg / (ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape, ncu --set full
    # (profiles/r01s_ctrgc_tc3_full.txt: 826.36 MB read + 252.31 MB written)
    traffic = 1078670080 if dtype == torch.bfloat16 else None
    return dict(bound='hbm', kernel='ctrgc_fwd_tc3_kernel (tcgen05 + mma.sync)' if dtype == torch.bfloat16 else 'ctrgc_fwd_kernel',
                achieved=ach, peak=peak_gbs, unit='GB/s', frac=ach / peak_gbs,
                traffic=traffic, peak_source=peak_src, algorithmic_bytes=alg, ms_per_launch=ms,
                shape=dict(N=N, Cout=Cout, T=T, V=V, K=K, R=R, dtype=str(dtype).replace('torch.', '')),
                note='inputs %.2f GB > 126 MB L2; back-to-back launches' % (x3.numel() * s / 1e9))


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json (measured)'
    except Exception:
        return 6650.0, 'B200_PROFILING.md fallback'


# ------------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if os.environ.get('BENCH_HANG_DUMP'):          # debugging aid: dump all Python stacks and exit if the run wedges
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['BENCH_HANG_DUMP']), exit=True)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        reference_arm(args, rank)
        return
    import torch.distributed as dist
    import tam_gcn_b200
    from tam_gcn_b200 import _C, ctrgcn, engine
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the B200-native path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    tam_gcn_b200.set_act_dtype(dtype)

    torch.manual_seed(0)
    model = ctrgcn.Model(**UCLA)
    perturb_(model.named_parameters())
    model = model.to(dev).train()
    trainer = engine.Trainer(model, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, use_graph=not args.no_graph)
    B = args.batch
    xh, yh = synthetic_batch(B, 1000 + rank)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    x, y = xh.to(dev), yh.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) --------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        loss = trainer.step(x, y)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = _C.launch_count()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush.zero_()
        e0.record()
        loss = trainer.step(x, y)
        e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    ms_per_step = total_ms / args.steps
    value = B * world * args.steps / (total_ms * 1e-3)
    launches = trainer.captured_launches * args.steps if trainer.graph is not None else _C.launch_count() - l0
    final_loss = float(loss)

    # ---- end to end through the public API, host buffers ----------------------------------------------------
    for _ in range(3):
        trainer.step_from_host(xh, yh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        trainer.step_from_host(xh, yh)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / float(e2e_s)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        roof = None
        if not args.no_roofline:
            roof = ctrgc_roofline(dtype, peak, peak_src)
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ts = cpu_train_steps(B, 2, 1, cores)
            cpu = dict(value=B / (sum(ts) / len(ts)), unit='samples/s', cores=cores, kind='port',
                       sample='2 fwd+bwd+SGD steps at batch %d after 1 warm-up, fp32, %d threads' % (B, cores))
        line = dict(metric='ctrgcn_nucla_train_samples_per_s', value=value, unit='samples/s', n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype=args.dtype, data='synthetic',
                    config=dict(workload='CTR-GCN NW-UCLA training fwd+bwd+SGD (BASELINE.json configs[1])',
                                batch_per_gpu=B, global_batch=B * world, C=3, T=52, V=20, M=1, num_class=10,
                                parallelism='dp%d' % world, cuda_graph=trainer.graph is not None,
                                l2='flushed between steps (256 MiB memset outside the timed CUDA events)',
                                weights='reference init, dead paths perturbed (alpha=0.7, offset conv N(0,0.05))'),
                    e2e=dict(value=e2e_val, unit='samples/s', h2d_bytes_per_step=xh.numel() * 4 + yh.numel() * 8,
                             d2h_bytes_per_step=4),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu,
                    loss=final_loss, wall_s_timed_region=t_wall)
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down: the captured graph holds NCCL work; release it first, and never let a wedged communicator
        # tear-down keep the (already printed) run alive
        sys.stdout.flush()
        dist.barrier()
        trainer.graph = None
        del trainer
        import gc
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Timer(15.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
