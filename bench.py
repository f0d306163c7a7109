#!/usr/bin/env python
"""bench.py — throughput of the B200-native CTR-GCN / ST-GCN hot path, with roofline and the reference's own numbers.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ucla_train|...]
                    [--dtype bf16|f32] [--batch B]

Workloads (BASELINE.json `configs`; the default is the one the headline metric is quoted on):
  ucla_train   configs[1]  CTR-GCN NW-UCLA (C=3, T=52, V=20, M=1, 10 classes) training step, batch 64 per GPU
  ntu_train    north_star  CTR-GCN NTU-60 shape (T=64, V=25, M=2, 60 classes) training step, batch 32 per GPU
  ntu_infer    configs[3]  CTR-GCN NTU-60 shape eval-mode forward, batch 256 per GPU (use --batch for the sweep)
  stgcn_train  configs[2]  ST-GCN on the NTU RGB+D graph (T=300, V=25, M=2, 60 classes) training step, batch 16 per GPU
  fusion_gcn   configs[4]  GCN branch of the cross-modal fusion model (frozen CTR-GCN + attention MLP), batch 64 per GPU

One training step = data_bn prologue + blocks forward + pooled classifier + cross-entropy + hand-written backward +
(N>1: NCCL gradient all-reduce, overlapped with backward) + fused SGD-nesterov, replayed as ONE CUDA graph.
Prints ONE JSON line (rank 0):

  value          device-resident inputs, CUDA-event timed per step, L2 flushed between steps (outside the events)
  e2e            the same step through the public API from PINNED HOST buffers: H2D copy of the batch -> step -> D2H of
                 the loss (logits for inference) inside the timed region, every step
  roofline       the fused CTRGC forward kernel (BASELINE metric "CTRGC HBM GB/s") at an HBM-resident size of the
                 headline shape, CUDA events on the launching stream, against MEASURED_PEAKS.json
  roofline_ntu   the same kernel at the NTU-60 shape (V=25) of SURVEY.md §8(d);  roofline_bwd: the fused CTRGC backward
  roofline_step  per kernel family of the timed step: device time per step (torch.profiler/CUPTI over graph replays —
                 a breakdown, never the headline), algorithmic bytes / FLOPs at the TRAINING shape, GB/s, TFLOP/s
  gpu_eager_baseline  the reference math (oracle port = the reference's ATen calls) in stock PyTorch on the SAME GPU:
                 fp32 and autocast(bf16), eager and as a CUDA graph — the number the hand-written kernels must beat
  cpu_baseline   the reference's CPU path on the host cores (N=1 only)

`--impl reference` times the reference's CPU path alone (the reference has no GPU kernels of its own).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the CPU arm wants every host core; torchrun exports OMP_NUM_THREADS=1, which must be undone BEFORE torch loads
    for _v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count() or 1)

import re  # noqa: E402
import statistics  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import torch  # noqa: E402

UCLA = dict(num_class=10, num_point=20, num_person=1, graph='graph.ucla.Graph', graph_args=dict(labeling_mode='spatial'))
NTU = dict(num_class=60, num_point=25, num_person=2, graph='graph.ntu_rgb_d.Graph', graph_args=dict(labeling_mode='spatial'))
STGCN = dict(in_channels=3, num_class=60, num_point=25, num_person=1, graph='graph.ntu_rgb_d.Graph',
             graph_args=dict(labeling_mode='spatial'))

WORKLOADS = {
    'ucla_train': dict(metric='ctrgcn_nucla_train_samples_per_s', family='ctrgcn', cfg=UCLA, C=3, T=52, V=20, M=1, batch=64, train=True,
                       desc='CTR-GCN NW-UCLA training fwd+bwd+SGD (BASELINE.json configs[1])'),
    'ntu_train': dict(metric='ctrgcn_ntu_train_samples_per_s', family='ctrgcn', cfg=NTU, C=3, T=64, V=25, M=2, batch=32, train=True,
                      desc='CTR-GCN NTU-60 shape training fwd+bwd+SGD (north_star NTU-shaped throughput)'),
    'ntu_infer': dict(metric='ctrgcn_ntu_infer_samples_per_s', family='ctrgcn', cfg=NTU, C=3, T=64, V=25, M=2, batch=256, train=False,
                      desc='CTR-GCN NTU-60 shape eval-mode inference (BASELINE.json configs[3])'),
    'stgcn_train': dict(metric='stgcn_ntu_train_samples_per_s', family='stgcn', cfg=STGCN, C=3, T=300, V=25, M=2, batch=16, train=True,
                        desc='ST-GCN NTU RGB+D graph training fwd+bwd+SGD (BASELINE.json configs[2])'),
    'fusion_gcn': dict(metric='fusion_gcn_branch_samples_per_s', family='fusion', cfg=UCLA, C=3, T=52, V=20, M=1, batch=64, train=False,
                       desc='GCN branch of the cross-modal fusion model: frozen CTR-GCN features -> attention gate '
                            '(BASELINE.json configs[4])'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='ucla_train', choices=sorted(WORKLOADS))
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--batch', type=int, default=0, help='samples per GPU (0 = the workload default)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-baseline', action='store_true')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--no-step-roofline', action='store_true')
    ap.add_argument('--no-side-stream', action='store_true')
    ap.add_argument('--no-overlap', action='store_true')
    return ap.parse_args()


def synthetic_batch(w, n, seed, device='cpu'):
    """randn*0.5 clipped to [-1,1] (mimics the feeder's min-max output, feeder/feeder_nucla_gcn.py:103-105)."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, w['C'], w['T'], w['V'], w['M'], generator=g) * 0.5).clamp_(-1, 1)
    y = torch.randint(0, w['cfg']['num_class'], (n,), generator=g)
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


def config_of(w, name, batch, world, **extra):
    c = dict(workload=w['desc'], name=name, batch_per_gpu=batch, global_batch=batch * world, C=w['C'], T=w['T'], V=w['V'],
             M=w['M'], num_class=w['cfg']['num_class'], parallelism='dp%d' % world)
    c.update(extra)
    return c


def build_model(w, device):
    from tam_gcn_b200 import ctrgcn, stgcn
    torch.manual_seed(0)
    if w['family'] == 'stgcn':
        model = stgcn.Model(**w['cfg'])
    elif w['family'] == 'fusion':
        from tam_gcn_b200 import fusion
        model = fusion.GcnAttentionBranch(**w['cfg'])
    else:
        model = ctrgcn.Model(**w['cfg'])
    perturb_(model.named_parameters())
    return model.to(device)


# ------------------------------------------------------------------------------------------------------------
# the reference math in stock PyTorch (oracle port): CPU arm / cpu_baseline leg / GPU eager baseline
# ------------------------------------------------------------------------------------------------------------
def _oracle_state(w, device, dtype=torch.float32):
    from oracle import gcn_oracle as O          # bench.py's baseline legs are the one place allowed to execute oracle/
    model = build_model(w, 'cpu')
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    p = O.clone_state(state, dtype, requires_grad=False)
    for k, v in p.items():
        if torch.is_tensor(v):
            p[k] = v.to(device)
            if v.is_floating_point() and not k.endswith(('running_mean', 'running_var')) and k != 'A':
                p[k].requires_grad_(True)
    return p


def _oracle_forward(w):
    from oracle import gcn_oracle as O
    V = w['V']
    if w['family'] == 'stgcn':
        return lambda x, p, train: O.stgcn_forward(x, p, V, train)
    return lambda x, p, train: O.ctrgcn_forward(x, p, V, train)


def cpu_steps(w, batch, steps, warmup, threads):
    """The reference's CPU path (same ATen CPU kernels the reference dispatches to): fwd+bwd+SGD or eval forward."""
    import torch.nn.functional as F
    torch.set_num_threads(threads)
    p = _oracle_state(w, 'cpu')
    fwd = _oracle_forward(w)
    params = [v for v in p.values() if torch.is_tensor(v) and v.requires_grad]
    mom = [torch.zeros_like(v) for v in params]
    x, y = synthetic_batch(w, batch, 0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if w['train']:
            loss = F.cross_entropy(fwd(x, p, True), y)
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            with torch.no_grad():                    # SGD nesterov, lr 0.1, wd 1e-4 (config/nucla/gcn.yaml:29-41)
                for v, g, m in zip(params, grads, mom):
                    if g is None:
                        continue
                    g = g.add(v, alpha=1e-4)
                    m.mul_(0.9).add_(g)
                    v.add_(g.add(m, alpha=0.9), alpha=-0.1)
            float(loss.detach())
        else:
            with torch.no_grad():
                fwd(x, p, False).sum().item()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def reference_arm(args, rank, world):
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    if w['family'] == 'fusion':
        w = dict(w, family='ctrgcn')               # the GCN branch dominates; its CPU path is the CTR-GCN forward
    cores = os.cpu_count() or 1
    batch = args.batch or w['batch']
    steps, warmup = max(1, min(args.steps, 40)), max(1, min(args.warmup, 5))
    per_step = 1.0 if args.workload == 'ucla_train' else 8.0           # rough seconds per step on 16 cores
    while steps > 2 and (steps + warmup) * per_step > 240:             # keep the arm within a few minutes
        steps //= 2
        warmup = min(warmup, 2)
    times = cpu_steps(w, batch, steps, warmup, cores)
    ms = 1e3 * sum(times) / len(times)
    val = batch / (ms / 1e3)
    what = 'fwd+bwd+SGD steps' if w['train'] else 'eval forwards'
    sample = '%d %s at batch %d after %d warm-up, fp32, %d threads' % (steps, what, batch, warmup, cores)
    line = dict(metric=w['metric'], value=val, unit='samples/s', n_gpus=args.gpus, steps=steps, warmup=warmup,
                ms_per_step=ms, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                impl='reference', config=config_of(w, args.workload, batch, 1, note='reference CPU path (oracle port)'),
                cpu_baseline=dict(value=val, unit='samples/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=val, unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def gpu_eager_baseline(w, batch, dev, steps=10):
    """Stock PyTorch (cuDNN / cuBLAS, default flags) running the reference math on this GPU."""
    import torch.nn.functional as F
    fwd = _oracle_forward(w)
    x, y = synthetic_batch(w, batch, 0, dev)
    out = {}
    for tag, autocast in (('fp32', False), ('autocast_bf16', True)):
        p = _oracle_state(w, dev)
        params = [v for v in p.values() if torch.is_tensor(v) and v.requires_grad]
        opt = torch.optim.SGD(params, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, fused=True) if w['train'] else None

        def step():
            if w['train']:
                opt.zero_grad(set_to_none=True)
                with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                    o = fwd(x, p, True)
                loss = F.cross_entropy(o.float(), y)
                loss.backward()
                opt.step()
                return loss
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                return fwd(x, p, False)

        def timed(fn):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        try:
            ms = timed(step)
            out[tag + '_eager'] = dict(ms_per_step=ms, samples_per_s=batch / ms * 1e3)
        except Exception as e:  # noqa: BLE001
            out[tag + '_eager'] = dict(error=str(e)[:200])
            continue
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(3):
                    step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            if opt is not None:
                opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g):
                step()
            ms = timed(g.replay)
            out[tag + '_cuda_graph'] = dict(ms_per_step=ms, samples_per_s=batch / ms * 1e3)
            del g
        except Exception as e:  # noqa: BLE001
            out[tag + '_cuda_graph'] = dict(error=str(e)[:200])
        del p, params, opt
        torch.cuda.synchronize()
    out['note'] = ('oracle port of the reference modules (identical ATen call sequence) in stock PyTorch %s on this GPU, '
                   'default TF32 flags, fused SGD; batch %d' % (torch.__version__, batch))
    return out


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
# roofline of the fused CTRGC kernels at HBM-resident sizes
# ------------------------------------------------------------------------------------------------------------
def _traffic(tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    kernel at this shape (profiles/traffic.json, written when the capture is taken); None when there is none."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get(tag)
    except Exception:
        return None


def ctrgc_roofline(dtype, peak_gbs, peak_src, N, Cout, T, V, K, R, backward=False, tag=None):
    from tam_gcn_b200 import ops
    dev = torch.device('cuda')
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
    if not backward:
        # algorithmic bytes per launch: read K x3 planes once, write y once, read fp32 x1/x2, weights
        alg = s * N * T * V * Cout * (K + 1) + K * 8 * N * R * V + K * 4 * (Cout * R + Cout + V * V)
        run = lambda: ops.ctrgc_fwd(x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y, stats=(st[0], st[1]))
    else:
        # read g once, read K x3 planes, write K dx3 planes, read + write fp32 x1/x2 and their gradients
        alg = s * N * T * V * Cout * (1 + 2 * K) + 2 * K * 8 * N * R * V + K * 4 * (Cout * R + Cout + V * V)
        dx3 = torch.empty_like(x3)
        dx12 = torch.zeros_like(x12)
        acc = [torch.zeros_like(W4), torch.zeros_like(b4), torch.zeros_like(PA), torch.zeros(1, device=dev)]
        run = lambda: ops.ctrgc_bwd(ops.Opnd(y), x3, x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, dx3, dx12[:, :K * R],
                                    dx12[:, K * R:], acc[0], acc[1], acc[2], acc[3])
        y.normal_()
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
    return dict(bound='hbm', kernel=('fused CTRGC backward' if backward else 'fused CTRGC forward') + ' (tamgcn_ctrgc_%s)' % ('bwd' if backward else 'fwd'),
                achieved=ach, peak=peak_gbs, unit='GB/s', frac=ach / peak_gbs, traffic=_traffic(tag) if tag else None,
                peak_source=peak_src, algorithmic_bytes=alg, ms_per_launch=ms,
                shape=dict(N=N, Cout=Cout, T=T, V=V, K=K, R=R, dtype=str(dtype).replace('torch.', '')),
                note='inputs %.2f GB > 126 MB L2; back-to-back launches' % (x3.numel() * s / 1e9))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), float(d.get('bf16_tflops_sustained', 1359.5)), 'MEASURED_PEAKS.json (measured)'
    except Exception:
        return 6650.0, 1400.0, 'B200_PROFILING.md fallback'


# ------------------------------------------------------------------------------------------------------------
# step-level roofline: device time per kernel family (CUPTI) against the algorithmic bytes / FLOPs of the step
# ------------------------------------------------------------------------------------------------------------
FAMILIES = [
    ('conv_wgrad', r'conv_wg2_kernel|tconv_wgrad_mma_kernel|conv_wgrad_kernel|conv_wgrad_tc_kernel|smallL_wgrad|tconv9_wgrad'),
    ('conv_dgrad', r'conv_tc2_kernel<\(int\)1|conv_tc2_kernel<1|conv_dgrad|tconv_mma_kernel.*dgrad|smallL_dgrad|tconv9_kernel<\(int\)1|tconv9_kernel<1'),
    ('conv_fwd', r'conv_tc2_kernel|tconv_mma_kernel|conv_fwd|smallL_fwd|pack_w2|tconv9_kernel'),
    ('ctrgc_bwd', r'ctrgc_bwd'),
    ('ctrgc_fwd', r'ctrgc_fwd'),
    ('epilogues+maxpool', r'epilogue|maxpool|gcn_mid|mean_t'),
    ('bn_coefficients', r'bn_finalize|bn_bwd_coef'),
    ('graph_agg', r'graph_agg'),
    ('head+sgd', r'data_bn|pool_fc|softmax_ce|sgd_step|scale_by_scalar'),
    ('nccl', r'nccl'),
    ('memset/copy', r'[Mm]emset|[Mm]emcpy'),
]


def step_roofline(replay, acct, n_steps, peak_gbs, peak_tf, ms_per_step):
    """replay(): one CUDA-graph step.  acct: {family: [bytes, flops]} accumulated by ops.account during the capture."""
    from torch.profiler import ProfilerActivity, profile
    for _ in range(2):
        replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_steps):
            replay()
        torch.cuda.synchronize()
    fam_us, fam_n, other, kern = {}, {}, {}, []
    total = 0.0
    for ev in prof.key_averages():
        us = float(getattr(ev, 'device_time_total', 0.0) or 0.0)
        if us <= 0:
            us = float(getattr(ev, 'cuda_time_total', 0.0) or 0.0)
        if us <= 0:
            continue
        total += us
        kern.append((us, ev.count, ev.key))
        for fam, pat in FAMILIES:
            if re.search(pat, ev.key):
                fam_us[fam] = fam_us.get(fam, 0.0) + us
                fam_n[fam] = fam_n.get(fam, 0) + ev.count
                break
        else:
            other[ev.key[:60]] = other.get(ev.key[:60], 0.0) + us
    rows = []
    for fam, us in sorted(fam_us.items(), key=lambda kv: -kv[1]):
        per_step_us = us / n_steps
        b, fl = acct.get(fam, (0, 0))
        row = dict(family=fam, us_per_step=round(per_step_us, 1), launches_per_step=round(fam_n[fam] / n_steps, 1),
                   share_of_kernel_time=round(us / total, 3))
        if b:
            gbs = b / (per_step_us * 1e-6) / 1e9
            row.update(algorithmic_mb_per_step=round(b / 1e6, 1), gb_per_s=round(gbs, 1), frac_hbm_peak=round(gbs / peak_gbs, 4))
        if fl:
            tf = fl / (per_step_us * 1e-6) / 1e12
            row.update(gflop_per_step=round(fl / 1e9, 2), tflop_per_s=round(tf, 2), frac_bf16_sustained=round(tf / peak_tf, 4))
        rows.append(row)
    if other:
        rows.append(dict(family='other', us_per_step=round(sum(other.values()) / n_steps, 1), kernels=sorted(other, key=other.get)[-5:]))
    top = [dict(kernel=re.sub(r'^void |tamgcn::', '', k)[:90], us_per_step=round(u / n_steps, 1), launches_per_step=round(c / n_steps, 1),
                us_per_launch=round(u / max(c, 1), 1)) for u, c, k in sorted(kern, reverse=True)[:16]]
    return dict(rows=rows, top_kernels=top, kernel_us_per_step=round(total / n_steps, 1), ms_per_step_timed=ms_per_step,
                note='device time summed over kernels from torch.profiler (CUPTI) over %d graph replays with warm L2; side-stream '
                     'kernels overlap, so the sum can exceed the step time. bytes/FLOPs: algorithmic, at the training shape' % n_steps)


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
        reference_arm(args, rank, world)
        return
    import torch.distributed as dist
    import tam_gcn_b200
    from tam_gcn_b200 import _C, engine, ops
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the B200-native path has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    tam_gcn_b200.set_act_dtype(dtype)
    w = WORKLOADS[args.workload]
    B = args.batch or w['batch']

    model = build_model(w, dev)
    xh, yh = synthetic_batch(w, B, 1000 + rank)
    xh, yh = xh.pin_memory(), yh.pin_memory()
    x, y = xh.to(dev), yh.to(dev)
    acct = {}
    if w['train']:
        model.train()
        runner = engine.Trainer(model, lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4, use_graph=not args.no_graph,
                                side_stream=not args.no_side_stream, overlap_allreduce=not args.no_overlap)
        runner.step(x, y)                            # captures the graph
        with ops.account(acct):                      # one more (eager) step so the accounting sees ONE step's calls
            runner._step_body(x, y)
        step = lambda: runner.step(x, y)
        step_host = lambda: runner.step_from_host(xh, yh)
        d2h = 4
    else:
        model.eval()
        for p in model.parameters():
            p.requires_grad_(False)
        if w['family'] != 'fusion':                  # calibrate running statistics (fresh ones overflow in eval mode)
            bns = [b for b in model.modules() if isinstance(b, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d))]
            for b in bns:
                b.momentum = 1.0
            model.train()
            with torch.no_grad():
                model(x[:min(B, 32)])
            for b in bns:
                b.momentum = 0.1
            model.eval()
        runner = engine.Predictor(model, use_graph=not args.no_graph, eval_mode=w['family'] != 'fusion')
        out0 = runner(x)
        with ops.account(acct), torch.no_grad():
            model(x)
        step = lambda: runner(x)
        step_host = lambda: runner.from_host(xh)
        d2h = out0.numel() * out0.element_size()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) --------------------------------------------------------------------
    W = max(args.warmup, 3)
    for _ in range(W):
        last = step()
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
        last = step()
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
    launches = runner.captured_launches * args.steps if runner.graph is not None else _C.launch_count() - l0
    final = float(last) if w['train'] else float(last.float().abs().mean())

    # ---- end to end through the public API, host buffers ----------------------------------------------------
    for _ in range(3):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / float(e2e_s)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_tf, peak_src = measured_peaks()
        roof = roof_ntu = roof_bwd = roof_step = None
        if not args.no_roofline:
            # the headline (NW-UCLA) l2-l4 block shape at an HBM-resident batch: N'=2048, Cout=64, T=52, V=20, K=3, R=8
            roof = ctrgc_roofline(dtype, peak, peak_src, 2048, 64, 52, 20, 3, 8, tag='ctrgc_fwd_ucla_2048')
            roof_bwd = ctrgc_roofline(dtype, peak, peak_src, 2048, 64, 52, 20, 3, 8, backward=True, tag='ctrgc_bwd_ucla_2048')
            # the cfg4 size, batch 1024 -> N'=2048, l2-l4: C=64, T=64, V=25
            roof_ntu = ctrgc_roofline(dtype, peak, peak_src, 2048, 64, 64, 25, 3, 8, tag='ctrgc_fwd_ntu_2048')
        # (single GPU only: with N > 1 the step graph holds the all-reduce, and replaying it on rank 0 alone would wait
        # for the other ranks forever)
        if not args.no_step_roofline and runner.graph is not None and world == 1:
            try:
                roof_step = step_roofline(runner.graph.replay, acct, 5, peak, peak_tf, ms_per_step)
            except Exception as e:  # noqa: BLE001
                roof_step = dict(error=str(e)[:300])
        gpu_base = None
        if not args.no_gpu_baseline and world == 1 and w['family'] != 'fusion':
            try:
                gpu_base = gpu_eager_baseline(w, B, dev)
            except Exception as e:  # noqa: BLE001
                gpu_base = dict(error=str(e)[:300])
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            wc = dict(w, family='ctrgcn') if w['family'] == 'fusion' else w
            n_cpu = 2 if args.workload == 'ucla_train' else 1
            ts = cpu_steps(wc, B if args.workload == 'ucla_train' else min(B, 16), n_cpu, 1, cores)
            bc = B if args.workload == 'ucla_train' else min(B, 16)
            cpu = dict(value=bc / (sum(ts) / len(ts)), unit='samples/s', cores=cores, kind='port',
                       sample='%d %s at batch %d after 1 warm-up, fp32, %d threads' %
                              (n_cpu, 'fwd+bwd+SGD steps' if w['train'] else 'eval forwards', bc, cores))
        line = dict(metric=w['metric'], value=value, unit='samples/s', n_gpus=world,
                    steps=args.steps, warmup=W, ms_per_step=ms_per_step, higher_is_better=True,
                    scaling='weak', vs_baseline=None, dtype=args.dtype, data='synthetic',
                    config=config_of(w, args.workload, B, world, cuda_graph=runner.graph is not None,
                                     l2='flushed between steps (256 MiB memset outside the timed CUDA events)',
                                     weights='reference init, dead paths perturbed (alpha=0.7, offset conv N(0,0.05))',
                                     side_stream=not args.no_side_stream, allreduce_overlap=(world > 1 and not args.no_overlap)),
                    e2e=dict(value=e2e_val, unit='samples/s', h2d_bytes_per_step=xh.numel() * 4 + (yh.numel() * 8 if w['train'] else 0),
                             d2h_bytes_per_step=d2h),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, roofline_bwd=roof_bwd, roofline_ntu=roof_ntu,
                    roofline_step=roof_step, gpu_eager_baseline=gpu_base, cpu_baseline=cpu,
                    loss=final, wall_s_timed_region=t_wall)
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear-down: the captured graph holds NCCL work; release it first, and never let a wedged communicator
        # tear-down keep the (already printed) run alive
        sys.stdout.flush()
        dist.barrier()
        runner.graph = None
        del runner
        import gc
        gc.collect()
        torch.cuda.synchronize()
        t = threading.Timer(15.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
