"""Shared test helpers: build B200-native modules for the golden cases and compare with the fixtures."""
import os

import torch

from oracle import gcn_oracle as O
from oracle.cases import CASES, build_case, oracle_forward  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN, name + '.pt'), map_location='cpu', weights_only=False)


def strip(p, pre):
    n = len(pre) + 1
    return {k[n:]: v for k, v in p.items() if k.startswith(pre + '.')}


def our_module(case):
    """Construct the tam_gcn_b200 module for a golden case (same ctor arguments the reference was given)."""
    import tam_gcn_b200.ctrgcn as C
    import tam_gcn_b200.stgcn as S
    kind, a, A = case['kind'], case['args'], case['A']
    if kind == 'ctrgc':
        return C.CTRGC(a['cin'], a['cout'])
    if kind == 'unit_gcn':
        return C.unit_gcn(a['cin'], a['cout'], A, residual=a.get('residual', True))
    if kind == 'ms_tcn':
        return C.MultiScale_TemporalConv(a['cin'], a['cout'], kernel_size=a['kernel_size'], stride=a['stride'],
                                         dilations=list(a['dilations']), residual=a['residual'],
                                         residual_kernel_size=a.get('residual_kernel_size', 1))
    if kind == 'unit_tcn':
        return C.unit_tcn(a['cin'], a['cout'], kernel_size=a['kernel_size'], stride=a['stride'])
    if kind == 'tcn_gcn_unit':
        return C.TCN_GCN_unit(a['cin'], a['cout'], A, stride=a['stride'], residual=a['residual'])
    if kind == 'ctrgcn_model':
        return C.Model(num_class=a['num_class'], num_point=a['num_point'], num_person=a['num_person'],
                       graph=a['graph'], graph_args=dict(labeling_mode='spatial'))
    if kind == 'ctg':
        return S.ConvTemporalGraphical(a['cin'], a['cout'], a['K'])
    if kind == 'st_gcn':
        return S.st_gcn(a['cin'], a['cout'], (9, a['K']), a['stride'], residual=a['residual'])
    if kind == 'stgcn_model':
        return S.Model(in_channels=3, num_class=a['num_class'], num_point=a['num_point'], num_person=1,
                       graph=a['graph'], graph_args=dict(labeling_mode='spatial'))
    if kind == 'fusion_model':
        import tam_gcn_b200.fusion as Fu
        return Fu.ResNet_GCN_Attention(num_class=a['num_class'], num_point=a['num_point'], num_person=a['num_person'],
                                       graph=a['graph'], graph_args=dict(labeling_mode='spatial'), in_channels_rgb=3,
                                       resnet=identity_backbone())
    raise KeyError(kind)


def identity_backbone():
    """Stand-in for the ResNet-50 of the fusion model (outside the hot path): every stage is the identity, so the
    `x_rgb` argument IS the (N, 2048, 7, 7) backbone output the head consumes."""
    r = torch.nn.Module()
    for nm in ('conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4'):
        setattr(r, nm, torch.nn.Identity())
    return r


def load_state(m, case, state):
    sd = state if case['kind'].endswith('_model') else strip(state, 'm')
    sd = {k: v.clone() for k, v in sd.items() if not k.startswith('__')}
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m


def run_module(case, m, x, extra):
    kind = case['kind']
    if kind == 'ctrgc':
        return m(x, extra['A'], extra['alpha'])
    if kind in ('ctg', 'st_gcn'):
        return m(x, extra['A'])[0]
    if kind == 'fusion_model':
        return m(x, extra['f_rgb'])
    return m(x)


def run_case(name, device='cpu', act_dtype=torch.float32):
    """Forward + backward of our module on a golden case.  Returns dict(y, dx, grads, buffers)."""
    import tam_gcn_b200
    case = CASES[name]
    built = build_case(case)
    m = load_state(our_module(case), case, built['state']).to(device)
    m.train(case['train'])
    is_model = case['kind'].endswith('_model')
    x = built['x'].to(device)
    if not is_model:
        x = x.to(act_dtype)
    x.requires_grad_(True)
    extra = {k: v.to(device).requires_grad_(True) for k, v in built['extra'].items()}
    with tam_gcn_b200.act_dtype(act_dtype):
        y = run_module(case, m, x, extra)
    y.backward(built['cot'].to(device).to(y.dtype))
    grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters() if p.grad is not None}
    for k, v in extra.items():
        if v.grad is not None:
            grads['__' + k] = v.grad.detach().float().cpu()
    bufs = {k: b.detach().cpu() for k, b in m.named_buffers()}
    return dict(y=y.detach().float().cpu(), dx=x.grad.detach().float().cpu(), grads=grads, buffers=bufs)


def compare(name, res, fx, tol_y, tol_dx, tol_g, tol_buf=None, report=None, tol_gall=None, tol_gk=None):
    """Relative-L2 comparison against a golden fixture.  Returns the list of failures (strings)."""
    fails = []
    e_y, e_dx = O.rel_err(res['y'], fx['y']), O.rel_err(res['dx'], fx['dx'])
    if not e_y <= tol_y:
        fails.append('%s: y rel err %.3e > %.1e' % (name, e_y, tol_y))
    if not e_dx <= tol_dx:
        fails.append('%s: dx rel err %.3e > %.1e' % (name, e_dx, tol_dx))
    worst_g, worst_k = 0.0, None
    gscale = max([float(v.norm()) for v in fx['grads'].values()] + [1e-30])
    for k, gref in fx['grads'].items():
        if k not in res['grads']:
            fails.append('%s: missing gradient %s' % (name, k))
            continue
        # biases that feed a train-mode BatchNorm have exactly-zero true gradient
        if float(gref.norm()) < 1e-6 * gscale:
            continue
        e = O.rel_err(res['grads'][k], gref)
        if e > worst_g:
            worst_g, worst_k = e, k
        tk = max(tol_g, (tol_gk or {}).get(k, 0.0))       # per-tensor bound (bf16: cancellation-heavy tensors)
        if not e <= tk:
            fails.append('%s: grad %s rel err %.3e > %.1e' % (name, k, e, tk))
    if tol_gall is not None:
        ks = [k for k in fx['grads'] if k in res['grads']]
        ga = torch.cat([res['grads'][k].reshape(-1) for k in ks])
        gb = torch.cat([fx['grads'][k].reshape(-1) for k in ks])
        e = O.rel_err(ga, gb)
        if not e <= tol_gall:
            fails.append('%s: whole-gradient rel err %.3e > %.1e' % (name, e, tol_gall))
    worst_b = 0.0
    for k, b in fx['buffers'].items():
        if k not in res['buffers']:
            if k == 'A' or k.endswith('.A'):
                continue
            fails.append('%s: missing buffer %s' % (name, k))
            continue
        if b.is_floating_point():
            e = O.rel_err(res['buffers'][k], b)
            worst_b = max(worst_b, e)
            if not e <= (tol_buf or tol_y):
                fails.append('%s: buffer %s rel err %.3e' % (name, k, e))
        elif int(res['buffers'][k]) != int(b):
            fails.append('%s: buffer %s = %d, expected %d' % (name, k, int(res['buffers'][k]), int(b)))
    if report is not None:
        report.append('%-28s y %.2e dx %.2e dW %.2e (%s) buf %.2e' % (name, e_y, e_dx, worst_g, worst_k, worst_b))
    return fails


# ---- reference-initialised whole models (fresh init is the regime the benchmark runs in) ---------------------------
UCLA = dict(num_class=10, num_point=20, num_person=1, graph='graph.ucla.Graph', graph_args=dict(labeling_mode='spatial'))
NTU = dict(num_class=60, num_point=25, num_person=2, graph='graph.ntu_rgb_d.Graph', graph_args=dict(labeling_mode='spatial'))


def wake_dead_paths(model, seed=0):
    """The reference init leaves alpha = 0, the offset conv = 0 and unit_gcn.bn.weight = 1e-6: give
    them trained-like magnitudes so every kernel sees non-trivial operands (same recipe as bench.py)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.endswith('gcn1.alpha'):
                p.fill_(0.7)
            elif k.endswith('offset_conv.0.weight'):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif k.endswith('gcn1.bn.weight'):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
    return model


def fresh_ctrgcn(seed=0, **cfg):
    import tam_gcn_b200.ctrgcn as C
    torch.manual_seed(seed)
    kw = dict(UCLA)
    kw.update(cfg)
    return wake_dead_paths(C.Model(**kw), seed)


def fresh_stgcn(seed=0, **cfg):
    import tam_gcn_b200.stgcn as S
    torch.manual_seed(seed)
    kw = dict(in_channels=3, num_class=60, num_point=25, num_person=1, graph='graph.ntu_rgb_d.Graph',
              graph_args=dict(labeling_mode='spatial'))
    kw.update(cfg)
    m = S.Model(**kw)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():                                   # edge importance away from exactly 1
        for p in m.edge_importance:
            p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
    return m


def state_of(model):
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def grad_vector(named_grads, keys):
    return torch.cat([named_grads[k].detach().double().reshape(-1).cpu() for k in keys])
