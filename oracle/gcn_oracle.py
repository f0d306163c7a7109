"""CPU oracle for the CTR-GCN / ST-GCN hot path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file.  The product path (`tam_gcn_b200/`) never routes through it and
has no CPU fallback.

What it is: a *functional* (stateless, parameter-dict driven) restatement of the reference's
forward math, in plain torch CPU ops, runnable in fp32 or fp64.  Gradients come from torch
autograd over these functions.  Each function cites the reference lines it restates
(paths relative to /root/reference).

Parity pinning: the reference ships no golden vectors or tests for this path (SURVEY.md §4,
§8c: "parity unpinned" upstream).  The oracle is therefore pinned against the reference
*itself*: `oracle/make_golden.py` imports the unmodified reference modules in the build
container, loads the same perturbed state, and (a) asserts oracle == reference to fp64
round-off and (b) writes small input/output fixtures to `tests/golden/`, which
`tests/test_oracle_golden.py` replays on every run (also on the GPU box, where the reference
is absent).

Parameter dictionaries use the reference's `state_dict` names, e.g.
`l5.gcn1.convs.0.conv3.weight`.  `p` maps name -> tensor, `pre` is the dotted prefix.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def conv(x, p, pre, stride=1, pad=0, dil=1):
    """nn.Conv2d with a (k,1) kernel acting on (N,C,T,V)."""
    return F.conv2d(x, p[pre + '.weight'], p.get(pre + '.bias'), stride=(stride, 1),
                    padding=(pad, 0), dilation=(dil, 1))


def batch_norm(x, p, pre, train):
    """nn.BatchNorm{1,2}d defaults: eps 1e-5, momentum 0.1, biased var for normalisation,
    unbiased var into running_var, num_batches_tracked += 1."""
    rm, rv = p.get(pre + '.running_mean'), p.get(pre + '.running_var')
    if train and (pre + '.num_batches_tracked') in p:
        p[pre + '.num_batches_tracked'] += 1
    use_batch = train or rm is None
    return F.batch_norm(x, rm, rv, p[pre + '.weight'], p[pre + '.bias'], use_batch, BN_MOMENTUM, BN_EPS)


# --------------------------------------------------------------------------------------
# CTR-GCN  (models/ctrgcn.py)
# --------------------------------------------------------------------------------------
def ctrgc(x, p, pre, A=None, alpha=1):
    """CTRGC.forward, models/ctrgcn.py:172-177."""
    x1 = conv(x, p, pre + '.conv1').mean(-2)                      # (N,R,V)
    x2 = conv(x, p, pre + '.conv2').mean(-2)
    x3 = conv(x, p, pre + '.conv3')                               # (N,C,T,V)
    d = torch.tanh(x1.unsqueeze(-1) - x2.unsqueeze(-2))           # (N,R,V,V)  d[n,r,u,v]
    q = conv(d, p, pre + '.conv4') * alpha                        # (N,C,V,V)
    if A is not None:
        q = q + A.unsqueeze(0).unsqueeze(0)
    return torch.einsum('ncuv,nctv->nctu', q, x3)


def unit_gcn(x, p, pre, train=True):
    """unit_gcn.forward (adaptive=True), models/ctrgcn.py:246-263."""
    PA, alpha = p[pre + '.PA'], p[pre + '.alpha']
    y = None
    for i in range(PA.shape[0]):
        z = ctrgc(x, p, f'{pre}.convs.{i}', PA[i], alpha)
        y = z if y is None else z + y
    y = batch_norm(y, p, pre + '.bn', train)
    if (pre + '.down.0.weight') in p:                             # :210-214
        res = batch_norm(conv(x, p, pre + '.down.0'), p, pre + '.down.1', train)
    elif p.get(pre + '.__residual__', True):                      # :216
        res = x
    else:                                                         # :218
        res = 0
    diff = res - y
    off = torch.tanh(batch_norm(conv(diff, p, pre + '.offset_conv.0'), p, pre + '.offset_conv.1', train))
    return torch.relu(y + off + res)


def temporal_conv(x, p, pre, ksize, stride=1, dil=1, train=True):
    """TemporalConv.forward, models/ctrgcn.py:52-69."""
    pad = (ksize + (ksize - 1) * (dil - 1) - 1) // 2
    return batch_norm(conv(x, p, pre + '.conv', stride, pad, dil), p, pre + '.bn', train)


def ms_tcn(x, p, pre, kernel_size=3, stride=1, dilations=(1, 2, 3, 4), residual=True,
           residual_kernel_size=1, train=True):
    """MultiScale_TemporalConv.forward, models/ctrgcn.py:72-147."""
    nd = len(dilations)
    ks = list(kernel_size) if isinstance(kernel_size, (list, tuple)) else [kernel_size] * nd
    outs = []
    for b, (k, d) in enumerate(zip(ks, dilations)):                # :93-110
        h = torch.relu(batch_norm(conv(x, p, f'{pre}.branches.{b}.0'), p, f'{pre}.branches.{b}.1', train))
        outs.append(temporal_conv(h, p, f'{pre}.branches.{b}.3', k, stride, d, train))
    b = nd                                                         # :113-119 max-pool branch
    h = torch.relu(batch_norm(conv(x, p, f'{pre}.branches.{b}.0'), p, f'{pre}.branches.{b}.1', train))
    h = F.max_pool2d(h, kernel_size=(3, 1), stride=(stride, 1), padding=(1, 0))
    outs.append(batch_norm(h, p, f'{pre}.branches.{b}.4', train))
    b = nd + 1                                                     # :121-124 strided 1x1 branch
    outs.append(batch_norm(conv(x, p, f'{pre}.branches.{b}.0', stride), p, f'{pre}.branches.{b}.1', train))
    out = torch.cat(outs, dim=1)
    if not residual:                                               # :127-132
        return out
    if (pre + '.residual.conv.weight') in p:
        return out + temporal_conv(x, p, pre + '.residual', residual_kernel_size, stride, 1, train)
    return out + x


def unit_tcn(x, p, pre, kernel_size=9, stride=1, train=True):
    """unit_tcn.forward, models/ctrgcn.py:179-193 (its ReLU is never applied)."""
    pad = int((kernel_size - 1) / 2)
    return batch_norm(conv(x, p, pre + '.conv', stride, pad), p, pre + '.bn', train)


def tcn_gcn_unit(x, p, pre, stride=1, residual=True, kernel_size=5, dilations=(1, 2), train=True):
    """TCN_GCN_unit.forward, models/ctrgcn.py:266-284."""
    y = ms_tcn(unit_gcn(x, p, pre + '.gcn1', train), p, pre + '.tcn1', kernel_size, stride,
               dilations, residual=False, train=train)
    if not residual:
        r = 0
    elif (pre + '.residual.conv.weight') in p:
        r = unit_tcn(x, p, pre + '.residual', 1, stride, train)
    else:
        r = x
    return torch.relu(y + r)


# (in, out, stride, residual) of l1..l10, models/ctrgcn.py:305-314
CTRGCN_LAYERS = [(None, 64, 1, False), (64, 64, 1, True), (64, 64, 1, True), (64, 64, 1, True),
                 (64, 128, 2, True), (128, 128, 1, True), (128, 128, 1, True),
                 (128, 256, 2, True), (256, 256, 1, True), (256, 256, 1, True)]


def _data_bn(x, p, train, ctr_layout):
    """Input reshuffle + data_bn, models/ctrgcn.py:325-332 / models/stgcn.py:171-184."""
    N, C, T, V, M = x.shape
    x = x.permute(0, 4, 3, 1, 2).contiguous()
    x = x.view(N, M * V * C, T) if ctr_layout else x.view(N * M, V * C, T)
    x = batch_norm(x, p, 'data_bn', train)
    return x.view(N, M, V, C, T).permute(0, 1, 3, 4, 2).contiguous().view(N * M, C, T, V)


def _as5d(x, num_point):
    if x.dim() == 3:                                               # models/ctrgcn.py:325-327
        N, T, _ = x.shape
        x = x.view(N, T, num_point, -1).permute(0, 3, 1, 2).contiguous().unsqueeze(-1)
    return x


def ctrgcn_features(x, p, num_point, train=True):
    x = _as5d(x, num_point)
    N, C, T, V, M = x.shape
    h = _data_bn(x, p, train, ctr_layout=True)
    for i, (_, _, stride, residual) in enumerate(CTRGCN_LAYERS):
        h = tcn_gcn_unit(h, p, f'l{i + 1}', stride, residual, train=train)
    return h, (N, M)


def ctrgcn_forward(x, p, num_point, train=True):
    """Model.forward, models/ctrgcn.py:324-348 (drop_out=0)."""
    h, (N, M) = ctrgcn_features(x, p, num_point, train)
    h = h.view(N, M, h.shape[1], -1).mean(3).mean(1)
    return F.linear(h, p['fc.weight'], p['fc.bias'])


def ctrgcn_extract_feature(x, p, num_point, train=True):
    """Model.extract_feature, models/ctrgcn.py:350-374."""
    h, (N, M) = ctrgcn_features(x, p, num_point, train)
    _, C, T, V = h.shape
    return h.view(N, M, C, T, V).permute(0, 2, 3, 4, 1).contiguous()


# --------------------------------------------------------------------------------------
# ST-GCN  (models/stgcn.py)
# --------------------------------------------------------------------------------------
def conv_temporal_graphical(x, A, p, pre, K):
    """ConvTemporalGraphical.forward, models/stgcn.py:57-63 (t_kernel_size=1)."""
    assert A.shape[0] == K
    y = conv(x, p, pre + '.conv')
    n, kc, t, v = y.shape
    y = y.view(n, K, kc // K, t, v)
    return torch.einsum('nkctv,kvw->nctw', y, A).contiguous()


def st_gcn_block(x, A, p, pre, stride=1, residual=True, tk=9, train=True):
    """st_gcn.forward, models/stgcn.py:95-99 (dropout=0)."""
    if not residual:
        res = 0
    elif (pre + '.residual.0.weight') in p:
        res = batch_norm(conv(x, p, pre + '.residual.0', stride), p, pre + '.residual.1', train)
    else:
        res = x
    h = conv_temporal_graphical(x, A, p, pre + '.gcn', A.shape[0])
    h = torch.relu(batch_norm(h, p, pre + '.tcn.0', train))
    h = batch_norm(conv(h, p, pre + '.tcn.2', stride, (tk - 1) // 2), p, pre + '.tcn.3', train)
    return torch.relu(h + res)


STGCN_LAYERS = [(1, False), (1, True), (1, True), (1, True), (2, True), (1, True), (1, True),
                (2, True), (1, True), (1, True)]                   # models/stgcn.py:140-151


def stgcn_forward(x, p, num_point, train=True):
    """Model.forward, models/stgcn.py:170-198 (edge_importance_weighting=True, dropout=0)."""
    x = _as5d(x, num_point)
    N, C, T, V, M = x.shape
    h = _data_bn(x, p, train, ctr_layout=False)
    for i, (stride, residual) in enumerate(STGCN_LAYERS):
        h = st_gcn_block(h, p['A'] * p[f'edge_importance.{i}'], p, f'st_gcn_networks.{i}', stride,
                         residual, train=train)
    h = h.mean((2, 3), keepdim=True).view(N, M, -1, 1, 1).mean(1)
    return conv(h, p, 'fcn').view(N, -1)


# --------------------------------------------------------------------------------------
# cross-modal fusion head  (models/resnet_gcn_attention.py)
# --------------------------------------------------------------------------------------
def fusion_attention(x_gcn, p, num_point, train=True):
    """GCN branch of ResNet_GCN_Attention.forward, models/resnet_gcn_attention.py:82-89: CTR-GCN feature, mean over
    (T, V, M), attention_transform = Linear -> BatchNorm1d -> ReLU -> Linear -> Sigmoid (:59-65).  -> (N, 2048)."""
    pg = {k[4:]: v for k, v in p.items() if k.startswith('gcn.')}
    f = ctrgcn_extract_feature(x_gcn, pg, num_point, train).mean(dim=(2, 3, 4))
    h = F.linear(f, p['attention_transform.0.weight'], p['attention_transform.0.bias'])
    h = torch.relu(batch_norm(h, p, 'attention_transform.1', train))
    return torch.sigmoid(F.linear(h, p['attention_transform.3.weight'], p['attention_transform.3.bias']))


def fusion_forward(x_gcn, f_rgb, p, num_point, train=True):
    """ResNet_GCN_Attention.forward from the backbone output f_rgb (N, 2048, 7, 7) on, :108-120: channel gate,
    global average pool, classifier."""
    att = fusion_attention(x_gcn, p, num_point, train)
    out = (f_rgb * att.unsqueeze(-1).unsqueeze(-1)).mean((2, 3))
    return F.linear(out, p['classifier.weight'], p['classifier.bias'])


def add_fusion_head(p, g, num_class=10, cg=256, cr=2048):
    def lin(pre, co, ci):
        p[pre + '.weight'] = torch.randn(co, ci, generator=g, dtype=torch.float64) * ci ** -0.5
        p[pre + '.bias'] = 0.1 * torch.randn(co, generator=g, dtype=torch.float64)
    lin('attention_transform.0', cr // 2, cg)
    _add_bn(p, 'attention_transform.1', cr // 2, g)
    lin('attention_transform.3', cr, cr // 2)
    lin('classifier', num_class, cr)


# --------------------------------------------------------------------------------------
# synthetic state (reference-named parameter dicts) and inputs  
# --------------------------------------------------------------------------------------
def _kaiming_fan_out(shape, g):
    fan_out = shape[0] * int(np.prod(shape[2:]))
    return torch.randn(shape, generator=g, dtype=torch.float64) * math.sqrt(2.0 / fan_out)


def _add_conv(p, pre, cout, cin, k, g, bias_scale=0.1):
    p[pre + '.weight'] = _kaiming_fan_out((cout, cin, k, 1), g)
    p[pre + '.bias'] = bias_scale * torch.randn(cout, generator=g, dtype=torch.float64)


def _add_bn(p, pre, c, g, dims=1):
    p[pre + '.weight'] = 1 + 0.1 * torch.randn(c, generator=g, dtype=torch.float64)
    p[pre + '.bias'] = 0.1 * torch.randn(c, generator=g, dtype=torch.float64)
    p[pre + '.running_mean'] = 0.05 * torch.randn(c, generator=g, dtype=torch.float64)
    p[pre + '.running_var'] = 1 + 0.2 * torch.rand(c, generator=g, dtype=torch.float64)
    p[pre + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)


def rel_channels(cin):
    return 8 if cin in (3, 9) else cin // 8                        # models/ctrgcn.py:155-160


def add_ctrgc(p, pre, cin, cout, g):
    r = rel_channels(cin)
    _add_conv(p, pre + '.conv1', r, cin, 1, g)
    _add_conv(p, pre + '.conv2', r, cin, 1, g)
    _add_conv(p, pre + '.conv3', cout, cin, 1, g)
    _add_conv(p, pre + '.conv4', cout, r, 1, g)


def add_unit_gcn(p, pre, cin, cout, A, g):
    """Perturbed (non-degenerate) unit_gcn state: alpha=0.7, offset conv N(0,0.05), bn ~ 1+0.1N
    (the reference init makes these paths numerically dead)."""
    for i in range(A.shape[0]):
        add_ctrgc(p, f'{pre}.convs.{i}', cin, cout, g)
    if cin != cout:
        _add_conv(p, pre + '.down.0', cout, cin, 1, g)
        _add_bn(p, pre + '.down.1', cout, g)
    p[pre + '.offset_conv.0.weight'] = 0.05 * torch.randn(cout, cout, 1, 1, generator=g, dtype=torch.float64)
    p[pre + '.offset_conv.0.bias'] = 0.1 * torch.randn(cout, generator=g, dtype=torch.float64)
    _add_bn(p, pre + '.offset_conv.1', cout, g)
    p[pre + '.PA'] = torch.as_tensor(np.asarray(A), dtype=torch.float64).clone() \
        + 0.02 * torch.randn(A.shape, generator=g, dtype=torch.float64)
    p[pre + '.alpha'] = torch.full((1,), 0.7, dtype=torch.float64)
    _add_bn(p, pre + '.bn', cout, g)


def add_ms_tcn(p, pre, cin, cout, kernel_size, dilations, g, residual_conv=False, residual_kernel_size=1):
    nd = len(dilations)
    cb = cout // (nd + 2)
    ks = list(kernel_size) if isinstance(kernel_size, (list, tuple)) else [kernel_size] * nd
    for b in range(nd):
        _add_conv(p, f'{pre}.branches.{b}.0', cb, cin, 1, g)
        _add_bn(p, f'{pre}.branches.{b}.1', cb, g)
        _add_conv(p, f'{pre}.branches.{b}.3.conv', cb, cb, ks[b], g)
        _add_bn(p, f'{pre}.branches.{b}.3.bn', cb, g)
    _add_conv(p, f'{pre}.branches.{nd}.0', cb, cin, 1, g)
    _add_bn(p, f'{pre}.branches.{nd}.1', cb, g)
    _add_bn(p, f'{pre}.branches.{nd}.4', cb, g)
    _add_conv(p, f'{pre}.branches.{nd + 1}.0', cb, cin, 1, g)
    _add_bn(p, f'{pre}.branches.{nd + 1}.1', cb, g)
    if residual_conv:
        _add_conv(p, pre + '.residual.conv', cout, cin, residual_kernel_size, g)
        _add_bn(p, pre + '.residual.bn', cout, g)


def add_tcn_gcn_unit(p, pre, cin, cout, A, stride, residual, g):
    add_unit_gcn(p, pre + '.gcn1', cin, cout, A, g)
    add_ms_tcn(p, pre + '.tcn1', cout, cout, 5, (1, 2), g)
    if residual and not (cin == cout and stride == 1):
        _add_conv(p, pre + '.residual.conv', cout, cin, 1, g)
        _add_bn(p, pre + '.residual.bn', cout, g)


def make_ctrgcn_state(A, num_class, num_person, in_channels=3, seed=0, dtype=torch.float32):
    """Full CTR-GCN state dict (reference names, App. B), perturbed per SURVEY.md §8d."""
    g = torch.Generator().manual_seed(seed)
    V = A.shape[1]
    p = {}
    _add_bn(p, 'data_bn', num_person * in_channels * V, g)
    cin = in_channels
    for i, (_, cout, stride, residual) in enumerate(CTRGCN_LAYERS):
        add_tcn_gcn_unit(p, f'l{i + 1}', cin, cout, A, stride, residual, g)
        cin = cout
    p['fc.weight'] = torch.randn(num_class, 256, generator=g, dtype=torch.float64) * math.sqrt(2.0 / num_class)
    p['fc.bias'] = 0.1 * torch.randn(num_class, generator=g, dtype=torch.float64)
    return cast_state(p, dtype)


def make_stgcn_state(A, num_class, in_channels=3, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    K, V, _ = A.shape
    p = {'A': torch.as_tensor(np.asarray(A), dtype=torch.float64).clone()}
    _add_bn(p, 'data_bn', in_channels * V, g)
    chans = [(in_channels, 64), (64, 64), (64, 64), (64, 64), (64, 128), (128, 128), (128, 128),
             (128, 256), (256, 256), (256, 256)]
    for i, ((cin, cout), (stride, residual)) in enumerate(zip(chans, STGCN_LAYERS)):
        pre = f'st_gcn_networks.{i}'
        _add_conv(p, pre + '.gcn.conv', cout * K, cin, 1, g)
        _add_bn(p, pre + '.tcn.0', cout, g)
        _add_conv(p, pre + '.tcn.2', cout, cout, 9, g)
        _add_bn(p, pre + '.tcn.3', cout, g)
        if residual and not (cin == cout and stride == 1):
            _add_conv(p, pre + '.residual.0', cout, cin, 1, g)
            _add_bn(p, pre + '.residual.1', cout, g)
    for i in range(10):
        p[f'edge_importance.{i}'] = 1 + 0.1 * torch.randn(K, V, V, generator=g, dtype=torch.float64)
    _add_conv(p, 'fcn', num_class, 256, 1, g)
    return cast_state(p, dtype)


def cast_state(p, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v.clone()) for k, v in p.items()}


def clone_state(p, dtype=None, requires_grad=False):
    out = {}
    for k, v in p.items():
        if not torch.is_tensor(v):
            out[k] = v
            continue
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        is_param = t.is_floating_point() and not k.endswith(('running_mean', 'running_var')) and k != 'A'
        if requires_grad and is_param:
            t.requires_grad_(True)
        out[k] = t
    return out


def synthetic_skeletons(N, T, V, M, C=3, seed=0, dtype=torch.float32):
    """randn*0.5 clipped to [-1,1] — mimics the feeder's min-max output
    (feeder/feeder_nucla_gcn.py:103-105), SURVEY.md §8d."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(N, C, T, V, M, generator=g, dtype=torch.float64) * 0.5).clamp_(-1, 1).to(dtype)


def rel_err(a, b):
    """relative L2 error ||a-b|| / ||b||."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
