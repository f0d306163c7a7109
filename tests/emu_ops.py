"""Pure-torch emulation of every function in tam_gcn_b200/ops.py  —  TEST INFRASTRUCTURE ONLY.

Same signatures and in-place output conventions as the real wrappers, written with ATen ops (and
torch autograd for the backward primitives).  Used two ways:

  * on CPU (no GPU in the build container): tests monkeypatch `tam_gcn_b200.ops` with these functions
    to check the host-side orchestration in tam_gcn_b200/functional.py (packing of parameters,
    BatchNorm coefficient algebra, gradient routing) against the golden fixtures;
  * on the GPU box: each CUDA kernel is compared against its emulation on the same random inputs
    (tests/test_kernels_gpu.py).

Nothing under tam_gcn_b200/ imports this file.
"""
import torch
import torch.nn.functional as F

RES_NONE, RES_IDENTITY, RES_AFFINE = 0, 1, 2


class Opnd:
    __slots__ = ('p', 'q', 'a', 'b', 'c', 'relu')

    def __init__(self, p, q=None, a=None, b=None, c=None, relu=False):
        self.p, self.q, self.a, self.b, self.c, self.relu = p, q, a, b, c, relu


def _cv(t):
    return t.view(1, -1, 1, 1)


def val(o, dtype=torch.float32):
    if torch.is_tensor(o):
        return o.to(dtype)
    v = o.p.to(dtype)
    if o.a is not None:
        v = v * _cv(o.a).to(dtype)
    if o.q is not None and o.b is not None:
        v = v + o.q.to(dtype) * _cv(o.b).to(dtype)
    if o.c is not None:
        v = v + _cv(o.c).to(dtype)
    if o.relu:
        v = torch.relu(v)
    return v


def _store(dst, v):
    dst.copy_(v.to(dst.dtype))
    return dst.to(torch.float32)


def _acc(stat, v):
    stat += v.double().sum((0, 2, 3))


def conv_fwd(x, W, bias, y, k=1, stride=1, dil=1, pad=0, stats=None, stat_c0=0, wpack=None):
    xv = val(x)
    Cout, Cin = y.shape[1], xv.shape[1]
    out = F.conv2d(xv, W.reshape(Cout, Cin, k, 1), bias, stride=(stride, 1), padding=(pad, 0), dilation=(dil, 1))
    out = _store(y, out)
    if stats is not None:
        _acc(stats[0], out[:, stat_c0:])
        _acc(stats[1], out[:, stat_c0:] ** 2)


def conv_dgrad(dy, W, dx, k=1, stride=1, dil=1, pad=0, addend=None, bcast=None, bcast_scale=0.0, mask=None,
               stats=None, wpack=None):
    dyv = val(dy)
    N, Cin, T, V = dx.shape
    Cout = dyv.shape[1]
    g = torch.nn.grad.conv2d_input((N, Cin, T, V), W.reshape(Cout, Cin, k, 1), dyv, stride=(stride, 1),
                                   padding=(pad, 0), dilation=(dil, 1))
    if addend is not None:
        g = g + addend.float()
    if bcast is not None:
        g = g + bcast.reshape(N, Cin, 1, V) * bcast_scale
    if mask is not None:
        pv = mask.p.float()
        mv = pv * (_cv(mask.a) if mask.a is not None else 1.0) + (_cv(mask.c) if mask.c is not None else 0.0)
        g = g * (mv > 0)
        g = _store(dx, g)
        if stats is not None:
            _acc(stats[0], g)
            _acc(stats[1], g * pv)
    else:
        _store(dx, g)


def conv_pack_weights(W, Cout, Cin, k):
    return None, None


def conv_wgrad(dy, x, dW, dbias, k=1, stride=1, dil=1, pad=0):
    dyv, xv = val(dy), val(x)
    Cout, Cin = dyv.shape[1], xv.shape[1]
    g = torch.nn.grad.conv2d_weight(xv, (Cout, Cin, k, 1), dyv, stride=(stride, 1), padding=(pad, 0),
                                    dilation=(dil, 1))
    dW += g.reshape(dW.shape)
    if dbias is not None:
        dbias += dyv.sum((0, 2, 3))


def mean_t(x, m):
    m.copy_(x.float().mean(2, keepdim=True))


def _ctrgc(x3, x1, x2, W4, b4, PA, alpha):
    N, KC, T, V = x3.shape
    K = PA.shape[0]
    Cc = KC // K
    R = x1.shape[1] // K
    D = torch.tanh(x1.reshape(N, K, R, V, 1) - x2.reshape(N, K, R, 1, V))
    Q = alpha.reshape(()) * (torch.einsum('kcr,nkruv->nkcuv', W4, D) + b4.view(1, K, Cc, 1, 1)) + PA.view(1, K, 1, V, V)
    return torch.einsum('nkcuv,nkctv->nctu', Q, x3.reshape(N, K, Cc, T, V))


def ctrgc_fwd(x3, x1, x2, W4, b4, PA, alpha, y, stats=None):
    out = _store(y, _ctrgc(x3.float(), x1, x2, W4, b4, PA, alpha))
    if stats is not None:
        _acc(stats[0], out)
        _acc(stats[1], out ** 2)


def ctrgc_bwd(g, x3, x1, x2, W4, b4, PA, alpha, dx3, dx1, dx2, dW4, db4, dPA, dalpha):
    with torch.enable_grad():
        leaves = [t.detach().float().clone().requires_grad_(True) for t in (x3, x1, x2, W4, b4, PA, alpha)]
        y = _ctrgc(*leaves)
        gr = torch.autograd.grad(y, leaves, val(g))
    dx3.copy_(gr[0].to(dx3.dtype))
    dx1 += gr[1]
    dx2 += gr[2]
    dW4 += gr[3]
    db4 += gr[4]
    dPA += gr[5]
    dalpha += gr[6].reshape(dalpha.shape)


def bn_finalize(descs, count, momentum, eps, train):
    for d in descs:
        if train:
            mean = d['sum'] / count
            var = (d['sumsq'] / count - mean * mean).clamp_min(0)
            if d.get('rmean') is not None:
                unb = var * count / (count - 1) if count > 1 else var
                d['rmean'].copy_(((1 - momentum) * d['rmean'].double() + momentum * mean).float())
                d['rvar'].copy_(((1 - momentum) * d['rvar'].double() + momentum * unb).float())
            if d.get('nbt') is not None:
                d['nbt'] += 1
        else:
            mean, var = d['rmean'].double(), d['rvar'].double()
        invstd = (1.0 / torch.sqrt(var + eps)).float()
        gamma = d['gamma'] if d.get('gamma') is not None else torch.ones_like(invstd)
        beta = d['beta'] if d.get('beta') is not None else torch.zeros_like(invstd)
        scale = gamma * invstd
        d['scale'].copy_(scale)
        d['shift'].copy_(beta - mean.float() * scale)
        if d.get('mean') is not None:
            d['mean'].copy_(mean.float())
        if d.get('invstd') is not None:
            d['invstd'].copy_(invstd)


def bn_bwd_coef(descs, count, train):
    for d in descs:
        mean, invstd = d['mean'].double(), d['invstd'].double()
        s1, s2 = d['s1'], d['s2']
        sx = invstd * (s2 - mean * s1)
        gamma = d['gamma'].double() if d.get('gamma') is not None else torch.ones_like(mean)
        A = gamma * invstd
        if train:
            B = -A * invstd * sx / count
            Cc = -A * s1 / count - B * mean
        else:
            B = torch.zeros_like(A)
            Cc = torch.zeros_like(A)
        d['A'].copy_(A.float())
        d['B'].copy_(B.float())
        d['Cc'].copy_(Cc.float())
        if d.get('dgamma') is not None:
            d['dgamma'].copy_(sx.float())
        if d.get('dbeta') is not None:
            d['dbeta'].copy_(s1.float())


def _resval(res_mode, r, sr, hr):
    if res_mode == RES_NONE:
        return 0.0
    if res_mode == RES_IDENTITY:
        return r.float()
    return r.float() * _cv(sr) + _cv(hr)


def gcn_epilogue_fwd(y0, sg, hg, z, so, ho, res_mode, r, sr, hr, out):
    v = y0.float() * _cv(sg) + _cv(hg) + torch.tanh(z.float() * _cv(so) + _cv(ho)) + _resval(res_mode, r, sr, hr)
    _store(out, torch.relu(v))


def gcn_epilogue_bwd(g, out, z, so, ho, G, DZ, s1o, s2o):
    gv = g.float() * (out.float() > 0)
    o = torch.tanh(z.float() * _cv(so) + _cv(ho))
    _store(G, gv)
    dz = _store(DZ, gv * (1 - o * o))
    _acc(s1o, dz)
    _acc(s2o, dz * z.float())


def coef_diff(sb, ha, hb, nb, c):
    nb.copy_(-sb)
    c.copy_((ha if ha is not None else 0) - hb)


def gcn_mid_bwd(G, DD, dr, y0, r, s1g, s2g, s1d, s2d, extra=None):
    gv, dd = G.float(), DD.float()
    dy = (gv - dd).to(G.dtype).float()
    drv = (gv + dd + (extra.float() if extra is not None else 0)).to(G.dtype).float()
    G.copy_(dy.to(G.dtype))
    _acc(s1g, dy)
    _acc(s2g, dy * y0.float())
    if dr is not None:
        dr.copy_(drv.to(dr.dtype))
    if r is not None:
        _acc(s1d, drv)
        _acc(s2d, drv * r.float())


def tcn_epilogue_fwd(u, su, hu, res_mode, r, sr, hr, relu, out):
    v = u.float() * _cv(su) + _cv(hu) + _resval(res_mode, r, sr, hr)
    _store(out, torch.relu(v) if relu else v)


def tcn_epilogue_bwd(g, out, relu, u, r, G, s1, s2u, s2r):
    gv = g.float()
    if relu:
        gv = gv * (out.float() > 0)
    if G is not None:
        G.copy_(gv.to(G.dtype))
    _acc(s1, gv)
    _acc(s2u, gv * u.float())
    if r is not None:
        _acc(s2r, gv * r.float())


def maxpool_fwd(x, y, stride, stats=None):
    out = _store(y, F.max_pool2d(val(x), (3, 1), (stride, 1), (1, 0)))
    if stats is not None:
        _acc(stats[0], out)
        _acc(stats[1], out ** 2)


def maxpool_bwd(dy, x, dh, stride, stats=None):
    with torch.enable_grad():
        xv = val(x).detach().requires_grad_(True)
        y = F.max_pool2d(xv, (3, 1), (stride, 1), (1, 0))
        g, = torch.autograd.grad(y, xv, val(dy))
    pv = x.p.float()
    mv = pv * (_cv(x.a) if x.a is not None else 1.0) + (_cv(x.c) if x.c is not None else 0.0)
    g = _store(dh, g * (mv > 0))
    if stats is not None:
        _acc(stats[0], g)
        _acc(stats[1], g * pv)


def graph_agg_fwd(y, A, out, stats=None):
    N, KC, T, V = y.shape
    K = A.shape[0]
    o = _store(out, torch.einsum('nkctv,kvw->nctw', y.float().reshape(N, K, KC // K, T, V), A))
    if stats is not None:
        _acc(stats[0], o)
        _acc(stats[1], o ** 2)


def graph_agg_bwd(dout, y, A, dy, dA):
    gv = val(dout)
    N, KC, T, V = y.shape
    K = A.shape[0]
    dy.copy_(torch.einsum('nctw,kvw->nkctv', gv, A).reshape(N, KC, T, V).to(dy.dtype))
    if dA is not None:
        dA += torch.einsum('nkctv,nctw->kvw', y.float().reshape(N, K, KC // K, T, V), gv)


def _x5(x, num_point):
    """(N, C, T, V, M) fp32 view of the model input (5-D as is, 3-D (N, T, V*C) as models/ctrgcn.py:325-327)."""
    if x.dim() == 5:
        return x
    N, T, VC = x.shape
    return x.view(N, T, num_point, -1).permute(0, 3, 1, 2).unsqueeze(-1)


def _data_bn_rows(x5, fold_m):
    N, C, T, V, M = x5.shape
    if fold_m:
        return x5.permute(0, 4, 3, 1, 2).reshape(N * M, V * C, T)       # models/stgcn.py:175-176
    return x5.permute(0, 4, 3, 1, 2).reshape(N, M * V * C, T)           # models/ctrgcn.py:329


def _rows_to_nchw(y, N, C, T, V, M):
    return y.reshape(N, M, V, C, T).permute(0, 1, 3, 4, 2).reshape(N * M, C, T, V)


def data_bn_fwd(x, num_point, fold_m, bn, train, out, save_mean, save_invstd):
    x5 = _x5(x, num_point)
    N, C, T, V, M = x5.shape
    rows = _data_bn_rows(x5, fold_m).double()
    if train:
        mean = rows.mean((0, 2))
        var = rows.var((0, 2), unbiased=False)
        cnt = rows.shape[0] * rows.shape[2]
        m = bn.momentum
        bn.running_mean.copy_(((1 - m) * bn.running_mean.double() + m * mean).float())
        bn.running_var.copy_(((1 - m) * bn.running_var.double() + m * var * cnt / max(cnt - 1, 1)).float())
        bn.num_batches_tracked += 1
    else:
        mean, var = bn.running_mean.double(), bn.running_var.double()
    invstd = 1.0 / torch.sqrt(var + bn.eps)
    y = (rows - mean.view(1, -1, 1)) * (invstd * bn.weight.double()).view(1, -1, 1) + bn.bias.double().view(1, -1, 1)
    out.copy_(_rows_to_nchw(y, N, C, T, V, M).to(out.dtype))
    save_mean.copy_(mean.float())
    save_invstd.copy_(invstd.float())


def data_bn_bwd(g, x, num_point, fold_m, gamma, mean, invstd, train, dgamma, dbeta, dx):
    x5 = _x5(x, num_point)
    N, C, T, V, M = x5.shape
    rows = _data_bn_rows(x5, fold_m).double()
    grow = g.double().reshape(N, M, C, T, V).permute(0, 1, 4, 2, 3)                    # (N, M, V, C, T)
    grow = grow.reshape(N * M, V * C, T) if fold_m else grow.reshape(N, M * V * C, T)
    xh = (rows - mean.double().view(1, -1, 1)) * invstd.double().view(1, -1, 1)
    s1, s2 = grow.sum((0, 2)), (grow * xh).sum((0, 2))
    if dgamma is not None:
        dgamma += s2.float()
    if dbeta is not None:
        dbeta += s1.float()
    if dx is not None:
        cnt = rows.shape[0] * rows.shape[2]
        a = (gamma.double() * invstd.double()).view(1, -1, 1)
        d = a * (grow - (s1 / cnt).view(1, -1, 1) - xh * (s2 / cnt).view(1, -1, 1)) if train else a * grow
        d = d.reshape(N, M, V, C, T).permute(0, 3, 4, 2, 1)                            # (N, C, T, V, M)
        dx.copy_(d.float())


def pool_fc_fwd(x, M, W, b, pooled, logits, gate=None):
    NM, C, T, V = x.shape
    p = x.float().reshape(NM // M, M, C, T * V).mean(3).mean(1)
    pooled.copy_(p)
    if W is not None:
        logits.copy_(F.linear(p if gate is None else p * gate, W, b))


def pool_fc_bwd(dlogits, pooled, W, M, g, dW, db, gate=None, dgate=None):
    dp = dlogits if W is None else dlogits @ W
    gt = 1.0 if gate is None else gate
    if dW is not None:
        dW += dlogits.t() @ (pooled * gt)
    if db is not None:
        db += dlogits.sum(0)
    if dgate is not None:
        dgate.copy_(dp * pooled)
    if g is not None:
        NM, C, T, V = g.shape
        g.copy_((dp * gt / (M * T * V)).view(NM // M, 1, C, 1, 1).expand(NM // M, M, C, T, V).reshape(g.shape).to(g.dtype))


def transpose_act(inp, out, mode=0, aux=None):
    v = inp
    if mode == 1:
        v = torch.sigmoid(inp)
    elif mode == 2:
        v = inp * aux * (1 - aux)
    out.copy_(v.t())


def softmax_ce_fwd(logits, labels, loss, dl):
    valid = (labels >= 0) & (labels < logits.shape[1])
    cnt = int(valid.sum())
    lsm = torch.log_softmax(logits.double(), 1)
    safe = labels.clamp(0, logits.shape[1] - 1)
    nll = -lsm.gather(1, safe.view(-1, 1)).view(-1) * valid
    loss.copy_((nll.sum() / cnt).float().reshape(1))
    if dl is not None:
        d = lsm.exp()
        d[torch.arange(len(labels)), safe] -= 1.0
        dl.copy_((d * valid.view(-1, 1) / cnt).float())


def softmax_ce_bwd(dl, gloss, out):
    out.copy_(dl * gloss.reshape(()))


def sgd_step(P, G, Mo, lr, momentum, weight_decay, nesterov, grad_scale=1.0):
    with torch.no_grad():
        g = G * grad_scale + weight_decay * P
        Mo.mul_(momentum).add_(g)
        P.sub_(float(lr) * (g + momentum * Mo if nesterov else Mo))


def feeder_nucla(raw, length, sample, view, frame_idx, bone_parent, mode, out):
    import math
    B, T = frame_idx.shape
    V = raw.shape[2]
    for b in range(B):
        s = int(sample[b])
        L = int(length[s])
        val = raw[s, :L].double()
        val = val - val[0, 1]
        ax, ay, sc = math.radians(float(view[b, 0])), math.radians(float(view[b, 1])), float(view[b, 2])
        Rx = torch.tensor([[1, 0, 0], [0, math.cos(ax), math.sin(ax)], [0, -math.sin(ax), math.cos(ax)]], dtype=torch.float64)
        Ry = torch.tensor([[math.cos(ay), 0, -math.sin(ay)], [0, 1, 0], [math.sin(ay), 0, math.cos(ay)]], dtype=torch.float64)
        X = val.reshape(-1, 3) @ (Ry @ Rx * sc).to(val.device)
        lo, hi = X.min(0).values, X.max(0).values
        X = ((X - lo) / (hi - lo + 1e-6) * 2 - 1).reshape(L, V, 3)
        d = X[frame_idx[b].long()]
        if mode == 1:
            par = bone_parent.long()
            d = torch.where((par >= 0)[None, :, None], d - d[:, par.clamp_min(0)], torch.zeros_like(d))
        elif mode == 2:
            d = torch.cat([d[1:] - d[:-1], torch.zeros_like(d[:1])])
        out[b] = d.permute(2, 0, 1).unsqueeze(-1).float()


ALL = ['conv_pack_weights', 'conv_fwd', 'conv_dgrad', 'conv_wgrad', 'mean_t', 'ctrgc_fwd', 'ctrgc_bwd', 'bn_finalize', 'bn_bwd_coef',
       'gcn_epilogue_fwd', 'gcn_epilogue_bwd', 'gcn_mid_bwd', 'tcn_epilogue_fwd', 'tcn_epilogue_bwd', 'maxpool_fwd',
       'maxpool_bwd', 'graph_agg_fwd', 'graph_agg_bwd', 'data_bn_fwd', 'data_bn_bwd', 'pool_fc_fwd', 'pool_fc_bwd',
       'softmax_ce_fwd', 'softmax_ce_bwd', 'sgd_step', 'transpose_act', 'feeder_nucla', 'coef_diff']


def install(monkeypatch=None):
    """Route tam_gcn_b200.functional through the emulation (CPU orchestration tests).  Without a monkeypatch fixture
    (spawned worker processes) the attributes are simply overwritten."""
    import tam_gcn_b200.ops as real
    import tam_gcn_b200.functional as Fn
    g = globals()
    put = monkeypatch.setattr if monkeypatch is not None else setattr
    for name in ALL:
        put(real, name, g[name])
    put(Fn, '_require_cuda', lambda x: None)
