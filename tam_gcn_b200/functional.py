"""Autograd functions of the CTR-GCN / ST-GCN hot path: host-side orchestration of the CUDA kernels.

Each `torch.autograd.Function` below covers one reference module end to end (forward AND a
hand-written backward) as a short sequence of fused kernels from libtamgcn.so:

    UnitGcnFn      unit_gcn                       models/ctrgcn.py:246-263
    CtrgcFn        CTRGC (stand-alone use)        models/ctrgcn.py:172-177
    MsTcnFn        MultiScale_TemporalConv (+ the residual add / ReLU tail of TCN_GCN_unit)
                                                  models/ctrgcn.py:137-147, 282-284
    ConvBnFn       unit_tcn / TemporalConv        models/ctrgcn.py:66-69, 191-193
    CtgFn          ConvTemporalGraphical          models/stgcn.py:57-63
    StGcnFn        st_gcn                         models/stgcn.py:95-99

BatchNorm never runs as its own pass: statistics are reduced in the epilogue of the kernel that
produces a tensor, tiny coefficient kernels turn them into per-channel (scale, shift) resp. the
backward affine (A, B, C), and the consumer kernel applies them while loading (`ops.Opnd`).

No arithmetic on activations happens in Python/ATen here; the only torch ops are allocations, the
packing (cat/stack) of a few KB of parameters, and O(C) coefficient algebra.
"""
import torch
import torch.nn as nn

from . import arena, ops
from .ops import Opnd, RES_NONE, RES_IDENTITY, RES_AFFINE


# ------------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------------
def _check_input(x):
    if not x.is_cuda:
        raise RuntimeError('tam_gcn_b200 modules run on CUDA (sm_100a) only; there is no CPU path')
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError('tam_gcn_b200 supports float32 / bfloat16 activations, got %s' % x.dtype)
    if x.dim() != 4:
        raise ValueError('expected (N, C, T, V) input, got shape %s' % (tuple(x.shape),))
    return x.contiguous()


def _empty(shape, like, dtype=None):
    return torch.empty(shape, device=like.device, dtype=dtype or like.dtype)


def _zeros(shape, like, dtype):
    """Zero-initialised accumulator; inside an engine step it is a slice of the once-per-step cleared arena."""
    return arena.zeros(shape, dtype, like.device)


def _packed(mod, key, params, shape):
    """torch.cat of `params` (flattened) viewed as `shape`, without the copy: the parameters' storage is moved, once,
    into one flat buffer that they become views of (same Parameter objects, same names and shapes, so state_dict,
    optimisers and load_state_dict are unaffected).  The pointers are re-checked on every call (`.to()`, `.data`
    swaps re-trigger the move); non-leaf tensors (nn.DataParallel replicas) fall back to torch.cat."""
    packs = mod.__dict__.setdefault('_tamgcn_packs', {})
    ent = packs.get(key)
    if ent is not None:
        flat, ptrs = ent
        if len(ptrs) == len(params) and all(p.data_ptr() == q for p, q in zip(params, ptrs)):
            return flat.view(shape)
    movable = all(isinstance(p, nn.Parameter) and p.is_leaf and p.is_contiguous() for p in params) \
        and not (params[0].is_cuda and torch.cuda.is_current_stream_capturing())
    if not movable:
        return torch.cat([p.reshape(-1) for p in params]).view(shape)
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        off = 0
        for p in params:
            n = p.numel()
            p.data = flat[off:off + n].view(p.shape)
            off += n
    packs[key] = (flat, [p.data_ptr() for p in params])
    return flat.view(shape)


def _w2(conv):
    """Conv2d weight (Cout, Cin, k, 1) -> (Cout, Cin*k) fp32 view."""
    w = conv.weight
    if w.dtype != torch.float32:
        raise TypeError('parameters must be float32 (master weights); got %s' % w.dtype)
    return w.reshape(w.shape[0], -1)


def _pack(W2d, k, like, stride=1):
    """Tensor-core weight tiles (wpack_fwd, wpack_dgrad) when activations are bf16; (None, None) in fp32 mode
    (and per entry for the shapes whose kernel reads the fp32 weights directly)."""
    if like.dtype != torch.bfloat16:
        return None, None
    return ops.conv_pack_weights(W2d, W2d.shape[0], W2d.shape[1] // k, k, stride, like.shape[3])


def _bias(conv, like):
    if conv.bias is None:
        return torch.zeros(conv.weight.shape[0], device=like.device, dtype=torch.float32)
    return conv.bias


def _bn_group(bns):
    m, e = bns[0].momentum, bns[0].eps
    for b in bns:
        if b.momentum != m or b.eps != e or b.momentum is None:
            raise NotImplementedError('BatchNorm layers fused in one kernel must share a float momentum and eps')
        if not b.track_running_stats or not b.affine:
            raise NotImplementedError('BatchNorm without affine / running statistics is not supported')
    return m, e


class _BnCoef:
    """Per-channel forward coefficients of a (concatenated) group of BatchNorms: rows scale, shift, mean, invstd."""

    def __init__(self, C, like):
        self.t = torch.empty(4, C, device=like.device, dtype=torch.float32)
        self.scale, self.shift, self.mean, self.invstd = self.t[0], self.t[1], self.t[2], self.t[3]


def _bn_forward(bns, slices, coef, stats, count, train):
    """Finalize BatchNorms `bns` (channel ranges `slices` of the concatenated coefficient rows)."""
    m, e = _bn_group(bns)
    descs = []
    for bn, sl in zip(bns, slices):
        d = dict(gamma=bn.weight, beta=bn.bias, rmean=bn.running_mean, rvar=bn.running_var,
                 nbt=bn.num_batches_tracked if train else None,
                 scale=coef.scale[sl], shift=coef.shift[sl], mean=coef.mean[sl], invstd=coef.invstd[sl])
        if train:
            d['sum'], d['sumsq'] = stats[0][sl], stats[1][sl]
        descs.append(d)
    ops.bn_finalize(descs, count, m, e, train)


class _BnBwd:
    """Backward affine dY = A*dYhat + B*Y + C of a group of BatchNorms, plus dgamma / dbeta."""

    def __init__(self, C, like):
        self.t = torch.empty(5, C, device=like.device, dtype=torch.float32)
        self.A, self.B, self.C, self.dgamma, self.dbeta = self.t[0], self.t[1], self.t[2], self.t[3], self.t[4]


def _bn_backward(bns, slices, coef, bw, s1, s2, count, train):
    descs = []
    for bn, sl in zip(bns, slices):
        descs.append(dict(s1=s1[sl], s2=s2[sl], gamma=bn.weight, mean=coef.mean[sl], invstd=coef.invstd[sl],
                          A=bw.A[sl], B=bw.B[sl], Cc=bw.C[sl], dgamma=bw.dgamma[sl], dbeta=bw.dbeta[sl]))
    ops.bn_bwd_coef(descs, count, train)


def _full(sl_c):
    return slice(0, sl_c)


def _conv_out_len(T, k, s, d, p):
    return (T + 2 * p - d * (k - 1) - 1) // s + 1


# ------------------------------------------------------------------------------------------------
# CTRGC group (K subsets sharing the input x): shared by unit_gcn (K=3) and stand-alone CTRGC (K=1)
# ------------------------------------------------------------------------------------------------
def _ctrgc_pack(convs, like, extra=None):
    """Stack the parameters of K CTRGC modules.  extra: optional Conv2d (unit_gcn.down[0]) appended to conv3."""
    K = len(convs)
    R, Cin = convs[0].conv1.weight.shape[:2]
    Cout = convs[0].conv3.weight.shape[0]
    for c in convs:
        _w2(c.conv1), _w2(c.conv2), _w2(c.conv3), _w2(c.conv4)                                # dtype checks
    own = convs[0]                                                                            # the packs hang off the first CTRGC
    W12 = _packed(own, 'W12', [c.conv1.weight for c in convs] + [c.conv2.weight for c in convs], (2 * K * R, Cin))
    b12 = _packed(own, 'b12', [_bias(c.conv1, like) for c in convs] + [_bias(c.conv2, like) for c in convs], (2 * K * R,))
    w3 = [c.conv3.weight for c in convs]
    b3 = [_bias(c.conv3, like) for c in convs]
    Cw = K * Cout
    if extra is not None:
        _w2(extra)
        w3.append(extra.weight)
        b3.append(_bias(extra, like))
        Cw += extra.weight.shape[0]
    W3 = _packed(own, 'W3', w3, (Cw, Cin))                                                    # (K*Cout [+Cd], Cin)
    b3 = _packed(own, 'b3', b3, (Cw,))
    W4 = _packed(own, 'W4', [c.conv4.weight for c in convs], (K, Cout, R))
    b4 = _packed(own, 'b4', [_bias(c.conv4, like) for c in convs], (K, Cout))
    return K, R, Cin, Cout, W12, b12, W3, b3, W4, b4


def _ctrgc_unpack_grads(convs, K, R, Cout, dW12, db12, dW3, db3, dW4, db4):
    """Per-module gradient views, in the order conv1.w, conv1.b, conv2.w, conv2.b, conv3.w, conv3.b, conv4.w, conv4.b."""
    out = []
    for i, c in enumerate(convs):
        out += [dW12[i * R:(i + 1) * R].view_as(c.conv1.weight), db12[i * R:(i + 1) * R] if c.conv1.bias is not None else None,
                dW12[(K + i) * R:(K + i + 1) * R].view_as(c.conv2.weight),
                db12[(K + i) * R:(K + i + 1) * R] if c.conv2.bias is not None else None,
                dW3[i * Cout:(i + 1) * Cout].view_as(c.conv3.weight),
                db3[i * Cout:(i + 1) * Cout] if c.conv3.bias is not None else None,
                dW4[i].view_as(c.conv4.weight), db4[i] if c.conv4.bias is not None else None]
    return out


def ctrgc_params(c):
    return [c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias, c.conv3.weight, c.conv3.bias,
            c.conv4.weight, c.conv4.bias]


# ------------------------------------------------------------------------------------------------
# unit_gcn
# ------------------------------------------------------------------------------------------------
def unit_gcn_params(mod):
    """Flat parameter list of a unit_gcn in the order UnitGcnFn.backward returns gradients."""
    ps = []
    for c in mod.convs:
        ps += ctrgc_params(c)
    if mod.has_down:
        ps += [mod.down[0].weight, mod.down[0].bias, mod.down[1].weight, mod.down[1].bias]
    ps += [mod.offset_conv[0].weight, mod.offset_conv[0].bias, mod.offset_conv[1].weight, mod.offset_conv[1].bias,
           mod.bn.weight, mod.bn.bias, mod.alpha]
    if mod.adaptive:
        ps.append(mod.PA)
    return ps


class UnitGcnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod, *params):
        x = _check_input(x)
        N, Cin, T, V = x.shape
        train = mod.training
        convs = list(mod.convs)
        has_down = mod.has_down
        K, R, Cin_w, Cout, W12, b12, W3, b3, W4, b4 = _ctrgc_pack(convs, x, mod.down[0] if has_down else None)
        if Cin_w != Cin:
            raise ValueError('unit_gcn: input has %d channels, module expects %d' % (Cin, Cin_w))
        PA = (mod.PA if mod.adaptive else mod.A).to(torch.float32).contiguous()
        alpha = mod.alpha
        count = N * T * V
        KC = K * Cout

        # x1 / x2: 1x1 convs on the T-mean of x
        m = _empty((N, Cin, 1, V), x, torch.float32)
        ops.mean_t(x, m)
        x12 = _empty((N, 2 * K * R, 1, V), x, torch.float32)
        ops.conv_fwd(m, W12, b12, x12)
        # conv3 of the K subsets (+ down conv) in one pass over x
        Cw = KC + (Cout if has_down else 0)
        xw = _empty((N, Cw, T, V), x)
        stats = _zeros((6, Cout), x, torch.float64) if train else None   # rows: down(sum,sq), bn(sum,sq), offset(sum,sq)
        pk3 = _pack(W3, 1, x)
        ops.conv_fwd(x, W3, b3, xw, stats=(stats[0], stats[1]) if (train and has_down) else None, stat_c0=KC,
                     wpack=pk3[0])
        # fused topology refinement + aggregation (+ BN statistics of y0)
        y0 = _empty((N, Cout, T, V), x)
        ops.ctrgc_fwd(xw[:, :KC], x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y0,
                      stats=(stats[2], stats[3]) if train else None)
        cg = _BnCoef(Cout, x)
        cd = _BnCoef(Cout, x) if has_down else None
        if has_down:
            if train:
                _bn_forward([mod.down[1]], [_full(Cout)], cd, (stats[0], stats[1]), count, True)
                _bn_forward([mod.bn], [_full(Cout)], cg, (stats[2], stats[3]), count, True)
            else:
                _bn_forward([mod.down[1]], [_full(Cout)], cd, None, count, False)
                _bn_forward([mod.bn], [_full(Cout)], cg, None, count, False)
        else:
            _bn_forward([mod.bn], [_full(Cout)], cg, (stats[2], stats[3]) if train else None, count, train)
        # offset branch: z = W_o (res - y) + b_o, with res - y formed while loading
        nsg = -cg.scale
        if has_down:
            diff = Opnd(xw[:, KC:], y0, a=cd.scale, b=nsg, c=cd.shift - cg.shift)
            res_mode, r, sr, hr = RES_AFFINE, xw[:, KC:], cd.scale, cd.shift
        elif mod.residual_identity:
            diff = Opnd(x, y0, a=None, b=nsg, c=-cg.shift)
            res_mode, r, sr, hr = RES_IDENTITY, x, None, None
        else:
            diff = Opnd(y0, None, a=nsg, c=-cg.shift)
            res_mode, r, sr, hr = RES_NONE, None, None, None
        oc = mod.offset_conv[0]
        Wo, bo = _w2(oc), _bias(oc, x)
        z = _empty((N, Cout, T, V), x)
        pko = _pack(Wo, 1, x)
        ops.conv_fwd(diff, Wo, bo, z, stats=(stats[4], stats[5]) if train else None, wpack=pko[0])
        co = _BnCoef(Cout, x)
        _bn_forward([mod.offset_conv[1]], [_full(Cout)], co, (stats[4], stats[5]) if train else None, count, train)
        out = _empty((N, Cout, T, V), x)
        ops.gcn_epilogue_fwd(y0, cg.scale, cg.shift, z, co.scale, co.shift, res_mode, r, sr, hr, out)

        ctx.mod, ctx.train, ctx.dims = mod, train, (N, Cin, Cout, T, V, K, R)
        ctx.res_mode = res_mode
        ctx.diff = diff
        ctx.coefs = (cg, cd, co)
        ctx.packed = (W12, W3, W4, b4, PA, Wo, pk3[1], pko[1])
        ctx.save_for_backward(x, m, x12, xw, y0, z, out)
        return out

    @staticmethod
    def backward(ctx, g):
        mod, train = ctx.mod, ctx.train
        N, Cin, Cout, T, V, K, R = ctx.dims
        x, m, x12, xw, y0, z, out = ctx.saved_tensors
        cg, cd, co = ctx.coefs
        W12, W3, W4, b4, PA, Wo, pk3d, pkod = ctx.packed
        has_down = cd is not None
        res_mode = ctx.res_mode
        KC, count = K * Cout, N * T * V
        g = g.contiguous()
        if g.dtype != x.dtype:
            g = g.to(x.dtype)
        sb = _zeros((6, Cout), x, torch.float64)      # rows: offset(s1,s2), bn(s1,s2), down(s1,s2)

        # tail: ReLU mask, tanh', BN_o backward sums
        G = _empty(g.shape, x)
        DZ = _empty(g.shape, x)
        ops.gcn_epilogue_bwd(g, out, z, co.scale, co.shift, G, DZ, sb[0], sb[1])
        bo_ = _BnBwd(Cout, x)
        _bn_backward([mod.offset_conv[1]], [_full(Cout)], co, bo_, sb[0], sb[1], count, train)
        dz = Opnd(DZ, z, a=bo_.A, b=bo_.B, c=bo_.C)
        # offset conv backward
        dWo = _zeros(Wo.shape, x, torch.float32)
        dbo = _zeros((Cout,), x, torch.float32)
        ops.conv_wgrad(dz, ctx.diff, dWo, dbo)
        DD = _empty(g.shape, x)
        ops.conv_dgrad(dz, Wo, DD, wpack=pkod)
        # dY = G - DD (grad wrt bn output), dRes = G + DD; BN / down.BN backward sums
        Cw = xw.shape[1]
        dxw = _empty(xw.shape, x)
        if has_down:
            ops.gcn_mid_bwd(G, DD, dxw[:, KC:], y0, xw[:, KC:], sb[2], sb[3], sb[4], sb[5])
            dres = None
        elif res_mode == RES_IDENTITY:
            dres = _empty(g.shape, x)
            ops.gcn_mid_bwd(G, DD, dres, y0, None, sb[2], sb[3], None, None)
        else:
            dres = None
            ops.gcn_mid_bwd(G, DD, None, y0, None, sb[2], sb[3], None, None)
        bg = _BnBwd(Cout, x)
        bd = _BnBwd(Cout, x) if has_down else None
        _bn_backward([mod.bn], [_full(Cout)], cg, bg, sb[2], sb[3], count, train)
        if has_down:
            _bn_backward([mod.down[1]], [_full(Cout)], cd, bd, sb[4], sb[5], count, train)
        # fused CTRGC backward
        dy = Opnd(G, y0, a=bg.A, b=bg.B, c=bg.C)
        dx12 = _zeros(x12.shape, x, torch.float32)
        dW4 = _zeros(W4.shape, x, torch.float32)
        db4 = _zeros(b4.shape, x, torch.float32)
        dPA = _zeros(PA.shape, x, torch.float32)
        dalpha = _zeros((1,), x, torch.float32)
        ops.ctrgc_bwd(dy, xw[:, :KC], x12[:, :K * R], x12[:, K * R:], W4, b4, PA, mod.alpha, dxw[:, :KC],
                      dx12[:, :K * R], dx12[:, K * R:], dW4, db4, dPA, dalpha)
        # conv1/conv2 backward on the T-mean
        dW12 = _zeros(W12.shape, x, torch.float32)
        db12 = _zeros((W12.shape[0],), x, torch.float32)
        ops.conv_wgrad(dx12, m, dW12, db12)
        dm = _empty(m.shape, x, torch.float32)
        ops.conv_dgrad(dx12, W12, dm)
        # conv3 (+down) backward: one wgrad and one dgrad over the concatenated channels
        if has_down:
            a = torch.ones(Cw, device=x.device, dtype=torch.float32)
            b = torch.zeros(Cw, device=x.device, dtype=torch.float32)
            c = torch.zeros(Cw, device=x.device, dtype=torch.float32)
            a[KC:], b[KC:], c[KC:] = bd.A, bd.B, bd.C
            dxw_op = Opnd(dxw, xw, a=a, b=b, c=c)
        else:
            dxw_op = Opnd(dxw)
        dW3 = _zeros(W3.shape, x, torch.float32)
        db3 = _zeros((Cw,), x, torch.float32)
        ops.conv_wgrad(dxw_op, x, dW3, db3)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dxw_op, W3, dx, addend=dres, bcast=dm, bcast_scale=1.0 / T, wpack=pk3d)

        grads = _ctrgc_unpack_grads(list(mod.convs), K, R, Cout, dW12, db12, dW3, db3, dW4, db4)
        if has_down:
            d0 = mod.down[0]
            grads += [dW3[KC:].view_as(d0.weight), db3[KC:] if d0.bias is not None else None, bd.dgamma, bd.dbeta]
        oc = mod.offset_conv[0]
        grads += [dWo.view_as(oc.weight), dbo if oc.bias is not None else None, bo_.dgamma, bo_.dbeta,
                  bg.dgamma, bg.dbeta, dalpha]
        if mod.adaptive:
            grads.append(dPA)
        return (dx, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# stand-alone CTRGC   forward(x, A=None, alpha=1)
# ------------------------------------------------------------------------------------------------
class CtrgcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, A, alpha, mod, *params):
        x = _check_input(x)
        N, Cin, T, V = x.shape
        K, R, Cin_w, Cout, W12, b12, W3, b3, W4, b4 = _ctrgc_pack([mod], x)
        if Cin_w != Cin:
            raise ValueError('CTRGC: input has %d channels, module expects %d' % (Cin, Cin_w))
        PA = A.to(torch.float32).reshape(1, V, V).contiguous()
        al = alpha.to(torch.float32).reshape(1).contiguous()
        m = _empty((N, Cin, 1, V), x, torch.float32)
        ops.mean_t(x, m)
        x12 = _empty((N, 2 * R, 1, V), x, torch.float32)
        ops.conv_fwd(m, W12, b12, x12)
        x3 = _empty((N, Cout, T, V), x)
        pk3 = _pack(W3, 1, x)
        ops.conv_fwd(x, W3, b3, x3, wpack=pk3[0])
        y = _empty((N, Cout, T, V), x)
        ops.ctrgc_fwd(x3, x12[:, :R], x12[:, R:], W4, b4, PA, al, y)
        ctx.mod, ctx.dims = mod, (N, Cin, Cout, T, V, R)
        ctx.packed = (W12, W3, W4, b4, PA, al, pk3[1])
        ctx.save_for_backward(x, m, x12, x3)
        ctx.a_shape, ctx.alpha_shape = A.shape, alpha.shape
        return y

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        N, Cin, Cout, T, V, R = ctx.dims
        x, m, x12, x3 = ctx.saved_tensors
        W12, W3, W4, b4, PA, al, pk3d = ctx.packed
        g = g.contiguous().to(x.dtype)
        dx3 = _empty(x3.shape, x)
        dx12 = _zeros(x12.shape, x, torch.float32)
        dW4 = _zeros(W4.shape, x, torch.float32)
        db4 = _zeros(b4.shape, x, torch.float32)
        dPA = _zeros(PA.shape, x, torch.float32)
        dalpha = _zeros((1,), x, torch.float32)
        ops.ctrgc_bwd(Opnd(g), x3, x12[:, :R], x12[:, R:], W4, b4, PA, al, dx3, dx12[:, :R], dx12[:, R:], dW4, db4,
                      dPA, dalpha)
        dW12 = _zeros(W12.shape, x, torch.float32)
        db12 = _zeros((2 * R,), x, torch.float32)
        ops.conv_wgrad(dx12, m, dW12, db12)
        dm = _empty(m.shape, x, torch.float32)
        ops.conv_dgrad(dx12, W12, dm)
        dW3 = _zeros(W3.shape, x, torch.float32)
        db3 = _zeros((Cout,), x, torch.float32)
        ops.conv_wgrad(dx3, x, dW3, db3)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dx3, W3, dx, bcast=dm, bcast_scale=1.0 / T, wpack=pk3d)
        grads = _ctrgc_unpack_grads([mod], 1, R, Cout, dW12, db12, dW3, db3, dW4, db4)
        return (dx, dPA.reshape(ctx.a_shape), dalpha.reshape(ctx.alpha_shape), None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# conv (k x 1) + BN   (unit_tcn, TemporalConv)
# ------------------------------------------------------------------------------------------------
def _conv_geom(conv):
    k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
    if conv.kernel_size[1] != 1 or conv.stride[1] != 1 or conv.padding[1] != 0 or conv.groups != 1:
        raise NotImplementedError('only (k x 1) ungrouped temporal convolutions are supported')
    return k, s, d, p


class ConvBnFn(torch.autograd.Function):
    """out = BN(conv(x));  params = (conv.weight, conv.bias, bn.weight, bn.bias)."""

    @staticmethod
    def forward(ctx, x, conv, bn, *params):
        x = _check_input(x)
        N, Cin, T, V = x.shape
        k, s, d, p = _conv_geom(conv)
        Cout = conv.weight.shape[0]
        To = _conv_out_len(T, k, s, d, p)
        train = bn.training
        W, b = _w2(conv), _bias(conv, x)
        raw = _empty((N, Cout, To, V), x)
        stats = _zeros((2, Cout), x, torch.float64) if train else None
        pk = _pack(W, k, x, s)
        ops.conv_fwd(x, W, b, raw, k, s, d, p, stats=stats, wpack=pk[0])
        cf = _BnCoef(Cout, x)
        _bn_forward([bn], [_full(Cout)], cf, stats, N * To * V, train)
        out = _empty(raw.shape, x)
        ops.tcn_epilogue_fwd(raw, cf.scale, cf.shift, RES_NONE, None, None, None, False, out)
        ctx.conv, ctx.bn, ctx.train, ctx.geom, ctx.cf, ctx.pkd = conv, bn, train, (k, s, d, p), cf, pk[1]
        ctx.save_for_backward(x, raw)
        return out

    @staticmethod
    def backward(ctx, g):
        conv, bn, train = ctx.conv, ctx.bn, ctx.train
        k, s, d, p = ctx.geom
        x, raw = ctx.saved_tensors
        N, Cout, To, V = raw.shape
        g = g.contiguous().to(x.dtype)
        sb = _zeros((2, Cout), x, torch.float64)
        ops.tcn_epilogue_bwd(g, None, False, raw, None, None, sb[0], sb[1], None)
        bw = _BnBwd(Cout, x)
        _bn_backward([bn], [_full(Cout)], ctx.cf, bw, sb[0], sb[1], N * To * V, train)
        dy = Opnd(g, raw, a=bw.A, b=bw.B, c=bw.C)
        W = _w2(conv)
        dW = _zeros(W.shape, x, torch.float32)
        db = _zeros((Cout,), x, torch.float32)
        ops.conv_wgrad(dy, x, dW, db, k, s, d, p)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dy, W, dx, k, s, d, p, wpack=ctx.pkd)
        return dx, None, None, dW.view_as(conv.weight), (db if conv.bias is not None else None), bw.dgamma, bw.dbeta


# ------------------------------------------------------------------------------------------------
# MultiScale_TemporalConv  (+ fused residual add / ReLU of the enclosing TCN_GCN_unit)
# ------------------------------------------------------------------------------------------------
def ms_tcn_params(mod):
    ps = []
    for j in range(mod.num_dil):
        br = mod.branches[j]
        ps += [br[0].weight, br[0].bias, br[1].weight, br[1].bias, br[3].conv.weight, br[3].conv.bias,
               br[3].bn.weight, br[3].bn.bias]
    br = mod.branches[mod.num_dil]
    ps += [br[0].weight, br[0].bias, br[1].weight, br[1].bias, br[4].weight, br[4].bias]
    br = mod.branches[mod.num_dil + 1]
    ps += [br[0].weight, br[0].bias, br[1].weight, br[1].bias]
    return ps


def res_conv_params(rm):
    return [rm.conv.weight, rm.conv.bias, rm.bn.weight, rm.bn.bias]


class MsTcnFn(torch.autograd.Function):
    """out = f( cat_j BN_j(branch_j(x)) + res ),  f = ReLU if relu.

    res_kind: 'none' | 'identity' (res = r_in) | 'conv' (res = BN(conv(r_in)), `res_mod` has .conv / .bn).
    r_in is None when the residual source is x itself (stand-alone module).
    params = ms_tcn_params(mod) + (res_conv_params(res_mod) if res_kind == 'conv').
    """

    @staticmethod
    def forward(ctx, x, r_in, mod, res_kind, res_mod, relu, *params):
        x = _check_input(x)
        N, Cin, T, V = x.shape
        train = mod.training
        nd, Cb, s = mod.num_dil, mod.branch_channels, mod.stride
        nb = nd + 2
        Cout, Ch = nb * Cb, (nd + 1) * Cb
        To = _conv_out_len(T, 1, s, 1, 0)
        r_src = x if r_in is None else _check_input(r_in)

        # all 1x1 branch heads that keep T in one pass over x
        heads = [mod.branches[j][0] for j in range(nd + 1)]
        for c in heads:
            _w2(c)
        Wh = _packed(mod, 'Wh', [c.weight for c in heads], ((nd + 1) * Cb, Cin))
        bh = _packed(mod, 'bh', [_bias(c, x) for c in heads], ((nd + 1) * Cb,))
        h = _empty((N, Ch, T, V), x)
        st_h = _zeros((2, Ch), x, torch.float64) if train else None
        pkh = _pack(Wh, 1, x)
        ops.conv_fwd(x, Wh, bh, h, stats=st_h, wpack=pkh[0])
        u = _empty((N, Cout, To, V), x)
        st_u = _zeros((2, Cout), x, torch.float64) if train else None
        # strided 1x1 branch straight into its slice of u
        c3 = mod.branches[nd + 1][0]
        W3, b3 = _w2(c3), _bias(c3, x)
        pk3 = _pack(W3, 1, x)
        ops.conv_fwd(x, W3, b3, u[:, Ch:], 1, s, 1, 0, stats=(st_u[0][Ch:], st_u[1][Ch:]) if train else None,
                     wpack=pk3[0])
        ch = _BnCoef(Ch, x)
        sl = [slice(j * Cb, (j + 1) * Cb) for j in range(nb)]
        _bn_forward([mod.branches[j][1] for j in range(nd + 1)], sl[:nd + 1], ch, st_h, N * T * V, train)
        geoms = []
        for j in range(nd):
            tc = mod.branches[j][3].conv
            k, cs, d, p = _conv_geom(tc)
            if cs != s or _conv_out_len(T, k, cs, d, p) != To:
                raise ValueError('MultiScale_TemporalConv: branch %d output length differs' % j)
            pkt = _pack(_w2(tc), k, x, cs)
            geoms.append((k, cs, d, p, pkt[1]))
            ops.conv_fwd(Opnd(h[:, sl[j]], a=ch.scale[sl[j]], c=ch.shift[sl[j]], relu=True), _w2(tc), _bias(tc, x),
                         u[:, sl[j]], k, cs, d, p, stats=(st_u[0][sl[j]], st_u[1][sl[j]]) if train else None,
                         wpack=pkt[0])
        if _conv_out_len(T, 3, s, 1, 1) != To:
            raise ValueError('MultiScale_TemporalConv: max-pool branch output length differs')
        ops.maxpool_fwd(Opnd(h[:, sl[nd]], a=ch.scale[sl[nd]], c=ch.shift[sl[nd]], relu=True), u[:, sl[nd]], s,
                        stats=(st_u[0][sl[nd]], st_u[1][sl[nd]]) if train else None)
        cu = _BnCoef(Cout, x)
        final_bns = [mod.branches[j][3].bn for j in range(nd)] + [mod.branches[nd][4], mod.branches[nd + 1][1]]
        _bn_forward(final_bns, sl, cu, st_u, N * To * V, train)
        # residual
        cr = r_raw = None
        if res_kind == 'conv':
            rk, rs, rd, rp = _conv_geom(res_mod.conv)
            if _conv_out_len(r_src.shape[2], rk, rs, rd, rp) != To or res_mod.conv.weight.shape[0] != Cout:
                raise ValueError('residual branch shape mismatch')
            r_raw = _empty((N, Cout, To, V), x)
            st_r = _zeros((2, Cout), x, torch.float64) if res_mod.bn.training else None
            pkr = _pack(_w2(res_mod.conv), rk, x, rs)
            ops.conv_fwd(r_src, _w2(res_mod.conv), _bias(res_mod.conv, x), r_raw, rk, rs, rd, rp, stats=st_r,
                         wpack=pkr[0])
            cr = _BnCoef(Cout, x)
            _bn_forward([res_mod.bn], [_full(Cout)], cr, st_r, N * To * V, res_mod.bn.training)
            res_mode, r, sr, hr = RES_AFFINE, r_raw, cr.scale, cr.shift
        elif res_kind == 'identity':
            if r_src.shape != (N, Cout, To, V):
                raise ValueError('identity residual shape mismatch')
            res_mode, r, sr, hr = RES_IDENTITY, r_src, None, None
        else:
            res_mode, r, sr, hr = RES_NONE, None, None, None
        out = _empty((N, Cout, To, V), x)
        ops.tcn_epilogue_fwd(u, cu.scale, cu.shift, res_mode, r, sr, hr, relu, out)

        ctx.mod, ctx.res_mod, ctx.res_kind, ctx.relu, ctx.train = mod, res_mod, res_kind, relu, train
        ctx.geoms, ctx.coefs = geoms, (ch, cu, cr)
        ctx.packed = (Wh, W3, pkh[1], pk3[1], pkr[1] if res_kind == 'conv' else None)
        ctx.r_is_x = r_in is None
        ctx.save_for_backward(x, r_src if res_kind == 'conv' else None, h, u, r_raw, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        mod, res_mod, res_kind, relu, train = ctx.mod, ctx.res_mod, ctx.res_kind, ctx.relu, ctx.train
        x, r_src, h, u, r_raw, out = ctx.saved_tensors
        ch, cu, cr = ctx.coefs
        Wh, W3, pkhd, pk3d, pkrd = ctx.packed
        N, Cin, T, V = x.shape
        nd, Cb, s = mod.num_dil, mod.branch_channels, mod.stride
        nb = nd + 2
        Cout, Ch = nb * Cb, (nd + 1) * Cb
        To = u.shape[2]
        sl = [slice(j * Cb, (j + 1) * Cb) for j in range(nb)]
        g = g.contiguous().to(x.dtype)

        sb = _zeros((3, Cout), x, torch.float64)
        G = _empty(g.shape, x) if relu else None
        ops.tcn_epilogue_bwd(g, out, relu, u, r_raw, G, sb[0], sb[1], sb[2] if r_raw is not None else None)
        Gt = G if relu else g
        final_bns = [mod.branches[j][3].bn for j in range(nd)] + [mod.branches[nd][4], mod.branches[nd + 1][1]]
        bu = _BnBwd(Cout, x)
        _bn_backward(final_bns, sl, cu, bu, sb[0], sb[1], N * To * V, train)

        def dy_op(c0, c1):
            return Opnd(Gt[:, c0:c1], u[:, c0:c1], a=bu.A[c0:c1], b=bu.B[c0:c1], c=bu.C[c0:c1])

        DH = _empty(h.shape, x)
        sh = _zeros((2, Ch), x, torch.float64)
        tgrads = []
        for j in range(nd):
            tc = mod.branches[j][3].conv
            k, cs, d, p, pktd = ctx.geoms[j]
            dyj = dy_op(sl[j].start, sl[j].stop)
            hj = h[:, sl[j]]
            W = _w2(tc)
            dW = _zeros(W.shape, x, torch.float32)
            db = _zeros((Cb,), x, torch.float32)
            ops.conv_wgrad(dyj, Opnd(hj, a=ch.scale[sl[j]], c=ch.shift[sl[j]], relu=True), dW, db, k, cs, d, p)
            ops.conv_dgrad(dyj, W, DH[:, sl[j]], k, cs, d, p, mask=Opnd(hj, a=ch.scale[sl[j]], c=ch.shift[sl[j]]),
                           stats=(sh[0][sl[j]], sh[1][sl[j]]), wpack=pktd)
            tgrads.append((dW.view_as(tc.weight), db if tc.bias is not None else None))
        ops.maxpool_bwd(dy_op(sl[nd].start, sl[nd].stop),
                        Opnd(h[:, sl[nd]], a=ch.scale[sl[nd]], c=ch.shift[sl[nd]], relu=True), DH[:, sl[nd]], s,
                        stats=(sh[0][sl[nd]], sh[1][sl[nd]]))
        bh_ = _BnBwd(Ch, x)
        _bn_backward([mod.branches[j][1] for j in range(nd + 1)], sl[:nd + 1], ch, bh_, sh[0], sh[1], N * T * V, train)
        dh = Opnd(DH, h, a=bh_.A, b=bh_.B, c=bh_.C)
        dWh = _zeros(Wh.shape, x, torch.float32)
        dbh = _zeros((Ch,), x, torch.float32)
        ops.conv_wgrad(dh, x, dWh, dbh)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dh, Wh, dx, wpack=pkhd)
        dy3 = dy_op(Ch, Cout)
        dW3 = _zeros(W3.shape, x, torch.float32)
        db3 = _zeros((Cb,), x, torch.float32)
        ops.conv_wgrad(dy3, x, dW3, db3, 1, s, 1, 0)
        ops.conv_dgrad(dy3, W3, dx, 1, s, 1, 0, addend=dx, wpack=pk3d)

        dr = None
        rgrads = []
        if res_kind == 'conv':
            br = _BnBwd(Cout, x)
            _bn_backward([res_mod.bn], [_full(Cout)], cr, br, sb[0], sb[2], N * To * V, res_mod.bn.training)
            rk, rs, rd, rp = _conv_geom(res_mod.conv)
            dyr = Opnd(Gt, r_raw, a=br.A, b=br.B, c=br.C)
            Wr = _w2(res_mod.conv)
            dWr = _zeros(Wr.shape, x, torch.float32)
            dbr = _zeros((Cout,), x, torch.float32)
            ops.conv_wgrad(dyr, r_src, dWr, dbr, rk, rs, rd, rp)
            if ctx.r_is_x:
                ops.conv_dgrad(dyr, Wr, dx, rk, rs, rd, rp, addend=dx, wpack=pkrd)
            else:
                dr = _empty(r_src.shape, x)
                ops.conv_dgrad(dyr, Wr, dr, rk, rs, rd, rp, wpack=pkrd)
            rgrads = [dWr.view_as(res_mod.conv.weight), dbr if res_mod.conv.bias is not None else None, br.dgamma,
                      br.dbeta]
        elif res_kind == 'identity':
            if ctx.r_is_x:
                dx = dx + Gt
            else:
                dr = Gt

        grads = []
        for j in range(nd):
            br_ = mod.branches[j]
            grads += [dWh[sl[j]].view_as(br_[0].weight), dbh[sl[j]] if br_[0].bias is not None else None,
                      bh_.dgamma[sl[j]], bh_.dbeta[sl[j]], tgrads[j][0], tgrads[j][1], bu.dgamma[sl[j]], bu.dbeta[sl[j]]]
        br_ = mod.branches[nd]
        grads += [dWh[sl[nd]].view_as(br_[0].weight), dbh[sl[nd]] if br_[0].bias is not None else None,
                  bh_.dgamma[sl[nd]], bh_.dbeta[sl[nd]], bu.dgamma[sl[nd]], bu.dbeta[sl[nd]]]
        br_ = mod.branches[nd + 1]
        grads += [dW3.view_as(br_[0].weight), db3 if br_[0].bias is not None else None, bu.dgamma[sl[nd + 1]],
                  bu.dbeta[sl[nd + 1]]]
        grads += rgrads
        return (dx, dr, None, None, None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# ST-GCN graph layers
# ------------------------------------------------------------------------------------------------
class CtgFn(torch.autograd.Function):
    """ConvTemporalGraphical: out[n,c,t,w] = sum_{k,v} conv(x)[n,k*C+c,t,v] * A[k,v,w]."""

    @staticmethod
    def forward(ctx, x, A, mod, *params):
        x = _check_input(x)
        N, Cin, T, V = x.shape
        K = mod.kernel_size
        k, s, d, p = _conv_geom(mod.conv)
        KC = mod.conv.weight.shape[0]
        To = _conv_out_len(T, k, s, d, p)
        Af = A.to(torch.float32).contiguous()
        W, b = _w2(mod.conv), _bias(mod.conv, x)
        y = _empty((N, KC, To, V), x)
        pk = _pack(W, k, x, s)
        ops.conv_fwd(x, W, b, y, k, s, d, p, wpack=pk[0])
        out = _empty((N, KC // K, To, V), x)
        ops.graph_agg_fwd(y, Af, out)
        ctx.mod, ctx.geom, ctx.pkd = mod, (k, s, d, p), pk[1]
        ctx.save_for_backward(x, y, Af)
        ctx.a_dtype = A.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        k, s, d, p = ctx.geom
        x, y, Af = ctx.saved_tensors
        g = g.contiguous().to(x.dtype)
        dy = _empty(y.shape, x)
        dA = _zeros(Af.shape, x, torch.float32)
        ops.graph_agg_bwd(Opnd(g), y, Af, dy, dA)
        W = _w2(mod.conv)
        dW = _zeros(W.shape, x, torch.float32)
        db = _zeros((W.shape[0],), x, torch.float32)
        ops.conv_wgrad(dy, x, dW, db, k, s, d, p)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dy, W, dx, k, s, d, p, wpack=ctx.pkd)
        return dx, dA.to(ctx.a_dtype), None, dW.view_as(mod.conv.weight), (db if mod.conv.bias is not None else None)


def st_gcn_params(mod):
    ps = [mod.gcn.conv.weight, mod.gcn.conv.bias, mod.tcn[0].weight, mod.tcn[0].bias, mod.tcn[2].weight,
          mod.tcn[2].bias, mod.tcn[3].weight, mod.tcn[3].bias]
    if mod.res_kind == 'conv':
        ps += [mod.residual[0].weight, mod.residual[0].bias, mod.residual[1].weight, mod.residual[1].bias]
    return ps


class StGcnFn(torch.autograd.Function):
    """st_gcn block: relu( BN(conv9x1(relu(BN(graph_conv(x, A))))) + residual(x) )."""

    @staticmethod
    def forward(ctx, x, A, mod, *params):
        x = _check_input(x)
        N, Cin, T, V = x.shape
        train = mod.training
        K = mod.gcn.kernel_size
        gk, gs, gd, gp = _conv_geom(mod.gcn.conv)
        KC = mod.gcn.conv.weight.shape[0]
        Cout = KC // K
        Tg = _conv_out_len(T, gk, gs, gd, gp)
        Af = A.to(torch.float32).contiguous()
        Wg, bg = _w2(mod.gcn.conv), _bias(mod.gcn.conv, x)
        y = _empty((N, KC, Tg, V), x)
        pkg = _pack(Wg, gk, x, gs)
        ops.conv_fwd(x, Wg, bg, y, gk, gs, gd, gp, wpack=pkg[0])
        agg = _empty((N, Cout, Tg, V), x)
        st_a = _zeros((2, Cout), x, torch.float64) if train else None
        ops.graph_agg_fwd(y, Af, agg, stats=st_a)
        ca = _BnCoef(Cout, x)
        _bn_forward([mod.tcn[0]], [_full(Cout)], ca, st_a, N * Tg * V, train)
        tc = mod.tcn[2]
        k, s, d, p = _conv_geom(tc)
        To = _conv_out_len(Tg, k, s, d, p)
        u = _empty((N, Cout, To, V), x)
        st_u = _zeros((2, Cout), x, torch.float64) if train else None
        pkt = _pack(_w2(tc), k, x, s)
        ops.conv_fwd(Opnd(agg, a=ca.scale, c=ca.shift, relu=True), _w2(tc), _bias(tc, x), u, k, s, d, p, stats=st_u,
                     wpack=pkt[0])
        cu = _BnCoef(Cout, x)
        _bn_forward([mod.tcn[3]], [_full(Cout)], cu, st_u, N * To * V, train)
        cr = r_raw = None
        pkr = (None, None)
        if mod.res_kind == 'conv':
            rc = mod.residual[0]
            rk, rs, rd, rp = _conv_geom(rc)
            r_raw = _empty((N, Cout, To, V), x)
            st_r = _zeros((2, Cout), x, torch.float64) if train else None
            pkr = _pack(_w2(rc), rk, x, rs)
            ops.conv_fwd(x, _w2(rc), _bias(rc, x), r_raw, rk, rs, rd, rp, stats=st_r, wpack=pkr[0])
            cr = _BnCoef(Cout, x)
            _bn_forward([mod.residual[1]], [_full(Cout)], cr, st_r, N * To * V, train)
            res_mode, r, sr, hr = RES_AFFINE, r_raw, cr.scale, cr.shift
        elif mod.res_kind == 'identity':
            res_mode, r, sr, hr = RES_IDENTITY, x, None, None
        else:
            res_mode, r, sr, hr = RES_NONE, None, None, None
        out = _empty((N, Cout, To, V), x)
        ops.tcn_epilogue_fwd(u, cu.scale, cu.shift, res_mode, r, sr, hr, True, out)
        ctx.mod, ctx.train, ctx.coefs, ctx.pkd = mod, train, (ca, cu, cr), (pkg[1], pkt[1], pkr[1])
        ctx.save_for_backward(x, y, Af, agg, u, r_raw, out)
        ctx.a_dtype = A.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        mod, train = ctx.mod, ctx.train
        ca, cu, cr = ctx.coefs
        x, y, Af, agg, u, r_raw, out = ctx.saved_tensors
        N, Cout, To, V = u.shape
        Tg = agg.shape[2]
        g = g.contiguous().to(x.dtype)
        sb = _zeros((3, Cout), x, torch.float64)
        G = _empty(g.shape, x)
        ops.tcn_epilogue_bwd(g, out, True, u, r_raw, G, sb[0], sb[1], sb[2] if r_raw is not None else None)
        bu = _BnBwd(Cout, x)
        _bn_backward([mod.tcn[3]], [_full(Cout)], cu, bu, sb[0], sb[1], N * To * V, train)
        tc = mod.tcn[2]
        k, s, d, p = _conv_geom(tc)
        dyu = Opnd(G, u, a=bu.A, b=bu.B, c=bu.C)
        Wt = _w2(tc)
        dWt = _zeros(Wt.shape, x, torch.float32)
        dbt = _zeros((Cout,), x, torch.float32)
        ops.conv_wgrad(dyu, Opnd(agg, a=ca.scale, c=ca.shift, relu=True), dWt, dbt, k, s, d, p)
        DA = _empty(agg.shape, x)
        sa = _zeros((2, Cout), x, torch.float64)
        pkgd, pktd, pkrd = ctx.pkd
        ops.conv_dgrad(dyu, Wt, DA, k, s, d, p, mask=Opnd(agg, a=ca.scale, c=ca.shift), stats=(sa[0], sa[1]),
                       wpack=pktd)
        ba = _BnBwd(Cout, x)
        _bn_backward([mod.tcn[0]], [_full(Cout)], ca, ba, sa[0], sa[1], N * Tg * V, train)
        dy = _empty(y.shape, x)
        dA = _zeros(Af.shape, x, torch.float32)
        ops.graph_agg_bwd(Opnd(DA, agg, a=ba.A, b=ba.B, c=ba.C), y, Af, dy, dA)
        gk, gs, gd, gp = _conv_geom(mod.gcn.conv)
        Wg = _w2(mod.gcn.conv)
        dWg = _zeros(Wg.shape, x, torch.float32)
        dbg = _zeros((Wg.shape[0],), x, torch.float32)
        ops.conv_wgrad(dy, x, dWg, dbg, gk, gs, gd, gp)
        dx = _empty(x.shape, x)
        rgrads = []
        if mod.res_kind == 'conv':
            ops.conv_dgrad(dy, Wg, dx, gk, gs, gd, gp, wpack=pkgd)
            rc = mod.residual[0]
            rk, rs, rd, rp = _conv_geom(rc)
            br = _BnBwd(Cout, x)
            _bn_backward([mod.residual[1]], [_full(Cout)], cr, br, sb[0], sb[2], N * To * V, train)
            dyr = Opnd(G, r_raw, a=br.A, b=br.B, c=br.C)
            Wr = _w2(rc)
            dWr = _zeros(Wr.shape, x, torch.float32)
            dbr = _zeros((Cout,), x, torch.float32)
            ops.conv_wgrad(dyr, x, dWr, dbr, rk, rs, rd, rp)
            ops.conv_dgrad(dyr, Wr, dx, rk, rs, rd, rp, addend=dx, wpack=pkrd)
            rgrads = [dWr.view_as(rc.weight), dbr if rc.bias is not None else None, br.dgamma, br.dbeta]
        elif mod.res_kind == 'identity':
            ops.conv_dgrad(dy, Wg, dx, gk, gs, gd, gp, addend=G, wpack=pkgd)
        else:
            ops.conv_dgrad(dy, Wg, dx, gk, gs, gd, gp, wpack=pkgd)
        grads = [dWg.view_as(mod.gcn.conv.weight), dbg if mod.gcn.conv.bias is not None else None, ba.dgamma, ba.dbeta,
                 dWt.view_as(tc.weight), dbt if tc.bias is not None else None, bu.dgamma, bu.dbeta] + rgrads
        return (dx, dA.to(ctx.a_dtype), None) + tuple(grads)
