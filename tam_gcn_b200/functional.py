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
packing (cat/stack) of a few KB of parameters when no `params.ParamStore` owns them, and O(C) negations of
BatchNorm coefficients.

Parameter gradients: by default every backward returns them to autograd (fresh zero-initialised accumulators).
Inside `ParamStore.direct_grads()` (the engine's training step) the kernels accumulate straight into the store's
flat gradient buffer — the slice that mirrors the parameter (or the whole pack of adjacent parameters) — and
autograd gets None for them; that buffer is what the optimiser kernel and the NCCL all-reduce consume.
"""
import contextlib

import torch
import torch.nn as nn

from . import arena, ops
from .ops import Opnd, RES_NONE, RES_IDENTITY, RES_AFFINE


_sink = None          # params.ParamStore while its direct_grads() context is active
_pack_cache = None    # ops.PackCache of the engine driving the current step (persistent tensor-core weight tiles)


@contextlib.contextmanager
def pack_cache(cache):
    """Inside the context `_pack` serves weight tiles from `cache` (refreshed by one batched launch per step)."""
    global _pack_cache
    old = _pack_cache
    _pack_cache = cache
    try:
        if cache is not None:
            cache.repack_all()
        yield cache
    finally:
        if cache is not None:
            cache.end_step()
        _pack_cache = old


# ------------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------------
class _GradOut:
    """Destination of the parameter gradients of one backward call.

    `buf(t)` returns the accumulator for the gradient of `t` (a parameter or a pack of adjacent parameters): the
    mirroring slice of the active ParamStore's gradient buffer (already cleared for the step) or a fresh zeroed
    tensor.  `ret(g, t)` is what the autograd Function must return for it: None when the kernels wrote into the
    store, `g` otherwise."""

    def __init__(self, like):
        self.like = like
        self.sink = _sink
        self.direct = set()

    def buf(self, t, shape=None):
        if t is None:
            return None
        if self.sink is not None:
            g = self.sink.grad_view(t)
            if g is not None:
                self.direct.add(t.data_ptr())
                return g if shape is None else g.view(shape)
        return _zeros(tuple(t.shape) if shape is None else shape, self.like, torch.float32)

    def is_direct(self, t):
        return t is not None and t.data_ptr() in self.direct

    def ret(self, g, t):
        return None if (t is None or self.is_direct(t)) else g


def _require_cuda(x):
    if not x.is_cuda:
        raise RuntimeError('tam_gcn_b200 modules run on CUDA (sm_100a) only; there is no CPU path')


def _check_input(x):
    _require_cuda(x)
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


def _ctrgc_group_lists(convs, extra=None):
    """Parameter lists of the packs of K CTRGC modules (+ the unit_gcn down conv appended to conv3)."""
    w3 = [c.conv3.weight for c in convs]
    b3 = [c.conv3.bias for c in convs]
    if extra is not None:
        w3.append(extra.weight)
        b3.append(extra.bias)
    return {'W12': [c.conv1.weight for c in convs] + [c.conv2.weight for c in convs],
            'b12': [c.conv1.bias for c in convs] + [c.conv2.bias for c in convs],
            'W3': w3, 'b3': b3,
            'W4': [c.conv4.weight for c in convs], 'b4': [c.conv4.bias for c in convs]}


def pack_groups(mod):
    """[(owner module, key, [parameters])] — the parameter packs `mod` itself asks for (params.ParamStore lays them
    out adjacently so `_packed` never copies).  Sub-modules are visited separately by the caller."""
    kind = type(mod).__name__
    out = []
    if kind == 'unit_gcn' and hasattr(mod, 'convs'):
        convs = list(mod.convs)
        for key, ps in _ctrgc_group_lists(convs, mod.down[0] if getattr(mod, 'has_down', False) else None).items():
            out.append((convs[0], key, ps))
    elif kind == 'CTRGC' and hasattr(mod, 'conv4'):
        for key, ps in _ctrgc_group_lists([mod]).items():
            out.append((mod, key, ps))
    elif kind == 'MultiScale_TemporalConv' and hasattr(mod, 'num_dil'):
        heads = [mod.branches[j][0] for j in range(mod.num_dil + 1)]
        out.append((mod, 'Wh', [c.weight for c in heads]))
        out.append((mod, 'bh', [c.bias for c in heads]))
    return [(o, k, ps) for o, k, ps in out if all(p is not None for p in ps)]


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
    if _pack_cache is not None:
        return _pack_cache.get(W2d, W2d.shape[0], W2d.shape[1] // k, k, stride, like.shape[3])
    return ops.conv_pack_weights(W2d, W2d.shape[0], W2d.shape[1] // k, k, stride, like.shape[3])


def _bias(conv, like):
    if conv.bias is None:
        return torch.zeros(conv.weight.shape[0], device=like.device, dtype=torch.float32)
    return conv.bias


def _bn_train(bns):
    """Common `training` flag of the BatchNorms fused into one kernel (a model in train() with single BatchNorm
    layers put in eval() — the freeze-BN idiom — is honoured per layer; a mixed group cannot be fused)."""
    t = bns[0].training
    for b in bns:
        if b.training != t:
            raise NotImplementedError('BatchNorm layers fused in one kernel must share their train/eval mode')
    return bool(t)


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


def _bn_forward(bns, slices, coef, stats, count, train, coefs=None):
    """Finalize BatchNorms `bns` in ONE launch.  Either they are channel ranges `slices` of one concatenated
    coefficient table `coef` with one (sum, sumsq) pair `stats`, or (coefs=[...]) each has its own table and its own
    entry of the list `stats`."""
    m, e = _bn_group(bns)
    descs = []
    for i, (bn, sl) in enumerate(zip(bns, slices)):
        cf = coef if coefs is None else coefs[i]
        st = stats if (coefs is None or stats is None) else stats[i]
        d = dict(gamma=bn.weight, beta=bn.bias, rmean=bn.running_mean, rvar=bn.running_var,
                 nbt=bn.num_batches_tracked if train else None,
                 scale=cf.scale[sl], shift=cf.shift[sl], mean=cf.mean[sl], invstd=cf.invstd[sl])
        if train:
            d['sum'], d['sumsq'] = st[0][sl], st[1][sl]
        descs.append(d)
    ops.bn_finalize(descs, count, m, e, train)


class _BnBwd:
    """Backward affine dY = A*dYhat + B*Y + C of a group of BatchNorms, plus dgamma / dbeta.
    `abc`: optional (3, C) storage for the A / B / C rows (a slice of a wider coefficient table)."""

    def __init__(self, C, like, abc=None):
        if abc is None:
            self.t = torch.empty(5, C, device=like.device, dtype=torch.float32)
            self.A, self.B, self.C, self.dgamma, self.dbeta = self.t[0], self.t[1], self.t[2], self.t[3], self.t[4]
        else:
            self.t = torch.empty(2, C, device=like.device, dtype=torch.float32)
            self.A, self.B, self.C = abc
            self.dgamma, self.dbeta = self.t[0], self.t[1]


def _identity_coef_table(mod, Cw, KC, like):
    """(3, Cw) operand coefficients (a, b, c) whose first KC channels are the identity (1, 0, 0): the cotangent of the
    concatenated conv3|down output, where only the `down` channels carry a BatchNorm backward.  Cached on the module
    (per device); the down rows are rewritten by bn_bwd_coef in every backward, the identity rows never change."""
    cache = mod.__dict__.setdefault('_tamgcn_coef3', {})
    key = (like.device, Cw, KC)
    t = cache.get(key)
    if t is None:
        t = torch.zeros(3, Cw, device=like.device, dtype=torch.float32)
        t[0, :KC] = 1.0
        if not (like.is_cuda and torch.cuda.is_current_stream_capturing()):
            cache[key] = t
    return t


def _bn_backward(bns, slices, coef, bw, s1, s2, count, train, go=None, per=None):
    """Backward coefficients of a group of BatchNorms in ONE launch (channel ranges `slices` of shared tables, or
    per=[(coef, bw, s1, s2), ...] with one entry per BatchNorm); dgamma / dbeta go to the rows of `bw`, or — for
    parameters owned by the active ParamStore — straight into its gradient buffer (`go`: the call's _GradOut)."""
    descs = []
    for i, (bn, sl) in enumerate(zip(bns, slices)):
        cf, b_, a1, a2 = (coef, bw, s1, s2) if per is None else per[i]
        dg, db = b_.dgamma[sl], b_.dbeta[sl]
        if go is not None and go.sink is not None:
            gw, gb = go.sink.grad_view(bn.weight), go.sink.grad_view(bn.bias)
            if gw is not None and gb is not None:
                dg, db = gw, gb
                go.direct.add(bn.weight.data_ptr())
                go.direct.add(bn.bias.data_ptr())
        descs.append(dict(s1=a1[sl], s2=a2[sl], gamma=bn.weight, mean=cf.mean[sl], invstd=cf.invstd[sl],
                          A=b_.A[sl], B=b_.B[sl], Cc=b_.C[sl], dgamma=dg, dbeta=db))
    ops.bn_bwd_coef(descs, count, train)


def _bwd_opnd(P, Y, bw, train, sl=None):
    """dY = A*dYhat + B*Y + C as a lazy operand; in eval mode B = C = 0 and Y is not read at all."""
    A, B, Cc = (bw.A, bw.B, bw.C) if sl is None else (bw.A[sl], bw.B[sl], bw.C[sl])
    if train:
        return Opnd(P, Y, a=A, b=B, c=Cc)
    return Opnd(P, None, a=A)


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
    gl = _ctrgc_group_lists(convs, extra)
    Cw = K * Cout
    if extra is not None:
        _w2(extra)
        Cw += extra.weight.shape[0]
    zb = lambda ps, cs: [p if p is not None else torch.zeros(c.weight.shape[0], device=like.device) for p, c in zip(ps, cs)]
    c12 = [c.conv1 for c in convs] + [c.conv2 for c in convs]
    c3 = [c.conv3 for c in convs] + ([extra] if extra is not None else [])
    W12 = _packed(own, 'W12', gl['W12'], (2 * K * R, Cin))
    b12 = _packed(own, 'b12', zb(gl['b12'], c12), (2 * K * R,))
    W3 = _packed(own, 'W3', gl['W3'], (Cw, Cin))                                              # (K*Cout [+Cd], Cin)
    b3 = _packed(own, 'b3', zb(gl['b3'], c3), (Cw,))
    W4 = _packed(own, 'W4', gl['W4'], (K, Cout, R))
    b4 = _packed(own, 'b4', zb(gl['b4'], [c.conv4 for c in convs]), (K, Cout))
    return K, R, Cin, Cout, W12, b12, W3, b3, W4, b4


def _ctrgc_unpack_grads(convs, K, R, Cout, go, packs, dW12, db12, dW3, db3, dW4, db4):
    """Per-module gradient views, in the order conv1.w, conv1.b, conv2.w, conv2.b, conv3.w, conv3.b, conv4.w, conv4.b
    (None where the kernels accumulated straight into the ParamStore's gradient buffer)."""
    W12, b12, W3, b3, W4, b4 = packs
    out = []
    n12, nb12, n3, nb3, n4, nb4 = (go.is_direct(t) for t in (W12, b12, W3, b3, W4, b4))
    for i, c in enumerate(convs):
        out += [None if n12 else dW12[i * R:(i + 1) * R].view_as(c.conv1.weight),
                None if (nb12 or c.conv1.bias is None) else db12[i * R:(i + 1) * R],
                None if n12 else dW12[(K + i) * R:(K + i + 1) * R].view_as(c.conv2.weight),
                None if (nb12 or c.conv2.bias is None) else db12[(K + i) * R:(K + i + 1) * R],
                None if n3 else dW3[i * Cout:(i + 1) * Cout].view_as(c.conv3.weight),
                None if (nb3 or c.conv3.bias is None) else db3[i * Cout:(i + 1) * Cout],
                None if n4 else dW4[i].view_as(c.conv4.weight),
                None if (nb4 or c.conv4.bias is None) else db4[i]]
    return out


def ctrgc_params(c):
    return [c.conv1.weight, c.conv1.bias, c.conv2.weight, c.conv2.bias, c.conv3.weight, c.conv3.bias,
            c.conv4.weight, c.conv4.bias]


# ------------------------------------------------------------------------------------------------
# residual cotangent hand-over inside a TCN_GCN_unit
# ------------------------------------------------------------------------------------------------
# out = relu(tcn1(gcn1(x)) + residual(x)): x receives a cotangent from gcn1 AND one from the residual.  Left to autograd
# the two are summed by an ATen add over the whole activation.  Instead MsTcnFn.backward (which autograd always runs
# before the UnitGcnFn.backward of the same unit: tcn1 consumes gcn1's output) parks the residual cotangent on the gcn1
# module, keyed by the input it belongs to, and UnitGcnFn.backward adds it inside its last kernel.
def _park_residual_cotangent(gcn_mod, key, dr):
    """key = (data_ptr, version) of the residual source tensor == the unit_gcn input"""
    gcn_mod.__dict__.setdefault('_tamgcn_res_cot', {})[key] = dr


def _take_residual_cotangent(gcn_mod, x):
    d = gcn_mod.__dict__.get('_tamgcn_res_cot')
    if not d:
        return None
    dr = d.pop((x.data_ptr(), x._version), None)
    if dr is not None and (dr.shape != x.shape or dr.dtype != x.dtype):
        raise RuntimeError('residual cotangent does not match the unit_gcn input')
    return dr


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
        convs = list(mod.convs)
        has_down = mod.has_down
        tr_g, tr_o = _bn_train([mod.bn]), _bn_train([mod.offset_conv[1]])
        tr_d = _bn_train([mod.down[1]]) if has_down else False
        K, R, Cin_w, Cout, W12, b12, W3, b3, W4, b4 = _ctrgc_pack(convs, x, mod.down[0] if has_down else None)
        if Cin_w != Cin:
            raise ValueError('unit_gcn: input has %d channels, module expects %d' % (Cin, Cin_w))
        PA = (mod.PA if mod.adaptive else mod.A).to(torch.float32).contiguous()
        alpha = mod.alpha
        count = N * T * V
        KC = K * Cout

        m = _empty((N, Cin, 1, V), x, torch.float32)
        x12 = _empty((N, 2 * K * R, 1, V), x, torch.float32)
        Cw = KC + (Cout if has_down else 0)
        xw = _empty((N, Cw, T, V), x)
        stats = _zeros((6, Cout), x, torch.float64) if (tr_g or tr_o or tr_d) else None
        # rows: down(sum,sq), bn(sum,sq), offset(sum,sq)
        pk3 = _pack(W3, 1, x)
        # x1 / x2: 1x1 convs on the T-mean of x — independent of conv3, on a parallel branch inside an engine step
        with ops.branch(1):
            ops.mean_t(x, m)
            ops.conv_fwd(m, W12, b12, x12)
        # conv3 of the K subsets (+ down conv) in one pass over x
        ops.conv_fwd(x, W3, b3, xw, stats=(stats[0], stats[1]) if tr_d else None, stat_c0=KC, wpack=pk3[0])
        ops.branch_join()
        # fused topology refinement + aggregation (+ BN statistics of y0)
        y0 = _empty((N, Cout, T, V), x)
        ops.ctrgc_fwd(xw[:, :KC], x12[:, :K * R], x12[:, K * R:], W4, b4, PA, alpha, y0,
                      stats=(stats[2], stats[3]) if tr_g else None)
        cg = _BnCoef(Cout, x)
        cd = _BnCoef(Cout, x) if has_down else None
        if has_down and tr_d == tr_g:
            _bn_forward([mod.down[1], mod.bn], [slice(0, Cout), slice(0, Cout)], None,
                        [(stats[0], stats[1]), (stats[2], stats[3])] if tr_g else None, count, tr_g, coefs=[cd, cg])
        else:
            if has_down:
                _bn_forward([mod.down[1]], [_full(Cout)], cd, (stats[0], stats[1]) if tr_d else None, count, tr_d)
            _bn_forward([mod.bn], [_full(Cout)], cg, (stats[2], stats[3]) if tr_g else None, count, tr_g)
        # offset branch: z = W_o (res - y) + b_o, with res - y formed while loading; its coefficients
        # (-scale_y, shift_res - shift_y) come from one tiny kernel
        dc = torch.empty(2, Cout, device=x.device, dtype=torch.float32)
        nsg, dsh = dc[0], dc[1]
        ops.coef_diff(cg.scale, cd.shift if has_down else None, cg.shift, nsg, dsh)
        if has_down:
            diff = Opnd(xw[:, KC:], y0, a=cd.scale, b=nsg, c=dsh)
            res_mode, r, sr, hr = RES_AFFINE, xw[:, KC:], cd.scale, cd.shift
        elif mod.residual_identity:
            diff = Opnd(x, y0, a=None, b=nsg, c=dsh)
            res_mode, r, sr, hr = RES_IDENTITY, x, None, None
        else:
            diff = Opnd(y0, None, a=nsg, c=dsh)
            res_mode, r, sr, hr = RES_NONE, None, None, None
        oc = mod.offset_conv[0]
        Wo, bo = _w2(oc), _bias(oc, x)
        z = _empty((N, Cout, T, V), x)
        pko = _pack(Wo, 1, x)
        ops.conv_fwd(diff, Wo, bo, z, stats=(stats[4], stats[5]) if tr_o else None, wpack=pko[0])
        co = _BnCoef(Cout, x)
        _bn_forward([mod.offset_conv[1]], [_full(Cout)], co, (stats[4], stats[5]) if tr_o else None, count, tr_o)
        out = _empty((N, Cout, T, V), x)
        ops.gcn_epilogue_fwd(y0, cg.scale, cg.shift, z, co.scale, co.shift, res_mode, r, sr, hr, out)

        ctx.mod, ctx.train, ctx.dims = mod, (tr_g, tr_o, tr_d), (N, Cin, Cout, T, V, K, R)
        ctx.res_mode = res_mode
        ctx.diff = diff
        ctx.coefs = (cg, cd, co)
        ctx.packed = (W12, b12, W3, b3, W4, b4, PA, Wo, pk3[1], pko[1])
        ctx.save_for_backward(x, m, x12, xw, y0, z, out)
        return out

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        tr_g, tr_o, tr_d = ctx.train
        N, Cin, Cout, T, V, K, R = ctx.dims
        x, m, x12, xw, y0, z, out = ctx.saved_tensors
        cg, cd, co = ctx.coefs
        W12, b12, W3, b3, W4, b4, PA, Wo, pk3d, pkod = ctx.packed
        has_down = cd is not None
        res_mode = ctx.res_mode
        KC, count = K * Cout, N * T * V
        g = g.contiguous()
        if g.dtype != x.dtype:
            g = g.to(x.dtype)
        go = _GradOut(x)
        sb = _zeros((6, Cout), x, torch.float64)      # rows: offset(s1,s2), bn(s1,s2), down(s1,s2)
        # cotangent that TCN_GCN_unit's residual sends to the same input x (handed over by MsTcnFn.backward, which always
        # runs first): folded into this backward's last kernel instead of an autograd add pass
        xdr = _take_residual_cotangent(mod, x)

        # tail: ReLU mask, tanh', BN_o backward sums
        G = _empty(g.shape, x)
        DZ = _empty(g.shape, x)
        ops.gcn_epilogue_bwd(g, out, z, co.scale, co.shift, G, DZ, sb[0], sb[1])
        bo_ = _BnBwd(Cout, x)
        obn = mod.offset_conv[1]
        _bn_backward([obn], [_full(Cout)], co, bo_, sb[0], sb[1], count, tr_o, go)
        dz = _bwd_opnd(DZ, z, bo_, tr_o)
        # offset conv backward
        oc = mod.offset_conv[0]
        dWo = go.buf(oc.weight, Wo.shape)
        dbo = go.buf(oc.bias) if oc.bias is not None else _zeros((Cout,), x, torch.float32)
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
            ops.gcn_mid_bwd(G, DD, dres, y0, None, sb[2], sb[3], None, None, extra=xdr)
            xdr = None
        else:
            dres = None
            ops.gcn_mid_bwd(G, DD, None, y0, None, sb[2], sb[3], None, None)
        bg = _BnBwd(Cout, x)
        bd = None
        if has_down:
            tab = _identity_coef_table(mod, Cw, KC, x)
            bd = _BnBwd(Cout, x, abc=(tab[0, KC:], tab[1, KC:], tab[2, KC:]))
            if tr_d == tr_g:
                _bn_backward([mod.bn, mod.down[1]], [_full(Cout), _full(Cout)], None, None, None, None, count, tr_g, go,
                             per=[(cg, bg, sb[2], sb[3]), (cd, bd, sb[4], sb[5])])
            else:
                _bn_backward([mod.bn], [_full(Cout)], cg, bg, sb[2], sb[3], count, tr_g, go)
                _bn_backward([mod.down[1]], [_full(Cout)], cd, bd, sb[4], sb[5], count, tr_d, go)
        else:
            _bn_backward([mod.bn], [_full(Cout)], cg, bg, sb[2], sb[3], count, tr_g, go)
        # fused CTRGC backward
        dy = _bwd_opnd(G, y0, bg, tr_g)
        dx12 = _zeros(x12.shape, x, torch.float32)
        dW4, db4 = go.buf(W4), go.buf(b4)
        dPA = go.buf(mod.PA, PA.shape) if mod.adaptive else _zeros(PA.shape, x, torch.float32)
        dalpha = go.buf(mod.alpha)
        ops.ctrgc_bwd(dy, xw[:, :KC], x12[:, :K * R], x12[:, K * R:], W4, b4, PA, mod.alpha, dxw[:, :KC],
                      dx12[:, :K * R], dx12[:, K * R:], dW4, db4, dPA, dalpha)
        # conv1/conv2 backward on the T-mean
        dW12, db12 = go.buf(W12), go.buf(b12)
        ops.conv_wgrad(dx12, m, dW12, db12)
        dm = _empty(m.shape, x, torch.float32)
        ops.conv_dgrad(dx12, W12, dm)
        # conv3 (+down) backward: one wgrad and one dgrad over the concatenated channels
        if has_down:
            dxw_op = Opnd(dxw, xw, a=tab[0], b=tab[1], c=tab[2])
        else:
            dxw_op = Opnd(dxw)
        dW3, db3 = go.buf(W3), go.buf(b3)
        ops.conv_wgrad(dxw_op, x, dW3, db3)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dxw_op, W3, dx, addend=dres if dres is not None else xdr, bcast=dm, bcast_scale=1.0 / T, wpack=pk3d)

        grads = _ctrgc_unpack_grads(list(mod.convs), K, R, Cout, go, (W12, b12, W3, b3, W4, b4), dW12, db12, dW3, db3,
                                    dW4, db4)
        if has_down:
            d0, d1 = mod.down[0], mod.down[1]
            n3, nb3 = go.is_direct(W3), go.is_direct(b3)
            grads += [None if n3 else dW3[KC:].view_as(d0.weight), None if (nb3 or d0.bias is None) else db3[KC:],
                      go.ret(bd.dgamma, d1.weight), go.ret(bd.dbeta, d1.bias)]
        grads += [go.ret(dWo.view_as(oc.weight), oc.weight), go.ret(dbo, oc.bias), go.ret(bo_.dgamma, obn.weight),
                  go.ret(bo_.dbeta, obn.bias), go.ret(bg.dgamma, mod.bn.weight), go.ret(bg.dbeta, mod.bn.bias),
                  go.ret(dalpha, mod.alpha)]
        if mod.adaptive:
            grads.append(go.ret(dPA, mod.PA))
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
        ctx.packed = (W12, b12, W3, b3, W4, b4, PA, al, pk3[1])
        ctx.save_for_backward(x, m, x12, x3)
        ctx.a_shape, ctx.alpha_shape = A.shape, alpha.shape
        return y

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        N, Cin, Cout, T, V, R = ctx.dims
        x, m, x12, x3 = ctx.saved_tensors
        W12, b12, W3, b3, W4, b4, PA, al, pk3d = ctx.packed
        g = g.contiguous().to(x.dtype)
        go = _GradOut(x)
        dx3 = _empty(x3.shape, x)
        dx12 = _zeros(x12.shape, x, torch.float32)
        dW4, db4 = go.buf(W4), go.buf(b4)
        dPA = _zeros(PA.shape, x, torch.float32)
        dalpha = _zeros((1,), x, torch.float32)
        ops.ctrgc_bwd(Opnd(g), x3, x12[:, :R], x12[:, R:], W4, b4, PA, al, dx3, dx12[:, :R], dx12[:, R:], dW4, db4,
                      dPA, dalpha)
        dW12, db12 = go.buf(W12), go.buf(b12)
        ops.conv_wgrad(dx12, m, dW12, db12)
        dm = _empty(m.shape, x, torch.float32)
        ops.conv_dgrad(dx12, W12, dm)
        dW3, db3 = go.buf(W3), go.buf(b3)
        ops.conv_wgrad(dx3, x, dW3, db3)
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dx3, W3, dx, bcast=dm, bcast_scale=1.0 / T, wpack=pk3d)
        grads = _ctrgc_unpack_grads([mod], 1, R, Cout, go, (W12, b12, W3, b3, W4, b4), dW12, db12, dW3, db3, dW4, db4)
        return (dx, dPA.reshape(ctx.a_shape), dalpha.reshape(ctx.alpha_shape), None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# conv (k x 1) + BN   (unit_tcn, TemporalConv)
# ------------------------------------------------------------------------------------------------
def _conv_geom(conv):
    k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
    if conv.kernel_size[1] != 1 or conv.stride[1] != 1 or conv.padding[1] != 0 or conv.groups != 1:
        raise NotImplementedError('only (k x 1) ungrouped temporal convolutions are supported')
    return k, s, d, p


def _conv_wgrad_to(go, conv, dy, xop, like, geom=(1, 1, 1, 0)):
    """Weight / bias gradient of `conv` into the call's gradient destination; returns (dW, db) as autograd wants them."""
    W = _w2(conv)
    dW = go.buf(conv.weight, W.shape)
    db = go.buf(conv.bias) if conv.bias is not None else _zeros((W.shape[0],), like, torch.float32)
    ops.conv_wgrad(dy, xop, dW, db, *geom)
    return go.ret(dW.view_as(conv.weight), conv.weight), go.ret(db, conv.bias)


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
        go = _GradOut(x)
        sb = _zeros((2, Cout), x, torch.float64)
        ops.tcn_epilogue_bwd(g, None, False, raw, None, None, sb[0], sb[1], None)
        bw = _BnBwd(Cout, x)
        _bn_backward([bn], [_full(Cout)], ctx.cf, bw, sb[0], sb[1], N * To * V, train, go)
        dy = _bwd_opnd(g, raw, bw, train)
        gW, gb = _conv_wgrad_to(go, conv, dy, x, x, (k, s, d, p))
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dy, _w2(conv), dx, k, s, d, p, wpack=ctx.pkd)
        return dx, None, None, gW, gb, go.ret(bw.dgamma, bn.weight), go.ret(bw.dbeta, bn.bias)


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
        # `res_sink` (set by TCN_GCN_unit through mod._tamgcn_res_sink): the unit_gcn whose input IS r_in; its backward
        # takes over the residual cotangent (see _park_residual_cotangent)
        ctx.res_sink = mod.__dict__.get('_tamgcn_res_sink') if r_in is not None else None
        ctx.r_in_key = (r_in.data_ptr(), r_in._version) if r_in is not None else None
        x = _check_input(x)
        N, Cin, T, V = x.shape
        nd, Cb, s = mod.num_dil, mod.branch_channels, mod.stride
        nb = nd + 2
        Cout, Ch = nb * Cb, (nd + 1) * Cb
        To = _conv_out_len(T, 1, s, 1, 0)
        r_src = x if r_in is None else _check_input(r_in)
        head_bns = [mod.branches[j][1] for j in range(nd + 1)]
        final_bns = [mod.branches[j][3].bn for j in range(nd)] + [mod.branches[nd][4], mod.branches[nd + 1][1]]
        tr_h, tr_u = _bn_train(head_bns), _bn_train(final_bns)

        # all 1x1 branch heads that keep T in one pass over x
        heads = [mod.branches[j][0] for j in range(nd + 1)]
        for c in heads:
            _w2(c)
        Wh = _packed(mod, 'Wh', [c.weight for c in heads], ((nd + 1) * Cb, Cin))
        bh = _packed(mod, 'bh', [_bias(c, x) for c in heads], ((nd + 1) * Cb,))
        h = _empty((N, Ch, T, V), x)
        st_h = _zeros((2, Ch), x, torch.float64) if tr_h else None
        pkh = _pack(Wh, 1, x)
        u = _empty((N, Cout, To, V), x)
        st_u = _zeros((2, Cout), x, torch.float64) if tr_u else None
        c3 = mod.branches[nd + 1][0]
        W3, b3 = _w2(c3), _bias(c3, x)
        pk3 = _pack(W3, 1, x)
        # strided 1x1 branch straight into its slice of u: independent of the heads (parallel branch in an engine step)
        with ops.branch(1):
            ops.conv_fwd(x, W3, b3, u[:, Ch:], 1, s, 1, 0, stats=(st_u[0][Ch:], st_u[1][Ch:]) if tr_u else None,
                         wpack=pk3[0])
        ops.conv_fwd(x, Wh, bh, h, stats=st_h, wpack=pkh[0])
        ch = _BnCoef(Ch, x)
        sl = [slice(j * Cb, (j + 1) * Cb) for j in range(nb)]
        _bn_forward(head_bns, sl[:nd + 1], ch, st_h, N * T * V, tr_h)
        geoms, keep = [], []
        for j in range(nd):
            tc = mod.branches[j][3].conv
            k, cs, d, p = _conv_geom(tc)
            if cs != s or _conv_out_len(T, k, cs, d, p) != To:
                raise ValueError('MultiScale_TemporalConv: branch %d output length differs' % j)
            pkt = _pack(_w2(tc), k, x, cs)
            geoms.append((k, cs, d, p, pkt[1]))
            bj = _bias(tc, x)
            keep.append((pkt, bj))
            with ops.branch(j):                      # the temporal branches write disjoint slices of u: run side by side
                ops.conv_fwd(Opnd(h[:, sl[j]], a=ch.scale[sl[j]], c=ch.shift[sl[j]], relu=True), _w2(tc), bj,
                             u[:, sl[j]], k, cs, d, p, stats=(st_u[0][sl[j]], st_u[1][sl[j]]) if tr_u else None,
                             wpack=pkt[0])
        if _conv_out_len(T, 3, s, 1, 1) != To:
            raise ValueError('MultiScale_TemporalConv: max-pool branch output length differs')
        with ops.branch(nd):
            ops.maxpool_fwd(Opnd(h[:, sl[nd]], a=ch.scale[sl[nd]], c=ch.shift[sl[nd]], relu=True), u[:, sl[nd]], s,
                            stats=(st_u[0][sl[nd]], st_u[1][sl[nd]]) if tr_u else None)
        ops.branch_join()
        cu = _BnCoef(Cout, x)
        # residual (its BatchNorm is finalized in the same launch as the branch BatchNorms)
        cr = r_raw = None
        tr_r = False
        if res_kind == 'conv':
            rk, rs, rd, rp = _conv_geom(res_mod.conv)
            if _conv_out_len(r_src.shape[2], rk, rs, rd, rp) != To or res_mod.conv.weight.shape[0] != Cout:
                raise ValueError('residual branch shape mismatch')
            tr_r = res_mod.bn.training
            r_raw = _empty((N, Cout, To, V), x)
            st_r = _zeros((2, Cout), x, torch.float64) if tr_r else None
            pkr = _pack(_w2(res_mod.conv), rk, x, rs)
            ops.conv_fwd(r_src, _w2(res_mod.conv), _bias(res_mod.conv, x), r_raw, rk, rs, rd, rp, stats=st_r,
                         wpack=pkr[0])
            cr = _BnCoef(Cout, x)
            if tr_r == tr_u:
                _bn_forward(final_bns + [res_mod.bn], sl + [_full(Cout)], None, [st_u] * nb + [st_r], N * To * V, tr_u,
                            coefs=[cu] * nb + [cr])
            else:
                _bn_forward(final_bns, sl, cu, st_u, N * To * V, tr_u)
                _bn_forward([res_mod.bn], [_full(Cout)], cr, st_r, N * To * V, tr_r)
            res_mode, r, sr, hr = RES_AFFINE, r_raw, cr.scale, cr.shift
        else:
            _bn_forward(final_bns, sl, cu, st_u, N * To * V, tr_u)
            if res_kind == 'identity':
                if r_src.shape != (N, Cout, To, V):
                    raise ValueError('identity residual shape mismatch')
                res_mode, r, sr, hr = RES_IDENTITY, r_src, None, None
            else:
                res_mode, r, sr, hr = RES_NONE, None, None, None
        out = _empty((N, Cout, To, V), x)
        ops.tcn_epilogue_fwd(u, cu.scale, cu.shift, res_mode, r, sr, hr, relu, out)

        ctx.mod, ctx.res_mod, ctx.res_kind, ctx.relu, ctx.train = mod, res_mod, res_kind, relu, (tr_h, tr_u, tr_r)
        ctx.geoms, ctx.coefs = geoms, (ch, cu, cr)
        ctx.packed = (Wh, bh, W3, pkh[1], pk3[1], pkr[1] if res_kind == 'conv' else None)
        ctx.r_is_x = r_in is None
        ctx.save_for_backward(x, r_src if res_kind == 'conv' else None, h, u, r_raw, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        mod, res_mod, res_kind, relu = ctx.mod, ctx.res_mod, ctx.res_kind, ctx.relu
        tr_h, tr_u, tr_r = ctx.train
        x, r_src, h, u, r_raw, out = ctx.saved_tensors
        ch, cu, cr = ctx.coefs
        Wh, bh, W3, pkhd, pk3d, pkrd = ctx.packed
        N, Cin, T, V = x.shape
        nd, Cb, s = mod.num_dil, mod.branch_channels, mod.stride
        nb = nd + 2
        Cout, Ch = nb * Cb, (nd + 1) * Cb
        To = u.shape[2]
        sl = [slice(j * Cb, (j + 1) * Cb) for j in range(nb)]
        g = g.contiguous().to(x.dtype)
        go = _GradOut(x)

        sb = _zeros((3, Cout), x, torch.float64)
        G = _empty(g.shape, x) if relu else None
        ops.tcn_epilogue_bwd(g, out, relu, u, r_raw, G, sb[0], sb[1], sb[2] if r_raw is not None else None)
        Gt = G if relu else g
        head_bns = [mod.branches[j][1] for j in range(nd + 1)]
        final_bns = [mod.branches[j][3].bn for j in range(nd)] + [mod.branches[nd][4], mod.branches[nd + 1][1]]
        bu = _BnBwd(Cout, x)
        br = None
        if res_kind == 'conv' and tr_r == tr_u:
            br = _BnBwd(Cout, x)
            _bn_backward(final_bns + [res_mod.bn], sl + [_full(Cout)], None, None, None, None, N * To * V, tr_u, go,
                         per=[(cu, bu, sb[0], sb[1])] * nb + [(cr, br, sb[0], sb[2])])
        else:
            _bn_backward(final_bns, sl, cu, bu, sb[0], sb[1], N * To * V, tr_u, go)
            if res_kind == 'conv':
                br = _BnBwd(Cout, x)
                _bn_backward([res_mod.bn], [_full(Cout)], cr, br, sb[0], sb[2], N * To * V, tr_r, go)

        def dy_op(c0, c1):
            return _bwd_opnd(Gt[:, c0:c1], u[:, c0:c1], bu, tr_u, slice(c0, c1))

        DH = _empty(h.shape, x)
        sh = _zeros((2, Ch), x, torch.float64)
        tgrads = []
        for j in range(nd):
            tc = mod.branches[j][3].conv
            k, cs, d, p, pktd = ctx.geoms[j]
            dyj = dy_op(sl[j].start, sl[j].stop)
            hj = h[:, sl[j]]
            tgrads.append(_conv_wgrad_to(go, tc, dyj, Opnd(hj, a=ch.scale[sl[j]], c=ch.shift[sl[j]], relu=True), x,
                                         (k, cs, d, p)))
            with ops.branch(j):                      # disjoint slices of DH / of the statistics: side by side
                ops.conv_dgrad(dyj, _w2(tc), DH[:, sl[j]], k, cs, d, p, mask=Opnd(hj, a=ch.scale[sl[j]], c=ch.shift[sl[j]]),
                               stats=(sh[0][sl[j]], sh[1][sl[j]]), wpack=pktd)
        with ops.branch(nd):
            ops.maxpool_bwd(dy_op(sl[nd].start, sl[nd].stop),
                            Opnd(h[:, sl[nd]], a=ch.scale[sl[nd]], c=ch.shift[sl[nd]], relu=True), DH[:, sl[nd]], s,
                            stats=(sh[0][sl[nd]], sh[1][sl[nd]]))
        ops.branch_join()
        bh_ = _BnBwd(Ch, x)
        _bn_backward(head_bns, sl[:nd + 1], ch, bh_, sh[0], sh[1], N * T * V, tr_h, go)
        dh = _bwd_opnd(DH, h, bh_, tr_h)
        dWh, dbh = go.buf(Wh), go.buf(bh)
        ops.conv_wgrad(dh, x, dWh, dbh)
        dx = _empty(x.shape, x)
        # identity residual of a stand-alone module: its cotangent is added by the first data-gradient kernel
        ops.conv_dgrad(dh, Wh, dx, addend=Gt if (res_kind == 'identity' and ctx.r_is_x) else None, wpack=pkhd)
        dy3 = dy_op(Ch, Cout)
        c3 = mod.branches[nd + 1][0]
        g3 = _conv_wgrad_to(go, c3, dy3, x, x, (1, s, 1, 0))
        ops.conv_dgrad(dy3, W3, dx, 1, s, 1, 0, addend=dx, wpack=pk3d)

        dr = None
        rgrads = []
        if res_kind == 'conv':
            rk, rs, rd, rp = _conv_geom(res_mod.conv)
            dyr = _bwd_opnd(Gt, r_raw, br, tr_r)
            gr = _conv_wgrad_to(go, res_mod.conv, dyr, r_src, x, (rk, rs, rd, rp))
            Wr = _w2(res_mod.conv)
            if ctx.r_is_x:
                ops.conv_dgrad(dyr, Wr, dx, rk, rs, rd, rp, addend=dx, wpack=pkrd)
            else:
                dr = _empty(r_src.shape, x)
                ops.conv_dgrad(dyr, Wr, dr, rk, rs, rd, rp, wpack=pkrd)
            rgrads = [gr[0], gr[1], go.ret(br.dgamma, res_mod.bn.weight), go.ret(br.dbeta, res_mod.bn.bias)]
        elif res_kind == 'identity' and not ctx.r_is_x:
            dr = Gt

        nh, nbh = go.is_direct(Wh), go.is_direct(bh)
        grads = []
        for j in range(nd):
            br_ = mod.branches[j]
            grads += [None if nh else dWh[sl[j]].view_as(br_[0].weight),
                      None if (nbh or br_[0].bias is None) else dbh[sl[j]],
                      go.ret(bh_.dgamma[sl[j]], br_[1].weight), go.ret(bh_.dbeta[sl[j]], br_[1].bias),
                      tgrads[j][0], tgrads[j][1],
                      go.ret(bu.dgamma[sl[j]], br_[3].bn.weight), go.ret(bu.dbeta[sl[j]], br_[3].bn.bias)]
        br_ = mod.branches[nd]
        grads += [None if nh else dWh[sl[nd]].view_as(br_[0].weight),
                  None if (nbh or br_[0].bias is None) else dbh[sl[nd]],
                  go.ret(bh_.dgamma[sl[nd]], br_[1].weight), go.ret(bh_.dbeta[sl[nd]], br_[1].bias),
                  go.ret(bu.dgamma[sl[nd]], br_[4].weight), go.ret(bu.dbeta[sl[nd]], br_[4].bias)]
        br_ = mod.branches[nd + 1]
        grads += [g3[0], g3[1], go.ret(bu.dgamma[sl[nd + 1]], br_[1].weight), go.ret(bu.dbeta[sl[nd + 1]], br_[1].bias)]
        grads += rgrads
        if dr is not None and ctx.res_sink is not None and ctx.needs_input_grad[1]:
            # hand the residual cotangent to the unit_gcn backward that follows (keyed by the residual source tensor)
            _park_residual_cotangent(ctx.res_sink, ctx.r_in_key, dr.contiguous())
            dr = None
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
        go = _GradOut(x)
        dy = _empty(y.shape, x)
        dA = _zeros(Af.shape, x, torch.float32)
        ops.graph_agg_bwd(Opnd(g), y, Af, dy, dA)
        gW, gb = _conv_wgrad_to(go, mod.conv, dy, x, x, (k, s, d, p))
        dx = _empty(x.shape, x)
        ops.conv_dgrad(dy, _w2(mod.conv), dx, k, s, d, p, wpack=ctx.pkd)
        return dx, dA.to(ctx.a_dtype), None, gW, gb


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
        tr_a, tr_u = mod.tcn[0].training, mod.tcn[3].training
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
        st_a = _zeros((2, Cout), x, torch.float64) if tr_a else None
        ops.graph_agg_fwd(y, Af, agg, stats=st_a)
        ca = _BnCoef(Cout, x)
        _bn_forward([mod.tcn[0]], [_full(Cout)], ca, st_a, N * Tg * V, tr_a)
        tc = mod.tcn[2]
        k, s, d, p = _conv_geom(tc)
        To = _conv_out_len(Tg, k, s, d, p)
        u = _empty((N, Cout, To, V), x)
        st_u = _zeros((2, Cout), x, torch.float64) if tr_u else None
        pkt = _pack(_w2(tc), k, x, s)
        ops.conv_fwd(Opnd(agg, a=ca.scale, c=ca.shift, relu=True), _w2(tc), _bias(tc, x), u, k, s, d, p, stats=st_u,
                     wpack=pkt[0])
        cu = _BnCoef(Cout, x)
        _bn_forward([mod.tcn[3]], [_full(Cout)], cu, st_u, N * To * V, tr_u)
        cr = r_raw = None
        tr_r = False
        pkr = (None, None)
        if mod.res_kind == 'conv':
            rc = mod.residual[0]
            tr_r = mod.residual[1].training
            rk, rs, rd, rp = _conv_geom(rc)
            r_raw = _empty((N, Cout, To, V), x)
            st_r = _zeros((2, Cout), x, torch.float64) if tr_r else None
            pkr = _pack(_w2(rc), rk, x, rs)
            ops.conv_fwd(x, _w2(rc), _bias(rc, x), r_raw, rk, rs, rd, rp, stats=st_r, wpack=pkr[0])
            cr = _BnCoef(Cout, x)
            _bn_forward([mod.residual[1]], [_full(Cout)], cr, st_r, N * To * V, tr_r)
            res_mode, r, sr, hr = RES_AFFINE, r_raw, cr.scale, cr.shift
        elif mod.res_kind == 'identity':
            res_mode, r, sr, hr = RES_IDENTITY, x, None, None
        else:
            res_mode, r, sr, hr = RES_NONE, None, None, None
        out = _empty((N, Cout, To, V), x)
        ops.tcn_epilogue_fwd(u, cu.scale, cu.shift, res_mode, r, sr, hr, True, out)
        ctx.mod, ctx.train, ctx.coefs, ctx.pkd = mod, (tr_a, tr_u, tr_r), (ca, cu, cr), (pkg[1], pkt[1], pkr[1])
        ctx.save_for_backward(x, y, Af, agg, u, r_raw, out)
        ctx.a_dtype = A.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        mod = ctx.mod
        tr_a, tr_u, tr_r = ctx.train
        ca, cu, cr = ctx.coefs
        x, y, Af, agg, u, r_raw, out = ctx.saved_tensors
        N, Cout, To, V = u.shape
        Tg = agg.shape[2]
        g = g.contiguous().to(x.dtype)
        go = _GradOut(x)
        sb = _zeros((3, Cout), x, torch.float64)
        G = _empty(g.shape, x)
        ops.tcn_epilogue_bwd(g, out, True, u, r_raw, G, sb[0], sb[1], sb[2] if r_raw is not None else None)
        bu = _BnBwd(Cout, x)
        _bn_backward([mod.tcn[3]], [_full(Cout)], cu, bu, sb[0], sb[1], N * To * V, tr_u, go)
        tc = mod.tcn[2]
        k, s, d, p = _conv_geom(tc)
        dyu = _bwd_opnd(G, u, bu, tr_u)
        gWt, gbt = _conv_wgrad_to(go, tc, dyu, Opnd(agg, a=ca.scale, c=ca.shift, relu=True), x, (k, s, d, p))
        DA = _empty(agg.shape, x)
        sa = _zeros((2, Cout), x, torch.float64)
        pkgd, pktd, pkrd = ctx.pkd
        ops.conv_dgrad(dyu, _w2(tc), DA, k, s, d, p, mask=Opnd(agg, a=ca.scale, c=ca.shift), stats=(sa[0], sa[1]),
                       wpack=pktd)
        ba = _BnBwd(Cout, x)
        _bn_backward([mod.tcn[0]], [_full(Cout)], ca, ba, sa[0], sa[1], N * Tg * V, tr_a, go)
        dy = _empty(y.shape, x)
        dA = _zeros(Af.shape, x, torch.float32)
        ops.graph_agg_bwd(_bwd_opnd(DA, agg, ba, tr_a), y, Af, dy, dA)
        gk, gs, gd, gp = _conv_geom(mod.gcn.conv)
        Wg = _w2(mod.gcn.conv)
        gWg, gbg = _conv_wgrad_to(go, mod.gcn.conv, dy, x, x, (gk, gs, gd, gp))
        dx = _empty(x.shape, x)
        rgrads = []
        if mod.res_kind == 'conv':
            ops.conv_dgrad(dy, Wg, dx, gk, gs, gd, gp, wpack=pkgd)
            rc, rb = mod.residual[0], mod.residual[1]
            rk, rs, rd, rp = _conv_geom(rc)
            br = _BnBwd(Cout, x)
            _bn_backward([rb], [_full(Cout)], cr, br, sb[0], sb[2], N * To * V, tr_r, go)
            dyr = _bwd_opnd(G, r_raw, br, tr_r)
            gWr, gbr = _conv_wgrad_to(go, rc, dyr, x, x, (rk, rs, rd, rp))
            ops.conv_dgrad(dyr, _w2(rc), dx, rk, rs, rd, rp, addend=dx, wpack=pkrd)
            rgrads = [gWr, gbr, go.ret(br.dgamma, rb.weight), go.ret(br.dbeta, rb.bias)]
        elif mod.res_kind == 'identity':
            ops.conv_dgrad(dy, Wg, dx, gk, gs, gd, gp, addend=G, wpack=pkgd)
        else:
            ops.conv_dgrad(dy, Wg, dx, gk, gs, gd, gp, wpack=pkgd)
        grads = [gWg, gbg, go.ret(ba.dgamma, mod.tcn[0].weight), go.ret(ba.dbeta, mod.tcn[0].bias), gWt, gbt,
                 go.ret(bu.dgamma, mod.tcn[3].weight), go.ret(bu.dbeta, mod.tcn[3].bias)] + rgrads
        return (dx, dA.to(ctx.a_dtype), None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# network ends: data_bn prologue, pooled classifier head, cross-entropy   (csrc/head.cu)
# ------------------------------------------------------------------------------------------------
class DataBnFn(torch.autograd.Function):
    """Model.forward prologue (models/ctrgcn.py:324-332, models/stgcn.py:168-181): (N,C,T,V,M) or (N,T,V*C) fp32 input
    -> BatchNorm1d over the (m, v, c) channels (fold_m: over (v, c) with the persons in the batch, ST-GCN) ->
    (N*M, C, T, V) in the activation dtype.  One kernel each way; the input is read in place through its strides."""

    @staticmethod
    def forward(ctx, x, bn, num_point, fold_m, act_dtype, *params):
        _require_cuda(x)
        if x.dtype != torch.float32:
            x = x.float()
        if x.dim() not in (3, 5):
            raise ValueError('expected (N, C, T, V, M) or (N, T, V*C) input, got shape %s' % (tuple(x.shape),))
        if not bn.affine or not bn.track_running_stats or bn.momentum is None:
            raise NotImplementedError('data_bn without affine / running statistics / float momentum is not supported')
        if x.dim() == 5:
            N, C, T, V, M = x.shape
        else:
            N, T, VC = x.shape
            V, M = num_point, 1
            C = VC // V
        nch = (1 if fold_m else M) * V * C
        if bn.weight.numel() != nch:
            raise ValueError('data_bn has %d channels, input needs %d' % (bn.weight.numel(), nch))
        train = bn.training
        out = torch.empty((N * M, C, T, V), device=x.device, dtype=act_dtype)
        save = torch.empty((2, nch), device=x.device, dtype=torch.float32)
        ops.data_bn_fwd(x, num_point, fold_m, bn, train, out, save[0], save[1])
        ctx.bn, ctx.num_point, ctx.fold_m, ctx.train = bn, num_point, fold_m, train
        ctx.save_for_backward(x, save)
        return out

    @staticmethod
    def backward(ctx, g):
        x, save = ctx.saved_tensors
        bn = ctx.bn
        go = _GradOut(x)
        g = g.contiguous()
        dgamma, dbeta = go.buf(bn.weight), go.buf(bn.bias)
        dx = torch.empty(x.shape if x.dim() == 5 else (x.shape[0], x.shape[2] // ctx.num_point, x.shape[1], ctx.num_point, 1),
                         device=x.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        ops.data_bn_bwd(g, x, ctx.num_point, ctx.fold_m, bn.weight, save[0], save[1], ctx.train, dgamma, dbeta, dx)
        if dx is not None and x.dim() == 3:                 # (N,C,T,V,1) -> (N,T,V*C)
            dx = dx[..., 0].permute(0, 2, 3, 1).reshape(x.shape)
        return dx, None, None, None, None, go.ret(dgamma, bn.weight), go.ret(dbeta, bn.bias)


class PoolFcFn(torch.autograd.Function):
    """logits = Linear(mean over persons and (T, V) of x)  (models/ctrgcn.py:343-348; models/stgcn.py:187-195 with
    the 1x1 `fcn` conv on the pooled feature being the same linear map).  x: (N*M, C, T, V).
    weight None: pooling only (returns the (N, C) pooled feature)."""

    @staticmethod
    def forward(ctx, x, M, weight, bias):
        x = _check_input(x)
        NM, C, T, V = x.shape
        N = NM // M
        W = None
        if weight is not None:
            W = weight.reshape(weight.shape[0], -1)
            if W.shape[1] != C or W.dtype != torch.float32:
                raise ValueError('classifier expects %d fp32 input features, got %d' % (W.shape[1], C))
        pooled = torch.empty((N, C), device=x.device, dtype=torch.float32)
        logits = torch.empty((N, W.shape[0]), device=x.device, dtype=torch.float32) if W is not None else None
        ops.pool_fc_fwd(x, M, W, bias, pooled, logits)
        ctx.M, ctx.xmeta = M, (x.shape, x.dtype)
        ctx.save_for_backward(pooled, weight, bias)
        return logits if W is not None else pooled

    @staticmethod
    def backward(ctx, dl):
        pooled, weight, bias = ctx.saved_tensors
        shape, dtype = ctx.xmeta
        go = _GradOut(pooled)
        dl = dl.contiguous().float()
        g = torch.empty(shape, device=pooled.device, dtype=dtype) if ctx.needs_input_grad[0] else None
        if weight is None:
            if g is not None:
                ops.pool_fc_bwd(dl, pooled, None, ctx.M, g, None, None)
            return g, None, None, None
        W = weight.reshape(weight.shape[0], -1)
        dW = go.buf(weight, W.shape)
        db = go.buf(bias) if bias is not None else None
        ops.pool_fc_bwd(dl, pooled, W, ctx.M, g, dW, db)
        return g, None, go.ret(dW.view_as(weight), weight), go.ret(db, bias)


class CrossEntropyFn(torch.autograd.Function):
    """nn.CrossEntropyLoss() of the reference's processors (processor/recognition_rgb.py:19,61): mean over the batch of
    -log softmax(logits)[label]; labels outside [0, K) (e.g. -100) are ignored."""

    @staticmethod
    def forward(ctx, logits, labels):
        _require_cuda(logits)
        logits = logits.contiguous().float()
        labels = labels.contiguous().long()
        loss = torch.empty((1,), device=logits.device, dtype=torch.float32)
        dl = torch.empty_like(logits)
        ops.softmax_ce_fwd(logits, labels, loss, dl)
        ctx.save_for_backward(dl)
        return loss.view(())

    @staticmethod
    def backward(ctx, gl):
        dl, = ctx.saved_tensors
        out = torch.empty_like(dl)
        ops.softmax_ce_bwd(dl, gl.reshape(1).contiguous().float(), out)
        return out, None


def cross_entropy(logits, labels):
    return CrossEntropyFn.apply(logits, labels)
