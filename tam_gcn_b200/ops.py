"""Tensor-level wrappers over the C-ABI (one Python function per entry point of include/tamgcn.h).

Everything here is a thin marshalling layer: it derives pointers / sizes / strides from torch CUDA
tensors and enqueues the kernel on torch's current stream.  No arithmetic happens in Python and there
is no fallback path — a missing library or a CPU tensor raises.

Activation arguments may be channel-slice VIEWS of wider contiguous (N, C, T, V) tensors; the sample
stride is taken from `stride(0)`.
"""
import contextlib
import ctypes as C

import os

import torch

from . import _C

RES_NONE, RES_IDENTITY, RES_AFFINE = _C.RES_NONE, _C.RES_IDENTITY, _C.RES_AFFINE


class Opnd:
    """Lazy operand  f(a[ch]*p + b[ch]*q + c[ch])  (see tamgcn_operand in include/tamgcn.h)."""
    __slots__ = ('p', 'q', 'a', 'b', 'c', 'relu')

    def __init__(self, p, q=None, a=None, b=None, c=None, relu=False):
        self.p, self.q, self.a, self.b, self.c, self.relu = p, q, a, b, c, relu


def _stream():
    if _branch_stream is not None:
        return _branch_stream.cuda_stream
    return torch.cuda.current_stream().cuda_stream


# ---- parallel branches ----------------------------------------------------------------------------------------------
# Independent kernels of one module (the branches of MultiScale_TemporalConv, the conv1/conv2 relation path next to
# conv3 in unit_gcn) are small and latency-bound; inside `branches(streams)` (the engine's training step) the code
# under `with branch(i):` is enqueued on extra stream i (ordered after everything already on the current stream) and
# `branch_join()` makes the current stream wait for all of them.  Without an active `branches` context the code simply
# runs inline on the current stream.  Rule for callers: every tensor touched inside a branch must stay referenced until
# the `branch_join()` that follows (allocate outputs BEFORE entering the branch; torch's current stream is not changed,
# so allocations still belong to the main stream).
_branch_pool = None          # list of torch.cuda.Stream
_branch_stream = None        # the stream of the branch being recorded, or None
_branch_used = []


@contextlib.contextmanager
def branches(streams):
    global _branch_pool
    old = _branch_pool
    _branch_pool = list(streams) if streams else None
    try:
        yield
    finally:
        branch_join()
        _branch_pool = old


@contextlib.contextmanager
def branch(i):
    """i = 0: the current stream itself; i >= 1: extra stream i - 1 (if the engine provided any)."""
    global _branch_stream
    if _branch_pool is None or i <= 0 or _branch_stream is not None:
        yield
        return
    st = _branch_pool[(i - 1) % len(_branch_pool)]
    st.wait_stream(torch.cuda.current_stream())
    if st not in _branch_used:
        _branch_used.append(st)
    _branch_stream = st
    try:
        yield
    finally:
        _branch_stream = None


def branch_join():
    if _branch_used:
        cur = torch.cuda.current_stream()
        for st in _branch_used:
            cur.wait_stream(st)
        del _branch_used[:]


# ---- side stream for work that is off the critical path of backward -------------------------------------------------
# Weight-gradient kernels feed nothing but the optimiser, so inside `side_stream(s)` they are enqueued on `s` (ordered
# after everything already on the current stream) and the current stream moves on to the next data-gradient kernel.
# Every tensor such a launch touches is kept referenced until `join_side_stream`, so the caching allocator cannot hand
# its memory to a later kernel of the main stream while the side stream still reads it.
class _Side:
    __slots__ = ('stream', 'keep', 'share')

    def __init__(self, stream, share=100):
        self.stream, self.keep, self.share = stream, [], int(share)


_side = None


@contextlib.contextmanager
def side_stream(stream, sm_share=50):
    """Weight-gradient kernels of the calling thread go to `stream` and may occupy `sm_share` percent of the SMs
    (measured on the NW-UCLA step: 100 % -> the two streams alternate, 50 % -> they overlap, 25 % -> the side stream
    becomes the critical path)."""
    global _side
    old = _side
    _side = _Side(stream, sm_share) if stream is not None else None
    old_share = _C.lib().tamgcn_set_wgrad_sm_share(int(sm_share)) if stream is not None else None
    try:
        yield
    finally:
        if old_share is not None:
            _C.lib().tamgcn_set_wgrad_sm_share(old_share)
        if _side is not None and _side.keep:
            join_side_stream(stream)
        _side = old


@contextlib.contextmanager
def main_sm_share(percent):
    """Cap the persistent convolution forward / data-gradient kernels at `percent` of the SMs inside the context."""
    if percent is None or percent >= 100:
        yield
        return
    old = _C.lib().tamgcn_set_main_sm_share(int(percent))
    try:
        yield
    finally:
        _C.lib().tamgcn_set_main_sm_share(old)


def join_side_stream(stream):
    """Make the current stream wait for the side stream and release the tensors held for it."""
    if stream is None:
        return
    torch.cuda.current_stream().wait_stream(stream)
    if _side is not None and _side.stream is stream:
        _side.keep.clear()


# ---- algorithmic bytes / FLOPs accounting (bench.py's step-level roofline) -----------------------------------------
_acct = None


@contextlib.contextmanager
def account(table):
    """Inside the context every wrapper adds [algorithmic bytes, FLOPs] of its call to table[family]."""
    global _acct
    old = _acct
    _acct = table
    try:
        yield table
    finally:
        _acct = old


def _count(family, nbytes, flops=0):
    if _acct is not None:
        e = _acct.setdefault(family, [0, 0])
        e[0] += int(nbytes)
        e[1] += int(flops)


def _opnd_elems(o):
    """elements read for a lazy operand (P, plus Q when it has a second term)"""
    if torch.is_tensor(o):
        return o.numel()
    return o.p.numel() * (2 if (o.q is not None and o.b is not None) else 1)


def _opnd_tensors(o):
    if torch.is_tensor(o):
        return (o,)
    return tuple(t for t in (o.p, o.q, o.a, o.b, o.c) if t is not None)


def _dt(t):
    if t.dtype == torch.float32:
        return _C.F32
    if t.dtype == torch.bfloat16:
        return _C.BF16
    raise TypeError('activation dtype %s not supported (float32 or bfloat16)' % t.dtype)


def _act(t, dtype=None):
    """(pointer, sample stride) of an (N, C, T, V) activation view with contiguous (T, V) planes."""
    if not t.is_cuda:
        raise RuntimeError('tam_gcn_b200 kernels need CUDA tensors (there is no CPU path)')
    if dtype is not None and t.dtype != dtype:
        raise TypeError('mixed activation dtypes: %s vs %s' % (t.dtype, dtype))
    N, Cc, T, V = t.shape
    s = t.stride()
    ok = (V == 1 or s[3] == 1) and (T == 1 or s[2] == V) and (Cc == 1 or s[1] == T * V)
    if not ok:
        raise ValueError('activation view must have contiguous (T,V) planes and channel stride T*V; '
                         'got shape %s strides %s' % (tuple(t.shape), s))
    return t.data_ptr(), (s[0] if N > 1 else Cc * T * V)


def _f32(t, n=None):
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise TypeError('expected a contiguous CUDA float32 tensor')
    if n is not None and t.numel() != n:
        raise ValueError('expected %d elements, got %d' % (n, t.numel()))
    return t.data_ptr()


def _f64(t, n=None):
    if t is None:
        return None
    if t.dtype != torch.float64 or not t.is_contiguous() or not t.is_cuda:
        raise TypeError('expected a contiguous CUDA float64 tensor')
    if n is not None and t.numel() != n:
        raise ValueError('expected %d elements, got %d' % (n, t.numel()))
    return t.data_ptr()


def _operand(o, nch):
    if torch.is_tensor(o):
        o = Opnd(o)
    dtype = o.p.dtype
    p, pns = _act(o.p)
    s = _C.Operand()
    s.p, s.p_nstride = p, pns
    if o.q is not None:
        if o.q.shape != o.p.shape:
            raise ValueError('operand p/q shape mismatch')
        s.q, s.q_nstride = _act(o.q, dtype)
    s.a, s.b, s.c = _f32(o.a, nch if o.a is not None else None), _f32(o.b, nch if o.b is not None else None), \
        _f32(o.c, nch if o.c is not None else None)
    s.relu = 1 if o.relu else 0
    return s


def _geom(N, Cin, Cout, T, To, V, k, stride, dil, pad):
    g = _C.ConvGeom()
    g.N, g.Cin, g.Cout, g.T, g.To, g.V, g.k, g.stride, g.dil, g.pad = N, Cin, Cout, T, To, V, k, stride, dil, pad
    return g


def _p(t):
    return None if t is None else t.data_ptr()


def conv_fwd(x, W, bias, y, k=1, stride=1, dil=1, pad=0, stats=None, stat_c0=0, wpack=None):
    """y = conv_{k x 1}(X) (+bias); optional (sum, sumsq) fp64 statistics of y for channels >= stat_c0."""
    xp = x.p if isinstance(x, Opnd) else x
    N, Cin, T, V = xp.shape
    _, Cout, To, _ = y.shape
    g = _geom(N, Cin, Cout, T, To, V, k, stride, dil, pad)
    xo = _operand(x, Cin)
    yp, yns = _act(y, xp.dtype)
    ssum = ssq = None
    if stats is not None:
        ssum, ssq = _f64(stats[0], Cout - stat_c0), _f64(stats[1], Cout - stat_c0)
    _C.check(_C.lib().tamgcn_conv_fwd(C.byref(g), _dt(xp), C.byref(xo), _f32(W, Cout * Cin * k), _p(wpack), _f32(bias), yp, yns,
                                      ssum, ssq, stat_c0, _stream()), 'tamgcn_conv_fwd')
    if _acct is not None:
        rows = min(T, To * k)
        _count('conv_fwd', xp.element_size() * (_opnd_elems(x) * rows // T + y.numel()) + 4 * Cout * Cin * k,
               2 * N * To * V * Cout * Cin * k)


def conv_dgrad(dy, W, dx, k=1, stride=1, dil=1, pad=0, addend=None, bcast=None, bcast_scale=0.0, mask=None,
               stats=None, wpack=None):
    """dx = conv_transpose(dY) (+addend) (+bcast*scale); optional ReLU mask + BN-backward sums."""
    dyp = dy.p if isinstance(dy, Opnd) else dy
    N, Cout, To, V = dyp.shape
    _, Cin, T, _ = dx.shape
    g = _geom(N, Cin, Cout, T, To, V, k, stride, dil, pad)
    dyo = _operand(dy, Cout)
    dxp, dxns = _act(dx, dyp.dtype)
    ap = ans = None
    if addend is not None:
        if addend.shape != dx.shape:
            raise ValueError('addend shape mismatch')
        ap, ans = _act(addend, dyp.dtype)
    mo = None
    if mask is not None:
        if mask.p.shape != dx.shape:
            raise ValueError('mask shape mismatch')
        mo = C.byref(_operand(mask, Cin))
    s1 = s2 = None
    if stats is not None:
        s1, s2 = _f64(stats[0], Cin), _f64(stats[1], Cin)
    _C.check(_C.lib().tamgcn_conv_dgrad(C.byref(g), _dt(dyp), C.byref(dyo), _f32(W, Cout * Cin * k), _p(wpack), dxp, dxns, ap,
                                        ans or 0, _f32(bcast, N * Cin * V if bcast is not None else None),
                                        float(bcast_scale), mo, s1, s2, _stream()), 'tamgcn_conv_dgrad')
    if _acct is not None:
        extra = (addend.numel() if addend is not None else 0) + (mask.p.numel() if mask is not None else 0)
        _count('conv_dgrad', dyp.element_size() * (_opnd_elems(dy) + dx.numel() * min(T, To * k) // T + extra) + 4 * Cout * Cin * k,
               2 * N * To * V * Cout * Cin * k)


def conv_pack_weights(W, Cout, Cin, k, stride=1, V=0):
    """bf16 tensor-core weight tiles for the forward and the data-gradient GEMM -> (wpack_fwd, wpack_dgrad).
    An entry is None when the kernel serving that shape reads the fp32 weights directly (tamgcn_conv_needs_pack)."""
    l = _C.lib()
    nf = V <= 0 or l.tamgcn_conv_needs_pack(Cin, Cout, k, stride, V, 0)
    nd = V <= 0 or l.tamgcn_conv_needs_pack(Cin, Cout, k, stride, V, 1)
    wf = torch.empty(l.tamgcn_conv_pack_bytes(Cout, Cin, k, 0), dtype=torch.uint8, device=W.device) if nf else None
    wd = torch.empty(l.tamgcn_conv_pack_bytes(Cout, Cin, k, 1), dtype=torch.uint8, device=W.device) if nd else None
    if nf or nd:
        _C.check(l.tamgcn_conv_pack_weights(_f32(W, Cout * Cin * k), Cout, Cin, k, _p(wf), _p(wd), _stream()),
                 'tamgcn_conv_pack_weights')
    return wf, wd


class PackCache:
    """Persistent tensor-core weight tiles of a model whose parameters change once per step (the optimiser) or never
    (inference).  `get` hands out the same buffers for the same weight matrix; `repack_all` refreshes ALL of them in one
    launch (tamgcn_conv_pack_weights_batched) — a step engine calls it at the start of the step, so the ~46 per-layer
    pack launches of a CTR-GCN step become one.  Until `repack_all` has run in the current step (`fresh`), `get` packs
    the matrix it returns itself, so a cache is never stale."""

    def __init__(self, model):
        self.model = model
        self.storages = set()
        self.entries = {}            # key -> (wf, wd, W)
        self.table = None
        self.fresh = False

    def _is_parameter_memory(self, W):
        """Only views of the model's own parameters have an address that means the same matrix next step (a torch.cat
        of DataParallel replica weights does not)."""
        sp = W.untyped_storage().data_ptr()
        if sp not in self.storages:
            self.storages = {p.untyped_storage().data_ptr() for p in self.model.parameters()}
        return sp in self.storages

    def get(self, W, Cout, Cin, k, stride, V):
        key = (W.data_ptr(), Cout, Cin, k, stride, V, W.device)
        e = self.entries.get(key)
        if e is None:
            wf, wd = conv_pack_weights(W, Cout, Cin, k, stride, V)
            if not self._is_parameter_memory(W):
                return wf, wd
            self.entries[key] = (wf, wd, W)
            self.table = None
            return wf, wd
        wf, wd, _ = e
        if not self.fresh and (wf is not None or wd is not None):
            _C.check(_C.lib().tamgcn_conv_pack_weights(_f32(W, Cout * Cin * k), Cout, Cin, k, _p(wf), _p(wd), _stream()),
                     'tamgcn_conv_pack_weights')
        return wf, wd

    def repack_all(self):
        jobs = [(W, wf, wd, key) for key, (wf, wd, W) in self.entries.items() if wf is not None or wd is not None]
        if not jobs:
            return
        if self.table is None:
            if jobs[0][0].is_cuda and torch.cuda.is_current_stream_capturing():
                return                               # cannot build the table now: `get` keeps packing per call
            rows = [[W.data_ptr(), _p(wf) or 0, _p(wd) or 0, key[1], key[2], key[3], 0, 0] for W, wf, wd, key in jobs]
            self.table = torch.tensor(rows, dtype=torch.int64).to(jobs[0][0].device)
        _C.check(_C.lib().tamgcn_conv_pack_weights_batched(self.table.data_ptr(), self.table.shape[0], _stream()),
                 'tamgcn_conv_pack_weights_batched')
        self.fresh = True

    def end_step(self):
        self.fresh = False


def conv_wgrad(dy, x, dW, dbias, k=1, stride=1, dil=1, pad=0):
    """dW += dY (*) X, dbias += sum dY  (fp32 accumulators, zeroed by the caller).  Runs on the side stream when one
    is active (see `side_stream`)."""
    dyp = dy.p if isinstance(dy, Opnd) else dy
    xp = x.p if isinstance(x, Opnd) else x
    N, Cout, To, V = dyp.shape
    _, Cin, T, _ = xp.shape
    g = _geom(N, Cin, Cout, T, To, V, k, stride, dil, pad)
    dyo, xo = _operand(dy, Cout), _operand(x, Cin)
    if xp.dtype != dyp.dtype:
        raise TypeError('conv_wgrad: mixed activation dtypes')
    args = (C.byref(g), _dt(dyp), C.byref(dyo), C.byref(xo), _f32(dW, Cout * Cin * k),
            _f32(dbias, Cout if dbias is not None else None))
    if _acct is not None:
        _count('conv_wgrad', dyp.element_size() * (_opnd_elems(dy) + _opnd_elems(x) * min(T, To * k) // T) + 4 * Cout * Cin * k,
               2 * N * To * V * Cout * Cin * k)
    sd = _side
    if sd is None:
        _C.check(_C.lib().tamgcn_conv_wgrad(*args, _stream()), 'tamgcn_conv_wgrad')
        return
    sd.stream.wait_stream(_branch_stream if _branch_stream is not None else torch.cuda.current_stream())
    sd.keep.append(_opnd_tensors(dy) + _opnd_tensors(x) + (dW, dbias))
    _C.check(_C.lib().tamgcn_conv_wgrad(*args, sd.stream.cuda_stream), 'tamgcn_conv_wgrad')


def mean_t(x, m):
    """m[n,c,0,v] = mean_t x[n,c,t,v]; m is fp32 (N,C,1,V)."""
    N, Cc, T, V = x.shape
    xp, xns = _act(x)
    _C.check(_C.lib().tamgcn_mean_t(_dt(x), xp, xns, N, Cc, T, V, _f32(m, N * Cc * V), _stream()), 'tamgcn_mean_t')
    _count('epilogues+maxpool', x.element_size() * x.numel() + 4 * m.numel())


def ctrgc_fwd(x3, x1, x2, W4, b4, PA, alpha, y, stats=None):
    """Fused CTRGC forward.  x3: (N,K*Cout,T,V); x1,x2: fp32 (N,K*R,1,V) views; y: (N,Cout,T,V)."""
    N, KC, T, V = x3.shape
    K = PA.shape[0]
    Cout = KC // K
    R = x1.shape[1] // K
    x3p, x3ns = _act(x3)
    yp, yns = _act(y, x3.dtype)
    x1p, x12ns = _act(x1, torch.float32)
    x2p, x12ns2 = _act(x2, torch.float32)
    assert x12ns == x12ns2
    ssum = ssq = None
    if stats is not None:
        ssum, ssq = _f64(stats[0], Cout), _f64(stats[1], Cout)
    _C.check(_C.lib().tamgcn_ctrgc_fwd(_dt(x3), x3p, x3ns, N, Cout, T, V, K, R, x1p, x2p, x12ns,
                                       _f32(W4, K * Cout * R), _f32(b4, K * Cout), _f32(PA, K * V * V),
                                       _f32(alpha, 1), yp, yns, ssum, ssq, _stream()), 'tamgcn_ctrgc_fwd')
    _count('ctrgc_fwd', x3.element_size() * N * T * V * Cout * (K + 1) + K * 8 * N * R * V + K * 4 * (Cout * R + Cout + V * V),
           2 * K * N * Cout * T * V * V)


def ctrgc_bwd(g, x3, x1, x2, W4, b4, PA, alpha, dx3, dx1, dx2, dW4, db4, dPA, dalpha):
    """Fused CTRGC backward (dx3 written; dx1, dx2, dW4, db4, dPA, dalpha accumulated)."""
    N, KC, T, V = x3.shape
    K = PA.shape[0]
    Cout = KC // K
    R = x1.shape[1] // K
    go = _operand(g, Cout)
    x3p, x3ns = _act(x3)
    dx3p, dx3ns = _act(dx3, x3.dtype)
    x1p, x12ns = _act(x1, torch.float32)
    x2p, _ = _act(x2, torch.float32)
    d1p, d12ns = _act(dx1, torch.float32)
    d2p, _ = _act(dx2, torch.float32)
    assert d12ns == x12ns
    _C.check(_C.lib().tamgcn_ctrgc_bwd(_dt(x3), C.byref(go), x3p, x3ns, N, Cout, T, V, K, R, x1p, x2p, x12ns,
                                       _f32(W4, K * Cout * R), _f32(b4, K * Cout), _f32(PA, K * V * V),
                                       _f32(alpha, 1), dx3p, dx3ns, d1p, d2p, _f32(dW4, K * Cout * R),
                                       _f32(db4, K * Cout), _f32(dPA, K * V * V), _f32(dalpha, 1), _stream()),
             'tamgcn_ctrgc_bwd')
    _count('ctrgc_bwd', x3.element_size() * N * T * V * Cout * (1 + 2 * K) + 2 * K * 8 * N * R * V + K * 4 * (Cout * R + Cout + V * V),
           4 * K * N * Cout * T * V * V)


def bn_finalize(descs, count, momentum, eps, train):
    """descs: list of dicts with keys sum, sumsq, gamma, beta, rmean, rvar, nbt, scale, shift, mean, invstd."""
    for i in range(0, len(descs), 8):
        chunk = descs[i:i + 8]
        arr = (_C.BnDesc * len(chunk))()
        for s, d in zip(arr, chunk):
            Cn = d['scale'].numel()
            s.sum, s.sumsq = _f64(d.get('sum'), Cn if d.get('sum') is not None else None), \
                _f64(d.get('sumsq'), Cn if d.get('sumsq') is not None else None)
            s.gamma, s.beta = _p(d.get('gamma')), _p(d.get('beta'))
            s.rmean, s.rvar, s.nbt = _p(d.get('rmean')), _p(d.get('rvar')), _p(d.get('nbt'))
            s.scale, s.shift = _f32(d['scale'], Cn), _f32(d['shift'], Cn)
            s.mean, s.invstd = _f32(d.get('mean'), Cn if d.get('mean') is not None else None), \
                _f32(d.get('invstd'), Cn if d.get('invstd') is not None else None)
            s.C = Cn
        _C.check(_C.lib().tamgcn_bn_finalize(len(chunk), arr, float(count), float(momentum), float(eps),
                                             1 if train else 0, _stream()), 'tamgcn_bn_finalize')


def bn_bwd_coef(descs, count, train):
    """descs: list of dicts with keys s1, s2, gamma, mean, invstd, A, B, Cc, dgamma, dbeta."""
    for i in range(0, len(descs), 8):
        chunk = descs[i:i + 8]
        arr = (_C.BnBwdDesc * len(chunk))()
        for s, d in zip(arr, chunk):
            Cn = d['A'].numel()
            s.s1, s.s2 = _f64(d['s1'], Cn), _f64(d['s2'], Cn)
            s.gamma, s.mean, s.invstd = _p(d.get('gamma')), _f32(d['mean'], Cn), _f32(d['invstd'], Cn)
            s.A, s.B, s.Cc = _f32(d['A'], Cn), _f32(d['B'], Cn), _f32(d['Cc'], Cn)
            s.dgamma, s.dbeta = _p(d.get('dgamma')), _p(d.get('dbeta'))
            s.C = Cn
        _C.check(_C.lib().tamgcn_bn_bwd_coef(len(chunk), arr, float(count), 1 if train else 0, _stream()),
                 'tamgcn_bn_bwd_coef')


def _full(t, dtype):
    if not t.is_contiguous() or t.dtype != dtype or not t.is_cuda:
        raise TypeError('expected a contiguous CUDA %s tensor' % dtype)
    return t.data_ptr()


def _res(res_mode, r, dtype):
    if res_mode == RES_NONE or r is None:
        return None, 0
    return _act(r, dtype)


def gcn_epilogue_fwd(y0, sg, hg, z, so, ho, res_mode, r, sr, hr, out):
    N, Cc, T, V = y0.shape
    dt = y0.dtype
    rp, rns = _res(res_mode, r, dt)
    _C.check(_C.lib().tamgcn_gcn_epilogue_fwd(_dt(y0), N, Cc, T * V, _full(y0, dt), _f32(sg, Cc), _f32(hg, Cc),
                                              _full(z, dt), _f32(so, Cc), _f32(ho, Cc), res_mode, rp, rns, _f32(sr),
                                              _f32(hr), _full(out, dt), _stream()), 'tamgcn_gcn_epilogue_fwd')
    _count('epilogues+maxpool', y0.element_size() * y0.numel() * (3 + (1 if r is not None else 0)))


def gcn_epilogue_bwd(g, out, z, so, ho, G, DZ, s1o, s2o):
    N, Cc, T, V = g.shape
    dt = g.dtype
    _C.check(_C.lib().tamgcn_gcn_epilogue_bwd(_dt(g), N, Cc, T * V, _full(g, dt), _full(out, dt), _full(z, dt),
                                              _f32(so, Cc), _f32(ho, Cc), _full(G, dt), _full(DZ, dt), _f64(s1o, Cc),
                                              _f64(s2o, Cc), _stream()), 'tamgcn_gcn_epilogue_bwd')
    _count('epilogues+maxpool', g.element_size() * g.numel() * 5)


def gcn_mid_bwd(G, DD, dr, y0, r, s1g, s2g, s1d, s2d, extra=None):
    N, Cc, T, V = G.shape
    dt = G.dtype
    drp, drns = (None, 0) if dr is None else _act(dr, dt)
    rp, rns = (None, 0) if r is None else _act(r, dt)
    ep, ens = (None, 0) if extra is None else _act(extra, dt)
    _C.check(_C.lib().tamgcn_gcn_mid_bwd(_dt(G), N, Cc, T * V, _full(G, dt), _full(DD, dt), drp, drns, _full(y0, dt),
                                         rp, rns, _f64(s1g, Cc), _f64(s2g, Cc), _f64(s1d), _f64(s2d), ep, ens, _stream()),
             'tamgcn_gcn_mid_bwd')
    _count('epilogues+maxpool', G.element_size() * G.numel() * (4 + (1 if dr is not None else 0) + (1 if r is not None else 0) +
                                                               (1 if extra is not None else 0)))


def coef_diff(sb, ha, hb, nb, c):
    """nb = -sb, c = (ha or 0) - hb  (per-channel fp32 coefficient rows)."""
    Cn = sb.numel()
    _C.check(_C.lib().tamgcn_coef_diff(Cn, _f32(sb, Cn), _f32(ha), _f32(hb, Cn), _f32(nb, Cn), _f32(c, Cn), _stream()),
             'tamgcn_coef_diff')


def tcn_epilogue_fwd(u, su, hu, res_mode, r, sr, hr, relu, out):
    N, Cc, T, V = u.shape
    dt = u.dtype
    up, uns = _act(u)
    rp, rns = _res(res_mode, r, dt)
    _C.check(_C.lib().tamgcn_tcn_epilogue_fwd(_dt(u), N, Cc, T * V, up, uns, _f32(su, Cc), _f32(hu, Cc), res_mode, rp,
                                              rns, _f32(sr), _f32(hr), 1 if relu else 0, _full(out, dt), _stream()),
             'tamgcn_tcn_epilogue_fwd')
    _count('epilogues+maxpool', u.element_size() * u.numel() * (2 + (1 if r is not None else 0)))


def tcn_epilogue_bwd(g, out, relu, u, r, G, s1, s2u, s2r):
    N, Cc, T, V = g.shape
    dt = g.dtype
    up, uns = _act(u, dt)
    rp, rns = (None, 0) if r is None else _act(r, dt)
    _C.check(_C.lib().tamgcn_tcn_epilogue_bwd(_dt(g), N, Cc, T * V, _full(g, dt), _p(out), 1 if relu else 0, up, uns,
                                              rp, rns, _p(G), _f64(s1, Cc), _f64(s2u, Cc), _f64(s2r), _stream()),
             'tamgcn_tcn_epilogue_bwd')
    _count('epilogues+maxpool', g.element_size() * g.numel() * (2 + (1 if out is not None else 0) + (1 if G is not None else 0) +
                                                               (1 if r is not None else 0)))


def maxpool_fwd(x, y, stride, stats=None):
    xp = x.p if isinstance(x, Opnd) else x
    N, Cc, T, V = xp.shape
    To = y.shape[2]
    xo = _operand(x, Cc)
    yp, yns = _act(y, xp.dtype)
    ssum = ssq = None
    if stats is not None:
        ssum, ssq = _f64(stats[0], Cc), _f64(stats[1], Cc)
    _C.check(_C.lib().tamgcn_maxpool_fwd(_dt(xp), N, Cc, T, To, V, stride, C.byref(xo), yp, yns, ssum, ssq,
                                         _stream()), 'tamgcn_maxpool_fwd')
    _count('epilogues+maxpool', xp.element_size() * (xp.numel() + y.numel()))


def maxpool_bwd(dy, x, dh, stride, stats=None):
    dyp = dy.p if isinstance(dy, Opnd) else dy
    N, Cc, To, V = dyp.shape
    T = x.p.shape[2]
    dyo, xo = _operand(dy, Cc), _operand(x, Cc)
    dhp, dhns = _act(dh, dyp.dtype)
    s1 = s2 = None
    if stats is not None:
        s1, s2 = _f64(stats[0], Cc), _f64(stats[1], Cc)
    _C.check(_C.lib().tamgcn_maxpool_bwd(_dt(dyp), N, Cc, T, To, V, stride, C.byref(dyo), C.byref(xo), dhp, dhns, s1,
                                         s2, _stream()), 'tamgcn_maxpool_bwd')
    _count('epilogues+maxpool', dyp.element_size() * (_opnd_elems(dy) + x.p.numel() + dh.numel()))


def graph_agg_fwd(y, A, out, stats=None):
    """out[n,c,t,w] = sum_{k,v} y[n,k*C+c,t,v] A[k,v,w]."""
    N, KC, T, V = y.shape
    K = A.shape[0]
    Cc = KC // K
    yp, yns = _act(y)
    op, ons = _act(out, y.dtype)
    ssum = ssq = None
    if stats is not None:
        ssum, ssq = _f64(stats[0], Cc), _f64(stats[1], Cc)
    _C.check(_C.lib().tamgcn_graph_agg_fwd(_dt(y), N, K, Cc, T, V, yp, yns, _f32(A, K * V * V), op, ons, ssum, ssq,
                                           _stream()), 'tamgcn_graph_agg_fwd')
    _count('graph_agg', y.element_size() * (y.numel() + out.numel()), 2 * N * KC * T * V * V)


def graph_agg_bwd(dout, y, A, dy, dA):
    """dy = dOut . A^T per subset; dA += y^T . dOut.  dA feeds only the optimiser: TAMGCN_AGG_DA_SIDE=1 computes it on
    the side stream (like the convolution weight gradients)."""
    dp = dout.p if isinstance(dout, Opnd) else dout
    N, Cc, T, V = dp.shape
    K = A.shape[0]
    do = _operand(dout, Cc)
    yp, yns = _act(y, dp.dtype)
    dyp, dyns = _act(dy, dp.dtype)
    lib = _C.lib()
    dAp = _f32(dA, K * V * V if dA is not None else None)
    _count('graph_agg', dp.element_size() * (_opnd_elems(dout) + y.numel() + dy.numel()), 4 * N * K * Cc * T * V * V)
    # (measured on the ST-GCN NTU step: 21.3 ms with dA on the side stream, 20.2 ms inline — the two streams share the
    # SMs, so the step is the sum of the kernels either way and the extra hand-off only costs; opt-in)
    sd = _side if os.environ.get('TAMGCN_AGG_DA_SIDE', '0') == '1' else None
    if sd is None or dA is None:
        _C.check(lib.tamgcn_graph_agg_bwd(_dt(dp), N, K, Cc, T, V, C.byref(do), yp, yns, _f32(A, K * V * V), dyp, dyns, dAp,
                                          _stream()), 'tamgcn_graph_agg_bwd')
        return
    _C.check(lib.tamgcn_graph_agg_bwd(_dt(dp), N, K, Cc, T, V, C.byref(do), yp, yns, _f32(A, K * V * V), dyp, dyns, None,
                                      _stream()), 'tamgcn_graph_agg_bwd')
    sd.stream.wait_stream(_branch_stream if _branch_stream is not None else torch.cuda.current_stream())
    sd.keep.append(_opnd_tensors(dout) + (y, A, dA))
    _C.check(lib.tamgcn_graph_agg_bwd(_dt(dp), N, K, Cc, T, V, C.byref(do), yp, yns, _f32(A, K * V * V), None, 0, dAp,
                                      sd.stream.cuda_stream), 'tamgcn_graph_agg_bwd')


# ---- network ends and optimiser (csrc/head.cu) ---------------------------------------------------------------------
def _strides5(x, fold_3d_num_point=None):
    """Element strides of the (n, c, t, v, m) axes of a 5-D (N,C,T,V,M) or 3-D (N,T,V*C) fp32 input."""
    if x.dim() == 5:
        return tuple(x.shape), tuple(x.stride())
    N, T, VC = x.shape
    V = fold_3d_num_point
    Cc = VC // V
    sn, st, s2 = x.stride()
    return (N, Cc, T, V, 1), (sn, s2, st, Cc * s2, 0)          # x.view(N,T,V,C).permute(0,3,1,2).unsqueeze(-1)


def data_bn_fwd(x, num_point, fold_m, bn, train, out, save_mean, save_invstd):
    """BatchNorm1d prologue of Model.forward; x: fp32 (N,C,T,V,M) or (N,T,V*C); out: (N*M, C, T, V)."""
    if x.dtype != torch.float32 or not x.is_cuda:
        raise TypeError('data_bn: expected a CUDA float32 input')
    (N, Cc, T, V, M), st = _strides5(x, num_point)
    arr = (_C.i64 * 5)(*st)
    _C.check(_C.lib().tamgcn_data_bn_fwd(_dt(out), x.data_ptr(), arr, N, Cc, T, V, M, 1 if fold_m else 0, _p(bn.weight),
                                         _p(bn.bias), _p(bn.running_mean), _p(bn.running_var),
                                         _p(bn.num_batches_tracked) if train else None, float(bn.momentum), float(bn.eps),
                                         1 if train else 0, _full(out, out.dtype), _f32(save_mean), _f32(save_invstd),
                                         _stream()), 'tamgcn_data_bn_fwd')
    _count('head+sgd', 4 * x.numel() + out.element_size() * out.numel())


def data_bn_bwd(g, x, num_point, fold_m, gamma, mean, invstd, train, dgamma, dbeta, dx):
    (N, Cc, T, V, M), st = _strides5(x, num_point)
    arr = (_C.i64 * 5)(*st)
    _C.check(_C.lib().tamgcn_data_bn_bwd(_dt(g), _full(g, g.dtype), x.data_ptr(), arr, N, Cc, T, V, M, 1 if fold_m else 0,
                                         _p(gamma), _f32(mean), _f32(invstd), 1 if train else 0, _f32(dgamma), _f32(dbeta),
                                         _f32(dx), _stream()), 'tamgcn_data_bn_bwd')
    _count('head+sgd', 4 * x.numel() + g.element_size() * g.numel() + (4 * x.numel() if dx is not None else 0))


def pool_fc_fwd(x, M, W, b, pooled, logits, gate=None):
    """x: (N*M, C, T, V); pooled (N, C) fp32 (raw mean); logits (N, K) fp32 = (gate .* pooled) W^T + b (W None: pooling only)."""
    NM, Cc, T, V = x.shape
    N = NM // M
    K = W.shape[0] if W is not None else 0
    _C.check(_C.lib().tamgcn_pool_fc_fwd(_dt(x), _full(x, x.dtype), N, M, Cc, T * V, K, _f32(gate, N * Cc if gate is not None else None),
                                         _f32(W, K * Cc if W is not None else None), _f32(b), _f32(pooled, N * Cc),
                                         _f32(logits), _stream()), 'tamgcn_pool_fc_fwd')
    _count('head+sgd', x.element_size() * x.numel())


def pool_fc_bwd(dlogits, pooled, W, M, g, dW, db, gate=None, dgate=None):
    """W None: pooling only (dlogits is the cotangent of pooled, K == C)."""
    N, K = dlogits.shape
    Cc = pooled.shape[1]
    if g is not None:
        TV = g.shape[2] * g.shape[3]
        dt = _dt(g)
    else:
        TV, dt = 1, _C.F32
    _C.check(_C.lib().tamgcn_pool_fc_bwd(dt, _f32(dlogits), _f32(pooled), _f32(gate), _f32(W, K * Cc if W is not None else None),
                                         N, M, Cc, TV, K, None if g is None else _full(g, g.dtype), _f32(dW), _f32(db),
                                         _f32(dgate), _stream()), 'tamgcn_pool_fc_bwd')
    _count('head+sgd', g.element_size() * g.numel() if g is not None else 0)


def transpose_act(inp, out, mode=0, aux=None):
    """out (C, R) = f(inp (R, C))^T;  mode 0 identity, 1 sigmoid, 2 sigmoid backward (inp * aux * (1 - aux))."""
    R, Cc = inp.shape
    _C.check(_C.lib().tamgcn_transpose_act(_f32(inp), _f32(aux, R * Cc if aux is not None else None), R, Cc, mode,
                                           _f32(out, R * Cc), _stream()), 'tamgcn_transpose_act')


def softmax_ce_fwd(logits, labels, loss, dl):
    N, K = logits.shape
    if labels.dtype != torch.int64 or not labels.is_contiguous():
        raise TypeError('softmax_ce: labels must be contiguous int64')
    _C.check(_C.lib().tamgcn_softmax_ce_fwd(_f32(logits), labels.data_ptr(), N, K, _f32(loss, 1), _f32(dl), _stream()),
             'tamgcn_softmax_ce_fwd')


def softmax_ce_bwd(dl, gloss, out):
    N, K = dl.shape
    _C.check(_C.lib().tamgcn_softmax_ce_bwd(_f32(dl), _f32(gloss, 1), N, K, _f32(out), _stream()), 'tamgcn_softmax_ce_bwd')


def sgd_step(P, G, Mo, lr, momentum, weight_decay, nesterov, grad_scale=1.0):
    """In-place SGD over flat fp32 buffers; lr is a 1-element CUDA tensor."""
    n = P.numel()
    _C.check(_C.lib().tamgcn_sgd_step(_f32(P), _f32(G, n), _f32(Mo, n), n, _f32(lr, 1), float(momentum),
                                      float(weight_decay), 1 if nesterov else 0, float(grad_scale), _stream()),
             'tamgcn_sgd_step')
    _count('head+sgd', 20 * n)


def feeder_nucla(raw, length, sample, view, frame_idx, bone_parent, mode, out):
    """GPU feeder (csrc/feeder.cu).  raw (S, Lmax, V, 3) fp32, length (S) int32, sample (B) int64, view (B, 3) fp32,
    frame_idx (B, T) int32, bone_parent (V) int32 or None, out (B, 3, T, V, 1) fp32."""
    S, Lmax, V, _ = raw.shape
    B, T = frame_idx.shape
    for t, dt in ((raw, torch.float32), (length, torch.int32), (sample, torch.int64), (view, torch.float32),
                  (frame_idx, torch.int32), (out, torch.float32)):
        if t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
            raise TypeError('feeder_nucla: expected contiguous CUDA %s tensors' % dt)
    if out.shape != (B, 3, T, V, 1) or view.shape != (B, 3) or sample.shape != (B,):
        raise ValueError('feeder_nucla: shape mismatch')
    _C.check(_C.lib().tamgcn_feeder_nucla(raw.data_ptr(), length.data_ptr(), sample.data_ptr(), view.data_ptr(),
                                          frame_idx.data_ptr(), _p(bone_parent), B, Lmax, V, T, int(mode), out.data_ptr(),
                                          _stream()), 'tamgcn_feeder_nucla')
