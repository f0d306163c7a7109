"""Flat parameter / gradient storage of one model (host-side runtime of the training step).

`ParamStore(model)` moves every parameter of `model` into ONE contiguous fp32 buffer `P` (the Parameter objects, their
names and shapes are untouched — they become views, so `state_dict`, `load_state_dict` and optimisers keep working)
and allocates a gradient buffer `G` with the same layout.  What that buys the step (SURVEY.md §8 e/f1):

  * the SGD update is one kernel over (P, G, momentum) instead of one multi-tensor launch per 50 tensors;
  * the data-parallel gradient all-reduce runs on slices of `G` directly — no gather copy, no re-pointing;
  * the parameter groups the kernels want stacked (conv1|conv2 of the K CTRGC subsets, conv3|down, the 1x1 heads of
    MultiScale_TemporalConv ...) are adjacent in `P`, so `functional._packed` finds them already packed, and their
    weight-gradient accumulators are the matching slices of `G`: the backward kernels accumulate straight into the
    bucket that is all-reduced (`functional.grad_sink`).

Layout: parameters in `model.named_parameters()` order (so a layer's parameters are contiguous and later layers sit at
higher offsets — the all-reduce of the upper part of `G` can start while the lower layers are still in backward); a pack
group is placed where its first member appears; every group / lone parameter starts on a 256-byte boundary.
"""
import contextlib

import torch
import torch.nn as nn

_ALIGN = 64          # floats


def _align(n):
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


class ParamStore:
    def __init__(self, model):
        from . import functional as Fn
        params = [(k, p) for k, p in model.named_parameters()]
        if not params:
            raise ValueError('model has no parameters')
        dev = params[0][1].device
        for k, p in params:
            if p.dtype != torch.float32:
                raise TypeError('ParamStore: parameter %s is %s; master parameters must be float32' % (k, p.dtype))
            if p.device != dev:
                raise ValueError('ParamStore: parameters live on different devices (%s vs %s)' % (p.device, dev))
        # pack groups declared by the modules (owner module, key, parameter list)
        group_of = {}
        groups = []
        for mod in model.modules():
            for owner, key, ps in Fn.pack_groups(mod):
                if all(isinstance(p, nn.Parameter) for p in ps) and not any(id(p) in group_of for p in ps):
                    gi = len(groups)
                    groups.append([owner, key, ps, None])
                    for p in ps:
                        group_of[id(p)] = gi
        self.offsets = {}                      # id(param) -> offset (floats)
        self.order = []                        # [(name, param, offset)]
        names = {id(p): k for k, p in params}
        off = 0
        for k, p in params:
            if id(p) in self.offsets:
                continue
            gi = group_of.get(id(p))
            off = _align(off)
            if gi is None:
                self.offsets[id(p)] = off
                self.order.append((k, p, off))
                off += p.numel()
            else:
                groups[gi][3] = off
                for q in groups[gi][2]:
                    self.offsets[id(q)] = off
                    self.order.append((names[id(q)], q, off))
                    off += q.numel()
        self.numel = _align(off)
        self.P = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.G = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for _, p, o in self.order:
                n = p.numel()
                self.P[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.P[o:o + n].view(p.shape)
        for owner, key, ps, start in groups:
            n = sum(p.numel() for p in ps)
            owner.__dict__.setdefault('_tamgcn_packs', {})[key] = (self.P[start:start + n], [p.data_ptr() for p in ps])
        self.device = dev
        self._base = self.P.data_ptr()

    # ---- queries ---------------------------------------------------------------------------------------------
    def valid(self):
        """True while every parameter still aliases its slot of P (`.to()`, `.data = ...` break that)."""
        b = self._base
        return all(p.data_ptr() == b + 4 * o for _, p, o in self.order)

    def grad_view(self, t):
        """The slice of G that mirrors `t` (a parameter, or a pack of adjacent parameters, living in P); None otherwise."""
        if t is None or t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
            return None
        a = t.data_ptr() - self._base
        n = t.numel()
        if a < 0 or a % 4 or a // 4 + n > self.numel:
            return None
        o = a // 4
        return self.G[o:o + n].view(t.shape)

    def offset_of(self, param):
        return self.offsets[id(param)]

    def attach_grads(self):
        """Point `.grad` of every trainable parameter at its slice of G (done once; G is cleared, never re-allocated)."""
        for _, p, o in self.order:
            p.grad = self.G[o:o + p.numel()].view(p.shape) if p.requires_grad else None

    def trainable_ranges(self):
        """Maximal [lo, hi) ranges of P (in floats, 4-aligned) that hold trainable parameters only (padding included)."""
        ranges = []
        for _, p, o in sorted(self.order, key=lambda e: e[2]):
            if not p.requires_grad:
                continue
            lo, hi = o, _align(o + p.numel())
            if ranges and lo <= ranges[-1][1]:
                ranges[-1][1] = max(ranges[-1][1], hi)
            else:
                ranges.append([lo - lo % 4, hi])
        # a frozen parameter packed right behind a trainable one must not be touched by the padded tail
        frozen = sorted((o, o + p.numel()) for _, p, o in self.order if not p.requires_grad)
        for lo, hi in ranges:
            for a, b in frozen:
                if a < hi and b > lo:
                    raise NotImplementedError('ParamStore: trainable and frozen parameters inside one pack group')
        return [(lo, hi) for lo, hi in ranges]

    @contextlib.contextmanager
    def direct_grads(self):
        """Inside this context the hand-written backward passes accumulate parameter gradients straight into G
        (and return None to autograd for them).  G must have been cleared for the step."""
        from . import functional as Fn
        old = Fn._sink
        Fn._sink = self
        try:
            yield self
        finally:
            Fn._sink = old
