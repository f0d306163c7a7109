"""Zero-initialised scratch for one optimisation step, cleared with ONE memset.

The hand-written backward accumulates into ~270 small zero-initialised buffers per step (fp64 BatchNorm sums, fp32
weight-gradient accumulators that are not served by a ParamStore).  Served by `torch.zeros` each of them is its own
fill kernel on the step's critical path; here they are consecutive slices of one buffer that `begin_step()` clears once.

An arena belongs to whoever drives the step (`engine.Trainer` owns one): `with arena.use(a): ...` makes `zeros()` serve
from `a` for tensors on its device.  Contract: a tensor handed out by `a.zeros()` is valid until the next
`a.begin_step()`.  Outside a `begin_step()` / `end_step()` bracket, without an active arena, and whenever the buffer is
too small, `zeros()` is plain `torch.zeros`.  After `freeze()` (a CUDA graph has captured the buffer's address) the
buffer is never re-allocated; overflow is served by `torch.zeros`.
"""
import contextlib

import torch

_ALIGN = 256


class ZeroArena:
    def __init__(self, device):
        self.device = torch.device(device)
        self.buf = None
        self.off = 0            # bytes handed out since begin_step()
        self.req = 0            # bytes requested since begin_step() (served or not)
        self.need = 0           # largest `req` seen in any step
        self.active = False
        self.frozen = False

    def begin_step(self):
        self.need = max(self.need, self.req)
        capturing = self.device.type == 'cuda' and torch.cuda.is_current_stream_capturing()
        if (self.buf is None or self.buf.numel() < self.need) and self.need > 0 and not self.frozen and not capturing:
            self.buf = torch.empty(self.need, dtype=torch.uint8, device=self.device)
        if self.buf is not None:
            self.buf.zero_()
        self.off = self.req = 0
        self.active = True

    def end_step(self):
        self.need = max(self.need, self.req)
        self.active = False

    def freeze(self):
        """Pin the current buffer: its address is baked into a captured graph."""
        self.frozen = True

    def zeros(self, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * dtype.itemsize
        padded = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        if self.active:
            self.req += padded
            if self.buf is not None and self.off + padded <= self.buf.numel() and nbytes > 0:
                t = self.buf[self.off:self.off + nbytes].view(dtype).view(shape)
                self.off += padded
                return t
        return torch.zeros(shape, dtype=dtype, device=self.device)


_current = None


@contextlib.contextmanager
def use(a):
    """Serve `zeros()` from arena `a` inside the context."""
    global _current
    old = _current
    _current = a
    try:
        yield a
    finally:
        _current = old


def zeros(shape, dtype, device):
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    a = _current
    if a is not None:
        device = torch.device(device)
        if device.type == 'cuda' and device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        ad = a.device if not (a.device.type == 'cuda' and a.device.index is None) else torch.device('cuda', torch.cuda.current_device())
        if ad == device:
            return a.zeros(shape, dtype)
    return torch.zeros(shape, dtype=dtype, device=device)
