"""Zero-initialised scratch for one optimisation step, cleared with ONE memset.

The hand-written backward accumulates into ~270 small zero-initialised buffers per step (fp64 BatchNorm sums, fp32
weight-gradient accumulators).  Served by `torch.zeros` each of them is its own fill kernel on the step's critical
path; here they are consecutive slices of one buffer that `begin_step()` clears once.

Contract: a tensor handed out by `zeros()` is valid until the next `begin_step()` on the same device (parameter
gradients of step i are therefore valid until step i+1 starts — `engine.Trainer` owns that cadence).  Outside a
`begin_step()` / `end_step()` bracket, and whenever the buffer is too small, `zeros()` is plain `torch.zeros`.
"""
import torch

_ALIGN = 256


class ZeroArena:
    def __init__(self, device):
        self.device = device
        self.buf = None
        self.off = 0            # bytes handed out since begin_step()
        self.req = 0            # bytes requested since begin_step() (served or not)
        self.need = 0           # largest `req` seen in any step
        self.active = False

    def begin_step(self):
        self.need = max(self.need, self.req)
        if (self.buf is None or self.buf.numel() < self.need) and self.need > 0 \
                and not (self.device.type == 'cuda' and torch.cuda.is_current_stream_capturing()):
            self.buf = torch.empty(self.need, dtype=torch.uint8, device=self.device)
        if self.buf is not None:
            self.buf.zero_()
        self.off = self.req = 0
        self.active = True

    def end_step(self):
        self.need = max(self.need, self.req)
        self.active = False

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


_arenas = {}


def arena(device):
    device = torch.device(device)
    if device.type == 'cuda' and device.index is None:
        device = torch.device(device.type, torch.cuda.current_device())
    a = _arenas.get(device)
    if a is None:
        a = _arenas[device] = ZeroArena(device)
    return a


def zeros(shape, dtype, device):
    return arena(device).zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype)
