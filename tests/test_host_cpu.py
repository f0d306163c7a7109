"""Host-side plumbing that needs no GPU: the once-per-step zero arena and the flat parameter packs."""
import torch
import torch.nn as nn

from tam_gcn_b200 import arena
from tam_gcn_b200 import functional as Fn


def test_zero_arena_serves_slices_after_sizing_step():
    a = arena.ZeroArena(torch.device('cpu'))
    t = a.zeros((4,), torch.float32)                     # outside a step: plain zeros
    assert a.buf is None and t.eq(0).all()
    a.begin_step()                                       # sizing step: nothing to serve from yet
    x = a.zeros((3, 5), torch.float32)
    y = a.zeros((7,), torch.float64)
    assert a.buf is None and x.shape == (3, 5) and y.dtype == torch.float64
    a.end_step()
    a.begin_step()
    x = a.zeros((3, 5), torch.float32)
    y = a.zeros((7,), torch.float64)
    assert a.buf is not None
    lo, hi = a.buf.data_ptr(), a.buf.data_ptr() + a.buf.numel()
    assert lo <= x.data_ptr() < hi and lo <= y.data_ptr() < hi
    assert x.data_ptr() % 256 == lo % 256 and y.data_ptr() - x.data_ptr() == 256
    x.fill_(3.0)
    y.fill_(-1.0)
    z = a.zeros((1000,), torch.float32)                  # does not fit: falls back, and the arena grows next step
    assert not (lo <= z.data_ptr() < hi) and z.eq(0).all()
    a.end_step()
    a.begin_step()                                       # ONE clear for everything handed out before
    x2 = a.zeros((3, 5), torch.float32)
    y2 = a.zeros((7,), torch.float64)
    z2 = a.zeros((1000,), torch.float32)
    lo, hi = a.buf.data_ptr(), a.buf.data_ptr() + a.buf.numel()
    assert x2.eq(0).all() and y2.eq(0).all() and z2.eq(0).all() and lo <= z2.data_ptr() < hi
    a.end_step()
    assert not (lo <= a.zeros((2,), torch.float32).data_ptr() < hi)
    # module-level zeros() serves from the arena made current by use(); a frozen arena never re-allocates
    assert arena.zeros((3,), torch.float32, 'cpu').eq(0).all()
    a.freeze()
    with arena.use(a):
        a.begin_step()
        big = arena.zeros((1 << 16,), torch.float32, 'cpu')
        small = arena.zeros((5,), torch.float32, 'cpu')
        a.end_step()
        a.begin_step()
        assert a.buf.data_ptr() == lo and lo <= small.data_ptr() < hi and not (lo <= big.data_ptr() < hi)
        a.end_step()


def test_packed_parameters_are_views_of_one_buffer():
    torch.manual_seed(0)
    holder = nn.Module()
    convs = nn.ModuleList([nn.Conv2d(4, 6, 1) for _ in range(3)])
    ref = torch.cat([c.weight.detach().reshape(6, 4) for c in convs]).clone()
    keys = [k for k, _ in convs.state_dict().items()]
    W = Fn._packed(holder, 'W', [c.weight for c in convs], (18, 4))
    assert torch.equal(W, ref) and [k for k, _ in convs.state_dict().items()] == keys
    assert all(isinstance(c.weight, nn.Parameter) and c.weight.shape == (6, 4, 1, 1) for c in convs)
    # the parameters now alias the pack: an in-place optimiser update shows through, and no copy is made next time
    opt = torch.optim.SGD([c.weight for c in convs], lr=1.0)
    for c in convs:
        c.weight.grad = torch.ones_like(c.weight)
    opt.step()
    W2 = Fn._packed(holder, 'W', [c.weight for c in convs], (18, 4))
    assert W2.data_ptr() == W.data_ptr() and torch.equal(W2, ref - 1.0)
    # a storage swap (what .to() / load-by-assignment do) is detected and re-packed
    convs[1].weight.data = torch.full_like(convs[1].weight, 5.0)
    W3 = Fn._packed(holder, 'W', [c.weight for c in convs], (18, 4))
    assert torch.equal(W3[6:12], torch.full((6, 4), 5.0)) and torch.equal(W3[:6], ref[:6] - 1.0)
    assert convs[1].weight.data_ptr() == W3[6:12].data_ptr()
    # non-leaf tensors (DataParallel replicas) are concatenated, never re-pointed
    reps = [c.weight * 1.0 for c in convs]
    W4 = Fn._packed(nn.Module(), 'W', reps, (18, 4))
    assert torch.equal(W4, W3) and W4.data_ptr() != W3.data_ptr()
