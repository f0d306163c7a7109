"""GPU-side NW-UCLA feeder (feeder/feeder_nucla_gcn.py:85-130).

CPU tier: the numpy oracle against the golden outputs of the UNMODIFIED reference feeder (tests/golden/feeder_ucla.npz,
written by oracle/make_feeder_golden.py with bit-exact agreement), and the host logic of
tam_gcn_b200.feeder.GpuSkeletonFeeder (random draws, evaluation frame grid) over the emulated kernel.
GPU tier: the CUDA kernel through the C-ABI against the same golden outputs, joint / bone / motion, train / val."""
import os

import numpy as np
import pytest
import torch

import emu_ops
import helpers as H
from oracle import feeder_oracle as FO

GOLD = np.load(os.path.join(H.GOLDEN, 'feeder_ucla.npz'))
SEQS = FO.synthetic_sequences()
KEYS = [(st, sp, i) for st in ('joint', 'bone', 'motion') for sp in ('train', 'val') for i in range(len(SEQS))]


@pytest.mark.parametrize('stream,split,i', KEYS)
def test_oracle_matches_reference_feeder(stream, split, i):
    key = '%s_%s_%d' % (stream, split, i)
    agx, agy, sc = GOLD[key + '_view']
    out = FO.skeleton_sample(SEQS[i], agx, agy, sc, GOLD[key + '_idx'], stream)
    assert np.array_equal(out.astype(np.float32), GOLD[key])
    if split == 'val':
        assert FO.draw_val(SEQS[i].shape[0])[3] == list(GOLD[key + '_idx'])


def _feeder(device, stream):
    from tam_gcn_b200.feeder import GpuSkeletonFeeder
    return GpuSkeletonFeeder(SEQS, [i % 10 for i in range(len(SEQS))], device, stream=stream)


def _check_against_golden(f, stream, tol):
    n = len(SEQS)
    for split in ('train', 'val'):
        view = torch.tensor(np.stack([GOLD['%s_%s_%d_view' % (stream, split, i)] for i in range(n)]), dtype=torch.float32)
        idx = torch.tensor(np.stack([GOLD['%s_%s_%d_idx' % (stream, split, i)] for i in range(n)]), dtype=torch.int32)
        x, y = f.batch(list(range(n)), train=(split == 'train'), view=view.to(f.device), frame_idx=idx.to(f.device))
        ref = np.stack([GOLD['%s_%s_%d' % (stream, split, i)] for i in range(n)])
        assert x.shape == (n, 3, 52, 20, 1) and x.dtype == torch.float32
        assert float(np.abs(x.cpu().numpy() - ref).max()) < tol, (stream, split)
        assert y.tolist() == [i % 10 for i in range(n)]
    # evaluation path end to end: the frame grid is np.linspace(0, L - 1, 52).astype(int), no view transform
    x, _ = f.batch(list(range(n)), train=False)
    ref = np.stack([GOLD['%s_val_%d' % (stream, i)] for i in range(n)])
    assert float(np.abs(x.cpu().numpy() - ref).max()) < tol


@pytest.mark.parametrize('stream', ['joint', 'bone', 'motion'])
def test_feeder_host_logic_cpu(stream, monkeypatch):
    emu_ops.install(monkeypatch)
    f = _feeder('cpu', stream)
    _check_against_golden(f, stream, 1e-6)
    # training draws: angles integers in [-60, 60], scale in [0.5, 1.5), frames sorted and inside the sequence;
    # seeded -> reproducible
    g = torch.Generator().manual_seed(5)
    sample = torch.arange(len(SEQS))
    view, idx = f.draw(sample, True, g)
    assert view.shape == (len(SEQS), 3) and idx.shape == (len(SEQS), 52)
    assert torch.equal(view[:, :2], view[:, :2].round()) and view[:, :2].abs().max() <= 60
    assert (view[:, 2] >= 0.5).all() and (view[:, 2] < 1.5).all()
    assert (idx[:, 1:] >= idx[:, :-1]).all() and (idx >= 0).all()
    assert (idx.max(1).values < f.length).all()
    # a 16-frame sequence sampled 52 times from {0..15} x 100 uses every frame at most 100 times and most frames at least once
    assert idx[0].unique().numel() >= 12
    view2, idx2 = f.draw(sample, True, torch.Generator().manual_seed(5))
    assert torch.equal(view, view2) and torch.equal(idx, idx2)
    x, _ = f.batch(sample, train=True, generator=torch.Generator().manual_seed(6))
    assert x.abs().max() <= (1.0 if stream == 'joint' else 2.0) + 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('stream', ['joint', 'bone', 'motion'])
def test_feeder_kernel_matches_reference(stream):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    f = _feeder('cuda', stream)
    _check_against_golden(f, stream, 2e-5)          # fp32 arithmetic against the reference's float64 -> float32
    x, y = f.batch(torch.arange(len(SEQS)).repeat(11)[:64], train=True, generator=torch.Generator(device='cuda').manual_seed(1))
    assert x.shape == (64, 3, 52, 20, 1) and torch.isfinite(x).all()
