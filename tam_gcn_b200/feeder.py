"""GPU-side NW-UCLA skeleton feeder — the work of feeder/feeder_nucla_gcn.py:85-130 for a whole batch
in one kernel launch, with the dataset resident in HBM (the whole NW-UCLA skeleton set is ~1500 sequences x <= 201 frames x
20 joints x 3 floats = a few MB).

    feeder = GpuSkeletonFeeder(sequences, labels, device, stream='joint')     # sequences: list of (L_i, 20, 3) arrays
    x, y = feeder.batch(indices, train=True, generator=g)                      # x: (B, 3, 52, 20, 1) fp32 on the device

The random draws follow the reference: view angles agx, agy ~ integers in [-60, 60], scale ~ U(0.5, 1.5)
(:92-95); training frames = 52 draws WITHOUT replacement from the multiset {every frame index x 100}, sorted (:112-113);
evaluation frames = linspace(0, L - 1, 52) truncated (:116).  `draw()` produces them with torch's generator on the
device; `batch(..., view=, frame_idx=)` accepts explicit draws (that is how the parity tests pin the arithmetic against
the reference feeder).
"""
import numpy as np
import torch

from . import ops

# feeder/feeder_nucla_gcn.py:27-28 — (joint, the joint subtracted from it), 1-based
UCLA_BONES = [(1, 2), (2, 3), (3, 3), (4, 3), (5, 3), (6, 5), (7, 6), (8, 7), (9, 3), (10, 9), (11, 10), (12, 11), (13, 1),
              (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19)]
MODES = {'joint': 0, 'bone': 1, 'motion': 2}


class GpuSkeletonFeeder:
    def __init__(self, sequences, labels, device, stream='joint', time_steps=52, bones=UCLA_BONES):
        if stream not in MODES:
            raise ValueError('stream must be one of %s' % sorted(MODES))
        self.device = torch.device(device)
        self.mode = MODES[stream]
        self.T = int(time_steps)
        S = len(sequences)
        self.V = int(np.asarray(sequences[0]).shape[1])
        Lmax = max(int(np.asarray(s).shape[0]) for s in sequences)
        raw = np.zeros((S, Lmax, self.V, 3), dtype=np.float32)
        length = np.zeros((S,), dtype=np.int32)
        for i, s in enumerate(sequences):
            s = np.asarray(s, dtype=np.float32)
            raw[i, :s.shape[0]] = s
            length[i] = s.shape[0]
        # evaluation frame grid of every sequence, computed once with numpy itself (feeder_nucla_gcn.py:116) — float64
        # rounding of i * (L - 1) / (T - 1) decides the truncation, and device arithmetic is free to round differently
        val_idx = np.stack([np.linspace(0, int(L) - 1, self.T).astype(int) for L in length]).astype(np.int32)
        self.val_idx = torch.from_numpy(val_idx).to(self.device)
        self.raw = torch.from_numpy(raw).to(self.device)
        self.length = torch.from_numpy(length).to(self.device)
        self.labels = torch.as_tensor(np.asarray(labels), dtype=torch.int64).to(self.device)
        parent = -np.ones((self.V,), dtype=np.int32)
        for a, b in bones:
            parent[a - 1] = b - 1
        self.bone_parent = torch.from_numpy(parent).to(self.device)

    def __len__(self):
        return self.raw.shape[0]

    def draw(self, sample, train, generator=None):
        """Random view transform and frame selection for the sequences `sample` (B,) -> view (B, 3), frame_idx (B, T)."""
        B = sample.numel()
        dev = self.device
        L = self.length[sample].to(torch.int64)                                        # (B,)
        if not train:
            view = torch.tensor([0.0, 0.0, 1.0], device=dev).repeat(B, 1)
            idx = self.val_idx[sample]
            return view, idx.contiguous()
        ang = torch.randint(-60, 61, (B, 2), device=dev, generator=generator).to(torch.float32)
        sc = torch.rand((B, 1), device=dev, generator=generator) + 0.5
        view = torch.cat([ang, sc], 1).contiguous()
        # T draws without replacement from {0..L-1} x 100: random keys over the 100 * Lmax slots, invalid slots pushed
        # to the end, the T smallest keys win; slot -> frame = slot % L
        Lmax = int(self.raw.shape[1])
        keys = torch.rand((B, 100 * Lmax), device=dev, generator=generator)
        slot = torch.arange(100 * Lmax, device=dev)[None, :]
        keys = torch.where(slot < 100 * L[:, None], keys, torch.full_like(keys, 2.0))
        pick = keys.topk(self.T, dim=1, largest=False).indices
        idx = (pick % L[:, None]).sort(dim=1).values.to(torch.int32)
        return view, idx.contiguous()

    def batch(self, indices, train=True, generator=None, view=None, frame_idx=None):
        sample = torch.as_tensor(indices, dtype=torch.int64, device=self.device).contiguous()
        if view is None or frame_idx is None:
            view, frame_idx = self.draw(sample, train, generator)
        B = sample.numel()
        out = torch.empty((B, 3, self.T, self.V, 1), device=self.device, dtype=torch.float32)
        ops.feeder_nucla(self.raw, self.length, sample, view.to(torch.float32).contiguous(),
                         frame_idx.to(torch.int32).contiguous(), self.bone_parent if self.mode == 1 else None, self.mode, out)
        return out, self.labels[sample]
