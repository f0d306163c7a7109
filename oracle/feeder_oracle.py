"""CPU restatement (numpy, float64) of the reference's NW-UCLA feeder arithmetic — TEST INFRASTRUCTURE.

Follows feeder/feeder_nucla_gcn.py:76-84 (rand_view_transform) and :85-130 (__getitem__, skeleton part); the random
draws are arguments, `draw_train` / `draw_val` reproduce the reference's own sequence of `random` calls (:92-95, :112-116).
Pinned against the imported reference Feeder by oracle/make_feeder_golden.py -> tests/golden/feeder_ucla.npz.
Only tests/ and the golden scripts import this file."""
import math
import random

import numpy as np

BONES = [(1, 2), (2, 3), (3, 3), (4, 3), (5, 3), (6, 5), (7, 6), (8, 7), (9, 3), (10, 9), (11, 10), (12, 11), (13, 1),
         (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19)]            # feeder_nucla_gcn.py:27-28


def view_matrix(agx, agy, s):
    """Ry . Rx . S, feeder_nucla_gcn.py:76-82."""
    agx, agy = math.radians(agx), math.radians(agy)
    Rx = np.asarray([[1, 0, 0], [0, math.cos(agx), math.sin(agx)], [0, -math.sin(agx), math.cos(agx)]])
    Ry = np.asarray([[math.cos(agy), 0, -math.sin(agy)], [0, 1, 0], [math.sin(agy), 0, math.cos(agy)]])
    Ss = np.asarray([[s, 0, 0], [0, s, 0], [0, 0, s]])
    return np.dot(Ry, np.dot(Rx, Ss))


def draw_train(length, time_steps=52, rng=random):
    """The reference's draws, in its order (:92-95 then :112-113)."""
    agx = rng.randint(-60, 60)
    agy = rng.randint(-60, 60)
    s = rng.uniform(0.5, 1.5)
    idx = rng.sample(list(np.arange(length)) * 100, time_steps)
    idx.sort()
    return agx, agy, s, [int(i) for i in idx]


def draw_val(length, time_steps=52):
    return 0, 0, 1.0, [int(i) for i in np.linspace(0, length - 1, time_steps).astype(int)]     # :97, :116


def skeleton_sample(value, agx, agy, s, idx, stream='joint', time_steps=52):
    """value: (L, 20, 3) -> (3, T, 20, 1) float64."""
    value = np.asarray(value, dtype=np.float64)
    value = value - value[0, 1, :]                                                        # :98-99
    X = np.dot(np.reshape(value, (-1, 3)), view_matrix(agx, agy, s))                       # :83, :100-101
    v_min, v_max = np.min(X, axis=0), np.max(X, axis=0)                                    # :102
    X = (X - v_min) / (v_max - v_min + 1e-6) * 2 - 1                                        # :103-104
    X = np.reshape(X, (-1, 20, 3))
    data = X[list(idx), :, :]                                                              # :107-117
    if stream == 'bone':                                                                   # :119-123
        out = np.zeros_like(data)
        for a, b in BONES:
            out[:, a - 1, :] = data[:, a - 1, :] - data[:, b - 1, :]
        data = out
    elif stream == 'motion':                                                               # :124-127
        out = np.zeros_like(data)
        out[:-1] = data[1:] - data[:-1]
        data = out
    return np.reshape(np.transpose(data, (2, 0, 1)), (3, time_steps, 20, 1))               # :129-130


def synthetic_sequences(seed=0, lengths=(16, 21, 52, 101, 201, 37)):
    """Seeded stand-ins for NW-UCLA skeleton sequences ((L, 20, 3) each): a random pose plus a random walk."""
    rng = np.random.RandomState(seed)
    seqs = []
    for L in lengths:
        base = rng.randn(1, 20, 3) * 0.3 + np.array([0.1, 0.9, 2.5])
        walk = np.cumsum(rng.randn(L, 20, 3) * 0.02, axis=0)
        seqs.append(base + walk)
    return seqs
