"""Adjacency construction helpers (semantics of reference graph/tools.py:10-43).

Vectorised numpy; float64 like the reference.  Convention (graph/tools.py:10-14):
an edge (i, j) sets A[j, i] = 1, i.e. column = source joint, row = destination joint.
"""
import numpy as np


def edges_to_matrix(edges, num_node):
    A = np.zeros((num_node, num_node), dtype=np.float64)
    if len(edges):
        e = np.asarray(edges, dtype=np.int64)
        A[e[:, 1], e[:, 0]] = 1.0
    return A


def column_normalise(A):
    """A @ diag(1/colsum) with empty columns left at zero (graph/tools.py:27-35)."""
    deg = A.sum(axis=0)
    inv = np.zeros_like(deg)
    nz = deg > 0
    inv[nz] = 1.0 / deg[nz]
    return A * inv[None, :]


def spatial_partition(num_node, self_link, inward, outward):
    """(3, V, V) stack [identity, normalised inward, normalised outward] (graph/tools.py:38-43)."""
    return np.stack((edges_to_matrix(self_link, num_node),
                     column_normalise(edges_to_matrix(inward, num_node)),
                     column_normalise(edges_to_matrix(outward, num_node))))
