"""Skeleton graphs (adjacency stacks) used by the CTR-GCN / ST-GCN hot path.

The reference builds these once at constructor time (graph/ucla.py:18-33,
graph/ntu_rgb_d.py:17-33, graph/tools.py:10-43).  `/root/reference` does not exist on
the GPU box, so the package carries its own builders; `tests/test_graph.py` checks them
element-for-element against golden arrays dumped from the reference.
"""
from . import tools, ucla, ntu_rgb_d  # noqa: F401
