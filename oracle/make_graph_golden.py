"""Dump the reference's adjacency stacks (graph/ucla.py, graph/ntu_rgb_d.py) to tests/golden/graphs.npz.
Build container only (needs /root/reference):  PYTHONDONTWRITEBYTECODE=1 python oracle/make_graph_golden.py"""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')
from graph import ucla, ntu_rgb_d  # noqa: E402  (reference)

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'graphs.npz')
np.savez_compressed(out, ucla=ucla.Graph('spatial').A, ntu=ntu_rgb_d.Graph('spatial').A)
print('wrote', out)
