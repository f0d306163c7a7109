"""CTRGC kernel rooflines only (the three entries bench.py reports), for quick kernel iteration on the GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
peak, peak_tf, src = bench.measured_peaks()
dt = torch.bfloat16
which = sys.argv[1:] or ['ucla', 'bwd', 'ntu']
if 'ucla' in which:
    r = bench.ctrgc_roofline(dt, peak, src, 2048, 64, 52, 20, 3, 8); print('fwd V=20 ', round(r['achieved']), round(r['frac'], 4), round(r['ms_per_launch'], 4))
if 'bwd' in which:
    r = bench.ctrgc_roofline(dt, peak, src, 2048, 64, 52, 20, 3, 8, backward=True); print('bwd V=20 ', round(r['achieved']), round(r['frac'], 4), round(r['ms_per_launch'], 4))
if 'ntu' in which:
    r = bench.ctrgc_roofline(dt, peak, src, 2048, 64, 64, 25, 3, 8); print('fwd V=25 ', round(r['achieved']), round(r['frac'], 4), round(r['ms_per_launch'], 4))
if 'ntubwd' in which:
    r = bench.ctrgc_roofline(dt, peak, src, 1024, 64, 64, 25, 3, 8, backward=True); print('bwd V=25 ', round(r['achieved']), round(r['frac'], 4), round(r['ms_per_launch'], 4))
if 'ntu128' in which:
    r = bench.ctrgc_roofline(dt, peak, src, 1024, 128, 64, 25, 3, 8); print('fwd V=25 C=128 N=1024 ', round(r['achieved']), round(r['frac'], 4), round(r['ms_per_launch'], 4))
