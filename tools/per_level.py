#!/usr/bin/env python
"""per_level.py -- time the fused suite level by level on the bench's IFS-shaped field (which levels are slow, and why)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "earthkit-meteo_b200"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402

import bench  # noqa: E402
from ek_thermo import fused  # noqa: E402

dev = torch.device("cuda", 0)
npl = bench.O1280_POINTS
t, q, p = bench.IfsField("tqp", npl, levels=137, seed=0, device=dev).materialise(0, 137, torch.float64)
out = {k: torch.empty(npl, device=dev, dtype=torch.float64) for k in fused.DEFAULT_TQP}
print("level  p_mean[Pa]  t_mean[K]  band_frac  ms      Gpt/s  frac")
for k in range(0, 137, 4):
    sl = slice(k * npl, (k + 1) * npl)
    tt, qq, pp = t[sl], q[sl], p[sl]
    for _ in range(3):
        fused.suite_tqp(tt, qq, pp, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        fused.suite_tqp(tt, qq, pp, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    band = float(((tt > 250.16) & (tt < 273.16)).double().mean())
    print(f"{k:5d}  {float(pp.mean()):10.1f}  {float(tt.mean()):8.2f}  {band:8.3f}  {ms:.4f}  {npl / ms / 1e6:6.1f}  {64 * npl / ms / 1e6 / 6551:.3f}")
