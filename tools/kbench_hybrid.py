#!/usr/bin/env python
"""kbench_hybrid.py -- timing of the hybrid-level kernels on O1280 x 137 levels, one GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "earthkit-meteo_b200"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ek_thermo import fused, vertical  # noqa: E402

dev = "cuda:0"
npl, nlev = 4 * 1280 * 1289, 137
dtype = torch.float64 if "--f32" not in sys.argv else torch.float32
esz = 8 if dtype == torch.float64 else 4
ab = np.load(os.path.join(ROOT, "tests", "golden", "ifs_l137_ab.npz"))
A, B = torch.tensor(ab["A"], dtype=dtype, device=dev), torch.tensor(ab["B"], dtype=dtype, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
sp = torch.empty(npl, device=dev, dtype=dtype).uniform_(5.0e4, 1.05e5, generator=g)
zs = torch.empty(npl, device=dev, dtype=dtype).uniform_(-300.0, 3.0e4, generator=g)
p = vertical.pressure_on_hybrid_levels(A, B, sp)
t = (288.15 * (p / 101325.0) ** 0.19 + torch.empty_like(p).uniform_(-15, 15, generator=g)).clamp_(180, 320)
q = torch.empty_like(p).uniform_(1.0e-6, 0.02, generator=g)
del p
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
n = npl * nlev
kernels = {
    # name: (callable, algorithmic bytes per [level, point] element)
    "pressure full": (lambda: vertical.pressure_on_hybrid_levels(A, B, sp), esz * (1 + 1 / nlev)),
    # delta and alpha are float64 arrays whatever the dtype of sp (as in the reference)
    "pressure full+half+delta+alpha": (lambda: vertical.pressure_on_hybrid_levels(A, B, sp, output=("full", "half", "delta", "alpha")), esz * (2 + 2 / nlev) + 16),
    "thickness (alpha/delta in registers)": (lambda: vertical.relative_geopotential_thickness_on_hybrid_levels(t, q, A, B, sp), esz * (3 + 1 / nlev)),
    "geometric height above ground": (lambda: vertical.height_on_hybrid_levels(t, q, zs, A, B, sp), esz * (3 + 2 / nlev)),
    "suite_tq_hybrid (5 outputs)": (lambda: fused.suite_tq_hybrid(t, q, sp, A, B), esz * (7 + 1 / nlev)),
}
print(f"dtype={dtype} points={n} peak={peak} GB/s")
for name, (fn, bpp) in kernels.items():
    for _ in range(2):
        r = fn()
    del r
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = fn()
        del r
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = bpp * n / ms / 1e6
    print(f"{name:40s} {ms:8.3f} ms {n / ms / 1e6:8.2f} Gpt/s {gbs:8.1f} GB/s frac={gbs / peak:.3f}")
