#!/usr/bin/env python
"""kbench_wind.py -- timing of the elementwise wind kernels and the height conversions on one GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "earthkit-meteo_b200")]

import torch  # noqa: E402

from ek_thermo import vertical, wind  # noqa: E402

dev = "cuda:0"
n = 6599680 * 24
dtype = torch.float32 if "--f32" in sys.argv else torch.float64
esz = 4 if dtype == torch.float32 else 8
g = torch.Generator(device=dev).manual_seed(0)
u = torch.empty(n, device=dev, dtype=dtype).normal_(0.0, 12.0, generator=g)
v = torch.empty(n, device=dev, dtype=dtype).normal_(0.0, 12.0, generator=g)
t = torch.empty(n, device=dev, dtype=dtype).uniform_(200.0, 310.0, generator=g)
p = torch.empty(n, device=dev, dtype=dtype).uniform_(1.0e3, 1.05e5, generator=g)
lat = torch.empty(n, device=dev, dtype=dtype).uniform_(-90.0, 90.0, generator=g)
z = torch.empty(n, device=dev, dtype=dtype).uniform_(-1.0e3, 6.0e5, generator=g)
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
kernels = {
    "wind.speed": (lambda: wind.speed(u, v), 3),
    "wind.direction": (lambda: wind.direction(u, v), 3),
    "wind.xy_to_polar": (lambda: wind.xy_to_polar(u, v), 4),
    "wind.polar_to_xy": (lambda: wind.polar_to_xy(t, lat), 4),
    "wind.w_from_omega": (lambda: wind.w_from_omega(u, t, p), 4),
    "wind.coriolis": (lambda: wind.coriolis(lat), 2),
    "vertical.geopotential_height_from_geopotential": (lambda: vertical.geopotential_height_from_geopotential(z), 2),
    "vertical.geometric_height_from_geopotential": (lambda: vertical.geometric_height_from_geopotential(z), 2),
}
print(f"dtype={dtype} n={n} peak={peak} GB/s")
for name, (fn, narr) in kernels.items():
    for _ in range(3):
        r = fn()
    del r
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r = fn()
        del r
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = narr * esz * n / ms / 1e6
    print(f"{name:48s} {ms:8.3f} ms {n / ms / 1e6:8.2f} Gpt/s {gbs:8.1f} GB/s frac={gbs / peak:.3f}")
