#!/usr/bin/env python
"""kbench_levels.py -- the per-level caller (BASELINE.json configs[0]): theta + rh on ERA5 0.25 degree levels (721 x 1440 points)
held ONE ALLOCATION PER LEVEL.  Times, per level: one eager launch per level, the same launches replayed from a CUDA graph, and
one batched launch over all levels (fused.suite_tqp_batch).  Enough levels that nothing is served from L2 (137 x 41.5 MB)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "earthkit-meteo_b200"), os.path.join(ROOT, "tools")]

import torch  # noqa: E402

from ek_thermo import fused  # noqa: E402
from synthetic import IfsField  # noqa: E402

dev = "cuda:0"
npl, nlev = 721 * 1440, 137
dtype = torch.float32 if "--f32" in sys.argv else torch.float64
outputs = ("theta", "rh")
field = IfsField("tqp", npl, levels=nlev, seed=0, device=dev)
ts, qs, ps = [], [], []
for k in range(nlev):
    t, q, p = (x.to(dtype) for x in field.slab(k))
    ts.append(t), qs.append(q), ps.append(p)
outs = [{name: torch.empty_like(ts[0]) for name in outputs} for _ in range(nlev)]
bpp = ts[0].element_size() * (3 + len(outputs))
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def per_level():
    for k in range(nlev):
        fused.suite_tqp(ts[k], qs[k], ps[k], outputs=outputs, out=outs[k])


def batched():
    fused.suite_tqp_batch(ts, qs, ps, outputs=outputs, out=outs)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = [("one launch per level, eager", timed(per_level))]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    per_level()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        per_level()
torch.cuda.synchronize()
rows.append(("one launch per level, CUDA-graph replay", timed(g.replay)))
# the levels are independent: captured on three forked streams, neighbouring launches overlap (one kernel's launch latency and
# tail hide behind the next kernel's body) -- what a per-level caller with its own streams gets without the batched entry point
side = [torch.cuda.Stream() for _ in range(3)]


def per_level_streams():
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(main)
    for st in side:
        st.wait_event(ev)
    for k in range(nlev):
        with torch.cuda.stream(side[k % 3]):
            fused.suite_tqp(ts[k], qs[k], ps[k], outputs=outputs, out=outs[k])
    for st in side:
        e2 = torch.cuda.Event()
        e2.record(st)
        main.wait_event(e2)


with torch.cuda.stream(s):
    per_level_streams()
    g3 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g3):
        per_level_streams()
torch.cuda.synchronize()
rows.append(("per level on 3 streams, CUDA-graph replay", timed(g3.replay)))
rows.append(("one batched launch over all levels", timed(batched)))
print(f"theta + rh, {nlev} levels x {npl} points, {dtype}, {bpp} B/pt, roofline {bpp * npl / peak / 1e3:.2f} us per level")
for name, ms in rows:
    us = ms * 1e3 / nlev
    print(f"{name:42s} {us:7.2f} us per level  {npl / us / 1e3:7.1f} Gpt/s  frac={bpp * npl / us / 1e3 / peak:.3f}")

# pressure-level data (ERA5 on pressure levels): t and q fields, the pressure ONE NUMBER per level -- 2 arrays in, 2 out
plev = [float(p[0]) for p in ps]
bpp_l = ts[0].element_size() * (2 + len(outputs))


def per_level_scalar():
    for k in range(nlev):
        fused.suite_tqp(ts[k], qs[k], plev[k], outputs=outputs, out=outs[k])


def batched_scalar():
    fused.suite_tqp_batch(ts, qs, plev, outputs=outputs, out=outs)


rows = [("one launch per level, eager", timed(per_level_scalar)), ("one batched launch, one pressure per level", timed(batched_scalar))]
print(f"pressure levels: theta + rh, {nlev} levels x {npl} points, scalar p per level, {bpp_l} B/pt, roofline {bpp_l * npl / peak / 1e3:.2f} us per level")
for name, ms in rows:
    us = ms * 1e3 / nlev
    print(f"{name:42s} {us:7.2f} us per level  {npl / us / 1e3:7.1f} Gpt/s  frac={bpp_l * npl / us / 1e3 / peak:.3f}")
