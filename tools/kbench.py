#!/usr/bin/env python
"""kbench.py -- quick per-kernel timing on one GPU: achieved algorithmic GB/s and fraction of the measured HBM peak.

    EK_THERMO_LIB=libek_thermo_lean.so python tools/kbench.py [--points N] [--dtype f64|f32] [--ctas 8,16,32] [--only name,...]

Each kernel is timed with CUDA events over `--iters` back-to-back launches on inputs far larger than L2.
Used to A/B library variants and launch configurations; bench.py stays the contract benchmark.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "earthkit-meteo_b200"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402

import ek_thermo  # noqa: E402
from ek_thermo import fused, thermo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=6599680 * 24)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--ctas", default="16")
    ap.add_argument("--only", default="")
    ap.add_argument("--masked", type=float, default=0.0, help="fraction of points (contiguous blocks) set to NaN in every input")
    ap.add_argument("--graph", action="store_true", help="also time the fused suites replayed from a CUDA graph")
    ap.add_argument("--realistic", action="store_true", help="IFS-like smooth t(p) instead of uniform random t")
    ap.add_argument("--smooth", action="store_true",
                    help="spatially smooth fields (neighbouring points differ by ~1e-4 relative, as analysed fields do): warps stay "
                         "phase-uniform and read neighbouring table entries; the default inputs are random per point (SURVEY 8(d))")
    a = ap.parse_args()
    dt = torch.float64 if a.dtype == "f64" else torch.float32
    esz = 8 if a.dtype == "f64" else 4
    dev = "cuda:0"
    n = a.points
    g = torch.Generator(device=dev).manual_seed(0)
    p = torch.empty(n, device=dev, dtype=torch.float64).uniform_(1.0e3, 1.05e5, generator=g)
    if a.smooth:  # long waves along the array + 0.01 % noise
        x = torch.arange(n, device=dev, dtype=torch.float64) * (2.0 * 3.141592653589793 / n)
        wob = lambda lo, hi: 1.0 + torch.empty(n, device=dev, dtype=torch.float64).uniform_(lo, hi, generator=g)  # noqa: E731
        p = (5.3e4 + 5.0e4 * torch.sin(7.0 * x)) * wob(-1e-4, 1e-4)
        t = (288.15 * (p / 101325.0) ** 0.19 + 10.0 * torch.sin(131.0 * x)).clamp_(180, 320) * wob(-1e-4, 1e-4)
        del x
    elif a.realistic:
        t = (288.15 * (p / 101325.0) ** 0.19 + torch.empty(n, device=dev, dtype=torch.float64).uniform_(-15, 15, generator=g)).clamp_(180, 320)
    else:
        t = torch.empty(n, device=dev, dtype=torch.float64).uniform_(200.0, 320.0, generator=g)
    if a.smooth:
        x = torch.arange(n, device=dev, dtype=torch.float64) * (2.0 * 3.141592653589793 / n)
        q = (0.0101 + 0.0099 * torch.sin(53.0 * x)) * wob(-1e-4, 1e-4)
        td = t - (15.0 + 14.0 * torch.sin(29.0 * x))
        r = (50.5 + 49.0 * torch.sin(17.0 * x)) * wob(-1e-4, 1e-4)
        del x
    else:
        q = torch.empty(n, device=dev, dtype=torch.float64).uniform_(1.0e-6, 0.02, generator=g)
        td = t - torch.empty(n, device=dev, dtype=torch.float64).uniform_(0.0, 30.0, generator=g)
        r = torch.empty(n, device=dev, dtype=torch.float64).uniform_(1.0, 100.0, generator=g)
    if a.masked > 0:  # missing values in contiguous blocks of 64 Ki points
        m = (torch.arange(n, device=dev) // 65536) % 100 < int(a.masked * 100)
        for x in (t, p, q, td, r):
            x[m] = float("nan")
        del m
    t, p, q, td, r = (x.to(dt) for x in (t, p, q, td, r))
    out5 = {k: torch.empty_like(t) for k in ("theta", "es", "rh", "td", "tv", "q", "w", "e", "thetav", "ept", "wbpt")}

    kernels = {
        # name: (callable, arrays touched)
        "theta": (lambda: thermo.potential_temperature(t, p), 3),
        "es_mixed": (lambda: thermo.saturation_vapour_pressure(t), 2),
        "rh_from_q": (lambda: thermo.relative_humidity_from_specific_humidity(t, q, p), 4),
        "td_from_q": (lambda: thermo.dewpoint_from_specific_humidity(q, p), 3),
        "q_from_td": (lambda: thermo.specific_humidity_from_dewpoint(td, p), 3),
        "rh_from_td": (lambda: thermo.relative_humidity_from_dewpoint(t, td), 3),
        "td_from_rh": (lambda: thermo.dewpoint_from_relative_humidity(t, r), 3),
        "tv": (lambda: thermo.virtual_temperature(t, q), 3),
        "suite_tqp5": (lambda: fused.suite_tqp(t, q, p, out=out5), 8),
        "suite_tqp_theta_rh": (lambda: fused.suite_tqp(t, q, p, outputs=("theta", "rh"), out=out5), 5),
        "suite_tqp_rh_td_w": (lambda: fused.suite_tqp(t, q, p, outputs=("rh", "td", "w"), out=out5), 6),
        "suite_ttdp5": (lambda: fused.suite_ttdp(t, td, p, out=out5), 8),
        "single_pass_tqp": (lambda: fused.suite_tqp(t, q, p, outputs=fused.SINGLE_PASS_TQP, out=out5), 8),
        "suite7_tqp": (lambda: fused.suite_tqp(t, q, p, outputs=fused.ALL7_TQP, out=out5), 10),
        "suite7_ttdp": (lambda: fused.suite_ttdp(t, td, p, outputs=fused.ALL7_TTDP, out=out5), 10),
        "suite_ept_wbpt": (lambda: fused.suite_tqp(t, q, p, outputs=("ept", "wbpt"), out=out5), 5),
        "ept_ifs_q": (lambda: thermo.ept_from_specific_humidity(t, q, p), 4),
        "ept_wbpt_direct": (lambda: fused.ept_wet_bulb(t, q, p, "q", "ifs", "direct"), 5),
        "wbpt_newton": (lambda: thermo.wet_bulb_potential_temperature_from_specific_humidity(t, q, p, t_method="newton"), 4),
        "wbpt_bisect": (lambda: thermo.wet_bulb_potential_temperature_from_specific_humidity(t, q, p, t_method="bisect"), 4),
        "lcl_davies": (lambda: thermo.lcl(t, td, p), 5),
    }
    only = [s for s in a.only.split(",") if s]
    peak = 6551.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(f"lib={ek_thermo._backend.LIB_PATH} dtype={a.dtype} n={n} realistic={a.realistic} smooth={a.smooth} peak={peak} GB/s")
    for ctas in [int(c) for c in a.ctas.split(",")]:
        ek_thermo.set_launch_config(0, ctas)
        for name, (fn, narr) in kernels.items():
            if only and name not in only:
                continue
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            gbs = narr * esz * n / ms / 1e6
            extra = ""
            if a.graph and name.startswith("suite"):  # the same call replayed from a CUDA graph (no host work per call)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    fn()
                for _ in range(3):
                    gr.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(a.iters):
                    gr.replay()
                e1.record()
                torch.cuda.synchronize()
                extra = f"  graph-replay {e0.elapsed_time(e1) / a.iters * 1e3:7.2f} us/call"
            print(f"ctas/SM={ctas:3d} {name:22s} {ms:8.3f} ms  {n / ms / 1e6:8.2f} Gpt/s  {gbs:8.1f} GB/s  frac={gbs / peak:.3f}{extra}")


if __name__ == "__main__":
    main()
