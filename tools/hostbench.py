#!/usr/bin/env python
"""hostbench.py -- throughput of the host-array paths (numpy in, numpy out) on one GPU.

    python tools/hostbench.py [--points 67108864]

Times, on the same float64 field: ek_thermo.host.thermo.<fn> with pageable numpy arrays and with page-locked ones
(two streams, Python per chunk), and the C pipeline hostpipe.HostSuite (three streams) for the five-output suite.
Wall-clock, best of 3, inputs and outputs in host memory.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "earthkit-meteo_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ek_thermo import host, hostpipe  # noqa: E402


LAST = []


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del r
    LAST[:] = ts
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 26)
    a = ap.parse_args()
    n = a.points
    rng = np.random.default_rng(0)
    t = rng.uniform(200.0, 310.0, n)
    p = rng.uniform(1.0e3, 1.05e5, n)
    q = rng.uniform(1.0e-6, 0.02, n)
    tp, pp, qp = (hostpipe.pinned_empty(n) for _ in range(3))
    tp[:], pp[:], qp[:] = t, p, q
    rows = []
    for name, fn, bpp in (("potential_temperature", lambda x, y, z: host.thermo.potential_temperature(x, y), 24),
                          ("relative_humidity_from_specific_humidity", lambda x, y, z: host.thermo.relative_humidity_from_specific_humidity(x, z, y), 32),
                          ("wet_bulb_temperature_from_specific_humidity (bisect)", lambda x, y, z: host.thermo.wet_bulb_temperature_from_specific_humidity(x, z, y), 32)):
        for kind, arrs in (("pageable", (t, p, q)), ("pinned in, pageable out", (tp, pp, qp))):
            dt = best(lambda: fn(*arrs))
            rows.append((f"host.thermo.{name}", kind, n / dt / 1e9, bpp * n / dt / 1e9))
    from ek_thermo import fused

    for outputs in (("ept", "wbpt"), fused.DEFAULT_TQP, fused.ALL7_TQP):
        for pinned_results in (True, False):
            prev = host.set_pinned_results(pinned_results)
            for kind, arrs in (("pageable in", (t, q, p)), ("pinned in", (tp, qp, pp))):
                dt = best(lambda: host.fused.suite_tqp(*arrs, outputs=outputs), reps=4)
                rows.append((f"host.fused.suite_tqp ({len(outputs)} outputs)", f"{kind}, {'page-locked' if pinned_results else 'pageable'} results",
                             n / dt / 1e9, 8 * (3 + len(outputs)) * n / dt / 1e9))
                print(f"  reps [s] {outputs} pinned_results={pinned_results} {kind}: " + " ".join(f"{x:.3f}" for x in LAST), flush=True)
            host.set_pinned_results(prev)
    t0 = time.perf_counter()
    blocks = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(4)]
    t1 = time.perf_counter()
    del blocks
    blocks = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(4)]
    t2 = time.perf_counter()
    del blocks
    print(f"  page-locked allocation of 4 x {8 * n / 1e6:.0f} MB: first {t1 - t0:.3f} s, again (torch's caching host allocator) {t2 - t1:.3f} s")
    hs = hostpipe.HostSuite("cuda:0", workspace_bytes=1536 << 20, n_slots=3)
    outs = {k: hostpipe.pinned_empty(n) for k in ("theta", "es", "rh", "td", "tv")}
    dt = best(lambda: hs.suite_tqp(tp, qp, pp, outputs=tuple(outs), out=outs))
    rows.append(("hostpipe.HostSuite.suite_tqp (5 outputs)", "pinned in and out", n / dt / 1e9, 64 * n / dt / 1e9))
    print(f"points={n} float64")
    for r in rows:
        print(f"{r[0]:<72s} {r[1]:<26s} {r[2]:7.3f} Gpt/s  {r[3]:7.1f} GB/s over PCIe")


if __name__ == "__main__":
    main()
