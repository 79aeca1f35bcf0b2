#!/usr/bin/env python
"""ab.py -- A/B of library variants on one GPU, robust to the box's power-state drift: the variants are timed round-robin
`--rounds` times (each a fresh kbench process, one library per process) and the median / best fraction per kernel is reported.

    python tools/ab.py --libs libek_thermo.so,libek_thermo_x.so --only suite_tqp5,ept_wbpt_direct [--rounds 3] [--args "--realistic"]
"""
import argparse
import collections
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--libs", required=True)
    ap.add_argument("--only", default="")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--args", default="--realistic")
    ap.add_argument("--tool", default="kbench.py")
    a = ap.parse_args()
    libs = a.libs.split(",")
    res = collections.defaultdict(list)
    for r in range(a.rounds):
        for lib in (libs if r % 2 == 0 else libs[::-1]):
            env = dict(os.environ, EK_THERMO_LIB=lib)
            cmd = [sys.executable, os.path.join(ROOT, "tools", a.tool)] + (["--only", a.only] if a.only else []) + a.args.split()
            out = subprocess.run(cmd, env=env, capture_output=True, text=True)
            for ln in out.stdout.splitlines():
                m = re.search(r"^(?:ctas/SM=\s*\d+\s+)?(.+?)\s+([\d.]+) ms\s+.*frac=([\d.]+)", ln)
                if m:
                    res[(m.group(1).strip(), lib)].append(float(m.group(3)))
            if out.returncode != 0:
                print(out.stderr[-2000:])
    kernels = sorted({k for k, _ in res})
    print(f"{'kernel':44s} " + " ".join(f"{lib.replace('libek_thermo', '').replace('.so', '') or '(default)':>22s}" for lib in libs))
    for k in kernels:
        cells = []
        for lib in libs:
            v = res.get((k, lib), [])
            cells.append(f"med {statistics.median(v):.3f} max {max(v):.3f}" if v else "-")
        print(f"{k:44s} " + " ".join(f"{c:>22s}" for c in cells))


if __name__ == "__main__":
    main()
