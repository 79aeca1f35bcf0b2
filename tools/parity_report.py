#!/usr/bin/env python
"""parity_report.py -- MEASURED error of every function of the path against the oracle, per case, dtype and input set.

    python tools/parity_report.py gpu  [--out gpurun_out/parity_gpu.json]      # on the B200 box
    python tools/parity_report.py cond [--in gpurun_out/parity_gpu.json] [--out profiles/r02_parity.json]   # anywhere (CPU)

Stage `gpu` runs every case of tests/cases.py (103 function x option combinations) plus the fused suites through the CUDA
path, float64 and float32, on two deterministic input sets -- "random" (tests/cases.random_inputs, the parity tests'
set) and "ifs" (tools/synthetic.ifs_point_inputs: the benchmark's IFS L137 column shape, top levels at 1-100 Pa where the
reference's NaN rule is live) -- and records, per case: n, NaN / inf position mismatches, max / p99 / p99.9 relative
difference, the number of points over the FLAT contract limit (1e-12 float64, 1e-5 float32) and those points' indices.
Stage `cond` needs no GPU: it regenerates the same inputs, evaluates the conditioning of the oracle (relative change under
+-1 / +-16 ulp input perturbations, tests/compare.conditioning) AT THE EXCEEDING POINTS ONLY, and reports how many of them
remain over max(limit, 4 x conditioning) -- i.e. how often the conditioning-aware rule of tests/compare.py is needed and
whether it ever fails.  The oracle is the checker here, never the thing measured.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("earthkit-meteo_b200", "oracle", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import numpy as np  # noqa: E402

N_RANDOM = 200_000
N_IFS_PER_LEVEL = 1500  # x 137 levels = 205 500 points
MAX_IDX = 4000


def input_sets():
    from cases import random_inputs
    from synthetic import ifs_point_inputs

    return {"random": random_inputs(N_RANDOM, seed=5), "ifs": ifs_point_inputs(N_IFS_PER_LEVEL, seed=3)}


def suite_cases():
    """(id, suite, ept_method, output name) for the fused kernels: each output is a 'case' of its own."""
    out = []
    for suite, names in (("tqp", ("theta", "es", "rh", "td", "tv", "w", "e", "thetav", "ept", "wbpt")),
                         ("ttdp", ("theta", "es", "rh", "q", "tv", "w", "e", "thetav", "ept", "wbpt"))):
        for em in ("ifs", "bolton35", "bolton39"):
            for name in names:
                if em != "ifs" and name not in ("ept", "wbpt"):
                    continue
                out.append((f"fused.suite_{suite}[{name};ept_method={em}]", suite, em, name))
    return out


def stats(got, want, limit):
    got = np.asarray(got).astype(np.float64)
    want = np.asarray(want).astype(np.float64)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    inf_any = (np.isinf(got) | np.isinf(want)) & ~(nan_g | nan_w)
    fin = np.isfinite(got) & np.isfinite(want)
    with np.errstate(all="ignore"):
        rel = np.where(fin, np.abs(got - want) / np.maximum(np.abs(want), 1e-300), 0.0)
    over = np.flatnonzero(rel > limit)
    r = rel[fin]
    return {
        "n": int(got.size), "n_nan": int(nan_w.sum()), "nan_mismatches": int(np.sum(nan_g != nan_w)),
        "inf_mismatches": int(np.sum(inf_any & (got != want))),
        "max_rel": float(r.max()) if r.size else 0.0,
        "p99": float(np.quantile(r, 0.99)) if r.size else 0.0, "p999": float(np.quantile(r, 0.999)) if r.size else 0.0,
        "n_over_limit": int(over.size), "over_idx": over[:MAX_IDX].tolist(), "over_rel": rel[over[:MAX_IDX]].tolist(),
    }


def stage_gpu(path, mock=False):
    import torch

    import thermo_oracle as oracle
    from cases import CASES
    from ek_thermo import fused, thermo

    dev = "cuda:0"
    if mock:  # script self-test without a GPU: the g++ build of the functors stands in for the device (tests/hostmath_backend.py)
        import hostmath_backend
        from ek_thermo import _backend

        _backend._call = hostmath_backend.fake_call
        _backend._check_device = lambda tensors: tensors[0].device
        _backend._on_device = lambda t: True
        dev = "cpu"
        global N_RANDOM, N_IFS_PER_LEVEL
        N_RANDOM, N_IFS_PER_LEVEL = 3000, 20
    sets = input_sets()
    report = {"limit": {"f64": 1e-12, "f32": 1e-5}, "sets": {k: int(v["t"].size) for k, v in sets.items()}, "cases": {}}
    for dname, npd in (("f64", np.float64), ("f32", np.float32)):
        limit = report["limit"][dname]
        for sname, inp in sets.items():
            a_np = {k: np.ascontiguousarray(v.astype(npd)) for k, v in inp.items()}
            a_dev = {k: torch.from_numpy(v).to(dev) for k, v in a_np.items()}
            for case in CASES:
                args = [a_np[a] for a in case.args]
                res = getattr(thermo, case.fn)(*[a_dev[a] for a in case.args], **case.kwargs)
                with np.errstate(all="ignore"):
                    want = getattr(oracle, case.fn)(*args, **case.kwargs)
                if not isinstance(res, tuple):
                    res, want = (res,), (want,)
                for k, (g, w) in enumerate(zip(res, want)):
                    key = f"{case.id}#{k}" if len(res) > 1 else case.id
                    report["cases"].setdefault(key, {"iterative": case.iterative})[f"{dname}/{sname}"] = stats(g.cpu().numpy(), w, limit)
            for sid, suite, em, name in suite_cases():
                names = ("t", "q", "p") if suite == "tqp" else ("t", "td", "p")
                fn = fused.suite_tqp if suite == "tqp" else fused.suite_ttdp
                g = fn(*[a_dev[a] for a in names], outputs=(name,), ept_method=em)[name]
                with np.errstate(all="ignore"):
                    w = (oracle.suite_tqp if suite == "tqp" else oracle.suite_ttdp)(*[a_np[a] for a in names], ept_method=em)[name]
                report["cases"].setdefault(sid, {"iterative": ""})[f"{dname}/{sname}"] = stats(g.cpu().numpy(), w, limit)
            print(f"[parity_report] {dname}/{sname} done", flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(report, f)
    print(f"[parity_report] wrote {path}")


def stage_cond(src, dst):
    import thermo_oracle as oracle
    from cases import CASE_BY_ID, Case
    from compare import conditioning

    with open(src) as f:
        report = json.load(f)
    sets = input_sets()
    summary = {"f64": {"cases": 0, "cases_with_points_over_flat_limit": 0, "points_over_flat_limit": 0, "points_over_after_conditioning": 0},
               "f32": {"cases": 0, "cases_with_points_over_flat_limit": 0, "points_over_flat_limit": 0, "points_over_after_conditioning": 0}}
    for cid, per in report["cases"].items():
        base_id = cid.split("#")[0]
        k_out = int(cid.split("#")[1]) if "#" in cid else 0
        for key, st in per.items():
            if key == "iterative":
                continue
            dname, sname = key.split("/")
            npd = np.float64 if dname == "f64" else np.float32
            limit = report["limit"][dname]
            summ = summary[dname]
            summ["cases"] += 1
            idx = np.asarray(st.pop("over_idx"), dtype=np.int64)
            rel = np.asarray(st.pop("over_rel"), dtype=np.float64)
            if idx.size == 1:  # the moist-adiabat functions need more than one point next to a scalar pressure
                idx, rel = np.repeat(idx, 2), np.repeat(rel, 2)
            st["n_over_after_conditioning"] = 0
            st["worst_ratio_to_conditioning"] = None
            if st["n_over_limit"]:
                summ["cases_with_points_over_flat_limit"] += 1
                summ["points_over_flat_limit"] += st["n_over_limit"]
                inp = sets[sname]
                if base_id in CASE_BY_ID:
                    case = CASE_BY_ID[base_id]
                    args = [np.ascontiguousarray(inp[a].astype(npd))[idx] for a in case.args]
                else:  # fused suite output: conditioning of the oracle's composition
                    suite = "tqp" if "suite_tqp" in base_id else "ttdp"
                    name = base_id.split("[")[1].split(";")[0]
                    em = base_id.split("ept_method=")[1].rstrip("]")
                    names = ("t", "q", "p") if suite == "tqp" else ("t", "td", "p")
                    fn_name = f"_suite_{suite}_{name}_{em}"
                    setattr(oracle, fn_name, (lambda suite, name, em: lambda a, b, c: (oracle.suite_tqp if suite == "tqp" else oracle.suite_ttdp)(a, b, c, ept_method=em)[name])(suite, name, em))
                    case = Case(fn_name, names)
                    args = [np.ascontiguousarray(inp[a].astype(npd))[idx] for a in names]
                if per.get("iterative") == "bisect":
                    st["note"] = "bisection: results are quantised to 0.0293 K and flip at sign ties (SURVEY 7.3-H3); judged by the bisect rule of tests/compare.py"
                    st["n_over_after_conditioning"] = None
                else:
                    cond = conditioning(case, args, k_out)
                    tol = np.maximum(limit, 4.0 * cond)
                    if dname == "f32":
                        # the reference's own float32 noise: its float32 result against its float64 result on the same
                        # (float32-valued) inputs.  Nobody can match the float32 oracle more closely than it matches itself.
                        fn = getattr(oracle, case.fn)
                        with np.errstate(all="ignore"):
                            w32 = fn(*args, **case.kwargs)
                            w64 = fn(*[a.astype(np.float64) for a in args], **case.kwargs)
                        w32 = np.asarray(w32[k_out] if isinstance(w32, tuple) else w32).astype(np.float64)
                        w64 = np.asarray(w64[k_out] if isinstance(w64, tuple) else w64).astype(np.float64)
                        with np.errstate(all="ignore"):
                            noise = np.abs(w32 - w64) / np.maximum(np.abs(w64), 1e-300)
                        noise = np.where(np.isfinite(noise), noise, np.inf)
                        st["n_over_after_conditioning_only"] = int((rel > tol).sum())
                        st["median_reference_f32_vs_f64_at_over_points"] = float(np.median(noise))
                        tol = np.maximum(tol, 4.0 * noise)
                    elif per.get("iterative") == "newton":
                        # float64 one-step Newton solve: the contract's bar is 1e-6 K (2e-9 relative = 6e-7 K at 300 K)
                        st["n_over_after_conditioning_only"] = int((rel > tol).sum())
                        tol = np.maximum(tol, 2e-9)
                    still = rel > tol
                    st["n_over_after_conditioning"] = int(still.sum())
                    st["sampled_over_points"] = int(idx.size)
                    with np.errstate(all="ignore"):
                        ratio = rel / np.maximum(cond, 1e-300)
                    st["worst_ratio_to_conditioning"] = float(np.min([np.max(ratio), 1e30]))
                    st["median_conditioning_at_over_points"] = float(np.median(cond))
                    summ["points_over_after_conditioning"] += int(still.sum())
    report["summary"] = summary
    report["how"] = ("stage gpu on a B200 (CUDA path through the C ABI vs oracle/thermo_oracle.py, identical inputs); stage cond on the CPU: conditioning = "
                     "relative change of the oracle under +-1 / +-16 ulp perturbations of each input, evaluated at the points over the flat limit")
    os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
    with open(dst, "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(summary, indent=1))
    # worst offenders table
    rows = []
    for cid, per in report["cases"].items():
        for key, st in per.items():
            if key != "iterative" and st["n_over_limit"]:
                rows.append((key, cid, st["n_over_limit"], st["n"], st["max_rel"], st["p999"], st["n_over_after_conditioning"]))
    for r in sorted(rows, key=lambda r: (r[0], -r[2]))[:400]:
        print(f"{r[0]:10s} over={r[2]:7d}/{r[3]}  max={r[4]:.2e} p99.9={r[5]:.2e} after_cond={r[6]}  {r[1]}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stage", choices=["gpu", "cond"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--in", dest="src", default=os.path.join(ROOT, "gpurun_out", "parity_gpu.json"))
    ap.add_argument("--mock", action="store_true", help="self-test of this script on the CPU mock device (no GPU, tiny inputs)")
    a = ap.parse_args()
    if a.stage == "gpu":
        stage_gpu(a.out or os.path.join(ROOT, "gpurun_out", "parity_gpu.json"), mock=a.mock)
    else:
        stage_cond(a.src, a.out or os.path.join(ROOT, "profiles", "r02_parity.json"))


if __name__ == "__main__":
    main()
