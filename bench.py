#!/usr/bin/env python
"""bench.py -- thermo grid-points/s and achieved HBM GB/s vs roofline (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch of synthetic input.  Default workload (N=1) is
BASELINE.json configs[1]: the fused thermo suite (t,q,p -> theta, es, rh, td, Tv) on IFS O1280
(6 599 680 points) x 137 model levels, float64 = 904 156 160 points, 64 algorithmic bytes per point.
Inputs are generated ON DEVICE before the timed region (58 GB resident, far larger than the 126 MB L2,
so every step streams from HBM).  For N>1 (torchrun, one process per GPU) every rank runs the same
per-GPU field (weak scaling); there is no data-path collective -- NCCL is only used for the barrier
and the max-over-ranks of the device time.

Printed JSON line: see the keys below; `roofline`, `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`
are described in DESIGN.md.  `--impl reference` times the CPU implementation (the numpy oracle port of
the reference, all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "earthkit-meteo_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

O1280_POINTS = 4 * 1280 * 1289  # 6 599 680 (octahedral reduced Gaussian grid O1280)
N_LEVELS = 137
METRIC = "thermo grid-points/s"  # BASELINE.json metric; the achieved HBM GB/s vs roofline half of it is the `roofline` object

WORKLOADS = {
    # name: (kind, outputs, levels, points per level, dtype)
    "suite_tqp_o1280x137_f64": ("tqp", ("theta", "es", "rh", "td", "tv"), N_LEVELS, O1280_POINTS, "f64"),
    "suite_tqp_o1280x137_f32": ("tqp", ("theta", "es", "rh", "td", "tv"), N_LEVELS, O1280_POINTS, "f32"),
    "suite_ttdp_o1280x137_f64": ("ttdp", ("theta", "es", "rh", "q", "tv"), N_LEVELS, O1280_POINTS, "f64"),
    "theta_rh_era5_f64": ("tqp", ("theta", "rh"), 1, 721 * 1440, "f64"),
    "ept_wbpt_o1280x137_f64": ("ept", ("ept", "wbpt"), N_LEVELS, O1280_POINTS, "f64"),
    "ept_wbpt_o1280x137_f32": ("ept", ("ept", "wbpt"), N_LEVELS, O1280_POINTS, "f32"),
    # BASELINE.json configs[3]: one GPU's shard (of 8) of ENS 51 members x O640 x 137 levels, humidity conversions
    # the same suite with the pressure computed in-kernel from sp and the L137 A/B (56 B/pt)
    "suite_tq_hybrid_o1280x137_f64": ("hybrid", ("theta", "es", "rh", "td", "tv"), N_LEVELS, O1280_POINTS, "f64"),
    "conv_ens_o640_shard_f64": ("tqp", ("rh", "td", "w"), 874, 4 * 640 * 649, "f64"),
}
DEFAULT_WORKLOAD = "suite_tqp_o1280x137_f64"


L2_FLUSH_BYTES = 384 << 20  # rotate small workloads over at least this many bytes (3 x the 126 MB L2)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload)
    return None


# --------------------------------------------------------------------------------------------------
# synthetic IFS-shaped input; generated on the device, never copied from the host
# --------------------------------------------------------------------------------------------------
def make_inputs_device(kind, levels, npl, dtype, device, seed):
    import torch

    from ek_thermo import thermo

    g = torch.Generator(device=device).manual_seed(seed)
    ab = np.load(os.path.join(ROOT, "tests", "golden", "ifs_l137_ab.npz"))
    a_full = 0.5 * (ab["A"][:-1] + ab["A"][1:])
    b_full = 0.5 * (ab["B"][:-1] + ab["B"][1:])
    if levels < a_full.size:  # short workloads use the lowest `levels` model levels
        a_full, b_full = a_full[-levels:], b_full[-levels:]
    elif levels > a_full.size:  # member x level slabs of an ensemble: the 137 levels repeat
        sel = np.arange(levels) % a_full.size
        a_full, b_full = a_full[sel], b_full[sel]
    a_l = torch.tensor(a_full, dtype=torch.float64, device=device).reshape(levels, 1)
    b_l = torch.tensor(b_full, dtype=torch.float64, device=device).reshape(levels, 1)
    sp = torch.empty(1, npl, dtype=torch.float64, device=device).uniform_(5.0e4, 1.05e5, generator=g)
    p = torch.empty(levels, npl, dtype=torch.float64, device=device)
    t = torch.empty(levels, npl, dtype=torch.float64, device=device)
    sp_keep = sp[0].clone() if kind == "hybrid" else None
    if kind == "hybrid":  # exactly the reference's full-level pressure: ph0 + 0.5 * (ph1 - ph0)  (vertical.py:663,708)
        a_h64 = torch.tensor(ab["A"], dtype=torch.float64, device=device)
        b_h64 = torch.tensor(ab["B"], dtype=torch.float64, device=device)
    for k in range(levels):  # level by level to bound temporaries
        if kind == "hybrid":
            ph0, ph1 = a_h64[k] + b_h64[k] * sp[0], a_h64[k + 1] + b_h64[k + 1] * sp[0]
            p[k] = ph0 + 0.5 * (ph1 - ph0)
        else:
            p[k] = a_l[k] + b_l[k] * sp[0]
        noise = torch.empty(npl, dtype=torch.float64, device=device).uniform_(-15.0, 15.0, generator=g)
        t[k] = (288.15 * (p[k] / 101325.0) ** 0.19 + noise).clamp_(180.0, 320.0)
    del sp
    h = torch.empty(levels, npl, dtype=torch.float64, device=device)
    for k in range(levels):
        if kind == "ttdp":
            h[k] = t[k] - torch.empty(npl, dtype=torch.float64, device=device).uniform_(0.0, 30.0, generator=g)
        else:
            u = torch.empty(npl, dtype=torch.float64, device=device).uniform_(1.0e-6, 0.02, generator=g)
            qs = thermo.saturation_specific_humidity(t[k], p[k])
            h[k] = torch.where(torch.isnan(qs), u, torch.minimum(u, 0.95 * qs.abs()))
    tdt = torch.float64 if dtype == "f64" else torch.float32
    if kind == "hybrid":  # hand the kernel sp and the half-level coefficients instead of the pressure field
        a_h = torch.tensor(ab["A"], dtype=tdt, device=device)
        b_h = torch.tensor(ab["B"], dtype=tdt, device=device)
        del p
        return [t.reshape(-1).to(tdt), h.reshape(-1).to(tdt), (sp_keep.reshape(-1).to(tdt), a_h, b_h)]
    return [x.reshape(-1).to(tdt) for x in (t, h, p)]


def make_inputs_host(kind, n, dtype, seed):
    """Same distribution for the CPU legs (numpy)."""
    from cases import random_inputs

    inp = random_inputs(n, seed=seed)
    npd = np.float64 if dtype == "f64" else np.float32
    h = inp["td"] if kind == "ttdp" else inp["q"]
    return [np.ascontiguousarray(x.astype(npd)) for x in (inp["t"], h, inp["p"])]


# --------------------------------------------------------------------------------------------------
# the hot-path step on the device
# --------------------------------------------------------------------------------------------------
def build_step(kind, outputs, arrays):
    """Returns (step callable, bytes per point).  Output buffers are allocated once and reused."""
    import torch

    from ek_thermo import _backend, fused

    t, h, p = arrays
    esz = t.element_size()
    if kind == "hybrid":  # p is (sp, A, B): [npl] + 2 x (nlev + 1)
        sp, a, b = p
        nlev = a.numel() - 1
        t2, h2 = t.reshape(nlev, -1), h.reshape(nlev, -1)
        out = {name: torch.empty_like(t2) for name in outputs}

        def step():
            fused.suite_tq_hybrid(t2, h2, sp, a, b, outputs=outputs, out=out)

        # per point: t, q read, outputs written; sp is read once per point COLUMN (8/nlev B per point)
        return step, esz * (2 + len(outputs)) + esz / nlev, {k: v.reshape(-1) for k, v in out.items()}
    if kind in ("tqp", "ttdp"):
        out = {name: torch.empty_like(t) for name in outputs}
        fn = fused.suite_tqp if kind == "tqp" else fused.suite_ttdp
        bpp = esz * (3 + len(outputs))
        # a field that fits the 126 MB L2 (one ERA5 level: 41.5 MB) is rotated over enough copies that no launch finds
        # its inputs in L2 ("rotating buffers > 126 MB")
        n_sets = 1 if bpp * t.numel() >= L2_FLUSH_BYTES else -(-L2_FLUSH_BYTES // (bpp * t.numel()))
        sets = [(t, h, p, out)] + [(t.clone(), h.clone(), p.clone(), {k: torch.empty_like(t) for k in outputs}) for _ in range(n_sets - 1)]
        state = {"i": 0}

        def step():
            a, b, c, o = sets[state["i"] % n_sets]
            state["i"] += 1
            fn(a, b, c, outputs=outputs, out=o)

        step.n_sets = n_sets
        return step, bpp, out
    # ept + wet-bulb potential temperature ("direct"), BASELINE.json configs[2]
    from ctypes import c_int, c_int64, c_void_p

    ept, wb = torch.empty_like(t), torch.empty_like(t)
    ops = [_backend.ek_operand(x.data_ptr(), 0.0) for x in (t, h, p)]
    c_args = ops + [c_int(1), c_int(0), c_int(1), c_int(1), c_void_p(ept.data_ptr()), c_void_p(wb.data_ptr()), c_int64(t.numel())]

    def step():
        _backend._call("ept_wet_bulb", t.dtype, t.device, c_args)

    return step, esz * 5, {"ept": ept, "wbpt": wb}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; only samples that arrive inside the timed
    region [mark_begin, mark_end] are reported."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for ts, ln in self.lines if self.t0 <= ts <= self.t1 + 0.06]
        if not inside:  # a timed region shorter than one sampling period: take the sample closest to it
            inside = [min(self.lines, key=lambda x: abs(x[0] - self.t1))[1]] if self.lines else []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_single_core(kind, outputs, npl, dtype, budget_s=12.0):
    """The oracle (numpy port of the reference, 1 core as the reference ships) on whole level slabs until ~budget_s."""
    import thermo_oracle as oracle

    t, h, p = make_inputs_host("ttdp" if kind == "ttdp" else "tqp", npl, dtype, seed=0)
    fn = _oracle_step(kind, outputs, oracle)
    with np.errstate(all="ignore"):
        fn(t[:10000], h[:10000], p[:10000])
        done, t0 = 0, time.perf_counter()
        while True:
            fn(t, h, p)
            done += 1
            el = time.perf_counter() - t0
            if el > budget_s or done >= 8:
                break
    return {"value": done * npl / el, "unit": "grid-points/s", "cores": 1, "kind": "port",
            "sample": f"{done} level slab(s) of {npl} points ({dtype}), numpy oracle of the reference, {el:.1f} s"}


def _oracle_step(kind, outputs, oracle):
    if kind == "tqp":
        fns = {"theta": lambda t, q, p: oracle.potential_temperature(t, p), "es": lambda t, q, p: oracle.saturation_vapour_pressure(t),
               "rh": oracle.relative_humidity_from_specific_humidity, "td": lambda t, q, p: oracle.dewpoint_from_specific_humidity(q, p),
               "tv": lambda t, q, p: oracle.virtual_temperature(t, q)}
    elif kind == "ttdp":
        fns = {"theta": lambda t, td, p: oracle.potential_temperature(t, p), "es": lambda t, td, p: oracle.saturation_vapour_pressure(t),
               "rh": lambda t, td, p: oracle.relative_humidity_from_dewpoint(t, td), "q": lambda t, td, p: oracle.specific_humidity_from_dewpoint(td, p),
               "tv": lambda t, td, p: oracle.virtual_temperature(t, oracle.specific_humidity_from_dewpoint(td, p))}
    else:
        fns = {"ept": oracle.ept_from_specific_humidity, "wbpt": oracle.wet_bulb_potential_temperature_from_specific_humidity}

    def step(a, b, c):
        return [fns[o](a, b, c) for o in outputs]

    return step


# --------------------------------------------------------------------------------------------------
# reference arm: CPU implementation on all host cores
# --------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(kind, outputs, dtype, npl, seed):
    import thermo_oracle as oracle

    _W["fn"] = _oracle_step(kind, outputs, oracle)
    _W["arr"] = make_inputs_host(kind, npl, dtype, seed)


def _worker_run(bounds):
    b, e = bounds
    with np.errstate(all="ignore"):
        res = _W["fn"](*(x[b:e] for x in _W["arr"]))
    return float(sum(np.nansum(r[:8]) for r in res))  # results stay in the worker (as they would stay in RAM)


def run_reference(args, kind, outputs, levels, npl, dtype):
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_step = min(npl, 6_599_680)  # one level slab per step (bounded sample of the workload)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_worker_init, initargs=(kind, outputs, dtype, n_step, 0)) as pool:
        edges = np.linspace(0, n_step, cores * 4 + 1).astype(np.int64)
        chunks = list(zip(edges[:-1], edges[1:]))
        for _ in range(max(args.warmup, 1)):
            pool.map(_worker_run, chunks)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_worker_run, chunks)
        el = time.perf_counter() - t0
    value = args.steps * n_step / el
    sample = f"{n_step} points per step (one O1280 level slab of the {levels}-level workload), numpy oracle port of the reference, {cores} processes"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "grid-points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": args.workload, "outputs": list(outputs), "points_per_step": int(n_step), "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "grid-points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "grid-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, kind, outputs, levels, npl, dtype):
    import torch

    import ek_thermo
    from ek_thermo import hostpipe

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_cpus = hostpipe.bind_host_to_device(device) if world > 1 else None  # pinned e2e buffers next to the rank's GPU
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)

    n = levels * npl
    arrays = make_inputs_device(kind, levels, npl, dtype, device, seed=rank)
    step, bytes_per_pt, out = build_step(kind, outputs, arrays)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ek_thermo.launch_count()
    barrier()
    sampler.mark_begin()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    sampler.mark_end()
    launches = ek_thermo.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        tt = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- e2e: the host-buffer call (pinned host arrays in, host arrays out), copies inside the timed region
    e2e_levels = max(1, min(levels, args.e2e_levels))
    n_e2e = e2e_levels * npl
    npd = np.float64 if dtype == "f64" else np.float32
    e2e = None
    if kind in ("tqp", "ttdp"):
        hs = hostpipe.HostSuite(device, workspace_bytes=args.e2e_workspace_mb << 20, n_slots=3)
        h_in = [hostpipe.pinned_empty(n_e2e, npd) for _ in range(3)]
        for hbuf, d in zip(h_in, arrays):
            hbuf[:] = d[:n_e2e].cpu().numpy()
        h_out = {name: hostpipe.pinned_empty(n_e2e, npd) for name in outputs}
        fn = hs.suite_tqp if kind == "tqp" else hs.suite_ttdp
        e2e_steps = max(2, min(args.steps, args.e2e_steps))
        fn(*h_in, outputs=outputs, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn(*h_in, outputs=outputs, out=h_out)  # blocks until the outputs are in host memory
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([el], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = float(tt.item())
        esz = np.dtype(npd).itemsize
        e2e = {"value": world * e2e_steps * n_e2e / el, "unit": "grid-points/s", "h2d_bytes_per_step": 3 * esz * n_e2e,
               "d2h_bytes_per_step": len(outputs) * esz * n_e2e, "steps": e2e_steps,
               "sample": f"{e2e_levels} of {levels} levels per step ({n_e2e} points) through ek_thermo.hostpipe.HostSuite, pinned host buffers"}
        # the device result of the timed steps and the host-pipeline result agree bit for bit on the shared slab
        name0 = outputs[0]
        assert np.array_equal(h_out[name0][:100000], out[name0][:100000].cpu().numpy(), equal_nan=True)
        del hs, h_in, h_out

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = bytes_per_pt * n / (ms_step * 1e-3) / 1e9  # per GPU: one kernel per step
        cpu = cpu_baseline_single_core("tqp" if kind == "hybrid" else kind, outputs, min(npl, O1280_POINTS), dtype) if world == 1 and not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": "grid-points/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": args.workload, "outputs": list(outputs), "levels": levels, "points_per_level": npl,
                       "points_per_gpu": n, "bytes_per_point": bytes_per_pt, "parallelism": f"shard x{world} (no collective)",
                       "host_cpus_rank0": (f"{len(numa_cpus)} CPUs local to the GPU" if numa_cpus else "unbound"),
                       "l2": ("inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed" % (bytes_per_pt * n / 1e9)
                              if getattr(step, "n_sets", 1) == 1 else
                              "inputs+outputs per step are %.1f MB (< 126 MB L2): steps rotate over %d buffer sets (%.0f MB), so no launch finds its inputs in L2"
                              % (bytes_per_pt * n / 1e6, step.n_sets, step.n_sets * bytes_per_pt * n / 1e6))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(args.workload), "peak_source": peak_src,
                         "kernel": {"ept": "ew_kernel<OpEptWb>", "tqp": "ew_kernel<OpSuiteTQPm>", "ttdp": "ew_kernel<OpSuiteTTdPm>",
                                    "hybrid": "suite_hybrid_kernel<OpSuiteTQPm>"}[kind],
                         "algorithmic_bytes_per_launch": bytes_per_pt * n, "avg_launch_ms": ms_step},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--levels", type=int, default=0, help="override the number of levels (smaller field)")
    ap.add_argument("--e2e-levels", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-workspace-mb", type=int, default=768)
    ap.add_argument("--no-cpu", action="store_true", help="skip the single-core CPU baseline leg")
    args = ap.parse_args()
    kind, outputs, levels, npl, dtype = WORKLOADS[args.workload]
    if args.levels:
        levels = args.levels
    if args.impl == "reference":
        run_reference(args, "tqp" if kind == "hybrid" else kind, outputs, levels, npl, dtype)
    else:
        run_ours(args, kind, outputs, levels, npl, dtype)


if __name__ == "__main__":
    main()
