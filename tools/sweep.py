#!/usr/bin/env python
"""sweep.py -- field-size sweep (BASELINE.json configs[4]): grid-points/s and fraction of the HBM roofline vs N.

    python tools/sweep.py [--dtype f64] [--max-bytes 1.2e11] > gpurun_out/sweep.jsonl

Single process: per-GPU numbers.  Under torchrun (`python -m torch.distributed.run --nproc-per-node G tools/sweep.py`)
the field of N points is cut into G contiguous shards (ek_thermo.partition), every rank runs its shard on its own
GPU, and rank 0 prints the aggregate: N / max-over-ranks of the device time (NCCL only carries the barrier and that
max; there is no data-path collective).
Small N is launch-latency bound (Python call + ~5 us launch); the table shows where the roofline regime starts.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "earthkit-meteo_b200")]

import torch  # noqa: E402

from ek_thermo import fused, partition, thermo  # noqa: E402


L2_FLUSH = 1 << 30  # bytes a rotation must cover per GPU (8 x the 126 MB L2; with 3 x, a 120 MB float32 set still read 8.2 TB/s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--max-bytes", type=float, default=1.2e11)
    a = ap.parse_args()
    dt = torch.float64 if a.dtype == "f64" else torch.float32
    esz = 8 if a.dtype == "f64" else 4
    rank, world, local = partition.env_rank_world()
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(dev))
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    sizes = [1_000_000, 3_000_000, 10_000_000, 30_000_000, 100_000_000, 300_000_000, 1_000_000_000, 3_000_000_000, 10_000_000_000]
    g = torch.Generator(device=dev).manual_seed(0)
    for n_total in sizes:
        b, e = partition.shard_range(n_total, world, rank, align=4096)
        n = e - b
        kernels = {"theta": 3, "rh_from_q": 4, "suite_tqp5": 8}
        for name, narr in kernels.items():
            if narr * esz * (n_total // world + 4096) > a.max_bytes:
                if rank == 0:
                    print(json.dumps({"n": n_total, "gpus": world, "kernel": name, "dtype": a.dtype,
                                      "skipped": "exceeds the per-GPU memory cap; shard it over more GPUs (ek_thermo.partition)"}), flush=True)
                continue
            # a working set that would sit in the 126 MB L2 is rotated over enough buffer sets (>= 1 GiB in total per GPU) that
            # every launch streams from HBM: small-N cells are launch-latency numbers on the HBM axis, not L2 numbers
            set_bytes = narr * esz * max(n, 1)
            n_sets = 1 if set_bytes >= L2_FLUSH else min(256, -(-L2_FLUSH // set_bytes))
            sets = []
            for _ in range(n_sets):
                t = torch.empty(n, device=dev, dtype=dt).uniform_(200.0, 320.0, generator=g)
                p = torch.empty(n, device=dev, dtype=dt).uniform_(1.0e3, 1.05e5, generator=g)
                q = torch.empty(n, device=dev, dtype=dt).uniform_(1.0e-6, 0.02, generator=g)
                out = {k: torch.empty_like(t) for k in fused.DEFAULT_TQP} if name == "suite_tqp5" else None
                sets.append((t, p, q, out))
            state = {"i": 0}
            ring = [None] * n_sets  # the single-output functions allocate their result: keep the last n_sets alive, so that the
            # allocator hands out a different block each call (one re-used 40 MB block would sit in L2 and absorb every store)

            def fn(name=name, sets=sets, state=state, ring=ring):
                k = state["i"] % len(sets)
                t, p, q, out = sets[k]
                state["i"] += 1
                if name == "theta":
                    ring[k] = thermo.potential_temperature(t, p)
                elif name == "rh_from_q":
                    ring[k] = thermo.relative_humidity_from_specific_humidity(t, q, p)
                else:
                    fused.suite_tqp(t, q, p, out=out)

            iters = max(5, min(2000, int(2e9 / max(n, 1))))
            for _ in range(n_sets + 3):  # once around the ring: the allocator has every result block before the timed loop
                fn()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            if dist is not None:
                tt = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
            gbs = narr * esz * n_total / ms / 1e6
            if rank == 0:
                print(json.dumps({"n": n_total, "gpus": world, "kernel": name, "dtype": a.dtype, "ms": round(ms, 5),
                                  "gpts": round(n_total / ms / 1e6, 3), "gbs": round(gbs, 1),
                                  "frac_of_measured_hbm": round(gbs / (peak * world), 4), "buffer_sets": n_sets,
                                  "regime": ("HBM: one field per launch, larger than L2" if n_sets == 1 else
                                             f"HBM by rotation over {n_sets} buffer sets ({n_sets * set_bytes / 1e6:.0f} MB per GPU); launch latency bounds small N"),
                                  "iters": iters}), flush=True)
            del sets, fn
            torch.cuda.empty_cache()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
