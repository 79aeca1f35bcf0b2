#!/usr/bin/env python
"""dma_ceiling.py -- what the BOX can move between host memory and its GPUs: every rank (one per GPU, torchrun) copies
page-locked host buffers to its GPU and back at the same time, in the suite's 24 : 40 byte ratio and at 1 : 1, and rank 0 prints
the aggregate GB/s (bytes of all ranks / max-over-ranks wall time).  The host pipelines (bench.py `e2e`) are judged against
this number: at 1 GPU it is the PCIe link, at 8 GPUs of a VM it is the host's DMA / DRAM rate."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "earthkit-meteo_b200")]

import torch  # noqa: E402

from ek_thermo import hostpipe  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist

    hostpipe.bind_host_to_device(dev)
    dist.init_process_group("nccl", device_id=dev)
MB = 1 << 20
chunk = 64 * MB // 8  # doubles per copy
h_in = [torch.empty(chunk, dtype=torch.float64, pin_memory=True).fill_(1.0) for _ in range(3)]
h_out = [torch.empty(chunk, dtype=torch.float64, pin_memory=True) for _ in range(5)]
d_in = [torch.empty(chunk, dtype=torch.float64, device=dev) for _ in range(3)]
d_out = [torch.ones(chunk, dtype=torch.float64, device=dev) for _ in range(5)]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(n_in, n_out, reps=12):
    def once():
        with torch.cuda.stream(s_in):
            for k in range(n_in):
                d_in[k % 3].copy_(h_in[k % 3], non_blocking=True)
        with torch.cuda.stream(s_out):
            for k in range(n_out):
                h_out[k % 5].copy_(d_out[k % 5], non_blocking=True)

    once()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([el], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        el = float(tt.item())
    b_in, b_out = world * reps * n_in * chunk * 8, world * reps * n_out * chunk * 8
    return b_in / el / 1e9, b_out / el / 1e9


rows = [("H2D only", run(3, 0)), ("D2H only", run(0, 5)), ("H2D : D2H = 3 : 5 (the suite's 24 : 40 B/pt)", run(3, 5)), ("H2D : D2H = 1 : 1", run(5, 5))]
if rank == 0:
    print(f"box DMA ceiling, {world} GPU(s), 64 MB page-locked copies, both directions on their own streams")
    for name, (gi, go) in rows:
        print(f"  {name:48s} H2D {gi:7.1f} GB/s  D2H {go:7.1f} GB/s  total {gi + go:7.1f} GB/s")
    gi, go = rows[2][1]
    print(f"  => the five-output suite (24 B in + 40 B out per point) cannot exceed {(gi + go) / 64:.2f} Gpt/s on this box")
if dist is not None:
    dist.destroy_process_group()
