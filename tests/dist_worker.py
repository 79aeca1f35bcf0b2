"""Worker of tests/test_partition.py::test_world_size_2_gloo: one process per rank (gloo, CPU).

Each rank computes ITS shard of a fused-suite call -- host logic of the package unchanged, per-point work done by
the mock device (tests/hostmath_backend.py) -- then the shards are gathered on rank 0 and compared with the
oracle on the whole field.  No collective is part of the data path; the gather exists only to check the result.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("earthkit-meteo_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import hostmath_backend  # noqa: E402
import thermo_oracle as oracle  # noqa: E402
from cases import random_inputs  # noqa: E402
from ek_thermo import _backend, fused, partition  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world, _ = partition.env_rank_world()
    assert world == dist.get_world_size() == 2 and rank == dist.get_rank()
    _backend._call = hostmath_backend.fake_call
    _backend._check_device = lambda tensors: tensors[0].device

    n, align = 100_003, 1024
    inp = random_inputs(n, seed=3)
    t, q, p = (torch.from_numpy(inp[k]) for k in ("t", "q", "p"))
    b, e = partition.shard_range(n, world, rank, align)
    mine = fused.suite_tqp(t[b:e], q[b:e], p[b:e])
    gathered = [None] * world
    dist.all_gather_object(gathered, {"range": (b, e), "out": {k: v.numpy() for k, v in mine.items()}})
    ok = True
    if rank == 0:
        assert [g["range"] for g in gathered] == partition.all_shards(n, world, align)
        with np.errstate(all="ignore"):
            want = oracle.suite_tqp(inp["t"], inp["q"], inp["p"])
        for name in fused.DEFAULT_TQP:
            got = np.concatenate([g["out"][name] for g in gathered])
            ok &= got.shape == (n,) and bool(np.allclose(got, want[name], rtol=1e-12, atol=0, equal_nan=True))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
