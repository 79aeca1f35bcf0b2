"""CPU tests of the boundary itself: exported symbols, the field partitioner, and a world-size-2 gloo run."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ek_thermo.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = []
    for m in re.finditer(r"EK_THERMO_FN\(\s*(\w+)\s*,", hdr):
        if m.group(1) != "name":
            names += [f"ek_thermo_{m.group(1)}_f64", f"ek_thermo_{m.group(1)}_f32"]
    names += re.findall(r"^\s*(?:int|uint64_t|const char\*)\s+(ek_thermo_\w+)\s*\(", hdr, flags=re.M)
    return sorted(set(names))


@pytest.mark.parametrize("lib", ["libek_thermo.so", "libek_thermo_exact.so"])
def test_library_loads_and_exports_every_declared_symbol(lib):
    path = os.path.join(ROOT, "earthkit-meteo_b200", "ek_thermo", lib)
    assert os.path.exists(path), "run __graft_entry__.build() first"
    so = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 2 * 43 + 5
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, missing
    so.ek_thermo_version.restype = ctypes.c_int
    assert so.ek_thermo_version() == int(re.search(r"#define EK_THERMO_VERSION (\d+)", open(os.path.join(ROOT, "include", "ek_thermo.h")).read()).group(1))


def test_python_binding_covers_the_reference_api():
    import thermo_oracle as oracle
    from ek_thermo import thermo

    assert sorted(thermo.__all__) == sorted(oracle.PUBLIC_NAMES) and len(thermo.__all__) == 39
    for name in oracle.PUBLIC_NAMES:
        assert callable(getattr(thermo, name)) and callable(getattr(thermo.array, name))


def test_argument_errors_without_a_gpu():
    """Error paths that never reach a kernel launch."""
    import torch

    from ek_thermo import _backend, thermo

    t = torch.ones(4, dtype=torch.float64)
    with pytest.raises(TypeError):
        thermo.potential_temperature(t, t)  # CPU tensor: no CPU path
    with pytest.raises(TypeError):
        thermo.potential_temperature(1.0, 2.0)  # no tensor at all
    with pytest.raises(ValueError):
        thermo.specific_humidity_from_vapour_pressure(t, t, eps=0.0)
    with pytest.raises(KeyError):
        thermo.ept_from_dewpoint(t, t, t, method="nope")
    with pytest.raises(ValueError):
        thermo.lcl(t, t, t, method="nope")
    assert thermo.saturation_vapour_pressure(t, phase="nope") is None
    with pytest.raises(ValueError):
        _backend.set_launch_config(threads=100)
    with pytest.raises(ValueError):
        _backend.shard_range(10, 0, 0)


@pytest.mark.parametrize("n,world,align", [(0, 1, 1), (10, 3, 1), (1000, 8, 16), (904156160, 8, 6599680), (11608481280, 8, 1661440),
                                           (7, 8, 1), (1_000_003, 4, 2048)])
def test_shard_ranges_partition_the_field(n, world, align):
    from ek_thermo import partition

    shards = partition.all_shards(n, world, align)
    assert shards[0][0] == 0 and shards[-1][1] == n
    for (b0, e0), (b1, e1) in zip(shards, shards[1:]):
        assert e0 == b1 and b0 <= e0
    for b, e in shards[:-1]:
        assert b % align == 0 and e % align == 0
    sizes = [e - b for b, e in shards[:-1]]
    if sizes:
        assert max(sizes) - min(sizes) <= align  # balanced to one slab


def test_world_size_2_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]


def test_host_binding_helper_without_a_gpu():
    """CPU-list parsing, and the NUMA binding is a no-op (None, no exception) when there is no device to ask."""
    import torch

    from ek_thermo import hostpipe

    assert hostpipe._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostpipe._parse_cpulist("") == set()
    if not torch.cuda.is_available():
        before = os.sched_getaffinity(0)
        assert hostpipe.bind_host_to_device("cuda:0") is None
        assert os.sched_getaffinity(0) == before


def test_host_array_front_end_has_no_cpu_path():
    """ek_thermo.host mirrors the reference's names for numpy input; without a GPU it raises instead of computing on the CPU."""
    import numpy as np
    import torch

    from ek_thermo import host, thermo, wind

    assert sorted(host.thermo.__all__) == sorted(thermo.__all__) and sorted(host.wind.__all__) == sorted(wind.__all__)
    assert host.thermo.array is host.thermo
    assert host.thermo.potential_temperature.__name__ == "potential_temperature"
    with pytest.raises(ValueError):
        host.set_chunk_elements(0)
    # what counts as an array argument: numpy arrays / scalars and nested numbers, not tuples of option strings
    assert host._is_arraylike(np.zeros(3)) and host._is_arraylike([1.0, 2.0]) and host._is_arraylike(np.float32(1.0)) and host._is_arraylike(())
    assert not host._is_arraylike(("theta", "rh")) and not host._is_arraylike(1.0) and not host._is_arraylike("mixed") and not host._is_arraylike(None)
    assert callable(host.fused.suite_tqp) and callable(host.fused.ept_wet_bulb)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            host.thermo.potential_temperature(np.full(4, 280.0), np.full(4, 9.0e4))


def test_bench_shard_plan_partitions_the_ensemble():
    """bench.py's partitioned workload (BASELINE.json configs[3]): the ranks' slab ranges tile the 51 x 137 slabs of the
    ensemble without gap or overlap for every world size the driver uses; a shard that does not fit in HBM is capped and
    says so; the weak workloads give every rank the whole per-GPU field."""
    import bench
    from ek_thermo import partition

    wl = bench.WORKLOADS["conv_ens_o640_f64"]
    total = 51 * 137
    for world in (1, 2, 4, 8):
        plans = [bench.shard_plan(wl, world, r) for r in range(world)]
        edges = []
        for r in range(world):
            b, e = partition.shard_range(total * wl["npl"], world, r, align=wl["npl"])
            edges.append((b // wl["npl"], e // wl["npl"]))
            assert plans[r][0] == edges[-1][0]
            assert plans[r][1] == min(edges[-1][1] - edges[-1][0], 1880) and plans[r][2] == (edges[-1][1] - edges[-1][0] > 1880)
        assert edges[0][0] == 0 and edges[-1][1] == total and all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
        assert all(p[2] for p in plans) == (world < 4)  # 4 and 8 GPUs hold the whole ensemble, 1 and 2 are capped at 150 GB
    assert bench.shard_plan(bench.WORKLOADS["suite_tqp_o1280x137_f64"], 8, 5) == (0, 137, False)
