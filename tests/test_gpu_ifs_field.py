"""GPU parity on the BENCHMARK'S OWN inputs: the IFS-shaped fields bench.py times (tools/synthetic.IfsField -- real L137 A/B
coefficients, top model levels at 1-100 Pa where the reference's NaN rule p - es < 1e-4 is live), at BASELINE.json's full
sizes through size-independent properties plus an oracle check of a strided sample (the checker is bench.parity_sample, the
very function that fills the `parity` field of the benchmark line), and level-slab checks of every ept formulation x solver
x humidity kind in float64 and float32.  Run on the B200 box with ``pytest -m gpu``."""
import numpy as np
import pytest
import torch

import bench
import thermo_oracle as oracle
from cases import CASE_BY_ID
from compare import compare, conditioning, reference_f32_noise
from synthetic import O640_POINTS, O1280_POINTS, IfsField, ifs_point_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ek():
    import ek_thermo

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return ek_thermo


@pytest.fixture(scope="module")
def o1280(ek):
    """BASELINE.json configs[1] / [2]: the O1280 x 137 float64 field of `bench.py` (seed 0 = rank 0's field)."""
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~95 GB of free HBM")
    field = IfsField("tqp", O1280_POINTS, levels=137, seed=0, device=DEV)
    arrays = field.materialise(0, 137, torch.float64)
    yield arrays
    del arrays
    torch.cuda.empty_cache()


def _assert_parity(res, kind, outputs, arrays, out, dtype):
    """Flat contract limit first; a point over it must be explained by the conditioning of the formula itself."""
    assert res["nan_mismatches"] == 0 and res["inf_mismatches"] == 0, res
    if res["n_over_limit"] == 0:
        return
    # which points?  recompute the sample and judge the exceeding points by max(limit, 4 x conditioning)
    n = arrays[0].numel()
    idx = torch.arange(0, n, res["stride"], device=arrays[0].device)
    t, h, p = (a[idx].cpu().numpy() for a in arrays)
    for name in outputs:
        if res["per_output"][name]["n_over_limit"] == 0:
            continue
        got = out[name][idx].cpu().numpy().astype(np.float64)
        with np.errstate(all="ignore"):
            want = np.asarray(bench.OUT_FNS[kind][name](oracle, t, h, p)).astype(np.float64)
            rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
        over = np.flatnonzero(rel > res["limit"])

        class _C:  # the oracle composition of this output as a "case" for compare.conditioning
            fn, kwargs = f"_bench_{kind}_{name}", {}

        setattr(oracle, _C.fn, lambda a, b, c, _f=bench.OUT_FNS[kind][name]: _f(oracle, a, b, c))
        cond = conditioning(_C, [t[over], h[over], p[over]])
        bad = rel[over] > np.maximum(res["limit"], 4.0 * cond)
        assert not bad.any(), (name, dtype, int(bad.sum()), float(rel[over][bad].max()))


def test_config1_suite_on_the_bench_field(ek, o1280):
    """configs[1]: theta, es, rh, td, Tv on O1280 x 137 (904 156 160 points), float64 -- the workload `bench.py` times by
    default.  (i) >= 1e6 strided points of the outputs equal the oracle (flat 1e-12, NaN positions identical);
    (ii) the fused outputs are the single-function kernels' outputs over the WHOLE field;
    (iii) theta <-> t and K <-> degC round trips over the whole field."""
    from ek_thermo import fused

    t, q, p = o1280
    out = fused.suite_tqp(t, q, p)
    res = bench.parity_sample("tqp", fused.DEFAULT_TQP, o1280, None, out, O1280_POINTS, "f64")
    assert res["n"] >= 1_000_000
    _assert_parity(res, "tqp", fused.DEFAULT_TQP, o1280, out, "f64")
    singles = {
        "theta": lambda: ek.thermo.potential_temperature(t, p),
        "es": lambda: ek.thermo.saturation_vapour_pressure(t),
        "rh": lambda: ek.thermo.relative_humidity_from_specific_humidity(t, q, p),
        "td": lambda: ek.thermo.dewpoint_from_specific_humidity(q, p),
        "tv": lambda: ek.thermo.virtual_temperature(t, q),
    }
    for name, fn in singles.items():
        one = fn()
        same = ((one - out[name]).abs() <= 1e-14 * one.abs()) | (torch.isnan(one) & torch.isnan(out[name]))
        assert bool(same.all()), name
        del one, same
    back = ek.thermo.temperature_from_potential_temperature(out["theta"], p)
    assert float(((back - t).abs() / t).max()) < 1e-13
    del back
    k2 = ek.thermo.celsius_to_kelvin(ek.thermo.kelvin_to_celsius(t))
    assert float((k2 - t).abs().max()) < 1e-12


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_single_pass_suite_on_the_bench_field(ek, o1280, dtype):
    """configs[1] + configs[2] in ONE launch (north_star item 3): theta, es, rh, td, Tv, theta_e, theta_w from one read of
    (t, q, p) over the full O1280 x 137 field.  The seven outputs equal the oracle on the strided sample; theta_e and
    theta_w are the bits of the two-output ept / wet-bulb kernel, the other five the bits of the five-output suite;
    theta_e >= theta and theta_w <= theta_e wherever both are finite."""
    from ek_thermo import fused

    tdt = torch.float64 if dtype == "f64" else torch.float32
    arrays = [x.to(tdt) for x in o1280] if dtype == "f32" else o1280
    t, q, p = arrays
    before = ek.launch_count()
    out = fused.suite_tqp(t, q, p, outputs=fused.ALL7_TQP)
    assert ek.launch_count() == before + 1
    res = bench.parity_sample("tqp", fused.ALL7_TQP, arrays, None, out, O1280_POINTS, dtype)
    _assert_parity(res, "tqp", fused.ALL7_TQP, arrays, out, dtype)

    def same_bits(a, b):
        return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())

    ept, wb = fused.ept_wet_bulb(t, q, p, humidity="q", ept_method="ifs", t_method="direct")
    assert same_bits(out["ept"], ept) and same_bits(out["wbpt"], wb)
    fin = torch.isfinite(ept) & torch.isfinite(wb)
    assert bool((ept[fin] >= out["theta"][fin] * (1 - 1e-6)).all()) and bool((wb[fin] <= ept[fin]).all())
    del ept, wb, fin
    five = fused.suite_tqp(t, q, p)
    for name in fused.DEFAULT_TQP:
        assert same_bits(out[name], five[name]), name
    del five
    sp5 = fused.suite_tqp(t, q, p, outputs=fused.SINGLE_PASS_TQP)  # theta, rh, td, theta_e, theta_w: the other static instantiation
    for name in fused.SINGLE_PASS_TQP:
        assert same_bits(out[name], sp5[name]), name


def test_config4_ens_shard_on_the_bench_field(ek):
    """configs[3]: one GPU's shard (874 member x level slabs, 1 452 098 560 points) of ENS 51 x O640 x 137, humidity /
    dewpoint conversions on the IFS-shaped ensemble field.  Strided sample = oracle; q -> td -> q, td -> r -> td and
    q -> w -> q round trips over the whole shard; the shard edges the partitioner produces."""
    from ek_thermo import fused, partition

    n_total, slab = 51 * 137 * O640_POINTS, O640_POINTS
    b, e = partition.shard_range(n_total, 8, 3, align=slab)
    assert (e - b) in (873 * slab, 874 * slab)
    torch.cuda.empty_cache()  # blocks the caching allocator still holds from earlier tests do not count as free
    free, _ = torch.cuda.mem_get_info()
    if free < 110e9:
        pytest.skip("needs ~105 GB of free HBM")
    field = IfsField("tqp", slab, levels=137, seed=0, device=DEV)
    arrays = field.materialise(b // slab, (e - b) // slab, torch.float64)
    t, q, p = arrays
    outputs = ("rh", "td", "w")
    out = fused.suite_tqp(t, q, p, outputs=outputs)
    res = bench.parity_sample("tqp", outputs, arrays, None, out, slab, "f64")
    _assert_parity(res, "tqp", outputs, arrays, out, "f64")
    th = ek.thermo
    q2 = th.specific_humidity_from_dewpoint(out["td"], p)
    ok = torch.isfinite(q2)
    assert float(ok.double().mean()) > 0.98  # the NaN rule p - es(td) < 1e-4 fires at the top levels only
    assert float(((q2 - q).abs() / q)[ok].max()) < 1e-10  # q -> td -> q
    del q2, ok
    r = th.relative_humidity_from_dewpoint(t, out["td"])
    td2 = th.dewpoint_from_relative_humidity(t, r)
    assert float(((td2 - out["td"]).abs() / out["td"]).max()) < 1e-11  # td -> r -> td
    del td2, r
    q3 = th.specific_humidity_from_mixing_ratio(out["w"])
    assert float(((q3 - q).abs() / q).max()) < 1e-14


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
@pytest.mark.parametrize("t_method", ["direct", "newton", "bisect"])
@pytest.mark.parametrize("hum", ["q", "td"])
def test_moist_adiabat_functions_on_ifs_columns(ek, dtype, ept_method, t_method, hum):
    """Every ept formulation x solver x humidity kind on IFS L137 columns (900 columns x 137 levels, top levels at
    1-100 Pa): ept, the wet-bulb potential temperature and (newton / bisect) the wet-bulb temperature against the oracle
    under the parity rule of tests/compare.py."""
    inp = ifs_point_inputs(900, seed=11)
    sfx = "specific_humidity" if hum == "q" else "dewpoint"
    ids = [f"ept_from_{sfx}(t,{hum},p;method={ept_method})",
           f"wet_bulb_potential_temperature_from_{sfx}(t,{hum},p;ept_method={ept_method},t_method={t_method})"]
    if t_method != "direct":
        ids.append(f"wet_bulb_temperature_from_{sfx}(t,{hum},p;ept_method={ept_method},t_method={t_method})")
    for cid in ids:
        case = CASE_BY_ID[cid]
        args_np = [np.ascontiguousarray(inp[a].astype(dtype)) for a in case.args]
        got = getattr(ek.thermo, case.fn)(*[torch.from_numpy(a).to(DEV) for a in args_np], **case.kwargs)
        with np.errstate(all="ignore"):
            want = getattr(oracle, case.fn)(*args_np, **case.kwargs)
        cond = None if case.iterative == "bisect" else conditioning(case, args_np)
        noise = reference_f32_noise(case, args_np) if (dtype == np.float32 and case.iterative != "bisect") else None
        compare(case, got.cpu().numpy(), want, dtype, cond=cond, noise=noise)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_hybrid_suite_on_the_bench_field(ek, dtype):
    """SURVEY 8(f)-1 on the benchmark's field shape (24 bottom levels of O1280): the hybrid-level suite, seven outputs, against
    the oracle with p = ph_k + 0.5 (ph_k+1 - ph_k) -- through bench.parity_sample's hybrid branch."""
    from ek_thermo import fused

    tdt = torch.float64 if dtype == "f64" else torch.float32
    nlev, npl = 24, O1280_POINTS
    field = IfsField("hybrid", npl, levels=nlev, seed=0, device=DEV)
    arrays = field.materialise(0, nlev, tdt)
    hyb = (field.sp(0).to(tdt), torch.tensor(field.A_half, dtype=tdt, device=DEV), torch.tensor(field.B_half, dtype=tdt, device=DEV))
    out = fused.suite_tq_hybrid(arrays[0].reshape(nlev, npl), arrays[1].reshape(nlev, npl), hyb[0], hyb[1], hyb[2], outputs=fused.ALL7_TQP)
    flat = {k: v.reshape(-1) for k, v in out.items()}
    res = bench.parity_sample("hybrid", fused.ALL7_TQP, arrays, hyb, flat, npl, dtype)
    assert res["nan_mismatches"] == 0 and res["inf_mismatches"] == 0, res
    if dtype == "f64":
        assert res["n_over_limit"] == 0, res
    else:
        assert res["n_over_limit"] <= 1e-3 * res["n"] * len(fused.ALL7_TQP), res
