"""GPU parity tests: the CUDA path, called through the C ABI by the drop-in package, against the oracle
on identical inputs, against the committed live-reference fixtures, the reference's golden CSVs and its
inline known-answer vectors.  Run on the B200 box with ``pytest -m gpu``.
"""
import numpy as np
import pytest
import torch

import thermo_oracle as oracle
from cases import CASES, edge_inputs, random_inputs
from compare import compare, conditioning, reference_f32_noise
from kat import KATS

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ek():
    import ek_thermo

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return ek_thermo


def _to_dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(DEV)


def _run_case(ek, case, inputs, dtype):
    args_np = [np.ascontiguousarray(inputs[a].astype(dtype)) for a in case.args]
    before = ek.launch_count()
    res = getattr(ek.thermo, case.fn)(*[torch.from_numpy(a).to(DEV) for a in args_np], **case.kwargs)
    # one kernel per call; the lean float64 bisection adds the ~3 us launch that tabulates its 4095-node tree
    assert ek.launch_count() - before in ((1, 2) if case.iterative == "bisect" else (1,)), "one kernel launch per call"
    with np.errstate(all="ignore"):
        want = getattr(oracle, case.fn)(*args_np, **case.kwargs)
    if not isinstance(res, tuple):
        res, want = (res,), (want,)
    conds = [None if case.iterative == "bisect" else conditioning(case, args_np, k) for k in range(len(res))]
    if dtype == np.float32 and case.iterative != "bisect":  # float32: (conditioning, the reference's own float32 noise)
        conds = [(c, reference_f32_noise(case, args_np, k)) for k, c in enumerate(conds)]
    return [r.cpu().numpy() for r in res], want, conds


# N chosen so that the vector body (several tiles), the scalar tail and a ragged end are all exercised
N_RANDOM = 256 * 4 * 37 + 77


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_cuda_matches_oracle_random(ek, case, dtype):
    got, want, conds = _run_case(ek, case, random_inputs(N_RANDOM, seed=5), dtype)
    for g, w, c in zip(got, want, conds):
        c, nz = c if isinstance(c, tuple) else (c, None)
        compare(case, g, w, dtype, cond=c, noise=nz)


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_cuda_matches_oracle_edge(ek, case, dtype):
    with np.errstate(all="ignore"):
        got, want, conds = _run_case(ek, case, edge_inputs(n=4099, seed=21), dtype)
    for g, w, c in zip(got, want, conds):
        c, nz = c if isinstance(c, tuple) else (c, None)
        compare(case, g, w, dtype, edge=True, cond=c, noise=nz)


@pytest.mark.parametrize("sname,dname", [("rand", "float64"), ("edge", "float64"), ("grid", "float64"), ("ma", "float64"),
                                         ("rand", "float32")])
def test_cuda_matches_live_reference_fixtures(ek, ref_live, sname, dname):
    """Outputs of the unmodified reference (tests/golden/make_golden.py) -- independent of the oracle."""
    dtype = np.dtype(dname).type
    pre = f"in/{sname}/"
    inputs = {k[len(pre):]: v for k, v in ref_live.items() if k.startswith(pre)}
    n = 0
    for case in CASES:
        if f"out/{sname}/{dname}/{case.id}/0" not in ref_live:
            continue
        args_np = [np.ascontiguousarray(inputs[a].astype(dtype)) for a in case.args]
        res = getattr(ek.thermo, case.fn)(*[torch.from_numpy(a).to(DEV) for a in args_np], **case.kwargs)
        res = res if isinstance(res, tuple) else (res,)
        for k, r in enumerate(res):
            cond = None if case.iterative == "bisect" else conditioning(case, args_np, k)
            noise = reference_f32_noise(case, args_np, k) if (dtype == np.float32 and case.iterative != "bisect") else None
            compare(case, r.cpu().numpy(), ref_live[f"out/{sname}/{dname}/{case.id}/{k}"], dtype, edge=(sname == "edge"), cond=cond,
                    grid=(sname in ("grid", "ma")), noise=noise)
            n += 1
    assert n > 50 or sname == "ma"


@pytest.mark.parametrize("kat", KATS, ids=[f"{i}-{k[0]}" for i, k in enumerate(KATS)])
def test_kat_reference_numbers(ek, kat):
    fn, args, kwargs, expected, rtol = kat
    got = getattr(ek.thermo, fn)(*[torch.tensor(a, dtype=torch.float64, device=DEV) for a in args], **kwargs)
    if not isinstance(got, tuple):
        got, expected = (got,), (expected,)
    for g, e in zip(got, expected):
        np.testing.assert_allclose(g.cpu().numpy(), np.asarray(e, dtype=np.float64), rtol=rtol, atol=1e-8, equal_nan=True)


# ---- the reference's golden CSVs with the reference tests' own tolerances (TT:169-356, 706-849) ----
def test_csv_goldens(ek, ref_csv):
    th = ek.thermo
    d = lambda k: torch.from_numpy(ref_csv[k]).to(DEV)  # noqa: E731
    close = lambda a, b, **kw: np.testing.assert_allclose(a.cpu().numpy(), ref_csv[b], equal_nan=True, **kw)  # noqa: E731
    for ph in ("mixed", "water", "ice"):
        close(th.saturation_vapour_pressure(d("sat_vp/t"), phase=ph), f"sat_vp/{ph}", rtol=1e-12)
        close(th.saturation_vapour_pressure_slope(d("sat_vp_slope/t"), phase=ph), f"sat_vp_slope/{ph}", rtol=1e-12)
        close(th.saturation_mixing_ratio(d("sat_mr/t"), d("sat_mr/p"), phase=ph), f"sat_mr/{ph}", rtol=1e-12)
        close(th.saturation_specific_humidity(d("sat_q/t"), d("sat_q/p"), phase=ph), f"sat_q/{ph}", rtol=1e-12)
        close(th.saturation_mixing_ratio_slope(d("sat_mr_slope/t"), d("sat_mr_slope/p"), phase=ph), f"sat_mr_slope/{ph}", rtol=1e-12)
        close(th.saturation_specific_humidity_slope(d("sat_q_slope/t"), d("sat_q_slope/p"), phase=ph), f"sat_q_slope/{ph}", rtol=1e-12)
    t, td, q, p = (d(f"t_hum_p_data/{k}") for k in ("t", "td", "q", "p"))
    for m in ("ifs", "bolton35", "bolton39"):
        close(th.ept_from_dewpoint(t, td, p, method=m), f"eqpt/{m}_td", rtol=1e-12)
        close(th.ept_from_specific_humidity(t, q, p, method=m), f"eqpt/{m}_q", rtol=1e-12)
        close(th.saturation_ept(t, p, method=m), f"seqpt/{m}", rtol=1e-12)
        for tm in ("bisect", "newton"):
            rt = 1e-3 if tm == "bisect" else 2e-9
            close(th.temperature_on_moist_adiabat(d("t_on_most_adiabat/ept"), d("t_on_most_adiabat/p"), ept_method=m, t_method=tm),
                  f"t_on_most_adiabat/{m}_{tm}", rtol=rt, atol=0)
            close(th.wet_bulb_temperature_from_dewpoint(t, td, p, ept_method=m, t_method=tm), f"t_wet/{m}_{tm}_td", rtol=rt, atol=0)
            close(th.wet_bulb_temperature_from_specific_humidity(t, q, p, ept_method=m, t_method=tm), f"t_wet/{m}_{tm}_q", rtol=rt, atol=0)
        for tm in ("bisect", "newton", "direct"):
            rt = 1e-3 if tm == "bisect" else 2e-9
            close(th.wet_bulb_potential_temperature_from_dewpoint(t, td, p, ept_method=m, t_method=tm), f"t_wetpt/{m}_{tm}_td", rtol=rt, atol=0)
            close(th.wet_bulb_potential_temperature_from_specific_humidity(t, q, p, ept_method=m, t_method=tm), f"t_wetpt/{m}_{tm}_q", rtol=rt, atol=0)


# ---- calling conventions ------------------------------------------------------------------------
def test_scalar_broadcast_unaligned_nd_noncontiguous(ek):
    th = ek.thermo
    inp = random_inputs(50021, seed=9)
    t, p, q = (torch.from_numpy(inp[k]).to(DEV) for k in ("t", "p", "q"))
    want = oracle.potential_temperature(inp["t"], 85000.0)
    np.testing.assert_allclose(th.potential_temperature(t, 85000.0).cpu().numpy(), want, rtol=1e-12)  # python scalar p (TT:622)
    np.testing.assert_allclose(th.potential_temperature(t, torch.tensor(85000.0, device=DEV, dtype=torch.float64)).cpu().numpy(), want, rtol=1e-12)
    # unaligned views: pointers offset by one element (8 bytes) take the scalar ld/st path
    want = oracle.relative_humidity_from_specific_humidity(inp["t"][1:], inp["q"][1:], inp["p"][1:])
    np.testing.assert_allclose(th.relative_humidity_from_specific_humidity(t[1:], q[1:], p[1:]).cpu().numpy(), want, rtol=1e-12)
    want = oracle.potential_temperature(inp["t"][:-1], inp["p"][1:])
    np.testing.assert_allclose(th.potential_temperature(t[:-1], p[1:]).cpu().numpy(), want, rtol=1e-12)
    # N-D and non-contiguous
    t2, p2 = t[:50000].reshape(50, 1000), p[:50000].reshape(50, 1000)
    got = th.potential_temperature(t2.t(), p2.t())
    assert got.shape == (1000, 50)
    np.testing.assert_allclose(got.cpu().numpy(), oracle.potential_temperature(inp["t"][:50000].reshape(50, 1000).T, inp["p"][:50000].reshape(50, 1000).T), rtol=1e-12)
    # broadcasting a per-level pressure column against a [level, point] field
    pl = p[:50].reshape(50, 1)
    got = th.potential_temperature(t2, pl)
    np.testing.assert_allclose(got.cpu().numpy(), oracle.potential_temperature(inp["t"][:50000].reshape(50, 1000), inp["p"][:50].reshape(50, 1)), rtol=1e-12)
    # N-D through the iterative solvers (a superset of the reference, which is 1-D only there)
    e2 = torch.from_numpy(inp["ept"][:50000]).to(DEV).reshape(50, 1000)
    got = th.temperature_on_moist_adiabat(e2, p2, t_method="newton")
    np.testing.assert_allclose(got.cpu().numpy().ravel(), oracle.temperature_on_moist_adiabat(inp["ept"][:50000], inp["p"][:50000], t_method="newton"),
                               rtol=2e-9, equal_nan=True)
    # empty input
    assert th.potential_temperature(t[:0], p[:0]).numel() == 0
    # optional precomputed es / es_slope
    es = th.saturation_vapour_pressure(t)
    des = th.saturation_vapour_pressure_slope(t)
    a = th.saturation_mixing_ratio_slope(t, p)
    b = th.saturation_mixing_ratio_slope(t, p, es=es, es_slope=des)
    torch.testing.assert_close(a, b, rtol=1e-13, atol=0, equal_nan=True)  # separate kernels: FMA contraction may differ by an ulp
    # float32 in -> float32 out; mixed dtypes promote
    assert th.potential_temperature(t.float(), p.float()).dtype == torch.float32
    assert th.potential_temperature(t.float(), p).dtype == torch.float64


def test_error_conventions(ek):
    th = ek.thermo
    t = torch.full((8,), 280.0, device=DEV, dtype=torch.float64)
    with pytest.raises(ValueError):
        th.specific_humidity_from_vapour_pressure(t, t, eps=0)  # T:189-190
    with pytest.raises(ValueError):
        th.saturation_mixing_ratio_slope(t, t, eps=-1.0)  # T:405-406
    with pytest.raises(ValueError):
        th.lcl_temperature(t, t, method="x")  # T:968
    with pytest.raises(KeyError):
        th.ept_from_dewpoint(t, t, t, method="x")  # T:1026
    with pytest.raises(ValueError):
        th.temperature_on_moist_adiabat(t, t, t_method="x")  # T:1509
    with pytest.raises(ValueError):
        th.wet_bulb_temperature_from_dewpoint(t, t, t, t_method="direct")
    assert th.saturation_vapour_pressure(t, phase="x") is None  # E:74-79
    with pytest.raises(TypeError):
        th.potential_temperature(t.cpu(), t.cpu())  # no CPU path
    with pytest.raises(TypeError):
        th.potential_temperature(np.ones(4), np.ones(4))
    # inputs are never modified
    t0 = t.clone()
    th.saturation_vapour_pressure(t)
    assert torch.equal(t, t0)


# ---- fused suites ---------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_fused_suites_match_oracle(ek, dtype):
    from ek_thermo import fused

    inp = random_inputs(N_RANDOM, seed=13)
    a = {k: np.ascontiguousarray(inp[k].astype(dtype)) for k in ("t", "q", "td", "p")}
    d = {k: torch.from_numpy(v).to(DEV) for k, v in a.items()}
    f32 = dtype == np.float32
    rtol = 1e-5 if f32 else 1e-12

    def check(got, suite_fn, args, names):
        with np.errstate(all="ignore"):
            want = suite_fn(*args)
            pert = [suite_fn(*[np.nextafter(a, np.asarray(np.inf, dtype=a.dtype)) if i == j else a for j, a in enumerate(args)])
                    for i in range(len(args))]
            want64 = suite_fn(*[a.astype(np.float64) for a in args]) if f32 else None  # the reference's own float32 noise
        for name in names:
            g = got[name].cpu().numpy().astype(np.float64)
            w = np.asarray(want[name]).astype(np.float64)
            if f32:  # the oracle's "direct" fit runs in float64 and overflows later than float32 does (SURVEY 8(c) caveat)
                assert np.mean(np.isfinite(g) != np.isfinite(w)) < 0.005, name
            else:
                np.testing.assert_array_equal(np.isnan(g), np.isnan(w), err_msg=name)
            fin = np.isfinite(w) & np.isfinite(g)
            den = np.maximum(np.abs(w[fin]), 1e-300)
            cond = np.zeros(den.shape)
            for pw in pert:  # 1-ulp input conditioning, see compare.conditioning
                dlt = np.abs(np.asarray(pw[name]).astype(np.float64)[fin] - w[fin]) / den
                cond = np.fmax(cond, np.where(np.isfinite(dlt), dlt, 0.0))
            tol = np.maximum(rtol, 4 * cond)
            if f32:
                nz = np.abs(np.asarray(want64[name])[fin] - w[fin]) / den
                tol = np.maximum(tol, 4 * np.where(np.isfinite(nz), nz, 0.0))
            rel = np.abs(g[fin] - w[fin]) / den
            assert np.mean(rel > tol) <= (0.0005 if f32 else 0.0), (name, rel.max(), int(np.sum(rel > tol)))

    before = ek.launch_count()
    got = fused.suite_tqp(d["t"], d["q"], d["p"], outputs=tuple(fused.SUITE_TQP_OUTPUTS))
    assert ek.launch_count() == before + 1
    check(got, oracle.suite_tqp, (a["t"], a["q"], a["p"]), fused.SUITE_TQP_OUTPUTS)
    got = fused.suite_ttdp(d["t"], d["td"], d["p"], outputs=tuple(fused.SUITE_TTDP_OUTPUTS))
    check(got, oracle.suite_ttdp, (a["t"], a["td"], a["p"]), fused.SUITE_TTDP_OUTPUTS)
    # every subset of outputs gives the same fields as the full run (the mask only skips work; the subsets with a
    # dedicated compile-time instantiation may contract a*b+c differently, hence "a few ulp" and not "bit-identical")
    full = fused.suite_tqp(d["t"], d["q"], d["p"], outputs=tuple(fused.SUITE_TQP_OUTPUTS))
    for names in (("theta",), ("rh",), ("td", "tv"), ("theta", "rh"), fused.DEFAULT_TQP, ("w", "e", "thetav")):
        part = fused.suite_tqp(d["t"], d["q"], d["p"], outputs=names)
        assert set(part) == set(names)
        for nme in names:
            torch.testing.assert_close(part[nme], full[nme], rtol=(1e-6 if f32 else 1e-14), atol=0, equal_nan=True, msg=nme)
    # scalar pressure (a pressure level) and preallocated outputs
    out = {"theta": torch.empty_like(d["t"])}
    r = fused.suite_tqp(d["t"], d["q"], 85000.0, outputs=("theta", "rh"), out=out)
    assert r["theta"].data_ptr() == out["theta"].data_ptr()
    np.testing.assert_allclose(r["theta"].cpu().numpy().astype(np.float64), oracle.potential_temperature(a["t"], dtype(85000.0)).astype(np.float64), rtol=rtol)


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
@pytest.mark.parametrize("t_method", ["direct", "newton", "bisect"])
def test_fused_ept_wet_bulb_equals_separate_calls(ek, ept_method, t_method):
    from ek_thermo import fused

    inp = random_inputs(20011, seed=17)
    t, q, td, p = (torch.from_numpy(inp[k]).to(DEV) for k in ("t", "q", "td", "p"))
    for hum, h in (("q", q), ("td", td)):
        ept, wb = fused.ept_wet_bulb(t, h, p, humidity=hum, ept_method=ept_method, t_method=t_method, potential=True)
        sfx = "specific_humidity" if hum == "q" else "dewpoint"
        e1 = getattr(ek.thermo, f"ept_from_{sfx}")(t, h, p, method=ept_method)
        w1 = getattr(ek.thermo, f"wet_bulb_potential_temperature_from_{sfx}")(t, h, p, ept_method=ept_method, t_method=t_method)
        # same formulas in two kernels: the compiler may contract a*b+c differently, so allow a few ulp
        # (bisect: a flipped near-tie sign moves the quantised result by up to two last steps, 0.06 K)
        torch.testing.assert_close(ept, e1, rtol=1e-11, atol=0, equal_nan=True)  # |exponent| reaches 200 on unphysical points
        if t_method == "bisect":
            d = (wb - w1).abs()
            ok = (d <= 1e-12 * w1.abs()) | (torch.isnan(wb) & torch.isnan(w1))
            assert float((~ok).float().mean()) < 0.005 and float(torch.nan_to_num(d).max()) < 0.0616
        else:
            torch.testing.assert_close(wb, w1, rtol=2e-9, atol=0, equal_nan=True)


# (the full-size properties on the benchmark's own IFS-shaped fields live in tests/test_gpu_ifs_field.py)


# ---- host-buffer pipeline and partitioner -----------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_host_pipeline_equals_device_path(ek, dtype):
    from ek_thermo import fused, hostpipe

    n = 3_000_017
    inp = random_inputs(n, seed=23)
    hs = hostpipe.HostSuite(DEV, workspace_bytes=64 << 20, n_slots=3)  # small workspace -> many chunks
    host = {}
    for k in ("t", "q", "td", "p"):
        host[k] = hostpipe.pinned_empty(n, dtype)
        host[k][:] = inp[k]
    res = hs.suite_tqp(host["t"], host["q"], host["p"])
    dev = fused.suite_tqp(*(torch.from_numpy(host[k]).to(DEV) for k in ("t", "q", "p")))
    for name in fused.DEFAULT_TQP:
        np.testing.assert_array_equal(res[name], dev[name].cpu().numpy(), err_msg=name)
    res = hs.suite_ttdp(host["t"], host["td"], host["p"], outputs=("rh", "q"))
    dev = fused.suite_ttdp(*(torch.from_numpy(host[k]).to(DEV) for k in ("t", "td", "p")), outputs=("rh", "q"))
    for name in ("rh", "q"):
        np.testing.assert_array_equal(res[name], dev[name].cpu().numpy(), err_msg=name)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
def test_single_pass_suite_equals_the_ept_kernels(ek, dtype, ept_method):
    """Slots 8 / 9 of the suites (theta_e, theta_w "direct") are the bits of the two-output ept / wet-bulb kernel and of the
    reference-named entry points, for (t, q, p) and (t, td, p), every formulation, the run-time-mask kernel and the
    compile-time-mask instantiations -- and equal the oracle composition under the parity rule."""
    from ek_thermo import fused

    inp = random_inputs(N_RANDOM, seed=31)
    a = {k: np.ascontiguousarray(inp[k].astype(dtype)) for k in ("t", "q", "td", "p")}
    d = {k: torch.from_numpy(v).to(DEV) for k, v in a.items()}

    def same_bits(x, y):
        return bool(((x == y) | (torch.isnan(x) & torch.isnan(y))).all())

    for suite, hname, hum, sfx in ((fused.suite_tqp, "q", "q", "specific_humidity"), (fused.suite_ttdp, "td", "td", "dewpoint")):
        ept, wb = fused.ept_wet_bulb(d["t"], d[hname], d["p"], humidity=hum, ept_method=ept_method, t_method="direct")
        table = fused.SUITE_TQP_OUTPUTS if hum == "q" else fused.SUITE_TTDP_OUTPUTS
        sets = [("ept", "wbpt"), tuple(table), ("wbpt",), ("rh", "ept")]
        if hum == "q":
            sets += [fused.ALL7_TQP, fused.SINGLE_PASS_TQP]
        else:
            sets += [fused.ALL7_TTDP, fused.SINGLE_PASS_TTDP]
        for outputs in sets:
            before = ek.launch_count()
            got = suite(d["t"], d[hname], d["p"], outputs=outputs, ept_method=ept_method)
            assert ek.launch_count() == before + 1 and tuple(got) == tuple(outputs)
            # a point with a NaN in ANY of its outputs is recomputed as a whole by the exact functor (libdevice math): there
            # the other outputs agree to rounding, not bit for bit; everywhere else the bits are those of the ept kernel
            fast = ~torch.isnan(ept) & ~torch.isnan(wb)  # (the two-output kernel follows the same rule: a NaN wet bulb recomputes its ept)
            for v in got.values():
                fast &= ~torch.isnan(v)
            for name, ref in (("ept", ept), ("wbpt", wb)):
                if name in got:
                    assert bool((got[name][fast] == ref[fast]).all()), (hum, outputs, name)
                    torch.testing.assert_close(got[name][~fast], ref[~fast], rtol=(1e-5 if dtype == np.float32 else 1e-11), atol=0, equal_nan=True)
        e1 = getattr(ek.thermo, f"ept_from_{sfx}")(d["t"], d[hname], d["p"], method=ept_method)
        w1 = getattr(ek.thermo, f"wet_bulb_potential_temperature_from_{sfx}")(d["t"], d[hname], d["p"], ept_method=ept_method)
        both = ~torch.isnan(ept) & ~torch.isnan(wb)  # (a NaN wet bulb sends the point's ept to the exact functor as well)
        assert same_bits(e1[both], ept[both]) and same_bits(w1[both], wb[both])
        torch.testing.assert_close(e1, ept, rtol=(1e-5 if dtype == np.float32 else 1e-11), atol=0, equal_nan=True)
        for fn_name, g in ((f"ept_from_{sfx}", ept), (f"wet_bulb_potential_temperature_from_{sfx}", wb)):
            kw = {"method": ept_method} if fn_name.startswith("ept") else {"ept_method": ept_method, "t_method": "direct"}
            case = next(c for c in CASES if c.fn == fn_name and c.kwargs == kw)
            args_np = [a[x] for x in case.args]
            with np.errstate(all="ignore"):
                want = getattr(oracle, fn_name)(*args_np, **kw)
            compare(case, g.cpu().numpy(), want, dtype, cond=conditioning(case, args_np),
                    noise=(reference_f32_noise(case, args_np) if dtype == np.float32 else None))
    with pytest.raises(KeyError):
        fused.suite_tqp(d["t"], d["q"], d["p"], outputs=("ept",), ept_method="nope")  # as the reference (T:1026)


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
def test_suites_on_special_values_equal_the_single_functions(ek, ept_method):
    """The special-value set (NaN, +-inf, 0, negative and subnormal inputs, band edges and their neighbours, p - e straddling the
    NaN rule) through the ten-output suites: NaN positions and infinities are those of the single-function kernels -- which
    test_cuda_matches_oracle_edge holds to the oracle -- and finite values agree with them to rounding (a point with a NaN in any
    output is recomputed as a whole by the exact functor, so bits may differ there)."""
    from ek_thermo import fused

    with np.errstate(all="ignore"):
        inp = edge_inputs(n=4099, seed=33)
    d = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).to(DEV) for k in ("t", "q", "td", "p")}
    th = ek.thermo
    singles_q = {
        "theta": lambda: th.potential_temperature(d["t"], d["p"]), "es": lambda: th.saturation_vapour_pressure(d["t"]),
        "rh": lambda: th.relative_humidity_from_specific_humidity(d["t"], d["q"], d["p"]),
        "td": lambda: th.dewpoint_from_specific_humidity(d["q"], d["p"]), "tv": lambda: th.virtual_temperature(d["t"], d["q"]),
        "w": lambda: th.mixing_ratio_from_specific_humidity(d["q"]), "e": lambda: th.vapour_pressure_from_specific_humidity(d["q"], d["p"]),
        "thetav": lambda: th.virtual_potential_temperature(d["t"], d["q"], d["p"]),
        "ept": lambda: th.ept_from_specific_humidity(d["t"], d["q"], d["p"], method=ept_method),
        "wbpt": lambda: th.wet_bulb_potential_temperature_from_specific_humidity(d["t"], d["q"], d["p"], ept_method=ept_method),
    }
    q_td = th.specific_humidity_from_dewpoint(d["td"], d["p"])
    singles_td = {
        "theta": singles_q["theta"], "es": singles_q["es"], "rh": lambda: th.relative_humidity_from_dewpoint(d["t"], d["td"]),
        "q": lambda: q_td, "tv": lambda: th.virtual_temperature(d["t"], q_td), "w": lambda: th.mixing_ratio_from_dewpoint(d["td"], d["p"]),
        "e": lambda: th.saturation_vapour_pressure(d["td"], phase="water"),
        "thetav": lambda: th.virtual_potential_temperature(d["t"], q_td, d["p"]),
        "ept": lambda: th.ept_from_dewpoint(d["t"], d["td"], d["p"], method=ept_method),
        "wbpt": lambda: th.wet_bulb_potential_temperature_from_dewpoint(d["t"], d["td"], d["p"], ept_method=ept_method),
    }
    for suite, hname, singles in ((fused.suite_tqp, "q", singles_q), (fused.suite_ttdp, "td", singles_td)):
        for outputs in (tuple(singles), tuple(singles)[:5] + ("ept", "wbpt")):  # run-time mask and (ifs) the seven-output instantiation
            got = suite(d["t"], d[hname], d["p"], outputs=outputs, ept_method=ept_method)
            for name in outputs:
                g, w = got[name], singles[name]()
                assert torch.equal(torch.isnan(g), torch.isnan(w)), (hname, name)
                inf = torch.isinf(g) | torch.isinf(w)
                assert torch.equal(g[inf], w[inf]), (hname, name)
                fin = torch.isfinite(g) & torch.isfinite(w)
                tiny = (g[fin] - w[fin]).abs() <= 1e-300  # subnormal results
                rel = (g[fin] - w[fin]).abs() / w[fin].abs().clamp_min(1e-300)
                assert bool(((rel <= 1e-10) | tiny).all()), (hname, name, float(rel.max()))


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_host_pipelines_for_every_suite(ek, dtype):
    """The host-buffer entry points (page-locked numpy in, numpy out, chunked on three streams) for the suites with the ept /
    wet-bulb slots and for the hybrid-level suite (2-D copies of column chunks; the pressure field crosses PCIe in neither
    direction): bit for bit the device call, ragged chunk ends included."""
    from ek_thermo import fused, hostpipe
    from synthetic import IfsField

    n = 1_200_011
    inp = random_inputs(n, seed=37)
    hs = hostpipe.HostSuite(DEV, workspace_bytes=48 << 20, n_slots=3)  # small workspace -> many chunks
    host = {}
    for k in ("t", "q", "td", "p"):
        host[k] = hostpipe.pinned_empty(n, dtype)
        host[k][:] = inp[k]
    dev = {k: torch.from_numpy(np.array(v)).to(DEV) for k, v in host.items()}
    for em in ("ifs", "bolton39"):
        res = hs.suite_tqp(host["t"], host["q"], host["p"], outputs=fused.ALL7_TQP, ept_method=em)
        want = fused.suite_tqp(dev["t"], dev["q"], dev["p"], outputs=fused.ALL7_TQP, ept_method=em)
        for name in fused.ALL7_TQP:
            np.testing.assert_array_equal(res[name], want[name].cpu().numpy(), err_msg=f"{name} {em}")
    res = hs.suite_ttdp(host["t"], host["td"], host["p"], outputs=("ept", "wbpt"))
    want = fused.suite_ttdp(dev["t"], dev["td"], dev["p"], outputs=("ept", "wbpt"))
    for name in ("ept", "wbpt"):
        np.testing.assert_array_equal(res[name], want[name].cpu().numpy(), err_msg=name)
    # hybrid levels: [nlev, npl] host arrays, npl not a multiple of the chunk (ragged last chunk, non-vector tail)
    nlev, npl = 19, 70_003
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    field = IfsField("hybrid", npl, levels=nlev, seed=5, device=DEV)
    t_d, q_d, _ = field.materialise(0, nlev, tdt)
    sp_d = field.sp(0).to(tdt)
    A, B = field.A_half.astype(dtype), field.B_half.astype(dtype)
    want = fused.suite_tq_hybrid(t_d.reshape(nlev, npl), q_d.reshape(nlev, npl), sp_d, A, B, outputs=fused.ALL7_TQP)
    hs_small = hostpipe.HostSuite(DEV, workspace_bytes=24 << 20, n_slots=3)
    res = hs_small.suite_tq_hybrid(t_d.reshape(nlev, npl).cpu().numpy(), q_d.reshape(nlev, npl).cpu().numpy(), sp_d.cpu().numpy(), A, B,
                                   outputs=fused.ALL7_TQP)
    for name in fused.ALL7_TQP:
        assert res[name].shape == (nlev, npl)
        np.testing.assert_array_equal(res[name], want[name].cpu().numpy(), err_msg=name)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_batched_suite_equals_per_field_launches(ek, dtype):
    """One launch over a list of separately allocated levels (the per-level caller's layout) = one suite launch per level, bit
    for bit: compile-time and run-time output sets, every ept formulation, a broadcast pressure level, more fields than one
    launch's pointer table holds, a ragged field length (tile body + tail) and 16-byte-unaligned fields (scalar path)."""
    from ek_thermo import fused

    n_seg, n = 150, 256 * 4 * 3 + 77  # > 128 segments: two launches
    inp = random_inputs(n_seg * n + 1, seed=41)
    flat = {k: torch.from_numpy(inp[k]).to(DEV).to(dtype) for k in ("t", "q", "td", "p")}

    def fields(name, off=0):
        return [flat[name][off + j * n: off + (j + 1) * n].clone() if off == 0 else flat[name][off + j * n: off + (j + 1) * n] for j in range(n_seg)]

    def same(a, b):
        return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())

    ts, qs, tds, ps = (fields(k) for k in ("t", "q", "td", "p"))
    for outputs, em in ((fused.DEFAULT_TQP, "ifs"), (("theta", "rh"), "ifs"), (fused.ALL7_TQP, "ifs"), (tuple(fused.SUITE_TQP_OUTPUTS), "bolton39"), (("w", "ept"), "bolton35")):
        before = ek.launch_count()
        got = fused.suite_tqp_batch(ts, qs, ps, outputs=outputs, ept_method=em)
        assert ek.launch_count() == before + 2 and len(got) == n_seg
        for j in (0, 1, 127, 128, n_seg - 1):
            want = fused.suite_tqp(ts[j], qs[j], ps[j], outputs=outputs, ept_method=em)
            for name in outputs:
                assert same(got[j][name], want[name]), (outputs, em, j, name)
    got = fused.suite_ttdp_batch(ts, tds, 85000.0, outputs=fused.ALL7_TTDP)  # a pressure level: scalar operand
    for j in (0, 77, n_seg - 1):
        want = fused.suite_ttdp(ts[j], tds[j], 85000.0, outputs=fused.ALL7_TTDP)
        for name in fused.ALL7_TTDP:
            assert same(got[j][name], want[name]), (j, name)
    # pressure-level data: one pressure PER FIELD (150 levels: two launches, the second one starts at level 128)
    levels = [float(x) for x in np.linspace(100.0, 101325.0, n_seg)]
    levels[3] = 3.0  # p - es < 1e-4 on most points: the NaN rule
    ts[64][::97] = float("nan")  # missing values and zeros in a field: the exact recompute inside the batched kernel
    ts[127][5::101] = 0.0
    for outputs, em in ((("theta", "rh"), "ifs"), (fused.ALL7_TQP, "ifs"), (fused.ALL7_TQP, "bolton39")):
        got = fused.suite_tqp_batch(ts, qs, levels, outputs=outputs, ept_method=em)
        for j in (0, 3, 64, 127, 128, n_seg - 1):
            want = fused.suite_tqp(ts[j], qs[j], levels[j], outputs=outputs, ept_method=em)
            for name in outputs:
                assert same(got[j][name], want[name]), ("levels", outputs, em, j, name)
    got = fused.suite_ttdp_batch(ts, tds, levels)
    for j in (1, 128, n_seg - 1):
        want = fused.suite_ttdp(ts[j], tds[j], levels[j])
        for name in fused.DEFAULT_TTDP:
            assert same(got[j][name], want[name]), ("levels ttdp", j, name)
    with pytest.raises(ValueError):
        fused.suite_tqp_batch(ts, qs, levels[:-1])
    # views at an odd element offset: not 16-byte aligned -> the scalar load/store path of the same kernel
    tv, qv, pv = (fields(k, off=1) for k in ("t", "q", "p"))
    got = fused.suite_tqp_batch(tv, qv, pv)
    for j in (0, 5, n_seg - 1):
        want = fused.suite_tqp(tv[j], qv[j], pv[j])
        for name in fused.DEFAULT_TQP:
            assert same(got[j][name], want[name]), (j, name)
    assert fused.suite_tqp_batch([], [], []) == []
    with pytest.raises(ValueError):
        fused.suite_tqp_batch(ts, qs[:-1], ps)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n", [4099, 1_038_240, 6_599_680], ids=["tiny", "era5-level", "o1280-level"])
def test_back_to_back_launches_keep_stream_order(ek, dtype, n):
    """The streaming kernels are launched with programmatic stream serialization: a kernel may be scheduled while its
    predecessor in the stream drains, and orders itself with griddepcontrol.wait before it touches a field.  A chain of launches
    where each reads what the previous one wrote (RAW), overwrites what the previous one read (WAR) and rewrites its output
    (WAW) must give the bits of the same chain run with a device synchronisation between the launches; a torch kernel in
    the middle of the chain (launched without the attribute) and a CUDA-graph replay of the chain must too."""
    from ek_thermo import fused, thermo

    inp = random_inputs(n, seed=77)
    t0, q, p = (torch.from_numpy(inp[k]).to(DEV).to(dtype) for k in ("t", "q", "p"))

    def chain(sync, bufs):
        a, b = bufs
        a.copy_(t0)
        for i in range(12):
            # theta(a) -> b, then t_from_theta(b) back into a (a is read by the first launch and written by the second)
            fused.suite_tqp(a, q, p, outputs=("theta",), out={"theta": b})
            if sync:
                torch.cuda.synchronize()
            if i == 5:
                b.mul_(1.0)  # an ordinary kernel in the chain
            r = thermo.temperature_from_potential_temperature(b, p)
            if sync:
                torch.cuda.synchronize()
            if i % 3 == 0:
                a.copy_(r)
            else:
                fused.suite_tqp(r, q, p, outputs=("tv",), out={"tv": a})
            if sync:
                torch.cuda.synchronize()
        return a.clone()

    def same(x, y):
        return bool(((x == y) | (torch.isnan(x) & torch.isnan(y))).all())

    want = chain(True, (torch.empty_like(t0), torch.empty_like(t0)))
    for _ in range(3):
        assert same(chain(False, (torch.empty_like(t0), torch.empty_like(t0))), want)
    # the same chain captured once and replayed (programmatic edges inside a graph)
    bufs = (torch.empty_like(t0), torch.empty_like(t0))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain(False, bufs)  # warm-up on the capture stream (allocator, per-kernel attributes)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = chain(False, bufs)
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        assert same(res, want)


def test_sharded_run_equals_single_run(ek):
    """The partitioner's shards, run one by one on this GPU, reproduce the unsharded result exactly."""
    from ek_thermo import fused, partition

    n = 1_661_440 * 3 + 5
    inp = random_inputs(n, seed=29)
    t, q, p = (torch.from_numpy(inp[k]).to(DEV) for k in ("t", "q", "p"))
    whole = fused.suite_tqp(t, q, p)
    for world in (2, 8):
        parts = {k: [] for k in fused.DEFAULT_TQP}
        for r in range(world):
            b, e = partition.shard_range(n, world, r, align=2048)
            o = fused.suite_tqp(t[b:e], q[b:e], p[b:e])
            for k in parts:
                parts[k].append(o[k])
        for k in parts:
            cat = torch.cat(parts[k])
            assert torch.equal(torch.nan_to_num(cat), torch.nan_to_num(whole[k])), k


def test_exact_build_variant_passes_the_same_parity_cases():
    """libek_thermo_exact.so (libdevice math, IEEE division) is the A/B twin of the product library; run the
    float64 random + edge parity cases and the fixtures against it in a child process (one library per process)."""
    import os
    import subprocess
    import sys

    import ek_thermo

    exact = os.path.join(os.path.dirname(ek_thermo._backend.LIB_PATH), "libek_thermo_exact.so")
    if os.path.basename(ek_thermo._backend.LIB_PATH) == "libek_thermo_exact.so":
        pytest.skip("already running against the exact build")
    assert os.path.exists(exact), "build it with `make -C earthkit-meteo_b200/csrc exact`"
    env = dict(os.environ, EK_THERMO_LIB="libek_thermo_exact.so")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu", "-x", "-k",
                        "(oracle_random and f64) or (oracle_edge and f64) or fixtures or csv_goldens or fused_suites or bit_identical"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_exact_build_is_bit_identical_on_arithmetic_functions(ek):
    """The exact build (IEEE division, no fused multiply-add: -fmad=false) performs the reference's operations one by one, so
    every function without a transcendental returns numpy's bits -- random and special-value inputs, float64 and float32.
    Runs in the child process of test_exact_build_variant_passes_the_same_parity_cases (the product library replaces the
    division by a reciprocal and is held to 1e-12 instead)."""
    import os

    if os.path.basename(ek._backend.LIB_PATH) != "libek_thermo_exact.so":
        pytest.skip("only meaningful for the exact build")
    arithmetic = {"celsius_to_kelvin", "kelvin_to_celsius", "specific_humidity_from_mixing_ratio", "mixing_ratio_from_specific_humidity",
                  "vapour_pressure_from_specific_humidity", "vapour_pressure_from_mixing_ratio", "specific_humidity_from_vapour_pressure",
                  "mixing_ratio_from_vapour_pressure", "virtual_temperature", "specific_gas_constant"}
    checked = 0
    for case in CASES:
        if case.fn not in arithmetic:
            continue
        for dtype in (np.float64, np.float32):
            for inputs in (random_inputs(N_RANDOM, seed=5), edge_inputs(n=4099, seed=21)):
                with np.errstate(all="ignore"):
                    got, want, _ = _run_case(ek, case, inputs, dtype)
                for g, w in zip(got, want):
                    assert np.array_equal(g, np.asarray(w, dtype=dtype), equal_nan=True), (case.id, dtype.__name__)
                    checked += 1
    assert checked >= 4 * len(arithmetic)


def test_cuda_graph_capture_and_replay(ek):
    """Every entry point is a plain asynchronous launch on the caller's stream (no allocation, no sync inside the
    library), so a call sequence can be captured into a CUDA graph -- the way to run the launch-latency-bound
    ERA5-level case (BASELINE.json configs[0], 1 038 240 points) without per-call host overhead."""
    from ek_thermo import fused

    n = 721 * 1440
    inp = random_inputs(n, seed=41)
    t, q, p = (torch.from_numpy(inp[k]).to(DEV) for k in ("t", "q", "p"))
    out = {k: torch.empty_like(t) for k in ("theta", "rh")}
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fused.suite_tqp(t, q, p, outputs=("theta", "rh"), out=out)  # warm-up on the side stream
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fused.suite_tqp(t, q, p, outputs=("theta", "rh"), out=out)
    for o in out.values():
        o.zero_()
    t.add_(1.5)  # new input values in the captured buffers
    g.replay()
    torch.cuda.synchronize()
    tn = inp["t"] + 1.5
    np.testing.assert_allclose(out["theta"].cpu().numpy(), oracle.potential_temperature(tn, inp["p"]), rtol=1e-12)
    np.testing.assert_allclose(out["rh"].cpu().numpy(), oracle.relative_humidity_from_specific_humidity(tn, inp["q"], inp["p"]), rtol=1e-12)


def test_missing_values_stay_nan_and_neighbours_stay_exact(ek):
    """Masked fields: NaN in every input (a land/sea or below-ground mask), NaN in only one input, and NaN-free points
    in the same warps.  Outputs must be NaN exactly where the oracle's are, and exact elsewhere."""
    from ek_thermo import fused

    n = 300_007
    inp = random_inputs(n, seed=43)
    t, q, p = (inp[k].copy() for k in ("t", "q", "p"))
    rng = np.random.default_rng(43)
    blk = (np.arange(n) // 5000) % 3 == 0  # contiguous masked blocks: all inputs missing
    t[blk] = q[blk] = p[blk] = np.nan
    only_q = rng.random(n) < 0.05  # scattered: one input missing
    q[only_q] = np.nan
    d = [torch.from_numpy(x).to(DEV) for x in (t, q, p)]
    got = fused.suite_tqp(*d, outputs=tuple(fused.SUITE_TQP_OUTPUTS))
    with np.errstate(all="ignore"):
        want = oracle.suite_tqp(t, q, p)
    for name in fused.SUITE_TQP_OUTPUTS:
        g, w = got[name].cpu().numpy(), np.asarray(want[name])
        np.testing.assert_array_equal(np.isnan(g), np.isnan(w), err_msg=name)
        np.testing.assert_allclose(g, w, rtol=1e-12, atol=0, equal_nan=True, err_msg=name)
    assert np.isnan(got["theta"].cpu().numpy()[blk]).all() and not np.isnan(got["tv"].cpu().numpy()[~blk & ~only_q]).any()
    # scalar pressure with a masked t/q field
    got = fused.suite_tqp(d[0], d[1], 85000.0, outputs=("theta", "rh", "td"))
    with np.errstate(all="ignore"):
        np.testing.assert_allclose(got["theta"].cpu().numpy(), oracle.potential_temperature(t, 85000.0), rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(got["td"].cpu().numpy(), oracle.dewpoint_from_specific_humidity(q, 85000.0), rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
def test_tabulated_bisection_reproduces_the_reference_iterates(ek, ept_method):
    """The lean float64 bisection reads es(t) / ln t of the data-independent 4095-node tree from a table and takes the
    sign in log space (ek_thermo_formulas.inc: t_on_ma_bisect_tab); the iterate itself moves exactly as in the reference.
    On physical fields the quantised results must therefore be IDENTICAL to the oracle's (a differing point needs an
    exact tie at the 1e-16 level), at p and at p0, NaN positions included."""
    n = 400_000
    rng = np.random.default_rng(77)
    p = rng.uniform(2.0e4, 1.05e5, n)
    t = 288.15 * (p / 101325.0) ** 0.19 + rng.uniform(-15, 15, n)
    es = 611.21 * np.exp(17.502 * (t - 273.16) / (t - 32.19))
    q = np.minimum(rng.uniform(1e-6, 0.02, n), 0.95 * 0.621981 * es / (p - 0.378019 * es))
    d = [torch.from_numpy(x).to(DEV) for x in (t, q, p)]
    for fn in ("wet_bulb_temperature_from_specific_humidity", "wet_bulb_potential_temperature_from_specific_humidity"):
        got = getattr(ek.thermo, fn)(*d, ept_method=ept_method, t_method="bisect").cpu().numpy()
        with np.errstate(all="ignore"):
            want = getattr(oracle, fn)(t, q, p, ept_method=ept_method, t_method="bisect")
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        differing = int(np.sum(np.abs(got - want) > 0))
        assert differing <= 2, f"{fn}/{ept_method}: {differing} of {n} points differ"


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
def test_regrouped_newton_step_on_a_physical_field(ek, ept_method):
    """The lean float64 one-step Newton solve regroups the reference's arithmetic (shared reciprocals, folded exponentials;
    ek_thermo_formulas.inc: t_on_ma_newton_lean).  On 400 000 physical points it stays within 2e-10 relative of the oracle
    (measured: ifs / bolton39 < 1e-11, bolton35 3.2e-11; the bar for this solve is 2e-9 relative = 6e-7 K), with identical
    NaN positions, at p and at p0."""
    n = 400_000
    rng = np.random.default_rng(78)
    p = rng.uniform(2.0e4, 1.05e5, n)
    t = 288.15 * (p / 101325.0) ** 0.19 + rng.uniform(-15, 15, n)
    es = 611.21 * np.exp(17.502 * (t - 273.16) / (t - 32.19))
    q = np.minimum(rng.uniform(1e-6, 0.02, n), 0.95 * 0.621981 * es / (p - 0.378019 * es))
    d = [torch.from_numpy(x).to(DEV) for x in (t, q, p)]
    for fn in ("wet_bulb_temperature_from_specific_humidity", "wet_bulb_potential_temperature_from_specific_humidity"):
        got = getattr(ek.thermo, fn)(*d, ept_method=ept_method, t_method="newton").cpu().numpy()
        with np.errstate(all="ignore"):
            want = getattr(oracle, fn)(t, q, p, ept_method=ept_method, t_method="newton")
        np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
        fin = ~np.isnan(want)
        rel = np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin]))
        assert rel < 2e-10, f"{fn}/{ept_method}: max relative difference {rel:.3e}"


def test_lean_math_accuracy_in_ulps(ek):
    """The lean float64 primitives (table log/exp, Newton reciprocal) through the entry points that isolate them, on 2 M
    physical points against the oracle (numpy libm): the error stays at the few-ulp level, four orders of magnitude
    inside the 1e-12 parity bar.  The measured maxima are written to gpurun_out/lean_accuracy.json when that exists."""
    import json
    import os

    n = 1 << 21
    rng = np.random.default_rng(77)
    t = rng.uniform(190.0, 320.0, n)
    p = np.exp(rng.uniform(np.log(1.0), np.log(1.08e5), n))  # 1 Pa .. 1080 hPa, log-uniform: every exponent of log_
    q = rng.uniform(1e-7, 0.03, n)
    dev = {k: torch.from_numpy(v).to(DEV) for k, v in (("t", t), ("p", p), ("q", q))}
    th = ek.thermo
    pairs = {
        "theta (log, exp)": (th.potential_temperature(dev["t"], dev["p"]), oracle.potential_temperature(t, p)),
        "es water (rcp, exp)": (th.saturation_vapour_pressure(dev["t"], phase="water"), oracle.saturation_vapour_pressure(t, phase="water")),
        "es mixed": (th.saturation_vapour_pressure(dev["t"]), oracle.saturation_vapour_pressure(t)),
        "td from q (rcp, log, rcp)": (th.dewpoint_from_specific_humidity(dev["q"], dev["p"]), oracle.dewpoint_from_specific_humidity(q, p)),
        "e from q (rcp)": (th.vapour_pressure_from_specific_humidity(dev["q"], dev["p"]), oracle.vapour_pressure_from_specific_humidity(q, p)),
    }
    ulps = {}
    for name, (got, want) in pairs.items():
        g = got.cpu().numpy()
        ulps[name] = float(np.max(np.abs(g - want) / np.abs(want)) / 2.0**-52)
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "lean_accuracy.json"), "w") as f:
            json.dump({"lib": os.path.basename(ek._backend.LIB_PATH), "max_rel_err_in_2^-52": ulps}, f, indent=1)
    # theta's exponent kappa*ln(p0/p) reaches 3.3, so an absolute error of 1 ulp(ln p) shows up as a few ulp of theta
    assert ulps["theta (log, exp)"] < 32 and ulps["td from q (rcp, log, rcp)"] < 32, ulps
    assert ulps["es water (rcp, exp)"] < 64 and ulps["es mixed"] < 64 and ulps["e from q (rcp)"] < 8, ulps


def test_host_array_front_end_equals_device_path(ek):
    """ek_thermo.host (numpy in, numpy out): every case of the table, streamed through the device in
    ragged chunks on two streams, returns exactly what the device call returns (same kernels), as numpy arrays."""
    from ek_thermo import host

    inputs = random_inputs(N_RANDOM, seed=5)
    old = host.set_chunk_elements(10_000)  # 37 965 points -> 4 chunks, the last one ragged
    try:
        for case in CASES:
            args_np = [np.ascontiguousarray(inputs[a].astype(np.float64)) for a in case.args]
            got = getattr(host.thermo, case.fn)(*args_np, **case.kwargs)
            want = getattr(ek.thermo, case.fn)(*[torch.from_numpy(a).to(DEV) for a in args_np], **case.kwargs)
            got, want = (got, want) if isinstance(got, tuple) else ((got,), (want,))
            assert len(got) == len(want), case.id
            for g, w in zip(got, want):
                assert isinstance(g, np.ndarray) and g.dtype == np.float64 and g.shape == args_np[0].shape, case.id
                assert np.array_equal(g, w.cpu().numpy(), equal_nan=True), case.id
    finally:
        host.set_chunk_elements(old)


def test_host_array_front_end_numpy_semantics(ek):
    """Broadcasting, dtype promotion, scalars, lists, pinned buffers and the wind functions through ek_thermo.host."""
    from ek_thermo import host, hostpipe

    rng = np.random.default_rng(3)
    t = rng.uniform(220.0, 310.0, (5, 1, 257))
    p = np.array([1000.0, 5.0e4, 7.0e4, 8.5e4, 1.0e5]).reshape(5, 1, 1) * np.ones((1, 3, 1))
    with np.errstate(all="ignore"):
        got = host.thermo.potential_temperature(t, p)  # (5,1,257) x (5,3,1) -> (5,3,257)
        assert got.shape == (5, 3, 257)
        np.testing.assert_allclose(got, oracle.potential_temperature(t, p), rtol=1e-12)
        np.testing.assert_allclose(host.thermo.potential_temperature(t, 8.5e4), oracle.potential_temperature(t, 8.5e4), rtol=1e-12)
        # float32 stays float32 (a Python scalar does not up-cast), mixed float32 / float64 arrays promote to float64
        t32 = t.astype(np.float32)
        g32 = host.thermo.saturation_vapour_pressure(t32)
        assert g32.dtype == np.float32
        np.testing.assert_allclose(g32, oracle.saturation_vapour_pressure(t32), rtol=1e-5)
        assert host.thermo.potential_temperature(t32, p).dtype == np.float64
        # lists and integers, Python scalars only (numpy scalar out), zero-size input
        np.testing.assert_allclose(host.thermo.celsius_to_kelvin([0, 10, 20]), [273.16, 283.16, 293.16], rtol=1e-15)
        s = host.thermo.potential_temperature(280.0, 9.0e4)
        assert np.ndim(s) == 0 and abs(float(s) - float(oracle.potential_temperature(np.float64(280.0), np.float64(9.0e4)))) < 1e-10
        assert host.thermo.potential_temperature(np.empty((0, 4)), np.empty((0, 4))).shape == (0, 4)
        # two outputs, keyword options, errors as in the reference
        t1, td1, p1 = t[0, 0], t[0, 0] - 5.0, np.full(257, 9.0e4)
        tl, pl = host.thermo.lcl(t1, td1, p1, method="bolton")
        wl = oracle.lcl(t1, td1, p1, method="bolton")
        np.testing.assert_allclose(tl, wl[0], rtol=1e-12)
        np.testing.assert_allclose(pl, wl[1], rtol=1e-12)
        with pytest.raises(KeyError):
            host.thermo.ept_from_dewpoint(t1, td1, p1, method="nope")
    # page-locked arrays take the asynchronous copy path
    n = 300_000
    tp, pp = hostpipe.pinned_empty(n), hostpipe.pinned_empty(n)
    tp[:] = rng.uniform(220.0, 310.0, n)
    pp[:] = rng.uniform(1.0e4, 1.0e5, n)
    old = host.set_chunk_elements(70_000)
    try:
        np.testing.assert_allclose(host.thermo.potential_temperature(tp, pp), oracle.potential_temperature(tp, pp), rtol=1e-12)
    finally:
        host.set_chunk_elements(old)
    # wind
    u, v = rng.normal(0, 10, 1000), rng.normal(0, 10, 1000)
    np.testing.assert_allclose(host.wind.speed(u, v), np.hypot(u, v), rtol=1e-14)
    sp_, dir_ = host.wind.xy_to_polar(u, v)
    np.testing.assert_allclose(sp_, np.hypot(u, v), rtol=1e-14)
    assert dir_.shape == u.shape and np.all((dir_ >= 0) & (dir_ <= 360))
    # the device namespace still refuses host arrays: no silent CPU path anywhere
    with pytest.raises(TypeError):
        ek.thermo.potential_temperature(t, p)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_pressure_level_calls_equal_array_calls(ek, dtype):
    """Pressure-level data: t (and q / td) arrays, p ONE number.  The kernels run a tile loop specialised for that shape, in
    which everything that depends on p alone is computed once per thread; the results must be the bits of the same call
    with p materialised as a constant array -- single functions, suites (every ept formulation), the two-output ept kernel,
    the iterative solvers, ragged and unaligned fields, special values in t and a pressure that makes the NaN rule fire."""
    from ek_thermo import fused, thermo

    inp = random_inputs(N_RANDOM + 1, seed=91)
    t, q, td, ept = (torch.from_numpy(inp[k]).to(DEV).to(dtype) for k in ("t", "q", "td", "ept"))
    t[::1001] = float("nan")
    t[5::1777] = float("inf")
    t[7::1999] = 0.0

    def same(a, b, what):
        assert a.shape == b.shape and bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all()), what

    # every function of the case table that takes two or more fields: its last field as one number
    np_dtype = np.float64 if dtype == torch.float64 else np.float32
    for case in CASES:
        if len(case.args) < 2:
            continue
        arrs = [torch.from_numpy(inp[a].astype(np_dtype)).to(DEV) for a in case.args[:-1]]
        arrs[0][::1001] = float("nan")
        arrs[0][3::1013] = 0.0
        fn = getattr(thermo, case.fn)
        for s0 in (float(np.median(inp[case.args[-1]])), 3.0):
            got = fn(*arrs, s0, **case.kwargs)
            want = fn(*arrs, torch.full_like(arrs[0], s0), **case.kwargs)
            for k, (g, w) in enumerate(zip(got if isinstance(got, tuple) else (got,), want if isinstance(want, tuple) else (want,))):
                same(g, w, (case.id, s0, k))

    for off in (0, 1):  # 16-byte aligned fields, and views that start one element later (the scalar load/store path)
        tt, qq, dd, ee = t[off:], q[off:], td[off:], ept[off:]
        for p0 in (85000.0, 1.0e5, 3.0, 20000.0):  # 3 Pa: p - es < 1e-4 on most points (T:194)
            pa = torch.full_like(tt, p0)
            same(thermo.potential_temperature(tt, p0), thermo.potential_temperature(tt, pa), ("theta", p0, off))
            same(thermo.relative_humidity_from_specific_humidity(tt, qq, p0), thermo.relative_humidity_from_specific_humidity(tt, qq, pa), ("rh", p0, off))
            same(thermo.dewpoint_from_specific_humidity(qq, p0), thermo.dewpoint_from_specific_humidity(qq, pa), ("td", p0, off))
            same(thermo.saturation_specific_humidity(tt, p0), thermo.saturation_specific_humidity(tt, pa), ("qs", p0, off))
            same(thermo.ept_from_specific_humidity(tt, qq, p0), thermo.ept_from_specific_humidity(tt, qq, pa), ("ept", p0, off))
            same(thermo.ept_from_dewpoint(tt, dd, p0, method="bolton39"), thermo.ept_from_dewpoint(tt, dd, pa, method="bolton39"), ("ept39", p0, off))
            for tm in ("bisect", "newton"):
                same(thermo.temperature_on_moist_adiabat(ee, p0, t_method=tm), thermo.temperature_on_moist_adiabat(ee, pa, t_method=tm), (tm, p0, off))
            for em in ("ifs", "bolton35", "bolton39"):
                for outputs in (fused.DEFAULT_TQP, ("theta", "rh"), fused.ALL7_TQP, tuple(fused.SUITE_TQP_OUTPUTS)):
                    a, b = fused.suite_tqp(tt, qq, p0, outputs=outputs, ept_method=em), fused.suite_tqp(tt, qq, pa, outputs=outputs, ept_method=em)
                    for name in outputs:
                        same(a[name], b[name], ("suite_tqp", em, name, p0, off))
                a, b = fused.suite_ttdp(tt, dd, p0, outputs=fused.ALL7_TTDP, ept_method=em), fused.suite_ttdp(tt, dd, pa, outputs=fused.ALL7_TTDP, ept_method=em)
                for name in fused.ALL7_TTDP:
                    same(a[name], b[name], ("suite_ttdp", em, name, p0, off))
                for x, y in zip(fused.ept_wet_bulb(tt, qq, p0, humidity="q", ept_method=em), fused.ept_wet_bulb(tt, qq, pa, humidity="q", ept_method=em)):
                    same(x, y, ("ept_wet_bulb", em, p0, off))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_fused_results_do_not_depend_on_position(ek, dtype):
    """A point's result is the same bits whether it is computed in the vector body of a tile or in the scalar tail, in a
    whole-field launch or in a launch over a piece of the field: the fused suites and the ept / wet-bulb kernel on
    pieces (ragged, 16-byte aligned and unaligned starts) equal the whole-field launch exactly."""
    from ek_thermo import fused

    inp = random_inputs(N_RANDOM, seed=11)
    t, q, p, td = (torch.from_numpy(inp[k]).to(DEV).to(dtype) for k in ("t", "q", "p", "td"))
    cuts = [0, 10_000, 10_003, 21_111, N_RANDOM]

    def pieces(fn):
        parts = [fn(slice(b, e)) for b, e in zip(cuts[:-1], cuts[1:])]
        return [torch.cat([pt[k] for pt in parts]) for k in range(len(parts[0]))]

    def same(whole, parts, what):
        for k, (w, s) in enumerate(zip(whole, parts)):
            assert torch.equal(torch.nan_to_num(w, nan=-1.0), torch.nan_to_num(s, nan=-1.0)), (what, k)

    everything = tuple(fused.SUITE_TQP_OUTPUTS)
    same(list(fused.suite_tqp(t, q, p, outputs=everything).values()),
         pieces(lambda s: list(fused.suite_tqp(t[s], q[s], p[s], outputs=everything).values())), "suite_tqp")
    everything = tuple(fused.SUITE_TTDP_OUTPUTS)
    same(list(fused.suite_ttdp(t, td, p, outputs=everything).values()),
         pieces(lambda s: list(fused.suite_ttdp(t[s], td[s], p[s], outputs=everything).values())), "suite_ttdp")
    for em in ("ifs", "bolton35", "bolton39"):
        for tm in ("direct", "bisect", "newton"):
            for potential in (True, False):
                if tm == "direct" and not potential:
                    continue
                kw = dict(humidity="q", ept_method=em, t_method=tm, potential=potential)
                same(list(fused.ept_wet_bulb(t, q, p, **kw)), pieces(lambda s: list(fused.ept_wet_bulb(t[s], q[s], p[s], **kw))), (em, tm, potential))


def test_more_than_2_to_31_points_in_one_launch(ek):
    """Maximum sizes: indexing is 64-bit end to end.  One launch over 2^31 + 4099 float32 points (8.6 GB in, 8.6 GB out),
    once through a one-input kernel and once through a two-input kernel with a broadcast scalar; the points beyond the
    32-bit boundary, the ragged tail and a strided sample are checked against the oracle."""
    free, _ = torch.cuda.mem_get_info()
    n = (1 << 31) + 4099
    if free < 3 * 4 * n + (1 << 30):
        pytest.skip("not enough free device memory for a 2^31-point field")
    t = torch.empty(n, dtype=torch.float32, device=DEV)
    t[: 1 << 20] = torch.linspace(200.0, 320.0, 1 << 20, device=DEV)
    t[1 << 20:] = 250.0
    t[-4099:] = torch.linspace(210.0, 310.0, 4099, device=DEV)
    t[(1 << 31) - 5: (1 << 31) + 5] = torch.arange(10, device=DEV, dtype=torch.float32) + 270.0
    k = ek.thermo.kelvin_to_celsius(t)
    th = ek.thermo.potential_temperature(t, 85000.0)
    torch.cuda.synchronize()
    idx = torch.cat([torch.arange(0, 4096, device=DEV), torch.arange((1 << 31) - 8, n, device=DEV), torch.arange(0, n, 104729, device=DEV)])
    tin = t[idx].cpu().numpy()
    np.testing.assert_array_equal(k[idx].cpu().numpy(), oracle.kelvin_to_celsius(tin))
    np.testing.assert_allclose(th[idx].cpu().numpy(), oracle.potential_temperature(tin, np.float32(85000.0)), rtol=2e-6)
    assert k.shape == t.shape and th.shape == t.shape
    del t, k, th
    torch.cuda.empty_cache()


def test_host_array_staged_pipeline_equals_device_path(ek):
    """Large pageable arrays go through page-locked staging buffers filled and drained by worker threads (three slots in
    flight).  With the thresholds lowered so that 37 965 points make four ragged chunks: same bits as the device call, for
    one- and two-output functions, float64 and float32, a page-locked input among pageable ones, and repeated calls that
    reuse the staging buffers."""
    from ek_thermo import host, hostpipe

    inputs = random_inputs(N_RANDOM, seed=9)
    old = (host._STAGE_MIN, host._STAGE_CHUNK)
    host._STAGE_MIN, host._STAGE_CHUNK = 20_000, 10_000
    host.release_staging()
    try:
        for dtype in (np.float64, np.float32):
            t, td, q, p = (np.ascontiguousarray(inputs[k].astype(dtype)) for k in ("t", "td", "q", "p"))
            tp = hostpipe.pinned_empty(t.size, dtype)
            tp[:] = t
            d = {k: torch.from_numpy(v).to(DEV) for k, v in (("t", t), ("td", td), ("q", q), ("p", p))}
            for _ in range(2):
                got = host.thermo.relative_humidity_from_specific_humidity(tp, q, p)  # pinned t, pageable q and p
                want = ek.thermo.relative_humidity_from_specific_humidity(d["t"], d["q"], d["p"])
                assert got.dtype == dtype and np.array_equal(got, want.cpu().numpy(), equal_nan=True)
                gl = host.thermo.lcl(t, td, p)
                wl = ek.thermo.lcl(d["t"], d["td"], d["p"])
                assert all(np.array_equal(g, w.cpu().numpy(), equal_nan=True) for g, w in zip(gl, wl))
                gw = host.thermo.wet_bulb_temperature_from_specific_humidity(t, q, 85000.0)  # scalar operand, bisection
                ww = ek.thermo.wet_bulb_temperature_from_specific_humidity(d["t"], d["q"], 85000.0)
                assert np.array_equal(gw, ww.cpu().numpy(), equal_nan=True)
    finally:
        host._STAGE_MIN, host._STAGE_CHUNK = old
        host.release_staging()


def test_host_functions_under_apply_ufunc_semantics(ek):
    """the reference's own high-level idiom is ``xr.apply_ufunc(potential_temperature, t, p)`` (reference
    tests/vertical/test_xr_theta.py:34).  xarray is not in this image, so the test hands ``host.thermo`` exactly what
    apply_ufunc hands a function, with plain numpy stand-ins: the ``.data`` of DataArrays as read-only arrays (netCDF-backed
    variables), non-contiguous views (after ``transpose`` / ``isel``), broadcast dimensions of length 1, float32 variables,
    0-d arrays (a reduced dimension), per-element calls (``vectorize=True``: numpy scalars in, one value out), and
    ``dask="parallelized"`` blocks mapped from a thread pool.  Results: a NEW array of the broadcast shape and the input
    dtype, inputs untouched, values those of the oracle."""
    from concurrent.futures import ThreadPoolExecutor

    from ek_thermo import host

    rng = np.random.default_rng(77)
    lev, lat, lon = 7, 37, 53
    t = rng.uniform(210.0, 310.0, (lev, lat, lon))
    p = rng.uniform(2.0e4, 1.02e5, (lev, lat, lon))
    q = rng.uniform(1e-6, 0.015, (lev, lat, lon))
    want = oracle.potential_temperature(t, p)
    fn = host.thermo.potential_temperature
    # read-only .data (a variable opened from a file), result is a fresh writable array
    t_ro, p_ro = t.copy(), p.copy()
    t_ro.flags.writeable = False
    p_ro.flags.writeable = False
    got = fn(t_ro, p_ro)
    assert got.shape == t.shape and got.dtype == np.float64 and got.flags.writeable and not np.shares_memory(got, t_ro)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    assert np.array_equal(t_ro, t) and np.array_equal(p_ro, p)
    # non-contiguous views: transposed (lat, lon, lev) and strided selections
    got = fn(np.transpose(t, (1, 2, 0)), np.transpose(p, (1, 2, 0)))
    np.testing.assert_allclose(got, np.transpose(want, (1, 2, 0)), rtol=1e-12)
    np.testing.assert_allclose(fn(t[::2, 3:, ::5], p[::2, 3:, ::5]), want[::2, 3:, ::5], rtol=1e-12)
    # a pressure coordinate broadcast against the field: (lev, 1, 1) with (lev, lat, lon), and the other way round
    p_lev = np.linspace(2.0e4, 1.0e5, lev).reshape(lev, 1, 1)
    np.testing.assert_allclose(fn(t, p_lev), oracle.potential_temperature(t, p_lev), rtol=1e-12)
    np.testing.assert_allclose(fn(t[:, :1, :1], p), oracle.potential_temperature(t[:, :1, :1], p), rtol=1e-12)
    # float32 variables stay float32; a float32 field with a float64 coordinate promotes like numpy
    t32, p32 = t.astype(np.float32), p.astype(np.float32)
    g32 = fn(t32, p32)
    assert g32.dtype == np.float32
    np.testing.assert_allclose(g32, oracle.potential_temperature(t32, p32), rtol=1e-5)
    assert fn(t32, p_lev).dtype == np.float64
    # 0-d arrays and numpy scalars (vectorize=True calls the function once per element)
    s0 = fn(np.asarray(t[0, 0, 0]), np.asarray(p[0, 0, 0]))
    assert np.ndim(s0) == 0 and abs(float(s0) - float(want[0, 0, 0])) <= 1e-12 * float(want[0, 0, 0])
    vec = np.vectorize(fn)(t[0, :3, :4], p[0, :3, :4])
    np.testing.assert_allclose(vec, want[0, :3, :4], rtol=1e-12)
    # dask="parallelized": blocks along the first dimension evaluated from a thread pool, concatenated by the caller
    with ThreadPoolExecutor(4) as pool:
        blocks = list(pool.map(lambda k: host.thermo.relative_humidity_from_specific_humidity(t[k], q[k], p[k]), range(lev)))
    np.testing.assert_allclose(np.stack(blocks), oracle.relative_humidity_from_specific_humidity(t, q, p), rtol=1e-12)
    # keyword options and multiple outputs pass through (apply_ufunc(..., kwargs=..., output_core_dims=[[], []]))
    tl, pl = host.thermo.lcl(t, t - 4.0, p, method="bolton")
    wl = oracle.lcl(t, t - 4.0, p, method="bolton")
    np.testing.assert_allclose(tl, wl[0], rtol=1e-12)
    np.testing.assert_allclose(pl, wl[1], rtol=1e-12)


def test_host_array_pipeline_is_reentrant(ek):
    """dask's threaded scheduler and xr.apply_ufunc call the host functions from several threads at once: every running
    call owns its staging buffers (checked out under a lock), so concurrent calls -- and a release_staging() in the middle
    of them -- return the same bits as the device path.  Both result flavours (page-locked and pageable arrays)."""
    import threading

    from ek_thermo import host

    inputs = random_inputs(N_RANDOM, seed=21)
    t, td, q, p = (np.ascontiguousarray(inputs[k]) for k in ("t", "td", "q", "p"))
    d = {k: torch.from_numpy(v).to(DEV) for k, v in (("t", t), ("td", td), ("q", q), ("p", p))}
    want = {
        "theta": ek.thermo.potential_temperature(d["t"], d["p"]).cpu().numpy(),
        "rh": ek.thermo.relative_humidity_from_specific_humidity(d["t"], d["q"], d["p"]).cpu().numpy(),
        "td": ek.thermo.dewpoint_from_specific_humidity(d["q"], d["p"]).cpu().numpy(),
        "ept": ek.thermo.ept_from_dewpoint(d["t"], d["td"], d["p"]).cpu().numpy(),
    }
    calls = {
        "theta": lambda: host.thermo.potential_temperature(t, p),
        "rh": lambda: host.thermo.relative_humidity_from_specific_humidity(t, q, p),
        "td": lambda: host.thermo.dewpoint_from_specific_humidity(q, p),
        "ept": lambda: host.thermo.ept_from_dewpoint(t, td, p),
    }
    old = (host._STAGE_MIN, host._STAGE_CHUNK)
    host._STAGE_MIN, host._STAGE_CHUNK = 20_000, 5_000
    host.release_staging()
    bad = []

    def worker(name, rounds):
        try:
            for r in range(rounds):
                got = calls[name]()
                if not np.array_equal(got, want[name], equal_nan=True):
                    bad.append((name, r))
                if name == "td" and r == 2:
                    host.release_staging()  # while the other threads are inside their calls
        except Exception as exc:  # noqa: BLE001 -- reported through `bad`
            bad.append((name, repr(exc)))

    try:
        for pinned_results in (True, False):
            prev = host.set_pinned_results(pinned_results)
            try:
                threads = [threading.Thread(target=worker, args=(name, 6)) for name in calls for _ in range(2)]
                for th in threads:
                    th.start()
                for th in threads:
                    th.join()
            finally:
                host.set_pinned_results(prev)
            assert not bad, bad
    finally:
        host._STAGE_MIN, host._STAGE_CHUNK = old
        host.release_staging()


def test_host_array_fused_kernels(ek):
    """host.fused: the fused suites and the ept / wet-bulb kernel over numpy arrays, one pass over PCIe for all fields;
    equal to the device call bit for bit, direct and staged pipelines."""
    from ek_thermo import fused, host

    inputs = random_inputs(N_RANDOM, seed=13)
    t, td, q, p = (np.ascontiguousarray(inputs[k]) for k in ("t", "td", "q", "p"))
    d = {k: torch.from_numpy(v).to(DEV) for k, v in (("t", t), ("td", td), ("q", q), ("p", p))}
    old = (host._STAGE_MIN, host._STAGE_CHUNK)
    try:
        for stage_min in (old[0], 20_000):
            host._STAGE_MIN, host._STAGE_CHUNK = stage_min, 10_000
            host.release_staging()
            got = host.fused.suite_tqp(t, q, p, outputs=("theta", "rh", "td", "thetav"))
            want = fused.suite_tqp(d["t"], d["q"], d["p"], outputs=("theta", "rh", "td", "thetav"))
            assert list(got) == ["theta", "rh", "td", "thetav"]
            for k in got:
                assert np.array_equal(got[k], want[k].cpu().numpy(), equal_nan=True), k
            got = host.fused.suite_ttdp(t, td, p)
            want = fused.suite_ttdp(d["t"], d["td"], d["p"])
            for k in want:
                assert np.array_equal(got[k], want[k].cpu().numpy(), equal_nan=True), k
            ge, gw = host.fused.ept_wet_bulb(t, q, p, t_method="bisect", potential=False)
            we, ww = fused.ept_wet_bulb(d["t"], d["q"], d["p"], t_method="bisect", potential=False)
            assert np.array_equal(ge, we.cpu().numpy(), equal_nan=True) and np.array_equal(gw, ww.cpu().numpy(), equal_nan=True)
    finally:
        host._STAGE_MIN, host._STAGE_CHUNK = old
        host.release_staging()
