"""The oracle against (a) the live-reference fixtures and (b) the reference's own golden CSVs.

(a) is bit-exact: the oracle keeps the reference's numpy operation order, so on the same numpy it
must reproduce the reference's outputs exactly (float64 and float32, NaN/inf positions included).
(b) uses the tolerances of the reference's own tests (tests/thermo/test_thermo.py): default
allclose (rtol 1e-5, atol 1e-8) and rtol=1e-3, atol=0 for the iterative wet-bulb columns
(TT:802-849).  Plus the reference's inline known-answer vectors.
"""
import numpy as np
import pytest

import thermo_oracle as oracle
from cases import CASES
from kat import KATS


def _call(case, inputs, dtype):
    args = [np.ascontiguousarray(inputs[a].astype(dtype)) for a in case.args]
    res = getattr(oracle, case.fn)(*args, **case.kwargs)
    return res if isinstance(res, tuple) else (res,)


def _inputs(ref_live, sname):
    pre = f"in/{sname}/"
    return {k[len(pre):]: v for k, v in ref_live.items() if k.startswith(pre)}


@pytest.mark.parametrize("sname,dname", [("rand", "float64"), ("edge", "float64"), ("grid", "float64"),
                                         ("ma", "float64"), ("rand", "float32"), ("edge", "float32")])
def test_oracle_bit_identical_to_live_reference(ref_live, sname, dname):
    inputs = _inputs(ref_live, sname)
    n = 0
    for case in CASES:
        if f"out/{sname}/{dname}/{case.id}/0" not in ref_live:
            continue
        got = _call(case, inputs, np.dtype(dname))
        for k, g in enumerate(got):
            want = ref_live[f"out/{sname}/{dname}/{case.id}/{k}"]
            g = np.asarray(g)
            assert g.dtype == want.dtype, case.id
            # bit-exact incl. NaN positions; numpy's SIMD exp/log may differ by <=2 ulp between
            # CPU generations (AVX512 vs AVX2 code paths), so allow that and nothing more
            np.testing.assert_array_equal(np.isnan(g), np.isnan(want), err_msg=case.id)
            tol = 8 * np.finfo(want.dtype).eps
            if case.iterative == "bisect":
                # a sign flip at an exact tie moves the iterate by 2*dt_k; only possible if libm differs
                tol = 1e-3
            np.testing.assert_allclose(g, want, rtol=tol, atol=0, equal_nan=True, err_msg=case.id)
            n += 1
    assert n > 0


# ---------------------------------------------------------------------------------------------
# (b) the reference's golden CSVs, consumed the way the reference's tests consume them
# ---------------------------------------------------------------------------------------------
def _grid(ref_csv):
    return {k: ref_csv[f"t_hum_p_data/{k}"] for k in ("t", "td", "r", "q", "p")}


def test_csv_saturation_vapour_pressure(ref_csv):  # TT:169-186, 346-356
    t = ref_csv["sat_vp/t"]
    for ph in ("mixed", "water", "ice"):
        np.testing.assert_allclose(oracle.saturation_vapour_pressure(t, phase=ph), ref_csv[f"sat_vp/{ph}"], rtol=1e-12)
        np.testing.assert_allclose(
            oracle.saturation_vapour_pressure_slope(ref_csv["sat_vp_slope/t"], phase=ph), ref_csv[f"sat_vp_slope/{ph}"], rtol=1e-12
        )


@pytest.mark.parametrize("stem,fn", [("sat_mr", "saturation_mixing_ratio"), ("sat_q", "saturation_specific_humidity"),
                                     ("sat_mr_slope", "saturation_mixing_ratio_slope"),
                                     ("sat_q_slope", "saturation_specific_humidity_slope")])
def test_csv_saturation_humidity(ref_csv, stem, fn):  # TT:206-331
    t, p = ref_csv[f"{stem}/t"], ref_csv[f"{stem}/p"]
    for ph in ("mixed", "water", "ice"):
        np.testing.assert_allclose(getattr(oracle, fn)(t, p, phase=ph), ref_csv[f"{stem}/{ph}"], rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("method", ["ifs", "bolton35", "bolton39"])
def test_csv_ept(ref_csv, method):  # TT:706-747
    g = _grid(ref_csv)
    np.testing.assert_allclose(oracle.ept_from_dewpoint(g["t"], g["td"], g["p"], method=method), ref_csv[f"eqpt/{method}_td"], rtol=1e-12)
    np.testing.assert_allclose(oracle.ept_from_specific_humidity(g["t"], g["q"], g["p"], method=method), ref_csv[f"eqpt/{method}_q"], rtol=1e-12)
    np.testing.assert_allclose(oracle.saturation_ept(g["t"], g["p"], method=method), ref_csv[f"seqpt/{method}"], rtol=1e-12)


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
@pytest.mark.parametrize("t_method", ["bisect", "newton"])
def test_csv_t_on_moist_adiabat(ref_csv, ept_method, t_method):  # TT:750-773
    got = oracle.temperature_on_moist_adiabat(ref_csv["t_on_most_adiabat/ept"], ref_csv["t_on_most_adiabat/p"],
                                              ept_method=ept_method, t_method=t_method)
    np.testing.assert_allclose(got, ref_csv[f"t_on_most_adiabat/{ept_method}_{t_method}"], rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
@pytest.mark.parametrize("t_method", ["bisect", "newton", "direct"])
def test_csv_wet_bulb(ref_csv, ept_method, t_method):  # TT:776-849 (rtol 1e-3, atol 0 as in the reference)
    g = _grid(ref_csv)
    kw = dict(ept_method=ept_method, t_method=t_method)
    # the tight bound holds except for the bisect sign-tie columns
    tight = 1e-12 if t_method != "bisect" else 1e-3
    got = oracle.wet_bulb_potential_temperature_from_dewpoint(g["t"], g["td"], g["p"], **kw)
    np.testing.assert_allclose(got, ref_csv[f"t_wetpt/{ept_method}_{t_method}_td"], rtol=tight, atol=0, equal_nan=True)
    got = oracle.wet_bulb_potential_temperature_from_specific_humidity(g["t"], g["q"], g["p"], **kw)
    np.testing.assert_allclose(got, ref_csv[f"t_wetpt/{ept_method}_{t_method}_q"], rtol=tight, atol=0, equal_nan=True)
    if t_method != "direct":
        got = oracle.wet_bulb_temperature_from_dewpoint(g["t"], g["td"], g["p"], **kw)
        np.testing.assert_allclose(got, ref_csv[f"t_wet/{ept_method}_{t_method}_td"], rtol=tight, atol=0, equal_nan=True)
        got = oracle.wet_bulb_temperature_from_specific_humidity(g["t"], g["q"], g["p"], **kw)
        np.testing.assert_allclose(got, ref_csv[f"t_wet/{ept_method}_{t_method}_q"], rtol=tight, atol=0, equal_nan=True)


# ---------------------------------------------------------------------------------------------
# inline known-answer vectors of the reference's tests (values quoted from TT, not computed here)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kat", KATS, ids=[f"{i}-{k[0]}" for i, k in enumerate(KATS)])
def test_kat_reference_numbers(kat):
    fn, args, kwargs, expected, rtol = kat
    got = getattr(oracle, fn)(*[np.asarray(a, dtype=np.float64) for a in args], **kwargs)
    if not isinstance(got, tuple):
        got, expected = (got,), (expected,)
    for g, e in zip(got, expected):
        np.testing.assert_allclose(g, np.asarray(e, dtype=np.float64), rtol=rtol, atol=1e-8, equal_nan=True)


def test_kat_nan_rules():
    # es = 0 -> NaN temperature (TT:359-371); q = 0 -> NaN dewpoint (TT:546-548); r = 0 -> NaN (TT:520-522)
    assert np.isnan(oracle.temperature_from_saturation_vapour_pressure(np.array([0.0])))[0]
    assert np.isnan(oracle.dewpoint_from_specific_humidity(np.array([0.0]), np.array([1e5])))[0]
    assert np.isnan(oracle.dewpoint_from_relative_humidity(np.array([290.0]), np.array([0.0])))[0]
    # p - e < eps -> NaN (T:194, T:231)
    assert np.isnan(oracle.specific_humidity_from_vapour_pressure(np.array([5e4]), np.array([5e4])))[0]
    assert np.isnan(oracle.mixing_ratio_from_vapour_pressure(np.array([5e4]), np.array([5e4])))[0]


def test_error_conventions():
    t = np.array([280.0])
    with pytest.raises(ValueError):
        oracle.specific_humidity_from_vapour_pressure(t, t, eps=0)
    with pytest.raises(ValueError):
        oracle.lcl_temperature(t, t, method="x")
    with pytest.raises(KeyError):
        oracle.ept_from_dewpoint(t, t, t, method="x")
    with pytest.raises(ValueError):
        oracle.temperature_on_moist_adiabat(t, t, t_method="x")
    assert oracle.saturation_vapour_pressure(t, phase="x") is None
