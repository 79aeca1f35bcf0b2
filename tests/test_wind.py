"""Wind: oracle vs the live-reference fixtures and the reference tests' known-answer vectors (CPU);
the package's host logic + the g++ build of the functors vs the oracle (CPU, mock device); the CUDA path (GPU)."""
import os

import numpy as np
import pytest
import torch

import hostmath_backend
import wind_oracle as woracle

nan = np.nan
CASES = [
    ("speed", ("u", "v"), {}), ("direction", ("u", "v"), {"convention": "meteo"}), ("direction", ("u", "v"), {"convention": "polar"}),
    ("direction", ("u", "v"), {"convention": "polar", "to_positive": False}), ("xy_to_polar", ("u", "v"), {"convention": "meteo"}),
    ("xy_to_polar", ("u", "v"), {"convention": "polar"}), ("polar_to_xy", ("mag", "dir"), {"convention": "meteo"}),
    ("polar_to_xy", ("mag", "dir"), {"convention": "polar"}), ("w_from_omega", ("omega", "t", "p"), {}), ("coriolis", ("lat",), {}),
]
U = [0, 1, 1, 1, 0, -1, -1, -1, 0, nan, 1, nan]
V = [1, 1, 0, -1, -1, -1, 0, 1, 0, 1, nan, nan]
SPD = [1.0, 1.4142135624, 1.0, 1.4142135624, 1.0, 1.4142135624, 1.0, 1.4142135624, 0.0, nan, nan, nan]
# known-answer vectors of the reference's tests/wind/test_wind.py (:14-140)
KATS = [
    ("speed", (U, V), {}, (SPD,)),
    ("direction", (U, V), {"convention": "meteo"}, ([180.0, 225, 270, 315, 0, 45, 90, 135, 270, nan, nan, nan],)),
    ("direction", (U, V), {"convention": "polar"}, ([90, 45, 0, 315, 270, 225, 180, 135.0, 0, nan, nan, nan],)),
    ("direction", (U, V), {"convention": "polar", "to_positive": False}, ([90, 45, 0, -45, -90, -135, 180, 135, 0, nan, nan, nan],)),
    ("xy_to_polar", (U, V), {}, (SPD, [180.0, 225, 270, 315, 0, 45, 90, 135, 270, nan, nan, nan])),
]


def case_id(fn, kw):
    return fn + "".join(f";{k}={v}" for k, v in sorted(kw.items()))


@pytest.fixture(scope="module")
def fx():
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wind.npz")) as z:
        return {k: z[k] for k in z.files}


def _tuple(x):
    return x if isinstance(x, tuple) else (x,)


@pytest.mark.parametrize("dname", ["float64", "float32"])
def test_oracle_matches_live_reference(fx, dname):
    """Same numpy calls in the same order as the reference, so the oracle was bit-identical to it where the fixtures were
    generated (tests/golden/PINNING.json: worst difference 0.0).  The bound here is 8 ulp, not 0: numpy's SIMD
    sin / cos / arctan2 / hypot kernels differ in the last bits between CPU generations, and this test also runs on the
    GPU box's host."""
    dt = np.dtype(dname).type
    for fn, args, kw in CASES:
        with np.errstate(all="ignore"):
            got = _tuple(getattr(woracle, fn)(*[fx[f"in/{a}"].astype(dt) for a in args], **kw))
        for k, g in enumerate(got):
            want = fx[f"out/{dname}/{case_id(fn, kw)}/{k}"]
            np.testing.assert_allclose(np.asarray(g), want, rtol=8 * np.finfo(want.dtype).eps, atol=0, equal_nan=True, err_msg=case_id(fn, kw))


@pytest.mark.parametrize("kat", KATS, ids=[case_id(k[0], k[2]) for k in KATS])
def test_oracle_reference_known_answers(kat):
    fn, args, kw, want = kat
    with np.errstate(all="ignore"):
        got = _tuple(getattr(woracle, fn)(*[np.asarray(a, dtype=np.float64) for a in args], **kw))
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, np.asarray(w, dtype=np.float64), rtol=1e-5, atol=1e-8, equal_nan=True)


def _check(wind, fx, dt, make):
    f32 = dt == np.float32
    for fn, args, kw in CASES:
        a_np = [fx[f"in/{a}"].astype(dt) for a in args]
        got = _tuple(getattr(wind, fn)(*[make(a) for a in a_np], **kw))
        with np.errstate(all="ignore"):
            want = _tuple(getattr(woracle, fn)(*a_np, **kw))
        for g, w in zip(got, want):
            g = g.cpu().numpy().astype(np.float64)
            w = np.asarray(w).astype(np.float64)
            np.testing.assert_array_equal(np.isnan(g), np.isnan(w), err_msg=case_id(fn, kw))
            # angles near 0/360 and components near 0 are differences of large numbers: absolute floor
            np.testing.assert_allclose(g, w, rtol=2e-5 if f32 else 1e-12, atol=2e-4 if f32 else 1e-12, equal_nan=True, err_msg=case_id(fn, kw))


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_functors_and_host_logic_match_oracle(monkeypatch, fx, dt):
    import ek_thermo

    hostmath_backend.install(monkeypatch)
    _check(ek_thermo.wind, fx, dt, lambda a: torch.from_numpy(a.copy()))
    t = torch.ones(3, dtype=torch.float64)
    with pytest.raises(ValueError):
        ek_thermo.wind.direction(t, t, convention="nope")
    with pytest.raises(ValueError):
        ek_thermo.wind.polar_to_xy(t, t, convention="nope")


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_cuda_matches_oracle_and_known_answers(fx, dt):
    import ek_thermo

    _check(ek_thermo.wind, fx, dt, lambda a: torch.from_numpy(a).to("cuda:0"))
    for fn, args, kw, want in KATS:
        got = _tuple(getattr(ek_thermo.wind, fn)(*[torch.tensor(a, dtype=torch.float64, device="cuda:0") for a in args], **kw))
        for g, w in zip(got, want):
            np.testing.assert_allclose(g.cpu().numpy(), np.asarray(w, dtype=np.float64), rtol=1e-5, atol=1e-8, equal_nan=True)
    # polar <-> xy round trip on a large field
    g = torch.Generator(device="cuda:0").manual_seed(4)
    u = torch.empty(5_000_011, device="cuda:0", dtype=torch.float64).normal_(0, 10, generator=g)
    v = torch.empty_like(u).normal_(0, 10, generator=g)
    for conv in ("meteo", "polar"):
        s, d = ek_thermo.wind.xy_to_polar(u, v, convention=conv)
        x, y = ek_thermo.wind.polar_to_xy(s, d, convention=conv)
        assert float((x - u).abs().max()) < 1e-11 and float((y - v).abs().max()) < 1e-11
