"""CPU check of everything except the GPU itself: the package's host logic + the g++ build of the
kernel functors (mock device, see hostmath_backend.py) against the oracle on the shared case table.

Tolerances: float64 closed forms 1e-12 relative (the build's bar, BASELINE.json north_star), NaN and
inf positions identical; float32 1e-5 relative.  Bisect results are quantised to 0.0293 K steps and a
1-ulp libm difference can flip an exact sign tie: mismatching points are counted
and must stay below 0.5 %, all others must agree to 1e-12.
"""
import numpy as np
import pytest
import torch

import hostmath_backend
import thermo_oracle as oracle
from cases import CASES, edge_inputs, random_inputs
from compare import compare, conditioning, reference_f32_noise
from kat import KATS


@pytest.fixture()
def thermo(monkeypatch):
    import ek_thermo

    hostmath_backend.install(monkeypatch)
    return ek_thermo.thermo


def _run_case(thermo, case, inputs, dtype):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    args_np = [np.ascontiguousarray(inputs[a].astype(dtype)) for a in case.args]
    res = getattr(thermo, case.fn)(*[torch.from_numpy(a.copy()).to(tdt) for a in args_np], **case.kwargs)
    want = getattr(oracle, case.fn)(*args_np, **case.kwargs)
    if not isinstance(res, tuple):
        res, want = (res,), (want,)
    conds = [None if case.iterative == "bisect" else conditioning(case, args_np, k) for k in range(len(res))]
    if dtype == np.float32 and case.iterative != "bisect":  # float32: (conditioning, the reference's own float32 noise)
        conds = [(c, reference_f32_noise(case, args_np, k)) for k, c in enumerate(conds)]
    return [r.numpy() for r in res], want, conds


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_functors_match_oracle_random(thermo, case, dtype):
    inputs = random_inputs(3000, seed=5)
    got, want, conds = _run_case(thermo, case, inputs, dtype)
    for g, w, c in zip(got, want, conds):
        c, nz = c if isinstance(c, tuple) else (c, None)
        compare(case, g, w, dtype, cond=c, noise=nz)


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_functors_match_oracle_edge(thermo, case, dtype):
    with np.errstate(all="ignore"):
        inputs = edge_inputs(n=600, seed=21)
        got, want, conds = _run_case(thermo, case, inputs, dtype)
    for g, w, c in zip(got, want, conds):
        c, nz = c if isinstance(c, tuple) else (c, None)
        compare(case, g, w, dtype, edge=True, cond=c, noise=nz)


@pytest.mark.parametrize("kat", KATS, ids=[f"{i}-{k[0]}" for i, k in enumerate(KATS)])
def test_kat_reference_numbers(thermo, kat):
    fn, args, kwargs, expected, rtol = kat
    got = getattr(thermo, fn)(*[torch.tensor(a, dtype=torch.float64) for a in args], **kwargs)
    if not isinstance(got, tuple):
        got, expected = (got,), (expected,)
    for g, e in zip(got, expected):
        np.testing.assert_allclose(g.numpy(), np.asarray(e, dtype=np.float64), rtol=rtol, atol=1e-8, equal_nan=True)


@pytest.mark.parametrize("ept_method", ["ifs", "bolton35", "bolton39"])
@pytest.mark.parametrize("suite", ["tqp", "ttdp"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_fused_suite_functors_match_oracle(monkeypatch, suite, ept_method, dtype):
    """The suite functors, slots 8 / 9 (ept, wbpt "direct") included: every output equals the reference function it stands
    for (the oracle's composition) -- the run-time-mask instantiation with all ten outputs and, for "ifs", the two
    compile-time-mask instantiations the library ships for the single pass (0x31F, 0x30D)."""
    from ek_thermo import fused

    hostmath_backend.install(monkeypatch)
    inp = random_inputs(3000, seed=17)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    names = ("t", "q", "p") if suite == "tqp" else ("t", "td", "p")
    a_np = [np.ascontiguousarray(inp[k].astype(dtype)) for k in names]
    a_t = [torch.from_numpy(a.copy()).to(tdt) for a in a_np]
    fn = fused.suite_tqp if suite == "tqp" else fused.suite_ttdp
    table = fused.SUITE_TQP_OUTPUTS if suite == "tqp" else fused.SUITE_TTDP_OUTPUTS
    with np.errstate(all="ignore"):
        want = (oracle.suite_tqp if suite == "tqp" else oracle.suite_ttdp)(*a_np, ept_method=ept_method)
    sets = [tuple(table)]
    if ept_method == "ifs":
        sets += [fused.ALL7_TQP if suite == "tqp" else fused.ALL7_TTDP, fused.SINGLE_PASS_TQP if suite == "tqp" else fused.SINGLE_PASS_TTDP]
    rtol = 1e-12 if dtype == np.float64 else 1e-5
    for outputs in sets:
        got = fn(*a_t, outputs=outputs, ept_method=ept_method)
        assert tuple(got) == tuple(outputs)
        for name in outputs:
            g, w = got[name].numpy(), np.asarray(want[name])
            assert g.dtype == dtype
            if dtype == np.float64:
                assert np.array_equal(np.isnan(g), np.isnan(w)), (name, outputs)
            else:  # float32 overflows where float64 does not (unphysical p - e -> 0 points; the oracle's "direct" fit runs in float64)
                assert np.mean(np.isfinite(g) != np.isfinite(w)) <= 0.01, (name, outputs)
            ok = np.isfinite(w) & np.isfinite(g)
            rel = np.abs(g[ok].astype(np.float64) - w[ok].astype(np.float64)) / np.abs(w[ok].astype(np.float64))
            # float32: a few points sit where the float32 formula itself is ill-conditioned (rh -> 0, p - e -> 0)
            assert np.quantile(rel, 0.995 if dtype == np.float32 else 1.0) <= rtol, (name, outputs, rel.max())
    with pytest.raises(KeyError):
        fn(*a_t, outputs=("ept",), ept_method="nope")


def test_batched_suite_host_logic(monkeypatch):
    """fused.suite_tqp_batch / suite_ttdp_batch on the mock device: the pointer tables, the broadcast scalar, preallocated outputs
    and the argument checks of the Python side (the CUDA kernel itself is covered by the GPU test of the same name)."""
    from ek_thermo import fused

    hostmath_backend.install(monkeypatch)
    inp = random_inputs(4 * 500, seed=3)
    ts, qs, tds, ps = ([torch.from_numpy(inp[k][j * 500:(j + 1) * 500].copy()) for j in range(4)] for k in ("t", "q", "td", "p"))
    got = fused.suite_tqp_batch(ts, qs, ps, outputs=("theta", "rh", "ept"), ept_method="bolton35")
    assert len(got) == 4
    for j in range(4):
        want = fused.suite_tqp(ts[j], qs[j], ps[j], outputs=("theta", "rh", "ept"), ept_method="bolton35")
        for name in want:
            assert torch.equal(torch.nan_to_num(got[j][name]), torch.nan_to_num(want[name])), (j, name)
    pre = [{"q": torch.empty_like(ts[0])} for _ in range(4)]
    got = fused.suite_ttdp_batch(ts, tds, 85000.0, outputs=("q", "wbpt"), out=pre)
    for j in range(4):
        assert got[j]["q"].data_ptr() == pre[j]["q"].data_ptr()
        want = fused.suite_ttdp(ts[j], tds[j], 85000.0, outputs=("q", "wbpt"))
        assert torch.equal(torch.nan_to_num(got[j]["wbpt"]), torch.nan_to_num(want["wbpt"]))
    # pressure-level data: one pressure per field
    levels = [100000.0, 85000.0, 50000, 3.0]
    got = fused.suite_tqp_batch(ts, qs, levels, outputs=("theta", "rh", "td", "ept", "wbpt"))
    for j in range(4):
        want = fused.suite_tqp(ts[j], qs[j], float(levels[j]), outputs=("theta", "rh", "td", "ept", "wbpt"))
        for name in want:
            assert torch.equal(torch.nan_to_num(got[j][name]), torch.nan_to_num(want[name])), (j, name)
    with pytest.raises(ValueError):
        fused.suite_tqp_batch(ts, qs, levels[:3])
    assert fused.suite_tqp_batch([], [], []) == []
    with pytest.raises(ValueError):
        fused.suite_tqp_batch(ts, qs[:3], ps)
    with pytest.raises(ValueError):
        fused.suite_tqp_batch(ts, qs, [p[:10] for p in ps])
    with pytest.raises(TypeError):
        fused.suite_tqp_batch(1.0, 2.0, 3.0)
    with pytest.raises(KeyError):
        fused.suite_tqp_batch(ts, qs, ps, outputs=("ept",), ept_method="nope")
