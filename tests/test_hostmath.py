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
from compare import compare, conditioning
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
    return [r.numpy() for r in res], want, conds


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_functors_match_oracle_random(thermo, case, dtype):
    inputs = random_inputs(3000, seed=5)
    got, want, conds = _run_case(thermo, case, inputs, dtype)
    for g, w, c in zip(got, want, conds):
        compare(case, g, w, dtype, cond=c)


@pytest.mark.parametrize("case", CASES, ids=[c.id for c in CASES])
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_functors_match_oracle_edge(thermo, case, dtype):
    with np.errstate(all="ignore"):
        inputs = edge_inputs(n=600, seed=21)
        got, want, conds = _run_case(thermo, case, inputs, dtype)
    for g, w, c in zip(got, want, conds):
        compare(case, g, w, dtype, edge=True, cond=c)


@pytest.mark.parametrize("kat", KATS, ids=[f"{i}-{k[0]}" for i, k in enumerate(KATS)])
def test_kat_reference_numbers(thermo, kat):
    fn, args, kwargs, expected, rtol = kat
    got = getattr(thermo, fn)(*[torch.tensor(a, dtype=torch.float64) for a in args], **kwargs)
    if not isinstance(got, tuple):
        got, expected = (got,), (expected,)
    for g, e in zip(got, expected):
        np.testing.assert_allclose(g.numpy(), np.asarray(e, dtype=np.float64), rtol=rtol, atol=1e-8, equal_nan=True)
