"""Shared comparison rule of the parity tests (CPU mock-device tests and GPU tests).

Tolerances (BASELINE.json north_star): float64 closed forms 1e-12 relative with NaN and inf positions
identical; the one-step Newton solve 1e-10 relative (= far below 1e-6 K); float32 2e-5 relative.
Bisect results are quantised to 0.0293 K steps and a 1-ulp difference in exp/pow can flip an exact sign
tie: such points are counted and bounded, all others must agree to 1e-12.
"""
import numpy as np


def compare(case, got, want, dtype, edge=False):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, case.id
    f32 = dtype == np.float32
    if f32:
        want = want.astype(np.float64)
        got = got.astype(np.float64)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    if case.iterative == "bisect" or f32:
        # float32 values near overflow/underflow boundaries: a 1-ulp difference decides inf vs finite (and the
        # reference's float32 "direct" path evaluates its polynomials in float64, SURVEY.md §8(c) caveat)
        assert np.mean(nan_g != nan_w) < 0.02, case.id
    else:
        np.testing.assert_array_equal(nan_g, nan_w, err_msg=f"NaN positions differ: {case.id}")
    ok = ~(nan_g | nan_w)
    inf = ok & (np.isinf(got) | np.isinf(want))
    fin = ok & ~inf
    if not (f32 and edge):
        np.testing.assert_array_equal(got[inf], want[inf], err_msg=f"inf differ: {case.id}")
    rtol = 2e-5 if f32 else 1e-12
    if case.iterative == "newton":
        rtol = 5e-5 if f32 else 1e-10
    with np.errstate(all="ignore"):
        rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300)
        small = np.abs(got[fin] - want[fin]) <= (1e-30 if not f32 else 1e-37)
    bad = (rel > rtol) & ~small
    if case.iterative == "bisect":
        assert bad.mean() < (0.03 if f32 else 0.005), f"{case.id}: {bad.sum()} of {bad.size} bisect points differ"
    elif f32 and edge:
        assert bad.mean() < 0.03, f"{case.id}: {bad.sum()} of {bad.size}"
    elif f32:
        # float32: 1e-5-class agreement, except where the formula itself is ill-conditioned in float32
        # (p - es -> 0 amplifies a 1-ulp difference of expf without bound): allow 1 % such points
        assert bad.mean() < 0.01, f"{case.id}: {bad.sum()} of {bad.size} beyond {rtol}"
    else:
        assert not bad.any(), f"{case.id}: max rel {rel.max():.3e} at {np.argmax(rel)}"


