"""Shared comparison rule of the parity tests (CPU mock-device tests and GPU tests).

Tolerances (BASELINE.json north_star): float64 closed forms 1e-12 relative with NaN and inf positions
identical; the one-step Newton solve 2e-9 relative (6e-7 K at 300 K, inside the 1e-6 K bar); float32 1e-5 relative.
A point may exceed the flat limit only where the FORMULA cannot do better: by 4 x its conditioning (relative change of
the oracle under a 1-ulp input perturbation) or, in float32, by 4 x the reference's own float32 noise (its float32
result against its float64 result on the same inputs).  How often that is needed is measured, per case, in
profiles/r02_parity.json (tools/parity_report.py): float64 -- no closed-form point on IFS-shaped fields, 418 of 1e8 sampled
values overall, all explained; float32 -- no closed-form point on IFS-shaped fields.
Bisect results are quantised to 0.0293 K steps and a 1-ulp difference in exp/pow can flip an exact sign
tie: such points are counted and bounded, all others must agree to 1e-12.
"""
import numpy as np


def _step_ulps(a, k):
    """`a` moved by k ulps of its dtype (k may be negative)."""
    out = a
    target = np.asarray(np.inf if k > 0 else -np.inf, dtype=a.dtype)
    for _ in range(abs(k)):
        out = np.nextafter(out, target)
    return out


def conditioning(case, args_np, k_out=0):
    """Relative change of the oracle output per ulp of input perturbation (max over inputs).

    A result that differs from the oracle by no more than a 1-ulp input perturbation does is as exact as the
    formula allows: e.g. theta(t, p - es) loses digits without bound as p - es -> 0, for ANY two correctly
    rounded exp implementations.  compare() accepts max(rtol, 4 x this) per point.

    Each input is moved by +-1 ulp and by +-16 ulps (change divided by 16): a single 1-ulp step can be absorbed by the
    rounding of an intermediate (e.g. ept inside the "direct" wet-bulb fit, whose exponential amplifies one ulp of ept
    to 1e-12 at unphysical edge points) and would then report a conditioning of zero."""
    import thermo_oracle as oracle

    fn = getattr(oracle, case.fn)
    with np.errstate(all="ignore"):
        base = fn(*args_np, **case.kwargs)
        base = np.asarray(base[k_out] if isinstance(base, tuple) else base).astype(np.float64)
        cond = np.zeros(base.shape)
        for i, a in enumerate(args_np):
            for k in (1, -1, 16, -16):
                pert = list(args_np)
                pert[i] = _step_ulps(a, k)
                r = fn(*pert, **case.kwargs)
                r = np.asarray(r[k_out] if isinstance(r, tuple) else r).astype(np.float64)
                d = np.abs(r - base) / np.maximum(np.abs(base), 1e-300) / abs(k)
                cond = np.fmax(cond, np.where(np.isfinite(d), d, 0.0))
    return cond


def reference_f32_noise(case, args_np, k_out=0):
    """Relative difference between the oracle evaluated in float32 and in float64 on the same (float32-valued) inputs: the
    rounding noise of the reference's own float32 arithmetic.  No implementation can be expected to match the float32
    oracle more closely than it matches itself."""
    import thermo_oracle as oracle

    fn = getattr(oracle, case.fn)
    with np.errstate(all="ignore"):
        w32 = fn(*args_np, **case.kwargs)
        w64 = fn(*[np.asarray(a).astype(np.float64) for a in args_np], **case.kwargs)
        w32 = np.asarray(w32[k_out] if isinstance(w32, tuple) else w32).astype(np.float64)
        w64 = np.asarray(w64[k_out] if isinstance(w64, tuple) else w64).astype(np.float64)
        noise = np.abs(w32 - w64) / np.maximum(np.abs(w64), 1e-300)
    return np.where(np.isfinite(noise), noise, 0.0)


DT_LAST = 120.0 / 2 ** 12  # last bisection step, 0.0293 K (T:1069-1075)


def compare(case, got, want, dtype, edge=False, cond=None, grid=False, noise=None):
    """Assert `got` (new build) equals `want` (oracle / reference) under the parity rule of this module.

    edge: the special-value input set (looser conditioning factor, float32 overflow boundaries tolerated);
    grid: the reference's own grid-aligned test data, which hits exact bisection sign ties (SURVEY.md 7.3-H3:
          the reference re-run against its own t_wet.csv already differs by 2.1e-4 relative there).
    """
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, case.id
    f32 = dtype == np.float32
    got = got.astype(np.float64)
    want = want.astype(np.float64)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    bisect = case.iterative == "bisect"
    if f32:
        # float32 near overflow/underflow: a 1-ulp difference decides inf vs finite vs NaN, and the reference's
        # float32 "direct" path evaluates its polynomials in float64 (SURVEY.md 8(c) caveat)
        nonfin_g, nonfin_w = ~np.isfinite(got), ~np.isfinite(want)
        assert np.mean(nonfin_g != nonfin_w) < (0.02 if edge else 0.005), f"non-finite positions differ: {case.id}"
        fin = ~(nonfin_g | nonfin_w)
    else:
        if bisect:
            assert np.mean(nan_g != nan_w) < 0.02, f"NaN positions differ: {case.id}"
        else:
            np.testing.assert_array_equal(nan_g, nan_w, err_msg=f"NaN positions differ: {case.id}")
        ok = ~(nan_g | nan_w)
        inf = ok & (np.isinf(got) | np.isinf(want))
        fin = ok & ~inf
        np.testing.assert_array_equal(got[inf], want[inf], err_msg=f"inf differ: {case.id}")
    rtol = 1e-5 if f32 else 1e-12
    if case.iterative == "newton":
        # 2e-9 * 300 K = 6e-7 K: inside the 1e-6 K bar for the iterative solves.  float32: measured p99.9 of the one-step
        # solve is 1.1e-5 ... 2.4e-5 (profiles/r02_parity.json) -- the step amplifies float32 rounding of its first guess
        rtol = 2.5e-5 if f32 else 2e-9
    with np.errstate(all="ignore"):
        diff = np.abs(got[fin] - want[fin])
        rel = diff / np.maximum(np.abs(want[fin]), 1e-300)
    # edge set: results that are ~0 by cancellation (e.g. the inverse es formula at t -> 0) carry absolute noise
    small = diff <= ((1e-6 if f32 else 1e-10) if edge else (1e-37 if f32 else 1e-30))
    tol = rtol if cond is None else np.maximum(rtol, (32.0 if edge else 4.0) * np.asarray(cond)[fin])
    if noise is not None:
        tol = np.maximum(tol, 4.0 * np.asarray(noise)[fin])
    bad = (rel > tol) & ~small
    if bisect:
        # a flipped sign at a near-tie is pulled back by the remaining halvings: never further than 2 last steps
        lim = 2.1 * DT_LAST * (4.0 if f32 else 1.0)
        far = bad & (diff > lim) & (np.abs(want[fin]) < 1e4)
        assert far.mean() < (0.02 if (f32 or edge) else 0.0) + 1e-12, f"{case.id}: {far.sum()} of {far.size} bisect points off by more than {lim:.3f} K"
        frac = 0.05 if (grid or f32 or edge) else 0.005
        assert bad.mean() < frac, f"{case.id}: {bad.sum()} of {bad.size} bisect points differ"
    elif f32:
        # float32: flat 1e-5 (or the formula's own limits, see the module docstring) on every point of the closed forms;
        # the one-step Newton solve is chaotic on a few points per thousand (its float32 first guess decides the regime and
        # the step is not a contraction there), the special-value set sits on float32 overflow boundaries
        frac = 0.03 if edge else (0.002 if case.iterative == "newton" else 0.0)
        assert bad.mean() <= frac, f"{case.id}: {bad.sum()} of {bad.size} beyond {rtol} (max {rel[bad].max():.2e})"
    else:
        assert not bad.any(), f"{case.id}: max rel {rel[bad].max():.3e} at {np.flatnonzero(fin)[np.argmax(np.where(bad, rel, 0))]}"
