"""CPU oracle for ``earthkit.meteo.vertical.pressure_on_hybrid_levels``.  TEST INFRASTRUCTURE, NOT PRODUCT.

SURVEY.md 8(f)-1: the step immediately before the thermo path on model levels -- it produces the
``[level, point]`` pressure field the thermo kernels read.  numpy restatement of the reference's
``src/earthkit/meteo/vertical/array/vertical.py:505-737`` ("V") with its operation order.  Pinned by
tests/golden/make_golden.py against the live reference (bit-identical, see PINNING.json "hybrid") and against the
reference's golden vectors (tests/vertical/_hybrid_core_data.py, consumed at tests/vertical/test_array_vertical.py:159-370).
"""
from __future__ import annotations

import numpy as np

OUTPUTS = ("full", "half", "alpha", "delta")
PRESSURE_TOA = 0.1  # V:667


def pressure_on_hybrid_levels(A, B, sp, levels=None, alpha_top="ifs", output="full", vertical_axis=0):
    """V:505-737."""
    if isinstance(output, str):
        output = (output,)
    if not output:
        raise ValueError("At least one output type must be specified.")  # V:617-618
    for out in output:
        if out not in OUTPUTS:
            raise ValueError(f"Unknown output type '{out}'. Allowed values are 'full', 'half', 'alpha' or 'delta'.")  # V:620-624
    if alpha_top not in ("ifs", "arpege"):
        raise ValueError(f"Unknown method '{alpha_top}' for pressure calculation. Use 'ifs' or 'arpege'.")  # V:626-627
    A = np.asarray(A)
    B = np.asarray(B)
    sp = np.asarray(sp)
    sel_half = sel_full = None
    if levels is not None:  # V:634-654: a contiguous band of half-levels is computed, then the requested rows picked
        nlev = A.shape[0] - 1
        levels = np.asarray(levels)
        lmax, lmin = int(levels.max()), int(levels.min())
        if lmax > nlev:
            raise ValueError(f"Requested level {lmax} exceeds the maximum number of levels {nlev}.")
        if lmin < 1:
            raise ValueError(f"Level numbering starts at 1. Found level={lmin} < 1.")
        half_idx = np.arange(lmin - 1, lmax + 1)
        A, B = A[half_idx], B[half_idx]
        sel_half = np.nonzero(levels[:, None] == half_idx[None, :])[1]
        sel_full = sel_half - 1
    shape_half = (A.shape[0],) + (1,) * sp.ndim
    with np.errstate(all="ignore"):
        ph = A.reshape(shape_half) + B.reshape(shape_half) * sp[np.newaxis, ...]  # V:663
        res = {}
        if "delta" in output or "alpha" in output:
            a_top = np.log(2) if alpha_top == "ifs" else 1.0  # V:669
            delta = np.zeros((A.shape[0] - 1,) + sp.shape)
            delta[1:, ...] = np.log(ph[2:, ...] / ph[1:-1, ...])  # V:675
            top_is_toa = bool(np.any(ph[0, ...] <= PRESSURE_TOA))  # V:678: one decision for the whole field
            if top_is_toa:
                delta[0, ...] = np.log(ph[1, ...] / PRESSURE_TOA)  # V:679
            else:
                delta[0, ...] = np.log(ph[1, ...] / ph[0, ...])  # V:682
            alpha = np.zeros((A.shape[0] - 1,) + sp.shape)
            alpha[1:, ...] = 1.0 - ph[1:-1, ...] / (ph[2:, ...] - ph[1:-1, ...]) * delta[1:, ...]  # V:687-689
            if top_is_toa:
                alpha[0, ...] = a_top  # V:693
            else:
                alpha[0, ...] = 1.0 - ph[0, ...] / (ph[1, ...] - ph[0, ...]) * delta[0, ...]  # V:696-698
            res["delta"], res["alpha"] = delta, alpha
        if "full" in output:
            res["full"] = ph[:-1, ...] + 0.5 * np.diff(ph, axis=0)  # V:708
        res["half"] = ph
    outs = []
    for out in output:  # V:713-729
        r = res[out]
        if levels is not None:
            r = r[sel_half if out == "half" else sel_full, ...]
        outs.append(r)
    if vertical_axis != 0 and outs[0].ndim > 1:  # V:731-733
        outs = [np.moveaxis(r, 0, vertical_axis) for r in outs]
    return outs[0] if len(outs) == 1 else tuple(outs)


# --------------------------------------------------------------------------------------------------
# geopotential thickness / geopotential / height on hybrid levels -- the in-tree consumer of
# thermo.specific_gas_constant (V:798-801): R(q) * t, then a bottom-up cumulative sum along the level axis.
# --------------------------------------------------------------------------------------------------
G = 9.80665  # constants.g (constants.py:53)
R_EARTH = 6371229  # constants.R_earth (constants.py:57)
RD, RV = 287.0597, 461.51  # constants.py:22,26


def geopotential_height_from_geopotential(z):
    """V:330-354"""
    return z / G


def geometric_height_from_geopotential(z, R_earth=R_EARTH):
    """V:472-502"""
    z = z / G
    return R_earth * z / (R_earth - z)


def _thickness(t, q, alpha, delta):
    """V:741-812 (_compute_relative_geopotential_thickness_on_hybrid_levels), vertical axis first."""
    R = RD + (RV - RD) * q  # thermo.specific_gas_constant (T:1706)
    d = R * t
    dphi_half = np.cumulative_sum(np.flip(d[1:, ...] * delta[1:, ...], axis=0), axis=0)  # V:804
    dphi_half = np.flip(dphi_half, axis=0)
    dphi = np.zeros_like(d)
    dphi[:-1, ...] = dphi_half + d[:-1, ...] * alpha[:-1, ...]  # V:808
    dphi[-1, ...] = d[-1, ...] * alpha[-1, ...]  # V:809
    return dphi


def relative_geopotential_thickness_on_hybrid_levels_from_alpha_delta(t, q, alpha, delta, vertical_axis=0):
    """V:815-891"""
    t, q, alpha, delta = (np.asarray(x) for x in (t, q, alpha, delta))
    if vertical_axis != 0:
        t, q, alpha, delta = (np.moveaxis(x, vertical_axis, 0) for x in (t, q, alpha, delta))
    dphi = _thickness(t, q, alpha, delta)
    return np.moveaxis(dphi, 0, vertical_axis) if vertical_axis != 0 else dphi


def _hybrid_subset(data, A, vertical_axis=0):
    """V:1191-1203: data on fewer levels than A/B describe = the bottom-most contiguous band."""
    nlev_t, nlev = data.shape[vertical_axis], A.shape[0] - 1
    return None if nlev_t == nlev else list(range(nlev - nlev_t + 1, nlev + 1))


def relative_geopotential_thickness_on_hybrid_levels(t, q, A, B, sp, alpha_top="ifs", vertical_axis=0):
    """V:894-994"""
    t, q, A, B, sp = (np.asarray(x) for x in (t, q, A, B, sp))
    levels = _hybrid_subset(t, A, vertical_axis)
    alpha, delta = pressure_on_hybrid_levels(A, B, sp, alpha_top=alpha_top, levels=levels, output=("alpha", "delta"))
    if vertical_axis != 0:  # V:981-986 (alpha/delta are moved too, exactly as the reference does)
        alpha, delta, t, q = (np.moveaxis(x, vertical_axis, 0) for x in (alpha, delta, t, q))
    dphi = _thickness(t, q, alpha, delta)
    return np.moveaxis(dphi, 0, vertical_axis) if vertical_axis != 0 else dphi


def geopotential_on_hybrid_levels(t, q, zs, A, B, sp, alpha_top="ifs", vertical_axis=0):
    """V:997-1069"""
    z = relative_geopotential_thickness_on_hybrid_levels(t, q, A, B, sp, vertical_axis=vertical_axis, alpha_top=alpha_top)
    return z + np.asarray(zs)


def height_on_hybrid_levels(t, q, zs, A, B, sp, alpha_top="ifs", h_type="geometric", h_reference="ground", vertical_axis=0):
    """V:1072-1188"""
    if h_reference not in ["sea", "ground"]:
        raise ValueError(f"Unknown '{h_reference=}'. Use 'sea' or 'ground'.")
    z_thickness = relative_geopotential_thickness_on_hybrid_levels(t, q, A, B, sp, alpha_top=alpha_top, vertical_axis=vertical_axis)
    if h_reference == "sea":
        z = z_thickness + np.asarray(zs)
        return geometric_height_from_geopotential(z) if h_type == "geometric" else geopotential_height_from_geopotential(z)
    if h_type == "geometric":
        zs = np.asarray(zs)
        return geometric_height_from_geopotential(z_thickness + zs) - geometric_height_from_geopotential(zs)
    return geopotential_height_from_geopotential(z_thickness)
