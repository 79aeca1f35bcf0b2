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
