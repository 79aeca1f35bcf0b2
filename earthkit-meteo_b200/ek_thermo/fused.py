"""Fused multi-output kernels: read the state of each grid point once, write every requested field.

These have no counterpart function in the reference; every output equals the reference function
named in ``SUITE_TQP_OUTPUTS`` / ``SUITE_TTDP_OUTPUTS`` applied to the same inputs (the oracle's
``suite_tqp`` / ``suite_ttdp`` state that composition).  HBM traffic per point is
``8 * (3 + len(outputs))`` bytes in float64, instead of one full read+write pass per function.
"""
from __future__ import annotations

from . import _backend as _b

# output name -> slot (include/ek_thermo.h EK_S_*), with the reference function each one equals
SUITE_TQP_OUTPUTS = {
    "theta": 0,  # potential_temperature(t, p)                          T:801-829
    "es": 1,  # saturation_vapour_pressure(t)                            T:235-279
    "rh": 2,  # relative_humidity_from_specific_humidity(t, q, p)        T:524-556
    "td": 3,  # dewpoint_from_specific_humidity(q, p)                    T:702-735
    "tv": 4,  # virtual_temperature(t, q)                                T:738-764
    "w": 5,  # mixing_ratio_from_specific_humidity(q)                    T:80-102
    "e": 6,  # vapour_pressure_from_specific_humidity(q, p)              T:105-131
    "thetav": 7,  # virtual_potential_temperature(t, q, p)               T:767-798
    "ept": 8,  # ept_from_specific_humidity(t, q, p, method=ept_method)  T:1390-1415
    "wbpt": 9,  # wet_bulb_potential_temperature_from_specific_humidity(t, q, p, ept_method, t_method="direct")  T:1637-1675
}
SUITE_TTDP_OUTPUTS = {
    "theta": 0,  # potential_temperature(t, p)
    "es": 1,  # saturation_vapour_pressure(t)
    "rh": 2,  # relative_humidity_from_dewpoint(t, td)                   T:494-521
    "q": 3,  # specific_humidity_from_dewpoint(td, p)                    T:559-591
    "tv": 4,  # virtual_temperature(t, q(td, p))
    "w": 5,  # mixing_ratio_from_dewpoint(td, p)                         T:594-626
    "e": 6,  # saturation_vapour_pressure(td, phase="water")
    "thetav": 7,  # virtual_potential_temperature(t, q(td, p), p)
    "ept": 8,  # ept_from_dewpoint(t, td, p, method=ept_method)          T:1326-1387
    "wbpt": 9,  # wet_bulb_potential_temperature_from_dewpoint(t, td, p, ept_method, t_method="direct")  T:1593-1634
}
# the single pass BASELINE.json names: "read t/q/p (or t/td/p) once and write theta, rh, td, theta_e and wbpt"
SINGLE_PASS_TQP = ("theta", "rh", "td", "ept", "wbpt")
SINGLE_PASS_TTDP = ("theta", "rh", "q", "ept", "wbpt")
# ... and with the two remaining fields of the default suite: 10 arrays = 80 bytes per point in float64
ALL7_TQP = ("theta", "es", "rh", "td", "tv", "ept", "wbpt")
ALL7_TTDP = ("theta", "es", "rh", "q", "tv", "ept", "wbpt")
DEFAULT_TQP = ("theta", "es", "rh", "td", "tv")
DEFAULT_TTDP = ("theta", "es", "rh", "q", "tv")

_EPT = {"ifs": 0, "bolton35": 1, "bolton39": 2}
_TM = {None: 0, "none": 0, "direct": 1, "bisect": 2, "newton": 3}


def _slots(table, outputs):
    try:
        return [table[o] for o in outputs]
    except KeyError as e:
        raise ValueError(f"unknown suite output {e.args[0]!r}; choose from {sorted(table)}") from None


def _ept_id(ept_method):
    return _EPT[ept_method]  # KeyError for an unknown method, as the reference (T:1026)


def suite_tqp(t, q, p, outputs=DEFAULT_TQP, out=None, ept_method="ifs"):
    """One pass over (t, q, p) producing the requested fields; returns ``{name: tensor}``.

    ``ept_method`` ("ifs", "bolton35", "bolton39") is the formulation of the ``"ept"`` / ``"wbpt"`` outputs."""
    outputs = tuple(outputs)
    return _b.execute_suite("suite_tqp", (t, q, p), outputs, _slots(SUITE_TQP_OUTPUTS, outputs), out, _ept_id(ept_method))


def suite_ttdp(t, td, p, outputs=DEFAULT_TTDP, out=None, ept_method="ifs"):
    """One pass over (t, td, p) producing the requested fields; returns ``{name: tensor}``."""
    outputs = tuple(outputs)
    return _b.execute_suite("suite_ttdp", (t, td, p), outputs, _slots(SUITE_TTDP_OUTPUTS, outputs), out, _ept_id(ept_method))


def suite_tqp_batch(ts, qs, ps, outputs=DEFAULT_TQP, out=None, ept_method="ifs"):
    """``suite_tqp`` over a LIST of separate fields (one tensor per level / member, each its own allocation) in one launch.

    ``ts`` / ``qs`` / ``ps`` are lists of same-shape contiguous CUDA tensors, or a Python number for a broadcast operand; ``ps``
    may also be a list of Python numbers, ONE PRESSURE PER FIELD -- pressure-level data, the reference user's loop
    ``for lev: potential_temperature(t[lev], p_lev)`` in one launch.  Returns ``[{name: tensor}, ...]``, one dict per field; every
    value is bit-identical to ``suite_tqp(ts[j], qs[j], ps[j])``.  A launch per 1 M-point level is bound by launch and pipeline-fill latency
    (12-16 us for 6 us of HBM time); one launch over all levels is not."""
    outputs = tuple(outputs)
    return _b.execute_suite_batch("suite_tqp_batch", (ts, qs, ps), outputs, _slots(SUITE_TQP_OUTPUTS, outputs), out, _ept_id(ept_method))


def suite_ttdp_batch(ts, tds, ps, outputs=DEFAULT_TTDP, out=None, ept_method="ifs"):
    """``suite_ttdp`` over a list of separate fields in one launch (see ``suite_tqp_batch``)."""
    outputs = tuple(outputs)
    return _b.execute_suite_batch("suite_ttdp_batch", (ts, tds, ps), outputs, _slots(SUITE_TTDP_OUTPUTS, outputs), out, _ept_id(ept_method))


def suite_tq_hybrid(t, q, sp, A, B, outputs=DEFAULT_TQP, out=None, want_p=False, ept_method="ifs"):
    """The (t, q, p) suite on hybrid model levels with the pressure computed inside the kernel.

    ``t`` and ``q`` are ``[nlev, ...]`` fields, ``sp`` the surface pressure ``[...]``, ``A`` / ``B`` the ``nlev + 1``
    half-level coefficients.  Equals ``suite_tqp(t, q, p)`` with
    ``p = earthkit.meteo.vertical.pressure_on_hybrid_levels(A, B, sp, output="full")`` (reference
    vertical/array/vertical.py:663,708) -- but the ``[nlev, ...]`` pressure array is never written or read:
    56 instead of 64 bytes per point for the five default outputs.  ``want_p=True`` also returns it (key ``"p"``).
    """
    import ctypes
    from ctypes import c_int, c_int64, c_uint32, c_void_p

    import torch

    outputs = tuple(outputs)
    slots = _slots(SUITE_TQP_OUTPUTS, outputs)
    for x in (t, q, sp):
        if not isinstance(x, torch.Tensor):
            raise TypeError("suite_tq_hybrid: t, q and sp must be torch CUDA tensors")
    dev = _b._check_device([t, q, sp])
    dtype = t.dtype
    if dtype not in (torch.float64, torch.float32) or q.dtype != dtype or sp.dtype != dtype:
        raise TypeError("suite_tq_hybrid: t, q and sp must share one dtype (float64 or float32)")
    if t.shape != q.shape or t.dim() < 1 or tuple(t.shape[1:]) != tuple(sp.shape):
        raise ValueError(f"suite_tq_hybrid: expected t, q of shape (nlev,) + sp.shape, got {tuple(t.shape)}, {tuple(q.shape)}, {tuple(sp.shape)}")
    nlev, npl = int(t.shape[0]), int(sp.numel())
    a = torch.as_tensor(A, dtype=dtype, device=dev).contiguous()
    b = torch.as_tensor(B, dtype=dtype, device=dev).contiguous()
    if a.numel() != nlev + 1 or b.numel() != nlev + 1:
        raise ValueError(f"suite_tq_hybrid: A and B need nlev + 1 = {nlev + 1} half-level values")
    tc, qc, spc = t.contiguous(), q.contiguous(), sp.contiguous()
    ptrs = (c_void_p * _b.N_SUITE_SLOTS)()
    mask = 0
    res = {}
    em = _ept_id(ept_method)
    for name, k in zip(outputs, slots):
        if out is not None and name in out:
            o = out[name]
            if o.dtype != dtype or o.shape != t.shape or not o.is_contiguous() or o.device != dev:
                raise ValueError(f"suite_tq_hybrid: preallocated output {name!r} must be a contiguous {dtype} tensor of shape {tuple(t.shape)}")
        else:
            o = torch.empty_like(tc)
        res[name] = o
        ptrs[k] = o.data_ptr()
        mask |= 1 << k
    p_ptr = c_void_p(None)
    if want_p:
        if out is not None and "p" in out:
            o = out["p"]
            if o.dtype != dtype or o.shape != t.shape or not o.is_contiguous() or o.device != dev:
                raise ValueError(f"suite_tq_hybrid: preallocated output 'p' must be a contiguous {dtype} tensor of shape {tuple(t.shape)}")
            res["p"] = o
        else:
            res["p"] = torch.empty_like(tc)
        p_ptr = c_void_p(res["p"].data_ptr())
    if npl > 0:
        _b.call_raw("suite_tq_hybrid", dtype, dev, c_void_p(tc.data_ptr()), c_void_p(qc.data_ptr()), c_void_p(spc.data_ptr()),
                    c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), c_int(nlev), c_int64(npl), ctypes.cast(ptrs, ctypes.POINTER(c_void_p)),
                    c_uint32(mask), c_int(em), p_ptr)
    return res


def ept_wet_bulb(t, h, p, humidity="q", ept_method="ifs", t_method="direct", potential=True, want_ept=True, want_wb=True):
    """Equivalent potential temperature and wet-bulb (potential) temperature in one pass.

    ``h`` is the specific humidity (``humidity="q"``) or the dewpoint (``"td"``).  Equals
    ``ept_from_*`` and ``wet_bulb[_potential]_temperature_from_*`` of the reference with the same
    ``ept_method`` / ``t_method``.  Returns ``(ept, wb)``, with None for an output not wanted.
    """
    if humidity not in ("q", "td"):
        raise ValueError(f"humidity={humidity!r} must be 'q' or 'td'")
    m = _EPT[ept_method]
    if t_method not in _TM or (t_method == "direct" and not potential):
        raise ValueError(f"temperature_on_moist_adiabat: invalid t_method={t_method} specified!")
    tm = _TM[t_method]
    if tm == 0:
        want_wb = False
    if not (want_ept or want_wb):
        raise ValueError("nothing to compute")
    res = _b.execute("ept_wet_bulb", (t, h, p), (int(humidity == "q"), m, tm, int(bool(potential))), want=(want_ept, want_wb))
    if want_ept and want_wb:
        return res
    return (res, None) if want_ept else (None, res)
