"""``ek_thermo.vertical`` -- hybrid (IFS model) level pressure on torch CUDA tensors.

Drop-in for ``earthkit.meteo.vertical.pressure_on_hybrid_levels`` (reference
src/earthkit/meteo/vertical/array/vertical.py:505-737, "V"): same signature, defaults, outputs and exceptions.
One kernel launch reads ``sp`` once and writes every requested [level, point] array from registers.
"""
from __future__ import annotations

import math
from ctypes import c_double, c_int, c_int64, c_void_p

import torch

from . import _backend as _b

_OUTPUTS = ("full", "half", "alpha", "delta")


def _coeff_tensors(A, B, dtype, device):
    a = torch.as_tensor(A, dtype=dtype, device=device).contiguous()
    b = torch.as_tensor(B, dtype=dtype, device=device).contiguous()
    if a.dim() != 1 or a.shape != b.shape or a.numel() < 2:
        raise ValueError("A and B must be 1-D arrays of the same size (one value per half-level, at least 2)")
    return a, b


def pressure_on_hybrid_levels(A, B, sp, levels=None, alpha_top="ifs", output="full", vertical_axis=0):
    """Pressure on full / half levels, ``delta`` and ``alpha`` of hybrid levels.  Reference V:505-737.

    ``sp`` is a torch CUDA tensor (float64 / float32) of any shape; ``A`` / ``B`` are array-likes with one value per
    half-level.  Returns a tensor or a tuple of tensors shaped ``(levels,) + sp.shape`` (vertical axis moved to
    ``vertical_axis``), in the dtype of ``sp``.
    """
    if isinstance(output, str):
        output = (output,)
    if not output:
        raise ValueError("At least one output type must be specified.")  # V:617-618
    for out in output:
        if out not in _OUTPUTS:
            raise ValueError(f"Unknown output type '{out}'. Allowed values are 'full', 'half', 'alpha' or 'delta'.")  # V:620-624
    if alpha_top not in ("ifs", "arpege"):
        raise ValueError(f"Unknown method '{alpha_top}' for pressure calculation. Use 'ifs' or 'arpege'.")  # V:626-627
    if not isinstance(sp, torch.Tensor):
        raise TypeError("ek_thermo.vertical: sp must be a torch CUDA tensor (no CPU path; use earthkit.meteo.vertical for host arrays)")
    dev = _b._check_device([sp])
    dtype = sp.dtype if sp.dtype in (torch.float64, torch.float32) else torch.float64
    spc = sp.to(dtype).contiguous()
    a, b = _coeff_tensors(A, B, dtype, dev)
    nhalf = a.numel()
    nlev = nhalf - 1
    if levels is not None:  # V:634-654
        lv = [int(x) for x in (levels.tolist() if hasattr(levels, "tolist") else list(levels))]
        lmax, lmin = max(lv), min(lv)
        if lmax > nlev:
            raise ValueError(f"Requested level {lmax} exceeds the maximum number of levels {nlev}.")
        if lmin < 1:
            raise ValueError(f"Level numbering starts at 1. Found level={lmin} < 1.")
        full_rows, half_rows, top_k = [x - 1 for x in lv], lv, lmin - 1
    else:
        full_rows, half_rows, top_k = list(range(nlev)), list(range(nhalf)), 0
    want = {o: True for o in output}
    npl = spc.numel()
    top_toa = 0
    if "delta" in want or "alpha" in want:  # V:678: one decision for the whole field
        a_top, b_top = float(a[top_k]), float(b[top_k])
        if b_top == 0.0 or npl == 0:
            top_toa = int(a_top <= 0.1)
        else:
            flag = torch.zeros(1, dtype=torch.int32, device=dev)
            _b.call_raw("hybrid_top_is_toa", dtype, dev, c_void_p(spc.data_ptr()), c_int64(npl), c_double(a_top), c_double(b_top),
                        c_void_p(flag.data_ptr()))
            top_toa = int(flag.item())
    rows_f = torch.tensor(full_rows, dtype=torch.int32, device=dev)
    rows_h = torch.tensor(half_rows, dtype=torch.int32, device=dev)
    res = {}
    for name in _OUTPUTS:
        if name in want:
            nrow = len(half_rows) if name == "half" else len(full_rows)
            res[name] = torch.empty((nrow,) + tuple(sp.shape), dtype=dtype, device=dev)
    ptr = lambda n: c_void_p(res[n].data_ptr()) if n in res else c_void_p(None)  # noqa: E731
    if npl > 0:
        _b.call_raw("pressure_on_hybrid_levels", dtype, dev, c_void_p(a.data_ptr()), c_void_p(b.data_ptr()), c_int(nhalf),
                    c_void_p(spc.data_ptr()), c_int64(npl), c_void_p(rows_f.data_ptr()), c_int(len(full_rows)), c_void_p(rows_h.data_ptr()),
                    c_int(len(half_rows)), c_int(top_k), c_int(top_toa), c_double(math.log(2.0) if alpha_top == "ifs" else 1.0),
                    ptr("full"), ptr("half"), ptr("delta"), ptr("alpha"))
    outs = [res[o] for o in output]
    if vertical_axis != 0 and outs[0].dim() > 1:  # V:731-733
        outs = [r.movedim(0, vertical_axis) for r in outs]
    return outs[0] if len(outs) == 1 else tuple(outs)
