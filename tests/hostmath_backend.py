"""TEST-ONLY stand-in for the C-ABI call of ``ek_thermo._backend`` (a mock of the device, not a product path).

``install(monkeypatch)`` replaces ``_backend._call`` and ``_backend._check_device`` so that the package's
host logic (option validation, enum mapping, broadcasting, scalar operands, output allocation) runs
unchanged on CPU tensors, while the per-point work is done by tests/_hostmath/libek_hostmath.so -- the
g++ build of the very functors the kernels instantiate (ek_thermo_ops.cuh).  The mapping below mirrors
the thin argument plumbing of csrc/ek_ops_*.cu.
"""
import ctypes
import os
import subprocess
from ctypes import c_double, c_int, c_int64, c_uint32, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostmath", "libek_hostmath.so")
SRC = os.path.join(HERE, "_hostmath", "hostmath.cpp")


def build(force=False):
    hdr_dir = os.path.join(os.path.dirname(HERE), "earthkit-meteo_b200", "csrc")
    deps = [SRC] + [os.path.join(hdr_dir, f) for f in ("ek_thermo_ops.inc", "ek_thermo_formulas.inc", "ek_thermo_math.cuh")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", SRC, "-o", SO])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hostmath_run.argtypes = [ctypes.c_char_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_double), ctypes.POINTER(c_void_p),
                                      c_int64, c_int, c_int, c_double, c_uint32, c_int, c_int]
        _lib.hostmath_run.restype = c_int
    return _lib


def _val(x):
    return x.value if hasattr(x, "value") else x


def _run(op, dtype, operands, outs, n, opt0=0, opt1=0, eps=1e-4, mask=1, m=0, tm=0):
    nin = len(operands)
    ins = (c_void_p * nin)(*[o.ptr for o in operands])
    sc = (c_double * nin)(*[o.value for o in operands])
    o = (c_void_p * len(outs))(*[_val(x) for x in outs])
    rc = lib().hostmath_run(op.encode(), int(dtype == torch.float32), ins, sc, o, n, opt0, opt1, eps, mask, m, tm)
    assert rc == 0, (op, rc)


def fake_call(symbol, dtype, device, c_args):
    from ek_thermo import _backend as b

    if symbol in ("suite_tqp_batch", "suite_ttdp_batch"):  # one mock "launch" per field: the per-field operands are read back from the pointer tables
        n_seg, pa, pb, pc, scalars, level_scalars, out_tab, mask, em, n = c_args
        mask = _val(mask)
        base = symbol[:-len("_batch")]
        for j in range(_val(n_seg)):
            ops = []
            for k, tab in enumerate((pa, pb, pc)):
                tab = _val(tab)
                if tab:
                    ops.append(b.ek_operand(ctypes.cast(tab, ctypes.POINTER(c_void_p))[j], 0.0))
                else:
                    ops.append(b.ek_operand(None, level_scalars[j] if (k == 2 and level_scalars is not None) else scalars[k]))
            o = [ctypes.cast(out_tab[k], ctypes.POINTER(c_void_p))[j] if (mask >> k) & 1 else None for k in range(b.N_SUITE_SLOTS)]
            _run(base, dtype, ops, o, _val(n), mask=mask, m=_val(em))
        return None
    if symbol in ("suite_tqp", "suite_ttdp"):
        operands, (outs, mask, em, n) = c_args[:3], c_args[3:]
        mask = _val(mask)
        o = [outs[k] if (mask >> k) & 1 else None for k in range(b.N_SUITE_SLOTS)]
        op = symbol + {0x31F: "_31f", 0x30D: "_30d"}.get(mask, "") if _val(em) == 0 else symbol  # the static-mask instantiations, like the library
        return _run(op, dtype, operands, o, _val(n), mask=mask, m=_val(em))
    nin, opt_types, nout = b.SIGNATURES[symbol]
    operands = c_args[:nin]
    opts = [_val(x) for x in c_args[nin:nin + len(opt_types)]]
    outs = c_args[nin + len(opt_types):nin + len(opt_types) + nout]
    n = _val(c_args[-1])
    kw = {}
    op = symbol
    if symbol in ("specific_humidity_from_vapour_pressure", "mixing_ratio_from_vapour_pressure"):
        kw = dict(eps=opts[0])
    elif symbol in ("saturation_vapour_pressure", "saturation_vapour_pressure_slope", "saturation_mixing_ratio",
                    "saturation_specific_humidity", "lcl_temperature", "lcl"):
        kw = dict(opt0=opts[0])
    elif symbol == "wind_direction":
        kw = dict(opt0=opts[0], opt1=opts[1])
    elif symbol in ("wind_xy_to_polar", "wind_polar_to_xy"):
        kw = dict(opt0=opts[0])
    elif symbol in ("saturation_mixing_ratio_slope", "saturation_specific_humidity_slope"):
        kw = dict(opt0=opts[2], opt1=opts[0] | (opts[1] << 1), eps=opts[3])
    elif symbol in ("ept_from_dewpoint", "ept_from_specific_humidity"):
        op, kw, outs = "ept_wet_bulb", dict(opt0=int(symbol.endswith("humidity")), m=opts[0], tm=0), [outs[0], None]
    elif symbol == "saturation_ept":
        kw = dict(m=opts[0])
    elif symbol == "temperature_on_moist_adiabat":
        kw = dict(m=opts[0], tm=opts[1])
    elif symbol.startswith("wet_bulb_"):
        op = "ept_wet_bulb"
        kw = dict(opt0=int(symbol.endswith("humidity")), opt1=int("potential" in symbol), m=opts[0], tm=opts[1])
        outs = [None, outs[0]]
    elif symbol == "ept_wet_bulb":
        kw = dict(opt0=opts[0], m=opts[1], tm=opts[2], opt1=opts[3])
    return _run(op, dtype, operands, list(outs), n, **kw)


def install(monkeypatch):
    from ek_thermo import _backend as b

    monkeypatch.setattr(b, "_call", fake_call)
    monkeypatch.setattr(b, "_check_device", lambda tensors: tensors[0].device)
