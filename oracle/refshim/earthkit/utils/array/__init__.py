"""numpy-only stand-in for the third-party ``earthkit.utils.array`` (earthkit-utils>=0.2).

TEST INFRASTRUCTURE ONLY.  The reference's thermo hot path needs three things from
earthkit-utils (pyproject.toml:33 of the reference; not vendored, not installable offline):
``array_namespace(*arrays)`` returning a numpy-like namespace with the extra helpers
``polyval(x, coeffs)`` (ascending coefficients), ``size(x)`` and ``device(x)``
(call sites: thermo/array/thermo.py:13,192,229,412,464,826,1048,1056,1082;
thermo/array/es_comp.py:12,73,100,128).  This module provides exactly that for numpy so that
``/root/reference/src`` imports unmodified in this container when golden vectors are generated
and the oracle is pinned against it (both by tests/golden/make_golden.py), and when ``oracle/build_ref.py``
installs the reference into ``oracle/_ref`` for the CPU arm of bench.py.
It is never imported by the product package.
"""
import numpy as _np
from numpy.polynomial import polynomial as _poly


class _NumpyNamespace:
    """Forwards everything to numpy, plus the three non-standard helpers."""

    nan = _np.nan

    def __getattr__(self, name):
        return getattr(_np, name)

    @staticmethod
    def polyval(x, c):
        return _poly.polyval(x, c)

    @staticmethod
    def size(x):
        return _np.size(x)

    @staticmethod
    def device(x):
        return getattr(x, "device", "cpu")


_NUMPY_NAMESPACE = _NumpyNamespace()


def array_namespace(*args):
    return _NUMPY_NAMESPACE
