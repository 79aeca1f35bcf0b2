"""Stand-in for ``earthkit.utils.array.testing`` (reference tests/thermo/test_thermo.py:18)."""
from . import _NUMPY_NAMESPACE

NAMESPACE_DEVICES = [(_NUMPY_NAMESPACE, "cpu")]
