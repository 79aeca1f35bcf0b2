"""Stand-in for ``earthkit.utils.array.namespace``."""
from . import _NUMPY_NAMESPACE  # noqa: F401
