"""``ek_thermo.thermo.array`` -- same role as ``earthkit.meteo.thermo.array`` (thermo/array/__init__.py:14)."""
from ._functions import *  # noqa: F401,F403
from ._functions import __all__  # noqa: F401
