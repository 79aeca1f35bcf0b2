"""``ek_thermo.thermo`` -- same role as ``earthkit.meteo.thermo`` (thermo/__init__.py:19, thermo/thermo.py:13-166):
the 39 functions at the top level and again under ``.array``."""
from . import array  # noqa: F401
from ._functions import *  # noqa: F401,F403
from ._functions import __all__  # noqa: F401
