"""``ek_thermo.wind`` -- the elementwise functions of ``earthkit.meteo.wind`` on torch CUDA tensors.

Same names, signatures, defaults and exceptions as the reference (src/earthkit/meteo/wind/array/wind.py, "W"); one
kernel launch per call.  ``windrose`` (a 2-D histogram, W:254-328) is not an elementwise function and is not provided.
"""
from __future__ import annotations

from . import _backend as _b

_CONV = {"meteo": 0, "polar": 1}

__all__ = ["speed", "direction", "xy_to_polar", "polar_to_xy", "w_from_omega", "coriolis"]


def speed(u, v):
    """Wind speed hypot(u, v).  Reference W:15-34."""
    return _b.execute("wind_speed", (u, v))


def direction(u, v, convention="meteo", to_positive=True):
    """Wind direction in degrees ("meteo": the direction the wind blows from; "polar").  Reference W:64-104."""
    if convention not in _CONV:
        raise ValueError(f"direction(): invalid convention={convention}!")
    return _b.execute("wind_direction", (u, v), (_CONV[convention], int(bool(to_positive))))


def xy_to_polar(x, y, convention="meteo"):
    """(speed, direction) in one kernel.  Reference W:107-135."""
    if convention not in _CONV:
        raise ValueError(f"direction(): invalid convention={convention}!")
    return _b.execute("wind_xy_to_polar", (x, y), (_CONV[convention],))


def polar_to_xy(magnitude, direction, convention="meteo"):
    """(x, y) components from magnitude and direction, one kernel.  Reference W:156-189."""
    if convention not in _CONV:
        raise ValueError(f"polar_to_xy(): invalid convention={convention}!")
    return _b.execute("wind_polar_to_xy", (magnitude, direction), (_CONV[convention],))


def w_from_omega(omega, t, p):
    """Hydrostatic vertical velocity (m/s) from pressure velocity.  Reference W:192-222."""
    return _b.execute("w_from_omega", (omega, t, p))


def coriolis(lat):
    """Coriolis parameter 2 Omega sin(lat).  Reference W:225-251."""
    return _b.execute("coriolis", (lat,))
