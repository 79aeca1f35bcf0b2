"""CPU oracle for the elementwise functions of ``earthkit.meteo.wind``.  TEST INFRASTRUCTURE, NOT PRODUCT.

SURVEY.md 8(f)-3.  numpy restatement of the reference's ``src/earthkit/meteo/wind/array/wind.py`` ("W") and the
constants it uses (``constants/constants.py`` "C"), same operation order.  Pinned bit-identical against the live
reference by tests/golden/make_golden.py (PINNING.json "wind") and against the reference tests' known-answer
vectors (tests/wind/test_wind.py).  ``windrose`` (a 2-D histogram, W:254-328) is not part of the elementwise path.
"""
import numpy as np

SOLAR_DAY = 86400  # C:60
SIDERAL_YEAR = 365.25 * SOLAR_DAY * 2 * np.pi / 6.283076  # C:63
SIDERAL_DAY = SOLAR_DAY / (1.0 + SOLAR_DAY / SIDERAL_YEAR)  # C:67
OMEGA = 2.0 * np.pi / SIDERAL_DAY  # C:71
DEGREE = 180.0 / np.pi  # C:75
RADIAN = 1.0 / DEGREE  # C:78
RD, G = 287.0597, 9.80665  # C:22, C:53


def speed(u, v):
    """W:15-34"""
    return np.hypot(np.asarray(u), np.asarray(v))


def _direction_meteo(u, v):
    """W:37-49"""
    minus_pi2 = -np.pi / 2.0
    d = np.asarray(np.arctan2(np.asarray(v), np.asarray(u)))
    d = np.array(d, copy=True, ndmin=0)
    m = d <= minus_pi2
    d[m] = (minus_pi2 - d[m]) * DEGREE
    m = ~m
    d[m] = (1.5 * np.pi - d[m]) * DEGREE
    return d


def _direction_polar(u, v, to_positive):
    """W:52-61"""
    d = np.arctan2(np.asarray(v), np.asarray(u)) * DEGREE
    if to_positive:
        d = np.array(d, copy=True, ndmin=0)
        m = d < 0
        d[m] = 360.0 + d[m]
    return d


def direction(u, v, convention="meteo", to_positive=True):
    """W:64-104"""
    if convention == "meteo":
        return _direction_meteo(u, v)
    if convention == "polar":
        return _direction_polar(u, v, to_positive)
    raise ValueError(f"direction(): invalid convention={convention}!")


def xy_to_polar(x, y, convention="meteo"):
    """W:107-135"""
    return speed(x, y), direction(x, y, convention=convention)


def polar_to_xy(magnitude, direction, convention="meteo"):
    """W:138-189"""
    magnitude, direction = np.asarray(magnitude), np.asarray(direction)
    if convention == "meteo":
        a = (270.0 - direction) * RADIAN
    elif convention == "polar":
        a = direction * RADIAN
    else:
        raise ValueError(f"polar_to_xy(): invalid convention={convention}!")
    return magnitude * np.cos(a), magnitude * np.sin(a)


def w_from_omega(omega, t, p):
    """W:192-222"""
    with np.errstate(all="ignore"):
        return (-RD / G) * (omega * t / p)


def coriolis(lat):
    """W:225-251"""
    return 2 * OMEGA * np.sin(np.asarray(lat) * RADIAN)
