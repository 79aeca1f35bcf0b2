"""CPU oracle for the earthkit-meteo thermo hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.

A numpy restatement of the algorithm in the reference's
``src/earthkit/meteo/thermo/array/thermo.py`` (T), ``.../thermo/array/es_comp.py`` (E) and
``src/earthkit/meteo/constants/constants.py`` (C).  Every function cites the reference lines it
follows and keeps the reference's *operation order* so that, on numpy float64, results are
bit-identical to the reference (pinned: ``tests/golden/make_golden.py`` compares this file
with the live reference on seeded inputs and with the reference's golden CSVs; the outcome is
recorded in tests/golden/PINNING.json and re-checked by tests/test_oracle_golden.py against the
committed fixtures).

Parity status: PINNED for float64 (reference golden CSVs + live reference outputs).
float32 is pinned only against live reference outputs (the reference has no fp32 thermo test).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product package (``ek_thermo``) never does; it has no CPU path at all.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import polynomial as _npoly

# --- constants, verbatim values (C:22-50, E:14-20) -------------------------------------------
RD = 287.0597  # C:22
RV = 461.51  # C:26
CPD = 1004.79  # C:30
LV = 2.5008e6  # C:38
KAPPA = 0.285691  # C:41 (a literal, not Rd/c_pd)
P0 = 1e5  # C:44
EPS = 0.621981  # C:47 (a literal, not Rd/Rv)
T0 = 273.16  # C:50, E:19
ES_C1 = 611.21  # E:14
ES_C3W = 17.502  # E:15
ES_C4W = 32.19  # E:16
ES_C3I = 22.587  # E:17
ES_C4I = -0.7  # E:18
TI = T0 - 23  # E:20  (250.16000000000003)
LAMBDA = 1.0 / KAPPA  # T:1022

EPT_METHODS = ("ifs", "bolton35", "bolton39")  # T:1319-1323
PHASES = ("mixed", "water", "ice")  # E:22

_QUIET = dict(divide="ignore", invalid="ignore", over="ignore", under="ignore")


def _quiet(fn):
    def wrapped(*a, **k):
        with np.errstate(**_QUIET):
            return fn(*a, **k)

    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


def _nan_where(v, cond):
    """``v[cond] = nan`` on a fresh array (the reference always masks a temporary)."""
    v = np.array(v, copy=True, ndmin=0)
    if v.dtype.kind != "f":
        v = v.astype(np.float64)
    v[np.broadcast_to(cond, v.shape)] = np.nan
    return v


# --- simple conversions -----------------------------------------------------------------------
def celsius_to_kelvin(t):
    """T:21-35"""
    return t + T0


def kelvin_to_celsius(t):
    """T:38-52"""
    return t - T0


@_quiet
def specific_humidity_from_mixing_ratio(w):
    """T:55-77"""
    return w / (1 + w)


@_quiet
def mixing_ratio_from_specific_humidity(q):
    """T:80-102"""
    return q / (1 - q)


@_quiet
def vapour_pressure_from_specific_humidity(q, p):
    """T:105-131"""
    c = EPS * (1.0 / EPS - 1.0)
    return (p * q) / (EPS + c * q)


@_quiet
def vapour_pressure_from_mixing_ratio(w, p):
    """T:134-159"""
    return (p * w) / (EPS + w)


@_quiet
def specific_humidity_from_vapour_pressure(e, p, eps=1e-4):
    """T:162-196 -- NaN where p - e < eps; eps <= 0 is a ValueError."""
    if eps <= 0:
        raise ValueError(f"specific_humidity_from_vapour_pressure(): eps={eps} must be > 0")
    v = _nan_where(p + (EPS - 1) * e, np.asarray(p - e) < eps)
    return EPS * e / v


@_quiet
def mixing_ratio_from_vapour_pressure(e, p, eps=1e-4):
    """T:199-232"""
    if eps <= 0:
        raise ValueError(f"mixing_ratio_from_vapour_pressure(): eps={eps} must be > 0")
    v = np.asarray(p - e)
    v = _nan_where(v, v < eps)
    return EPS * e / v


# --- saturation vapour pressure (E) -----------------------------------------------------------
def _es_water(t):
    """E:133-134"""
    return ES_C1 * np.exp(ES_C3W * (t - T0) / (t - ES_C4W))


def _es_ice(t):
    """E:137-138"""
    return ES_C1 * np.exp(ES_C3I * (t - T0) / (t - ES_C4I))


def _es_water_slope(t):
    """E:169-170"""
    return _es_water(t) * (ES_C3W * (T0 - ES_C4W)) / np.square(t - ES_C4W)


def _es_ice_slope(t):
    """E:173-174"""
    return _es_ice(t) * (ES_C3I * (T0 - ES_C4I)) / np.square(t - ES_C4I)


def _phase_split(t):
    """Masks of E:154,158,162: ice (t<=TI), water (t>=T0), blend (everything else, NaN included)."""
    ice = t <= TI
    water = t >= T0
    return ice, water, ~(ice | water)


def _es_mixed(t):
    """E:141-166 -- gather/compute/scatter per band, exactly as the reference does."""
    t = np.asarray(t)
    out = np.zeros_like(t, dtype=t.dtype)
    ice, water, blend = _phase_split(t)
    out[ice] = _es_ice(t[ice])
    out[water] = _es_water(t[water])
    tb = t[blend]
    alpha = np.square((tb - TI) / (T0 - TI))
    out[blend] = alpha * _es_water(tb) + (1.0 - alpha) * _es_ice(tb)
    return out


def _es_mixed_slope(t):
    """E:177-200"""
    t = np.asarray(t)
    out = np.zeros_like(t, dtype=t.dtype)
    ice, water, blend = _phase_split(t)
    out[ice] = _es_ice_slope(t[ice])
    out[water] = _es_water_slope(t[water])
    tb = t[blend]
    alpha = np.square((tb - TI) / (T0 - TI))
    d_alpha = (2.0 / (T0 - TI) ** 2) * (tb - TI)
    out[blend] = (
        d_alpha * _es_water(tb) + alpha * _es_water_slope(tb) - d_alpha * _es_ice(tb) + (1.0 - alpha) * _es_ice_slope(tb)
    )
    return out


_ES = {"mixed": _es_mixed, "water": _es_water, "ice": _es_ice}
_ES_SLOPE = {"mixed": _es_mixed_slope, "water": _es_water_slope, "ice": _es_ice_slope}


@_quiet
def saturation_vapour_pressure(t, phase="mixed"):
    """T:235-279 -> E:31-79.  An unknown phase silently yields None (E:74-79; check_phase is never called)."""
    fn = _ES.get(phase)
    return None if fn is None else fn(t)


@_quiet
def saturation_vapour_pressure_slope(t, phase="mixed"):
    """T:344-364 -> E:82-106"""
    fn = _ES_SLOPE.get(phase)
    return None if fn is None else fn(t)


def saturation_mixing_ratio(t, p, phase="mixed"):
    """T:282-310"""
    return mixing_ratio_from_vapour_pressure(saturation_vapour_pressure(t, phase=phase), p)


def saturation_specific_humidity(t, p, phase="mixed"):
    """T:313-341"""
    return specific_humidity_from_vapour_pressure(saturation_vapour_pressure(t, phase=phase), p)


@_quiet
def saturation_mixing_ratio_slope(t, p, es=None, es_slope=None, phase="mixed", eps=1e-4):
    """T:367-415"""
    if eps <= 0:
        raise ValueError(f"saturation_mixing_ratio_slope(): eps={eps} must be > 0")
    if es is None:
        es = saturation_vapour_pressure(t, phase=phase)
    if es_slope is None:
        es_slope = saturation_vapour_pressure_slope(t, phase=phase)
    v = np.asarray(p - es)
    v = _nan_where(v, v < eps)
    return EPS * es_slope * p / np.square(v)


@_quiet
def saturation_specific_humidity_slope(t, p, es=None, es_slope=None, phase="mixed", eps=1e-4):
    """T:418-467"""
    if eps <= 0:
        raise ValueError(f"saturation_specific_humidity_slope(): eps={eps} must be > 0")
    if es is None:
        es = saturation_vapour_pressure(t, phase=phase)
    if es_slope is None:
        es_slope = saturation_vapour_pressure_slope(t, phase=phase)
    v = _nan_where(np.square(p + es * (EPS - 1.0)), np.asarray(p - es) < eps)
    return EPS * es_slope * p / v


@_quiet
def temperature_from_saturation_vapour_pressure(es):
    """T:470-491 -> E:109-130 (always the water formula; es=0 -> NaN)."""
    v = np.log(es / ES_C1)
    return (v * ES_C4W - ES_C3W * T0) / (v - ES_C3W)


# --- humidity / dewpoint conversions ----------------------------------------------------------
@_quiet
def relative_humidity_from_dewpoint(t, td):
    """T:494-521"""
    e = saturation_vapour_pressure(td, phase="water")
    es = saturation_vapour_pressure(t, phase="water")
    return 100.0 * e / es


@_quiet
def relative_humidity_from_specific_humidity(t, q, p):
    """T:524-556"""
    svp = saturation_vapour_pressure(t)
    e = vapour_pressure_from_specific_humidity(q, p)
    return 100.0 * e / svp


def specific_humidity_from_dewpoint(td, p):
    """T:559-591"""
    return specific_humidity_from_vapour_pressure(saturation_vapour_pressure(td, phase="water"), p)


def mixing_ratio_from_dewpoint(td, p):
    """T:594-626"""
    return mixing_ratio_from_vapour_pressure(saturation_vapour_pressure(td, phase="water"), p)


@_quiet
def specific_humidity_from_relative_humidity(t, r, p):
    """T:629-663"""
    e = r * saturation_vapour_pressure(t) / 100.0
    return specific_humidity_from_vapour_pressure(e, p)


@_quiet
def dewpoint_from_relative_humidity(t, r):
    """T:666-699"""
    es = saturation_vapour_pressure(t, phase="water") * r / 100.0
    return temperature_from_saturation_vapour_pressure(es)


def dewpoint_from_specific_humidity(q, p):
    """T:702-735"""
    return temperature_from_saturation_vapour_pressure(vapour_pressure_from_specific_humidity(q, p))


# --- temperatures on dry adiabats -------------------------------------------------------------
def virtual_temperature(t, q):
    """T:738-764"""
    c1 = (1.0 - EPS) / EPS
    return t * (1.0 + c1 * q)


def virtual_potential_temperature(t, q, p):
    """T:767-798"""
    c1 = (1.0 - EPS) / EPS
    return potential_temperature(t, p) * (1.0 + c1 * q)


@_quiet
def potential_temperature(t, p):
    """T:801-829"""
    t = np.asarray(t)
    p = np.asarray(p)
    return t * np.power(P0 / p, KAPPA)


@_quiet
def temperature_from_potential_temperature(th, p):
    """T:832-858"""
    return th * np.power(p / P0, KAPPA)


@_quiet
def pressure_on_dry_adiabat(t, t_def, p_def):
    """T:861-889"""
    return p_def * np.power(t / t_def, 1 / KAPPA)


@_quiet
def temperature_on_dry_adiabat(p, t_def, p_def):
    """T:892-920"""
    return t_def * np.power(p / p_def, KAPPA)


@_quiet
def lcl_temperature(t, td, method="davies"):
    """T:923-968 -- both variants are closed-form."""
    if method == "davies":
        return td - (0.212 + 1.571e-3 * (td - T0) - 4.36e-4 * (t - T0)) * (t - td)
    if method == "bolton":
        return 56.0 + 1 / (1 / (td - 56) + np.log(t / td) / 800)
    raise ValueError(f"lcl_temperature: invalid method={method} specified!")


def lcl(t, td, p, method="davies"):
    """T:971-1000 -- returns the tuple (t_lcl, p_lcl)."""
    t_lcl = lcl_temperature(t, td, method=method)
    return t_lcl, pressure_on_dry_adiabat(t_lcl, t, p)


def specific_gas_constant(q):
    """T:1678-1707"""
    return RD + (RV - RD) * q


# --- equivalent potential temperature: three formulations (T:1162-1323) ------------------------
# The reference keeps per-call scratch in a _ThermoState object (T:1003-1017) and three _EptComp
# subclasses.  Here each formulation is a small record of plain functions over a dict ``s`` that
# plays the role of the scratch state (keys t, p, td, q, es, ws, qs, c_tw).
K0_IFS = LV / CPD  # T:1164
B35_K0, B35_K3 = 2675.0, 0.28  # T:1202-1203
B39_K0, B39_K1, B39_K2, B39_K4 = 3036.0, 1.78, 0.448, 0.28  # T:1263-1266


def _state(**kw):
    s = dict(t=None, td=None, q=None, p=None, es=None, ws=None, qs=None, c_tw=None)
    s.update(kw)
    return s


class _Ifs:
    mixing_ratio_based = False  # T:1166-1167

    @staticmethod
    def ept(s):  # T:1169-1175
        th = potential_temperature(s["t"], s["p"])
        t_lcl = lcl_temperature(s["t"], s["td"], method="davies")
        if s["q"] is None:
            s["q"] = specific_humidity_from_dewpoint(s["td"], s["p"])
        return th * np.exp(K0_IFS * s["q"] / t_lcl)

    @staticmethod
    def th_sat(s):  # T:1177-1178
        return potential_temperature(s["t"], s["p"])

    @staticmethod
    def g_sat(s, scale=1.0):  # T:1180-1182 (recomputes qs on every call)
        qs = saturation_specific_humidity(s["t"], s["p"])
        return (scale * K0_IFS) * qs / s["t"]

    @staticmethod
    def d_g_sat(s):  # T:1184-1190
        if s["qs"] is None:
            s["qs"] = saturation_specific_humidity(s["t"], s["p"])
        return -K0_IFS * s["qs"] / (s["t"] ** 2) + K0_IFS * saturation_specific_humidity_slope(s["t"], s["p"]) / s["t"]

    @classmethod
    def f(cls, s):  # T:1192-1194
        return s["c_tw"] * np.exp(cls.g_sat(s, scale=-LAMBDA))

    @classmethod
    def d_lnf(cls, s):  # T:1196-1197
        return -LAMBDA * (1 / s["t"] + cls.d_g_sat(s))


class _Bolton35:
    mixing_ratio_based = True

    @staticmethod
    def ept(s):  # T:1205-1213
        t_lcl = lcl_temperature(s["t"], s["td"], method="bolton")
        if s["q"] is None:
            w = mixing_ratio_from_dewpoint(s["td"], s["p"])
        else:
            w = mixing_ratio_from_specific_humidity(s["q"])
        th = s["t"] * np.power(P0 / s["p"], KAPPA * (1 - B35_K3 * w))
        return th * np.exp(B35_K0 * w / t_lcl)

    @staticmethod
    def _ws(s):
        if s["ws"] is None:
            s["ws"] = saturation_mixing_ratio(s["t"], s["p"])
        return s["ws"]

    @classmethod
    def th_sat(cls, s):  # T:1215-1219
        ws = cls._ws(s)
        return s["t"] * np.power(P0 / s["p"], KAPPA * (1 - B35_K3 * ws))

    @classmethod
    def g_sat(cls, s, scale=1.0):  # T:1221-1224
        return (scale * B35_K0) * cls._ws(s) / s["t"]

    @staticmethod
    def d_g_sat(s):  # T:1226-1231
        return -B35_K0 * s["ws"] / np.square(s["t"]) + B35_K0 * saturation_mixing_ratio_slope(s["t"], s["p"]) / s["t"]

    @classmethod
    def f(cls, s):  # T:1233-1242
        return s["c_tw"] * np.power(s["p"] / P0, B35_K3 * s["ws"]) * np.exp(cls.g_sat(s, scale=-LAMBDA))

    @classmethod
    def d_lnf(cls, s):  # T:1244-1250 -- the es slope (not the ws slope) multiplies K3*log(p/p0): kept as is
        return -LAMBDA * (
            1 / s["t"] + B35_K3 * np.log(s["p"] / P0) * saturation_vapour_pressure_slope(s["t"]) + cls.d_g_sat(s)
        )


class _Bolton39:
    mixing_ratio_based = True

    @staticmethod
    def ept(s):  # T:1268-1278
        t_lcl = lcl_temperature(s["t"], s["td"], method="bolton")
        if s["q"] is None:
            w = mixing_ratio_from_dewpoint(s["td"], s["p"])
        else:
            w = mixing_ratio_from_specific_humidity(s["q"])
        e = vapour_pressure_from_mixing_ratio(w, s["p"])
        th = potential_temperature(s["t"], s["p"] - e) * np.power(s["t"] / t_lcl, B39_K4 * w)
        return th * np.exp((B39_K0 / t_lcl - B39_K1) * w * (1.0 + B39_K2 * w))

    @staticmethod
    def _es(s):  # T:1281-1284 / T:1288-1291: masked only when not already cached
        if s["es"] is None:
            es = saturation_vapour_pressure(s["t"])
            s["es"] = _nan_where(es, s["p"] - es < 1e-4)
        return s["es"]

    @classmethod
    def th_sat(cls, s):  # T:1280-1285
        return potential_temperature(s["t"], s["p"] - cls._es(s))

    @classmethod
    def g_sat(cls, s, scale=1.0):  # T:1287-1295
        ws = mixing_ratio_from_vapour_pressure(cls._es(s), s["p"])
        return ((scale * B39_K0) / s["t"] - (scale * B39_K1)) * ws * (1.0 + B39_K2 * ws)

    @staticmethod
    def d_g_sat(s):  # T:1297-1302
        ws, t = s["ws"], s["t"]
        return -B39_K0 * (ws + B39_K2 * np.square(ws)) / (np.square(t)) + (B39_K0 / t - B39_K1) * (
            1 + (2 * B39_K2) * ws
        ) * saturation_mixing_ratio_slope(t, s["p"])

    @classmethod
    def f(cls, s):  # T:1304-1309
        return s["c_tw"] * (1 - s["es"] / s["p"]) * np.exp(cls.g_sat(s, scale=-LAMBDA))

    @classmethod
    def d_lnf(cls, s):  # T:1311-1316
        return -LAMBDA * (
            1 / s["t"] + KAPPA * saturation_vapour_pressure_slope(s["t"]) / (s["p"] - s["es"]) + cls.d_g_sat(s)
        )


_FORMULATIONS = {"ifs": _Ifs, "bolton35": _Bolton35, "bolton39": _Bolton39}


def _formulation(method):
    """T:1024-1026: an unknown method is a KeyError (dict lookup)."""
    return _FORMULATIONS[method]


@_quiet
def _compute_ept(method, t=None, td=None, q=None, p=None):
    """T:1031-1040"""
    fm = _formulation(method)
    if td is None and q is None:
        raise ValueError("ept: either td or q must have a valid value!")
    if td is None:
        td = dewpoint_from_specific_humidity(q, p)
    return fm.ept(_state(t=t, td=td, q=q, p=p))


def ept_from_dewpoint(t, td, p, method="ifs"):
    """T:1326-1387"""
    return _compute_ept(method, t=t, td=td, p=p)


def ept_from_specific_humidity(t, q, p, method="ifs"):
    """T:1390-1415"""
    return _compute_ept(method, t=t, q=q, p=p)


@_quiet
def saturation_ept(t, p, method="ifs"):
    """T:1418-1469 -> T:1042-1045 (th_sat first, then G_sat, sharing the scratch state)."""
    fm = _formulation(method)
    s = _state(t=t, p=p)
    return fm.th_sat(s) * np.exp(fm.g_sat(s))


_WBPT_A = [7.101574, -20.68208, 16.11182, 2.574631, -5.205688]  # T:1051
_WBPT_B = [1.0, -3.552497, 3.781782, -0.6899655, -0.5929340]  # T:1052


@_quiet
def _wbpt_direct(ept):
    """T:1047-1053 -- rational fit of Davies-Jones (2008) Eq 3.8; ascending-order coefficients."""
    x = ept / 273.16
    return ept - np.exp(_npoly.polyval(x, _WBPT_A) / _npoly.polyval(x, _WBPT_B))


@_quiet
def _t_on_ma_bisect(fm, ept, p):
    """T:1055-1079 -- 12 fixed halvings of dt=120 K starting at T0-20; flat 1-D iterate."""
    ept = np.asarray(ept)
    p = np.asarray(p)
    size = np.size(ept) if np.size(ept) > np.size(p) else np.size(p)
    t = np.full(size, T0 - 20, dtype=ept.dtype)
    dt = 120.0
    for _ in range(12):
        s = _state(t=t, p=p)
        dt /= 2.0
        t += np.sign(ept * np.exp(fm.g_sat(s, scale=-1.0)) - fm.th_sat(s)) * dt
    return t


_K1 = [-53.737, 137.81, -38.5]  # T:1092
_K2 = [-0.384, 56.831, -4.392]  # T:1097


@_quiet
def _t_on_ma_newton(fm, ept, p):
    """T:1081-1159 -- Davies-Jones (2008) first guess by regime, then exactly one Newton step."""
    ept = np.asarray(ept)
    p = np.asarray(p)
    if np.size(ept) > np.size(p):
        p = np.full(np.size(ept), p, dtype=ept.dtype)
    A = 2675
    t0 = 273.16

    def d_of_p(pv):
        return 1.0 / (0.1859e-5 * pv + 0.6512)

    tw = np.array(ept, copy=True)
    pp = np.power(p / P0, KAPPA)
    te = ept * pp
    c_te = np.power(t0 / te, LAMBDA)

    m = c_te > d_of_p(p)
    if np.any(m):
        es = saturation_vapour_pressure(te[m])
        ws = mixing_ratio_from_vapour_pressure(es, p[m])
        d_es = saturation_vapour_pressure_slope(te[m])
        tw[m] = te[m] - t0 - (A * ws) / (1 + A * ws * d_es / es)
    m = (1 <= c_te) & (c_te <= d_of_p(p))
    tw[m] = _npoly.polyval(pp[m], _K1) - _npoly.polyval(pp[m], _K2) * c_te[m]
    m = (0.4 <= c_te) & (c_te < 1)
    tw[m] = (_npoly.polyval(pp[m], _K1) - 1.21) - (_npoly.polyval(pp[m], _K2) - 1.21) * c_te[m]
    m = c_te < 0.4
    tw[m] = (_npoly.polyval(pp[m], _K1) - 2.66) - (_npoly.polyval(pp[m], _K2) - 1.21) * c_te[m] + 0.58 / c_te[m]
    tw = celsius_to_kelvin(tw)

    for _ in range(1):  # max_iter = 1 (T:1104)
        s = _state(t=tw, p=p)
        s["c_tw"] = np.power(t0 / s["t"], LAMBDA)
        s["es"] = saturation_vapour_pressure(s["t"])
        if fm.mixing_ratio_based:
            s["ws"] = mixing_ratio_from_vapour_pressure(s["es"], s["p"])
        else:
            s["qs"] = specific_humidity_from_vapour_pressure(s["es"], s["p"])
        f_val = fm.f(s)
        tw -= (f_val - c_te) / (f_val * fm.d_lnf(s))
    tw[tw <= 0] = np.nan  # T:1155
    return tw


def temperature_on_moist_adiabat(ept, p, ept_method="ifs", t_method="bisect"):
    """T:1472-1509"""
    fm = _formulation(ept_method)
    if t_method == "bisect":
        return _t_on_ma_bisect(fm, ept, p)
    if t_method == "newton":
        return _t_on_ma_newton(fm, ept, p)
    raise ValueError(f"temperature_on_moist_adiabat: invalid t_method={t_method} specified!")


def wet_bulb_temperature_from_dewpoint(t, td, p, ept_method="ifs", t_method="bisect"):
    """T:1512-1549"""
    ept = ept_from_dewpoint(t, td, p, method=ept_method)
    return temperature_on_moist_adiabat(ept, p, ept_method=ept_method, t_method=t_method)


def wet_bulb_temperature_from_specific_humidity(t, q, p, ept_method="ifs", t_method="bisect"):
    """T:1552-1590"""
    ept = ept_from_specific_humidity(t, q, p, method=ept_method)
    return temperature_on_moist_adiabat(ept, p, ept_method=ept_method, t_method=t_method)


def wet_bulb_potential_temperature_from_dewpoint(t, td, p, ept_method="ifs", t_method="direct"):
    """T:1593-1634"""
    ept = ept_from_dewpoint(t, td, p, method=ept_method)
    if t_method == "direct":
        _formulation(ept_method)
        return _wbpt_direct(ept)
    return temperature_on_moist_adiabat(ept, P0, ept_method=ept_method, t_method=t_method)


def wet_bulb_potential_temperature_from_specific_humidity(t, q, p, ept_method="ifs", t_method="direct"):
    """T:1637-1675"""
    ept = ept_from_specific_humidity(t, q, p, method=ept_method)
    if t_method == "direct":
        _formulation(ept_method)
        return _wbpt_direct(ept)
    return temperature_on_moist_adiabat(ept, P0, ept_method=ept_method, t_method=t_method)


# --- the fused suites of the new build, stated as compositions of the reference functions -----
def suite_tqp(t, q, p, ept_method="ifs"):
    """What the fused (t,q,p) kernel must equal, output by output (BASELINE.json configs[1]; "ept" / "wbpt" are the
    configs[2] pair: the reference chain T:1637-1675 -> T:1390-1415 -> T:1031-1040 with its default t_method)."""
    return {
        "theta": potential_temperature(t, p),
        "es": saturation_vapour_pressure(t),
        "rh": relative_humidity_from_specific_humidity(t, q, p),
        "td": dewpoint_from_specific_humidity(q, p),
        "tv": virtual_temperature(t, q),
        "w": mixing_ratio_from_specific_humidity(q),
        "e": vapour_pressure_from_specific_humidity(q, p),
        "thetav": virtual_potential_temperature(t, q, p),
        "ept": ept_from_specific_humidity(t, q, p, method=ept_method),
        "wbpt": wet_bulb_potential_temperature_from_specific_humidity(t, q, p, ept_method=ept_method, t_method="direct"),
    }


def suite_ttdp(t, td, p, ept_method="ifs"):
    """What the fused (t,td,p) kernel must equal, output by output."""
    q = specific_humidity_from_dewpoint(td, p)
    return {
        "theta": potential_temperature(t, p),
        "es": saturation_vapour_pressure(t),
        "rh": relative_humidity_from_dewpoint(t, td),
        "q": q,
        "tv": virtual_temperature(t, q),
        "w": mixing_ratio_from_dewpoint(td, p),
        "e": saturation_vapour_pressure(td, phase="water"),
        "thetav": virtual_potential_temperature(t, q, p),
        "ept": ept_from_dewpoint(t, td, p, method=ept_method),
        "wbpt": wet_bulb_potential_temperature_from_dewpoint(t, td, p, ept_method=ept_method, t_method="direct"),
    }


PUBLIC_NAMES = [
    "celsius_to_kelvin",
    "kelvin_to_celsius",
    "specific_humidity_from_mixing_ratio",
    "mixing_ratio_from_specific_humidity",
    "vapour_pressure_from_specific_humidity",
    "vapour_pressure_from_mixing_ratio",
    "specific_humidity_from_vapour_pressure",
    "mixing_ratio_from_vapour_pressure",
    "saturation_vapour_pressure",
    "saturation_mixing_ratio",
    "saturation_specific_humidity",
    "saturation_vapour_pressure_slope",
    "saturation_mixing_ratio_slope",
    "saturation_specific_humidity_slope",
    "temperature_from_saturation_vapour_pressure",
    "relative_humidity_from_dewpoint",
    "relative_humidity_from_specific_humidity",
    "specific_humidity_from_dewpoint",
    "mixing_ratio_from_dewpoint",
    "specific_humidity_from_relative_humidity",
    "dewpoint_from_relative_humidity",
    "dewpoint_from_specific_humidity",
    "virtual_temperature",
    "virtual_potential_temperature",
    "potential_temperature",
    "temperature_from_potential_temperature",
    "pressure_on_dry_adiabat",
    "temperature_on_dry_adiabat",
    "lcl_temperature",
    "lcl",
    "ept_from_dewpoint",
    "ept_from_specific_humidity",
    "saturation_ept",
    "temperature_on_moist_adiabat",
    "wet_bulb_temperature_from_dewpoint",
    "wet_bulb_temperature_from_specific_humidity",
    "wet_bulb_potential_temperature_from_dewpoint",
    "wet_bulb_potential_temperature_from_specific_humidity",
    "specific_gas_constant",
]
