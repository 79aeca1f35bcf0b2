"""Shared table of parity cases for the thermo hot path.

One ``Case`` = one public function of the reference API with one choice of
keyword options, plus the names of the input fields it is fed.  The same table drives

* tests/golden/make_golden.py   (runs the live reference in the build container -> fixtures),
* tests/test_oracle_golden.py   (oracle vs fixtures, CPU),
* tests/test_hostmath.py        (host instantiation of the device math header vs oracle, CPU),
* tests/test_gpu_parity.py      (CUDA path through the C-ABI vs oracle and fixtures, GPU).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

T0 = 273.16
TI = T0 - 23


@dataclass(frozen=True)
class Case:
    fn: str
    args: tuple
    kwargs: dict = field(default_factory=dict)
    n_out: int = 1
    iterative: str = ""  # "", "bisect" or "newton"

    @property
    def id(self):
        kw = ",".join(f"{k}={v}" for k, v in sorted(self.kwargs.items()))
        a = ",".join(self.args)
        return f"{self.fn}({a}{';' + kw if kw else ''})"


def _cases():
    c = []
    add = c.append
    add(Case("celsius_to_kelvin", ("tc",)))
    add(Case("kelvin_to_celsius", ("t",)))
    add(Case("specific_humidity_from_mixing_ratio", ("w",)))
    add(Case("mixing_ratio_from_specific_humidity", ("q",)))
    add(Case("vapour_pressure_from_specific_humidity", ("q", "p")))
    add(Case("vapour_pressure_from_mixing_ratio", ("w", "p")))
    add(Case("specific_humidity_from_vapour_pressure", ("e", "p")))
    add(Case("specific_humidity_from_vapour_pressure", ("e", "p"), {"eps": 50.0}))
    add(Case("mixing_ratio_from_vapour_pressure", ("e", "p")))
    add(Case("mixing_ratio_from_vapour_pressure", ("e", "p"), {"eps": 50.0}))
    for ph in ("mixed", "water", "ice"):
        add(Case("saturation_vapour_pressure", ("t",), {"phase": ph}))
        add(Case("saturation_vapour_pressure_slope", ("t",), {"phase": ph}))
        add(Case("saturation_mixing_ratio", ("t", "p"), {"phase": ph}))
        add(Case("saturation_specific_humidity", ("t", "p"), {"phase": ph}))
        add(Case("saturation_mixing_ratio_slope", ("t", "p"), {"phase": ph}))
        add(Case("saturation_specific_humidity_slope", ("t", "p"), {"phase": ph}))
    add(Case("saturation_mixing_ratio_slope", ("t", "p"), {"eps": 50.0}))
    add(Case("saturation_specific_humidity_slope", ("t", "p"), {"eps": 50.0}))
    add(Case("temperature_from_saturation_vapour_pressure", ("es",)))
    add(Case("relative_humidity_from_dewpoint", ("t", "td")))
    add(Case("relative_humidity_from_specific_humidity", ("t", "q", "p")))
    add(Case("specific_humidity_from_dewpoint", ("td", "p")))
    add(Case("mixing_ratio_from_dewpoint", ("td", "p")))
    add(Case("specific_humidity_from_relative_humidity", ("t", "r", "p")))
    add(Case("dewpoint_from_relative_humidity", ("t", "r")))
    add(Case("dewpoint_from_specific_humidity", ("q", "p")))
    add(Case("virtual_temperature", ("t", "q")))
    add(Case("virtual_potential_temperature", ("t", "q", "p")))
    add(Case("potential_temperature", ("t", "p")))
    add(Case("temperature_from_potential_temperature", ("th", "p")))
    add(Case("pressure_on_dry_adiabat", ("t", "t_def", "p_def")))
    add(Case("temperature_on_dry_adiabat", ("p", "t_def", "p_def")))
    for m in ("davies", "bolton"):
        add(Case("lcl_temperature", ("t", "td"), {"method": m}))
        add(Case("lcl", ("t", "td", "p"), {"method": m}, n_out=2))
    for m in ("ifs", "bolton35", "bolton39"):
        add(Case("ept_from_dewpoint", ("t", "td", "p"), {"method": m}))
        add(Case("ept_from_specific_humidity", ("t", "q", "p"), {"method": m}))
        add(Case("saturation_ept", ("t", "p"), {"method": m}))
        for tm in ("bisect", "newton"):
            kw = {"ept_method": m, "t_method": tm}
            add(Case("temperature_on_moist_adiabat", ("ept", "p"), kw, iterative=tm))
            add(Case("wet_bulb_temperature_from_dewpoint", ("t", "td", "p"), kw, iterative=tm))
            add(Case("wet_bulb_temperature_from_specific_humidity", ("t", "q", "p"), kw, iterative=tm))
        for tm in ("direct", "bisect", "newton"):
            kw = {"ept_method": m, "t_method": tm}
            it = "" if tm == "direct" else tm
            add(Case("wet_bulb_potential_temperature_from_dewpoint", ("t", "td", "p"), kw, iterative=it))
            add(Case("wet_bulb_potential_temperature_from_specific_humidity", ("t", "q", "p"), kw, iterative=it))
    add(Case("specific_gas_constant", ("q",)))
    return c


CASES = _cases()
CASE_BY_ID = {c.id: c for c in CASES}


# ------------------------------------------------------------------------------------------
# Input sets.  All are plain float64 numpy dicts; callers cast to float32 when they need to.
# ------------------------------------------------------------------------------------------
def _es_mixed_simple(t):
    """Local helper for building *physical* q (not used as an oracle)."""
    ew = 611.21 * np.exp(17.502 * (t - T0) / (t - 32.19))
    ei = 611.21 * np.exp(22.587 * (t - T0) / (t + 0.7))
    a = np.clip((t - TI) / (T0 - TI), 0.0, 1.0) ** 2
    return a * ew + (1 - a) * ei


def random_inputs(n, seed=0):
    """ERA5/IFS-like random points: t~U(200,320) K, p~U(1e3,1.05e5) Pa, physical q."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(200.0, 320.0, n)
    p = rng.uniform(1.0e3, 1.05e5, n)
    es = _es_mixed_simple(t)
    qs = 0.621981 * es / np.maximum(p - 0.378019 * es, 1.0)
    q = np.minimum(rng.uniform(1e-6, 0.02, n), 0.95 * np.abs(qs))
    td = t - rng.uniform(0.0, 30.0, n)
    r = rng.uniform(1.0, 100.0, n)
    w = q / (1 - q)
    e = p * q / (0.621981 + 0.378019 * q)
    ept = rng.uniform(220.0, 500.0, n)
    th = t * (1e5 / p) ** 0.285691
    t_def = rng.uniform(250.0, 310.0, n)
    p_def = rng.uniform(7.0e4, 1.05e5, n)
    return dict(t=t, tc=t - T0, td=td, q=q, r=r, p=p, w=w, e=e, es=es, ept=ept, th=th, t_def=t_def, p_def=p_def)


def edge_inputs(n=384, seed=7):
    """Special values mixed at random: NaN/inf/0/negatives, the TI and T0 band edges and their
    floating-point neighbours, p-e straddling the eps rule, subnormals."""
    rng = np.random.default_rng(seed)
    nxt = np.nextafter
    tvals = [np.nan, np.inf, -np.inf, 0.0, -10.0, TI, nxt(TI, 0), nxt(TI, 1e3), T0, nxt(T0, 0), nxt(T0, 1e3),
             32.19, 56.0, 180.0, 233.16, 250.0, 260.0, 273.15, 290.0, 320.0, 373.15, 1e-310, 5000.0]
    pvals = [np.nan, np.inf, 0.0, -1.0e5, 1e-5, 1.0, 50.0, 611.21, 611.2101, 1.0e3, 5.0e4, 8.5e4, 1.0e5, 1.02e5,
             1.876e5, 3.0e5, 1e-310, 1e300]
    qvals = [np.nan, 0.0, -0.01, 1.0, 0.5, 1e-320, 1e-8, 1e-4, 0.003, 0.01, 0.02, 2.0, np.inf]
    rvals = [np.nan, 0.0, -5.0, 1e-3, 10.0, 50.0, 100.0, 120.0, np.inf]
    evals = [np.nan, 0.0, -1.0, 1e-4, 611.21, 2.0e3, 4.99e4, 5.0e4 - 1e-4, 5.0e4 - 5e-5, 5.0e4, 1.0e5, np.inf, 1e-310]
    eptvals = [np.nan, 0.0, -300.0, 150.0, 220.0, 273.16, 300.0, 330.0, 400.0, 600.0, 900.0, 2000.0, np.inf]

    def pick(v):
        return np.asarray(v, dtype=np.float64)[rng.integers(0, len(v), n)]

    t = pick(tvals)
    td = np.where(rng.random(n) < 0.5, t - rng.uniform(0, 20, n), pick(tvals))
    out = dict(t=t, tc=t - T0, td=td, q=pick(qvals), r=pick(rvals), p=pick(pvals), w=pick(qvals), e=pick(evals),
               es=pick(evals), ept=pick(eptvals), th=pick(tvals), t_def=pick(tvals), p_def=pick(pvals))
    return out

