"""Synthetic input fields shared by bench.py, tools/ and tests/ (SURVEY.md 8(d)).  No product code, no oracle code.

* ``random_inputs`` / ``edge_inputs`` live in tests/cases.py (the fixtures in tests/golden were generated from them).
* ``IfsField`` -- the ERA5/IFS-shaped generator of the benchmark workloads: real IFS L137 A/B coefficients, per-column
  surface pressure, a standard-atmosphere temperature profile with noise, physical humidity.  Every (member, level) slab
  has its own counter-style seed, so any rank -- or the CPU arm -- can regenerate any slab of a field on its own:
  the reference arm of bench.py times the reference on slabs of the very field the GPU arm times, and the sharded
  multi-GPU check recomputes slabs that another rank owns.  The generator is plain torch (never the product library).
  torch's CUDA and CPU generators produce different streams: "the same field" holds per device type (the GPU box has a
  GPU for both arms; without one the CPU generator draws from the same distribution).
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O1280_POINTS = 4 * 1280 * 1289  # 6 599 680 (octahedral reduced Gaussian grid O1280)
O640_POINTS = 4 * 640 * 649     # 1 661 440
N_LEVELS = 137

T0 = 273.16
TI = T0 - 23.0
EPS = 0.621981


def ifs_ab():
    """IFS L137 half-level coefficients (138 values each), re-packed from the reference's conf JSON by make_golden.py."""
    ab = np.load(os.path.join(ROOT, "tests", "golden", "ifs_l137_ab.npz"))
    return np.asarray(ab["A"], dtype=np.float64), np.asarray(ab["B"], dtype=np.float64)


def _qs_mixed(t, p):
    """Mixed-phase saturation specific humidity, only used to keep the synthetic q physical (not an oracle).  NaN where
    the reference's rule p - es < 1e-4 applies."""
    import torch

    ew = 611.21 * torch.exp(17.502 * (t - T0) / (t - 32.19))
    ei = 611.21 * torch.exp(22.587 * (t - T0) / (t + 0.7))
    a = ((t - TI) / (T0 - TI)).clamp(0.0, 1.0) ** 2
    es = a * ew + (1.0 - a) * ei
    qs = EPS * es / (p + (EPS - 1.0) * es)
    return torch.where((p - es) < 1e-4, torch.full_like(qs, float("nan")), qs)


class IfsField:
    """[members x levels, npl] synthetic model-level field, generated slab by slab.

    kind: "tqp" (t, q, p), "ttdp" (t, td, p) or "hybrid" (t, q with p = ph_k + 0.5 (ph_k+1 - ph_k), the reference's
    full-level pressure, vertical.py:663,708; the kernel is handed sp, A, B instead of p).
    Slab s = member s // levels, model level (s % levels) of the lowest `levels` levels (all 137 by default).
    """

    def __init__(self, kind, npl, levels=N_LEVELS, seed=0, device="cpu"):
        import torch

        self.kind, self.npl, self.levels, self.seed = kind, int(npl), int(levels), int(seed)
        self.device = torch.device(device)
        A, B = ifs_ab()
        if levels > N_LEVELS:
            raise ValueError("at most 137 levels per member")
        self.A_half, self.B_half = A[N_LEVELS - levels:], B[N_LEVELS - levels:]  # levels + 1 half-level values
        self._sp_cache = (None, None)

    def _gen(self, stream_id):
        import torch

        return torch.Generator(device=self.device).manual_seed((self.seed * 1_000_003 + stream_id) & 0x7FFFFFFFFFFF)

    def sp(self, member=0):
        """Surface pressure of a member, U(5e4, 1.05e5) Pa, float64 [npl]."""
        import torch

        if self._sp_cache[0] != member:
            g = self._gen(0x40000000 + member)
            self._sp_cache = (member, torch.empty(self.npl, dtype=torch.float64, device=self.device).uniform_(5.0e4, 1.05e5, generator=g))
        return self._sp_cache[1]

    def slab(self, s):
        """(t, h, p) of slab s as float64 tensors on the field's device; h is q (tqp, hybrid) or td (ttdp)."""
        import torch

        member, k = divmod(int(s), self.levels)
        sp = self.sp(member)
        g = self._gen(s)
        a0, a1, b0, b1 = self.A_half[k], self.A_half[k + 1], self.B_half[k], self.B_half[k + 1]
        if self.kind == "hybrid":
            ph0, ph1 = a0 + b0 * sp, a1 + b1 * sp
            p = ph0 + 0.5 * (ph1 - ph0)
        else:
            p = 0.5 * (a0 + a1) + 0.5 * (b0 + b1) * sp
        noise = torch.empty(self.npl, dtype=torch.float64, device=self.device).uniform_(-15.0, 15.0, generator=g)
        t = (288.15 * (p / 101325.0) ** 0.19 + noise).clamp_(180.0, 320.0)
        u = torch.empty(self.npl, dtype=torch.float64, device=self.device)
        if self.kind == "ttdp":
            h = t - u.uniform_(0.0, 30.0, generator=g)
        else:
            u.uniform_(1.0e-6, 0.02, generator=g)
            qs = _qs_mixed(t, p)
            h = torch.where(torch.isnan(qs), u, torch.minimum(u, 0.95 * qs.abs()))
        return t, h, p

    def slabs_numpy(self, slabs, dtype=np.float64):
        """The named slabs concatenated, as host numpy arrays [len(slabs) * npl] (for the CPU arms)."""
        parts = [[x.cpu().numpy().astype(dtype) for x in self.slab(s)] for s in slabs]
        return [np.ascontiguousarray(np.concatenate([pt[i] for pt in parts])) for i in range(3)]

    def materialise(self, first_slab, n_slabs, torch_dtype):
        """[n_slabs * npl] flat device arrays (t, h, p) of slabs first_slab ... first_slab + n_slabs - 1."""
        import torch

        n = n_slabs * self.npl
        out = [torch.empty(n, dtype=torch_dtype, device=self.device) for _ in range(3)]
        for j in range(n_slabs):
            for dst, src in zip(out, self.slab(first_slab + j)):
                dst[j * self.npl:(j + 1) * self.npl] = src  # rounds to the working dtype once, as a cast of the float64 field
        return out


def sample_levels(levels, n):
    """n model levels spread evenly over a column (top, mixed-phase band and boundary layer all represented)."""
    n = max(1, min(n, levels))
    return [int(round((j + 0.5) * levels / n - 0.5)) for j in range(n)]


def ifs_point_inputs(n_per_level, seed=0, levels=N_LEVELS):
    """A numpy dict with the names of tests/cases.random_inputs (t, td, q, r, p, w, e, es, ept, th, t_def, p_def, tc) drawn from
    the IFS-shaped distribution of IfsField: n_per_level columns x `levels` model levels, top levels at 1-100 Pa included
    (where the reference's NaN rule p - es < 1e-4 is live).  Pure numpy (PCG64): the same points on every machine, for
    the parity tests and tools/parity_report.py."""
    rng = np.random.default_rng(seed)
    A, B = ifs_ab()
    A, B = A[N_LEVELS - levels:], B[N_LEVELS - levels:]
    sp = rng.uniform(5.0e4, 1.05e5, n_per_level)
    a_f, b_f = 0.5 * (A[:-1] + A[1:]), 0.5 * (B[:-1] + B[1:])
    p = (a_f[:, None] + b_f[:, None] * sp[None, :]).reshape(-1)
    n = p.size
    t = np.clip(288.15 * (p / 101325.0) ** 0.19 + rng.uniform(-15.0, 15.0, n), 180.0, 320.0)
    ew = 611.21 * np.exp(17.502 * (t - T0) / (t - 32.19))
    ei = 611.21 * np.exp(22.587 * (t - T0) / (t + 0.7))
    al = np.clip((t - TI) / (T0 - TI), 0.0, 1.0) ** 2
    es = al * ew + (1.0 - al) * ei
    with np.errstate(all="ignore"):
        qs = np.where((p - es) < 1e-4, np.nan, EPS * es / (p + (EPS - 1.0) * es))
    u = rng.uniform(1.0e-6, 0.02, n)
    q = np.where(np.isnan(qs), u, np.minimum(u, 0.95 * np.abs(qs)))
    td = t - rng.uniform(0.0, 30.0, n)
    r = rng.uniform(1.0, 100.0, n)
    w = q / (1.0 - q)
    e = p * q / (EPS + (1.0 - EPS) * q)
    th = t * (1.0e5 / p) ** 0.285691
    ept = np.minimum(th * np.exp(2490.0 * q / np.maximum(td - 5.0, 100.0)), 1.0e5)  # plausible theta_e along the column (an input of the moist-adiabat functions)
    t_def = rng.uniform(250.0, 310.0, n)
    p_def = rng.uniform(7.0e4, 1.05e5, n)
    return dict(t=t, tc=t - T0, td=td, q=q, r=r, p=p, w=w, e=e, es=es, ept=ept, th=th, t_def=t_def, p_def=p_def)
