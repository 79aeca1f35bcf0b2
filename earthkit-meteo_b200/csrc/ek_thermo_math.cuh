// ek_thermo_math.cuh -- per-point thermodynamic formulas of the earthkit-meteo thermo hot path,
// written as templated __host__ __device__ functions over T in {double, float}.
//
// Reference citations: T = src/earthkit/meteo/thermo/array/thermo.py, E = .../array/es_comp.py,
// C = src/earthkit/meteo/constants/constants.py (all in the reference tree).  The formulas keep the
// reference's operation order; constants are the reference's literals (C:22-50, E:14-20) and the
// derived ones are evaluated in double exactly as Python evaluates them, then rounded to T once
// (which is what NumPy-2 weak-scalar promotion does for float32 arrays).
//
// Everything a thread needs lives in registers: the reference's _ThermoState scratch object
// (T:1003-1017) becomes local variables here.
//
// Math primitives: ek::m_exp / m_log / m_pow / m_div wrap libdevice (exact) by default.  The lean
// device-only variants are selected with -DEK_LEAN_MATH=1 (see ek_thermo_lean.cuh).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

#if defined(__CUDACC__)
#define EK_HD __host__ __device__ __forceinline__
#else
#define EK_HD inline
#endif

#ifndef EK_LEAN_MATH
#define EK_LEAN_MATH 0
#endif

namespace ek {

// ------------------------------------------------------------------------------------------
// math primitives
// ------------------------------------------------------------------------------------------
EK_HD double m_exp_exact(double x) { return ::exp(x); }
EK_HD float m_exp_exact(float x) { return ::expf(x); }
EK_HD double m_log_exact(double x) { return ::log(x); }
EK_HD float m_log_exact(float x) { return ::logf(x); }
EK_HD double m_pow_exact(double x, double y) { return ::pow(x, y); }
EK_HD float m_pow_exact(float x, float y) { return ::powf(x, y); }

}  // namespace ek

#if EK_LEAN_MATH && defined(__CUDA_ARCH__)
#include "ek_thermo_lean.cuh"
#endif

namespace ek {

#if EK_LEAN_MATH && defined(__CUDA_ARCH__)
template <typename T> EK_HD T m_exp(T x) { return lean::exp_(x); }
template <typename T> EK_HD T m_log(T x) { return lean::log_(x); }
template <typename T> EK_HD T m_pow(T x, T y) { return lean::pow_(x, y); }
template <typename T> EK_HD T m_div(T a, T b) { return lean::div_(a, b); }
#else
template <typename T> EK_HD T m_exp(T x) { return m_exp_exact(x); }
template <typename T> EK_HD T m_log(T x) { return m_log_exact(x); }
template <typename T> EK_HD T m_pow(T x, T y) { return m_pow_exact(x, y); }
template <typename T> EK_HD T m_div(T a, T b) { return a / b; }
#endif

template <typename T> EK_HD T m_nan() { return static_cast<T>(NAN); }
template <typename T> EK_HD T sq(T x) { return x * x; }
// numpy.sign: -1, 0, +1, NaN for NaN (T:1075 relies on the NaN case to poison the iterate)
template <typename T> EK_HD T m_sign(T x) { return (x > T(0)) ? T(1) : ((x < T(0)) ? T(-1) : ((x == T(0)) ? T(0) : x)); }

// ------------------------------------------------------------------------------------------
// constants (double expressions evaluated as Python does, rounded to T once)
// ------------------------------------------------------------------------------------------
namespace c {
constexpr double Rd = 287.0597;      // C:22
constexpr double Rv = 461.51;        // C:26
constexpr double c_pd = 1004.79;     // C:30
constexpr double Lv = 2.5008e6;      // C:38
constexpr double kappa = 0.285691;   // C:41
constexpr double p0 = 1e5;           // C:44
constexpr double eps = 0.621981;     // C:47
constexpr double T0 = 273.16;        // C:50, E:19
constexpr double C1 = 611.21;        // E:14
constexpr double C3W = 17.502;       // E:15
constexpr double C4W = 32.19;        // E:16
constexpr double C3I = 22.587;       // E:17
constexpr double C4I = -0.7;         // E:18
constexpr double TI = T0 - 23;       // E:20
constexpr double lambda = 1.0 / kappa;                 // T:1022
constexpr double c_vp = eps * (1.0 / eps - 1.0);       // T:130
constexpr double c1_tv = (1.0 - eps) / eps;            // T:763
constexpr double eps_m1 = eps - 1;                     // T:193, T:465
constexpr double K0_ifs = Lv / c_pd;                   // T:1164
constexpr double slope_w = C3W * (T0 - C4W);           // E:170
constexpr double slope_i = C3I * (T0 - C4I);           // E:174
constexpr double band = T0 - TI;                       // E:164
constexpr double d_alpha_c = 2.0 / (band * band);      // E:192  (2.0 / (T0 - TI) ** 2)
constexpr double C3W_T0 = C3W * T0;                    // E:130
constexpr double t0_dj = 273.16;                       // T:1049, T:1106
}  // namespace c

enum Phase : int { PHASE_MIXED = 0, PHASE_WATER = 1, PHASE_ICE = 2 };
enum LclMethod : int { LCL_DAVIES = 0, LCL_BOLTON = 1 };
enum EptMethod : int { EPT_IFS = 0, EPT_BOLTON35 = 1, EPT_BOLTON39 = 2 };
enum TMethod : int { TM_NONE = 0, TM_DIRECT = 1, TM_BISECT = 2, TM_NEWTON = 3 };

// ------------------------------------------------------------------------------------------
// saturation vapour pressure (E)
// ------------------------------------------------------------------------------------------
template <typename T> EK_HD T es_water(T t) {  // E:133-134
    return T(c::C1) * m_exp(m_div(T(c::C3W) * (t - T(c::T0)), t - T(c::C4W)));
}
template <typename T> EK_HD T es_ice(T t) {  // E:137-138
    return T(c::C1) * m_exp(m_div(T(c::C3I) * (t - T(c::T0)), t - T(c::C4I)));
}
template <typename T> EK_HD T es_water_slope(T t) {  // E:169-170
    return m_div(es_water(t) * T(c::slope_w), sq(t - T(c::C4W)));
}
template <typename T> EK_HD T es_ice_slope(T t) {  // E:173-174
    return m_div(es_ice(t) * T(c::slope_i), sq(t - T(c::C4I)));
}
// E:141-166.  t<=TI -> ice, t>=T0 -> water, everything else (NaN included) -> blend.
template <typename T> EK_HD T es_mixed(T t) {
    if (t <= T(c::TI)) return es_ice(t);
    if (t >= T(c::T0)) return es_water(t);
    T alpha = sq(m_div(t - T(c::TI), T(c::band)));
    return alpha * es_water(t) + (T(1.0) - alpha) * es_ice(t);
}
// E:177-200
template <typename T> EK_HD T es_mixed_slope(T t) {
    if (t <= T(c::TI)) return es_ice_slope(t);
    if (t >= T(c::T0)) return es_water_slope(t);
    T alpha = sq(m_div(t - T(c::TI), T(c::band)));
    T d_alpha = T(c::d_alpha_c) * (t - T(c::TI));
    return d_alpha * es_water(t) + alpha * es_water_slope(t) - d_alpha * es_ice(t) + (T(1.0) - alpha) * es_ice_slope(t);
}
// both at once for the callers that need es and its slope at the same t (shares the exps)
template <typename T> EK_HD void es_mixed_both(T t, T& es, T& des) {
    if (t <= T(c::TI)) {
        es = es_ice(t);
        des = m_div(es * T(c::slope_i), sq(t - T(c::C4I)));
    } else if (t >= T(c::T0)) {
        es = es_water(t);
        des = m_div(es * T(c::slope_w), sq(t - T(c::C4W)));
    } else {
        T ew = es_water(t), ei = es_ice(t);
        T dw = m_div(ew * T(c::slope_w), sq(t - T(c::C4W)));
        T di = m_div(ei * T(c::slope_i), sq(t - T(c::C4I)));
        T alpha = sq(m_div(t - T(c::TI), T(c::band)));
        T d_alpha = T(c::d_alpha_c) * (t - T(c::TI));
        es = alpha * ew + (T(1.0) - alpha) * ei;
        des = d_alpha * ew + alpha * dw - d_alpha * ei + (T(1.0) - alpha) * di;
    }
}
template <typename T> EK_HD T es_phase(T t, int phase) {  // E:31-79
    return phase == PHASE_WATER ? es_water(t) : (phase == PHASE_ICE ? es_ice(t) : es_mixed(t));
}
template <typename T> EK_HD T es_slope_phase(T t, int phase) {  // E:82-106
    return phase == PHASE_WATER ? es_water_slope(t) : (phase == PHASE_ICE ? es_ice_slope(t) : es_mixed_slope(t));
}
template <typename T> EK_HD T t_from_es(T es) {  // E:109-130 (always the water formula)
    T v = m_log(m_div(es, T(c::C1)));
    return m_div(v * T(c::C4W) - T(c::C3W_T0), v - T(c::C3W));
}

// ------------------------------------------------------------------------------------------
// humidity conversions
// ------------------------------------------------------------------------------------------
template <typename T> EK_HD T q_from_w(T w) { return m_div(w, T(1) + w); }                       // T:77
template <typename T> EK_HD T w_from_q(T q) { return m_div(q, T(1) - q); }                       // T:102
template <typename T> EK_HD T e_from_q(T q, T p) { return m_div(p * q, T(c::eps) + T(c::c_vp) * q); }  // T:130-131
template <typename T> EK_HD T e_from_w(T w, T p) { return m_div(p * w, T(c::eps) + w); }         // T:159
template <typename T> EK_HD T q_from_e(T e, T p, T eps_arg) {                                    // T:193-196
    T v = p + T(c::eps_m1) * e;
    if ((p - e) < eps_arg) v = m_nan<T>();
    return m_div(T(c::eps) * e, v);
}
template <typename T> EK_HD T w_from_e(T e, T p, T eps_arg) {  // T:230-232
    T v = p - e;
    if (v < eps_arg) v = m_nan<T>();
    return m_div(T(c::eps) * e, v);
}
template <typename T> EK_HD T ws_slope_from(T es, T des, T p, T eps_arg) {  // T:412-415
    T v = p - es;
    if (v < eps_arg) v = m_nan<T>();
    return m_div(T(c::eps) * des * p, sq(v));
}
template <typename T> EK_HD T qs_slope_from(T es, T des, T p, T eps_arg) {  // T:464-467
    T v = sq(p + es * T(c::eps_m1));
    if ((p - es) < eps_arg) v = m_nan<T>();
    return m_div(T(c::eps) * des * p, v);
}
template <typename T> EK_HD T rh_from_td(T t, T td) {  // T:519-521
    return m_div(T(100.0) * es_water(td), es_water(t));
}
template <typename T> EK_HD T rh_from_q(T t, T q, T p) {  // T:554-556
    T svp = es_mixed(t);
    return m_div(T(100.0) * e_from_q(q, p), svp);
}
template <typename T> EK_HD T q_from_rh(T t, T r, T p) {  // T:662-663
    T e = m_div(r * es_mixed(t), T(100.0));
    return q_from_e(e, p, T(1e-4));
}
template <typename T> EK_HD T td_from_rh(T t, T r) {  // T:698-699
    return t_from_es(m_div(es_water(t) * r, T(100.0)));
}
template <typename T> EK_HD T td_from_q(T q, T p) { return t_from_es(e_from_q(q, p)); }  // T:735

// ------------------------------------------------------------------------------------------
// dry adiabats, virtual temperature, lcl
// ------------------------------------------------------------------------------------------
template <typename T> EK_HD T tv_factor(T q) { return T(1.0) + T(c::c1_tv) * q; }                 // T:763-764
template <typename T> EK_HD T theta(T t, T p) { return t * m_pow(m_div(T(c::p0), p), T(c::kappa)); }  // T:829
template <typename T> EK_HD T t_from_theta(T th, T p) { return th * m_pow(m_div(p, T(c::p0)), T(c::kappa)); }  // T:858
template <typename T> EK_HD T p_on_dry_adiabat(T t, T t_def, T p_def) {  // T:889
    return p_def * m_pow(m_div(t, t_def), T(c::lambda));
}
template <typename T> EK_HD T t_on_dry_adiabat(T p, T t_def, T p_def) {  // T:920
    return t_def * m_pow(m_div(p, p_def), T(c::kappa));
}
template <typename T> EK_HD T lcl_t_davies(T t, T td) {  // T:961
    return td - (T(0.212) + T(1.571e-3) * (td - T(c::T0)) - T(4.36e-4) * (t - T(c::T0))) * (t - td);
}
template <typename T> EK_HD T lcl_t_bolton(T t, T td) {  // T:966
    return T(56.0) + m_div(T(1), m_div(T(1), td - T(56)) + m_div(m_log(m_div(t, td)), T(800)));
}
template <typename T> EK_HD T lcl_t(T t, T td, int method) {
    return method == LCL_BOLTON ? lcl_t_bolton(t, td) : lcl_t_davies(t, td);
}

// numpy.polynomial.polynomial.polyval (ascending coefficients): c0 = c[-1] + x*0; c0 = c[-i] + c0*x
template <typename T, int N> EK_HD T polyval_asc(T x, const double (&cf)[N]) {
    T c0 = T(cf[N - 1]) + x * T(0);
#pragma unroll
    for (int i = N - 2; i >= 0; --i) c0 = T(cf[i]) + c0 * x;
    return c0;
}

// ------------------------------------------------------------------------------------------
// equivalent potential temperature: the three formulations (T:1162-1323)
// ------------------------------------------------------------------------------------------
namespace k {
constexpr double b35_K0 = 2675.0, b35_K3 = 0.28;                            // T:1202-1203
constexpr double b39_K0 = 3036.0, b39_K1 = 1.78, b39_K2 = 0.448, b39_K4 = 0.28;  // T:1263-1266
}  // namespace k

// ept from (t, td, p) and optionally q.  When has_q is false q is derived from td exactly as each
// formulation does (T:1173-1174, T:1208-1211, T:1271-1274).
template <int M, typename T> EK_HD T ept_point(T t, T td, T q, bool has_q, T p) {
    if (M == EPT_IFS) {  // T:1169-1175
        T th = theta(t, p);
        T t_lcl = lcl_t_davies(t, td);
        if (!has_q) q = q_from_e(es_water(td), p, T(1e-4));
        return th * m_exp(m_div(T(c::K0_ifs) * q, t_lcl));
    }
    T t_lcl = lcl_t_bolton(t, td);
    T w = has_q ? w_from_q(q) : w_from_e(es_water(td), p, T(1e-4));
    if (M == EPT_BOLTON35) {  // T:1205-1213
        T th = t * m_pow(m_div(T(c::p0), p), T(c::kappa) * (T(1) - T(k::b35_K3) * w));
        return th * m_exp(m_div(T(k::b35_K0) * w, t_lcl));
    }
    // T:1268-1278
    T e = e_from_w(w, p);
    T th = theta(t, p - e) * m_pow(m_div(t, t_lcl), T(k::b39_K4) * w);
    return th * m_exp((m_div(T(k::b39_K0), t_lcl) - T(k::b39_K1)) * w * (T(1.0) + T(k::b39_K2) * w));
}

// ept when only q is known: td = dewpoint_from_specific_humidity(q, p) first (T:1036-1037)
template <int M, typename T> EK_HD T ept_from_q_point(T t, T q, T p) {
    return ept_point<M>(t, td_from_q(q, p), q, true, p);
}

// f(t) = ept*exp(G_sat(t,p,scale=-1)) - th_sat(t,p): the function whose sign drives the bisection
// (T:1075).  pk = pow(p0/p, kappa) and lp = p0/p are loop invariants hoisted by the caller.
template <int M, typename T> EK_HD T bisect_residual(T ept, T t, T p, T p0_over_p, T pk) {
    T es = es_mixed(t);
    if (M == EPT_IFS) {  // T:1177-1182
        T qs = q_from_e(es, p, T(1e-4));
        T g = m_div(T(-1.0 * c::K0_ifs) * qs, t);
        return ept * m_exp(g) - t * pk;
    }
    if (M == EPT_BOLTON35) {  // T:1215-1224
        T ws = w_from_e(es, p, T(1e-4));
        T g = m_div(T(-1.0 * k::b35_K0) * ws, t);
        T th = t * m_pow(p0_over_p, T(c::kappa) * (T(1) - T(k::b35_K3) * ws));
        return ept * m_exp(g) - th;
    }
    // bolton39, T:1280-1295: es masked first, then cached for th_sat
    if ((p - es) < T(1e-4)) es = m_nan<T>();
    T ws = w_from_e(es, p, T(1e-4));
    T g = (m_div(T(-1.0 * k::b39_K0), t) - T(-1.0 * k::b39_K1)) * ws * (T(1.0) + T(k::b39_K2) * ws);
    return ept * m_exp(g) - theta(t, p - es);
}

// T:1055-1079: 12 fixed halvings, iterate kept in a register
template <int M, typename T> EK_HD T t_on_ma_bisect(T ept, T p) {
    T t = T(c::T0 - 20);
    T dt = T(120.0);
    const T lp = m_div(T(c::p0), p);
    const T pk = m_pow(lp, T(c::kappa));
#pragma unroll 1
    for (int i = 0; i < 12; ++i) {
        dt = dt / T(2.0);
        t += m_sign(bisect_residual<M>(ept, t, p, lp, pk)) * dt;
    }
    return t;
}

// saturation ept, T:1042-1045: th_sat * exp(G_sat)
template <int M, typename T> EK_HD T sat_ept_point(T t, T p) {
    T es = es_mixed(t);
    if (M == EPT_IFS) {
        T qs = q_from_e(es, p, T(1e-4));
        return theta(t, p) * m_exp(m_div(T(1.0 * c::K0_ifs) * qs, t));
    }
    if (M == EPT_BOLTON35) {
        T ws = w_from_e(es, p, T(1e-4));
        T th = t * m_pow(m_div(T(c::p0), p), T(c::kappa) * (T(1) - T(k::b35_K3) * ws));
        return th * m_exp(m_div(T(1.0 * k::b35_K0) * ws, t));
    }
    if ((p - es) < T(1e-4)) es = m_nan<T>();
    T th = theta(t, p - es);
    T ws = w_from_e(es, p, T(1e-4));
    T g = (m_div(T(1.0 * k::b39_K0), t) - T(1.0 * k::b39_K1)) * ws * (T(1.0) + T(k::b39_K2) * ws);
    return th * m_exp(g);
}

// T:1047-1053 ("direct"): rational fit in x = ept/273.16
template <typename T> EK_HD T wbpt_direct(T ept) {
    const double a[5] = {7.101574, -20.68208, 16.11182, 2.574631, -5.205688};
    const double b[5] = {1.0, -3.552497, 3.781782, -0.6899655, -0.5929340};
    T x = m_div(ept, T(c::t0_dj));
    return ept - m_exp(m_div(polyval_asc(x, a), polyval_asc(x, b)));
}

// T:1081-1159 ("newton"): Davies-Jones first guess by regime + exactly one Newton step
template <int M, typename T> EK_HD T t_on_ma_newton(T ept, T p) {
    const double k1c[3] = {-53.737, 137.81, -38.5};
    const double k2c[3] = {-0.384, 56.831, -4.392};
    const T t0 = T(c::t0_dj);
    const T A = T(2675);
    T tw = ept;
    T pp = m_pow(m_div(p, T(c::p0)), T(c::kappa));
    T te = ept * pp;
    T c_te = m_pow(m_div(t0, te), T(c::lambda));
    T D = m_div(T(1.0), T(0.1859e-5) * p + T(0.6512));
    // the four masks are applied in the reference's order; a later one overwrites an earlier one
    if (c_te > D) {  // T:1114-1119
        T es, d_es;
        es_mixed_both(te, es, d_es);
        T ws = w_from_e(es, p, T(1e-4));
        tw = te - t0 - m_div(A * ws, T(1) + m_div(A * ws * d_es, es));
    }
    T k1 = polyval_asc(pp, k1c), k2 = polyval_asc(pp, k2c);
    if (T(1) <= c_te && c_te <= D) tw = k1 - k2 * c_te;                                  // T:1121-1122
    if (T(0.4) <= c_te && c_te < T(1)) tw = (k1 - T(1.21)) - (k2 - T(1.21)) * c_te;      // T:1124-1125
    if (c_te < T(0.4)) tw = (k1 - T(2.66)) - (k2 - T(1.21)) * c_te + m_div(T(0.58), c_te);  // T:1127-1128
    tw = tw + T(c::T0);  // T:1130

    // one Newton step, T:1132-1149
    {
        T t = tw;
        T c_tw = m_pow(m_div(t0, t), T(c::lambda));
        T es, des;
        es_mixed_both(t, es, des);  // ths.es is cached UNMASKED (T:1135)
        T f, d_lnf;
        if (M == EPT_IFS) {
            T qs = q_from_e(es, p, T(1e-4));
            T g = m_div(T(-c::lambda * c::K0_ifs) * qs, t);                               // T:1180-1182
            f = c_tw * m_exp(g);                                                          // T:1192-1194
            T dqs = qs_slope_from(es, des, p, T(1e-4));
            T dG = m_div(T(-c::K0_ifs) * qs, sq(t)) + m_div(T(c::K0_ifs) * dqs, t);        // T:1184-1190
            d_lnf = T(-c::lambda) * (m_div(T(1), t) + dG);                                // T:1196-1197
        } else if (M == EPT_BOLTON35) {
            T ws = w_from_e(es, p, T(1e-4));
            T g = m_div(T(-c::lambda * k::b35_K0) * ws, t);                               // T:1221-1224
            f = c_tw * m_pow(m_div(p, T(c::p0)), T(k::b35_K3) * ws) * m_exp(g);           // T:1233-1242
            T dws = ws_slope_from(es, des, p, T(1e-4));
            T dG = m_div(T(-k::b35_K0) * ws, sq(t)) + m_div(T(k::b35_K0) * dws, t);        // T:1226-1231
            // the es slope (not the ws slope) multiplies K3*log(p/p0): reference behaviour, kept (T:1246-1250)
            d_lnf = T(-c::lambda) * (m_div(T(1), t) + T(k::b35_K3) * m_log(m_div(p, T(c::p0))) * des + dG);
        } else {
            T ws = w_from_e(es, p, T(1e-4));
            T g = (m_div(T(-c::lambda * k::b39_K0), t) - T(-c::lambda * k::b39_K1)) * ws * (T(1.0) + T(k::b39_K2) * ws);  // T:1287-1295
            f = c_tw * (T(1) - m_div(es, p)) * m_exp(g);                                  // T:1304-1309
            T dws = ws_slope_from(es, des, p, T(1e-4));
            T dG = m_div(T(-k::b39_K0) * (ws + T(k::b39_K2) * sq(ws)), sq(t)) +
                   (m_div(T(k::b39_K0), t) - T(k::b39_K1)) * (T(1) + T(2 * k::b39_K2) * ws) * dws;  // T:1297-1302
            d_lnf = T(-c::lambda) * (m_div(T(1), t) + m_div(T(c::kappa) * des, p - es) + dG);       // T:1311-1316
        }
        tw = tw - m_div(f - c_te, f * d_lnf);  // T:1149
    }
    if (tw <= T(0)) tw = m_nan<T>();  // T:1155
    return tw;
}

template <int M, int TM, typename T> EK_HD T t_on_ma(T ept, T p) {
    if (TM == TM_BISECT) return t_on_ma_bisect<M>(ept, p);
    return t_on_ma_newton<M>(ept, p);
}

}  // namespace ek
