// ek_thermo_ops.cuh -- one functor per kernel body: "read NIN values of one grid point, write NOUT".
//
// The same functors are instantiated (a) inside the sm_100a kernels (ek_thermo_kernels.cuh) and
// (b) by the host-only check harness tests/_hostmath (plain g++), so the per-point logic, option
// handling and output masks that run on the GPU are exactly what the CPU tests verify against
// the oracle.  Reference citations are on the formulas in ek_thermo_math.cuh.
#pragma once
#include "ek_thermo_math.cuh"

namespace ek {

// Per-launch options, passed by value to the kernel (uniform across the grid).
struct Params {
    int opt0 = 0;           // phase | lcl method | humidity kind (0 = dewpoint, 1 = specific humidity)
    int opt1 = 0;           // flags: bit0 "es given", bit1 "es_slope given" | wet-bulb level (0 = at p, 1 = at p0)
    double eps = 1e-4;      // the eps argument of the NaN rule (T:162,199,367,418)
    uint32_t out_mask = 1;  // which outputs are wanted (bit k = output k)
};

#define EK_OP_BEGIN(NAME, NIN_, NOUT_)                                               \
    struct NAME {                                                                    \
        static constexpr int NIN = NIN_;                                             \
        static constexpr int NOUT = NOUT_;                                           \
        template <typename T> static EK_HD void apply(const T* in, T* out, const Params& P) {
#define EK_OP_END \
    }             \
    }             \
    ;

// ---- one-output closed forms -----------------------------
EK_OP_BEGIN(OpCelsiusToKelvin, 1, 1) (void)P; out[0] = in[0] + T(c::T0); EK_OP_END                 // T:35
EK_OP_BEGIN(OpKelvinToCelsius, 1, 1) (void)P; out[0] = in[0] - T(c::T0); EK_OP_END                 // T:52
EK_OP_BEGIN(OpQFromW, 1, 1) (void)P; out[0] = q_from_w(in[0]); EK_OP_END                           // T:77
EK_OP_BEGIN(OpWFromQ, 1, 1) (void)P; out[0] = w_from_q(in[0]); EK_OP_END                           // T:102
EK_OP_BEGIN(OpEFromQ, 2, 1) (void)P; out[0] = e_from_q(in[0], in[1]); EK_OP_END                    // T:131
EK_OP_BEGIN(OpEFromW, 2, 1) (void)P; out[0] = e_from_w(in[0], in[1]); EK_OP_END                    // T:159
EK_OP_BEGIN(OpQFromE, 2, 1) out[0] = q_from_e(in[0], in[1], T(P.eps)); EK_OP_END                   // T:193-196
EK_OP_BEGIN(OpWFromE, 2, 1) out[0] = w_from_e(in[0], in[1], T(P.eps)); EK_OP_END                   // T:230-232
EK_OP_BEGIN(OpEs, 1, 1) out[0] = es_phase(in[0], P.opt0); EK_OP_END                                // T:279
EK_OP_BEGIN(OpEsSlope, 1, 1) out[0] = es_slope_phase(in[0], P.opt0); EK_OP_END                     // T:364
EK_OP_BEGIN(OpWs, 2, 1) out[0] = w_from_e(es_phase(in[0], P.opt0), in[1], T(1e-4)); EK_OP_END      // T:309-310
EK_OP_BEGIN(OpQs, 2, 1) out[0] = q_from_e(es_phase(in[0], P.opt0), in[1], T(1e-4)); EK_OP_END      // T:340-341
// in = t, p, es, es_slope; opt1 bit0/bit1 say whether es / es_slope were supplied (T:407-410, T:459-462)
EK_OP_BEGIN(OpWsSlope, 4, 1)
    T es = (P.opt1 & 1) ? in[2] : es_phase(in[0], P.opt0);
    T des = (P.opt1 & 2) ? in[3] : es_slope_phase(in[0], P.opt0);
    out[0] = ws_slope_from(es, des, in[1], T(P.eps));
EK_OP_END
EK_OP_BEGIN(OpQsSlope, 4, 1)
    T es = (P.opt1 & 1) ? in[2] : es_phase(in[0], P.opt0);
    T des = (P.opt1 & 2) ? in[3] : es_slope_phase(in[0], P.opt0);
    out[0] = qs_slope_from(es, des, in[1], T(P.eps));
EK_OP_END
EK_OP_BEGIN(OpTFromEs, 1, 1) (void)P; out[0] = t_from_es(in[0]); EK_OP_END                         // T:491
EK_OP_BEGIN(OpRhFromTd, 2, 1) (void)P; out[0] = rh_from_td(in[0], in[1]); EK_OP_END                // T:519-521
EK_OP_BEGIN(OpRhFromQ, 3, 1) (void)P; out[0] = rh_from_q(in[0], in[1], in[2]); EK_OP_END           // T:554-556
EK_OP_BEGIN(OpQFromTd, 2, 1) (void)P; out[0] = q_from_e(es_water(in[0]), in[1], T(1e-4)); EK_OP_END  // T:590-591
EK_OP_BEGIN(OpWFromTd, 2, 1) (void)P; out[0] = w_from_e(es_water(in[0]), in[1], T(1e-4)); EK_OP_END  // T:625-626
EK_OP_BEGIN(OpQFromRh, 3, 1) (void)P; out[0] = q_from_rh(in[0], in[1], in[2]); EK_OP_END           // T:662-663
EK_OP_BEGIN(OpTdFromRh, 2, 1) (void)P; out[0] = td_from_rh(in[0], in[1]); EK_OP_END                // T:698-699
EK_OP_BEGIN(OpTdFromQ, 2, 1) (void)P; out[0] = td_from_q(in[0], in[1]); EK_OP_END                  // T:735
EK_OP_BEGIN(OpTv, 2, 1) (void)P; out[0] = in[0] * tv_factor(in[1]); EK_OP_END                      // T:763-764
EK_OP_BEGIN(OpThetaV, 3, 1) (void)P; out[0] = theta(in[0], in[2]) * tv_factor(in[1]); EK_OP_END    // T:797-798
EK_OP_BEGIN(OpTheta, 2, 1) (void)P; out[0] = theta(in[0], in[1]); EK_OP_END                        // T:829
EK_OP_BEGIN(OpTFromTheta, 2, 1) (void)P; out[0] = t_from_theta(in[0], in[1]); EK_OP_END            // T:858
EK_OP_BEGIN(OpPOnDryAdiabat, 3, 1) (void)P; out[0] = p_on_dry_adiabat(in[0], in[1], in[2]); EK_OP_END  // T:889
EK_OP_BEGIN(OpTOnDryAdiabat, 3, 1) (void)P; out[0] = t_on_dry_adiabat(in[0], in[1], in[2]); EK_OP_END  // T:920
EK_OP_BEGIN(OpLclT, 2, 1) out[0] = lcl_t(in[0], in[1], P.opt0); EK_OP_END                          // T:960-966
// lcl returns the pair (t_lcl, p_lcl) (T:998-1000)
EK_OP_BEGIN(OpLcl, 3, 2)
    T tl = lcl_t(in[0], in[1], P.opt0);
    out[0] = tl;
    out[1] = p_on_dry_adiabat(tl, in[0], in[2]);
EK_OP_END
EK_OP_BEGIN(OpGasConstant, 1, 1) (void)P; out[0] = T(c::Rd) + T(c::Rv - c::Rd) * in[0]; EK_OP_END  // T:1706

// ---- ept / wet-bulb family (rows A31-A44) ---------------------------------------------------
// in = t, h, p where h is the dewpoint (opt0 = 0) or the specific humidity (opt0 = 1).
// out[0] = ept, out[1] = temperature on the moist adiabat through ept, taken at p (opt1 = 0, wet-bulb
// temperature T:1548-1549) or at p0 (opt1 = 1, wet-bulb potential temperature T:1630-1634).
template <int M, int TM> struct OpEptWb {
    static constexpr int NIN = 3;
    static constexpr int NOUT = 2;
    template <typename T> static EK_HD void apply(const T* in, T* out, const Params& P) {
        T ept = P.opt0 ? ept_from_q_point<M>(in[0], in[1], in[2]) : ept_point<M>(in[0], in[1], T(0), false, in[2]);
        out[0] = ept;
        if (TM == TM_NONE) return;
        if (TM == TM_DIRECT) {
            out[1] = wbpt_direct(ept);
        } else {
            T p = P.opt1 ? T(c::p0) : in[2];
            out[1] = t_on_ma<M, TM>(ept, p);
        }
    }
};
template <int M, int TM> struct OpTOnMa {  // in = ept, p (T:1472-1509)
    static constexpr int NIN = 2;
    static constexpr int NOUT = 1;
    template <typename T> static EK_HD void apply(const T* in, T* out, const Params&) {
        out[0] = t_on_ma<M, TM>(in[0], in[1]);
    }
};
template <int M> struct OpSatEpt {  // in = t, p (T:1418-1469)
    static constexpr int NIN = 2;
    static constexpr int NOUT = 1;
    template <typename T> static EK_HD void apply(const T* in, T* out, const Params&) {
        out[0] = sat_ept_point<M>(in[0], in[1]);
    }
};

// ---- fused suites: read the state of a point once, write every requested diagnostic ---------
// Output slots (same for both suites so callers can share a mask vocabulary):
//   0 theta   potential_temperature(t, p)                         T:829
//   1 es      saturation_vapour_pressure(t)  [mixed phase]        E:141-166
//   2 rh      relative humidity [%]                               T:556 (from q) / T:521 (from td)
//   3 td | q  dewpoint_from_specific_humidity(q,p) T:735  |  specific_humidity_from_dewpoint(td,p) T:591
//   4 tv      virtual_temperature(t, q)                           T:764
//   5 w       mixing ratio: T:102 (from q)  |  T:626 (from td)
//   6 e       vapour pressure: T:131 (from q)  |  es_water(td) T:590
//   7 thetav  virtual_potential_temperature(t, q, p)              T:798
enum SuiteSlot : int { S_THETA = 0, S_ES = 1, S_RH = 2, S_TDQ = 3, S_TV = 4, S_W = 5, S_E = 6, S_THETAV = 7, S_NSLOTS = 8 };

struct OpSuiteTQP {  // in = t, q, p
    static constexpr int NIN = 3;
    static constexpr int NOUT = S_NSLOTS;
    template <typename T> static EK_HD void apply(const T* in, T* out, const Params& P) {
        const uint32_t m = P.out_mask;
        const T t = in[0], q = in[1], p = in[2];
        T th = T(0), tvf = T(0), e = T(0);
        if (m & ((1u << S_THETA) | (1u << S_THETAV))) th = theta(t, p);
        if (m & ((1u << S_TV) | (1u << S_THETAV))) tvf = tv_factor(q);
        if (m & ((1u << S_RH) | (1u << S_TDQ) | (1u << S_E))) e = e_from_q(q, p);
        if (m & (1u << S_THETA)) out[S_THETA] = th;
        if (m & ((1u << S_ES) | (1u << S_RH))) {
            T es = es_mixed(t);
            out[S_ES] = es;
            if (m & (1u << S_RH)) out[S_RH] = m_div(T(100.0) * e, es);
        }
        if (m & (1u << S_TDQ)) out[S_TDQ] = t_from_es(e);
        if (m & (1u << S_TV)) out[S_TV] = t * tvf;
        if (m & (1u << S_W)) out[S_W] = w_from_q(q);
        if (m & (1u << S_E)) out[S_E] = e;
        if (m & (1u << S_THETAV)) out[S_THETAV] = th * tvf;
    }
};

struct OpSuiteTTdP {  // in = t, td, p
    static constexpr int NIN = 3;
    static constexpr int NOUT = S_NSLOTS;
    template <typename T> static EK_HD void apply(const T* in, T* out, const Params& P) {
        const uint32_t m = P.out_mask;
        const T t = in[0], td = in[1], p = in[2];
        T th = T(0), e = T(0), q = T(0);
        if (m & ((1u << S_THETA) | (1u << S_THETAV))) th = theta(t, p);
        if (m & ((1u << S_RH) | (1u << S_TDQ) | (1u << S_TV) | (1u << S_W) | (1u << S_E) | (1u << S_THETAV))) e = es_water(td);
        if (m & ((1u << S_TDQ) | (1u << S_TV) | (1u << S_THETAV))) q = q_from_e(e, p, T(1e-4));
        if (m & (1u << S_THETA)) out[S_THETA] = th;
        if (m & (1u << S_ES)) out[S_ES] = es_mixed(t);
        if (m & (1u << S_RH)) out[S_RH] = m_div(T(100.0) * e, es_water(t));
        if (m & (1u << S_TDQ)) out[S_TDQ] = q;
        if (m & (1u << S_TV)) out[S_TV] = t * tv_factor(q);
        if (m & (1u << S_W)) out[S_W] = w_from_e(e, p, T(1e-4));
        if (m & (1u << S_E)) out[S_E] = e;
        if (m & (1u << S_THETAV)) out[S_THETAV] = th * tv_factor(q);
    }
};

}  // namespace ek
