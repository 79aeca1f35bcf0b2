// ek_ops_basic.cu -- entry points of the closed-form single-output functions.
#include "ek_launch.cuh"

using namespace ek;

#define EK_SIMPLE1(NAME, OP)                                                                         \
    template <typename T> static int impl_##NAME(ek_operand a, void* out, int64_t n, void* stream) { \
        ek_operand ins[1] = {a};                                                                     \
        void* outs[1] = {out};                                                                       \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, Params{}, stream);                                 \
    }                                                                                                \
    EK_API(NAME, (ek_operand a, void* out, int64_t n, void* stream), (a, out, n, stream))

#define EK_SIMPLE2(NAME, OP)                                                                                       \
    template <typename T> static int impl_##NAME(ek_operand a, ek_operand b, void* out, int64_t n, void* stream) { \
        ek_operand ins[2] = {a, b};                                                                                \
        void* outs[1] = {out};                                                                                     \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, Params{}, stream);                                               \
    }                                                                                                              \
    EK_API(NAME, (ek_operand a, ek_operand b, void* out, int64_t n, void* stream), (a, b, out, n, stream))

#define EK_SIMPLE3(NAME, OP)                                                                                                     \
    template <typename T> static int impl_##NAME(ek_operand a, ek_operand b, ek_operand c, void* out, int64_t n, void* stream) { \
        ek_operand ins[3] = {a, b, c};                                                                                           \
        void* outs[1] = {out};                                                                                                   \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, Params{}, stream);                                                             \
    }                                                                                                                            \
    EK_API(NAME, (ek_operand a, ek_operand b, ek_operand c, void* out, int64_t n, void* stream), (a, b, c, out, n, stream))

EK_SIMPLE1(celsius_to_kelvin, OpCelsiusToKelvin)
EK_SIMPLE1(kelvin_to_celsius, OpKelvinToCelsius)
EK_SIMPLE1(specific_humidity_from_mixing_ratio, OpQFromW)
EK_SIMPLE1(mixing_ratio_from_specific_humidity, OpWFromQ)
EK_SIMPLE2(vapour_pressure_from_specific_humidity, OpEFromQ)
EK_SIMPLE2(vapour_pressure_from_mixing_ratio, OpEFromW)
EK_SIMPLE1(temperature_from_saturation_vapour_pressure, OpTFromEs)
EK_SIMPLE2(relative_humidity_from_dewpoint, OpRhFromTd)
EK_SIMPLE3(relative_humidity_from_specific_humidity, OpRhFromQ)
EK_SIMPLE2(specific_humidity_from_dewpoint, OpQFromTd)
EK_SIMPLE2(mixing_ratio_from_dewpoint, OpWFromTd)
EK_SIMPLE3(specific_humidity_from_relative_humidity, OpQFromRh)
EK_SIMPLE2(dewpoint_from_relative_humidity, OpTdFromRh)
EK_SIMPLE2(dewpoint_from_specific_humidity, OpTdFromQ)
EK_SIMPLE2(virtual_temperature, OpTv)
EK_SIMPLE3(virtual_potential_temperature, OpThetaV)
EK_SIMPLE2(potential_temperature, OpTheta)
EK_SIMPLE2(temperature_from_potential_temperature, OpTFromTheta)
EK_SIMPLE3(pressure_on_dry_adiabat, OpPOnDryAdiabat)
EK_SIMPLE3(temperature_on_dry_adiabat, OpTOnDryAdiabat)
EK_SIMPLE1(specific_gas_constant, OpGasConstant)

// ---- eps-rule conversions (T:162-232) ---------------------------------------------------------
#define EK_EPS2(NAME, OP)                                                                                                      \
    template <typename T> static int impl_##NAME(ek_operand a, ek_operand b, double eps, void* out, int64_t n, void* stream) { \
        if (!(eps > 0)) return set_error(EK_ERR_EPS, #NAME "(): eps=%g must be > 0", eps);                                     \
        ek_operand ins[2] = {a, b};                                                                                            \
        void* outs[1] = {out};                                                                                                 \
        Params P;                                                                                                              \
        P.eps = eps;                                                                                                           \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, P, stream);                                                                  \
    }                                                                                                                          \
    EK_API(NAME, (ek_operand a, ek_operand b, double eps, void* out, int64_t n, void* stream), (a, b, eps, out, n, stream))

EK_EPS2(specific_humidity_from_vapour_pressure, OpQFromE)
EK_EPS2(mixing_ratio_from_vapour_pressure, OpWFromE)

// ---- phase-dependent saturation functions (T:235-364) -------------------------------------------
#define EK_PHASE1(NAME, OP)                                                                                     \
    template <typename T> static int impl_##NAME(ek_operand a, int phase, void* out, int64_t n, void* stream) { \
        if (!valid_phase(phase)) return set_error(EK_ERR_ENUM, #NAME ": invalid phase id %d", phase);           \
        ek_operand ins[1] = {a};                                                                                \
        void* outs[1] = {out};                                                                                  \
        Params P;                                                                                               \
        P.opt0 = phase;                                                                                         \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, P, stream);                                                   \
    }                                                                                                           \
    EK_API(NAME, (ek_operand a, int phase, void* out, int64_t n, void* stream), (a, phase, out, n, stream))

#define EK_PHASE2(NAME, OP)                                                                                                   \
    template <typename T> static int impl_##NAME(ek_operand a, ek_operand b, int phase, void* out, int64_t n, void* stream) { \
        if (!valid_phase(phase)) return set_error(EK_ERR_ENUM, #NAME ": invalid phase id %d", phase);                         \
        ek_operand ins[2] = {a, b};                                                                                           \
        void* outs[1] = {out};                                                                                                \
        Params P;                                                                                                             \
        P.opt0 = phase;                                                                                                       \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, P, stream);                                                                 \
    }                                                                                                                         \
    EK_API(NAME, (ek_operand a, ek_operand b, int phase, void* out, int64_t n, void* stream), (a, b, phase, out, n, stream))

EK_PHASE1(saturation_vapour_pressure, OpEs)
EK_PHASE1(saturation_vapour_pressure_slope, OpEsSlope)
EK_PHASE2(saturation_mixing_ratio, OpWs)
EK_PHASE2(saturation_specific_humidity, OpQs)

#define EK_SLOPE(NAME, OP)                                                                                                      \
    template <typename T>                                                                                                       \
    static int impl_##NAME(ek_operand t, ek_operand p, ek_operand es, ek_operand des, int has_es, int has_des, int phase,       \
                           double eps, void* out, int64_t n, void* stream) {                                                    \
        if (!(eps > 0)) return set_error(EK_ERR_EPS, #NAME "(): eps=%g must be > 0", eps);                                      \
        if (!valid_phase(phase)) return set_error(EK_ERR_ENUM, #NAME ": invalid phase id %d", phase);                           \
        ek_operand none = {nullptr, 0.0};                                                                                       \
        ek_operand ins[4] = {t, p, has_es ? es : none, has_des ? des : none};                                                   \
        void* outs[1] = {out};                                                                                                  \
        Params P;                                                                                                               \
        P.opt0 = phase;                                                                                                         \
        P.opt1 = (has_es ? 1 : 0) | (has_des ? 2 : 0);                                                                          \
        P.eps = eps;                                                                                                            \
        return launch<EK_OPS(OP), T>(#NAME, ins, outs, n, P, stream);                                                                   \
    }                                                                                                                           \
    EK_API(NAME,                                                                                                                \
           (ek_operand t, ek_operand p, ek_operand es, ek_operand des, int has_es, int has_des, int phase, double eps, void* out, \
            int64_t n, void* stream),                                                                                           \
           (t, p, es, des, has_es, has_des, phase, eps, out, n, stream))

EK_SLOPE(saturation_mixing_ratio_slope, OpWsSlope)
EK_SLOPE(saturation_specific_humidity_slope, OpQsSlope)

// ---- lcl (T:923-1000) -----------------------------------------------------------------------------
template <typename T> static int impl_lcl_temperature(ek_operand t, ek_operand td, int method, void* out, int64_t n, void* stream) {
    if (method != EK_LCL_DAVIES && method != EK_LCL_BOLTON) return set_error(EK_ERR_ENUM, "lcl_temperature: invalid method id %d", method);
    ek_operand ins[2] = {t, td};
    void* outs[1] = {out};
    Params P;
    P.opt0 = method;
    return launch<EK_OPS(OpLclT), T>("lcl_temperature", ins, outs, n, P, stream);
}
EK_API(lcl_temperature, (ek_operand t, ek_operand td, int method, void* out, int64_t n, void* stream), (t, td, method, out, n, stream))

template <typename T>
static int impl_lcl(ek_operand t, ek_operand td, ek_operand p, int method, void* t_lcl, void* p_lcl, int64_t n, void* stream) {
    if (method != EK_LCL_DAVIES && method != EK_LCL_BOLTON) return set_error(EK_ERR_ENUM, "lcl: invalid method id %d", method);
    if (!t_lcl || !p_lcl) return set_error(EK_ERR_ARG, "lcl: both output buffers are required");
    ek_operand ins[3] = {t, td, p};
    void* outs[2] = {t_lcl, p_lcl};
    Params P;
    P.opt0 = method;
    return launch<EK_OPS(OpLcl), T>("lcl", ins, outs, n, P, stream);
}
EK_API(lcl, (ek_operand t, ek_operand td, ek_operand p, int method, void* t_lcl, void* p_lcl, int64_t n, void* stream),
       (t, td, p, method, t_lcl, p_lcl, n, stream))

// ---- height forms of a geopotential ---------------------------------------------------
template <typename T> static int impl_height_from_thickness(ek_operand dphi, ek_operand zs, int mode, void* out, int64_t n, void* stream) {
    if (mode < 0 || mode > 6) return set_error(EK_ERR_ENUM, "height_from_thickness: invalid mode %d", mode);
    ek_operand ins[2] = {dphi, zs};
    void* outs[1] = {out};
    Params P;
    P.opt0 = mode;
    return launch<EK_OPS(OpHeightForm), T>("height_from_thickness", ins, outs, n, P, stream);
}
EK_API(height_from_thickness, (ek_operand dphi, ek_operand zs, int mode, void* out, int64_t n, void* stream), (dphi, zs, mode, out, n, stream))
