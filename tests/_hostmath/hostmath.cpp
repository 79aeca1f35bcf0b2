// hostmath.cpp -- TEST-ONLY host instantiation of the device op functors (ek_thermo_ops.cuh).
//
// Compiled with plain g++ (no CUDA) into tests/_hostmath/libek_hostmath.so and loaded ONLY by
// tests/test_hostmath.py, which compares it with the oracle.  Purpose: the build container has no GPU,
// so this is how the per-point logic that the kernels execute (operation order, option handling,
// NaN rules, output masks) is verified on the CPU before a GPU run.  It is not part of the product:
// the ek_thermo package never loads it and has no CPU execution path.
#include <cstring>
#include <string>

#include "../../earthkit-meteo_b200/csrc/ek_thermo_math.cuh"

using namespace ek;
using namespace ek::exactm;

template <class Op, typename T>
static void run(const void* const* ins, const double* scalars, void* const* outs, int64_t n, const Params& P) {
    for (int64_t i = 0; i < n; ++i) {
        T a[Op::NIN], r[Op::NOUT];
        for (int k = 0; k < Op::NIN; ++k) a[k] = ins[k] ? static_cast<const T*>(ins[k])[i] : static_cast<T>(scalars[k]);
        for (int o = 0; o < Op::NOUT; ++o) r[o] = T(0);
        Op::template apply<T>(a, r, P);
        for (int o = 0; o < Op::NOUT; ++o)
            if (outs[o]) static_cast<T*>(outs[o])[i] = r[o];
    }
}

template <class Op> static int both(int f32, const void* const* ins, const double* sc, void* const* outs, int64_t n, const Params& P) {
    if (f32) run<Op, float>(ins, sc, outs, n, P); else run<Op, double>(ins, sc, outs, n, P);
    return 0;
}

template <int M> static int ept_wb(int tm, int f32, const void* const* ins, const double* sc, void* const* outs, int64_t n, const Params& P) {
    switch (tm) {
        case TM_NONE: return both<OpEptWb<M, TM_NONE>>(f32, ins, sc, outs, n, P);
        case TM_DIRECT: return both<OpEptWb<M, TM_DIRECT>>(f32, ins, sc, outs, n, P);
        case TM_BISECT: return both<OpEptWb<M, TM_BISECT>>(f32, ins, sc, outs, n, P);
        case TM_NEWTON: return both<OpEptWb<M, TM_NEWTON>>(f32, ins, sc, outs, n, P);
    }
    return -2;
}
template <int M> static int t_on_ma_(int tm, int f32, const void* const* ins, const double* sc, void* const* outs, int64_t n, const Params& P) {
    if (tm == TM_BISECT) return both<OpTOnMa<M, TM_BISECT>>(f32, ins, sc, outs, n, P);
    if (tm == TM_NEWTON) return both<OpTOnMa<M, TM_NEWTON>>(f32, ins, sc, outs, n, P);
    return -2;
}

extern "C" int hostmath_run(const char* op, int f32, const void* const* ins, const double* scalars, void* const* outs, int64_t n,
                            int opt0, int opt1, double eps, uint32_t out_mask, int m, int tm) {
    Params P;
    P.opt0 = opt0;
    P.opt1 = opt1;
    P.eps = eps;
    P.out_mask = out_mask;
    const std::string s(op);
#define EK_CASE(NAME, OP) \
    if (s == NAME) return both<OP>(f32, ins, scalars, outs, n, P);
    EK_CASE("celsius_to_kelvin", OpCelsiusToKelvin)
    EK_CASE("kelvin_to_celsius", OpKelvinToCelsius)
    EK_CASE("specific_humidity_from_mixing_ratio", OpQFromW)
    EK_CASE("mixing_ratio_from_specific_humidity", OpWFromQ)
    EK_CASE("vapour_pressure_from_specific_humidity", OpEFromQ)
    EK_CASE("vapour_pressure_from_mixing_ratio", OpEFromW)
    EK_CASE("specific_humidity_from_vapour_pressure", OpQFromE)
    EK_CASE("mixing_ratio_from_vapour_pressure", OpWFromE)
    EK_CASE("saturation_vapour_pressure", OpEs)
    EK_CASE("saturation_vapour_pressure_slope", OpEsSlope)
    EK_CASE("saturation_mixing_ratio", OpWs)
    EK_CASE("saturation_specific_humidity", OpQs)
    EK_CASE("saturation_mixing_ratio_slope", OpWsSlope)
    EK_CASE("saturation_specific_humidity_slope", OpQsSlope)
    EK_CASE("temperature_from_saturation_vapour_pressure", OpTFromEs)
    EK_CASE("relative_humidity_from_dewpoint", OpRhFromTd)
    EK_CASE("relative_humidity_from_specific_humidity", OpRhFromQ)
    EK_CASE("specific_humidity_from_dewpoint", OpQFromTd)
    EK_CASE("mixing_ratio_from_dewpoint", OpWFromTd)
    EK_CASE("specific_humidity_from_relative_humidity", OpQFromRh)
    EK_CASE("dewpoint_from_relative_humidity", OpTdFromRh)
    EK_CASE("dewpoint_from_specific_humidity", OpTdFromQ)
    EK_CASE("virtual_temperature", OpTv)
    EK_CASE("virtual_potential_temperature", OpThetaV)
    EK_CASE("potential_temperature", OpTheta)
    EK_CASE("temperature_from_potential_temperature", OpTFromTheta)
    EK_CASE("pressure_on_dry_adiabat", OpPOnDryAdiabat)
    EK_CASE("temperature_on_dry_adiabat", OpTOnDryAdiabat)
    EK_CASE("lcl_temperature", OpLclT)
    EK_CASE("lcl", OpLcl)
    EK_CASE("specific_gas_constant", OpGasConstant)
    EK_CASE("wind_speed", OpWindSpeed)
    EK_CASE("wind_direction", OpWindDirection)
    EK_CASE("wind_xy_to_polar", OpXyToPolar)
    EK_CASE("wind_polar_to_xy", OpPolarToXy)
    EK_CASE("w_from_omega", OpWFromOmega)
    EK_CASE("coriolis", OpCoriolis)
#undef EK_CASE
    if (s == "suite_tqp" || s == "suite_ttdp") {  // generic (run-time mask) instantiation, per ept formulation
        const bool q = s == "suite_tqp";
        switch (m) {
            case EPT_IFS: return q ? both<OpSuiteTQPm<0, EPT_IFS>>(f32, ins, scalars, outs, n, P) : both<OpSuiteTTdPm<0, EPT_IFS>>(f32, ins, scalars, outs, n, P);
            case EPT_BOLTON35: return q ? both<OpSuiteTQPm<0, EPT_BOLTON35>>(f32, ins, scalars, outs, n, P) : both<OpSuiteTTdPm<0, EPT_BOLTON35>>(f32, ins, scalars, outs, n, P);
            case EPT_BOLTON39: return q ? both<OpSuiteTQPm<0, EPT_BOLTON39>>(f32, ins, scalars, outs, n, P) : both<OpSuiteTTdPm<0, EPT_BOLTON39>>(f32, ins, scalars, outs, n, P);
        }
        return -2;
    }
    // the compile-time-mask instantiations the library ships (0x31F: the seven-output single pass, 0x30D: theta, rh, td|q, ept, wbpt)
    if (s == "suite_tqp_31f") return both<OpSuiteTQPm<0x31F, EPT_IFS>>(f32, ins, scalars, outs, n, P);
    if (s == "suite_ttdp_31f") return both<OpSuiteTTdPm<0x31F, EPT_IFS>>(f32, ins, scalars, outs, n, P);
    if (s == "suite_tqp_30d") return both<OpSuiteTQPm<0x30D, EPT_IFS>>(f32, ins, scalars, outs, n, P);
    if (s == "suite_ttdp_30d") return both<OpSuiteTTdPm<0x30D, EPT_IFS>>(f32, ins, scalars, outs, n, P);
    if (s == "hyb_full" || s == "hyb_delta_alpha") {  // hybrid-level formulas (ek_thermo_formulas.inc), per point
        for (int64_t i = 0; i < n; ++i) {
            auto in = [&](int k) { return f32 ? (double)static_cast<const float*>(ins[k])[i] : static_cast<const double*>(ins[k])[i]; };
            auto put = [&](int o, double v) {
                if (!outs[o]) return;
                if (f32) static_cast<float*>(outs[o])[i] = (float)v; else static_cast<double*>(outs[o])[i] = v;
            };
            if (s == "hyb_full") {  // in: a0, b0, a1, b1, sp
                if (f32) put(0, hyb_full<float>(hyb_half<float>(in(0), in(1), in(4)), hyb_half<float>(in(2), in(3), in(4))));
                else put(0, hyb_full<double>(hyb_half<double>(in(0), in(1), in(4)), hyb_half<double>(in(2), in(3), in(4))));
            } else {  // in: ph0, ph1; opt0 = top-is-toa, eps = alpha_top
                if (f32) { float d, a; hyb_delta_alpha<float>(in(0), in(1), opt0 != 0, (float)eps, d, a); put(0, d); put(1, a); }
                else { double d, a; hyb_delta_alpha<double>(in(0), in(1), opt0 != 0, eps, d, a); put(0, d); put(1, a); }
            }
        }
        return 0;
    }
    if (s == "ept_wet_bulb") {
        if (m == EPT_IFS) return ept_wb<EPT_IFS>(tm, f32, ins, scalars, outs, n, P);
        if (m == EPT_BOLTON35) return ept_wb<EPT_BOLTON35>(tm, f32, ins, scalars, outs, n, P);
        if (m == EPT_BOLTON39) return ept_wb<EPT_BOLTON39>(tm, f32, ins, scalars, outs, n, P);
        return -2;
    }
    if (s == "temperature_on_moist_adiabat") {
        if (m == EPT_IFS) return t_on_ma_<EPT_IFS>(tm, f32, ins, scalars, outs, n, P);
        if (m == EPT_BOLTON35) return t_on_ma_<EPT_BOLTON35>(tm, f32, ins, scalars, outs, n, P);
        if (m == EPT_BOLTON39) return t_on_ma_<EPT_BOLTON39>(tm, f32, ins, scalars, outs, n, P);
        return -2;
    }
    if (s == "saturation_ept") {
        if (m == EPT_IFS) return both<OpSatEpt<EPT_IFS>>(f32, ins, scalars, outs, n, P);
        if (m == EPT_BOLTON35) return both<OpSatEpt<EPT_BOLTON35>>(f32, ins, scalars, outs, n, P);
        if (m == EPT_BOLTON39) return both<OpSatEpt<EPT_BOLTON39>>(f32, ins, scalars, outs, n, P);
        return -2;
    }
    return -1;
}
