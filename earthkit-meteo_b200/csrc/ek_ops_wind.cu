// ek_ops_wind.cu -- entry points of the elementwise wind functions (reference wind/array/wind.py).
#include "ek_launch.cuh"

using namespace ek;

template <typename T> static int impl_wind_speed(ek_operand u, ek_operand v, void* out, int64_t n, void* stream) {
    ek_operand ins[2] = {u, v};
    void* outs[1] = {out};
    return launch<EK_OPS(OpWindSpeed), T>("wind_speed", ins, outs, n, Params{}, stream);
}
EK_API(wind_speed, (ek_operand u, ek_operand v, void* out, int64_t n, void* stream), (u, v, out, n, stream))

static bool valid_convention(int c) { return c == 0 || c == 1; }

template <typename T> static int impl_wind_direction(ek_operand u, ek_operand v, int convention, int to_positive, void* out, int64_t n, void* stream) {
    if (!valid_convention(convention)) return set_error(EK_ERR_ENUM, "wind_direction: invalid convention id %d", convention);
    ek_operand ins[2] = {u, v};
    void* outs[1] = {out};
    Params P;
    P.opt0 = convention;
    P.opt1 = to_positive ? 1 : 0;
    return launch<EK_OPS(OpWindDirection), T>("wind_direction", ins, outs, n, P, stream);
}
EK_API(wind_direction, (ek_operand u, ek_operand v, int convention, int to_positive, void* out, int64_t n, void* stream),
       (u, v, convention, to_positive, out, n, stream))

template <typename T> static int impl_wind_xy_to_polar(ek_operand x, ek_operand y, int convention, void* speed, void* dir, int64_t n, void* stream) {
    if (!valid_convention(convention)) return set_error(EK_ERR_ENUM, "wind_xy_to_polar: invalid convention id %d", convention);
    if (!speed || !dir) return set_error(EK_ERR_ARG, "wind_xy_to_polar: both output buffers are required");
    ek_operand ins[2] = {x, y};
    void* outs[2] = {speed, dir};
    Params P;
    P.opt0 = convention;
    return launch<EK_OPS(OpXyToPolar), T>("wind_xy_to_polar", ins, outs, n, P, stream);
}
EK_API(wind_xy_to_polar, (ek_operand x, ek_operand y, int convention, void* speed, void* dir, int64_t n, void* stream),
       (x, y, convention, speed, dir, n, stream))

template <typename T> static int impl_wind_polar_to_xy(ek_operand mag, ek_operand dir, int convention, void* x, void* y, int64_t n, void* stream) {
    if (!valid_convention(convention)) return set_error(EK_ERR_ENUM, "wind_polar_to_xy: invalid convention id %d", convention);
    if (!x || !y) return set_error(EK_ERR_ARG, "wind_polar_to_xy: both output buffers are required");
    ek_operand ins[2] = {mag, dir};
    void* outs[2] = {x, y};
    Params P;
    P.opt0 = convention;
    return launch<EK_OPS(OpPolarToXy), T>("wind_polar_to_xy", ins, outs, n, P, stream);
}
EK_API(wind_polar_to_xy, (ek_operand mag, ek_operand dir, int convention, void* x, void* y, int64_t n, void* stream),
       (mag, dir, convention, x, y, n, stream))

template <typename T> static int impl_w_from_omega(ek_operand omega, ek_operand t, ek_operand p, void* out, int64_t n, void* stream) {
    ek_operand ins[3] = {omega, t, p};
    void* outs[1] = {out};
    return launch<EK_OPS(OpWFromOmega), T>("w_from_omega", ins, outs, n, Params{}, stream);
}
EK_API(w_from_omega, (ek_operand omega, ek_operand t, ek_operand p, void* out, int64_t n, void* stream), (omega, t, p, out, n, stream))

template <typename T> static int impl_coriolis(ek_operand lat, void* out, int64_t n, void* stream) {
    ek_operand ins[1] = {lat};
    void* outs[1] = {out};
    return launch<EK_OPS(OpCoriolis), T>("coriolis", ins, outs, n, Params{}, stream);
}
EK_API(coriolis, (ek_operand lat, void* out, int64_t n, void* stream), (lat, out, n, stream))
