// ek_launch.cuh -- host side of a launch: argument checks, grid sizing, error reporting.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ek_thermo.h"
#include "ek_thermo_kernels.cuh"

#define EK_EXPORT __attribute__((visibility("default")))

namespace ek {

// defined in ek_api.cu
int set_error(int code, const char* fmt, ...);
int sm_count_current_device();
extern std::atomic<int> g_ctas_per_sm;
extern std::atomic<uint64_t> g_launches;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Launch Op over n points.  ins[k].ptr == NULL means broadcast scalar ins[k].value; outs[o] == NULL means
// "output o not wanted" (P.out_mask must agree).
template <class Op, class OpE, typename T>
int launch(const char* what, const ek_operand* ins, void* const* outs, int64_t n, Params P, void* stream) {
    if (n < 0) return set_error(EK_ERR_ARG, "%s: n=%lld must be >= 0", what, (long long)n);
    InArgs<Op::NIN> in;
    OutArgs<Op::NOUT> out;
    int vec_ok = 1;
    for (int k = 0; k < Op::NIN; ++k) {
        in.p[k] = ins[k].ptr;
        in.s[k] = ins[k].value;
        if (ins[k].ptr) {
            if (!aligned16(ins[k].ptr)) vec_ok = 0;
            if (reinterpret_cast<uintptr_t>(ins[k].ptr) % sizeof(T))
                return set_error(EK_ERR_ARG, "%s: input %d is not aligned to its element size", what, k);
        }
    }
    uint32_t mask = 0;
    for (int o = 0; o < Op::NOUT; ++o) {
        out.p[o] = outs[o];
        if (outs[o]) {
            mask |= 1u << o;
            if (!aligned16(outs[o])) vec_ok = 0;
            if (reinterpret_cast<uintptr_t>(outs[o]) % sizeof(T))
                return set_error(EK_ERR_ARG, "%s: output %d is not aligned to its element size", what, o);
        }
    }
    if (mask == 0) return set_error(EK_ERR_ARG, "%s: no output buffer given", what);
    P.out_mask = mask;
    if (n == 0) return EK_OK;

    constexpr int threads = kThreads;
    const int64_t tile = (int64_t)threads * Vec16<T>::N * EK_UNROLL;
    const int64_t ntiles = n / tile;
    const int64_t tail_blocks = ((n - ntiles * tile) + threads - 1) / threads;
    const int sms = sm_count_current_device();
    if (sms <= 0) return set_error(EK_ERR_ARG, "%s: no CUDA device is current", what);
    const int64_t cap = (int64_t)sms * g_ctas_per_sm.load(std::memory_order_relaxed);
    int64_t blocks = ntiles > tail_blocks ? ntiles : tail_blocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;

    ew_kernel<Op, OpE, T, EK_UNROLL><<<(unsigned)blocks, threads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(in, out, n, P, vec_ok);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return set_error((int)err, "%s: kernel launch failed: %s", what, cudaGetErrorString(err));
    return EK_OK;
}

inline bool valid_phase(int v) { return v >= 0 && v <= 2; }

// defines ek_thermo_<NAME>_f64 / _f32 forwarding to the template function impl_<NAME><T>
#define EK_API(NAME, PARAMS, ARGS)                                                    \
    extern "C" EK_EXPORT int ek_thermo_##NAME##_f64 PARAMS { return impl_##NAME<double> ARGS; } \
    extern "C" EK_EXPORT int ek_thermo_##NAME##_f32 PARAMS { return impl_##NAME<float> ARGS; }

}  // namespace ek
