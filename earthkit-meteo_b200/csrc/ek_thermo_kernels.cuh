// ek_thermo_kernels.cuh -- the one streaming kernel template behind every entry point.
//
// The path is elementwise with no reuse, so the design is the HBM-streaming one:
//   * 128-bit coalesced loads/stores (double2 / float4) with streaming cache hints (ld.global.cs /
//     st.global.cs: the data is touched once, keep it out of the way in L1/L2);
//   * each thread front-loads EK_UNROLL vectors per input before any math so that enough bytes are in
//     flight per SM to cover HBM latency (needs ~35 KB/SM in flight at 6.5 TB/s);
//   * all intermediates of a point live in registers; outputs are stored as soon as a vector is done;
//   * tiles are handed to CTAs round-robin (grid-stride), grid sized as SMs x CTAs-per-SM;
//   * no shared memory, no TMA, no tensor cores: there is nothing to stage or contract.
// Broadcast scalars arrive by value (never materialised); unaligned views use scalar ld/st inside the
// same kernel (uniform branch), and the sub-tile tail is a scalar grid-stride loop.
#pragma once
#include <cuda_runtime.h>

#include "ek_thermo_ops.cuh"

#ifndef EK_UNROLL
#define EK_UNROLL 2
#endif
#ifndef EK_MAX_THREADS
#define EK_MAX_THREADS 256
#endif

namespace ek {

template <int N> struct InArgs {
    const void* p[N];
    double s[N];
};
template <int N> struct OutArgs {
    void* p[N];
};

template <typename T> struct Vec16;
template <> struct Vec16<double> {
    using type = double2;
    static constexpr int N = 2;
    static __device__ __forceinline__ void load(const double* src, double* dst) {
        double2 v = __ldcs(reinterpret_cast<const double2*>(src));
        dst[0] = v.x;
        dst[1] = v.y;
    }
    static __device__ __forceinline__ void store(double* dst, const double* src) {
        __stcs(reinterpret_cast<double2*>(dst), make_double2(src[0], src[1]));
    }
};
template <> struct Vec16<float> {
    using type = float4;
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* src, float* dst) {
        float4 v = __ldcs(reinterpret_cast<const float4*>(src));
        dst[0] = v.x;
        dst[1] = v.y;
        dst[2] = v.z;
        dst[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* dst, const float* src) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(src[0], src[1], src[2], src[3]));
    }
};

template <class Op, typename T, int UNROLL>
__global__ void __launch_bounds__(EK_MAX_THREADS)
    ew_kernel(const InArgs<Op::NIN> in, const OutArgs<Op::NOUT> out, const int64_t n, const Params P, const int vec_ok) {
    constexpr int NIN = Op::NIN;
    constexpr int NOUT = Op::NOUT;
    constexpr int VEC = Vec16<T>::N;
    const int64_t vec_stride = (int64_t)blockDim.x * VEC;  // elements between a thread's successive vectors
    const int64_t tile_elems = vec_stride * UNROLL;
    const int64_t ntiles = n / tile_elems;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * tile_elems + (int64_t)threadIdx.x * VEC;
        T x[NIN][UNROLL][VEC];
#pragma unroll
        for (int k = 0; k < NIN; ++k) {
            if (in.p[k] != nullptr) {
                const T* src = static_cast<const T*>(in.p[k]) + base;
                if (vec_ok) {
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u) Vec16<T>::load(src + u * vec_stride, x[k][u]);
                } else {
#pragma unroll
                    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) x[k][u][v] = __ldcs(src + u * vec_stride + v);
                }
            } else {
                const T s = static_cast<T>(in.s[k]);
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[k][u][v] = s;
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            T y[NOUT][VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T a[NIN], r[NOUT];
#pragma unroll
                for (int k = 0; k < NIN; ++k) a[k] = x[k][u][v];
#pragma unroll
                for (int o = 0; o < NOUT; ++o) r[o] = T(0);
                Op::template apply<T>(a, r, P);
#pragma unroll
                for (int o = 0; o < NOUT; ++o) y[o][v] = r[o];
            }
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                if (out.p[o] != nullptr) {
                    T* dst = static_cast<T*>(out.p[o]) + base + u * vec_stride;
                    if (vec_ok) {
                        Vec16<T>::store(dst, y[o]);
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) __stcs(dst + v, y[o][v]);
                    }
                }
            }
        }
    }

    // tail: fewer than one tile of points, one point per thread, grid-stride
    for (int64_t i = ntiles * tile_elems + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        T a[NIN], r[NOUT];
#pragma unroll
        for (int k = 0; k < NIN; ++k) a[k] = (in.p[k] != nullptr) ? __ldcs(static_cast<const T*>(in.p[k]) + i) : static_cast<T>(in.s[k]);
#pragma unroll
        for (int o = 0; o < NOUT; ++o) r[o] = T(0);
        Op::template apply<T>(a, r, P);
#pragma unroll
        for (int o = 0; o < NOUT; ++o)
            if (out.p[o] != nullptr) __stcs(static_cast<T*>(out.p[o]) + i, r[o]);
    }
}

}  // namespace ek
