// probe_rw_mix.cu -- HBM bandwidth ceiling of a streaming kernel as a function of its read:write mix.
// The roofline denominator (MEASURED_PEAKS.json) is a 1R:1W copy; the fused thermo suite is 3R:5W.  This probe runs
// the SAME access pattern as ek_thermo_kernels.cuh (256-thread CTAs, 2 x 16-byte vectors per array per thread,
// ld.global.cs / st.global.cs, grid = SMs x 16) with no math, for several (R, W) mixes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_rw_mix probe_rw_mix.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

// cache-hint variants: 0 = ld.global.cs / st.global.cs (what the product uses), 1 = plain ld / st, 2 = ld.global.nc (read-only
// path, L1 no-allocate) / st.global.cs, 3 = ld.global.cs / st.global.wt, 4 = plain ld / st.global.cg
template <int H> __device__ __forceinline__ double2 ldx(const double2* p) {
    if (H == 1 || H == 4) return *p;
    if (H == 2) {
        double2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
        return v;
    }
    return __ldcs(p);
}
template <int H> __device__ __forceinline__ void stx(double2* p, double2 v) {
    if (H == 1) { *p = v; return; }
    if (H == 3) { __stwt(p, v); return; }
    if (H == 4) { __stcg(p, v); return; }
    __stcs(p, v);
}

template <int R, int W, int H = 0>
__global__ void __launch_bounds__(256, 4) mix(const double* const* in, double* const* out, long n) {
    const long tile = 256 * 2 * 2;
    for (long t = blockIdx.x; t < n / tile; t += gridDim.x) {
        const long base = t * tile + threadIdx.x * 2;
        double2 x[R > 0 ? R : 1][2];
#pragma unroll
        for (int k = 0; k < R; ++k)
#pragma unroll
            for (int u = 0; u < 2; ++u) x[k][u] = ldx<H>(reinterpret_cast<const double2*>(in[k] + base + u * 512));
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double2 s = make_double2(1.0 + t, 2.0);
#pragma unroll
            for (int k = 0; k < R; ++k) { s.x += x[k][u].x; s.y += x[k][u].y; }
#pragma unroll
            for (int o = 0; o < W; ++o) stx<H>(reinterpret_cast<double2*>(out[o] + base + u * 512), make_double2(s.x + o, s.y - o));
        }
    }
}

template <int R, int W, int H = 0> void run(double** d_in, double** d_out, const double* const* in, double* const* out, long n, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) mix<R, W, H><<<sms * 16, 256>>>(in, out, n);
    cudaEventRecord(e0);
    const int iters = 10;
    for (int i = 0; i < iters; ++i) mix<R, W, H><<<sms * 16, 256>>>(in, out, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
    printf("R=%d W=%d hint=%d  %8.3f ms  %8.1f GB/s\n", R, W, H, ms, (R + W) * 8.0 * n / ms / 1e6);
}

int main() {
    const long n = 6599680L * 24;  // 158 M doubles per array, 1.27 GB each
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* h[8];
    for (int i = 0; i < 8; ++i) { cudaMalloc(&h[i], n * 8); cudaMemset(h[i], 0, n * 8); }
    double **d_in, **d_out;
    cudaMalloc(&d_in, 8 * sizeof(double*)); cudaMalloc(&d_out, 8 * sizeof(double*));
    cudaMemcpy(d_in, h, 3 * sizeof(double*), cudaMemcpyHostToDevice);
    cudaMemcpy(d_out, h + 3, 5 * sizeof(double*), cudaMemcpyHostToDevice);
    run<1, 1>(d_in, d_out, d_in, d_out, n, sms);
    run<2, 1>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 1>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 2>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 3>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 5>(d_in, d_out, d_in, d_out, n, sms);
    run<2, 5>(d_in, d_out, d_in, d_out, n, sms);
    run<0, 1>(d_in, d_out, d_in, d_out, n, sms);
    run<0, 5>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 5, 1>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 5, 2>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 5, 3>(d_in, d_out, d_in, d_out, n, sms);
    run<3, 5, 4>(d_in, d_out, d_in, d_out, n, sms);
    run<2, 1, 1>(d_in, d_out, d_in, d_out, n, sms);
    run<2, 1, 2>(d_in, d_out, d_in, d_out, n, sms);
    cudaError_t err = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
