// probe_mufu.cu -- measures the accuracy of the fp64 reciprocal seed (MUFU.RCP64H via rcp.approx.ftz.f64)
// and of 1/2 Newton refinements, to size ek_thermo_lean.cuh's division.  Build: nvcc -arch=sm_100a, run on a B200.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__global__ void probe(const double* b, double* e0, double* e1, double* e2, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = b[i], r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    double e = fma(-x, r0, 1.0);
    double r1 = fma(fma(e, e, e), r0, r0);  // cubic step
    double f = fma(-x, r1, 1.0);
    double r2 = fma(r1, f, r1);             // + quadratic step
    double t = 1.0 / x;
    e0[i] = fabs(r0 - t) / t;
    e1[i] = fabs(r1 - t) / t;
    e2[i] = fabs(r2 - t) / t;
}

int main() {
    const int n = 1 << 22;
    double *b, *e0, *e1, *e2;
    cudaMallocManaged(&b, n * 8); cudaMallocManaged(&e0, n * 8); cudaMallocManaged(&e1, n * 8); cudaMallocManaged(&e2, n * 8);
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        b[i] = (1.0 + (double)(s >> 11) / 9007199254740992.0) * (i % 2 ? 3.7e4 : 1.3e-2);
    }
    probe<<<(n + 255) / 256, 256>>>(b, e0, e1, e2, n);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(err)); return 1; }
    double m0 = 0, m1 = 0, m2 = 0;
    for (int i = 0; i < n; ++i) { m0 = fmax(m0, e0[i]); m1 = fmax(m1, e1[i]); m2 = fmax(m2, e2[i]); }
    printf("rcp.approx.ftz.f64 seed: max rel err %.3e (2^%.1f)\n", m0, log2(m0));
    printf("after cubic step       : max rel err %.3e (%.2f ulp)\n", m1, m1 / 1.11e-16);
    printf("after +quadratic step  : max rel err %.3e (%.2f ulp)\n", m2, m2 / 1.11e-16);
    return 0;
}
