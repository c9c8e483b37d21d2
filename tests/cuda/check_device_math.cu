// Accuracy of the branch-free reciprocal used by the generated model code and the integrators
// (sysbio_modeling_b200/csrc/sbm_device_math.h) against IEEE division, over random normal doubles
// spanning many binades.  Prints: max relative error in ulps for 1, 2 and 3 Newton steps, and of sbm_rcp.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include "../../sysbio_modeling_b200/csrc/sbm_device_math.h"

__device__ __forceinline__ double rcp_n(double x, int n)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    for (int k = 0; k < n; ++k) { double e = fma(-x, r, 1.0); r = fma(r, e, r); }
    return r;
}

__global__ void k(const double* x, double* err, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i], ref = 1.0 / v;
    const double ulp = fabs(ref) * 2.220446049250313e-16;
    err[0 * n + i] = fabs(rcp_n(v, 0) - ref) / fabs(ref);
    err[1 * n + i] = fabs(rcp_n(v, 1) - ref) / ulp;
    err[2 * n + i] = fabs(rcp_n(v, 2) - ref) / ulp;
    err[3 * n + i] = fabs(rcp_n(v, 3) - ref) / ulp;
    err[4 * n + i] = fabs(sbm_rcp(v) - ref) / ulp;
}

int main()
{
    const int n = 1 << 22;
    double* hx = (double*)malloc(n * sizeof(double));
    srand(12345);
    for (int i = 0; i < n; ++i) {
        double m = 1.0 + (double)rand() / RAND_MAX + (double)rand() / RAND_MAX / RAND_MAX;
        int e = rand() % 600 - 300;
        hx[i] = ldexp(m, e) * ((rand() & 1) ? 1.0 : -1.0);
    }
    double *dx, *de;
    cudaMalloc(&dx, n * sizeof(double));
    cudaMalloc(&de, 5 * n * sizeof(double));
    cudaMemcpy(dx, hx, n * sizeof(double), cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(dx, de, n);
    double* he = (double*)malloc(5 * n * sizeof(double));
    if (cudaMemcpy(he, de, 5 * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
    double mx[5] = {0, 0, 0, 0, 0};
    for (int j = 0; j < 5; ++j)
        for (int i = 0; i < n; ++i) if (he[j * n + i] > mx[j]) mx[j] = he[j * n + i];
    printf("seed_relerr %.3e newton1_ulp %.3e newton2_ulp %.3e newton3_ulp %.3e sbm_rcp_ulp %.3e\n", mx[0], mx[1], mx[2], mx[3], mx[4]);
    return 0;
}
