// Accuracy of the branch-free reciprocal used by the generated model code and the integrators
// (sysbio_modeling_b200/csrc/sbm_device_math.h) against IEEE division, over random normal doubles
// spanning many binades.  Prints: max relative error in ulps for 1, 2 and 3 Newton steps, and of sbm_rcp; then the
// branch-free sbm_log / sbm_exp against the device's libm (itself <= 1 ulp), and their special cases.
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

__global__ void k2(const double* x, double* err, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // log over |x| (many binades), exp over arguments in [-700, 700] derived from the same numbers
    const double v = fabs(x[i]);
    const double rl = log(v);
    err[0 * n + i] = fabs(sbm_log(v) - rl) / fmax(fabs(rl) * 2.220446049250313e-16, 4.9e-324);
    const double m = 1.0 + 1e-6 * (v / ldexp(1.0, ilogb(v)) - 1.5);         // near 1: log -> 0
    const double r1 = log(m);
    err[1 * n + i] = fabs(sbm_log(m) - r1) / fmax(fabs(r1) * 2.220446049250313e-16, 4.9e-324);
    const double a = fmod(v / ldexp(1.0, ilogb(v)) * 7919.0, 1400.0) - 700.0;
    const double re = exp(a);
    err[2 * n + i] = fabs(sbm_exp(a) - re) / (re * 2.220446049250313e-16);
    // pow through exp(b log x): relative error against pow()
    const double b = 0.25 + 3.0 * (v / ldexp(1.0, ilogb(v)) - 1.0), xx = 0.01 + 20.0 * (m - 1.0 + 1.5e-6) / 1e-6 * 0.3;
    const double rp = pow(xx, b);
    err[3 * n + i] = fabs(sbm_exp(b * sbm_log(xx)) - rp) / rp;
}

__global__ void k3(double* out)
{
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    out[0] = sbm_log(0.0); out[1] = sbm_log(-1.0); out[2] = sbm_log(inf); out[3] = sbm_log(1e-310);
    out[4] = sbm_log(1.0); out[5] = sbm_exp(0.0); out[6] = sbm_exp(800.0); out[7] = sbm_exp(-800.0);
    out[8] = sbm_exp(1e12); out[9] = sbm_exp(-1e12); out[10] = sbm_exp(-740.0); out[11] = sbm_exp(inf - inf);
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
    k2<<<(n + 255) / 256, 256>>>(dx, de, n);
    if (cudaMemcpy(he, de, 4 * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
    double m2[4] = {0, 0, 0, 0};
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < n; ++i) if (!(he[j * n + i] <= m2[j])) m2[j] = he[j * n + i];
    printf("sbm_log_ulp %.3e sbm_log_near1_ulp %.3e sbm_exp_ulp %.3e pow_relerr %.3e\n", m2[0], m2[1], m2[2], m2[3]);
    k3<<<1, 1>>>(de);
    double sp[12];
    if (cudaMemcpy(sp, de, sizeof(sp), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
    const char* names[12] = {"log_zero", "log_neg", "log_inf", "log_subnormal", "log_one", "exp_zero", "exp_800", "exp_m800",
                             "exp_1e12", "exp_m1e12", "exp_m740", "exp_nan"};
    for (int j = 0; j < 12; ++j) printf("%s %.17g ", names[j], sp[j]);
    printf("\n");
    return 0;
}
