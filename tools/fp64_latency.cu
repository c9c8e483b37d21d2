// Dependent-issue latencies that bound the single model warp of sbm_dp5_ct (one warp per SM, clock64 around unrolled chains):
// DFMA chain (1, 2, 4 independent chains), DMUL+DADD, MUFU.RCP64H seed + Newton, shared-memory load -> DFMA, and the same
// DFMA chain while 2 other warps of the scheduler stream independent DFMAs.  nvcc -arch=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k(double* out, long long* clk, double a, double b)
{
    __shared__ double sm[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sm[threadIdx.x & 63] = a;
    __syncthreads();
    double x0 = a + lane, x1 = a - lane, x2 = b + lane, x3 = b - lane;
    long long t0, t1;
    if (warp == 0) {
        t0 = clock64();
#pragma unroll 64
        for (int i = 0; i < N; ++i) x0 = fma(x0, a, b);
        t1 = clock64(); if (lane == 0) clk[0] = t1 - t0;
        t0 = clock64();
#pragma unroll 32
        for (int i = 0; i < N; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); }
        t1 = clock64(); if (lane == 0) clk[1] = t1 - t0;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < N; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
        t1 = clock64(); if (lane == 0) clk[2] = t1 - t0;
        t0 = clock64();
#pragma unroll 64
        for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x0)); x0 = r; }
        t1 = clock64(); if (lane == 0) clk[3] = t1 - t0;
        t0 = clock64();
#pragma unroll 64
        for (int i = 0; i < N; ++i) x0 = fma(x0, sm[(__double2loint(x0) & 31)], b);     // LDS (address depends on x0) -> DFMA
        t1 = clock64(); if (lane == 0) clk[4] = t1 - t0;
        t0 = clock64();
#pragma unroll 64
        for (int i = 0; i < N; ++i) x0 = x0 * a + x1;                                      // DFMA again (contended pass below)
        t1 = clock64(); if (lane == 0) clk[5] = t1 - t0;
    } else if (warp == 4 || warp == 8) {
        // same scheduler as warp 0 (warp % 4): stream independent DFMAs while warp 0 runs all of the above
        for (int i = 0; i < 40 * N; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
    }
    out[threadIdx.x] = x0 + x1 + x2 + x3;
}
int main()
{
    double* o; long long* c;
    cudaMalloc(&o, 512 * 8); cudaMalloc(&c, 64);
    for (int nw = 1; nw <= 12; nw += 11) {
        k<<<1, 32 * nw>>>(o, c, 0.999999, 1e-7);
        long long h[6];
        if (cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error\n"); return 1; }
        printf("warps %2d: clk per step: DFMA chain %.1f | 2 chains %.1f | 4 chains %.1f | MUFU.RCP64H chain %.1f | LDS->DFMA %.1f | DFMA chain (tail) %.1f\n",
               nw, h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N, h[4] / (double)N, h[5] / (double)N);
    }
    return 0;
}
