// Microbenchmark: tensor-memory load/store throughput per SM with the 32x32b shape (thread t <-> lane t),
// as used by sbm_dp5_ct for thread-private vectors.   nvcc -arch=sm_100a -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cuda_runtime.h>

template <int X>
__device__ __forceinline__ void ld(unsigned ta, unsigned (&r)[X]);
template <> __device__ __forceinline__ void ld<4>(unsigned ta, unsigned (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(ta));
}
template <> __device__ __forceinline__ void ld<16>(unsigned ta, unsigned (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta));
}
__device__ __forceinline__ void st4(unsigned ta, unsigned a, unsigned b, unsigned c, unsigned d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(ta), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <int X, int MODE>   // MODE 0: loads, 1: stores (x4 only), 2: load + wait each (latency-bound)
__global__ void k(unsigned long long* out, int iters)
{
    __shared__ unsigned base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = base_s + ((unsigned)(32 * (warp & 3)) << 16);
    unsigned acc = 0;
    // initialise the columns this warp touches
    for (int c = 0; c < 512; c += 4) st4(base + c, 1u, 2u, 3u, 4u);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned col = (unsigned)(((it * 8 + u) * X) & 511) & ~(unsigned)(X - 1);
            if (MODE == 1) {
                st4(base + col, acc, (unsigned)it, (unsigned)u, 7u);
            } else {
                unsigned r[X];
                ld<X>(base + col, r);
                if (MODE == 2) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("" : "+r"(r[0]));
                acc += r[0];
            }
        }
        if (MODE == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) out[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base_s), "r"(512) : "memory");
}

template <int X, int MODE>
void run(const char* name, int warps)
{
    unsigned long long* d;
    cudaMalloc(&d, 148 * 8);
    const int iters = 4096;
    k<X, MODE><<<148, warps * 32>>>(d, 64);
    cudaDeviceSynchronize();
    k<X, MODE><<<148, warps * 32>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double clk = 0;
    for (int i = 0; i < 148; ++i) clk += (double)h[i];
    clk /= 148;
    const double bytes = (double)iters * 8 * X * 4 * 32 * warps;     // per SM
    printf("%-28s warps %2d  %8.1f B/clk/SM  (%.1f clk per warp-instruction)  %s\n", name, warps, bytes / clk,
           clk / (iters * 8.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main()
{
    for (int w : {1, 2, 4, 8, 12, 16}) run<4, 0>("LDTM.x4 (batched wait)", w);
    for (int w : {1, 4, 8, 16}) run<16, 0>("LDTM.x16 (batched wait)", w);
    for (int w : {1, 4, 12}) run<4, 2>("LDTM.x4 + wait each", w);
    for (int w : {1, 4, 8, 12, 16}) run<4, 1>("STTM.x4", w);
    return 0;
}
