// tcp_probe.cu - what does tcgen05.cp .32x128b.warpx4 copy where?  (experiment for passing J to the column warps through
// tensor memory: one instruction broadcasts a [32 lanes][16 bytes] block of shared memory to the four lane quadrants)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tcp_probe tools/tcp_probe.cu && tools/tcp_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(unsigned* out, unsigned lbo, unsigned sbo, int ncp, unsigned stride_bytes)
{
    __shared__ __align__(1024) unsigned buf[4096];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ unsigned tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((unsigned)__cvta_generic_to_shared(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"((unsigned)__cvta_generic_to_shared(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tbase;
    // zero the TMEM window first
    for (int c = 0; c < 64; c += 4)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" :: "r"(tb + ((unsigned)(32 * warp) << 16) + c), "r"(0xdeadbeefu) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int k = 0; k < ncp; ++k) {
            const unsigned addr = (unsigned)__cvta_generic_to_shared(buf) + k * stride_bytes;
            unsigned long long desc = (unsigned long long)((addr & 0x3ffff) >> 4) | ((unsigned long long)(lbo >> 4) << 16) |
                                      ((unsigned long long)(sbo >> 4) << 32) | (1ull << 46);
            asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" :: "r"(tb + 4 * k), "l"(desc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"((unsigned)__cvta_generic_to_shared(&mbar)) : "memory");
    }
    {
        unsigned ok = 0;
        while (!ok)
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(&mbar)) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 4 * ncp; c += 4) {
        unsigned r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                     : "r"(tb + ((unsigned)(32 * warp) << 16) + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        unsigned* o = out + ((warp * 32 + lane) * 16 + c);
        o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(tb) : "memory");
}

int main()
{
    unsigned* d;
    cudaMalloc(&d, 128 * 16 * 4);
    unsigned h[128 * 16];
    const unsigned cfg[][4] = {{16, 128, 2, 512}, {128, 16, 2, 512}, {0, 128, 2, 512}, {128, 256, 2, 1024}, {16, 128, 3, 16}};
    for (auto& c : cfg) {
        cudaMemset(d, 0xff, sizeof h);
        probe<<<1, 128>>>(d, c[0], c[1], (int)c[2], c[3]);
        cudaError_t e = cudaDeviceSynchronize();
        printf("== LBO %u SBO %u ncp %u stride %u : %s\n", c[0], c[1], c[2], c[3], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        for (int w = 0; w < 4; w += 3)
            for (int l = 0; l < 32; l += (l < 9 ? 1 : 7)) {
                printf("  warp %d lane %2d:", w, l);
                for (int c2 = 0; c2 < 4 * (int)c[2]; ++c2) printf(" %4u", h[(w * 32 + l) * 16 + c2]);
                printf("\n");
            }
        bool same = true;
        for (int l = 0; l < 32; ++l) for (int c2 = 0; c2 < 4 * (int)c[2]; ++c2) for (int w = 1; w < 4; ++w)
            if (h[(w * 32 + l) * 16 + c2] != h[l * 16 + c2]) same = false;
        printf("  all four quadrants equal: %s\n", same ? "yes" : "NO");
    }
    return 0;
}
