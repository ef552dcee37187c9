// Micro-benchmark (development aid, not product): tcgen05.mma execution rate (M=128, K=16, bf16) as a function of N and the
// shared-memory operand layout.  The issue loop is convergent (whole warp, elect.sync inside) with descriptors advanced
// by one 32-bit add, 8x unrolled, so the issuing thread is not the limiter (an earlier version of this file measured
// its own loop overhead: 76 cycles / iteration).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}

template <int NACC>
__global__ void __launch_bounds__(128) rate_kernel(int n_iter, int N, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
                                                   uint32_t b_sbo, uint32_t layout, uint32_t a_step, int commit_every,
                                                   long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
        const uint64_t da0 = make_desc(a0, a_lbo, a_sbo, layout), db0 = make_desc(b0, b_lbo, b_sbo, layout);
        const uint32_t astep16 = a_step >> 4;
        uint32_t parity = 0;
        const long long t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    mma(tmem + (uint32_t)((u % NACC) * N), da0 + (uint64_t)(u * astep16), db0 + (uint64_t)((u & 3) * 2), idesc,
                        (it | (u >= NACC)) ? 1u : 0u);
                if (commit_every && (it % commit_every) == commit_every - 1 && it + 1 < n_iter)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            if (commit_every && (it % commit_every) == commit_every - 1 && it + 1 < n_iter) {
                // emulate a pipeline that only runs `commit_every` iterations ahead of the completion
                uint32_t ok = 0;
                const long long tw = clock64();
                while (!ok && clock64() - tw < 400000000ll)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
                parity ^= 1;
            }
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        const long long tw = clock64();
        while (!ok && clock64() - tw < 400000000ll)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    long long *d_out, h_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { const char *name; int N; uint32_t a_lbo, a_sbo, b_lbo, b_sbo, layout, a_step; };
    const Cfg cfgs[] = {
        {"none  N=64  conv halo A (lbo 9856 sbo 544)", 64, 9856, 544, 1024, 128, 0, 16},
        {"none  N=128 conv halo A (lbo 9856 sbo 544)", 128, 9856, 544, 2048, 128, 0, 16},
        {"none  N=256 conv halo A (lbo 9856 sbo 544)", 256, 9856, 544, 4096, 128, 0, 16},
        {"none  N=32  dense", 32, 2048, 128, 512, 128, 0, 0},
        {"none  N=64  dense", 64, 2048, 128, 1024, 128, 0, 0},
        {"none  N=96  dense", 96, 2048, 128, 1536, 128, 0, 0},
        {"none  N=128 dense", 128, 2048, 128, 2048, 128, 0, 0},
        {"none  N=160 dense", 160, 2048, 128, 2560, 128, 0, 0},
        {"none  N=192 dense", 192, 2048, 128, 3072, 128, 0, 0},
        {"none  N=256 dense", 256, 2048, 128, 4096, 128, 0, 0},
        {"sw128 N=64 ", 64, 16, 1024, 16, 1024, 2, 0},
        {"sw128 N=128", 128, 16, 1024, 16, 1024, 2, 0},
        {"sw128 N=256", 256, 16, 1024, 16, 1024, 2, 0},
    };
    for (const Cfg &c : cfgs) {
        for (int nacc : {1, 2, 4}) {
            if (nacc * c.N > 512) continue;
            for (int ce : {0, 1, 4}) {
                const int n = 500, grid = 148;
                auto k = nacc == 1 ? rate_kernel<1> : nacc == 2 ? rate_kernel<2> : rate_kernel<4>;
                k<<<grid, 128, 200 * 1024>>>(n, c.N, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, c.layout, c.a_step, ce, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost);
                printf("%-46s nacc=%d wait-every=%d x8  %7.1f cyc/MMA (ideal %d)\n", c.name, nacc, ce, (double)h_out / (n * 8), c.N / 2);
                fflush(stdout);
            }
        }
    }
    return 0;
}
