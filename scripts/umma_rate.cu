// Micro-benchmark (development aid, not product): tcgen05.mma issue/execute rate as a function of the shared-memory
// operand layout.  Measures cycles per MMA (M=128, K=16, bf16) for no-swizzle K-major descriptors with various
// LBO/SBO and for the canonical 128B-swizzle layout.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128) rate_kernel(int n_mma, int N, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
                                                   uint32_t b_sbo, uint32_t layout, uint32_t a_step, long long *out) {
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
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const uint64_t da = make_desc(a0 + (uint32_t)(i & 7) * a_step, a_lbo, a_sbo, layout);
            const uint64_t db = make_desc(b0, b_lbo, b_sbo, layout);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (uint32_t)((i & 1) * N)),
                "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(i > 1))
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    long long *d_out, h_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct Cfg { const char *name; int N; uint32_t a_lbo, a_sbo, b_lbo, b_sbo, layout, a_step; };
    const Cfg cfgs[] = {
        {"none  N=64  A:lbo=9856 sbo=544 (conv halo)  B:lbo=1024 sbo=128", 64, 9856, 544, 1024, 128, 0, 16},
        {"none  N=128 A:lbo=9856 sbo=544 (conv halo)  B:lbo=2048 sbo=128", 128, 9856, 544, 2048, 128, 0, 16},
        {"none  N=64  A:lbo=2048 sbo=128 (dense)      B:lbo=1024 sbo=128", 64, 2048, 128, 1024, 128, 0, 0},
        {"none  N=128 A:lbo=2048 sbo=128 (dense)      B:lbo=2048 sbo=128", 128, 2048, 128, 2048, 128, 0, 0},
        {"none  N=256 A:lbo=2048 sbo=128 (dense)      B:lbo=4096 sbo=128", 256, 2048, 128, 4096, 128, 0, 0},
        {"none  N=128 A:lbo=128 sbo=256 (k-interleaved) B:lbo=128 sbo=256", 128, 128, 256, 128, 256, 0, 0},
        {"none  N=128 A:lbo=9872 sbo=560              B:lbo=2064 sbo=144", 128, 9872, 560, 2064, 144, 0, 16},
        {"sw128 N=64  sbo=1024", 64, 16, 1024, 16, 1024, 2, 0},
        {"sw128 N=128 sbo=1024", 128, 16, 1024, 16, 1024, 2, 0},
        {"sw128 N=256 sbo=1024", 256, 16, 1024, 16, 1024, 2, 0},
    };
    for (const Cfg &c : cfgs) {
        for (int grid : {1, 148}) {
            const int n = 2000;
            rate_kernel<<<grid, 128, 200 * 1024>>>(n, c.N, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, c.layout, c.a_step, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost);
            printf("%-70s grid=%3d  %7.1f cyc/MMA (ideal %d)\n", c.name, grid, (double)h_out / n, c.N / 2);
        }
    }
    return 0;
}
