// Micro-benchmark (development aid): issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, bf16) as a
// function of N, next to cta_group::1 (M = 128) in the same harness.  Operands are whatever is in shared memory; only the
// timing matters.  Each CTA holds its 128 rows of A and, in pair mode, N/2 rows of B.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <bool kPair>
__global__ void __launch_bounds__(128) rate2(int n_iter, int N, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint32_t rank = 0;
    if (kPair) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (kPair) cg::this_cluster().sync(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int nb = kPair ? N / 2 : N;                 // B rows staged per CTA
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((kPair ? 16u : 8u) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        const uint64_t da0 = (uint64_t)((a0 & 0x3FFFF) >> 4) | ((uint64_t)((128 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint64_t db0 = (uint64_t)((b0 & 0x3FFFF) >> 4) | ((uint64_t)((nb * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const long long t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint64_t da = da0 + (uint64_t)(u * 8), db = db0 + (uint64_t)((u & 3) * 2);
                    const uint32_t d = tmem + (uint32_t)((u & 1) * N), acc = (it | (u >= 2)) ? 1u : 0u;
                    if (kPair)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                    else
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
            }
            __syncwarp();
        }
        if (elect_one()) {
            if (kPair)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
            else
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        uint32_t ok = 0;
        const long long tw = clock64();
        while (!ok && clock64() - tw < 400000000ll)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = ok ? t1 - t0 : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (kPair) cg::this_cluster().sync(); else __syncthreads();
    if (warp == 0) {
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

// A operand as the conv kernel stages it: SWIZZLE_64B rows of 64 B (one pixel, 32 channels), an 8-row group = 8 consecutive
// pixels of a plane row, SBO = the plane's row pitch, start address shifted by the tap.  Does pitch / shift cost tensor-pipe cycles?
__global__ void __launch_bounds__(128) rate_sw64(int n_iter, int N, int sbo_bytes, int shift_bytes, long long *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
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
        uint32_t a0 = smem_u32(smem);
        a0 = (a0 + 1023u) & ~1023u;
        const uint32_t b0 = a0 + 96 * 1024;
        const uint64_t da0 = (uint64_t)(((a0 + (uint32_t)shift_bytes) & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (4ull << 61);
        const uint64_t db0 = (uint64_t)((b0 & 0x3FFFF) >> 4) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const long long t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    // 8 MMAs: two K16 halves x four "taps" (start shifted by 0, 1, 2 pixels and one row)
                    const uint32_t tapoff = (u >> 1) == 3 ? (uint32_t)sbo_bytes : (uint32_t)((u >> 1) * 64);
                    const uint64_t da = da0 + (uint64_t)((tapoff >> 4) + (u & 1) * 2), db = db0 + (uint64_t)((u & 1) * 2 * N);
                    const uint32_t d = tmem + (uint32_t)((u >> 2) * N), acc = (it | (u >= 4)) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        const long long tw = clock64();
        while (!ok && clock64() - tw < 400000000ll)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = ok ? t1 - t0 : -1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    long long *d_out, h_out;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(rate2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(rate2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int n = 500;
    cudaFuncSetAttribute(rate_sw64, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int N : {16, 64, 128})
        for (int sbo : {512, 1024, 1152, 2176})
            for (int shift : {0, 64, 1152 + 64}) {
                rate_sw64<<<148, 128, 200 * 1024>>>(n, N, sbo, shift, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sw64 N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost);
                printf("SW64 A  N=%3d  SBO=%4d B  start shift=%4d B: %6.1f cycles / MMA\n", N, sbo, shift, (double)h_out / (n * 8));
                fflush(stdout);
            }
    for (int N : {64, 128, 160, 256}) {
        rate2<false><<<148, 128, 160 * 1024>>>(n, N, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("single N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost);
        printf("cta_group::1  M=128 N=%3d: %7.1f cycles / MMA (ideal %d)\n", N, (double)h_out / (n * 8), N / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 160 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, rate2<true>, n, N, d_out);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h_out, d_out, 8, cudaMemcpyDeviceToHost);
        printf("cta_group::2  M=256 N=%3d: %7.1f cycles / MMA (ideal %d per SM pair, i.e. the same per-SM rate)\n", N, (double)h_out / (n * 8), N / 2);
        fflush(stdout);
    }
    return 0;
}
