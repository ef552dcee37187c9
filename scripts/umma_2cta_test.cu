// Experiment (development aid): cta_group::2 basics on sm_100a.  A cluster of two CTAs computes D[256 x N] = A[256 x 32] * B[N x 32]^T
// with ONE tcgen05.mma.cta_group::2 stream issued by the leader: each CTA holds its own 128 rows of A and HALF of B
// (N/2 rows), accumulators land in each CTA's TMEM (its 128 rows x N columns).  Checks the result against the CPU for
// a few hypotheses about which half of B lives where.  No-swizzle K-major operands filled with plain stores.
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int N = 64, K = 32;

__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok && clock64() - t0 < 200000000ll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// a: [256][K] bf16 row-major (global), b: [N][K]; out: [256][N] fp32.  variant: how B's halves are placed
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) test2(const __nv_bfloat16 *a, const __nv_bfloat16 *b, float *out, int variant, int *status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int warp = threadIdx.x >> 5;
    __nv_bfloat16 *sA = reinterpret_cast<__nv_bfloat16 *>(smem);               // [K/8][128][8]
    __nv_bfloat16 *sB = reinterpret_cast<__nv_bfloat16 *>(smem + 16 * 1024);   // [K/8][N/2][8]
    // A: this CTA's 128 rows
    for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        sA[((k / 8) * 128 + r) * 8 + k % 8] = a[(rank * 128 + r) * K + k];
    }
    // B: this CTA's half of the N rows
    for (int i = threadIdx.x; i < (N / 2) * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        const int src_n = variant == 0 ? rank * (N / 2) + n : (1 - rank) * (N / 2) + n;
        sB[((k / 8) * (N / 2) + n) * 8 + k % 8] = b[src_n * K + k];
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (rank == 0 && threadIdx.x == 0) {
        // M = 256 (cta_group::2), N, K = 16 per instruction
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        for (int k2 = 0; k2 < K / 16; ++k2) {
            const uint32_t a_addr = smem_u32(sA) + (uint32_t)k2 * 2u * 128u * 16u, b_addr = smem_u32(sB) + (uint32_t)k2 * 2u * (N / 2) * 16u;
            const uint64_t da = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)((128 * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint64_t db = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)(((N / 2) * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                         "l"(db), "r"(idesc), "r"((uint32_t)k2)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar_done)),
                     "h"((uint16_t)3)
                     : "memory");
    }
    // both CTAs wait on their own copy of the barrier (multicast commit arrives on both)
    const bool ok = wait_bar(smem_u32(&bar_done), 0);
    if (!ok && threadIdx.x == 0) status[rank] = 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int n0 = 0; n0 < N; n0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)n0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 16; ++q) out[(rank * 128 + threadIdx.x) * N + n0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

int main() {
    std::vector<__nv_bfloat16> a(256 * K), b(N * K);
    std::vector<float> af(256 * K), bf(N * K), ref(256 * N, 0.f);
    srand(2);
    for (int i = 0; i < 256 * K; ++i) { af[i] = (float)(rand() % 7 - 3); a[i] = __float2bfloat16(af[i]); }
    for (int i = 0; i < N * K; ++i) { bf[i] = (float)(rand() % 5 - 2); b[i] = __float2bfloat16(bf[i]); }
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) ref[m * N + n] += af[m * K + k] * bf[n * K + k];
    __nv_bfloat16 *da, *db;
    float *dout;
    int *dstat;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 256 * N * 4); cudaMalloc(&dstat, 8);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(test2, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(dout, 0, 256 * N * 4); cudaMemset(dstat, 0, 8);
        test2<<<2, 128, 48 * 1024>>>(da, db, dout, variant, dstat);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> got(256 * N);
        int st[2];
        cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(st, dstat, 8, cudaMemcpyDeviceToHost);
        int bad = 0, bad_lo = 0, bad_hi = 0;
        for (int i = 0; i < 256 * N; ++i) if (fabs(got[i] - ref[i]) > 0.5) { ++bad; if (i < 128 * N) ++bad_lo; else ++bad_hi; }
        // also test the column-swapped hypothesis
        int bad_sw = 0;
        for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) if (fabs(got[m * N + (n + N / 2) % N] - ref[m * N + n]) > 0.5) ++bad_sw;
        printf("variant %d (B half %s): timeouts %d %d, mismatches %d / %d (rows 0-127: %d, rows 128-255: %d); with halves swapped: %d\n", variant,
               variant == 0 ? "rank r holds rows r*N/2.." : "rank r holds the other half", st[0], st[1], bad, 256 * N, bad_lo, bad_hi, bad_sw);
        fflush(stdout);
    }
    return 0;
}
