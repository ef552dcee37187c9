// Experiment (development aid): the filter-gradient GEMM on tcgen05.  dW_tap[o][i] = sum over pixels G[pix][o] * I[pix + tap][i]
// has K = pixels, so with NHWC tensors BOTH operands are "MN-major" (the M / N index is the contiguous one).  This test pins down,
// against a CPU sum: (1) TMA boxes [pixels][64 channels] with SWIZZLE_128B as MN-major SW128 operands (8 pixels x 128 B = one swizzle
// atom, SBO = 1024 between groups of 8 pixels, LBO = distance between blocks of 64 channels); (2) filter taps as start-address
// shifts by whole pixels (128 B) into the staged input patch, for the descriptor's base-offset field = 0 or (addr >> 7) & 7;
// (3) where an M = 64 accumulator lives in TMEM and whether a second one can sit at lane offset 16; (4) MMA rates per (M, N).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_mn_test umma_mn_test.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int H = 10, W = 48, CG = 128, CI = 64;
constexpr int TR = 2, TW = 32, IR = TR + 2, IC = TW + 2, GY0 = 3, GX0 = 5;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t bo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)bo << 49) |
           ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                 "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok && clock64() - t0 < 400000000ll)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// M = 128: A = both 64-channel blocks of G; M = 64: block 0 only.  tap2 >= 0 (M = 64 only): a second accumulator for that tap at lane offset 16.
__global__ void __launch_bounds__(128) test_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI, float *out, int M,
                                                   int tap, int tap2, int bo_mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    unsigned char *sG0 = smem, *sG1 = smem + 8192, *sI = smem + 16384;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // clear the accumulator columns so that untouched lanes read as zero
    {
        const uint32_t z = 0;
        for (int c = 0; c < 64; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(2 * TR * TW * 128 + IR * IC * 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sG0)), "l"(&tmG),
                     "r"(0), "r"(GX0), "r"(GY0), "r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sG1)), "l"(&tmG),
                     "r"(64), "r"(GX0), "r"(GY0), "r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sI)), "l"(&tmI),
                     "r"(0), "r"(GX0 - 1), "r"(GY0 - 1), "r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(CI >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int pass = 0; pass < (tap2 >= 0 ? 2 : 1); ++pass) {
            const int t = pass ? tap2 : tap, ky = t / 3, kx = t % 3;
            for (int r = 0; r < TR; ++r)
                for (int h = 0; h < TW / 16; ++h) {
                    const uint32_t a_addr = smem_u32(sG0) + (uint32_t)(r * TW + 16 * h) * 128u;
                    const uint32_t b_addr = smem_u32(sI) + (uint32_t)((r + ky) * IC + 16 * h + kx) * 128u;
                    const uint32_t bo = bo_mode == 1 ? (b_addr >> 7) & 7u : 0u;
                    mma(tmem + (pass ? (16u << 16) : 0u), make_desc(a_addr, 8192, 1024, 2, 0), make_desc(b_addr, 8192, 1024, 2, bo), idesc, (r | h) ? 1u : 0u);
                }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        wait_bar(&bar2, 0);
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 16; ++q) out[threadIdx.x * 64 + c0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

// The planned wgrad tile: A = the input patch, M = 128 = [64 input channels under tap ta ; the same under tap tb] through LBO = the address
// distance of the two taps; B = G with N = 64 or 96 (two channel boxes, LBO = box size).  out[m][n], m < 128, n < N.
__global__ void __launch_bounds__(128) pair_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI, float *out, int N, int ta,
                                                   int tb) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    unsigned char *sG0 = smem, *sG1 = smem + 8192, *sI = smem + 16384;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(2 * TR * TW * 128 + IR * IC * 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sG0)), "l"(&tmG),
                     "r"(0), "r"(GX0), "r"(GY0), "r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sG1)), "l"(&tmG),
                     "r"(64), "r"(GX0), "r"(GY0), "r"(smem_u32(&bar)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sI)), "l"(&tmI),
                     "r"(0), "r"(GX0 - 1), "r"(GY0 - 1), "r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t offa = (uint32_t)((ta / 3) * IC + ta % 3) * 128u, offb = (uint32_t)((tb / 3) * IC + tb % 3) * 128u;
        for (int r = 0; r < TR; ++r)
            for (int h = 0; h < TW / 16; ++h) {
                const uint32_t a_addr = smem_u32(sI) + (uint32_t)(r * IC + 16 * h) * 128u + offa;
                const uint32_t b_addr = smem_u32(sG0) + (uint32_t)(r * TW + 16 * h) * 128u;
                mma(tmem, make_desc(a_addr, offb - offa, 1024, 2, 0), make_desc(b_addr, 8192, 1024, 2, 0), idesc, (r | h) ? 1u : 0u);
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        wait_bar(&bar2, 0);
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 96; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                       "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 16; ++q) out[threadIdx.x * 96 + c0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

// rate: n_iter x 8 MMAs of shape (M, N, 16), both operands MN-major SW128, start addresses walking like a wgrad K loop
__global__ void __launch_bounds__(128) rate_kernel(int n_iter, int M, int N, int nacc, int mn_major, long long *out, uint32_t a_shift = 0, uint32_t a_lbo = 16384,
                                                   uint32_t b_shift = 128) {
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
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (1u << 15) | (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
        const long long t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // MN-major: 16 pixels = 2048 B per K step, LBO 16 KB between channel blocks; K-major SW128: 8-row groups 1024 B apart
                const uint64_t da = mn_major ? make_desc(a0 + (uint32_t)(u & 7) * 2048u + a_shift * (uint32_t)(1 + u % 3), a_lbo, 1024, 2, 0) : make_desc(a0 + (uint32_t)(u & 3) * 32u, 16, 1024, 2, 0);
                const uint64_t db = mn_major ? make_desc(b0 + (uint32_t)(u & 7) * 2048u + b_shift * (uint32_t)(u % 3), 16384, 1024, 2, 0)
                                             : make_desc(b0 + (uint32_t)(u & 3) * 32u, 16, 1024, 2, 0);
                mma(tmem + (uint32_t)((u % nacc) * N), da, db, idesc, (it | (u >= nacc)) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> g(H * W * CG), x(H * W * CI);
    std::vector<float> gf(H * W * CG), xf(H * W * CI);
    srand(3);
    for (size_t i = 0; i < g.size(); ++i) { gf[i] = (float)(rand() % 7 - 3); g[i] = __float2bfloat16(gf[i]); }
    for (size_t i = 0; i < x.size(); ++i) { xf[i] = (float)(rand() % 5 - 2); x[i] = __float2bfloat16(xf[i]); }
    // reference: ref[tap][o][i]
    std::vector<float> ref(9 * CG * CI, 0.f);
    for (int t = 0; t < 9; ++t)
        for (int r = 0; r < TR; ++r)
            for (int c = 0; c < TW; ++c) {
                const int gy = GY0 + r, gx = GX0 + c, iy = gy - 1 + t / 3, ix = gx - 1 + t % 3;
                for (int o = 0; o < CG; ++o)
                    for (int i = 0; i < CI; ++i) ref[(t * CG + o) * CI + i] += gf[(gy * W + gx) * CG + o] * xf[(iy * W + ix) * CI + i];
            }
    __nv_bfloat16 *dg, *dx;
    float *dout;
    cudaMalloc(&dg, g.size() * 2); cudaMalloc(&dx, x.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dg, g.data(), g.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tmG, tmI;
    {
        const cuuint64_t gdim[3] = {CG, W, H}, gstr[2] = {CG * 2, (cuuint64_t)W * CG * 2};
        const cuuint32_t box[3] = {64, TW, TR}, estr[3] = {1, 1, 1};
        CUresult r = encode(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode G failed %d\n", (int)r); return 1; }
    }
    {
        const cuuint64_t gdim[3] = {CI, W, H}, gstr[2] = {CI * 2, (cuuint64_t)W * CI * 2};
        const cuuint32_t box[3] = {64, IC, IR}, estr[3] = {1, 1, 1};
        CUresult r = encode(&tmI, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode I failed %d\n", (int)r); return 1; }
    }
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> got(128 * 64);
    auto run = [&](int M, int tap, int tap2, int bo) {
        cudaMemset(dout, 0, 128 * 64 * 4);
        test_kernel<<<1, 128, 64 * 1024>>>(tmG, tmI, dout, M, tap, tap2, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M %d tap %d bo %d: %s\n", M, tap, bo, cudaGetErrorString(e)); exit(1); }
        cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    };
    for (int bo = 0; bo < 2; ++bo)
        for (int tap = 0; tap < 9; ++tap) {
            run(128, tap, -1, bo);
            double maxerr = 0; int bad = 0;
            for (int o = 0; o < 128; ++o)
                for (int i = 0; i < 64; ++i) { const double d = fabs(got[o * 64 + i] - ref[(tap * CG + o) * CI + i]); if (d > maxerr) maxerr = d; if (d > 0.5) ++bad; }
            printf("M=128 MN-major SW128, base-offset mode %d, tap %d: max |err| %.1f, mismatches %d / 8192\n", bo, tap, maxerr, bad);
        }
    // M = 64: where do the rows go?  (two accumulators: tap 4 at lane offset 0, tap 7 at lane offset 16)
    run(64, 4, 7, 0);
    for (int which = 0; which < 2; ++which) {
        const int tap = which ? 7 : 4;
        printf("M=64 accumulator %d (tap %d): row -> lane:", which, tap);
        int found = 0;
        for (int o = 0; o < 64; ++o) {
            int at = -1;
            for (int lane = 0; lane < 128 && at < 0; ++lane) {
                bool same = true;
                for (int i = 0; i < 64 && same; ++i) same = fabs(got[lane * 64 + i] - ref[(tap * CG + o) * CI + i]) < 0.5;
                if (same) at = lane;
            }
            if (o % 16 == 0 || at < 0) printf(" %d->%d", o, at);
            found += at >= 0;
        }
        printf("  (%d / 64 rows found)\n", found);
    }
    // tap pairs through LBO
    {
        float *dout2;
        cudaMalloc(&dout2, 128 * 96 * 4);
        cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        std::vector<float> got2(128 * 96);
        const int pairs[5][2] = {{0, 1}, {0, 3}, {2, 6}, {4, 8}, {7, 8}};
        for (int N : {64, 96})
            for (auto &pr : pairs) {
                pair_kernel<<<1, 128, 64 * 1024>>>(tmG, tmI, dout2, N, pr[0], pr[1]);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("pair N %d: %s\n", N, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(got2.data(), dout2, got2.size() * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0; int bad = 0;
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < N; ++n) {
                        const int tap = pr[m / 64], i = m % 64;
                        const double d = fabs(got2[m * 96 + n] - ref[(tap * CG + n) * CI + i]);
                        if (d > maxerr) maxerr = d;
                        if (d > 0.5) ++bad;
                    }
                printf("A = input patch, taps (%d, %d) through LBO, N = %d: max |err| %.1f, mismatches %d / %d\n", pr[0], pr[1], N, maxerr, bad, 128 * N);
            }
    }
    // rates
    long long *d_cyc, h_cyc;
    cudaMalloc(&d_cyc, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int mn = 1; mn >= 0; --mn)
        for (int M : {64, 128})
            for (int N : {64, 128, 256})
                for (int nacc : {1, 3}) {
                    if (nacc * N > 512) continue;
                    rate_kernel<<<148, 128, 160 * 1024>>>(400, M, N, nacc, mn, d_cyc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("rate M %d N %d: %s\n", M, N, cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(&h_cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
                    printf("rate %s M=%d N=%d nacc=%d: %.1f cycles / MMA (math at full rate: %d)\n", mn ? "MN-major" : "K-major ", M, N, nacc, (double)h_cyc / (400 * 8), M * N / 256);
                }
    // operand alignment: tap shifts move the start address off the 1024-byte swizzle atom; LBO = 128 overlaps the two halves of A
    struct V { const char *name; uint32_t a_shift, a_lbo, b_shift; };
    const V vs[] = {{"A aligned, LBO 16 KB, B aligned", 0, 16384, 0}, {"A aligned, LBO 16 KB, B shifted 128", 0, 16384, 128}, {"A shifted 128, LBO 16 KB, B aligned", 128, 16384, 0},
                    {"A aligned, LBO 128, B aligned", 0, 128, 0}, {"A shifted 128, LBO 128, B aligned", 128, 128, 0}, {"A shifted 128, LBO 4352, B aligned", 128, 4352, 0},
                    {"A shifted 512, LBO 16 KB, B aligned", 512, 16384, 0}, {"A shifted 1024, LBO 1024, B aligned", 1024, 1024, 0}};
    for (const V &v : vs)
        for (int N : {64, 96}) {
            rate_kernel<<<148, 128, 160 * 1024>>>(400, 128, N, 3, 1, d_cyc, v.a_shift, v.a_lbo, v.b_shift);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rate variant: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(&h_cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("rate MN-major M=128 N=%d, %s: %.1f cycles / MMA\n", N, v.name, (double)h_cyc / (400 * 8));
        }
    return 0;
}
