// Experiment (development aid), stride-2 variant of tma_swz_test.cu: ONE dense TMA box [rows][cols][32 channels = 64 B] with SWIZZLE_128B, read by
// the MMA as SWIZZLE_128B K-major rows of TWO pixels (row pitch 128 B = input stride 2; odd-pixel taps start 64 B into the row).
// Original header: can the conv A operand be staged by ONE tensor-map TMA load per k-block (box = 32
// channels x halo columns x halo rows, SWIZZLE_64B) and the 3x3 taps still be "a different descriptor start address
// into the same staged plane"?  Compares a 3x3 conv computed that way against a CPU reference for several settings
// of the descriptor's base-offset field.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 (no -lcuda needed).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int H = 44, W = 30, C = 32, N = 16, ROWS = 33, COLS = 18, TY0 = 2, TX0 = 3;     // tile: 16 x 8 output pixels at (TY0, TX0), stride 2, pad 1

__global__ void __launch_bounds__(128) test_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16 *wpk, float *out,
                                                   int bo_mode, int swz, unsigned short *dump) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    unsigned char *sA = smem;                       // [ROWS][COLS][64 B] swizzled by TMA
    unsigned char *sB = smem + 100 * 1024;           // 9 taps x [C/8][N][8] no-swizzle K-major
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 9 * (C / 8) * N * 8; i += blockDim.x) reinterpret_cast<__nv_bfloat16 *>(sB)[i] = wpk[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(ROWS * COLS * 64)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sA)),
                     "l"(&tmap), "r"(0), "r"(2 * TX0 - 2), "r"(2 * TY0 - 1), "r"(smem_u32(&bar))
                     : "memory");
        uint32_t ok = 0;
        const long long t0 = clock64();
        while (!ok && clock64() - t0 < 400000000ll)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint32_t row_b = swz ? 64u : 64u;
        for (int tap = 0; tap < (bo_mode >= 10 ? 0 : 9); ++tap)
            for (int k2 = 0; k2 < 2; ++k2) {
                const uint32_t a_addr = smem_u32(sA) + (uint32_t)((tap / 3) * COLS + 1 + tap % 3) * row_b + (uint32_t)k2 * 32u;
                uint32_t bo = 0;
                if (bo_mode == 1) bo = (a_addr >> 7) & 7u;
                if (bo_mode == 2) bo = (a_addr >> 7) & 3u;
                if (bo_mode == 3) bo = ((a_addr >> 6) & 7u);
                // swizzled K-major: LBO ignored for K <= swizzle width, SBO = pitch between 8-row groups = plane row pitch
                const uint64_t da = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)(1) << 16) | ((uint64_t)((2 * COLS * row_b) >> 4) << 32) |
                                    (1ull << 46) | ((uint64_t)bo << 49) | ((uint64_t)2 << 61);
                const uint32_t b_addr = smem_u32(sB) + (uint32_t)tap * (C / 8) * N * 16 + (uint32_t)k2 * 2u * N * 16;
                const uint64_t db = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)((N * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                             "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(tap | k2))
                             : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        ok = 0;
        const long long t1 = clock64();
        while (!ok && clock64() - t1 < 400000000ll)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2)) : "memory");
    }
    __syncthreads();
    if (bo_mode == 10) for (int i = threadIdx.x; i < 50 * 1024; i += blockDim.x) dump[i] = reinterpret_cast<unsigned short *>(sA)[i];
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(32 * warp) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int q = 0; q < 16; ++q) out[threadIdx.x * 16 + q] = __uint_as_float(v[q]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u));
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    std::vector<__nv_bfloat16> x(H * W * C), wpk(9 * (C / 8) * N * 8);
    std::vector<float> xf(H * W * C), wf(9 * C * N);
    srand(1);
    for (int i = 0; i < H * W * C; ++i) { xf[i] = (float)(rand() % 255 - 127); x[i] = __float2bfloat16(xf[i]); }
    for (int t = 0; t < 9; ++t)
        for (int c = 0; c < C; ++c)
            for (int n = 0; n < N; ++n) {
                const float v = (float)(rand() % 5 - 2);
                wf[(t * C + c) * N + n] = v;
                wpk[((t * (C / 8) + c / 8) * N + n) * 8 + c % 8] = __float2bfloat16(v);
            }
    // CPU reference: out[m = 8 r + c][n] for output pixel (r, c), r < 16, c < 8, pad 1
    std::vector<float> ref(128 * N, 0.f);
    for (int r = 0; r < 16; ++r)
        for (int c = 0; c < 8; ++c)
            for (int t = 0; t < 9; ++t) {
                const int iy = 2 * (TY0 + r) - 1 + t / 3, ix = 2 * (TX0 + c) - 1 + t % 3;
                if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
                for (int ch = 0; ch < C; ++ch)
                    for (int n = 0; n < N; ++n) ref[(8 * r + c) * N + n] += xf[(iy * W + ix) * C + ch] * wf[(t * C + ch) * N + n];
            }
    __nv_bfloat16 *dx, *dw;
    float *dout;
    cudaMalloc(&dx, x.size() * 2); cudaMalloc(&dw, wpk.size() * 2); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, wpk.data(), wpk.size() * 2, cudaMemcpyHostToDevice);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    unsigned short *ddump; cudaMalloc(&ddump, 100 * 1024);
    for (int swz = 1; swz >= 1; --swz)
        for (int bo_mode : {10, 0, 1}) {
            CUtensorMap tmap;
            const cuuint64_t gdim[3] = {C, W, H};
            const cuuint64_t gstr[2] = {C * 2, (cuuint64_t)W * C * 2};
            const cuuint32_t box[3] = {32, COLS, ROWS};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            cudaMemset(dout, 0, 128 * 16 * 4);
            cudaMemset(ddump, 0xff, 100 * 1024);
            test_kernel<<<1, 128, 160 * 1024>>>(tmap, dw, dout, bo_mode, swz, ddump);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("swz %d bo_mode %d: %s\n", swz, bo_mode, cudaGetErrorString(e)); return 1; }
            if (bo_mode == 10) {
                std::vector<unsigned short> hd(50 * 1024);
                cudaMemcpy(hd.data(), ddump, 100 * 1024, cudaMemcpyDeviceToHost);
                // locate box pixel (row 0, col q) chunk ch by comparing 8 values with the source
                for (int q = 0; q < 6; ++q)
                    for (int ch = 0; ch < 4; ++ch) {
                        const int iy = 2 * TY0 - 1, ix = 2 * TX0 - 2 + q;
                        int found = -1;
                        for (int off = 0; off + 8 <= 50 * 1024 && found < 0; off += 8) {
                            bool same = true;
                            for (int e = 0; e < 8 && same; ++e) { unsigned short b; __nv_bfloat16 v = x[(iy * W + ix) * C + ch * 8 + e]; memcpy(&b, &v, 2); same = hd[off + e] == b; }
                            if (same) found = off * 2;
                        }
                        printf("box pixel (0,%d) chunk %d -> smem byte %d\n", q, ch, found);
                    }
                const int iy = 2 * TY0, ix = 2 * TX0 - 2;
                for (int off = 0; off + 8 <= 50 * 1024; off += 8) { bool same = true; for (int e = 0; e < 8 && same; ++e) { unsigned short b; __nv_bfloat16 v = x[(iy * W + ix) * C + e]; memcpy(&b, &v, 2); same = hd[off + e] == b; } if (same) { printf("box pixel (1,0) chunk 0 -> smem byte %d\n", off * 2); break; } }
            }
            std::vector<float> got(128 * 16);
            cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0; int bad = 0;
            for (int i = 0; i < 128 * 16; ++i) { const double d = fabs(got[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 0.5) ++bad; }
            printf("dense stride-2 plane, SWIZZLE_128B pair rows (%d), base_offset mode %d: max |err| = %.1f, mismatching outputs %d / 2048\n", swz, bo_mode, maxerr, bad);
            fflush(stdout);
        }
    return 0;
}
