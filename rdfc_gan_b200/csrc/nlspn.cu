// NLSPN: fused offset/affinity stage and the propagation loop (fp32, NCHW, channels_f == 1, k_f == 3).
//
// Replaces nlspn/nlspn_model.py:68-138 (NLPSN._get_offset_affinity: a 3x3 conv, 8 ModulatedDeformConvFunction calls
// with a 1x1 kernel and ~25 small ATen kernels) by ONE kernel, and nlspn_model.py:140-144,166-173 (prop_time x
// {columns alloc, im2col, addmm(K=9,N=1), permute+contiguous}) by one gather kernel per iteration with no
// intermediate buffer.  The propagation is HBM / L2 bound: per pixel and iteration it streams 16 offsets + 9
// affinities (the centre tap's offsets are identically zero and are not read), gathers 36 feature corners through
// L1, and writes one float.  The host wrapper walks the batch in L2-sized image groups so that iterations 2..T of a
// group find their offset/affinity planes in the 126 MB L2 instead of HBM.
#include <stdlib.h>

#include "common.cuh"

namespace rdfc {
namespace {

constexpr int TX = 32, TY = 8;   // pixel tile of a CTA (one warp = one 32-pixel row segment -> coalesced planes)

// streaming load: read-only, do not allocate in L1 (keeps L1 for the gathered feature map)
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// bilinear sample with the DCN validity rule and per-corner zeroing (deformconv/src/cuda/modulated_deform_im2col_cuda.cuh:25-54,180)
__device__ __forceinline__ float bilinear(const float *__restrict__ im, int H, int W, float y, float x) {
    if (!(y > -1.f && x > -1.f && y < (float)H && x < (float)W)) return 0.f;
    const float fy = floorf(y), fx = floorf(x);
    const int yl = (int)fy, xl = (int)fx;
    const float ly = y - fy, lx = x - fx, hy = 1.f - ly, hx = 1.f - lx;
    const bool y0 = yl >= 0, y1 = yl + 1 <= H - 1, x0 = xl >= 0, x1 = xl + 1 <= W - 1;
    const float *r0 = im + (long long)yl * W + xl;
    const float v1 = (y0 && x0) ? __ldg(r0) : 0.f;
    const float v2 = (y0 && x1) ? __ldg(r0 + 1) : 0.f;
    const float v3 = (y1 && x0) ? __ldg(r0 + W) : 0.f;
    const float v4 = (y1 && x1) ? __ldg(r0 + W + 1) : 0.f;
    return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

// ------------------------------------------------------------------------------------------------------------------
// Offset / affinity stage.  grid (ceil(W/TX), ceil(H/TY), B), block TX*TY.
//   1. conv_offset_aff: 8 -> 24 channels, 3x3, pad 1 (guidance tile + halo staged in shared memory, weights too)
//   2. offsets: tap j (of the 8 non-centre taps) takes conv channels (2j, 2j+1) as (dy, dx)   (nlspn_model.py:76)
//   3. affinity: AS/ASS raw, TC tanh/scale, TGASS tanh/(scale+1e-8)                           (:82-87)
//   4. conf_prop: aff_j *= bilinear(confidence, h + dy_j, w + dx_j)  (1x1 kernel, pad 0)       (:96-119)
//   5. normalise by max(sum|aff| + 1e-4, 1) (ASS/TGASS) or sum|aff| + 1e-4 (AS); aff_ref = 1 - sum  (:122-136)
__global__ void __launch_bounds__(TX *TY) nlspn_affinity_kernel(const float *__restrict__ guidance,
                                                                const float *__restrict__ confidence,
                                                                const float *__restrict__ conv_w,
                                                                const float *__restrict__ conv_b,
                                                                const float *__restrict__ aff_scale, int affinity,
                                                                int conf_prop, float *__restrict__ offset,
                                                                float *__restrict__ aff, int H, int W) {
    __shared__ __align__(16) float s_w[72 * 24];     // [(ci*9 + tap)][24 outputs]: one LDS.128 feeds 4 FMAs
    __shared__ float s_b[24];
    __shared__ float s_g[8][TY + 2][TX + 2];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const long long P = (long long)H * W;
    for (int e = tid; e < 24 * 72; e += TX * TY) s_w[(e % 72) * 24 + e / 72] = conv_w[e];
    if (tid < 24) s_b[tid] = conv_b[tid];
    for (int e = tid; e < 8 * (TY + 2) * (TX + 2); e += TX * TY) {
        const int c = e / ((TY + 2) * (TX + 2)), r = e % ((TY + 2) * (TX + 2));
        const int yy = y0 + r / (TX + 2) - 1, xx = x0 + r % (TX + 2) - 1;
        s_g[c][r / (TX + 2)][r % (TX + 2)] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(guidance + ((long long)b * 8 + c) * P + (long long)yy * W + xx) : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;

    float o[24];
#pragma unroll
    for (int c = 0; c < 24; ++c) o[c] = s_b[c];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float g = s_g[ci][threadIdx.y + t / 3][threadIdx.x + t % 3];
            const float4 *w4 = reinterpret_cast<const float4 *>(s_w + (ci * 9 + t) * 24);
#pragma unroll
            for (int c4 = 0; c4 < 6; ++c4) {
                const float4 w = w4[c4];
                o[4 * c4 + 0] = fmaf(w.x, g, o[4 * c4 + 0]);
                o[4 * c4 + 1] = fmaf(w.y, g, o[4 * c4 + 1]);
                o[4 * c4 + 2] = fmaf(w.z, g, o[4 * c4 + 2]);
                o[4 * c4 + 3] = fmaf(w.w, g, o[4 * c4 + 3]);
            }
        }

    const float scale = __ldg(aff_scale);
    float a[8];
    float abs_sum = 0.f;
    const float *conf = confidence + (long long)b * P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = o[16 + j];
        if (affinity == RDFC_AFF_TC) v = tanhf(v) / scale;
        else if (affinity == RDFC_AFF_TGASS) v = tanhf(v) / (scale + 1e-8f);
        if (conf_prop) v *= bilinear(conf, H, W, (float)y + o[2 * j], (float)x + o[2 * j + 1]);
        a[j] = v;
        abs_sum += fabsf(v);
    }
    abs_sum += 1e-4f;
    if (affinity == RDFC_AFF_ASS || affinity == RDFC_AFF_TGASS) abs_sum = abs_sum < 1.f ? 1.f : abs_sum;
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (affinity != RDFC_AFF_TC) a[j] = a[j] / abs_sum;
        sum += a[j];
    }
    const long long pix = (long long)y * W + x;
    float *offp = offset + (long long)b * 18 * P + pix;
    float *affp = aff + (long long)b * 9 * P + pix;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = j < 4 ? j : j + 1;   // centre tap inserted at index 4
        offp[(long long)(2 * k) * P] = o[2 * j];
        offp[(long long)(2 * k + 1) * P] = o[2 * j + 1];
        affp[(long long)k * P] = a[j];
    }
    offp[8 * P] = 0.f;
    offp[9 * P] = 0.f;
    affp[4 * P] = 1.f - sum;
}

// ------------------------------------------------------------------------------------------------------------------
// One propagation iteration over images [0, nb) of the pointers given.  grid (ceil(W/TX), ceil(H/TY), nb).
//   out(h,w) = sum_k aff_k(h,w) * bilinear(in, h - 1 + k/3 + dy_k, w - 1 + k%3 + dx_k)
template <bool kClamp>
__global__ void __launch_bounds__(TX *TY) nlspn_prop_kernel(const float *__restrict__ in,
                                                            const float *__restrict__ offset,
                                                            const float *__restrict__ aff, float *__restrict__ out,
                                                            float *__restrict__ inter, int H, int W) {
    const int x = blockIdx.x * TX + threadIdx.x, y = blockIdx.y * TY + threadIdx.y, b = blockIdx.z;
    if (x >= W || y >= H) return;
    const long long P = (long long)H * W, pix = (long long)y * W + x;
    const float *offp = offset + (long long)b * 18 * P + pix;
    const float *affp = aff + (long long)b * 9 * P + pix;
    const float *im = in + (long long)b * P;

    // issue all streaming loads first (25 independent requests in flight per thread)
    float dy[9], dx[9], a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        a[k] = ld_stream(affp + (long long)k * P);
        if (k == 4) {
            dy[k] = 0.f;
            dx[k] = 0.f;
        } else {
            dy[k] = ld_stream(offp + (long long)(2 * k) * P);
            dx[k] = ld_stream(offp + (long long)(2 * k + 1) * P);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        acc = fmaf(a[k], bilinear(im, H, W, (float)(y - 1 + k / 3) + dy[k], (float)(x - 1 + k % 3) + dx[k]), acc);
    if (kClamp) acc = fminf(fmaxf(acc, -1.f), 1.f);
    out[(long long)b * P + pix] = acc;
    if (inter) inter[(long long)b * P + pix] = acc;
}


// ------------------------------------------------------------------------------------------------------------------
// Band kernel (the default): the thread-per-pixel kernel above is bound by the L1 tag stage (3.3 wavefronts per
// gather request with 2-px offsets, 769 instructions per warp, DRAM only 41 % busy: profiles/).  Here a CTA owns a
// contiguous run of image rows and first stages the feature rows its taps can reach (+- halo) in shared memory with
// a zero border, so that
//   * the 36 corner reads per pixel are LDS (bank-limited, no tag lookups) with NO bounds predicates: the DCN
//     validity rule and per-corner zeroing (modulated_deform_im2col_cuda.cuh:25-54,180) fall out of clamping the
//     sample position to [-1, H] x [-1, W] and reading zeros from the border;
//   * the 25 streamed planes are read as 8-byte vectors, two pixels per thread, straight from HBM (no L1 allocation);
//   * taps that leave the staged band (|offset| > halo) take the global-memory path of `bilinear` (rare).
// Rows are dealt evenly over one wave of CTAs (a run may span two images).
// Nine-tap gather from a staged band with the escape test hoisted out of the tap loop: the common case (no tap leaves
// the band) is straight-line code with 36 independent LDS, so the loads of all taps overlap instead of serialising
// behind a per-tap branch.  dy/dx/a: the pixel's 9 offsets and affinities; (y, x): the pixel.
struct BandView {
    const float *tile;      // [nrows][pitch], tile row 0 = image row ty0, tile column tc = image column tc - 1
    const float *im;        // the image in global memory (fallback path)
    int pitch, ty0, nrows, H, W;
    float Hf, Wf;
};
template <typename FDy, typename FDx, typename FA>
__device__ __forceinline__ float gather9(const BandView &v, FDy dyf, FDx dxf, FA af, int y, int x) {
    // pass 1: does any tap leave the staged band?  (rows only: columns are staged over the full width)
    float ymin = 1e30f, ymax = -1e30f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float cy = fminf(fmaxf((float)(y - 1 + k / 3) + dyf(k), -1.f), v.Hf);
        ymin = fminf(ymin, cy);
        ymax = fmaxf(ymax, cy);
    }
    const bool esc = ((int)floorf(ymin) < v.ty0) | ((int)floorf(ymax) + 1 >= v.ty0 + v.nrows);
    float acc = 0.f;
    if (!esc) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float cy = fminf(fmaxf((float)(y - 1 + k / 3) + dyf(k), -1.f), v.Hf);
            const float cx = fminf(fmaxf((float)(x - 1 + k % 3) + dxf(k), -1.f), v.Wf);
            const float fy = floorf(cy), fx = floorf(cx);
            const float ly = cy - fy, lx = cx - fx, hy = 1.f - ly, hx = 1.f - lx;
            // tile index in float (exact: < 2^24), one conversion per tap instead of two plus an integer multiply-add
            const float *q = v.tile + (int)fmaf(fy - (float)v.ty0, (float)v.pitch, fx + 1.f);
            acc = fmaf(af(k), hy * hx * q[0] + hy * lx * q[1] + ly * hx * q[v.pitch] + ly * lx * q[v.pitch + 1], acc);
        }
    } else {                     // some tap left the staged band: global-memory path
#pragma unroll
        for (int k = 0; k < 9; ++k)
            acc = fmaf(af(k), bilinear(v.im, v.H, v.W, (float)(y - 1 + k / 3) + dyf(k), (float)(x - 1 + k % 3) + dxf(k)), acc);
    }
    return acc;
}

constexpr int BAND_THREADS = 256;

__device__ __forceinline__ float2 ld_stream2(const float *p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

template <bool kClamp, int kPix, int kMinBlocks>
__global__ void __launch_bounds__(BAND_THREADS, kMinBlocks) nlspn_prop_band_kernel(const float *__restrict__ in,
                                                                          const float *__restrict__ offset,
                                                                          const float *__restrict__ aff,
                                                                          float *__restrict__ out, float *__restrict__ inter,
                                                                          int B, int H, int W, int rows_per_cta, int halo, int prefetch,
                                                                          int pre) {
    extern __shared__ float band_tile[];
    const int pitch = W + 4;                         // tile column tc <-> image column tc - 1 (-1 .. W + 2)
    const long long total = (long long)B * H, P = (long long)H * W;
    long long row = (long long)blockIdx.x * rows_per_cta;
    const long long row_end = row + rows_per_cta < total ? row + rows_per_cta : total;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = W / kPix;
    const float Hf = (float)H, Wf = (float)W;
    // Programmatic dependent launch: the next iteration's grid may take SM slots as this grid's CTAs retire (its CTAs then
    // sit in griddepcontrol.wait until this grid has completed and flushed), which removes the launch gap and the CTA
    // ramp-up between the 18 launches.  Nothing of the previous iteration is read before the wait; the offsets /
    // affinities do not depend on it, so the lines of the thread's first `pre` items are sent on their way to L2 first.
    if (pre && row < row_end) {
        const int b = (int)(row / H), y_lo = (int)(row - (long long)b * H);
        const int nr = (int)((row_end - row) < (long long)(H - y_lo) ? (row_end - row) : (long long)(H - y_lo));
        for (int i = 0; i < pre; ++i) {
            const int idx = threadIdx.x + i * BAND_THREADS;
            if (idx < nr * Wp && (lane & 15) == 0) {         // one prefetch per 128-byte line (16 lanes x 8 bytes)
                const int ry = idx / Wp, x = kPix * (idx - ry * Wp);
                const long long pix = (long long)(y_lo + ry) * W + x;
                const float *offp = offset + (long long)b * 18 * P + pix;
                const float *affp = aff + (long long)b * 9 * P + pix;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(affp + (long long)k * P));
                    if (k != 4) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(offp + (long long)(2 * k) * P));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(offp + (long long)(2 * k + 1) * P));
                    }
                }
            }
        }
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    while (row < row_end) {
        const int b = (int)(row / H), y_lo = (int)(row - (long long)b * H);
        const int nr = (int)((row_end - row) < (long long)(H - y_lo) ? (row_end - row) : (long long)(H - y_lo));   // rows of this image
        const int ty0 = y_lo - halo - 1;             // image row of tile row 0
        const int nrows = nr + 2 * halo + 3;
        const float *im = in + (long long)b * P;
        __syncthreads();                              // previous sub-band is done with the tile
        for (int tr = warp; tr < nrows; tr += BAND_THREADS / 32) {
            const int iy = ty0 + tr;
            const bool yok = iy >= 0 && iy < H;
            const float *src = im + (long long)iy * W - 1;
            float *dst = band_tile + tr * pitch;
            for (int tc = lane; tc < pitch; tc += 32) dst[tc] = (yok && tc >= 1 && tc <= W) ? __ldg(src + tc) : 0.f;
        }
        __syncthreads();
        const BandView bv{band_tile, im, pitch, ty0, nrows, H, W, Hf, Wf};
        const int nitems = nr * Wp;
        for (int idx = threadIdx.x; idx < nitems; idx += BAND_THREADS) {
            const int ry = idx / Wp, x = kPix * (idx - ry * Wp), y = y_lo + ry;
            const long long pix = (long long)y * W + x;
            const float *offp = offset + (long long)b * 18 * P + pix;
            const float *affp = aff + (long long)b * 9 * P + pix;
            const long long o = (long long)b * P + pix;
            if (prefetch && idx + BAND_THREADS < nitems) {
                // the thread's next item: start its 25 plane lines on their way from HBM to L2 while this one is gathered
                const int idn = idx + BAND_THREADS, ryn = idn / Wp;
                const long long dpix = (long long)(y_lo + ryn) * W + kPix * (idn - ryn * Wp) - pix;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(affp + (long long)k * P + dpix));
                    if (k != 4) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(offp + (long long)(2 * k) * P + dpix));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(offp + (long long)(2 * k + 1) * P + dpix));
                    }
                }
            }
            if (kPix == 2) {
                float2 a[9], dy[9], dx[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    a[k] = ld_stream2(affp + (long long)k * P);
                    if (k == 4) {
                        dy[k] = dx[k] = make_float2(0.f, 0.f);
                    } else {
                        dy[k] = ld_stream2(offp + (long long)(2 * k) * P);
                        dx[k] = ld_stream2(offp + (long long)(2 * k + 1) * P);
                    }
                }
                float acc0 = gather9(bv, [&](int k) { return dy[k].x; }, [&](int k) { return dx[k].x; }, [&](int k) { return a[k].x; }, y, x);
                float acc1 = gather9(bv, [&](int k) { return dy[k].y; }, [&](int k) { return dx[k].y; }, [&](int k) { return a[k].y; }, y, x + 1);
                if (kClamp) {
                    acc0 = fminf(fmaxf(acc0, -1.f), 1.f);
                    acc1 = fminf(fmaxf(acc1, -1.f), 1.f);
                }
                *reinterpret_cast<float2 *>(out + o) = make_float2(acc0, acc1);
                if (inter) *reinterpret_cast<float2 *>(inter + o) = make_float2(acc0, acc1);
            } else {
                float a[9], dy[9], dx[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    a[k] = ld_stream(affp + (long long)k * P);
                    if (k == 4) {
                        dy[k] = dx[k] = 0.f;
                    } else {
                        dy[k] = ld_stream(offp + (long long)(2 * k) * P);
                        dx[k] = ld_stream(offp + (long long)(2 * k + 1) * P);
                    }
                }
                float acc = gather9(bv, [&](int k) { return dy[k]; }, [&](int k) { return dx[k]; }, [&](int k) { return a[k]; }, y, x);
                if (kClamp) acc = fminf(fmaxf(acc, -1.f), 1.f);
                out[o] = acc;
                if (inter) inter[o] = acc;
            }
        }
        row += nr;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row-streaming propagation kernel (the default when W % 4 == 0): persistent CTAs, one image row per trip, one
// thread per pixel.  The 25 offset/affinity planes of a row are staged in shared memory by TMA 1-D bulk copies
// (cp.async.bulk, W*4 bytes each) onto an mbarrier, `stages` rows ahead of the compute, so HBM requests stay in
// flight while the warps gather feature corners through L1.  A CTA owns a contiguous band of rows: the feature
// rows its taps touch stay L1-resident from one trip to the next.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if (clock64() - t0 > 8000000000ll) __trap();      // protocol bug: fail loudly instead of hanging the GPU
    }
}

template <bool kClamp, int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) nlspn_prop_rows_kernel(const float *__restrict__ in,
                                                               const float *__restrict__ offset,
                                                               const float *__restrict__ aff, float *__restrict__ out,
                                                               float *__restrict__ inter, int B, int H, int W,
                                                               int stages, int rows_per_cta) {
    extern __shared__ __align__(128) unsigned char nl_smem[];
    float *stage_base = reinterpret_cast<float *>(nl_smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(nl_smem + (size_t)stages * 25 * W * sizeof(float));
    const long long total = (long long)B * H, P = (long long)H * W;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < total ? r0 + rows_per_cta : total;
    if (r0 >= r1) return;
    const int x = threadIdx.x;
    if (x == 0) {
        for (int s = 0; s < stages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t row_bytes = (uint32_t)W * 4u;
    // threads 0..24 each issue ONE plane's bulk copy (a single issuing thread serialises ~25 x 80 cycles per row)
    auto issue = [&](long long row, int s) {
        if (x >= 25) return;
        const int b = (int)(row / H), y = (int)(row % H);
        const uint32_t bar = smem_u32(bars + s);
        if (x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(25u * row_bytes) : "memory");
        const uint32_t dst = smem_u32(stage_base + (size_t)s * 25 * W) + (uint32_t)x * row_bytes;
        const float *src;
        if (x < 16) {
            const int j = x >> 1, k = j < 4 ? j : j + 1;       // slot 2j / 2j+1 <- offset channels 2k / 2k+1
            src = offset + ((long long)b * 18 + 2 * k + (x & 1)) * P + (long long)y * W;
        } else {
            src = aff + ((long long)b * 9 + (x - 16)) * P + (long long)y * W;
        }
        bulk_g2s(dst, src, row_bytes, bar);
    };
    for (int s = 0; s < stages; ++s)
        if (r0 + s < r1) issue(r0 + s, s);

    int it = 0;
    for (long long row = r0; row < r1; ++row, ++it) {
        const int s = it % stages;
        mbar_wait_parity(smem_u32(bars + s), (uint32_t)((it / stages) & 1));
        if (x < W) {
            const int b = (int)(row / H), y = (int)(row % H);
            const float *st = stage_base + (size_t)s * 25 * W + x;
            const float *im = in + (long long)b * P;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int j = k < 4 ? k : k - 1;
                const float dy = k == 4 ? 0.f : st[(2 * j) * W], dx = k == 4 ? 0.f : st[(2 * j + 1) * W];
                acc = fmaf(st[(16 + k) * W], bilinear(im, H, W, (float)(y - 1 + k / 3) + dy, (float)(x - 1 + k % 3) + dx), acc);
            }
            if (kClamp) acc = fminf(fmaxf(acc, -1.f), 1.f);
            const long long o = (long long)b * P + (long long)y * W + x;
            out[o] = acc;
            if (inter) inter[o] = acc;
        }
        __syncthreads();                               // everyone is done reading stage s
        if (row + stages < r1) issue(row + stages, s);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Ring kernel (the default when W % 4 == 0): the band kernel's shared-memory gathers + the row kernel's TMA stream,
// warp-specialised so that neither waits for the other.  One persistent CTA per SM owns a contiguous run of rows:
//   * producer warp: for every stage (G rows) 25 x G bulk copies (cp.async.bulk, one image row of one plane each)
//     onto an mbarrier, S stages ahead, released by the consumer warps through an "empty" mbarrier;
//   * consumer warps: stage the feature rows their taps can reach (+- halo, zero border) in shared memory once per
//     sub-band, then per stage read the 25 streamed values of a pixel from shared memory (conflict-free) and gather
//     the 36 corners with LDS; taps outside the staged band fall back to `bilinear` on global memory.
// HBM requests therefore stay in flight during the gathers (the band kernel's load -> wait -> compute phases left DRAM
// 47 % busy), and no thread holds the streamed values in registers.
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <bool kClamp>
__global__ void __launch_bounds__(992, 1) nlspn_prop_ring_kernel(const float *__restrict__ in, const float *__restrict__ offset,
                                                                 const float *__restrict__ aff, float *__restrict__ out,
                                                                 float *__restrict__ inter, int B, int H, int W,
                                                                 int rows_per_cta, int halo, int G, int S, int nr_max) {
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int pitch = W + 4;
    float *ring = reinterpret_cast<float *>(ring_smem);                                  // [S][G][25][W]
    float *tile = ring + (size_t)S * G * 25 * W;                                         // [nr_max + 2 halo + 3][pitch]
    uint64_t *bars = reinterpret_cast<uint64_t *>(tile + (size_t)(nr_max + 2 * halo + 3) * pitch);   // full[S], empty[S]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncw = (blockDim.x >> 5) - 1;                                               // consumer warps; warp ncw produces
    const long long total = (long long)B * H, P = (long long)H * W;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < total ? r0 + rows_per_cta : total;
    if (r0 >= r1) return;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + S + s)), "r"(ncw));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t row_bytes = (uint32_t)W * 4u;
    if (warp == ncw) {
        // ---------------- producer
        int i = 0;
        for (long long row = r0; row < r1; row += G, ++i) {
            const int s = i % S;
            const int ng = (int)((r1 - row) < G ? (r1 - row) : G);
            mbar_wait_parity(smem_u32(bars + S + s), (uint32_t)(((i / S) & 1) ^ 1));      // consumers released the slot
            const uint32_t full = smem_u32(bars + s);
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(25u * row_bytes * (uint32_t)ng) : "memory");
            __syncwarp();
            for (int c = lane; c < 25 * ng; c += 32) {
                const int g = c / 25, k = c - 25 * g;
                const long long rg = row + g;
                const int b = (int)(rg / H), y = (int)(rg - (long long)b * H);
                const float *src;
                if (k < 16) {
                    const int j = k >> 1, kk = j < 4 ? j : j + 1;      // slot 2j / 2j+1 <- offset channels 2kk / 2kk+1
                    src = offset + ((long long)b * 18 + 2 * kk + (k & 1)) * P + (long long)y * W;
                } else {
                    src = aff + ((long long)b * 9 + (k - 16)) * P + (long long)y * W;
                }
                bulk_g2s(smem_u32(ring + ((size_t)(s * G + g) * 25 + k) * W), src, row_bytes, full);
            }
        }
        return;
    }
    // ---------------- consumers
    const int nseg = (W + 31) >> 5;
    const float Hf = (float)H, Wf = (float)W;
    const int nthr_c = ncw * 32;
    int i = 0;
    long long row = r0;
    while (row < r1) {
        const int b = (int)(row / H), y_lo = (int)(row - (long long)b * H);
        int nr = (int)((r1 - row) < (long long)(H - y_lo) ? (r1 - row) : (long long)(H - y_lo));
        if (nr > nr_max) nr = nr_max;
        const int ty0 = y_lo - halo - 1, nrows = nr + 2 * halo + 3;
        const float *im = in + (long long)b * P;
        asm volatile("bar.sync 1, %0;" ::"r"(nthr_c) : "memory");          // previous sub-band is done with the tile
        for (int tr = warp; tr < nrows; tr += ncw) {
            const int iy = ty0 + tr;
            const bool yok = iy >= 0 && iy < H;
            const float *src = im + (long long)iy * W - 1;
            float *dst = tile + tr * pitch;
            for (int tc = lane; tc < pitch; tc += 32) dst[tc] = (yok && tc >= 1 && tc <= W) ? __ldg(src + tc) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nthr_c) : "memory");
        const BandView bv{tile, im, pitch, ty0, nrows, H, W, Hf, Wf};
        for (int rr = 0; rr < nr; rr += G, ++i) {
            const int s = i % S;
            mbar_wait_parity(smem_u32(bars + s), (uint32_t)((i / S) & 1));
            for (int seg = warp; seg < G * nseg; seg += ncw) {
                const int g = seg / nseg, x = (seg - g * nseg) * 32 + lane, y = y_lo + rr + g;
                if (rr + g < nr && x < W) {
                    const float *st = ring + (size_t)(s * G + g) * 25 * W + x;
                    float acc = gather9(bv,
                                        [&](int k) { return k == 4 ? 0.f : st[(2 * (k < 4 ? k : k - 1)) * W]; },
                                        [&](int k) { return k == 4 ? 0.f : st[(2 * (k < 4 ? k : k - 1) + 1) * W]; },
                                        [&](int k) { return st[(16 + k) * W]; }, y, x);
                    if (kClamp) acc = fminf(fmaxf(acc, -1.f), 1.f);
                    const long long o = (long long)b * P + (long long)y * W + x;
                    out[o] = acc;
                    if (inter) inter[o] = acc;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(smem_u32(bars + S + s));        // this warp is done reading the stage
        }
        row += nr;
    }
}

// feat = (1 - m) * feat + m * fix, m = fix > 0   (nlspn_model.py:159-160,169)
__global__ void nlspn_preserve_kernel(const float *__restrict__ in, const float *__restrict__ fix,
                                      float *__restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float f = __ldg(fix + i);
        const float m = f > 0.f ? 1.f : 0.f;
        out[i] = (1.f - m) * in[i] + m * f;
    }
}

__global__ void fuse_depth_kernel(const float *__restrict__ d1, const float *__restrict__ c1,
                                  const float *__restrict__ d2, const float *__restrict__ c2,
                                  float *__restrict__ d2c, float *__restrict__ pred, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v2 = fminf(fmaxf(d2[i], -1.f), 1.f);
        const float a = c1[i], b = c2[i], m = fmaxf(a, b);
        const float e1 = expf(a - m), e2 = expf(b - m), inv = 1.f / (e1 + e2);
        if (d2c) d2c[i] = v2;
        pred[i] = d1[i] * (e1 * inv) + v2 * (e2 * inv);
    }
}

}  // namespace
}  // namespace rdfc

using namespace rdfc;

extern "C" int rdfc_nlspn_affinity_forward(const float *guidance, const float *confidence, const float *conv_w,
                                           const float *conv_b, const float *aff_scale, int affinity, int conf_prop,
                                           float *offset, float *aff, int B, int H, int W, void *stream) {
    RDFC_REQUIRE(guidance && conv_w && conv_b && aff_scale && offset && aff, "NULL pointer argument");
    RDFC_REQUIRE(!conf_prop || confidence, "conf_prop requires a confidence map (nlspn_model.py:150-151)");
    RDFC_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "bad shape (%d,%d,%d)", B, H, W);
    RDFC_REQUIRE(affinity >= RDFC_AFF_AS && affinity <= RDFC_AFF_TGASS, "unknown affinity mode %d", affinity);
    dim3 grid(cdiv(W, TX), cdiv(H, TY), B), block(TX, TY);
    nlspn_affinity_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(guidance, confidence, conv_w, conv_b, aff_scale,
                                                                     affinity, conf_prop, offset, aff, H, W);
    RDFC_CHECK_LAUNCH("nlspn_affinity_kernel");
    return 0;
}

extern "C" int rdfc_nlspn_propagate_forward(const float *feat_init, const float *offset, const float *aff,
                                            const float *feat_fix, int preserve_input, float *out, float *scratch,
                                            float *inter, int B, int H, int W, int prop_time, int clamp_out,
                                            void *stream) {
    RDFC_REQUIRE(feat_init && offset && aff && out && scratch, "NULL pointer argument");
    RDFC_REQUIRE(!preserve_input || feat_fix, "preserve_input requires feat_fix (nlspn_model.py:157-158)");
    RDFC_REQUIRE(B > 0 && H > 0 && W > 0 && prop_time >= 0, "bad shape (%d,%d,%d) x %d", B, H, W, prop_time);
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)H * W;
    if (prop_time == 0) {
        RDFC_CUDA(cudaMemcpyAsync(out, feat_init, sizeof(float) * B * P, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    // L2 blocking: offset+aff of a group (27 planes, 25 read) should stay L2 resident across the iterations.
    const long long bytes_per_img = 27 * P * 4;
    // Measured on B200 (profiles/): walking the batch in L2-sized groups does NOT pay -- a 60 MB group still misses
    // L2 (21 % hit rate, LRU thrash) and small launches lose more to launch/tail effects -- so the default is one
    // launch per iteration over the whole batch.  RDFC_NLSPN_GROUP_MB re-enables grouping for experiments.
    long long budget = 1ll << 40;
    if (const char *e = getenv("RDFC_NLSPN_GROUP_MB")) budget = (long long)atoi(e) << 20;
    long long gsz = budget / bytes_per_img;
    if (gsz < 1) gsz = 1;
    if (gsz > 65535) gsz = 65535;
    // extra buffer for preserve_input (blend result); reuse: blend writes into the buffer not being read
    for (long long b0 = 0; b0 < B; b0 += gsz) {
        const int nb = (int)((B - b0) < gsz ? (B - b0) : gsz);
        dim3 grid(cdiv(W, TX), cdiv(H, TY), nb), block(TX, TY);
        const float *off_g = offset + b0 * 18 * P, *aff_g = aff + b0 * 9 * P;
        const float *fix_g = feat_fix ? feat_fix + b0 * P : nullptr;
        float *bufs[2] = {scratch + b0 * P, out + b0 * P};
        // choose ping-pong parity so that the last iteration lands in `out`
        const float *cur = feat_init + b0 * P;
        for (int t = 0; t < prop_time; ++t) {
            float *dst = bufs[(prop_time - 1 - t) % 2 == 0 ? 1 : 0];
            if (preserve_input) {
                // blend into dst's sibling is unsafe (cur may live there); blend in place needs cur writable:
                // iteration 0 reads feat_init (const) -> blend into the other buffer first.
                float *tmp = (dst == bufs[0]) ? bufs[1] : bufs[0];
                const int nblk = (int)min((long long)cdiv(nb * P, 256), (long long)sm_count() * 8);
                nlspn_preserve_kernel<<<nblk, 256, 0, st>>>(cur, fix_g, tmp, nb * P);
                RDFC_CHECK_LAUNCH("nlspn_preserve_kernel");
                cur = tmp;
            }
            float *it = inter ? inter + ((long long)t * B + b0) * P : nullptr;
            const bool clamp = clamp_out && t == prop_time - 1;
            // row-streaming kernel: needs 16-byte row granularity for the bulk copies
            const size_t stage_bytes = (size_t)25 * W * sizeof(float);
            const int nt = (W + 31) / 32 * 32;
            // many small CTAs per SM (1 stage each) hide the bulk-copy latency by interleaving; wide rows get 2 stages
            int stages = 2;
            if (const char *e = getenv("RDFC_NLSPN_STAGES")) stages = atoi(e);
            while (stages > 1 && stages * stage_bytes + 64 > 200 * 1024) --stages;
            const size_t smem = stages * stage_bytes + 64;
            // ring kernel: TMA row stream + shared-memory band gathers, one persistent CTA per SM
            {
                int halo_r = 8;
                if (const char *e = getenv("RDFC_NLSPN_HALO")) halo_r = atoi(e);
                const int nseg = (W + 31) / 32;
                int G = (H % 3 == 0 && nseg <= 10) ? 3 : ((H % 2 == 0 && nseg <= 15) ? 2 : 1);   // rows per stage ~ 30 warps of work
                if (const char *e = getenv("RDFC_NLSPN_G")) G = atoi(e);
                int ncw = G * nseg > 30 ? 30 : G * nseg;
                if (const char *e = getenv("RDFC_NLSPN_NCW")) ncw = atoi(e);
                const size_t stage_b = (size_t)G * 25 * W * 4;
                int S = stage_b >= 48 * 1024 ? 2 : 3;
                if (const char *e = getenv("RDFC_NLSPN_STAGES")) S = atoi(e);
                const long long total_rows = (long long)nb * H;
                long long nctas = sm_count();
                if (const char *e = getenv("RDFC_NLSPN_RING_CTAS")) nctas = atoll(e);
                long long rpc = (total_rows + nctas - 1) / nctas;
                rpc = (rpc + G - 1) / G * G;
                const long long left = 220 * 1024 - (long long)S * stage_b - 2 * S * 8 - 128;
                long long nr_max = left / ((W + 4) * 4) - 2 * halo_r - 3;
                if (nr_max > rpc) nr_max = rpc;
                nr_max = nr_max / G * G;
                const bool ring_ok = getenv("RDFC_NLSPN_RING") && !getenv("RDFC_NLSPN_SIMPLE") && !getenv("RDFC_NLSPN_ROWS") &&
                                     W % 4 == 0 && H % G == 0 && nr_max >= G && nr_max >= 4 && ncw >= 1 && ncw <= 30 &&
                                     ((uintptr_t)off_g % 16) == 0 && ((uintptr_t)aff_g % 16) == 0 && (P % 4) == 0;
                if (ring_ok) {
                    const size_t smem = (size_t)S * stage_b + (size_t)(nr_max + 2 * halo_r + 3) * (W + 4) * 4 + 2 * S * 8 + 128;
                    const int grid_r = (int)((total_rows + rpc - 1) / rpc);
                    if (clamp) {
                        RDFC_CUDA(cudaFuncSetAttribute(nlspn_prop_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                        nlspn_prop_ring_kernel<true><<<grid_r, (ncw + 1) * 32, smem, st>>>(cur, off_g, aff_g, dst, it, nb, H, W, (int)rpc, halo_r, G, S, (int)nr_max);
                    } else {
                        RDFC_CUDA(cudaFuncSetAttribute(nlspn_prop_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                        nlspn_prop_ring_kernel<false><<<grid_r, (ncw + 1) * 32, smem, st>>>(cur, off_g, aff_g, dst, it, nb, H, W, (int)rpc, halo_r, G, S, (int)nr_max);
                    }
                    RDFC_CHECK_LAUNCH("nlspn_prop_ring_kernel");
                    cur = dst;
                    continue;
                }
            }
            // band kernel: even W (8-byte plane vectors), tile must fit shared memory
            int halo = 6;      // rows of band halo: taps beyond it take the global path (measured at sigma = 2.2 px offsets: 8 -> 49.5, 6 -> 47.8, 4 -> 47.2 us per launch)
            if (const char *e = getenv("RDFC_NLSPN_HALO")) halo = atoi(e);
            const long long total_rows_b = (long long)nb * H;
            int band_pix = 2;                      // pixels per thread: 2 -> 3 CTAs / SM (measured best), 1 -> 5 CTAs / SM
            if (W % 2 != 0) band_pix = 1;
            if (const char *e = getenv("RDFC_NLSPN_PIX")) band_pix = atoi(e);
            const int band_per_sm = band_pix == 2 ? 3 : 5;
            int band_prefetch = 0;        // measured: L2 prefetch of the next item costs more LSU slots than it hides (1058 vs 957 us)
            if (const char *e = getenv("RDFC_NLSPN_PREFETCH")) band_prefetch = atoi(e);
            int band_pdl = 1;             // programmatic dependent launch between the iterations (see the kernel)
            if (const char *e = getenv("RDFC_NLSPN_PDL")) band_pdl = atoi(e) != 0;
            int band_pre = 0;             // items per thread whose offset / affinity lines are prefetched to L2 before the wait
            if (const char *e = getenv("RDFC_NLSPN_PRE")) band_pre = atoi(e);
            long long band_ctas = (long long)sm_count() * band_per_sm;
            if (const char *e = getenv("RDFC_NLSPN_BAND_CTAS")) band_ctas = atoll(e);
            if (band_ctas > total_rows_b) band_ctas = total_rows_b;
            const int band_rpc = (int)((total_rows_b + band_ctas - 1) / band_ctas);
            const size_t band_smem = (size_t)(band_rpc + 2 * halo + 3) * (W + 4) * sizeof(float);
            const size_t band_smem_max = (size_t)(220 * 1024) / band_per_sm;
            const bool band_ok = !getenv("RDFC_NLSPN_SIMPLE") && !getenv("RDFC_NLSPN_ROWS") && W % band_pix == 0 && band_smem <= band_smem_max &&
                                 ((uintptr_t)off_g % 8) == 0 && ((uintptr_t)aff_g % 8) == 0 && ((uintptr_t)dst % 8) == 0 &&
                                 (!it || ((uintptr_t)it % 8) == 0) && (P % 2) == 0;
            if (band_ok) {
                const int nctas = (int)((total_rows_b + band_rpc - 1) / band_rpc);
#define RDFC_BAND(CL, PX, MB)                                                                                                  \
    do {                                                                                                                       \
        RDFC_CUDA(cudaFuncSetAttribute(nlspn_prop_band_kernel<CL, PX, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)); \
        cudaLaunchConfig_t cfg = {};                                                                                       \
        cfg.gridDim = dim3(nctas); cfg.blockDim = dim3(BAND_THREADS); cfg.dynamicSmemBytes = band_smem; cfg.stream = st;       \
        cudaLaunchAttribute attr[1];                                                                                           \
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                       \
        attr[0].val.programmaticStreamSerializationAllowed = band_pdl;                                                         \
        cfg.attrs = attr; cfg.numAttrs = 1;                                                                                    \
        RDFC_CUDA(cudaLaunchKernelEx(&cfg, nlspn_prop_band_kernel<CL, PX, MB>, (const float *)cur, (const float *)off_g,       \
                                     (const float *)aff_g, (float *)dst, (float *)it, (int)nb, H, W, band_rpc, halo, band_prefetch, \
                                     band_pre));                                                                               \
    } while (0)
                if (band_pix == 2) { if (clamp) RDFC_BAND(true, 2, 3); else RDFC_BAND(false, 2, 3); }
                else { if (clamp) RDFC_BAND(true, 1, 5); else RDFC_BAND(false, 1, 5); }
#undef RDFC_BAND
                RDFC_CHECK_LAUNCH("nlspn_prop_band_kernel");
                cur = dst;
                continue;
            }
            const bool rows_ok = getenv("RDFC_NLSPN_ROWS") && W % 4 == 0 && W <= 1024 && smem <= 200 * 1024 &&
                                 ((uintptr_t)off_g % 16) == 0 && ((uintptr_t)aff_g % 16) == 0;
            if (rows_ok) {
                int per_sm = (int)((226 * 1024) / (smem + 1024));
                if (per_sm * nt > 2048) per_sm = 2048 / nt;
                if (nt <= 320 && per_sm > 3) per_sm = 3;       // register budget of the <320,3> instantiation (64 regs)
                if (per_sm < 1) per_sm = 1;
                if (const char *e = getenv("RDFC_NLSPN_CTAS_PER_SM")) per_sm = atoi(e);
                const long long total_rows = (long long)nb * H;
                long long nctas = (long long)sm_count() * per_sm;
                if (nctas > total_rows) nctas = total_rows;
                const int rpc = (int)((total_rows + nctas - 1) / nctas);
                nctas = (total_rows + rpc - 1) / rpc;
#define RDFC_ROWS(CL, MT, MB)                                                                                        \
    do {                                                                                                             \
        RDFC_CUDA(cudaFuncSetAttribute(nlspn_prop_rows_kernel<CL, MT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       200 * 1024));                                                                 \
        nlspn_prop_rows_kernel<CL, MT, MB><<<(int)nctas, nt, smem, st>>>(cur, off_g, aff_g, dst, it, nb, H, W, stages, rpc); \
    } while (0)
                if (nt <= 320) {
                    if (clamp) RDFC_ROWS(true, 320, 3); else RDFC_ROWS(false, 320, 3);
                } else {
                    if (clamp) RDFC_ROWS(true, 1024, 1); else RDFC_ROWS(false, 1024, 1);
                }
#undef RDFC_ROWS
                RDFC_CHECK_LAUNCH("nlspn_prop_rows_kernel");
            } else {
                if (clamp)
                    nlspn_prop_kernel<true><<<grid, block, 0, st>>>(cur, off_g, aff_g, dst, it, H, W);
                else
                    nlspn_prop_kernel<false><<<grid, block, 0, st>>>(cur, off_g, aff_g, dst, it, H, W);
                RDFC_CHECK_LAUNCH("nlspn_prop_kernel");
            }
            cur = dst;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the offset / affinity stage up to the conv output (training): given dL/d offset (B,18,H,W) and dL/d aff (B,9,H,W)
// it produces dL/d conv_offset_aff(guidance) (B,24,H,W), scatters dL/d confidence (the 1x1 DCN gathers of nlspn_model.py:96-119,
// offsets detached there) and accumulates dL/d aff_scale_const (TGASS).  The raw affinity channels are recomputed (8 of the 24
// conv outputs); the offsets are read back from the forward's output.  The conv's own backward (8 <-> 24 channels, 0.24 GFLOP per
// image) is left to the caller.
__global__ void __launch_bounds__(TX *TY) nlspn_affinity_bwd_kernel(const float *__restrict__ guidance, const float *__restrict__ confidence,
                                                                    const float *__restrict__ conv_w, const float *__restrict__ conv_b,
                                                                    const float *__restrict__ aff_scale, int affinity, int conf_prop,
                                                                    const float *__restrict__ offset, const float *__restrict__ g_offset,
                                                                    const float *__restrict__ g_aff, float *__restrict__ g_conv,
                                                                    float *__restrict__ g_conf, float *__restrict__ g_scale, int H, int W) {
    __shared__ __align__(16) float s_w[72 * 8];      // the 8 affinity channels: [(ci*9 + tap)][8]
    __shared__ float s_b[8];
    __shared__ float s_g[8][TY + 2][TX + 2];
    __shared__ float s_red[TX * TY / 32];
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int b = blockIdx.z, x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const long long P = (long long)H * W;
    for (int e = tid; e < 8 * 72; e += TX * TY) s_w[(e % 72) * 8 + e / 72] = conv_w[16 * 72 + e];
    if (tid < 8) s_b[tid] = conv_b[16 + tid];
    for (int e = tid; e < 8 * (TY + 2) * (TX + 2); e += TX * TY) {
        const int c = e / ((TY + 2) * (TX + 2)), r = e % ((TY + 2) * (TX + 2));
        const int yy = y0 + r / (TX + 2) - 1, xx = x0 + r % (TX + 2) - 1;
        s_g[c][r / (TX + 2)][r % (TX + 2)] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(guidance + ((long long)b * 8 + c) * P + (long long)yy * W + xx) : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const bool live = x < W && y < H;
    float gs = 0.f;                                  // this pixel's contribution to dL/d aff_scale_const
    if (live) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = s_b[j];
#pragma unroll
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float g = s_g[ci][threadIdx.y + t / 3][threadIdx.x + t % 3];
                const float4 *w4 = reinterpret_cast<const float4 *>(s_w + (ci * 9 + t) * 8);
                const float4 w0 = w4[0], w1 = w4[1];
                r[0] = fmaf(w0.x, g, r[0]); r[1] = fmaf(w0.y, g, r[1]); r[2] = fmaf(w0.z, g, r[2]); r[3] = fmaf(w0.w, g, r[3]);
                r[4] = fmaf(w1.x, g, r[4]); r[5] = fmaf(w1.y, g, r[5]); r[6] = fmaf(w1.z, g, r[6]); r[7] = fmaf(w1.w, g, r[7]);
            }
        const long long pix = (long long)y * W + x;
        const float *offp = offset + (long long)b * 18 * P + pix;
        const float *gop = g_offset + (long long)b * 18 * P + pix;
        const float *gap = g_aff + (long long)b * 9 * P + pix;
        float *gc = g_conv + (long long)b * 24 * P + pix;
        const float scale = __ldg(aff_scale);
        const float den = affinity == RDFC_AFF_TGASS ? scale + 1e-8f : scale;
        const float *conf = conf_prop ? confidence + (long long)b * P : nullptr;
        float t[8], v[8], c[8], u[8], dy[8], dx[8];
        float abs_sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = j < 4 ? j : j + 1;
            dy[j] = __ldg(offp + (long long)(2 * k) * P);
            dx[j] = __ldg(offp + (long long)(2 * k + 1) * P);
            t[j] = (affinity == RDFC_AFF_TC || affinity == RDFC_AFF_TGASS) ? tanhf(r[j]) : r[j];
            v[j] = (affinity == RDFC_AFF_TC || affinity == RDFC_AFF_TGASS) ? t[j] / den : r[j];
            c[j] = conf_prop ? bilinear(conf, H, W, (float)y + dy[j], (float)x + dx[j]) : 1.f;
            u[j] = v[j] * c[j];
            abs_sum += fabsf(u[j]);
        }
        abs_sum += 1e-4f;
        bool clamped = false;
        if ((affinity == RDFC_AFF_ASS || affinity == RDFC_AFF_TGASS) && abs_sum < 1.f) { abs_sum = 1.f; clamped = true; }
        const float g_ref = __ldg(gap + 4 * P);           // aff_ref = 1 - sum(aff)  (nlspn_model.py:132-136)
        float ga[8], dot = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = j < 4 ? j : j + 1;
            ga[j] = __ldg(gap + (long long)k * P) - g_ref;
            dot += ga[j] * u[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float gu = ga[j];
            if (affinity != RDFC_AFF_TC) {                // aff = u / S, S = sum|u| + 1e-4 (clamped from below: constant)
                const float sgn = u[j] > 0.f ? 1.f : (u[j] < 0.f ? -1.f : 0.f);
                gu = ga[j] / abs_sum - (clamped ? 0.f : dot / (abs_sum * abs_sum) * sgn);
            }
            const float gv = gu * c[j];
            if (conf_prop) {                              // dL/d confidence: transposed bilinear gather at (y + dy, x + dx)
                const float gcj = gu * v[j], sy = (float)y + dy[j], sx = (float)x + dx[j];
                if (gcj != 0.f && sy > -1.f && sx > -1.f && sy < (float)H && sx < (float)W) {
                    const float fy = floorf(sy), fx = floorf(sx);
                    const int yl = (int)fy, xl = (int)fx;
                    const float ly = sy - fy, lx = sx - fx;
                    float *gim = g_conf + (long long)b * P;
#pragma unroll
                    for (int aa = 0; aa < 2; ++aa)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            const int yy = yl + aa, xx = xl + cc;
                            if (yy >= 0 && yy <= H - 1 && xx >= 0 && xx <= W - 1)
                                atomicAdd(gim + (long long)yy * W + xx, (aa ? ly : 1.f - ly) * (cc ? lx : 1.f - lx) * gcj);
                        }
                }
            }
            float gr = gv;
            if (affinity == RDFC_AFF_TC || affinity == RDFC_AFF_TGASS) {
                gr = gv * (1.f - t[j] * t[j]) / den;
                gs -= gv * t[j] / (den * den);
            }
            gc[(long long)(16 + j) * P] = gr;
            const int k = j < 4 ? j : j + 1;
            gc[(long long)(2 * j) * P] = __ldg(gop + (long long)(2 * k) * P);          // conv channel 2j / 2j+1 = (dy, dx) of neighbour j
            gc[(long long)(2 * j + 1) * P] = __ldg(gop + (long long)(2 * k + 1) * P);
        }
    }
    if (g_scale) {                                    // block reduction, one atomic per CTA
        for (int o = 16; o > 0; o >>= 1) gs += __shfl_down_sync(0xffffffffu, gs, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = gs;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.f;
            for (int w = 0; w < TX * TY / 32; ++w) tot += s_red[w];
            if (tot != 0.f) atomicAdd(g_scale, tot);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward of the propagation (training).  One iteration is y = A x' with
//   (A x')[p] = sum_k aff_k[p] * bilinear(x', p + tap_k + offset_k[p]),   x' = preserve_input ? blend(x, fix) : x,
// the reference's 18 ModulatedDeformConvFunction calls with w = 1, b = 0 (nlspn_model.py:140-175).  Its backward
// (modulated_deform_conv_cuda.cu:124-280, kernels modulated_deform_im2col_cuda.cuh:197-328) is split in two phases:
//   1. g_{t-1} = A^T g_t for t = T..1: the col2im scatter (cuh:197-254); it needs no feature values, only offsets / affinities.
//   2. ONE pass over the pixels that accumulates, over all T iterations, grad_aff_k += g_t * bilinear(x'_{t-1}) and
//      grad_offset_k += g_t * aff_k * d bilinear / d(y, x) (col2im_coord, cuh:257-328) in registers and writes the 27 planes once,
//      instead of 18 read-modify-write rounds over them.
struct BwdPos { int yl, xl; float ly, lx; bool valid; };
__device__ __forceinline__ BwdPos bwd_pos(float y, float x, int H, int W) {
    BwdPos q;
    q.valid = y > -1.f && x > -1.f && y < (float)H && x < (float)W;      // cuh:180 / :228
    const float yf = floorf(y), xf = floorf(x);
    q.yl = (int)yf; q.xl = (int)xf; q.ly = y - yf; q.lx = x - xf;
    return q;
}

// g (raw) -> finalised in place: iterations below the last one add the gradient that arrived through list_feat and apply
// the (1 - mask_fix) factor of the blend (nlspn_model.py:169); then scattered into gprev (zero-initialised).
__global__ void __launch_bounds__(256) nlspn_bwd_scatter_kernel(float *__restrict__ g, const float *__restrict__ ginter,
                                                                const float *__restrict__ fix, const float *__restrict__ offset,
                                                                const float *__restrict__ aff, float *__restrict__ gprev,
                                                                int finalize, int H, int W) {
    const long long P = (long long)H * W;
    const int b = blockIdx.y;
    const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
    if (pix >= P) return;
    const long long o = (long long)b * P + pix;
    float gv = g[o];
    if (finalize) {
        if (fix && __ldg(fix + o) > 0.f) gv = 0.f;
        if (ginter) gv += __ldg(ginter + o);
        g[o] = gv;
    }
    if (gv == 0.f) return;
    const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
    const float *offp = offset + (long long)b * 18 * P + pix;
    const float *affp = aff + (long long)b * 9 * P + pix;
    float *gim = gprev + (long long)b * P;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float sy = (float)(y - 1 + k / 3) + __ldg(offp + (long long)(2 * k) * P);
        const float sx = (float)(x - 1 + k % 3) + __ldg(offp + (long long)(2 * k + 1) * P);
        const BwdPos q = bwd_pos(sy, sx, H, W);
        if (!q.valid) continue;
        const float top = gv * __ldg(affp + (long long)k * P);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int yy = q.yl + a, xx = q.xl + c;
                if (yy >= 0 && yy <= H - 1 && xx >= 0 && xx <= W - 1)
                    atomicAdd(gim + (long long)yy * W + xx, (a ? q.ly : 1.f - q.ly) * (c ? q.lx : 1.f - q.lx) * top);
            }
    }
}

// feats: x_0 = feat_init, x_t = inter[t-1]; G: finalised g_1..g_T at G + t * B * P.
template <bool kPreserve>
__global__ void __launch_bounds__(256) nlspn_bwd_coord_kernel(const float *__restrict__ feat_init, const float *__restrict__ inter,
                                                              const float *__restrict__ fix, const float *__restrict__ G,
                                                              const float *__restrict__ offset, const float *__restrict__ aff,
                                                              float *__restrict__ goff, float *__restrict__ gaff,
                                                              int B, int H, int W, int T) {
    const long long P = (long long)H * W, BP = (long long)B * P;
    const int b = blockIdx.y;
    const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
    if (pix >= P) return;
    const long long o = (long long)b * P + pix;
    const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
    const float *offp = offset + (long long)b * 18 * P + pix;
    const float *affp = aff + (long long)b * 9 * P + pix;
    const float *fx = kPreserve ? fix + (long long)b * P : nullptr;
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        const float sy = (float)(y - 1 + k / 3) + __ldg(offp + (long long)(2 * k) * P);
        const float sx = (float)(x - 1 + k % 3) + __ldg(offp + (long long)(2 * k + 1) * P);
        const float m = __ldg(affp + (long long)k * P);
        float a_m = 0.f, a_y = 0.f, a_x = 0.f;
        // cuh:84-125 / :295-322: outside (-1, H) x (-1, W) every derivative is zero
        if (!(sy <= -1.f || sy >= (float)H || sx <= -1.f || sx >= (float)W)) {
            const BwdPos q = bwd_pos(sy, sx, H, W);
            const int yh = q.yl + 1, xh = q.xl + 1;
            const bool ok1 = q.yl >= 0 && q.xl >= 0, ok2 = q.yl >= 0 && xh <= W - 1, ok3 = yh <= H - 1 && q.xl >= 0,
                       ok4 = yh <= H - 1 && xh <= W - 1;
            const long long i1 = (long long)q.yl * W + q.xl, i2 = i1 + 1, i3 = i1 + W, i4 = i3 + 1;
            const float hy = 1.f - q.ly, hx = 1.f - q.lx;
            for (int t = 1; t <= T; ++t) {
                const float gv = __ldg(G + (long long)t * BP + o);
                if (gv == 0.f) continue;
                const float *im = (t == 1 ? feat_init : inter + (long long)(t - 2) * BP) + (long long)b * P;
                float v1 = ok1 ? __ldg(im + i1) : 0.f, v2 = ok2 ? __ldg(im + i2) : 0.f, v3 = ok3 ? __ldg(im + i3) : 0.f,
                      v4 = ok4 ? __ldg(im + i4) : 0.f;
                if (kPreserve) {        // x' = fix where fix > 0 (nlspn_model.py:159-160,169)
                    const float f1 = ok1 ? __ldg(fx + i1) : 0.f, f2 = ok2 ? __ldg(fx + i2) : 0.f, f3 = ok3 ? __ldg(fx + i3) : 0.f,
                                f4 = ok4 ? __ldg(fx + i4) : 0.f;
                    v1 = f1 > 0.f ? f1 : v1; v2 = f2 > 0.f ? f2 : v2; v3 = f3 > 0.f ? f3 : v3; v4 = f4 > 0.f ? f4 : v4;
                }
                const float val = hy * hx * v1 + hy * q.lx * v2 + q.ly * hx * v3 + q.ly * q.lx * v4;
                a_m += gv * val;
                a_y += gv * (hx * (v3 - v1) + q.lx * (v4 - v2));
                a_x += gv * (hy * (v2 - v1) + q.ly * (v4 - v3));
            }
        }
        gaff[(long long)b * 9 * P + (long long)k * P + pix] = a_m;
        goff[(long long)b * 18 * P + (long long)(2 * k) * P + pix] = a_y * m;
        goff[(long long)b * 18 * P + (long long)(2 * k + 1) * P + pix] = a_x * m;
    }
}

__global__ void nlspn_bwd_final_kernel(const float *__restrict__ g0, const float *__restrict__ fix, float *__restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (fix && __ldg(fix + i) > 0.f) ? 0.f : g0[i];
}

extern "C" int rdfc_nlspn_affinity_backward(const float *guidance, const float *confidence, const float *conv_w, const float *conv_b,
                                            const float *aff_scale, int affinity, int conf_prop, const float *offset,
                                            const float *grad_offset, const float *grad_aff, float *grad_conv, float *grad_confidence,
                                            float *grad_aff_scale, int B, int H, int W, void *stream) {
    RDFC_REQUIRE(guidance && conv_w && conv_b && aff_scale && offset && grad_offset && grad_aff && grad_conv, "NULL pointer argument");
    RDFC_REQUIRE(!conf_prop || (confidence && grad_confidence), "conf_prop requires confidence and grad_confidence");
    RDFC_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535, "bad shape (%d,%d,%d)", B, H, W);
    RDFC_REQUIRE(affinity >= RDFC_AFF_AS && affinity <= RDFC_AFF_TGASS, "unknown affinity mode %d", affinity);
    cudaStream_t st = (cudaStream_t)stream;
    if (conf_prop) RDFC_CUDA(cudaMemsetAsync(grad_confidence, 0, sizeof(float) * (size_t)B * H * W, st));
    if (grad_aff_scale) RDFC_CUDA(cudaMemsetAsync(grad_aff_scale, 0, sizeof(float), st));
    dim3 grid(cdiv(W, TX), cdiv(H, TY), B), block(TX, TY);
    nlspn_affinity_bwd_kernel<<<grid, block, 0, st>>>(guidance, confidence, conv_w, conv_b, aff_scale, affinity, conf_prop, offset,
                                                      grad_offset, grad_aff, grad_conv, grad_confidence, grad_aff_scale, H, W);
    RDFC_CHECK_LAUNCH("nlspn_affinity_bwd_kernel");
    return 0;
}

extern "C" int rdfc_nlspn_propagate_backward(const float *grad_out, const float *grad_inter, const float *feat_init,
                                             const float *inter, const float *offset, const float *aff, const float *feat_fix,
                                             int preserve_input, float *grad_feat_init, float *grad_offset, float *grad_aff,
                                             float *scratch, int B, int H, int W, int prop_time, void *stream) {
    RDFC_REQUIRE(grad_out && feat_init && offset && aff && grad_feat_init && grad_offset && grad_aff && scratch, "NULL pointer argument");
    RDFC_REQUIRE(!preserve_input || feat_fix, "preserve_input requires feat_fix (nlspn_model.py:157-158)");
    RDFC_REQUIRE(B > 0 && H > 0 && W > 0 && prop_time >= 0 && B <= 65535, "bad shape (%d,%d,%d) x %d", B, H, W, prop_time);
    RDFC_REQUIRE(prop_time <= 1 || inter, "the backward needs the forward's intermediate results (inter)");
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)H * W, BP = (long long)B * P;
    const int T = prop_time;
    const float *fix = preserve_input ? feat_fix : nullptr;
    if (T == 0) {
        RDFC_CUDA(cudaMemcpyAsync(grad_feat_init, grad_out, sizeof(float) * BP, cudaMemcpyDeviceToDevice, st));
        RDFC_CUDA(cudaMemsetAsync(grad_offset, 0, sizeof(float) * 18 * BP, st));
        RDFC_CUDA(cudaMemsetAsync(grad_aff, 0, sizeof(float) * 9 * BP, st));
        return 0;
    }
    // scratch = G[0..T]: G[T] = grad_out (the caller's grad_inter[T-1], if any, is added by the first scatter), the rest zero
    float *G = scratch;
    RDFC_CUDA(cudaMemsetAsync(G, 0, sizeof(float) * T * BP, st));
    RDFC_CUDA(cudaMemcpyAsync(G + (long long)T * BP, grad_out, sizeof(float) * BP, cudaMemcpyDeviceToDevice, st));
    dim3 grid((unsigned)cdiv(P, 256), (unsigned)B);
    for (int t = T; t >= 1; --t) {
        const float *gi = grad_inter ? grad_inter + (long long)(t - 1) * BP : nullptr;
        // the last iteration's output is not blended: only list_feat's gradient is added there
        nlspn_bwd_scatter_kernel<<<grid, 256, 0, st>>>(G + (long long)t * BP, gi, t < T ? fix : nullptr, offset, aff,
                                                       G + (long long)(t - 1) * BP, (t < T || gi) ? 1 : 0, H, W);
        RDFC_CHECK_LAUNCH("nlspn_bwd_scatter_kernel");
    }
    const int nblk = (int)min((long long)cdiv(BP, 256), (long long)sm_count() * 8);
    nlspn_bwd_final_kernel<<<nblk, 256, 0, st>>>(G, fix, grad_feat_init, BP);
    RDFC_CHECK_LAUNCH("nlspn_bwd_final_kernel");
    if (fix)
        nlspn_bwd_coord_kernel<true><<<grid, 256, 0, st>>>(feat_init, inter, fix, G, offset, aff, grad_offset, grad_aff, B, H, W, T);
    else
        nlspn_bwd_coord_kernel<false><<<grid, 256, 0, st>>>(feat_init, inter, fix, G, offset, aff, grad_offset, grad_aff, B, H, W, T);
    RDFC_CHECK_LAUNCH("nlspn_bwd_coord_kernel");
    return 0;
}

extern "C" int rdfc_fuse_depth_forward(const float *d1, const float *c1, const float *d2, const float *c2,
                                       float *d2_clamped, float *pred, size_t n, void *stream) {
    RDFC_REQUIRE(d1 && c1 && d2 && c2 && pred, "NULL pointer argument");
    if (n == 0) return 0;
    const int nblk = (int)min((long long)cdiv((long long)n, 256), (long long)sm_count() * 8);
    fuse_depth_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(d1, c1, d2, c2, d2_clamped, pred, n);
    RDFC_CHECK_LAUNCH("fuse_depth_kernel");
    return 0;
}
