// Instance-norm statistics and the RGB<-depth fusion epilogues (W-AdaIN / AdaIN / IN) on NHWC views.
// Reference: model_utils.py:53-90 (AdaptiveInstanceNorm), :92-116 (AdaIN, unbiased variance), :119-129 (IN).
#include "common.cuh"

namespace rdfc {
namespace {

constexpr int CHUNK_PIX = 256;

// Pass 1: per (b, pixel chunk, c) shifted sums -> (mean, M2) partials.
// grid (nchunk, B, ceil(C/64)), block 256 = 8 channel groups (8 channels = one 16/32-byte vector) x 32 pixel lanes;
// each thread strides over the chunk's pixels, then the 32 lanes are reduced through shared memory.
template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float *p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16 *p, float (&v)[8]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[q]));
        v[2 * q] = f.x;
        v[2 * q + 1] = f.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) in_partial_kernel(const T *__restrict__ x, int C, int stride, int npix,
                                                         float *__restrict__ partial, int nchunk) {
    __shared__ float red[32][8][17];
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int cg = threadIdx.x & 7, lane = threadIdx.x >> 3;            // 8 channel groups x 32 pixel lanes
    const int c0 = blockIdx.z * 64 + cg * 8;
    const int p0 = chunk * CHUNK_PIX, p1 = min(npix, p0 + CHUNK_PIX), n = p1 - p0;
    float s[8], ss[8], pivot[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = ss[q] = pivot[q] = 0.f;
    const bool active = c0 < C;                                          // C is a multiple of 8 (checked on the host)
    if (active) {
        const T *base = x + ((long long)b * npix + p0) * stride + c0;
        load8<T>(base, pivot);
        for (int p = lane; p < n; p += 32) {
            float v[8];
            load8<T>(base + (long long)p * stride, v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float d = v[q] - pivot[q];
                s[q] += d;
                ss[q] = fmaf(d, d, ss[q]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        red[lane][cg][q] = s[q];
        red[lane][cg][8 + q] = ss[q];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int g2 = threadIdx.x >> 3, q = threadIdx.x & 7, c = blockIdx.z * 64 + g2 * 8 + q;
        if (c < C) {
            float ts = 0.f, tss = 0.f;
            for (int l = 0; l < 32; ++l) {
                ts += red[l][g2][q];
                tss += red[l][g2][8 + q];
            }
            // every lane used the same pivot (pixel p0 of the chunk)
            const float pv = ldf(x + ((long long)b * npix + p0) * stride + c);
            const float mean = pv + ts / n, m2 = fmaxf(tss - ts * ts / n, 0.f);
            float *o = partial + (((long long)b * nchunk + chunk) * C + c) * 2;
            o[0] = mean;
            o[1] = m2;
        }
    }
}

// Pass 2: Chan combination in chunk order (deterministic).  grid (B), threads over c.
__global__ void __launch_bounds__(256) in_final_kernel(const float *__restrict__ partial, int C, int npix, int nchunk,
                                                       float eps, int unbiased, int want_std,
                                                       float *__restrict__ mean_o, float *__restrict__ rstd_o) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean = 0.f, m2 = 0.f, n = 0.f;
        // the combination is a serial chain, the loads are not: fetch eight chunks' partials at once (one L2 round trip per
        // eight chunks instead of one per chunk), then combine them in chunk order as before
        for (int k0 = 0; k0 < nchunk; k0 += 8) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < nchunk) v[u] = __ldg(reinterpret_cast<const float2 *>(partial + (((long long)b * nchunk + k0 + u) * C + c) * 2));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < nchunk) {
                    const int k = k0 + u;
                    const float nb = (float)min(CHUNK_PIX, npix - k * CHUNK_PIX);
                    const float delta = v[u].x - mean, nn = n + nb;
                    mean += delta * nb / nn;
                    m2 += v[u].y + delta * delta * n * nb / nn;
                    n = nn;
                }
        }
        const float var = m2 / (unbiased ? (n - 1.f) : n) + eps;
        mean_o[(long long)b * C + c] = mean;
        rstd_o[(long long)b * C + c] = want_std ? sqrtf(var) : rsqrtf(var);
    }
}

struct V {
    const void *ptr;
    int stride;
};

template <typename TX, typename TG, typename TO>
__global__ void __launch_bounds__(256) wadain_kernel(V x, V gb, V gw, V bw, const float *__restrict__ mean,
                                                     const float *__restrict__ rstd, void *out, int out_stride,
                                                     int C, long long npix_total, int npix) {
    const long long total = npix_total * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C;
        const int c = (int)(i % C), b = (int)(p / npix);
        const float xv = ldf((const TX *)x.ptr + p * x.stride + c);
        float gamma = ldf((const TG *)gb.ptr + p * gb.stride + c);
        float beta = ldf((const TG *)gb.ptr + p * gb.stride + C + c);
        if (gw.ptr) gamma *= ldf((const TG *)gw.ptr + p * gw.stride + c);
        if (bw.ptr) beta *= ldf((const TG *)bw.ptr + p * bw.stride + c);
        const float nrm = (xv - mean[(long long)b * C + c]) * rstd[(long long)b * C + c];
        stf((TO *)out + p * out_stride + c, gamma * nrm + beta);
    }
}

// out = (x - m0) * r0 * s1 + m1   (s1/m1 NULL -> plain normalise)
template <typename TX, typename TO>
__global__ void __launch_bounds__(256) affine_norm_kernel(V x, const float *__restrict__ m0,
                                                          const float *__restrict__ r0, const float *__restrict__ m1,
                                                          const float *__restrict__ s1, int divide, void *out,
                                                          int out_stride, int C, long long npix_total, int npix) {
    const long long total = npix_total * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C;
        const int c = (int)(i % C), b = (int)(p / npix);
        const long long bc = (long long)b * C + c;
        float v = ldf((const TX *)x.ptr + p * x.stride + c) - m0[bc];
        v = divide ? v / r0[bc] : v * r0[bc];
        if (s1) v = v * s1[bc] + m1[bc];
        stf((TO *)out + p * out_stride + c, v);
    }
}

int grid_for(long long total) { return (int)min((long long)cdiv(total, 256), (long long)sm_count() * 16); }

}  // namespace
}  // namespace rdfc

using namespace rdfc;

extern "C" int rdfc_instnorm_nchunk(int npix) { return cdiv(npix, CHUNK_PIX); }

extern "C" int rdfc_instnorm_stats(const rdfc_view *x, int B, int H, int W, float eps, int unbiased, int want_std,
                                   float *partial, float *mean, float *rstd, void *stream) {
    RDFC_REQUIRE(x && x->ptr && partial && mean && rstd, "NULL pointer argument");
    RDFC_REQUIRE(!x->nchw, "instnorm_stats expects an NHWC view");
    RDFC_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "bad shape");
    const int npix = H * W, nchunk = cdiv(npix, CHUNK_PIX);
    cudaStream_t st = (cudaStream_t)stream;
    RDFC_REQUIRE(x->C % 8 == 0 && x->pix_stride % 8 == 0 && ((uintptr_t)x->ptr % 32) == 0,
                 "instnorm_stats: channels / pixel stride must be multiples of 8 and the view 32-byte aligned");
    dim3 grid(nchunk, B, cdiv(x->C, 64));
    if (x->dtype == RDFC_F32)
        in_partial_kernel<float><<<grid, 256, 0, st>>>((const float *)x->ptr, x->C, x->pix_stride, npix, partial, nchunk);
    else if (x->dtype == RDFC_BF16)
        in_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x->ptr, x->C, x->pix_stride, npix,
                                                               partial, nchunk);
    else
        return fail(RDFC_ERR_UNSUPPORTED, "instnorm_stats: dtype %d", x->dtype);
    RDFC_CHECK_LAUNCH("in_partial_kernel");
    in_final_kernel<<<B, 256, 0, st>>>(partial, x->C, npix, nchunk, eps, unbiased, want_std, mean, rstd);
    RDFC_CHECK_LAUNCH("in_final_kernel");
    return 0;
}

extern "C" int rdfc_wadain_apply(const rdfc_view *x, const rdfc_view *gb, const rdfc_view *gw, const rdfc_view *bw,
                                 const float *mean, const float *rstd, const rdfc_view *out, int B, int H, int W,
                                 void *stream) {
    RDFC_REQUIRE(x && gb && out && x->ptr && gb->ptr && out->ptr && mean && rstd, "NULL pointer argument");
    RDFC_REQUIRE(gb->C == 2 * x->C && out->C == x->C, "wadain: channel mismatch (x %d, gb %d, out %d)", x->C, gb->C, out->C);
    RDFC_REQUIRE(x->dtype == out->dtype && gb->dtype == x->dtype, "wadain: x, gb and out must share a dtype");
    const V vx{x->ptr, x->pix_stride}, vg{gb->ptr, gb->pix_stride};
    const V vgw{gw ? gw->ptr : nullptr, gw ? gw->pix_stride : 0}, vbw{bw ? bw->ptr : nullptr, bw ? bw->pix_stride : 0};
    RDFC_REQUIRE((!gw || gw->dtype == x->dtype) && (!bw || bw->dtype == x->dtype), "wadain: weighting dtype mismatch");
    const long long np = (long long)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(np * x->C);
    if (x->dtype == RDFC_F32)
        wadain_kernel<float, float, float><<<g, 256, 0, st>>>(vx, vg, vgw, vbw, mean, rstd, out->ptr, out->pix_stride, x->C, np, H * W);
    else if (x->dtype == RDFC_BF16)
        wadain_kernel<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>(vx, vg, vgw, vbw, mean, rstd, out->ptr,
                                                                                     out->pix_stride, x->C, np, H * W);
    else
        return fail(RDFC_ERR_UNSUPPORTED, "wadain: dtype %d", x->dtype);
    RDFC_CHECK_LAUNCH("wadain_kernel");
    return 0;
}

static int affine_norm(const rdfc_view *x, const float *m0, const float *r0, const float *m1, const float *s1,
                       int divide, const rdfc_view *out, int B, int H, int W, void *stream) {
    RDFC_REQUIRE(x && out && x->ptr && out->ptr && m0 && r0, "NULL pointer argument");
    RDFC_REQUIRE(out->C == x->C && x->dtype == out->dtype, "norm apply: view mismatch");
    const V vx{x->ptr, x->pix_stride};
    const long long np = (long long)B * H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(np * x->C);
    if (x->dtype == RDFC_F32)
        affine_norm_kernel<float, float><<<g, 256, 0, st>>>(vx, m0, r0, m1, s1, divide, out->ptr, out->pix_stride, x->C, np, H * W);
    else if (x->dtype == RDFC_BF16)
        affine_norm_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>(vx, m0, r0, m1, s1, divide, out->ptr,
                                                                           out->pix_stride, x->C, np, H * W);
    else
        return fail(RDFC_ERR_UNSUPPORTED, "norm apply: dtype %d", x->dtype);
    RDFC_CHECK_LAUNCH("affine_norm_kernel");
    return 0;
}

extern "C" int rdfc_adain_apply(const rdfc_view *x, const float *cmean, const float *cstd, const float *smean,
                                const float *sstd, const rdfc_view *out, int B, int H, int W, void *stream) {
    RDFC_REQUIRE(smean && sstd, "NULL style statistics");
    return affine_norm(x, cmean, cstd, smean, sstd, 1, out, B, H, W, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 NCHW stem inputs -> one bf16 NHWC tensor [in0's C0 channels | in1's channel | zeros] for the tensor-core stems of
// generators whose stem input has many channels (RDF-GAN: 40-channel guidance, rdf_gan_generator.py:235-245).
// A block transposes 64 pixels x all channels through shared memory: coalesced plane reads, 16-byte NHWC writes.
namespace rdfc {
namespace {
__global__ void __launch_bounds__(256) pack_stem_kernel(const float *__restrict__ in0, int C0, const float *__restrict__ in1,
                                                        __nv_bfloat16 *__restrict__ out, int Cpad, long long HW, long long total) {
    extern __shared__ float ps_tile[];               // [Cpad][65]
    const long long p0 = (long long)blockIdx.x * 64;
    const int nsrc = C0 + (in1 ? 1 : 0);
    for (int e = threadIdx.x; e < Cpad * 64; e += 256) {
        const int c = e >> 6, px = e & 63;
        const long long p = p0 + px;
        float v = 0.f;
        if (c < nsrc && p < total) {
            const long long b = p / HW, q = p - b * HW;
            v = c < C0 ? __ldg(in0 + (b * C0 + c) * HW + q) : __ldg(in1 + b * HW + q);
        }
        ps_tile[c * 65 + px] = v;
    }
    __syncthreads();
    const int nchunk = Cpad >> 3;
    for (int e = threadIdx.x; e < 64 * nchunk; e += 256) {
        const int px = e / nchunk, ch = e - px * nchunk;
        const long long p = p0 + px;
        if (p >= total) continue;
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(ps_tile[(ch * 8 + 2 * k) * 65 + px], ps_tile[(ch * 8 + 2 * k + 1) * 65 + px]);
            w[k] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        *reinterpret_cast<uint4 *>(out + p * Cpad + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
}  // namespace
}  // namespace rdfc

extern "C" int rdfc_pack_stem_input(const float *in0, int C0, const float *in1, void *out_bf16, int Cpad, int B, int H, int W,
                                    void *stream) {
    RDFC_REQUIRE(in0 && out_bf16, "pack_stem_input: NULL argument");
    RDFC_REQUIRE(C0 >= 1 && Cpad % 8 == 0 && Cpad >= C0 + (in1 ? 1 : 0) && Cpad <= 128, "pack_stem_input: bad channel counts (%d -> %d)", C0, Cpad);
    RDFC_REQUIRE(B > 0 && H > 0 && W > 0 && ((uintptr_t)out_bf16 % 16) == 0, "pack_stem_input: bad shape / alignment");
    const long long HW = (long long)H * W, total = HW * B;
    const size_t smem = (size_t)Cpad * 65 * sizeof(float);
    rdfc::pack_stem_kernel<<<(unsigned)((total + 63) / 64), 256, smem, (cudaStream_t)stream>>>(in0, C0, in1, (__nv_bfloat16 *)out_bf16, Cpad,
                                                                                              HW, total);
    RDFC_CHECK_LAUNCH("pack_stem_kernel");
    return 0;
}

extern "C" int rdfc_norm_apply(const rdfc_view *x, const float *mean, const float *rstd, const rdfc_view *out, int B,
                               int H, int W, void *stream) {
    return affine_norm(x, mean, rstd, nullptr, nullptr, 0, out, B, H, W, stream);
}
