// Glue kernels of the ESANet guidance network (RDF-GAN's global_guidance_module,
// F/lib/models/segmentator/esa_net/esa_net_one_modality.py:11-194, F = /root/reference/RDF-GAN), bf16 NHWC.
//
// The GEMM-shaped layers of ESANet (ResNet BasicBlocks, 1x1 skip / context convs, decoder 3x3 and factorised 3x1 / 1x3 convs,
// conv_out) run on conv_umma_kernel; what is left are HBM-bound layers with a handful of input channels or no contraction at
// all: the 7x7 stride-2 stem, max pooling, squeeze-and-excitation, the pyramid-pooling module's average pools and nearest
// up-sampling, and the decoder's "learned" x2 up-sampling (nearest + depth-wise 3x3, + skip add).
#include "common.cuh"

namespace rdfc {
namespace {

// ---- encoder.conv1 + bn1 + ReLU: k x k (k <= 7) stride-s conv from fp32 NCHW (Cin <= 4) to bf16 NHWC (Cout = 64) ----------
// block = 64 threads (one per output channel) x 4 pixels; the filter bank sits in shared memory as [tap][ci][cout].
__global__ void __launch_bounds__(256) first_conv_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ scale,
                                                         const float *__restrict__ shift, __nv_bfloat16 *__restrict__ out, int out_stride, int B,
                                                         int Cin, int Hi, int Wi, int Ho, int Wo, int k, int s, int pad, int Cout, int relu) {
    extern __shared__ float sw[];                       // [k*k*Cin][Cout]
    const int ntaps = k * k * Cin;
    for (int e = threadIdx.x; e < ntaps * Cout; e += blockDim.x) {
        const int co = e % Cout, t = e / Cout;           // t = (ky*k + kx)*Cin + ci ; torch layout w[co][ci][ky][kx]
        const int ci = t % Cin, kk = t / Cin;
        sw[e] = w[((long long)co * Cin + ci) * k * k + kk];
    }
    __syncthreads();
    const int co = threadIdx.x % Cout, sub = threadIdx.x / Cout, npb = blockDim.x / Cout;
    const long long npix = (long long)B * Ho * Wo;
    for (long long p = (long long)blockIdx.x * npb + sub; p < npix; p += (long long)gridDim.x * npb) {
        const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), b = (int)(p / ((long long)Wo * Ho));
        float acc = 0.f;
        for (int ky = 0; ky < k; ++ky) {
            const int iy = oy * s - pad + ky;
            if (iy < 0 || iy >= Hi) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int ix = ox * s - pad + kx;
                if (ix < 0 || ix >= Wi) continue;
                for (int ci = 0; ci < Cin; ++ci)
                    acc = fmaf(__ldg(x + (((long long)b * Cin + ci) * Hi + iy) * Wi + ix), sw[((ky * k + kx) * Cin + ci) * Cout + co], acc);
            }
        }
        float y = fmaf(acc, scale[co], shift[co]);
        if (relu) y = fmaxf(y, 0.f);
        out[p * out_stride + co] = __float2bfloat16_rn(y);
    }
}

// ---- max pooling 3x3 / stride 2 / pad 1 over bf16 NHWC, 8 channels per thread -------------------------------------------
__global__ void __launch_bounds__(256) maxpool_kernel(const __nv_bfloat16 *__restrict__ x, int x_stride, __nv_bfloat16 *__restrict__ out,
                                                      int out_stride, int B, int C, int Hi, int Wi, int Ho, int Wo) {
    const int cgs = C >> 3;
    const long long total = (long long)B * Ho * Wo * cgs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % cgs);
        const long long p = i / cgs;
        const int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), b = (int)(p / ((long long)Wo * Ho));
        float m[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) m[q] = -3.0e38f;
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= Hi) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= Wi) continue;
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(x + (((long long)b * Hi + iy) * Wi + ix) * x_stride + cg * 8));
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&wv[q]));
                    m[2 * q] = fmaxf(m[2 * q], f.x);
                    m[2 * q + 1] = fmaxf(m[2 * q + 1], f.y);
                }
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * q], m[2 * q + 1]);
            o[q] = *reinterpret_cast<const uint32_t *>(&h);
        }
        *reinterpret_cast<uint4 *>(out + p * out_stride + cg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---- squeeze-and-excitation weights: w[b, :] = sigmoid(W2 relu(W1 mean[b, :] + b1) + b2), one CTA per image ----------------
__global__ void __launch_bounds__(256) se_weights_kernel(const float *__restrict__ mean, const float *__restrict__ w1, const float *__restrict__ b1,
                                                         const float *__restrict__ w2, const float *__restrict__ b2, float *__restrict__ out, int C,
                                                         int R) {
    extern __shared__ float sm[];                       // mean[C], hidden[R]
    float *m = sm, *h = sm + C;
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) m[c] = mean[(long long)b * C + c];
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float a = b1[r];
        for (int c = 0; c < C; ++c) a = fmaf(w1[(long long)r * C + c], m[c], a);
        h[r] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = b2[c];
        for (int r = 0; r < R; ++r) a = fmaf(w2[(long long)c * R + r], h[r], a);
        out[(long long)b * C + c] = 1.f / (1.f + expf(-a));
    }
}

// ---- adaptive average pooling to bins x bins (torch's window rule: [floor(i*H/n), ceil((i+1)*H/n))) ------------------------
__global__ void __launch_bounds__(256) adaptive_avgpool_kernel(const __nv_bfloat16 *__restrict__ x, int x_stride, __nv_bfloat16 *__restrict__ out,
                                                               int out_stride, int B, int C, int H, int W, int bins) {
    const long long total = (long long)B * bins * bins * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long p = i / C;
        const int bx = (int)(p % bins), by = (int)((p / bins) % bins), b = (int)(p / ((long long)bins * bins));
        const int y0 = (by * H) / bins, y1 = ((by + 1) * H + bins - 1) / bins, x0 = (bx * W) / bins, x1 = ((bx + 1) * W + bins - 1) / bins;
        float s = 0.f;
        for (int y = y0; y < y1; ++y)
            for (int xx = x0; xx < x1; ++xx) s += __bfloat162float(x[(((long long)b * H + y) * W + xx) * x_stride + c]);
        out[p * out_stride + c] = __float2bfloat16_rn(s / (float)((y1 - y0) * (x1 - x0)));
    }
}

// ---- nearest up-sampling of a (hs x ws) map into a channel slice of a (H x W) NHWC buffer (torch 'nearest': floor(dst*in/out)) --
__global__ void __launch_bounds__(256) upsample_nearest_kernel(const __nv_bfloat16 *__restrict__ x, int x_stride, __nv_bfloat16 *__restrict__ out,
                                                               int out_stride, int B, int C, int hs, int ws, int H, int W) {
    const int cgs = C >> 3;
    const long long total = (long long)B * H * W * cgs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % cgs);
        const long long p = i / cgs;
        const int ox = (int)(p % W), oy = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
        const int sy = min((int)floorf(oy * ((float)hs / H)), hs - 1), sx = min((int)floorf(ox * ((float)ws / W)), ws - 1);
        *reinterpret_cast<uint4 *>(out + p * out_stride + cg * 8) =
            __ldg(reinterpret_cast<const uint4 *>(x + (((long long)b * hs + sy) * ws + sx) * x_stride + cg * 8));
    }
}

// ---- the decoder's learned up-sampling (Upsample, decoder.py:137-191, mode 'learned-3x3-zeropad'): nearest resize to (Ho, Wo),
// depth-wise 3x3 conv with zero padding + bias, optional skip add.  Output: bf16 NHWC, or fp32 NCHW planes (the network's result).
template <bool kNCHW>
__global__ void __launch_bounds__(256) upsample_dw_kernel(const __nv_bfloat16 *__restrict__ x, int x_stride, const float *__restrict__ w,
                                                          const float *__restrict__ bias, const __nv_bfloat16 *__restrict__ skip, int skip_stride,
                                                          void *__restrict__ out, int out_stride, int B, int C, int Hi, int Wi, int Ho, int Wo) {
    const long long total = (long long)B * Ho * Wo * C;
    const float fy = (float)Hi / Ho, fx = (float)Wi / Wo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // NHWC output: channel fastest; NCHW output: x fastest (coalesced stores either way)
        int c, ox, oy, b;
        if (kNCHW) {
            ox = (int)(i % Wo); oy = (int)((i / Wo) % Ho); c = (int)((i / ((long long)Wo * Ho)) % C); b = (int)(i / ((long long)Wo * Ho * C));
        } else {
            c = (int)(i % C); ox = (int)((i / C) % Wo); oy = (int)((i / ((long long)C * Wo)) % Ho); b = (int)(i / ((long long)C * Wo * Ho));
        }
        float acc = bias[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int uy = oy - 1 + ky;                 // position in the up-sampled map
            if (uy < 0 || uy >= Ho) continue;
            const int sy = min((int)floorf(uy * fy), Hi - 1);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ux = ox - 1 + kx;
                if (ux < 0 || ux >= Wo) continue;
                const int sx = min((int)floorf(ux * fx), Wi - 1);
                acc = fmaf(w[c * 9 + ky * 3 + kx], __bfloat162float(x[(((long long)b * Hi + sy) * Wi + sx) * x_stride + c]), acc);
            }
        }
        const long long p = ((long long)b * Ho + oy) * Wo + ox;
        if (skip) acc += __bfloat162float(skip[p * skip_stride + c]);
        if (kNCHW) reinterpret_cast<float *>(out)[(((long long)b * C + c) * Ho + oy) * Wo + ox] = acc;
        else reinterpret_cast<__nv_bfloat16 *>(out)[p * out_stride + c] = __float2bfloat16_rn(acc);
    }
}

int grid_for(long long total) { return (int)min((long long)cdiv(total, 256), (long long)sm_count() * 16); }

}  // namespace
}  // namespace rdfc

using namespace rdfc;

static int bf16_view_ok(const rdfc_view *v, const char *what) {
    RDFC_REQUIRE(v && v->ptr && v->dtype == RDFC_BF16 && !v->nchw && v->C % 8 == 0 && v->pix_stride % 8 == 0 && ((uintptr_t)v->ptr % 16) == 0,
                 "%s: 16-byte aligned bf16 NHWC view with C %% 8 == 0 expected", what);
    return 0;
}

extern "C" int rdfc_first_conv_forward(const float *x_nchw, int B, int Cin, int Hi, int Wi, const float *weight, int k, int stride, int pad,
                                       const float *scale, const float *shift, int relu, const rdfc_view *out, void *stream) {
    RDFC_REQUIRE(x_nchw && weight && scale && shift && out, "first conv: NULL argument");
    if (int rc = bf16_view_ok(out, "first conv out")) return rc;
    RDFC_REQUIRE(B > 0 && Cin >= 1 && Cin <= 4 && k >= 1 && k <= 7 && stride >= 1 && out->C <= 256 && 256 % out->C == 0,
                 "first conv: Cin <= 4, k <= 7, Cout a divisor of 256");
    const int Ho = (Hi + 2 * pad - k) / stride + 1, Wo = (Wi + 2 * pad - k) / stride + 1;
    const size_t smem = (size_t)k * k * Cin * out->C * sizeof(float);
    RDFC_REQUIRE(smem <= 48 * 1024, "first conv: filter bank exceeds 48 KB of shared memory");
    const long long npix = (long long)B * Ho * Wo;
    const int npb = 256 / out->C;
    const int grid = (int)min((long long)cdiv(npix, npb), (long long)sm_count() * 8);
    first_conv_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x_nchw, weight, scale, shift, (__nv_bfloat16 *)out->ptr, out->pix_stride, B, Cin, Hi,
                                                                 Wi, Ho, Wo, k, stride, pad, out->C, relu);
    RDFC_CHECK_LAUNCH("first_conv_kernel");
    return 0;
}

extern "C" int rdfc_maxpool3x3s2_forward(const rdfc_view *x, const rdfc_view *out, int B, int Hi, int Wi, void *stream) {
    if (int rc = bf16_view_ok(x, "maxpool x")) return rc;
    if (int rc = bf16_view_ok(out, "maxpool out")) return rc;
    RDFC_REQUIRE(x->C == out->C && B > 0 && Hi > 0 && Wi > 0, "maxpool: channel mismatch / empty");
    const int Ho = (Hi - 1) / 2 + 1, Wo = (Wi - 1) / 2 + 1;
    maxpool_kernel<<<grid_for((long long)B * Ho * Wo * (x->C / 8)), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x->ptr, x->pix_stride,
                                                                                                   (__nv_bfloat16 *)out->ptr, out->pix_stride, B,
                                                                                                   x->C, Hi, Wi, Ho, Wo);
    RDFC_CHECK_LAUNCH("maxpool_kernel");
    return 0;
}

extern "C" int rdfc_se_weights(const float *mean, const float *w1, const float *b1, const float *w2, const float *b2, float *out, int B, int C,
                               int R, void *stream) {
    RDFC_REQUIRE(mean && w1 && b1 && w2 && b2 && out && B > 0 && C > 0 && R > 0 && (size_t)(C + R) * 4 <= 48 * 1024, "SE weights: bad argument");
    se_weights_kernel<<<B, 256, (size_t)(C + R) * sizeof(float), (cudaStream_t)stream>>>(mean, w1, b1, w2, b2, out, C, R);
    RDFC_CHECK_LAUNCH("se_weights_kernel");
    return 0;
}

extern "C" int rdfc_adaptive_avgpool_forward(const rdfc_view *x, const rdfc_view *out, int B, int H, int W, int bins, void *stream) {
    if (int rc = bf16_view_ok(x, "adaptive avgpool x")) return rc;
    if (int rc = bf16_view_ok(out, "adaptive avgpool out")) return rc;
    RDFC_REQUIRE(x->C == out->C && B > 0 && bins > 0 && H > 0 && W > 0, "adaptive avgpool: bad shape");
    adaptive_avgpool_kernel<<<grid_for((long long)B * bins * bins * x->C), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)x->ptr, x->pix_stride, (__nv_bfloat16 *)out->ptr, out->pix_stride, B, x->C, H, W, bins);
    RDFC_CHECK_LAUNCH("adaptive_avgpool_kernel");
    return 0;
}

extern "C" int rdfc_upsample_nearest_forward(const rdfc_view *x, const rdfc_view *out, int B, int hs, int ws, int H, int W, void *stream) {
    if (int rc = bf16_view_ok(x, "upsample x")) return rc;
    if (int rc = bf16_view_ok(out, "upsample out")) return rc;
    RDFC_REQUIRE(x->C == out->C && B > 0 && hs > 0 && ws > 0 && H > 0 && W > 0, "upsample nearest: bad shape");
    upsample_nearest_kernel<<<grid_for((long long)B * H * W * (x->C / 8)), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)x->ptr, x->pix_stride, (__nv_bfloat16 *)out->ptr, out->pix_stride, B, x->C, hs, ws, H, W);
    RDFC_CHECK_LAUNCH("upsample_nearest_kernel");
    return 0;
}

extern "C" int rdfc_upsample_dw_forward(const rdfc_view *x, const float *weight, const float *bias, const rdfc_view *skip, const rdfc_view *out,
                                        float *out_nchw, int B, int Hi, int Wi, int Ho, int Wo, void *stream) {
    if (int rc = bf16_view_ok(x, "upsample dw x")) return rc;
    RDFC_REQUIRE(weight && bias && B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "upsample dw: bad argument");
    RDFC_REQUIRE((out && out->ptr) != (out_nchw != nullptr), "upsample dw: exactly one of out (bf16 NHWC) / out_nchw (fp32 NCHW)");
    const __nv_bfloat16 *sk = nullptr;
    int sk_stride = 0;
    if (skip && skip->ptr) {
        if (int rc = bf16_view_ok(skip, "upsample dw skip")) return rc;
        RDFC_REQUIRE(skip->C == x->C, "upsample dw: skip channel mismatch");
        sk = (const __nv_bfloat16 *)skip->ptr; sk_stride = skip->pix_stride;
    }
    const long long total = (long long)B * Ho * Wo * x->C;
    if (out_nchw) {
        upsample_dw_kernel<true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x->ptr, x->pix_stride, weight, bias, sk,
                                                                                    sk_stride, out_nchw, 0, B, x->C, Hi, Wi, Ho, Wo);
    } else {
        if (int rc = bf16_view_ok(out, "upsample dw out")) return rc;
        RDFC_REQUIRE(out->C == x->C, "upsample dw: output channel mismatch");
        upsample_dw_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x->ptr, x->pix_stride, weight, bias, sk,
                                                                                     sk_stride, out->ptr, out->pix_stride, B, x->C, Hi, Wi, Ho, Wo);
    }
    RDFC_CHECK_LAUNCH("upsample_dw_kernel");
    return 0;
}
