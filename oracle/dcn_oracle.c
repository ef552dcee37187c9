/*
 * oracle/dcn_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double accumulation, OpenMP over independent
 * outputs) of the reference's deformable-convolution boundary.  It exists so
 * that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs can check (and time) the CUDA path against the
 * reference's arithmetic.  Nothing under rdfc_gan_b200/ may import, link or
 * call it.
 *
 * Reference (citations are into
 * /root/reference/RDFC-GAN/lib/models/generator/rdf_generator/nlspn/deformconv/):
 *   forward   src/cuda/modulated_deform_im2col_cuda.cuh:25-54  (bilinear, corner zeroing)
 *             src/cuda/modulated_deform_im2col_cuda.cuh:128-194 (sampling grid, validity rule)
 *             src/cuda/modulated_deform_conv_cuda.cu:75-118     (output size, grouped contraction + bias)
 *   backward  src/cuda/modulated_deform_conv_cuda.cu:211-275    (gcol, grad_weight, grad_bias)
 *             src/cuda/modulated_deform_im2col_cuda.cuh:57-81,197-254   (grad_input scatter)
 *             src/cuda/modulated_deform_im2col_cuda.cuh:84-125,257-328  (grad_offset / grad_mask)
 *   DCN v1    src/cuda/deform_im2col_cuda.cuh:127,192,249 -- identical with mask == 1
 *             (pass mask == NULL here).
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against the
 * fixtures under tests/golden/ that tests/golden/make_golden.py produced by
 * importing the reference's own Python (its DCN extension has no CPU kernel,
 * so the reference Function runs on torchvision.ops.deform_conv2d there, as
 * BASELINE.json prescribes), and against the reference's known-answer
 * properties from deformconv/test.py.
 *
 * Layouts (all contiguous, as the reference asserts):
 *   input  (B, Cin, H, W)            weight (Cout, Cin/group, kh, kw)   bias (Cout)
 *   offset (B, dg*2*kh*kw, Ho, Wo)   channel 2k = dy, 2k+1 = dx of tap k = i*kw + j
 *   mask   (B, dg*kh*kw, Ho, Wo)     output (B, Cout, Ho, Wo)
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int B, Cin, H, W, Cout, kh, kw, sh, sw, ph, pw, dh, dw, group, dg;
} dcn_shape;

static int out_h(const dcn_shape *s) { return (s->H + 2 * s->ph - (s->dh * (s->kh - 1) + 1)) / s->sh + 1; }
static int out_w(const dcn_shape *s) { return (s->W + 2 * s->pw - (s->dw * (s->kw - 1) + 1)) / s->sw + 1; }

/* modulated_deform_im2col_cuda.cuh:25-54 */
#define DEF_BILINEAR(T, NAME)                                                                   \
    static double NAME(const T *im, int H, int W, double y, double x) {                         \
        int yl = (int)floor(y), xl = (int)floor(x), yh = yl + 1, xh = xl + 1;                   \
        double ly = y - yl, lx = x - xl, hy = 1.0 - ly, hx = 1.0 - lx;                          \
        double v1 = (yl >= 0 && xl >= 0) ? (double)im[(size_t)yl * W + xl] : 0.0;               \
        double v2 = (yl >= 0 && xh <= W - 1) ? (double)im[(size_t)yl * W + xh] : 0.0;           \
        double v3 = (yh <= H - 1 && xl >= 0) ? (double)im[(size_t)yh * W + xl] : 0.0;           \
        double v4 = (yh <= H - 1 && xh <= W - 1) ? (double)im[(size_t)yh * W + xh] : 0.0;       \
        return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;                       \
    }

/* modulated_deform_im2col_cuda.cuh:84-125 ; dir 0 = d/dy, 1 = d/dx */
#define DEF_COORD(T, NAME)                                                                      \
    static double NAME(const T *im, int H, int W, double y, double x, int dir) {                \
        if (y <= -1 || y >= H || x <= -1 || x >= W) return 0.0;                                 \
        int yl = (int)floor(y), xl = (int)floor(x), yh = yl + 1, xh = xl + 1;                   \
        double w = 0.0;                                                                         \
        if (dir == 0) {                                                                         \
            if (yl >= 0 && xl >= 0) w += -1.0 * (xl + 1 - x) * im[(size_t)yl * W + xl];         \
            if (yl >= 0 && xh <= W - 1) w += -1.0 * (x - xl) * im[(size_t)yl * W + xh];         \
            if (yh <= H - 1 && xl >= 0) w += (xl + 1 - x) * im[(size_t)yh * W + xl];            \
            if (yh <= H - 1 && xh <= W - 1) w += (x - xl) * im[(size_t)yh * W + xh];            \
        } else {                                                                                \
            if (yl >= 0 && xl >= 0) w += -1.0 * (yl + 1 - y) * im[(size_t)yl * W + xl];         \
            if (yl >= 0 && xh <= W - 1) w += (yl + 1 - y) * im[(size_t)yl * W + xh];            \
            if (yh <= H - 1 && xl >= 0) w += -1.0 * (y - yl) * im[(size_t)yh * W + xl];         \
            if (yh <= H - 1 && xh <= W - 1) w += (y - yl) * im[(size_t)yh * W + xh];            \
        }                                                                                       \
        return w;                                                                               \
    }

#define DEF_ALL(T, SFX)                                                                                     \
    DEF_BILINEAR(T, bilinear_##SFX)                                                                         \
    DEF_COORD(T, coord_##SFX)                                                                               \
                                                                                                            \
    /* forward: modulated_deform_im2col_cuda.cuh:128-194 + modulated_deform_conv_cuda.cu:106-118 */         \
    int orc_dcn_forward_##SFX(const T *input, const T *weight, const T *bias, const T *offset,              \
                              const T *mask /* NULL => DCN v1 */, T *output, const dcn_shape *s) {          \
        const int Ho = out_h(s), Wo = out_w(s), K = s->kh * s->kw;                                          \
        const int cpg = s->Cin / s->group, opg = s->Cout / s->group, cpdg = s->Cin / s->dg;                 \
        if (s->Cin % s->group || s->Cout % s->group || s->Cin % s->dg) return 1;                            \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                            \
        for (int b = 0; b < s->B; ++b)                                                                      \
            for (int ho = 0; ho < Ho; ++ho) {                                                               \
                double *col = (double *)malloc(sizeof(double) * (size_t)s->Cin * K);                        \
                for (int wo = 0; wo < Wo; ++wo) {                                                           \
                    for (int ci = 0; ci < s->Cin; ++ci) {                                                   \
                        const int gd = ci / cpdg;                                                           \
                        const T *im = input + ((size_t)b * s->Cin + ci) * s->H * s->W;                      \
                        const T *off = offset + ((size_t)b * s->dg + gd) * 2 * K * Ho * Wo;                 \
                        const T *msk = mask ? mask + ((size_t)b * s->dg + gd) * K * Ho * Wo : NULL;         \
                        for (int k = 0; k < K; ++k) {                                                       \
                            const int i = k / s->kw, j = k % s->kw;                                         \
                            const size_t pix = (size_t)ho * Wo + wo;                                        \
                            const double dy = off[(size_t)(2 * k) * Ho * Wo + pix];                         \
                            const double dx = off[(size_t)(2 * k + 1) * Ho * Wo + pix];                     \
                            const double m = msk ? (double)msk[(size_t)k * Ho * Wo + pix] : 1.0;            \
                            const double y = ho * s->sh - s->ph + i * s->dh + dy;                           \
                            const double x = wo * s->sw - s->pw + j * s->dw + dx;                           \
                            double val = 0.0;                                                               \
                            if (y > -1 && x > -1 && y < s->H && x < s->W)                                   \
                                val = bilinear_##SFX(im, s->H, s->W, y, x);                                 \
                            col[(size_t)ci * K + k] = val * m;                                              \
                        }                                                                                   \
                    }                                                                                       \
                    for (int co = 0; co < s->Cout; ++co) {                                                  \
                        const int g = co / opg;                                                             \
                        double acc = bias ? (double)bias[co] : 0.0;                                         \
                        const T *wrow = weight + (size_t)co * cpg * K;                                      \
                        const double *c = col + (size_t)g * cpg * K;                                        \
                        for (int q = 0; q < cpg * K; ++q) acc += (double)wrow[q] * c[q];                    \
                        output[(((size_t)b * s->Cout + co) * Ho + ho) * Wo + wo] = (T)acc;                  \
                    }                                                                                       \
                }                                                                                           \
                free(col);                                                                                  \
            }                                                                                               \
        return 0;                                                                                           \
    }                                                                                                       \
                                                                                                            \
    /* backward: modulated_deform_conv_cuda.cu:211-275 and the col2im / col2im_coord kernels */             \
    int orc_dcn_backward_##SFX(const T *input, const T *weight, const T *offset, const T *mask,             \
                               const T *grad_output, T *grad_input, T *grad_offset, T *grad_mask,           \
                               T *grad_weight, T *grad_bias, const dcn_shape *s) {                          \
        const int Ho = out_h(s), Wo = out_w(s), K = s->kh * s->kw;                                          \
        const int cpg = s->Cin / s->group, opg = s->Cout / s->group, cpdg = s->Cin / s->dg;                 \
        const size_t n_in = (size_t)s->B * s->Cin * s->H * s->W;                                            \
        const size_t n_off = (size_t)s->B * s->dg * 2 * K * Ho * Wo;                                        \
        const size_t n_w = (size_t)s->Cout * cpg * K;                                                       \
        double *gin = (double *)calloc(n_in, sizeof(double));                                               \
        double *goff = (double *)calloc(n_off, sizeof(double));                                             \
        double *gmsk = (double *)calloc(n_off / 2, sizeof(double));                                         \
        double *gw = (double *)calloc(n_w, sizeof(double));                                                 \
        double *gb = (double *)calloc((size_t)s->Cout, sizeof(double));                                     \
        /* serial over outputs: scatter-adds into gin/gw are order dependent only in rounding (double) */   \
        for (int b = 0; b < s->B; ++b)                                                                      \
            for (int ho = 0; ho < Ho; ++ho)                                                                 \
                for (int wo = 0; wo < Wo; ++wo) {                                                           \
                    const size_t pix = (size_t)ho * Wo + wo;                                                \
                    for (int co = 0; co < s->Cout; ++co)                                                    \
                        gb[co] += (double)grad_output[((size_t)b * s->Cout + co) * Ho * Wo + pix];          \
                    for (int ci = 0; ci < s->Cin; ++ci) {                                                   \
                        const int g = ci / cpg, cil = ci % cpg, gd = ci / cpdg;                             \
                        const T *im = input + ((size_t)b * s->Cin + ci) * s->H * s->W;                      \
                        double *gim = gin + ((size_t)b * s->Cin + ci) * s->H * s->W;                        \
                        const size_t obase = ((size_t)b * s->dg + gd) * 2 * K * Ho * Wo;                    \
                        const size_t mbase = ((size_t)b * s->dg + gd) * K * Ho * Wo;                        \
                        for (int k = 0; k < K; ++k) {                                                       \
                            const int i = k / s->kw, j = k % s->kw;                                         \
                            const double dy = offset[obase + (size_t)(2 * k) * Ho * Wo + pix];              \
                            const double dx = offset[obase + (size_t)(2 * k + 1) * Ho * Wo + pix];          \
                            const double m = mask ? (double)mask[mbase + (size_t)k * Ho * Wo + pix] : 1.0;  \
                            const double y = ho * s->sh - s->ph + i * s->dh + dy;                           \
                            const double x = wo * s->sw - s->pw + j * s->dw + dx;                           \
                            const int valid = (y > -1 && x > -1 && y < s->H && x < s->W);                   \
                            const double val = valid ? bilinear_##SFX(im, s->H, s->W, y, x) : 0.0;          \
                            /* gcol = W_g^T grad_out  (modulated_deform_conv_cuda.cu:217-222) */            \
                            double gcol = 0.0;                                                              \
                            for (int col = 0; col < opg; ++col) {                                           \
                                const int co = g * opg + col;                                               \
                                const double go =                                                           \
                                    grad_output[((size_t)b * s->Cout + co) * Ho * Wo + pix];                \
                                gcol += (double)weight[((size_t)co * cpg + cil) * K + k] * go;              \
                                /* grad_weight += grad_out . col^T  (:265-270), col = val*m */              \
                                gw[((size_t)co * cpg + cil) * K + k] += go * val * m;                       \
                            }                                                                               \
                            /* col2im_coord (:257-328) */                                                   \
                            gmsk[mbase + (size_t)k * Ho * Wo + pix] += gcol * val;                          \
                            goff[obase + (size_t)(2 * k) * Ho * Wo + pix] +=                                \
                                gcol * m * coord_##SFX(im, s->H, s->W, y, x, 0);                            \
                            goff[obase + (size_t)(2 * k + 1) * Ho * Wo + pix] +=                            \
                                gcol * m * coord_##SFX(im, s->H, s->W, y, x, 1);                            \
                            /* col2im (:197-254, weights :57-81) */                                         \
                            if (valid) {                                                                    \
                                const int yl = (int)floor(y), xl = (int)floor(x);                           \
                                const double ly = y - yl, lx = x - xl;                                      \
                                const double top = gcol * m;                                                \
                                for (int a = 0; a < 2; ++a)                                                 \
                                    for (int c = 0; c < 2; ++c) {                                           \
                                        const int yy = yl + a, xx = xl + c;                                 \
                                        if (yy < 0 || yy > s->H - 1 || xx < 0 || xx > s->W - 1) continue;   \
                                        const double wgt = (a ? ly : 1.0 - ly) * (c ? lx : 1.0 - lx);       \
                                        gim[(size_t)yy * s->W + xx] += wgt * top;                           \
                                    }                                                                       \
                            }                                                                               \
                        }                                                                                   \
                    }                                                                                       \
                }                                                                                           \
        if (grad_input) for (size_t q = 0; q < n_in; ++q) grad_input[q] = (T)gin[q];                        \
        if (grad_offset) for (size_t q = 0; q < n_off; ++q) grad_offset[q] = (T)goff[q];                    \
        if (grad_mask) for (size_t q = 0; q < n_off / 2; ++q) grad_mask[q] = (T)gmsk[q];                    \
        if (grad_weight) for (size_t q = 0; q < n_w; ++q) grad_weight[q] = (T)gw[q];                        \
        if (grad_bias) for (int q = 0; q < s->Cout; ++q) grad_bias[q] = (T)gb[q];                           \
        free(gin); free(goff); free(gmsk); free(gw); free(gb);                                              \
        return 0;                                                                                           \
    }

DEF_ALL(float, f32)
DEF_ALL(double, f64)

/*
 * NLSPN propagation loop, restating nlspn/nlspn_model.py:140-144,157-173 for
 * channels_f == 1: `prop_time` applications of the 3x3 (k_f x k_f) modulated
 * deformable convolution with weight == 1, bias == 0, stride 1, pad (k_f-1)/2,
 * optionally re-imposing the sparse input before every step (preserve_input,
 * mask = feat_fix > 0).  Written as a dedicated loop (fp32 in/out per step,
 * exactly as the reference materialises fp32 tensors between steps) so that
 * bench.py's CPU leg is not dominated by call overhead.
 */
int orc_nlspn_propagate_f32(const float *feat_init, const float *offset, const float *aff,
                            const float *feat_fix /* may be NULL */, int preserve_input,
                            float *feat_out, float *scratch /* B*H*W */, int B, int H, int W,
                            int k_f, int prop_time) {
    const int K = k_f * k_f, pad = (k_f - 1) / 2;
    const size_t P = (size_t)H * W;
    float *cur = scratch, *nxt = feat_out;
    memcpy(cur, feat_init, sizeof(float) * B * P);
    for (int t = 0; t < prop_time; ++t) {
        if (preserve_input && feat_fix) {
            _Pragma("omp parallel for schedule(static)")
            for (long q = 0; q < (long)(B * P); ++q) {
                const float mfix = feat_fix[q] > 0.0f ? 1.0f : 0.0f; /* nlspn_model.py:159-160 */
                cur[q] = (1.0f - mfix) * cur[q] + mfix * feat_fix[q]; /* :169 */
            }
        }
        _Pragma("omp parallel for collapse(2) schedule(static)")
        for (int b = 0; b < B; ++b)
            for (int h = 0; h < H; ++h) {
                const float *im = cur + (size_t)b * P;
                const float *off = offset + (size_t)b * 2 * K * P;
                const float *a = aff + (size_t)b * K * P;
                for (int w = 0; w < W; ++w) {
                    const size_t pix = (size_t)h * W + w;
                    double acc = 0.0;
                    for (int k = 0; k < K; ++k) {
                        const double y = h - pad + k / k_f + (double)off[(size_t)(2 * k) * P + pix];
                        const double x = w - pad + k % k_f + (double)off[(size_t)(2 * k + 1) * P + pix];
                        if (y > -1 && x > -1 && y < H && x < W)
                            acc += (double)a[(size_t)k * P + pix] * bilinear_f32(im, H, W, y, x);
                    }
                    nxt[(size_t)b * P + pix] = (float)acc;
                }
            }
        float *tmp = cur; cur = nxt; nxt = tmp;
    }
    if (cur != feat_out) memcpy(feat_out, cur, sizeof(float) * B * P);
    return 0;
}
