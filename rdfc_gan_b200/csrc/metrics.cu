// Evaluation glue on the device: the eight depth metrics of RDFGANMetric
// (lib/metrics/rdf_gan_metric.py:59-151) and the de-normalisation of Eval.inference (lib/evaluator/evaluator.py:27-29) as ONE
// pass over (pred, gt): per image nine sums, reduced deterministically (fixed chunking, tree reduction, no atomics).
// HBM-bound: 8 bytes per pixel (+1 with an evaluate mask).
#include "common.cuh"

namespace rdfc {
namespace {

constexpr int NT = 256, NSUM = 9;       // count, sum d^2, sum |d|, sum dinv^2, sum |dinv|, sum rel, #(ratio < 1.25^k) k = 1..3
constexpr int CHUNK = 16384;            // pixels per CTA

__global__ void __launch_bounds__(NT) depth_metric_partial_kernel(const float *__restrict__ pred, const float *__restrict__ gt,
                                                                  const unsigned char *__restrict__ emask, float std, float mean,
                                                                  float t_valid, double *__restrict__ partial, long long n) {
    const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
    const long long base = (long long)b * n, lo = (long long)chunk * CHUNK, hi = lo + CHUNK < n ? lo + CHUNK : n;
    double acc[NSUM];
#pragma unroll
    for (int i = 0; i < NSUM; ++i) acc[i] = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += NT) {
        // fp32 per-pixel arithmetic exactly as the reference's torch ops (evaluator.py:28-29, rdf_gan_metric.py:73-131)
        const float p = __fadd_rn(__fmul_rn(__ldg(pred + base + i), std), mean), g = __fadd_rn(__fmul_rn(__ldg(gt + base + i), std), mean);   // two roundings, like torch
        if (!(g > t_valid) || (emask && !emask[base + i])) continue;
        const float pinv = p <= t_valid ? 0.f : 1.0f / (p + 1e-8f), ginv = 1.0f / (g + 1e-8f);     // g > t_valid here
        const float d = p - g, da = fabsf(d), di = pinv - ginv;
        const float ratio = fmaxf(g / (p + 1e-8f), p / (g + 1e-8f));
        acc[0] += 1.0;
        acc[1] += (double)(d * d);
        acc[2] += (double)da;
        acc[3] += (double)(di * di);
        acc[4] += (double)fabsf(di);
        acc[5] += (double)(da / (g + 1e-8f));
        acc[6] += ratio < 1.25f ? 1.0 : 0.0;
        acc[7] += ratio < 1.5625f ? 1.0 : 0.0;
        acc[8] += ratio < 1.953125f ? 1.0 : 0.0;
    }
    __shared__ double red[NSUM][NT / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NSUM; ++i) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < NSUM) {
        double v = 0.0;
        for (int w = 0; w < NT / 32; ++w) v += red[threadIdx.x][w];
        partial[((long long)b * nchunk + chunk) * NSUM + threadIdx.x] = v;
    }
}

__global__ void depth_metric_final_kernel(const double *__restrict__ partial, double *__restrict__ sums, int nchunk) {
    const int b = blockIdx.x;
    if (threadIdx.x < NSUM) {
        double v = 0.0;
        for (int c = 0; c < nchunk; ++c) v += partial[((long long)b * nchunk + c) * NSUM + threadIdx.x];
        sums[b * NSUM + threadIdx.x] = v;
    }
}

}  // namespace
}  // namespace rdfc

using namespace rdfc;

extern "C" int rdfc_depth_metric_nchunk(long long n) { return (int)((n + CHUNK - 1) / CHUNK); }

extern "C" int rdfc_depth_metric_sums(const float *pred, const float *gt, const unsigned char *evaluate_mask, float std, float mean,
                                      float t_valid, double *sums, double *partial, int B, long long n, void *stream) {
    RDFC_REQUIRE(pred && gt && sums && partial, "NULL pointer argument");
    RDFC_REQUIRE(B > 0 && B <= 65535 && n > 0, "bad shape (%d, %lld)", B, n);
    const int nchunk = rdfc_depth_metric_nchunk(n);
    depth_metric_partial_kernel<<<dim3(nchunk, B), NT, 0, (cudaStream_t)stream>>>(pred, gt, evaluate_mask, std, mean, t_valid, partial, n);
    RDFC_CHECK_LAUNCH("depth_metric_partial_kernel");
    depth_metric_final_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(partial, sums, nchunk);
    RDFC_CHECK_LAUNCH("depth_metric_final_kernel");
    return 0;
}
