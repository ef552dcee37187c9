// rdfc_conv_forward: validates the descriptor and dispatches to the tcgen05 (bf16) or CUDA-core (fp32) kernel.
#include "common.cuh"

namespace rdfc {
int conv_simt_forward(const rdfc_conv_desc *d, cudaStream_t st);
int conv_umma_forward(const rdfc_conv_desc *d, cudaStream_t st, const rdfc_heads_desc *heads);
int conv_umma_read_dbg(long long *host, int n);
}  // namespace rdfc

using namespace rdfc;

extern "C" int rdfc_conv_forward(const rdfc_conv_desc *d, void *stream) {
    RDFC_REQUIRE(d != nullptr, "descriptor is NULL");
    RDFC_REQUIRE(d->in.ptr && d->out.ptr && d->weight, "conv: input / output / weight must not be NULL");
    RDFC_REQUIRE(d->B > 0 && d->Hi > 0 && d->Wi > 0 && d->Ho > 0 && d->Wo > 0 && d->in.C > 0 && d->out.C > 0,
                 "conv: empty dimension");
    RDFC_REQUIRE(d->kh > 0 && d->kw > 0 && d->kh * d->kw <= 9 && d->stride > 0 && d->pad >= 0, "conv: bad kernel geometry");
    if (d->transposed) {
        RDFC_REQUIRE(d->Ho <= 2 * d->Hi && d->Wo <= 2 * d->Wi, "conv: transposed output larger than 2x input");
    } else {
        const int ho = (d->Hi + 2 * d->pad - d->kh) / d->stride + 1, wo = (d->Wi + 2 * d->pad - d->kw) / d->stride + 1;
        RDFC_REQUIRE(d->Ho == ho && d->Wo == wo, "conv: output size (%d,%d) != expected (%d,%d)", d->Ho, d->Wo, ho, wo);
    }
    RDFC_REQUIRE(d->in.nchw || d->in.pix_stride >= d->in.C, "conv: input pixel stride smaller than its channel count");
    RDFC_REQUIRE(d->out.nchw || d->out.pix_stride >= d->out.C, "conv: output pixel stride smaller than its channel count");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->path == RDFC_PATH_UMMA_BF16) return conv_umma_forward(d, st, nullptr);
    if (d->path == RDFC_PATH_SIMT_F32) return conv_simt_forward(d, st);
    return fail(RDFC_ERR_INVALID, "conv: unknown path %d", d->path);
}

extern "C" int rdfc_heads_forward(const rdfc_heads_desc *h, void *stream) {
    RDFC_REQUIRE(h != nullptr && h->in.ptr && h->weight && h->shift, "heads: NULL argument");
    RDFC_REQUIRE(h->ncols >= 1 && h->ncols <= 16, "heads: ncols must be 1..16");
    RDFC_REQUIRE(h->B > 0 && h->H > 0 && h->W > 0, "heads: empty dimension");
    for (int q = 0; q < h->ncols; ++q) RDFC_REQUIRE(h->out[q] != nullptr, "heads: output plane %d is NULL", q);
    rdfc_conv_desc d{};
    d.B = h->B; d.Hi = d.Ho = h->H; d.Wi = d.Wo = h->W;
    d.kh = d.kw = 3; d.stride = 1; d.pad = 1; d.act = RDFC_ACT_NONE; d.path = RDFC_PATH_UMMA_BF16;
    d.in = h->in;
    d.out.ptr = h->out[0]; d.out.dtype = RDFC_F32; d.out.C = 16; d.out.pix_stride = 16;
    d.weight = h->weight; d.scale = nullptr; d.shift = h->shift;
    return conv_umma_forward(&d, (cudaStream_t)stream, h);
}

// development aid (not part of the public ABI): role timers of the last conv_umma launch under RDFC_UMMA_DBG=1
extern "C" int rdfc_dev_umma_timers(long long *host, int n) { return conv_umma_read_dbg(host, n); }
