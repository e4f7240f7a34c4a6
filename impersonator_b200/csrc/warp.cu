// Bilinear flow warp on NCHW fp32 tensors: ImpersonatorGenerator.transform / stn / resize_trans
// (networks/generator.py:303-320) and the image-level warp of models/imitator.py:259.
//
// HBM-bound gather: one thread owns one output pixel, derives the (optionally resized) flow
// once, then sweeps its slice of channels so that a warp writes 32 consecutive x of one plane
// (coalesced stores; the 4 taps of neighbouring pixels share 32B sectors through L1/L2).
#include "common.cuh"
#include "sample.cuh"

namespace {

__global__ void __launch_bounds__(256) k_warp_nchw(
        const float* __restrict__ x, int src_batch, int C, int h, int w,
        const float* __restrict__ T, int B, int th, int tw, int align_corners,
        float* __restrict__ out, int accumulate, int c_per_block)
{
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.z;
    if (pix >= h * w) return;
    const int y = pix / w, xo = pix % w;
    float gx, gy;
    lwb::flow_at(T + (size_t)b * th * tw * 2, th, tw, h, w, y, xo, gx, gy);
    lwb::Taps tp;
    lwb::make_taps(gx, gy, h, w, align_corners, tp);
    const size_t plane = (size_t)h * w;
    const int c0 = blockIdx.y * c_per_block, c1 = min(C, c0 + c_per_block);
    const float* src = x + (size_t)(src_batch == 1 ? 0 : b) * C * plane;
    float* dst = out + (size_t)b * C * plane + pix;
    for (int c = c0; c < c1; c++) {
        const float* pl = src + (size_t)c * plane;
        float v = 0.f;
        if (tp.m & 1) v += __ldg(pl + tp.o00) * tp.w00;
        if (tp.m & 2) v += __ldg(pl + tp.o00 + 1) * tp.w01;
        if (tp.m & 4) v += __ldg(pl + tp.o00 + w) * tp.w10;
        if (tp.m & 8) v += __ldg(pl + tp.o00 + w + 1) * tp.w11;
        if (accumulate) v += dst[(size_t)c * plane];
        dst[(size_t)c * plane] = v;
    }
}

}  // namespace

extern "C" int lwb_warp_nchw(const float* x, int src_batch, int channels, int h, int w,
                             const float* T, int batch, int th, int tw, int align_corners,
                             float* out, int accumulate, lwb_stream_t stream)
{
    LWB_CHECK_ARG(x && T && out, "null pointer");
    LWB_CHECK_ARG(channels > 0 && h > 0 && w > 0 && batch > 0 && th > 0 && tw > 0, "non-positive size");
    LWB_CHECK_ARG(src_batch == 1 || src_batch == batch, "src_batch must be 1 or batch");
    LWB_CHECK_ARG(batch <= 65535, "batch too large");
    const int t = 256;
    const int c_per_block = channels >= 64 ? 16 : channels;
    dim3 grid(lwb::ceil_div((long)h * w, t), lwb::ceil_div(channels, c_per_block), batch);
    k_warp_nchw<<<grid, t, 0, (cudaStream_t)stream>>>(x, src_batch, channels, h, w, T, batch, th, tw,
                                                      align_corners, out, accumulate, c_per_block);
    LWB_LAUNCH_OK();
    return LWB_OK;
}
