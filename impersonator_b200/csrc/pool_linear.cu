// Small fp32 kernels around the conv engine for the HMR encoder (networks/hmr.py:119-166, 214-252, 275-300):
// the stem's max-pool, the global average pool behind post_bn + ReLU, and the fully connected layers of the
// iterative theta regressor.  All three are far from any roofline-relevant size (per image: 0.8 M max-pool
// outputs, a 49 x 2048 mean, 3 x 3.3 M multiply-adds); they exist so that no torch operator sits on the path.
#include "common.cuh"

namespace {

// F.max_pool2d(x, kernel_size=k, stride=s, ceil_mode=True), no padding (networks/hmr.py:150):
// out = ceil((in - k) / s) + 1 windows, the last one clipped at the border.  NCHW in -> NHWC out.
__global__ void __launch_bounds__(256) k_maxpool_nchw_to_nhwc(const float* __restrict__ x, int n, int c, int h, int w,
                                                              int k, int s, int ho, int wo, float* __restrict__ out)
{
    const long total = (long)n * ho * wo * c;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    // consecutive threads walk x (coalesced reads of the NCHW plane); the NHWC write is strided by c
    const int ox = (int)(i % wo), oy = (int)((i / wo) % ho), ch = (int)((i / ((long)wo * ho)) % c), b = (int)(i / ((long)wo * ho * c));
    const float* p = x + ((size_t)b * c + ch) * h * w;
    float m = -INFINITY;
    for (int dy = 0; dy < k; dy++) {
        const int y = oy * s + dy;
        if (y >= h) break;
        for (int dx = 0; dx < k; dx++) {
            const int xx = ox * s + dx;
            if (xx >= w) break;
            m = fmaxf(m, __ldg(p + (size_t)y * w + xx));
        }
    }
    out[(((size_t)b * ho + oy) * wo + ox) * c + ch] = m;
}

// out[b, ch] = mean over hw of relu?(x[b, p, ch] * scale[ch] + shift[ch])   (post_bn + ReLU + avg_pool2d(7), hmr.py:160-163)
__global__ void __launch_bounds__(256) k_global_avgpool_nhwc(const float* __restrict__ x, int n, int hw, int c,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             int relu, float* __restrict__ out, int ld_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * c) return;
    const int ch = i % c, b = i / c;
    const float sc = scale ? scale[ch] : 1.f, sh = shift ? shift[ch] : 0.f;
    const float* p = x + (size_t)b * hw * c + ch;
    float acc = 0.f;
    for (int q = 0; q < hw; q++) {
        float v = fmaf(__ldg(p + (size_t)q * c), sc, sh);
        if (relu) v = fmaxf(v, 0.f);
        acc += v;
    }
    out[(size_t)b * ld_out + ch] = acc / (float)hw;
}

// nn.Linear: out[b, m] (+)= relu?(sum_k x[b, k] * w[m, k] + bias[m]); one warp per (b, m).
__global__ void __launch_bounds__(256) k_linear(const float* __restrict__ x, int ld_x, const float* __restrict__ w,
                                                const float* __restrict__ bias, int n, int k, int m, int relu, int accumulate,
                                                float* __restrict__ out, int ld_out)
{
    const int warp = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (warp >= n * m) return;
    const int b = warp / m, j = warp % m;
    const float* xr = x + (size_t)b * ld_x;
    const float* wr = w + (size_t)j * k;
    float acc = 0.f;
    for (int q = lane; q < k; q += 32) acc = fmaf(__ldg(xr + q), __ldg(wr + q), acc);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        float v = acc + (bias ? bias[j] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        float* o = out + (size_t)b * ld_out + j;
        *o = accumulate ? *o + v : v;
    }
}

}  // namespace

extern "C" int lwb_maxpool_nchw_to_nhwc(const float* x, int n, int c, int h, int w, int k, int stride, float* out, lwb_stream_t stream)
{
    LWB_CHECK_ARG(x && out, "null pointer");
    LWB_CHECK_ARG(n > 0 && c > 0 && h >= k && w >= k && k > 0 && stride > 0, "bad sizes");
    const int ho = lwb::ceil_div(h - k, stride) + 1, wo = lwb::ceil_div(w - k, stride) + 1;
    const long total = (long)n * ho * wo * c;
    k_maxpool_nchw_to_nhwc<<<lwb::ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, n, c, h, w, k, stride, ho, wo, out);
    LWB_LAUNCH_OK();
    return LWB_OK;
}

extern "C" int lwb_global_avgpool_nhwc(const float* x, int n, int hw, int c, const float* scale, const float* shift, int relu,
                                       float* out, int ld_out, lwb_stream_t stream)
{
    LWB_CHECK_ARG(x && out, "null pointer");
    LWB_CHECK_ARG(n > 0 && hw > 0 && c > 0 && ld_out >= c, "bad sizes");
    LWB_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift go together");
    k_global_avgpool_nhwc<<<lwb::ceil_div((long)n * c, 256), 256, 0, (cudaStream_t)stream>>>(x, n, hw, c, scale, shift, relu, out, ld_out);
    LWB_LAUNCH_OK();
    return LWB_OK;
}

extern "C" int lwb_linear(const float* x, int ld_x, const float* w, const float* bias, int n, int k, int m, int relu, int accumulate,
                          float* out, int ld_out, lwb_stream_t stream)
{
    LWB_CHECK_ARG(x && w && out, "null pointer");
    LWB_CHECK_ARG(n > 0 && k > 0 && m > 0 && ld_x >= k && ld_out >= m, "bad sizes");
    const long threads = (long)n * m * 32;
    k_linear<<<lwb::ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(x, ld_x, w, bias, n, k, m, relu, accumulate, out, ld_out);
    LWB_LAUNCH_OK();
    return LWB_OK;
}
