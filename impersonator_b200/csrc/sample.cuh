// Flow resize + bilinear tap set-up shared by the warp kernels.
//
// Formulas are torch's (the reference calls F.interpolate / F.grid_sample, whose arithmetic
// lives in ATen):
//   resize_trans  = upsample_bilinear2d(align_corners=True)   ATen/native/UpSample.h:271-296,442-476
//   stn           = grid_sampler_2d(bilinear, zeros)          ATen/native/GridSampler.h:27-36
#pragma once
#include <cuda_runtime.h>

namespace lwb {

// Flow T [th,tw,2] evaluated at output pixel (y,x) of an (h,w) grid.  Equal sizes: plain read.
// Otherwise bilinear resize with align_corners=True (networks/generator.py:307):
//   scale = (in-1)/(out-1); src = scale*dst; i0 = int(src); i1 = i0 + (i0 < in-1); l1 = src-i0; l0 = 1-l1
//   out = l0y*(l0x*v00 + l1x*v01) + l1y*(l0x*v10 + l1x*v11)
__device__ __forceinline__ void flow_at(const float* __restrict__ T, int th, int tw, int h, int w,
                                        int y, int x, float& gx, float& gy)
{
    if (th == h && tw == w) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(T) + (size_t)y * tw + x);
        gx = v.x; gy = v.y;
        return;
    }
    const float sh = h > 1 ? (float)(th - 1) / (float)(h - 1) : 0.f;
    const float sw = w > 1 ? (float)(tw - 1) / (float)(w - 1) : 0.f;
    const float fy = sh * (float)y, fx = sw * (float)x;
    const int y0 = min((int)fy, th - 1), x0 = min((int)fx, tw - 1);
    const int yp = (y0 < th - 1) ? 1 : 0, xp = (x0 < tw - 1) ? 1 : 0;
    const float ly1 = fminf(fmaxf(fy - (float)y0, 0.f), 1.f), ly0 = 1.f - ly1;
    const float lx1 = fminf(fmaxf(fx - (float)x0, 0.f), 1.f), lx0 = 1.f - lx1;
    const float2* p = reinterpret_cast<const float2*>(T) + (size_t)y0 * tw + x0;
    const float2 v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + (size_t)yp * tw), v11 = __ldg(p + (size_t)yp * tw + xp);
    gx = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
    gy = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
}

struct Taps {
    int   o00;                  // linear offset (y0*w + x0) of the north-west tap (may be out of range)
    int   m;                     // validity bits: 1 nw, 2 ne, 4 sw, 8 se
    float w00, w01, w10, w11;
};

// grid_sampler_2d, bilinear, padding_mode=zeros.
__device__ __forceinline__ void make_taps(float gx, float gy, int h, int w, int align_corners, Taps& t)
{
    const float ix = align_corners ? ((gx + 1.f) / 2.f) * (float)(w - 1) : ((gx + 1.f) * (float)w - 1.f) / 2.f;
    const float iy = align_corners ? ((gy + 1.f) / 2.f) * (float)(h - 1) : ((gy + 1.f) * (float)h - 1.f) / 2.f;
    // Clamp far-away (or non-finite) coordinates so the int conversion is defined; all taps are then invalid.
    const float cx = fminf(fmaxf(ix, -4.f), (float)w + 4.f), cy = fminf(fmaxf(iy, -4.f), (float)h + 4.f);
    const bool finite = (ix == ix) && (iy == iy);
    const float x0f = floorf(cx), y0f = floorf(cy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float wx1 = cx - x0f, wx0 = (x0f + 1.f) - cx, wy1 = cy - y0f, wy0 = (y0f + 1.f) - cy;
    t.w00 = wx0 * wy0; t.w01 = wx1 * wy0; t.w10 = wx0 * wy1; t.w11 = wx1 * wy1;
    const bool vx0 = x0 >= 0 && x0 < w, vx1 = x0 + 1 >= 0 && x0 + 1 < w;
    const bool vy0 = y0 >= 0 && y0 < h, vy1 = y0 + 1 >= 0 && y0 + 1 < h;
    t.m = finite ? ((vy0 && vx0) ? 1 : 0) | ((vy0 && vx1) ? 2 : 0) | ((vy1 && vx0) ? 4 : 0) | ((vy1 && vx1) ? 8 : 0) : 0;
    t.o00 = y0 * w + x0;
}

}  // namespace lwb
