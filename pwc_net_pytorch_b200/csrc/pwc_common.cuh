// Shared device helpers for the PWC-Net warp + cost-volume kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pwc {

__host__ __device__ constexpr int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int round_up(int a, int b) { return cdiv(a, b) * b; }

// Bilinear tap of one warped pixel: four corner weights (already zeroed for corners that fall
// outside the image, i.e. grid_sample's padding_mode='zeros') and the clamped corner offsets.
// Corner addresses are base+off, base+off+dx, base+off+dyw, base+off+dyw+dx, all in range.
struct Tap {
    float w00, w01, w10, w11;
    int off;    // ya*W + xa, or -1 when every weight is zero (nothing to read)
    int dx;     // 0 or 1
    int dyw;    // 0 or W
};

// Sample position (sx, sy) in source-pixel units -> tap.  This is F.grid_sample(bilinear, zeros,
// align_corners=True) at the un-normalised coordinate, which is what modules.py:36-41 evaluates
// under torch 0.4.0: out = sum over the 4 corners of w * x[corner],
// corners outside [0,W)x[0,H) contribute 0.
__device__ __forceinline__ Tap make_tap(float sx, float sy, int H, int W)
{
    Tap t;
    // Rejects NaN/Inf and anything whose 4 corners are all outside (also keeps the int
    // conversion below well defined).
    const bool live = (sx > -1.0f) && (sx < (float)W) && (sy > -1.0f) && (sy < (float)H);
    if (!live) {
        t.w00 = t.w01 = t.w10 = t.w11 = 0.0f;
        t.off = -1; t.dx = 0; t.dyw = 0;
        return t;
    }
    const float fx = floorf(sx), fy = floorf(sy);
    const float ax = sx - fx, ay = sy - fy;
    const int x0 = (int)fx, y0 = (int)fy;
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool inx0 = (x0 >= 0), inx1 = (x1 < W);   // x0 < W and x1 >= 0 are implied by `live`
    const bool iny0 = (y0 >= 0), iny1 = (y1 < H);
    const float bx = 1.0f - ax, by = 1.0f - ay;
    t.w00 = (inx0 && iny0) ? bx * by : 0.0f;
    t.w01 = (inx1 && iny0) ? ax * by : 0.0f;
    t.w10 = (inx0 && iny1) ? bx * ay : 0.0f;
    t.w11 = (inx1 && iny1) ? ax * ay : 0.0f;
    const int xa = inx0 ? x0 : 0, xb = inx1 ? x1 : W - 1;
    const int ya = iny0 ? y0 : 0, yb = iny1 ? y1 : H - 1;
    t.off = ya * W + xa;
    t.dx = xb - xa;
    t.dyw = (yb - ya) * W;
    return t;
}

__device__ __forceinline__ float tap_sample(const Tap& t, const float* __restrict__ plane)
{
    const float* p = plane + t.off;
    const float v00 = __ldg(p), v01 = __ldg(p + t.dx);
    const float v10 = __ldg(p + t.dyw), v11 = __ldg(p + t.dyw + t.dx);
    return fmaf(t.w11, v11, fmaf(t.w10, v10, fmaf(t.w01, v01, t.w00 * v00)));
}

__device__ __forceinline__ float leaky(float v, float slope) { return v < 0.0f ? v * slope : v; }

}  // namespace pwc
