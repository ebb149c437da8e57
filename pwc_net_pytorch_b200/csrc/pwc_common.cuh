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

// Pixel (x, y) displaced by the flow (u, v) -> tap.  This is F.grid_sample(bilinear, zeros,
// align_corners=True) at the un-normalised coordinate (x+u, y+v), which is what modules.py:36-41
// evaluates under torch 0.4.0: out = sum over the 4 corners of
// w * x[corner], corners outside [0,W)x[0,H) contribute 0.
// The flow is split into floor + fraction *before* the pixel coordinate is added: u - floor(u) is
// exact in fp32, so the bilinear weights carry no rounding from the image width (the reference's own
// normalise/denormalise round trip costs ~W * 2^-24 px).
// ax/ay (optional) receive the fractions, x0/y0 the top-left corner.
__device__ __forceinline__ Tap make_tap(int x, int y, float u, float v, int H, int W,
                                        float* ax_out = nullptr, float* ay_out = nullptr,
                                        int* x0_out = nullptr, int* y0_out = nullptr)
{
    Tap t;
    t.w00 = t.w01 = t.w10 = t.w11 = 0.0f;
    t.off = -1; t.dx = 0; t.dyw = 0;
    // rejects NaN / Inf / absurd displacements (also keeps the int conversions well defined)
    if (!(fabsf(u) < 1.0e6f) || !(fabsf(v) < 1.0e6f)) return t;
    const float fu = floorf(u), fv = floorf(v);
    const float ax = u - fu, ay = v - fv;
    const int x0 = x + (int)fu, y0 = y + (int)fv;
    if (ax_out) *ax_out = ax;
    if (ay_out) *ay_out = ay;
    if (x0_out) *x0_out = x0;
    if (y0_out) *y0_out = y0;
    if (x0 < -1 || x0 >= W || y0 < -1 || y0 >= H) return t;      // all four corners outside
    const int x1 = x0 + 1, y1 = y0 + 1;
    const bool inx0 = (x0 >= 0), inx1 = (x1 < W);
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

// ---- model.py:78: flow = F.upsample(flow_coarse, scale_factor=2, mode='bilinear') * 2, folded into the flow read.
// F.upsample(bilinear) is align_corners=False in torch 0.4.0 and 2.x alike.  The
// arithmetic below is ATen's upsample_bilinear2d_out_frame as compiled for sm_100 (read from the SASS of torch
// 2.11's libtorch_cuda.so): source index s = max(fma(i + 0.5, 0.5, -0.5), 0), i0 = (int)s, i1 = i0 + (i0 < n-1),
// lambda1 = s - i0, lambda0 = 1 - lambda1; a blend a*l0 + b*l1 is evaluated as fma(l0, a, l1*b) with the second
// product rounded on its own, for the horizontal and then the vertical pair; `* 2` is exact.  The result equals
// `F.interpolate(c, scale_factor=2, mode='bilinear', align_corners=False) * 2` bit for bit.
__device__ __forceinline__ float up2_source(int i, int n_coarse, int& i0, int& i1)
{
    const float s = fmaxf(__fmaf_rn((float)i + 0.5f, 0.5f, -0.5f), 0.0f);
    i0 = (int)s;
    i1 = i0 + (i0 < n_coarse - 1 ? 1 : 0);
    return s - (float)i0;
}
__device__ __forceinline__ float up2_blend(float v00, float v01, float v10, float v11, float lx1, float ly1)
{
    const float lx0 = 1.0f - lx1, ly0 = 1.0f - ly1;
    const float top = __fmaf_rn(lx0, v00, __fmul_rn(lx1, v01));
    const float bot = __fmaf_rn(lx0, v10, __fmul_rn(lx1, v11));
    return __fmul_rn(2.0f, __fmaf_rn(ly0, top, __fmul_rn(ly1, bot)));
}
// (u, v) of fine pixel (x, y) from image n's coarse flow planes cu, cv ([Hc][Wc] each), through global memory
__device__ __forceinline__ void up2_flow_at(const float* __restrict__ cu, const float* __restrict__ cv, int Hc,
                                            int Wc, int x, int y, float& u, float& v)
{
    int x0, x1, y0, y1;
    const float lx1 = up2_source(x, Wc, x0, x1), ly1 = up2_source(y, Hc, y0, y1);
    const int a = y0 * Wc + x0, b = y0 * Wc + x1, c = y1 * Wc + x0, d = y1 * Wc + x1;
    u = up2_blend(__ldg(cu + a), __ldg(cu + b), __ldg(cu + c), __ldg(cu + d), lx1, ly1);
    v = up2_blend(__ldg(cv + a), __ldg(cv + b), __ldg(cv + c), __ldg(cv + d), lx1, ly1);
}

// Programmatic dependent launch: blocks until the grid this one was launched behind has completed and
// its memory is visible (a no-op for a normally launched kernel).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace pwc
