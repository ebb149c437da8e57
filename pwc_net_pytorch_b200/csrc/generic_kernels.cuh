// Generic (any pad/kernel/displacement/stride) correlation kernels and the standalone warp.
// These cover the parts of the reference operator's parameter space that the tiled fast paths
// do not (kernel_size > 1, stride1 > 1, pad != max_displacement).  One thread per result
// element, direct NCHW addressing: no padded channels-last scratch copy
// (correlation_cuda_kernel.cu:10-32 and the fills of correlation_cuda.c:36-42 are not needed).
#pragma once
#include "pwc_common.cuh"

namespace pwc {

struct CorrGeom {
    int B, C, H, W;
    int pad, k, md, s1, s2;
    int kr, r, D, oc, oh, ow;
};

// value of the zero-padded second operand at padded coordinates (yp, xp); warped on the fly
// when flow != nullptr (the pixel's own flow decides where f2 is sampled, model.py:80).
__device__ __forceinline__ float second_operand(const float* __restrict__ f2n,
                                                const float* __restrict__ flown, int c, int yp,
                                                int xp, const CorrGeom& g)
{
    const int y = yp - g.pad, x = xp - g.pad;
    if (y < 0 || y >= g.H || x < 0 || x >= g.W) return 0.0f;
    const size_t HW = (size_t)g.H * g.W;
    const float* plane = f2n + (size_t)c * HW;
    if (flown == nullptr) return __ldg(plane + (size_t)y * g.W + x);
    const float u = __ldg(flown + (size_t)y * g.W + x);
    const float v = __ldg(flown + HW + (size_t)y * g.W + x);
    const Tap t = make_tap(x, y, u, v, g.H, g.W);
    return t.off < 0 ? 0.0f : tap_sample(t, plane);
}

__device__ __forceinline__ float first_operand(const float* __restrict__ f1n, int c, int yp,
                                               int xp, const CorrGeom& g)
{
    const int y = yp - g.pad, x = xp - g.pad;
    if (y < 0 || y >= g.H || x < 0 || x >= g.W) return 0.0f;
    return __ldg(f1n + ((size_t)c * g.H + y) * g.W + x);
}

// correlation_cuda_kernel.cu:45-101 for arbitrary parameters, optionally fused with the warp
// and the activation.
__global__ void __launch_bounds__(256)
corr_fwd_generic_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                        const float* __restrict__ flow, float* __restrict__ out, CorrGeom g,
                        int act, float slope, long long obs, long long fbs)
{
    const size_t total = (size_t)g.B * g.oc * g.oh * g.ow;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int bx = (int)(idx % g.ow);
    const int by = (int)((idx / g.ow) % g.oh);
    const int tc = (int)((idx / ((size_t)g.ow * g.oh)) % g.oc);
    const int n = (int)(idx / ((size_t)g.ow * g.oh * g.oc));
    const int ti = tc % g.D - g.r, tj = tc / g.D - g.r;
    const int y1 = by * g.s1 + g.md + g.kr, x1 = bx * g.s1 + g.md + g.kr;
    const int y2 = y1 + tj * g.s2, x2 = x1 + ti * g.s2;
    const size_t HW = (size_t)g.H * g.W;
    const float* f1n = f1 + (size_t)n * g.C * HW;
    const float* f2n = f2 + (size_t)n * g.C * HW;
    const float* flown = flow ? flow + (size_t)n * (size_t)fbs : nullptr;   // fbs: flow batch stride
    float s = 0.0f;
    for (int j = -g.kr; j <= g.kr; ++j)
        for (int i = -g.kr; i <= g.kr; ++i)
            for (int c = 0; c < g.C; ++c)
                s = fmaf(first_operand(f1n, c, y1 + j, x1 + i, g),
                         second_operand(f2n, flown, c, y2 + j, x2 + i, g), s);
    float v = s / (float)(g.k * g.k * g.C);
    if (act) v = leaky(v, slope);
    out[(size_t)n * (size_t)obs + (idx - (size_t)n * g.oc * g.oh * g.ow)] = v;   // obs: output batch stride
}

// model.py:78 as a stand-alone pass: fine = F.upsample(coarse, 2, 'bilinear') * 2, written with a batch stride
// (image n's [2][H][W] block starts at fine + n * fobs -- the last two channels of the flow estimator's
// concatenated input).  Used where the fold into the fused kernel's flow read does not apply.
__global__ void __launch_bounds__(256)
flow_up2_kernel(const float* __restrict__ coarse, float* __restrict__ fine, long long fobs, int B, int H, int W)
{
    const int Hc = H >> 1, Wc = W >> 1;
    const size_t HW = (size_t)H * W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)B * HW) return;
    const int n = (int)(idx / HW);
    const int pix = (int)(idx - (size_t)n * HW);
    const int y = pix / W, x = pix - y * W;
    const float* cu = coarse + (size_t)n * 2 * Hc * Wc;
    float u, v;
    up2_flow_at(cu, cu + (size_t)Hc * Wc, Hc, Wc, x, y, u, v);
    float* o = fine + (size_t)n * (size_t)fobs + pix;
    o[0] = u;
    o[HW] = v;
}

// correlation_cuda_kernel.cu:119-196 (gradInput1) and :211-288 (gradInput2) for stride1 == 1,
// arbitrary pad / kernel_size / displacement / stride2; same window arithmetic as the reference
// (C truncating division).  One thread per (n, c, y, x); both gradients in one pass.
// `second` is the (already warped) second operand.  gate != nullptr applies leaky_relu_'s
// derivative by the sign of the forward output.
__global__ void __launch_bounds__(256)
corr_bwd_generic_kernel(const float* __restrict__ gout, const float* __restrict__ gate,
                        const float* __restrict__ f1, const float* __restrict__ second,
                        float* __restrict__ g1, float* __restrict__ g2, CorrGeom g, float slope, long long gbs,
                        long long gate_bs)
{
    const size_t total = (size_t)g.B * g.C * g.H * g.W;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int xx = (int)(idx % g.W);
    const int yy = (int)((idx / g.W) % g.H);
    const int c = (int)((idx / ((size_t)g.W * g.H)) % g.C);
    const int n = (int)(idx / ((size_t)g.W * g.H * g.C));
    const size_t HW = (size_t)g.H * g.W;
    const float* f1n = f1 + (size_t)n * g.C * HW;
    const float* f2n = second + (size_t)n * g.C * HW;
    const size_t ohw = (size_t)g.oh * g.ow;
    const float* gon = gout + (size_t)n * (size_t)gbs;           // batch strides of the output gradient / the gate
    const float* gaten = gate ? gate + (size_t)n * (size_t)gate_bs : nullptr;
    const int y = yy * g.s1 + g.pad, x = xx * g.s1 + g.pad;
    const float nelems = (float)(g.k * g.k * g.C);

    float a1 = 0.0f, a2 = 0.0f;
    for (int tc = 0; tc < g.oc; ++tc) {
        const int i2 = (tc % g.D - g.r) * g.s2, j2 = (tc / g.D - g.r) * g.s2;
        // ---- input1 window (:129-149): independent of tc ----
        {
            int xmin = (x - g.kr - g.md) / g.s1, ymin = (y - g.kr - g.md) / g.s1;
            int xmax = (x + g.kr - g.md) / g.s1, ymax = (y + g.kr - g.md) / g.s1;
            const bool skip = (xmax < 0 || ymax < 0 || xmin >= g.ow || ymin >= g.oh) ||
                              (xmin > xmax || ymin > ymax);
            if (!skip) {
                xmin = max(0, xmin); xmax = min(g.ow - 1, xmax);
                ymin = max(0, ymin); ymax = min(g.oh - 1, ymax);
                const float v2 = second_operand(f2n, nullptr, c, y + j2, x + i2, g);
                float s = 0.0f;
                for (int j = ymin; j <= ymax; ++j)
                    for (int i = xmin; i <= xmax; ++i) {
                        const size_t o = (size_t)tc * ohw + (size_t)j * g.ow + i;
                        float gv = __ldg(gon + o);
                        if (gaten && __ldg(gaten + o) < 0.0f) gv *= slope;
                        s += gv;
                    }
                a1 = fmaf(s, v2, a1);
            }
        }
        // ---- input2 window (:246-266) ----
        {
            int xmin = (x - g.kr - g.md - i2) / g.s1, ymin = (y - g.kr - g.md - j2) / g.s1;
            int xmax = (x + g.kr - g.md - i2) / g.s1, ymax = (y + g.kr - g.md - j2) / g.s1;
            const bool skip = (xmax < 0 || ymax < 0 || xmin >= g.ow || ymin >= g.oh) ||
                              (xmin > xmax || ymin > ymax);
            if (!skip) {
                xmin = max(0, xmin); xmax = min(g.ow - 1, xmax);
                ymin = max(0, ymin); ymax = min(g.oh - 1, ymax);
                const float v1 = first_operand(f1n, c, y - j2, x - i2, g);
                float s = 0.0f;
                for (int j = ymin; j <= ymax; ++j)
                    for (int i = xmin; i <= xmax; ++i) {
                        const size_t o = (size_t)tc * ohw + (size_t)j * g.ow + i;
                        float gv = __ldg(gon + o);
                        if (gaten && __ldg(gaten + o) < 0.0f) gv *= slope;
                        s += gv;
                    }
                a2 = fmaf(s, v1, a2);
            }
        }
    }
    g1[idx] = a1 / nelems;
    g2[idx] = a2 / nelems;
}

// WarpingLayer.forward (modules.py:31-42): one thread per (n, channel group, y, x).
template <int CPT>
__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow,
                float* __restrict__ out, int B, int C, int H, int W)
{
    const int cgroups = cdiv(C, CPT);
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)B * cgroups * HW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t pix = idx % HW;
    const int cg = (int)((idx / HW) % cgroups);
    const int n = (int)(idx / (HW * cgroups));
    const int yy = (int)(pix / W), xx = (int)(pix % W);
    const float u = __ldg(flow + (size_t)n * 2 * HW + pix);
    const float v = __ldg(flow + (size_t)n * 2 * HW + HW + pix);
    const Tap t = make_tap(xx, yy, u, v, H, W);
    const int c0 = cg * CPT;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = c0 + k;
        if (c < C) {
            const size_t plane = ((size_t)n * C + c) * HW;
            out[plane + pix] = t.off < 0 ? 0.0f : tap_sample(t, x + plane);
        }
    }
}

// Autograd of WarpingLayer.  grad_x must be zeroed beforehand
// (the ABI entry does it on the same stream).  One thread per (n, channel group, y, x): the feature
// gradient is a scatter-add; the flow gradient is a sum over channels -- written directly when there
// is a single channel group (deterministic), accumulated with atomicAdd into a zeroed buffer otherwise.
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x,
                const float* __restrict__ flow, float* __restrict__ gx,
                float* __restrict__ gflow, int B, int C, int H, int W, int cpt, int cgroups)
{
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)B * HW * cgroups;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t pix = idx % HW;
    const int cg = (int)((idx / HW) % cgroups);
    const int n = (int)(idx / (HW * cgroups));
    const int yy = (int)(pix / W), xx = (int)(pix % W);
    const float u = __ldg(flow + (size_t)n * 2 * HW + pix);
    const float v = __ldg(flow + (size_t)n * 2 * HW + HW + pix);
    float ax = 0.0f, ay = 0.0f;
    int x0 = 0, y0 = 0;
    const Tap t = make_tap(xx, yy, u, v, H, W, &ax, &ay, &x0, &y0);
    float gu = 0.0f, gv = 0.0f;
    if (t.off >= 0) {
        // corners outside the image read as 0 (their clamped addresses are masked out)
        const bool inx0 = x0 >= 0, inx1 = x0 + 1 < W, iny0 = y0 >= 0, iny1 = y0 + 1 < H;
        const float m00 = (inx0 && iny0) ? 1.0f : 0.0f, m01 = (inx1 && iny0) ? 1.0f : 0.0f;
        const float m10 = (inx0 && iny1) ? 1.0f : 0.0f, m11 = (inx1 && iny1) ? 1.0f : 0.0f;
        const int c_end = min(C, (cg + 1) * cpt);
        for (int c = cg * cpt; c < c_end; ++c) {
            const size_t plane = ((size_t)n * C + c) * HW;
            const float g = __ldg(gout + plane + pix);
            const float* p = x + plane + t.off;
            const float v00 = m00 * __ldg(p), v01 = m01 * __ldg(p + t.dx);
            const float v10 = m10 * __ldg(p + t.dyw), v11 = m11 * __ldg(p + t.dyw + t.dx);
            gu = fmaf(g, fmaf(v11 - v10, ay, (v01 - v00) * (1.0f - ay)), gu);
            gv = fmaf(g, fmaf(v11 - v01, ax, (v10 - v00) * (1.0f - ax)), gv);
            if (gx) {
                float* q = gx + plane + t.off;
                if (t.w00 != 0.0f) atomicAdd(q, g * t.w00);
                if (t.w01 != 0.0f) atomicAdd(q + t.dx, g * t.w01);
                if (t.w10 != 0.0f) atomicAdd(q + t.dyw, g * t.w10);
                if (t.w11 != 0.0f) atomicAdd(q + t.dyw + t.dx, g * t.w11);
            }
        }
    }
    if (gflow) {
        float* gf = gflow + (size_t)n * 2 * HW + pix;
        if (cgroups == 1) {
            gf[0] = gu;
            gf[HW] = gv;
        } else if (t.off >= 0) {
            atomicAdd(gf, gu);
            atomicAdd(gf + HW, gv);
        }
    }
}

// Feature-gradient scatter with 128-bit reductions.  The scalar scatter above is bound by the SM's
// reduction-issue rate (one RED.E.ADD.F32 per lane, corner and channel: 1.29 cycles each, measured).
// Here four channels share one REDG.E.ADD.F32x4 (red.global.add.v4.f32, sm_90+): the gradient is
// accumulated in a scratch buffer laid out [B][ceil(C/8)][H][W][8] (zeroed by the caller), and
// deinterleave8_kernel then writes the NCHW result (which therefore needs no memset).
// Two adjacent lanes own the two channel quads of one (pixel, channel octet), so their two reductions
// fall into the SAME 32-byte sector: one sector request per lane pair instead of one per lane (5.7 M
// instead of 10.9 M reduction sectors at the level-2 shape).  The pair also sums its flow-gradient
// terms with one shuffle before the atomic.  One thread per (n, channel octet, y, x, half).
// (Fallback of warp_bwd_tile.cuh for W % 4 != 0 or unaligned pointers.)
__global__ void __launch_bounds__(256)
warp_bwd_v8_kernel(const float* __restrict__ gout, const float* __restrict__ x,
                   const float* __restrict__ flow, float* __restrict__ gx8,
                   float* __restrict__ gflow, float* __restrict__ warped_out, int B, int C, int H, int W, int cocts)
{
    // grid = (ceil(2*H*W / 256), channel octets, images): no 64-bit divisions on the index path (with a flat
    // 64-bit index they were most of the kernel's instructions: 60 of 95 us, measured by ablation)
    const int HWi = H * W;
    const size_t HW = (size_t)HWi;
    const int t2 = blockIdx.x * blockDim.x + threadIdx.x;
    // lanes that take part in the pair shuffle at the end: taken while the warp is still converged (the
    // shuffle names its participants explicitly instead of trusting __activemask() after divergent code)
    const unsigned pair_mask = __ballot_sync(0xffffffffu, t2 < 2 * HWi);
    if (t2 >= 2 * HWi) return;                        // 2*H*W is even: a lane pair is in or out together
    const int half = t2 & 1;
    const int pix = t2 >> 1;
    const int co = blockIdx.y;
    const int n = blockIdx.z;
    (void)B;
    const int cq = 2 * co + half;                     // this lane's channel quad
    const int yy = pix / W, xx = pix - yy * W;
    const float u = __ldg(flow + (size_t)n * 2 * HW + pix);
    const float v = __ldg(flow + (size_t)n * 2 * HW + HW + pix);
    float ax = 0.0f, ay = 0.0f;
    int x0 = 0, y0 = 0;
    const Tap t = make_tap(xx, yy, u, v, H, W, &ax, &ay, &x0, &y0);
    const bool live = t.off >= 0;   // else: nothing to scatter, zero flow gradient (the buffers are zeroed)
    if (!live && warped_out) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * cq + k < C) warped_out[((size_t)n * C + 4 * cq + k) * HW + pix] = 0.0f;
    }
    const bool inx0 = x0 >= 0, inx1 = x0 + 1 < W, iny0 = y0 >= 0, iny1 = y0 + 1 < H;
    const float m00 = (inx0 && iny0) ? 1.0f : 0.0f, m01 = (inx1 && iny0) ? 1.0f : 0.0f;
    const float m10 = (inx0 && iny1) ? 1.0f : 0.0f, m11 = (inx1 && iny1) ? 1.0f : 0.0f;
    float g[4];
    float gu = 0.0f, gv = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = 4 * cq + k;
        g[k] = 0.0f;
        if (live && c < C) {
            const size_t plane = ((size_t)n * C + c) * HW;
            g[k] = __ldg(gout + plane + pix);
            const float* p = x + plane + t.off;
            const float v00 = m00 * __ldg(p), v01 = m01 * __ldg(p + t.dx);
            const float v10 = m10 * __ldg(p + t.dyw), v11 = m11 * __ldg(p + t.dyw + t.dx);
            gu = fmaf(g[k], fmaf(v11 - v10, ay, (v01 - v00) * (1.0f - ay)), gu);
            gv = fmaf(g[k], fmaf(v11 - v01, ax, (v10 - v00) * (1.0f - ax)), gv);
            if (warped_out)   // same expression as tap_sample (weights of masked corners are 0)
                warped_out[plane + pix] = fmaf(t.w11, v11, fmaf(t.w10, v10, fmaf(t.w01, v01, t.w00 * v00)));
        }
    }
    if (live && gx8 && 4 * cq < C) {
        float* q = gx8 + (((size_t)n * cocts + co) * HW + t.off) * 8 + 4 * half;
        const size_t sdx = (size_t)t.dx * 8, sdy = (size_t)t.dyw * 8;
        if (t.w00 != 0.0f)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(g[0] * t.w00), "f"(g[1] * t.w00),
                         "f"(g[2] * t.w00), "f"(g[3] * t.w00) : "memory");
        if (t.w01 != 0.0f)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + sdx), "f"(g[0] * t.w01), "f"(g[1] * t.w01),
                         "f"(g[2] * t.w01), "f"(g[3] * t.w01) : "memory");
        if (t.w10 != 0.0f)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + sdy), "f"(g[0] * t.w10), "f"(g[1] * t.w10),
                         "f"(g[2] * t.w10), "f"(g[3] * t.w10) : "memory");
        if (t.w11 != 0.0f)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + sdy + sdx), "f"(g[0] * t.w11),
                         "f"(g[1] * t.w11), "f"(g[2] * t.w11), "f"(g[3] * t.w11) : "memory");
    }
    if (gflow) {
        // every in-range lane reaches this point (no early exit above), dead taps carry zeros
        gu += __shfl_xor_sync(pair_mask, gu, 1);
        gv += __shfl_xor_sync(pair_mask, gv, 1);
        if (half == 0 && live) {
            float* gf = gflow + (size_t)n * 2 * HW + pix;
            if (cocts == 1) {
                gf[0] = gu;
                gf[HW] = gv;
            } else {
                atomicAdd(gf, gu);
                atomicAdd(gf + HW, gv);
            }
        }
    }
}

// [B][ceil(C/8)][H][W][8] scratch -> NCHW gradient (fully overwrites the destination).
__global__ void __launch_bounds__(256)
deinterleave8_kernel(const float* __restrict__ src8, float* __restrict__ dst, int B, int C, int H, int W, int cocts)
{
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)B * HW * cocts;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t pix = idx % HW;
    const int co = (int)((idx / HW) % cocts);
    const int n = (int)(idx / (HW * cocts));
    const float4* s = reinterpret_cast<const float4*>(src8) + (((size_t)n * cocts + co) * HW + pix) * 2;
    const float4 a = __ldg(s), b = __ldg(s + 1);
    const float vals[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = 8 * co + k;
        if (c < C) dst[((size_t)n * C + c) * HW + pix] = vals[k];
    }
    griddep_wait();      // see zero2_kernel: explicit ordering when launched into a predecessor's tail
}

// Zero fill of two float ranges (na, nb multiples of 4, 16-byte aligned bases) as a KERNEL, so that it can
// be launched with programmatic dependent launch into the tail of the persistent kernel before it (a
// cudaMemsetAsync node cannot).  Grid-stride.
__global__ void __launch_bounds__(256)
zero2_kernel(float4* __restrict__ a, size_t na4, float4* __restrict__ b, size_t nb4)
{
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < na4 + nb4; i += stride) {
        if (i < na4) a[i] = z;
        else b[i - na4] = z;
    }
    // launched into the tail of the kernel before it (programmatic dependent launch): this kernel does not
    // read that kernel's output, but the NEXT kernel in the stream does, and it only waits for this one.
    // Waiting here -- after our own stores -- makes "this grid complete" imply "the predecessor complete
    // and flushed" by the PTX rules instead of by driver behaviour; the overlap with the tail is kept.
    griddep_wait();
}

// LeakyReLU backward (model.py:84) as a stand-alone pass: dst = grad_out * (out < 0 ? slope : 1).
// Feeds the TMA correlation-backward kernels, whose taps are staged by asynchronous copies and cannot
// be gated on the way.  n4 = number of float4 elements.
__global__ void __launch_bounds__(256)
gate_grad_kernel(const float4* __restrict__ gout, const float4* __restrict__ out, float4* __restrict__ dst,
                 size_t n4, float slope, size_t per_image4, size_t gbs4, size_t obs4)
{
    // dst is dense; gout / out may carry a batch stride (per_image4, gbs4, obs4 in float4 units)
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const size_t n = i / per_image4, r = i - n * per_image4;
    float4 g = __ldg(gout + n * gbs4 + r);
    const float4 o = __ldg(out + n * obs4 + r);
    if (o.x < 0.0f) g.x *= slope;
    if (o.y < 0.0f) g.y *= slope;
    if (o.z < 0.0f) g.z *= slope;
    if (o.w < 0.0f) g.w *= slope;
    dst[i] = g;
}

}  // namespace pwc
