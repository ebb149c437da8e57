// Tiled correlation backward for kernel_size == 1, stride1 == 1, pad_size == max_displacement.
//
// Replaces correlation_cuda.c:113-121 (4 fills), correlation_cuda_kernel.cu:435-436 (2 layout
// copies) and the 2*B per-item launches of :441-463 with two launches (one per gradient):
//
//   g1[n,c,y,x] = 1/C * sum_d gO[n,d,y,x]             * X2[n,c,y+dy,x+dx]      (SIGN = +1)
//   g2[n,c,y,x] = 1/C * sum_d gO[n,d,y-dy,x-dx]       * X1[n,c,y-dy,x-dx]      (SIGN = -1)
//   (dx,dy) = ((d%D - r)*S2, (d/D - r)*S2); terms whose pixel falls outside the image are 0.
//
// One thread per pixel of a tile (8x32, or 16x16 / 8x8 for small images) keeps its D*D gradient taps G[d] in registers (read once,
// coalesced); per channel chunk the CTA stages the tile+halo of the other operand in shared
// memory and every thread reduces over the D*D displacements for each channel.
#pragma once
#include "pwc_common.cuh"

namespace pwc {

template <int D_, int S2_, int CK_, int TW_ = 32, int TH_ = 8>
struct BwdCfg {
    static constexpr int D = D_, S2 = S2_, CK = CK_;
    static constexpr int r = (D - 1) / 2, R = r * S2;
    // one thread per tile pixel; small tiles (8x8, 16x16) keep the 6x7 / 12x14 pyramid levels from
    // running mostly-idle 256-thread CTAs
    static constexpr int TW = TW_, TH = TH_, NT = TW * TH;
    static constexpr int HH = TH + 2 * R, HWD = TW + 2 * R;
    static constexpr int HP = HWD | 1;   // odd pitch: rows of a warp never collide
    static constexpr size_t smem_bytes() { return sizeof(float) * (size_t)CK * HH * HP; }
};

template <class Cfg, int SIGN>
__global__ void __launch_bounds__(Cfg::NT)
corr_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ gate,
                const float* __restrict__ X, float* __restrict__ res,
                int C, int H, int W, int tiles_x, int tiles_y, int cgroup, float slope, long long gbs,
                long long gate_bs)
{
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, r = Cfg::r, R = Cfg::R;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, NT = Cfg::NT, HH = Cfg::HH, HP = Cfg::HP;
    extern __shared__ __align__(16) float sX[];

    const int tid = threadIdx.x, lx = tid % TW, ly = tid / TW;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int n = t / tiles_y;
    const int y0t = ty * TH, x0t = tx * TW;
    const int y = y0t + ly, x = x0t + lx;
    const bool inside = (y < H) && (x < W);
    const size_t HW = (size_t)H * W;
    const float* gon = gout + (size_t)n * (size_t)gbs;          // batch strides of the output gradient / the gate
    const float* gaten = gate ? gate + (size_t)n * (size_t)gate_bs : nullptr;
    const float* Xn = X + (size_t)n * C * HW;
    float* resn = res + (size_t)n * C * HW;

    float G[D * D];
#pragma unroll
    for (int d = 0; d < D * D; ++d) {
        const int dy = (d / D - r) * S2, dx = (d % D - r) * S2;
        const int gy = (SIGN > 0) ? y : y - dy, gx = (SIGN > 0) ? x : x - dx;
        float g = 0.0f;
        if (inside && gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = (size_t)d * HW + (size_t)gy * W + gx;
            g = __ldg(gon + o);
            if (gaten && __ldg(gaten + o) < 0.0f) g *= slope;
        }
        G[d] = g;
    }

    // output channels are independent: blockIdx.y selects a group of `cgroup` channels, which gives
    // small images (6x7, 12x14 pyramid levels) enough CTAs to fill the machine
    const float nelems = (float)C;
    const int c_begin = blockIdx.y * cgroup;
    const int c_end = min(C, c_begin + cgroup);
    // The halo tile is zero outside the image: zero the buffer once, then every chunk only rewrites the
    // rectangle that intersects the image (for the 6x7 / 12x14 levels that is a small part of the tile).
    for (int i = tid; i < CK * HH * HP; i += NT) sX[i] = 0.0f;
    const int ry0 = max(0, y0t - R), ry1 = min(H, y0t + TH + R);
    const int rx0 = max(0, x0t - R), rx1 = min(W, x0t + TW + R);
    const int rw = rx1 - rx0, rh = ry1 - ry0, rarea = rw * rh;
    for (int c0 = c_begin; c0 < c_end; c0 += CK) {
        __syncthreads();
        const int nch = min(CK, c_end - c0);
        for (int i = tid; i < nch * rarea; i += NT) {
            const int c = i / rarea, rem = i - c * rarea;
            const int ry = rem / rw, rx = rem - ry * rw;
            const int yy = ry0 + ry, xx = rx0 + rx;
            sX[c * (HH * HP) + (yy - (y0t - R)) * HP + (xx - (x0t - R))] =
                __ldg(Xn + (size_t)(c0 + c) * HW + (size_t)yy * W + xx);
        }
        __syncthreads();
        const float* base = sX + (ly + R) * HP + (lx + R);
        for (int c = 0; c < CK; ++c) {
            if (c0 + c >= c_end) break;
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
#pragma unroll
            for (int d = 0; d < D * D; d += 3) {
                {
                    const int dy = (d / D - r) * S2, dx = (d % D - r) * S2;
                    s0 = fmaf(G[d], base[c * (HH * HP) + SIGN * (dy * HP + dx)], s0);
                }
                if (d + 1 < D * D) {
                    const int dy = ((d + 1) / D - r) * S2, dx = ((d + 1) % D - r) * S2;
                    s1 = fmaf(G[d + 1], base[c * (HH * HP) + SIGN * (dy * HP + dx)], s1);
                }
                if (d + 2 < D * D) {
                    const int dy = ((d + 2) / D - r) * S2, dx = ((d + 2) % D - r) * S2;
                    s2 = fmaf(G[d + 2], base[c * (HH * HP) + SIGN * (dy * HP + dx)], s2);
                }
            }
            if (inside) resn[(size_t)(c0 + c) * HW + (size_t)y * W + x] = (s0 + s1 + s2) / nelems;
        }
    }
}

}  // namespace pwc
