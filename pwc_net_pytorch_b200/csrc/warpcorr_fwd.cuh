// Fused bilinear backward-warp + local correlation (cost volume) + 1/C + LeakyReLU, forward.
//
// Replaces, in one launch, the reference sequence
//   modules.py:31-42 (WarpingLayer: ~10 elementwise launches + grid_sample)
//   -> correlation_cuda.c:36-42 (3 fills) -> correlation_cuda_kernel.cu:10-32 x2 (layout copies)
//   -> correlation_cuda_kernel.cu:34-106 (forward) -> model.py:84 (leaky_relu_)
// for the case kernel_size == 1, stride1 == 1, pad_size == max_displacement (both configs the
// model uses).  The warped feature map is never written to HBM
// unless the caller asks for it (warped_out, model.py:107,113).
//
//   out[n, (tj+r)*D + (ti+r), y, x] = act( 1/C * sum_c f1[n,c,y,x] * W2[n,c,y+tj*S2, x+ti*S2] )
//   W2[n,c,y',x'] = 0 outside the image, else bilinear_zero(f2[n,c], x'+u(y',x'), y'+v(y',x'))
//
// Work decomposition: one CTA per (image, TH x TW pixel tile); D warps, warp `wd` owns
// displacement row tj = wd - r; lane (lr, ls) owns the PX-pixel strip at tile row lr, column
// ls*PX, and keeps its PX x D accumulators in registers across the whole channel loop.  Per
// channel chunk the CTA stages the f1 tile and the *warped* f2 tile-plus-halo in shared memory
// (bilinear taps are computed once per CTA from the flow and reused for every channel).
#pragma once
#include <cooperative_groups.h>

#include "pwc_common.cuh"

namespace pwc {

// Smallest pitch >= width (multiple of 4 floats) for which the 8 lanes of a quarter warp, each
// issuing a 128-bit shared load at row*pitch + strip*PX, fall into 8 distinct 16-byte bank groups.
__host__ __device__ constexpr bool pitch_ok(int pitch, int px, int lw)
{
    int used = 0;
    for (int q = 0; q < 8; ++q) {
        const int row = q / lw, strip = q % lw;
        const int grp = ((row * pitch + strip * px) / 4) % 8;
        if (used & (1 << grp)) return false;
        used |= 1 << grp;
    }
    return true;
}
__host__ __device__ constexpr int pick_pitch(int width, int px, int lw)
{
    int p = round_up(width, 4);
    for (int k = 0; k < 16; ++k, p += 4)
        if (pitch_ok(p, px, lw)) return p;
    return round_up(width, 4);
}

template <int D_, int S2_, int PX_, int LW_, int CK_>
struct FwdCfg {
    static constexpr int D = D_, S2 = S2_, PX = PX_, LW = LW_, CK = CK_;
    static constexpr int r = (D - 1) / 2;
    static constexpr int R = r * S2;               // halo radius in pixels
    static constexpr int LR = 32 / LW;
    static constexpr int TW = LW * PX, TH = LR;    // output tile
    static constexpr int HH = TH + 2 * R, HWD = TW + 2 * R;
    static constexpr int HP = pick_pitch(HWD, PX, LW);   // warped-tile pitch
    static constexpr int FP = pick_pitch(TW, PX, LW);    // f1-tile pitch
    static constexpr int NT = 32 * D;
    static constexpr int NHALO = HH * HWD;
    static constexpr int WSPAN = PX + 2 * R;       // warped values one strip needs per row
    static constexpr int W2_ELEMS = CK * HH * HP;
    static constexpr int F1_ELEMS = CK * TH * FP;
    static constexpr int RED_ELEMS = PX * D * NT;      // split-K staging: one accumulator set per thread
    static constexpr int W2_ALLOC = W2_ELEMS > RED_ELEMS ? W2_ELEMS : RED_ELEMS;   // the staging aliases the warped tile
    static constexpr size_t smem_bytes(bool has_flow)
    {
        return sizeof(float) * (size_t)(W2_ALLOC + F1_ELEMS) +
               (has_flow ? (size_t)NHALO * (sizeof(float4) + sizeof(int2)) : 0);
    }
    static_assert(PX % 4 == 0, "strips are loaded with 128-bit shared loads");
    static_assert(32 % LW == 0, "lanes tile the strip grid");
};

template <class Cfg, bool HAS_FLOW>
__global__ void __launch_bounds__(Cfg::NT)
warpcorr_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                    const float* __restrict__ flow, float* __restrict__ out,
                    float* __restrict__ warped_out,
                    int C, int H, int W, int tiles_x, int tiles_y, int act, float slope, int ksplit, int cper,
                    long long obs, long long fbs)
{
    // ksplit > 1: the kernel is launched in thread-block clusters of ksplit CTAs; the CTAs of a cluster
    // share one tile and split its channels (cper each); partial accumulators are summed through
    // distributed shared memory in rank order, so the result stays deterministic.  This is what gives
    // the 6x7 / 12x14 pyramid levels (few tiles, 128-196 channels) enough CTAs to fill the machine.
    constexpr int D = Cfg::D, S2 = Cfg::S2, PX = Cfg::PX, LW = Cfg::LW, CK = Cfg::CK;
    constexpr int R = Cfg::R, TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, HWD = Cfg::HWD;
    constexpr int HP = Cfg::HP, FP = Cfg::FP, NT = Cfg::NT, NHALO = Cfg::NHALO;
    constexpr int WSPAN = Cfg::WSPAN;

    extern __shared__ __align__(16) float smem[];
    float* sW2 = smem;
    float* sF1 = sW2 + Cfg::W2_ALLOC;
    float4* sTapW = reinterpret_cast<float4*>(sF1 + Cfg::F1_ELEMS);
    int2* sTapO = reinterpret_cast<int2*>(sTapW + NHALO);

    const int tid = threadIdx.x, lane = tid & 31, wd = tid >> 5;
    const int ls = lane % LW, lr = lane / LW;
    const int rank = (ksplit > 1) ? (int)(blockIdx.x % ksplit) : 0;
    const int c_begin = rank * cper;
    const int c_end = min(C, c_begin + cper);
    int t = blockIdx.x / ksplit;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int n = t / tiles_y;
    const int y0t = ty * TH, x0t = tx * TW;
    const size_t HW = (size_t)H * W;
    const float* f1n = f1 + (size_t)n * C * HW;
    const float* f2n = f2 + (size_t)n * C * HW;

    if (HAS_FLOW) {
        const float* un = flow + (size_t)n * (size_t)fbs;      // fbs: flow batch stride
        for (int i = tid; i < NHALO; i += NT) {
            const int hy = i / HWD, hx = i - hy * HWD;
            const int y = y0t - R + hy, x = x0t - R + hx;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            int2 o = make_int2(-1, 0);
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float u = __ldg(un + (size_t)y * W + x);
                const float v = __ldg(un + HW + (size_t)y * W + x);
                const Tap tp = make_tap(x, y, u, v, H, W);
                w = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
                o = make_int2(tp.off, (tp.dyw << 1) | tp.dx);
            }
            sTapW[i] = w;
            sTapO[i] = o;
        }
    }

    float acc[PX][D];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int d = 0; d < D; ++d) acc[p][d] = 0.0f;

    for (int c0 = c_begin; c0 < c_end; c0 += CK) {
        __syncthreads();   // previous chunk fully consumed (and taps visible on the first pass)

        // ---- stage the f1 tile (zero outside the image / beyond C) ----
        for (int i = tid; i < CK * TH * TW; i += NT) {
            const int c = i / (TH * TW), rem = i - c * (TH * TW);
            const int ly = rem / TW, lx = rem - ly * TW;
            const int y = y0t + ly, x = x0t + lx;
            float v = 0.0f;
            if (c0 + c < c_end && y < H && x < W) v = __ldg(f1n + (size_t)(c0 + c) * HW + (size_t)y * W + x);
            sF1[c * (TH * FP) + ly * FP + lx] = v;
        }

        // ---- stage the warped f2 tile + halo ----
        if (HAS_FLOW) {
            for (int i = tid; i < NHALO; i += NT) {
                const int hy = i / HWD, hx = i - hy * HWD;
                const float4 w = sTapW[i];
                const int2 o = sTapO[i];
                float* dst = sW2 + hy * HP + hx;
                // x2_warp export (model.py:107,113): interior pixels of the tile that are in the image
                const int gy = y0t - R + hy, gx = x0t - R + hx;
                const bool interior = warped_out != nullptr && hy >= R && hy < R + TH &&
                                      hx >= R && hx < R + TW && gy < H && gx < W;
                float* wo = interior ? warped_out + ((size_t)n * C + c0) * HW + (size_t)gy * W + gx
                                     : nullptr;
                if (o.x < 0) {
#pragma unroll 8
                    for (int c = 0; c < CK; ++c) {
                        dst[c * (HH * HP)] = 0.0f;
                        if (wo && c0 + c < c_end) wo[(size_t)c * HW] = 0.0f;
                    }
                } else {
                    const int dx = o.y & 1, dyw = o.y >> 1;
                    const float* p00 = f2n + (size_t)c0 * HW + o.x;
#pragma unroll 8
                    for (int c = 0; c < CK; ++c) {
                        float v = 0.0f;
                        if (c0 + c < c_end) {
                            const float* p = p00 + (size_t)c * HW;
                            const float v00 = __ldg(p), v01 = __ldg(p + dx);
                            const float v10 = __ldg(p + dyw), v11 = __ldg(p + dyw + dx);
                            v = fmaf(w.w, v11, fmaf(w.z, v10, fmaf(w.y, v01, w.x * v00)));
                            if (wo) wo[(size_t)c * HW] = v;
                        }
                        dst[c * (HH * HP)] = v;
                    }
                }
            }
        } else {
            for (int i = tid; i < CK * NHALO; i += NT) {
                const int c = i / NHALO, rem = i - c * NHALO;
                const int hy = rem / HWD, hx = rem - hy * HWD;
                const int y = y0t - R + hy, x = x0t - R + hx;
                float v = 0.0f;
                if (c0 + c < c_end && y >= 0 && y < H && x >= 0 && x < W)
                    v = __ldg(f2n + (size_t)(c0 + c) * HW + (size_t)y * W + x);
                sW2[c * (HH * HP) + hy * HP + hx] = v;
            }
        }
        __syncthreads();

        // ---- correlate out of shared memory ----
        const float* pf = sF1 + lr * FP + ls * PX;
        const float* pw = sW2 + (lr + wd * S2) * HP + ls * PX;
#pragma unroll 2
        for (int c = 0; c < CK; ++c) {
            float f[PX], w[WSPAN];
#pragma unroll
            for (int q = 0; q < PX / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(pf + c * (TH * FP) + 4 * q);
                f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < WSPAN / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(pw + c * (HH * HP) + 4 * q);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int p = 0; p < PX; ++p)
#pragma unroll
                for (int d = 0; d < D; ++d) acc[p][d] = fmaf(f[p], w[p + d * S2], acc[p][d]);
        }
    }

    if (ksplit > 1) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        __syncthreads();                       // every warp is done reading the warped tile it aliases
        float* sRed = sW2;
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int d = 0; d < D; ++d) sRed[(p * D + d) * NT + tid] = acc[p][d];
        cluster.sync();
        if (rank == 0) {
            for (int r = 1; r < ksplit; ++r) {
                const float* remote = cluster.map_shared_rank(sRed, r);
#pragma unroll
                for (int p = 0; p < PX; ++p)
#pragma unroll
                    for (int d = 0; d < D; ++d) acc[p][d] += remote[(p * D + d) * NT + tid];
            }
        }
        cluster.sync();                        // remote shared memory stays valid until rank 0 has read it
        if (rank != 0) return;
    }

    // ---- epilogue: 1/C (correlation_cuda_kernel.cu:65,100), optional LeakyReLU (model.py:84) ----
    const int y = y0t + lr;
    if (y < H) {
        const float nelems = (float)C;
        const int xs = x0t + ls * PX;
        // 128-bit stores need every image's base 16-byte aligned: the tensor base AND the batch stride
        const bool vec = ((W & 3) == 0) && (xs + PX <= W) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                         ((obs & 3) == 0);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float* o = out + (size_t)n * (size_t)obs + ((size_t)(wd * D + d) * H + y) * W + xs;   // obs: output batch stride
            float v[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                v[p] = acc[p][d] / nelems;
                if (act) v[p] = leaky(v[p], slope);
            }
            if (vec) {
#pragma unroll
                for (int q = 0; q < PX / 4; ++q)
                    *reinterpret_cast<float4*>(o + 4 * q) =
                        make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            } else {
#pragma unroll
                for (int p = 0; p < PX; ++p)
                    if (xs + p < W) o[p] = v[p];
            }
        }
    }
}

}  // namespace pwc
