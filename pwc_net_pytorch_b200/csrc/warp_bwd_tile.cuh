// Backward of the bilinear warp, tiled (sm_100a): grad w.r.t. f2 (scatter) and w.r.t. the flow, plus the
// warped features themselves for the g1 kernel.  Replaces ATen's grid_sample backward + the autograd of
// modules.py:36-40.
//
// Same arithmetic as warp_bwd_v8_kernel (generic_kernels.cuh); what changes is where the four bilinear
// corner values come from.  There every lane gathered them from global memory: ~2 L1 sector requests
// per (pixel, channel), 65 % of the kernel's LSU wavefronts.  Here a CTA owns a 16x16 pixel tile, the
// source window of f2 that its samples fall into is fetched by TMA (placed from the bounding box of the
// sample positions, as in the forward kernel; 4 channels per ring slot), and the corners are read from
// shared memory.  A thread owns one pixel for all channels, so the flow gradient is a register sum that
// is stored once (no atomics, no memset); the feature gradient still goes to the 8-channel-interleaved
// scratch with red.global.add.v4.f32 (two channel quads of a pixel share a 32-byte sector).
//
//   T 1 warp : window TMA requests (ring of NWIN), one per 4-channel chunk, as soon as the tile's window
//              origin is known.
//   B 8 warps: tap from the flow (make_tap: same code as every other kernel), bounding box (named
//              barrier among the B warps), then per chunk 16 LDS + 4 coalesced LDG (gradient w.r.t. the
//              warped features) -> x2_warp store, flow-gradient terms, 4 vector reductions.
// Persistent, several CTAs per SM (the ring is small), W % 4 == 0 and 16-byte aligned bases (TMA).
#pragma once
#include "warpcorr_fwd_tma.cuh"

namespace pwc {

struct WarpBwdCfg {
    static constexpr int CK = 4, TW = 16, TH = 16, NPX = TW * TH;
    static constexpr int NB = NPX, NT = NB + 32;             // 8 B warps + T warp
    static constexpr int MARGIN = 8;
    static constexpr int WW = TW + 1 + 2 * MARGIN + 3, WH = TH + 1 + 2 * MARGIN;   // 36 x 33 (footprints are 2 x 2)
    static constexpr int NWIN = 3;
    static constexpr int WIN_ELEMS = CK * WH * WW;                   // TMA box
    static constexpr uint32_t WIN_BYTES = WIN_ELEMS * 4;
    static constexpr int SLOT_ELEMS = (WIN_ELEMS + 31) / 32 * 32;    // ring slots stay 128-byte aligned (TMA destination)
    static constexpr int CTRL_BYTES = 384;                           // 8 mbarriers, window origins, 2 x 8 warp boxes
    static_assert(8 * 8 + 4 * 4 + 2 * 8 * 4 * 4 <= CTRL_BYTES, "control block too small");
    static_assert((WW * 4) % 16 == 0, "TMA box rows are 16B multiples");
    static constexpr size_t smem_bytes() { return CTRL_BYTES + (size_t)NWIN * SLOT_ELEMS * 4; }
};

__global__ void __launch_bounds__(WarpBwdCfg::NT, 3)
warp_bwd_tile_kernel(const __grid_constant__ CUtensorMap tmF2, const float* __restrict__ gwarp,
                     const float* __restrict__ f2, const float* __restrict__ flow, float* __restrict__ gx8,
                     float* __restrict__ gflow, float* __restrict__ warped_out, int C, int H, int W, int tiles_x,
                     int tiles_y, int ntiles, int cocts)
{
    using Cfg = WarpBwdCfg;
    constexpr int CK = Cfg::CK, TW = Cfg::TW, TH = Cfg::TH, WW = Cfg::WW, WH = Cfg::WH, NWIN = Cfg::NWIN, NB = Cfg::NB;

    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* barWin = reinterpret_cast<uint64_t*>(base);     // [NWIN] TMA: window chunk landed      (T -> B)
    uint64_t* barWinFree = barWin + NWIN;                     // [NWIN] window chunk consumed          (B -> T)
    uint64_t* barOrg = barWinFree + NWIN;                     // [2]    window origin of a tile known  (B -> T)
    int* worg = reinterpret_cast<int*>(barOrg + 2);           // [2][2] window origin per tile parity
    int* sbox = worg + 4;                                     // [8][4] per-warp bounding boxes
    float* sWin = reinterpret_cast<float*>(base + Cfg::CTRL_BYTES);

    const int tid = threadIdx.x;
    const int HWi = H * W;
    const size_t HW = (size_t)HWi;
    const int nchunks = (C + CK - 1) / CK;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * nchunks;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NWIN; ++i) {
            mbar_init(&barWin[i], 1);
            mbar_init(&barWinFree[i], NB);
        }
        mbar_init(&barOrg[0], 1);
        mbar_init(&barOrg[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NB) {
        // ================================ T: window TMA ================================
        if (tid != NB) return;
        prefetch_tmap(&tmF2);
        int wx0 = 0, wy0 = 0;
        for (int g = 0; g < total; ++g) {
            const int lt = g / nchunks, k = g - lt * nchunks, slot = g % NWIN;
            if (k == 0) {
                mbar_wait(&barOrg[lt & 1], (lt >> 1) & 1);
                wx0 = worg[2 * (lt & 1)];
                wy0 = worg[2 * (lt & 1) + 1];
            }
            if (g >= NWIN) mbar_wait(&barWinFree[slot], ((g / NWIN) - 1) & 1);
            const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
            mbar_expect_tx(&barWin[slot], Cfg::WIN_BYTES);
            tma_load_4d(sWin + slot * Cfg::SLOT_ELEMS, &tmF2, &barWin[slot], wx0, wy0, k * CK, tc.n);
        }
        return;
    }

    // ================================ B: one pixel per thread ================================
    const int lane = tid & 31, wid = tid >> 5;
    const int ly = tid >> 4, lx = tid & 15;                   // a warp covers two tile rows
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
        const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
        const int y = tc.y0 + ly, x = tc.x0 + lx;
        const bool inside = y < H && x < W;
        const int pix = y * W + x;
        float u = 0.0f, v = 0.0f;
        if (inside) {
            u = __ldg(flow + (size_t)tc.n * 2 * HW + pix);
            v = __ldg(flow + (size_t)tc.n * 2 * HW + HW + pix);
        }
        float ax = 0.0f, ay = 0.0f;
        int x0 = 0, y0 = 0;
        Tap t = make_tap(x, y, u, v, H, W, &ax, &ay, &x0, &y0);
        if (!inside) t.off = -1;
        // ---- bounding box of the 2x2 footprints (top-left corners), window origin ----
        int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -0x7fffffff, mxy = -0x7fffffff;
        if (t.off >= 0) { mnx = mxx = x0; mny = mxy = y0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        const int par = lt & 1;
        if (lane == 0) {
            int* sb = sbox + (par * 8 + wid) * 4;
            sb[0] = mnx; sb[1] = mny; sb[2] = mxx; sb[3] = mxy;
        }
        producer_sync(NB);                                    // bar.sync 1 among the B warps
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int* sb = sbox + (par * 8 + w) * 4;
            mnx = min(mnx, sb[0]); mny = min(mny, sb[1]); mxx = max(mxx, sb[2]); mxy = max(mxy, sb[3]);
        }
        int wx0 = 0, wy0 = 0;
        if (mnx <= mxx) {
            wx0 = mnx & ~3;                                   // 16-byte aligned TMA start (also for x < 0)
            if (mxx + 1 - wx0 + 1 > WW) wx0 = ((mnx + mxx + 2 - WW) >> 1) & ~3;
            wy0 = (mxy + 1 - mny + 1 <= WH) ? mny : (mny + mxy + 2 - WH) / 2;
        }
        if (tid == 0) {
            worg[2 * par] = wx0;
            worg[2 * par + 1] = wy0;
            mbar_arrive(&barOrg[par]);                        // release: the T warp may read the origin
        }
        // window-relative offset of this pixel's footprint, or -1: gather from global memory (outlier)
        const int rx = x0 - wx0, ry = y0 - wy0;
        const bool in_win = rx >= 0 && rx + 1 < WW && ry >= 0 && ry + 1 < WH;
        const int woff = in_win ? ry * WW + rx : -1;
        const bool inx0 = x0 >= 0, inx1 = x0 + 1 < W, iny0 = y0 >= 0, iny1 = y0 + 1 < H;
        const float m00 = (inx0 && iny0) ? 1.0f : 0.0f, m01 = (inx1 && iny0) ? 1.0f : 0.0f;
        const float m10 = (inx0 && iny1) ? 1.0f : 0.0f, m11 = (inx1 && iny1) ? 1.0f : 0.0f;
        const float bx = 1.0f - ax, by = 1.0f - ay;
        float gu = 0.0f, gv = 0.0f;

        // gradient w.r.t. the warped features: coalesced loads, fetched one chunk ahead of their use
        float gnext[CK];
#pragma unroll
        for (int c = 0; c < CK; ++c)
            gnext[c] = (t.off >= 0 && c < C) ? __ldg(gwarp + ((size_t)tc.n * C + c) * HW + pix) : 0.0f;
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++g) {
            const int slot = g % NWIN, c0 = k * CK;
            float gq[CK];
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                gq[c] = gnext[c];
                const int cn = c0 + CK + c;
                gnext[c] = (t.off >= 0 && cn < C) ? __ldg(gwarp + ((size_t)tc.n * C + cn) * HW + pix) : 0.0f;
            }
            mbar_wait(&barWin[slot], (g / NWIN) & 1);
            if (k > 0) mbar_arrive(&barWinFree[(g - 1) % NWIN]);     // previous chunk: all its loads have issued and landed
            if (t.off >= 0) {
                float wv[CK];
#pragma unroll
                for (int c = 0; c < CK; ++c) {
                    float v00, v01, v10, v11;
                    if (woff >= 0) {          // zero fill of the TMA box = grid_sample's zero padding
                        const float* q = sWin + slot * Cfg::SLOT_ELEMS + c * (WH * WW) + woff;
                        v00 = q[0]; v01 = q[1]; v10 = q[WW]; v11 = q[WW + 1];
                    } else if (c0 + c < C) {
                        const float* p = f2 + ((size_t)tc.n * C + c0 + c) * HW + t.off;
                        v00 = m00 * __ldg(p); v01 = m01 * __ldg(p + t.dx);
                        v10 = m10 * __ldg(p + t.dyw); v11 = m11 * __ldg(p + t.dyw + t.dx);
                    } else {
                        v00 = v01 = v10 = v11 = 0.0f;
                    }
                    gu = fmaf(gq[c], fmaf(v11 - v10, ay, (v01 - v00) * by), gu);
                    gv = fmaf(gq[c], fmaf(v11 - v01, ax, (v10 - v00) * bx), gv);
                    wv[c] = fmaf(t.w11, v11, fmaf(t.w10, v10, fmaf(t.w01, v01, t.w00 * v00)));
                }
                if (warped_out) {
#pragma unroll
                    for (int c = 0; c < CK; ++c)
                        if (c0 + c < C) warped_out[((size_t)tc.n * C + c0 + c) * HW + pix] = wv[c];
                }
                float* qd = gx8 + (((size_t)tc.n * cocts + (k >> 1)) * HW + t.off) * 8 + 4 * (k & 1);
                const size_t sdx = (size_t)t.dx * 8, sdy = (size_t)t.dyw * 8;
                if (t.w00 != 0.0f)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(qd), "f"(gq[0] * t.w00), "f"(gq[1] * t.w00),
                                 "f"(gq[2] * t.w00), "f"(gq[3] * t.w00) : "memory");
                if (t.w01 != 0.0f)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(qd + sdx), "f"(gq[0] * t.w01), "f"(gq[1] * t.w01),
                                 "f"(gq[2] * t.w01), "f"(gq[3] * t.w01) : "memory");
                if (t.w10 != 0.0f)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(qd + sdy), "f"(gq[0] * t.w10), "f"(gq[1] * t.w10),
                                 "f"(gq[2] * t.w10), "f"(gq[3] * t.w10) : "memory");
                if (t.w11 != 0.0f)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(qd + sdy + sdx), "f"(gq[0] * t.w11),
                                 "f"(gq[1] * t.w11), "f"(gq[2] * t.w11), "f"(gq[3] * t.w11) : "memory");
            } else if (inside && warped_out) {
#pragma unroll
                for (int c = 0; c < CK; ++c)
                    if (c0 + c < C) warped_out[((size_t)tc.n * C + c0 + c) * HW + pix] = 0.0f;
            }
        }
        // flow gradient: complete in registers (this thread saw every channel); the stores depend on every
        // window load of the tile, so the last chunk's slot may be released after them
        if (inside) {
            gflow[(size_t)tc.n * 2 * HW + pix] = gu;
            gflow[(size_t)tc.n * 2 * HW + HW + pix] = gv;
        }
        mbar_arrive(&barWinFree[(g - 1) % NWIN]);
    }
}

}  // namespace pwc
