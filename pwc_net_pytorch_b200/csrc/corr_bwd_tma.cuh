// Correlation backward on the forward kernel's TMA / mbarrier machinery (sm_100a).
//
// Replaces correlation_cuda.c:113-121 (4 fills), correlation_cuda_kernel.cu:435-436 (2 layout
// copies) and the 2*B per-item launches of :441-463 for kernel_size == 1, stride1 == 1,
// pad_size == max_displacement, 81 displacements.  One launch per gradient:
//
//   SIGN = +1 : g1[n,c,y,x] = 1/C * sum_d gO[n,d,y,x]       * X[n,c,y+dy,x+dx]      X = second operand
//   SIGN = -1 : g2[n,c,y,x] = 1/C * sum_d gO[n,d,y-dy,x-dx] * X[n,c,y-dy,x-dx]      X = first operand
//   (dx,dy) = ((d%9 - 4)*S2, (d/9 - 4)*S2); out-of-image terms are 0 (TMA zero fill / masked taps).
//
// Persistent CTA (one per SM) walking 16x16 tiles; work item = (tile, chunk of CK channels).
//   T  1 warp : TMA requests for the X tile + halo chunks (ring of NS).
//   C  9 warps: warp wd owns displacement row tj = wd - 4; lane (lr, ls) owns an 8-pixel strip and
//               keeps its 8x9 output-gradient taps in registers for the whole tile (read once from
//               global, LeakyReLU gate applied on the way); per channel 4 LDS.128 feed 72 FFMA and
//               leave 8 partial sums, written to the warp's own slice of a double-buffered
//               partial-sum buffer (no atomics: the result is deterministic).
//   R  2 warps: sum the 9 slices, scale by 1/C and store the gradient chunk (128-bit, full sectors).
// Roles meet only through mbarriers.  Requires W % 4 == 0 and 16-byte aligned bases (TMA).
#pragma once
#include "warpcorr_fwd_tma.cuh"

namespace pwc {

// TMA prefetch of a 4-D box into L2 (no shared memory involved): used to pull the next tile's
// output-gradient taps towards the SM while the current tile is being processed.
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int x, int y, int c, int n)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(x), "r"(y),
                 "r"(c), "r"(n)
                 : "memory");
}

template <int S2_, int CK_>
struct BwdTmaCfg {
    static constexpr int D = 9, S2 = S2_, CK = CK_, PX = 8;
    static constexpr int r = 4, R = r * S2;
    static constexpr int TW = 16, TH = 16;
    static constexpr int NCONS = 32 * D, NRED = 64;
    static constexpr int NT = NCONS + NRED + 32;            // 12 warps
    static constexpr int HH = TH + 2 * R, HWD = TW + 2 * R;
    static constexpr int WP = HWD + 4;                      // X tile pitch (TMA box width), 4 mod 8
    static constexpr int PP = TW + 4;                       // partial-slice pitch, 4 mod 8
    static constexpr int WSPAN = PX + 2 * R;
    static constexpr int NS = 5;                            // X-chunk ring: deep enough to cover the TMA latency
    static constexpr int X_ELEMS = CK * HH * WP;
    static constexpr int SLICE_ELEMS = CK * TH * PP;        // one warp's partial sums for one chunk
    static constexpr int PART_ELEMS = D * SLICE_ELEMS;
    static constexpr uint32_t X_BYTES = X_ELEMS * 4;
    static constexpr int CTRL_BYTES = 256;
    static constexpr int GBOX_C = 27;                       // gradient-tap prefetch box: 27 of the 81 channels
    static_assert(WP % 8 == 4 && PP % 8 == 4, "pitches must be 4 mod 8 floats");
    static_assert(X_BYTES % 128 == 0 && (PART_ELEMS * 4) % 128 == 0, "buffers stay 128B aligned");
    static_assert((2 * NS + 4) * 8 <= CTRL_BYTES, "control block too small");
    static constexpr size_t smem_bytes() { return CTRL_BYTES + (size_t)NS * X_BYTES + 2 * (size_t)PART_ELEMS * 4; }
};

template <class Cfg, int SIGN>
__global__ void __launch_bounds__(Cfg::NT, 1)
corr_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                    const float* __restrict__ gout,
                    const float* __restrict__ gate, float* __restrict__ res,
                    int C, int H, int W, int tiles_x, int tiles_y, int ntiles, float slope)
{
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, PX = Cfg::PX, R = Cfg::R;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, WP = Cfg::WP, PP = Cfg::PP;
    constexpr int WSPAN = Cfg::WSPAN, NCONS = Cfg::NCONS, NRED = Cfg::NRED, NS = Cfg::NS;

    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* barFull = reinterpret_cast<uint64_t*>(base);   // [NS] TMA: X chunk landed       (T -> C)
    uint64_t* barEmpty = barFull + NS;                       // [NS] X chunk consumed           (C -> T)
    uint64_t* barPart = barEmpty + NS;                       // [2]  partial sums written       (C -> R)
    uint64_t* barPartFree = barPart + 2;                     // [2]  partial sums read          (R -> C)
    float* sX = reinterpret_cast<float*>(base + Cfg::CTRL_BYTES);
    float* sPart = sX + NS * Cfg::X_ELEMS;                   // [2][D][CK][TH][PP]

    const int tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    const int nchunks = (C + CK - 1) / CK;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * nchunks;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            mbar_init(&barFull[i], 1);
            mbar_init(&barEmpty[i], NCONS);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&barPart[i], NCONS);
            mbar_init(&barPartFree[i], NRED);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NCONS + NRED) {
        // ================================ T: TMA issue ================================
        if (tid != NCONS + NRED) return;
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmG);
        for (int g = 0; g < total; ++g) {
            const int s = g % NS;
            const TileCoord tc = tile_coord(blockIdx.x + (g / nchunks) * gridDim.x, tiles_x, tiles_y, TH, TW);
            if (g % nchunks == 0 && g / nchunks + 1 < my_tiles) {
                // first chunk of a tile: pull the NEXT tile's output-gradient taps into L2, so that the
                // consumers' per-tile tap loads (their only global loads) hit L2 instead of DRAM
                const TileCoord tn = tile_coord(blockIdx.x + (g / nchunks + 1) * gridDim.x, tiles_x, tiles_y, TH, TW);
                const int off = (SIGN > 0) ? 0 : R;
#pragma unroll
                for (int j = 0; j < (D * D) / Cfg::GBOX_C; ++j)
                    tma_prefetch_4d(&tmG, tn.x0 - off, tn.y0 - off, j * Cfg::GBOX_C, tn.n);
            }
            if (g >= NS) mbar_wait(&barEmpty[s], ((g / NS) - 1) & 1);
            mbar_expect_tx(&barFull[s], Cfg::X_BYTES);
            tma_load_4d(sX + s * Cfg::X_ELEMS, &tmX, &barFull[s], tc.x0 - R, tc.y0 - R, (g % nchunks) * CK, tc.n);
        }
        return;
    }

    if (tid >= NCONS) {
        // ================================ R: reduce + store ================================
        const int rt = tid - NCONS;
        const float inv_nelems = __frcp_rn((float)C);      // 1/C, correlation_cuda_kernel.cu:194,286
        constexpr int QUADS = CK * TH * (TW / 4);          // 128-bit outputs per chunk
        for (int g = 0; g < total; ++g) {
            const int lt = g / nchunks, k = g - lt * nchunks, pb = g & 1;
            const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
            mbar_wait(&barPart[pb], (g >> 1) & 1);
            const float* part = sPart + pb * Cfg::PART_ELEMS;
#pragma unroll
            for (int j = 0; j < (QUADS + NRED - 1) / NRED; ++j) {
                const int q = rt + j * NRED;
                if (q < QUADS) {
                    const int c = q / (TH * (TW / 4)), rem = q - c * (TH * (TW / 4));
                    const int row = rem / (TW / 4), quad = rem - row * (TW / 4);
                    const float* p = part + c * (TH * PP) + row * PP + 4 * quad;
                    float4 a = *reinterpret_cast<const float4*>(p);
#pragma unroll
                    for (int wdx = 1; wdx < D; ++wdx) {
                        const float4 b = *reinterpret_cast<const float4*>(p + wdx * Cfg::SLICE_ELEMS);
                        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                    }
                    const int y = tc.y0 + row, x = tc.x0 + 4 * quad, cc = k * CK + c;
                    if (cc < C && y < H && x < W) {         // W % 4 == 0: a quad is inside or outside as a whole
                        a.x *= inv_nelems; a.y *= inv_nelems; a.z *= inv_nelems; a.w *= inv_nelems;
                        *reinterpret_cast<float4*>(res + ((size_t)tc.n * C + cc) * HW + (size_t)y * W + x) = a;
                    }
                }
            }
            mbar_arrive(&barPartFree[pb]);
        }
        return;
    }

    // ================================ C: partial sums ================================
    const int lane = tid & 31, wd = tid >> 5;
    const int lr = lane & 15, ls = lane >> 4;
    const int rowsel = (SIGN > 0) ? wd : (D - 1 - wd);     // X row offset inside the halo tile, in units of S2
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
        const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
        // ---- this thread's 8 x 9 output-gradient taps (registers for the whole tile) ----
        float G[PX][D];
        {
            // 128-bit loads: a strip starts on a multiple of 8 pixels and W % 4 == 0, so every aligned
            // quad is entirely inside or outside the row.  For SIGN < 0 the 8 wanted values start at
            // xs - dx; they are picked out of 2 or 3 aligned quads with compile-time indices.
            const int y = tc.y0 + lr, xs = tc.x0 + ls * PX;
            const int dy = (wd - 4) * S2;
            const int gy = (SIGN > 0) ? y : y - dy;
            const bool row_ok = (y < H) && gy >= 0 && gy < H;
            const float* gon = gout + (size_t)tc.n * (D * D) * HW + (size_t)(wd * D) * HW + (size_t)(row_ok ? gy : 0) * W;
            const float* gaten = gate ? gate + (size_t)tc.n * (D * D) * HW + (size_t)(wd * D) * HW + (size_t)(row_ok ? gy : 0) * W
                                      : nullptr;
            // phase 1: every aligned quad this thread needs, all requests in flight together (one L2 round
            // trip per tile instead of one per displacement column)
            constexpr int NQ = 3;
            float4 raw[D][NQ];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int dx = (SIGN > 0) ? 0 : (d - 4) * S2;
                const int sh = ((-dx) % 4 + 4) % 4;            // (xs - dx) mod 4, compile time
                const int xb = xs - dx - sh;                   // aligned start
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int x = xb + 4 * q;
                    raw[d][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if ((q < 2 || sh > 0) && row_ok && x >= 0 && x < W)
                        raw[d][q] = __ldg(reinterpret_cast<const float4*>(gon + (size_t)d * HW + x));
                }
            }
            // phase 2 (only with the LeakyReLU gate, model.py:84): scale by the sign of the forward output
            if (gaten) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const int dx = (SIGN > 0) ? 0 : (d - 4) * S2;
                    const int sh = ((-dx) % 4 + 4) % 4;
                    const int xb = xs - dx - sh;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int x = xb + 4 * q;
                        if ((q < 2 || sh > 0) && row_ok && x >= 0 && x < W) {
                            const float4 o = __ldg(reinterpret_cast<const float4*>(gaten + (size_t)d * HW + x));
                            if (o.x < 0.0f) raw[d][q].x *= slope;
                            if (o.y < 0.0f) raw[d][q].y *= slope;
                            if (o.z < 0.0f) raw[d][q].z *= slope;
                            if (o.w < 0.0f) raw[d][q].w *= slope;
                        }
                    }
                }
            }
            // phase 3: pick the 8 wanted values (compile-time shifts)
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int dx = (SIGN > 0) ? 0 : (d - 4) * S2;
                const int sh = ((-dx) % 4 + 4) % 4;
                const float t[4 * NQ] = {raw[d][0].x, raw[d][0].y, raw[d][0].z, raw[d][0].w, raw[d][1].x, raw[d][1].y,
                                         raw[d][1].z, raw[d][1].w, raw[d][2].x, raw[d][2].y, raw[d][2].z, raw[d][2].w};
#pragma unroll
                for (int p = 0; p < PX; ++p) G[p][d] = (xs + p < W) ? t[sh + p] : 0.0f;
            }
        }

        for (int k = 0; k < nchunks; ++k, ++g) {
            const int s = g % NS, pb = g & 1;
            mbar_wait(&barFull[s], (g / NS) & 1);
            const float* pw = sX + s * Cfg::X_ELEMS + (lr + rowsel * S2) * WP + ls * PX;
            float part[CK][PX];
#pragma unroll
            for (int c = 0; c < CK; ++c) {
#pragma unroll
                for (int p = 0; p < PX; ++p) part[c][p] = 0.0f;
#pragma unroll
                for (int q = 0; q < WSPAN / 4; ++q) {
                    const float4 v4 = *reinterpret_cast<const float4*>(pw + c * (HH * WP) + 4 * q);
                    const float wq[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * q + e;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            // X column (relative to the strip) used by displacement column d
                            const int col = (SIGN > 0) ? d * S2 : (D - 1 - d) * S2;
                            const int p = j - col;
                            if (p >= 0 && p < PX) part[c][p] = fmaf(G[p][d], wq[e], part[c][p]);
                        }
                    }
                }
            }
            mbar_arrive(&barEmpty[s]);                         // X chunk consumed
            if (g >= 2) mbar_wait(&barPartFree[pb], ((g >> 1) - 1) & 1);   // R has read this buffer's previous use
            float* dst = sPart + pb * Cfg::PART_ELEMS + wd * Cfg::SLICE_ELEMS + lr * PP + ls * PX;
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                *reinterpret_cast<float4*>(dst + c * (TH * PP)) = make_float4(part[c][0], part[c][1], part[c][2], part[c][3]);
                *reinterpret_cast<float4*>(dst + c * (TH * PP) + 4) = make_float4(part[c][4], part[c][5], part[c][6], part[c][7]);
            }
            mbar_arrive(&barPart[pb]);                         // release: this warp's slice of chunk g is written
        }
    }
}

}  // namespace pwc
