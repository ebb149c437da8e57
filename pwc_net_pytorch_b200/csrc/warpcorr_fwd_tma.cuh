// Fused warp + correlation forward, TMA-staged variant (the main sm_100a kernel).
//
// Same arithmetic as warpcorr_fwd.cuh (see there for the reference lines it replaces); what
// changes is how operands reach shared memory:
//
//   * the f1 tile and a *source window* of f2 are fetched by TMA (cp.async.bulk.tensor.4d over a
//     [B][C][H][W] tensor map; out-of-image and beyond-C elements arrive as zeros, so no border
//     or channel-tail code exists in the main loop), multi-buffered over channel chunks and
//     tracked with mbarriers -- no global load sits in the steady-state loop;
//   * the window is placed per CTA from the bounding box of the tile+halo sample positions
//     (x+u, y+v), so large smooth flows cost nothing; a sample whose 2x2 footprint falls outside
//     the window (rare outlier) is gathered from global memory instead;
//   * the CTA is warp-specialised: 3 producer warps issue the TMA copies and evaluate the bilinear
//     warp out of the window into a double-buffered warped tile; 9 consumer warps (warp wd owns
//     displacement row tj = wd - 4, lane (lr, ls) owns an 8-pixel strip and its 8x9 accumulators)
//     correlate out of shared memory.  Producers and consumers meet only through mbarriers
//     (full/empty per buffer); there is no __syncthreads in the channel loop.
// Requires W % 4 == 0 and 16-byte aligned bases (TMA global strides are multiples of 16 bytes);
// every TMA start coordinate along x is kept a multiple of 4 pixels (see TmaCfg::WW).
#pragma once
#include <cuda.h>

#include "pwc_common.cuh"

namespace pwc {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int x,
                                            int y, int c, int n)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(n)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producer_sync(int nthreads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <int S2_, int CK_>
struct TmaCfg {
    static constexpr int D = 9, S2 = S2_, CK = CK_, PX = 8;
    static constexpr int r = 4, R = r * S2;
    static constexpr int TW = 16, TH = 16;                 // output tile: 2 strips x 16 rows = 32 lanes
    static constexpr int NCONS = 32 * D;                   // consumer threads: one warp per tj
    static constexpr int NPROD = 96;                       // producer threads: TMA issue + bilinear warp
    static constexpr int NT = NCONS + NPROD;               // 12 warps
    static constexpr int HH = TH + 2 * R, HWD = TW + 2 * R;   // warped tile + halo
    static constexpr int WP = HWD + 4;                     // pitch = 4 (mod 8): conflict-free 128-bit rows
    static constexpr int MARGIN = 8;                       // extra source pixels each side of the halo
    // f2 source window (TMA box).  +4 columns: the innermost TMA start coordinate must be a multiple
    // of 16 bytes (measured on B200: a 4-D tiled fp32 load at x % 4 != 0 raises "illegal
    // instruction"), so the window origin is rounded down to a multiple of 4 pixels.
    static constexpr int WW = HWD + 2 * MARGIN + 4, WH = HH + 2 * MARGIN;
    static constexpr int F1W = TW + 4, F1H = TH;           // f1 box, pitch 20 = 4 (mod 8)
    static constexpr int NHALO = HH * HWD;
    static constexpr int WSPAN = PX + 2 * R;
    static constexpr int NWIN = 2, NF1 = 3, NW2F = 2, NW2P = 3;   // ring depths (flow / plain)
    static constexpr int WIN_ELEMS = CK * WH * WW;
    static constexpr int F1_ELEMS = CK * F1H * F1W;
    static constexpr int W2_ELEMS = CK * HH * WP;
    static constexpr uint32_t WIN_BYTES = WIN_ELEMS * 4, F1_BYTES = F1_ELEMS * 4, W2_BYTES = W2_ELEMS * 4;
    static_assert(WP % 8 == 4 && F1W % 8 == 4, "pitches must be 4 mod 8 floats");
    static_assert((WW * 4) % 16 == 0 && (F1W * 4) % 16 == 0 && (WP * 4) % 16 == 0, "TMA box rows are 16B multiples");
    static_assert(WIN_BYTES % 128 == 0 && F1_BYTES % 128 == 0 && W2_BYTES % 128 == 0, "buffers stay 128B aligned");

    static constexpr size_t smem_bytes(bool has_flow)
    {
        size_t b = 256;   // mbarriers + bbox scratch
        b += (size_t)NF1 * F1_BYTES;
        if (has_flow) {
            b += (size_t)NWIN * WIN_BYTES + (size_t)NW2F * W2_BYTES;
            b += (size_t)NHALO * (sizeof(float4) + sizeof(int));
        } else {
            b += (size_t)NW2P * W2_BYTES;
        }
        return b + 128;   // slack for the manual 128-byte alignment of the dynamic segment
    }
};

constexpr int TAP_EMPTY = -1;      // every corner outside the image (or non-finite flow): value 0
constexpr int TAP_GLOBAL = -2;     // footprint outside the staged window: gather from global memory

template <class Cfg, bool HAS_FLOW>
__global__ void __launch_bounds__(Cfg::NT, 1)
warpcorr_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmF1, const __grid_constant__ CUtensorMap tmF2,
                        const float* __restrict__ f2, const float* __restrict__ flow,
                        float* __restrict__ out, float* __restrict__ warped_out,
                        int C, int H, int W, int tiles_x, int tiles_y, int act, float slope)
{
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, PX = Cfg::PX, R = Cfg::R;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, HWD = Cfg::HWD;
    constexpr int WP = Cfg::WP, WW = Cfg::WW, WH = Cfg::WH, F1W = Cfg::F1W, F1H = Cfg::F1H;
    constexpr int NHALO = Cfg::NHALO, WSPAN = Cfg::WSPAN, NCONS = Cfg::NCONS, NPROD = Cfg::NPROD;
    constexpr int NWIN = Cfg::NWIN, NF1 = Cfg::NF1, NW2 = HAS_FLOW ? Cfg::NW2F : Cfg::NW2P;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* barF1 = reinterpret_cast<uint64_t*>(base);           // [NF1]  TMA f1 landed
    uint64_t* barWin = barF1 + NF1;                                // [NWIN] TMA window landed
    uint64_t* barW2Full = barWin + NWIN;                           // [3]    warped tile ready
    uint64_t* barW2Empty = barW2Full + 3;                          // [3]    warped tile (and its f1) consumed
    int* bbox = reinterpret_cast<int*>(base + 128);                // minx, miny, maxx, maxy
    float* sF1 = reinterpret_cast<float*>(base + 256);
    float* sW2 = sF1 + NF1 * Cfg::F1_ELEMS;
    float* sWin = sW2 + NW2 * Cfg::W2_ELEMS;                       // HAS_FLOW only
    float4* sTapW = reinterpret_cast<float4*>(sWin + NWIN * Cfg::WIN_ELEMS);
    int* sTapM = reinterpret_cast<int*>(sTapW + NHALO);

    const int tid = threadIdx.x;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int n = t / tiles_y;
    const int y0t = ty * TH, x0t = tx * TW;
    const size_t HW = (size_t)H * W;
    const int nchunks = (C + CK - 1) / CK;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NF1; ++i) mbar_init(&barF1[i], 1);
#pragma unroll
        for (int i = 0; i < NWIN; ++i) mbar_init(&barWin[i], 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            mbar_init(&barW2Full[i], HAS_FLOW ? NPROD : 1);
            mbar_init(&barW2Empty[i], NCONS);
        }
        fence_mbar_init();
        bbox[0] = bbox[1] = 0x7fffffff;
        bbox[2] = bbox[3] = -0x7fffffff;
    }
    __syncthreads();     // the only block-wide barrier; roles split below

    if (tid >= NCONS) {
        // =========================== producer warps ===========================
        const int ptid = tid - NCONS;
        if (ptid == 0) {
            prefetch_tmap(&tmF1);
            prefetch_tmap(&tmF2);
            for (int k = 0; k < NF1 && k < nchunks; ++k) {
                mbar_expect_tx(&barF1[k], Cfg::F1_BYTES);
                tma_load_4d(sF1 + k * Cfg::F1_ELEMS, &tmF1, &barF1[k], x0t, y0t, k * CK, n);
            }
        }
        if (!HAS_FLOW) {
            // plain correlation: the f2 tile + halo is the "warped" tile; TMA writes it directly
            if (ptid == 0) {
                for (int k = 0; k < nchunks; ++k) {
                    const int b = k % NW2;
                    if (k >= NW2) {
                        mbar_wait(&barW2Empty[b], ((k / NW2) - 1) & 1);
                        mbar_expect_tx(&barF1[k % NF1], Cfg::F1_BYTES);
                        tma_load_4d(sF1 + (k % NF1) * Cfg::F1_ELEMS, &tmF1, &barF1[k % NF1], x0t, y0t, k * CK, n);
                    }
                    mbar_expect_tx(&barW2Full[b], Cfg::W2_BYTES);
                    tma_load_4d(sW2 + b * Cfg::W2_ELEMS, &tmF2, &barW2Full[b], x0t - R, y0t - R, k * CK, n);
                }
            }
            return;
        }

        // ---- pass 1: sample positions of tile + halo, their bounding box ----
        const float* un = flow + (size_t)n * 2 * HW;
        int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -0x7fffffff, mxy = -0x7fffffff;
        for (int i = ptid; i < NHALO; i += NPROD) {
            const int hy = i / HWD, hx = i - hy * HWD;
            const int y = y0t - R + hy, x = x0t - R + hx;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            int meta = TAP_EMPTY;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float u = __ldg(un + (size_t)y * W + x);
                const float v = __ldg(un + HW + (size_t)y * W + x);
                const float sx = (float)x + u, sy = (float)y + v;
                if (sx > -1.0f && sx < (float)W && sy > -1.0f && sy < (float)H) {
                    const float fx = floorf(sx), fy = floorf(sy);
                    const float ax = sx - fx, ay = sy - fy;
                    const int x0 = (int)fx, y0 = (int)fy;
                    w = make_float4((1.0f - ax) * (1.0f - ay), ax * (1.0f - ay), (1.0f - ax) * ay, ax * ay);
                    meta = ((y0 + 1) << 16) | (x0 + 1);     // x0, y0 >= -1; H, W < 32760 checked on the host
                    mnx = min(mnx, x0); mxx = max(mxx, x0 + 1);
                    mny = min(mny, y0); mxy = max(mxy, y0 + 1);
                }
            }
            sTapW[i] = w;
            sTapM[i] = meta;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if ((ptid & 31) == 0) {
            atomicMin(&bbox[0], mnx); atomicMin(&bbox[1], mny);
            atomicMax(&bbox[2], mxx); atomicMax(&bbox[3], mxy);
        }
        producer_sync(NPROD);
        // window origin: the bounding box if it fits, else centred on it (outliers go to global)
        int wx0 = 0, wy0 = 0;
        {
            const int bx0 = bbox[0], by0 = bbox[1], bx1 = bbox[2], by1 = bbox[3];
            if (bx0 <= bx1) {
                wx0 = bx0 & ~3;                                   // 16-byte aligned TMA start (also for x < 0)
                if (bx1 - wx0 + 1 > WW) wx0 = ((bx0 + bx1 + 1 - WW) >> 1) & ~3;
                wy0 = (by1 - by0 + 1 <= WH) ? by0 : (by0 + by1 + 1 - WH) / 2;
            }
        }
        if (ptid == 0) {
            for (int k = 0; k < NWIN && k < nchunks; ++k) {
                mbar_expect_tx(&barWin[k], Cfg::WIN_BYTES);
                tma_load_4d(sWin + k * Cfg::WIN_ELEMS, &tmF2, &barWin[k], wx0, wy0, k * CK, n);
            }
        }
        // ---- pass 2: positions -> window-relative offsets (each thread owns the same taps in every pass) ----
        for (int i = ptid; i < NHALO; i += NPROD) {
            const int meta = sTapM[i];
            if (meta >= 0) {
                const int x0 = (meta & 0xffff) - 1, y0 = (meta >> 16) - 1;
                const int rx = x0 - wx0, ry = y0 - wy0;
                sTapM[i] = (rx >= 0 && rx + 1 < WW && ry >= 0 && ry + 1 < WH) ? ry * WW + rx : TAP_GLOBAL;
            }
        }

        for (int k = 0; k < nchunks; ++k) {
            const int c0 = k * CK, b = k % NW2;
            if (k >= NW2) {
                // consumers are done with chunk k - NW2: its warped tile and its f1 buffer are free
                mbar_wait(&barW2Empty[b], ((k / NW2) - 1) & 1);
                const int kf = k + 1;     // f1 chunks 0..NF1-1 were issued up front
                if (ptid == 0 && kf >= NF1 && kf < nchunks) {
                    mbar_expect_tx(&barF1[kf % NF1], Cfg::F1_BYTES);
                    tma_load_4d(sF1 + (kf % NF1) * Cfg::F1_ELEMS, &tmF1, &barF1[kf % NF1], x0t, y0t, kf * CK, n);
                }
            }
            const float* win = sWin + (k % NWIN) * Cfg::WIN_ELEMS;
            float* w2buf = sW2 + b * Cfg::W2_ELEMS;
            mbar_wait(&barWin[k % NWIN], (k / NWIN) & 1);
            for (int i = ptid; i < NHALO; i += NPROD) {
                const int hy = i / HWD, hx = i - hy * HWD;
                const float4 w = sTapW[i];
                const int meta = sTapM[i];
                float v[CK];
                if (meta >= 0) {
                    const float* p = win + meta;
#pragma unroll
                    for (int c = 0; c < CK; ++c) {
                        const float* q = p + c * (WH * WW);
                        v[c] = fmaf(w.w, q[WW + 1], fmaf(w.z, q[WW], fmaf(w.y, q[1], w.x * q[0])));
                    }
                } else if (meta == TAP_EMPTY) {
#pragma unroll
                    for (int c = 0; c < CK; ++c) v[c] = 0.0f;
                } else {
                    // outlier: recompute the tap from the flow and gather from global memory
                    const int y = y0t - R + hy, x = x0t - R + hx;
                    const Tap tp = make_tap((float)x + __ldg(un + (size_t)y * W + x),
                                            (float)y + __ldg(un + HW + (size_t)y * W + x), H, W);
#pragma unroll
                    for (int c = 0; c < CK; ++c)
                        v[c] = (c0 + c < C && tp.off >= 0)
                                   ? tap_sample(tp, f2 + ((size_t)n * C + c0 + c) * HW) : 0.0f;
                }
                float* dst = w2buf + hy * WP + hx;
#pragma unroll
                for (int c = 0; c < CK; ++c) dst[c * (HH * WP)] = v[c];
                if (warped_out != nullptr) {     // x2_warp export (model.py:107,113)
                    const int gy = y0t - R + hy, gx = x0t - R + hx;
                    if (hy >= R && hy < R + TH && hx >= R && hx < R + TW && gy < H && gx < W) {
                        float* wo = warped_out + ((size_t)n * C + c0) * HW + (size_t)gy * W + gx;
#pragma unroll
                        for (int c = 0; c < CK; ++c)
                            if (c0 + c < C) wo[(size_t)c * HW] = v[c];
                    }
                }
            }
            mbar_arrive(&barW2Full[b]);          // release: this thread's part of warped tile k is written
            if (k + NWIN < nchunks) {
                producer_sync(NPROD);            // every producer has finished reading window k
                if (ptid == 0) {
                    const int kk = k + NWIN;
                    mbar_expect_tx(&barWin[kk % NWIN], Cfg::WIN_BYTES);
                    tma_load_4d(sWin + (kk % NWIN) * Cfg::WIN_ELEMS, &tmF2, &barWin[kk % NWIN], wx0, wy0, kk * CK, n);
                }
            }
        }
        return;
    }

    // =========================== consumer warps ===========================
    const int lane = tid & 31, wd = tid >> 5;     // wd: displacement row, tj = wd - r
    const int lr = lane & 15, ls = lane >> 4;     // rows fastest: a quarter warp spans 8 rows of one strip
    float acc[PX][D];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int d = 0; d < D; ++d) acc[p][d] = 0.0f;

    for (int k = 0; k < nchunks; ++k) {
        const int b = k % NW2;
        mbar_wait(&barF1[k % NF1], (k / NF1) & 1);
        mbar_wait(&barW2Full[b], (k / NW2) & 1);
        const float* pf = sF1 + (k % NF1) * Cfg::F1_ELEMS + lr * F1W + ls * PX;
        const float* pw = sW2 + b * Cfg::W2_ELEMS + (lr + wd * S2) * WP + ls * PX;
#pragma unroll
        for (int c = 0; c < CK; ++c) {
            float f[PX];
#pragma unroll
            for (int q = 0; q < PX / 4; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(pf + c * (F1H * F1W) + 4 * q);
                f[4 * q] = v4.x; f[4 * q + 1] = v4.y; f[4 * q + 2] = v4.z; f[4 * q + 3] = v4.w;
            }
#pragma unroll
            for (int q = 0; q < WSPAN / 4; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(pw + c * (HH * WP) + 4 * q);
                const float wq[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = 4 * q + e;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const int p = j - d * S2;
                        if (p >= 0 && p < PX) acc[p][d] = fmaf(f[p], wq[e], acc[p][d]);
                    }
                }
            }
        }
        mbar_arrive(&barW2Empty[b]);     // chunk k (warped tile b and f1 buffer k % NF1) consumed
    }

    // ---- epilogue: 1/C (correlation_cuda_kernel.cu:65,100), optional LeakyReLU (model.py:84) ----
    const int y = y0t + lr;
    const int xs = x0t + ls * PX;
    if (y < H && xs < W) {     // W % 4 == 0 and xs % 8 == 0: a strip is fully inside or ends on a multiple of 4
        const float nelems = (float)C;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float* o = out + (((size_t)n * (D * D) + (wd * D + d)) * H + y) * W + xs;
            float v[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                v[p] = acc[p][d] / nelems;
                if (act) v[p] = leaky(v[p], slope);
            }
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            if (xs + 4 < W) *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

}  // namespace pwc
