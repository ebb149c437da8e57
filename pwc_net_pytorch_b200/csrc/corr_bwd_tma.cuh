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
//   S  3 warps: stage the NEXT tile's output-gradient taps in shared memory (stage_tap_row: 16-byte
//               cp.async, or LDG.128 pairs + register shift + STS.128 for the shifted positions y-dy, x-dx,
//               zero fill outside the image) while the consumers work on the current tile, so the taps
//               never cost an exposed global round trip.
//   C  9 warps: warp wd owns displacement row tj = wd - 4; lane (lr, ls) owns an 8-pixel strip and
//               keeps its 8x9 output-gradient taps in registers for the whole tile (18 conflict-free
//               LDS.128 from the staging buffer); per channel 4 LDS.128 feed 72 FFMA and
//               leave 8 partial sums, written to the warp's own slice of a double-buffered
//               partial-sum buffer (no atomics: the result is deterministic).
//   The LeakyReLU gate (model.py:84) is applied to the output gradient by a separate elementwise pass
//   (gate_grad_kernel) before these kernels run.
//   R  3 warps: sum the 9 slices, scale by 1/C and store the gradient chunk (128-bit, full sectors).
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

// Ampere-style asynchronous copy global -> shared of BYTES (4, 8 or 16) per lane; src_bytes == 0 writes zeros.
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* dst, const void* src, int src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(smem_u32(dst)), "l"(src), "n"(BYTES), "r"(src_bytes)
                 : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Tap staging buffer: [81][16 rows][16 px], the four 16-byte chunks of a row XOR-swizzled with
// (row >> 1) & 3, so that the consumers' LDS.128 (8 rows x one chunk per quarter warp) are conflict-free.
__device__ __forceinline__ int tap_slot(int d, int row, int x)
{
    return d * 256 + row * 16 + ((((x >> 2) ^ (row >> 1)) & 3) << 2) + (x & 3);
}


// One displacement row (9 planes, dxi = 0..8) of output-gradient taps for one 16x16 tile -> staging
// buffer `dst` (9 x 256 floats, tap_slot layout).  plane(dxi)[row][x] = gO[dyi*9+dxi][y0+row-dy][x0+x-dx]
// (dx = dy = 0 for SIGN > 0), zero outside the image.  Executed by one warp.
//   * dx % 4 == 0: 16-byte cp.async straight from global memory (zero fill by src-size 0);
//   * otherwise  : two aligned LDG.128 per output quad, a compile-time register shift, STS.128.
//     (4-byte cp.async for these planes cost ~6x more LSU cycles and starved the consumers: measured.)
// All loads of the row are issued before the first store.  The caller signals completion with BOTH
// cp_async_mbar_arrive() and a normal mbar_arrive() (release for the STS).
template <int S2, int SIGN>
__device__ __forceinline__ void stage_tap_row(float* dst, const float* __restrict__ gon, int dyi, const TileCoord& tc,
                                              int H, int W, size_t HW, int lane, const float* safe)
{
    constexpr int D = 9;
    const int dy = (SIGN > 0) ? 0 : (dyi - 4) * S2;
    // lane -> two output quads per plane: (row, xq) = (lane >> 2, lane & 3) and (8 + (lane >> 2), lane & 3)
    const int r0 = lane >> 2, xq = lane & 3;
    float4 A[D][2], Bq[D][2];
#pragma unroll
    for (int dxi = 0; dxi < D; ++dxi) {
        const int dx = (SIGN > 0) ? 0 : (dxi - 4) * S2;
        const int sh = ((-dx) % 4 + 4) % 4;
        const float* plane = gon + (size_t)(dyi * D + dxi) * HW;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int row = r0 + 8 * q;
            const int sy = tc.y0 + row - dy;
            const int ab = tc.x0 + 4 * xq - dx - sh;                 // aligned source column of quad A
            const bool rok = sy >= 0 && sy < H;
            if (sh == 0) {
                const bool ok = rok && ab >= 0 && ab < W;            // W % 4 == 0: whole quads
                cp_async_zfill<16>(dst + tap_slot(dxi, row, 4 * xq), ok ? plane + (size_t)sy * W + ab : safe, ok ? 16 : 0);
            } else {
                A[dxi][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                Bq[dxi][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rok && ab >= 0 && ab < W) A[dxi][q] = __ldg(reinterpret_cast<const float4*>(plane + (size_t)sy * W + ab));
                if (rok && ab + 4 >= 0 && ab + 4 < W) Bq[dxi][q] = __ldg(reinterpret_cast<const float4*>(plane + (size_t)sy * W + ab + 4));
            }
        }
    }
#pragma unroll
    for (int dxi = 0; dxi < D; ++dxi) {
        const int dx = (SIGN > 0) ? 0 : (dxi - 4) * S2;
        const int sh = ((-dx) % 4 + 4) % 4;
        if (sh != 0) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 a = A[dxi][q], b = Bq[dxi][q];
                const float4 o = sh == 1 ? make_float4(a.y, a.z, a.w, b.x)
                               : sh == 2 ? make_float4(a.z, a.w, b.x, b.y)
                                         : make_float4(a.w, b.x, b.y, b.z);
                *reinterpret_cast<float4*>(dst + tap_slot(dxi, r0 + 8 * q, 4 * xq)) = o;
            }
        }
    }
}

template <int S2_, int CK_>
struct BwdTmaCfg {
    static constexpr int D = 9, S2 = S2_, CK = CK_, PX = 8;
    static constexpr int r = 4, R = r * S2;
    static constexpr int TW = 16, TH = 16;
    static constexpr int NCONS = 32 * D, NRED = 96, NSTAGE = 96;
    static constexpr int NT = NCONS + NRED + 32 + NSTAGE;   // 16 warps: C 9, R 3, T 1, S 3 (128 registers each)
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
    static constexpr int TAP_ELEMS = D * D * TH * TW;       // staged output-gradient taps of one tile
    static constexpr int GBOX_C = 27;                       // gradient-tap prefetch box: 27 of the 81 channels
    static_assert(WP % 8 == 4 && PP % 8 == 4, "pitches must be 4 mod 8 floats");
    static_assert(X_BYTES % 128 == 0 && (PART_ELEMS * 4) % 128 == 0, "buffers stay 128B aligned");
    static_assert((2 * NS + 6) * 8 <= CTRL_BYTES, "control block too small");
    static constexpr size_t smem_bytes() { return CTRL_BYTES + (size_t)NS * X_BYTES + 2 * (size_t)PART_ELEMS * 4 + (size_t)TAP_ELEMS * 4; }
};

template <class Cfg, int SIGN>
__global__ void __launch_bounds__(Cfg::NT, 1)
corr_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                    const float* __restrict__ gout, float* __restrict__ res,
                    int C, int H, int W, int tiles_x, int tiles_y, int ntiles, long long gbs)
{
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, PX = Cfg::PX, R = Cfg::R;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, WP = Cfg::WP, PP = Cfg::PP;
    constexpr int WSPAN = Cfg::WSPAN, NCONS = Cfg::NCONS, NRED = Cfg::NRED, NS = Cfg::NS;

    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* barFull = reinterpret_cast<uint64_t*>(base);   // [NS] TMA: X chunk landed       (T -> C)
    uint64_t* barEmpty = barFull + NS;                       // [NS] X chunk consumed           (C -> T)
    uint64_t* barPart = barEmpty + NS;                       // [2]  partial sums written       (C -> R)
    uint64_t* barPartFree = barPart + 2;                     // [2]  partial sums read          (R -> C)
    uint64_t* barTap = barPartFree + 2;                      // [1]  taps of a tile staged      (T cp.async -> C)
    uint64_t* barTapFree = barTap + 1;                       // [1]  taps copied to registers   (C -> T)
    float* sX = reinterpret_cast<float*>(base + Cfg::CTRL_BYTES);
    float* sPart = sX + NS * Cfg::X_ELEMS;                   // [2][D][CK][TH][PP]
    float* sTap = sPart + 2 * Cfg::PART_ELEMS;               // [81][TH][TW] swizzled (tap_slot)

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
        mbar_init(barTap, 2 * Cfg::NSTAGE);     // per staging thread: one cp.async arrival + one plain arrival
        mbar_init(barTapFree, NCONS);
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NCONS + NRED + 32) {
        // ================================ S: tap staging ================================
        const int lane = tid & 31, sw = (tid - (NCONS + NRED + 32)) >> 5;      // sw: 0..2, owns rows dyi = sw, sw+3, sw+6
        for (int lt = 0; lt < my_tiles; ++lt) {
            const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
            if (lt >= 1) mbar_wait(barTapFree, (lt - 1) & 1);      // tile lt-1's taps are in the consumers' registers
            const float* gon = gout + (size_t)tc.n * (size_t)gbs;      // gbs: batch stride of the output gradient
            for (int dyi = sw; dyi < D; dyi += Cfg::NSTAGE / 32)
                stage_tap_row<S2, SIGN>(sTap + dyi * D * (TH * TW), gon, dyi, tc, H, W, HW, lane, gout);
            cp_async_mbar_arrive(barTap);       // arrives once this thread's asynchronous copies have landed
            mbar_arrive(barTap);                // release: this thread's shifted quads are stored
        }
        return;
    }

    if (tid >= NCONS + NRED) {
        // ================================ T: TMA issue ================================
        if (tid != NCONS + NRED) return;
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmG);
        for (int g = 0; g < total; ++g) {
            const int s = g % NS;
            const TileCoord tc = tile_coord(blockIdx.x + (g / nchunks) * gridDim.x, tiles_x, tiles_y, TH, TW);
            if (g % nchunks == 0 && g / nchunks + 2 < my_tiles) {
                // pull the output-gradient region of the tile after next into L2 (TMA prefetch, no shared
                // memory): the S warps copy it one tile ahead of the consumers
                const TileCoord tn = tile_coord(blockIdx.x + (g / nchunks + 2) * gridDim.x, tiles_x, tiles_y, TH, TW);
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
        // ---- this thread's 8 x 9 output-gradient taps (registers for the whole tile) ----
        // G[p][d] = gO[n, wd*9+d, y, xs+p] (SIGN > 0) or gO[n, wd*9+d, y-dy, xs+p-dx] (SIGN < 0), zero outside
        // the image: staged by the T warp, already shifted, so both signs read the same two quads per d.
        float G[PX][D];
        mbar_wait(barTap, lt & 1);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float4 a = *reinterpret_cast<const float4*>(sTap + tap_slot(wd * D + d, lr, ls * PX));
            const float4 b = *reinterpret_cast<const float4*>(sTap + tap_slot(wd * D + d, lr, ls * PX + 4));
            G[0][d] = a.x; G[1][d] = a.y; G[2][d] = a.z; G[3][d] = a.w;
            G[4][d] = b.x; G[5][d] = b.y; G[6][d] = b.z; G[7][d] = b.w;
        }
        // (barTapFree is signalled after the first chunk's partial sums are stored: see below)

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
            if (g >= 2) mbar_wait(&barPartFree[pb], ((g >> 1) - 1) & 1);   // R has read this buffer's previous use
            float* dst = sPart + pb * Cfg::PART_ELEMS + wd * Cfg::SLICE_ELEMS + lr * PP + ls * PX;
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                *reinterpret_cast<float4*>(dst + c * (TH * PP)) = make_float4(part[c][0], part[c][1], part[c][2], part[c][3]);
                *reinterpret_cast<float4*>(dst + c * (TH * PP) + 4) = make_float4(part[c][4], part[c][5], part[c][6], part[c][7]);
            }
            mbar_arrive(&barPart[pb]);                         // release: this warp's slice of chunk g is written
            // The "consumed" signals come only now, after the stores above: those depend (through the FFMAs)
            // on every shared load of this chunk and on every tap register, so all of them have landed.
            // An mbarrier arrive issued right behind a still-pending LDS can overtake it, and the next TMA /
            // cp.async write then corrupts the value being read (measured: last LDS.128 of a chunk).
            mbar_arrive(&barEmpty[s]);                         // X chunk consumed
            if (k == 0) mbar_arrive(barTapFree);               // the tap staging buffer may be overwritten
        }
    }
}

}  // namespace pwc
