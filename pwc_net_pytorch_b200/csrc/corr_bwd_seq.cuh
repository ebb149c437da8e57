// Correlation backward, stride2 == 1, "sequential displacement rows" variant (sm_100a).
//
// Same arithmetic as corr_bwd_tma.cuh (the reference lines it replaces are listed there):
//   SIGN = +1 : g1[n,c,y,x] = 1/C * sum_d gO[n,d,y,x]       * X[n,c,y+dy,x+dx]      X = second operand
//   SIGN = -1 : g2[n,c,y,x] = 1/C * sum_d gO[n,d,y-dy,x-dx] * X[n,c,y-dy,x-dx]      X = first operand
// but a thread now owns COMPLETE outputs: an 8-pixel strip of 4 channels, accumulated over all 81
// displacements, so there is no cross-warp reduction, no partial-sum buffer and no reducer role.
// The displacement rows are visited one after the other; the 8x9 output-gradient taps of the current
// row live in registers (18 conflict-free LDS.128 from a small ring), the X rows are read out of a
// tile that holds 32 channels at once.
//
// Persistent CTA (one per SM); work item = (16x16 tile, group of 32 channels).
//   C  8 warps: warp w owns channels 4w..4w+3 of the item; lane (lr, ls) owns the 8-pixel strip at
//               tile row lr, column 8*ls.  Per displacement row: 18 LDS.128 (taps) + 4 x (4 LDS.128 +
//               72 FFMA); at the end of the item 8 STG.128.
//   T  1 warp : one TMA request per item for the X tile + halo of 32 channels (double buffered).
//   S  3 warps: stream the taps, one ring slot per (item, displacement row): 16-byte cp.async where the
//               source is 16-byte aligned (always for SIGN > 0), else aligned LDG.128 pairs + a register
//               shift + STS.128 (stage_tap_row); zero fill outside the image; runs NTS slots ahead.
// "Consumed" signals are given only after every instruction that reads the buffer has issued (one
// loop iteration late, or after dependent stores): an mbarrier arrive can overtake a pending LDS.
#pragma once
#include "corr_bwd_tma.cuh"

namespace pwc {

struct BwdSeqCfg {
    static constexpr int D = 9, S2 = 1, CK = 4, PX = 8, CPI = 32;     // CPI: channels per item
    static constexpr int r = 4, R = r * S2;
    static constexpr int TW = 16, TH = 16;
    static constexpr int NCONS = 256, NSTAGE = 96;
    static constexpr int NT = NCONS + 32 + NSTAGE;                      // 12 warps (168 registers each)
    static constexpr int HH = TH + 2 * R, HWD = TW + 2 * R;
    static constexpr int WP = HWD + 4;                                  // X tile pitch (TMA box width), 4 mod 8
    static constexpr int WSPAN = PX + 2 * R;
    static constexpr int X_ELEMS = CPI * HH * WP;
    static constexpr uint32_t X_BYTES = X_ELEMS * 4;
    static constexpr int NTS = 6;                                       // tap ring slots
    static constexpr int SLOT_ELEMS = D * TH * TW;                      // 9 planes of one displacement row
    static constexpr int CTRL_BYTES = 512;                              // tap slots start 512-byte aligned (64-byte TMA swizzle)
    static constexpr int GBOX_C = 27;
    static_assert(WP % 8 == 4, "pitch must be 4 mod 8 floats");
    static_assert((4 + 2 * NTS) * 8 <= CTRL_BYTES, "control block too small");
    static_assert((CTRL_BYTES + 2 * X_BYTES) % 512 == 0 && (SLOT_ELEMS * 4) % 512 == 0, "swizzle atom alignment of the tap slots");
    static_assert(D % (NSTAGE / 32) == 0, "each staging warp owns the same displacement rows in every item");
    static constexpr size_t smem_bytes() { return CTRL_BYTES + 2 * (size_t)X_BYTES + (size_t)NTS * SLOT_ELEMS * 4; }
};

template <int SIGN>
__global__ void __launch_bounds__(BwdSeqCfg::NT, 1)
corr_bwd_seq_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                    const __grid_constant__ CUtensorMap tmTap, int tma_taps,
                    const float* __restrict__ gout, float* __restrict__ res,
                    int C, int H, int W, int tiles_x, int tiles_y, int nitems, int nsc, long long gbs)
{
    using Cfg = BwdSeqCfg;
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, PX = Cfg::PX, R = Cfg::R, CPI = Cfg::CPI;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, WP = Cfg::WP;
    constexpr int WSPAN = Cfg::WSPAN, NCONS = Cfg::NCONS, NTS = Cfg::NTS;

    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* barX = reinterpret_cast<uint64_t*>(base);      // [2]   TMA: X item landed              (T -> C)
    uint64_t* barXFree = barX + 2;                           // [2]   X item consumed                 (C -> T)
    uint64_t* barTap = barXFree + 2;                         // [NTS] taps of a displacement row staged (S -> C)
    uint64_t* barTapFree = barTap + NTS;                     // [NTS] taps are in registers and used   (C -> S)
    float* sX = reinterpret_cast<float*>(base + Cfg::CTRL_BYTES);   // [2][CPI][HH][WP]
    float* sTap = sX + 2 * Cfg::X_ELEMS;                             // [NTS][9][TH][TW] swizzled (tap_slot)

    const int tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    const int my_items = ((int)blockIdx.x < nitems) ? (nitems - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&barX[i], 1);
            mbar_init(&barXFree[i], NCONS);
        }
#pragma unroll
        for (int i = 0; i < NTS; ++i) {
            // staging warps: per lane one cp.async arrival + one plain arrival; TMA taps: the expect_tx arrival
            mbar_init(&barTap[i], (SIGN > 0 && tma_taps) ? 1 : 64);
            mbar_init(&barTapFree[i], NCONS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= NCONS + 32) {
        // ================================ S: tap streaming ================================
        const int lane = tid & 31, sw = (tid - (NCONS + 32)) >> 5;      // sw owns rows dyi = sw, sw+3, sw+6
        if (SIGN > 0 && tma_taps) {
            // g1: one TMA request per displacement row ([9][16][16] box of the output gradient at the tile, zero
            // fill outside the image, 64-byte swizzle = tap_slot()), NTS rows ahead of the consumers
            if (sw != 0 || lane != 0) return;
            prefetch_tmap(&tmTap);
            int j = 0;
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const TileCoord tc = tile_coord(item / nsc, tiles_x, tiles_y, TH, TW);
                for (int dyi = 0; dyi < D; ++dyi, ++j) {
                    const int slot = j % NTS;
                    if (j >= NTS) mbar_wait(&barTapFree[slot], ((j / NTS) - 1) & 1);
                    mbar_expect_tx(&barTap[slot], Cfg::SLOT_ELEMS * 4);
                    tma_load_4d(sTap + slot * Cfg::SLOT_ELEMS, &tmTap, &barTap[slot], tc.x0, tc.y0, dyi * D, tc.n);
                }
            }
            return;
        }
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const TileCoord tc = tile_coord(item / nsc, tiles_x, tiles_y, TH, TW);
            const float* gon = gout + (size_t)tc.n * (size_t)gbs;      // gbs: batch stride of the output gradient
            for (int dyi = sw; dyi < D; dyi += Cfg::NSTAGE / 32) {
                const int j = it * D + dyi, slot = j % NTS;
                if (j >= NTS) mbar_wait(&barTapFree[slot], ((j / NTS) - 1) & 1);
                stage_tap_row<S2, SIGN>(sTap + slot * Cfg::SLOT_ELEMS, gon, dyi, tc, H, W, HW, lane, gout);
                cp_async_mbar_arrive(&barTap[slot]);       // arrives once this thread's asynchronous copies have landed
                mbar_arrive(&barTap[slot]);                // release: this thread's shifted quads are stored
            }
        }
        return;
    }

    if (tid >= NCONS) {
        // ================================ T: TMA issue ================================
        if (tid != NCONS) return;
        prefetch_tmap(&tmX);
        prefetch_tmap(&tmG);
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x, xb = it & 1;
            const TileCoord tc = tile_coord(item / nsc, tiles_x, tiles_y, TH, TW);
            if (it + 1 < my_items) {
                // pull the next item's output-gradient region into L2 (TMA prefetch, no shared memory)
                const TileCoord tn = tile_coord((item + gridDim.x) / nsc, tiles_x, tiles_y, TH, TW);
                const int off = (SIGN > 0) ? 0 : R;
#pragma unroll
                for (int q = 0; q < (D * D) / Cfg::GBOX_C; ++q)
                    tma_prefetch_4d(&tmG, tn.x0 - off, tn.y0 - off, q * Cfg::GBOX_C, tn.n);
            }
            if (it >= 2) mbar_wait(&barXFree[xb], ((it >> 1) - 1) & 1);
            mbar_expect_tx(&barX[xb], Cfg::X_BYTES);
            tma_load_4d(sX + xb * Cfg::X_ELEMS, &tmX, &barX[xb], tc.x0 - R, tc.y0 - R, (item % nsc) * CPI, tc.n);
        }
        return;
    }

    // ================================ C: complete outputs ================================
    // Programmatic dependent launch: once every CTA has started its LAST item (or has none), a kernel that
    // was launched behind this one with the programmatic-serialization attribute may start on the SMs
    // that fall idle during the tail (9.08 items per CTA at the level-2 shape: 136 of 148 SMs idle for
    // the last item).  Such a dependent must not consume this kernel's output (see pwc_abi.cu).
    if (tid == 0 && my_items <= 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int lane = tid & 31, cg = tid >> 5;              // cg: channel group inside the item
    const int lr = lane & 15, ls = lane >> 4;
    const float inv_nelems = __frcp_rn((float)C);          // 1/C, correlation_cuda_kernel.cu:194,286
    int j = 0;                                             // running tap-slot counter
    for (int it = 0; it < my_items; ++it) {
        const int item = blockIdx.x + it * gridDim.x, xb = it & 1;
        if (tid == 0 && it == my_items - 1 && my_items > 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const TileCoord tc = tile_coord(item / nsc, tiles_x, tiles_y, TH, TW);
        const int c_base = (item % nsc) * CPI + cg * CK;
        float part[CK][PX];
#pragma unroll
        for (int c = 0; c < CK; ++c)
#pragma unroll
            for (int p = 0; p < PX; ++p) part[c][p] = 0.0f;

        mbar_wait(&barX[xb], (it >> 1) & 1);
        const float* px = sX + xb * Cfg::X_ELEMS + (cg * CK) * (HH * WP) + lr * WP + ls * PX;
#pragma unroll 1
        for (int dyi = 0; dyi < D; ++dyi, ++j) {
            const int slot = j % NTS;
            mbar_wait(&barTap[slot], (j / NTS) & 1);
            const float* tp = sTap + slot * Cfg::SLOT_ELEMS;
            float G[PX][D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float4 a = *reinterpret_cast<const float4*>(tp + tap_slot(d, lr, ls * PX));
                const float4 b = *reinterpret_cast<const float4*>(tp + tap_slot(d, lr, ls * PX + 4));
                G[0][d] = a.x; G[1][d] = a.y; G[2][d] = a.z; G[3][d] = a.w;
                G[4][d] = b.x; G[5][d] = b.y; G[6][d] = b.z; G[7][d] = b.w;
            }
            // the previous row's taps were consumed by the previous iteration (all of it has issued)
            if (dyi > 0) mbar_arrive(&barTapFree[(j - 1) % NTS]);
            const int rowsel = (SIGN > 0) ? dyi : (D - 1 - dyi);
            const float* pw = px + rowsel * S2 * WP;
#pragma unroll
            for (int c = 0; c < CK; ++c) {
#pragma unroll
                for (int q = 0; q < WSPAN / 4; ++q) {
                    const float4 v4 = *reinterpret_cast<const float4*>(pw + c * (HH * WP) + 4 * q);
                    const float wq[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int jj = 4 * q + e;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const int col = (SIGN > 0) ? d * S2 : (D - 1 - d) * S2;
                            const int p = jj - col;
                            if (p >= 0 && p < PX) part[c][p] = fmaf(G[p][d], wq[e], part[c][p]);
                        }
                    }
                }
            }
        }

        // ---- store the finished strip (W % 4 == 0: a quad is inside or outside as a whole) ----
        const int y = tc.y0 + lr, xs = tc.x0 + ls * PX;
#pragma unroll
        for (int c = 0; c < CK; ++c) {
            const int cc = c_base + c;
            if (cc < C && y < H && xs < W) {
                float* o = res + ((size_t)tc.n * C + cc) * HW + (size_t)y * W + xs;
                *reinterpret_cast<float4*>(o) = make_float4(part[c][0] * inv_nelems, part[c][1] * inv_nelems,
                                                            part[c][2] * inv_nelems, part[c][3] * inv_nelems);
                if (xs + 4 < W)
                    *reinterpret_cast<float4*>(o + 4) = make_float4(part[c][4] * inv_nelems, part[c][5] * inv_nelems,
                                                                    part[c][6] * inv_nelems, part[c][7] * inv_nelems);
            }
        }
        // the accumulators just stored depend on every load of the item: both buffers are free now
        mbar_arrive(&barTapFree[(j - 1) % NTS]);
        mbar_arrive(&barXFree[xb]);
    }
}

}  // namespace pwc
