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
//   * the CTA is persistent (one per SM, looping over 16x16 tiles) and warp-specialised into four
//     roles that meet only through mbarriers (no __syncthreads after start-up):
//       T  1 warp : issues every TMA request (flow tiles, f1 chunks, f2 windows), running ahead;
//       P  1 warp : turns a flow tile into bilinear taps + the window origin (bbox by warp
//                   shuffles), one tile ahead of the bilinear warps;
//       B  5 warps: evaluate the bilinear warp out of the window into a ring of warped chunks
//                   (taps live in registers for the whole tile, loads are batched for MLP);
//       C  9 warps: a thread owns one (tile row, 8-pixel strip, displacement row) task and its 8x9
//                   accumulators; correlate out of shared memory (packed FFMA2, quad-shared LDS.128
//                   addresses for stride2 = 1), then write the cost volume.
//     The stream of (tile, channel-chunk) work items is continuous across tiles.  For stride2 = 1 the
//     last group of four warps hands 48 registers per thread to the first twelve (setmaxnreg).
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

// 256-bit global store (sm_100: STG.E.ENL2.256): one full 32-byte sector per lane and instruction.
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8])
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
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
    static constexpr int NCONS = 32 * D;                   // C: one warp per tj
    static constexpr int NBIL = 160;                       // B: bilinear warps
    static constexpr int NT = NCONS + NBIL + 64;           // + P warp + T warp = 16 warps (regs are granted per 4 warps)
    static constexpr int HH = TH + 2 * R, HWD = TW + 2 * R;   // warped tile + halo
    static constexpr int WP = HWD + 4;                     // pitch = 4 (mod 8): conflict-free 128-bit rows
    static constexpr int MARGIN = 8;                       // extra source pixels each side of the halo
    // f2 source window (TMA box).  +4 columns: the innermost TMA start coordinate must be a multiple
    // of 16 bytes (measured on B200: a 4-D tiled fp32 load at x % 4 != 0 raises "illegal
    // instruction"), so the window origin is rounded down to a multiple of 4 pixels.
    static constexpr int WW = HWD + 2 * MARGIN + 4, WH = HH + 2 * MARGIN;
    static constexpr int F1W = TW + 4, F1H = TH;           // f1 box, pitch 20 = 4 (mod 8)
    static constexpr int NHALO = HH * HWD;
    static constexpr int PXB = (NHALO + NBIL - 1) / NBIL;  // halo pixels per bilinear thread
    static constexpr int PXP = (NHALO + 31) / 32;          // halo pixels per lane of the taps warp
    static constexpr int WSPAN = PX + 2 * R;
    // Ring depths.  The T warp walks ONE in-order loop over both rings: in iteration g it first waits for the f1 slot
    // of item g + NS (free once the consumers are past chunk g + NS - NF1 + 1) and only then requests the warped / window
    // slot of item g, so the operands of item g are requested NF1 - NS - 1 chunks ahead of their use at most.  Measured
    // at the level-2 shape without a flow (TMA fills the warped ring itself, scripts/dev_fwd.cu): (NS, NF1) = (3,6) 95 us,
    // (5,8) 95 us, (4,8) 83 us, (4,12) 82 us, (6,8) / (8,10) / (12,14) 137 us -- a lead of 3 chunks hides the TMA latency,
    // a lead of 1 exposes it in every chunk.  With a flow the bilinear stage sits in between and the depths do not matter
    // (107.5 - 111.6 us for all of the above).
    static constexpr int NS = 4;   // warped-chunk ring (B -> C); a slot is released one chunk late (see C)
    static constexpr int NF1 = 8;                          // f1 ring
    static constexpr int NWIN = 3;                         // f2 window ring (T -> B)
    static constexpr int WIN_ELEMS = CK * WH * WW;
    static constexpr int F1_ELEMS = CK * F1H * F1W;
    static constexpr int W2_ELEMS = CK * HH * WP;
    static constexpr int FLOW_ELEMS = 2 * HH * HWD;        // flow tile + halo (u and v), TMA box
    static constexpr uint32_t WIN_BYTES = WIN_ELEMS * 4, F1_BYTES = F1_ELEMS * 4, W2_BYTES = W2_ELEMS * 4;
    static constexpr uint32_t FLOW_BYTES = FLOW_ELEMS * 4;
    // model.py:78 folded into the flow read: instead of the fine flow tile + halo the T warp fetches the
    // coarse flow's [CHB][CWB] box around it (per component) and the P warp evaluates 2 * bilinear-up2 from
    // shared memory.  Fine columns x0-R .. x0+TW+R-1 need coarse columns x0/2 - R/2 - 1 .. x0/2 + (TW+R)/2;
    // the box starts CW_OFF (a multiple of 4: 16-byte aligned TMA start, see WW) left of x0/2.
    static constexpr int CW_OFF = (R / 2 + 1 + 3) / 4 * 4, CH_OFF = R / 2 + 1;
    static constexpr int CWB = (CW_OFF + (TW + R) / 2 + 1 + 3) / 4 * 4, CHB = CH_OFF + (TH + R) / 2 + 1;
    static constexpr uint32_t CFLOW_BYTES = 2 * CHB * CWB * 4;
    static_assert(2 * CHB * CWB <= FLOW_ELEMS, "the coarse flow box reuses the fine flow tile's buffer");
    static_assert(TW % 8 == 0 && TH % 2 == 0 && R % 2 == 0, "tile origins must halve to multiples of 4 / integers");
    static constexpr int NBARS = 2 * NF1 + 2 * NS + 2 * NWIN + 8;
    static constexpr int CTRL_BYTES = 512;                 // mbarriers + window origins
    static_assert(WP % 8 == 4 && F1W % 8 == 4, "pitches must be 4 mod 8 floats");
    static_assert(HH % 8 == 0 && HWD % 4 == 0 && NHALO % 32 == 0, "the bilinear warps walk 4x8 pixel patches");
    static_assert((WW * 4) % 16 == 0 && (F1W * 4) % 16 == 0 && (WP * 4) % 16 == 0 && (HWD * 4) % 16 == 0,
                  "TMA box rows are 16B multiples");
    static_assert(WIN_BYTES % 128 == 0 && F1_BYTES % 128 == 0 && W2_BYTES % 128 == 0 && FLOW_BYTES % 128 == 0,
                  "buffers stay 128B aligned");
    static_assert(NBARS * 8 + 16 <= CTRL_BYTES, "control block too small");
    // Deadlock freedom of the T warp: in iteration g it waits for f1 slot (g + NS) % NF1, i.e. for item
    // g + NS - NF1, which the consumers release one chunk late (after chunk g + NS - NF1 + 1); the window of
    // that chunk must already have been requested (iteration g - 1 requested windows below g - 1 + NWIN).
    static_assert(NF1 > NS + 2 - NWIN && NS >= 2, "ring depths would deadlock the TMA warp (measured: a hang)");

    static constexpr size_t smem_bytes(bool has_flow)
    {
        size_t b = CTRL_BYTES + (size_t)NF1 * F1_BYTES + (size_t)NS * W2_BYTES;
        if (has_flow)
            b += (size_t)NWIN * WIN_BYTES + 2 * (size_t)FLOW_BYTES +
                 2 * (size_t)NHALO * (sizeof(float4) + sizeof(int));
        return b;
    }
};

constexpr int TAP_EMPTY = -1;      // every corner outside the image (or non-finite flow): value 0
constexpr int TAP_GLOBAL = -2;     // footprint outside the staged window: gather from global memory
constexpr int TAP_NONE = -3;       // padding entry of a thread's tap list (beyond NHALO)

struct TileCoord { int n, y0, x0; };

// Bilinear-warp work item i (0 .. NHALO-1, padded to whole warps) -> halo pixel.  A warp's 32 items form
// a 4-wide x 8-tall patch: with the window pitch of 44 (and 52) floats the eight rows start in banks
// {0,12,24,4,16,28,8,20} (+4 columns each), so for a locally uniform displacement -- every real flow
// field -- the four corner loads of a warp hit 32 distinct banks; the warped-tile pitch of 28 (36) makes
// the stores conflict-free as well.  (Consecutive pixels of 24-wide rows put lanes 24-31 on the banks of
// lanes 12-19: 2-way conflicts, measured 5.3 M of 19.6 M wavefronts with a smooth flow.)
template <int HWD>
__device__ __forceinline__ void halo_item(int i, int& hy, int& hx)
{
    static_assert(HWD % 4 == 0, "patches are 4 pixels wide");
    const int patch = i >> 5, l = i & 31;
    hx = (patch % (HWD / 4)) * 4 + (l & 3);
    hy = (patch / (HWD / 4)) * 8 + (l >> 2);
}
__device__ __forceinline__ TileCoord tile_coord(int tile, int tiles_x, int tiles_y, int TH, int TW)
{
    TileCoord t;
    const int tx = tile % tiles_x;
    const int rest = tile / tiles_x;
    t.x0 = tx * TW;
    t.y0 = (rest % tiles_y) * TH;
    t.n = rest / tiles_y;
    return t;
}

// Ring-slot counter (slot index + phase parity of the slot's current use), advanced per work item: the roles walk
// their streams of (tile, chunk) items with increments only.  Measured (ncu source view, level-2 shape): the integer
// divisions of `g / nchunks`, `g % NS` and of tile_coord(), re-evaluated per chunk, were ~40% of the bilinear
// role's dependent instruction chain, and that role is the one the correlation warps wait for.
template <int N>
struct Ring {
    int slot = 0, phase = 0;
    __device__ __forceinline__ void next()
    {
        if (++slot == N) { slot = 0; phase ^= 1; }
    }
};
// (local tile, chunk) cursor of a role's item stream; the tile coordinate is recomputed once per tile
struct ItemCursor {
    int lt = 0, k = 0;
    TileCoord tc;
    __device__ __forceinline__ void start(int tiles_x, int tiles_y, int TH, int TW)
    {
        tc = tile_coord(blockIdx.x, tiles_x, tiles_y, TH, TW);
    }
    __device__ __forceinline__ void next(int nchunks, int tiles_x, int tiles_y, int TH, int TW)
    {
        if (++k == nchunks) {
            k = 0;
            ++lt;
            tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
        }
    }
};

// Cold path of the bilinear role: a sample whose 2x2 footprint lies outside the staged window is gathered from
// global memory (correctness never depends on the flow magnitude).  The flow comes from `flow` (batch stride fbs) or
// is evaluated from `coarse` (model.py:78 folded in).  vv[c] = warped f2[n, c0 + c] at pixel (x, y), 0 beyond C.
template <int CK>
__device__ __noinline__ void global_tap_values(float* vv, const float* __restrict__ f2, const float* __restrict__ flow,
                                               long long fbs, const float* __restrict__ coarse, int n, int x, int y,
                                               int c0, int C, int H, int W)
{
    const size_t HW = (size_t)H * W;
    float fu, fv;
    if (coarse != nullptr) {
        const int Hc = H >> 1, Wc = W >> 1;
        const float* cu = coarse + (size_t)n * 2 * Hc * Wc;
        up2_flow_at(cu, cu + Hc * Wc, Hc, Wc, x, y, fu, fv);
    } else {
        const float* un = flow + (size_t)n * (size_t)fbs;
        fu = __ldg(un + (size_t)y * W + x);
        fv = __ldg(un + HW + (size_t)y * W + x);
    }
    const Tap tp = make_tap(x, y, fu, fv, H, W);
#pragma unroll
    for (int c = 0; c < CK; ++c)
        vv[c] = (c0 + c < C && tp.off >= 0) ? tap_sample(tp, f2 + ((size_t)n * C + c0 + c) * HW) : 0.0f;
}

#ifndef PWC_ROLE_ATTR
#define PWC_ROLE_ATTR __forceinline__
#endif

// Kernel arguments as one record: the producer roles are compiled as separate (non-inlined) functions, so that their
// register allocation is independent of the correlation role, which fills the 128-register budget of a 512-thread
// CTA on its own.  (With every role inlined, values of the common prologue were spilled to local memory and reloaded
// inside the producer loops -- measured: 180-280 bytes of spills once the loops kept their cursors in registers.)
struct FwdArgs {
    const CUtensorMap *tmF1, *tmF2, *tmFlow;
    const float *f2, *flow;
    float *out, *warped_out;
    int C, H, W, tiles_x, tiles_y, ntiles, act;
    float slope;
    long long obs, fbs;
    const float* coarse;
    float* flow_out;
    long long fobs;
};

// The context every role derives for itself: compile-time geometry, the shared-memory carve-up (declared from the
// extern array in each function, so the accesses stay LDS/STS), and this CTA's share of the work.
#define PWC_FWD_CTX(A)                                                                                             \
    const CUtensorMap& tmF1 = *(A).tmF1; const CUtensorMap& tmF2 = *(A).tmF2; const CUtensorMap& tmFlow = *(A).tmFlow;  \
    const float* __restrict__ f2 = (A).f2; const float* __restrict__ flow = (A).flow; float* __restrict__ out = (A).out;\
    float* __restrict__ warped_out = (A).warped_out; const int C = (A).C, H = (A).H, W = (A).W;                         \
    const int tiles_x = (A).tiles_x, tiles_y = (A).tiles_y, ntiles = (A).ntiles, act = (A).act; const float slope = (A).slope;\
    const long long obs = (A).obs, fbs = (A).fbs, fobs = (A).fobs; const float* __restrict__ coarse = (A).coarse;       \
    float* __restrict__ flow_out = (A).flow_out;                                                                        \
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, PX = Cfg::PX, R = Cfg::R;                                     \
    constexpr int TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, HWD = Cfg::HWD;                                             \
    constexpr int WP = Cfg::WP, WW = Cfg::WW, WH = Cfg::WH, F1W = Cfg::F1W, F1H = Cfg::F1H;                             \
    constexpr int NHALO = Cfg::NHALO, WSPAN = Cfg::WSPAN, NCONS = Cfg::NCONS, NBIL = Cfg::NBIL;                         \
    constexpr int NS = Cfg::NS, NF1 = Cfg::NF1, NWIN = Cfg::NWIN, PXB = Cfg::PXB, PXP = Cfg::PXP;                       \
    extern __shared__ __align__(1024) uint8_t base[];                                                                   \
    uint64_t* barF1 = reinterpret_cast<uint64_t*>(base);                                                                \
    uint64_t* barF1Free = barF1 + NF1;                                                                                  \
    uint64_t* barFull = barF1Free + NF1;                                                                                \
    uint64_t* barEmpty = barFull + NS;                                                                                  \
    uint64_t* barWin = barEmpty + NS;                                                                                   \
    uint64_t* barWinFree = barWin + NWIN;                                                                               \
    uint64_t* barFlow = barWinFree + NWIN;                                                                              \
    uint64_t* barFlowFree = barFlow + 2;                                                                                \
    uint64_t* barTaps = barFlowFree + 2;                                                                                \
    uint64_t* barTapsFree = barTaps + 2;                                                                                \
    int* worg = reinterpret_cast<int*>(base + Cfg::NBARS * 8);                                                          \
    float* sF1 = reinterpret_cast<float*>(base + Cfg::CTRL_BYTES);                                                      \
    float* sW2 = sF1 + NF1 * Cfg::F1_ELEMS;                                                                             \
    float* sWin = sW2 + NS * Cfg::W2_ELEMS;                                                                             \
    float* sFlow = sWin + NWIN * Cfg::WIN_ELEMS;                                                                        \
    float4* sTapW = reinterpret_cast<float4*>(sFlow + 2 * Cfg::FLOW_ELEMS);                                             \
    int* sTapM = reinterpret_cast<int*>(sTapW + 2 * NHALO);                                                             \
    const int tid = threadIdx.x;                                                                                        \
    const size_t HW = (size_t)H * W;                                                                                    \
    const int nchunks = (C + CK - 1) / CK;                                                                              \
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;          \
    const int total = my_tiles * nchunks;                                                                               \
    (void)tmF1; (void)tmF2; (void)tmFlow; (void)f2; (void)flow; (void)out; (void)warped_out; (void)act; (void)slope; (void)obs; (void)fbs; (void)fobs; (void)coarse; (void)flow_out; (void)barF1; (void)barF1Free; (void)barFull; (void)barEmpty; (void)barWin; (void)barWinFree; (void)barFlow; (void)barFlowFree; (void)barTaps; (void)barTapsFree; (void)worg; (void)sF1; (void)sW2; (void)sWin; (void)sFlow; (void)sTapW; (void)sTapM; (void)HW; (void)total; (void)tid; (void)H; (void)W;

// ================================ T: TMA issue warp (one thread) ================================
template <class Cfg, bool HAS_FLOW>
__device__ PWC_ROLE_ATTR void fwd_role_tma(const FwdArgs& A)
{
    PWC_FWD_CTX(A)
    // ================================ T: TMA issue warp ================================
    if (tid != NCONS + NBIL + 32) return;
    prefetch_tmap(&tmF1);
    prefetch_tmap(&tmF2);
    ItemCursor cf;                        // next f1 chunk to request
    Ring<NF1> rf;
    cf.start(tiles_x, tiles_y, TH, TW);
    int nf = 0;                           // f1 chunks requested so far
    auto request_f1 = [&]() {             // f1 chunk of the next work item -> its ring slot
        if (nf >= NF1) mbar_wait(&barF1Free[rf.slot], rf.phase ^ 1);      // item nf - NF1 consumed
        mbar_expect_tx(&barF1[rf.slot], Cfg::F1_BYTES);
        tma_load_4d(sF1 + rf.slot * Cfg::F1_ELEMS, &tmF1, &barF1[rf.slot], cf.tc.x0, cf.tc.y0, cf.k * CK, cf.tc.n);
        cf.next(nchunks, tiles_x, tiles_y, TH, TW);
        rf.next();
        ++nf;
    };
    if (!HAS_FLOW) {
        // plain correlation: the f2 tile + halo *is* the warped chunk; TMA writes it directly
        for (int j = 0; j < NS && j < total; ++j) request_f1();
        ItemCursor cw;
        Ring<NS> rs;
        cw.start(tiles_x, tiles_y, TH, TW);
        for (int g = 0; g < total; ++g) {
            if (nf < total) request_f1();
            if (g >= NS) mbar_wait(&barEmpty[rs.slot], rs.phase ^ 1);
            mbar_expect_tx(&barFull[rs.slot], Cfg::W2_BYTES);
            tma_load_4d(sW2 + rs.slot * Cfg::W2_ELEMS, &tmF2, &barFull[rs.slot], cw.tc.x0 - R, cw.tc.y0 - R, cw.k * CK,
                        cw.tc.n);
            cw.next(nchunks, tiles_x, tiles_y, TH, TW);
            rs.next();
        }
        return;
    }
    prefetch_tmap(&tmFlow);
    auto request_flow = [&](int lt) {    // flow tile + halo of local tile lt -> sFlow[lt & 1]
        const TileCoord tj = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
        if (coarse != nullptr) {         // the coarse box that covers the tile + halo
            mbar_expect_tx(&barFlow[lt & 1], Cfg::CFLOW_BYTES);
            tma_load_4d(sFlow + (lt & 1) * Cfg::FLOW_ELEMS, &tmFlow, &barFlow[lt & 1], tj.x0 / 2 - Cfg::CW_OFF,
                        tj.y0 / 2 - Cfg::CH_OFF, 0, tj.n);
        } else {
            mbar_expect_tx(&barFlow[lt & 1], Cfg::FLOW_BYTES);
            tma_load_4d(sFlow + (lt & 1) * Cfg::FLOW_ELEMS, &tmFlow, &barFlow[lt & 1], tj.x0 - R, tj.y0 - R, 0, tj.n);
        }
    };
    if (my_tiles > 0) request_flow(0);
    if (my_tiles > 1) request_flow(1);
    for (int j = 0; j < NS && j < total; ++j) request_f1();
    int wx0 = 0, wy0 = 0;
    int jw = 0;             // next work item whose f2 window has not been requested
    ItemCursor cw;
    Ring<NWIN> rw;
    cw.start(tiles_x, tiles_y, TH, TW);
    for (int g = 0; g < total; ++g) {
        if (nf < total) request_f1();
        for (; jw < total && jw < g + NWIN; ++jw) {
            if (cw.k == 0) {             // first chunk of a tile: its window origin comes from the taps warp
                const int lt = cw.lt;
                if (lt + 1 < my_tiles && lt + 1 >= 2) {
                    mbar_wait(&barFlowFree[(lt + 1) & 1], ((lt - 1) >> 1) & 1);   // P is done with tile lt - 1's flow
                    request_flow(lt + 1);
                }
                mbar_wait(&barTaps[lt & 1], (lt >> 1) & 1);
                wx0 = worg[2 * (lt & 1)]; wy0 = worg[2 * (lt & 1) + 1];
            }
            if (jw >= NWIN) mbar_wait(&barWinFree[rw.slot], rw.phase ^ 1);   // B is done with item jw - NWIN
            mbar_expect_tx(&barWin[rw.slot], Cfg::WIN_BYTES);
            tma_load_4d(sWin + rw.slot * Cfg::WIN_ELEMS, &tmF2, &barWin[rw.slot], wx0, wy0, cw.k * CK, cw.tc.n);
            cw.next(nchunks, tiles_x, tiles_y, TH, TW);
            rw.next();
        }
    }
    return;
}

// ================================ P: taps warp ================================
template <class Cfg, bool HAS_FLOW>
__device__ PWC_ROLE_ATTR void fwd_role_taps(const FwdArgs& A)
{
    PWC_FWD_CTX(A)
    // ================================ P: taps warp ================================
    if (!HAS_FLOW) return;
    const int lane = tid & 31;
    for (int lt = 0; lt < my_tiles; ++lt) {
        const int par = lt & 1;
        const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
        mbar_wait(&barFlow[par], (lt >> 1) & 1);
        if (lt >= 2) mbar_wait(&barTapsFree[par], ((lt >> 1) - 1) & 1);   // B finished tile lt - 2
        const float* sfl = sFlow + par * Cfg::FLOW_ELEMS;
        float4* tapW = sTapW + par * NHALO;
        int* tapM = sTapM + par * NHALO;
        // pass 1: sample positions of tile + halo and their bounding box (pixels outside the image are
        // skipped by coordinate; the TMA zero fill of the flow tile is never interpreted)
        int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -0x7fffffff, mxy = -0x7fffffff;
#pragma unroll 4
        for (int j = 0; j < PXP; ++j) {
            const int i = lane + 32 * j;
            if (i < NHALO) {
                const int hy = i / HWD, hx = i - hy * HWD;
                const int y = tc.y0 - R + hy, x = tc.x0 - R + hx;
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                int meta = TAP_EMPTY;
                if (y >= 0 && y < H && x >= 0 && x < W) {
                    float u, v;
                    if (coarse != nullptr) {
                        constexpr int CWB = Cfg::CWB, CPL = Cfg::CHB * Cfg::CWB;
                        int xl, xr, yl, yr;
                        const float lx1 = up2_source(x, W >> 1, xl, xr), ly1 = up2_source(y, H >> 1, yl, yr);
                        const int cx0 = tc.x0 / 2 - Cfg::CW_OFF, cy0 = tc.y0 / 2 - Cfg::CH_OFF;
                        const int ia = (yl - cy0) * CWB + (xl - cx0), ib = (yl - cy0) * CWB + (xr - cx0);
                        const int ic = (yr - cy0) * CWB + (xl - cx0), id = (yr - cy0) * CWB + (xr - cx0);
                        u = up2_blend(sfl[ia], sfl[ib], sfl[ic], sfl[id], lx1, ly1);
                        v = up2_blend(sfl[CPL + ia], sfl[CPL + ib], sfl[CPL + ic], sfl[CPL + id], lx1, ly1);
                        if (flow_out != nullptr && hy >= R && hy < R + TH && hx >= R && hx < R + TW) {
                            float* fo = flow_out + (size_t)tc.n * (size_t)fobs + (size_t)y * W + x;
                            fo[0] = u;
                            fo[HW] = v;
                        }
                    } else {
                        u = sfl[i];
                        v = sfl[NHALO + i];
                    }
                    if (fabsf(u) < 1.0e6f && fabsf(v) < 1.0e6f) {          // rejects NaN / Inf as well
                        // floor + fraction of the flow first: the fraction is exact in fp32 (pwc_common.cuh)
                        const float fu = floorf(u), fv = floorf(v);
                        const float ax = u - fu, ay = v - fv;
                        const int x0 = x + (int)fu, y0 = y + (int)fv;
                        if (x0 >= -1 && x0 < W && y0 >= -1 && y0 < H) {
                            w = make_float4((1.0f - ax) * (1.0f - ay), ax * (1.0f - ay), (1.0f - ax) * ay, ax * ay);
                            meta = ((y0 + 1) << 16) | (x0 + 1);   // x0, y0 >= -1; H, W < 32760 (host check)
                            mnx = min(mnx, x0); mxx = max(mxx, x0 + 1);
                            mny = min(mny, y0); mxy = max(mxy, y0 + 1);
                        }
                    }
                }
                tapW[i] = w;
                tapM[i] = meta;
            }
        }
        mbar_arrive(&barFlowFree[par]);      // this lane no longer reads sFlow[par]
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        // window origin: the bounding box if it fits, else centred on it (outliers -> global gather)
        int wx0 = 0, wy0 = 0;
        if (mnx <= mxx) {
            wx0 = mnx & ~3;                                   // 16-byte aligned TMA start (also for x < 0)
            if (mxx - wx0 + 1 > WW) wx0 = ((mnx + mxx + 1 - WW) >> 1) & ~3;
            wy0 = (mxy - mny + 1 <= WH) ? mny : (mny + mxy + 1 - WH) / 2;
        }
        // pass 2: positions -> window-relative offsets (each lane revisits exactly the taps it wrote)
#pragma unroll 4
        for (int j = 0; j < PXP; ++j) {
            const int i = lane + 32 * j;
            if (i < NHALO) {
                const int meta = tapM[i];
                if (meta >= 0) {
                    const int x0 = (meta & 0xffff) - 1, y0 = (meta >> 16) - 1;
                    const int rx = x0 - wx0, ry = y0 - wy0;
                    tapM[i] = (rx >= 0 && rx + 1 < WW && ry >= 0 && ry + 1 < WH) ? ry * WW + rx : TAP_GLOBAL;
                }
            }
        }
        if (lane == 0) { worg[2 * par] = wx0; worg[2 * par + 1] = wy0; }
        mbar_arrive(&barTaps[par]);          // release: taps[par] and worg[par] are complete
    }
    return;
}

// ================================ B: bilinear warps ================================
template <class Cfg, bool HAS_FLOW>
__device__ PWC_ROLE_ATTR void fwd_role_bilinear(const FwdArgs& A)
{
    PWC_FWD_CTX(A)
    // ================================ B: bilinear warps ================================
    if (!HAS_FLOW) return;
    const int btid = tid - NCONS;
    float4 tw[PXB];          // this thread's taps for the current tile (registers for all its chunks)
    int toff[PXB];           // window offset (0 for empty taps: weights are 0), or TAP_NONE / TAP_GLOBAL
    int tdst[PXB];           // destination offset in the warped chunk
    bool any_global = false;
    Ring<NS> rs;             // warped-chunk ring
    Ring<NWIN> rw;           // f2 window ring
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
        const int par = lt & 1;
        mbar_wait(&barTaps[par], (lt >> 1) & 1);
        any_global = false;
#pragma unroll
        for (int j = 0; j < PXB; ++j) {
            const int i = btid + j * NBIL;
            const bool valid = i < NHALO;
            const int ii = valid ? i : 0;
            int hy, hx;
            halo_item<HWD>(ii, hy, hx);
            tw[j] = sTapW[par * NHALO + hy * HWD + hx];
            const int meta = sTapM[par * NHALO + hy * HWD + hx];
            tdst[j] = hy * WP + hx;
            toff[j] = !valid ? TAP_NONE : (meta == TAP_EMPTY ? 0 : meta);
            any_global |= valid && meta == TAP_GLOBAL;
        }
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++g) {
        const int c0 = k * CK;
        if (g >= NS) mbar_wait(&barEmpty[rs.slot], rs.phase ^ 1);    // consumers released warped slot
        mbar_wait(&barWin[rw.slot], rw.phase);
        const float* win = sWin + rw.slot * Cfg::WIN_ELEMS;
        float* w2buf = sW2 + rs.slot * Cfg::W2_ELEMS;
        float v[PXB][CK];
        // The corner loads go out in groups of 8 (two channels of one pixel; the empty asm is a compiler-
        // level fence that keeps the groups apart).  64 loads per thread in one burst kept the LSU queue
        // full of this role's conflicting scalar loads, and the consumers' LDS.128 waited behind them:
        // bursts of 64 / 16 / 8 / 4 loads -> 118.5 / 117.3 / 112.8 / 118.2 us at the level-2 shape.  The
        // stores stay together at the end (storing each pixel at once: 128 us).
#pragma unroll
        for (int j = 0; j < PXB; ++j) {
            const float* p = win + (toff[j] >= 0 ? toff[j] : 0);     // empty / none / global: a safe address
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                const float* q = p + c * (WH * WW);
#if defined(PWC_DEV_X) && PWC_DEV_X == 3
                v[j][c] = tw[j].x;                             // (dev ablation) no loads at all
#else
                v[j][c] = fmaf(tw[j].w, q[WW + 1], fmaf(tw[j].z, q[WW], fmaf(tw[j].y, q[1], tw[j].x * q[0])));
#endif
#ifndef PWC_BGROUP
#define PWC_BGROUP 2
#endif
                if (PWC_BGROUP > 0 && (c % PWC_BGROUP) == PWC_BGROUP - 1) asm volatile("" ::: "memory");
            }
        }
        if (any_global) {
            // outliers (cold path, its own function so that it costs the loop no registers): recompute the tap
            // from the flow and gather from global memory
#pragma unroll
            for (int j = 0; j < PXB; ++j) {
                if (toff[j] == TAP_GLOBAL) {
                    const int i = btid + j * NBIL;
                    int hy, hx;
                    halo_item<HWD>(i, hy, hx);
                    float vv[CK];
                    const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);   // (cold)
                    global_tap_values<CK>(vv, f2, flow, fbs, coarse, tc.n, tc.x0 - R + hx, tc.y0 - R + hy, c0, C, H, W);
#pragma unroll
                    for (int c = 0; c < CK; ++c) v[j][c] = vv[c];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PXB; ++j) {
            if (toff[j] != TAP_NONE) {
                float* dst = w2buf + tdst[j];
#pragma unroll
                for (int c = 0; c < CK; ++c) dst[c * (HH * WP)] = v[j][c];
            }
        }
        if (warped_out != nullptr) {     // x2_warp export (model.py:107,113)
            const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
#pragma unroll
            for (int j = 0; j < PXB; ++j) {
                const int i = btid + j * NBIL;
                int hy, hx;
                halo_item<HWD>(i < NHALO ? i : 0, hy, hx);
                const int gy = tc.y0 - R + hy, gx = tc.x0 - R + hx;
                if (toff[j] != TAP_NONE && hy >= R && hy < R + TH && hx >= R && hx < R + TW && gy < H && gx < W) {
                    float* wo = warped_out + ((size_t)tc.n * C + c0) * HW + (size_t)gy * W + gx;
#pragma unroll
                    for (int c = 0; c < CK; ++c)
                        if (c0 + c < C) wo[(size_t)c * HW] = v[j][c];
                }
            }
        }
        mbar_arrive(&barFull[rs.slot]);           // release: this thread's part of the warped chunk is written
        mbar_arrive(&barWinFree[rw.slot]);        // and it no longer reads this window
        rs.next();
        rw.next();
        }
        mbar_arrive(&barTapsFree[par]);
    }
    return;
}

template <class Cfg, bool HAS_FLOW>
__global__ void __launch_bounds__(Cfg::NT, 1)
warpcorr_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmF1, const __grid_constant__ CUtensorMap tmF2,
                        const __grid_constant__ CUtensorMap tmFlow, const float* __restrict__ f2,
                        const float* __restrict__ flow, float* __restrict__ out, float* __restrict__ warped_out,
                        int C, int H, int W, int tiles_x, int tiles_y, int ntiles, int act, float slope,
                        long long obs, long long fbs, const float* __restrict__ coarse,
                        float* __restrict__ flow_out, long long fobs)
{
    // flow source: `flow` (image n's [2][H][W] block at flow + n*fbs; tmFlow maps it), or, when `coarse` is
    // given ([B][2][H/2][W/2] dense; tmFlow maps THAT), flow = F.upsample(coarse, 2, 'bilinear') * 2
    // (model.py:78) evaluated by the P warp, which also writes the tile's fine flow to flow_out + n*fobs.
    constexpr int D = Cfg::D, S2 = Cfg::S2, CK = Cfg::CK, PX = Cfg::PX, R = Cfg::R;
    constexpr int TW = Cfg::TW, TH = Cfg::TH, HH = Cfg::HH, HWD = Cfg::HWD;
    constexpr int WP = Cfg::WP, WW = Cfg::WW, WH = Cfg::WH, F1W = Cfg::F1W, F1H = Cfg::F1H;
    constexpr int NHALO = Cfg::NHALO, WSPAN = Cfg::WSPAN, NCONS = Cfg::NCONS, NBIL = Cfg::NBIL;
    constexpr int NS = Cfg::NS, NF1 = Cfg::NF1, NWIN = Cfg::NWIN, PXB = Cfg::PXB, PXP = Cfg::PXP;

    // No pointer<->integer round trips on this pointer: the compiler must keep the shared address space
    // (otherwise every access below becomes a generic LD/ST instead of LDS/STS).
    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* barF1 = reinterpret_cast<uint64_t*>(base);   // [NF1]  TMA: f1 chunk landed           (T -> C)
    uint64_t* barF1Free = barF1 + NF1;                     // [NF1]  f1 chunk consumed               (C -> T)
    uint64_t* barFull = barF1Free + NF1;                   // [NS]   warped chunk ready              (B|TMA -> C)
    uint64_t* barEmpty = barFull + NS;                     // [NS]   warped chunk consumed           (C -> B|T)
    uint64_t* barWin = barEmpty + NS;                      // [NWIN] TMA: f2 window landed           (T -> B)
    uint64_t* barWinFree = barWin + NWIN;                  // [NWIN] window read by every B thread   (B -> T)
    uint64_t* barFlow = barWinFree + NWIN;                 // [2]    TMA: flow tile landed           (T -> P)
    uint64_t* barFlowFree = barFlow + 2;                   // [2]    flow tile read                  (P -> T)
    uint64_t* barTaps = barFlowFree + 2;                   // [2]    taps + window origin ready      (P -> B, T)
    uint64_t* barTapsFree = barTaps + 2;                   // [2]    taps no longer needed           (B -> P)
    int* worg = reinterpret_cast<int*>(base + Cfg::NBARS * 8);     // [2][2] window origin per tile parity
    float* sF1 = reinterpret_cast<float*>(base + Cfg::CTRL_BYTES);
    float* sW2 = sF1 + NF1 * Cfg::F1_ELEMS;
    float* sWin = sW2 + NS * Cfg::W2_ELEMS;                        // HAS_FLOW only from here on
    float* sFlow = sWin + NWIN * Cfg::WIN_ELEMS;                   // [2][2][HH][HWD]
    float4* sTapW = reinterpret_cast<float4*>(sFlow + 2 * Cfg::FLOW_ELEMS);    // [2][NHALO]
    int* sTapM = reinterpret_cast<int*>(sTapW + 2 * NHALO);                     // [2][NHALO]

    const int tid = threadIdx.x;
    const size_t HW = (size_t)H * W;
    const int nchunks = (C + CK - 1) / CK;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; work items are (tile, chunk) pairs
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * nchunks;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NF1; ++i) {
            mbar_init(&barF1[i], 1);
            mbar_init(&barF1Free[i], NCONS);
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            mbar_init(&barFull[i], HAS_FLOW ? NBIL : 1);
            mbar_init(&barEmpty[i], NCONS);
        }
#pragma unroll
        for (int i = 0; i < NWIN; ++i) {
            mbar_init(&barWin[i], 1);
            mbar_init(&barWinFree[i], NBIL);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mbar_init(&barFlow[i], 1);
            mbar_init(&barFlowFree[i], 32);
            mbar_init(&barTaps[i], 32);
            mbar_init(&barTapsFree[i], NBIL);
        }
        fence_mbar_init();
    }
    __syncthreads();     // the only block-wide barrier; roles split below

    FwdArgs A;
    A.tmF1 = &tmF1; A.tmF2 = &tmF2; A.tmFlow = &tmFlow; A.f2 = f2; A.flow = flow; A.out = out; A.warped_out = warped_out;
    A.C = C; A.H = H; A.W = W; A.tiles_x = tiles_x; A.tiles_y = tiles_y; A.ntiles = ntiles; A.act = act; A.slope = slope;
    A.obs = obs; A.fbs = fbs; A.coarse = coarse; A.flow_out = flow_out; A.fobs = fobs;
#ifndef PWC_REG_SPLIT
#define PWC_REG_SPLIT 1
#endif
    // Register split (setmaxnreg moves registers between groups of 4 warps inside the CTA's launch allocation of
    // 16 x 32 x 128): warps 12..15 (two bilinear warps, P, T) give up 48 registers each, warps 0..11 (the nine
    // correlation warps and three bilinear warps) take 16 more -- 12 * 144 + 4 * 80 = 16 * 128.  At 128 registers the
    // correlation role spilled loop-invariant values and re-read tid (S2R) in every tile's epilogue.  Measured at the
    // level-2 shape: 111 -> 109 us (i.i.d. flow), 107 -> 105 (smooth), 84 -> 81 (no flow); 152 / 56 makes the bilinear
    // warps spill (135 us).  stride2 = 2 keeps the plain layout: its bilinear warps hold 7 taps each and spill at 80
    // registers (162 -> 188 us).
    if (PWC_REG_SPLIT && Cfg::NT == 512 && Cfg::S2 == 1) {
        if (tid >= 384) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
            if (tid >= NCONS + NBIL + 32) {
                if (tid == NCONS + NBIL + 32) fwd_role_tma<Cfg, HAS_FLOW>(A);
            } else if (tid >= NCONS + NBIL) {
                if (HAS_FLOW) fwd_role_taps<Cfg, HAS_FLOW>(A);
            } else {
                if (HAS_FLOW) fwd_role_bilinear<Cfg, HAS_FLOW>(A);
            }
            return;
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        if (tid >= NCONS) {
            if (HAS_FLOW) fwd_role_bilinear<Cfg, HAS_FLOW>(A);
            return;
        }
    } else {
    if (tid >= NCONS + NBIL + 32) {
        if (tid == NCONS + NBIL + 32) fwd_role_tma<Cfg, HAS_FLOW>(A);
        return;
    }
    if (tid >= NCONS + NBIL) {
        if (HAS_FLOW) fwd_role_taps<Cfg, HAS_FLOW>(A);
        return;
    }
    if (tid >= NCONS) {
        if (HAS_FLOW) fwd_role_bilinear<Cfg, HAS_FLOW>(A);
        return;
    }
    }

    // ================================ C: correlation warps ================================
    // Task of this thread: displacement row wd (tj = wd - r), tile row lr, 8-pixel strip ls.
    const int lane = tid & 31;
    int wd, lr, ls;
#ifndef PWC_QUAD_MAP
#define PWC_QUAD_MAP 1
#endif
    if (PWC_QUAD_MAP && S2 == 1) {
        // Quad sharing: an LDS.128 costs 4 clocks when all 32 lanes read different words but 2.1 when the lanes of
        // every aligned group of four read at most two different addresses (equal addresses are merged inside a quad
        // only, scripts/lds_probe.cu).  A quad = 2 pixel rows x 2 warped rows -- tasks (y, d0), (y, d0+1), (y+1, d0-1),
        // (y+1, d0) -- reads 2 f1 rows and 2 warped rows; a half warp stacks 4 such row pairs.  d0 = 1, 3, 5, 7 covers
        // every task except (even row, 0) and (odd row, 8), which the ninth warp takes without sharing.
        const int w9 = tid >> 5;
        if (w9 < 8) {
            const int hw = 2 * w9 + (lane >> 4), qd = (lane >> 2) & 3, e = lane & 3;
            lr = 8 * (hw >> 3) + 2 * qd + (e >> 1);
            ls = (hw >> 2) & 1;
            wd = 1 + 2 * (hw & 3) + (e & 1) - (e >> 1);
        } else {
            lr = lane & 15; ls = lane >> 4; wd = (lr & 1) ? D - 1 : 0;
        }
    } else {
        wd = tid >> 5; lr = lane & 15; ls = lane >> 4;     // rows fastest: a quarter warp spans 8 rows of one strip
    }
    // 1/C (correlation_cuda_kernel.cu:65,100 divide by nelems; a correctly rounded reciprocal and one
    // multiply differ from the division by at most 1 ulp, far inside the 1e-5 tolerance)
    const float inv_nelems = __frcp_rn((float)C);
    const bool out_32B_aligned = ((W & 7) == 0) && ((reinterpret_cast<uintptr_t>(out) & 31) == 0) && ((obs & 7) == 0);
    int g = 0;      // (one register: this role sits at the 128-register cap, ring structs here spilled into the FFMA loop)
    for (int lt = 0; lt < my_tiles; ++lt) {
        const TileCoord tc = tile_coord(blockIdx.x + lt * gridDim.x, tiles_x, tiles_y, TH, TW);
        float acc[PX][D];
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int d = 0; d < D; ++d) acc[p][d] = 0.0f;
#ifndef PWC_FFMA2
#define PWC_FFMA2 1
#endif
        // Packed variant (stride2 = 1): the accumulators (p, d) and (p + 1, d - 1), p even, d = 1..8, both multiply the
        // warped value w[p + d], which FFMA2 takes as its scalar (broadcast) operand, with the natural register pair
        // (f[p], f[p + 1]); (even p, d = 0) and (odd p, d = 8) stay scalar.  32 FFMA2 + 8 FFMA instead of 72 FFMA per channel.
        constexpr bool PACKED = PWC_FFMA2 && S2 == 1;
        float2 ap[PX / 2][D - 1];
#pragma unroll
        for (int p = 0; p < PX / 2; ++p)
#pragma unroll
            for (int d = 0; d < D - 1; ++d) ap[p][d] = make_float2(0.0f, 0.0f);

        // A chunk's "consumed" signals are given one loop iteration late (or after the epilogue stores for
        // the last chunk of a tile): an mbarrier arrive issued right behind a still-pending LDS can overtake
        // it, and the producer's next write then corrupts the value being read (measured in corr_bwd_tma.cuh).
        // One iteration later every instruction of the chunk has issued, so all its loads have landed.
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++g) {
            const int s = g % NS, sf = g % NF1;
            mbar_wait(&barF1[sf], (g / NF1) & 1);
            mbar_wait(&barFull[s], (g / NS) & 1);
            const float* pf = sF1 + sf * Cfg::F1_ELEMS + lr * F1W + ls * PX;
            const float* pw = sW2 + s * Cfg::W2_ELEMS + (lr + wd * S2) * WP + ls * PX;
            // Software-pipelined at 128-bit granularity: a warped-row quad is reloaded for the next
            // channel right after its last use, so every LDS has a whole channel of FFMAs to land.
            float f[PX], w[WSPAN];
#pragma unroll
            for (int q = 0; q < PX / 4; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(pf + 4 * q);
                f[4 * q] = v4.x; f[4 * q + 1] = v4.y; f[4 * q + 2] = v4.z; f[4 * q + 3] = v4.w;
            }
#pragma unroll
            for (int q = 0; q < WSPAN / 4; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(pw + 4 * q);
                w[4 * q] = v4.x; w[4 * q + 1] = v4.y; w[4 * q + 2] = v4.z; w[4 * q + 3] = v4.w;
            }
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                float fn[PX];
                if (c + 1 < CK) {
#pragma unroll
                    for (int q = 0; q < PX / 4; ++q) {
                        const float4 v4 = *reinterpret_cast<const float4*>(pf + (c + 1) * (F1H * F1W) + 4 * q);
                        fn[4 * q] = v4.x; fn[4 * q + 1] = v4.y; fn[4 * q + 2] = v4.z; fn[4 * q + 3] = v4.w;
                    }
                }
#pragma unroll
                for (int q = 0; q < WSPAN / 4; ++q) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int jj = 4 * q + e;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const int p = jj - d * S2;
#if defined(PWC_DEV_X) && PWC_DEV_X == 4
                            if (d > 1) continue;                // (dev ablation) ~1/5 of the FMAs, all loads kept
#endif
                            if (PACKED) {
                                if (p >= 0 && p < PX && (p & 1) == 0 && d >= 1)
                                    ap[p >> 1][d - 1] = __ffma2_rn(make_float2(f[p], f[p + 1]), make_float2(w[jj], w[jj]), ap[p >> 1][d - 1]);
                                if (p >= 0 && p < PX && (((p & 1) == 0 && d == 0) || ((p & 1) == 1 && d == D - 1)))
                                    acc[p][d] = fmaf(f[p], w[jj], acc[p][d]);
                                continue;
                            }
                            if (p >= 0 && p < PX) acc[p][d] = fmaf(f[p], w[jj], acc[p][d]);
                        }
                    }
                    if (c + 1 < CK) {
                        const float4 v4 = *reinterpret_cast<const float4*>(pw + (c + 1) * (HH * WP) + 4 * q);
                        w[4 * q] = v4.x; w[4 * q + 1] = v4.y; w[4 * q + 2] = v4.z; w[4 * q + 3] = v4.w;
                    }
                }
                if (c + 1 < CK) {
#pragma unroll
                    for (int p = 0; p < PX; ++p) f[p] = fn[p];
                }
            }
            if (k > 0) {
                mbar_arrive(&barEmpty[(g - 1) % NS]);       // warped chunk slot of the previous chunk consumed
                mbar_arrive(&barF1Free[(g - 1) % NF1]);     // f1 chunk slot of the previous chunk consumed
            }
        }

        // ---- epilogue: 1/C, optional LeakyReLU (model.py:84) ----
        const int y = tc.y0 + lr;
        const int xs = tc.x0 + ls * PX;
        if (y < H && xs < W) {   // W % 4 == 0 and xs % 8 == 0: a strip is fully inside or ends on a multiple of 4
            const bool wide = out_32B_aligned && xs + PX <= W;     // the strip is one aligned 32-byte sector
            // obs: output batch stride; the nine displacement planes of this thread are H * W floats apart
            float* o = out + (size_t)tc.n * (size_t)obs + ((size_t)(wd * D) * H + y) * W + xs;
            const size_t plane = (size_t)H * W;
#pragma unroll
            for (int d = 0; d < D; ++d, o += plane) {
                float v[PX];
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    float a = acc[p][d];
                    if (PACKED) {          // (even p, d >= 1) -> ap[p/2][d-1].x, (odd p, d <= 7) -> ap[p/2][d].y
                        if ((p & 1) == 0 && d >= 1) a = ap[p >> 1][d - 1].x;
                        if ((p & 1) == 1 && d <= D - 2) a = ap[p >> 1][d].y;
                    }
                    v[p] = a * inv_nelems;
                }
                if (act) {      // one uniform branch per displacement, selects inside (not a branch per element)
#pragma unroll
                    for (int p = 0; p < PX; ++p) v[p] = v[p] < 0.0f ? v[p] * slope : v[p];
                }
                if (wide) {
                    st_global_v8(o, v);
                } else {
                    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                    if (xs + 4 < W) *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        // last chunk of the tile: its loads fed the accumulators that were just stored (threads outside the
        // image store nothing, but they are a whole epilogue past the loads as well)
        mbar_arrive(&barEmpty[(g - 1) % NS]);
        mbar_arrive(&barF1Free[(g - 1) % NF1]);
    }
}

}  // namespace pwc
