// Consumer-loop probe (development only, not part of the product): how fast can the correlate stage of the fused
// forward kernel run when its operands already sit in shared memory?  No TMA, no mbarriers: the staged chunks are
// static, so the time is the consumer loop's own (LDS.128 + FFMA + epilogue stores).
//   MAP 0: the round-1 mapping -- warp = displacement row, lane = (tile row, 8-pixel strip): every lane of a warp
//          reads different shared-memory words (24 wavefronts per channel and warp).
//   MAP 1: slab mapping -- 36 consecutive threads share a 4-row block of one strip (9 displacement rows x 4 rows),
//          so the f1 rows and most warped rows are read by several lanes of a warp at once (broadcast: ~10
//          wavefronts per channel and warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o /tmp/cons_probe scripts/cons_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int PX = 8, D = 9, CK = 4, TH = 16, TW = 16, R = 4;
constexpr int HH = TH + 2 * R, HWD = TW + 2 * R, WP = HWD + 4, F1W = TW + 4;
constexpr int F1_ELEMS = CK * TH * F1W, W2_ELEMS = CK * HH * WP;
constexpr int STAGE = F1_ELEMS + W2_ELEMS;
constexpr int WSPAN = PX + 2 * R;

__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8])
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

template <int MAP, int NT, int NST, bool STORE, int NW = 9, int MODE = 0, int FF2 = 0>
__global__ void __launch_bounds__(NT, 1) probe(float* __restrict__ out, int my_tiles, int nchunks, int H, int W)
{
    extern __shared__ __align__(1024) float sm[];
    for (int i = threadIdx.x; i < NST * STAGE; i += NT) sm[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    __syncthreads();
    const int tid = threadIdx.x;
    if (tid >= 32 * NW) return;
    int y, s, dy;
    if (MAP == 3) {
        // quad sharing: the hardware merges equal addresses only inside aligned groups of 4 lanes (lds_probe.cu).
        // A quad = 2 pixel rows x 2 warped rows: tasks (y,d0) (y,d0+1) (y+1,d0-1) (y+1,d0) read 2 distinct f1 rows and 2
        // distinct warped rows -> every LDS.128 of the warp touches 16 distinct chunks: 2 clocks instead of 4.
        const int w9 = (tid >> 5) % 9, lane = tid & 31;
        if (w9 < 8) {
            const int hw = 2 * w9 + (lane >> 4), qd = (lane >> 2) & 3, e = lane & 3;
            y = 8 * (hw >> 3) + 2 * qd + (e >> 1); s = (hw >> 2) & 1; dy = 1 + 2 * (hw & 3) + (e & 1) - (e >> 1);
        } else {
            y = lane & 15; s = lane >> 4; dy = (y & 1) ? 8 : 0;     // the leftovers: (even row, dy 0), (odd row, dy 8)
        }
    } else if (MAP == 2 && (tid >> 5) % 9 < 8) {
        // half-warp = 4 pixel rows x 4 displacement rows of one strip: 4 distinct f1 rows and 7 distinct warped rows
        // per LDS.128, i.e. <= 128 distinct bytes per half-warp -> 2 clocks instead of 4 (scripts/lds_probe.cu)
        const int hw = 2 * ((tid >> 5) % 9) + ((tid & 31) >> 4), l = tid & 15;
        dy = 4 * (hw & 1) + (l >> 2); s = (hw >> 1) & 1; y = 4 * (hw >> 2) + (l & 3);
    } else if (MAP == 0 || MAP == 2) {
        dy = (tid >> 5) % 9; y = tid & 15; s = (tid & 31) >> 4;
    } else {
        const int grp = (tid % 288) / 36, rem = tid % 36;
        dy = rem >> 2; y = (grp & 3) * 4 + (rem & 3); s = grp >> 2;
    }
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
        float acc[PX][D];
#pragma unroll
        for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int d = 0; d < D; ++d) acc[p][d] = 0.0f;
        // FF2: anti-diagonal accumulator pairs (acc[p][d], acc[p+1][d-1]), p even, d = 1..8: both take the warped value
        // w[p+d] (FFMA2 scalar-broadcast operand) and the natural register pair (f[p], f[p+1])
        float2 ap[PX / 2][D - 1];
#pragma unroll
        for (int p = 0; p < PX / 2; ++p)
#pragma unroll
            for (int d = 0; d < D - 1; ++d) ap[p][d] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++g) {
            const float* st = sm + (g % NST) * STAGE;
            const float* pf = st + y * F1W + s * PX;
            const float* pw = st + F1_ELEMS + (y + dy) * WP + s * PX;
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
                if (c + 1 < CK && MODE != 1) {
#pragma unroll
                    for (int q = 0; q < PX / 4; ++q) {
                        const float4 v4 = *reinterpret_cast<const float4*>(pf + (c + 1) * (TH * F1W) + 4 * q);
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
                            const int p = jj - d;
                            if (FF2) {
                                if (p >= 0 && p < PX && (p & 1) == 0 && d >= 1)
                                    ap[p >> 1][d - 1] = __ffma2_rn(make_float2(f[p], f[p + 1]), make_float2(w[jj], w[jj]), ap[p >> 1][d - 1]);
                                if (p >= 0 && p < PX && (((p & 1) == 0 && d == 0) || ((p & 1) == 1 && d == 8)))
                                    acc[p][d] = fmaf(f[p], w[jj], acc[p][d]);
                                continue;
                            }
                            if (MODE != 2 && p >= 0 && p < PX) acc[p][d] = fmaf(f[p], w[jj], acc[p][d]);
                        }
                        if (MODE == 2) acc[jj & 7][jj >> 3] += w[jj] + (jj < PX ? f[jj] : 0.0f);
                    }
                    if (c + 1 < CK && MODE != 1) {
                        const float4 v4 = *reinterpret_cast<const float4*>(pw + (c + 1) * (HH * WP) + 4 * q);
                        w[4 * q] = v4.x; w[4 * q + 1] = v4.y; w[4 * q + 2] = v4.z; w[4 * q + 3] = v4.w;
                    }
                }
                if (c + 1 < CK && MODE != 1) {
#pragma unroll
                    for (int p = 0; p < PX; ++p) f[p] = fn[p];
                }
            }
        }
        // epilogue: the tile's 81 x 16 x 16 outputs
        const int tile = blockIdx.x + lt * gridDim.x;
        const int tiles_x = W / TW, tiles_y = H / TH;
        const int x0 = (tile % tiles_x) * TW, y0 = ((tile / tiles_x) % tiles_y) * TH, n = tile / (tiles_x * tiles_y);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float v[PX];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                float a = acc[p][d];
                if (FF2) {
                    if ((p & 1) == 0 && d >= 1) a = ap[p >> 1][d - 1].x;
                    if ((p & 1) == 1 && d <= 7) a = ap[p >> 1][d].y;
                }
                v[p] = a * 0.03125f;
            }
            float* o = out + (size_t)n * 81 * H * W + ((size_t)(dy * D + d) * H + y0 + y) * W + x0 + s * PX;
            if (STORE) st_global_v8(o, v);
            else if (v[0] == 123.456f) st_global_v8(o, v);
        }
    }
}

template <int MAP, int NT, int NST, bool STORE, int NW = 9, int MODE = 0, int FF2 = 0>
static void run(const char* name, float* dout, int B, int C, int H, int W)
{
    auto kern = probe<MAP, NT, NST, STORE, NW, MODE, FF2>;
    const size_t smem = (size_t)NST * STAGE * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ntiles = B * (H / TH) * (W / TW);
    const int grid = 148, my_tiles = (ntiles + grid - 1) / grid;      // every CTA runs the maximum (the tail round)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 8; ++it) {
        cudaEventRecord(e0);
        kern<<<grid, NT, smem>>>(dout, my_tiles, C / CK, H, W);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double fma = 81.0 * C * 256 * my_tiles * NW / 9.0;                    // per SM
    printf("%-34s B%d C%d %dx%d: %7.2f us  (%.0f cycles/chunk @1.965 GHz, FMA pipe %.0f%%) %s\n", name, B, C, H, W,
           1e3 * best, best * 1e-3 * 1.965e9 / (my_tiles * (C / CK)), 100.0 * fma / 128.0 / (best * 1e-3 * 1.965e9),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const int B = 32, H = 96, W = 112;
    float* dout;
    cudaMalloc(&dout, (size_t)(B + 4) * 81 * H * W * 4);
    run<0, 288, 8, true, 9>("map0 9 warps FFMA, stores", dout, B, 32, H, W);
    run<3, 288, 8, true, 9>("map3 (quad sharing) 9 warps", dout, B, 32, H, W);
    run<3, 288, 8, true, 9, 0, 1>("map3 9 warps FFMA2", dout, B, 32, H, W);
    run<0, 256, 8, true, 8>("map0 8 warps FFMA, stores", dout, B, 32, H, W);
    run<3, 256, 8, true, 8>("map3 8 warps FFMA, stores", dout, B, 32, H, W);
    run<3, 256, 8, true, 8, 0, 1>("map3 8 warps FFMA2, stores", dout, B, 32, H, W);
    run<0, 512, 8, true, 16>("map0 16 warps FFMA, stores", dout, B, 32, H, W);
    run<3, 512, 8, true, 16>("map3 16 warps FFMA, stores", dout, B, 32, H, W);
    run<3, 512, 8, true, 16, 0, 1>("map3 16 warps FFMA2, stores", dout, B, 32, H, W);
    return 0;
}
