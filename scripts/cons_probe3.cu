// Consumer-loop probe, second design (development only): W-row-sharing tasks + packed FFMA2.
// A thread owns one warped row r of a 4-pixel strip and up to three (pixel row y, displacement row dy = r - y)
// tasks that all read that warped row: per channel 12 + 12 loaded floats feed 108 FMAs (4.5 per loaded float
// instead of 3.0), issued as 54 FFMA2 + 10 MOV.  Static operands in shared memory: no TMA, no mbarriers.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int D = 9, CK = 4, TH = 16, TW = 16, R = 4;
constexpr int HH = TH + 2 * R, HWD = TW + 2 * R, WP = HWD + 4, F1W = TW + 4;
constexpr int F1_ELEMS = CK * TH * F1W, W2_ELEMS = CK * HH * WP;
constexpr int STAGE = F1_ELEMS + W2_ELEMS;
constexpr int NROWTHR = 54;        // (warped row, task group) pairs per strip

struct RowTab { unsigned char r[NROWTHR], y0[NROWTHR], cnt[NROWTHR]; };
static RowTab make_tab()
{
    RowTab t; int n = 0;
    for (int r = 0; r < HH; ++r) {
        const int ylo = r - 8 > 0 ? r - 8 : 0, yhi = r < TH - 1 ? r : TH - 1;
        for (int y0 = ylo; y0 <= yhi; y0 += 3) {
            t.r[n] = r; t.y0[n] = y0; t.cnt[n] = (yhi - y0 + 1 < 3) ? yhi - y0 + 1 : 3; ++n;
        }
    }
    if (n != NROWTHR) printf("table size %d\n", n);
    return t;
}
__constant__ RowTab c_tab;

__device__ unsigned long long g_clk[2];
__global__ void spin(float* o, int n) { float a = o[threadIdx.x]; for (int i = 0; i < n; ++i) a = fmaf(a, 1.0001f, 0.5f); o[threadIdx.x] = a; }
// MODE 0: full loop; 1: FMA only (operands loaded once per chunk); 2: LDS only
template <int NT, int NST, bool STORE, int MODE = 0, int VAR = 0>
__global__ void __launch_bounds__(NT, 1) probe3(float* __restrict__ out, int my_tiles, int nchunks, int H, int W)
{
    unsigned long long c0 = clock64(), t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    extern __shared__ __align__(1024) float sm[];
    for (int i = threadIdx.x; i < NST * STAGE; i += NT) sm[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
    __syncthreads();
    const int tid = threadIdx.x;
    const int s4 = (tid >> 3) & 3, m = (tid >> 5) * 8 + (tid & 7);
    if (m >= NROWTHR) return;
    const int r = c_tab.r[m], y0 = c_tab.y0[m], cnt = c_tab.cnt[m];
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
        // accumulators: pixel p of task t holds displacement pairs (d, d+1) with p + d even -- the warped pair
        // (w[p+d], w[p+d+1]) is then an aligned register pair of the LDS.128 result and f[p] is the FFMA2
        // scalar-broadcast operand -- plus one single displacement (d = 8 for even p, d = 0 for odd p)
        float2 acc2[3][4][4];
        float acc1[3][4];
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                acc1[t][p] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc2[t][p][i] = make_float2(0.f, 0.f);
            }
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k, ++g) {
            const float* st = sm + (g % NST) * STAGE;
            const float* pw = st + F1_ELEMS + r * WP + 4 * s4;
            const float* pf[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) pf[t] = st + (y0 + (t < cnt ? t : 0)) * F1W + 4 * s4;
            float w[12], f[3][4];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(pw + 4 * q);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const float4 v = *reinterpret_cast<const float4*>(pf[t]);
                f[t][0] = v.x; f[t][1] = v.y; f[t][2] = v.z; f[t][3] = v.w;
            }
#pragma unroll
            for (int c = 0; c < CK; ++c) {
                float wn[12], fn[3][4];
                if (c + 1 < CK && MODE != 1) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(pw + (c + 1) * (HH * WP) + 4 * q);
                        wn[4 * q] = v.x; wn[4 * q + 1] = v.y; wn[4 * q + 2] = v.z; wn[4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const float4 v = *reinterpret_cast<const float4*>(pf[t] + (c + 1) * (TH * F1W));
                        fn[t][0] = v.x; fn[t][1] = v.y; fn[t][2] = v.z; fn[t][3] = v.w;
                    }
                }
                if (MODE == 2) {
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc1[t][p] += f[t][p] + w[t * 4 + p];
                } else if (VAR == 2) {
                    // warped pair outermost: consecutive FFMA2 share the 64-bit b operand (register reuse cache), read one
                    // scalar a and one accumulator pair -- the "b outer" pattern of scripts/fma_probe2.cu (2.0 clk)
#pragma unroll
                    for (int jp = 0; jp < 6; ++jp) {
                        const float2 b = make_float2(w[2 * jp], w[2 * jp + 1]);
#pragma unroll
                        for (int t = 0; t < 3; ++t)
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const int i2 = 2 * jp - p - (p & 1);
                                if (i2 >= 0 && i2 <= 6) acc2[t][p][i2 >> 1] = __ffma2_rn(make_float2(f[t][p], f[t][p]), b, acc2[t][p][i2 >> 1]);
                            }
                    }
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc1[t][p] = fmaf(f[t][p], w[(p & 1) ? p : p + 8], acc1[t][p]);
                } else {
#pragma unroll
                for (int t = 0; t < 3; ++t)
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float2 fb = make_float2(f[t][p], f[t][p]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int j = p + 2 * i + (p & 1);
                            if (VAR == 0) acc2[t][p][i] = __ffma2_rn(fb, make_float2(w[j], w[j + 1]), acc2[t][p][i]);
                            else {
                                acc2[t][p][i].x = fmaf(f[t][p], w[j], acc2[t][p][i].x);
                                acc2[t][p][i].y = fmaf(f[t][p], w[j + 1], acc2[t][p][i].y);
                            }
                        }
                        acc1[t][p] = fmaf(f[t][p], w[(p & 1) ? p : p + 8], acc1[t][p]);
                    }
                }
                if (c + 1 < CK && MODE != 1) {
#pragma unroll
                    for (int j = 0; j < 12; ++j) w[j] = wn[j];
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int p = 0; p < 4; ++p) f[t][p] = fn[t][p];
                }
            }
        }
        const int tile = blockIdx.x + lt * gridDim.x;
        const int tiles_x = W / TW, tiles_y = H / TH;
        const int x0 = (tile % tiles_x) * TW, ty0 = ((tile / tiles_x) % tiles_y) * TH, n = tile / (tiles_x * tiles_y);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (t < cnt) {
                const int y = y0 + t, dy = r - y;
                float* o = out + (size_t)n * 81 * H * W + ((size_t)(dy * D) * H + ty0 + y) * W + x0 + 4 * s4;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    float v[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        // displacement d of pixel p: even p -> pair d/2 (.x even d, .y odd d), single d = 8;
                        // odd p -> single d = 0, pair (d-1)/2 (.x odd d, .y even d)
                        float a;
                        if ((p & 1) == 0) a = d == 8 ? acc1[t][p] : ((d & 1) ? acc2[t][p][d >> 1].y : acc2[t][p][d >> 1].x);
                        else a = d == 0 ? acc1[t][p] : ((d & 1) ? acc2[t][p][(d - 1) >> 1].x : acc2[t][p][(d - 1) >> 1].y);
                        v[p] = a * 0.03125f;
                    }
                    if (STORE || v[0] == 123.456f) *reinterpret_cast<float4*>(o + (size_t)d * H * W) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        g_clk[0] = clock64() - c0; g_clk[1] = t1 - t0;
    }
}

template <int NT, int NST, bool STORE, int MODE = 0, int VAR = 0>
static void run(const char* name, float* dout, int B, int C, int H, int W)
{
    auto kern = probe3<NT, NST, STORE, MODE, VAR>;
    const size_t smem = (size_t)NST * STAGE * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ntiles = B * (H / TH) * (W / TW);
    const int grid = 148, my_tiles = (ntiles + grid - 1) / grid;
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
    unsigned long long clk[2];
    cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk));
    const double mhz = 1e3 * (double)clk[0] / (double)clk[1];
    const double fma = 81.0 * C * 256 * my_tiles;
    printf("%-34s B%d C%d %dx%d: %7.2f us  SM clock %.0f MHz  %.0f cycles/chunk, useful FMA %.0f%% of 128/clk %s\n", name, B, C, H, W,
           1e3 * best, mhz, (double)clk[0] / (my_tiles * (C / CK)), 100.0 * fma / 128.0 / (double)clk[0],
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main()
{
    const RowTab t = make_tab();
    cudaMemcpyToSymbol(c_tab, &t, sizeof(t));
    const int B = 32, H = 96, W = 112;
    float* dout;
    cudaMalloc(&dout, (size_t)(B + 4) * 81 * H * W * 4);
    spin<<<148, 256>>>(dout, 20000000);
    cudaDeviceSynchronize();
    run<224, 8, true>("row-share FFMA2 224 thr, stores", dout, B, 32, H, W);
    run<224, 8, false>("row-share FFMA2, no stores", dout, B, 32, H, W);
    run<224, 8, false, 1>("row-share FFMA2, FMA only", dout, B, 32, H, W);
    run<224, 8, false, 2>("row-share FFMA2, LDS only", dout, B, 32, H, W);
    run<224, 8, true, 0, 2>("row-share FFMA2 b-outer, stores", dout, B, 32, H, W);
    run<224, 8, false, 1, 2>("row-share FFMA2 b-outer, FMA only", dout, B, 32, H, W);
    run<224, 8, true, 0, 1>("row-share scalar FFMA, stores", dout, B, 32, H, W);
    run<224, 8, false, 1, 1>("row-share scalar FFMA, FMA only", dout, B, 32, H, W);
    return 0;
}
