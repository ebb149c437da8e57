// FP32 issue-rate probe 2 (development only): how many register operands can an FFMA / FFMA2 read per clock?
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long g_clk;
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ in, float* __restrict__ out, int iters)
{
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 1000]; }
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = in[i];
    float2* acc2 = reinterpret_cast<float2*>(acc);
    const unsigned long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (MODE == 0) {          // one varying operand: acc only
#pragma unroll
                for (int i = 0; i < 64; ++i) acc[i] = fmaf(a[0], b[0], acc[i]);
            } else if (MODE == 1) {   // a varies slowly, b fixed
#pragma unroll
                for (int i = 0; i < 64; ++i) acc[i] = fmaf(a[i >> 3], b[0], acc[i]);
            } else if (MODE == 2) {   // 8x8 outer product
#pragma unroll
                for (int i = 0; i < 64; ++i) acc[i] = fmaf(a[i >> 3], b[i & 7], acc[i]);
            } else if (MODE == 3) {   // FFMA2, a and b fixed
#pragma unroll
                for (int i = 0; i < 32; ++i) acc2[i] = __ffma2_rn(make_float2(a[0], a[0]), make_float2(b[0], b[1]), acc2[i]);
            } else if (MODE == 4) {   // FFMA2 8x4 outer product, scalar-broadcast a
#pragma unroll
                for (int i = 0; i < 32; ++i) acc2[i] = __ffma2_rn(make_float2(a[i >> 2], a[i >> 2]), make_float2(b[2 * (i & 3)], b[2 * (i & 3) + 1]), acc2[i]);
            } else if (MODE == 6 || MODE == 7) {   // mixed: FFMA2 on one half of the accumulators, scalar FFMA on the other
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    acc2[(i & 7) * 2 + (i >> 3)] = __ffma2_rn(make_float2(a[i & 7], a[i & 7]), make_float2(b[2 * (i >> 3)], b[2 * (i >> 3) + 1]), acc2[(i & 7) * 2 + (i >> 3)]);
#pragma unroll
                    for (int u = 0; u < (MODE == 6 ? 2 : 1); ++u) acc[32 + 2 * i + u] = fmaf(a[0], b[0], acc[32 + 2 * i + u]);
                }
            } else if (MODE == 5) {   // FFMA2 8x4, b outer (b pair reused), a scalar varies
#pragma unroll
                for (int i = 0; i < 32; ++i) acc2[(i & 7) * 4 + (i >> 3)] = __ffma2_rn(make_float2(a[i & 7], a[i & 7]), make_float2(b[2 * (i >> 3)], b[2 * (i >> 3) + 1]), acc2[(i & 7) * 4 + (i >> 3)]);
            }
        }
    }
    const unsigned long long c1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) g_clk = c1 - c0;
}
template <int MODE>
static void run(const char* name, int nthr, float* din, float* dout)
{
    const int iters = 2000;
    for (int t = 0; t < 3; ++t) k<MODE><<<148, nthr>>>(din, dout, iters);
    cudaDeviceSynchronize();
    unsigned long long clk;
    cudaMemcpyFromSymbol(&clk, g_clk, sizeof(clk));
    const double fmas = MODE == 6 ? 256.0 : MODE == 7 ? 192.0 : 256.0;           // lane-FMAs per thread and iteration
    const double insts = MODE == 6 ? 4 * 48.0 : MODE == 7 ? 4 * 32.0 : (MODE >= 3 ? 128.0 : 256.0);
    const double winst = insts * iters * nthr / 32 / 4;     // warp instructions per scheduler
    printf("%-44s %4d thr: %.2f clk per warp instruction and scheduler, %.1f lane-FMA/clk/SM  %s\n", name, nthr,
           clk / winst, fmas * iters * nthr / clk, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float *din, *dout;
    cudaMalloc(&din, 1 << 20); cudaMemset(din, 0, 1 << 20); cudaMalloc(&dout, 1 << 22);
    for (int nthr : {128, 512}) {
        run<0>("FFMA acc only varies", nthr, din, dout);
        run<1>("FFMA a slow, b fixed", nthr, din, dout);
        run<2>("FFMA 8x8 outer", nthr, din, dout);
        run<3>("FFMA2 a,b fixed", nthr, din, dout);
        run<4>("FFMA2 8x4 outer, a outer", nthr, din, dout);
        run<5>("FFMA2 8x4 outer, b outer", nthr, din, dout);
        run<6>("mixed: 16 FFMA2 + 32 FFMA per round", nthr, din, dout);
        run<7>("mixed: 16 FFMA2 + 16 FFMA per round", nthr, din, dout);
    }
    return 0;
}
