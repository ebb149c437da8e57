// FP32 issue-rate probe (development only): scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on an 8x8 register outer
// product, the pattern of every SIMT correlation kernel in this repository.  Reports lane-FMAs per clock and SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/dev/fma_probe scripts/fma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ in, float* __restrict__ out, int iters)
{
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; b[i] = in[threadIdx.x + 32 * i + 1000]; }
    if (MODE == 0) {
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += acc[i][j];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        float2 acc[8][4], ad[8], b2[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) ad[i] = make_float2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) b2[j] = make_float2(b[2 * j], b[2 * j + 1]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(ad[i], b2[j], acc[i][j]);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <int MODE>
static void run(const char* name, int nthr, float* din, float* dout)
{
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int t = 0; t < 5; ++t) {
        cudaEventRecord(e0);
        k<MODE><<<148, nthr>>>(din, dout, iters);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fma = 256.0 * iters * nthr;      // lane-FMAs per SM
    printf("%-28s %4d thr: %8.3f ms  %.1f lane-FMA/clk/SM (@1.965 GHz)  %s\n", name, nthr, best,
           fma / (best * 1e-3 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float *din, *dout;
    cudaMalloc(&din, 1 << 20); cudaMemset(din, 0, 1 << 20); cudaMalloc(&dout, 1 << 22);
    for (int nthr : {128, 256, 512}) { run<0>("FFMA 8x8", nthr, din, dout); run<1>("FFMA2 8x4 pairs", nthr, din, dout); }
    return 0;
}
