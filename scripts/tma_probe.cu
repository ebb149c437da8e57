// Standalone probe: 4-D TMA box load from a [B][C][H][W] fp32 tensor into smem (zero fill out of range).
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/tma_probe.cu && /tmp/tma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../pwc_net_pytorch_b200/csrc/warpcorr_fwd_tma.cuh"

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int bw, int bh, int bc, int x, int y, int c, int n, int mode)
{
    extern __shared__ uint8_t raw[];
    uint8_t* base = (uint8_t*)(((uintptr_t)raw + 127) & ~uintptr_t(127));
    uint64_t* bar = (uint64_t*)base;
    float* buf = (float*)(base + 128);
    if (threadIdx.x == 0) {
        pwc::mbar_init(bar, 1);
        pwc::fence_mbar_init();
    }
    __syncthreads();
    if (mode >= 1 && threadIdx.x == 0) pwc::prefetch_tmap(&tm);
    if (threadIdx.x == 0) {
        pwc::mbar_expect_tx(bar, bw * bh * bc * 4);
        pwc::tma_load_4d(buf, &tm, bar, x, y, c, n);
    }
    pwc::mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bw * bh * bc; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv)
{
    int only = argc > 1 ? atoi(argv[1]) : -1; int idx = -1;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    printf("encode fn %p status %d\n", p, (int)q);
    struct Case { int B, C, H, W, bw, bh, bc, x, y, c, n, mode; } cases[] = {
        {2, 8, 96, 112, 40, 40, 4, 10, 7, 0, 1, 0},
        {2, 8, 96, 112, 40, 40, 4, -3, -5, 4, 1, 1},
        {2, 6, 96, 112, 20, 16, 4, 100, 90, 4, 0, 1},
        {1, 3, 24, 28, 48, 64, 2, -9, -9, 2, 0, 1},
        {1, 3, 8, 16, 28, 24, 4, -4, -4, 0, 0, 1},
        {2, 8, 96, 112, 32, 8, 1, 0, 0, 0, 0, 0},
        {2, 8, 96, 112, 32, 8, 4, 16, 8, 0, 1, 0},
        {2, 8, 96, 112, 16, 40, 4, 16, 8, 0, 1, 0},
        {2, 8, 96, 112, 64, 4, 1, 16, 8, 0, 1, 0},
        {2, 8, 96, 112, 40, 4, 1, 16, 8, 0, 1, 0},
        {2, 8, 96, 112, 40, 40, 1, -3, 8, 0, 1, 0},   // 10: 6.4 KB
        {2, 8, 96, 112, 40, 40, 2, -3, 8, 0, 1, 0},   // 11: 12.8 KB
        {2, 8, 96, 112, 64, 64, 1, -3, 8, 0, 1, 0},   // 12: 16 KB
        {2, 8, 96, 112, 48, 48, 2, -3, 8, 0, 1, 0},   // 13: 18.4 KB
        {2, 8, 96, 112, 40, 40, 3, -3, 8, 0, 1, 0},   // 14: 19.2 KB
        {2, 8, 96, 112, 64, 65, 1, -3, 8, 0, 1, 0},   // 15: 16.6 KB
        {2, 8, 96, 112, 56, 72, 1, -3, 8, 0, 1, 0},   // 16: 16.1 KB
        {2, 8, 96, 112, 40, 40, 4, -4, -5, 0, 1, 0},  // 17: aligned x, large box
        {2, 8, 96, 112, 64, 64, 1, 8, 3, 0, 1, 0},    // 18: aligned x
        {2, 8, 96, 112, 32, 8, 1, 3, 0, 0, 1, 0},     // 19: unaligned x, small box
        {2, 8, 96, 112, 44, 40, 4, -8, -5, 4, 1, 0},  // 20: aligned x, large box
    };
    for (auto& cs : cases) {
        ++idx; if (only >= 0 && idx != only) continue;
        size_t N = (size_t)cs.B * cs.C * cs.H * cs.W;
        std::vector<float> h(N);
        for (size_t i = 0; i < N; ++i) h[i] = (float)(i % 100003) + 1.0f;
        float *d, *o;
        cudaMalloc(&d, N * 4); cudaMemcpy(d, h.data(), N * 4, cudaMemcpyHostToDevice);
        int nb = cs.bw * cs.bh * cs.bc;
        cudaMalloc(&o, nb * 4);
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)cs.W, (cuuint64_t)cs.H, (cuuint64_t)cs.C, (cuuint64_t)cs.B};
        cuuint64_t strides[3] = {(cuuint64_t)cs.W * 4, (cuuint64_t)cs.W * cs.H * 4, (cuuint64_t)cs.W * cs.H * cs.C * 4};
        cuuint32_t box[4] = {(cuuint32_t)cs.bw, (cuuint32_t)cs.bh, (cuuint32_t)cs.bc, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("case W=%d H=%d box=%dx%dx%d at (%d,%d,%d,%d) mode %d: encode=%d ", cs.W, cs.H, cs.bw, cs.bh, cs.bc, cs.x, cs.y, cs.c, cs.n, cs.mode, (int)r);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        size_t smem = nb * 4 + 256 + 128;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe<<<1, 128, smem>>>(tm, o, cs.bw, cs.bh, cs.bc, cs.x, cs.y, cs.c, cs.n, cs.mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> got(nb);
        cudaMemcpy(got.data(), o, nb * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int cc = 0; cc < cs.bc; ++cc) for (int yy = 0; yy < cs.bh; ++yy) for (int xx = 0; xx < cs.bw; ++xx) {
            int gx = cs.x + xx, gy = cs.y + yy, gc = cs.c + cc;
            float want = 0.f;
            if (gx >= 0 && gx < cs.W && gy >= 0 && gy < cs.H && gc < cs.C)
                want = h[(((size_t)cs.n * cs.C + gc) * cs.H + gy) * cs.W + gx];
            if (got[(cc * cs.bh + yy) * cs.bw + xx] != want) ++bad;
        }
        printf("mismatches=%d\n", bad);
        cudaFree(d); cudaFree(o);
    }
    return 0;
}
