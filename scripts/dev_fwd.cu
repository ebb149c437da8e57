// Development harness for the fused forward kernel (not part of the product): builds in seconds, runs the
// TMA kernel at one shape, times it with CUDA events and checks it against the plain tiled kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o /tmp/dev_fwd scripts/dev_fwd.cu
//   /tmp/dev_fwd [B C H W] [iid|smooth] [iters]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../pwc_net_pytorch_b200/csrc/warpcorr_fwd.cuh"
#include "../pwc_net_pytorch_b200/csrc/warpcorr_fwd_tma.cuh"

#ifndef DEV_S2
#define DEV_S2 1
#endif
#ifndef DEV_CK
#define DEV_CK 4
#endif

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_enc;

static bool make_map(CUtensorMap* map, const float* ptr, int B, int C, int H, int W, int bw, int bh, int bc)
{
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return g_enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static unsigned long long g_rng = 0x9E3779B97F4A7C15ull;
static float urand()
{
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
    return (float)((g_rng >> 11) * (1.0 / 9007199254740992.0));
}
static float nrand()
{
    const float a = fmaxf(urand(), 1e-12f), b = urand();
    return sqrtf(-2.0f * logf(a)) * cosf(6.2831853f * b);
}

#define CK_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv)
{
    int B = 32, C = 32, H = 96, W = 112, iters = 20;
    bool smooth = false, noflow = false;
    int ai = 1;
    if (argc > 4 && atoi(argv[1]) > 0) { B = atoi(argv[1]); C = atoi(argv[2]); H = atoi(argv[3]); W = atoi(argv[4]); ai = 5; }
    if (argc > ai) smooth = strcmp(argv[ai], "smooth") == 0;
    if (argc > ai) noflow = strcmp(argv[ai], "noflow") == 0;     // plain correlation: zero flow for the check
    if (argc > ai + 1) iters = atoi(argv[ai + 1]);
    cudaFree(0);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    g_enc = (EncodeTiledFn)p;
    const size_t HW = (size_t)H * W, N = (size_t)B * C * HW, NF = (size_t)B * 2 * HW, NO = (size_t)B * 81 * HW;
    std::vector<float> h1(N), h2(N), hf(NF);
    for (auto& v : h1) v = nrand();
    for (auto& v : h2) v = nrand();
    if (noflow) {
        for (auto& v : hf) v = 0.0f;
    } else if (!smooth) {
        for (auto& v : hf) v = 2.0f * nrand();
    } else {
        const int ch = H / 8 + 2, cw = W / 8 + 2;
        std::vector<float> cz((size_t)B * 2 * ch * cw);
        for (auto& v : cz) v = 2.0f * nrand();
        for (int n = 0; n < B * 2; ++n)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x) {
                    const float fy = y / 8.0f, fx = x / 8.0f;
                    const int y0 = (int)fy, x0 = (int)fx;
                    const float ay = fy - y0, ax = fx - x0;
                    const float* c0 = &cz[(size_t)n * ch * cw];
                    hf[(size_t)n * HW + (size_t)y * W + x] =
                        (1 - ay) * ((1 - ax) * c0[y0 * cw + x0] + ax * c0[y0 * cw + x0 + 1]) +
                        ay * ((1 - ax) * c0[(y0 + 1) * cw + x0] + ax * c0[(y0 + 1) * cw + x0 + 1]);
                }
    }
    float *d1, *d2, *df, *dout, *dref;
    CK_(cudaMalloc(&d1, N * 4)); CK_(cudaMalloc(&d2, N * 4)); CK_(cudaMalloc(&df, NF * 4));
    CK_(cudaMalloc(&dout, NO * 4)); CK_(cudaMalloc(&dref, NO * 4));
    CK_(cudaMemcpy(d1, h1.data(), N * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(d2, h2.data(), N * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(df, hf.data(), NF * 4, cudaMemcpyHostToDevice));
    // L2 flush buffer
    float* dflush; const size_t FL = 256u << 20;
    CK_(cudaMalloc(&dflush, FL));

    // ---- reference: plain tiled kernel ----
    {
        using RC = pwc::FwdCfg<9, DEV_S2, 8, 4, 8>;
        auto kern = pwc::warpcorr_fwd_kernel<RC, true>;
        const size_t smem = RC::smem_bytes(true);
        CK_(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tx = pwc::cdiv(W, RC::TW), ty = pwc::cdiv(H, RC::TH);
        kern<<<tx * ty * B, RC::NT, smem>>>(d1, d2, df, dref, nullptr, C, H, W, tx, ty, 0, 0.0f, 1, C, (long long)81 * HW,
                                            (long long)2 * HW);
        CK_(cudaGetLastError());
        CK_(cudaDeviceSynchronize());
    }
    // ---- kernel under test ----
    using Cfg = pwc::TmaCfg<DEV_S2, DEV_CK>;
    CUtensorMap m1, m2, m3;
    if (!make_map(&m1, d1, B, C, H, W, Cfg::F1W, Cfg::F1H, Cfg::CK) || !make_map(&m2, d2, B, C, H, W, Cfg::WW, Cfg::WH, Cfg::CK) ||
        !make_map(&m3, df, B, 2, H, W, Cfg::HWD, Cfg::HH, 2)) { printf("tensor map failed\n"); return 1; }
    auto kern = noflow ? pwc::warpcorr_fwd_tma_kernel<Cfg, false> : pwc::warpcorr_fwd_tma_kernel<Cfg, true>;
    const size_t smem = Cfg::smem_bytes(!noflow);
    if (noflow && !make_map(&m2, d2, B, C, H, W, Cfg::WP, Cfg::HH, Cfg::CK)) { printf("tensor map failed\n"); return 1; }
    CK_(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = pwc::cdiv(W, Cfg::TW), tiles_y = pwc::cdiv(H, Cfg::TH);
    const int ntiles = tiles_x * tiles_y * B;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const unsigned grid = ntiles < sms ? ntiles : sms;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, sum = 0.0f;
    for (int it = 0; it < iters + 3; ++it) {
        if (!getenv("DEV_NOFLUSH")) CK_(cudaMemsetAsync(dflush, it & 255, FL));     // evict the inputs from L2 between iterations
        cudaEventRecord(e0);
        kern<<<grid, Cfg::NT, smem>>>(m1, m2, m3, d2, noflow ? nullptr : df, dout, nullptr, C, H, W, tiles_x, tiles_y, ntiles, 0, 0.0f,
                                      (long long)81 * HW, (long long)2 * HW, nullptr, nullptr, 0);
        cudaEventRecord(e1);
        CK_(cudaGetLastError());
        CK_(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 3) { best = fminf(best, ms); sum += ms; }
    }
    std::vector<float> ho(NO), hr(NO);
    CK_(cudaMemcpy(ho.data(), dout, NO * 4, cudaMemcpyDeviceToHost));
    CK_(cudaMemcpy(hr.data(), dref, NO * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < NO; ++i) { maxerr = fmax(maxerr, fabs((double)ho[i] - hr[i])); maxref = fmax(maxref, fabs((double)hr[i])); }
    const double bytes = 4.0 * (2 * C + 2 + 81) * B * HW;
    printf("shape %dx%dx%dx%d %s S2=%d CK=%d smem=%zu: mean %.2f us best %.2f us  %.1f GB/s (mean)  frac(6453)=%.3f  max rel err vs tiled %.2e %s\n",
           B, C, H, W, noflow ? "noflow" : smooth ? "smooth" : "iid", DEV_S2, DEV_CK, smem, 1e3 * sum / iters, 1e3 * best, bytes / (sum / iters) / 1e6,
           bytes / (sum / iters) / 1e6 / 6453.1, maxerr / maxref, maxerr / maxref < 2e-6 ? "OK" : "MISMATCH");
    return maxerr / maxref < 2e-6 ? 0 : 2;
}
