// tcgen05 decision probe (not part of the product): the correlate stage of the cost volume as a banded GEMM on
// the 5th-generation tensor cores, so that the question "do tensor cores win on the C = 96 / 128 / 196 levels?"
// (north_star) is answered by a measurement instead of an estimate.
//
//   out[p, d] = 1/C * sum_c f1[p, c] * f2[p + d, c],   d in [-4, 4]^2          (no warp: the plain Correlation)
//
// One CTA (16 warps) per 8 x 16 pixel tile: M = 128 pixels (TMEM lanes), N = 16 x 24 = 384 halo pixels (TMEM columns),
// K = channels in chunks of 8.  fp32 parity (1e-5) needs the 3xTF32 split: a = a_hi + a_lo with a_hi = tf32(a),
// D += A_hi B_hi + A_hi B_lo + A_lo B_hi (the dropped lo*lo term is 2^-22 relative).  Of the 128 x 384 products
// of a tile, 128 x 81 are outputs (21 %): the band is extracted in the epilogue (tcgen05.ld, predicated stores).
// Operands are written to shared memory by the threads in the K-major, no-swizzle canonical layout
// (core matrix = 8 rows x 16 bytes), double buffered; one elected thread issues tcgen05.mma kind::tf32.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/tc_probe scripts/tc_probe.cu
//   build/tc_probe [B C H W] [mode]     mode 0: full (default), 1: no epilogue stores, 2: swap LBO/SBO (layout check)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../pwc_net_pytorch_b200/csrc/warpcorr_fwd.cuh"

namespace tc {

constexpr int TH = 8, TW = 16, R = 4, HH = TH + 2 * R, HWD = TW + 2 * R;     // 8x16 tile, 16x24 halo
constexpr int M = TH * TW, N = HH * HWD, NH = N / 2, KC = 8;                  // 128, 384, two MMAs of N = 192, K chunk
constexpr int A_ELEMS = M * KC, B_ELEMS = N * KC;                             // per precision part
constexpr int BUF_FLOATS = 2 * (A_ELEMS + B_ELEMS);                           // hi + lo
constexpr size_t SMEM = 1024 + 2 * BUF_FLOATS * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: 8-row x 16-byte core matrices; `lbo` = bytes between core
// matrices along K, `sbo` = bytes between 8-row groups along M/N
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;       // descriptor version (sm_100)
    return d;                     // layout type 0 = no swizzle, base offset 0
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// a = hi + lo, hi = a with the 13 low mantissa bits cleared (what kind::tf32 reads of a 32-bit operand)
__device__ __forceinline__ void split(float a, float& hi, float& lo)
{
    hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
    lo = a - hi;
}

// mode bit 0: skip the epilogue's stores; bit 1: swap LBO / SBO (layout check)
constexpr int NT = 512;      // 16 warps: operand preparation over all of them, epilogue columns split over the four warp groups

__global__ void __launch_bounds__(NT, 1)
corr_tc_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, int C, int H, int W,
               int tiles_x, int tiles_y, int mode)
{
    extern __shared__ __align__(1024) uint8_t base[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(base);               // [2] MMAs of a buffer have completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 64);
    float* buf = reinterpret_cast<float*>(base + 1024);              // [2][A_hi | A_lo | B_hi | B_lo]
    const int tid = threadIdx.x, warp = tid >> 5;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int n = t / tiles_y;
    const int y0 = ty * TH, x0 = tx * TW;
    const size_t HW = (size_t)H * W;

    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 192, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int nchunks = C / KC;
    // operand preparation: work item = (row of A or B, k-group of 4 channels): (128 + 384) * 2 = 1024 items = 2 per thread
    for (int k = 0; k < nchunks; ++k) {
        const int b = k & 1;
        if (k >= 2) mbar_wait(&bar[b], ((k >> 1) - 1) & 1);          // MMAs that read buffer b (chunk k-2) are done
        float* Ahi = buf + b * BUF_FLOATS;
        float* Alo = Ahi + A_ELEMS;
        float* Bhi = Alo + A_ELEMS;
        float* Blo = Bhi + B_ELEMS;
        float v[2][4];
        int dst[2];
        bool isA[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int item = tid + NT * j;                 // 0 .. 1023
            const int g = item & 1, row = item >> 1;       // row 0..127: A (pixel), 128..511: B (halo pixel)
            isA[j] = row < M;
            int y, x;
            const float* src;
            if (isA[j]) {
                y = y0 + (row >> 4); x = x0 + (row & 15); src = f1;
                dst[j] = (g * M + row) * 4;
            } else {
                const int q = row - M, hy = q / HWD, hx = q - hy * HWD;
                y = y0 - R + hy; x = x0 - R + hx; src = f2;
                dst[j] = (g * N + q) * 4;
            }
            const bool in = y >= 0 && y < H && x >= 0 && x < W;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                v[j][c] = in ? __ldg(src + ((size_t)n * C + k * KC + 4 * g + c) * HW + (size_t)y * W + x) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float4 h, l;
            split(v[j][0], h.x, l.x); split(v[j][1], h.y, l.y); split(v[j][2], h.z, l.z); split(v[j][3], h.w, l.w);
            *reinterpret_cast<float4*>((isA[j] ? Ahi : Bhi) + dst[j]) = h;
            *reinterpret_cast<float4*>((isA[j] ? Alo : Blo) + dst[j]) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> async proxy (tcgen05.mma)
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            // K-major: LBO = distance between the two k-groups of an MMA's K = 8, SBO = distance between 8-row groups
            uint32_t lboA = M * 16, lboB = N * 16, sbo = 128;
            if (mode & 2) { const uint32_t s = sbo; lboA = s; lboB = s; sbo = M * 16; }
            const uint64_t ah = make_desc(smem_u32(Ahi), lboA, sbo), al = make_desc(smem_u32(Alo), lboA, sbo);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint64_t bh = make_desc(smem_u32(Bhi) + h * NH * 16, lboB, sbo);
                const uint64_t bl = make_desc(smem_u32(Blo) + h * NH * 16, lboB, sbo);
                const uint32_t d = tmem + h * NH;
                mma_tf32(d, ah, bh, idesc, k > 0);
                mma_tf32(d, ah, bl, idesc, 1);
                mma_tf32(d, al, bh, idesc, 1);
            }
            mma_commit(&bar[b]);       // arrives when every MMA issued so far has completed
        }
    }
    // ---- epilogue: wait for the last chunk's MMAs, extract the band ----
    {
        const int k = nchunks - 1;
        mbar_wait(&bar[k & 1], (k >> 1) & 1);
        if (nchunks >= 2) mbar_wait(&bar[(k - 1) & 1], ((k - 1) >> 1) & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    const float inv = 1.0f / (float)C;
    // warp w reads TMEM lanes 32*(w % 4) ..: pixel of this thread; the four warp groups split the 384 columns
    const int lane_row = (warp & 3) * 32 + (tid & 31);
    const int pr = lane_row >> 4, pc = lane_row & 15;
    const int y = y0 + pr, x = x0 + pc;
    const int wg = warp >> 2;
    float keep = 0.0f;
#pragma unroll 1
    for (int cb = wg * (N / 64); cb < (wg + 1) * (N / 64); ++cb) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + cb * 16;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int col = cb * 16 + j, hy = col / HWD, hx = col - hy * HWD;
            const int tj = hy - pr, ti = hx - pc;
            if (tj >= 0 && tj < 9 && ti >= 0 && ti < 9 && y < H && x < W) {
                const float val = __uint_as_float(r[j]) * inv;
                if (mode & 1) keep += val;
                else out[((size_t)n * 81 + tj * 9 + ti) * HW + (size_t)y * W + x] = val;
            }
        }
    }
    if ((mode & 1) && keep == 123.456f) out[0] = keep;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace tc

static unsigned long long g_rng = 0x9E3779B97F4A7C15ull;
static float urand() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (float)((g_rng >> 11) * (1.0 / 9007199254740992.0)); }
static float nrand() { const float a = fmaxf(urand(), 1e-12f), b = urand(); return sqrtf(-2.0f * logf(a)) * cosf(6.2831853f * b); }
#define CK_(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv)
{
    int B = 32, C = 96, H = 24, W = 28, mode = 0, ai = 1;
    if (argc > 4) { B = atoi(argv[1]); C = atoi(argv[2]); H = atoi(argv[3]); W = atoi(argv[4]); ai = 5; }
    if (argc > ai) mode = atoi(argv[ai]);
    if (C % 8) { printf("C must be a multiple of 8\n"); return 1; }
    const size_t HW = (size_t)H * W, N = (size_t)B * C * HW, NO = (size_t)B * 81 * HW;
    std::vector<float> h1(N), h2(N);
    for (auto& v : h1) v = nrand();
    for (auto& v : h2) v = nrand();
    float *d1, *d2, *dout, *dref;
    CK_(cudaMalloc(&d1, N * 4)); CK_(cudaMalloc(&d2, N * 4)); CK_(cudaMalloc(&dout, NO * 4)); CK_(cudaMalloc(&dref, NO * 4));
    CK_(cudaMemcpy(d1, h1.data(), N * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemcpy(d2, h2.data(), N * 4, cudaMemcpyHostToDevice));
    CK_(cudaMemset(dout, 0, NO * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_ref = 0, ms_tc = 0;
    {   // SIMT reference: the plain tiled kernel of the library, no flow
        using RC = pwc::FwdCfg<9, 1, 8, 4, 8>;
        auto kern = pwc::warpcorr_fwd_kernel<RC, false>;
        const size_t smem = RC::smem_bytes(false);
        CK_(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int tx = pwc::cdiv(W, RC::TW), ty = pwc::cdiv(H, RC::TH);
        for (int it = 0; it < 6; ++it) {
            if (it == 1) cudaEventRecord(e0);
            kern<<<tx * ty * B, RC::NT, smem>>>(d1, d2, nullptr, dref, nullptr, C, H, W, tx, ty, 0, 0.0f, 1, C, (long long)81 * HW,
                                                (long long)2 * HW);
        }
        cudaEventRecord(e1);
        CK_(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms_ref, e0, e1); ms_ref /= 5;
    }
    CK_(cudaFuncSetAttribute(tc::corr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM));
    const int tiles_x = pwc::cdiv(W, tc::TW), tiles_y = pwc::cdiv(H, tc::TH);
    for (int it = 0; it < 6; ++it) {
        if (it == 1) cudaEventRecord(e0);
        tc::corr_tc_kernel<<<tiles_x * tiles_y * B, tc::NT, tc::SMEM>>>(d1, d2, dout, C, H, W, tiles_x, tiles_y, mode);
    }
    cudaEventRecord(e1);
    CK_(cudaGetLastError());
    CK_(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms_tc, e0, e1); ms_tc /= 5;
    std::vector<float> ho(NO), hr(NO);
    CK_(cudaMemcpy(ho.data(), dout, NO * 4, cudaMemcpyDeviceToHost));
    CK_(cudaMemcpy(hr.data(), dref, NO * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < NO; ++i) { maxerr = fmax(maxerr, fabs((double)ho[i] - hr[i])); maxref = fmax(maxref, fabs((double)hr[i])); }
    const double useful = 2.0 * 81 * C * B * HW, issued = 2.0 * 3 * tc::M * tc::N * C * (double)(tiles_x * tiles_y * B);
    printf("B=%d C=%d %dx%d mode=%d: tcgen05 3xTF32 banded GEMM %.1f us (%.1f TFLOP/s issued tf32, %.1f TFLOP/s useful) | SIMT tiled "
           "%.1f us | max rel err %.2e %s\n", B, C, H, W, mode, 1e3 * ms_tc, issued / ms_tc / 1e9, useful / ms_tc / 1e9, 1e3 * ms_ref,
           maxerr / maxref, (mode & 1) ? "(stores skipped)" : (maxerr / maxref < 1e-5 ? "OK" : "MISMATCH"));
    return 0;
}
