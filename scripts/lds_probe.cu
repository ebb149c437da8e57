// Shared-memory load bandwidth probe (development only): bytes per clock and SM for LDS.32 / .64 / .128,
// conflict-free addresses, with and without all lanes reading the same address (broadcast).
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long g_clk[2];
template <int VEC, bool BCAST, int SHARE = 1, int PAT = 0>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters)
{
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const unsigned long long c0 = clock64();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // SHARE lanes read the same address; PAT 0: sharing lanes are adjacent (lane / SHARE), PAT 1: sharing lanes are
    // SHARE groups apart (lane % (32 / SHARE)), i.e. spread over the quarter / half warps
    // PAT >= 2: lanes l and l ^ (1 << (PAT - 2)) share an address, 16 distinct addresses per warp
    int slot = PAT == 0 ? lane / SHARE : lane % (32 / SHARE);
    if (PAT >= 2) { const int bit = 1 << (PAT - 2); slot = ((lane >> (PAT - 1)) << (PAT - 2)) | (lane & (bit - 1)); }
    const float* p = sm + (BCAST ? 0 : slot * VEC) + (warp & 7) * 128;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned a = (unsigned)__cvta_generic_to_shared(p + u * 1024);
            if (VEC == 1) { float x; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); acc[u] += x; }
            else if (VEC == 2) { float x, y; asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); acc[u] += x + y; }
            else { float x, y, z, w; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); acc[u] += (x + y) + (z + w); }
        }
    }
    float s = 0;
    for (int u = 0; u < 8; ++u) s += acc[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) g_clk[0] = clock64() - c0;
}
template <int VEC, bool BCAST, int SHARE = 1, int PAT = 0>
static void run(const char* name, int nthr, float* dout)
{
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int t = 0; t < 3; ++t) { cudaEventRecord(e0); k<VEC, BCAST, SHARE, PAT><<<148, nthr>>>(dout, iters); cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1); }
    unsigned long long clk[2];
    cudaMemcpyFromSymbol(clk, g_clk, sizeof(clk));
    const double bytes = 8.0 * iters * nthr * VEC * 4;
    printf("%-22s %4d thr: %.1f B/clk/SM, %.2f clk per warp instruction (clock64 %llu cycles, event %.1f us)  %s\n", name, nthr, bytes / clk[0],
           (double)clk[0] / (8.0 * iters * nthr / 32), clk[0], ms * 1e3, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float* dout; cudaMalloc(&dout, 1 << 22);
    for (int nthr : {512}) {
        run<4, false>("LDS.128 distinct", nthr, dout);
        run<4, false, 2, 2>("LDS.128 share l^1", nthr, dout);
        run<4, false, 2, 3>("LDS.128 share l^2", nthr, dout);
        run<4, false, 2, 4>("LDS.128 share l^4", nthr, dout);
        run<4, false, 2, 5>("LDS.128 share l^8", nthr, dout);
        run<4, false, 2, 6>("LDS.128 share l^16", nthr, dout);
        run<2, false, 2, 2>("LDS.64 share l^1", nthr, dout);
        run<2, false, 2, 3>("LDS.64 share l^2", nthr, dout);
        run<2, false, 2, 4>("LDS.64 share l^4", nthr, dout);
        run<2, false, 2, 5>("LDS.64 share l^8", nthr, dout);
        run<2, false, 2, 6>("LDS.64 share l^16", nthr, dout);
    }
    return 0;
}
