// Whole-image kernels for the coarse pyramid levels (6x7, 12x14, 6x8, 12x16 ...: H*W <= SMALL_MAX_PX).
//
// At these sizes the path is pure latency (2.5 MB per launch at the 6x7 level),
// so the goal is the shortest dependent chain, not bandwidth: ONE launch for the forward and ONE for
// the whole backward (correlation backward w.r.t. both operands + grid_sample backward, i.e.
// correlation_cuda_kernel.cu:108-290 and SURVEY.md section 8 a10), no workspace, no memset, no
// global atomics.
//
// Decomposition: a thread-block cluster per image; CTA `rank` of the cluster owns the channel slice
// [rank*cs, rank*cs + cs).  The slice of f1 and of the warped f2 lives in shared memory with the
// channel index innermost ([pixel][channel], pitch csp with csp/4 odd), so one 128-bit shared load
// feeds four FMAs and the lanes of a warp (consecutive pixels) never collide.  What has to be summed
// over all channels (the cost volume in the forward, the flow gradient in the backward) is reduced
// across the cluster through distributed shared memory in rank order.
#pragma once
#include <cooperative_groups.h>

#include "pwc_common.cuh"

namespace pwc {

constexpr int SMALL_MAX_PX = 256;
constexpr int SMALL_NT = 768;
constexpr int SMALL_MAX_KS = 8;      // portable cluster size

// Channel slice per CTA (cs, a multiple of 4; pitch csp with csp/4 odd) and cluster size ks (a power
// of two).  The largest ks whose B*ks CTAs still run as ONE wave (one CTA per SM) wins: these levels
// are latency-bound, a second wave costs a whole kernel time.
struct SmallPlan {
    int cs, csp, ks;
    size_t smem;
    bool ok;
};
inline size_t small_smem(int HW, int csp, bool has_flow, bool backward)
{
    const size_t common = (size_t)round_up(81 * HW, 4);
    if (!backward) return sizeof(float) * ((size_t)2 * HW * csp + common + (has_flow ? (size_t)6 * HW : 0));
    return sizeof(float) * ((size_t)(has_flow ? 3 : 2) * HW * csp + common + (has_flow ? (size_t)12 * HW : 0));
}
inline SmallPlan small_plan(int B, int C, int HW, bool has_flow, bool backward, size_t smem_limit, int sm_count)
{
    SmallPlan best = {0, 0, 0, 0, false};
    for (int ks = SMALL_MAX_KS; ks >= 1; ks >>= 1) {
        SmallPlan p;
        p.cs = 4 * cdiv(C, 4 * ks);
        p.csp = ((p.cs / 4) & 1) ? p.cs : p.cs + 4;
        p.ks = 1;
        while (p.ks < cdiv(C, p.cs)) p.ks *= 2;      // few channels: fewer slices than asked for
        p.smem = small_smem(HW, p.csp, has_flow, backward);
        p.ok = p.smem <= smem_limit;
        if (!p.ok) break;                            // fewer slices only need more shared memory
        best = p;
        if ((long long)B * p.ks <= sm_count) break;
    }
    return best;
}

// valid displacement steps t in [-4, 4] with 0 <= v + t*S2 < n
template <int S2>
__device__ __forceinline__ void disp_range(int v, int n, int& lo, int& hi)
{
    lo = max(-4, -(v / S2));
    hi = min(4, (n - 1 - v) / S2);
}

// ---------------------------------------------------------------------------------------------
// forward: out[n, d, p] = act( 1/C * sum_c f1[n,c,p] * W2[n,c,p+d] )
// ---------------------------------------------------------------------------------------------
template <int S2, bool HAS_FLOW>
__global__ void __launch_bounds__(SMALL_NT)
warpcorr_fwd_small_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                          const float* __restrict__ flow, float* __restrict__ out,
                          float* __restrict__ warped_out, int C, int H, int W, int cs, int csp, int act,
                          float slope, long long obs, long long fbs, const float* __restrict__ coarse,
                          float* __restrict__ flow_out, long long fobs)
{
    // flow source: `flow` ([2][H][W] per image, batch stride fbs), or -- model.py:78 folded into this read --
    // `coarse` ([B][2][H/2][W/2], dense): flow = F.upsample(coarse, 2, 'bilinear') * 2, which rank 0 of the
    // cluster also writes to flow_out (batch stride fobs) because the flow estimator needs it as a tensor
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int D = 9, r = 4, NT = SMALL_NT;
    const int ks = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int n = blockIdx.x / ks;
    const int tid = threadIdx.x;
    const int HW = H * W;
    const int c_begin = rank * cs;
    const int nch = max(0, min(cs, C - c_begin));       // a trailing rank may own no channel at all
    const int nq = (nch + 3) >> 2;                      // channel quads in use

    extern __shared__ __align__(16) float smem[];
    float* sF1 = smem;                                   // [HW][csp]
    float* sW2 = sF1 + HW * csp;                         // [HW][csp]
    float* sOut = sW2 + HW * csp;                        // [81][HW] partial cost volume of this slice
    float4* sTapW = reinterpret_cast<float4*>(sOut + round_up(D * D * HW, 4));   // [HW]
    int2* sTapO = reinterpret_cast<int2*>(sTapW + HW);                            // [HW]

    const float* f1n = f1 + ((size_t)n * C + c_begin) * HW;
    const float* f2n = f2 + ((size_t)n * C + c_begin) * HW;

    if (HAS_FLOW) {
        const float* un = coarse ? nullptr : flow + (size_t)n * (size_t)fbs;
        const int Hc = H >> 1, Wc = W >> 1;
        const float* cu = coarse ? coarse + (size_t)n * 2 * Hc * Wc : nullptr;
        for (int q = tid; q < HW; q += NT) {
            const int y = q / W, x = q - y * W;
            float u, v;
            if (coarse) {
                up2_flow_at(cu, cu + Hc * Wc, Hc, Wc, x, y, u, v);
                if (rank == 0 && flow_out != nullptr) {
                    float* fo = flow_out + (size_t)n * (size_t)fobs + q;
                    fo[0] = u;
                    fo[HW] = v;
                }
            } else {
                u = __ldg(un + q);
                v = __ldg(un + HW + q);
            }
            const Tap tp = make_tap(x, y, u, v, H, W);
            sTapW[q] = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
            sTapO[q] = make_int2(tp.off, (tp.dyw << 1) | tp.dx);
        }
    }
    // loads are unconditional (channel index clamped into the slice) so that all of a thread's
    // requests are in flight together; the tail channels of the last quad are zeroed afterwards
    for (int i = tid; i < HW * nq; i += NT) {
        const int q4 = i / HW, p = i - q4 * HW;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(f1n + (size_t)min(4 * q4 + k, nch - 1) * HW + p);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (4 * q4 + k < nch) ? v[k] : 0.0f;
        *reinterpret_cast<float4*>(sF1 + p * csp + 4 * q4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (HAS_FLOW) __syncthreads();

    // ---- warped slice of f2 -> [pixel][channel] ----
    for (int i = tid; i < HW * nq; i += NT) {
        const int q4 = i / HW, q = i - q4 * HW;
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (HAS_FLOW) {
            const float4 w = sTapW[q];
            const int2 o = sTapO[q];
            if (o.x >= 0) {
                const int dx = o.y & 1, dyw = o.y >> 1;
                float c00[4], c01[4], c10[4], c11[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float* p = f2n + (size_t)min(4 * q4 + k, nch - 1) * HW + o.x;
                    c00[k] = __ldg(p); c01[k] = __ldg(p + dx);
                    c10[k] = __ldg(p + dyw); c11[k] = __ldg(p + dyw + dx);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    v[k] = (4 * q4 + k < nch) ? fmaf(w.w, c11[k], fmaf(w.z, c10[k], fmaf(w.y, c01[k], w.x * c00[k]))) : 0.0f;
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __ldg(f2n + (size_t)min(4 * q4 + k, nch - 1) * HW + q);
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (4 * q4 + k < nch) ? v[k] : 0.0f;
        }
        *reinterpret_cast<float4*>(sW2 + q * csp + 4 * q4) = make_float4(v[0], v[1], v[2], v[3]);
        if (HAS_FLOW && warped_out != nullptr) {     // x2_warp export (model.py:107,113)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * q4 + k < nch) warped_out[((size_t)n * C + c_begin + 4 * q4 + k) * HW + q] = v[k];
        }
    }
    __syncthreads();

    // ---- partial cost volume of this slice: item = (displacement row tj, pixel p) ----
    for (int i = tid; i < D * HW; i += NT) {
        const int tji = i / HW, p = i - tji * HW;
        const int y = p / W, x = p - y * W;
        const int y2 = y + (tji - r) * S2;
        float acc[D];
#pragma unroll
        for (int ti = 0; ti < D; ++ti) acc[ti] = 0.0f;
        if (y2 >= 0 && y2 < H) {
            unsigned mask = 0;
#pragma unroll
            for (int ti = 0; ti < D; ++ti) {
                const int x2 = x + (ti - r) * S2;
                if (x2 >= 0 && x2 < W) mask |= 1u << ti;
            }
            const float* pf = sF1 + p * csp;
            const float* pw = sW2 + (y2 * W + x) * csp;
            for (int q4 = 0; q4 < nq; ++q4) {
                const float4 f = *reinterpret_cast<const float4*>(pf + 4 * q4);
#pragma unroll
                for (int ti = 0; ti < D; ++ti) {
                    if ((mask >> ti) & 1u) {
                        const float4 w = *reinterpret_cast<const float4*>(pw + (ti - r) * S2 * csp + 4 * q4);
                        acc[ti] = fmaf(f.x, w.x, fmaf(f.y, w.y, fmaf(f.z, w.z, fmaf(f.w, w.w, acc[ti]))));
                    }
                }
            }
        }
#pragma unroll
        for (int ti = 0; ti < D; ++ti) sOut[(tji * D + ti) * HW + p] = acc[ti];
    }
    cluster.sync();

    // ---- reduce-scatter over the cluster: rank k finishes outputs [k*per, (k+1)*per) ----
    {
        const int NO = D * D * HW;
        const int per = cdiv(NO, ks);
        const int lo = rank * per, hi = min(NO, lo + per);
        const float* remote[SMALL_MAX_KS];
#pragma unroll
        for (int rr = 0; rr < SMALL_MAX_KS; ++rr) remote[rr] = cluster.map_shared_rank(sOut, rr < ks ? rr : 0);
        const float nelems = (float)C;   // correlation_cuda_kernel.cu:65,100
        float* on = out + (size_t)n * (size_t)obs;
        for (int o = lo + tid; o < hi; o += NT) {
            float part[SMALL_MAX_KS];
#pragma unroll
            for (int rr = 0; rr < SMALL_MAX_KS; ++rr) part[rr] = (rr < ks) ? remote[rr][o] : 0.0f;
            float s = 0.0f;
#pragma unroll
            for (int rr = 0; rr < SMALL_MAX_KS; ++rr) s += part[rr];
            s = s / nelems;
            if (act) s = leaky(s, slope);
            on[o] = s;
        }
    }
    cluster.sync();      // remote shared memory stays valid until every rank has read it
}

// ---------------------------------------------------------------------------------------------
// backward, everything in one launch:
//   gO'      = grad_out * (gate < 0 ? slope : 1)                         (LeakyReLU, model.py:84)
//   gW2[c,q] = 1/C * sum_d gO'[d, q-d] * f1[c, q-d]                      (correlation_cuda_kernel.cu:200-290)
//   g1[c,p]  = 1/C * sum_d gO'[d, p]   * W2[c, p+d]                      (correlation_cuda_kernel.cu:108-198)
//   gf2      = bilinear scatter of gW2, gflow = sum_c gW2 * dW2/d(u,v)   (grid_sample backward)
// Without flow W2 = f2 and gf2 = gW2 (the legacy Correlation backward).
// Only displacements that stay inside the image are visited (a quarter of the 81 at the 6x7 level).
// ---------------------------------------------------------------------------------------------
template <int S2, bool HAS_FLOW>
__global__ void __launch_bounds__(SMALL_NT)
warpcorr_bwd_small_kernel(const float* __restrict__ gout, const float* __restrict__ gate,
                          const float* __restrict__ f1, const float* __restrict__ f2,
                          const float* __restrict__ flow, float* __restrict__ gf1, float* __restrict__ gf2,
                          float* __restrict__ gflow, int C, int H, int W, int cs, int csp, float slope,
                          long long gbs, long long gate_bs)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int D = 9, r = 4, NT = SMALL_NT;
    const int ks = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int n = blockIdx.x / ks;
    const int tid = threadIdx.x;
    const int HW = H * W;
    const int c_begin = rank * cs;
    const int nch = max(0, min(cs, C - c_begin));
    const int nq = (nch + 3) >> 2;

    extern __shared__ __align__(16) float smem[];
    float* sF1 = smem;                                   // [HW][csp]
    float* sW2 = sF1 + HW * csp;                         // [HW][csp]
    float* sG = sW2 + HW * csp;                          // [81][HW] gated output gradient
    float* sGF2 = sG + round_up(D * D * HW, 4);          // [HW][csp] scatter target     (HAS_FLOW only from here)
    float4* sTapW = reinterpret_cast<float4*>(sGF2 + HW * csp);   // [HW] corner weights
    int4* sTapO = reinterpret_cast<int4*>(sTapW + HW);            // [HW] off, dx, dyw, corner mask
    float2* sTapA = reinterpret_cast<float2*>(sTapO + HW);        // [HW] fractions ax, ay
    float* sGFl = reinterpret_cast<float*>(sTapA + HW);           // [2][HW] flow gradient of this slice

    const float* f1n = f1 + ((size_t)n * C + c_begin) * HW;
    const float* f2n = f2 + ((size_t)n * C + c_begin) * HW;
    const float nelems = (float)C;

    // ---- stage: gated output gradient, f1 slice, taps; zero the accumulators ----
    {
        const float* gon = gout + (size_t)n * (size_t)gbs;       // batch strides of the output gradient / the gate
        const float* gaten = gate ? gate + (size_t)n * (size_t)gate_bs : nullptr;
        for (int i = tid; i < D * D * HW; i += NT) {
            float g = __ldg(gon + i);
            if (gaten && __ldg(gaten + i) < 0.0f) g *= slope;
            sG[i] = g;
        }
    }
    for (int i = tid; i < HW * nq; i += NT) {
        const int q4 = i / HW, p = i - q4 * HW;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(f1n + (size_t)min(4 * q4 + k, nch - 1) * HW + p);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (4 * q4 + k < nch) ? v[k] : 0.0f;
        *reinterpret_cast<float4*>(sF1 + p * csp + 4 * q4) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (HAS_FLOW) {
        const float* un = flow + (size_t)n * 2 * HW;
        for (int q = tid; q < HW; q += NT) {
            const int y = q / W, x = q - y * W;
            float ax = 0.0f, ay = 0.0f;
            int x0 = 0, y0 = 0;
            const Tap tp = make_tap(x, y, __ldg(un + q), __ldg(un + HW + q), H, W, &ax, &ay, &x0, &y0);
            int mask = 0;
            if (tp.off >= 0) {
                const bool inx0 = x0 >= 0, inx1 = x0 + 1 < W, iny0 = y0 >= 0, iny1 = y0 + 1 < H;
                mask = (inx0 && iny0 ? 1 : 0) | (inx1 && iny0 ? 2 : 0) | (inx0 && iny1 ? 4 : 0) | (inx1 && iny1 ? 8 : 0);
            }
            sTapW[q] = make_float4(tp.w00, tp.w01, tp.w10, tp.w11);
            sTapO[q] = make_int4(tp.off, tp.dx, tp.dyw, mask);
            sTapA[q] = make_float2(ax, ay);
            sGFl[q] = 0.0f;
            sGFl[HW + q] = 0.0f;
        }
        for (int i = tid; i < HW * csp; i += NT) sGF2[i] = 0.0f;
    }
    __syncthreads();

    // ---- pass 1, item = (pixel q, channel quad): W2, gW2, scatter, flow gradient ----
    for (int i = tid; i < HW * nq; i += NT) {
        const int q4 = i / HW, q = i - q4 * HW;
        const int y = q / W, x = q - y * W;

        // issue the global gathers first: they land while gW2 is computed out of shared memory
        float c00[4], c01[4], c10[4], c11[4];
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 o = make_int4(-1, 0, 0, 0);
        float2 a = make_float2(0.f, 0.f);
        if (HAS_FLOW) {
            w = sTapW[q];
            o = sTapO[q];
            a = sTapA[q];
            const int off = max(o.x, 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float* p = f2n + (size_t)min(4 * q4 + k, nch - 1) * HW + off;
                c00[k] = __ldg(p); c01[k] = __ldg(p + o.y);
                c10[k] = __ldg(p + o.z); c11[k] = __ldg(p + o.z + o.y);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) c00[k] = __ldg(f2n + (size_t)min(4 * q4 + k, nch - 1) * HW + q);
        }

        // gradient w.r.t. the warped features at q: source pixels p = q - d inside the image
        float gw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        {
            int tlo, thi, slo, shi;
            disp_range<S2>(H - 1 - y, H, tlo, thi);      // 0 <= y - t*S2 < H  <=>  range of (H-1-y) + t*S2
            disp_range<S2>(W - 1 - x, W, slo, shi);
            for (int t = tlo; t <= thi; ++t) {
                const int prow = (y - t * S2) * W + x;
                const float* gr = sG + ((t + r) * D + r) * HW + prow;
                for (int s = slo; s <= shi; ++s) {
                    const float g = gr[s * (HW - S2)];                       // sG[d][prow - s*S2]
                    const float4 f = *reinterpret_cast<const float4*>(sF1 + (prow - s * S2) * csp + 4 * q4);
                    gw[0] = fmaf(g, f.x, gw[0]); gw[1] = fmaf(g, f.y, gw[1]);
                    gw[2] = fmaf(g, f.z, gw[2]); gw[3] = fmaf(g, f.w, gw[3]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) gw[k] = (4 * q4 + k < nch) ? gw[k] / nelems : 0.0f;

        if (!HAS_FLOW) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * q4 + k < nch) gf2[((size_t)n * C + c_begin + 4 * q4 + k) * HW + q] = gw[k];
            *reinterpret_cast<float4*>(sW2 + q * csp + 4 * q4) = make_float4(c00[0], c00[1], c00[2], c00[3]);
        } else {
            float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (o.x >= 0) {
                const float m00 = (o.w & 1) ? 1.0f : 0.0f, m01 = (o.w & 2) ? 1.0f : 0.0f;
                const float m10 = (o.w & 4) ? 1.0f : 0.0f, m11 = (o.w & 8) ? 1.0f : 0.0f;
                float gu = 0.0f, gv = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float v00 = m00 * c00[k], v01 = m01 * c01[k], v10 = m10 * c10[k], v11 = m11 * c11[k];
                    v[k] = fmaf(w.w, v11, fmaf(w.z, v10, fmaf(w.y, v01, w.x * v00)));
                    gu = fmaf(gw[k], fmaf(v11 - v10, a.y, (v01 - v00) * (1.0f - a.y)), gu);
                    gv = fmaf(gw[k], fmaf(v11 - v01, a.x, (v10 - v00) * (1.0f - a.x)), gv);
                }
                float* t = sGF2 + o.x * csp + 4 * q4;
                const int sdx = o.y * csp, sdy = o.z * csp;   // o.y is 0 or 1 pixel, o.z is 0 or W pixels
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (w.x != 0.0f) atomicAdd(t + k, gw[k] * w.x);
                    if (w.y != 0.0f) atomicAdd(t + sdx + k, gw[k] * w.y);
                    if (w.z != 0.0f) atomicAdd(t + sdy + k, gw[k] * w.z);
                    if (w.w != 0.0f) atomicAdd(t + sdy + sdx + k, gw[k] * w.w);
                }
                atomicAdd(&sGFl[q], gu);
                atomicAdd(&sGFl[HW + q], gv);
            }
            *reinterpret_cast<float4*>(sW2 + q * csp + 4 * q4) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();

    // ---- pass 2, item = (pixel p, channel quad): g1 from W2; write the scattered gf2 slice ----
    for (int i = tid; i < HW * nq; i += NT) {
        const int q4 = i / HW, p = i - q4 * HW;
        const int y = p / W, x = p - y * W;
        float g1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        int tlo, thi, slo, shi;
        disp_range<S2>(y, H, tlo, thi);
        disp_range<S2>(x, W, slo, shi);
        for (int t = tlo; t <= thi; ++t) {
            const float* gr = sG + ((t + r) * D + r) * HW + p;
            const float* wr = sW2 + ((y + t * S2) * W + x) * csp + 4 * q4;
            for (int s = slo; s <= shi; ++s) {
                const float g = gr[s * HW];
                const float4 wv = *reinterpret_cast<const float4*>(wr + s * S2 * csp);
                g1[0] = fmaf(g, wv.x, g1[0]); g1[1] = fmaf(g, wv.y, g1[1]);
                g1[2] = fmaf(g, wv.z, g1[2]); g1[3] = fmaf(g, wv.w, g1[3]);
            }
        }
        const float4 sc = HAS_FLOW ? *reinterpret_cast<const float4*>(sGF2 + p * csp + 4 * q4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float sv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (4 * q4 + k < nch) {
                const size_t o = ((size_t)n * C + c_begin + 4 * q4 + k) * HW + p;
                gf1[o] = g1[k] / nelems;
                if (HAS_FLOW) gf2[o] = sv[k];
            }
        }
    }

    if (HAS_FLOW) {
        // flow gradient: sum of the slices' partial sums, in rank order
        cluster.sync();
        if (rank == 0) {
            const float* remote[SMALL_MAX_KS];
#pragma unroll
            for (int rr = 0; rr < SMALL_MAX_KS; ++rr) remote[rr] = cluster.map_shared_rank(sGFl, rr < ks ? rr : 0);
            for (int i = tid; i < 2 * HW; i += NT) {
                float s = 0.0f;
#pragma unroll
                for (int rr = 0; rr < SMALL_MAX_KS; ++rr)
                    if (rr < ks) s += remote[rr][i];
                gflow[(size_t)n * 2 * HW + i] = s;
            }
        }
        cluster.sync();
    }
}

}  // namespace pwc
