// extern "C" surface of libpwc_b200.so (declared in include/pwc_b200.h) and kernel dispatch.
// Plain pointers and sizes only; no torch types.  Never allocates, frees or synchronises.
#include "../../include/pwc_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "corr_bwd.cuh"
#include "corr_bwd_seq.cuh"
#include "corr_bwd_tma.cuh"
#include "generic_kernels.cuh"
#include "pwc_common.cuh"
#include "small_image.cuh"
#include "warp_bwd_tile.cuh"
#include "warpcorr_fwd.cuh"
#include "warpcorr_fwd_tma.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_force_generic{0};
std::atomic<int> g_disable_tma{0};
std::atomic<int> g_disable_small{0};
std::atomic<int> g_disable_seq{0};

int fail(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 0;
}

int check_launch(const char* what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
    return 1;
}

bool make_geom(pwc::CorrGeom& g, int B, int C, int H, int W, int pad, int k, int md, int s1, int s2)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("non-positive tensor size") != 0;
    if (pad < 0 || k < 1 || (k & 1) == 0 || md < 0 || s1 < 1 || s2 < 1)
        return fail("bad correlation parameters (pad=%d k=%d md=%d s1=%d s2=%d); kernel_size must "
                    "be odd and >= 1, strides >= 1", pad, k, md, s1, s2) != 0;
    g.B = B; g.C = C; g.H = H; g.W = W;
    g.pad = pad; g.k = k; g.md = md; g.s1 = s1; g.s2 = s2;
    g.kr = (k - 1) / 2;
    g.r = md / s2;
    g.D = 2 * g.r + 1;
    g.oc = g.D * g.D;
    // ceil((padded - 2*border) / stride1), correlation_cuda.c:25-34
    const int nh = H + 2 * pad - 2 * (g.kr + md), nw = W + 2 * pad - 2 * (g.kr + md);
    g.oh = nh > 0 ? (nh + s1 - 1) / s1 : 0;
    g.ow = nw > 0 ? (nw + s1 - 1) / s1 : 0;
    if (g.oh <= 0 || g.ow <= 0) return fail("empty correlation output (%d x %d)", g.oh, g.ow) != 0;
    if ((size_t)B * g.oc * g.oh * g.ow >= ((size_t)1 << 40)) return fail("output too large") != 0;
    return true;
}

bool fast_path(const pwc::CorrGeom& g)
{
    return !g_force_generic.load() && g.k == 1 && g.s1 == 1 && g.pad == g.md && g.D == 9 &&
           (g.s2 == 1 || g.s2 == 2);
}

constexpr int B200_SMS = 148;      // only the fallback when the attribute query fails

// SM count of the current device (cached per thread and device): every launch heuristic sizes its grid from it
int sm_count_of_current_device()
{
    static thread_local int dev_cached = -1, count = B200_SMS;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) {
        if (cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) count = B200_SMS;
        (void)cudaGetLastError();
        dev_cached = dev;
    }
    return count;
}

// The shared-memory opt-in of a kernel (cudaFuncSetAttribute) is per (function, device).  Every launcher keeps one bit per
// device in a static of its own template instantiation: set once per device, for all host threads, without re-setting
// the attribute when a thread alternates between devices.  (cudaFuncSetAttribute is idempotent, so a race between two
// first callers is harmless.)
inline bool device_configured(const std::atomic<unsigned long long>& mask, int dev)
{
    return (mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull;
}
inline void mark_device_configured(std::atomic<unsigned long long>& mask, int dev)
{
    mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
}

// Where the forward kernels take the flow from (model.py:74-80):
//   flow   : [2][H][W] per image, image n at flow + n * fbs floats (fbs == 2*H*W when dense), or NULL (no warp)
//   coarse : when non-NULL, the previous level's flow [B][2][H/2][W/2] (dense); the kernel evaluates
//            model.py:78 itself (flow = F.upsample(coarse, 2, 'bilinear') * 2) and writes the fine flow to
//            flow_out + n * fobs.  Kernels without the fold are given flow = flow_out after a prepass.
struct FlowSpec {
    const float* flow;
    long long fbs;
    const float* coarse;
    float* flow_out;
    long long fobs;
};

template <class Cfg, bool HAS_FLOW>
int launch_fwd_tiled(const float* f1, const float* f2, const float* flow, long long fbs, float* out,
                     float* warped, const pwc::CorrGeom& g, int act, float slope, long long obs, cudaStream_t st)
{
    auto kern = pwc::warpcorr_fwd_kernel<Cfg, HAS_FLOW>;
    const size_t smem = Cfg::smem_bytes(HAS_FLOW);
    static std::atomic<unsigned long long> configured_devs{0};   // one bit per device (the opt-in is per function and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!device_configured(configured_devs, dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(cudaGetLastError()));
        mark_device_configured(configured_devs, dev);
    }
    const int tiles_x = pwc::cdiv(g.W, Cfg::TW), tiles_y = pwc::cdiv(g.H, Cfg::TH);
    const long long blocks = (long long)tiles_x * tiles_y * g.B;
    if (blocks > 0x0fffffffLL) return fail("grid too large");
    // few tiles but many channels (small pyramid levels): split the channels over a thread-block
    // cluster of up to 8 CTAs per tile (deterministic DSMEM reduction inside the kernel)
    int ksplit = 1;
    while (ksplit < 8 && blocks * ksplit < sm_count_of_current_device() && g.C / (ksplit * 2) >= 8) ksplit *= 2;
    const int cper = pwc::cdiv(g.C, ksplit);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(blocks * ksplit));
    cfg.blockDim = dim3(Cfg::NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ksplit;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, f1, f2, flow, out, warped, g.C, g.H, g.W, tiles_x, tiles_y, act, slope,
                           ksplit, cper, obs, fbs) != cudaSuccess)
        return fail("warpcorr_fwd_kernel launch: %s", cudaGetErrorString(cudaGetLastError()));
    return check_launch("warpcorr_fwd_kernel");
}


// ---- whole-image path for the coarse levels (small_image.cuh) -----------------------------------
constexpr size_t SMALL_SMEM_LIMIT = 200 * 1024;

pwc::SmallPlan small_plan_for(const pwc::CorrGeom& g, bool has_flow, bool backward)
{
    pwc::SmallPlan none = {0, 0, 0, 0, false};
    if (g_disable_small.load() || !fast_path(g) || (long long)g.H * g.W > pwc::SMALL_MAX_PX) return none;
    const pwc::SmallPlan p = pwc::small_plan(g.B, g.C, g.H * g.W, has_flow, backward, SMALL_SMEM_LIMIT,
                                             sm_count_of_current_device());
    if (!p.ok || (long long)g.B * p.ks > 0x3fffffffLL) return none;
    return p;
}

template <class Kern, class... Args>
int launch_small(Kern kern, const char* what, size_t smem, int B, int ks, cudaStream_t st, Args... args)
{
    // the shared-memory opt-in is per (function, device); this launcher is shared by several kernels of one
    // signature, so it keeps a small per-thread table of (kernel, devices already configured)
    struct Entry { const void* fn; unsigned long long devs; };
    static thread_local Entry configured[16];
    static thread_local int nconfigured = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* fn = reinterpret_cast<const void*>(kern);
    Entry* e = nullptr;
    for (int i = 0; i < nconfigured; ++i)
        if (configured[i].fn == fn) e = &configured[i];
    if (e == nullptr && nconfigured < 16) {
        e = &configured[nconfigured++];
        e->fn = fn;
        e->devs = 0;
    }
    if (e == nullptr || !((e->devs >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_LIMIT) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", SMALL_SMEM_LIMIT, cudaGetErrorString(cudaGetLastError()));
        if (e != nullptr) e->devs |= 1ull << (dev & 63);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * ks));
    cfg.blockDim = dim3(pwc::SMALL_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, args...) != cudaSuccess)
        return fail("%s launch: %s", what, cudaGetErrorString(cudaGetLastError()));
    return check_launch(what);
}

template <int S2, bool HAS_FLOW>
int launch_fwd_small(const float* f1, const float* f2, const FlowSpec& fs, float* out, float* warped,
                     const pwc::CorrGeom& g, int act, float slope, long long obs, cudaStream_t st)
{
    const pwc::SmallPlan p = small_plan_for(g, HAS_FLOW, false);
    return launch_small(pwc::warpcorr_fwd_small_kernel<S2, HAS_FLOW>, "warpcorr_fwd_small_kernel", p.smem, g.B, p.ks, st, f1, f2, fs.flow, out, warped, g.C, g.H,
                        g.W, p.cs, p.csp, act, slope, obs, fs.fbs, fs.coarse, fs.flow_out, fs.fobs);
}

template <int S2, bool HAS_FLOW>
int launch_bwd_small(const float* gout, const float* gate, const float* f1, const float* f2, const float* flow,
                     float* gf1, float* gf2, float* gflow, const pwc::CorrGeom& g, float slope, long long gbs,
                     long long gate_bs, cudaStream_t st)
{
    const pwc::SmallPlan p = small_plan_for(g, HAS_FLOW, true);
    return launch_small(pwc::warpcorr_bwd_small_kernel<S2, HAS_FLOW>, "warpcorr_bwd_small_kernel", p.smem, g.B, p.ks, st, gout, gate, f1, f2, flow, gf1, gf2,
                        gflow, g.C, g.H, g.W, p.cs, p.csp, slope, gbs, gate_bs);
}


// ---- TMA path ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// 4-D tensor map over a dense [B][C][H][W] fp32 tensor, box = (bw, bh, bc, 1); out-of-range box
// elements (image border, channel tail) are filled with zeros by the TMA unit.
// batch_stride: floats between consecutive images (0 = dense, C*H*W); must be a multiple of 4 floats.
bool make_nchw_map(CUtensorMap* map, const float* ptr, int B, int C, int H, int W, int bw, int bh, int bc,
                   long long batch_stride = 0, bool swizzle64 = false)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    if (batch_stride == 0) batch_stride = (long long)W * H * C;
    if ((batch_stride & 3) != 0 || ((W * 4) & 15) != 0) return false;      // TMA global strides are multiples of 16 bytes
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)batch_stride * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// obs: batch stride (floats) of the tensor the kernel stores to with 128-/256-bit stores; every image's
// base must be 16-byte aligned, i.e. the stride a multiple of 4 floats (0 = dense)
bool tma_eligible(const float* f1, const float* f2, const float* out, const pwc::CorrGeom& g, long long obs = 0)
{
    if (g_disable_tma.load()) return false;
    if ((obs & 3) != 0) return false;
    if ((g.W & 3) != 0 || g.W < 16 || g.H < 8 || g.W >= 32760 || g.H >= 32760) return false;
    if (((uintptr_t)f1 | (uintptr_t)f2 | (uintptr_t)out) & 15) return false;
    return true;
}

// returns 1 ok, 0 error, -1 "not taken" (caller falls back to the plain tiled kernel)
template <int S2, int CK, bool HAS_FLOW>
int launch_fwd_tma(const float* f1, const float* f2, const FlowSpec& fs, float* out, float* warped,
                   const pwc::CorrGeom& g, int act, float slope, long long obs, cudaStream_t st)
{
    using Cfg = pwc::TmaCfg<S2, CK>;
    const float* flow = fs.flow;
    CUtensorMap m1, m2;
    if (!make_nchw_map(&m1, f1, g.B, g.C, g.H, g.W, Cfg::F1W, Cfg::F1H, CK)) return -1;
    if (!make_nchw_map(&m2, f2, g.B, g.C, g.H, g.W, HAS_FLOW ? Cfg::WW : Cfg::WP, HAS_FLOW ? Cfg::WH : Cfg::HH, CK))
        return -1;
    CUtensorMap m3 = m2;   // unused without flow
    if (HAS_FLOW) {
        if (fs.coarse) {   // model.py:78 folded into the flow read: the map covers the coarse flow
            if (((uintptr_t)fs.coarse & 15) != 0 || (g.H & 1) || (g.W & 7)) return -1;
            if (!make_nchw_map(&m3, fs.coarse, g.B, 2, g.H / 2, g.W / 2, Cfg::CWB, Cfg::CHB, 2)) return -1;
        } else {
            if (((uintptr_t)flow & 15) != 0) return -1;
            if (!make_nchw_map(&m3, flow, g.B, 2, g.H, g.W, Cfg::HWD, Cfg::HH, 2, fs.fbs)) return -1;
        }
    }
    auto kern = pwc::warpcorr_fwd_tma_kernel<Cfg, HAS_FLOW>;
    const size_t smem = Cfg::smem_bytes(HAS_FLOW);
    static std::atomic<unsigned long long> configured_devs{0};   // one bit per device (the opt-in is per function and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!device_configured(configured_devs, dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(cudaGetLastError()));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        mark_device_configured(configured_devs, dev);
    }
    const int tiles_x = pwc::cdiv(g.W, Cfg::TW), tiles_y = pwc::cdiv(g.H, Cfg::TH);
    const long long ntiles = (long long)tiles_x * tiles_y * g.B;
    if (ntiles > 0x3fffffffLL) return fail("grid too large");
    const int sm_count = sm_count_of_current_device();
    // persistent: one CTA per SM (148 on B200), each walks tiles blockIdx.x, blockIdx.x + grid, ...
    const unsigned grid = (unsigned)(ntiles < sm_count ? ntiles : sm_count);
    kern<<<grid, Cfg::NT, smem, st>>>(m1, m2, m3, f2, flow, out, warped, g.C, g.H, g.W, tiles_x, tiles_y, (int)ntiles,
                                      act, slope, obs, fs.fbs, fs.coarse, fs.flow_out, fs.fobs);
    return check_launch("warpcorr_fwd_tma_kernel");
}

template <int S2, bool HAS_FLOW>
int dispatch_fwd_tiled(const float* f1, const float* f2, const float* flow, long long fbs, float* out,
                       float* warped, const pwc::CorrGeom& g, int act, float slope, long long obs, cudaStream_t st)
{
    if (g.W > 16)
        return launch_fwd_tiled<pwc::FwdCfg<9, S2, 8, 4, 8>, HAS_FLOW>(f1, f2, flow, fbs, out, warped, g, act, slope, obs, st);
    // small images (6x7, 12x14 levels: few CTAs, many channels): deep channel chunks, so that the
    // serial chunk loop is short and each chunk keeps many gathers in flight
    if (g.C >= 64)
        return launch_fwd_tiled<pwc::FwdCfg<9, S2, 4, 4, (S2 == 1 ? 28 : 16)>, HAS_FLOW>(f1, f2, flow, fbs, out, warped, g, act, slope, obs, st);
    return launch_fwd_tiled<pwc::FwdCfg<9, S2, 4, 4, 8>, HAS_FLOW>(f1, f2, flow, fbs, out, warped, g, act, slope, obs, st);
}

// model.py:78 as its own launch (where the fold into the fused kernel does not apply): fs.coarse -> fs.flow_out,
// after which the flow is an ordinary strided input.
int flow_prepass(FlowSpec& fs, const pwc::CorrGeom& g, cudaStream_t st)
{
    const size_t total = (size_t)g.B * g.H * g.W;
    pwc::flow_up2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(fs.coarse, fs.flow_out, fs.fobs, g.B, g.H, g.W);
    if (!check_launch("flow_up2_kernel")) return 0;
    fs.flow = fs.flow_out;
    fs.fbs = fs.fobs;
    fs.coarse = nullptr;
    return 1;
}

int forward_impl(const float* f1, const float* f2, FlowSpec fs, float* out, float* warped,
                 const pwc::CorrGeom& g, int act, float slope, long long obs, cudaStream_t st)
{
    const long long dense = (long long)g.oc * g.oh * g.ow;
    if (obs == 0) obs = dense;
    if (obs < dense) return fail("output batch stride %lld is smaller than one image's output (%lld)", obs, dense);
    const long long fdense = 2LL * g.H * g.W;
    if (fs.fbs == 0) fs.fbs = fdense;
    if (fs.fobs == 0) fs.fobs = fdense;
    if (fs.flow && fs.fbs < fdense) return fail("flow batch stride %lld is smaller than one image's flow (%lld)", fs.fbs, fdense);
    if (fs.coarse) {
        if ((g.H & 1) || (g.W & 1)) return fail("coarse flow: H and W must be even (got %d x %d)", g.H, g.W);
        if (!fs.flow_out) return fail("coarse flow: flow_out is required (the flow estimator needs the fine flow)");
        if (fs.fobs < fdense) return fail("flow_out batch stride %lld is smaller than one image's flow (%lld)", fs.fobs, fdense);
    }
    const bool has_flow = fs.flow != nullptr || fs.coarse != nullptr;
    if (fast_path(g)) {
        if (warped && !has_flow) {   // no warp: x2_warp is x2 itself (model.py:80 with zero displacement)
            if (cudaMemcpyAsync(warped, f2, sizeof(float) * (size_t)g.B * g.C * g.H * g.W,
                                cudaMemcpyDeviceToDevice, st) != cudaSuccess)
                return fail("cudaMemcpyAsync(warped_out): %s", cudaGetErrorString(cudaGetLastError()));
            warped = nullptr;
        }
        if (small_plan_for(g, has_flow, false).ok) {      // folds the coarse flow itself
            if (g.s2 == 1)
                return has_flow ? launch_fwd_small<1, true>(f1, f2, fs, out, warped, g, act, slope, obs, st)
                                : launch_fwd_small<1, false>(f1, f2, fs, out, warped, g, act, slope, obs, st);
            return has_flow ? launch_fwd_small<2, true>(f1, f2, fs, out, warped, g, act, slope, obs, st)
                            : launch_fwd_small<2, false>(f1, f2, fs, out, warped, g, act, slope, obs, st);
        }
        if (tma_eligible(f1, f2, out, g, obs)) {          // folds it when W % 8 == 0 (TMA stride rule), else -1
            int rc;
            if (g.s2 == 1)
                rc = has_flow ? launch_fwd_tma<1, 4, true>(f1, f2, fs, out, warped, g, act, slope, obs, st)
                              : launch_fwd_tma<1, 4, false>(f1, f2, fs, out, warped, g, act, slope, obs, st);
            else
                rc = has_flow ? launch_fwd_tma<2, 2, true>(f1, f2, fs, out, warped, g, act, slope, obs, st)
                              : launch_fwd_tma<2, 2, false>(f1, f2, fs, out, warped, g, act, slope, obs, st);
            if (rc >= 0) return rc;
            if (fs.coarse) {                              // not foldable here: prepass, then the TMA kernel again
                if (!flow_prepass(fs, g, st)) return 0;
                rc = g.s2 == 1 ? launch_fwd_tma<1, 4, true>(f1, f2, fs, out, warped, g, act, slope, obs, st)
                               : launch_fwd_tma<2, 2, true>(f1, f2, fs, out, warped, g, act, slope, obs, st);
                if (rc >= 0) return rc;
            }
        }
        if (fs.coarse && !flow_prepass(fs, g, st)) return 0;
        const float* flow = fs.flow;
        if (g.s2 == 1)
            return flow ? dispatch_fwd_tiled<1, true>(f1, f2, flow, fs.fbs, out, warped, g, act, slope, obs, st)
                        : dispatch_fwd_tiled<1, false>(f1, f2, flow, fs.fbs, out, warped, g, act, slope, obs, st);
        return flow ? dispatch_fwd_tiled<2, true>(f1, f2, flow, fs.fbs, out, warped, g, act, slope, obs, st)
                    : dispatch_fwd_tiled<2, false>(f1, f2, flow, fs.fbs, out, warped, g, act, slope, obs, st);
    }
    if (fs.coarse && !flow_prepass(fs, g, st)) return 0;
    const float* flow = fs.flow;
    const size_t total = (size_t)g.B * g.oc * g.oh * g.ow;
    pwc::corr_fwd_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(f1, f2, flow, out, g, act, slope, obs, fs.fbs);
    if (!check_launch("corr_fwd_generic_kernel")) return 0;
    if (warped) {
        if (flow && fs.fbs != fdense) return fail("warped_out with a strided flow is only available on the 81-displacement fast path");
        if (flow) return pwc_warp_forward(f2, flow, warped, g.B, g.C, g.H, g.W, st);
        if (cudaMemcpyAsync(warped, f2, sizeof(float) * (size_t)g.B * g.C * g.H * g.W,
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return fail("cudaMemcpyAsync(warped_out): %s", cudaGetErrorString(cudaGetLastError()));
    }
    return 1;
}

template <int S2, int SIGN, int TW = 32, int TH = 8>
int launch_bwd_tiled_cfg(const float* gout, const float* gate, const float* X, float* res,
                         const pwc::CorrGeom& g, float slope, long long gbs, long long gate_bs, cudaStream_t st)
{
    using Cfg = pwc::BwdCfg<9, S2, 16, TW, TH>;
    auto kern = pwc::corr_bwd_kernel<Cfg, SIGN>;
    const size_t smem = Cfg::smem_bytes();
    static std::atomic<unsigned long long> configured_devs{0};   // one bit per device (the opt-in is per function and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!device_configured(configured_devs, dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(cudaGetLastError()));
        mark_device_configured(configured_devs, dev);
    }
    const int tiles_x = pwc::cdiv(g.W, Cfg::TW), tiles_y = pwc::cdiv(g.H, Cfg::TH);
    const long long blocks = (long long)tiles_x * tiles_y * g.B;
    if (blocks > 0x7fffffffLL) return fail("grid too large");
    // split the (independent) output channels over blockIdx.y until the grid covers ~2 waves
    int cgroup = g.C;
    while (cgroup > Cfg::CK && blocks * pwc::cdiv(g.C, cgroup) < 2 * sm_count_of_current_device()) cgroup = pwc::round_up(pwc::cdiv(cgroup, 2), Cfg::CK);
    const dim3 grid((unsigned)blocks, (unsigned)pwc::cdiv(g.C, cgroup));
    kern<<<grid, Cfg::NT, smem, st>>>(gout, gate, X, res, g.C, g.H, g.W, tiles_x, tiles_y, cgroup, slope, gbs, gate_bs);
    return check_launch("corr_bwd_kernel");
}

// returns 1 ok, 0 error, -1 "not taken" (caller falls back to the plain tiled kernel)
template <int S2, int CK, int SIGN>
int launch_bwd_tma(const float* gout, long long gbs, const float* X, float* res, const pwc::CorrGeom& g, cudaStream_t st)
{
    using Cfg = pwc::BwdTmaCfg<S2, CK>;
    CUtensorMap mX, mG;
    if (!make_nchw_map(&mX, X, g.B, g.C, g.H, g.W, Cfg::WP, Cfg::HH, CK)) return -1;
    // output-gradient prefetch box: the tile (g1) or the tile + halo (gradient w.r.t. the second operand)
    if (!make_nchw_map(&mG, gout, g.B, 81, g.H, g.W, SIGN > 0 ? Cfg::TW : Cfg::HWD, SIGN > 0 ? Cfg::TH : Cfg::HH,
                       Cfg::GBOX_C, gbs))
        return -1;
    auto kern = pwc::corr_bwd_tma_kernel<Cfg, SIGN>;
    const size_t smem = Cfg::smem_bytes();
    static std::atomic<unsigned long long> configured_devs{0};   // one bit per device (the opt-in is per function and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!device_configured(configured_devs, dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(cudaGetLastError()));
        mark_device_configured(configured_devs, dev);
    }
    const int tiles_x = pwc::cdiv(g.W, Cfg::TW), tiles_y = pwc::cdiv(g.H, Cfg::TH);
    const long long ntiles = (long long)tiles_x * tiles_y * g.B;
    if (ntiles > 0x3fffffffLL) return fail("grid too large");
    const int sm_count = sm_count_of_current_device();
    const unsigned grid = (unsigned)(ntiles < sm_count ? ntiles : sm_count);
    kern<<<grid, Cfg::NT, smem, st>>>(mX, mG, gout, res, g.C, g.H, g.W, tiles_x, tiles_y, (int)ntiles, gbs);
    return check_launch("corr_bwd_tma_kernel");
}

// stride2 == 1: threads own complete outputs (corr_bwd_seq.cuh).  returns 1 ok, 0 error, -1 "not taken"
template <int SIGN>
int launch_bwd_seq(const float* gout, long long gbs, const float* X, float* res, const pwc::CorrGeom& g, cudaStream_t st)
{
    using Cfg = pwc::BwdSeqCfg;
    CUtensorMap mX, mG;
    if (!make_nchw_map(&mX, X, g.B, g.C, g.H, g.W, Cfg::WP, Cfg::HH, Cfg::CPI)) return -1;
    if (!make_nchw_map(&mG, gout, g.B, 81, g.H, g.W, SIGN > 0 ? Cfg::TW : Cfg::HWD, SIGN > 0 ? Cfg::TH : Cfg::HH,
                       Cfg::GBOX_C, gbs))
        return -1;
    // g1: the taps of a displacement row are an unshifted [9][16][16] box of the output gradient -- TMA writes it
    // straight into the tap ring (64-byte swizzle == tap_slot()); the other gradient reads them at x - dx, where a
    // tiled load cannot start (16-byte rule), and keeps the staging warps
    CUtensorMap mTap = mG;
    int tma_taps = 0;
    if (SIGN > 0 && make_nchw_map(&mTap, gout, g.B, 81, g.H, g.W, Cfg::TW, Cfg::TH, Cfg::D, gbs, true)) tma_taps = 1;
    auto kern = pwc::corr_bwd_seq_kernel<SIGN>;
    const size_t smem = Cfg::smem_bytes();
    static std::atomic<unsigned long long> configured_devs{0};   // one bit per device (the opt-in is per function and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!device_configured(configured_devs, dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(cudaGetLastError()));
        mark_device_configured(configured_devs, dev);
    }
    const int tiles_x = pwc::cdiv(g.W, Cfg::TW), tiles_y = pwc::cdiv(g.H, Cfg::TH);
    const int nsc = pwc::cdiv(g.C, Cfg::CPI);
    const long long nitems = (long long)tiles_x * tiles_y * g.B * nsc;
    if (nitems > 0x3fffffffLL) return fail("grid too large");
    const int sms = sm_count_of_current_device();
    const unsigned grid = (unsigned)(nitems < sms ? nitems : sms);
    kern<<<grid, Cfg::NT, smem, st>>>(mX, mG, mTap, tma_taps, gout, res, g.C, g.H, g.W, tiles_x, tiles_y, (int)nitems, nsc, gbs);
    return check_launch("corr_bwd_seq_kernel");
}

template <int S2, int SIGN>
int launch_bwd_tiled(const float* gout, const float* gate, const float* X, float* res,
                     const pwc::CorrGeom& g, float slope, long long gbs, long long gate_bs, cudaStream_t st)
{
    if (g.W <= 8 && g.H <= 8) return launch_bwd_tiled_cfg<S2, SIGN, 8, 8>(gout, gate, X, res, g, slope, gbs, gate_bs, st);
    if (g.W <= 16) return launch_bwd_tiled_cfg<S2, SIGN, 16, 16>(gout, gate, X, res, g, slope, gbs, gate_bs, st);
    return launch_bwd_tiled_cfg<S2, SIGN, 32, 8>(gout, gate, X, res, g, slope, gbs, gate_bs, st);
}

// g1 (w.r.t. f1) and g2 (w.r.t. the second operand as given, i.e. the warped features).
// which: 1 = g1 only, 2 = g2 only, 3 = both
// gated_scratch (optional, B*81*H*W floats, 16-byte aligned): room for the LeakyReLU-gated output
// gradient; without it a gated backward stays on the plain tiled kernels.  *gated_done tells a caller
// that splits the two gradients over two calls that the scratch already holds the gated gradient.
// Strides of the backward's strided inputs: gbs = floats between consecutive images of grad_out, gate_bs = the
// same for the forward output that gates it (act); 0 = dense.  This is how the gradient of the flow estimator's
// concatenated input [x1 | corr | flow] (model.py:89-91) is consumed in place: grad_out = grad_in + C*H*W with
// batch stride (C + 83)*H*W, no slice copy.
struct GradStrides { long long gbs, gate_bs; };

int corr_backward_impl(const float* gout, const float* gate, const float* f1, const float* second,
                       float* g1, float* g2, const pwc::CorrGeom& g, float slope, cudaStream_t st, int which = 3,
                       float* gated_scratch = nullptr, bool* gated_done = nullptr, GradStrides gs = {0, 0})
{
    if (g.s1 != 1)
        return fail("correlation backward requires stride1 == 1 (got %d): the reference kernels "
                    "address gradInput out of range otherwise", g.s1);
    const long long dense = (long long)g.oc * g.oh * g.ow;
    long long gbs = gs.gbs ? gs.gbs : dense, gate_bs = gs.gate_bs ? gs.gate_bs : dense;
    if (gbs < dense || gate_bs < dense) return fail("grad_out / out batch stride smaller than one image (%lld)", dense);
    if (which == 3 && small_plan_for(g, false, true).ok)
        return g.s2 == 1 ? launch_bwd_small<1, false>(gout, gate, f1, second, nullptr, g1, g2, nullptr, g, slope, gbs, gate_bs, st)
                         : launch_bwd_small<2, false>(gout, gate, f1, second, nullptr, g1, g2, nullptr, g, slope, gbs, gate_bs, st);
    if (fast_path(g)) {
        const float* any_in = second ? second : f1;
        float* any_out = g1 ? g1 : g2;
        const bool gate_ok = !gate || (gated_scratch && (((uintptr_t)gate | (uintptr_t)gated_scratch) & 15) == 0 &&
                                       (gate_bs & 3) == 0);
        if (gate_ok && (gbs & 3) == 0 && tma_eligible(f1, any_in, any_out, g) &&
            (((uintptr_t)g2 | (uintptr_t)g1 | (uintptr_t)gout) & 15) == 0) {
            const float* go = gout;
            long long go_bs = gbs;
            if (gate) {      // LeakyReLU backward as its own pass (the TMA kernels stage raw taps); dense result
                if (!gated_done || !*gated_done) {
                    const size_t per4 = (size_t)dense / 4;                     // W % 4 == 0 on this path
                    const size_t n4 = (size_t)g.B * per4;
                    pwc::gate_grad_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
                        reinterpret_cast<const float4*>(gout), reinterpret_cast<const float4*>(gate),
                        reinterpret_cast<float4*>(gated_scratch), n4, slope, per4, (size_t)gbs / 4, (size_t)gate_bs / 4);
                    if (!check_launch("gate_grad_kernel")) return 0;
                    if (gated_done) *gated_done = true;
                }
                go = gated_scratch;
                go_bs = dense;
            }
            int r1 = 1, r2 = 1;
            if (g.s2 == 1 && !g_disable_seq.load()) {
                // threads own complete outputs (corr_bwd_seq.cuh): 90 / 100 us vs 111 / 108 us for the
                // slice/reduce kernel at the level-2 shape
                if (which & 1) r1 = launch_bwd_seq<+1>(go, go_bs, second, g1, g, st);
                if ((which & 2) && r1 > 0) r2 = launch_bwd_seq<-1>(go, go_bs, f1, g2, g, st);
            } else if (g.s2 == 1) {
                if (which & 1) r1 = launch_bwd_tma<1, 4, +1>(go, go_bs, second, g1, g, st);
                if ((which & 2) && r1 > 0) r2 = launch_bwd_tma<1, 4, -1>(go, go_bs, f1, g2, g, st);
            } else {
                if (which & 1) r1 = launch_bwd_tma<2, 2, +1>(go, go_bs, second, g1, g, st);
                if ((which & 2) && r1 > 0) r2 = launch_bwd_tma<2, 2, -1>(go, go_bs, f1, g2, g, st);
            }
            if (r1 == 0 || r2 == 0) return 0;
            if (r1 > 0 && r2 > 0) return 1;
        }
        if (g.s2 == 1)
            return (!(which & 1) || launch_bwd_tiled<1, +1>(gout, gate, second, g1, g, slope, gbs, gate_bs, st)) &&
                   (!(which & 2) || launch_bwd_tiled<1, -1>(gout, gate, f1, g2, g, slope, gbs, gate_bs, st));
        return (!(which & 1) || launch_bwd_tiled<2, +1>(gout, gate, second, g1, g, slope, gbs, gate_bs, st)) &&
               (!(which & 2) || launch_bwd_tiled<2, -1>(gout, gate, f1, g2, g, slope, gbs, gate_bs, st));
    }
    const size_t total = (size_t)g.B * g.C * g.H * g.W;
    pwc::corr_bwd_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gout, gate, f1, second, g1, g2, g, slope,
                                                                                 gbs, gate_bs);
    return check_launch("corr_bwd_generic_kernel");
}

// Feature-gradient scatter through the 8-channel-interleaved scratch (see warp_bwd_v8_kernel).
// scratch holds B * ceil(C/8) * H * W * 8 floats, 16-byte aligned.
// Launch `kern` so that it may start while the kernel before it in the stream is still running its tail
// (programmatic dependent launch).  Only for kernels that do NOT read what that kernel writes.
template <class Kern, class... Args>
cudaError_t launch_into_tail(Kern kern, dim3 grid, dim3 block, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

// Zeroes the scatter scratch (B * ceil(C/8) * H * W * 8 floats) and the flow gradient.  Launched right
// behind the kernel that computes the gradient w.r.t. the warped features, into its tail.
int scatter_zero(float* scratch, float* grad_flow, int B, int C, int H, int W, cudaStream_t stream)
{
    const size_t n8 = (size_t)B * pwc::cdiv(C, 8) * H * W * 8, nf = (size_t)B * 2 * H * W;
    if ((nf & 3) == 0 && (((uintptr_t)scratch | (uintptr_t)grad_flow) & 15) == 0) {
        const unsigned grid = (unsigned)std::min<size_t>((n8 / 4 + nf / 4 + 255) / 256, (size_t)sm_count_of_current_device() * 8);
        if (launch_into_tail(pwc::zero2_kernel, dim3(grid), dim3(256), stream, reinterpret_cast<float4*>(scratch), n8 / 4,
                             reinterpret_cast<float4*>(grad_flow), nf / 4) != cudaSuccess)
            return fail("zero2_kernel launch: %s", cudaGetErrorString(cudaGetLastError()));
        return check_launch("zero2_kernel");
    }
    if (cudaMemsetAsync(scratch, 0, sizeof(float) * n8, stream) != cudaSuccess ||
        cudaMemsetAsync(grad_flow, 0, sizeof(float) * nf, stream) != cudaSuccess)
        return fail("cudaMemsetAsync(scatter scratch): %s", cudaGetErrorString(cudaGetLastError()));
    return 1;
}

// warp_bwd_tile_kernel: the tiled backward of the bilinear warp.
// returns 1 ok, 0 error, -1 "not taken" (TMA does not apply: caller falls back)
int launch_warp_tile(const float* grad_out, const float* x, const float* flow, float* scratch, float* grad_flow,
                     float* warped_out, int B, int C, int H, int W, cudaStream_t stream)
{
    if (g_disable_tma.load() || (W & 3) != 0 || W < 16 || H < 8 || W >= 32760 || H >= 32760 ||
        (((uintptr_t)x | (uintptr_t)scratch) & 15) != 0)
        return -1;
    using Cfg = pwc::WarpBwdCfg;
    CUtensorMap mX;
    if (!make_nchw_map(&mX, x, B, C, H, W, Cfg::WW, Cfg::WH, Cfg::CK)) return -1;
    auto kern = pwc::warp_bwd_tile_kernel;
    const size_t smem = Cfg::smem_bytes();
    static std::atomic<unsigned long long> configured_devs{0};   // one bit per device (the opt-in is per function and device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (!device_configured(configured_devs, dev)) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(cudaGetLastError()));
        mark_device_configured(configured_devs, dev);
    }
    const int cocts = pwc::cdiv(C, 8);
    const int tiles_x = pwc::cdiv(W, Cfg::TW), tiles_y = pwc::cdiv(H, Cfg::TH);
    const long long ntiles = (long long)tiles_x * tiles_y * B;
    if (ntiles > 0x3fffffffLL) return fail("grid too large");
    const long long cap = 3LL * sm_count_of_current_device();
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    kern<<<grid, Cfg::NT, smem, stream>>>(mX, grad_out, x, flow, scratch, grad_flow, warped_out, C, H, W, tiles_x,
                                         tiles_y, (int)ntiles, cocts);
    return check_launch("warp_bwd_tile_kernel");
}

// Feature-gradient scatter through the 8-channel-interleaved scratch (see warp_bwd_v8_kernel); the scratch
// and grad_flow must have been zeroed (scatter_zero).  scratch is 16-byte aligned.
int scatter_accumulate(const float* grad_out, const float* x, const float* flow, float* grad_flow, float* scratch,
                       float* warped_out, int B, int C, int H, int W, cudaStream_t stream)
{
    const int cocts = pwc::cdiv(C, 8);
    if ((long long)H * W > 0x3fffffffLL || B > 65535 || cocts > 65535) return fail("warp backward: tensor too large");
    // tiled kernel (corner values from a TMA-staged shared-memory window) where TMA applies
    {
        const int rc = launch_warp_tile(grad_out, x, flow, scratch, grad_flow, warped_out, B, C, H, W, stream);
        if (rc >= 0) return rc;
    }
    const dim3 grid((unsigned)pwc::cdiv(2 * H * W, 256), (unsigned)cocts, (unsigned)B);
    pwc::warp_bwd_v8_kernel<<<grid, 256, 0, stream>>>(grad_out, x, flow, scratch, grad_flow, warped_out, B, C, H, W, cocts);
    return check_launch("warp_bwd_v8_kernel");
}

// scratch -> NCHW grad_x.  `into_tail`: the kernel before it in the stream does not write the scratch, so
// the de-interleave may start in that kernel's tail.
int scatter_finish(const float* scratch, float* grad_x, int B, int C, int H, int W, bool into_tail, cudaStream_t stream)
{
    const int cocts = pwc::cdiv(C, 8);
    const size_t total = (size_t)B * H * W * cocts;
    const dim3 grid((unsigned)((total + 255) / 256));
    if (into_tail) {
        if (launch_into_tail(pwc::deinterleave8_kernel, grid, dim3(256), stream, scratch, grad_x, B, C, H, W, cocts) != cudaSuccess)
            return fail("deinterleave8_kernel launch: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
        pwc::deinterleave8_kernel<<<grid, 256, 0, stream>>>(scratch, grad_x, B, C, H, W, cocts);
    }
    return check_launch("deinterleave8_kernel");
}

int warp_backward_scratch(const float* grad_out, const float* x, const float* flow, float* grad_x, float* grad_flow,
                     float* scratch, float* warped_out, int B, int C, int H, int W, cudaStream_t stream)
{
    return scatter_zero(scratch, grad_flow, B, C, H, W, stream) &&
           scatter_accumulate(grad_out, x, flow, grad_flow, scratch, warped_out, B, C, H, W, stream) &&
           scatter_finish(scratch, grad_x, B, C, H, W, false, stream);
}

}  // namespace

extern "C" {

const char* pwc_last_error(void) { return g_err; }
int pwc_abi_version(void) { return PWC_B200_ABI_VERSION; }
long long pwc_launch_count(void) { return g_launches.load(); }
int pwc_set_force_generic(int on) { return g_force_generic.exchange(on ? 1 : 0); }
int pwc_set_disable_tma(int on) { return g_disable_tma.exchange(on ? 1 : 0); }
int pwc_set_disable_small(int on) { return g_disable_small.exchange(on ? 1 : 0); }
int pwc_set_disable_seq(int on) { return g_disable_seq.exchange(on ? 1 : 0); }

int pwc_corr_output_shape(int H, int W, int pad_size, int kernel_size, int max_displacement,
                          int stride1, int stride2, int* out_channels, int* out_h, int* out_w)
{
    pwc::CorrGeom g;
    if (!make_geom(g, 1, 1, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    if (out_channels) *out_channels = g.oc;
    if (out_h) *out_h = g.oh;
    if (out_w) *out_w = g.ow;
    return 1;
}

int pwc_warp_forward(const float* x, const float* flow, float* out, int B, int C, int H, int W,
                     cudaStream_t stream)
{
    if (!x || !flow || !out) return fail("pwc_warp_forward: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("pwc_warp_forward: non-positive size");
    // (The tiled kernel of the backward, run forward-only, is no faster here: 50.0 vs 49.4 us at the level-2 shape, 36 vs
    // 29 us at level 3 -- the forward has no reductions to save, and its window ring adds hand-offs.)
    constexpr int CPT = 8;      // channels per thread: 8 measured best on B200 (2: 61, 4: 47, 8: 44, 16: 46 us at level 2)
    const size_t total = (size_t)B * pwc::cdiv(C, CPT) * H * W;
    pwc::warp_fwd_kernel<CPT><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(x, flow, out, B, C, H, W);
    return check_launch("warp_fwd_kernel");
}

int pwc_warp_backward(const float* grad_out, const float* x, const float* flow, float* grad_x,
                      float* grad_flow, int B, int C, int H, int W, cudaStream_t stream)
{
    if (!grad_out || !x || !flow) return fail("pwc_warp_backward: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("pwc_warp_backward: non-positive size");
    if (!grad_x && !grad_flow) return 1;
    if (grad_x &&
        cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)B * C * H * W, stream) != cudaSuccess)
        return fail("cudaMemsetAsync(grad_x): %s", cudaGetErrorString(cudaGetLastError()));
    // channel groups in parallel; the flow gradient is then a sum over groups (atomicAdd on a zeroed buffer)
    int cpt = C;
    while (cpt > 8 && (size_t)B * H * W * pwc::cdiv(C, cpt) < (size_t)sm_count_of_current_device() * 2048) cpt = pwc::cdiv(cpt, 2);
    const int cgroups = pwc::cdiv(C, cpt);
    if (grad_flow && cgroups > 1 &&
        cudaMemsetAsync(grad_flow, 0, sizeof(float) * (size_t)B * 2 * H * W, stream) != cudaSuccess)
        return fail("cudaMemsetAsync(grad_flow): %s", cudaGetErrorString(cudaGetLastError()));
    const size_t total = (size_t)B * H * W * cgroups;
    pwc::warp_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(grad_out, x, flow, grad_x, grad_flow, B, C, H, W, cpt, cgroups);
    return check_launch("warp_bwd_kernel");
}

long long pwc_warp_backward_workspace(int B, int C, int H, int W)
{
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)sizeof(float) * B * pwc::cdiv(C, 8) * 8 * H * W;
}

int pwc_warp_backward_ws(const float* grad_out, const float* x, const float* flow, float* grad_x, float* grad_flow,
                         int B, int C, int H, int W, void* workspace, long long workspace_bytes, cudaStream_t stream)
{
    if (!grad_out || !x || !flow) return fail("pwc_warp_backward_ws: null pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return fail("pwc_warp_backward_ws: non-positive size");
    // the tiled path needs both outputs (the flow gradient falls out of the same corner loads) and the scratch
    if (grad_x && grad_flow && workspace && workspace_bytes >= pwc_warp_backward_workspace(B, C, H, W) &&
        ((uintptr_t)workspace & 15) == 0 && (long long)H * W <= 0x3fffffffLL && B <= 65535)
        return warp_backward_scratch(grad_out, x, flow, grad_x, grad_flow, static_cast<float*>(workspace), nullptr, B, C, H, W,
                                     stream);
    return pwc_warp_backward(grad_out, x, flow, grad_x, grad_flow, B, C, H, W, stream);
}

int pwc_warpcorr_forward(const float* f1, const float* f2, const float* flow, float* out,
                         float* warped_out, int B, int C, int H, int W, int pad_size,
                         int kernel_size, int max_displacement, int stride1, int stride2, int act,
                         float slope, cudaStream_t stream)
{
    if (!f1 || !f2 || !out) return fail("pwc_warpcorr_forward: null pointer");
    pwc::CorrGeom g;
    if (!make_geom(g, B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    return forward_impl(f1, f2, FlowSpec{flow, 0, nullptr, nullptr, 0}, out, warped_out, g, act, slope, 0, stream);
}

int pwc_warpcorr_forward_strided(const float* f1, const float* f2, const float* flow, float* out,
                                 long long out_batch_stride, float* warped_out, int B, int C, int H, int W,
                                 int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                 int act, float slope, cudaStream_t stream)
{
    if (!f1 || !f2 || !out) return fail("pwc_warpcorr_forward_strided: null pointer");
    pwc::CorrGeom g;
    if (!make_geom(g, B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    return forward_impl(f1, f2, FlowSpec{flow, 0, nullptr, nullptr, 0}, out, warped_out, g, act, slope, out_batch_stride, stream);
}

int pwc_warpcorr_forward_coarse(const float* f1, const float* f2, const float* coarse_flow, float* out,
                                long long out_batch_stride, float* flow_out, long long flow_out_batch_stride,
                                float* warped_out, int B, int C, int H, int W, int pad_size, int kernel_size,
                                int max_displacement, int stride1, int stride2, int act, float slope,
                                cudaStream_t stream)
{
    if (!f1 || !f2 || !out || !coarse_flow || !flow_out) return fail("pwc_warpcorr_forward_coarse: null pointer");
    pwc::CorrGeom g;
    if (!make_geom(g, B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    return forward_impl(f1, f2, FlowSpec{nullptr, 0, coarse_flow, flow_out, flow_out_batch_stride}, out, warped_out, g,
                        act, slope, out_batch_stride, stream);
}

long long pwc_warpcorr_backward_workspace(int B, int C, int H, int W, int has_flow, int, int, int,
                                          int, int)
{
    if (!has_flow) return 0;
    // warped second operand + its gradient + the 8-channel-interleaved scatter scratch + the gated output gradient
    return (long long)sizeof(float) * (2LL * B * C * H * W + 8LL * B * ((C + 7) / 8) * H * W + 81LL * B * H * W);
}

static int warpcorr_backward_impl(const float* grad_out, GradStrides gs, const float* f1, const float* f2,
                          const float* flow, const float* out, const float* warped_in, float* grad_f1, float* grad_f2,
                          float* grad_flow, void* workspace, long long workspace_bytes, int B,
                          int C, int H, int W, int pad_size, int kernel_size, int max_displacement,
                          int stride1, int stride2, int act, float slope, cudaStream_t stream)
{
    if (!grad_out || !f1 || !f2 || !grad_f1 || !grad_f2) return fail("pwc_warpcorr_backward: null pointer");
    if (act && !out) return fail("pwc_warpcorr_backward: act != 0 needs the forward output");
    pwc::CorrGeom g;
    if (!make_geom(g, B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    const float* gate = act ? out : nullptr;
    const long long dense = (long long)g.oc * g.oh * g.ow;
    const long long gbs = gs.gbs ? gs.gbs : dense, gate_bs = gs.gate_bs ? gs.gate_bs : dense;
    if (!flow) return corr_backward_impl(grad_out, gate, f1, f2, grad_f1, grad_f2, g, slope, stream, 3, nullptr, nullptr, gs);
    if (!grad_flow) return fail("pwc_warpcorr_backward: grad_flow is required when flow is given");
    if (small_plan_for(g, true, true).ok)      // coarse levels: the whole backward is one launch, no workspace
        return g.s2 == 1 ? launch_bwd_small<1, true>(grad_out, gate, f1, f2, flow, grad_f1, grad_f2, grad_flow, g, slope, gbs, gate_bs, stream)
                         : launch_bwd_small<2, true>(grad_out, gate, f1, f2, flow, grad_f1, grad_f2, grad_flow, g, slope, gbs, gate_bs, stream);
    const long long need = pwc_warpcorr_backward_workspace(B, C, H, W, 1, pad_size, kernel_size,
                                                           max_displacement, stride1, stride2);
    if (!workspace || workspace_bytes < need)
        return fail("pwc_warpcorr_backward: workspace of %lld bytes required, got %lld", need, workspace_bytes);
    const size_t N = (size_t)B * C * H * W;
    float* wbuf = static_cast<float*>(workspace);
    float* gwarped = wbuf + N;
    float* scratch = gwarped + N;
    float* gated = scratch + 8 * (size_t)B * ((C + 7) / 8) * H * W;      // used only with act on the TMA kernels
    if (g.oc != 81) gated = nullptr;                                     // (sized for the 81-displacement fast path)
    bool gated_done = false;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(scratch) & 15) == 0;
    const bool split_ok = fast_path(g);      // the tiled kernels compute the two gradients in separate launches
    if (vec_ok && split_ok) {
        // 1. gradient w.r.t. the warped features (needs f1 and grad_out only)
        if (!corr_backward_impl(grad_out, gate, f1, nullptr, nullptr, gwarped, g, slope, stream, 2, gated, &gated_done, gs)) return 0;
        // 2. zero the scatter scratch: independent of step 1, launched into its tail
        if (!scatter_zero(scratch, grad_flow, B, C, H, W, stream)) return 0;
        // 3. scatter to the scratch + flow gradient; the same pass re-materialises x2_warp when the forward did
        //    not export it (both need the four bilinear corner values)
        if (!scatter_accumulate(gwarped, f2, flow, grad_flow, scratch, warped_in ? nullptr : wbuf, B, C, H, W, stream))
            return 0;
        // 4. gradient w.r.t. f1 (needs x2_warp, does not touch the scratch)
        if (!corr_backward_impl(grad_out, gate, f1, warped_in ? warped_in : wbuf, grad_f1, nullptr, g, slope, stream, 1, gated, &gated_done, gs))
            return 0;
        // 5. scratch -> grad_f2 (NCHW): independent of step 4, launched into its tail
        return scatter_finish(scratch, grad_f2, B, C, H, W, true, stream);
    }
    const float* warped = warped_in;
    if (!warped) {      // re-materialise x2_warp (the forward never stored it)
        if (!pwc_warp_forward(f2, flow, wbuf, B, C, H, W, stream)) return 0;
        warped = wbuf;
    }
    if (!corr_backward_impl(grad_out, gate, f1, warped, grad_f1, gwarped, g, slope, stream, 3, gated, &gated_done, gs)) return 0;
    if (vec_ok) return warp_backward_scratch(gwarped, f2, flow, grad_f2, grad_flow, scratch, nullptr, B, C, H, W, stream);
    return pwc_warp_backward(gwarped, f2, flow, grad_f2, grad_flow, B, C, H, W, stream);
}

int pwc_warpcorr_backward(const float* grad_out, const float* f1, const float* f2,
                          const float* flow, const float* out, const float* warped_in, float* grad_f1, float* grad_f2,
                          float* grad_flow, void* workspace, long long workspace_bytes, int B,
                          int C, int H, int W, int pad_size, int kernel_size, int max_displacement,
                          int stride1, int stride2, int act, float slope, cudaStream_t stream)
{
    return warpcorr_backward_impl(grad_out, GradStrides{0, 0}, f1, f2, flow, out, warped_in, grad_f1, grad_f2, grad_flow,
                                  workspace, workspace_bytes, B, C, H, W, pad_size, kernel_size, max_displacement, stride1,
                                  stride2, act, slope, stream);
}

int pwc_warpcorr_backward_strided(const float* grad_out, long long grad_out_batch_stride, const float* f1,
                                  const float* f2, const float* flow, const float* out, long long out_batch_stride,
                                  const float* warped_in, float* grad_f1, float* grad_f2, float* grad_flow,
                                  void* workspace, long long workspace_bytes, int B, int C, int H, int W,
                                  int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                  int act, float slope, cudaStream_t stream)
{
    return warpcorr_backward_impl(grad_out, GradStrides{grad_out_batch_stride, out_batch_stride}, f1, f2, flow, out,
                                  warped_in, grad_f1, grad_f2, grad_flow, workspace, workspace_bytes, B, C, H, W, pad_size,
                                  kernel_size, max_displacement, stride1, stride2, act, slope, stream);
}

int Correlation_forward_cuda_kernel(float* output, int ob, int oc, int oh, int ow, int, int, int,
                                    int, float* input1, int ic, int ih, int iw, int, int, int, int,
                                    float* input2, int gc, int, int, int, int, float*, float*,
                                    int pad_size, int kernel_size, int max_displacement,
                                    int stride1, int stride2, int, cudaStream_t stream)
{
    if (!output || !input1 || !input2) return fail("Correlation_forward_cuda_kernel: null pointer");
    if (gc != ic) return fail("input channel mismatch (%d vs %d)", ic, gc);
    pwc::CorrGeom g;
    if (!make_geom(g, ob, ic, ih, iw, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    if (oc != g.oc || oh != g.oh || ow != g.ow)
        return fail("output is [%d,%d,%d,%d] but the parameters give [%d,%d,%d,%d]", ob, oc, oh, ow,
                    ob, g.oc, g.oh, g.ow);
    return forward_impl(input1, input2, FlowSpec{nullptr, 0, nullptr, nullptr, 0}, output, nullptr, g, 0, 0.0f, 0, stream);
}

int Correlation_backward_cuda_kernel(float* gradOutput, int gob, int goc, int goh, int gow, int,
                                     int, int, int, float* input1, int ic, int ih, int iw, int, int,
                                     int, int, float* input2, int, int, int, int,
                                     float* gradInput1, int, int, int, int, float* gradInput2,
                                     int ggc, int, int, int, int, float*, float*, int pad_size,
                                     int kernel_size, int max_displacement, int stride1,
                                     int stride2, int, cudaStream_t stream)
{
    if (!gradOutput || !input1 || !input2 || !gradInput1 || !gradInput2)
        return fail("Correlation_backward_cuda_kernel: null pointer");
    if (ggc != ic) return fail("gradInput2 channel mismatch (%d vs %d)", ic, ggc);
    pwc::CorrGeom g;
    if (!make_geom(g, gob, ic, ih, iw, pad_size, kernel_size, max_displacement, stride1, stride2)) return 0;
    if (goc != g.oc || goh != g.oh || gow != g.ow)
        return fail("gradOutput is [%d,%d,%d,%d] but the parameters give [%d,%d,%d,%d]", gob, goc,
                    goh, gow, gob, g.oc, g.oh, g.ow);
    return corr_backward_impl(gradOutput, nullptr, input1, input2, gradInput1, gradInput2, g, 0.0f, stream);
}

}  // extern "C"
