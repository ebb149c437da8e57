/*
 * pwc_b200.h -- C ABI of libpwc_b200.so: the B200 (sm_100a) warp + cost-volume hot path of PWC-Net.
 *
 * Drop-in boundary for daigo0927/PWC-Net_pytorch (paths below are under the reference root):
 *
 *   - Correlation_forward_cuda_kernel / Correlation_backward_cuda_kernel keep the exact symbol
 *     names and argument lists of correlation_package/src/correlation_cuda_kernel.h:5-39 and
 *     :41-88, so the reference's wrapper correlation_package/src/correlation_cuda.c:44-81 /
 *     :124-171 (or any ctypes/cffi binding of those launchers) links against this library
 *     unchanged.
 *   - pwc_* are the entry points a binding of the fused path adds: they replace the call
 *     sequence modules.py:31-42 (WarpingLayer) -> correlation_cuda.c:11-93 -> model.py:84
 *     (leaky_relu_) and its autograd (functions/correlation.py:38-56 + grid_sample backward).
 *
 * Conventions (same as the reference launchers unless stated):
 *   - fp32, dense NCHW-contiguous device buffers (functions/correlation.py:17-18); stride
 *     arguments of the two legacy entry points are accepted and, as in the reference
 *     (correlation_cuda_kernel.cu:296-369 never reads them), ignored.
 *   - The caller owns every buffer.  The library never allocates, frees or synchronises; work is
 *     enqueued on `stream` and every entry point is CUDA-graph capturable.
 *   - Outputs are fully overwritten; no pre-zeroing is required of the caller.
 *     The legacy scratch pointers rInput1/rInput2 are not used and may be NULL.
 *   - Return value: 1 on success, 0 on failure (correlation_cuda_kernel.cu:362-368).  After a 0,
 *     pwc_last_error() returns a thread-local message.
 *   - corr_type_multiply is accepted and ignored: the reference never reads it
 *     (SURVEY.md section 0 fact 6).
 *
 * Semantics (r = max_displacement / stride2, D = 2r+1, kr = (kernel_size-1)/2):
 *   out[n, (tj+r)*D + (ti+r), y, x] =
 *       1/(k*k*C) * sum_{j,i in [-kr,kr]} sum_c  P1[n,c,y1+j,x1+i] * P2[n,c,y1+j+tj*s2, x1+i+ti*s2]
 *   with y1 = y*stride1 + max_displacement + kr and P* the inputs zero-padded by pad_size
 *   (correlation_cuda_kernel.cu:52-101); output height = ceil((H + 2*pad - 2*(kr+md)) / stride1)
 *   (correlation_cuda.c:25-34).
 *   The warp samples input2 bilinearly with zero padding at (x + flow[n,0,y,x], y + flow[n,1,y,x])
 *   (modules.py:36-41 under torch 0.4.0 grid_sample == align_corners=True).
 */
#ifndef PWC_B200_H
#define PWC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st *cudaStream_t;
#endif

#define PWC_B200_ABI_VERSION 5

/* ---- legacy launchers: correlation_cuda_kernel.h:5-39 ------------------------------------- */
int Correlation_forward_cuda_kernel(
    float *output, int ob, int oc, int oh, int ow, int osb, int osc, int osh, int osw,
    float *input1, int ic, int ih, int iw, int isb, int isc, int ish, int isw,
    float *input2, int gc, int gsb, int gsc, int gsh, int gsw,
    float *rInput1, float *rInput2,
    int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
    int corr_type_multiply, cudaStream_t stream);

/* ---- correlation_cuda_kernel.h:41-88.  stride1 must be 1 (the reference's backward kernels
 * address gradInput out of range otherwise, correlation_cuda_kernel.cu:121-122,196). -------- */
int Correlation_backward_cuda_kernel(
    float *gradOutput, int gob, int goc, int goh, int gow, int gosb, int gosc, int gosh, int gosw,
    float *input1, int ic, int ih, int iw, int isb, int isc, int ish, int isw,
    float *input2, int gsb, int gsc, int gsh, int gsw,
    float *gradInput1, int gisb, int gisc, int gish, int gisw,
    float *gradInput2, int ggc, int ggsb, int ggsc, int ggsh, int ggsw,
    float *rInput1, float *rInput2,
    int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
    int corr_type_multiply, cudaStream_t stream);

/* ---- output geometry: correlation_cuda.c:20-34 --------------------------------------------- */
int pwc_corr_output_shape(int H, int W, int pad_size, int kernel_size, int max_displacement,
                          int stride1, int stride2, int *out_channels, int *out_h, int *out_w);

/* ---- WarpingLayer.forward, modules.py:31-42.  x,out:[B,C,H,W]  flow:[B,2,H,W] --------------- */
int pwc_warp_forward(const float *x, const float *flow, float *out,
                     int B, int C, int H, int W, cudaStream_t stream);

/* ---- autograd of WarpingLayer (grid_sample backward + modules.py:36-40).
 * grad_x:[B,C,H,W] and grad_flow:[B,2,H,W] are fully overwritten (grad_x is zeroed on `stream`
 * by the library before the scatter-add).  Either may be NULL to skip it. -------------------- */
int pwc_warp_backward(const float *grad_out, const float *x, const float *flow,
                      float *grad_x, float *grad_flow,
                      int B, int C, int H, int W, cudaStream_t stream);

/* ---- same with a caller-owned scratch (ABI v5): the tiled kernel of the fused backward -- corner values from a
 * TMA-staged shared-memory window, 128-bit vector reductions into an 8-channel-interleaved scratch, flow gradient
 * without atomics -- instead of one scalar atomic per corner and channel (2.3x faster at the level-2 shape).
 * workspace: pwc_warp_backward_workspace(...) bytes, 16-byte aligned, contents undefined on entry and exit; with a
 * smaller / NULL workspace, or when only one of the two gradients is requested, this is pwc_warp_backward. ------- */
long long pwc_warp_backward_workspace(int B, int C, int H, int W);
int pwc_warp_backward_ws(const float *grad_out, const float *x, const float *flow,
                         float *grad_x, float *grad_flow,
                         int B, int C, int H, int W,
                         void *workspace, long long workspace_bytes, cudaStream_t stream);

/* ---- fused warp + correlation + (optional) LeakyReLU: model.py:80-84 in one launch.
 * flow == NULL      : no warp (plain Correlation of f1 with f2).
 * warped_out != NULL: also writes x2_warp [B,C,H,W] (model.py:107,113 exports it).
 * act != 0          : out = leaky_relu(out, slope) (model.py:84 uses slope 0.01). ------------ */
int pwc_warpcorr_forward(const float *f1, const float *f2, const float *flow,
                         float *out, float *warped_out,
                         int B, int C, int H, int W,
                         int pad_size, int kernel_size, int max_displacement,
                         int stride1, int stride2,
                         int act, float slope, cudaStream_t stream);

/* ---- same, writing into a larger buffer: image n's [D*D, oh, ow] block starts at
 * out + n * out_batch_stride (floats; >= D*D*oh*ow, 0 means dense).  This is how the cost volume is
 * written straight into the flow estimator's concatenated input [x1 | corr | flow] (model.py:89-91)
 * instead of being copied there by torch.cat. ------------------------------------------------- */
int pwc_warpcorr_forward_strided(const float *f1, const float *f2, const float *flow,
                                 float *out, long long out_batch_stride, float *warped_out,
                                 int B, int C, int H, int W,
                                 int pad_size, int kernel_size, int max_displacement,
                                 int stride1, int stride2,
                                 int act, float slope, cudaStream_t stream);

/* ---- same, with model.py:78 folded in: the flow is given at the previous (coarser) pyramid level,
 * coarse_flow:[B,2,H/2,W/2] dense (H, W even), and the kernel itself evaluates
 *     flow = F.upsample(coarse_flow, scale_factor=2, mode='bilinear') * 2        (align_corners=False)
 * bit for bit as torch does, warps input2 by it, and writes it to flow_out (required: the flow estimator
 * takes the fine flow as an input, model.py:89-91): image n's [2,H,W] block at flow_out +
 * n * flow_out_batch_stride floats (0 = dense) -- typically the last two channels of the estimator's
 * concatenated input, next to the cost volume written through out / out_batch_stride.
 * Replaces one F.upsample launch, one multiply launch and a [B,2,H,W] round trip per pyramid level.
 * Forward only (inference); training keeps the flow as a tensor and uses pwc_warpcorr_forward. ------- */
int pwc_warpcorr_forward_coarse(const float *f1, const float *f2, const float *coarse_flow,
                                float *out, long long out_batch_stride,
                                float *flow_out, long long flow_out_batch_stride, float *warped_out,
                                int B, int C, int H, int W,
                                int pad_size, int kernel_size, int max_displacement,
                                int stride1, int stride2,
                                int act, float slope, cudaStream_t stream);

/* ---- backward of pwc_warpcorr_forward.
 * out        : the forward result, read only when act != 0 (sign gate of leaky_relu_).
 * warped     : optional x2_warp as written by pwc_warpcorr_forward(warped_out) for the same inputs;
 *              when given, the backward does not re-evaluate the warp (NULL: it does).
 * workspace  : device scratch of pwc_warpcorr_backward_workspace(...) bytes (may be NULL if 0).
 * grad_f1, grad_f2 : [B,C,H,W]; grad_flow: [B,2,H,W] (NULL allowed iff flow == NULL).
 * stride1 must be 1. ------------------------------------------------------------------------- */
long long pwc_warpcorr_backward_workspace(int B, int C, int H, int W, int has_flow,
                                          int pad_size, int kernel_size, int max_displacement,
                                          int stride1, int stride2);
int pwc_warpcorr_backward(const float *grad_out, const float *f1, const float *f2,
                          const float *flow, const float *out, const float *warped,
                          float *grad_f1, float *grad_f2, float *grad_flow,
                          void *workspace, long long workspace_bytes,
                          int B, int C, int H, int W,
                          int pad_size, int kernel_size, int max_displacement,
                          int stride1, int stride2,
                          int act, float slope, cudaStream_t stream);

/* ---- same, reading the output gradient (and, for act != 0, the forward output) through a batch stride:
 * image n's [D*D,H,W] block of grad_out starts at grad_out + n * grad_out_batch_stride floats, that of out
 * at out + n * out_batch_stride (0 = dense).  This is how the gradient of the flow estimator's concatenated
 * input [x1 | corr | flow] (model.py:89-91) is consumed in place, without a slice copy, when the forward
 * wrote the cost volume there with pwc_warpcorr_forward_strided. ------------------------------------------ */
int pwc_warpcorr_backward_strided(const float *grad_out, long long grad_out_batch_stride,
                                  const float *f1, const float *f2, const float *flow,
                                  const float *out, long long out_batch_stride, const float *warped,
                                  float *grad_f1, float *grad_f2, float *grad_flow,
                                  void *workspace, long long workspace_bytes,
                                  int B, int C, int H, int W,
                                  int pad_size, int kernel_size, int max_displacement,
                                  int stride1, int stride2,
                                  int act, float slope, cudaStream_t stream);

/* ---- diagnostics --------------------------------------------------------------------------- */
const char *pwc_last_error(void);          /* thread-local, never NULL */
int pwc_abi_version(void);                 /* == PWC_B200_ABI_VERSION */
/* number of kernels launched by this library in this process since load (bench.py gpu_launches) */
long long pwc_launch_count(void);
/* forces the generic (any kernel_size/stride) kernels even where a tiled fast path exists;
 * test hook, returns the previous value. */
int pwc_set_force_generic(int on);
/* disables the TMA-staged forward kernel (falls back to the plain tiled kernel); test hook,
 * returns the previous value. */
int pwc_set_disable_tma(int on);
/* disables the whole-image kernels of the coarse pyramid levels (H*W <= 256), so that the tiled
 * kernels serve those shapes too; test hook, returns the previous value. */
int pwc_set_disable_small(int on);
/* disables the complete-output kernel of the gradient w.r.t. input1 (corr_bwd_seq_kernel), so that the
 * slice/reduce kernel serves stride2 == 1 too; test hook, returns the previous value. */
int pwc_set_disable_seq(int on);

#ifdef __cplusplus
}
#endif
#endif /* PWC_B200_H */
