/*
 * oracle/pwc_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT A PRODUCT PATH.
 *
 * Plain-C CPU restatement of the reference's warp + cost-volume hot path, used only as the
 * checker in tests/, __graft_entry__.smoke() and as bench.py's cpu_baseline leg.  Nothing in
 * pwc_net_pytorch_b200/ may import, link or call it (the product path fails loudly when the
 * CUDA library is missing; there is no CPU fallback).
 *
 * What each function restates (paths under /root/reference):
 *   pwc_oracle_corr_shape     correlation_package/src/correlation_cuda.c:20-34
 *   pwc_oracle_corr_forward   correlation_package/src/correlation_cuda.c:36-42 (zero padding) +
 *                             correlation_package/src/correlation_cuda_kernel.cu:10-32 (layout copy,
 *                             expressed here as a padded accessor) and :45-101 (forward kernel)
 *   pwc_oracle_corr_backward  correlation_cuda.c:105-121, correlation_cuda_kernel.cu:119-196 (input1)
 *                             and :211-288 (input2)
 *   pwc_oracle_warp_forward   modules.py:31-42 + utils.py:3-7, with F.grid_sample in its torch-0.4.0
 *                             meaning (bilinear, zero padding, align_corners=True; SURVEY.md section 0
 *                             fact 3).  The arithmetic of grid_sample lives in PyTorch
 *                             (torch==0.4.0, requirements.txt:62), not under /root/reference.
 *   pwc_oracle_warp_backward  autograd of the above
 *   pwc_oracle_warpcorr_*     model.py:80-84: warp, correlation, optional leaky_relu_
 *
 * Parity pins (see tests/test_oracle_golden.py): outputs of the reference's own Python modules
 * (modules.WarpingLayer, modules.CostVolumeLayer) imported from /root/reference by
 * tests/golden/make_golden.py, and outputs of the reference's own CUDA kernels compiled unchanged
 * for sm_100a (oracle/build_ref.sh -> oracle/_ref/libref_corr.so) run on a B200 by
 * tests/golden/make_golden_gpu.py.  The reference itself ships no tests or golden vectors.
 *
 * "mode": 0 = exact (double accumulation, sample at x+u directly),
 *         1 = literal (fp32, same summation order as the reference kernels: 32 lane-partials then a
 *             serial add; same normalise/denormalise round trip as WarpingLayer).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LANES 32

static inline long idx4(int n, int c, int y, int x, int C, int H, int W)
{
    return (((long)n * C + c) * H + y) * (long)W + x;
}

/* value of the zero-padded, channels-last copy rInput[n, yp, xp, c]
 * (correlation_cuda_kernel.cu:28-31 after the zero fill of correlation_cuda.c:39-40).
 * Positions outside the padded array itself would be out-of-bounds reads in the reference;
 * they are defined as 0 here. */
static inline float padded(const float *in, int n, int c, int yp, int xp,
                           int C, int H, int W, int pad)
{
    int y = yp - pad, x = xp - pad;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
    return in[idx4(n, c, y, x, C, H, W)];
}

/* correlation_cuda.c:20-34 */
void pwc_oracle_corr_shape(int H, int W, int pad, int k, int md, int s1, int s2,
                           int *oc, int *oh, int *ow)
{
    int kr = (k - 1) / 2;
    int border = kr + md;
    int pH = H + 2 * pad, pW = W + 2 * pad;
    int r = md / s2;
    *oc = (2 * r + 1) * (2 * r + 1);
    *oh = (int)ceilf((float)(pH - 2 * border) / (float)s1);
    *ow = (int)ceilf((float)(pW - 2 * border) / (float)s1);
}

/* correlation_cuda_kernel.cu:45-101.  out is [B, D*D, oh, ow]. */
int pwc_oracle_corr_forward(const float *in1, const float *in2, float *out,
                            int B, int C, int H, int W,
                            int pad, int k, int md, int s1, int s2, int mode)
{
    int oc, oh, ow;
    pwc_oracle_corr_shape(H, W, pad, k, md, s1, s2, &oc, &oh, &ow);
    if (oh <= 0 || ow <= 0) return 0;
    const int kr = (k - 1) / 2;
    const int r = md / s2;
    const int D = 2 * r + 1;
    const float nelems = (float)(k * k * C);

#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n) {
        for (int by = 0; by < oh; ++by) {
            for (int bx = 0; bx < ow; ++bx) {
                int y1 = by * s1 + md + kr;
                int x1 = bx * s1 + md + kr;
                for (int tj = -r; tj <= r; ++tj) {
                    for (int ti = -r; ti <= r; ++ti) {
                        int x2 = x1 + ti * s2;
                        int y2 = y1 + tj * s2;
                        float res;
                        if (mode == 1) {
                            float lane[LANES];
                            for (int l = 0; l < LANES; ++l) lane[l] = 0.0f;
                            for (int j = -kr; j <= kr; ++j)
                                for (int i = -kr; i <= kr; ++i)
                                    for (int l = 0; l < LANES; ++l)
                                        for (int ch = l; ch < C; ch += LANES)
                                            lane[l] += padded(in1, n, ch, y1 + j, x1 + i, C, H, W, pad) *
                                                       padded(in2, n, ch, y2 + j, x2 + i, C, H, W, pad);
                            float s = 0.0f;
                            for (int l = 0; l < LANES; ++l) s += lane[l];
                            res = s / nelems;
                        } else {
                            double s = 0.0;
                            for (int j = -kr; j <= kr; ++j)
                                for (int i = -kr; i <= kr; ++i)
                                    for (int ch = 0; ch < C; ++ch)
                                        s += (double)padded(in1, n, ch, y1 + j, x1 + i, C, H, W, pad) *
                                             (double)padded(in2, n, ch, y2 + j, x2 + i, C, H, W, pad);
                            res = (float)(s / (double)nelems);
                        }
                        int tc = (tj + r) * D + (ti + r);
                        out[idx4(n, tc, by, bx, oc, oh, ow)] = res;
                    }
                }
            }
        }
    }
    return 1;
}

/* correlation_cuda_kernel.cu:119-196 and :211-288.  The reference launches one block per input
 * pixel with y = blockIdx.x*stride1 + pad (:121-122), which only addresses the input correctly
 * for stride1 == 1; for stride1 > 1 it writes outside gradInput.  The oracle therefore refuses
 * stride1 != 1 (returns 0).  g1, g2 are [B, C, H, W] and fully overwritten. */
int pwc_oracle_corr_backward(const float *gout, const float *in1, const float *in2,
                             float *g1, float *g2,
                             int B, int C, int H, int W,
                             int pad, int k, int md, int s1, int s2, int mode)
{
    if (s1 != 1) return 0;
    int oc, oh, ow;
    pwc_oracle_corr_shape(H, W, pad, k, md, s1, s2, &oc, &oh, &ow);
    if (oh <= 0 || ow <= 0) return 0;
    const int kr = (k - 1) / 2;
    const int r = md / s2;
    const int D = 2 * r + 1;
    const float nelems = (float)(k * k * C);
    memset(g1, 0, sizeof(float) * (size_t)B * C * H * W);   /* correlation_cuda.c:120-121 */
    memset(g2, 0, sizeof(float) * (size_t)B * C * H * W);

#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n) {
        for (int c = 0; c < C; ++c) {
            for (int yy = 0; yy < H; ++yy) {
                for (int xx = 0; xx < W; ++xx) {
                    const int y = yy * s1 + pad, x = xx * s1 + pad;
                    /* ---- gradInput1 (:129-196) ---- */
                    {
                        int xmin = (x - kr - md) / s1, ymin = (y - kr - md) / s1;
                        int xmax = (x + kr - md) / s1, ymax = (y + kr - md) / s1;
                        int skip = (xmax < 0 || ymax < 0 || xmin >= ow || ymin >= oh) ||
                                   (xmin > xmax || ymin > ymax);
                        if (!skip) {
                            if (xmin < 0) xmin = 0;
                            if (xmax > ow - 1) xmax = ow - 1;
                            if (ymin < 0) ymin = 0;
                            if (ymax > oh - 1) ymax = oh - 1;
                            float lane[LANES];
                            double acc = 0.0;
                            for (int l = 0; l < LANES; ++l) lane[l] = 0.0f;
                            for (int tc = 0; tc < oc; ++tc) {
                                int i2 = (tc % D - r) * s2, j2 = (tc / D - r) * s2;
                                float v2 = padded(in2, n, c, y + j2, x + i2, C, H, W, pad);
                                for (int j = ymin; j <= ymax; ++j)
                                    for (int i = xmin; i <= xmax; ++i) {
                                        float g = gout[idx4(n, tc, j, i, oc, oh, ow)];
                                        lane[tc % LANES] += g * v2;
                                        acc += (double)g * (double)v2;
                                    }
                            }
                            float res;
                            if (mode == 1) {
                                float s = 0.0f;
                                for (int l = 0; l < LANES; ++l) s += lane[l];
                                res = s / nelems;
                            } else {
                                res = (float)(acc / (double)nelems);
                            }
                            g1[idx4(n, c, yy, xx, C, H, W)] = res;
                        }
                    }
                    /* ---- gradInput2 (:236-288) ---- */
                    {
                        float lane[LANES];
                        double acc = 0.0;
                        for (int l = 0; l < LANES; ++l) lane[l] = 0.0f;
                        for (int tc = 0; tc < oc; ++tc) {
                            int i2 = (tc % D - r) * s2, j2 = (tc / D - r) * s2;
                            int xmin = (x - kr - md - i2) / s1, ymin = (y - kr - md - j2) / s1;
                            int xmax = (x + kr - md - i2) / s1, ymax = (y + kr - md - j2) / s1;
                            if (xmax < 0 || ymax < 0 || xmin >= ow || ymin >= oh) continue;
                            if (xmin > xmax || ymin > ymax) continue;
                            if (xmin < 0) xmin = 0;
                            if (xmax > ow - 1) xmax = ow - 1;
                            if (ymin < 0) ymin = 0;
                            if (ymax > oh - 1) ymax = oh - 1;
                            float v1 = padded(in1, n, c, y - j2, x - i2, C, H, W, pad);
                            for (int j = ymin; j <= ymax; ++j)
                                for (int i = xmin; i <= xmax; ++i) {
                                    float g = gout[idx4(n, tc, j, i, oc, oh, ow)];
                                    lane[tc % LANES] += g * v1;
                                    acc += (double)g * (double)v1;
                                }
                        }
                        float res;
                        if (mode == 1) {
                            float s = 0.0f;
                            for (int l = 0; l < LANES; ++l) s += lane[l];
                            res = s / nelems;
                        } else {
                            res = (float)(acc / (double)nelems);
                        }
                        g2[idx4(n, c, yy, xx, C, H, W)] = res;
                    }
                }
            }
        }
    }
    return 1;
}

/* Sample position of output pixel (y, x) in source-pixel units.
 * literal: the fp32 round trip of modules.py:36-40 (flow / ((W-1)/2) added to linspace(-1,1,W))
 *          followed by grid_sample's align_corners=True un-normalisation ((g+1)/2*(W-1)).
 * exact  : x + u in double. */
static inline void sample_pos(float u, float v, int y, int x, int H, int W, int mode,
                              double *sx, double *sy)
{
    if (mode == 1) {
        float hx = (float)(((double)W - 1.0) / 2.0), hy = (float)(((double)H - 1.0) / 2.0);
        float gx0 = -1.0f + (2.0f / (float)(W - 1)) * (float)x;   /* linspace(-1,1,W)[x], utils.py:4 */
        float gy0 = -1.0f + (2.0f / (float)(H - 1)) * (float)y;   /* utils.py:5 */
        float gx = gx0 + u / hx, gy = gy0 + v / hy;               /* modules.py:37-40 */
        float ix = ((gx + 1.0f) / 2.0f) * (float)(W - 1);
        float iy = ((gy + 1.0f) / 2.0f) * (float)(H - 1);
        *sx = (double)ix; *sy = (double)iy;
    } else {
        *sx = (double)x + (double)u; *sy = (double)y + (double)v;
    }
}

/* modules.py:31-42.  x:[B,C,H,W], flow:[B,2,H,W] (ch0 = horizontal px, ch1 = vertical px). */
int pwc_oracle_warp_forward(const float *x, const float *flow, float *out,
                            int B, int C, int H, int W, int mode)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < B; ++n) {
        for (int yy = 0; yy < H; ++yy) {
            for (int xx = 0; xx < W; ++xx) {
                float u = flow[idx4(n, 0, yy, xx, 2, H, W)];
                float v = flow[idx4(n, 1, yy, xx, 2, H, W)];
                double sx, sy;
                sample_pos(u, v, yy, xx, H, W, mode, &sx, &sy);
                double fx0 = floor(sx), fy0 = floor(sy);
                double ax = sx - fx0, ay = sy - fy0;
                double w00 = (1.0 - ax) * (1.0 - ay), w01 = ax * (1.0 - ay);
                double w10 = (1.0 - ax) * ay, w11 = ax * ay;
                /* guard the int conversion against huge / non-finite flows */
                int valid = (sx > -2.0 && sx < (double)W + 1.0 && sy > -2.0 && sy < (double)H + 1.0);
                int x0 = valid ? (int)fx0 : -4, y0 = valid ? (int)fy0 : -4;
                int x1 = x0 + 1, y1 = y0 + 1;
                int in00 = (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H);
                int in01 = (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H);
                int in10 = (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H);
                int in11 = (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H);
                for (int c = 0; c < C; ++c) {
                    double a = 0.0;
                    if (in00) a += w00 * (double)x[idx4(n, c, y0, x0, C, H, W)];
                    if (in01) a += w01 * (double)x[idx4(n, c, y0, x1, C, H, W)];
                    if (in10) a += w10 * (double)x[idx4(n, c, y1, x0, C, H, W)];
                    if (in11) a += w11 * (double)x[idx4(n, c, y1, x1, C, H, W)];
                    out[idx4(n, c, yy, xx, C, H, W)] = (float)a;
                }
            }
        }
    }
    return 1;
}

/* autograd of the warp: gx (scatter-add into the 4 corners) and gflow.  The (W-1)/2 factors of
 * modules.py:37-38 and of grid_sample's un-normalisation cancel. */
int pwc_oracle_warp_backward(const float *gout, const float *x, const float *flow,
                             float *gx, float *gflow, int B, int C, int H, int W)
{
    double *acc = (double *)calloc((size_t)B * C * H * W, sizeof(double));
    if (!acc) return 0;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < B; ++n) {
        for (int yy = 0; yy < H; ++yy) {
            for (int xx = 0; xx < W; ++xx) {
                float u = flow[idx4(n, 0, yy, xx, 2, H, W)];
                float v = flow[idx4(n, 1, yy, xx, 2, H, W)];
                double sx = (double)xx + (double)u, sy = (double)yy + (double)v;
                double fx0 = floor(sx), fy0 = floor(sy);
                double ax = sx - fx0, ay = sy - fy0;
                int valid = (sx > -2.0 && sx < (double)W + 1.0 && sy > -2.0 && sy < (double)H + 1.0);
                int x0 = valid ? (int)fx0 : -4, y0 = valid ? (int)fy0 : -4;
                int x1 = x0 + 1, y1 = y0 + 1;
                int in00 = (x0 >= 0 && x0 < W && y0 >= 0 && y0 < H);
                int in01 = (x1 >= 0 && x1 < W && y0 >= 0 && y0 < H);
                int in10 = (x0 >= 0 && x0 < W && y1 >= 0 && y1 < H);
                int in11 = (x1 >= 0 && x1 < W && y1 >= 0 && y1 < H);
                double gu = 0.0, gv = 0.0;
                for (int c = 0; c < C; ++c) {
                    double g = (double)gout[idx4(n, c, yy, xx, C, H, W)];
                    double v00 = in00 ? (double)x[idx4(n, c, y0, x0, C, H, W)] : 0.0;
                    double v01 = in01 ? (double)x[idx4(n, c, y0, x1, C, H, W)] : 0.0;
                    double v10 = in10 ? (double)x[idx4(n, c, y1, x0, C, H, W)] : 0.0;
                    double v11 = in11 ? (double)x[idx4(n, c, y1, x1, C, H, W)] : 0.0;
                    if (in00) acc[idx4(n, c, y0, x0, C, H, W)] += g * (1.0 - ax) * (1.0 - ay);
                    if (in01) acc[idx4(n, c, y0, x1, C, H, W)] += g * ax * (1.0 - ay);
                    if (in10) acc[idx4(n, c, y1, x0, C, H, W)] += g * (1.0 - ax) * ay;
                    if (in11) acc[idx4(n, c, y1, x1, C, H, W)] += g * ax * ay;
                    gu += g * ((v01 - v00) * (1.0 - ay) + (v11 - v10) * ay);
                    gv += g * ((v10 - v00) * (1.0 - ax) + (v11 - v01) * ax);
                }
                gflow[idx4(n, 0, yy, xx, 2, H, W)] = (float)gu;
                gflow[idx4(n, 1, yy, xx, 2, H, W)] = (float)gv;
            }
        }
    }
    long total = (long)B * C * H * W;
    for (long i = 0; i < total; ++i) gx[i] = (float)acc[i];
    free(acc);
    return 1;
}

/* model.py:80-84: x2_warp = warp(x2, flow); corr = Correlation(x1, x2_warp); optional
 * leaky_relu_(corr) with slope `slope` when act != 0.  flow == NULL means no warp (x2 used as is).
 * warped_out (may be NULL) receives x2_warp. */
int pwc_oracle_warpcorr_forward(const float *f1, const float *f2, const float *flow,
                                float *out, float *warped_out,
                                int B, int C, int H, int W,
                                int pad, int k, int md, int s1, int s2,
                                int act, float slope, int mode)
{
    const float *second = f2;
    float *tmp = NULL;
    if (flow) {
        tmp = warped_out ? warped_out : (float *)malloc(sizeof(float) * (size_t)B * C * H * W);
        if (!tmp) return 0;
        pwc_oracle_warp_forward(f2, flow, tmp, B, C, H, W, mode);
        second = tmp;
    } else if (warped_out) {
        memcpy(warped_out, f2, sizeof(float) * (size_t)B * C * H * W);
    }
    int ok = pwc_oracle_corr_forward(f1, second, out, B, C, H, W, pad, k, md, s1, s2, mode);
    if (ok && act) {
        int oc, oh, ow;
        pwc_oracle_corr_shape(H, W, pad, k, md, s1, s2, &oc, &oh, &ow);
        long total = (long)B * oc * oh * ow;
        for (long i = 0; i < total; ++i)
            if (out[i] < 0.0f) out[i] *= slope;
    }
    if (tmp && tmp != warped_out) free(tmp);
    return ok;
}

/* backward of the above.  `out` (the forward result) is only read when act != 0 (to gate the
 * gradient by the sign, as leaky_relu_'s backward does).  gflow may be NULL when flow is NULL. */
int pwc_oracle_warpcorr_backward(const float *gout, const float *f1, const float *f2,
                                 const float *flow, const float *out,
                                 float *g1, float *g2, float *gflow,
                                 int B, int C, int H, int W,
                                 int pad, int k, int md, int s1, int s2,
                                 int act, float slope)
{
    int oc, oh, ow;
    pwc_oracle_corr_shape(H, W, pad, k, md, s1, s2, &oc, &oh, &ow);
    long nout = (long)B * oc * oh * ow, nin = (long)B * C * H * W;
    float *g = (float *)malloc(sizeof(float) * (size_t)nout);
    float *warped = NULL, *gw = NULL;
    int ok = 0;
    if (!g) return 0;
    for (long i = 0; i < nout; ++i)
        g[i] = (act && out[i] < 0.0f) ? gout[i] * slope : gout[i];
    if (flow) {
        warped = (float *)malloc(sizeof(float) * (size_t)nin);
        gw = (float *)malloc(sizeof(float) * (size_t)nin);
        if (!warped || !gw) goto done;
        pwc_oracle_warp_forward(f2, flow, warped, B, C, H, W, 0);
        ok = pwc_oracle_corr_backward(g, f1, warped, g1, gw, B, C, H, W, pad, k, md, s1, s2, 0);
        if (ok) ok = pwc_oracle_warp_backward(gw, f2, flow, g2, gflow, B, C, H, W);
    } else {
        ok = pwc_oracle_corr_backward(g, f1, f2, g1, g2, B, C, H, W, pad, k, md, s1, s2, 0);
    }
done:
    free(g); free(warped); free(gw);
    return ok;
}
