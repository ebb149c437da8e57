"""Closed-form PyTorch restatement of the hot path -- TEST INFRASTRUCTURE ONLY.

Runs on any device and dtype (fp64 for gradcheck / error floors).  Two groups:

* `corr_ref`, `warp_ref`, `warpcorr_ref`: the numeric oracle for the CUDA `Correlation` semantics
  (correlation_package/src/correlation_cuda_kernel.cu:45-101 forward; autograd of the closed form
  equals :119-196 / :211-288 for kernel_size=1, stride1=1) and for `WarpingLayer`
  (modules.py:31-42 with torch-0.4.0 grid_sample == align_corners=True).
* `warping_layer_port`, `cost_volume_layer_port`: a port of the reference's PyTorch-level path
  (modules.py:25-42 and :45-74).  CostVolumeLayer computes a stride-1, /81, differently ordered
  volume, so it is the CPU *timing* baseline only; `COSTVOLUME_PERM`
  maps its channels onto Correlation(pad=4, md=4, stride2=1).
"""
import math

import torch
import torch.nn.functional as F


def corr_out_shape(H, W, pad, k, md, s1, s2):
    """correlation_cuda.c:20-34."""
    kr = (k - 1) // 2
    border = kr + md
    r = md // s2
    D = 2 * r + 1
    oh = math.ceil((H + 2 * pad - 2 * border) / s1)
    ow = math.ceil((W + 2 * pad - 2 * border) / s1)
    return D * D, oh, ow


def corr_ref(f1, f2, pad, k, md, s1=1, s2=1):
    """out[n,(tj+r)D+(ti+r),y,x] = 1/(k*k*C) sum_{j,i,c} p1[n,c,y1+j,x1+i] p2[n,c,y1+j+tj*s2,x1+i+ti*s2],
    y1 = y*s1 + md + kr, p = zero-padded by `pad` (correlation_cuda_kernel.cu:52-101)."""
    B, C, H, W = f1.shape
    kr = (k - 1) // 2
    r = md // s2
    D = 2 * r + 1
    oc, oh, ow = corr_out_shape(H, W, pad, k, md, s1, s2)
    R = r * s2
    # extra margin so every shifted slice is in range even when pad < R + kr
    extra = max(0, R + kr - pad) + md + kr
    p1 = F.pad(f1, (pad + extra,) * 4)
    p2 = F.pad(f2, (pad + extra,) * 4)
    y0 = md + kr + extra
    outs = []
    for tj in range(-r, r + 1):
        for ti in range(-r, r + 1):
            acc = 0
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    a = p1[:, :, y0 + j: y0 + j + (oh - 1) * s1 + 1: s1,
                           y0 + i: y0 + i + (ow - 1) * s1 + 1: s1]
                    b = p2[:, :, y0 + j + tj * s2: y0 + j + tj * s2 + (oh - 1) * s1 + 1: s1,
                           y0 + i + ti * s2: y0 + i + ti * s2 + (ow - 1) * s1 + 1: s1]
                    acc = acc + (a * b).sum(1)
            outs.append(acc)
    return torch.stack(outs, 1) / float(k * k * C)


def warp_ref(x, flow):
    """Bilinear, zero-padded sample of x at (x+u, y+v): the torch-0.4.0 meaning of
    modules.py:36-41.  Written as an explicit gather so it is exact in any dtype."""
    B, C, H, W = x.shape
    dev, dt = x.device, x.dtype
    xs = torch.arange(W, device=dev, dtype=dt).view(1, 1, W) + flow[:, 0]
    ys = torch.arange(H, device=dev, dtype=dt).view(1, H, 1) + flow[:, 1]
    x0f, y0f = torch.floor(xs), torch.floor(ys)
    ax, ay = xs - x0f, ys - y0f
    big = 4.0 * max(H, W) + 8.0
    x0 = x0f.clamp(-big, big).long()
    y0 = y0f.clamp(-big, big).long()
    flat = x.reshape(B, C, H * W)
    out = 0
    for dy, dx, wgt in ((0, 0, (1 - ax) * (1 - ay)), (0, 1, ax * (1 - ay)),
                        (1, 0, (1 - ax) * ay), (1, 1, ax * ay)):
        xi, yi = x0 + dx, y0 + dy
        ok = ((xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)).to(dt)
        lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).reshape(B, 1, H * W).expand(B, C, H * W)
        val = torch.gather(flat, 2, lin).reshape(B, C, H, W)
        out = out + val * (wgt * ok).unsqueeze(1)
    return out


def warpcorr_ref(f1, f2, flow, pad, k, md, s1=1, s2=1, act=False, slope=0.01):
    """model.py:80-84."""
    second = f2 if flow is None else warp_ref(f2, flow)
    out = corr_ref(f1, second, pad, k, md, s1, s2)
    if act:
        out = F.leaky_relu(out, slope)
    return out


# ----------------------------------------------------------------------------------------------
# Port of the reference's PyTorch-level path (CPU timing baseline, bench.py cpu_baseline / --impl
# reference).  Same operation sequence as modules.py, written independently.
# ----------------------------------------------------------------------------------------------

def grid_port(x):
    """utils.py:3-7: a [B,2,H,W] fp32 grid of linspace(-1,1) coordinates, built on the CPU."""
    B, _, H, W = x.shape
    gx = torch.linspace(-1.0, 1.0, W).view(1, 1, 1, W).expand(B, 1, H, W)
    gy = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat([gx, gy], 1)


def warping_layer_port(x, flow):
    """modules.py:31-42 with the torch-0.4.0 grid_sample semantics spelled out."""
    B, _, H, W = flow.shape
    norm = torch.zeros_like(flow)
    norm[:, 0] = flow[:, 0] / ((W - 1.0) / 2.0)
    norm[:, 1] = flow[:, 1] / ((H - 1.0) / 2.0)
    grid = (grid_port(x).to(x.device) + norm).permute(0, 2, 3, 1)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def _costvolume_order(search_range):
    """(dy, dx) of every CostVolumeLayer channel, in its own order (modules.py:56-72).
    Channel for shift (a, b) multiplies tgt[y-a, x-b] with src[y, x], i.e. displacement (-a, -b)."""
    order = [(0, 0)]
    for i in range(1, search_range + 1):
        order += [(-i, 0), (i, 0), (0, -i), (0, i)]
        for j in range(1, search_range + 1):
            order += [(-i, -j), (i, j), (-i, j), (i, -j)]
    return order


def costvolume_perm(search_range=4):
    """Raster channel tc=(dy+r)*D+(dx+r) of Correlation(pad=r, md=r, stride2=1) for each
    CostVolumeLayer channel."""
    D = 2 * search_range + 1
    return [(dy + search_range) * D + (dx + search_range) for dy, dx in _costvolume_order(search_range)]


COSTVOLUME_PERM = costvolume_perm(4)


def cost_volume_layer_port(src, tgt, search_range=4):
    """modules.py:52-74: 81 shifted channel-sum slices, divided by 81."""
    B, C, H, W = src.shape
    order = _costvolume_order(search_range)
    out = torch.zeros((B, len(order), H, W), dtype=src.dtype, device=src.device)
    for I, (dy, dx) in enumerate(order):
        # out[y, x] = sum_c src[y, x] * tgt[y+dy, x+dx] where both are in range
        ys0, ys1 = max(0, -dy), min(H, H - dy)
        xs0, xs1 = max(0, -dx), min(W, W - dx)
        if ys1 <= ys0 or xs1 <= xs0:      # shift larger than the image: the slice stays zero
            continue
        out[:, I, ys0:ys1, xs0:xs1] = (src[:, :, ys0:ys1, xs0:xs1] *
                                       tgt[:, :, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]).sum(1)
    return out / float(len(order))
