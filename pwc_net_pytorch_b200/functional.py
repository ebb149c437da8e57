"""autograd Functions over the C ABI (PyTorch is plumbing here: device memory, streams, autograd).

  CorrelationFunction      drop-in for correlation_package/functions/correlation.py:5-56
  WarpFunction             WarpingLayer's arithmetic, modules.py:31-42 (+ its autograd)
  WarpCorrelationFunction  model.py:80-84 fused: warp -> correlation -> optional leaky_relu_

All tensors must be fp32 CUDA; inputs are made contiguous (the reference asserts contiguity for
the inputs, functions/correlation.py:17-18, and silently assumes it for grad_output, :40-41).
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager is a
    measurable part of the launch overhead of the small pyramid levels)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_inputs(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("pwc_net_pytorch_b200 ops are CUDA-only (sm_100a); there is no CPU path")
        if t.dtype != torch.float32:
            raise TypeError(f"expected float32, got {t.dtype} (the reference operator is fp32-only)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("all tensors must be on the same device")
    return dev


_shape_cache = {}


def corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2):
    """correlation_cuda.c:20-34 (answered by the library, memoised: it is on every call's path)."""
    key = (H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    hit = _shape_cache.get(key)
    if hit is not None:
        return hit
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    ok = _lib.load().pwc_corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1,
                                           stride2, ctypes.byref(oc), ctypes.byref(oh),
                                           ctypes.byref(ow))
    _lib.check(ok, "pwc_corr_output_shape")
    if len(_shape_cache) < 4096:
        _shape_cache[key] = (oc.value, oh.value, ow.value)
    return oc.value, oh.value, ow.value


def _dense_strides(t):
    return [int(s) for s in t.stride()]


class CorrelationFunction(Function):
    """Same call signature and return convention as the reference Function; goes through the
    legacy launcher symbols so the boundary exercised is exactly the reference's."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1,
                stride2=2, corr_multiply=1):
        dev = _check_inputs(input1, input2)
        if input1.shape != input2.shape or input1.dim() != 4:
            raise ValueError("input1/input2 must be 4-D tensors of identical shape")
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        B, C, H, W = input1.shape
        oc, oh, ow = corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
        output = torch.empty((B, oc, oh, ow), dtype=torch.float32, device=dev)
        L = _lib.load()
        with _on_device(dev):
            ok = L.Correlation_forward_cuda_kernel(
                _ptr(output), B, oc, oh, ow, *_dense_strides(output),
                _ptr(input1), C, H, W, *_dense_strides(input1),
                _ptr(input2), C, *_dense_strides(input2),
                None, None,
                pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply, _stream())
        _lib.check(ok, "Correlation_forward_cuda_kernel")
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply = ctx.params
        grad_output = grad_output.contiguous()
        _check_inputs(grad_output)
        B, C, H, W = input1.shape
        grad_input1 = torch.empty_like(input1)
        grad_input2 = torch.empty_like(input2)
        L = _lib.load()
        with _on_device(input1.device):
            ok = L.Correlation_backward_cuda_kernel(
                _ptr(grad_output), *grad_output.shape, *_dense_strides(grad_output),
                _ptr(input1), C, H, W, *_dense_strides(input1),
                _ptr(input2), *_dense_strides(input2),
                _ptr(grad_input1), *_dense_strides(grad_input1),
                _ptr(grad_input2), C, *_dense_strides(grad_input2),
                None, None,
                pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply, _stream())
        _lib.check(ok, "Correlation_backward_cuda_kernel")
        return (grad_input1, grad_input2) + (None,) * 6


class WarpFunction(Function):
    @staticmethod
    def forward(ctx, x, flow):
        dev = _check_inputs(x, flow)
        if x.dim() != 4 or flow.dim() != 4 or flow.shape[1] != 2 or \
                flow.shape[0] != x.shape[0] or flow.shape[2:] != x.shape[2:]:
            raise ValueError(f"x {tuple(x.shape)} / flow {tuple(flow.shape)}: expected [B,C,H,W] and [B,2,H,W]")
        x = x.contiguous()
        flow = flow.contiguous()
        ctx.save_for_backward(x, flow)
        out = torch.empty_like(x)
        B, C, H, W = x.shape
        with _on_device(dev):
            ok = _lib.load().pwc_warp_forward(_ptr(x), _ptr(flow), _ptr(out), B, C, H, W, _stream())
        _lib.check(ok, "pwc_warp_forward")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, flow = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        _check_inputs(grad_out)
        B, C, H, W = x.shape
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gflow = torch.empty_like(flow) if ctx.needs_input_grad[1] else None
        L = _lib.load()
        with _on_device(x.device):
            if gx is not None and gflow is not None:
                # both gradients: the tiled kernel with its 8-channel-interleaved scratch (pwc_b200.h, ABI v5)
                nbytes = L.pwc_warp_backward_workspace(B, C, H, W)
                ws = torch.empty(max(nbytes, 16) // 4, dtype=torch.float32, device=x.device)
                ok = L.pwc_warp_backward_ws(_ptr(grad_out), _ptr(x), _ptr(flow), _ptr(gx), _ptr(gflow), B, C, H, W,
                                            ws.data_ptr(), nbytes, _stream())
                what = "pwc_warp_backward_ws"
            else:
                ok = L.pwc_warp_backward(_ptr(grad_out), _ptr(x), _ptr(flow), _ptr(gx), _ptr(gflow), B, C, H, W,
                                         _stream())
                what = "pwc_warp_backward"
        _lib.check(ok, what)
        return gx, gflow


class WarpCorrelationFunction(Function):
    """out = [leaky_relu]( Correlation(x1, warp(x2, flow)) ) in one launch; flow may be None."""

    @staticmethod
    def forward(ctx, x1, x2, flow, pad_size, kernel_size, max_displacement, stride1, stride2,
                act, slope, want_warped):
        dev = _check_inputs(x1, x2, flow)
        if x1.shape != x2.shape or x1.dim() != 4:
            raise ValueError("x1/x2 must be 4-D tensors of identical shape")
        B, C, H, W = x1.shape
        if flow is not None and tuple(flow.shape) != (B, 2, H, W):
            raise ValueError(f"flow must be [B,2,H,W] = {(B, 2, H, W)}, got {tuple(flow.shape)}")
        x1 = x1.contiguous()
        x2 = x2.contiguous()
        flow = None if flow is None else flow.contiguous()
        oc, oh, ow = corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
        out = torch.empty((B, oc, oh, ow), dtype=torch.float32, device=dev)
        # If the caller asks for x2_warp anyway (model.py:107,113) and a backward will follow, the exported
        # tensor is handed to the backward, which then skips re-evaluating the warp.  It is not exported
        # just for that: measured on B200 the export costs the forward about as much as it saves.
        needs_grad = want_warped and flow is not None and any(ctx.needs_input_grad[:3])
        warped = torch.empty_like(x2) if want_warped else None
        with _on_device(dev):
            ok = _lib.load().pwc_warpcorr_forward(
                _ptr(x1), _ptr(x2), _ptr(flow), _ptr(out), _ptr(warped), B, C, H, W,
                pad_size, kernel_size, max_displacement, stride1, stride2,
                int(bool(act)), float(slope), _stream())
        _lib.check(ok, "pwc_warpcorr_forward")
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2, bool(act), float(slope))
        ctx.has_flow = flow is not None
        if flow is None:
            ctx.save_for_backward(x1, x2, out if act else None)
        else:
            ctx.save_for_backward(x1, x2, out if act else None, flow, warped if needs_grad else None)
        if want_warped:
            ctx.mark_non_differentiable(warped)
            return out, warped
        return out

    @staticmethod
    def backward(ctx, grad_out, *unused):
        pad_size, kernel_size, max_displacement, stride1, stride2, act, slope = ctx.params
        if ctx.has_flow:
            x1, x2, out, flow, warped = ctx.saved_tensors
        else:
            x1, x2, out = ctx.saved_tensors
            flow = warped = None
        grad_out = grad_out.contiguous()
        _check_inputs(grad_out)
        B, C, H, W = x1.shape
        g1 = torch.empty_like(x1)
        g2 = torch.empty_like(x2)
        gflow = torch.empty_like(flow) if flow is not None else None
        L = _lib.load()
        ws_bytes = int(L.pwc_warpcorr_backward_workspace(B, C, H, W, int(flow is not None), pad_size,
                                                         kernel_size, max_displacement, stride1, stride2))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x1.device) if ws_bytes else None
        with _on_device(x1.device):
            ok = L.pwc_warpcorr_backward(
                _ptr(grad_out), _ptr(x1), _ptr(x2), _ptr(flow), _ptr(out), _ptr(warped), _ptr(g1), _ptr(g2),
                _ptr(gflow), _ptr(ws), ws_bytes, B, C, H, W,
                pad_size, kernel_size, max_displacement, stride1, stride2,
                int(act), float(slope), _stream())
        _lib.check(ok, "pwc_warpcorr_backward")
        return (g1, g2, gflow) + (None,) * 8


class WarpCorrelationConcatFunction(Function):
    """est_in = cat([x1, [leaky_relu](Correlation(x1, warp(x2, flow))), flow], 1) -- the flow estimator's
    input (model.py:80-91) -- with the cost volume written by the kernel straight into its channel slice and,
    in the backward, the gradient of that slice read in place (pwc_warpcorr_forward_strided /
    pwc_warpcorr_backward_strided): no torch.cat copy of the 81 channels, no slice copy of their gradient.
    Returns (est_in, x2_warp or None); flow may be None only together with a `flow_fill` tensor that supplies
    the last two channels (level 0: zero flow, no warp)."""

    @staticmethod
    def forward(ctx, x1, x2, flow, flow_fill, pad_size, kernel_size, max_displacement, stride1, stride2,
                act, slope, want_warped):
        dev = _check_inputs(x1, x2, flow, flow_fill)
        if x1.shape != x2.shape or x1.dim() != 4:
            raise ValueError("x1/x2 must be 4-D tensors of identical shape")
        B, C, H, W = x1.shape
        tail = flow if flow is not None else flow_fill
        if tail is None or tuple(tail.shape) != (B, 2, H, W):
            raise ValueError(f"flow (or flow_fill) must be [B,2,H,W] = {(B, 2, H, W)}")
        oc, oh, ow = corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
        if (oh, ow) != (H, W):
            raise ValueError("the concatenated form needs an output of the input's size (pad == max_displacement + (k-1)/2, stride1 == 1)")
        x1 = x1.contiguous()
        x2 = x2.contiguous()
        flow = None if flow is None else flow.contiguous()
        est_in = torch.empty((B, C + oc + 2, H, W), dtype=torch.float32, device=dev)
        est_in[:, :C].copy_(x1)
        est_in[:, C + oc:].copy_(tail)
        warped = torch.empty_like(x2) if want_warped else None
        with _on_device(dev):
            ok = _lib.load().pwc_warpcorr_forward_strided(
                _ptr(x1), _ptr(x2), _ptr(flow), ctypes.c_void_p(est_in.data_ptr() + 4 * C * H * W),
                (C + oc + 2) * H * W, _ptr(warped), B, C, H, W, pad_size, kernel_size, max_displacement, stride1,
                stride2, int(bool(act)), float(slope), _stream())
        _lib.check(ok, "pwc_warpcorr_forward_strided")
        ctx.params = (pad_size, kernel_size, max_displacement, stride1, stride2, bool(act), float(slope), oc)
        ctx.has_flow = flow is not None
        keep_warp = want_warped and flow is not None
        # est_in is saved only for the sign gate of the activation (its corr slice is the forward output)
        ctx.save_for_backward(x1, x2, flow, est_in if act else None, warped if keep_warp else None)
        if want_warped:
            ctx.mark_non_differentiable(warped)
            return est_in, warped
        return est_in

    @staticmethod
    def backward(ctx, grad_est, *unused):
        pad_size, kernel_size, max_displacement, stride1, stride2, act, slope, oc = ctx.params
        x1, x2, flow, est_in, warped = ctx.saved_tensors
        grad_est = grad_est.contiguous()
        _check_inputs(grad_est)
        B, C, H, W = x1.shape
        per = (C + oc + 2) * H * W
        g1 = torch.empty_like(x1)
        g2 = torch.empty_like(x2)
        gflow = torch.empty_like(flow) if flow is not None else None
        L = _lib.load()
        ws_bytes = int(L.pwc_warpcorr_backward_workspace(B, C, H, W, int(flow is not None), pad_size,
                                                         kernel_size, max_displacement, stride1, stride2))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=x1.device) if ws_bytes else None
        off = 4 * C * H * W
        with _on_device(x1.device):
            ok = L.pwc_warpcorr_backward_strided(
                ctypes.c_void_p(grad_est.data_ptr() + off), per, _ptr(x1), _ptr(x2), _ptr(flow),
                None if est_in is None else ctypes.c_void_p(est_in.data_ptr() + off), per, _ptr(warped),
                _ptr(g1), _ptr(g2), _ptr(gflow), _ptr(ws), ws_bytes, B, C, H, W,
                pad_size, kernel_size, max_displacement, stride1, stride2, int(act), float(slope), _stream())
        _lib.check(ok, "pwc_warpcorr_backward_strided")
        g1 += grad_est[:, :C]                    # x1 is also the first C channels of est_in
        gtail = grad_est[:, C + oc:]
        if flow is not None:
            gflow += gtail
            return (g1, g2, gflow, None) + (None,) * 8
        return (g1, g2, None, gtail.contiguous() if ctx.needs_input_grad[3] else None) + (None,) * 8


def warp_correlation_concat(x1, x2, flow, flow_fill=None, pad_size=4, kernel_size=1, max_displacement=4, stride1=1,
                            stride2=1, act=False, slope=0.01, return_warped=False):
    return WarpCorrelationConcatFunction.apply(x1, x2, flow, flow_fill, pad_size, kernel_size, max_displacement,
                                               stride1, stride2, act, slope, return_warped)


def correlation(input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1,
                stride2=2, corr_multiply=1):
    return CorrelationFunction.apply(input1, input2, pad_size, kernel_size, max_displacement,
                                     stride1, stride2, corr_multiply)


def warp(x, flow):
    return WarpFunction.apply(x, flow)


def _dense_image_view(t, shape, what):
    """`t` must be a view of `shape` whose images are dense ([C,H,W] contiguous) -- only the batch stride is free."""
    B, C, H, W = shape
    if tuple(t.shape) != tuple(shape) or t.stride()[1:] != (H * W, W, 1) or (B > 1 and t.stride(0) < C * H * W):
        raise ValueError(f"{what} must be a [{B},{C},{H},{W}] view with dense images, got shape "
                         f"{tuple(t.shape)} strides {t.stride()}")
    return int(t.stride(0)) if B > 1 else 0


def warp_correlation_coarse_into(out, flow_out, x1, x2, coarse_flow, pad_size=4, kernel_size=1,
                                 max_displacement=4, stride1=1, stride2=1, act=False, slope=0.01,
                                 return_warped=False):
    """model.py:78-84 in one launch (inference only): the flow arrives at the previous pyramid level's
    resolution (`coarse_flow` [B,2,H/2,W/2]); the kernel evaluates `F.upsample(coarse_flow, 2, 'bilinear') * 2`
    itself (bit for bit as torch does), writes that fine flow to the view `flow_out` ([B,2,H,W], dense images,
    free batch stride) and the cost volume to the view `out` -- typically the last two channels and the middle
    81 channels of the flow estimator's concatenated input (model.py:89-91)."""
    dev = _check_inputs(x1, x2, coarse_flow, out, flow_out)
    if torch.is_grad_enabled() and any(t.requires_grad for t in (x1, x2, coarse_flow)):
        raise RuntimeError("warp_correlation_coarse_into is inference-only; under autograd keep the flow as a tensor "
                           "(F.interpolate) and use warp_correlation")
    if x1.shape != x2.shape or x1.dim() != 4:
        raise ValueError("x1/x2 must be 4-D tensors of identical shape")
    B, C, H, W = x1.shape
    if H % 2 or W % 2 or tuple(coarse_flow.shape) != (B, 2, H // 2, W // 2):
        raise ValueError(f"coarse_flow must be [B,2,H/2,W/2] = {(B, 2, H // 2, W // 2)}, got {tuple(coarse_flow.shape)}")
    oc, oh, ow = corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    obs = _dense_image_view(out, (B, oc, oh, ow), "out")
    fobs = _dense_image_view(flow_out, (B, 2, H, W), "flow_out")
    x1, x2 = x1.detach().contiguous(), x2.detach().contiguous()
    coarse_flow = coarse_flow.detach().contiguous()
    warped = torch.empty_like(x2) if return_warped else None
    with _on_device(dev):
        ok = _lib.load().pwc_warpcorr_forward_coarse(
            _ptr(x1), _ptr(x2), _ptr(coarse_flow), _ptr(out), obs, _ptr(flow_out), fobs, _ptr(warped),
            B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
            int(bool(act)), float(slope), _stream())
    _lib.check(ok, "pwc_warpcorr_forward_coarse")
    return (out, warped) if return_warped else out


def warp_correlation_into(out, x1, x2, flow, pad_size=4, kernel_size=1, max_displacement=4, stride1=1,
                          stride2=1, act=False, slope=0.01, return_warped=False):
    """Inference-only variant that writes the cost volume into `out`, a [B, D*D, oh, ow] *view* whose
    images are dense but may be separated by a larger batch stride -- typically the channel slice
    `buf[:, C:C+81]` of the flow estimator's concatenated input (model.py:89-91), so that torch.cat
    never copies the cost volume.  No autograd graph is recorded."""
    dev = _check_inputs(x1, x2, flow, out)
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x1, x2, flow)):
        raise RuntimeError("warp_correlation_into is inference-only; use warp_correlation when gradients are needed")
    if x1.shape != x2.shape or x1.dim() != 4:
        raise ValueError("x1/x2 must be 4-D tensors of identical shape")
    B, C, H, W = x1.shape
    if flow is not None and tuple(flow.shape) != (B, 2, H, W):
        raise ValueError(f"flow must be [B,2,H,W] = {(B, 2, H, W)}, got {tuple(flow.shape)}")
    oc, oh, ow = corr_output_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2)
    if tuple(out.shape) != (B, oc, oh, ow) or out.stride()[1:] != (oh * ow, ow, 1) or \
            (B > 1 and out.stride(0) < oc * oh * ow):
        raise ValueError(f"out must be a [B,{oc},{oh},{ow}] view with dense images, got shape "
                         f"{tuple(out.shape)} strides {out.stride()}")
    x1, x2 = x1.detach().contiguous(), x2.detach().contiguous()
    flow = None if flow is None else flow.detach().contiguous()
    warped = torch.empty_like(x2) if return_warped else None
    with _on_device(dev):
        ok = _lib.load().pwc_warpcorr_forward_strided(
            _ptr(x1), _ptr(x2), _ptr(flow), _ptr(out), int(out.stride(0)) if B > 1 else 0, _ptr(warped),
            B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
            int(bool(act)), float(slope), _stream())
    _lib.check(ok, "pwc_warpcorr_forward_strided")
    return (out, warped) if return_warped else out


def warp_correlation(x1, x2, flow, pad_size=4, kernel_size=1, max_displacement=4, stride1=1,
                     stride2=1, act=False, slope=0.01, return_warped=False):
    return WarpCorrelationFunction.apply(x1, x2, flow, pad_size, kernel_size, max_displacement,
                                         stride1, stride2, act, slope, return_warped)
