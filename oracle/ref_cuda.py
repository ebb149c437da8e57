"""Driver for the reference's OWN CUDA kernels -- TEST INFRASTRUCTURE ONLY.

oracle/_ref/libref_corr.so is /root/reference/correlation_package/src/correlation_cuda_kernel.cu
compiled unchanged for sm_100a by oracle/build_ref.sh.  The TH/cffi glue around it cannot be
built with torch 2.x, so the ~20 lines of correlation_cuda.c that matter are restated here:
shapes (correlation_cuda.c:20-34), resize + zero-fill of the padded channels-last scratch and of
the outputs (:36-42, :113-121), pointer/stride unpacking (:44-81, :124-171).

Used as (a) the strongest parity pin (tests -m gpu and tests/golden/make_golden_gpu.py) and
(b) the "GPU reference bar" timed by bench.py next to the product kernels.
"""
import ctypes
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_corr.so")
_lib = None
_vp, _i = ctypes.c_void_p, ctypes.c_int


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(SO)
        L.Correlation_forward_cuda_kernel.argtypes = (
            [_vp] + [_i] * 8 + [_vp] + [_i] * 7 + [_vp] + [_i] * 5 + [_vp, _vp] + [_i] * 6 + [_vp])
        L.Correlation_forward_cuda_kernel.restype = _i
        L.Correlation_backward_cuda_kernel.argtypes = (
            [_vp] + [_i] * 8 + [_vp] + [_i] * 7 + [_vp] + [_i] * 4 + [_vp] + [_i] * 4 + [_vp] +
            [_i] * 5 + [_vp, _vp] + [_i] * 6 + [_vp])
        L.Correlation_backward_cuda_kernel.restype = _i
        _lib = L
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def out_shape(H, W, pad, k, md, s1, s2):
    kr = (k - 1) // 2
    border = kr + md
    r = md // s2
    oc = (2 * r + 1) ** 2
    oh = math.ceil(float(H + 2 * pad - 2 * border) / float(s1))
    ow = math.ceil(float(W + 2 * pad - 2 * border) / float(s1))
    return oc, oh, ow


def correlation_forward(in1, in2, pad, k, md, s1, s2):
    """correlation_cuda.c:11-93."""
    B, C, H, W = in1.shape
    oc, oh, ow = out_shape(H, W, pad, k, md, s1, s2)
    r1 = torch.zeros((B, H + 2 * pad, W + 2 * pad, C), dtype=torch.float32, device=in1.device)
    r2 = torch.zeros_like(r1)
    out = torch.zeros((B, oc, oh, ow), dtype=torch.float32, device=in1.device)
    ok = lib().Correlation_forward_cuda_kernel(
        _p(out), B, oc, oh, ow, *out.stride(),
        _p(in1), C, H, W, *in1.stride(),
        _p(in2), C, *in2.stride(),
        _p(r1), _p(r2), pad, k, md, s1, s2, 1, _st())
    if not ok:
        raise RuntimeError("reference Correlation_forward_cuda_kernel failed")
    return out


def correlation_backward(gout, in1, in2, pad, k, md, s1, s2):
    """correlation_cuda.c:95-180."""
    B, C, H, W = in1.shape
    r1 = torch.zeros((B, H + 2 * pad, W + 2 * pad, C), dtype=torch.float32, device=in1.device)
    r2 = torch.zeros_like(r1)
    g1 = torch.zeros_like(in1)
    g2 = torch.zeros_like(in2)
    ok = lib().Correlation_backward_cuda_kernel(
        _p(gout), *gout.shape, *gout.stride(),
        _p(in1), C, H, W, *in1.stride(),
        _p(in2), *in2.stride(),
        _p(g1), *g1.stride(),
        _p(g2), C, *g2.stride(),
        _p(r1), _p(r2), pad, k, md, s1, s2, 1, _st())
    if not ok:
        raise RuntimeError("reference Correlation_backward_cuda_kernel failed")
    return g1, g2
