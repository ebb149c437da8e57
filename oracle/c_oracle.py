"""ctypes/numpy wrapper around oracle/pwc_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Every function mirrors one C function; see pwc_oracle.c for the reference file:line each follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpwc_oracle.so")
_lib = None

_fp = ctypes.POINTER(ctypes.c_float)
_i = ctypes.c_int


def build(force=False):
    src = os.path.join(_HERE, "pwc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["sh", os.path.join(_HERE, "build_oracle.sh")],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.pwc_oracle_corr_shape.argtypes = [_i] * 7 + [ctypes.POINTER(_i)] * 3
        L.pwc_oracle_corr_shape.restype = None
        L.pwc_oracle_corr_forward.argtypes = [_fp, _fp, _fp] + [_i] * 10
        L.pwc_oracle_corr_backward.argtypes = [_fp] * 5 + [_i] * 10
        L.pwc_oracle_warp_forward.argtypes = [_fp, _fp, _fp] + [_i] * 5
        L.pwc_oracle_warp_backward.argtypes = [_fp] * 5 + [_i] * 4
        L.pwc_oracle_warpcorr_forward.argtypes = [_fp] * 5 + [_i] * 10 + [ctypes.c_float, _i]
        L.pwc_oracle_warpcorr_backward.argtypes = [_fp] * 8 + [_i] * 10 + [ctypes.c_float]
        for name in ("pwc_oracle_corr_forward", "pwc_oracle_corr_backward",
                     "pwc_oracle_warp_forward", "pwc_oracle_warp_backward",
                     "pwc_oracle_warpcorr_forward", "pwc_oracle_warpcorr_backward"):
            getattr(L, name).restype = _i
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


def corr_shape(H, W, pad, k, md, s1, s2):
    oc, oh, ow = _i(), _i(), _i()
    lib().pwc_oracle_corr_shape(H, W, pad, k, md, s1, s2,
                                ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow))
    return oc.value, oh.value, ow.value


def corr_forward(in1, in2, pad, k, md, s1, s2, mode=0):
    in1, in2 = _f32(in1), _f32(in2)
    B, C, H, W = in1.shape
    oc, oh, ow = corr_shape(H, W, pad, k, md, s1, s2)
    out = np.zeros((B, oc, max(oh, 0), max(ow, 0)), np.float32)
    ok = lib().pwc_oracle_corr_forward(_p(in1), _p(in2), _p(out), B, C, H, W,
                                       pad, k, md, s1, s2, mode)
    if not ok:
        raise RuntimeError("oracle corr_forward: empty output")
    return out


def corr_backward(gout, in1, in2, pad, k, md, s1, s2, mode=0):
    gout, in1, in2 = _f32(gout), _f32(in1), _f32(in2)
    B, C, H, W = in1.shape
    g1 = np.zeros_like(in1)
    g2 = np.zeros_like(in1)
    ok = lib().pwc_oracle_corr_backward(_p(gout), _p(in1), _p(in2), _p(g1), _p(g2),
                                        B, C, H, W, pad, k, md, s1, s2, mode)
    if not ok:
        raise RuntimeError("oracle corr_backward: unsupported (stride1 != 1 or empty output)")
    return g1, g2


def warp_forward(x, flow, mode=0):
    x, flow = _f32(x), _f32(flow)
    B, C, H, W = x.shape
    out = np.zeros_like(x)
    lib().pwc_oracle_warp_forward(_p(x), _p(flow), _p(out), B, C, H, W, mode)
    return out


def warp_backward(gout, x, flow):
    gout, x, flow = _f32(gout), _f32(x), _f32(flow)
    B, C, H, W = x.shape
    gx = np.zeros_like(x)
    gflow = np.zeros_like(flow)
    ok = lib().pwc_oracle_warp_backward(_p(gout), _p(x), _p(flow), _p(gx), _p(gflow), B, C, H, W)
    if not ok:
        raise MemoryError("oracle warp_backward")
    return gx, gflow


def warpcorr_forward(f1, f2, flow, pad, k, md, s1, s2, act=False, slope=0.01, mode=0,
                     return_warped=False):
    f1, f2 = _f32(f1), _f32(f2)
    flow = None if flow is None else _f32(flow)
    B, C, H, W = f1.shape
    oc, oh, ow = corr_shape(H, W, pad, k, md, s1, s2)
    out = np.zeros((B, oc, oh, ow), np.float32)
    warped = np.zeros_like(f2) if return_warped else None
    ok = lib().pwc_oracle_warpcorr_forward(_p(f1), _p(f2), _p(flow), _p(out), _p(warped),
                                           B, C, H, W, pad, k, md, s1, s2,
                                           int(bool(act)), float(slope), mode)
    if not ok:
        raise RuntimeError("oracle warpcorr_forward failed")
    return (out, warped) if return_warped else out


def warpcorr_backward(gout, f1, f2, flow, out, pad, k, md, s1, s2, act=False, slope=0.01):
    gout, f1, f2 = _f32(gout), _f32(f1), _f32(f2)
    flow = None if flow is None else _f32(flow)
    out = None if out is None else _f32(out)
    if act and out is None:
        raise ValueError("act=True needs the forward output")
    B, C, H, W = f1.shape
    g1, g2 = np.zeros_like(f1), np.zeros_like(f2)
    gflow = None if flow is None else np.zeros_like(flow)
    ok = lib().pwc_oracle_warpcorr_backward(_p(gout), _p(f1), _p(f2), _p(flow), _p(out),
                                            _p(g1), _p(g2), _p(gflow), B, C, H, W,
                                            pad, k, md, s1, s2, int(bool(act)), float(slope))
    if not ok:
        raise RuntimeError("oracle warpcorr_backward failed")
    return g1, g2, gflow
