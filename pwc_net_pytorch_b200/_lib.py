"""ctypes binding of libpwc_b200.so (include/pwc_b200.h).

Replaces the reference's cffi loader correlation_package/_ext/correlation/__init__.py:1-15
(torch.utils.ffi._wrap_function, removed from PyTorch).  There is no CPU fallback: if the
library is missing or a launch fails, a RuntimeError is raised (the reference's wrapper turns a
failed launch into THError("aborting"), correlation_cuda.c:87-89).
"""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libpwc_b200.so")

_c_float_p = ctypes.c_void_p   # raw device pointers (tensor.data_ptr())
_int = ctypes.c_int
_stream = ctypes.c_void_p

_lock = threading.Lock()
_lib = None

# every symbol include/pwc_b200.h declares; tests check the .so exports exactly these
EXPORTS = (
    "Correlation_forward_cuda_kernel",
    "Correlation_backward_cuda_kernel",
    "pwc_corr_output_shape",
    "pwc_warp_forward",
    "pwc_warp_backward",
    "pwc_warpcorr_forward",
    "pwc_warp_backward_workspace",
    "pwc_warp_backward_ws",
    "pwc_warpcorr_forward_strided",
    "pwc_warpcorr_forward_coarse",
    "pwc_warpcorr_backward_workspace",
    "pwc_warpcorr_backward",
    "pwc_warpcorr_backward_strided",
    "pwc_last_error",
    "pwc_abi_version",
    "pwc_launch_count",
    "pwc_set_force_generic",
    "pwc_set_disable_tma",
    "pwc_set_disable_small",
    "pwc_set_disable_seq",
)


def _declare(L):
    L.Correlation_forward_cuda_kernel.argtypes = (
        [_c_float_p] + [_int] * 8 + [_c_float_p] + [_int] * 7 + [_c_float_p] + [_int] * 5 +
        [_c_float_p, _c_float_p] + [_int] * 6 + [_stream])
    L.Correlation_forward_cuda_kernel.restype = _int
    L.Correlation_backward_cuda_kernel.argtypes = (
        [_c_float_p] + [_int] * 8 + [_c_float_p] + [_int] * 7 + [_c_float_p] + [_int] * 4 +
        [_c_float_p] + [_int] * 4 + [_c_float_p] + [_int] * 5 + [_c_float_p, _c_float_p] +
        [_int] * 6 + [_stream])
    L.Correlation_backward_cuda_kernel.restype = _int
    L.pwc_corr_output_shape.argtypes = [_int] * 7 + [ctypes.POINTER(_int)] * 3
    L.pwc_corr_output_shape.restype = _int
    L.pwc_warp_forward.argtypes = [_c_float_p] * 3 + [_int] * 4 + [_stream]
    L.pwc_warp_forward.restype = _int
    L.pwc_warp_backward.argtypes = [_c_float_p] * 5 + [_int] * 4 + [_stream]
    L.pwc_warp_backward.restype = _int
    L.pwc_warp_backward_workspace.argtypes = [_int] * 4
    L.pwc_warp_backward_workspace.restype = ctypes.c_longlong
    L.pwc_warp_backward_ws.argtypes = [_c_float_p] * 5 + [_int] * 4 + [ctypes.c_void_p, ctypes.c_longlong, _stream]
    L.pwc_warp_backward_ws.restype = _int
    L.pwc_warpcorr_forward.argtypes = ([_c_float_p] * 5 + [_int] * 9 + [_int, ctypes.c_float] +
                                       [_stream])
    L.pwc_warpcorr_forward.restype = _int
    L.pwc_warpcorr_forward_strided.argtypes = ([_c_float_p] * 4 + [ctypes.c_longlong, _c_float_p] + [_int] * 9 +
                                               [_int, ctypes.c_float] + [_stream])
    L.pwc_warpcorr_forward_strided.restype = _int
    L.pwc_warpcorr_forward_coarse.argtypes = ([_c_float_p] * 4 + [ctypes.c_longlong, _c_float_p, ctypes.c_longlong,
                                               _c_float_p] + [_int] * 9 + [_int, ctypes.c_float] + [_stream])
    L.pwc_warpcorr_forward_coarse.restype = _int
    L.pwc_warpcorr_backward_workspace.argtypes = [_int] * 10
    L.pwc_warpcorr_backward_workspace.restype = ctypes.c_longlong
    L.pwc_warpcorr_backward.argtypes = ([_c_float_p] * 9 + [ctypes.c_void_p, ctypes.c_longlong] +
                                        [_int] * 9 + [_int, ctypes.c_float] + [_stream])
    L.pwc_warpcorr_backward.restype = _int
    L.pwc_warpcorr_backward_strided.argtypes = ([_c_float_p, ctypes.c_longlong] + [_c_float_p] * 4 + [ctypes.c_longlong] +
                                                [_c_float_p] * 4 + [ctypes.c_void_p, ctypes.c_longlong] +
                                                [_int] * 9 + [_int, ctypes.c_float] + [_stream])
    L.pwc_warpcorr_backward_strided.restype = _int
    L.pwc_last_error.argtypes = []
    L.pwc_last_error.restype = ctypes.c_char_p
    L.pwc_abi_version.argtypes = []
    L.pwc_abi_version.restype = _int
    L.pwc_launch_count.argtypes = []
    L.pwc_launch_count.restype = ctypes.c_longlong
    L.pwc_set_force_generic.argtypes = [_int]
    L.pwc_set_force_generic.restype = _int
    L.pwc_set_disable_tma.argtypes = [_int]
    L.pwc_set_disable_tma.restype = _int
    L.pwc_set_disable_small.argtypes = [_int]
    L.pwc_set_disable_small.restype = _int
    L.pwc_set_disable_seq.argtypes = [_int]
    L.pwc_set_disable_seq.restype = _int


def load():
    """Loads (once) and returns the ctypes handle.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m pwc_net_pytorch_b200.build` "
                    "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
            L = ctypes.CDLL(LIB_PATH)
            _declare(L)
            if L.pwc_abi_version() != 5:
                raise RuntimeError("libpwc_b200.so ABI version mismatch; rebuild it")
            _lib = L
    return _lib


def last_error():
    return load().pwc_last_error().decode("utf-8", "replace")


def check(ok, what):
    """0 from a launcher -> RuntimeError (reference: THError('aborting'), correlation_cuda.c:87-89)."""
    if not ok:
        raise RuntimeError(f"{what} failed: {last_error()}")


def launch_count():
    return int(load().pwc_launch_count())
