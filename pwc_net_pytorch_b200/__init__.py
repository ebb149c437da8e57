"""pwc_net_pytorch_b200 -- B200 (sm_100a) warp + cost-volume hot path of PWC-Net behind the
reference's own operator API (daigo0927/PWC-Net_pytorch).

Public surface:
    Correlation, CorrelationFunction        (correlation_package mirror)
    WarpingLayer, FusedWarpCorrelation      (modules.py mirror + fused entry point)
    functional.correlation / warp / warp_correlation
    build.build()                           compiles lib/libpwc_b200.so (nvcc, sm_100a)

The compute path is hand-written CUDA reached through the C ABI in include/pwc_b200.h; importing
this package does not need a GPU, calling an op does (there is no CPU fallback).
"""
import sys

from . import functional
from .correlation_package.functions.correlation import CorrelationFunction
from .correlation_package.modules.correlation import Correlation
from .modules import FusedWarpCorrelation, WarpingLayer

__all__ = ["Correlation", "CorrelationFunction", "WarpingLayer", "FusedWarpCorrelation",
           "functional", "install_as_reference_modules"]
# model.Net / graphed.GraphedForward are imported on demand (they pull in the convolution stack)


def install_as_reference_modules():
    """Registers this package's correlation_package under the reference's top-level module names
    so that unmodified reference code (`model.py:8`) imports the B200 implementation."""
    from . import correlation_package as cp
    sys.modules.setdefault("correlation_package", cp)
    sys.modules.setdefault("correlation_package.modules", cp.modules)
    sys.modules.setdefault("correlation_package.modules.correlation", cp.modules.correlation)
    sys.modules.setdefault("correlation_package.functions", cp.functions)
    sys.modules.setdefault("correlation_package.functions.correlation", cp.functions.correlation)
