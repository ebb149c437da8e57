#!/bin/sh
# Compiles the reference's OWN correlation CUDA kernels, unchanged and in place, for sm_100a.
# Source stays under /root/reference (never copied into the repo); only the binary lands in
# oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
#
# The file includes only <stdio.h> and its own header (correlation_cuda_kernel.cu:1-3) and
# exports the two extern "C" launchers declared in correlation_cuda_kernel.h:5-88.
# The TH/cffi glue (correlation_cuda.c, build.py) cannot be built with torch 2.x; its ~20 lines
# of shape/zero-fill logic are restated in oracle/ref_cuda.py.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
ref="${PWC_REFERENCE_ROOT:-/root/reference}"
src="$ref/correlation_package/src/correlation_cuda_kernel.cu"
if [ ! -f "$src" ]; then
    echo "reference not present at $ref; keeping any prebuilt oracle/_ref/libref_corr.so" >&2
    exit 0
fi
mkdir -p "$here/_ref"
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Xcompiler -fPIC -shared \
     -I"$ref/correlation_package/src" -o "$here/_ref/libref_corr.so" "$src"
echo "built $here/_ref/libref_corr.so"
