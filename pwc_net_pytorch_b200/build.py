"""Builds libpwc_b200.so in-tree with nvcc for sm_100a.

Replaces the reference's correlation_package/make.sh:11 (nvcc -arch=sm_52 -> .o) and
correlation_package/build.py:18-28 (torch.utils.ffi.create_extension, removed from PyTorch):
the product is a plain C-ABI shared library (include/pwc_b200.h) with no TH/ATen dependency,
loaded with ctypes by pwc_net_pytorch_b200/_lib.py.

    python -m pwc_net_pytorch_b200.build [--force] [--verbose]
"""
import argparse
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(OUT_DIR, "libpwc_b200.so")

def _sources():
    return [os.path.join(CSRC, "pwc_abi.cu")]


def _deps():
    deps = [os.path.join(ROOT, "include", "pwc_b200.h"), os.path.abspath(__file__)]
    for f in os.listdir(CSRC):
        deps.append(os.path.join(CSRC, f))
    return deps


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libpwc_b200.so cannot be built")


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
           "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include")]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH] + _sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
