"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every symbol that
include/pwc_b200.h declares (and nothing is declared that is not exported), the host-side mirror
keeps the reference's names/signatures, and the product path refuses to run without CUDA."""
import inspect
import os
import re
import subprocess

import pytest
import torch

import pwc_net_pytorch_b200 as pkg
from pwc_net_pytorch_b200 import _lib
from pwc_net_pytorch_b200 import functional as PF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pwc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)) - {"defined"}


def test_header_and_exports_agree():
    declared = _header_functions()
    assert declared == set(_lib.EXPORTS)
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, out


def test_abi_version_and_shape_on_cpu():
    L = _lib.load()
    assert L.pwc_abi_version() == 5
    assert PF.corr_output_shape(6, 7, 9, 1, 9, 1, 2) == (81, 6, 7)       # README.md:128
    assert PF.corr_output_shape(96, 112, 9, 1, 9, 1, 2) == (81, 96, 112)  # README.md:184
    assert PF.corr_output_shape(96, 112, 4, 1, 4, 1, 1) == (81, 96, 112)
    assert PF.corr_output_shape(9, 10, 3, 3, 4, 1, 2) == (25, 5, 6)
    assert PF.corr_output_shape(9, 10, 2, 1, 4, 2, 1) == (81, 3, 3)
    with pytest.raises(RuntimeError, match="empty correlation output"):
        PF.corr_output_shape(4, 4, 0, 1, 4, 1, 1)
    with pytest.raises(RuntimeError, match="kernel_size must be odd"):
        PF.corr_output_shape(8, 8, 4, 2, 4, 1, 1)


def test_reference_api_surface():
    # modules/correlation.py:8-21
    sig = inspect.signature(pkg.Correlation.__init__)
    assert list(sig.parameters)[1:] == ["pad_size", "kernel_size", "max_displacement", "stride1",
                                        "stride2", "corr_multiply"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [0, 0, 0, 1, 2, 1]
    # functions/correlation.py:8-16
    sig = inspect.signature(pkg.CorrelationFunction.forward)
    assert list(sig.parameters) == ["ctx", "input1", "input2", "pad_size", "kernel_size",
                                    "max_displacement", "stride1", "stride2", "corr_multiply"]
    assert [p.default for p in list(sig.parameters.values())[3:]] == [3, 3, 20, 1, 2, 1]
    m = pkg.Correlation(pad_size=9, kernel_size=1, max_displacement=9, stride1=1, stride2=2, corr_multiply=1)
    assert not list(m.parameters()) and not m.state_dict()
    w = pkg.WarpingLayer(args=None)
    assert not w.state_dict()
    # reference import paths (model.py:7-8)
    pkg.install_as_reference_modules()
    from correlation_package.modules.correlation import Correlation as C2
    assert C2 is pkg.Correlation


def test_no_cpu_fallback():
    a = torch.zeros(1, 2, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pkg.Correlation(4, 1, 4, 1, 1, 1)(a, a)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pkg.WarpingLayer(None)(a, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        pkg.FusedWarpCorrelation()(a, a, torch.zeros(1, 2, 8, 8))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pwc_net_pytorch_b200/ may reference it."""
    pkgdir = os.path.join(ROOT, "pwc_net_pytorch_b200")
    for dp, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)
                assert "pwc_oracle" not in txt, os.path.join(dp, f)


def test_header_compiles_as_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: include/pwc_b200.h must be valid C (no CUDA headers needed), and a plain C
    program must link against libpwc_b200.so and call it (version + output-shape arithmetic run without a GPU)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "caller.c"
    src.write_text(
        '#include <stdio.h>\n#include "pwc_b200.h"\n'
        "int main(void) {\n"
        "    int oc = 0, oh = 0, ow = 0;\n"
        "    if (pwc_abi_version() != PWC_B200_ABI_VERSION) return 2;\n"
        "    if (!pwc_corr_output_shape(96, 112, 4, 1, 4, 1, 1, &oc, &oh, &ow)) return 3;\n"
        "    if (oc != 81 || oh != 96 || ow != 112) return 4;\n"
        "    if (pwc_corr_output_shape(96, 112, 4, 2, 4, 1, 1, &oc, &oh, &ow)) return 5;   /* even kernel_size is refused */\n"
        '    printf("%s\\n", pwc_last_error());\n'
        "    return 0;\n}\n")
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lpwc_b200", "-Wl,-rpath," + libdir])
    p = subprocess.run([str(exe)], capture_output=True, text=True)
    assert p.returncode == 0, (p.returncode, p.stdout, p.stderr)
    assert "kernel_size" in p.stdout


def test_cost_volume_channel_order_is_the_reference_layers():
    """`--corr CostVolumeLayer` (model.py:21-22): the channel order the product uses to turn the fused
    kernel's raster volume into the layer's own order (modules.py:58-72) equals the permutation the
    oracle derived from the reference-authored golden outputs."""
    from oracle import torch_ref as tr
    from pwc_net_pytorch_b200.modules import cost_volume_channel_order
    assert cost_volume_channel_order(4) == [int(v) for v in tr.COSTVOLUME_PERM]
    assert sorted(cost_volume_channel_order(3)) == list(range(49))
