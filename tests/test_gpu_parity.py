"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes) behind the
reference-shaped Python API, against the oracle on identical seeded inputs.

Tolerance (north_star): per-level cost volumes and warped features within 1e-5 relative (fp32),
judged in the max norm: ||new - ref||_inf <= 1e-5 * ||ref||_inf.
Backward passes use fp32 atomics for the feature scatter, so their bound is the same relative
tolerance rather than bit equality."""
import os

import numpy as np
import pytest
import torch

import pwc_net_pytorch_b200 as pkg
from pwc_net_pytorch_b200 import _lib
from pwc_net_pytorch_b200 import functional as PF
from oracle import c_oracle as co
from oracle import ref_cuda
from oracle import torch_ref as tr
from util import CANON_CFG, GENERIC_CFGS, REF_CFG, make_inputs, max_rel

pytestmark = pytest.mark.gpu
TOL = 1e-5
HERE = os.path.dirname(os.path.abspath(__file__))


def dev():
    return torch.device("cuda:0")


def to_dev(*arrs):
    return [None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev()) for a in arrs]


def test_extension_loaded_and_counts_launches():
    L = _lib.load()
    before = L.pwc_launch_count()
    a, b, f = to_dev(*make_inputs(1, 4, 8, 8)[:3])
    pkg.FusedWarpCorrelation()(a, b, f)
    torch.cuda.synchronize()
    assert L.pwc_launch_count() == before + 1      # one fused launch, nothing else
    maps = open("/proc/self/maps").read()
    assert "libpwc_b200.so" in maps


# pyramid level shapes of a 384x448 pair (README.md:127-184) at small batch, plus ragged ones
LEVEL_SHAPES = [(2, 192, 6, 7), (2, 196, 6, 7), (2, 128, 12, 14), (1, 96, 24, 28), (1, 64, 48, 56),
                (1, 32, 96, 112), (1, 5, 9, 11), (1, 3, 33, 37), (3, 1, 8, 40), (1, 17, 2, 2)]


@pytest.mark.parametrize("shape", LEVEL_SHAPES)
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
def test_fused_forward_vs_oracle(shape, cfg):
    B, C, H, W = shape
    f1, f2, flow, _ = make_inputs(B, C, H, W, seed=B * 1000 + C + H)
    ref, ref_w = co.warpcorr_forward(f1, f2, flow, *cfg, return_warped=True)
    a, b, f = to_dev(f1, f2, flow)
    m = pkg.FusedWarpCorrelation(*cfg, return_warped=True)
    out, warped = m(a, b, f)
    assert tuple(out.shape) == ref.shape
    assert max_rel(out.cpu().numpy(), ref) < TOL
    assert max_rel(warped.cpu().numpy(), ref_w) < TOL
    # without the warped export the result is identical
    out2 = pkg.FusedWarpCorrelation(*cfg)(a, b, f)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("kind", ["iid", "smooth", "integer", "zero"])
@pytest.mark.parametrize("act", [False, True])
def test_fused_forward_flow_kinds_and_activation(kind, act):
    f1, f2, flow, _ = make_inputs(2, 32, 24, 28, seed=17, flow_sigma=3.0, flow_kind=kind)
    for cfg in (REF_CFG, CANON_CFG):
        ref = co.warpcorr_forward(f1, f2, flow, *cfg, act=act, slope=0.01)
        out = pkg.FusedWarpCorrelation(*cfg, activation=act, negative_slope=0.01)(*to_dev(f1, f2, flow))
        assert max_rel(out.cpu().numpy(), ref) < TOL
        if kind == "zero":      # l == 0 of the pyramid loop (model.py:74-76): warp is the identity
            plain = pkg.Correlation(*cfg)(*to_dev(f1, f2))
            if act:
                plain = torch.nn.functional.leaky_relu(plain, 0.01)
            assert torch.equal(out, plain)


@pytest.mark.parametrize("shape", LEVEL_SHAPES)
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
def test_correlation_module_vs_oracle(shape, cfg):
    B, C, H, W = shape
    f1, f2, _, _ = make_inputs(B, C, H, W, seed=7)
    ref = co.corr_forward(f1, f2, *cfg)
    out = pkg.Correlation(*cfg, corr_multiply=1)(*to_dev(f1, f2))
    assert max_rel(out.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("cfg", GENERIC_CFGS)
def test_generic_parameters_vs_oracle(cfg):
    f1, f2, flow, rng = make_inputs(2, 7, 13, 15, seed=23)
    ref = co.corr_forward(f1, f2, *cfg)
    a, b, f = to_dev(f1, f2, flow)
    a.requires_grad_(); b.requires_grad_()
    out = pkg.Correlation(*cfg)(a, b)
    assert max_rel(out.detach().cpu().numpy(), ref) < TOL
    reff = co.warpcorr_forward(f1, f2, flow, *cfg)
    assert max_rel(pkg.FusedWarpCorrelation(*cfg)(a.detach(), b.detach(), f).cpu().numpy(), reff) < TOL
    go = rng.standard_normal(ref.shape).astype(np.float32)
    if cfg[3] == 1:
        out.backward(to_dev(go)[0])
        g1, g2 = co.corr_backward(go, f1, f2, *cfg)
        assert max_rel(a.grad.cpu().numpy(), g1) < TOL
        assert max_rel(b.grad.cpu().numpy(), g2) < TOL
    else:
        with pytest.raises(RuntimeError, match="stride1 == 1"):
            out.backward(to_dev(go)[0])


@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
@pytest.mark.parametrize("shape", [(2, 32, 48, 56), (1, 7, 16, 16), (1, 10, 24, 28), (2, 5, 9, 64), (1, 33, 40, 20)])
def test_tma_kernel_equals_plain_tiled_kernel(cfg, shape):
    """The TMA-staged warp-specialised kernel and the plain tiled kernel are two implementations
    of the same arithmetic; W % 4 == 0 shapes take the TMA path by default."""
    B, C, H, W = shape
    f1, f2, flow, _ = make_inputs(B, C, H, W, seed=53)
    a, b, f = to_dev(f1, f2, flow)
    L = _lib.load()
    outs = {}
    for use_flow in (True, False):
        tma = pkg.FusedWarpCorrelation(*cfg, return_warped=True)(a, b, f if use_flow else None)
        prev = L.pwc_set_disable_tma(1)
        try:
            plain = pkg.FusedWarpCorrelation(*cfg, return_warped=True)(a, b, f if use_flow else None)
        finally:
            L.pwc_set_disable_tma(prev)
        assert max_rel(tma[0].cpu().numpy(), plain[0].cpu().numpy()) < 2e-6
        assert max_rel(tma[1].cpu().numpy(), plain[1].cpu().numpy()) < 2e-6


@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
@pytest.mark.parametrize("kind", ["large_smooth", "outliers", "huge_iid"])
def test_tma_window_placement_and_global_fallback(cfg, kind):
    """The f2 source window follows the flow's bounding box (large coherent motion), and samples
    whose footprint leaves the window are gathered from global memory (outliers / huge spread)."""
    B, C, H, W = 2, 12, 48, 64
    f1, f2, flow, rng = make_inputs(B, C, H, W, seed=59, flow_sigma=1.0, flow_kind="smooth")
    if kind == "large_smooth":
        flow = flow + np.array([17.3, -11.6], np.float32).reshape(1, 2, 1, 1)
    elif kind == "outliers":
        idx = rng.integers(0, H * W, size=40)
        flat = flow.reshape(B, 2, -1)
        flat[:, :, idx] += rng.choice([-30.0, 25.5, 40.25], size=(B, 2, 40)).astype(np.float32)
    else:
        flow = (12.0 * rng.standard_normal(flow.shape)).astype(np.float32)
    ref, ref_w = co.warpcorr_forward(f1, f2, flow, *cfg, return_warped=True)
    out, warped = pkg.FusedWarpCorrelation(*cfg, return_warped=True)(*to_dev(f1, f2, flow))
    assert max_rel(warped.cpu().numpy(), ref_w) < TOL
    assert max_rel(out.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
def test_forced_generic_equals_tiled(cfg):
    f1, f2, flow, rng = make_inputs(2, 9, 17, 19, seed=29)
    a, b, f = to_dev(f1, f2, flow)
    L = _lib.load()
    tiled = pkg.FusedWarpCorrelation(*cfg)(a, b, f)
    prev = L.pwc_set_force_generic(1)
    try:
        generic = pkg.FusedWarpCorrelation(*cfg)(a, b, f)
    finally:
        L.pwc_set_force_generic(prev)
    assert max_rel(tiled.cpu().numpy(), generic.cpu().numpy()) < 2e-6


BWD_SHAPES = [(2, 192, 6, 7), (2, 196, 6, 7), (1, 96, 24, 28), (1, 32, 48, 56), (1, 5, 9, 11), (2, 3, 33, 37)]


@pytest.mark.parametrize("shape", BWD_SHAPES)
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
def test_correlation_backward_vs_oracle(shape, cfg):
    B, C, H, W = shape
    f1, f2, _, rng = make_inputs(B, C, H, W, seed=31)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    g1, g2 = co.corr_backward(go, f1, f2, *cfg)
    a, b, g = to_dev(f1, f2, go)
    a.requires_grad_(); b.requires_grad_()
    pkg.Correlation(*cfg)(a, b).backward(g)
    assert max_rel(a.grad.cpu().numpy(), g1) < TOL
    assert max_rel(b.grad.cpu().numpy(), g2) < TOL


@pytest.mark.parametrize("shape", BWD_SHAPES)
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
@pytest.mark.parametrize("act", [False, True])
def test_fused_backward_vs_oracle(shape, cfg, act):
    B, C, H, W = shape
    f1, f2, flow, rng = make_inputs(B, C, H, W, seed=37, flow_sigma=2.0)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    a, b, f, g = to_dev(f1, f2, flow, go)
    for t in (a, b, f):
        t.requires_grad_()
    out = pkg.FusedWarpCorrelation(*cfg, activation=act, negative_slope=0.01)(a, b, f)
    out.backward(g)
    fwd = co.warpcorr_forward(f1, f2, flow, *cfg, act=act, slope=0.01)
    assert max_rel(out.detach().cpu().numpy(), fwd) < TOL
    # leaky_relu_'s backward is gated by the sign of the *stored* forward output; hand the oracle the
    # same output so that elements within rounding of zero take the same branch on both sides
    g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow, out.detach().cpu().numpy(), *cfg, act=act, slope=0.01)
    assert max_rel(a.grad.cpu().numpy(), g1) < TOL
    assert max_rel(b.grad.cpu().numpy(), g2) < TOL
    assert max_rel(f.grad.cpu().numpy(), gf) < TOL


def test_warping_layer_vs_oracle_and_torch():
    for shape in [(2, 192, 6, 7), (1, 32, 96, 112), (1, 3, 33, 37)]:
        B, C, H, W = shape
        _, f2, flow, rng = make_inputs(B, C, H, W, seed=41, flow_sigma=3.0)
        x, f = to_dev(f2, flow)
        x.requires_grad_(); f.requires_grad_()
        y = pkg.WarpingLayer(None)(x, f)
        assert max_rel(y.detach().cpu().numpy(), co.warp_forward(f2, flow, 0)) < TOL
        # torch's own grid_sample with the 0.4.0 semantics, on the same device
        assert max_rel(y.detach().cpu().numpy(), tr.warping_layer_port(x.detach(), f.detach()).cpu().numpy()) < 3e-5
        go = rng.standard_normal(f2.shape).astype(np.float32)
        y.backward(to_dev(go)[0])
        gx, gf = co.warp_backward(go, f2, flow)
        assert max_rel(x.grad.cpu().numpy(), gx) < TOL
        assert max_rel(f.grad.cpu().numpy(), gf) < TOL


def test_golden_reference_python_fixtures(golden):
    """Reference-authored outputs (modules.WarpingLayer / CostVolumeLayer)."""
    for name in ("tiny", "lvl6", "mid", "big_flow"):
        f1, f2, flow = (golden[f"{name}/{k}"] for k in ("f1", "f2", "flow"))
        C = f1.shape[1]
        out, warped = pkg.FusedWarpCorrelation(*CANON_CFG, return_warped=True)(*to_dev(f1, f2, flow))
        assert max_rel(warped.cpu().numpy(), golden[f"{name}/warp"]) < TOL
        got = out.cpu().numpy()[:, tr.COSTVOLUME_PERM] * C
        assert max_rel(got, golden[f"{name}/costvolume_of_warp"] * 81.0) < TOL


@pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref/libref_corr.so not built")
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG, (4, 3, 4, 1, 2), (4, 1, 4, 1, 2)])
def test_against_reference_cuda_kernels_live(cfg):
    """The reference's own kernels (compiled unchanged for sm_100a) on the same device."""
    f1, f2, _, rng = make_inputs(2, 37, 12, 14, seed=43)
    a, b = to_dev(f1, f2)
    ref = ref_cuda.correlation_forward(a, b, *cfg)
    a.requires_grad_(); b.requires_grad_()
    out = pkg.Correlation(*cfg)(a, b)
    assert max_rel(out.detach().cpu().numpy(), ref.cpu().numpy()) < TOL
    go = torch.from_numpy(rng.standard_normal(tuple(ref.shape)).astype(np.float32)).to(dev())
    r1, r2 = ref_cuda.correlation_backward(go, a.detach(), b.detach(), *cfg)
    out.backward(go)
    assert max_rel(a.grad.cpu().numpy(), r1.cpu().numpy()) < TOL
    assert max_rel(b.grad.cpu().numpy(), r2.cpu().numpy()) < TOL


def test_edge_cases():
    f1, f2, flow, _ = make_inputs(1, 3, 6, 7)
    a, b = to_dev(f1, f2)
    far = torch.full((1, 2, 6, 7), 1000.0, device=dev())
    assert not pkg.FusedWarpCorrelation()(a, b, far).any()
    bad = torch.from_numpy(flow).to(dev())
    bad[0, 0, 2, 3] = float("nan")
    bad[0, 1, 1, 1] = float("inf")
    w = pkg.WarpingLayer(None)(b, bad)
    assert torch.isfinite(w).all() and not w[0, :, 2, 3].any() and not w[0, :, 1, 1].any()
    out = pkg.FusedWarpCorrelation()(a, b, bad)
    assert torch.isfinite(out).all()
    # non-contiguous inputs are accepted (made dense), wrong dtypes / shapes are refused
    nc = torch.randn(1, 6, 7, 3, device=dev()).permute(0, 3, 1, 2)
    assert max_rel(pkg.Correlation(4, 1, 4, 1, 1)(nc, nc).cpu().numpy(),
                   co.corr_forward(nc.cpu().numpy(), nc.cpu().numpy(), 4, 1, 4, 1, 1)) < TOL
    with pytest.raises(TypeError):
        pkg.Correlation(4, 1, 4, 1, 1)(a.double(), b.double())
    with pytest.raises(ValueError):
        pkg.FusedWarpCorrelation()(a, b, torch.zeros(1, 2, 5, 7, device=dev()))
    with pytest.raises(RuntimeError, match="empty correlation output"):
        pkg.Correlation(0, 1, 4, 1, 1)(a, b)


def test_full_size_properties():
    """BASELINE.json config 2 at full size (B=32, C=32, 96x112 and C=196, 6x7), through
    size-independent properties instead of the (slow) CPU oracle:
      * linearity in f1 and in f2;  * the centre channel with zero flow is mean_c f1*f2;
      * zero flow == plain Correlation;  * <grad_out, J v> == <J^T grad_out, v> (adjointness);
      * a spot check of 2 images against the oracle."""
    torch.manual_seed(0)
    for (B, C, H, W) in [(32, 32, 96, 112), (32, 196, 6, 7)]:
        d = dev()
        f1 = torch.randn(B, C, H, W, device=d)
        f2 = torch.randn(B, C, H, W, device=d)
        g2 = torch.randn(B, C, H, W, device=d)
        flow = 2.0 * torch.randn(B, 2, H, W, device=d)
        op = pkg.FusedWarpCorrelation(*CANON_CFG)
        o1 = op(f1, f2, flow)
        lin = op(f1, 2.0 * f2 + 0.5 * g2, flow)
        assert max_rel((lin - 2.0 * o1 - 0.5 * op(f1, g2, flow)).cpu().numpy() + o1.cpu().numpy(), o1.cpu().numpy()) < 5e-6
        z = torch.zeros_like(flow)
        oz = op(f1, f2, z)
        assert torch.equal(oz, pkg.Correlation(*CANON_CFG)(f1, f2))
        assert max_rel(oz[:, 40].cpu().numpy(), (f1 * f2).mean(1).cpu().numpy()) < TOL
        # adjointness of the backward kernels w.r.t. the forward (f1 and f2 directions)
        a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_()
        go = torch.randn_like(o1)
        op(a, b, flow).backward(go)
        v1 = torch.randn_like(f1)
        lhs = (go.double() * op(v1, f2, flow).double()).sum().item()
        rhs = (a.grad.double() * v1.double()).sum().item()
        assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)
        v2 = torch.randn_like(f2)
        lhs = (go.double() * op(f1, v2, flow).double()).sum().item()
        rhs = (b.grad.double() * v2.double()).sum().item()
        assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0)
        # oracle spot check on two images
        sel = [0, B - 1]
        ref = co.warpcorr_forward(f1[sel].cpu().numpy(), f2[sel].cpu().numpy(), flow[sel].cpu().numpy(), *CANON_CFG)
        assert max_rel(o1[sel].cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("shape", [(1, 32, 112, 256), (1, 32, 96, 320), (2, 64, 56, 128), (1, 16, 50, 52), (3, 5, 17, 36),
                                   (1, 40, 16, 16), (1, 9, 8, 24)])
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
def test_highres_and_ragged_shapes_fwd_bwd(shape, cfg):
    """BASELINE.json config 5 level-2 shapes (Sintel 1024x448 -> 112x256, KITTI 1280x384 -> 96x320) and
    sizes that are not multiples of the 16x16 tile (TMA path, W % 4 == 0) -- forward and all gradients."""
    B, C, H, W = shape
    f1, f2, flow, rng = make_inputs(B, C, H, W, seed=61, flow_sigma=2.5)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    a, b, f, g = to_dev(f1, f2, flow, go)
    for t in (a, b, f):
        t.requires_grad_()
    out = pkg.FusedWarpCorrelation(*cfg)(a, b, f)
    out.backward(g)
    ref = co.warpcorr_forward(f1, f2, flow, *cfg)
    g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow, None, *cfg)
    assert max_rel(out.detach().cpu().numpy(), ref) < TOL
    assert max_rel(a.grad.cpu().numpy(), g1) < TOL
    assert max_rel(b.grad.cpu().numpy(), g2) < TOL
    assert max_rel(f.grad.cpu().numpy(), gf) < TOL


def test_backward_with_exported_warp_equals_recomputed_warp():
    """ABI v2: the backward may consume the x2_warp exported by the forward instead of re-evaluating it."""
    f1, f2, flow, rng = make_inputs(2, 24, 32, 48, seed=67)
    go = rng.standard_normal((2, 81, 32, 48)).astype(np.float32)
    grads = []
    for export in (False, True):
        a, b, f, g = to_dev(f1, f2, flow, go)
        for t in (a, b, f):
            t.requires_grad_()
        res = pkg.FusedWarpCorrelation(*REF_CFG, return_warped=export)(a, b, f)
        (res[0] if export else res).backward(g)
        grads.append([t.grad.clone() for t in (a, b, f)])
    for x, y in zip(*grads):
        assert max_rel(x.cpu().numpy(), y.cpu().numpy()) < 2e-6


def test_tma_backward_equals_plain_tiled_backward():
    f1, f2, flow, rng = make_inputs(2, 20, 40, 44, seed=71)
    go = rng.standard_normal((2, 81, 40, 44)).astype(np.float32)
    L = _lib.load()
    grads = []
    for disable in (0, 1):
        prev = L.pwc_set_disable_tma(disable)
        try:
            a, b, f, g = to_dev(f1, f2, flow, go)
            for t in (a, b, f):
                t.requires_grad_()
            out = pkg.FusedWarpCorrelation(*CANON_CFG, activation=True)(a, b, f)
            out.backward(g)
            grads.append([t.grad.clone() for t in (a, b, f)])
        finally:
            L.pwc_set_disable_tma(prev)
    for x, y in zip(*grads):
        assert max_rel(x.cpu().numpy(), y.cpu().numpy()) < 5e-6


def test_unaligned_buffers_through_the_c_abi():
    """Pointers that are only 4-byte aligned (views with a storage offset) must take the non-TMA kernels
    and still be exact: the C ABI promises nothing about alignment beyond float."""
    import ctypes
    f1, f2, flow, rng = make_inputs(2, 6, 16, 24, seed=73)
    go = rng.standard_normal((2, 81, 16, 24)).astype(np.float32)
    L = _lib.load()

    def shifted(a):      # device copy living at base + 4 bytes
        t = torch.from_numpy(a).to(dev())
        buf = torch.empty(t.numel() + 1, device=dev())
        v = buf[1:].view(t.shape)
        v.copy_(t)
        return v

    a, b, f, g = (shifted(x) for x in (f1, f2, flow, go))
    out = shifted(np.zeros((2, 81, 16, 24), np.float32))
    g1, g2, gf = (shifted(np.zeros_like(x)) for x in (f1, f2, flow))
    need = L.pwc_warpcorr_backward_workspace(2, 6, 16, 24, 1, 4, 1, 4, 1, 1)
    ws = torch.empty(need // 4 + 1, device=dev())[1:]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert all(t.data_ptr() % 16 == 4 for t in (a, b, f, g, out, g1, g2, gf))
    assert L.pwc_warpcorr_forward(p(a), p(b), p(f), p(out), None, 2, 6, 16, 24, 4, 1, 4, 1, 1, 0, 0.0, st)
    assert L.pwc_warpcorr_backward(p(g), p(a), p(b), p(f), None, None, p(g1), p(g2), p(gf), p(ws), ws.numel() * 4,
                                   2, 6, 16, 24, 4, 1, 4, 1, 1, 0, 0.0, st)
    torch.cuda.synchronize()
    ref = co.warpcorr_forward(f1, f2, flow, *CANON_CFG)
    r1, r2, rf = co.warpcorr_backward(go, f1, f2, flow, None, *CANON_CFG)
    assert max_rel(out.cpu().numpy(), ref) < TOL
    assert max_rel(g1.cpu().numpy(), r1) < TOL
    assert max_rel(g2.cpu().numpy(), r2) < TOL
    assert max_rel(gf.cpu().numpy(), rf) < TOL
    # bad arguments are refused with a message, not a crash
    assert not L.pwc_warpcorr_forward(None, p(b), p(f), p(out), None, 2, 6, 16, 24, 4, 1, 4, 1, 1, 0, 0.0, st)
    assert b"null pointer" in L.pwc_last_error()
    assert not L.pwc_warpcorr_backward(p(g), p(a), p(b), p(f), None, None, p(g1), p(g2), p(gf), None, 0,
                                       2, 6, 16, 24, 4, 1, 4, 1, 1, 0, 0.0, st)
    assert b"workspace" in L.pwc_last_error()


@pytest.mark.parametrize("shape", [(3, 16, 24, 28), (2, 12, 32, 48), (2, 40, 6, 7)])
def test_strided_output_into_concat_buffer(shape):
    """pwc_warpcorr_forward_strided: the cost volume lands inside [x1 | corr | flow] (model.py:89-91)."""
    B, C, H, W = shape
    f1, f2, flow, _ = make_inputs(B, C, H, W, seed=79)
    a, b, f = to_dev(f1, f2, flow)
    op = pkg.FusedWarpCorrelation(*REF_CFG, activation=True, return_warped=True)
    buf = torch.full((B, C + 81 + 2, H, W), float("nan"), device=dev())
    buf[:, :C] = a
    buf[:, C + 81:] = f
    with torch.no_grad():
        out, warped = op(a, b, f, out=buf[:, C:C + 81])
        dense, warped2 = op(a, b, f)
    assert out.data_ptr() == buf[:, C:C + 81].data_ptr()
    assert torch.equal(buf, torch.cat([a, dense, f], dim=1))
    assert torch.equal(warped, warped2)
    a.requires_grad_()
    with pytest.raises(RuntimeError, match="inference-only"):
        op(a, b, f, out=buf[:, C:C + 81])
    with pytest.raises(ValueError):
        with torch.no_grad():
            op(a, b, f, out=torch.empty(B, 81, H, W + 1, device=dev())[..., :W])


def test_random_shapes_fuzz():
    """Seeded random shapes through every dispatch path (TMA / plain tiled / cluster split / generic
    geometry is covered elsewhere): forward and all gradients against the oracle."""
    rng = np.random.Generator(np.random.PCG64(2026))
    for it in range(28):
        B = int(rng.integers(1, 4))
        C = int(rng.choice([1, 2, 3, 7, 8, 13, 32, 33, 70]))
        H = int(rng.integers(2, 41))
        W = int(rng.choice([2, 3, 5, 8, 12, 16, 20, 24, 31, 36, 44, 52])) if it % 3 else 4 * int(rng.integers(4, 14))
        cfg = REF_CFG if it % 2 else CANON_CFG
        act = bool(it % 4 == 1)
        use_flow = it % 5 != 0
        sigma = float(rng.choice([0.5, 2.0, 6.0]))
        f1, f2, flow, r2 = make_inputs(B, C, H, W, seed=1000 + it, flow_sigma=sigma,
                                       flow_kind="smooth" if it % 7 == 3 else "iid")
        go = r2.standard_normal((B, 81, H, W)).astype(np.float32)
        a, b, f, g = to_dev(f1, f2, flow if use_flow else None, go)
        for t in (a, b, f):
            if t is not None:
                t.requires_grad_()
        out = pkg.FusedWarpCorrelation(*cfg, activation=act, negative_slope=0.01)(a, b, f)
        out.backward(g)
        tag = (it, B, C, H, W, cfg, act, use_flow, sigma)
        ref = co.warpcorr_forward(f1, f2, flow if use_flow else None, *cfg, act=act, slope=0.01)
        assert max_rel(out.detach().cpu().numpy(), ref) < TOL, tag
        g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow if use_flow else None, out.detach().cpu().numpy(), *cfg,
                                          act=act, slope=0.01)
        assert max_rel(a.grad.cpu().numpy(), g1) < TOL, tag
        assert max_rel(b.grad.cpu().numpy(), g2) < TOL, tag
        if use_flow:
            assert max_rel(f.grad.cpu().numpy(), gf) < TOL, tag


def test_cuda_graph_capture_and_stream():
    """The entry points enqueue on the caller's current stream and are graph-capturable
    (no allocation, no sync inside the library)."""
    f1, f2, flow, _ = make_inputs(2, 16, 24, 28, seed=47)
    a, b, f = to_dev(f1, f2, flow)
    op = pkg.FusedWarpCorrelation(*REF_CFG, activation=True)
    eager = op(a, b, f)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            op(a, b, f)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        captured = op(a, b, f)
    captured.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(captured, eager)


SMALL_SHAPES = [(2, 196, 6, 7), (2, 192, 6, 8), (2, 128, 12, 14), (1, 128, 12, 16), (1, 5, 9, 11), (1, 17, 2, 2),
                (2, 64, 16, 16), (1, 3, 1, 1), (1, 260, 6, 8), (1, 33, 7, 16)]


@pytest.mark.parametrize("shape", SMALL_SHAPES)
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
@pytest.mark.parametrize("act", [False, True])
def test_small_image_kernels_vs_oracle_and_tiled(shape, cfg, act):
    """Coarse pyramid levels (H*W <= 256) run on the whole-image cluster kernels (one launch forward,
    one launch backward); they must agree with the oracle and with the tiled kernels that serve the
    same shapes when the small path is switched off."""
    B, C, H, W = shape
    f1, f2, flow, rng = make_inputs(B, C, H, W, seed=C + 7 * H + W)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    L = _lib.load()
    res = []
    for disable in (0, 1):
        prev = L.pwc_set_disable_small(disable)
        try:
            a, b, f, g = to_dev(f1, f2, flow, go)
            for t in (a, b, f):
                t.requires_grad_()
            before = L.pwc_launch_count()
            out, warped = pkg.FusedWarpCorrelation(*cfg, activation=act, return_warped=True)(a, b, f)
            out.backward(g)
            torch.cuda.synchronize()
            launches = L.pwc_launch_count() - before
            res.append([out.detach(), warped.detach(), a.grad.clone(), b.grad.clone(), f.grad.clone()])
        finally:
            L.pwc_set_disable_small(prev)
        if not disable:
            assert launches == 2, launches          # one forward + one backward launch, nothing else
    ref, ref_w = co.warpcorr_forward(f1, f2, flow, *cfg, act=act, slope=0.01, return_warped=True)
    gate = res[0][0].cpu().numpy()      # sign gate of the GPU's own forward (values near 0 may flip sign)
    g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow, gate, *cfg, act=act, slope=0.01)
    for got, want in zip(res[0], (ref, ref_w, g1, g2, gf)):
        assert max_rel(got.cpu().numpy(), want) < TOL
    for x, y in zip(*res):
        assert max_rel(x.cpu().numpy(), y.cpu().numpy()) < 5e-6


@pytest.mark.parametrize("shape", [(2, 196, 6, 7), (1, 40, 12, 14), (1, 3, 5, 5)])
@pytest.mark.parametrize("cfg", [REF_CFG, CANON_CFG])
def test_small_image_plain_correlation(shape, cfg):
    """Legacy Correlation (no warp) at the coarse levels: forward and both gradients in one launch each."""
    B, C, H, W = shape
    f1, f2, _, rng = make_inputs(B, C, H, W, seed=91)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    a, b, g = to_dev(f1, f2, go)
    a.requires_grad_(); b.requires_grad_()
    L = _lib.load()
    before = L.pwc_launch_count()
    out = pkg.Correlation(*cfg)(a, b)
    out.backward(g)
    torch.cuda.synchronize()
    assert L.pwc_launch_count() - before == 2
    assert max_rel(out.detach().cpu().numpy(), co.corr_forward(f1, f2, *cfg)) < TOL
    g1, g2 = co.corr_backward(go, f1, f2, *cfg)
    assert max_rel(a.grad.cpu().numpy(), g1) < TOL
    assert max_rel(b.grad.cpu().numpy(), g2) < TOL


@pytest.mark.parametrize("shape", [(150, 8, 16, 32), (3, 40, 40, 44), (2, 70, 24, 28)])
@pytest.mark.parametrize("act", [False, True])
def test_seq_backward_equals_slice_backward_multi_tile(shape, act):
    """stride2 == 1: the gradient w.r.t. f1 comes from corr_bwd_seq_kernel (threads own complete outputs);
    with it switched off the slice/reduce kernel computes the same thing.  (150, 8, 16, 32) gives every
    persistent CTA several tiles (the tap ring, the X double buffer and their release order matter
    there), 40 and 70 channels exercise partial and multiple 32-channel items."""
    B, C, H, W = shape
    f1, f2, flow, rng = make_inputs(B, C, H, W, seed=97)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    L = _lib.load()
    grads = []
    for disable in (0, 1):
        prev = L.pwc_set_disable_seq(disable)
        try:
            a, b, f, g = to_dev(f1, f2, flow, go)
            for t in (a, b, f):
                t.requires_grad_()
            out = pkg.FusedWarpCorrelation(*CANON_CFG, activation=act)(a, b, f)
            out.backward(g)
            torch.cuda.synchronize()
            grads.append([out.detach(), a.grad.clone(), b.grad.clone(), f.grad.clone()])
        finally:
            L.pwc_set_disable_seq(prev)
    for x, y in zip(*grads):
        assert max_rel(x.cpu().numpy(), y.cpu().numpy()) < 5e-6
    if B <= 3:
        gate = grads[0][0].cpu().numpy()
        g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow, gate, *CANON_CFG, act=act, slope=0.01)
        for got, want in zip(grads[0][1:], (g1, g2, gf)):
            assert max_rel(got.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("shape,cfg", [((40, 8, 32, 48), CANON_CFG), ((3, 20, 40, 44), REF_CFG), ((4, 196, 6, 7), CANON_CFG)])
def test_cuda_graph_replay_of_forward_and_backward(shape, cfg):
    """Forward + backward captured in one CUDA graph and replayed on fresh data: the backward chain uses
    programmatic dependent launches (zero fill and de-interleave start in the tails of the persistent
    kernels), which the graph must keep as edges that preserve the results."""
    B, C, H, W = shape
    f1, f2, flow, rng = make_inputs(B, C, H, W, seed=101)
    go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
    a, b, f, g = to_dev(f1, f2, flow, go)
    for t in (a, b, f):
        t.requires_grad_()
    op = pkg.FusedWarpCorrelation(*cfg, activation=True)

    def run():
        a.grad = b.grad = f.grad = None
        out = op(a, b, f)
        out.backward(g)
        return out

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            run()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_c = run()
    grads_c = [a.grad, b.grad, f.grad]
    # new inputs, same buffers
    f1n, f2n, flown, rng2 = make_inputs(B, C, H, W, seed=103)
    gon = rng2.standard_normal((B, 81, H, W)).astype(np.float32)
    with torch.no_grad():
        for dst, src in ((a, f1n), (b, f2n), (f, flown), (g, gon)):
            dst.copy_(torch.from_numpy(src))
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    got = [out_c.detach().clone()] + [t.clone() for t in grads_c]
    # eager run on the same data
    a2, b2, f2_, g2_ = to_dev(f1n, f2n, flown, gon)
    for t in (a2, b2, f2_):
        t.requires_grad_()
    out_e = op(a2, b2, f2_)
    out_e.backward(g2_)
    torch.cuda.synchronize()
    want = [out_e.detach(), a2.grad, b2.grad, f2_.grad]
    assert torch.equal(got[0], want[0])
    assert torch.equal(got[1], want[1])             # g1: no atomics anywhere on its path
    for x, y in zip(got[2:], want[2:]):             # scatter: fp32 reductions in a different order
        assert max_rel(x.cpu().numpy(), y.cpu().numpy()) < 2e-6
