"""Tiny driver for ncu captures: runs one op at one shape a few times.
usage: python scripts/prof_case.py fwd|bwd|fwdbwd level2|level6|B,C,H,W [iid|smooth] [canon|ref] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pwc_net_pytorch_b200 as pkg  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
shape = sys.argv[2] if len(sys.argv) > 2 else "level2"
kind = sys.argv[3] if len(sys.argv) > 3 else "iid"
cfgname = sys.argv[4] if len(sys.argv) > 4 else "canon"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
B, C, H, W = {"level2": (32, 32, 96, 112), "level6": (32, 196, 6, 7)}.get(shape) or tuple(map(int, shape.split(",")))
dev = torch.device("cuda:0")
torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device=dev)
f2 = torch.randn(B, C, H, W, device=dev)
if kind == "iid":
    flow = 2.0 * torch.randn(B, 2, H, W, device=dev)
else:
    coarse = 2.0 * torch.randn(B, 2, max(2, H // 8 + 1), max(2, W // 8 + 1), device=dev)
    flow = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True).contiguous()
gout = torch.randn(B, 81, H, W, device=dev)
op = pkg.FusedWarpCorrelation() if cfgname == "canon" else pkg.FusedWarpCorrelation.from_search_range(4)
if os.environ.get("PWC_FORCE_GENERIC"):
    from pwc_net_pytorch_b200 import _lib
    _lib.load().pwc_set_force_generic(1)
if what != "fwd":
    f1.requires_grad_(); f2.requires_grad_(); flow.requires_grad_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(iters):
    if it == iters - 1:
        e0.record()
    if what == "fwd":
        with torch.no_grad():
            out = op(f1, f2, flow)
    else:
        f1.grad = f2.grad = flow.grad = None
        out = op(f1, f2, flow)
        out.backward(gout)
e1.record()
torch.cuda.synchronize()
print(f"{what} {shape} {kind} {cfgname}: last iter {e0.elapsed_time(e1):.4f} ms")
