"""GPU-vs-GPU check at the full level-2 size: backward through the TMA / seq kernels against the plain tiled
kernels (pwc_set_disable_tma), printing the positions of any mismatch.  usage: python scripts/check_backward_fullsize.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pwc_net_pytorch_b200 as pkg
from pwc_net_pytorch_b200 import _lib
L = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
for (B, C, H, W) in [(32, 32, 96, 112), (8, 32, 96, 112), (32, 8, 96, 112), (32, 20, 96, 112)]:
    f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
    flow = 2.0 * torch.randn(B, 2, H, W, device=dev); go = torch.randn(B, 81, H, W, device=dev)
    res = []
    for disable in (1, 0, 0):
        prev = L.pwc_set_disable_tma(disable)
        a = f1.clone().requires_grad_(); b = f2.clone().requires_grad_(); f = flow.clone().requires_grad_()
        pkg.FusedWarpCorrelation()(a, b, f).backward(go)
        torch.cuda.synchronize()
        L.pwc_set_disable_tma(prev)
        res.append((a.grad.clone(), b.grad.clone(), f.grad.clone()))
    for k in (1, 2):
        for name, x, y in zip(("g1", "g2", "gflow"), res[0], res[k]):
            err = (x - y).abs()
            bad = (err > 1e-4 * x.abs().max()).nonzero()
            print((B, C, H, W), "run", k, name, "maxerr %.3e" % (err.max().item() / x.abs().max().item()), "nbad", len(bad), bad[:6].tolist())
