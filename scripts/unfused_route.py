"""Times the zero-edit drop-in route: the reference's two modules kept separate --
WarpingLayer, then Correlation, then leaky_relu_ -- against the fused op, forward and forward + backward.
usage: python scripts/unfused_route.py [B C H W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pwc_net_pytorch_b200 as pkg  # noqa: E402
from pwc_net_pytorch_b200 import _lib  # noqa: E402

B, C, H, W = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (32, 32, 96, 112)
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
flow = 2.0 * torch.randn(B, 2, H, W, device=dev)
gout = torch.randn(B, 81, H, W, device=dev)
gw = torch.randn(B, C, H, W, device=dev)
warp, corr, fused = pkg.WarpingLayer(None), pkg.Correlation(4, 1, 4, 1, 1), pkg.FusedWarpCorrelation()
L = _lib.load()
for label, tma in (("tiled/TMA kernels", 0), ("plain kernels (TMA off)", 1)):
    L.pwc_set_disable_tma(tma)
    with torch.no_grad():
        t_w = timed(lambda: warp(f2, flow))
        t_c = timed(lambda: corr(f1, f2))
        t_u = timed(lambda: corr(f1, warp(f2, flow)))
        t_f = timed(lambda: fused(f1, f2, flow))
    a, b, c = f1.clone().requires_grad_(), f2.clone().requires_grad_(), flow.clone().requires_grad_()

    def wb():
        b.grad = c.grad = None
        warp(b, c).backward(gw)

    def ub():
        a.grad = b.grad = c.grad = None
        corr(a, warp(b, c)).backward(gout)

    def fb():
        a.grad = b.grad = c.grad = None
        fused(a, b, c).backward(gout)
    print(f"{label}: WarpingLayer fwd {t_w:.1f} us, fwd+bwd {timed(wb):.1f} us | Correlation fwd {t_c:.1f} us | "
          f"unfused route fwd {t_u:.1f}, fwd+bwd {timed(ub):.1f} us | fused fwd {t_f:.1f}, fwd+bwd {timed(fb):.1f} us")
L.pwc_set_disable_tma(0)
