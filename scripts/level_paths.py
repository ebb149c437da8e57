"""Times the forward / backward of one pyramid level under each kernel path the library can take (development aid).
usage: python scripts/level_paths.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pwc_net_pytorch_b200 as pkg  # noqa: E402
from pwc_net_pytorch_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
L = _lib.load()
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


for name, (C, H, W) in {"L2": (32, 96, 112), "L3": (64, 48, 56), "L4": (96, 24, 28), "L5": (128, 12, 14), "L6": (196, 6, 7)}.items():
    torch.manual_seed(0)
    f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
    flow = 2.0 * torch.randn(B, 2, H, W, device=dev)
    gout = torch.randn(B, 81, H, W, device=dev)
    op = pkg.FusedWarpCorrelation()
    row = []
    for label, tma, small in (("default", 0, 0), ("no-tma", 1, 0), ("no-small", 0, 1), ("no-tma,no-small", 1, 1)):
        L.pwc_set_disable_tma(tma); L.pwc_set_disable_small(small)
        with torch.no_grad():
            tf = timed(lambda: op(f1, f2, flow))
        a, b, c = f1.clone().requires_grad_(), f2.clone().requires_grad_(), flow.clone().requires_grad_()

        def fb():
            a.grad = b.grad = c.grad = None
            op(a, b, c).backward(gout)
        tfb = timed(fb)
        row.append(f"{label}: fwd {tf:6.1f} fwd+bwd {tfb:6.1f}")
    L.pwc_set_disable_tma(0); L.pwc_set_disable_small(0)
    print(f"{name} B={B} C={C} {H}x{W} | " + " | ".join(row))
