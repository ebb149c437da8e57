import sys, os, torch
sys.path.insert(0, "/root/repo")
import pwc_net_pytorch_b200 as pkg
dev = torch.device("cuda:0")
def timed(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    flush = torch.empty(64 << 20, device=dev)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3
for (B, C, H, W) in [(32, 32, 96, 112), (32, 64, 48, 56), (32, 96, 24, 28)]:
    f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
    flow = 2.0 * torch.randn(B, 2, H, W, device=dev)
    op = pkg.FusedWarpCorrelation()
    warp = pkg.WarpingLayer(None)
    with torch.no_grad():
        t_f = timed(lambda: op(f1, f2, flow))
        t_n = timed(lambda: op(f1, f2, None))
        t_w = timed(lambda: warp(f2, flow))
    print(f"{(B,C,H,W)}: fused {t_f:.1f} us | no-flow corr {t_n:.1f} us | warp alone {t_w:.1f} us | sum {t_n+t_w:.1f}")
