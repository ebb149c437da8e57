"""Conv-stack policy probe: full forward at 384x448 under different cuDNN
policies; the hot path stays fp32.  usage: python scripts/conv_policy_probe.py [B]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwc_net_pytorch_b200.model import Net, default_args  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = Net(default_args(device=dev)).eval()
x = torch.rand(B, 3, 2, 384, 448, device=dev) * 255.0


def timed(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    ref, _ = net(x)
    ref = [f.clone() for f in ref]


def report(tag, fn):
    with torch.no_grad():
        out, _ = fn()
        epe = max(float(torch.norm(a.float() - b, p=2, dim=1).max()) for a, b in zip(out, ref))
        ms = timed(fn)
    print(f"{tag:48s} B={B}: {ms:8.2f} ms  {B / ms * 1e3:8.1f} pairs/s   max EPE vs tf32 eager {epe:.2e}", flush=True)


def plain():
    with torch.no_grad():
        return net(x)


def amp(dtype):
    def f():
        with torch.no_grad(), torch.autocast("cuda", dtype=dtype):
            return net(x)
    return f


for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    tag = "cudnn.benchmark" if bench else "default heuristics"
    net = net.to(memory_format=torch.contiguous_format)
    torch.backends.cudnn.allow_tf32 = True
    report(f"tf32 NCHW, {tag}", plain)
    net = net.to(memory_format=torch.channels_last)
    report(f"tf32 channels_last weights, {tag}", plain)
    report(f"bf16 autocast channels_last, {tag}", amp(torch.bfloat16))
    report(f"fp16 autocast channels_last, {tag}", amp(torch.float16))
    net = net.to(memory_format=torch.contiguous_format)
    report(f"bf16 autocast NCHW, {tag}", amp(torch.bfloat16))
