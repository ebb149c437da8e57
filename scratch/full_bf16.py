import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pwc_net_pytorch_b200.model import Net, default_args
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
net = Net(default_args(device=dev)).eval()
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for Bf in (1, 16, 64):
    xin = torch.rand(Bf, 3, 2, 384, 448, device=dev) * 255.0
    with torch.no_grad():
        ref = net(xin)[0][-1]
        ms = timeit(lambda: net(xin))
        print("tf32 NCHW", Bf, ms, Bf / ms * 1e3)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms = timeit(lambda: net(xin))
            o = net(xin)[0][-1]
        print("bf16 NCHW", Bf, ms, Bf / ms * 1e3, "EPE vs tf32", (o.float() - ref).norm(dim=1).mean().item())
    net_cl = net.to(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ms = timeit(lambda: net_cl(xin))
        o = net_cl(xin)[0][-1]
    print("bf16 channels_last", Bf, ms, Bf / ms * 1e3, "EPE vs tf32", (o.float() - ref).norm(dim=1).mean().item())
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ms = timeit(lambda: net_cl(xin))
    print("fp16 channels_last", Bf, ms, Bf / ms * 1e3)
    net = net.to(memory_format=torch.contiguous_format)
