#!/usr/bin/env python
"""BASELINE.json config 4: PWC-Net training step (multi-scale L1 loss, forward + backward through the
fused warp/correlation op) data-parallel over N GPUs with the NCCL gradient all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        examples/train_step_ddp.py [--batch 8] [--steps 10]

The hot path itself never communicates (it shards by image pair); the only collective is DDP's
bucketed all-reduce of the convolution weights' gradients (~17 MB fp32).
FlowEstimator(Lv5/Lv6) are constructed but unused at output_level 4 (model.py:27-30,101-108), hence
find_unused_parameters=True.  Loss = reference MultiScale (losses.py:62-98): sum_l w_l * mean|o_l - t_l|
with targets AvgPool(2^s)(gt)/2^s for s = 6..3 and the full-resolution gt; Adam lr 1e-4, wd 4e-4
(main.py:91,94,187-188).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pwc_net_pytorch_b200.model import Net, default_args  # noqa: E402

WEIGHTS = [0.32, 0.08, 0.02, 0.01, 0.005]   # main.py:84


def multiscale_l1(flows, gt):
    targets = [F.avg_pool2d(gt, 2 ** s) / 2 ** s for s in (6, 5, 4, 3)] + [gt]
    return sum(w * (o - t).abs().mean() for w, o, t in zip(WEIGHTS, flows, targets))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--height", type=int, default=384)
    ap.add_argument("--width", type=int, default=448)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                       # identical initial weights on every rank
    net = Net(default_args(device=dev)).train()
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True) \
        if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=4e-4)
    g = torch.Generator(device=dev).manual_seed(1 + rank)      # each rank owns its own image pairs
    x = torch.rand(a.batch, 3, 2, a.height, a.width, device=dev, generator=g) * 255.0
    gt = torch.randn(a.batch, 2, a.height, a.width, device=dev, generator=g) * 3.0
    losses = []

    def step():
        flows, _ = model(x)
        loss = multiscale_l1(flows, gt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss.detach()

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        losses.append(step())
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": f"PWC-Net training step, {a.batch} pairs/GPU {a.height}x{a.width}, DDP",
                          "n_gpus": world, "ms_per_step": float(ms) / a.steps,
                          "pairs_per_s": world * a.batch * a.steps / (float(ms) * 1e-3),
                          "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
                          "tf32_convs": bool(torch.backends.cudnn.allow_tf32)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
