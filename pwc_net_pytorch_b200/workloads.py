"""Whole-network workloads around the hot path: BASELINE.json configs 3-5.

  multiscale_l1        the optimised quantity of the reference's MultiScale criterion
                       (losses.py:62-98: sum_l w_l * mean|o_l - t_l|, targets AvgPool(2^s)(gt) / 2^s
                       for the coarse outputs plus the full-resolution gt; weights main.py:84)
  TrainStep            config 4: forward + loss + backward + Adam (main.py:187-188, 217-246), one
                       process per GPU, DistributedDataParallel over whatever backend the process
                       group has (nccl on GPUs; the 2-rank test uses gloo) -- the gradient all-reduce of
                       the convolution weights is the only collective of the whole project
  PyramidInference     configs 3 and 5: full coarse-to-fine forward (model.py:72-113) of a batch of
                       image pairs, eager or replayed from one CUDA graph (batch-1 latency)

The hot path inside both is the fused CUDA operator (no CPU fallback); the convolution stacks are
plain cuDNN (out of scope).
"""
import contextlib

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .model import Net, default_args

LOSS_WEIGHTS = (0.32, 0.08, 0.02, 0.01, 0.005)      # main.py:84
ADAM = dict(lr=1e-4, weight_decay=4e-4)              # main.py:91,94,187-188


def multiscale_l1(flows, gt, weights=LOSS_WEIGHTS, num_levels=7):
    """losses.py:80-98 with norm 'L1' (losses.py:24,74): the coarse outputs l = 0.. are compared with
    AvgPool2d(2^(num_levels-l-1))(gt) / 2^(num_levels-l-1), the last output with gt itself."""
    n = len(flows)
    targets = [F.avg_pool2d(gt, 2 ** (num_levels - l - 1)) / 2 ** (num_levels - l - 1) for l in range(n - 1)] + [gt]
    return sum(w * (o - t).abs().mean() for w, o, t in zip(weights, flows, targets))


def unused_estimators(net):
    """FlowEstimator(Lv5) / (Lv6) are constructed (model.py:27-30) but never run when the loop breaks at
    output_level (model.py:101-108): their parameters get no gradient."""
    return [m for l, m in enumerate(net.flow_estimators) if l > net.args.output_level]


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class TrainStep:
    """One data-parallel training step of config 4 on this rank's image pairs.

    unused = 'find'   : DDP(find_unused_parameters=True)  -- the reference's modules as they are
             'freeze' : FlowEstimator(Lv5/Lv6).requires_grad_(False), plain DDP (no graph traversal,
                        smaller all-reduce)
    """

    def __init__(self, device, batch=8, height=384, width=448, unused="find", seed=0, args_over=None,
                 bucket_cap_mb=25, channels_last=False):
        self.device = torch.device(device)
        torch.manual_seed(seed)                       # identical initial weights on every rank
        self.net = Net(default_args(device=self.device, **(args_over or {}))).train()
        if channels_last:     # conv-stack policy: cuDNN's NHWC kernels, no per-conv layout conversions
            self.net.to(memory_format=torch.channels_last)
        self.channels_last = channels_last
        if unused == "freeze":
            for m in unused_estimators(self.net):
                m.requires_grad_(False)
        self.unused = unused
        self.world = _world()
        self.bucket_cap_mb = bucket_cap_mb
        if self.world > 1:
            ids = [self.device.index] if self.device.type == "cuda" else None
            self.model = torch.nn.parallel.DistributedDataParallel(
                self.net, device_ids=ids, find_unused_parameters=(unused == "find"),
                bucket_cap_mb=bucket_cap_mb, gradient_as_bucket_view=True)
        else:
            self.model = self.net
        self.opt = torch.optim.Adam([p for p in self.model.parameters() if p.requires_grad], **ADAM)
        rank = dist.get_rank() if self.world > 1 else 0
        g = torch.Generator(device=self.device).manual_seed(1000 + rank)      # each rank owns its own pairs
        self.x = torch.rand(batch, 3, 2, height, width, device=self.device, generator=g) * 255.0
        self.gt = torch.randn(batch, 2, height, width, device=self.device, generator=g) * 3.0
        self.batch = batch

    def grad_bytes(self):
        """fp32 bytes the all-reduce moves per step (parameters that receive a gradient)."""
        used = set(id(p) for p in self.net.parameters() if p.requires_grad)
        if self.unused == "find":
            for m in unused_estimators(self.net):
                used -= set(id(p) for p in m.parameters())
        return 4 * sum(p.numel() for p in self.net.parameters() if id(p) in used)

    def step(self, sync=True):
        """forward, loss, zero_grad, backward, optimizer step (main.py:217-246).  sync=False skips the
        gradient all-reduce (DDP.no_sync) -- used only to measure what the collective costs."""
        ctx = self.model.no_sync() if (not sync and self.world > 1) else contextlib.nullcontext()
        with ctx:
            flows, _ = self.model(self.x)
            loss = multiscale_l1(flows, self.gt)
            self.opt.zero_grad(set_to_none=True)
            loss.backward()
        self.opt.step()
        return loss.detach()


class PyramidInference:
    """Full-pyramid inference of `batch` synthetic image pairs of height x width (multiples of 64,
    SURVEY.md section 5: KITTI 375x1242 is padded to 384x1280 before the network)."""

    def __init__(self, device, batch, height, width, graphed=False, seed=0, net=None, args_over=None):
        if height % 64 or width % 64:
            raise ValueError("height and width must be multiples of 64 (six stride-2 stages, model.py:78)")
        self.device = torch.device(device)
        if net is None:
            torch.manual_seed(seed)
            net = Net(default_args(device=self.device, **(args_over or {}))).eval()
        self.net = net
        g = torch.Generator(device=self.device).manual_seed(2000 + seed)
        self.x = torch.rand(batch, 3, 2, height, width, device=self.device, generator=g) * 255.0
        self.batch = batch
        self.graph = None
        if graphed:
            from .graphed import GraphedForward
            self.graph = GraphedForward(net, self.x)

    def run(self):
        if self.graph is not None:
            return self.graph(self.x)
        with torch.no_grad():
            return self.net(self.x)
