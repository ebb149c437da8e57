"""PWC-Net with the B200 hot path in its pyramid loop (SURVEY.md section 8 rows a11 and f).

Same architecture, argument Namespace and state_dict keys as the reference `Net` (model.py:11-115,
modules.py:77-152), so reference checkpoints load unchanged; the only functional difference is that
the three hot statements of the coarse-to-fine loop (model.py:80-84: WarpingLayer -> Correlation ->
leaky_relu_) are one fused CUDA launch.  The convolution stacks stay plain `nn.Conv2d` (cuDNN): they
are out of scope for this path.

    args = default_args(device='cuda')           # same fields main.py builds (main.py:42-78)
    net = Net(args).eval()
    flows, summaries = net(x)                    # x: [B, 3, 2, H, W], H and W multiples of 64

`ops` lets a caller (the parity tests) swap the hot-path operators for another implementation
with the same call signature; the default is the CUDA path and there is no CPU fallback.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modules import FusedWarpCorrelation, FusedWarpCostVolume


def default_args(**overrides):
    """The fields of the reference's argparse Namespace that the model reads
    (main.py:42-46,53-58,71-78)."""
    args = SimpleNamespace(search_range=4, device="cuda", rgb_max=255.0, residual=False, flow_norm=False,
                           num_levels=7, lv_chs=[16, 32, 64, 96, 128, 192], output_level=4, batch_norm=False,
                           corr="cost_volume", corr_activation=False, input_norm=True)
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


def _conv_block(batch_norm, cin, cout, stride=1):
    """3x3 convolution + LeakyReLU(0.1), optionally with BatchNorm (modules.py:11-22)."""
    layers = [nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=not batch_norm)]
    if batch_norm:
        layers.append(nn.BatchNorm2d(cout))
    layers.append(nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(*layers)


class FeaturePyramidExtractor(nn.Module):
    """modules.py:77-98: num_levels-1 stages of (stride-2 conv, conv); returned coarse -> fine."""

    def __init__(self, args):
        super().__init__()
        self.convs = []
        cin = 3
        for l, ch in enumerate(args.lv_chs[:args.num_levels - 1]):
            stage = nn.Sequential(_conv_block(args.batch_norm, cin, ch, stride=2),
                                  _conv_block(args.batch_norm, ch, ch))
            self.add_module(f"Feature(Lv{l + 1})", stage)
            self.convs.append(stage)
            cin = ch

    def forward(self, x):
        feats = []
        for stage in self.convs:
            x = stage(x)
            feats.append(x)
        return feats[::-1]


class OpticalFlowEstimator(nn.Module):
    """modules.py:101-124."""

    def __init__(self, args, ch_in):
        super().__init__()
        self.flow_norm = args.flow_norm
        bn = args.batch_norm
        self.convs = nn.Sequential(_conv_block(bn, ch_in, 128), _conv_block(bn, 128, 128), _conv_block(bn, 128, 96),
                                   _conv_block(bn, 96, 64), _conv_block(bn, 64, 32),
                                   nn.Conv2d(32, 2, 3, padding=1))

    def forward(self, x):
        y = self.convs(x)
        if self.flow_norm:   # modules.py:116-121 (both channels scaled by the width, as in the reference)
            y = torch.tanh(y) * ((x.size(3) - 1.0) / 2.0)
        return y


class ContextNetwork(nn.Module):
    """modules.py:127-152: dilated 3x3 stack 128-128-128-96-64-32-2, LeakyReLU(0.01)."""

    def __init__(self, args, ch_in):
        super().__init__()
        layers, cin = [], ch_in
        for cout, dil in ((128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1)):
            layers += [nn.Conv2d(cin, cout, 3, padding=dil, dilation=dil), nn.LeakyReLU(inplace=True)]
            cin = cout
        layers.append(nn.Conv2d(cin, 2, 3, padding=1))
        self.convs = nn.Sequential(*layers)

    def forward(self, x):
        return self.convs(x)


class Net(nn.Module):
    """Reference `Net` (model.py:11-115) with the fused warp + cost-volume operator."""

    def __init__(self, args, ops=None):
        super().__init__()
        self.args = args
        self.feature_pyramid_extractor = FeaturePyramidExtractor(args)
        # model.py:24: Correlation(pad = md = 2*search_range+1, kernel 1, stride1 1, stride2 2); the
        # optional leaky_relu_ of model.py:84 (slope 0.01) is the kernel's epilogue
        # model.py:21-22: `--corr CostVolumeLayer` selects the pure-PyTorch layer, which samples +-search_range
        # at stride 1, orders its channels axis-first and divides by 81 -- a different volume from the CUDA
        # op's, so a checkpoint trained with it needs exactly that volume
        cost_volume_layer = getattr(args, "corr", None) == "CostVolumeLayer"
        if ops is not None:
            self.warp_corr = ops
        elif cost_volume_layer:
            self.warp_corr = FusedWarpCostVolume(args.search_range, activation=bool(args.corr_activation),
                                                 negative_slope=0.01, return_warped=True)
        else:
            self.warp_corr = FusedWarpCorrelation.from_search_range(
                args.search_range, activation=bool(args.corr_activation), negative_slope=0.01, return_warped=True)
        self._own_ops = ops is None
        self._direct_concat = ops is None and args.search_range == 4 and not cost_volume_layer
        self._fold_upsample = True      # SURVEY.md section 8 row f1 (inference; switched off by the parity tests)
        self.flow_estimators = []
        for l, ch in enumerate(args.lv_chs[::-1] + [3]):
            est = OpticalFlowEstimator(args, ch + (args.search_range * 2 + 1) ** 2 + 2)
            self.add_module(f"FlowEstimator(Lv{l})", est)
            self.flow_estimators.append(est)
        self.context_network = ContextNetwork(args, 3 + 2)
        for m in self.modules():   # model.py:39-46
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                if m.bias is not None:
                    nn.init.uniform_(m.bias)
                nn.init.xavier_uniform_(m.weight)
        self.to(args.device)

    def forward(self, x):
        args = self.args
        if args.input_norm:   # model.py:51-53
            mean = x.contiguous().view(x.size()[:2] + (-1,)).mean(dim=-1).view(x.size()[:2] + (1, 1, 1))
            x = (x - mean) / args.rgb_max
        x1_raw = x[:, :, 0].contiguous()
        x2_raw = x[:, :, 1].contiguous()
        pyr1 = self.feature_pyramid_extractor(x1_raw) + [x1_raw]
        pyr2 = self.feature_pyramid_extractor(x2_raw) + [x2_raw]

        flows, warps = [], []
        flow = None
        for l, (x1, x2) in enumerate(zip(pyr1, pyr2)):
            if (l > 0 and self._direct_concat and self._fold_upsample and not torch.is_grad_enabled()
                    and x1.dtype == torch.float32 and flow.dtype == torch.float32):
                # inference: model.py:78-91 without the upsample / multiply / cat launches -- the kernel takes
                # the previous level's flow, upsamples it itself (bit for bit F.interpolate(...) * 2) and writes
                # both the fine flow and the cost volume into the estimator's input [x1 | corr | flow]
                C = x1.size(1)
                est_in = torch.empty((x1.size(0), C + 81 + 2, x1.size(2), x1.size(3)), dtype=x1.dtype, device=x1.device)
                est_in[:, :C] = x1
                _, x2_warp = self.warp_corr(x1, x2, None, out=est_in[:, C:C + 81], coarse_flow=flow,
                                            flow_out=est_in[:, C + 81:])
                flow = est_in[:, C + 81:]
                flow_coarse = self.flow_estimators[l](est_in)
                if args.residual:
                    flow_coarse = flow_coarse + flow
                if l == args.output_level:   # model.py:101-108
                    scale = 2 ** (args.num_levels - args.output_level - 1)
                    flow = F.interpolate(flow_coarse, scale_factor=scale, mode="bilinear", align_corners=False) * scale
                    flow = flow + self.context_network(torch.cat([pyr1[-1], flow], dim=1))
                    flows.append(flow)
                    warps.append(x2_warp.detach())
                    break
                flow = flow_coarse
                flows.append(flow)
                warps.append(x2_warp.detach())
                continue
            if l == 0:   # model.py:74-76: zero flow at the coarsest level (the warp is then the identity)
                flow = torch.zeros((x1.size(0), 2, x1.size(2), x1.size(3)), dtype=x1.dtype, device=x1.device)
            else:        # model.py:78: F.upsample(..., 'bilinear') == align_corners=False
                flow = F.interpolate(flow, scale_factor=2, mode="bilinear", align_corners=False) * 2
            # ---- hot path (model.py:80-84): one fused launch ----
            # l == 0: the flow is identically zero, the warp is the identity (model.py:74-76 still runs it);
            # the kernel is told so and skips taps, window and gather (x2_warp is then x2 itself, bit for bit)
            wflow = None if (l == 0 and self._own_ops) else flow
            if x1.dtype != torch.float32:      # reduced-precision convolutions (autocast): the hot path is fp32
                x1, x2 = x1.float(), x2.float()
                flow = flow.float()
                wflow = None if wflow is None else flow
            if self._direct_concat and not torch.is_grad_enabled():
                # inference: the kernel writes the cost volume straight into the estimator's input
                # [x1 | corr | flow] (model.py:89-91); torch.cat would copy its 81 channels once more
                C = x1.size(1)
                est_in = torch.empty((x1.size(0), C + 81 + 2, x1.size(2), x1.size(3)), dtype=x1.dtype, device=x1.device)
                est_in[:, :C] = x1
                est_in[:, C + 81:] = flow
                _, x2_warp = self.warp_corr(x1, x2, wflow, out=est_in[:, C:C + 81])
            elif self._direct_concat:
                # training: same buffer, recorded by autograd (the backward reads the gradient of the corr
                # slice in place) -- no torch.cat copy and no slice copy of the 81 channels
                est_in, x2_warp = self.warp_corr(x1, x2, wflow, concat=flow)
            else:
                corr, x2_warp = self.warp_corr(x1, x2, wflow)
                est_in = torch.cat([x1, corr, flow], dim=1)
            flow_coarse = self.flow_estimators[l](est_in)
            if args.residual:
                flow_coarse = flow_coarse + flow
            if l == args.output_level:   # model.py:101-108
                scale = 2 ** (args.num_levels - args.output_level - 1)
                flow = F.interpolate(flow_coarse, scale_factor=scale, mode="bilinear", align_corners=False) * scale
                flow = flow + self.context_network(torch.cat([pyr1[-1], flow], dim=1))
                flows.append(flow)
                warps.append(x2_warp.detach())
                break
            flow = flow_coarse
            flows.append(flow)
            warps.append(x2_warp.detach())
        return flows, {"x2_warps": warps}
