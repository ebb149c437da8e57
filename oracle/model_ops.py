"""Model-level checkers -- TEST INFRASTRUCTURE ONLY: the hot-path operators of `Net` restated with the torch oracle
(swapped in through `Net(args, ops=...)`), shared by tests/, tests/golden/make_golden_model.py and the parity
assertions of bench.py's whole-network legs (the checker there, never the thing measured)."""
import zlib

import torch
import torch.nn.functional as F

from . import torch_ref as tr


def deterministic_init(module, seed=0):
    """Fills every parameter from a generator seeded by (seed, parameter name): identical values for any
    two modules with the same state_dict keys/shapes, without shipping a checkpoint."""
    with torch.no_grad():
        for name, p in module.state_dict().items():
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) % (2 ** 31))
            v = torch.randn(p.shape, generator=g, dtype=torch.float32)
            if name.endswith("weight") and p.dim() == 4:
                fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                v = v * (1.0 / fan_in) ** 0.5
            else:
                v = v * 0.05
            p.copy_(v.to(p.device))


class CostVolumeOps(torch.nn.Module):
    """What the reference Net computes with --corr CostVolumeLayer (model.py:20-22,80-84):
    WarpingLayer -> CostVolumeLayer -> optional leaky_relu_.  Returns (corr, x2_warp)."""

    def __init__(self, search_range=4, activation=False):
        super().__init__()
        self.search_range, self.activation = search_range, activation

    def forward(self, x1, x2, flow):
        w = tr.warp_ref(x2, flow)
        c = tr.cost_volume_layer_port(x1, w, self.search_range)
        return (F.leaky_relu(c, 0.01) if self.activation else c), w


class TorchCorrelationOps(torch.nn.Module):
    """Closed-form torch restatement of the default reference path (CUDA Correlation with
    pad = md = 2*sr+1, stride2 = 2, model.py:24) on the warped features."""

    def __init__(self, search_range=4, activation=False):
        super().__init__()
        self.sr, self.activation = search_range, activation

    def forward(self, x1, x2, flow):
        w = tr.warp_ref(x2, flow)
        md = 2 * self.sr + 1
        c = tr.corr_ref(x1, w, md, 1, md, 1, 2)
        return (F.leaky_relu(c, 0.01) if self.activation else c), w
