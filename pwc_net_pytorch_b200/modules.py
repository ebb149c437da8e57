"""Host-side mirror of the hot-path modules of the reference's modules.py.

  WarpingLayer(args).forward(x, flow)      modules.py:25-42 -- same constructor and call
  FusedWarpCorrelation(...)(x1, x2, flow)  model.py:80-84 in one launch (new entry point)
"""
import torch
import torch.nn as nn

from . import functional as PF


class WarpingLayer(nn.Module):
    """Backward warp of `x` by `flow` (pixel units of this level; channel 0 horizontal, channel 1
    vertical): out[n,c,y,x] = bilinear_zero(x[n,c], x+u, y+v), i.e. modules.py:31-42 with
    torch-0.4.0 grid_sample semantics.  `args` is kept for signature compatibility
    (modules.py:27-29); only the tensors' own device is used -- no CPU grid is built or copied
    per call (utils.py:3-7, modules.py:40)."""

    def __init__(self, args=None):
        super(WarpingLayer, self).__init__()
        self.args = args

    def forward(self, x, flow):
        return PF.warp(x, flow)


class FusedWarpCorrelation(nn.Module):
    """corr = [leaky_relu_]( Correlation(x1, WarpingLayer(x2, flow)) ) without materialising the
    warped features.  Defaults are the canonical PWC-Net cost volume (81 displacements, +-4 px);
    `FusedWarpCorrelation.from_search_range(4)` gives the literal reference configuration
    (model.py:24: pad 9, md 9, stride2 2 -> displacements {-8,-6,...,8})."""

    def __init__(self, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                 corr_multiply=1, activation=False, negative_slope=0.01, return_warped=False):
        super(FusedWarpCorrelation, self).__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply
        self.activation = activation
        self.negative_slope = negative_slope
        self.return_warped = return_warped

    @classmethod
    def from_search_range(cls, search_range, **kw):
        return cls(pad_size=search_range * 2 + 1, kernel_size=1,
                   max_displacement=search_range * 2 + 1, stride1=1, stride2=2, **kw)

    def forward(self, x1, x2, flow=None, out=None, coarse_flow=None, flow_out=None, concat=None):
        """`out` (inference only): a [B, 81, H, W] view to write into, e.g. the channel slice of the flow
        estimator's concatenated input; see functional.warp_correlation_into.
        `coarse_flow` + `flow_out` (inference only, with `out`): the flow at the previous pyramid level; the
        kernel upsamples it itself (model.py:78) and writes the fine flow to `flow_out`; see
        functional.warp_correlation_coarse_into."""
        if concat is not None:
            # training path of model.py:89-91: returns (cat([x1, corr, flow]), x2_warp); `concat` is the tensor
            # that fills the last two channels (the flow itself, or the zero flow of level 0 when flow is None)
            return PF.warp_correlation_concat(x1, x2, flow, concat if flow is None else None, self.pad_size,
                                              self.kernel_size, self.max_displacement, self.stride1, self.stride2,
                                              self.activation, self.negative_slope, self.return_warped)
        if coarse_flow is not None:
            if out is None or flow_out is None or flow is not None:
                raise ValueError("coarse_flow needs out= and flow_out= views and no flow=")
            return PF.warp_correlation_coarse_into(out, flow_out, x1, x2, coarse_flow, self.pad_size, self.kernel_size,
                                                   self.max_displacement, self.stride1, self.stride2,
                                                   self.activation, self.negative_slope, self.return_warped)
        if out is not None:
            return PF.warp_correlation_into(out, x1, x2, flow, self.pad_size, self.kernel_size,
                                            self.max_displacement, self.stride1, self.stride2,
                                            self.activation, self.negative_slope, self.return_warped)
        return PF.warp_correlation(x1, x2, flow, self.pad_size, self.kernel_size,
                                   self.max_displacement, self.stride1, self.stride2,
                                   self.activation, self.negative_slope, self.return_warped)


def cost_volume_channel_order(search_range):
    """Raster channel tc = (dy + r) * D + (dx + r) of `Correlation(pad r, md r, stride2 1)` that holds
    channel I of the reference's pure-PyTorch `CostVolumeLayer` (modules.py:58-72): (0, 0) first, then for
    i = 1..r the axis displacements (-i,0), (+i,0), (0,-i), (0,+i) followed, for j = 1..r, by the diagonal
    ones (-i,-j), (+i,+j), (-i,+j), (+i,-j)."""
    r = search_range
    D = 2 * r + 1
    order = [(0, 0)]
    for i in range(1, r + 1):
        order += [(-i, 0), (i, 0), (0, -i), (0, i)]
        for j in range(1, r + 1):
            order += [(-i, -j), (i, j), (-i, j), (i, -j)]
    return [(dy + r) * D + (dx + r) for dy, dx in order]


class FusedWarpCostVolume(nn.Module):
    """`CostVolumeLayer(x1, WarpingLayer(x2, flow))` of the reference (modules.py:45-74, selected by
    `--corr CostVolumeLayer`, model.py:21-22) on the same fused kernel: displacements +-search_range at
    stride 1, the layer's own channel order, and division by D*D instead of by C (SURVEY.md section 0
    fact 5).  The LeakyReLU of model.py:84 commutes with the positive rescale, so it stays in the kernel's
    epilogue; the permutation + rescale is one gather pass over the 81 channels."""

    def __init__(self, search_range=4, activation=False, negative_slope=0.01, return_warped=False):
        super(FusedWarpCostVolume, self).__init__()
        self.search_range = search_range
        self.return_warped = return_warped
        self.op = FusedWarpCorrelation(pad_size=search_range, kernel_size=1, max_displacement=search_range,
                                       stride1=1, stride2=1, activation=activation,
                                       negative_slope=negative_slope, return_warped=return_warped)
        self.register_buffer("order", torch.tensor(cost_volume_channel_order(search_range), dtype=torch.long),
                             persistent=False)

    def forward(self, x1, x2, flow=None):
        res = self.op(x1, x2, flow)
        corr, warped = res if self.return_warped else (res, None)
        D2 = (2 * self.search_range + 1) ** 2
        corr = corr.index_select(1, self.order.to(corr.device)) * (float(x1.size(1)) / D2)
        return (corr, warped) if self.return_warped else corr
