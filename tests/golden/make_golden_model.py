"""Generates tests/golden/ref_model.npz from the reference's OWN model.py (needs /root/reference).

The reference `Net` (model.py:11-115) is imported unmodified and run on the CPU with
--corr CostVolumeLayer (its PyTorch-level path; the default CUDA Correlation cannot run without a
GPU).  Weights are filled deterministically from parameter names (tests/model_util.py), so only the
input and the outputs are stored.  grid_sample is forced to align_corners=True (torch 0.4.0).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pwc_net_pytorch_b200 as pwc  # noqa: E402
from model_util import deterministic_init  # noqa: E402

pwc.install_as_reference_modules()          # model.py:8 imports correlation_package
sys.path.insert(0, os.environ.get("PWC_REFERENCE_ROOT", "/root/reference"))
import modules as ref_modules  # noqa: E402
import model as ref_model  # noqa: E402

_orig = F.grid_sample
ref_modules.F.grid_sample = lambda inp, grid, *a, **k: _orig(inp, grid, mode="bilinear", padding_mode="zeros",
                                                            align_corners=True)
import warnings  # noqa: E402
warnings.filterwarnings("ignore")

out = {}
for name, over in {"plain": {}, "act_residual": {"corr_activation": True, "residual": True}}.items():
    args = types.SimpleNamespace(search_range=4, device=torch.device("cpu"), rgb_max=255.0, residual=False,
                                 flow_norm=False, num_levels=7, lv_chs=[16, 32, 64, 96, 128, 192], output_level=4,
                                 batch_norm=False, corr="CostVolumeLayer", corr_activation=False, input_norm=True)
    for k, v in over.items():
        setattr(args, k, v)
    net = ref_model.Net(args).eval()
    deterministic_init(net, seed=1)
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, 256, (1, 3, 2, 128, 192), generator=g).float()   # >= 128 per side: the reference divides by (H-1)/2 at every level
    with torch.no_grad():
        flows, summ = net(x)
    out[f"{name}/x"] = x.numpy().astype(np.uint8)
    for i, f in enumerate(flows):
        out[f"{name}/flow{i}"] = f.numpy()
    out[f"{name}/warp_last"] = summ["x2_warps"][-1].numpy()
    out[f"{name}/keys"] = np.array(sorted(net.state_dict().keys()))
    out[f"{name}/shapes"] = np.array([str(tuple(net.state_dict()[k].shape)) for k in sorted(net.state_dict().keys())])
np.savez_compressed(os.path.join(HERE, "ref_model.npz"), **out)
print("wrote ref_model.npz", {k: getattr(v, "shape", None) for k, v in out.items() if "flow" in k})
