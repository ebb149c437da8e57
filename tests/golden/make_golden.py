"""Generates tests/golden/ref_python.npz from the reference's OWN Python modules.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What it pins (the reference ships no tests or golden vectors of its own):
  * modules.WarpingLayer (modules.py:25-42) imported unmodified.  The only intervention is that
    F.grid_sample is called with align_corners=True, which is what torch 0.4.0
    (requirements.txt:62) did and what the reference was written against (SURVEY.md section 0
    fact 3); torch 2.x changed the default.
  * its backward: gradients w.r.t. x and flow from torch autograd through that unmodified forward.
  * modules.CostVolumeLayer (modules.py:45-74) imported unmodified; it equals the CUDA
    Correlation(pad=4, kernel=1, md=4, stride1=1, stride2=1) up to the channel permutation of
    SURVEY.md appendix B and the factor C/81, which pins the channel order / displacement sign of
    the correlation oracle against reference-authored code.
Inputs are drawn from numpy's PCG64 with fixed seeds, so the file is reproducible.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("PWC_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_modules():
    sys.path.insert(0, REF)
    import modules as ref_modules  # noqa: E402  (reference's modules.py)
    return ref_modules


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "ref_python.npz"))
    a = ap.parse_args()
    ref_modules = load_reference_modules()

    orig_grid_sample = F.grid_sample

    def grid_sample_040(inp, grid, *args, **kw):
        kw.setdefault("mode", "bilinear")
        kw.setdefault("padding_mode", "zeros")
        kw["align_corners"] = True
        return orig_grid_sample(inp, grid, *args, **kw)

    ref_modules.F.grid_sample = grid_sample_040
    args = types.SimpleNamespace(device=torch.device("cpu"), search_range=4)
    warp = ref_modules.WarpingLayer(args)
    cvl = ref_modules.CostVolumeLayer(args)

    out = {}
    cases = {
        # name: (B, C, H, W, flow_sigma)
        "tiny": (2, 5, 6, 7, 1.5),
        "lvl6": (2, 20, 6, 7, 2.0),
        "mid": (1, 12, 24, 28, 2.0),
        "big_flow": (1, 3, 10, 12, 8.0),
    }
    for name, (B, C, H, W, sig) in cases.items():
        rng = np.random.Generator(np.random.PCG64(sum(map(ord, name))))
        f1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
        f2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
        flow = (sig * rng.standard_normal((B, 2, H, W))).astype(np.float32)
        with torch.no_grad():
            t1, t2, tf = map(torch.from_numpy, (f1, f2, flow))
            w = warp(t2, tf)
            cv_plain = cvl(t1, t2)
            cv_warp = cvl(t1, w)
        out[f"{name}/f1"] = f1
        out[f"{name}/f2"] = f2
        out[f"{name}/flow"] = flow
        out[f"{name}/warp"] = w.numpy()
        out[f"{name}/costvolume_plain"] = cv_plain.numpy()
        out[f"{name}/costvolume_of_warp"] = cv_warp.numpy()
        # autograd of the reference's own WarpingLayer.forward (grid_sample backward + the gradient of
        # modules.py:36-40): pins the warp *backward*  to reference code
        gw = rng.standard_normal((B, C, H, W)).astype(np.float32)
        x2 = torch.from_numpy(f2).clone().requires_grad_()
        fl = torch.from_numpy(flow).clone().requires_grad_()
        warp(x2, fl).backward(torch.from_numpy(gw))
        out[f"{name}/warp_gout"] = gw
        out[f"{name}/warp_gx"] = x2.grad.numpy()
        out[f"{name}/warp_gflow"] = fl.grad.numpy()
    # zero flow must be the identity under the 0.4.0 semantics
    z = torch.zeros(1, 2, 6, 7)
    x = torch.from_numpy(out["tiny/f2"][:1])
    out["tiny/warp_zero_flow_maxdiff"] = np.float32((warp(x, z) - x).abs().max().item())
    np.savez_compressed(a.out, **out)
    print("wrote", a.out, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
