"""Generates tests/golden/ref_cuda.npz from the reference's OWN CUDA kernels on a B200.

The kernels are /root/reference/correlation_package/src/correlation_cuda_kernel.cu compiled
unchanged for sm_100a (oracle/build_ref.sh -> oracle/_ref/libref_corr.so, built in the container
that has /root/reference and shipped to the GPU box as a binary).  Run on the GPU box:

    gpurun -- python tests/golden/make_golden_gpu.py --out gpurun_out/ref_cuda.npz

then copy gpurun_out/ref_cuda.npz to tests/golden/ref_cuda.npz.  Inputs come from numpy PCG64
with fixed seeds so the file is reproducible.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_cuda  # noqa: E402

CASES = {
    # name: (B, C, H, W, (pad, k, md, s1, s2))
    "ref_cfg_tiny": (2, 5, 6, 7, (9, 1, 9, 1, 2)),
    "canon_cfg_tiny": (2, 5, 6, 7, (4, 1, 4, 1, 1)),
    "ref_cfg_c37": (1, 37, 12, 14, (9, 1, 9, 1, 2)),
    "canon_cfg_c37": (1, 37, 12, 14, (4, 1, 4, 1, 1)),
    "k3_s2_2": (1, 4, 9, 10, (4, 3, 4, 1, 2)),   # pad >= r*s2 keeps the reference bwd in bounds
    "stride1_2_fwd_only": (1, 4, 9, 10, (2, 1, 4, 2, 1)),
    "d5": (1, 6, 8, 9, (4, 1, 4, 1, 2)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_cuda.npz"))
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {}
    for name, (B, C, H, W, cfg) in CASES.items():
        rng = np.random.Generator(np.random.PCG64(sum(map(ord, name))))
        f1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
        f2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
        t1, t2 = torch.from_numpy(f1).to(dev), torch.from_numpy(f2).to(dev)
        o = ref_cuda.correlation_forward(t1, t2, *cfg)
        go = rng.standard_normal(tuple(o.shape)).astype(np.float32)
        out[f"{name}/cfg"] = np.asarray(cfg, np.int32)
        out[f"{name}/f1"], out[f"{name}/f2"], out[f"{name}/gout"] = f1, f2, go
        out[f"{name}/out"] = o.cpu().numpy()
        if cfg[3] == 1:
            g1, g2 = ref_cuda.correlation_backward(torch.from_numpy(go).to(dev), t1, t2, *cfg)
            out[f"{name}/g1"], out[f"{name}/g2"] = g1.cpu().numpy(), g2.cpu().numpy()
        torch.cuda.synchronize()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    np.savez_compressed(a.out, **out)
    print("wrote", a.out, sorted({k.split('/')[0] for k in out}))


if __name__ == "__main__":
    main()
