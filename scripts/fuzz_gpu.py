"""Extended seeded fuzz of every dispatch path against the C oracle (one-off robustness run, larger than
tests/test_gpu_parity.py::test_random_shapes_fuzz).  usage: python scripts/fuzz_gpu.py [iterations] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import pwc_net_pytorch_b200 as pkg
from oracle import c_oracle as co
from util import CANON_CFG, REF_CFG, make_inputs, max_rel

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device("cuda:0")
rng = np.random.Generator(np.random.PCG64(seed))
worst = 0.0
for it in range(iters):
    big = it % 6 == 0
    B = int(rng.integers(20, 60)) if big else int(rng.integers(1, 5))
    C = int(rng.choice([1, 3, 4, 8, 13, 31, 32, 33, 40, 64, 70, 96]))
    if big:
        C = int(rng.choice([4, 8, 12, 33]))
    H = int(rng.integers(1, 72 if not big else 40))
    W = 4 * int(rng.integers(1, 18)) if it % 3 else int(rng.integers(1, 70))
    cfg = REF_CFG if it % 2 else CANON_CFG
    act = it % 4 == 1
    use_flow = it % 5 != 0
    sigma = float(rng.choice([0.3, 2.0, 5.0, 12.0]))
    f1, f2, flow, r2 = make_inputs(B, C, H, W, seed=5000 + it, flow_sigma=sigma, flow_kind="smooth" if it % 7 == 3 else "iid")
    go = r2.standard_normal((B, 81, H, W)).astype(np.float32)
    ts = [None if x is None else torch.from_numpy(x).to(dev) for x in (f1, f2, flow if use_flow else None, go)]
    a, b, f, g = ts
    for t in (a, b, f):
        if t is not None:
            t.requires_grad_()
    out = pkg.FusedWarpCorrelation(*cfg, activation=act, negative_slope=0.01)(a, b, f)
    out.backward(g)
    ref = co.warpcorr_forward(f1, f2, flow if use_flow else None, *cfg, act=act, slope=0.01)
    g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow if use_flow else None, out.detach().cpu().numpy(), *cfg, act=act, slope=0.01)
    errs = [max_rel(out.detach().cpu().numpy(), ref), max_rel(a.grad.cpu().numpy(), g1), max_rel(b.grad.cpu().numpy(), g2)]
    if use_flow:
        errs.append(max_rel(f.grad.cpu().numpy(), gf))
    worst = max(worst, max(errs))
    if max(errs) >= 1e-5:
        print("FAIL", (it, B, C, H, W, cfg, act, use_flow, sigma), errs)
        sys.exit(1)
print(f"fuzz ok: {iters} cases, worst relative error {worst:.2e}")
