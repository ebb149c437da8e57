"""Oracle check of shapes that give every persistent CTA several tiles (ring / release-order bugs show up
only there).  usage: python scripts/check_backward_multitile.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pwc_net_pytorch_b200 as pkg
from oracle import c_oracle as co
from util import CANON_CFG, REF_CFG, make_inputs, max_rel
dev = torch.device("cuda:0")
for shape in [(80, 4, 16, 32), (40, 8, 32, 48), (150, 4, 16, 32)]:
    for cfg in (CANON_CFG, REF_CFG):
        B, C, H, W = shape
        f1, f2, flow, rng = make_inputs(B, C, H, W, seed=3)
        go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
        ref = co.warpcorr_forward(f1, f2, flow, *cfg)
        g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow, ref, *cfg)
        for rep in range(3):
            a, b, f, g = [torch.from_numpy(x).to(dev) for x in (f1, f2, flow, go)]
            for t in (a, b, f): t.requires_grad_()
            out = pkg.FusedWarpCorrelation(*cfg)(a, b, f)
            out.backward(g)
            e1 = np.abs(a.grad.cpu().numpy() - g1).reshape(B, -1).max(1) / np.abs(g1).max()
            e2 = np.abs(b.grad.cpu().numpy() - g2).reshape(B, -1).max(1) / np.abs(g2).max()
            print(shape, cfg, "out %.2e g1 %.2e g2 %.2e gflow %.2e" % (max_rel(out.detach().cpu().numpy(), ref),
                  e1.max(), e2.max(), max_rel(f.grad.cpu().numpy(), gf)), "bad g1 imgs", np.nonzero(e1 > 1e-5)[0][:12], "bad g2 imgs", np.nonzero(e2 > 1e-5)[0][:12])
