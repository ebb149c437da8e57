import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import pwc_net_pytorch_b200 as pkg
from oracle import c_oracle as co
from util import CANON_CFG, REF_CFG, make_inputs, max_rel
dev = torch.device("cuda:0")
for shape in [(1, 4, 6, 7), (2, 192, 6, 7), (1, 8, 3, 3), (1,4,1,5), (1,4,5,1)]:
    for cfg in (REF_CFG, CANON_CFG):
        B, C, H, W = shape
        f1, f2, flow, rng = make_inputs(B, C, H, W, seed=3)
        go = rng.standard_normal((B, 81, H, W)).astype(np.float32)
        a, b, f, g = [torch.from_numpy(x).to(dev) for x in (f1, f2, flow, go)]
        for t in (a, b, f): t.requires_grad_()
        out = pkg.FusedWarpCorrelation(*cfg)(a, b, f)
        out.backward(g)
        ref = co.warpcorr_forward(f1, f2, flow, *cfg)
        g1, g2, gf = co.warpcorr_backward(go, f1, f2, flow, ref, *cfg)
        print(shape, cfg, "out %.2e g1 %.2e g2 %.2e gflow %.2e" % (max_rel(out.detach().cpu().numpy(), ref),
              max_rel(a.grad.cpu().numpy(), g1), max_rel(b.grad.cpu().numpy(), g2), max_rel(f.grad.cpu().numpy(), gf)))
        if shape == (1, 4, 6, 7) and cfg == REF_CFG:
            print("g1 gpu\n", a.grad.cpu().numpy()[0, 0], "\nref\n", g1[0, 0])
            print("g2 gpu\n", b.grad.cpu().numpy()[0, 0], "\nref\n", g2[0, 0])
