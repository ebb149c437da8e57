"""Seeded synthetic inputs shared by the tests."""
import numpy as np

# (pad, kernel, max_disp, stride1, stride2)
REF_CFG = (9, 1, 9, 1, 2)      # the literal reference configuration (model.py:24, search_range=4)
CANON_CFG = (4, 1, 4, 1, 1)    # canonical PWC-Net cost volume ("md=4" in BASELINE.json)
GENERIC_CFGS = [
    (3, 3, 4, 1, 2),   # kernel_size 3
    (2, 1, 4, 2, 1),   # stride1 2, pad < md
    (5, 3, 3, 2, 1),   # pad > md, kernel 3, stride1 2
    (0, 1, 2, 1, 1),   # no padding
    (4, 1, 4, 1, 2),   # D = 5
    (20, 1, 20, 1, 2), # FlowNetC-style 21x21 volume
]


def smooth_flow(rng, B, H, W, sigma):
    """8x-downsampled noise, bilinearly upsampled: a 'typical' smooth flow field."""
    h, w = max(2, H // 8 + 1), max(2, W // 8 + 1)
    coarse = rng.standard_normal((B, 2, h, w)) * sigma
    ys = np.linspace(0, h - 1, H)
    xs = np.linspace(0, w - 1, W)
    y0 = np.floor(ys).astype(int).clip(0, h - 2)
    x0 = np.floor(xs).astype(int).clip(0, w - 2)
    ay = (ys - y0)[None, None, :, None]
    ax = (xs - x0)[None, None, None, :]
    c = coarse
    top = c[:, :, y0][:, :, :, x0] * (1 - ax) + c[:, :, y0][:, :, :, x0 + 1] * ax
    bot = c[:, :, y0 + 1][:, :, :, x0] * (1 - ax) + c[:, :, y0 + 1][:, :, :, x0 + 1] * ax
    return (top * (1 - ay) + bot * ay).astype(np.float32)


def make_inputs(B, C, H, W, seed=0, flow_sigma=2.0, flow_kind="iid"):
    rng = np.random.Generator(np.random.PCG64(seed))
    f1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    if flow_kind == "iid":
        flow = (flow_sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
    elif flow_kind == "smooth":
        flow = smooth_flow(rng, B, H, W, flow_sigma)
    elif flow_kind == "zero":
        flow = np.zeros((B, 2, H, W), np.float32)
    elif flow_kind == "integer":
        flow = np.round(flow_sigma * rng.standard_normal((B, 2, H, W))).astype(np.float32)
    else:
        raise ValueError(flow_kind)
    return f1, f2, flow, rng


def max_rel(a, b):
    """||a-b||_inf / ||b||_inf  (relative error per element is meaningless for a dot product that
    cancels to ~0)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0))
