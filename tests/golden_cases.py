"""Seeded inputs of the committed golden fixtures (mirrors oracle/gen_golden_raster.py)."""
import numpy as np
import torch

from b200gs import synthetic as syn

CASES = {"small": dict(P=3000, W=96, H=64, mu=0.02, seed=11, depth_grad=True),
         "medium": dict(P=20000, W=200, H=120, mu=0.01, seed=12, depth_grad=False)}
BG = [0.1, 0.2, 0.3]


def scene(name, device="cpu"):
    c = CASES[name]
    raw = syn.make_gaussians(c["P"], scale_mu=c["mu"], seed=c["seed"], device="cpu")
    act = {k: v.to(device) for k, v in syn.activated(raw).items()}     # activations always on the CPU (as the fixture)
    return c, act, syn.make_camera(c["W"], c["H"], device=device)


def upstream(c):
    g = torch.Generator().manual_seed(c["seed"] + 100)
    W, H = c["W"], c["H"]
    dLc = (torch.rand(3, H, W, generator=g) - 0.5) / (3 * H * W)
    dLd = torch.randn(1, H, W, generator=g) / (H * W) if c["depth_grad"] else torch.zeros(1, H, W)
    return dLc, dLd


def knn_points():
    g = torch.Generator().manual_seed(21)
    return torch.cat([torch.rand(4000, 3, generator=g) * 3 - 1.5, torch.randn(1000, 3, generator=g) * 0.01 + 0.5,
                      torch.zeros(8, 3)])


def load(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
