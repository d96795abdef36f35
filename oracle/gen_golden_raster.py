"""TEST INFRASTRUCTURE. Generates tests/golden/raster_*.npz and knn_*.npz by running the
REFERENCE's own CUDA code (oracle/_ref/*.so, built from /root/reference by oracle/build_ref.sh)
on a B200:   gpurun -- 'python oracle/gen_golden_raster.py gpurun_out/golden'
then copy gpurun_out/golden/*.npz into tests/golden/. Inputs are regenerated from seeds by the
tests (b200gs.synthetic), only the reference's outputs are stored."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("iclr2025_3d-mom_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import ref_harness as rh  # noqa: E402
from b200gs import synthetic as syn  # noqa: E402

CASES = {"small": dict(P=3000, W=96, H=64, mu=0.02, seed=11, depth_grad=True),
         "medium": dict(P=20000, W=200, H=120, mu=0.01, seed=12, depth_grad=False)}


def upstream(case):
    g = torch.Generator().manual_seed(case["seed"] + 100)
    W, H = case["W"], case["H"]
    dLc = (torch.rand(3, H, W, generator=g) - 0.5) / (3 * H * W)
    dLd = torch.randn(1, H, W, generator=g) / (H * W) if case["depth_grad"] else torch.zeros(1, H, W)
    return dLc, dLd


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        # activations are evaluated on the CPU so the CPU tests see bit-identical inputs
        raw = syn.make_gaussians(c["P"], scale_mu=c["mu"], seed=c["seed"], device="cpu")
        act = {k: v.cuda() for k, v in syn.activated(raw).items()}
        cam = syn.make_camera(c["W"], c["H"], device="cuda")
        bg = torch.tensor([0.1, 0.2, 0.3], device="cuda")
        R, color, depth, radii = rh.ref_forward(cam, bg, act["means3D"], act["opacities"], shs=act["shs"],
                                                scales=act["scales"], rotations=act["rotations"])
        out = dict(R=np.int64(R), color=color.cpu().numpy(), depth=depth.cpu().numpy(), radii=radii.cpu().numpy())
        tiles = ((c["W"] + 15) // 16) * ((c["H"] + 15) // 16)
        for f in ("tiles_touched", "keys", "point_list", "n_contrib", "accum_alpha", "depths", "means2D", "conic_opacity",
                  "rgb", "cov3D", "clamped"):
            out[f] = rh.ref_get(f).copy()
        out["ranges"] = rh.ref_get("ranges")[: 2 * tiles].copy()
        dLc, dLd = upstream(c)
        g = rh.ref_backward(cam, bg, R, radii, dLc.cuda(), dLd.cuda(), act["means3D"], shs=act["shs"],
                            scales=act["scales"], rotations=act["rotations"])
        for k, v in g.items():
            out["grad_" + k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, f"raster_{name}.npz"), **out)
        print(name, "R =", R, "visible =", int((radii > 0).sum()))
    g = torch.Generator().manual_seed(21)
    pts = torch.cat([torch.rand(4000, 3, generator=g) * 3 - 1.5, torch.randn(1000, 3, generator=g) * 0.01 + 0.5,
                     torch.zeros(8, 3)]).cuda()
    np.savez_compressed(os.path.join(out_dir, "knn_5008.npz"), dist2=rh.ref_dist2(pts).cpu().numpy())
    print("knn done")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
