"""Scratch: run ours fwd+bwd once or twice at a given size (for ncu captures)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("iclr2025_3d-mom_b200", "iclr2025_3d-mom_b200/dropin", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from b200gs import synthetic as syn
from b200gs.rasterizer import _C
P, W, H, mu = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
raw = syn.make_gaussians(P, scale_mu=mu, device="cuda"); act = syn.activated(raw)
cam = syn.make_camera(W, H, device="cuda")
bg = torch.zeros(3, device="cuda"); E = torch.Tensor([])
gt = torch.rand(3, H, W, device="cuda")
for _ in range(reps):
    R, color, depth, radii, geom, binb, img = _C.rasterize_gaussians(bg, act["means3D"], E, act["opacities"], act["scales"], act["rotations"], 1.0, E,
        cam.viewmatrix, cam.projmatrix, cam.tanfovx, cam.tanfovy, H, W, act["shs"], 3, cam.campos, False, False)
    dLc = torch.sign(color - gt) / (3 * H * W); dLd = torch.zeros(1, H, W, device="cuda")
    _C.rasterize_gaussians_backward(bg, act["means3D"], radii, E, act["scales"], act["rotations"], 1.0, E, cam.viewmatrix, cam.projmatrix,
        cam.tanfovx, cam.tanfovy, dLc, dLd, act["shs"], 3, cam.campos, geom, R, binb, img, False)
torch.cuda.synchronize()
print("done R=", R)
