"""TEST INFRASTRUCTURE. Generates tests/golden/cpu_path.pt from the REAL reference functions (utils/sh_utils.py::eval_sh,
utils/general_utils.py::{build_rotation, build_scaling_rotation, strip_symmetric}, utils/graphics_utils.py::{geom_transform_points,
getProjectionMatrix}) imported from /root/reference in this container; run here (no GPU):  python oracle/gen_golden_cpu_path.py
The reference hard-codes device="cuda" inside three of those helpers (general_utils.py:71,89,108); for the run here
torch.zeros is wrapped to drop that keyword -- the arithmetic is untouched."""
import os
import sys

import torch

REF = os.environ.get("REF_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_field import stub_modules
    stub_modules()
    sys.path.insert(0, REF)
    from utils import general_utils, graphics_utils, sh_utils
    real_zeros = torch.zeros

    def zeros_anywhere(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)
    g = torch.Generator().manual_seed(6666)
    P = 301
    sh = torch.randn(P, 3, 16, generator=g)
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, generator=g))
    scaling = torch.exp(torch.randn(P, 3, generator=g) * 0.6 - 4.0)
    rot = torch.randn(P, 4, generator=g)
    pts = torch.rand(P, 3, generator=g) * 3 - 1.5
    fovx, fovy = 1.1, 0.7
    proj = graphics_utils.getProjectionMatrix(0.01, 100.0, fovx, fovy)
    R = torch.linalg.qr(torch.randn(3, 3, generator=g))[0].numpy()
    t = torch.tensor([0.1, -0.2, 4.5]).numpy()
    wv = torch.tensor(graphics_utils.getWorld2View2(R, t)).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(proj.transpose(0, 1).unsqueeze(0))).squeeze(0)
    out = {"sh": sh, "dirs": dirs, "scaling": scaling, "rot": rot, "pts": pts, "fovx": fovx, "fovy": fovy, "proj": proj,
           "world_view": wv, "full_proj": full}
    for deg in range(4):
        out[f"eval_sh_{deg}"] = sh_utils.eval_sh(deg, sh, dirs)
    torch.zeros = zeros_anywhere
    try:
        L = general_utils.build_scaling_rotation(0.7 * scaling, rot)
        out["cov_0p7"] = general_utils.strip_symmetric(L @ L.transpose(1, 2))
        L = general_utils.build_scaling_rotation(1.0 * scaling, rot)
        out["cov_1"] = general_utils.strip_symmetric(L @ L.transpose(1, 2))
        out["rotation"] = general_utils.build_rotation(rot)
    finally:
        torch.zeros = real_zeros
    out["ndc"] = graphics_utils.geom_transform_points(pts, full)
    path = os.path.join(ROOT, "tests", "golden", "cpu_path.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
