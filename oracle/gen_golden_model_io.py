"""TEST INFRASTRUCTURE. Generates tests/golden/model_io.ply and tests/golden/model_io.pt by running the REAL reference
`GaussianModel.save_ply` / `load_ply` (scene/gaussian_model.py:342-360, 367-407), imported from /root/reference in this container,
on CPU tensors:   python oracle/gen_golden_model_io.py
`plyfile` (a third-party dependency of the reference, not installed here and not vendored under /root/reference) is stood in for
by iclr2025_3d-mom_b200/compat/plyfile.py, which writes the published binary little-endian PLY layout (header lines `property
float <name>` in dtype order, then the packed records).  The reference hard-codes device="cuda" in load_ply's torch.tensor calls;
for the run here that keyword is dropped (values untouched)."""
import os
import sys

import torch

REF = os.environ.get("REF_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_field import stub_modules, hyper
    stub_modules()
    sys.path.insert(0, os.path.join(ROOT, "iclr2025_3d-mom_b200", "compat"))
    sys.modules.pop("plyfile", None)
    import plyfile                                                    # noqa: F401  (the compat module, see the header)
    sys.path.insert(0, REF)
    real_tensor = torch.tensor

    def tensor_anywhere(*a, **k):
        k.pop("device", None)
        return real_tensor(*a, **k)
    from scene.gaussian_model import GaussianModel
    torch.manual_seed(6666)
    gm = GaussianModel(3, hyper([1, 2], [6, 5, 7, 4]))
    N = 11
    g = torch.Generator().manual_seed(23)
    rnd = lambda *s: torch.randn(*s, generator=g)
    inputs = {"xyz": rnd(N, 3), "f_dc": rnd(N, 1, 3), "f_rest": rnd(N, 15, 3) * 0.1, "opacity": rnd(N, 1), "scaling": rnd(N, 3) - 4.0,
              "rotation": rnd(N, 4)}
    gm._xyz = torch.nn.Parameter(inputs["xyz"].clone())
    gm._features_dc = torch.nn.Parameter(inputs["f_dc"].clone())
    gm._features_rest = torch.nn.Parameter(inputs["f_rest"].clone())
    gm._opacity = torch.nn.Parameter(inputs["opacity"].clone())
    gm._scaling = torch.nn.Parameter(inputs["scaling"].clone())
    gm._rotation = torch.nn.Parameter(inputs["rotation"].clone())
    out = os.path.join(ROOT, "tests", "golden")
    ply = os.path.join(out, "model_io.ply")
    gm.save_ply(ply)
    gm2 = GaussianModel(3, hyper([1, 2], [6, 5, 7, 4]))
    torch.tensor = tensor_anywhere
    try:
        gm2.load_ply(ply)
    finally:
        torch.tensor = real_tensor
    loaded = {"xyz": gm2._xyz, "f_dc": gm2._features_dc, "f_rest": gm2._features_rest, "opacity": gm2._opacity,
              "scaling": gm2._scaling, "rotation": gm2._rotation}
    torch.save({"inputs": inputs, "loaded": {k: v.detach().clone() for k, v in loaded.items()},
                "active_sh_degree": gm2.active_sh_degree}, os.path.join(out, "model_io.pt"))
    print("wrote", ply, os.path.getsize(ply), "bytes")


if __name__ == "__main__":
    main()
