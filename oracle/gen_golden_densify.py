"""TEST INFRASTRUCTURE. Generates tests/golden/densify.pt by running the REAL reference `GaussianModel` bookkeeping
(scene/gaussian_model.py: training_setup, add_densification_stats, densify -> densify_and_clone / densify_and_split /
densification_postfix / prune_points, prune, reset_opacity) with torch.optim.Adam, imported from /root/reference in this
container, on CPU tensors:   python oracle/gen_golden_densify.py
The reference hard-codes device="cuda" in its torch.zeros calls; for the run here torch.zeros drops that keyword (arithmetic
untouched).  torch.normal is wrapped only to RECORD the split offsets it draws, so that the GPU parity test can replay them
(the CPU and CUDA generators produce different streams)."""
import os
import sys
import types

import torch

REF = os.environ.get("REF_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def snapshot(gm):
    st = gm.optimizer.state
    out = {}
    for name, attr in (("xyz", "_xyz"), ("f_dc", "_features_dc"), ("f_rest", "_features_rest"), ("opacity", "_opacity"),
                       ("scaling", "_scaling"), ("rotation", "_rotation")):
        p = getattr(gm, attr)
        out[name] = p.detach().clone()
        out[name + ".exp_avg"] = st[p]["exp_avg"].clone()
        out[name + ".exp_avg_sq"] = st[p]["exp_avg_sq"].clone()
        out[name + ".step"] = float(st[p]["step"])
    for attr in ("xyz_gradient_accum", "denom", "max_radii2D", "_deformation_accum", "_deformation_table", "_scene_flow"):
        out[attr] = getattr(gm, attr).detach().clone()
    return out


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from gen_golden_field import stub_modules, hyper
    stub_modules()
    sys.path.insert(0, REF)
    real_zeros, real_normal = torch.zeros, torch.normal

    def zeros_anywhere(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)
    drawn = []

    def recording_normal(*a, **k):
        r = real_normal(*a, **k)
        drawn.append(r.clone())
        return r
    torch.zeros, torch.normal = zeros_anywhere, recording_normal
    try:
        from scene.gaussian_model import GaussianModel
        torch.manual_seed(6666)
        gm = GaussianModel(3, hyper([1, 2], [6, 5, 7, 4]))
        N = 400
        g = torch.Generator().manual_seed(11)
        rnd = lambda *s: torch.randn(*s, generator=g)
        gm._xyz = torch.nn.Parameter(rnd(N, 3))
        gm._features_dc = torch.nn.Parameter(rnd(N, 1, 3))
        gm._features_rest = torch.nn.Parameter(rnd(N, 15, 3) * 0.1)
        gm._opacity = torch.nn.Parameter(rnd(N, 1) * 2.5 - 1.0)
        gm._scaling = torch.nn.Parameter(rnd(N, 3) * 0.7 - 4.6)          # exp ~ 0.01: straddles percent_dense * extent = 0.01
        gm._rotation = torch.nn.Parameter(rnd(N, 4))
        gm._scene_flow = rnd(N, 3) * 1e-3
        gm._deformation_table = torch.rand(N, generator=g) > 0.2
        gm.max_radii2D = real_zeros(N)
        opt_args = types.SimpleNamespace(
            percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=20000,
            deformation_lr_init=0.00016, deformation_lr_final=0.0000016, deformation_lr_delay_mult=0.01, grid_lr_init=0.0016,
            grid_lr_final=0.000016, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)
        gm.spatial_lr_scale = 1.0
        gm.training_setup(opt_args)
        for _ in range(2):                         # two Adam steps so that the moments are non-trivial
            for p in (gm._xyz, gm._features_dc, gm._features_rest, gm._opacity, gm._scaling, gm._rotation):
                p.grad = rnd(*p.shape) * 1e-2
            gm.optimizer.step()
        out = {"extent": 1.0, "max_grad": 0.0002, "min_opacity": 0.005}
        # ---- add_densification_stats over three "views" ----
        stats_in = []
        for v in range(3):
            vg = rnd(N, 3) * 3e-4
            flt = torch.rand(N, generator=g) > 0.4
            stats_in.append((vg.clone(), flt.clone()))
            gm.add_densification_stats(vg, flt)
        out["stats_in"] = stats_in
        gm.max_radii2D = torch.rand(N, generator=g) * 30.0
        out["before_densify"] = snapshot(gm)
        gm.densify(out["max_grad"], out["min_opacity"], out["extent"], None, 5, 5, None, 1, "fine")
        out["normal_samples"] = [d.clone() for d in drawn]
        out["after_densify"] = snapshot(gm)
        # ---- prune with the screen-size / world-size criteria switched on ----
        gm.max_radii2D = torch.rand(gm._xyz.shape[0], generator=g) * 30.0
        gm.xyz_gradient_accum = rnd(gm._xyz.shape[0], 1).abs()
        gm.denom = torch.rand(gm._xyz.shape[0], 1, generator=g) * 3
        gm._deformation_accum = rnd(gm._xyz.shape[0], 3)
        out["before_prune"] = snapshot(gm)
        gm.prune(out["max_grad"], out["min_opacity"], out["extent"], 20)
        out["after_prune"] = snapshot(gm)
        gm.prune(out["max_grad"], 0.3, out["extent"], None)           # opacity criterion alone
        out["after_prune_opacity_only"] = snapshot(gm)
        gm.reset_opacity()
        out["after_reset_opacity"] = snapshot(gm)
    finally:
        torch.zeros, torch.normal = real_zeros, real_normal
    for k in ("before_densify", "after_densify", "after_prune", "after_reset_opacity"):
        print(k, out[k]["xyz"].shape[0], "points")
    print("split offsets drawn:", [tuple(d.shape) for d in out["normal_samples"]])
    path = os.path.join(ROOT, "tests", "golden", "densify.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
