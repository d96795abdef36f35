"""C5 stress (BASELINE.json configs[4]): 5M Gaussians at 3840x2160 — distCUDA2 initialisation on 5M points, then
training steps with a prune + clone event every few steps through the fused bookkeeping. Prints timings and checks
that everything stays finite. Usage: python tools/stress_c5.py [points] [width] [height] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("iclr2025_3d-mom_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from b200gs import engine, synthetic as syn, densify
from b200gs.knn import distCUDA2

P = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3840
H = int(sys.argv[3]) if len(sys.argv) > 3 else 2160
STEPS = int(sys.argv[4]) if len(sys.argv) > 4 else 12
dev = torch.device("cuda", 0)
ev = lambda: torch.cuda.Event(enable_timing=True)

raw = syn.make_gaussians(P, scale_mu=0.004, device="cpu")
xyz = raw["xyz"].to(dev)
a, b = ev(), ev(); a.record()
d2 = distCUDA2(xyz)
b.record(); torch.cuda.synchronize()
print(f"distCUDA2 on {P} points: {a.elapsed_time(b):.1f} ms; mean nn dist^2 {d2.mean().item():.3e}; finite {bool(torch.isfinite(d2).all())}")
# scales from the kNN distances, as create_from_pcd does (scene/gaussian_model.py:160-161)
raw["log_scale"] = torch.log(torch.sqrt(torch.clamp_min(d2, 1e-7)))[..., None].repeat(1, 3).cpu()
torch.manual_seed(6666)
model = engine.GaussianState({k: v.to(dev) for k, v in raw.items()}, hyper=engine.default_hyper()).to(dev)
with torch.no_grad():
    model._deformation.deformation_net.set_aabb(xyz.max(0).values.tolist(), xyz.min(0).values.tolist())
model.training_setup()
cams = syn.orbit_cameras(2, W, H, device=dev)
gts = [torch.rand(3, H, W, device=dev) for _ in cams]
h = engine.default_hyper()
def make_trainer():
    return engine.ViewParallelTrainer(model, torch.zeros(3, device=dev), stage="fine",
                                      regulation=(h.time_smoothness_weight, h.l1_time_planes, h.plane_tv_weight))
tr = make_trainer()
for it in range(STEPS):
    a, b = ev(), ev(); a.record()
    loss = tr.step(cams, gts)
    b.record(); torch.cuda.synchronize()
    n = model._xyz.shape[0]
    print(f"step {it:3d}: {a.elapsed_time(b):8.1f} ms for {len(cams)} views, {n} Gaussians, loss {float(loss):.5f}, "
          f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    assert torch.isfinite(loss).all() and torch.isfinite(model._xyz).all()
    if it % 4 == 3:
        # densify / prune event on the statistics of this step (gaussian_model.py:541-581, 681-692, restated minimally):
        # prune low-opacity Gaussians, clone the ones with the largest screen-space gradient
        a, b = ev(), ev(); a.record()
        with torch.no_grad():
            g = tr.viewspace_grad[:, :2].norm(dim=-1)
            keep = torch.sigmoid(model._opacity).squeeze(-1) > 0.005
            opt = model.optimizer
            g = g[keep]
            new = densify.prune_optimizer(opt, keep)
            model._scene_flow = model._scene_flow[keep]
            thr = torch.quantile(g[:: max(1, g.numel() // 1_000_000)], 0.98)
            sel = g >= thr
            ext = {k: v.detach()[sel] for k, v in new.items()}
            flow_ext = model._scene_flow[sel]
            new = densify.cat_tensors_to_optimizer(opt, ext)
            model._scene_flow = torch.cat((model._scene_flow, flow_ext), 0)
            model._xyz, model._features_dc, model._features_rest = new["xyz"], new["f_dc"], new["f_rest"]
            model._opacity, model._scaling, model._rotation = new["opacity"], new["scaling"], new["rotation"]
        tr = make_trainer()
        b.record(); torch.cuda.synchronize()
        print(f"   densify/prune event: {a.elapsed_time(b):.1f} ms, kept {int(keep.sum())}, cloned {int(sel.sum())} -> {model._xyz.shape[0]}")
print("stress ok")
