#!/usr/bin/env python
"""Headline benchmark: 4DGS training throughput, 1M Gaussians, 1280x720 (BASELINE.json C3).

    python bench.py --gpus N --steps K --warmup W              # this repo (sm_100a kernels via the C ABI)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own code path

One step = one fine-stage training iteration over this rank's batch of views
(train_4DGS.py:172-297): per view HexPlane deformation -> exp/normalize/sigmoid -> rasterize ->
L1 -> backward through all of it; then (N > 1) ONE flat NCCL all-reduce of the gradients and one
fused Adam step over every parameter.  Views are sharded over ranks (weak scaling: the per-GPU
batch is fixed), `value` = view-iterations per second summed over all ranks, timed on the device
with CUDA events between barriers, max over ranks.  `e2e` repeats the measurement through the same
public API with the per-view ground-truth images living in pinned HOST memory (copied inside the
timed region, as train_4DGS.py:194 does) and the loss read back to the host every step.

The reference arm runs the reference's own CUDA rasterizer (oracle/_ref/libref_rast.so, built
unmodified from /root/reference by oracle/build_ref.sh), a plain-PyTorch port of its HexPlane /
deformation modules (oracle/field_torch.py, pinned bit-exact against the real modules) and
torch.optim.Adam, on the same GPU, same scene, same loop.  `cpu_baseline` is the CPU port
(oracle/raster_cpu.c + the torch field on CPU tensors) timed on the host cores for ONE view.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in ("iclr2025_3d-mom_b200", os.path.join("iclr2025_3d-mom_b200", "dropin"), "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cpu"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--views-per-gpu", type=int, default=8)
    ap.add_argument("--scale-mu", type=float, default=0.010, help="S-coarse 0.010 / S-fine 0.004 (SURVEY.md 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--render-frames", type=int, default=30, help="frames per GPU for the render-FPS side measurement (0 = skip)")
    ap.add_argument("--no-shared-spatial", action="store_true",
                    help="render every frame with the full six-plane HexPlane pass instead of sharing the spatial product over the sequence")
    ap.add_argument("--cpu-points", type=int, default=0, help="override the cpu_baseline sample size")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "samples": len(self.rows), "reasons": sorted(reasons)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
NCU_TRAFFIC = {"hexplane_bwd_kernel": 348.4e6, "deform_mlp_bwd_tc5_kernel": 1561.1e6, "deform_mlp_fwd_tc5v2_kernel": 1317.6e6,
               "hexplane_time_fwd_kernel": 485.5e6, "hexplane_time_bwd_kernel": 1037.8e6}
ROOFLINE_NOTES = {
    "deform_mlp_bwd_tc5_kernel": "HBM is the binding roofline (1.58 KB/point: 1 KB activation stash + features + d_features; the tensor pipe is 15 % "
                                 "busy), but the kernel is SIMT / shared-memory-port bound today: per phase the gradient math runs at 0.4 IPC per scheduler (two warps each) "
                                 "and 164 KB of operand stores go through the 128 B/clk shared-memory port (DESIGN.md section 7); traffic = ncu dram bytes of the first-generation kernel",
    "deform_mlp_fwd_tc5v2_kernel": "HBM is the binding roofline (1.37 KB/point, 1 KB of it the activation stash); tensor pipe 26 % busy",
    "hexplane_bwd_kernel": "average over the step's V time-plane passes and its one spatial pass; algorithmic HBM bytes only (xyz, order, "
                           "d_feature, shared spatial product in; d_xyz, its gradient accumulator, plane gradients out); the 3 KB/point/pass of "
                           "plane texel gathers + vector REDs are served by L1/L2 (planes are 11.6 MB), which is what bounds this kernel; "
                           "traffic = ncu dram bytes of the full six-plane launch",
    "hexplane_fwd_kernel": "plane texel gathers (6.1 KB/point) are L1/L2 traffic, not HBM",
}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


# ---------------------------------------------------------------------------------------------
def build_scene(args, device, world, rank, impl):
    from b200gs import synthetic as syn
    raw = syn.make_gaussians(args.points, scale_mu=args.scale_mu, device="cpu")
    n_global = args.views_per_gpu * world
    cams_all = syn.orbit_cameras(n_global, args.width, args.height, device=device)
    mine = list(range(rank, n_global, world))
    g = torch.Generator().manual_seed(1234)
    gts_host = []
    for b in range(n_global):
        img = torch.rand(3, args.height, args.width, generator=g)
        if b in mine:
            gts_host.append(img.pin_memory())
    cams = [cams_all[b] for b in mine]
    return raw, cams, gts_host, n_global


def make_b200_trainer(args, raw, device, world, rank):
    from b200gs import engine
    torch.manual_seed(6666)
    model = engine.GaussianState({k: v.to(device) for k, v in raw.items()}, hyper=engine.default_hyper()).to(device)
    with torch.no_grad():
        xyz = model._xyz
        model._deformation.deformation_net.set_aabb(xyz.max(0).values.tolist(), xyz.min(0).values.tolist())   # scene/__init__.py:72-78
        for p in model._deformation.deformation_net.grid.grids.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    model.training_setup()
    pg = None
    bg = torch.zeros(3, device=device)
    h = engine.default_hyper()
    return model, engine.ViewParallelTrainer(model, bg, stage="fine", process_group=pg, world_size=world, rank=rank,
                                             regulation=(h.time_smoothness_weight, h.l1_time_planes, h.plane_tv_weight))


# ---- reference arm -------------------------------------------------------------------------------
def make_reference_trainer(args, raw, device, world, rank):
    """Reference code path on the GPU: unmodified reference CUDA rasterizer (oracle/_ref) behind a
    torch.autograd.Function that does what RAST/diff_gaussian_rasterization/__init__.py and
    rasterize_points.cu do (torch.zeros gradient tensors included), the PyTorch port of the
    HexPlane field, torch.optim.Adam, and the same loop as ViewParallelTrainer."""
    import ref_harness as rh
    from b200gs import engine
    from oracle import field_torch
    if not rh.have_ref():
        return None, None
    L = rh.rast()
    p = rh._p

    class RefRaster(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, opac, scales, rots, cam, bg):
            P = means3D.shape[0]
            H, W = cam.image_height, cam.image_width
            color = torch.zeros(3, H, W, device=device); depth = torch.zeros(1, H, W, device=device)
            radii = torch.zeros(P, dtype=torch.int32, device=device)
            means3D, sh, opac, scales, rots = (t.contiguous() for t in (means3D, sh, opac, scales, rots))
            R = L.ref_rast_forward(P, 3, sh.shape[1], p(bg), W, H, p(means3D), p(sh), None, p(opac), p(scales), 1.0, p(rots), None,
                                   p(cam.viewmatrix), p(cam.projmatrix), p(cam.campos), cam.tanfovx, cam.tanfovy, 0, p(color),
                                   p(depth), p(radii), 0)
            assert R >= 0, L.ref_last_error()
            ctx.save_for_backward(means3D, sh, scales, rots, radii)
            ctx.cam, ctx.bg, ctx.R = cam, bg, R
            ctx.mark_non_differentiable(radii)
            return color, radii, depth

        @staticmethod
        def backward(ctx, dcolor, _dr, ddepth):
            means3D, sh, scales, rots, radii = ctx.saved_tensors
            cam, bg = ctx.cam, ctx.bg
            P = means3D.shape[0]; H, W = cam.image_height, cam.image_width
            z = lambda *s: torch.zeros(*s, device=device)
            g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacity=z(P, 1), colors=z(P, 3), depths=z(P, 1), means3D=z(P, 3),
                     cov3D=z(P, 6), sh=z(P, sh.shape[1], 3), scales=z(P, 3), rotations=z(P, 4))
            ddepth = ddepth.contiguous() if ddepth is not None else z(1, H, W)
            rc = L.ref_rast_backward(P, 3, sh.shape[1], ctx.R, p(bg), W, H, p(means3D), p(sh), None, p(scales), 1.0, p(rots), None,
                                     p(cam.viewmatrix), p(cam.projmatrix), p(cam.campos), cam.tanfovx, cam.tanfovy, p(radii),
                                     p(dcolor.contiguous()), p(ddepth), p(g["means2D"]), p(g["conic"]), p(g["opacity"]),
                                     p(g["colors"]), p(g["depths"]), p(g["means3D"]), p(g["cov3D"]), p(g["sh"]), p(g["scales"]),
                                     p(g["rotations"]), 0)
            assert rc == 0
            return g["means3D"], g["means2D"], g["sh"], g["opacity"], g["scales"], g["rotations"], None, None

    class RefModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(6666)
            ours = engine.GaussianState({k: v for k, v in raw.items()}, hyper=engine.default_hyper())
            self._xyz, self._features_dc, self._features_rest = ours._xyz, ours._features_dc, ours._features_rest
            self._scaling, self._rotation, self._opacity = ours._scaling, ours._rotation, ours._opacity
            self.register_buffer("_scene_flow", ours._scene_flow)
            # same initial field parameters, held as plain contiguous tensors keyed like the state_dict
            self.field = torch.nn.ParameterDict()
            self.keys = {}
            for k, v in ours._deformation.state_dict().items():
                if v.dtype.is_floating_point and "poc" not in k:
                    name = k.replace(".", "__")
                    self.field[name] = torch.nn.Parameter(v.detach().clone().contiguous(), requires_grad=not k.endswith("aabb"))
                    self.keys[k] = name
            self.levels = len(ours._deformation.deformation_net.grid.grids)
            self.optimizer = None

        @property
        def get_xyz(self):
            return self._xyz

        def named_parameters(self, *a, **k):        # names as in the product model so the arena filter applies
            for n, q in super().named_parameters(*a, **k):
                yield ("_deformation." + n[len("field."):].replace("__", ".") if n.startswith("field.") else n), q

    model = RefModel().to(device)
    with torch.no_grad():
        xyz = model._xyz
        aabb = model.field[model.keys["deformation_net.grid.aabb"]]
        aabb.copy_(torch.stack([xyz.max(0).values, xyz.min(0).values]))
        torch.manual_seed(6666)
        for k, n in model.keys.items():
            if ".grids." in k:
                model.field[n].add_(torch.randn_like(model.field[n]) * 0.01)
    o = engine.default_opt()
    mlp = [q for k, n in model.keys.items() if "grid" not in k for q in [model.field[n]]]
    grid = [q for k, n in model.keys.items() if "grid" in k for q in [model.field[n]]]
    model.optimizer = torch.optim.Adam([
        {'params': [model._xyz], 'lr': o.position_lr_init, "name": "xyz"}, {'params': mlp, 'lr': o.deformation_lr_init, "name": "deformation"},
        {'params': grid, 'lr': o.grid_lr_init, "name": "grid"}, {'params': [model._features_dc], 'lr': o.feature_lr, "name": "f_dc"},
        {'params': [model._features_rest], 'lr': o.feature_lr / 20.0, "name": "f_rest"},
        {'params': [model._opacity], 'lr': o.opacity_lr, "name": "opacity"}, {'params': [model._scaling], 'lr': o.scaling_lr, "name": "scaling"},
        {'params': [model._rotation], 'lr': o.rotation_lr, "name": "rotation"}], lr=0.0, eps=1e-15)

    def ref_render(cam, m, bg, stage):
        P = m._xyz.shape[0]
        sp = torch.zeros_like(m._xyz, requires_grad=True)
        shs = torch.cat((m._features_dc, m._features_rest), dim=1)
        time_ = torch.full((P, 1), float(cam.time), device=device)
        sd = {k: m.field[n] for k, n in m.keys.items()}
        pts, sc, rt, op, sh = field_torch.deform_forward(sd, m.levels, m._xyz, m._scaling, m._rotation, m._opacity, shs, time_,
                                                         m._scene_flow, cam.frame_num, 1)
        color, radii, depth = RefRaster.apply(pts, sp, sh, torch.sigmoid(op), torch.exp(sc), torch.nn.functional.normalize(rt), cam, bg)
        return {"render": color, "viewspace_points": sp, "radii": radii, "depth": depth}

    def ref_regulation():
        """compute_regulation exactly as scene/gaussian_model.py:730-769 + scene/regulation.py:22-28 spell it (PyTorch ops)."""
        h = engine.default_hyper()
        grids = [[model.field[model.keys[f"deformation_net.grid.grids.{l}.{k}"]] for k in range(6)] for l in range(model.levels)]

        def smooth(t):
            hh = t.shape[2]
            first = t[..., 1:, :] - t[..., :hh - 1, :]
            second = first[..., 1:, :] - first[..., :hh - 2, :]
            return torch.square(second).mean()
        plane = sum(smooth(g[k]) for g in grids for k in (0, 1, 3))
        tsm = sum(smooth(g[k]) for g in grids for k in (2, 4, 5))
        l1 = sum(torch.abs(1 - g[k]).mean() for g in grids for k in (2, 4, 5))
        return h.plane_tv_weight * plane + h.time_smoothness_weight * tsm + h.l1_time_planes * l1

    bg = torch.zeros(3, device=device)
    return model, engine.ViewParallelTrainer(model, bg, stage="fine", world_size=world, rank=rank, render_fn=ref_render,
                                             regulation_fn=ref_regulation)


# ---- render FPS (BASELINE.json metric, second half; config C4 style) ---------------------------------
def render_fps(args, model, device, world, rank, impl, ref_render=None):
    """Video rendering (render_4DGS.py:41-76): frames of an orbit with advancing time, sharded round-robin over
    ranks, no collective. `fps` = frames rendered per second on the device (deformation + rasterizer forward),
    `e2e_fps` adds the output path per frame: to8b + D2H into pinned host memory (our arm: GPU quantise + async
    3 B/pixel copy through a pinned ring; reference arm: the blocking float copy + host clip/cast of render_4DGS.py:49)."""
    import numpy as np
    from b200gs import engine, synthetic as syn
    out = {}
    bg = torch.zeros(3, device=device)
    for tag, (W, H) in (("1280x720", (args.width, args.height)), ("1920x1080", (1920, 1080))):
        n = args.render_frames
        cams = syn.orbit_cameras(n * world, W, H, device=device)[rank::world]
        fn = (lambda c: engine.render(c, model, bg, stage="fine")) if impl == "b200" else (lambda c: ref_render(c, model, bg, "fine"))
        # our arm renders the sequence the way engine.render_frames does: the spatial half of the HexPlane field is evaluated
        # once for the whole trajectory (the Gaussians do not move between frames), every frame samples its time planes only
        # (its one-off cost is inside both timed regions: a fresh context is entered after the start event / clock)
        import contextlib

        def shared():
            if impl == "b200" and not args.no_shared_spatial:
                from b200gs import field as _field
                return _field.shared_spatial_product(model._deformation, model._xyz)
            return contextlib.nullcontext()
        with torch.no_grad():
            with shared():
                for c in cams[:3]:
                    fn(c)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with shared():
                for c in cams:
                    fn(c)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            ms_plain = None
            if impl == "b200" and not args.no_shared_spatial:          # the same frames with the full six-plane pass per frame, for reference
                e0.record()
                for c in cams:
                    fn(c)
                e1.record()
                torch.cuda.synchronize()
                ms_plain = e0.elapsed_time(e1)
            if impl == "b200":
                from b200gs import output
                ring = output.FrameRing(H, W, depth=4, device=device)      # pinned buffers are allocated once, outside the loop
                ring.push(fn(cams[0])["render"]); ring.pop()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if impl == "b200":
                with shared():
                    for c in cams:
                        if ring.count == 4:
                            ring.pop()
                        ring.push(fn(c)["render"])
                    while ring.count:
                        ring.pop()
            else:
                for c in cams:
                    (255 * np.clip(fn(c)["render"].cpu().numpy(), 0, 1)).astype(np.uint8)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms, wall * 1e3, ms_plain if ms_plain is not None else 0.0], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
            if ms_plain is not None:
                ms_plain = float(t[2])
        out[tag] = {"fps": n * world / (ms / 1e3), "e2e_fps": n * world / wall, "frames": n * world,
                    "d2h_bytes_per_frame": 3 * W * H if impl == "b200" else 12 * W * H}
        if ms_plain is not None:
            out[tag]["fps_full_field_per_frame"] = n * world / (ms_plain / 1e3)
    return out


# ---- cpu baseline ---------------------------------------------------------------------------------
def cpu_baseline(args, raw, cam):
    """CPU port of ONE view-iteration (field fwd/bwd in torch on CPU tensors, rasterizer fwd/bwd in
    oracle/raster_cpu.c with OpenMP, Adam in C) on the host cores; sample bounded by --cpu-points."""
    import numpy as np
    from b200gs import engine
    from oracle import field_torch, raster_cpu as rc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = args.cpu_points or min(args.points, 200_000)
    frac = P / args.points
    W = max(16, int(round(args.width * frac ** 0.5 / 16)) * 16); H = max(16, int(round(args.height * frac ** 0.5 / 16)) * 16)
    from b200gs import synthetic as syn
    camc = syn.make_camera(W, H)
    sub = {k: v[:P].clone() for k, v in raw.items()}
    torch.manual_seed(6666)
    hyper = engine.default_hyper()
    # reference-initialised field parameters without touching CUDA: build the product module on CPU (parameters only)
    from b200gs.field import deform_network
    net = deform_network(hyper)
    sd = {k: v.detach().clone().contiguous().requires_grad_(v.dtype.is_floating_point and "poc" not in k and not k.endswith("aabb"))
          for k, v in net.state_dict().items()}
    xyz = sub["xyz"].clone().requires_grad_(True); scl = sub["log_scale"].clone().requires_grad_(True)
    rot = sub["rot"].clone().requires_grad_(True); opa = sub["opacity_logit"].clone().requires_grad_(True)
    shs = sub["shs"].clone().requires_grad_(True)
    t0 = time.perf_counter()
    tt = torch.full((P, 1), 0.5)
    pts, sc, rt, op, sh = field_torch.deform_forward(sd, 2, xyz, scl, rot, opa, shs, tt, sub["scene_flow"], 3, 1)
    a_sc, a_rt, a_op = torch.exp(sc), torch.nn.functional.normalize(rt), torch.sigmoid(op)
    s = rc.forward(pts.detach().numpy(), a_op.detach().numpy(), camc.viewmatrix.numpy(), camc.projmatrix.numpy(), camc.campos.numpy(),
                   W, H, camc.tanfovx, camc.tanfovy, np.zeros(3, np.float32), shs=sh.detach().numpy(), scales=a_sc.detach().numpy(),
                   rots=a_rt.detach().numpy())
    gt = np.random.default_rng(0).random((3, H, W), dtype=np.float32)
    dL = np.sign(s["color"] - gt).astype(np.float32) / (3 * H * W)
    g = rc.backward(s, dL, np.zeros((1, H, W), np.float32))
    torch.autograd.backward([pts, a_sc, a_rt, a_op, sh],
                            [torch.from_numpy(g["means3D"]), torch.from_numpy(g["scales"]), torch.from_numpy(g["rotations"]),
                             torch.from_numpy(g["opacity"]).reshape(P, 1), torch.from_numpy(g["sh"])])
    for q in [xyz, scl, rot, opa, shs] + [v for v in sd.values() if v.requires_grad and v.grad is not None]:
        pn = q.detach().numpy(); m = np.zeros_like(pn); v = np.zeros_like(pn)
        rc.adam_step(pn, q.grad.numpy().copy(), m, v, 1e-3, 1)
    dt = time.perf_counter() - t0
    # one view-iteration on the sample; the per-view cost scales ~linearly with Gaussians and pixels
    return {"value": (1.0 / dt) * frac, "unit": "view-iters/s", "cores": cores, "kind": "port",
            "sample": f"1 view-iteration (field fwd/bwd + raster fwd/bwd + Adam) at {P} Gaussians, {W}x{H}, measured {dt:.2f} s, "
                      f"scaled by {frac:.3f} to the {args.points}-Gaussian {args.width}x{args.height} workload",
            "measured_seconds": dt, "instances": int(s["R"])}


# ---------------------------------------------------------------------------------------------
def main():
    import contextlib
    # library chatter (Python prints AND C-level writes such as NCCL's version banner) goes to stderr;
    # the original stdout carries ONE JSON line
    sys.stdout.flush()
    real_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        with contextlib.redirect_stdout(sys.stderr):
            line = _main()
    finally:
        sys.stdout.flush()
        os.dup2(real_fd, 1)
    if line is not None:
        os.write(real_fd, (line + "\n").encode())
    os.close(real_fd)


def _main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    impl = args.impl
    if impl == "reference-cpu":
        if rank != 0:
            return None
        from b200gs import synthetic as syn
        raw = syn.make_gaussians(args.points, scale_mu=args.scale_mu, device="cpu")
        cb = cpu_baseline(args, raw, None)
        return json.dumps({"impl": "reference", "metric": "train_iters_per_s", "value": cb["value"], "unit": "view-iters/s",
                           "n_gpus": 0, "steps": 1, "warmup": 0, "higher_is_better": True, "cpu_baseline": cb,
                           "e2e": {"value": cb["value"], "unit": "view-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    launched_world = world
    if impl == "reference" and world > 1:
        # the reference has no multi-GPU path: under torchrun rank 0 alone runs it, on one GPU,
        # and prints the line; the other ranks leave without work
        if rank != 0:
            return None
        world = 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    raw, cams, gts_host, n_global = build_scene(args, device, world, rank, impl)
    if impl == "b200":
        model, trainer = make_b200_trainer(args, raw, device, world, rank)
    else:
        model, trainer = make_reference_trainer(args, raw, device, world, rank)
        if trainer is None:
            if rank == 0:
                return json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_rast.so missing (run oracle/build_ref.sh where /root/reference exists)"})
            return None
    gts_dev = [g.to(device, non_blocking=True) for g in gts_host]
    V = len(cams)
    pairs_per_view = None
    if impl == "b200":
        # evaluated (pixel, Gaussian) pairs = sum of n_contrib, counted once, outside the timed region
        from b200gs import engine as _eng
        from b200gs.rasterizer import _C as _rc
        with torch.no_grad():
            pk0 = _eng.render(cams[0], model, torch.zeros(3, device=device), stage="fine")
        try:
            sv = _eng.LAST_RASTER_STATE
            nc = _rc.export_state("n_contrib", args.points, sv[0], args.width, args.height, sv[1], sv[2], sv[3])
            pairs_per_view = int(nc.view(torch.int32).to(torch.int64).sum())
        except Exception:
            pairs_per_view = None

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return ms

    # Per-entry-point device timing (CUDA events on the launching stream) is taken INSIDE the timed steps:
    # our arm through b200gs._lib.CallTimer, the reference arm for its optimiser step only.
    adam_ms = []
    opt_step = model.optimizer.step
    timer = None
    if impl == "b200":
        from b200gs import _lib as _b200lib
        timer = _b200lib.CallTimer()
        timer.__enter__()
    else:
        def timed_opt_step(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = opt_step(*a, **k); e1.record()
            adam_ms.append((e0, e1))
            return r
        model.optimizer.step = timed_opt_step

    for _ in range(max(args.warmup, 3)):
        trainer.step(cams, gts_dev, global_batch=n_global)
    adam_ms.clear()
    if timer:
        timer.reset()
    barrier()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            trainer.step(cams, gts_dev, global_batch=n_global)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        calls = timer.summary() if timer else {}
        if timer:
            timer.__exit__()
            adam_t = calls.get("b200gs_adam_multi", {}).get("ms_avg", 0.0)
        else:
            adam_t = sum(a.elapsed_time(b) for a, b in adam_ms) / max(len(adam_ms), 1)
        # end to end: ground truth in pinned host memory, copied per view inside the timed region; loss read back
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.perf_counter()
        f0.record()
        last = 0.0
        for _ in range(args.steps):
            g_step = [g.to(device, non_blocking=True) for g in gts_host]
            last = float(trainer.step(cams, g_step, global_batch=n_global))
        f1.record()
        barrier()
        ms_e2e = max_over_ranks(f0.elapsed_time(f1))
        wall_e2e = time.perf_counter() - t_wall
    if not timer:
        model.optimizer.step = opt_step
    render = None
    render_note = None
    if args.render_frames > 0:
        try:
            render = render_fps(args, model, device, world, rank, impl, ref_render=None if impl == "b200" else trainer.render_fn)
        except Exception as ex:                      # the side measurement must not take the headline down: fall back to per-frame passes
            if impl != "b200" or args.no_shared_spatial:
                raise
            render_note = f"shared spatial product failed ({type(ex).__name__}: {ex}); per-frame six-plane passes measured instead"
            args.no_shared_spatial = True
            from b200gs import field as _field
            _field._SHARED = None
            render = render_fps(args, model, device, world, rank, impl, ref_render=None)
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return None

    views_total = V * world * args.steps
    value = views_total / (ms / 1e3)
    e2e_value = views_total / (ms_e2e / 1e3)
    pk, pk_kind = peaks()
    n_params = sum(p.numel() for p in trainer.trainable)
    adam_bytes = 28.0 * n_params                                   # p, m, v read+written, g read
    adam_gbs = adam_bytes / (adam_t * 1e-3) / 1e9 if adam_t > 0 else None
    hbm = pk.get("hbm_gbs")
    P_, F_ = args.points, 64
    plane_params = sum(p.numel() for n, p in model.named_parameters() if ".grids." in n)
    # algorithmic HBM bytes per launch: what each kernel must move when every
    # re-used operand (planes, weights) stays on chip
    model_bytes = {
        "b200gs_adam_multi": ("adam_multi_kernel", adam_bytes),
        "b200gs_hexplane_forward": ("hexplane_fwd_kernel", P_ * (12 + 4 + 4 * F_)),
        "b200gs_hexplane_backward": ("hexplane_bwd_kernel", P_ * (12 + 4 + 4 * F_ + 12) + 4 * plane_params),
        # shared-spatial step: V time-plane passes (factor S in, d(S) accumulated) + one spatial pass per optimiser step
        "b200gs_hexplane_forward_masked": ("hexplane_fwd_kernel", (V * P_ * (12 + 4 + 4 * F_ + 4 * F_) + P_ * (12 + 4 + 4 * F_)) / (V + 1)),
        "b200gs_hexplane_backward_masked": ("hexplane_bwd_kernel", (V * P_ * (12 + 4 + 4 * F_ + 4 * F_ + 8 * F_ + 12)
                                                                     + P_ * (12 + 4 + 4 * F_ + 12) + 4 * plane_params) / (V + 1)),
        # one-timestamp views: time planes from shared memory; streams xyz, S and the feature rows (+ the d(S) accumulator)
        "b200gs_hexplane_time_forward": ("hexplane_time_fwd_kernel", P_ * (12 + 4 * F_ + 4 * F_)),
        "b200gs_hexplane_time_backward": ("hexplane_time_bwd_kernel", P_ * (12 + 4 * F_ + 4 * F_ + 8 * F_ + 12)),
        "b200gs_deform_mlp_forward": ("deform_mlp_fwd_tc5v2_kernel", P_ * (4 * F_ + 52 + 4 * 4 * 64 + 40)),
        "b200gs_deform_mlp_backward": ("deform_mlp_bwd_tc5_kernel", P_ * (4 * 4 * 64 + 4 * F_ + 40 + 4 * F_)),
        "b200gs_activations_forward": ("activations_fwd_kernel", P_ * 64),
        "b200gs_activations_backward": ("activations_bwd_kernel", P_ * 96),
        "b200gs_l1_loss_fwd_bwd": ("l1_fwd_bwd_kernel", 12 * 3 * args.width * args.height),
    }
    kernels = []
    for name, c in sorted(calls.items(), key=lambda kv: -kv[1]["ms_total"]):
        row = {"entry": name, "calls": c["calls"], "ms_avg": round(c["ms_avg"], 4), "share_of_step": round(c["ms_total"] / ms, 4)}
        if name in model_bytes and c["ms_avg"] > 0:
            kname, nbytes = model_bytes[name]
            gbs = nbytes / (c["ms_avg"] * 1e-3) / 1e9
            row.update({"kernel": kname, "bound": "hbm", "algorithmic_bytes_per_launch": nbytes, "achieved_gbs": round(gbs, 1),
                        "frac_of_measured_hbm_peak": round(gbs / hbm, 4)})
        kernels.append(row)
    if pairs_per_view:
        # compositing is FP32-issue bound: SURVEY.md 8d cost model (reference SASS) = 85 FP32 lane-instructions per evaluated
        # pair in the backward; peak = 148 SMs x 128 lanes x 1.965 GHz. The entry also contains preprocess_bwd (HBM-bound, ~15 %).
        for row in kernels:
            if row["entry"] in ("b200gs_rast_backward", "b200gs_rast_backward_accumulate_sh"):
                rate = pairs_per_view * 85 / (row["ms_avg"] * 1e-3)
                row.update({"kernel": "composite_bwd_kernel (+ preprocess_bwd_kernel)", "bound": "fp32-issue", "pairs_per_launch": pairs_per_view,
                            "achieved_lane_instr_per_s": rate, "frac_of_fp32_peak_on_reference_cost_model": round(rate / (148 * 128 * 1.965e9), 4)})
    single = [k for k in kernels if "frac_of_measured_hbm_peak" in k]
    dom = single[0] if single else None
    launches_per_step = None
    if impl == "b200":
        # kernels of libb200gs launched in the timed region, counted per entry point by CallTimer (KERNELS table)
        launches_per_step = sum(c["kernels"] for c in calls.values()) / args.steps
    res = {
        "metric": "train_iters_per_s", "value": value, "unit": "view-iters/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C3 train_4DGS fine-stage iteration: HexPlane deform + raster fwd/bwd + plane regulariser + Adam, {args.points} Gaussians, "
                               f"{args.width}x{args.height}, {args.views_per_gpu} views per GPU per optimizer step (global batch {n_global})",
                   "scene": f"seeded synthetic, scale_mu={args.scale_mu}", "views_per_gpu": args.views_per_gpu,
                   "iters_per_s_at_batch": value / n_global,
                   "l2": "per-step working set (>= 236 MB of parameters + Adam state + 1M-splat records) exceeds the 126 MB L2",
                   "parallelism": f"view-parallel dp{world}"},
        "e2e": {"value": e2e_value, "unit": "view-iters/s", "h2d_bytes_per_step": V * 3 * args.height * args.width * 4,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall_e2e * 1e3 / args.steps,
                "last_loss": last},
        "clocks": clk.summary(),
        "gpu_launches": int(round(launches_per_step * args.steps)) if launches_per_step else 0,
    }
    peak_src = pk_kind + " (MEASURED_PEAKS.json hbm_gbs)" if pk_kind == "measured" else "fallback 6650"
    if impl == "b200" and dom is not None:
        # the dominant kernel of the step by device time (largest share among the single-kernel entry points)
        res["roofline"] = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": hbm, "unit": "GB/s",
                           "frac": dom["frac_of_measured_hbm_peak"], "traffic": NCU_TRAFFIC.get(dom["kernel"]), "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"], "ms_per_launch": dom["ms_avg"],
                           "share_of_step": dom["share_of_step"], "note": ROOFLINE_NOTES.get(dom["kernel"], "")}
        res["kernels"] = kernels
    else:
        res["roofline"] = {"kernel": "torch foreach Adam", "bound": "hbm", "achieved": adam_gbs, "peak": hbm, "unit": "GB/s",
                           "frac": (adam_gbs / hbm) if adam_gbs else None, "traffic": None, "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": adam_bytes, "ms_per_launch": adam_t, "params": n_params}
    if impl == "b200":
        # which kernel generations ran (include/b200gs.h: b200gs_set_option; environment B200GS_* overrides)
        try:
            from b200gs import _lib as _l
            res["config"]["kernel_options"] = {n: int(_l.lib().b200gs_get_option(n.encode()))
                                               for n in ("mlp_fwd_elect", "mlp_bwd_v2", "hexplane_time_fwd", "hexplane_time_bwd", "lookback_parallel")}
        except Exception as ex:              # informational only
            res["config"]["kernel_options"] = f"unavailable: {ex}"
    if render is not None:
        res["render"] = render
        res["config"]["render"] = (f"video rendering, {args.render_frames} frames per GPU of an orbit with advancing time, frames sharded "
                                   "round-robin over ranks; fps = device time, e2e_fps = wall clock incl. to8b + D2H per frame"
                                   + ("" if impl != "b200" or args.no_shared_spatial else
                                      "; spatial HexPlane product evaluated once per sequence (engine.render_frames), time planes per frame"))
        if render_note:
            res["config"]["render_note"] = render_note
    if impl != "b200":
        res["impl"] = "reference"
        res["reference_stack"] = "reference CUDA rasterizer (oracle/_ref, unmodified) + PyTorch port of HexPlane/deformation + torch.optim.Adam, on GPU"
        if launched_world > 1:
            res["n_gpus"] = launched_world
            res["ranks_used"] = 1
            res["config"]["parallelism"] = f"single GPU (the reference has no multi-GPU path; launched with {launched_world} ranks, rank 0 ran)"
        # schema completeness: this arm is the reference's own implementation of the path, which is CUDA (it has no CPU
        # implementation); one host thread drives it. The CPU port is what `cpu_baseline` on the product line times.
        res["cpu_baseline"] = {"value": value, "unit": "view-iters/s", "cores": 1, "kind": "reference",
                               "sample": "the reference's own code path for this workload (its CUDA rasterizer from oracle/_ref + its PyTorch "
                                         "field + torch Adam) on the same GPU, full workload; host side single-threaded"}
    if not args.no_cpu_baseline and world == 1 and impl == "b200":
        try:
            res["cpu_baseline"] = cpu_baseline(args, raw, cams[0])
        except Exception as ex:          # the checker must never take the product line down
            res["cpu_baseline"] = {"value": None, "unit": "view-iters/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return json.dumps(res)


if __name__ == "__main__":
    main()
