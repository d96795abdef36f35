"""Host-side hot loop: the model state, `render()` and the view-parallel training step.

This is the part of the reference's L2/L3 layers that the hot path needs, restated so that
bench.py, the tests and the multi-GPU trainer can drive the kernels without the reference
tree (which does not exist on the GPU box):
  * `GaussianState`  — the tensors and optimiser groups of scene/gaussian_model.py:55-70, :190-209
                       (same attribute names: _xyz, _features_dc, _features_rest, _scaling,
                       _rotation, _opacity, _scene_flow, _deformation; same 8 Adam groups / lrs)
  * `render()`       — gaussian_renderer/__init__.py:22-178 (settings, deformation in the fine
                       stage, exp / normalize / sigmoid activations, rasterizer call, result dict)
  * `ViewParallelTrainer` — train_4DGS.py:172-297 for a batch of views, with the batch sharded
                       over ranks: every rank renders its views, one flat NCCL all-reduce sums the
                       gradients (Gaussians | planes | MLP | screen-space xy), every rank takes
                       the identical fused Adam step.
With the real reference tree on sys.path, train_4DGS.py / render_4DGS.py run unchanged through
b200gs.launcher instead.
"""
import math
import os
import types

import torch
import torch.nn as nn

from . import fusedops
from . import rasterizer as _rast
from .adam import FusedAdam
from .field import deform_network
from .rasterizer import GaussianRasterizationSettings, GaussianRasterizer


def default_hyper(multires=(1, 2), time_res=50):
    """ModelHiddenParams as train_4DGS.py sees them with its default config
    (arguments/__init__.py:75-106 overridden by arguments/dnerf/hellwarrior.py + dnerf_default.py)."""
    return types.SimpleNamespace(
        net_width=64, timebase_pe=4, defor_depth=0, posebase_pe=10, scale_rotation_pe=2, opacity_pe=2,
        timenet_width=64, timenet_output=32, bounds=1.6, plane_tv_weight=0.0001, time_smoothness_weight=0.01,
        l1_time_planes=0.0001, grid_pe=0,
        kplanes_config={'grid_dimensions': 2, 'input_coordinate_dim': 4, 'output_coordinate_dim': 32,
                        'resolution': [64, 64, 64, time_res]},
        multires=list(multires), no_dx=False, no_grid=False, no_ds=False, no_dr=False, no_do=True, no_dshs=True,
        empty_voxel=False, static_mlp=False, apply_rotation=False)


def default_opt():
    """OptimizationParams (arguments/__init__.py:110-151 + dnerf_default.py)."""
    return types.SimpleNamespace(
        position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=20000,
        deformation_lr_init=0.00016, deformation_lr_final=0.0000016, deformation_lr_delay_mult=0.01,
        grid_lr_init=0.0016, grid_lr_final=0.000016, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005,
        rotation_lr=0.001, percent_dense=0.01)


def expon_lr(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1.0):
    """The reference's learning-rate schedule (utils/general_utils.py:35-68, get_expon_lr_func): log-linear interpolation from
    lr_init at step 0 to lr_final at max_steps, optionally eased in by a reverse-cosine delay. Python floats, like the
    reference (the optimiser reads `param_group['lr']` on the host)."""
    if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
        return 0.0
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * min(max(step / lr_delay_steps, 0.0), 1.0))
    else:
        delay_rate = 1.0
    t = min(max(step / max_steps, 0.0), 1.0)
    return delay_rate * math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


class GaussianState(nn.Module):
    def __init__(self, raw, sh_degree=3, hyper=None, spatial_lr_scale=1.0):
        """raw: dict from b200gs.synthetic.make_gaussians (xyz, log_scale, rot, opacity_logit, shs, scene_flow)."""
        super().__init__()
        self.max_sh_degree = sh_degree
        self.active_sh_degree = sh_degree
        self.spatial_lr_scale = spatial_lr_scale
        self._xyz = nn.Parameter(raw["xyz"].clone())
        self._features_dc = nn.Parameter(raw["shs"][:, :1, :].clone().contiguous())
        self._features_rest = nn.Parameter(raw["shs"][:, 1:, :].clone().contiguous())
        self._scaling = nn.Parameter(raw["log_scale"].clone())
        self._rotation = nn.Parameter(raw["rot"].clone())
        self._opacity = nn.Parameter(raw["opacity_logit"].clone())
        self.register_buffer("_scene_flow", raw["scene_flow"].clone())
        self._deformation = deform_network(hyper or default_hyper())
        self.optimizer = None

    # accessors named as in scene/gaussian_model.py:117-147
    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_flow(self):
        return self._scene_flow

    scaling_activation = staticmethod(torch.exp)
    opacity_activation = staticmethod(torch.sigmoid)
    rotation_activation = staticmethod(torch.nn.functional.normalize)

    def training_setup(self, opt=None):
        opt = opt or default_opt()
        s = self.spatial_lr_scale
        groups = [
            {'params': [self._xyz], 'lr': opt.position_lr_init * s, "name": "xyz"},
            {'params': list(self._deformation.get_mlp_parameters()), 'lr': opt.deformation_lr_init * s, "name": "deformation"},
            {'params': list(self._deformation.get_grid_parameters()), 'lr': opt.grid_lr_init * s, "name": "grid"},
            {'params': [self._features_dc], 'lr': opt.feature_lr, "name": "f_dc"},
            {'params': [self._features_rest], 'lr': opt.feature_lr / 20.0, "name": "f_rest"},
            {'params': [self._opacity], 'lr': opt.opacity_lr, "name": "opacity"},
            {'params': [self._scaling], 'lr': opt.scaling_lr, "name": "scaling"},
            {'params': [self._rotation], 'lr': opt.rotation_lr, "name": "rotation"}]
        self.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15)
        self._lr_args = {"xyz": (opt.position_lr_init * s, opt.position_lr_final * s),
                         "deformation": (opt.deformation_lr_init * s, opt.deformation_lr_final * s),
                         "grid": (opt.grid_lr_init * s, opt.grid_lr_final * s)}
        self._lr_max_steps = opt.position_lr_max_steps
        return self.optimizer

    def update_learning_rate(self, iteration):
        """scene/gaussian_model.py:284-298: per-iteration schedule of the xyz / grid / deformation groups (all three decay
        over position_lr_max_steps; the other groups keep their constant rates)."""
        for group in self.optimizer.param_groups:
            name = group["name"]
            key = "xyz" if name == "xyz" else ("grid" if "grid" in name else ("deformation" if name == "deformation" else None))
            if key is not None:
                lr_init, lr_final = self._lr_args[key]
                group["lr"] = expon_lr(iteration, lr_init, lr_final, self._lr_max_steps)


LAST_RASTER_STATE = None      # (R, geomBuffer, binningBuffer, imgBuffer) of the last no-grad render (bench.py counts pairs from it)


def render(cam, pc, bg_color, scaling_modifier=1.0, stage="fine", delta_scale=1, debug=False, shs=None):
    """gaussian_renderer/__init__.py:22-178 for a b200gs.synthetic.SynthCamera-like `cam`.
    `shs`: optional pre-concatenated [P,16,3] SH tensor standing in for `pc.get_features` (the trainer
    concatenates once per optimiser step instead of once per view; the values are identical)."""
    means3D = pc.get_xyz
    screenspace_points = torch.zeros_like(means3D, requires_grad=True)
    settings = GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=cam.tanfovx, tanfovy=cam.tanfovy,
        bg=bg_color, scale_modifier=scaling_modifier, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix,
        sh_degree=pc.active_sh_degree, campos=cam.campos, prefiltered=False, debug=debug)
    rasterizer = GaussianRasterizer(raster_settings=settings)
    opacity, scales, rotations = pc._opacity, pc._scaling, pc._rotation
    if shs is None:
        shs = pc.get_features
    if stage == "coarse":
        m3, sc, rt, op, sh = means3D, scales, rotations, opacity, shs
    else:
        # the reference repeats the camera's timestamp into a [P,1] tensor (gaussian_renderer/__init__.py:56); our field
        # takes the scalar itself (CUDA path only), which also tells its backward that the whole view shares one time
        time = float(cam.time) if means3D.is_cuda else torch.full((means3D.shape[0], 1), float(cam.time), device=means3D.device)
        m3, sc, rt, op, sh = pc._deformation(means3D, scales, rotations, opacity, shs, time, pc.get_flow,
                                             cam.frame_num, delta_scale)
    if sc.is_cuda:
        sc, rt, op = fusedops.activations(sc, rt, op)      # exp / F.normalize / sigmoid in one launch each way
    else:                                                   # CPU tensors only reach here from the gloo plumbing test
        sc = pc.scaling_activation(sc); rt = pc.rotation_activation(rt); op = pc.opacity_activation(op)
    image, radii, depth = rasterizer(means3D=m3, means2D=screenspace_points, shs=sh, colors_precomp=None,
                                     opacities=op, scales=sc, rotations=rt, cov3D_precomp=None)
    if not torch.is_grad_enabled():
        global LAST_RASTER_STATE
        LAST_RASTER_STATE = getattr(rasterizer, "last_state", None)
    return {"render": image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
            "depth": depth}


def render_frames(cams, pc, bg_color, stage="fine", **kw):
    """render() for every camera of a sequence over ONE static model (render_4DGS.py:41-76), as a generator: in the fine
    stage the time-independent half of the HexPlane field is evaluated once for the whole sequence
    (field.shared_spatial_product) instead of once per frame."""
    import contextlib
    from . import field as _field
    with contextlib.ExitStack() as stack:
        if stage == "fine" and pc.get_xyz.is_cuda:
            with torch.no_grad():
                stack.enter_context(_field.shared_spatial_product(pc._deformation, pc._xyz))
        for cam in cams:
            with torch.no_grad():              # per frame, not across the yield: the consumer keeps its own grad mode
                out = render(cam, pc, bg_color, stage=stage, **kw)
            yield out


def _receives_grad(name, stage):
    """Parameters the reference's loss reaches (SURVEY.md Appendix C iii): timenet, the opacity /
    SH heads and the aabb never do; in the coarse stage the whole deformation field is bypassed."""
    if name.startswith("_deformation."):
        if stage == "coarse":
            return False
        return not ("timenet" in name or "opacity_deform" in name or "shs_deform" in name or name.endswith("grid.aabb"))
    return True


class ViewParallelTrainer:
    """One optimiser step over a batch of views sharded across ranks (replicated model).

    `render_fn(cam, model, bg, stage)` defaults to this module's `render`; tests inject a CPU
    stand-in to exercise the sharding / flat-arena / collective logic over gloo."""

    def __init__(self, model, bg_color, stage="fine", process_group=None, world_size=1, rank=0, render_fn=None,
                 regulation=None, regulation_fn=None, shared_shs=None, overlap_sh_reduce=None):
        self.model = model
        self.bg = bg_color
        self.stage = stage
        # (time_smoothness_weight, l1_time_planes, plane_tv_weight) of train_4DGS.py:215-218, or None to leave it out
        self.regulation = regulation if stage == "fine" else None
        self.regulation_fn = regulation_fn if stage == "fine" else None      # PyTorch-op variant (the reference arm of bench.py)
        self.pg = process_group
        self.world_size = world_size
        self.rank = rank
        # our own render() takes the per-step SH tensor; an injected render_fn keeps the 4-argument form
        self.shared_shs = (render_fn is None) if shared_shs is None else bool(shared_shs)
        self.render_fn = render_fn or (lambda cam, m, bg, st, shs=None: render(cam, m, bg, stage=st, shs=shs))
        # Opt-in (B200GS_OVERLAP_SH_REDUCE=1, not yet measured): the SH gradient -- 192 of the 248 MB a rank contributes at 1M
        # Gaussians -- is final as soon as the LAST view's rasterizer backward has run, so its all-reduce starts there, on NCCL's
        # own stream, and overlaps that view's field backward and the deferred spatial pass; the rest of the arena (everything
        # but the SH slices, which then sit at its end) is reduced after the loop as before.
        if overlap_sh_reduce is None:
            overlap_sh_reduce = os.environ.get("B200GS_OVERLAP_SH_REDUCE") == "1"
        self.overlap_sh_reduce = bool(overlap_sh_reduce) and self.shared_shs and world_size > 1
        self._sh_work = None
        P = model.get_xyz.shape[0]
        # flat gradient arena: [every parameter the loss reaches | screen-space xy per Gaussian];
        # p.grad are views into it, so autograd accumulates in place and ONE collective (fp32 sum
        # over NVLink/NVSwitch) reduces everything. Parameters outside it keep grad None and the
        # optimiser skips them, as torch does in the reference.
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and _receives_grad(n, stage)]
        is_sh = lambda n: n in ("_features_dc", "_features_rest")
        if self.overlap_sh_reduce:                     # SH slices last: [other parameters | screen-space xy | f_dc | f_rest]
            named = [x for x in named if not is_sh(x[0])] + [x for x in named if is_sh(x[0])]
        self.trainable = [p for _, p in named]
        # every slice starts on a 256-byte boundary: the kernels use 128-bit loads / vector reductions on them
        al = lambda x: (x + 63) // 64 * 64
        n = sum(al(p.numel()) for p in self.trainable) + al(3 * P)
        self.arena = torch.zeros(n, dtype=torch.float32, device=model.get_xyz.device)
        self.views, off = [], 0
        self.viewspace_grad = None
        for name, p in named:
            if self.overlap_sh_reduce and is_sh(name) and self.viewspace_grad is None:
                self.viewspace_grad = self.arena[off:off + 3 * P].view(P, 3)
                off += al(3 * P)
                self._reduce_end = off                 # the post-loop all-reduce covers arena[:_reduce_end]
            flat = self.arena[off:off + p.numel()]
            if p.dim() == 4 and p.stride(1) == 1 and not p.is_contiguous():     # channels_last plane
                N, C, H, W = p.shape
                v = flat.view(N, H, W, C).permute(0, 3, 1, 2)
            else:
                v = flat.view(p.shape)
            self.views.append(v)
            off += al(p.numel())
        if self.viewspace_grad is None:
            self.viewspace_grad = self.arena[off:off + 3 * P].view(P, 3)
            self._reduce_end = n
        self.max_radii = torch.zeros(P, dtype=torch.int32, device=self.arena.device)

    def _bind(self):
        self.arena.zero_()
        self.max_radii.zero_()
        for p, v in zip(self.trainable, self.views):
            p.grad = v

    def _start_sh_reduce(self, shs):
        if self._sh_work is None and shs is not None and shs.grad is not None:
            import torch.distributed as dist
            self._sh_work = dist.all_reduce(shs.grad, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def local_views(self, n_global):
        """Indices of the global batch this rank renders: view b goes to rank b mod world_size."""
        return list(range(self.rank, n_global, self.world_size))

    def step(self, cams, gts, global_batch=None):
        """cams / gts: THIS rank's views. Loss = mean over the GLOBAL batch of per-view L1 means
        (train_4DGS.py:205-210: l1 over the concatenated [B,3,H,W] tensor). Returns the summed
        local loss (already scaled by 1/B) as a tensor."""
        B = global_batch or (len(cams) * self.world_size)
        self._bind()
        total = None
        shs = None
        if self.shared_shs:
            # get_features (scene/gaussian_model.py:136-140) is the same tensor for every view of the step: build it once
            # as a leaf, let the views' SH gradients accumulate in it, and split them back after the last view
            m = self.model
            shs = torch.cat((m._features_dc, m._features_rest), dim=1).detach().requires_grad_(True)
            if shs.is_cuda:
                # the rasterizer backward adds every view's SH gradient straight into this buffer (no per-view allocation,
                # no AccumulateGrad pass over 192 B per Gaussian)
                shs.grad = torch.zeros_like(shs)
        from . import field as _field
        share_spatial = self.shared_shs and self.stage == "fine" and len(cams) > 1 and self.model._xyz.is_cuda
        if share_spatial:
            _field.begin_shared_step(self.model._deformation, self.model._xyz)
        self._sh_work = None
        try:
            for vi, (cam, gt) in enumerate(zip(cams, gts)):
                pkg = self.render_fn(cam, self.model, self.bg, self.stage, shs) if self.shared_shs else \
                    self.render_fn(cam, self.model, self.bg, self.stage)
                _field.ACCUMULATE_INTO_GRAD = self.shared_shs     # p.grad are arena views: let the field kernels add into them
                _rast.SH_GRAD_ACCUMULATOR = shs.grad if (shs is not None and shs.is_cuda) else None
                if self.overlap_sh_reduce and vi == len(cams) - 1:
                    # called by the rasterizer backward right after it has queued the kernel that adds this view's SH gradient
                    _rast.AFTER_SH_ACCUMULATE = lambda: self._start_sh_reduce(shs)
                try:
                    if self.shared_shs and gt.is_cuda:
                        # fused L1 (utils/loss_utils.py:23-24) + its gradient, then backward from the image
                        if total is None:
                            total = torch.zeros(1, device=gt.device)
                        img = pkg["render"]
                        d_img = fusedops.l1_loss_and_grad(img, gt, 1.0 / (img.numel() * B), total)
                        img.backward(d_img)
                        loss = None
                    else:
                        loss = (pkg["render"] - gt).abs().mean() / B
                        loss.backward()
                finally:
                    _field.ACCUMULATE_INTO_GRAD = False
                    _rast.SH_GRAD_ACCUMULATOR = None
                    _rast.AFTER_SH_ACCUMULATE = None
                vg = pkg["viewspace_points"].grad
                if vg is not None:
                    self.viewspace_grad += vg
                torch.maximum(self.max_radii, pkg["radii"], out=self.max_radii)
                if loss is not None:
                    total = loss.detach() if total is None else total + loss.detach()
        except BaseException:
            _field.drop_shared()               # never leave a half-used spatial product behind (a later render() would reuse it)
            raise
        if share_spatial:
            _field.finish_shared_step()
        if self.overlap_sh_reduce and shs is not None:
            if shs.grad is None:                       # a rank without views still takes part in the collective
                shs.grad = torch.zeros_like(shs)
            self._start_sh_reduce(shs)                 # no-op when the rasterizer backward has already started it
            self._sh_work.wait()
            self._sh_work = None
        if shs is not None and shs.grad is not None:
            self.model._features_dc.grad += shs.grad[:, :1]
            self.model._features_rest.grad += shs.grad[:, 1:]
        if self.world_size > 1:
            import torch.distributed as dist
            # with the SH slices already reduced (above) only the front of the arena is left
            dist.all_reduce(self.arena[:self._reduce_end] if self.overlap_sh_reduce else self.arena, op=dist.ReduceOp.SUM, group=self.pg)
            dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=self.pg)
        if self.regulation is not None:
            # plane-only term, identical on every rank: added once, after the reduce
            _field.accumulate_regulation(self.model._deformation.deformation_net.grid, *self.regulation,
                                         loss_accum=total if (total is not None and total.numel() == 1 and total.dim() == 1) else None)
        if self.regulation_fn is not None:
            reg = self.regulation_fn()
            reg.backward()
            total = total + reg.detach()
        self.model.optimizer.step()
        return total.reshape(()) if total is not None else total
