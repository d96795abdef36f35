"""Densify / prune bookkeeping: one decision kernel + one multi-tensor row gather per event.

Drop-in bodies for the reference's `GaussianModel` methods (scene/gaussian_model.py):
  * `_prune_optimizer` / `cat_tensors_to_optimizer` (:424-442, :461-482): same group iteration, same `optimizer.state`
    re-attachment (the `step` counter survives, exp_avg / exp_avg_sq are gathered or zero-extended), same returned
    `{group name: new nn.Parameter}` dict -- ONE multi-tensor row gather (`b200gs_gather_rows_multi`) instead of ~18
    boolean-mask / cat kernels;
  * `densify` (:693-698 -> densify_and_clone :541-565 + densify_and_split :511-539 + two densification_postfix :484-509 +
    prune_points :444-459): the clone / split flags come from one pass over the Gaussians (`b200gs_densify_select`), the final
    row order [kept originals | clones | split children x 2] is built as ONE index and every parameter, both Adam moments,
    `_scene_flow` and the bookkeeping tensors move in ONE gather; the split offsets are drawn by the same `torch.normal`
    call on the same shapes, so a run with the reference's seed consumes the RNG identically;
  * `prune` (:681-690), `add_densification_stats` (:713-715), `reset_opacity` (:362-365): one kernel each (+ the gather).
`patch_gaussian_model(cls)` installs them on the reference class; `b200gs.engine.GaussianState` uses the same functions.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, current_stream


class _GatherTensor(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("row_floats", ctypes.c_int), ("zero_tail_rows", ctypes.c_int)]


_lib.register("b200gs_gather_rows_multi", ctypes.c_int,
              [ctypes.c_int, ctypes.POINTER(_GatherTensor), ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p])


_P = ctypes.c_void_p
_lib.register("b200gs_densify_select", ctypes.c_int, [ctypes.c_longlong, _P, _P, _P, ctypes.c_float, ctypes.c_float, _P, _P, _P])
_lib.register("b200gs_prune_select", ctypes.c_int, [ctypes.c_longlong, _P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_float, _P, _P])
_lib.register("b200gs_densification_stats", ctypes.c_int, [ctypes.c_longlong, _P, _P, _P, _P, _P])
_lib.register("b200gs_reset_opacity", ctypes.c_int, [ctypes.c_longlong, _P, _P, _P])


def gather_rows(tensors, index, n_out, out_rows=None, zero_tail=None):
    """For every float32 CUDA tensor t ([N, ...], contiguous): returns new tensors whose first n_out rows
    are t[index[i]] (index=None: identity copy). out_rows >= n_out sizes the outputs (extra rows are
    left for the caller to fill).  zero_tail[i] > 0: the last zero_tail[i] of the n_out rows of tensor i are zeros."""
    out_rows = n_out if out_rows is None else out_rows
    outs, arr = [], (_GatherTensor * len(tensors))()
    for i, t in enumerate(tensors):
        if t.dtype != torch.float32 or not t.is_cuda:
            raise RuntimeError("gather_rows works on float32 CUDA tensors")
        t = t.contiguous()
        o = torch.empty((out_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        row = int(t[0].numel()) if t.shape[0] > 0 else int(torch.Size(t.shape[1:]).numel())
        arr[i] = _GatherTensor(t.data_ptr(), o.data_ptr(), row, int(zero_tail[i]) if zero_tail is not None else 0)
        outs.append(o)
        tensors[i] = t
    if n_out > 0 and len(tensors) > 0:
        check(_lib.lib().b200gs_gather_rows_multi(len(tensors), arr, index.data_ptr() if index is not None else None,
                                                  n_out, current_stream()), "gather_rows_multi")
    return outs


def prune_optimizer(optimizer, mask):
    """gaussian_model.py:424-442."""
    _lib.COUNTERS["prune_events"] += 1
    index = torch.nonzero(mask, as_tuple=False).reshape(-1).contiguous()          # int64 rows to keep, ascending
    n_out = int(index.numel())
    jobs = []
    for group in optimizer.param_groups:
        if len(group["params"]) > 1:
            continue
        p = group["params"][0]
        st = optimizer.state.get(p, None)
        jobs.append((group, p, st))
    srcs = []
    for group, p, st in jobs:
        srcs.append(p.detach())
        if st is not None:
            srcs += [st["exp_avg"], st["exp_avg_sq"]]
    outs = gather_rows(srcs, index, n_out)
    optimizable, k = {}, 0
    for group, p, st in jobs:
        new_p = nn.Parameter(outs[k].requires_grad_(True)); k += 1
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = outs[k], outs[k + 1]; k += 2
            del optimizer.state[p]
            group["params"][0] = new_p
            optimizer.state[new_p] = st
        else:
            group["params"][0] = new_p
        optimizable[group["name"]] = new_p
    return optimizable


def cat_tensors_to_optimizer(optimizer, tensors_dict):
    """gaussian_model.py:461-482."""
    _lib.COUNTERS["densify_cat_events"] += 1
    jobs = []
    for group in optimizer.param_groups:
        if len(group["params"]) > 1:
            continue
        p = group["params"][0]
        jobs.append((group, p, optimizer.state.get(p, None), tensors_dict[group["name"]]))
    srcs = []
    for group, p, st, ext in jobs:
        srcs.append(p.detach())
        if st is not None:
            srcs += [st["exp_avg"], st["exp_avg_sq"]]
    n_old = int(jobs[0][1].shape[0]) if jobs else 0
    n_new = int(jobs[0][3].shape[0]) if jobs else 0
    outs = gather_rows(srcs, None, n_old, out_rows=n_old + n_new)
    optimizable, k = {}, 0
    for group, p, st, ext in jobs:
        outs[k][n_old:] = ext
        new_p = nn.Parameter(outs[k].requires_grad_(True)); k += 1
        if st is not None:
            outs[k][n_old:] = 0; outs[k + 1][n_old:] = 0
            st["exp_avg"], st["exp_avg_sq"] = outs[k], outs[k + 1]; k += 2
            del optimizer.state[p]
            group["params"][0] = new_p
            optimizer.state[new_p] = st
        else:
            group["params"][0] = new_p
        optimizable[group["name"]] = new_p
    return optimizable


# ---- whole events ---------------------------------------------------------------------------------------------------------
_PARAM_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
               "rotation": "_rotation"}


def _single_groups(optimizer):
    return [(g, g["params"][0], optimizer.state.get(g["params"][0], None)) for g in optimizer.param_groups
            if len(g["params"]) == 1 and g["name"] in _PARAM_ATTR]


def _regather_model(model, index, n_out, zero_tail_rows, extra):
    """Moves every per-Gaussian parameter, its Adam moments and the float tensors in `extra` ({attr: tensor}) through ONE gather:
    out[i] = src[index[i]]; the last `zero_tail_rows` rows of the moments are zeros.  Re-attaches parameters / optimiser state the
    way _prune_optimizer / cat_tensors_to_optimizer do.  Returns {attr: gathered tensor} for `extra`."""
    opt = model.optimizer
    jobs = _single_groups(opt)
    srcs, zt = [], []
    for group, p, st in jobs:
        srcs.append(p.detach()); zt.append(0)
        if st is not None:
            srcs += [st["exp_avg"], st["exp_avg_sq"]]; zt += [zero_tail_rows, zero_tail_rows]
    names = list(extra.keys())
    for n in names:
        srcs.append(extra[n]); zt.append(0)
    outs = gather_rows(srcs, index, n_out, zero_tail=zt)
    k = 0
    for group, p, st in jobs:
        new_p = nn.Parameter(outs[k].requires_grad_(True)); k += 1
        if st is not None:
            st["exp_avg"], st["exp_avg_sq"] = outs[k], outs[k + 1]; k += 2
            del opt.state[p]
            group["params"][0] = new_p
            opt.state[new_p] = st
        else:
            group["params"][0] = new_p
        setattr(model, _PARAM_ATTR[group["name"]], new_p)
    return {n: outs[k + i] for i, n in enumerate(names)}


def _flags(N, device):
    return torch.empty((N,), dtype=torch.bool, device=device)


def densify(model, max_grad, min_opacity, extent, max_screen_size, density_threshold=None, displacement_scale=None, model_path=None,
            iteration=None, stage=None):
    """GaussianModel.densify (scene/gaussian_model.py:693-698): same final tensors, same RNG consumption."""
    xyz = model._xyz
    N = int(xyz.shape[0])
    if not (float(max_grad) > 0.0) or N == 0:
        raise RuntimeError("fused densify needs a positive gradient threshold and a non-empty model")
    dev = xyz.device
    _lib.COUNTERS["densify_cat_events"] += 1
    clone_f, split_f = _flags(N, dev), _flags(N, dev)
    dense_extent = float(model.percent_dense * extent)
    accum, denom = model.xyz_gradient_accum.contiguous(), model.denom.contiguous()
    scaling = model._scaling.detach().contiguous()
    check(_lib.lib().b200gs_densify_select(N, accum.data_ptr(), denom.data_ptr(), scaling.data_ptr(), float(max_grad), dense_extent,
                                           clone_f.data_ptr(), split_f.data_ptr(), current_stream()), "densify_select")
    split_idx = torch.nonzero(split_f, as_tuple=False).reshape(-1)
    clone_idx = torch.nonzero(clone_f, as_tuple=False).reshape(-1)
    keep_idx = torch.nonzero(~split_f, as_tuple=False).reshape(-1)
    ns, nc = int(split_idx.numel()), int(clone_idx.numel())
    index = torch.cat((keep_idx, clone_idx, split_idx, split_idx)).contiguous()
    n_out = int(index.numel())
    children = None
    if ns > 0:
        # densify_and_split (:524-531), on the split rows only; torch.normal is called exactly as the reference calls it
        s_sel = torch.exp(scaling[split_idx])                                     # get_scaling[selected_pts_mask]
        stds = s_sel.repeat(2, 1)
        samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds)
        r = model._rotation.detach()[split_idx]
        q = r / torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])[:, None]
        w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                         2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), dim=1).reshape(-1, 3, 3)
        new_xyz = torch.bmm(R.repeat(2, 1, 1), samples.unsqueeze(-1)).squeeze(-1) + xyz.detach()[split_idx].repeat(2, 1)
        new_scaling = torch.log(stds / (0.8 * 2))
        children = (new_xyz, new_scaling)
    extra = {"_scene_flow": model._scene_flow.contiguous()}
    moved = _regather_model(model, index, n_out, nc + 2 * ns, extra)
    model._scene_flow = moved["_scene_flow"]
    model._deformation_table = model._deformation_table[index]
    if children is not None:
        with torch.no_grad():
            model._xyz[n_out - 2 * ns:] = children[0]
            model._scaling[n_out - 2 * ns:] = children[1]
    # densification_postfix (:504-507): the statistics restart from zero at the new size
    model.xyz_gradient_accum = torch.zeros((n_out, 1), device=dev)
    model._deformation_accum = torch.zeros((n_out, 3), device=dev)
    model.denom = torch.zeros((n_out, 1), device=dev)
    model.max_radii2D = torch.zeros((n_out,), device=dev)


def prune(model, max_grad, min_opacity, extent, max_screen_size):
    """GaussianModel.prune (scene/gaussian_model.py:681-692) -> prune_points (:444-459)."""
    N = int(model._xyz.shape[0])
    dev = model._xyz.device
    _lib.COUNTERS["prune_events"] += 1
    flag = _flags(N, dev)
    mss = float(max_screen_size) if max_screen_size else 0.0
    radii = model.max_radii2D.to(torch.float32).contiguous()
    check(_lib.lib().b200gs_prune_select(N, model._opacity.detach().contiguous().data_ptr(), model._scaling.detach().contiguous().data_ptr(),
                                         radii.data_ptr(), float(min_opacity), mss, float(0.1 * extent), flag.data_ptr(), current_stream()),
          "prune_select")
    prune_points(model, flag)


def prune_points(model, mask):
    """GaussianModel.prune_points (scene/gaussian_model.py:444-459): drop the rows where mask is True."""
    index = torch.nonzero(~mask, as_tuple=False).reshape(-1).contiguous()
    n_out = int(index.numel())
    f32 = lambda t: t.to(torch.float32).contiguous()
    extra = {"_deformation_accum": f32(model._deformation_accum), "xyz_gradient_accum": f32(model.xyz_gradient_accum), "denom": f32(model.denom),
             "max_radii2D": f32(model.max_radii2D), "_scene_flow": f32(model._scene_flow)}
    moved = _regather_model(model, index, n_out, 0, extra)
    for k, v in moved.items():
        setattr(model, k, v)
    model._deformation_table = model._deformation_table[index]


def add_densification_stats(model, viewspace_point_tensor, update_filter):
    """GaussianModel.add_densification_stats (scene/gaussian_model.py:713-715)."""
    N = int(model.xyz_gradient_accum.shape[0])
    g = viewspace_point_tensor.detach()
    if not (g.is_cuda and g.dtype == torch.float32 and g.dim() == 2 and g.shape[1] >= 2 and update_filter.dtype == torch.bool):
        raise RuntimeError("add_densification_stats needs a float32 CUDA [N,>=2] gradient and a bool filter")
    if g.shape[1] != 3 or not g.is_contiguous():
        g3 = torch.zeros((N, 3), device=g.device); g3[:, :2] = g[:, :2]; g = g3
    if not (model.xyz_gradient_accum.is_contiguous() and model.denom.is_contiguous()):
        model.xyz_gradient_accum = model.xyz_gradient_accum.contiguous(); model.denom = model.denom.contiguous()
    f = update_filter.contiguous()
    check(_lib.lib().b200gs_densification_stats(N, g.data_ptr(), f.data_ptr(), model.xyz_gradient_accum.data_ptr(), model.denom.data_ptr(),
                                                current_stream()), "densification_stats")


def reset_opacity(model):
    """GaussianModel.reset_opacity (scene/gaussian_model.py:362-365) + replace_tensor_to_optimizer (:409-422)."""
    old = model._opacity
    new = torch.empty_like(old.detach(), memory_format=torch.contiguous_format)
    check(_lib.lib().b200gs_reset_opacity(old.numel(), old.detach().contiguous().data_ptr(), new.data_ptr(), current_stream()), "reset_opacity")
    opt = model.optimizer
    for group in opt.param_groups:
        if group["name"] == "opacity":
            st = opt.state.get(group["params"][0], None)
            new_p = nn.Parameter(new.requires_grad_(True))
            if st is not None:
                st["exp_avg"] = torch.zeros_like(new)
                st["exp_avg_sq"] = torch.zeros_like(new)
                del opt.state[group["params"][0]]
                opt.state[new_p] = st
            group["params"][0] = new_p
            model._opacity = new_p


def warmup(model, growth=1.06):
    """Makes the FIRST densification / pruning event as cheap as the later ones (at 5M Gaussians it cost 131 ms against 12.5 ms:
    first-use kernel loading plus cudaMalloc of ~3.6 GB of new parameter / moment buffers).  (1) Every kernel and torch operator of
    an event runs once on a 64-row toy; (2) one buffer per per-Gaussian tensor that an event re-creates, `growth` times its
    current size, is allocated and released, so the caching allocator already owns blocks of the right size class when the
    event asks for them (a standing reserve of about one model copy: 0.7 KB per Gaussian, nothing on a 180 GB device)."""
    dev = model._xyz.device
    if dev.type != "cuda":
        return
    n = 64
    L = _lib.lib()
    st = current_stream()
    z = lambda *sh: torch.zeros(*sh, device=dev)
    acc, den, sc, op, rad = z(n, 1) + 1e-3, z(n, 1) + 1.0, z(n, 3) - 4.0, z(n, 1), z(n)
    f1, f2 = _flags(n, dev), _flags(n, dev)
    check(L.b200gs_densify_select(n, acc.data_ptr(), den.data_ptr(), sc.data_ptr(), 0.0002, 0.01, f1.data_ptr(), f2.data_ptr(), st), "densify_select")
    check(L.b200gs_prune_select(n, op.data_ptr(), sc.data_ptr(), rad.data_ptr(), 0.005, 20.0, 0.1, f1.data_ptr(), st), "prune_select")
    check(L.b200gs_densification_stats(n, z(n, 3).data_ptr(), f2.data_ptr(), acc.data_ptr(), den.data_ptr(), st), "densification_stats")
    check(L.b200gs_reset_opacity(n, op.data_ptr(), z(n, 1).data_ptr(), st), "reset_opacity")
    idx = torch.nonzero(~f1, as_tuple=False).reshape(-1)
    idx = torch.cat((idx, idx[:4], idx[:2], idx[:2])).contiguous()
    gather_rows([z(n, 3), z(n, 48)], idx, int(idx.numel()), zero_tail=[0, 3])
    std = torch.exp(sc[:4]).repeat(2, 1)
    smp = torch.normal(mean=torch.zeros((8, 3), device=dev), std=std)
    torch.log(std / 1.6); torch.bmm(z(8, 3, 3), smp.unsqueeze(-1)); f1[idx[:3]]
    rows = int(model._xyz.shape[0] * growth) + 1024
    hold = []
    for group, p, stt in _single_groups(model.optimizer):
        per = int(p[0].numel()) if p.shape[0] else 1
        for _ in range(3 if stt is not None or True else 1):          # parameter + exp_avg + exp_avg_sq
            hold.append(torch.empty((rows, per), device=dev))
    hold.append(torch.empty((rows, 3), device=dev))                    # _scene_flow
    del hold


def patch_gaussian_model(cls):
    """Install the fused bookkeeping on the reference's GaussianModel class (methods keep their names
    and signatures; everything that calls them — prune_points, densification_postfix — is unchanged)."""
    cls._prune_optimizer = lambda self, mask: prune_optimizer(self.optimizer, mask)
    cls.cat_tensors_to_optimizer = lambda self, tensors_dict: cat_tensors_to_optimizer(self.optimizer, tensors_dict)
    # whole events (B200GS_FUSED_DENSIFY=0 keeps the reference's own methods on top of the two gather helpers above)
    import os
    if os.environ.get("B200GS_FUSED_DENSIFY", "1") != "0":
        orig_densify = cls.densify

        def _densify(self, max_grad, min_opacity, extent, max_screen_size, *a, **k):
            if self._xyz.is_cuda and float(max_grad) > 0.0 and self._xyz.shape[0] > 0:
                return densify(self, max_grad, min_opacity, extent, max_screen_size, *a, **k)
            return orig_densify(self, max_grad, min_opacity, extent, max_screen_size, *a, **k)
        cls.densify = _densify
        cls.prune = lambda self, max_grad, min_opacity, extent, max_screen_size: prune(self, max_grad, min_opacity, extent, max_screen_size)
        cls.prune_points = lambda self, mask: prune_points(self, mask)
        cls.add_densification_stats = lambda self, viewspace_point_tensor, update_filter: add_densification_stats(self, viewspace_point_tensor, update_filter)
        cls.reset_opacity = lambda self: reset_opacity(self)
    return cls
