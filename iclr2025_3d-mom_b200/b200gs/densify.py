"""Densify / prune bookkeeping on one kernel launch per event.

Drop-in bodies for `GaussianModel._prune_optimizer` and `GaussianModel.cat_tensors_to_optimizer`
(scene/gaussian_model.py:424-442, :461-482): same group iteration, same `optimizer.state`
re-attachment (the `step` counter survives, exp_avg / exp_avg_sq are gathered or zero-extended),
same returned `{group name: new nn.Parameter}` dict — but the ~18 boolean-mask / cat kernels per
event become ONE multi-tensor row gather (`b200gs_gather_rows_multi`) moving parameter and both
Adam moments together.  `patch_gaussian_model(cls)` installs them on the reference class.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, current_stream


class _GatherTensor(ctypes.Structure):
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("row_floats", ctypes.c_int), ("reserved", ctypes.c_int)]


_lib.register("b200gs_gather_rows_multi", ctypes.c_int,
              [ctypes.c_int, ctypes.POINTER(_GatherTensor), ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p])


def gather_rows(tensors, index, n_out, out_rows=None):
    """For every float32 CUDA tensor t ([N, ...], contiguous): returns new tensors whose first n_out rows
    are t[index[i]] (index=None: identity copy). out_rows >= n_out sizes the outputs (extra rows are
    left for the caller to fill)."""
    out_rows = n_out if out_rows is None else out_rows
    outs, arr = [], (_GatherTensor * len(tensors))()
    for i, t in enumerate(tensors):
        if t.dtype != torch.float32 or not t.is_cuda:
            raise RuntimeError("gather_rows works on float32 CUDA tensors")
        t = t.contiguous()
        o = torch.empty((out_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        row = int(t[0].numel()) if t.shape[0] > 0 else int(torch.Size(t.shape[1:]).numel())
        arr[i] = _GatherTensor(t.data_ptr(), o.data_ptr(), row, 0)
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


def patch_gaussian_model(cls):
    """Install the fused bookkeeping on the reference's GaussianModel class (methods keep their names
    and signatures; everything that calls them — prune_points, densification_postfix — is unchanged)."""
    cls._prune_optimizer = lambda self, mask: prune_optimizer(self.optimizer, mask)
    cls.cat_tensors_to_optimizer = lambda self, tensors_dict: cat_tensors_to_optimizer(self.optimizer, tensors_dict)
    return cls
