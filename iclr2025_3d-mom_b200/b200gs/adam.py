"""Fused multi-tensor Adam behind the torch.optim.Optimizer interface.

Drop-in for the optimiser the reference builds at scene/gaussian_model.py:209
(`torch.optim.Adam(l, lr=0.0, eps=1e-15)`): same constructor defaults, same
`param_groups` / `state[p] = {"step", "exp_avg", "exp_avg_sq"}` layout (the reference's
densify / prune / reset_opacity code reaches into both, gaussian_model.py:409-482), same
arithmetic as torch's foreach path — but `step()` is one CUDA launch
over every parameter tensor (libb200gs `b200gs_adam_multi`).
"""
import ctypes

import torch
from torch.optim.optimizer import Optimizer

from . import _lib
from ._lib import check, current_stream


class _AdamTensor(ctypes.Structure):
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p),
                ("exp_avg_sq", ctypes.c_void_p), ("numel", ctypes.c_longlong), ("neg_step_size", ctypes.c_float),
                ("bias_correction2_sqrt", ctypes.c_float)]


_lib.register("b200gs_adam_multi", ctypes.c_int,
              [ctypes.c_int, ctypes.POINTER(_AdamTensor), ctypes.c_double, ctypes.c_double, ctypes.c_double,
               ctypes.c_void_p])


def _same_layout(a, b):
    return a.shape == b.shape and a.stride() == b.stride()


def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


class FusedAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise ValueError("FusedAdam implements the configuration the reference uses: "
                             "weight_decay=0, amsgrad=False")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        # one launch per distinct (betas, eps) (the reference has a single combination)
        buckets = {}
        keep = []
        touched = []
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            lr = float(group["lr"])
            eps = float(group["eps"])
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam needs CUDA parameters (there is no CPU path)")
                if p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam supports float32 parameters")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if not _dense(p):
                    raise RuntimeError("FusedAdam needs dense (contiguous or channels_last) parameters")
                if not _same_layout(m, p):
                    m = torch.empty_like(p).copy_(m); state["exp_avg"] = m
                if not _same_layout(v, p):
                    v = torch.empty_like(p).copy_(v); state["exp_avg_sq"] = v
                if not _same_layout(g, p):
                    g = torch.empty_like(p).copy_(g); keep.append(g)
                touched.append(p)
                state["step"] += 1
                step = float(state["step"])
                bc1 = 1 - beta1 ** step
                bc2 = 1 - beta2 ** step
                buckets.setdefault((beta1, beta2, eps), []).append(
                    (p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), (lr / bc1) * -1, bc2 ** 0.5))
        stream = current_stream()
        if buckets:
            # the kernel writes the parameters through raw pointers: tell autograd (saved-tensor checks) and every cache keyed
            # on `_version` (field._cell_order, the shared spatial HexPlane product) that they changed
            from . import field as _field
            _field.drop_shared()
            torch.autograd.graph.increment_version(touched)
        for (beta1, beta2, eps), items in buckets.items():
            arr = (_AdamTensor * len(items))()
            for i, it in enumerate(items):
                arr[i] = _AdamTensor(*it)
            check(L.b200gs_adam_multi(len(items), arr, beta1, beta2, eps, stream), "adam_multi")
        return loss
