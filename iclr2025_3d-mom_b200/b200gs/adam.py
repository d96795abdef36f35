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


_lib.register("b200gs_adam_sh", ctypes.c_int,
              [ctypes.c_longlong, ctypes.c_int, ctypes.POINTER(_AdamTensor), ctypes.POINTER(_AdamTensor), ctypes.c_void_p,
               ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_void_p])


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

    def _group_of(self, p):
        for group in self.param_groups:
            if any(q is p for q in group["params"]):
                return group
        raise KeyError("parameter is not in this optimiser")

    @torch.no_grad()
    def step_sh(self, f_dc, f_rest, sh_grad):
        """The step of the two SH parameters ([P,1,3] and [P,M-1,3], scene/gaussian_model.py:136-140) with their gradient read
        from ONE [P,M,3] buffer (what the rasterizer backward writes): same arithmetic and state as `step()`, one launch, no
        split of the gradient.  The trainer calls it on a side stream as soon as the last view's SH gradient is complete and
        passes `skip=(f_dc, f_rest)` to the `step()` that follows."""
        P, M = int(sh_grad.shape[0]), int(sh_grad.shape[1])
        if not (sh_grad.is_cuda and sh_grad.dtype == torch.float32 and sh_grad.is_contiguous() and tuple(f_dc.shape) == (P, 1, 3)
                and tuple(f_rest.shape) == (P, M - 1, 3) and f_dc.is_contiguous() and f_rest.is_contiguous()):
            raise RuntimeError("step_sh: expected contiguous float32 CUDA tensors [P,1,3], [P,M-1,3] and a [P,M,3] gradient")
        descs = []
        betas_eps = None
        for p in (f_dc, f_rest):
            group = self._group_of(p)
            state = self.state[p]
            if len(state) == 0:
                state["step"] = torch.tensor(0.0, dtype=torch.float32)
                state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            state["step"] += 1
            step = float(state["step"])
            beta1, beta2 = group["betas"]
            if betas_eps is None:
                betas_eps = (beta1, beta2, float(group["eps"]))
            elif betas_eps != (beta1, beta2, float(group["eps"])):
                raise RuntimeError("step_sh: the two SH groups must share betas / eps")
            descs.append(_AdamTensor(p.data_ptr(), None, state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr(), p.numel(),
                                     (float(group["lr"]) / (1 - beta1 ** step)) * -1, (1 - beta2 ** step) ** 0.5))
        check(_lib.lib().b200gs_adam_sh(P, M, ctypes.byref(descs[0]), ctypes.byref(descs[1]), sh_grad.data_ptr(), *betas_eps,
                                        current_stream()), "adam_sh")
        torch.autograd.graph.increment_version([f_dc, f_rest])

    @torch.no_grad()
    def step(self, closure=None, skip=()):
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
                if g is None or any(p is q for q in skip):
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
                    # e.g. the reference's _xyz right after create_from_pcd: torch.tensor(pcd_points.T) keeps numpy's
                    # column-major strides (scene/gaussian_model.py:153) until the first densify / prune re-builds it.
                    # Same Parameter object, same values, row-major storage from here on (moments follow below).
                    p.data = p.data.contiguous()
                    if g is not None and not _same_layout(g, p):
                        g = g.contiguous()
                        p.grad = g
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
        _lib.COUNTERS["adam_steps"] += 1
        import time as _time
        now = _time.perf_counter()
        _lib.TIMES.setdefault("first_adam_step", now)
        _lib.TIMES["last_adam_step"] = now
        for (beta1, beta2, eps), items in buckets.items():
            arr = (_AdamTensor * len(items))()
            for i, it in enumerate(items):
                arr[i] = _AdamTensor(*it)
            check(L.b200gs_adam_multi(len(items), arr, beta1, beta2, eps, stream), "adam_multi")
        return loss
