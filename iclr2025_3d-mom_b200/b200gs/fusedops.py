"""Fused elementwise pieces of the training loop (csrc/elementwise.cu) for b200gs.engine.

The reference spells these with stock PyTorch ops inside code that runs unchanged on top of the
drop-ins (gaussian_renderer/__init__.py:130-132 activations; utils/loss_utils.py:23-24 L1), so they
stay PyTorch there. Our own trainer (`engine.render` / `ViewParallelTrainer`) calls these instead:
one launch each way for exp / normalize / sigmoid, one launch for the loss and its gradient.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, current_stream

_P = ctypes.c_void_p
_lib.register("b200gs_activations_forward", ctypes.c_int, [ctypes.c_longlong, _P, _P, _P, _P, _P, _P, _P])
_lib.register("b200gs_activations_backward", ctypes.c_int, [ctypes.c_longlong] + [_P] * 10)
_lib.register("b200gs_l1_loss_fwd_bwd", ctypes.c_int, [ctypes.c_longlong, _P, _P, ctypes.c_float, _P, _P, _P, _P])
_lib.register("b200gs_l1_loss_fwd_bwd_u8", ctypes.c_int, [ctypes.c_int, ctypes.c_int, _P, _P, ctypes.c_float, _P, _P, _P, _P])


def _req(t, shape_tail):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("fused activations need float32 CUDA tensors (there is no CPU path)")
    if t.dim() != 2 or t.shape[1] != shape_tail:
        raise RuntimeError(f"expected a [P,{shape_tail}] tensor, got {tuple(t.shape)}")
    return t.contiguous()


class _Activations(torch.autograd.Function):
    """(exp(scales), F.normalize(rotations), sigmoid(opacity)) — scene/gaussian_model.py:37-47."""

    @staticmethod
    def forward(ctx, scales, rotations, opacity):
        s, r, o = _req(scales, 3), _req(rotations, 4), _req(opacity, 1)
        P = int(s.shape[0])
        so, ro, oo = torch.empty_like(s), torch.empty_like(r), torch.empty_like(o)
        check(_lib.lib().b200gs_activations_forward(P, s.data_ptr(), r.data_ptr(), o.data_ptr(), so.data_ptr(), ro.data_ptr(),
                                                    oo.data_ptr(), current_stream()), "activations_forward")
        ctx.save_for_backward(so, r, oo)
        return so, ro, oo

    @staticmethod
    def backward(ctx, gs, gr, go):
        so, r, oo = ctx.saved_tensors
        P = int(so.shape[0])
        need = ctx.needs_input_grad
        ds = torch.empty_like(so) if need[0] else None
        dr = torch.empty_like(r) if need[1] else None
        do = torch.empty_like(oo) if need[2] else None
        # contiguous copies of the upstream gradients stay referenced until the launch has been queued
        gs, gr, go = (t.contiguous() if t is not None else None for t in (gs, gr, go))
        p = lambda t: t.data_ptr() if t is not None else None
        check(_lib.lib().b200gs_activations_backward(P, so.data_ptr(), r.data_ptr(), oo.data_ptr(), p(gs), p(gr), p(go),
                                                     p(ds), p(dr), p(do), current_stream()), "activations_backward")
        return ds, dr, do


def activations(scales, rotations, opacity):
    return _Activations.apply(scales, rotations, opacity)


def psnr_from_sse(sse, n):
    """utils/image_utils.py:17-38 from the per-image sums of squared errors the L1 kernel leaves behind:
    20 log10(1 / sqrt(mse)) per image (a device tensor; reading it is the caller's sync)."""
    return 20.0 * torch.log10(1.0 / torch.sqrt(sse / float(n)))


def l1_loss_and_grad(render, target, scale, loss_accum, sse_accum=None):
    """loss_accum[0] += scale * sum|render - target|; returns d(loss)/d(render) (= scale * sign).
    sse_accum (optional, a 1-element float32 CUDA tensor the caller zeroes): += sum (render - target)^2, see psnr_from_sse.
    target: float32 [3,H,W], or the dataset's own uint8 [H,W,3] image (converted on the device as PILtoTorch does: u8 / 255)."""
    sse = None
    if sse_accum is not None:
        if not (sse_accum.is_cuda and sse_accum.dtype == torch.float32 and sse_accum.numel() >= 1):
            raise RuntimeError("l1_loss_and_grad: sse_accum must be a float32 CUDA tensor")
        sse = sse_accum.data_ptr()
    if target.dtype == torch.uint8:
        if not (render.is_cuda and target.is_cuda and render.dtype == torch.float32 and render.dim() == 3 and render.shape[0] == 3
                and tuple(target.shape) == (render.shape[1], render.shape[2], 3)):
            raise RuntimeError("l1_loss_and_grad: a uint8 target must be [H,W,3] on the GPU for a float32 [3,H,W] render")
        r, t = render.detach().contiguous(), target.contiguous()
        d = torch.empty_like(r)
        check(_lib.lib().b200gs_l1_loss_fwd_bwd_u8(int(r.shape[1]), int(r.shape[2]), r.data_ptr(), t.data_ptr(), float(scale),
                                                   loss_accum.data_ptr(), sse, d.data_ptr(), current_stream()), "l1_loss_u8")
        return d
    if not (render.is_cuda and target.is_cuda and render.dtype == torch.float32 and target.dtype == torch.float32):
        raise RuntimeError("l1_loss_and_grad needs float32 CUDA tensors (there is no CPU path)")
    if render.shape != target.shape:
        raise RuntimeError("render / target shape mismatch")
    r, t = render.detach().contiguous(), target.contiguous()
    d = torch.empty_like(r)
    check(_lib.lib().b200gs_l1_loss_fwd_bwd(r.numel(), r.data_ptr(), t.data_ptr(), float(scale), loss_accum.data_ptr(), sse,
                                            d.data_ptr(), current_stream()), "l1_loss")
    return d
