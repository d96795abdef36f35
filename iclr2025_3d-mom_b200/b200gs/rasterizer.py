"""Drop-in for the reference's `diff_gaussian_rasterization` package.

Mirrors RAST/diff_gaussian_rasterization/__init__.py (GaussianRasterizationSettings :158-170,
GaussianRasterizer :172-221, _RasterizeGaussians :44-156, rasterize_gaussians :21-42) and the
pybind module `_C` (RAST/ext.cpp:15-19, RAST/rasterize_points.cu) — same names, argument
order, defaults, return order (color, radii, depth) and error behaviour — on top of the C ABI.
"""
import ctypes
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, current_stream, ptr

NUM_CHANNELS = 3

# Set by b200gs.engine.ViewParallelTrainer for the duration of a step: a [P,M,3] buffer that the backward ADDS the SH
# gradient into (b200gs_rast_backward_accumulate_sh); autograd then gets None for `sh`. None = ordinary behaviour.
SH_GRAD_ACCUMULATOR = None
# Optional callable the backward invokes right after queueing the kernel that adds into SH_GRAD_ACCUMULATOR (the trainer starts
# the SH gradient's all-reduce there on the last view of a step); None = nothing.
AFTER_SH_ACCUMULATE = None
# True for the first view of a step: the backward WRITES the accumulator (b200gs_rast_backward fills every row, zeros for culled
# Gaussians) instead of adding into it, so the buffer never needs a zero fill.
SH_GRAD_OVERWRITE = False

_pinned = {}


def _host_counters(device):
    key = (device.index if device.index is not None else torch.cuda.current_device())
    t = _pinned.get(key)
    if t is None:
        t = torch.zeros(2, dtype=torch.int64).pin_memory()
        _pinned[key] = t
    return t


def _f32c(t):
    if t is None:
        return None
    if t.numel() == 0:
        return t
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected a float32 tensor, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError("expected a CUDA tensor (the rasterizer has no CPU path)")
    return t.contiguous()


def _sizes(P, R, W, H):
    out = (ctypes.c_size_t * 3)()
    check(_lib.lib().b200gs_rast_buffer_sizes(P, R, W, H, out), "rast_buffer_sizes")
    return out[0], out[1], out[2]


class _CModule:
    """Function-for-function stand-in for the reference's compiled `_C` module."""

    @staticmethod
    def rasterize_gaussians(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                            cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                            image_width, sh, degree, campos, prefiltered, debug):
        # rasterize_points.cu:56-59
        if means3D.ndimension() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        L = _lib.lib()
        _lib.COUNTERS["raster_forward_calls"] += 1
        P = int(means3D.size(0))
        H, W = int(image_height), int(image_width)
        dev = means3D.device
        means3D = _f32c(means3D); colors = _f32c(colors); opacity = _f32c(opacity)
        scales = _f32c(scales); rotations = _f32c(rotations); cov3D_precomp = _f32c(cov3D_precomp)
        viewmatrix = _f32c(viewmatrix); projmatrix = _f32c(projmatrix); campos = _f32c(campos)
        background = _f32c(background); sh = _f32c(sh)
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0) else 0
        out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
        out_depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        stream = current_stream()
        rendered = 0
        visible = 0
        gbytes, _, ibytes = _sizes(P, 0, W, H)
        geomBuffer = torch.empty((gbytes,), dtype=torch.uint8, device=dev)
        imgBuffer = torch.empty((ibytes,), dtype=torch.uint8, device=dev)
        if P != 0:
            hc = _host_counters(dev)
            check(L.b200gs_rast_forward_stage1(
                P, int(degree), M, W, H, ptr(means3D), ptr(sh), ptr(colors), ptr(opacity), ptr(scales),
                float(scale_modifier), ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix),
                ptr(campos), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)), ptr(radii),
                geomBuffer.data_ptr(), gbytes, hc.data_ptr(), stream), "rasterize_gaussians")
            rendered, visible = int(hc[0]), int(hc[1])
        _, bbytes, _ = _sizes(P, rendered, W, H)
        # the instance count differs from view to view: round the request up to a few size classes (1/8 octave above 8 MiB)
        # so that the caching allocator re-uses the blocks of earlier views instead of growing by a cudaMalloc every time a
        # slightly larger count shows up (at 5M Gaussians / 4K that was a 30-80 ms stall every few iterations)
        if bbytes > (8 << 20):
            step = 1 << (int(bbytes).bit_length() - 4)
            bbytes = (bbytes + step - 1) // step * step
        binningBuffer = torch.empty((bbytes,), dtype=torch.uint8, device=dev)
        check(L.b200gs_rast_forward_stage2(
            P, rendered, visible, W, H, ptr(background), geomBuffer.data_ptr(), binningBuffer.data_ptr(), bbytes,
            imgBuffer.data_ptr(), ibytes, out_color.data_ptr(), out_depth.data_ptr(), stream),
            "rasterize_gaussians")
        if debug:
            torch.cuda.synchronize(dev)     # surface asynchronous CUDA errors here, like CHECK_CUDA(debug)
        return rendered, out_color, out_depth, radii, geomBuffer, binningBuffer, imgBuffer

    @staticmethod
    def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                     cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                     dL_dout_depth, sh, degree, campos, geomBuffer, R, binningBuffer,
                                     imageBuffer, debug):
        L = _lib.lib()
        P = int(means3D.size(0))
        H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
        dev = means3D.device
        means3D = _f32c(means3D); colors = _f32c(colors); scales = _f32c(scales)
        rotations = _f32c(rotations); cov3D_precomp = _f32c(cov3D_precomp)
        viewmatrix = _f32c(viewmatrix); projmatrix = _f32c(projmatrix); campos = _f32c(campos)
        background = _f32c(background); sh = _f32c(sh)
        dL_dout_color = _f32c(dL_dout_color); dL_dout_depth = _f32c(dL_dout_depth)
        M = int(sh.size(1)) if (sh is not None and sh.numel() != 0) else 0
        opts = dict(dtype=torch.float32, device=dev)
        # every row of these is written by the kernel, so no torch.zeros fills (rasterize_points.cu:154-163)
        dL_dmeans3D = torch.empty((P, 3), **opts)
        dL_dmeans2D = torch.empty((P, 3), **opts)
        dL_dcolors = torch.empty((P, NUM_CHANNELS), **opts)
        dL_dopacity = torch.empty((P, 1), **opts)
        dL_dcov3D = torch.empty((P, 6), **opts)
        has_sh = M != 0 and (colors is None or colors.numel() == 0)
        acc = SH_GRAD_ACCUMULATOR
        use_acc = has_sh and acc is not None and acc.shape == (P, M, 3) and acc.is_contiguous() and acc.device == dev
        overwrite = use_acc and SH_GRAD_OVERWRITE      # first view of a step: the kernel writes every row, nothing to zero
        dL_dsh = acc if use_acc else (torch.empty((P, M, 3), **opts) if has_sh else torch.zeros((P, M, 3), **opts))
        dL_dscales = torch.empty((P, 3), **opts)
        dL_drotations = torch.empty((P, 4), **opts)
        if P != 0:
            arena = torch.empty((P, 12), **opts)
            entry = L.b200gs_rast_backward_accumulate_sh if (use_acc and not overwrite) else L.b200gs_rast_backward
            check(entry(
                P, int(degree), M, int(R), W, H, ptr(background), ptr(means3D), ptr(sh), ptr(colors), ptr(scales),
                float(scale_modifier), ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix),
                ptr(campos), float(tan_fovx), float(tan_fovy), ptr(radii), geomBuffer.data_ptr(),
                binningBuffer.data_ptr(), imageBuffer.data_ptr(), ptr(dL_dout_color), ptr(dL_dout_depth),
                arena.data_ptr(), dL_dmeans2D.data_ptr(), dL_dcolors.data_ptr(), dL_dopacity.data_ptr(),
                dL_dmeans3D.data_ptr(), dL_dcov3D.data_ptr(), ptr(dL_dsh) if has_sh else None,
                dL_dscales.data_ptr(), dL_drotations.data_ptr(), current_stream()),
                "rasterize_gaussians_backward")
            if use_acc and AFTER_SH_ACCUMULATE is not None:
                AFTER_SH_ACCUMULATE()
            if debug:
                torch.cuda.synchronize(dev)
        return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, (None if use_acc else dL_dsh), dL_dscales, dL_drotations

    @staticmethod
    def mark_visible(means3D, viewmatrix, projmatrix):
        P = int(means3D.size(0))
        present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
        if P != 0:
            means3D = _f32c(means3D)
            check(_lib.lib().b200gs_mark_visible(P, ptr(means3D), ptr(_f32c(viewmatrix)), ptr(_f32c(projmatrix)),
                                                 present.data_ptr(), current_stream()), "mark_visible")
        return present

    @staticmethod
    def export_state(field, P, R, W, H, geomBuffer, binningBuffer, imgBuffer):
        """Parity/debug: internal state re-expressed in the reference's layout (see b200gs.h)."""
        L = _lib.lib()
        args = (field.encode(), P, R, W, H, geomBuffer.data_ptr(), binningBuffer.data_ptr(), imgBuffer.data_ptr())
        n = L.b200gs_rast_export(*args, None, 0, current_stream())
        if n < 0:
            check(-1, "rast_export")
        dst = torch.empty((max(int(n), 1),), dtype=torch.uint8, device=geomBuffer.device)
        if n > 0 and L.b200gs_rast_export(*args, dst.data_ptr(), n, current_stream()) < 0:
            check(-1, "rast_export")
        return dst[:n]


_C = _CModule


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings):
        args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                raster_settings.image_height, raster_settings.image_width, sh, raster_settings.sh_degree,
                raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        else:
            num_rendered, color, depth, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        num_rendered = ctx.num_rendered
        raster_settings = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer,
         imgBuffer) = ctx.saved_tensors
        args = (raster_settings.bg, means3D, radii, colors_precomp, scales, rotations,
                raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy, grad_out_color,
                grad_depth, sh, raster_settings.sh_degree, raster_settings.campos, geomBuffer, num_rendered,
                binningBuffer, imgBuffer, raster_settings.debug)
        if raster_settings.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
                 grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        else:
            (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh,
             grad_scales, grad_rotations) = _C.rasterize_gaussians_backward(*args)
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales,
                grad_rotations, grad_cov3Ds_precomp, None)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            raster_settings = self.raster_settings
            visible = _C.mark_visible(positions, raster_settings.viewmatrix, raster_settings.projmatrix)
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        if not torch.is_grad_enabled() and not raster_settings.debug:
            # inference: keep the opaque state buffers reachable for debugging / statistics (export_state)
            args = (raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations, raster_settings.scale_modifier,
                    cov3D_precomp, raster_settings.viewmatrix, raster_settings.projmatrix, raster_settings.tanfovx,
                    raster_settings.tanfovy, raster_settings.image_height, raster_settings.image_width, shs,
                    raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered, raster_settings.debug)
            R, color, depth, radii, geom, binning, img = _C.rasterize_gaussians(*args)
            self.last_state = (R, geom, binning, img)
            return color, radii, depth
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)
