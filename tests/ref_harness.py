"""TEST INFRASTRUCTURE: ctypes access to the reference's own CUDA code compiled into
oracle/_ref/*.so by oracle/build_ref.sh (GPU box only). Plays the part of the reference's
torch glue (RAST/rasterize_points.cu) so the parity tests can call the unmodified
CudaRasterizer::Rasterizer::{forward,backward} and SimpleKNN::knn on the same tensors."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
_rast = None
_knn = None


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libref_rast.so"))


def rast():
    global _rast
    if _rast is None:
        L = ctypes.CDLL(os.path.join(REF_DIR, "libref_rast.so"))
        P = c_void_p
        L.ref_last_error.restype = c_char_p
        L.ref_rast_forward.restype = c_int
        L.ref_rast_forward.argtypes = [c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, P, c_float, P, P, P, P, P,
                                       c_float, c_float, c_int, P, P, P, c_int]
        L.ref_rast_backward.restype = c_int
        L.ref_rast_backward.argtypes = [c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, P, P, c_float, P, P, P, P,
                                        P, c_float, c_float, P, P, P] + [P] * 10 + [c_int]
        L.ref_mark_visible.restype = c_int
        L.ref_mark_visible.argtypes = [c_int, P, P, P, P]
        L.ref_rast_get.restype = c_longlong
        L.ref_rast_get.argtypes = [c_char_p, P, c_longlong]
        _rast = L
    return _rast


def knn():
    global _knn
    if _knn is None:
        L = ctypes.CDLL(os.path.join(REF_DIR, "libref_knn.so"))
        L.ref_dist2.restype = c_int
        L.ref_dist2.argtypes = [c_int, c_void_p, c_void_p]
        _knn = L
    return _knn


def _p(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


_FIELD_DTYPES = dict(depths=np.float32, clamped=np.uint8, means2D=np.float32, cov3D=np.float32,
                     conic_opacity=np.float32, rgb=np.float32, tiles_touched=np.uint32, point_offsets=np.uint32,
                     point_list=np.uint32, point_list_unsorted=np.uint32, keys=np.uint64, keys_unsorted=np.uint64,
                     accum_alpha=np.float32, n_contrib=np.uint32, ranges=np.uint32)


def ref_get(name):
    L = rast()
    n = L.ref_rast_get(name.encode(), None, 0)
    assert n >= 0, L.ref_last_error()
    buf = np.empty(max(n, 1), dtype=np.uint8)
    if n:
        assert L.ref_rast_get(name.encode(), buf.ctypes.data, n) == n, L.ref_last_error()
    return buf[:n].view(_FIELD_DTYPES[name])


def ref_forward(cam, bg, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, sh_degree=3, scale_modifier=1.0):
    """Reference forward (legacy default stream; synchronises). Returns R, color, depth, radii."""
    L = rast()
    P = means3D.shape[0]
    H, W = cam.image_height, cam.image_width
    dev = means3D.device
    color = torch.zeros(3, H, W, device=dev)
    depth = torch.zeros(1, H, W, device=dev)
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    M = shs.shape[1] if shs is not None and shs.numel() else 0
    torch.cuda.synchronize()
    R = L.ref_rast_forward(P, sh_degree, M, _p(bg), W, H, _p(means3D), _p(shs), _p(colors_precomp), _p(opacities),
                           _p(scales), scale_modifier, _p(rotations), _p(cov3D_precomp), _p(cam.viewmatrix),
                           _p(cam.projmatrix), _p(cam.campos), cam.tanfovx, cam.tanfovy, 0, _p(color), _p(depth),
                           _p(radii), 0)
    assert R >= 0, L.ref_last_error()
    torch.cuda.synchronize()
    return R, color, depth, radii


def ref_backward(cam, bg, R, radii, dL_dcolor, dL_ddepth, means3D, shs=None, colors_precomp=None, scales=None,
                 rotations=None, cov3D_precomp=None, sh_degree=3, scale_modifier=1.0):
    """Reference backward on the state left by the last ref_forward. Returns the dict of the
    ten gradient tensors RasterizeGaussiansBackwardCUDA allocates (rasterize_points.cu:154-163)."""
    L = rast()
    P = means3D.shape[0]
    H, W = cam.image_height, cam.image_width
    dev = means3D.device
    M = shs.shape[1] if shs is not None and shs.numel() else 0
    z = lambda *s: torch.zeros(*s, device=dev)
    g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacity=z(P, 1), colors=z(P, 3), depths=z(P, 1), means3D=z(P, 3),
             cov3D=z(P, 6), sh=z(P, M, 3), scales=z(P, 3), rotations=z(P, 4))
    torch.cuda.synchronize()
    rc = L.ref_rast_backward(P, sh_degree, M, R, _p(bg), W, H, _p(means3D), _p(shs), _p(colors_precomp), _p(scales),
                             scale_modifier, _p(rotations), _p(cov3D_precomp), _p(cam.viewmatrix), _p(cam.projmatrix),
                             _p(cam.campos), cam.tanfovx, cam.tanfovy, _p(radii), _p(dL_dcolor.contiguous()),
                             _p(dL_ddepth.contiguous()), _p(g["means2D"]), _p(g["conic"]), _p(g["opacity"]),
                             _p(g["colors"]), _p(g["depths"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["sh"]),
                             _p(g["scales"]), _p(g["rotations"]), 0)
    assert rc == 0, L.ref_last_error()
    torch.cuda.synchronize()
    return g


def ref_mark_visible(cam, means3D):
    """CudaRasterizer::Rasterizer::markVisible (rasterizer_impl.cu:141-153) -> bool[P]."""
    L = rast()
    P = means3D.shape[0]
    out = torch.zeros(P, dtype=torch.bool, device=means3D.device)
    torch.cuda.synchronize()
    assert L.ref_mark_visible(P, means3D.contiguous().data_ptr(), cam.viewmatrix.data_ptr(), cam.projmatrix.data_ptr(),
                              out.data_ptr()) == 0, L.ref_last_error()
    torch.cuda.synchronize()
    return out


def ref_dist2(points):
    out = torch.zeros(points.shape[0], device=points.device)
    torch.cuda.synchronize()
    assert knn().ref_dist2(points.shape[0], points.contiguous().data_ptr(), out.data_ptr()) == 0
    return out
