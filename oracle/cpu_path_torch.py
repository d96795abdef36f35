"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Plain-PyTorch restatement of the pieces of the reference's PyTorch (CPU-runnable) path that BASELINE.json's config 1 times --
"HexPlane deformation + SH eval + projection forward on CPU" -- next to oracle/field_torch.py (the deformation field):
  * eval_sh:     utils/sh_utils.py:57-112 (called from gaussian_renderer/__init__.py:139-144 with `clamp_min(sh2rgb + 0.5, 0)`)
  * covariance:  scene/gaussian_model.py:31-35 + utils/general_utils.py:71-116 (strip_lowerdiag, build_rotation,
                 build_scaling_rotation; those three hard-code device="cuda" in the reference, here they follow the input)
  * project:     utils/graphics_utils.py:22-29 (geom_transform_points with the camera's full_proj_transform)
  * projection_matrix / world_view: utils/graphics_utils.py:38-71, scene/cameras.py:63-68
  * c1_forward:  the whole config-1 forward for one timestep.
Parity pinned: tests/test_oracle_cpu_path.py checks every function bit-for-bit (same operators, same order) against
tests/golden/cpu_path.pt, produced by oracle/gen_golden_cpu_path.py from the REAL reference functions in this container.
"""
import math

import torch

from . import field_torch

# real SH basis constants, bands 0..3 (the published 3DGS / svox2 values the reference hard-codes in utils/sh_utils.py:24-53)
_K0 = 0.28209479177387814
_K1 = 0.4886025119029199
_K2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_K3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def eval_sh(deg, sh, dirs):
    """sh [..., C, (deg+1)^2 or more], dirs [..., 3] unit vectors -> [..., C]; same term order as the reference."""
    if not 0 <= deg <= 3:
        raise ValueError("SH degree 0..3")
    c = lambda i: sh[..., i]
    out = _K0 * c(0)
    if deg == 0:
        return out
    x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
    out = (out - _K1 * y * c(1) + _K1 * z * c(2) - _K1 * x * c(3))
    if deg == 1:
        return out
    xx, yy, zz = x * x, y * y, z * z
    xy, yz, xz = x * y, y * z, x * z
    # every term is the constant times its factors, multiplied left to right, then the coefficient (the reference's operator order)
    def term(k, factors, coeff):
        t = k
        for f in factors:
            t = t * f
        return t * coeff
    band2 = ((xy,), (yz,), (2.0 * zz - xx - yy,), (xz,), (xx - yy,))
    for k, factors in enumerate(band2):
        out = out + term(_K2[k], factors, c(4 + k))
    if deg == 2:
        return out
    band3 = ((y, 3 * xx - yy), (xy, z), (y, 4 * zz - xx - yy), (z, 2 * zz - 3 * xx - 3 * yy), (x, 4 * zz - xx - yy),
             (z, xx - yy), (x, xx - 3 * yy))
    for k, factors in enumerate(band3):
        out = out + term(_K3[k], factors, c(9 + k))
    return out


def rotation_matrix(q):
    """[P,4] (w, x, y, z), normalised inside like the reference's build_rotation -> [P,3,3]."""
    n = torch.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    u = q / n[:, None]
    w, x, y, z = u[:, 0], u[:, 1], u[:, 2], u[:, 3]
    R = torch.zeros((q.shape[0], 3, 3), dtype=q.dtype, device=q.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - w * z); R[:, 0, 2] = 2 * (x * z + w * y)
    R[:, 1, 0] = 2 * (x * y + w * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - w * x)
    R[:, 2, 0] = 2 * (x * z - w * y); R[:, 2, 1] = 2 * (y * z + w * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def covariance(scaling, scaling_modifier, rotation):
    """Upper triangle [P,6] of (R S)(R S)^T."""
    S = torch.zeros((scaling.shape[0], 3, 3), dtype=torch.float, device=scaling.device)
    s = scaling_modifier * scaling
    S[:, 0, 0] = s[:, 0]; S[:, 1, 1] = s[:, 1]; S[:, 2, 2] = s[:, 2]
    L = rotation_matrix(rotation) @ S
    full = L @ L.transpose(1, 2)
    out = torch.zeros((full.shape[0], 6), dtype=torch.float, device=scaling.device)
    for k, (i, j) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2))):
        out[:, k] = full[:, i, j]
    return out


def project(points, full_proj_transform):
    """Homogeneous transform by the (transposed, row-vector convention) full projection; perspective divide with +1e-7."""
    ones = torch.ones(points.shape[0], 1, dtype=points.dtype, device=points.device)
    hom = torch.cat([points, ones], dim=1)
    o = torch.matmul(hom, full_proj_transform.unsqueeze(0))
    return (o[..., :3] / (o[..., 3:] + 0.0000001)).squeeze(dim=0)


def projection_matrix(znear, zfar, fovX, fovY):
    ty, tx = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = ty * znear, tx * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (right - (-right))
    Pm[1, 1] = 2.0 * znear / (top - (-top))
    Pm[0, 2] = (right + (-right)) / (right - (-right))
    Pm[1, 2] = (top + (-top)) / (top - (-top))
    Pm[3, 2] = 1.0
    Pm[2, 2] = 1.0 * zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def c1_forward(sd, levels, xyz, log_scale, rot, opacity_logit, shs, scene_flow, time, frame_num, view, full_proj, campos, sh_degree=3):
    """BASELINE.json config 1, one timestep: deform_network.forward (gaussian_renderer/__init__.py:101-103) -> exp / normalize /
    sigmoid (:130-132) -> eval_sh + 0.5, clamp (:139-144) -> cov3D (scene/gaussian_model.py:146-147) -> projection."""
    P = xyz.shape[0]
    tt = torch.full((P, 1), float(time), dtype=xyz.dtype, device=xyz.device)
    pts, sc, rt, op, sh = field_torch.deform_forward(sd, levels, xyz, log_scale, rot, opacity_logit, shs, tt, scene_flow, frame_num, 1)
    scales, rots, opac = torch.exp(sc), torch.nn.functional.normalize(rt), torch.sigmoid(op)
    shs_view = sh.transpose(1, 2).view(-1, 3, 16)
    d = pts - campos.repeat(P, 1)
    d = d / d.norm(dim=1, keepdim=True)
    rgb = torch.clamp_min(eval_sh(sh_degree, shs_view, d) + 0.5, 0.0)
    cov = covariance(scales, 1.0, rots)
    ndc = project(pts, full_proj)
    return {"means3D": pts, "scales": scales, "rotations": rots, "opacity": opac, "rgb": rgb, "cov3D": cov, "ndc": ndc}
