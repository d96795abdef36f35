"""Seeded synthetic scenes and cameras shared by the tests and bench.py.

The camera matrices follow the reference's conventions (scene/cameras.py:63-68,
utils/graphics_utils.py:38-71): world_view_transform and full_proj_transform are handed to
the rasterizer TRANSPOSED (row-vector convention), campos = inverse(world_view)[3, :3].
"""
import math
from typing import NamedTuple

import numpy as np
import torch

SEED = 6666          # the reference's own seed (train_4DGS.py:416)


class SynthCamera(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor      # [4,4] float32, transposed world->view
    projmatrix: torch.Tensor      # [4,4] float32, transposed full projection
    campos: torch.Tensor          # [3]
    time: float
    frame_num: int


def _world2view(R, t):
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R.T
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return Rt.astype(np.float32)


def _projection(znear, zfar, fovx, fovy):
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_camera(W, H, R=None, t=None, time=0.0, frame_num=0, distance=4.5, device="cpu"):
    """Pin-hole camera `distance` in front of the cube centre looking at it; focal length
    582.69*(H/512) px as in scene/dataset_readers.py:994."""
    focal = 582.69 * (H / 512.0)
    fovx = 2 * math.atan(W / (2 * focal))
    fovy = 2 * math.atan(H / (2 * focal))
    if R is None:
        R = np.eye(3)
    if t is None:
        t = np.array([0.0, 0.0, distance])
    wv = torch.tensor(_world2view(np.asarray(R, dtype=np.float64), np.asarray(t, dtype=np.float64))).transpose(0, 1)
    proj = _projection(0.01, 100.0, fovx, fovy).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(proj.unsqueeze(0))).squeeze(0)
    campos = wv.inverse()[3, :3]
    return SynthCamera(H, W, math.tan(fovx * 0.5), math.tan(fovy * 0.5), wv.contiguous().to(device),
                       full.contiguous().to(device), campos.contiguous().to(device), float(time), int(frame_num))


def orbit_cameras(n, W, H, distance=4.5, device="cpu"):
    """n cameras on a circle around the cube, all looking at the origin (multi-view batches)."""
    cams = []
    for k in range(n):
        ang = 2 * math.pi * k / max(n, 1) * 0.25 - 0.3     # a quarter-orbit keeps every view in front
        c, s = math.cos(ang), math.sin(ang)
        R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])      # camera-to-world rotation
        t = np.array([0.0, 0.0, distance])
        cams.append(make_camera(W, H, R=R, t=t, time=k / max(n - 1, 1), frame_num=k, device=device))
    return cams


def video_trajectories(W, H, frames=60, distance=4.5, device="cpu"):
    """The render_4DGS.py video paths (scene/dataset_readers.py:1168-1190: up-down, side, zoom-in, circle from
    test_trajectory/*_{R,t}_list, plus vfx) restated for the synthetic scene: as in those files the rotation is the identity
    and only the camera translation moves -- linearly between +a and -a (up-down: y, side: x), from 0 to -b in z (zoom-in),
    on an ellipse in x / z (circle), or rising while moving in (vfx) -- with the amplitudes (0.08-0.09 and 0.24 scene units
    at a scene depth of ~1) scaled to this scene's camera distance.  Frame i of a path carries time = linspace(0, 2, frames)[i] / 2
    and frame_num = i (scene/dataset_readers.py:1004-1018, :1150-1159).  Returns {name: [SynthCamera] * frames}."""
    k = distance
    lin = np.linspace(1.0, -1.0, frames)
    ang = np.linspace(0.0, 2.0 * math.pi, frames)
    ramp = np.linspace(0.0, 1.0, frames)
    paths = {
        "up_down": [(0.0, 0.08 * k * a, 0.0) for a in lin],
        "side": [(0.09 * k * a, 0.0, 0.0) for a in lin],
        "zoom_in": [(0.0, 0.0, -0.24 * k * r) for r in ramp],
        "circle": [(-0.04 * k * math.cos(a), -0.003 * k * math.sin(a), 0.09 * k * math.cos(a)) for a in ang],
        "vfx": [(0.0, 0.16 * k * r, -0.24 * k * r) for r in ramp],
    }
    times = np.linspace(0.0, 2.0, frames, dtype=np.float32) / 2.0
    out = {}
    for name, offs in paths.items():
        out[name] = [make_camera(W, H, R=np.eye(3), t=np.array([o[0], o[1], distance + o[2]]), time=float(times[i]), frame_num=i,
                                 device=device) for i, o in enumerate(offs)]
    return out


def make_gaussians(P, scale_mu=0.004, sh_degree=3, seed=SEED, device="cpu"):
    """Raw (pre-activation) Gaussian parameters, generated on the CPU for reproducibility."""
    g = torch.Generator().manual_seed(seed)
    M = (sh_degree + 1) ** 2
    xyz = (torch.rand(P, 3, generator=g) * 3.0 - 1.5)
    log_scale = torch.randn(P, 3, generator=g) * 0.6 + math.log(scale_mu)
    rot = torch.randn(P, 4, generator=g)
    opacity_logit = torch.randn(P, 1, generator=g) * 1.5
    f_dc = torch.rand(P, 1, 3, generator=g) * 3.0 - 1.5
    f_rest = torch.randn(P, M - 1, 3, generator=g) * 0.05
    scene_flow = torch.randn(P, 3, generator=g) * 1e-3
    out = dict(xyz=xyz, log_scale=log_scale, rot=rot, opacity_logit=opacity_logit,
               shs=torch.cat([f_dc, f_rest], dim=1).contiguous(), scene_flow=scene_flow)
    return {k: v.to(device) for k, v in out.items()}


def activated(raw):
    """What gaussian_renderer/__init__.py:130-132 hands to the rasterizer."""
    return dict(means3D=raw["xyz"].contiguous(), scales=torch.exp(raw["log_scale"]).contiguous(),
                rotations=torch.nn.functional.normalize(raw["rot"]).contiguous(),
                opacities=torch.sigmoid(raw["opacity_logit"]).contiguous(), shs=raw["shs"].contiguous())
