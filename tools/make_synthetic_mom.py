"""Synthetic stand-in for the stage-1 output that train_4DGS.py / render_4DGS.py read: writes
    <out>/MOM/train_data.pth   dict(camera_angle_x, camera_angle_y, W, H, pcd_points [3,N], pcd_colors [N,3],
                                    frames = [{transform_matrix: 4x4 camera-to-world (OpenGL axes), image: PIL}, ...])
                               (written by train_motion.py:463; read by scene/dataset_readers.py:1022-1057, :802-868, :1176-1187)
    <out>/MOM/scene_flow.pth   float tensor [3,N]                      (train_motion.py:464; scene/gaussian_model.py:183)
    <out>/MOM/video/00000.png ...  the animated frames of the centre view (frames[2]; scene/dataset_readers.py:802-843)
so that the reference's own scripts can be run, unchanged, on a box without the stage-1 checkpoints or any dataset.
The images are procedural (smooth colour fields + blobs that drift with the frame index): enough for the optimisation to have
something to fit and for densify / prune to fire; they are not meant to look like anything.

    python tools/make_synthetic_mom.py <out_dir> [--points 210000] [--width 320] [--height 192] [--views 5] [--video-frames 60]
"""
import argparse
import math
import os

import numpy as np
import torch
from PIL import Image


def _image(W, H, seed, shift=0.0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    x = x / W; y = y / H
    img = np.zeros((H, W, 3), np.float32)
    for c in range(3):
        a, b, ph = rng.uniform(1.0, 4.0), rng.uniform(1.0, 4.0), rng.uniform(0, 2 * math.pi)
        img[..., c] = 0.5 + 0.3 * np.sin(2 * math.pi * (a * x + b * y) + ph + shift)
    for _ in range(6):                      # a few blobs that move with `shift`
        cx, cy, r = rng.uniform(0.1, 0.9), rng.uniform(0.1, 0.9), rng.uniform(0.03, 0.12)
        col = rng.uniform(0, 1, 3)
        d = np.exp(-(((x - cx - 0.05 * shift) ** 2 + (y - cy) ** 2) / (2 * r * r)))
        img = img * (1 - d[..., None]) + col * d[..., None]
    return Image.fromarray((np.clip(img, 0, 1) * 255).astype(np.uint8), "RGB")


def _c2w_opengl(angle_y, offset, distance):
    """Camera `distance` in front of the origin, rotated by angle_y about the vertical axis and shifted by `offset`; returned as the
    OpenGL-convention camera-to-world matrix the reader expects (it flips the y / z camera axes and inverts)."""
    c, s = math.cos(angle_y), math.sin(angle_y)
    R_c2w = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])          # COLMAP axes: x right, y down, z forward
    centre = R_c2w @ np.array([offset[0], offset[1], -distance + offset[2]])
    m = np.eye(4)
    m[:3, :3] = R_c2w
    m[:3, 3] = centre
    m[:3, 1:3] *= -1                                                          # COLMAP -> OpenGL camera axes
    return m


def write(out_dir, points=210000, width=320, height=192, views=5, video_frames=60, seed=6666, distance=4.5):
    mom = os.path.join(out_dir, "MOM")
    os.makedirs(os.path.join(mom, "video"), exist_ok=True)
    rng = np.random.default_rng(seed)
    focal = 582.69 * (height / 512.0)
    pts = rng.uniform(-1.5, 1.5, size=(3, points)).astype(np.float32)
    cols = rng.uniform(0.05, 0.95, size=(points, 3)).astype(np.float32)
    frames = []
    for v in range(max(views, 3)):
        ang = (v - 2) * 0.06
        off = (0.05 * (v - 2), 0.02 * ((v % 2) * 2 - 1), 0.0)
        frames.append({"transform_matrix": _c2w_opengl(ang, off, distance).tolist(), "image": _image(width, height, seed + 10 + v)})
    data = {"camera_angle_x": 2 * math.atan(width / (2 * focal)), "camera_angle_y": 2 * math.atan(height / (2 * focal)),
            "W": width, "H": height, "pcd_points": pts, "pcd_colors": cols, "frames": frames}
    torch.save(data, os.path.join(mom, "train_data.pth"))
    torch.save(torch.from_numpy(rng.normal(0.0, 1e-3, size=(3, points)).astype(np.float32)), os.path.join(mom, "scene_flow.pth"))
    for i in range(video_frames):
        _image(width, height, seed + 12, shift=2.0 * i / max(video_frames - 1, 1)).save(os.path.join(mom, "video", f"{i:05d}.png"))
    return mom


def write_config(path, coarse_iterations=60, iterations=160, batch_size=2, time_res=50):
    """A short-run variant of arguments/dnerf/hellwarrior.py (same model; iteration counts and densify / prune cadence shrunk so
    that a few hundred iterations pass through coarse -> fine, one densification and one pruning event)."""
    with open(path, "w") as f:
        f.write(f'''ModelHiddenParams = dict(
    kplanes_config = {{'grid_dimensions': 2, 'input_coordinate_dim': 4, 'output_coordinate_dim': 32, 'resolution': [64, 64, 64, {time_res}]}},
    multires = [1, 2], defor_depth = 0, net_width = 64, plane_tv_weight = 0.0001, time_smoothness_weight = 0.01,
    l1_time_planes = 0.0001, weight_decay_iteration = 0, bounds = 1.6)
OptimizationParams = dict(
    coarse_iterations = {coarse_iterations}, iterations = {iterations}, batch_size = {batch_size},
    deformation_lr_init = 0.00016, deformation_lr_final = 0.0000016, deformation_lr_delay_mult = 0.01,
    grid_lr_init = 0.0016, grid_lr_final = 0.000016, percent_dense = 0.01, render_process = False,
    densify_from_iter = 20, densification_interval = 40, densify_until_iter = 15000,
    pruning_from_iter = 20, pruning_interval = 50, opacity_reset_interval = 3000)
''')
    return path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out_dir")
    ap.add_argument("--points", type=int, default=210000)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=192)
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--video-frames", type=int, default=60)
    a = ap.parse_args()
    print(write(a.out_dir, a.points, a.width, a.height, a.views, a.video_frames))
