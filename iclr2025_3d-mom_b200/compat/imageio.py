"""Minimal `imageio.mimwrite` on top of OpenCV (no ffmpeg binary in this image); used only for
the preview videos of train_4DGS.py:350 / render_4DGS.py:76."""
import numpy as np


def mimwrite(path, frames, fps=30, quality=8, **kw):
    import cv2
    frames = [np.asarray(f) for f in frames]
    if not frames:
        return
    h, w = frames[0].shape[:2]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    for f in frames:
        vw.write(cv2.cvtColor(f[..., :3].astype(np.uint8), cv2.COLOR_RGB2BGR))
    vw.release()


def imwrite(path, img, **kw):
    import cv2
    cv2.imwrite(path, cv2.cvtColor(np.asarray(img)[..., :3].astype(np.uint8), cv2.COLOR_RGB2BGR))
