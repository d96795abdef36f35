"""CPU: the off-critical-path frame writer (b200gs.output.FrameWriter): PNGs named like the reference's
(render_4DGS.py:58: '{0:05d}.png'), lossless, in any arrival order; one MP4 in index order on close."""
import os

import numpy as np


def test_frame_writer_pngs_are_lossless_and_video_is_written(tmp_path):
    import cv2
    from b200gs.output import FrameWriter
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, size=(48, 64, 3), dtype=np.uint8) for _ in range(9)]
    with FrameWriter(png_dir=str(tmp_path / "png"), video_path=str(tmp_path / "v.mp4"), fps=30, workers=3) as w:
        for i in (3, 0, 8, 1, 2, 7, 4, 6, 5):                 # frames of a sharded render arrive out of order
            w.put(i, frames[i])
    for i, f in enumerate(frames):
        back = cv2.cvtColor(cv2.imread(str(tmp_path / "png" / f"{i:05d}.png")), cv2.COLOR_BGR2RGB)
        assert np.array_equal(back, f)
    assert os.path.getsize(tmp_path / "v.mp4") > 500
    cap = cv2.VideoCapture(str(tmp_path / "v.mp4"))
    n = 0
    while cap.read()[0]:
        n += 1
    assert n == len(frames)
