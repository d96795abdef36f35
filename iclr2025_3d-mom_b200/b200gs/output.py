"""Render output path: quantise on the GPU, copy 3 bytes per pixel to pinned host
memory on a side stream, hand finished frames to the writer off the critical path.

The reference does `to8b = lambda x: (255*np.clip(x.cpu().numpy(),0,1)).astype(np.uint8)` per frame
(render_4DGS.py:49, train_4DGS.py:335): a blocking 12 B/pixel D2H copy plus host-side clip/scale/cast inside the
render loop. Here the frame is clipped, scaled and cast by `b200gs_to8b_hwc` (same truncation), and a ring of
pinned buffers lets frame i's copy overlap frame i+1's rendering.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, current_stream

_lib.register("b200gs_to8b_hwc", ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p])


def to8b(image):
    """[3,H,W] float32 CUDA image -> [H,W,3] uint8 CUDA tensor, (255 * clip(x, 0, 1)).astype(uint8)."""
    if not (image.is_cuda and image.dtype == torch.float32 and image.dim() == 3 and image.shape[0] == 3):
        raise RuntimeError("to8b needs a [3,H,W] float32 CUDA image (there is no CPU path)")
    img = image.detach().contiguous()
    H, W = int(img.shape[1]), int(img.shape[2])
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=img.device)
    check(_lib.lib().b200gs_to8b_hwc(H, W, img.data_ptr(), out.data_ptr(), current_stream()), "to8b_hwc")
    return out


class FrameRing:
    """Pinned-host ring for finished frames. `push(image)` quantises on the current stream and starts an async
    D2H copy on a side stream; `pop()` returns the oldest frame as a numpy array once its copy has landed.  By default the
    array is a COPY (safe to append to a list, as render_4DGS.py:60-76 does before `imageio.mimwrite`); `pop(copy=False)`
    returns a view of the pinned slot, valid only until `depth` more frames have been pushed (for a writer that consumes
    the frame immediately)."""

    def __init__(self, H, W, depth=4, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.host = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.head = 0          # next slot to fill
        self.count = 0
        self.bytes_per_frame = H * W * 3

    def push(self, image):
        if self.count == len(self.host):
            raise RuntimeError("FrameRing full: pop() a frame first")
        q = to8b(image)
        ready = torch.cuda.Event()
        ready.record()
        slot = self.head
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[slot].copy_(q, non_blocking=True)
            q.record_stream(self.copy_stream)
            self.done[slot].record(self.copy_stream)
        self.head = (self.head + 1) % len(self.host)
        self.count += 1

    def pop(self, copy=True):
        if self.count == 0:
            raise RuntimeError("FrameRing empty")
        slot = (self.head - self.count) % len(self.host)
        self.done[slot].synchronize()
        self.count -= 1
        frame = self.host[slot].numpy()
        return frame.copy() if copy else frame


class FrameStore:
    """All frames of a camera path in ONE pinned host array (render_4DGS.py:41-76 knows the frame count up front: it renders the
    whole trajectory into a list and then calls `imageio.mimwrite`).  `put(i, image)` quantises on the current stream and starts
    the async 3 B/pixel copy of frame i straight into its final place, so there is no host-side copy at all (FrameRing.pop()
    costs one 6 MB memcpy per 1080p frame on the thread that also launches the kernels); `array()` waits for the copies and
    returns the [n,H,W,3] uint8 numpy view (`list(store.array())` is the reference's `render_images`).  The pinned allocation
    (cudaHostAlloc, ~0.2 s per GB) is made once in the constructor; a store can be refilled for the next path once the previous
    one has been consumed."""

    def __init__(self, n_frames, H, W, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.host = torch.empty((n_frames, H, W, 3), dtype=torch.uint8).pin_memory()
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bytes_per_frame = H * W * 3
        self.filled = 0

    def put(self, index, image):
        if not 0 <= index < self.host.shape[0]:
            raise IndexError(f"FrameStore.put: frame {index} outside [0, {self.host.shape[0]})")
        q = to8b(image)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            self.host[index].copy_(q, non_blocking=True)
            q.record_stream(self.copy_stream)
        self.filled += 1

    def array(self):
        self.copy_stream.synchronize()
        return self.host.numpy()


class FrameWriter:
    """PNG / MP4 encoding off the critical path.  render_4DGS.py:58-76 calls
    `torchvision.utils.save_image` per frame INSIDE its render loop and `imageio.mimwrite` after it, so its "FPS" is PNG-encode
    bound; here the render loop only hands finished uint8 [H,W,3] frames (what `FrameRing.pop()` returns) to a queue and a few
    worker threads encode them (OpenCV releases the GIL while encoding): PNG files named `{index:05d}.png` like the reference's,
    and, on `close()`, one MP4 of all frames in index order.  Host-only code: nothing here touches the GPU."""

    def __init__(self, png_dir=None, video_path=None, fps=30, workers=4, max_pending=64):
        import queue
        import threading
        self.png_dir, self.video_path, self.fps = png_dir, video_path, fps
        if png_dir is not None:
            import os
            os.makedirs(png_dir, exist_ok=True)
        self.q = queue.Queue(maxsize=max_pending)          # back-pressure: rendering stalls only if encoding falls this far behind
        self.frames = {} if video_path is not None else None
        self.errors = []
        self.threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, workers))]
        for t in self.threads:
            t.start()

    def _work(self):
        import os
        import cv2
        while True:
            item = self.q.get()
            if item is None:
                self.q.task_done()
                return
            idx, frame = item
            try:
                if self.png_dir is not None:
                    cv2.imwrite(os.path.join(self.png_dir, "{0:05d}.png".format(idx)), cv2.cvtColor(frame, cv2.COLOR_RGB2BGR))
                if self.frames is not None:
                    self.frames[idx] = frame
            except Exception as ex:                        # reported by close()
                self.errors.append((idx, ex))
            self.q.task_done()

    def put(self, index, frame):
        """frame: uint8 [H,W,3] numpy array that the caller will not modify (FrameRing.pop() hands out a copy)."""
        self.q.put((int(index), frame))

    def close(self):
        """Waits for the queue to drain, writes the video, re-raises the first encoding error."""
        for _ in self.threads:
            self.q.put(None)
        for t in self.threads:
            t.join()
        if self.errors:
            raise RuntimeError(f"FrameWriter: frame {self.errors[0][0]} failed: {self.errors[0][1]}")
        if self.frames:
            import cv2
            order = sorted(self.frames)
            h, w = self.frames[order[0]].shape[:2]
            vw = cv2.VideoWriter(self.video_path, cv2.VideoWriter_fourcc(*"mp4v"), self.fps, (w, h))
            for i in order:
                vw.write(cv2.cvtColor(self.frames[i], cv2.COLOR_RGB2BGR))
            vw.release()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
