"""Drop-in for `simple_knn._C.distCUDA2` (KNN/spatial.cu:15-26; used at
scene/gaussian_model.py:164)."""
import torch

from . import _lib
from ._lib import check, current_stream


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("distCUDA2 expects a CUDA tensor (there is no CPU path)")
    P = int(points.size(0))
    pts = points.contiguous()
    if pts.dtype != torch.float32:
        raise RuntimeError("distCUDA2 expects float32 points")
    means = torch.empty((P,), dtype=torch.float32, device=points.device)
    if P == 0:
        return means
    L = _lib.lib()
    nbytes = L.b200gs_dist2_scratch_bytes(P)
    scratch = torch.empty((nbytes,), dtype=torch.uint8, device=points.device)
    check(L.b200gs_dist2(P, pts.data_ptr(), means.data_ptr(), scratch.data_ptr(), nbytes, current_stream()),
          "distCUDA2")
    return means
