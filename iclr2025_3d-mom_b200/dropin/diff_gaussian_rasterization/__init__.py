"""`diff_gaussian_rasterization` as the reference imports it
(gaussian_renderer/__init__.py:14), served by the B200-native engine."""
import os
import sys

_pkg_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

from b200gs.rasterizer import (  # noqa: E402,F401
    GaussianRasterizationSettings, GaussianRasterizer, _RasterizeGaussians, rasterize_gaussians,
    cpu_deep_copy_tuple, _C)
