"""Stand-in for the compiled module diff_gaussian_rasterization._C (RAST/ext.cpp:15-19)."""
from b200gs.rasterizer import _CModule

rasterize_gaussians = _CModule.rasterize_gaussians
rasterize_gaussians_backward = _CModule.rasterize_gaussians_backward
mark_visible = _CModule.mark_visible
