"""eogs2_b200 — B200-native (sm_100a) drop-in for EOGS++'s differentiable affine-camera
Gaussian-splatting rasterizer (gardiens/EOGS2, submodules/diff-gaussian-rasterization).

Public surface = the reference's (DGR/diff_gaussian_rasterization/__init__.py):
    GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians
The shim package `diff_gaussian_rasterization/` at the repo root re-exports it under the
reference's import name, so `gaussian_renderer/renderer.py:15-18` runs unchanged.
"""
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    _RasterizeGaussians,
    rasterize_forward_raw,
    rasterize_backward_raw,
    assemble_grad_viewmatrix,
    export_state,
    ForwardState,
)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
