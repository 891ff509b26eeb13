"""Drop-in alias: `from diff_gaussian_rasterization import GaussianRasterizationSettings,
GaussianRasterizer` (EOGS++ gaussian_renderer/renderer.py:15-18) resolves to the sm_100a
implementation in eogs2_b200 when this repository root is on sys.path ahead of the
reference's compiled extension."""
from eogs2_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    rasterize_gaussians,
    _RasterizeGaussians,
)
