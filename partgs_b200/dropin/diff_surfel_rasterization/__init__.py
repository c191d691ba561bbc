"""Import-name shim: ``from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer``
(renderer/gaussian_renderer/__init__.py:5 in PartGS) resolves to partgs_b200."""
from partgs_b200.diff_surfel_rasterization import (  # noqa: F401
    GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _RasterizeGaussians, _C)
