"""Import-name shim: ``from diff_surfel_rasterization_part import GaussianRasterizationSettings, GaussianRasterizer``
(renderer/gaussian_renderer_2d/__init__.py:3 in PartGS) resolves to partgs_b200."""
from partgs_b200.diff_surfel_rasterization_part import (  # noqa: F401
    GaussianRasterizationSettings, GaussianRasterizer, rasterize_gaussians, _RasterizeGaussians, _C)
