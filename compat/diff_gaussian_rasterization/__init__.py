"""Import alias for the stock package name (commented import at gaussian_renderer/__init__.py:14)."""
from opengaussian_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                          rasterize_gaussians)
