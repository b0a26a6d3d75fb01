"""Import alias: OpenGaussian does ``from ashawkey_diff_gaussian_rasterization import
GaussianRasterizationSettings, GaussianRasterizer`` (gaussian_renderer/__init__.py:15,
utils/sam_refinement_utils.py:21).  Put ``<repo>/compat`` and ``<repo>`` on PYTHONPATH and the
reference runs on the B200 rasterizer unchanged."""
from opengaussian_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                          rasterize_gaussians)
