"""Import-name shim: `from diff_gaussian_rasterization_df import GaussianRasterizationSettings,
GaussianRasterizer` (gaussian_renderer/__init__.py:15 of the reference) resolves to the B200
implementation in ex4dgs_b200.  Put the repository root on PYTHONPATH instead of installing the
reference's submodules/diff_gaussian_rasterization_df."""
from ex4dgs_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,  # noqa: F401
                                    _RasterizeGaussians, rasterize_gaussians, cpu_deep_copy_tuple)
