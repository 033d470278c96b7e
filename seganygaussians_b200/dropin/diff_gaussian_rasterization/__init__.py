"""Drop-in for the reference package ``diff_gaussian_rasterization`` (BASE variant, 3 channels;
``submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py``), backed by libsagars."""
from seganygaussians_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeBase as _RasterizeGaussians,
    cpu_deep_copy_tuple,
)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)
