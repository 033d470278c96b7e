"""Drop-in for the reference package ``diff_gaussian_rasterization_depth`` (DEPTH variant: RGB + per-Gaussian
mask + view depth; ``submodules/diff-gaussian-rasterization-depth/.../__init__.py``), backed by libsagars."""
from seganygaussians_b200.rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizerDepth as GaussianRasterizer,
    _RasterizeGaussiansDepth as _RasterizeGaussians,
    _RasterizeMaskGaussians,
    cpu_deep_copy_tuple,
)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, mask, scales, rotations, cov3Ds_precomp,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, mask, scales, rotations,
                                     cov3Ds_precomp, raster_settings)


def rasterize_mask_gaussians(means3D, means2D, opacities, mask, scales, rotations, cov3Ds_precomp, raster_settings):
    return _RasterizeMaskGaussians.apply(means3D, means2D, opacities, mask, scales, rotations, cov3Ds_precomp,
                                         raster_settings)
