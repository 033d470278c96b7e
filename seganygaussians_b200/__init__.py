"""seganygaussians_b200 -- B200-native (sm_100a) differentiable Gaussian feature rasterizer.

One hot path of Jumpat/SegAnyGAussians, rebuilt from scratch: preprocess -> tile binning (device radix
sort) -> per-tile alpha-composited forward of RGB / depth / K-dim features -> per-pixel backward.
The compute lives in ``lib/libsagars.so`` (hand-written CUDA behind the C ABI of ``include/sagars.h``);
this package is the host-side mirror of the reference's operator API:

* :mod:`seganygaussians_b200.rasterizer`  -- ``GaussianRasterizationSettings`` / ``GaussianRasterizer``
  for the three reference variants (BASE, CF, DEPTH);
* :mod:`seganygaussians_b200.dropin`      -- import-compatible packages with the reference's names
  (``diff_gaussian_rasterization``, ``diff_gaussian_rasterization_contrastive_f``,
  ``diff_gaussian_rasterization_depth``, ``gaussian_renderer``); call :func:`activate` or put the
  directory on ``PYTHONPATH`` so the reference's scripts run unmodified;
* :mod:`seganygaussians_b200.shims`       -- stand-ins for the third-party imports the reference's scripts need to start
  (``plyfile``, ``simple_knn._C.distCUDA2``, ``pytorch3d.ops.knn_points``), used only when the real ones are absent;
* :mod:`seganygaussians_b200.data_parallel` -- image-batch data parallelism (one camera per GPU).
"""
import os
import sys

from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    GaussianRasterizerContrastiveF,
    GaussianRasterizerDepth,
)

__version__ = "0.1.0"

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
SHIMS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def activate(shims: bool = True) -> str:
    """Make the reference-named packages importable.

    The rasterizer packages (``dropin/``) are PREPENDED to ``sys.path``: they replace the reference's extensions.
    The stand-ins for third-party imports of the reference's scripts (``shims/``: ``plyfile``, ``simple_knn``,
    ``pytorch3d.ops.knn_points`` -- SURVEY.md section 8(f) rank 1) are APPENDED, so a real installation of those
    packages always wins."""
    if DROPIN_DIR not in sys.path:
        sys.path.insert(0, DROPIN_DIR)
    if shims and SHIMS_DIR not in sys.path:
        sys.path.append(SHIMS_DIR)
    return DROPIN_DIR
