"""Stand-in for the part of ``pytorch3d`` the reference uses: ``pytorch3d.ops.knn_points``
(scene/gaussian_model_ff.py:13, 326-331, 345-350, 378-384)."""
__version__ = "0.0.0+sagars.shim"
