"""On-disk formats of the reference's trained scenes and a synthetic-scene writer.

Two PLY layouts, both one ``vertex`` element of float32 properties (binary little endian):

* 3DGS scene  (``scene/gaussian_model.py:196-234``):  x y z  nx ny nz  f_dc_0..2  f_rest_0..(3*(M-1)-1)  opacity
  scale_0..2  rot_0..3   -- ``f_dc`` / ``f_rest`` are the SH tensors ``[P, M, 3]`` transposed to channel-major;
* feature scene (``scene/gaussian_model_ff.py:552-592``):  x y z  nx ny nz  f_0..f_{K-1}  opacity  scale_0..2  rot_0..3.

Stored values are the RAW parameters: opacity before the sigmoid, scales before the exp, rotations un-normalised --
exactly what ``load_ply`` of the reference puts back into ``_opacity`` / ``_scaling`` / ``_rotation``
(``gaussian_model_ff.py:603-640``).  ``write_synthetic_model`` lays out a model directory the reference's ``Scene`` can
load (``scene/__init__.py:150-205``: ``point_cloud/iteration_<n>/{scene,feature,contrastive_feature}_point_cloud.ply``),
filled with the SYN(P, ...) cloud of ``seganygaussians_b200.synthetic`` so BASELINE's config-3-like runs need no dataset.
Host-side file I/O only (numpy + ``plyfile``; the stand-in in ``shims/`` is used when the real package is absent).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np


def _plyfile():
    try:
        import plyfile
    except ImportError:
        import sys
        sys.path.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims"))
        import plyfile
    return plyfile


def feature_ply_attributes(K: int):
    """Property list of ``FeatureGaussianModel.construct_list_of_attributes`` (gaussian_model_ff.py:552-564)."""
    return ["x", "y", "z", "nx", "ny", "nz"] + [f"f_{i}" for i in range(K)] + ["opacity"] + \
           [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def scene_ply_attributes(n_rest: int):
    """Property list of ``GaussianModel.construct_list_of_attributes`` (gaussian_model.py:196-209)."""
    return ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(n_rest)] + \
           ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def _write(path: str, names, columns: np.ndarray) -> None:
    ply = _plyfile()
    assert columns.shape[1] == len(names)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    el = np.empty(columns.shape[0], dtype=[(n, "f4") for n in names])
    col = np.ascontiguousarray(columns, dtype=np.float32)
    for i, n in enumerate(names):
        el[n] = col[:, i]
    ply.PlyData([ply.PlyElement.describe(el, "vertex")]).write(path)


def save_feature_ply(path: str, xyz, point_features, opacity_raw, scaling_raw, rotation_raw) -> None:
    """Feature PLY as written by ``FeatureGaussianModel.save_ply`` (normals are zeros there too)."""
    xyz = np.asarray(xyz, np.float32)
    f = np.asarray(point_features, np.float32)
    cols = np.concatenate([xyz, np.zeros_like(xyz), f, np.asarray(opacity_raw, np.float32).reshape(-1, 1),
                           np.asarray(scaling_raw, np.float32), np.asarray(rotation_raw, np.float32)], axis=1)
    _write(path, feature_ply_attributes(f.shape[1]), cols)


def save_scene_ply(path: str, xyz, features_dc, features_rest, opacity_raw, scaling_raw, rotation_raw) -> None:
    """3DGS PLY as written by ``GaussianModel.save_ply``: ``features_dc [P,1,3]``, ``features_rest [P,M-1,3]`` (the
    reference flattens their transposes, i.e. channel-major)."""
    xyz = np.asarray(xyz, np.float32)
    dc = np.asarray(features_dc, np.float32).transpose(0, 2, 1).reshape(xyz.shape[0], -1)
    rest = np.asarray(features_rest, np.float32).transpose(0, 2, 1).reshape(xyz.shape[0], -1)
    cols = np.concatenate([xyz, np.zeros_like(xyz), dc, rest, np.asarray(opacity_raw, np.float32).reshape(-1, 1),
                           np.asarray(scaling_raw, np.float32), np.asarray(rotation_raw, np.float32)], axis=1)
    _write(path, scene_ply_attributes(rest.shape[1]), cols)


def _sorted_cols(el, prefix: str, exclude=()):
    names = [p.name for p in el.properties if p.name.startswith(prefix) and p.name not in exclude]
    names = sorted(names, key=lambda x: int(x.split("_")[-1]))
    return np.stack([np.asarray(el[n]) for n in names], axis=1) if names else np.zeros((len(el), 0), np.float32)


def load_feature_ply(path: str) -> Dict[str, np.ndarray]:
    """What ``FeatureGaussianModel.load_ply`` reads (gaussian_model_ff.py:603-640), as numpy arrays."""
    el = _plyfile().PlyData.read(path).elements[0]
    return dict(xyz=np.stack([np.asarray(el[a]) for a in "xyz"], axis=1), point_features=_sorted_cols(el, "f_"),
                opacity=np.asarray(el["opacity"])[:, None], scaling=_sorted_cols(el, "scale_"), rotation=_sorted_cols(el, "rot"))


def load_scene_ply(path: str, max_sh_degree: Optional[int] = None) -> Dict[str, np.ndarray]:
    """What ``GaussianModel.load_ply`` reads (gaussian_model.py:271-320): ``features_dc [P,3,1]``, ``features_rest
    [P,3,M-1]`` in the reference's channel-major file order."""
    el = _plyfile().PlyData.read(path).elements[0]
    P = len(el)
    dc = np.stack([np.asarray(el[f"f_dc_{i}"]) for i in range(3)], axis=1)[:, :, None]
    rest = _sorted_cols(el, "f_rest_")
    if max_sh_degree is not None:
        assert rest.shape[1] == 3 * (max_sh_degree + 1) ** 2 - 3
    return dict(xyz=np.stack([np.asarray(el[a]) for a in "xyz"], axis=1), features_dc=dc,
                features_rest=rest.reshape(P, 3, -1), opacity=np.asarray(el["opacity"])[:, None],
                scaling=_sorted_cols(el, "scale_"), rotation=_sorted_cols(el, "rot"))


def write_synthetic_model(model_path: str, P: int, K: int = 32, W: int = 1600, iteration: int = 30000, sh_degree: int = 3,
                          seed: int = 0) -> Dict[str, str]:
    """A model directory with the SYN(P, ...) cloud in the reference's layout: ``scene_point_cloud.ply`` (3DGS, SH degree
    ``sh_degree``), ``feature_point_cloud.ply`` and ``contrastive_feature_point_cloud.ply`` (K features) under
    ``point_cloud/iteration_<iteration>/``.  Returns the paths."""
    import torch
    from . import synthetic
    M = (sh_degree + 1) ** 2
    g = synthetic.make_gaussians(P, K, W, seed=seed, sh_coeffs=M)
    opac_raw = torch.logit(g.opacities.clamp(1e-6, 1 - 1e-6)).numpy()        # inverse_sigmoid, utils/general_utils.py
    scale_raw = torch.log(g.scales).numpy()
    shs = g.shs.numpy()                                                       # [P, M, 3]
    d = os.path.join(model_path, "point_cloud", f"iteration_{iteration}")
    out = {k: os.path.join(d, f"{k}_point_cloud.ply") for k in ("scene", "feature", "contrastive_feature")}
    save_scene_ply(out["scene"], g.means3D.numpy(), shs[:, :1, :], shs[:, 1:, :], opac_raw, scale_raw, g.rotations.numpy())
    for k in ("feature", "contrastive_feature"):
        save_feature_ply(out[k], g.means3D.numpy(), g.colors.numpy(), opac_raw, scale_raw, g.rotations.numpy())
    return out
