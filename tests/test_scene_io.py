"""On-disk formats: property lists and value layout of the reference's two PLY kinds,
round trips, and the synthetic model directory -- checked both through this package's readers and by re-doing the
reference's own read sequence (gaussian_model_ff.py:603-640, gaussian_model.py:271-306) on the files."""
import os
import sys

import numpy as np
import torch

from seganygaussians_b200 import scene_io, synthetic

sys.path.append(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seganygaussians_b200", "shims"))


def test_feature_ply_layout_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    P, K = 257, 32
    xyz, f = rng.standard_normal((P, 3)).astype(np.float32), rng.standard_normal((P, K)).astype(np.float32)
    op, sc, rot = rng.standard_normal((P, 1)).astype(np.float32), rng.standard_normal((P, 3)).astype(np.float32), rng.standard_normal((P, 4)).astype(np.float32)
    path = str(tmp_path / "point_cloud" / "iteration_7" / "feature_point_cloud.ply")
    scene_io.save_feature_ply(path, xyz, f, op, sc, rot)
    from plyfile import PlyData
    el = PlyData.read(path).elements[0]
    assert [p.name for p in el.properties] == scene_io.feature_ply_attributes(K)
    assert np.all(np.asarray(el["nx"]) == 0)
    # the reference's read sequence
    f_names = sorted([p.name for p in el.properties if p.name.startswith("f_")], key=lambda x: int(x.split("_")[-1]))
    assert len(f_names) == K
    assert np.array_equal(np.stack([np.asarray(el[n]) for n in f_names], axis=1), f)
    back = scene_io.load_feature_ply(path)
    for k, v in dict(xyz=xyz, point_features=f, opacity=op, scaling=sc, rotation=rot).items():
        assert np.array_equal(back[k], v), k


def test_scene_ply_is_channel_major_like_the_reference(tmp_path):
    rng = np.random.default_rng(1)
    P, M = 100, 16
    shs = rng.standard_normal((P, M, 3)).astype(np.float32)
    xyz = rng.standard_normal((P, 3)).astype(np.float32)
    op, sc, rot = rng.standard_normal((P, 1)).astype(np.float32), rng.standard_normal((P, 3)).astype(np.float32), rng.standard_normal((P, 4)).astype(np.float32)
    path = str(tmp_path / "scene_point_cloud.ply")
    scene_io.save_scene_ply(path, xyz, shs[:, :1], shs[:, 1:], op, sc, rot)
    from plyfile import PlyData
    el = PlyData.read(path).elements[0]
    assert [p.name for p in el.properties] == scene_io.scene_ply_attributes(3 * (M - 1))
    # GaussianModel.save_ply writes features_rest.transpose(1, 2).flatten(1): f_rest_j = channel j // (M-1), coefficient j % (M-1)
    assert np.array_equal(np.asarray(el["f_rest_0"]), shs[:, 1, 0]) and np.array_equal(np.asarray(el[f"f_rest_{M - 1}"]), shs[:, 1, 1])
    back = scene_io.load_scene_ply(path, max_sh_degree=3)
    # GaussianModel.load_ply: features_extra.reshape(P, 3, M-1), later transposed to [P, M-1, 3]
    assert np.array_equal(back["features_rest"].transpose(0, 2, 1), shs[:, 1:])
    assert np.array_equal(back["features_dc"][:, :, 0], shs[:, 0, :])


def test_synthetic_model_directory(tmp_path):
    out = scene_io.write_synthetic_model(str(tmp_path / "model"), P=2000, K=32, iteration=123)
    assert set(out) == {"scene", "feature", "contrastive_feature"}
    assert all(os.path.exists(p) and "iteration_123" in p for p in out.values())
    g = synthetic.make_gaussians(2000, 32, 1600, seed=0, sh_coeffs=16)
    feat = scene_io.load_feature_ply(out["contrastive_feature"])
    assert np.array_equal(feat["xyz"], g.means3D.numpy()) and np.array_equal(feat["point_features"], g.colors.numpy())
    # raw parameters: the reference applies sigmoid / exp / normalize on access (gaussian_model_ff.py:96-118)
    assert np.allclose(torch.sigmoid(torch.from_numpy(feat["opacity"])).numpy(), g.opacities.numpy(), rtol=1e-5, atol=1e-6)
    assert np.allclose(np.exp(feat["scaling"]), g.scales.numpy(), rtol=1e-5)
    scn = scene_io.load_scene_ply(out["scene"], max_sh_degree=3)
    assert np.array_equal(scn["features_dc"][:, :, 0], g.shs.numpy()[:, 0, :])
