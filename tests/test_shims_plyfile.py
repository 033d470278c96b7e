"""The plyfile stand-in: byte-level format, round trips of the property lists the
reference writes (3DGS PLY: gaussian_model.py:213-234; feature PLY: gaussian_model_ff.py:552-592), and reading back
exactly the way the reference does (gaussian_model_ff.py:603-640, dataset_readers.py:122-131)."""
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seganygaussians_b200", "shims"))
from plyfile import PlyData, PlyElement  # noqa: E402


def _feature_ply_array(P, K, seed=0):
    rng = np.random.default_rng(seed)
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_{i}" for i in range(K)] + ["opacity"] + \
            [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    arr = np.empty(P, dtype=[(n, "f4") for n in names])
    vals = rng.standard_normal((P, len(names))).astype(np.float32)
    arr[:] = list(map(tuple, vals))
    return names, vals, arr


def test_binary_layout_is_header_plus_raw_little_endian_rows(tmp_path):
    names, vals, arr = _feature_ply_array(5, 4)
    path = tmp_path / "a.ply"
    PlyData([PlyElement.describe(arr, "vertex")]).write(str(path))
    raw = path.read_bytes()
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex 5\n" +
              "".join(f"property float {n}\n" for n in names) + "end_header\n").encode("ascii")
    assert raw.startswith(header)
    assert raw[len(header):] == vals.astype("<f4").tobytes()


def test_feature_ply_round_trip_the_way_the_reference_reads_it(tmp_path):
    P, K = 1000, 32
    names, vals, arr = _feature_ply_array(P, K, seed=3)
    path = str(tmp_path / "point_cloud.ply")
    PlyData([PlyElement.describe(arr, "vertex")]).write(path)
    plydata = PlyData.read(path)
    el = plydata.elements[0]
    xyz = np.stack((np.asarray(el["x"]), np.asarray(el["y"]), np.asarray(el["z"])), axis=1)
    assert np.array_equal(xyz, vals[:, 0:3])
    f_names = sorted([p.name for p in el.properties if p.name.startswith("f_")], key=lambda x: int(x.split("_")[-1]))
    assert len(f_names) == K
    feats = np.stack([np.asarray(el[n]) for n in f_names], axis=1)
    assert np.array_equal(feats, vals[:, 6:6 + K])
    assert np.array_equal(np.asarray(plydata["vertex"]["opacity"]), vals[:, 6 + K])
    assert len(plydata["vertex"]) == P and "vertex" in plydata


def test_mixed_types_ascii_and_big_endian():
    arr = np.empty(3, dtype=[("x", "f4"), ("y", "f8"), ("red", "u1"), ("id", "i4")])
    arr["x"], arr["y"], arr["red"], arr["id"] = [0.5, -1.25, 3.0], [1e-3, 2.0, -7.5], [0, 128, 255], [-1, 0, 7]
    for kw in (dict(text=True), dict(byte_order=">"), dict(byte_order="<")):
        buf = io.BytesIO()
        PlyData([PlyElement.describe(arr, "vertex")], **kw).write(buf)
        buf.seek(0)
        back = PlyData.read(buf)["vertex"]
        for n in arr.dtype.names:
            assert np.array_equal(np.asarray(back[n]), arr[n]), (kw, n)
    head = PlyData([PlyElement.describe(arr, "vertex")]).header
    assert "property float x" in head and "property double y" in head and "property uchar red" in head and "property int id" in head


def test_list_properties_are_read(tmp_path):
    path = tmp_path / "mesh.ply"
    path.write_text("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                    "element face 1\nproperty list uchar int vertex_indices\nend_header\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    ply = PlyData.read(str(path))
    assert ply.comments == ["made by hand"]
    assert np.array_equal(np.asarray(ply["vertex"]["x"]), np.array([0, 1, 0], np.float32))
    assert list(ply["face"]["vertex_indices"][0]) == [0, 1, 2]


def test_errors():
    with pytest.raises(Exception):
        PlyData.read(io.BytesIO(b"not a ply\n"))
    with pytest.raises(Exception):
        PlyData.read(io.BytesIO(b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nend_header\n\x00\x00"))
    with pytest.raises(TypeError):
        PlyElement.describe(np.zeros((3, 3), np.float32), "vertex")
