"""The exact-KNN kernels behind the `simple_knn._C.distCUDA2` / `pytorch3d.ops.knn_points` stand-ins (seganygaussians_b200/csrc/
knn_kernels.cuh: bounding box, Morton codes, the library's radix sort, two-level boxes, pruned search) executed on the CPU under
the CUDA execution shim, against the brute-force oracle (oracle/knn_oracle.py, itself pinned to golden outputs of the reference's
simple_knn and to scipy's cKDTree in tests/test_knn_oracle.py)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import knn_oracle
from tests import knn_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    d = tempfile.mkdtemp(prefix="sagars_emu_")
    so = os.path.join(d, "libemu_knn.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "seganygaussians_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda_emu", "emu_knn.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_knn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def _run(emu, pts, K, exclude_self, queries=None):
    n = len(pts)
    nq = n if queries is None else len(queries)
    idx = np.full((nq, K), -7, np.int64)
    d2 = np.zeros((nq, K), np.float32)
    mean = np.zeros(nq, np.float32)
    p = lambda a: None if a is None else a.ctypes.data
    pts = np.ascontiguousarray(pts, np.float32)
    q = None if queries is None else np.ascontiguousarray(queries, np.float32)
    emu.emu_knn(n, p(pts), nq, p(q), K, int(exclude_self), p(idx), p(d2), p(mean))
    return idx, d2, mean


@pytest.mark.parametrize("name", ["uniform_2000", "clustered_3000", "duplicates_1000", "flat_1500", "tiny_5"])
def test_dist_cuda2_stand_in(emu, name):
    """K = 3, self excluded, mean of the squared distances: what simple_knn._C.distCUDA2 returns."""
    pts = knn_cases.clouds()[name]
    _, d2, mean = _run(emu, pts, 3, True)
    want = knn_oracle.dist_cuda2(pts)
    assert np.allclose(mean, want, rtol=2e-6, atol=1e-12), float(np.abs(mean - want).max())


@pytest.mark.parametrize("name,K", [("uniform_2000", 16), ("clustered_3000", 16), ("flat_1500", 8), ("tiny_5", 8)])
def test_knn_points_stand_in(emu, name, K):
    """self included, K up to 16: the smoothing map's neighbour indices and squared distances, ascending."""
    pts = knn_cases.clouds()[name]
    idx, d2, _ = _run(emu, pts, K, False)
    widx, wd2 = knn_oracle.knn_bruteforce(pts, None, K=K)
    valid = widx >= 0
    assert np.array_equal(idx >= 0, valid)
    assert np.allclose(d2[valid], wd2[valid], rtol=2e-6, atol=1e-12)
    for row, ok in zip(d2, valid):
        assert np.all(np.diff(row[ok]) >= 0)
    # identical neighbour sets wherever the K-th distance is not tied with the next one
    if len(pts) > K:
        _, d_next = knn_oracle.knn_bruteforce(pts, None, K=K + 1)
        clear = d_next[:, K] > d_next[:, K - 1] * (1 + 1e-5)
        assert np.array_equal(np.sort(idx[clear], axis=1), np.sort(widx[clear], axis=1))


def test_separate_queries(emu):
    rng = np.random.default_rng(3)
    pts = knn_cases.clouds()["clustered_3000"]
    q = (rng.standard_normal((400, 3)) * 3).astype(np.float32)
    idx, d2, _ = _run(emu, pts, 4, False, queries=q)
    widx, wd2 = knn_oracle.knn_bruteforce(pts, q, K=4)
    assert np.allclose(d2, wd2, rtol=2e-6, atol=1e-12)
