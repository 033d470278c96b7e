"""CPU: the brute-force neighbour-search oracle is pinned against (1) golden outputs of the unmodified reference
``simple_knn._C.distCUDA2`` on seeded clouds and (2) scipy's cKDTree (an independent exact method)."""
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import knn_oracle
from tests import knn_cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "knn", "simple_knn_small.npz")


@pytest.mark.parametrize("name", list(knn_cases.clouds()))
def test_oracle_matches_reference_golden(name):
    if not os.path.exists(GOLDEN):
        pytest.skip("golden vectors of the reference distCUDA2 not generated yet")
    gold = np.load(GOLDEN)[name]
    ours = knn_oracle.dist_cuda2(knn_cases.clouds()[name])
    # host arithmetic has no FMA contraction: allow the last bits of each squared distance
    assert np.allclose(ours, gold, rtol=2e-6, atol=1e-12), float(np.abs(ours - gold).max())


@pytest.mark.parametrize("name", ["uniform_2000", "clustered_3000", "flat_1500"])
def test_oracle_matches_kdtree(name):
    pts = knn_cases.clouds()[name]
    idx, d2 = knn_oracle.knn_bruteforce(pts, None, K=8)
    dist, tidx = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=8)
    assert np.allclose(np.sqrt(d2.astype(np.float64)), dist, rtol=1e-5, atol=1e-7)
    assert np.array_equal(idx[:, 0], np.arange(len(pts)))          # a cloud against itself: the point comes first
    # neighbour sets agree wherever the k-th and (k+1)-th distances are not tied
    _, d9 = knn_oracle.knn_bruteforce(pts, None, K=9)
    clear = d9[:, 8] > d9[:, 7] * (1 + 1e-5)
    assert np.array_equal(np.sort(idx[clear], axis=1), np.sort(tidx[clear], axis=1))


def test_exclude_self_and_padding():
    pts = knn_cases.clouds()["tiny_5"]
    idx, d2 = knn_oracle.knn_bruteforce(pts, None, K=8, exclude_self=True)
    assert (idx[:, :4] >= 0).all() and (idx[:, 4:] == -1).all()
    assert not (idx[:, :4] == np.arange(5)[:, None]).any()
    dup = knn_cases.clouds()["duplicates_1000"]
    assert np.all(knn_oracle.knn_bruteforce(dup, None, K=1, exclude_self=True)[1] == 0)   # the twin is at distance 0
