"""CPU restatement of the neighbour searches behind ``sagars_knn`` -- TEST INFRASTRUCTURE ONLY (imported by tests/ and
tools/, never by the product).

* ``dist_cuda2`` follows the reference's ``distCUDA2`` (submodules/simple-knn/simple_knn.cu:146-219): for every point the
  three smallest squared distances to OTHER points (exclusion by index, so exact duplicates count with distance 0),
  result ``(best[0] + best[1] + best[2]) / 3.0f`` with ``best`` ascending, everything in float32.
* ``knn_bruteforce`` is the semantics of ``pytorch3d.ops.knn_points`` for 3-D points (squared L2, ascending, the point
  itself included when a cloud is searched against itself) -- pytorch3d is not vendored in the reference; its call
  sites are scene/gaussian_model_ff.py:326-331, 345-350.

Brute force over the full distance matrix in blocks: exact, O(N^2), meant for N up to a few 10^4.
Pinned by tests/test_knn_oracle.py against golden outputs of the unmodified reference extension
(tests/golden/knn/simple_knn_*.npz, made on a B200 by tests/golden/make_golden_knn.py) and against scipy's cKDTree."""
import numpy as np


def _sqdist_block(q: np.ndarray, p: np.ndarray) -> np.ndarray:
    """float32 squared distances, same expression shape as the kernels: dx*dx + dy*dy + dz*dz (no FMA on the host, so
    the last bit may differ from the GPU's contracted form)."""
    d = p[None, :, :] - q[:, None, :]
    d = d.astype(np.float32)
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def knn_bruteforce(points: np.ndarray, queries=None, K: int = 1, exclude_self: bool = False, block: int = 1024):
    """Returns (idx [Q,K] int64, dist2 [Q,K] float32); rows with fewer than K eligible points are padded with -1 / FLT_MAX."""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    same = queries is None
    qs = pts if same else np.ascontiguousarray(queries, dtype=np.float32)
    if exclude_self and not same:
        raise ValueError("exclude_self needs the cloud to be searched against itself")
    Q, N = qs.shape[0], pts.shape[0]
    idx = np.full((Q, K), -1, dtype=np.int64)
    d2 = np.full((Q, K), np.finfo(np.float32).max, dtype=np.float32)
    for s in range(0, Q, block):
        e = min(Q, s + block)
        D = _sqdist_block(qs[s:e], pts)
        if exclude_self:
            D[np.arange(e - s), np.arange(s, e)] = np.inf
        k = min(K, N - (1 if exclude_self else 0))
        if k <= 0:
            continue
        order = np.argsort(D, axis=1, kind="stable")[:, :k]
        idx[s:e, :k] = order
        d2[s:e, :k] = np.take_along_axis(D, order, axis=1)
    return idx, d2


def dist_cuda2(points: np.ndarray) -> np.ndarray:
    """simple_knn.cu:183: (best[0] + best[1] + best[2]) / 3.0f, float32."""
    _, d2 = knn_bruteforce(points, None, K=3, exclude_self=True)
    return ((d2[:, 0] + d2[:, 1]) + d2[:, 2]) / np.float32(3.0)
