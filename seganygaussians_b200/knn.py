"""Exact K-nearest-neighbour search on the GPU (``sagars_knn`` in ``include/sagars.h``; kernels in ``csrc/knn.cu``).

Host-side wrapper used by the stand-ins for ``simple_knn._C.distCUDA2`` and ``pytorch3d.ops.knn_points``
(``seganygaussians_b200/shims``).  No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def knn(points: torch.Tensor, queries: Optional[torch.Tensor] = None, K: int = 1, exclude_self: bool = False,
        want_idx: bool = True, want_dists: bool = True, want_mean: bool = False
        ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[torch.Tensor]]:
    """K nearest neighbours of every query among ``points`` ([N,3] float32 CUDA), squared distances ascending.

    ``queries=None``: the cloud against itself (``exclude_self`` then removes point i from its own list).
    Returns ``(idx [Q,K] int64 or None, dist2 [Q,K] float32 or None, mean_dist2 [Q] float32 or None)``;
    ``idx`` is -1 (and ``dist2`` FLT_MAX) where the cloud has fewer than K eligible points."""
    lib = _lib.load()
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor (libsagars has no CPU path)")
    if not 1 <= int(K) <= 32:
        raise RuntimeError("K must be in 1..32")
    dev = points.device
    pts = points.detach().to(torch.float32).contiguous()
    qs = None
    if queries is not None and queries is not points:
        if queries.dim() != 2 or queries.shape[1] != 3:
            raise RuntimeError("queries must have dimensions (num_queries, 3)")
        if exclude_self:
            raise RuntimeError("exclude_self needs queries to be the point set itself")
        qs = queries.detach().to(device=dev, dtype=torch.float32).contiguous()
    N = int(pts.shape[0])
    Q = N if qs is None else int(qs.shape[0])
    with torch.cuda.device(dev):
        idx = torch.empty((Q, K), dtype=torch.int64, device=dev) if want_idx else None
        d2 = torch.empty((Q, K), dtype=torch.float32, device=dev) if want_dists else None
        mean = torch.empty((Q,), dtype=torch.float32, device=dev) if want_mean else None
        if Q == 0:
            return idx, d2, mean
        if N == 0:
            raise RuntimeError("knn: empty reference cloud")
        temp = torch.empty(int(lib.sagars_knn_temp_bytes(N)), dtype=torch.uint8, device=dev)
        ptr = lambda t: None if t is None else t.data_ptr()
        rc = lib.sagars_knn(dev.index if dev.index is not None else torch.cuda.current_device(), N, pts.data_ptr(), Q, ptr(qs),
                            int(K), 1 if exclude_self else 0, ptr(idx), ptr(d2), ptr(mean), temp.data_ptr(),
                            int(torch.cuda.current_stream(dev).cuda_stream))
        _lib.check(rc)
    return idx, d2, mean


def dist_cuda2(points: torch.Tensor) -> torch.Tensor:
    """``simple_knn._C.distCUDA2``: mean squared distance of every point to its 3 nearest other points."""
    return knn(points, None, K=3, exclude_self=True, want_idx=False, want_dists=False, want_mean=True)[2]
