"""``pytorch3d.ops.knn_points`` for 3-D points, backed by libsagars' exact grid KNN (``csrc/knn.cu``).

Signature and return type follow pytorch3d (``knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1,
return_nn=False, return_sorted=True) -> _KNN(dists, idx, knn)``; dists are SQUARED L2 distances, ascending).  The
reference calls it as ``knn_points(xyz.unsqueeze(0), xyz.unsqueeze(0), K=K).idx.squeeze()``."""
import os
import sys
from collections import namedtuple

import torch

try:
    from seganygaussians_b200.knn import knn as _knn
except ImportError:   # shims/ was put on sys.path without the package root
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))))
    from seganygaussians_b200.knn import knn as _knn

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, lengths1=None, lengths2=None, norm: int = 2, K: int = 1, version: int = -1,
               return_nn: bool = False, return_sorted: bool = True):
    if norm != 2:
        raise NotImplementedError("only the L2 norm is provided by this stand-in")
    if p1.dim() != 3 or p2.dim() != 3 or p1.shape[2] != 3 or p2.shape[2] != 3 or p1.shape[0] != p2.shape[0]:
        raise ValueError("p1, p2 must be [N, P, 3] point clouds with the same batch size")
    N, P1 = p1.shape[0], p1.shape[1]
    if K > 32:
        raise NotImplementedError("K <= 32 in this stand-in")
    dists, idxs = [], []
    for b in range(N):
        n1 = P1 if lengths1 is None else int(lengths1[b])
        n2 = p2.shape[1] if lengths2 is None else int(lengths2[b])
        a, c = p1[b, :n1], p2[b, :n2]
        same = (a.data_ptr() == c.data_ptr()) and (n1 == n2)
        idx, d2, _ = _knn(c, None if same else a, K=K)
        if n1 < P1:   # pad like pytorch3d: zeros beyond the valid length
            idx = torch.cat([idx, idx.new_zeros((P1 - n1, K))]); d2 = torch.cat([d2, d2.new_zeros((P1 - n1, K))])
        missing = idx < 0
        if missing.any():
            idx = idx.masked_fill(missing, 0); d2 = d2.masked_fill(missing, 0.0)
        dists.append(d2); idxs.append(idx)
    dists, idx = torch.stack(dists), torch.stack(idxs)
    nn = None
    if return_nn:
        nn = torch.stack([p2[b][idx[b]] for b in range(N)])
    return _KNN(dists=dists, idx=idx, knn=nn)
