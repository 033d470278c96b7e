#!/usr/bin/env python
"""Golden vectors for the neighbour-search shims: outputs of the UNMODIFIED reference extension ``simple_knn._C.distCUDA2``
(built by oracle/build_ref.py into oracle/_ref/simple_knn) on seeded point clouds.  Run on a GPU box:

    python tests/golden/make_golden_knn.py gpurun_out/simple_knn_small.npz      # then copy into tests/golden/knn/

The clouds are regenerated from the seeds by tests/knn_cases.py, so only the reference's outputs are stored."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import knn_cases  # noqa: E402


def main(out):
    so = os.path.join(ROOT, "oracle", "_ref", "simple_knn", "_C.so")
    spec = importlib.util.spec_from_file_location("_C", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = {}
    for name, pts in knn_cases.clouds().items():
        d = mod.distCUDA2(torch.from_numpy(pts).cuda()).float().cpu().numpy()
        res[name] = d
        print(name, pts.shape, float(d.mean()))
    np.savez_compressed(out, **res)


if __name__ == "__main__":
    main(sys.argv[1])
