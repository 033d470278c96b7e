#!/usr/bin/env python
"""Developer diagnostic: sagars_knn at scene scale -- time vs the unmodified reference simple_knn (K = 3) and the cost of
the K = 16 map the feature smoothing uses; agreement of the two distCUDA2 implementations."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seganygaussians_b200.knn import knn, dist_cuda2  # noqa: E402


def ref_distcuda2():
    so = os.path.join(ROOT, "oracle", "_ref", "simple_knn", "_C.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("_C", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.distCUDA2


def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ref = ref_distcuda2()
    rng = np.random.default_rng(0)
    for N in (1_000_000, 5_000_000):
        # scene-like: surfaces (anisotropic clusters) + a thin cloud of floaters
        c = rng.standard_normal((64, 3)).astype(np.float32) * 4
        pts = c[rng.integers(0, 64, N)] + rng.standard_normal((N, 3)).astype(np.float32) * np.array([0.6, 0.05, 0.4], np.float32)
        pts[: N // 200] = rng.standard_normal((N // 200, 3)).astype(np.float32) * 30
        t = torch.from_numpy(pts.astype(np.float32)).cuda()
        ms3 = timeit(lambda: dist_cuda2(t))
        ms16 = timeit(lambda: knn(t, None, K=16, want_dists=False), n=3)
        line = f"N={N}: distCUDA2 (K=3, no self) {ms3:.2f} ms; knn_points K=16 {ms16:.2f} ms"
        if ref is not None:
            msr = timeit(lambda: ref(t), n=3)
            a, b = dist_cuda2(t), ref(t).float()
            rel = float(((a - b).abs() / b.clamp_min(1e-30)).max())
            line += f"; reference simple_knn {msr:.2f} ms; max rel diff {rel:.2e}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
