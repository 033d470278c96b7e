#!/usr/bin/env python
"""Developer diagnostic: fused feature smoothing vs the reference's tensor expression at the bench's point count
(P = 1M, C = 32, K = 16 -> Ks = 8): CUDA-event times and the HBM-roofline view of the fused kernels."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seganygaussians_b200.smoothing import smooth_point_features, reference_expression  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    P, C, Ks = 1_000_000, 32, 8
    g = torch.Generator().manual_seed(0)
    F = torch.randn(P, C, generator=g).cuda().requires_grad_(True)
    # neighbours of a Morton-ordered cloud are mostly nearby rows: local indices with a few far ones
    base = torch.arange(P).unsqueeze(1)
    idx = (base + torch.randint(-2000, 2000, (P, Ks), generator=g)).clamp_(0, P - 1)
    idx[:, 0] = base[:, 0]
    idx = idx.cuda()
    w = torch.randn(P, C, generator=g).cuda()
    peak = 6580.3
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for norm in (False, True):
        def fwd_f(): return smooth_point_features(F, idx, norm)
        def fwd_r(): return reference_expression(F, idx, norm)
        def fb_f():
            F.grad = None; (smooth_point_features(F, idx, norm) * w).sum().backward()
        def fb_r():
            F.grad = None; (reference_expression(F, idx, norm) * w).sum().backward()
        tf, tr = timeit(fwd_f), timeit(fwd_r)
        tbf, tbr = timeit(fb_f), timeit(fb_r)
        # algorithmic bytes of the forward, per point: Ks int64 indices + Ks gathered rows + the output row (+ its norm)
        fwd_bytes = P * (8 * Ks + 4 * C * (Ks + 1) + (4 if norm else 0))
        print(f"normalize_output={norm}: forward fused {tf:.3f} ms ({fwd_bytes / tf / 1e6:.0f} GB/s algorithmic = "
              f"{100 * fwd_bytes / tf / 1e6 / peak:.0f}% of {peak:.0f} GB/s) vs torch {tr:.3f} ms;  "
              f"fwd+bwd(+loss) fused {tbf:.3f} ms vs torch {tbr:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
