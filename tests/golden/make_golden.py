#!/usr/bin/env python
"""Generate the golden vectors of tests/golden/ from the UNMODIFIED reference CUDA extension.

Run on a GPU box (the reference has no CPU implementation):
    python oracle/build_ref.py            # once, where /root/reference is mounted (this container)
    gpurun -- python tests/golden/make_golden.py --out gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.

Each file holds the config (enough to regenerate the seeded inputs with seganygaussians_b200.synthetic) and the
reference's outputs: images, final_T, the integer binning state decoded from its scratch buffers
 and all gradients.  tests/test_oracle_golden.py pins the CPU oracle against them
(no GPU needed); tests/test_parity_gpu.py pins the CUDA path against them and against the oracle.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import common  # noqa: E402
from seganygaussians_b200 import synthetic  # noqa: E402

# name: (P, H, W, K, depth, use_sh, sh_degree, sh_coeffs)
CONFIGS = {
    "cf_small": (3000, 72, 104, 32, False, False, 0, 0),
    "base_small": (3000, 72, 104, 3, False, False, 0, 0),
    "base_sh_small": (2000, 64, 80, 3, False, True, 3, 16),
    "depth_sh_small": (2000, 64, 80, 3, True, True, 3, 16),
    "depth_small": (3000, 72, 104, 3, True, False, 0, 0),
}

FIELDS = common.FLOAT_FWD + common.INT_FWD + common.GRADS + ("means2D", "conic_opacity", "depths", "cov3D", "keys")


def make(name, out_dir):
    P, H, W, K, depth, use_sh, deg, M = CONFIGS[name]
    sc = synthetic.scene(P, H, W, K, sh_coeffs=M)
    r = common.run_torch_impl("ref", sc, K, depth=depth, use_sh=use_sh, sh_degree=deg)
    d = {"config": np.array([P, H, W, K, int(depth), int(use_sh), deg, M], np.int64)}
    for f in FIELDS:
        v = getattr(r, f, None)
        if v is not None:
            d[f] = np.asarray(v)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"[golden] {name}: R={r.num_rendered} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.dirname(os.path.abspath(__file__)))
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for n in (a.only or CONFIGS):
        make(n, a.out)
