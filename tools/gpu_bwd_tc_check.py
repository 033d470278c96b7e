"""GPU check of the tcgen05 backward (rasterizer.set_blend_kernels(backward="tc")) at BASELINE's full c2 size: every gradient against
the mma.sync warp kernel and, when oracle/_ref is present, against the unmodified reference extension; then CUDA-event times of
both backward kernels (library stage profile)."""
import sys, os, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import common
from seganygaussians_b200 import synthetic, rasterizer as R, _lib

sizes = [(20000, 270, 480), (1000000, 1080, 1920)] if "--small" not in sys.argv else [(20000, 270, 480)]
for (P, H, W) in sizes:
    K = 32
    sc = synthetic.scene(P, H, W, K)
    R.set_blend_kernels(backward="default")
    a = common.run_torch_impl("ours", sc, K)
    R.set_blend_kernels(backward="tc")
    b = common.run_torch_impl("ours", sc, K)
    R.set_blend_kernels()
    ok, lines = common.compare(b, a, ints=common.INT_FWD, floats=common.FLOAT_FWD + common.GRADS, verbose=True)
    print(f"[tc vs warp] P={P} {H}x{W}: {'OK' if ok else 'FAIL'}")
    print("\n".join(lines))
    if common.have_ref("cf"):
        ref = common.run_torch_impl("ref", sc, K)
        ok2, lines2 = common.compare(b, ref, ints=common.INT_FWD, floats=common.FLOAT_FWD + common.GRADS, verbose=True)
        print(f"[tc vs reference] P={P} {H}x{W}: {'OK' if ok2 else 'FAIL'}")
        print("\n".join(lines2))
