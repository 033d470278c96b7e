"""CPU: pin the oracle (oracle/sagars_oracle.c) against golden vectors produced by the UNMODIFIED reference CUDA
extension on a B200 (tests/golden/make_golden.py).  Integer state must match exactly; fp32 within RTOL=1e-4."""
import glob
import os

import numpy as np
import pytest

from tests import common
from seganygaussians_b200 import synthetic

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    P, H, W, K, depth, use_sh, deg, M = [int(v) for v in z["config"]]
    return z, dict(P=P, H=H, W=W, K=K, depth=bool(depth), use_sh=bool(use_sh), deg=deg, M=M)


def test_golden_files_present():
    names = {os.path.basename(p) for p in GOLDEN}
    assert {"cf_small.npz", "base_small.npz", "base_sh_small.npz"} <= names, names


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    z, c = _load(path)
    sc = synthetic.scene(c["P"], c["H"], c["W"], c["K"], sh_coeffs=c["M"])
    o = common.run_oracle(sc, c["K"], depth=c["depth"], use_sh=c["use_sh"], sh_degree=c["deg"])
    ref = common.SimpleNamespace(kind="golden(reference)", variant=o.variant)
    for f in z.files:
        if f != "config":
            setattr(ref, f, z[f])
    ref.num_rendered = int(z["num_rendered"])
    ok, lines = common.compare(o, ref, ints=common.INT_FWD,
                               floats=common.FLOAT_FWD + common.GRADS + ("means2D", "conic_opacity", "depths", "cov3D"),
                               verbose=False)
    assert ok, "\n".join(lines)
