"""Randomised inputs straight into the binning kernels (both paths) under the CPU execution shim, against an independent NumPy
statement of what the reference's binning produces: instances = every (Gaussian, tile) of the Gaussian's tile rectangle
(`getRect`, CF auxiliary.h:46-57), ordered by (tile id, depth bits, emission order) -- i.e. a stable sort of the 64-bit keys
`tile << 32 | float_bits(depth)` in emission order (CF rasterizer_impl.cu:70-111, 298-317).  Sizes around the kernels' block
boundaries, invisible Gaussians, exact depth ties, a splat covering the whole grid, an empty frame."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOCK = 16


@pytest.fixture(scope="module")
def emu():
    d = tempfile.mkdtemp(prefix="sagars_emu_")
    so = os.path.join(d, "libemu_binning.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "seganygaussians_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda_emu", "emu_binning.cpp"), "-o", so])
    L = C.CDLL(so)
    for f in (L.emu_binning, L.emu_binning_depth_first):
        f.restype = C.c_int
        f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_longlong] + [C.c_void_p] * 5
    return L


def _higher_msb(n):      # the reference's getHigherMsb (CF rasterizer_impl.cu:35-50)
    msb, step = 16, 16
    while step > 1:
        step //= 2
        msb = msb + step if (n >> msb) else msb - step
    return msb + 1 if (n >> msb) else msb


def _rects(xy, radii, gx, gy):
    """getRect in fp32 with C's truncating float -> int conversion."""
    r = radii.astype(np.float32)
    def lo(p, g):
        return np.clip(np.trunc((p - r) / np.float32(BLOCK)).astype(np.int64), 0, g)
    def hi(p, g):
        return np.clip(np.trunc((p + r + np.float32(BLOCK - 1)) / np.float32(BLOCK)).astype(np.int64), 0, g)
    return lo(xy[:, 0], gx), lo(xy[:, 1], gy), hi(xy[:, 0], gx), hi(xy[:, 1], gy)


def _expected(xy, depths, radii, gx, gy):
    x0, y0, x1, y1 = _rects(xy, radii, gx, gy)
    touched = np.where(radii > 0, (x1 - x0) * (y1 - y0), 0).astype(np.uint32)
    keys, vals = [], []
    bits = depths.view(np.uint32).astype(np.uint64)
    for i in np.nonzero(touched)[0]:
        for y in range(y0[i], y1[i]):
            for x in range(x0[i], x1[i]):
                keys.append((np.uint64(y * gx + x) << np.uint64(32)) | bits[i])
                vals.append(i)
    keys = np.array(keys, np.uint64)
    vals = np.array(vals, np.uint32)
    order = np.argsort(keys, kind="stable")
    keys, vals = keys[order], vals[order]
    ranges = np.zeros((gx * gy, 2), np.uint32)
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        idx = np.nonzero(tiles == t)[0]
        ranges[t] = (idx[0], idx[-1] + 1)
    return touched, np.cumsum(touched, dtype=np.uint64).astype(np.uint32), keys, vals, ranges


CASES = [  # name, P, W, H, max radius, fraction invisible, seed
    ("one", 1, 40, 40, 30, 0.0, 1), ("block_minus_one", 255, 64, 48, 12, 0.2, 2), ("block", 256, 64, 48, 12, 0.2, 3),
    ("block_plus_one", 257, 64, 48, 12, 0.2, 4), ("ragged_image", 700, 75, 101, 20, 0.3, 5), ("wide_grid", 500, 400, 40, 25, 0.1, 6),
    ("mostly_invisible", 900, 96, 96, 10, 0.95, 7), ("all_invisible", 300, 64, 64, 10, 1.0, 8), ("large_splats", 120, 128, 128, 200, 0.0, 9),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("path", ["global_sort", "depth_first"])
def test_binning_kernels_on_random_inputs(emu, case, path):
    name, P, W, H, rmax, invisible, seed = case
    rng = np.random.default_rng(seed)
    gx, gy = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    xy = np.stack([rng.uniform(-20, W + 20, P), rng.uniform(-20, H + 20, P)], 1).astype(np.float32)
    radii = rng.integers(1, rmax + 1, P).astype(np.int32)
    radii[rng.random(P) < invisible] = 0
    depths = rng.uniform(0.2, 30.0, P).astype(np.float32)
    depths[1::5] = depths[0::5][: len(depths[1::5])]                       # exact ties: emission (index) order decides
    if name == "large_splats":
        xy[0], radii[0] = (W / 2, H / 2), 4 * max(W, H)                    # one splat on every tile
    touched, offsets, keys, vals, ranges = _expected(xy, depths, radii, gx, gy)
    R = int(touched.sum())
    geo = np.zeros((P, 8), np.float32)
    geo[:, 0:2] = xy
    p = lambda a: a.ctypes.data
    fn = emu.emu_binning if path == "global_sort" else emu.emu_binning_depth_first
    bits = 32 + _higher_msb(gx * gy) if path == "global_sort" else _higher_msb(gx * gy)
    for cap, n_dev in ((R, -1), (R + 100, R)):                             # exact layout; speculative capacity with the device count
        got_off = np.zeros(P, np.uint32)
        got_keys = np.full(cap + 1, 0xFFFFFFFFFFFFFFFF, np.uint64)
        got_vals = np.full(cap + 1, 0xFFFFFFFF, np.uint32)
        got_ranges = np.zeros((gx * gy, 2), np.uint32)
        nr = np.zeros(1, np.uint32)
        rc = fn(P, p(geo), p(depths), p(touched), p(radii), gx, gy, bits, cap, n_dev, p(got_off), p(got_keys), p(got_vals), p(got_ranges), p(nr))
        assert rc == 0 and int(nr[0]) == R
        assert np.array_equal(got_off, offsets)
        assert np.array_equal(got_vals[:R], vals)
        assert np.array_equal(got_keys[:R], keys)
        assert np.array_equal(got_ranges, ranges)
