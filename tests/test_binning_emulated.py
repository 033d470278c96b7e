"""The kernels of the DEFAULT binning path (seganygaussians_b200/csrc/binning_kernels.cuh: block-sum scan, key emission, the
library's own LSD radix sort, tile ranges) executed on the CPU under the CUDA execution shim (tests/cuda_emu/) and compared
with the CPU oracle's binning state.  These kernels are also covered on the GPU (tests/test_parity_gpu.py); running the same
source here keeps the product's integer path under test in the CPU-only suite."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from tests import common
from seganygaussians_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    d = tempfile.mkdtemp(prefix="sagars_emu_")
    so = os.path.join(d, "libemu_binning.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "seganygaussians_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda_emu", "emu_binning.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_binning.restype = C.c_int
    L.emu_binning.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_longlong] + [C.c_void_p] * 5
    L.emu_binning_depth_first.restype = C.c_int
    L.emu_binning_depth_first.argtypes = L.emu_binning.argtypes
    return L


def _higher_msb(n):      # the reference's getHigherMsb (CF rasterizer_impl.cu:35-50), as api.cu restates it
    msb, step = 16, 16
    while step > 1:
        step //= 2
        msb = msb + step if (n >> msb) else msb - step
    return msb + 1 if (n >> msb) else msb


@pytest.mark.parametrize("case", [("typical", 1500, 56, 72, 2.0), ("one_tile", 200, 16, 16, 2.0), ("many_passes", 700, 130, 250, 3.0),
                                  ("two_sort_blocks", 1200, 32, 32, 20.0)], ids=lambda c: c[0])
def test_emulated_default_binning_reproduces_the_oracle(emu, case):
    name, P, H, W, sigma = case
    sc = synthetic.scene(P, H, W, 3, sigma_px=sigma)
    o = common.run_oracle(sc, 3, backward=False)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R = o.num_rendered
    if name == "two_sort_blocks":
        assert R > 4096
    geo = np.zeros((P, 8), np.float32)
    geo[:, 0:2] = o.means2D
    tt = np.ascontiguousarray(o.tiles_touched.astype(np.uint32))
    depths = np.ascontiguousarray(o.depths.astype(np.float32))
    radii = np.ascontiguousarray(o.radii.astype(np.int32))
    point_offsets = np.zeros(P, np.uint32)
    keys = np.zeros(R + 1, np.uint64)
    vals = np.zeros(R + 1, np.uint32)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    nr = np.zeros(1, np.uint32)
    p = lambda a: a.ctypes.data
    rc = emu.emu_binning(P, p(geo), p(depths), p(tt), p(radii), gx, gy, 32 + _higher_msb(gx * gy), R, -1,
                         p(point_offsets), p(keys), p(vals), p(ranges), p(nr))
    assert rc == 0 and int(nr[0]) == R
    assert np.array_equal(point_offsets, o.point_offsets)
    assert np.array_equal(keys[:R], o.keys)
    assert np.array_equal(vals[:R], o.point_list)
    assert np.array_equal(ranges, o.ranges)


def test_speculative_layout_of_the_default_path(emu):
    """Capacity hint larger than the count: same result in the first R slots; hint too small: the kernels must not write a single
    key, value or range (api.cu then re-issues the stages with the exact size)."""
    P, H, W = 900, 48, 64
    sc = synthetic.scene(P, H, W, 3, sigma_px=3.0)
    o = common.run_oracle(sc, 3, backward=False)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R = o.num_rendered
    geo = np.zeros((P, 8), np.float32)
    geo[:, 0:2] = o.means2D
    tt = np.ascontiguousarray(o.tiles_touched.astype(np.uint32))
    depths = np.ascontiguousarray(o.depths.astype(np.float32))
    radii = np.ascontiguousarray(o.radii.astype(np.int32))
    p = lambda a: a.ctypes.data
    for cap, expect_written in ((R + R // 4 + 4096, True), (R - 1, False)):
        point_offsets = np.zeros(P, np.uint32)
        keys = np.full(cap + 1, 0xFFFFFFFFFFFFFFFF, np.uint64)
        vals = np.full(cap + 1, 0xFFFFFFFF, np.uint32)
        ranges = np.zeros((gx * gy, 2), np.uint32)
        nr = np.zeros(1, np.uint32)
        rc = emu.emu_binning(P, p(geo), p(depths), p(tt), p(radii), gx, gy, 32 + _higher_msb(gx * gy), cap, R,
                             p(point_offsets), p(keys), p(vals), p(ranges), p(nr))
        assert rc == 0 and int(nr[0]) == R
        assert np.array_equal(point_offsets, o.point_offsets)          # written either way (they do not live in the binning buffer)
        if expect_written:
            assert np.array_equal(keys[:R], o.keys) and np.array_equal(vals[:R], o.point_list) and np.array_equal(ranges, o.ranges)
        else:
            assert np.all(keys == 0xFFFFFFFFFFFFFFFF) and np.all(vals == 0xFFFFFFFF) and not ranges.any()


@pytest.mark.parametrize("case", [("typical", 1500, 56, 72, 2.0), ("one_tile", 200, 16, 16, 2.0), ("many_tiles", 700, 130, 250, 3.0),
                                  ("two_sort_blocks", 1200, 32, 32, 20.0), ("odd_pass_count", 900, 64, 64, 3.0), ("two_tile_passes", 900, 272, 304, 4.0)],
                         ids=lambda c: c[0])
def test_emulated_depth_first_binning_reproduces_the_oracle(emu, case):
    """SAGARS_FLAG_DEPTH_FIRST: Gaussians sorted by depth, instances emitted in that order, one stable sort on the tile bits --
    point_offsets, point_list, the rebuilt 64-bit keys and the ranges must equal the reference's (oracle's) bit for bit, ties in
    depth included (duplicated Gaussians below)."""
    name, P, H, W, sigma = case
    sc = synthetic.scene(P, H, W, 3, sigma_px=sigma)
    g = sc.gauss
    g.means3D[1::7] = g.means3D[0::7][: len(g.means3D[1::7])]          # exact depth ties: order must fall back to the Gaussian index
    o = common.run_oracle(sc, 3, backward=False)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R = o.num_rendered
    geo = np.zeros((P, 8), np.float32)
    geo[:, 0:2] = o.means2D
    tt = np.ascontiguousarray(o.tiles_touched.astype(np.uint32))
    depths = np.ascontiguousarray(o.depths.astype(np.float32))
    radii = np.ascontiguousarray(o.radii.astype(np.int32))
    p = lambda a: a.ctypes.data
    for cap, n_dev, expect_written in ((R, -1, True), (R + R // 4 + 64, R, True), (R - 1, R, False)):
        point_offsets = np.zeros(P, np.uint32)
        keys = np.full(cap + 1, 0xFFFFFFFFFFFFFFFF, np.uint64)
        vals = np.full(cap + 1, 0xFFFFFFFF, np.uint32)
        ranges = np.zeros((gx * gy, 2), np.uint32)
        nr = np.zeros(1, np.uint32)
        rc = emu.emu_binning_depth_first(P, p(geo), p(depths), p(tt), p(radii), gx, gy, _higher_msb(gx * gy), cap, n_dev,
                                         p(point_offsets), p(keys), p(vals), p(ranges), p(nr))
        assert rc == 0 and int(nr[0]) == R
        assert np.array_equal(point_offsets, o.point_offsets)
        if expect_written:
            assert np.array_equal(vals[:R], o.point_list)
            assert np.array_equal(keys[:R], o.keys)
            assert np.array_equal(ranges, o.ranges)
        else:
            assert np.all(keys == 0xFFFFFFFFFFFFFFFF) and np.all(vals == 0xFFFFFFFF) and not ranges.any()
