"""The ACTUAL kernels of the sort-free binning (seganygaussians_b200/csrc/tile_sort_kernels.cuh) executed on the CPU.

The kernel header is compiled with g++ against a small CUDA execution shim (tests/cuda_emu/cuda_runtime.h: one OS thread per
CUDA thread, real atomics, barriers, warp shuffles) and run in the order tile_sort.cu queues the launches; the resulting
point_offsets / point_list / sorted keys / ranges must equal the CPU oracle's binning state bit for bit.  This is not a
substitute for the GPU test (tests/test_parity_gpu.py::test_tile_sort_binning_is_bit_identical) -- the host launch code, the
memory model and the speculative-capacity plumbing only exist there -- but every line of device code of the variant runs here,
with genuinely racing threads."""
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
    so = os.path.join(d, "libemu_tile_sort.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC",
                           "-I", os.path.join(ROOT, "tests", "cuda_emu"), "-I", os.path.join(ROOT, "seganygaussians_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cuda_emu", "emu_tile_sort.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_tile_binning.restype = C.c_int
    L.emu_tile_binning.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int] + [C.c_void_p] * 6
    return L


def _run(emu, o, W, H, cap=None, n_dev=-1, big_grid=2):
    P = len(o.radii)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    R = o.num_rendered
    cap = R if cap is None else cap
    geo = np.zeros((P, 8), np.float32)
    geo[:, 0:2] = o.means2D
    tt = np.ascontiguousarray(o.tiles_touched.astype(np.uint32))
    nblk = (P + 255) // 256
    sums = np.add.reduceat(np.concatenate([tt, np.zeros(nblk * 256 - P, np.uint32)]).astype(np.uint64), np.arange(0, nblk * 256, 256))
    block_excl = np.concatenate([[0], np.cumsum(sums)[:-1]]).astype(np.uint32)
    out = dict(point_offsets=np.zeros(P, np.uint32), ranges=np.zeros((T, 2), np.uint32), pairs=np.zeros(max(cap, 1), np.uint64),
               point_list=np.full(max(cap, 1), 0xFFFFFFFF, np.uint32), keys=np.zeros(max(cap, 1), np.uint64), queue=np.zeros(T + 1, np.uint32))
    depths = np.ascontiguousarray(o.depths.astype(np.float32))
    radii = np.ascontiguousarray(o.radii.astype(np.int32))
    p = lambda a: a.ctypes.data
    nq = emu.emu_tile_binning(P, p(geo), p(depths), p(tt), p(block_excl), p(radii), gx, gy, cap, n_dev, big_grid,
                              p(out["point_offsets"]), p(out["ranges"]), p(out["pairs"]), p(out["point_list"]), p(out["keys"]), p(out["queue"]))
    return out, nq


CASES = [
    # name, P, H, W, sigma_px, expected class of the longest segment
    ("typical", 1500, 56, 72, 2.0, "small"),
    ("ragged_sparse", 120, 50, 70, 1.0, "small"),
    ("mid_segments", 2600, 32, 32, 30.0, "large"),      # tiles with 1024 < n <= 8192 instances -> queue + 64 KB shared path
    ("huge_segments", 9000, 16, 32, 80.0, "huge"),      # n > 8192 -> chunks through shared memory + long-span stages in global memory
    ("huger_segments", 21000, 16, 16, 200.0, "huge"),   # one tile with ~20 k instances: k = 16384 and 32768 of the long-span path
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_emulated_kernels_reproduce_the_oracle_binning(emu, case):
    name, P, H, W, sigma, cls = case
    sc = synthetic.scene(P, H, W, 3, sigma_px=sigma)
    o = common.run_oracle(sc, 3, backward=False)
    n = o.ranges[:, 1].astype(np.int64) - o.ranges[:, 0]
    assert {"small": n.max() <= 1024, "large": 1024 < n.max() <= 8192, "huge": n.max() > 8192}[cls], n.max()
    out, nq = _run(emu, o, W, H)
    assert nq == int((n > 1024).sum())
    R = o.num_rendered
    assert np.array_equal(out["point_offsets"], o.point_offsets)
    assert np.array_equal(out["ranges"], o.ranges)
    assert np.array_equal(out["point_list"][:R], o.point_list)
    assert np.array_equal(out["keys"][:R], o.keys)


def test_speculative_layout_larger_than_needed_and_too_small(emu):
    sc = synthetic.scene(800, 48, 64, 3, sigma_px=3.0)
    o = common.run_oracle(sc, 3, backward=False)
    R = o.num_rendered
    # capacity hint larger than the count: identical result in the first R slots
    out, _ = _run(emu, o, 64, 48, cap=R + R // 4 + 4096, n_dev=R)
    assert np.array_equal(out["ranges"], o.ranges) and np.array_equal(out["point_list"][:R], o.point_list)
    assert np.array_equal(out["keys"][:R], o.keys)
    # hint too small: nothing may be written into the binning arrays, every tile reads as empty for the queued blend kernels
    out, nq = _run(emu, o, 64, 48, cap=R - 1, n_dev=R)
    assert nq == 0 and np.all(out["point_list"] == 0xFFFFFFFF) and np.all(out["pairs"] == 0)
    assert np.all(out["ranges"][:, 0] == out["ranges"][:, 1])
