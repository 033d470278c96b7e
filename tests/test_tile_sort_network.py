"""The comparator schedule of the per-tile sort (seganygaussians_b200/csrc/sort_network.cuh) on the CPU.

The header is written to compile for the host too: this test builds a tiny shared library from it with g++ and drives
``network_stage`` exactly as the CUDA kernel does -- stage by stage, every "thread" of a CTA in turn, a barrier between
stages -- so the index arithmetic that runs on the GPU is the arithmetic checked here (the GPU-side equivalence of the whole
binning path with the radix sort is tests/test_parity_gpu.py::test_tile_sort_binning_is_bit_identical)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "seganygaussians_b200", "csrc", "sort_network.cuh")

HARNESS = r'''
#include "%s"
extern "C" void cta_sort(uint64_t* a, uint32_t n, uint32_t nthreads)
{
    using namespace sagars;
    const uint32_t N = network_width(n);
    for (uint32_t k = 2; k <= N; k <<= 1) {                 // the loop nest of network_sort<THREADS> in tile_sort.cu
        for (uint32_t t = 0; t < nthreads; t++) network_stage(a, n, N, k, 0u, t, nthreads);
        for (uint32_t j = k >> 2; j > 0; j >>= 1)
            for (uint32_t t = 0; t < nthreads; t++) network_stage(a, n, N, k, j, t, nthreads);
    }
}
// the schedule tile_sort_big_kernel uses for segments longer than its shared buffer (CH pairs): stages with span <= CH chunk by
// chunk through a scratch copy, only the longer spans on the whole array
extern "C" void cta_sort_chunked(uint64_t* seg, uint32_t n, uint32_t nthreads, uint32_t CH)
{
    using namespace sagars;
    const uint32_t N = network_width(n);
    uint64_t* sh = new uint64_t[CH];
    auto chunk_pass = [&](uint32_t base, auto body) {
        const uint32_t m = n - base < CH ? n - base : CH;
        for (uint32_t i = 0; i < m; i++) sh[i] = seg[base + i];
        body(m);
        for (uint32_t i = 0; i < m; i++) seg[base + i] = sh[i];
    };
    for (uint32_t base = 0; base < n; base += CH)
        chunk_pass(base, [&](uint32_t m) { cta_sort(sh, m, nthreads); });
    for (uint32_t k = 2 * CH; k <= N; k <<= 1) {
        for (uint32_t t = 0; t < nthreads; t++) network_stage(seg, n, N, k, 0u, t, nthreads);
        for (uint32_t j = k >> 2; j >= CH; j >>= 1)
            for (uint32_t t = 0; t < nthreads; t++) network_stage(seg, n, N, k, j, t, nthreads);
        for (uint32_t base = 0; base < n; base += CH)
            chunk_pass(base, [&](uint32_t m) {
                for (uint32_t j = CH >> 1; j > 0; j >>= 1)
                    for (uint32_t t = 0; t < nthreads; t++) network_stage(sh, m, CH, k, j, t, nthreads);
            });
    }
    delete[] sh;
}
// every comparator of every stage touches a distinct pair of indices (so the threads of a stage cannot race)
extern "C" int stages_are_disjoint(uint32_t N)
{
    using namespace sagars;
    unsigned char* seen = new unsigned char[N];
    int ok = 1;
    for (uint32_t k = 2; k <= N && ok; k <<= 1) {
        for (int pass = 0; ok; pass++) {
            const uint32_t j = pass == 0 ? 0u : (k >> (pass + 1));
            if (pass > 0 && j == 0) break;
            for (uint32_t i = 0; i < N; i++) seen[i] = 0;
            for (uint32_t c = 0; c < N / 2; c++) {
                uint32_t i, l;
                if (j == 0) flip_pair(c, k, i, l); else clean_pair(c, j, i, l);
                if (!(i < l) || l >= N || seen[i] || seen[l]) { ok = 0; break; }
                seen[i] = seen[l] = 1;
            }
        }
    }
    delete[] seen;
    return ok;
}
'''


@pytest.fixture(scope="module")
def lib():
    d = tempfile.mkdtemp(prefix="sagars_net_")
    src = os.path.join(d, "harness.cpp")
    with open(src, "w") as f:
        f.write(HARNESS % HDR)
    so = os.path.join(d, "libnet.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    L = ctypes.CDLL(so)
    L.cta_sort.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32]
    L.cta_sort_chunked.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    L.stages_are_disjoint.argtypes = [ctypes.c_uint32]
    L.stages_are_disjoint.restype = ctypes.c_int
    return L


def _pairs(rng, n, ties=True):
    """(depth bits << 32 | id) with unique ids and (optionally) many equal depths, as a tile's segment holds them."""
    depth = rng.randint(0, 7 if ties else 1 << 30, size=n).astype(np.uint64)
    ids = rng.permutation(4 * n + 1)[:n].astype(np.uint64)
    return (depth << np.uint64(32)) | ids


def test_every_length_up_to_1100(lib):
    rng = np.random.RandomState(0)
    for n in range(0, 1101):
        a = _pairs(rng, n, ties=(n % 2 == 0))
        want = np.sort(a)
        got = a.copy()
        lib.cta_sort(got.ctypes.data, n, 256)
        assert np.array_equal(got, want), n


@pytest.mark.parametrize("n,threads", [(1024, 256), (1025, 1024), (4097, 1024), (8192, 1024), (8193, 1024), (20011, 1024), (65537, 1024)])
def test_class_boundaries_and_long_segments(lib, n, threads):
    rng = np.random.RandomState(n)
    a = _pairs(rng, n)
    got = a.copy()
    lib.cta_sort(got.ctypes.data, n, threads)
    assert np.array_equal(got, np.sort(a))


def test_chunked_schedule_for_segments_longer_than_the_shared_buffer(lib):
    """Small chunk sizes make the long-span machinery run on short arrays: every n in a range, several chunk sizes."""
    rng = np.random.RandomState(7)
    for CH in (8, 32, 64):
        for n in list(range(CH + 1, 6 * CH + 3)) + [17 * CH + 5, 33 * CH - 1]:
            a = _pairs(rng, n, ties=(n % 3 == 0))
            got = a.copy()
            lib.cta_sort_chunked(got.ctypes.data, n, 16, CH)
            assert np.array_equal(got, np.sort(a)), (CH, n)
    for n in (8193, 20011, 70001):                       # the kernel's own chunk size
        a = _pairs(rng, n)
        got = a.copy()
        lib.cta_sort_chunked(got.ctypes.data, n, 1024, 8192)
        assert np.array_equal(got, np.sort(a)), n


def test_already_sorted_reverse_and_constant_depth(lib):
    n = 777
    ids = np.arange(n, dtype=np.uint64)
    for a in (ids.copy(), ids[::-1].copy(), (np.uint64(5) << np.uint64(32)) | np.random.RandomState(1).permutation(n).astype(np.uint64)):
        got = a.copy()
        lib.cta_sort(got.ctypes.data, n, 256)
        assert np.array_equal(got, np.sort(a))


def test_comparators_of_a_stage_are_disjoint(lib):
    for N in (2, 4, 8, 64, 1024, 8192):
        assert lib.stages_are_disjoint(N) == 1, N


def test_stable_order_equals_sort_by_depth_then_id():
    """The equivalence the device code relies on: within one tile ids are unique and the reference emits them in ascending
    order, so a STABLE sort by depth bits alone equals an ordinary sort by (depth bits, id)."""
    rng = np.random.RandomState(3)
    n = 5000
    ids = np.sort(rng.permutation(10 * n)[:n]).astype(np.uint64)          # emission order: ascending id
    depth = rng.randint(0, 50, size=n).astype(np.uint64)
    stable = ids[np.argsort(depth, kind="stable")]
    composite = np.sort((depth << np.uint64(32)) | ids) & np.uint64(0xFFFFFFFF)
    assert np.array_equal(stable, composite)


def test_count_scan_scatter_sort_reproduces_the_oracle_binning(lib):
    """The whole tile_sort.cu algorithm emulated in numpy on a real scene (scatter order randomised, as the atomics leave it)
    against the CPU oracle's binning state: point_list, sorted keys and ranges must come out identical."""
    from tests import common
    from seganygaussians_b200 import synthetic
    for (P, H, W, sigma) in [(3000, 72, 104, 2.0), (1500, 40, 40, 25.0)]:
        sc = synthetic.scene(P, H, W, 3, sigma_px=sigma)
        o = common.run_oracle(sc, 3, backward=False)
        gx, gy = (W + 15) // 16, (H + 15) // 16
        T = gx * gy
        # what the scatter kernel sees: per Gaussian its tile rectangle (recovered from the oracle's keys) and depth bits
        tiles = (o.keys >> np.uint64(32)).astype(np.int64)
        depth_bits = (o.keys & np.uint64(0xFFFFFFFF))
        ids = o.point_list.astype(np.uint64)
        rng = np.random.RandomState(P)
        perm = rng.permutation(len(ids))                                     # arbitrary arrival order of the atomics
        counts = np.bincount(tiles, minlength=T)
        start = np.concatenate([[0], np.cumsum(counts)[:-1]])
        cursor = start.copy()
        pairs = np.zeros(len(ids), np.uint64)
        for q in perm:
            t = tiles[q]
            pairs[cursor[t]] = (depth_bits[q] << np.uint64(32)) | ids[q]
            cursor[t] += 1
        point_list = np.zeros(len(ids), np.uint32)
        keys = np.zeros(len(ids), np.uint64)
        ranges = np.zeros((T, 2), np.uint32)
        for t in range(T):
            n = int(cursor[t] - start[t])
            if n == 0:
                continue
            seg = pairs[start[t]:start[t] + n].copy()
            lib.cta_sort(seg.ctypes.data, n, 256)
            point_list[start[t]:start[t] + n] = (seg & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            keys[start[t]:start[t] + n] = (np.uint64(t) << np.uint64(32)) | (seg >> np.uint64(32))
            ranges[t] = (start[t], start[t] + n)
        assert np.array_equal(point_list, o.point_list)
        assert np.array_equal(keys, o.keys)
        assert np.array_equal(ranges, o.ranges)
