// sort_network.cuh -- the comparator schedule of the per-tile sort (tile_sort.cu), written so that the SAME index arithmetic
// compiles for the device and, with a plain C++ compiler, for the host (tests/test_tile_sort_network.py runs it on the CPU).
//
// Normalised bitonic network: for k = 2, 4, ..., N (N = next power of two >= n): one "flip" stage (comparator c pairs index
// i with its mirror inside the block of k), then half-cleaner stages with strides j = k/4, k/8, ..., 1.  EVERY comparator is
// ascending (lower index keeps the minimum), so an array of any length n sorts as if padded with +inf up to N: comparators
// whose upper index is >= n are skipped.  The N/2 comparators of a stage touch disjoint pairs: any assignment to threads
// works, with a barrier between stages.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SAGARS_SN_HD __host__ __device__ inline
#else
#define SAGARS_SN_HD inline
#endif

namespace sagars {

SAGARS_SN_HD uint32_t network_width(uint32_t n)
{
    uint32_t N = 1;
    while (N < n) N <<= 1;
    return N;
}

// comparator c (0 <= c < N/2) of the flip stage with block size k: lower index i, upper index l
SAGARS_SN_HD void flip_pair(uint32_t c, uint32_t k, uint32_t& i, uint32_t& l)
{
    const uint32_t hk = k >> 1, b = c / hk, o = c - b * hk;
    i = b * k + o;
    l = b * k + (k - 1u - o);
}

// comparator c of the half-cleaner stage with stride j
SAGARS_SN_HD void clean_pair(uint32_t c, uint32_t j, uint32_t& i, uint32_t& l)
{
    i = (c / j) * 2u * j + (c % j);
    l = i + j;
}

SAGARS_SN_HD void compare_exchange(uint64_t* a, uint32_t n, uint32_t i, uint32_t l)
{
    if (l < n) {
        const uint64_t x = a[i], y = a[l];
        if (x > y) { a[i] = y; a[l] = x; }
    }
}

// the share of thread `tid` (of `nthreads`) in one stage; j == 0 selects the flip stage of block size k
SAGARS_SN_HD void network_stage(uint64_t* a, uint32_t n, uint32_t N, uint32_t k, uint32_t j, uint32_t tid, uint32_t nthreads)
{
    const uint32_t half = N >> 1;
    for (uint32_t c = tid; c < half; c += nthreads) {
        uint32_t i, l;
        if (j == 0) flip_pair(c, k, i, l);
        else clean_pair(c, j, i, l);
        compare_exchange(a, n, i, l);
    }
}

}  // namespace sagars
