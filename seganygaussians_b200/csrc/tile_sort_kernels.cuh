// tile_sort_kernels.cuh -- the device code of tile_sort.cu (see there for the algorithm).  Kept free of host-side runtime
// calls so that tests/test_tile_sort_emulated.py can compile these very kernels for the CPU against a small CUDA execution
// shim (tests/cuda_emu/) and run them multi-threaded against the oracle's binning state.
#pragma once
#include <stdint.h>
#include "math.cuh"
#include "sort_network.cuh"

#ifndef SAGARS_DYNAMIC_SMEM
#define SAGARS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sagars {


// ---- 1. point_offsets + per-tile counts --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_count_kernel(int P, const float* __restrict__ geo, const uint32_t* __restrict__ tiles_touched,
                  const uint32_t* __restrict__ block_excl, const int32_t* __restrict__ radii,
                  uint32_t* __restrict__ point_offsets, uint2* __restrict__ ranges, int tiles_x, int tiles_y)
{
    __shared__ uint32_t warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int idx = blockIdx.x * 256 + tid;
    const uint32_t n = (idx < P) ? tiles_touched[idx] : 0u;
    uint32_t inc = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) wbase += (w < warp) ? warp_tot[w] : 0u;
    if (idx >= P) return;
    point_offsets[idx] = block_excl[blockIdx.x] + wbase + inc;
    if (n == 0) return;
    const float4 r0 = *reinterpret_cast<const float4*>(geo + 8 * (size_t)idx);
    uint2 rmin, rmax;
    tile_rect(make_float2(r0.x, r0.y), radii[idx], rmin, rmax, tiles_x, tiles_y);
    for (uint32_t y = rmin.y; y < rmax.y; y++)
        for (uint32_t x = rmin.x; x < rmax.x; x++) atomicAdd(&ranges[y * (uint32_t)tiles_x + x].y, 1u);
}

// ---- 2. exclusive scan of the tile counts: ranges[t] = (start, start) -------------------------------------------------------
__global__ void __launch_bounds__(1024)
tile_scan_kernel(uint2* __restrict__ ranges, int num_tiles, uint32_t* __restrict__ queue)
{
    if (threadIdx.x == 0) queue[0] = 0u;   // length of the long-segment queue the sort kernels use
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s, slab_total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < num_tiles; base += 1024) {
        const int i = base + tid;
        const uint32_t v = (i < num_tiles) ? ranges[i].y : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = warp_tot[lane];
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_tot[lane] = winc - w;
            if (lane == 31) slab_total_s = winc;
        }
        __syncthreads();
        const uint32_t start = carry_s + warp_tot[warp] + (inc - v);
        if (i < num_tiles) ranges[i] = make_uint2(start, start);
        __syncthreads();
        if (tid == 0) carry_s += slab_total_s;
        __syncthreads();
    }
}

// ---- 3. scatter the (depth bits, id) pairs into their tile's segment ------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_scatter_kernel(int P, const float* __restrict__ geo, const float* __restrict__ depths,
                    const uint32_t* __restrict__ tiles_touched, const int32_t* __restrict__ radii,
                    uint2* __restrict__ ranges, uint64_t* __restrict__ pairs, int tiles_x, int tiles_y,
                    const uint32_t* __restrict__ n_dev, int cap)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P || tiles_touched[idx] == 0) return;
    if (n_dev != nullptr && *n_dev > (uint32_t)cap) return;   // layout too small: nothing may be written
    const float4 r0 = *reinterpret_cast<const float4*>(geo + 8 * (size_t)idx);
    uint2 rmin, rmax;
    tile_rect(make_float2(r0.x, r0.y), radii[idx], rmin, rmax, tiles_x, tiles_y);
    const uint64_t pair = ((uint64_t)__float_as_uint(depths[idx]) << 32) | (uint64_t)(uint32_t)idx;
    for (uint32_t y = rmin.y; y < rmax.y; y++)
        for (uint32_t x = rmin.x; x < rmax.x; x++) {
            const uint32_t slot = atomicAdd(&ranges[y * (uint32_t)tiles_x + x].y, 1u);
            pairs[slot] = pair;
        }
}

// ---- 4. per-tile sort (comparator schedule: sort_network.cuh) ----------------------------------------------------------------
// sorts a[0, n) ascending; a may be shared or global memory of this CTA's tile.  All threads of the CTA take part.
template <int THREADS>
__device__ __forceinline__ void network_sort(uint64_t* a, uint32_t n)
{
    const uint32_t N = network_width(n);
    for (uint32_t k = 2; k <= N; k <<= 1) {
        network_stage(a, n, N, k, 0u, threadIdx.x, THREADS);
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            network_stage(a, n, N, k, j, threadIdx.x, THREADS);
            __syncthreads();
        }
    }
}

constexpr uint32_t TSORT_SMALL = 1024;   // pairs sorted by a 256-thread CTA in 8 KB of shared memory
constexpr uint32_t TSORT_LARGE = 8192;   // pairs sorted by a 1024-thread CTA in 64 KB of shared memory
constexpr int TSORT_BIG_CTAS = 2 * 148;  // persistent grid of the long-segment kernel (it usually finds an empty queue)

template <int THREADS>
__device__ __forceinline__ void write_sorted(const uint64_t* a, uint32_t n, uint32_t tile, uint32_t start,
                                             uint32_t* __restrict__ point_list, uint64_t* __restrict__ keys)
{
    const uint64_t hi = (uint64_t)tile << 32;
    for (uint32_t i = threadIdx.x; i < n; i += THREADS) {
        const uint64_t p = a[i];
        point_list[start + i] = (uint32_t)p;
        keys[start + i] = hi | (p >> 32);
    }
}

// One CTA per tile.  Segments of up to TSORT_SMALL pairs are sorted here, in shared memory; empty tiles get the reference's
// (0, 0) range; longer segments are queued for tile_sort_big_kernel (queue[0] = count, queue[1 + i] = tile id).
__global__ void __launch_bounds__(256)
tile_sort_small_kernel(uint2* __restrict__ ranges, const uint64_t* __restrict__ pairs, uint32_t* __restrict__ point_list,
                       uint64_t* __restrict__ keys, uint32_t* __restrict__ queue, const uint32_t* __restrict__ n_dev, int cap)
{
    __shared__ uint64_t a[TSORT_SMALL];
    if (n_dev != nullptr && *n_dev > (uint32_t)cap) return;
    const uint32_t tile = blockIdx.x;
    const uint2 rg = ranges[tile];
    const uint32_t n = rg.y - rg.x;
    if (n == 0) {
        if (threadIdx.x == 0) ranges[tile] = make_uint2(0u, 0u);
        return;
    }
    if (n > TSORT_SMALL) {
        if (threadIdx.x == 0) queue[1u + atomicAdd(&queue[0], 1u)] = tile;
        return;
    }
    for (uint32_t i = threadIdx.x; i < n; i += 256) a[i] = pairs[rg.x + i];
    __syncthreads();
    network_sort<256>(a, n);
    write_sorted<256>(a, n, tile, rg.x, point_list, keys);
}

// Persistent CTAs over the queue of long segments: up to TSORT_LARGE pairs in 64 KB of shared memory; longer ones chunk by chunk
// through the same buffer, with only the long-span stages of the network in global memory (__syncthreads orders the CTA's own
// global accesses between stages).
__global__ void __launch_bounds__(1024)
tile_sort_big_kernel(const uint2* __restrict__ ranges, uint64_t* __restrict__ pairs, uint32_t* __restrict__ point_list,
                     uint64_t* __restrict__ keys, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ n_dev, int cap)
{
    SAGARS_DYNAMIC_SMEM(tsort_smem);
    uint64_t* sh = reinterpret_cast<uint64_t*>(tsort_smem);
    if (n_dev != nullptr && *n_dev > (uint32_t)cap) return;
    const uint32_t count = queue[0];
    for (uint32_t q = blockIdx.x; q < count; q += gridDim.x) {
        const uint32_t tile = queue[1u + q];
        const uint2 rg = ranges[tile];
        const uint32_t n = rg.y - rg.x;
        uint64_t* seg = pairs + rg.x;
        if (n <= TSORT_LARGE) {
            for (uint32_t i = threadIdx.x; i < n; i += 1024) sh[i] = seg[i];
            __syncthreads();
            network_sort<1024>(sh, n);
            write_sorted<1024>(sh, n, tile, rg.x, point_list, keys);
        } else {
            // Longer than the shared buffer.  The network's stages with span <= TSORT_LARGE only ever pair elements of the same
            // aligned TSORT_LARGE-chunk, so they run chunk by chunk in shared memory; only the few stages with a longer span
            // (one flip + log2(k / TSORT_LARGE) - 1 half-cleaners per k > TSORT_LARGE) touch global memory.  Same comparators
            // in the same order as network_sort on the whole segment -- grouped by where their operands live.
            constexpr uint32_t CH = TSORT_LARGE;
            const uint32_t N = network_width(n);
            for (uint32_t base = 0; base < n; base += CH) {            // every stage with k <= CH: sort each chunk
                const uint32_t m = n - base < CH ? n - base : CH;
                for (uint32_t i = threadIdx.x; i < m; i += 1024) sh[i] = seg[base + i];
                __syncthreads();
                network_sort<1024>(sh, m);
                for (uint32_t i = threadIdx.x; i < m; i += 1024) seg[base + i] = sh[i];
                __syncthreads();
            }
            for (uint32_t k = 2 * CH; k <= N; k <<= 1) {
                network_stage(seg, n, N, k, 0u, threadIdx.x, 1024);     // flip, span k
                __syncthreads();
                for (uint32_t j = k >> 2; j >= CH; j >>= 1) {           // half-cleaners that cross chunk boundaries
                    network_stage(seg, n, N, k, j, threadIdx.x, 1024);
                    __syncthreads();
                }
                for (uint32_t base = 0; base < n; base += CH) {         // strides CH/2 ... 1: inside a chunk
                    const uint32_t m = n - base < CH ? n - base : CH;
                    for (uint32_t i = threadIdx.x; i < m; i += 1024) sh[i] = seg[base + i];
                    __syncthreads();
                    for (uint32_t j = CH >> 1; j > 0; j >>= 1) {
                        network_stage(sh, m, CH, k, j, threadIdx.x, 1024);
                        __syncthreads();
                    }
                    for (uint32_t i = threadIdx.x; i < m; i += 1024) seg[base + i] = sh[i];
                    __syncthreads();
                }
            }
            write_sorted<1024>(seg, n, tile, rg.x, point_list, keys);
        }
        __syncthreads();   // the shared buffer is reused by the next queue entry
    }
}

}  // namespace sagars
