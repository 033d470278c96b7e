// binning_kernels.cuh -- the device code of binning.cu (see there).  Free of host-side runtime calls so that
// tests/test_binning_emulated.py can compile these kernels for the CPU against tests/cuda_emu/ and run them against the oracle.
#pragma once
#include <stdint.h>
#include "math.cuh"

namespace sagars {

#ifndef SAGARS_SORT_CONSTANTS
#define SAGARS_SORT_CONSTANTS
constexpr int SORT_CHUNK = 4096;          // keys per sort block
constexpr int SORT_RADIX_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
#endif

// Element count of the binning arrays as the kernels see it.  `n_dev == nullptr`: the host knows it (`cap`).
// Otherwise the count lives in device memory (written by the scan) and `cap` is the capacity the arrays were
// laid out for; a count above the capacity means the speculative layout was too small: every kernel then
// does nothing and the host re-issues the stages with the exact size (api.cu).
__device__ __forceinline__ int live_count(const uint32_t* __restrict__ n_dev, int cap)
{
    if (n_dev == nullptr) return cap;
    const uint32_t n = *n_dev;
    return n > (uint32_t)cap ? 0 : (int)n;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of the per-preprocess-block sums (<= a few thousand entries): one CTA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
scan_block_sums_kernel(uint32_t* __restrict__ block_sums, int nblk, uint32_t* __restrict__ status)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s, slab_total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + tid;
        const uint32_t v = (i < nblk) ? block_sums[i] : 0u;
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
            warp_tot[lane] = winc - w;            // exclusive prefix over the 32 warps
            if (lane == 31) slab_total_s = winc;  // sum of this 1024-entry slab
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        if (i < nblk) block_sums[i] = carry + warp_tot[warp] + (inc - v);
        __syncthreads();
        if (tid == 0) carry_s = carry + slab_total_s;
        __syncthreads();
    }
    if (tid == 0) {
        block_sums[nblk] = carry_s;
        status[1] = carry_s;   // num_rendered
    }
}

// ---------------------------------------------------------------------------------------------
// key emission.  Same 256-Gaussian blocks as the preprocess kernel: local scan + block prefix gives
// each Gaussian its offset (and materialises point_offsets), then every Gaussian writes its tiles.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
duplicate_kernel(int P, const float* __restrict__ geo, const float* __restrict__ depths,
                 const uint32_t* __restrict__ tiles_touched,
                 const uint32_t* __restrict__ block_excl, const int32_t* __restrict__ radii,
                 uint32_t* __restrict__ point_offsets, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                 int tiles_x, int tiles_y, const uint32_t* __restrict__ n_dev, int cap)
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
    const uint32_t incl = block_excl[blockIdx.x] + wbase + inc;
    if (idx >= P) return;
    point_offsets[idx] = incl;
    if (n == 0) return;
    if (n_dev != nullptr && *n_dev > (uint32_t)cap) return;   // layout too small: nothing may be written

    uint32_t off = incl - n;
    const float4 r0 = *reinterpret_cast<const float4*>(geo + 8 * (size_t)idx);
    uint2 rmin, rmax;
    tile_rect(make_float2(r0.x, r0.y), radii[idx], rmin, rmax, tiles_x, tiles_y);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(depths[idx]);
    for (uint32_t y = rmin.y; y < rmax.y; y++) {
        for (uint32_t x = rmin.x; x < rmax.x; x++) {
            const uint64_t key = ((uint64_t)(y * (uint32_t)tiles_x + x) << 32) | depth_bits;
            keys[off] = key;
            vals[off] = (uint32_t)idx;
            off++;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort of (u64 key, u32 value), 8-bit digits
// ---------------------------------------------------------------------------------------------
template <class KeyT>
__global__ void __launch_bounds__(256)
radix_hist_kernel(const KeyT* __restrict__ keys, const uint32_t* __restrict__ n_dev, int cap, int shift,
                  uint32_t* __restrict__ counts, int nblk)
{
    __shared__ uint32_t hist[SORT_RADIX];
    const int tid = threadIdx.x;
    const int n = live_count(n_dev, cap);
    hist[tid] = 0;
    __syncthreads();
    const int start = blockIdx.x * SORT_CHUNK;
    const int end = min(n, start + SORT_CHUNK);
    for (int i = start + tid; i < end; i += 256) {
        const uint32_t dgt = (uint32_t)(keys[i] >> shift) & (SORT_RADIX - 1);
        atomicAdd(&hist[dgt], 1u);
    }
    __syncthreads();
    counts[(size_t)tid * nblk + blockIdx.x] = hist[tid];
}

// one warp per digit: exclusive scan of that digit's row of per-block counts; totals[d] = row sum
__global__ void __launch_bounds__(256)
radix_rowscan_kernel(uint32_t* __restrict__ counts, int nblk, uint32_t* __restrict__ totals)
{
    const int lane = threadIdx.x & 31;
    const int dgt = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (dgt >= SORT_RADIX) return;
    uint32_t* row = counts + (size_t)dgt * nblk;
    uint32_t carry = 0;
    for (int base = 0; base < nblk; base += 32) {
        const int i = base + lane;
        const uint32_t v = (i < nblk) ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (i < nblk) row[i] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) totals[dgt] = carry;
}

template <class KeyT>
__global__ void __launch_bounds__(256)
radix_scatter_kernel(const KeyT* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     KeyT* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     const uint32_t* __restrict__ n_dev, int cap, int shift,
                     const uint32_t* __restrict__ counts, const uint32_t* __restrict__ totals, int nblk)
{
    const int n = live_count(n_dev, cap);
    __shared__ uint32_t digit_base[SORT_RADIX];      // next output slot of each digit for this block
    __shared__ uint32_t warp_cnt[2][8][SORT_RADIX];  // per-round per-warp digit counts -> offsets
    __shared__ uint32_t scan_tmp[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // exclusive scan of the 256 digit totals (block-wide), plus this block's row prefix
    {
        const uint32_t v = totals[tid];
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) scan_tmp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) wbase += (w < warp) ? scan_tmp[w] : 0u;
        digit_base[tid] = wbase + inc - v + counts[(size_t)tid * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < 8; w++) warp_cnt[0][w][tid] = 0;
    }
    __syncthreads();

    const int start = blockIdx.x * SORT_CHUNK;
    const int rounds = (min(n, start + SORT_CHUNK) - start + 255) / 256;
    for (int r = 0; r < rounds; r++) {
        const int buf = r & 1;
        const int i = start + r * 256 + tid;
        const bool valid = i < n;
        KeyT key = 0;
        uint32_t val = 0;
        uint32_t dgt = 0xffffffffu - (uint32_t)lane;   // unique per lane: never matches a real digit
        if (valid) {
            key = keys_in[i];
            val = vals_in[i];
            dgt = (uint32_t)(key >> shift) & (SORT_RADIX - 1);
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, dgt);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[buf][warp][dgt] = __popc(peers);
        __syncthreads();
        {   // thread `tid` owns digit `tid`: turn the 8 per-warp counts into output offsets
            uint32_t s = digit_base[tid];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t c = warp_cnt[buf][w][tid];
                warp_cnt[buf][w][tid] = s;
                s += c;
            }
            digit_base[tid] = s;
        }
        __syncthreads();
        if (valid) {
            const uint32_t pos = warp_cnt[buf][warp][dgt] + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
#pragma unroll
        for (int w = 0; w < 8; w++) warp_cnt[buf ^ 1][w][tid] = 0;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// DEPTH-FIRST binning (SAGARS_FLAG_DEPTH_FIRST, the default): the same point_list / keys / ranges with a quarter of the sort
// traffic.  The reference sorts the R = sum(tiles_touched) duplicated instances on 32 depth bits + the tile bits (6 passes over
// R 12-byte pairs).  Equivalent, because an LSD radix sort is stable and so is every step below:
//   1. sort the P GAUSSIANS by depth bits (4 passes over P 8-byte pairs; culled Gaussians get the key 0xFFFFFFFF and sort last);
//   2. emit each Gaussian's instances in that order -> the instance array is already ordered by (depth, Gaussian index);
//   3. a stable sort of the instances on the TILE bits only (2 passes for up to 65,536 tiles) leaves every tile's segment in
//      (depth, index) order = exactly the order of the reference's stable sort on (tile | depth) keys emitted in index order;
//   4. the 64-bit keys of the reference are rebuilt from (tile, depth of the Gaussian) while the tile ranges are detected.
// ---------------------------------------------------------------------------------------------

// Step 1a (P threads, the preprocess kernel's 256-Gaussian blocks): point_offsets (inclusive scan of tiles_touched in index
// order: part of the geometry state) and the Gaussian's sort key.
__global__ void __launch_bounds__(256)
order_keys_kernel(int P, const float* __restrict__ depths, const uint32_t* __restrict__ tiles_touched,
                  const uint32_t* __restrict__ block_excl, uint32_t* __restrict__ point_offsets,
                  uint32_t* __restrict__ okeys, uint32_t* __restrict__ ovals)
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
    okeys[idx] = n ? __float_as_uint(depths[idx]) : 0xFFFFFFFFu;     // view depth > 0.2: raw bits are monotone
    ovals[idx] = (uint32_t)idx;
}

// Step 2a: per-256-block sums of tiles_touched taken in depth order (input of scan_block_sums_kernel)
__global__ void __launch_bounds__(256)
sorted_block_sums_kernel(int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched,
                         uint32_t* __restrict__ block_sums)
{
    __shared__ uint32_t warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 256 + tid;
    uint32_t n = (i < P) ? tiles_touched[order[i]] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) warp_tot[warp] = n;
    __syncthreads();
    if (tid == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) s += warp_tot[w];
        block_sums[blockIdx.x] = s;
    }
}

// Step 2b: instance emission in depth order: (tile id, Gaussian index) pairs, tiles y-outer / x-inner as the reference
__global__ void __launch_bounds__(256)
emit_sorted_kernel(int P, const uint32_t* __restrict__ order, const float* __restrict__ geo, const uint32_t* __restrict__ tiles_touched,
                   const uint32_t* __restrict__ block_excl, const int32_t* __restrict__ radii,
                   uint32_t* __restrict__ tkeys, uint32_t* __restrict__ vals, int tiles_x, int tiles_y,
                   const uint32_t* __restrict__ n_dev, int cap)
{
    __shared__ uint32_t warp_tot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 256 + tid;
    const uint32_t idx = (i < P) ? order[i] : 0u;
    const uint32_t n = (i < P) ? tiles_touched[idx] : 0u;
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
    if (i >= P || n == 0) return;
    if (n_dev != nullptr && *n_dev > (uint32_t)cap) return;   // layout too small: nothing may be written
    uint32_t off = block_excl[blockIdx.x] + wbase + inc - n;
    const float4 r0 = *reinterpret_cast<const float4*>(geo + 8 * (size_t)idx);
    uint2 rmin, rmax;
    tile_rect(make_float2(r0.x, r0.y), radii[idx], rmin, rmax, tiles_x, tiles_y);
    for (uint32_t y = rmin.y; y < rmax.y; y++) {
        for (uint32_t x = rmin.x; x < rmax.x; x++) {
            tkeys[off] = y * (uint32_t)tiles_x + x;
            vals[off] = idx;
            off++;
        }
    }
}

// Step 4: the reference's 64-bit keys (tile << 32 | depth bits) and the tile ranges from the tile-sorted instances
__global__ void __launch_bounds__(256)
finalize_bins_kernel(const uint32_t* __restrict__ n_dev, int cap, const uint32_t* __restrict__ tkeys, const uint32_t* __restrict__ point_list,
                     const float* __restrict__ depths, uint64_t* __restrict__ keys, uint2* __restrict__ ranges)
{
    const int R = live_count(n_dev, cap);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t cur = tkeys[i];
    keys[i] = ((uint64_t)cur << 32) | (uint64_t)__float_as_uint(depths[point_list[i]]);
    if (i == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = tkeys[i - 1];
        if (cur != prev) {
            ranges[prev].y = (uint32_t)i;
            ranges[cur].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[cur].y = (uint32_t)R;
}

// ---------------------------------------------------------------------------------------------
// tile ranges from the sorted keys (CF rasterizer_impl.cu:116-138)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t* __restrict__ n_dev, int cap, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges)
{
    const int R = live_count(n_dev, cap);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R) return;
    const uint32_t cur = (uint32_t)(keys[idx] >> 32);
    if (idx == 0) {
        ranges[cur].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
        if (cur != prev) {
            ranges[prev].y = (uint32_t)idx;
            ranges[cur].x = (uint32_t)idx;
        }
    }
    if (idx == R - 1) ranges[cur].y = (uint32_t)R;
}

}  // namespace sagars
