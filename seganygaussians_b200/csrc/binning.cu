// binning.cu -- tile binning: offsets scan, (tile|depth) key emission, stable radix sort, tile ranges.
//
// Replaces the reference's CUB-based binning (CF cuda_rasterizer/rasterizer_impl.cu:70-138,277-317;
// SURVEY.md Appendix A.9).  Semantics that must hold bit-exactly:
//   * point_offsets = inclusive prefix sum of tiles_touched;
//   * key = (tile_y * tiles_x + tile_x) << 32 | float_bits(view depth), emitted y-outer / x-inner
//     starting at the Gaussian's exclusive offset;
//   * a STABLE ascending sort on key bits [0, 32 + msb(tiles)) -- ties keep emission order;
//   * ranges[tile] = [first, last+1) in the sorted list, (0,0) for empty tiles.
// The sort here is the library's own LSD radix sort (8-bit digits; per pass: per-block digit
// histogram -> per-digit row scan -> stable scatter using warp match ranking).  cub::DeviceRadixSort
// is kept behind SAGARS_FLAG_CUB_SORT purely as a cross-check for tests.
#include "common.cuh"
#include "math.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace sagars {

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

int launch_scan_block_sums(const Dims& d, GeomView g, cudaStream_t s, bool debug)
{
    const int nblk = (d.P + 255) / 256;
    scan_block_sums_kernel<<<1, 1024, 0, s>>>(g.block_sums, nblk, g.status);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
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

int launch_duplicate(const Dims& d, GeomView g, const int32_t* radii, uint64_t* keys, uint32_t* vals,
                     const uint32_t* n_dev, int cap, cudaStream_t s, bool debug)
{
    const int nblk = (d.P + 255) / 256;
    duplicate_kernel<<<nblk, 256, 0, s>>>(d.P, g.geo, g.depths, g.tiles_touched, g.block_sums, radii, g.point_offsets,
                                          keys, vals, d.tiles_x, d.tiles_y, n_dev, cap);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort of (u64 key, u32 value), 8-bit digits
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
radix_hist_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ n_dev, int cap, int shift,
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

__global__ void __launch_bounds__(256)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
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
        uint64_t key = 0;
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

int sort_num_passes(int end_bit) { return (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS; }

// Sorts the first n pairs of arrays laid out for `cap` pairs; n = *n_dev (device) or, with n_dev == nullptr, cap.
// Own sort: input must be in (keys_a, vals_a) when the pass count is even and in (keys_b, vals_b) when it is
// odd; the result always lands in (keys_a, vals_a).  CUB (host-side count only): input in A, *result_in_a
// tells where the result is.
int launch_sort_pairs(const uint32_t* n_dev, int cap, int end_bit, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b,
                      uint32_t* vals_b, void* temp, size_t temp_bytes, bool use_cub, bool* result_in_a,
                      cudaStream_t s, bool debug)
{
    *result_in_a = true;
    const int n = cap;
    if (n <= 0) return SAGARS_OK;
    if (use_cub) {
        if (n_dev != nullptr) { set_error("the CUB cross-check sort needs the host-side count"); return SAGARS_EINVAL; }
        cub::DoubleBuffer<uint64_t> dk(keys_a, keys_b);
        cub::DoubleBuffer<uint32_t> dv(vals_a, vals_b);
        size_t need = 0;
        SAGARS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, n, 0, end_bit, s));
        if (need > temp_bytes) {
            set_error("cub sort temp: need %zu bytes, reserved %zu", need, temp_bytes);
            return SAGARS_EINVAL;
        }
        SAGARS_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, dk, dv, n, 0, end_bit, s));
        count_launch(sort_num_passes(end_bit) + 1);
        if (debug) SAGARS_CUDA(cudaStreamSynchronize(s));
        *result_in_a = (dk.Current() == keys_a);
        return SAGARS_OK;
    }
    const int nblk = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    uint32_t* counts = (uint32_t*)temp;
    uint32_t* totals = (uint32_t*)((char*)temp + align_up((size_t)SORT_RADIX * (nblk + 1) * 4));
    const int npass = sort_num_passes(end_bit);
    uint64_t* kin = (npass & 1) ? keys_b : keys_a;
    uint32_t* vin = (npass & 1) ? vals_b : vals_a;
    uint64_t* kout = (npass & 1) ? keys_a : keys_b;
    uint32_t* vout = (npass & 1) ? vals_a : vals_b;
    for (int p = 0; p < npass; p++) {
        const int shift = p * SORT_RADIX_BITS;
        radix_hist_kernel<<<nblk, 256, 0, s>>>(kin, n_dev, cap, shift, counts, nblk);
        SAGARS_LAUNCH_CHECK(s, debug);
        radix_rowscan_kernel<<<SORT_RADIX * 32 / 256, 256, 0, s>>>(counts, nblk, totals);
        SAGARS_LAUNCH_CHECK(s, debug);
        radix_scatter_kernel<<<nblk, 256, 0, s>>>(kin, vin, kout, vout, n_dev, cap, shift, counts, totals, nblk);
        SAGARS_LAUNCH_CHECK(s, debug);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return SAGARS_OK;
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

int launch_tile_ranges(const uint32_t* n_dev, int cap, int num_tiles, const uint64_t* keys, uint2* ranges, cudaStream_t s, bool debug)
{
    SAGARS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s));
    if (cap > 0) {
        tile_ranges_kernel<<<(cap + 255) / 256, 256, 0, s>>>(n_dev, cap, keys, ranges);
        SAGARS_LAUNCH_CHECK(s, debug);
    }
    return SAGARS_OK;
}

}  // namespace sagars
