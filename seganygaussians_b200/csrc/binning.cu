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
#include "binning_kernels.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace sagars {

int launch_scan_block_sums(const Dims& d, GeomView g, cudaStream_t s, bool debug)
{
    const int nblk = (d.P + 255) / 256;
    scan_block_sums_kernel<<<1, 1024, 0, s>>>(g.block_sums, nblk, g.status);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
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
        radix_hist_kernel<uint64_t><<<nblk, 256, 0, s>>>(kin, n_dev, cap, shift, counts, nblk);
        SAGARS_LAUNCH_CHECK(s, debug);
        radix_rowscan_kernel<<<SORT_RADIX * 32 / 256, 256, 0, s>>>(counts, nblk, totals);
        SAGARS_LAUNCH_CHECK(s, debug);
        radix_scatter_kernel<uint64_t><<<nblk, 256, 0, s>>>(kin, vin, kout, vout, n_dev, cap, shift, counts, totals, nblk);
        SAGARS_LAUNCH_CHECK(s, debug);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return SAGARS_OK;
}

// 32-bit keys: the pairs must be in (keys_a, vals_a) when the pass count is even and in (keys_b, vals_b) when it is odd; the result
// lands in (keys_a, vals_a).  n = *n_dev (device) or, with n_dev == nullptr, cap.
int launch_sort_pairs32(const uint32_t* n_dev, int cap, int end_bit, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                        void* temp, cudaStream_t s, bool debug)
{
    if (cap <= 0) return SAGARS_OK;
    const int nblk = (cap + SORT_CHUNK - 1) / SORT_CHUNK;
    uint32_t* counts = (uint32_t*)temp;
    uint32_t* totals = (uint32_t*)((char*)temp + align_up((size_t)SORT_RADIX * (nblk + 1) * 4));
    const int npass = sort_num_passes(end_bit);
    uint32_t* kin = (npass & 1) ? keys_b : keys_a;
    uint32_t* vin = (npass & 1) ? vals_b : vals_a;
    uint32_t* kout = (npass & 1) ? keys_a : keys_b;
    uint32_t* vout = (npass & 1) ? vals_a : vals_b;
    for (int p = 0; p < npass; p++) {
        const int shift = p * SORT_RADIX_BITS;
        radix_hist_kernel<uint32_t><<<nblk, 256, 0, s>>>(kin, n_dev, cap, shift, counts, nblk);
        SAGARS_LAUNCH_CHECK(s, debug);
        radix_rowscan_kernel<<<SORT_RADIX * 32 / 256, 256, 0, s>>>(counts, nblk, totals);
        SAGARS_LAUNCH_CHECK(s, debug);
        radix_scatter_kernel<uint32_t><<<nblk, 256, 0, s>>>(kin, vin, kout, vout, n_dev, cap, shift, counts, totals, nblk);
        SAGARS_LAUNCH_CHECK(s, debug);
        uint32_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return SAGARS_OK;
}

// Gaussians in depth order (g.ovals[0]) and the exclusive scan of their tile counts in that order (g.order_sums); also
// point_offsets.  Needs the preprocess outputs and the scanned block sums only, not the instance count.
int launch_depth_order(const Dims& d, GeomView g, cudaStream_t s, bool debug)
{
    const int nblk = (d.P + 255) / 256;
    order_keys_kernel<<<nblk, 256, 0, s>>>(d.P, g.depths, g.tiles_touched, g.block_sums, g.point_offsets, g.okeys[0], g.ovals[0]);
    SAGARS_LAUNCH_CHECK(s, debug);
    int rc = launch_sort_pairs32(nullptr, d.P, 32, g.okeys[0], g.ovals[0], g.okeys[1], g.ovals[1], g.osort_temp, s, debug);
    if (rc) return rc;
    sorted_block_sums_kernel<<<nblk, 256, 0, s>>>(d.P, g.ovals[0], g.tiles_touched, g.order_sums);
    SAGARS_LAUNCH_CHECK(s, debug);
    scan_block_sums_kernel<<<1, 1024, 0, s>>>(g.order_sums, nblk, g.status);      // status[1] = the same total once more
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_emit_sorted(const Dims& d, GeomView g, const int32_t* radii, uint32_t* tkeys, uint32_t* vals, const uint32_t* n_dev, int cap,
                       cudaStream_t s, bool debug)
{
    const int nblk = (d.P + 255) / 256;
    emit_sorted_kernel<<<nblk, 256, 0, s>>>(d.P, g.ovals[0], g.geo, g.tiles_touched, g.order_sums, radii, tkeys, vals, d.tiles_x, d.tiles_y,
                                            n_dev, cap);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_finalize_bins(const uint32_t* n_dev, int cap, int num_tiles, const uint32_t* tkeys, const uint32_t* point_list, const float* depths,
                         uint64_t* keys, uint2* ranges, cudaStream_t s, bool debug)
{
    SAGARS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s));
    if (cap > 0) {
        finalize_bins_kernel<<<(cap + 255) / 256, 256, 0, s>>>(n_dev, cap, tkeys, point_list, depths, keys, ranges);
        SAGARS_LAUNCH_CHECK(s, debug);
    }
    return SAGARS_OK;
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
