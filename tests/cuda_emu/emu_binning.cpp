// Host harness: the kernels of seganygaussians_b200/csrc/binning_kernels.cuh (the DEFAULT binning path: block-sum scan, key
// emission, the library's LSD radix sort, tile ranges) compiled against the execution shim and run in the order api.cu /
// binning.cu queue them.  The ping-pong bookkeeping of launch_sort_pairs is restated here (the launcher itself uses <<< >>>).
// TEST INFRASTRUCTURE ONLY.
#include "../../include/sagars.h"
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "binning_kernels.cuh"
#include <vector>

using namespace sagars;

extern "C" int emu_binning(int P, const float* geo, const float* depths, const uint32_t* tiles_touched, const int32_t* radii,
                           int tiles_x, int tiles_y, int end_bit, int cap, long long n_dev_value,
                           uint32_t* point_offsets, uint64_t* keys_out, uint32_t* vals_out, uint2* ranges, uint32_t* num_rendered)
{
    const int nblk_p = (P + 255) / 256;
    // what the preprocess kernel leaves behind: per-256-block sums of tiles_touched
    std::vector<uint32_t> block_sums(nblk_p + 1, 0u);
    for (int i = 0; i < P; i++) block_sums[i / 256] += tiles_touched[i];
    uint32_t status[8] = {0};
    cuda_emu::launch(1, 1024, 0, scan_block_sums_kernel, block_sums.data(), nblk_p, status);
    *num_rendered = status[1];
    uint32_t n_dev_store = n_dev_value < 0 ? 0u : (uint32_t)n_dev_value;
    const uint32_t* n_dev = n_dev_value < 0 ? nullptr : &n_dev_store;
    const int n = cap;
    std::vector<uint64_t> keys_b((size_t)n + 1);
    std::vector<uint32_t> vals_b((size_t)n + 1);
    const int npass = (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    uint64_t* kin = (npass & 1) ? keys_b.data() : keys_out;
    uint32_t* vin = (npass & 1) ? vals_b.data() : vals_out;
    uint64_t* kout = (npass & 1) ? keys_out : keys_b.data();
    uint32_t* vout = (npass & 1) ? vals_out : vals_b.data();
    if (n > 0) {
        cuda_emu::launch(nblk_p, 256, 0, duplicate_kernel, P, geo, depths, tiles_touched, (const uint32_t*)block_sums.data(), radii,
                         point_offsets, kin, vin, tiles_x, tiles_y, n_dev, cap);
        const int nblk = (n + SORT_CHUNK - 1) / SORT_CHUNK;
        std::vector<uint32_t> counts((size_t)SORT_RADIX * (nblk + 1)), totals(SORT_RADIX);
        for (int p = 0; p < npass; p++) {
            const int shift = p * SORT_RADIX_BITS;
            cuda_emu::launch(nblk, 256, 0, radix_hist_kernel<uint64_t>, (const uint64_t*)kin, n_dev, cap, shift, counts.data(), nblk);
            cuda_emu::launch(SORT_RADIX * 32 / 256, 256, 0, radix_rowscan_kernel, counts.data(), nblk, totals.data());
            cuda_emu::launch(nblk, 256, 0, radix_scatter_kernel<uint64_t>, (const uint64_t*)kin, (const uint32_t*)vin, kout, vout, n_dev, cap, shift,
                             (const uint32_t*)counts.data(), (const uint32_t*)totals.data(), nblk);
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        if (kin != keys_out) return -1;   // the ping-pong must end in the final arrays
    }
    std::memset(ranges, 0, sizeof(uint2) * (size_t)tiles_x * tiles_y);
    if (n > 0) cuda_emu::launch((n + 255) / 256, 256, 0, tile_ranges_kernel, n_dev, cap, (const uint64_t*)keys_out, ranges);
    return 0;
}

// The depth-first path (SAGARS_FLAG_DEPTH_FIRST) in launch order: order keys -> sort of the P Gaussians on 32 depth bits ->
// block sums in depth order -> scan -> emission -> sort of the instances on the tile bits -> keys / ranges.
static void sort32(const uint32_t* n_dev, int cap, int end_bit, uint32_t* ka, uint32_t* va, uint32_t* kb, uint32_t* vb)
{
    const int nblk = (cap + SORT_CHUNK - 1) / SORT_CHUNK;
    std::vector<uint32_t> counts((size_t)SORT_RADIX * (nblk + 1)), totals(SORT_RADIX);
    const int npass = (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    uint32_t* kin = (npass & 1) ? kb : ka; uint32_t* vin = (npass & 1) ? vb : va;
    uint32_t* kout = (npass & 1) ? ka : kb; uint32_t* vout = (npass & 1) ? va : vb;
    for (int p = 0; p < npass; p++) {
        const int shift = p * SORT_RADIX_BITS;
        cuda_emu::launch(nblk, 256, 0, radix_hist_kernel<uint32_t>, (const uint32_t*)kin, n_dev, cap, shift, counts.data(), nblk);
        cuda_emu::launch(SORT_RADIX * 32 / 256, 256, 0, radix_rowscan_kernel, counts.data(), nblk, totals.data());
        cuda_emu::launch(nblk, 256, 0, radix_scatter_kernel<uint32_t>, (const uint32_t*)kin, (const uint32_t*)vin, kout, vout, n_dev, cap, shift,
                         (const uint32_t*)counts.data(), (const uint32_t*)totals.data(), nblk);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
}

extern "C" int emu_binning_depth_first(int P, const float* geo, const float* depths, const uint32_t* tiles_touched, const int32_t* radii,
                                       int tiles_x, int tiles_y, int tile_bits, int cap, long long n_dev_value,
                                       uint32_t* point_offsets, uint64_t* keys_out, uint32_t* vals_out, uint2* ranges, uint32_t* num_rendered)
{
    const int nblk_p = (P + 255) / 256;
    std::vector<uint32_t> block_sums(nblk_p + 1, 0u), order_sums(nblk_p + 1, 0u);
    for (int i = 0; i < P; i++) block_sums[i / 256] += tiles_touched[i];
    uint32_t status[8] = {0};
    cuda_emu::launch(1, 1024, 0, scan_block_sums_kernel, block_sums.data(), nblk_p, status);
    *num_rendered = status[1];
    uint32_t n_dev_store = n_dev_value < 0 ? 0u : (uint32_t)n_dev_value;
    const uint32_t* n_dev = n_dev_value < 0 ? nullptr : &n_dev_store;
    std::vector<uint32_t> ok_a(P + 1), ok_b(P + 1), ov_a(P + 1), ov_b(P + 1);
    cuda_emu::launch(nblk_p, 256, 0, order_keys_kernel, P, depths, tiles_touched, (const uint32_t*)block_sums.data(), point_offsets, ok_a.data(), ov_a.data());
    sort32(nullptr, P, 32, ok_a.data(), ov_a.data(), ok_b.data(), ov_b.data());
    cuda_emu::launch(nblk_p, 256, 0, sorted_block_sums_kernel, P, (const uint32_t*)ov_a.data(), tiles_touched, order_sums.data());
    cuda_emu::launch(1, 1024, 0, scan_block_sums_kernel, order_sums.data(), nblk_p, status);
    if (status[1] != *num_rendered) return -2;
    std::memset(ranges, 0, sizeof(uint2) * (size_t)tiles_x * tiles_y);
    if (cap <= 0) return 0;
    std::vector<uint32_t> tk_a((size_t)cap + 1, 0xFFFFFFFFu), tk_b((size_t)cap + 1, 0xFFFFFFFFu), vals_b((size_t)cap + 1, 0xFFFFFFFFu);
    const int npass = (tile_bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    const bool start_b = (npass & 1) != 0;
    cuda_emu::launch(nblk_p, 256, 0, emit_sorted_kernel, P, (const uint32_t*)ov_a.data(), geo, tiles_touched, (const uint32_t*)order_sums.data(), radii,
                     start_b ? tk_b.data() : tk_a.data(), start_b ? vals_b.data() : vals_out, tiles_x, tiles_y, n_dev, cap);
    sort32(n_dev, cap, tile_bits, tk_a.data(), vals_out, tk_b.data(), vals_b.data());
    cuda_emu::launch((cap + 255) / 256, 256, 0, finalize_bins_kernel, n_dev, cap, (const uint32_t*)tk_a.data(), (const uint32_t*)vals_out, depths, keys_out, ranges);
    return 0;
}
