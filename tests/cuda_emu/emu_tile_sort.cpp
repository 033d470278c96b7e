// Host harness: the kernels of seganygaussians_b200/csrc/tile_sort_kernels.cuh compiled against the execution shim and run in
// the order launch_tile_bin / launch_tile_sort (tile_sort.cu) queue them.  TEST INFRASTRUCTURE ONLY.
#include "../../include/sagars.h"
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "tile_sort_kernels.cuh"

using namespace sagars;

// n_dev_value < 0: exact layout (the kernels get n_dev == nullptr); otherwise the device-side instance count the
// speculative layout compares with `cap`.  small_cap / large_cap are fixed by the kernels (TSORT_SMALL / TSORT_LARGE).
extern "C" int emu_tile_binning(int P, const float* geo, const float* depths, const uint32_t* tiles_touched,
                                const uint32_t* block_excl, const int32_t* radii, int tiles_x, int tiles_y, int cap,
                                long long n_dev_value, int big_grid,
                                uint32_t* point_offsets, uint2* ranges, uint64_t* pairs, uint32_t* point_list, uint64_t* keys,
                                uint32_t* queue)
{
    const int num_tiles = tiles_x * tiles_y;
    const int nblk = (P + 255) / 256;
    uint32_t n_dev_store = n_dev_value < 0 ? 0u : (uint32_t)n_dev_value;
    const uint32_t* n_dev = n_dev_value < 0 ? nullptr : &n_dev_store;
    std::memset(ranges, 0, sizeof(uint2) * (size_t)num_tiles);
    cuda_emu::launch(nblk, 256, 0, tile_count_kernel, P, geo, tiles_touched, block_excl, radii, point_offsets, ranges, tiles_x, tiles_y);
    cuda_emu::launch(1, 1024, 0, tile_scan_kernel, ranges, num_tiles, queue);
    cuda_emu::launch(nblk, 256, 0, tile_scatter_kernel, P, geo, depths, tiles_touched, radii, ranges, pairs, tiles_x, tiles_y, n_dev, cap);
    cuda_emu::launch(num_tiles, 256, 0, tile_sort_small_kernel, ranges, (const uint64_t*)pairs, point_list, keys, queue, n_dev, cap);
    cuda_emu::launch(big_grid, 1024, TSORT_LARGE * 8, tile_sort_big_kernel, (const uint2*)ranges, pairs, point_list, keys,
                     (const uint32_t*)queue, n_dev, cap);
    return (int)queue[0];
}
