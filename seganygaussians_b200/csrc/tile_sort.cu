// tile_sort.cu -- tile binning WITHOUT a global sort (SAGARS_FLAG_TILE_SORT): count -> scan -> scatter -> per-tile sort.
//
// Same contract as binning.cu (CF cuda_rasterizer/rasterizer_impl.cu:70-138,277-317): the final
// point_list / point_list_keys / ranges are bit-identical to a STABLE ascending sort of the emitted (tile << 32 | depth bits)
// keys.  What is exploited: the high half of the key is the tile id, and how many instances every tile receives is known
// before a single key is written.  So instead of 6 global radix passes over all R instances (6 x 24 B/instance of HBM
// traffic, 18 launches):
//   1. tile_count_kernel   : per Gaussian, point_offsets (as duplicate_kernel) and one atomic per touched tile into
//                            ranges[tile].y                                                            (R atomics, no writes)
//   2. tile_scan_kernel    : exclusive scan of the T tile counts (one CTA) -> ranges[tile] = (start, start)
//   3. tile_scatter_kernel : per Gaussian, per touched tile: slot = atomicAdd(&ranges[tile].y, 1); the pair
//                            (depth bits << 32 | Gaussian id) goes to slot.  Afterwards ranges[tile] = [start, end) -- the
//                            reference's tile ranges fall out for free -- and every tile's segment holds its instances in
//                            arbitrary order                                                           (8 B/instance written)
//   4. tile_sort_{small,big}_kernel: one CTA per tile sorts its segment by the 64-bit pair.  Within a tile every Gaussian occurs at
//                            most once and the reference's emission order is ascending Gaussian id, so ascending
//                            (depth bits, id) IS the stable order.  Segments of <= 1024 pairs are sorted in 8 KB of shared
//                            memory by the tile's own CTA; longer ones are queued for a small persistent grid (<= 8192 pairs:
//                            64 KB of shared memory; longer: chunks of 8192 through that buffer, only the network's long-span
//                            stages in global memory); the sorted ids and the re-assembled keys are written once
//                                                                                              (8 B read + 12 B written)
// The sort network is the normalised bitonic network (every comparator ascending: a "flip" stage with partner i ^ (k - 1),
// then half-cleaners with partner i ^ j).  With all comparators ascending a segment of any length n sorts as if padded with
// +inf to the next power of two: comparators whose upper index is >= n are skipped (sort_network.cuh; tests/test_tile_sort_network.py compiles
// that header for the host and checks the schedule for every n <= 1100 and random larger n).
#include "common.cuh"
#include "math.cuh"
#include "tile_sort_kernels.cuh"

namespace sagars {

size_t tile_sort_queue_bytes(int num_tiles) { return ((size_t)num_tiles + 1) * sizeof(uint32_t); }

// memset(ranges) + count + scan + scatter.  `pairs` holds cap u64; `queue`: tile_sort_queue_bytes(num_tiles) bytes.
int launch_tile_bin(const Dims& d, GeomView g, const int32_t* radii, uint64_t* pairs, uint2* ranges, uint32_t* queue,
                    const uint32_t* n_dev, int cap, cudaStream_t s, bool debug)
{
    const int num_tiles = d.tiles_x * d.tiles_y;
    const int nblk = (d.P + 255) / 256;
    SAGARS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), s));
    tile_count_kernel<<<nblk, 256, 0, s>>>(d.P, g.geo, g.tiles_touched, g.block_sums, radii, g.point_offsets, ranges,
                                           d.tiles_x, d.tiles_y);
    SAGARS_LAUNCH_CHECK(s, debug);
    tile_scan_kernel<<<1, 1024, 0, s>>>(ranges, num_tiles, queue);
    SAGARS_LAUNCH_CHECK(s, debug);
    tile_scatter_kernel<<<nblk, 256, 0, s>>>(d.P, g.geo, g.depths, g.tiles_touched, radii, ranges, pairs, d.tiles_x, d.tiles_y,
                                             n_dev, cap);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

// queue[0] was zeroed by launch_tile_bin's scan kernel
int launch_tile_sort(int num_tiles, uint2* ranges, uint64_t* pairs, uint32_t* point_list, uint64_t* keys, uint32_t* queue,
                     const uint32_t* n_dev, int cap, cudaStream_t s, bool debug)
{
    {   // opt in to 64 KB of dynamic shared memory once per device
        static DeviceOnce once;
        int dev = 0;
        SAGARS_CUDA(cudaGetDevice(&dev));
        if (once.need(dev)) {
            SAGARS_CUDA(cudaFuncSetAttribute(tile_sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TSORT_LARGE * 8)));
            once.done(dev);
        }
    }
    tile_sort_small_kernel<<<num_tiles, 256, 0, s>>>(ranges, pairs, point_list, keys, queue, n_dev, cap);
    SAGARS_LAUNCH_CHECK(s, debug);
    // the host cannot know whether any tile holds more than 1024 instances without a read-back: a small persistent grid is
    // always queued and usually finds the queue empty
    const int grid = num_tiles < TSORT_BIG_CTAS ? num_tiles : TSORT_BIG_CTAS;
    tile_sort_big_kernel<<<grid, 1024, TSORT_LARGE * 8, s>>>(ranges, pairs, point_list, keys, queue, n_dev, cap);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

}  // namespace sagars
