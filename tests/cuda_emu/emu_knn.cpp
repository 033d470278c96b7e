// Host harness: the exact-KNN kernels (seganygaussians_b200/csrc/knn_kernels.cuh; Morton order via the library's radix sort,
// two-level boxes, pruned search) under the execution shim, with the orchestration of launch_knn (knn.cu) and the ping-pong of
// launch_sort_pairs (binning.cu) restated.  TEST INFRASTRUCTURE ONLY.
#define SAGARS_CUDA_EMU 1
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "knn_kernels.cuh"
#include "binning_kernels.cuh"
#include <algorithm>
#include <vector>

using namespace sagars;

static void sort_pairs(int n, int end_bit, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b)
{
    const int nblk = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    std::vector<uint32_t> counts((size_t)SORT_RADIX * (nblk + 1)), totals(SORT_RADIX);
    const int npass = (end_bit + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    uint64_t* kin = (npass & 1) ? keys_b : keys_a;
    uint32_t* vin = (npass & 1) ? vals_b : vals_a;
    uint64_t* kout = (npass & 1) ? keys_a : keys_b;
    uint32_t* vout = (npass & 1) ? vals_a : vals_b;
    for (int p = 0; p < npass; p++) {
        const int shift = p * SORT_RADIX_BITS;
        cuda_emu::launch(nblk, 256, 0, radix_hist_kernel<uint64_t>, (const uint64_t*)kin, (const uint32_t*)nullptr, n, shift, counts.data(), nblk);
        cuda_emu::launch(SORT_RADIX * 32 / 256, 256, 0, radix_rowscan_kernel, counts.data(), nblk, totals.data());
        cuda_emu::launch(nblk, 256, 0, radix_scatter_kernel<uint64_t>, (const uint64_t*)kin, (const uint32_t*)vin, kout, vout, (const uint32_t*)nullptr, n,
                         shift, (const uint32_t*)counts.data(), (const uint32_t*)totals.data(), nblk);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
}

template <int K>
static void search(bool self, bool excl, int n, int nq, const float* queries, const KnnTemp& t, int nb1, int nb2, int k_out,
                   long long* idx_out, float* dist_out, float* mean_out)
{
    const unsigned blocks = (nq + 127) / 128;
    if (self && excl) cuda_emu::launch(blocks, 128, 0, knn_search_kernel<K, true, true>, n, nq, queries, (const float4*)t.sorted_pts, (const KnnBox*)t.box1, nb1, (const KnnBox*)t.box2, nb2, k_out, idx_out, dist_out, mean_out);
    else if (self) cuda_emu::launch(blocks, 128, 0, knn_search_kernel<K, true, false>, n, nq, queries, (const float4*)t.sorted_pts, (const KnnBox*)t.box1, nb1, (const KnnBox*)t.box2, nb2, k_out, idx_out, dist_out, mean_out);
    else cuda_emu::launch(blocks, 128, 0, knn_search_kernel<K, false, false>, n, nq, queries, (const float4*)t.sorted_pts, (const KnnBox*)t.box1, nb1, (const KnnBox*)t.box2, nb2, k_out, idx_out, dist_out, mean_out);
}

extern "C" int emu_knn(int n, const float* points, int nq, const float* queries, int K, int exclude_self, long long* idx_out,
                       float* dist_out, float* mean_out)
{
    std::vector<char> temp(knn_temp_layout((size_t)n, nullptr, nullptr) + 512);
    char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(temp.data()) + 255) & ~uintptr_t(255));
    KnnTemp t;
    knn_temp_layout((size_t)n, &t, base);
    const int nb1 = (n + KNN_L1 - 1) / KNN_L1, nb2 = (nb1 + KNN_FAN - 1) / KNN_FAN;
    if (queries == nullptr) nq = n;
    cuda_emu::launch(1, 32, 0, knn_bbox_init_kernel, t.bbox);
    cuda_emu::launch(std::min((n + 255) / 256, 148 * 8), 256, 0, knn_bbox_kernel, n, points, t.bbox);
    constexpr int bits = 30;
    const int npass = (bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
    const bool start_alt = (npass & 1) != 0;
    cuda_emu::launch((n + 255) / 256, 256, 0, knn_morton_kernel, n, points, (const int*)t.bbox, start_alt ? t.keys_b : t.keys_a,
                     start_alt ? t.vals_b : t.vals_a);
    sort_pairs(n, bits, t.keys_a, t.vals_a, t.keys_b, t.vals_b);
    cuda_emu::launch((n + 255) / 256, 256, 0, knn_gather_kernel, n, (const uint32_t*)t.vals_a, points, t.sorted_pts);
    cuda_emu::launch((nb1 * 32 + 255) / 256, 256, 0, knn_box1_kernel, n, (const float4*)t.sorted_pts, t.box1, nb1);
    cuda_emu::launch((nb2 + 127) / 128, 128, 0, knn_box2_kernel, (const KnnBox*)t.box1, nb1, t.box2, nb2);
    const bool self = queries == nullptr;
    if (K <= 4) search<4>(self, exclude_self != 0, n, nq, queries, t, nb1, nb2, K, idx_out, dist_out, mean_out);
    else if (K <= 8) search<8>(self, exclude_self != 0, n, nq, queries, t, nb1, nb2, K, idx_out, dist_out, mean_out);
    else if (K <= 16) search<16>(self, exclude_self != 0, n, nq, queries, t, nb1, nb2, K, idx_out, dist_out, mean_out);
    else search<32>(self, exclude_self != 0, n, nq, queries, t, nb1, nb2, K, idx_out, dist_out, mean_out);
    return 0;
}
