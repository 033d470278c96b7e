// common.cuh -- shared declarations of the sm_100a Gaussian feature rasterizer (libsagars).
//
// Private to the library.  The public C ABI is include/sagars.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "../../include/sagars.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsagars carries sm_100a code only (build with -gencode arch=compute_100a,code=sm_100a)"
#endif

// Keeps two floats opaque to the optimiser (nvcc otherwise rematerialises the pixel coordinates from %ctaid / %tid inside the
// blend kernels' hot loops).  The CPU execution shim of the tests has no such problem and no "f" register class.
#if defined(SAGARS_CUDA_EMU)
#define SAGARS_PIN_F2(a, b) ((void)0)
#else
#define SAGARS_PIN_F2(a, b) asm volatile("" : "+f"(a), "+f"(b))
#endif

namespace sagars {

constexpr int TILE_X = SAGARS_TILE_X;
constexpr int TILE_Y = SAGARS_TILE_Y;
constexpr int TILE_PIX = TILE_X * TILE_Y;   // 256 pixels = one CTA of the blend kernels

// ---------------------------------------------------------------------------------------------
// Scratch layouts.  Every field starts on a 256-byte boundary.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct GeomView {
    float* depths;            // [P]
    float* geo;               // [P][8]  x, y, conic.x, conic.y, conic.z, opacity, accept_threshold, 0
    float* cov3D;             // [P][6]
    float* rgb;               // [P][3]
    uint8_t* clamped;         // [P][3]
    uint32_t* tiles_touched;  // [P]
    uint32_t* point_offsets;  // [P]  inclusive prefix sum of tiles_touched
    uint32_t* block_sums;     // [ceil(P/256)+1]  per-preprocess-block sums, then their exclusive scan
    uint32_t* status;         // [8]  0: prefilter violation flag, 1: num_rendered
    // depth-first binning (binning_kernels.cuh): the Gaussians ordered by depth
    uint32_t* okeys[2];       // [P] depth bits (0xFFFFFFFF when culled), ping-pong
    uint32_t* ovals[2];       // [P] Gaussian indices, ping-pong; after the sort ovals[0] is the depth order
    uint32_t* order_sums;     // [ceil(P/256)+1] block sums of tiles_touched taken in depth order, then their exclusive scan
    void* osort_temp;         // radix-sort scratch for P pairs
};

struct GeomOffsets {
    sagars_geom_layout pub;
    size_t block_sums;
    size_t okeys[2], ovals[2], order_sums, osort_temp;
};
#ifndef SAGARS_SORT_CONSTANTS
#define SAGARS_SORT_CONSTANTS
constexpr int SORT_CHUNK = 4096;          // keys per sort block
constexpr int SORT_RADIX_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
#endif
// own radix sort: counts[RADIX][nblk] + totals[RADIX]
inline size_t own_sort_temp_bytes(size_t n) {
    size_t nblk = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    return align_up((size_t)SORT_RADIX * (nblk + 1) * 4) + align_up(SORT_RADIX * 4);
}
inline GeomOffsets geom_offsets(size_t P) {
    GeomOffsets G;
    sagars_geom_layout& L = G.pub;
    size_t o = 0;
    L.depths = o;        o = align_up(o + P * 4);
    L.geo = o;           o = align_up(o + P * 32);
    L.cov3D = o;         o = align_up(o + P * 24);
    L.rgb = o;           o = align_up(o + P * 12);
    L.clamped = o;       o = align_up(o + P * 3);
    L.tiles_touched = o; o = align_up(o + P * 4);
    L.point_offsets = o; o = align_up(o + P * 4);
    size_t nblk = (P + 255) / 256 + 1;
    G.block_sums = o;    o = align_up(o + nblk * 4);
    L.status = o;        o = align_up(o + 64);
    for (int k = 0; k < 2; k++) { G.okeys[k] = o; o = align_up(o + P * 4); }
    for (int k = 0; k < 2; k++) { G.ovals[k] = o; o = align_up(o + P * 4); }
    G.order_sums = o;    o = align_up(o + nblk * 4);
    G.osort_temp = o;    o = align_up(o + own_sort_temp_bytes(P));
    L.total = o + 256;
    return G;
}
inline sagars_geom_layout geom_layout(size_t P) { return geom_offsets(P).pub; }
inline GeomView geom_view(void* base, size_t P) {
    GeomOffsets G = geom_offsets(P);
    const sagars_geom_layout& L = G.pub;
    char* b = (char*)base;
    GeomView g;
    g.depths = (float*)(b + L.depths);
    g.geo = (float*)(b + L.geo);
    g.cov3D = (float*)(b + L.cov3D);
    g.rgb = (float*)(b + L.rgb);
    g.clamped = (uint8_t*)(b + L.clamped);
    g.tiles_touched = (uint32_t*)(b + L.tiles_touched);
    g.point_offsets = (uint32_t*)(b + L.point_offsets);
    g.block_sums = (uint32_t*)(b + G.block_sums);
    g.status = (uint32_t*)(b + L.status);
    for (int k = 0; k < 2; k++) { g.okeys[k] = (uint32_t*)(b + G.okeys[k]); g.ovals[k] = (uint32_t*)(b + G.ovals[k]); }
    g.order_sums = (uint32_t*)(b + G.order_sums);
    g.osort_temp = (void*)(b + G.osort_temp);
    return g;
}

struct ImageView {
    float* final_T;        // [H*W]
    uint32_t* n_contrib;   // [H*W]
    uint2* ranges;         // [tiles]
};
inline sagars_image_layout image_layout(int W, int H) {
    sagars_image_layout L;
    size_t N = (size_t)W * H;
    size_t T = (size_t)((W + TILE_X - 1) / TILE_X) * ((H + TILE_Y - 1) / TILE_Y);
    size_t o = 0;
    L.final_T = o;   o = align_up(o + N * 4);
    L.n_contrib = o; o = align_up(o + N * 4);
    L.ranges = o;    o = align_up(o + T * 8);
    L.total = o + 256;
    return L;
}
inline ImageView image_view(void* base, int W, int H) {
    sagars_image_layout L = image_layout(W, H);
    char* b = (char*)base;
    ImageView v;
    v.final_T = (float*)(b + L.final_T);
    v.n_contrib = (uint32_t*)(b + L.n_contrib);
    v.ranges = (uint2*)(b + L.ranges);
    return v;
}

// Radix sort scratch: ping-pong key/value buffers + per-block digit counts.
#ifndef SAGARS_SORT_CONSTANTS
#define SAGARS_SORT_CONSTANTS
constexpr int SORT_CHUNK = 4096;          // keys per sort block
constexpr int SORT_RADIX_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
#endif

struct BinningView {
    uint32_t* point_list;        // [R] sorted values (final)
    uint64_t* point_list_keys;   // [R] sorted keys (final)
    uint32_t* vals_alt;          // [R]
    uint64_t* keys_alt;          // [R]
    void* sort_temp;             // sort_temp_bytes(R)
};
inline size_t sort_temp_bytes(size_t n) {
    // own sort: counts[RADIX][nblk] + totals[RADIX];  CUB (DoubleBuffer) needs O(n / tile) look-back state
    size_t own = own_sort_temp_bytes(n);
    size_t cub = align_up(n / 2 + (1u << 20));
    return own > cub ? own : cub;
}
inline sagars_binning_layout binning_layout(size_t R) {
    sagars_binning_layout L;
    size_t o = 0;
    L.point_list = o;      o = align_up(o + R * 4);
    L.point_list_keys = o; o = align_up(o + R * 8);
    L.total = o;   // (extended below; `total` is patched by binning_total)
    return L;
}
inline size_t binning_total(size_t R) {
    size_t o = binning_layout(R).total;
    o = align_up(o + R * 4);   // vals_alt
    o = align_up(o + R * 8);   // keys_alt
    o = align_up(o + sort_temp_bytes(R));
    return o + 256;
}
inline BinningView binning_view(void* base, size_t R) {
    sagars_binning_layout L = binning_layout(R);
    char* b = (char*)base;
    BinningView v;
    v.point_list = (uint32_t*)(b + L.point_list);
    v.point_list_keys = (uint64_t*)(b + L.point_list_keys);
    size_t o = L.total;
    v.vals_alt = (uint32_t*)(b + o); o = align_up(o + R * 4);
    v.keys_alt = (uint64_t*)(b + o); o = align_up(o + R * 8);
    v.sort_temp = (void*)(b + o);
    return v;
}

// Per-Gaussian accumulators of the blend backward: 8 floats per Gaussian
//   0: dL/dmean2D.x  1: dL/dmean2D.y  2: dL/dconic.x  3: dL/dconic.y  4: dL/dconic.w(zz)
//   5: dL/dopacity   6: dL/dmask      7: unused
constexpr int GG_STRIDE = 8;
inline size_t grad_scratch_bytes(size_t P) { return align_up(P * GG_STRIDE * 4) + 256; }

// ---------------------------------------------------------------------------------------------
// error handling (thread-local message, returned through sagars_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);

#define SAGARS_CUDA(call)                                                         \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return ::sagars::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// "has this kernel's attribute opt-in been done on the current device?"  One atomic word per template instantiation (bit = device
// ordinal; ordinals >= 64 simply repeat the idempotent cudaFuncSetAttribute calls): the forward thread and autograd's backward
// thread may ask at the same time.
struct DeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool need(int dev) const { return dev >= 64 || !((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
    void done(int dev) { if (dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release); }
};

// after a kernel launch: always check the launch; in debug mode also synchronise and check execution
#define SAGARS_LAUNCH_CHECK(stream, debug)                                        \
    do {                                                                          \
        ::sagars::count_launch();                                                 \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e == cudaSuccess && (debug)) _e = cudaStreamSynchronize(stream);     \
        if (_e != cudaSuccess) return ::sagars::cuda_fail(_e, "kernel", __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// stage entry points (host side), one per .cu
// ---------------------------------------------------------------------------------------------
struct Dims {
    int P, D, M, C, W, H;
    int tiles_x, tiles_y;
    float tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
};

int launch_preprocess(const sagars_forward_args& a, const Dims& d, GeomView g, cudaStream_t s, bool debug);
int launch_scan_block_sums(const Dims& d, GeomView g, cudaStream_t s, bool debug);
// n_dev / cap: element count in device memory (or nullptr: cap is the count) and the capacity of the arrays
int launch_duplicate(const Dims& d, GeomView g, const int32_t* radii, uint64_t* keys, uint32_t* vals,
                     const uint32_t* n_dev, int cap, cudaStream_t s, bool debug);
int launch_sort_pairs(const uint32_t* n_dev, int cap, int end_bit, uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b,
                      uint32_t* vals_b, void* temp, size_t temp_bytes, bool use_cub, bool* result_in_a,
                      cudaStream_t s, bool debug);
int sort_num_passes(int end_bit);
int launch_tile_ranges(const uint32_t* n_dev, int cap, int num_tiles, const uint64_t* keys, uint2* ranges, cudaStream_t s, bool debug);
// depth-first binning (SAGARS default): Gaussians ordered by depth -> instances emitted in that order -> stable sort on the tile bits
int launch_depth_order(const Dims& d, GeomView g, cudaStream_t s, bool debug);            // needs only P: may run before R is known
int launch_emit_sorted(const Dims& d, GeomView g, const int32_t* radii, uint32_t* tkeys, uint32_t* vals, const uint32_t* n_dev, int cap,
                       cudaStream_t s, bool debug);
int launch_sort_pairs32(const uint32_t* n_dev, int cap, int end_bit, uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                        void* temp, cudaStream_t s, bool debug);                           // input in A (even pass count) or B (odd); result in A
int launch_finalize_bins(const uint32_t* n_dev, int cap, int num_tiles, const uint32_t* tkeys, const uint32_t* point_list, const float* depths,
                         uint64_t* keys, uint2* ranges, cudaStream_t s, bool debug);
// tile_sort.cu (SAGARS_FLAG_TILE_SORT): count -> scan -> scatter (launch_tile_bin), then one CTA per tile sorts its segment
int launch_tile_bin(const Dims& d, GeomView g, const int32_t* radii, uint64_t* pairs, uint2* ranges, uint32_t* queue,
                    const uint32_t* n_dev, int cap, cudaStream_t s, bool debug);
int launch_tile_sort(int num_tiles, uint2* ranges, uint64_t* pairs, uint32_t* point_list, uint64_t* keys, uint32_t* queue,
                     const uint32_t* n_dev, int cap, cudaStream_t s, bool debug);
size_t tile_sort_queue_bytes(int num_tiles);   // scratch of the long-segment queue (taken from the sort scratch)
int launch_smooth_forward(int P, int C, int Ks, const float* F, const long long* idx, int normalize_out, float* out,
                          float* mean_norm, cudaStream_t s);
int launch_smooth_backward(int P, int C, int Ks, const float* F, const long long* idx, int normalize_out,
                           const float* mean_norm, const float* out, const float* dL_dout, float* dL_dn, float* dL_dF,
                           cudaStream_t s);
int launch_sample_rays_forward(int C, int H, int W, int h, int w, const float* img, const long long* rays, int S, float* out,
                               float* norm_sum, cudaStream_t s);
int launch_sample_rays_backward(int C, int H, int W, int h, int w, const float* img, const long long* rays, int S, const float* g_out,
                                const float* g_norm, float* grad_img, cudaStream_t s);
size_t knn_temp_bytes(size_t n);
int launch_knn(int n, const float* points, int nq, const float* queries, int K, bool exclude_self, long long* idx_out,
               float* dist_out, float* mean_out, void* temp, cudaStream_t s);
int launch_render_forward_warp(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                               const uint32_t* point_list, cudaStream_t s, bool debug);
int launch_render_forward(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                          const uint32_t* point_list, cudaStream_t s, bool debug);
int launch_render_forward_tc(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                             const uint32_t* point_list, cudaStream_t s, bool debug);
int launch_render_backward(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                           const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug);
int launch_render_backward_warp(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                                const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug);
int launch_render_backward_tc(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                              const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug);
int launch_render_backward_mma(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                               const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug);
int launch_geom_backward(const sagars_backward_args& a, const Dims& d, GeomView g, const float* ggrad,
                         cudaStream_t s, bool debug);
int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                        cudaStream_t s);

}  // namespace sagars
