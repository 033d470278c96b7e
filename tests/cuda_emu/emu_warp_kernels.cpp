// Host harness: the DEFAULT K = 32 blend kernels (one warp per 8x4 pixel block; seganygaussians_b200/csrc/
// render_forward_warp_kernels.cuh, render_backward_warp_kernels.cuh) under the execution shim.  TEST INFRASTRUCTURE ONLY.
#define SAGARS_CUDA_EMU 1
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "render_forward_warp_kernels.cuh"
#include "render_backward_warp_kernels.cuh"
#include "render_backward_kernels.cuh"
#include "render_backward_mma_kernels.cuh"
#include "render_forward_tc_kernels.cuh"
#include "render_backward_tc_kernels.cuh"
#include "math.cuh"

using namespace sagars;

extern "C" void emu_make_geo(int P, const float* means2D, const float* conic_opacity, float* geo)
{
    for (int i = 0; i < P; i++) {
        float* g = geo + 8 * (size_t)i;
        g[0] = means2D[2 * i]; g[1] = means2D[2 * i + 1];
        g[2] = conic_opacity[4 * i]; g[3] = conic_opacity[4 * i + 1]; g[4] = conic_opacity[4 * i + 2]; g[5] = conic_opacity[4 * i + 3];
        g[6] = accept_threshold(g[5]); g[7] = 0.f;
    }
}

extern "C" int emu_forward_warp(int W, int H, int K, const uint2* ranges, const uint32_t* point_list, const float* geo,
                                const float* features, const float* bg, float* final_T, uint32_t* n_contrib, float* out_color)
{
    const unsigned tx = (W + TILE_X - 1) / TILE_X, ty = (H + TILE_Y - 1) / TILE_Y;
    const bool vec = (K % 4) == 0;
    const int nq = (K + 3) / 4;
#define GO(NQ_, VEC_) cuda_emu::launch2d(2 * tx, 4 * ty, 32, sizeof(FwSmem<NQ_>), render_forward_warp_kernel<NQ_, VEC_>, ranges, point_list, \
                                         W, H, K, geo, features, bg, final_T, n_contrib, out_color)
    // the launcher's dispatch (render_forward_warp.cu): smallest NQ that holds the channels, float4 or scalar feature loads
#define CASE(NQ_) if (nq <= NQ_) { if (vec) { GO(NQ_, true); } else { GO(NQ_, false); } return 0; }
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16)
#undef CASE
#undef GO
    return -1;
}

extern "C" int emu_backward_warp(int md, int W, int H, int K, const uint2* ranges, const uint32_t* point_list, const float* bg,
                                 const float* geo, const float* features, const float* final_T, const uint32_t* n_contrib,
                                 const float* dL_dpix, const float* dL_dout_mask, float* ggrad, float* dL_dcolors)
{
    const unsigned tx = (W + TILE_X - 1) / TILE_X, ty = (H + TILE_Y - 1) / TILE_Y;
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define GO(NQ_, VEC_, MD_) cuda_emu::launch2d(2 * tx, 4 * ty, 32, sizeof(BwSmem<NQ_>), render_backward_warp_kernel<NQ_, VEC_, MD_, true>, ranges, \
                                              point_list, W, H, K, bg, geo, features, final_T, n_contrib, dL_dpix, dL_dout_mask, ggrad, dL_dcolors)
    // the launcher's dispatch (render_backward_warp.cu): smallest NQ that holds the channels, MD / VEC / scalar instance of it
#define CASE(NQ_) if (nq <= NQ_) { if (md) { GO(NQ_, false, true); } else if (vec) { GO(NQ_, true, false); } else { GO(NQ_, false, false); } return 0; }
    CASE(1) CASE(2) CASE(4) CASE(8) CASE(16)
#undef CASE
#undef GO
    return -1;
}

// the two CTA-per-tile backward kernels: kind 1 = fp32 SIMT (render_backward_kernels.cuh), kind 2 = mma.sync tile kernel
extern "C" int emu_backward_tile(int kind, int md, int W, int H, int K, const uint2* ranges, const uint32_t* point_list, const float* bg,
                                 const float* geo, const float* features, const float* final_T, const uint32_t* n_contrib,
                                 const float* dL_dpix, const float* dL_dout_mask, float* ggrad, float* dL_dcolors)
{
    cuda_emu::thread_exit_hook = emu_async::flush_thread;
    const unsigned tx = (W + TILE_X - 1) / TILE_X, ty = (H + TILE_Y - 1) / TILE_Y;
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define GO1(NQ_, VEC_, MD_) cuda_emu::launch2d(tx, ty, TILE_PIX, sizeof(BwdSmem<NQ_>), render_backward_kernel<NQ_, VEC_, MD_, true>, ranges, point_list, \
                                               W, H, K, bg, geo, features, final_T, n_contrib, dL_dpix, dL_dout_mask, ggrad, dL_dcolors)
#define GO2(NQ_, VEC_, MD_) cuda_emu::launch2d(tx, ty, TILE_PIX, sizeof(BmSmem<NQ_>), render_backward_mma_kernel<NQ_, VEC_, MD_, true>, ranges, point_list, \
                                               W, H, K, bg, geo, features, final_T, n_contrib, dL_dpix, dL_dout_mask, ggrad, dL_dcolors)
    if (kind == 1) {
        if (md) { if (nq <= 1) { GO1(1, false, true); return 0; } return -1; }
        if (nq <= 1 && !vec) { GO1(1, false, false); return 0; }
        if (nq <= 8 && vec) { GO1(8, true, false); return 0; }
    } else if (kind == 2) {
        if (md) { if (nq <= 1) { GO2(1, false, true); return 0; } return -1; }
        if (nq <= 1 && !vec) { GO2(1, false, false); return 0; }
        if (nq <= 8 && vec) { GO2(8, true, false); return 0; }
    }
#undef GO1
#undef GO2
    return -1;
}

// the tcgen05 / TMEM tile forward (K = 32 only)
extern "C" int emu_forward_tc(int W, int H, const uint2* ranges, const uint32_t* point_list, const float* geo, const float* features,
                              const float* bg, float* final_T, uint32_t* n_contrib, float* out_color)
{
    cuda_emu::thread_exit_hook = emu_async::flush_thread;
    const unsigned tx = (W + TILE_X - 1) / TILE_X, ty = (H + TILE_Y - 1) / TILE_Y;
    cuda_emu::launch2d(tx, ty, TILE_PIX, sizeof(FwdTcSmem) + 1024, render_forward_tc_kernel, ranges, point_list, W, H, geo, features, bg,
                       final_T, n_contrib, out_color);
    return 0;
}

// the tcgen05 / TMEM backward (C = 32 precomputed colours): one CTA per 16x8 pixel group
extern "C" int emu_backward_tc(int W, int H, const uint2* ranges, const uint32_t* point_list, const float* bg, const float* geo,
                               const float* features, const float* final_T, const uint32_t* n_contrib, const float* dL_dpix,
                               float* ggrad, float* dL_dcolors)
{
    cuda_emu::thread_exit_hook = emu_async::flush_thread;
    const unsigned tx = (W + TILE_X - 1) / TILE_X, ty = (H + TILE_Y - 1) / TILE_Y;
    cuda_emu::launch2d(tx, 2 * ty, BT_THREADS, sizeof(BtSmem) + 1024, render_backward_tc_kernel, ranges, point_list, W, H, bg, geo, features,
                       final_T, n_contrib, dL_dpix, ggrad, dL_dcolors);
    return 0;
}
