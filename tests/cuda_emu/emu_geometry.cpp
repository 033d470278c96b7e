// Host harness: the per-Gaussian kernels (seganygaussians_b200/csrc/preprocess_kernels.cuh, geom_backward_kernels.cuh) under the
// execution shim.  TEST INFRASTRUCTURE ONLY.
#define SAGARS_CUDA_EMU 1
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "preprocess_kernels.cuh"
#include "geom_backward_kernels.cuh"
#include <vector>

using namespace sagars;

extern "C" size_t emu_geom_bytes(int P) { return geom_layout((size_t)P).total; }
extern "C" void emu_geom_offsets(int P, size_t* out)   // depths, geo, cov3D, rgb, clamped, tiles_touched, block_sums, status
{
    GeomOffsets G = geom_offsets((size_t)P);
    out[0] = G.pub.depths; out[1] = G.pub.geo; out[2] = G.pub.cov3D; out[3] = G.pub.rgb; out[4] = G.pub.clamped;
    out[5] = G.pub.tiles_touched; out[6] = G.block_sums; out[7] = G.pub.status;
}

extern "C" void emu_preprocess(int P, int D, int M, int C, const float* means3D, const float* scales, float scale_modifier,
                               const float* rotations, const float* opacities, const float* shs, const float* cov3D_precomp,
                               int have_colors, const float* view, const float* proj, const float* campos, int W, int H,
                               float tan_fovx, float tan_fovy, int32_t* radii, void* geom_buffer, unsigned prefiltered)
{
    GeomView g = geom_view(geom_buffer, (size_t)P);
    std::memset(g.status, 0, 64);
    const int tx = (W + TILE_X - 1) / TILE_X, ty = (H + TILE_Y - 1) / TILE_Y;
    const float fy = H / (2.0f * tan_fovy), fx = W / (2.0f * tan_fovx);
    cuda_emu::launch((P + 255) / 256, 256, 0, preprocess_kernel, P, D, M, C, means3D, scales, scale_modifier, rotations, opacities, shs,
                     cov3D_precomp, have_colors != 0, view, proj, campos, W, H, tx, ty, tan_fovx, tan_fovy, fx, fy, radii, g, prefiltered);
}

extern "C" void emu_geom_backward(int P, int D, int M, const float* means3D, const int32_t* radii, const float* cov3Ds, const float* shs,
                                  const uint8_t* clamped, const float* scales, const float* rotations, float scale_modifier,
                                  const float* view, const float* proj, const float* campos, int W, int H, float tan_fovx, float tan_fovy,
                                  const float* ggrad, const float* dL_dcolor, float* dL_dmeans2D, float* dL_dopacity, float* dL_dmask,
                                  float* dL_dmeans3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots)
{
    const float fy = H / (2.0f * tan_fovy), fx = W / (2.0f * tan_fovx);
    cuda_emu::launch((P + 255) / 256, 256, 0, geom_backward_kernel, P, D, M, means3D, radii, cov3Ds, shs, clamped, scales, rotations,
                     scale_modifier, view, proj, campos, fx, fy, tan_fovx, tan_fovy, ggrad, dL_dcolor, dL_dmeans2D, dL_dopacity, dL_dmask,
                     dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drots);
}
