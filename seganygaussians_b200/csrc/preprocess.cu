// preprocess.cu -- per-Gaussian projection stage (forward) and the frustum test.
//
// Behaviour follows the reference's per-Gaussian stage (CF cuda_rasterizer/forward.cu:158-259 with
// helpers forward.cu:23-155 and auxiliary.h:41-164).  The integer
// results (radius, tile rectangle, tiles_touched) feed the tile binning and must be bit-exact, so
// the fp32 expression trees below keep the reference's association order, its explicit zero terms in
// the 3x3 products and its fp64 NDC->pixel conversion; everything else is this library's own
// structure: one packed 32-byte record per Gaussian for the blend kernels, the per-block partial
// sums of tiles_touched produced in the same pass (so the scan needs no extra read of the array),
// and a status word instead of the reference's __trap() for the `prefiltered` contract.
#include "common.cuh"
#include "math.cuh"
#include "preprocess_kernels.cuh"

namespace sagars {

int launch_preprocess(const sagars_forward_args& a, const Dims& d, GeomView g, cudaStream_t s, bool debug)
{
    const int blocks = (d.P + 255) / 256;
    preprocess_kernel<<<blocks, 256, 0, s>>>(
        d.P, d.D, d.M, d.C, a.means3D, a.scales, d.scale_modifier, a.rotations, a.opacities, a.shs,
        a.cov3D_precomp, a.colors_precomp != nullptr || (a.flags & SAGARS_FLAG_MASK_ONLY),
        a.viewmatrix, a.projmatrix, a.cam_pos, d.W, d.H, d.tiles_x, d.tiles_y,
        d.tan_fovx, d.tan_fovy, d.focal_x, d.focal_y, a.radii, g,
        (a.flags & SAGARS_FLAG_PREFILTERED) ? 1u : 0u);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* view, const float* proj, uint8_t* present,
                        cudaStream_t s)
{
    (void)proj;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
    SAGARS_LAUNCH_CHECK(s, false);
    return SAGARS_OK;
}

}  // namespace sagars
