// geom_backward.cu -- per-Gaussian tail of the backward pass, one fused kernel.
//
// The reference runs two kernels here (computeCov2DCUDA, CF cuda_rasterizer/backward.cu:144-274, then
// preprocessCUDA, backward.cu:346-396) over nine zero-initialised gradient tensors.  This library
// fuses them: one thread per Gaussian reads the eight blend-stage accumulators of `ggrad`
// (dL/dmean2D, dL/dconic, dL/dopacity, dL/dmask), pushes them through conic -> cov2D -> cov3D ->
// scale/rotation and mean2D -> mean3D (+ SH -> mean3D), and WRITES every output row exactly once
// (zeros for culled Gaussians), so no output tensor needs a memset.  Formulas: SURVEY.md Appendix
// A.18-A.24.
#include "common.cuh"
#include "math.cuh"
#include "geom_backward_kernels.cuh"

namespace sagars {

int launch_geom_backward(const sagars_backward_args& a, const Dims& d, GeomView g, const float* ggrad,
                         cudaStream_t s, bool debug)
{
    const float* cov3D = a.cov3D_precomp != nullptr ? a.cov3D_precomp : g.cov3D;
    geom_backward_kernel<<<(d.P + 255) / 256, 256, 0, s>>>(
        d.P, d.D, d.M, a.means3D, a.radii, cov3D, a.shs, g.clamped, a.scales, a.rotations, d.scale_modifier,
        a.viewmatrix, a.projmatrix, a.cam_pos, d.focal_x, d.focal_y, d.tan_fovx, d.tan_fovy, ggrad,
        a.dL_dcolors, a.dL_dmeans2D, a.dL_dopacity, a.dL_dmask, a.dL_dmeans3D, a.dL_dcov3D, a.dL_dsh,
        a.dL_dscales, a.dL_drotations);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

}  // namespace sagars
