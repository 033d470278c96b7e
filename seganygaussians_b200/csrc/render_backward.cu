// render_backward.cu -- per-tile back-to-front gradient of the alpha compositing.
//
// Semantics: CF cuda_rasterizer/backward.cu:399-559 (DEPTH backward.cu:400-564 adds dL_dmask),
// SURVEY.md Appendix A.13-A.17.  The reference gives every (pixel, Gaussian) pair C+6 scalar global
// atomicAdds; all 256 threads of a CTA hit the same C+6 addresses.  This kernel is organised
// differently -- two phases per batch of NB instances, both inside one CTA per 16x16 tile:
//
//   phase A (thread = pixel, exactly the reference's traversal): recompute alpha, undo T, and reduce
//     the C-channel work of a pair to ONE dot product s = f_j . g_p.  Because dL/dalpha is linear in
//     the upstream gradient, the reference's per-channel recurrence accum_rec[ch] collapses to a
//     scalar recurrence on a = accum_rec . g_p (same association order, Appendix D).  A (warp, splat)
//     pair in which no pixel can pass the reference's tests (conservative lower bound on `power`, see
//     accept_threshold) is skipped with one warp vote before expf.  Each blended pair appends (w = alpha*T, q = G*dL/dalpha, pixel) to a
//     compact per-splat list in shared memory (warp ballot + one shared atomic per (warp, splat)).
//   phase B (warps take splats from a shared counter; lane = (list slot, channel quad)): every per-Gaussian gradient is a sparse
//     product over the tile's pixels,
//         dL/dcolour[j][:] = sum_p w[j][p] * g[p][:]          (g tile resident in smem, float4 reads)
//         moments[j][:]    = sum_p q[j][p] * (1, x, y, x^2, xy, y^2)(p)   (tile-centred pixel coordinates)
//     accumulated in registers by walking the compact list; the six moments give dL/dopacity, dL/dmean2D
//     and dL/dconic in closed form.  No per-pair atomics, no cross-lane reduction per pair.
//   One vectorised `red.global.add.v4.f32` per (tile, Gaussian, channel quad) and six scalar reds per
//   (tile, Gaussian) then replace the reference's (C+6) x (blended pixels) atomics.
//
// Instance records are staged with cp.async (double buffered), feature rows with cp.async behind phase B.
#include "common.cuh"
#include "cp_async.cuh"
#include "render_backward_kernels.cuh"

namespace sagars {

template <int NQ, bool VEC, bool MD, bool COLOR>
static int launch_bwd_t(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                        const uint32_t* point_list, const float* features, float* ggrad, cudaStream_t s, bool debug)
{
    auto kern = render_backward_kernel<NQ, VEC, MD, COLOR>;
    const size_t smem = sizeof(BwdSmem<NQ>);
    {   // opt in to the dynamic shared-memory size once per device (not on every launch: the call takes the context lock)
        static DeviceOnce once;
        int dev = 0;
        SAGARS_CUDA(cudaGetDevice(&dev));
        if (once.need(dev)) {
            SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            once.done(dev);
        }
    }
    dim3 grid(d.tiles_x, d.tiles_y);
    kern<<<grid, TILE_PIX, smem, s>>>(im.ranges, point_list, d.W, d.H, d.C, a.background, g.geo, features,
                                      im.final_T, im.n_contrib, a.dL_dout_color, a.dL_dout_mask, ggrad,
                                      a.dL_dcolors);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_render_backward(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                           const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug)
{
    const bool md = (a.flags & SAGARS_FLAG_MASK_DEPTH) != 0;
    const bool mask_only = (a.flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    if (mask_only) return launch_bwd_t<1, false, true, false>(a, d, g, im, point_list, features, ggrad, s, debug);
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define SAGARS_BWD_CASE(NQ_)                                                                                     \
    if (nq <= NQ_) {                                                                                             \
        if (md) return launch_bwd_t<NQ_, false, true, true>(a, d, g, im, point_list, features, ggrad, s, debug);  \
        return vec ? launch_bwd_t<NQ_, true, false, true>(a, d, g, im, point_list, features, ggrad, s, debug)     \
                   : launch_bwd_t<NQ_, false, false, true>(a, d, g, im, point_list, features, ggrad, s, debug);   \
    }
    SAGARS_BWD_CASE(1)
    SAGARS_BWD_CASE(2)
    SAGARS_BWD_CASE(4)
    SAGARS_BWD_CASE(8)
    SAGARS_BWD_CASE(16)
#undef SAGARS_BWD_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
