// render_backward_mma.cu -- per-tile back-to-front gradient of the alpha compositing; the per-Gaussian reductions over
// the tile's pixels are warp-level tensor-core GEMMs on COMPACTED rows.
//
// Semantics: CF cuda_rasterizer/backward.cu:399-559 (DEPTH backward.cu:400-564 adds dL_dmask), SURVEY.md Appendix
// A.13-A.17 / D.  One CTA per 16x16 tile, one thread per pixel, splats taken back to front in staged batches of 64.
//
//   phase A: first, lane = splat: 32 splats at a time are tested against the warp's 8x4 pixel block (exact minimum of
//     the splat's quadratic form over the block vs. the conservative accept threshold of math.cuh) -- about 70 % of the
//     (warp, splat) pairs end here, 32 per instruction.  Then thread = pixel over the surviving splats (the
//     reference's traversal): accepted pairs recompute alpha, undo T and need ONE dot product s = f_j . g_p because the
//     reference's per-channel recurrence accum_rec[ch] collapses to a scalar recurrence on a = accum_rec . g_p.  A
//     (warp, splat) instance with at least one accepted pixel appends ONE row to the warp's private tiles:
//     W[row][p] = alpha*T (weight of dL/dcolour) and Q[row][p] = G * dL/dalpha (weight of every geometric gradient),
//     zero for the pixels that did not blend.
//   phase B (per warp, whenever 8 rows are waiting; no CTA barrier, no cross-warp traffic):
//         dL/dcolour^T[:, row] = G^T (C x 32 pixels) * W^T (32 x 8 rows)           mma.sync m16n8k8 TF32, 3xTF32 split
//         moments[row][:]      = Q (8 x 32) * (1, x, y, x^2, xy, y^2)(pixel)       tile-centred basis, exact in tf32
//     accumulators in registers, results leave as fire-and-forget `red.global.add.f32`: C + 6 per (warp, splat) instance
//     instead of the reference's C + 6 atomics per blended (pixel, splat) pair.  dL/dopacity, dL/dmean2D, dL/dconic follow
//     in closed form from the six moments.
//
// The only CTA-wide synchronisation left is ONE barrier per staged batch (cp.async double buffer of records + feature
// rows).  Shared-memory rows are rotation swizzled instead of padded (g rows by quad, operand rows by 4 columns per
// row), which keeps the thread-per-pixel float4 reads and the mma fragment gathers (nearly) conflict-free
// and three CTAs per SM.
#include "common.cuh"
#include "cp_async.cuh"
#include "candidate.cuh"
#include "mma.cuh"
#include "render_backward_mma_kernels.cuh"

namespace sagars {

template <int NQ, bool VEC, bool MD, bool COLOR>
static int launch_bwd_mma_t(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                            const uint32_t* point_list, const float* features, float* ggrad, cudaStream_t s, bool debug)
{
    auto kern = render_backward_mma_kernel<NQ, VEC, MD, COLOR>;
    const size_t smem = sizeof(BmSmem<NQ>);
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

int launch_render_backward_mma(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                               const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug)
{
    const bool md = (a.flags & SAGARS_FLAG_MASK_DEPTH) != 0;
    const bool mask_only = (a.flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    if (mask_only) return launch_bwd_mma_t<1, false, true, false>(a, d, g, im, point_list, features, ggrad, s, debug);
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define SAGARS_BWDM_CASE(NQ_)                                                                                        \
    if (nq <= NQ_) {                                                                                                 \
        if (md) return launch_bwd_mma_t<NQ_, false, true, true>(a, d, g, im, point_list, features, ggrad, s, debug);  \
        return vec ? launch_bwd_mma_t<NQ_, true, false, true>(a, d, g, im, point_list, features, ggrad, s, debug)     \
                   : launch_bwd_mma_t<NQ_, false, false, true>(a, d, g, im, point_list, features, ggrad, s, debug);   \
    }
    SAGARS_BWDM_CASE(1)
    SAGARS_BWDM_CASE(2)
    SAGARS_BWDM_CASE(4)
    SAGARS_BWDM_CASE(8)
    SAGARS_BWDM_CASE(16)
#undef SAGARS_BWDM_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
