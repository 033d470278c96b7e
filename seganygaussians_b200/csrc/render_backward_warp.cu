// render_backward_warp.cu -- back-to-front gradient of the alpha compositing, ONE WARP PER CTA.
//
// Semantics: CF cuda_rasterizer/backward.cu:399-559 (DEPTH backward.cu:400-564 adds dL_dmask), SURVEY.md Appendix
// A.13-A.17 / D.  Same arithmetic as render_backward_mma.cu; different decomposition:
//
//   * a CTA is one warp = one 8x4 pixel block of a 16x16 tile (grid = 2*tiles_x by 4*tiles_y).  Nothing is shared
//     between the blocks of a tile, so there is no CTA barrier anywhere: the hardware block scheduler balances the
//     65 k independent blocks of a 1080p image, and a block only walks the tile's list as deep as ITS OWN pixels need
//     (max n_contrib of the block), not as deep as the tile's deepest pixel.  The price is that every block reads the
//     tile's records itself (8x the L2 reads of 32-byte records; feature rows are only read for candidates).
//   * per 32 list entries, lane = splat: the lane loads the splat's record straight from global memory (L2) into
//     registers -- the next chunk's record is already in flight -- and tests it against the block (candidate.cuh);
//     candidates are appended, in list order, to a small shared-memory table (record, Gaussian id, list position).
//   * whenever 8 candidates are waiting (leftovers are carried over to the next chunk, so the GEMMs always run on full
//     groups): (1) their feature rows are gathered and the dot products s = f_j . g_p of the group are formed as a
//     mma.sync m16n8k8 TF32 (3xTF32) product S = G F^T; (2) thread = pixel traversal with the scalar recurrence of
//     Appendix D, an accepted pair leaves w = alpha*T and q = G * dL/dalpha in row j of the W / Q tiles; (3)
//     dL/dcolour^T = G^T W^T and the six geometry moments Q (1, x, y, x^2, xy, y^2) (block-centred pixel coordinates,
//     exact in tf32) as mma.sync GEMMs; the results leave as red.global.add.f32: C + 6 per (block, splat) instance.
//
// Shared memory per CTA: 4 KB gradient rows + 2 KB W / Q tiles + 1.6 KB candidate table (K = 32) -> register-limited
// occupancy (25 warps per SM at 80 registers), no barriers, no tail inside a tile.
#include "common.cuh"
#include "cp_async.cuh"
#include "candidate.cuh"
#include "mma.cuh"
#include "render_backward_warp_kernels.cuh"

namespace sagars {

template <int NQ, bool VEC, bool MD, bool COLOR>
static int launch_bwd_warp_t(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                             const uint32_t* point_list, const float* features, float* ggrad, cudaStream_t s, bool debug)
{
    auto kern = render_backward_warp_kernel<NQ, VEC, MD, COLOR>;
    const size_t smem = sizeof(BwSmem<NQ>);
    {   // opt in to the dynamic shared-memory size once per device (not on every launch: the call takes the context lock)
        static DeviceOnce once;
        int dev = 0;
        SAGARS_CUDA(cudaGetDevice(&dev));
        if (once.need(dev)) {
            SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            once.done(dev);
        }
    }
    dim3 grid(2 * d.tiles_x, 4 * d.tiles_y);
    kern<<<grid, 32, smem, s>>>(im.ranges, point_list, d.W, d.H, d.C, a.background, g.geo, features,
                                im.final_T, im.n_contrib, a.dL_dout_color, a.dL_dout_mask, ggrad, a.dL_dcolors);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_render_backward_warp(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                                const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug)
{
    const bool md = (a.flags & SAGARS_FLAG_MASK_DEPTH) != 0;
    const bool mask_only = (a.flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    if (mask_only) return launch_bwd_warp_t<1, false, true, false>(a, d, g, im, point_list, features, ggrad, s, debug);
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define SAGARS_BWDW_CASE(NQ_)                                                                                         \
    if (nq <= NQ_) {                                                                                                  \
        if (md) return launch_bwd_warp_t<NQ_, false, true, true>(a, d, g, im, point_list, features, ggrad, s, debug);  \
        return vec ? launch_bwd_warp_t<NQ_, true, false, true>(a, d, g, im, point_list, features, ggrad, s, debug)     \
                   : launch_bwd_warp_t<NQ_, false, false, true>(a, d, g, im, point_list, features, ggrad, s, debug);   \
    }
    SAGARS_BWDW_CASE(1)
    SAGARS_BWDW_CASE(2)
    SAGARS_BWDW_CASE(4)
    SAGARS_BWDW_CASE(8)
    SAGARS_BWDW_CASE(16)
#undef SAGARS_BWDW_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
