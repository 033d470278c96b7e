// render_forward_warp.cu -- front-to-back alpha compositing of C feature channels, ONE WARP PER CTA.
//
// Semantics: CF cuda_rasterizer/forward.cu:264-385, SURVEY.md Appendix A.10-A.12.  Same per-pixel scalar arithmetic
// as the reference for power / alpha / T / the 1/255 and 1e-4 tests (final_T and n_contrib stay bit-identical); the
// colour accumulation C[p][ch] = sum_j w[p][j] f[j][ch], w = alpha * T, runs on the tensor cores per warp.
//
//   * a CTA is one warp = one 8x4 pixel block of a 16x16 tile (grid = 2*tiles_x by 4*tiles_y): no CTA barrier, no
//     shared staging, the block scheduler balances ~65 k independent blocks of a 1080p image, and a block stops
//     walking the tile's list as soon as ITS 32 pixels are saturated;
//   * per 32 list entries, lane = splat: the record comes straight from global memory (L2) into registers (the next
//     chunk's is already in flight) and is tested against the block (candidate.cuh); candidates join a small
//     shared-memory table in list order, leftovers are carried over so the products always run on 8 candidates;
//   * per group of 8 candidates: the 8 feature rows are gathered, thread = pixel runs the reference's test / blend
//     chain and leaves w in row j of the W tile, then  C (32 x C) += W^T (32 x 8) F (8 x C)  as mma.sync m16n8k8 TF32
//     in 3xTF32 split precision (~2^-21), accumulators (32 pixels x C channels) in registers for the whole list;
//   * epilogue: fragment -> planar image (+ T * bg), 32-byte segments per store.
//
// Colours differ from the fp32 SIMT kernel / the reference by fp32-level rounding only (3xTF32, summation order).
#include "common.cuh"
#include "candidate.cuh"
#include "mma.cuh"
#include "render_forward_warp_kernels.cuh"

namespace sagars {

template <int NQ, bool VEC>
static int launch_fwd_warp_t(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                             const uint32_t* point_list, const float* features, cudaStream_t s, bool debug)
{
    auto kern = render_forward_warp_kernel<NQ, VEC>;
    const size_t smem = sizeof(FwSmem<NQ>);
    {   // once per device (not on every launch: the call takes the context lock)
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
    kern<<<grid, 32, smem, s>>>(im.ranges, point_list, d.W, d.H, d.C, g.geo, features, a.background,
                                im.final_T, im.n_contrib, a.out_color);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

// colour-only forward (no mask / depth channels), any channel count up to 64
int launch_render_forward_warp(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                               const uint32_t* point_list, cudaStream_t s, bool debug)
{
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    const bool vec = (K % 4) == 0;
    const int nq = (K + 3) / 4;
#define SAGARS_FWDW_CASE(NQ_)                                                                                   \
    if (nq <= NQ_) {                                                                                            \
        return vec ? launch_fwd_warp_t<NQ_, true>(a, d, g, im, point_list, features, s, debug)                  \
                   : launch_fwd_warp_t<NQ_, false>(a, d, g, im, point_list, features, s, debug);                \
    }
    SAGARS_FWDW_CASE(1)
    SAGARS_FWDW_CASE(2)
    SAGARS_FWDW_CASE(4)
    SAGARS_FWDW_CASE(8)
    SAGARS_FWDW_CASE(16)
#undef SAGARS_FWDW_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
