// render_forward_tc.cu -- forward blend with the C-channel contraction on the 5th-generation tensor cores.
//
// Same semantics as render_forward.cu (CF cuda_rasterizer/forward.cu:264-385), same per-pixel scalar arithmetic
// for power / alpha / T / the 1/255 and 1e-4 tests (so final_T and n_contrib stay bit-identical), but the colour
// accumulation  C[p][ch] = sum_j w[p][j] * f[j][ch],  w = alpha * T,  is expressed as the dense per-tile GEMM of
// SURVEY.md Appendix D and runs on tcgen05:
//
//   * a CTA is one 16x16 tile = two independent 128-pixel groups (4 warps each).  Thread t of a group owns pixel row t
//     of that group's A operand and TMEM lane t of its accumulator D[128 x 32] (fp32, 32 TMEM columns per group);
//   * per staged batch of 64 splats every warp first tests the splats, one per lane, against its own 8x4 pixel block
//     (candidate.cuh); the four warps of a group OR their candidate masks, and only the group's candidates (about half
//     of the tile's list) become k-slots of the GEMM.  Taken four slots at a time in list order: a warp runs the
//     reference's per-pixel test only for its own candidates (w = alpha * T, 0 otherwise), every thread writes its row
//     of the K-major tf32 operand tiles W_hi / W_lo (SWIZZLE_NONE canonical layout) with one 16-byte store each, and
//     the group's 128 threads transpose-split the four feature rows into F_hi / F_lo;
//   * every 16 slots one elected thread per group issues  D += W_lo F_hi + W_hi F_lo + W_hi F_hi  (3xTF32: 6
//     tcgen05.mma.kind::tf32 of 128x32x8, ~2^-21 relative error) and commits to the group's mbarrier; the next slots'
//     scalar work overlaps the MMAs and only waits for them right before it overwrites the operand tiles;
//   * at the end each thread reads its 32 channel sums with one tcgen05.ld and adds T * bg.
//
// The colour image differs from the SIMT kernel / the reference by fp32-level rounding only (3xTF32 and a different
// summation order), far inside the 1e-4 parity tolerance; integer state is unaffected.
#include "common.cuh"
#include "cp_async.cuh"
#include "tc.cuh"
#include "candidate.cuh"
#include "render_forward_tc_kernels.cuh"

namespace sagars {

int launch_render_forward_tc(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                             const uint32_t* point_list, cudaStream_t s, bool debug)
{
    auto kern = render_forward_tc_kernel;
    const size_t smem = sizeof(FwdTcSmem) + 1024;
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
    kern<<<grid, TILE_PIX, smem, s>>>(im.ranges, point_list, d.W, d.H, g.geo, a.colors_precomp, a.background,
                                      im.final_T, im.n_contrib, a.out_color);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

}  // namespace sagars
