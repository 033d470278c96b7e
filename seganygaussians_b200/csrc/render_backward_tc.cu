// render_backward_tc.cu -- back-to-front gradient of the alpha compositing on the 5th-generation tensor cores (tcgen05 / TMEM),
// C = 32 colour channels.
//
// Semantics: CF cuda_rasterizer/backward.cu:399-559, SURVEY.md Appendix A.13-A.17 / D.  Same arithmetic as
// render_backward_warp.cu; what changes is WHO multiplies.  The warp kernel spends ~75 of its ~210 warp instructions per
// (block, splat) candidate on issuing mma.sync products and on loading / splitting their fragments; here one elected thread
// issues tcgen05.mma for the whole CTA and the other threads only produce and consume operand tiles:
//
//   * a CTA = 128 threads = one 16 x 8 pixel group (half a tile) = four warps, each owning an 8x4 pixel block.  The pixels'
//     upstream gradient rows are split once into tf32 high parts and remainders and stay in shared memory for the whole
//     list as two K-major A operands: rows = pixels for S = G F^T, and rows = channels (+ the six-row moment basis) with the
//     pixels as contraction index for the gradient product.
//   * per 32 list entries (lane = splat, every warp reads the chunk) each warp tests the splats against its own block
//     (candidate.cuh); the union of the four masks joins a small ring in list order, with a 4-bit membership mask.
//   * per batch of 16 candidates: (1) their feature rows are gathered, split and written as the B operand; 12 tcgen05.mma
//     (3xTF32) leave S (128 pixels x 16) in TMEM, read back with tcgen05.ld: thread = pixel = accumulator lane, so the 16 dot
//     products of a pixel arrive in its registers with no shared-memory round trip; (2) thread = pixel back-to-front traversal
//     with the scalar recurrence of Appendix D (warps skip candidates that are not theirs); w = alpha T and q = G dL/dalpha go,
//     split hi / lo, into the B operand [w_hi | w_lo | q_hi | q_lo] (64 columns x 128 pixels); (3) 16 tcgen05.mma with
//     M = 128 rows (32 gradient high parts, 32 remainders, 6 basis rows) give dL/dcolour (all four hi / lo cross terms, i.e.
//     better than 3xTF32) and the six geometry moments in one accumulator; its epilogue (tcgen05.ld, one coalesced 128-byte
//     red.global.add per candidate and row group, moments -> the six geometry gradients) runs one batch later, while the
//     tensor core works on the next S.
//
// Shared memory per CTA: 32 + 37 KB gradient tiles (one per contraction: kind::tf32 MN-major operands measured unusable, see the
// kernel header) + 33 KB w / q tile + 4 KB feature tiles + 5 KB tables = 112 KB -> two CTAs per SM; TMEM: 128 columns per CTA.
#include "common.cuh"
#include "render_backward_tc_kernels.cuh"

namespace sagars {

// two CTAs per SM: 2 x (dynamic + 1 KB reserved per CTA) must fit the SM's 228 KB
static_assert(sizeof(BtSmem) + 1024 <= 115712, "tcgen05 backward: shared memory of one CTA exceeds half an SM");

int launch_render_backward_tc(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                              const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug)
{
    if (d.C != BT_C || a.colors_precomp == nullptr) {
        set_error("the tcgen05 backward kernel handles C = %d precomputed colours only", BT_C);
        return SAGARS_EINVAL;
    }
    auto kern = render_backward_tc_kernel;
    const size_t smem = sizeof(BtSmem) + 1024;
    {   // opt in to the dynamic shared-memory size once per device
        static DeviceOnce once;
        int dev = 0;
        SAGARS_CUDA(cudaGetDevice(&dev));
        if (once.need(dev)) {
            SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            once.done(dev);
        }
    }
    dim3 grid(d.tiles_x, 2 * d.tiles_y);
    kern<<<grid, BT_THREADS, smem, s>>>(im.ranges, point_list, d.W, d.H, a.background, g.geo, a.colors_precomp, im.final_T, im.n_contrib,
                                    a.dL_dout_color, ggrad, a.dL_dcolors);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

}  // namespace sagars

#if defined(SAGARS_BT_TIMELINE)
// developer hook (only in -DSAGARS_BT_TIMELINE builds): copy and clear the kernel's phase counters
extern "C" __attribute__((visibility("default"))) int sagars_bt_timeline_read(unsigned long long* out32)
{
    if (cudaMemcpyFromSymbol(out32, sagars_bt_timeline, sizeof(unsigned long long) * 32) != cudaSuccess) return 1;
    unsigned long long z[32] = {0};
    return cudaMemcpyToSymbol(sagars_bt_timeline, z, sizeof(z)) == cudaSuccess ? 0 : 1;
}
#endif
