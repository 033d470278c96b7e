// smooth.cu -- fused feature smoothing in front of the rasterizer.
//
// The reference builds the K = 32 "colours" of a training step as (scene/gaussian_model_ff.py:338-364 and
// gaussian_renderer/__init__.py:355-363)
//     normed = F.normalize(point_features, dim=-1)                  # x / max(||x||, 1e-12)
//     ret    = normed[nbr_idx[:, select], :].mean(dim=1)            # gather of Ks = 8 neighbour rows per point
//     colors = ret / (ret.norm(dim=1, keepdim=True) + 1e-9)         # when norm_point_features
// i.e. five full-size tensor passes plus a [P, Ks, C] gather forward, and an index_put / scatter-add of the same size
// backward.  Here: ONE forward kernel (C / 4 lanes per row, float4 loads of the Ks neighbour rows issued together, the
// neighbour's norm reduced with shuffles on the fly, average and optional renormalisation in registers), and for the
// backward one scatter kernel (red.global.add.v4 into a [P, C] accumulator) plus one finalising kernel that applies the
// Jacobian of the row normalisation once per row.  HBM / L2 bound: (Ks + 1) rows of C floats + 8 Ks index bytes per point
// forward.  Channel counts that are not 4, 8, 16, 32 or 64 use the scalar warp-per-row variants.
#include "common.cuh"
#include "cp_async.cuh"
#include "smooth_kernels.cuh"

namespace sagars {

static int smooth_lanes_per_row(int C) { return (C % 4 == 0 && (C == 4 || C == 8 || C == 16 || C == 32 || C == 64)) ? C / 4 : 0; }

int launch_smooth_forward(int P, int C, int Ks, const float* F, const long long* idx, int normalize_out, float* out,
                          float* mean_norm, cudaStream_t s)
{
    const int lr = smooth_lanes_per_row(C);
    if (lr > 0) {
        const int blocks = (int)(((size_t)P * lr + 255) / 256);
        switch (lr) {
#define SAGARS_SMF(LR_) case LR_: smooth_forward_vec_kernel<LR_><<<blocks, 256, 0, s>>>(P, Ks, F, idx, normalize_out, out, mean_norm); break;
            SAGARS_SMF(1) SAGARS_SMF(2) SAGARS_SMF(4) SAGARS_SMF(8) SAGARS_SMF(16)
#undef SAGARS_SMF
        }
    } else {
        const int blocks = (int)(((size_t)P * 32 + 255) / 256);
        if (C <= 32) smooth_forward_kernel<1><<<blocks, 256, 0, s>>>(P, C, Ks, F, idx, normalize_out, out, mean_norm);
        else smooth_forward_kernel<2><<<blocks, 256, 0, s>>>(P, C, Ks, F, idx, normalize_out, out, mean_norm);
    }
    SAGARS_LAUNCH_CHECK(s, false);
    return SAGARS_OK;
}

int launch_smooth_backward(int P, int C, int Ks, const float* F, const long long* idx, int normalize_out,
                           const float* mean_norm, const float* out, const float* dL_dout, float* dL_dn, float* dL_dF,
                           cudaStream_t s)
{
    SAGARS_CUDA(cudaMemsetAsync(dL_dn, 0, (size_t)P * C * sizeof(float), s));
    const int lr = smooth_lanes_per_row(C);
    if (lr > 0) {
        const int blocks = (int)(((size_t)P * lr + 255) / 256);
        switch (lr) {
#define SAGARS_SMB(LR_)                                                                                                        \
    case LR_:                                                                                                                  \
        smooth_backward_scatter_vec_kernel<LR_><<<blocks, 256, 0, s>>>(P, Ks, idx, normalize_out, mean_norm, out, dL_dout, dL_dn); \
        SAGARS_LAUNCH_CHECK(s, false);                                                                                         \
        smooth_backward_finalize_vec_kernel<LR_><<<blocks, 256, 0, s>>>(P, F, dL_dn, dL_dF);                                   \
        break;
            SAGARS_SMB(1) SAGARS_SMB(2) SAGARS_SMB(4) SAGARS_SMB(8) SAGARS_SMB(16)
#undef SAGARS_SMB
        }
    } else {
        const int blocks = (int)(((size_t)P * 32 + 255) / 256);
        if (C <= 32) {
            smooth_backward_scatter_kernel<1><<<blocks, 256, 0, s>>>(P, C, Ks, idx, normalize_out, mean_norm, out, dL_dout, dL_dn);
            SAGARS_LAUNCH_CHECK(s, false);
            smooth_backward_finalize_kernel<1><<<blocks, 256, 0, s>>>(P, C, F, dL_dn, dL_dF);
        } else {
            smooth_backward_scatter_kernel<2><<<blocks, 256, 0, s>>>(P, C, Ks, idx, normalize_out, mean_norm, out, dL_dout, dL_dn);
            SAGARS_LAUNCH_CHECK(s, false);
            smooth_backward_finalize_kernel<2><<<blocks, 256, 0, s>>>(P, C, F, dL_dn, dL_dF);
        }
    }
    SAGARS_LAUNCH_CHECK(s, false);
    return SAGARS_OK;
}

}  // namespace sagars
