// Host harness: the fused feature-smoothing kernels (seganygaussians_b200/csrc/smooth_kernels.cuh) under the execution shim, with
// the channel-count dispatch of launch_smooth_forward / launch_smooth_backward (smooth.cu) restated.  TEST INFRASTRUCTURE ONLY.
#define SAGARS_CUDA_EMU 1
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "smooth_kernels.cuh"

using namespace sagars;

static int lanes_per_row(int C) { return (C % 4 == 0 && (C == 4 || C == 8 || C == 16 || C == 32 || C == 64)) ? C / 4 : 0; }

extern "C" int emu_smooth_forward(int P, int C, int Ks, const float* F, const long long* idx, int normalize_out, float* out, float* mean_norm)
{
    const int lr = lanes_per_row(C);
    if (lr > 0) {
        const unsigned blocks = (unsigned)(((size_t)P * lr + 255) / 256);
        switch (lr) {
#define SMF(LR_) case LR_: cuda_emu::launch(blocks, 256, 0, smooth_forward_vec_kernel<LR_>, P, Ks, F, idx, normalize_out, out, mean_norm); break;
            SMF(1) SMF(2) SMF(4) SMF(8) SMF(16)
#undef SMF
        }
    } else {
        const unsigned blocks = (unsigned)(((size_t)P * 32 + 255) / 256);
        if (C <= 32) cuda_emu::launch(blocks, 256, 0, smooth_forward_kernel<1>, P, C, Ks, F, idx, normalize_out, out, mean_norm);
        else cuda_emu::launch(blocks, 256, 0, smooth_forward_kernel<2>, P, C, Ks, F, idx, normalize_out, out, mean_norm);
    }
    return 0;
}

extern "C" int emu_smooth_backward(int P, int C, int Ks, const float* F, const long long* idx, int normalize_out, const float* mean_norm,
                                   const float* out, const float* dL_dout, float* dL_dn, float* dL_dF)
{
    std::memset(dL_dn, 0, (size_t)P * C * sizeof(float));
    const int lr = lanes_per_row(C);
    if (lr > 0) {
        const unsigned blocks = (unsigned)(((size_t)P * lr + 255) / 256);
        switch (lr) {
#define SMB(LR_)                                                                                                                       \
    case LR_:                                                                                                                          \
        cuda_emu::launch(blocks, 256, 0, smooth_backward_scatter_vec_kernel<LR_>, P, Ks, idx, normalize_out, mean_norm, out, dL_dout, dL_dn); \
        cuda_emu::launch(blocks, 256, 0, smooth_backward_finalize_vec_kernel<LR_>, P, F, (const float*)dL_dn, dL_dF);                  \
        break;
            SMB(1) SMB(2) SMB(4) SMB(8) SMB(16)
#undef SMB
        }
    } else {
        const unsigned blocks = (unsigned)(((size_t)P * 32 + 255) / 256);
        if (C <= 32) {
            cuda_emu::launch(blocks, 256, 0, smooth_backward_scatter_kernel<1>, P, C, Ks, idx, normalize_out, mean_norm, out, dL_dout, dL_dn);
            cuda_emu::launch(blocks, 256, 0, smooth_backward_finalize_kernel<1>, P, C, F, (const float*)dL_dn, dL_dF);
        } else {
            cuda_emu::launch(blocks, 256, 0, smooth_backward_scatter_kernel<2>, P, C, Ks, idx, normalize_out, mean_norm, out, dL_dout, dL_dn);
            cuda_emu::launch(blocks, 256, 0, smooth_backward_finalize_kernel<2>, P, C, F, (const float*)dL_dn, dL_dF);
        }
    }
    return 0;
}
