// Host harness: the kernels of seganygaussians_b200/csrc/sample_kernels.cuh (fused norm + bilinear ray sampling) under the execution
// shim, in the order sample.cu launches them.  TEST INFRASTRUCTURE ONLY.
#include <cuda_runtime.h>            // the shim (this directory comes first on the include path)
#include "sample_kernels.cuh"

using namespace sagars;

extern "C" int emu_sample_forward(int C, int H, int W, int h, int w, const float* img, const long long* rays, int S, float* out, float* norm_sum)
{
    *norm_sum = 0.f;
    cuda_emu::launch((H * W + 255) / 256, 256, 0, pixel_norm_sum_kernel, img, C, H * W, norm_sum);
    if (S > 0) cuda_emu::launch((unsigned)(((long long)S * C + 255) / 256), 256, 0, sample_rays_forward_kernel, img, C, H, W, h, w, rays, S, out);
    return 0;
}
extern "C" int emu_sample_backward(int C, int H, int W, int h, int w, const float* img, const long long* rays, int S, const float* g_out,
                                   const float* g_norm, float* grad_img)
{
    cuda_emu::launch((H * W + 255) / 256, 256, 0, sample_rays_backward_dense_kernel, img, C, H * W, g_norm, grad_img);
    if (S > 0) cuda_emu::launch((unsigned)(((long long)S * C + 255) / 256), 256, 0, sample_rays_backward_taps_kernel, C, H, W, h, w, rays, S, g_out, grad_img);
    return 0;
}
