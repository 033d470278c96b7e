// sample.cu -- the loss-side consumer of the K-feature render, fused:
//
//   reference (train_contrastive_feature.py:232-254):  norm = render.norm(dim=0).mean()                       [C,H,W] -> scalar
//                                                      up   = interpolate(render[None], (h, w), 'bilinear')[0] [C,h,w]  (265 MB at 1080p)
//                                                      samp = (up * gate)[:, :, sampled_ray]                   ~1000 rays are read
//   here: ONE pass over the render for the norm term and a direct four-tap read of the sampled rays -- the resized image is never
//   materialised (forward: 2 x 265 MB less HBM traffic; backward: the dense gradient of the norm term is written once and the rays'
//   taps are added into it, instead of interpolate's full-image backward).
#include "common.cuh"
#include "sample_kernels.cuh"

namespace sagars {

int launch_sample_rays_forward(int C, int H, int W, int h, int w, const float* img, const long long* rays, int S, float* out,
                               float* norm_sum, cudaStream_t s)
{
    const int HW = H * W;
    SAGARS_CUDA(cudaMemsetAsync(norm_sum, 0, sizeof(float), s));
    pixel_norm_sum_kernel<<<(HW + 255) / 256, 256, 0, s>>>(img, C, HW, norm_sum);
    SAGARS_LAUNCH_CHECK(s, false);
    if (S > 0) {
        const long long n = (long long)S * C;
        sample_rays_forward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(img, C, H, W, h, w, rays, S, out);
        SAGARS_LAUNCH_CHECK(s, false);
    }
    return SAGARS_OK;
}

int launch_sample_rays_backward(int C, int H, int W, int h, int w, const float* img, const long long* rays, int S, const float* g_out,
                                const float* g_norm, float* grad_img, cudaStream_t s)
{
    const int HW = H * W;
    sample_rays_backward_dense_kernel<<<(HW + 255) / 256, 256, 0, s>>>(img, C, HW, g_norm, grad_img);
    SAGARS_LAUNCH_CHECK(s, false);
    if (S > 0) {
        const long long n = (long long)S * C;
        sample_rays_backward_taps_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(C, H, W, h, w, rays, S, g_out, grad_img);
        SAGARS_LAUNCH_CHECK(s, false);
    }
    return SAGARS_OK;
}

}  // namespace sagars
