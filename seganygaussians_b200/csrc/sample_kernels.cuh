// sample_kernels.cuh -- the device code of sample.cu (see there).  Free of host-side runtime calls so that the CPU suite can run
// these kernels under tests/cuda_emu/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sagars {

// source coordinate of torch.nn.functional.interpolate(mode='bilinear', align_corners=False) for output index `dst`:
// src = scale * (dst + 0.5) - 0.5 clamped at 0 (ATen area_pixel_compute_source_index), scale = in / out
__device__ __forceinline__ void bilinear_tap(int dst, float scale, int in_size, int& i0, int& i1, float& w1)
{
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i0 = i0 > in_size - 1 ? in_size - 1 : i0;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    w1 = src - (float)i0;
}

// sum over the pixels of ||x_p||_2 (x_p = the C channel values of pixel p of a planar [C, H*W] image): one partial sum per block
// into norm_sum[0] (one atomic per block).
__global__ void __launch_bounds__(256)
pixel_norm_sum_kernel(const float* __restrict__ img, int C, int HW, float* __restrict__ norm_sum)
{
    __shared__ float warp_part[8];
    const int p = blockIdx.x * 256 + threadIdx.x;
    float n = 0.f;
    if (p < HW) {
        float s = 0.f;
        for (int c = 0; c < C; c++) { const float v = img[(size_t)c * HW + p]; s += v * v; }
        n = sqrtf(s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += warp_part[w];
        atomicAdd(norm_sum, t);
    }
}

// out[c, s] = bilinear resize of img [C, H, W] to [h, w], read at flat output position rays[s]: thread = (ray, channel)
__global__ void __launch_bounds__(256)
sample_rays_forward_kernel(const float* __restrict__ img, int C, int H, int W, int h, int w, const long long* __restrict__ rays, int S,
                           float* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)S * C) return;
    const int s = (int)(t % S), c = (int)(t / S);            // consecutive threads: consecutive rays of one channel (coalesced store)
    const long long r = rays[s];
    const int oy = (int)(r / w), ox = (int)(r % w);
    int y0, y1, x0, x1;
    float wy, wx;
    bilinear_tap(oy, (float)H / (float)h, H, y0, y1, wy);
    bilinear_tap(ox, (float)W / (float)w, W, x0, x1, wx);
    const float* pl = img + (size_t)c * H * W;
    const float v00 = pl[(size_t)y0 * W + x0], v01 = pl[(size_t)y0 * W + x1], v10 = pl[(size_t)y1 * W + x0], v11 = pl[(size_t)y1 * W + x1];
    out[(size_t)c * S + s] = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
}

// dense part of the gradient: d(mean_p ||x_p||)/dx[c, p] * g_norm = g_norm / HW * x[c, p] / ||x_p||  (0 where the norm is 0, as
// torch.linalg.vector_norm's backward); writes every element of grad_img
__global__ void __launch_bounds__(256)
sample_rays_backward_dense_kernel(const float* __restrict__ img, int C, int HW, const float* __restrict__ g_norm, float* __restrict__ grad_img)
{
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= HW) return;
    float s = 0.f;
    for (int c = 0; c < C; c++) { const float v = img[(size_t)c * HW + p]; s += v * v; }
    const float n = sqrtf(s);
    const float k = (n > 0.f) ? (g_norm[0] / (float)HW) / n : 0.f;
    for (int c = 0; c < C; c++) grad_img[(size_t)c * HW + p] = k * img[(size_t)c * HW + p];
}

// sparse part: the four taps of every sampled ray
__global__ void __launch_bounds__(256)
sample_rays_backward_taps_kernel(int C, int H, int W, int h, int w, const long long* __restrict__ rays, int S, const float* __restrict__ g_out,
                                 float* __restrict__ grad_img)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)S * C) return;
    const int s = (int)(t % S), c = (int)(t / S);
    const long long r = rays[s];
    const int oy = (int)(r / w), ox = (int)(r % w);
    int y0, y1, x0, x1;
    float wy, wx;
    bilinear_tap(oy, (float)H / (float)h, H, y0, y1, wy);
    bilinear_tap(ox, (float)W / (float)w, W, x0, x1, wx);
    const float g = g_out[(size_t)c * S + s];
    float* pl = grad_img + (size_t)c * H * W;
    atomicAdd(pl + (size_t)y0 * W + x0, g * (1.f - wy) * (1.f - wx));
    atomicAdd(pl + (size_t)y0 * W + x1, g * (1.f - wy) * wx);
    atomicAdd(pl + (size_t)y1 * W + x0, g * wy * (1.f - wx));
    atomicAdd(pl + (size_t)y1 * W + x1, g * wy * wx);
}

}  // namespace sagars
