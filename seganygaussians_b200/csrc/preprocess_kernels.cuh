// preprocess_kernels.cuh -- the device code of preprocess.cu (see there); free of host-side runtime calls so that the CPU suite
// can run it under tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include "math.cuh"

namespace sagars {


__global__ void __launch_bounds__(256)
preprocess_kernel(int P, int D, int M, int C,
                  const float* __restrict__ means3D,
                  const float* __restrict__ scales, float scale_modifier,
                  const float* __restrict__ rotations,
                  const float* __restrict__ opacities,
                  const float* __restrict__ shs,
                  const float* __restrict__ cov3D_precomp,
                  const bool have_colors,
                  const float* __restrict__ viewmatrix,
                  const float* __restrict__ projmatrix,
                  const float* __restrict__ cam_pos,
                  int W, int H, int tiles_x, int tiles_y,
                  float tan_fovx, float tan_fovy, float focal_x, float focal_y,
                  int32_t* __restrict__ radii,
                  GeomView g, uint32_t prefiltered)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t my_tiles = 0;

    if (idx < P) {
        int my_radius_i = 0;
        do {
            const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
            const float4 p_hom = xform4x4(p, projmatrix);
            const float p_w = 1.0f / (p_hom.w + 0.0000001f);
            const float3 p_proj = make_float3(p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w);
            const float3 p_view = xform4x3(p, viewmatrix);

            // near cull only (the lateral test is disabled in the reference, auxiliary.h:154)
            if (p_view.z <= 0.2f) {
                if (prefiltered) atomicOr(&g.status[0], 1u);
                break;
            }

            // 3D covariance: given, or from scale / rotation
            float c3[6];
            if (cov3D_precomp != nullptr) {
#pragma unroll
                for (int i = 0; i < 6; i++) c3[i] = cov3D_precomp[6 * idx + i];
            } else {
                const float3 sc = make_float3(scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]);
                const float4 q = *reinterpret_cast<const float4*>(rotations + 4 * idx);
                cov3d_from_scale_rot(sc, scale_modifier, q, c3);
#pragma unroll
                for (int i = 0; i < 6; i++) g.cov3D[6 * idx + i] = c3[i];
            }

            // EWA projection to a 2D covariance (+0.3 low-pass on the diagonal)
            const float3 cov = cov2d_project(p, focal_x, focal_y, tan_fovx, tan_fovy, c3, viewmatrix);

            const float det = (cov.x * cov.z - cov.y * cov.y);
            if (det == 0.0f) break;
            const float det_inv = 1.f / det;
            const float3 conic = make_float3(cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv);

            // screen-space extent from the larger eigenvalue
            const float mid = 0.5f * (cov.x + cov.z);
            const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
            const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
            const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
            const float2 pix = make_float2(ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H));

            uint2 rmin, rmax;
            tile_rect(pix, (int)my_radius, rmin, rmax, tiles_x, tiles_y);
            const uint32_t ntiles = (rmax.x - rmin.x) * (rmax.y - rmin.y);
            if (ntiles == 0) break;

            // SH -> RGB when no colours were supplied (legal only for C == 3; checked on the host)
            if (!have_colors) {
                bool cl[3];
                const float3 rgb = sh_to_rgb(idx, D, M, p, cam_pos, shs, cl);
                g.rgb[3 * idx + 0] = rgb.x;
                g.rgb[3 * idx + 1] = rgb.y;
                g.rgb[3 * idx + 2] = rgb.z;
                g.clamped[3 * idx + 0] = cl[0];
                g.clamped[3 * idx + 1] = cl[1];
                g.clamped[3 * idx + 2] = cl[2];
            }

            g.depths[idx] = p_view.z;
            const float opac = opacities[idx];
            float4* rec = reinterpret_cast<float4*>(g.geo + 8 * (size_t)idx);
            rec[0] = make_float4(pix.x, pix.y, conic.x, conic.y);
            rec[1] = make_float4(conic.z, opac, accept_threshold(opac), 0.f);
            my_radius_i = (int)my_radius;
            my_tiles = ntiles;
        } while (false);

        radii[idx] = my_radius_i;
        g.tiles_touched[idx] = my_tiles;
    }

    // per-block sum of tiles_touched -> block_sums[blockIdx.x] (input of the offsets scan)
    uint32_t v = my_tiles;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __shared__ uint32_t warp_sums[8];
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += warp_sums[w];
        g.block_sums[blockIdx.x] = t;
    }
}


// present[i] = view-space z > 0.2  (CF rasterizer_impl.cu:54-66)
__global__ void __launch_bounds__(256)
mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                    uint8_t* __restrict__ present)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 p_view = xform4x3(p, viewmatrix);
    present[idx] = (p_view.z <= 0.2f) ? 0 : 1;
}

}  // namespace sagars
