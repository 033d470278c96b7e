// smooth_kernels.cuh -- the device code of smooth.cu (see there); free of host-side runtime calls so that the CPU suite can run
// it under tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include "cp_async.cuh"

namespace sagars {


constexpr float SMOOTH_EPS_IN = 1e-12f;    // F.normalize eps
constexpr float SMOOTH_EPS_OUT = 1e-9f;    // the renderer's "+ 1e-9"

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per output row: mean of the Ks normalised neighbour rows, optional renormalisation
template <int CU>
__global__ void __launch_bounds__(256)
smooth_forward_kernel(int P, int C, int Ks, const float* __restrict__ F, const long long* __restrict__ idx,
                      int normalize_out, float* __restrict__ out, float* __restrict__ mean_norm)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= P) return;
    float acc[CU];
#pragma unroll
    for (int u = 0; u < CU; u++) acc[u] = 0.f;
    for (int k0 = 0; k0 < Ks; k0 += 8) {
        long long j[8];
#pragma unroll
        for (int k = 0; k < 8; k++) j[k] = (k0 + k < Ks) ? idx[(size_t)row * Ks + k0 + k] : -1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (j[k] < 0) continue;                       // warp-uniform
            float x[CU], ss = 0.f;
#pragma unroll
            for (int u = 0; u < CU; u++) {
                const int c = lane + 32 * u;
                x[u] = c < C ? F[(size_t)j[k] * C + c] : 0.f;
                ss += x[u] * x[u];
            }
            const float nrm = fmaxf(sqrtf(warp_sum(ss)), SMOOTH_EPS_IN);
#pragma unroll
            for (int u = 0; u < CU; u++) acc[u] += x[u] / nrm;
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int u = 0; u < CU; u++) { acc[u] = acc[u] / (float)Ks; ss += acc[u] * acc[u]; }
    float s = 0.f;
    if (normalize_out) {
        s = sqrtf(warp_sum(ss));
        if (lane == 0) mean_norm[row] = s;
    }
#pragma unroll
    for (int u = 0; u < CU; u++) {
        const int c = lane + 32 * u;
        if (c < C) out[(size_t)row * C + c] = normalize_out ? acc[u] / (s + SMOOTH_EPS_OUT) : acc[u];
    }
}

// backward, part 1: one warp per output row i: dL/dm_i (through the optional renormalisation), then
// dL/dn_j += dL/dm_i / Ks for its Ks neighbours j
template <int CU>
__global__ void __launch_bounds__(256)
smooth_backward_scatter_kernel(int P, int C, int Ks, const long long* __restrict__ idx, int normalize_out,
                               const float* __restrict__ mean_norm, const float* __restrict__ out,
                               const float* __restrict__ dL_dout, float* __restrict__ dL_dn)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= P) return;
    float g[CU], o[CU];
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < CU; u++) {
        const int c = lane + 32 * u;
        g[u] = c < C ? dL_dout[(size_t)row * C + c] : 0.f;
        o[u] = (normalize_out && c < C) ? out[(size_t)row * C + c] : 0.f;
        dot += g[u] * o[u];
    }
    if (normalize_out) {
        // out = m / (s + eps), s = ||m||:  dL/dm = g / (s + eps) - out * (g . out) / s
        dot = warp_sum(dot);
        const float s = mean_norm[row];
#pragma unroll
        for (int u = 0; u < CU; u++) g[u] = g[u] / (s + SMOOTH_EPS_OUT) - (s > 0.f ? o[u] * dot / s : 0.f);
    }
#pragma unroll
    for (int u = 0; u < CU; u++) g[u] = g[u] / (float)Ks;
    for (int k = 0; k < Ks; k++) {
        const long long j = idx[(size_t)row * Ks + k];
#pragma unroll
        for (int u = 0; u < CU; u++) {
            const int c = lane + 32 * u;
            if (c < C) red_add(dL_dn + (size_t)j * C + c, g[u]);
        }
    }
}

// backward, part 2: one warp per row j: dL/dF_j = J_normalize(F_j)^T dL/dn_j
template <int CU>
__global__ void __launch_bounds__(256)
smooth_backward_finalize_kernel(int P, int C, const float* __restrict__ F, const float* __restrict__ dL_dn,
                                float* __restrict__ dL_dF)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= P) return;
    float x[CU], d[CU];
    float dot = 0.f, ss = 0.f;
#pragma unroll
    for (int u = 0; u < CU; u++) {
        const int c = lane + 32 * u;
        x[u] = c < C ? F[(size_t)row * C + c] : 0.f;
        d[u] = c < C ? dL_dn[(size_t)row * C + c] : 0.f;
        dot += x[u] * d[u];
        ss += x[u] * x[u];
    }
    dot = warp_sum(dot);
    const float nrm = fmaxf(sqrtf(warp_sum(ss)), SMOOTH_EPS_IN);
    // n = x / max(||x||, eps).  ||x|| > eps: dL/dx = (d - n (n . d)) / ||x||;  clamped: the denominator is a constant
    const bool clamped = !(nrm > SMOOTH_EPS_IN);
#pragma unroll
    for (int u = 0; u < CU; u++) {
        const int c = lane + 32 * u;
        if (c < C) dL_dF[(size_t)row * C + c] = clamped ? d[u] / nrm : (d[u] - x[u] * (dot / (nrm * nrm))) / nrm;
    }
}

// ---- vectorised variants for C % 4 == 0: LR = C / 4 lanes per row (power of two), 32 / LR rows per warp ----
template <int LR>
__device__ __forceinline__ float group_sum(float v)
{
#pragma unroll
    for (int o = LR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, LR);
    return v;
}

template <int LR>
__global__ void __launch_bounds__(256)
smooth_forward_vec_kernel(int P, int Ks, const float* __restrict__ F, const long long* __restrict__ idx,
                          int normalize_out, float* __restrict__ out, float* __restrict__ mean_norm)
{
    constexpr int C = 4 * LR;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / LR, l = threadIdx.x % LR;
    const bool live = gid < P;
    const int row = live ? gid : P - 1;                 // whole groups stay converged for the shuffles
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < Ks; k0 += 8) {
        float4 x[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + k < Ks) {
                const long long j = idx[(size_t)row * Ks + k0 + k];
                x[k] = *reinterpret_cast<const float4*>(F + (size_t)j * C + 4 * l);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k0 + k < Ks) {
                const float ss = group_sum<LR>(x[k].x * x[k].x + x[k].y * x[k].y + x[k].z * x[k].z + x[k].w * x[k].w);
                const float nrm = fmaxf(sqrtf(ss), SMOOTH_EPS_IN);
                acc.x += x[k].x / nrm; acc.y += x[k].y / nrm; acc.z += x[k].z / nrm; acc.w += x[k].w / nrm;
            }
        }
    }
    const float kf = (float)Ks;
    acc.x /= kf; acc.y /= kf; acc.z /= kf; acc.w /= kf;
    if (normalize_out) {
        const float s = sqrtf(group_sum<LR>(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w));
        if (live && l == 0) mean_norm[row] = s;
        const float dn = s + SMOOTH_EPS_OUT;
        acc.x /= dn; acc.y /= dn; acc.z /= dn; acc.w /= dn;
    }
    if (live) *reinterpret_cast<float4*>(out + (size_t)row * C + 4 * l) = acc;
}

template <int LR>
__global__ void __launch_bounds__(256)
smooth_backward_scatter_vec_kernel(int P, int Ks, const long long* __restrict__ idx, int normalize_out,
                                   const float* __restrict__ mean_norm, const float* __restrict__ out,
                                   const float* __restrict__ dL_dout, float* __restrict__ dL_dn)
{
    constexpr int C = 4 * LR;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / LR, l = threadIdx.x % LR;
    const bool live = gid < P;
    const int row = live ? gid : P - 1;
    float4 g = *reinterpret_cast<const float4*>(dL_dout + (size_t)row * C + 4 * l);
    if (normalize_out) {
        const float4 o = *reinterpret_cast<const float4*>(out + (size_t)row * C + 4 * l);
        const float dot = group_sum<LR>(g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w);
        const float s = mean_norm[row];
        const float dn = s + SMOOTH_EPS_OUT, t = s > 0.f ? dot / s : 0.f;
        g.x = g.x / dn - o.x * t; g.y = g.y / dn - o.y * t; g.z = g.z / dn - o.z * t; g.w = g.w / dn - o.w * t;
    }
    const float kf = (float)Ks;
    g.x /= kf; g.y /= kf; g.z /= kf; g.w /= kf;
    if (!live) return;
    for (int k = 0; k < Ks; k++) {
        const long long j = idx[(size_t)row * Ks + k];
        red_add_v4(dL_dn + (size_t)j * C + 4 * l, g.x, g.y, g.z, g.w);
    }
}

template <int LR>
__global__ void __launch_bounds__(256)
smooth_backward_finalize_vec_kernel(int P, const float* __restrict__ F, const float* __restrict__ dL_dn,
                                    float* __restrict__ dL_dF)
{
    constexpr int C = 4 * LR;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / LR, l = threadIdx.x % LR;
    const bool live = gid < P;
    const int row = live ? gid : P - 1;
    const float4 x = *reinterpret_cast<const float4*>(F + (size_t)row * C + 4 * l);
    const float4 d = *reinterpret_cast<const float4*>(dL_dn + (size_t)row * C + 4 * l);
    const float dot = group_sum<LR>(x.x * d.x + x.y * d.y + x.z * d.z + x.w * d.w);
    const float nrm = fmaxf(sqrtf(group_sum<LR>(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w)), SMOOTH_EPS_IN);
    float4 r;
    if (!(nrm > SMOOTH_EPS_IN)) {
        r = make_float4(d.x / nrm, d.y / nrm, d.z / nrm, d.w / nrm);
    } else {
        const float t = dot / (nrm * nrm);
        r = make_float4((d.x - x.x * t) / nrm, (d.y - x.y * t) / nrm, (d.z - x.z * t) / nrm, (d.w - x.w * t) / nrm);
    }
    if (live) *reinterpret_cast<float4*>(dL_dF + (size_t)row * C + 4 * l) = r;
}

}  // namespace sagars
