// geom_backward_kernels.cuh -- the device code of geom_backward.cu (see there); free of host-side runtime calls so that the CPU
// suite can run it under tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include "math.cuh"

namespace sagars {


__device__ __forceinline__ float3 dnormvdv3(const float3 v, const float3 dv)
{
    const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    float3 r;
    r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
    r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
    r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
    return r;
}

// SH backward for one Gaussian (CF backward.cu:20-139): writes dL_dsh rows, returns the part of
// dL/dmean that flows through the view direction.
__device__ __forceinline__ float3 sh_backward(int idx, int deg, int M, const float3 pos, const float* __restrict__ cam_pos,
                                              const float* __restrict__ shs, const uint8_t* __restrict__ clamped,
                                              const float* __restrict__ dL_dcolor /* [P,3] */, float* __restrict__ dL_dsh)
{
    const float3 dir_orig = make_float3(pos.x - cam_pos[0], pos.y - cam_pos[1], pos.z - cam_pos[2]);
    const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
    const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
    const float* sh = shs + (size_t)idx * M * 3;
    float* out = dL_dsh + (size_t)idx * M * 3;

    float dRGB[3];
#pragma unroll
    for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor[3 * idx + c] * (clamped[3 * idx + c] ? 0.f : 1.f);

    float ddir[3] = {0.f, 0.f, 0.f};   // dL/d(dir) accumulated over the three colour channels
#pragma unroll
    for (int c = 0; c < 3; c++) {
#define SHC(k) sh[(k) * 3 + c]
#define OUT(k) out[(k) * 3 + c]
        const float dL = dRGB[c];
        float dx = 0.f, dy = 0.f, dz = 0.f;   // dRGB/d(x,y,z) of this channel
        OUT(0) = SAGARS_SH_C0 * dL;
        if (deg > 0) {
            OUT(1) = (-SAGARS_SH_C1 * y) * dL;
            OUT(2) = (SAGARS_SH_C1 * z) * dL;
            OUT(3) = (-SAGARS_SH_C1 * x) * dL;
            dx = -SAGARS_SH_C1 * SHC(3);
            dy = -SAGARS_SH_C1 * SHC(1);
            dz = SAGARS_SH_C1 * SHC(2);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z;
                const float xy = x * y, yz = y * z, xz = x * z;
                OUT(4) = (kSH_C2[0] * xy) * dL;
                OUT(5) = (kSH_C2[1] * yz) * dL;
                OUT(6) = (kSH_C2[2] * (2.f * zz - xx - yy)) * dL;
                OUT(7) = (kSH_C2[3] * xz) * dL;
                OUT(8) = (kSH_C2[4] * (xx - yy)) * dL;
                dx += kSH_C2[0] * y * SHC(4) + kSH_C2[2] * 2.f * -x * SHC(6) + kSH_C2[3] * z * SHC(7) + kSH_C2[4] * 2.f * x * SHC(8);
                dy += kSH_C2[0] * x * SHC(4) + kSH_C2[1] * z * SHC(5) + kSH_C2[2] * 2.f * -y * SHC(6) + kSH_C2[4] * 2.f * -y * SHC(8);
                dz += kSH_C2[1] * y * SHC(5) + kSH_C2[2] * 2.f * 2.f * z * SHC(6) + kSH_C2[3] * x * SHC(7);
                if (deg > 2) {
                    OUT(9) = (kSH_C3[0] * y * (3.f * xx - yy)) * dL;
                    OUT(10) = (kSH_C3[1] * xy * z) * dL;
                    OUT(11) = (kSH_C3[2] * y * (4.f * zz - xx - yy)) * dL;
                    OUT(12) = (kSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL;
                    OUT(13) = (kSH_C3[4] * x * (4.f * zz - xx - yy)) * dL;
                    OUT(14) = (kSH_C3[5] * z * (xx - yy)) * dL;
                    OUT(15) = (kSH_C3[6] * x * (xx - 3.f * yy)) * dL;
                    dx += (kSH_C3[0] * SHC(9) * 3.f * 2.f * xy + kSH_C3[1] * SHC(10) * yz + kSH_C3[2] * SHC(11) * -2.f * xy +
                           kSH_C3[3] * SHC(12) * -3.f * 2.f * xz + kSH_C3[4] * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
                           kSH_C3[5] * SHC(14) * 2.f * xz + kSH_C3[6] * SHC(15) * 3.f * (xx - yy));
                    dy += (kSH_C3[0] * SHC(9) * 3.f * (xx - yy) + kSH_C3[1] * SHC(10) * xz +
                           kSH_C3[2] * SHC(11) * (-3.f * yy + 4.f * zz - xx) + kSH_C3[3] * SHC(12) * -3.f * 2.f * yz +
                           kSH_C3[4] * SHC(13) * -2.f * xy + kSH_C3[5] * SHC(14) * -2.f * yz +
                           kSH_C3[6] * SHC(15) * -3.f * 2.f * xy);
                    dz += (kSH_C3[1] * SHC(10) * xy + kSH_C3[2] * SHC(11) * 4.f * 2.f * yz +
                           kSH_C3[3] * SHC(12) * 3.f * (2.f * zz - xx - yy) + kSH_C3[4] * SHC(13) * 4.f * 2.f * xz +
                           kSH_C3[5] * SHC(14) * (xx - yy));
                }
            }
        }
        // coefficients above the active degree receive no gradient
        for (int k = (deg + 1) * (deg + 1); k < M; k++) OUT(k) = 0.f;
        ddir[0] += dx * dL;
        ddir[1] += dy * dL;
        ddir[2] += dz * dL;
#undef SHC
#undef OUT
    }
    return dnormvdv3(dir_orig, make_float3(ddir[0], ddir[1], ddir[2]));
}

__global__ void __launch_bounds__(256)
geom_backward_kernel(int P, int D, int M,
                     const float* __restrict__ means3D, const int32_t* __restrict__ radii,
                     const float* __restrict__ cov3Ds, const float* __restrict__ shs,
                     const uint8_t* __restrict__ clamped,
                     const float* __restrict__ scales, const float* __restrict__ rotations, float scale_modifier,
                     const float* __restrict__ view, const float* __restrict__ proj, const float* __restrict__ cam_pos,
                     float h_x, float h_y, float tan_fovx, float tan_fovy,
                     const float* __restrict__ ggrad, const float* __restrict__ dL_dcolor,
                     float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dopacity, float* __restrict__ dL_dmask,
                     float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh,
                     float* __restrict__ dL_dscales, float* __restrict__ dL_drots)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;

    const float4 gg0 = *reinterpret_cast<const float4*>(ggrad + (size_t)idx * GG_STRIDE);
    const float4 gg1 = *reinterpret_cast<const float4*>(ggrad + (size_t)idx * GG_STRIDE + 4);
    const bool visible = radii[idx] > 0;

    // outputs that are plain copies of the blend-stage accumulators
    dL_dmeans2D[3 * idx + 0] = visible ? gg0.x : 0.f;
    dL_dmeans2D[3 * idx + 1] = visible ? gg0.y : 0.f;
    dL_dmeans2D[3 * idx + 2] = 0.f;
    dL_dopacity[idx] = visible ? gg1.y : 0.f;
    if (dL_dmask != nullptr) dL_dmask[idx] = visible ? gg1.z : 0.f;

    if (!visible) {
#pragma unroll
        for (int i = 0; i < 3; i++) dL_dmeans3D[3 * idx + i] = 0.f;
#pragma unroll
        for (int i = 0; i < 6; i++) dL_dcov3D[6 * idx + i] = 0.f;
#pragma unroll
        for (int i = 0; i < 3; i++) dL_dscales[3 * idx + i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) dL_drots[4 * idx + i] = 0.f;
        if (dL_dsh != nullptr)
            for (int i = 0; i < 3 * M; i++) dL_dsh[(size_t)idx * 3 * M + i] = 0.f;
        return;
    }

    // ---- conic -> cov2D -> cov3D and the first part of dL/dmean3D (CF backward.cu:144-274) ----
    const float* cov3D = cov3Ds + 6 * (size_t)idx;
    const float3 mean = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    const float3 dL_dconic = make_float3(gg0.z, gg0.w, gg1.x);
    float3 t = xform4x3(mean, view);

    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    const float x_grad_mul = txtz < -limx || txtz > limx ? 0.f : 1.f;
    const float y_grad_mul = tytz < -limy || tytz > limy ? 0.f : 1.f;

    const Mat3 J = mat3_cols(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z),
                             0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z),
                             0.f, 0.f, 0.f);
    const Mat3 Wm = mat3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    const Mat3 Vrk = mat3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    const Mat3 T = mat3_mul(Wm, J);
    Mat3 cov2D = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(Vrk)), T);

    const float a = cov2D.c[0][0] += 0.3f;
    const float b = cov2D.c[0][1];
    const float c = cov2D.c[1][1] += 0.3f;

    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

    float dcov[6];
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
        dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
        dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

        dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
        dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
        dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
        dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
        dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
        dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) dcov[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) dL_dcov3D[6 * idx + i] = dcov[i];

    // gradient w.r.t. the upper 2x3 of T
    const float dL_dT00 = 2 * (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_da +
                          (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_db;
    const float dL_dT01 = 2 * (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_da +
                          (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_db;
    const float dL_dT02 = 2 * (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_da +
                          (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_db;
    const float dL_dT10 = 2 * (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_dc +
                          (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_db;
    const float dL_dT11 = 2 * (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_dc +
                          (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_db;
    const float dL_dT12 = 2 * (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_dc +
                          (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_db;

    // T = W * J
    const float dL_dJ00 = Wm.c[0][0] * dL_dT00 + Wm.c[0][1] * dL_dT01 + Wm.c[0][2] * dL_dT02;
    const float dL_dJ02 = Wm.c[2][0] * dL_dT00 + Wm.c[2][1] * dL_dT01 + Wm.c[2][2] * dL_dT02;
    const float dL_dJ11 = Wm.c[1][0] * dL_dT10 + Wm.c[1][1] * dL_dT11 + Wm.c[1][2] * dL_dT12;
    const float dL_dJ12 = Wm.c[2][0] * dL_dT10 + Wm.c[2][1] * dL_dT11 + Wm.c[2][2] * dL_dT12;

    const float tz = 1.f / t.z;
    const float tz2 = tz * tz;
    const float tz3 = tz2 * tz;

    // gradient w.r.t. the view-space mean; the fov clamp zeroes x/y and adds no d/dt.z term (Appendix A.20)
    const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
    const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
    float3 dmean = xform_vec4x3_transpose(make_float3(dL_dtx, dL_dty, dL_dtz), view);

    // ---- mean2D -> mean3D (CF backward.cu:373-387) ----
    {
        const float4 m_hom = xform4x4(mean, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * m_w * m_w;
        const float gx = gg0.x, gy = gg0.y;
        float3 d2;
        d2.x = (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
        d2.y = (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
        d2.z = (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;
        dmean.x += d2.x;
        dmean.y += d2.y;
        dmean.z += d2.z;
    }

    // ---- SH -> colour gradient (only when SH were the colour source) ----
    if (shs != nullptr) {
        const float3 d3 = sh_backward(idx, D, M, mean, cam_pos, shs, clamped, dL_dcolor, dL_dsh);
        dmean.x += d3.x;
        dmean.y += d3.y;
        dmean.z += d3.z;
    } else if (dL_dsh != nullptr) {
        for (int i = 0; i < 3 * M; i++) dL_dsh[(size_t)idx * 3 * M + i] = 0.f;
    }
    dL_dmeans3D[3 * idx + 0] = dmean.x;
    dL_dmeans3D[3 * idx + 1] = dmean.y;
    dL_dmeans3D[3 * idx + 2] = dmean.z;

    // ---- cov3D -> scale / rotation (CF backward.cu:278-341), only when they were the inputs ----
    if (scales != nullptr) {
        const float4 q = *reinterpret_cast<const float4*>(rotations + 4 * idx);
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const Mat3 R = quat_to_mat3(q);
        const float3 s = make_float3(scale_modifier * scales[3 * idx], scale_modifier * scales[3 * idx + 1],
                                     scale_modifier * scales[3 * idx + 2]);
        Mat3 S = mat3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
        S.c[0][0] = s.x;
        S.c[1][1] = s.y;
        S.c[2][2] = s.z;
        const Mat3 Mm = mat3_mul(S, R);
        const Mat3 dL_dSigma = mat3_cols(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2],
                                         0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                         0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
        Mat3 M2;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) M2.c[i][j] = 2.0f * Mm.c[i][j];
        const Mat3 dL_dM = mat3_mul(M2, dL_dSigma);
        const Mat3 Rt = mat3_transpose(R);
        Mat3 dL_dMt = mat3_transpose(dL_dM);

        dL_dscales[3 * idx + 0] = Rt.c[0][0] * dL_dMt.c[0][0] + Rt.c[0][1] * dL_dMt.c[0][1] + Rt.c[0][2] * dL_dMt.c[0][2];
        dL_dscales[3 * idx + 1] = Rt.c[1][0] * dL_dMt.c[1][0] + Rt.c[1][1] * dL_dMt.c[1][1] + Rt.c[1][2] * dL_dMt.c[1][2];
        dL_dscales[3 * idx + 2] = Rt.c[2][0] * dL_dMt.c[2][0] + Rt.c[2][1] * dL_dMt.c[2][1] + Rt.c[2][2] * dL_dMt.c[2][2];

#pragma unroll
        for (int j = 0; j < 3; j++) {
            dL_dMt.c[0][j] *= s.x;
            dL_dMt.c[1][j] *= s.y;
            dL_dMt.c[2][j] *= s.z;
        }
        float4 dq;
        dq.x = 2 * z * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * y * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) + 2 * x * (dL_dMt.c[1][2] - dL_dMt.c[2][1]);
        dq.y = 2 * y * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * z * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) + 2 * r * (dL_dMt.c[1][2] - dL_dMt.c[2][1]) - 4 * x * (dL_dMt.c[2][2] + dL_dMt.c[1][1]);
        dq.z = 2 * x * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * r * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) + 2 * z * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * y * (dL_dMt.c[2][2] + dL_dMt.c[0][0]);
        dq.w = 2 * r * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * x * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) + 2 * y * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * z * (dL_dMt.c[1][1] + dL_dMt.c[0][0]);
        *reinterpret_cast<float4*>(dL_drots + 4 * idx) = dq;   // no normalisation Jacobian (Appendix A.23)
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) dL_dscales[3 * idx + i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; i++) dL_drots[4 * idx + i] = 0.f;
    }
}

}  // namespace sagars
