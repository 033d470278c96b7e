// math.cuh -- small fp32 helpers of the per-Gaussian stages.
//
// The reference evaluates its 3x3 algebra through glm (column-major `mat3`, product
// R[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2], evaluated left to right,
// third_party/glm/glm/detail/type_mat3x3.inl:486-518).  Radii and tile rectangles are integer
// functions of these fp32 results, so `Mat3` below keeps that exact association (including the
// products with structural zeros: dropping a `x*0` term changes which product nvcc fuses into an FMA
// and therefore the last bit).  SURVEY.md Appendix A.2-A.8 lists the semantics being reproduced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sagars {

// column-major 3x3: c[col][row]
struct Mat3 {
    float c[3][3];
};

__device__ __forceinline__ Mat3 mat3_cols(float a0, float a1, float a2, float b0, float b1, float b2,
                                          float c0, float c1, float c2)
{
    Mat3 m;
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}

__device__ __forceinline__ Mat3 mat3_mul(const Mat3& A, const Mat3& B)
{
    Mat3 R;
#pragma unroll
    for (int col = 0; col < 3; col++)
#pragma unroll
        for (int row = 0; row < 3; row++)
            R.c[col][row] = A.c[0][row] * B.c[col][0] + A.c[1][row] * B.c[col][1] + A.c[2][row] * B.c[col][2];
    return R;
}

__device__ __forceinline__ Mat3 mat3_transpose(const Mat3& A)
{
    Mat3 R;
#pragma unroll
    for (int col = 0; col < 3; col++)
#pragma unroll
        for (int row = 0; row < 3; row++) R.c[col][row] = A.c[row][col];
    return R;
}

// 4x4 (row-vector convention, matrix stored transposed, read column-major): CF auxiliary.h:58-77
__device__ __forceinline__ float3 xform4x3(const float3& p, const float* __restrict__ m)
{
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float4 xform4x4(const float3& p, const float* __restrict__ m)
{
    return make_float4(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
                       m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]);
}
// transpose of the upper-left 3x3 applied to a vector (CF auxiliary.h:89-97)
__device__ __forceinline__ float3 xform_vec4x3_transpose(const float3& p, const float* __restrict__ m)
{
    return make_float3(m[0] * p.x + m[1] * p.y + m[2] * p.z,
                       m[4] * p.x + m[5] * p.y + m[6] * p.z,
                       m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

// NDC -> pixel.  The reference's literals are doubles, so this is evaluated in fp64 and rounded
// once (CF auxiliary.h:41-44).
__device__ __forceinline__ float ndc_to_pix(float v, int S)
{
    return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}

// rotation matrix of an (unnormalised) quaternion q = (r, x, y, z), columns as in CF forward.cu:137-141
__device__ __forceinline__ Mat3 quat_to_mat3(const float4& q)
{
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    return mat3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                     2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                     2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
}

// Sigma = (S R)^T (S R), upper triangle (CF forward.cu:121-155; Appendix A.3)
__device__ __forceinline__ void cov3d_from_scale_rot(const float3& scale, float mod, const float4& q, float* out6)
{
    Mat3 S = mat3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
    S.c[0][0] = mod * scale.x;
    S.c[1][1] = mod * scale.y;
    S.c[2][2] = mod * scale.z;
    const Mat3 R = quat_to_mat3(q);
    const Mat3 Mm = mat3_mul(S, R);
    const Mat3 Sigma = mat3_mul(mat3_transpose(Mm), Mm);
    out6[0] = Sigma.c[0][0];
    out6[1] = Sigma.c[0][1];
    out6[2] = Sigma.c[0][2];
    out6[3] = Sigma.c[1][1];
    out6[4] = Sigma.c[1][2];
    out6[5] = Sigma.c[2][2];
}

// EWA projection of the 3D covariance (CF forward.cu:77-116; Appendix A.4). Returns (a, b, c) of
// [[a b][b c]] with the 0.3 low-pass already added.
__device__ __forceinline__ float3 cov2d_project(const float3& mean, float focal_x, float focal_y,
                                                float tan_fovx, float tan_fovy, const float* cov3D,
                                                const float* __restrict__ view)
{
    float3 t = xform4x3(mean, view);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;

    const Mat3 J = mat3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
                             0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
                             0.f, 0.f, 0.f);
    const Mat3 Wm = mat3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
    const Mat3 T = mat3_mul(Wm, J);
    const Mat3 Vrk = mat3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    Mat3 cov = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(Vrk)), T);
    cov.c[0][0] += 0.3f;
    cov.c[1][1] += 0.3f;
    return make_float3(cov.c[0][0], cov.c[0][1], cov.c[1][1]);
}

// tile rectangle of a splat (CF auxiliary.h:46-56; Appendix A.8): float arithmetic, C truncation
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, uint2& rmin, uint2& rmax,
                                          int tiles_x, int tiles_y)
{
    rmin.x = (unsigned)min(tiles_x, max(0, (int)((p.x - max_radius) / SAGARS_TILE_X)));
    rmin.y = (unsigned)min(tiles_y, max(0, (int)((p.y - max_radius) / SAGARS_TILE_Y)));
    rmax.x = (unsigned)min(tiles_x, max(0, (int)((p.x + max_radius + SAGARS_TILE_X - 1) / SAGARS_TILE_X)));
    rmax.y = (unsigned)min(tiles_y, max(0, (int)((p.y + max_radius + SAGARS_TILE_Y - 1) / SAGARS_TILE_Y)));
}

// Conservative lower bound on `power` below which a (pixel, splat) pair is certainly rejected by the blend
// kernels' exact test  alpha = min(0.99, opacity * expf(power)) >= 1/255 :
//     opacity * exp(power) >= 1/255   <=>   power >= -ln(255 * opacity),
// so any pair with power < -ln(255*opacity) - margin is skipped without evaluating expf.  The margin (1e-4
// absolute + 1e-5 relative) dwarfs the error of logf here and of expf / the product rounding there (~1e-6), and the
// blend kernels compare the SAME fp32 `power` value they would exponentiate, so no conditioning enters.
// +inf = the splat can never be accepted (opacity <= 1/255); -inf = never skip (NaN opacity).
// This is purely an acceleration structure: skipped pairs are exactly pairs the reference skips with
// `continue` (CF forward.cu:340-349), so results are unchanged.
__device__ __forceinline__ float accept_threshold(float opacity)
{
    const float inf = __int_as_float(0x7f800000);
    if (!(opacity == opacity)) return -inf;
    if (!(opacity * 255.0f > 1.0f)) return inf;
    const float t = logf(255.0f * opacity);          // > 0
    return -(t + 1e-5f * t + 1e-4f);
}

// real spherical-harmonics constants (CF auxiliary.h:22-39)
#define SAGARS_SH_C0 0.28209479177387814f
#define SAGARS_SH_C1 0.4886025119029199f
__device__ __constant__ const float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                                  -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ const float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                                  0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                                  -0.5900435899266435f};

// view-dependent colour of one Gaussian from its SH coefficients, +0.5, clamped at 0
// (CF forward.cu:23-74; Appendix X4).  `sh` is [P][M][3].
__device__ __forceinline__ float3 sh_to_rgb(int idx, int deg, int M, const float3& pos, const float* __restrict__ cam_pos,
                                            const float* __restrict__ shs, bool* clamped)
{
    float3 dir = make_float3(pos.x - cam_pos[0], pos.y - cam_pos[1], pos.z - cam_pos[2]);
    const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir = make_float3(dir.x / len, dir.y / len, dir.z / len);
    const float* sh = shs + (size_t)idx * M * 3;
    float res[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
#define SHC(k) sh[(k) * 3 + ch]
        float result = SAGARS_SH_C0 * SHC(0);
        if (deg > 0) {
            const float x = dir.x, y = dir.y, z = dir.z;
            result = result - SAGARS_SH_C1 * y * SHC(1) + SAGARS_SH_C1 * z * SHC(2) - SAGARS_SH_C1 * x * SHC(3);
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z;
                const float xy = x * y, yz = y * z, xz = x * z;
                result = result + kSH_C2[0] * xy * SHC(4) + kSH_C2[1] * yz * SHC(5) +
                         kSH_C2[2] * (2.0f * zz - xx - yy) * SHC(6) + kSH_C2[3] * xz * SHC(7) +
                         kSH_C2[4] * (xx - yy) * SHC(8);
                if (deg > 2) {
                    result = result + kSH_C3[0] * y * (3.0f * xx - yy) * SHC(9) + kSH_C3[1] * xy * z * SHC(10) +
                             kSH_C3[2] * y * (4.0f * zz - xx - yy) * SHC(11) +
                             kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SHC(12) +
                             kSH_C3[4] * x * (4.0f * zz - xx - yy) * SHC(13) + kSH_C3[5] * z * (xx - yy) * SHC(14) +
                             kSH_C3[6] * x * (xx - 3.0f * yy) * SHC(15);
                }
            }
        }
#undef SHC
        result += 0.5f;
        clamped[ch] = (result < 0);
        res[ch] = fmaxf(result, 0.0f);
    }
    return make_float3(res[0], res[1], res[2]);
}

}  // namespace sagars
