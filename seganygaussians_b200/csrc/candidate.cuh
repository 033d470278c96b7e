// candidate.cuh -- block-level candidate test shared by the blend kernels.
//
// A (pixel, splat) pair contributes only if power = -Q(d) >= accept_threshold (math.cuh: a conservative restatement of
// the reference's alpha >= 1/255 test, CF forward.cu:336-345 / backward.cu:470-480), with
//     Q(d) = 0.5 * (a dx^2 + c dy^2) + b dx dy,   d = splat centre - pixel,   (a, b, c) = conic.
// For a convex Q the minimum over a rectangle of pixel centres is 0 when the splat centre lies inside, otherwise it sits
// on one of the four edges, where Q is a 1-D parabola whose minimiser is clamped to the edge.  One lane evaluates this
// for one splat, so a warp rejects 32 (block, splat) pairs per instruction instead of one.  The test is conservative
// (continuous rectangle instead of the pixel lattice, relative + absolute margin, non-convex conics never rejected);
// the per-pixel test that follows is the reference's own and decides what is blended, so results do not change.
#pragma once
#include <cuda_runtime.h>
#include "mma.cuh"      // rcp_approx

namespace sagars {

// g0 = (x, y, conic.x, conic.y), g1 = (conic.z, opacity, accept_threshold, -); rectangle [bx0, bx1] x [by0, by1]
// in pixel-centre coordinates.  Returns true when NO point of the rectangle can reach the accept threshold.
// FAST: branch-free with rcp.approx instead of __frcp_rn (whose denormal slow path is a branch) -- a minimiser that is off by an
// ulp moves the edge minimum in second order only, far inside the margins below.  Measured: faster in the backward warp kernel;
// in the forward warp kernel (32 accumulator registers live) it makes nvcc spill 104 bytes and costs 10 %, so that one keeps the
// branchy form.
template <bool FAST = false>
__device__ __forceinline__ bool block_rejects(const float4 g0, const float4 g1, float bx0, float bx1, float by0, float by1)
{
    const float ca = g0.z, cb = g0.w, cc = g1.x, thr = g1.z;
    const float dxl = g0.x - bx1, dxh = g0.x - bx0, dyl = g0.y - by1, dyh = g0.y - by0;   // range of d = centre - pixel
    const bool inside = dxl <= 0.f && dxh >= 0.f && dyl <= 0.f && dyh >= 0.f;
    float qmin = 0.f;
    if (FAST || !inside) {
        const float ia = FAST ? rcp_approx(ca) : __frcp_rn(ca), ic = FAST ? rcp_approx(cc) : __frcp_rn(cc);
        const float x0 = fminf(fmaxf(-cb * dyl * ia, dxl), dxh);
        const float x1 = fminf(fmaxf(-cb * dyh * ia, dxl), dxh);
        const float y0 = fminf(fmaxf(-cb * dxl * ic, dyl), dyh);
        const float y1 = fminf(fmaxf(-cb * dxh * ic, dyl), dyh);
        const float e0 = 0.5f * (ca * x0 * x0 + cc * dyl * dyl) + cb * x0 * dyl;
        const float e1 = 0.5f * (ca * x1 * x1 + cc * dyh * dyh) + cb * x1 * dyh;
        const float e2 = 0.5f * (ca * dxl * dxl + cc * y0 * y0) + cb * dxl * y0;
        const float e3 = 0.5f * (ca * dxh * dxh + cc * y1 * y1) + cb * dxh * y1;
        const float q = fminf(fminf(e0, e1), fminf(e2, e3));
        qmin = inside ? 0.f : q;
    }
    const bool convex = ca > 0.f && cc > 0.f && ca * cc - cb * cb > 0.f;
    // sentinel records (threshold = +inf) and NaN thresholds can never be accepted by the per-pixel test either
    return (convex && (qmin * (1.f - 1e-4f) - 1e-4f > -thr)) || !(thr < __int_as_float(0x7f800000));
}

}  // namespace sagars
