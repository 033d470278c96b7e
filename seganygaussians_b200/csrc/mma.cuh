// mma.cuh -- warp-level TF32 tensor-core helpers (mma.sync m16n8k8) shared by the backward blend kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(SAGARS_CUDA_EMU)
#include "mma_emu.h"   // CPU execution shim of the test suite (tests/cuda_emu/): the same two entry points on the host
#else

namespace sagars {

// x = hi + lo with hi exact in tf32 (truncation) and lo = x - hi exact in fp32; the tensor core reads the top 19 bits
// of lo, so hi*hi' + hi*lo' + lo*hi' carries ~2^-21 relative error (cvt.rna.tf32 would cost ~5 instructions each)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// D(16x8, f32) += A(16x8, tf32, row) * B(8x8, tf32, col)
// fragments (g = lane / 4, t = lane % 4): a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4); b0=(k=t,n=g) b1=(k=t+4,n=g);
// d0,d1=(g, 2t, 2t+1) d2,d3=(g+8, 2t, 2t+1)
__device__ __forceinline__ void mma_16n8k8(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// reciprocal without the IEEE slow path (__frcp_rn / division compile to MUFU.RCP plus a BRANCH to a denormal handler, which breaks
// up a sequence of otherwise independent chains): one MUFU.RCP, <= 1 ulp
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

}  // namespace sagars

#endif  // SAGARS_CUDA_EMU
