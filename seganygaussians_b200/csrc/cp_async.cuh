// cp_async.cuh -- thin wrappers over the asynchronous global->shared copy instructions used to stage
// per-tile instance lists (LDGSTS) and, for contiguous runs, bulk copies completed on an mbarrier.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(SAGARS_CUDA_EMU)
// CPU execution shim of the test suite (tests/cuda_emu/): the same entry points with the instructions' semantics restated on
// the host -- copies land as LATE as the programming model allows, so a read before the matching wait shows up as a failure.
#include "cp_async_emu.h"
#else

namespace sagars {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte asynchronous copy, L2 only (streaming data: each row is used by one CTA)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// 4-byte asynchronous copy (list entries prefetched many chunks ahead)
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- bulk asynchronous copies (the TMA engine's non-tensor form) completing on an mbarrier ----
// One instruction moves `bytes` (a multiple of 16; source and destination 16-byte aligned) from global to shared memory;
// the mbarrier's transaction count drops by `bytes` when the data has landed.  Used to GATHER per-instance rows
// (32-byte records, K*4-byte feature rows): one bulk copy per row, no per-thread address arithmetic or register staging.
__device__ __forceinline__ void mbarrier_init(uint64_t* bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the initialised barrier (and earlier generic-proxy writes to shared memory) visible to the async proxy
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbarrier_arrive_expect_tx(uint64_t* bar, uint32_t tx_bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Wait for the phase of the given parity.  A wait that does not end within ~2^26 polls (seconds; a copy that can never
// complete, e.g. a misaligned row) traps: a CUDA error the caller sees instead of a hung GPU.
__device__ __forceinline__ void mbarrier_wait_parity(uint64_t* bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; spins++) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && spins > (1u << 26)) __trap();
    }
}

// vectorised fire-and-forget global reductions (sm_90+): one instruction adds 4 / 2 floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};\n" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a)
{
    asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(a) : "memory");
}

}  // namespace sagars

#endif  // SAGARS_CUDA_EMU
