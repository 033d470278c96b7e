// tc.cuh -- hand-written tcgen05 / TMEM / mbarrier plumbing (sm_100a) used by the tensor-core blend kernels.
//
// Conventions (validated stand-alone by tools/probes/tcgen05_probe.cu on a B200):
//   * shared-memory matrix descriptors for the SWIZZLE_NONE canonical layouts
//       K-major  operand [rows x k]: 16-byte chunk = 4 tf32 along k;  element (r, k) at
//                (k/4)*LBO + (r/8)*SBO + (r%8)*16 + (k%4)*4          (LBO = byte step between k-chunks, SBO between 8-row groups)
//       MN-major operand [k x n]   : 16-byte chunk = 4 tf32 along n for one k;  element (n, k) at
//                (k/8)*LBO + (n/4)*SBO + (k%8)*16 + (n%4)*4          (LBO = byte step between groups of 8 k, SBO between n-chunks)
//   * instruction descriptor bits: c_format[4,6)=1 (f32), a/b_format[7,10)/[10,13)=2 (tf32), a_major bit 15, b_major bit 16
//     (1 = MN-major), n_dim[17,23) = N>>3, m_dim[24,29) = M>>4;
//   * one thread issues tcgen05.mma; completion is signalled to an mbarrier with tcgen05.commit;
//   * accumulators live in TMEM (row i of an M=128 tile = lane i, column j = column base+j) and are read back with
//     tcgen05.ld.32x32b (warp w may touch lanes 32*(w%4) .. +31 only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cp_async.cuh"
#include "mma.cuh"      // rcp_approx

namespace sagars {
namespace tc {

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell); base_offset = 0, lbo_mode = 0, layout = SWIZZLE_NONE
    return d;
}

// D = f32, A = B = tf32, A K-major, B MN-major (b_mn = 1) or K-major (0)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#if !defined(SAGARS_CUDA_EMU)
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// the same with the A operand (M = 128 rows = TMEM lanes, 8 tf32 = 8 columns per instruction) in tensor memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// make all previously issued MMAs of this thread arrive on `bar` when they have completed
__device__ __forceinline__ void commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival of the calling thread (release semantics at CTA scope: its earlier shared-memory writes are visible to a waiter)
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// wait for a phase that is expected to take long (thousands of cycles): the hardware suspends the thread for up to `ns`
// nanoseconds per attempt, so the waiting warp leaves the issue slots to the warps that share its scheduler (a __nanosleep loop
// was measured to poll every ~17 cycles: a third of all instructions the kernel executed)
__device__ __forceinline__ void mbar_wait_long(uint64_t* bar, uint32_t parity, uint32_t ns = 20000u)
{
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}

// non-blocking test of the same condition (a role that polls several barriers in turn)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// A published counter in shared memory.  Writer: data stores, __syncwarp(), then this store, all from ONE warp; reader: this load,
// then the data loads.  Shared-memory accesses of a warp are performed in program order and there is a single copy of shared
// memory per CTA, so plain volatile accesses (plus compiler barriers) are enough; a formal st.release / ld.acquire pair costs a
// MEMBAR that also waits for the writer's outstanding GLOBAL loads (measured: ~1,000 cycles per publish with prefetches in flight).
__device__ __forceinline__ void st_release_cta(int* p, int v)
{
    asm volatile("" ::: "memory");
    *reinterpret_cast<volatile int*>(p) = v;
    asm volatile("" ::: "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int* p)
{
    const int v = *reinterpret_cast<const volatile int*>(p);
    asm volatile("" ::: "memory");
    return v;
}
__device__ __forceinline__ void backoff(unsigned ns) { __nanosleep(ns); }

// wait for the phase with the given parity to complete (try_wait blocks in hardware for a bounded time per call)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// generic-proxy writes to shared memory -> visible to the tensor-core (async) proxy
__device__ __forceinline__ void fence_smem_to_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// TMEM allocation: one full warp; ncols = power of two >= 32; the base address lands in *slot (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}

// this thread's TMEM lane (row), 32 consecutive columns starting at taddr's column
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v)
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// this thread's TMEM lane (row), 8 consecutive columns starting at taddr's column
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}

// registers -> this thread's TMEM lane, 32 consecutive columns (complete when the function returns)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
                   "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
                   "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
                   "f"(v[30]), "f"(v[31]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// this thread's TMEM lane (row), 16 consecutive columns starting at taddr's column
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// 128-thread named barrier (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync_128(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// "f" constraints on tcgen05.st need b32 registers: floats are fine (bit pattern is stored)

#endif  // SAGARS_CUDA_EMU (tests/cuda_emu/tc_emu.h restates the functions above for the CPU execution shim)

// split an fp32 value into a tf32-exact high part and the remainder (kept to ~22 significant bits by the MMA)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

}  // namespace tc
}  // namespace sagars

#if defined(SAGARS_CUDA_EMU)
#include "tc_emu.h"
#endif
