// tcgen05_latency_probe2.cu -- follow-up: (a) do MMA chains issued by DIFFERENT warps of one CTA overlap?  (b) kind::f16 vs tf32,
// (c) M = 64, (d) N = 256.  n MMAs per issuing warp; `issuers` warps issue concurrently into their own accumulators and mbarriers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(128) probe(int n, int M, int N, int issuers, int f16, long long* out, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192; i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
    if (tid == 0) { for (int w = 0; w < 4; w++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[w])), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t fmt = f16 ? 1u : 2u;     // a/b format: 1 = bf16 (kind::f16), 2 = tf32
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    long long total = 0;
    for (int r = 0; r < reps; r++) {
        __syncthreads();
        long long t0 = clock64();
        if (lane == 0 && warp < issuers) {
            const uint64_t da = make_desc(smem_u32(smem), 2048, 128), db = make_desc(smem_u32(smem + 8192), 4096, 128);
            const uint32_t d = tmem + (uint32_t)(warp * (512 / issuers));
            for (int i = 0; i < n; i++) {
                const uint32_t accum = (i > 0) ? 1u : 0u;
                if (f16) asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
                else asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
        }
        for (int w = 0; w < issuers; w++) {
            uint32_t ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar[w])), "r"(r & 1) : "memory");
        }
        total += clock64() - t0;
    }
    if (tid == 0) out[0] = total;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main()
{
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    const int reps = 200;
    printf("%6s %4s %5s %8s %6s %10s %12s\n", "n_mma", "M", "N", "issuers", "kind", "cycles", "per_mma/issuer");
    struct C { int M, N, issuers, f16; } cs[] = {{128, 32, 1, 0}, {128, 32, 2, 0}, {128, 32, 4, 0}, {128, 128, 4, 0}, {64, 32, 1, 0}, {128, 256, 1, 0},
                                                 {128, 32, 1, 1}, {128, 128, 1, 1}, {128, 256, 1, 1}, {128, 32, 4, 1}};
    for (auto c : cs)
        for (int n : {4, 16, 32}) {
            if (c.issuers * c.N > 512) continue;
            probe<<<1, 128, 40000>>>(n, c.M, c.N, c.issuers, c.f16, d, reps);
            cudaError_t e = cudaDeviceSynchronize();
            long long cy = 0; cudaMemcpy(&cy, d, 8, cudaMemcpyDeviceToHost);
            printf("%6d %4d %5d %8d %6s %10.0f %12.1f %s\n", n, c.M, c.N, c.issuers, c.f16 ? "bf16" : "tf32", (double)cy / reps, (double)cy / reps / n,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
