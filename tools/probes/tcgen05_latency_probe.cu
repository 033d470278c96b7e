// tcgen05_latency_probe.cu -- how long does a chain of small kind::tf32 MMAs take on B200?  For n MMAs (M = 128, K = 8) with N
// columns each, accumulating round-robin into `acc` independent accumulators: cycles from the first issue to the mbarrier arrival.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(128) probe(int n, int N, int acc, int a_tmem, long long* out, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float* A = reinterpret_cast<float*>(smem);                 // 128 x 8 K-major: 2 k-chunks x 16 row groups x 128 B = 4 KB
    float* B = reinterpret_cast<float*>(smem + 8192);          // up to 256 x 8: 2 k-chunks x 32 row groups x 128 B = 8 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long total = 0;
    for (int r = 0; r < reps; r++) {
        long long t0 = clock64();
        if (tid == 0) {
            const uint64_t da = make_desc(smem_u32(A), 2048, 128), db = make_desc(smem_u32(B), 4096, 128);
            for (int i = 0; i < n; i++) {
                const uint32_t d = tmem + 64 + (uint32_t)((i % acc) * N);
                const uint32_t accum = (i >= acc) ? 1u : 0u;
                if (a_tmem)
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }" ::"r"(d), "r"(tmem), "l"(db), "r"(idesc), "r"(accum) : "memory");
                else
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        uint32_t ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(r & 1) : "memory");
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
    printf("%6s %5s %4s %6s %10s %10s\n", "n_mma", "N", "acc", "A", "cycles", "per_mma");
    for (int a_tmem = 0; a_tmem < 2; a_tmem++)
    for (int N : {16, 32, 64, 128})
    for (int acc : {1, 2, 4})
    for (int n : {1, 4, 8, 16, 32}) {
        if (acc * N > 448) continue;
        probe<<<1, 128, 40000>>>(n, N, acc, a_tmem, d, reps);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%6d %5d %4d %6s %10.0f %10.1f %s\n", n, N, acc, a_tmem ? "tmem" : "smem", (double)c / reps, (double)c / reps / n, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
