// tcgen05_probe.cu -- stand-alone check of the hand-built tcgen05 plumbing before it goes into the blend kernels:
// smem matrix descriptors (K-major, SWIZZLE_NONE canonical layout), instruction descriptor, TMEM alloc / ld,
// commit -> mbarrier.  D[128 x 32] (fp32, TMEM) = A[128 x K] * B[32 x K]^T with bf16 operands written by threads.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int max_spins)
{
    for (int i = 0; i < max_spins; i++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

constexpr int M = 128, N = 32, KT = 64;   // K tile = 64 -> 4 MMAs of K=16 (bf16)

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                    int* __restrict__ status, int reps, long long* cycles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);                      // [KT/8][M/8][8][8] = 16 KB
    __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + M * KT * 2);         // [KT/8][N/8][8][8] = 4 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + M * KT * 2 + N * KT * 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t SBO = 128, LBO_A = 128 * (M / 8), LBO_B = 128 * (N / 8);

    // thread m writes row m of A; threads 0..31 also write row n of B
    for (int k = 0; k < KT; k++) {
        sA[(k / 8) * (LBO_A / 2) + (tid / 8) * (SBO / 2) + (tid % 8) * 8 + (k % 8)] = __float2bfloat16(A[tid * KT + k]);
        if (tid < N) sB[(k / 8) * (LBO_B / 2) + (tid / 8) * (SBO / 2) + (tid % 8) * 8 + (k % 8)] = __float2bfloat16(B[tid * KT + k]);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // make the generic-proxy smem writes visible to the async (tensor core) proxy, then sync the CTA
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // instruction descriptor: D=f32, A=B=bf16, both K-major, N=32, M=128
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    long long t0 = clock64();
    uint32_t parity = 0;
    bool ok = true;
    for (int r = 0; r < reps && ok; r++) {
        if (tid == 0) {
#pragma unroll
            for (int ks = 0; ks < KT / 16; ks++) {
                const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO_B, LBO_B, SBO);
                const uint32_t accum = (ks > 0) ? 1u : 0u;
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                             ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        ok = mbar_wait(bar, parity, 20000000);
        parity ^= 1;
    }
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) {
        if (tid == 0) status[0] = 1;   // timed out waiting for the MMA commit
    } else {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < N; n++) D[tid * N + n] = __uint_as_float(v[n]);
        if (tid == 0) { status[0] = 0; cycles[0] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
}

// ---- variant 2: kind::tf32, 3-term split (hi/lo), A K-major [128 x 16], B MN-major (feature rows [k][n]) ----
constexpr int K2 = 16;   // instances per batch -> 2 MMAs of K=8 per term
template <int BMN>
__global__ void __launch_bounds__(128) probe_tf32_kernel(const float* __restrict__ A, const float* __restrict__ F, float* __restrict__ D,
                                                         int* __restrict__ status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    // A tiles (hi, lo): K-major, 16-byte chunk = 4 tf32 along k; [k/4][m/8][m%8][4]: SBO = 128, LBO = 128 * 16 = 2048
    float* sAh = reinterpret_cast<float*>(smem);
    float* sAl = sAh + M * K2;
    // B tiles (hi, lo): MN-major, 16-byte chunk = 4 tf32 along n for one k; [k/8][n/4][k%8][4]: SBO = 128 (next n-chunk), LBO = 8*128 (next 8 k)
    float* sBh = sAl + M * K2;
    float* sBl = sBh + N * K2;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + N * K2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t A_SBO = 128, A_LBO = 128 * (M / 8);
    constexpr uint32_t B_SBO = 128, B_LBO = BMN ? 128 * (N / 4) : 128 * (N / 8);

    for (int k = 0; k < K2; k++) {
        const float x = A[tid * K2 + k];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const int off = (k / 4) * (A_LBO / 4) + (tid / 8) * (A_SBO / 4) + (tid % 8) * 4 + (k % 4);
        sAh[off] = h;
        sAl[off] = x - h;
    }
    for (int e = tid; e < K2 * N; e += 128) {   // F is [k][n] row-major (a feature row per instance)
        const int k = e / N, n = e % N;
        const float x = F[e];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const int off = BMN ? (k / 8) * (B_LBO / 4) + (n / 4) * (B_SBO / 4) + (k % 8) * 4 + (n % 4)
                            : (k / 4) * (B_LBO / 4) + (n / 8) * (B_SBO / 4) + (n % 8) * 4 + (k % 4);
        sBh[off] = h;
        sBl[off] = x - h;
    }
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), A K-major (bit15=0), B MN-major (bit16=1), N, M
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)BMN << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        int first = 1;
        for (int term = 0; term < 3; term++) {            // lo*hi, hi*lo, hi*hi
            const float* a = (term == 0) ? sAl : sAh;
            const float* b = (term == 1) ? sBl : sBh;
            for (int ks = 0; ks < K2 / 8; ks++) {
                const uint64_t da = make_desc(smem_u32(a) + ks * 2 * A_LBO, A_LBO, A_SBO);
                const uint64_t db = make_desc(smem_u32(b) + (BMN ? ks * B_LBO : ks * 2 * B_LBO), B_LBO, B_SBO);
                const uint32_t accum = first ? 0u : 1u;
                first = 0;
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
                             ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    const bool ok = mbar_wait(bar, 0, 20000000);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) status[0] = 1; }
    else {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < N; n++) D[tid * N + n] = __uint_as_float(v[n]);
        if (tid == 0) status[0] = 0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32) : "memory");
}

int main()
{
    float *hA = new float[M * KT], *hB = new float[N * KT], *hD = new float[M * N], *ref = new float[M * N];
    srand(3);
    auto bf = [](float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000u; float y; memcpy(&y, &u, 4); return y; };
    for (int i = 0; i < M * KT; i++) hA[i] = bf((rand() / (float)RAND_MAX - 0.5f) * 2.f);
    for (int i = 0; i < N * KT; i++) hB[i] = bf((rand() / (float)RAND_MAX - 0.5f) * 2.f);
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double s = 0; for (int k = 0; k < KT; k++) s += (double)hA[m * KT + k] * hB[n * KT + k]; ref[m * N + n] = (float)s; }
    float *dA, *dB, *dD; int* dS; long long* dC;
    cudaMalloc(&dA, M * KT * 4); cudaMalloc(&dB, N * KT * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dS, 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4); int st = -1; cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
    const int smem = M * KT * 2 + N * KT * 2 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int reps : {1, 1000}) {
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dS, reps, dC);
        cudaError_t e = cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0; int bad = 0;
        for (int i = 0; i < M * N; i++) { double d = fabs(hD[i] - ref[i]); if (d > err) err = d; if (fabs(ref[i]) > mx) mx = fabs(ref[i]); if (d > 1e-3) bad++; }
        printf("reps=%d: cuda=%s status=%d  max|ref|=%.3f max err=%.3e bad=%d  %s   cycles/rep=%.1f (4 MMAs 128x32x16 + commit + wait)\n", reps,
               cudaGetErrorString(e), st, mx, err, bad, (e == cudaSuccess && st == 0 && bad == 0) ? "TCGEN05 OK" : "TCGEN05 WRONG", (double)cyc / reps);
        if (bad) { printf(" D[0][0..3]= %f %f %f %f  ref %f %f %f %f\n", hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
                   printf(" D[1][0..3]= %f %f %f %f  ref %f %f %f %f\n", hD[32], hD[33], hD[34], hD[35], ref[32], ref[33], ref[34], ref[35]); }
    }
    {   // variant 2
        float *hA2 = new float[M * K2], *hF = new float[K2 * N], *ref2 = new float[M * N];
        for (int i = 0; i < M * K2; i++) hA2[i] = (rand() / (float)RAND_MAX) * 0.9f;
        for (int i = 0; i < K2 * N; i++) hF[i] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
        for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double s2 = 0; for (int k = 0; k < K2; k++) s2 += (double)hA2[m * K2 + k] * hF[k * N + n]; ref2[m * N + n] = (float)s2; }
        float *dA2, *dF;
        cudaMalloc(&dA2, M * K2 * 4); cudaMalloc(&dF, K2 * N * 4);
        cudaMemcpy(dA2, hA2, M * K2 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dF, hF, K2 * N * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, M * N * 4); st = -1; cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
        const int smem2 = (2 * M * K2 + 2 * N * K2) * 4 + 64;
        for (int bmn = 0; bmn < 2; bmn++) {
        cudaMemset(dD, 0, M * N * 4); st = -1; cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
        if (bmn) { cudaFuncSetAttribute(probe_tf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2); probe_tf32_kernel<1><<<1, 128, smem2>>>(dA2, dF, dD, dS); }
        else { cudaFuncSetAttribute(probe_tf32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2); probe_tf32_kernel<0><<<1, 128, smem2>>>(dA2, dF, dD, dS); }
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
        double err = 0, mx = 0;
        for (int i = 0; i < M * N; i++) { double d = fabs(hD[i] - ref2[i]); if (d > err) err = d; if (fabs(ref2[i]) > mx) mx = fabs(ref2[i]); }
        printf("tf32x3 (A K-major, B %s): cuda=%s status=%d max|ref|=%.3f max err=%.3e  %s\n", bmn ? "MN-major" : "K-major", cudaGetErrorString(e), st, mx, err,
               (e == cudaSuccess && st == 0 && err < 2e-5) ? "TF32X3 OK" : "TF32X3 WRONG");
        if (err >= 2e-5) { printf(" D[0][0..3]= %f %f %f %f  ref %f %f %f %f\n", hD[0], hD[1], hD[2], hD[3], ref2[0], ref2[1], ref2[2], ref2[3]);
                           printf(" D[5][0..3]= %f %f %f %f  ref %f %f %f %f\n", hD[160], hD[161], hD[162], hD[163], ref2[160], ref2[161], ref2[162], ref2[163]); }
        }
    }
    return 0;
}
