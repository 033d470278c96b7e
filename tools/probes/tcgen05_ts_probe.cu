// tcgen05_ts_probe.cu -- stand-alone check (B200): kind::tf32 MMA with the A operand in TENSOR MEMORY (written once with tcgen05.st,
// thread = row), B K-major in shared memory: S[128 x 16] = G[128 x 32] * F[16 x 32]^T as 3xTF32 (lo*hi, hi*lo, hi*hi).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ float hi_of(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

constexpr int NPX = 128, NCH = 32, NB = 16, F_LBO = 256, F_BYTES = 8 * F_LBO;

__global__ void __launch_bounds__(128) probe(const float* __restrict__ G, const float* __restrict__ F, float* __restrict__ S_out,
                                             int* __restrict__ status, long long* __restrict__ cyc)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Fh = reinterpret_cast<float*>(smem);
    float* Fl = reinterpret_cast<float*>(smem + F_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * F_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    {
        const int cand = tid & 15, quad = tid >> 4;
        const float* f = F + cand * NCH + 4 * quad;
        float4 h, l;
        h.x = hi_of(f[0]); l.x = f[0] - h.x; h.y = hi_of(f[1]); l.y = f[1] - h.y;
        h.z = hi_of(f[2]); l.z = f[2] - h.z; h.w = hi_of(f[3]); l.w = f[3] - h.w;
        const int off = quad * (F_LBO / 4) + (cand >> 3) * 32 + (cand & 7) * 4;
        *reinterpret_cast<float4*>(Fh + off) = h;
        *reinterpret_cast<float4*>(Fl + off) = l;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t tG = tmem, tS = tmem + 64;       // G_hi at columns [0,32), G_lo at [32,64), S at [64,80)
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // the pixel's gradient row -> its TMEM lane: 32 high parts, then 32 remainders
    {
        uint32_t h[32], l[32];
        for (int k = 0; k < 32; k++) { const float g = G[tid * NCH + k]; const float gh = hi_of(g); h[k] = __float_as_uint(gh); l[k] = __float_as_uint(g - gh); }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(tG + lane_base), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]), "r"(h[8]), "r"(h[9]), "r"(h[10]), "r"(h[11]),
                       "r"(h[12]), "r"(h[13]), "r"(h[14]), "r"(h[15]), "r"(h[16]), "r"(h[17]), "r"(h[18]), "r"(h[19]), "r"(h[20]), "r"(h[21]), "r"(h[22]), "r"(h[23]), "r"(h[24]),
                       "r"(h[25]), "r"(h[26]), "r"(h[27]), "r"(h[28]), "r"(h[29]), "r"(h[30]), "r"(h[31]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                     ::"r"(tG + lane_base + 32), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]), "r"(l[8]), "r"(l[9]), "r"(l[10]), "r"(l[11]),
                       "r"(l[12]), "r"(l[13]), "r"(l[14]), "r"(l[15]), "r"(l[16]), "r"(l[17]), "r"(l[18]), "r"(l[19]), "r"(l[20]), "r"(l[21]), "r"(l[22]), "r"(l[23]), "r"(l[24]),
                       "r"(l[25]), "r"(l[26]), "r"(l[27]), "r"(l[28]), "r"(l[29]), "r"(l[30]), "r"(l[31]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr uint32_t ID = idesc_tf32(128, NB);
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t fh = smem_u32(Fh), fl = smem_u32(Fl);
        int first = 1;
        for (int term = 0; term < 3; term++) {
            const uint32_t a0 = tG + (term == 0 ? 32u : 0u);
            const uint32_t b0 = (term == 1) ? fl : fh;
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t db = make_desc(b0 + ks * 2 * F_LBO, F_LBO, 128);
                const uint32_t acc = first ? 0u : 1u; first = 0;
                asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p; }"
                             ::"r"(tS), "r"(a0 + ks * 8), "l"(db), "r"(ID), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    bool ok = false;
    for (int i = 0; i < 20000000 && !ok; i++) {
        uint32_t r;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(r) : "r"(smem_u32(bar)), "r"(0) : "memory");
        ok = r != 0;
    }
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) status[0] = 1; }
    else {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tS + lane_base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < 16; n++) S_out[tid * 16 + n] = __uint_as_float(v[n]);
        if (tid == 0) { status[0] = 0; cyc[0] = t1 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main()
{
    srand(11);
    float *G = new float[NPX * NCH], *F = new float[NB * NCH];
    for (int i = 0; i < NPX * NCH; i++) G[i] = (rand() / (float)RAND_MAX - 0.5f) * 1e-6f;
    for (int i = 0; i < NB * NCH; i++) F[i] = (rand() / (float)RAND_MAX - 0.5f);
    float *dG, *dF, *dS; int* dSt; long long* dC;
    cudaMalloc(&dG, NPX * NCH * 4); cudaMalloc(&dF, NB * NCH * 4); cudaMalloc(&dS, NPX * 16 * 4); cudaMalloc(&dSt, 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dG, G, NPX * NCH * 4, cudaMemcpyHostToDevice); cudaMemcpy(dF, F, NB * NCH * 4, cudaMemcpyHostToDevice);
    int st = -1; cudaMemcpy(dSt, &st, 4, cudaMemcpyHostToDevice); cudaMemset(dS, 0, NPX * 16 * 4);
    const int smem = 2 * F_BYTES + 64;
    probe<<<1, 128, smem>>>(dG, dF, dS, dSt, dC);
    cudaError_t e = cudaDeviceSynchronize();
    float* S = new float[NPX * 16]; long long cyc = 0;
    cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost); cudaMemcpy(S, dS, NPX * 16 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double e1 = 0, m1 = 0;
    for (int p = 0; p < NPX; p++) for (int j = 0; j < NB; j++) {
        double s = 0; for (int c = 0; c < NCH; c++) s += (double)G[p * NCH + c] * F[j * NCH + c];
        e1 = fmax(e1, fabs(S[p * 16 + j] - s)); m1 = fmax(m1, fabs(s));
    }
    printf("A-from-TMEM tf32x3: cuda=%s status=%d rel err %.2e cycles %lld  %s\n", cudaGetErrorString(e), st, e1 / m1, cyc,
           (e == cudaSuccess && st == 0 && e1 < 2e-5 * m1) ? "TS OK" : "TS WRONG");
    if (!(e1 < 2e-5 * m1)) printf(" S[0][0..3] %g %g %g %g | S[5][0..3] %g %g %g %g\n", S[0], S[1], S[2], S[3], S[80], S[81], S[82], S[83]);
    return 0;
}
