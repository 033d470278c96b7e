// tcgen05_bwd_probe.cu -- stand-alone check (B200) of the operand layouts the tcgen05 backward blend kernel uses, before they
// go into the kernel:
//   GEMM1  S[128 px x 16 cand]   = G[128 x 32 ch] * F[16 x 32]^T        3xTF32, A K-major with a NON-dense 8-row-group stride
//   GEMM3  D[128 rows x 64]      = A3[rows x 128 px] * B3[64 x 128 px]^T  A3 = the SAME shared-memory tile read MN-major:
//          rows 0..31 = G_hi channels, 32..63 = G_lo channels, 64..71 = moment basis X, 72..127 = whatever follows in shared
//          memory (their accumulator rows are never read);  B3 = [W_hi | W_lo | Q_hi | Q_lo] K-major, padded k-chunk stride.
// Prints max errors against a double-precision host product and the issue->completion latency of both GEMMs.
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
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int max_spins)
{
    for (int i = 0; i < max_spins; i++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ float hi_of(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

constexpr int NPX = 128, NCH = 32, NB = 16;
constexpr int G_CHUNKS = 18;                        // 8 hi + 8 lo + 2 basis chunks of 4 rows
constexpr int G_PG = G_CHUNKS * 128;                // bytes per group of 8 pixels
constexpr int G_BYTES = (NPX / 8) * G_PG + 14 * 128;   // + slack read by the unused accumulator rows of the last k-groups
constexpr int B3_LBO = 1040, B3_BYTES = 32 * B3_LBO;   // k-chunk (4 px) stride, padded against bank conflicts
constexpr int F_LBO = 256, F_BYTES = 8 * F_LBO;        // per hi / lo

__global__ void __launch_bounds__(128) probe(const float* __restrict__ G, const float* __restrict__ X, const float* __restrict__ F,
                                             const float* __restrict__ WQ, float* __restrict__ S_out, float* __restrict__ D_out,
                                             int* __restrict__ status, long long* __restrict__ cyc, int reps, int variant)
{
    // variant 0: as designed (A3 = the G tile read MN-major, overlapping row chunks, NaN slack)
    //         1: same, slack zeroed            2: A3 = a separate K-major copy [px/4][row/8][row%8][4]  (rows 0..71 valid, rest zero)
    //         3: MN-major, LBO / SBO swapped   4: MN-major, M = 64 instruction (rows 0..63 only)
    //         5: MN-major from a dense copy [px/8][32 chunks][px%8][4] (no overlap)
    //         6: TRANSPOSED product: D'[128 x 80] = A = the w/q tile (K-major, rows 64..127 alias) x B = the G tile read MN-major (N = 80)
    //         7: as 6 with the dense copy of variant 5 as B
    extern __shared__ __align__(1024) unsigned char smem[];
    float* Gt = reinterpret_cast<float*>(smem);
    float* B3 = reinterpret_cast<float*>(smem + G_BYTES);
    float* Fh = reinterpret_cast<float*>(smem + G_BYTES + B3_BYTES);
    float* Fl = reinterpret_cast<float*>(smem + G_BYTES + B3_BYTES + F_BYTES);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + G_BYTES + B3_BYTES + 2 * F_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    float* A3 = reinterpret_cast<float*>(smem + G_BYTES + B3_BYTES + 2 * F_BYTES + 64);   // 64 KB: K-major or dense MN-major copy
    const int tid = threadIdx.x, warp = tid >> 5;

    // G tile: thread = pixel
    {
        float* base = Gt + (tid >> 3) * (G_PG / 4) + (tid & 7) * 4;
        for (int c = 0; c < 8; c++) {
            float4 h, l;
            const float* g = G + tid * NCH + 4 * c;
            h.x = hi_of(g[0]); l.x = g[0] - h.x; h.y = hi_of(g[1]); l.y = g[1] - h.y;
            h.z = hi_of(g[2]); l.z = g[2] - h.z; h.w = hi_of(g[3]); l.w = g[3] - h.w;
            *reinterpret_cast<float4*>(base + c * 32) = h;
            *reinterpret_cast<float4*>(base + (8 + c) * 32) = l;
        }
        *reinterpret_cast<float4*>(base + 16 * 32) = make_float4(X[tid * 8 + 0], X[tid * 8 + 1], X[tid * 8 + 2], X[tid * 8 + 3]);
        *reinterpret_cast<float4*>(base + 17 * 32) = make_float4(X[tid * 8 + 4], X[tid * 8 + 5], X[tid * 8 + 6], X[tid * 8 + 7]);
        for (int i = tid; i < 14 * 32; i += 128) Gt[(NPX / 8) * (G_PG / 4) + i] = (variant == 0) ? __int_as_float(0x7fc00000) : 0.f;
        for (int i = tid; i < 16384; i += 128) A3[i] = 0.f;
        __syncthreads();
        for (int row = 0; row < 72; row++) {        // value of A3 row `row` at pixel tid
            const float v = Gt[(tid >> 3) * (G_PG / 4) + (row >> 2) * 32 + (tid & 7) * 4 + (row & 3)];
            if (variant == 2) A3[(tid >> 2) * 512 + (row >> 3) * 32 + (row & 7) * 4 + (tid & 3)] = v;          // K-major, LBO = 2048, SBO = 128
            else A3[(tid >> 3) * 1024 + (row >> 2) * 32 + (tid & 7) * 4 + (row & 3)] = v;                       // dense MN-major, LBO = 4096, SBO = 128
        }
    }
    // F tile: thread t -> cand = t % 16, quad = t / 16
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
    // B3 tile: thread = pixel writes w / q of the 16 candidates (hi, lo): column n at (n/8)*128 + (n%8)*16 bytes
    {
        float* base = B3 + (tid >> 2) * (B3_LBO / 4) + (tid & 3);
        for (int j = 0; j < NB; j++) {
            const float w = WQ[tid * 2 * NB + j], q = WQ[tid * 2 * NB + NB + j];
            const float wh = hi_of(w), qh = hi_of(q);
            base[((j >> 3) + 0) * 32 + (j & 7) * 4] = wh;
            base[((j >> 3) + 2) * 32 + (j & 7) * 4] = w - wh;
            base[((j >> 3) + 4) * 32 + (j & 7) * 4] = qh;
            base[((j >> 3) + 6) * 32 + (j & 7) * 4] = q - qh;
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1));
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
    const uint32_t tS = tmem, tD = tmem + 64;

    constexpr uint32_t ID1 = idesc_tf32(128, NB, 0, 0);          // A K-major, B K-major
    const uint32_t ID3 = (variant >= 6) ? idesc_tf32(128, 80, 0, 1) : idesc_tf32(variant == 4 ? 64 : 128, 4 * NB, variant == 2 ? 0 : 1, 0);
    bool ok = true;
    long long c1 = 0, c3 = 0;
    for (int r = 0; r < reps && ok; r++) {
        long long t0 = clock64();
        if (tid == 0) {
            const uint32_t g0 = smem_u32(Gt), fh = smem_u32(Fh), fl = smem_u32(Fl);
            int first = 1;
            for (int term = 0; term < 3; term++) {                 // lo*hi, hi*lo, hi*hi
                const uint32_t a0 = g0 + (term == 0 ? 1024u : 0u);
                const uint32_t b0 = (term == 1) ? fl : fh;
                for (int ks = 0; ks < 4; ks++) {
                    mma_tf32(tS, make_desc(a0 + ks * 256, 128, G_PG), make_desc(b0 + ks * 2 * F_LBO, F_LBO, 128), ID1, first ? 0u : 1u);
                    first = 0;
                }
            }
            commit(&bar[0]);
        }
        ok = mbar_wait(&bar[0], r & 1, 20000000);
        long long t1 = clock64();
        if (tid == 0) {
            const uint32_t g0 = smem_u32(Gt), b0 = smem_u32(B3);
            const uint32_t a3 = smem_u32(A3);
            for (int ks = 0; ks < NPX / 8; ks++) {
                uint64_t da;
                if (variant >= 6) {
                    const uint64_t db = (variant == 6) ? make_desc(g0 + ks * G_PG, G_PG, 128) : make_desc(a3 + ks * 4096, 4096, 128);
                    mma_tf32(tS + 32, make_desc(b0 + ks * 2 * B3_LBO, B3_LBO, 128), db, ID3, ks > 0 ? 1u : 0u);
                    continue;
                }
                if (variant == 2) da = make_desc(a3 + ks * 2 * 2048, 2048, 128);
                else if (variant == 3) da = make_desc(g0 + ks * G_PG, 128, G_PG);
                else if (variant == 5) da = make_desc(a3 + ks * 4096, 4096, 128);
                else da = make_desc(g0 + ks * G_PG, G_PG, 128);
                mma_tf32(tD, da, make_desc(b0 + ks * 2 * B3_LBO, B3_LBO, 128), ID3, ks > 0 ? 1u : 0u);
            }
            commit(&bar[1]);
        }
        ok = ok && mbar_wait(&bar[1], r & 1, 20000000);
        long long t2 = clock64();
        c1 += t1 - t0; c3 += t2 - t1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) status[0] = 1; }
    else {
        uint32_t v[16];
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tS + lane_base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < 16; n++) S_out[tid * 16 + n] = __uint_as_float(v[n]);
        const int ncol16 = (variant >= 6) ? 5 : 4;
        const uint32_t tRead = (variant >= 6) ? tS + 32 : tD;
        const int ostride = (variant >= 6) ? 80 : 64;
        for (int c = 0; c < ncol16; c++) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tRead + lane_base + 16 * c));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int n = 0; n < 16; n++) D_out[tid * ostride + 16 * c + n] = __uint_as_float(v[n]);
        }
        if (tid == 0) { status[0] = 0; cyc[0] = c1; cyc[1] = c3; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main()
{
    srand(7);
    float *G = new float[NPX * NCH], *X = new float[NPX * 8], *F = new float[NB * NCH], *WQ = new float[NPX * 2 * NB];
    for (int i = 0; i < NPX * NCH; i++) G[i] = (rand() / (float)RAND_MAX - 0.5f) * 1e-6f;
    for (int p = 0; p < NPX; p++) {
        const float x = (float)(p & 15) - 7.5f, y = (float)(p >> 4) - 3.5f;
        const float b[8] = {1.f, x, y, x * x, x * y, y * y, 0.f, 0.f};
        for (int m = 0; m < 8; m++) X[p * 8 + m] = b[m];
    }
    for (int i = 0; i < NB * NCH; i++) F[i] = (rand() / (float)RAND_MAX - 0.5f);
    for (int i = 0; i < NPX * 2 * NB; i++) WQ[i] = (rand() / (float)RAND_MAX - 0.3f);
    float *dG, *dX, *dF, *dWQ, *dS, *dD; int* dSt; long long* dC;
    cudaMalloc(&dG, NPX * NCH * 4); cudaMalloc(&dX, NPX * 8 * 4); cudaMalloc(&dF, NB * NCH * 4); cudaMalloc(&dWQ, NPX * 2 * NB * 4);
    cudaMalloc(&dS, NPX * 16 * 4); cudaMalloc(&dD, NPX * 80 * 4); cudaMalloc(&dSt, 4); cudaMalloc(&dC, 16);
    cudaMemcpy(dG, G, NPX * NCH * 4, cudaMemcpyHostToDevice); cudaMemcpy(dX, X, NPX * 8 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dF, F, NB * NCH * 4, cudaMemcpyHostToDevice); cudaMemcpy(dWQ, WQ, NPX * 2 * NB * 4, cudaMemcpyHostToDevice);
    const int smem = G_BYTES + B3_BYTES + 2 * F_BYTES + 64 + 65536;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int variant = 0; variant < 8; variant++)
    for (int reps : {1, 500}) {
        int st = -1; cudaMemcpy(dSt, &st, 4, cudaMemcpyHostToDevice);
        cudaMemset(dS, 0, NPX * 16 * 4); cudaMemset(dD, 0, NPX * 80 * 4);
        probe<<<1, 128, smem>>>(dG, dX, dF, dWQ, dS, dD, dSt, dC, reps, variant);
        cudaError_t e = cudaDeviceSynchronize();
        float* S = new float[NPX * 16]; float* D = new float[NPX * 80]; long long cyc[2] = {0, 0};
        cudaMemcpy(&st, dSt, 4, cudaMemcpyDeviceToHost); cudaMemcpy(S, dS, NPX * 16 * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(D, dD, NPX * 80 * 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost);
        double e1 = 0, m1 = 0, e3 = 0, m3 = 0, e3x = 0, m3x = 0;
        for (int p = 0; p < NPX; p++) for (int j = 0; j < NB; j++) {
            double s = 0; for (int c = 0; c < NCH; c++) s += (double)G[p * NCH + c] * F[j * NCH + c];
            e1 = fmax(e1, fabs(S[p * 16 + j] - s)); m1 = fmax(m1, fabs(s));
        }
        // rows 0..31 (+32..63): sum over columns (W_hi + W_lo) of both row groups = sum_p g[p][ch] w[p][j]; rows 64..69: columns Q
        if (variant >= 6) {
            // D'[row n][col m]: n = j (w_hi), 16 + j (w_lo), 32 + j (q_hi), 48 + j (q_lo); m = ch (hi), 32 + ch (lo), 64.. basis
            for (int ch = 0; ch < NCH; ch++) for (int j = 0; j < NB; j++) {
                double s = 0; for (int p = 0; p < NPX; p++) s += (double)G[p * NCH + ch] * WQ[p * 2 * NB + j];
                const double got = (double)D[j * 80 + ch] + D[j * 80 + 32 + ch] + D[(16 + j) * 80 + ch] + D[(16 + j) * 80 + 32 + ch];
                e3 = fmax(e3, fabs(got - s)); m3 = fmax(m3, fabs(s));
            }
            for (int m = 0; m < 6; m++) for (int j = 0; j < NB; j++) {
                double s = 0; for (int p = 0; p < NPX; p++) s += (double)X[p * 8 + m] * WQ[p * 2 * NB + NB + j];
                const double got = (double)D[(32 + j) * 80 + 64 + m] + D[(48 + j) * 80 + 64 + m];
                e3x = fmax(e3x, fabs(got - s) / fmax(1.0, fabs(s))); m3x = fmax(m3x, fabs(s));
            }
        } else {
        for (int ch = 0; ch < NCH; ch++) for (int j = 0; j < NB; j++) {
            double s = 0; for (int p = 0; p < NPX; p++) s += (double)G[p * NCH + ch] * WQ[p * 2 * NB + j];
            const double got = (double)D[ch * 64 + j] + D[ch * 64 + 16 + j] + D[(32 + ch) * 64 + j] + D[(32 + ch) * 64 + 16 + j];
            e3 = fmax(e3, fabs(got - s)); m3 = fmax(m3, fabs(s));
        }
        for (int m = 0; m < 6; m++) for (int j = 0; j < NB; j++) {
            double s = 0; for (int p = 0; p < NPX; p++) s += (double)X[p * 8 + m] * WQ[p * 2 * NB + NB + j];
            const double got = (double)D[(64 + m) * 64 + 32 + j] + D[(64 + m) * 64 + 48 + j];
            e3x = fmax(e3x, fabs(got - s) / fmax(1.0, fabs(s))); m3x = fmax(m3x, fabs(s));
        }
        }
        const bool good = e == cudaSuccess && st == 0 && e1 < 2e-5 * m1 && e3 < 2e-5 * m3 && e3x < 2e-5;
        printf("variant=%d reps=%d cuda=%s status=%d | GEMM1 rel err %.2e | GEMM3 colour rel err %.2e, moments rel err %.2e | cycles: GEMM1 %.0f GEMM3 %.0f | %s\n",
               variant, reps, cudaGetErrorString(e), st, e1 / m1, e3 / m3, e3x, (double)cyc[0] / reps, (double)cyc[1] / reps, good ? "BWD LAYOUTS OK" : "BWD LAYOUTS WRONG");
        if (!good) {
            printf(" S[0][0..3] %g %g %g %g\n", S[0], S[1], S[2], S[3]);
            printf(" D[0][0..3] %g %g %g %g | D[64][32..35] %g %g %g %g\n", D[0], D[1], D[2], D[3], D[64 * 64 + 32], D[64 * 64 + 33], D[64 * 64 + 34], D[64 * 64 + 35]);
            double want0 = 0; for (int p2 = 0; p2 < NPX; p2++) want0 += (double)G[p2 * NCH + 0] * WQ[p2 * 2 * NB + 0];
            printf(" expected colour[0][0] %g (sum of the four D entries %g); nonzero D entries:", want0, (double)D[0] + D[16] + D[32 * 64] + D[32 * 64 + 16]);
            int nz = 0; for (int i = 0; i < NPX * 64; i++) if (D[i] != 0.f) nz++;
            printf(" %d of %d\n", nz, NPX * 64);
        }
    }
    return 0;
}
