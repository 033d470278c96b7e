// render_forward_tc.cu -- forward blend with the C-channel contraction on the 5th-generation tensor cores.
//
// Same semantics as render_forward.cu (CF cuda_rasterizer/forward.cu:264-385), same per-pixel scalar arithmetic
// for power / alpha / T / the 1/255 and 1e-4 tests (so final_T and n_contrib stay bit-identical), but the colour
// accumulation  C[p][ch] = sum_j w[p][j] * f[j][ch],  w = alpha * T,  is expressed as the dense per-tile GEMM of
// SURVEY.md Appendix D and runs on tcgen05:
//
//   * a CTA is one 16x16 tile = two independent 128-pixel groups (4 warps each).  Thread t of a group owns pixel row t
//     of that group's A operand and TMEM lane t of its accumulator D[128 x 32] (fp32, 32 TMEM columns per group);
//   * per sub-batch of 16 splats every thread runs the scalar loop and keeps w[16] in registers (0 where the pair
//     is rejected / the pixel is saturated), then writes its row of the K-major tf32 operand tiles W_hi / W_lo
//     (SWIZZLE_NONE canonical layout) while 128 threads transpose-split the 16 feature rows into F_hi / F_lo;
//   * one elected thread per group issues  D += W_lo F_hi + W_hi F_lo + W_hi F_hi  (3xTF32: 6 tcgen05.mma.kind::tf32
//     of 128x32x8, ~2^-21 relative error) and commits to the group's mbarrier; the next sub-batch's scalar loop
//     overlaps the MMAs and only waits for them right before it overwrites the operand tiles;
//   * at the end each thread reads its 32 channel sums with one tcgen05.ld and adds T * bg.
//
// The colour image differs from the SIMT kernel / the reference by fp32-level rounding only (3xTF32 and a different
// summation order), far inside the 1e-4 parity tolerance; integer state is unaffected.
#include "common.cuh"
#include "cp_async.cuh"
#include "tc.cuh"

namespace sagars {

constexpr int TCF_BATCH = 64;   // instances staged per cp.async stage
constexpr int TCF_SUB = 16;     // instances per MMA sub-batch (2 k-steps of 8)
constexpr int TCF_N = 32;       // channels

struct FwdTcSmem {
    float A[2][2][TCF_SUB * 128];       // [group][hi, lo]  K-major: (k/4)*512 + (r/8)*32 + (r%8)*4 + k%4      (floats)
    float B[2][2][TCF_SUB * TCF_N];     // [group][hi, lo]  K-major: (k/4)*128 + (n/8)*32 + (n%8)*4 + k%4
    float4 geo[2][TCF_BATCH][2];        // x, y, cx, cy | cz, opacity, accept_threshold, -
    float4 feat[2][TCF_BATCH][TCF_N / 4];
    uint32_t ids[2][TCF_BATCH];
    uint64_t mbar[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tcf_issue_batch(FwdTcSmem& sm, int stage, int idbuf, int cnt, const float* __restrict__ geo,
                                                const float* __restrict__ features)
{
    const int tid = threadIdx.x;
    for (int c = tid; c < cnt * 2; c += TILE_PIX) {
        const int j = c >> 1, h = c & 1;
        cp_async16(&sm.geo[stage][j][h], geo + 8 * (size_t)sm.ids[idbuf][j] + 4 * h);
    }
    for (int c = tid; c < cnt * (TCF_N / 4); c += TILE_PIX) {
        const int j = c >> 3, q = c & 7;
        cp_async16(&sm.feat[stage][j][q], features + (size_t)sm.ids[idbuf][j] * TCF_N + 4 * q);
    }
}

// records past the end of the tile's list: can never be accepted (threshold = +inf), contribute w = 0
__device__ __forceinline__ void tcf_pad_batch(FwdTcSmem& sm, int stage, int cnt)
{
    const int tid = threadIdx.x;
    if (tid >= cnt && tid < TCF_BATCH) {
        sm.geo[stage][tid][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.geo[stage][tid][1] = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
    }
}

__global__ void __launch_bounds__(TILE_PIX, 3)
render_forward_tc_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                         const float* __restrict__ geo, const float* __restrict__ features, const float* __restrict__ bg,
                         float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, float* __restrict__ out_color)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FwdTcSmem& sm = *reinterpret_cast<FwdTcSmem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 2;          // 128-pixel group
    const int gt = tid & 127;           // row of this thread in its group's operand / accumulator
    const int tiles_x = gridDim.x;
    const uint32_t px = blockIdx.x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = blockIdx.y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    asm volatile("" : "+f"(pixx), "+f"(pixy));

    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    const int nbatch = (total + TCF_BATCH - 1) / TCF_BATCH;

    // one-time setup: mbarriers, TMEM (64 columns = two 128x32 fp32 accumulators), zeroed feature staging
    {
        float* f = reinterpret_cast<float*>(&sm.feat[0][0][0]);
        for (int c = tid; c < 2 * TCF_BATCH * TCF_N; c += TILE_PIX) f[c] = 0.f;
    }
    if (tid == 0) {
        tc::mbar_init(&sm.mbar[0], 1);
        tc::mbar_init(&sm.mbar[1], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc<64>(&sm.tmem_base);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = sm.tmem_base + (uint32_t)(grp * TCF_N);          // this group's accumulator columns

    float T = 1.0f;
    uint32_t last_contributor = 0;
    bool done = !inside;
    uint32_t n_issued = 0;               // MMA sub-batches committed by this group so far (group-uniform)

    constexpr uint32_t A_SBO = 128, A_LBO = 128 * 16;     // bytes
    constexpr uint32_t B_SBO = 128, B_LBO = 128 * (TCF_N / 8);
    constexpr uint32_t IDESC = tc::idesc_tf32(128, TCF_N, 0, 0);

    if (nbatch > 0) {
        if (tid < min(TCF_BATCH, total)) sm.ids[0][tid] = point_list[range.x + tid];
        __syncthreads();
        tcf_issue_batch(sm, 0, 0, min(TCF_BATCH, total), geo, features);
        cp_async_commit();
        if (nbatch > 1 && tid < min(TCF_BATCH, total - TCF_BATCH)) sm.ids[1][tid] = point_list[range.x + TCF_BATCH + tid];
        cp_async_wait_all();
        tcf_pad_batch(sm, 0, min(TCF_BATCH, total));
        __syncthreads();
    }

    for (int b = 0; b < nbatch; b++) {
        const int stage = b & 1;
        const int cnt = min(TCF_BATCH, total - b * TCF_BATCH);
        if (__syncthreads_and(done)) break;

        if (b + 1 < nbatch) {
            tcf_issue_batch(sm, stage ^ 1, (b + 1) & 1, min(TCF_BATCH, total - (b + 1) * TCF_BATCH), geo, features);
            cp_async_commit();
        }
        uint32_t next_id = 0;
        const int rem2 = total - (b + 2) * TCF_BATCH;
        const bool have_next_id = (b + 2 < nbatch) && tid < min(TCF_BATCH, rem2);
        if (have_next_id) next_id = point_list[range.x + (b + 2) * TCF_BATCH + tid];

        for (int sb = 0; sb < cnt; sb += TCF_SUB) {
            // ---- scalar part: w[j] = alpha * T of this pixel for the 16 splats of the sub-batch (0 if rejected) ----
            float w[TCF_SUB];
            const float4* gp = &sm.geo[stage][sb][0];
            // four splats at a time: their `power` tests are independent (ILP, one vote per four); only the accepted
            // ones are then taken in order, because T and `done` chain through them.  Records beyond the tile's list
            // are sentinels (accept_threshold = +inf), so no bounds checks are needed here.
#pragma unroll
            for (int j0 = 0; j0 < TCF_SUB; j0 += 4) {
                float pw[4], op[4];
                bool cd[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 g0 = gp[2 * (j0 + i)];
                    const float4 g1 = gp[2 * (j0 + i) + 1];
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    pw[i] = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    cd[i] = !(pw[i] > 0.0f) && (pw[i] >= g1.z);
                    op[i] = g1.y;
                    w[j0 + i] = 0.f;
                }
                const bool anyc = (cd[0] || cd[1] || cd[2] || cd[3]) && !done;
                if (__any_sync(0xffffffffu, anyc)) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (cd[i] && !done) {
                            const float alpha = fminf(0.99f, op[i] * expf(pw[i]));
                            if (!(alpha < 1.0f / 255.0f)) {
                                const float test_T = T * (1 - alpha);
                                if (test_T < 0.0001f) {
                                    done = true;
                                } else {
                                    w[j0 + i] = alpha * T;
                                    T = test_T;
                                    last_contributor = (uint32_t)(b * TCF_BATCH + sb + j0 + i + 1);
                                }
                            }
                        }
                    }
                }
            }

            // ---- operand tiles: wait until the previous sub-batch's MMAs have consumed them, then rewrite ----
            if (n_issued > 0) tc::mbar_wait(&sm.mbar[grp], (n_issued - 1) & 1);
            {
                float* Ah = &sm.A[grp][0][(gt >> 3) * 32 + (gt & 7) * 4];
                float* Al = &sm.A[grp][1][(gt >> 3) * 32 + (gt & 7) * 4];
#pragma unroll
                for (int c = 0; c < TCF_SUB / 4; c++) {
                    float4 h, l;
                    h.x = tc::tf32_hi(w[4 * c + 0]); l.x = w[4 * c + 0] - h.x;
                    h.y = tc::tf32_hi(w[4 * c + 1]); l.y = w[4 * c + 1] - h.y;
                    h.z = tc::tf32_hi(w[4 * c + 2]); l.z = w[4 * c + 2] - h.z;
                    h.w = tc::tf32_hi(w[4 * c + 3]); l.w = w[4 * c + 3] - h.w;
                    *reinterpret_cast<float4*>(Ah + c * 512) = h;
                    *reinterpret_cast<float4*>(Al + c * 512) = l;
                }
                // feature rows -> F^T tiles (rows = channels, k = splat): thread (j, chunk) handles 4 channels of splat j
                const int j = gt >> 3, c4 = gt & 7;
                const float4 raw = sm.feat[stage][sb + j][c4];   // rows beyond cnt hold stale finite data; their w is 0
                const float r[4] = {raw.x, raw.y, raw.z, raw.w};
                float* Bh = &sm.B[grp][0][(j >> 2) * 128 + (j & 3)];
                float* Bl = &sm.B[grp][1][(j >> 2) * 128 + (j & 3)];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int n = 4 * c4 + i;
                    const float h = tc::tf32_hi(r[i]);
                    Bh[(n >> 3) * 32 + (n & 7) * 4] = h;
                    Bl[(n >> 3) * 32 + (n & 7) * 4] = r[i] - h;
                }
            }
            tc::fence_smem_to_async_proxy();
            tc::fence_before_sync();
            tc::bar_sync_128(1 + grp);
            if (gt == 0) {
                tc::fence_after_sync();
                const uint32_t a_hi = smem_u32(&sm.A[grp][0][0]), a_lo = smem_u32(&sm.A[grp][1][0]);
                const uint32_t b_hi = smem_u32(&sm.B[grp][0][0]), b_lo = smem_u32(&sm.B[grp][1][0]);
#pragma unroll
                for (int term = 0; term < 3; term++) {            // lo*hi, hi*lo, hi*hi (small terms first)
                    const uint32_t a0 = (term == 0) ? a_lo : a_hi;
                    const uint32_t b0 = (term == 1) ? b_lo : b_hi;
#pragma unroll
                    for (int ks = 0; ks < TCF_SUB / 8; ks++) {
                        const uint64_t da = tc::smem_desc(a0 + ks * 2 * A_LBO, A_LBO, A_SBO);
                        const uint64_t db = tc::smem_desc(b0 + ks * 2 * B_LBO, B_LBO, B_SBO);
                        tc::mma_tf32(tmem_d, da, db, IDESC, (n_issued > 0 || term > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                tc::commit(&sm.mbar[grp]);
            }
            n_issued++;
        }

        if (have_next_id) sm.ids[b & 1][tid] = next_id;
        cp_async_wait_all();
        if (b + 1 < nbatch) tcf_pad_batch(sm, stage ^ 1, min(TCF_BATCH, total - (b + 1) * TCF_BATCH));
        __syncthreads();
    }

    // ---- epilogue: accumulator row -> registers -> planar image ----
    float Cacc[TCF_N];
    if (n_issued > 0) {
        tc::mbar_wait(&sm.mbar[grp], (n_issued - 1) & 1);
        tc::fence_after_sync();
        tc::tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16), Cacc);
    } else {
#pragma unroll
        for (int k = 0; k < TCF_N; k++) Cacc[k] = 0.f;
    }
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int k = 0; k < TCF_N; k++) out_color[(size_t)k * plane + pix_id] = Cacc[k] + T * bg[k];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(sm.tmem_base);
}

int launch_render_forward_tc(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                             const uint32_t* point_list, cudaStream_t s, bool debug)
{
    auto kern = render_forward_tc_kernel;
    const size_t smem = sizeof(FwdTcSmem) + 1024;
    SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(d.tiles_x, d.tiles_y);
    kern<<<grid, TILE_PIX, smem, s>>>(im.ranges, point_list, d.W, d.H, g.geo, a.colors_precomp, a.background,
                                      im.final_T, im.n_contrib, a.out_color);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

}  // namespace sagars
