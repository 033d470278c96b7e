// render_forward_tc_kernels.cuh -- the device code of render_forward_tc.cu (see there); free of host-side runtime calls so that the
// CPU suite can run it under tests/cuda_emu/ (tcgen05 / TMEM restated in tests/cuda_emu/tc_emu.h).
#pragma once
#include "common.cuh"
#include "cp_async.cuh"
#include "tc.cuh"
#include "candidate.cuh"

// the operand tiles of tcgen05.mma want generous alignment: this kernel asks for 1024 bytes (the CPU shim aligns its buffer likewise)
#ifndef SAGARS_DYNAMIC_SMEM_1024
#if defined(SAGARS_CUDA_EMU)
#define SAGARS_DYNAMIC_SMEM_1024(name) SAGARS_DYNAMIC_SMEM(name)
#else
#define SAGARS_DYNAMIC_SMEM_1024(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
#endif

namespace sagars {


constexpr int TCF_BATCH = 64;   // instances staged per cp.async stage
constexpr int TCF_SUB = 16;     // k-slots per MMA issue (2 k-steps of 8)
constexpr int TCF_N = 32;       // channels

struct FwdTcSmem {
    float A[2][2][TCF_SUB * 128];       // [group][hi, lo]  K-major: (k/4)*512 + (r/8)*32 + (r%8)*4 + k%4      (floats)
    float B[2][2][TCF_SUB * TCF_N];     // [group][hi, lo]  K-major: (k/4)*128 + (n/8)*32 + (n%8)*4 + k%4
    float4 geo[2][TCF_BATCH][2];        // x, y, cx, cy | cz, opacity, accept_threshold, -
    float4 feat[2][TCF_BATCH][TCF_N / 4];
    uint32_t ids[2][TCF_BATCH];
    uint32_t cmask[2][2][4][2];         // [batch parity][group][warp of the group][low, high 32 splats]: candidate masks
    uint8_t glist[8][TCF_BATCH + 4];    // per warp: its copy of the group's candidate list (+ padding of the last chunk)
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
    SAGARS_DYNAMIC_SMEM_1024(smem_raw);
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
    SAGARS_PIN_F2(pixx, pixy);

    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    const int nbatch = (total + TCF_BATCH - 1) / TCF_BATCH;

    // one-time setup: mbarriers, TMEM (64 columns = two 128x32 fp32 accumulators), zeroed feature staging
    {
        float* f = reinterpret_cast<float*>(&sm.feat[0][0][0]);
        for (int c = tid; c < 2 * TCF_BATCH * TCF_N; c += TILE_PIX) f[c] = 0.f;
        float* bt = &sm.B[0][0][0];   // slots of a partially filled last chunk keep whatever is here: must be finite
        for (int c = tid; c < 2 * 2 * TCF_SUB * TCF_N; c += TILE_PIX) bt[c] = 0.f;
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
    uint32_t n_issued = 0;               // MMA issues committed by this group so far (group-uniform)
    int nslot = 0;                       // k-slots of the current operand tiles already written (group-uniform, multiple of 4)

    constexpr uint32_t A_SBO = 128, A_LBO = 128 * 16;     // bytes
    constexpr uint32_t B_SBO = 128, B_LBO = 128 * (TCF_N / 8);
    constexpr uint32_t IDESC = tc::idesc_tf32(128, TCF_N, 0, 0);

    // the warp's 8x4 pixel block (pixel centres), for the block-level candidate test
    const float bx0 = (float)(blockIdx.x * TILE_X + (warp & 1) * 8), bx1 = bx0 + 7.f;
    const float by0 = (float)(blockIdx.y * TILE_Y + (warp >> 1) * 4), by1 = by0 + 3.f;

    float* const Ah = &sm.A[grp][0][(gt >> 3) * 32 + (gt & 7) * 4];
    float* const Al = &sm.A[grp][1][(gt >> 3) * 32 + (gt & 7) * 4];
    // F^T tiles: lane handles slot (lane & 3) of a chunk and channel 8 * (warp of the group) + lane / 4 -> 32 consecutive floats per warp
    const int b_slot = lane & 3, b_ch = 8 * (warp & 3) + (lane >> 2);
    float* const Bh = &sm.B[grp][0][(b_ch >> 3) * 32 + (b_ch & 7) * 4 + b_slot];
    float* const Bl = &sm.B[grp][1][(b_ch >> 3) * 32 + (b_ch & 7) * 4 + b_slot];

    // 16 slots are complete (or padded): hand the operand tiles to the tensor core
    auto issue_mma = [&]() {
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
        nslot = 0;
    };

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
        if (__syncthreads_and(done)) break;

        if (b + 1 < nbatch) {
            tcf_issue_batch(sm, stage ^ 1, (b + 1) & 1, min(TCF_BATCH, total - (b + 1) * TCF_BATCH), geo, features);
            cp_async_commit();
        }
        uint32_t next_id = 0;
        const int rem2 = total - (b + 2) * TCF_BATCH;
        const bool have_next_id = (b + 2 < nbatch) && tid < min(TCF_BATCH, rem2);
        if (have_next_id) next_id = point_list[range.x + (b + 2) * TCF_BATCH + tid];

        // ---- candidates: lane = splat against this warp's pixel block; the group ORs its four masks ----
        // Every warp then writes its own copy of the group's candidate list (in list order): entry = batch-local splat
        // index | 0x80 when the splat is also a candidate of THIS warp's block.  Four entries = one 32-bit word.
        int ng;
        {
            uint32_t own_lo = 0u, own_hi = 0u, grp_lo = 0u, grp_hi = 0u;
            if (!__all_sync(0xffffffffu, done)) {
                own_lo = __ballot_sync(0xffffffffu, !block_rejects(sm.geo[stage][lane][0], sm.geo[stage][lane][1], bx0, bx1, by0, by1));
                own_hi = __ballot_sync(0xffffffffu, !block_rejects(sm.geo[stage][32 + lane][0], sm.geo[stage][32 + lane][1], bx0, bx1, by0, by1));
            }
            if (lane == 0) { sm.cmask[stage][grp][warp & 3][0] = own_lo; sm.cmask[stage][grp][warp & 3][1] = own_hi; }
            tc::bar_sync_128(1 + grp);
#pragma unroll
            for (int w4 = 0; w4 < 4; w4++) { grp_lo |= sm.cmask[stage][grp][w4][0]; grp_hi |= sm.cmask[stage][grp][w4][1]; }
            const uint32_t lt = (1u << lane) - 1u;
            const int n_lo = __popc(grp_lo);
            ng = n_lo + __popc(grp_hi);
            uint8_t* gl = &sm.glist[warp][0];
            if ((grp_lo >> lane) & 1u) gl[__popc(grp_lo & lt)] = (uint8_t)(lane | (((own_lo >> lane) & 1u) << 7));
            if ((grp_hi >> lane) & 1u) gl[n_lo + __popc(grp_hi & lt)] = (uint8_t)((32 + lane) | (((own_hi >> lane) & 1u) << 7));
            if (lane < 3) gl[ng + lane] = 0;   // padding of the last chunk: staged row 0, not a candidate -> weight 0
            __syncwarp();
        }

        // ---- the group's candidates, four k-slots at a time, in list order ----
        const float4* gp = &sm.geo[stage][0][0];
        const float* fp = reinterpret_cast<const float*>(&sm.feat[stage][0][0]) + b_ch;
#pragma unroll 1
        for (int k0 = 0; k0 < ng; k0 += 4) {
            const uint32_t ent = *reinterpret_cast<const uint32_t*>(&sm.glist[warp][k0]);   // 4 entries
            float wq[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                wq[i] = 0.f;
                const uint32_t e = (ent >> (8 * i)) & 0xffu;
                if (e & 0x80u) {                                                       // warp-uniform: candidate of this warp's block
                    const int jj = (int)(e & 0x3fu);
                    const float4 g0 = gp[2 * jj];
                    const float4 g1 = gp[2 * jj + 1];
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    const float pw = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    const bool cd = !done && !(pw > 0.0f) && (pw >= g1.z);
                    if (cd) {
                        const float alpha = fminf(0.99f, g1.y * expf(pw));
                        if (!(alpha < 1.0f / 255.0f)) {
                            const float test_T = T * (1 - alpha);
                            if (test_T < 0.0001f) {
                                done = true;
                            } else {
                                wq[i] = alpha * T;
                                T = test_T;
                                last_contributor = (uint32_t)(b * TCF_BATCH + jj + 1);
                            }
                        }
                    }
                }
            }
            // operand tiles: wait until the previous issue's MMAs have consumed them before the first rewrite
            if (nslot == 0 && n_issued > 0) tc::mbar_wait(&sm.mbar[grp], (n_issued - 1) & 1);
            {
                float4 h, l;
                h.x = tc::tf32_hi(wq[0]); l.x = wq[0] - h.x;
                h.y = tc::tf32_hi(wq[1]); l.y = wq[1] - h.y;
                h.z = tc::tf32_hi(wq[2]); l.z = wq[2] - h.z;
                h.w = tc::tf32_hi(wq[3]); l.w = wq[3] - h.w;
                *reinterpret_cast<float4*>(Ah + (nslot >> 2) * 512) = h;
                *reinterpret_cast<float4*>(Al + (nslot >> 2) * 512) = l;
                // this lane transposes channel b_ch of the splat in slot (lane & 3); padded slots read a finite staged row
                const int jsel = (int)((ent >> (8 * b_slot)) & 0x3fu);
                const float f = fp[jsel * TCF_N];
                const float fh = tc::tf32_hi(f);
                Bh[(nslot >> 2) * 128] = fh;
                Bl[(nslot >> 2) * 128] = f - fh;
            }
            nslot += 4;
            if (nslot == TCF_SUB) issue_mma();
        }

        if (have_next_id) sm.ids[b & 1][tid] = next_id;
        cp_async_wait_all();
        if (b + 1 < nbatch) tcf_pad_batch(sm, stage ^ 1, min(TCF_BATCH, total - (b + 1) * TCF_BATCH));
        __syncthreads();
    }
    if (nslot > 0) {   // last, partially filled operand tiles: zero weights in the remaining slots
        for (int c = nslot >> 2; c < TCF_SUB / 4; c++) {
            *reinterpret_cast<float4*>(Ah + c * 512) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(Al + c * 512) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        issue_mma();
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

}  // namespace sagars
