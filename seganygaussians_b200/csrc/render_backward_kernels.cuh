// render_backward_kernels.cuh -- the device code of render_backward.cu (see there); free of host-side runtime calls so that the CPU
// suite can run it under tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include "cp_async.cuh"

#ifndef SAGARS_DYNAMIC_SMEM
#define SAGARS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sagars {


constexpr int BWD_NB = 16;   // instances per batch

template <int NQ>
struct BwdCfg {
    static constexpr int TPI = NQ > 8 ? NQ : 8;   // phase-B lanes per list slot (one per channel quad / moment)
    static constexpr int SLOTS = 32 / TPI;        // list entries a warp consumes per iteration
};

template <int NQ>
struct BwdSmem {
    float4 Gs[TILE_PIX][NQ];            // upstream gradient rows, indexed by raster-local pixel (y*16 + x); padded channels = 0
    float2 ent[BWD_NB][TILE_PIX];       // compact list of blended pairs per instance: (w, q)
    uint8_t entp[BWD_NB][TILE_PIX];     //   ... and their raster-local pixel index
    uint32_t cnt[2][BWD_NB];            // list lengths (double buffered across batches)
    uint32_t next_inst[2];              // phase-B work counter: warps take instances first come, first served
    float4 geo[2][BWD_NB][2];           // x, y, cx, cy | cz, opacity, accept_threshold, -
    float4 feat[BWD_NB][NQ];            // feature rows, zero padded
    float tabx[16][8];                  // moment basis factors in x: 1, x, 1, x^2, x, 1, 0, 0   (x = xl - 7.5)
    float taby[16][8];                  //                     in y: 1, 1, y, 1, y, y^2, 0, 0
    uint32_t ids[3][BWD_NB];
    uint32_t max_contrib;
};

template <int NQ>
__device__ __forceinline__ void bwd_issue_geo(BwdSmem<NQ>& sm, int stage, int idbuf, int cnt, const float* __restrict__ geo)
{
    const int tid = threadIdx.x;
    if (tid < cnt * 2) {
        const int j = tid >> 1, h = tid & 1;
        cp_async16(&sm.geo[stage][j][h], geo + 8 * (size_t)sm.ids[idbuf][j] + 4 * h);
    }
}

// records past the end of the batch: never accepted (threshold = +inf)
template <int NQ>
__device__ __forceinline__ void bwd_pad_geo(BwdSmem<NQ>& sm, int stage, int cnt)
{
    const int tid = threadIdx.x;
    if (tid >= cnt && tid < BWD_NB) {
        sm.geo[stage][tid][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.geo[stage][tid][1] = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
    }
}

template <int NQ, bool VEC>
__device__ __forceinline__ void bwd_issue_feat(BwdSmem<NQ>& sm, int idbuf, int cnt, int K, const float* __restrict__ features)
{
    const int tid = threadIdx.x;
    if (VEC) {
        const int nq = K >> 2;
        for (int c = tid; c < cnt * nq; c += TILE_PIX) {
            const int j = c / nq, q = c - j * nq;
            cp_async16(&sm.feat[j][q], features + (size_t)sm.ids[idbuf][j] * K + 4 * q);
        }
    } else {
        float* f = reinterpret_cast<float*>(&sm.feat[0][0]);
        for (int c = tid; c < cnt * K; c += TILE_PIX) {
            const int j = c / K, k = c - j * K;
            f[j * (4 * NQ) + k] = features[(size_t)sm.ids[idbuf][j] * K + k];
        }
    }
}

// NQ : float4 groups covering the gradient channels (K colour channels [+ 1 mask channel when MD])
// VEC: K % 4 == 0 and no mask channel -> dL_dcolors rows are 16-byte aligned, use red.v4
template <int NQ, bool VEC, bool MD, bool COLOR>
__global__ void __launch_bounds__(TILE_PIX, (NQ <= 8) ? 3 : 1)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       int W, int H, int K,
                       const float* __restrict__ bg, const float* __restrict__ geo,
                       const float* __restrict__ features,
                       const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dout_mask,
                       float* __restrict__ ggrad, float* __restrict__ dL_dcolors)
{
    using Cfg = BwdCfg<NQ>;
    SAGARS_DYNAMIC_SMEM(smem_raw);
    BwdSmem<NQ>& sm = *reinterpret_cast<BwdSmem<NQ>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = gridDim.x;
    const uint32_t tile_x0 = blockIdx.x * TILE_X, tile_y0 = blockIdx.y * TILE_Y;
    const int xl = (warp & 1) * 8 + (lane & 7), yl = (warp >> 1) * 4 + (lane >> 3);   // this thread's pixel in the tile
    const int rl = yl * TILE_X + xl;                                                   // raster-local index
    const uint32_t px = tile_x0 + xl, py = tile_y0 + yl;
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    // opaque to the optimiser: otherwise nvcc rematerialises both from %ctaid / %tid inside the hot loop
    SAGARS_PIN_F2(pixx, pixy);
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;

    // the tile only needs instances [0, max over its pixels of n_contrib); a warp only [0, its own max)
    if (tid == 0) sm.max_contrib = 0;
    if (tid < 2 * BWD_NB) sm.cnt[tid / BWD_NB][tid % BWD_NB] = 0;
    if (tid < 2) sm.next_inst[tid] = 0;
    if (tid < 128) {   // moment basis tables
        const int c = tid >> 3, k = tid & 7;
        const float v = (float)c - 7.5f;
        sm.tabx[c][k] = (k == 1 || k == 4) ? v : (k == 3) ? v * v : (k <= 5) ? 1.f : 0.f;
        sm.taby[c][k] = (k == 2 || k == 4) ? v : (k == 5) ? v * v : (k <= 5) ? 1.f : 0.f;
    }
    __syncthreads();
    int warp_n = my_n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_n = max(warp_n, __shfl_xor_sync(0xffffffffu, warp_n, o));
    if (lane == 0 && warp_n > 0) atomicMax(&sm.max_contrib, (uint32_t)warp_n);

    // upstream gradient of this pixel: registers for phase A, smem row for phase B
    float g[4 * NQ];
    float bgdot = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * NQ; k++) {
        float v = 0.f;
        if (COLOR && k < K) {
            v = inside ? dL_dpix[(size_t)k * plane + pix_id] : 0.f;
            bgdot += bg[k] * v;
        }
        g[k] = v;
    }
    float gmask = 0.f;
    if (MD) gmask = inside ? dL_dout_mask[pix_id] : 0.f;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        float4 v = make_float4(g[4 * q], g[4 * q + 1], g[4 * q + 2], g[4 * q + 3]);
        if (MD) {   // the mask gradient rides as channel K of the colour product
            if (K == 4 * q + 0) v.x = gmask;
            if (K == 4 * q + 1) v.y = gmask;
            if (K == 4 * q + 2) v.z = gmask;
            if (K == 4 * q + 3) v.w = gmask;
        }
        sm.Gs[rl][q] = v;
    }
    // zero the padded feature channels once
    if (!VEC || (K >> 2) < NQ) {
        float* f = reinterpret_cast<float*>(&sm.feat[0][0]);
        for (int c = tid; c < BWD_NB * 4 * NQ; c += TILE_PIX) f[c] = 0.f;
    }
    __syncthreads();

    const int maxc = min((int)sm.max_contrib, total);
    if (maxc <= 0) return;
    const int nbatch = (maxc + BWD_NB - 1) / BWD_NB;
    // batch b covers list positions pos_hi(b) - jj, jj = 0 .. cnt(b)-1, with pos_hi(b) = maxc - 1 - b*NB
    auto batch_cnt = [&](int b) { return min(BWD_NB, maxc - b * BWD_NB); };
    auto load_id = [&](int b, int jj) { return point_list[range.x + (maxc - 1 - b * BWD_NB - jj)]; };

    // prologue
    if (tid < batch_cnt(0)) sm.ids[0][tid] = load_id(0, tid);
    __syncthreads();
    bwd_issue_geo<NQ>(sm, 0, 0, batch_cnt(0), geo);
    if (COLOR) bwd_issue_feat<NQ, VEC>(sm, 0, batch_cnt(0), K, features);
    cp_async_commit();
    if (nbatch > 1 && tid < batch_cnt(1)) sm.ids[1][tid] = load_id(1, tid);
    cp_async_wait_all();
    bwd_pad_geo<NQ>(sm, 0, batch_cnt(0));
    __syncthreads();

    float T = T_final;
    float acc = 0.f, last_alpha = 0.f, last_s = 0.f;

    // phase-B role of this lane
    const int b_slot = lane / Cfg::TPI;
    const int b_k = lane % Cfg::TPI;
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;   // (0.5 * W) rounded to float, as the reference
    const float tcx = (float)tile_x0 + 7.5f, tcy = (float)tile_y0 + 7.5f;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int b = 0; b < nbatch; b++) {
        const int stage = b & 1;
        const int idb = b % 3;
        const int cnt = batch_cnt(b);
        const int pos_hi = maxc - 1 - b * BWD_NB;

        // records of batch b+1 start flying; ids of batch b+2 into a register; next batch's list lengths cleared
        if (b + 1 < nbatch) {
            bwd_issue_geo<NQ>(sm, stage ^ 1, (b + 1) % 3, batch_cnt(b + 1), geo);
            cp_async_commit();
        }
        uint32_t next_id = 0;
        const bool have_next_id = (b + 2 < nbatch) && tid < batch_cnt(b + 2);
        if (have_next_id) next_id = load_id(b + 2, tid);
        if (tid < BWD_NB) sm.cnt[stage ^ 1][tid] = 0;
        if (tid == BWD_NB) sm.next_inst[stage ^ 1] = 0;

        // ---------------- phase A: thread = pixel ----------------
        if (pos_hi - (cnt - 1) < warp_n) {   // some pixel of this warp still has contributors in this batch
            // four splats at a time: independent `power` tests (ILP, one vote per four); the accepted ones are then
            // taken in order.  Records beyond the batch are sentinels (accept_threshold = +inf).
            for (int j0 = 0; j0 < cnt; j0 += 4) {
                if (pos_hi - (j0 + 3) >= warp_n) continue;                              // warp-uniform
                float pw[4], op[4];
                bool cd[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 g0 = sm.geo[stage][j0 + i][0];
                    const float4 g1 = sm.geo[stage][j0 + i][1];
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    pw[i] = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    cd[i] = (pos_hi - (j0 + i) < my_n) && !(pw[i] > 0.0f) && (pw[i] >= g1.z);
                    op[i] = g1.y;
                }
                if (!__any_sync(0xffffffffu, cd[0] || cd[1] || cd[2] || cd[3])) continue;   // warp-uniform
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (!__any_sync(0xffffffffu, cd[i])) continue;                      // warp-uniform
                    const int jj = j0 + i;
                    bool blended = false;
                    float w_out = 0.f, q_out = 0.f;
                    if (cd[i]) {
                        const float G = expf(pw[i]);
                        const float alpha = fminf(0.99f, op[i] * G);
                        if (!(alpha < 1.0f / 255.0f)) {
                            T = T / (1.f - alpha);
                            w_out = alpha * T;
                            float s = 0.f;
                            if (COLOR) {
                                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four chains instead of one 4*NQ-long one
#pragma unroll
                                for (int q = 0; q < NQ; q++) {
                                    const float4 f = sm.feat[jj][q];
                                    s0 += f.x * g[4 * q + 0];
                                    s1 += f.y * g[4 * q + 1];
                                    s2 += f.z * g[4 * q + 2];
                                    s3 += f.w * g[4 * q + 3];
                                }
                                s = (s0 + s1) + (s2 + s3);
                            }
                            acc = last_alpha * last_s + (1.f - last_alpha) * acc;
                            last_s = s;
                            float dL_dalpha = (s - acc) * T;
                            last_alpha = alpha;
                            dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
                            q_out = G * dL_dalpha;
                            blended = true;
                        }
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, blended);
                    if (m != 0u) {
                        const int leader = __ffs(m) - 1;
                        uint32_t base = 0;
                        if (lane == leader) base = atomicAdd(&sm.cnt[stage][jj], (uint32_t)__popc(m));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        if (blended) {
                            const uint32_t e = base + __popc(m & lt_mask);
                            sm.ent[jj][e] = make_float2(w_out, q_out);
                            sm.entp[jj][e] = (uint8_t)rl;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // feature rows of batch b+1 (the single feature buffer is free now)
        if (COLOR && b + 1 < nbatch) {
            bwd_issue_feat<NQ, VEC>(sm, (b + 1) % 3, batch_cnt(b + 1), K, features);
            cp_async_commit();
        }

        // ---------------- phase B: a warp per instance, lane = (list slot, channel quad / moment) ----------------
        for (;;) {
            int jj = 0;
            if (lane == 0) jj = (int)atomicAdd(&sm.next_inst[stage], 1u);
            jj = __shfl_sync(0xffffffffu, jj, 0);
            if (jj >= cnt) break;                                                       // warp-uniform
            const int n = (int)sm.cnt[stage][jj];
            if (n == 0) continue;                                                       // warp-uniform
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            float mom = 0.f;
            for (int i = b_slot; i < n; i += Cfg::SLOTS) {
                const float2 wq = sm.ent[jj][i];
                const int p = sm.entp[jj][i];
                if (b_k < NQ) {
                    const float4 gv = sm.Gs[p][b_k < NQ ? b_k : 0];
                    a.x += wq.x * gv.x;
                    a.y += wq.x * gv.y;
                    a.z += wq.x * gv.z;
                    a.w += wq.x * gv.w;
                }
                if (b_k < 8) mom += wq.y * (sm.tabx[p & 15][b_k] * sm.taby[p >> 4][b_k]);
            }
#pragma unroll
            for (int o = Cfg::TPI; o < 32; o <<= 1) {
                a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
                a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
                a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
                a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
                mom += __shfl_xor_sync(0xffffffffu, mom, o);
            }
            // raw moments about the tile centre (lanes 0..5 of every slot group hold m0, mx, my, mxx, mxy, myy)
            const float m0 = __shfl_sync(0xffffffffu, mom, 0, Cfg::TPI);
            const float mx = __shfl_sync(0xffffffffu, mom, 1, Cfg::TPI);
            const float my = __shfl_sync(0xffffffffu, mom, 2, Cfg::TPI);
            const float mxx = __shfl_sync(0xffffffffu, mom, 3, Cfg::TPI);
            const float mxy = __shfl_sync(0xffffffffu, mom, 4, Cfg::TPI);
            const float myy = __shfl_sync(0xffffffffu, mom, 5, Cfg::TPI);
            if (b_slot == 0) {
                const uint32_t id = sm.ids[idb][jj];
                if (b_k < NQ) {
                    if (VEC) {
                        if (4 * b_k < K) red_add_v4(dL_dcolors + (size_t)id * K + 4 * b_k, a.x, a.y, a.z, a.w);
                    } else {
                        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const int ch = 4 * b_k + c;
                            if (COLOR && ch < K) red_add(dL_dcolors + (size_t)id * K + ch, av[c]);
                            else if (MD && ch == K) red_add(ggrad + (size_t)id * GG_STRIDE + 6, av[c]);
                        }
                    }
                }
                if (b_k < 6) {
                    const float4 g0 = sm.geo[stage][jj][0];
                    const float4 g1 = sm.geo[stage][jj][1];
                    const float conx = g0.z, cony = g0.w, conz = g1.x, o = g1.y;
                    // sums over the pixels of q * (1, dx, dy, dx^2, dx dy, dy^2) with d = centre - pixel = c - x'
                    const float cx = g0.x - tcx, cy = g0.y - tcy;
                    const float Sx = cx * m0 - mx;
                    const float Sy = cy * m0 - my;
                    const float Sxx = cx * cx * m0 - 2.f * cx * mx + mxx;
                    const float Sxy = cx * cy * m0 - cx * my - cy * mx + mxy;
                    const float Syy = cy * cy * m0 - 2.f * cy * my + myy;
                    float v;
                    int slot;
                    if (b_k == 0) { v = m0; slot = 5; }                                           // dL/dopacity
                    else if (b_k == 1) { v = -o * half_W * (conx * Sx + cony * Sy); slot = 0; }   // dL/dmean2D.x
                    else if (b_k == 2) { v = -o * half_H * (conz * Sy + cony * Sx); slot = 1; }   // dL/dmean2D.y
                    else if (b_k == 3) { v = -0.5f * o * Sxx; slot = 2; }                         // dL/dconic.x
                    else if (b_k == 4) { v = -0.5f * o * Sxy; slot = 3; }                         // dL/dconic.y
                    else { v = -0.5f * o * Syy; slot = 4; }                                       // dL/dconic.w
                    red_add(ggrad + (size_t)id * GG_STRIDE + slot, v);
                }
            }
        }

        // publish ids(b+2); wait for the copies of batch b+1
        if (have_next_id) sm.ids[(b + 2) % 3][tid] = next_id;
        cp_async_wait_all();
        if (b + 1 < nbatch) bwd_pad_geo<NQ>(sm, stage ^ 1, batch_cnt(b + 1));
        __syncthreads();
    }
}

}  // namespace sagars
