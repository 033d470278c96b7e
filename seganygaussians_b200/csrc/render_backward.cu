// render_backward.cu -- per-tile back-to-front gradient of the alpha compositing.
//
// Semantics: CF cuda_rasterizer/backward.cu:399-559 (DEPTH backward.cu:400-564 adds dL_dmask),
// SURVEY.md Appendix A.13-A.17.  The reference gives every (pixel, Gaussian) pair C+6 scalar global
// atomicAdds; all 256 threads of a CTA hit the same C+6 addresses.  This kernel is organised
// differently -- two phases per batch of NB instances, both inside one CTA per 16x16 tile:
//
//   phase A (thread = pixel, exactly the reference's traversal): recompute alpha, undo T, and reduce
//     the C-channel work of a pair to ONE dot product s = f_j . g_p.  Because dL/dalpha is linear in
//     the upstream gradient, the reference's per-channel recurrence accum_rec[ch] collapses to a
//     scalar recurrence on a = accum_rec . g_p (same association order, Appendix D).  Each blended
//     pair emits two scalars into shared memory: w = alpha*T (weight of dL/dcolour) and
//     q = G * dL/dalpha (weight of every geometric gradient), plus a ballot bit.
//   phase B (thread = (instance, channel quad)): every per-Gaussian gradient is a sparse
//     matrix product over the tile's pixels,
//         dL/dcolour[j][:] = sum_p w[j][p] * g[p][:]          (g tile resident in smem, float4 reads)
//         moments[j][:]    = sum_p q[j][p] * (1, dx, dy, dx^2, dx*dy, dy^2)
//     accumulated in registers by walking the ballot bits, so there is no cross-lane reduction and no
//     shared-memory atomic; the six moments give dL/dopacity, dL/dmean2D and dL/dconic in closed form.
//   One vectorised `red.global.add.v4.f32` per (tile, Gaussian, channel quad) and six scalar reds per
//   (tile, Gaussian) then replace the reference's (C+6) x (blended pixels) atomics.
//
// Feature rows and instance records are staged with cp.async, double buffered.
#include "common.cuh"
#include "cp_async.cuh"

namespace sagars {

constexpr int BWD_NB = 16;   // instances per batch

template <int NQ>
struct BwdCfg {
    static constexpr int TPI = NQ > 8 ? NQ : 8;                 // phase-B threads per instance
    static constexpr int SPLIT = TILE_PIX / (BWD_NB * TPI);     // pixel-range splits per instance
    static constexpr int WARPS_PER_SPLIT = 8 / SPLIT;
};

template <int NQ>
struct BwdSmem {
    float4 Gs[TILE_PIX][NQ];            // upstream gradient rows of the tile's pixels (padded channels = 0)
    float2 Wq[BWD_NB][TILE_PIX];        // (w, q) per (instance, pixel); valid where the ballot bit is set
    uint32_t masks[BWD_NB][8];          // ballot of blended pixels per (instance, warp)
    float4 geo[2][BWD_NB][2];           // x, y, cx, cy | cz, opacity, depth, -
    float4 feat[2][BWD_NB][NQ];         // feature rows, zero padded
    uint32_t ids[3][BWD_NB];
    uint32_t max_contrib;
};

template <int NQ, bool VEC>
__device__ __forceinline__ void bwd_issue_batch(BwdSmem<NQ>& sm, int stage, int idbuf, int cnt, int K,
                                                const float* __restrict__ geo, const float* __restrict__ features)
{
    const int tid = threadIdx.x;
    if (tid < cnt * 2) {
        const int j = tid >> 1, h = tid & 1;
        const uint32_t id = sm.ids[idbuf][j];
        cp_async16(&sm.geo[stage][j][h], geo + 8 * (size_t)id + 4 * h);
    }
    if (VEC) {
        const int nq = K >> 2;
        for (int c = tid; c < cnt * nq; c += TILE_PIX) {
            const int j = c / nq, q = c - j * nq;
            const uint32_t id = sm.ids[idbuf][j];
            cp_async16(&sm.feat[stage][j][q], features + (size_t)id * K + 4 * q);
        }
    } else {
        float* f = reinterpret_cast<float*>(&sm.feat[stage][0][0]);
        for (int c = tid; c < cnt * K; c += TILE_PIX) {
            const int j = c / K, k = c - j * K;
            const uint32_t id = sm.ids[idbuf][j];
            f[j * (4 * NQ) + k] = features[(size_t)id * K + k];
        }
    }
}

// NQ : float4 groups covering the gradient channels (K colour channels [+ 1 mask channel when MD])
// VEC: K % 4 == 0 and no mask channel -> dL_dcolors rows are 16-byte aligned, use red.v4
template <int NQ, bool VEC, bool MD, bool COLOR>
__global__ void __launch_bounds__(TILE_PIX)
render_backward_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                       int W, int H, int K,
                       const float* __restrict__ bg, const float* __restrict__ geo,
                       const float* __restrict__ features,
                       const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                       const float* __restrict__ dL_dpix, const float* __restrict__ dL_dout_mask,
                       float* __restrict__ ggrad, float* __restrict__ dL_dcolors)
{
    using Cfg = BwdCfg<NQ>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem<NQ>& sm = *reinterpret_cast<BwdSmem<NQ>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = gridDim.x;
    const uint32_t tile_x0 = blockIdx.x * TILE_X, tile_y0 = blockIdx.y * TILE_Y;
    const uint32_t px = tile_x0 + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    const float pixx = (float)px, pixy = (float)py;
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;

    // the tile only needs instances [0, max over its pixels of n_contrib)
    if (tid == 0) sm.max_contrib = 0;
    __syncthreads();
    {
        int m = my_n;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0) atomicMax(&sm.max_contrib, (uint32_t)m);
    }

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
        sm.Gs[tid][q] = v;
    }
    // zero the padded feature channels once
    if (!VEC || (K >> 2) < NQ) {
        float* f = reinterpret_cast<float*>(&sm.feat[0][0][0]);
        for (int c = tid; c < 2 * BWD_NB * 4 * NQ; c += TILE_PIX) f[c] = 0.f;
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
    bwd_issue_batch<NQ, VEC>(sm, 0, 0, batch_cnt(0), K, geo, features);
    cp_async_commit();
    if (nbatch > 1 && tid < batch_cnt(1)) sm.ids[1][tid] = load_id(1, tid);
    cp_async_wait_all();
    __syncthreads();

    float T = T_final;
    float acc = 0.f, last_alpha = 0.f, last_s = 0.f;

    // phase-B role of this thread
    const int b_split = tid / (BWD_NB * Cfg::TPI);
    const int b_jj = (tid % (BWD_NB * Cfg::TPI)) / Cfg::TPI;
    const int b_k = tid % Cfg::TPI;
    // moment basis of lane k: (a0 + a1 dx + a2 dy) * (b0 + b1 dx + b2 dy)
    const float ma0 = (b_k == 0) ? 1.f : 0.f;
    const float ma1 = (b_k == 1 || b_k == 3 || b_k == 4) ? 1.f : 0.f;
    const float ma2 = (b_k == 2 || b_k == 5) ? 1.f : 0.f;
    const float mb0 = (b_k <= 2) ? 1.f : 0.f;
    const float mb1 = (b_k == 3) ? 1.f : 0.f;
    const float mb2 = (b_k == 4 || b_k == 5) ? 1.f : 0.f;
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;   // (0.5 * W) rounded to float, as the reference

    for (int b = 0; b < nbatch; b++) {
        const int stage = b & 1;
        const int idb = b % 3;
        const int cnt = batch_cnt(b);
        const int pos_hi = maxc - 1 - b * BWD_NB;

        // (A) copies of batch b+1, (B) ids of batch b+2
        if (b + 1 < nbatch) {
            bwd_issue_batch<NQ, VEC>(sm, stage ^ 1, (b + 1) % 3, batch_cnt(b + 1), K, geo, features);
            cp_async_commit();
        }
        uint32_t next_id = 0;
        const bool have_next_id = (b + 2 < nbatch) && tid < batch_cnt(b + 2);
        if (have_next_id) next_id = load_id(b + 2, tid);

        // ---------------- phase A: thread = pixel ----------------
        for (int jj = 0; jj < cnt; jj++) {
            bool blended = false;
            if (pos_hi - jj < my_n) {
                const float4 g0 = sm.geo[stage][jj][0];
                const float4 g1 = sm.geo[stage][jj][1];
                const float dx = g0.x - pixx, dy = g0.y - pixy;
                const float power = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                if (!(power > 0.0f)) {
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, g1.y * G);
                    if (!(alpha < 1.0f / 255.0f)) {
                        T = T / (1.f - alpha);
                        const float w = alpha * T;
                        float s = 0.f;
                        if (COLOR) {
#pragma unroll
                            for (int q = 0; q < NQ; q++) {
                                const float4 f = sm.feat[stage][jj][q];
                                s += f.x * g[4 * q + 0];
                                s += f.y * g[4 * q + 1];
                                s += f.z * g[4 * q + 2];
                                s += f.w * g[4 * q + 3];
                            }
                        }
                        acc = last_alpha * last_s + (1.f - last_alpha) * acc;
                        last_s = s;
                        float dL_dalpha = (s - acc) * T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
                        sm.Wq[jj][tid] = make_float2(w, G * dL_dalpha);
                        blended = true;
                    }
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, blended);
            if (lane == 0) sm.masks[jj][warp] = m;
        }
        __syncthreads();

        // ---------------- phase B: thread = (instance, channel quad / moment) ----------------
        {
            const bool active = b_jj < cnt;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            float mom = 0.f;
            uint32_t any = 0;
            float xg = 0.f, yg = 0.f;
            if (active) {
                const float4 g0 = sm.geo[stage][b_jj][0];
                xg = g0.x;
                yg = g0.y;
#pragma unroll
                for (int wi = 0; wi < Cfg::WARPS_PER_SPLIT; wi++) {
                    const int wrp = b_split * Cfg::WARPS_PER_SPLIT + wi;
                    uint32_t m = sm.masks[b_jj][wrp];
                    any |= m;
                    const float wx = (float)(tile_x0 + (wrp & 1) * 8);
                    const float wy = (float)(tile_y0 + (wrp >> 1) * 4);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const int p = wrp * 32 + l;
                        const float2 wq = sm.Wq[b_jj][p];
                        if (b_k < NQ) {
                            const float4 gv = sm.Gs[p][b_k < NQ ? b_k : 0];
                            a.x += wq.x * gv.x;
                            a.y += wq.x * gv.y;
                            a.z += wq.x * gv.z;
                            a.w += wq.x * gv.w;
                        }
                        const float dx = xg - (wx + (float)(l & 7));
                        const float dy = yg - (wy + (float)(l >> 3));
                        mom += wq.y * ((ma0 + ma1 * dx + ma2 * dy) * (mb0 + mb1 * dx + mb2 * dy));
                    }
                }
            }
            // moments of lanes 1 and 2 of this instance's lane group (dmean needs both)
            const float m1 = __shfl_sync(0xffffffffu, mom, 1, Cfg::TPI);
            const float m2 = __shfl_sync(0xffffffffu, mom, 2, Cfg::TPI);
            if (active && any) {
                const uint32_t id = sm.ids[idb][b_jj];
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
                    const float4 g0 = sm.geo[stage][b_jj][0];
                    const float4 g1 = sm.geo[stage][b_jj][1];
                    const float cx = g0.z, cy = g0.w, cz = g1.x, o = g1.y;
                    float v;
                    int slot;
                    if (b_k == 0) { v = mom; slot = 5; }                                          // dL/dopacity
                    else if (b_k == 1) { v = -o * half_W * (cx * m1 + cy * m2); slot = 0; }       // dL/dmean2D.x
                    else if (b_k == 2) { v = -o * half_H * (cz * m2 + cy * m1); slot = 1; }       // dL/dmean2D.y
                    else { v = -0.5f * o * mom; slot = b_k - 1; }                                 // dL/dconic x,y,w
                    red_add(ggrad + (size_t)id * GG_STRIDE + slot, v);
                }
            }
        }

        // (D) publish ids(b+2); wait for the copies of batch b+1
        if (have_next_id) sm.ids[(b + 2) % 3][tid] = next_id;
        cp_async_wait_all();
        __syncthreads();
    }
}

template <int NQ, bool VEC, bool MD, bool COLOR>
static int launch_bwd_t(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                        const uint32_t* point_list, const float* features, float* ggrad, cudaStream_t s, bool debug)
{
    auto kern = render_backward_kernel<NQ, VEC, MD, COLOR>;
    const size_t smem = sizeof(BwdSmem<NQ>);
    SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(d.tiles_x, d.tiles_y);
    kern<<<grid, TILE_PIX, smem, s>>>(im.ranges, point_list, d.W, d.H, d.C, a.background, g.geo, features,
                                      im.final_T, im.n_contrib, a.dL_dout_color, a.dL_dout_mask, ggrad,
                                      a.dL_dcolors);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_render_backward(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                           const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug)
{
    const bool md = (a.flags & SAGARS_FLAG_MASK_DEPTH) != 0;
    const bool mask_only = (a.flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    if (mask_only) return launch_bwd_t<1, false, true, false>(a, d, g, im, point_list, features, ggrad, s, debug);
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define SAGARS_BWD_CASE(NQ_)                                                                                     \
    if (nq <= NQ_) {                                                                                             \
        if (md) return launch_bwd_t<NQ_, false, true, true>(a, d, g, im, point_list, features, ggrad, s, debug);  \
        return vec ? launch_bwd_t<NQ_, true, false, true>(a, d, g, im, point_list, features, ggrad, s, debug)     \
                   : launch_bwd_t<NQ_, false, false, true>(a, d, g, im, point_list, features, ggrad, s, debug);   \
    }
    SAGARS_BWD_CASE(1)
    SAGARS_BWD_CASE(2)
    SAGARS_BWD_CASE(4)
    SAGARS_BWD_CASE(8)
    SAGARS_BWD_CASE(16)
#undef SAGARS_BWD_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
