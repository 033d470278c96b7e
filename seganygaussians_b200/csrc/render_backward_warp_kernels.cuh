// render_backward_warp_kernels.cuh -- the device code of render_backward_warp.cu (see there).  Free of host-side runtime calls
// so that tests/test_warp_kernels_emulated.py can compile it for the CPU against tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include "cp_async.cuh"
#include "candidate.cuh"
#include "mma.cuh"

#ifndef SAGARS_DYNAMIC_SMEM
#define SAGARS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sagars {


constexpr int BW_RS = 36;     // row stride (words) of the W / Q tiles
constexpr int BW_N = 8;      // candidates per group = rows of the W / Q tiles = N of the colour product
constexpr int BW_TAB = 40;   // candidate table: up to 7 carried over + 32 new

template <int NQ>
struct BwCfg {
    static constexpr int NQE = NQ < 2 ? 2 : NQ;     // quads per gradient row (power of two)
    static constexpr int ROW = 4 * NQE;             // floats per gradient row
    static constexpr int MT = (ROW + 15) / 16;      // 16-channel m-tiles of the transposed colour product
};

template <int NQ>
struct BwSmem {
    // Gradient tile of the block (32 pixels x ROW channels) in the order the S product's A fragments read it: element (p, c) at
    //   i = (((c >> 3) * 2 + (p >> 4)) * 4 + ((p >> 3) & 1) + 2 * ((c >> 2) & 1)) * 32 + 4 * (p & 7) + (c & 3),   i ^= ((i >> 6) & 1) << 4
    // so that every fragment address of BOTH products (S = G F^T reads rows = pixels, dL/dcolour^T = G^T W^T reads rows =
    // channels) is one of two per-lane bases plus a compile-time offset, and both are free of bank conflicts.
    float Gs[32 * BwCfg<NQ>::ROW];
    // W / Q tiles: candidate row r, block pixel p at r * BW_RS + p.  The row stride of 36 words makes the scalar pass (lane = pixel,
    // row = immediate) and the fragment reads of the gradient product (lane = (candidate fg, pixel ft), 36 fg = 4 fg mod 32) both
    // conflict-free with addresses of the form per-lane base + immediate.
    float rowW[BW_N][BW_RS];        // also holds the gathered feature rows during (1)
    float rowQ[BW_N][BW_RS];        // holds S during (2)
    float4 ctab[BW_TAB][2];         // candidate records (x, y, cx, cy | cz, opacity, accept_threshold, -), list order
    uint32_t cid[BW_TAB];           // their Gaussian ids
};

// Moment basis of the geometry products: B fragment values X[p][mm] of fragment column mm = lane / 4 (1, x, y, x^2, x y, y^2, 0, 0)
// at the block-centred pixel (x, y) = (lane % 4 - 3.5 [+4], ks - 1.5): entry [lane][2 ks + (0: x, 1: x + 4)].  Multiples of 1/4,
// exact in tf32.  A table (two 16-byte L1 hits per group) because nvcc would otherwise rematerialise these eight values from
// threadIdx in every group (~45 instructions) rather than hold them in registers.
__device__ const float BW_MOMENT_BASIS[32][8] = {
    1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f,
    1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f,
    1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f,
    1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f,
    -3.5f, 0.5f, -3.5f, 0.5f, -3.5f, 0.5f, -3.5f, 0.5f,
    -2.5f, 1.5f, -2.5f, 1.5f, -2.5f, 1.5f, -2.5f, 1.5f,
    -1.5f, 2.5f, -1.5f, 2.5f, -1.5f, 2.5f, -1.5f, 2.5f,
    -0.5f, 3.5f, -0.5f, 3.5f, -0.5f, 3.5f, -0.5f, 3.5f,
    -1.5f, -1.5f, -0.5f, -0.5f, 0.5f, 0.5f, 1.5f, 1.5f,
    -1.5f, -1.5f, -0.5f, -0.5f, 0.5f, 0.5f, 1.5f, 1.5f,
    -1.5f, -1.5f, -0.5f, -0.5f, 0.5f, 0.5f, 1.5f, 1.5f,
    -1.5f, -1.5f, -0.5f, -0.5f, 0.5f, 0.5f, 1.5f, 1.5f,
    12.25f, 0.25f, 12.25f, 0.25f, 12.25f, 0.25f, 12.25f, 0.25f,
    6.25f, 2.25f, 6.25f, 2.25f, 6.25f, 2.25f, 6.25f, 2.25f,
    2.25f, 6.25f, 2.25f, 6.25f, 2.25f, 6.25f, 2.25f, 6.25f,
    0.25f, 12.25f, 0.25f, 12.25f, 0.25f, 12.25f, 0.25f, 12.25f,
    5.25f, -0.75f, 1.75f, -0.25f, -1.75f, 0.25f, -5.25f, 0.75f,
    3.75f, -2.25f, 1.25f, -0.75f, -1.25f, 0.75f, -3.75f, 2.25f,
    2.25f, -3.75f, 0.75f, -1.25f, -0.75f, 1.25f, -2.25f, 3.75f,
    0.75f, -5.25f, 0.25f, -1.75f, -0.25f, 1.75f, -0.75f, 5.25f,
    2.25f, 2.25f, 0.25f, 0.25f, 0.25f, 0.25f, 2.25f, 2.25f,
    2.25f, 2.25f, 0.25f, 0.25f, 0.25f, 0.25f, 2.25f, 2.25f,
    2.25f, 2.25f, 0.25f, 0.25f, 0.25f, 0.25f, 2.25f, 2.25f,
    2.25f, 2.25f, 0.25f, 0.25f, 0.25f, 0.25f, 2.25f, 2.25f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
    0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
};

// NQ : float4 groups covering the gradient channels (K colour channels [+ 1 mask channel when MD])
// VEC: K % 4 == 0 and no mask channel -> feature rows are read as float4
template <int NQ, bool VEC, bool MD, bool COLOR>
__global__ void __launch_bounds__(32, (NQ <= 8) ? 28 : 12)
render_backward_warp_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                            int W, int H, int K,
                            const float* __restrict__ bg, const float* __restrict__ geo,
                            const float* __restrict__ features,
                            const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                            const float* __restrict__ dL_dpix, const float* __restrict__ dL_dout_mask,
                            float* __restrict__ ggrad, float* __restrict__ dL_dcolors)
{
    using Cfg = BwCfg<NQ>;
    constexpr int NQE = Cfg::NQE, ROW = Cfg::ROW, MT = Cfg::MT;
    SAGARS_DYNAMIC_SMEM(smem_raw);
    BwSmem<NQ>& sm = *reinterpret_cast<BwSmem<NQ>*>(smem_raw);

    const int lane = threadIdx.x;
    const int tiles_x = (int)(gridDim.x >> 1);
    const uint32_t blk_x0 = blockIdx.x * 8, blk_y0 = blockIdx.y * 4;
    const uint32_t px = blk_x0 + (lane & 7), py = blk_y0 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    SAGARS_PIN_F2(pixx, pixy);   // keep nvcc from rematerialising them in the hot loop
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[(blockIdx.y >> 2) * tiles_x + (blockIdx.x >> 1)];
    const int total = (int)(range.y - range.x);
    if (total <= 0) return;   // empty tile

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;
    int blk_n = my_n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) blk_n = max(blk_n, __shfl_xor_sync(0xffffffffu, blk_n, o));
    const int maxc = min(blk_n, total);     // list positions [0, maxc) can contribute to this block
    if (maxc <= 0) return;
    const int nchunk = (maxc + 31) >> 5;

    // chunk c, lane l <-> list position maxc - 1 - 32 c - l (back to front)
    auto chunk_pos = [&](int c) { return maxc - 1 - 32 * c - lane; };
    // first chunk's id and record, then the gradient row of this pixel: independent latency chains
    int pos_cur = chunk_pos(0);
    uint32_t id_cur = pos_cur >= 0 ? point_list[range.x + pos_cur] : 0u;

    float bgdot = 0.f;
    {
        float gmask = 0.f;
        if (MD) gmask = inside ? dL_dout_mask[pix_id] : 0.f;
        float gr[ROW];
#pragma unroll
        for (int k = 0; k < ROW; k++) {
            float x = 0.f;
            if (COLOR && k < K) x = inside ? dL_dpix[(size_t)k * plane + pix_id] : 0.f;
            if (MD && k == K) x = gmask;   // the mask gradient rides as channel K of the colour product
            gr[k] = x;
        }
        const int wl0 = 128 * (lane >> 4) + 32 * ((lane >> 3) & 1) + 4 * (lane & 7), wl1 = wl0 ^ 16;
#pragma unroll
        for (int k = 0; k < ROW; k++) {
            if (COLOR && k < K) bgdot += bg[k] * gr[k];
            sm.Gs[(((k >> 2) & 1) ? wl1 : wl0) + (k >> 3) * 256 + 64 * ((k >> 2) & 1) + (k & 3)] = gr[k];
        }
    }
    float4 r0_cur = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_cur));
    float4 r1_cur = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_cur + 4));
    {   // rows past the fill level are multiplied too (their products are never used): start from finite values
        float* w = &sm.rowW[0][0];
        float* q = &sm.rowQ[0][0];
#pragma unroll
        for (int i = 0; i < BW_N; i++) { w[i * BW_RS + lane] = 0.f; q[i * BW_RS + lane] = 0.f; }
    }
    __syncwarp();

    float T = T_final;
    const bool any_bg = __any_sync(0xffffffffu, bgdot != 0.f);   // zero background: the term is exactly 0
    float u = 0.f;   // see the scalar pass
    float* const rowW = &sm.rowW[0][0];
    float* const rowQ = &sm.rowQ[0][0];

    // mma fragment coordinates of this lane
    const int fg = lane >> 2, ft = lane & 3;
    // per-lane bases into the gradient tile (see BwSmem::Gs)
    const float* const gS0 = sm.Gs + lane;
    const float* const gS1 = sm.Gs + (lane ^ 16);
    const float* const gD0 = sm.Gs + 64 * (fg >> 2) + 4 * ft + (fg & 3) + 16 * (fg >> 2);
    const float* const gD1 = sm.Gs + 64 * (fg >> 2) + 4 * ft + (fg & 3) + 16 * (1 - (fg >> 2));
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;   // (0.5 * W) rounded to float, as the reference
    const float bx0 = (float)blk_x0, bx1 = bx0 + 7.f, by0 = (float)blk_y0, by1 = by0 + 3.f;   // block of pixel centres
    const float bcx = bx0 + 3.5f, bcy = by0 + 1.5f;                                            // its centre
    const uint32_t lt = (1u << lane) - 1u;

    // one group: candidates in table slots [gs, gs + m), m <= 8
    auto process_group = [&](int gs, int m) {
        // ---- (1) S (32 pixels x 8) = G (32 x C) * F^T (C x 8): feature rows gathered into the W tile's storage
        //          (B-fragment order, see the gather), S written into the Q tile's storage in the Q layout ----
        if (COLOR) {
            float sacc[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) sacc[mt][0] = sacc[mt][1] = sacc[mt][2] = sacc[mt][3] = 0.f;
            constexpr int QR = (ROW < 32 ? ROW : 32) / 4;      // feature quads per row and channel block
#pragma unroll 1
            for (int cb = 0; cb < ROW; cb += 32) {
                if (cb > 0) __syncwarp();
                for (int idx = lane; idx < BW_N * QR; idx += 32) {
                    // feature tile in B-fragment order: quad qd of candidate r at qd * 32 + 4 r (8 lanes = 8 candidates write one
                    // 128-byte line; the fragment of k-step ks is line 2 ks [+1], word = lane)
                    const int r = idx & (BW_N - 1), qd = idx / BW_N;
                    const uint32_t id = sm.cid[gs + min(r, m - 1)];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int c0 = cb + 4 * qd;
                    if (VEC && K == ROW) {      // full float4 rows: straight into the tile
                        cp_async16(rowW + qd * 32 + 4 * r, features + (size_t)id * K + c0);
                        continue;
                    }
                    if (VEC) {
                        if (c0 < K) v = __ldg(reinterpret_cast<const float4*>(features + (size_t)id * K + c0));
                    } else {
                        const float* f = features + (size_t)id * K;
                        if (c0 + 0 < K) v.x = __ldg(f + c0 + 0);
                        if (c0 + 1 < K) v.y = __ldg(f + c0 + 1);
                        if (c0 + 2 < K) v.z = __ldg(f + c0 + 2);
                        if (c0 + 3 < K) v.w = __ldg(f + c0 + 3);
                    }
                    *reinterpret_cast<float4*>(rowW + qd * 32 + 4 * r) = v;
                }
                if (VEC && K == ROW) { cp_async_commit(); cp_async_wait_all(); }
                __syncwarp();
                const float* Fr = rowW + lane;
#pragma unroll
                for (int ks = 0; ks < QR / 2; ks++) {
                    uint32_t bh0, bl0, bh1, bl1;
                    split_tf32(Fr[64 * ks], bh0, bl0);
                    split_tf32(Fr[64 * ks + 32], bh1, bl1);
                    const int ch0 = cb + ks * 8 + ft, ch1 = ch0 + 4;
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const int p0 = 16 * mt + fg, p1 = p0 + 8;      // block pixels of fragment rows g and g + 8
                        uint32_t ah[4], al[4];
                        const int fo = (((cb >> 3) + ks) * 2 + mt) * 128;      // (pixel 16 mt + fg [+8], channel cb + 8 ks + ft [+4])
                        split_tf32(gS0[fo], ah[0], al[0]);
                        split_tf32(gS0[fo + 32], ah[1], al[1]);
                        split_tf32(gS1[fo + 64], ah[2], al[2]);
                        split_tf32(gS1[fo + 96], ah[3], al[3]);
                        mma_16n8k8(sacc[mt], al[0], al[1], al[2], al[3], bh0, bh1);
                        mma_16n8k8(sacc[mt], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                        mma_16n8k8(sacc[mt], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                    }
                }
            }
            __syncwarp();   // the feature rows are consumed; (2) rewrites this storage row by row
            // fragment (pixel 16 mt + fg [+8], candidates 2 ft, 2 ft + 1) -> S[i][p] in the Q tile (row stride BW_RS)
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const int pa = 16 * mt + fg, pb = pa + 8;
                rowQ[(2 * ft) * BW_RS + pa] = sacc[mt][0];
                rowQ[(2 * ft + 1) * BW_RS + pa] = sacc[mt][1];
                rowQ[(2 * ft) * BW_RS + pb] = sacc[mt][2];
                rowQ[(2 * ft + 1) * BW_RS + pb] = sacc[mt][3];
            }
            __syncwarp();
        }
        // ---- (2) thread = pixel over the group's candidates (the reference's traversal): row i of W / Q ----
        //      The chain as selects (no divergent branch; a rejected pair computes on garbage and selects nothing), full groups
        //      unrolled so that table / tile addresses are immediates.  1 / (1 - alpha) through rcp.approx (<= 1 ulp; IEEE
        //      division costs a slow-path branch per pair); the background term is skipped when no pixel of the block has one.
        auto one = [&](int i) {
            const float4 g0 = sm.ctab[gs + i][0];
            const float4 g1 = sm.ctab[gs + i][1];
            const float dx = g0.x - pixx, dy = g0.y - pixy;
            const float pw = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
            const float G0 = expf(pw);
            const float alpha0 = fminf(0.99f, g1.y * G0);
            const bool cd = (__float_as_int(g1.w) < my_n) && !(pw > 0.0f) && (pw >= g1.z) && !(alpha0 < 1.0f / 255.0f);
            // a pair that does not blend runs the same arithmetic with G = alpha = 0: u and the products then keep / produce
            // exactly what they should (u' = 0 s + 1 u, w = 0, q = 0); only T is selected (rcp.approx(1) need not be exactly 1)
            const float G = cd ? G0 : 0.f, alpha = cd ? alpha0 : 0.f;
            const float oma = 1.f - alpha;
            const float rcp = rcp_approx(oma);
            const float Tn = T * rcp;
            const float s = COLOR ? rowQ[i * BW_RS + lane] : 0.f;
            // u = "colour behind" dotted with the pixel's gradient: the reference's accum_rec recurrence (last_alpha * last_color +
            // (1 - last_alpha) * accum_rec) evaluated when a pair is accepted instead of when the next one is
            float dL_dalpha = (s - u) * Tn;
            if (any_bg) dL_dalpha += (-T_final * rcp) * bgdot;
            u = alpha * s + oma * u;
            T = cd ? Tn : T;
            rowW[i * BW_RS + lane] = alpha * Tn;          // all lanes write: zero where the pixel did not blend
            rowQ[i * BW_RS + lane] = G * dL_dalpha;
        };
        if (m == BW_N) {
#pragma unroll
            for (int i = 0; i < BW_N; i++) one(i);
        } else {
#pragma unroll 1
            for (int i = 0; i < m; i++) one(i);
        }
        __syncwarp();

        // ---- (3) dL/dcolour^T = G^T W^T, moments = Q X; results -> global memory ----
        float dc[MT][4];
#pragma unroll
        for (int mm = 0; mm < MT; mm++) dc[mm][0] = dc[mm][1] = dc[mm][2] = dc[mm][3] = 0.f;
        float dm[4] = {0.f, 0.f, 0.f, 0.f};
        {
            // this lane's row of the tiles: column n = fg of W^T, row fg of Q; pixel p sits at word p of the row (ft is in the base)
            const float* Wr = rowW + fg * BW_RS + ft;
            const float* Qr = rowQ + fg * BW_RS + ft;
            const float4 xb_lo = __ldg(reinterpret_cast<const float4*>(&BW_MOMENT_BASIS[lane][0]));
            const float4 xb_hi = __ldg(reinterpret_cast<const float4*>(&BW_MOMENT_BASIS[lane][4]));
            const float xv[8] = {xb_lo.x, xb_lo.y, xb_lo.z, xb_lo.w, xb_hi.x, xb_hi.y, xb_hi.z, xb_hi.w};
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const int c0 = ks * 8, c1 = c0 + 4;                      // block pixels ks*8 + ft and + 4 (ft is in the base)
                uint32_t wh0, wl0, wh1, wl1, qh0, ql0, qh1, ql1;
                split_tf32(Wr[c0], wh0, wl0);
                split_tf32(Wr[c1], wh1, wl1);
                split_tf32(Qr[c0], qh0, ql0);
                split_tf32(Qr[c1], qh1, ql1);
                const int r0 = ks * 8 + ft, r1 = r0 + 4;     // block pixels ks*8 + ft and + 4 (same row, x and x + 4)
                if (COLOR || MD) {
#pragma unroll
                    for (int mm = 0; mm < MT; mm++) {
                        const int cl = 16 * mm + fg, chh = cl + 8;     // channels of fragment rows g and g + 8
                        uint32_t ah[4], al[4];
                        const int go = ((4 * mm + (ks >> 1)) * 4 + (ks & 1)) * 32;   // (channel 16 mm + fg, pixel 8 ks + ft [+4])
                        split_tf32(gD0[go], ah[0], al[0]);
                        split_tf32(gD1[go], ah[2], al[2]);
                        if (16 * mm + 8 < ROW) {
                            split_tf32(gD0[go + 256], ah[1], al[1]);                 // channel + 8: next 8-channel step
                            split_tf32(gD1[go + 256], ah[3], al[3]);
                        } else {
                            ah[1] = al[1] = ah[3] = al[3] = 0u;
                        }
                        mma_16n8k8(dc[mm], al[0], al[1], al[2], al[3], wh0, wh1);
                        mma_16n8k8(dc[mm], ah[0], ah[1], ah[2], ah[3], wl0, wl1);
                        mma_16n8k8(dc[mm], ah[0], ah[1], ah[2], ah[3], wh0, wh1);
                    }
                }
                const float v0 = xv[2 * ks], v1 = xv[2 * ks + 1];   // basis at pixels (ks*8 + ft, ks*8 + ft + 4)
                mma_16n8k8(dm, ql0, 0u, ql1, 0u, __float_as_uint(v0), __float_as_uint(v1));
                mma_16n8k8(dm, qh0, 0u, qh1, 0u, __float_as_uint(v0), __float_as_uint(v1));
            }
        }
        // colour product: this lane holds channels (16 mm + fg, + 8) of rows 2 ft and 2 ft + 1
        if (COLOR || MD) {
            const uint32_t ida = sm.cid[gs + min(2 * ft, m - 1)], idb = sm.cid[gs + min(2 * ft + 1, m - 1)];
            const bool va = 2 * ft < m, vb = 2 * ft + 1 < m;
            auto emit = [&](uint32_t id, int ch, float v) {
                if (COLOR && ch < K) red_add(dL_dcolors + (size_t)id * K + ch, v);
                else if (MD && ch == K) red_add(ggrad + (size_t)id * GG_STRIDE + 6, v);
            };
            if (COLOR && !MD && ROW >= 16 && K == ROW) {     // full rows (warp-uniform): one base per candidate, no per-channel bounds
                float* const pa = dL_dcolors + (size_t)ida * K + fg;
                float* const pb = dL_dcolors + (size_t)idb * K + fg;
                if (va) {
#pragma unroll
                    for (int mm = 0; mm < MT; mm++) { red_add(pa + 16 * mm, dc[mm][0]); red_add(pa + 16 * mm + 8, dc[mm][2]); }
                }
                if (vb) {
#pragma unroll
                    for (int mm = 0; mm < MT; mm++) { red_add(pb + 16 * mm, dc[mm][1]); red_add(pb + 16 * mm + 8, dc[mm][3]); }
                }
            } else {
#pragma unroll
                for (int mm = 0; mm < MT; mm++) {
                    const int cl = 16 * mm + fg;
                    if (va) { emit(ida, cl, dc[mm][0]); emit(ida, cl + 8, dc[mm][2]); }
                    if (vb) { emit(idb, cl, dc[mm][1]); emit(idb, cl + 8, dc[mm][3]); }
                }
            }
        }
        // moments of row fg: (m0, mx) in lane ft = 0, (my, mxx) in ft = 1, (mxy, myy) in ft = 2 of the quad
        {
            const int q0 = lane & ~3;
            const float m0 = __shfl_sync(0xffffffffu, dm[0], q0);
            const float mx = __shfl_sync(0xffffffffu, dm[1], q0);
            const float my = __shfl_sync(0xffffffffu, dm[0], q0 + 1);
            const float mxx = __shfl_sync(0xffffffffu, dm[1], q0 + 1);
            const float mxy = __shfl_sync(0xffffffffu, dm[0], q0 + 2);
            const float myy = __shfl_sync(0xffffffffu, dm[1], q0 + 2);
            if (fg < m && ft < 3) {
                const uint32_t id = sm.cid[gs + fg];
                const float4 g0 = sm.ctab[gs + fg][0];
                const float4 g1 = sm.ctab[gs + fg][1];
                const float conx = g0.z, cony = g0.w, conz = g1.x, o = g1.y;
                // sums over the pixels of q * (1, dx, dy, dx^2, dx dy, dy^2) with d = centre - pixel = c - x'
                const float cx = g0.x - bcx, cy = g0.y - bcy;
                // every lane evaluates all five sums and selects its two outputs: the three-way split by ft would run serially
                const float Sx = cx * m0 - mx;
                const float Sy = cy * m0 - my;
                const float Sxx = cx * cx * m0 - 2.f * cx * mx + mxx;
                const float Sxy = cx * cy * m0 - cx * my - cy * mx + mxy;
                const float Syy = cy * cy * m0 - 2.f * cy * my + myy;
                const float dmx = -o * half_W * (conx * Sx + cony * Sy);              // dL/dmean2D.x
                const float dmy = -o * half_H * (conz * Sy + cony * Sx);              // dL/dmean2D.y
                const float hno = -0.5f * o;
                // ft = 0: (dL/dopacity -> 5, dmean2D.x -> 0); 1: (dmean2D.y -> 1, dconic.x -> 2); 2: (dconic.y -> 3, dconic.w -> 4)
                const float ua = (ft == 0) ? m0 : (ft == 1) ? dmy : hno * Sxy;
                const float ub = (ft == 0) ? dmx : (ft == 1) ? hno * Sxx : hno * Syy;
                const int sa = (ft == 0) ? 5 : 2 * ft - 1, sb = 2 * ft;
                red_add(ggrad + (size_t)id * GG_STRIDE + sa, ua);
                red_add(ggrad + (size_t)id * GG_STRIDE + sb, ub);
            }
        }
        __syncwarp();   // every lane is done with the tiles and the table slots before they are overwritten
    };

    int ntab = 0;   // candidates waiting in table slots [0, ntab)
    for (int c = 0; c < nchunk; c++) {
        // the next chunk's id, then its record, are in flight while this chunk is worked on
        const int pos_nxt = (c + 1 < nchunk) ? chunk_pos(c + 1) : -1;
        const uint32_t id_nxt = pos_nxt >= 0 ? point_list[range.x + pos_nxt] : 0u;

        // block-level candidate test (candidate.cuh), lane = splat; survivors join the table in list order
        const bool keep = pos_cur >= 0 && !block_rejects<true>(r0_cur, r1_cur, bx0, bx1, by0, by1);
        const uint32_t km = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int slot = ntab + __popc(km & lt);
            sm.ctab[slot][0] = r0_cur;
            float4 r1p = r1_cur;
            r1p.w = __int_as_float(pos_cur);
            sm.ctab[slot][1] = r1p;
            sm.cid[slot] = id_cur;
        }
        ntab += __popc(km);
        float4 r0_nxt = make_float4(0.f, 0.f, 0.f, 0.f), r1_nxt = r0_nxt;
        if (pos_nxt >= 0) {
            r0_nxt = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_nxt));
            r1_nxt = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_nxt + 4));
        }
        __syncwarp();

        // full groups now, the rest is carried over (the last chunk flushes everything)
        int gs = 0;
        const bool last = (c + 1 == nchunk);
        while (ntab - gs >= BW_N || (last && ntab - gs > 0)) {
            process_group(gs, min(BW_N, ntab - gs));
            gs += BW_N;
        }
        if (gs > 0 && gs < ntab) {   // carry the leftovers (< 8) to the front: sources are slots >= 8, destinations < 7
            const int left = ntab - gs;
            float4 a0, a1;
            uint32_t ci = 0;
            if (lane < left) { a0 = sm.ctab[gs + lane][0]; a1 = sm.ctab[gs + lane][1]; ci = sm.cid[gs + lane]; }
            __syncwarp();
            if (lane < left) { sm.ctab[lane][0] = a0; sm.ctab[lane][1] = a1; sm.cid[lane] = ci; }
            __syncwarp();
        }
        ntab = (gs >= ntab) ? 0 : ntab - gs;

        pos_cur = pos_nxt;
        id_cur = id_nxt;
        r0_cur = r0_nxt;
        r1_cur = r1_nxt;
    }
}

}  // namespace sagars
