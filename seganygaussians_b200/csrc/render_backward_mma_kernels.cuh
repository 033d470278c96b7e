// render_backward_mma_kernels.cuh -- the device code of render_backward_mma.cu (see there); free of host-side runtime calls so
// that the CPU suite can run it under tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include "cp_async.cuh"
#include "candidate.cuh"
#include "mma.cuh"

#ifndef SAGARS_DYNAMIC_SMEM
#define SAGARS_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace sagars {


constexpr int BM_NB = 64;     // splats per staged batch (two 32-wide candidate masks per warp)
constexpr int BM_N = 8;       // rows a warp collects per GEMM = N of the colour tiles = used M of the moment tile

template <int NQ>
struct BmCfg {
    static constexpr int NQE = NQ < 2 ? 2 : NQ;     // quads per gradient row (power of two)
    static constexpr int ROW = 4 * NQE;             // floats per gradient row
    static constexpr int MT = (ROW + 15) / 16;      // 16-channel m-tiles of the transposed colour product
};

template <int NQ>
struct BmSmem {
    float Gs[TILE_PIX][BmCfg<NQ>::ROW];             // gradient rows by raster-local pixel; quad q of row r lives at quad (q + r) % NQE
    float rowW[8][BM_N][32];                        // per warp: row r, pixel p at (p + 4 r) & 31
    float rowQ[8][BM_N][32];
    uint32_t row_id[8][BM_N];                       // Gaussian index of the row
    uint8_t clist[8][BM_NB];                        // per warp: batch-local indices of its candidate splats, in list order
    float4 geo[2][BM_NB][2];                        // staged records: x, y, cx, cy | cz, opacity, accept_threshold, -
    float4 feat[2][BM_NB][NQ];                      // staged feature rows, zero padded
    uint32_t ids[3][BM_NB];
    uint32_t warp_max[8];                           // deepest contributing list position per warp
};

template <int NQ>
__device__ __forceinline__ void bm_issue_batch(BmSmem<NQ>& sm, int stage, int idbuf, int cnt, int K, bool vec, bool color,
                                               const float* __restrict__ geo, const float* __restrict__ features)
{
    const int tid = threadIdx.x;
    if (tid < cnt * 2) {
        const int j = tid >> 1, h = tid & 1;
        cp_async16(&sm.geo[stage][j][h], geo + 8 * (size_t)sm.ids[idbuf][j] + 4 * h);
    }
    if (!color) return;
    if (vec) {
        const int nq = K >> 2;
        for (int c = tid; c < cnt * nq; c += TILE_PIX) {
            const int j = c / nq, q = c - j * nq;
            cp_async16(&sm.feat[stage][j][q], features + (size_t)sm.ids[idbuf][j] * K + 4 * q);
        }
    } else {
        float* f = reinterpret_cast<float*>(&sm.feat[stage][0][0]);
        for (int c = tid; c < cnt * K; c += TILE_PIX) {
            const int j = c / K, k = c - j * K;
            f[j * (4 * NQ) + k] = features[(size_t)sm.ids[idbuf][j] * K + k];
        }
    }
}

template <int NQ>
__device__ __forceinline__ void bm_pad_geo(BmSmem<NQ>& sm, int stage, int cnt)
{
    const int tid = threadIdx.x;
    if (tid >= cnt && tid < BM_NB) {   // records past the end of the batch: never accepted (threshold = +inf)
        sm.geo[stage][tid][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.geo[stage][tid][1] = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
    }
}

// NQ : float4 groups covering the gradient channels (K colour channels [+ 1 mask channel when MD])
// VEC: K % 4 == 0 and no mask channel -> feature rows are 16-byte aligned, staged with cp.async
template <int NQ, bool VEC, bool MD, bool COLOR>
__global__ void __launch_bounds__(TILE_PIX, (NQ <= 8) ? 3 : 1)
render_backward_mma_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                           int W, int H, int K,
                           const float* __restrict__ bg, const float* __restrict__ geo,
                           const float* __restrict__ features,
                           const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                           const float* __restrict__ dL_dpix, const float* __restrict__ dL_dout_mask,
                           float* __restrict__ ggrad, float* __restrict__ dL_dcolors)
{
    using Cfg = BmCfg<NQ>;
    constexpr int NQE = Cfg::NQE, ROW = Cfg::ROW, MT = Cfg::MT;
    SAGARS_DYNAMIC_SMEM(smem_raw);
    BmSmem<NQ>& sm = *reinterpret_cast<BmSmem<NQ>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = gridDim.x;
    const uint32_t tile_x0 = blockIdx.x * TILE_X, tile_y0 = blockIdx.y * TILE_Y;
    const int xl = (warp & 1) * 8 + (lane & 7), yl = (warp >> 1) * 4 + (lane >> 3);   // this thread's pixel in the tile
    const int rl = yl * TILE_X + xl;                                                   // raster-local index
    const uint32_t px = tile_x0 + xl, py = tile_y0 + yl;
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    SAGARS_PIN_F2(pixx, pixy);   // keep nvcc from rematerialising them in the hot loop
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    if (total <= 0) return;   // empty tile: nothing to differentiate (before the gradient rows are fetched)

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;

    // ---- one-time setup -------------------------------------------------------------------------------------------
    // Two independent latency chains start here and overlap: (a) the upstream gradient row of this pixel (C strided
    // loads from the planar image, held in registers until (b) is under way), (b) n_contrib -> deepest list position
    // of the tile -> Gaussian ids of the first batch -> their records / feature rows.
    float gr[ROW];
    {
        float gmask = 0.f;
        if (MD) gmask = inside ? dL_dout_mask[pix_id] : 0.f;
#pragma unroll
        for (int k = 0; k < ROW; k++) {
            float x = 0.f;
            if (COLOR && k < K) x = inside ? dL_dpix[(size_t)k * plane + pix_id] : 0.f;
            if (MD && k == K) x = gmask;   // the mask gradient rides as channel K of the colour product
            gr[k] = x;
        }
    }
    int warp_n = my_n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_n = max(warp_n, __shfl_xor_sync(0xffffffffu, warp_n, o));
    if (lane == 0) sm.warp_max[warp] = (uint32_t)warp_n;
    {   // rows past the fill level are multiplied too (their products are never used): start from finite values
        float* w = &sm.rowW[warp][0][0];
        float* q = &sm.rowQ[warp][0][0];
#pragma unroll
        for (int i = 0; i < BM_N; i++) { w[i * 32 + lane] = 0.f; q[i * 32 + lane] = 0.f; }
        if (lane < BM_N) sm.row_id[warp][lane] = 0;
    }
    if (!VEC || (K >> 2) < NQ) {   // zero the padded feature channels once
        float* f = reinterpret_cast<float*>(&sm.feat[0][0][0]);
        for (int c = tid; c < 2 * BM_NB * 4 * NQ; c += TILE_PIX) f[c] = 0.f;
    }
    __syncthreads();
    int maxc = 0;
#pragma unroll
    for (int w8 = 0; w8 < 8; w8++) maxc = max(maxc, (int)sm.warp_max[w8]);
    maxc = min(maxc, total);
    if (maxc <= 0) return;
    const int nbatch = (maxc + BM_NB - 1) / BM_NB;
    // batch b covers list positions pos_hi(b) - jj, jj = 0 .. cnt(b)-1, with pos_hi(b) = maxc - 1 - b*NB
    auto batch_cnt = [&](int b) { return min(BM_NB, maxc - b * BM_NB); };
    auto load_id = [&](int b, int jj) { return point_list[range.x + (maxc - 1 - b * BM_NB - jj)]; };

    // ids(0), ids(1); records + features of batch 0
    if (tid < batch_cnt(0)) sm.ids[0][tid] = load_id(0, tid);
    __syncthreads();
    bm_issue_batch<NQ>(sm, 0, 0, batch_cnt(0), K, VEC, COLOR, geo, features);
    cp_async_commit();
    if (nbatch > 1 && tid < batch_cnt(1)) sm.ids[1][tid] = load_id(1, tid);

    // the gradient row -> swizzled smem row (quad q at physical quad (q + rl) % NQE)
    float bgdot = 0.f;
#pragma unroll
    for (int q = 0; q < NQE; q++) {
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (COLOR && 4 * q + c < K) bgdot += bg[4 * q + c] * gr[4 * q + c];
        *reinterpret_cast<float4*>(&sm.Gs[rl][4 * ((q + rl) & (NQE - 1))]) = make_float4(gr[4 * q], gr[4 * q + 1], gr[4 * q + 2], gr[4 * q + 3]);
    }
    cp_async_wait_all();
    bm_pad_geo<NQ>(sm, 0, batch_cnt(0));
    __syncthreads();

    float T = T_final;
    float acc_r = 0.f, last_alpha = 0.f, last_s = 0.f;

    float* const rowW = &sm.rowW[warp][0][0];
    float* const rowQ = &sm.rowQ[warp][0][0];

    // mma fragment coordinates of this lane
    const int fg = lane >> 2, ft = lane & 3;
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;   // (0.5 * W) rounded to float, as the reference
    const float tcx = (float)tile_x0 + 7.5f, tcy = (float)tile_y0 + 7.5f;
    // the warp's 8x4 pixel block (pixel centres), for the block-level candidate test
    const float bx0 = (float)(tile_x0 + (warp & 1) * 8), bx1 = bx0 + 7.f;
    const float by0 = (float)(tile_y0 + (warp >> 1) * 4), by1 = by0 + 3.f;

    // multiply the waiting rows [0, nrows) (nrows <= 8) with the warp's gradient rows / the moment basis and send the
    // results to global memory
    auto flush_rows = [&](int nrows) {
        __syncwarp();
        float dc[MT][4];
#pragma unroll
        for (int m = 0; m < MT; m++) dc[m][0] = dc[m][1] = dc[m][2] = dc[m][3] = 0.f;
        float dm[4] = {0.f, 0.f, 0.f, 0.f};
        // this lane's row of the tiles: column n = fg of W^T, row fg of Q; pixel p sits at column (p + 4 fg) & 31
        const float* Wr = rowW + fg * 32;
        const float* Qr = rowQ + fg * 32;
        // moment basis X[p][m] of fragment column m = fg at pixel (x, y): value = xa + (xb + xc * y) * y
        const float x0 = (float)((warp & 1) * 8 + ft) - 7.5f, x1 = x0 + 4.f;
        const float xa0 = (fg == 0) ? 1.f : (fg == 1) ? x0 : (fg == 3) ? x0 * x0 : 0.f;
        const float xa1 = (fg == 0) ? 1.f : (fg == 1) ? x1 : (fg == 3) ? x1 * x1 : 0.f;
        const float xb0 = (fg == 2) ? 1.f : (fg == 4) ? x0 : 0.f;
        const float xb1 = (fg == 2) ? 1.f : (fg == 4) ? x1 : 0.f;
        const float xc = (fg == 5) ? 1.f : 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int c0 = (ks * 8 + ft + 4 * fg) & 31, c1 = (c0 + 4) & 31;
            uint32_t wh0, wl0, wh1, wl1, qh0, ql0, qh1, ql1;
            split_tf32(Wr[c0], wh0, wl0);
            split_tf32(Wr[c1], wh1, wl1);
            split_tf32(Qr[c0], qh0, ql0);
            split_tf32(Qr[c1], qh1, ql1);
            // rows of the gradient tile: warp-local pixels ks*8 + ft and + 4  (same tile row, x and x+4)
            const int r0 = ((warp >> 1) * 4 + ks) * TILE_X + (warp & 1) * 8 + ft;
            const int r1 = r0 + 4;
            if (COLOR || MD) {
#pragma unroll
                for (int m = 0; m < MT; m++) {
                    const int cl = 16 * m + fg, chh = cl + 8;     // channels of fragment rows g and g + 8
                    uint32_t ah[4], al[4];
                    split_tf32(sm.Gs[r0][4 * (((cl >> 2) + r0) & (NQE - 1)) + (cl & 3)], ah[0], al[0]);
                    split_tf32(sm.Gs[r1][4 * (((cl >> 2) + r1) & (NQE - 1)) + (cl & 3)], ah[2], al[2]);
                    if (16 * m + 8 < ROW) {
                        split_tf32(sm.Gs[r0][4 * (((chh >> 2) + r0) & (NQE - 1)) + (chh & 3)], ah[1], al[1]);
                        split_tf32(sm.Gs[r1][4 * (((chh >> 2) + r1) & (NQE - 1)) + (chh & 3)], ah[3], al[3]);
                    } else {
                        ah[1] = al[1] = ah[3] = al[3] = 0u;
                    }
                    mma_16n8k8(dc[m], al[0], al[1], al[2], al[3], wh0, wh1);
                    mma_16n8k8(dc[m], ah[0], ah[1], ah[2], ah[3], wl0, wl1);
                    mma_16n8k8(dc[m], ah[0], ah[1], ah[2], ah[3], wh0, wh1);
                }
            }
            // B fragments of the moment basis, p = warp-local pixel ks*8 + ft (+4); exact in tf32
            const float yb = (float)((warp >> 1) * 4 + ks) - 7.5f;
            const float v0 = xa0 + (xb0 + xc * yb) * yb;
            const float v1 = xa1 + (xb1 + xc * yb) * yb;
            mma_16n8k8(dm, ql0, 0u, ql1, 0u, __float_as_uint(v0), __float_as_uint(v1));
            mma_16n8k8(dm, qh0, 0u, qh1, 0u, __float_as_uint(v0), __float_as_uint(v1));
        }

        // colour product: this lane holds channels (16 m + fg, + 8) of rows 2 ft and 2 ft + 1
        if (COLOR || MD) {
            const uint32_t ida = sm.row_id[warp][2 * ft], idb2 = sm.row_id[warp][2 * ft + 1];
            const bool va = 2 * ft < nrows, vb = 2 * ft + 1 < nrows;
            auto emit = [&](uint32_t id, int ch, float v) {
                if (COLOR && ch < K) red_add(dL_dcolors + (size_t)id * K + ch, v);
                else if (MD && ch == K) red_add(ggrad + (size_t)id * GG_STRIDE + 6, v);
            };
#pragma unroll
            for (int m = 0; m < MT; m++) {
                const int cl = 16 * m + fg;
                if (va) { emit(ida, cl, dc[m][0]); emit(ida, cl + 8, dc[m][2]); }
                if (vb) { emit(idb2, cl, dc[m][1]); emit(idb2, cl + 8, dc[m][3]); }
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
            if (fg < nrows && ft < 3) {
                const uint32_t id = sm.row_id[warp][fg];
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id));       // L1/L2 hit: staged a moment ago
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id + 4));
                const float conx = g0.z, cony = g0.w, conz = g1.x, o = g1.y;
                // sums over the pixels of q * (1, dx, dy, dx^2, dx dy, dy^2) with d = centre - pixel = c - x'
                const float cx = g0.x - tcx, cy = g0.y - tcy;
                const float Sx = cx * m0 - mx;
                const float Sy = cy * m0 - my;
                float ua, ub;
                int sa, sb;
                if (ft == 0) {
                    ua = m0; sa = 5;                                                  // dL/dopacity
                    ub = -o * half_W * (conx * Sx + cony * Sy); sb = 0;               // dL/dmean2D.x
                } else if (ft == 1) {
                    const float Sxx = cx * cx * m0 - 2.f * cx * mx + mxx;
                    ua = -o * half_H * (conz * Sy + cony * Sx); sa = 1;               // dL/dmean2D.y
                    ub = -0.5f * o * Sxx; sb = 2;                                     // dL/dconic.x
                } else {
                    const float Sxy = cx * cy * m0 - cx * my - cy * mx + mxy;
                    const float Syy = cy * cy * m0 - 2.f * cy * my + myy;
                    ua = -0.5f * o * Sxy; sa = 3;                                     // dL/dconic.y
                    ub = -0.5f * o * Syy; sb = 4;                                     // dL/dconic.w
                }
                red_add(ggrad + (size_t)id * GG_STRIDE + sa, ua);
                red_add(ggrad + (size_t)id * GG_STRIDE + sb, ub);
            }
        }
        __syncwarp();   // every lane is done with the rows before they are overwritten
    };

    for (int b = 0; b < nbatch; b++) {
        const int stage = b & 1, idb = b % 3;
        const int cnt = batch_cnt(b);
        const int pos_hi = maxc - 1 - b * BM_NB;

        // data of batch b+1 starts flying; ids of batch b+2 into a register
        if (b + 1 < nbatch) {
            bm_issue_batch<NQ>(sm, stage ^ 1, (b + 1) % 3, batch_cnt(b + 1), K, VEC, COLOR, geo, features);
            cp_async_commit();
        }
        uint32_t next_id = 0;
        const bool have_next_id = (b + 2 < nbatch) && tid < batch_cnt(b + 2);
        if (have_next_id) next_id = load_id(b + 2, tid);

        // ---------------- phase A ----------------
        if (pos_hi - (cnt - 1) < warp_n) {   // some pixel of this warp still has contributors in this batch
            // block-level candidate test (candidate.cuh), lane = splat: can ANY point of the warp's 8x4 pixel block reach
            // the splat's accept threshold?  Survivors are compacted, in list order, into the warp's candidate list.
            int ncand;
            {
                const uint32_t lt = (1u << lane) - 1u;
                const bool k0 = !block_rejects(sm.geo[stage][lane][0], sm.geo[stage][lane][1], bx0, bx1, by0, by1) &&
                                (pos_hi - lane < warp_n);
                const bool k1 = !block_rejects(sm.geo[stage][32 + lane][0], sm.geo[stage][32 + lane][1], bx0, bx1, by0, by1) &&
                                (pos_hi - (32 + lane) < warp_n);
                const uint32_t c0 = __ballot_sync(0xffffffffu, k0), c1 = __ballot_sync(0xffffffffu, k1);
                const int n0 = __popc(c0);
                if (k0) sm.clist[warp][__popc(c0 & lt)] = (uint8_t)lane;
                if (k1) sm.clist[warp][n0 + __popc(c1 & lt)] = (uint8_t)(32 + lane);
                ncand = n0 + __popc(c1);
                __syncwarp();
            }
#pragma unroll 1
            for (int k0 = 0; k0 < ncand; k0 += BM_N) {
                const int m = min(BM_N, ncand - k0);   // candidates of this group = rows of the operand tiles
                // ---- (1) s[p][i] = f_i . g_p for the group's candidates as a tensor-core product (3xTF32):
                //          S (32 pixels x 8) = G (32 x C) * F^T (C x 8).  The feature rows are first gathered into the W
                //          tile's storage (row i, channel c at (c + 4 i) & 31), S lands in the Q tile's storage with the
                //          Q layout, so the lane that later writes Q[i][p] is the one that reads S[i][p].
                if (COLOR) {
                    float sacc[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) sacc[mt][0] = sacc[mt][1] = sacc[mt][2] = sacc[mt][3] = 0.f;
                    constexpr int QR = (ROW < 32 ? ROW : 32) / 4;      // feature quads per row and channel block
#pragma unroll 1
                    for (int cb = 0; cb < ROW; cb += 32) {
                        if (cb > 0) __syncwarp();
                        for (int idx = lane; idx < BM_N * QR; idx += 32) {
                            const int r = idx / QR, qd = idx - r * QR;
                            const int jr = sm.clist[warp][min(k0 + r, ncand - 1)];
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((cb >> 2) + qd < NQ) v = sm.feat[stage][jr][(cb >> 2) + qd];
                            *reinterpret_cast<float4*>(rowW + r * 32 + 4 * ((qd + r) & 7)) = v;
                        }
                        __syncwarp();
                        const float* Fr = rowW + fg * 32;
#pragma unroll
                        for (int ks = 0; ks < QR / 2; ks++) {
                            uint32_t bh0, bl0, bh1, bl1;
                            split_tf32(Fr[(ks * 8 + ft + 4 * fg) & 31], bh0, bl0);
                            split_tf32(Fr[(ks * 8 + ft + 4 + 4 * fg) & 31], bh1, bl1);
                            const int ch0 = cb + ks * 8 + ft, ch1 = ch0 + 4;
#pragma unroll
                            for (int mt = 0; mt < 2; mt++) {
                                // pixels 16 mt + fg and + 8 of the warp's block: raster rows 2 mt and 2 mt + 1, x = fg
                                const int p0 = ((warp >> 1) * 4 + 2 * mt) * TILE_X + (warp & 1) * 8 + fg, p1 = p0 + TILE_X;
                                uint32_t ah[4], al[4];
                                split_tf32(sm.Gs[p0][4 * (((ch0 >> 2) + p0) & (NQE - 1)) + (ch0 & 3)], ah[0], al[0]);
                                split_tf32(sm.Gs[p1][4 * (((ch0 >> 2) + p1) & (NQE - 1)) + (ch0 & 3)], ah[1], al[1]);
                                split_tf32(sm.Gs[p0][4 * (((ch1 >> 2) + p0) & (NQE - 1)) + (ch1 & 3)], ah[2], al[2]);
                                split_tf32(sm.Gs[p1][4 * (((ch1 >> 2) + p1) & (NQE - 1)) + (ch1 & 3)], ah[3], al[3]);
                                mma_16n8k8(sacc[mt], al[0], al[1], al[2], al[3], bh0, bh1);
                                mma_16n8k8(sacc[mt], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
                                mma_16n8k8(sacc[mt], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
                            }
                        }
                    }
                    // fragment (pixel 16 mt + fg [+8], candidates 2 ft, 2 ft + 1) -> S[i][(p + 4 i) & 31]
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        const int pa = 16 * mt + fg, pb = pa + 8;
                        rowQ[(2 * ft) * 32 + ((pa + 8 * ft) & 31)] = sacc[mt][0];
                        rowQ[(2 * ft + 1) * 32 + ((pa + 8 * ft + 4) & 31)] = sacc[mt][1];
                        rowQ[(2 * ft) * 32 + ((pb + 8 * ft) & 31)] = sacc[mt][2];
                        rowQ[(2 * ft + 1) * 32 + ((pb + 8 * ft + 4) & 31)] = sacc[mt][3];
                    }
                    __syncwarp();
                }
                // ---- (2) thread = pixel over the group's candidates (the reference's traversal): row i of W / Q ----
#pragma unroll 1
                for (int i = 0; i < m; i++) {
                    const int jj = sm.clist[warp][k0 + i];
                    const float4 g0 = sm.geo[stage][jj][0];
                    const float4 g1 = sm.geo[stage][jj][1];
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    const float pw = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    const bool cd = (pos_hi - jj < my_n) && !(pw > 0.0f) && (pw >= g1.z);
                    const int col = (lane + 4 * i) & 31;
                    float w = 0.f, q = 0.f;
                    if (cd) {
                        const float G = expf(pw);
                        const float alpha = fminf(0.99f, g1.y * G);
                        if (!(alpha < 1.0f / 255.0f)) {
                            T = T / (1.f - alpha);
                            const float s = COLOR ? rowQ[i * 32 + col] : 0.f;
                            acc_r = last_alpha * last_s + (1.f - last_alpha) * acc_r;
                            last_s = s;
                            float dL_dalpha = (s - acc_r) * T;
                            last_alpha = alpha;
                            if (bgdot != 0.f) dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;   // zero background: the term is exactly 0
                            w = alpha * T;
                            q = G * dL_dalpha;
                        }
                    }
                    rowW[i * 32 + col] = w;   // all lanes write: zero where the pixel did not blend
                    rowQ[i * 32 + col] = q;
                    if (lane == 0) sm.row_id[warp][i] = sm.ids[idb][jj];
                }
                // ---- (3) the rows' products leave for global memory ----
                flush_rows(m);
            }
        }

        // publish ids(b+2); wait for the copies of batch b+1
        if (have_next_id) sm.ids[(b + 2) % 3][tid] = next_id;
        cp_async_wait_all();
        if (b + 1 < nbatch) bm_pad_geo<NQ>(sm, stage ^ 1, batch_cnt(b + 1));
        __syncthreads();   // batch b+1 is visible; nobody reads the buffers of batch b any more
    }
}

}  // namespace sagars
