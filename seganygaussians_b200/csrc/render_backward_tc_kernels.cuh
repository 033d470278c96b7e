// render_backward_tc_kernels.cuh -- the device code of render_backward_tc.cu (see there); free of host-side runtime calls so that
// the CPU suite can run it under tests/cuda_emu/ (tcgen05 / TMEM restated in tests/cuda_emu/tc_emu.h).
#pragma once
#include "common.cuh"
#include "cp_async.cuh"
#include "tc.cuh"
#include "candidate.cuh"

#ifndef SAGARS_DYNAMIC_SMEM_1024
#if defined(SAGARS_CUDA_EMU)
#define SAGARS_DYNAMIC_SMEM_1024(name) SAGARS_DYNAMIC_SMEM(name)
#else
#define SAGARS_DYNAMIC_SMEM_1024(name) extern __shared__ __align__(1024) unsigned char name[]
#endif
#endif

namespace sagars {

constexpr int BT_C = 32;            // colour channels (this kernel: C = 32 only)
constexpr int BT_PIX = 128;         // pixels per CTA: half a tile, 16 wide x 8 high = four 8x4 blocks, one warp each
constexpr int BT_RING = 64;         // candidate ring (up to NB - 1 carried over + 32 new)

// Shared-memory operand tiles (all SWIZZLE_NONE canonical K-major layouts, tc.cuh; offsets in bytes).  Measured on B200
// (tools/probes/tcgen05_bwd_probe.cu, profiles/r2_tcgen05_bwd_probe.md): kind::tf32 with an MN-major operand (either one, dense
// or not) leaves an all-zero accumulator, so the gradient tile is kept twice, once per contraction:
//
//  G1  [pixel group p / 8][chunk 0..15][p % 8][4 floats]: chunk c < 8: tf32-exact high parts of channels 4c..4c+3 of the pixel's
//      upstream gradient, c >= 8: the remainders.  A operand of S = G F^T (rows = pixels, k = channels).
//  G3  [k-chunk = p / 4][row / 8][row % 8][4 floats]: rows 0..31 high parts of the gradient channels, 32..63 remainders, 64..69 the
//      moment basis (1, x, y, x^2, x y, y^2), 70..71 zero.  A operand of the gradient product (k = pixels).  The k-chunk step is
//      9 row groups + 16 B (bank spread of the one-time transposing stores); an M = 128 instruction reads 16 row groups from each
//      k-chunk: groups 9..15 alias the next chunk (or the slack behind the tile); those accumulator rows are never read.
//  B3  [k-chunk = p / 4][n / 8][n % 8][4 floats], n = column: [0,NB) w_hi, [NB,2NB) w_lo, [2NB,3NB) q_hi, [3NB,4NB) q_lo of the
//      batch's candidates; k-chunk step padded by 16 B so that a warp's 32 scalar stores hit 32 different banks.
//  F   [k-chunk = channel / 4][n / 8][n % 8][4 floats] feature rows of the batch (hi, lo).
template <int NB>
struct BtCfg {
    static constexpr int N3 = 4 * NB;                       // columns of the gradient product
    static constexpr int G1_PG = 16 * 128;                  // bytes per group of 8 pixels (= 8-row-group step of G1)
    static constexpr int G1_BYTES = (BT_PIX / 8) * G1_PG;
    static constexpr int G3_LBO = 9 * 128 + 16;             // bytes between k-chunks of G3
    static constexpr int G3_BYTES = (BT_PIX / 4) * G3_LBO + 7 * 128;
    static constexpr int B3_LBO = (N3 / 8) * 128 + 16;      // bytes between k-chunks
    static constexpr int B3_BYTES = (BT_PIX / 4) * B3_LBO;
    static constexpr int F_LBO = 2 * 128;                   // 16 rows
    static constexpr int F_BYTES = (BT_C / 4) * F_LBO;
    static constexpr int TMEM_COLS = (64 + N3 <= 128) ? 128 : 256;
    static constexpr int D3_COL = 64;                       // S at columns [0, 16), the gradient product at [64, 64 + N3)
};

template <int NB>
struct BtSmem {
    unsigned char G1[BtCfg<NB>::G1_BYTES];
    unsigned char G3[BtCfg<NB>::G3_BYTES];
    unsigned char B3[BtCfg<NB>::B3_BYTES];
    unsigned char Fh[BtCfg<NB>::F_BYTES];
    unsigned char Fl[BtCfg<NB>::F_BYTES];
    float4 ctab[BT_RING][2];        // candidate ring: record (x, y, cx, cy | cz, opacity, accept_threshold, list position)
    uint32_t cid[BT_RING];          // Gaussian id
    uint32_t cmem[BT_RING];         // bit w: candidate of warp w's pixel block
    float4 brec[2][16][2];          // the batch's records and ids, kept for its epilogue (which runs one batch later)
    uint32_t bid[2][16];
    float mom[8][16];               // moment rows of the batch (warp 2)
    uint32_t wmask[2][4];           // per-warp candidate masks of a 32-splat chunk
    int32_t red_n[4];
    uint64_t mbar[2];               // [0]: S = G F^T done, [1]: gradient product done
    uint32_t tmem_base;
};

// C = 32, colour only (no mask / depth channel).  Grid = (tiles_x, 2 * tiles_y), 128 threads.
template <int NB>
__global__ void __launch_bounds__(BT_PIX, 2)
render_backward_tc_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                          const float* __restrict__ bg, const float* __restrict__ geo, const float* __restrict__ features,
                          const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                          const float* __restrict__ dL_dpix, float* __restrict__ ggrad, float* __restrict__ dL_dcolors)
{
    using Cfg = BtCfg<NB>;
    static_assert(NB == 16, "the batch size is tied to the 16-column S tile and the F gather mapping");
    SAGARS_DYNAMIC_SMEM_1024(smem_raw);
    BtSmem<NB>& sm = *reinterpret_cast<BtSmem<NB>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile_x0 = blockIdx.x * TILE_X, half_y0 = blockIdx.y * 8;
    const uint32_t blk_x0 = tile_x0 + (warp & 1) * 8, blk_y0 = half_y0 + (warp >> 1) * 4;
    const uint32_t px = blk_x0 + (lane & 7), py = blk_y0 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    SAGARS_PIN_F2(pixx, pixy);
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[(blockIdx.y >> 1) * gridDim.x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    if (total <= 0) return;   // empty tile (uniform over the CTA)

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;
    int wn = my_n;            // deepest contributor of this warp's block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wn = max(wn, __shfl_xor_sync(0xffffffffu, wn, o));
    if (lane == 0) sm.red_n[warp] = wn;
    __syncthreads();
    const int maxc = min(max(max(sm.red_n[0], sm.red_n[1]), max(sm.red_n[2], sm.red_n[3])), total);
    if (maxc <= 0) return;    // uniform
    const int nchunk = (maxc + 31) >> 5;

    // chunk c, lane l <-> list position maxc - 1 - 32 c - l (back to front); every warp reads the same chunk
    auto chunk_pos = [&](int c) { return maxc - 1 - 32 * c - lane; };
    int pos_cur = chunk_pos(0);
    uint32_t id_cur = pos_cur >= 0 ? point_list[range.x + pos_cur] : 0u;

    // ---- one-time: the pixel's gradient row -> G tile (hi, lo, basis); mbarriers; TMEM ----
    float bgdot = 0.f;
    {
        float* g1 = reinterpret_cast<float*>(sm.G1) + (tid >> 3) * (Cfg::G1_PG / 4) + (tid & 7) * 4;
        float* g3 = reinterpret_cast<float*>(sm.G3) + (tid >> 2) * (Cfg::G3_LBO / 4) + (tid & 3);       // row r at (r / 8) * 32 + (r % 8) * 4
#pragma unroll
        for (int c = 0; c < BT_C / 4; c++) {
            float g[4], h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                g[k] = inside ? dL_dpix[(size_t)(4 * c + k) * plane + pix_id] : 0.f;
                bgdot += bg[4 * c + k] * g[k];
                h[k] = tc::tf32_hi(g[k]);
                l[k] = g[k] - h[k];
                const int r = 4 * c + k;
                g3[(r >> 3) * 32 + (r & 7) * 4] = h[k];
                g3[((32 + r) >> 3) * 32 + (r & 7) * 4] = l[k];
            }
            *reinterpret_cast<float4*>(g1 + c * 32) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(g1 + (8 + c) * 32) = make_float4(l[0], l[1], l[2], l[3]);
        }
        // pixel coordinates relative to the centre of the 16 x 8 pixel group: multiples of 0.5, squares exact in tf32
        const float xr = (float)((warp & 1) * 8 + (lane & 7)) - 7.5f, yr = (float)((warp >> 1) * 4 + (lane >> 3)) - 3.5f;
        const float basis[8] = {1.f, xr, yr, xr * xr, xr * yr, yr * yr, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; k++) g3[8 * 32 + k * 4] = basis[k];
        float* slack = reinterpret_cast<float*>(sm.G3) + (BT_PIX / 4) * (Cfg::G3_LBO / 4);
        for (int i = tid; i < 7 * 32; i += BT_PIX) slack[i] = 0.f;
        float* fz = reinterpret_cast<float*>(sm.Fh);
        for (int i = tid; i < 2 * Cfg::F_BYTES / 4; i += BT_PIX) fz[i] = 0.f;     // Fh and Fl are adjacent
    }
    float4 r0_cur = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_cur));
    float4 r1_cur = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_cur + 4));
    if (tid == 0) {
        tc::mbar_init(&sm.mbar[0], 1);
        tc::mbar_init(&sm.mbar[1], 1);
        tc::mbar_init_fence();
    }
    if (warp == 0) tc::tmem_alloc<Cfg::TMEM_COLS>(&sm.tmem_base);
    tc::fence_smem_to_async_proxy();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_s = sm.tmem_base, tmem_d3 = sm.tmem_base + Cfg::D3_COL;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

    constexpr uint32_t ID1 = tc::idesc_tf32(128, 16, 0, 0);          // S: A K-major, B K-major
    constexpr uint32_t ID3 = tc::idesc_tf32(128, Cfg::N3, 0, 0);     // gradient product: both K-major (k = pixels)
    const uint32_t g1_addr = smem_u32(sm.G1), g3_addr = smem_u32(sm.G3), b3_addr = smem_u32(sm.B3), fh_addr = smem_u32(sm.Fh), fl_addr = smem_u32(sm.Fl);

    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;
    const float bx0 = (float)blk_x0, bx1 = bx0 + 7.f, by0 = (float)blk_y0, by1 = by0 + 3.f;
    const float gcx = (float)tile_x0 + 7.5f, gcy = (float)half_y0 + 3.5f;     // centre of the pixel group
    const uint32_t lt = (1u << lane) - 1u;

    float T = T_final;
    float acc_r = 0.f, last_alpha = 0.f, last_s = 0.f;
    int head = 0, ntab = 0;              // candidate ring: slots [head, head + ntab)
    uint32_t nbatch = 0;                 // batches issued so far
    bool pending = false;                // a gradient product is in flight; its epilogue has not run
    int pend_m = 0;

    float* const b3w = reinterpret_cast<float*>(sm.B3) + (tid >> 2) * (Cfg::B3_LBO / 4) + (tid & 3);

    // gradient product of batch `par`: accumulator rows -> global memory
    auto epilogue = [&](int par, int m) {
        tc::mbar_wait(&sm.mbar[1], (uint32_t)par);
        tc::fence_after_sync();
        if (warp < 2) {
            // rows 0..31: high parts of the gradient channels, rows 32..63: remainders; columns j (w_hi) and NB + j (w_lo)
            float v[32];
            tc::tmem_ld32(tmem_d3 + lane_sel, v);
#pragma unroll
            for (int j = 0; j < NB; j++) {
                if (j < m) red_add(dL_dcolors + (size_t)sm.bid[par][j] * BT_C + lane, v[j] + v[NB + j]);
            }
        } else if (warp == 2) {
            float v[32];
            tc::tmem_ld32(tmem_d3 + lane_sel + 2 * NB, v);
            if (lane < 6) {
#pragma unroll
                for (int j = 0; j < NB; j++) sm.mom[lane][j] = v[j] + v[NB + j];
            }
            __syncwarp();
            if (lane < m) {
                const float m0 = sm.mom[0][lane], mx = sm.mom[1][lane], my = sm.mom[2][lane];
                const float mxx = sm.mom[3][lane], mxy = sm.mom[4][lane], myy = sm.mom[5][lane];
                const float4 g0 = sm.brec[par][lane][0];
                const float4 g1 = sm.brec[par][lane][1];
                const float conx = g0.z, cony = g0.w, conz = g1.x, o = g1.y;
                // sums over the pixels of q * (1, dx, dy, dx^2, dx dy, dy^2) with d = centre - pixel = c - x'
                const float cx = g0.x - gcx, cy = g0.y - gcy;
                const float Sx = cx * m0 - mx;
                const float Sy = cy * m0 - my;
                const float Sxx = cx * cx * m0 - 2.f * cx * mx + mxx;
                const float Sxy = cx * cy * m0 - cx * my - cy * mx + mxy;
                const float Syy = cy * cy * m0 - 2.f * cy * my + myy;
                float* gg = ggrad + (size_t)sm.bid[par][lane] * GG_STRIDE;
                red_add(gg + 0, -o * half_W * (conx * Sx + cony * Sy));      // dL/dmean2D.x
                red_add(gg + 1, -o * half_H * (conz * Sy + cony * Sx));      // dL/dmean2D.y
                red_add(gg + 2, -0.5f * o * Sxx);                            // dL/dconic.x
                red_add(gg + 3, -0.5f * o * Sxy);                            // dL/dconic.y
                red_add(gg + 4, -0.5f * o * Syy);                            // dL/dconic.w
                red_add(gg + 5, m0);                                         // dL/dopacity
            }
            __syncwarp();
        }
        tc::fence_before_sync();
    };

    // one batch: ring slots [head, head + m), m <= NB
    auto process_batch = [&](int m) {
        const int par = (int)(nbatch & 1u);
        // ---- (a) feature rows of the batch -> F tiles (hi, lo); the batch's ids / records for its epilogue ----
        {
            const int j = tid & 15, quad = tid >> 4;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < m) f = __ldg(reinterpret_cast<const float4*>(features + (size_t)sm.cid[(head + j) & (BT_RING - 1)] * BT_C + 4 * quad));
            float4 h, l;
            h.x = tc::tf32_hi(f.x); l.x = f.x - h.x;
            h.y = tc::tf32_hi(f.y); l.y = f.y - h.y;
            h.z = tc::tf32_hi(f.z); l.z = f.z - h.z;
            h.w = tc::tf32_hi(f.w); l.w = f.w - h.w;
            const int off = quad * (Cfg::F_LBO / 4) + (j >> 3) * 32 + (j & 7) * 4;
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(sm.Fh) + off) = h;
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(sm.Fl) + off) = l;
            if (tid < m) {
                const int slot = (head + tid) & (BT_RING - 1);
                sm.bid[par][tid] = sm.cid[slot];
                sm.brec[par][tid][0] = sm.ctab[slot][0];
                sm.brec[par][tid][1] = sm.ctab[slot][1];
            }
        }
        // ---- (b) S (128 pixels x 16) = G F^T, 3xTF32: lo*hi, hi*lo, hi*hi ----
        tc::fence_smem_to_async_proxy();
        tc::fence_before_sync();
        tc::bar_sync_128(1);
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int term = 0; term < 3; term++) {
                const uint32_t a0 = g1_addr + (term == 0 ? 1024u : 0u);
                const uint32_t f0 = (term == 1) ? fl_addr : fh_addr;
#pragma unroll
                for (int ks = 0; ks < BT_C / 8; ks++)
                    tc::mma_tf32(tmem_s, tc::smem_desc(a0 + ks * 256, 128, Cfg::G1_PG), tc::smem_desc(f0 + ks * 2 * Cfg::F_LBO, Cfg::F_LBO, 128),
                                 ID1, (term > 0 || ks > 0) ? 1u : 0u);
            }
            tc::commit(&sm.mbar[0]);
        }
        // ---- (c) the previous batch's gradient product leaves while the tensor core works on S ----
        if (pending) epilogue(par ^ 1, pend_m);
        // ---- (d) S -> registers: thread = pixel = accumulator row ----
        float s[16];
        tc::mbar_wait(&sm.mbar[0], (uint32_t)par);
        tc::fence_after_sync();
        tc::tmem_ld16(tmem_s + lane_sel, s);
        // ---- (e) thread = pixel over the batch (the reference's traversal); w = alpha T and q = G dL/dalpha -> B3 ----
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const int slot = (head + j) & (BT_RING - 1);
            float w = 0.f, q = 0.f;
            if (j < m && ((sm.cmem[slot] >> warp) & 1u)) {          // warp-uniform
                const float4 g0 = sm.ctab[slot][0];
                const float4 g1 = sm.ctab[slot][1];
                const float dx = g0.x - pixx, dy = g0.y - pixy;
                const float pw = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const bool cd = (__float_as_int(g1.w) < my_n) && !(pw > 0.0f) && (pw >= g1.z);
                if (cd) {
                    const float G = expf(pw);
                    const float alpha = fminf(0.99f, g1.y * G);
                    if (!(alpha < 1.0f / 255.0f)) {
                        T = T / (1.f - alpha);
                        acc_r = last_alpha * last_s + (1.f - last_alpha) * acc_r;
                        last_s = s[j];
                        float dL_dalpha = (s[j] - acc_r) * T;
                        last_alpha = alpha;
                        if (bgdot != 0.f) dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
                        w = alpha * T;
                        q = G * dL_dalpha;
                    }
                }
            }
            const float wh = tc::tf32_hi(w), qh = tc::tf32_hi(q);
            float* col = b3w + (j >> 3) * 32 + (j & 7) * 4;
            col[0 * (NB / 8) * 32] = wh;
            col[1 * (NB / 8) * 32] = w - wh;
            col[2 * (NB / 8) * 32] = qh;
            col[3 * (NB / 8) * 32] = q - qh;
        }
        // ---- (f) rows (G_hi | G_lo | basis) x columns (w_hi | w_lo | q_hi | q_lo), contraction over the 128 pixels ----
        tc::fence_smem_to_async_proxy();
        tc::fence_before_sync();
        tc::bar_sync_128(1);
        if (tid == 0) {
            tc::fence_after_sync();
#pragma unroll
            for (int ks = 0; ks < BT_PIX / 8; ks++)
                tc::mma_tf32(tmem_d3, tc::smem_desc(g3_addr + ks * 2 * Cfg::G3_LBO, Cfg::G3_LBO, 128),
                             tc::smem_desc(b3_addr + ks * 2 * Cfg::B3_LBO, Cfg::B3_LBO, 128), ID3, ks > 0 ? 1u : 0u);
            tc::commit(&sm.mbar[1]);
        }
        pending = true;
        pend_m = m;
        nbatch++;
        head = (head + m) & (BT_RING - 1);
        ntab -= m;
    };

    for (int c = 0; c < nchunk; c++) {
        const int pos_nxt = (c + 1 < nchunk) ? chunk_pos(c + 1) : -1;
        const uint32_t id_nxt = pos_nxt >= 0 ? point_list[range.x + pos_nxt] : 0u;

        // block-level candidate test (candidate.cuh), lane = splat, against this warp's 8x4 block
        const bool own = pos_cur >= 0 && pos_cur < wn && !block_rejects(r0_cur, r1_cur, bx0, bx1, by0, by1);
        const uint32_t own_mask = __ballot_sync(0xffffffffu, own);
        if (lane == 0) sm.wmask[c & 1][warp] = own_mask;
        float4 r0_nxt = make_float4(0.f, 0.f, 0.f, 0.f), r1_nxt = r0_nxt;
        if (pos_nxt >= 0) {
            r0_nxt = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_nxt));
            r1_nxt = __ldg(reinterpret_cast<const float4*>(geo + 8 * (size_t)id_nxt + 4));
        }
        tc::bar_sync_128(1);
        const uint32_t m0 = sm.wmask[c & 1][0], m1 = sm.wmask[c & 1][1], m2 = sm.wmask[c & 1][2], m3 = sm.wmask[c & 1][3];
        const uint32_t any = m0 | m1 | m2 | m3;
        // the group's candidates join the ring in list order; warp (rank & 3) writes entry `rank` (every warp holds the chunk)
        if ((any >> lane) & 1u) {
            const int rank = __popc(any & lt);
            if ((rank & 3) == warp) {
                const int slot = (head + ntab + rank) & (BT_RING - 1);
                float4 r1p = r1_cur;
                r1p.w = __int_as_float(pos_cur);
                sm.ctab[slot][0] = r0_cur;
                sm.ctab[slot][1] = r1p;
                sm.cid[slot] = id_cur;
                sm.cmem[slot] = ((m0 >> lane) & 1u) | (((m1 >> lane) & 1u) << 1) | (((m2 >> lane) & 1u) << 2) | (((m3 >> lane) & 1u) << 3);
            }
        }
        ntab += __popc(any);
        tc::bar_sync_128(1);

        const bool last = (c + 1 == nchunk);
        while (ntab >= NB || (last && ntab > 0)) process_batch(min(NB, ntab));

        pos_cur = pos_nxt;
        id_cur = id_nxt;
        r0_cur = r0_nxt;
        r1_cur = r1_nxt;
    }
    if (pending) epilogue((int)((nbatch - 1u) & 1u), pend_m);

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<Cfg::TMEM_COLS>(sm.tmem_base);
}

}  // namespace sagars
