// render_backward_mma.cu -- per-tile back-to-front gradient of the alpha compositing; the per-Gaussian reductions over
// the tile's pixels are warp-level tensor-core GEMMs.
//
// Semantics: CF cuda_rasterizer/backward.cu:399-559 (DEPTH backward.cu:400-564 adds dL_dmask), SURVEY.md Appendix
// A.13-A.17 / D.  One CTA per 16x16 tile, one thread per pixel, splats taken back to front in batches of 16.
//
//   phase A (thread = pixel, the reference's traversal): four `power` tests at a time, one warp vote rejects splats no
//     pixel of the warp can accept (conservative lower bound on power, math.cuh accept_threshold); accepted pairs
//     recompute alpha, undo T and need ONE dot product s = f_j . g_p because the reference's per-channel recurrence
//     accum_rec[ch] collapses to a scalar recurrence on a = accum_rec . g_p.  Each pair leaves two scalars in the
//     warp's private operand tiles: W[j][p] = alpha*T (weight of dL/dcolour) and Q[j][p] = G * dL/dalpha (weight of
//     every geometric gradient); everything else stays 0.
//   phase B (per warp, no CTA barrier): the warp multiplies its own 16 x 32 tiles against the 32 gradient rows of
//     its pixels and against the pixel-coordinate basis,
//         dL/dcolour[j][:] += W[j][p] * g[p][:]                       (16 x 32 x C)
//         moments[j][:]    += Q[j][p] * (1, x, y, x^2, xy, y^2)(p)    (16 x 32 x 8, tile-centred coordinates)
//     with mma.sync.m16n8k8 TF32 in 3xTF32 split precision (~2^-21), accumulators in registers, and adds the result
//     result in its own (now dead) operand slab.
//   output (after the batch's CTA barrier): 16 threads per splat sum the warps' partial rows and turn them into ONE
//     `red.global.add.v4.f32` per channel quad and six scalar reds (dL/dopacity, dL/dmean2D, dL/dconic in closed form
//     from the six moments) -- instead of the reference's (C+6) atomics per blended (pixel, splat) pair.
//
// Shared-memory rows are XOR/rotation swizzled instead of padded (g rows by quad, W/Q rows by 4 columns per row), which
// makes both the thread-per-pixel float4 reads and the mma fragment gathers conflict-free and keeps 3 CTAs per SM.
#include "common.cuh"
#include "cp_async.cuh"

namespace sagars {

constexpr int BM_NB = 16;   // instances per batch = M of the mma tiles

template <int NQ>
struct BmCfg {
    static constexpr int NQE = NQ < 2 ? 2 : NQ;     // quads per gradient row (>= 8 channels for one n-tile)
    static constexpr int ROW = 4 * NQE;             // floats per gradient row
    static constexpr int NT = NQE / 2;              // 8-channel n-tiles
    static constexpr int ACC_N = 8 * NT + 8;        // colour columns + 8 moment columns
    static constexpr int SLAB = (BM_NB * ACC_N > 2 * BM_NB * 32) ? BM_NB * ACC_N : 2 * BM_NB * 32;   // floats per warp slab
};

template <int NQ>
struct BmSmem {
    float Gs[TILE_PIX][BmCfg<NQ>::ROW];             // gradient rows by raster-local pixel; quad q of row r lives at quad (q + r) % NQE
    // per-warp slab: operand tiles W[j][(lane + 4 j) & 31], Q[...] during phase A / the mma; after the mma the same 4 KB
    // hold the warp's partial result P[16][ACC_N] until the CTA has summed the partials of all warps
    float WQ[8][BmCfg<NQ>::SLAB];                   // W tile at [0, 512), Q tile at [512, 1024)
    uint32_t touched[BM_NB];                        // splat had a blended pixel in this tile (this batch)
    uint32_t contrib;                               // bit w: warp w left a partial result for this batch
    float4 geo[2][BM_NB][2];                        // x, y, cx, cy | cz, opacity, accept_threshold, -
    float4 feat[2][BM_NB][NQ];                      // feature rows, zero padded
    uint32_t ids[3][BM_NB];
    uint32_t max_contrib;
};

__device__ __forceinline__ uint32_t f2tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo)
{
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
// D(16x8, f32) += A(16x8, tf32, row) * B(8x8, tf32, col)
__device__ __forceinline__ void mma_16n8k8(float* d, const uint32_t* a, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NQ>
__device__ __forceinline__ void bm_issue_batch(BmSmem<NQ>& sm, int gstage, int fstage, int idbuf, int cnt, int K, bool vec, bool color,
                                               const float* __restrict__ geo, const float* __restrict__ features)
{
    const int tid = threadIdx.x;
    if (tid < cnt * 2) {
        const int j = tid >> 1, h = tid & 1;
        cp_async16(&sm.geo[gstage][j][h], geo + 8 * (size_t)sm.ids[idbuf][j] + 4 * h);
    }
    if (!color) return;
    if (vec) {
        const int nq = K >> 2;
        for (int c = tid; c < cnt * nq; c += TILE_PIX) {
            const int j = c / nq, q = c - j * nq;
            cp_async16(&sm.feat[fstage][j][q], features + (size_t)sm.ids[idbuf][j] * K + 4 * q);
        }
    } else {
        float* f = reinterpret_cast<float*>(&sm.feat[fstage][0][0]);
        for (int c = tid; c < cnt * K; c += TILE_PIX) {
            const int j = c / K, k = c - j * K;
            f[j * (4 * NQ) + k] = features[(size_t)sm.ids[idbuf][j] * K + k];
        }
    }
}

template <int NQ>
__device__ __forceinline__ void bm_pad_geo(BmSmem<NQ>& sm, int gstage, int cnt)
{
    const int tid = threadIdx.x;
    if (tid >= cnt && tid < BM_NB) {   // records past the end of the batch: never accepted (threshold = +inf)
        sm.geo[gstage][tid][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.geo[gstage][tid][1] = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
    }
}

// NQ : float4 groups covering the gradient channels (K colour channels [+ 1 mask channel when MD])
// VEC: K % 4 == 0 and no mask channel -> dL_dcolors rows are 16-byte aligned, use red.v4
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
    constexpr int NQE = Cfg::NQE, NT = Cfg::NT, ACC_N = Cfg::ACC_N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    asm volatile("" : "+f"(pixx), "+f"(pixy));   // keep nvcc from rematerialising them in the hot loop
    const size_t plane = (size_t)H * W;

    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;

    // ---- one-time setup -------------------------------------------------------------------------------------------
    if (tid == 0) { sm.max_contrib = 0; sm.contrib = 0; }
    if (tid < BM_NB) sm.touched[tid] = 0;
    {   // this warp's operand tiles start out zero
        float4* w4 = reinterpret_cast<float4*>(&sm.WQ[warp][0]);
#pragma unroll
        for (int i = 0; i < 8; i++) w4[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!VEC || (K >> 2) < NQ) {   // zero the padded feature channels once
        float* f = reinterpret_cast<float*>(&sm.feat[0][0][0]);
        for (int c = tid; c < 2 * BM_NB * 4 * NQ; c += TILE_PIX) f[c] = 0.f;
    }
    __syncthreads();
    int warp_n = my_n;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_n = max(warp_n, __shfl_xor_sync(0xffffffffu, warp_n, o));
    if (lane == 0 && warp_n > 0) atomicMax(&sm.max_contrib, (uint32_t)warp_n);

    // upstream gradient row of this pixel -> swizzled smem row (quad q at physical quad (q + rl) % NQE)
    float bgdot = 0.f;
    {
        float gmask = 0.f;
        if (MD) gmask = inside ? dL_dout_mask[pix_id] : 0.f;
#pragma unroll
        for (int q = 0; q < NQE; q++) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int k = 4 * q + c;
                float x = 0.f;
                if (COLOR && k < K) {
                    x = inside ? dL_dpix[(size_t)k * plane + pix_id] : 0.f;
                    bgdot += bg[k] * x;
                }
                if (MD && k == K) x = gmask;   // the mask gradient rides as channel K of the colour product
                v[c] = x;
            }
            *reinterpret_cast<float4*>(&sm.Gs[rl][4 * ((q + rl) & (NQE - 1))]) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();

    const int maxc = min((int)sm.max_contrib, total);
    if (maxc <= 0) return;
    const int nbatch = (maxc + BM_NB - 1) / BM_NB;
    // batch b covers list positions pos_hi(b) - jj, jj = 0 .. cnt(b)-1, with pos_hi(b) = maxc - 1 - b*NB
    auto batch_cnt = [&](int b) { return min(BM_NB, maxc - b * BM_NB); };
    auto load_id = [&](int b, int jj) { return point_list[range.x + (maxc - 1 - b * BM_NB - jj)]; };

    // prologue: ids(0), ids(1); records + features of batch 0
    if (tid < batch_cnt(0)) sm.ids[0][tid] = load_id(0, tid);
    __syncthreads();
    bm_issue_batch<NQ>(sm, 0, 0, 0, batch_cnt(0), K, VEC, COLOR, geo, features);
    cp_async_commit();
    if (nbatch > 1 && tid < batch_cnt(1)) sm.ids[1][tid] = load_id(1, tid);
    cp_async_wait_all();
    bm_pad_geo<NQ>(sm, 0, batch_cnt(0));
    __syncthreads();

    float T = T_final;
    float acc_r = 0.f, last_alpha = 0.f, last_s = 0.f;
    bool dirty = false;   // this warp's operand tiles hold non-zero entries

    // mma fragment coordinates of this lane
    const int fg = lane >> 2, ft = lane & 3;
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;   // (0.5 * W) rounded to float, as the reference
    const float tcx = (float)tile_x0 + 7.5f, tcy = (float)tile_y0 + 7.5f;
    // partial results of all warps -> final sums of one batch -> global reductions (16 threads per splat)
    auto flush_batch = [&](int pb) {
        const int st = pb & 1, idb = pb % 3;
        const int pcnt = batch_cnt(pb);
        const int jj = tid >> 4, l16 = tid & 15;
        const bool live = jj < pcnt && sm.touched[jj] != 0u;
        const uint32_t cmask = sm.contrib;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);      // colour quad `l16` of splat jj
        float4 mq = make_float4(0.f, 0.f, 0.f, 0.f);     // moment quad (lanes NQE and NQE+1 of the 16-lane group)
        const int mh = (l16 == (NQE & 15)) ? 0 : (l16 == ((NQE + 1) & 15)) ? 1 : -1;
        if (live) {
            for (int w = 0; w < 8; w++) {
                if (!((cmask >> w) & 1u)) continue;
                const float* P = &sm.WQ[w][0] + jj * ACC_N;
                if (l16 < NQE) {
                    const float4 v = *reinterpret_cast<const float4*>(P + 4 * l16);
                    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                }
                if (mh >= 0) {
                    const float4 v = *reinterpret_cast<const float4*>(P + 8 * NT + 4 * mh);
                    mq.x += v.x; mq.y += v.y; mq.z += v.z; mq.w += v.w;
                }
            }
        }
        const float m0 = __shfl_sync(0xffffffffu, mq.x, NQE & 15, 16);
        const float mx = __shfl_sync(0xffffffffu, mq.y, NQE & 15, 16);
        const float my = __shfl_sync(0xffffffffu, mq.z, NQE & 15, 16);
        const float mxx = __shfl_sync(0xffffffffu, mq.w, NQE & 15, 16);
        const float mxy = __shfl_sync(0xffffffffu, mq.x, (NQE + 1) & 15, 16);
        const float myy = __shfl_sync(0xffffffffu, mq.y, (NQE + 1) & 15, 16);
        if (live) {
            const uint32_t id = sm.ids[idb][jj];
            if (l16 < NQE) {
                if (VEC) {
                    if (4 * l16 < K) red_add_v4(dL_dcolors + (size_t)id * K + 4 * l16, a.x, a.y, a.z, a.w);
                } else {
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int ch = 4 * l16 + c;
                        if (COLOR && ch < K) red_add(dL_dcolors + (size_t)id * K + ch, av[c]);
                        else if (MD && ch == K) red_add(ggrad + (size_t)id * GG_STRIDE + 6, av[c]);
                    }
                }
            }
            if (l16 < 6) {
                const float4 g0 = sm.geo[st][jj][0];
                const float4 g1 = sm.geo[st][jj][1];
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
                if (l16 == 0) { v = m0; slot = 5; }                                           // dL/dopacity
                else if (l16 == 1) { v = -o * half_W * (conx * Sx + cony * Sy); slot = 0; }   // dL/dmean2D.x
                else if (l16 == 2) { v = -o * half_H * (conz * Sy + cony * Sx); slot = 1; }   // dL/dmean2D.y
                else if (l16 == 3) { v = -0.5f * o * Sxx; slot = 2; }                         // dL/dconic.x
                else if (l16 == 4) { v = -0.5f * o * Sxy; slot = 3; }                         // dL/dconic.y
                else { v = -0.5f * o * Syy; slot = 4; }                                       // dL/dconic.w
                red_add(ggrad + (size_t)id * GG_STRIDE + slot, v);
            }
        }
    };

    for (int b = 0; b < nbatch; b++) {
        const int fstage = b & 1, gstage = b & 1;
        const int cnt = batch_cnt(b);
        const int pos_hi = maxc - 1 - b * BM_NB;

        // data of batch b+1 starts flying; ids of batch b+2 into a register
        if (b + 1 < nbatch) {
            bm_issue_batch<NQ>(sm, gstage ^ 1, fstage ^ 1, (b + 1) % 3, batch_cnt(b + 1), K, VEC, COLOR, geo, features);
            cp_async_commit();
        }
        uint32_t next_id = 0;
        const bool have_next_id = (b + 2 < nbatch) && tid < batch_cnt(b + 2);
        if (have_next_id) next_id = load_id(b + 2, tid);

        // ---------------- phase A: thread = pixel ----------------
        bool warp_any = false;
        if (pos_hi - (cnt - 1) < warp_n) {   // some pixel of this warp still has contributors in this batch
            if (dirty) {   // the slab still holds last batch's operands / partial result
                float4* w4 = reinterpret_cast<float4*>(&sm.WQ[warp][0]);
#pragma unroll
                for (int i = 0; i < 8; i++) w4[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
                dirty = false;
                __syncwarp();
            }
            for (int j0 = 0; j0 < cnt; j0 += 4) {
                if (pos_hi - (j0 + 3) >= warp_n) continue;                              // warp-uniform
                float pw[4], op[4];
                bool cd[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 g0 = sm.geo[gstage][j0 + i][0];
                    const float4 g1 = sm.geo[gstage][j0 + i][1];
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
                    if (cd[i]) {
                        const float G = expf(pw[i]);
                        const float alpha = fminf(0.99f, op[i] * G);
                        if (!(alpha < 1.0f / 255.0f)) {
                            T = T / (1.f - alpha);
                            float s = 0.f;
                            if (COLOR) {
                                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                                for (int q = 0; q < NQ; q++) {
                                    const float4 f = sm.feat[fstage][jj][q];
                                    const float4 gv = *reinterpret_cast<const float4*>(&sm.Gs[rl][4 * ((q + rl) & (NQE - 1))]);
                                    s0 += f.x * gv.x;
                                    s1 += f.y * gv.y;
                                    s2 += f.z * gv.z;
                                    s3 += f.w * gv.w;
                                }
                                s = (s0 + s1) + (s2 + s3);
                            }
                            acc_r = last_alpha * last_s + (1.f - last_alpha) * acc_r;
                            last_s = s;
                            float dL_dalpha = (s - acc_r) * T;
                            last_alpha = alpha;
                            dL_dalpha += (-T_final / (1.f - alpha)) * bgdot;
                            const int col = (lane + 4 * jj) & 31;
                            sm.WQ[warp][jj * 32 + col] = alpha * T;
                            sm.WQ[warp][512 + jj * 32 + col] = G * dL_dalpha;
                            blended = true;
                        }
                    }
                    if (__any_sync(0xffffffffu, blended)) {
                        warp_any = true;
                        if (lane == 0) sm.touched[jj] = 1u;
                    }
                }
            }
        }

        // ---------------- phase B: this warp's 16 x 32 tiles times its 32 gradient rows / the moment basis ----------------
        if (warp_any) {
            dirty = true;
            __syncwarp();
            float d[NT + 1][4];
#pragma unroll
            for (int n = 0; n <= NT; n++) d[n][0] = d[n][1] = d[n][2] = d[n][3] = 0.f;
            const float* Wt = &sm.WQ[warp][0];
            const float* Qt = &sm.WQ[warp][512];
#pragma unroll 1
            for (int ks = 0; ks < 4; ks++) {
                const int c0 = (ks * 8 + ft + 4 * fg) & 31, c1 = (ks * 8 + ft + 4 + 4 * fg) & 31;
                uint32_t awh[4], awl[4], aqh[4], aql[4];
                split_tf32(Wt[fg * 32 + c0], awh[0], awl[0]);
                split_tf32(Wt[(fg + 8) * 32 + c0], awh[1], awl[1]);
                split_tf32(Wt[fg * 32 + c1], awh[2], awl[2]);
                split_tf32(Wt[(fg + 8) * 32 + c1], awh[3], awl[3]);
                split_tf32(Qt[fg * 32 + c0], aqh[0], aql[0]);
                split_tf32(Qt[(fg + 8) * 32 + c0], aqh[1], aql[1]);
                split_tf32(Qt[fg * 32 + c1], aqh[2], aql[2]);
                split_tf32(Qt[(fg + 8) * 32 + c1], aqh[3], aql[3]);
                // rows of the gradient tile: warp-local pixels ks*8 + ft and + 4  (same tile row, x and x+4)
                const int r0 = ((warp >> 1) * 4 + ks) * TILE_X + (warp & 1) * 8 + ft;
                const int r1 = r0 + 4;
#pragma unroll
                for (int n = 0; n < NT; n++) {
                    const int ch = 8 * n + fg;
                    const float b0f = sm.Gs[r0][4 * (((ch >> 2) + r0) & (NQE - 1)) + (ch & 3)];
                    const float b1f = sm.Gs[r1][4 * (((ch >> 2) + r1) & (NQE - 1)) + (ch & 3)];
                    uint32_t b0h, b0l, b1h, b1l;
                    split_tf32(b0f, b0h, b0l);
                    split_tf32(b1f, b1h, b1l);
                    mma_16n8k8(d[n], awl, b0h, b1h);
                    mma_16n8k8(d[n], awh, b0l, b1l);
                    mma_16n8k8(d[n], awh, b0h, b1h);
                }
                // B fragments of the moment basis X[p][m], p = warp-local pixel ks*8 + ft (+4), m = fg; exact in tf32
                const float yb = (float)((warp >> 1) * 4 + ks) - 7.5f;
                const float xb0 = (float)((warp & 1) * 8 + ft) - 7.5f, xb1 = xb0 + 4.f;
                const float v0 = (fg == 0) ? 1.f : (fg == 1) ? xb0 : (fg == 2) ? yb : (fg == 3) ? xb0 * xb0 : (fg == 4) ? xb0 * yb : (fg == 5) ? yb * yb : 0.f;
                const float v1 = (fg == 0) ? 1.f : (fg == 1) ? xb1 : (fg == 2) ? yb : (fg == 3) ? xb1 * xb1 : (fg == 4) ? xb1 * yb : (fg == 5) ? yb * yb : 0.f;
                mma_16n8k8(d[NT], aql, __float_as_uint(v0), __float_as_uint(v1));
                mma_16n8k8(d[NT], aqh, __float_as_uint(v0), __float_as_uint(v1));
            }
            // the operand tiles are dead now: the slab takes the warp's partial result P[16][ACC_N]
            __syncwarp();
            float* P = &sm.WQ[warp][0];
#pragma unroll
            for (int n = 0; n <= NT; n++) {
                *reinterpret_cast<float2*>(P + fg * ACC_N + 8 * n + 2 * ft) = make_float2(d[n][0], d[n][1]);
                *reinterpret_cast<float2*>(P + (fg + 8) * ACC_N + 8 * n + 2 * ft) = make_float2(d[n][2], d[n][3]);
            }
            if (lane == 0) atomicOr(&sm.contrib, 1u << warp);
        }

        // publish ids(b+2); wait for the copies of batch b+1
        if (have_next_id) sm.ids[(b + 2) % 3][tid] = next_id;
        cp_async_wait_all();
        if (b + 1 < nbatch) bm_pad_geo<NQ>(sm, gstage ^ 1, batch_cnt(b + 1));
        __syncthreads();         // all partial results of batch b are in the slabs (the imbalanced part ends here)
        flush_batch(b);          // sum them, one global reduction per (splat, quad) + six per splat
        __syncthreads();         // slabs may be rewritten (short, balanced section between the two barriers)
        if (tid < BM_NB) sm.touched[tid] = 0;
        if (tid == BM_NB) sm.contrib = 0;
    }
}

template <int NQ, bool VEC, bool MD, bool COLOR>
static int launch_bwd_mma_t(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                            const uint32_t* point_list, const float* features, float* ggrad, cudaStream_t s, bool debug)
{
    auto kern = render_backward_mma_kernel<NQ, VEC, MD, COLOR>;
    const size_t smem = sizeof(BmSmem<NQ>);
    SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(d.tiles_x, d.tiles_y);
    kern<<<grid, TILE_PIX, smem, s>>>(im.ranges, point_list, d.W, d.H, d.C, a.background, g.geo, features,
                                      im.final_T, im.n_contrib, a.dL_dout_color, a.dL_dout_mask, ggrad,
                                      a.dL_dcolors);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

int launch_render_backward_mma(const sagars_backward_args& a, const Dims& d, GeomView g, ImageView im,
                               const uint32_t* point_list, float* ggrad, cudaStream_t s, bool debug)
{
    const bool md = (a.flags & SAGARS_FLAG_MASK_DEPTH) != 0;
    const bool mask_only = (a.flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    if (mask_only) return launch_bwd_mma_t<1, false, true, false>(a, d, g, im, point_list, features, ggrad, s, debug);
    const bool vec = (K % 4) == 0 && !md;
    const int nq = (K + (md ? 1 : 0) + 3) / 4;
#define SAGARS_BWDM_CASE(NQ_)                                                                                        \
    if (nq <= NQ_) {                                                                                                 \
        if (md) return launch_bwd_mma_t<NQ_, false, true, true>(a, d, g, im, point_list, features, ggrad, s, debug);  \
        return vec ? launch_bwd_mma_t<NQ_, true, false, true>(a, d, g, im, point_list, features, ggrad, s, debug)     \
                   : launch_bwd_mma_t<NQ_, false, false, true>(a, d, g, im, point_list, features, ggrad, s, debug);   \
    }
    SAGARS_BWDM_CASE(1)
    SAGARS_BWDM_CASE(2)
    SAGARS_BWDM_CASE(4)
    SAGARS_BWDM_CASE(8)
    SAGARS_BWDM_CASE(16)
#undef SAGARS_BWDM_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
