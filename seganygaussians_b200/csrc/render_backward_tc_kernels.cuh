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

// Optional timeline instrumentation (build with -DSAGARS_BT_TIMELINE): lane 0 of warps 0, 4 and 7 accumulates clock64() deltas per
// phase into sagars_bt_timeline[] (tools/bt_timeline.py prints them).  Off by default: no code is generated.
#if defined(SAGARS_BT_TIMELINE) && !defined(SAGARS_CUDA_EMU)
__device__ unsigned long long sagars_bt_timeline[32];
#define BT_TL_DECL long long tl_t0 = clock64(), tl_t1
#define BT_TL(slot) do { if (lane == 0) { tl_t1 = clock64(); atomicAdd(&sagars_bt_timeline[slot], (unsigned long long)(tl_t1 - tl_t0)); tl_t0 = tl_t1; } } while (0)
#else
#define BT_TL_DECL do {} while (0)
#define BT_TL(slot) do {} while (0)
#endif

namespace sagars {

constexpr int BT_C = 32;            // colour channels (this kernel: C = 32 only)
constexpr int BT_PIX = 128;         // pixels per CTA: half a tile, 16 wide x 8 high = four 8x4 blocks, one consumer warp each
constexpr int BT_THREADS = 256;     // warps 0..3 consumers (thread = pixel), 4..6 epilogue / gradient-product issuers, 7 producer
constexpr int BT_NB = 16;           // candidates per batch
constexpr int BT_RING = 64;         // candidate ring of the producer (up to NB - 1 carried over + 32 new)
constexpr int BT_PRE_ID = 16, BT_PRE_REC = 8;   // prefetch depth of the producer in chunks of 32 list entries (ids / records)
constexpr int BT_TABLES = 16;       // published batch tables (ring): the producer selects several batches ahead of the tensor core

// Measured on B200 before this layout was fixed (tools/probes/, profiles/r2_tcgen05_probes.md):
//  * kind::tf32 with an MN-major shared-memory operand leaves an all-zero accumulator; an A operand in tensor memory works;
//  * ONE thread issues a tcgen05.mma every ~120-200 cycles whatever its size (N = 16 .. 128 cost the same), chains issued by
//    different warps of a CTA overlap.  Small products are therefore bound by the number of MMA instructions per issuing thread:
//    few, wide instructions, and more than one issuer.
//
// Operand tiles (SWIZZLE_NONE canonical K-major layouts, tc.cuh; byte offsets):
//  G (TMEM, columns [0,64))  row = pixel: 32 tf32-exact high parts of the pixel's upstream gradient, then the 32 remainders.
//      A operand of S = G F^T (contraction over the channels); written once with tcgen05.st, thread = pixel.
//  F[2] (shared) [k-chunk = channel / 4][row / 8][row % 8][4 floats], rows 0..15 the high parts of the batch's feature rows, rows
//      16..31 the remainders: ONE N = 32 instruction per k-step gives G.F_hi (columns j) and G.F_lo (columns 16 + j); issued for
//      A = G_hi and A = G_lo, 8 instructions leave all four hi / lo cross terms in S.
//  G3 (shared) [k-chunk = p / 4][row / 8][row % 8][4 floats]: rows 0..31 high parts of the gradient channels, 32..63 remainders,
//      64..69 the moment basis (1, x, y, x^2, x y, y^2), 70..71 zero.  A operand of the gradient product (contraction over the
//      pixels).  The k-chunk step is 9 row groups + 16 B (bank spread of the one-time transposing stores); an M = 128 instruction
//      reads 16 row groups from each k-chunk: groups 9..15 alias the next chunk (or the slack behind the tile); those accumulator
//      rows are never read.
//  B3 (shared) [k-chunk = p / 4][n / 8][n % 8][4 floats], n = column: [0,16) w_hi, [16,32) w_lo, [32,48) q_hi, [48,64) q_lo of the
//      batch's candidates; k-chunk step padded by 16 B so that a warp's 32 scalar stores hit 32 different banks.
//  D3 (TMEM): two partial accumulators of the gradient product (pixels 0..63 and 64..127), one per issuing warp.
struct BtCfg {
    static constexpr int N3 = 4 * BT_NB;                    // columns of the gradient product
    static constexpr int G3_LBO = 9 * 128 + 16;             // bytes between k-chunks of G3
    static constexpr int G3_BYTES = (BT_PIX / 4) * G3_LBO + 7 * 128;
    static constexpr int B3_LBO = (N3 / 8) * 128 + 16;      // bytes between k-chunks
    static constexpr int B3_BYTES = (BT_PIX / 4) * B3_LBO;
    static constexpr int F_LBO = 4 * 128;                   // 32 rows
    static constexpr int F_BYTES = (BT_C / 4) * F_LBO;
    // tensor memory columns: G 64 | S 2 x 32 | D3 2 partials x 64
    static constexpr int TM_G = 0, TM_S = 64, TM_D3 = 128, TMEM_COLS = 256;
};

struct BtTable {                    // one batch, written by the producer, read by the consumers and the epilogue warps
    float4 rec[BT_NB][2];           // record (x, y, cx, cy | cz, opacity, accept_threshold, list position)
    uint32_t member[BT_NB];         // bit w: candidate of consumer warp w's pixel block
    uint32_t id[BT_NB];             // Gaussian id
    int32_t count, last;
    int32_t pad[2];
};

struct BtSmem {
    unsigned char G3[BtCfg::G3_BYTES];
    unsigned char B3[BtCfg::B3_BYTES];
    unsigned char F[2][BtCfg::F_BYTES];
    BtTable tab[BT_TABLES];
    uint32_t pre_id[BT_PRE_ID][32];    // producer-private prefetch rings: list entries 16 chunks ahead,
    float4 pre_rec[BT_PRE_REC][32][2]; // their records 8 chunks ahead (cp.async; lane = splat)
    float4 ring_rec[BT_RING][2];    // producer-private candidate ring
    uint32_t ring_id[BT_RING];
    uint32_t ring_mem[BT_RING];
    float mom[8][BT_NB];            // moment rows of a batch (epilogue warp 6)
    int32_t red_n[4];
    int32_t npub;                   // batches published so far (release / acquire)
    int32_t nfinal;                 // -1 until the list is exhausted, then the total number of batches
    uint64_t bar_setup;             // the gradient tiles (shared + tensor memory) are written (four consumer warps)
    uint64_t bar_f[2];              // F tile of batch parity p written (four consumer warps)
    uint64_t bar_s[2];              // S of batch parity p ready (tcgen05.commit)
    uint64_t bar_b3;                // B3 written by the four consumer warps (one phase per batch)
    uint64_t bar_d3;                // both halves of the gradient product complete (two tcgen05.commit; one phase per batch)
    uint64_t bar_d3free;            // the three epilogue warps have read D3 (one phase per batch)
    uint32_t tmem_base;
};

// C = 32, colour only (no mask / depth channel).  Grid = (tiles_x, 2 * tiles_y), 256 threads.
__global__ void __launch_bounds__(BT_THREADS, 2)
render_backward_tc_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H,
                          const float* __restrict__ bg, const float* __restrict__ geo, const float* __restrict__ features,
                          const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                          const float* __restrict__ dL_dpix, float* __restrict__ ggrad, float* __restrict__ dL_dcolors)
{
    using Cfg = BtCfg;
    constexpr int NB = BT_NB;
    SAGARS_DYNAMIC_SMEM_1024(smem_raw);
    BtSmem& sm = *reinterpret_cast<BtSmem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile_x0 = blockIdx.x * TILE_X, half_y0 = blockIdx.y * 8;

    const uint2 range = ranges[(blockIdx.y >> 1) * gridDim.x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    if (total <= 0) return;   // empty tile (uniform over the CTA)
    BT_TL_DECL;

    // ---- consumer identity: warp w < 4 owns the 8x4 pixel block w of the group, thread = pixel ----
    const int cw = warp & 3;
    const uint32_t blk_x0 = tile_x0 + (cw & 1) * 8, blk_y0 = half_y0 + (cw >> 1) * 4;
    const uint32_t px = blk_x0 + (lane & 7), py = blk_y0 + (lane >> 3);
    const bool inside = warp < 4 && px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    SAGARS_PIN_F2(pixx, pixy);
    const size_t plane = (size_t)H * W;
    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    const int my_n = inside ? (int)n_contrib[pix_id] : 0;
    if (warp < 4) {
        int wn = my_n;            // deepest contributor of this warp's block
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wn = max(wn, __shfl_xor_sync(0xffffffffu, wn, o));
        if (lane == 0) sm.red_n[warp] = wn;
    }
    // ---- one-time setup, part 1: mbarriers, TMEM ----
    if (tid == 0) {
#pragma unroll
        for (int p = 0; p < 2; p++) {
            tc::mbar_init(&sm.bar_f[p], 4);
            tc::mbar_init(&sm.bar_s[p], 1);
        }
        tc::mbar_init(&sm.bar_setup, 4);
        tc::mbar_init(&sm.bar_b3, 4);
        tc::mbar_init(&sm.bar_d3, 2);
        tc::mbar_init(&sm.bar_d3free, 3);
        sm.npub = 0;
        sm.nfinal = -1;
        tc::mbar_init_fence();
    }
    if (warp == 7) {
        tc::tmem_alloc<Cfg::TMEM_COLS>(&sm.tmem_base);
        float* z = reinterpret_cast<float*>(sm.G3) + (BT_PIX / 4) * (Cfg::G3_LBO / 4);     // slack behind G3
        for (int i = lane; i < 7 * 32; i += 32) z[i] = 0.f;
        tc::fence_smem_to_async_proxy();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_sel = (uint32_t)(cw * 32) << 16;      // TMEM lanes a warp may touch: 32 * (warp % 4) ..
    const int wn0 = sm.red_n[0], wn1 = sm.red_n[1], wn2 = sm.red_n[2], wn3 = sm.red_n[3];
    const int maxc = min(max(max(wn0, wn1), max(wn2, wn3)), total);
    if (maxc <= 0) {          // uniform: no pixel of the group has a contributor
        if (warp == 7) tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem);
        return;
    }

    // ---- part 2 (consumers only; the producer is already selecting): the pixel's gradient row -> G3 (shared) and G (TMEM) ----
    float bgdot = 0.f;
    if (warp < 4) {
        const int gt = warp * 32 + lane;                      // pixel index in the group = accumulator row of S = k index of G3 / B3
        float* g3 = reinterpret_cast<float*>(sm.G3) + (gt >> 2) * (Cfg::G3_LBO / 4) + (gt & 3);       // row r at (r / 8) * 32 + (r % 8) * 4
        float gh[BT_C], gl[BT_C];
#pragma unroll
        for (int k = 0; k < BT_C; k++) gh[k] = inside ? dL_dpix[(size_t)k * plane + pix_id] : 0.f;      // 32 loads in flight
#pragma unroll
        for (int k = 0; k < BT_C; k++) {
            const float g = gh[k];
            bgdot += bg[k] * g;
            gh[k] = tc::tf32_hi(g);
            gl[k] = g - gh[k];
            g3[(k >> 3) * 32 + (k & 7) * 4] = gh[k];
            g3[((32 + k) >> 3) * 32 + (k & 7) * 4] = gl[k];
        }
        tc::tmem_st32(tmem + Cfg::TM_G + lane_sel, gh);
        tc::tmem_st32(tmem + Cfg::TM_G + 32 + lane_sel, gl);
        // pixel coordinates relative to the centre of the 16 x 8 pixel group: multiples of 0.5, squares exact in tf32
        const float xr = (float)((warp & 1) * 8 + (lane & 7)) - 7.5f, yr = (float)((warp >> 1) * 4 + (lane >> 3)) - 3.5f;
        const float basis[8] = {1.f, xr, yr, xr * xr, xr * yr, yr * yr, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; k++) g3[8 * 32 + k * 4] = basis[k];
        tc::fence_smem_to_async_proxy();
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&sm.bar_setup);      // the issuers wait for the four consumer warps before their first MMA
    }

    if (warp == 0) BT_TL(0);          // setup
    if (warp == 7) {
        // =====================================================================================================================
        // PRODUCER.  Two duties, polled in turn (neither blocks on another role):
        //  (1) selection: walk the tile's list back to front in chunks of 32 (lane = splat), test every splat against the four pixel
        //      blocks, keep the union in a ring and PUBLISH batches of NB candidates (table ring, release store of the count) ahead
        //      of the tensor core -- the consumers gather a batch's feature rows two batches before they blend it;
        //  (2) S = G F^T of the next batch whose F tile the consumers have written: 8 tcgen05.mma (A = G_hi, then G_lo, from
        //      tensor memory; B = [F_hi; F_lo], N = 32).
        // =====================================================================================================================
        constexpr uint32_t ID1 = tc::idesc_tf32(128, 32, 0, 0);
        const int nchunk = (maxc + 31) >> 5;
        auto chunk_pos = [&](int c) { return maxc - 1 - 32 * c - lane; };     // back to front
        const float gx0 = (float)tile_x0, gy0 = (float)half_y0;
        const uint32_t lt = (1u << lane) - 1u;
        int head = 0, ntab = 0;
        int chunk = 0;                               // next chunk to test
        int pub = 0, g1 = 0;                         // batches published / S issued
        bool list_done = false;

        // Deep prefetch (the tile's list is walked once, ~2,000-cycle loads each depending on the previous one: list entry -> record):
        // entries 16 chunks ahead and records 8 chunks ahead travel with cp.async into two private rings; every test_chunk commits
        // exactly one group {entries of chunk c + 16, records of chunk c + 8}, so "all but the 7 newest groups have landed" is
        // precisely "the records of chunk c and the entries of chunk c + 8 are here".
        auto issue_ids = [&](int c) {
            const int pos = (c < nchunk) ? chunk_pos(c) : -1;
            if (pos >= 0) cp_async4(&sm.pre_id[c & (BT_PRE_ID - 1)][lane], point_list + range.x + pos);
        };
        auto issue_recs = [&](int c) {            // the entries of chunk c have landed
            const int pos = (c < nchunk) ? chunk_pos(c) : -1;
            if (pos >= 0) {
                const uint32_t id = sm.pre_id[c & (BT_PRE_ID - 1)][lane];
                cp_async16(&sm.pre_rec[c & (BT_PRE_REC - 1)][lane][0], geo + 8 * (size_t)id);
                cp_async16(&sm.pre_rec[c & (BT_PRE_REC - 1)][lane][1], geo + 8 * (size_t)id + 4);
            }
        };
        for (int c = 0; c < BT_PRE_ID; c++) issue_ids(c);
        cp_async_commit();
        cp_async_wait_all();
        for (int c = 0; c < BT_PRE_REC; c++) { issue_recs(c); cp_async_commit(); }

        auto publish = [&](int m, bool last) {       // ring slots [head, head + m) -> table pub
            BtTable& t = sm.tab[pub & (BT_TABLES - 1)];
            if (lane < NB) {
                const int slot = (head + min(lane, max(m - 1, 0))) & (BT_RING - 1);
                t.rec[lane][0] = sm.ring_rec[slot][0];
                t.rec[lane][1] = sm.ring_rec[slot][1];
                t.member[lane] = lane < m ? sm.ring_mem[slot] : 0u;
                t.id[lane] = sm.ring_id[slot];
            }
            if (lane == 0) { t.count = m; t.last = last ? 1 : 0; }
            __syncwarp();
            pub++;
            if (lane == 0) {
                if (last) tc::st_release_cta(&sm.nfinal, pub);
                tc::st_release_cta(&sm.npub, pub);
            }
            head = (head + m) & (BT_RING - 1);
            ntab -= m;
            if (last) list_done = true;
        };
        auto test_chunk = [&]() {
            cp_async_wait_group<BT_PRE_REC - 1>();
            const int pos_cur = chunk_pos(chunk);
            const uint32_t id_cur = pos_cur >= 0 ? sm.pre_id[chunk & (BT_PRE_ID - 1)][lane] : 0u;
            float4 r0_cur = make_float4(0.f, 0.f, 0.f, 0.f), r1_cur = r0_cur;
            if (pos_cur >= 0) {
                r0_cur = sm.pre_rec[chunk & (BT_PRE_REC - 1)][lane][0];
                r1_cur = sm.pre_rec[chunk & (BT_PRE_REC - 1)][lane][1];
            }
            // only now may the slots of this chunk be overwritten (a copy can land at any time after it is issued)
            issue_ids(chunk + BT_PRE_ID);
            issue_recs(chunk + BT_PRE_REC);
            cp_async_commit();
            // candidate test (candidate.cuh), lane = splat, against the 16 x 8 pixel group (the consumers' per-pixel test decides)
            const uint32_t mem = (pos_cur >= 0 && !block_rejects(r0_cur, r1_cur, gx0, gx0 + 15.f, gy0, gy0 + 7.f)) ? 1u : 0u;
            const uint32_t any = __ballot_sync(0xffffffffu, mem != 0u);
            if (mem != 0u) {
                const int slot = (head + ntab + __popc(any & lt)) & (BT_RING - 1);
                float4 r1p = r1_cur;
                r1p.w = __int_as_float(pos_cur);
                sm.ring_rec[slot][0] = r0_cur;
                sm.ring_rec[slot][1] = r1p;
                sm.ring_id[slot] = id_cur;
                sm.ring_mem[slot] = mem;
            }
            ntab += __popc(any);
            __syncwarp();
            chunk++;
        };

        BT_TL(20);
        while (true) {
            bool progress = false;
            // (1) selection.  A table slot is reused 16 batches later: by then the epilogue that read it is long done (it
            //     finishes before the consumers complete two more batches, which S of batch g1 - 1 being issued implies).
            if (!list_done) {
                const bool more = chunk < nchunk;
                if ((ntab > NB || (ntab == NB && more)) || !more) {
                    if (pub <= g1 + 8) { publish(more ? NB : min(ntab, NB), !more && ntab <= NB); progress = true; }
                } else {
                    test_chunk();
                    progress = true;
                }
            }
            if (progress) BT_TL(21);       // selection / publishing
            // (2) S = G F^T (the first one also needs the gradient rows in tensor memory)
            if (g1 < pub) {
                int ok = 0;
                if (lane == 0) ok = (tc::mbar_test_wait(&sm.bar_f[g1 & 1], (uint32_t)((g1 >> 1) & 1)) &&
                                     (g1 > 0 || tc::mbar_test_wait(&sm.bar_setup, 0u))) ? 1 : 0;
                if (__shfl_sync(0xffffffffu, ok, 0)) {
                    tc::fence_after_sync();
                    if (lane == 0) {
                        const int p = g1 & 1;
                        const uint32_t f0 = smem_u32(sm.F[p]);
                        const uint32_t ts = tmem + Cfg::TM_S + p * 32;
#pragma unroll
                        for (int term = 0; term < 2; term++) {
                            const uint32_t a0 = tmem + Cfg::TM_G + term * 32;
#pragma unroll
                            for (int ks = 0; ks < BT_C / 8; ks++)
                                tc::mma_tf32_ts(ts, a0 + ks * 8, tc::smem_desc(f0 + ks * 2 * Cfg::F_LBO, Cfg::F_LBO, 128), ID1,
                                                (term > 0 || ks > 0) ? 1u : 0u);
                        }
                        tc::commit(&sm.bar_s[p]);
                    }
                    __syncwarp();
                    g1++;
                    progress = true;
                    BT_TL(22);             // S issue
                }
            }
            if (list_done && g1 == pub) break;
            if (!progress) { tc::backoff(64); BT_TL(23); }     // idle
        }
    } else if (warp < 4) {
        // =====================================================================================================================
        // CONSUMERS: thread = pixel over the batch (the reference's back-to-front traversal), w = alpha T and q = G dL/dalpha
        // -> B3 (hi / lo).  No CTA barrier: a warp arrives on bar_b3 and goes on to the next batch.  Every thread also gathers one
        // float4 of the feature rows of the batch two ahead (load issued before the blend, consumed after it) into its F tile.
        // =====================================================================================================================
        const int gt = warp * 32 + lane;
        const int fj = gt & 15, fq = gt >> 4;          // F gather: candidate and float4 of its feature row
        // number of batches once known (all published), else "at least `need` are published"
        auto wait_published = [&](int need) {
            int total_b = -1;
            if (lane == 0) {
                while (true) {
                    const int f = tc::ld_acquire_cta(&sm.nfinal);
                    if (f >= 0) { total_b = f; break; }
                    if (tc::ld_acquire_cta(&sm.npub) >= need) break;
                    tc::backoff(32);
                }
            }
            return __shfl_sync(0xffffffffu, total_b, 0);
        };
        auto load_f = [&](int b) {                     // this thread's float4 of batch b's feature rows
            const BtTable& t = sm.tab[b & (BT_TABLES - 1)];
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fj < t.count) f = __ldg(reinterpret_cast<const float4*>(features + (size_t)t.id[fj] * BT_C + 4 * fq));
            return f;
        };
        auto store_f = [&](int p, float4 f) {          // rows fj (high parts) and 16 + fj (remainders)
            float4 hi, lo;
            hi.x = tc::tf32_hi(f.x); lo.x = f.x - hi.x;
            hi.y = tc::tf32_hi(f.y); lo.y = f.y - hi.y;
            hi.z = tc::tf32_hi(f.z); lo.z = f.z - hi.z;
            hi.w = tc::tf32_hi(f.w); lo.w = f.w - hi.w;
            float* base = reinterpret_cast<float*>(sm.F[p]) + fq * (Cfg::F_LBO / 4) + (fj >> 3) * 32 + (fj & 7) * 4;
            *reinterpret_cast<float4*>(base) = hi;
            *reinterpret_cast<float4*>(base + 64) = lo;
        };
        if (warp == 0) BT_TL(1);
        {                                              // prologue: the F tiles of the first two batches, both loads in flight
            const int t0 = wait_published(2);
            const bool have1 = t0 < 0 || t0 >= 2;
            const float4 fa = load_f(0);
            float4 fb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have1) fb = load_f(1);
            store_f(0, fa);
            if (have1) store_f(1, fb);
            tc::fence_smem_to_async_proxy();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&sm.bar_f[0]);
                if (have1) tc::mbar_arrive(&sm.bar_f[1]);
            }
        }
        if (warp == 0) BT_TL(2);          // prologue
        float T = T_final, acc_r = 0.f, pend = 0.f, om = 1.f;      // pend = last_alpha * last_s, om = 1 - last_alpha
        float* const b3w = reinterpret_cast<float*>(sm.B3) + (gt >> 2) * (Cfg::B3_LBO / 4) + (gt & 3);
        for (int b = 0;; b++) {
            const int p = b & 1;
            // the batch two ahead: its table is published well before (the producer runs ahead); start its feature-row load now
            // (four batches ahead through a register queue was measured no faster: 2.04 vs 1.96 ms)
            const int total_b = wait_published(b + 3);
            const bool have2 = total_b < 0 || b + 2 < total_b;
            float4 f2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have2) f2 = load_f(b + 2);
            if (warp == 0) BT_TL(3);      // table wait + load issue
            tc::mbar_wait_long(&sm.bar_s[p], (uint32_t)((b >> 1) & 1));
            if (warp == 0) BT_TL(4);      // wait S
            tc::fence_after_sync();
            const BtTable& bt = sm.tab[b & (BT_TABLES - 1)];
            const bool last = bt.last != 0;
            float s[2 * NB];                                      // columns j: G . F_hi, 16 + j: G . F_lo
            tc::tmem_ld32(tmem + Cfg::TM_S + p * 32 + lane_sel, s);
            // pass 1, independent per candidate: alpha, G, 1 / (1 - alpha); zero / one when the pair does not blend
            float al[NB], Gv[NB], rv[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) {                        // no branches: sixteen independent dependency chains
                s[j] += s[NB + j];
                const float4 g0 = bt.rec[j][0];
                const float4 g1 = bt.rec[j][1];
                const float dx = g0.x - pixx, dy = g0.y - pixy;
                const float pw = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                const bool cd = (__float_as_int(g1.w) < my_n) && !(pw > 0.0f) && (pw >= g1.z);
                const float G = expf(cd ? pw : 0.f);
                const float alpha = fminf(0.99f, g1.y * G);
                const bool ok = cd && !(alpha < 1.0f / 255.0f);
                al[j] = ok ? alpha : 0.f;
                Gv[j] = ok ? G : 0.f;
                rv[j] = rcp_approx(1.f - al[j]);
            }
            // B3 is free once the gradient product of the previous batch has read it (it was issued as soon as this warp and its
            // three neighbours had arrived, a whole pass 1 ago)
            if (warp == 0) BT_TL(5);      // ld S + pass 1
            if (b >= 1) tc::mbar_wait_long(&sm.bar_d3, (uint32_t)((b - 1) & 1));
            if (warp == 0) BT_TL(6);      // wait B3 free
            // pass 2, the recurrences: T_j = T_{j-1} / (1 - alpha_j);  a_j = alpha_{j-1} s_{j-1} + (1 - alpha_{j-1}) a_{j-1}
            // (a candidate that does not blend has alpha = 0: it folds the pending term and contributes nothing itself)
#pragma unroll
            for (int j = 0; j < NB; j++) {
                T = T * rv[j];
                acc_r = fmaf(om, acc_r, pend);
                float dL_dalpha = (s[j] - acc_r) * T;
                if (bgdot != 0.f) dL_dalpha += (-T_final * rv[j]) * bgdot;     // zero background: the term is exactly 0
                const float w = al[j] * T;
                const float q = Gv[j] * dL_dalpha;
                pend = al[j] * s[j];
                om = 1.f - al[j];
                const float wh = tc::tf32_hi(w), qh = tc::tf32_hi(q);
                float* col = b3w + (j >> 3) * 32 + (j & 7) * 4;
                col[0 * 64] = wh;
                col[1 * 64] = w - wh;
                col[2 * 64] = qh;
                col[3 * 64] = q - qh;
            }
            if (warp == 0) BT_TL(7);      // pass 2 + B3 stores
            if (have2) store_f(p, f2);                 // F[p] is free: S of batch b (its last reader) completed before this batch began
            if (warp == 0) BT_TL(16);     // F rows -> tile
            tc::fence_smem_to_async_proxy();
            if (warp == 0) BT_TL(17);     // proxy fence
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&sm.bar_b3);
                if (have2) tc::mbar_arrive(&sm.bar_f[p]);
            }
            if (warp == 0) BT_TL(8);      // F store + fence + arrive
            if (last) break;
        }
    } else {
        // =====================================================================================================================
        // EPILOGUE WARPS 4, 5, 6.  Warps 4 and 5 each ISSUE one half of the gradient product (pixels 0..63 / 64..127, 8 tcgen05.mma
        // into their own partial accumulator) as soon as the consumers have filled B3, then all three move the accumulator rows
        // they can reach (0..31 gradient high parts, 32..63 remainders, 64..69 moments) of both partials to global memory.
        // =====================================================================================================================
        constexpr uint32_t ID3 = tc::idesc_tf32(128, Cfg::N3, 0, 0);
        const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;
        const float gcx = (float)tile_x0 + 7.5f, gcy = (float)half_y0 + 3.5f;     // centre of the pixel group
        const uint32_t g3_addr = smem_u32(sm.G3), b3_addr = smem_u32(sm.B3);
        const int half = warp - 4;                                                // issuer of pixels [64 half, 64 half + 64)
        if (warp == 4) BT_TL(10);
        for (int b = 0;; b++) {
            if (warp < 6) {
                if (lane == 0) {
                    tc::mbar_wait_long(&sm.bar_b3, (uint32_t)(b & 1));                              // a batch takes thousands of cycles
                    if (warp == 4) BT_TL(11);      // wait B3 full
                    if (b >= 1) tc::mbar_wait_long(&sm.bar_d3free, (uint32_t)((b - 1) & 1));
                    if (warp == 4) BT_TL(12);      // wait D3 free
                    tc::fence_after_sync();
#pragma unroll
                    for (int k = 0; k < BT_PIX / 16; k++) {
                        const int ks = half * (BT_PIX / 16) + k;
                        tc::mma_tf32(tmem + Cfg::TM_D3 + half * Cfg::N3, tc::smem_desc(g3_addr + ks * 2 * Cfg::G3_LBO, Cfg::G3_LBO, 128),
                                     tc::smem_desc(b3_addr + ks * 2 * Cfg::B3_LBO, Cfg::B3_LBO, 128), ID3, k > 0 ? 1u : 0u);
                    }
                    tc::commit(&sm.bar_d3);
                    if (warp == 4) BT_TL(13);      // gradient product issue
                }
                __syncwarp();
            }
            tc::mbar_wait_long(&sm.bar_d3, (uint32_t)(b & 1));                             // long waits by design: do not burn issue slots
            if (warp == 4) BT_TL(14);              // wait gradient product
            tc::fence_after_sync();
            const BtTable& et = sm.tab[b & (BT_TABLES - 1)];
            const int m = et.count;
            const bool last = et.last != 0;
            // both partial accumulators; columns: w_hi | w_lo | q_hi | q_lo
            float v0[32], v1[32];
            const uint32_t col0 = (warp < 6) ? 0u : 2u * NB;
            tc::tmem_ld32(tmem + Cfg::TM_D3 + col0 + lane_sel, v0);
            tc::tmem_ld32(tmem + Cfg::TM_D3 + Cfg::N3 + col0 + lane_sel, v1);
            // everything this warp needs from the batch table is read BEFORE the accumulator is handed back (the table slot is
            // recycled 16 batches later)
            uint32_t ids[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) ids[j] = et.id[j];
            const float4 g0 = et.rec[lane & (NB - 1)][0];
            const float4 g1 = et.rec[lane & (NB - 1)][1];
            const uint32_t my_id = et.id[lane & (NB - 1)];
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm.bar_d3free);
            if (warp < 6) {
#pragma unroll
                for (int j = 0; j < NB; j++)
                    if (j < m) red_add(dL_dcolors + (size_t)ids[j] * BT_C + lane, (v0[j] + v0[NB + j]) + (v1[j] + v1[NB + j]));
            } else {
                if (lane < 6) {
#pragma unroll
                    for (int j = 0; j < NB; j++) sm.mom[lane][j] = (v0[j] + v0[NB + j]) + (v1[j] + v1[NB + j]);
                }
                __syncwarp();
                if (lane < m) {
                    const float m0 = sm.mom[0][lane], mx = sm.mom[1][lane], my = sm.mom[2][lane];
                    const float mxx = sm.mom[3][lane], mxy = sm.mom[4][lane], myy = sm.mom[5][lane];
                    const float conx = g0.z, cony = g0.w, conz = g1.x, o = g1.y;
                    // sums over the pixels of q * (1, dx, dy, dx^2, dx dy, dy^2) with d = centre - pixel = c - x'
                    const float cx = g0.x - gcx, cy = g0.y - gcy;
                    const float Sx = cx * m0 - mx;
                    const float Sy = cy * m0 - my;
                    const float Sxx = cx * cx * m0 - 2.f * cx * mx + mxx;
                    const float Sxy = cx * cy * m0 - cx * my - cy * mx + mxy;
                    const float Syy = cy * cy * m0 - 2.f * cy * my + myy;
                    float* gg = ggrad + (size_t)my_id * GG_STRIDE;
                    red_add(gg + 0, -o * half_W * (conx * Sx + cony * Sy));      // dL/dmean2D.x
                    red_add(gg + 1, -o * half_H * (conz * Sy + cony * Sx));      // dL/dmean2D.y
                    red_add(gg + 2, -0.5f * o * Sxx);                            // dL/dconic.x
                    red_add(gg + 3, -0.5f * o * Sxy);                            // dL/dconic.y
                    red_add(gg + 4, -0.5f * o * Syy);                            // dL/dconic.w
                    red_add(gg + 5, m0);                                         // dL/dopacity
                }
                __syncwarp();
            }
            if (warp == 4) BT_TL(15);              // accumulator rows -> global memory
            if (last) break;
        }
    }
    if (warp == 0) BT_TL(9);                       // consumer: waiting for the CTA to finish
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 7) tc::tmem_dealloc<Cfg::TMEM_COLS>(sm.tmem_base);
}

}  // namespace sagars
