// render_forward.cu -- per-tile front-to-back alpha compositing of C feature channels
// (+ optional mask / depth channels).
//
// Semantics: CF cuda_rasterizer/forward.cu:264-385 (DEPTH forward.cu:262-387 adds the mask/depth
// accumulators), SURVEY.md Appendix A.10-A.12.  One CTA per 16x16 tile, one thread per pixel.
// What is this library's own:
//   * the per-instance records ({xy, conic, opacity, depth} = 32 B) and the C-float feature rows of
//     a batch of instances are staged in shared memory by asynchronous 16-byte copies
//     (cp.async / LDGSTS), double buffered so the copies of batch b+1 overlap the blending of batch
//     b; the reference re-reads every feature from global memory per pixel per channel;
//   * a warp covers an 8x4 pixel block (the reference: 16x2), which cuts the number of
//     (warp, Gaussian) pairs that have any pixel to blend;
//   * a warp stops as soon as its 32 pixels are saturated (the reference only stops per CTA, per
//     256-instance batch);
//   * a (warp, splat) pair in which no pixel can pass the reference's tests (power <= 0 and a conservative
//     per-splat lower bound on power that implies alpha >= 1/255) is rejected with one warp vote, before expf;
//   * C is a run-time value (<= 64), dispatched onto float4-group templates;
//   * SAGARS_FLAG_STAGE_TMA selects a second staging engine for the same kernel body: the 32-byte records and
//     (K % 4 == 0) the K*4-byte feature rows of a batch are gathered by bulk asynchronous copies (cp.async.bulk, the
//     TMA unit's non-tensor form: ONE instruction per row, issued by warp 0, no per-thread 16-byte pieces) that complete
//     on an mbarrier per pipeline stage; consumers wait on the barrier's phase instead of cp.async.wait_all.
// The per-pixel arithmetic (power, alpha, the 1/255 and 1e-4 tests, the order of accumulation)
// is kept operation for operation so that n_contrib / final_T / colours match the reference.
#include "common.cuh"
#include <cstdlib>
#include "cp_async.cuh"

namespace sagars {

constexpr int FWD_BATCH = 64;   // instances staged per pipeline stage

template <int NQ>
struct FwdSmem {
    float4 geo[2][FWD_BATCH][2];        // x, y, cx, cy | cz, opacity, accept_threshold, -
    float4 feat[2][FWD_BATCH][NQ];      // feature rows, zero padded to 4*NQ channels
    uint32_t ids[2][FWD_BATCH];
    float maskv[2][FWD_BATCH];          // DEPTH variant: per-instance mask value
    float depthv[2][FWD_BATCH];         // DEPTH variant: per-instance view depth
};

template <int NQ>
struct FwdSmemTma : FwdSmem<NQ> {
    uint64_t bar[2];                    // one mbarrier per pipeline stage (bulk-copy staging only)
};
template <int NQ, bool TMA>
struct FwdSmemSel { using type = FwdSmem<NQ>; };
template <int NQ>
struct FwdSmemSel<NQ, true> { using type = FwdSmemTma<NQ>; };

// bulk-copy staging of one batch: warp 0 announces the bytes of the batch on the stage's mbarrier and gathers the
// rows, one cp.async.bulk per row (record: 32 B; feature row: K*4 B when K % 4 == 0)
template <int NQ, bool VEC, bool COLOR>
__device__ __forceinline__ void fwd_issue_batch_bulk(FwdSmemTma<NQ>& sm, int stage, int idbuf, int cnt, int K,
                                                     const float* __restrict__ geo, const float* __restrict__ features)
{
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const uint32_t row_bytes = (VEC && COLOR) ? (uint32_t)K * 4u : 0u;
    uint64_t* bar = &sm.bar[stage];
    if (lane == 0) mbarrier_arrive_expect_tx(bar, (uint32_t)cnt * (32u + row_bytes));
    __syncwarp();
    for (int j = lane; j < cnt; j += 32) {
        const uint32_t id = sm.ids[idbuf][j];
        bulk_copy_g2s(&sm.geo[stage][j][0], geo + 8 * (size_t)id, 32u, bar);
        if (VEC && COLOR) bulk_copy_g2s(&sm.feat[stage][j][0], features + (size_t)id * K, row_bytes, bar);
    }
}

// issue the asynchronous copies of one batch (ids already in smem)
template <int NQ, bool VEC, bool MD, bool COLOR, bool TMA = false>
__device__ __forceinline__ void fwd_issue_batch(typename FwdSmemSel<NQ, TMA>::type& sm, int stage, int idbuf, int cnt, int K,
                                                const float* __restrict__ geo, const float* __restrict__ features,
                                                const float* __restrict__ mask, const float* __restrict__ depths)
{
    const int tid = threadIdx.x;
    if constexpr (TMA) {
        fwd_issue_batch_bulk<NQ, VEC, COLOR>(sm, stage, idbuf, cnt, K, geo, features);
        if (MD) {
            if (tid < cnt) {
                const uint32_t id = sm.ids[idbuf][tid];
                sm.maskv[stage][tid] = mask[id];
                sm.depthv[stage][tid] = depths[id];
            }
        }
        if (COLOR && !VEC) {
            float* f = reinterpret_cast<float*>(&sm.feat[stage][0][0]);
            for (int c = tid; c < cnt * K; c += TILE_PIX) {
                const int j = c / K, k = c - j * K;
                const uint32_t id = sm.ids[idbuf][j];
                f[j * (4 * NQ) + k] = features[(size_t)id * K + k];
            }
        }
        return;
    }
    // geometry records: 2 x 16 B per instance
    for (int c = tid; c < cnt * 2; c += TILE_PIX) {
        const int j = c >> 1, h = c & 1;
        const uint32_t id = sm.ids[idbuf][j];
        cp_async16(&sm.geo[stage][j][h], geo + 8 * (size_t)id + 4 * h);
    }
    if (MD) {
        if (tid < cnt) {
            const uint32_t id = sm.ids[idbuf][tid];
            sm.maskv[stage][tid] = mask[id];
            sm.depthv[stage][tid] = depths[id];
        }
    }
    if (!COLOR) return;
    if (VEC) {
        const int nq = K >> 2;   // == NQ or fewer (remaining quads stay zero)
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

// records past the end of the tile's list (up to the next multiple of 4): never accepted (threshold = +inf)
template <int NQ>
__device__ __forceinline__ void fwd_pad_batch(FwdSmem<NQ>& sm, int stage, int cnt)
{
    const int tid = threadIdx.x;
    if (tid >= cnt && tid < FWD_BATCH) {
        sm.geo[stage][tid][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.geo[stage][tid][1] = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 0.f);
    }
}

template <int NQ, bool VEC, bool MD, bool COLOR, bool TMA>
__device__ __forceinline__ void
render_forward_body(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                    int W, int H, int K,
                    const float* __restrict__ geo, const float* __restrict__ features,
                    const float* __restrict__ mask, const float* __restrict__ depths, const float* __restrict__ bg,
                    float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                    float* __restrict__ out_color, float* __restrict__ out_mask, float* __restrict__ out_depth)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = typename FwdSmemSel<NQ, TMA>::type;
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = gridDim.x;
    const uint32_t px = blockIdx.x * TILE_X + (warp & 1) * 8 + (lane & 7);
    const uint32_t py = blockIdx.y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    float pixx = (float)px, pixy = (float)py;
    // opaque to the optimiser: otherwise nvcc rematerialises both from %ctaid / %tid inside the hot loop
    asm volatile("" : "+f"(pixx), "+f"(pixy));
    const uint2 range = ranges[blockIdx.y * tiles_x + blockIdx.x];
    const int total = (int)(range.y - range.x);
    const int nbatch = (total + FWD_BATCH - 1) / FWD_BATCH;

    // zero the padded feature channels once (cp.async only ever writes the first K of each row)
    if (!VEC || (K >> 2) < NQ) {
        float* f = reinterpret_cast<float*>(&sm.feat[0][0][0]);
        for (int c = tid; c < 2 * FWD_BATCH * 4 * NQ; c += TILE_PIX) f[c] = 0.f;
    }
    if constexpr (TMA) {
        if (tid == 0) {
            mbarrier_init(&sm.bar[0], 1);
            mbarrier_init(&sm.bar[1], 1);
        }
        // the barriers and the zero fill above (generic proxy) before the first bulk copy (async proxy) touches them;
        // the __syncthreads of the prologue below orders every thread's fence before warp 0 issues
        fence_proxy_async_smem();
    }

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C[4 * NQ];
#pragma unroll
    for (int k = 0; k < 4 * NQ; k++) C[k] = 0.f;
    float Macc = 0.f, Dacc = 0.f;
    bool done = !inside;

    // prologue: ids(0) -> smem, copies of batch 0, ids(1) -> smem
    if (nbatch > 0) {
        if (tid < min(FWD_BATCH, total)) sm.ids[0][tid] = point_list[range.x + tid];
        __syncthreads();
        fwd_issue_batch<NQ, VEC, MD, COLOR, TMA>(sm, 0, 0, min(FWD_BATCH, total), K, geo, features, mask, depths);
        if constexpr (!TMA) cp_async_commit();
        if (nbatch > 1 && tid < min(FWD_BATCH, total - FWD_BATCH)) sm.ids[1][tid] = point_list[range.x + FWD_BATCH + tid];
        if constexpr (TMA) mbarrier_wait_parity(&sm.bar[0], 0u);
        else cp_async_wait_all();
        fwd_pad_batch<NQ>(sm, 0, min(FWD_BATCH, total));
        __syncthreads();
    }

    for (int b = 0; b < nbatch; b++) {
        const int stage = b & 1;
        const int cnt = min(FWD_BATCH, total - b * FWD_BATCH);
        // all pixels of the tile saturated -> nothing left to do (block-uniform)
        if (__syncthreads_and(done)) break;

        // (A) start the copies of batch b+1 (its ids were stored one iteration ago)
        if (b + 1 < nbatch) {
            fwd_issue_batch<NQ, VEC, MD, COLOR, TMA>(sm, stage ^ 1, (b + 1) & 1, min(FWD_BATCH, total - (b + 1) * FWD_BATCH),
                                              K, geo, features, mask, depths);
            if constexpr (!TMA) cp_async_commit();
        }
        // (B) ids of batch b+2 into a register
        uint32_t next_id = 0;
        const int rem2 = total - (b + 2) * FWD_BATCH;
        const bool have_next_id = (b + 2 < nbatch) && tid < min(FWD_BATCH, rem2);
        if (have_next_id) next_id = point_list[range.x + (b + 2) * FWD_BATCH + tid];

        // (C) blend batch b
        if (!__all_sync(0xffffffffu, done)) {
            // four splats at a time: independent `power` tests (ILP, one vote per four), accepted ones taken in order.
            // Records beyond the tile's list are sentinels (accept_threshold = +inf): no bounds checks needed.
            const float4* gp = &sm.geo[stage][0][0];
            for (int j0 = 0; j0 < cnt; j0 += 4, gp += 8) {
                float pw[4], op[4];
                bool cd[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 g0 = gp[2 * i];
                    const float4 g1 = gp[2 * i + 1];
                    const float dx = g0.x - pixx, dy = g0.y - pixy;
                    pw[i] = -0.5f * (g0.z * dx * dx + g1.x * dy * dy) - g0.w * dx * dy;
                    cd[i] = !(pw[i] > 0.0f) && (pw[i] >= g1.z);
                    op[i] = g1.y;
                }
                const bool anyc = (cd[0] || cd[1] || cd[2] || cd[3]) && !done;
                if (!__any_sync(0xffffffffu, anyc)) continue;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (cd[i] && !done) {
                        const int j = j0 + i;
                        const float alpha = fminf(0.99f, op[i] * expf(pw[i]));
                        if (!(alpha < 1.0f / 255.0f)) {
                            const float test_T = T * (1 - alpha);
                            if (test_T < 0.0001f) {
                                done = true;
                            } else {
                                if (COLOR) {
#pragma unroll
                                    for (int q = 0; q < NQ; q++) {
                                        const float4 f = sm.feat[stage][j][q];
                                        C[4 * q + 0] += f.x * alpha * T;
                                        C[4 * q + 1] += f.y * alpha * T;
                                        C[4 * q + 2] += f.z * alpha * T;
                                        C[4 * q + 3] += f.w * alpha * T;
                                    }
                                }
                                if (MD) {
                                    Macc += sm.maskv[stage][j] * alpha * T;
                                    Dacc += sm.depthv[stage][j] * alpha * T;
                                }
                                T = test_T;
                                last_contributor = (uint32_t)(b * FWD_BATCH + j + 1);
                            }
                        }
                    }
                }
                if (__all_sync(0xffffffffu, done)) break;   // only reached when some pixel was a candidate
            }
        }

        // (D) publish ids(b+2); wait for batch b+1
        if (have_next_id) sm.ids[b & 1][tid] = next_id;
        if constexpr (TMA) {
            // stage s is filled by batches b = s, s + 2, ...: its (b >> 1)-th fill completes phase parity (b >> 1) & 1
            if (b + 1 < nbatch) mbarrier_wait_parity(&sm.bar[stage ^ 1], (uint32_t)(((b + 1) >> 1) & 1));
        } else {
            cp_async_wait_all();
        }
        if (b + 1 < nbatch) fwd_pad_batch<NQ>(sm, stage ^ 1, min(FWD_BATCH, total - (b + 1) * FWD_BATCH));
        __syncthreads();
    }

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const size_t plane = (size_t)H * W;
        if (COLOR) {
#pragma unroll
            for (int k = 0; k < 4 * NQ; k++)
                if (k < K) out_color[(size_t)k * plane + pix_id] = C[k] + T * bg[k];
        }
        if (MD) {
            out_mask[pix_id] = Macc;
            if (out_depth != nullptr) out_depth[pix_id] = Dacc;
        }
    }
}

#define SAGARS_FWD_PARAMS                                                                                        \
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int K,                 \
    const float* __restrict__ geo, const float* __restrict__ features, const float* __restrict__ mask,              \
    const float* __restrict__ depths, const float* __restrict__ bg, float* __restrict__ final_T,                    \
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_mask,                  \
    float* __restrict__ out_depth
#define SAGARS_FWD_ARGS ranges, point_list, W, H, K, geo, features, mask, depths, bg, final_T, n_contrib, out_color, out_mask, out_depth

// cp.async (LDGSTS) staging -- the default
template <int NQ, bool VEC, bool MD, bool COLOR>
__global__ void __launch_bounds__(TILE_PIX) render_forward_kernel(SAGARS_FWD_PARAMS)
{
    render_forward_body<NQ, VEC, MD, COLOR, false>(SAGARS_FWD_ARGS);
}
// bulk-copy (TMA unit) staging completing on mbarriers -- SAGARS_FLAG_STAGE_TMA
template <int NQ, bool VEC, bool MD, bool COLOR>
__global__ void __launch_bounds__(TILE_PIX) render_forward_tma_kernel(SAGARS_FWD_PARAMS)
{
    render_forward_body<NQ, VEC, MD, COLOR, true>(SAGARS_FWD_ARGS);
}
#undef SAGARS_FWD_PARAMS
#undef SAGARS_FWD_ARGS

template <int NQ, bool VEC, bool MD, bool COLOR, bool TMA>
static int launch_fwd_tt(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                         const uint32_t* point_list, const float* features, cudaStream_t s, bool debug)
{
    auto kern = TMA ? render_forward_tma_kernel<NQ, VEC, MD, COLOR> : render_forward_kernel<NQ, VEC, MD, COLOR>;
    const size_t smem = sizeof(typename FwdSmemSel<NQ, TMA>::type);
    {   // opt in to the dynamic shared-memory size once per device (not on every launch: the call takes the context lock)
        static uint64_t done_mask = 0;
        int dev = 0;
        SAGARS_CUDA(cudaGetDevice(&dev));
        if (!((done_mask >> (dev & 63)) & 1ull)) {
            SAGARS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            done_mask |= 1ull << (dev & 63);
        }
    }
    dim3 grid(d.tiles_x, d.tiles_y);
    kern<<<grid, TILE_PIX, smem, s>>>(im.ranges, point_list, d.W, d.H, d.C, g.geo, features, a.mask, g.depths, a.background,
                                      im.final_T, im.n_contrib, a.out_color, a.out_mask, a.out_depth);
    SAGARS_LAUNCH_CHECK(s, debug);
    return SAGARS_OK;
}

template <int NQ, bool VEC, bool MD, bool COLOR>
static int launch_fwd_t(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                        const uint32_t* point_list, const float* features, cudaStream_t s, bool debug)
{
    if (a.flags & SAGARS_FLAG_STAGE_TMA) return launch_fwd_tt<NQ, VEC, MD, COLOR, true>(a, d, g, im, point_list, features, s, debug);
    return launch_fwd_tt<NQ, VEC, MD, COLOR, false>(a, d, g, im, point_list, features, s, debug);
}

int launch_render_forward(const sagars_forward_args& a, const Dims& d, GeomView g, ImageView im,
                          const uint32_t* point_list, cudaStream_t s, bool debug)
{
    const bool md = (a.flags & SAGARS_FLAG_MASK_DEPTH) != 0;
    const bool mask_only = (a.flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const float* features = a.colors_precomp != nullptr ? a.colors_precomp : g.rgb;
    const int K = d.C;
    if (mask_only) return launch_fwd_t<1, false, true, false>(a, d, g, im, point_list, features, s, debug);
    // colour-only rendering with the channel contraction on the tensor cores.  K = 32 (SAGA's feature rendering): the
    // warp-per-block mma.sync kernel (render_forward_warp.cu), or with SAGARS_FLAG_FWD_TILE the tile-per-CTA tcgen05 / TMEM
    // kernel (render_forward_tc.cu); SAGARS_FLAG_FWD_WARP_ANY extends the warp kernel to every channel count.  Everything
    // else (DEPTH, other channel counts, SAGARS_FLAG_NO_TENSOR_CORES) is the fp32 SIMT kernel below, whose colours are
    // bit-identical to the reference's.
    if (!md && !(a.flags & SAGARS_FLAG_NO_TENSOR_CORES)) {
        if (K == 32 && a.colors_precomp != nullptr) {
            if (a.flags & SAGARS_FLAG_FWD_TILE) return launch_render_forward_tc(a, d, g, im, point_list, s, debug);
            return launch_render_forward_warp(a, d, g, im, point_list, s, debug);
        }
        if (a.flags & SAGARS_FLAG_FWD_WARP_ANY) return launch_render_forward_warp(a, d, g, im, point_list, s, debug);
    }
    const bool vec = (K % 4) == 0;
    const int nq = (K + 3) / 4;
#define SAGARS_FWD_CASE(NQ_)                                                                               \
    if (nq <= NQ_) {                                                                                       \
        if (md) return vec ? launch_fwd_t<NQ_, true, true, true>(a, d, g, im, point_list, features, s, debug) \
                           : launch_fwd_t<NQ_, false, true, true>(a, d, g, im, point_list, features, s, debug); \
        return vec ? launch_fwd_t<NQ_, true, false, true>(a, d, g, im, point_list, features, s, debug)       \
                   : launch_fwd_t<NQ_, false, false, true>(a, d, g, im, point_list, features, s, debug);     \
    }
    SAGARS_FWD_CASE(1)
    SAGARS_FWD_CASE(2)
    SAGARS_FWD_CASE(4)
    SAGARS_FWD_CASE(8)
    SAGARS_FWD_CASE(16)
#undef SAGARS_FWD_CASE
    set_error("unsupported channel count %d (max %d)", K, SAGARS_MAX_CHANNELS);
    return SAGARS_EINVAL;
}

}  // namespace sagars
