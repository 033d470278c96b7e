// api.cu -- the C ABI (include/sagars.h): host orchestration of the forward / backward stages.
//
// Replaces the host side of the reference: CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (CF cuda_rasterizer/rasterizer_impl.cu:141-153,198-434) and the scratch carving of
// GeometryState / ImageState / BinningState (rasterizer_impl.cu:155-194).
#include "common.cuh"
#include <sched.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <atomic>
#include <mutex>
#include <vector>

namespace sagars {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs backward on its own thread

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
    const char* base = strrchr(file, '/');
    set_error("[CUDA ERROR] %s (%s) at %s:%d", cudaGetErrorString(e), what, base ? base + 1 : file, line);
    return SAGARS_ECUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional per-stage device timing (CUDA events on the caller's stream), for bench.py's roofline ----
enum Stage { ST_PREPROCESS = 0, ST_SCAN, ST_DUPLICATE, ST_SORT, ST_RANGES, ST_RENDER_FWD, ST_RENDER_BWD, ST_GEOM_BWD, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {"preprocess", "scan_block_sums", "duplicate_keys", "radix_sort", "tile_ranges",
                                            "render_forward", "render_backward", "geom_backward"};
struct ProfRec { int stage; cudaEvent_t a, b; };
static std::atomic<int> g_profile{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static double g_prof_ms[ST_COUNT];
static int64_t g_prof_n[ST_COUNT];

static cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    int stage; cudaStream_t s; cudaEvent_t a = nullptr, b = nullptr; bool on;
    ProfScope(int stage_, cudaStream_t s_) : stage(stage_), s(s_), on(g_profile.load() != 0)
    {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        a = prof_event(); b = prof_event();
        cudaEventRecord(a, s);
    }
    ~ProfScope()
    {
        if (!on) return;
        cudaEventRecord(b, s);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_recs.push_back({stage, a, b});
    }
};

// host-side trace of the forward's phases (SAGARS_TRACE=1): where does wall time go between launches?
static bool trace_on()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("SAGARS_TRACE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
static double now_us()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

// per-thread pinned landing buffer and event for the forward's single host read-back
static uint32_t* pinned_status()
{
    thread_local uint32_t* p = nullptr;
    if (!p && cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault) != cudaSuccess) p = nullptr;
    return p;
}
static cudaEvent_t sync_event()
{
    thread_local cudaEvent_t ev = nullptr;
    thread_local int ev_dev = -1;
    int dev = -1;
    cudaGetDevice(&dev);
    if (ev && ev_dev != dev) { cudaEventDestroy(ev); ev = nullptr; }
    if (!ev) {
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) ev = nullptr;
        ev_dev = dev;
    }
    return ev;
}
// SAGARS_SYNC=block: cudaStreamSynchronize; default: poll the event
static int sync_mode()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("SAGARS_SYNC"); v = (e && e[0] == 'b') ? 1 : 0; }
    return v;
}

// same rule as the reference's getHigherMsb (CF rasterizer_impl.cu:35-50): bits needed for tile ids
static uint32_t higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step;
        else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

static int fill_dims(Dims& d, int P, int D, int M, int C, int W, int H, float tfx, float tfy, float mod)
{
    if (P < 0 || W <= 0 || H <= 0) {
        set_error("bad dimensions P=%d W=%d H=%d", P, W, H);
        return SAGARS_EINVAL;
    }
    if (C < 1 || C > SAGARS_MAX_CHANNELS) {
        set_error("unsupported channel count %d (1..%d)", C, SAGARS_MAX_CHANNELS);
        return SAGARS_EINVAL;
    }
    d.P = P; d.D = D; d.M = M; d.C = C; d.W = W; d.H = H;
    d.tiles_x = (W + TILE_X - 1) / TILE_X;
    d.tiles_y = (H + TILE_Y - 1) / TILE_Y;
    d.tan_fovx = tfx; d.tan_fovy = tfy;
    d.focal_y = H / (2.0f * tfy);
    d.focal_x = W / (2.0f * tfx);
    d.scale_modifier = mod;
    return SAGARS_OK;
}

}  // namespace sagars

using namespace sagars;

extern "C" {

int sagars_abi_version(void) { return SAGARS_ABI_VERSION; }
size_t sagars_sizeof_forward_args(void) { return sizeof(sagars_forward_args); }
size_t sagars_sizeof_backward_args(void) { return sizeof(sagars_backward_args); }

void sagars_profile_enable(int on) { g_profile.store(on ? 1 : 0); }
int sagars_profile_num_stages(void) { return ST_COUNT; }
const char* sagars_profile_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
int sagars_profile_read(double* ms_out, int64_t* count_out, int reset)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double gap_ms[ST_COUNT] = {0.0};
    for (size_t i = 0; i < g_prof_recs.size(); i++) {
        auto& r = g_prof_recs[i];
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_prof_ms[r.stage] += ms;
            g_prof_n[r.stage] += 1;
        }
        float gap = 0.f;   // stream time between the previous stage's end and this stage's start (SAGARS_TRACE only)
        if (trace_on() && i > 0 && cudaEventElapsedTime(&gap, g_prof_recs[i - 1].b, r.a) == cudaSuccess) gap_ms[r.stage] += gap;
    }
    if (trace_on() && !g_prof_recs.empty()) {
        for (int i = 0; i < ST_COUNT; i++)
            fprintf(stderr, "[sagars] stream gap before %-16s total %9.3f ms over %zu records\n", kStageNames[i], gap_ms[i], g_prof_recs.size());
    }
    for (auto& r : g_prof_recs) {
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    for (int i = 0; i < ST_COUNT; i++) {
        if (ms_out) ms_out[i] = g_prof_ms[i];
        if (count_out) count_out[i] = g_prof_n[i];
        if (reset) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
    }
    return SAGARS_OK;
}
const char* sagars_arch(void) { return "sm_100a"; }
const char* sagars_last_error(void) { return g_err; }
int64_t sagars_launch_count(void) { return g_launches.load(); }
void sagars_reset_launch_count(void) { g_launches.store(0); }

size_t sagars_geom_bytes(int32_t P) { return geom_layout((size_t)(P < 0 ? 0 : P)).total; }
size_t sagars_image_bytes(int32_t W, int32_t H) { return image_layout(W, H).total; }
size_t sagars_binning_bytes(int32_t R) { return binning_total((size_t)(R < 0 ? 0 : R)); }
size_t sagars_grad_scratch_bytes(int32_t P) { return grad_scratch_bytes((size_t)(P < 0 ? 0 : P)); }
size_t sagars_sort_temp_bytes(int32_t n)
{
    const size_t m = (size_t)(n < 0 ? 0 : n);
    return align_up(m * 8) + align_up(m * 4) + sort_temp_bytes(m) + 256;
}

int sagars_get_geom_layout(int32_t P, sagars_geom_layout* out)
{
    if (!out || P < 0) { set_error("bad argument"); return SAGARS_EINVAL; }
    *out = geom_layout((size_t)P);
    return SAGARS_OK;
}
int sagars_get_image_layout(int32_t W, int32_t H, sagars_image_layout* out)
{
    if (!out || W <= 0 || H <= 0) { set_error("bad argument"); return SAGARS_EINVAL; }
    *out = image_layout(W, H);
    return SAGARS_OK;
}
int sagars_get_binning_layout(int32_t R, sagars_binning_layout* out)
{
    if (!out || R < 0) { set_error("bad argument"); return SAGARS_EINVAL; }
    *out = binning_layout((size_t)R);
    out->total = binning_total((size_t)R);
    return SAGARS_OK;
}

int sagars_forward(const sagars_forward_args* a,
                   sagars_alloc_fn geom_alloc, void* geom_user,
                   sagars_alloc_fn binning_alloc, void* binning_user,
                   sagars_alloc_fn image_alloc, void* image_user,
                   int32_t* num_rendered, void* stream_)
{
    g_err[0] = 0;
    if (!a || !geom_alloc || !binning_alloc || !image_alloc || !num_rendered) {
        set_error("sagars_forward: NULL argument");
        return SAGARS_EINVAL;
    }
    cudaStream_t s = (cudaStream_t)stream_;
    const bool debug = (a->flags & SAGARS_FLAG_DEBUG) != 0;
    const bool mask_only = (a->flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const bool md = (a->flags & (SAGARS_FLAG_MASK_DEPTH | SAGARS_FLAG_MASK_ONLY)) != 0;
    *num_rendered = 0;

    Dims d;
    int rc = fill_dims(d, a->P, a->D, a->M, a->num_channels, a->width, a->height, a->tan_fovx, a->tan_fovy, a->scale_modifier);
    if (rc) return rc;
    if (a->P == 0) return SAGARS_OK;   // nothing is launched; caller pre-zeroes outputs (CF rasterize_points.cu:81)
    if (!a->means3D || !a->opacities || !a->viewmatrix || !a->projmatrix || !a->radii || !a->background ||
        (!mask_only && !a->out_color)) {
        set_error("sagars_forward: missing required pointer");
        return SAGARS_EINVAL;
    }
    if (!mask_only && a->colors_precomp == nullptr) {
        // the reference's message for the non-RGB build (CF rasterizer_impl.cu:242-245)
        if (a->num_channels != 3) {
            set_error("For non-RGB, provide precomputed Gaussian colors!");
            return SAGARS_ENOCOLOR;
        }
        if (!a->shs || !a->cam_pos || a->M <= 0) {
            set_error("sagars_forward: neither colors_precomp nor shs given");
            return SAGARS_EINVAL;
        }
    }
    if (a->cov3D_precomp == nullptr && (!a->scales || !a->rotations)) {
        set_error("sagars_forward: neither cov3D_precomp nor scales/rotations given");
        return SAGARS_EINVAL;
    }
    if (md && (!a->mask || !a->out_mask)) {
        set_error("sagars_forward: mask / out_mask required for the depth variant");
        return SAGARS_EINVAL;
    }
    const bool tr = trace_on();
    double t0 = tr ? now_us() : 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    SAGARS_CUDA(cudaSetDevice(a->device));

    void* geom_mem = geom_alloc(geom_user, geom_layout(d.P).total);
    void* img_mem = image_alloc(image_user, image_layout(d.W, d.H).total);
    if (!geom_mem || !img_mem) { set_error("allocator callback returned NULL"); return SAGARS_EALLOC; }
    GeomView g = geom_view(geom_mem, d.P);
    ImageView im = image_view(img_mem, d.W, d.H);

    if (tr) t1 = now_us();
    SAGARS_CUDA(cudaMemsetAsync(g.status, 0, 64, s));
    { ProfScope ps(ST_PREPROCESS, s); rc = launch_preprocess(*a, d, g, s, debug); }
    if (rc) return rc;
    { ProfScope ps(ST_SCAN, s); rc = launch_scan_block_sums(d, g, s, debug); }
    if (rc) return rc;

    // The one host read-back of the forward pass: R sizes the binning buffer (CF rasterizer_impl.cu:280-285).
    // The copy lands in pinned memory and the wait polls an event.  With a capacity hint the remaining stages are
    // queued BEFORE the wait (they read R from device memory), so the GPU keeps working while the host wakes up.
    if (tr) t2 = now_us();
    uint32_t* pin = pinned_status();
    cudaEvent_t ev = sync_event();
    if (!pin || !ev) { set_error("cannot allocate the pinned status buffer"); return SAGARS_ECUDA; }
    SAGARS_CUDA(cudaMemcpyAsync(pin, g.status, 8, cudaMemcpyDeviceToHost, s));
    SAGARS_CUDA(cudaEventRecord(ev, s));
    auto wait_count = [&](int* R_out) -> int {
        if (sync_mode() == 1) {
            SAGARS_CUDA(cudaStreamSynchronize(s));
        } else {
            // poll briefly (the count is usually there within a few microseconds), then yield the core between polls: one
            // process per GPU plus data-loader workers should not pin a core per rank on this wait
            cudaError_t q;
            unsigned spins = 0;
            while ((q = cudaEventQuery(ev)) == cudaErrorNotReady) {
                if (++spins > 2000) sched_yield();
                else __builtin_ia32_pause();
            }
            SAGARS_CUDA(q);
        }
        if (pin[0] != 0) {
            set_error("Point is filtered although prefiltered is set. This shouldn't happen!");
            return SAGARS_EPREFILTER;
        }
        *R_out = (int)pin[1];
        return SAGARS_OK;
    };

    const int num_tiles = d.tiles_x * d.tiles_y;
    const int end_bit = 32 + (int)higher_msb((uint32_t)num_tiles);
    const bool use_cub = (a->flags & SAGARS_FLAG_CUB_SORT) != 0;
    // depth-first binning: the Gaussians' depth order needs nothing but P, so it is queued BEFORE the host waits for the count
    const bool depth_first = (a->flags & SAGARS_FLAG_DEPTH_FIRST) != 0 && !use_cub && !(a->flags & SAGARS_FLAG_TILE_SORT);
    if (depth_first) {
        ProfScope ps(ST_SORT, s);
        rc = launch_depth_order(d, g, s, debug);
        if (rc) return rc;
    }
    bool speculative = a->binning_capacity_hint > 0 && !use_cub && !debug;
    int R = 0, cap = 0;
    if (speculative) {
        cap = a->binning_capacity_hint;
    } else {
        rc = wait_count(&R);
        if (rc) return rc;
        cap = R;
    }
    if (tr) t3 = now_us();
    for (;;) {
        void* bin_mem = binning_alloc(binning_user, binning_total((size_t)cap));
        if (!bin_mem) { set_error("allocator callback returned NULL"); return SAGARS_EALLOC; }
        BinningView bv = binning_view(bin_mem, (size_t)cap);
        const uint32_t* n_dev = speculative ? g.status + 1 : nullptr;   // exact layout: the host-side count is the count
        // the tile sort's queue of long segments lives in the radix sort's scratch (>= 1 MB); absurdly many tiles: radix path
        const bool tile_sort = (a->flags & SAGARS_FLAG_TILE_SORT) != 0 && !use_cub &&
                               tile_sort_queue_bytes(num_tiles) <= sort_temp_bytes((size_t)cap);
        if (depth_first) {
            if (cap > 0) {
                // tile-id keys live in the (unused) 64-bit alternate key array; values ping-pong between point_list and vals_alt
                uint32_t* tk_a = reinterpret_cast<uint32_t*>(bv.keys_alt);
                uint32_t* tk_b = tk_a + cap;
                const int tile_bits = end_bit - 32;
                const bool start_b = (sort_num_passes(tile_bits) & 1) != 0;
                { ProfScope ps(ST_DUPLICATE, s); rc = launch_emit_sorted(d, g, a->radii, start_b ? tk_b : tk_a, start_b ? bv.vals_alt : bv.point_list, n_dev, cap, s, debug); }
                if (rc) return rc;
                { ProfScope ps(ST_SORT, s); rc = launch_sort_pairs32(n_dev, cap, tile_bits, tk_a, bv.point_list, tk_b, bv.vals_alt, bv.sort_temp, s, debug); }
                if (rc) return rc;
                { ProfScope ps(ST_RANGES, s); rc = launch_finalize_bins(n_dev, cap, num_tiles, tk_a, bv.point_list, g.depths, bv.point_list_keys, im.ranges, s, debug); }
                if (rc) return rc;
            } else {
                SAGARS_CUDA(cudaMemsetAsync(im.ranges, 0, (size_t)num_tiles * sizeof(uint2), s));
            }
        } else if (cap > 0 && tile_sort) {
            // no global sort: per-tile counts -> scan -> scatter (the tile ranges fall out), then one CTA per tile sorts its
            // own segment (tile_sort.cu).  keys_alt holds the unsorted (depth bits, id) pairs.
            { ProfScope ps(ST_DUPLICATE, s); rc = launch_tile_bin(d, g, a->radii, bv.keys_alt, im.ranges, (uint32_t*)bv.sort_temp, n_dev, cap, s, debug); }
            if (rc) return rc;
            { ProfScope ps(ST_SORT, s); rc = launch_tile_sort(num_tiles, im.ranges, bv.keys_alt, bv.point_list, bv.point_list_keys, (uint32_t*)bv.sort_temp, n_dev, cap, s, debug); }
            if (rc) return rc;
        } else {
            if (cap > 0) {
                // own sort: emit into the buffer from which an npass-long ping-pong ends in the final arrays
                const bool start_alt = !use_cub && (sort_num_passes(end_bit) & 1);
                uint64_t* k0 = start_alt ? bv.keys_alt : bv.point_list_keys;
                uint32_t* v0 = start_alt ? bv.vals_alt : bv.point_list;
                { ProfScope ps(ST_DUPLICATE, s); rc = launch_duplicate(d, g, a->radii, k0, v0, n_dev, cap, s, debug); }
                if (rc) return rc;
                bool in_a = true;
                {
                    ProfScope ps(ST_SORT, s);
                    rc = launch_sort_pairs(n_dev, cap, end_bit, bv.point_list_keys, bv.point_list, bv.keys_alt, bv.vals_alt,
                                           bv.sort_temp, sort_temp_bytes((size_t)cap), use_cub, &in_a, s, debug);
                }
                if (rc) return rc;
                if (!in_a) {
                    SAGARS_CUDA(cudaMemcpyAsync(bv.point_list_keys, bv.keys_alt, (size_t)cap * 8, cudaMemcpyDeviceToDevice, s));
                    SAGARS_CUDA(cudaMemcpyAsync(bv.point_list, bv.vals_alt, (size_t)cap * 4, cudaMemcpyDeviceToDevice, s));
                }
            }
            { ProfScope ps(ST_RANGES, s); rc = launch_tile_ranges(n_dev, cap, num_tiles, bv.point_list_keys, im.ranges, s, debug); }
            if (rc) return rc;
        }
        // the blend is the first stage that reads the colours / features: an optional event gates it (include/sagars.h)
        if (a->blend_wait_event) SAGARS_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)a->blend_wait_event, 0));
        { ProfScope ps(ST_RENDER_FWD, s); rc = launch_render_forward(*a, d, g, im, bv.point_list, s, debug); }
        if (rc) return rc;
        if (!speculative) break;
        rc = wait_count(&R);
        if (rc) return rc;
        if (R <= cap) break;
        // the hint was too small: the queued stages saw the overflow and did nothing; lay the buffer out exactly
        speculative = false;
        cap = R;
    }
    if (tr) t4 = now_us();
    *num_rendered = R;
    if (a->binning_capacity_out) *a->binning_capacity_out = cap;
    if (tr) {
        t5 = now_us();
        fprintf(stderr, "[sagars trace] fwd host us: alloc(geom,img)=%.0f launch(pre,scan)=%.0f wait-before=%.0f rest(+wait)=%.0f "
                        "total=%.0f returned-at=%.0f\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t0, t5);
    }
    return SAGARS_OK;
}

int sagars_backward(const sagars_backward_args* a, void* stream_)
{
    g_err[0] = 0;
    if (!a) { set_error("sagars_backward: NULL argument"); return SAGARS_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream_;
    const bool debug = (a->flags & SAGARS_FLAG_DEBUG) != 0;
    const bool mask_only = (a->flags & SAGARS_FLAG_MASK_ONLY) != 0;
    const bool md = (a->flags & (SAGARS_FLAG_MASK_DEPTH | SAGARS_FLAG_MASK_ONLY)) != 0;

    Dims d;
    int rc = fill_dims(d, a->P, a->D, a->M, a->num_channels, a->width, a->height, a->tan_fovx, a->tan_fovy, a->scale_modifier);
    if (rc) return rc;
    if (a->P == 0) return SAGARS_OK;
    if (!a->geom_buffer || !a->image_buffer || (a->R > 0 && !a->binning_buffer) || !a->grad_scratch || !a->radii ||
        !a->means3D || !a->viewmatrix || !a->projmatrix || !a->background) {
        set_error("sagars_backward: missing required pointer");
        return SAGARS_EINVAL;
    }
    if (!mask_only && (!a->dL_dout_color || !a->dL_dcolors || !a->dL_dmeans2D || !a->dL_dopacity || !a->dL_dmeans3D ||
                       !a->dL_dcov3D || !a->dL_dscales || !a->dL_drotations)) {
        set_error("sagars_backward: missing gradient pointer");
        return SAGARS_EINVAL;
    }
    if (md && (!a->dL_dout_mask || !a->dL_dmask)) {
        set_error("sagars_backward: dL_dout_mask / dL_dmask required for the depth variant");
        return SAGARS_EINVAL;
    }
    if (a->M > 0 && a->shs != nullptr && !a->dL_dsh) {
        set_error("sagars_backward: dL_dsh required when shs are given");
        return SAGARS_EINVAL;
    }
    SAGARS_CUDA(cudaSetDevice(a->device));

    GeomView g = geom_view(const_cast<void*>(a->geom_buffer), d.P);
    ImageView im = image_view(const_cast<void*>(a->image_buffer), d.W, d.H);
    const uint32_t* point_list = nullptr;
    if (a->R > 0) point_list = binning_view(const_cast<void*>(a->binning_buffer), (size_t)a->R).point_list;
    float* ggrad = (float*)a->grad_scratch;

    const bool tr = trace_on();
    const double t0 = tr ? now_us() : 0;
    SAGARS_CUDA(cudaMemsetAsync(ggrad, 0, (size_t)d.P * GG_STRIDE * sizeof(float), s));
    if (!mask_only) SAGARS_CUDA(cudaMemsetAsync(a->dL_dcolors, 0, (size_t)d.P * d.C * sizeof(float), s));
    const double t1 = tr ? now_us() : 0;
    if (a->R > 0) {
        {
            ProfScope ps(ST_RENDER_BWD, s);
            if (a->flags & SAGARS_FLAG_NO_TENSOR_CORES) rc = launch_render_backward(*a, d, g, im, point_list, ggrad, s, debug);
            else if (a->flags & SAGARS_FLAG_BWD_TILE) rc = launch_render_backward_mma(*a, d, g, im, point_list, ggrad, s, debug);
            else if ((a->flags & SAGARS_FLAG_BWD_TC) && !md && d.C == 32 && a->colors_precomp != nullptr)
                rc = launch_render_backward_tc(*a, d, g, im, point_list, ggrad, s, debug);
            else rc = launch_render_backward_warp(*a, d, g, im, point_list, ggrad, s, debug);
        }
        if (rc) return rc;
    }
    if (tr) fprintf(stderr, "[sagars trace] bwd host us: memsets=%.0f launch(render_backward)=%.0f at=%.0f\n", t1 - t0, now_us() - t1, t0);
    if (mask_only) {
        // mask-only path: the only gradient is dL_dmask (DEPTH __init__.py:280-289)
        SAGARS_CUDA(cudaMemcpy2DAsync(a->dL_dmask, sizeof(float), ggrad + 6, GG_STRIDE * sizeof(float), sizeof(float),
                                      (size_t)d.P, cudaMemcpyDeviceToDevice, s));
        return SAGARS_OK;
    }
    { ProfScope ps(ST_GEOM_BWD, s); rc = launch_geom_backward(*a, d, g, ggrad, s, debug); }
    return rc;
}

int sagars_mark_visible(int32_t device, int32_t P, const float* means3D, const float* viewmatrix,
                        const float* projmatrix, uint8_t* present, void* stream)
{
    g_err[0] = 0;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("sagars_mark_visible: bad argument");
        return SAGARS_EINVAL;
    }
    if (P == 0) return SAGARS_OK;
    SAGARS_CUDA(cudaSetDevice(device));
    return launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, (cudaStream_t)stream);
}

int sagars_sort_pairs(int32_t device, int32_t n, int32_t end_bit, const uint64_t* keys_in, const uint32_t* vals_in,
                      uint64_t* keys_out, uint32_t* vals_out, void* temp, int32_t use_cub, void* stream_)
{
    // stand-alone entry for tests.  `temp` layout: [keys_alt n*8][vals_alt n*4][sort temp]
    g_err[0] = 0;
    if (n < 0 || end_bit < 1 || end_bit > 64 || (n > 0 && (!keys_in || !vals_in || !keys_out || !vals_out || !temp))) {
        set_error("sagars_sort_pairs: bad argument");
        return SAGARS_EINVAL;
    }
    if (n == 0) return SAGARS_OK;
    cudaStream_t s = (cudaStream_t)stream_;
    SAGARS_CUDA(cudaSetDevice(device));
    char* t = (char*)temp;
    uint64_t* keys_alt = (uint64_t*)t;
    uint32_t* vals_alt = (uint32_t*)(t + align_up((size_t)n * 8));
    void* stemp = t + align_up((size_t)n * 8) + align_up((size_t)n * 4);
    const bool start_alt = !use_cub && (sort_num_passes(end_bit) & 1);
    SAGARS_CUDA(cudaMemcpyAsync(start_alt ? keys_alt : keys_out, keys_in, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    SAGARS_CUDA(cudaMemcpyAsync(start_alt ? vals_alt : vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    bool in_a = true;
    int rc = launch_sort_pairs(nullptr, n, end_bit, keys_out, vals_out, keys_alt, vals_alt, stemp, sort_temp_bytes((size_t)n),
                               use_cub != 0, &in_a, s, false);
    if (rc) return rc;
    if (!in_a) {
        SAGARS_CUDA(cudaMemcpyAsync(keys_out, keys_alt, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        SAGARS_CUDA(cudaMemcpyAsync(vals_out, vals_alt, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    }
    return SAGARS_OK;
}

size_t sagars_knn_temp_bytes(int32_t num_points) { return knn_temp_bytes((size_t)(num_points < 0 ? 0 : num_points)); }

int sagars_knn(int32_t device, int32_t num_points, const float* points, int32_t num_queries, const float* queries,
               int32_t K, int32_t exclude_self, int64_t* idx_out, float* dist2_out, float* mean_dist2_out, void* temp,
               void* stream)
{
    g_err[0] = 0;
    if (num_points < 0 || num_queries < 0 || K < 1 || K > 32 || (num_points > 0 && (!points || !temp)) ||
        (exclude_self && queries != nullptr && queries != points)) {
        set_error("sagars_knn: bad argument (1 <= K <= 32; exclude_self needs queries == points)");
        return SAGARS_EINVAL;
    }
    if (queries == nullptr || queries == points) { queries = nullptr; num_queries = num_points; }
    if (num_queries == 0) return SAGARS_OK;
    if (num_points == 0) { set_error("sagars_knn: empty reference cloud"); return SAGARS_EINVAL; }
    SAGARS_CUDA(cudaSetDevice(device));
    return launch_knn(num_points, points, num_queries, queries, K, exclude_self != 0, (long long*)idx_out, dist2_out,
                      mean_dist2_out, temp, (cudaStream_t)stream);
}

int sagars_smooth_forward(int32_t device, int32_t P, int32_t C, int32_t Ks, const float* features, const int64_t* nbr_idx,
                          int32_t normalize_out, float* out, float* mean_norm, void* stream)
{
    g_err[0] = 0;
    if (P < 0 || C < 1 || C > SAGARS_MAX_CHANNELS || Ks < 1 ||
        (P > 0 && (!features || !nbr_idx || !out || (normalize_out && !mean_norm)))) {
        set_error("sagars_smooth_forward: bad argument (1 <= C <= %d, Ks >= 1)", SAGARS_MAX_CHANNELS);
        return SAGARS_EINVAL;
    }
    if (P == 0) return SAGARS_OK;
    SAGARS_CUDA(cudaSetDevice(device));
    return launch_smooth_forward(P, C, Ks, features, (const long long*)nbr_idx, normalize_out, out, mean_norm,
                                 (cudaStream_t)stream);
}

int sagars_smooth_backward(int32_t device, int32_t P, int32_t C, int32_t Ks, const float* features, const int64_t* nbr_idx,
                           int32_t normalize_out, const float* mean_norm, const float* out,
                           const float* dL_dout, float* dL_dn_scratch, float* dL_dfeatures, void* stream)
{
    g_err[0] = 0;
    if (P < 0 || C < 1 || C > SAGARS_MAX_CHANNELS || Ks < 1 ||
        (P > 0 && (!features || !nbr_idx || !dL_dout || !dL_dn_scratch || !dL_dfeatures ||
                   (normalize_out && (!mean_norm || !out))))) {
        set_error("sagars_smooth_backward: bad argument");
        return SAGARS_EINVAL;
    }
    if (P == 0) return SAGARS_OK;
    SAGARS_CUDA(cudaSetDevice(device));
    return launch_smooth_backward(P, C, Ks, features, (const long long*)nbr_idx, normalize_out, mean_norm, out,
                                  dL_dout, dL_dn_scratch, dL_dfeatures, (cudaStream_t)stream);
}

int sagars_sample_rays_forward(int32_t device, int32_t C, int32_t H, int32_t W, int32_t out_h, int32_t out_w, const float* image,
                               const int64_t* ray_index, int32_t num_rays, float* samples, float* norm_sum, void* stream)
{
    g_err[0] = 0;
    if (C < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1 || num_rays < 0 || !image || !norm_sum || (num_rays > 0 && (!ray_index || !samples))) {
        set_error("sagars_sample_rays_forward: bad argument");
        return SAGARS_EINVAL;
    }
    SAGARS_CUDA(cudaSetDevice(device));
    return launch_sample_rays_forward(C, H, W, out_h, out_w, image, (const long long*)ray_index, num_rays, samples, norm_sum, (cudaStream_t)stream);
}

int sagars_sample_rays_backward(int32_t device, int32_t C, int32_t H, int32_t W, int32_t out_h, int32_t out_w, const float* image,
                                const int64_t* ray_index, int32_t num_rays, const float* dL_dsamples, const float* dL_dnorm_mean,
                                float* dL_dimage, void* stream)
{
    g_err[0] = 0;
    if (C < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1 || num_rays < 0 || !image || !dL_dnorm_mean || !dL_dimage ||
        (num_rays > 0 && (!ray_index || !dL_dsamples))) {
        set_error("sagars_sample_rays_backward: bad argument");
        return SAGARS_EINVAL;
    }
    SAGARS_CUDA(cudaSetDevice(device));
    return launch_sample_rays_backward(C, H, W, out_h, out_w, image, (const long long*)ray_index, num_rays, dL_dsamples, dL_dnorm_mean,
                                       dL_dimage, (cudaStream_t)stream);
}

}  // extern "C"
