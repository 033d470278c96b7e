// knn_kernels.cuh -- the device code and the scratch layout of knn.cu (see there); free of host-side runtime calls so that the
// CPU suite can run it under tests/cuda_emu/.
#pragma once
#include "common.cuh"
#include <cfloat>

namespace sagars {


constexpr int KNN_L1 = 128;    // points per level-1 box
constexpr int KNN_FAN = 32;    // level-1 boxes per level-2 box

struct KnnBox { float lo[3], hi[3]; };   // 24 bytes

struct KnnTemp {
    int* bbox;             // 6 ordered-int encoded floats: min xyz, max xyz
    uint64_t* keys_a;      // sorted (Morton code, point index) pairs land here
    uint64_t* keys_b;
    uint32_t* vals_a;
    uint32_t* vals_b;
    void* sort_temp;
    float4* sorted_pts;    // (x, y, z, original index as bits) in Morton order
    KnnBox* box1;          // ceil(n / 128)
    KnnBox* box2;          // ceil(nbox1 / 32)
};

static size_t knn_temp_layout(size_t n, KnnTemp* t, char* base)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return base ? base + at : (char*)nullptr; };
    const size_t nb1 = (n + KNN_L1 - 1) / KNN_L1, nb2 = (nb1 + KNN_FAN - 1) / KNN_FAN;
    char* p;
    p = take(64);      if (t) t->bbox = (int*)p;
    p = take(n * 8);   if (t) t->keys_a = (uint64_t*)p;
    p = take(n * 8);   if (t) t->keys_b = (uint64_t*)p;
    p = take(n * 4);   if (t) t->vals_a = (uint32_t*)p;
    p = take(n * 4);   if (t) t->vals_b = (uint32_t*)p;
    p = take(sort_temp_bytes(n)); if (t) t->sort_temp = (void*)p;
    p = take(n * 16);  if (t) t->sorted_pts = (float4*)p;
    p = take((nb1 + 1) * sizeof(KnnBox)); if (t) t->box1 = (KnnBox*)p;
    p = take((nb2 + 1) * sizeof(KnnBox)); if (t) t->box2 = (KnnBox*)p;
    return o + 256;
}


// monotone float <-> int encoding, so atomicMin / atomicMax on ints order floats
__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void knn_bbox_init_kernel(int* bbox)
{
    if (threadIdx.x < 3) bbox[threadIdx.x] = f2ord(FLT_MAX);
    else if (threadIdx.x < 6) bbox[threadIdx.x] = f2ord(-FLT_MAX);
}

__global__ void __launch_bounds__(256)
knn_bbox_kernel(int n, const float* __restrict__ pts, int* __restrict__ bbox)
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = pts[3 * (size_t)i + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicMin(&bbox[c], f2ord(mn[c]));
            atomicMax(&bbox[3 + c], f2ord(mx[c]));
        }
    }
}

__device__ __forceinline__ uint32_t knn_spread3(uint32_t x)   // 10 bits -> every third bit
{
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

__global__ void __launch_bounds__(256)
knn_morton_kernel(int n, const float* __restrict__ pts, const int* __restrict__ bbox, uint64_t* __restrict__ keys,
                  uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float lo = ord2f(bbox[c]), hi = ord2f(bbox[3 + c]);
        const float ext = hi - lo;
        const float u = ext > 0.f ? (pts[3 * (size_t)i + c] - lo) / ext : 0.f;
        const uint32_t q = (uint32_t)fminf(fmaxf(u * 1023.f, 0.f), 1023.f);
        code |= knn_spread3(q) << c;
    }
    keys[i] = (uint64_t)code;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
knn_gather_kernel(int n, const uint32_t* __restrict__ vals, const float* __restrict__ pts, float4* __restrict__ sorted_pts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = vals[i];
    sorted_pts[i] = make_float4(pts[3 * (size_t)j], pts[3 * (size_t)j + 1], pts[3 * (size_t)j + 2], __uint_as_float(j));
}

// level-1 boxes: one warp per 128 consecutive points
__global__ void __launch_bounds__(256)
knn_box1_kernel(int n, const float4* __restrict__ sorted_pts, KnnBox* __restrict__ box1, int nb1)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= nb1) return;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = b * KNN_L1 + lane; i < min(n, (b + 1) * KNN_L1); i += 32) {
        const float4 p = sorted_pts[i];
        lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
        lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
        lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if (lane == 0) {
        KnnBox bx;
#pragma unroll
        for (int c = 0; c < 3; c++) { bx.lo[c] = lo[c]; bx.hi[c] = hi[c]; }
        box1[b] = bx;
    }
}

// level-2 boxes: one thread per 32 level-1 boxes
__global__ void __launch_bounds__(128)
knn_box2_kernel(const KnnBox* __restrict__ box1, int nb1, KnnBox* __restrict__ box2, int nb2)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb2) return;
    KnnBox bx;
#pragma unroll
    for (int c = 0; c < 3; c++) { bx.lo[c] = FLT_MAX; bx.hi[c] = -FLT_MAX; }
    for (int i = b * KNN_FAN; i < min(nb1, (b + 1) * KNN_FAN); i++) {
        const KnnBox a = box1[i];
#pragma unroll
        for (int c = 0; c < 3; c++) { bx.lo[c] = fminf(bx.lo[c], a.lo[c]); bx.hi[c] = fmaxf(bx.hi[c], a.hi[c]); }
    }
    box2[b] = bx;
}

// squared distance from a point to a box (0 inside); a lower bound for every point of the box
__device__ __forceinline__ float knn_box_dist2(const KnnBox& b, float x, float y, float z)
{
    const float dx = fmaxf(fmaxf(b.lo[0] - x, x - b.hi[0]), 0.f);
    const float dy = fmaxf(fmaxf(b.lo[1] - y, y - b.hi[1]), 0.f);
    const float dz = fmaxf(fmaxf(b.lo[2] - z, z - b.hi[2]), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// One thread per query.  SELF: the query set is the point set (thread t takes the t-th point in Morton order); with
// EXCL a point is never its own neighbour (simple_knn).  Otherwise thread t takes external query t.
template <int K, bool SELF, bool EXCL>
__global__ void __launch_bounds__(128)
knn_search_kernel(int n, int nq, const float* __restrict__ queries, const float4* __restrict__ sorted_pts,
                  const KnnBox* __restrict__ box1, int nb1, const KnnBox* __restrict__ box2, int nb2,
                  int k_out, long long* __restrict__ idx_out, float* __restrict__ dist_out, float* __restrict__ mean_out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    float qx, qy, qz;
    uint32_t qid;            // output row (original index of the query)
    int own1 = -1;           // the query's own level-1 box (SELF)
    if (SELF) {
        const float4 p = sorted_pts[t];
        qx = p.x; qy = p.y; qz = p.z; qid = __float_as_uint(p.w);
        own1 = t / KNN_L1;
    } else {
        qx = queries[3 * (size_t)t]; qy = queries[3 * (size_t)t + 1]; qz = queries[3 * (size_t)t + 2];
        qid = (uint32_t)t;
    }

    float bd[K];
    uint32_t bi[K];
#pragma unroll
    for (int j = 0; j < K; j++) { bd[j] = FLT_MAX; bi[j] = 0xffffffffu; }

    auto scan_box1 = [&](int b) {
        const int e = min(n, (b + 1) * KNN_L1);
        for (int i = b * KNN_L1; i < e; i++) {
            const float4 p = sorted_pts[i];
            if (EXCL && i == t) continue;
            // same expression as simple_knn.cu:135-137 (d = candidate - query)
            const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            float d = dx * dx + dy * dy + dz * dz;
            if (!(d < bd[K - 1])) continue;
            uint32_t id = __float_as_uint(p.w);
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (bd[j] > d) {
                    const float td = bd[j]; bd[j] = d; d = td;
                    const uint32_t ti = bi[j]; bi[j] = id; id = ti;
                }
            }
        }
    };

    if (SELF) scan_box1(own1);
    for (int b2 = 0; b2 < nb2; b2++) {
        if (knn_box_dist2(box2[b2], qx, qy, qz) > bd[K - 1]) continue;
        const int e1 = min(nb1, (b2 + 1) * KNN_FAN);
        for (int b1 = b2 * KNN_FAN; b1 < e1; b1++) {
            if (b1 == own1) continue;
            if (knn_box_dist2(box1[b1], qx, qy, qz) > bd[K - 1]) continue;
            scan_box1(b1);
        }
    }

    if (idx_out || dist_out) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (j < k_out) {
                if (idx_out) idx_out[(size_t)qid * k_out + j] = (bi[j] == 0xffffffffu) ? -1ll : (long long)bi[j];
                if (dist_out) dist_out[(size_t)qid * k_out + j] = bd[j];
            }
        }
    }
    if (mean_out) {
        // simple_knn.cu:183: (best[0] + best[1] + best[2]) / 3.0f, generalised to k_out terms in ascending order
        float sacc = 0.f;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (j < k_out) sacc += bd[j];
        mean_out[qid] = sacc / (float)k_out;
    }
}

}  // namespace sagars
