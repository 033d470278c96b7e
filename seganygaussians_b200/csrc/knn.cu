// knn.cu -- exact K nearest neighbours on a uniform grid.
//
// Serves the two neighbour searches the reference's scripts import from third-party packages that are not installed:
//   * simple_knn._C.distCUDA2(points)       -> mean squared distance to the 3 nearest OTHER points
//       (submodules/simple-knn/simple_knn.cu:146-219: Morton order + box pruning, exact; the result is
//        (best[0] + best[1] + best[2]) / 3.0f with best[] ascending and d2 = d.x*d.x + d.y*d.y + d.z*d.z);
//   * pytorch3d.ops.knn_points(xyz, xyz, K) -> indices / squared distances of the K nearest points, the point itself
//       included (scene/gaussian_model_ff.py:326-331, 345-350: K = 16 for the feature smoothing map).
//
// Method: bounding box -> cell size chosen on the device so that there are about N/2 cells (never more than N) ->
// (cell id, point index) pairs sorted with the library's radix sort -> per-cell [first, last) ranges -> one thread
// per query walks the cells in shells of growing Chebyshev radius r around its own cell, keeps the K best in
// registers, and stops when the K-th best squared distance is <= (r * cell)^2, the closest anything outside the
// visited cube can be (or when the cube covers the grid).  Exact for any input, including duplicates and outliers.
// Everything runs on the caller's stream without host synchronisation.
#include "common.cuh"
#include <cfloat>

namespace sagars {

struct KnnGrid {
    float ox, oy, oz;      // grid origin (bounding-box minimum)
    float h, inv_h;        // cell size
    int nx, ny, nz;        // grid dimensions, nx * ny * nz <= num_points
    int ncell;
};

struct KnnTemp {
    int* bbox;             // 6 ordered-int encoded floats: min xyz, max xyz
    KnnGrid* grid;
    uint64_t* keys_a;      // sorted (cell id, point index) pairs land here
    uint64_t* keys_b;
    uint32_t* vals_a;
    uint32_t* vals_b;
    void* sort_temp;
    uint2* cell_range;     // [num_points] (first, last + 1) in the sorted order, (0, 0) for empty cells
    float4* sorted_pts;    // (x, y, z, original index as bits) in sorted order
};

static size_t knn_temp_layout(size_t n, KnnTemp* t, char* base)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return base ? base + at : (char*)nullptr; };
    char* p;
    p = take(64);      if (t) t->bbox = (int*)p;
    p = take(64);      if (t) t->grid = (KnnGrid*)p;
    p = take(n * 8);   if (t) t->keys_a = (uint64_t*)p;
    p = take(n * 8);   if (t) t->keys_b = (uint64_t*)p;
    p = take(n * 4);   if (t) t->vals_a = (uint32_t*)p;
    p = take(n * 4);   if (t) t->vals_b = (uint32_t*)p;
    p = take(sort_temp_bytes(n)); if (t) t->sort_temp = (void*)p;
    p = take(n * 8);   if (t) t->cell_range = (uint2*)p;
    p = take(n * 16);  if (t) t->sorted_pts = (float4*)p;
    return o + 256;
}

size_t knn_temp_bytes(size_t n) { return knn_temp_layout(n, nullptr, nullptr); }

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

// one thread: cell size for ~n/2 cells, at most n cells and 1024 per axis
__global__ void knn_grid_setup_kernel(int n, const int* __restrict__ bbox, KnnGrid* __restrict__ g)
{
    const float ox = ord2f(bbox[0]), oy = ord2f(bbox[1]), oz = ord2f(bbox[2]);
    float ex = ord2f(bbox[3]) - ox, ey = ord2f(bbox[4]) - oy, ez = ord2f(bbox[5]) - oz;
    const float emax = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-30f));
    // degenerate (flat / collinear) clouds: give the thin axes a small non-zero extent
    ex = fmaxf(ex, emax * 1e-6f); ey = fmaxf(ey, emax * 1e-6f); ez = fmaxf(ez, emax * 1e-6f);
    const float target = fmaxf(1.f, 0.5f * (float)n);
    float h = cbrtf(ex * ey * ez / target);
    h = fmaxf(h, emax / 1024.f);
    int nx, ny, nz;
    for (int it = 0; it < 64; it++) {
        nx = min(1024, (int)(ex / h) + 1);
        ny = min(1024, (int)(ey / h) + 1);
        nz = min(1024, (int)(ez / h) + 1);
        if ((long long)nx * ny * nz <= (long long)max(n, 1)) break;
        h *= 1.2599211f;   // halves the cell count
    }
    g->ox = ox; g->oy = oy; g->oz = oz;
    g->h = h; g->inv_h = 1.f / h;
    g->nx = nx; g->ny = ny; g->nz = nz;
    g->ncell = nx * ny * nz;
}

__device__ __forceinline__ void knn_cell_of(const KnnGrid& g, float x, float y, float z, int& cx, int& cy, int& cz)
{
    cx = min(g.nx - 1, max(0, (int)((x - g.ox) * g.inv_h)));
    cy = min(g.ny - 1, max(0, (int)((y - g.oy) * g.inv_h)));
    cz = min(g.nz - 1, max(0, (int)((z - g.oz) * g.inv_h)));
}

__global__ void __launch_bounds__(256)
knn_cell_keys_kernel(int n, const float* __restrict__ pts, const KnnGrid* __restrict__ gp, uint64_t* __restrict__ keys,
                     uint32_t* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const KnnGrid g = *gp;
    int cx, cy, cz;
    knn_cell_of(g, pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], cx, cy, cz);
    keys[i] = (uint64_t)((cz * g.ny + cy) * g.nx + cx);
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
knn_ranges_gather_kernel(int n, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                         const float* __restrict__ pts, uint2* __restrict__ cell_range, float4* __restrict__ sorted_pts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cur = (uint32_t)keys[i];
    if (i == 0) cell_range[cur].x = 0;
    else {
        const uint32_t prev = (uint32_t)keys[i - 1];
        if (cur != prev) { cell_range[prev].y = (uint32_t)i; cell_range[cur].x = (uint32_t)i; }
    }
    if (i == n - 1) cell_range[cur].y = (uint32_t)n;
    const uint32_t j = vals[i];
    sorted_pts[i] = make_float4(pts[3 * (size_t)j], pts[3 * (size_t)j + 1], pts[3 * (size_t)j + 2], __uint_as_float(j));
}

// One thread per query.  SELF: the query set is the point set and query i may not return point i (simple_knn).
template <int K, bool SELF>
__global__ void __launch_bounds__(128)
knn_search_kernel(int nq, const float* __restrict__ queries, const KnnGrid* __restrict__ gp,
                  const uint2* __restrict__ cell_range, const float4* __restrict__ sorted_pts,
                  int k_out, long long* __restrict__ idx_out, float* __restrict__ dist_out, float* __restrict__ mean_out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const KnnGrid g = *gp;
    const float qx = queries[3 * (size_t)q], qy = queries[3 * (size_t)q + 1], qz = queries[3 * (size_t)q + 2];
    int cx, cy, cz;
    knn_cell_of(g, qx, qy, qz, cx, cy, cz);

    float bd[K];
    uint32_t bi[K];
#pragma unroll
    for (int j = 0; j < K; j++) { bd[j] = FLT_MAX; bi[j] = 0xffffffffu; }

    auto visit_cell = [&](int x, int y, int z) {
        const uint2 r = cell_range[(z * g.ny + y) * g.nx + x];
        for (uint32_t i = r.x; i < r.y; i++) {
            const float4 p = sorted_pts[i];
            const uint32_t pid = __float_as_uint(p.w);
            if (SELF && pid == (uint32_t)q) continue;
            // same expression as simple_knn.cu:135-137 (d = candidate - query), so nvcc contracts it the same way
            const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            float d = dx * dx + dy * dy + dz * dz;
            if (!(d < bd[K - 1])) continue;
            uint32_t id = pid;
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (bd[j] > d) {
                    const float td = bd[j]; bd[j] = d; d = td;
                    const uint32_t ti = bi[j]; bi[j] = id; id = ti;
                }
            }
        }
    };

    const int rmax = max(max(cx, g.nx - 1 - cx), max(max(cy, g.ny - 1 - cy), max(cz, g.nz - 1 - cz)));
    for (int r = 0; r <= rmax; r++) {
        const int z0 = max(0, cz - r), z1 = min(g.nz - 1, cz + r);
        const int y0 = max(0, cy - r), y1 = min(g.ny - 1, cy + r);
        for (int z = z0; z <= z1; z++) {
            const bool zface = (z == cz - r) || (z == cz + r);
            for (int y = y0; y <= y1; y++) {
                const bool yface = (y == cy - r) || (y == cy + r);
                if (zface || yface) {
                    const int x0 = max(0, cx - r), x1 = min(g.nx - 1, cx + r);
                    for (int x = x0; x <= x1; x++) visit_cell(x, y, z);
                } else {
                    if (cx - r >= 0) visit_cell(cx - r, y, z);
                    if (r > 0 && cx + r <= g.nx - 1) visit_cell(cx + r, y, z);
                }
            }
        }
        // everything outside the visited cube is at least r * h away from the query (it lies inside its own cell)
        const float reach = (float)r * g.h * 0.9999f;   // margin for the rounding of the cell assignment
        if (bd[K - 1] <= reach * reach) break;
    }

    if (idx_out || dist_out) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (j < k_out) {
                if (idx_out) idx_out[(size_t)q * k_out + j] = (bi[j] == 0xffffffffu) ? -1ll : (long long)bi[j];
                if (dist_out) dist_out[(size_t)q * k_out + j] = bd[j];
            }
        }
    }
    if (mean_out) {
        // simple_knn.cu:183: (best[0] + best[1] + best[2]) / 3.0f, generalised to k_out terms in ascending order
        float sacc = 0.f;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (j < k_out) sacc += bd[j];
        mean_out[q] = sacc / (float)k_out;
    }
}

template <int K>
static void knn_launch_search(bool self, int nq, const float* queries, const KnnTemp& t, int k_out, long long* idx_out,
                              float* dist_out, float* mean_out, cudaStream_t s)
{
    const int blocks = (nq + 127) / 128;
    if (self) knn_search_kernel<K, true><<<blocks, 128, 0, s>>>(nq, queries, t.grid, t.cell_range, t.sorted_pts, k_out, idx_out, dist_out, mean_out);
    else knn_search_kernel<K, false><<<blocks, 128, 0, s>>>(nq, queries, t.grid, t.cell_range, t.sorted_pts, k_out, idx_out, dist_out, mean_out);
}

int launch_knn(int n, const float* points, int nq, const float* queries, int K, bool exclude_self, long long* idx_out,
               float* dist_out, float* mean_out, void* temp, cudaStream_t s)
{
    KnnTemp t;
    knn_temp_layout((size_t)n, &t, (char*)temp);
    knn_bbox_init_kernel<<<1, 32, 0, s>>>(t.bbox);
    SAGARS_LAUNCH_CHECK(s, false);
    knn_bbox_kernel<<<min((n + 255) / 256, 148 * 8), 256, 0, s>>>(n, points, t.bbox);
    SAGARS_LAUNCH_CHECK(s, false);
    knn_grid_setup_kernel<<<1, 1, 0, s>>>(n, t.bbox, t.grid);
    SAGARS_LAUNCH_CHECK(s, false);
    // cell ids < n: sort bits [0, bits(n))
    int bits = 1;
    while (bits < 32 && (1ll << bits) < (long long)n + 1) bits++;
    const bool start_alt = (sort_num_passes(bits) & 1) != 0;
    knn_cell_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, points, t.grid, start_alt ? t.keys_b : t.keys_a, start_alt ? t.vals_b : t.vals_a);
    SAGARS_LAUNCH_CHECK(s, false);
    bool in_a = true;
    int rc = launch_sort_pairs(nullptr, n, bits, t.keys_a, t.vals_a, t.keys_b, t.vals_b, t.sort_temp, sort_temp_bytes((size_t)n),
                               false, &in_a, s, false);
    if (rc) return rc;
    SAGARS_CUDA(cudaMemsetAsync(t.cell_range, 0, (size_t)n * sizeof(uint2), s));
    knn_ranges_gather_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, t.keys_a, t.vals_a, points, t.cell_range, t.sorted_pts);
    SAGARS_LAUNCH_CHECK(s, false);
    const float* qp = queries ? queries : points;
    if (K <= 4) knn_launch_search<4>(exclude_self, nq, qp, t, K, idx_out, dist_out, mean_out, s);
    else if (K <= 8) knn_launch_search<8>(exclude_self, nq, qp, t, K, idx_out, dist_out, mean_out, s);
    else if (K <= 16) knn_launch_search<16>(exclude_self, nq, qp, t, K, idx_out, dist_out, mean_out, s);
    else knn_launch_search<32>(exclude_self, nq, qp, t, K, idx_out, dist_out, mean_out, s);
    SAGARS_LAUNCH_CHECK(s, false);
    return SAGARS_OK;
}

}  // namespace sagars
