// A minimal CUDA *execution shim* for the host: just enough of the device-side vocabulary (vector types, thread indices,
// barriers, warp shuffles, atomics, bit casts) to compile simple kernels with g++ and run them with one OS thread per CUDA
// thread.  TEST INFRASTRUCTURE ONLY (tests/test_tile_sort_emulated.py); it shadows <cuda_runtime.h> for the kernel headers
// it is used with.  Semantics: blocks run one after the other, the threads of a block concurrently (real races and real
// atomics), __shared__ variables are function-local statics (uninitialised across blocks, like the real thing),
// __syncthreads / full-mask warp shuffles are barriers that tolerate threads that have already left the kernel.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) alignas(n)
#define SAGARS_DYNAMIC_SMEM(name) unsigned char* name = ::cuda_emu::dynamic_smem

typedef void* cudaStream_t;      // host-side declarations of the product headers only need the names
typedef int cudaError_t;

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float __frcp_rn(float x) { return 1.0f / x; }

inline thread_local uint3 threadIdx, blockIdx;
inline uint3 blockDim, gridDim;

namespace cuda_emu {

// barrier over the threads of a group that are still inside the kernel.  The bookkeeping sits under a mutex (arrivals and
// departures must be decided atomically together), the waiting does not: a block has up to 1024 OS threads on a handful of
// cores, and waking them through a condition variable spent most of the suite's time in futex calls -- they yield on the
// generation counter instead.
struct Barrier {
    std::mutex m;
    int active = 0, waiting = 0;
    std::atomic<uint64_t> gen{0};
    void reset(int n) { std::lock_guard<std::mutex> lk(m); active = n; waiting = 0; }
    void sync()
    {
        uint64_t g;
        {
            std::lock_guard<std::mutex> lk(m);
            g = gen.load();
            if (++waiting == active) { waiting = 0; gen.store(g + 1); return; }
        }
        while (gen.load() == g) std::this_thread::yield();
    }
    // a thread that has left the kernel no longer takes part; if everybody else is already waiting, let them go
    void drop()
    {
        std::lock_guard<std::mutex> lk(m);
        --active;
        if (active > 0 && waiting == active) { waiting = 0; gen.fetch_add(1); }
    }
};

struct Warp {
    Barrier bar;
    uint64_t slot[32];
};

inline Barrier block_bar, outer_bar;
inline void (*thread_exit_hook)() = nullptr;     // e.g. "outstanding asynchronous copies of this thread land now"
inline std::vector<Warp>* warps = nullptr;
inline unsigned char* dynamic_smem = nullptr;

template <class T>
inline T shfl_from(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Warp& w = (*warps)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.slot[lane] = bits;
    w.bar.sync();
    T r = v;
    if (src_lane >= 0 && src_lane < 32) std::memcpy(&r, &w.slot[src_lane], sizeof(T));
    w.bar.sync();
    return r;
}

// kernel<<<dim3(grid_x, grid_y), block, smem>>>(args...) with a 1-D block
template <class K, class... A>
void launch2d(unsigned grid_x, unsigned grid_y, unsigned block, size_t smem_bytes, K kernel, A... args)
{
    const unsigned grid = grid_x * grid_y;
    std::vector<unsigned char> smem(smem_bytes + 2048);
    dynamic_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem.data()) + 1023u) & ~uintptr_t(1023));
    std::vector<Warp> w((block + 31) / 32);
    warps = &w;
    blockDim = {block, 1, 1};
    gridDim = {grid_x, grid_y, 1};
    outer_bar.reset((int)block);
    auto arm = [&] {
        block_bar.reset((int)block);
        for (unsigned i = 0; i < w.size(); i++) w[i].bar.reset((int)std::min(32u, block - 32 * i));
    };
    arm();
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < block; t++) {
        pool.emplace_back([&, t] {
            threadIdx = {t, 0, 0};
            for (unsigned b = 0; b < grid; b++) {
                blockIdx = {b % grid_x, b / grid_x, 0};
                kernel(args...);
                if (thread_exit_hook) thread_exit_hook();
                block_bar.drop();                  // this thread has left the kernel: later barriers do not wait for it
                (*warps)[t >> 5].bar.drop();
                outer_bar.sync();
                if (t == 0) arm();
                outer_bar.sync();
            }
        });
    }
    for (auto& th : pool) th.join();
    warps = nullptr;
    dynamic_smem = nullptr;
}

template <class K, class... A>
void launch(unsigned grid, unsigned block, size_t smem_bytes, K kernel, A... args)
{
    launch2d(grid, 1u, block, smem_bytes, kernel, args...);
}

}  // namespace cuda_emu

inline void __syncthreads() { cuda_emu::block_bar.sync(); }
// barrier + AND over the threads still inside the kernel.  Two alternating flags: call k uses flag k & 1 and every thread
// clears it again after the second barrier, long before call k + 2 can set it.
inline int __syncthreads_and(int pred)
{
    static unsigned char some_false[2];
    static thread_local unsigned call = 0;
    unsigned char& f = some_false[call++ & 1u];
    if (!pred) __atomic_store_n(&f, 1, __ATOMIC_RELAXED);
    cuda_emu::block_bar.sync();
    const int r = !__atomic_load_n(&f, __ATOMIC_RELAXED);
    cuda_emu::block_bar.sync();
    __atomic_store_n(&f, 0, __ATOMIC_RELAXED);
    return r;
}
// width: the warp is split into segments of `width` lanes and the source lane is taken inside the caller's segment
template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const int lane = threadIdx.x & 31;
    return cuda_emu::shfl_from(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32)
{
    const int lane = threadIdx.x & 31, src = lane ^ lane_mask;
    return cuda_emu::shfl_from(v, (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned delta)
{
    const int lane = threadIdx.x & 31;
    return cuda_emu::shfl_from(v, lane - (int)delta >= 0 ? lane - (int)delta : lane);
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float* p, float v)
{
    unsigned int* u = reinterpret_cast<unsigned int*>(p);
    unsigned int o = __atomic_load_n(u, __ATOMIC_RELAXED), w;
    float old, want;
    do {
        std::memcpy(&old, &o, 4);
        want = old + v;
        std::memcpy(&w, &want, 4);
    } while (!__atomic_compare_exchange_n(u, &o, w, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&old, &o, 4);
    return old;
}
inline unsigned atomicMax(unsigned* p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicMax(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline int atomicMin(int* p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
// full-mask warp votes / matches over the lanes that take part (every lane of the warp must call, as on the device)
template <class T> inline unsigned __match_any_sync(unsigned, T v)
{
    cuda_emu::Warp& w = (*cuda_emu::warps)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.slot[lane] = bits;
    w.bar.sync();
    unsigned m = 0;
    const unsigned lanes = blockDim.x - (threadIdx.x & ~31u) < 32u ? blockDim.x - (threadIdx.x & ~31u) : 32u;
    for (unsigned l = 0; l < lanes; l++) m |= (w.slot[l] == bits) ? (1u << l) : 0u;
    w.bar.sync();
    return m;
}
inline unsigned __ballot_sync(unsigned, bool pred)
{
    cuda_emu::Warp& w = (*cuda_emu::warps)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    w.slot[lane] = pred ? 1u : 0u;
    w.bar.sync();
    unsigned m = 0;
    const unsigned lanes = blockDim.x - (threadIdx.x & ~31u) < 32u ? blockDim.x - (threadIdx.x & ~31u) : 32u;
    for (unsigned l = 0; l < lanes; l++) m |= w.slot[l] ? (1u << l) : 0u;
    w.bar.sync();
    return m;
}
inline unsigned emu_warp_lanes() { const unsigned rest = blockDim.x - (threadIdx.x & ~31u); return rest < 32u ? rest : 32u; }
inline bool __any_sync(unsigned, bool pred) { return __ballot_sync(0xffffffffu, pred) != 0u; }
inline bool __all_sync(unsigned, bool pred)
{
    const unsigned lanes = emu_warp_lanes();
    return __ballot_sync(0xffffffffu, pred) == (lanes == 32u ? 0xffffffffu : ((1u << lanes) - 1u));
}
inline void __syncwarp(unsigned = 0xffffffffu) { (*cuda_emu::warps)[threadIdx.x >> 5].bar.sync(); }
template <class T> inline T __ldg(const T* p) { return *p; }
