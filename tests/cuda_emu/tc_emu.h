// Host restatement of the tcgen05 / TMEM entry points of seganygaussians_b200/csrc/tc.cuh for the CPU execution shim.
// TEST INFRASTRUCTURE ONLY.  What is modelled:
//   * tcgen05.mma.kind::tf32 with shared-memory matrix descriptors in the SWIZZLE_NONE canonical K-major layout documented in
//     tc.cuh (start address, leading / stride byte offsets in 16-byte units; instruction descriptor: N >> 3 at bit 17, M >> 4 at
//     bit 24): D[m][n] (+)= sum_k tf32(A[m][k]) * tf32(B[n][k]) over the 8 k of one instruction, accumulator in a 128-lane x
//     512-column TMEM array (address = lane << 16 | column);
//   * the MMAs are ASYNCHRONOUS: they are queued at issue, handed to an mbarrier by tcgen05.commit, and only executed -- reading
//     their operand tiles from shared memory at that moment -- when some thread waits for that barrier.  Overwriting an operand
//     tile before the wait therefore corrupts the result here as it would on the tensor core;
//   * mbarrier phase parity, tcgen05.ld.32x32b (thread = TMEM lane of the warp's 32-lane window), bar.sync id, 128.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace sagars {
namespace tc {

namespace emu {
struct Mma { uint32_t tmem_d; uint64_t a_desc, b_desc; uint32_t idesc, accumulate; int a_tmem = 0; };
struct Bar { int init = 0, pending = 0; unsigned phase = 0; std::vector<Mma> in_flight; };
inline std::mutex mu;
inline std::map<const void*, Bar> bars;
inline thread_local std::vector<Mma> issued;           // MMAs of this thread not yet committed
inline float tmem[128][512];

inline float tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

inline void execute(const Mma& m)
{
    const int N = (int)((m.idesc >> 17) & 0x3F) << 3, M = (int)((m.idesc >> 24) & 0x1F) << 4;
    const bool a_mn = ((m.idesc >> 15) & 1u) != 0;
    if (((m.idesc >> 16) & 1u) || M != 128) {
        std::fprintf(stderr, "[cuda_emu] tcgen05.mma: only a K-major B operand and M = 128 are modelled\n");
        std::abort();
    }
    auto field = [](uint64_t d, int sh) { return (size_t)((d >> sh) & 0x3FFF) << 4; };
    const unsigned char* base = ::cuda_emu::dynamic_smem;
    const size_t a0 = field(m.a_desc, 0), a_lbo = field(m.a_desc, 16), a_sbo = field(m.a_desc, 32);
    const size_t b0 = field(m.b_desc, 0), b_lbo = field(m.b_desc, 16), b_sbo = field(m.b_desc, 32);
    auto elem = [&](size_t start, size_t lbo, size_t sbo, int r, int k) {
        float v;
        std::memcpy(&v, base + start + (size_t)(k / 4) * lbo + (size_t)(r / 8) * sbo + (size_t)(r % 8) * 16 + (size_t)(k % 4) * 4, 4);
        return (double)tf32(v);
    };
    // MN-major operand (tc.cuh): element (row, k) at (k/8)*LBO + (row/4)*SBO + (k%8)*16 + (row%4)*4; rows may alias other data
    // (the caller then ignores their accumulator rows): read as raw bits, NaN / Inf stay confined to their own row
    auto elem_mn = [&](size_t start, size_t lbo, size_t sbo, int r, int k) {
        float v;
        std::memcpy(&v, base + start + (size_t)(k / 8) * lbo + (size_t)(r / 4) * sbo + (size_t)(k % 8) * 16 + (size_t)(r % 4) * 4, 4);
        return (double)tf32(v);
    };
    const int lane0 = (int)(m.tmem_d >> 16), col0 = (int)(m.tmem_d & 0xFFFF);
    for (int r = 0; r < M; r++)
        for (int n = 0; n < N; n++) {
            double acc = m.accumulate ? (double)tmem[lane0 + r][col0 + n] : 0.0;
            for (int k = 0; k < 8; k++) {
                // A in tensor memory: row r = lane (lane field of the address) + r, 8 consecutive columns
                const double av = m.a_tmem ? (double)tf32(tmem[(int)((uint32_t)m.a_desc >> 16) + r][(int)(m.a_desc & 0xFFFF) + k])
                                           : (a_mn ? elem_mn(a0, a_lbo, a_sbo, r, k) : elem(a0, a_lbo, a_sbo, r, k));
                acc += av * elem(b0, b_lbo, b_sbo, n, k);
            }
            tmem[lane0 + r][col0 + n] = (float)acc;
        }
}

struct NamedBarrier {
    std::mutex m;
    int waiting = 0;
    std::atomic<uint64_t> gen{0};
    void sync(int count)
    {
        uint64_t g;
        {
            std::lock_guard<std::mutex> lk(m);
            g = gen.load();
            if (++waiting == count) { waiting = 0; gen.store(g + 1); return; }
        }
        while (gen.load() == g) std::this_thread::yield();
    }
};
inline NamedBarrier named[16];
}  // namespace emu

inline void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    emu::issued.push_back({tmem_d, a_desc, b_desc, idesc, accumulate});
}
inline void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    emu::Mma m{tmem_d, (uint64_t)tmem_a, b_desc, idesc, accumulate};
    m.a_tmem = 1;
    emu::issued.push_back(m);
}
inline void commit(uint64_t* bar)
{
    std::lock_guard<std::mutex> lk(emu::mu);
    emu::Bar& b = emu::bars.at(bar);
    b.in_flight.insert(b.in_flight.end(), emu::issued.begin(), emu::issued.end());
    emu::issued.clear();
    emu::Mma marker{0xFFFFFFFFu, 0, 0, 0, 0};
    b.in_flight.push_back(marker);                             // marker: one arrival once everything before it has executed
}
inline void mbar_init(uint64_t* bar, uint32_t count)
{
    std::lock_guard<std::mutex> lk(emu::mu);
    emu::Bar& b = emu::bars[bar];
    b = emu::Bar();
    b.init = b.pending = (int)count;
}
inline void mbar_init_fence() {}
inline void mbar_arrive(uint64_t* bar)                         // a thread's own arrival: immediate
{
    std::lock_guard<std::mutex> lk(emu::mu);
    emu::Bar& b = emu::bars.at(bar);
    if (--b.pending == 0) { b.phase ^= 1u; b.pending = b.init; }
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    std::lock_guard<std::mutex> lk(emu::mu);
    emu::Bar& b = emu::bars.at(bar);
    if (b.phase != (parity & 1u)) return true;
    std::vector<emu::Mma> work;                                // the tensor core gets to it only now
    work.swap(b.in_flight);
    for (const emu::Mma& m : work) {
        if (m.tmem_d == 0xFFFFFFFFu) {
            if (--b.pending == 0) { b.phase ^= 1u; b.pending = b.init; }
        } else {
            emu::execute(m);
        }
    }
    return b.phase != (parity & 1u);
}
inline bool mbar_test_wait(uint64_t* bar, uint32_t parity) { return mbar_try_wait(bar, parity); }
inline void st_release_cta(int* p, int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline int ld_acquire_cta(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void backoff(unsigned) { std::this_thread::sleep_for(std::chrono::microseconds(20)); }
inline void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(20);
    while (!mbar_try_wait(bar, parity)) {
        if (std::chrono::steady_clock::now() > deadline) {
            {
                std::lock_guard<std::mutex> lk(emu::mu);
                const emu::Bar& b = emu::bars.at(bar);
                std::fprintf(stderr, "[cuda_emu] tcgen05 mbarrier wait timed out: thread %u block (%u,%u) parity %u, barrier at smem+%td: count %d pending %d "
                             "phase %u, %zu operations in flight\n", threadIdx.x, blockIdx.x, blockIdx.y, parity,
                             (const unsigned char*)bar - ::cuda_emu::dynamic_smem, b.init, b.pending, b.phase, b.in_flight.size());
                for (const auto& kv : emu::bars) {
                    const ptrdiff_t off = (const unsigned char*)kv.first - ::cuda_emu::dynamic_smem;
                    if (off >= 0 && off < (1 << 20))
                        std::fprintf(stderr, "[cuda_emu]   barrier smem+%td: count %d pending %d phase %u in flight %zu\n", off, kv.second.init,
                                     kv.second.pending, kv.second.phase, kv.second.in_flight.size());
                }
            }
            std::abort();
        }
        // back off: hundreds of polling threads otherwise starve the one thread everybody is waiting for (emu::mu is not fair)
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}
inline void mbar_wait_long(uint64_t* bar, uint32_t parity, uint32_t = 0) { mbar_wait(bar, parity); }
inline void fence_smem_to_async_proxy() {}
inline void fence_before_sync() {}
inline void fence_after_sync() {}
// warp-collective on the device (one write); here every lane of the warp stores the same value -- atomically, so that a
// ThreadSanitizer build of the suite (tests/conftest.py, SAGARS_EMU_TSAN) reports the kernels' races, not the shim's
template <int NCOLS> inline void tmem_alloc(uint32_t* slot) { __atomic_store_n(slot, 0u, __ATOMIC_RELAXED); }
template <int NCOLS> inline void tmem_dealloc(uint32_t) {}
inline void tmem_ld32(uint32_t taddr, float* v)
{
    const int lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < 32; i++) v[i] = emu::tmem[lane][col + i];
}
inline void tmem_ld8(uint32_t taddr, float* v)
{
    const int lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < 8; i++) v[i] = emu::tmem[lane][col + i];
}
inline void tmem_st32(uint32_t taddr, const float* v)
{
    const int lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < 32; i++) emu::tmem[lane][col + i] = v[i];
}
inline void tmem_ld16(uint32_t taddr, float* v)
{
    const int lane = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xFFFF);
    for (int i = 0; i < 16; i++) v[i] = emu::tmem[lane][col + i];
}
inline void bar_sync_128(int id) { emu::named[id & 15].sync(128); }

}  // namespace tc
}  // namespace sagars
