// cuda_emu.h - a small host emulation of the CUDA execution model, for CPU tests.
//
// TEST INFRASTRUCTURE ONLY.  It lets `g++` compile a kernel translation unit of
// pyqed_b200/csrc unchanged (the hardware wrappers of heom_device.cuh are
// replaced below) and run it with one OS thread per CUDA thread:
//   * a CTA = blockDim.x std::threads; CTAs of a grid run one after the other;
//   * __syncthreads / __syncwarp(mask) / __reduce_*_sync / __shfl_sync are real
//     barriers among the participating threads, so lanes run in arbitrary order
//     between synchronisation points (more adversarial than the hardware: a
//     missing __syncwarp shows up as a data race / wrong result);
//   * dynamic shared memory is a per-CTA buffer filled with NaN bit patterns;
//   * cp.async / cp.async.bulk + mbarrier are modelled at the two extremes of
//     their legal timing (copies land at issue, or only when waited for - see
//     "asynchronous copies" below); proxy fences are no-ops (cross-proxy ordering
//     is NOT modelled - that is checked on the GPU by the parity tests and
//     compute-sanitizer).
// What it checks is a kernel's indexing, table formats, stage algebra and
// edge-case handling against the oracle, without a GPU.
#pragma once
#define HEOM_HOST_EMU 1

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct double2 { double x, y; };
struct int2 { int x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __grid_constant__

using std::max;
using std::min;

namespace emu {

struct Dim { unsigned x = 1, y = 1, z = 1; };

class Barrier {
public:
    void wait(int n) {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned long g = gen_;
        if (++count_ == n) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen_ != g; });
        }
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int count_ = 0;
    unsigned long gen_ = 0;
};

struct Warp {
    std::mutex m;
    std::map<unsigned, std::unique_ptr<Barrier>> bars;   // one barrier per participation mask
    long long xchg[32];
    Barrier& bar(unsigned mask) {
        std::lock_guard<std::mutex> lk(m);
        auto& b = bars[mask];
        if (!b) b.reset(new Barrier);
        return *b;
    }
};

struct Cta {
    std::vector<char> smem;
    std::vector<Warp> warps;
    Barrier all;
    int nthreads = 0;
};

inline thread_local Dim t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local Cta* t_cta = nullptr;

inline Warp& my_warp() { return t_cta->warps[t_threadIdx.x >> 5]; }
inline int popc(unsigned m) { return __builtin_popcount(m); }

template <typename Kernel, typename... Args>
void launch(Kernel kernel, unsigned grid, unsigned block, size_t smem_bytes, Args... args) {
    if (block == 0 || block > 1024 || smem_bytes > 232448) {   // sm_100a: 1024 threads, 227 KB per CTA
        std::fprintf(stderr, "cuda_emu: illegal launch (%u threads, %zu bytes of shared memory)\n", block, smem_bytes);
        std::abort();
    }
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta;
        cta.smem.assign(smem_bytes, (char)0xff);   // NaN patterns: stale reads poison the result
        cta.warps = std::vector<Warp>((block + 31) / 32);
        cta.nthreads = (int)block;
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t, b] {
                t_threadIdx = Dim{t, 0, 0};
                t_blockIdx = Dim{b, 0, 0};
                t_blockDim = Dim{block, 1, 1};
                t_gridDim = Dim{grid, 1, 1};
                t_cta = &cta;
                kernel(args...);
            });
        for (auto& x : th) x.join();
    }
}

}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

#define HEOM_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::t_cta->smem.data())
#define HEOM_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(kernel, (unsigned)(grid), (unsigned)(block), (size_t)(smem), __VA_ARGS__)

// ---- synchronisation and warp collectives ------------------------------------
inline void __syncthreads() { emu::t_cta->all.wait(emu::t_cta->nthreads); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::my_warp().bar(mask).wait(emu::popc(mask)); }

template <typename F>
inline long long emu_collective(unsigned mask, long long mine, F combine) {
    emu::Warp& w = emu::my_warp();
    const int lane = threadIdx.x & 31;
    w.xchg[lane] = mine;
    __syncwarp(mask);
    const long long r = combine(w.xchg);
    __syncwarp(mask);
    return r;
}
inline int __reduce_max_sync(unsigned mask, int v) {
    return (int)emu_collective(mask, v, [&](const long long* x) {
        long long r = INT32_MIN;
        for (int l = 0; l < 32; ++l)
            if (mask >> l & 1u) r = std::max(r, x[l]);
        return r;
    });
}
inline int __reduce_add_sync(unsigned mask, int v) {
    return (int)emu_collective(mask, v, [&](const long long* x) {
        long long r = 0;
        for (int l = 0; l < 32; ++l)
            if (mask >> l & 1u) r += x[l];
        return r;
    });
}
inline int __shfl_sync(unsigned mask, int v, int src) {
    return (int)emu_collective(mask, v, [&](const long long* x) { return x[src & 31]; });
}

// ---- memory access wrappers ---------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, const T v) { *p = v; }

inline unsigned smem_u32(const void* p) {
    return (unsigned)(reinterpret_cast<const char*>(p) - emu::t_cta->smem.data());
}

// ---- asynchronous copies --------------------------------------------------------
// Two timing models, chosen per launch with emu::g_async_late:
//   0  a copy lands the moment it is issued (the earliest legal time: exposes
//      write-after-read hazards on a buffer that is still being read);
//   1  a cp.async copy lands when its group is waited for, a bulk copy when a
//      thread waits on its mbarrier (the latest legal time: exposes reads before
//      the wait, wrong wait counts, wrong expect_tx byte counts and phase-parity
//      mistakes - a wait that can never complete aborts the process).
namespace emu {
inline int g_async_late = 0;
struct Copy { void* dst; const void* src; unsigned bytes; };
inline thread_local std::vector<std::vector<Copy>> t_groups;   // committed cp.async groups, oldest first
inline thread_local std::vector<Copy> t_open;                  // copies of the group not yet committed
struct MBar {
    unsigned long completed = 0;   // phases completed so far
    long long expected = 0, arrived = 0;
    bool armed = false;
    std::vector<Copy> copies;
};
struct MBars {
    std::mutex m;
    std::condition_variable cv;
    std::map<const void*, MBar> bars;
};
inline MBars& mbars() {
    static MBars all;   // keyed by shared-memory address; CTAs run one after the other
    return all;
}
inline void apply(const Copy& c) { std::memcpy(c.dst, c.src, c.bytes); }
// Operand rules of the hardware copies that a memcpy would silently forgive: cp.async ..., 16
// needs both addresses 16-byte aligned; cp.async.bulk needs 16-byte aligned addresses and a
// size that is a non-zero multiple of 16; the shared-memory side must lie inside the CTA's
// dynamic shared memory; an mbarrier is 8-byte aligned and its tx-count is below 2^20.
inline void require(bool ok, const char* what) {
    if (ok) return;
    std::fprintf(stderr, "cuda_emu: illegal operand: %s\n", what);
    std::abort();
}
inline bool in_smem(const void* p, size_t bytes) {
    const char* b = t_cta->smem.data();
    const char* q = reinterpret_cast<const char*>(p);
    return q >= b && q + bytes <= b + t_cta->smem.size();
}
inline bool smem_aligned(const void* p, size_t a) {
    return (size_t)(reinterpret_cast<const char*>(p) - t_cta->smem.data()) % a == 0;
}
inline bool gmem_aligned(const void* p, size_t a) { return reinterpret_cast<uintptr_t>(p) % a == 0; }
}  // namespace emu

inline void cp_async16(void* smem_dst, const void* gsrc) {
    emu::require(emu::in_smem(smem_dst, 16) && emu::smem_aligned(smem_dst, 16), "cp.async 16: shared destination");
    emu::require(emu::gmem_aligned(gsrc, 16), "cp.async 16: global source not 16-byte aligned");
    if (emu::g_async_late) emu::t_open.push_back(emu::Copy{smem_dst, gsrc, 16});
    else std::memcpy(smem_dst, gsrc, 16);
}
inline void cp_async16_s(unsigned smem_dst_u32, const void* gsrc) {
    cp_async16(emu::t_cta->smem.data() + smem_dst_u32, gsrc);
}
inline void cp_async16_s_if(unsigned smem_dst_u32, const void* gsrc, bool pred) {
    if (pred) cp_async16_s(smem_dst_u32, gsrc);
}
inline void cp_async_commit() {
    if (!emu::g_async_late) return;
    emu::t_groups.push_back(std::move(emu::t_open));
    emu::t_open.clear();
}
template <int NWAIT> inline void cp_async_wait() {
    while ((int)emu::t_groups.size() > NWAIT) {
        for (const auto& c : emu::t_groups.front()) emu::apply(c);
        emu::t_groups.erase(emu::t_groups.begin());
    }
}
inline void mbar_init(unsigned long long* bar, unsigned) {
    emu::require(emu::in_smem(bar, 8) && emu::smem_aligned(bar, 8), "mbarrier.init: address");
    std::lock_guard<std::mutex> lk(emu::mbars().m);
    emu::mbars().bars[bar] = emu::MBar{};
}
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    emu::require(bytes < (1u << 20), "mbarrier expect_tx: tx-count out of range");
    std::lock_guard<std::mutex> lk(emu::mbars().m);
    emu::MBar& b = emu::mbars().bars[bar];
    b.expected += bytes;
    b.armed = true;
    emu::mbars().cv.notify_all();
}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    emu::require(bytes > 0 && bytes % 16 == 0, "cp.async.bulk global->shared: size not a non-zero multiple of 16");
    emu::require(emu::in_smem(dst, bytes) && emu::smem_aligned(dst, 16), "cp.async.bulk global->shared: shared destination");
    emu::require(emu::gmem_aligned(src, 16), "cp.async.bulk global->shared: global source not 16-byte aligned");
    std::lock_guard<std::mutex> lk(emu::mbars().m);
    emu::MBar& b = emu::mbars().bars[bar];
    if (emu::g_async_late) b.copies.push_back(emu::Copy{dst, src, bytes});
    else std::memcpy(dst, src, bytes);
    b.arrived += bytes;
    emu::mbars().cv.notify_all();
}
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
    std::unique_lock<std::mutex> lk(emu::mbars().m);
    emu::MBar& b = emu::mbars().bars[bar];
    for (;;) {
        if ((b.completed & 1u) != parity) return;             // that phase is over
        if (b.armed && b.arrived == b.expected) {              // every byte is in: the phase completes
            for (const auto& c : b.copies) emu::apply(c);
            b.copies.clear();
            b.expected = b.arrived = 0;
            b.armed = false;
            ++b.completed;
            emu::mbars().cv.notify_all();
            return;
        }
        if (emu::mbars().cv.wait_for(lk, std::chrono::seconds(20)) == std::cv_status::timeout) {
            std::fprintf(stderr, "cuda_emu: mbarrier wait cannot complete (expected %lld bytes, arrived %lld)\n",
                         b.expected, b.arrived);
            std::abort();
        }
    }
}
inline void fence_proxy_async() {}
// shared -> global bulk stores: performed at issue, or (late model) when the issuing thread
// waits for its bulk groups - the source must stay intact until then
namespace emu {
inline thread_local std::vector<std::pair<Copy, std::vector<char>>> t_bulk_stores;
}
inline void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    emu::require(bytes > 0 && bytes % 16 == 0, "cp.async.bulk shared->global: size not a non-zero multiple of 16");
    emu::require(emu::in_smem(smem_src, bytes) && emu::smem_aligned(smem_src, 16), "cp.async.bulk shared->global: shared source");
    emu::require(emu::gmem_aligned(gdst, 16), "cp.async.bulk shared->global: global destination not 16-byte aligned");
    if (emu::g_async_late) emu::t_bulk_stores.push_back({emu::Copy{gdst, smem_src, bytes}, {}});
    else std::memcpy(gdst, smem_src, bytes);
}
inline void bulk_commit() {}
inline void bulk_wait_read() {
    for (auto& c : emu::t_bulk_stores) emu::apply(c.first);   // reads the source as late as allowed
    emu::t_bulk_stores.clear();
}
inline void bulk_wait_all() { bulk_wait_read(); }
inline void __threadfence_system() {}
// L2 eviction-priority hints have no effect on results: the hinted copies are the plain ones
inline unsigned long long l2_policy_evict_first() { return 1; }
inline unsigned long long l2_policy_evict_last() { return 2; }
inline void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar, unsigned long long) {
    bulk_g2s(dst, src, bytes, bar);
}
inline void cp_async16_s_if_hint(unsigned smem_dst_u32, const void* gsrc, bool pred, unsigned long long) {
    cp_async16_s_if(smem_dst_u32, gsrc, pred);
}
// global atomics: the emulated CUDA threads are OS threads
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
