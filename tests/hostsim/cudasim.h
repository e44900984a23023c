// cudasim.h — TEST INFRASTRUCTURE ONLY.  A minimal host-side emulation of the CUDA execution model so that the
// product kernels (visioncortex_b200/csrc/*.cuh) can be compiled with g++ and exercised against the oracle on a
// machine without a GPU (the build container has none).  It is never linked into, loaded by, or shipped with the
// product library: libvcb200.so is compiled by nvcc for sm_100a only and has no CPU path.
//
// Model: one std::thread per CUDA thread of a block (blocks run one after another; cooperative launches run all
// blocks at once), std::barrier for __syncthreads / warp collectives / grid sync, GCC __atomic builtins for atomics.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
struct uchar4 { unsigned char x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace sim {
struct WarpState {
    std::barrier<> bar{32};
    unsigned long long slot[32];
    explicit WarpState(int n) : bar(n) {}
};
struct BlockState {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<WarpState>> warps;
    std::vector<unsigned char> smem;
};
struct ThreadCtx {
    dim3 tid, bid, bdim, gdim;
    BlockState *blk = nullptr;
    std::barrier<> *grid_bar = nullptr;
    std::barrier<> *cluster_bar = nullptr;
    int lane = 0, warp = 0, warp_size = 32;
};
inline thread_local ThreadCtx tc;

// Run `fn` for every thread of every block.  mode 0: sequential (no collectives allowed), 1: threaded per block,
// 2: cooperative (all blocks concurrently, grid sync allowed).
inline void launch(dim3 grid, dim3 block, size_t smem, int mode, const std::function<void()> &fn, unsigned cluster = 1) {
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nblocks = grid.x * grid.y * grid.z;
    if (nblocks == 0 || nthreads == 0) return;
    auto setup = [&](ThreadCtx &c, unsigned b, unsigned t) {
        c.bdim = block; c.gdim = grid;
        c.bid = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
        c.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
        c.lane = (int)(t % 32); c.warp = (int)(t / 32);
    };
    if (mode == 0) {
        BlockState bs; bs.smem.resize(smem + 16);
        for (unsigned b = 0; b < nblocks; ++b)
            for (unsigned t = 0; t < nthreads; ++t) {
                setup(tc, b, t); tc.blk = &bs; tc.grid_bar = nullptr;
                fn();
            }
        return;
    }
    auto make_block = [&](BlockState &bs) {
        bs.bar.reset(new std::barrier<>(nthreads));
        bs.smem.assign(smem + 16, 0);
        for (unsigned w = 0; w * 32 < nthreads; ++w)
            bs.warps.emplace_back(new WarpState((int)std::min(32u, nthreads - w * 32)));
    };
    if (mode == 1) {
        for (unsigned b = 0; b < nblocks; ++b) {
            BlockState bs; make_block(bs);
            std::vector<std::thread> th;
            th.reserve(nthreads);
            for (unsigned t = 0; t < nthreads; ++t)
                th.emplace_back([&, b, t] { setup(tc, b, t); tc.blk = &bs; tc.grid_bar = nullptr; fn(); });
            for (auto &x : th) x.join();
        }
        return;
    }
    std::vector<BlockState> bss(nblocks);
    for (auto &bs : bss) make_block(bs);
    std::barrier<> gbar(nthreads * nblocks);
    // thread-block clusters: consecutive groups of `cluster` blocks share a hardware barrier
    std::vector<std::unique_ptr<std::barrier<>>> cbars;
    for (unsigned c = 0; c * cluster < nblocks; ++c) cbars.emplace_back(new std::barrier<>(nthreads * cluster));
    std::vector<std::thread> th;
    for (unsigned b = 0; b < nblocks; ++b)
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, b, t] { setup(tc, b, t); tc.blk = &bss[b]; tc.grid_bar = &gbar; tc.cluster_bar = cbars[b / cluster].get(); fn(); });
    for (auto &x : th) x.join();
}
inline void *smem_ptr() {
    auto p = (uintptr_t)tc.blk->smem.data();
    return (void *)((p + 15) & ~(uintptr_t)15);
}
inline WarpState &warp() { return *tc.blk->warps[tc.warp]; }
inline unsigned long long xchg(unsigned long long v, int src) {
    WarpState &w = warp();
    w.slot[tc.lane] = v;
    w.bar.arrive_and_wait();
    unsigned long long r = w.slot[src & 31];
    w.bar.arrive_and_wait();
    return r;
}
}  // namespace sim

#define threadIdx (sim::tc.tid)
#define blockIdx (sim::tc.bid)
#define blockDim (sim::tc.bdim)
#define gridDim (sim::tc.gdim)
static const int warpSize = 32;

static inline void __syncthreads() { sim::tc.blk->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { sim::warp().bar.arrive_and_wait(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    unsigned long long u = 0; std::memcpy(&u, &v, sizeof(T));
    u = sim::xchg(u, src); T r; std::memcpy(&r, &u, sizeof(T)); return r;
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    int src = sim::tc.lane - (int)d; T r = __shfl_sync(m, v, src < 0 ? sim::tc.lane : src); return src < 0 ? v : r;
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
    int src = sim::tc.lane + (int)d; T r = __shfl_sync(m, v, src > 31 ? sim::tc.lane : src); return src > 31 ? v : r;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { return __shfl_sync(m, v, sim::tc.lane ^ x); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    sim::WarpState &w = sim::warp();
    w.slot[sim::tc.lane] = pred ? 1 : 0;
    w.bar.arrive_and_wait();
    unsigned r = 0;
    int n = (int)std::min<unsigned>(32, blockDim.x * blockDim.y * blockDim.z - sim::tc.warp * 32);
    for (int i = 0; i < n; ++i) r |= (unsigned)(w.slot[i] & 1) << i;
    w.bar.arrive_and_wait();
    return r;
}
static inline unsigned __match_any_sync(unsigned, unsigned v) {
    sim::WarpState &w = sim::warp();
    w.slot[sim::tc.lane] = v;
    w.bar.arrive_and_wait();
    unsigned r = 0;
    int n = (int)std::min<unsigned>(32, blockDim.x * blockDim.y * blockDim.z - sim::tc.warp * 32);
    for (int i = 0; i < n; ++i) r |= (unsigned)((unsigned)w.slot[i] == v) << i;
    w.bar.arrive_and_wait();
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    int n = (int)std::min<unsigned>(32, blockDim.x * blockDim.y * blockDim.z - sim::tc.warp * 32);
    unsigned full = n == 32 ? 0xffffffffu : ((1u << n) - 1);
    return __ballot_sync(m, p) == full;
}
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) { int x = (a >> (8 * k)) & 255, y = (b >> (8 * k)) & 255; r |= (unsigned)(x > y ? x - y : y - x) << (8 * k); }
    return r;
}
static inline unsigned __vsubus4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; ++k) { int x = (a >> (8 * k)) & 255, y = (b >> (8 * k)) & 255; r |= (unsigned)(x > y ? x - y : 0) << (8 * k); }
    return r;
}
static inline long long clock64() { return 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }

#define SIM_ATOMIC_OPS(T)                                                                                         \
    static inline T atomicAdd(T *a, T v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }                   \
    static inline T atomicSub(T *a, T v) { return __atomic_fetch_sub(a, v, __ATOMIC_SEQ_CST); }                   \
    static inline T atomicExch(T *a, T v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }                 \
    static inline T atomicOr(T *a, T v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }                     \
    static inline T atomicAnd(T *a, T v) { return __atomic_fetch_and(a, v, __ATOMIC_SEQ_CST); }                   \
    static inline T atomicCAS(T *a, T c, T v) { __atomic_compare_exchange_n(a, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return c; } \
    static inline T atomicMin(T *a, T v) { T o = __atomic_load_n(a, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; } \
    static inline T atomicMax(T *a, T v) { T o = __atomic_load_n(a, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
SIM_ATOMIC_OPS(unsigned)
SIM_ATOMIC_OPS(int)
SIM_ATOMIC_OPS(unsigned long long)
#undef SIM_ATOMIC_OPS

template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
