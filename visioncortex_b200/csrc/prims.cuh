// prims.cuh — hand-written device-wide primitives used by the clustering pipeline:
//   * ordered exclusive scan / stream compaction over a functor-defined domain whose length may live in device memory
//     (north_star stage (2): "canonical relabeling ... using a prefix-scan compaction");
//   * stable LSD radix sort of (u64 key, u32 value) pairs (reference call sites: the stable `sort_by_key` of
//     color_clusters/builder.rs:375 and :452, and the ordered `indices` / `points` lists).
// All kernels are plain sm_100a CUDA; no library (CUB/Thrust) calls.
#pragma once
#include "rt.h"

namespace prims {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;   // 2048

// Exclusive scan of one value per thread across a 256-thread block.  `sh` needs 9 u32.  Returns the exclusive prefix;
// *total receives the block sum.
template <class T> __device__ static __forceinline__ T block_exscan(T v, T *sh, T *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (SCAN_THREADS / 32) ? sh[lane] : 0;
        T winc = w;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < (SCAN_THREADS / 32)) sh[lane] = winc - w;
        if (lane == (SCAN_THREADS / 32) - 1) sh[8] = winc;
    }
    __syncthreads();
    T res = sh[warp] + inc - v;
    *total = sh[8];
    __syncthreads();
    return res;
}

// optional functor methods: value_first(i) in the summing pass, value_again(i) in the applying pass (default: value(i))
template <class F> __device__ static __forceinline__ auto scan_value_first(const F &f, u32 i, int) -> decltype(f.value_first(i)) { return f.value_first(i); }
template <class F> __device__ static __forceinline__ u32 scan_value_first(const F &f, u32 i, long) { return f.value(i); }
template <class F> __device__ static __forceinline__ auto scan_value_again(const F &f, u32 i, int) -> decltype(f.value_again(i)) { return f.value_again(i); }
template <class F> __device__ static __forceinline__ u32 scan_value_again(const F &f, u32 i, long) { return f.value(i); }

// ---------------------------------------------------------------------------------------------------------------
// Functor protocol for scans:  u32 F::count() const  (number of items; may read device memory)
//                               u32 F::value(u32 i) const
//                               void F::emit(u32 i, u32 exclusive_prefix, u32 value) const
// The scan is global over [0, count).  chunk c covers items [c*2048, (c+1)*2048).
template <class F> __global__ void k_scan_sums(F f, u32 *chunksum) {
    VCB_SMEM(u32, sh);
    const u32 cnt = f.count();
    const u32 nchunks = (cnt + SCAN_CHUNK - 1) / SCAN_CHUNK;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const u32 i0 = c * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (i0 + k < cnt) s += scan_value_first(f, i0 + k, 0);
        u32 total;
        block_exscan(s, sh, &total);
        if (threadIdx.x == 0) chunksum[c] = total;
    }
}

// single block: exclusive scan of chunksum[0..nchunks) in place; chunksum[nchunks] receives the grand total.
template <class F> __global__ void k_scan_chunks(F f, u32 *chunksum) {
    VCB_SMEM(u32, sh);
    const u32 cnt = f.count();
    const u32 nchunks = (cnt + SCAN_CHUNK - 1) / SCAN_CHUNK;
    u32 carry = 0;
    for (u32 base = 0; base < nchunks; base += SCAN_THREADS) {
        const u32 i = base + threadIdx.x;
        const u32 v = i < nchunks ? chunksum[i] : 0;
        u32 total;
        const u32 ex = block_exscan(v, sh, &total);
        if (i < nchunks) chunksum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) chunksum[nchunks] = carry;
}

template <class F> __global__ void k_scan_apply(F f, const u32 *chunkpref) {
    VCB_SMEM(u32, sh);
    const u32 cnt = f.count();
    const u32 nchunks = (cnt + SCAN_CHUNK - 1) / SCAN_CHUNK;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const u32 i0 = c * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
        u32 v[SCAN_ITEMS];
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (i0 + k < cnt) ? scan_value_again(f, i0 + k, 0) : 0;
            s += v[k];
        }
        u32 total;
        u32 ex = block_exscan(s, sh, &total) + chunkpref[c];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (i0 + k < cnt) f.emit(i0 + k, ex, v[k]);
            ex += v[k];
        }
    }
}

// Host driver.  `chunksum` must hold max_items/2048 + 2 entries.  After the call chunksum[nchunks] = total (on device).
template <class F> inline void scan(rt::Stream &st, const char *name, F f, u32 *chunksum, u32 max_items, int grid_cap) {
    u32 maxchunks = (max_items + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (maxchunks == 0) maxchunks = 1;
    const u32 grid = maxchunks < (u32)grid_cap ? maxchunks : (u32)grid_cap;
    rt::launch(st, name, rt::THR, k_scan_sums<F>, dim3(grid), dim3(SCAN_THREADS), 64, f, chunksum);
    rt::launch(st, name, rt::THR, k_scan_chunks<F>, dim3(1), dim3(SCAN_THREADS), 64, f, chunksum);
    rt::launch(st, name, rt::THR, k_scan_apply<F>, dim3(grid), dim3(SCAN_THREADS), 64, f, chunksum);
}

// first two passes only: chunksum[c] = exclusive prefix of chunk c, chunksum[nchunks] = total; the caller applies them
template <class F> inline void scan_prefixes(rt::Stream &st, const char *name, F f, u32 *chunksum, u32 max_items, int grid_cap) {
    u32 maxchunks = (max_items + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (maxchunks == 0) maxchunks = 1;
    const u32 grid = maxchunks < (u32)grid_cap ? maxchunks : (u32)grid_cap;
    rt::launch(st, name, rt::THR, k_scan_sums<F>, dim3(grid), dim3(SCAN_THREADS), 64, f, chunksum);
    rt::launch(st, name, rt::THR, k_scan_chunks<F>, dim3(1), dim3(SCAN_THREADS), 64, f, chunksum);
}

// ---------------------------------------------------------------------------------------------------------------
// Stable LSD radix sort, 8-bit digits.  Item count is read from device memory (*countp).
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_CHUNK = RS_THREADS * RS_ROUNDS;   // 2048

struct RsArgs {
    const u64 *kin; const u32 *vin; u64 *kout; u32 *vout;
    const u32 *countp; u32 *table;   // table[256 * nchunks] (+1)
    int shift;
};

__global__ void k_rs_hist(RsArgs a) {
    VCB_SMEM(u32, hist);
    const u32 cnt = *a.countp;
    const u32 nchunks = (cnt + RS_CHUNK - 1) / RS_CHUNK;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        for (int r = 0; r < RS_ROUNDS; ++r) {
            const u32 i = c * RS_CHUNK + r * RS_THREADS + threadIdx.x;
            if (i < cnt) atomicAdd(&hist[(u32)(a.kin[i] >> a.shift) & 255u], 1u);
        }
        __syncthreads();
        a.table[threadIdx.x * nchunks + c] = hist[threadIdx.x];
        __syncthreads();
    }
}

struct RsTableScan {
    const u32 *countp; u32 *table;
    __device__ u32 count() const { const u32 cnt = *countp; return 256u * ((cnt + RS_CHUNK - 1) / RS_CHUNK); }
    __device__ u32 value(u32 i) const { return table[i]; }
    __device__ void emit(u32 i, u32 ex, u32) const { table[i] = ex; }
};

__global__ void k_rs_scatter(RsArgs a) {
    VCB_SMEM(u32, sh);                 // running[256] + warpcnt[8][256]
    u32 *running = sh;
    u32 *warpcnt = sh + 256;
    const u32 cnt = *a.countp;
    const u32 nchunks = (cnt + RS_CHUNK - 1) / RS_CHUNK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        running[threadIdx.x] = a.table[threadIdx.x * nchunks + c];
        for (int r = 0; r < RS_ROUNDS; ++r) {
            for (int k = 0; k < 8; ++k) warpcnt[k * 256 + threadIdx.x] = 0;
            __syncthreads();
            const u32 i = c * RS_CHUNK + r * RS_THREADS + threadIdx.x;
            const bool ok = i < cnt;
            u64 key = 0; u32 val = 0; u32 d = 0;
            if (ok) { key = a.kin[i]; val = a.vin ? a.vin[i] : 0; d = (u32)(key >> a.shift) & 255u; }
            // lanes with the same digit (invalid lanes excluded)
            u32 m = __ballot_sync(0xffffffffu, ok);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const u32 bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                m &= ((d >> b) & 1u) ? bal : ~bal;
            }
            const u32 rank = __popc(m & ((1u << lane) - 1u));
            if (ok && rank == 0) warpcnt[warp * 256 + d] = __popc(m);
            __syncthreads();
            u32 off = 0;
            if (ok) {
                off = running[d];
                for (int w2 = 0; w2 < warp; ++w2) off += warpcnt[w2 * 256 + d];
                a.kout[off + rank] = key;
                if (a.vout) a.vout[off + rank] = val;
            }
            __syncthreads();
            {   // advance running[] by this round's totals
                u32 t = 0;
                for (int w2 = 0; w2 < 8; ++w2) t += warpcnt[w2 * 256 + threadIdx.x];
                running[threadIdx.x] += t;
            }
            __syncthreads();
        }
    }
}

// Sorts by key bits [0, nbits).  Buffers ping-pong; returns true if the result ended in the "out" buffers (k1/v1).
// table: 256 * (max_items/2048 + 1) + 1 entries; chunksum: scan scratch for 256*(max_items/2048+1) items.
inline bool radix_sort(rt::Stream &st, u64 *k0, u32 *v0, u64 *k1, u32 *v1, const u32 *countp, u32 max_items, int nbits,
                       u32 *table, u32 *chunksum, int grid_cap) {
    u32 maxchunks = (max_items + RS_CHUNK - 1) / RS_CHUNK;
    if (maxchunks == 0) maxchunks = 1;
    const u32 grid = maxchunks < (u32)grid_cap ? maxchunks : (u32)grid_cap;
    bool flipped = false;
    for (int shift = 0; shift < nbits; shift += 8) {
        RsArgs a{flipped ? k1 : k0, flipped ? v1 : v0, flipped ? k0 : k1, flipped ? v0 : v1, countp, table, shift};
        rt::launch(st, "rs_hist", rt::THR, k_rs_hist, dim3(grid), dim3(RS_THREADS), 256 * 4, a);
        RsTableScan ts{countp, table};
        scan(st, "rs_scan", ts, chunksum, 256u * maxchunks, grid_cap);
        rt::launch(st, "rs_scatter", rt::THR, k_rs_scatter, dim3(grid), dim3(RS_THREADS), (256 + 8 * 256) * 4, a);
        flipped = !flipped;
    }
    return flipped;
}

}  // namespace prims
