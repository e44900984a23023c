// listing.cuh — the exact order of `Cluster.indices` (colour, color_clusters/builder.rs:526-533: `to.indices.append(from.indices)`)
// and of binary `Cluster.points` (clusters.rs:330-343).  Every merge appends the loser's list to the winner's, and every
// pixel joining later is pushed at the end, so with the reconstruction tree of ids.cuh:
//     listing(leaf) = own(leaf)            listing(v) = listing(winner child) ++ listing(loser child) ++ own(v)
// where own(c) = the pixels attributed to node c (k_attribute), ascending.  Hence the position of pixel p inside its
// cluster = start(attr(p)) [+ both child areas if attr(p) is an event] + rank of p among own(attr(p)), and
// start(c) = sum over the path c -> root of delta, delta = 0 for a winner child, area(winner) for a loser child.
// The path sums are computed by in-place asynchronous pointer jumping on packed (ancestor, offset) pairs; the ranks by
// one stable radix sort of the pixel ids keyed by their block offset.
#pragma once
#include "records.cuh"

namespace vcb {

__device__ static __forceinline__ unsigned long long *pair_ptr(const Bufs &b, u32 node) {
    return (node & TAG) ? (unsigned long long *)&b.ls[node & ~TAG].pad[1] : (unsigned long long *)&b.ev[node].carL;
}
__device__ static __forceinline__ unsigned long long ld_pair(unsigned long long *p) {
#ifdef VCB_HOSTSIM
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#else
    return __ldcg(p);
#endif
}
__device__ static __forceinline__ void st_pair(unsigned long long *p, unsigned long long v) {
#ifdef VCB_HOSTSIM
    __atomic_store_n(p, v, __ATOMIC_RELAXED);
#else
    __stcg(p, v);
#endif
}

__device__ static __forceinline__ void list_init_node(const Bufs &b, u32 node) {
    const u32 v = node_par(b, node);
    u32 delta = 0;
    if (v != NONE) {
        const EvRec *e = b.ev + v;
        const bool up = e->childU == node;
        const bool lwin = (e->arrive & EV_LWIN) != 0;
        const bool winner = up ? !lwin : lwin;
        if (!winner) delta = up ? e->a_wl : e->b_wu;      // area of the winning sibling
    }
    st_pair(pair_ptr(b, node), ((unsigned long long)delta << 32) | v);
}

constexpr int LST_THREADS = 256;
// work items: [0, kt) = leaves (newlist), [kt, kt + ncand) = candidate events (effective ones only)
__device__ static __forceinline__ u32 list_node(const Bufs &b, u32 t, u32 kt) {
    if (t < kt) return b.leaflist[t] | TAG;
    const u32 p = b.cand[t - kt] & ~TAG;
    return (b.flags[p] & F_MSF) ? p : NONE;
}

__global__ void __launch_bounds__(LST_THREADS) k_list_init(Bufs b) {
    const u32 kt = ldcg(&b.ctr->lt), total = kt + ldcg(&b.ctr->ncand);
    const u32 stride = gridDim.x * LST_THREADS;
    for (u32 t = blockIdx.x * LST_THREADS + threadIdx.x; t < total; t += stride) {
        const u32 node = list_node(b, t, kt);
        if (node != NONE) list_init_node(b, node);
    }
}

__global__ void __launch_bounds__(LST_THREADS) k_list_jump(Bufs b) {
    const u32 kt = ldcg(&b.ctr->lt), total = kt + ldcg(&b.ctr->ncand);
    const u32 stride = gridDim.x * LST_THREADS;
    for (u32 t = blockIdx.x * LST_THREADS + threadIdx.x; t < total; t += stride) {
        const u32 node = list_node(b, t, kt);
        if (node == NONE) continue;
        unsigned long long *mine = pair_ptr(b, node);
        unsigned long long pr = ld_pair(mine);
        u32 j = (u32)pr, a = (u32)(pr >> 32);
        while (j != NONE) {
            const unsigned long long q = ld_pair(pair_ptr(b, j));   // (ancestor of j, offset of j below it)
            a += (u32)(q >> 32);
            j = (u32)q;
            st_pair(mine, ((unsigned long long)a << 32) | j);
        }
    }
}

// per image: pool base = total list length of the earlier images (tiny, sequential)
__global__ void k_pool_base(const vcb_cluster_rec *rec, const u32 *slot_base, u32 nimg, u32 *pool_base /* nimg + 1 */) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        u32 acc = 0;
        for (u32 i = 0; i < nimg; ++i) {
            pool_base[i] = acc;
            if (slot_base[i + 1] > slot_base[i]) {
                const vcb_cluster_rec *l = rec + slot_base[i + 1] - 1;
                acc += (u32)l->indices_off + l->indices_len;
            }
        }
        pool_base[nimg] = acc;
    }
}

__global__ void __launch_bounds__(LST_THREADS) k_list_keys(Dims d, Bufs b, const vcb_cluster_rec *rec, const u32 *slot_base,
                                                           const u32 *pool_base, int keep_key, u64 *keys, u32 *vals) {
    const u32 stride = gridDim.x * LST_THREADS;
    for (u32 g = blockIdx.x * LST_THREADS + threadIdx.x; g < d.N; g += stride) {
        const u32 img = g / d.n;
        const u32 f = b.flags[g];
        u64 key = ~0ull;
        if ((f & J_MASK) == J_NONE) {
            if (keep_key) key = pool_base[img];                              // keyed pixel -> clusters[0].indices: raster order
        } else {
            const u32 node = b.attr[g];
            u32 own = (u32)(ld_pair(pair_ptr(b, node)) >> 32);
            if (!(node & TAG)) own += b.ev[node].a_wl + b.ev[node].b_wu;
            const vcb_cluster_rec *r = rec + slot_base[img] + b.la[b.pc[g]].flabel;
            key = (u64)pool_base[img] + r->indices_off + own;
        }
        keys[g] = key;
        vals[g] = g - img * d.n;
    }
}

}  // namespace vcb
