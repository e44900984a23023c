// queries.cuh — device-resident `Cluster` queries on a finished result: perimeter, neighbours and
// to_image_with_hole computed from what the result already keeps in HBM (stage-1 labels, final labels, the stage-2 merge
// forest, the records), so a caller that stays on the GPU between clustering and tracing needs no indices / holes pools on
// the host.  Reference: color_clusters/cluster.rs:46-80 (perimeter, to_image_with_hole), :132-165 (neighbours),
// shape/geometry.rs:52-75 (image_boundary_list).
//
// Membership without the pools.  After the hierarchical stage `Cluster.indices` of X holds every pixel whose merge chain
// (stage-1 slot -> mpar -> mpar ...) passes through X, provided X kept its pixels (a plain merge moves them out:
// combine_clusters, builder.rs:526-539; a deepen merge keeps a clone: combine_clusters_clone :514-524); `Cluster.holes` of X
// holds the pixels whose hop INTO X was a hollow deepen merge (merge_cluster_into, builder.rs:495-512).  So a pixel is set in
// X.to_image_with_hole(w, hole) iff X is on its chain and (not hole, or the hop into X was not hollow).  Without a
// hierarchical stage the chain is the label itself.
#pragma once
#include "stage2.cuh"

namespace vcb {

struct QArgs {
    Dims d;                        // one image: nimg = 1
    const u32 *l1;                 // stage-1 labels of the image (image-local slot ids)
    const u32 *fin;                // final labels of the image
    const u32 *mpar;               // merge forest of the image's slots (null: no hierarchical stage)
    const vcb_cluster_rec *rec;    // the image's records
    u32 nslots;
};

// is X on the chain that starts at stage-1 slot s?  *hollow = the hop into X was a hollow deepen merge
__device__ static __forceinline__ bool q_on_chain(const QArgs &q, u32 s, u32 X, bool *hollow) {
    bool h = false;
    while (true) {
        if (s == X) { *hollow = h; return true; }
        if (!q.mpar) return false;
        const u32 m = q.mpar[s];
        if (m == NONE) return false;
        h = (m & HOLLOWBIT) != 0;
        s = m & MPAR_MASK;
    }
}
__device__ static __forceinline__ bool q_set_in(const QArgs &q, u32 s, u32 X) {      // pixel of slot s set in X.to_image (holes removed)
    bool h;
    return q_on_chain(q, s, X, &h) && !h;
}

// Cluster::perimeter of EVERY cluster that holds pixels, in one pass over the image: perim[X] = number of pixels set in
// X.to_image(parent) with a 4-neighbour that is not (cluster.rs:46-48, shape/geometry.rs:52-75; outside the rect = outside the
// image of the cluster = not set).  A pixel whose four neighbours carry its own stage-1 label is interior to every cluster on
// its chain and is skipped after five loads.
__global__ void __launch_bounds__(256) k_q_perimeters(QArgs q, u32 *perim) {
    const Dims d = q.d;
    const u32 stride = gridDim.x * 256;
    for (u32 p = blockIdx.x * 256 + threadIdx.x; p < d.n; p += stride) {
        const u32 x = p % d.w, y = p / d.w;
        const u32 s = q.l1[p];
        const u32 nb[4] = {x > 0 ? q.l1[p - 1] : NONE, x + 1 < d.w ? q.l1[p + 1] : NONE, y > 0 ? q.l1[p - d.w] : NONE, y + 1 < d.h ? q.l1[p + d.w] : NONE};
        if (nb[0] == s && nb[1] == s && nb[2] == s && nb[3] == s) continue;
        u32 c = s;
        bool hollow_in = false;
        while (true) {
            if (q.rec[c].indices_len != 0 && !hollow_in) {                    // p is set in c's image
                bool edge = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) edge = edge || nb[k] == NONE || (nb[k] != s && !q_set_in(q, nb[k], c));
                if (edge) atomicAdd(&perim[c], 1u);
            }
            if (!q.mpar) break;
            const u32 m = q.mpar[c];
            if (m == NONE) break;
            hollow_in = (m & HOLLOWBIT) != 0;
            c = m & MPAR_MASK;
        }
    }
}

// Cluster::neighbours (cluster.rs:132-165) of cluster X: final labels of the 4-neighbours of X's pixels (holes included:
// the reference iterates `indices`), without ZERO and without `myself` = the final label of X's pixels.  flags[] is per slot.
__global__ void __launch_bounds__(256) k_q_neighbours(QArgs q, u32 X, i32 left, i32 top, u32 rw, u32 rh, u32 *flags) {
    const Dims d = q.d;
    const u32 stride = gridDim.x * 256, cells = rw * rh;
    for (u32 t = blockIdx.x * 256 + threadIdx.x; t < cells; t += stride) {
        const u32 x = (u32)left + t % rw, y = (u32)top + t / rw;
        const u32 p = y * d.w + x;
        bool h;
        if (!q_on_chain(q, q.l1[p], X, &h)) continue;
        const u32 myself = q.fin[p];
        const u32 nl[4] = {y > 0 ? q.fin[p - d.w] : 0u, y + 1 < d.h ? q.fin[p + d.w] : 0u, x > 0 ? q.fin[p - 1] : 0u, x + 1 < d.w ? q.fin[p + 1] : 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) if (nl[k] != 0u && nl[k] != myself) flags[nl[k]] = 1u;
    }
}
// ordered compaction of the flagged slots (one block; the list is short)
__global__ void __launch_bounds__(256) k_q_compact(const u32 *flags, u32 nslots, u32 *out, u32 cap, u32 *count) {
    VCB_SMEM(u32, sh);                                           // 257 words (dynamic)
    u32 carry = 0;
    for (u32 s0 = 0; s0 < nslots; s0 += 256) {
        const u32 s = s0 + threadIdx.x;
        const u32 f = s < nslots && flags[s] ? 1u : 0u;
        sh[threadIdx.x + 1] = f;
        if (threadIdx.x == 0) sh[0] = 0;
        __syncthreads();
        if (threadIdx.x == 0) for (int k = 1; k <= 256; ++k) sh[k] += sh[k - 1];
        __syncthreads();
        if (f) { const u32 pos = carry + sh[threadIdx.x]; if (pos < cap) out[pos] = s; }
        carry += sh[256];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

// Cluster::to_image_with_hole(parent.width, hole) (cluster.rs:60-80) as BitVec storage (u32 blocks, pixel i of the
// rect-sized image at block i / 32, bit i % 32): one warp per 32 pixels
__global__ void __launch_bounds__(256) k_q_to_image(QArgs q, u32 X, i32 left, i32 top, u32 rw, u32 rh, int hole, u32 *bits) {
    const Dims d = q.d;
    const u32 cells = rw * rh, nwords = (cells + 31) / 32;
    const u32 lane = threadIdx.x & 31;
    for (u32 wd = (blockIdx.x * 256 + threadIdx.x) >> 5; wd < nwords; wd += (gridDim.x * 256) >> 5) {
        const u32 t = wd * 32 + lane;
        bool set = false;
        if (t < cells) {
            const u32 x = (u32)left + t % rw, y = (u32)top + t / rw;
            bool h;
            set = q_on_chain(q, q.l1[y * d.w + x], X, &h) && !(hole && h);
        }
        const u32 m = __ballot_sync(0xffffffffu, set);
        if (lane == 0) bits[wd] = m;
    }
}

// The images of MANY clusters at once (row f1: Cluster::to_compound_path starts with to_image_with_hole(...).to_clusters(false),
// cluster.rs:108-128): item k = cluster slots[k] with rect rects[4k..], written as a Wp x Hp bit image (pixels outside the
// rect unset) of words_per_img words -- the layout the batched binary pipeline reads.  One warp per 32 pixels.
__global__ void __launch_bounds__(256) k_q_to_image_batch(QArgs q, const u32 *slots, const i32 *rects, u32 nitems, u32 Wp, u32 Hp,
                                                          u32 words_per_img, int hole, u32 *bits) {
    const Dims d = q.d;
    const u32 lane = threadIdx.x & 31;
    const u64 nwords = (u64)nitems * words_per_img;
    for (u64 wd = ((u64)blockIdx.x * 256 + threadIdx.x) >> 5; wd < nwords; wd += ((u64)gridDim.x * 256) >> 5) {
        const u32 k = (u32)(wd / words_per_img), t = (u32)(wd - (u64)k * words_per_img) * 32 + lane;
        bool set = false;
        if (t < Wp * Hp) {
            const u32 lx = t % Wp, ly = t / Wp;
            const i32 left = rects[4 * k], top = rects[4 * k + 1], right = rects[4 * k + 2], bottom = rects[4 * k + 3];
            if ((i32)lx < right - left && (i32)ly < bottom - top) {
                bool h;
                set = q_on_chain(q, q.l1[((u32)top + ly) * d.w + (u32)left + lx], slots[k], &h) && !(hole && h);
            }
        }
        const u32 m = __ballot_sync(0xffffffffu, set);
        if (lane == 0) bits[wd] = m;
    }
}

}  // namespace vcb
