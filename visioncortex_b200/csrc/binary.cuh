// binary.cuh — result shaping for BinaryImage::to_clusters (clusters.rs:345-348): the non-empty slots in slot order,
// each with its rect and its points (listing order), plus the overall rect of all set pixels (clusters.rs:302).
#pragma once
#include "listing.cuh"

namespace vcb {

struct BinCompact {
    const vcb_cluster_rec *rec; u32 total; vcb_bin_cluster_rec *out; i32 *overall /* minx, miny, maxx, maxy */; u32 *ncl;
    __device__ u32 count() const { return total; }
    __device__ u32 value(u32 s) const { return rec[s].indices_len ? 1u : 0u; }
    __device__ void emit(u32 s, u32 ex, u32 v) const {
        if (v) {
            const vcb_cluster_rec &r = rec[s];
            vcb_bin_cluster_rec o;
            o.rect[0] = r.rect[0]; o.rect[1] = r.rect[1]; o.rect[2] = r.rect[2]; o.rect[3] = r.rect[3];
            o.points_off = r.indices_off; o.points_len = r.indices_len; o._pad = 0;
            out[ex] = o;
            atomicMin(&overall[0], r.rect[0]); atomicMin(&overall[1], r.rect[1]);
            atomicMax(&overall[2], r.rect[2]); atomicMax(&overall[3], r.rect[3]);
        }
        if (s == total - 1) *ncl = ex + v;
    }
};

__global__ void k_bin_points(const u32 *pool, u32 total, u32 w, i32 *xy) {
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const u32 p = pool[i];
        xy[2 * i] = (i32)(p % w);
        xy[2 * i + 1] = (i32)(p / w);
    }
}

}  // namespace vcb
