// records.cuh — per-cluster records (north_star stage (3): area, bounding rect, colour sums), `clusters_output` for
// hierarchical == 0 (color_clusters/builder.rs:366-377) and ClustersView::to_color_image (container.rs:104-116).
#pragma once
#include "../../include/vcb200.h"
#include "ids.cuh"

namespace vcb {

constexpr int REC_THREADS = 256;

// per slot: zero, rect sentinels for the atomic min/max accumulation.  The array is written as a flat run of 16-byte words
// (a record is six of them: words 10-11 = rect left/top = INT_MAX, words 12-13 = rect right/bottom = -1).
static_assert(sizeof(vcb_cluster_rec) == 96, "record layout");
__global__ void __launch_bounds__(REC_THREADS) k_rec_init(vcb_cluster_rec *rec, u32 total) {
    uint4 *p = reinterpret_cast<uint4 *>(rec);
    const u64 nvec = (u64)total * 6, stride = (u64)gridDim.x * REC_THREADS;
    for (u64 i = (u64)blockIdx.x * REC_THREADS + threadIdx.x; i < nvec; i += stride) {
        const u32 q = (u32)(i % 6);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q == 2) { v.z = 0x7fffffffu; v.w = 0x7fffffffu; }
        else if (q == 3) { v.x = 0xffffffffu; v.y = 0xffffffffu; }
        p[i] = v;
    }
}

// per leaf: fold the provisional cluster's sums / rect into the slot of its final cluster (ColorSum::merge color.rs:278-284,
// BoundingRect::merge bound.rs:204-219; both order-independent)
__global__ void __launch_bounds__(REC_THREADS) k_rec_accum(Dims d, Bufs b, vcb_cluster_rec *rec, const u32 *slot_base) {
    const u32 lt = ldcg(&b.ctr->lt);
    const u32 stride = gridDim.x * REC_THREADS;
    for (u32 k = blockIdx.x * REC_THREADS + threadIdx.x; k < lt; k += stride) {
        const u32 p = b.leaflist[k];
        const LeafA a = b.la[p];
        vcb_cluster_rec *r = rec + slot_base[p / d.n] + a.flabel;
        const LeafS s = b.ls[p];
        const LeafR q = b.lr[p];
        if (a.parL == NONE) {
            // no effective event ever touched this leaf: it is the whole cluster and the only writer of its slot
            *reinterpret_cast<uint4 *>(&r->sum[0]) = make_uint4(s.sum[0], s.sum[1], s.sum[2], s.sum[3]);
            r->sum[4] = s.npx;
            *reinterpret_cast<int2 *>(&r->rect[0]) = make_int2(q.minx, q.miny);
            *reinterpret_cast<int2 *>(&r->rect[2]) = make_int2(q.maxx, q.maxy);
            continue;
        }
        if (s.sum[0]) atomicAdd(&r->sum[0], s.sum[0]);
        if (s.sum[1]) atomicAdd(&r->sum[1], s.sum[1]);
        if (s.sum[2]) atomicAdd(&r->sum[2], s.sum[2]);
        if (s.sum[3]) atomicAdd(&r->sum[3], s.sum[3]);
        atomicAdd(&r->sum[4], s.npx);
        atomicMin(&r->rect[0], q.minx); atomicMin(&r->rect[1], q.miny);
        atomicMax(&r->rect[2], q.maxx); atomicMax(&r->rect[3], q.maxy);
    }
}

// per image: clusters[0] receives the keyed pixels under KeyingAction::Keep (builder.rs:330-334)
__global__ void k_rec_key(Dims d, Bufs b, vcb_cluster_rec *rec, const u32 *slot_base) {
    const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= d.nimg) return;
    const ImgMeta m = b.meta[img];
    if (m.key_cnt == 0) return;
    vcb_cluster_rec *r = rec + slot_base[img];
    for (int k = 0; k < 4; ++k) r->sum[k] = m.key_sum[k];
    r->sum[4] = m.key_cnt;
    r->rect[0] = m.key_rect[0]; r->rect[1] = m.key_rect[2]; r->rect[2] = m.key_rect[1]; r->rect[3] = m.key_rect[3];
}

// stage-1 finalisation fused into the offset scan: half-open rect, residue_sum = sum (prepare_stage_2, builder.rs:380-382),
// indices_len = area, indices_off = exclusive scan (rebased per image afterwards)
struct OffScanFinal {
    vcb_cluster_rec *rec; u32 total;
    __device__ u32 count() const { return total; }
    __device__ u32 value(u32 s) const { return rec[s].sum[4]; }
    __device__ void emit(u32 s, u32 ex, u32 v) const {
        vcb_cluster_rec *r = rec + s;
        if (v == 0) { r->rect[0] = r->rect[1] = r->rect[2] = r->rect[3] = 0; }
        else { r->rect[2] += 1; r->rect[3] += 1; }
        for (int k = 0; k < 5; ++k) r->residue_sum[k] = r->sum[k];
        r->indices_len = v;
        r->indices_off = ex;
    }
};

// total area of the images before image i (= global exclusive prefix at the image's first slot): the prefix of the 2048-slot
// chunk the slot lies in plus the areas between the chunk start and the slot.  One warp per image.
__global__ void __launch_bounds__(REC_THREADS) k_img_base(const vcb_cluster_rec *rec, const u32 *slot_base, u32 nimg,
                                                          const u32 *chunkpref, u32 *img_base) {
    const u32 img = (blockIdx.x * REC_THREADS + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (img >= nimg) return;                               // whole warps
    const u32 s1 = slot_base[img], c = s1 / (u32)prims::SCAN_CHUNK;
    u32 acc = 0;
    for (u32 s = c * (u32)prims::SCAN_CHUNK + lane; s < s1; s += 32) acc += rec[s].sum[4];
#pragma unroll
    for (int dlt = 16; dlt; dlt >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, dlt);
    if (lane == 0) img_base[img] = chunkpref[c] + acc;      // c == number of chunks is the grand total (k_scan_chunks)
}

// Last pass of the stage-1 record scan (OffScanFinal::emit for every slot, rebased per image) with the records moved
// through shared memory: 256 records = 1536 16-byte words are loaded and stored fully coalesced, each thread edits its own.
constexpr int RF_WORDS = 24;
constexpr size_t RF_SMEM = (size_t)REC_THREADS * RF_WORDS * 4 + 64;
// With keys != nullptr the sort keys of stage_1_output (builder.rs:366-377: slots with area > 0, stable sort by
// area * 65535 + index) and the per-image live counts are produced in the same pass.
__global__ void __launch_bounds__(REC_THREADS) k_rec_finalize(vcb_cluster_rec *rec, u32 total, const u32 *chunkpref,
                                                              const u32 *slot_img, const u32 *img_base, const u32 *slot_base,
                                                              int key_bits, u64 *keys, u32 *vals, u32 *live) {
    VCB_SMEM(u32, sm);
    u32 *sh = sm + REC_THREADS * RF_WORDS;
    uint4 *sm4 = reinterpret_cast<uint4 *>(sm);
    const u32 nchunks = (total + prims::SCAN_CHUNK - 1) / prims::SCAN_CHUNK;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        u32 carry = chunkpref[c];
        for (u32 s0 = c * (u32)prims::SCAN_CHUNK; s0 < (c + 1) * (u32)prims::SCAN_CHUNK && s0 < total; s0 += REC_THREADS) {
            const u32 cnt = total - s0 < (u32)REC_THREADS ? total - s0 : (u32)REC_THREADS;
            uint4 *g4 = reinterpret_cast<uint4 *>(rec + s0);
            for (u32 i = threadIdx.x; i < cnt * 6; i += REC_THREADS) sm4[i] = g4[i];
            __syncthreads();
            u32 *r = sm + threadIdx.x * RF_WORDS;
            const bool mine = threadIdx.x < cnt;
            const u32 v = mine ? r[4] : 0u;
            u32 tot;
            const u32 ex = prims::block_exscan(v, sh, &tot) + carry;
            u32 img = 0xffffffffu;
            if (mine) {
                if (v == 0) { r[10] = r[11] = r[12] = r[13] = 0; }
                else { r[12] += 1; r[13] += 1; }
#pragma unroll
                for (int k = 0; k < 5; ++k) r[5 + k] = r[k];
                r[17] = v;
                img = slot_img[s0 + threadIdx.x];
                r[18] = ex - img_base[img]; r[19] = 0;
            }
            if (keys) {                                  // uniform; stage_1_output key (builder.rs:366-377)
                const u32 s = s0 + threadIdx.x;
                if (mine) {
                    const u32 idx = s - slot_base[img];
                    const u64 key = v ? (u64)v * 65535ull + idx : ((1ull << key_bits) - 1ull);
                    keys[s] = ((u64)img << key_bits) | key;
                    vals[s] = idx;
                }
                const int lane = threadIdx.x & 31;
                const bool alive = mine && v != 0;
                const u32 img0 = __shfl_sync(0xffffffffu, img, 0);
                const bool uniform = __all_sync(0xffffffffu, img == img0 || !mine);
                const u32 m = __ballot_sync(0xffffffffu, alive);
                if (uniform) { if (lane == 0 && m && img0 != 0xffffffffu) atomicAdd(&live[img0], (u32)__popc(m)); }
                else if (alive) atomicAdd(&live[img], 1u);
            }
            __syncthreads();
            for (u32 i = threadIdx.x; i < cnt * 6; i += REC_THREADS) g4[i] = sm4[i];
            __syncthreads();
            carry += tot;
        }
    }
}

// indices_off: exclusive scan of indices_len inside each image; live count per image
struct OffScan {
    vcb_cluster_rec *rec; const u32 *slot_base; const u32 *slot_img; u32 total; u64 *img_off;
    __device__ u32 count() const { return total; }
    __device__ u32 value(u32 s) const { return rec[s].indices_len; }
    __device__ void emit(u32 s, u32 ex, u32) const { rec[s].indices_off = ex; }   // global; rebased by k_off_rebase
};
__global__ void __launch_bounds__(REC_THREADS) k_off_rebase(vcb_cluster_rec *rec, const u32 *slot_base, const u32 *slot_img, u32 total) {
    const u32 stride = gridDim.x * REC_THREADS;
    for (u32 s = blockIdx.x * REC_THREADS + threadIdx.x; s < total; s += stride) {
        const u32 first = slot_base[slot_img[s]];
        if (s != first) rec[s].indices_off -= rec[first].indices_off;
    }
    // the first slot of every image is rebased last, by k_off_rebase0
}
__global__ void k_off_rebase0(vcb_cluster_rec *rec, const u32 *slot_base, u32 nimg) {
    const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img < nimg && slot_base[img] < slot_base[img + 1]) rec[slot_base[img]].indices_off = 0;   // an image may own no slot at all
}
__global__ void __launch_bounds__(REC_THREADS) k_slot_img(u32 *slot_img, const u32 *slot_base, u32 nimg) {
    // slot_img[s] = image owning slot s
    const u32 img = blockIdx.x;
    if (img >= nimg) return;
    for (u32 s = slot_base[img] + threadIdx.x; s < slot_base[img + 1]; s += REC_THREADS) slot_img[s] = img;
}

// ClustersView::to_color_image for hierarchical == 0: every pixel belongs to exactly one output cluster (its own);
// residue_color = residue_sum.average() (cluster.rs:42-44, color.rs:295-302).  Keyed pixels under Discard stay 0.
// `nx` (sequential resolver only, else null): pixels marked 0xFFFFFFFE are in no cluster's list and stay unpainted.
__global__ void __launch_bounds__(REC_THREADS) k_color_image_flat(u32 n, const u32 *labels, const vcb_cluster_rec *rec, u32 *out, const u32 *nx) {
    const u32 stride = gridDim.x * REC_THREADS;
    for (u32 g = blockIdx.x * REC_THREADS + threadIdx.x; g < n; g += stride) {
        const vcb_cluster_rec *r = rec + labels[g];
        u32 c = 0;
        const u32 cnt = r->residue_sum[4];
        if (r->indices_len && cnt && !(nx && nx[g] == 0xFFFFFFFEu))
            c = ((r->residue_sum[0] / cnt) & 255u) | (((r->residue_sum[1] / cnt) & 255u) << 8) |
                (((r->residue_sum[2] / cnt) & 255u) << 16) | (((r->residue_sum[3] / cnt) & 255u) << 24);
        out[g] = c;
    }
}

}  // namespace vcb
