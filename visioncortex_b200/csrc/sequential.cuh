// sequential.cuh — device-side sequential resolver for the inputs whose stage-1 result depends on the reference's scan
// order in ways the parallel id model (ids.cuh) does not cover:
//   * "unclean" key colours: a non-key pixel is color_same to the key, so label 0 (clusters[0], builder.rs:44-45,189) takes
//     part in joins and merges, is emptied by combine_clusters and filled again;
//   * diagonal = true inputs where a slot freed by builder.rs:315-317 is resurrected through the stale cluster_upleft label
//     and then overwritten by the next new cluster (builder.rs:301-305, 341-348; Appendix E item 5, second half), or whose
//     stale-upleft fixpoint does not settle.
// The parallel pipeline detects these on the device (Counters.err); the affected device batch is then replayed here: one
// CTA per image whose thread 0 restates builder.rs:273-364 literally on device memory -- cluster_indices is the primary
// state, clusters are (list head, list tail, area, ColorSum, BoundingRect) records per slot, `indices` are linked lists
// threaded through one word per pixel, combine_clusters relabels by walking the absorbed list (builder.rs:526-539).
// Images are independent, so a batch still runs one image per SM; a 1080p image takes on the order of a second.  This is
// a product path on the GPU (no host fallback); it only exists so that every input has an exact answer.
#pragma once
#include "listing.cuh"

namespace vcb {

struct SeqSlot {
    u32 head, tail, area, pad;
    u32 sum[4];                    // ColorSum r, g, b, a (counter == area); u32 wrapping (color.rs:261-284)
    i32 rect[4];                   // BoundingRect left, top, right, bottom, half open (bound.rs:15-21)
};
static_assert(sizeof(SeqSlot) == 48, "SeqSlot layout");

__device__ static __forceinline__ bool seq_rect_empty(const i32 *r) { return r[2] - r[0] == 0 && r[3] - r[1] == 0; }   // bound.rs:101-103

// Cluster::add (cluster.rs:24-28): indices.push(i); sum.add(color); rect.add_x_y(x, y)
__device__ static __forceinline__ void seq_add(SeqSlot *cl, u32 *nx, u32 s, u32 i, u32 color, i32 x, i32 y) {
    SeqSlot &c = cl[s];
    nx[i] = NONE;
    if (c.area == 0) c.head = i; else nx[c.tail] = i;
    c.tail = i;
    c.area += 1;
    c.sum[0] += color & 255u; c.sum[1] += (color >> 8) & 255u; c.sum[2] += (color >> 16) & 255u; c.sum[3] += color >> 24;
    if (seq_rect_empty(c.rect)) { c.rect[0] = x; c.rect[2] = x + 1; c.rect[1] = y; c.rect[3] = y + 1; }   // bound.rs:172-190
    else {
        if (x < c.rect[0]) c.rect[0] = x; else if (x + 1 > c.rect[2]) c.rect[2] = x + 1;
        if (y < c.rect[1]) c.rect[1] = y; else if (y + 1 > c.rect[3]) c.rect[3] = y + 1;
    }
}

// combine_clusters (builder.rs:526-539)
__device__ static __forceinline__ void seq_combine(SeqSlot *cl, u32 *nx, u32 *lab, u32 from, u32 to) {
    SeqSlot &f = cl[from];
    SeqSlot &t = cl[to];
    if (f.area) {
        for (u32 p = f.head; p != NONE; p = nx[p]) lab[p] = to;               // :527-529
        if (t.area == 0) t.head = f.head; else nx[t.tail] = f.head;           // :531-532 to.indices.append(from.indices)
        t.tail = f.tail;
        t.area += f.area;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) t.sum[k] += f.sum[k];                         // :535 ColorSum::merge
    if (!seq_rect_empty(f.rect)) {                                            // :536 BoundingRect::merge (bound.rs:204-219)
        if (seq_rect_empty(t.rect)) { for (int k = 0; k < 4; ++k) t.rect[k] = f.rect[k]; }
        else {
            t.rect[0] = t.rect[0] < f.rect[0] ? t.rect[0] : f.rect[0]; t.rect[2] = t.rect[2] > f.rect[2] ? t.rect[2] : f.rect[2];
            t.rect[1] = t.rect[1] < f.rect[1] ? t.rect[1] : f.rect[1]; t.rect[3] = t.rect[3] > f.rect[3] ? t.rect[3] : f.rect[3];
        }
    }
    f.head = f.tail = NONE; f.area = 0;                                       // mem::take, :537-538 clear()
#pragma unroll
    for (int k = 0; k < 4; ++k) { f.sum[k] = 0; f.rect[k] = 0; }
}

constexpr int SEQ_THREADS = 32;
// nx[] of a pixel that is in no cluster's list: a keyed pixel under KeyingAction::Discard (builder.rs:333) and the pixels of a
// resurrected slot that a new cluster overwrote (builder.rs:347-348).  ClustersView::to_color_image walks the lists
// (container.rs:107-113), so these pixels stay unpainted whatever their label says.
constexpr u32 SEQ_ORPHAN = 0xFFFFFFFEu;
// One CTA per image of the group [img0, img0 + gridDim.x); thread 0 replays builder.rs:273-364.  `cl` holds n + 1 slot
// records per image of the group, `nx` / `labels` one word per pixel of the batch.
__global__ void __launch_bounds__(SEQ_THREADS) k_seq_stage1(Dims d, ColorPolicy pol, int keep_key, u32 *labels, u32 *nx_all, SeqSlot *cl_all,
                                                            ImgMeta *meta, u32 img0) {
    if (threadIdx.x != 0) return;
    const u32 img = img0 + blockIdx.x;
    if (img >= d.nimg) return;
    const u32 *px = pol.rgba + (size_t)img * d.n;
    u32 *lab = labels + (size_t)img * d.n;
    u32 *nx = nx_all + (size_t)img * d.n;
    SeqSlot *cl = cl_all + (size_t)blockIdx.x * ((size_t)d.n + 1);
    const SeqSlot fresh = {NONE, NONE, 0u, 0u, {0u, 0u, 0u, 0u}, {0, 0, 0, 0}};
    cl[0] = fresh;                                           // clusters = vec![Cluster::new()] (builder.rs:189)
    u32 len = 1, next_index = 1;                             // next_index: builder.rs:190
    u32 orphans = 0;                                         // overwritten resurrected slots (labels and lists disagree afterwards)
    const bool has_key = pol.has_key != 0, diagonal = pol.diagonal != 0;
    const u32 key = pol.key, w = d.w;
    u32 i = 0;
    for (u32 y = 0; y < d.h; ++y) {
        for (u32 x = 0; x < w; ++x, ++i) {
            const bool vu = y > 0, vl = x > 0, vul = vu && vl;                // pixel_at -> None outside (builder.rs:549-556)
            const u32 color = px[i];
            const u32 up = vu ? px[i - w] : 0u, left = vl ? px[i - 1] : 0u, upleft = vul ? px[i - w - 1] : 0u;
            u32 cluster_up = vu ? lab[i - w] : 0u;                            // :291-305
            u32 cluster_left = vl ? lab[i - 1] : 0u;
            const u32 cluster_upleft = vul ? lab[i - w - 1] : 0u;
            const bool s_cu = vu && pol.same(color, up), s_cl = vl && pol.same(color, left), s_cul = vul && pol.same(color, upleft);
            if (cluster_left != cluster_up && vul && pol.same(left, up) && (diagonal || (s_cl && s_cu))) {   // :307-311
                if (cl[cluster_left].area <= cl[cluster_up].area) {           // :313-325
                    seq_combine(cl, nx, lab, cluster_left, cluster_up);
                    if (cluster_left == next_index - 1 && next_index == len) next_index -= 1;
                    cluster_left = cluster_up;
                } else {
                    seq_combine(cl, nx, lab, cluster_up, cluster_left);
                    cluster_up = cluster_left;
                }
            }
            if (has_key && color == key) {                                    // :330-334
                lab[i] = 0u;                                                  // cluster_indices keeps its initial ZERO
                if (keep_key) seq_add(cl, nx, 0u, i, color, (i32)x, (i32)y);
                else nx[i] = SEQ_ORPHAN;
            } else if (s_cu && s_cul) {                                       // :335-337
                lab[i] = cluster_up; seq_add(cl, nx, cluster_up, i, color, (i32)x, (i32)y);
            } else if (s_cl && s_cul) {                                       // :338-340
                lab[i] = cluster_left; seq_add(cl, nx, cluster_left, i, color, (i32)x, (i32)y);
            } else if (diagonal && s_cul) {                                   // :341-343 (cluster_upleft was read before the merge)
                lab[i] = cluster_upleft; seq_add(cl, nx, cluster_upleft, i, color, (i32)x, (i32)y);
            } else {                                                          // :344-353
                if (next_index < len && cl[next_index].area) {                // a resurrected slot is overwritten: its pixels keep their label
                    for (u32 p = cl[next_index].head; p != NONE;) { const u32 q = nx[p]; nx[p] = SEQ_ORPHAN; p = q; }
                    ++orphans;
                }
                cl[next_index] = fresh;                                       // overwrite a freed slot or push
                if (next_index >= len) len = next_index + 1;
                seq_add(cl, nx, next_index, i, color, (i32)x, (i32)y);
                lab[i] = next_index;
                next_index += 1;
            }
        }
    }
    meta[img].slots = len;
    meta[img].pad[0] = orphans;
}

// Discarded key pixels get the INERT stage-1 label when a hierarchical stage follows (they are label 0 in the reference but in
// no cluster's list; clusters[0] has members of its own here, and stage 2 is built on labels == membership)
__global__ void __launch_bounds__(256) k_seq_inert(Dims d, ColorPolicy pol, u32 *labels) {
    const u32 stride = gridDim.x * 256;
    for (u32 g = blockIdx.x * 256 + threadIdx.x; g < d.N; g += stride)
        if (pol.rgba[g] == pol.key) labels[g] = 0xFFFFFFFEu;               // == INERT (stage2.cuh)
}

// slot records of the replay -> the accumulation format the common record tail expects (k_rec_init / k_rec_finalize:
// sum[0..3], sum[4] = pixel count, rect = inclusive min / max; an empty slot keeps the k_rec_init sentinels)
__global__ void __launch_bounds__(256) k_seq_rec_fill(Dims d, const SeqSlot *cl_all, vcb_cluster_rec *rec, const u32 *slot_base, u32 img0, u32 nimg_grp) {
    const u32 stride = gridDim.x * 256;
    for (u32 k = 0; k < nimg_grp; ++k) {
        const u32 img = img0 + k;
        const u32 s0 = slot_base[img], ns = slot_base[img + 1] - s0;
        const SeqSlot *cl = cl_all + (size_t)k * ((size_t)d.n + 1);
        for (u32 s = blockIdx.x * 256 + threadIdx.x; s < ns; s += stride) {
            const SeqSlot c = cl[s];
            if (c.area == 0) continue;
            vcb_cluster_rec *r = rec + s0 + s;
            for (int q = 0; q < 4; ++q) r->sum[q] = c.sum[q];
            r->sum[4] = c.area;
            r->rect[0] = c.rect[0]; r->rect[1] = c.rect[1]; r->rect[2] = c.rect[2] - 1; r->rect[3] = c.rect[3] - 1;
        }
    }
}

// Cluster.indices of every slot: the linked list in order, written at the slot's offset of the image's pool
__global__ void __launch_bounds__(256) k_seq_pool(Dims d, const SeqSlot *cl_all, const u32 *nx_all, const vcb_cluster_rec *rec, const u32 *slot_base,
                                                  const u32 *pool_base, u32 *pool, u32 img0, u32 nimg_grp) {
    const u32 stride = gridDim.x * 256;
    for (u32 k = 0; k < nimg_grp; ++k) {
        const u32 img = img0 + k;
        const u32 s0 = slot_base[img], ns = slot_base[img + 1] - s0;
        const SeqSlot *cl = cl_all + (size_t)k * ((size_t)d.n + 1);
        const u32 *nx = nx_all + (size_t)img * d.n;
        for (u32 s = blockIdx.x * 256 + threadIdx.x; s < ns; s += stride) {
            if (cl[s].area == 0) continue;
            u32 *out = pool + pool_base[img] + (u32)rec[s0 + s].indices_off;
            u32 q = 0;
            for (u32 p = cl[s].head; p != NONE; p = nx[p]) out[q++] = p;
        }
    }
}

}  // namespace vcb
