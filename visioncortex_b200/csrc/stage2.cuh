// stage2.cuh — the hierarchical stage of color_clusters::Builder (builder.rs:379-539).
//
// The reference pops clusters in (current area, index) order; each pop merges the cluster into the neighbour with the
// closest average colour (or outputs it).  The pop order is reproduced exactly:
//   * pixels are never relabelled: a per-slot forwarding table (union-find with a "hollow" flag on the last hop)
//     maps stage-1 labels to current clusters; the region adjacency graph is built once from the stage-1 labels;
//   * pops are grouped in epochs e = floor(log2(area)); the clusters of an epoch are sorted on the device
//     (prims::radix_sort) and one CTA per image replays them in order; a merge creates a key with a larger area, which
//     belongs to a later epoch or (rarely) to the small in-epoch "urgent" list;
//   * inside a pop the neighbour scan (cluster.rs:132-171), the arg-min over `diff * 65535 + index`
//     (builder.rs:452-454) and the perimeter of patch_good (runner.rs:131-146, shape/geometry.rs:52-75) are
//     data-parallel over the CTA.
// Images of a batch are independent and run on different SMs.
#pragma once
#include "listing.cuh"

namespace vcb {

struct S2Cfg {
    u32 hierarchical; u64 good_min, good_max; i32 deepen_diff; u64 hollow_neighbours; int can_discard; int hier_max;
    u64 help_min_px;   // rects at least this large are counted by the whole thread-block cluster
};

struct S2ImgState {
    u32 stamp, nout, maxarea, nover;
    u32 nmerge, task_c, task_quit, perim_acc;
    i32 task_rect[4];
};

struct S2 {
    Dims d; S2Cfg cfg;
    const u32 *l1;                 // stage-1 labels (image-local slot ids), N
    vcb_cluster_rec *rec;          // total_slots
    const u32 *slot_base;          // nimg + 1
    const u32 *slot_img;           // total_slots
    u32 total_slots;
    u32 *fw;                       // total_slots: current cluster of the slot | hollow << 31 (roots: own index)
    u32 *mpar;                     // total_slots: merge parent (to | deepen << 31 | hollow-deepen << 30), NONE = never merged
    u32 *childoff, *holeoff;       // offset of the absorbed list inside to.indices / to.holes at merge time
    u32 *mnext, *mtail;            // member lists (stage-1 slots of a cluster)
    u32 *mark;                     // distinct-neighbour stamps
    u32 *outpos;                   // first position in clusters_output, NONE if absent
    u32 *adj_off;                  // total_slots + 1
    u32 *adj;                      // adjacency entries (image-local slot ids)
    u32 *cursor;                   // total_slots (fill cursors / degree counts)
    u32 *outputs;                  // total_slots: clusters_output per image at slot_base[img]
    u64 *over;                     // total_slots: per-image overflow of the urgent list at slot_base[img]
    S2ImgState *ist;               // nimg
    u64 *keys; u32 *vals;          // epoch sort input/output
    u32 *count;                    // number of keys collected this epoch
    int idx_bits, area_bits;
};

constexpr u32 FLAGBIT = 0x80000000u;
constexpr u32 HOLLOWBIT = 0x40000000u;
constexpr u32 MPAR_MASK = 0x3fffffffu;

// Every slot points DIRECTLY at its current cluster (roots point at themselves): a merge rewrites the pointers of all
// members of the absorbed cluster (the counterpart of the reference's relabel loop, builder.rs:527-529), so resolving a
// stage-1 label is one load.  Bit 31 = "the hop into the root was a hollow deepen merge" (the pixel is in root.holes).
__device__ static __forceinline__ u32 s2_find(const u32 *fw, u32 base, u32 s, u32 *flag) {
    const u32 v = ldcg(&fw[base + s]);
    *flag = v >> 31;
    return v & ~FLAGBIT;
}

// ------------------------------------------------------------------------------------------------- RAG
// pairs of different stage-1 labels on horizontally / vertically adjacent pixels; a pair is skipped when the previous
// pixel along the border carries the same pair (most duplicates of straight borders)
template <bool FILL> __global__ void __launch_bounds__(256) k_rag(S2 a) {
    const Dims d = a.d;
    const u32 stride = gridDim.x * 256;
    for (u32 g = blockIdx.x * 256 + threadIdx.x; g < d.N; g += stride) {
        const u32 img = g / d.n, p = g - img * d.n, x = p % d.w, y = p / d.w;
        const u32 base = a.slot_base[img];
        const u32 la = a.l1[g];
        if (x + 1 < d.w) {
            const u32 lb = a.l1[g + 1];
            if (la != lb && !(y > 0 && a.l1[g - d.w] == la && a.l1[g + 1 - d.w] == lb)) {
                if (FILL) {
                    a.adj[a.adj_off[base + la] + atomicAdd(&a.cursor[base + la], 1u)] = lb;
                    a.adj[a.adj_off[base + lb] + atomicAdd(&a.cursor[base + lb], 1u)] = la;
                } else { atomicAdd(&a.cursor[base + la], 1u); atomicAdd(&a.cursor[base + lb], 1u); }
            }
        }
        if (y + 1 < d.h) {
            const u32 lb = a.l1[g + d.w];
            if (la != lb && !(x > 0 && a.l1[g - 1] == la && a.l1[g + d.w - 1] == lb)) {
                if (FILL) {
                    a.adj[a.adj_off[base + la] + atomicAdd(&a.cursor[base + la], 1u)] = lb;
                    a.adj[a.adj_off[base + lb] + atomicAdd(&a.cursor[base + lb], 1u)] = la;
                } else { atomicAdd(&a.cursor[base + la], 1u); atomicAdd(&a.cursor[base + lb], 1u); }
            }
        }
    }
}

struct DegScan {
    S2 a;
    __device__ u32 count() const { return a.total_slots; }
    __device__ u32 value(u32 s) const { return a.cursor[s]; }
    __device__ void emit(u32 s, u32 ex, u32 v) const {
        a.adj_off[s] = ex;
        a.cursor[s] = 0;
        if (s == a.total_slots - 1) a.adj_off[a.total_slots] = ex + v;
    }
};

__global__ void __launch_bounds__(256) k_s2_init(S2 a) {
    const u32 stride = gridDim.x * 256;
    for (u32 s = blockIdx.x * 256 + threadIdx.x; s < a.total_slots; s += stride) {
        a.fw[s] = s - a.slot_base[a.slot_img[s]]; a.mpar[s] = NONE; a.mnext[s] = NONE; a.mtail[s] = s - a.slot_base[a.slot_img[s]];
        a.mark[s] = 0; a.outpos[s] = NONE; a.cursor[s] = 0;
        const u32 area = a.rec[s].indices_len;
        if (area) atomicMax(&a.ist[a.slot_img[s]].maxarea, area);
    }
}
__global__ void k_s2_init_img(S2 a) {
    const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= a.d.nimg) return;
    S2ImgState z{};
    a.ist[img] = z;
}

// ------------------------------------------------------------------------------------------------- epochs
__global__ void __launch_bounds__(256) k_s2_collect(S2 a, int epoch) {
    const u32 stride = gridDim.x * 256;
    const int lane = threadIdx.x & 31;
    const u32 nloop = (a.total_slots + stride - 1) / stride;
    for (u32 it = 0; it < nloop; ++it) {
        const u32 s = it * stride + blockIdx.x * 256 + threadIdx.x;
        bool take = false;
        u64 key = 0; u32 idx = 0;
        if (s < a.total_slots && a.fw[s] == s - a.slot_base[a.slot_img[s]]) {
            const u32 area = a.rec[s].indices_len;
            if (area && (31 - __clz((int)area)) == epoch) {
                const u32 img = a.slot_img[s];
                idx = s - a.slot_base[img];
                key = ((u64)img << (a.idx_bits + epoch)) | ((u64)(area - (1u << epoch)) << a.idx_bits) | idx;   // area has `epoch` free bits
                take = true;
            }
        }
        const u32 m = __ballot_sync(0xffffffffu, take);
        if (m) {
            u32 pos = 0;
            const int leader = __ffs((int)m) - 1;
            if (lane == leader) pos = atomicAdd(a.count, (u32)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, leader);
            if (take) { pos += __popc(m & ((1u << lane) - 1u)); a.keys[pos] = key; a.vals[pos] = idx; }
        }
    }
}

__device__ static __forceinline__ u32 avg_color(const u32 *sum) {   // ColorSum::average (color.rs:295-302), `as u8`
    const u32 c = sum[4];
    return ((sum[0] / c) & 255u) | (((sum[1] / c) & 255u) << 8) | (((sum[2] / c) & 255u) << 16) | (((sum[3] / c) & 255u) << 24);
}
__device__ static __forceinline__ i32 color_diff(u32 a, u32 b) {     // runner.rs:110-114
    i32 dsum = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const i32 x = (i32)((a >> (8 * c)) & 255u) - (i32)((b >> (8 * c)) & 255u);
        dsum += x < 0 ? -x : x;
    }
    return dsum;
}

constexpr int S2_MAX_THREADS = 1024;   // block size is chosen at launch: 1024 for a few large images, 128 for big batches
constexpr int S2_MEMCHUNK = 256;
constexpr int S2_URGENT = 64;
constexpr u32 S2_MAPBYTES = 40960;

struct S2Shared {
    unsigned long long best;
    unsigned long long urgent[S2_URGENT];
    u32 mem[S2_MEMCHUNK];
    u32 eoff[S2_MEMCHUNK];
    u32 deg[S2_MEMCHUNK + 1];
    u32 cur, area, ndist, perim, nmem, more, mycolor, nurg, lo, hi, stamp, nchunks, newfw, target, want_perim, deepen, single;
    i32 rect[4];
    u8 map[S2_MAPBYTES];
};

// Perimeter of cluster c restricted to rect rows [row_lo, row_hi) (rect-relative), shape/geometry.rs:52-75 on
// Cluster::to_image_with_hole: pixels of c (holes removed) with a 4-neighbour outside c.  Block-parallel; the membership
// of the rows (+1 halo) is staged band by band into a shared byte map; pixels outside the rect are outside the cluster by
// definition (the rect bounds it), pixels outside the image likewise.  Adds the block's count to sh->perim.
__device__ static __forceinline__ void s2_perimeter_rows(const S2 &a, S2Shared *sh, u32 img, u32 base, u32 c, const i32 *rect,
                                                         u32 row_lo, u32 row_hi) {
    const Dims d = a.d;
    const u32 tid = threadIdx.x, S2_THREADS = blockDim.x;
    const i32 L = rect[0], T = rect[1], R = rect[2], B = rect[3];
    const u32 rw = (u32)(R - L);
    const u32 *l1 = a.l1 + (size_t)img * d.n;
    u32 cntp = 0;
    const u32 pw = rw + 2;
    u32 band = S2_MAPBYTES / pw;
    if (band < 3) band = 3;
    const u32 inner = band - 2;                                   // rows counted per band
    for (u32 r0 = row_lo; r0 < row_hi; r0 += inner) {
        const u32 rows = (row_hi - r0 < inner) ? row_hi - r0 : inner;
        const u32 cells = (rows + 2) * pw;
        if (cells > S2_MAPBYTES) {                                // rect wider than the map: direct (slow) evaluation
            for (u64 q = tid; q < (u64)rw * rows; q += S2_THREADS) {
                const u32 x = (u32)L + (u32)(q % rw), y = (u32)T + r0 + (u32)(q / rw);
                u32 fl;
                if (s2_find(a.fw, base, l1[y * d.w + x], &fl) != c || fl) continue;
                bool edge = (x == 0) || (y == 0) || (x + 1 == d.w) || (y + 1 == d.h);
                if (!edge) {
                    const i32 off[4] = {-1, 1, -(i32)d.w, (i32)d.w};
#pragma unroll
                    for (int k2 = 0; k2 < 4; ++k2) {
                        u32 f2;
                        if (s2_find(a.fw, base, l1[(i32)(y * d.w + x) + off[k2]], &f2) != c || f2) { edge = true; break; }
                    }
                }
                if (edge) cntp++;
            }
            continue;
        }
        __syncthreads();
        for (u32 q0 = tid; q0 < cells; q0 += 4 * S2_THREADS) {
            // four independent cells per step: the label loads, then the forwarding loads, are all in flight together
            u32 lab[4]; bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const u32 q = q0 + u * S2_THREADS;
                const i32 mx = (i32)(q % pw) - 1, my = (i32)(q / pw) - 1;    // rect-relative
                const i32 x = L + mx, y = T + (i32)r0 + my;
                ok[u] = q < cells && mx >= 0 && mx < (i32)rw && y >= T && y < B;
                lab[u] = ok[u] ? __ldg(&l1[(u32)y * d.w + (u32)x]) : 0u;
            }
            u32 fwv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) fwv[u] = ok[u] ? ldcg(&a.fw[base + lab[u]]) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const u32 q = q0 + u * S2_THREADS;
                if (q < cells) sh->map[q] = (ok[u] && fwv[u] == c) ? 1 : 0;   // == c: root is c and the hollow flag is clear
            }
        }
        __syncthreads();
        for (u32 q = tid; q < rw * rows; q += S2_THREADS) {
            const u32 mx = q % rw + 1, my = q / rw + 1;
            const u32 idx = my * pw + mx;
            if (sh->map[idx] && !(sh->map[idx - 1] && sh->map[idx + 1] && sh->map[idx - pw] && sh->map[idx + pw])) cntp++;
        }
    }
    if (cntp) atomicAdd(&sh->perim, cntp);
}

__global__ void __launch_bounds__(S2_MAX_THREADS) k_s2_epoch(S2 a, const u64 *keys, int epoch, u32 G) {
    const u32 S2_THREADS = blockDim.x;
    VCB_SMEM(S2Shared, sh);
    const u32 img = blockIdx.x / G, crank = blockIdx.x % G, base = a.slot_base[img];   // G CTAs (one cluster) per image
    const u32 nslots = a.slot_base[img + 1] - base;
    const u32 tid = threadIdx.x;
    S2ImgState *ist = a.ist + img;
    const int kshift = a.idx_bits + epoch;                      // sorted keys: image | (area - 2^epoch) | index
    const u64 idx_mask = (1ull << a.idx_bits) - 1ull;
    const u32 area_base = 1u << epoch;
    u64 *over = a.over + base;

    if (crank != 0) {
        // helper CTA of the image's cluster: sleeps in the cluster barrier; woken for a perimeter task or to quit
        while (true) {
            VCB_CLUSTER_SYNC();
            if (ldcg(&ist->task_quit)) break;
            if (tid == 0) sh->perim = 0;
            i32 rect[4];
            for (int k2 = 0; k2 < 4; ++k2) rect[k2] = ldcg(&ist->task_rect[k2]);
            const u32 tc = ldcg(&ist->task_c);
            __syncthreads();
            const u32 rh = (u32)(rect[3] - rect[1]);
            s2_perimeter_rows(a, sh, img, base, tc, rect, (u32)((u64)rh * crank / G), (u32)((u64)rh * (crank + 1) / G));
            __syncthreads();
            if (tid == 0) { if (sh->perim) atomicAdd(&ist->perim_acc, sh->perim); __threadfence(); }
            __syncthreads();
            VCB_CLUSTER_SYNC();
        }
        return;
    }

    if (tid == 0) {
        const u32 cnt = *a.count;
        u32 lo = 0, hi = cnt;                       // first key of this image
        while (lo < hi) { const u32 mid = (lo + hi) / 2; if ((u32)(keys[mid] >> kshift) < img) lo = mid + 1; else hi = mid; }
        sh->lo = lo;
        hi = cnt;
        u32 l2 = lo;
        while (l2 < hi) { const u32 mid = (l2 + hi) / 2; if ((u32)(keys[mid] >> kshift) <= img) l2 = mid + 1; else hi = mid; }
        sh->hi = l2;
        sh->nurg = 0;
    }
    __syncthreads();

    while (true) {
        if (tid == 0) {
            // next pop = smallest valid key among: sorted segment head, urgent list, overflow list
            u32 c = NONE, carea = 0;
            while (true) {
                u64 kb = ~0ull; int src = -1; u32 which = 0;
                if (sh->lo < sh->hi) { kb = (keys[sh->lo] & ((1ull << kshift) - 1ull)) + ((u64)area_base << a.idx_bits); src = 0; }
                for (u32 i = 0; i < sh->nurg; ++i) if (sh->urgent[i] < kb) { kb = sh->urgent[i]; src = 1; which = i; }
                for (u32 i = 0; i < ist->nover; ++i) if (over[i] < kb) { kb = over[i]; src = 2; which = i; }
                if (src < 0) break;
                if (src == 0) sh->lo++;
                else if (src == 1) { sh->urgent[which] = sh->urgent[--sh->nurg]; }
                else { over[which] = over[--ist->nover]; }
                const u32 idx = (u32)(kb & idx_mask), ar = (u32)(kb >> a.idx_bits);   // kb = (area << idx_bits) | index
                if (a.fw[base + idx] == idx && a.rec[base + idx].indices_len == ar) { c = idx; carea = ar; break; }
            }
            sh->cur = c; sh->area = carea;
            if (c != NONE) {
                sh->best = ~0ull; sh->ndist = 0; sh->perim = 0;
                sh->stamp = ++ist->stamp;
                sh->mycolor = avg_color(a.rec[base + c].sum);
                sh->more = c; sh->nchunks = 0;
                // fast path for a cluster that never absorbed anything (most pops): its member list is itself, so the
                // adjacency range is loaded here and the gather / degree-prefix rounds are skipped
                sh->single = 0;
                if (a.mnext[base + c] == NONE) {
                    const u32 e0 = a.adj_off[base + c], e1 = a.adj_off[base + c + 1];
                    sh->mem[0] = c; sh->eoff[0] = e0; sh->deg[0] = 0; sh->deg[1] = e1 - e0;
                    sh->nmem = 1; sh->more = NONE; sh->nchunks = 1; sh->single = 1;
                }
            }
        }
        __syncthreads();
        const u32 c = sh->cur;
        if (c == NONE) break;
        const u32 area = sh->area;
        if (area > a.cfg.hierarchical) {                              // builder.rs:429-432
            if (tid == 0) {
                const u32 pos = ist->nout++;
                a.outputs[base + pos] = c;
                if (a.outpos[base + c] == NONE) a.outpos[base + c] = pos;
            }
            __syncthreads();
            continue;
        }
        // ---- neighbours of c (cluster.rs:132-171) over the member lists and the static adjacency
        const u32 stamp = sh->stamp, mycolor = sh->mycolor;
        const bool single = sh->single != 0;                          // uniform (written before the barrier above)
        while (true) {
            if (!single) {
                if (tid == 0) {
                    u32 n = 0, m = sh->more;
                    while (m != NONE && n < S2_MEMCHUNK) { sh->mem[n++] = m; m = a.mnext[base + m]; }
                    sh->nmem = n; sh->more = m; sh->nchunks++;
                }
                __syncthreads();
                // edge-parallel: prefix of the member degrees in shared memory, then one (member, edge) item per thread step
                for (u32 i = tid; i < sh->nmem; i += S2_THREADS) {
                    const u32 s = sh->mem[i];
                    sh->eoff[i] = a.adj_off[base + s];
                    sh->deg[i] = a.adj_off[base + s + 1] - sh->eoff[i];
                }
                __syncthreads();
                if (tid == 0) { const u32 nm0 = sh->nmem; u32 acc = 0; for (u32 i = 0; i < nm0; ++i) { const u32 dg = sh->deg[i]; sh->deg[i] = acc; acc += dg; } sh->deg[nm0] = acc; }
                __syncthreads();
            }
            const u32 nm = sh->nmem;
            const u32 nitems = sh->deg[nm];
            for (u32 it = tid; it < nitems; it += S2_THREADS) {
                u32 lo = 0, hi = nm;                                  // last member with prefix <= it
                while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (sh->deg[mid] <= it) lo = mid; else hi = mid; }
                const u32 e = sh->eoff[lo] + (it - sh->deg[lo]);
                u32 fl;
                const u32 o = s2_find(a.fw, base, a.adj[e], &fl);
                if (o == c || o == 0) continue;                     // label 0 and self are not neighbours (cluster.rs:150)
                if (atomicExch(&a.mark[base + o], stamp) != stamp) atomicAdd(&sh->ndist, 1u);
                const i32 df = color_diff(mycolor, avg_color(a.rec[base + o].sum));
                const unsigned long long k = ((unsigned long long)((u32)df * 65535u + o) << 32) | o;
                atomicMin(&sh->best, k);
            }
            __syncthreads();
            const bool last = sh->more == NONE;
            __syncthreads();                                          // thread 0 rewrites sh->more in the next round
            if (last) break;
        }
        const u32 ndist = sh->ndist;
        if (ndist == 0) {                                             // builder.rs:444-450
            if (tid == 0 && (area == ist->maxarea || a.cfg.can_discard)) {
                const u32 pos = ist->nout++;
                a.outputs[base + pos] = c;
                if (a.outpos[base + c] == NONE) a.outpos[base + c] = pos;
            }
            __syncthreads();
            continue;
        }
        // thread 0 alone reads the target's record here: it goes on to modify it while the others may still be behind
        if (tid == 0) {
            const u32 tg = (u32)sh->best;
            const i32 md = (i32)(((u32)(sh->best >> 32) - tg) / 65535u);   // key = diff * 65535 + index (builder.rs:452)
            sh->target = tg;
            sh->want_perim = 0; sh->deepen = 0;
            if (a.cfg.hier_max && a.cfg.good_min < (u64)area && (u64)area < a.cfg.good_max && md > a.cfg.deepen_diff) {
                if (a.cfg.good_min == 0) sh->deepen = 1; else sh->want_perim = 1;
            }
            const vcb_cluster_rec *rc = a.rec + base + c;
            sh->rect[0] = rc->rect[0]; sh->rect[1] = rc->rect[1]; sh->rect[2] = rc->rect[2]; sh->rect[3] = rc->rect[3];
        }
        __syncthreads();
        const u32 target = sh->target;
        bool deepen = sh->deepen != 0;
        if (sh->want_perim) {
            const u32 rh = (u32)(sh->rect[3] - sh->rect[1]), rwid = (u32)(sh->rect[2] - sh->rect[0]);
            if (G > 1 && (u64)rh * rwid >= a.cfg.help_min_px) {
                // large rect: the other CTAs of the cluster take row slices.  They sleep in the cluster barrier until now.
                if (tid == 0) {
                    stcg(&ist->task_c, c); stcg(&ist->task_quit, 0u);
                    for (int k2 = 0; k2 < 4; ++k2) stcg(&ist->task_rect[k2], sh->rect[k2]);
                    __threadfence();
                }
                __syncthreads();
                VCB_CLUSTER_SYNC();                                   // task published (relabel stores of earlier pops included)
                s2_perimeter_rows(a, sh, img, base, c, sh->rect, 0, rh / G);
                __syncthreads();
                VCB_CLUSTER_SYNC();                                   // all slices counted into ist->perim_acc
                if (tid == 0) { sh->perim += ldcg(&ist->perim_acc); stcg(&ist->perim_acc, 0u); }
                __syncthreads();
            } else {
                s2_perimeter_rows(a, sh, img, base, c, sh->rect, 0, rh);
                __syncthreads();
            }
            deepen = (u64)sh->perim < (u64)area;
        }
        const bool hollow = (u64)ndist <= a.cfg.hollow_neighbours;
        if (tid == 0) {
            vcb_cluster_rec *rf = a.rec + base + c, *rt_ = a.rec + base + target;
            if (deepen) {                                             // builder.rs:463-465
                const u32 pos = ist->nout++;
                a.outputs[base + pos] = c;
                if (a.outpos[base + c] == NONE) a.outpos[base + c] = pos;
            }
            // merge_cluster_into (builder.rs:495-539)
            if (!deepen) for (int k2 = 0; k2 < 5; ++k2) rt_->residue_sum[k2] += rf->residue_sum[k2];
            for (int k2 = 0; k2 < 5; ++k2) rt_->sum[k2] += rf->sum[k2];
            if (!(rf->rect[2] - rf->rect[0] == 0 && rf->rect[3] - rf->rect[1] == 0)) {                 // bound.rs:204-219
                if (rt_->rect[2] - rt_->rect[0] == 0 && rt_->rect[3] - rt_->rect[1] == 0) {
                    for (int k2 = 0; k2 < 4; ++k2) rt_->rect[k2] = rf->rect[k2];
                } else {
                    rt_->rect[0] = rt_->rect[0] < rf->rect[0] ? rt_->rect[0] : rf->rect[0];
                    rt_->rect[1] = rt_->rect[1] < rf->rect[1] ? rt_->rect[1] : rf->rect[1];
                    rt_->rect[2] = rt_->rect[2] > rf->rect[2] ? rt_->rect[2] : rf->rect[2];
                    rt_->rect[3] = rt_->rect[3] > rf->rect[3] ? rt_->rect[3] : rf->rect[3];
                }
            }
            a.childoff[base + c] = rt_->indices_len;                 // builder.rs:529: to.indices.append(from.indices)
            a.holeoff[base + c] = rt_->holes_len;
            rt_->indices_len += area;
            if (!deepen) {
                for (int k2 = 0; k2 < 5; ++k2) rf->sum[k2] = 0;
                for (int k2 = 0; k2 < 4; ++k2) rf->rect[k2] = 0;
                rf->indices_len = 0;
            } else {
                if (hollow) { rt_->holes_len += area; rt_->num_holes += 1; }
                rf->merged_into = target;
                rt_->depth += 1;
            }
            sh->newfw = target | ((deepen && hollow) ? FLAGBIT : 0u);
            a.mpar[base + c] = target | (deepen ? FLAGBIT : 0u) | ((deepen && hollow) ? HOLLOWBIT : 0u);
            a.mnext[base + a.mtail[base + target]] = c;
            a.mtail[base + target] = a.mtail[base + c];
            ist->nmerge++;
            const u32 na = rt_->indices_len;
            if (na > ist->maxarea) ist->maxarea = na;
            if ((31 - __clz((int)na)) == epoch) {                     // same epoch: must still be popped in order
                const u64 k = ((u64)na << a.idx_bits) | target;
                if (sh->nurg < S2_URGENT) sh->urgent[sh->nurg++] = k; else over[ist->nover++] = k;
            }
            if (sh->nchunks > 1) sh->more = c;                        // the member list no longer fits sh->mem: walk it again
        }
        __syncthreads();
        // relabel: every member slot of c now resolves to the target (builder.rs:527-529)
        {
            const u32 nf = sh->newfw;
            if (sh->nchunks <= 1) {
                const u32 nm = sh->nmem;
                for (u32 i = tid; i < nm; i += S2_THREADS) stcg(&a.fw[base + sh->mem[i]], nf);
            } else {
                // c's list is c -> ... -> old tail; it was appended to the target's list, so stop at the old tail of c
                const u32 stop = a.mtail[base + target];              // == old tail of c
                while (true) {
                    __syncthreads();
                    if (tid == 0) {
                        u32 n = 0, m = sh->more;
                        while (m != NONE && n < S2_MEMCHUNK) { sh->mem[n++] = m; m = (m == stop) ? NONE : a.mnext[base + m]; }
                        sh->nmem = n; sh->more = m;
                    }
                    __syncthreads();
                    const u32 nm = sh->nmem;
                    for (u32 i = tid; i < nm; i += S2_THREADS) stcg(&a.fw[base + sh->mem[i]], nf);
                    __syncthreads();
                    const bool last = sh->more == NONE;
                    __syncthreads();
                    if (last) break;
                }
            }
        }
        __syncthreads();
    }
    (void)nslots;
    if (G > 1) {                                                      // release the helper CTAs
        if (tid == 0) { stcg(&ist->task_quit, 1u); __threadfence(); }
        __syncthreads();
        VCB_CLUSTER_SYNC();
    }
}

// ------------------------------------------------------------------------------------------------- finalisation
__global__ void __launch_bounds__(256) k_s2_labels(S2 a, u32 *labels) {
    const Dims d = a.d;
    const u32 stride = gridDim.x * 256;
    for (u32 g = blockIdx.x * 256 + threadIdx.x; g < d.N; g += stride) {
        const u32 base = a.slot_base[g / d.n];
        labels[g] = a.fw[base + a.l1[g]] & ~FLAGBIT;
    }
}

struct HolesScan {
    vcb_cluster_rec *rec; u32 total;
    __device__ u32 count() const { return total; }
    __device__ u32 value(u32 s) const { return rec[s].holes_len; }
    __device__ void emit(u32 s, u32 ex, u32) const { rec[s].holes_off = ex; }
};
__global__ void __launch_bounds__(256) k_holes_rebase(vcb_cluster_rec *rec, const u32 *slot_base, const u32 *slot_img, u32 total, u32 *first_off) {
    const u32 stride = gridDim.x * 256;
    for (u32 s = blockIdx.x * 256 + threadIdx.x; s < total; s += stride) rec[s].holes_off -= first_off[slot_img[s]];
}
__global__ void k_holes_first(const vcb_cluster_rec *rec, const u32 *slot_base, u32 nimg, u32 *first_off) {
    const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img < nimg) first_off[img] = (u32)rec[slot_base[img]].holes_off;
}

// ClustersView::to_color_image (container.rs:104-116): clusters_output painted in reverse with residue_color, so a pixel
// ends with the colour of the FIRST output cluster whose `indices` contain it.  The clusters containing the pixels of a
// stage-1 slot are the slot itself and its merge ancestors reached through the chain, restricted to those that still
// own their indices (alive roots and deepened clusters, builder.rs:514-524).
__global__ void __launch_bounds__(256) k_s2_slot_color(S2 a, u32 *slot_color) {
    const u32 stride = gridDim.x * 256;
    for (u32 s = blockIdx.x * 256 + threadIdx.x; s < a.total_slots; s += stride) {
        const u32 base = a.slot_base[a.slot_img[s]];
        u32 x = s - base, best = NONE, bestx = NONE;
        while (true) {
            const u32 mp = a.mpar[base + x];
            const bool owns = (mp == NONE) || (mp & FLAGBIT);
            if (owns) { const u32 op = a.outpos[base + x]; if (op < best) { best = op; bestx = x; } }
            if (mp == NONE) break;
            x = mp & MPAR_MASK;
        }
        u32 col = 0;
        if (bestx != NONE) {
            const vcb_cluster_rec *r = a.rec + base + bestx;
            if (r->residue_sum[4]) col = avg_color(r->residue_sum);
        }
        slot_color[s] = col;
    }
}
__global__ void __launch_bounds__(256) k_s2_paint(S2 a, const u32 *slot_color, u32 img, u32 *out) {
    const u32 stride = gridDim.x * 256;
    const u32 base = a.slot_base[img];
    for (u32 p = blockIdx.x * 256 + threadIdx.x; p < a.d.n; p += stride) out[p] = slot_color[base + a.l1[(size_t)img * a.d.n + p]];
}

// Ordered `Cluster.indices` / `Cluster.holes` after the hierarchical stage.  A merge appends the absorbed cluster's whole
// list to the target's (builder.rs:526-533), a deepened cluster also keeps its own copy (builder.rs:514-524) and a hollow
// deepen merge appends the same list to target.holes (builder.rs:503-507).  So every list is a contiguous run of the
// merge forest's preorder, and a stage-1 pixel with rank r in its stage-1 list lands at offset r + sum(childoff) in every
// owning ancestor.  One thread per entry of the stage-1 pool walks its merge chain.
__global__ void __launch_bounds__(256) k_s2_pool_scatter(S2 a, const u32 *pool1, const u32 *pool1_base, const u32 *off1,
                                                         const u32 *ipool_base, const u32 *hpool_base, u32 *out_i, u32 *out_h) {
    const Dims d = a.d;
    const u32 total = pool1_base[d.nimg];
    const u32 stride = gridDim.x * 256;
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
        u32 lo = 0, hi = d.nimg;                                      // image owning pool entry i
        while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (pool1_base[mid] <= i) lo = mid; else hi = mid; }
        const u32 img = lo, base = a.slot_base[img];
        const u32 p = pool1[i];
        u32 x = a.l1[(size_t)img * d.n + p];
        u32 rel = (i - pool1_base[img]) - off1[base + x];
        while (true) {
            const u32 mp = a.mpar[base + x];
            if (mp == NONE || (mp & FLAGBIT)) out_i[ipool_base[img] + (u32)a.rec[base + x].indices_off + rel] = p;
            if (mp == NONE) break;
            const u32 z = mp & MPAR_MASK;
            if (mp & HOLLOWBIT) out_h[hpool_base[img] + (u32)a.rec[base + z].holes_off + a.holeoff[base + x] + rel] = p;
            rel += a.childoff[base + x];
            x = z;
        }
    }
}
__global__ void k_save_off1(const vcb_cluster_rec *rec, u32 total, u32 *off1) {
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < total; s += stride) off1[s] = (u32)rec[s].indices_off;
}
__global__ void k_hpool_base(const vcb_cluster_rec *rec, const u32 *slot_base, u32 nimg, u32 *hpool_base) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        u32 acc = 0;
        for (u32 i = 0; i < nimg; ++i) {
            hpool_base[i] = acc;
            if (slot_base[i + 1] > slot_base[i]) {
                const vcb_cluster_rec *l = rec + slot_base[i + 1] - 1;
                acc += (u32)l->holes_off + l->holes_len;
            }
        }
        hpool_base[nimg] = acc;
    }
}

}  // namespace vcb
