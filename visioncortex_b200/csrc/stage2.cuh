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
    int batch_rounds;  // parallel batch rounds on (default) / off
    int rect_perim;    // 1: perimeter by scanning the bounding rect (cluster helpers); 0: by the boundary-pixel lists
    int prefetch;      // warm L2 with the image's records / adjacency / boundary lists when the fused kernel starts
    u32 boff_min;      // batch rounds are given up for the rest of an epoch when they commit fewer pops per round than this
    int replay_loads;  // batch-round replay resolves roots / colours by loads even when the state is in global memory (runs that do not
                       // share the device with other batches: measured C3 pass 2 4.31 -> 4.01 ms, C5 76.1 -> 72.9 ms; overlapped C4: slower)
    int inert;         // the stage-1 labels hold INERT pixels (sequential resolver, unclean key under KeyingAction::Discard)
    int timing;        // VCB_S2_STATS: thread 0 accumulates the cycles spent in each phase of a pop / batch round (S2ImgState.tph)
};

struct S2ImgState {
    u32 stamp, nout, maxarea, nover;
    u32 nmerge, task_c, task_quit, perim_acc;
    i32 task_rect[4];
    u32 st_rounds, st_committed, st_cands, st_singles;   // statistics of the batch rounds / single pops
    unsigned long long tph[16];                          // cycles per phase (S2Cfg.timing)
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
    u32 *tmin;                     // batch rounds: smallest batch position that modifies the slot (NONE outside a round)
    u32 *outpos;                   // first position in clusters_output, NONE if absent
    u32 *adj_off;                  // total_slots + 1
    u32 *adj;                      // adjacency entries (image-local slot ids)
    u32 *cursor;                   // total_slots (fill cursors / degree counts)
    u32 *bnd_off, *bcur;           // per stage-1 slot: its boundary pixels (a neighbour with another label, or the image edge) ...
    uint4 *bnd_nb;                 // ... each stored as the stage-1 labels of its four neighbours (NONE = outside the image)
    u32 *col;                      // average colour (ColorSum::average) of every live root, kept current by the merges
    u32 *outputs;                  // clusters_output per image at out_mul * slot_base[img]
    u32 out_mul;                   // 2 after a hierarchical stage: a cluster that is output as isolated (builder.rs:444-450) and grows
                                   // afterwards -- clusters[0] merging into it, unclean key -- is output again at its new area, so
                                   // an image can have more outputs than slots (at most one per pop: <= slots + merges <= 2 * slots)
    u64 *over;                     // total_slots: per-image overflow of the urgent list at slot_base[img]
    S2ImgState *ist;               // nimg
    u64 *keys; u32 *vals;          // epoch sort input/output
    u32 *count;                    // number of keys collected this epoch
    int idx_bits, area_bits;
    u32 *live;                     // fused kernel: per image, its live roots in index order (compacted as clusters merge)
    u32 key_cap;                   // fused kernel: capacity of the shared-memory key array
    u32 state_cap, state_mask;     // fused kernel: slots per shared-memory state array; bit 0 fw, bit 1 col, bit 2 mnext + mtail
    u32 state_live;                // ... and which of them this CTA really holds in shared memory (set inside the kernel)
};

// Where the pop keys of the current epoch come from: the globally sorted (image | area - 2^epoch | index) array of the
// multi-launch path, or the CTA's own shared-memory array of ready-made (area << idx_bits | index) keys (fused kernel).
struct S2Keys { const u64 *g; const u64 *s; u64 mask, add; };
__device__ static __forceinline__ u64 s2_key(const S2Keys &k, u32 i) { return k.s ? k.s[i] : ((k.g[i] & k.mask) + k.add); }

// phase timing (thread 0, after a barrier): the cycles since the previous tick are added to phase k
enum { TP_SELECT = 0, TP_NBR, TP_DECIDE, TP_PERIM, TP_MERGE, TP_RELABEL, TP_BF, TP_BP, TP_BC, TP_BD, TP_BE, TP_COLLECT, TP_SORT, TP_GATHER, TP_DEG, TP_BP2 };
#define S2_TICK(k) do { if (a.cfg.timing && threadIdx.x == 0) { const unsigned long long t_ = clock64(); sh->tph[k] += t_ - sh->tlast; sh->tlast = t_; } } while (0)

// Per-slot state of the image being replayed (forwarding words, colours, member links) lives in the CTA's shared memory
// when it fits (k_s2_run) and in global memory otherwise; both are reached through these accessors.  Writers and readers
// of a word are always separated by a barrier (or are the same warp), so plain volatile accesses are enough.
template <class T> __device__ static __forceinline__ T ldv(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <class T> __device__ static __forceinline__ void stv(T *p, T v) { *reinterpret_cast<volatile T *>(p) = v; }

// Stage-1 label of a pixel that belongs to no cluster's list although the reference labels it 0: a discarded key pixel of an
// image whose clusters[0] has members of its own (unclean key, sequential.cuh).  For the graph build it is outside every
// cluster (no adjacency, no boundary record of its own; as a neighbour it reads as NONE, like the image edge -- the
// reference skips label-0 neighbours, cluster.rs:150, and to_image never sets it); its final label stays 0.
constexpr u32 INERT = 0xFFFFFFFEu;
constexpr u32 FLAGBIT = 0x80000000u;
constexpr u32 HOLLOWBIT = 0x40000000u;
constexpr u32 MPAR_MASK = 0x3fffffffu;

// Every slot points DIRECTLY at its current cluster (roots point at themselves): a merge rewrites the pointers of all
// members of the absorbed cluster (the counterpart of the reference's relabel loop, builder.rs:527-529), so resolving a
// stage-1 label is one load.  Bit 31 = "the hop into the root was a hollow deepen merge" (the pixel is in root.holes).
__device__ static __forceinline__ u32 s2_find(const u32 *fw, u32 base, u32 s, u32 *flag) {
    const u32 v = ldv(&fw[base + s]);
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
        if (a.cfg.inert) {                                  // rare path (see INERT): the row-sweep kernel does the same
            if (la == INERT) continue;
            auto L = [&](u32 i) { const u32 v = a.l1[i]; return v == INERT ? NONE : v; };
            const u32 nl = x > 0 ? L(g - 1) : NONE, nr = x + 1 < d.w ? L(g + 1) : NONE;
            const u32 nu = y > 0 ? L(g - d.w) : NONE, nd = y + 1 < d.h ? L(g + d.w) : NONE;
            if (nl != la || nr != la || nu != la || nd != la) {
                if (FILL) a.bnd_nb[a.bnd_off[base + la] + atomicAdd(&a.bcur[base + la], 1u)] = make_uint4(nl, nr, nu, nd);
                else atomicAdd(&a.bcur[base + la], 1u);
            }
            if (nr != NONE && la != nr && !(y > 0 && L(g - d.w) == la && L(g + 1 - d.w) == nr)) {
                if (FILL) {
                    a.adj[a.adj_off[base + la] + atomicAdd(&a.cursor[base + la], 1u)] = nr;
                    a.adj[a.adj_off[base + nr] + atomicAdd(&a.cursor[base + nr], 1u)] = la;
                } else { atomicAdd(&a.cursor[base + la], 1u); atomicAdd(&a.cursor[base + nr], 1u); }
            }
            if (nd != NONE && la != nd && !(x > 0 && L(g - 1) == la && L(g + d.w - 1) == nd)) {
                if (FILL) {
                    a.adj[a.adj_off[base + la] + atomicAdd(&a.cursor[base + la], 1u)] = nd;
                    a.adj[a.adj_off[base + nd] + atomicAdd(&a.cursor[base + nd], 1u)] = la;
                } else { atomicAdd(&a.cursor[base + la], 1u); atomicAdd(&a.cursor[base + nd], 1u); }
            }
            continue;
        }
        {   // boundary pixel of its stage-1 slot: only these can ever be on the perimeter of a merged cluster.  What the
            // perimeter test needs of such a pixel is the labels of its four neighbours (outside the image: NONE)
            const u32 nl = x > 0 ? a.l1[g - 1] : NONE, nr = x + 1 < d.w ? a.l1[g + 1] : NONE;
            const u32 nu = y > 0 ? a.l1[g - d.w] : NONE, nd = y + 1 < d.h ? a.l1[g + d.w] : NONE;
            if (nl != la || nr != la || nu != la || nd != la) {
                if (FILL) a.bnd_nb[a.bnd_off[base + la] + atomicAdd(&a.bcur[base + la], 1u)] = make_uint4(nl, nr, nu, nd);
                else atomicAdd(&a.bcur[base + la], 1u);
            }
        }
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

// Row-sweep form of k_rag for widths that are a multiple of 4: one warp walks down a strip of 128 columns x RAG_ROWS rows
// with the rows above / at / below the current one in registers (one 128-bit label load per four pixels, left and right
// strip halos on lanes 0 / 31), so every label is loaded once instead of seven times, and a thread whose 4 x 3 window
// holds a single label -- the interior of a cluster, most pixels -- does nothing else.  Emits exactly the items of k_rag.
constexpr int RAG_ROWS = 32;
template <bool FILL> __global__ void __launch_bounds__(256) k_rag_rows(S2 a, u32 strips_x, u32 strips_y) {
    const Dims d = a.d;
    const u32 lane = threadIdx.x & 31;
    const u32 wid = blockIdx.x * 8 + (threadIdx.x >> 5);
    const u32 per_img = strips_x * strips_y;
    if (wid >= per_img * d.nimg) return;
    const u32 img = wid / per_img, rr = wid - img * per_img, sy = rr / strips_x, sx = rr - sy * strips_x;
    const u32 base = a.slot_base[img];
    const u32 *L = a.l1 + (size_t)img * d.n;
    const u32 x0 = sx * 128 + lane * 4;
    const bool act = x0 < d.w;
    const bool halo_l = lane == 0 && sx > 0, halo_r = lane == 31 && x0 + 4 < d.w;
    const u32 y0 = sy * RAG_ROWS, y1 = y0 + RAG_ROWS < d.h ? y0 + RAG_ROWS : d.h;
    const uint4 none4 = make_uint4(NONE, NONE, NONE, NONE);
    uint4 up = none4, cur = none4, dn = none4;
    u32 upl = NONE, upr = NONE, curl = NONE, curr = NONE, dnl = NONE, dnr = NONE;
    // rows outside the image read as NONE (y0 - 1 wraps to 0xffffffff when y0 == 0)
#define RAG_LOAD(y, v, hl, hr) do { const u32 y_ = (y); v = none4; hl = NONE; hr = NONE; if (y_ < d.h) { const u32 *row_ = L + (size_t)y_ * d.w; \
        if (act) v = __ldg(reinterpret_cast<const uint4 *>(row_ + x0)); if (halo_l) hl = __ldg(row_ + x0 - 1); if (halo_r) hr = __ldg(row_ + x0 + 4); } } while (0)
    RAG_LOAD(y0 - 1, up, upl, upr);
    (void)upl;
    RAG_LOAD(y0, cur, curl, curr);
    for (u32 y = y0; y < y1; ++y) {
        RAG_LOAD(y + 1, dn, dnl, dnr);
        u32 cl = __shfl_up_sync(0xffffffffu, cur.w, 1), cr = __shfl_down_sync(0xffffffffu, cur.x, 1);
        u32 ur = __shfl_down_sync(0xffffffffu, up.x, 1), dl = __shfl_up_sync(0xffffffffu, dn.w, 1);
        if (lane == 0) { cl = curl; dl = dnl; }
        if (lane == 31) { cr = curr; ur = upr; }
        const u32 la0 = cur.x;
        const bool flat = cl == la0 && cur.y == la0 && cur.z == la0 && cur.w == la0 && cr == la0 && up.x == la0 && up.y == la0 &&
                          up.z == la0 && up.w == la0 && dn.x == la0 && dn.y == la0 && dn.z == la0 && dn.w == la0;
        if (act && !flat) {
            const u32 c[6] = {cl, cur.x, cur.y, cur.z, cur.w, cr};
            const u32 u[5] = {up.x, up.y, up.z, up.w, ur};
            const u32 dd[5] = {dl, dn.x, dn.y, dn.z, dn.w};
            // all positions are claimed first (independent atomics in flight together), the records are written afterwards
            bool fb[4], fh[4], fv[4];
            u32 pb[4], ph0[4], ph1[4], pv0[4], pv1[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u32 la = c[i + 1], nl = c[i], nr = c[i + 2], nu = u[i], nd = dd[i + 1];
                fb[i] = nl != la || nr != la || nu != la || nd != la;
                fh[i] = nr != NONE && la != nr && !(u[i] == la && u[i + 1] == nr);
                fv[i] = nd != NONE && la != nd && !(c[i] == la && dd[i] == nd);
                if (fb[i]) pb[i] = atomicAdd(&a.bcur[base + la], 1u);
                if (fh[i]) { ph0[i] = atomicAdd(&a.cursor[base + la], 1u); ph1[i] = atomicAdd(&a.cursor[base + nr], 1u); }
                if (fv[i]) { pv0[i] = atomicAdd(&a.cursor[base + la], 1u); pv1[i] = atomicAdd(&a.cursor[base + nd], 1u); }
            }
            if (FILL) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const u32 la = c[i + 1], nl = c[i], nr = c[i + 2], nu = u[i], nd = dd[i + 1];
                    if (fb[i]) a.bnd_nb[a.bnd_off[base + la] + pb[i]] = make_uint4(nl, nr, nu, nd);
                    if (fh[i]) { a.adj[a.adj_off[base + la] + ph0[i]] = nr; a.adj[a.adj_off[base + nr] + ph1[i]] = la; }
                    if (fv[i]) { a.adj[a.adj_off[base + la] + pv0[i]] = nd; a.adj[a.adj_off[base + nd] + pv1[i]] = la; }
                }
            }
        }
        up = cur; upl = curl; upr = curr;
        cur = dn; curl = dnl; curr = dnr;
    }
#undef RAG_LOAD
}

struct BndScan {
    S2 a;
    __device__ u32 count() const { return a.total_slots; }
    __device__ u32 value(u32 s) const { return a.bcur[s]; }
    __device__ void emit(u32 s, u32 ex, u32 v) const {
        a.bnd_off[s] = ex;
        a.bcur[s] = 0;
        if (s == a.total_slots - 1) a.bnd_off[a.total_slots] = ex + v;
    }
};

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

__global__ void __launch_bounds__(256) k_s2_init(S2 a) {
    const u32 stride = gridDim.x * 256;
    for (u32 s = blockIdx.x * 256 + threadIdx.x; s < a.total_slots; s += stride) {
        a.fw[s] = s - a.slot_base[a.slot_img[s]]; a.mpar[s] = NONE; a.mnext[s] = NONE; a.mtail[s] = s - a.slot_base[a.slot_img[s]];
        a.mark[s] = 0; a.tmin[s] = NONE; a.outpos[s] = NONE; a.cursor[s] = 0; a.bcur[s] = 0;
        const u32 area = a.rec[s].indices_len;
        if (area) { atomicMax(&a.ist[a.slot_img[s]].maxarea, area); a.col[s] = avg_color(a.rec[s].sum); }
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

constexpr size_t S2_FUSED_SMEM_MAX = 200 * 1024;   // fused kernel: fixed part + key array must fit one SM's shared memory
constexpr int S2_MAX_THREADS = 1024;   // block size is chosen at launch: 1024 for a few large images, 128 for big batches
constexpr int S2_MEMCHUNK = 256;
constexpr int S2_URGENT = 64;
constexpr u32 S2_MAPBYTES = 28672;
constexpr int S2_NB = 32;                 // candidates per batch round (one warp each)
constexpr int S2_CAPN = 32;               // neighbour roots kept per candidate
constexpr int S2_CAPM = 32;               // member slots kept per candidate
enum : u32 { BF_SLOW = 1, BF_OUTONLY = 2, BF_ISOLATED = 4, BF_DEEPEN = 8, BF_HOLLOW = 16, BF_CONFLICT = 32, BF_MERGE = 64, BF_NEEDP = 128 };

struct S2Shared {
    unsigned long long best;
    unsigned long long urgent[S2_URGENT];
    u32 mem[S2_MEMCHUNK];
    u32 eoff[S2_MEMCHUNK];
    u32 deg[S2_MEMCHUNK + 1];
    u32 cur, area, ndist, perim, nmem, more, mycolor, nurg, lo, hi, stamp, nchunks, newfw, target, want_perim, deepen, single;
    u32 wpos[32], full, pending, nslots;   // member scan (s2_scan_members): per-warp cursors over the image's slots
    i32 rect[4];
    // batch rounds
    unsigned long long bbest[S2_NB], bkey[S2_NB];
    u32 bc[S2_NB], barea[S2_NB], btarget[S2_NB], bndist[S2_NB], bflags[S2_NB], bnewarea[S2_NB], bpos[S2_NB], bseg[S2_NB], bstamp[S2_NB], bnmem[S2_NB];
    u32 bnbr[S2_NB][S2_CAPN];
    u32 bmem[S2_NB][S2_CAPM];
    u32 bb0[S2_NB][S2_CAPM], bpre[S2_NB][S2_CAPM + 1];   // perimeter work of a round: first boundary record / running count per member
    u32 bperim[S2_NB], bpoff[S2_NB + 1];
    u32 bcolor[S2_NB];                 // colour of the candidate when the round started
    u32 frec[S2_NB][24];               // the candidate's own record (vcb_cluster_rec as 24 words) when the round started
    u32 prec[S2_NB][24];               // the record of the target chosen on the state the round started from
    u32 trec[S2_NB][24], tslot[S2_NB]; // records of the targets modified in this round (written back when the round ends)
    u32 bncol[S2_NB][S2_CAPN];         // colours of the candidate's neighbour roots when the round started
    u32 mcs[S2_NB], mts[S2_NB];        // merges committed in this round, in order: absorbed root -> target (global-state mode)
    u32 tcol[S2_NB], tmtail[S2_NB];    // per cached target: current colour, current tail of its member list
    u32 rtmp[32];
    u32 fmtail[S2_NB], pmtail[S2_NB];  // member-list tails of the candidate / of its S0 target when the round started
    u32 nb, ncommit, mode, force_single, brounds, bdone, boff;
    u32 nlive, wsum[33];
    unsigned long long tph[16], tlast;
    S2ImgState ist_l;                  // the image's counters while the fused kernel runs (written back at the end)
    alignas(16) u8 map[S2_MAPBYTES];   // last member: allocated only in the rect-scan perimeter mode
};

// ---------------------------------------------------------------------------------------------------------------
// Batch round: the next candidates of the sorted stream are evaluated IN PARALLEL (one warp each) and the longest prefix
// that is provably independent is committed in parallel.  Candidate k depends on an earlier candidate i of the batch
// iff i modifies a slot that k reads: i's own slot c_i (absorbed) or its target t_i (sums / area / neighbourhood change),
// where k reads c_k and its neighbour roots N(c_k) -- the distance-2 rule of SURVEY Appendix D.  tmin[slot] holds the
// smallest batch position modifying the slot, so the test is `min over {c_k} U N(c_k) of tmin < k`.  A merge also creates
// a new key (new area, target); the prefix additionally stops before the first candidate whose key exceeds a new key made
// earlier in the batch, so the global (area, index) pop order is never violated.  Everything else of a pop (outputs in
// candidate order, the running maximum area of builder.rs:446, list offsets) is a prefix scan over the batch.
__device__ static __forceinline__ bool s2_in_cluster(const S2 &a, const u32 *l1, u32 base, u32 q, u32 c) {
    return ldv(&a.fw[base + __ldg(&l1[q])]) == c;
}
__device__ static __forceinline__ unsigned long long warp_min64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}

__device__ static __forceinline__ bool s2_batch_round(const S2 &a, S2Shared *sh, S2ImgState *ist, const S2Keys &keys, int epoch,
                                                      u32 img, u32 base, u64 *over) {
    const Dims d = a.d;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 nwarps = blockDim.x >> 5;
    // a warp evaluates the candidates q = warp, warp + nwarps, ... one after the other: a round always takes up to S2_NB
    // candidates, whatever the block size
    const u32 W = (u32)S2_NB;
    const u64 idx_mask = (1ull << a.idx_bits) - 1ull;
    // ---- F: warp 0 takes the next entries of the sorted segment that precede every pending urgent key
    if (warp == 0) {
        unsigned long long umin = ~0ull;
        for (u32 i = lane; i < sh->nurg; i += 32) umin = sh->urgent[i] < umin ? sh->urgent[i] : umin;
        const u32 nover = ist->nover;
        for (u32 i = lane; i < nover; i += 32) umin = over[i] < umin ? over[i] : umin;
        umin = warp_min64(umin);
        const bool use = sh->force_single == 0 && sh->boff == 0;
        const u32 lo = sh->lo, hi = sh->hi;
        unsigned long long kb = ~0ull;
        bool inrange = false, valid = false;
        u32 idx = 0, ar = 0;
        if (use && lane < W && lo + lane < hi) {
            kb = s2_key(keys, lo + lane);
            inrange = kb < umin;
            if (inrange) {
                idx = (u32)(kb & idx_mask); ar = (u32)(kb >> a.idx_bits);
                valid = a.fw[base + idx] == idx && a.rec[base + idx].indices_len == ar;
            }
        }
        const u32 m_in = __ballot_sync(0xffffffffu, inrange);
        const u32 npre = (~m_in) ? (u32)(__ffs((int)~m_in) - 1) : 32u;       // consecutive entries from the head
        const u32 vmask = __ballot_sync(0xffffffffu, valid && lane < npre);
        const u32 nb = (u32)__popc(vmask);
        if (use && npre > 0) {
            if (valid && lane < npre) {
                const u32 slot = (u32)__popc(vmask & ((1u << lane) - 1u));
                sh->bc[slot] = idx; sh->barea[slot] = ar; sh->bseg[slot] = lo + lane; sh->bkey[slot] = kb;
                sh->bstamp[slot] = ist->stamp + 1 + slot; sh->bflags[slot] = 0; sh->bbest[slot] = ~0ull; sh->bndist[slot] = 0;
                sh->btarget[slot] = NONE; sh->bnewarea[slot] = 0; sh->bnmem[slot] = 0; sh->bperim[slot] = 0;
            }
            __syncwarp();
            if (lane == 0) { ist->stamp += nb; sh->lo = lo + npre; sh->nb = nb; sh->mode = nb ? 1u : 2u; }
        } else if (lane == 0) { sh->mode = 0; sh->force_single = 0; }
    }
    __syncthreads();
    S2_TICK(TP_BF);
    if (sh->mode == 0) return false;
    if (sh->mode == 2) return true;
    const u32 nb = sh->nb;
    // ---- P: one warp per candidate: neighbours, arg-min, deepen / hollow, new area
    for (u32 q = warp; q < nb; q += nwarps) {
        const u32 c = sh->bc[q], area = sh->barea[q];
        u32 flags = 0;
        if (area > a.cfg.hierarchical) flags = BF_OUTONLY;
        else if (c == 0) flags = BF_SLOW;                            // clusters[0]: its merge turns label-0 pixels into neighbours
        else {
            bool slow = false;
            // member slots of c (most candidates never absorbed anything: one load)
            if (lane == 0) {
                u32 n = 0, m = c;
                while (m != NONE && n < (u32)S2_CAPM) { sh->bmem[q][n++] = m; m = ldv(&a.mnext[base + m]); }
                sh->bnmem[q] = n;
                if (m != NONE) sh->bnmem[q] = 0xffffffffu;
            }
            __syncwarp();
            const u32 nm = sh->bnmem[q];
            if (nm == 0xffffffffu) slow = true;
            else {
                const u32 mycolor = ldv(&a.col[base + c]);
                if (lane < 24) sh->frec[q][lane] = ldv(reinterpret_cast<const u32 *>(a.rec + base + c) + lane);
                if (lane == 0) { sh->bcolor[q] = mycolor; sh->fmtail[q] = ldv(&a.mtail[base + c]); }
                // adjacency ranges and boundary ranges of all members at once (lane i = member i): one round trip
                u32 e0 = 0, e1 = 0, b0 = 0, b1 = 0;
                if (lane < nm) {
                    const u32 sl = sh->bmem[q][lane];
                    e0 = a.adj_off[base + sl]; e1 = a.adj_off[base + sl + 1];
                    if (ldv(&a.fw[base + sl]) == c) { b0 = a.bnd_off[base + sl]; b1 = a.bnd_off[base + sl + 1]; }   // hole members: no pixels of c
                }
                // distinct neighbour roots are collected in the candidate's own list (the global mark[] stamps cannot be
                // shared by candidates that are scanned concurrently); no load depends on another inside this loop nest
                // except adjacency entry -> forwarding word
                u32 n = 0;                                           // warp-uniform
                for (u32 i = 0; i < nm; ++i) {
                    const u32 me0 = __shfl_sync(0xffffffffu, e0, (int)i), me1 = __shfl_sync(0xffffffffu, e1, (int)i);
                    for (u32 eb = me0; eb < me1; eb += 32) {
                        const u32 e = eb + lane;
                        u32 o = NONE;
                        if (e < me1) {
                            u32 fl;
                            o = s2_find(a.fw, base, a.adj[e], &fl);
                            if (o == c || o == 0) o = NONE;
                        }
                        u32 pending = __ballot_sync(0xffffffffu, o != NONE);
                        while (pending) {
                            const int src = __ffs((int)pending) - 1;
                            const u32 v = __shfl_sync(0xffffffffu, o, src);
                            const bool found = __any_sync(0xffffffffu, lane < n && lane < (u32)S2_CAPN && sh->bnbr[q][lane] == v);
                            if (!found) {
                                if ((int)lane == src && n < (u32)S2_CAPN) sh->bnbr[q][n] = v;
                                ++n;
                                __syncwarp();
                            }
                            pending &= ~__ballot_sync(0xffffffffu, o == v);
                        }
                    }
                }
                const u32 nd = n;
                if (nd > (u32)S2_CAPN) slow = true;
                else if (nd == 0) flags = BF_ISOLATED;
                else {
                    // arg-min of diff * 65535 + index (builder.rs:452-454) over the distinct neighbours: lane i = neighbour i
                    unsigned long long kbest = ~0ull;
                    if (lane < nd) {
                        const u32 v = sh->bnbr[q][lane];
                        const u32 vc = ldv(&a.col[base + v]);
                        sh->bncol[q][lane] = vc;
                        const i32 df = color_diff(mycolor, vc);
                        kbest = ((unsigned long long)((u32)df * 65535u + v) << 32) | v;
                    }
                    const unsigned long long best = warp_min64(kbest);
                    const u32 tg = (u32)best;
                    const i32 md = (i32)(((u32)(best >> 32) - tg) / 65535u);
                    bool deepen = false, needp = false;
                    if (a.cfg.hier_max && a.cfg.good_min < (u64)area && (u64)area < a.cfg.good_max && md > a.cfg.deepen_diff) {
                        if (a.cfg.good_min == 0) deepen = true;
                        else {
                            // the perimeter (patch_good, runner.rs:131-146) is counted from the members' boundary-pixel records
                            // by the whole block after this phase: one flat work list over all candidates of the round, so that a
                            // long boundary does not hold up the round on a single warp
                            const u32 len = b1 - b0;
                            u32 inc = len;
#pragma unroll
                            for (int o2 = 1; o2 < 32; o2 <<= 1) { const u32 t2 = __shfl_up_sync(0xffffffffu, inc, o2); if ((int)lane >= o2) inc += t2; }
                            sh->bb0[q][lane] = b0;
                            sh->bpre[q][lane] = inc - len;
                            if (lane == 31) sh->bpre[q][32] = inc;
                            needp = true;
                        }
                    }
                    if (!slow) {
                        flags = BF_MERGE | (deepen ? BF_DEEPEN : 0u) | (needp ? BF_NEEDP : 0u) | (((u64)nd <= a.cfg.hollow_neighbours) ? BF_HOLLOW : 0u);
                        if (lane == 0) { sh->bndist[q] = nd; sh->btarget[q] = tg; sh->pmtail[q] = ldv(&a.mtail[base + tg]); }
                        if (lane < 24) sh->prec[q][lane] = ldv(reinterpret_cast<const u32 *>(a.rec + base + tg) + lane);
                    }
                }
            }
            if (slow) flags = BF_SLOW;
        }
        if (lane == 0) sh->bflags[q] = flags;
    }
    __threadfence_block();
    __syncthreads();
    S2_TICK(TP_BP);
    // ---- P2: perimeters of the candidates that need one, block-wide over the flat list of their boundary records
    if (warp == 0) {
        const u32 v = (lane < nb && (sh->bflags[lane] & BF_NEEDP)) ? sh->bpre[lane][32] : 0u;
        u32 inc = v;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) { const u32 t2 = __shfl_up_sync(0xffffffffu, inc, o2); if ((int)lane >= o2) inc += t2; }
        sh->bpoff[lane] = inc - v;
        if (lane == 31) sh->bpoff[32] = inc;
    }
    __syncthreads();
    {
        const u32 total = sh->bpoff[32];
        if (total) {
            // (candidate q, member lo) of item `it`: the last q with bpoff[q] <= it, the last lo with bpre[q][lo] <= it - bpoff[q].
            // A thread's items only move forward, so both are advanced from where the previous item left them instead of two
            // binary searches (ten dependent shared-memory loads) per item.
            u32 q = 0, lo = 0;
            for (u32 it = tid; it < total; it += blockDim.x) {
                while (q < 31u && sh->bpoff[q + 1] <= it) { ++q; lo = 0; }
                const u32 loc = it - sh->bpoff[q];
                while (lo < 31u && sh->bpre[q][lo + 1] <= loc) ++lo;
                const uint4 nbq = a.bnd_nb[sh->bb0[q][lo] + (loc - sh->bpre[q][lo])];
                const u32 c = sh->bc[q];
                const u32 f0 = nbq.x == NONE ? NONE : ldv(&a.fw[base + nbq.x]), f1 = nbq.y == NONE ? NONE : ldv(&a.fw[base + nbq.y]);
                const u32 f2 = nbq.z == NONE ? NONE : ldv(&a.fw[base + nbq.z]), f3 = nbq.w == NONE ? NONE : ldv(&a.fw[base + nbq.w]);
                if (f0 != c || f1 != c || f2 != c || f3 != c) atomicAdd(&sh->bperim[q], 1u);
            }
            __syncthreads();
            if (tid < nb && (sh->bflags[tid] & BF_NEEDP) && (u64)sh->bperim[tid] < (u64)sh->barea[tid]) sh->bflags[tid] |= BF_DEEPEN;
            __syncthreads();
        }
    }
    S2_TICK(TP_BP2);
    // ---- R: warp 0 replays the candidates in pop order.  Each was evaluated on the state the round started from (S0); what
    // an earlier candidate of the round changes is patched in before the next one decides, all of it in shared memory:
    //   * a candidate that has meanwhile been the TARGET of a merge is no longer a pop (its area grew: the key is stale);
    //   * its neighbour roots are re-resolved through the forwarding words (a root absorbed by an earlier merge of the round
    //     carries all its slots along, so resolving the S0 roots again is exact), deduplicated, and the arg-min of
    //     diff * 65535 + index (builder.rs:452-454) is taken over the colours as they are now;
    //   * the perimeter of patch_good depends only on the candidate's own pixels, which no other pop touches;
    //   * the records of the targets live in a small cache (trec) for the round, so a chain of merges into one big
    //     neighbour costs shared-memory updates instead of dependent global round trips.
    // The replay stops in front of a candidate that needs the single-pop path, whose key exceeds a key created earlier in the
    // round (the new cluster pops first), or whose deepen test now needs a perimeter that was not counted.
    if (warp == 0) {
        const u32 FULLM = 0xffffffffu;
        u32 J = nb, ntc = 0, nmerge = 0, nmc = 0;
        u32 nout = ist->nout, maxarea = ist->maxarea;
        unsigned long long nkmin = ~0ull;
        for (u32 q = 0; q < nb; ++q) {
            __syncwarp();
            const u32 c = sh->bc[q], f = sh->bflags[q], area = sh->barea[q];
            if (__any_sync(FULLM, lane < ntc && sh->tslot[lane] == c)) continue;      // absorbed something in this round: stale key
            if (f & BF_SLOW) { J = q; break; }
            if (sh->bkey[q] > nkmin) { J = q; break; }
            if (f & BF_OUTONLY) {                                                     // builder.rs:429-432
                if (lane == 0) { a.outputs[a.out_mul * base + nout] = c; atomicMin(&a.outpos[base + c], nout); }
                ++nout;
                continue;
            }
            if (f & BF_ISOLATED) {                                                    // builder.rs:444-450
                if (area == maxarea || a.cfg.can_discard) {
                    if (lane == 0) { a.outputs[a.out_mul * base + nout] = c; atomicMin(&a.outpos[base + c], nout); }
                    ++nout;
                }
                continue;
            }
            const u32 nd0 = sh->bndist[q];
            u32 v = NONE;
            if (lane < nd0) v = sh->bnbr[q][lane];
            // A root absorbed earlier in this round: every commit rewrites the forwarding words of the absorbed members
            // (the S0 root itself among them) before the next candidate is looked at, and a root that absorbed something
            // is never absorbed in the same round (stale key, above), so one look at the forwarding word resolves it.
            // With the per-slot state in shared memory that look-up (and the colour below) is a few cycles; with the state
            // in global memory (overlapped device batches) the round's own small tables are cheaper than two L2 round trips.
            const bool local = (a.state_live & 3u) == 3u || a.cfg.replay_loads;
            if (local) { if (nmc && v != NONE) v = ldv(&a.fw[base + v]) & ~FLAGBIT; }
            else for (u32 i = 0; i < nmc; ++i) if (v == sh->mcs[i]) v = sh->mts[i];
            const u32 grp = __match_any_sync(FULLM, v);                               // lanes holding the same root
            const bool lead = v != NONE && (u32)(__ffs((int)grp) - 1) == lane;
            const u32 nd = (u32)__popc(__ballot_sync(FULLM, lead));
            unsigned long long kbest = ~0ull;
            if (lead) {
                // colour now: col[] is kept current by every commit (below); a root no merge touched is as it was at S0
                u32 vc = sh->bncol[q][lane];
                if (local) { if (nmc) vc = ldv(&a.col[base + v]); }
                else for (u32 i = 0; i < ntc; ++i) if (sh->tslot[i] == v) vc = sh->tcol[i];
                const i32 df = color_diff(sh->bcolor[q], vc);
                kbest = ((unsigned long long)((u32)df * 65535u + v) << 32) | v;
            }
            const unsigned long long best = warp_min64(kbest);
            const u32 tg = (u32)best;
            const i32 md = (i32)(((u32)(best >> 32) - tg) / 65535u);
            bool deepen = false;
            if (a.cfg.hier_max && a.cfg.good_min < (u64)area && (u64)area < a.cfg.good_max && md > a.cfg.deepen_diff) {
                if (a.cfg.good_min == 0) deepen = true;
                else if (f & BF_NEEDP) deepen = (u64)sh->bperim[q] < (u64)area;
                else { J = q; break; }
            }
            const bool hollow = (u64)nd <= a.cfg.hollow_neighbours;
            if (deepen) {                                                             // builder.rs:463-465
                if (lane == 0) { a.outputs[a.out_mul * base + nout] = c; atomicMin(&a.outpos[base + c], nout); }
                ++nout;
            }
            // the target's record: already cached (modified earlier in the round), the S0 copy read by the candidate's
            // warp (the usual case: the choice did not change), or a fresh load
            const u32 hitm = __ballot_sync(FULLM, lane < ntc && sh->tslot[lane] == tg);
            u32 ti;
            if (hitm) ti = (u32)__ffs((int)hitm) - 1u;
            else {
                ti = ntc++;
                const bool predicted = sh->btarget[q] == tg;
                if (lane == 0) { sh->tslot[ti] = tg; sh->tmtail[ti] = predicted ? sh->pmtail[q] : ldv(&a.mtail[base + tg]); }
                if (lane < 24) sh->trec[ti][lane] = predicted ? sh->prec[q][lane] : ldv(reinterpret_cast<const u32 *>(a.rec + base + tg) + lane);
                __syncwarp();
            }
            u32 *tw = sh->trec[ti];
            const u32 *fr = sh->frec[q];
            // merge_cluster_into (builder.rs:495-539) on the cached words: 0-4 sum, 5-9 residue_sum, 10-13 rect, 14 num_holes,
            // 15 depth, 16 merged_into, 17 indices_len, 22 holes_len
            if (lane < 5) {
                if (!deepen) tw[5 + lane] += fr[5 + lane];
                tw[lane] += fr[lane];
            } else if (lane == 5) {
                const i32 f0 = (i32)fr[10], f1 = (i32)fr[11], f2 = (i32)fr[12], f3 = (i32)fr[13];
                if (!(f2 - f0 == 0 && f3 - f1 == 0)) {                                // bound.rs:204-219
                    const i32 t0 = (i32)tw[10], t1 = (i32)tw[11], t2 = (i32)tw[12], t3 = (i32)tw[13];
                    if (t2 - t0 == 0 && t3 - t1 == 0) { tw[10] = (u32)f0; tw[11] = (u32)f1; tw[12] = (u32)f2; tw[13] = (u32)f3; }
                    else {
                        tw[10] = (u32)(t0 < f0 ? t0 : f0); tw[11] = (u32)(t1 < f1 ? t1 : f1);
                        tw[12] = (u32)(t2 > f2 ? t2 : f2); tw[13] = (u32)(t3 > f3 ? t3 : f3);
                    }
                }
            } else if (lane == 6) {
                a.childoff[base + c] = tw[17];                                        // builder.rs:529: to.indices.append(from.indices)
                a.holeoff[base + c] = tw[22];
                tw[17] += area;
                if (deepen) {
                    if (hollow) { tw[22] += area; tw[14] += 1; }
                    tw[15] += 1;
                }
                a.mpar[base + c] = tg | (deepen ? FLAGBIT : 0u) | ((deepen && hollow) ? HOLLOWBIT : 0u);
            } else if (lane == 7) {
                stv(&a.mnext[base + sh->tmtail[ti]], c);                              // the member lists are concatenated
                sh->tmtail[ti] = sh->fmtail[q];
                sh->mcs[nmc] = c; sh->mts[nmc] = tg;
            } else if (lane >= 8 && lane < 18) {
                // the absorbed cluster's own record (global): cleared unless it keeps its pixels as a deepened layer
                u32 *rw = reinterpret_cast<u32 *>(a.rec + base + c);
                const u32 k = lane - 8;                                               // 0-4 sum, 5-8 rect, 9 indices_len
                if (!deepen) rw[k < 5 ? k : (k < 9 ? k + 5 : 17)] = 0u;
                else if (k == 0) rw[16] = tg;
            }
            __syncwarp();
            {   // ColorSum::average of the target's new sum
                u32 ch = 0;
                if (lane < 4) ch = ((tw[lane] / tw[4]) & 255u) << (8 * lane);
                ch |= __shfl_xor_sync(FULLM, ch, 1);
                ch |= __shfl_xor_sync(FULLM, ch, 2);
                if (lane == 0) { stv(&a.col[base + tg], ch); sh->tcol[ti] = ch; }
            }
            ++nmc;
            const u32 newfw = tg | ((deepen && hollow) ? FLAGBIT : 0u);
            const u32 nm = sh->bnmem[q];
            if (lane < nm) stv(&a.fw[base + sh->bmem[q][lane]], newfw);
            const u32 na = tw[17];
            if (na > maxarea) maxarea = na;
            ++nmerge;
            const unsigned long long nk = ((unsigned long long)na << a.idx_bits) | tg;
            if (nk < nkmin) nkmin = nk;
            if (lane == 0 && (31 - __clz((int)na)) == epoch) {                        // same epoch: must still be popped in order
                if (sh->nurg < (u32)S2_URGENT) sh->urgent[sh->nurg++] = nk; else over[ist->nover++] = nk;
            }
        }
        __syncwarp();
        for (u32 e = 0; e < ntc; ++e) {
            if (lane < 24) reinterpret_cast<u32 *>(a.rec + base + sh->tslot[e])[lane] = sh->trec[e][lane];
            if (lane == 24) stv(&a.mtail[base + sh->tslot[e]], sh->tmtail[e]);
        }
        if (lane == 0) {
            ist->nout = nout; ist->maxarea = maxarea; ist->nmerge += nmerge;
            sh->ncommit = J;
            if (J < nb) sh->lo = sh->bseg[J];                             // the rest is read again next round
            if (J == 0) sh->force_single = 1;                             // head candidate needs the single-pop path
            sh->brounds += 1; sh->bdone += J;
            ist->st_rounds += 1; ist->st_committed += J; ist->st_cands += nb;
        }
    }
    __threadfence_block();
    __syncthreads();
    S2_TICK(TP_BE);
    return true;
}

// Members of cluster c WITHOUT walking its linked list: every slot points directly at its current cluster, so the members
// are the slots whose forwarding word (hollow flag ignored) is c.  Each warp scans its own slice of the image's slots with
// four coalesced loads in flight and appends the matches to sh->mem (order is irrelevant to every user: distinct-neighbour
// sets, perimeters and relabelling); a chunk holds S2_MEMCHUNK members, the per-warp cursors carry on from there.  A list
// walk costs one dependent load per member on a single thread -- 190 k cycles per pop of a region that has absorbed a few
// hundred speckles (C5) -- the scan costs nslots / (32 * warps) steps.  Used when the image has few enough slots.
constexpr u32 S2_SCAN_PER_THREAD = 256;     // scan when nslots <= this many slots per thread of the CTA
__device__ static __forceinline__ bool s2_scan_mode(const S2Shared *sh) { return sh->nslots <= S2_SCAN_PER_THREAD * blockDim.x; }
__device__ static __forceinline__ void s2_scan_begin(S2Shared *sh) {            // call by all threads; followed by a barrier in s2_scan_members
    const u32 nw = blockDim.x >> 5;
    const u32 slice = ((sh->nslots + nw - 1) / nw + 31u) & ~31u;
    if (threadIdx.x < 32) sh->wpos[threadIdx.x] = threadIdx.x * slice;
}
// fills sh->mem / sh->nmem with the next chunk; sh->more = NONE when every slice is exhausted.  `exact`: only slots whose
// word is exactly c (no hollow flag) -- the slots that hold pixels of c.
__device__ static __forceinline__ void s2_scan_members(const S2 &a, S2Shared *sh, u32 base, u32 c, bool exact) {
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const u32 nslots = sh->nslots;
    const u32 slice = ((nslots + nw - 1) / nw + 31u) & ~31u;
    __syncthreads();
    if (tid == 0) { sh->nmem = 0; sh->full = 0; sh->pending = 0; }
    __syncthreads();
    u32 pos = sh->wpos[warp];
    const u32 end = (warp + 1) * slice < nslots ? (warp + 1) * slice : nslots;
    bool stop = false;
    // (the flag is read by one lane and broadcast: the loop condition must be uniform across the warp)
    while (pos < end && !stop && !__shfl_sync(0xffffffffu, ldv(&sh->full), 0)) {
        u32 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const u32 sl = pos + u * 32 + lane; v[u] = sl < end ? ldv(&a.fw[base + sl]) : NONE; }
        u32 done = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (stop) break;
            const u32 sl = pos + u * 32 + lane;
            const bool match = sl < end && (exact ? v[u] == c : (v[u] & ~FLAGBIT) == c);
            const u32 m = __ballot_sync(0xffffffffu, match);
            if (m) {
                const u32 cnt = (u32)__popc(m);
                u32 k = 0;
                if (lane == 0) {
                    // reserve [k, k + cnt) only if it fits (compare-and-swap: an add that is taken back later would let
                    // another warp's reservation land behind the hole)
                    k = ldv(&sh->nmem);
                    while (true) {
                        if (k + cnt > (u32)S2_MEMCHUNK) { stv(&sh->full, 1u); k = NONE; break; }
                        const u32 seen = atomicCAS(&sh->nmem, k, k + cnt);
                        if (seen == k) break;
                        k = seen;
                    }
                }
                k = __shfl_sync(0xffffffffu, k, 0);
                if (k == NONE) { stop = true; break; }
                if (match) sh->mem[k + (u32)__popc(m & ((1u << lane) - 1u))] = sl;
            }
            done = (u32)u + 1;
        }
        pos += done * 32;
    }
    if (lane == 0) { sh->wpos[warp] = pos; if (pos < end) stv(&sh->pending, 1u); }
    __syncthreads();
    if (tid == 0) sh->more = sh->pending ? c : NONE;
    __syncthreads();
}

// Perimeter of cluster c from the boundary-pixel lists of its member slots (block-parallel).  A pixel whose four
// neighbours carry its own stage-1 label stays interior whatever merges happen, so only the stage-1 boundary pixels of the
// members need to be looked at: the cost is the boundary length of the members, not the area of the bounding rect.
// Members that are holes of c (forwarding word != c) are skipped whole.  Adds the block's count to sh->perim.
__device__ static __forceinline__ void s2_perimeter_bnd(const S2 &a, S2Shared *sh, u32 img, u32 base, u32 c) {
    const Dims d = a.d;
    const u32 tid = threadIdx.x, T = blockDim.x;
    const u32 *l1 = a.l1 + (size_t)img * d.n;
    u32 cntp = 0;
    const bool scan = s2_scan_mode(sh) && sh->single == 0;     // a cluster that never absorbed anything is its own only member
    if (tid == 0) sh->more = c;
    if (scan) s2_scan_begin(sh);
    __syncthreads();
    while (true) {
        if (scan) s2_scan_members(a, sh, base, c, true);
        else if (tid == 0) {
            u32 n = 0, m = sh->more;
            while (m != NONE && n < S2_MEMCHUNK) { sh->mem[n++] = m; m = a.mnext[base + m]; }
            sh->nmem = n; sh->more = m;
        }
        __syncthreads();
        for (u32 i = tid; i < sh->nmem; i += T) {
            const u32 m = sh->mem[i];
            const bool in = ldv(&a.fw[base + m]) == c;              // root is c and the slot is not one of its holes
            sh->eoff[i] = a.bnd_off[base + m];
            sh->deg[i] = in ? a.bnd_off[base + m + 1] - a.bnd_off[base + m] : 0u;
        }
        __syncthreads();
        if (tid == 0) { const u32 nm0 = sh->nmem; u32 acc = 0; for (u32 i = 0; i < nm0; ++i) { const u32 dg = sh->deg[i]; sh->deg[i] = acc; acc += dg; } sh->deg[nm0] = acc; }
        __syncthreads();
        const u32 nm = sh->nmem, nitems = sh->deg[nm];
        for (u32 it = tid; it < nitems; it += 2 * T) {
            u32 kk[2]; bool ok[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const u32 t2 = it + u * T;
                ok[u] = t2 < nitems;
                u32 lo = 0, hi = nm;
                if (ok[u]) while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (sh->deg[mid] <= t2) lo = mid; else hi = mid; }
                kk[u] = ok[u] ? sh->eoff[lo] + (t2 - sh->deg[lo]) : 0u;
            }
            const uint4 q0 = ok[0] ? a.bnd_nb[kk[0]] : make_uint4(NONE, NONE, NONE, NONE);
            const uint4 q1 = ok[1] ? a.bnd_nb[kk[1]] : q0;
            const u32 nb[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
            u32 f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = nb[j] == NONE ? NONE : ldv(&a.fw[base + nb[j]]);
            if (ok[0] && (f[0] != c || f[1] != c || f[2] != c || f[3] != c)) cntp++;
            if (ok[1] && (f[4] != c || f[5] != c || f[6] != c || f[7] != c)) cntp++;
        }
        __syncthreads();
        const bool last = sh->more == NONE;
        __syncthreads();
        if (last) break;
    }
    if (cntp) atomicAdd(&sh->perim, cntp);
}

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
            for (int u = 0; u < 4; ++u) fwv[u] = ok[u] ? ldv(&a.fw[base + lab[u]]) : 0xffffffffu;
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

// helper CTAs of an image's thread-block cluster (rect-scan perimeter mode): they sleep in the cluster barrier and are
// woken for a perimeter task or to quit
__device__ static __forceinline__ void s2_helper_loop(const S2 &a, S2Shared *sh, S2ImgState *ist, u32 img, u32 base, u32 crank, u32 G) {
    const u32 tid = threadIdx.x;
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
}
__device__ static __forceinline__ void s2_release_helpers(S2ImgState *ist, u32 G) {
    if (G > 1) {
        if (threadIdx.x == 0) { stcg(&ist->task_quit, 1u); __threadfence(); }
        __syncthreads();
        VCB_CLUSTER_SYNC();
    }
}

// All pops of one epoch of one image, in the reference's exact order.  sh->lo / sh->hi delimit the image's sorted keys.
__device__ static __forceinline__ void s2_epoch_leader(const S2 &a, S2Shared *sh, S2ImgState *ist, const S2Keys &keys, int epoch,
                                                       u32 img, u32 base, u64 *over, u32 G) {
    const u32 S2_THREADS = blockDim.x;
    const u32 tid = threadIdx.x;
    const u64 idx_mask = (1ull << a.idx_bits) - 1ull;
    while (true) {
        if (a.cfg.batch_rounds && sh->boff == 0 && s2_batch_round(a, sh, ist, keys, epoch, img, base, over)) continue;
        if (tid == 0) {
            // next pop = smallest valid key among: sorted segment head, urgent list, overflow list
            u32 c = NONE, carea = 0;
            while (true) {
                u64 kb = ~0ull; int src = -1; u32 which = 0;
                if (sh->lo < sh->hi) { kb = s2_key(keys, sh->lo); src = 0; }
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
                ist->st_singles += 1;
                sh->best = ~0ull; sh->ndist = 0; sh->perim = 0;
                sh->stamp = ++ist->stamp;
                sh->mycolor = a.col[base + c];
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
        S2_TICK(TP_SELECT);
        const u32 c = sh->cur;
        if (c == NONE) break;
        const u32 area = sh->area;
        if (area > a.cfg.hierarchical) {                              // builder.rs:429-432
            if (tid == 0) {
                const u32 pos = ist->nout++;
                a.outputs[a.out_mul * base + pos] = c;
                if (a.outpos[base + c] == NONE) a.outpos[base + c] = pos;
            }
            __syncthreads();
            continue;
        }
        // ---- neighbours of c (cluster.rs:132-171) over the member lists and the static adjacency
        const u32 stamp = sh->stamp, mycolor = sh->mycolor;
        const bool single = sh->single != 0;                          // uniform (written before the barrier above)
        const bool scan = !single && s2_scan_mode(sh);
        if (scan) s2_scan_begin(sh);
        while (true) {
            if (!single) {
                if (scan) { s2_scan_members(a, sh, base, c, false); if (tid == 0) sh->nchunks++; }
                else if (tid == 0) {
                    u32 n = 0, m = sh->more;
                    while (m != NONE && n < S2_MEMCHUNK) { sh->mem[n++] = m; m = a.mnext[base + m]; }
                    sh->nmem = n; sh->more = m; sh->nchunks++;
                }
                __syncthreads();
                S2_TICK(TP_GATHER);
                // edge-parallel: prefix of the member degrees in shared memory, then one (member, edge) item per thread step
                for (u32 i = tid; i < sh->nmem; i += S2_THREADS) {
                    const u32 s = sh->mem[i];
                    sh->eoff[i] = a.adj_off[base + s];
                    sh->deg[i] = a.adj_off[base + s + 1] - sh->eoff[i];
                }
                __syncthreads();
                if (tid == 0) { const u32 nm0 = sh->nmem; u32 acc = 0; for (u32 i = 0; i < nm0; ++i) { const u32 dg = sh->deg[i]; sh->deg[i] = acc; acc += dg; } sh->deg[nm0] = acc; }
                __syncthreads();
                S2_TICK(TP_DEG);
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
                const i32 df = color_diff(mycolor, ldv(&a.col[base + o]));
                const unsigned long long k = ((unsigned long long)((u32)df * 65535u + o) << 32) | o;
                atomicMin(&sh->best, k);
            }
            __syncthreads();
            const bool last = sh->more == NONE;
            __syncthreads();                                          // thread 0 rewrites sh->more in the next round
            if (last) break;
        }
        S2_TICK(TP_NBR);
        const u32 ndist = sh->ndist;
        if (ndist == 0) {                                             // builder.rs:444-450
            if (tid == 0 && (area == ist->maxarea || a.cfg.can_discard)) {
                const u32 pos = ist->nout++;
                a.outputs[a.out_mul * base + pos] = c;
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
        S2_TICK(TP_DECIDE);
        const u32 target = sh->target;
        bool deepen = sh->deepen != 0;
        if (sh->want_perim) {
            const u32 rh = (u32)(sh->rect[3] - sh->rect[1]), rwid = (u32)(sh->rect[2] - sh->rect[0]);
            if (!a.cfg.rect_perim) {
                s2_perimeter_bnd(a, sh, img, base, c);
                __syncthreads();
            } else if (G > 1 && (u64)rh * rwid >= a.cfg.help_min_px) {
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
            S2_TICK(TP_PERIM);
        }
        const bool hollow = (u64)ndist <= a.cfg.hollow_neighbours;
        if (tid == 0) {
            vcb_cluster_rec *rf = a.rec + base + c, *rt_ = a.rec + base + target;
            if (deepen) {                                             // builder.rs:463-465
                const u32 pos = ist->nout++;
                a.outputs[a.out_mul * base + pos] = c;
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
            a.col[base + target] = avg_color(rt_->sum);
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
        S2_TICK(TP_MERGE);
        // relabel: every member slot of c now resolves to the target (builder.rs:527-529)
        {
            const u32 nf = sh->newfw;
            if (scan) {
                // every slot that resolves to c (hole members included), wherever it sits in the list
                const u32 ns = sh->nslots;
                for (u32 sl = tid; sl < ns; sl += S2_THREADS) if ((ldv(&a.fw[base + sl]) & ~FLAGBIT) == c) stv(&a.fw[base + sl], nf);
            } else if (sh->nchunks <= 1) {
                const u32 nm = sh->nmem;
                for (u32 i = tid; i < nm; i += S2_THREADS) stv(&a.fw[base + sh->mem[i]], nf);
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
                    for (u32 i = tid; i < nm; i += S2_THREADS) stv(&a.fw[base + sh->mem[i]], nf);
                    __syncthreads();
                    const bool last = sh->more == NONE;
                    __syncthreads();
                    if (last) break;
                }
            }
        }
        __syncthreads();
        S2_TICK(TP_RELABEL);
    }
}

// multi-launch path: one launch per epoch over the globally sorted keys of all images (prims::radix_sort between launches)
__global__ void __launch_bounds__(S2_MAX_THREADS) k_s2_epoch(S2 a, const u64 *gkeys, int epoch, u32 G) {
    VCB_SMEM(S2Shared, sh);
    const u32 img = blockIdx.x / G, crank = blockIdx.x % G, base = a.slot_base[img];   // G CTAs (one cluster) per image
    const u32 tid = threadIdx.x;
    S2ImgState *ist = a.ist + img;
    const int kshift = a.idx_bits + epoch;                      // sorted keys: image | (area - 2^epoch) | index
    if (crank != 0) { s2_helper_loop(a, sh, ist, img, base, crank, G); return; }
    if (tid == 0) {
        const u32 cnt = *a.count;
        u32 lo = 0, hi = cnt;                       // first key of this image
        while (lo < hi) { const u32 mid = (lo + hi) / 2; if ((u32)(gkeys[mid] >> kshift) < img) lo = mid + 1; else hi = mid; }
        sh->lo = lo;
        hi = cnt;
        u32 l2 = lo;
        while (l2 < hi) { const u32 mid = (l2 + hi) / 2; if ((u32)(gkeys[mid] >> kshift) <= img) l2 = mid + 1; else hi = mid; }
        sh->hi = l2;
        sh->nurg = 0; sh->force_single = 0; sh->mode = 0; sh->brounds = 0; sh->bdone = 0; sh->boff = 0;
        sh->nslots = a.slot_base[img + 1] - base;
    }
    __syncthreads();
    const S2Keys keys{gkeys, nullptr, (1ull << kshift) - 1ull, (u64)(1u << epoch) << a.idx_bits};
    s2_epoch_leader(a, sh, ist, keys, epoch, img, base, a.over + base, G);
    s2_release_helpers(ist, G);
}

// ---------------------------------------------------------------------------------------------------------------
// Fused path: ONE launch replays every epoch.  The CTA keeps the image's live roots as a compact list (index order); at the
// start of an epoch it drops the absorbed ones, emits the pop keys of the roots whose area lies in [2^e, 2^(e+1)) into
// shared memory and sorts them there (bitonic; index order makes epoch 0 sorted as collected).  No per-epoch host work,
// no global sort: the hierarchical stage of a batch is one kernel, which lets several batches overlap on the device.
__device__ static __forceinline__ u32 s2_collect_epoch(const S2 &a, S2Shared *sh, u64 *skeys, u32 base, u32 nslots, int ep, bool initial) {
    const u32 T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (T + 31) >> 5;
    const u32 nin = initial ? nslots : sh->nlive;
    u32 kept = 0, taken = 0;                                    // uniform running totals
    for (u32 t0 = 0; t0 < nin; t0 += T) {
        const u32 i = t0 + tid;
        u32 s = NONE, area = 0;
        bool keep = false, take = false;
        if (i < nin) {
            s = initial ? i : a.live[base + i];
            if (ldv(&a.fw[base + s]) == s) {
                area = a.rec[base + s].indices_len;
                keep = area != 0;
                take = keep && (31 - __clz((int)area)) == ep;
            }
        }
        const u32 mk = __ballot_sync(0xffffffffu, keep), mt = __ballot_sync(0xffffffffu, take);
        if (lane == 0) sh->wsum[warp] = (u32)__popc(mk) | ((u32)__popc(mt) << 16);
        __syncthreads();
        u32 pk = 0, pt = 0, tk = 0, tt = 0;
        for (u32 w = 0; w < nw; ++w) {
            const u32 v = sh->wsum[w];
            if (w < warp) { pk += v & 0xffffu; pt += v >> 16; }
            tk += v & 0xffffu; tt += v >> 16;
        }
        // in place: the entries of this tile were read before the barrier, later tiles start at t0 + T >= kept + tk
        if (keep) a.live[base + kept + pk + (u32)__popc(mk & ((1u << lane) - 1u))] = s;
        if (take) skeys[taken + pt + (u32)__popc(mt & ((1u << lane) - 1u))] = ((u64)area << a.idx_bits) | s;
        kept += tk; taken += tt;
        __syncthreads();
    }
    if (tid == 0) sh->nlive = kept;
    __syncthreads();
    return taken;
}
// Bitonic sort, ascending, in the "normalised" form whose every compare-exchange moves the smaller key to the lower
// index: keys beyond n behave as +infinity without being stored, so any n works in place (shared or global memory).
__device__ static __forceinline__ void s2_sort_keys(u64 *k, u32 n) {
    const u32 T = blockDim.x, tid = threadIdx.x;
    u32 m = 1;
    while (m < n) m <<= 1;
    for (u32 size = 2; size <= m; size <<= 1) {
        const u32 half = size >> 1;
        for (u32 t = tid; t < (m >> 1); t += T) {
            const u32 blk = t / half, off = t - blk * half;
            const u32 i = blk * size + off, j = blk * size + size - 1u - off;
            if (j < n) { const u64 x = k[i], y = k[j]; if (x > y) { k[i] = y; k[j] = x; } }
        }
        __syncthreads();
        for (u32 stride = half >> 1; stride > 0; stride >>= 1) {
            for (u32 t = tid; t < (m >> 1); t += T) {
                const u32 i = ((t & ~(stride - 1u)) << 1) | (t & (stride - 1u)), j = i + stride;
                if (j < n) { const u64 x = k[i], y = k[j]; if (x > y) { k[i] = y; k[j] = x; } }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(S2_MAX_THREADS) k_s2_run(S2 a_in, int max_epoch, u32 G) {
    VCB_SMEM(S2Shared, sh);
    S2 a = a_in;
    const u32 img = blockIdx.x / G, crank = blockIdx.x % G, base = a.slot_base[img];
    const u32 nslots = a.slot_base[img + 1] - base;
    const u32 tid = threadIdx.x, T = blockDim.x;
    S2ImgState *gist = a.ist + img;
    if (crank != 0) { s2_helper_loop(a, sh, gist, img, base, crank, G); return; }
    // the key array follows the fixed part of the shared block (and the byte map of the rect-scan perimeter mode), the state
    // arrays follow the keys
    u64 *skeys = reinterpret_cast<u64 *>(reinterpret_cast<u8 *>(sh) + ((offsetof(S2Shared, map) + (a.cfg.rect_perim ? S2_MAPBYTES : 0u) + 15u) & ~(size_t)15));
    u32 *sstate = reinterpret_cast<u32 *>(skeys + a.key_cap);
    // State in shared memory: the replay is a chain of dependent lookups (adjacency entry -> forwarding word -> colour, member
    // -> next member), so the words it chases are kept next to the SM.  Helper CTAs of a cluster (rect-scan mode) read the
    // forwarding words too, so that mode keeps everything in global memory.
    const u32 smask = (G == 1 && nslots <= a.state_cap) ? a.state_mask : 0u;
    u32 *gfw = a.fw;
    a.state_live = smask;
    if (smask & 1u) { u32 *p = sstate; sstate += a.state_cap; for (u32 s = tid; s < nslots; s += T) p[s] = a.fw[base + s]; a.fw = p - base; }
    if (smask & 2u) { u32 *p = sstate; sstate += a.state_cap; for (u32 s = tid; s < nslots; s += T) p[s] = a.col[base + s]; a.col = p - base; }
    if (smask & 4u) {
        u32 *p = sstate; sstate += a.state_cap; for (u32 s = tid; s < nslots; s += T) p[s] = a.mnext[base + s]; a.mnext = p - base;
        u32 *q = sstate; sstate += a.state_cap; for (u32 s = tid; s < nslots; s += T) q[s] = a.mtail[base + s]; a.mtail = q - base;
    }
#ifndef VCB_HOSTSIM
    if (a.cfg.prefetch) {
        // one pass over everything the replay will chase: the first touch of a record is then an L2 hit, not a DRAM miss
        auto pf = [&](const void *p0, size_t bytes) {
            const char *p = reinterpret_cast<const char *>(p0);
            for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
        };
        pf(a.rec + base, (size_t)nslots * sizeof(vcb_cluster_rec));
        pf(a.adj_off + base, (size_t)(nslots + 1) * 4);
        pf(a.bnd_off + base, (size_t)(nslots + 1) * 4);
        const u32 e0 = a.adj_off[base], e1 = a.adj_off[base + nslots];
        pf(a.adj + e0, (size_t)(e1 - e0) * 4);
        if (a.cfg.prefetch > 1) { const u32 b0 = a.bnd_off[base], b1 = a.bnd_off[base + nslots]; pf(a.bnd_nb + b0, (size_t)(b1 - b0) * 16); }
        pf(a.mark + base, (size_t)nslots * 4);
        if (!(smask & 1u)) pf(a.fw + base, (size_t)nslots * 4);
        if (!(smask & 2u)) pf(a.col + base, (size_t)nslots * 4);
        if (!(smask & 4u)) { pf(a.mnext + base, (size_t)nslots * 4); pf(a.mtail + base, (size_t)nslots * 4); }
    }
#endif
    // the image's counters stay in shared memory while the kernel runs (helper CTAs exchange tasks through the global copy)
    S2ImgState *ist = gist;
    if (G == 1) { if (tid == 0) sh->ist_l = *gist; ist = &sh->ist_l; }
    if (tid == 0) { sh->nlive = nslots; sh->nslots = nslots; for (int k = 0; k < 16; ++k) sh->tph[k] = 0; sh->tlast = clock64(); }
    __syncthreads();
    for (int ep = 0; ep <= max_epoch; ++ep) {
        // the roots still alive bound the number of keys: shared memory when they fit, the image's global key range otherwise
        u64 *kbuf = (ep == 0 ? nslots : sh->nlive) <= a.key_cap ? skeys : a.keys + base;
        __syncthreads();
        const u32 n = s2_collect_epoch(a, sh, kbuf, base, nslots, ep, ep == 0);
        S2_TICK(TP_COLLECT);
        if (n == 0) continue;
        if (ep > 0 && n > 1) s2_sort_keys(kbuf, n);              // epoch 0: every area is 1, collected in index order
        S2_TICK(TP_SORT);
        if (tid == 0) { sh->lo = 0; sh->hi = n; sh->nurg = 0; sh->force_single = 0; sh->mode = 0; sh->brounds = 0; sh->bdone = 0; sh->boff = 0; }
        __syncthreads();
        const S2Keys keys{nullptr, kbuf, 0ull, 0ull};
        s2_epoch_leader(a, sh, ist, keys, ep, img, base, a.over + base, G);
        __syncthreads();
    }
    if (smask & 1u) for (u32 s = tid; s < nslots; s += T) gfw[base + s] = a.fw[base + s];   // k_s2_labels, to_color_image read it
    if (tid == 0) {
        if (a.cfg.timing) for (int k = 0; k < 16; ++k) ist->tph[k] = sh->tph[k];
        if (G == 1) *gist = sh->ist_l;
    }
    s2_release_helpers(gist, G);
}

// ------------------------------------------------------------------------------------------------- finalisation
__global__ void __launch_bounds__(256) k_s2_labels(S2 a, u32 *labels) {
    const Dims d = a.d;
    const u32 stride = gridDim.x * 256;
    for (u32 g = blockIdx.x * 256 + threadIdx.x; g < d.N; g += stride) {
        const u32 base = a.slot_base[g / d.n];
        const u32 l = a.l1[g];
        labels[g] = (a.cfg.inert && l == INERT) ? 0u : (a.fw[base + l] & ~FLAGBIT);
    }
}

__global__ void __launch_bounds__(256) k_s2_labels4(S2 a, u32 *labels) {
    const Dims d = a.d;
    const u32 stride = gridDim.x * 256, ngrp = d.N / 4, gpi = d.n / 4;      // d.n % 4 == 0: a group never straddles two images
    for (u32 q = blockIdx.x * 256 + threadIdx.x; q < ngrp; q += stride) {
        const u32 base = a.slot_base[q / gpi];
        const uint4 l = __ldg(reinterpret_cast<const uint4 *>(a.l1) + q);
        uint4 o;
        o.x = a.fw[base + l.x] & ~FLAGBIT;
        o.y = l.y == l.x ? o.x : (a.fw[base + l.y] & ~FLAGBIT);
        o.z = l.z == l.y ? o.y : (a.fw[base + l.z] & ~FLAGBIT);
        o.w = l.w == l.z ? o.z : (a.fw[base + l.w] & ~FLAGBIT);
        reinterpret_cast<uint4 *>(labels)[q] = o;
    }
}

// Compact records for the host pipeline (vcb_host_result.live_slots): the slots that are not Cluster::default(), in slot order,
// with their image-local index; coff[img] = position of the image's first live slot, coff[nimg] = total.
struct LiveScan {
    const vcb_cluster_rec *rec; u32 total; const u32 *slot_base; const u32 *slot_img; u32 nimg;
    vcb_cluster_rec *crec; u32 *cidx; u32 *coff;
    __device__ u32 count() const { return total; }
    __device__ u32 value(u32 s) const {
        const vcb_cluster_rec &r = rec[s];
        return (r.indices_len | r.holes_len | r.residue_sum[4] | r.merged_into | r.depth | r.num_holes) ? 1u : 0u;
    }
    __device__ void emit(u32 s, u32 ex, u32 v) const {
        const u32 img = slot_img[s], b0 = slot_base[img];
        if (v) { crec[ex] = rec[s]; cidx[ex] = s - b0; }
        if (s == b0) coff[img] = ex;
        if (s == total - 1) coff[nimg] = ex + v;
    }
};

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
    for (u32 p = blockIdx.x * 256 + threadIdx.x; p < a.d.n; p += stride) {
        const u32 l = a.l1[(size_t)img * a.d.n + p];
        out[p] = (a.cfg.inert && l == INERT) ? 0u : slot_color[base + l];      // a discarded key pixel is in no list: unpainted
    }
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
