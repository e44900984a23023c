// ids.cuh — stage 1 of the clustering path as data-parallel kernels (colour: color_clusters/builder.rs:273-364;
// binary: clusters.rs:270-349).  The reference scans pixels in raster order and merges clusters by *dynamic* area, so
// its cluster ids are not those of a textbook connected-component labeling.  DESIGN.md ("id model") explains the
// decomposition implemented here; in short, with time = raster index:
//   K1  k_edges_tile      static per-pixel join code J and merge flag M (pure functions of 4 pixels), tile-local
//                          resolution of the J forest in shared memory, leaf (= provisional cluster) records
//   K2  k_flatten_events  global resolution of the J forest; candidate merge events (M set, endpoints in different PCs)
//   K3  k_boruvka         which candidate events are *effective* = minimum spanning forest under weight = time
//   K4  k_chain_insert    Kruskal reconstruction tree (parent of every leaf / effective event), lock-free
//   K5  k_attribute       every pixel is attributed to the tree node alive when it joined; per-node counts,
//                          the up-side child of every event, per-leaf colour sums / rects
//   K6  k_tournament      bottom-up: areas at merge time -> winners -> carriers -> label deaths
//   K7  numbering         prefix-scan compaction of NEW events minus immediately-reused slots (builder.rs:315-317)
//   K8  k_labels          final cluster_indices
#pragma once
#include "prims.cuh"

namespace vcb {

constexpr u32 NONE = 0xFFFFFFFFu;
constexpr u32 TAG = 0x80000000u;   // node ids: leaf = TAG | pixel, event = pixel; also "dead" mark in the candidate list

enum : u32 { J_NEW = 0, J_UP = 1, J_LEFT = 2, J_UPLEFT = 3, J_NONE = 4, J_RUP = 5, J_RES = 6, J_MASK = 7,
             F_M = 8, F_CAND = 16, F_MSF = 32, F_FINAL = 64, F_STALE = 128 };
// J_RES / F_STALE (colour path, diagonal = true): builder.rs:301-305 reads cluster_upleft BEFORE the merge step of the same
// pixel, so an UPLEFT join whose up-left cluster has just lost that merge puts the pixel into the dead slot and brings it
// back to life (builder.rs:341-343).  Such a pixel starts a provisional cluster of its own that
// carries the dead label instead of a fresh one: join code J_RES.  Whether it happens depends on merge outcomes earlier in
// the scan, so the set of these pixels is found as the fixpoint of "run the pipeline, F_STALE marks the pixels where the
// stale side lost, the next pass turns exactly those into J_RES leaves" (run_stage1; the fixpoint is unique because a
// pixel's outcome depends only on earlier pixels, and every pass fixes at least the earliest wrong one).
constexpr u32 EV_LWIN = 0x100u;   // EvRec.arrive: bits 0-1 = side of the up-left cluster (stale check), bit 8 = left child won
// F_FINAL: pc[] already holds the leaf after the tile kernel (its tile-local root is a NEW pixel); J_NONE on the colour path = keyed pixel

struct Dims { u32 w, h, n, nimg, N; };

// one 32-byte sector per record; records are indexed by *pixel* (sparse): only NEW pixels own a leaf record and only
// candidate-event pixels own an event record, and each is initialised by the kernel that creates it (no memsets).
struct LeafA { u32 ufp, best0, best1, parL, cnt, death, label, flabel; };
struct LeafS { u32 sum[4]; u32 npx; u32 pad[3]; };   // pad[0]: NEW leaf: slot freed by a later incarnation; J_RES leaf: the leaf whose label it carries
                                                     // pad[1..2]: listing pair (listing.cuh)
struct LeafR { i32 minx, maxx, miny, maxy; };
struct EvRec { u32 parE, cnt, childU, arrive, a_wl, b_wu, carL, carU; };

struct ImgMeta {
    u32 leaf_begin;        // index of the image's first leaf in newlist
    u32 notd_base;         // scan value (leaves not in D) at leaf_begin
    u32 slots;             // clusters.len()
    u32 live;              // slots with area > 0
    u32 key_sum[4]; u32 key_cnt; i32 key_rect[4];   // clusters[0] under KeyingAction::Keep (builder.rs:330-334)
    u32 maxlabel;          // label of the image's last NEW event = the largest label ever handed out (clusters.rs:320-324)
    u32 pad[2];
};

struct Counters {
    u32 ncand;             // candidate events
    u32 live[2];           // boruvka: any live edge this round
    u32 err;               // bit 0: unclean key colour, bit 2: a slot freed by builder.rs:315-317 is resurrected at the same pixel
    u32 kt;                // total leaves (all images)
    u32 rounds;            // boruvka rounds executed
    u32 lt;                // total real leaves (NEW events that were not contracted)
    u32 twork;             // k_tournament work counter
    u32 cwork;             // k_chain_insert work counter
    u32 res_diff;          // pixels whose stale-upleft status differs from the J_RES set this pass ran with (0 = fixpoint)
    u32 nres;              // J_RES leaves of this pass
};

struct Bufs {
    u8 *flags; u32 *pc; LeafA *la; LeafS *ls; LeafR *lr; EvRec *ev;
    u32 *cand; u32 *ea; u32 *eb; u32 *newlist; u32 *leaflist; u32 *attr; u32 *labex; ImgMeta *meta; Counters *ctr;
};

// --------------------------------------------------------------------------------------------------------------
// pixel policies for K1
struct ColorPolicy {
    const u32 *rgba;       // RGBA8 packed little-endian: r | g<<8 | b<<16 | a<<24 (image/format.rs:29-33)
    int shift, thres, diagonal, has_key;
    u32 key;
    int use_res = 0;       // second and later passes of the stale-upleft fixpoint: flags[] still holds the previous pass
    __device__ __forceinline__ bool resurrected(const u8 *flags, u32 g) const { return use_res && (flags[g] & F_STALE); }
    bool fixpoint() const { return diagonal != 0; }
    void set_use_res() { use_res = 1; }
    static constexpr bool row_sweep = true;               // k_edges_rows (32-bit pixels, two per lane)

    __device__ __forceinline__ u32 load(u32 g) const { return __ldg(rgba + g); }
    // color_same (color_clusters/runner.rs:116-129): alpha ignored.  The three channel tests run as one byte-wise SIMD
    // absolute difference: m4 keeps the shifted r, g, b bytes (alpha cleared), t4 holds the threshold in every byte.
    u32 m4, t4;
    // The row-sweep kernel tests color_same on channels spread to 10-bit fields (r | g << 10 | b << 20): with
    //   PG(a) = P(a) | 512 per field   and   Q(b) = t per field - P(b)
    // the sum PG(a) + Q(b) holds 512 + a - b + t in every field (257 .. 1022: no field borrows from its neighbour), and
    // |a - b| <= t  <=>  512 <= field <= 512 + 2t  <=>  bit 9 set and (low nine bits + 511 - 2t) leaves bit 9 clear.
    // Five integer instructions per test instead of the emulated saturating byte subtraction of same().
    u32 t3, k3;
    static constexpr u32 F3 = 0x00100401u, G3 = 0x200u * F3, L3 = 0x1ffu * F3;
    __host__ __device__ void prepare() {
        m4 = 0x00010101u * (u32)(255 >> shift);
        const u32 t = (u32)(thres > 255 ? 255 : (thres < 0 ? 0 : thres));
        t4 = 0x00010101u * t;
        t3 = F3 * t; k3 = F3 * (511u - 2u * t);
    }
    __device__ __forceinline__ u32 pack3(u32 c) const {
        const u32 v = (c >> shift) & m4;
        return (v & 0xffu) | ((v & 0xff00u) << 2) | ((v & 0xff0000u) << 4);
    }
    __device__ __forceinline__ bool same3(u32 pg_a, u32 q_b) const {   // pg_a = pack3(a) | G3, q_b = t3 - pack3(b); thres >= 0
        const u32 x = pg_a + q_b;
        const u32 w = (x & L3) + k3;
        return ((w | ~x) & G3) == 0u;
    }
    __device__ __forceinline__ bool same(u32 a, u32 b) const {
        if (thres < 0) return false;                      // |x - y| > thres always
        const u32 dlt = __vabsdiffu4((a >> shift) & m4, (b >> shift) & m4);
        return __vsubus4(dlt, t4) == 0u;                  // no byte exceeds the threshold
    }
    // builder.rs:286-354 restated as static predicates
    __device__ __forceinline__ u32 classify(u32 c, u32 up, u32 left, u32 ul, bool vu, bool vl, u32 *err) const {
        const bool vul = vu && vl;
        const bool kc = has_key && c == key;
        u32 f = 0;
        if (has_key && !kc && same(c, key)) *err |= 1u;          // unclean key: order-dependent in the reference
        // keyed neighbours are never joined: with a clean key they are never color_same to a non-key pixel anyway, and an
        // unclean key is rejected (VCB_ERR_UNSUPPORTED) -- this only keeps the J forest well formed until it is
        const bool ku = has_key && up == key, kl = has_key && left == key, kul = has_key && ul == key;
        const bool s_cu = vu && !ku && same(c, up), s_cl = vl && !kl && same(c, left), s_cul = vul && !kul && same(c, ul);
        if (vul) {
            if (!ku && !kl && same(left, up) && (diagonal || (s_cl && s_cu))) f |= F_M;
        }
        if (kc) return f | J_NONE;
        if (s_cu && s_cul) return f | J_UP;
        if (s_cl && s_cul) return f | J_LEFT;
        if (diagonal && s_cul) return f | J_UPLEFT;
        return f | J_NEW;
    }
    __device__ __forceinline__ void stats(u32 c, u32 *s) const {
        s[0] = c & 255u; s[1] = (c >> 8) & 255u; s[2] = (c >> 16) & 255u; s[3] = c >> 24;
    }
};

struct BinaryPolicy {
    const u32 *bits;       // bit_vec::BitVec storage, pixel i at word i/32 bit i%32; one image = ceil(n/32) words
    u32 n, words_per_img;
    int diagonal;
    __device__ __forceinline__ u32 load(u32 g) const {
        const u32 img = g / n, i = g - img * n;
        return (__ldg(bits + (size_t)img * words_per_img + (i >> 5)) >> (i & 31)) & 1u;
    }
    // clusters.rs:279-326 as static predicates
    __device__ __forceinline__ u32 classify(u32 v, u32 up, u32 left, u32 ul, bool vu, bool vl, u32 *) const {
        const bool b_u = vu && up, b_l = vl && left, b_ul = vu && vl && ul;
        u32 f = 0;
        if ((v || diagonal) && b_u && b_l) f |= F_M;
        if (!v) return f | J_NONE;
        if (b_u) return f | J_UP;
        if (b_l) return f | J_LEFT;
        if (b_ul && diagonal) return f | J_UPLEFT;
        return f | J_NEW;
    }
    __device__ __forceinline__ void stats(u32, u32 *s) const { s[0] = s[1] = s[2] = s[3] = 0; }
    // immune: its upleft branch excludes v_up && v_left (clusters.rs:303-311), so an UPLEFT join never sits on a merge
    __device__ __forceinline__ bool resurrected(const u8 *, u32) const { return false; }
    bool fixpoint() const { return false; }
    void set_use_res() {}
    static constexpr bool row_sweep = false;
};

// --------------------------------------------------------------------------------------------------------------
// K1: one CTA per TW x TH tile.
//
// Contraction of trivial leaves: a NEW pixel p whose right neighbour carries a merge flag starts a cluster that is merged
// away by the very next scan step -- as the LEFT side with area 1, so it always loses (builder.rs:313: 1 <= area(up)),
// and its slot is handed to the next new cluster (builder.rs:315-317).  Nothing can happen between the two steps, so
// p is re-parented to the pixel above its right neighbour (J_RUP): it joins that cluster one step early, which changes
// no area seen by any comparison, no label and no list order, and removes one leaf and one event from the global graph
// (half of all leaves on noisy inputs).
constexpr int TW = 64, TH = 32, K1_THREADS = 256, K1_PPT = TW * TH / K1_THREADS;
constexpr int PW = TW + 8, FW = TW + 1, XO = 4;   // box = 72 x 33 pixels starting at (x0 - 4, y0 - 1): TMA needs the inner
                                                    // start coordinate and the box row to be multiples of 16 bytes (probed on B200)
constexpr size_t K1_PIX_BYTES = (size_t)(TH + 1) * PW * 4;
constexpr size_t K1_SMEM = K1_PIX_BYTES + (size_t)TW * TH * 2 + (((size_t)FW * TH + 15) & ~(size_t)15) + 16;

template <class P> __device__ __forceinline__ void edges_tile_body(const Dims &d, const P &pol, const Bufs &b, u32 *pix, u32 img,
                                                                    u32 x0, u32 y0);

// staging by plain loads (any width, binary images, and the host simulator)
template <class P> __global__ void __launch_bounds__(K1_THREADS) k_edges_tile(Dims d, P pol, Bufs b, u32 tile0) {
    VCB_SMEM(u32, pix);
    const u32 tiles_x = (d.w + TW - 1) / TW, tiles_y = (d.h + TH - 1) / TH;
    const u32 tile = blockIdx.x + tile0;
    const u32 img = tile / (tiles_x * tiles_y), tr = tile % (tiles_x * tiles_y);
    const u32 x0 = (tr % tiles_x) * TW, y0 = (tr / tiles_x) * TH;
    const u32 base = img * d.n;
    for (int idx = threadIdx.x; idx < (TH + 1) * PW; idx += K1_THREADS) {
        const int ly = idx / PW, lx = idx % PW;
        const i64 gy = (i64)y0 - 1 + ly, gx = (i64)x0 - XO + lx;
        u32 v = 0;
        if (gy >= 0 && gx >= 0 && gy < d.h && gx < d.w) v = pol.load(base + (u32)gy * d.w + (u32)gx);
        pix[idx] = v;
    }
    __syncthreads();
    edges_tile_body(d, pol, b, pix, img, x0, y0);
}

#ifndef VCB_HOSTSIM
// staging by TMA: one elected thread issues a 3-D bulk tensor copy (x, y, image) of the 68 x 33 pixel box -- halo included,
// out-of-image coordinates zero-filled by the hardware -- into shared memory and the CTA waits on an mbarrier.
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
template <class P> __global__ void __launch_bounds__(K1_THREADS)
k_edges_tile_tma(Dims d, P pol, Bufs b, const __grid_constant__ CUtensorMap tmap, u32 tile0) {
    VCB_SMEM(u32, pix);
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(pix) + K1_SMEM - 16);
    const u32 tiles_x = (d.w + TW - 1) / TW, tiles_y = (d.h + TH - 1) / TH;
    const u32 tile = blockIdx.x + tile0;
    const u32 img = tile / (tiles_x * tiles_y), tr = tile % (tiles_x * tiles_y);
    const u32 x0 = (tr % tiles_x) * TW, y0 = (tr / tiles_x) * TH;
    const u32 bar = smem_addr(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)K1_PIX_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_addr(pix)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"((int)x0 - XO), "r"((int)y0 - 1),
            "r"((int)img), "r"(bar)
            : "memory");
    }
    {
        u32 done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }
    edges_tile_body(d, pol, b, pix, img, x0, y0);
}
#endif

template <class P> __device__ __forceinline__ void edges_tile_body(const Dims &d, const P &pol, const Bufs &b, u32 *pix, u32 img,
                                                                    u32 x0, u32 y0) {
    u16 *lp = (u16 *)(pix + (TH + 1) * PW);
    u8 *fl = (u8 *)(lp + TW * TH);
    const u32 base = img * d.n;

    u32 err = 0;
    for (int i = threadIdx.x; i < FW * TH; i += K1_THREADS) {      // tile pixels + one halo column to the right (M flag only)
        const int ly = i / FW, lx = i % FW;
        const u32 gx = x0 + lx, gy = y0 + ly;
        u32 f = J_NONE;
        if (gx < d.w && gy < d.h) {
            const u32 c = pix[(ly + 1) * PW + lx + XO], up = pix[ly * PW + lx + XO];
            const u32 left = pix[(ly + 1) * PW + lx + XO - 1], ul = pix[ly * PW + lx + XO - 1];
            f = pol.classify(c, up, left, ul, gy > 0, gx > 0, &err);
            if ((f & J_MASK) == J_UPLEFT && lx < TW && pol.resurrected(b.flags, base + gy * d.w + gx)) f = (f & ~J_MASK) | J_RES;
        }
        fl[i] = (u8)f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K1_PPT; ++k) {
        const int i = threadIdx.x + k * K1_THREADS;
        const int ly = i / TW, lx = i % TW;
        const u32 gx = x0 + lx, gy = y0 + ly;
        u32 f = fl[ly * FW + lx];
        u16 l = (u16)i;
        if (gx < d.w && gy < d.h) {
            if ((f & J_MASK) == J_NEW && gx + 1 < d.w && (fl[ly * FW + lx + 1] & F_M)) {
                f = (f & ~J_MASK) | J_RUP;
                fl[ly * FW + lx] = (u8)f;                            // only the J bits of the own entry change
            }
            const u32 j = f & J_MASK;
            if (j == J_UP && ly > 0) l = (u16)(i - TW);
            else if (j == J_LEFT && lx > 0) l = (u16)(i - 1);
            else if (j == J_UPLEFT && ly > 0 && lx > 0) l = (u16)(i - TW - 1);
            else if (j == J_RUP && ly > 0 && lx + 1 < TW) l = (u16)(i - TW + 1);
        }
        lp[i] = l;
    }
    __syncthreads();
    // pointer jumping inside the tile: chains are at most TW+TH-1 long
#pragma unroll 1
    for (int it = 0; it < 7; ++it) {
        u16 nv[K1_PPT];
#pragma unroll
        for (int k = 0; k < K1_PPT; ++k) {
            const int i = threadIdx.x + k * K1_THREADS;
            nv[k] = lp[lp[i]];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K1_PPT; ++k) lp[threadIdx.x + k * K1_THREADS] = nv[k];
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < K1_PPT; ++k) {
        const int i = threadIdx.x + k * K1_THREADS;
        const int ly = i / TW, lx = i % TW;
        const u32 gx = x0 + lx, gy = y0 + ly;
        if (gx >= d.w || gy >= d.h) continue;
        const u32 g = base + gy * d.w + gx;
        const u32 f = fl[ly * FW + lx], j = f & J_MASK;
        u32 root = NONE, fin = 0;
        if (j != J_NONE) {
            const int r = lp[i];
            const u32 jr = fl[(r / TW) * FW + r % TW] & J_MASK;
            const u32 gr = base + (y0 + r / TW) * d.w + x0 + r % TW;
            const bool leaf = jr == J_NEW || jr == J_RES;
            root = leaf ? gr : jr == J_UP ? gr - d.w : jr == J_LEFT ? gr - 1 : jr == J_UPLEFT ? gr - d.w - 1 : gr - d.w + 1;
            if (leaf) fin = F_FINAL;
        }
        b.flags[g] = (u8)(f | fin);
        b.pc[g] = root;
        if (j == J_NEW || j == J_RES) {
            LeafA a; a.ufp = g; a.best0 = NONE; a.best1 = NONE; a.parL = NONE; a.cnt = 1; a.death = NONE; a.label = 0; a.flabel = 0;
            b.la[g] = a;
            LeafS s; pol.stats(pix[(ly + 1) * PW + lx + XO], s.sum); s.npx = 1; s.pad[0] = j == J_RES ? NONE : 0u; s.pad[1] = s.pad[2] = 0;
            b.ls[g] = s;
            LeafR r; r.minx = r.maxx = (i32)gx; r.miny = r.maxy = (i32)gy;
            b.lr[g] = r;
        }
    }
    if (err) atomicOr(&b.ctr->err, err);
}

// --------------------------------------------------------------------------------------------------------------
// K1, row-sweep form (colour images whose width is a multiple of 4): one WARP per TW x TH tile, no block-wide barrier.
// The warp's 72 x 33 pixel box (tile + halo) is staged in shared memory by its own TMA bulk tensor copy; the lanes then
// walk down the rows, lane l owning the columns 2l and 2l + 1.  Everything a pixel needs from its neighbourhood is in
// registers: the row above is the lane's previous iteration, the pixels to the left / right belong to the neighbouring
// lanes (one shuffle).  The tile-local J forest is resolved on the way down instead of by pointer jumping afterwards:
//   * UP / UPLEFT / RUP joins take the root the row above already carries (own register or a neighbour's);
//   * a run of LEFT joins takes the root of the run's head, found with two ballots: the chain walks left through the
//     lanes whose both pixels are LEFT joins, so the head sits in the nearest lane to the left that is not all-LEFT.
// A root is either a NEW (or resurrected) pixel of the tile -- final, F_FINAL -- or the id of the pixel outside the tile
// that the chain leaves through, exactly what k_edges_tile writes; K2 finishes those.  Per pixel this is about a sixth of
// the instructions of the pointer-jumping kernel, which was issue-bound (82 % SM throughput).
constexpr int RW_WARPS = 4, RW_THREADS = RW_WARPS * 32;
constexpr size_t RW_BOX_PITCH = (K1_PIX_BYTES + 127) & ~(size_t)127;
constexpr size_t RW_SMEM = RW_WARPS * RW_BOX_PITCH + RW_WARPS * 8 + 16;
static_assert(TW == 64, "two pixels per lane");

template <class P> __device__ __forceinline__ void edges_rows_body(const Dims &d, const P &pol, const Bufs &b, const u32 *pix, u32 img,
                                                                    u32 x0, u32 y0) {
    const u32 FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const u32 base = img * d.n;
    const u32 gx0 = x0 + 2u * (u32)lane;                 // the lane's pixels: (gx0, gy) and (gx0 + 1, gy); w is even: both or none
    const bool act = gx0 < d.w;
    const bool vl0 = gx0 > 0;
    const bool can = pol.thres >= 0;                     // a negative threshold: nothing is the same colour
    const bool diag = pol.diagonal != 0, hk = pol.has_key != 0;
    const bool halo_r = lane == 31 && x0 + TW < d.w;     // this lane also evaluates the merge flag of the column right of the tile
    u32 err = 0;
    // row above: pixels (gx0, gy - 1), (gx0 + 1, gy - 1), (gx0 - 1, gy - 1) as raw values (key tests) and as Q words;
    // box row 0 is the halo (zero-filled above the image: masked out by vu)
    u32 u0, u1, ul0, qu0, qu1, qul0;
    {
        const uint2 t = *reinterpret_cast<const uint2 *>(pix + 2 * lane + XO);
        u0 = t.x; u1 = t.y;
        ul0 = __shfl_up_sync(FULL, u1, 1);
        if (lane == 0) ul0 = pix[XO - 1];
        qu0 = pol.t3 - pol.pack3(u0); qu1 = pol.t3 - pol.pack3(u1); qul0 = pol.t3 - pol.pack3(ul0);
    }
    // roots carried by the row above: pixel id | TAG when final (a leaf of this tile), plain id of an outside pixel otherwise
    u32 pv0 = NONE, pv1 = NONE;
    const u32 rows = d.h - y0 < (u32)TH ? d.h - y0 : (u32)TH;
    for (u32 ly = 0; ly < rows; ++ly) {
        const u32 gy = y0 + ly;
        const u32 *row = pix + (ly + 1) * PW;
        const uint2 t = *reinterpret_cast<const uint2 *>(row + 2 * lane + XO);
        const u32 c0 = t.x, c1 = t.y;
        u32 l0 = __shfl_up_sync(FULL, c1, 1);
        if (lane == 0) l0 = row[XO - 1];
        const u32 p0 = pol.pack3(c0), p1 = pol.pack3(c1), pl0 = pol.pack3(l0);
        const u32 g_0 = p0 | P::G3, g_1 = p1 | P::G3, g_l0 = pl0 | P::G3;
        const u32 q_0 = pol.t3 - p0, ql0 = pol.t3 - pl0;
        // everything below is written with non-short-circuit operators and select chains: the compiler turns `&&` / nested
        // `?:` over these predicates into divergent branches, which cost far more than evaluating both sides
        const bool vu = (gy > 0) & can, v0 = vu & vl0, vl = vl0 & can;
        // color_same of (c, up), (c, left), (c, upleft), (left, up) for both pixels (runner.rs:116-129), validity folded in
        bool s_cu0 = vu & pol.same3(g_0, qu0), s_cl0 = vl & pol.same3(g_0, ql0), s_cul0 = v0 & pol.same3(g_0, qul0), s_lu0 = v0 & pol.same3(g_l0, qu0);
        bool s_cu1 = vu & pol.same3(g_1, qu1), s_cl1 = can & pol.same3(g_1, q_0), s_cul1 = vu & pol.same3(g_1, qu0), s_lu1 = vu & pol.same3(g_0, qu1);
        bool kc0 = false, kc1 = false;
        if (hk) {
            // builder.rs:330-334 and the clean-key rule of ColorPolicy::classify: keyed pixels are never joined
            const u32 key = pol.key;
            kc0 = c0 == key; kc1 = c1 == key;
            const bool ku0 = u0 == key, ku1 = u1 == key, kl0 = l0 == key, kul0 = ul0 == key;
            const u32 qk = pol.t3 - pol.pack3(key);
            if (act & can & ((!kc0 & pol.same3(g_0, qk)) | (!kc1 & pol.same3(g_1, qk)))) err |= 1u;
            s_cu0 &= !ku0; s_cl0 &= !kl0; s_cul0 &= !kul0; s_lu0 &= !ku0 & !kl0;
            s_cu1 &= !ku1; s_cl1 &= !kc0; s_cul1 &= !ku0; s_lu1 &= !ku1 & !kc0;
        }
        const bool m0 = s_lu0 & (diag | (s_cl0 & s_cu0)), m1 = s_lu1 & (diag | (s_cl1 & s_cu1));
        u32 j0 = J_NEW, j1 = J_NEW;
        j0 = (diag & s_cul0) ? (u32)J_UPLEFT : j0; j0 = (s_cl0 & s_cul0) ? (u32)J_LEFT : j0; j0 = (s_cu0 & s_cul0) ? (u32)J_UP : j0;
        j1 = (diag & s_cul1) ? (u32)J_UPLEFT : j1; j1 = (s_cl1 & s_cul1) ? (u32)J_LEFT : j1; j1 = (s_cu1 & s_cul1) ? (u32)J_UP : j1;
        j0 = (kc0 | !act) ? (u32)J_NONE : j0; j1 = (kc1 | !act) ? (u32)J_NONE : j1;
        // merge flag of the pixel to the right of pixel 1 (trivial-leaf contraction): the next lane's pixel 0, or the halo column
        bool mr = __shfl_down_sync(FULL, (int)(act & m0), 1) != 0;
        mr &= lane != 31;
        if (halo_r) {
            const u32 hc = row[XO + TW], hu = row[XO + TW - PW];
            const u32 gh = pol.pack3(hc) | P::G3, qhu = pol.t3 - pol.pack3(hu);
            bool s_lu = vu & pol.same3(g_1, qhu);
            const bool s_cl = can & pol.same3(gh, pol.t3 - p1), s_cu = vu & pol.same3(gh, qhu);
            if (hk) s_lu &= (hu != pol.key) & !kc1;
            mr = s_lu & (diag | (s_cl & s_cu));
        }
        const u32 g0 = base + gy * d.w + gx0;
        if (diag & act & pol.use_res) {
            if (j0 == J_UPLEFT && pol.resurrected(b.flags, g0)) j0 = J_RES;
            if (j1 == J_UPLEFT && pol.resurrected(b.flags, g0 + 1)) j1 = J_RES;
        }
        j0 = ((j0 == J_NEW) & m1) ? (u32)J_RUP : j0;
        j1 = ((j1 == J_NEW) & mr) ? (u32)J_RUP : j1;
        const u32 pvL = __shfl_up_sync(FULL, pv1, 1);        // root of (gx0 - 1, gy - 1)
        const u32 pvR = __shfl_down_sync(FULL, pv0, 1);      // root of (gx0 + 2, gy - 1)
        const bool top = ly == 0;
        const u32 gup = g0 - d.w;
        // root candidates of the three upward joins (outside the tile: the id of the pixel joined), then one select chain
        const u32 a_up0 = top ? gup : pv0, a_ul0 = (top | (lane == 0)) ? gup - 1 : pvL, a_ru0 = top ? gup + 1 : pv1;
        const u32 a_up1 = top ? gup + 1 : pv1, a_ul1 = top ? gup : pv0, a_ru1 = (top | (lane == 31)) ? gup + 2 : pvR;
        u32 r0 = g0 | TAG, r1 = (g0 + 1) | TAG;
        r0 = j0 == J_UP ? a_up0 : r0; r0 = j0 == J_UPLEFT ? a_ul0 : r0; r0 = j0 == J_RUP ? a_ru0 : r0;
        r1 = j1 == J_UP ? a_up1 : r1; r1 = j1 == J_UPLEFT ? a_ul1 : r1; r1 = j1 == J_RUP ? a_ru1 : r1;
        // LEFT joins: the chain from pixel 0 goes to pixel 1 of the lane to the left and on through all-LEFT lanes
        const u32 mA = __ballot_sync(FULL, j0 == J_LEFT), mB = __ballot_sync(FULL, j1 == J_LEFT);
        const u32 open = ~(mA & mB) & ((1u << lane) - 1u);  // lanes to the left where the chain stops
        const int hl = open ? 31 - __clz((int)open) : 0;
        const u32 h0 = __shfl_sync(FULL, r0, hl), h1 = __shfl_sync(FULL, r1, hl);
        u32 head = ((mB >> hl) & 1u) ? h0 : h1;             // the head is pixel 1 of that lane unless that is itself a LEFT join
        head = open ? head : g0 - 2u * (u32)lane - 1u;      // no such lane: the chain leaves the tile through its left edge
        r0 = j0 == J_LEFT ? head : r0;
        r1 = j1 == J_LEFT ? r0 : r1;
        if (act) {
            const u32 f0 = j0 | (m0 ? (u32)F_M : 0u) | (((j0 != J_NONE) & ((r0 & TAG) != 0u)) ? (u32)F_FINAL : 0u);
            const u32 f1 = j1 | (m1 ? (u32)F_M : 0u) | (((j1 != J_NONE) & ((r1 & TAG) != 0u)) ? (u32)F_FINAL : 0u);
            *reinterpret_cast<u16 *>(b.flags + g0) = (u16)(f0 | (f1 << 8));
            *reinterpret_cast<uint2 *>(b.pc + g0) = make_uint2(j0 == J_NONE ? NONE : (r0 & ~TAG), j1 == J_NONE ? NONE : (r1 & ~TAG));
            const bool n0 = (j0 == J_NEW) | (j0 == J_RES), n1 = (j1 == J_NEW) | (j1 == J_RES);
            if (n0 | n1) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (!(k ? n1 : n0)) continue;
                    const u32 g = g0 + k, c = k ? c1 : c0;
                    uint4 *pa = reinterpret_cast<uint4 *>(b.la + g), *ps = reinterpret_cast<uint4 *>(b.ls + g);
                    pa[0] = make_uint4(g, NONE, NONE, NONE);         // ufp, best0, best1, parL
                    pa[1] = make_uint4(1u, NONE, 0u, 0u);            // cnt, death, label, flabel
                    ps[0] = make_uint4(c & 255u, (c >> 8) & 255u, (c >> 16) & 255u, c >> 24);     // ColorPolicy::stats
                    ps[1] = make_uint4(1u, (k ? j1 : j0) == J_RES ? NONE : 0u, 0u, 0u);           // npx, pad[0..2]
                    *reinterpret_cast<uint4 *>(b.lr + g) = make_uint4(gx0 + k, gx0 + k, gy, gy);      // minx, maxx, miny, maxy
                }
            }
        }
        ul0 = l0; u0 = c0; u1 = c1; qul0 = ql0; qu0 = q_0; qu1 = pol.t3 - p1; pv0 = r0; pv1 = r1;
    }
    if (err) atomicOr(&b.ctr->err, err);
}

// staging by plain loads (host simulator, and when the tensor map cannot be encoded)
template <class P> __global__ void __launch_bounds__(RW_THREADS) k_edges_rows(Dims d, P pol, Bufs b, u32 tile0, u32 ntiles) {
    VCB_SMEM(unsigned char, sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 t = blockIdx.x * RW_WARPS + warp;
    if (t >= ntiles) return;
    const u32 tiles_x = (d.w + TW - 1) / TW, tiles_y = (d.h + TH - 1) / TH;
    const u32 tile = t + tile0;
    const u32 img = tile / (tiles_x * tiles_y), tr = tile % (tiles_x * tiles_y);
    const u32 x0 = (tr % tiles_x) * TW, y0 = (tr / tiles_x) * TH;
    u32 *pix = reinterpret_cast<u32 *>(sm + warp * RW_BOX_PITCH);
    const u32 base = img * d.n;
    for (int idx = lane; idx < (TH + 1) * PW; idx += 32) {
        const int ly = idx / PW, lx = idx % PW;
        const i64 gy = (i64)y0 - 1 + ly, gx = (i64)x0 - XO + lx;
        u32 v = 0;
        if (gy >= 0 && gx >= 0 && gy < d.h && gx < d.w) v = pol.load(base + (u32)gy * d.w + (u32)gx);
        pix[idx] = v;
    }
    __syncwarp();
    edges_rows_body(d, pol, b, pix, img, x0, y0);
}

#ifndef VCB_HOSTSIM
template <class P> __global__ void __launch_bounds__(RW_THREADS)
k_edges_rows_tma(Dims d, P pol, Bufs b, const __grid_constant__ CUtensorMap tmap, u32 tile0, u32 ntiles) {
    VCB_SMEM(unsigned char, sm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 t = blockIdx.x * RW_WARPS + warp;
    if (t >= ntiles) return;                              // whole warps leave: nothing below synchronises the block
    const u32 tiles_x = (d.w + TW - 1) / TW, tiles_y = (d.h + TH - 1) / TH;
    const u32 tile = t + tile0;
    const u32 img = tile / (tiles_x * tiles_y), tr = tile % (tiles_x * tiles_y);
    const u32 x0 = (tr % tiles_x) * TW, y0 = (tr / tiles_x) * TH;
    u32 *pix = reinterpret_cast<u32 *>(sm + warp * RW_BOX_PITCH);
    const u32 bar = smem_addr(sm + RW_WARPS * RW_BOX_PITCH + warp * 8);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)K1_PIX_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_addr(pix)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"((int)x0 - XO), "r"((int)y0 - 1),
            "r"((int)img), "r"(bar)
            : "memory");
    }
    __syncwarp();
    {
        u32 done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }
    edges_rows_body(d, pol, b, pix, img, x0, y0);
}
#endif

// --------------------------------------------------------------------------------------------------------------
// K2: one thread per pixel.  pc[] -> provisional-cluster root (leaf id); candidate events.
__device__ static __forceinline__ u32 pc_root(const u32 *pc, u32 q) {
    u32 r = ldcg(pc + q);
    while (true) {
        const u32 t = ldcg(pc + r);
        if (t == r) return r;
        r = t;
    }
}
// leaf of pixel q during K2: one load when the tile kernel already resolved it
__device__ static __forceinline__ u32 pc_root_f(const u32 *pc, const u8 *flags, u32 q) {
    return (flags[q] & F_FINAL) ? ldcg(pc + q) : pc_root(pc, q);
}

constexpr int K2_THREADS = 256;
__global__ void __launch_bounds__(K2_THREADS) k_flatten_events(Dims d, Bufs b, u32 p_lo, u32 p_hi) {
    const u32 stride = gridDim.x * K2_THREADS;
    const int lane = threadIdx.x & 31;
    const u32 nloop = (p_hi - p_lo + stride - 1) / stride;
    for (u32 it = 0; it < nloop; ++it) {
        const u32 g = p_lo + it * stride + blockIdx.x * K2_THREADS + threadIdx.x;
        bool is_cand = false;
        u32 ea = 0, eb = 0;
        if (g < p_hi) {
            const u32 f = b.flags[g];
            if ((f & J_MASK) != J_NONE && !(f & F_FINAL)) {
                const u32 r = pc_root(b.pc, g);
                stcg(b.pc + g, r);
            }
            if (f & F_M) {
                // duplicate filter: the event one row up joined the same pair of provisional clusters earlier
                const u32 fl = b.flags[g - 1], fu = b.flags[g - d.w];
                const bool dup = (fl & J_MASK) == J_UP && (fu & J_MASK) == J_UP && (fu & F_M);
                if (!dup) {
                    ea = pc_root_f(b.pc, b.flags, g - 1);
                    eb = pc_root_f(b.pc, b.flags, g - d.w);
                    is_cand = ea != eb;
                }
            }
        }
        const u32 m = __ballot_sync(0xffffffffu, is_cand);
        if (m) {
            u32 pos = 0;
            if (lane == __ffs((int)m) - 1) pos = atomicAdd(&b.ctr->ncand, (u32)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, __ffs((int)m) - 1);
            if (is_cand) {
                const u32 slot = pos + __popc(m & ((1u << lane) - 1u));
                b.cand[slot] = g; b.ea[slot] = ea; b.eb[slot] = eb;     // dense endpoint arrays for the spanning-forest rounds
                EvRec e; e.parE = NONE; e.cnt = 0; e.childU = NONE; e.arrive = 0; e.a_wl = 0; e.b_wu = 0; e.carL = NONE; e.carU = NONE;
                b.ev[g] = e;
                b.flags[g] = (u8)(b.flags[g] | F_CAND);
            }
        }
    }
}

// K2 for widths that are a multiple of 4: one thread per 4 consecutive pixels of a row, flags read as one 32-bit word
// (own row and the row above), candidates compacted with one atomic per warp.
__global__ void __launch_bounds__(K2_THREADS) k_flatten_events4(Dims d, Bufs b, u32 grp_lo, u32 grp_hi) {
    const u32 groups = grp_hi;
    const u32 stride = gridDim.x * K2_THREADS;
    const int lane = threadIdx.x & 31;
    const u32 nloop = (grp_hi - grp_lo + stride - 1) / stride;
    for (u32 it = 0; it < nloop; ++it) {
        const u32 grp = grp_lo + it * stride + blockIdx.x * K2_THREADS + threadIdx.x;
        u32 ncand = 0, cg[4], ca[4], cb[4];
        if (grp < groups) {
            const u32 g4 = grp * 4;
            const u32 f4 = *reinterpret_cast<const u32 *>(b.flags + g4);
            // bit 8k set: pixel k still needs the global resolution (joined, and its tile-local root is not a NEW pixel)
            u32 open = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 f = (f4 >> (8 * k)) & 255u;
                if ((f & J_MASK) != J_NONE && !(f & F_FINAL)) open |= 1u << k;
            }
            const bool any_m = (f4 & 0x08080808u) != 0;
            u32 root[4] = {NONE, NONE, NONE, NONE};
            if (open || (f4 & 0x08080800u)) {             // own leaves are needed: to be resolved, or as the left side of an event
                const uint4 raw = *reinterpret_cast<const uint4 *>(b.pc + g4);
                root[0] = raw.x; root[1] = raw.y; root[2] = raw.z; root[3] = raw.w;
                // neighbouring pixels of one region share the tile-local root: walk each distinct chain once
                u32 last_raw = NONE, last_root = NONE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (!(open & (1u << k))) continue;
                    const u32 rw = root[k];
                    if (rw != last_raw) {
                        u32 r = rw;
                        while (true) { const u32 t = ldcg(b.pc + r); if (t == r) break; r = t; }
                        last_raw = rw; last_root = r;
                    }
                    root[k] = last_root;
                    stcg(b.pc + g4 + k, last_root);
                }
            }
            if (any_m) {                                  // some pixel carries a merge flag (implies a row above)
                const u32 fu4 = *reinterpret_cast<const u32 *>(b.flags + g4 - d.w);
                const uint4 upraw = *reinterpret_cast<const uint4 *>(b.pc + g4 - d.w);     // w % 4 == 0: aligned
                const u32 up[4] = {upraw.x, upraw.y, upraw.z, upraw.w};
                u32 fleft = 0;
                if (f4 & F_M) fleft = b.flags[g4 - 1];
                u32 last_raw = NONE, last_root = NONE;    // the pixels above usually share one tile-local root as well
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const u32 f = (f4 >> (8 * k)) & 255u;
                    if (!(f & F_M)) continue;
                    const u32 fl = k ? (f4 >> (8 * (k - 1))) & 255u : fleft;
                    const u32 fu = (fu4 >> (8 * k)) & 255u;
                    if ((fl & J_MASK) == J_UP && (fu & J_MASK) == J_UP && (fu & F_M)) continue;   // duplicate of the event one row up
                    const u32 g = g4 + k;
                    // a merge flag implies joined (non-keyed / set) left and up pixels, so both have a leaf
                    const u32 ea = k ? root[k - 1] : ((fl & F_FINAL) ? ldcg(b.pc + g - 1) : pc_root(b.pc, g - 1));
                    u32 eb = up[k];                       // a stale or a resolved value: both lead to the same root
                    if (!(fu & F_FINAL)) {
                        if (eb != last_raw) {
                            u32 r = eb;
                            while (true) { const u32 t = ldcg(b.pc + r); if (t == r) break; r = t; }
                            last_raw = eb; last_root = r;
                        }
                        eb = last_root;
                    }
                    if (ea != eb) { cg[ncand] = g; ca[ncand] = ea; cb[ncand] = eb; ++ncand; }
                }
            }
        }
        // warp-aggregated append
        u32 inc = ncand;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, inc, dlt); if (lane >= dlt) inc += t; }
        const u32 total = __shfl_sync(0xffffffffu, inc, 31);
        if (total) {
            u32 pos = 0;
            if (lane == 31) pos = atomicAdd(&b.ctr->ncand, total);
            pos = __shfl_sync(0xffffffffu, pos, 31) + inc - ncand;
            for (u32 k = 0; k < ncand; ++k) {
                const u32 g = cg[k];
                b.cand[pos + k] = g; b.ea[pos + k] = ca[k]; b.eb[pos + k] = cb[k];
                EvRec e; e.parE = NONE; e.cnt = 0; e.childU = NONE; e.arrive = 0; e.a_wl = 0; e.b_wu = 0; e.carL = NONE; e.carU = NONE;
                b.ev[g] = e;
                b.flags[g] = (u8)(b.flags[g] | F_CAND);
            }
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// K3: Boruvka minimum spanning forest over (leaves, candidate events), weight = event pixel index (= time; unique).
// An event of the reference is effective (its two sides are different clusters when the scan reaches it) iff it is
// an MSF edge.  Cooperative kernel: two grid syncs per round, O(log K) rounds.
__device__ static __forceinline__ u32 uf_find(LeafA *la, u32 x) {
    while (true) {
        const u32 p = ldcg(&la[x].ufp);
        if (p == x) return x;
        const u32 gp = ldcg(&la[p].ufp);
        if (gp == p) return p;
        stcg(&la[x].ufp, gp);
        x = gp;
    }
}

constexpr int K3_THREADS = 256;
__global__ void __launch_bounds__(K3_THREADS) k_boruvka(Bufs b) {
    const u32 ncand = ldcg(&b.ctr->ncand);
    const u32 tid = blockIdx.x * K3_THREADS + threadIdx.x, stride = gridDim.x * K3_THREADS;
    u32 round = 0;
    while (true) {
        const int par = (int)(round & 1u);
        if (tid == 0) stcg(&b.ctr->live[par ^ 1], 0u);
        bool any = false;
        for (u32 e = tid; e < ncand; e += stride) {
            const u32 p = ldcg(b.cand + e);
            if (p & TAG) continue;
            const u32 ra = uf_find(b.la, b.ea[e]), rb = uf_find(b.la, b.eb[e]);
            if (ra == rb) { b.cand[e] = p | TAG; continue; }
            b.ea[e] = ra; b.eb[e] = rb;
            // (best0, best1) sit in one sector: one load tells whether this edge can still win and whether the other
            // parity still has to be cleared, so most edges of a big component issue no atomic and no store at all
            u32 *ba = par ? &b.la[ra].best1 : &b.la[ra].best0, *bb = par ? &b.la[rb].best1 : &b.la[rb].best0;
            u32 *oa = par ? &b.la[ra].best0 : &b.la[ra].best1, *ob = par ? &b.la[rb].best0 : &b.la[rb].best1;
            if (ldcg(ba) > p) atomicMin(ba, p);
            if (ldcg(bb) > p) atomicMin(bb, p);
            if (ldcg(oa) != NONE) stcg(oa, NONE);
            if (ldcg(ob) != NONE) stcg(ob, NONE);
            any = true;
        }
        if (any) stcg(&b.ctr->live[par], 1u);
        __threadfence();
        VCB_GRID_SYNC();
        if (ldcg(&b.ctr->live[par]) == 0) break;
        for (u32 e = tid; e < ncand; e += stride) {
            const u32 p = ldcg(b.cand + e);
            if (p & TAG) continue;
            const u32 ra = b.ea[e], rb = b.eb[e];
            const bool sa = ldcg(par ? &b.la[ra].best1 : &b.la[ra].best0) == p;
            const bool sb = ldcg(par ? &b.la[rb].best1 : &b.la[rb].best0) == p;
            if (sa || sb) {
                if (sa && sb) { const u32 hi = ra > rb ? ra : rb, lo = ra > rb ? rb : ra; stcg(&b.la[hi].ufp, lo); }
                else if (sa) stcg(&b.la[ra].ufp, rb);
                else stcg(&b.la[rb].ufp, ra);
                b.flags[p] = (u8)(b.flags[p] | F_MSF);
                b.cand[e] = p | TAG;
            }
        }
        __threadfence();
        VCB_GRID_SYNC();
        ++round;
    }
    if (tid == 0) b.ctr->rounds = round;
    // epilogue: every leaf points directly at the root of its final cluster (later finds are one load)
    const u32 lt = ldcg(&b.ctr->lt);
    for (u32 k = tid; k < lt; k += stride) {
        const u32 leaf = b.leaflist[k];
        const u32 r = uf_find(b.la, leaf);
        if (r != leaf) stcg(&b.la[leaf].ufp, r);
    }
}

// --------------------------------------------------------------------------------------------------------------
// K4: Kruskal reconstruction tree by lock-free sorted-chain insertion.  Fact(f -> g): event g is an ancestor of node
// f.  Every effective event g = (a, b) contributes Fact(leaf a -> g) and Fact(leaf b -> g).  par[f] always holds the
// lightest ancestor of f known so far; inserting g above f either walks up (par[f] < g: g is also an ancestor of
// par[f]) or installs g and re-inserts the displaced ancestor above g.  At quiescence par[] is the tree.
__device__ static __forceinline__ void chain_insert(Bufs b, u32 leaf, u32 g) {
    u32 *slot = &b.la[leaf].parL;
    while (true) {
        const u32 old = atomicMin(slot, g);               // one round trip per step
        if (old == g) return;
        if (old > g) {                                    // installed g below `old` (NONE included)
            if (old == NONE) return;
            slot = &b.ev[g].parE; g = old;                // now insert the displaced ancestor above g
        } else {
            slot = &b.ev[old].parE;                       // a lighter ancestor is already there: g is also its ancestor
        }
    }
}

// Warp-level dynamic work claiming: the warp takes WORK_CHUNK consecutive items from a global counter with one atomic and
// hands them to its lanes as they become free.  Returns the item index for this lane or NONE.
constexpr u32 WORK_CHUNK = 2048;
struct WarpWork {
    u32 pos = 0, end = 0;      // warp-uniform
    bool exhausted = false;
    __device__ __forceinline__ u32 claim(u32 *counter, u32 total, bool want, int lane) {
        const u32 need = __ballot_sync(0xffffffffu, want);
        if (!need) return NONE;
        if (pos >= end && !exhausted) {
            u32 base = 0;
            if (lane == 0) base = atomicAdd(counter, WORK_CHUNK);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= total) exhausted = true;
            else { pos = base; end = base + WORK_CHUNK < total ? base + WORK_CHUNK : total; }
        }
        u32 idx = NONE;
        if (pos < end) {
            const u32 mine = pos + __popc(need & ((1u << lane) - 1u));
            if (want && mine < end) idx = mine;
            pos += __popc(need);
        }
        return idx;
    }
};

constexpr int K4_THREADS = 256;
// One insertion step per lane per iteration; a lane whose insertion is complete claims the next (event, side) item from
// a global counter, so the few long walks do not idle the rest of the warp.
__global__ void __launch_bounds__(K4_THREADS) k_chain_insert(Dims d, Bufs b) {
    const u32 nitems = 2 * ldcg(&b.ctr->ncand);
    const int lane = threadIdx.x & 31;
    bool have = false;
    u32 *slot = nullptr;
    u32 g = 0;
    WarpWork ww;
    while (true) {
        const u32 t = ww.claim(&b.ctr->cwork, nitems, !have, lane);
        if (t != NONE) {
            const u32 p = b.cand[t >> 1] & ~TAG;
            if (b.flags[p] & F_MSF) {
                const u32 leaf = (t & 1u) ? b.pc[p - d.w] : b.pc[p - 1];
                slot = &b.la[leaf].parL; g = p; have = true;
            }
        }
        if (__all_sync(0xffffffffu, !have)) {
            // nobody holds a live item: either the list is exhausted or this batch held only non-effective candidates
            if (ww.exhausted && ww.pos >= ww.end) break;
            continue;
        }
        const bool drained = ww.exhausted && ww.pos >= ww.end;   // nothing left to claim: finish the own walk in a tight loop
        while (have) {
            // walking up past lighter ancestors needs only loads; the atomic is issued when g has to be installed
            u32 old = ldcg(slot);
            if (old > g) old = atomicMin(slot, g);
            if (old == g || old == NONE) { have = false; break; }
            if (old > g) { slot = &b.ev[g].parE; g = old; }   // installed g below `old`: now insert the displaced ancestor above g
            else slot = &b.ev[old].parE;                  // a lighter ancestor is already there: g is also its ancestor
            if (!drained) break;
        }
        if (drained) break;
    }
}

// --------------------------------------------------------------------------------------------------------------
// K5: attribution.  One thread per (image, row segment, column); a warp covers 32 adjacent columns and walks down
// its rows in lock step.  g(p) = highest ancestor of leaf(pc[p]) born no later than p; it is reached by climbing from
// the node of the pixel's J-parent (the pixel above: kept in a register; the pixel to the left: warp shuffle).
__device__ static __forceinline__ u32 node_par(const Bufs &b, u32 node) {
    return (node & TAG) ? ldcg(&b.la[node & ~TAG].parL) : ldcg(&b.ev[node].parE);
}
// the tree is final and read-only from K5 on: parents through the read-only path
__device__ static __forceinline__ u32 node_par_ro(const Bufs &b, u32 node) {
    return (node & TAG) ? __ldg(&b.la[node & ~TAG].parL) : __ldg(&b.ev[node].parE);
}
__device__ static __forceinline__ u32 climb(const Bufs &b, u32 node, u32 bound) {   // parents born <= bound
    while (true) {
        const u32 q = node_par(b, node);
        if (q > bound) return node;                      // NONE > anything
        node = q;
    }
}
// (node, par) pair kept in registers: no load at all while the pixel column stays on the same node
// The record of the next event up is requested as soon as its id is known: the column usually reaches it rows later, and the
// climb step is then an L2 hit instead of a first-touch DRAM miss (k_attribute 3.49 -> 2.87 ms at 64 x 1080p).
__device__ static __forceinline__ void prefetch_event(const Bufs &b, u32 ev_id) {
#ifndef VCB_HOSTSIM
    if (ev_id != NONE) asm volatile("prefetch.global.L2 [%0];" ::"l"(&b.ev[ev_id].parE));
#endif
}
__device__ static __forceinline__ void climb2(const Bufs &b, u32 &node, u32 &par, u32 bound) {
    if (par > bound) return;
    do { node = par; par = __ldg(&b.ev[node].parE); } while (par <= bound);
    prefetch_event(b, par);
}

struct LeafRun {
    u32 leaf; u32 s[4]; u32 n; i32 miny, maxy;
    __device__ __forceinline__ void flush(const Bufs &b, i32 x) {
        if (leaf == NONE || n == 0) return;
        LeafS *ls = b.ls + leaf;
        if (s[0]) atomicAdd(&ls->sum[0], s[0]);
        if (s[1]) atomicAdd(&ls->sum[1], s[1]);
        if (s[2]) atomicAdd(&ls->sum[2], s[2]);
        if (s[3]) atomicAdd(&ls->sum[3], s[3]);
        atomicAdd(&ls->npx, n);
        LeafR *lr = b.lr + leaf;
        atomicMin(&lr->minx, x); atomicMax(&lr->maxx, x);
        atomicMin(&lr->miny, miny); atomicMax(&lr->maxy, maxy);
        n = 0;
    }
};

constexpr int K5_THREADS = 128;
template <class P, bool WRITE_ATTR>
__global__ void __launch_bounds__(K5_THREADS) k_attribute(Dims d, P pol, Bufs b, u32 rows_per_seg, int keep_key, u64 t_lo, u64 t_hi) {
    const u32 wpad = (d.w + 31) & ~31u;
    const u32 nseg = (d.h + rows_per_seg - 1) / rows_per_seg;
    const u64 total = t_hi;                               // work items [t_lo, t_hi) of nimg * nseg * wpad (multiples of 32)
    const u64 stride = (u64)gridDim.x * K5_THREADS;
    const int lane = threadIdx.x & 31;
    for (u64 t0 = t_lo + (u64)blockIdx.x * K5_THREADS + threadIdx.x - lane; t0 < total; t0 += stride) {
        const u64 t = t0 + lane;
        const u32 x = (u32)(t % wpad);
        const u32 seg = (u32)((t / wpad) % nseg), img = (u32)(t / ((u64)wpad * nseg));
        const bool valid = x < d.w;
        const u32 y0 = seg * rows_per_seg, y1 = (y0 + rows_per_seg < d.h) ? y0 + rows_per_seg : d.h;
        u32 gprev = NONE, gprev_par = NONE;              // node of the pixel above and its parent (NONE: unknown / keyed)
        u32 run_node = NONE, run_cnt = 0;
        LeafRun lr; lr.leaf = NONE; lr.n = 0; lr.s[0] = lr.s[1] = lr.s[2] = lr.s[3] = 0; lr.miny = 0; lr.maxy = 0;
        u32 ks[4] = {0, 0, 0, 0}, kn = 0; i32 kminy = 0x7fffffff, kmaxy = -1;
        // the loads of row y + 1 are issued before row y is processed (the per-row warp votes would otherwise serialise them)
        u32 nf = J_NONE, npc = NONE, npix = 0;
        if (valid && y0 < y1) {
            const u32 p0 = img * d.n + y0 * d.w + x;
            nf = b.flags[p0]; npc = b.pc[p0]; npix = pol.load(p0);
        }
        for (u32 y = y0; y < y1; ++y) {
            const u32 p = img * d.n + y * d.w + (valid ? x : 0);
            const u32 f = nf, pc_p = npc, pix_p = npix;
            if (valid && y + 1 < y1) { nf = b.flags[p + d.w]; npc = b.pc[p + d.w]; npix = pol.load(p + d.w); }
            const u32 j = f & J_MASK;
            u32 g = NONE, gpar = NONE;
            const u32 gl_n = __shfl_up_sync(0xffffffffu, gprev, 1), gl_p = __shfl_up_sync(0xffffffffu, gprev_par, 1);     // (x-1, y-1)
            const u32 gr_n = __shfl_down_sync(0xffffffffu, gprev, 1), gr_p = __shfl_down_sync(0xffffffffu, gprev_par, 1); // (x+1, y-1)
            // the up-side child of an effective event at p: climb the chain of pixel p-w to just below p
            u32 upnode = gprev, uppar = gprev_par;
            if (valid && (f & F_MSF)) {
                if (upnode == NONE) { upnode = b.pc[p - d.w] | TAG; uppar = node_par_ro(b, upnode); }
                climb2(b, upnode, uppar, p - 1);
                stcg(&b.ev[p].childU, upnode);
                if (j == J_UPLEFT || j == J_RES) {
                    // builder.rs:301-305 reads cluster_upleft BEFORE the merge step: if the up-left pixel's cluster is one of
                    // the two merging sides and loses, the reference labels this pixel with the dead slot (SURVEY Appendix E
                    // item 5).  Record which side it is on; k_tournament marks the pixel F_STALE if that side loses.
                    u32 ul = gl_n, ulp = gl_p;
                    if (lane == 0 || ul == NONE) { ul = b.pc[p - d.w - 1] | TAG; ulp = node_par_ro(b, ul); }
                    climb2(b, ul, ulp, p - 1);
                    u32 side = 0;
                    if (ulp == p) side = (ul == upnode) ? 2u : 1u;
                    stcg(&b.ev[p].arrive, side);
                }
            }
            bool resolved = true;
            if (valid && j != J_NONE) {
                if (j == J_NEW || j == J_RES) { g = p | TAG; gpar = __ldg(&b.la[p].parL); prefetch_event(b, gpar); }
                else if (j == J_LEFT && lane > 0) resolved = false;          // wait for the lane to the left
                else {
                    if (j == J_UP && upnode != NONE) { g = upnode; gpar = uppar; }
                    else if (j == J_UPLEFT && lane > 0 && gl_n != NONE) { g = gl_n; gpar = gl_p; }
                    else if (j == J_RUP && lane < 31 && gr_n != NONE) { g = gr_n; gpar = gr_p; }
                    else { g = pc_p | TAG; gpar = node_par_ro(b, g); prefetch_event(b, gpar); }      // fresh start from the leaf
                    climb2(b, g, gpar, p);
                }
            }
            while (true) {
                const u32 gl = __shfl_up_sync(0xffffffffu, g, 1), glp = __shfl_up_sync(0xffffffffu, gpar, 1);
                const int rl = __shfl_up_sync(0xffffffffu, (int)resolved, 1);
                if (!resolved && rl) { g = gl; gpar = glp; climb2(b, g, gpar, p); resolved = true; }
                if (__all_sync(0xffffffffu, resolved)) break;
            }
            if (valid) {
                if (j == J_NONE) {
                    if (keep_key) {                                  // colour path: J_NONE = keyed pixel (builder.rs:330-334)
                        u32 s[4]; pol.stats(pix_p, s);
                        ks[0] += s[0]; ks[1] += s[1]; ks[2] += s[2]; ks[3] += s[3]; kn++;
                        if ((i32)y < kminy) kminy = (i32)y;
                        kmaxy = (i32)y;
                    }
                } else {
                    if (WRITE_ATTR) b.attr[p] = g;
                    if (j != J_NEW && j != J_RES) {
                        if (g == run_node) run_cnt++;
                        else {
                            if (run_cnt) atomicAdd((run_node & TAG) ? &b.la[run_node & ~TAG].cnt : &b.ev[run_node].cnt, run_cnt);
                            run_node = g; run_cnt = 1;
                        }
                        const u32 leaf = pc_p;
                        if (leaf != lr.leaf) { lr.flush(b, (i32)x); lr.leaf = leaf; lr.miny = (i32)y; lr.s[0] = lr.s[1] = lr.s[2] = lr.s[3] = 0; }
                        u32 s[4]; pol.stats(pix_p, s);
                        lr.s[0] += s[0]; lr.s[1] += s[1]; lr.s[2] += s[2]; lr.s[3] += s[3]; lr.n++; lr.maxy = (i32)y;
                    }
                }
            }
            gprev = g; gprev_par = gpar;
        }
        if (run_cnt) atomicAdd((run_node & TAG) ? &b.la[run_node & ~TAG].cnt : &b.ev[run_node].cnt, run_cnt);
        lr.flush(b, (i32)x);
        if (kn) {
            ImgMeta *m = b.meta + img;
            atomicAdd(&m->key_sum[0], ks[0]); atomicAdd(&m->key_sum[1], ks[1]);
            atomicAdd(&m->key_sum[2], ks[2]); atomicAdd(&m->key_sum[3], ks[3]);
            atomicAdd(&m->key_cnt, kn);
            atomicMin(&m->key_rect[0], (i32)x); atomicMax(&m->key_rect[1], (i32)x);
            atomicMin(&m->key_rect[2], kminy); atomicMax(&m->key_rect[3], kmaxy);
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// ordered compaction of NEW pixels (scan functor) -> newlist, per-image leaf_begin
struct NewScan {
    Dims d; Bufs b;
    __device__ u32 count() const { return d.N; }
    __device__ u32 value(u32 i) const { const u32 j = b.flags[i] & J_MASK; return (j == J_NEW || j == J_RUP) ? 1u : 0u; }
    __device__ void emit(u32 i, u32 ex, u32 v) const {
        // contracted (trivial) NEW events stay in the list, tagged: they own no leaf record but they are NEW events of the
        // reference, so they bound the "next NEW" window of the numbering and they are always in D
        if (v) b.newlist[ex] = i | (((b.flags[i] & J_MASK) == J_RUP) ? TAG : 0u);
        if (i % d.n == 0) b.meta[i / d.n].leaf_begin = ex;
        if (i == d.N - 1) b.ctr->kt = ex + v;
    }
};

// Both compactions in one two-pass scan: 8 flag bytes per thread read as one 64-bit word, (NEW events, real leaves)
// counted side by side in the two halves of a 64-bit sum.  Used for whole batches; NewScan / LeafScan remain the
// restatement the simulator tests compare against (VCB_NO_FUSED_SCAN=1).
constexpr int NL_THREADS = 256, NL_ITEMS = 8, NL_WORDS = 4, NL_CHUNK = NL_THREADS * NL_ITEMS * NL_WORDS;   // 32 pixels per thread
__device__ static __forceinline__ u64 nl_load(const Bufs &b, u32 i0, u32 cnt) {
    if (i0 + NL_ITEMS <= cnt) return *reinterpret_cast<const u64 *>(b.flags + i0);
    u64 w = 0;
    for (int k = 0; k < NL_ITEMS; ++k) w |= (u64)((i0 + k < cnt) ? b.flags[i0 + k] : (u8)J_NONE) << (8 * k);
    return w;
}
__device__ static __forceinline__ u64 nl_count(u64 w) {
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < NL_ITEMS; ++k) {
        const u32 j = (u32)(w >> (8 * k)) & J_MASK;
        s += (j == J_NEW ? 0x100000001ull : 0ull) + (j == J_RUP ? 1ull : 0ull) + (j == J_RES ? 0x100000000ull : 0ull);   // low half: NEW events, high half: leaves
    }
    return s;
}
__global__ void __launch_bounds__(NL_THREADS) k_newleaf_sums(Dims d, Bufs b, u64 *chunksum) {
    VCB_SMEM(u64, sh);
    const u32 nchunks = (d.N + NL_CHUNK - 1) / NL_CHUNK;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const u32 i0 = c * NL_CHUNK + threadIdx.x * NL_ITEMS * NL_WORDS;
        u64 s = 0;
#pragma unroll
        for (int q = 0; q < NL_WORDS; ++q) s += nl_count(nl_load(b, i0 + q * NL_ITEMS, d.N));
        u64 total;
        prims::block_exscan(s, sh, &total);
        if (threadIdx.x == 0) chunksum[c] = total;
    }
}
__global__ void __launch_bounds__(NL_THREADS) k_newleaf_chunks(Dims d, u64 *chunksum) {
    VCB_SMEM(u64, sh);
    const u32 nchunks = (d.N + NL_CHUNK - 1) / NL_CHUNK;
    u64 carry = 0;
    for (u32 base = 0; base < nchunks; base += NL_THREADS) {
        const u32 i = base + threadIdx.x;
        const u64 v = i < nchunks ? chunksum[i] : 0;
        u64 total;
        const u64 ex = prims::block_exscan(v, sh, &total);
        if (i < nchunks) chunksum[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(NL_THREADS, 6) k_newleaf_apply(Dims d, Bufs b, const u64 *chunkpref) {
    VCB_SMEM(u64, sh);
    const u32 nchunks = (d.N + NL_CHUNK - 1) / NL_CHUNK;
    for (u32 c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const u32 i0 = c * NL_CHUNK + threadIdx.x * NL_ITEMS * NL_WORDS;
        u64 w[NL_WORDS], s = 0;
#pragma unroll
        for (int q = 0; q < NL_WORDS; ++q) { w[q] = nl_load(b, i0 + q * NL_ITEMS, d.N); s += nl_count(w[q]); }
        u64 total;
        const u64 ex = prims::block_exscan(s, sh, &total) + chunkpref[c];
        u32 xn = (u32)ex, xl = (u32)(ex >> 32);
        // image starts inside this thread's 32 pixels (at most a few per launch: the division runs once per thread)
        u32 img_next = i0 < d.N ? (i0 + d.n - 1) / d.n : 0;        // first image starting at or after i0
        u32 next_start = img_next * d.n;
#pragma unroll
        for (int q = 0; q < NL_WORDS; ++q) {
#pragma unroll
            for (int k = 0; k < NL_ITEMS; ++k) {
                const u32 i = i0 + q * NL_ITEMS + k;
                if (i < d.N) {
                    if (i == next_start) { b.meta[img_next].leaf_begin = xn; ++img_next; next_start += d.n; }
                    const u32 j = (u32)(w[q] >> (8 * k)) & J_MASK;
                    if (j == J_NEW) { b.newlist[xn++] = i; b.leaflist[xl++] = i; }
                    else if (j == J_RUP) b.newlist[xn++] = i | TAG;
                    else if (j == J_RES) b.leaflist[xl++] = i;
                    if (i == d.N - 1) { b.ctr->kt = xn; b.ctr->lt = xl; }
                }
            }
        }
    }
}

// ordered compaction of the real leaves only (the per-leaf kernels run over this list)
struct LeafScan {
    Dims d; Bufs b;
    __device__ u32 count() const { return d.N; }
    __device__ u32 value(u32 i) const { const u32 j = b.flags[i] & J_MASK; return (j == J_NEW || j == J_RES) ? 1u : 0u; }
    __device__ void emit(u32 i, u32 ex, u32 v) const {
        if (v) b.leaflist[ex] = i;
        if (i == d.N - 1) b.ctr->lt = ex + v;
    }
};

// --------------------------------------------------------------------------------------------------------------
// K6: tournament, bottom-up over the reconstruction tree (one thread per leaf; at every event the first arriving
// child parks its (area, carrier) and leaves, the second one decides the winner and carries on).
// builder.rs:313-325: `if area(left) <= area(up) { combine(left -> up) } else { combine(up -> left) }`.
constexpr int K6_THREADS = 256;
struct TWalk { bool have; u32 child, W, car, v; };

// one level of a walk whose exchange result `other` and record head are already in registers
__device__ static __forceinline__ void tournament_step(const Bufs &b, TWalk &w, const uint4 head, const unsigned long long other) {
    if (other == ~0ull) { w.have = false; return; }       // first to arrive: the sibling will pick this up
    EvRec *e = b.ev + w.v;
    const bool up = head.z == w.child;                    // childU
    const u32 stale = head.w & 3u;                        // side of the up-left pixel's cluster (UPLEFT join at an event), 0 = none
    const u32 Wo = (u32)other, co = (u32)(other >> 32);
    const u32 WL = up ? Wo : w.W, WU = up ? w.W : Wo;
    const u32 cL = up ? co : w.car, cU = up ? w.car : co;
    u32 lwin = 0, dead = NONE;
    if (WL <= WU) {                                       // left loses: its label dies as the left side (builder.rs:313-318)
        stcg(&b.la[cL].death, w.v | TAG);
        w.car = cU;
        if (stale == 1) dead = cL;
    } else {
        stcg(&b.la[cU].death, w.v);
        w.car = cL;
        lwin = EV_LWIN;
        if (stale == 2) dead = cU;
    }
    if (dead != NONE) {
        // the pixel of this event joins its up-left cluster through a label read before the merge (builder.rs:301-305): it
        // resurrects the slot that has just died.  This pass either already runs with the pixel as a J_RES leaf (record the
        // leaf whose label it carries) or has to be repeated with it
        const u32 f = b.flags[w.v];
        b.flags[w.v] = (u8)(f | F_STALE);
        if ((f & J_MASK) == J_RES) stcg(&b.ls[w.v].pad[0], dead);
        else atomicAdd(&b.ctr->res_diff, 1u);
    }
    stcg(&e->arrive, stale | lwin);
    stcg(&e->a_wl, WL); stcg(&e->b_wu, WU);               // both child areas, for the listing pass
    w.W = WL + WU + head.y;                               // + cnt(v)
    w.child = w.v;
    w.v = head.x;                                         // parE
}
__device__ static __forceinline__ void tournament_root(const Bufs &b, TWalk &w) {
    const u32 root = uf_find(b.la, w.car);                // any leaf of the cluster; ufp is (nearly) flat after k_boruvka's epilogue
    stcg(&b.la[root].best0, w.car);                       // carrier of the component (best0 is free after K3)
    stcg(&b.la[root].best1, w.W);                         // final area
    w.have = false;
}

__global__ void __launch_bounds__(K6_THREADS) k_tournament(Dims d, Bufs b) {
    const u32 lt = ldcg(&b.ctr->lt);
    const int lane = threadIdx.x & 31;
    // Every lane walks leaf-to-root paths one level per iteration, TWO walks at a time so that the round trips of one overlap
    // those of the other; a walk that ends (first to arrive at an event, or root reached) is replaced at once by the next
    // leaf claimed from a global counter, so the warp stays full although most walks are one level long and a few are
    // thousands.  Per level: the head of the event record (parE, cnt, childU, stale side) is loaded while the 64-bit exchange
    // that parks / collects the sibling's (area, carrier) in the (carL, carU) mailbox is in flight: one round trip.
    TWalk w[2];
    w[0].have = w[1].have = false;
    w[0].v = w[1].v = NONE;
    WarpWork ww;
    while (true) {
        // 1. issue the level step of the walks in progress
        bool act[2];
        uint4 head[2];
        unsigned long long other[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            act[s] = w[s].have && w[s].v != NONE;
            if (act[s]) {
                // the whole 32-byte record in one sector: head (parE, cnt, childU, stale side) + tail (child areas, mailbox)
                const EvRec *e = b.ev + w[s].v;
                head[s] = __ldcg(reinterpret_cast<const uint4 *>(e));
                const uint4 tl = __ldcg(reinterpret_cast<const uint4 *>(e) + 1);
                other[s] = ((unsigned long long)tl.w << 32) | tl.z;
            }
        }
        // 2. walks that reached a root; empty slots claim new leaves (their loads overlap the exchanges above)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (w[s].have && !act[s]) tournament_root(b, w[s]);
            const bool was_empty = !w[s].have;
            const u32 k = ww.claim(&b.ctr->twork, lt, was_empty, lane);
            if (k != NONE) {
                const u32 leaf = b.leaflist[k];
                w[s].child = leaf | TAG; w[s].W = ldcg(&b.la[leaf].cnt); w[s].car = leaf;
                w[s].v = ldcg(&b.la[leaf].parL);
                w[s].have = true;
            }
        }
        // 3. consume the exchanges
        // a filled mailbox is final (only the sibling writes it, once) and is consumed as loaded; an empty one is exchanged
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (act[s] && other[s] == ~0ull)
                other[s] = atomicExch(reinterpret_cast<unsigned long long *>(&b.ev[w[s].v].carL), ((unsigned long long)w[s].car << 32) | w[s].W);
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (act[s]) tournament_step(b, w[s], head[s], other[s]);
        if (!__any_sync(0xffffffffu, w[0].have || w[1].have)) {
            if (ww.exhausted && ww.pos >= ww.end) break;
            continue;
        }
        if (ww.exhausted && ww.pos >= ww.end) {
            // nothing left to claim: finish the own walks in a tight loop
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                while (w[s].have) {
                    if (w[s].v == NONE) { tournament_root(b, w[s]); break; }
                    EvRec *e = b.ev + w[s].v;
                    const uint4 hd = __ldcg(reinterpret_cast<const uint4 *>(e));
                    const uint4 tl = __ldcg(reinterpret_cast<const uint4 *>(e) + 1);      // a_wl, b_wu, carL, carU (same 32-byte sector)
                    unsigned long long ot = ((unsigned long long)tl.w << 32) | tl.z;
                    // a filled mailbox is final (only the sibling writes it, once): consume it without an atomic; otherwise park
                    if (ot == ~0ull) ot = atomicExch(reinterpret_cast<unsigned long long *>(&e->carL), ((unsigned long long)w[s].car << 32) | w[s].W);
                    tournament_step(b, w[s], hd, ot);
                }
            }
            break;
        }
    }
}

// --------------------------------------------------------------------------------------------------------------
// K7: numbering.  D = leaves whose label dies as the left side no later than the next NEW event (builder.rs:315-317:
// the freed slot is handed to the next new cluster).  label = base + #{earlier leaves of the image not in D}.
struct LabelScan {
    Dims d; Bufs b;
    int res_mode;                                         // J_RES leaves exist: a later incarnation of the label may have freed the slot
    __device__ u32 count() const { return ldcg(&b.ctr->kt); }
    __device__ u32 value(u32 k) const {
        const u32 e = b.newlist[k];
        if (e & TAG) return 0u;                           // contracted NEW event: in D by construction
        if (res_mode && b.ls[e].pad[0]) return 0u;        // k_res_resolve
        const u32 death = b.la[e].death;
        if (death == NONE || !(death & TAG)) return 1u;
        const u32 kt = ldcg(&b.ctr->kt);
        const u32 img = e / d.n;
        if (k + 1 >= kt) return 0u;
        const u32 nxt = b.newlist[k + 1] & ~TAG;
        if (nxt / d.n != img) return 0u;
        return (death & ~TAG) <= nxt ? 0u : 1u;
    }
    // the first pass parks the value in labex[] (rewritten by emit) so that the second pass does not chase death again
    __device__ u32 value_first(u32 k) const { const u32 v = value(k); b.labex[k] = v; return v; }
    __device__ u32 value_again(u32 k) const { return b.labex[k]; }
    __device__ void emit(u32 k, u32 ex, u32 v) const {
        const u32 e = b.newlist[k];
        b.labex[k] = ex;
        if (!(e & TAG)) { b.la[e].label = ex; b.la[e].flabel = v; }   // raw scan value + not-in-D flag
        const u32 img = (e & ~TAG) / d.n;
        if (k == b.meta[img].leaf_begin) b.meta[img].notd_base = ex;
    }
};

// J_RES leaves (see the enum above).  A resurrected leaf r carries the label of the leaf that lost at r's own pixel; chains of
// resurrections end at the NEW leaf o that was handed the slot.  The slot is handed on to the next new cluster
// (builder.rs:315-317) when ANY incarnation of the label dies as the left side before the next NEW event after o, because
// the label is then still the most recently allocated one: that is recorded on o for the numbering scan.  If the
// incarnation that dies this way is resurrected at the very same pixel, the next NEW event overwrites a live slot
// (builder.rs:347-348) and its pixels are orphaned: err bit 2, such inputs go to the sequential replay.
__device__ static __forceinline__ u32 next_new_after(const Dims &d, const Bufs &b, u32 o) {
    const u32 kt = ldcg(&b.ctr->kt);
    u32 lo = 0, hi = kt;                                  // newlist is ascending in pixel id (contracted entries carry TAG)
    while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if ((b.newlist[mid] & ~TAG) <= o) lo = mid; else hi = mid; }
    if (lo + 1 >= kt) return NONE;
    const u32 nxt = b.newlist[lo + 1] & ~TAG;
    return nxt / d.n != o / d.n ? NONE : nxt;
}
__global__ void __launch_bounds__(256) k_res_resolve(Dims d, Bufs b) {
    const u32 lt = ldcg(&b.ctr->lt);
    const u32 stride = gridDim.x * 256;
    for (u32 k = blockIdx.x * 256 + threadIdx.x; k < lt; k += stride) {
        const u32 r = b.leaflist[k];
        const u32 f = b.flags[r];
        if ((f & J_MASK) != J_RES) continue;
        atomicAdd(&b.ctr->nres, 1u);
        if (!(f & F_STALE)) { atomicAdd(&b.ctr->res_diff, 1u); b.la[r].label = NONE; continue; }   // not resurrected after all
        const u32 src = ldcg(&b.ls[r].pad[0]);
        u32 o = src;
        while (o != NONE && (b.flags[o] & J_MASK) == J_RES) o = ldcg(&b.ls[o].pad[0]);
        b.la[r].label = o;                                // origin; replaced by its label after the numbering scan
        if (o == NONE) continue;
        const u32 nxt = next_new_after(d, b, o);
        const u32 death = b.la[r].death;
        if (death != NONE && (death & TAG) && (death & ~TAG) <= nxt) stcg(&b.ls[o].pad[0], 1u);
        if (b.la[src].death == (r | TAG) && r <= nxt) atomicOr(&b.ctr->err, 4u);
    }
}
__global__ void __launch_bounds__(256) k_res_labels(Bufs b) {
    const u32 lt = ldcg(&b.ctr->lt);
    const u32 stride = gridDim.x * 256;
    for (u32 k = blockIdx.x * 256 + threadIdx.x; k < lt; k += stride) {
        const u32 r = b.leaflist[k];
        if ((b.flags[r] & J_MASK) != J_RES) continue;
        const u32 o = b.la[r].label;
        b.la[r].label = o == NONE ? 0u : b.la[o].label;
    }
}

// per image: slot count: base + #(leaves not in D) + [last leaf in D]
__global__ void k_image_meta(Dims d, Bufs b, u32 base) {
    const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= d.nimg) return;
    const u32 kt = b.ctr->kt;
    const u32 lb = b.meta[img].leaf_begin;
    const u32 le = img + 1 < d.nimg ? b.meta[img + 1].leaf_begin : kt;
    u32 slots = base, maxlabel = 0;
    if (le > lb) {
        const u32 last = b.newlist[le - 1];
        const u32 last_ex = b.labex[le - 1], last_notd = (last & TAG) ? 0u : b.la[last].flabel;
        slots = base + (last_ex + last_notd - b.meta[img].notd_base) + (last_notd ? 0u : 1u);
        maxlabel = base + last_ex - b.meta[img].notd_base;
    }
    b.meta[img].slots = slots;
    b.meta[img].maxlabel = maxlabel;
}

constexpr int K8_THREADS = 256;
// per leaf: image-relative label of the carrier of its final cluster (la[].label holds the raw scan value of LabelScan)
__global__ void __launch_bounds__(K8_THREADS) k_leaf_final(Dims d, Bufs b, u32 base) {
    const u32 lt = ldcg(&b.ctr->lt);
    const u32 stride = gridDim.x * K8_THREADS;
    for (u32 k = blockIdx.x * K8_THREADS + threadIdx.x; k < lt; k += stride) {
        const u32 p = b.leaflist[k];
        const u32 root = uf_find(b.la, p);
        const u32 car = ldcg(&b.la[root].best0);
        b.la[p].flabel = base + ldcg(&b.la[car].label) - b.meta[p / d.n].notd_base;
    }
}
// per pixel: cluster_indices (builder.rs:161; label 0 for keyed pixels, builder.rs:330-334)
__global__ void __launch_bounds__(K8_THREADS) k_labels(Dims d, Bufs b, u32 *labels, u32 p_lo, u32 p_hi) {
    const u32 stride = gridDim.x * K8_THREADS;
    for (u32 g = p_lo + blockIdx.x * K8_THREADS + threadIdx.x; g < p_hi; g += stride) {
        const u32 leaf = b.pc[g];
        labels[g] = leaf == NONE ? 0u : b.la[leaf].flabel;
    }
}

// four pixels per thread (pixel counts and ranges that are multiples of 4): 128-bit loads / stores, one record lookup per
// run of equal leaves
__global__ void __launch_bounds__(K8_THREADS) k_labels4(Bufs b, u32 *labels, u32 grp_lo, u32 grp_hi) {
    const u32 stride = gridDim.x * K8_THREADS;
    for (u32 q = grp_lo + blockIdx.x * K8_THREADS + threadIdx.x; q < grp_hi; q += stride) {
        const uint4 leaf = __ldg(reinterpret_cast<const uint4 *>(b.pc) + q);
        uint4 o;
        o.x = leaf.x == NONE ? 0u : b.la[leaf.x].flabel;
        o.y = leaf.y == leaf.x ? o.x : (leaf.y == NONE ? 0u : b.la[leaf.y].flabel);
        o.z = leaf.z == leaf.y ? o.y : (leaf.z == NONE ? 0u : b.la[leaf.z].flabel);
        o.w = leaf.w == leaf.z ? o.z : (leaf.w == NONE ? 0u : b.la[leaf.w].flabel);
        reinterpret_cast<uint4 *>(labels)[q] = o;
    }
}

}  // namespace vcb
