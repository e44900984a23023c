// engine.cuh — host-side orchestration of the kernels: workspace arena, stage-1 pipeline, result objects.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "stage2.cuh"
#include "sequential.cuh"
#include "vmm.h"

namespace vcb {
struct BatchResult;
// One set of stage-2 workspace with its own stream, scan scratch and pinned read-back block.  Stage 2 is one persistent
// CTA per image and latency bound, so it leaves most of every SM idle: running it on a side stream lets it overlap the
// stage-1 kernels of the next device batch (and the stage 2 of the previous ones).
struct S2Set {
    rt::Stream st;
    rt::Event ev_in, ev_done;
    bool ready = false;
    size_t cap_slots = 0, cap_adj = 0, cap_img = 0, cap_scan = 0;
    u32 *fw = nullptr, *mnext = nullptr, *mtail = nullptr, *mark = nullptr, *adj_off = nullptr, *cursor = nullptr, *adj = nullptr,
        *tmin = nullptr, *bnd_off = nullptr, *bcur = nullptr, *live = nullptr, *col = nullptr, *chunksum = nullptr, *small = nullptr;
    uint4 *bnd_nb = nullptr;
    u64 *over = nullptr, *keys = nullptr;
    S2ImgState *ist = nullptr;
    void *h_pinned = nullptr;
    size_t h_pinned_cap = 0;
    BatchResult *busy = nullptr;      // result whose stage 2 is in flight on this set
};
}  // namespace vcb

struct vcb_ctx {
    int device = 0;
    int sms = 1;
    rt::Stream st;
    std::string err;
    // workspace arena (grow-only, sized by the largest batch seen)
    size_t capN = 0, capN_px = 0, cap_slots = 0;
    vcb::Bufs b{};
    u32 *chunksum = nullptr;          // scan scratch
    u32 *rs_table = nullptr;          // radix sort digit table
    u64 *sk0 = nullptr, *sk1 = nullptr;
    u32 *sv0 = nullptr, *sv1 = nullptr;
    u32 *d_small = nullptr;           // small device scratch (counts, slot_base ...)
    size_t cap_img = 0;
    void *h_pinned = nullptr;         // pinned staging for small D2H reads
    size_t h_pinned_cap = 0;
    u8 *d_stage = nullptr;            // device staging of host inputs
    size_t d_stage_cap = 0;
    int opt_materialize = 0;          // VCB_OPT_MATERIALIZE_INDICES
    u32 last_counters[8] = {};
    // host-to-host pipeline (vcb_runner_run_batch_host)
    rt::Stream st_h2d, st_d2h;
    bool pipe_ready = false;
    u8 *d_in[2] = {nullptr, nullptr};
    size_t d_in_cap = 0;
    // row-band mode: workspace arrays striped over the GPUs of the box (vmm.h)
    rt::VmmArray bw_flags, bw_pc, bw_la, bw_ls, bw_lr, bw_ev, bw_rgba;
    size_t bw_capN = 0;
    int bw_ndev = 0;
    int bw_devs[8] = {};
    u32 bw_split = 0;
    // stage-2 workspace sets: the hierarchical stage of one device batch runs on its set's own stream while the next batch's
    // stage 1 runs on `st` (see S2Set)
    static constexpr int MAX_S2_SETS = 8;
    vcb::S2Set *s2_sets[MAX_S2_SETS] = {};
    int s2_next = 0;
    int coop_blocks = 0;              // occupancy of the cooperative spanning-forest kernel on this device
    vcb::SeqSlot *seq_slots = nullptr; // sequential resolver (sequential.cuh): n + 1 slot records per image, allocated on first use
    size_t seq_cap = 0;
    bool s2_overlap = false;          // the current call overlaps stage 2 with the following device batches
    int live_results = 0;             // result objects that still point at this context
    bool zombie = false;              // vcb_ctx_destroy was called while results were alive: destroyed by the last vcb_clusters_free
    bool host_compact = false;        // vcb_runner_run_batch_host with live_slots: stage2_launch also compacts the records
    bool s2_tail = false;             // ... and this is its last device batch: no stage-1 kernels follow on the main stream
#ifndef VCB_HOSTSIM
    cudaEvent_t events[16] = {};
#endif
};

namespace vcb {

struct BatchResult {
    vcb_ctx *ctx = nullptr;
    Dims d{};
    bool binary = false;
    vcb_runner_config cfg{};
    u32 *labels = nullptr;            // N
    vcb_cluster_rec *rec = nullptr;   // total_slots
    u32 *outputs = nullptr;           // out_mul * total_slots (per image: [out_mul * slot_base, ... + nout))
    u32 out_mul = 1;                  // 2 after a hierarchical stage (see S2::out_mul)
    u32 *d_slot_base = nullptr;       // nimg + 1
    u32 *d_slot_img = nullptr;        // total_slots
    u32 *labels1 = nullptr;           // stage-1 labels when a hierarchical stage ran (labels then holds the final ones)
    u32 *mpar = nullptr, *outpos = nullptr;   // stage-2 merge forest / first output position per slot
    u32 *childoff = nullptr, *holeoff = nullptr, *off1 = nullptr, *hpool = nullptr, *pool1 = nullptr;
    std::vector<u32> hpool_base, pool1_base;
    bool hier = false;
    u32 *seq_nx = nullptr;            // sequential resolver: list links per pixel (0xFFFFFFFE = in no cluster's list)
    vcb_cluster_rec *crec = nullptr;  // compact records of the host pipeline (LiveScan), with their slot indices and per-image offsets
    u32 *cidx = nullptr, *coff = nullptr;
    std::vector<u32> live_off;        // nimg + 1 (host copy of coff)
    bool inert_key = false;           // sequential resolver + Discard + hierarchical stage: discarded key pixels carry INERT in labels1
    u32 *pool = nullptr;              // concatenated Cluster.indices of all images (exact order), when materialised
    u8 *pixels = nullptr;             // retained copy of the input (Clusters.pixels, container.rs:8; take_image :34-40)
    const u8 *ext_pixels = nullptr;   // device-resident runs: the caller's own buffer instead of a copy
    std::vector<u32> slot_base, nout, pool_base;
    std::vector<u64> indices_total, holes_total;
    u32 total_slots = 0;
    bool keep_key = false;
    S2Set *pending = nullptr;         // stage 2 launched on this set, not yet finished (stage2_finish)
    S2 s2args{};
    void attach(vcb_ctx *c);          // binds the result to its context (counted: see vcb_ctx_destroy)
    ~BatchResult();
};

inline int stage2_finish(vcb_ctx *c, BatchResult *res, bool sync_main = true);
inline void ctx_destroy_now(vcb_ctx *c);
inline void BatchResult::attach(vcb_ctx *c) { ctx = c; if (c) c->live_results++; }
inline BatchResult::~BatchResult() {
    if (!ctx) return;
    rt::set_device(ctx->device);
    if (pending) stage2_finish(ctx, this, true);
    for (void *p : {(void *)labels, (void *)rec, (void *)outputs, (void *)d_slot_base, (void *)d_slot_img, (void *)pool,
                    (void *)pixels, (void *)labels1, (void *)mpar, (void *)outpos, (void *)childoff, (void *)holeoff, (void *)off1,
                    (void *)hpool, (void *)pool1, (void *)seq_nx, (void *)crec, (void *)cidx, (void *)coff})
        rt::dev_free_async(ctx->st, p);
    // a context destroyed before its results lives on until the last of them is freed (include/vcb200.h)
    if (--ctx->live_results == 0 && ctx->zombie) ctx_destroy_now(ctx);
}

inline int grid_for(const vcb_ctx *c, u64 items, int threads, int per_sm = 8) {
    u64 g = (items + threads - 1) / threads;
    const u64 cap = (u64)c->sms * per_sm;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (int)g;
}

template <class T> inline bool grow(T *&p, size_t n) {
    rt::dev_free(p);
    p = (T *)rt::dev_malloc(n * sizeof(T));
    return p != nullptr;
}

// banded: the per-pixel arrays that the row-band mode stripes over the devices (flags, pc, leaf and event records) come from
// the striped allocation instead and are not allocated here
inline int ensure_workspace(vcb_ctx *c, size_t N, size_t nimg, bool banded = false) {
    if (!banded && N > c->capN_px) {
        const size_t cap = N + 64;
        bool ok = true;
        ok &= grow(c->b.flags, cap); ok &= grow(c->b.pc, cap); ok &= grow(c->b.la, cap); ok &= grow(c->b.ls, cap);
        ok &= grow(c->b.lr, cap); ok &= grow(c->b.ev, cap);
        if (!ok) { c->capN_px = 0; return VCB_ERR_NOMEM; }
        c->capN_px = N;
    }
    if (N > c->capN) {
        const size_t cap = N + 64;
        bool ok = true;
        ok &= grow(c->b.cand, cap); ok &= grow(c->b.ea, cap); ok &= grow(c->b.eb, cap); ok &= grow(c->b.newlist, cap);
        ok &= grow(c->b.attr, cap); ok &= grow(c->b.labex, cap); ok &= grow(c->b.leaflist, cap);
        if (!c->b.ctr) ok &= grow(c->b.ctr, 1);
        if (!ok) { c->capN = 0; return VCB_ERR_NOMEM; }
        c->capN = N;
    }
    // buffers indexed by slots as well as by pixels: a batch can own up to N + nimg slots (every pixel its own cluster plus the
    // ZERO slot of every image), so they follow their own capacity -- many tiny images after one large image must regrow them
    const size_t slots = N + nimg + 64;
    if (slots > c->cap_slots) {
        const size_t chunks = slots / 2048 + 4;
        bool ok = grow(c->chunksum, 256 * chunks / 2048 + chunks + 1024);
        ok &= grow(c->rs_table, 256 * chunks + 16);
        ok &= grow(c->sk0, slots); ok &= grow(c->sk1, slots);
        ok &= grow(c->sv0, slots); ok &= grow(c->sv1, slots);
        if (!ok) { c->cap_slots = 0; return VCB_ERR_NOMEM; }
        c->cap_slots = slots;
    }
    if (nimg > c->cap_img) {
        bool ok = grow(c->b.meta, nimg + 1);
        ok &= grow(c->d_small, 4 * (nimg + 4) + 32);
        if (!ok) { c->cap_img = 0; return VCB_ERR_NOMEM; }
        c->cap_img = nimg;
    }
    const size_t need = (nimg + 4) * (sizeof(ImgMeta) + sizeof(S2ImgState)) + 4096;
    if (need > c->h_pinned_cap) {
        rt::host_free(c->h_pinned);
        c->h_pinned = rt::host_malloc(need);
        c->h_pinned_cap = c->h_pinned ? need : 0;
        if (!c->h_pinned) return VCB_ERR_NOMEM;
    }
    return VCB_OK;
}

__global__ void k_meta_init(Dims d, Bufs b) {
    const u32 img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img == 0) { Counters z{}; *b.ctr = z; }
    if (img >= d.nimg) return;
    ImgMeta m{};
    m.key_rect[0] = 0x7fffffff; m.key_rect[1] = -1; m.key_rect[2] = 0x7fffffff; m.key_rect[3] = -1;
    b.meta[img] = m;
}

// Row-band mode: the devices that share one image.  Device k physically owns pixels [k*split, (k+1)*split) of every striped
// array and runs the per-pixel kernels on the rows that start inside that range; everything else runs on ctx[0].
struct DevSet {
    int n = 1;
    vcb_ctx *ctx[8] = {};
    u32 split = 0;
    // first row handled by device k (k = n gives h)
    u32 row_begin(int k, const Dims &d) const {
        if (n <= 1) return k == 0 ? 0 : d.h;
        if (k >= n) return d.h;
        const u64 r = ((u64)k * split + d.w - 1) / d.w;
        return r > d.h ? d.h : (u32)r;
    }
};
inline int sync_devices(const DevSet &ds) {
    int rc = 0;
    for (int k = 0; k < ds.n; ++k) {
        rt::set_device(ds.ctx[k]->device);
        if (rt::stream_sync(ds.ctx[k]->st) != 0 || ds.ctx[k]->st.last_error) {
            if (!rc) ds.ctx[0]->err = ds.ctx[k]->st.last_error ? ds.ctx[k]->st.last_error_text : std::string("stream synchronize failed");
            ds.ctx[k]->st.last_error = 0;
            rc = VCB_ERR_CUDA;
        }
    }
    rt::set_device(ds.ctx[0]->device);
    return rc;
}

// K1 with TMA staging: colour images whose row pitch is a multiple of 16 bytes (TMA global-stride rule).  Returns false
// when the plain-load kernel has to be used instead (binary images, odd widths, host simulator).
template <class P> inline bool launch_edges_tma(vcb_ctx *, const Dims &, const P &, const Bufs &, u32, u32) { return false; }
#ifndef VCB_HOSTSIM
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode_tiled() {
    // initialised once, thread-safe (distinct contexts may be used from distinct threads)
    static const PFN_encodeTiled fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            return (PFN_encodeTiled)p;
        cudaGetLastError();
        return (PFN_encodeTiled) nullptr;
    }();
    return fn;
}
template <> inline bool launch_edges_tma<ColorPolicy>(vcb_ctx *c, const Dims &d, const ColorPolicy &pol, const Bufs &b, u32 tile0, u32 tiles) {
    static const bool no_tma = std::getenv("VCB_NO_TMA") != nullptr;
    if (no_tma || d.w % 4 != 0 || ((uintptr_t)pol.rgba & 15) != 0) return false;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {d.w, d.h, d.nimg};
    const cuuint64_t gstride[2] = {(cuuint64_t)d.w * 4, (cuuint64_t)d.n * 4};
    const cuuint32_t box[3] = {(cuuint32_t)PW, (cuuint32_t)(TH + 1), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)pol.rgba, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    static const bool old_k1 = std::getenv("VCB_K1_TILE") != nullptr;      // the pointer-jumping tile kernel, for comparison
    if (old_k1) rt::launch(c->st, "k_edges_tile", rt::THR, k_edges_tile_tma<ColorPolicy>, dim3(tiles), dim3(K1_THREADS), K1_SMEM, d, pol, b, tmap, tile0);
    else rt::launch(c->st, "k_edges_tile", rt::THR, k_edges_rows_tma<ColorPolicy>, dim3((tiles + RW_WARPS - 1) / RW_WARPS), dim3(RW_THREADS), RW_SMEM, d, pol, b, tmap, tile0, tiles);
    return true;
}
#endif

// Runs K1..K8 for `nimg` equally sized images resident on the device and builds the per-slot records.
// base = 1 for the colour path (slot 0 = ZERO, builder.rs:44-45,189), 0 for the binary path.
template <class P>
inline int run_stage1(vcb_ctx *c, const Dims &d, const P &pol_in, u32 base, bool want_attr, int keep_key, BatchResult *res,
                      std::vector<ImgMeta> &meta_out, Counters &ctr_out, const DevSet *dsp = nullptr) {
    rt::Stream &st = c->st;
    Bufs b = c->b;
    DevSet one; one.n = 1; one.ctx[0] = c;
    const DevSet &ds = dsp ? *dsp : one;
    const bool multi = ds.n > 1;
    // colour path with diagonal = true: the pipeline is repeated until the set of resurrected pixels (J_RES, ids.cuh) is the
    // one the pass itself observes.  Inputs without such pixels (every BASELINE configuration) take exactly one pass.
    const bool res_mode = pol_in.fixpoint();
    P pol = pol_in;
    ImgMeta *hm = nullptr;
    Counters *hc = nullptr;
    int npass = 0;
    for (int pass = 0;; ++pass) {
        npass = pass + 1;
        const int g_img = (int)((d.nimg + 127) / 128);
        rt::launch(st, "k_meta_init", rt::SEQ, k_meta_init, dim3(g_img), dim3(128), 0, d, b);

        const u32 tiles_x = (d.w + TW - 1) / TW, tiles_y = (d.h + TH - 1) / TH;
        if (multi) { const int e = sync_devices(ds); if (e) return e; }
        for (int k = 0; k < ds.n; ++k) {                      // K1: device k takes the tile rows that start in its band
            vcb_ctx *ck = ds.ctx[k];
            u32 t0 = multi ? (ds.row_begin(k, d) + TH - 1) / TH : 0, t1 = multi ? (ds.row_begin(k + 1, d) + TH - 1) / TH : tiles_y * d.nimg;
            if (multi && t1 > tiles_y) t1 = tiles_y;
            if (t1 <= t0) continue;
            rt::set_device(ck->device);
            const u32 tile0 = t0 * tiles_x, tiles = (t1 - t0) * tiles_x;
            if (!launch_edges_tma(ck, d, pol, b, tile0, tiles)) {
                bool done = false;
                if constexpr (P::row_sweep) {
                    if (d.w % 4 == 0 && !std::getenv("VCB_K1_TILE")) {
                        rt::launch(ck->st, "k_edges_tile", rt::THR, k_edges_rows<P>, dim3((tiles + RW_WARPS - 1) / RW_WARPS), dim3(RW_THREADS), RW_SMEM, d, pol, b, tile0, tiles);
                        done = true;
                    }
                }
                if (!done) rt::launch(ck->st, "k_edges_tile", rt::THR, k_edges_tile<P>, dim3(tiles), dim3(K1_THREADS), K1_SMEM, d, pol, b, tile0);
            }
        }
        if (multi) { const int e = sync_devices(ds); if (e) return e; }
        for (int k = 0; k < ds.n; ++k) {                      // K2
            vcb_ctx *ck = ds.ctx[k];
            const u32 p0 = multi ? ds.row_begin(k, d) * d.w : 0, p1 = multi ? ds.row_begin(k + 1, d) * d.w : d.N;
            if (p1 <= p0) continue;
            rt::set_device(ck->device);
            if (d.w % 4 == 0)
                rt::launch(ck->st, "k_flatten_events", rt::THR, k_flatten_events4, dim3(grid_for(ck, (p1 - p0) / 4, K2_THREADS, 16)), dim3(K2_THREADS), 0, d, b, p0 / 4, p1 / 4);
            else
                rt::launch(ck->st, "k_flatten_events", rt::THR, k_flatten_events, dim3(grid_for(ck, p1 - p0, K2_THREADS, 16)), dim3(K2_THREADS), 0, d, b, p0, p1);
        }
        if (multi) { const int e = sync_devices(ds); if (e) return e; }
        {
            static const bool plain = getenv("VCB_NO_FUSED_SCAN") != nullptr;
            if (plain) {
                prims::scan(st, "scan_new", NewScan{d, b}, c->chunksum, d.N, c->sms * 8);
                prims::scan(st, "scan_new", LeafScan{d, b}, c->chunksum, d.N, c->sms * 8);
            } else {
                u64 *cs = reinterpret_cast<u64 *>(c->rs_table);           // free until the first sort
                const u32 nchunks = (d.N + NL_CHUNK - 1) / NL_CHUNK;
                const u32 grid = std::min<u32>(nchunks, (u32)c->sms * 8);
                rt::launch(st, "scan_new", rt::THR, k_newleaf_sums, dim3(grid), dim3(NL_THREADS), 128, d, b, cs);
                rt::launch(st, "scan_new", rt::THR, k_newleaf_chunks, dim3(1), dim3(NL_THREADS), 128, d, cs);
                rt::launch(st, "scan_new", rt::THR, k_newleaf_apply, dim3(grid), dim3(NL_THREADS), 128, d, b, cs);
            }
        }
        {
            if (!c->coop_blocks) c->coop_blocks = rt::max_coop_blocks((const void *)k_boruvka, K3_THREADS, 0, c->device);
            int g = c->coop_blocks;
            // a cooperative grid must be resident as a whole: while the hierarchical stage of earlier device batches occupies part of
            // every SM (S2Set), a full-occupancy grid would wait for those kernels to drain, so it is sized to fit beside them
            if (c->s2_overlap && g > c->sms * 2) g = c->sms * 2;
            const int want = (int)((d.N / 4 + K3_THREADS - 1) / K3_THREADS);
            if (g > want) g = want < 1 ? 1 : want;
            rt::launch(st, "k_boruvka", rt::COOP, k_boruvka, dim3(g), dim3(K3_THREADS), 0, b);
        }
        rt::launch(st, "k_chain_insert", rt::THR, k_chain_insert, dim3(grid_for(c, d.N / 2, K4_THREADS, 16)), dim3(K4_THREADS), 0, d, b);
        {
            // rows per segment: enough threads to fill the machine, as few restarts as possible
            const u64 wpad = (d.w + 31) & ~31u;
            const u64 target = (u64)c->sms * 2048 * (u64)ds.n;   // enough column walkers to fill every device
            u32 nseg = (u32)((target + wpad * d.nimg - 1) / (wpad * d.nimg));
            if (nseg < 1) nseg = 1;
            if (nseg > d.h) nseg = d.h;
            u32 rows = (d.h + nseg - 1) / nseg;
            static const int env_rows = getenv("VCB_ATTR_ROWS") ? atoi(getenv("VCB_ATTR_ROWS")) : 0;
            static const int env_per_sm = getenv("VCB_ATTR_PER_SM") ? atoi(getenv("VCB_ATTR_PER_SM")) : 16;
            if (env_rows > 0 && (u32)env_rows < rows) rows = (u32)env_rows;
            if (rows < 8 && d.h >= 8) rows = 8;
            const u32 nseg_t = (d.h + rows - 1) / rows;
            if (multi) { const int e = sync_devices(ds); if (e) return e; }
            for (int k = 0; k < ds.n; ++k) {                  // K5: device k takes the row segments that start in its band
                vcb_ctx *ck = ds.ctx[k];
                u32 s0 = multi ? (ds.row_begin(k, d) + rows - 1) / rows : 0, s1 = multi ? (ds.row_begin(k + 1, d) + rows - 1) / rows : nseg_t * d.nimg;
                if (multi && s1 > nseg_t) s1 = nseg_t;
                if (s1 <= s0) continue;
                rt::set_device(ck->device);
                const u64 t_lo = (u64)s0 * wpad, t_hi = (u64)s1 * wpad;
                const int g = grid_for(ck, t_hi - t_lo, K5_THREADS, env_per_sm);
                if (want_attr)
                    rt::launch(ck->st, "k_attribute", rt::THR, k_attribute<P, true>, dim3(g), dim3(K5_THREADS), 0, d, pol, b, rows, keep_key, t_lo, t_hi);
                else
                    rt::launch(ck->st, "k_attribute", rt::THR, k_attribute<P, false>, dim3(g), dim3(K5_THREADS), 0, d, pol, b, rows, keep_key, t_lo, t_hi);
            }
            if (multi) { const int e = sync_devices(ds); if (e) return e; }
        }
        rt::launch(st, "k_tournament", rt::THR, k_tournament, dim3(grid_for(c, d.N / 4, K6_THREADS, 16)), dim3(K6_THREADS), 0, d, b);
        if (res_mode) rt::launch(st, "k_res_resolve", rt::SEQ, k_res_resolve, dim3(grid_for(c, d.N / 8, 256, 8)), dim3(256), 0, d, b);
        prims::scan(st, "scan_label", LabelScan{d, b, res_mode ? 1 : 0}, c->chunksum, d.N, c->sms * 8);
        if (res_mode) rt::launch(st, "k_res_labels", rt::SEQ, k_res_labels, dim3(grid_for(c, d.N / 8, 256, 8)), dim3(256), 0, b);
        rt::launch(st, "k_image_meta", rt::SEQ, k_image_meta, dim3(g_img), dim3(128), 0, d, b, base);
        rt::launch(st, "k_leaf_final", rt::SEQ, k_leaf_final, dim3(grid_for(c, d.N / 4, K8_THREADS, 16)), dim3(K8_THREADS), 0, d, b, base);
        if (res->labels) {
            if (multi) { const int e = sync_devices(ds); if (e) return e; }
            for (int k = 0; k < ds.n; ++k) {
                vcb_ctx *ck = ds.ctx[k];
                const u32 p0 = multi ? ds.row_begin(k, d) * d.w : 0, p1 = multi ? ds.row_begin(k + 1, d) * d.w : d.N;
                if (p1 <= p0) continue;
                rt::set_device(ck->device);
                if (d.w % 4 == 0)
                    rt::launch(ck->st, "k_labels", rt::SEQ, k_labels4, dim3(grid_for(ck, (p1 - p0) / 4, K8_THREADS, 16)), dim3(K8_THREADS), 0, b, res->labels, p0 / 4, p1 / 4);
                else
                    rt::launch(ck->st, "k_labels", rt::SEQ, k_labels, dim3(grid_for(ck, p1 - p0, K8_THREADS, 16)), dim3(K8_THREADS), 0, d, b, res->labels, p0, p1);
            }
            if (multi) { const int e = sync_devices(ds); if (e) return e; }
        }

        // one small read-back: per-image slot counts (sizes of the result) and the error flags
        hm = (ImgMeta *)c->h_pinned;
        hc = (Counters *)((char *)c->h_pinned + (d.nimg + 1) * sizeof(ImgMeta));
        rt::copy_d2h(st, hm, b.meta, d.nimg * sizeof(ImgMeta));
        rt::copy_d2h(st, hc, b.ctr, sizeof(Counters));
        if (rt::stream_sync(st) != 0 || st.last_error) {
            c->err = st.last_error ? st.last_error_text : std::string("stream synchronize failed");
            st.last_error = 0;
            return VCB_ERR_CUDA;
        }
        if (!res_mode || hc->res_diff == 0) break;
        if (pass >= 200) { c->err = "stale-upleft fixpoint did not converge"; return VCB_ERR_UNSUPPORTED; }
        pol.set_use_res();
    }
    meta_out.assign(hm, hm + d.nimg);
    ctr_out = *hc;
    c->last_counters[0] = hc->ncand; c->last_counters[1] = hc->kt; c->last_counters[2] = hc->lt; c->last_counters[3] = hc->rounds;
    c->last_counters[4] = hc->err; c->last_counters[5] = hc->nres; c->last_counters[6] = (u32)npass; c->last_counters[7] = 0;
    return VCB_OK;
}

// Sequential resolver (sequential.cuh) for a device batch whose parallel pass reported an order-dependent quirk: replays
// builder.rs:273-364 literally, one image per CTA.  Fills res->labels and the slot counts; the slot records stay in
// c->seq_slots for build_records / materialize_pool_seq.
inline int run_stage1_seq(vcb_ctx *c, const Dims &d, const ColorPolicy &pol, int keep_key, BatchResult *res, std::vector<ImgMeta> &meta_out,
                          bool *orphans) {
    rt::Stream &st = c->st;
    Bufs b = c->b;
    const size_t need = (size_t)d.N + d.nimg;
    if (need > c->seq_cap) {
        if (!grow(c->seq_slots, need + 16)) { c->seq_cap = 0; return VCB_ERR_NOMEM; }
        c->seq_cap = need;
    }
    res->seq_nx = (u32 *)rt::dev_malloc_async(st, (size_t)d.N * sizeof(u32));
    if (!res->seq_nx) return VCB_ERR_NOMEM;
    rt::launch(st, "k_seq_stage1", rt::SEQ, k_seq_stage1, dim3(d.nimg), dim3(SEQ_THREADS), 0, d, pol, keep_key, res->labels, res->seq_nx, c->seq_slots, b.meta, 0u);
    if (res->inert_key)   // a hierarchical stage follows: discarded key pixels leave label 0 to the members of clusters[0] (stage2.cuh INERT)
        rt::launch(st, "k_seq_inert", rt::SEQ, k_seq_inert, dim3(grid_for(c, d.N, 256, 16)), dim3(256), 0, d, pol, res->labels);
    ImgMeta *hm = (ImgMeta *)c->h_pinned;
    rt::copy_d2h(st, hm, b.meta, d.nimg * sizeof(ImgMeta));
    if (rt::stream_sync(st) != 0 || st.last_error) {
        c->err = st.last_error ? st.last_error_text : std::string("stream synchronize failed");
        st.last_error = 0;
        return VCB_ERR_CUDA;
    }
    meta_out.assign(hm, hm + d.nimg);
    *orphans = false;
    for (u32 i = 0; i < d.nimg; ++i) if (hm[i].pad[0]) *orphans = true;
    c->last_counters[7] = 1;
    return VCB_OK;
}

// records + (optionally) the hierarchical == 0 output list
inline int build_records(vcb_ctx *c, const Dims &d, const std::vector<ImgMeta> &meta, BatchResult *res, bool flat_outputs,
                         bool keep_key, bool seq = false) {
    rt::Stream &st = c->st;
    Bufs b = c->b;
    res->slot_base.assign(d.nimg + 1, 0);
    for (u32 i = 0; i < d.nimg; ++i) res->slot_base[i + 1] = res->slot_base[i] + meta[i].slots;
    const u32 total = res->slot_base[d.nimg];
    res->total_slots = total;
    res->rec = (vcb_cluster_rec *)rt::dev_malloc_async(st, (size_t)total * sizeof(vcb_cluster_rec) + 16);
    res->d_slot_base = (u32 *)rt::dev_malloc_async(st, (d.nimg + 1) * sizeof(u32));
    res->d_slot_img = (u32 *)rt::dev_malloc_async(st, (size_t)total * sizeof(u32) + 16);
    res->outputs = (u32 *)rt::dev_malloc_async(st, (size_t)total * sizeof(u32) + 16);
    if (!res->rec || !res->d_slot_base || !res->d_slot_img || !res->outputs) return VCB_ERR_NOMEM;
    rt::copy_h2d(st, res->d_slot_base, res->slot_base.data(), (d.nimg + 1) * sizeof(u32));
    const int g = grid_for(c, total, REC_THREADS, 16);
    rt::launch(st, "k_rec_init", rt::SEQ, k_rec_init, dim3(g), dim3(REC_THREADS), 0, res->rec, total);
    rt::launch(st, "k_slot_img", rt::SEQ, k_slot_img, dim3(d.nimg), dim3(REC_THREADS), 0, res->d_slot_img, res->d_slot_base, d.nimg);
    if (seq)           // the sequential resolver already holds the per-slot sums (clusters[0] included)
        rt::launch(st, "k_seq_rec_fill", rt::SEQ, k_seq_rec_fill, dim3(grid_for(c, total, 256, 16)), dim3(256), 0, d, (const SeqSlot *)c->seq_slots,
                   res->rec, (const u32 *)res->d_slot_base, 0u, d.nimg);
    else {
        rt::launch(st, "k_rec_accum", rt::SEQ, k_rec_accum, dim3(grid_for(c, d.N / 4, REC_THREADS, 16)), dim3(REC_THREADS), 0, d, b, res->rec, res->d_slot_base);
        if (keep_key)
            rt::launch(st, "k_rec_key", rt::SEQ, k_rec_key, dim3((d.nimg + 127) / 128), dim3(128), 0, d, b, res->rec, res->d_slot_base);
    }
    u32 *d_live = c->d_small;            // nimg counters
    u32 *d_total = c->d_small + d.nimg + 1;
    int key_bits = 1, img_bits = 1;
    if (flat_outputs) {
        while ((1ull << key_bits) <= (u64)d.n * 65535ull + d.n + 1) ++key_bits;
        while ((1u << img_bits) < d.nimg) ++img_bits;
        rt::fill(st, d_live, 0, (d.nimg + 2) * sizeof(u32));
        rt::copy_h2d(st, d_total, &res->total_slots, sizeof(u32));
    }
    {
        u32 *d_img_base = c->d_small + 2 * (d.nimg + 2);   // nimg entries; the counters above use the first 2 * (nimg + 2)
        prims::scan_prefixes(st, "scan_off", OffScanFinal{res->rec, total}, c->chunksum, total, c->sms * 8);
        rt::launch(st, "k_img_base", rt::THR, k_img_base, dim3((d.nimg * 32 + REC_THREADS - 1) / REC_THREADS), dim3(REC_THREADS), 0,
                   (const vcb_cluster_rec *)res->rec, (const u32 *)res->d_slot_base, d.nimg, (const u32 *)c->chunksum, d_img_base);
        const u32 nchunks = (total + prims::SCAN_CHUNK - 1) / prims::SCAN_CHUNK;
        const u32 gf = std::max<u32>(1, std::min<u32>(nchunks, (u32)c->sms * 8));
        // the sort keys of stage_1_output come out of the same pass
        rt::launch(st, "k_rec_finalize", rt::THR, k_rec_finalize, dim3(gf), dim3(REC_THREADS), RF_SMEM, res->rec, total,
                   (const u32 *)c->chunksum, (const u32 *)res->d_slot_img, (const u32 *)d_img_base, (const u32 *)res->d_slot_base,
                   key_bits, flat_outputs ? c->sk0 : (u64 *)nullptr, c->sv0, d_live);
    }

    res->nout.assign(d.nimg, 0);
    res->indices_total.assign(d.nimg, 0);
    res->holes_total.assign(d.nimg, 0);
    if (flat_outputs) {
        const bool flipped = prims::radix_sort(st, c->sk0, c->sv0, c->sk1, c->sv1, d_total, total, key_bits + (d.nimg > 1 ? img_bits : 0),
                                               c->rs_table, c->chunksum, c->sms * 8);
        rt::copy_d2d(st, res->outputs, flipped ? c->sv1 : c->sv0, (size_t)total * sizeof(u32));
        rt::copy_d2h(st, c->h_pinned, d_live, d.nimg * sizeof(u32));
        if (rt::stream_sync(st) != 0 || st.last_error) { c->err = st.last_error_text; st.last_error = 0; return VCB_ERR_CUDA; }
        for (u32 i = 0; i < d.nimg; ++i) res->nout[i] = ((u32 *)c->h_pinned)[i];
    }
    return VCB_OK;
}

// Cluster.indices / points in the reference's exact order (listing.cuh).  Requires k_attribute<.., true> to have run.
inline int materialize_pool(vcb_ctx *c, const Dims &d, BatchResult *res, bool keep_key) {
    rt::Stream &st = c->st;
    Bufs b = c->b;
    const int g = grid_for(c, d.N / 2, LST_THREADS, 16);
    rt::launch(st, "k_list_init", rt::SEQ, k_list_init, dim3(g), dim3(LST_THREADS), 0, b);
    rt::launch(st, "k_list_jump", rt::SEQ, k_list_jump, dim3(g), dim3(LST_THREADS), 0, b);
    u32 *d_pool_base = c->d_small + 2 * (d.nimg + 2);
    u32 *d_count = c->d_small + 3 * (d.nimg + 2) + 2;
    rt::launch(st, "k_pool_base", rt::SEQ, k_pool_base, dim3(1), dim3(32), 0, (const vcb_cluster_rec *)res->rec,
               (const u32 *)res->d_slot_base, d.nimg, d_pool_base);
    rt::launch(st, "k_list_keys", rt::SEQ, k_list_keys, dim3(grid_for(c, d.N, LST_THREADS, 16)), dim3(LST_THREADS), 0, d, b,
               (const vcb_cluster_rec *)res->rec, (const u32 *)res->d_slot_base, (const u32 *)d_pool_base, keep_key ? 1 : 0, c->sk0, c->sv0);
    rt::copy_h2d(st, d_count, &d.N, sizeof(u32));
    int nbits = 1;
    while (((1ull << nbits) - 1ull) <= (u64)d.N) ++nbits;
    const bool flipped = prims::radix_sort(st, c->sk0, c->sv0, c->sk1, c->sv1, d_count, d.N, nbits, c->rs_table, c->chunksum, c->sms * 8);
    rt::copy_d2h(st, c->h_pinned, d_pool_base, (d.nimg + 1) * sizeof(u32));
    if (rt::stream_sync(st) != 0 || st.last_error) { c->err = st.last_error_text; st.last_error = 0; return VCB_ERR_CUDA; }
    res->pool_base.assign((u32 *)c->h_pinned, (u32 *)c->h_pinned + d.nimg + 1);
    const u32 total = res->pool_base[d.nimg];
    res->pool = (u32 *)rt::dev_malloc_async(st, (size_t)total * sizeof(u32) + 4);
    if (!res->pool) return VCB_ERR_NOMEM;
    rt::copy_d2d(st, res->pool, flipped ? c->sv1 : c->sv0, (size_t)total * sizeof(u32));
    return VCB_OK;
}

// the same for a batch replayed by the sequential resolver: every slot's linked list, in order
inline int materialize_pool_seq(vcb_ctx *c, const Dims &d, BatchResult *res) {
    rt::Stream &st = c->st;
    u32 *d_pool_base = c->d_small + 2 * (d.nimg + 2);
    rt::launch(st, "k_pool_base", rt::SEQ, k_pool_base, dim3(1), dim3(32), 0, (const vcb_cluster_rec *)res->rec,
               (const u32 *)res->d_slot_base, d.nimg, d_pool_base);
    rt::copy_d2h(st, c->h_pinned, d_pool_base, (d.nimg + 1) * sizeof(u32));
    if (rt::stream_sync(st) != 0 || st.last_error) { c->err = st.last_error_text; st.last_error = 0; return VCB_ERR_CUDA; }
    res->pool_base.assign((u32 *)c->h_pinned, (u32 *)c->h_pinned + d.nimg + 1);
    const u32 total = res->pool_base[d.nimg];
    res->pool = (u32 *)rt::dev_malloc_async(st, (size_t)total * sizeof(u32) + 4);
    if (!res->pool) return VCB_ERR_NOMEM;
    rt::launch(st, "k_seq_pool", rt::SEQ, k_seq_pool, dim3(grid_for(c, res->total_slots, 256, 16)), dim3(256), 0, d, (const SeqSlot *)c->seq_slots,
               (const u32 *)res->seq_nx, (const vcb_cluster_rec *)res->rec, (const u32 *)res->d_slot_base, (const u32 *)d_pool_base, res->pool, 0u, d.nimg);
    return VCB_OK;
}

// Stage 2 (builder.rs:379-539) on the records / labels produced by stage 1.  See stage2.cuh.
inline int s2_set_count();
inline S2Set *acquire_s2_set(vcb_ctx *c, bool async);
inline int stage2_finish(vcb_ctx *c, BatchResult *res, bool sync_main);

inline int ensure_s2_set(vcb_ctx *c, S2Set *s, const Dims &d, u32 total) {
    if (!s->ready) {
        if (rt::stream_create(s->st) != 0 || rt::event_create(s->ev_in) != 0 || rt::event_create(s->ev_done) != 0) return VCB_ERR_CUDA;
        s->ready = true;
    }
    if (total > s->cap_slots) {
        const size_t cap = (size_t)total + 64;
        bool ok = grow(s->fw, cap) & grow(s->mnext, cap) & grow(s->mtail, cap) & grow(s->mark, cap) & grow(s->tmin, cap) & grow(s->bnd_off, cap + 1) &
                  grow(s->bcur, cap) & grow(s->adj_off, cap + 1) & grow(s->cursor, cap) & grow(s->over, cap) & grow(s->live, cap) & grow(s->keys, cap) & grow(s->col, cap);
        if (!ok) { s->cap_slots = 0; return VCB_ERR_NOMEM; }
        s->cap_slots = total;
    }
    if ((size_t)d.N * 4 > s->cap_adj) {
        if (!grow(s->adj, (size_t)d.N * 4 + 64) || !grow(s->bnd_nb, (size_t)d.N + 64)) { s->cap_adj = 0; return VCB_ERR_NOMEM; }
        s->cap_adj = (size_t)d.N * 4;
    }
    if (d.nimg > s->cap_img) {
        if (!grow(s->ist, (size_t)d.nimg + 1) || !grow(s->small, 4 * ((size_t)d.nimg + 4) + 32)) { s->cap_img = 0; return VCB_ERR_NOMEM; }
        s->cap_img = d.nimg;
    }
    const size_t scan_items = std::max<size_t>(total, 1);
    if (scan_items > s->cap_scan) {
        if (!grow(s->chunksum, scan_items / 2048 + 1032)) { s->cap_scan = 0; return VCB_ERR_NOMEM; }
        s->cap_scan = scan_items;
    }
    const size_t need = ((size_t)d.nimg + 4) * (sizeof(S2ImgState) + 16) + 4096;
    if (need > s->h_pinned_cap) {
        rt::host_free(s->h_pinned);
        s->h_pinned = rt::host_malloc(need);
        s->h_pinned_cap = s->h_pinned ? need : 0;
        if (!s->h_pinned) return VCB_ERR_NOMEM;
    }
    return VCB_OK;
}

inline void free_s2_set(S2Set *s) {
    if (!s) return;
    for (void *p : {(void *)s->fw, (void *)s->mnext, (void *)s->mtail, (void *)s->mark, (void *)s->adj_off, (void *)s->cursor, (void *)s->adj,
                    (void *)s->tmin, (void *)s->bnd_off, (void *)s->bcur, (void *)s->bnd_nb, (void *)s->live, (void *)s->col, (void *)s->chunksum,
                    (void *)s->small, (void *)s->over, (void *)s->keys, (void *)s->ist})
        rt::dev_free(p);
    rt::host_free(s->h_pinned);
    if (s->ready) { rt::stream_destroy(s->st); rt::event_destroy(s->ev_in); rt::event_destroy(s->ev_done); }
    delete s;
}

// Launches the hierarchical stage of `res` (whose stage 1 and records are queued on c->st).  With `async` the work goes to
// the stream of a free workspace set and the call returns at once; stage2_finish completes the result.
inline int stage2_launch(vcb_ctx *c, const Dims &d, const vcb_runner_config *cfg, bool has_key, BatchResult *res, bool async) {
    const u32 total = res->total_slots;
    const bool legacy = std::getenv("VCB_S2_LEGACY") != nullptr;           // per-epoch launches + global sorts (uses the ctx scratch)
    if (legacy) async = false;
    S2Set *set = acquire_s2_set(c, async);
    if (!set) return VCB_ERR_NOMEM;
    int e = ensure_s2_set(c, set, d, total);
    if (e) return e;
    res->mpar = (u32 *)rt::dev_malloc_async(c->st, (size_t)total * 4 + 4);
    res->outpos = (u32 *)rt::dev_malloc_async(c->st, (size_t)total * 4 + 4);
    res->childoff = (u32 *)rt::dev_malloc_async(c->st, (size_t)total * 4 + 4);
    res->holeoff = (u32 *)rt::dev_malloc_async(c->st, (size_t)total * 4 + 4);
    res->labels1 = res->labels;
    rt::dev_free_async(c->st, res->outputs);                  // room for two outputs per slot (S2::out_mul)
    res->outputs = (u32 *)rt::dev_malloc_async(c->st, (size_t)total * 8 + 16);
    res->out_mul = 2;
    if (!res->outputs) return VCB_ERR_NOMEM;
    res->labels = (u32 *)rt::dev_malloc_async(c->st, (size_t)d.N * 4);
    if (!res->mpar || !res->outpos || !res->labels || !res->childoff || !res->holeoff) return VCB_ERR_NOMEM;
    res->hier = true;
    // everything stage 2 reads was produced (or allocated) on c->st: the set's stream starts after it
    rt::Stream &st = async ? set->st : c->st;
    if (async) { rt::event_record(set->ev_in, c->st); rt::stream_wait(set->st, set->ev_in); }

    u32 max_slots = 1;
    for (u32 i = 0; i < d.nimg; ++i) max_slots = std::max(max_slots, res->slot_base[i + 1] - res->slot_base[i]);
    S2 a{};
    a.d = d;
    a.cfg.hierarchical = cfg->hierarchical; a.cfg.good_min = cfg->good_min_area; a.cfg.good_max = cfg->good_max_area;
    a.cfg.deepen_diff = cfg->deepen_diff; a.cfg.hollow_neighbours = cfg->hollow_neighbours;
    a.cfg.can_discard = (cfg->keying_action == VCB_KEYING_DISCARD && has_key) ? 1 : 0;
    a.cfg.hier_max = cfg->hierarchical == VCB_HIERARCHICAL_MAX;
    a.cfg.batch_rounds = std::getenv("VCB_S2_NOBATCH") ? 0 : 1;
    a.cfg.inert = res->inert_key ? 1 : 0;
    a.cfg.rect_perim = (std::getenv("VCB_S2_RECT_PERIM") && !a.cfg.inert) ? 1 : 0;
    a.cfg.timing = std::getenv("VCB_S2_STATS") ? 1 : 0;
    a.cfg.replay_loads = c->s2_overlap ? 0 : 1;
    a.cfg.prefetch = std::getenv("VCB_S2_PREFETCH") ? std::atoi(std::getenv("VCB_S2_PREFETCH")) : 0;
    a.cfg.boff_min = std::getenv("VCB_S2_BOFF") ? (u32)std::atoi(std::getenv("VCB_S2_BOFF")) : 6u;
#ifdef VCB_HOSTSIM
    a.cfg.help_min_px = 12;
#else
    a.cfg.help_min_px = std::getenv("VCB_S2_HELP_MIN") ? (u64)std::atoll(std::getenv("VCB_S2_HELP_MIN")) : 8192;
#endif
    a.l1 = res->labels1; a.rec = res->rec; a.slot_base = res->d_slot_base; a.slot_img = res->d_slot_img; a.total_slots = total;
    a.childoff = res->childoff; a.holeoff = res->holeoff;
    a.fw = set->fw; a.mpar = res->mpar; a.mnext = set->mnext; a.mtail = set->mtail; a.mark = set->mark; a.tmin = set->tmin; a.outpos = res->outpos;
    a.bnd_off = set->bnd_off; a.bcur = set->bcur; a.bnd_nb = set->bnd_nb; a.col = set->col;
    a.adj_off = set->adj_off; a.adj = set->adj; a.cursor = set->cursor; a.outputs = res->outputs; a.out_mul = res->out_mul; a.over = set->over;
    a.ist = set->ist; a.keys = legacy ? c->sk0 : set->keys; a.vals = c->sv0; a.live = set->live;
    u32 *d_count = set->small + 3 * (d.nimg + 2) + 3;
    a.count = d_count;
    a.idx_bits = 1; while ((1u << a.idx_bits) < max_slots) ++a.idx_bits;
    a.area_bits = 1; while ((1ull << a.area_bits) <= (u64)d.n) ++a.area_bits;
    int img_bits = 1; while ((1u << img_bits) < d.nimg) ++img_bits;

    const int gs = grid_for(c, total, 256, 16), gp = grid_for(c, d.N, 256, 16);
    rt::launch(st, "k_s2_init_img", rt::SEQ, k_s2_init_img, dim3((d.nimg + 127) / 128), dim3(128), 0, a);
    rt::launch(st, "k_s2_init", rt::SEQ, k_s2_init, dim3(gs), dim3(256), 0, a);
    // widths that are a multiple of 4 (128-bit label loads): the row-sweep form; VCB_RAG_PIXEL keeps the per-pixel kernel
    static const bool rag_pixel = std::getenv("VCB_RAG_PIXEL") != nullptr;
    const bool rag_rows = !rag_pixel && !a.cfg.inert && d.w % 4 == 0 && ((uintptr_t)a.l1 & 15) == 0 && d.n % 4 == 0;
    const u32 rag_sx = (d.w + 127) / 128, rag_sy = (d.h + RAG_ROWS - 1) / RAG_ROWS;
    const u32 rag_grid = (u32)(((u64)rag_sx * rag_sy * d.nimg + 7) / 8);
    if (rag_rows) rt::launch(st, "k_rag_count", rt::THR, k_rag_rows<false>, dim3(rag_grid), dim3(256), 0, a, rag_sx, rag_sy);
    else rt::launch(st, "k_rag_count", rt::SEQ, k_rag<false>, dim3(gp), dim3(256), 0, a);
    prims::scan(st, "scan_deg", DegScan{a}, set->chunksum, total, c->sms * 8);
    prims::scan(st, "scan_deg", BndScan{a}, set->chunksum, total, c->sms * 8);
    if (rag_rows) rt::launch(st, "k_rag_fill", rt::THR, k_rag_rows<true>, dim3(rag_grid), dim3(256), 0, a, rag_sx, rag_sy);
    else rt::launch(st, "k_rag_fill", rt::SEQ, k_rag<true>, dim3(gp), dim3(256), 0, a);
    int max_epoch = 0;
    while ((2ull << max_epoch) <= (u64)d.n) ++max_epoch;
    // few images: fat CTAs (and, in the rect-scan perimeter mode, a thread-block cluster per image whose other CTAs help with
    // the large perimeter scans); many images: one slim CTA per image, several per SM
#ifdef VCB_HOSTSIM
    // 4 warps by default (fast on the thread-per-CUDA-thread simulator); VCB_S2_THREADS selects the block sizes the GPU runs
    // (256 / 512 / 1024), so that the same passes-per-round are exercised on the CPU as well
    int s2_threads = 128;
    if (const char *ev = std::getenv("VCB_S2_THREADS")) s2_threads = std::max(32, std::min(1024, std::atoi(ev) / 32 * 32));
    const u32 G = (d.nimg == 1 && a.cfg.rect_perim) ? 2u : 1u;
#else
    int s2_threads = d.nimg * 8 <= (u32)c->sms ? 1024 : d.nimg * 2 <= (u32)c->sms ? 512 : 256;
    // a run that is one device batch with at most one image per two SMs (the per-rank share of an 8-GPU batch): nothing shares
    // the SMs with the replay, so every CTA takes a whole SM -- 32 warps evaluate a round's 32 candidates in one pass (measured,
    // 64 images: 512 / 1024 threads -> 17.8 / 16.3-16.7 ms per call; at 128 images the gain was not reproducible: 22.2 and
    // 27.8 ms against 25.4-25.9 with 256 threads, so that size keeps 256)
    if (!c->s2_overlap && d.nimg * 2 <= (u32)c->sms) s2_threads = 1024;
    // the last device batch of an overlapped call shares the SMs with nothing that follows: one fat CTA per SM (a round
    // evaluates its 32 candidates in one pass of 32 warps instead of four passes of 8) shortens the tail of the call
    static const int tail_threads = std::getenv("VCB_S2_TAIL_THREADS") ? std::atoi(std::getenv("VCB_S2_TAIL_THREADS")) : 512;   // measured: 256 -> 76.9 ms, 512 -> 72.6 ms, 1024 -> 76.1 ms per 512-image step
    if (c->s2_tail && d.nimg <= (u32)c->sms) s2_threads = std::max(s2_threads, std::max(32, std::min(1024, tail_threads / 32 * 32)));
    else if (const char *ev = std::getenv("VCB_S2_THREADS")) s2_threads = std::max(32, std::min(1024, std::atoi(ev) / 32 * 32));
    u32 G = 1;
    static const u32 maxg = std::getenv("VCB_S2_MAXG") ? (u32)std::atoi(std::getenv("VCB_S2_MAXG")) : 8u;
    while (a.cfg.rect_perim && G < maxg && d.nimg * (G * 2) * 2 <= (u32)c->sms) G *= 2;   // helpers only serve the rect-scan perimeter
#endif
    const size_t smem_fixed = ((offsetof(S2Shared, map) + (a.cfg.rect_perim ? S2_MAPBYTES : 0u) + 15u) & ~(size_t)15) + 16;
    if (!legacy) {
        // fused path: every epoch in one launch, keys sorted in shared memory (an epoch with more live roots than the shared
        // key array holds is sorted in the image's global key range instead)
        u32 key_cap = d.nimg * 2 <= (u32)c->sms ? 16384u : 2048u;
        if (const char *ev = std::getenv("VCB_S2_KEYCAP")) key_cap = (u32)std::max(1, std::atoi(ev));
        else key_cap = std::min(key_cap, (max_slots + 63u) & ~63u);          // an epoch never holds more keys than the image has slots
        a.key_cap = key_cap;
        // per-slot state in shared memory (k_s2_run): as many of (fw, col, mnext + mtail) as fit for the largest image.  While
        // other device batches overlap this kernel, part of every SM's shared memory is left to their stage-1 kernels
        size_t budget = (c->s2_overlap ? 0u : 216u) * 1024u;               // overlapped batches: measured slower (stage-1 kernels starve for shared memory)
        static const int tail_smem_kb = std::getenv("VCB_S2_TAIL_SMEM_KB") ? std::atoi(std::getenv("VCB_S2_TAIL_SMEM_KB")) : 0;
        if (c->s2_tail) budget = (size_t)std::max(0, tail_smem_kb) * 1024u;
        if (const char *ev = std::getenv("VCB_S2_SMEM_KB")) budget = (size_t)std::max(0, std::atoi(ev)) * 1024u;
        const size_t used = smem_fixed + (size_t)key_cap * 8;
        const size_t avail = budget > used ? budget - used : 0;
        const u32 cap = (max_slots + 31u) & ~31u;
        a.state_cap = cap; a.state_mask = 0;
        if (G == 1) {
            if ((size_t)cap * 16 <= avail) a.state_mask = 7;
            else if ((size_t)cap * 8 <= avail) a.state_mask = 3;
            else if ((size_t)cap * 4 <= avail) a.state_mask = 1;
            else if (avail >= 4096 * 4) { a.state_cap = (u32)(avail / 4) & ~31u; a.state_mask = 1; }   // the smaller images of the batch
        }
        const size_t state_bytes = (size_t)a.state_cap * 4 * ((a.state_mask & 1) + ((a.state_mask >> 1) & 1) + ((a.state_mask >> 2) & 1) * 2);
        const size_t smem = used + state_bytes;
        if (G > 1)
            rt::launch_cluster(st, "k_s2_epoch", k_s2_run, dim3(d.nimg * G), dim3(s2_threads), smem, G, a, max_epoch, G);
        else
            rt::launch(st, "k_s2_epoch", rt::THR, k_s2_run, dim3(d.nimg), dim3(s2_threads), smem, a, max_epoch, G);
    } else {
        for (int ep = 0; ep <= max_epoch; ++ep) {
            rt::fill(st, d_count, 0, sizeof(u32));
            rt::launch(st, "k_s2_collect", rt::THR, k_s2_collect, dim3(gs), dim3(256), 0, a, ep);
            const int nbits = a.idx_bits + ep + (d.nimg > 1 ? img_bits : 0);      // within an epoch only `ep` bits of the area vary
            const bool flipped = prims::radix_sort(st, c->sk0, c->sv0, c->sk1, c->sv1, d_count, total, nbits, c->rs_table, c->chunksum, c->sms * 8);
            if (G > 1)
                rt::launch_cluster(st, "k_s2_epoch", k_s2_epoch, dim3(d.nimg * G), dim3(s2_threads), smem_fixed, G, a,
                                   (const u64 *)(flipped ? c->sk1 : c->sk0), ep, G);
            else
                rt::launch(st, "k_s2_epoch", rt::THR, k_s2_epoch, dim3(d.nimg), dim3(s2_threads), smem_fixed, a,
                           (const u64 *)(flipped ? c->sk1 : c->sk0), ep, G);
        }
    }
    if (d.n % 4 == 0 && !a.cfg.inert) rt::launch(st, "k_s2_labels", rt::SEQ, k_s2_labels4, dim3(grid_for(c, d.N / 4, 256, 16)), dim3(256), 0, a, res->labels);
    else rt::launch(st, "k_s2_labels", rt::SEQ, k_s2_labels, dim3(gp), dim3(256), 0, a, res->labels);
    // offsets of the per-slot lists after the merges
    prims::scan(st, "scan_off", OffScan{res->rec, res->d_slot_base, res->d_slot_img, total, nullptr}, set->chunksum, total, c->sms * 8);
    rt::launch(st, "k_off_rebase", rt::SEQ, k_off_rebase, dim3(gs), dim3(REC_THREADS), 0, res->rec, res->d_slot_base, res->d_slot_img, total);
    rt::launch(st, "k_off_rebase0", rt::SEQ, k_off_rebase0, dim3((d.nimg + 127) / 128), dim3(128), 0, res->rec, res->d_slot_base, d.nimg);
    u32 *d_first = set->small;
    prims::scan(st, "scan_holes", HolesScan{res->rec, total}, set->chunksum, total, c->sms * 8);
    rt::launch(st, "k_holes_first", rt::SEQ, k_holes_first, dim3((d.nimg + 127) / 128), dim3(128), 0, (const vcb_cluster_rec *)res->rec, (const u32 *)res->d_slot_base, d.nimg, d_first);
    rt::launch(st, "k_holes_rebase", rt::SEQ, k_holes_rebase, dim3(gs), dim3(256), 0, res->rec, (const u32 *)res->d_slot_base, (const u32 *)res->d_slot_img, total, d_first);
    rt::copy_d2h(st, set->h_pinned, set->ist, d.nimg * sizeof(S2ImgState));
    if (c->host_compact && total) {
        res->crec = (vcb_cluster_rec *)rt::dev_malloc_async(c->st, (size_t)total * sizeof(vcb_cluster_rec));
        res->cidx = (u32 *)rt::dev_malloc_async(c->st, (size_t)total * 4);
        res->coff = (u32 *)rt::dev_malloc_async(c->st, ((size_t)d.nimg + 1) * 4);
        if (!res->crec || !res->cidx || !res->coff) return VCB_ERR_NOMEM;
        if (async) { rt::event_record(set->ev_in, c->st); rt::stream_wait(set->st, set->ev_in); }   // the allocations are ordered on c->st
        prims::scan(st, "scan_live", LiveScan{res->rec, total, res->d_slot_base, res->d_slot_img, d.nimg, res->crec, res->cidx, res->coff},
                    set->chunksum, total, c->sms * 8);
        rt::copy_d2h(st, (char *)set->h_pinned + (size_t)d.nimg * sizeof(S2ImgState), res->coff, ((size_t)d.nimg + 1) * 4);
    }
    if (async) rt::event_record(set->ev_done, set->st);
    res->s2args = a;
    res->pending = set;
    set->busy = res;
    if (!async) return stage2_finish(c, res, true);
    return VCB_OK;
}

// Completes a result whose stage 2 was launched on a workspace set: waits for the set's stream, reads the output counts and
// (when the ordered lists were requested) scatters the indices / holes pools.  c->st is ordered after the set's work.
// sync_main = false (host pipeline, no pools requested): only the set's stream is waited for -- the main stream keeps the stage-1
// work of the following sub-batches queued and the host must not drain it.
inline int stage2_finish(vcb_ctx *c, BatchResult *res, bool sync_main) {
    S2Set *set = res->pending;
    if (!set) return VCB_OK;
    res->pending = nullptr;
    set->busy = nullptr;
    const Dims d = res->d;
    const bool on_set = set->st.s != nullptr && set->st.launches > 0 && !std::getenv("VCB_S2_LEGACY");
    rt::Stream &st = c->st;
    // the set's stream may hold the work: wait for it, then continue on c->st (ordered by the host-side wait)
    if (set->ready && rt::stream_sync(set->st) != 0) { c->err = set->st.last_error_text; set->st.last_error = 0; return VCB_ERR_CUDA; }
    if (set->st.last_error) { c->err = set->st.last_error_text; set->st.last_error = 0; return VCB_ERR_CUDA; }
    (void)on_set;
    if (sync_main || res->pool1) {
        if (rt::stream_sync(st) != 0 || st.last_error) { c->err = st.last_error_text; st.last_error = 0; return VCB_ERR_CUDA; }
    }
    S2ImgState *h = (S2ImgState *)set->h_pinned;
    for (u32 i = 0; i < d.nimg; ++i) res->nout[i] = h[i].nout;
    if (res->coff) {
        const u32 *lo = (const u32 *)((const char *)set->h_pinned + (size_t)d.nimg * sizeof(S2ImgState));
        res->live_off.assign(lo, lo + d.nimg + 1);
    }
    if (std::getenv("VCB_S2_STATS"))
        std::fprintf(stderr, "[vcb] stage 2 image 0: merges %u, outputs %u, batch rounds %u (candidates %u, committed %u), single pops %u\n",
                     h[0].nmerge, h[0].nout, h[0].st_rounds, h[0].st_cands, h[0].st_committed, h[0].st_singles);
    if (std::getenv("VCB_S2_STATS")) {
        static const char *names[16] = {"select", "nbr", "decide", "perim", "merge", "relabel", "bF", "bP", "bC", "bD", "bE", "collect", "sort", "gather", "deg", "bP2"};
        std::fprintf(stderr, "[vcb] stage 2 image 0 kcycles:");
        for (int k = 0; k < 16; ++k) std::fprintf(stderr, " %s %llu", names[k], h[0].tph[k] / 1000ull);
        std::fprintf(stderr, "\n");
    }
    if (res->pool1) {
        // ordered indices / holes pools from the stage-1 pool and the merge forest
        const S2 &a = res->s2args;
        u32 *d_ib = set->small, *d_hb = set->small + (d.nimg + 2), *d_p1b = set->small + 2 * (d.nimg + 2);
        rt::launch(st, "k_pool_base", rt::SEQ, k_pool_base, dim3(1), dim3(32), 0, (const vcb_cluster_rec *)res->rec, (const u32 *)res->d_slot_base, d.nimg, d_ib);
        rt::launch(st, "k_hpool_base", rt::SEQ, k_hpool_base, dim3(1), dim3(32), 0, (const vcb_cluster_rec *)res->rec, (const u32 *)res->d_slot_base, d.nimg, d_hb);
        rt::copy_h2d(st, d_p1b, res->pool1_base.data(), (d.nimg + 1) * sizeof(u32));
        u32 *hb = (u32 *)c->h_pinned;
        rt::copy_d2h(st, hb, d_ib, (d.nimg + 1) * sizeof(u32));
        rt::copy_d2h(st, hb + d.nimg + 1, d_hb, (d.nimg + 1) * sizeof(u32));
        if (rt::stream_sync(st) != 0 || st.last_error) { c->err = st.last_error_text; st.last_error = 0; return VCB_ERR_CUDA; }
        res->pool_base.assign(hb, hb + d.nimg + 1);
        res->hpool_base.assign(hb + d.nimg + 1, hb + 2 * (d.nimg + 1));
        res->pool = (u32 *)rt::dev_malloc_async(st, (size_t)res->pool_base[d.nimg] * 4 + 4);
        res->hpool = (u32 *)rt::dev_malloc_async(st, (size_t)res->hpool_base[d.nimg] * 4 + 4);
        if (!res->pool || !res->hpool) return VCB_ERR_NOMEM;
        rt::launch(st, "k_s2_pool_scatter", rt::SEQ, k_s2_pool_scatter, dim3(grid_for(c, res->pool1_base[d.nimg], 256, 16)), dim3(256), 0, a,
                   (const u32 *)res->pool1, (const u32 *)d_p1b, (const u32 *)res->off1, (const u32 *)d_ib, (const u32 *)d_hb, res->pool, res->hpool);
        if (rt::stream_sync(st) != 0 || st.last_error) { c->err = st.last_error_text; st.last_error = 0; return VCB_ERR_CUDA; }
    }
    return VCB_OK;
}

// next workspace set in round-robin order; a set still serving an earlier result is completed first
inline int s2_set_count() {
    static const int n = std::max(1, std::min((int)vcb_ctx::MAX_S2_SETS, std::getenv("VCB_S2_SETS") ? std::atoi(std::getenv("VCB_S2_SETS")) : 4));
    return n;
}
inline S2Set *acquire_s2_set(vcb_ctx *c, bool async) {
    const int nsets = s2_set_count();
    int k = 0;                                                  // a run that does not overlap anything always takes set 0
    if (async) { k = c->s2_next % nsets; c->s2_next = (k + 1) % nsets; }
    if (!c->s2_sets[k]) c->s2_sets[k] = new (std::nothrow) S2Set();
    S2Set *s = c->s2_sets[k];
    if (s && s->busy) stage2_finish(c, s->busy, true);
    return s;
}

}  // namespace vcb
