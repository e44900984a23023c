// vcb200.cu — the C ABI of include/vcb200.h over the sm_100a kernels.  No CPU compute path exists in this library:
// every entry point that produces results launches kernels on the context's device and fails if there is none.
#include "binary.cuh"
#include "engine.cuh"
#include "queries.cuh"

#include <algorithm>
#include <map>
#include <mutex>
#include <new>

using namespace vcb;

struct vcb_clusters {
    std::shared_ptr<BatchResult> br;
    u32 img = 0;
    // lazily fetched host copies
    std::vector<vcb_cluster_rec> h_rec;
    bool have_rec = false;
};

struct vcb_bin_clusters {
    vcb_bin_clusters_info info{};
    std::vector<vcb_bin_cluster_rec> rec;
    std::vector<int32_t> xy;
};

// vcb_clusters_components: for every output cluster, the binary clusters of its image (row f1)
struct vcb_components {
    std::vector<uint32_t> comp_off;               // num_outputs + 1
    std::vector<vcb_bin_cluster_rec> rec;         // all components, output cluster by output cluster
    std::vector<int32_t> xy;                      // their points (x, y pairs), relative to the cluster's image
};

namespace {

int fail(vcb_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}

int check_stream(vcb_ctx *c) {
    if (rt::stream_sync(c->st) != 0 || c->st.last_error) {
        c->err = c->st.last_error ? c->st.last_error_text : std::string("stream synchronize failed");
        c->st.last_error = 0;
        return VCB_ERR_CUDA;
    }
    return VCB_OK;
}

int validate_cfg(vcb_ctx *c, const vcb_runner_config *cfg) {
    if (!cfg) return fail(c, VCB_ERR_INVALID_ARG, "null config");
    // assert!(is_same_color_a < 8) (runner.rs:78); a negative shift would be an overflow panic in the reference
    if (cfg->is_same_color_a < 0 || cfg->is_same_color_a >= 8) return fail(c, VCB_ERR_INVALID_ARG, "is_same_color_a must be in 0..8");
    if (cfg->batch_size == 0) return fail(c, VCB_ERR_INVALID_ARG, "batch_size == 0 never terminates in the reference");
    return VCB_OK;
}

// the last record of the image gives the pool totals (indices_off + indices_len, holes_off + holes_len)
int fetch_last_record(vcb_clusters *h) {
    if (h->have_rec) return VCB_OK;
    BatchResult *br = h->br.get();
    vcb_ctx *c = br->ctx;
    const u32 s0 = br->slot_base[h->img], s1 = br->slot_base[h->img + 1];
    h->h_rec.resize(1);
    std::memset(&h->h_rec[0], 0, sizeof(vcb_cluster_rec));
    if (s1 > s0) {
        rt::copy_d2h(c->st, h->h_rec.data(), br->rec + s1 - 1, sizeof(vcb_cluster_rec));
        const int e = check_stream(c);
        if (e) return e;
    }
    h->have_rec = true;
    return VCB_OK;
}

// colour path on device-resident images
// An image without pixels: the reference builds `clusters = [Cluster::new()]` (builder.rs:189), labels and outputs empty.
int run_color_empty(vcb_ctx *c, const vcb_runner_config *cfg, u32 w, u32 h, size_t nimg, std::shared_ptr<BatchResult> &out_br) {
    rt::set_device(c->device);
    auto br = std::make_shared<BatchResult>();
    br->attach(c); br->cfg = *cfg;
    br->d = Dims{w, h, 0u, (u32)nimg, 0u};
    br->total_slots = (u32)nimg;
    br->rec = (vcb_cluster_rec *)rt::dev_malloc_async(c->st, nimg * sizeof(vcb_cluster_rec));
    if (!br->rec) return fail(c, VCB_ERR_NOMEM, "device out of memory");
    rt::fill(c->st, br->rec, 0, nimg * sizeof(vcb_cluster_rec));
    br->slot_base.resize(nimg + 1);
    for (size_t i = 0; i <= nimg; ++i) br->slot_base[i] = (u32)i;
    br->nout.assign(nimg, 0);
    br->pool_base.assign(nimg + 1, 0);
    br->hpool_base.assign(nimg + 1, 0);
    out_br = br;
    return check_stream(c);
}

// retain: keep a private copy of the input in the result (host inputs are staged in a buffer the next call reuses);
// otherwise the result only remembers the caller's device pointer.  async_s2: the hierarchical stage is queued on a
// workspace set's stream and the result must be completed with stage2_finish before it is handed out.
int run_color(vcb_ctx *c, const vcb_runner_config *cfg, const u8 *d_rgba, u32 w, u32 h, size_t nimg, bool want_pool,
              bool retain, std::shared_ptr<BatchResult> &out_br, const DevSet *ds = nullptr, bool async_s2 = false) {
    int e = validate_cfg(c, cfg);
    if (e) return e;
    if ((u64)w * h == 0) return run_color_empty(c, cfg, w, h, nimg, out_br);
    const u64 n64 = (u64)w * h;
    if (cfg->hierarchical != 0 && n64 * nimg >= (1ull << 30))
        return fail(c, VCB_ERR_INVALID_ARG, "hierarchical runs take at most 2^30 pixels per device batch (32-bit adjacency offsets)");
    if (n64 == 0 || n64 * nimg >= 0x7fffff00ull) return fail(c, VCB_ERR_INVALID_ARG, "width*height*n must be in 1..2^31");
    if (!d_rgba) return fail(c, VCB_ERR_INVALID_ARG, "null image");
    rt::set_device(c->device);
    Dims d{w, h, (u32)n64, (u32)nimg, (u32)(n64 * nimg)};
    e = ensure_workspace(c, d.N, d.nimg, ds != nullptr);        // row-band mode: the per-pixel arrays are the striped ones (not ours to grow)
    if (e) return fail(c, e, "device out of memory (workspace)");

    auto br = std::make_shared<BatchResult>();
    br->attach(c); br->d = d; br->cfg = *cfg;
    br->labels = (u32 *)rt::dev_malloc_async(c->st, (size_t)d.N * sizeof(u32));
    if (!br->labels) return fail(c, VCB_ERR_NOMEM, "device out of memory (labels)");
    if (retain) {
        br->pixels = (u8 *)rt::dev_malloc_async(c->st, (size_t)d.N * 4);
        if (!br->pixels) return fail(c, VCB_ERR_NOMEM, "device out of memory (image copy)");
        rt::copy_d2d(c->st, br->pixels, d_rgba, (size_t)d.N * 4);
    } else
        br->ext_pixels = d_rgba;

    ColorPolicy pol;
    pol.rgba = (const u32 *)d_rgba;
    pol.shift = cfg->is_same_color_a; pol.thres = cfg->is_same_color_b; pol.diagonal = cfg->diagonal != 0;
    pol.prepare();
    pol.key = (u32)cfg->key_color[0] | ((u32)cfg->key_color[1] << 8) | ((u32)cfg->key_color[2] << 16) | ((u32)cfg->key_color[3] << 24);
    pol.has_key = pol.key != 0;
    const bool keep = pol.has_key && cfg->keying_action == VCB_KEYING_KEEP;
    br->keep_key = keep;

    std::vector<ImgMeta> meta;
    Counters ctr{};
    e = run_stage1(c, d, pol, 1u, want_pool, keep ? 1 : 0, br.get(), meta, ctr, ds);
    // Order-dependent quirk inputs (unclean key colour; a resurrected slot that is overwritten; a stale-upleft fixpoint that
    // does not settle) are replayed by the sequential resolver on the device (sequential.cuh) instead of being rejected.
    bool seq = e == VCB_ERR_UNSUPPORTED;
    if (e && !seq) return e;
    if (!seq && (ctr.err & 5u)) seq = true;
    if (seq) {
        if (ds) return fail(c, VCB_ERR_UNSUPPORTED, "row-band mode: the input needs the sequential resolver (unclean key colour or an overwritten resurrected slot)");
        // What stays rejected: the hierarchical stage when a pixel's label and the cluster lists disagree because a resurrected
        // slot was overwritten (the orphan keeps a label whose cluster does not list it).  Stage 2 is built on labels ==
        // membership; discarded key pixels, the other such case, are given the INERT stage-1 label instead (stage2.cuh).
        if (cfg->hierarchical != 0 && ((ctr.err & 4u) || e == VCB_ERR_UNSUPPORTED))
            return fail(c, VCB_ERR_UNSUPPORTED, "hierarchical stage after a resurrected slot was overwritten (builder.rs:347-348): a pixel's label and the cluster lists disagree");
        // unclean key under KeyingAction::Discard with a hierarchical stage: the discarded pixels are taken out of label 0
        br->inert_key = cfg->hierarchical != 0 && pol.has_key && !keep;
        bool orphans = false;
        e = run_stage1_seq(c, d, pol, keep ? 1 : 0, br.get(), meta, &orphans);
        if (e) return fail(c, e, "sequential resolver failed");
        // the replay itself knows whether a resurrected slot was overwritten (with an unclean key the parallel pass's own check is
        // not reliable); the reference's hierarchical stage then divides by zero on the empty cluster the orphan's label names
        if (orphans && cfg->hierarchical != 0)
            return fail(c, VCB_ERR_UNSUPPORTED, "hierarchical stage after a resurrected slot was overwritten (builder.rs:347-348): a pixel's label and the cluster lists disagree");
    }
    e = build_records(c, d, meta, br.get(), cfg->hierarchical == 0, keep, seq);
    if (e) return e;
    if (want_pool) {
        e = seq ? materialize_pool_seq(c, d, br.get()) : materialize_pool(c, d, br.get(), keep);
        if (e) return fail(c, e, "materialising the indices pool failed");
    }
    if (cfg->hierarchical != 0) {
        if (want_pool) {
            // the stage-1 pool and its per-slot offsets feed the stage-2 pools
            br->pool1 = br->pool; br->pool = nullptr;
            br->pool1_base = br->pool_base;
            br->off1 = (u32 *)rt::dev_malloc_async(c->st, (size_t)br->total_slots * 4 + 4);
            if (!br->off1) return fail(c, VCB_ERR_NOMEM, "device out of memory");
            rt::launch(c->st, "k_save_off1", rt::SEQ, k_save_off1, dim3(grid_for(c, br->total_slots, 256, 16)), dim3(256), 0,
                       (const vcb_cluster_rec *)br->rec, br->total_slots, br->off1);
        }
        e = stage2_launch(c, d, cfg, pol.has_key != 0, br.get(), async_s2 && !ds);
        if (e) return fail(c, e, c->err.empty() ? "hierarchical stage failed" : c->err);
    }
    out_br = br;
    return VCB_OK;
}

// Device batches are processed in chunks of at most MAX_CHUNK_PX pixels: the workspace is ~140 bytes per pixel and the
// per-image stage-2 CTAs want >= 148 images in flight, so one chunk is sized to fill the SMs without exhausting HBM.
constexpr u64 MAX_CHUNK_PX = 320ull << 20;

// images per device batch of a hierarchical run over many images: small enough that the stage 2 of several batches overlaps
// the stage 1 of the following ones (S2Set), large enough to amortise the fixed cost of the ~40 launches per batch
size_t chunk_images(const vcb_runner_config *cfg, u64 n64, size_t nimg) {
    size_t per = n64 ? (size_t)(MAX_CHUNK_PX / n64) : nimg;
    if (per < 1) per = 1;
    if (cfg->hierarchical != 0) {
        if (const char *ev = std::getenv("VCB_CHUNK_IMAGES")) per = std::min(per, (size_t)std::max(1, std::atoi(ev)));
        else if (nimg > 96) per = std::min<size_t>(per, 128);
    }
    return per;
}

int run_color_device(vcb_ctx *c, const vcb_runner_config *cfg, const u8 *d_rgba, u32 w, u32 h, size_t nimg,
                     vcb_clusters **out, bool retain) {
    if (!out) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    if (nimg == 0) return VCB_OK;
    for (size_t i = 0; i < nimg; ++i) out[i] = nullptr;         // on failure every entry is null: nothing for the caller to free
    const u64 n64 = (u64)w * h;
    const size_t per_chunk = chunk_images(cfg, n64, nimg);
    std::vector<std::shared_ptr<BatchResult>> brs;
    int e = VCB_OK;
    c->s2_overlap = cfg->hierarchical != 0 && nimg > per_chunk;
    for (size_t i0 = 0; i0 < nimg && !e; i0 += per_chunk) {
        const size_t cnt = nimg - i0 < per_chunk ? nimg - i0 : per_chunk;
        std::shared_ptr<BatchResult> br;
        c->s2_tail = c->s2_overlap && i0 + per_chunk >= nimg;
        e = run_color(c, cfg, d_rgba + i0 * n64 * 4, w, h, cnt, c->opt_materialize != 0, retain, br, nullptr, nimg > per_chunk);
        c->s2_tail = false;
        if (!e) brs.push_back(br);
    }
    for (auto &br : brs) {                                      // complete the hierarchical stages still in flight
        const int e2 = stage2_finish(c, br.get());
        if (!e) e = e2;
    }
    c->s2_overlap = false;
    if (e) return e;
    size_t i0 = 0;
    for (auto &br : brs) {
        for (u32 i = 0; i < br->d.nimg; ++i) {
            vcb_clusters *hd = new (std::nothrow) vcb_clusters();
            if (!hd) { for (size_t k = 0; k < nimg; ++k) { delete out[k]; out[k] = nullptr; } return fail(c, VCB_ERR_NOMEM, "host out of memory"); }
            hd->br = br; hd->img = i;
            out[i0 + i] = hd;
        }
        i0 += br->d.nimg;
    }
    return VCB_OK;
}

// first request for Cluster.indices on a result computed without them: recompute from the retained image
int ensure_pool(BatchResult *br) {
    if (br->pool || br->d.n == 0) return VCB_OK;
    vcb_ctx *c = br->ctx;
    const u8 *px = br->pixels ? br->pixels : br->ext_pixels;
    if (!px) return fail(c, VCB_ERR_INVALID_ARG, "result holds no image to rebuild the indices from");
    std::shared_ptr<BatchResult> tmp;
    int e = run_color(c, &br->cfg, px, br->d.w, br->d.h, br->d.nimg, true, false, tmp);
    if (e) return e;
    std::swap(br->pool, tmp->pool);
    br->pool_base = tmp->pool_base;
    std::swap(br->hpool, tmp->hpool);
    br->hpool_base = tmp->hpool_base;
    return check_stream(c);
}

u8 *stage_input(vcb_ctx *c, size_t bytes) {
    if (bytes > c->d_stage_cap) {
        rt::dev_free(c->d_stage);
        c->d_stage = (u8 *)rt::dev_malloc(bytes);
        c->d_stage_cap = c->d_stage ? bytes : 0;
    }
    return c->d_stage;
}

// BinaryImage::to_clusters(diagonal) (clusters.rs:270-349) for nimg equally sized bit images stored back to back
// (words_per_img words each): the same kernels as vcb_binary_to_clusters_device in one launch sequence; per image the
// non-empty slots in slot order with their rect and their points in listing order (image-local coordinates).
struct BinImage {
    std::vector<vcb_bin_cluster_rec> rec;
    std::vector<int32_t> xy;
};
int run_binary_batch(vcb_ctx *c, const u32 *d_bits, u32 w, u32 h, u32 nimg, int diagonal, std::vector<BinImage> &out) {
    out.assign(nimg, BinImage());
    const u64 n64 = (u64)w * h;
    if (n64 == 0 || nimg == 0) return VCB_OK;
    if (n64 * nimg >= 0x7fffff00ull) return VCB_ERR_INVALID_ARG;
    Dims d{w, h, (u32)n64, nimg, (u32)(n64 * nimg)};
    int e = ensure_workspace(c, d.N, d.nimg);
    if (e) return e;
    BatchResult br;
    br.attach(c); br.d = d; br.binary = true;
    BinaryPolicy pol;
    pol.bits = d_bits; pol.n = d.n; pol.words_per_img = (d.n + 31) / 32; pol.diagonal = diagonal != 0;
    std::vector<ImgMeta> meta;
    Counters ctr{};
    e = run_stage1(c, d, pol, 0u, true, 0, &br, meta, ctr);
    if (e) return e;
    for (u32 i = 0; i < nimg; ++i)                          // panic!("overflow") (clusters.rs:322-324) in any of the images
        if (meta[i].slots > 0 && meta[i].maxlabel >= 65534u) return fail(c, VCB_ERR_LABEL_OVERFLOW, "overflow");
    e = build_records(c, d, meta, &br, false, false);
    if (!e) e = materialize_pool(c, d, &br, false);
    if (e) return e;
    const u32 total = br.total_slots, npts = br.pool_base[nimg];
    std::vector<vcb_cluster_rec> hrec(total);
    std::vector<u32> hpool(npts);
    if (total) rt::copy_d2h(c->st, hrec.data(), br.rec, (size_t)total * sizeof(vcb_cluster_rec));
    if (npts) rt::copy_d2h(c->st, hpool.data(), br.pool, (size_t)npts * 4);
    e = check_stream(c);
    if (e) return e;
    const bool pow2 = (w & (w - 1)) == 0;
    u32 lg = 0;
    while (pow2 && (1u << lg) < w) ++lg;
    for (u32 i = 0; i < nimg; ++i) {
        BinImage &bi = out[i];
        size_t npi = 0, nci = 0;
        for (u32 s = br.slot_base[i]; s < br.slot_base[i + 1]; ++s) if (hrec[s].indices_len) { npi += hrec[s].indices_len; ++nci; }
        bi.xy.resize(npi * 2);
        bi.rec.reserve(nci);
        int32_t *xy = bi.xy.data();
        size_t at = 0;
        for (u32 s = br.slot_base[i]; s < br.slot_base[i + 1]; ++s) {
            const vcb_cluster_rec &r = hrec[s];
            if (!r.indices_len) continue;
            vcb_bin_cluster_rec o{};
            for (int k = 0; k < 4; ++k) o.rect[k] = r.rect[k];
            o.points_off = at; o.points_len = r.indices_len;
            const u32 *pp = hpool.data() + br.pool_base[i] + r.indices_off;
            if (pow2) for (u32 k = 0; k < r.indices_len; ++k) { xy[2 * (at + k)] = (int32_t)(pp[k] & (w - 1)); xy[2 * (at + k) + 1] = (int32_t)(pp[k] >> lg); }
            else for (u32 k = 0; k < r.indices_len; ++k) { xy[2 * (at + k)] = (int32_t)(pp[k] % w); xy[2 * (at + k) + 1] = (int32_t)(pp[k] / w); }
            at += r.indices_len;
            bi.rec.push_back(o);
        }
    }
    return VCB_OK;
}

template <class F> void for_each_stream(vcb_ctx *c, F f) {
    f(c->st);
    for (S2Set *set : c->s2_sets) if (set && set->ready) f(set->st);
}

}  // namespace

namespace vcb {
inline void ctx_destroy_now(vcb_ctx *c) {
    rt::set_device(c->device);
    rt::stream_sync(c->st);
    Bufs &b = c->b;
    rt::dev_free(b.flags); rt::dev_free(b.pc); rt::dev_free(b.la); rt::dev_free(b.ls); rt::dev_free(b.lr); rt::dev_free(b.ev);
    rt::dev_free(b.cand); rt::dev_free(b.ea); rt::dev_free(b.eb); rt::dev_free(b.newlist); rt::dev_free(b.attr); rt::dev_free(b.labex); rt::dev_free(b.leaflist); rt::dev_free(b.meta); rt::dev_free(b.ctr);
    rt::dev_free(c->chunksum); rt::dev_free(c->rs_table); rt::dev_free(c->sk0); rt::dev_free(c->sk1);
    rt::dev_free(c->sv0); rt::dev_free(c->sv1); rt::dev_free(c->d_small); rt::dev_free(c->d_stage);
    for (S2Set *&set : c->s2_sets) { if (set && set->ready) rt::stream_sync(set->st); free_s2_set(set); set = nullptr; }
    rt::dev_free(c->d_in[0]); rt::dev_free(c->d_in[1]);
    if (c->pipe_ready) { rt::stream_destroy(c->st_h2d); rt::stream_destroy(c->st_d2h); }
    for (rt::VmmArray *a : {&c->bw_flags, &c->bw_pc, &c->bw_la, &c->bw_ls, &c->bw_lr, &c->bw_ev, &c->bw_rgba}) rt::vmm_free(*a);
    rt::host_free(c->h_pinned);
#ifndef VCB_HOSTSIM
    for (auto &ev : c->events) if (ev) cudaEventDestroy(ev);
#endif
    rt::stream_destroy(c->st);
    delete c;
}
}  // namespace vcb

extern "C" {

const char *vcb_strerror(int status) {
    switch (status) {
        case VCB_OK: return "ok";
        case VCB_ERR_INVALID_ARG: return "invalid argument";
        case VCB_ERR_LABEL_OVERFLOW: return "overflow";
        case VCB_ERR_UNSUPPORTED: return "unsupported input";
        case VCB_ERR_CUDA: return "CUDA error";
        case VCB_ERR_NCCL: return "NCCL error";
        case VCB_ERR_NOMEM: return "out of memory";
        default: return "unknown status";
    }
}

const char *vcb_last_error(const vcb_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
uint32_t vcb_abi_version(void) { return 1; }

void vcb_runner_config_default(vcb_runner_config *cfg) {   // runner.rs:23-39
    if (!cfg) return;
    cfg->diagonal = 0; cfg->hierarchical = VCB_HIERARCHICAL_MAX; cfg->batch_size = 25600;
    cfg->good_min_area = 16; cfg->good_max_area = 256 * 256;
    cfg->is_same_color_a = 4; cfg->is_same_color_b = 1; cfg->deepen_diff = 64; cfg->hollow_neighbours = 1;
    cfg->key_color[0] = cfg->key_color[1] = cfg->key_color[2] = cfg->key_color[3] = 0;
    cfg->keying_action = VCB_KEYING_KEEP;
}

int vcb_ctx_create(int device, vcb_ctx **out) {
    if (!out) return VCB_ERR_INVALID_ARG;
    *out = nullptr;
    if (device < 0 || device >= rt::device_count()) return VCB_ERR_CUDA;   // no device, no library: there is no CPU path
    if (rt::set_device(device) != 0) return VCB_ERR_CUDA;
    vcb_ctx *c = new (std::nothrow) vcb_ctx();
    if (!c) return VCB_ERR_NOMEM;
    c->device = device;
    c->sms = rt::sm_count(device);
    if (rt::stream_create(c->st) != 0) { delete c; return VCB_ERR_CUDA; }
    *out = c;
    return VCB_OK;
}

void vcb_ctx_destroy(vcb_ctx *c) {
    if (!c) return;
    // results still alive keep device buffers and the stream of this context: the context becomes a zombie and is destroyed by the
    // last vcb_clusters_free (no further call may be made on it)
    if (c->live_results > 0) { c->zombie = true; return; }
    vcb::ctx_destroy_now(c);
}

int vcb_ctx_device(const vcb_ctx *c) { return c ? c->device : -1; }

int vcb_runner_run_device(vcb_ctx *c, const vcb_runner_config *cfg, const uint8_t *d_rgba, uint32_t width,
                          uint32_t height, size_t n, vcb_clusters **out) {
    if (!c) return VCB_ERR_INVALID_ARG;
    // the caller owns d_rgba and keeps it alive while it uses take_image / the ordered lists of these handles
    return run_color_device(c, cfg, d_rgba, width, height, n, out, false);
}

int vcb_runner_run(vcb_ctx *c, const vcb_runner_config *cfg, const uint8_t *rgba, uint32_t width, uint32_t height,
                   vcb_clusters **out) {
    if (!c) return VCB_ERR_INVALID_ARG;
    if (!out) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    *out = nullptr;
    int e = validate_cfg(c, cfg);
    if (e) return e;
    const size_t bytes = (size_t)width * height * 4;
    if (bytes == 0) return run_color_device(c, cfg, nullptr, width, height, 1, out, true);
    if (!rgba) return fail(c, VCB_ERR_INVALID_ARG, "null image");
    rt::set_device(c->device);
    u8 *d = stage_input(c, bytes);
    if (!d) return fail(c, VCB_ERR_NOMEM, "device out of memory (input)");
    rt::copy_h2d(c->st, d, rgba, bytes);
    return run_color_device(c, cfg, d, width, height, 1, out, true);
}

int vcb_runner_run_batch_host(vcb_ctx *c, const vcb_runner_config *cfg, const uint8_t *const *rgba, uint32_t width,
                              uint32_t height, size_t n, vcb_host_result *results) {
    if (!c) return VCB_ERR_INVALID_ARG;
    if (!rgba || !results) return fail(c, VCB_ERR_INVALID_ARG, "null argument");
    int e = validate_cfg(c, cfg);
    if (e) return e;
    const u64 npx = (u64)width * height;
    if (npx == 0) return fail(c, VCB_ERR_INVALID_ARG, "empty image");
    if (n == 0) return VCB_OK;
    rt::set_device(c->device);
    if (!c->pipe_ready) {
        if (rt::stream_create(c->st_h2d) != 0 || rt::stream_create(c->st_d2h) != 0) return fail(c, VCB_ERR_CUDA, "stream creation failed");
        c->pipe_ready = true;
    }
    // sub-batches so that the copies overlap the compute, at most MAX_CHUNK_PX pixels each.  Without a hierarchical stage: 8 per
    // call.  With one, the download of a sub-batch can only start once its latency-bound stage 2 has finished, `depth`
    // sub-batches later: small sub-batches start the 5.5 GB of downloads early, but `depth` of them must still cover one
    // stage-2 latency -- about 40 Mpixel each (19 x 1080p) measured best (512 x 1080p: 8 / 16 / 32 / 48 / 64 sub-batches ->
    // 6.6 / 7.2 / 7.4 / 6.2 / 4.9 Gpix/s end to end).  VCB_HOST_CHUNKS overrides the count.
    u64 want_chunks = 8;
    if (cfg->hierarchical != 0) want_chunks = std::max<u64>(1, ((u64)n * npx + (40ull << 20) - 1) / (40ull << 20));
    if (const char *ev = getenv("VCB_HOST_CHUNKS")) want_chunks = std::max<u64>(1, (u64)atoll(ev));
    size_t per = (size_t)std::max<u64>(1, std::min<u64>((n + want_chunks - 1) / want_chunks, MAX_CHUNK_PX / npx));
    const size_t bytes_img = (size_t)npx * 4;
    if (per * bytes_img > c->d_in_cap) {
        rt::dev_free(c->d_in[0]); rt::dev_free(c->d_in[1]);
        c->d_in[0] = (u8 *)rt::dev_malloc(per * bytes_img);
        c->d_in[1] = (u8 *)rt::dev_malloc(per * bytes_img);
        c->d_in_cap = (c->d_in[0] && c->d_in[1]) ? per * bytes_img : 0;
        if (!c->d_in_cap) return fail(c, VCB_ERR_NOMEM, "device out of memory (input staging)");
    }
    const size_t nchunks = (n + per - 1) / per;
    std::vector<rt::Event> ev_in(nchunks), ev_done(nchunks);
    for (auto &x : ev_in) rt::event_create(x);
    for (auto &x : ev_done) rt::event_create(x);
    std::vector<std::shared_ptr<BatchResult>> keep(nchunks);
    auto upload = [&](size_t k) {
        const size_t i0 = k * per, cnt = std::min(per, n - i0);
        for (size_t i = 0; i < cnt; ++i) rt::copy_h2d(c->st_h2d, c->d_in[k & 1] + i * bytes_img, rgba[i0 + i], bytes_img);
        rt::event_record(ev_in[k], c->st_h2d);
    };
    int rc = VCB_OK;
    // results of sub-batch k go back to the host once its hierarchical stage (queued on a workspace set's stream) is complete:
    // that wait is taken after sub-batch k + 1 has been queued, so the device never idles on the host
    auto download = [&](size_t k) -> int {
        const size_t i0 = k * per, cnt = std::min(per, n - i0);
        BatchResult *br = keep[k].get();
        S2Set *set = br->pending;
        int e2 = stage2_finish(c, br, false);
        if (e2) return e2;
        if (set && set->ready) rt::stream_wait(c->st_d2h, set->ev_done);   // everything of this sub-batch ends on its set's stream
        else { rt::event_record(ev_done[k], c->st); rt::stream_wait(c->st_d2h, ev_done[k]); }
        // the last record of every image carries the list totals: one small read for the whole sub-batch
        for (size_t i = 0; i < cnt; ++i) {
            vcb_host_result &r = results[i0 + i];
            const u32 s0 = br->slot_base[i], s1 = br->slot_base[i + 1];
            r.info.width = width; r.info.height = height;
            r.info.num_slots = s1 - s0; r.info.num_outputs = br->nout[i];
            r.info.indices_total = 0; r.info.holes_total = 0;
            r.status = VCB_OK;
            if (r.cluster_indices) rt::copy_d2h(c->st_d2h, r.cluster_indices, br->labels + i * npx, (size_t)npx * 4);
            r.num_live = 0; r.compact = 0;
            if (r.records && r.live_slots && !br->live_off.empty()) {           // compact form: the slots that are not default
                const u32 l0 = br->live_off[i], nl = br->live_off[i + 1] - l0;
                r.num_live = nl; r.compact = 1;
                if (r.records_cap < nl) r.status = VCB_ERR_NOMEM;
                else if (nl) {
                    rt::copy_d2h(c->st_d2h, r.records, br->crec + l0, (size_t)nl * sizeof(vcb_cluster_rec));
                    rt::copy_d2h(c->st_d2h, r.live_slots, br->cidx + l0, (size_t)nl * 4);
                }
            } else if (r.records) {
                if (r.records_cap < s1 - s0) r.status = VCB_ERR_NOMEM;
                else rt::copy_d2h(c->st_d2h, r.records, br->rec + s0, (size_t)(s1 - s0) * sizeof(vcb_cluster_rec));
            }
            if (r.outputs) {
                if (r.outputs_cap < br->nout[i]) r.status = VCB_ERR_NOMEM;
                else if (br->nout[i]) rt::copy_d2h(c->st_d2h, r.outputs, br->outputs + (size_t)br->out_mul * s0, (size_t)br->nout[i] * 4);
            }
        }
        return VCB_OK;
    };
    upload(0);
    c->s2_overlap = cfg->hierarchical != 0 && nchunks > 1;
    c->host_compact = cfg->hierarchical != 0 && results[0].live_slots != nullptr;
    // The hierarchical stage of a sub-batch is a long latency-bound kernel on a side stream (S2Set): its results are fetched
    // `depth` sub-batches later, after that many further stage-1 passes have been queued behind it, so neither the host nor
    // the main stream ever waits for it.  Without a hierarchical stage the result is complete when run_color returns.
    size_t depth = 1;
    if (cfg->hierarchical != 0) {
        depth = (size_t)std::max(1, s2_set_count() - 1);
        if (const char *ev = getenv("VCB_HOST_DEPTH")) depth = (size_t)std::max(1, std::min(s2_set_count() - 1, atoi(ev)));
    }
    size_t next_dl = 0;
    for (size_t k = 0; k < nchunks && rc == VCB_OK; ++k) {
        const size_t cnt = std::min(per, n - k * per);
        if (k + 1 < nchunks) upload(k + 1);                     // the other input buffer: its last reader (chunk k-1) is finished
        rt::stream_wait(c->st, ev_in[k]);
        c->s2_tail = c->s2_overlap && k + 1 == nchunks;
        e = run_color(c, cfg, c->d_in[k & 1], width, height, cnt, false, false, keep[k], nullptr, nchunks > 1);
        c->s2_tail = false;
        if (e) { rc = e; break; }
        if (k >= depth) { e = download(next_dl++); if (e) { rc = e; break; } }
    }
    while (rc == VCB_OK && next_dl < nchunks) rc = download(next_dl++);
    c->s2_overlap = false;
    c->host_compact = false;
    if (rt::stream_sync(c->st_d2h) != 0 || rt::stream_sync(c->st_h2d) != 0) rc = rc ? rc : VCB_ERR_CUDA;
    if (rc == VCB_OK) {
        for (size_t i = 0; i < n; ++i) {
            vcb_host_result &r = results[i];
            // the last record carries the list totals (offsets are exclusive scans over the slots; in the compact form the slots
            // behind the last live one add nothing)
            const u32 nrec = r.compact ? r.num_live : r.info.num_slots;
            if (r.records && r.status == VCB_OK && nrec) {
                const vcb_cluster_rec &l = r.records[nrec - 1];
                r.info.indices_total = l.indices_off + l.indices_len; r.info.holes_total = l.holes_off + l.holes_len;
            }
        }
    }
    for (auto &x : ev_in) rt::event_destroy(x);
    for (auto &x : ev_done) rt::event_destroy(x);
    keep.clear();
    const int e2 = check_stream(c);
    return rc ? rc : e2;
}

int vcb_runner_run_banded(vcb_ctx *const *ctxs, int nctx, const vcb_runner_config *cfg, const uint8_t *rgba, uint32_t width,
                          uint32_t height, vcb_clusters **out) {
    if (!ctxs || nctx < 1 || nctx > 8 || !ctxs[0]) return VCB_ERR_INVALID_ARG;
    vcb_ctx *c = ctxs[0];
    if (!out) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    *out = nullptr;
    int e = validate_cfg(c, cfg);
    if (e) return e;
    const u64 n64 = (u64)width * height;
    if (n64 == 0 || !rgba) return fail(c, VCB_ERR_INVALID_ARG, "empty image");
    if (n64 >= 0x7fffff00ull) return fail(c, VCB_ERR_INVALID_ARG, "width*height must be below 2^31");
    if (nctx == 1) return vcb_runner_run(c, cfg, rgba, width, height, out);
    int devs[8];
    for (int k = 0; k < nctx; ++k) {
        if (!ctxs[k]) return fail(c, VCB_ERR_INVALID_ARG, "null context");
        devs[k] = ctxs[k]->device;
#ifndef VCB_HOSTSIM
        for (int j = 0; j < k; ++j) if (devs[j] == devs[k]) return fail(c, VCB_ERR_INVALID_ARG, "contexts must be on distinct devices");
#endif
    }
    if (!rt::vmm_available()) return fail(c, VCB_ERR_CUDA, "virtual memory management API unavailable");
    if (rt::enable_peer_access(nctx, devs) != 0) return fail(c, VCB_ERR_CUDA, "peer access between the devices is not available");
    rt::set_device(c->device);
    const u32 N = (u32)n64;
    // striped workspace: device k owns pixels [k * split, (k + 1) * split); split is a multiple of the mapping granularity
    // for 1-byte elements, hence for every element size
    // (the pages are mapped on, and accessible from, one particular set of devices: a different set needs a new mapping)
    bool same_devs = nctx == c->bw_ndev;
    for (int k = 0; same_devs && k < nctx; ++k) same_devs = c->bw_devs[k] == devs[k];
    if (N > c->bw_capN || !same_devs) {
        for (rt::VmmArray *a : {&c->bw_flags, &c->bw_pc, &c->bw_la, &c->bw_ls, &c->bw_lr, &c->bw_ev, &c->bw_rgba}) rt::vmm_free(*a);
        size_t gran = 1;
        for (int k = 0; k < nctx; ++k) gran = std::max(gran, rt::vmm_granularity(devs[k]));
        size_t split = ((size_t)N + nctx - 1) / nctx;
        split = (split + gran - 1) / gran * gran;
        int rc = 0;
        rc |= rt::vmm_alloc(c->bw_flags, split, 1, nctx, devs, N);
        rc |= rt::vmm_alloc(c->bw_pc, split, 4, nctx, devs, N);
        rc |= rt::vmm_alloc(c->bw_la, split, sizeof(LeafA), nctx, devs, N);
        rc |= rt::vmm_alloc(c->bw_ls, split, sizeof(LeafS), nctx, devs, N);
        rc |= rt::vmm_alloc(c->bw_lr, split, sizeof(LeafR), nctx, devs, N);
        rc |= rt::vmm_alloc(c->bw_ev, split, sizeof(EvRec), nctx, devs, N);
        rc |= rt::vmm_alloc(c->bw_rgba, split, 4, nctx, devs, N);
        if (rc) { c->bw_capN = 0; return fail(c, VCB_ERR_NOMEM, "striped workspace allocation failed"); }
        c->bw_capN = N; c->bw_ndev = nctx; c->bw_split = (u32)split;
        for (int k = 0; k < nctx; ++k) c->bw_devs[k] = devs[k];
    }
    e = ensure_workspace(c, N, 1, true);                        // only the arrays that are not striped
    if (e) return fail(c, e, "device out of memory (workspace)");
    DevSet ds;
    ds.n = nctx; ds.split = c->bw_split;
    for (int k = 0; k < nctx; ++k) ds.ctx[k] = ctxs[k];
    // every device loads its own band of the image from the host (no halo exchange: the row above a band is read in place
    // through the shared address range)
    for (int k = 0; k < nctx; ++k) {
        const u64 p0 = std::min<u64>((u64)k * ds.split, N), p1 = std::min<u64>((u64)(k + 1) * ds.split, N);
        if (p1 <= p0) continue;
        rt::set_device(devs[k]);
        rt::copy_h2d(ctxs[k]->st, (u8 *)c->bw_rgba.ptr + p0 * 4, rgba + p0 * 4, (size_t)(p1 - p0) * 4);
    }
    e = sync_devices(ds);
    if (e) return e;
    const Bufs saved = c->b;
    c->b.flags = (u8 *)c->bw_flags.ptr; c->b.pc = (u32 *)c->bw_pc.ptr; c->b.la = (LeafA *)c->bw_la.ptr;
    c->b.ls = (LeafS *)c->bw_ls.ptr; c->b.lr = (LeafR *)c->bw_lr.ptr; c->b.ev = (EvRec *)c->bw_ev.ptr;
    std::shared_ptr<BatchResult> br;
    e = run_color(c, cfg, (const u8 *)c->bw_rgba.ptr, width, height, 1, c->opt_materialize != 0, true, br, &ds);
    if (!e) e = check_stream(c);
    c->b = saved;
    if (e) return e;
    vcb_clusters *hd = new (std::nothrow) vcb_clusters();
    if (!hd) return fail(c, VCB_ERR_NOMEM, "host out of memory");
    hd->br = br; hd->img = 0;
    *out = hd;
    return VCB_OK;
}

int vcb_runner_run_batch(vcb_ctx *c, const vcb_runner_config *cfg, const uint8_t *const *rgba, const uint32_t *width,
                         const uint32_t *height, size_t n, vcb_clusters **out) {
    if (!c) return VCB_ERR_INVALID_ARG;
    if (!rgba || !width || !height || !out) return fail(c, VCB_ERR_INVALID_ARG, "null argument");
    int e = validate_cfg(c, cfg);
    if (e) return e;
    rt::set_device(c->device);
    for (size_t k = 0; k < n; ++k) out[k] = nullptr;
    // runs of equally sized images are processed as one device batch; on failure every handle already produced is released
    auto release_all = [&]() { for (size_t k = 0; k < n; ++k) { delete out[k]; out[k] = nullptr; } };
    size_t i = 0;
    while (i < n) {
        size_t j = i + 1;
        const size_t bytes = (size_t)width[i] * height[i] * 4;
        while (j < n && width[j] == width[i] && height[j] == height[i] && (j - i + 1) * (bytes / 4) <= MAX_CHUNK_PX) ++j;
        u8 *d = nullptr;
        if (bytes) {
            d = stage_input(c, bytes * (j - i));
            if (!d) { release_all(); return fail(c, VCB_ERR_NOMEM, "device out of memory (input)"); }
            for (size_t k = i; k < j; ++k) {
                if (!rgba[k]) { release_all(); return fail(c, VCB_ERR_INVALID_ARG, "null image"); }
                rt::copy_h2d(c->st, d + (k - i) * bytes, rgba[k], bytes);
            }
        }
        e = run_color_device(c, cfg, d, width[i], height[i], j - i, out + i, true);
        if (e) { out[i] = nullptr; release_all(); return e; }
        i = j;
    }
    return VCB_OK;
}

int vcb_clusters_info_get(const vcb_clusters *hc, vcb_clusters_info *info) {
    if (!hc || !info) return VCB_ERR_INVALID_ARG;
    vcb_clusters *h = const_cast<vcb_clusters *>(hc);
    int e = fetch_last_record(h);
    if (e) return e;
    const BatchResult *br = h->br.get();
    info->width = br->d.w; info->height = br->d.h;
    info->num_slots = br->slot_base[h->img + 1] - br->slot_base[h->img];
    info->num_outputs = br->nout[h->img];
    u64 it = 0, ht = 0;
    if (!h->h_rec.empty()) {
        const vcb_cluster_rec &l = h->h_rec.back();
        it = l.indices_off + l.indices_len; ht = l.holes_off + l.holes_len;
    }
    info->indices_total = it; info->holes_total = ht;
    return VCB_OK;
}

int vcb_clusters_cluster_indices(const vcb_clusters *h, uint32_t *out) {
    if (!h || !out) return VCB_ERR_INVALID_ARG;
    const BatchResult *br = h->br.get();
    if (br->d.n == 0) return VCB_OK;
    rt::copy_d2h(br->ctx->st, out, br->labels + (size_t)h->img * br->d.n, (size_t)br->d.n * 4);
    return check_stream(br->ctx);
}

int vcb_clusters_outputs(const vcb_clusters *h, uint32_t *out) {
    if (!h) return VCB_ERR_INVALID_ARG;
    const BatchResult *br = h->br.get();
    const u32 n = br->nout[h->img];
    if (n == 0) return VCB_OK;
    if (!out) return VCB_ERR_INVALID_ARG;
    rt::copy_d2h(br->ctx->st, out, br->outputs + (size_t)br->out_mul * br->slot_base[h->img], (size_t)n * 4);
    return check_stream(br->ctx);
}

int vcb_clusters_records(const vcb_clusters *hc, vcb_cluster_rec *out) {
    if (!hc || !out) return VCB_ERR_INVALID_ARG;
    const BatchResult *br = hc->br.get();
    const u32 s0 = br->slot_base[hc->img], s1 = br->slot_base[hc->img + 1];
    rt::copy_d2h(br->ctx->st, out, br->rec + s0, (size_t)(s1 - s0) * sizeof(vcb_cluster_rec));   // straight into the caller's buffer
    return check_stream(br->ctx);
}

int vcb_clusters_indices_pool(const vcb_clusters *h, uint32_t *out) {
    if (!h) return VCB_ERR_INVALID_ARG;
    BatchResult *br = h->br.get();
    int e = ensure_pool(br);
    if (e) return e;
    const u32 n = br->pool_base[h->img + 1] - br->pool_base[h->img];
    if (n == 0) return VCB_OK;
    if (!out) return VCB_ERR_INVALID_ARG;
    rt::copy_d2h(br->ctx->st, out, br->pool + br->pool_base[h->img], (size_t)n * 4);
    return check_stream(br->ctx);
}

int vcb_clusters_take_image(const vcb_clusters *h, uint8_t *rgba_out) {
    if (!h) return VCB_ERR_INVALID_ARG;
    if (h->br->d.n == 0) return VCB_OK;
    if (!rgba_out) return VCB_ERR_INVALID_ARG;
    const BatchResult *br = h->br.get();
    const u8 *px = br->pixels ? br->pixels : br->ext_pixels;
    if (!px) return fail(br->ctx, VCB_ERR_INVALID_ARG, "result holds no image");
    rt::copy_d2h(br->ctx->st, rgba_out, px + (size_t)h->img * br->d.n * 4, (size_t)br->d.n * 4);
    return check_stream(br->ctx);
}

int vcb_ctx_set_option(vcb_ctx *c, int option, int64_t value) {
    if (!c) return VCB_ERR_INVALID_ARG;
    if (option == VCB_OPT_MATERIALIZE_INDICES) { c->opt_materialize = value != 0; return VCB_OK; }
    return fail(c, VCB_ERR_INVALID_ARG, "unknown option");
}

int vcb_clusters_holes_pool(const vcb_clusters *h, uint32_t *out) {
    if (!h) return VCB_ERR_INVALID_ARG;
    BatchResult *br = h->br.get();
    if (!br->hier) return VCB_OK;                        // hierarchical == 0: no holes
    int e = ensure_pool(br);
    if (e) return e;
    const u32 n = br->hpool_base[h->img + 1] - br->hpool_base[h->img];
    if (n == 0) return VCB_OK;
    if (!out) return VCB_ERR_INVALID_ARG;
    rt::copy_d2h(br->ctx->st, out, br->hpool + br->hpool_base[h->img], (size_t)n * 4);
    return check_stream(br->ctx);
}

int vcb_clusters_to_color_image(const vcb_clusters *h, uint8_t *rgba_out) {
    if (!h || !rgba_out) return VCB_ERR_INVALID_ARG;
    const BatchResult *br = h->br.get();
    vcb_ctx *c = br->ctx;
    rt::set_device(c->device);
    const u32 n = br->d.n;
    if (n == 0) return VCB_OK;
    u8 *d = stage_input(c, (size_t)n * 4);
    if (!d) return fail(c, VCB_ERR_NOMEM, "device out of memory");
    if (br->hier) {
        S2 a{};
        a.d = br->d; a.l1 = br->labels1; a.rec = br->rec; a.slot_base = br->d_slot_base; a.slot_img = br->d_slot_img;
        a.total_slots = br->total_slots; a.mpar = br->mpar; a.outpos = br->outpos;
        a.cfg.inert = br->inert_key ? 1 : 0;
        u32 *slot_color = (u32 *)rt::dev_malloc_async(c->st, (size_t)br->total_slots * 4 + 4);
        if (!slot_color) return fail(c, VCB_ERR_NOMEM, "device out of memory");
        rt::launch(c->st, "k_s2_slot_color", rt::SEQ, k_s2_slot_color, dim3(grid_for(c, br->total_slots, 256, 16)), dim3(256), 0, a, slot_color);
        rt::launch(c->st, "k_s2_paint", rt::SEQ, k_s2_paint, dim3(grid_for(c, n, 256, 16)), dim3(256), 0, a, (const u32 *)slot_color, h->img, (u32 *)d);
        rt::dev_free_async(c->st, slot_color);
    } else
    rt::launch(c->st, "k_color_image_flat", rt::SEQ, k_color_image_flat, dim3(grid_for(c, n, REC_THREADS, 16)), dim3(REC_THREADS), 0,
               n, (const u32 *)(br->labels + (size_t)h->img * n), (const vcb_cluster_rec *)(br->rec + br->slot_base[h->img]), (u32 *)d,
               (const u32 *)(br->seq_nx ? br->seq_nx + (size_t)h->img * n : nullptr));
    rt::copy_d2h(c->st, rgba_out, d, (size_t)n * 4);
    return check_stream(c);
}

// ---- device-resident Cluster queries (queries.cuh)
namespace {
int query_args(const vcb_clusters *h, QArgs &q, vcb_ctx *&c) {
    if (!h || !h->br || !h->br->ctx) return VCB_ERR_INVALID_ARG;
    BatchResult *br = h->br.get();
    c = br->ctx;
    rt::set_device(c->device);
    if (br->pending) { const int e = stage2_finish(c, br); if (e) return e; }
    if (br->binary || br->d.n == 0 || !br->labels) return fail(c, VCB_ERR_INVALID_ARG, "no label map in this result");
    if (br->seq_nx) return fail(c, VCB_ERR_UNSUPPORTED, "device-resident queries on a result of the sequential resolver (labels and cluster lists may disagree): use the indices pool");
    const size_t n = br->d.n;
    q.d = Dims{br->d.w, br->d.h, br->d.n, 1u, br->d.n};
    q.fin = br->labels + (size_t)h->img * n;
    q.l1 = br->hier ? br->labels1 + (size_t)h->img * n : q.fin;
    q.mpar = br->hier ? br->mpar + br->slot_base[h->img] : nullptr;
    q.rec = br->rec + br->slot_base[h->img];
    q.nslots = br->slot_base[h->img + 1] - br->slot_base[h->img];
    return VCB_OK;
}
int query_rect(vcb_ctx *c, const vcb_clusters *h, uint32_t slot, const QArgs &q, vcb_cluster_rec &r) {
    if (slot >= q.nslots) return fail(c, VCB_ERR_INVALID_ARG, "cluster index out of range");
    rt::copy_d2h(c->st, c->h_pinned, q.rec + slot, sizeof(vcb_cluster_rec));
    const int e = check_stream(c);
    if (e) return e;
    r = *(const vcb_cluster_rec *)c->h_pinned;
    (void)h;
    return VCB_OK;
}
}  // namespace

int vcb_clusters_perimeters(const vcb_clusters *h, uint32_t *out) {
    QArgs q{}; vcb_ctx *c = nullptr;
    int e = query_args(h, q, c);
    if (e) return e;
    if (!out) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    u32 *d = (u32 *)rt::dev_malloc_async(c->st, (size_t)q.nslots * 4 + 4);
    if (!d) return fail(c, VCB_ERR_NOMEM, "device out of memory");
    rt::fill(c->st, d, 0, (size_t)q.nslots * 4);
    rt::launch(c->st, "k_q_perimeters", rt::SEQ, k_q_perimeters, dim3(grid_for(c, q.d.n, 256, 16)), dim3(256), 0, q, d);
    rt::copy_d2h(c->st, out, d, (size_t)q.nslots * 4);
    rt::dev_free_async(c->st, d);
    return check_stream(c);
}

int vcb_clusters_neighbours(const vcb_clusters *h, uint32_t slot, uint32_t *out, size_t cap, size_t *count) {
    QArgs q{}; vcb_ctx *c = nullptr;
    int e = query_args(h, q, c);
    if (e) return e;
    if (!count || (cap && !out)) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    *count = 0;
    vcb_cluster_rec r;
    e = query_rect(c, h, slot, q, r);
    if (e) return e;
    if (r.indices_len == 0) return fail(c, VCB_ERR_INVALID_ARG, "cluster holds no pixels (the reference unwraps indices.first())");
    const u32 rw = (u32)(r.rect[2] - r.rect[0]), rh = (u32)(r.rect[3] - r.rect[1]);
    u32 *flags = (u32 *)rt::dev_malloc_async(c->st, ((size_t)q.nslots + cap + 2) * 4);
    if (!flags) return fail(c, VCB_ERR_NOMEM, "device out of memory");
    u32 *list = flags + q.nslots, *cnt = list + cap;
    rt::fill(c->st, flags, 0, (size_t)q.nslots * 4);
    rt::launch(c->st, "k_q_neighbours", rt::SEQ, k_q_neighbours, dim3(grid_for(c, (u64)rw * rh, 256, 16)), dim3(256), 0, q, slot, r.rect[0], r.rect[1], rw, rh, flags);
    rt::launch(c->st, "k_q_compact", rt::THR, k_q_compact, dim3(1), dim3(256), 260 * 4, (const u32 *)flags, q.nslots, list, (u32)cap, cnt);
    u32 *hp = (u32 *)c->h_pinned;
    rt::copy_d2h(c->st, hp, cnt, 4);
    e = check_stream(c);
    if (!e) {
        *count = hp[0];
        const size_t take = std::min<size_t>(hp[0], cap);
        if (take) { rt::copy_d2h(c->st, out, list, take * 4); e = check_stream(c); }
        if (!e && hp[0] > cap) e = fail(c, VCB_ERR_NOMEM, "neighbour list larger than the caller's buffer (count holds the size needed)");
    }
    rt::dev_free_async(c->st, flags);
    return e;
}

int vcb_clusters_to_image(const vcb_clusters *h, uint32_t slot, int hole, uint32_t *bits_out, size_t cap_words, int32_t rect_out[4]) {
    QArgs q{}; vcb_ctx *c = nullptr;
    int e = query_args(h, q, c);
    if (e) return e;
    vcb_cluster_rec r;
    e = query_rect(c, h, slot, q, r);
    if (e) return e;
    if (rect_out) for (int k = 0; k < 4; ++k) rect_out[k] = r.rect[k];
    const u32 rw = (u32)(r.rect[2] - r.rect[0]), rh = (u32)(r.rect[3] - r.rect[1]);
    const size_t nwords = ((size_t)rw * rh + 31) / 32;
    if (nwords == 0) return VCB_OK;
    if (!bits_out || cap_words < nwords) return fail(c, VCB_ERR_NOMEM, "bit buffer too small: (rect width * rect height + 31) / 32 words are needed");
    u32 *d = (u32 *)rt::dev_malloc_async(c->st, nwords * 4);
    if (!d) return fail(c, VCB_ERR_NOMEM, "device out of memory");
    rt::launch(c->st, "k_q_to_image", rt::THR, k_q_to_image, dim3(grid_for(c, nwords * 32, 256, 16)), dim3(256), 0, q, slot, r.rect[0], r.rect[1], rw, rh, hole, d);
    rt::copy_d2h(c->st, bits_out, d, nwords * 4);
    rt::dev_free_async(c->st, d);
    return check_stream(c);
}

int vcb_clusters_components(const vcb_clusters *h, int hole, vcb_components **out) {
    QArgs q{}; vcb_ctx *c = nullptr;
    int e = query_args(h, q, c);
    if (e) return e;
    if (!out) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    *out = nullptr;
    BatchResult *br = h->br.get();
    const u32 s0 = br->slot_base[h->img], nout = br->nout[h->img];
    std::vector<u32> outs(nout);
    std::vector<vcb_cluster_rec> recs(q.nslots);
    if (nout) rt::copy_d2h(c->st, outs.data(), br->outputs + (size_t)br->out_mul * s0, (size_t)nout * 4);
    rt::copy_d2h(c->st, recs.data(), q.rec, (size_t)q.nslots * sizeof(vcb_cluster_rec));
    e = check_stream(c);
    if (e) return e;
    // size classes: rect width / height rounded up to a power of two (at least 32); a class is one batched run of the binary
    // pipeline over equally sized bit images (pixels beyond a cluster's rect are unset, which changes nothing for
    // to_clusters(false): an unset pixel joins nothing)
    auto cls = [](i32 v) { u32 p = 32; while ((i32)p < v) p <<= 1; return p; };
    std::map<std::pair<u32, u32>, std::vector<u32>> buckets;          // (Wp, Hp) -> positions in clusters_output
    for (u32 k = 0; k < nout; ++k) {
        const vcb_cluster_rec &r = recs[outs[k]];
        const i32 rw = r.rect[2] - r.rect[0], rh = r.rect[3] - r.rect[1];
        if (rw <= 0 || rh <= 0) continue;                             // empty image: no clusters
        buckets[{cls(rw), cls(rh)}].push_back(k);
    }
    std::vector<BinImage> per_out(nout);
    for (auto &kv : buckets) {
        const u32 Wp = kv.first.first, Hp = kv.first.second;
        const u32 wpi = (Wp * Hp + 31) / 32;
        const size_t max_items = std::max<size_t>(1, (size_t)(64ull << 20) / ((size_t)Wp * Hp));      // <= 64 Mpixel per run
        for (size_t i0 = 0; i0 < kv.second.size(); i0 += max_items) {
            const u32 cnt = (u32)std::min(max_items, kv.second.size() - i0);
            std::vector<u32> hs(cnt);
            std::vector<i32> hr((size_t)cnt * 4);
            for (u32 k = 0; k < cnt; ++k) {
                const u32 slot = outs[kv.second[i0 + k]];
                hs[k] = slot;
                for (int j = 0; j < 4; ++j) hr[4 * k + j] = recs[slot].rect[j];
            }
            u32 *d_slots = (u32 *)rt::dev_malloc_async(c->st, (size_t)cnt * 4);
            i32 *d_rects = (i32 *)rt::dev_malloc_async(c->st, (size_t)cnt * 16);
            u32 *d_bits = (u32 *)rt::dev_malloc_async(c->st, (size_t)cnt * wpi * 4);
            if (!d_slots || !d_rects || !d_bits) return fail(c, VCB_ERR_NOMEM, "device out of memory");
            rt::copy_h2d(c->st, d_slots, hs.data(), (size_t)cnt * 4);
            rt::copy_h2d(c->st, d_rects, hr.data(), (size_t)cnt * 16);
            rt::launch(c->st, "k_q_to_image_batch", rt::THR, k_q_to_image_batch, dim3(grid_for(c, (u64)cnt * wpi * 32, 256, 16)), dim3(256), 0,
                       q, (const u32 *)d_slots, (const i32 *)d_rects, cnt, Wp, Hp, wpi, hole, d_bits);
            e = check_stream(c);                                       // hs / hr are read by the copies above
            std::vector<BinImage> got;
            if (!e) e = run_binary_batch(c, d_bits, Wp, Hp, cnt, 0, got);
            rt::dev_free_async(c->st, d_slots); rt::dev_free_async(c->st, d_rects); rt::dev_free_async(c->st, d_bits);
            if (e) return e;
            for (u32 k = 0; k < cnt; ++k) per_out[kv.second[i0 + k]] = std::move(got[k]);
        }
    }
    vcb_components *res = new (std::nothrow) vcb_components();
    if (!res) return fail(c, VCB_ERR_NOMEM, "host out of memory");
    res->comp_off.assign(nout + 1, 0);
    {
        size_t nr = 0, np = 0;
        for (u32 k = 0; k < nout; ++k) { nr += per_out[k].rec.size(); np += per_out[k].xy.size(); }
        res->rec.reserve(nr); res->xy.reserve(np);
    }
    for (u32 k = 0; k < nout; ++k) {
        const uint64_t pbase = res->xy.size() / 2;
        for (vcb_bin_cluster_rec r : per_out[k].rec) { r.points_off += pbase; res->rec.push_back(r); }
        res->xy.insert(res->xy.end(), per_out[k].xy.begin(), per_out[k].xy.end());
        res->comp_off[k + 1] = (uint32_t)res->rec.size();
    }
    *out = res;
    return VCB_OK;
}
int vcb_components_info(const vcb_components *r, uint32_t *num_outputs, uint64_t *num_components, uint64_t *points_total) {
    if (!r) return VCB_ERR_INVALID_ARG;
    if (num_outputs) *num_outputs = (uint32_t)(r->comp_off.size() - 1);
    if (num_components) *num_components = r->rec.size();
    if (points_total) *points_total = r->xy.size() / 2;
    return VCB_OK;
}
int vcb_components_get(const vcb_components *r, uint32_t *comp_off, vcb_bin_cluster_rec *recs, int32_t *xy) {
    if (!r) return VCB_ERR_INVALID_ARG;
    if (comp_off) std::copy(r->comp_off.begin(), r->comp_off.end(), comp_off);
    if (recs) std::copy(r->rec.begin(), r->rec.end(), recs);
    if (xy) std::copy(r->xy.begin(), r->xy.end(), xy);
    return VCB_OK;
}
void vcb_components_free(vcb_components *r) { delete r; }

const uint32_t *vcb_clusters_device_labels(const vcb_clusters *h) {
    if (!h) return nullptr;
    return h->br->labels + (size_t)h->img * h->br->d.n;
}

void vcb_clusters_free(vcb_clusters *h) { delete h; }

int vcb_binary_to_clusters_device(vcb_ctx *c, const uint32_t *d_bits, uint32_t width, uint32_t height, int diagonal,
                                  vcb_bin_clusters **out) {
    if (!c) return VCB_ERR_INVALID_ARG;
    if (!out) return fail(c, VCB_ERR_INVALID_ARG, "null output");
    *out = nullptr;
    const u64 n64 = (u64)width * height;
    if (n64 >= 0x7fffff00ull) return fail(c, VCB_ERR_INVALID_ARG, "width*height must be below 2^31");
    vcb_bin_clusters *hb = new (std::nothrow) vcb_bin_clusters();
    if (!hb) return fail(c, VCB_ERR_NOMEM, "host out of memory");
    hb->info.width = width; hb->info.height = height;
    if (n64 == 0) { *out = hb; return VCB_OK; }                 // empty image: no clusters, default rect
    if (!d_bits) { delete hb; return fail(c, VCB_ERR_INVALID_ARG, "null image"); }
    rt::set_device(c->device);
    Dims d{width, height, (u32)n64, 1u, (u32)n64};
    int e = ensure_workspace(c, d.N, 1);
    if (e) { delete hb; return fail(c, e, "device out of memory (workspace)"); }
    BatchResult br;
    br.attach(c); br.d = d; br.binary = true;
    BinaryPolicy pol;
    pol.bits = d_bits; pol.n = d.n; pol.words_per_img = (d.n + 31) / 32; pol.diagonal = diagonal != 0;
    std::vector<ImgMeta> meta;
    Counters ctr{};
    e = run_stage1(c, d, pol, 0u, true, 0, &br, meta, ctr);
    if (e) { delete hb; return e; }
    // panic!("overflow") (clusters.rs:322-324): the running label counter reaches 65535, i.e. some new cluster gets label 65534
    if (meta[0].slots > 0 && meta[0].maxlabel >= 65534u) { delete hb; return fail(c, VCB_ERR_LABEL_OVERFLOW, "overflow"); }
    e = build_records(c, d, meta, &br, false, false);
    if (!e) e = materialize_pool(c, d, &br, false);
    if (e) { delete hb; return e; }
    const u32 total = br.total_slots, npts = br.pool_base[1];
    vcb_bin_cluster_rec *d_out = (vcb_bin_cluster_rec *)rt::dev_malloc_async(c->st, (size_t)(total + 1) * sizeof(vcb_bin_cluster_rec));
    i32 *d_xy = (i32 *)rt::dev_malloc_async(c->st, (size_t)npts * 8 + 8);
    i32 *d_overall = (i32 *)c->d_small;
    u32 *d_count = c->d_small + 4;
    if (!d_out || !d_xy) { rt::dev_free_async(c->st, d_out); rt::dev_free_async(c->st, d_xy); delete hb; return fail(c, VCB_ERR_NOMEM, "device out of memory"); }
    const i32 init[5] = {0x7fffffff, 0x7fffffff, -1, -1, 0};
    rt::copy_h2d(c->st, d_overall, init, sizeof(init));
    if (total)
        prims::scan(c->st, "scan_bin", BinCompact{br.rec, total, d_out, d_overall, d_count}, c->chunksum, total, c->sms * 8);
    rt::launch(c->st, "k_bin_points", rt::SEQ, k_bin_points, dim3(grid_for(c, npts, 256, 16)), dim3(256), 0,
               (const u32 *)br.pool, npts, width, d_xy);
    i32 h_small[5];
    rt::copy_d2h(c->st, h_small, d_overall, sizeof(h_small));
    e = check_stream(c);
    if (!e) {
        const u32 ncl = total ? (u32)h_small[4] : 0;
        hb->rec.resize(ncl);
        hb->xy.resize((size_t)npts * 2);
        rt::copy_d2h(c->st, hb->rec.data(), d_out, (size_t)ncl * sizeof(vcb_bin_cluster_rec));
        rt::copy_d2h(c->st, hb->xy.data(), d_xy, (size_t)npts * 8);
        e = check_stream(c);
        hb->info.num_clusters = ncl; hb->info.points_total = npts;
        if (ncl) { for (int k = 0; k < 4; ++k) hb->info.rect[k] = h_small[k]; }
    }
    rt::dev_free_async(c->st, d_out); rt::dev_free_async(c->st, d_xy);
    if (e) { delete hb; return e; }
    *out = hb;
    return VCB_OK;
}

int vcb_binary_to_clusters(vcb_ctx *c, const uint32_t *bits, uint32_t width, uint32_t height, int diagonal,
                           vcb_bin_clusters **out) {
    if (!c) return VCB_ERR_INVALID_ARG;
    const u64 n64 = (u64)width * height;
    if (n64 == 0) return vcb_binary_to_clusters_device(c, nullptr, width, height, diagonal, out);
    if (!bits) return fail(c, VCB_ERR_INVALID_ARG, "null image");
    rt::set_device(c->device);
    const size_t bytes = (size_t)((n64 + 31) / 32) * 4;
    u8 *d = stage_input(c, bytes);
    if (!d) return fail(c, VCB_ERR_NOMEM, "device out of memory (input)");
    rt::copy_h2d(c->st, d, bits, bytes);
    return vcb_binary_to_clusters_device(c, (const uint32_t *)d, width, height, diagonal, out);
}

int vcb_bin_clusters_info_get(const vcb_bin_clusters *h, vcb_bin_clusters_info *info) {
    if (!h || !info) return VCB_ERR_INVALID_ARG;
    *info = h->info;
    return VCB_OK;
}
int vcb_bin_clusters_records(const vcb_bin_clusters *h, vcb_bin_cluster_rec *out) {
    if (!h || (!out && !h->rec.empty())) return VCB_ERR_INVALID_ARG;
    std::copy(h->rec.begin(), h->rec.end(), out);
    return VCB_OK;
}
int vcb_bin_clusters_points(const vcb_bin_clusters *h, int32_t *xy) {
    if (!h || (!xy && !h->xy.empty())) return VCB_ERR_INVALID_ARG;
    std::copy(h->xy.begin(), h->xy.end(), xy);
    return VCB_OK;
}
void vcb_bin_clusters_free(vcb_bin_clusters *h) { delete h; }

int vcb_ctx_synchronize(vcb_ctx *c) {
    if (!c) return VCB_ERR_INVALID_ARG;
    return check_stream(c);
}

uint64_t vcb_ctx_kernel_launches(const vcb_ctx *c) {
    if (!c) return 0;
    uint64_t n = c->st.launches;
    for (const S2Set *set : c->s2_sets) if (set) n += set->st.launches;      // the hierarchical stage runs on the sets' streams
    return n;
}

int vcb_ctx_profile_reset(vcb_ctx *c, int enable) {
    if (!c) return VCB_ERR_INVALID_ARG;
    for_each_stream(c, [&](rt::Stream &st) {
        rt::stream_sync(st);
#ifndef VCB_HOSTSIM
        for (auto &p : st.prof) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
#endif
        st.prof.clear();
        st.profiling = enable != 0;
    });
    return VCB_OK;
}

int vcb_ctx_profile_read(vcb_ctx *c, const char *kernel_name, double *total_ms, uint64_t *launches) {
    if (!c || !total_ms || !launches) return VCB_ERR_INVALID_ARG;
    double t = 0; uint64_t n = 0;
    for_each_stream(c, [&](rt::Stream &st) {
        rt::stream_sync(st);
#ifndef VCB_HOSTSIM
        for (auto &p : st.prof) {
            if (kernel_name && kernel_name[0] && std::strcmp(kernel_name, p.name) != 0) continue;
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) { t += ms; ++n; }
        }
#else
        (void)kernel_name;
#endif
    });
    *total_ms = t; *launches = n;
    return VCB_OK;
}

int vcb_ctx_last_counters(const vcb_ctx *c, uint32_t out[8]) {
    if (!c || !out) return VCB_ERR_INVALID_ARG;
    for (int k = 0; k < 8; ++k) out[k] = c->last_counters[k];
    return VCB_OK;
}

int vcb_ctx_event_record(vcb_ctx *c, int slot) {
    if (!c || slot < 0 || slot >= 16) return VCB_ERR_INVALID_ARG;
#ifndef VCB_HOSTSIM
    rt::set_device(c->device);
    if (!c->events[slot] && cudaEventCreate(&c->events[slot]) != cudaSuccess) return VCB_ERR_CUDA;
    if (cudaEventRecord(c->events[slot], c->st.s) != cudaSuccess) return VCB_ERR_CUDA;
#endif
    return VCB_OK;
}

int vcb_ctx_event_elapsed_ms(vcb_ctx *c, int a, int b, double *ms) {
    if (!c || !ms || a < 0 || b < 0 || a >= 16 || b >= 16) return VCB_ERR_INVALID_ARG;
    *ms = 0;
#ifndef VCB_HOSTSIM
    if (!c->events[a] || !c->events[b]) return VCB_ERR_INVALID_ARG;
    if (cudaEventSynchronize(c->events[b]) != cudaSuccess) return VCB_ERR_CUDA;
    float f = 0;
    if (cudaEventElapsedTime(&f, c->events[a], c->events[b]) != cudaSuccess) return VCB_ERR_CUDA;
    *ms = f;
#endif
    return VCB_OK;
}

}  // extern "C"
