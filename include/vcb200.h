/*
 * vcb200.h — C ABI of the B200-native clustering path of visioncortex.
 *
 * The reference (visioncortex v0.8.10, pure Rust) has no FFI of its own; the path is reached through
 * ordinary Rust methods.  Every entry point below names the reference method it replaces (file:line under
 * the reference tree).  A replacement crate (see rust/ and INTEGRATION.md) keeps the Rust API surface
 * (`ColorImage`, `BinaryImage`, `color_clusters::{Runner, RunnerConfig, Clusters, ClustersView, Cluster}`,
 * `clusters::{Clusters, Cluster}`) and forwards into these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a vcb_status (0 = OK) and never aborts or unwinds;
 *   - a vcb_ctx owns one CUDA device, its streams and a workspace arena; it is NOT thread-safe, distinct
 *     contexts may be used from distinct threads; result handles are immutable once returned;
 *   - accessors copy into caller-owned buffers (sizes come from vcb_clusters_info / vcb_bin_clusters_info);
 *   - there is no CPU fallback: if the device or the CUDA runtime is unavailable, vcb_ctx_create fails.
 */
#ifndef VCB200_H
#define VCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vcb_status {
    VCB_OK = 0,
    VCB_ERR_INVALID_ARG = 1,     /* mirrors assert!(is_same_color_a < 8) (color_clusters/runner.rs:78), size mismatch, batch_size == 0 */
    VCB_ERR_LABEL_OVERFLOW = 2,  /* mirrors panic!("overflow") (clusters.rs:322-324) */
    VCB_ERR_UNSUPPORTED = 3,     /* hierarchical stage on a diagonal=true input whose resurrected slot was overwritten (DESIGN.md
                                    section 1); every other order-dependent quirk input is replayed exactly by the sequential resolver */
    VCB_ERR_CUDA = 4,
    VCB_ERR_NCCL = 5,            /* reserved, never returned: no path makes an NCCL call (the row-band exchange is in-kernel peer
                                    memory traffic; batches shard without a collective) -- kept so the numbering stays stable */
    VCB_ERR_NOMEM = 6
} vcb_status;

/* KeyingAction (color_clusters/builder.rs:6-11) */
enum { VCB_KEYING_KEEP = 0, VCB_KEYING_DISCARD = 1 };

/* HIERARCHICAL_MAX (color_clusters/builder.rs:46) */
#define VCB_HIERARCHICAL_MAX 0xFFFFFFFFu

/* RunnerConfig (color_clusters/runner.rs:9-21); field meaning identical. */
typedef struct vcb_runner_config {
    int32_t  diagonal;
    uint32_t hierarchical;
    int32_t  batch_size;
    uint64_t good_min_area;
    uint64_t good_max_area;
    int32_t  is_same_color_a;
    int32_t  is_same_color_b;
    int32_t  deepen_diff;
    uint64_t hollow_neighbours;
    uint8_t  key_color[4];       /* r,g,b,a; (0,0,0,0) = no key (Color::default()) */
    int32_t  keying_action;
} vcb_runner_config;

/* One slot of `Clusters.clusters` (color_clusters/cluster.rs:7-17).  area() == indices_len. */
typedef struct vcb_cluster_rec {
    uint32_t sum[5];             /* ColorSum r,g,b,a,counter (color.rs:40-47), u32 wrapping */
    uint32_t residue_sum[5];
    int32_t  rect[4];            /* BoundingRect left,top,right,bottom (bound.rs:15-21) */
    uint32_t num_holes;
    uint32_t depth;
    uint32_t merged_into;
    uint32_t indices_len;
    uint64_t indices_off;        /* offset into the indices pool */
    uint64_t holes_off;          /* offset into the holes pool */
    uint32_t holes_len;
    uint32_t _pad;
} vcb_cluster_rec;

typedef struct vcb_clusters_info {
    uint32_t width, height;
    uint32_t num_slots;          /* clusters.len(), dead slots included */
    uint32_t num_outputs;        /* clusters_output.len(); after a hierarchical stage up to 2 * num_slots (a cluster output as isolated can
                                    grow and be output again, builder.rs:444-450) */
    uint64_t indices_total;      /* sum of indices_len over all slots */
    uint64_t holes_total;
} vcb_clusters_info;

/* clusters::Cluster (clusters.rs:7-11) */
typedef struct vcb_bin_cluster_rec {
    int32_t  rect[4];
    uint64_t points_off;         /* offset (in points) into the points pool */
    uint32_t points_len;
    uint32_t _pad;
} vcb_bin_cluster_rec;

typedef struct vcb_bin_clusters_info {
    uint32_t width, height;
    uint32_t num_clusters;       /* clusters.len() after the empty-slot filter (clusters.rs:345) */
    uint32_t _pad;
    uint64_t points_total;
    int32_t  rect[4];            /* Clusters.rect (clusters.rs:14-18,302) */
} vcb_bin_clusters_info;

typedef struct vcb_ctx vcb_ctx;
typedef struct vcb_clusters vcb_clusters;
typedef struct vcb_bin_clusters vcb_bin_clusters;

/* library */
const char *vcb_strerror(int status);
const char *vcb_last_error(const vcb_ctx *ctx);      /* detail text of the last failure on this context */
uint32_t    vcb_abi_version(void);

/* RunnerConfig::default() (color_clusters/runner.rs:23-39) */
void vcb_runner_config_default(vcb_runner_config *cfg);

/* context: one device, streams, workspace arena */
int  vcb_ctx_create(int device, vcb_ctx **out);
/* Result handles of the context may outlive this call: the context is then released by the last vcb_clusters_free (no other
 * call may be made on a destroyed context). */
void vcb_ctx_destroy(vcb_ctx *ctx);
int  vcb_ctx_device(const vcb_ctx *ctx);
/* options.  VCB_OPT_MATERIALIZE_INDICES (default 0): build the ordered `Cluster.indices` pool (color_clusters/cluster.rs:8)
 * during the run; when 0 it is built on the first vcb_clusters_indices_pool call from the retained image. */
enum { VCB_OPT_MATERIALIZE_INDICES = 1 };
int  vcb_ctx_set_option(vcb_ctx *ctx, int option, int64_t value);

/* Runner::new(config, image).run() (color_clusters/runner.rs:52-57,104-106): host RGBA8 row-major in, result handle out. */
int vcb_runner_run(vcb_ctx *ctx, const vcb_runner_config *cfg, const uint8_t *rgba,
                   uint32_t width, uint32_t height, vcb_clusters **out);

/* The same call for n independent images (how a vtracer-style caller batches).  out[] receives n handles.
 * Images are pipelined over the context's streams; there is no cross-image communication. */
int vcb_runner_run_batch(vcb_ctx *ctx, const vcb_runner_config *cfg, const uint8_t *const *rgba,
                         const uint32_t *width, const uint32_t *height, size_t n, vcb_clusters **out);

/* Host-to-host batch: n equally sized images in, the caller's result buffers out.  The images are processed in sub-batches;
 * the host->device copy of the next sub-batch and the device->host copy of the previous results run on their own streams
 * while the current sub-batch computes (full-duplex PCIe), so this is the call to use when everything lives in host memory
 * (what the Rust shim does per image with vcb_runner_run + accessors).  records_cap / outputs_cap are capacities in
 * elements; an image whose result does not fit gets status VCB_ERR_NOMEM (info is still filled).  Pinned buffers copy faster. */
typedef struct vcb_host_result {
    uint32_t        *cluster_indices;   /* width*height */
    vcb_cluster_rec *records;           /* records_cap */
    uint32_t        *outputs;           /* outputs_cap */
    uint32_t         records_cap, outputs_cap;
    vcb_clusters_info info;             /* out */
    int              status;            /* out */
    /* Optional compact records (set live_slots on EVERY result of the call, or on none).  After a hierarchical stage most slots
     * of `clusters` are Cluster::default() (merged away in stage 1: every field zero); their 96-byte records are most of the
     * device->host traffic of a batch.  With live_slots != NULL only the slots that are not default -- any of indices_len,
     * holes_len, residue_sum.counter, merged_into, depth, num_holes non-zero -- are transferred: records[k] is the record of
     * slot live_slots[k] (ascending), k < num_live <= records_cap (= capacity of both arrays); every slot not listed is default
     * (its indices_off / holes_off, this library's own layout fields, are the running totals and can be rebuilt as exclusive scans).
     * compact = 1 when the result came back in this form, 0 when all num_slots records were written as usual (runs without a
     * hierarchical stage). */
    uint32_t        *live_slots;        /* records_cap, or NULL */
    uint32_t         num_live;          /* out */
    uint32_t         compact;           /* out */
} vcb_host_result;
int vcb_runner_run_batch_host(vcb_ctx *ctx, const vcb_runner_config *cfg, const uint8_t *const *rgba, uint32_t width,
                              uint32_t height, size_t n, vcb_host_result *results);

/* One giant image split into row bands over the GPUs of one box (BASELINE config 5).  ctxs[0..nctx) are contexts on
 * distinct devices with peer access; ctxs[0] owns the result.  The workspace arrays are ONE virtual address range striped
 * over the devices; each device runs the per-pixel kernels on its band and reaches the other bands' records (border label
 * equivalences included) through NVLink peer loads / stores / atomics; the spanning-forest "union pass" over all bands
 * runs once on ctxs[0].  Result: identical to vcb_runner_run. */
int vcb_runner_run_banded(vcb_ctx *const *ctxs, int nctx, const vcb_runner_config *cfg, const uint8_t *rgba,
                          uint32_t width, uint32_t height, vcb_clusters **out);

/* Device-resident variant: d_rgba points to n images of identical size stored back to back in device memory
 * of ctx's device.  Results stay on the device until an accessor is called. */
int vcb_runner_run_device(vcb_ctx *ctx, const vcb_runner_config *cfg, const uint8_t *d_rgba,
                          uint32_t width, uint32_t height, size_t n, vcb_clusters **out);

/* Clusters / ClustersView (color_clusters/container.rs:4-117) */
int  vcb_clusters_info_get(const vcb_clusters *c, vcb_clusters_info *info);
int  vcb_clusters_cluster_indices(const vcb_clusters *c, uint32_t *out /* width*height */);
int  vcb_clusters_outputs(const vcb_clusters *c, uint32_t *out /* num_outputs */);
int  vcb_clusters_records(const vcb_clusters *c, vcb_cluster_rec *out /* num_slots */);
int  vcb_clusters_indices_pool(const vcb_clusters *c, uint32_t *out /* indices_total */);
int  vcb_clusters_holes_pool(const vcb_clusters *c, uint32_t *out /* holes_total */);
/* ClustersView::to_color_image (color_clusters/container.rs:104-116) */
int  vcb_clusters_to_color_image(const vcb_clusters *c, uint8_t *rgba_out /* width*height*4 */);
/* Device-resident Cluster queries: computed on the GPU from what the result keeps in HBM
 * (stage-1 / final labels, the stage-2 merge forest, the records); no indices / holes pool is built or copied.
 *   perimeters : Cluster::perimeter (color_clusters/cluster.rs:46-48 -> shape/geometry.rs:52-75) of EVERY cluster of the image in one
 *                pass; out[slot] for slot < num_slots, 0 for clusters that hold no pixels
 *   neighbours : Cluster::neighbours (cluster.rs:132-165): sorted indices of the clusters adjacent to `slot`; *count = length
 *                (VCB_ERR_NOMEM with *count set if cap is too small; VCB_ERR_INVALID_ARG for a cluster without pixels, where
 *                the reference unwraps indices.first())
 *   to_image   : Cluster::to_image_with_hole(parent.width, hole) (cluster.rs:60-80) as BitVec storage of the rect-sized image
 *                (u32 blocks, pixel i at block i/32 bit i%32, image/format.rs:10-14); rect_out = the cluster's BoundingRect
 * Results of the sequential resolver (unclean key colours) return VCB_ERR_UNSUPPORTED here: use the pools. */
int  vcb_clusters_perimeters(const vcb_clusters *c, uint32_t *out /* num_slots */);
int  vcb_clusters_neighbours(const vcb_clusters *c, uint32_t slot, uint32_t *out, size_t cap, size_t *count);
int  vcb_clusters_to_image(const vcb_clusters *c, uint32_t slot, int hole, uint32_t *bits_out, size_t cap_words, int32_t rect_out[4]);
/* SURVEY.md section 8(f) row f1 -- the clustering step of Cluster::to_compound_path (color_clusters/cluster.rs:108-128):
 * `self.to_image_with_hole(parent.width, hole).to_clusters(false)` for EVERY output cluster of the image, batched on the device:
 * the clusters' images are written straight from the result in HBM (no pools), grouped into size classes, and each class runs
 * through the BinaryImage::to_clusters kernels (clusters.rs:270-349) as one batch.  Per output cluster k (position in
 * clusters_output) the components comp_off[k] .. comp_off[k+1]: rect and points are relative to the cluster's image, exactly as
 * the reference's inner `cluster.rect` / `cluster.points` (the caller adds self.rect.left / top, cluster.rs:120-123). */
typedef struct vcb_components vcb_components;
int  vcb_clusters_components(const vcb_clusters *c, int hole, vcb_components **out);
int  vcb_components_info(const vcb_components *r, uint32_t *num_outputs, uint64_t *num_components, uint64_t *points_total);
int  vcb_components_get(const vcb_components *r, uint32_t *comp_off /* num_outputs + 1 */, vcb_bin_cluster_rec *recs, int32_t *xy);
void vcb_components_free(vcb_components *r);
/* device pointer of the label map (width*height u32) for callers that stay on the GPU; NULL if released */
const uint32_t *vcb_clusters_device_labels(const vcb_clusters *c);
/* Clusters::take_image (container.rs:34-40): the retained RGBA8 input, copied to the caller */
int  vcb_clusters_take_image(const vcb_clusters *c, uint8_t *rgba_out /* width*height*4 */);
void vcb_clusters_free(vcb_clusters *c);

/* BinaryImage::to_clusters(diagonal) (clusters.rs:270-349).  bits = bit_vec::BitVec storage: u32 blocks,
 * pixel i = y*width+x lives in block i/32 at bit i%32 (LSB first); ceil(width*height/32) blocks. */
int  vcb_binary_to_clusters(vcb_ctx *ctx, const uint32_t *bits, uint32_t width, uint32_t height,
                            int diagonal, vcb_bin_clusters **out);
int  vcb_binary_to_clusters_device(vcb_ctx *ctx, const uint32_t *d_bits, uint32_t width, uint32_t height,
                                   int diagonal, vcb_bin_clusters **out);
int  vcb_bin_clusters_info_get(const vcb_bin_clusters *c, vcb_bin_clusters_info *info);
int  vcb_bin_clusters_records(const vcb_bin_clusters *c, vcb_bin_cluster_rec *out /* num_clusters */);
int  vcb_bin_clusters_points(const vcb_bin_clusters *c, int32_t *xy_out /* 2*points_total, PointI32{x,y} */);
void vcb_bin_clusters_free(vcb_bin_clusters *c);

/* measurement hooks used by bench.py (not part of the reference surface) */
int  vcb_ctx_synchronize(vcb_ctx *ctx);
uint64_t vcb_ctx_kernel_launches(const vcb_ctx *ctx);            /* kernels launched by this context so far */
/* device time (ms) of the dominant kernel, accumulated by CUDA events on its launching stream since the last reset */
int  vcb_ctx_profile_reset(vcb_ctx *ctx, int enable);
int  vcb_ctx_profile_read(vcb_ctx *ctx, const char *kernel_name, double *total_ms, uint64_t *launches);
/* counters of the last stage-1 run on this context: out[0] = candidate events, out[1] = NEW events, out[2] = leaves after
 * the trivial-leaf contraction, out[3] = spanning-forest rounds, out[4] = error bits, out[5] = leaves resurrected through the
 * stale cluster_upleft label (diagonal = true, builder.rs:301-305), out[6] = passes of the stage-1 pipeline (1 unless out[5] > 0) */
int  vcb_ctx_last_counters(const vcb_ctx *ctx, uint32_t out[8]);
/* CUDA events on the context's stream: record into slot 0..15, elapsed milliseconds between two recorded slots */
int  vcb_ctx_event_record(vcb_ctx *ctx, int slot);
int  vcb_ctx_event_elapsed_ms(vcb_ctx *ctx, int slot_a, int slot_b, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* VCB200_H */
