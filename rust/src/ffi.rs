//! Raw bindings of include/vcb200.h (hand-written; `bindgen include/vcb200.h` produces the same).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const VCB_OK: c_int = 0;
pub const VCB_ERR_LABEL_OVERFLOW: c_int = 2;
pub const VCB_HIERARCHICAL_MAX: u32 = 0xFFFF_FFFF;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct vcb_runner_config {
    pub diagonal: i32,
    pub hierarchical: u32,
    pub batch_size: i32,
    pub good_min_area: u64,
    pub good_max_area: u64,
    pub is_same_color_a: i32,
    pub is_same_color_b: i32,
    pub deepen_diff: i32,
    pub hollow_neighbours: u64,
    pub key_color: [u8; 4],
    pub keying_action: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vcb_cluster_rec {
    pub sum: [u32; 5],
    pub residue_sum: [u32; 5],
    pub rect: [i32; 4],
    pub num_holes: u32,
    pub depth: u32,
    pub merged_into: u32,
    pub indices_len: u32,
    pub indices_off: u64,
    pub holes_off: u64,
    pub holes_len: u32,
    pub _pad: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vcb_clusters_info {
    pub width: u32,
    pub height: u32,
    pub num_slots: u32,
    pub num_outputs: u32,
    pub indices_total: u64,
    pub holes_total: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vcb_bin_cluster_rec {
    pub rect: [i32; 4],
    pub points_off: u64,
    pub points_len: u32,
    pub _pad: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vcb_bin_clusters_info {
    pub width: u32,
    pub height: u32,
    pub num_clusters: u32,
    pub _pad: u32,
    pub points_total: u64,
    pub rect: [i32; 4],
}

pub enum vcb_ctx {}
pub enum vcb_clusters {}
pub enum vcb_bin_clusters {}
pub enum vcb_components {}

extern "C" {
    pub fn vcb_strerror(status: c_int) -> *const c_char;
    pub fn vcb_last_error(ctx: *const vcb_ctx) -> *const c_char;
    pub fn vcb_ctx_create(device: c_int, out: *mut *mut vcb_ctx) -> c_int;
    pub fn vcb_ctx_destroy(ctx: *mut vcb_ctx);
    pub fn vcb_ctx_set_option(ctx: *mut vcb_ctx, option: c_int, value: i64) -> c_int;
    pub fn vcb_runner_run(ctx: *mut vcb_ctx, cfg: *const vcb_runner_config, rgba: *const u8, width: u32, height: u32,
                          out: *mut *mut vcb_clusters) -> c_int;
    pub fn vcb_clusters_info_get(c: *const vcb_clusters, info: *mut vcb_clusters_info) -> c_int;
    pub fn vcb_clusters_cluster_indices(c: *const vcb_clusters, out: *mut u32) -> c_int;
    pub fn vcb_clusters_outputs(c: *const vcb_clusters, out: *mut u32) -> c_int;
    pub fn vcb_clusters_records(c: *const vcb_clusters, out: *mut vcb_cluster_rec) -> c_int;
    pub fn vcb_clusters_indices_pool(c: *const vcb_clusters, out: *mut u32) -> c_int;
    pub fn vcb_clusters_holes_pool(c: *const vcb_clusters, out: *mut u32) -> c_int;
    pub fn vcb_clusters_free(c: *mut vcb_clusters);
    // device-resident Cluster queries (include/vcb200.h): no pools are built or copied
    pub fn vcb_clusters_perimeters(c: *const vcb_clusters, out: *mut u32) -> c_int;
    pub fn vcb_clusters_neighbours(c: *const vcb_clusters, slot: u32, out: *mut u32, cap: usize, count: *mut usize) -> c_int;
    pub fn vcb_clusters_to_image(c: *const vcb_clusters, slot: u32, hole: c_int, bits_out: *mut u32, cap_words: usize, rect_out: *mut i32) -> c_int;
    // row f1: to_image_with_hole(..).to_clusters(false) of every output cluster, batched on the device (cluster.rs:108-128)
    pub fn vcb_clusters_components(c: *const vcb_clusters, hole: c_int, out: *mut *mut vcb_components) -> c_int;
    pub fn vcb_components_info(r: *const vcb_components, num_outputs: *mut u32, num_components: *mut u64, points_total: *mut u64) -> c_int;
    pub fn vcb_components_get(r: *const vcb_components, comp_off: *mut u32, recs: *mut vcb_bin_cluster_rec, xy: *mut i32) -> c_int;
    pub fn vcb_components_free(r: *mut vcb_components);
    pub fn vcb_binary_to_clusters(ctx: *mut vcb_ctx, bits: *const u32, width: u32, height: u32, diagonal: c_int,
                                  out: *mut *mut vcb_bin_clusters) -> c_int;
    pub fn vcb_bin_clusters_info_get(c: *const vcb_bin_clusters, info: *mut vcb_bin_clusters_info) -> c_int;
    pub fn vcb_bin_clusters_records(c: *const vcb_bin_clusters, out: *mut vcb_bin_cluster_rec) -> c_int;
    pub fn vcb_bin_clusters_points(c: *const vcb_bin_clusters, xy: *mut i32) -> c_int;
    pub fn vcb_bin_clusters_free(c: *mut vcb_bin_clusters);
}
