//! Drop-in replacement of the clustering path of `visioncortex` (v0.8.10) over libvcb200.so (sm_100a CUDA).
//!
//! Same module paths and public items as the reference for this path: `ColorImage`, `BinaryImage`, `Color`, `ColorSum`,
//! `BoundingRect`, `PointI32`, `color_clusters::{Runner, RunnerConfig, KeyingAction, HIERARCHICAL_MAX, Clusters,
//! ClustersView, Cluster, ClusterIndex}`, `clusters::{Clusters, Cluster}`.  `Runner::run` and
//! `BinaryImage::to_clusters` run on the GPU; there is no CPU fallback.  UNCOMPILED in this repository (no Rust toolchain in
//! the build image): kept mechanical so that it can be checked by reading against include/vcb200.h.
pub mod ffi;

use std::ffi::CStr;
use std::sync::Mutex;

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct Color { pub r: u8, pub g: u8, pub b: u8, pub a: u8 }                  // color.rs:8-14

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct PointI32 { pub x: i32, pub y: i32 }                                    // point.rs:8,381

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct BoundingRect { pub left: i32, pub top: i32, pub right: i32, pub bottom: i32 }   // bound.rs:15-21

#[derive(Clone, Copy, Default, Debug)]
pub struct ColorSum { pub r: u32, pub g: u32, pub b: u32, pub a: u32, pub counter: u32 }   // color.rs:40-47

impl ColorSum {
    pub fn average(&self) -> Color {                                               // color.rs:295-302
        Color { r: (self.r / self.counter) as u8, g: (self.g / self.counter) as u8,
                b: (self.b / self.counter) as u8, a: (self.a / self.counter) as u8 }
    }
}

pub struct ColorImage { pub pixels: Vec<u8>, pub width: usize, pub height: usize }   // image/format.rs:29-33

/// image/format.rs:10-14 with the bit container inlined: `bit_vec::BitVec` storage = u32 blocks, LSB first.
pub struct BinaryImage { pub bits: Vec<u32>, pub width: usize, pub height: usize }

struct Ctx(*mut ffi::vcb_ctx);
unsafe impl Send for Ctx {}
static CTX: Mutex<Option<Ctx>> = Mutex::new(None);

fn with_ctx<R>(f: impl FnOnce(*mut ffi::vcb_ctx) -> R) -> R {
    let mut g = CTX.lock().unwrap();
    if g.is_none() {
        let mut p = std::ptr::null_mut();
        let st = unsafe { ffi::vcb_ctx_create(0, &mut p) };
        if st != ffi::VCB_OK { panic!("visioncortex-b200: no CUDA device (there is no CPU fallback)"); }
        *g = Some(Ctx(p));
    }
    f(g.as_ref().unwrap().0)
}

fn check(ctx: *mut ffi::vcb_ctx, st: i32) {
    if st == ffi::VCB_OK { return; }
    if st == ffi::VCB_ERR_LABEL_OVERFLOW { panic!("overflow"); }                    // clusters.rs:322-324
    let msg = unsafe { CStr::from_ptr(ffi::vcb_last_error(ctx)) }.to_string_lossy().into_owned();
    panic!("visioncortex-b200: {}", msg);
}

pub mod color_clusters {
    use super::*;

    pub const HIERARCHICAL_MAX: u32 = u32::MAX;                                    // builder.rs:46
    pub const ZERO: ClusterIndex = ClusterIndex(0);                                // builder.rs:45

    #[derive(Clone, Copy, Default, PartialEq, Eq)]
    pub enum KeyingAction { #[default] Keep, Discard }                             // builder.rs:6-11

    #[derive(Copy, Clone, Default, Eq, Ord, Hash, PartialEq, PartialOrd)]
    pub struct ClusterIndex(pub u32);                                              // container.rs:13-16

    pub struct RunnerConfig {                                                      // runner.rs:9-21
        pub diagonal: bool, pub hierarchical: u32, pub batch_size: i32, pub good_min_area: usize, pub good_max_area: usize,
        pub is_same_color_a: i32, pub is_same_color_b: i32, pub deepen_diff: i32, pub hollow_neighbours: usize,
        pub key_color: Color, pub keying_action: KeyingAction,
    }
    impl Default for RunnerConfig {                                                // runner.rs:23-39
        fn default() -> Self {
            Self { diagonal: false, hierarchical: HIERARCHICAL_MAX, batch_size: 25600, good_min_area: 16,
                   good_max_area: 256 * 256, is_same_color_a: 4, is_same_color_b: 1, deepen_diff: 64, hollow_neighbours: 1,
                   key_color: Color::default(), keying_action: KeyingAction::default() }
        }
    }

    #[derive(Default)]
    pub struct Cluster {                                                           // cluster.rs:7-17
        pub indices: Vec<u32>, pub holes: Vec<u32>, pub num_holes: u32, pub depth: u32,
        pub sum: ColorSum, pub residue_sum: ColorSum, pub rect: BoundingRect, pub merged_into: ClusterIndex,
    }
    impl Cluster {
        pub fn area(&self) -> usize { self.indices.len() }
        pub fn color(&self) -> Color { self.sum.average() }
        pub fn residue_color(&self) -> Color { self.residue_sum.average() }
    }

    pub struct Clusters {                                                          // container.rs:4-11
        pub width: u32, pub height: u32,
        pub(crate) pixels: Vec<u8>, pub(crate) clusters: Vec<Cluster>,
        pub(crate) cluster_indices: Vec<ClusterIndex>, pub(crate) clusters_output: Vec<ClusterIndex>,
    }
    pub struct ClustersView<'a> {                                                  // container.rs:43-50
        pub width: u32, pub height: u32, pub pixels: &'a [u8], pub clusters: &'a [Cluster],
        pub cluster_indices: &'a [ClusterIndex], pub clusters_output: &'a [ClusterIndex],
    }
    impl Clusters {
        pub fn output_len(&self) -> usize { self.clusters_output.len() }
        pub fn view(&self) -> ClustersView<'_> {
            ClustersView { width: self.width, height: self.height, pixels: &self.pixels, clusters: &self.clusters,
                           cluster_indices: &self.cluster_indices, clusters_output: &self.clusters_output }
        }
        pub fn take_image(self) -> ColorImage {
            ColorImage { pixels: self.pixels, width: self.width as usize, height: self.height as usize }
        }
    }
    impl ClustersView<'_> {
        pub fn iter(&self) -> impl Iterator<Item = &Cluster> {
            self.clusters_output.iter().map(move |i| &self.clusters[i.0 as usize])
        }
        pub fn get_cluster(&self, index: ClusterIndex) -> &Cluster { &self.clusters[index.0 as usize] }
    }

    pub struct Runner { config: RunnerConfig, image: ColorImage }                 // runner.rs:4-7
    impl Runner {
        pub fn new(config: RunnerConfig, image: ColorImage) -> Self { Self { config, image } }
        pub fn init(&mut self, image: ColorImage) { self.image = image; }

        /// runner.rs:104-106 — the whole computation runs on the GPU (vcb_runner_run)
        pub fn run(self) -> Clusters {
            assert!(self.config.is_same_color_a < 8);                              // runner.rs:78
            let c = &self.config;
            let cfg = ffi::vcb_runner_config {
                diagonal: c.diagonal as i32, hierarchical: c.hierarchical, batch_size: c.batch_size,
                good_min_area: c.good_min_area as u64, good_max_area: c.good_max_area as u64,
                is_same_color_a: c.is_same_color_a, is_same_color_b: c.is_same_color_b, deepen_diff: c.deepen_diff,
                hollow_neighbours: c.hollow_neighbours as u64,
                key_color: [c.key_color.r, c.key_color.g, c.key_color.b, c.key_color.a],
                keying_action: match c.keying_action { KeyingAction::Keep => 0, KeyingAction::Discard => 1 },
            };
            let (w, h) = (self.image.width as u32, self.image.height as u32);
            with_ctx(|ctx| unsafe {
                let mut hnd = std::ptr::null_mut();
                check(ctx, ffi::vcb_runner_run(ctx, &cfg, self.image.pixels.as_ptr(), w, h, &mut hnd));
                let mut info = ffi::vcb_clusters_info::default();
                check(ctx, ffi::vcb_clusters_info_get(hnd, &mut info));
                let mut labels = vec![0u32; (w * h) as usize];
                let mut outs = vec![0u32; info.num_outputs as usize];
                let mut recs = vec![ffi::vcb_cluster_rec::default(); info.num_slots as usize];
                let mut ipool = vec![0u32; info.indices_total as usize];
                let mut hpool = vec![0u32; info.holes_total as usize];
                check(ctx, ffi::vcb_clusters_cluster_indices(hnd, labels.as_mut_ptr()));
                check(ctx, ffi::vcb_clusters_outputs(hnd, outs.as_mut_ptr()));
                check(ctx, ffi::vcb_clusters_records(hnd, recs.as_mut_ptr()));
                check(ctx, ffi::vcb_clusters_indices_pool(hnd, ipool.as_mut_ptr()));
                check(ctx, ffi::vcb_clusters_holes_pool(hnd, hpool.as_mut_ptr()));
                ffi::vcb_clusters_free(hnd);
                let clusters = recs.iter().map(|r| Cluster {
                    indices: ipool[r.indices_off as usize..(r.indices_off + r.indices_len as u64) as usize].to_vec(),
                    holes: hpool[r.holes_off as usize..(r.holes_off + r.holes_len as u64) as usize].to_vec(),
                    num_holes: r.num_holes, depth: r.depth,
                    sum: ColorSum { r: r.sum[0], g: r.sum[1], b: r.sum[2], a: r.sum[3], counter: r.sum[4] },
                    residue_sum: ColorSum { r: r.residue_sum[0], g: r.residue_sum[1], b: r.residue_sum[2],
                                            a: r.residue_sum[3], counter: r.residue_sum[4] },
                    rect: BoundingRect { left: r.rect[0], top: r.rect[1], right: r.rect[2], bottom: r.rect[3] },
                    merged_into: ClusterIndex(r.merged_into),
                }).collect();
                Clusters { width: w, height: h, pixels: self.image.pixels, clusters,
                           cluster_indices: labels.into_iter().map(ClusterIndex).collect(),
                           clusters_output: outs.into_iter().map(ClusterIndex).collect() }
            })
        }
    }
}

pub mod clusters {
    use super::*;

    #[derive(Default)]
    pub struct Cluster { pub points: Vec<PointI32>, pub rect: BoundingRect }      // clusters.rs:7-11
    #[derive(Default)]
    pub struct Clusters { pub clusters: Vec<Cluster>, pub rect: BoundingRect }    // clusters.rs:14-18

    impl BinaryImage {
        /// clusters.rs:270-349 — runs on the GPU (vcb_binary_to_clusters)
        pub fn to_clusters(&self, diagonal: bool) -> Clusters {
            with_ctx(|ctx| unsafe {
                let mut hnd = std::ptr::null_mut();
                check(ctx, ffi::vcb_binary_to_clusters(ctx, self.bits.as_ptr(), self.width as u32, self.height as u32,
                                                       diagonal as i32, &mut hnd));
                let mut info = ffi::vcb_bin_clusters_info::default();
                check(ctx, ffi::vcb_bin_clusters_info_get(hnd, &mut info));
                let mut recs = vec![ffi::vcb_bin_cluster_rec::default(); info.num_clusters as usize];
                let mut xy = vec![0i32; 2 * info.points_total as usize];
                check(ctx, ffi::vcb_bin_clusters_records(hnd, recs.as_mut_ptr()));
                check(ctx, ffi::vcb_bin_clusters_points(hnd, xy.as_mut_ptr()));
                ffi::vcb_bin_clusters_free(hnd);
                let clusters = recs.iter().map(|r| Cluster {
                    points: (r.points_off..r.points_off + r.points_len as u64)
                        .map(|i| PointI32 { x: xy[2 * i as usize], y: xy[2 * i as usize + 1] }).collect(),
                    rect: BoundingRect { left: r.rect[0], top: r.rect[1], right: r.rect[2], bottom: r.rect[3] },
                }).collect();
                Clusters { clusters, rect: BoundingRect { left: info.rect[0], top: info.rect[1], right: info.rect[2], bottom: info.rect[3] } }
            })
        }
    }
}
