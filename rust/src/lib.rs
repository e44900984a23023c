//! Drop-in replacement of the clustering path of `visioncortex` (v0.8.10) over libvcb200.so (sm_100a CUDA).
//!
//! Same module paths and public items as the reference for this path (file:line under the reference's `src/`):
//!   `Color`, `ColorSum`, `PointI32`, `BoundingRect`                                   color.rs:8-47, point.rs:381, bound.rs:15-21
//!   `ColorImage { pixels, width, height }` + `new_w_h / get_pixel / set_pixel`         image/format.rs:29-33,268-327
//!   `BinaryImage { pixels: BitVec, width, height }` + `new_w_h / get_pixel(_safe) / set_pixel / to_clusters`
//!                                                                                     image/format.rs:10-14,43-77; clusters.rs:270-349
//!   `color_clusters::{Runner, RunnerConfig, KeyingAction, HIERARCHICAL_MAX, ZERO, Builder, IncrementalBuilder, NeighbourInfo,
//!                     Clusters, ClustersView, Cluster, ClusterIndex}`                  color_clusters/mod.rs:12-20
//!   `clusters::{Clusters, Cluster}`                                                    clusters.rs:7-18
//! `Runner::run` / `Runner::start` and `BinaryImage::to_clusters` run on the GPU; there is no CPU fallback.  The query
//! methods of `ClustersView` / `Cluster` are host code over the vectors the run returned, as in the reference.
//! `visioncortex_b200/api.py` is the executable twin of this file (same structure, tested through the same C ABI).
//!
//! UNCOMPILED in this repository (no Rust toolchain in the build image): kept mechanical so that it can be checked by
//! reading against include/vcb200.h.  The only dependency is `bit-vec = "0.6"`, as in the reference.
//!
//! Documented deviations: `Builder::{same, diff, deepen, hollow}` take arbitrary host closures in
//! the reference and cannot be offloaded: they panic here.  `IncrementalBuilder::tick()` performs the whole computation on
//! the first call and returns `true` on the second; `progress()` is 0 before and 100 after; intermediate `view()` states are
//! not reproduced (the final `result()` is bit-exact).
pub mod ffi;

pub use bit_vec::BitVec;
use std::ffi::CStr;
use std::sync::Mutex;

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct Color { pub r: u8, pub g: u8, pub b: u8, pub a: u8 }                  // color.rs:8-14
impl Color {
    pub fn new(r: u8, g: u8, b: u8) -> Self { Self::new_rgba(r, g, b, 255) }     // color.rs:58-60
    pub fn new_rgba(r: u8, g: u8, b: u8, a: u8) -> Self { Self { r, g, b, a } }  // color.rs:62-64
}

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct PointI32 { pub x: i32, pub y: i32 }                                    // point.rs:8,381

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct BoundingRect { pub left: i32, pub top: i32, pub right: i32, pub bottom: i32 }   // bound.rs:15-21
impl BoundingRect {
    pub fn width(self) -> i32 { self.right - self.left }                         // bound.rs:86-88
    pub fn height(self) -> i32 { self.bottom - self.top }                        // bound.rs:91-93
    pub fn is_empty(self) -> bool { self.width() == 0 && self.height() == 0 }    // bound.rs:101-103
}

#[derive(Clone, Copy, Default, PartialEq, Eq, Debug)]
pub struct ColorSum { pub r: u32, pub g: u32, pub b: u32, pub a: u32, pub counter: u32 }   // color.rs:40-47
impl ColorSum {
    pub fn average(&self) -> Color {                                               // color.rs:295-302
        Color { r: (self.r / self.counter) as u8, g: (self.g / self.counter) as u8,
                b: (self.b / self.counter) as u8, a: (self.a / self.counter) as u8 }
    }
}

/// image/format.rs:29-33
#[derive(Clone, Default)]
pub struct ColorImage { pub pixels: Vec<u8>, pub width: usize, pub height: usize }
impl ColorImage {
    pub fn new() -> Self { Default::default() }                                    // format.rs:269-271
    pub fn new_w_h(width: usize, height: usize) -> Self {                          // format.rs:273-279
        Self { pixels: vec![0; width * height * 4], width, height }
    }
    pub fn get_pixel(&self, x: usize, y: usize) -> Color { self.get_pixel_at(y * self.width + x) }   // format.rs:289-292
    pub fn get_pixel_safe(&self, x: i32, y: i32) -> Option<Color> {               // format.rs:298-304
        if x >= 0 && x < self.width as i32 && y >= 0 && y < self.height as i32 { return Some(self.get_pixel(x as usize, y as usize)); }
        None
    }
    pub fn get_pixel_at(&self, index: usize) -> Color {                           // format.rs:306-314
        let i = index * 4;
        Color::new_rgba(self.pixels[i], self.pixels[i + 1], self.pixels[i + 2], self.pixels[i + 3])
    }
    pub fn set_pixel(&mut self, x: usize, y: usize, color: &Color) { self.set_pixel_at(y * self.width + x, color); }   // format.rs:316-319
    pub fn set_pixel_at(&mut self, index: usize, color: &Color) {                 // format.rs:321-327
        let i = index * 4;
        self.pixels[i] = color.r; self.pixels[i + 1] = color.g; self.pixels[i + 2] = color.b; self.pixels[i + 3] = color.a;
    }
}

/// image/format.rs:10-14.  `pixels.storage()` (u32 blocks, pixel i at block i/32, bit i%32) is what crosses the C ABI.
#[derive(Debug, Clone, Default)]
pub struct BinaryImage { pub pixels: BitVec, pub width: usize, pub height: usize }
impl BinaryImage {
    pub fn new_w_h(width: usize, height: usize) -> BinaryImage {                  // format.rs:44-50
        BinaryImage { pixels: BitVec::from_elem(width * height, false), width, height }
    }
    pub fn get_pixel(&self, x: usize, y: usize) -> bool { self.pixels.get(y * self.width + x).unwrap() }   // format.rs:58-61
    pub fn get_pixel_safe(&self, x: i32, y: i32) -> bool {                        // format.rs:66-72
        if x >= 0 && x < self.width as i32 && y >= 0 && y < self.height as i32 { return self.get_pixel(x as usize, y as usize); }
        false
    }
    pub fn set_pixel(&mut self, x: usize, y: usize, v: bool) { self.pixels.set(y * self.width + x, v); }   // format.rs:74-77
}

struct Ctx(*mut ffi::vcb_ctx);
unsafe impl Send for Ctx {}
static CTX: Mutex<Option<Ctx>> = Mutex::new(None);

/// One context on device 0, opened on first use.  A panic raised while the lock is held (e.g. `panic!("overflow")`, which
/// callers may catch) must not poison every later call: the guard is recovered from a poisoned mutex.
fn with_ctx<R>(f: impl FnOnce(*mut ffi::vcb_ctx) -> R) -> R {
    let ctx = {
        let mut g = CTX.lock().unwrap_or_else(|e| e.into_inner());
        if g.is_none() {
            let mut p = std::ptr::null_mut();
            let st = unsafe { ffi::vcb_ctx_create(0, &mut p) };
            if st != ffi::VCB_OK { panic!("visioncortex-b200: no CUDA device (there is no CPU fallback)"); }
            *g = Some(Ctx(p));
        }
        g.as_ref().unwrap().0
    };                                                                             // lock released before the call: check() may panic
    f(ctx)
}

fn check(ctx: *mut ffi::vcb_ctx, st: i32) {
    if st == ffi::VCB_OK { return; }
    if st == ffi::VCB_ERR_LABEL_OVERFLOW { panic!("overflow"); }                    // clusters.rs:322-324
    let msg = unsafe { CStr::from_ptr(ffi::vcb_last_error(ctx)) }.to_string_lossy().into_owned();
    panic!("visioncortex-b200: {}", msg);
}

pub mod color_clusters {
    use super::*;
    use std::collections::HashSet;

    pub const HIERARCHICAL_MAX: u32 = u32::MAX;                                    // builder.rs:46
    pub const ZERO: ClusterIndex = ClusterIndex(0);                                // builder.rs:45

    #[derive(Clone, Copy, Default, PartialEq, Eq)]
    pub enum KeyingAction { #[default] Keep, Discard }                             // builder.rs:6-11

    pub type ClusterIndexElem = u32;
    #[derive(Copy, Clone, Default, Eq, Ord, Hash, PartialEq, PartialOrd)]
    pub struct ClusterIndex(pub ClusterIndexElem);                                 // container.rs:13-16

    pub struct NeighbourInfo { pub index: ClusterIndex, pub diff: i32 }            // builder.rs:34-37

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

    #[derive(Clone, Default)]
    pub struct Cluster {                                                           // cluster.rs:7-17
        pub indices: Vec<u32>, pub holes: Vec<u32>, pub num_holes: u32, pub depth: u32,
        pub sum: ColorSum, pub residue_sum: ColorSum, pub rect: BoundingRect, pub merged_into: ClusterIndex,
    }
    impl Cluster {
        pub fn new() -> Self { Self::default() }
        pub fn area(&self) -> usize { self.indices.len() }                         // cluster.rs:30-32
        pub fn iter(&self) -> impl Iterator<Item = &u32> { self.indices.iter() }   // cluster.rs:34-36
        pub fn color(&self) -> Color { self.sum.average() }                        // cluster.rs:38-40
        pub fn residue_color(&self) -> Color { self.residue_sum.average() }        // cluster.rs:42-44

        /// cluster.rs:46-48 -> shape/geometry.rs:52-75: set pixels with an unset or out-of-image 4-neighbour
        pub fn perimeter(&self, parent: &ClustersView) -> u32 {
            let image = self.to_image(parent);
            let mut n = 0;
            for y in 0..image.height as i32 {
                for x in 0..image.width as i32 {
                    if image.get_pixel(x as usize, y as usize)
                        && (!image.get_pixel_safe(x - 1, y) || !image.get_pixel_safe(x + 1, y)
                            || !image.get_pixel_safe(x, y - 1) || !image.get_pixel_safe(x, y + 1)) {
                        n += 1;
                    }
                }
            }
            n
        }
        pub fn to_image(&self, parent: &ClustersView) -> BinaryImage { self.to_image_with_hole(parent.width, true) }   // cluster.rs:53-55
        pub fn to_image_with_hole(&self, parent_width: u32, hole: bool) -> BinaryImage {                             // cluster.rs:60-80
            let width = self.rect.width() as usize;
            let height = self.rect.height() as usize;
            let mut image = BinaryImage::new_w_h(width, height);
            for &i in self.iter() {
                let x = (i as i32 % parent_width as i32) - self.rect.left;
                let y = (i as i32 / parent_width as i32) - self.rect.top;
                image.set_pixel(x as usize, y as usize, true);
            }
            if hole {
                for &i in self.holes.iter() {
                    let x = (i as i32 % parent_width as i32) - self.rect.left;
                    let y = (i as i32 / parent_width as i32) - self.rect.top;
                    image.set_pixel(x as usize, y as usize, false);
                }
            }
            image
        }
        pub fn render_to_binary_image(&self, parent: &ClustersView, image: &mut BinaryImage) {   // cluster.rs:82-88
            for &i in self.iter() { image.set_pixel((i % parent.width) as usize, (i / parent.width) as usize, true); }
        }
        pub fn render_to_color_image(&self, parent: &ClustersView, image: &mut ColorImage) {      // cluster.rs:90-93
            let color = self.residue_color();
            self.render_to_color_image_with_color(parent, image, &color);
        }
        pub fn render_to_color_image_with_color(&self, parent: &ClustersView, image: &mut ColorImage, color: &Color) {   // cluster.rs:95-101
            for &i in self.iter() { image.set_pixel((i % parent.width) as usize, (i / parent.width) as usize, color); }
        }
        /// cluster.rs:132-165: distinct labels != ZERO, != self among the 4-neighbours of every pixel, ascending
        pub fn neighbours(&self, parent: &ClustersView) -> Vec<ClusterIndex> {
            let myself = parent.get_cluster_at(*self.indices.first().unwrap());
            let mut neighbours = HashSet::new();
            for &i in self.iter() {
                let x = i % parent.width;
                let y = i / parent.width;
                for k in 0..4 {
                    let index = match k {
                        0 => if y > 0 { parent.cluster_indices[(parent.width * (y - 1) + x) as usize] } else { ZERO },
                        1 => if y < parent.height - 1 { parent.cluster_indices[(parent.width * (y + 1) + x) as usize] } else { ZERO },
                        2 => if x > 0 { parent.cluster_indices[(parent.width * y + (x - 1)) as usize] } else { ZERO },
                        _ => if x < parent.width - 1 { parent.cluster_indices[(parent.width * y + (x + 1)) as usize] } else { ZERO },
                    };
                    if index != ZERO && index != myself { neighbours.insert(index); }
                }
            }
            let mut list: Vec<ClusterIndex> = neighbours.into_iter().collect();
            list.sort();
            list
        }
        // to_compound_path (cluster.rs:108-128) belongs to the path-tracing layer: SURVEY.md section 8(f), not part of this path.
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
        pub fn output_len(&self) -> usize { self.clusters_output.len() }           // container.rs:18-20
        pub fn view(&self) -> ClustersView<'_> {                                   // container.rs:22-31
            ClustersView { width: self.width, height: self.height, pixels: &self.pixels, clusters: &self.clusters,
                           cluster_indices: &self.cluster_indices, clusters_output: &self.clusters_output }
        }
        pub fn take_image(self) -> ColorImage {                                    // container.rs:33-40
            ColorImage { pixels: self.pixels, width: self.width as usize, height: self.height as usize }
        }
    }
    impl ClustersView<'_> {
        pub fn iter(&self) -> impl Iterator<Item = &Cluster> {                     // container.rs:58-64
            self.clusters_output.iter().map(move |i| &self.clusters[i.0 as usize])
        }
        pub fn get_cluster(&self, index: ClusterIndex) -> &Cluster { &self.clusters[index.0 as usize] }   // container.rs:66-68
        pub fn get_cluster_at_point(&self, point: PointI32) -> ClusterIndex {      // container.rs:70-73
            self.get_cluster_at((point.y * self.width as i32 + point.x) as u32)
        }
        pub fn get_cluster_at(&self, index: u32) -> ClusterIndex { self.cluster_indices[index as usize] }   // container.rs:75-77
        pub fn get_pixel(&self, x: i32, y: i32) -> Option<Color> {                 // container.rs:79-87
            if x < 0 || y < 0 { return None; }
            if x as u32 >= self.width { return None; }
            self.get_pixel_at_index((y * self.width as i32 + x) as u32)
        }
        pub fn get_pixel_at_index(&self, index: u32) -> Option<Color> {            // container.rs:89-102
            let i = index as usize * 4;
            if i >= self.pixels.len() { return None; }
            Some(Color::new_rgba(self.pixels[i], self.pixels[i + 1], self.pixels[i + 2], self.pixels[i + 3]))
        }
        /// container.rs:104-116: clusters_output painted in reverse with residue_color.  Host code over the returned vectors,
        /// as in the reference (callers that still hold the device result can use vcb_clusters_to_color_image instead).
        pub fn to_color_image(&self) -> ColorImage {
            let mut image = ColorImage::new_w_h(self.width as usize, self.height as usize);
            self.clusters_output.iter().rev().for_each(|&u| self.get_cluster(u).render_to_color_image(self, &mut image));
            image
        }
    }

    fn to_ffi_config(c: &RunnerConfig) -> ffi::vcb_runner_config {
        ffi::vcb_runner_config {
            diagonal: c.diagonal as i32, hierarchical: c.hierarchical, batch_size: c.batch_size,
            good_min_area: c.good_min_area as u64, good_max_area: c.good_max_area as u64,
            is_same_color_a: c.is_same_color_a, is_same_color_b: c.is_same_color_b, deepen_diff: c.deepen_diff,
            hollow_neighbours: c.hollow_neighbours as u64,
            key_color: [c.key_color.r, c.key_color.g, c.key_color.b, c.key_color.a],
            keying_action: match c.keying_action { KeyingAction::Keep => 0, KeyingAction::Discard => 1 },
        }
    }

    /// The whole computation on the GPU (vcb_runner_run), results copied into the reference's containers.
    fn run_on_device(config: &RunnerConfig, image: ColorImage) -> Clusters {
        let cfg = to_ffi_config(config);
        let (w, h) = (image.width as u32, image.height as u32);
        assert!(image.pixels.len() == (w as usize) * (h as usize) * 4);            // index out of bounds in the reference (builder.rs:190,328)
        with_ctx(|ctx| unsafe {
            let mut hnd = std::ptr::null_mut();
            check(ctx, ffi::vcb_runner_run(ctx, &cfg, image.pixels.as_ptr(), w, h, &mut hnd));
            let mut info = ffi::vcb_clusters_info::default();
            check(ctx, ffi::vcb_clusters_info_get(hnd, &mut info));
            let mut labels = vec![0u32; (w as usize) * (h as usize)];
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
            Clusters { width: w, height: h, pixels: image.pixels, clusters,
                       cluster_indices: labels.into_iter().map(ClusterIndex).collect(),
                       clusters_output: outs.into_iter().map(ClusterIndex).collect() }
        })
    }

    /// builder.rs:48-110 restricted to what the GPU path can honour: the configuration setters work, the closure setters
    /// (arbitrary host code per pixel pair / per cluster) cannot be offloaded and panic.
    pub struct Builder { config: RunnerConfig, image: Option<ColorImage> }
    impl Builder {
        pub fn from(mut self, image: ColorImage) -> Self { self.image = Some(image); self }
        pub fn diagonal(mut self, v: bool) -> Self { self.config.diagonal = v; self }
        pub fn hierarchical(mut self, v: u32) -> Self { self.config.hierarchical = v; self }
        pub fn batch_size(mut self, v: u32) -> Self { self.config.batch_size = v as i32; self }
        pub fn key(mut self, v: Color) -> Self { self.config.key_color = v; self }
        pub fn keying_action(mut self, v: KeyingAction) -> Self { self.config.keying_action = v; self }
        pub fn same(self, _f: impl Fn(Color, Color) -> bool + 'static) -> Self { Self::unsupported("same") }
        pub fn diff(self, _f: impl Fn(Color, Color) -> i32 + 'static) -> Self { Self::unsupported("diff") }
        pub fn deepen(self, _f: impl Fn(&Cluster, &[NeighbourInfo]) -> bool + 'static) -> Self { Self::unsupported("deepen") }
        pub fn hollow(self, _f: impl Fn(&Cluster, &[NeighbourInfo]) -> bool + 'static) -> Self { Self::unsupported("hollow") }
        fn unsupported(name: &str) -> Self {
            panic!("visioncortex-b200: Builder::{} takes a host closure and cannot be offloaded to the GPU; use RunnerConfig (no CPU fallback)", name)
        }
        pub fn run(self) -> Clusters { run_on_device(&self.config, self.image.unwrap()) }          // builder.rs:90-94
        pub fn start(self) -> IncrementalBuilder {                                                 // builder.rs:96-98
            IncrementalBuilder { pending: Some((self.config, self.image.unwrap())), done: None, taken: false }
        }
    }

    /// builder.rs:58-60,112-141.  `tick()` computes everything on the first call (returns false) and reports completion on
    /// the second (returns true); `progress()` is 0 / 100; `view()` exists once the first tick has run.
    pub struct IncrementalBuilder { pending: Option<(RunnerConfig, ColorImage)>, done: Option<Clusters>, taken: bool }
    impl IncrementalBuilder {
        pub fn tick(&mut self) -> bool {
            if let Some((config, image)) = self.pending.take() {
                self.done = Some(run_on_device(&config, image));
                return false;
            }
            true
        }
        pub fn view(&self) -> ClustersView<'_> { self.done.as_ref().expect("view() before the first tick()").view() }
        pub fn result(&mut self) -> Clusters {
            if self.pending.is_some() { self.tick(); }
            self.taken = true;
            self.done.take().unwrap()
        }
        pub fn progress(&self) -> u32 {
            if self.taken { return 0; }                                            // builder.rs:131-135: builder_impl is None
            if self.done.is_some() { 100 } else { 0 }
        }
    }

    pub struct Runner { config: RunnerConfig, image: ColorImage }                 // runner.rs:4-7
    impl Default for Runner {
        fn default() -> Self { Self { config: RunnerConfig::default(), image: ColorImage::new() } }   // runner.rs:41-48
    }
    impl Runner {
        pub fn new(config: RunnerConfig, image: ColorImage) -> Self { Self { config, image } }        // runner.rs:52-57
        pub fn init(&mut self, image: ColorImage) { self.image = image; }                             // runner.rs:59-61
        pub fn builder(self) -> Builder {                                                             // runner.rs:63-98
            assert!(self.config.is_same_color_a < 8);                                                 // runner.rs:78
            Builder { config: self.config, image: Some(self.image) }
        }
        pub fn start(self) -> IncrementalBuilder { self.builder().start() }                           // runner.rs:100-102
        pub fn run(self) -> Clusters { self.builder().run() }                                         // runner.rs:104-106
    }
}

pub mod clusters {
    use super::*;

    #[derive(Default)]
    pub struct Cluster { pub points: Vec<PointI32>, pub rect: BoundingRect }      // clusters.rs:7-11
    #[derive(Default)]
    pub struct Clusters { pub clusters: Vec<Cluster>, pub rect: BoundingRect }    // clusters.rs:14-18

    impl Cluster {
        pub fn iter(&self) -> std::slice::Iter<'_, PointI32> { self.points.iter() }                  // clusters.rs:21-23
        pub fn size(&self) -> usize { self.points.len() }                                             // clusters.rs:30-32
        pub fn to_binary_image(&self) -> BinaryImage {                                                // clusters.rs:34-45
            let mut image = BinaryImage::new_w_h(self.rect.width() as usize, self.rect.height() as usize);
            for p in self.points.iter() {
                image.set_pixel(p.x as usize - self.rect.left as usize, p.y as usize - self.rect.top as usize, true);
            }
            image
        }
    }
    impl Clusters {
        pub fn iter(&self) -> std::slice::Iter<'_, Cluster> { self.clusters.iter() }                 // clusters.rs:239-241
        pub fn len(&self) -> usize { self.clusters.len() }
        pub fn is_empty(&self) -> bool { self.clusters.is_empty() }
        pub fn get_cluster(&self, index: usize) -> &Cluster { &self.clusters[index] }
    }

    impl BinaryImage {
        /// clusters.rs:270-349 — runs on the GPU (vcb_binary_to_clusters); `panic!("overflow")` mirrors clusters.rs:322-324
        pub fn to_clusters(&self, diagonal: bool) -> Clusters {
            with_ctx(|ctx| unsafe {
                let mut hnd = std::ptr::null_mut();
                check(ctx, ffi::vcb_binary_to_clusters(ctx, self.pixels.storage().as_ptr(), self.width as u32, self.height as u32,
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
