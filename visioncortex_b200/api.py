"""Host-side mirror of the reference's Rust API for the clustering path, over the C ABI (include/vcb200.h).

Same names, argument meaning and error behaviour as visioncortex v0.8.10 (paths relative to /root/reference/src):
  ColorImage / BinaryImage / Color                 image/format.rs:10-33,43-77,268-327; color.rs:8-14
  color_clusters::{Runner, RunnerConfig, KeyingAction, HIERARCHICAL_MAX, ZERO, Builder, IncrementalBuilder,
                   Clusters, ClustersView, Cluster, ClusterIndex, NeighbourInfo}       color_clusters/*.rs
  clusters::{Clusters, Cluster}, BinaryImage::to_clusters                              clusters.rs:7-18,270-349
`Runner.run()` and `BinaryImage.to_clusters()` run on the GPU; there is no CPU fallback.  The query methods of
`ClustersView` / `Cluster` (perimeter, neighbours, to_image ...) are plain host code over the vectors the run returned,
exactly as in the reference.  The Rust replacement crate (rust/src/lib.rs) has the same structure; this module is its
executable twin, used by the tests.

Documented deviations: the custom-closure setters of `Builder` (`same/diff/deepen/hollow`,
builder.rs:106-109) cannot be offloaded and raise; `IncrementalBuilder.tick()` performs the whole computation on the
first call and returns True on the second; `progress()` is 0 before and 100 after; intermediate `view()` states are not
reproduced (the final `result()` is bit-exact).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _ffi
from ._ffi import pack_bits
from .runtime import Context, make_config

HIERARCHICAL_MAX = 0xFFFFFFFF          # color_clusters/builder.rs:46
ZERO = 0                               # color_clusters/builder.rs:45 (ClusterIndex(0))


class KeyingAction:                    # color_clusters/builder.rs:6-11
    Keep = 0
    Discard = 1


_default_ctx = None


def default_context() -> Context:
    """One context on device 0, opened on first use (the Rust shim does the same behind a Mutex)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def set_default_context(ctx: Context):
    global _default_ctx
    _default_ctx = ctx


@dataclass(frozen=True)
class Color:                           # color.rs:8-14
    r: int = 0
    g: int = 0
    b: int = 0
    a: int = 0

    def tuple(self):
        return (self.r, self.g, self.b, self.a)


@dataclass
class PointI32:                        # point.rs:381
    x: int = 0
    y: int = 0


@dataclass
class BoundingRect:                    # bound.rs:15-21
    left: int = 0
    top: int = 0
    right: int = 0
    bottom: int = 0

    def width(self):
        return self.right - self.left

    def height(self):
        return self.bottom - self.top

    def is_empty(self):
        return self.width() == 0 and self.height() == 0


@dataclass
class ColorSum:                        # color.rs:40-47
    r: int = 0
    g: int = 0
    b: int = 0
    a: int = 0
    counter: int = 0

    def average(self) -> Color:        # color.rs:295-302 (`as u8`); division by zero panics in the reference
        n = self.counter
        if n == 0:
            raise ZeroDivisionError("attempt to divide by zero")
        return Color((self.r // n) & 255, (self.g // n) & 255, (self.b // n) & 255, (self.a // n) & 255)


class ColorImage:                      # image/format.rs:29-33
    def __init__(self, pixels=None, width: int = 0, height: int = 0):
        self.width, self.height = int(width), int(height)
        self.pixels = np.zeros(self.width * self.height * 4, dtype=np.uint8) if pixels is None else np.ascontiguousarray(pixels, dtype=np.uint8).reshape(-1)

    @staticmethod
    def new_w_h(width: int, height: int) -> "ColorImage":       # format.rs:273-279
        return ColorImage(None, width, height)

    @staticmethod
    def from_array(rgba: np.ndarray) -> "ColorImage":
        h, w = rgba.shape[:2]
        return ColorImage(rgba, w, h)

    def get_pixel(self, x: int, y: int) -> Color:               # format.rs:289-292,306-314 (panics out of bounds)
        i = (y * self.width + x) * 4
        if not (0 <= i and i + 3 < len(self.pixels)):
            raise IndexError("index out of bounds")
        return Color(*[int(v) for v in self.pixels[i:i + 4]])

    def set_pixel(self, x: int, y: int, color: Color):          # format.rs:316-327
        i = (y * self.width + x) * 4
        if not (0 <= i and i + 3 < len(self.pixels)):
            raise IndexError("index out of bounds")
        self.pixels[i:i + 4] = color.tuple()

    def array(self) -> np.ndarray:
        return self.pixels.reshape(self.height, self.width, 4)


class BinaryImage:                     # image/format.rs:10-14; `pixels` = bit_vec::BitVec (one bool per pixel here)
    def __init__(self, width: int = 0, height: int = 0, pixels=None):
        self.width, self.height = int(width), int(height)
        self.pixels = np.zeros(self.width * self.height, dtype=bool) if pixels is None else np.ascontiguousarray(pixels, dtype=bool).reshape(-1)

    @staticmethod
    def new_w_h(width: int, height: int) -> "BinaryImage":      # format.rs:44-50
        return BinaryImage(width, height)

    @staticmethod
    def from_string(string: str) -> "BinaryImage":              # format.rs:153-169 ('*' = set, rows by '\n')
        rows = string.split("\n")
        width = len(rows[0]) if rows else 0
        img = BinaryImage(width, len(rows))
        for y, row in enumerate(rows):
            for x, ch in enumerate(row):
                img.set_pixel(x, y, ch == "*")
        return img

    def get_pixel(self, x: int, y: int) -> bool:                # format.rs:58-61 (panics out of bounds)
        return bool(self.pixels[y * self.width + x])

    def get_pixel_safe(self, x: int, y: int) -> bool:           # format.rs:66-72
        return 0 <= x < self.width and 0 <= y < self.height and self.get_pixel(x, y)

    def set_pixel(self, x: int, y: int, v: bool):               # format.rs:74-77
        self.pixels[y * self.width + x] = v

    def storage(self) -> np.ndarray:
        """bit_vec::BitVec::storage(): u32 blocks, pixel i at block i / 32, bit i % 32."""
        return pack_bits(self.pixels.reshape(self.height, self.width)) if self.pixels.size else np.zeros(0, dtype=np.uint32)

    def to_clusters(self, diagonal: bool, ctx: Context | None = None) -> "BinaryClusters":      # clusters.rs:270-349, on the GPU
        ctx = ctx or default_context()
        try:
            res = ctx.to_clusters(self.pixels.reshape(self.height, self.width), diagonal)
        except _ffi.VcbError as e:
            if e.status == _ffi.VCB_ERR_LABEL_OVERFLOW:
                raise OverflowError("overflow") from e          # panic!("overflow"), clusters.rs:322-324
            raise
        out = BinaryClusters(rect=BoundingRect(*[int(v) for v in res["rect"]]))
        for r in res["records"]:
            o, n = int(r["points_off"]), int(r["points_len"])
            out.clusters.append(BinaryCluster(points=[PointI32(int(x), int(y)) for x, y in res["points"][o:o + n]],
                                              rect=BoundingRect(*[int(v) for v in r["rect"]])))
        return out


@dataclass
class BinaryCluster:                   # clusters.rs:7-11
    points: list = field(default_factory=list)
    rect: BoundingRect = field(default_factory=BoundingRect)

    def size(self):
        return len(self.points)

    def to_binary_image(self) -> BinaryImage:                   # clusters.rs:34-45
        img = BinaryImage.new_w_h(self.rect.width(), self.rect.height())
        for p in self.points:
            img.set_pixel(p.x - self.rect.left, p.y - self.rect.top, True)
        return img


@dataclass
class BinaryClusters:                  # clusters.rs:14-18
    clusters: list = field(default_factory=list)
    rect: BoundingRect = field(default_factory=BoundingRect)

    def __len__(self):
        return len(self.clusters)

    def __iter__(self):
        return iter(self.clusters)


@dataclass
class RunnerConfig:                    # color_clusters/runner.rs:9-39
    diagonal: bool = False
    hierarchical: int = HIERARCHICAL_MAX
    batch_size: int = 25600
    good_min_area: int = 16
    good_max_area: int = 256 * 256
    is_same_color_a: int = 4
    is_same_color_b: int = 1
    deepen_diff: int = 64
    hollow_neighbours: int = 1
    key_color: Color = Color()
    keying_action: int = KeyingAction.Keep

    def to_c(self):
        return make_config(diagonal=int(self.diagonal), hierarchical=self.hierarchical, batch_size=self.batch_size,
                           good_min_area=self.good_min_area, good_max_area=self.good_max_area,
                           is_same_color_a=self.is_same_color_a, is_same_color_b=self.is_same_color_b,
                           deepen_diff=self.deepen_diff, hollow_neighbours=self.hollow_neighbours,
                           key_color=self.key_color.tuple(), keying_action=self.keying_action)


@dataclass
class NeighbourInfo:                   # color_clusters/builder.rs:34-37
    index: int
    diff: int


class Cluster:                         # color_clusters/cluster.rs:7-17
    def __init__(self, rec, indices: np.ndarray, holes: np.ndarray):
        self.indices = indices
        self.holes = holes
        self.num_holes = int(rec["num_holes"])
        self.depth = int(rec["depth"])
        self.sum = ColorSum(*[int(v) for v in rec["sum"]])
        self.residue_sum = ColorSum(*[int(v) for v in rec["residue_sum"]])
        self.rect = BoundingRect(*[int(v) for v in rec["rect"]])
        self.merged_into = int(rec["merged_into"])

    def area(self) -> int:                                      # cluster.rs:30-32
        return len(self.indices)

    def iter(self):                                             # cluster.rs:34-36
        return iter(self.indices)

    def color(self) -> Color:                                   # cluster.rs:38-40
        return self.sum.average()

    def residue_color(self) -> Color:                           # cluster.rs:42-44
        return self.residue_sum.average()

    def to_image_with_hole(self, parent_width: int, hole: bool) -> BinaryImage:     # cluster.rs:60-80
        img = BinaryImage.new_w_h(self.rect.width(), self.rect.height())
        for lst, v in ((self.indices, True),) + (((self.holes, False),) if hole else ()):
            if len(lst):
                i = lst.astype(np.int64)
                x = i % parent_width - self.rect.left
                y = i // parent_width - self.rect.top
                img.pixels[y * img.width + x] = v
        return img

    def to_image(self, parent: "ClustersView") -> BinaryImage:                      # cluster.rs:53-55
        return self.to_image_with_hole(parent.width, True)

    def perimeter(self, parent: "ClustersView") -> int:         # cluster.rs:46-48 -> shape/geometry.rs:52-75
        img = self.to_image(parent)
        m = np.zeros((img.height + 2, img.width + 2), dtype=bool)
        m[1:-1, 1:-1] = img.pixels.reshape(img.height, img.width)
        c = m[1:-1, 1:-1]
        return int((c & ~(m[1:-1, :-2] & m[1:-1, 2:] & m[:-2, 1:-1] & m[2:, 1:-1])).sum())

    def render_to_binary_image(self, parent: "ClustersView", image: BinaryImage):   # cluster.rs:82-88
        i = self.indices.astype(np.int64)
        image.pixels[(i // parent.width) * image.width + i % parent.width] = True

    def render_to_color_image_with_color(self, parent: "ClustersView", image: ColorImage, color: Color):   # cluster.rs:95-101
        i = self.indices.astype(np.int64)
        k = ((i // parent.width) * image.width + i % parent.width) * 4
        for ch, v in enumerate(color.tuple()):
            image.pixels[k + ch] = v

    def render_to_color_image(self, parent: "ClustersView", image: ColorImage):     # cluster.rs:90-93
        self.render_to_color_image_with_color(parent, image, self.residue_color())

    def neighbours(self, parent: "ClustersView") -> list:       # cluster.rs:132-165
        w, h = parent.width, parent.height
        ci = parent.cluster_indices
        myself = int(ci[self.indices[0]])
        i = self.indices.astype(np.int64)
        x, y = i % w, i // w
        found = []
        for ok, off in ((y > 0, -w), (y < h - 1, w), (x > 0, -1), (x < w - 1, 1)):
            found.append(ci[(i + off)[ok]])
        lab = np.unique(np.concatenate(found)) if found else np.zeros(0, dtype=np.uint32)
        return [int(v) for v in lab if v != ZERO and v != myself]


class ClustersView:                    # color_clusters/container.rs:43-117
    def __init__(self, owner: "Clusters"):
        self._o = owner
        self.width, self.height = owner.width, owner.height
        self.pixels = owner._pixels
        self.clusters = owner._clusters
        self.cluster_indices = owner._cluster_indices
        self.clusters_output = owner._clusters_output

    def iter(self):                                             # container.rs:58-64
        return (self.clusters[int(i)] for i in self.clusters_output)

    def get_cluster(self, index: int) -> Cluster:               # container.rs:66-68
        return self.clusters[int(index)]

    def get_cluster_at(self, index: int) -> int:                # container.rs:75-77
        return int(self.cluster_indices[index])

    def get_cluster_at_point(self, point: PointI32) -> int:     # container.rs:70-73
        return self.get_cluster_at(point.y * self.width + point.x)

    def get_pixel_at_index(self, index: int):                   # container.rs:89-102
        i = index * 4
        if i >= len(self.pixels):
            return None
        return Color(*[int(v) for v in self.pixels[i:i + 4]])

    def get_pixel(self, x: int, y: int):                        # container.rs:79-87
        if x < 0 or y < 0 or x >= self.width:
            return None
        return self.get_pixel_at_index(y * self.width + x)

    def to_color_image(self) -> ColorImage:                     # container.rs:104-116, on the GPU (vcb_clusters_to_color_image)
        return self._o._to_color_image()


class Clusters:                        # color_clusters/container.rs:4-41
    def __init__(self, ctx: Context, handle, pixels: np.ndarray, width: int, height: int):
        self._ctx, self._handle = ctx, handle
        self.width, self.height = width, height
        self._pixels = pixels
        res = ctx.lib.read_clusters(handle, pools=True, free=False, ctx=ctx.handle)
        self._cluster_indices = res["cluster_indices"]
        self._clusters_output = res["clusters_output"]
        rec, ip, hp = res["records"], res["indices_pool"], res["holes_pool"]
        self._clusters = [Cluster(r, ip[int(r["indices_off"]):int(r["indices_off"]) + int(r["indices_len"])],
                                  hp[int(r["holes_off"]):int(r["holes_off"]) + int(r["holes_len"])]) for r in rec]

    def __del__(self):
        try:
            if self._handle:
                self._ctx.lib.fn("clusters_free")(self._handle)
                self._handle = None
        except Exception:
            pass

    def output_len(self) -> int:                                # container.rs:18-20
        return len(self._clusters_output)

    # ---- device-resident queries (include/vcb200.h: vcb_clusters_perimeters / _neighbours / _to_image): computed on the GPU from
    # the labels, the merge forest and the records the result keeps in HBM -- the host pools above are not involved
    def device_perimeters(self) -> np.ndarray:                  # Cluster::perimeter (cluster.rs:46-48) of every cluster, one pass
        out = np.zeros(len(self._clusters), dtype=np.uint32)
        lib = self._ctx.lib
        lib.check(lib.fn("clusters_perimeters")(self._handle, out.ctypes.data_as(C.c_void_p)), self._ctx.handle)
        return out

    def device_neighbours(self, index: int) -> list:            # Cluster::neighbours (cluster.rs:132-165)
        lib = self._ctx.lib
        cap = 64
        while True:
            out = np.zeros(cap, dtype=np.uint32)
            n = C.c_size_t(0)
            st = lib.fn("clusters_neighbours")(self._handle, int(index), out.ctypes.data_as(C.c_void_p), cap, C.byref(n))
            if st == 6 and n.value > cap:                       # VCB_ERR_NOMEM: the list is longer than the buffer
                cap = int(n.value)
                continue
            lib.check(st, self._ctx.handle)
            return [int(v) for v in out[:n.value]]

    def device_components(self, hole: bool = True) -> list:
        """The clustering step of Cluster::to_compound_path (cluster.rs:108-128) for every output cluster, batched on the device
        (vcb_clusters_components): element k = the binary clusters of `clusters_output[k].to_image_with_hole(width, hole)
        .to_clusters(false)` as (rect tuple, points (n, 2) int32 array), relative to the cluster's image."""
        from ._ffi import BIN_CLUSTER_REC_DTYPE

        lib = self._ctx.lib
        hnd = C.c_void_p()
        lib.check(lib.fn("clusters_components")(self._handle, int(hole), C.byref(hnd)), self._ctx.handle)
        try:
            nout, ncomp, npts = C.c_uint32(), C.c_uint64(), C.c_uint64()
            lib.check(lib.fn("components_info")(hnd, C.byref(nout), C.byref(ncomp), C.byref(npts)), self._ctx.handle)
            off = np.zeros(nout.value + 1, dtype=np.uint32)
            rec = np.zeros(max(1, ncomp.value), dtype=BIN_CLUSTER_REC_DTYPE)
            xy = np.zeros((max(1, npts.value), 2), dtype=np.int32)
            lib.check(lib.fn("components_get")(hnd, off.ctypes.data_as(C.c_void_p), rec.ctypes.data_as(C.c_void_p), xy.ctypes.data_as(C.c_void_p)), self._ctx.handle)
        finally:
            lib.fn("components_free")(hnd)
        out = []
        for k in range(nout.value):
            comps = []
            for r in rec[off[k]:off[k + 1]]:
                o, n = int(r["points_off"]), int(r["points_len"])
                comps.append((tuple(int(v) for v in r["rect"]), xy[o:o + n].copy()))
            out.append(comps)
        return out

    def device_to_image(self, index: int, hole: bool = True) -> BinaryImage:   # Cluster::to_image_with_hole (cluster.rs:60-80)
        lib = self._ctx.lib
        rect = (C.c_int32 * 4)()
        r = self._clusters[int(index)].rect
        nwords = (r.width() * r.height() + 31) // 32
        bits = np.zeros(max(1, nwords), dtype=np.uint32)
        lib.check(lib.fn("clusters_to_image")(self._handle, int(index), int(hole), bits.ctypes.data_as(C.c_void_p), nwords, C.byref(rect)), self._ctx.handle)
        img = BinaryImage.new_w_h(rect[2] - rect[0], rect[3] - rect[1])
        n = img.width * img.height
        if n:
            img.pixels[:] = np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)
        return img

    def view(self) -> ClustersView:                             # container.rs:22-31
        return ClustersView(self)

    def take_image(self) -> ColorImage:                         # container.rs:33-40
        return ColorImage(self._pixels, self.width, self.height)

    def _to_color_image(self) -> ColorImage:
        out = ColorImage.new_w_h(self.width, self.height)
        if self.width * self.height:
            self._ctx.lib.check(self._ctx.lib.fn("clusters_to_color_image")(self._handle, out.pixels.ctypes.data_as(C.c_void_p)), self._ctx.handle)
        return out


class ClosureNotOffloadable(TypeError):
    """Builder::{same, diff, deepen, hollow} take arbitrary host closures in the reference; the GPU path only implements the
    closures Runner::builder binds (runner.rs:87-97)."""


class IncrementalBuilder:              # color_clusters/builder.rs:58-60,112-141
    def __init__(self, run):
        self._run, self._clusters, self._ticks, self._taken = run, None, 0, False

    def tick(self) -> bool:
        """builder.rs:119-121.  The whole computation runs on the GPU on the first call; the second call reports completion."""
        if self._clusters is None:
            self._clusters = self._run()
            return False
        return True

    def progress(self) -> int:                                  # builder.rs:131-140: 0 once the result was taken, else 0 / 100
        if self._taken:
            return 0
        return 0 if self._clusters is None else 100

    def view(self) -> ClustersView:                             # builder.rs:123-125 (final state only)
        if self._clusters is None:
            raise RuntimeError("view() before the first tick(): intermediate states are not reproduced")
        return self._clusters.view()

    def result(self) -> Clusters:                               # builder.rs:127-129
        if self._clusters is None:
            self._clusters = self._run()
        self._taken = True
        return self._clusters


class Builder:                         # color_clusters/builder.rs:48-110 restricted to what the GPU path can honour
    def __init__(self, runner: "Runner"):
        self._runner = runner

    def _unsupported(self, *_a, **_k):
        raise ClosureNotOffloadable("custom same/diff/deepen/hollow closures (builder.rs:106-109) cannot be offloaded to the GPU; "
                                    "use RunnerConfig (there is no CPU fallback)")

    same = diff = deepen = hollow = _unsupported

    def diagonal(self, v: bool):
        self._runner.config.diagonal = bool(v)
        return self

    def hierarchical(self, v: int):
        self._runner.config.hierarchical = int(v)
        return self

    def batch_size(self, v: int):
        self._runner.config.batch_size = int(v)
        return self

    def key(self, v: Color):
        self._runner.config.key_color = v
        return self

    def keying_action(self, v: int):
        self._runner.config.keying_action = v
        return self

    def run(self) -> Clusters:                                  # builder.rs:90-94
        return self._runner._run()

    def start(self) -> IncrementalBuilder:                      # builder.rs:96-98
        return IncrementalBuilder(self._runner._run)


class Runner:                          # color_clusters/runner.rs:4-7,52-106
    def __init__(self, config: RunnerConfig | None = None, image: ColorImage | None = None, ctx: Context | None = None):
        self.config = config or RunnerConfig()
        self.image = image or ColorImage()
        self._ctx = ctx

    @staticmethod
    def new(config: RunnerConfig, image: ColorImage) -> "Runner":
        return Runner(config, image)

    def init(self, image: ColorImage):                          # runner.rs:59-61
        self.image = image

    def builder(self) -> Builder:                               # runner.rs:63-98
        assert self.config.is_same_color_a < 8                  # runner.rs:78
        return Builder(self)

    def start(self) -> IncrementalBuilder:                      # runner.rs:100-102
        return self.builder().start()

    def run(self) -> Clusters:                                  # runner.rs:104-106
        return self.builder().run()

    def _run(self) -> Clusters:
        ctx = self._ctx or default_context()
        img = self.image
        handle = ctx.runner_run_handle(img.pixels.reshape(img.height, img.width, 4), self.config.to_c())
        return Clusters(ctx, handle, img.pixels, img.width, img.height)
