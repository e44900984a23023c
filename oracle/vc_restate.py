"""vc_restate.py — SECOND, INDEPENDENT CPU restatement of the reference's clustering path, in pure Python.

TEST INFRASTRUCTURE ONLY.  Nothing under visioncortex_b200/ imports this file; only tests/ does.

Why it exists: `color_clusters` has no unit tests in the reference, so the C++ oracle
(oracle/vc_oracle.cpp) is pinned only by being a careful reading of the Rust.  This file is a second reading, written
from the Rust sources directly (not from vc_oracle.cpp), in a different language, with different data structures, by
walking the Rust statement by statement.  tests/test_two_oracles.py fuzzes the two against each other on thousands of
small images including the order-dependent quirk families (diagonal = true with the stale `cluster_upleft` read,
unclean key colours).  Two independent readings that agree bit for bit are the strongest pin available in an image
without rustc; it is still not the Rust binary.

Semantics fixed explicitly: Rust *release* profile (u32 arithmetic wraps, no overflow panics), `sort_by_key` is
stable, `HashSet` iteration order is irrelevant (the set is sorted before use).  Situations in which the Rust would
panic (division by zero in ColorSum::average, out-of-range BitVec::set, `unwrap` on a failed binary search) raise
`ReferencePanic` here.

Every function cites the reference lines it follows (paths relative to /root/reference/src).
"""
from bisect import bisect_left

U32 = 0xFFFFFFFF
HIERARCHICAL_MAX = 0xFFFFFFFF          # color_clusters/builder.rs:46
KEEP, DISCARD = 0, 1                   # color_clusters/builder.rs:6-11


class ReferencePanic(Exception):
    pass


# ------------------------------------------------------------------------------------------------ color.rs / bound.rs
class ColorSum:                        # color.rs:40-47
    __slots__ = ("r", "g", "b", "a", "counter")

    def __init__(self):
        self.r = self.g = self.b = self.a = self.counter = 0

    def add(self, c):                  # color.rs:261-267
        self.r = (self.r + c[0]) & U32
        self.g = (self.g + c[1]) & U32
        self.b = (self.b + c[2]) & U32
        self.a = (self.a + c[3]) & U32
        self.counter = (self.counter + 1) & U32

    def merge(self, o):                # color.rs:278-284
        self.r = (self.r + o.r) & U32
        self.g = (self.g + o.g) & U32
        self.b = (self.b + o.b) & U32
        self.a = (self.a + o.a) & U32
        self.counter = (self.counter + o.counter) & U32

    def average(self):                 # color.rs:295-302 (`as u8` truncates)
        if self.counter == 0:
            raise ReferencePanic("attempt to divide by zero (ColorSum::average)")
        n = self.counter
        return ((self.r // n) & 255, (self.g // n) & 255, (self.b // n) & 255, (self.a // n) & 255)

    def clear(self):                   # color.rs:304-310
        self.r = self.g = self.b = self.a = self.counter = 0

    def copy(self):
        o = ColorSum()
        o.r, o.g, o.b, o.a, o.counter = self.r, self.g, self.b, self.a, self.counter
        return o

    def tup(self):
        return (self.r, self.g, self.b, self.a, self.counter)


class Rect:                            # bound.rs:15-21
    __slots__ = ("left", "top", "right", "bottom")

    def __init__(self):
        self.left = self.top = self.right = self.bottom = 0

    def width(self):                   # bound.rs:86-88
        return self.right - self.left

    def height(self):                  # bound.rs:91-93
        return self.bottom - self.top

    def is_empty(self):                # bound.rs:101-103
        return self.width() == 0 and self.height() == 0

    def add_x_y(self, x, y):           # bound.rs:172-190
        if self.is_empty():
            self.left, self.right, self.top, self.bottom = x, x + 1, y, y + 1
            return
        if x < self.left:
            self.left = x
        elif x + 1 > self.right:
            self.right = x + 1
        if y < self.top:
            self.top = y
        elif y + 1 > self.bottom:
            self.bottom = y + 1

    def merge(self, o):                # bound.rs:204-219
        if o.is_empty():
            return
        if self.is_empty():
            self.left, self.right, self.top, self.bottom = o.left, o.right, o.top, o.bottom
            return
        self.left = min(self.left, o.left)
        self.right = max(self.right, o.right)
        self.top = min(self.top, o.top)
        self.bottom = max(self.bottom, o.bottom)

    def clear(self):                   # bound.rs:222-227
        self.left = self.top = self.right = self.bottom = 0

    def copy(self):
        o = Rect()
        o.left, o.top, o.right, o.bottom = self.left, self.top, self.right, self.bottom
        return o

    def tup(self):
        return (self.left, self.top, self.right, self.bottom)


# ------------------------------------------------------------------------------------------------ cluster.rs
class Cluster:                         # color_clusters/cluster.rs:7-17
    __slots__ = ("indices", "holes", "num_holes", "depth", "sum", "residue_sum", "rect", "merged_into")

    def __init__(self):
        self.indices = []
        self.holes = []
        self.num_holes = 0
        self.depth = 0
        self.sum = ColorSum()
        self.residue_sum = ColorSum()
        self.rect = Rect()
        self.merged_into = 0

    def add(self, i, c, x, y):         # cluster.rs:24-28
        self.indices.append(i)
        self.sum.add(c)
        self.rect.add_x_y(x, y)

    def area(self):                    # cluster.rs:30-32
        return len(self.indices)


def color_diff(a, b):                  # color_clusters/runner.rs:110-114
    return abs(a[0] - b[0]) + abs(a[1] - b[1]) + abs(a[2] - b[2])


def color_same(a, b, shift, thres):    # color_clusters/runner.rs:116-129 (alpha ignored)
    return (abs((a[0] >> shift) - (b[0] >> shift)) <= thres and abs((a[1] >> shift) - (b[1] >> shift)) <= thres
            and abs((a[2] >> shift) - (b[2] >> shift)) <= thres)


def to_image_with_hole(cl, parent_width, hole=True):     # cluster.rs:60-80 (BitVec of rect.width * rect.height bits)
    w, h = cl.rect.width(), cl.rect.height()
    if w < 0 or h < 0:
        raise ReferencePanic("negative rect size")
    bits = bytearray(w * h)

    def setp(i, v):
        x = (i % parent_width) - cl.rect.left
        y = (i // parent_width) - cl.rect.top
        if x < 0 or y < 0:                               # `as usize` of a negative value: a huge index, BitVec::set panics
            raise ReferencePanic("pixel outside the cluster rect")
        k = y * w + x                                    # image/format.rs:74-77
        if k >= len(bits):
            raise ReferencePanic("BitVec::set out of bounds")
        bits[k] = v

    for i in cl.indices:
        setp(i, 1)
    if hole:
        for i in cl.holes:
            setp(i, 0)
    return bits, w, h


def image_boundary_count(bits, w, h):  # shape/geometry.rs:52-75 with transpose = false; only the count is used
    def safe(x, y):                    # image/format.rs:66-72
        return 0 <= x < w and 0 <= y < h and bits[y * w + x]

    n = 0
    for y in range(h):
        for x in range(w):
            if bits[y * w + x] and (not safe(x - 1, y) or not safe(x + 1, y) or not safe(x, y - 1) or not safe(x, y + 1)):
                n += 1
    return n


# ------------------------------------------------------------------------------------------------ builder.rs
class Config:
    """RunnerConfig (color_clusters/runner.rs:9-39)."""

    def __init__(self, **kw):
        self.diagonal = False
        self.hierarchical = HIERARCHICAL_MAX
        self.batch_size = 25600
        self.good_min_area = 16
        self.good_max_area = 256 * 256
        self.is_same_color_a = 4
        self.is_same_color_b = 1
        self.deepen_diff = 64
        self.hollow_neighbours = 1
        self.key_color = (0, 0, 0, 0)
        self.keying_action = KEEP
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(k)
            setattr(self, k, v)


class Builder:
    """BuilderImpl (color_clusters/builder.rs:148-569) bound to the Runner's closures (runner.rs:63-98)."""

    def __init__(self, cfg, pixels, width, height):
        assert cfg.is_same_color_a < 8                   # runner.rs:78
        self.cfg = cfg
        self.width, self.height = width, height
        self.pixels = pixels                             # list of (r, g, b, a), row-major
        self.clusters = [Cluster()]                      # builder.rs:189
        self.cluster_indices = [0] * len(pixels)         # builder.rs:190
        self.cluster_areas = []                          # [area, count] pairs
        self.clusters_output = []
        self.next_index = 1                              # builder.rs:195

    # builder.rs:541-569
    def same(self, a, b):
        if a is None or b is None:
            return False
        return color_same(a, b, self.cfg.is_same_color_a, self.cfg.is_same_color_b)

    def pixel_at(self, x, y):
        if x < 0 or y < 0:
            return None
        i = y * self.width + x
        return self.pixels[i] if i < len(self.pixels) else None

    def combine_clusters(self, frm, to):                 # builder.rs:526-539
        cf, ct = self.clusters[frm], self.clusters[to]
        for i in cf.indices:
            self.cluster_indices[i] = to
        moved, cf.indices = cf.indices, []               # std::mem::take
        ct.indices.extend(moved)                         # (frm == to would simply restore the list, as in the Rust)
        s, r = cf.sum.copy(), cf.rect.copy()
        ct.sum.merge(s)
        ct.rect.merge(r)
        cf.sum.clear()
        cf.rect.clear()

    def combine_clusters_clone(self, frm, to):           # builder.rs:514-524
        cf = self.clusters[frm]
        s, r, idx = cf.sum.copy(), cf.rect.copy(), list(cf.indices)
        self.combine_clusters(frm, to)
        cf.sum, cf.rect, cf.indices = s, r, idx

    def merge_cluster_into(self, frm, to, deepen, hollow):   # builder.rs:495-512
        if not deepen:
            self.clusters[to].residue_sum.merge(self.clusters[frm].residue_sum)
            self.combine_clusters(frm, to)
        else:
            self.combine_clusters_clone(frm, to)
            if hollow:
                self.clusters[to].holes.extend(self.clusters[frm].indices)
                self.clusters[to].num_holes += 1
            self.clusters[frm].merged_into = to
            self.clusters[to].depth += 1

    def stage_1(self):                                   # builder.rs:273-364 (batch_size only slices the loop)
        cfg = self.cfg
        diagonal = bool(cfg.diagonal)
        key = tuple(cfg.key_color)
        has_key = key != (0, 0, 0, 0)
        W = self.width
        ci = self.cluster_indices
        for i in range(len(ci)):
            x, y = i % W, i // W
            color, up, left, upleft = self.pixel_at(x, y), self.pixel_at(x, y - 1), self.pixel_at(x - 1, y), self.pixel_at(x - 1, y - 1)
            cluster_up = ci[W * (y - 1) + x] if y > 0 else 0
            cluster_left = ci[W * y + x - 1] if x > 0 else 0
            cluster_upleft = ci[W * (y - 1) + x - 1] if (x > 0 and y > 0) else 0
            if cluster_left != cluster_up and self.same(left, up) and (diagonal or (self.same(color, left) and self.same(color, up))):
                if self.clusters[cluster_left].area() <= self.clusters[cluster_up].area():
                    self.combine_clusters(cluster_left, cluster_up)
                    if cluster_left == ((self.next_index - 1) & U32) and self.next_index == len(self.clusters):
                        self.next_index = (self.next_index - 1) & U32
                    cluster_left = cluster_up
                else:
                    self.combine_clusters(cluster_up, cluster_left)
                    cluster_up = cluster_left
            c = color
            if has_key and c == key:
                if cfg.keying_action == KEEP:
                    self.clusters[0].add(i, c, x, y)
            elif self.same(color, up) and self.same(color, upleft):
                ci[i] = cluster_up
                self.clusters[cluster_up].add(i, c, x, y)
            elif self.same(color, left) and self.same(color, upleft):
                ci[i] = cluster_left
                self.clusters[cluster_left].add(i, c, x, y)
            elif diagonal and self.same(color, upleft):
                ci[i] = cluster_upleft
                self.clusters[cluster_upleft].add(i, c, x, y)
            else:
                nc = Cluster()
                nc.add(i, c, x, y)
                if self.next_index < len(self.clusters):
                    self.clusters[self.next_index] = nc
                else:
                    self.clusters.append(nc)
                ci[i] = self.next_index
                self.next_index = (self.next_index + 1) & U32
        self.prepare_stage_2()

    def stage_1_output(self):                            # builder.rs:366-377
        out = [(idx, cl.area()) for idx, cl in enumerate(self.clusters) if cl.area() > 0]
        out.sort(key=lambda c: c[1] * 65535 + c[0])      # stable
        self.clusters_output.extend(c[0] for c in out)

    def prepare_stage_2(self):                           # builder.rs:379-403
        counts = {}
        for cl in self.clusters:
            cl.residue_sum = cl.sum.copy()
            a = cl.area()
            if a > 0:
                counts[a] = counts.get(a, 0) + 1
        self.cluster_areas = [[a, n] for a, n in sorted(counts.items())]

    def neighbours_internal(self, cl):                   # cluster.rs:132-171
        W, H, ci = self.width, self.height, self.cluster_indices
        myself = ci[cl.indices[0]]
        found = set()
        for i in cl.indices:
            x, y = i % W, i // W
            for idx in (ci[W * (y - 1) + x] if y > 0 else 0, ci[W * (y + 1) + x] if y < H - 1 else 0,
                        ci[W * y + x - 1] if x > 0 else 0, ci[W * y + x + 1] if x < W - 1 else 0):
                if idx != 0 and idx != myself:
                    found.add(idx)
        return sorted(found)

    def patch_good(self, cl):                            # runner.rs:131-146
        cfg = self.cfg
        if cfg.good_min_area < cl.area() < cfg.good_max_area:
            if cfg.good_min_area == 0:
                return True
            bits, w, h = to_image_with_hole(cl, self.width, True)      # cluster.rs:49-58
            if image_boundary_count(bits, w, h) < cl.area():
                return True
        return False

    def stage_2_level(self, it):
        """One call of stage_2 (builder.rs:405-493) at self.iteration == it; returns (next iteration, finished)."""
        cfg = self.cfg
        areas = self.cluster_areas
        if not areas:
            return it, True
        if areas[it][1] == 0:
            it += 1
            return it, it == len(areas)
        cur_area = areas[it][0]
        can_discard = cfg.keying_action == DISCARD and tuple(cfg.key_color) != (0, 0, 0, 0)
        for index in range(len(self.clusters)):
            my = self.clusters[index]
            if my.area() != cur_area:
                continue
            if cur_area > cfg.hierarchical:
                self.clusters_output.append(index)
                continue
            mycolor = my.sum.average()
            infos = [(o, color_diff(mycolor, self.clusters[o].sum.average())) for o in self.neighbours_internal(my)]
            if not infos:
                if it == ((len(areas) - 1) & U32) or can_discard:
                    self.clusters_output.append(index)
                continue
            infos.sort(key=lambda t: t[1] * 65535 + t[0])    # stable sort_by_key
            target = infos[0][0]
            if cfg.hierarchical == HIERARCHICAL_MAX:
                deepen = self.patch_good(my) and infos[0][1] > cfg.deepen_diff      # runner.rs:91-94
            else:
                deepen = False
            hollow = len(infos) <= cfg.hollow_neighbours                            # runner.rs:95-97
            if deepen:
                self.clusters_output.append(index)
            tarea = self.clusters[target].area()
            pos = bisect_left(areas, tarea, key=lambda a: a[0])
            if pos >= len(areas) or areas[pos][0] != tarea:
                raise ReferencePanic("binary_search(...).unwrap() on a missing area")
            areas[pos][1] = (areas[pos][1] - 1) & 0xFFFFFFFFFFFFFFFF
            self.merge_cluster_into(index, target, deepen, hollow)
            upd = self.clusters[target].area()
            pos = bisect_left(areas, upd, key=lambda a: a[0])
            if pos < len(areas) and areas[pos][0] == upd:
                areas[pos][1] += 1
            else:
                areas.insert(pos, [upd, 1])
        it += 1
        return it, it == len(areas)

    def run(self):                                       # builder.rs:90-94, 201-227
        self.stage_1()
        if self.cfg.hierarchical != 0:
            it, done = 0, False
            while not done:
                it, done = self.stage_2_level(it)
        else:
            self.stage_1_output()
        return self

    def to_color_image(self):                            # container.rs:104-116, cluster.rs:90-101
        img = [(0, 0, 0, 0)] * (self.width * self.height)
        for u in reversed(self.clusters_output):
            cl = self.clusters[u]
            col = cl.residue_sum.average()
            for i in cl.indices:
                img[i] = col
        return img


def runner_run(rgba, cfg, color_image=False):
    """rgba: numpy (h, w, 4) uint8.  Returns the same dict shape as tests/oracle_lib.Oracle.runner_run (pools included)."""
    import numpy as np

    from visioncortex_b200._ffi import CLUSTER_REC_DTYPE

    h, w = rgba.shape[:2]
    px = [tuple(int(v) for v in p) for p in rgba.reshape(-1, 4)]
    b = Builder(cfg, px, w, h).run()
    n = len(b.clusters)
    rec = np.zeros(n, dtype=CLUSTER_REC_DTYPE)
    ipool, hpool = [], []
    for k, cl in enumerate(b.clusters):
        rec[k]["sum"] = cl.sum.tup()
        rec[k]["residue_sum"] = cl.residue_sum.tup()
        rec[k]["rect"] = cl.rect.tup()
        rec[k]["num_holes"], rec[k]["depth"], rec[k]["merged_into"] = cl.num_holes, cl.depth, cl.merged_into
        rec[k]["indices_len"], rec[k]["indices_off"] = len(cl.indices), len(ipool)
        rec[k]["holes_len"], rec[k]["holes_off"] = len(cl.holes), len(hpool)
        ipool.extend(cl.indices)
        hpool.extend(cl.holes)
    res = {"width": w, "height": h, "num_slots": n, "cluster_indices": np.array(b.cluster_indices, dtype=np.uint32),
           "clusters_output": np.array(b.clusters_output, dtype=np.uint32), "records": rec,
           "indices_pool": np.array(ipool, dtype=np.uint32), "holes_pool": np.array(hpool, dtype=np.uint32)}
    if color_image:
        res["color_image"] = np.array(b.to_color_image(), dtype=np.uint8).reshape(h, w, 4)
    return res


# ------------------------------------------------------------------------------------------------ clusters.rs
def to_clusters(mask, diagonal):
    """BinaryImage::to_clusters (clusters.rs:270-349).  mask: numpy (h, w) bool.  Same dict shape as Oracle.to_clusters."""
    import numpy as np

    from visioncortex_b200._ffi import BIN_CLUSTER_REC_DTYPE

    h, w = mask.shape
    m = [[bool(v) for v in row] for row in mask]

    def safe(x, y):                    # image/format.rs:66-72
        return 0 <= x < w and 0 <= y < h and m[y][x]

    clusters = []                      # [points, Rect]
    rect = Rect()
    cmap = [[0] * w for _ in range(h)]
    clusterindex = 0

    def combine(frm, to):              # clusters.rs:330-343
        for (px, py) in clusters[frm][0]:
            cmap[py][px] = to
        moved, clusters[frm][0] = clusters[frm][0], []
        clusters[to][0].extend(moved)
        clusters[to][1].merge(clusters[frm][1])

    for y in range(h):
        for x in range(w):
            v, v_up, v_left, v_ul = safe(x, y), safe(x, y - 1), safe(x - 1, y), safe(x - 1, y - 1)
            c_up = cmap[y - 1][x] if y > 0 else 0
            c_left = cmap[y][x - 1] if x > 0 else 0
            c_ul = cmap[y - 1][x - 1] if (x > 0 and y > 0) else 0
            if (v or diagonal) and v_up and v_left and c_left != c_up:
                if len(clusters[c_left][0]) <= len(clusters[c_up][0]):
                    combine(c_left, c_up)
                    if clusterindex > 0 and c_left == clusterindex - 1 and clusterindex == len(clusters):
                        clusterindex -= 1
                    c_left = c_up
                else:
                    combine(c_up, c_left)
                    c_up = c_left
            if v:
                rect.add_x_y(x, y)
                if v_up:
                    cmap[y][x] = c_up
                    clusters[c_up][0].append((x, y)); clusters[c_up][1].add_x_y(x, y)
                elif v_left:
                    cmap[y][x] = c_left
                    clusters[c_left][0].append((x, y)); clusters[c_left][1].add_x_y(x, y)
                elif v_ul and diagonal:
                    cmap[y][x] = c_ul
                    clusters[c_ul][0].append((x, y)); clusters[c_ul][1].add_x_y(x, y)
                else:
                    r = Rect()
                    r.add_x_y(x, y)
                    nc = [[(x, y)], r]
                    if clusterindex < len(clusters):
                        clusters[clusterindex] = nc
                    else:
                        clusters.append(nc)
                    cmap[y][x] = clusterindex
                    clusterindex += 1
                    if clusterindex == 65535:
                        raise ReferencePanic("overflow")
    live = [c for c in clusters if len(c[0]) != 0]       # clusters.rs:345
    rec = np.zeros(len(live), dtype=BIN_CLUSTER_REC_DTYPE)
    pts = []
    for k, (p, r) in enumerate(live):
        rec[k]["rect"] = r.tup()
        rec[k]["points_off"], rec[k]["points_len"] = len(pts), len(p)
        pts.extend(p)
    return {"width": w, "height": h, "rect": np.array(rect.tup(), dtype=np.int32), "records": rec,
            "points": np.array(pts, dtype=np.int32).reshape(-1, 2)}
