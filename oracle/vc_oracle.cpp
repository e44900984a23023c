// vc_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A single-threaded CPU restatement (C++17) of the clustering path of visioncortex v0.8.10, written from the
// reference's Rust sources line by line so that it can serve as the parity oracle for the CUDA path.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// The product (visioncortex_b200/) never links, imports or calls it.
//
// Pinning: `BinaryImage::to_clusters` is pinned by the reference's own unit tests (src/clusters.rs:351-526),
// replayed in tests/test_oracle_kat.py.  `color_clusters::Runner` has NO tests in the reference, so for the
// colour path the oracle is "parity unpinned by reference vectors": the pin is this restatement itself
// (no Rust toolchain exists in the build image, so the reference cannot be executed).
//
// Semantics fixed explicitly: Rust *release* profile (u32/usize wrapping arithmetic, no overflow panics),
// stable sort_by_key, HashSet -> sorted Vec (hash order is irrelevant, color_clusters/cluster.rs:156-158).
//
// Every function cites the reference file:line it follows (paths relative to the reference's src/).

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../include/vcb200.h"

namespace vco {

// ---------------------------------------------------------------- color.rs:8-14
struct Color {
    uint8_t r = 0, g = 0, b = 0, a = 0;
    bool operator==(const Color &o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
    bool operator!=(const Color &o) const { return !(*this == o); }
};

struct ReferencePanic { const char *what; };          // the reference would panic on this input

// ---------------------------------------------------------------- color.rs:40-47, 256-311
struct ColorSum {
    uint32_t r = 0, g = 0, b = 0, a = 0, counter = 0;
    void add(const Color &c) {                      // color.rs:261-267 (u32 wrapping in release)
        r += c.r; g += c.g; b += c.b; a += c.a; counter += 1;
    }
    void merge(const ColorSum &o) {                 // color.rs:278-284
        r += o.r; g += o.g; b += o.b; a += o.a; counter += o.counter;
    }
    Color average() const {                         // color.rs:295-302 (integer division, `as u8` truncation)
        // counter == 0: Rust panics ("attempt to divide by zero") in every profile.  Reachable in stage 2 through the
        // diagonal = true orphan pixel: its label names a slot that has been merged away, which
        // then shows up as a neighbour with an empty cluster (builder.rs:436-442).
        if (counter == 0) throw ReferencePanic{"attempt to divide by zero (color.rs:295-302)"};
        Color c;
        c.r = (uint8_t)(r / counter); c.g = (uint8_t)(g / counter);
        c.b = (uint8_t)(b / counter); c.a = (uint8_t)(a / counter);
        return c;
    }
    void clear() { r = g = b = a = counter = 0; }   // color.rs:304-310
};

// ---------------------------------------------------------------- bound.rs:15-21, 86-103, 172-227
struct BoundingRect {
    int32_t left = 0, top = 0, right = 0, bottom = 0;
    int32_t width() const { return right - left; }
    int32_t height() const { return bottom - top; }
    bool is_empty() const { return width() == 0 && height() == 0; }   // bound.rs:101-103
    void add_x_y(int32_t x, int32_t y) {                              // bound.rs:172-190
        if (is_empty()) { left = x; right = x + 1; top = y; bottom = y + 1; return; }
        if (x < left) left = x; else if (x + 1 > right) right = x + 1;
        if (y < top) top = y; else if (y + 1 > bottom) bottom = y + 1;
    }
    void merge(const BoundingRect &o) {                               // bound.rs:204-219
        if (o.is_empty()) return;
        if (is_empty()) { *this = o; return; }
        left = std::min(left, o.left); right = std::max(right, o.right);
        top = std::min(top, o.top); bottom = std::max(bottom, o.bottom);
    }
    void clear() { left = top = right = bottom = 0; }                 // bound.rs:222-227
};

// ---------------------------------------------------------------- image/format.rs:10-14, 43-77
struct BinaryImage {
    std::vector<uint8_t> px;
    size_t width = 0, height = 0;
    BinaryImage() {}
    BinaryImage(size_t w, size_t h) : px(w * h, 0), width(w), height(h) {}    // format.rs:44-50
    bool get_pixel(size_t x, size_t y) const { return px[y * width + x] != 0; } // format.rs:58-61
    bool get_pixel_safe(int32_t x, int32_t y) const {                          // format.rs:66-72
        if (x >= 0 && x < (int32_t)width && y >= 0 && y < (int32_t)height) return get_pixel((size_t)x, (size_t)y);
        return false;
    }
    void set_pixel(size_t x, size_t y, bool v) { px[y * width + x] = v ? 1 : 0; } // format.rs:74-77
};

// shape/geometry.rs:45-75 — Shape::image_boundary_list(image).len() (transpose = false): number of set pixels
// with an unset or out-of-image 4-neighbour.
static uint32_t image_boundary_count(const BinaryImage &im) {
    uint32_t n = 0;
    for (int32_t y = 0; y < (int32_t)im.height; ++y)
        for (int32_t x = 0; x < (int32_t)im.width; ++x)
            if (im.get_pixel((size_t)x, (size_t)y) &&
                (!im.get_pixel_safe(x - 1, y) || !im.get_pixel_safe(x + 1, y) ||
                 !im.get_pixel_safe(x, y - 1) || !im.get_pixel_safe(x, y + 1)))
                ++n;
    return n;
}

// ================================================================ colour path
// color_clusters/cluster.rs:7-17
struct Cluster {
    std::vector<uint32_t> indices, holes;
    uint32_t num_holes = 0, depth = 0;
    ColorSum sum, residue_sum;
    BoundingRect rect;
    uint32_t merged_into = 0;
    void add(uint32_t i, const Color &c, int32_t x, int32_t y) {      // cluster.rs:24-28
        indices.push_back(i); sum.add(c); rect.add_x_y(x, y);
    }
    size_t area() const { return indices.size(); }                    // cluster.rs:30-32
    Color color() const { return sum.average(); }                     // cluster.rs:38-40
    Color residue_color() const { return residue_sum.average(); }     // cluster.rs:42-44
    // cluster.rs:60-80
    BinaryImage to_image_with_hole(uint32_t parent_width, bool hole) const {
        BinaryImage image((size_t)rect.width(), (size_t)rect.height());
        for (uint32_t i : indices) {
            int32_t x = ((int32_t)i % (int32_t)parent_width) - rect.left;
            int32_t y = ((int32_t)i / (int32_t)parent_width) - rect.top;
            image.set_pixel((size_t)x, (size_t)y, true);
        }
        if (hole)
            for (uint32_t i : holes) {
                int32_t x = ((int32_t)i % (int32_t)parent_width) - rect.left;
                int32_t y = ((int32_t)i / (int32_t)parent_width) - rect.top;
                image.set_pixel((size_t)x, (size_t)y, false);
            }
        return image;
    }
};

struct NeighbourInfo { uint32_t index; int32_t diff; };               // builder.rs:34-37
struct Area { size_t area; size_t count; };                           // builder.rs:143-146

// runner.rs:110-114
static int32_t color_diff(const Color &a, const Color &b) {
    return std::abs((int32_t)a.r - (int32_t)b.r) + std::abs((int32_t)a.g - (int32_t)b.g) +
           std::abs((int32_t)a.b - (int32_t)b.b);
}
// runner.rs:116-129
static bool color_same(const Color &a, const Color &b, int32_t shift, int32_t thres) {
    int32_t dr = (int32_t)(a.r >> shift) - (int32_t)(b.r >> shift);
    int32_t dg = (int32_t)(a.g >> shift) - (int32_t)(b.g >> shift);
    int32_t db = (int32_t)(a.b >> shift) - (int32_t)(b.b >> shift);
    return std::abs(dr) <= thres && std::abs(dg) <= thres && std::abs(db) <= thres;
}

// color_clusters/builder.rs:148-569 (BuilderImpl) with the Runner closures of runner.rs:63-98 bound in.
struct BuilderImpl {
    // runner config
    bool diagonal; uint32_t hierarchical; uint32_t batch_size;
    Color key; int keying_action;
    size_t good_min_area, good_max_area; int32_t same_a, same_b, deepen_diff; size_t hollow_neighbours;
    // state (builder.rs:158-167)
    uint32_t width, height;
    std::vector<uint8_t> pixels;
    std::vector<Cluster> clusters;
    std::vector<uint32_t> cluster_indices;
    std::vector<Area> cluster_areas;
    std::vector<uint32_t> clusters_output;
    uint32_t stage = 1, iteration = 0, next_index = 1;

    struct OptColor { bool some; Color c; };

    // builder.rs:557-569
    OptColor get_pixel(uint32_t i) const {
        size_t k = (size_t)i * 4;
        if (k < pixels.size()) return {true, Color{pixels[k], pixels[k + 1], pixels[k + 2], pixels[k + 3]}};
        return {false, Color{}};
    }
    // builder.rs:549-555
    OptColor pixel_at(int32_t x, int32_t y) const {
        if (x < 0 || y < 0) return {false, Color{}};
        return get_pixel((uint32_t)y * width + (uint32_t)x);
    }
    // builder.rs:541-547 with the `same` closure of runner.rs:87-89
    bool is_same(const OptColor &l, const OptColor &r) const {
        if (l.some && r.some) return color_same(l.c, r.c, same_a, same_b);
        return false;
    }

    // builder.rs:526-539
    void combine_clusters(uint32_t from, uint32_t to) {
        for (uint32_t i : clusters[from].indices) cluster_indices[i] = to;
        std::vector<uint32_t> moved;
        moved.swap(clusters[from].indices);
        clusters[to].indices.insert(clusters[to].indices.end(), moved.begin(), moved.end());
        ColorSum sum = clusters[from].sum;
        BoundingRect rect = clusters[from].rect;
        clusters[to].sum.merge(sum);
        clusters[to].rect.merge(rect);
        clusters[from].sum.clear();
        clusters[from].rect.clear();
    }
    // builder.rs:514-524
    void combine_clusters_clone(uint32_t from, uint32_t to) {
        ColorSum sum = clusters[from].sum;
        BoundingRect rect = clusters[from].rect;
        std::vector<uint32_t> indices = clusters[from].indices;
        combine_clusters(from, to);
        clusters[from].sum = sum;
        clusters[from].rect = rect;
        clusters[from].indices = indices;
    }
    // builder.rs:495-512
    void merge_cluster_into(uint32_t from, uint32_t to, bool deepen, bool hollow) {
        if (!deepen) {
            ColorSum residue = clusters[from].residue_sum;
            clusters[to].residue_sum.merge(residue);
            combine_clusters(from, to);
        } else {
            combine_clusters_clone(from, to);
            if (hollow) {
                std::vector<uint32_t> holes = clusters[from].indices;
                clusters[to].holes.insert(clusters[to].holes.end(), holes.begin(), holes.end());
                clusters[to].num_holes += 1;
            }
            clusters[from].merged_into = to;
            clusters[to].depth += 1;
        }
    }

    // builder.rs:273-364
    bool stage_1() {
        const bool has_key = key != Color{};
        const size_t len = cluster_indices.size();
        // (self.iteration..(self.iteration + batch_size)).take_while(|&i| (i as usize) < len), u32 wrapping add
        const uint32_t end = iteration + batch_size;
        for (uint32_t i = iteration; i < end && (size_t)i < len; ++i) {
            const int32_t x = (int32_t)(i % width), y = (int32_t)(i / width);
            const OptColor color = pixel_at(x, y), up = pixel_at(x, y - 1), left = pixel_at(x - 1, y),
                           upleft = pixel_at(x - 1, y - 1);
            uint32_t cluster_up = y > 0 ? cluster_indices[(size_t)((int32_t)width * (y - 1) + x)] : 0;
            uint32_t cluster_left = x > 0 ? cluster_indices[(size_t)((int32_t)width * y + (x - 1))] : 0;
            const uint32_t cluster_upleft =
                (x > 0 && y > 0) ? cluster_indices[(size_t)((int32_t)width * (y - 1) + (x - 1))] : 0;

            if (cluster_left != cluster_up && is_same(left, up) &&
                (diagonal || (is_same(color, left) && is_same(color, up)))) {
                if (clusters[cluster_left].area() <= clusters[cluster_up].area()) {
                    combine_clusters(cluster_left, cluster_up);
                    if (cluster_left == next_index - 1 && (size_t)next_index == clusters.size()) next_index -= 1;
                    cluster_left = cluster_up;
                } else {
                    combine_clusters(cluster_up, cluster_left);
                    cluster_up = cluster_left;
                }
            }

            const Color c = color.c;   // color.unwrap()
            if (has_key && c == key) {
                if (keying_action == VCB_KEYING_KEEP) clusters[0].add(i, c, x, y);
            } else if (is_same(color, up) && is_same(color, upleft)) {
                cluster_indices[i] = cluster_up;
                clusters[cluster_up].add(i, c, x, y);
            } else if (is_same(color, left) && is_same(color, upleft)) {
                cluster_indices[i] = cluster_left;
                clusters[cluster_left].add(i, c, x, y);
            } else if (diagonal && is_same(color, upleft)) {
                cluster_indices[i] = cluster_upleft;
                clusters[cluster_upleft].add(i, c, x, y);
            } else {
                Cluster nc;
                nc.add(i, c, x, y);
                if ((size_t)next_index < clusters.size()) clusters[next_index] = std::move(nc);
                else clusters.push_back(std::move(nc));
                cluster_indices[i] = next_index;
                next_index += 1;
            }
        }
        iteration += batch_size;
        if ((size_t)iteration >= cluster_indices.size()) { prepare_stage_2(); return true; }
        return false;
    }

    // builder.rs:366-377
    void stage_1_output() {
        std::vector<std::pair<uint32_t, size_t>> output;
        for (size_t index = 0; index < clusters.size(); ++index) {
            size_t area = clusters[index].area();
            if (area > 0) output.push_back({(uint32_t)index, area});
        }
        std::stable_sort(output.begin(), output.end(), [](const auto &a, const auto &b) {
            return (uint64_t)a.second * 65535 + a.first < (uint64_t)b.second * 65535 + b.first;
        });
        for (auto &c : output) clusters_output.push_back(c.first);
    }

    // builder.rs:379-403
    void prepare_stage_2() {
        for (auto &c : clusters) c.residue_sum = c.sum;
        std::vector<size_t> areas;
        for (auto &c : clusters) if (c.area() > 0) areas.push_back(c.area());
        std::sort(areas.begin(), areas.end());
        cluster_areas.clear();
        for (size_t a : areas) {
            if (!cluster_areas.empty() && cluster_areas.back().area == a) cluster_areas.back().count += 1;
            else cluster_areas.push_back({a, 1});
        }
    }

    // cluster.rs:132-171 (neighbours_impl! via neighbours_internal)
    std::vector<uint32_t> neighbours_internal(const Cluster &cl) const {
        const uint32_t myself = cluster_indices[cl.indices.front()];
        std::vector<uint32_t> list;
        for (uint32_t i : cl.indices) {
            const uint32_t x = i % width, y = i / width;
            for (int k = 0; k < 4; ++k) {
                uint32_t index = 0;
                switch (k) {
                    case 0: index = y > 0 ? cluster_indices[(size_t)width * (y - 1) + x] : 0; break;
                    case 1: index = y < height - 1 ? cluster_indices[(size_t)width * (y + 1) + x] : 0; break;
                    case 2: index = x > 0 ? cluster_indices[(size_t)width * y + (x - 1)] : 0; break;
                    case 3: index = x < width - 1 ? cluster_indices[(size_t)width * y + (x + 1)] : 0; break;
                }
                if (index != 0 && index != myself) list.push_back(index);
            }
        }
        std::sort(list.begin(), list.end());
        list.erase(std::unique(list.begin(), list.end()), list.end());
        return list;
    }

    // runner.rs:131-146 via cluster.rs:49-51,56-58
    bool patch_good(const Cluster &patch) const {
        if (good_min_area < patch.area() && patch.area() < good_max_area) {
            if (good_min_area == 0 ||
                (size_t)image_boundary_count(patch.to_image_with_hole(width, true)) < patch.area())
                return true;
        }
        return false;
    }

    // builder.rs:405-493
    bool stage_2() {
        if (cluster_areas.empty()) return true;
        if (cluster_areas[iteration].count == 0) {
            iteration += 1;
            return (size_t)iteration == cluster_areas.size();
        }
        const size_t cur_area = cluster_areas[iteration].area;
        const bool can_discard_pixels = keying_action == VCB_KEYING_DISCARD && key != Color{};

        for (size_t idx = 0; idx < clusters.size(); ++idx) {
            const uint32_t index = (uint32_t)idx;
            if (clusters[index].area() != cur_area) continue;
            if (cur_area > (size_t)hierarchical) { clusters_output.push_back(index); continue; }

            const Color mycolor = clusters[index].color();
            std::vector<NeighbourInfo> infos;
            for (uint32_t other : neighbours_internal(clusters[index]))
                infos.push_back({other, color_diff(mycolor, clusters[other].color())});

            if (infos.empty()) {
                if (iteration == (uint32_t)cluster_areas.size() - 1 || can_discard_pixels)
                    clusters_output.push_back(index);
                continue;
            }
            std::stable_sort(infos.begin(), infos.end(), [](const NeighbourInfo &a, const NeighbourInfo &b) {
                return (int64_t)a.diff * 65535 + (int64_t)a.index < (int64_t)b.diff * 65535 + (int64_t)b.index;
            });
            const uint32_t target = infos[0].index;

            // deepen closure runner.rs:91-94, hollow closure runner.rs:95-97
            bool deepen = false;
            if (hierarchical == VCB_HIERARCHICAL_MAX)
                deepen = patch_good(clusters[index]) && infos[0].diff > deepen_diff;
            const bool hollow = infos.size() <= hollow_neighbours;

            if (deepen) clusters_output.push_back(index);

            auto find_area = [&](size_t a) -> std::pair<bool, size_t> {
                auto it = std::lower_bound(cluster_areas.begin(), cluster_areas.end(), a,
                                           [](const Area &x, size_t v) { return x.area < v; });
                size_t pos = (size_t)(it - cluster_areas.begin());
                return {it != cluster_areas.end() && it->area == a, pos};
            };
            auto tin = find_area(clusters[target].area());     // .unwrap(): always present
            cluster_areas[tin.second].count -= 1;
            merge_cluster_into(index, target, deepen, hollow);
            const size_t updated_area = clusters[target].area();
            auto up = find_area(updated_area);
            if (up.first) cluster_areas[up.second].count += 1;
            else cluster_areas.insert(cluster_areas.begin() + (ptrdiff_t)up.second, Area{updated_area, 1});
        }
        iteration += 1;
        return (size_t)iteration == cluster_areas.size();
    }

    // builder.rs:201-227
    bool tick() {
        switch (stage) {
            case 1:
                if (stage_1()) {
                    if (hierarchical != 0) { stage += 1; iteration = 0; }
                    else { stage_1_output(); stage += 2; }
                }
                return false;
            case 2: {
                uint32_t reps = std::max<uint32_t>(1, iteration / 16);
                for (uint32_t k = 0; k < reps; ++k)
                    if (stage_2()) { stage += 1; iteration = 0; break; }
                return false;
            }
            default: return true;
        }
    }
};

// ================================================================ binary path (clusters.rs)
struct BinCluster {                                  // clusters.rs:7-11, 25-32
    std::vector<int32_t> points;                     // x,y pairs
    BoundingRect rect;
    void add(int32_t x, int32_t y) { points.push_back(x); points.push_back(y); rect.add_x_y(x, y); }
    size_t size() const { return points.size() / 2; }
};

struct BinResult {
    std::vector<BinCluster> clusters;
    BoundingRect rect;
    uint32_t width = 0, height = 0;
};

// clusters.rs:270-349; returns false on the reference's panic!("overflow") (clusters.rs:322-324)
static bool to_clusters(const BinaryImage &im, bool diagonal, BinResult &res) {
    std::vector<BinCluster> clusters;
    BoundingRect rect;
    std::vector<uint16_t> clustermap(im.width * im.height, 0);        // MonoImage::new_w_h (format.rs:236-242)
    uint16_t clusterindex = 0;
    auto combine_cluster = [&](uint16_t from, uint16_t to) {          // clusters.rs:330-343
        auto &pts = clusters[from].points;
        for (size_t k = 0; k < pts.size(); k += 2) clustermap[(size_t)pts[k + 1] * im.width + (size_t)pts[k]] = to;
        std::vector<int32_t> drain;
        drain.swap(clusters[from].points);
        clusters[to].points.insert(clusters[to].points.end(), drain.begin(), drain.end());
        BoundingRect r = clusters[from].rect;
        clusters[to].rect.merge(r);
    };
    for (size_t y = 0; y < im.height; ++y) {
        for (size_t x = 0; x < im.width; ++x) {
            const bool v = im.get_pixel_safe((int32_t)x, (int32_t)y);
            const bool v_up = im.get_pixel_safe((int32_t)x, (int32_t)y - 1);
            const bool v_left = im.get_pixel_safe((int32_t)x - 1, (int32_t)y);
            const bool v_up_left = im.get_pixel_safe((int32_t)x - 1, (int32_t)y - 1);
            uint16_t cluster_up = y > 0 ? clustermap[(y - 1) * im.width + x] : 0;
            uint16_t cluster_left = x > 0 ? clustermap[y * im.width + x - 1] : 0;
            const uint16_t cluster_up_left = (x > 0 && y > 0) ? clustermap[(y - 1) * im.width + x - 1] : 0;
            if ((v || diagonal) && v_up && v_left && cluster_left != cluster_up) {
                if (clusters[cluster_left].size() <= clusters[cluster_up].size()) {
                    combine_cluster(cluster_left, cluster_up);
                    if (clusterindex > 0 && cluster_left == clusterindex - 1 && (size_t)clusterindex == clusters.size())
                        clusterindex -= 1;
                    cluster_left = cluster_up;
                } else {
                    combine_cluster(cluster_up, cluster_left);
                    cluster_up = cluster_left;
                }
            }
            if (v) {
                rect.add_x_y((int32_t)x, (int32_t)y);
                if (v_up) {
                    clustermap[y * im.width + x] = cluster_up;
                    clusters[cluster_up].add((int32_t)x, (int32_t)y);
                } else if (v_left) {
                    clustermap[y * im.width + x] = cluster_left;
                    clusters[cluster_left].add((int32_t)x, (int32_t)y);
                } else if (v_up_left && diagonal) {
                    clustermap[y * im.width + x] = cluster_up_left;
                    clusters[cluster_up_left].add((int32_t)x, (int32_t)y);
                } else {
                    BinCluster nc;
                    nc.add((int32_t)x, (int32_t)y);
                    if ((size_t)clusterindex < clusters.size()) clusters[clusterindex] = std::move(nc);
                    else clusters.push_back(std::move(nc));
                    clustermap[y * im.width + x] = clusterindex;
                    clusterindex += 1;
                    if (clusterindex == 0xFFFF) return false;       // panic!("overflow")
                }
            }
        }
    }
    res.clusters.clear();
    for (auto &c : clusters) if (c.size() != 0) res.clusters.push_back(std::move(c));   // clusters.rs:345
    res.rect = rect;
    res.width = (uint32_t)im.width; res.height = (uint32_t)im.height;
    return true;
}

}  // namespace vco

// ================================================================ C API (same shapes as include/vcb200.h, prefix vco_)
struct vco_clusters {
    uint32_t width, height;
    std::vector<vco::Cluster> clusters;
    std::vector<uint32_t> cluster_indices, clusters_output;
};
struct vco_bin_clusters { vco::BinResult res; };

extern "C" {

int vco_runner_run(const vcb_runner_config *cfg, const uint8_t *rgba, uint32_t width, uint32_t height,
                   vco_clusters **out) {
    if (!cfg || !out || (!rgba && (size_t)width * height != 0)) return VCB_ERR_INVALID_ARG;
    if (!(cfg->is_same_color_a < 8)) return VCB_ERR_INVALID_ARG;            // runner.rs:78
    if (cfg->is_same_color_a < 0) return VCB_ERR_INVALID_ARG;               // u8 >> negative: overflow panic
    if (cfg->batch_size == 0) return VCB_ERR_INVALID_ARG;                   // reference never terminates
    vco::BuilderImpl b;
    b.diagonal = cfg->diagonal != 0; b.hierarchical = cfg->hierarchical; b.batch_size = (uint32_t)cfg->batch_size;
    b.key = vco::Color{cfg->key_color[0], cfg->key_color[1], cfg->key_color[2], cfg->key_color[3]};
    b.keying_action = cfg->keying_action;
    b.good_min_area = (size_t)cfg->good_min_area; b.good_max_area = (size_t)cfg->good_max_area;
    b.same_a = cfg->is_same_color_a; b.same_b = cfg->is_same_color_b; b.deepen_diff = cfg->deepen_diff;
    b.hollow_neighbours = (size_t)cfg->hollow_neighbours;
    b.width = width; b.height = height;
    b.pixels.assign(rgba, rgba + (size_t)width * height * 4);
    b.clusters.assign(1, vco::Cluster());                                   // builder.rs:189
    b.cluster_indices.assign((size_t)width * height, 0);                    // builder.rs:190
    try {
        while (!b.tick()) {}                                                // builder.rs:92
    } catch (const vco::ReferencePanic &) {
        return VCB_ERR_UNSUPPORTED;                                         // the reference panics on this input
    }
    auto *r = new (std::nothrow) vco_clusters();
    if (!r) return VCB_ERR_NOMEM;
    r->width = width; r->height = height;
    r->clusters = std::move(b.clusters);
    r->cluster_indices = std::move(b.cluster_indices);
    r->clusters_output = std::move(b.clusters_output);
    *out = r;
    return VCB_OK;
}

int vco_clusters_info_get(const vco_clusters *c, vcb_clusters_info *info) {
    info->width = c->width; info->height = c->height;
    info->num_slots = (uint32_t)c->clusters.size();
    info->num_outputs = (uint32_t)c->clusters_output.size();
    uint64_t it = 0, ht = 0;
    for (auto &cl : c->clusters) { it += cl.indices.size(); ht += cl.holes.size(); }
    info->indices_total = it; info->holes_total = ht;
    return VCB_OK;
}
int vco_clusters_cluster_indices(const vco_clusters *c, uint32_t *out) {
    std::memcpy(out, c->cluster_indices.data(), c->cluster_indices.size() * 4); return VCB_OK;
}
int vco_clusters_outputs(const vco_clusters *c, uint32_t *out) {
    std::memcpy(out, c->clusters_output.data(), c->clusters_output.size() * 4); return VCB_OK;
}
int vco_clusters_records(const vco_clusters *c, vcb_cluster_rec *out) {
    uint64_t io = 0, ho = 0;
    for (size_t k = 0; k < c->clusters.size(); ++k) {
        const auto &cl = c->clusters[k];
        vcb_cluster_rec &r = out[k];
        std::memset(&r, 0, sizeof r);
        r.sum[0] = cl.sum.r; r.sum[1] = cl.sum.g; r.sum[2] = cl.sum.b; r.sum[3] = cl.sum.a; r.sum[4] = cl.sum.counter;
        r.residue_sum[0] = cl.residue_sum.r; r.residue_sum[1] = cl.residue_sum.g; r.residue_sum[2] = cl.residue_sum.b;
        r.residue_sum[3] = cl.residue_sum.a; r.residue_sum[4] = cl.residue_sum.counter;
        r.rect[0] = cl.rect.left; r.rect[1] = cl.rect.top; r.rect[2] = cl.rect.right; r.rect[3] = cl.rect.bottom;
        r.num_holes = cl.num_holes; r.depth = cl.depth; r.merged_into = cl.merged_into;
        r.indices_len = (uint32_t)cl.indices.size(); r.indices_off = io; io += cl.indices.size();
        r.holes_len = (uint32_t)cl.holes.size(); r.holes_off = ho; ho += cl.holes.size();
    }
    return VCB_OK;
}
int vco_clusters_indices_pool(const vco_clusters *c, uint32_t *out) {
    for (auto &cl : c->clusters) { std::memcpy(out, cl.indices.data(), cl.indices.size() * 4); out += cl.indices.size(); }
    return VCB_OK;
}
int vco_clusters_holes_pool(const vco_clusters *c, uint32_t *out) {
    for (auto &cl : c->clusters) { std::memcpy(out, cl.holes.data(), cl.holes.size() * 4); out += cl.holes.size(); }
    return VCB_OK;
}
// container.rs:104-116 + cluster.rs:90-101: reverse-order paint with residue_color
int vco_clusters_to_color_image(const vco_clusters *c, uint8_t *rgba_out) {
    std::memset(rgba_out, 0, (size_t)c->width * c->height * 4);              // ColorImage::new_w_h zero-filled
    for (size_t k = c->clusters_output.size(); k-- > 0;) {
        const auto &cl = c->clusters[c->clusters_output[k]];
        vco::Color col;
        try { col = cl.residue_color(); } catch (const vco::ReferencePanic &) { return VCB_ERR_UNSUPPORTED; }   // the reference panics
        for (uint32_t i : cl.indices) {
            rgba_out[(size_t)i * 4 + 0] = col.r; rgba_out[(size_t)i * 4 + 1] = col.g;
            rgba_out[(size_t)i * 4 + 2] = col.b; rgba_out[(size_t)i * 4 + 3] = col.a;
        }
    }
    return VCB_OK;
}
// Cluster::perimeter(view) (cluster.rs:46-48) for one slot — exposed for tests of the boundary-count kernel
int vco_clusters_perimeter(const vco_clusters *c, uint32_t slot, uint32_t *out) {
    if (slot >= c->clusters.size()) return VCB_ERR_INVALID_ARG;
    *out = vco::image_boundary_count(c->clusters[slot].to_image_with_hole(c->width, true));
    return VCB_OK;
}
void vco_clusters_free(vco_clusters *c) { delete c; }

int vco_binary_to_clusters(const uint32_t *bits, uint32_t width, uint32_t height, int diagonal,
                           vco_bin_clusters **out) {
    if (!out || (!bits && (size_t)width * height != 0)) return VCB_ERR_INVALID_ARG;
    vco::BinaryImage im(width, height);
    for (size_t i = 0; i < (size_t)width * height; ++i) im.px[i] = (bits[i >> 5] >> (i & 31)) & 1u;
    auto *r = new (std::nothrow) vco_bin_clusters();
    if (!r) return VCB_ERR_NOMEM;
    if (!vco::to_clusters(im, diagonal != 0, r->res)) { delete r; return VCB_ERR_LABEL_OVERFLOW; }
    *out = r;
    return VCB_OK;
}
int vco_bin_clusters_info_get(const vco_bin_clusters *c, vcb_bin_clusters_info *info) {
    std::memset(info, 0, sizeof *info);
    info->width = c->res.width; info->height = c->res.height;
    info->num_clusters = (uint32_t)c->res.clusters.size();
    uint64_t t = 0;
    for (auto &cl : c->res.clusters) t += cl.size();
    info->points_total = t;
    info->rect[0] = c->res.rect.left; info->rect[1] = c->res.rect.top;
    info->rect[2] = c->res.rect.right; info->rect[3] = c->res.rect.bottom;
    return VCB_OK;
}
int vco_bin_clusters_records(const vco_bin_clusters *c, vcb_bin_cluster_rec *out) {
    uint64_t off = 0;
    for (size_t k = 0; k < c->res.clusters.size(); ++k) {
        const auto &cl = c->res.clusters[k];
        std::memset(&out[k], 0, sizeof out[k]);
        out[k].rect[0] = cl.rect.left; out[k].rect[1] = cl.rect.top;
        out[k].rect[2] = cl.rect.right; out[k].rect[3] = cl.rect.bottom;
        out[k].points_off = off; out[k].points_len = (uint32_t)cl.size(); off += cl.size();
    }
    return VCB_OK;
}
int vco_bin_clusters_points(const vco_bin_clusters *c, int32_t *xy_out) {
    for (auto &cl : c->res.clusters) { std::memcpy(xy_out, cl.points.data(), cl.points.size() * 4); xy_out += cl.points.size(); }
    return VCB_OK;
}
void vco_bin_clusters_free(vco_bin_clusters *c) { delete c; }

}  // extern "C"
