#include "bvh.h"

#include <algorithm>
#include <cmath>
#include <future>
#include <limits>

namespace lrh {

namespace {

constexpr int kBins = 32;
// Depth cap of the SAH recursion: binned SAH can peel one bin at a time on geometrically graded input (a huge ground plane next to
// tiny detail, instance scales spanning decades), giving trees as deep as the primitive count.  Below this depth every split is
// a median split by count, so a hierarchy is never deeper than kSahDepth + ceil(log2(n)): the traversal stacks (oracle: 512
// entries, device: keyed stack with an overflow flag) and this function's own recursion stay bounded.
constexpr uint32_t kSahDepth = 48;
constexpr float kTraversalCost = 1.0f;// cost of visiting a two-box node relative to one triangle test

struct Builder {
    const Aabb *bounds;
    std::vector<float3> centroid;
    std::vector<uint32_t> idx;
    std::vector<lrk_bvh_node> nodes;
    uint32_t max_leaf;
    bool tlas;

    // Stored boxes are padded: the traversal's slab test (fmaf(lo, 1/d, -o/d) per axis, then min / max) rounds, and a ray
    // that grazes a box edge or lies in the plane of an axis-aligned (zero-thickness) leaf box could otherwise be culled
    // although the triangle test would accept it.  Found by rendering BASELINE config C1 at full size with the unmodified
    // reference (tools/gen_ref_full_size.py): 58 of 262 144 pixels differed with exact boxes; padded, the film is bit-identical.
    // Size of the pad: for an origin inside the root box the rounding of o/d, 1/d and the fmaf moves each of tn and tf by at
    // most ~1.8e-7 x extent in position units, so 5e-7 x extent covers the gap (1e-7 already makes the large renders
    // bit-identical).  It must also stay BELOW the spawned-ray offset (offset_ray_origin: 1/65536 near the origin planes):
    // a larger pad (2e-6 was measured) puts every ray that leaves a flat surface inside that surface's own padded leaf box and
    // costs +32 % triangle tests and +47 % instance transforms on the headline scene; 5e-7 costs +0.01 %.
    float pad{0.f};
    void set_box(lrk_bvh_node &n, int child, const Aabb &b) const {
        float *lo = child == 0 ? n.lo0 : n.lo1;
        float *hi = child == 0 ? n.hi0 : n.hi1;
        for (int a = 0; a < 3; a++) { lo[a] = b.lo[a] - pad; hi[a] = b.hi[a] + pad; }
    }
    static void set_empty(lrk_bvh_node &n, int child) {
        float *lo = child == 0 ? n.lo0 : n.lo1;
        float *hi = child == 0 ? n.hi0 : n.hi1;
        for (int a = 0; a < 3; a++) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -std::numeric_limits<float>::infinity(); }
        (child == 0 ? n.ref0 : n.ref1) = LRK_BVH_EMPTY;
    }
    uint32_t leaf_ref(uint32_t begin, uint32_t end) const {
        if (tlas) return LRK_BVH_LEAF | idx[begin];
        return LRK_BVH_LEAF | ((end - begin - 1u) << 28) | begin;
    }
    Aabb range_bounds(uint32_t begin, uint32_t end) const {
        Aabb b;
        for (auto i = begin; i < end; i++) b.grow(bounds[idx[i]]);
        return b;
    }

    // choose a split of [begin, end); returns mid (begin < mid < end) or begin when a leaf is cheaper/forced
    uint32_t split(uint32_t begin, uint32_t end, const Aabb &node_bounds) {
        auto n = end - begin;
        Aabb cb;
        for (auto i = begin; i < end; i++) cb.grow(centroid[idx[i]]);
        auto extent = cb.hi - cb.lo;
        float best_cost = std::numeric_limits<float>::infinity();
        int best_axis = -1;
        int best_bin = -1;
        auto parent_area = std::max(node_bounds.half_area(), 1e-30f);
        for (int axis = 0; axis < 3; axis++) {
            if (!(extent[axis] > 0.f)) continue;
            Aabb bin_bounds[kBins];
            uint32_t bin_count[kBins]{};
            auto scale = static_cast<float>(kBins) / extent[axis];
            for (auto i = begin; i < end; i++) {
                auto p = idx[i];
                auto bi = std::min(kBins - 1, std::max(0, static_cast<int>((centroid[p][axis] - cb.lo[axis]) * scale)));
                bin_count[bi]++;
                bin_bounds[bi].grow(bounds[p]);
            }
            float right_area[kBins];
            uint32_t right_count[kBins];
            Aabb acc;
            uint32_t cnt = 0;
            for (int bi = kBins - 1; bi > 0; bi--) {
                acc.grow(bin_bounds[bi]);
                cnt += bin_count[bi];
                right_area[bi] = acc.half_area();
                right_count[bi] = cnt;
            }
            Aabb left;
            uint32_t lc = 0;
            for (int bi = 0; bi < kBins - 1; bi++) {
                left.grow(bin_bounds[bi]);
                lc += bin_count[bi];
                auto rc = right_count[bi + 1];
                if (lc == 0 || rc == 0) continue;
                auto cost = kTraversalCost + (left.half_area() * static_cast<float>(lc) + right_area[bi + 1] * static_cast<float>(rc)) / parent_area;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = bi; }
            }
        }
        if (n <= max_leaf && static_cast<float>(n) <= best_cost) return begin;// leaf is cheaper
        if (best_axis < 0) {
            if (n <= max_leaf) return begin;
            return begin + n / 2;// coincident centroids: split by count
        }
        auto scale = static_cast<float>(kBins) / extent[best_axis];
        auto lo = cb.lo[best_axis];
        auto it = std::partition(idx.begin() + begin, idx.begin() + end, [&](uint32_t p) {
            auto bi = std::min(kBins - 1, std::max(0, static_cast<int>((centroid[p][best_axis] - lo) * scale)));
            return bi <= best_bin;
        });
        auto mid = static_cast<uint32_t>(it - idx.begin());
        if (mid == begin || mid == end) mid = begin + n / 2;
        return mid;
    }

    // builds [begin,end) and returns the reference to put into the parent (inner node index or leaf ref)
    uint32_t build(uint32_t begin, uint32_t end, const Aabb &range_box, uint32_t parent, uint32_t depth = 0u) {
        auto n = end - begin;
        if (n == 1u) return leaf_ref(begin, end);
        uint32_t mid;
        if (depth < kSahDepth) {
            mid = split(begin, end, range_box);
            if (mid == begin) return leaf_ref(begin, end);
        } else {
            if (n <= max_leaf) return leaf_ref(begin, end);
            mid = begin + n / 2u;// too deep for SAH: halve by count (order within the range is whatever the splits above left)
        }
        auto node_index = static_cast<uint32_t>(nodes.size());
        nodes.emplace_back();
        auto lb = range_bounds(begin, mid);
        auto rb = range_bounds(mid, end);
        auto r0 = build(begin, mid, lb, node_index, depth + 1u);
        auto r1 = build(mid, end, rb, node_index, depth + 1u);
        auto &node = nodes[node_index];
        set_box(node, 0, lb);
        set_box(node, 1, rb);
        node.ref0 = r0;
        node.ref1 = r1;
        node.parent = parent;
        node.reserved = 0u;
        return node_index;
    }
};

}// namespace

BvhBuildResult build_bvh(const Aabb *bounds, uint32_t n, uint32_t max_leaf, bool tlas_leaf_form) {
    Builder b;
    b.bounds = bounds;
    b.max_leaf = tlas_leaf_form ? 1u : std::min(std::max(max_leaf, 1u), 8u);
    b.tlas = tlas_leaf_form;
    b.centroid.resize(n);
    b.idx.resize(n);
    Aabb all;
    for (uint32_t i = 0; i < n; i++) {
        b.idx[i] = i;
        b.centroid[i] = (bounds[i].lo + bounds[i].hi) * 0.5f;
        all.grow(bounds[i]);
    }
    if (n != 0u) {
        float extent = 0.f;
        for (int a = 0; a < 3; a++) extent = std::max({extent, std::fabs(all.lo[a]), std::fabs(all.hi[a]), all.hi[a] - all.lo[a]});
        b.pad = 5e-7f * extent;
    }
    b.nodes.reserve(n);
    BvhBuildResult out;
    if (n == 0u) {
        lrk_bvh_node root{};
        Builder::set_empty(root, 0);
        Builder::set_empty(root, 1);
        root.parent = LRK_BVH_EMPTY;
        out.nodes.push_back(root);
        return out;
    }
    auto ref = b.build(0u, n, all, LRK_BVH_EMPTY);
    if (ref & LRK_BVH_LEAF) {
        // the whole range became one leaf: wrap it so that the root is always an inner node
        lrk_bvh_node root{};
        b.set_box(root, 0, all);
        root.ref0 = ref;
        Builder::set_empty(root, 1);
        root.parent = LRK_BVH_EMPTY;
        b.nodes.insert(b.nodes.begin(), root);
    }
    out.nodes = std::move(b.nodes);
    out.prim_order = std::move(b.idx);
    return out;
}

}// namespace lrh
