// Host BVH2 builder (binned SAH).  The reference has no BVH code of its own — it delegates to
// OptiX / Embree (src/compute/src/backends/cuda/cuda_accel.cpp:38-135, cpu/accel.rs:449-535) — so this
// is new host work on the scene-flattening side of the C-ABI (north_star: "the C++ host keeps scene
// parsing, BVH build and film I/O").  Output is the 64-byte two-child node layout of include/lrk.h.
#pragma once
#include <cstdint>
#include <vector>

#include "../../../include/lrk.h"
#include "vecmath.h"

namespace lrh {

struct Aabb {
    float3 lo{1e30f, 1e30f, 1e30f};
    float3 hi{-1e30f, -1e30f, -1e30f};
    void grow(float3 p) { lo = min3(lo, p); hi = max3(hi, p); }
    void grow(const Aabb &b) { lo = min3(lo, b.lo); hi = max3(hi, b.hi); }
    float half_area() const {
        auto d = hi - lo;
        if (d.x < 0.f || d.y < 0.f || d.z < 0.f) return 0.f;
        return d.x * d.y + d.y * d.z + d.z * d.x;
    }
};

struct BvhBuildResult {
    std::vector<lrk_bvh_node> nodes;  // nodes[0] is the root; child refs are local node indices / leaf refs
    std::vector<uint32_t> prim_order; // leaf i covers prim_order[first .. first+count)
};

// Build over `n` primitives with the given bounds. Leaves hold at most `max_leaf` primitives (<= 8).
// Leaf refs are LRK_BVH_LEAF | (count-1) << 28 | first, `first` indexing prim_order (caller rebases).
// With max_leaf == 1 the leaf ref is LRK_BVH_LEAF | prim index (TLAS form).
BvhBuildResult build_bvh(const Aabb *bounds, uint32_t n, uint32_t max_leaf, bool tlas_leaf_form);

}// namespace lrh
