// traverse_host.cpp — the DEVICE traversal code (luisarender_b200/csrc/device/traverse.cuh: the inner-node step with its
// ordering and stack rules, the leaf step with instance entry / exit, the triangle test and its tie rule)
// compiled for the host, so that the per-ray logic the sm_100a kernels run can be checked against the oracle's BVH2 traversal
// without a GPU (tests/test_device_traversal_on_host.py).  TEST INFRASTRUCTURE: nothing here is part of the product.  What the
// GPU adds on top is warp scheduling only (ray refill, descent / leaf phases), which does not touch per-ray results.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

template<typename T>
static inline T __ldg(const T *p) { return *p; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
#ifndef __noinline__
#define __noinline__
#endif
#include <algorithm>
using std::isinf;
using std::isnan;
using std::max;
using std::min;

#include "../../luisarender_b200/csrc/device/traverse.cuh"
#include "../../luisarender_b200/csrc/device/shading.cuh"// alpha_skip (stochastic alpha test of traversal candidates)

namespace {
using namespace lrk;

// host stand-ins for the device's LaneStack / LaneWorld
struct HostStack {
    std::vector<uint32_t> e;
    size_t max_depth{0};
    void reset() { e.clear(); }
    void push(uint32_t ref) {
        e.push_back(ref);
        max_depth = std::max(max_depth, e.size());
    }
    uint32_t pop() {
        uint32_t v = e.back();
        e.pop_back();
        return v;
    }
    void push_if(bool p, uint32_t ref) {
        if (p) push(ref);
    }
    uint32_t top() const { return e.back(); }
    void drop_if(bool p) {
        if (p) e.pop_back();
    }
};

struct HostWorld {
    RaySetup saved;
    void save(const RaySetup &c) { saved = c; }
    void load(RaySetup &c) const { c = saved; }
};

// the device's per-ray control flow: inner steps while the ray stands on an inner node, leaf steps otherwise
template<bool ANY_HIT, bool ALPHA>
void trace_one(const DeviceScene &sc, const float *ray, uint32_t *hit, TraversalCounters &cnt, size_t &max_stack) {
    RayState r;
    HostStack stack;
    HostWorld world;
    start_ray(sc, r, stack, make_float4(ray[0], ray[1], ray[2], ray[3]), make_float4(ray[4], ray[5], ray[6], ray[7]));
    for (;;) {
        if (!(r.node & LRK_BVH_LEAF)) {
            inner_step<true>(sc, r, stack, cnt);
        } else if (leaf_step<ANY_HIT, true, ALPHA>(sc, r, stack, world, cnt)) {
            break;
        }
    }
    max_stack = std::max(max_stack, stack.max_depth);
    hit[0] = r.best_inst;
    hit[1] = r.best_prim;
    hit[2] = __float_as_uint(r.best_u);
    hit[3] = __float_as_uint(r.best_v);
}

}// namespace

// rays: n x {o.xyz, tmin, d.xyz, tmax}; hits: n x {inst, prim, bary.u bits, bary.v bits}; counters: {wide nodes, triangles, instance
// entries, deepest stack}
extern "C" int wide_trace_host(const lrk_scene_desc *s, const float *rays, uint64_t n, int any_hit, uint32_t *hits, uint64_t *counters) {
    std::vector<float4> xform(static_cast<size_t>(s->instance_count) * 4u);
    for (uint32_t i = 0; i < s->instance_count; i++) {// as lrk_upload_scene lays the traversal instance records out
        const auto &inst = s->instances[i];
        std::memcpy(&xform[i * 4u], inst.world_to_object, 48);
        xform[i * 4u + 3u] = make_float4(__uint_as_float(s->meshes[inst.mesh].bvh_root), 0.f, 0.f, 0.f);
    }
    std::vector<uint4> handles(s->instance_count);
    bool alpha = false;
    for (uint32_t i = 0; i < s->instance_count; i++) {
        std::memcpy(&handles[i], s->instances[i].handle, 16);
        const uint32_t flags = s->instances[i].handle[0] & 1023u;
        if ((flags & LRK_SHAPE_MAYBE_NON_OPAQUE) && (flags & LRK_SHAPE_HAS_SURFACE)) alpha = true;// lrk_upload_scene's rule
    }
    DeviceScene sc{};
    sc.self = &sc;
    sc.inst_handles = handles.data();
    sc.surfaces = s->surfaces;
    sc.meshes = s->meshes;
    sc.triangles = s->triangles;
    sc.vertices = s->vertices;
    sc.textures = s->textures;
    sc.texels = reinterpret_cast<const float4 *>(s->texels);
    sc.bvh_nodes = reinterpret_cast<const float4 *>(s->bvh_nodes);
    sc.tri_verts = reinterpret_cast<const float4 *>(s->tri_verts);
    sc.inst_xform = xform.data();
    sc.tlas_root = s->tlas_root;
    TraversalCounters cnt{0u, 0u, 0u};
    uint64_t totals[3]{0u, 0u, 0u};
    size_t max_stack = 0;
    for (uint64_t i = 0; i < n; i++) {
        cnt = TraversalCounters{0u, 0u, 0u};
        if (alpha) {
            if (any_hit) trace_one<true, true>(sc, rays + i * 8u, hits + i * 4u, cnt, max_stack);
            else trace_one<false, true>(sc, rays + i * 8u, hits + i * 4u, cnt, max_stack);
        } else {
            if (any_hit) trace_one<true, false>(sc, rays + i * 8u, hits + i * 4u, cnt, max_stack);
            else trace_one<false, false>(sc, rays + i * 8u, hits + i * 4u, cnt, max_stack);
        }
        totals[0] += cnt.nodes;
        totals[1] += cnt.tris;
        totals[2] += cnt.xforms;
    }
    if (counters) {
        counters[0] = totals[0];
        counters[1] = totals[1];
        counters[2] = totals[2];
        counters[3] = max_stack;
    }
    return 0;
}
