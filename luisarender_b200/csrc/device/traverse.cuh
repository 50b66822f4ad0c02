// Two-level wide-BVH traversal + ray/triangle intersection for sm_100a (B200 has no RT cores).
//
// Replaces Geometry::trace_closest / trace_any -> Accel::intersect / intersect_any
// (reference src/base/geometry.cpp:218-279), which the reference delegates to OptiX
// (src/compute/src/backends/cuda/cuda_builtin/cuda_device_resource.h:1603-1693) or Embree
// (src/compute/src/rust/luisa_compute_backend_impl/src/cpu/accel.rs:449-535).
//
// Per-ray rules (box and triangle arithmetic shared verbatim with the CPU oracle, oracle/oracle.cpp "BVH traversal"):
//   * 4-wide nodes collapsed on the device from the host's BVH2 (wide_bvh.cuh); slab test t = fma(plane, 1/d, -o/d) with
//     |d| clamped to >= 1e-30, near / far plane chosen by the sign of d; a child is hit when
//     max(t_entry, tmin) <= min(t_exit, t_best);
//   * closest-hit: the nearest hit child is visited first, the others are deferred with their entry distance and dropped
//     when popped if the ray has been shortened below it; any-hit: slot order, the first accepted triangle ends the ray;
//   * Moeller-Trumbore in object space with fma dot/cross; accept tmin < t < t_best, u,v >= 0, u+v <= 1; an exact tie
//     t == t_best goes to the lower (instance, primitive), so that the result does not depend on the visiting order;
//   * instance entry transforms the ray by world_to_object without renormalising d (t is shared).
//
// TLAS FIRST.  A ray walks the top-level hierarchy to the end before it enters any instance: instances whose box it hits are
// appended (index + entry distance) to a small per-lane candidate queue, and are then entered one after the other, skipping those
// whose box starts behind the hit found so far.  Compared with the usual nested walk (enter an instance as soon as its leaf is
// reached, come back to the TLAS through an exit sentinel) this removes the world-ray restore and the sentinel traffic, and - what
// matters on a SIMT machine - it lines the lanes of a warp up: freshly fetched rays do their TLAS steps together, then reach
// their first instance entry together.  The queue holds kListSize candidates; a TLAS node is only expanded while four slots are
// free, otherwise the TLAS walk is SUSPENDED (its stack entries stay where they are, the instance stack grows above them) and
// resumed once the queue has been worked off - by then with a shortened ray - so any number of overlapping instances is handled.
//
// Memory: the stack holds (ref, entry-distance key) pairs, the first kSmemStack entries per thread in shared memory (lane-
// interleaved, conflict-free), deeper ones in local memory; the candidate queue and the world-space origin / direction (needed
// for every instance transform) live in shared memory too.
//
// Warp scheduling (no oracle counterpart — it does not change any per-ray result): incoherent rays have very different
// traversal lengths, so rays are pulled from the queue through a per-launch atomic cursor and a warp REFILLS its idle lanes
// with fresh rays whenever fewer than `refill_below` lanes are still traversing (persistent warps with dynamic ray
// replacement); lanes that stand on a leaf wait while at least `inner_min` lanes still descend (descent / leaf phases).
#pragma once
#include "scene.cuh"
#include "wide_bvh.cuh"

namespace lrk {

constexpr uint32_t kSentinelDone = 0xfffffffdu;// bottom of a hierarchy's stack: popping it means "this hierarchy is exhausted"
constexpr int kSmemStack = 8;  // stack entries per thread kept in shared memory
constexpr int kLocalStack = 56;// further entries in local memory (a 4-wide tree defers at most 3 children per level)
constexpr uint32_t kListSize = 8u;// instance candidates per thread (a ring in shared memory); power of two
constexpr int kRefillBelow = 16;// refill when fewer than this many lanes of the warp hold a live ray
constexpr int kInnerMin = 8;   // leave the inner-node phase when fewer lanes than this still descend

struct TraversalCounters {
    uint32_t nodes, tris, xforms;
};

__device__ __forceinline__ float safe_rcp(float d) {
    float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
#if defined(LRK_FAST_RCP) && defined(__CUDA_ARCH__)
    float r;// experiment: 1-ulp MUFU reciprocal for the slab test only (the triangle test never sees it)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}

// The ray in the space of the hierarchy being walked (world space in the TLAS, object space inside an instance).
struct RaySetup {
    V3 o, d, inv, ood;
    uint32_t near_x, near_y, near_z;// byte offsets of the near planes' rows inside a wide node
    __device__ __forceinline__ void set(V3 oo, V3 dd) {
        o = oo;
        d = dd;
        inv = v3(safe_rcp(dd.x), safe_rcp(dd.y), safe_rcp(dd.z));
        ood = v3(oo.x * inv.x, oo.y * inv.y, oo.z * inv.z);
        near_x = ((__float_as_uint(dd.x) >> 31u) << 6u);
        near_y = ((__float_as_uint(dd.y) >> 31u) << 6u) + 16u;
        near_z = ((__float_as_uint(dd.z) >> 31u) << 6u) + 32u;
    }
};

// Per-ray traversal state (one per lane on the device).
struct RayState {
    RaySetup cur;
    float tmin, tbest;
    uint32_t node;    // the reference this ray stands on: inner node index, triangle-leaf ref, or kSentinelDone
    uint32_t cur_inst;// instance being walked, ~0u while in the TLAS
    uint32_t best_inst, best_prim;
    float best_u, best_v;
    uint32_t list;    // instance-candidate ring: bits 0..7 = count, bits 8..15 = index of the oldest entry
    uint32_t tlas_sp; // != 0: the TLAS walk is suspended, its entries occupy stack[0, tlas_sp)
};

__device__ __forceinline__ float key_distance(uint32_t key) { return __uint_as_float(key & 0x7ffffffcu); }

// Next reference to visit after `top` (the entry just taken off the stack) turned out to be culled: keep popping until an
// entry survives whose box starts within the (shortened) ray, or the hierarchy's bottom sentinel comes up.
template<typename Mem>
__device__ __forceinline__ uint32_t pop_live(Mem &mem, uint2 top, float tbest) {
    while (key_distance(top.y) > tbest && top.x != kSentinelDone) top = mem.pop();
    return top.x;
}

// One step on an inner node: test the four child boxes; continue with the nearest hit child (ORDERED) or the first one in slot
// order; defer the other hit children in slot order, each with its entry distance; pop when nothing was hit.  TLAS: hit
// children that are leaves are instances - they go to the candidate queue instead (the caller has made sure four slots are
// free).  Straight-line code: stores are predicated, the top of the stack is read whether or not it is needed, and the only
// loop (pop_live) runs when a popped entry is culled.
template<bool ORDERED, bool COUNT, bool TLAS, typename Mem>
__device__ __forceinline__ void inner_step(const DeviceScene &sc, RayState &r, Mem &mem, TraversalCounters &cnt) {
    const char *nb = reinterpret_cast<const char *>(sc.wide_nodes) + static_cast<size_t>(r.node) * (kWideRows * sizeof(float4));
    const float4 nx = __ldg(reinterpret_cast<const float4 *>(nb + r.cur.near_x));
    const float4 fx = __ldg(reinterpret_cast<const float4 *>(nb + (r.cur.near_x ^ 64u)));
    const float4 ny = __ldg(reinterpret_cast<const float4 *>(nb + r.cur.near_y));
    const float4 fy = __ldg(reinterpret_cast<const float4 *>(nb + (r.cur.near_y ^ 64u)));
    const float4 nz = __ldg(reinterpret_cast<const float4 *>(nb + r.cur.near_z));
    const float4 fz = __ldg(reinterpret_cast<const float4 *>(nb + (r.cur.near_z ^ 64u)));
    const float4 rf = __ldg(reinterpret_cast<const float4 *>(nb + 48u));
    if (COUNT) cnt.nodes++;
    const float ix = r.cur.inv.x, iy = r.cur.inv.y, iz = r.cur.inv.z;
    const float ox = -r.cur.ood.x, oy = -r.cur.ood.y, oz = -r.cur.ood.z;
    const float tmin = r.tmin, tbest = r.tbest;
#define LRK_SLAB(C, TN, HIT)                                                                                              \
    const float TN = fmaxf(fmaxf(fmaf(nx.C, ix, ox), fmaf(ny.C, iy, oy)), fmaxf(fmaf(nz.C, iz, oz), tmin));              \
    bool HIT = TN <= fminf(fminf(fmaf(fx.C, ix, ox), fmaf(fy.C, iy, oy)), fminf(fmaf(fz.C, iz, oz), tbest));
    LRK_SLAB(x, tn0, h0)
    LRK_SLAB(y, tn1, h1)
    LRK_SLAB(z, tn2, h2)
    LRK_SLAB(w, tn3, h3)
#undef LRK_SLAB
    const uint32_t r0 = __float_as_uint(rf.x), r1 = __float_as_uint(rf.y), r2 = __float_as_uint(rf.z), r3 = __float_as_uint(rf.w);
    // keys: entry distance (clamped at 0 so that its bit pattern orders like the value) with the slot in the two low bits
    const uint32_t k0 = (__float_as_uint(fmaxf(tn0, 0.f)) & 0x7ffffffcu) | 0u;
    const uint32_t k1 = (__float_as_uint(fmaxf(tn1, 0.f)) & 0x7ffffffcu) | 1u;
    const uint32_t k2 = (__float_as_uint(fmaxf(tn2, 0.f)) & 0x7ffffffcu) | 2u;
    const uint32_t k3 = (__float_as_uint(fmaxf(tn3, 0.f)) & 0x7ffffffcu) | 3u;
    if (TLAS) {
        // hit instances -> candidate queue, in slot order
        uint32_t count = r.list & 0xffu;
        const uint32_t head = (r.list >> 8u) & 0xffu;
        const bool i0 = h0 && (r0 & LRK_BVH_LEAF), i1 = h1 && (r1 & LRK_BVH_LEAF), i2 = h2 && (r2 & LRK_BVH_LEAF), i3 = h3 && (r3 & LRK_BVH_LEAF);
        if (i0) { mem.list_store((head + count) & (kListSize - 1u), r0 & 0x7fffffffu, k0); count++; }
        if (i1) { mem.list_store((head + count) & (kListSize - 1u), r1 & 0x7fffffffu, k1); count++; }
        if (i2) { mem.list_store((head + count) & (kListSize - 1u), r2 & 0x7fffffffu, k2); count++; }
        if (i3) { mem.list_store((head + count) & (kListSize - 1u), r3 & 0x7fffffffu, k3); count++; }
        r.list = (head << 8u) | count;
        h0 = h0 && !i0;
        h1 = h1 && !i1;
        h2 = h2 && !i2;
        h3 = h3 && !i3;
    }
    const uint32_t m0 = h0 ? k0 : 0xffffffffu, m1 = h1 ? k1 : 0xffffffffu, m2 = h2 ? k2 : 0xffffffffu, m3 = h3 ? k3 : 0xffffffffu;
    // closest-hit: the nearest hit child; any-hit: the first hit child in slot order (every hit child has to be looked at unless
    // an occluder turns up first, so there is nothing to gain from ordering)
    const uint32_t first = ORDERED ? min(min(m0, m1), min(m2, m3)) : (h0 ? m0 : h1 ? m1 : h2 ? m2 : m3);
    mem.push4_if(h3 && m3 != first, r3, k3, h2 && m2 != first, r2, k2, h1 && m1 != first, r1, k1, h0 && m0 != first, r0, k0);
    const uint32_t slot = first & 3u;
    const uint32_t nearest = slot == 0u ? r0 : slot == 1u ? r1 : slot == 2u ? r2 : r3;
    const uint2 top = mem.peek();// the bottom sentinel is always there
    if (first != 0xffffffffu) {
        r.node = nearest;
    } else {
        mem.drop();
        r.node = pop_live(mem, top, r.tbest);
    }
}

// ALPHA: stochastic alpha test of every accepted candidate (scenes with non-opaque surfaces only; alpha_skip is in shading.cuh)
__device__ bool alpha_skip(const DeviceScene &sc, uint32_t inst_id, uint32_t prim_id, float bu, float bv);

// The hierarchy the ray was walking is exhausted (its bottom sentinel came up): enter the next instance candidate whose box
// starts within the ray, else resume a suspended TLAS walk, else the ray is finished (returns true).
template<bool COUNT, typename Mem>
__device__ __forceinline__ bool next_hierarchy(const DeviceScene &sc, RayState &r, Mem &mem, TraversalCounters &cnt) {
    for (;;) {
        const uint32_t count = r.list & 0xffu;
        if (count != 0u) {
            const uint32_t head = (r.list >> 8u) & 0xffu;
            const uint2 e = mem.list_load(head);
            r.list = (((head + 1u) & (kListSize - 1u)) << 8u) | (count - 1u);
            if (key_distance(e.y) > r.tbest) continue;// the instance's box starts behind the hit found meanwhile
            r.cur_inst = e.x;
            if (COUNT) cnt.xforms++;
            const float4 *x = sc.inst_xform + static_cast<size_t>(e.x) * 4u;
            float4 m0 = __ldg(x + 0), m1 = __ldg(x + 1), m2 = __ldg(x + 2), m3 = __ldg(x + 3);
            V3 wo, wd;
            mem.world_load(wo, wd);
            V3 oo = v3(fmaf(m0.x, wo.x, fmaf(m0.y, wo.y, fmaf(m0.z, wo.z, m0.w))),
                       fmaf(m1.x, wo.x, fmaf(m1.y, wo.y, fmaf(m1.z, wo.z, m1.w))),
                       fmaf(m2.x, wo.x, fmaf(m2.y, wo.y, fmaf(m2.z, wo.z, m2.w))));
            V3 dd = v3(fmaf(m0.x, wd.x, fmaf(m0.y, wd.y, m0.z * wd.z)),
                       fmaf(m1.x, wd.x, fmaf(m1.y, wd.y, m1.z * wd.z)),
                       fmaf(m2.x, wd.x, fmaf(m2.y, wd.y, m2.z * wd.z)));
            r.cur.set(oo, dd);
            mem.push_if(true, kSentinelDone, 0u);
            r.node = __float_as_uint(m3.x);
            return false;
        }
        if (r.tlas_sp != 0u) {// back to the suspended TLAS walk: the node it stopped at is on top of the stack
            V3 wo, wd;
            mem.world_load(wo, wd);
            r.cur.set(wo, wd);
            r.cur_inst = ~0u;
            r.tlas_sp = 0u;
            r.node = mem.pop().x;
            return false;
        }
        return true;
    }
}

// A TLAS walk may only expand a node while four candidate slots are free; otherwise it parks the node on its stack and hands
// over to the queued instances (next_hierarchy comes back here when they are done).
template<typename Mem>
__device__ __forceinline__ bool tlas_suspend_if_full(RayState &r, Mem &mem) {
    if ((r.list & 0xffu) + 4u <= kListSize) return false;
    mem.push_if(true, r.node, 0u);
    r.tlas_sp = static_cast<uint32_t>(mem.depth());
    r.node = kSentinelDone;
    return true;
}

// One step on a leaf-like reference (bit 31 set): a triangle range of the instance being walked, or the bottom sentinel of an
// exhausted hierarchy.  Returns true when the ray is finished.
template<bool ANY_HIT, bool COUNT, bool ALPHA, typename Mem>
__device__ __forceinline__ bool leaf_step(const DeviceScene &sc, RayState &r, Mem &mem, TraversalCounters &cnt) {
    const uint32_t node = r.node;
    if (node == kSentinelDone) return next_hierarchy<COUNT>(sc, r, mem, cnt);
    if (node != LRK_BVH_EMPTY) {
        const uint32_t first = node & 0x0fffffffu;
        const uint32_t count = ((node >> 28u) & 7u) + 1u;
        for (uint32_t k = 0; k < count; k++) {
            const float4 *tv = sc.tri_verts + static_cast<size_t>(first + k) * 3u;
            float4 a = __ldg(tv + 0), b = __ldg(tv + 1), c = __ldg(tv + 2);
            if (COUNT) cnt.tris++;
            V3 p0 = v3(a.x, a.y, a.z), p1 = v3(b.x, b.y, b.z), p2 = v3(c.x, c.y, c.z);
            V3 e1 = p1 - p0, e2 = p2 - p0;
            V3 pvec = fcross(r.cur.d, e2);
            float det = fdot(e1, pvec);
            if (!(det != 0.0f)) continue;
            float inv_det = 1.0f / det;
            V3 tvec = r.cur.o - p0;
            float u = fdot(tvec, pvec) * inv_det;
            if (!(u >= 0.0f && u <= 1.0f)) continue;
            V3 qvec = fcross(tvec, e1);
            float v = fdot(r.cur.d, qvec) * inv_det;
            if (!(v >= 0.0f && u + v <= 1.0f)) continue;
            float t = fdot(e2, qvec) * inv_det;
            const uint32_t prim = __float_as_uint(a.w);
            // exact ties in t (coincident faces of two shapes) go to the lower (instance, primitive): the result does not depend
            // on the visiting order, which differs between this 4-wide walk, the oracle's BVH2 walk and a brute-force loop
            const bool tie = t == r.tbest && r.best_inst != ~0u && (r.cur_inst < r.best_inst || (r.cur_inst == r.best_inst && prim < r.best_prim));
            if (!(t > r.tmin && (t < r.tbest || tie))) continue;
            if (ALPHA) {// on_surface_candidate: commit only if not skipped (geometry.cpp:248-279)
                if (alpha_skip(sc, r.cur_inst, prim, u, v)) continue;
            }
            r.tbest = t;
            r.best_inst = r.cur_inst;
            r.best_prim = prim;
            r.best_u = u;
            r.best_v = v;
            if (ANY_HIT) return true;
        }
    }
    r.node = pop_live(mem, mem.pop(), r.tbest);
    return false;
}

template<typename Mem>
__device__ __forceinline__ void start_ray(const DeviceScene &sc, RayState &r, Mem &mem, float4 o, float4 d) {
    const V3 wo = v3(o.x, o.y, o.z), wd = v3(d.x, d.y, d.z);
    r.cur.set(wo, wd);
    mem.world_save(wo, wd);
    r.tmin = o.w;
    r.tbest = d.w;
    r.best_inst = ~0u;
    r.best_prim = ~0u;
    r.best_u = r.best_v = 0.f;
    r.cur_inst = ~0u;
    r.list = 0u;
    r.tlas_sp = 0u;
    mem.reset();
    mem.push_if(true, kSentinelDone, 0u);
    r.node = sc.tlas_root;
}

#ifdef __CUDACC__

constexpr int kTraceBlock = 256;// threads per block of every kernel that calls trace_queue

// Per-lane traversal memory.  The struct holds scalars only (shared-window addresses, depth, pointers) so that it lives in
// registers; shared memory is addressed through its 32-bit window address with explicit ld/st.shared (a pointer member would
// degrade to generic loads and stores).  Columns are lane-interleaved: entry k of thread t at base + (k * kTraceBlock + t) * size.
struct LaneMem {
    uint32_t stack;// (ref, key) stack, kSmemStack entries of 8 bytes
    uint32_t list; // instance candidates, kListSize entries of 8 bytes
    uint32_t world;// world-space ray: {o.x, o.y, o.z, d.x}, {d.y, d.z, -, -}
    uint2 *local;  // kLocalStack further stack entries in local memory
    int sp;
    uint32_t *overflow;
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ int depth() const { return sp; }
    __device__ __forceinline__ void push_if(bool valid, uint32_t ref, uint32_t key) {
        if (valid) {
            if (sp < kSmemStack) {
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stack + static_cast<uint32_t>(sp) * (kTraceBlock * 8u)), "r"(ref), "r"(key) : "memory");
            } else if (sp < kSmemStack + kLocalStack) {
                local[sp - kSmemStack] = make_uint2(ref, key);
            } else {
                *overflow = 1u;// reported by lrk_render / lrk_trace as an error (hierarchy too deep); the entry is lost
            }
            sp++;
        }
    }
    // up to four entries at once, in argument order: predicated shared stores when all of them fit into the shared part
    __device__ __forceinline__ void push4_if(bool pa, uint32_t ra, uint32_t ka, bool pb, uint32_t rb, uint32_t kb,
                                             bool pc, uint32_t rc, uint32_t kc, bool pd, uint32_t rd, uint32_t kd) {
        const int sa = sp, sb = sa + (pa ? 1 : 0), sc = sb + (pb ? 1 : 0), sd = sc + (pc ? 1 : 0), end = sd + (pd ? 1 : 0);
        if (end <= kSmemStack) {
            if (pa) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stack + static_cast<uint32_t>(sa) * (kTraceBlock * 8u)), "r"(ra), "r"(ka) : "memory");
            if (pb) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stack + static_cast<uint32_t>(sb) * (kTraceBlock * 8u)), "r"(rb), "r"(kb) : "memory");
            if (pc) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stack + static_cast<uint32_t>(sc) * (kTraceBlock * 8u)), "r"(rc), "r"(kc) : "memory");
            if (pd) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stack + static_cast<uint32_t>(sd) * (kTraceBlock * 8u)), "r"(rd), "r"(kd) : "memory");
            sp = end;
        } else {
            push_if(pa, ra, ka);
            push_if(pb, rb, kb);
            push_if(pc, rc, kc);
            push_if(pd, rd, kd);
        }
    }
    __device__ __forceinline__ uint2 at(int i) const {
        if (i >= kSmemStack) return local[min(i, kSmemStack + kLocalStack - 1) - kSmemStack];
        uint2 e;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(stack + static_cast<uint32_t>(i) * (kTraceBlock * 8u)) : "memory");
        return e;
    }
    __device__ __forceinline__ uint2 peek() const { return at(sp - 1); }
    __device__ __forceinline__ void drop() { --sp; }
    __device__ __forceinline__ uint2 pop() {
        --sp;
        return at(sp);
    }
    __device__ __forceinline__ void list_store(uint32_t slot, uint32_t inst, uint32_t key) {
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(list + slot * (kTraceBlock * 8u)), "r"(inst), "r"(key) : "memory");
    }
    __device__ __forceinline__ uint2 list_load(uint32_t slot) const {
        uint2 e;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(list + slot * (kTraceBlock * 8u)) : "memory");
        return e;
    }
    __device__ __forceinline__ void world_save(V3 o, V3 d) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(world), "f"(o.x), "f"(o.y), "f"(o.z), "f"(d.x) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(world + kTraceBlock * 16u), "f"(d.y), "f"(d.z) : "memory");
    }
    __device__ __forceinline__ void world_load(V3 &o, V3 &d) const {
        float4 a;
        float2 b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(world) : "memory");
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(b.x), "=f"(b.y) : "r"(world + kTraceBlock * 16u) : "memory");
        o = v3(a.x, a.y, a.z);
        d = v3(a.w, b.x, b.y);
    }
};

struct TraceShared {
    uint2 stack[kSmemStack][kTraceBlock];
    uint2 list[kListSize][kTraceBlock];
    float4 world[2][kTraceBlock];// row 1 uses its first 8 bytes only
};

// Traces rays [0, n) of the queue (ray_o / ray_d). `cursor` is a zero-initialised device counter private to
// this launch. `sink(finished, ray_index, hit)` is called by ALL 32 lanes together (warp-convergent) after every
// leaf phase; a lane passes finished = true exactly once per ray, with
// hit = {inst, prim, bary.u bits, bary.v bits} (miss <=> inst == ~0u; ANY_HIT: first hit found).
template<bool ANY_HIT, bool COUNT, int STRIDE, bool ALPHA = false, typename Sink>
__device__ __forceinline__ void trace_queue(const DeviceScene &sc, const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d,
                                            uint32_t n, uint32_t *cursor, TraversalCounters &cnt, Sink &&sink) {
    __shared__ TraceShared shared;
#ifdef LRK_ANYHIT_ORDERED
    constexpr bool ORDERED = true;
#else
    constexpr bool ORDERED = !ANY_HIT;
#endif
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint2 deep_stack[kLocalStack];
    LaneMem mem;
    mem.stack = static_cast<uint32_t>(__cvta_generic_to_shared(&shared.stack[0][threadIdx.x]));
    mem.list = static_cast<uint32_t>(__cvta_generic_to_shared(&shared.list[0][threadIdx.x]));
    mem.world = static_cast<uint32_t>(__cvta_generic_to_shared(&shared.world[0][threadIdx.x]));
    mem.local = deep_stack;
    mem.overflow = sc.traversal_overflow;
    mem.reset();
    RayState r;
    start_ray(sc, r, mem, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 1.f, 0.f));
    r.node = kSentinelDone;
    bool active = false, exhausted = false;
    uint32_t ray_index = 0u;
    for (;;) {
        // ---- refill idle lanes from the queue ------------------------------------------------------------
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle != 0u && !exhausted) {
            const uint32_t want = __popc(idle);
            uint32_t base = 0u;
            if (lane == 0u) base = atomicAdd(cursor, want);
            base = __shfl_sync(0xffffffffu, base, 0);
            const uint32_t pos = base + __popc(idle & lane_lt);
            if (!active && pos < n) {
                float4 o = ray_o[static_cast<size_t>(pos) * STRIDE], d = ray_d[static_cast<size_t>(pos) * STRIDE];
                start_ray(sc, r, mem, o, d);
                ray_index = pos;
                active = true;
            }
            exhausted = base + want >= n;
        }
        if (!__any_sync(0xffffffffu, active)) break;
        // ---- traverse until too few lanes are busy (and fresh rays are available) ---------------------------
        for (;;) {
            // TLAS phase: the lanes that walk the top level (freshly fetched rays, rarely a resumed walk) step together
            for (;;) {
                const bool tlas = active && r.cur_inst == ~0u && !(r.node & LRK_BVH_LEAF);
                if (!__any_sync(0xffffffffu, tlas)) break;
                if (tlas && !tlas_suspend_if_full(r, mem)) inner_step<ORDERED, COUNT, true>(sc, r, mem, cnt);
            }
            // inner phase: step the lanes that stand on an inner node; lanes that have reached a leaf wait, but only while at
            // least `inner_min` lanes still have inner work (inner_min = 1 is the classic while-while loop, 32 is if-if)
            for (;;) {
                const bool inner = !(r.node & LRK_BVH_LEAF);
                const uint32_t n_inner = __popc(__ballot_sync(0xffffffffu, inner));
                if (n_inner == 0u) break;
                if (n_inner < sc.inner_min && __any_sync(0xffffffffu, active && !inner)) break;
                if (inner) inner_step<ORDERED, COUNT, false>(sc, r, mem, cnt);
            }
            bool finished = false;
            if (active && (r.node & LRK_BVH_LEAF)) finished = leaf_step<ANY_HIT, COUNT, ALPHA>(sc, r, mem, cnt);
            sink(finished, ray_index, make_uint4(r.best_inst, r.best_prim, __float_as_uint(r.best_u), __float_as_uint(r.best_v)));
            if (finished) {
                active = false;
                r.node = kSentinelDone;// idle lanes skip the inner loops
            }
            const uint32_t busy = __popc(__ballot_sync(0xffffffffu, active));
            if (busy == 0u || (busy < sc.refill_below && !exhausted)) break;
        }
    }
}

#endif// __CUDACC__

}// namespace lrk
