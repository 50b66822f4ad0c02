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
//   * closest-hit: the nearest hit child is visited first, the others are deferred in slot order; any-hit: slot order, the
//     first accepted triangle ends the ray;
//   * Moeller-Trumbore in object space with fma dot/cross; accept tmin < t < t_best, u,v >= 0, u+v <= 1; an exact tie
//     t == t_best goes to the lower (instance, primitive), so that the result does not depend on the visiting order;
//   * instance entry transforms the ray by world_to_object without renormalising d (t is shared); leaving the instance (an exit
//     sentinel on the stack) brings the world-space ray back from shared memory.
//
// SIMT design.  The kernel is bound by instruction issue at partial warp width, not by memory (the hierarchy is L1 / L2
// resident): with one ray per lane about a third of a warp's lanes have work in any issued instruction, because rays are in
// different phases (descending, at a triangle leaf, entering an instance, finished).  So every lane owns kRaysPerLane ray SLOTS
// whose state lives in shared memory (lane-interleaved rows: conflict-free), and in each step of the descent loop a lane works
// on whichever of its rays stands on an inner node: the probability that a lane has something to do goes from p to
// 1 - (1 - p)^K.  Leaf work (triangle tests, instance entry / exit, finished rays) is done in a separate phase for all slots
// that wait for it; rays are fetched from the queue in batches through a per-launch atomic cursor (persistent warps).
// The per-ray rules do not depend on any of this: the step functions below are also compiled for the host and checked against
// the oracle ray by ray (tests/test_device_traversal_on_host.py).
#pragma once
#include "scene.cuh"
#include "wide_bvh.cuh"

namespace lrk {

constexpr uint32_t kSentinelDone = 0xfffffffdu;// bottom of the stack: the ray is finished
constexpr uint32_t kSentinelExit = 0xfffffffeu;// pushed on instance entry: back to the TLAS
constexpr uint32_t kSlotFree = 0xfffffffcu;    // node value of a slot that holds no ray
constexpr int kSmemStack = 8;  // stack entries per ray kept in shared memory
constexpr int kLocalStack = 56;// further entries in local memory (a 4-wide tree defers at most 3 children per level)
constexpr int kRefillBelow = 16;// fetch new rays when fewer than this many rays per ray slot index are live in the warp
constexpr int kInnerMin = 8;   // leave the descent phase when fewer lanes than this can still descend
#ifndef LRK_RAYS_PER_LANE
#define LRK_RAYS_PER_LANE 2
#endif
constexpr int kRaysPerLane = LRK_RAYS_PER_LANE;

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

__device__ __forceinline__ uint32_t near_offsets(V3 d) {
    return ((__float_as_uint(d.x) >> 31u) << 6u) | ((((__float_as_uint(d.y) >> 31u) << 6u) + 16u) << 8u) |
           ((((__float_as_uint(d.z) >> 31u) << 6u) + 32u) << 16u);
}

// What the descent needs of a ray in the space of the hierarchy being walked: 1/d, -o/d, the ray interval, and where to find
// the near planes of a wide node (byte offsets of the rows, from the signs of d).
struct RayHot {
    float ix, iy, iz, tmin;
    float ox, oy, oz, tbest;// ox = -(o.x / d.x) ...
    uint32_t node;    // the reference this ray stands on: inner node index, leaf ref, or a sentinel
    int sp;           // stack depth
    uint32_t near;    // byte 0 / 1 / 2: row offset of the near x / y / z planes (0x00 or 0x40, + 0x10 * axis)
    uint32_t cur_inst;// instance being walked, ~0u while in the TLAS
    __device__ __forceinline__ void set_ray(V3 o, V3 d) {
        ix = safe_rcp(d.x);
        iy = safe_rcp(d.y);
        iz = safe_rcp(d.z);
        ox = -(o.x * ix);
        oy = -(o.y * iy);
        oz = -(o.z * iz);
        near = near_offsets(d);
    }
};

// The rest of a ray's state: origin / direction in the current space (triangle tests), the best hit, the queue position.
struct RayCold {
    V3 o, d;
    float best_u, best_v;
    uint32_t best_inst, best_prim;
    uint32_t ray_index;
};

// One step on an inner node: test the four child boxes; continue with the nearest hit child (ORDERED) or the first one in slot
// order; defer the other hit children in slot order; pop when nothing was hit.  Straight-line code: stores are predicated.
template<bool ORDERED, bool COUNT, typename Stack>
__device__ __forceinline__ void inner_step(const DeviceScene &sc, RayHot &r, Stack &stack, TraversalCounters &cnt) {
    const char *nb = reinterpret_cast<const char *>(sc.wide_nodes) + static_cast<size_t>(r.node) * (kWideRows * sizeof(float4));
    const uint32_t near_x = r.near & 0xffu, near_y = (r.near >> 8u) & 0xffu, near_z = (r.near >> 16u) & 0xffu;
    const float4 nx = __ldg(reinterpret_cast<const float4 *>(nb + near_x));
    const float4 fx = __ldg(reinterpret_cast<const float4 *>(nb + (near_x ^ 64u)));
    const float4 ny = __ldg(reinterpret_cast<const float4 *>(nb + near_y));
    const float4 fy = __ldg(reinterpret_cast<const float4 *>(nb + (near_y ^ 64u)));
    const float4 nz = __ldg(reinterpret_cast<const float4 *>(nb + near_z));
    const float4 fz = __ldg(reinterpret_cast<const float4 *>(nb + (near_z ^ 64u)));
    const float4 rf = __ldg(reinterpret_cast<const float4 *>(nb + 48u));
    if (COUNT) cnt.nodes++;
#define LRK_SLAB(C, TN, HIT)                                                                                                     \
    const float TN = fmaxf(fmaxf(fmaf(nx.C, r.ix, r.ox), fmaf(ny.C, r.iy, r.oy)), fmaxf(fmaf(nz.C, r.iz, r.oz), r.tmin));        \
    const bool HIT = TN <= fminf(fminf(fmaf(fx.C, r.ix, r.ox), fmaf(fy.C, r.iy, r.oy)), fminf(fmaf(fz.C, r.iz, r.oz), r.tbest));
    LRK_SLAB(x, tn0, h0)
    LRK_SLAB(y, tn1, h1)
    LRK_SLAB(z, tn2, h2)
    LRK_SLAB(w, tn3, h3)
#undef LRK_SLAB
    const uint32_t r0 = __float_as_uint(rf.x), r1 = __float_as_uint(rf.y), r2 = __float_as_uint(rf.z), r3 = __float_as_uint(rf.w);
    // keys: entry distance bits (non-negative floats order like their bit patterns; a negative entry distance - only possible for
    // rays with tmin < 0 - merely spoils the ORDER, never a result) with the slot in the two low bits; a missed child is ~0
    const uint32_t k0 = h0 ? ((__float_as_uint(tn0) & 0x7ffffffcu) | 0u) : 0xffffffffu;
    const uint32_t k1 = h1 ? ((__float_as_uint(tn1) & 0x7ffffffcu) | 1u) : 0xffffffffu;
    const uint32_t k2 = h2 ? ((__float_as_uint(tn2) & 0x7ffffffcu) | 2u) : 0xffffffffu;
    const uint32_t k3 = h3 ? ((__float_as_uint(tn3) & 0x7ffffffcu) | 3u) : 0xffffffffu;
    // closest-hit: the nearest hit child; any-hit: the first hit child in slot order (every hit child has to be looked at unless
    // an occluder turns up first, so there is nothing to gain from ordering)
    const uint32_t first = ORDERED ? min(min(k0, k1), min(k2, k3)) : (h0 ? k0 : h1 ? k1 : h2 ? k2 : k3);
    stack.push4_if(r.sp, h3 && k3 != first, r3, h2 && k2 != first, r2, h1 && k1 != first, r1, h0 && k0 != first, r0);
    const uint32_t slot = first & 3u;
    const uint32_t nearest = slot == 0u ? r0 : slot == 1u ? r1 : slot == 2u ? r2 : r3;
    if (first != 0xffffffffu) r.node = nearest;
    else r.node = stack.pop(r.sp);
}

// ALPHA: stochastic alpha test of every accepted candidate (scenes with non-opaque surfaces only; alpha_skip is in shading.cuh)
__device__ bool alpha_skip(const DeviceScene &sc, uint32_t inst_id, uint32_t prim_id, float bu, float bv);

// One step on a leaf-like reference (bit 31 set): exit sentinel, instance (TLAS leaf) or triangle range (BLAS leaf).
// `world` parks the world-space ray while it is inside an instance.  Returns true when the ray is finished (bottom sentinel, or
// an any-hit ray that found an occluder).
template<bool ANY_HIT, bool COUNT, bool ALPHA, typename Stack, typename World>
__device__ __forceinline__ bool leaf_step(const DeviceScene &sc, RayHot &r, RayCold &c, Stack &stack, World &world, TraversalCounters &cnt) {
    const uint32_t node = r.node;
    if (node == kSentinelDone) return true;
    if (node == kSentinelExit) {
        world.load(r, c);
        r.cur_inst = ~0u;
        r.node = stack.pop(r.sp);
        return false;
    }
    if (node == LRK_BVH_EMPTY) {
        r.node = stack.pop(r.sp);
        return false;
    }
    if (r.cur_inst != ~0u) {
        const uint32_t first = node & 0x0fffffffu;
        const uint32_t count = ((node >> 28u) & 7u) + 1u;
        for (uint32_t k = 0; k < count; k++) {
            const float4 *tv = sc.tri_verts + static_cast<size_t>(first + k) * 3u;
            float4 a = __ldg(tv + 0), b = __ldg(tv + 1), cc = __ldg(tv + 2);
            if (COUNT) cnt.tris++;
            V3 p0 = v3(a.x, a.y, a.z), p1 = v3(b.x, b.y, b.z), p2 = v3(cc.x, cc.y, cc.z);
            V3 e1 = p1 - p0, e2 = p2 - p0;
            V3 pvec = fcross(c.d, e2);
            float det = fdot(e1, pvec);
            if (!(det != 0.0f)) continue;
            float inv_det = 1.0f / det;
            V3 tvec = c.o - p0;
            float u = fdot(tvec, pvec) * inv_det;
            if (!(u >= 0.0f && u <= 1.0f)) continue;
            V3 qvec = fcross(tvec, e1);
            float v = fdot(c.d, qvec) * inv_det;
            if (!(v >= 0.0f && u + v <= 1.0f)) continue;
            float t = fdot(e2, qvec) * inv_det;
            const uint32_t prim = __float_as_uint(a.w);
            // exact ties in t (coincident faces of two shapes) go to the lower (instance, primitive): the result does not depend
            // on the visiting order, which differs between this 4-wide walk, the oracle's BVH2 walk and a brute-force loop
            const bool tie = t == r.tbest && c.best_inst != ~0u && (r.cur_inst < c.best_inst || (r.cur_inst == c.best_inst && prim < c.best_prim));
            if (!(t > r.tmin && (t < r.tbest || tie))) continue;
            if (ALPHA) {// on_surface_candidate: commit only if not skipped (geometry.cpp:248-279)
                if (alpha_skip(sc, r.cur_inst, prim, u, v)) continue;
            }
            r.tbest = t;
            c.best_inst = r.cur_inst;
            c.best_prim = prim;
            c.best_u = u;
            c.best_v = v;
            if (ANY_HIT) return true;
        }
        r.node = stack.pop(r.sp);
        return false;
    }
    // TLAS leaf: enter the instance
    r.cur_inst = node & 0x7fffffffu;
    if (COUNT) cnt.xforms++;
    const float4 *x = sc.inst_xform + static_cast<size_t>(r.cur_inst) * 4u;
    float4 m0 = __ldg(x + 0), m1 = __ldg(x + 1), m2 = __ldg(x + 2), m3 = __ldg(x + 3);
    world.save(r, c);
    const V3 wo = c.o, wd = c.d;
    c.o = v3(fmaf(m0.x, wo.x, fmaf(m0.y, wo.y, fmaf(m0.z, wo.z, m0.w))),
             fmaf(m1.x, wo.x, fmaf(m1.y, wo.y, fmaf(m1.z, wo.z, m1.w))),
             fmaf(m2.x, wo.x, fmaf(m2.y, wo.y, fmaf(m2.z, wo.z, m2.w))));
    c.d = v3(fmaf(m0.x, wd.x, fmaf(m0.y, wd.y, m0.z * wd.z)),
             fmaf(m1.x, wd.x, fmaf(m1.y, wd.y, m1.z * wd.z)),
             fmaf(m2.x, wd.x, fmaf(m2.y, wd.y, m2.z * wd.z)));
    r.set_ray(c.o, c.d);
    stack.push(r.sp, kSentinelExit);
    r.node = __float_as_uint(m3.x);
    return false;
}

template<typename Stack>
__device__ __forceinline__ void start_ray(const DeviceScene &sc, RayHot &r, RayCold &c, Stack &stack, float4 o, float4 d, uint32_t ray_index) {
    c.o = v3(o.x, o.y, o.z);
    c.d = v3(d.x, d.y, d.z);
    c.best_inst = ~0u;
    c.best_prim = ~0u;
    c.best_u = c.best_v = 0.f;
    c.ray_index = ray_index;
    r.set_ray(c.o, c.d);
    r.tmin = o.w;
    r.tbest = d.w;
    r.cur_inst = ~0u;
    r.sp = 0;
    stack.push(r.sp, kSentinelDone);
    r.node = sc.tlas_root;
}

#ifdef __CUDACC__

constexpr int kTraceBlock = 128;// threads per block of every kernel that calls trace_queue

// Shared-memory image of one ray slot per thread; rows are lane-interleaved (row k of thread t at (k * kTraceBlock + t) * 16).
//   rows 0..2  hot   {ix, iy, iz, tmin} {ox, oy, oz, tbest} {node, sp, near, cur_inst}
//   rows 3..5  cold  {o.xyz, d.x} {d.y, d.z, best_u, best_v} {best_inst, best_prim, ray_index, -}
//   rows 6..8  world {o.xyz, d.x} {d.y, d.z, ix, iy} {iz, ox, oy, oz}   (the world-space ray while inside an instance)
//   rows 9..   stack kSmemStack x u32, four entries per row
constexpr uint32_t kSlotRows = 9u + kSmemStack / 4u;
constexpr uint32_t kRowStride = kTraceBlock * 16u;
struct TraceShared {
    float4 rows[kRaysPerLane][kSlotRows][kTraceBlock];
};

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

// One ray slot of this thread: shared-window base address + its overflow stack in local memory.  Scalars only, so that the
// struct lives in registers.
struct Slot {
    uint32_t base; // shared-window address of row 0 of this thread's column
    uint32_t *deep;// kLocalStack further stack entries (local memory)
    uint32_t *overflow;
    __device__ __forceinline__ uint32_t stack_addr(int i) const {
        return base + (9u + (static_cast<uint32_t>(i) >> 2u)) * kRowStride + (static_cast<uint32_t>(i) & 3u) * 4u;
    }
    __device__ __forceinline__ void push(int &sp, uint32_t ref) const {
        if (sp < kSmemStack) sts32(stack_addr(sp), ref);
        else if (sp < kSmemStack + kLocalStack) deep[sp - kSmemStack] = ref;
        else *overflow = 1u;// reported by lrk_render / lrk_trace as an error (hierarchy too deep); the entry is lost
        sp++;
    }
    // up to four entries at once, in argument order: predicated shared stores when all of them fit into the shared part
    __device__ __forceinline__ void push4_if(int &sp, bool pa, uint32_t ra, bool pb, uint32_t rb, bool pc, uint32_t rc, bool pd, uint32_t rd) const {
        const int sa = sp, sb = sa + (pa ? 1 : 0), sc = sb + (pb ? 1 : 0), sd = sc + (pc ? 1 : 0), end = sd + (pd ? 1 : 0);
        if (end <= kSmemStack) {
            if (pa) sts32(stack_addr(sa), ra);
            if (pb) sts32(stack_addr(sb), rb);
            if (pc) sts32(stack_addr(sc), rc);
            if (pd) sts32(stack_addr(sd), rd);
            sp = end;
        } else {
            if (pa) push(sp, ra);
            if (pb) push(sp, rb);
            if (pc) push(sp, rc);
            if (pd) push(sp, rd);
        }
    }
    __device__ __forceinline__ uint32_t pop(int &sp) const {
        --sp;
        if (sp >= kSmemStack) return deep[min(sp, kSmemStack + kLocalStack - 1) - kSmemStack];
        return lds32(stack_addr(sp));
    }
    __device__ __forceinline__ void load_hot(RayHot &r) const {
        const float4 a = lds128(base), b = lds128(base + kRowStride), c = lds128(base + 2u * kRowStride);
        r.ix = a.x; r.iy = a.y; r.iz = a.z; r.tmin = a.w;
        r.ox = b.x; r.oy = b.y; r.oz = b.z; r.tbest = b.w;
        r.node = __float_as_uint(c.x);
        r.sp = static_cast<int>(__float_as_uint(c.y));
        r.near = __float_as_uint(c.z);
        r.cur_inst = __float_as_uint(c.w);
    }
    __device__ __forceinline__ void store_node_sp(const RayHot &r) const { sts64(base + 2u * kRowStride, r.node, static_cast<uint32_t>(r.sp)); }
    __device__ __forceinline__ void store_hot(const RayHot &r) const {
        sts128(base, r.ix, r.iy, r.iz, r.tmin);
        sts128(base + kRowStride, r.ox, r.oy, r.oz, r.tbest);
        sts128(base + 2u * kRowStride, __uint_as_float(r.node), __uint_as_float(static_cast<uint32_t>(r.sp)), __uint_as_float(r.near),
               __uint_as_float(r.cur_inst));
    }
    __device__ __forceinline__ void load_cold(RayCold &c) const {
        const float4 a = lds128(base + 3u * kRowStride), b = lds128(base + 4u * kRowStride), d = lds128(base + 5u * kRowStride);
        c.o = v3(a.x, a.y, a.z);
        c.d = v3(a.w, b.x, b.y);
        c.best_u = b.z;
        c.best_v = b.w;
        c.best_inst = __float_as_uint(d.x);
        c.best_prim = __float_as_uint(d.y);
        c.ray_index = __float_as_uint(d.z);
    }
    __device__ __forceinline__ void store_cold(const RayCold &c) const {
        sts128(base + 3u * kRowStride, c.o.x, c.o.y, c.o.z, c.d.x);
        sts128(base + 4u * kRowStride, c.d.y, c.d.z, c.best_u, c.best_v);
        sts128(base + 5u * kRowStride, __uint_as_float(c.best_inst), __uint_as_float(c.best_prim), __uint_as_float(c.ray_index), 0.f);
    }
    // the world-space ray, parked while the ray is inside an instance (the near-plane offsets come back from the signs of d)
    __device__ __forceinline__ void save(const RayHot &r, const RayCold &c) const {
        sts128(base + 6u * kRowStride, c.o.x, c.o.y, c.o.z, c.d.x);
        sts128(base + 7u * kRowStride, c.d.y, c.d.z, r.ix, r.iy);
        sts128(base + 8u * kRowStride, r.iz, r.ox, r.oy, r.oz);
    }
    __device__ __forceinline__ void load(RayHot &r, RayCold &c) const {
        const float4 a = lds128(base + 6u * kRowStride), b = lds128(base + 7u * kRowStride), d = lds128(base + 8u * kRowStride);
        c.o = v3(a.x, a.y, a.z);
        c.d = v3(a.w, b.x, b.y);
        r.ix = b.z; r.iy = b.w; r.iz = d.x;
        r.ox = d.y; r.oy = d.z; r.oz = d.w;
        r.near = near_offsets(c.d);
    }
};

// Traces rays [0, n) of the queue (ray_o / ray_d). `cursor` is a zero-initialised device counter private to
// this launch. `sink(finished, ray_index, hit)` is called by ALL 32 lanes together (warp-convergent), once per ray slot index
// after every leaf phase; a lane passes finished = true exactly once per ray, with
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
    constexpr int K = kRaysPerLane;
    constexpr uint32_t kSlotStride = kSlotRows * kRowStride;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t deep_stack[K][kLocalStack];
    const uint32_t base0 = static_cast<uint32_t>(__cvta_generic_to_shared(&shared.rows[0][0][threadIdx.x]));
    uint32_t node[K];// each slot's current reference, mirrored in registers: scheduling needs no shared loads
#pragma unroll
    for (int s = 0; s < K; s++) node[s] = kSlotFree;
    bool exhausted = false;
    auto make_slot = [&](int s) {
        Slot sl;
        sl.base = base0 + static_cast<uint32_t>(s) * kSlotStride;
        sl.deep = deep_stack[s];
        sl.overflow = sc.traversal_overflow;
        return sl;
    };
    for (;;) {
        // ---- fetch new rays into the free slots: one atomic per warp, positions handed out slot index by slot index -----
        if (!exhausted) {
            uint32_t free_mask[K], want = 0u;
#pragma unroll
            for (int s = 0; s < K; s++) {
                free_mask[s] = __ballot_sync(0xffffffffu, node[s] == kSlotFree);
                want += __popc(free_mask[s]);
            }
            uint32_t base = 0u;
            if (lane == 0u && want != 0u) base = atomicAdd(cursor, want);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint32_t offset = 0u;
#pragma unroll
            for (int s = 0; s < K; s++) {
                const uint32_t pos = base + offset + __popc(free_mask[s] & lane_lt);
                if (node[s] == kSlotFree && pos < n) {
                    RayHot r;
                    RayCold c;
                    Slot sl = make_slot(s);
                    float4 o = ray_o[static_cast<size_t>(pos) * STRIDE], d = ray_d[static_cast<size_t>(pos) * STRIDE];
                    start_ray(sc, r, c, sl, o, d, pos);
                    sl.store_hot(r);
                    sl.store_cold(c);
                    node[s] = r.node;
                }
                offset += __popc(free_mask[s]);
            }
            exhausted = base + want >= n;
        }
        {
            bool any_live = false;
#pragma unroll
            for (int s = 0; s < K; s++) any_live = any_live || node[s] != kSlotFree;
            if (!__any_sync(0xffffffffu, any_live)) break;
        }
        // ---- traverse until too many slots are free (and fresh rays are available) ---------------------------
        for (;;) {
            // descent phase: every lane steps ONE of its rays that stands on an inner node; rays that have reached a leaf wait, but
            // only while at least `inner_min` lanes still have a ray to descend with
            for (;;) {
                int sel = -1;
                bool waiting = false;
#pragma unroll
                for (int s = K - 1; s >= 0; s--) {
                    if (!(node[s] & LRK_BVH_LEAF)) sel = s;
                    else if (node[s] != kSlotFree) waiting = true;
                }
                const uint32_t n_inner = __popc(__ballot_sync(0xffffffffu, sel >= 0));
                if (n_inner == 0u) break;
                if (n_inner < sc.inner_min && __any_sync(0xffffffffu, waiting)) break;
                if (sel >= 0) {
                    Slot sl = make_slot(sel);
                    RayHot r;
                    sl.load_hot(r);
                    inner_step<ORDERED, COUNT>(sc, r, sl, cnt);
                    sl.store_node_sp(r);
#pragma unroll
                    for (int s = 0; s < K; s++)
                        if (s == sel) node[s] = r.node;
                }
            }
            // leaf phase: one round per slot index (one copy of the code, K trips)
            uint32_t live = 0u;
#pragma unroll 1
            for (int s = 0; s < K; s++) {
                uint32_t nd = kSlotFree;
#pragma unroll
                for (int q = 0; q < K; q++)
                    if (q == s) nd = node[q];
                bool finished = false;
                uint4 hit = make_uint4(~0u, ~0u, 0u, 0u);
                uint32_t ray_index = 0u;
                if (nd != kSlotFree && (nd & LRK_BVH_LEAF)) {
                    Slot sl = make_slot(s);
                    RayHot r;
                    RayCold c;
                    sl.load_hot(r);
                    sl.load_cold(c);
                    finished = leaf_step<ANY_HIT, COUNT, ALPHA>(sc, r, c, sl, sl, cnt);
                    if (finished) {
                        hit = make_uint4(c.best_inst, c.best_prim, __float_as_uint(c.best_u), __float_as_uint(c.best_v));
                        ray_index = c.ray_index;
                        nd = kSlotFree;
                    } else {
                        sl.store_hot(r);
                        sl.store_cold(c);
                        nd = r.node;
                    }
#pragma unroll
                    for (int q = 0; q < K; q++)
                        if (q == s) node[q] = nd;
                }
                sink(finished, ray_index, hit);
                live += __popc(__ballot_sync(0xffffffffu, nd != kSlotFree));
            }
            if (live == 0u || (live < sc.refill_below * K && !exhausted)) break;
        }
    }
}

#endif// __CUDACC__

}// namespace lrk
