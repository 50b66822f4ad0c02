// Two-level BVH2 traversal + ray/triangle intersection for sm_100a (B200 has no RT cores).
//
// Replaces Geometry::trace_closest / trace_any -> Accel::intersect / intersect_any
// (reference src/base/geometry.cpp:218-279), which the reference delegates to OptiX
// (src/compute/src/backends/cuda/cuda_builtin/cuda_device_resource.h:1603-1693) or Embree
// (src/compute/src/rust/luisa_compute_backend_impl/src/cpu/accel.rs:449-535).
//
// Per-ray rules (shared verbatim with the CPU oracle, oracle/oracle.cpp "BVH traversal"):
//   * two-level BVH2 over 64-byte nodes (both child boxes in one node), ordered traversal: when both
//     children are hit the nearer entry is visited first (ties: child 0), the other is deferred;
//   * slab test t = fma(plane, 1/d, -o/d) with |d| clamped to >= 1e-30; a child is hit when
//     max(t_entry, tmin) <= min(t_exit, t_best);
//   * Moeller-Trumbore in object space with fma dot/cross; accept tmin < t < t_best, u,v >= 0, u+v <= 1; an exact tie
//     t == t_best goes to the lower (instance, primitive), so that the result does not depend on the visiting order;
//   * instance entry transforms the ray by world_to_object without renormalising d (t is shared); a TLAS->BLAS transition
//     pushes an exit sentinel, and leaving the instance brings the world-space ray (o, d, 1/d, o/d) back from shared memory.
// Deferred children live in a per-thread stack: the first kSmemStack entries in shared memory (lane-interleaved, conflict-free),
// deeper ones in local memory; a stack that would outgrow both raises an error flag instead of writing out of bounds.
//
// Warp scheduling (no oracle counterpart — it does not change any per-ray result): incoherent rays have very different
// traversal lengths (ncu on the 1.39M-triangle scene: 5.8 of 32 lanes active at bounce 1 with one-ray-per-thread scheduling),
// so rays are pulled from the queue through a per-launch atomic cursor and a warp REFILLS its idle lanes with fresh rays
// whenever fewer than `refill_below` lanes are still traversing (persistent warps with dynamic ray replacement); lanes that
// stand on a leaf wait while at least `inner_min` lanes still descend (descent / leaf phases).
//
// Round 2 measured three alternatives to this kernel on the B200 and kept none of them (profiles/r02[a-d]_*): a 4-wide
// hierarchy collapsed on the device (same work per box, heavier steps: 27.7 / 18.5 ms against 24.0 / 16.8 ms per 64-spp pass),
// a TLAS-first walk with an instance candidate queue (29.7 / 20.0 ms), and two ray slots per lane with the ray state in shared
// memory (37.8 / 24.8 ms: active lanes rose from 11 to 14 of 32, occupancy and latency hiding fell more).
#pragma once
#include "scene.cuh"

namespace lrk {

constexpr uint32_t kSentinelDone = 0xfffffffdu;
constexpr uint32_t kSentinelExit = 0xfffffffeu;
#ifndef LRK_SMEM_STACK
#define LRK_SMEM_STACK 16
#endif
#ifndef LRK_INNER_UNROLL
#define LRK_INNER_UNROLL 2// inner-node steps per warp vote of the descent phase (trace_queue)
#endif
constexpr int kSmemStack = LRK_SMEM_STACK;// stack entries per thread kept in shared memory (0: all in local memory)
constexpr int kLocalStack = 160;// further entries in local memory: the host builder caps a hierarchy's depth at 48 + log2(n)
constexpr int kRefillBelow = 16;// refill when fewer than this many lanes of the warp hold a live ray (swept on B200: 13-19 is a plateau)
constexpr int kInnerMin = 8;   // leave the inner-node phase when fewer lanes than this still descend (swept: 6-10 is a plateau)

struct TraversalCounters {
    uint32_t nodes, tris, xforms;
};

// 1 / d for the slab test, |d| clamped to >= 1e-30.  On the device this is the 1-ulp MUFU reciprocal (the IEEE division with its
// Newton step and slow-path guards was 6.4 % of the kernel's issued instructions at 8 of 32 lanes: three of them per instance
// entry): it feeds the BOX test only - boxes are padded by 5e-7 of the hierarchy's extent against exactly this kind of rounding
// (bvh.cpp) - never the triangle test, so hit records stay bit-identical to the oracle's (tests/test_gpu_parity.py: 200 k rays
// per scene, plus the full-size scene).  LRK_IEEE_RCP restores the division.
__device__ __forceinline__ float safe_rcp(float d) {
    float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
#if defined(__CUDA_ARCH__) && !defined(LRK_IEEE_RCP)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
#else
    return 1.0f / a;
#endif
}

struct RaySetup {
    V3 o, d, inv, ood;
    __device__ __forceinline__ void set(V3 oo, V3 dd) {
        o = oo;
        d = dd;
        inv = v3(safe_rcp(dd.x), safe_rcp(dd.y), safe_rcp(dd.z));
        ood = v3(oo.x * inv.x, oo.y * inv.y, oo.z * inv.z);
    }
};

// Per-ray traversal state (one per lane on the device).
struct RayState {
    RaySetup cur;     // the ray in the space of the hierarchy being walked (world space in the TLAS, object space inside an instance)
    float tmin, tbest;
    uint32_t node;    // the reference this ray stands on: inner node index, leaf ref, or a sentinel
    uint32_t cur_inst;// instance being walked, ~0u while in the TLAS
    uint32_t best_inst, best_prim;
    float best_u, best_v;
};

__device__ __forceinline__ bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz, const RaySetup &r,
                                     float tmin, float tbest, float &tnear) {
    float t0x = fmaf(lox, r.inv.x, -r.ood.x), t1x = fmaf(hix, r.inv.x, -r.ood.x);
    float t0y = fmaf(loy, r.inv.y, -r.ood.y), t1y = fmaf(hiy, r.inv.y, -r.ood.y);
    float t0z = fmaf(loz, r.inv.z, -r.ood.z), t1z = fmaf(hiz, r.inv.z, -r.ood.z);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tbest));
    tnear = tn;
    return tn <= tf;
}

// One step on an inner node: test both child boxes, continue with the nearer hit child, defer the other, pop when nothing was hit.
template<bool COUNT, typename Stack>
__device__ __forceinline__ void inner_step(const DeviceScene &sc, RayState &r, Stack &stack, TraversalCounters &cnt) {
    const float4 *np = sc.bvh_nodes + static_cast<size_t>(r.node) * 4u;
#if defined(LRK_NODE_EVICT_LAST) && defined(__CUDA_ARCH__)
    float4 n0, n1, n2, n3;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w) : "l"(np + 0));
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(np + 1));
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w) : "l"(np + 2));
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w) : "l"(np + 3));
#else
    float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
#endif
    if (COUNT) cnt.nodes++;
    float tn0, tn1;
    bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, r.cur, r.tmin, r.tbest, tn0);
    bool h1 = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, r.cur, r.tmin, r.tbest, tn1);
    uint32_t ref0 = __float_as_uint(n3.x), ref1 = __float_as_uint(n3.y);
    // straight-line code: the deferred child is stored under a predicate, the top of the stack is read whether or not it is
    // needed (a divergent pop branch costs a full warp instruction per instruction for the two or three lanes that take it)
    const bool both = h0 && h1, none = !h0 && !h1;
    const bool first0 = tn0 <= tn1;
    stack.push_if(both, first0 ? ref1 : ref0);
    const uint32_t top = stack.top();
    r.node = both ? (first0 ? ref0 : ref1) : h0 ? ref0 : h1 ? ref1 : top;
    stack.drop_if(none);
}

// ALPHA: stochastic alpha test of every accepted candidate (scenes with non-opaque surfaces only; alpha_skip is in shading.cuh)
__device__ inline bool alpha_skip(const DeviceScene &sc, uint32_t inst_id, uint32_t prim_id, float bu, float bv);

// One step on a leaf-like reference (bit 31 set): sentinel, instance (TLAS leaf) or triangle range (BLAS leaf).
// `world` parks the world-space ray while the lane is inside an instance.  Returns true when the ray is finished.
template<bool ANY_HIT, bool COUNT, bool ALPHA, typename Stack, typename World>
__device__ __forceinline__ bool leaf_step(const DeviceScene &sc, RayState &r, Stack &stack, World &world, TraversalCounters &cnt) {
    const uint32_t node = r.node;
    if (node == kSentinelDone) return true;
    if (node == kSentinelExit) {
        world.load(r.cur);
        r.cur_inst = ~0u;
        r.node = stack.pop();
        return false;
    }
    if (node == LRK_BVH_EMPTY) {
        r.node = stack.pop();
        return false;
    }
    if (r.cur_inst != ~0u) {
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
            // on the visiting order (the oracle's rule, and the one of oracle/ref's backend for the reference itself)
            const bool tie = t == r.tbest && r.best_inst != ~0u && (r.cur_inst < r.best_inst || (r.cur_inst == r.best_inst && prim < r.best_prim));
            if (!(t > r.tmin && (t < r.tbest || tie))) continue;
            if (ALPHA) {// on_surface_candidate: commit only if not skipped (geometry.cpp:248-279)
                if (alpha_skip(*sc.self, r.cur_inst, prim, u, v)) continue;
            }
            r.tbest = t;
            r.best_inst = r.cur_inst;
            r.best_prim = prim;
            r.best_u = u;
            r.best_v = v;
            if (ANY_HIT) return true;
        }
        r.node = stack.pop();
        return false;
    }
    // TLAS leaf: enter the instance
    r.cur_inst = node & 0x7fffffffu;
    if (COUNT) cnt.xforms++;
    const float4 *x = sc.inst_xform + static_cast<size_t>(r.cur_inst) * 4u;
    float4 m0 = __ldg(x + 0), m1 = __ldg(x + 1), m2 = __ldg(x + 2), m3 = __ldg(x + 3);
    world.save(r.cur);
    const V3 wo = r.cur.o, wd = r.cur.d;
    V3 oo = v3(fmaf(m0.x, wo.x, fmaf(m0.y, wo.y, fmaf(m0.z, wo.z, m0.w))),
               fmaf(m1.x, wo.x, fmaf(m1.y, wo.y, fmaf(m1.z, wo.z, m1.w))),
               fmaf(m2.x, wo.x, fmaf(m2.y, wo.y, fmaf(m2.z, wo.z, m2.w))));
    V3 dd = v3(fmaf(m0.x, wd.x, fmaf(m0.y, wd.y, m0.z * wd.z)),
               fmaf(m1.x, wd.x, fmaf(m1.y, wd.y, m1.z * wd.z)),
               fmaf(m2.x, wd.x, fmaf(m2.y, wd.y, m2.z * wd.z)));
    r.cur.set(oo, dd);
    stack.push(kSentinelExit);
    r.node = __float_as_uint(m3.x);
    return false;
}

template<typename Stack>
__device__ __forceinline__ void start_ray(const DeviceScene &sc, RayState &r, Stack &stack, float4 o, float4 d) {
    r.cur.set(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z));
    r.tmin = o.w;
    r.tbest = d.w;
    r.best_inst = ~0u;
    r.best_prim = ~0u;
    r.best_u = r.best_v = 0.f;
    r.cur_inst = ~0u;
    stack.reset();
    stack.push(kSentinelDone);
    r.node = sc.tlas_root;
}

#ifdef __CUDACC__

#ifndef LRK_TRACE_BLOCK
#define LRK_TRACE_BLOCK 256
#endif
constexpr int kTraceBlock = LRK_TRACE_BLOCK;// threads per block of every kernel that calls trace_queue

// Cache policy (profiles/r02x_traversal_occupancy_and_cache_policy.jsonl):
//   ray records are read once and hit records written once per launch: ld.global.cs / st.global.cs (evict first), so that they do
//   not push hierarchy nodes out of L1 / L2 (closest 23.36 -> 23.16 ms per pass; -DLRK_NO_STREAM_RAYS restores plain loads);
//   LRK_NODE_EVICT_LAST (experiment, off): ld.global.nc.L1::evict_last for the nodes on top of that changed nothing.
#ifndef LRK_NO_STREAM_RAYS
#define LRK_STREAM_RAYS 1
#endif
__device__ __forceinline__ float4 load_ray_record(const float4 *p) {
#ifdef LRK_STREAM_RAYS
    return __ldcs(p);
#else
    return __ldg(p);
#endif
}
template<typename T>
__device__ __forceinline__ void store_result_record(T *p, T v) {
#ifdef LRK_STREAM_RAYS
    __stcs(p, v);
#else
    *p = v;
#endif
}

// Deferred-children stack of a lane: kSmemStack entries in shared memory (explicit st/ld.shared through the 32-bit window address:
// the struct holds scalars only and lives in registers), then local memory, then the overflow flag.
struct LaneStack {
    uint32_t smem;// shared-window address of this thread's column: entry k at smem + k * kTraceBlock * 4
    uint32_t *local;
    int sp;
    uint32_t *overflow;
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(uint32_t ref) {
        if (sp < kSmemStack) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem + static_cast<uint32_t>(sp) * (kTraceBlock * 4u)), "r"(ref) : "memory");
        } else if (sp < kSmemStack + kLocalStack) {
            local[sp - kSmemStack] = ref;
        } else {
            *overflow = 1u;// reported by lrk_render / lrk_trace as an error (hierarchy too deep); the entry is lost
        }
        sp++;
    }
    __device__ __forceinline__ uint32_t at(int i) const {
        if (i >= kSmemStack) return local[min(i, kSmemStack + kLocalStack - 1) - kSmemStack];
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem + static_cast<uint32_t>(i) * (kTraceBlock * 4u)) : "memory");
        return v;
    }
    __device__ __forceinline__ uint32_t pop() {
        --sp;
        return at(sp);
    }
    __device__ __forceinline__ void push_if(bool p, uint32_t ref) {
        if (p) push(ref);
    }
    __device__ __forceinline__ uint32_t top() const { return at(sp - 1); }// the bottom sentinel is always there
    __device__ __forceinline__ void drop_if(bool p) { sp -= p ? 1 : 0; }
};

// The world-space ray of a lane that is inside an instance: three float4 per thread in shared memory, so that leaving the
// instance costs three 16-byte shared loads instead of three IEEE divisions.
struct LaneWorld {
    uint32_t smem;// shared-window address of this thread's column: row k at smem + k * kTraceBlock * 16
#ifdef LRK_WORLD_RECOMPUTE
    V3 o, d;
    __device__ __forceinline__ void save(const RaySetup &c) { o = c.o; d = c.d; }
    __device__ __forceinline__ void load(RaySetup &c) { c.set(o, d); }
#else
    __device__ __forceinline__ void save(const RaySetup &c) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem), "f"(c.o.x), "f"(c.o.y), "f"(c.o.z), "f"(c.d.x) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem + kTraceBlock * 16u), "f"(c.d.y), "f"(c.d.z), "f"(c.inv.x), "f"(c.inv.y) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem + kTraceBlock * 32u), "f"(c.inv.z), "f"(c.ood.x), "f"(c.ood.y), "f"(c.ood.z) : "memory");
    }
    __device__ __forceinline__ void load(RaySetup &c) {
        float4 a, b, d;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(smem) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(smem + kTraceBlock * 16u) : "memory");
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "r"(smem + kTraceBlock * 32u) : "memory");
        c.o = v3(a.x, a.y, a.z);
        c.d = v3(a.w, b.x, b.y);
        c.inv = v3(b.z, b.w, d.x);
        c.ood = v3(d.y, d.z, d.w);
    }
#endif
};

struct TraceShared {
    uint32_t stack[kSmemStack > 0 ? kSmemStack : 1][kTraceBlock];
    float4 world[3][kTraceBlock];
};

// Traces rays [0, n) of the queue (ray_o / ray_d). `cursor` is a zero-initialised device counter private to
// this launch. `sink(finished, ray_index, hit)` is called by ALL 32 lanes together (warp-convergent) after every
// traversal step; a lane passes finished = true exactly once per ray, with
// hit = {inst, prim, bary.u bits, bary.v bits} (miss <=> inst == ~0u; ANY_HIT: first hit found).
template<bool ANY_HIT, bool COUNT, int STRIDE, bool ALPHA = false, typename Sink>
__device__ __forceinline__ void trace_queue(const DeviceScene &sc, const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d,
                                            uint32_t n, uint32_t *cursor, TraversalCounters &cnt, Sink &&sink) {
    __shared__ TraceShared shared;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t deep_stack[kLocalStack];
    LaneStack stack;
    stack.smem = static_cast<uint32_t>(__cvta_generic_to_shared(&shared.stack[0][threadIdx.x]));
    stack.local = deep_stack;
    stack.overflow = sc.traversal_overflow;
    stack.reset();
    LaneWorld world;
    world.smem = static_cast<uint32_t>(__cvta_generic_to_shared(&shared.world[0][threadIdx.x]));
    RayState r;
    start_ray(sc, r, stack, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 1.f, 0.f));
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
                float4 o = load_ray_record(ray_o + static_cast<size_t>(pos) * STRIDE), d = load_ray_record(ray_d + static_cast<size_t>(pos) * STRIDE);
                start_ray(sc, r, stack, o, d);
                ray_index = pos;
                active = true;
            }
            exhausted = base + want >= n;
        }
        if (!__any_sync(0xffffffffu, active)) break;
        // ---- traverse until too few lanes are busy (and fresh rays are available) ---------------------------
        for (;;) {
            // inner phase: step the lanes that stand on an inner node; lanes that have reached a leaf wait, but only while at
            // least `inner_min` lanes still have inner work (inner_min = 1 is the classic while-while loop, 32 is if-if)
            for (;;) {
                const bool inner = !(r.node & LRK_BVH_LEAF);
                const uint32_t n_inner = __popc(__ballot_sync(0xffffffffu, inner));
                if (n_inner == 0u) break;
                if (n_inner < sc.inner_min && __any_sync(0xffffffffu, active && !inner)) break;
                if (inner) {
                    inner_step<COUNT>(sc, r, stack, cnt);
#pragma unroll
                    for (int extra = 1; extra < LRK_INNER_UNROLL; extra++)// further steps without a new warp vote
                        if (!(r.node & LRK_BVH_LEAF)) inner_step<COUNT>(sc, r, stack, cnt);
                }
            }
            bool finished = false;
            if (active && (r.node & LRK_BVH_LEAF)) finished = leaf_step<ANY_HIT, COUNT, ALPHA>(sc, r, stack, world, cnt);
            sink(finished, ray_index, make_uint4(r.best_inst, r.best_prim, __float_as_uint(r.best_u), __float_as_uint(r.best_v)));
            if (finished) {
                active = false;
                r.node = kSentinelDone;// idle lanes skip the inner loop
            }
            const uint32_t busy = __popc(__ballot_sync(0xffffffffu, active));
            if (busy == 0u || (busy < sc.refill_below && !exhausted)) break;
        }
    }
}

#endif// __CUDACC__

// ---- one ray per thread ---------------------------------------------------------------------------------------------------
// The same step functions driven by a plain per-thread loop with a local-memory stack: for code that traces from inside a
// longer per-thread computation (the general volume integrator walks its transmittance rays surface by surface).  Results are
// those of trace_queue bit for bit; only the scheduling differs.
struct ThreadStack {
    uint32_t *e;
    int sp;
    uint32_t *overflow;
    static constexpr int kCapacity = kSmemStack + kLocalStack;
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(uint32_t ref) {
        if (sp < kCapacity) e[sp] = ref;
        else *overflow = 1u;
        sp++;
    }
    __device__ __forceinline__ uint32_t at(int i) const { return e[min(i, kCapacity - 1)]; }
    __device__ __forceinline__ uint32_t pop() {
        --sp;
        return at(sp);
    }
    __device__ __forceinline__ void push_if(bool p, uint32_t ref) {
        if (p) push(ref);
    }
    __device__ __forceinline__ uint32_t top() const { return at(sp - 1); }
    __device__ __forceinline__ void drop_if(bool p) { sp -= p ? 1 : 0; }
};

struct ThreadWorld {
    RaySetup saved;
    __device__ __forceinline__ void save(const RaySetup &c) { saved = c; }
    __device__ __forceinline__ void load(RaySetup &c) const { c = saved; }
};

// hit = {inst, prim, bary.u bits, bary.v bits}; miss <=> inst == ~0u
template<bool ANY_HIT, bool ALPHA>
__device__ __noinline__ inline uint4 trace_single(const DeviceScene &sc, float4 o, float4 d) {
    uint32_t entries[ThreadStack::kCapacity];
    ThreadStack stack;
    stack.e = entries;
    stack.overflow = sc.traversal_overflow;
    ThreadWorld world;
    TraversalCounters cnt{0u, 0u, 0u};
    RayState r;
    start_ray(sc, r, stack, o, d);
    for (;;) {
        if (!(r.node & LRK_BVH_LEAF)) inner_step<false>(sc, r, stack, cnt);
        else if (leaf_step<ANY_HIT, false, ALPHA>(sc, r, stack, world, cnt)) break;
    }
    return make_uint4(r.best_inst, r.best_prim, __float_as_uint(r.best_u), __float_as_uint(r.best_v));
}

}// namespace lrk
