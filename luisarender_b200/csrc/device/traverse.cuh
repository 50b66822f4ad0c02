// BVH2 traversal + ray/triangle intersection for sm_100a (B200 has no RT cores).
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
//   * Moeller-Trumbore in object space with fma dot/cross; accept tmin < t < t_best, u,v >= 0, u+v <= 1;
//   * instance entry transforms the ray by world_to_object without renormalising d (t is shared).
// Deferred children live in a per-thread stack kept in local memory (L1-resident, lane-interleaved);
// a TLAS->BLAS transition pushes an exit sentinel so the world-space ray is restored on return.
//
// Warp scheduling (this file, no oracle counterpart — it does not change any per-ray result): incoherent
// rays have very different traversal lengths (ncu on the 1.39M-triangle scene: 5.8 of 32 lanes active at
// bounce 1 with one-ray-per-thread scheduling), so rays are pulled from the queue through a per-launch
// atomic cursor and a warp REFILLS its idle lanes with fresh rays whenever fewer than REFILL_BELOW lanes
// are still traversing (persistent warps with dynamic ray replacement).
#pragma once
#include "scene.cuh"

namespace lrk {

constexpr uint32_t kSentinelDone = 0xfffffffdu;
constexpr uint32_t kSentinelExit = 0xfffffffeu;
constexpr int kStackSize = 96;
constexpr int kRefillBelow = 16;// refill when fewer than this many lanes of the warp hold a live ray (swept on B200: 13-19 is a plateau)
constexpr int kInnerMin = 8;   // leave the inner-node phase when fewer lanes than this still descend (swept: 6-10 is a plateau)

struct TraversalCounters {
    uint32_t nodes, tris, xforms;
};

__device__ __forceinline__ float safe_rcp(float d) {
    float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
    return 1.0f / a;
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

// Traces rays [0, n) of the queue (ray_o / ray_d). `cursor` is a zero-initialised device counter private to
// this launch. `sink(finished, ray_index, hit)` is called by ALL 32 lanes together (warp-convergent) after every
// traversal step; a lane passes finished = true exactly once per ray, with
// hit = {inst, prim, bary.u bits, bary.v bits} (miss <=> inst == ~0u; ANY_HIT: first hit found).
// Optional visiting order: rays binned by direction octant (bin_rays_kernel).  Position p of the launch's cursor is
// mapped to the p-th entry of the concatenated bins; with `order == nullptr` the queue is visited in storage order.
constexpr uint32_t kRayBins = 8u;
struct RayOrder {
    const uint32_t *order;     // [kRayBins][capacity] ray indices
    const uint32_t *bin_counts;// [kRayBins]
    uint32_t capacity;
};
__device__ __forceinline__ uint32_t ordered_index(const RayOrder &ro, uint32_t p) {
    if (ro.order == nullptr) return p;
    uint32_t start = 0u;
#pragma unroll
    for (uint32_t b = 0; b < kRayBins; b++) {
        const uint32_t c = __ldg(ro.bin_counts + b);
        if (p < start + c) return __ldg(ro.order + static_cast<size_t>(b) * ro.capacity + (p - start));
        start += c;
    }
    return p;// unreachable when the bins cover [0, n)
}

// ALPHA: stochastic alpha test of every accepted candidate (scenes with non-opaque surfaces only; alpha_skip is in shading.cuh)
__device__ bool alpha_skip(const DeviceScene &sc, uint32_t inst_id, uint32_t prim_id, float bu, float bv);
template<bool ANY_HIT, bool COUNT, int STRIDE, bool ALPHA = false, typename Sink>
__device__ __forceinline__ void trace_queue(const DeviceScene &sc, const float4 *__restrict__ ray_o, const float4 *__restrict__ ray_d,
                                            uint32_t n, uint32_t *cursor, TraversalCounters &cnt, Sink &&sink,
                                            RayOrder ray_order = RayOrder{nullptr, nullptr, 0u}) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lane_lt = (1u << lane) - 1u;
    uint32_t stack[kStackSize];
    int sp = 0;
    uint32_t node = kSentinelDone;
    bool active = false, exhausted = false, in_blas = false;
    uint32_t ray_index = 0u, cur_inst = ~0u, best_inst = ~0u, best_prim = ~0u;
    float best_u = 0.f, best_v = 0.f, tbest = 0.f, tmin = 0.f;
    V3 world_o = v3(0.f), world_d = v3(0.f, 0.f, 1.f);
    RaySetup cur;
    cur.set(world_o, world_d);
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
                const uint32_t idx = ordered_index(ray_order, pos);
                float4 o = ray_o[static_cast<size_t>(idx) * STRIDE], d = ray_d[static_cast<size_t>(idx) * STRIDE];
                world_o = v3(o.x, o.y, o.z);
                world_d = v3(d.x, d.y, d.z);
                tmin = o.w;
                tbest = d.w;
                cur.set(world_o, world_d);
                ray_index = idx;
                best_inst = ~0u;
                best_prim = ~0u;
                best_u = best_v = 0.f;
                sp = 0;
                stack[sp++] = kSentinelDone;
                node = sc.tlas_root;
                in_blas = false;
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
                const bool inner = !(node & LRK_BVH_LEAF);
                const uint32_t n_inner = __popc(__ballot_sync(0xffffffffu, inner));
                if (n_inner == 0u) break;
                if (n_inner < sc.inner_min && __any_sync(0xffffffffu, active && !inner)) break;
                if (!inner) continue;
                const float4 *np = sc.bvh_nodes + static_cast<size_t>(node) * 4u;
                float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
                if (COUNT) cnt.nodes++;
                float tn0, tn1;
                bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, cur, tmin, tbest, tn0);
                bool h1 = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, cur, tmin, tbest, tn1);
                uint32_t ref0 = __float_as_uint(n3.x), ref1 = __float_as_uint(n3.y);
                if (h0 && h1) {
                    bool first0 = tn0 <= tn1;
                    stack[sp++] = first0 ? ref1 : ref0;
                    node = first0 ? ref0 : ref1;
                } else if (h0) {
                    node = ref0;
                } else if (h1) {
                    node = ref1;
                } else {
                    node = stack[--sp];
                }
            }
            bool finished = false;
            if (active && (node & LRK_BVH_LEAF)) {
                if (node == kSentinelDone) {
                    finished = true;
                } else if (node == kSentinelExit) {
                    cur.set(world_o, world_d);
                    in_blas = false;
                    node = stack[--sp];
                } else if (node == LRK_BVH_EMPTY) {
                    node = stack[--sp];
                } else if (in_blas) {
                    uint32_t first = node & 0x0fffffffu;
                    uint32_t count = ((node >> 28u) & 7u) + 1u;
                    for (uint32_t k = 0; k < count; k++) {
                        const float4 *tv = sc.tri_verts + static_cast<size_t>(first + k) * 3u;
                        float4 a = __ldg(tv + 0), b = __ldg(tv + 1), c = __ldg(tv + 2);
                        if (COUNT) cnt.tris++;
                        V3 p0 = v3(a.x, a.y, a.z), p1 = v3(b.x, b.y, b.z), p2 = v3(c.x, c.y, c.z);
                        V3 e1 = p1 - p0, e2 = p2 - p0;
                        V3 pvec = fcross(cur.d, e2);
                        float det = fdot(e1, pvec);
                        if (!(det != 0.0f)) continue;
                        float inv_det = 1.0f / det;
                        V3 tvec = cur.o - p0;
                        float u = fdot(tvec, pvec) * inv_det;
                        if (!(u >= 0.0f && u <= 1.0f)) continue;
                        V3 qvec = fcross(tvec, e1);
                        float v = fdot(cur.d, qvec) * inv_det;
                        if (!(v >= 0.0f && u + v <= 1.0f)) continue;
                        float t = fdot(e2, qvec) * inv_det;
                        if (!(t > tmin && t < tbest)) continue;
                        if (ALPHA) {// on_surface_candidate: commit only if not skipped (geometry.cpp:248-279)
                            if (alpha_skip(sc, cur_inst, __float_as_uint(a.w), u, v)) continue;
                        }
                        tbest = t;
                        best_inst = cur_inst;
                        best_prim = __float_as_uint(a.w);
                        best_u = u;
                        best_v = v;
                        if (ANY_HIT) {
                            finished = true;
                            break;
                        }
                    }
                    node = stack[--sp];
                } else {
                    cur_inst = node & 0x7fffffffu;
                    if (COUNT) cnt.xforms++;
                    const float4 *x = sc.inst_xform + static_cast<size_t>(cur_inst) * 4u;
                    float4 r0 = __ldg(x + 0), r1 = __ldg(x + 1), r2 = __ldg(x + 2), r3 = __ldg(x + 3);
                    V3 oo = v3(fmaf(r0.x, world_o.x, fmaf(r0.y, world_o.y, fmaf(r0.z, world_o.z, r0.w))),
                               fmaf(r1.x, world_o.x, fmaf(r1.y, world_o.y, fmaf(r1.z, world_o.z, r1.w))),
                               fmaf(r2.x, world_o.x, fmaf(r2.y, world_o.y, fmaf(r2.z, world_o.z, r2.w))));
                    V3 dd = v3(fmaf(r0.x, world_d.x, fmaf(r0.y, world_d.y, r0.z * world_d.z)),
                               fmaf(r1.x, world_d.x, fmaf(r1.y, world_d.y, r1.z * world_d.z)),
                               fmaf(r2.x, world_d.x, fmaf(r2.y, world_d.y, r2.z * world_d.z)));
                    cur.set(oo, dd);
                    stack[sp++] = kSentinelExit;
                    in_blas = true;
                    node = __float_as_uint(r3.x);
                }
            }
            sink(finished, ray_index, make_uint4(best_inst, best_prim, __float_as_uint(best_u), __float_as_uint(best_v)));
            if (finished) {
                active = false;
                node = kSentinelDone;// idle lanes skip the inner loop
            }
            const uint32_t busy = __popc(__ballot_sync(0xffffffffu, active));
            if (busy == 0u || (busy < sc.refill_below && !exhausted)) break;
        }
    }
}

}// namespace lrk
