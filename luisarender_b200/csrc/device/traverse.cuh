// BVH2 traversal + ray/triangle intersection for sm_100a (B200 has no RT cores).
//
// Replaces Geometry::trace_closest / trace_any -> Accel::intersect / intersect_any
// (reference src/base/geometry.cpp:218-279), which the reference delegates to OptiX
// (src/compute/src/backends/cuda/cuda_builtin/cuda_device_resource.h:1603-1693) or Embree
// (src/compute/src/rust/luisa_compute_backend_impl/src/cpu/accel.rs:449-535).
//
// Rules (shared verbatim with the CPU oracle, oracle/oracle.cpp "BVH traversal"):
//   * two-level BVH2 over 64-byte nodes (both child boxes in one node), ordered traversal: when both
//     children are hit the nearer entry is visited first (ties: child 0), the other is deferred;
//   * slab test t = fma(plane, 1/d, -o/d) with |d| clamped to >= 1e-30; a child is hit when
//     max(t_entry, tmin) <= min(t_exit, t_best);
//   * Moeller-Trumbore in object space with fma dot/cross; accept tmin < t < t_best, u,v >= 0, u+v <= 1;
//   * instance entry transforms the ray by world_to_object without renormalising d (t is shared).
// Deferred children live in a per-thread stack kept in local memory (L1-resident, lane-interleaved);
// a TLAS->BLAS transition pushes an exit sentinel so the world-space ray is restored on return.
#pragma once
#include "scene.cuh"

namespace lrk {

constexpr uint32_t kSentinelDone = 0xfffffffdu;
constexpr uint32_t kSentinelExit = 0xfffffffeu;
constexpr int kStackSize = 96;

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

// Returns the closest hit (ANY_HIT = false) or the first hit found (ANY_HIT = true); miss <=> inst == ~0u.
template<bool ANY_HIT, bool COUNT>
__device__ __forceinline__ uint4 trace_ray(const DeviceScene &sc, float4 ray_o_tmin, float4 ray_d_tmax, TraversalCounters &cnt) {
    uint32_t best_inst = ~0u, best_prim = ~0u;
    float best_u = 0.f, best_v = 0.f;
    float tbest = ray_d_tmax.w;
    const float tmin = ray_o_tmin.w;
    const V3 world_o = v3(ray_o_tmin.x, ray_o_tmin.y, ray_o_tmin.z);
    const V3 world_d = v3(ray_d_tmax.x, ray_d_tmax.y, ray_d_tmax.z);
    RaySetup cur;
    cur.set(world_o, world_d);
    uint32_t stack[kStackSize];
    int sp = 0;
    stack[sp++] = kSentinelDone;
    uint32_t node = sc.tlas_root;
    bool in_blas = false;
    uint32_t cur_inst = ~0u;
    for (;;) {
        while (!(node & LRK_BVH_LEAF)) {
            const float4 *n = sc.bvh_nodes + static_cast<size_t>(node) * 4u;
            float4 n0 = __ldg(n + 0), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3);
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
        if (node == kSentinelDone) break;
        if (node == kSentinelExit) {
            cur.set(world_o, world_d);
            in_blas = false;
            node = stack[--sp];
            continue;
        }
        if (node == LRK_BVH_EMPTY) {
            node = stack[--sp];
            continue;
        }
        if (in_blas) {
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
                tbest = t;
                best_inst = cur_inst;
                best_prim = __float_as_uint(a.w);
                best_u = u;
                best_v = v;
                if (ANY_HIT) return make_uint4(best_inst, best_prim, __float_as_uint(best_u), __float_as_uint(best_v));
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
    return make_uint4(best_inst, best_prim, __float_as_uint(best_u), __float_as_uint(best_v));
}

}// namespace lrk
