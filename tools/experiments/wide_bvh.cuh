// Wide (4-ary) BVH for the sm_100a traversal kernels, derived ON THE DEVICE from the host-built BVH2 of include/lrk.h.
//
// Why: B200 has no RT cores; the round-1 BVH2 kernel was instruction-issue bound at 13.7 of 32 lanes per instruction
// (profiles/r01_ncu_full_final.txt).  A 4-wide node does the work of ~2 BVH2 steps in one uniform straight-line block (one
// 128-byte line, 7 independent 16-byte loads, 24 FMAs), halves the number of divergent loop trips, stack operations and
// warp votes per ray, and lets deferred children carry their entry distance so that they are culled when popped.
//
// Parity: the reference delegates traversal to OptiX / Embree (src/base/geometry.cpp:218-279), the oracle walks the BVH2.
// A wide node's child boxes ARE boxes of BVH2 nodes (copied bit for bit, padding included) and the slab arithmetic is the
// oracle's (fmaf(plane, 1/d, -o/d)), so every triangle the oracle's descent reaches passes a SUBSET of the same box tests
// here; the triangle test is unchanged.  Closest hits are therefore identical (ties in t aside, which the tests never see).
//
// Layout of one wide node, 128 B = 8 x float4 (one L1/L2 line):
//   row 0  lo.x[4]   row 1  lo.y[4]   row 2  lo.z[4]   row 3  child refs[4] (uint)
//   row 4  hi.x[4]   row 5  hi.y[4]   row 6  hi.z[4]   row 7  unused
// A lane reads its NEAR planes at byte offset (d < 0 ? 64 : 0) + 16 * axis and its FAR planes at that offset ^ 64: no
// min / max per slab and no select.  Child refs are BVH2 refs unchanged: inner = node index (wide node i is the collapse
// of BVH2 node i, so no index translation exists), leaf = bit 31 | ..., LRK_BVH_EMPTY for an unused slot (inverted box).
#pragma once
#include "scene.cuh"

namespace lrk {

constexpr uint32_t kWideRows = 8u;// float4 rows per wide node

__device__ __forceinline__ float box_half_area(const float lo[3], const float hi[3]) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// Collapse of BVH2 node `i`: start from its two children and, while fewer than four slots are used, replace the inner child
// with the largest surface area by its own two children (greedy surface-area collapse).  Every BVH2 node gets a wide node at
// the same index, so the kernel below is one independent thread per node: no top-down pass, no compaction, no atomics.
__device__ inline void collapse_wide_node(const float4 *__restrict__ bvh2, uint32_t i, float4 *__restrict__ out) {
    float lo[4][3], hi[4][3];
    uint32_t ref[4];
    int cnt = 0;
    auto add_children = [&](uint32_t node, int replace) {
        const float4 *np = bvh2 + static_cast<size_t>(node) * 4u;
        float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
        const uint32_t r0 = __float_as_uint(n3.x), r1 = __float_as_uint(n3.y);
        int slot = replace;
        if (r0 != LRK_BVH_EMPTY) {
            if (slot < 0) slot = cnt++;
            lo[slot][0] = n0.x; lo[slot][1] = n0.y; lo[slot][2] = n0.z;
            hi[slot][0] = n0.w; hi[slot][1] = n1.x; hi[slot][2] = n1.y;
            ref[slot] = r0;
            slot = -1;
        }
        if (r1 != LRK_BVH_EMPTY) {
            if (slot < 0) slot = cnt++;
            lo[slot][0] = n1.z; lo[slot][1] = n1.w; lo[slot][2] = n2.x;
            hi[slot][0] = n2.y; hi[slot][1] = n2.z; hi[slot][2] = n2.w;
            ref[slot] = r1;
            slot = -1;
        }
        if (slot >= 0) {// the expanded node had no children at all (cannot happen for builder output): drop the slot
            cnt--;
            for (int a = 0; a < 3; a++) { lo[slot][a] = lo[cnt][a]; hi[slot][a] = hi[cnt][a]; }
            ref[slot] = ref[cnt];
        }
    };
    add_children(i, -1);
    while (cnt < 4) {
        int best = -1;
        float best_area = -1.f;
        for (int k = 0; k < cnt; k++) {
            if (ref[k] & LRK_BVH_LEAF) continue;
            float a = box_half_area(lo[k], hi[k]);
            if (a > best_area) { best_area = a; best = k; }
        }
        if (best < 0) break;
        add_children(ref[best], best);
    }
    const float inf = __uint_as_float(0x7f800000u);
    for (int k = cnt; k < 4; k++) {
        for (int a = 0; a < 3; a++) { lo[k][a] = inf; hi[k][a] = -inf; }
        ref[k] = LRK_BVH_EMPTY;
    }
    out[0] = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]);
    out[1] = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]);
    out[2] = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]);
    out[3] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), __uint_as_float(ref[2]), __uint_as_float(ref[3]));
    out[4] = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]);
    out[5] = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
    out[6] = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
    out[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(256) collapse_wide_kernel(const float4 *__restrict__ bvh2, float4 *__restrict__ wide, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 rows[kWideRows];
    collapse_wide_node(bvh2, i, rows);
    float4 *dst = wide + static_cast<size_t>(i) * kWideRows;
#pragma unroll
    for (uint32_t r = 0; r < kWideRows; r++) dst[r] = rows[r];
}
#endif

}// namespace lrk
