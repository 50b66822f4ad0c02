// The path state of the wavefront kernels: queue layout and launch constants shared by the two translation units of the device
// library - lrk.cu (ray generation, traversal, classification, film; IEEE arithmetic, bit-exact with the oracle) and shade.cu
// (the closure kernels; compiled with the arithmetic the reference's own CUDA backend uses, see shade.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "scene.cuh"

namespace lrk {


constexpr int kBlock = 256;
#ifndef LRK_TRACE_MIN_BLOCKS
#define LRK_TRACE_MIN_BLOCKS 4// blocks of kTraceBlock threads per SM the traversal kernels are compiled for (register budget)
#endif
#ifndef LRK_SHADOW_MIN_BLOCKS
// the any-hit kernel is the one that gains from a fifth resident block (51 registers, 22 bytes spilled): 16.54 -> 16.04 ms per pass;
// the closest-hit kernel loses at 5 (23.36 -> 24.13) and both lose at 6 (profiles/r02x_traversal_occupancy_and_cache_policy.jsonl)
#define LRK_SHADOW_MIN_BLOCKS 5
#endif
#ifndef LRK_SHADE_BLOCK
#define LRK_SHADE_BLOCK 256
#endif
#ifndef LRK_SHADE_MIN_BLOCKS
#define LRK_SHADE_MIN_BLOCKS 2
#endif
constexpr int kShadeBlock = LRK_SHADE_BLOCK;// threads per block of the surface shade kernels (register-bound: see DESIGN.md)
constexpr uint32_t kCountSlots = 16u;   // rows of 64 counters in PathBuffers::counts: 4 queue / cursor rows + one per hit bucket
constexpr uint32_t kMaxDepthSlots = 64u;// counts[0..63]: path queue size per depth, counts[64..127]: shadow queue size

struct PathBuffers {
    float4 *ray_o[2];
    float4 *ray_d[2];
    float4 *beta_pdf[2];
    uint2 *id_rng[2];
    uint4 *hit;// {inst, prim, bary} per ray of the current queue (inst == ~0u: escaped)
    uint32_t *hit_index[11];// per closure kind: indices (into the current ray queue) of the rays that hit such a surface
    float4 *sray_o;
    float4 *sray_d;
    float4 *scontrib;// rgb + path id bits
    float4 *li;
    uint32_t *counts;
    uint32_t capacity;
    // the pass being rendered: generation slot id -> (pixel, sample index) for the table-driven samplers (samplers.cuh)
    const uint32_t *pass_pixel_list;
    uint32_t pass_pixel_offset, pass_npix, pass_spp_begin;
    // volume path integrator only (config C4)
    ulonglong2 *pcg[2]; // per-path PCG32 {state, inc}
    float *u_rr[2];     // Russian-roulette number of the coming bounce (drawn at the top of the loop, mega_vpt_naive.cpp:256-257)
    float4 *s1ray_o;    // in-medium direct-light shadow ray of the coming bounce (from the ray origin)
    float4 *s1ray_d;
    uint32_t *occl1;    // ... and whether it hit a surface (advances the PCG32 stream by three draws)
    uint32_t *occl2[2]; // same for the surface NEE shadow ray of the previous bounce
    uint32_t *s2_target;// queue slot (next bounce) that receives occl2 for each shadow record, ~0u if the path ended
    // cost probe of lrk_balance_shards: live rays per pixel tile, counted by classify_hits_kernel (nullptr outside the probe)
    uint32_t *tile_cost;
    uint32_t tile_cost_size, tile_cost_tiles_x;
    unsigned long long *stats;// [0] closest rays, [1] shadow rays, [2..4] closest nodes/tris/xforms, [5..7] shadow nodes/tris/xforms
};
constexpr uint32_t kHitKinds = 11u;// hit buckets: emitter-only, Matte, Disney, Mirror, Glass, Plastic, Metal, Mix, transmissive Disney, Layered, thin Disney
static_assert(4u + kHitKinds <= kCountSlots, "PathBuffers::counts has one row of counters per hit bucket behind the four queue rows");

}// namespace lrk
