// The wavefront kernels (sm_100a).  One pass renders S samples of a chunk of pixels:
//
//   generate_rays                                   (reference wave_path.cpp:254-272)
//   for depth in 0 .. max_depth-1:
//       trace_closest   ray queue  -> hits          (wave_path.cpp:275-302, geometry.cpp:218-223)
//       shade           hits -> emitter MIS, light sample + shadow-ray record, closure evaluate/sample,
//                       Russian roulette, compacted next ray queue
//                                                   (wave_path.cpp:326-459 = evaluate_lights + sample_lights
//                                                    + evaluate_surfaces fused; hit reconstruction once)
//       trace_shadow    shadow queue -> Li[path] += contribution if unoccluded   (geometry.cpp:262-266)
//   accumulate          Li -> film, per-sample clamp (wave_path.cpp:462-469, films/color.cpp:107-130)
//
// Queues are dense SoA arrays of float4/uint2 records, ping-ponged per bounce; queue sizes live in device
// memory and every kernel is a grid-stride loop over `*count`, so the host never synchronises inside a
// pass (the reference's v2 reads six counters back per step, wave_path_v2.cpp:396-405).  Compaction uses
// warp ballots + one atomic per block iteration.  Radiance is carried in Li[path_id] (path_id is the
// generation slot), so film accumulation needs no atomics and is run-to-run deterministic.
#pragma once
#include <type_traits>
#include <utility>
#include "samplers.cuh"
#include "shading.cuh"
#include "traverse.cuh"
#include "pathcommon.cuh"

namespace lrk {

constexpr int kBlock = 256;
#ifndef LRK_TRACE_MIN_BLOCKS
#define LRK_TRACE_MIN_BLOCKS 4// blocks of kTraceBlock threads per SM the traversal kernels are compiled for (register budget)
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
    uint32_t *hit_index[9];// per closure kind: indices (into the current ray queue) of the rays that hit such a surface
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
    unsigned long long *stats;// [0] closest rays, [1] shadow rays, [2..4] closest nodes/tris/xforms, [5..7] shadow nodes/tris/xforms
};

__global__ void __launch_bounds__(kBlock) generate_rays_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ pixel_list,
                                                               uint32_t pixel_offset, uint32_t npix, uint32_t spp_begin, uint32_t n) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id == 0u) {
        pb.counts[0] = n;
        for (uint32_t d = 1u; d < kCountSlots * kMaxDepthSlots; d++) pb.counts[d] = 0u;// sizes + the traversal kernels' fetch cursors
    }
    if (id >= n) return;
    uint32_t k = id % npix;
    uint32_t s = id / npix;
    uint32_t pixel = __ldg(pixel_list + pixel_offset + k);
    uint32_t px = pixel & 0xffffu, py = pixel >> 16u;
    PathSampler smp;
    smp.start(sampler_ref(sc), px, py, spp_begin + s);
    const float2 uf = smp.pixel2d(sampler_ref(sc));
    float4 ro, rd;
    float weight;
    camera_ray(sc.camera, px, py, uf.x, uf.y, ro, rd, weight);
    pb.ray_o[0][id] = ro;
    pb.ray_d[0][id] = rd;
    pb.beta_pdf[0][id] = make_float4(weight, weight, weight, 1e16f);
    pb.id_rng[0][id] = make_uint2(id, smp.state);
    pb.li[id] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- traversal kernels -------------------------------------------------------------------------------
// counts[] layout (all zeroed by generate_rays_kernel at the start of a pass), 64 slots (one per depth) each:
//   [0] path-queue sizes  [1] shadow-queue sizes  [2] closest-hit fetch cursors  [3] shadow fetch cursors
//   [4],[5],[6] hit-bucket sizes: light-only hits, Matte hits, Disney hits
template<bool COUNT, bool ALPHA = false>
__global__ void __launch_bounds__(kTraceBlock, LRK_TRACE_MIN_BLOCKS) trace_closest_kernel(DeviceScene sc, const float4 *__restrict__ ray_o,
                                                               const float4 *__restrict__ ray_d, uint4 *__restrict__ hits,
                                                               const uint32_t *__restrict__ count, uint32_t *cursor,
                                                               unsigned long long *stats) {
    const uint32_t n = *count;
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<false, COUNT, 1, ALPHA>(sc, ray_o, ray_d, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
        if (finished) hits[i] = h;
    });
    if (COUNT) {
        atomicAdd(stats + 2, static_cast<unsigned long long>(tc.nodes));
        atomicAdd(stats + 3, static_cast<unsigned long long>(tc.tris));
        atomicAdd(stats + 4, static_cast<unsigned long long>(tc.xforms));
    }
}

// Sorted-by-material dispatch, step 1: bucket the hits of this bounce by closure kind (the reference sorts its
// SURFACE queue by surface tag, wave_path_v2.cpp:891-928,1255-1260, with a one-thread prefix sum; here a near-stable
// block-aggregated partition).  Escaped rays are dropped, so the shade kernels only ever see real work; every bucket
// keeps the ray-queue order inside a block chunk, which keeps the shade kernels' gathers coalesced.
//   kind 0: hit has no surface (emitter only)   kind 1: Matte closure   kind 2: Disney closure
//   kinds 3..6: Mirror, Glass, Plastic, Metal (MicrofacetFamilyClosure<type>, kind = type + 1)   kind 7: Mix
//   kind 8: transmissive Disney closure ("disney_trans": LRK_SURFACE_DISNEY_TRANSMISSIVE records)
constexpr uint32_t kHitKinds = 9u;
__global__ void __launch_bounds__(kBlock) classify_hits_kernel(DeviceScene sc, PathBuffers pb, uint32_t depth) {
    __shared__ uint32_t s_warp[kHitKinds][kBlock / 32];
    __shared__ uint32_t s_base[kHitKinds];
    const uint32_t n = pb.counts[depth];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5u, lane_lt = (1u << lane) - 1u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t kind = ~0u;
        if (i < n) {
            const uint32_t inst = pb.hit[i].x;
            if (inst != ~0u) kind = __ldg(sc.inst_kind + inst);
        }
        uint32_t masks[kHitKinds];
#pragma unroll
        for (uint32_t k = 0; k < kHitKinds; k++) {
            masks[k] = __ballot_sync(0xffffffffu, kind == k);
            if (lane == 0u) s_warp[k][warp] = __popc(masks[k]);
        }
        __syncthreads();
        if (threadIdx.x < kHitKinds) {
            const uint32_t k = threadIdx.x;
            uint32_t total = 0u;
            for (int w = 0; w < kBlock / 32; w++) {
                uint32_t c = s_warp[k][w];
                s_warp[k][w] = total;
                total += c;
            }
            s_base[k] = total ? atomicAdd(pb.counts + (4u + k) * kMaxDepthSlots + depth, total) : 0u;
        }
        __syncthreads();
        if (kind != ~0u) pb.hit_index[kind][s_base[kind] + s_warp[kind][warp] + __popc(masks[kind] & lane_lt)] = i;
        __syncthreads();
    }
}

template<bool COUNT, bool ALPHA = false>
__global__ void __launch_bounds__(kTraceBlock, LRK_TRACE_MIN_BLOCKS) trace_shadow_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ count,
                                                              uint32_t *cursor) {
    const uint32_t n = *count;
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<true, COUNT, 1, ALPHA>(sc, pb.sray_o, pb.sray_d, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
        if (finished && h.x == ~0u) {// unoccluded: add the pending next-event contribution to the path's radiance
            float4 c = pb.scontrib[i];
            uint32_t path = __float_as_uint(c.w);
            float4 li = pb.li[path];
            li.x += c.x;
            li.y += c.y;
            li.z += c.z;
            pb.li[path] = li;
        }
    });
    if (COUNT) {
        atomicAdd(pb.stats + 5, static_cast<unsigned long long>(tc.nodes));
        atomicAdd(pb.stats + 6, static_cast<unsigned long long>(tc.tris));
        atomicAdd(pb.stats + 7, static_cast<unsigned long long>(tc.xforms));
    }
}

// stand-alone queries (lrk_trace) on interleaved lrk_ray records: any-hit result is written as inst = 1 (occluded) / 0 (free)
template<bool ANY_HIT, bool ALPHA = false>
__global__ void __launch_bounds__(kTraceBlock) trace_query_kernel(DeviceScene sc, const float4 *__restrict__ rays, uint4 *__restrict__ hits,
                                                             uint32_t n, uint32_t *cursor) {
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<ANY_HIT, false, 2, ALPHA>(sc, rays, rays + 1, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
        if (ANY_HIT) h = make_uint4(h.x != ~0u ? 1u : 0u, 0u, 0u, 0u);
        if (finished) hits[i] = h;
    });
}

// ---- escaped rays -----------------------------------------------------------------------------------------
// Only launched for scenes with an environment light: environment radiance with MIS for every ray of the depth that left the
// scene (mega_path.cpp:68-75, UniformLightSamplerInstance::evaluate_miss uniform.cpp:67-76).  Without an environment the
// escaped rays are simply never looked at again (classify_hits_kernel buckets hits only).
__global__ void __launch_bounds__(kBlock) shade_miss_kernel(DeviceScene sc, PathBuffers pb, uint32_t depth) {
    const uint32_t n = pb.counts[depth];
    const int in = depth & 1u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (pb.hit[i].x != ~0u) continue;
        float4 rd = pb.ray_d[in][i];
        float4 bp = pb.beta_pdf[in][i];
        uint2 ir = pb.id_rng[in][i];
        LightEval e = environment_evaluate(*sc.self, v3(rd.x, rd.y, rd.z));
        e.pdf *= sc.env_prob;
        V3 add = v3(bp.x, bp.y, bp.z) * e.L * balance_heuristic(bp.w, e.pdf);
        float4 li = pb.li[ir.x];
        li.x += add.x;
        li.y += add.y;
        li.z += add.z;
        pb.li[ir.x] = li;
    }
}

// The numbers one bounce consumes, drawn from a table-driven sampler (row f2) for the path with generation slot `id`: pixel and
// sample index follow from the slot (PathBuffers::pass_*), `word` is the dimension counter the path carries.  Out of line, so
// that the Independent sampler's path through shade_kernel is the code it was before the table-driven samplers existed.
struct BounceDraws {
    float u_sel, ul0, ul1, u_lobe, ub0, ub1, u_rr;
    uint32_t state;
};
__device__ __noinline__ BounceDraws draw_bounce_from_tables(SamplerRef sc, const uint32_t *pixel_list, uint32_t pixel_offset, uint32_t npix,
                                                            uint32_t spp_begin, uint32_t id, uint32_t word, bool draw_rr) {
    const uint32_t pixel = __ldg(pixel_list + pixel_offset + id % npix);
    PathSampler smp;
    smp.resume(sc, word, pixel & 0xffffu, pixel >> 16u, spp_begin + id / npix);
    BounceDraws d;
    d.u_sel = smp.next1d(sc);
    const float2 ul = smp.next2d(sc);
    d.ul0 = ul.x;
    d.ul1 = ul.y;
    d.u_lobe = smp.next1d(sc);
    const float2 ub = smp.next2d(sc);
    d.ub0 = ub.x;
    d.ub1 = ub.y;
    d.u_rr = draw_rr ? smp.next1d(sc) : 0.f;
    d.state = smp.state;
    return d;
}

// Sorted-by-material dispatch, step 2: one shade kernel per closure kind, each over its own hit bucket.
template<uint32_t KIND, bool TEXTURED = false>
__global__ void __launch_bounds__(kShadeBlock, LRK_SHADE_MIN_BLOCKS) shade_kernel(DeviceScene sc, PathBuffers pb, uint32_t depth) {
    __shared__ uint32_t s_warp_next[kShadeBlock / 32], s_warp_shadow[kShadeBlock / 32];
    __shared__ uint32_t s_base_next, s_base_shadow;
    const uint32_t n = pb.counts[(4u + KIND) * kMaxDepthSlots + depth];// size of this kind's hit bucket
    const int in = depth & 1u, out = in ^ 1;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        bool push_next = false, push_shadow = false;
        float4 nro, nrd, nbeta, sro, srd, scon;
        uint2 nid;
        if (j < n) {
            const uint32_t i = pb.hit_index[KIND][j];
            const uint4 hit = pb.hit[i];
            {
                float4 ro = pb.ray_o[in][i], rd = pb.ray_d[in][i];
                float4 bp = pb.beta_pdf[in][i];
                uint2 ir = pb.id_rng[in][i];
                V3 beta = v3(bp.x, bp.y, bp.z);
                float pdf_bsdf = bp.w;
                uint32_t state = ir.y;
                V3 wo = -v3(rd.x, rd.y, rd.z);
                float bu = __uint_as_float(hit.z), bv = __uint_as_float(hit.w);
                Interaction it = make_interaction(sc, hit.x, hit.y, v3(1.f - bu - bv, bu, bv));
                it.back_facing = dot(wo, it.ng) < 0.0f;
                // emitter hit with MIS: mega_path.cpp:80-87, uniform.cpp:50-65
                if (sc.light_count != 0u && it.shape.has_light()) {
                    LightEval e = evaluate_hit(sc, it, v3(ro.x, ro.y, ro.z));
                    V3 add = beta * e.L * balance_heuristic(pdf_bsdf, e.pdf);
                    float4 li = pb.li[ir.x];
                    li.x += add.x;
                    li.y += add.y;
                    li.z += add.z;
                    pb.li[ir.x] = li;
                }
                if (KIND != 0u) {// kind 0 = emitter-only hit (no surface): the path ends here (mega_path.cpp:89)
                    // draw order is normative: mega_path.cpp:91-98
                    float u_sel, ul0, ul1, u_lobe, ub0, ub1, u_rr = 0.f;
                    if (sc.sampler_type == LRK_SAMPLER_INDEPENDENT) {// the headline path: an LCG word in a register
                        u_sel = lcg(state);
                        ul0 = lcg(state);
                        ul1 = lcg(state);
                        u_lobe = lcg(state);
                        ub0 = lcg(state);
                        ub1 = lcg(state);
                        if (depth + 1u >= sc.rr_depth) u_rr = lcg(state);
                    } else {// table-driven samplers: all of the bounce's numbers from one out-of-line call
                        BounceDraws dr = draw_bounce_from_tables(sampler_ref(sc), pb.pass_pixel_list, pb.pass_pixel_offset, pb.pass_npix, pb.pass_spp_begin,
                                                                 ir.x, state, depth + 1u >= sc.rr_depth);
                        u_sel = dr.u_sel; ul0 = dr.ul0; ul1 = dr.ul1; u_lobe = dr.u_lobe; ub0 = dr.ub0; ub1 = dr.ub1; u_rr = dr.u_rr;
                        state = dr.state;
                    }
                    LightSample ls;
                    ls.eval.L = v3(0.f);
                    ls.eval.pdf = 0.f;
                    ls.ray_o_tmin = make_float4(0.f, 0.f, 0.f, 0.f);
                    ls.ray_d_tmax = make_float4(0.f, 0.f, 1.f, 0.f);
                    if (sc.light_count != 0u || sc.env_prob != 0.f) ls = sample_light(sc, it, u_sel, ul0, ul1);
                    const lrk_surface *surf = sc.surfaces + it.shape.surface_tag;
                    V3 contrib, wi, f;
                    float pdf;
                    float eta_scale = 1.f;// mega_path.cpp:113,133-138
                    if (KIND == 1u) {
                        MatteClosure cl;
                        init_closure<TEXTURED>(sc, cl, surf, it);
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                    } else if (KIND == 2u) {
                        DisneyClosure cl;
                        init_closure<TEXTURED>(sc, cl, surf, it);
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                    } else if (KIND == 8u) {
                        DisneyTransClosure cl;
                        init_closure<TEXTURED>(sc, cl, surf, it);
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                        if (surf->lobes & LRK_DISNEY_LOBE_SPEC_TRANS) eta_scale = cl.rr_eta_scale;// closure->eta() (disney.cpp:531-533)
                    } else if (KIND == 7u) {
                        MixClosure cl;
                        cl.init(*surf, sc.surfaces);
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                        eta_scale = cl.rr_eta_scale;
                    } else {
                        MicrofacetFamilyClosure<(KIND >= 3u && KIND <= 6u) ? KIND - 1u : LRK_SURFACE_MIRROR> cl;// kind = surface type + 1
                        cl.init(*surf);// constant parameters only (include/lrk.h)
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                        eta_scale = cl.rr_eta_scale;
                    }
                    if (contrib.x != 0.f || contrib.y != 0.f || contrib.z != 0.f) {
                        // a zero (or NaN-free zero) contribution needs no shadow ray; NaNs must reach the film filter
                        push_shadow = true;
                        sro = ls.ray_o_tmin;
                        srd = ls.ray_d_tmax;
                        scon = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(ir.x));
                    }
                    // continue the path: mega_path.cpp:128-151
                    V3 po = p_robust(it, wi);
                    float w = pdf > 0.f ? 1.f / pdf : 0.f;
                    beta = beta * (w * f);
                    if (isnan(beta.x) || isnan(beta.y) || isnan(beta.z)) beta = v3(0.f);
                    bool alive = !(beta.x <= 0.f && beta.y <= 0.f && beta.z <= 0.f);
                    if (alive) {
                        float q = fmaxf(max3(beta) * eta_scale, .05f);
                        if (depth + 1u >= sc.rr_depth) {
                            if (q < sc.rr_threshold && u_rr >= q) alive = false;
                            beta = beta * (q < sc.rr_threshold ? 1.0f / q : 1.f);
                        }
                    }
                    if (alive && depth + 1u < sc.max_depth) {
                        push_next = true;
                        nro = make_float4(po.x, po.y, po.z, 0.f);
                        nrd = make_float4(wi.x, wi.y, wi.z, kFltMax);
                        nbeta = make_float4(beta.x, beta.y, beta.z, pdf);
                        nid = make_uint2(ir.x, state);
                    }
                }
            }
        }
        // block-aggregated compaction: ballots inside the warp, one atomic per queue per block iteration
        uint32_t m_next = __ballot_sync(0xffffffffu, push_next);
        uint32_t m_shadow = __ballot_sync(0xffffffffu, push_shadow);
        if (lane == 0u) {
            s_warp_next[warp] = __popc(m_next);
            s_warp_shadow[warp] = __popc(m_shadow);
        }
        __syncthreads();
        if (threadIdx.x == 0u) {
            uint32_t tn = 0u, ts = 0u;
            for (int w = 0; w < kShadeBlock / 32; w++) {
                uint32_t a = s_warp_next[w], b = s_warp_shadow[w];
                s_warp_next[w] = tn;
                s_warp_shadow[w] = ts;
                tn += a;
                ts += b;
            }
            s_base_next = tn ? atomicAdd(pb.counts + depth + 1u, tn) : 0u;
            s_base_shadow = ts ? atomicAdd(pb.counts + kMaxDepthSlots + depth, ts) : 0u;
        }
        __syncthreads();
        const uint32_t lt = (1u << lane) - 1u;
        if (push_next) {
            uint32_t slot = s_base_next + s_warp_next[warp] + __popc(m_next & lt);
            pb.ray_o[out][slot] = nro;
            pb.ray_d[out][slot] = nrd;
            pb.beta_pdf[out][slot] = nbeta;
            pb.id_rng[out][slot] = nid;
        }
        if (push_shadow) {
            uint32_t slot = s_base_shadow + s_warp_shadow[warp] + __popc(m_shadow & lt);
            pb.sray_o[slot] = sro;
            pb.sray_d[slot] = srd;
            pb.scontrib[slot] = scon;
        }
        __syncthreads();
    }
}

// ---- volume path (config C4): mega_vpt_naive.cpp:170-485 for ONE homogeneous environment medium ------------------------
// Scope and the reference quirks that are reproduced on purpose are listed next to the oracle's volume_path_li
// (oracle/oracle.cpp).  For opaque closures every transmittance ray ends at the first surface with f = 0, so both
// shadow rays of a bounce reduce to any-hit queries whose only side effect on the path is three PCG32 draws when
// they hit something (homogeneous.cpp:119-125).
//
// Wavefront schedule per depth d:   T1 any-hit(s1 rays) -> occl1 | T0 closest(main rays) -> hits |
//   volume_shade(d): PCG catch-up (occl2 of d-1, occl1 of d), distance sampling, scatter/absorb or surface shading,
//                    next ray + next bounce's s1 ray + surface NEE shadow record | T2 any-hit(shadow records) -> Li, occl2
// shadow ray from a point in the medium towards a sampled light point: LightSampler::sample with
// Interaction{ray->origin()} (mega_vpt_naive.cpp:270-273): zero offset factor, so the origin is the point itself
__device__ __forceinline__ void medium_light_shadow_ray(const DeviceScene &sc, V3 p_from, float u_sel, float u0, float u1,
                                                        float4 &ro, float4 &rd) {
    float n = static_cast<float>(sc.light_count);
    uint32_t tag = static_cast<uint32_t>(clampf(u_sel * n, 0.f, n - 1.f));
    const lrk_light_handle handle = sc.light_handles[tag];
    ShapeHandle light_inst = decode_handle(__ldg(sc.inst_handles + handle.instance_id));
    const lrk_mesh mesh = sc.meshes[light_inst.mesh];
    float u = u0 * static_cast<float>(light_inst.tri_count);
    uint32_t i = min(max(static_cast<uint32_t>(u), 0u), light_inst.tri_count - 1u);
    float u_remapped = u - floorf(u);
    lrk_alias_entry entry = sc.alias[mesh.triangle_offset + i];
    bool keep = u_remapped < entry.prob;
    uint32_t triangle_id = keep ? i : entry.alias;
    float ux = keep ? u_remapped / entry.prob : (u_remapped - entry.prob) / (1.0f - entry.prob);
    V3 uvw = sample_uniform_triangle(ux, u1);
    V3 Lv = hit_position(sc, handle.instance_id, triangle_id, uvw) - p_from;
    float d = length(Lv);
    V3 dir = Lv * (1.f / d);
    ro = make_float4(p_from.x, p_from.y, p_from.z, 0.f);
    rd = make_float4(dir.x, dir.y, dir.z, d * .9999f);
}

__global__ void __launch_bounds__(kBlock) generate_rays_volume_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ pixel_list,
                                                                      uint32_t pixel_offset, uint32_t npix, uint32_t spp_begin, uint32_t n) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id == 0u) {
        pb.counts[0] = n;
        for (uint32_t d = 1u; d < kCountSlots * kMaxDepthSlots; d++) pb.counts[d] = 0u;
    }
    if (id >= n) return;
    uint32_t k = id % npix;
    uint32_t s = id / npix;
    uint32_t pixel = __ldg(pixel_list + pixel_offset + k);
    uint32_t px = pixel & 0xffffu, py = pixel >> 16u;
    uint32_t state = xxhash32_uint4(px, py, sc.sampler_seed, spp_begin + s);
    float ux = lcg(state);
    float uy = lcg(state);
    float4 ro, rd;
    float weight;
    camera_ray(sc.camera, px, py, ux, uy, ro, rd, weight);
    // PCG32 rng(U64(as<UInt2>(generate_2d()))): first float = high word (src/util/u64.h:48,58-59)
    float s0 = lcg(state), s1 = lcg(state);
    PCG32 rng;
    rng.set_sequence((static_cast<unsigned long long>(__float_as_uint(s0)) << 32u) | __float_as_uint(s1));
    float u_rr = 0.f;
    if (1u >= sc.rr_depth) u_rr = lcg(state);
    float u_sel = lcg(state), ul0 = lcg(state), ul1 = lcg(state);
    float4 so, sd;
    medium_light_shadow_ray(sc, v3(ro.x, ro.y, ro.z), u_sel, ul0, ul1, so, sd);
    pb.ray_o[0][id] = ro;
    pb.ray_d[0][id] = rd;
    pb.beta_pdf[0][id] = make_float4(weight, weight, weight, 1e16f);
    pb.id_rng[0][id] = make_uint2(id, state);
    pb.li[id] = make_float4(0.f, 0.f, 0.f, 0.f);
    pb.pcg[0][id] = make_ulonglong2(rng.state, rng.inc);
    pb.u_rr[0][id] = u_rr;
    pb.s1ray_o[id] = so;
    pb.s1ray_d[id] = sd;
    pb.occl2[0][id] = 0u;
}

// T1 / T2: any-hit queries that record occlusion.  T2 also applies the pending surface NEE contribution.
template<bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) trace_medium_shadow_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ count,
                                                                     uint32_t *cursor) {
    const uint32_t n = *count;
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<true, COUNT, 1>(sc, pb.s1ray_o, pb.s1ray_d, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
        if (finished) pb.occl1[i] = h.x != ~0u ? 1u : 0u;
    });
    if (COUNT) {
        atomicAdd(pb.stats + 5, static_cast<unsigned long long>(tc.nodes));
        atomicAdd(pb.stats + 6, static_cast<unsigned long long>(tc.tris));
        atomicAdd(pb.stats + 7, static_cast<unsigned long long>(tc.xforms));
    }
}

template<bool COUNT>
__global__ void __launch_bounds__(kTraceBlock) trace_volume_nee_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ count,
                                                                  uint32_t *cursor, uint32_t *__restrict__ occl_out) {
    const uint32_t n = *count;
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<true, COUNT, 1>(sc, pb.sray_o, pb.sray_d, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
        if (!finished) return;
        const bool occluded = h.x != ~0u;
        const uint32_t target = pb.s2_target[i];
        if (target != ~0u) occl_out[target] = occluded ? 1u : 0u;
        if (!occluded) {
            float4 c = pb.scontrib[i];
            if (c.x != 0.f || c.y != 0.f || c.z != 0.f) {
                uint32_t path = __float_as_uint(c.w);
                float4 li = pb.li[path];
                li.x += c.x;
                li.y += c.y;
                li.z += c.z;
                pb.li[path] = li;
            }
        }
    });
    if (COUNT) {
        atomicAdd(pb.stats + 5, static_cast<unsigned long long>(tc.nodes));
        atomicAdd(pb.stats + 6, static_cast<unsigned long long>(tc.tris));
        atomicAdd(pb.stats + 7, static_cast<unsigned long long>(tc.xforms));
    }
}

// What a path carries into the next depth of the volume integrator (the queues of the next wave).
struct VolumeNext {
    bool push;
    float4 ro, rd, beta, s1o, s1d;
    uint2 id;
    ulonglong2 pcg;
    float u_rr;
};

// End of one loop iteration of mega_vpt_naive.cpp:439-452 (NaN guard, Russian roulette) and, for the survivors, the draws at
// the top of the next iteration (:256-273): u_rr, then the in-medium light sample whose shadow ray the next wave traces.
__device__ __forceinline__ void volume_continue(const DeviceScene &sc, uint32_t depth, bool alive, V3 beta, float pdf_bsdf, float u_rr,
                                                V3 next_o, V3 next_d, uint32_t path_id, uint32_t state, const PCG32 &rng, VolumeNext &nx) {
    nx.push = false;
    nx.u_rr = 0.f;
    if (alive) {
        if (isnan(beta.x) || isnan(beta.y) || isnan(beta.z)) beta = v3(0.f);
        alive = !(beta.x <= 0.f && beta.y <= 0.f && beta.z <= 0.f);
        if (alive) {
            float q = fmaxf(max3(beta) * 1.f, .05f);
            if (depth + 1u >= sc.rr_depth) {
                if (q < sc.rr_threshold && u_rr >= q) alive = false;
                beta = beta * (q < sc.rr_threshold ? 1.0f / q : 1.f);
            }
        }
    }
    if (alive && depth + 1u < sc.max_depth) {
        nx.push = true;
        if (depth + 2u >= sc.rr_depth) nx.u_rr = lcg(state);
        float u_sel = lcg(state), ul0 = lcg(state), ul1 = lcg(state);
        medium_light_shadow_ray(sc, next_o, u_sel, ul0, ul1, nx.s1o, nx.s1d);
        nx.ro = make_float4(next_o.x, next_o.y, next_o.z, 0.f);
        nx.rd = make_float4(next_d.x, next_d.y, next_d.z, kFltMax);
        nx.beta = make_float4(beta.x, beta.y, beta.z, pdf_bsdf);
        nx.id = make_uint2(path_id, state);
        nx.pcg = make_ulonglong2(rng.state, rng.inc);
    }
}

__device__ __forceinline__ void volume_store_next(const PathBuffers &pb, int out, uint32_t slot, const VolumeNext &nx) {
    pb.ray_o[out][slot] = nx.ro;
    pb.ray_d[out][slot] = nx.rd;
    pb.beta_pdf[out][slot] = nx.beta;
    pb.id_rng[out][slot] = nx.id;
    pb.pcg[out][slot] = nx.pcg;
    pb.u_rr[out][slot] = nx.u_rr;
    pb.s1ray_o[slot] = nx.s1o;
    pb.s1ray_d[slot] = nx.s1d;
    pb.occl2[out][slot] = 0u;
}

// Volume wave, step 1 (every path of the depth): advance the PCG32 stream by the occlusion results, sample the medium
// along the ray (homogeneous.cpp:48-118).  Absorption / scattering events finish here; paths that reach their surface hit
// (event 3) write their updated throughput, pdf, PCG state and the MOVED ray origin back in place and are appended to the
// hit bucket of their closure kind - the same material sort as the surface integrator - for volume_surface_kernel.
__global__ void __launch_bounds__(kBlock) volume_medium_kernel(DeviceScene sc, PathBuffers pb, uint32_t depth) {
    constexpr uint32_t kLists = 1u + kHitKinds;// list 0 = next wave, 1 + k = hit bucket k
    __shared__ uint32_t s_warp[kLists][kBlock / 32];
    __shared__ uint32_t s_base[kLists];
    const uint32_t n = pb.counts[depth];
    const int in = depth & 1u, out = in ^ 1;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5u, lane_lt = (1u << lane) - 1u;
    const V3 sigma_a = v3(sc.sigma_a[0], sc.sigma_a[1], sc.sigma_a[2]), sigma_s = v3(sc.sigma_s[0], sc.sigma_s[1], sc.sigma_s[2]);
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t list = ~0u;// which list this item is appended to, if any
        VolumeNext nx;
        nx.push = false;
        if (i < n) {
            const uint4 hit = pb.hit[i];
            float4 ro = pb.ray_o[in][i], rd = pb.ray_d[in][i];
            float4 bp = pb.beta_pdf[in][i];
            uint2 ir = pb.id_rng[in][i];
            ulonglong2 pc = pb.pcg[in][i];
            PCG32 rng{pc.x, pc.y};
            // transmittance rays that hit a surface consumed three draws each: previous bounce's surface NEE, then this
            // bounce's in-medium direct light (their contribution is f = Tr * bsdf(-d, d) = 0 for opaque closures)
            if (pb.occl2[in][i] != 0u) { rng.uniform_uint(); rng.uniform_uint(); rng.uniform_uint(); }
            if (pb.occl1[i] != 0u) { rng.uniform_uint(); rng.uniform_uint(); rng.uniform_uint(); }
            V3 beta = v3(bp.x, bp.y, bp.z);
            V3 o = v3(ro.x, ro.y, ro.z), d = v3(rd.x, rd.y, rd.z);
            const bool valid = hit.x != ~0u;
            const float bu = __uint_as_float(hit.z), bv = __uint_as_float(hit.w);
            float t_max = kFltMax;
            if (valid) t_max = length(hit_position(sc, hit.x, hit.y, v3(1.f - bu - bv, bu, bv)) - o);
            uint32_t event;
            V3 mf, no, nd;
            float mpdf;
            homogeneous_medium_sample(sigma_a, sigma_s, sc.medium_g, o, d, t_max, rng, event, mf, mpdf, no, nd);
            {
                float w = mpdf > 0.f ? 1.f / mpdf : 0.f;
                beta = beta * (mf * w);
            }
            if (event == 3u) {
                if (valid) {// surface event: handed to volume_surface_kernel<kind>
                    list = 1u + __ldg(sc.inst_kind + hit.x);
                    pb.ray_o[in][i] = make_float4(no.x, no.y, no.z, ro.w);
                    pb.beta_pdf[in][i] = make_float4(beta.x, beta.y, beta.z, mpdf);
                    pb.pcg[in][i] = make_ulonglong2(rng.state, rng.inc);
                }
            } else {
                volume_continue(sc, depth, true, beta, mpdf, pb.u_rr[in][i], no, nd, ir.x, ir.y, rng, nx);
                if (nx.push) list = 0u;
            }
        }
        uint32_t masks[kLists];
#pragma unroll
        for (uint32_t k = 0; k < kLists; k++) {
            masks[k] = __ballot_sync(0xffffffffu, list == k);
            if (lane == 0u) s_warp[k][warp] = __popc(masks[k]);
        }
        __syncthreads();
        if (threadIdx.x < kLists) {
            const uint32_t k = threadIdx.x;
            uint32_t total = 0u;
            for (int w = 0; w < kBlock / 32; w++) {
                uint32_t c = s_warp[k][w];
                s_warp[k][w] = total;
                total += c;
            }
            uint32_t *counter = k == 0u ? pb.counts + depth + 1u : pb.counts + (4u + (k - 1u)) * kMaxDepthSlots + depth;
            s_base[k] = total ? atomicAdd(counter, total) : 0u;
        }
        __syncthreads();
        if (list != ~0u) {
            uint32_t slot = s_base[list] + s_warp[list][warp];
#pragma unroll
            for (uint32_t k = 0; k < kLists; k++)
                if (list == k) slot += __popc(masks[k] & lane_lt);
            if (list == 0u) volume_store_next(pb, out, slot, nx);
            else pb.hit_index[list - 1u][slot] = i;
        }
        __syncthreads();
    }
}

// Volume wave, step 2 (surface events of one closure kind): emitter hit seen from the moved origin, surface NEE +
// closure sample (mega_vpt_naive.cpp:300-437), then the common end of the iteration.
template<uint32_t KIND, bool TEXTURED = false>
__global__ void __launch_bounds__(kBlock, 2) volume_surface_kernel(DeviceScene sc, PathBuffers pb, uint32_t depth) {
    __shared__ uint32_t s_warp_next[kBlock / 32], s_warp_shadow[kBlock / 32];
    __shared__ uint32_t s_base_next, s_base_shadow;
    const uint32_t n = pb.counts[(4u + KIND) * kMaxDepthSlots + depth];
    const int in = depth & 1u, out = in ^ 1;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5u;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t j = base + threadIdx.x;
        bool push_shadow = false;
        float4 sro, srd, scon;
        VolumeNext nx;
        nx.push = false;
        if (j < n) {
            const uint32_t i = pb.hit_index[KIND][j];
            const uint4 hit = pb.hit[i];
            float4 ro = pb.ray_o[in][i], rd = pb.ray_d[in][i];// ro = the origin moved onto the surface by the medium step
            float4 bp = pb.beta_pdf[in][i];
            uint2 ir = pb.id_rng[in][i];
            ulonglong2 pc = pb.pcg[in][i];
            PCG32 rng{pc.x, pc.y};
            V3 beta = v3(bp.x, bp.y, bp.z);
            float pdf_bsdf = bp.w;
            uint32_t state = ir.y;
            V3 no = v3(ro.x, ro.y, ro.z), d = v3(rd.x, rd.y, rd.z);
            const float bu = __uint_as_float(hit.z), bv = __uint_as_float(hit.w);
            Interaction it = make_interaction(sc, hit.x, hit.y, v3(1.f - bu - bv, bu, bv));
            it.back_facing = dot(-d, it.ng) < 0.0f;
            if (it.shape.has_light()) {// evaluate_hit from the MOVED ray origin (mega_vpt_naive.cpp:308,319)
                LightEval e = evaluate_hit(sc, it, no);
                V3 add = beta * e.L * balance_heuristic(pdf_bsdf, e.pdf);
                float4 li = pb.li[ir.x];
                li.x += add.x;
                li.y += add.y;
                li.z += add.z;
                pb.li[ir.x] = li;
            }
            if (KIND != 0u) {
                float u_sel = lcg(state);
                float ul0 = lcg(state), ul1 = lcg(state);
                float u_lobe = lcg(state);
                float ub0 = lcg(state), ub1 = lcg(state);
                LightSample ls = sample_light(sc, it, u_sel, ul0, ul1);
                const lrk_surface *surf = sc.surfaces + it.shape.surface_tag;
                V3 wo = -d;
                V3 contrib = v3(0.f), wi, f;
                float pdf;
                // true_hit(medium_tag = 0) <=> 0 <= priority of the environment medium: always true (medium_tracker.cpp:19-21)
                if (KIND == 1u) {
                    MatteClosure cl;
                    init_closure<TEXTURED>(sc, cl, surf, it);
                    shade_surface<true>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                } else {
                    DisneyClosure cl;
                    init_closure<TEXTURED>(sc, cl, surf, it);
                    shade_surface<true>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                }
                push_shadow = true;// traced even with a zero contribution: its occlusion advances the PCG stream of the next bounce
                sro = ls.ray_o_tmin;
                srd = ls.ray_d_tmax;
                scon = make_float4(contrib.x, contrib.y, contrib.z, __uint_as_float(ir.x));
                V3 next_o = p_robust(it, wi);
                float w = pdf > 0.f ? 1.f / pdf : 0.f;
                beta = beta * (w * f);
                volume_continue(sc, depth, true, beta, pdf, pb.u_rr[in][i], next_o, wi, ir.x, state, rng, nx);
                if (!nx.push && scon.x == 0.f && scon.y == 0.f && scon.z == 0.f) push_shadow = false;// nothing depends on it
            }
        }
        uint32_t m_next = __ballot_sync(0xffffffffu, nx.push);
        uint32_t m_shadow = __ballot_sync(0xffffffffu, push_shadow);
        if (lane == 0u) {
            s_warp_next[warp] = __popc(m_next);
            s_warp_shadow[warp] = __popc(m_shadow);
        }
        __syncthreads();
        if (threadIdx.x == 0u) {
            uint32_t tn = 0u, ts = 0u;
            for (int w = 0; w < kBlock / 32; w++) {
                uint32_t a = s_warp_next[w], b = s_warp_shadow[w];
                s_warp_next[w] = tn;
                s_warp_shadow[w] = ts;
                tn += a;
                ts += b;
            }
            s_base_next = tn ? atomicAdd(pb.counts + depth + 1u, tn) : 0u;
            s_base_shadow = ts ? atomicAdd(pb.counts + kMaxDepthSlots + depth, ts) : 0u;
        }
        __syncthreads();
        const uint32_t lt = (1u << lane) - 1u;
        uint32_t next_slot = ~0u;
        if (nx.push) {
            next_slot = s_base_next + s_warp_next[warp] + __popc(m_next & lt);
            volume_store_next(pb, out, next_slot, nx);
        }
        if (push_shadow) {
            uint32_t slot = s_base_shadow + s_warp_shadow[warp] + __popc(m_shadow & lt);
            pb.sray_o[slot] = sro;
            pb.sray_d[slot] = srd;
            pb.scontrib[slot] = scon;
            pb.s2_target[slot] = next_slot;
        }
        __syncthreads();
    }
}

// ---- film ---------------------------------------------------------------------------------------------------
// One thread per pixel of the chunk: adds the S samples of this pass in sample order (deterministic).
// Per-sample clamp / NaN filter: src/films/color.cpp:107-130 with effective_spp = 1.
__global__ void __launch_bounds__(kBlock) accumulate_kernel(DeviceScene sc, const float4 *__restrict__ li, float4 *__restrict__ film,
                                                            const uint32_t *__restrict__ pixel_list, uint32_t pixel_offset, uint32_t npix,
                                                            uint32_t spp, const uint32_t *__restrict__ counts, unsigned long long *stats) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0u) {// ray totals of this pass = sum of the per-depth queue sizes
        unsigned long long closest = 0ull, shadow = 0ull;
        for (uint32_t d = 0; d < sc.max_depth; d++) {
            closest += counts[d];
            shadow += counts[kMaxDepthSlots + d];
        }
        stats[0] += closest;
        stats[1] += shadow;
    }
    if (k >= npix) return;
    uint32_t pixel = __ldg(pixel_list + pixel_offset + k);
    uint32_t px = pixel & 0xffffu, py = pixel >> 16u;
    size_t pid = static_cast<size_t>(py) * sc.width + px;
    float4 acc = film[pid];
    const float threshold = sc.film_clamp * fmaxf(1.f, 1.f);
    for (uint32_t s = 0; s < spp; s++) {
        float4 v = li[static_cast<size_t>(s) * npix + k];
        V3 rgb = v3(v.x, v.y, v.z) * 1.0f;// shutter weight
        bool bad = isnan(rgb.x) || isnan(rgb.y) || isnan(rgb.z) || isinf(rgb.x) || isinf(rgb.y) || isinf(rgb.z);
        if (bad) continue;
        float strength = fmaxf(fmaxf(fmaxf(fabsf(rgb.x), fabsf(rgb.y)), fabsf(rgb.z)), 0.f);
        V3 c = rgb * (threshold / fmaxf(strength, threshold));
        if (c.x != 0.f || c.y != 0.f || c.z != 0.f) {
            acc.x += c.x;
            acc.y += c.y;
            acc.z += c.z;
        }
        acc.w += 1.f;
    }
    film[pid] = acc;
}

// convert_image: src/films/color.cpp:87-93
__global__ void __launch_bounds__(kBlock) convert_film_kernel(DeviceScene sc, const float4 *__restrict__ raw, float4 *__restrict__ out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 c = raw[i];
    float nrm = fmaxf(c.w, 1.f);
    float inv = 1.f / nrm;
    out[i] = make_float4((inv * sc.film_scale[0]) * c.x, (inv * sc.film_scale[1]) * c.y, (inv * sc.film_scale[2]) * c.z, 1.f);
}

}// namespace lrk

#include "volume_general.cuh"
