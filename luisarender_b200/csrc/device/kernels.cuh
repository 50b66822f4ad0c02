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
#include "shading.cuh"
#include "traverse.cuh"

namespace lrk {

constexpr int kBlock = 256;
constexpr uint32_t kMaxDepthSlots = 64u;// counts[0..63]: path queue size per depth, counts[64..127]: shadow queue size

struct PathBuffers {
    float4 *ray_o[2];
    float4 *ray_d[2];
    float4 *beta_pdf[2];
    uint2 *id_rng[2];
    uint4 *hit;// {inst, prim, bary} per ray of the current queue (inst == ~0u: escaped)
    uint32_t *hit_index[3];// per closure kind: indices (into the current ray queue) of the rays that hit such a surface
    float4 *sray_o;
    float4 *sray_d;
    float4 *scontrib;// rgb + path id bits
    float4 *li;
    uint32_t *counts;
    unsigned long long *stats;// [0] closest rays, [1] shadow rays, [2..4] closest nodes/tris/xforms, [5..7] shadow nodes/tris/xforms
};

// ---- Camera: src/base/filter.cpp:50-64, src/base/camera.cpp:212-224, src/cameras/pinhole.cpp:60-67 ----
__device__ __forceinline__ void sample_alias_filter(const lrk_camera *cam, float u_in, uint32_t &index, float &uu) {
    constexpr uint32_t n = LRK_FILTER_LUT_SIZE - 1u;
    float u = u_in * static_cast<float>(n);
    uint32_t i = min(max(static_cast<uint32_t>(u), 0u), n - 1u);
    float u_remapped = u - floorf(u);
    float prob = cam->filter_alias_probs[i];
    bool keep = u_remapped < prob;
    index = keep ? i : cam->filter_alias_indices[i];
    uu = keep ? u_remapped / prob : (u_remapped - prob) / (1.0f - prob);
}

__global__ void __launch_bounds__(kBlock) generate_rays_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ pixel_list,
                                                               uint32_t pixel_offset, uint32_t npix, uint32_t spp_begin, uint32_t n) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id == 0u) {
        pb.counts[0] = n;
        for (uint32_t d = 1u; d < 7u * kMaxDepthSlots; d++) pb.counts[d] = 0u;// sizes + the traversal kernels' fetch cursors
    }
    if (id >= n) return;
    uint32_t k = id % npix;
    uint32_t s = id / npix;
    uint32_t pixel = __ldg(pixel_list + pixel_offset + k);
    uint32_t px = pixel & 0xffffu, py = pixel >> 16u;
    uint32_t state = xxhash32_uint4(px, py, sc.sampler_seed, spp_begin + s);
    float ux = lcg(state);
    float uy = lcg(state);
    const lrk_camera *cam = sc.camera;
    uint32_t iy, ix;
    float fy, fx;
    sample_alias_filter(cam, ux, iy, fy);
    sample_alias_filter(cam, uy, ix, fx);
    float pdf = cam->filter_pdf[iy] * cam->filter_pdf[ix];
    float f = lerp(cam->filter_lut[ix], cam->filter_lut[ix + 1u], fx) * lerp(cam->filter_lut[iy], cam->filter_lut[iy + 1u], fy);
    float ppx = static_cast<float>(ix) + fx, ppy = static_cast<float>(iy) + fy;
    constexpr float inv_size = 1.0f / static_cast<float>(LRK_FILTER_LUT_SIZE);
    float ox = (ppx * inv_size * 2.0f - 1.0f) * cam->filter_radius + cam->filter_shift[0];
    float oy = (ppy * inv_size * 2.0f - 1.0f) * cam->filter_radius + cam->filter_shift[1];
    float weight = 1.f * (f / pdf);
    float pixel_x = static_cast<float>(px) + .5f + ox;
    float pixel_y = static_cast<float>(py) + .5f + oy;
    float rx = static_cast<float>(cam->resolution[0]), ry = static_cast<float>(cam->resolution[1]);
    float kk = cam->tan_half_fov / ry;
    float p_x = (pixel_x * 2.0f - rx) * kk;
    float p_y = (pixel_y * 2.0f - ry) * kk;
    V3 direction = normalize(v3(p_x, -p_y, -1.f));
    const float *m = cam->camera_to_world;
    V3 c0 = v3(m[0], m[4], m[8]), c1 = v3(m[1], m[5], m[9]), c2 = v3(m[2], m[6], m[10]), c3 = v3(m[3], m[7], m[11]);
    V3 o = 0.f * c0 + 0.f * c1 + 0.f * c2 + 1.f * c3;
    V3 d = normalize(direction.x * c0 + direction.y * c1 + direction.z * c2);
    pb.ray_o[0][id] = make_float4(o.x, o.y, o.z, 0.f);
    pb.ray_d[0][id] = make_float4(d.x, d.y, d.z, kFltMax);
    pb.beta_pdf[0][id] = make_float4(weight, weight, weight, 1e16f);
    pb.id_rng[0][id] = make_uint2(id, state);
    pb.li[id] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- traversal kernels -------------------------------------------------------------------------------
// counts[] layout (all zeroed by generate_rays_kernel at the start of a pass), 64 slots (one per depth) each:
//   [0] path-queue sizes  [1] shadow-queue sizes  [2] closest-hit fetch cursors  [3] shadow fetch cursors
//   [4],[5],[6] hit-bucket sizes: light-only hits, Matte hits, Disney hits
template<bool COUNT>
__global__ void __launch_bounds__(kBlock) trace_closest_kernel(DeviceScene sc, const float4 *__restrict__ ray_o,
                                                               const float4 *__restrict__ ray_d, uint4 *__restrict__ hits,
                                                               const uint32_t *__restrict__ count, uint32_t *cursor,
                                                               unsigned long long *stats) {
    const uint32_t n = *count;
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<false, COUNT, 1>(sc, ray_o, ray_d, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
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
constexpr uint32_t kHitKinds = 3u;
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

template<bool COUNT>
__global__ void __launch_bounds__(kBlock) trace_shadow_kernel(DeviceScene sc, PathBuffers pb, const uint32_t *__restrict__ count,
                                                              uint32_t *cursor) {
    const uint32_t n = *count;
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<true, COUNT, 1>(sc, pb.sray_o, pb.sray_d, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
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
template<bool ANY_HIT>
__global__ void __launch_bounds__(kBlock) trace_query_kernel(DeviceScene sc, const float4 *__restrict__ rays, uint4 *__restrict__ hits,
                                                             uint32_t n, uint32_t *cursor) {
    TraversalCounters tc{0u, 0u, 0u};
    trace_queue<ANY_HIT, false, 2>(sc, rays, rays + 1, n, cursor, tc, [&](bool finished, uint32_t i, uint4 h) {
        if (ANY_HIT) h = make_uint4(h.x != ~0u ? 1u : 0u, 0u, 0u, 0u);
        if (finished) hits[i] = h;
    });
}

// ---- shade ----------------------------------------------------------------------------------------------
template<typename Closure>
__device__ __forceinline__ void shade_surface(const Closure &cl, const Interaction &it, V3 wo, const LightSample &ls, V3 beta,
                                              float u_lobe, float ub0, float ub1, V3 &contrib, V3 &wi_world, V3 &f_over, float &pdf_bsdf) {
    V3 wo_local = it.shading.world_to_local(wo);
    contrib = v3(0.f);
    if (ls.eval.pdf > 0.0f) {
        V3 wi = v3(ls.ray_d_tmax.x, ls.ray_d_tmax.y, ls.ray_d_tmax.z);
        SurfEval ev = cl.evaluate_local(wo_local, it.shading.world_to_local(wi));
        if (!validate_surface_sides(it.ng, it.shading.n, wo, wi)) {
            ev.f = v3(0.f);
            ev.pdf = 0.f;
        }
        float w = balance_heuristic(ls.eval.pdf, ev.pdf) / ls.eval.pdf;
        contrib = w * beta * ev.f * ls.eval.L;
    }
    V3 wi_local;
    SurfEval s = cl.sample_local(wo_local, u_lobe, ub0, ub1, wi_local);
    wi_world = it.shading.local_to_world(wi_local);
    if (!validate_surface_sides(it.ng, it.shading.n, wo, wi_world)) {
        s.f = v3(0.f);
        s.pdf = 0.f;
    }
    f_over = s.f;
    pdf_bsdf = s.pdf;
}

// Sorted-by-material dispatch, step 2: one shade kernel per closure kind, each over its own hit bucket.
template<uint32_t KIND>
__global__ void __launch_bounds__(kBlock) shade_kernel(DeviceScene sc, PathBuffers pb, uint32_t depth) {
    __shared__ uint32_t s_warp_next[kBlock / 32], s_warp_shadow[kBlock / 32];
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
                    float u_sel = lcg(state);
                    float ul0 = lcg(state), ul1 = lcg(state);
                    float u_lobe = lcg(state);
                    float ub0 = lcg(state), ub1 = lcg(state);
                    float u_rr = 0.f;
                    if (depth + 1u >= sc.rr_depth) u_rr = lcg(state);
                    LightSample ls;
                    ls.eval.L = v3(0.f);
                    ls.eval.pdf = 0.f;
                    ls.ray_o_tmin = make_float4(0.f, 0.f, 0.f, 0.f);
                    ls.ray_d_tmax = make_float4(0.f, 0.f, 1.f, 0.f);
                    if (sc.light_count != 0u) ls = sample_light(sc, it, u_sel, ul0, ul1);
                    const lrk_surface *surf = sc.surfaces + it.shape.surface_tag;
                    V3 contrib, wi, f;
                    float pdf;
                    if (KIND == 1u) {
                        MatteClosure cl;
                        cl.init(*surf);
                        shade_surface(cl, it, wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                    } else {
                        DisneyClosure cl;
                        cl.init(*surf);
                        shade_surface(cl, it, wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
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
                        float q = fmaxf(max3(beta) * 1.f, .05f);
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
