// The closure kernels of the wavefront integrators: shade_kernel<KIND> (surface integrator: emitter MIS, light sample + shadow
// record, closure evaluate / sample, Russian roulette, compaction) and the volume integrator's medium / surface steps.
// Included by shade.cu ONLY.  shade.cu is compiled twice, with different arithmetic (see its header); LRK_SHADE_VARIANT names the
// namespace the kernels of each compilation live in, so that the two sets of __global__ functions stay distinct symbols.
#pragma once
#ifndef LRK_SHADE_VARIANT
#error "shade_kernels.cuh is compiled through shade.cu with -DLRK_SHADE_VARIANT=fast|strict"
#endif
#include "pathcommon.cuh"
#include "pathstate.cuh"
#include "samplers.cuh"

namespace lrk {
namespace LRK_SHADE_VARIANT {

// The numbers one bounce consumes, drawn from a table-driven sampler (row f2) for the path with generation slot `id`: pixel and
// sample index follow from the slot (PathBuffers::pass_*), `word` is the dimension counter the path carries.  Out of line, so
// that the Independent sampler's path through shade_kernel is the code it was before the table-driven samplers existed.
struct BounceDraws {
    float u_sel, ul0, ul1, u_lobe, ub0, ub1, u_rr;
    uint32_t state;
};
__device__ __noinline__ inline BounceDraws draw_bounce_from_tables(SamplerRef sc, const uint32_t *pixel_list, uint32_t pixel_offset, uint32_t npix,
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
                    LightEval e = evaluate_hit<TEXTURED>(sc, it, v3(ro.x, ro.y, ro.z));
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
                    if (sc.light_count != 0u || sc.env_prob != 0.f) ls = sample_light<TEXTURED>(sc, it, u_sel, ul0, ul1);
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
                    } else if (KIND == 10u) {// closure->eta() is empty (disney.cpp:779): no Russian-roulette eta scale
                        DisneyThinClosure cl;
                        init_closure<TEXTURED>(sc, cl, surf, it);
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                    } else if (KIND == 9u) {
                        LayeredClosure cl;
                        const Frame fr = closure_frame<TEXTURED>(sc, surf, it, wo);
                        cl.init(*surf, sc.surfaces, it, fr, wo);
                        shade_surface<false>(cl, it, fr, wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                        eta_scale = cl.rr_eta_scale;
                    } else if (KIND == 7u) {
                        MixClosure cl;
                        cl.init(*surf, sc.surfaces);
                        shade_surface<false>(cl, it, closure_frame<TEXTURED>(sc, surf, it, wo), wo, ls, beta, u_lobe, ub0, ub1, contrib, wi, f, pdf);
                        eta_scale = cl.rr_eta_scale;
                    } else {
                        MicrofacetFamilyClosure<(KIND >= 3u && KIND <= 6u) ? KIND - 1u : LRK_SURFACE_MIRROR> cl;// kind = surface type + 1
                        init_closure<TEXTURED>(sc, cl, surf, it);
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
                LightEval e = evaluate_hit<TEXTURED>(sc, it, no);
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
                LightSample ls = sample_light<TEXTURED>(sc, it, u_sel, ul0, ul1);
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

}// namespace LRK_SHADE_VARIANT
}// namespace lrk
