// closures_host.cpp — the DEVICE shading code (luisarender_b200/csrc/device/shading.cuh) compiled for the host, so that the
// very expressions the sm_100a shade kernels evaluate can be checked against the reference pins without a GPU
// (tests/test_device_closures_on_host.py).  TEST INFRASTRUCTURE: nothing here is part of the product; the product path
// runs these closures on the GPU only.
//
// g++ sees the CUDA headers' host side: __device__ / __forceinline__ are attributes it ignores; the few device intrinsics
// the shading code uses are given their obvious host meaning below.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include <cuda_runtime.h>

template<typename T>
static inline T __ldg(const T *p) { return *p; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
#ifndef __noinline__
#define __noinline__
#endif
#include <algorithm>
using std::isinf;
using std::isnan;
using std::max;
using std::min;

#include "../../luisarender_b200/csrc/device/shading.cuh"

namespace {
using namespace lrk;

struct Words {
    const uint32_t *in;
    uint32_t *out;
    float f() { float v; std::memcpy(&v, in++, 4); return v; }
    V3 v() { float x = f(), y = f(), z = f(); return v3(x, y, z); }
    void put(float v) { std::memcpy(out++, &v, 4); }
    void put(uint32_t v) { *out++ = v; }
    void put(V3 v) { put(v.x); put(v.y); put(v.z); }
};

// what shade_surface (kernels.cuh) does with a closure, for one direction
template<typename Closure>
void run(Closure &cl, Words &w, bool is_eval) {
    V3 ng = w.v(), ns = w.v(), tg = w.v(), wo = w.v();
    Frame shading = Frame::make(ns, tg);
    V3 wo_local = shading.world_to_local(wo);
    cl.prepare(wo_local);
    if (is_eval) {
        V3 wi = w.v();
        SurfEval e = cl.evaluate_local(wo_local, shading.world_to_local(wi));
        if (!validate_surface_sides(ng, shading.n, wo, wi)) { e.f = v3(0.f); e.pdf = 0.f; }
        w.put(e.f);
        w.put(e.pdf);
    } else {
        float u_lobe = w.f(), u0 = w.f(), u1 = w.f();
        V3 wi_local;
        bool valid = cl.sample_direction(wo_local, u_lobe, u0, u1, wi_local);
        V3 wi = shading.local_to_world(wi_local);
        SurfEval e;
        e.f = v3(0.f);
        e.pdf = 0.f;
        if (valid) {
            e = cl.evaluate_local(wo_local, wi_local);
            if (!validate_surface_sides(ng, shading.n, wo, wi)) { e.f = v3(0.f); e.pdf = 0.f; }
        }
        w.put(wi);
        w.put(e.f);
        w.put(e.pdf);
        w.put(0u);// the event is carried as rr_eta_scale on the device; compared separately
    }
}
}// namespace

// same word packing as oracle_unit / oracle/ref/pins.cpp closure pins
extern "C" int device_closure_unit(const char *name_c, const uint32_t *in, uint32_t *out, int count) {
    const std::string name{name_c};
    Words w{in, out};
    const bool is_eval = name.find("_evaluate") != std::string::npos;
    for (int n = 0; n < count; n++) {
        lrk_surface sf{};
        auto take = [&](int c) { for (int i = 0; i < c; i++) sf.p[i] = w.f(); };
        if (name.rfind("matte_", 0) == 0) {
            sf.type = LRK_SURFACE_MATTE; take(4);
            MatteClosure cl; cl.init(sf); run(cl, w, is_eval);
        } else if (name.rfind("disney_", 0) == 0) {
            sf.type = LRK_SURFACE_DISNEY; take(15);
            sf.lobes = static_cast<uint32_t>(std::stoul(name.substr(name.rfind('_') + 1)));
            DisneyClosure cl; cl.init(sf); run(cl, w, is_eval);
        } else if (name.rfind("disneytrans_", 0) == 0) {// the transmissive closure class (hit bucket 8)
            sf.type = LRK_SURFACE_DISNEY; take(15);
            sf.flags |= LRK_SURFACE_DISNEY_TRANSMISSIVE;
            sf.lobes = static_cast<uint32_t>(std::stoul(name.substr(name.rfind('_') + 1)));
            DisneyTransClosure cl; cl.init(sf); run(cl, w, is_eval);
        } else if (name.rfind("disneythin_", 0) == 0) {// the thin closure class (hit bucket 10)
            sf.type = LRK_SURFACE_DISNEY; take(16);
            sf.flags |= LRK_SURFACE_DISNEY_THIN;
            sf.lobes = static_cast<uint32_t>(std::stoul(name.substr(name.rfind('_') + 1)));
            DisneyThinClosure cl; cl.init(sf); run(cl, w, is_eval);
        } else {
            if (name.rfind("mirror_", 0) == 0) { take(5); MicrofacetFamilyClosure<LRK_SURFACE_MIRROR> cl; cl.init(sf); run(cl, w, is_eval); }
            else if (name.rfind("glass_", 0) == 0) { take(10); MicrofacetFamilyClosure<LRK_SURFACE_GLASS> cl; cl.init(sf); run(cl, w, is_eval); }
            else if (name.rfind("plastic_", 0) == 0) { take(10); MicrofacetFamilyClosure<LRK_SURFACE_PLASTIC> cl; cl.init(sf); run(cl, w, is_eval); }
            else if (name.rfind("metal_", 0) == 0) { take(11); MicrofacetFamilyClosure<LRK_SURFACE_METAL> cl; cl.init(sf); run(cl, w, is_eval); }
            else return -1;
        }
    }
    return 0;
}

// The Layered closure (hit bucket 9) against the oracle's restatement, on the records of a flattened scene.
// rows of 16 floats in: pg(3), shading normal n(3) (= ng), wo(3), wi(3), u_lobe, u0, u1, unused;
// rows of 12 floats out: evaluate f(3), pdf, sample wi(3), sample f(3), sample pdf, event
extern "C" int device_layered_unit(const lrk_surface *records, uint32_t index, const float *in, float *out, int count) {
    for (int i = 0; i < count; i++, in += 16, out += 12) {
        Interaction it{};
        it.pg = v3(in[0], in[1], in[2]);
        it.ng = normalize(v3(in[3], in[4], in[5]));
        const Frame fr = Frame::make(it.ng);
        const V3 wo = normalize(v3(in[6], in[7], in[8])), wi = normalize(v3(in[9], in[10], in[11]));
        LayeredClosure cl;
        cl.init(records[index], records, it, fr, wo);
        const SurfEval e = cl.evaluate_world(wi);
        cl.sample_world(in[12], in[13], in[14]);
        out[0] = e.f.x; out[1] = e.f.y; out[2] = e.f.z; out[3] = e.pdf;
        out[4] = cl.sampled_wi.x; out[5] = cl.sampled_wi.y; out[6] = cl.sampled_wi.z;
        out[7] = cl.sampled.f.x; out[8] = cl.sampled.f.y; out[9] = cl.sampled.f.z; out[10] = cl.sampled.pdf;
        out[11] = static_cast<float>(cl.event);
    }
    return 0;
}
