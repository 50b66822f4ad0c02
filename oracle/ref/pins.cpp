// pins.cpp — C-ABI around the REFERENCE's own functions (TEST INFRASTRUCTURE; see oracle/ref/README.md).
//
// Every pin below wraps one function of /root/reference/src/util (compiled from the sources where they lie, see Makefile)
// in a LuisaCompute Callable; calling the reference function records its AST, and refinterp executes that AST on the host.
// Nothing numerical is written here: arguments are forwarded, results are packed into a float array.
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <luisa/dsl/sugar.h>
#include <luisa/dsl/syntax.h>

#include <util/frame.h>
#include <util/loop_subdiv.h>
#include <util/rng.h>
#include <util/sampling.h>
#include <util/scattering.h>
#include <util/spec.h>
#include <util/u64.h>

#include "interp.h"
#include "pins.h"

using namespace luisa;
using namespace luisa::compute;
using namespace luisa::render;

using namespace refpins;

namespace refpins {
std::map<std::string, std::function<Pin()>> &factories() {
    static std::map<std::string, std::function<Pin()>> f;
    return f;
}
}// namespace refpins

namespace {

std::map<std::string, Pin> &built() {
    static std::map<std::string, Pin> b;
    return b;
}
thread_local std::string g_error;

void register_all() {
    static bool done = false;
    if (done) { return; }
    done = true;
    /* ---- src/util/rng.cpp ---- */
    add("xxhash32_1", [](UInt p) { return xxhash32(p); });
    add("xxhash32_2", [](UInt2 p) { return xxhash32(p); });
    add("xxhash32_3", [](UInt3 p) { return xxhash32(p); });
    add("xxhash32_4", [](UInt4 p) { return xxhash32(p); });
    add("pcg", [](UInt p) { return pcg(p); });
    add("pcg2d", [](UInt2 p) { return pcg2d(p); });
    add("pcg3d", [](UInt3 p) { return pcg3d(p); });
    add("pcg4d", [](UInt4 p) { return pcg4d(p); });
    add("uniform_uint_to_float", [](UInt u) { return uniform_uint_to_float(u); });
    add("lcg", [](UInt state) {
        auto u = lcg(state);
        return pack(u, bits(state));
    });
    add("pcg32_seq", [](UInt2 seq) {// PCG32(U64 seq_index) as mega_vpt_naive.cpp:184 constructs it, then 4 uints + 2 floats
        PCG32 rng{U64{seq}};
        auto a = rng.uniform_uint();
        auto b = rng.uniform_uint();
        auto c = rng.uniform_uint();
        auto d = rng.uniform_uint();
        auto e = rng.uniform_float();
        auto f = rng.uniform_float();
        return pack(bits(a), bits(b), bits(c), bits(d), e, f, bits(rng.state().hi()), bits(rng.state().lo()),
                    bits(rng.inc().hi()), bits(rng.inc().lo()));
    });
    add("u64_mul", [](UInt2 a, UInt2 b) { return (U64{a} * U64{b}).bits(); });
    add("u64_add", [](UInt2 a, UInt2 b) { return (U64{a} + U64{b}).bits(); });
    add("u64_shr", [](UInt2 a, UInt s) { return (U64{a} >> s).bits(); });
    add("u64_shl", [](UInt2 a, UInt s) { return (U64{a} << s).bits(); });
    /* ---- src/util/sampling.cpp ---- */
    add("sample_uniform_triangle", [](Float2 u) { return sample_uniform_triangle(u); });
    add("sample_uniform_disk_concentric", [](Float2 u) { return sample_uniform_disk_concentric(u); });
    add("sample_cosine_hemisphere", [](Float2 u) { return sample_cosine_hemisphere(u); });
    add("cosine_hemisphere_pdf", [](Float c) { return cosine_hemisphere_pdf(c); });
    add("sample_uniform_sphere", [](Float2 u) { return sample_uniform_sphere(u); });
    add("invert_uniform_sphere_sample", [](Float3 w) { return invert_uniform_sphere_sample(w); });
    add("uniform_cone_pdf", [](Float c) { return uniform_cone_pdf(c); });
    add("sample_uniform_cone", [](Float2 u, Float c) { return sample_uniform_cone(u, c); });
    add("balance_heuristic", [](Float f, Float g) { return balance_heuristic(f, g); });
    add("power_heuristic", [](Float f, Float g) { return power_heuristic(f, g); });
    add("sample_discrete_3", [](Float3 w, Float u) { return sample_discrete(w, u); });
    add("sample_exponential", [](Float u, Float a) { return sample_exponential(u, a); });
    add("sample_alias_table", [](BufferVar<AliasEntry> table, UInt n, Float u) {
        auto [index, uu] = sample_alias_table(table, n, u);
        return pack(bits(index), uu);
    });
    /* ---- src/util/frame.cpp ---- */
    add("frame_make_n", [](Float3 n) {
        auto f = Frame::make(n);
        return pack(f.s().x, f.s().y, f.s().z, f.t().x, f.t().y, f.t().z, f.n().x, f.n().y, f.n().z);
    });
    add("frame_make_ns", [](Float3 n, Float3 s) {
        auto f = Frame::make(n, s);
        return pack(f.s().x, f.s().y, f.s().z, f.t().x, f.t().y, f.t().z, f.n().x, f.n().y, f.n().z);
    });
    add("frame_local_to_world", [](Float3 s, Float3 t, Float3 n, Float3 d) { return Frame{s, t, n}.local_to_world(d); });
    add("frame_world_to_local", [](Float3 s, Float3 t, Float3 n, Float3 d) { return Frame{s, t, n}.world_to_local(d); });
    add("clamp_shading_normal", [](Float3 ns, Float3 ng, Float3 w) { return clamp_shading_normal(ns, ng, w); });
    /* ---- src/util/scattering.cpp ---- */
    add("refract", [](Float3 wi, Float3 n, Float eta) {
        Float3 wt = make_float3(0.f);
        auto valid = refract(wi, n, eta, &wt);
        return pack(flag(valid), wt.x, wt.y, wt.z);
    });
    add("face_forward", [](Float3 v, Float3 n) { return face_forward(v, n); });
    add("spherical_direction", [](Float s, Float c, Float phi) { return spherical_direction(s, c, phi); });
    add("spherical_theta", [](Float3 v) { return spherical_theta(v); });
    add("spherical_phi", [](Float3 v) { return spherical_phi(v); });
    add("tr_roughness_to_alpha", [](Float r) { return TrowbridgeReitzDistribution::roughness_to_alpha(r); });
    add("tr_D", [](Float2 alpha, Float3 wh) { return TrowbridgeReitzDistribution{alpha}.D(wh); });
    add("tr_Lambda", [](Float2 alpha, Float3 w) { return TrowbridgeReitzDistribution{alpha}.Lambda(w); });
    add("tr_G1", [](Float2 alpha, Float3 w) { return TrowbridgeReitzDistribution{alpha}.G1(w); });
    add("tr_G", [](Float2 alpha, Float3 wo, Float3 wi) { return TrowbridgeReitzDistribution{alpha}.G(wo, wi); });
    add("tr_sample_wh", [](Float2 alpha, Float3 wo, Float2 u) { return TrowbridgeReitzDistribution{alpha}.sample_wh(wo, u); });
    add("tr_pdf", [](Float2 alpha, Float3 wo, Float3 wh) { return TrowbridgeReitzDistribution{alpha}.pdf(wo, wh); });
    add("fresnel_dielectric", [](Float c, Float ei, Float et) { return fresnel_dielectric(c, ei, et); });
    add("fresnel_conductor", [](Float c, Float ei, Float3 et, Float3 k) { return pack_spec(fresnel_conductor(c, ei, spec3(et), spec3(k))); });
    add("fresnel_dielectric_integral", [](Float eta) { return fresnel_dielectric_integral(eta); });
    add("lambert_reflection_evaluate", [](Float3 r, Float3 wo, Float3 wi) {
        return pack_spec(LambertianReflection{spec3(r)}.evaluate(wo, wi, TransportMode::RADIANCE));
    });
    add("lambert_reflection_sample", [](Float3 r, Float3 wo, Float2 u) {
        LambertianReflection bxdf{spec3(r)};
        Float3 wi = make_float3(0.f);
        Float pdf = 0.f;
        auto f = bxdf.sample(wo, &wi, u, &pdf, TransportMode::RADIANCE);
        return pack(wi.x, wi.y, wi.z, pdf, f[0u], f[1u], f[2u]);
    });
    add("lambert_reflection_pdf", [](Float3 r, Float3 wo, Float3 wi) { return LambertianReflection{spec3(r)}.pdf(wo, wi, TransportMode::RADIANCE); });
    add("oren_nayar_evaluate", [](Float3 r, Float sigma, Float3 wo, Float3 wi) {
        return pack_spec(OrenNayar{spec3(r), sigma}.evaluate(wo, wi, TransportMode::RADIANCE));
    });
    add("microfacet_reflection_dielectric_evaluate", [](Float3 r, Float2 alpha, Float ei, Float et, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        FresnelDielectric fr{ei, et};
        return pack_spec(MicrofacetReflection{spec3(r), &d, &fr}.evaluate(wo, wi, TransportMode::RADIANCE));
    });
    add("microfacet_reflection_dielectric_sample", [](Float3 r, Float2 alpha, Float ei, Float et, Float3 wo, Float2 u) {
        TrowbridgeReitzDistribution d{alpha};
        FresnelDielectric fr{ei, et};
        MicrofacetReflection bxdf{spec3(r), &d, &fr};
        Float3 wi = make_float3(0.f);
        Float pdf = 0.f;
        auto f = bxdf.sample(wo, &wi, u, &pdf, TransportMode::RADIANCE);
        return pack(wi.x, wi.y, wi.z, pdf, f[0u], f[1u], f[2u]);
    });
    add("microfacet_reflection_dielectric_pdf", [](Float3 r, Float2 alpha, Float ei, Float et, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        FresnelDielectric fr{ei, et};
        return MicrofacetReflection{spec3(r), &d, &fr}.pdf(wo, wi, TransportMode::RADIANCE);
    });
    add("microfacet_reflection_conductor_evaluate", [](Float3 r, Float2 alpha, Float3 eta, Float3 k, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        FresnelConductor fr{1.f, spec3(eta), spec3(k)};
        return pack_spec(MicrofacetReflection{spec3(r), &d, &fr}.evaluate(wo, wi, TransportMode::RADIANCE));
    });
    add("microfacet_reflection_conductor_sample", [](Float3 r, Float2 alpha, Float3 eta, Float3 k, Float3 wo, Float2 u) {
        TrowbridgeReitzDistribution d{alpha};
        FresnelConductor fr{1.f, spec3(eta), spec3(k)};
        MicrofacetReflection bxdf{spec3(r), &d, &fr};
        Float3 wi = make_float3(0.f);
        Float pdf = 0.f;
        auto f = bxdf.sample(wo, &wi, u, &pdf, TransportMode::RADIANCE);
        return pack(wi.x, wi.y, wi.z, pdf, f[0u], f[1u], f[2u]);
    });
    add("microfacet_transmission_evaluate", [](Float3 t, Float2 alpha, Float ea, Float eb, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        return pack_spec(MicrofacetTransmission{spec3(t), &d, ea, eb}.evaluate(wo, wi, TransportMode::RADIANCE));
    });
    add("microfacet_transmission_sample", [](Float3 t, Float2 alpha, Float ea, Float eb, Float3 wo, Float2 u) {
        TrowbridgeReitzDistribution d{alpha};
        MicrofacetTransmission bxdf{spec3(t), &d, ea, eb};
        Float3 wi = make_float3(0.f);
        Float pdf = 0.f;
        auto f = bxdf.sample(wo, &wi, u, &pdf, TransportMode::RADIANCE);
        return pack(wi.x, wi.y, wi.z, pdf, f[0u], f[1u], f[2u]);
    });
    add("microfacet_transmission_pdf", [](Float3 t, Float2 alpha, Float ea, Float eb, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        return MicrofacetTransmission{spec3(t), &d, ea, eb}.pdf(wo, wi, TransportMode::RADIANCE);
    });
    add("fresnel_blend_evaluate", [](Float3 rd, Float3 rs, Float2 alpha, Float ratio, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        return pack_spec(FresnelBlend{spec3(rd), spec3(rs), &d, ratio}.evaluate(wo, wi, TransportMode::RADIANCE));
    });
    add("fresnel_blend_sample", [](Float3 rd, Float3 rs, Float2 alpha, Float ratio, Float3 wo, Float2 u) {
        TrowbridgeReitzDistribution d{alpha};
        FresnelBlend bxdf{spec3(rd), spec3(rs), &d, ratio};
        Float3 wi = make_float3(0.f);
        Float pdf = 0.f;
        auto f = bxdf.sample(wo, &wi, u, &pdf, TransportMode::RADIANCE);
        return pack(wi.x, wi.y, wi.z, pdf, f[0u], f[1u], f[2u]);
    });
    add("fresnel_blend_pdf", [](Float3 rd, Float3 rs, Float2 alpha, Float ratio, Float3 wo, Float3 wi) {
        TrowbridgeReitzDistribution d{alpha};
        return FresnelBlend{spec3(rd), spec3(rs), &d, ratio}.pdf(wo, wi, TransportMode::RADIANCE);
    });
}

const Pin &get(const char *name) {
    register_all();
    auto it = built().find(name);
    if (it != built().end()) { return it->second; }
    auto f = factories().find(name);
    if (f == factories().end()) { throw std::runtime_error(std::string{"unknown pin '"} + name + "'"); }
    return built().emplace(name, f->second()).first->second;
}

/* 32-bit words <-> typed bytes (bool = one word; vectors / matrices / arrays / structs lane by lane) */
size_t word_count(const Type *t) {
    if (t == nullptr) { return 0u; }
    if (t->is_scalar()) { return 1u; }
    if (t->is_vector()) { return t->dimension(); }
    if (t->is_matrix()) { return t->dimension() * t->dimension(); }
    if (t->is_array()) { return t->dimension() * word_count(t->element()); }
    if (t->is_structure()) {
        auto n = size_t{0u};
        for (auto m : t->members()) { n += word_count(m); }
        return n;
    }
    throw std::runtime_error("unsupported argument type");
}
size_t align_up(size_t x, size_t a) { return (x + a - 1u) / a * a; }
void words_to_bytes(const Type *t, const uint32_t *&w, std::byte *p) {
    if (t->is_scalar()) {
        if (t->tag() == Type::Tag::BOOL) { *reinterpret_cast<bool *>(p) = *w++ != 0u; }
        else if (t->size() == 4u) { std::memcpy(p, w++, 4u); }
        else { throw std::runtime_error("unsupported scalar argument"); }
    } else if (t->is_vector()) {
        for (auto i = 0u; i < t->dimension(); i++) { words_to_bytes(t->element(), w, p + i * t->element()->size()); }
    } else if (t->is_matrix()) {
        auto n = t->dimension();
        for (auto c = 0u; c < n; c++) {
            for (auto r = 0u; r < n; r++) { std::memcpy(p + (c * (n == 3u ? 4u : n) + r) * 4u, w++, 4u); }
        }
    } else if (t->is_array()) {
        auto stride = align_up(t->element()->size(), t->element()->alignment());
        for (auto i = 0u; i < t->dimension(); i++) { words_to_bytes(t->element(), w, p + i * stride); }
    } else if (t->is_structure()) {
        auto off = size_t{0u};
        for (auto m : t->members()) {
            off = align_up(off, m->alignment());
            words_to_bytes(m, w, p + off);
            off += m->size();
        }
    }
}
void bytes_to_words(const Type *t, const std::byte *p, uint32_t *&w) {
    if (t->is_scalar()) {
        if (t->tag() == Type::Tag::BOOL) { *w++ = *reinterpret_cast<const bool *>(p) ? 1u : 0u; }
        else if (t->size() == 4u) { std::memcpy(w++, p, 4u); }
        else { throw std::runtime_error("unsupported scalar result"); }
    } else if (t->is_vector()) {
        for (auto i = 0u; i < t->dimension(); i++) { bytes_to_words(t->element(), p + i * t->element()->size(), w); }
    } else if (t->is_matrix()) {
        auto n = t->dimension();
        for (auto c = 0u; c < n; c++) {
            for (auto r = 0u; r < n; r++) { std::memcpy(w++, p + (c * (n == 3u ? 4u : n) + r) * 4u, 4u); }
        }
    } else if (t->is_array()) {
        auto stride = align_up(t->element()->size(), t->element()->alignment());
        for (auto i = 0u; i < t->dimension(); i++) { bytes_to_words(t->element(), p + i * stride, w); }
    } else if (t->is_structure()) {
        auto off = size_t{0u};
        for (auto m : t->members()) {
            off = align_up(off, m->alignment());
            bytes_to_words(m, p + off, w);
            off += m->size();
        }
    }
}

}// namespace

extern "C" {

const char *refpin_last_error() { return g_error.c_str(); }

int refpin_count() {
    register_all();
    return static_cast<int>(factories().size());
}

const char *refpin_name(int index) {
    register_all();
    auto it = factories().begin();
    std::advance(it, index);
    return it->first.c_str();
}

/* number of 32-bit input words (buffers excluded) and output words of a pin; < 0 on error */
int refpin_signature(const char *name, int *n_in, int *n_out) {
    try {
        Function f{get(name).builder.get()};
        auto in = size_t{0u};
        for (auto a : f.arguments()) {
            if (!a.is_resource()) { in += word_count(a.type()); }
        }
        *n_in = static_cast<int>(in);
        *n_out = static_cast<int>(word_count(f.return_type()));
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

/* Evaluates pin `name` `count` times: inputs `in` (count x n_in words), outputs `out` (count x n_out words).  `buffer`
 * (optional) backs the pin's single buffer argument with `buffer_count` elements. */
int refpin_eval(const char *name, const uint32_t *in, uint32_t *out, int count, void *buffer, uint64_t buffer_count) {
    try {
        Function f{get(name).builder.get()};
        for (auto n = 0; n < count; n++) {
            std::vector<refinterp::Arg> args;
            for (auto a : f.arguments()) {
                refinterp::Arg arg;
                if (a.is_resource()) {
                    auto elem = a.type()->element();
                    arg.kind = refinterp::Arg::Kind::BUFFER;
                    arg.buffer = {static_cast<std::byte *>(buffer), static_cast<size_t>(buffer_count) * align_up(elem->size(), elem->alignment())};
                } else {
                    arg.bytes.assign(a.type()->size(), std::byte{0});
                    words_to_bytes(a.type(), in, arg.bytes.data());
                }
                args.push_back(std::move(arg));
            }
            auto r = refinterp::call(f, args);
            if (f.return_type() != nullptr) { bytes_to_words(f.return_type(), r.data(), out); }
        }
        return 0;
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
}

/* host-side reference functions (plain C++, no DSL) */
int refpin_create_alias_table(const float *values, uint32_t n, float *prob, uint32_t *alias, float *pdf) {
    auto [table, p] = create_alias_table(luisa::span<const float>{values, n});// src/util/sampling.cpp:38-87
    for (auto i = 0u; i < n; i++) {
        prob[i] = table[i].prob;
        alias[i] = table[i].alias;
        pdf[i] = p[i];
    }
    return 0;
}

}// extern "C"
