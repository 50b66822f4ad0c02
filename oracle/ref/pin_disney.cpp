// pin_disney.cpp — the reference's 'disney' surface plugin (src/surfaces/disney.cpp, compiled from where it lies) driven through
// Surface::Closure::{evaluate,sample}.  TEST INFRASTRUCTURE; see oracle/ref/README.md.
#include <base/scene_node.h>
#undef LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN
#define LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(cls)
#define LUISA_RENDER_PLUGIN_NAME "disney"
#include <surfaces/disney.cpp>

#include "pin_surface.h"

namespace luisa::render {
namespace {
using namespace refpins;
/* the 15 context values in the order of lrk_surface.p for LRK_SURFACE_DISNEY (include/lrk.h) + the lobe mask */
auto make_closure(const SampledWavelengths &swl, Expr<float3> color, Expr<float4> a, Expr<float4> b, Expr<float4> c, Expr<uint> lobes_unused,
                  uint lobes, Expr<float3> ng, Expr<float3> ns, Expr<float3> tangent, bool transmissive = false, bool thin = false,
                  Expr<float> diffuse_trans = 0.f) {
    auto closure = luisa::make_unique<DisneySurfaceClosure>(unused_pipeline(), swl, 0.f, thin, transmissive);
    closure->bind(DisneyContext{.it = make_interaction(ng, ns, tangent), .color = spec3(color), .color_lum = a.x, .metallic = a.y,
                                .eta_i = 1.f, .eta_t = a.z, .roughness = a.w, .specular_tint = b.x, .anisotropic = b.y, .sheen = b.z,
                                .sheen_tint = b.w, .clearcoat = c.x, .clearcoat_gloss = c.y, .specular_trans = c.z, .flatness = c.w,
                                .diffuse_trans = diffuse_trans});
    closure->enable_lobes(lobes);
    return closure;
}
/* the lobe mask is a HOST-side constant in the reference (populate_closure, disney.cpp:968-989): one pin per mask used */
void register_mask(uint lobes) {
    auto suffix = std::to_string(lobes);
    static std::vector<std::string> names;
    names.push_back("disney_evaluate_" + suffix);
    add(names.back().c_str(), [lobes](Float3 color, Float4 a, Float4 b, Float4 c, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto cl = make_closure(swl, color, a, b, c, 0u, lobes, ng, ns, tangent);
        return closure_evaluate(*cl, wo, wi);
    });
    names.push_back("disney_sample_" + suffix);
    add(names.back().c_str(), [lobes](Float3 color, Float4 a, Float4 b, Float4 c, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto cl = make_closure(swl, color, a, b, c, 0u, lobes, ng, ns, tangent);
        return closure_sample(*cl, wo, u_lobe, u);
    });
}
/* the closure class "disney_trans" of transmissive nodes (is_transmissive, disney.cpp:61-75,1001-1007): fourth technique */
void register_transmissive_mask(uint lobes) {
    auto suffix = std::to_string(lobes);
    static std::vector<std::string> names;
    names.push_back("disneytrans_evaluate_" + suffix);
    add(names.back().c_str(), [lobes](Float3 color, Float4 a, Float4 b, Float4 c, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto cl = make_closure(swl, color, a, b, c, 0u, lobes, ng, ns, tangent, true);
        return closure_evaluate(*cl, wo, wi);
    });
    names.push_back("disneytrans_sample_" + suffix);
    add(names.back().c_str(), [lobes](Float3 color, Float4 a, Float4 b, Float4 c, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto cl = make_closure(swl, color, a, b, c, 0u, lobes, ng, ns, tangent, true);
        return closure_sample(*cl, wo, u_lobe, u);
    });
}
/* the closure class "disney_thin" of thin nodes with a transmission (is_thin, disney.cpp:61-69,925-930): five techniques; the
 * 16th context value is diffuse_trans (lrk_surface.p[15]) */
void register_thin_mask(uint lobes) {
    auto suffix = std::to_string(lobes);
    static std::vector<std::string> names;
    names.push_back("disneythin_evaluate_" + suffix);
    add(names.back().c_str(), [lobes](Float3 color, Float4 a, Float4 b, Float4 c, Float dt, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto cl = make_closure(swl, color, a, b, c, 0u, lobes, ng, ns, tangent, false, true, dt);
        return closure_evaluate(*cl, wo, wi);
    });
    names.push_back("disneythin_sample_" + suffix);
    add(names.back().c_str(), [lobes](Float3 color, Float4 a, Float4 b, Float4 c, Float dt, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto cl = make_closure(swl, color, a, b, c, 0u, lobes, ng, ns, tangent, false, true, dt);
        return closure_sample(*cl, wo, u_lobe, u);
    });
}
void register_pins() {
    constexpr auto base = disney_lobe_diffuse_bit | disney_lobe_retro_bit | disney_lobe_specular_bit;
    register_transmissive_mask(base | disney_lobe_spec_trans_bit);
    register_transmissive_mask(base | disney_lobe_sheen_bit | disney_lobe_clearcoat_bit | disney_lobe_fake_ss_bit | disney_lobe_spec_trans_bit);
    register_thin_mask(base | disney_lobe_spec_trans_bit | disney_lobe_diff_trans_bit);
    register_thin_mask(base | disney_lobe_sheen_bit | disney_lobe_clearcoat_bit | disney_lobe_fake_ss_bit | disney_lobe_spec_trans_bit | disney_lobe_diff_trans_bit);
    register_thin_mask(disney_lobe_specular_bit | disney_lobe_sheen_bit | disney_lobe_fake_ss_bit | disney_lobe_diff_trans_bit);
    register_mask(base);
    register_mask(base | disney_lobe_sheen_bit | disney_lobe_clearcoat_bit);
    register_mask(base | disney_lobe_sheen_bit | disney_lobe_clearcoat_bit | disney_lobe_fake_ss_bit);
    register_mask(disney_lobe_specular_bit);
}
Registrar registrar{register_pins};
}// namespace
}// namespace luisa::render
