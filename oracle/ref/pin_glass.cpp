// pin_glass.cpp — the reference's 'glass' surface plugin (src/surfaces/glass.cpp, compiled from where it lies) driven through
// Surface::Closure::{evaluate,sample}.  TEST INFRASTRUCTURE; see oracle/ref/README.md.
#include <base/scene_node.h>
#undef LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN
#define LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(cls)
#define LUISA_RENDER_PLUGIN_NAME "glass"
#include <surfaces/glass.cpp>

#include "pin_surface.h"

namespace luisa::render {
namespace {
using namespace refpins;
auto make_closure(const SampledWavelengths &swl, Expr<float3> kr, Expr<float3> kt, Expr<float> eta_t, Expr<float2> alpha, Expr<float> kr_ratio,
                  Expr<float3> ng, Expr<float3> ns, Expr<float3> tangent) {
    auto closure = luisa::make_unique<GlassClosure>(unused_pipeline(), swl, 0.f);
    closure->bind(GlassClosure::Context{.it = make_interaction(ng, ns, tangent), .Kr = spec3(kr), .Kt = spec3(kt), .eta_i = 1.f,
                                        .eta_t = eta_t, .dispersive = false, .alpha = alpha, .Kr_ratio = kr_ratio});
    return closure;
}
void register_pins() {
    add("glass_evaluate", [](Float3 kr, Float3 kt, Float eta_t, Float2 alpha, Float kr_ratio, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, kr, kt, eta_t, alpha, kr_ratio, ng, ns, tangent);
        return closure_evaluate(*c, wo, wi);
    });
    add("glass_sample", [](Float3 kr, Float3 kt, Float eta_t, Float2 alpha, Float kr_ratio, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, kr, kt, eta_t, alpha, kr_ratio, ng, ns, tangent);
        return closure_sample(*c, wo, u_lobe, u);
    });
}
Registrar registrar{register_pins};
}// namespace
}// namespace luisa::render
