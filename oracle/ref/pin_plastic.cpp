// pin_plastic.cpp — the reference's 'plastic' surface plugin (src/surfaces/plastic.cpp, compiled from where it lies) driven through
// Surface::Closure::{evaluate,sample}.  TEST INFRASTRUCTURE; see oracle/ref/README.md.
#include <base/scene_node.h>
#undef LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN
#define LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(cls)
#define LUISA_RENDER_PLUGIN_NAME "plastic"
#include <surfaces/plastic.cpp>

#include "pin_surface.h"

namespace luisa::render {
namespace {
using namespace refpins;
auto make_closure(const SampledWavelengths &swl, Expr<float3> kd, Expr<float> kd_weight, Expr<float3> sigma_a, Expr<float> eta, Expr<float2> alpha,
                  Expr<float3> ng, Expr<float3> ns, Expr<float3> tangent) {
    auto closure = luisa::make_unique<PlasticClosure>(unused_pipeline(), swl, 0.f);
    closure->bind(PlasticContext{.it = make_interaction(ng, ns, tangent), .Kd = spec3(kd), .Kd_weight = kd_weight,
                                 .sigma_a = spec3(sigma_a), .eta = eta, .roughness = alpha});
    return closure;
}
void register_pins() {
    add("plastic_evaluate", [](Float3 kd, Float kd_weight, Float3 sigma_a, Float eta, Float2 alpha, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, kd, kd_weight, sigma_a, eta, alpha, ng, ns, tangent);
        return closure_evaluate(*c, wo, wi);
    });
    add("plastic_sample", [](Float3 kd, Float kd_weight, Float3 sigma_a, Float eta, Float2 alpha, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, kd, kd_weight, sigma_a, eta, alpha, ng, ns, tangent);
        return closure_sample(*c, wo, u_lobe, u);
    });
}
Registrar registrar{register_pins};
}// namespace
}// namespace luisa::render
