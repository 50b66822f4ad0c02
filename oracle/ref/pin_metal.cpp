// pin_metal.cpp — the reference's 'metal' surface plugin (src/surfaces/metal.cpp, compiled from where it lies) driven through
// Surface::Closure::{evaluate,sample}.  TEST INFRASTRUCTURE; see oracle/ref/README.md.
#include <base/scene_node.h>
#undef LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN
#define LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(cls)
#define LUISA_RENDER_PLUGIN_NAME "metal"
#include <surfaces/metal.cpp>

#include "pin_surface.h"

namespace luisa::render {
namespace {
using namespace refpins;
auto make_closure(const SampledWavelengths &swl, Expr<float3> n, Expr<float3> k, Expr<float3> refl, Expr<float2> alpha,
                  Expr<float3> ng, Expr<float3> ns, Expr<float3> tangent) {
    auto closure = luisa::make_unique<MetalClosure>(unused_pipeline(), swl, 0.f);
    closure->bind(MetalClosure::Context{.it = make_interaction(ng, ns, tangent), .eta_i = 1.f, .n = spec3(n), .k = spec3(k),
                                        .refl = spec3(refl), .alpha = alpha});
    return closure;
}
void register_pins() {
    add("metal_evaluate", [](Float3 n, Float3 k, Float3 refl, Float2 alpha, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, n, k, refl, alpha, ng, ns, tangent);
        return closure_evaluate(*c, wo, wi);
    });
    add("metal_sample", [](Float3 n, Float3 k, Float3 refl, Float2 alpha, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, n, k, refl, alpha, ng, ns, tangent);
        return closure_sample(*c, wo, u_lobe, u);
    });
}
Registrar registrar{register_pins};
}// namespace
}// namespace luisa::render
