// pin_mirror.cpp — the reference's 'mirror' surface plugin (src/surfaces/mirror.cpp, compiled from where it lies) driven through
// Surface::Closure::{evaluate,sample}.  TEST INFRASTRUCTURE; see oracle/ref/README.md.
#include <base/scene_node.h>
#undef LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN
#define LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(cls)
#define LUISA_RENDER_PLUGIN_NAME "mirror"
#include <surfaces/mirror.cpp>

#include "pin_surface.h"

namespace luisa::render {
namespace {
using namespace refpins;
auto make_closure(const SampledWavelengths &swl, Expr<float3> refl, Expr<float2> alpha, Expr<float3> ng, Expr<float3> ns, Expr<float3> tangent) {
    auto closure = luisa::make_unique<MirrorClosure>(unused_pipeline(), swl, 0.f);
    closure->bind(MirrorClosure::Context{.it = make_interaction(ng, ns, tangent), .refl = spec3(refl), .alpha = alpha});
    return closure;
}
void register_pins() {
    add("mirror_evaluate", [](Float3 refl, Float2 alpha, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float3 wi) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, refl, alpha, ng, ns, tangent);
        return closure_evaluate(*c, wo, wi);
    });
    add("mirror_sample", [](Float3 refl, Float2 alpha, Float3 ng, Float3 ns, Float3 tangent, Float3 wo, Float u_lobe, Float2 u) {
        SampledWavelengths swl{3u};
        auto c = make_closure(swl, refl, alpha, ng, ns, tangent);
        return closure_sample(*c, wo, u_lobe, u);
    });
}
Registrar registrar{register_pins};
}// namespace
}// namespace luisa::render
