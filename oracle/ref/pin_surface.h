// pin_surface.h — drives a reference Surface::Closure subclass through the reference's public
// Surface::Closure::{evaluate,sample} (src/base/surface.cpp:46-68), exactly the calls mega_path.cpp:122-131 makes.
// TEST INFRASTRUCTURE; see oracle/ref/README.md.  The including translation unit has already included the plugin source.
#pragma once

#include <base/interaction.h>
#include <base/surface.h>

#include "pins.h"

namespace refpins {

using luisa::render::Interaction;
using luisa::render::Pipeline;
using luisa::render::SampledWavelengths;
using luisa::render::Surface;

/* Closures only STORE the pipeline reference (surface.h:83-85); nothing on the evaluate / sample path reads it. */
inline const Pipeline &unused_pipeline() {
    alignas(64) static std::byte storage[16384]{};
    return *reinterpret_cast<const Pipeline *>(storage);
}

/* the interaction Geometry::interaction builds for a hit (interaction.h:86-91): geometric normal, shading frame from
 * the interpolated normal and dpdu */
inline Interaction make_interaction(Expr<float3> ng, Expr<float3> ns, Expr<float3> tangent) {
    return Interaction{luisa::render::Shape::Handle{}, 0u, 0u, 1.f, make_float3(0.f), ng, make_float2(0.f),
                       make_float3(0.f), ns, tangent, false};
}

template<typename Closure>
auto closure_evaluate(Closure &closure, Expr<float3> wo, Expr<float3> wi) {
    closure.pre_eval();
    auto e = closure.evaluate(wo, wi);
    auto r = pack(e.f[0u], e.f[1u], e.f[2u], e.pdf);
    closure.post_eval();
    return r;
}

template<typename Closure>
auto closure_sample(Closure &closure, Expr<float3> wo, Expr<float> u_lobe, Expr<float2> u) {
    closure.pre_eval();
    auto s = closure.sample(wo, u_lobe, u);
    auto r = pack(s.wi.x, s.wi.y, s.wi.z, s.eval.f[0u], s.eval.f[1u], s.eval.f[2u], s.eval.pdf, bits(s.event));
    closure.post_eval();
    return r;
}

}// namespace refpins

/* every plugin source ends with LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(cls), which defines the module's extern "C"
 * create/destroy; several plugins are linked into ONE library here, so the pin translation units neutralise it */
#define REFPIN_INCLUDE_PLUGIN_PROLOGUE
