// pins.h — registration helpers shared by the pin translation units (TEST INFRASTRUCTURE; see oracle/ref/README.md).
#pragma once

#include <functional>
#include <map>
#include <string>

#include <luisa/dsl/sugar.h>
#include <luisa/dsl/syntax.h>

#include <util/spec.h>

namespace refpins {

using namespace luisa;
using namespace luisa::compute;
using luisa::render::SampledSpectrum;

struct Pin {
    luisa::shared_ptr<const luisa::compute::detail::FunctionBuilder> builder;
};

std::map<std::string, std::function<Pin()>> &factories();

/* registers `def` (a lambda over DSL variables) under `name`; the callable is recorded on first use */
template<typename F>
void add(const char *name, F &&def) {
    factories()[name] = [def = std::forward<F>(def)]() mutable {
        Callable c = def;
        return Pin{c.function_builder()};
    };
}

template<typename... T>
auto pack(T &&...xs) {
    ArrayFloat<sizeof...(T)> a;
    auto i = 0u;
    ((a[i++] = std::forward<T>(xs)), ...);
    return a;
}
inline Float bits(Expr<uint> x) { return as<float>(x); }
inline Float flag(Expr<bool> x) { return ite(x, 1.f, 0.f); }
inline SampledSpectrum spec3(Expr<float3> v) {
    SampledSpectrum s{3u};
    s[0u] = v.x;
    s[1u] = v.y;
    s[2u] = v.z;
    return s;
}
inline auto pack_spec(const SampledSpectrum &s) { return pack(s[0u], s[1u], s[2u]); }

/* one registrar object per translation unit: its constructor runs the TU's add() calls at load time */
struct Registrar {
    explicit Registrar(void (*f)()) { f(); }
};

}// namespace refpins
