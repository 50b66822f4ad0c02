/*
 * interp.h — a small CPU interpreter for LuisaCompute ASTs (TEST INFRASTRUCTURE, see oracle/ref/README.md).
 *
 * LuisaRender's numerical code is written in the LuisaCompute DSL: calling e.g. `fresnel_dielectric(...)` does not
 * compute anything, it RECORDS an AST (luisa::compute::Function) that a backend (CUDA/NVRTC+OptiX, Rust/LLVM+Embree)
 * would compile.  Neither backend can be built in this environment, but the recording layer itself
 * (src/compute/src/{core,ast,dsl}) and LuisaRender's src/util are plain C++ and compile from the sources where they
 * lie under /root/reference (oracle/ref/Makefile).  This interpreter executes the recorded AST on the host, so the
 * reference's OWN functions run here, unmodified, and their outputs pin the oracle (tests/test_ref_pins.py).
 *
 * Builtin semantics follow the reference's CUDA backend header
 * src/compute/src/backends/cuda/cuda_builtin/cuda_device_math.h (lerp :3353, fract :3368, clamp = min(max(v,lo),hi)
 * :3301, normalize = v * rsqrt(dot(v,v)) :3509, reflect :3680-3682, pow with integral exponents = repeated
 * multiplication :24-36, step/smoothstep :3361-3365); transcendental functions are glibc's, rsqrt(x) = 1/sqrt(x).
 */
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include <luisa/ast/function.h>

namespace refinterp {

struct BufferArg {
    std::byte *data{nullptr};
    size_t count{0u};
};

/* one argument of the entry function: either a value (copied in; written back when the parameter is a reference)
 * or a buffer */
struct Arg {
    std::vector<std::byte> bytes;
    BufferArg buffer;
};

/* Executes `f` (a callable) and returns the bytes of its return value (empty for void).  Throws std::runtime_error on
 * AST constructs the interpreter does not implement. */
std::vector<std::byte> call(luisa::compute::Function f, std::vector<Arg> &args);

}// namespace refinterp
