/*
 * interp.h — a small CPU interpreter for LuisaCompute ASTs (TEST INFRASTRUCTURE, see oracle/ref/README.md).
 *
 * LuisaRender's numerical code is written in the LuisaCompute DSL: calling e.g. `fresnel_dielectric(...)` does not
 * compute anything, it RECORDS an AST (luisa::compute::Function) that a backend (CUDA/NVRTC+OptiX, Rust/LLVM+Embree)
 * would compile.  Neither backend can be built in this environment, but the recording layer itself
 * (src/compute/src/{core,ast,dsl}) and LuisaRender's sources are plain C++ and compile from the sources where they
 * lie under /root/reference (oracle/ref/Makefile).  This interpreter executes the recorded AST on the host, so the
 * reference's OWN functions — and, through refdevice.cpp, its whole renderer — run here unmodified, and their outputs
 * pin the oracle (tests/test_ref_pins.py, tests/test_ref_render.py).
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
    size_t size_bytes{0u};
};

/* LuisaCompute's Ray (include/luisa/runtime/rtx/ray.h) and SurfaceHit (rtx/hit.h:30-35) */
struct RayData {
    float origin[3];
    float t_min;
    float direction[3];
    float t_max;
};
struct HitData {
    uint32_t inst;
    uint32_t prim;
    float bary[2];
    float committed_ray_t;
    uint32_t pad;
};

/* what a kernel can reach on the "device": implemented by refdevice.cpp */
struct DeviceResources {
    virtual ~DeviceResources() = default;
    [[nodiscard]] virtual BufferArg bindless_buffer(uint64_t array, uint32_t slot) = 0;
    [[nodiscard]] virtual HitData trace_closest(uint64_t accel, const RayData &ray, uint32_t mask) = 0;
    [[nodiscard]] virtual bool trace_any(uint64_t accel, const RayData &ray, uint32_t mask) = 0;
    virtual void instance_transform(uint64_t accel, uint32_t index, float out_column_major[16]) = 0;
    /* every triangle the ray crosses inside (t_min, t_max), in traversal order, with its instance's opaque flag: the
     * candidates of a ray query (include/luisa/dsl/rtx/ray_query.h) */
    struct Candidate {
        HitData hit;
        bool opaque;
    };
    [[nodiscard]] virtual std::vector<Candidate> candidates(uint64_t accel, const RayData &ray, uint32_t mask) = 0;
    /* bindless 2D texture `slot` of `array`, sampled at uv with the slot's sampler (level 0), as float4 */
    virtual void bindless_tex2d_sample(uint64_t array, uint32_t slot, float u, float v, float out[4]) = 0;
    virtual void bindless_tex2d_size(uint64_t array, uint32_t slot, uint32_t out[2]) = 0;
    virtual void bindless_tex2d_read(uint64_t array, uint32_t slot, uint32_t x, uint32_t y, float out[4]) = 0;
    virtual void bindless_tex3d_read(uint64_t array, uint32_t slot, uint32_t x, uint32_t y, uint32_t z, float out[4]) = 0;
};

/* one argument of the entry function */
struct Arg {
    enum struct Kind { VALUE, BUFFER, BINDLESS_ARRAY, ACCEL } kind{Kind::VALUE};
    std::vector<std::byte> bytes;/* VALUE: copied in; written back when the parameter is a reference */
    BufferArg buffer;            /* BUFFER */
    uint64_t handle{0u};         /* BINDLESS_ARRAY / ACCEL */
};

/* Executes callable `f` and returns the bytes of its return value (empty for void).  Throws std::runtime_error on
 * AST constructs the interpreter does not implement. */
std::vector<std::byte> call(luisa::compute::Function f, std::vector<Arg> &args, DeviceResources *resources = nullptr);

/* Executes kernel `f` for every dispatch id in [0, size) on `threads` host threads. */
void launch(luisa::compute::Function f, const std::vector<Arg> &args, const uint32_t size[3], DeviceResources *resources,
            unsigned threads);

}// namespace refinterp
