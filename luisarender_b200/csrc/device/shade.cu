// shade.cu - the closure kernels of the device library, as their own translation unit, compiled TWICE.
//
// lrk.cu (ray generation, BVH traversal, hit classification, film) is compiled with IEEE arithmetic and no FMA contraction: its
// results are compared with the oracle bit for bit.  The kernels in THIS file evaluate closures, lights and media - thousands of
// fp32 operations per path vertex whose exact rounding no interface depends on.
//   -DLRK_SHADE_VARIANT=fast   --use_fast_math: the arithmetic the reference's own CUDA backend uses for every kernel (it passes
//                              -use_fast_math to NVRTC by default: src/compute/src/backends/cuda/cuda_device.cpp:697-698,
//                              ShaderOption::enable_fast_math{true} in src/compute/include/luisa/runtime/rhi/resource.h:109).
//                              The default for the emitter / Matte / Disney buckets and the volume integrator's steps.  Measured on
//                              the 1.39 M-triangle scene: shade 18.3 -> 11.0 ms per 64-spp pass (profiles/r02k_shade_arithmetic.jsonl).
//   -DLRK_SHADE_VARIANT=strict -fmad=false, IEEE div / sqrt: bit-compatible with lrk.cu and the oracle.  Always used for the
//                              Mirror / Glass / Plastic / Metal / Mix buckets (near-specular lobes amplify a 1e-6 error in the half
//                              vector into percents of the lobe value), for the Layered and thin Disney buckets, and for everything with lrk_set_option("strict_math", 1) -
//                              the configuration the tight parity tests run (films equal the oracle's to rel-L2 ~ 1e-7).
// The closures' source is the same in both, and is checked against the reference's closures bit for bit when compiled for the host
// (tests/test_device_closures_on_host.py).  build.py compiles the two objects; cross products and the uv determinant are written so
// that contraction cannot turn their exact zeros into rounding residue (vecmath.cuh: mul_exact).
#include "shade_launch.h"
#include "shade_kernels.cuh"

namespace lrk {
namespace LRK_SHADE_VARIANT {

namespace {

template<typename Kernel>
int grid_of(Kernel kernel, int block, int sm_count) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
    return (per_sm > 0 ? per_sm : 1) * sm_count;
}

// calls f(kernel) with the shade kernel instantiation of (kind, textured)
template<typename F>
void with_shade_kernel(uint32_t kind, bool textured, F &&f) {
    switch (kind) {
        case 0u: textured ? f(shade_kernel<0u, true>) : f(shade_kernel<0u, false>); break;// emitter-only hits: a light's image emission
        case 1u: textured ? f(shade_kernel<1u, true>) : f(shade_kernel<1u, false>); break;
        case 2u: textured ? f(shade_kernel<2u, true>) : f(shade_kernel<2u, false>); break;
        case 3u: textured ? f(shade_kernel<3u, true>) : f(shade_kernel<3u, false>); break;
        case 4u: textured ? f(shade_kernel<4u, true>) : f(shade_kernel<4u, false>); break;
        case 5u: textured ? f(shade_kernel<5u, true>) : f(shade_kernel<5u, false>); break;
        case 6u: textured ? f(shade_kernel<6u, true>) : f(shade_kernel<6u, false>); break;
        case 7u: textured ? f(shade_kernel<7u, true>) : f(shade_kernel<7u, false>); break;
        case 8u: textured ? f(shade_kernel<8u, true>) : f(shade_kernel<8u, false>); break;
        case 9u: textured ? f(shade_kernel<9u, true>) : f(shade_kernel<9u, false>); break;
        default: textured ? f(shade_kernel<10u, true>) : f(shade_kernel<10u, false>); break;
    }
}

template<typename F>
void with_volume_surface_kernel(uint32_t kind, bool textured, F &&f) {
    switch (kind) {
        case 0u: textured ? f(volume_surface_kernel<0u, true>) : f(volume_surface_kernel<0u, false>); break;
        case 1u: textured ? f(volume_surface_kernel<1u, true>) : f(volume_surface_kernel<1u, false>); break;
        default: textured ? f(volume_surface_kernel<2u, true>) : f(volume_surface_kernel<2u, false>); break;
    }
}

}// namespace

int shade_grid(uint32_t kind, int sm_count) {
    int g = 1;
    with_shade_kernel(kind, false, [&](auto kernel) { g = grid_of(kernel, kShadeBlock, sm_count); });
    return g;
}

int volume_medium_grid(int sm_count) { return grid_of(volume_medium_kernel, kBlock, sm_count); }

int volume_surface_grid(uint32_t kind, int sm_count) {
    int g = 1;
    with_volume_surface_kernel(kind, false, [&](auto kernel) { g = grid_of(kernel, kBlock, sm_count); });
    return g;
}

void launch_shade(uint32_t kind, bool textured, int blocks, cudaStream_t stream, const DeviceScene &sc, const PathBuffers &pb, uint32_t depth) {
    with_shade_kernel(kind, textured, [&](auto kernel) { kernel<<<blocks, kShadeBlock, 0, stream>>>(sc, pb, depth); });
}

void launch_volume_medium(int blocks, cudaStream_t stream, const DeviceScene &sc, const PathBuffers &pb, uint32_t depth) {
    volume_medium_kernel<<<blocks, kBlock, 0, stream>>>(sc, pb, depth);
}

void launch_volume_surface(uint32_t kind, bool textured, int blocks, cudaStream_t stream, const DeviceScene &sc, const PathBuffers &pb, uint32_t depth) {
    with_volume_surface_kernel(kind, textured, [&](auto kernel) { kernel<<<blocks, kBlock, 0, stream>>>(sc, pb, depth); });
}

}// namespace LRK_SHADE_VARIANT
}// namespace lrk
