// shim.cpp — the few symbols the reference objects expect from parts of the reference that are NOT built here
// (TEST INFRASTRUCTURE; see oracle/ref/README.md).  Nothing here is on a numerical path.
#include <cstddef>
#include <cstdlib>
#include <new>

// EASTL expects the application to provide these two allocation operators (EASTL/allocator.h; the reference provides them in
// its EASTL fork's allocator sources when built through CMake with EASTL_USER_DEFINED_ALLOCATOR).
void *operator new[](size_t size, const char *, int, unsigned, const char *, int) { return ::operator new[](size); }
void *operator new[](size_t size, size_t alignment, size_t, const char *, int, unsigned, const char *, int) {
    return ::operator new[](size, std::align_val_t{alignment});
}

// src/util/imageio.cpp (EXR / HDR output through tinyexr / stb) is not built; the film the reference hands to save_image
// (src/base/integrator.cpp:34-49) is written as raw floats instead: "<path>.f32" = one text line "width height channels\n"
// followed by width * height * channels little-endian float32, row 0 first.
#include <cstdio>
#include <filesystem>
#include <luisa/core/basic_types.h>
namespace luisa::render {
void save_image(std::filesystem::path path, const float *pixels, luisa::uint2 resolution, uint components) {
    path += ".f32";
    auto file = std::fopen(path.string().c_str(), "wb");
    if (file == nullptr) { std::abort(); }
    std::fprintf(file, "%u %u %u\n", resolution.x, resolution.y, components);
    std::fwrite(pixels, sizeof(float), static_cast<size_t>(resolution.x) * resolution.y * components, file);
    std::fclose(file);
}
}// namespace luisa::render
