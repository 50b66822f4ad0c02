#pragma once
#include <cstdint>
#include <filesystem>
#include <stdexcept>

namespace lrh {
// writes W*H RGBA float pixels; returns the path actually written (extension may fall back to .exr)
std::filesystem::path save_image(std::filesystem::path path, const float *rgba, uint32_t width, uint32_t height);
}// namespace lrh
