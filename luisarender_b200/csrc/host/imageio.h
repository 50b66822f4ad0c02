#pragma once
#include <cstdint>
#include <filesystem>
#include <stdexcept>
#include <vector>

namespace lrh {
// writes W*H RGBA float pixels; returns the path actually written (extension may fall back to .exr)
std::filesystem::path save_image(std::filesystem::path path, const float *rgba, uint32_t width, uint32_t height);

// A decoded image file: RGBA float texels, row 0 = top row; `channels` = the storage channel count the reference would
// pick for the file (1, 2 or 4 — LoadedImage::parse_storage, src/util/imageio.cpp:347-405).
struct LoadedImage {
    uint32_t width{0}, height{0}, channels{0};
    std::vector<float> rgba;
};
LoadedImage load_image(const std::filesystem::path &path);// throws std::runtime_error
// jpegload.cpp: nc = 1 (grey) or 3 (RGB) interleaved bytes, row 0 on top; throws std::runtime_error
void decode_jpeg(const std::filesystem::path &path, const std::vector<uint8_t> &data, uint32_t &w, uint32_t &h, uint32_t &nc,
                 std::vector<uint8_t> &pixels);
}// namespace lrh
