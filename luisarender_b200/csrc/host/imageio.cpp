// Film output: RGBA float images as OpenEXR (uncompressed scanlines, 32-bit float), Radiance HDR
// (flat RGBE) or PFM.  Same extension policy as save_image (src/util/imageio.cpp:694-726): anything
// that is not .exr / .hdr (/.pfm, our addition for tests) falls back to .exr.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "imageio.h"

namespace lrh {

namespace {

void put(std::vector<char> &b, const void *p, size_t n) {
    auto c = static_cast<const char *>(p);
    b.insert(b.end(), c, c + n);
}
void put_str(std::vector<char> &b, const char *s) { put(b, s, std::strlen(s) + 1); }
void put_i32(std::vector<char> &b, int32_t v) { put(b, &v, 4); }
void put_f32(std::vector<char> &b, float v) { put(b, &v, 4); }
void attr(std::vector<char> &b, const char *name, const char *type, const std::vector<char> &value) {
    put_str(b, name);
    put_str(b, type);
    put_i32(b, static_cast<int32_t>(value.size()));
    put(b, value.data(), value.size());
}

bool write_exr(const std::filesystem::path &path, const float *rgba, uint32_t w, uint32_t h) {
    std::vector<char> hd;
    const uint32_t magic = 20000630u, version = 2u;
    put(hd, &magic, 4);
    put(hd, &version, 4);
    {// channels, alphabetical: A B G R, pixel type FLOAT = 2
        std::vector<char> v;
        for (auto name : {"A", "B", "G", "R"}) {
            put_str(v, name);
            put_i32(v, 2);
            const char plinear[4]{0, 0, 0, 0};
            put(v, plinear, 4);
            put_i32(v, 1);
            put_i32(v, 1);
        }
        v.push_back(0);
        attr(hd, "channels", "chlist", v);
    }
    { std::vector<char> v{0}; attr(hd, "compression", "compression", v); }
    {
        std::vector<char> v;
        put_i32(v, 0); put_i32(v, 0); put_i32(v, static_cast<int32_t>(w) - 1); put_i32(v, static_cast<int32_t>(h) - 1);
        attr(hd, "dataWindow", "box2i", v);
        attr(hd, "displayWindow", "box2i", v);
    }
    { std::vector<char> v{0}; attr(hd, "lineOrder", "lineOrder", v); }
    { std::vector<char> v; put_f32(v, 1.f); attr(hd, "pixelAspectRatio", "float", v); }
    { std::vector<char> v; put_f32(v, 0.f); put_f32(v, 0.f); attr(hd, "screenWindowCenter", "v2f", v); }
    { std::vector<char> v; put_f32(v, 1.f); attr(hd, "screenWindowWidth", "float", v); }
    hd.push_back(0);

    std::ofstream f{path, std::ios::binary};
    if (!f) return false;
    f.write(hd.data(), static_cast<std::streamsize>(hd.size()));
    const uint64_t row_bytes = static_cast<uint64_t>(w) * 16u;
    uint64_t offset = hd.size() + static_cast<uint64_t>(h) * 8u;
    for (uint32_t y = 0; y < h; y++) {
        f.write(reinterpret_cast<const char *>(&offset), 8);
        offset += 8u + row_bytes;
    }
    std::vector<float> row(static_cast<size_t>(w) * 4u);
    for (uint32_t y = 0; y < h; y++) {
        auto yi = static_cast<int32_t>(y);
        auto sz = static_cast<int32_t>(row_bytes);
        f.write(reinterpret_cast<const char *>(&yi), 4);
        f.write(reinterpret_cast<const char *>(&sz), 4);
        const float *src = rgba + static_cast<size_t>(y) * w * 4u;
        const int order[4]{3, 2, 1, 0};// A B G R
        for (int c = 0; c < 4; c++)
            for (uint32_t x = 0; x < w; x++) row[static_cast<size_t>(c) * w + x] = src[x * 4u + order[c]];
        f.write(reinterpret_cast<const char *>(row.data()), static_cast<std::streamsize>(row_bytes));
    }
    return static_cast<bool>(f);
}

bool write_pfm(const std::filesystem::path &path, const float *rgba, uint32_t w, uint32_t h) {
    std::ofstream f{path, std::ios::binary};
    if (!f) return false;
    f << "PF\n" << w << " " << h << "\n-1.0\n";
    std::vector<float> row(static_cast<size_t>(w) * 3u);
    for (uint32_t y = 0; y < h; y++) {// PFM stores bottom row first
        const float *src = rgba + static_cast<size_t>(h - 1u - y) * w * 4u;
        for (uint32_t x = 0; x < w; x++)
            for (int c = 0; c < 3; c++) row[x * 3u + c] = src[x * 4u + c];
        f.write(reinterpret_cast<const char *>(row.data()), static_cast<std::streamsize>(row.size() * 4u));
    }
    return static_cast<bool>(f);
}

bool write_hdr(const std::filesystem::path &path, const float *rgba, uint32_t w, uint32_t h) {
    std::ofstream f{path, std::ios::binary};
    if (!f) return false;
    f << "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y " << h << " +X " << w << "\n";
    std::vector<unsigned char> row(static_cast<size_t>(w) * 4u);
    for (uint32_t y = 0; y < h; y++) {
        const float *src = rgba + static_cast<size_t>(y) * w * 4u;
        for (uint32_t x = 0; x < w; x++) {
            float r = src[x * 4u], g = src[x * 4u + 1], b = src[x * 4u + 2];
            float m = std::max(r, std::max(g, b));
            unsigned char *o = &row[x * 4u];
            if (!(m > 1e-32f)) {
                o[0] = o[1] = o[2] = o[3] = 0;
            } else {
                int e;
                float s = std::frexp(m, &e) * 256.0f / m;
                o[0] = static_cast<unsigned char>(std::max(r, 0.f) * s);
                o[1] = static_cast<unsigned char>(std::max(g, 0.f) * s);
                o[2] = static_cast<unsigned char>(std::max(b, 0.f) * s);
                o[3] = static_cast<unsigned char>(e + 128);
            }
        }
        f.write(reinterpret_cast<const char *>(row.data()), static_cast<std::streamsize>(row.size()));
    }
    return static_cast<bool>(f);
}

}// namespace

std::filesystem::path save_image(std::filesystem::path path, const float *rgba, uint32_t width, uint32_t height) {
    auto ext = path.extension().string();
    for (auto &c : ext) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    if (ext != ".exr" && ext != ".hdr" && ext != ".pfm") {
        path.replace_extension(".exr");
        ext = ".exr";
    }
    if (auto dir = path.parent_path(); !dir.empty()) {
        std::error_code ec;
        std::filesystem::create_directories(dir, ec);
    }
    bool ok = ext == ".exr" ? write_exr(path, rgba, width, height) :
              ext == ".hdr" ? write_hdr(path, rgba, width, height) :
                              write_pfm(path, rgba, width, height);
    if (!ok) throw std::runtime_error("Failed to save film to '" + path.string() + "'.");
    return path;
}

}// namespace lrh
