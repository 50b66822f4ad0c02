#include "envmap.h"

#include <algorithm>
#include <cmath>
#include <thread>

#include "scene.h"

namespace lrh {

namespace {

constexpr float kOneMinusEpsilon = 0x1.fffffep-1f;
constexpr float kPiF = 3.14159265358979323846264338327950288f;

inline float fractf(float x) { return x - std::floor(x); }
inline float coord_point(uint32_t address, float uv, float s) {// cpu_texture.h:418-433
    switch (address) {
        case LRK_TEX_ADDRESS_EDGE: return std::fmin(std::fmax(uv, 0.0f), kOneMinusEpsilon) * s;
        case LRK_TEX_ADDRESS_REPEAT: return fractf(uv) * s;
        case LRK_TEX_ADDRESS_MIRROR: {
            uv = std::fmod(std::fabs(uv), 2.0f);
            uv = uv < 1.f ? uv : 2.f - uv;
            return std::fmin(uv, kOneMinusEpsilon) * s;
        }
        default: return (uv < 0.f || uv >= 1.f) ? 65536.f : uv * s;
    }
}
inline void read_texel(const lrk_texture &t, const float *texels, uint32_t x, uint32_t y, float out[4]) {
    if (!(x < t.width && y < t.height)) {
        out[0] = out[1] = out[2] = out[3] = 0.f;
        return;
    }
    const float *p = texels + 4u * (static_cast<size_t>(y) * t.width + x);
    out[0] = p[0], out[1] = p[1], out[2] = p[2], out[3] = p[3];
}
inline float lerpf(float a, float b, float t) { return t * (b - a) + a; }
inline float decode(const lrk_texture &t, float x) {// image.cpp:143-158
    if (t.encoding == LRK_TEX_ENCODING_SRGB) {
        float lin = x <= 0.04045f ? x * (1.0f / 12.92f) : std::pow((x + 0.055f) * (1.0f / 1.055f), 2.4f);
        return t.scale * lin;
    }
    if (t.encoding == LRK_TEX_ENCODING_GAMMA) return t.scale * std::pow(x, t.gamma);
    return t.scale * x;
}

}// namespace

void host_texture_evaluate(const lrk_texture &t, const float *texels, float u_in, float v_in, float out[4]) {
    const float u = u_in * t.uv_scale[0] + t.uv_offset[0], v = v_in * t.uv_scale[1] + t.uv_offset[1];
    const float sx = static_cast<float>(t.width), sy = static_cast<float>(t.height);
    float s[4];
    if (t.filter == LRK_TEX_FILTER_POINT) {
        read_texel(t, texels, static_cast<uint32_t>(coord_point(t.address, u, sx)), static_cast<uint32_t>(coord_point(t.address, v, sy)), s);
    } else {
        const float inv_sx = 1.f / sx, inv_sy = 1.f / sy;
        float ax = coord_point(t.address, u - .5f * inv_sx, sx), bx = coord_point(t.address, u + .5f * inv_sx, sx);
        float ay = coord_point(t.address, v - .5f * inv_sy, sy), by = coord_point(t.address, v + .5f * inv_sy, sy);
        float x_min = std::fmin(ax, bx), x_max = std::fmax(ax, bx), y_min = std::fmin(ay, by), y_max = std::fmax(ay, by);
        float tx = fractf(x_max), ty = fractf(y_max);
        uint32_t x0 = static_cast<uint32_t>(x_min), y0 = static_cast<uint32_t>(y_min), x1 = static_cast<uint32_t>(x_max), y1 = static_cast<uint32_t>(y_max);
        float v00[4], v01[4], v10[4], v11[4];
        read_texel(t, texels, x0, y0, v00);
        read_texel(t, texels, x1, y0, v01);
        read_texel(t, texels, x0, y1, v10);
        read_texel(t, texels, x1, y1, v11);
        for (int c = 0; c < 4; c++) s[c] = lerpf(lerpf(v00[c], v01[c], tx), lerpf(v10[c], v11[c], tx), ty);
    }
    for (int c = 0; c < 4; c++) out[c] = decode(t, s[c]);
}

void build_environment_map(const lrk_texture &texture, const float *texels, bool compensate_mis,
                           std::vector<lrk_alias_entry> &alias, std::vector<float> &pdf) {
    constexpr uint32_t W = kEnvMapWidth, H = kEnvMapHeight;
    constexpr uint32_t pixel_count = W * H;
    std::vector<float> scale_map(pixel_count);
    // generate_weight_map_kernel, spherical.cpp:149-177: 17 x 17 taps at 1/8-pixel steps, Gaussian weights exp(-4 |offset|^2)
    constexpr float filter_step = .125f;
    constexpr int n = 8;// ceil(filter_radius / filter_step)
    float weights[2 * n + 1][2 * n + 1];
    for (int dy = -n; dy <= n; dy++)
        for (int dx = -n; dx <= n; dx++) {
            float ox = static_cast<float>(dx) * filter_step, oy = static_cast<float>(dy) * filter_step;
            weights[dy + n][dx + n] = std::exp(-4.f * (ox * ox + oy * oy));
        }
    auto rows = [&](uint32_t y_begin, uint32_t y_end) {
        for (uint32_t y = y_begin; y < y_end; y++) {
            for (uint32_t x = 0; x < W; x++) {
                float cx = static_cast<float>(x) + .5f, cy = static_cast<float>(y) + .5f;
                float sum_weight = 0.f, sum_scale = 0.f;
                for (int dy = -n; dy <= n; dy++) {
                    for (int dx = -n; dx <= n; dx++) {
                        float u = (cx + static_cast<float>(dx) * filter_step) / static_cast<float>(W);
                        float v = (cy + static_cast<float>(dy) * filter_step) / static_cast<float>(H);
                        float rgba[4];
                        host_texture_evaluate(texture, texels, u, v, rgba);
                        // evaluate_illuminant_spectrum(...).strength = srgb_to_cie_y(max(rgb, 0)) (texture.cpp:50-62, srgb.cpp:48-54)
                        float r = std::fmax(rgba[0], 0.f), g = std::fmax(rgba[1], 0.f), b = std::fmax(rgba[2], 0.f);
                        float scale = 0.212671f * r + 0.715160f * g + 0.072169f * b;
                        float sin_theta = std::sin(v * kPiF);
                        float weight = weights[dy + n][dx + n];
                        float value = weight * std::fmin(scale * sin_theta, 1e8f);
                        sum_weight += weight;
                        sum_scale += value;
                    }
                }
                scale_map[static_cast<size_t>(y) * W + x] = sum_scale / sum_weight;
            }
        }
    };
    {
        uint32_t threads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        uint32_t per = (H + threads - 1u) / threads;
        for (uint32_t t = 0; t < threads; t++) {
            uint32_t b = t * per, e = std::min(H, b + per);
            if (b < e) pool.emplace_back(rows, b, e);
        }
        for (auto &th : pool) th.join();
    }
    if (compensate_mis) {// :182-187
        double sum_scale = 0.;
        for (auto s : scale_map) sum_scale += s;
        auto average_scale = static_cast<float>(sum_scale / pixel_count);
        for (auto &s : scale_map) s = std::max(s - average_scale, 0.f);
    }
    std::vector<float> row_averages(H);
    pdf.assign(pixel_count, 0.f);
    alias.assign(static_cast<size_t>(H) + pixel_count, lrk_alias_entry{});
    std::vector<lrk_alias_entry> table;
    std::vector<float> pdf_table;
    for (uint32_t i = 0; i < H; i++) {// conditional tables, :191-205
        double sum = 0.;
        const float *values = scale_map.data() + static_cast<size_t>(i) * W;
        for (uint32_t x = 0; x < W; x++) sum += values[x];
        row_averages[i] = static_cast<float>(sum * (1.0 / W));
        create_alias_table(values, W, table, pdf_table);
        std::copy_n(pdf_table.data(), W, pdf.data() + static_cast<size_t>(i) * W);
        std::copy_n(table.data(), W, alias.data() + H + static_cast<size_t>(i) * W);
    }
    create_alias_table(row_averages.data(), H, table, pdf_table);// marginal table, :206-216
    std::copy_n(table.data(), H, alias.data());
    for (uint32_t y = 0; y < H; y++) {
        float scale = static_cast<float>(pdf_table[y] * pixel_count);
        for (uint32_t x = 0; x < W; x++) pdf[static_cast<size_t>(y) * W + x] *= scale;
    }
}

}// namespace lrh
