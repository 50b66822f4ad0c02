// Importance map of an image-textured Spherical environment, built on the host the way Spherical::build does it on the device
// (src/environments/spherical.cpp:140-236): a 2048 x 1024 map of Gaussian-filtered luminance * sin(theta), optional MIS
// compensation, one alias table per row + the marginal table over the rows, and the matching pdf table.
#pragma once
#include <cstdint>
#include <vector>

#include "../../../include/lrk.h"

namespace lrh {

constexpr uint32_t kEnvMapWidth = 2048u, kEnvMapHeight = 1024u;// Spherical::sample_map_size, spherical.cpp:21

// `texture` / `texels`: the emission texture exactly as the device will sample it (record + its RGBA float texels).
void build_environment_map(const lrk_texture &texture, const float *texels, bool compensate_mis,
                           std::vector<lrk_alias_entry> &alias, std::vector<float> &pdf);

// ImageTextureInstance::evaluate on the host (src/textures/image.cpp:132-166 with the software sampler of
// cpu_texture.h:418-493) — the same arithmetic as the device's texture_evaluate.
void host_texture_evaluate(const lrk_texture &texture, const float *texels, float u, float v, float out[4]);

}// namespace lrh
