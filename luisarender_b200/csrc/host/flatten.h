// Flattening of the host scene graph into the POD arrays of include/lrk.h.  This is the host half of
// Pipeline::create (src/base/pipeline.cpp:44-99): Geometry::build / _process_shape
// (src/base/geometry.cpp:12-163), register_surface/light (pipeline.cpp:20-42), the uniform light
// sampler's handle list (src/lightsamplers/uniform.cpp:32-48), camera + filter tables
// (src/base/filter.cpp:24-48) — plus the BVH build the reference leaves to OptiX/Embree.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "scene.h"

namespace lrh {

struct FlatCamera {
    lrk_camera camera;
    lrk_film film;
    std::filesystem::path file;
    // what Sampler::Instance::reset(resolution, spp) prepares for this camera (pointers of `sampler` are filled by desc())
    lrk_sampler sampler{};
    std::vector<float> pmj_pixel_samples;       // PMJ02BN: float2[tile * tile * spp]
    std::vector<uint64_t> vdc, vdc_inv;         // SOBOL: the rows of the van-der-Corput matrices for this resolution
    std::vector<uint32_t> zsobol_hash;          // ZSOBOL: uint2[1024]
};

struct FlatScene {
    std::vector<lrk_vertex> vertices;
    std::vector<lrk_triangle> triangles;
    std::vector<lrk_alias_entry> alias;
    std::vector<float> pdf;
    std::vector<lrk_mesh> meshes;
    std::vector<lrk_instance> instances;
    std::vector<lrk_bvh_node> bvh_nodes;
    std::vector<float> tri_verts;
    std::vector<lrk_surface> surfaces;
    std::vector<lrk_texture> textures;// image textures referenced by surfaces (lrk_surface::tex)
    std::vector<float> texels;        // RGBA float texels of all of them
    std::vector<lrk_light> lights;
    std::vector<lrk_light_handle> light_handles;// all instanced lights; the desc exposes the first lights.size()
    std::vector<FlatCamera> cameras;
    lrk_integrator integrator{};
    lrk_medium environment_medium{};
    std::vector<lrk_medium> media;             // indexed by the handles' medium tags; the environment medium is registered last
    uint32_t environment_medium_tag{LRK_MEDIUM_INVALID_TAG};
    lrk_environment environment{};           // pointers are filled by desc()
    std::vector<lrk_alias_entry> env_alias;  // importance map of an image-textured environment (spherical.cpp:140-236)
    std::vector<float> env_pdf;
    uint32_t tlas_root{0};
    float world_min[3]{}, world_max[3]{};
    uint64_t total_instanced_triangles{0};
    double bvh_build_ms{0.0};

    // a view (no ownership) for camera `index`
    lrk_scene_desc desc(uint32_t camera_index = 0) const;
};

std::unique_ptr<FlatScene> flatten_scene(const Scene &scene);

// the tables of the quasi-Monte-Carlo samplers (luisarender_b200/data/sampler_tables.bin), loaded on first use
struct SamplerTables {
    std::vector<uint32_t> sobol_matrices;// [1024][52]
    std::vector<uint64_t> vdc, vdc_inv;  // [25][52], [26][52]
    std::vector<uint32_t> pmj_samples;   // [5][65536][2]
    std::vector<uint16_t> blue_noise;    // [48][128][128]
};
const SamplerTables &sampler_tables();

// the Metal surface's named metals (luisarender_b200/data/metal_ior.bin, tools/extract_metal_ior.py): measured complex refractive
// index at 360 .. 830 nm in 5 nm steps; `name` as the reference calls its tables (Ag Al Au Cu CuZn Fe Ti V VN Li Cr)
struct MetalIor {
    std::vector<float> n, k;// 95 samples each
};
const MetalIor &metal_ior_table(const std::string &name);
std::filesystem::path data_directory();// <library dir>/../data or $LRH_DATA_DIR

}// namespace lrh
