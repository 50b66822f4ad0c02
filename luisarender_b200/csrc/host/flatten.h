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

}// namespace lrh
