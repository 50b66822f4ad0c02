// Triangle-mesh file readers for the `Mesh` shape plugin (reference: MeshLoader::load, src/shapes/mesh.cpp:29-146,
// which delegates to assimp — not available here).  Wavefront OBJ and Stanford PLY are read directly; the
// post-processing the reference asks assimp for is reproduced where it changes what the renderer sees:
//   * polygons are triangulated (fan), identical vertices are joined (aiProcess_JoinIdenticalVertices)
//   * v is flipped to 1 - v unless `flip_uv` is set (the reference passes aiProcess_FlipUVs when flip_uv is FALSE, :62)
//   * missing normals are generated as smooth normals with a 45 degree crease angle (aiProcess_GenSmoothNormals +
//     AI_CONFIG_PP_GSN_MAX_SMOOTHING_ANGLE = 45, :47,70) unless `drop_normal`; given normals are normalised (:113)
#pragma once
#include <cstdint>
#include <filesystem>
#include <vector>

#include "../../../include/lrk.h"

namespace lrh {

struct MeshData {
    std::vector<lrk_vertex> vertices;
    std::vector<lrk_triangle> triangles;
    uint32_t properties{0};// LRK_SHAPE_HAS_VERTEX_NORMAL | LRK_SHAPE_HAS_VERTEX_UV
};

MeshData load_mesh(const std::filesystem::path &path, bool flip_uv, bool drop_normal, bool drop_uv);// throws std::runtime_error

}// namespace lrh
