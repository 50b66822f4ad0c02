/*
 * lrh.h — C-ABI of the host front-end (libluisa_render_host.so): scene description parsing, scene
 * graph construction, flattening to the POD scene of lrk.h (including the BVH2 build) and film output.
 * It replaces, for the node types in scope, the reference's
 *   SceneParser::parse            src/sdl/scene_parser.cpp:400-407
 *   Scene::create                 src/base/scene.cpp:201-233
 *   Pipeline::create (host part)  src/base/pipeline.cpp:44-99, src/base/geometry.cpp:12-163
 *   save_image                    src/util/imageio.cpp:694-726
 * Everything is extern "C"; errors return a negative value and a message through lrh_last_error()
 * (thread-local).  This library contains no device code and no radiance computation.
 */
#ifndef LRH_H
#define LRH_H

#include <stdint.h>

#include "lrk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lrh_scene lrh_scene;

typedef struct lrh_scene_info {
    uint64_t unique_triangles;
    uint64_t instanced_triangles;
    uint64_t vertices;
    uint64_t bvh_nodes;
    uint32_t meshes;
    uint32_t instances;
    uint32_t surfaces;
    uint32_t lights;
    uint32_t cameras;
    uint32_t reserved;
    double bvh_build_ms;
    float world_min[3];
    float world_max[3];
} lrh_scene_info;

const char *lrh_last_error(void);

/* Load a .luisa / .json scene file.  macro_keys/values are the CLI "-D key=value" definitions. */
int lrh_scene_load(const char *path, const char *const *macro_keys, const char *const *macro_values,
                   uint32_t macro_count, lrh_scene **out);
/* Same from a source string; base_dir resolves relative imports and output paths. */
int lrh_scene_load_source(const char *source, int is_json, const char *base_dir, const char *const *macro_keys,
                          const char *const *macro_values, uint32_t macro_count, lrh_scene **out);
void lrh_scene_destroy(lrh_scene *scene);

int lrh_scene_get_info(const lrh_scene *scene, lrh_scene_info *info);
/* Fill `out` with a view of the flattened scene for camera `camera`; pointers stay valid until
 * lrh_scene_destroy. */
int lrh_scene_get_desc(const lrh_scene *scene, uint32_t camera, lrk_scene_desc *out);
/* Output image path of camera `camera` (the camera's `file` property, default <scene dir>/render.exr). */
const char *lrh_scene_camera_file(const lrh_scene *scene, uint32_t camera);

/* RGBA float image writer (.exr / .hdr / .pfm; other extensions fall back to .exr). */
int lrh_save_image(const char *path, const float *rgba, uint32_t width, uint32_t height);

/* Image file reader of the `Image` texture plugin (PNG / JPEG / BMP / TGA / PPM / PGM / PFM / HDR / EXR): replaces LoadedImage::load
 * (src/util/imageio.cpp:480-560).  Writes the decoded RGBA float texels (row 0 = top row) into `rgba` when it is non-null
 * and holds at least width * height * 4 floats; call once with rgba = NULL to query the size. */
int lrh_load_image(const char *path, uint32_t *width, uint32_t *height, uint32_t *channels, float *rgba, uint64_t rgba_capacity);

/* Registered node implementations, as "<tag>-<impl>" (the reference's plugin name minus "luisa-render-"). */
uint32_t lrh_plugin_count(void);
const char *lrh_plugin_name(uint32_t index);

/* The alias-table builder used for mesh light sampling, filter importance sampling and the environment map: replaces
 * create_alias_table (src/util/sampling.cpp:38-87).  prob / alias / pdf are caller-allocated arrays of n elements. */
int lrh_create_alias_table(const float *values, uint32_t n, float *prob, uint32_t *alias, float *pdf);

#ifdef __cplusplus
}
#endif
#endif /* LRH_H */
