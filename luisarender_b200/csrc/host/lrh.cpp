#include "../../../include/lrh.h"

#include <cstring>
#include <string>

#include "flatten.h"
#include "imageio.h"

struct lrh_scene {
    std::unique_ptr<lrh::SceneDesc> desc;
    std::unique_ptr<lrh::Scene> scene;
    std::unique_ptr<lrh::FlatScene> flat;
    std::vector<std::string> camera_files;
};

namespace {
thread_local std::string g_error;
int fail(const std::string &msg) {
    g_error = msg;
    return -1;
}
lrh::MacroMap make_macros(const char *const *keys, const char *const *values, uint32_t n) {
    lrh::MacroMap m;
    for (uint32_t i = 0; i < n; i++)
        if (keys && values && keys[i] && values[i]) m[keys[i]] = values[i];
    return m;
}
int finish(std::unique_ptr<lrh::SceneDesc> desc, lrh_scene **out) {
    auto s = std::make_unique<lrh_scene>();
    s->desc = std::move(desc);
    s->scene = lrh::Scene::create(s->desc.get());
    s->flat = lrh::flatten_scene(*s->scene);
    for (auto &c : s->flat->cameras) s->camera_files.push_back(c.file.string());
    *out = s.release();
    return 0;
}
}// namespace

extern "C" {

const char *lrh_last_error(void) { return g_error.c_str(); }

int lrh_scene_load(const char *path, const char *const *macro_keys, const char *const *macro_values,
                   uint32_t macro_count, lrh_scene **out) {
    if (!path || !out) return fail("lrh_scene_load: null argument.");
    try {
        return finish(lrh::parse_scene_file(path, make_macros(macro_keys, macro_values, macro_count)), out);
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}

int lrh_scene_load_source(const char *source, int is_json, const char *base_dir, const char *const *macro_keys,
                          const char *const *macro_values, uint32_t macro_count, lrh_scene **out) {
    if (!source || !out) return fail("lrh_scene_load_source: null argument.");
    try {
        std::filesystem::path dir = base_dir ? std::filesystem::path{base_dir} : std::filesystem::current_path();
        return finish(lrh::parse_scene_source(source, dir, make_macros(macro_keys, macro_values, macro_count), is_json != 0), out);
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}

void lrh_scene_destroy(lrh_scene *scene) { delete scene; }

int lrh_scene_get_info(const lrh_scene *scene, lrh_scene_info *info) {
    if (!scene || !info) return fail("lrh_scene_get_info: null argument.");
    auto &f = *scene->flat;
    std::memset(info, 0, sizeof(*info));
    info->unique_triangles = f.triangles.size();
    info->instanced_triangles = f.total_instanced_triangles;
    info->vertices = f.vertices.size();
    info->bvh_nodes = f.bvh_nodes.size();
    info->meshes = static_cast<uint32_t>(f.meshes.size());
    info->instances = static_cast<uint32_t>(f.instances.size());
    info->surfaces = static_cast<uint32_t>(f.surfaces.size());
    info->lights = static_cast<uint32_t>(f.lights.size());
    info->cameras = static_cast<uint32_t>(f.cameras.size());
    info->bvh_build_ms = f.bvh_build_ms;
    for (int a = 0; a < 3; a++) {
        info->world_min[a] = f.world_min[a];
        info->world_max[a] = f.world_max[a];
    }
    return 0;
}

int lrh_scene_get_desc(const lrh_scene *scene, uint32_t camera, lrk_scene_desc *out) {
    if (!scene || !out) return fail("lrh_scene_get_desc: null argument.");
    try {
        *out = scene->flat->desc(camera);
        return 0;
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}

const char *lrh_scene_camera_file(const lrh_scene *scene, uint32_t camera) {
    if (!scene || camera >= scene->camera_files.size()) return nullptr;
    return scene->camera_files[camera].c_str();
}

int lrh_save_image(const char *path, const float *rgba, uint32_t width, uint32_t height) {
    if (!path || !rgba) return fail("lrh_save_image: null argument.");
    try {
        lrh::save_image(path, rgba, width, height);
        return 0;
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}

uint32_t lrh_plugin_count(void) { return static_cast<uint32_t>(lrh::registered_plugins().size()); }

const char *lrh_plugin_name(uint32_t index) {
    static thread_local std::string name;
    auto all = lrh::registered_plugins();
    if (index >= all.size()) return nullptr;
    name = all[index];
    return name.c_str();
}

int lrh_load_image(const char *path, uint32_t *width, uint32_t *height, uint32_t *channels, float *rgba, uint64_t rgba_capacity) {
    if (!path || !width || !height || !channels) return fail("lrh_load_image: null argument.");
    try {
        auto image = lrh::load_image(path);
        *width = image.width;
        *height = image.height;
        *channels = image.channels;
        if (rgba != nullptr) {
            if (rgba_capacity < image.rgba.size()) return fail("lrh_load_image: output buffer too small.");
            std::copy(image.rgba.begin(), image.rgba.end(), rgba);
        }
        return 0;
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}

int lrh_create_alias_table(const float *values, uint32_t n, float *prob, uint32_t *alias, float *pdf) {
    if (!values || !prob || !alias || !pdf) return fail("lrh_create_alias_table: null argument.");
    try {
        std::vector<lrk_alias_entry> table;
        std::vector<float> p;
        lrh::create_alias_table(values, n, table, p);
        for (uint32_t i = 0u; i < n; i++) {
            prob[i] = table[i].prob;
            alias[i] = table[i].alias;
            pdf[i] = p[i];
        }
        return 0;
    } catch (const std::exception &e) {
        return fail(e.what());
    }
}

}// extern "C"
