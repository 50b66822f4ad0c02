// luisa-render-cli — same command line as the reference's src/apps/cli.cpp:59-185:
//   luisa-render-cli -b <backend> [-d <index>] [-D key=value]... <scene file>
// Parses the scene (host library), uploads the flattened scene to the B200 radiance library and renders
// every camera to its `file` (default <scene dir>/render.exr).  `-b cuda` selects the sm_100a backend; any
// other backend name is an error here (north_star: no multi-backend dispatch, no CPU fallback).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/lrh.h"

namespace {

void usage() {
    std::printf("Usage:\n  luisa-render-cli [OPTION...] <file>\n\n"
                "  -b, --backend <backend>      Compute backend name (cuda)\n"
                "  -d, --device <index>         Compute device index (default: -1)\n"
                "      --scene <file>           Path to scene description file\n"
                "  -D, --define <key>=<value>   Parameter definitions to override scene description macros.\n"
                "  -h, --help                   Display this help message\n");
}

[[noreturn]] void die(const std::string &msg) {
    std::fprintf(stderr, "[error] %s\n", msg.c_str());
    std::abort();// the reference's LUISA_ERROR logs and aborts (logging.h:63,107)
}

}// namespace

int main(int argc, char *argv[]) {
    std::string backend, scene_path;
    int device = -1;
    std::vector<std::string> keys, values;
    auto add_macro = [&](const std::string &d) {
        auto p = d.find('=');
        if (p == std::string::npos) {
            std::fprintf(stderr, "[warning] Invalid definition: %s\n", d.c_str());
            return;
        }
        auto k = d.substr(0, p), v = d.substr(p + 1);
        for (size_t i = 0; i < keys.size(); i++) {
            if (keys[i] == k) {
                std::fprintf(stderr, "[warning] Duplicate definition: %s = '%s'. Ignoring the previous one.\n", k.c_str(), v.c_str());
                values[i] = v;
                return;
            }
        }
        keys.push_back(k);
        values.push_back(v);
    };
    for (int i = 1; i < argc; i++) {
        std::string a{argv[i]};
        auto need = [&](const char *what) -> std::string {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "[warning] Missing %s after %s.\n", what, a.c_str());
                usage();
                std::exit(-1);
            }
            return argv[++i];
        };
        if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (a == "-b" || a == "--backend") backend = need("backend");
        else if (a.rfind("--backend=", 0) == 0) backend = a.substr(10);
        else if (a == "-d" || a == "--device") device = std::atoi(need("index").c_str());
        else if (a.rfind("--device=", 0) == 0) device = std::atoi(a.substr(9).c_str());
        else if (a == "--scene") scene_path = need("file");
        else if (a == "-D" || a == "--define") add_macro(need("definition"));
        else if (a.rfind("-D", 0) == 0) add_macro(a.substr(2));
        else if (!a.empty() && a[0] == '-') std::fprintf(stderr, "[warning] Unrecognized options: %s\n", a.c_str());
        else scene_path = a;
    }
    if (scene_path.empty()) {
        std::fprintf(stderr, "[warning] Scene file not specified.\n");
        usage();
        return -1;
    }
    if (backend.empty()) {
        std::fprintf(stderr, "[warning] Failed to parse command line arguments: Option 'backend' has no value.\n");
        usage();
        return -1;
    }
    for (auto &c : backend) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    if (backend != "cuda") die("Backend '" + backend + "' is not available: this build ships the sm_100a CUDA backend only (-b cuda).");

    std::vector<const char *> k, v;
    for (size_t i = 0; i < keys.size(); i++) {
        std::printf("[info] Found CLI Macro: %s = %s\n", keys[i].c_str(), values[i].c_str());
        k.push_back(keys[i].c_str());
        v.push_back(values[i].c_str());
    }
    auto t0 = std::chrono::steady_clock::now();
    lrh_scene *scene = nullptr;
    if (lrh_scene_load(scene_path.c_str(), k.data(), v.data(), static_cast<uint32_t>(k.size()), &scene) != 0) die(lrh_last_error());
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    lrh_scene_info info{};
    lrh_scene_get_info(scene, &info);
    std::printf("[info] Parsed and flattened scene '%s' in %.1f ms (BVH build %.1f ms).\n", scene_path.c_str(),
                ms(t0, std::chrono::steady_clock::now()), info.bvh_build_ms);
    std::printf("[info] Geometry built with %llu triangles (%llu unique, %u meshes, %u instances).\n",
                static_cast<unsigned long long>(info.instanced_triangles), static_cast<unsigned long long>(info.unique_triangles),
                info.meshes, info.instances);

    lrk_device_cfg cfg{};
    cfg.device_index = device;
    lrk_ctx *ctx = nullptr;
    if (int rc = lrk_create(&cfg, &ctx); rc != 0) die("Failed to create the CUDA device context (lrk_create = " + std::to_string(rc) + ").");
    for (uint32_t cam = 0; cam < info.cameras; cam++) {
        lrk_scene_desc desc{};
        if (lrh_scene_get_desc(scene, cam, &desc) != 0) die(lrh_last_error());
        if (lrk_upload_scene(ctx, &desc) != 0) {
            std::string msg = lrk_last_error(ctx);
            if (msg.find("No lights in scene") != std::string::npos) {// reference: warn and skip (wave_path.cpp:224-228)
                std::fprintf(stderr, "[warning] %s\n", msg.c_str());
                continue;
            }
            die(msg);
        }
        const uint32_t w = desc.camera.resolution[0], h = desc.camera.resolution[1], spp = desc.camera.spp;
        std::printf("[info] Wavefront path tracing configurations: resolution = %ux%u, spp = %u.\n", w, h, spp);
        std::printf("[info] Rendering started.\n");
        if (lrk_render(ctx, 0u, spp) != 0) die(lrk_last_error(ctx));
        lrk_stats st{};
        lrk_get_stats(ctx, &st);
        std::printf("[info] Rendering finished in %.3f ms.\n", st.render_ms);
        std::printf("[info] %.2f Msamples/s, %.2f Mrays/s (%llu closest + %llu shadow rays).\n",
                    static_cast<double>(st.samples) / st.render_ms * 1e-3,
                    static_cast<double>(st.closest_rays + st.shadow_rays) / st.render_ms * 1e-3,
                    static_cast<unsigned long long>(st.closest_rays), static_cast<unsigned long long>(st.shadow_rays));
        std::vector<float> pixels(static_cast<size_t>(w) * h * 4u);
        if (lrk_download_film(ctx, pixels.data()) != 0) die(lrk_last_error(ctx));
        const char *file = lrh_scene_camera_file(scene, cam);
        if (lrh_save_image(file, pixels.data(), w, h) != 0) std::fprintf(stderr, "[warning] %s\n", lrh_last_error());
        else std::printf("[info] Saved film to '%s'.\n", file);
    }
    lrk_destroy(ctx);
    lrh_scene_destroy(scene);
    return 0;
}
