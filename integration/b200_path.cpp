// luisa-render-integrator-b200path — the reference-side binding of this repository's radiance library.
//
// A LuisaRender scene selects it like any other integrator plugin (src/base/scene.cpp:64-75 resolves the lower-cased
// "luisa-render-integrator-b200path" module next to the executable):
//
//     integrator : B200Path { depth { 10 } rr_depth { 0 } rr_threshold { 0.95 } sampler : Independent { seed { 19980810 } } }
//
// and luisa-render-cli renders the scene file unchanged otherwise.  The class pair below mirrors the reference's own
// integrators (src/integrators/wave_path.cpp:36-52,208-218): a scene node that reads the properties, and an Instance whose
// render(Stream &) - the one virtual an integrator must implement (src/base/integrator.h:44,55-56) - produces the film of
// every camera and writes it where the camera says (src/base/integrator.cpp:34-49).  Everything BELOW that virtual is replaced:
// no DSL recording, no JIT, no OptiX / Embree.  The per-sample radiance loop runs in libb200pt.so (include/lrk.h), fed by the
// flattened scene that libluisa_render_host.so (include/lrh.h) builds from the very scene file the reference's parser has
// just read.
//
// Why the flattened scene comes from the FILE and not from the reference's host objects: every node implementation of the
// reference is a class private to its plugin module (src/surfaces/disney.cpp, src/lights/diffuse.cpp, ...), whose parameters
// are reachable only through DSL-recording virtuals (Surface::Instance::closure, Texture::Instance::evaluate ...); the scene
// nodes do not keep their SceneNodeDesc and the Scene does not keep the root description (src/base/scene_node.h:35-54,
// src/base/scene.h:40-90).  What an integrator can reach without recording a kernel is the command line of its process
// (scene file + -D definitions, exactly what src/apps/cli.cpp:59-152 parsed) and the source location of its own description
// node; both lead back to the scene text, which the host library parses with the same grammar, plugin names and defaults.
//
// Built by oracle/ref/Makefile against the reference's headers where they lie (this file is the only source of the module),
// into oracle/_ref/bin/ next to the reference's own plugins; tests/test_reference_plugin.py runs the UNMODIFIED reference CLI with
// it on the GPU and compares the EXR with the committed WavePath render of the reference itself.
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include <base/integrator.h>
#include <base/pipeline.h>
#include <base/camera.h>
#include <base/film.h>
#include <util/imageio.h>

#include <lrh.h>
#include <lrk.h>

namespace luisa::render {

class B200Path final : public Integrator {

private:
    uint _max_depth;
    uint _rr_depth;
    float _rr_threshold;
    std::filesystem::path _scene_file;// file in which this node is written: the fallback when the process has no scene argument

public:
    B200Path(Scene *scene, const SceneNodeDesc *desc) noexcept
        : Integrator{scene, desc},
          _max_depth{std::max(desc->property_uint_or_default("depth", 10u), 1u)},                 // wave_path.cpp:43
          _rr_depth{desc->property_uint_or_default("rr_depth", 0u)},                              // :44
          _rr_threshold{std::max(desc->property_float_or_default("rr_threshold", 0.95f), 0.05f)} {// :45
        if (auto loc = desc->source_location()) { _scene_file = *loc.file(); }
    }
    [[nodiscard]] auto max_depth() const noexcept { return _max_depth; }
    [[nodiscard]] auto rr_depth() const noexcept { return _rr_depth; }
    [[nodiscard]] auto rr_threshold() const noexcept { return _rr_threshold; }
    [[nodiscard]] const auto &scene_file() const noexcept { return _scene_file; }
    [[nodiscard]] luisa::string_view impl_type() const noexcept override { return LUISA_RENDER_PLUGIN_NAME; }
    [[nodiscard]] luisa::unique_ptr<Integrator::Instance> build(Pipeline &pipeline, CommandBuffer &command_buffer) const noexcept override;
};

namespace {

// The arguments luisa-render-cli was started with: positional scene file and "-D key=value" / "--define key=value" macros
// (src/apps/cli.cpp:59-152).  Options that take a value: -b/--backend, -d/--device.
struct CommandLine {
    std::string scene;
    std::vector<std::string> keys, values;
};

CommandLine read_command_line() noexcept {
    CommandLine cl;
    std::ifstream f{"/proc/self/cmdline", std::ios::binary};
    std::vector<std::string> args;
    for (std::string a; std::getline(f, a, '\0');) { args.push_back(a); }
    auto add_macro = [&](const std::string &kv) {
        auto eq = kv.find('=');
        if (eq == std::string::npos || eq == 0u) { return; }
        cl.keys.push_back(kv.substr(0u, eq));
        cl.values.push_back(kv.substr(eq + 1u));
    };
    for (size_t i = 1u; i < args.size(); i++) {
        const auto &a = args[i];
        if (a == "-b" || a == "--backend" || a == "-d" || a == "--device") { i++; continue; }
        if (a == "-D" || a == "--define") { if (i + 1u < args.size()) { add_macro(args[++i]); } continue; }
        if (a.rfind("-D", 0) == 0 && a.size() > 2u) { add_macro(a.substr(2u)); continue; }
        if (a.rfind("--define=", 0) == 0) { add_macro(a.substr(9u)); continue; }
        if (!a.empty() && a[0] == '-') { continue; }
        cl.scene = a;
    }
    return cl;
}

}// namespace

class B200PathInstance final : public Integrator::Instance {

public:
    B200PathInstance(Pipeline &pipeline, CommandBuffer &command_buffer, const B200Path *node) noexcept
        : Integrator::Instance{pipeline, command_buffer, node} {}

    void render(Stream &stream) noexcept override {
        auto node = this->node<B200Path>();
        auto cl = read_command_line();
        std::filesystem::path scene_path = cl.scene.empty() ? node->scene_file() : std::filesystem::path{cl.scene};
        if (scene_path.empty()) { LUISA_ERROR_WITH_LOCATION("B200Path: cannot locate the scene file."); }
        std::vector<const char *> keys, values;
        for (auto &k : cl.keys) { keys.push_back(k.c_str()); }
        for (auto &v : cl.values) { values.push_back(v.c_str()); }
        lrh_scene *host_scene = nullptr;
        if (lrh_scene_load(scene_path.string().c_str(), keys.data(), values.data(), static_cast<uint32_t>(keys.size()), &host_scene) != 0) {
            LUISA_ERROR_WITH_LOCATION("B200Path: {}", lrh_last_error());
        }
        lrk_device_cfg cfg{};
        cfg.device_index = -1;
        lrk_ctx *ctx = nullptr;
        if (lrk_create(&cfg, &ctx) != LRK_OK) { LUISA_ERROR_WITH_LOCATION("B200Path: no usable CUDA device (there is no CPU fallback)."); }
        for (auto i = 0u; i < pipeline().camera_count(); i++) {
            auto camera = pipeline().camera(i);
            lrk_scene_desc desc{};
            if (lrh_scene_get_desc(host_scene, i, &desc) != 0) { LUISA_ERROR_WITH_LOCATION("B200Path: {}", lrh_last_error()); }
            auto resolution = camera->film()->node()->resolution();
            LUISA_ASSERT(desc.camera.resolution[0] == resolution.x && desc.camera.resolution[1] == resolution.y &&
                             desc.integrator.max_depth == node->max_depth() && desc.integrator.rr_depth == node->rr_depth(),
                         "B200Path: the host library read a different scene than the reference's parser.");
            auto spp = camera->node()->spp();
            Clock clock;
            if (lrk_upload_scene(ctx, &desc) != LRK_OK) { LUISA_ERROR_WITH_LOCATION("B200Path: {}", lrk_last_error(ctx)); }
            if (lrk_render(ctx, 0u, spp) != LRK_OK) { LUISA_ERROR_WITH_LOCATION("B200Path: {}", lrk_last_error(ctx)); }
            luisa::vector<float4> pixels(static_cast<size_t>(resolution.x) * resolution.y);
            if (lrk_download_film(ctx, reinterpret_cast<float *>(pixels.data())) != LRK_OK) {// == Film::download, color.cpp:99-105
                LUISA_ERROR_WITH_LOCATION("B200Path: {}", lrk_last_error(ctx));
            }
            LUISA_INFO("Rendering finished in {} ms.", clock.toc());// wave_path.cpp:565-566
            save_image(camera->node()->file(), reinterpret_cast<const float *>(pixels.data()), resolution);// integrator.cpp:47
        }
        lrk_destroy(ctx);
        lrh_scene_destroy(host_scene);
    }
};

luisa::unique_ptr<Integrator::Instance> B200Path::build(Pipeline &pipeline, CommandBuffer &command_buffer) const noexcept {
    return luisa::make_unique<B200PathInstance>(pipeline, command_buffer, this);
}

}// namespace luisa::render

LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN(luisa::render::B200Path)
