#include "scene.h"
#include "flatten.h"

#include "imageio.h"
#include "meshload.h"

#include <algorithm>
#include <cmath>

namespace lrh {

// ---------------------------------------------------------------- plugin registry

namespace {
struct Registry {
    std::mutex mutex;
    std::unordered_map<std::string, Plugin> plugins;
};
Registry &registry() {
    static Registry r;
    return r;
}
}// namespace

void register_plugin(const std::string &key, Plugin plugin) {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    r.plugins[key] = plugin;
}

const Plugin *find_plugin(const std::string &key) {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    auto it = r.plugins.find(key);
    return it == r.plugins.end() ? nullptr : &it->second;
}

std::vector<std::string> registered_plugins() {
    auto &r = registry();
    std::scoped_lock lock{r.mutex};
    std::vector<std::string> keys;
    for (auto &kv : r.plugins) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    return keys;
}

// same shape as LUISA_RENDER_MAKE_SCENE_NODE_PLUGIN (src/base/scene_node.h:58-67)
#define LRH_PLUGIN(key, cls)                                                                          \
    namespace {                                                                                       \
    SceneNode *create_##cls(Scene *scene, const NodeDesc *desc) { return new cls{scene, desc}; }      \
    void destroy_##cls(SceneNode *node) { delete node; }                                              \
    struct Register_##cls {                                                                           \
        Register_##cls() { register_plugin(key, Plugin{create_##cls, destroy_##cls}); }               \
    } register_##cls##_instance;                                                                      \
    }

// ---------------------------------------------------------------- Scene

Scene::~Scene() {
    for (auto it = _internal_nodes.rbegin(); it != _internal_nodes.rend(); ++it) it->destroy(it->node);
    for (auto &kv : _nodes) kv.second.destroy(kv.second.node);
}

const NodeDesc *Scene::shared_default(Tag tag, const std::string &impl) {
    std::string key{tag_description(tag)};
    key.append("-").append(impl);
    for (auto &c : key) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    std::scoped_lock lock{_mutex};
    if (auto it = _default_lookup.find(key); it != _default_lookup.end()) return it->second;
    auto d = std::make_unique<NodeDesc>("__shared_default_" + key, tag);
    d->define(tag, impl, "shared default", {});
    auto p = _default_descs.emplace_back(std::move(d)).get();
    _default_lookup.emplace(key, p);
    return p;
}

SceneNode *Scene::load_node(Tag tag, const NodeDesc *desc) {
    if (desc == nullptr) return nullptr;
    if (!desc->is_defined()) {
        throw Error("Undefined scene description node '" + desc->identifier() + "' (type = " +
                    std::string{tag_description(desc->tag())} + "::" + desc->impl_type() + ").");
    }
    std::string key{tag_description(tag)};
    key.append("-").append(desc->impl_type());
    auto plugin = find_plugin(key);
    if (plugin == nullptr) {
        throw Error("Failed to load plugin 'luisa-render-" + key + "' for scene node '" + desc->identifier() +
                    "' (not implemented in this build). [" + desc->location() + "]");
    }
    std::scoped_lock lock{_mutex};
    if (desc->is_internal()) {
        auto node = plugin->create(this, desc);
        _internal_nodes.push_back({node, plugin->destroy});
        return node;
    }
    if (desc->tag() != tag) {
        throw Error("Invalid tag " + std::string{tag_description(desc->tag())} + " of scene description node '" +
                    desc->identifier() + "' (expected " + std::string{tag_description(tag)} + "). [" + desc->location() + "]");
    }
    if (auto it = _nodes.find(desc->identifier()); it != _nodes.end()) {
        auto node = it->second.node;
        if (node->tag() != tag || node->impl_type() != desc->impl_type()) {
            throw Error("Scene node `" + desc->identifier() + "` is already in the graph with another type. [" +
                        desc->location() + "]");
        }
        return node;
    }
    auto node = plugin->create(this, desc);
    _nodes.emplace(desc->identifier(), Handle{node, plugin->destroy});
    return node;
}

std::unique_ptr<Scene> Scene::create(const SceneDesc *desc) {
    auto root = desc->root();
    if (!root->is_defined()) throw Error("Root node is not defined in the scene description.");
    std::unique_ptr<Scene> scene{new Scene};
    // order of loading follows src/base/scene.cpp:201-233
    scene->_shadow_terminator = root->f("shadow_terminator", 0.f);
    scene->_intersection_offset = root->f("intersection_offset", 0.f);
    auto spectrum_desc = root->node("spectrum");
    if (!spectrum_desc) spectrum_desc = scene->shared_default(Tag::SPECTRUM, "sRGB");
    scene->_spectrum = scene->load<Spectrum>(Tag::SPECTRUM, spectrum_desc);
    scene->_integrator = scene->load<Integrator>(Tag::INTEGRATOR, root->required_node("integrator"));
    if (auto env = root->node("environment")) scene->_environment = scene->load<Environment>(Tag::ENVIRONMENT, env);
    scene->_environment_medium = scene->load_medium(root->node("environment_medium"));
    for (auto c : root->required_nodes("cameras")) scene->_cameras.push_back(scene->load<Camera>(Tag::CAMERA, c));
    for (auto s : root->required_nodes("shapes")) scene->_shapes.push_back(scene->load_shape(s));
    return scene;
}

// ---------------------------------------------------------------- base-class constructors

static float4x4 view_matrix(float3 origin, float3 front, float3 up) {
    // src/transforms/view.cpp:27-48
    auto w = normalize(-front);
    auto u = normalize(cross(up, w));
    auto v = normalize(cross(w, u));
    float4x4 m;
    m.c[0] = {u.x, u.y, u.z, 0.f};
    m.c[1] = {v.x, v.y, v.z, 0.f};
    m.c[2] = {w.x, w.y, w.z, 0.f};
    m.c[3] = {origin.x, origin.y, origin.z, 1.f};
    return m;
}

Filter::Filter(const Scene *s, const NodeDesc *d) : SceneNode{s, d, Tag::FILTER} {
    // src/base/filter.cpp:11-17
    radius = std::max(d->f("radius", .5f), 1e-3f);
    if (!d->fN("shift", 2, shift)) shift[0] = shift[1] = d->f("shift", 0.f);
}

Sampler::Sampler(const Scene *s, const NodeDesc *d)
    : SceneNode{s, d, Tag::SAMPLER}, seed{d->u("seed", 19980810u)} {}// src/base/sampler.cpp:11

Integrator::Integrator(Scene *s, const NodeDesc *d) : SceneNode{s, d, Tag::INTEGRATOR} {
    // src/base/integrator.cpp:13-18
    auto sd = d->node("sampler");
    if (!sd) sd = s->shared_default(Tag::SAMPLER, "independent");
    sampler = s->load_sampler(sd);
    auto ld = d->node("light_sampler");
    if (!ld) ld = s->shared_default(Tag::LIGHT_SAMPLER, "uniform");
    light_sampler = s->load_light_sampler(ld);
}

Camera::Camera(Scene *s, const NodeDesc *d) : SceneNode{s, d, Tag::CAMERA} {
    // src/base/camera.cpp:16-50,137-147
    film = s->load_film(d->required_node("film"));
    auto fd = d->node("filter");
    if (!fd) fd = s->shared_default(Tag::FILTER, "Box");
    filter = s->load_filter(fd);
    transform = s->load_transform(d->node("transform"));
    camera_to_world = transform ? transform->matrix() : float4x4::identity();
    if (transform == nullptr) {
        // compatibility with older scene files: position / front | look_at / up (src/base/camera.cpp:30-50)
        float p[3]{0.f, 0.f, 0.f}, f[3], up[3]{0.f, 1.f, 0.f};
        d->fN("position", 3, p);
        if (!d->fN("front", 3, f)) {
            float la[3]{p[0], p[1], p[2] - 1.f};
            d->fN("look_at", 3, la);
            auto n = normalize(float3{la[0] - p[0], la[1] - p[1], la[2] - p[2]});
            f[0] = n.x; f[1] = n.y; f[2] = n.z;
        }
        d->fN("up", 3, up);
        bool is_default = p[0] == 0.f && p[1] == 0.f && p[2] == 0.f && f[0] == 0.f && f[1] == 0.f && f[2] == -1.f &&
                          up[0] == 0.f && up[1] == 1.f && up[2] == 0.f;
        if (!is_default) camera_to_world = view_matrix({p[0], p[1], p[2]}, {f[0], f[1], f[2]}, {up[0], up[1], up[2]});
    }
    spp = d->u("spp", 1024u);
    float span[2];
    if (!d->fN("shutter_span", 2, span)) span[0] = span[1] = d->f("shutter_span", 0.f);
    if (span[0] != span[1]) throw Error("Motion blur (shutter_span) is not supported. [" + d->location() + "]");
    if (auto f = d->string("file")) {
        std::filesystem::path p{*f};
        file = p.is_absolute() ? p : d->source_dir() / p;
    } else {
        file = (d->source_dir().empty() ? std::filesystem::current_path() : d->source_dir()) / "render.exr";
    }
}

Shape::Shape(Scene *s, const NodeDesc *d) : SceneNode{s, d, Tag::SHAPE} {
    // src/base/shape.cpp:15-20
    surface = s->load_surface(d->node("surface"));
    light = s->load_light(d->node("light"));
    transform = s->load_transform(d->node("transform"));
    medium = s->load_medium(d->node("medium"));
}

const std::vector<lrk_vertex> &Shape::vertices() const {
    static const std::vector<lrk_vertex> empty;
    return empty;
}
const std::vector<lrk_triangle> &Shape::triangles() const {
    static const std::vector<lrk_triangle> empty;
    return empty;
}

// ---------------------------------------------------------------- textures / spectrum

namespace {

struct ConstantTexture final : Texture {
    // src/textures/constant.cpp:24-47
    float4 v{};
    uint32_t n{0};
    bool black{false};
    ConstantTexture(Scene *s, const NodeDesc *d) : Texture{s, d, Tag::TEXTURE} {
        auto scale = d->f("scale", 1.f);
        auto values = d->float_list("v");
        if (values.empty()) values.push_back(0.f);
        if (values.size() > 4) values.resize(4);
        n = static_cast<uint32_t>(values.size());
        for (uint32_t i = 0; i < n; i++) v[i] = scale * values[i];
        black = v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f;
    }
    bool is_black() const override { return black; }
    bool is_constant() const override { return true; }
    uint32_t channels() const override { return n; }
    float4 value() const override { return v; }
};

struct ImageTexture final : Texture {
    // src/textures/image.cpp:16-131
    LoadedImage image;
    lrk_texture rec{};
    ImageTexture(Scene *s, const NodeDesc *d) : Texture{s, d, Tag::TEXTURE} {
        auto lower = [](std::string v) {
            for (auto &c : v) c = static_cast<char>(std::tolower(c));
            return v;
        };
        auto filter = lower(d->s("filter", "bilinear")), address = lower(d->s("address", "repeat"));
        if (address == "zero") rec.address = LRK_TEX_ADDRESS_ZERO;
        else if (address == "edge") rec.address = LRK_TEX_ADDRESS_EDGE;
        else if (address == "mirror") rec.address = LRK_TEX_ADDRESS_MIRROR;
        else if (address == "repeat") rec.address = LRK_TEX_ADDRESS_REPEAT;
        else throw Error("Invalid texture address mode '" + address + "'. [" + d->location() + "]");
        // the reference samples level 0 without LOD for every non-point filter (image.cpp:165, mipmap generation is a TODO
        // there, :190-200), so bilinear / trilinear / anisotropic all reduce to the bilinear fetch
        if (filter == "point") rec.filter = LRK_TEX_FILTER_POINT;
        else if (filter == "bilinear" || filter == "trilinear" || filter == "anisotropic" || filter == "aniso") rec.filter = LRK_TEX_FILTER_LINEAR;
        else throw Error("Invalid texture filter mode '" + filter + "'. [" + d->location() + "]");
        auto two = [&](const char *name, float dflt, float out[2]) {
            auto v = d->float_list(name);
            if (v.size() >= 2) {
                out[0] = v[0];
                out[1] = v[1];
            } else {
                out[0] = out[1] = v.empty() ? dflt : v[0];
            }
        };
        two("uv_scale", 1.f, rec.uv_scale);
        two("uv_offset", 0.f, rec.uv_offset);
        auto path = d->path("file");
        auto ext = lower(path.extension().string());
        auto encoding = lower(d->s("encoding", (ext == ".exr" || ext == ".hdr") ? "linear" : "srgb"));
        rec.gamma = 1.f;
        if (encoding == "srgb") {
            rec.encoding = LRK_TEX_ENCODING_SRGB;
        } else if (encoding == "gamma") {
            rec.encoding = LRK_TEX_ENCODING_GAMMA;
            rec.gamma = d->f("gamma", 1.f);
        } else {
            rec.encoding = LRK_TEX_ENCODING_LINEAR;// unknown encodings fall back to linear with a warning in the reference
        }
        rec.scale = d->f("scale", 1.f);
        try {
            image = load_image(path);
        } catch (const std::exception &e) {
            throw Error(std::string{e.what()} + " [" + d->location() + "]");
        }
        rec.width = image.width;
        rec.height = image.height;
        rec.channels = image.channels;
    }
    bool is_black() const override { return rec.scale == 0.f; }
    bool is_constant() const override { return false; }
    bool is_image() const override { return true; }
    uint32_t channels() const override { return image.channels; }
    float4 value() const override { throw Error("ImageTexture has no constant value."); }
    void emit(lrk_texture &out, std::vector<float> &texels) const override {
        auto offset = out.texel_offset;
        out = rec;
        out.texel_offset = offset;
        texels.insert(texels.end(), image.rgba.begin(), image.rgba.end());
    }
};

float3 extend_color_to_rgb(float4 c, uint32_t n);

struct SwizzleTexture final : Texture {
    // src/textures/swizzle.cpp:17-104: picks / reorders the channels of another texture.  The decode of an image texture
    // (encoding, scale) acts per channel, so swizzling an image is a permutation of its texels' channels: host-only.
    const Texture *base;
    std::vector<uint32_t> pick;
    SwizzleTexture(Scene *s, const NodeDesc *d) : Texture{s, d, Tag::TEXTURE} {
        base = s->load_texture(d->required_node("base"));
        if (d->numbers("swizzle")) {
            pick = d->uint_list("swizzle");// throws on negative / fractional entries, like the reference's property_uint_list
        } else {
            for (auto c : d->s("swizzle", "rgba")) {
                switch (c) {
                    case 'r': case 'x': pick.push_back(0u); break;
                    case 'g': case 'y': pick.push_back(1u); break;
                    case 'b': case 'z': pick.push_back(2u); break;
                    case 'a': case 'w': pick.push_back(3u); break;
                    default: throw Error(std::string{"Invalid swizzle channel '"} + c + "'. [" + d->location() + "]");
                }
            }
        }
        if (pick.size() > 4u) pick.resize(4u);// the reference warns and discards the rest
        if (pick.empty()) throw Error("Swizzle channel index out of range. [" + d->location() + "]");
        for (auto c : pick)
            if (c >= 4u) throw Error("Swizzle channel '" + std::to_string(c) + "' out of range. [" + d->location() + "]");
    }
    // SwizzleTextureInstance::evaluate (:87-94): 1 -> (a,a,a,a), 2 -> (a,b,0,1), 3 -> (a,b,c,1), 4 -> (a,b,c,d)
    void apply(const float in[4], float out[4]) const {
        float v[4]{in[0], in[1], in[2], in[3]};
        switch (pick.size()) {
            case 1u: out[0] = out[1] = out[2] = out[3] = v[pick[0]]; break;
            case 2u: out[0] = v[pick[0]]; out[1] = v[pick[1]]; out[2] = 0.f; out[3] = 1.f; break;
            case 3u: out[0] = v[pick[0]]; out[1] = v[pick[1]]; out[2] = v[pick[2]]; out[3] = 1.f; break;
            default: out[0] = v[pick[0]]; out[1] = v[pick[1]]; out[2] = v[pick[2]]; out[3] = v[pick[3]]; break;
        }
    }
    bool is_black() const override { return base->is_black(); }
    bool is_constant() const override { return base->is_constant(); }
    bool is_image() const override { return base->is_image(); }
    uint32_t channels() const override { return static_cast<uint32_t>(pick.size()); }
    float4 value() const override {// evaluate_static (:65-72): unpicked channels stay 0
        auto b = base->value();
        float in[4]{b.x, b.y, b.z, b.w};
        float4 out{0.f, 0.f, 0.f, 0.f};
        float *o = &out.x;
        for (size_t i = 0; i < pick.size(); i++) o[i] = in[pick[i]];
        return out;
    }
    void emit(lrk_texture &out, std::vector<float> &texels) const override {
        auto first = texels.size();
        base->emit(out, texels);
        out.channels = channels();
        for (auto i = first; i + 4u <= texels.size(); i += 4u) apply(&texels[i], &texels[i]);
    }
};

struct CheckerboardTexture final : Texture {
    // src/textures/checkerboard.cpp:19-74 with a CONSTANT on texture and a constant or absent (black) off texture: the selector
    // (int(floor(u sx)) + int(floor(v sy))) % 2 == 0 is what a point-sampled, repeating 2x2 image gives at uv * (scale / 2)
    // (halving commutes with every rounding involved), so the checkerboard is baked into four texels: host-only.
    const Texture *on{}, *off{};
    float scale[2]{1.f, 1.f};
    CheckerboardTexture(Scene *s, const NodeDesc *d) : Texture{s, d, Tag::TEXTURE} {
        if (auto n = d->node("on")) on = s->load_texture(n);
        if (auto n = d->node("off")) off = s->load_texture(n);
        // an absent 'on' is 1 through evaluate() (scalar parameters) but a ZERO spectrum through evaluate_albedo_spectrum /
        // evaluate_illuminant_spectrum (checkerboard.cpp:78-83, 93-98: SampledSpectrum{n} is all zeros): one texture record
        // cannot be both, so the ambiguous form is refused; an absent 'off' is zero either way
        if (on == nullptr)
            throw Error("Checkerboard: 'on' must be given (the reference evaluates an absent 'on' as 1 for scalar parameters and as "
                        "black for colours). [" + d->location() + "]");
        if ((on && !on->is_constant()) || (off && !off->is_constant()))
            throw Error("Checkerboard: only constant 'on' / 'off' textures are supported. [" + d->location() + "]");
        auto v = d->float_list("scale");
        if (v.size() >= 2u) { scale[0] = v[0]; scale[1] = v[1]; }
        else if (v.size() == 1u) { scale[0] = scale[1] = v[0]; }
    }
    bool is_black() const override { return (on != nullptr && on->is_black()) && (off == nullptr || off->is_black()); }
    bool is_constant() const override { return false; }// no static value in the reference either (no evaluate_static)
    bool is_image() const override { return true; }
    // The reference decodes each square by ITS OWN channel count (checkerboard.cpp:76-105: extend_color_to_rgb per child), so a
    // grey `on` next to an RGB `off` stays grey and RGB.  The baked texels are therefore extended per child, and the record
    // reports at least three channels; scalar consumers read .x, which the extension keeps.
    uint32_t channels() const override { return std::max(3u, std::min(on ? on->channels() : 4u, off ? off->channels() : 4u)); }
    float4 value() const override { throw Error("Checkerboard has no constant value."); }
    void emit(lrk_texture &out, std::vector<float> &texels) const override {
        auto offset = out.texel_offset;
        out = lrk_texture{};
        out.texel_offset = offset;
        out.width = out.height = 2u;
        out.channels = channels();
        out.address = LRK_TEX_ADDRESS_REPEAT;
        out.filter = LRK_TEX_FILTER_POINT;
        out.encoding = LRK_TEX_ENCODING_LINEAR;
        out.gamma = 1.f;
        out.scale = 1.f;
        out.uv_scale[0] = scale[0] * 0.5f;
        out.uv_scale[1] = scale[1] * 0.5f;
        out.uv_offset[0] = out.uv_offset[1] = 0.f;
        auto square = [](const Texture *t, float4 absent) {
            if (t == nullptr) return absent;
            auto v = t->value();
            if (t->channels() < 3u) {
                auto rgb = extend_color_to_rgb(v, t->channels());
                v = float4{rgb.x, rgb.y, rgb.z, v.w};
            }
            return v;
        };
        auto a = square(on, float4{1.f, 1.f, 1.f, 1.f});
        auto b = square(off, float4{0.f, 0.f, 0.f, 0.f});
        for (int y = 0; y < 2; y++)
            for (int x = 0; x < 2; x++) {
                auto c = ((x + y) % 2 == 0) ? a : b;
                texels.insert(texels.end(), {c.x, c.y, c.z, c.w});
            }
    }
};

struct SRGBSpectrum final : Spectrum {
    SRGBSpectrum(Scene *s, const NodeDesc *d) : Spectrum{s, d, Tag::SPECTRUM} {}
};

// src/base/texture.cpp:15-19
float3 extend_color_to_rgb(float4 c, uint32_t n) {
    if (n == 1u) return {c.x, c.x, c.x};
    if (n == 2u) return {c.x, c.y, 1.f};
    return {c.x, c.y, c.z};
}
float saturate(float x) { return std::min(std::max(x, 0.f), 1.f); }
// albedo decode with the sRGB spectrum: src/spectra/srgb.cpp:18,34-40; returns luminance strength
float3 decode_albedo(const Texture *t, float *strength) {
    float3 rgb{1.f, 1.f, 1.f};
    if (t != nullptr) {
        auto c = extend_color_to_rgb(t->value(), t->channels());
        rgb = {saturate(c.x), saturate(c.y), saturate(c.z)};
    }
    // src/util/colorspace.h:21-25
    if (strength) *strength = 0.212671f * rgb.x + 0.715160f * rgb.y + 0.072169f * rgb.z;
    return rgb;
}
const Texture *constant_or_null(Scene *s, const NodeDesc *d, const char *name) {
    auto t = s->load_texture(d->node(name));
    if (t && !t->is_constant()) throw Error("Only constant textures are supported ('" + std::string{name} + "').");
    return t;
}
// surface parameters: constant or image texture
const Texture *surface_texture(Scene *s, const NodeDesc *d, const char *name) {
    auto t = s->load_texture(d->node(name));
    if (t && !t->is_constant() && !t->is_image()) throw Error("Only constant and image textures are supported ('" + std::string{name} + "').");
    return t;
}

}// namespace
LRH_PLUGIN("texture-constant", ConstantTexture)
LRH_PLUGIN("texture-image", ImageTexture)
LRH_PLUGIN("texture-swizzle", SwizzleTexture)
LRH_PLUGIN("texture-checkerboard", CheckerboardTexture)
LRH_PLUGIN("spectrum-srgb", SRGBSpectrum)

// ---------------------------------------------------------------- transforms

namespace {

struct IdentityTransform final : Transform {
    IdentityTransform(Scene *s, const NodeDesc *d) : Transform{s, d, Tag::TRANSFORM} {}
    float4x4 matrix() const override { return float4x4::identity(); }
};

struct MatrixTransform final : Transform {
    // src/transforms/matrix.cpp:17-43 : row-major list of 16
    float4x4 m{float4x4::identity()};
    MatrixTransform(Scene *s, const NodeDesc *d) : Transform{s, d, Tag::TRANSFORM} {
        auto v = d->float_list("m");
        if (v.size() == 16u) {
            v[12] = 0.f; v[13] = 0.f; v[14] = 0.f; v[15] = 1.f;
            for (int row = 0; row < 4; row++)
                for (int col = 0; col < 4; col++) m.c[col][row] = v[row * 4 + col];
        } else if (!v.empty()) {
            throw Error("Invalid matrix entries. [" + d->location() + "]");
        }
    }
    float4x4 matrix() const override { return m; }
};

float4x4 translation(float3 v) {
    auto m = float4x4::identity();
    m.c[3] = {v.x, v.y, v.z, 1.f};
    return m;
}
float4x4 scaling(float3 s) {
    auto m = float4x4::identity();
    m.c[0].x = s.x; m.c[1].y = s.y; m.c[2].z = s.z;
    return m;
}
// src/compute/include/luisa/core/mathematics.h:470-481
float4x4 rotation(float3 axis, float angle) {
    if (angle == 0.0f) return float4x4::identity();
    auto c = std::cos(angle);
    auto s = std::sin(angle);
    auto a = normalize(axis);
    auto t = (1.0f - c) * a;
    float4x4 m;
    m.c[0] = {c + t.x * a.x, t.x * a.y + s * a.z, t.x * a.z - s * a.y, 0.0f};
    m.c[1] = {t.y * a.x - s * a.z, c + t.y * a.y, t.y * a.z + s * a.x, 0.0f};
    m.c[2] = {t.z * a.x + s * a.y, t.z * a.y - s * a.x, c + t.z * a.z, 0.0f};
    m.c[3] = {0.0f, 0.0f, 0.0f, 1.0f};
    return m;
}

struct SRTTransform final : Transform {
    // src/transforms/srt.cpp:18-28
    float4x4 m;
    SRTTransform(Scene *s, const NodeDesc *d) : Transform{s, d, Tag::TRANSFORM} {
        float sc[3], rot[4], tr[3];
        if (!d->fN("scale", 3, sc)) sc[0] = sc[1] = sc[2] = d->f("scale", 1.f);
        if (!d->fN("rotate", 4, rot)) { rot[0] = 0.f; rot[1] = 0.f; rot[2] = 1.f; rot[3] = 0.f; }
        if (!d->fN("translate", 3, tr)) tr[0] = tr[1] = tr[2] = 0.f;
        m = translation({tr[0], tr[1], tr[2]}) *
            rotation(normalize(float3{rot[0], rot[1], rot[2]}), radians(rot[3])) *
            scaling({sc[0], sc[1], sc[2]});
    }
    float4x4 matrix() const override { return m; }
};

struct ViewTransform final : Transform {
    float4x4 m;
    ViewTransform(Scene *s, const NodeDesc *d) : Transform{s, d, Tag::TRANSFORM} {
        float o[3], f[3], up[3];
        if (!d->fN("origin", 3, o) && !d->fN("position", 3, o)) o[0] = o[1] = o[2] = 0.f;
        if (!d->fN("front", 3, f)) { f[0] = 0.f; f[1] = 0.f; f[2] = -1.f; }
        if (!d->fN("up", 3, up)) { up[0] = 0.f; up[1] = 1.f; up[2] = 0.f; }
        m = view_matrix({o[0], o[1], o[2]}, {f[0], f[1], f[2]}, {up[0], up[1], up[2]});
    }
    float4x4 matrix() const override { return m; }
};

struct StackTransform final : Transform {
    // src/transforms/stack.cpp:24-37 : later entries are applied after earlier ones
    float4x4 m{float4x4::identity()};
    bool identity{true};
    StackTransform(Scene *s, const NodeDesc *d) : Transform{s, d, Tag::TRANSFORM} {
        for (auto c : d->nodes("transforms")) {
            auto t = s->load_transform(c);
            identity = identity && t->is_identity();
            m = t->matrix() * m;
        }
    }
    float4x4 matrix() const override { return m; }
    bool is_identity() const override { return identity; }
};

}// namespace
LRH_PLUGIN("transform-identity", IdentityTransform)
LRH_PLUGIN("transform-matrix", MatrixTransform)
LRH_PLUGIN("transform-srt", SRTTransform)
LRH_PLUGIN("transform-view", ViewTransform)
LRH_PLUGIN("transform-stack", StackTransform)

// ---------------------------------------------------------------- filters

namespace {
struct BoxFilter final : Filter {
    BoxFilter(Scene *s, const NodeDesc *d) : Filter{s, d} {}
    float evaluate(float) const override { return 1.0f; }
};
struct TriangleFilter final : Filter {
    TriangleFilter(Scene *s, const NodeDesc *d) : Filter{s, d} {}
    float evaluate(float x) const override { return std::max(1.0f - std::abs(x / radius), 0.0f); }
};
struct GaussianFilter final : Filter {
    // src/filters/gaussian.cpp:16-37
    float sigma;
    GaussianFilter(Scene *s, const NodeDesc *d) : Filter{s, d}, sigma{d->f("sigma", 0.f)} {
        if (sigma <= 0.f) sigma = radius / 3.f;
    }
    float G(float x) const {
        auto s2 = 2.0f * sigma * sigma;
        return 1.0f / std::sqrt(kPi * s2) * std::exp(-x * x / s2);
    }
    float evaluate(float x) const override { return G(x) - G(radius); }
};
struct MitchellFilter final : Filter {
    // src/filters/mitchell.cpp:17-38
    float B, C;
    MitchellFilter(Scene *s, const NodeDesc *d) : Filter{s, d}, B{d->f("b", 1.0f / 3.0f)}, C{d->f("c", 1.0f / 3.0f)} {}
    float evaluate(float x) const override {
        x = 2.f * std::abs(x / radius);
        if (x <= 1.0f)
            return ((12.0f - 9.0f * B - 6.0f * C) * x * x * x + (-18.0f + 12.0f * B + 6.0f * C) * x * x + (6.0f - 2.0f * B)) * (1.f / 6.f);
        if (x <= 2.0f)
            return ((-B - 6.0f * C) * x * x * x + (6.0f * B + 30.0f * C) * x * x + (-12.0f * B - 48.0f * C) * x + (8.0f * B + 24.0f * C)) * (1.0f / 6.0f);
        return 0.0f;
    }
};
struct LanczosSincFilter final : Filter {
    // src/filters/lanczos_sinc.cpp:16-31
    float tau;
    LanczosSincFilter(Scene *s, const NodeDesc *d) : Filter{s, d}, tau{d->f("tau", 3.0f)} {}
    static float sinc(float x) {
        x = kPi * x;
        return 1.0f + x * x == 1.0f ? 1.0f : std::sin(x) / x;
    }
    float evaluate(float x) const override {
        x = x / radius;
        if (std::abs(x) > 1.0f) return 0.0f;
        return sinc(x) * sinc(x / tau);
    }
};
}// namespace
LRH_PLUGIN("filter-box", BoxFilter)
LRH_PLUGIN("filter-triangle", TriangleFilter)
LRH_PLUGIN("filter-gaussian", GaussianFilter)
LRH_PLUGIN("filter-mitchell", MitchellFilter)
LRH_PLUGIN("filter-lanczossinc", LanczosSincFilter)

// ---------------------------------------------------------------- film / sampler / light sampler / integrators

namespace {

struct ColorFilm final : Film {
    // src/films/color.cpp:26-42
    ColorFilm(Scene *s, const NodeDesc *d) : Film{s, d, Tag::FILM} {
        auto r = d->uint_list("resolution");
        if (r.size() >= 2) { resolution[0] = r[0]; resolution[1] = r[1]; }
        else { resolution[0] = resolution[1] = d->u("resolution", 1024u); }
        float e[3];
        if (!d->fN("exposure", 3, e)) e[0] = e[1] = e[2] = d->f("exposure", 0.f);
        for (int i = 0; i < 3; i++) scale[i] = std::pow(2.0f, e[i]);
        clamp = std::max(1.f, d->f("clamp", 256.f));
        if (resolution[0] == 0 || resolution[1] == 0) throw Error("Invalid film resolution. [" + d->location() + "]");
        // color.cpp:32,124-129: a debugging aid that overwrites a pixel with (inf, 0, 0, 1) when a NaN / infinite sample arrives; the
        // accumulate kernel drops such samples like the default does and has no marking pass - refuse rather than differ silently
        if (d->b("warn_nan", false)) throw Error("Film 'Color': warn_nan { true } is not supported (NaN / infinite samples are dropped, not marked). [" + d->location() + "]");
    }
};

struct IndependentSampler final : Sampler {
    IndependentSampler(Scene *s, const NodeDesc *d) : Sampler{s, d} {}
};
// the table-driven samplers of row f2: nothing but `seed` in their descriptions (src/samplers/*.cpp); tables and the values
// Sampler::Instance::reset derives from resolution / spp are attached per camera by flatten_sampler (flatten.cpp)
template<uint32_t TYPE>
struct TableSampler final : Sampler {
    TableSampler(Scene *s, const NodeDesc *d) : Sampler{s, d} { type = TYPE; }
};
using PMJ02BNSampler = TableSampler<LRK_SAMPLER_PMJ02BN>;
using SobolSampler = TableSampler<LRK_SAMPLER_SOBOL>;
using PaddedSobolSampler = TableSampler<LRK_SAMPLER_PADDED_SOBOL>;
using ZSobolSampler = TableSampler<LRK_SAMPLER_ZSOBOL>;

struct UniformLightSampler final : LightSampler {
    UniformLightSampler(Scene *s, const NodeDesc *d) : LightSampler{s, d, Tag::LIGHT_SAMPLER} {
        environment_weight = d->f("environment_weight", .5f);// src/lightsamplers/uniform.cpp:17-19
    }
};

// wavepath / wavepath_v2 / megapath share one estimator; parameters and defaults from
// src/integrators/wave_path.cpp:41-46 (identical in mega_path.cpp:22-29 and wave_path_v2.cpp:60-72).
struct PathIntegrator final : Integrator {
    PathIntegrator(Scene *s, const NodeDesc *d) : Integrator{s, d} {
        kind = LRK_INTEGRATOR_PATH;
        max_depth = std::max(d->u("depth", 10u), 1u);
        rr_depth = d->u("rr_depth", 0u);
        rr_threshold = std::max(d->f("rr_threshold", 0.95f), 0.05f);
        samples_per_pass = std::max(d->u("samples_per_pass", 16u), 1u);
    }
};

// src/integrators/mega_vpt_naive.cpp:33-42
struct VolumePathIntegrator final : Integrator {
    VolumePathIntegrator(Scene *s, const NodeDesc *d) : Integrator{s, d} {
        kind = LRK_INTEGRATOR_VOLUME_PATH;
        max_depth = std::max(d->u("depth", 10u), 1u);
        rr_depth = d->u("rr_depth", 0u);
        rr_threshold = std::max(d->f("rr_threshold", 0.95f), 0.05f);
        samples_per_pass = std::max(d->u("samples_per_pass", 16u), 1u);
    }
};

}// namespace
LRH_PLUGIN("film-color", ColorFilm)
LRH_PLUGIN("sampler-independent", IndependentSampler)
LRH_PLUGIN("sampler-pmj02bn", PMJ02BNSampler)
LRH_PLUGIN("sampler-sobol", SobolSampler)
LRH_PLUGIN("sampler-paddedsobol", PaddedSobolSampler)
LRH_PLUGIN("sampler-zsobol", ZSobolSampler)
LRH_PLUGIN("lightsampler-uniform", UniformLightSampler)
LRH_PLUGIN("integrator-wavepath", PathIntegrator)
namespace {
SceneNode *create_path_alias(Scene *scene, const NodeDesc *desc) { return new PathIntegrator{scene, desc}; }
void destroy_path_alias(SceneNode *n) { delete n; }
struct RegisterPathAliases {
    RegisterPathAliases() {
        register_plugin("integrator-wavepath_v2", Plugin{create_path_alias, destroy_path_alias});
        register_plugin("integrator-megapath", Plugin{create_path_alias, destroy_path_alias});
        // the name under which the reference loads this repository's integrator plugin (integration/b200_path.cpp): a scene
        // written for it parses here as the path integrator it is
        register_plugin("integrator-b200path", Plugin{create_path_alias, destroy_path_alias});
    }
} register_path_aliases_instance;
}// namespace
LRH_PLUGIN("integrator-megavptnaive", VolumePathIntegrator)

// ---------------------------------------------------------------- media

namespace {

struct HenyeyGreenstein final : PhaseFunction {
    HenyeyGreenstein(Scene *s, const NodeDesc *d) : PhaseFunction{s, d, Tag::PHASE_FUNCTION} {
        g = std::min(std::max(d->f("g", 0.f), -1.f), 1.f);// src/phasefunctions/henyey_greenstein.cpp:66-68
    }
};

struct HomogeneousMedium final : Medium {
    // src/media/homogeneous.cpp:188-199
    HomogeneousMedium(Scene *s, const NodeDesc *d) : Medium{s, d, Tag::MEDIUM} {
        priority = d->u("priority", 0u);
        eta = d->f("eta", 1.f);
        auto sa = constant_or_null(s, d, "sigma_a");
        auto ss = constant_or_null(s, d, "sigma_s");
        auto le_t = constant_or_null(s, d, "Le");
        phase = s->load_phase_function(d->node("phasefunction"));
        if (!sa) throw Error("sigma_a must be specified as constant. [" + d->location() + "]");
        if (!ss) throw Error("sigma_s must be specified as constant. [" + d->location() + "]");
        if (!phase) throw Error("Phase function must be specified. [" + d->location() + "]");
        auto a = extend_color_to_rgb(sa->value(), sa->channels());
        auto sc = extend_color_to_rgb(ss->value(), ss->channels());
        for (int i = 0; i < 3; i++) { sigma_a[i] = a[i]; sigma_s[i] = sc[i]; }
        if (le_t) {
            auto e = extend_color_to_rgb(le_t->value(), le_t->channels());
            for (int i = 0; i < 3; i++) le[i] = std::max(e[i], 0.f);
        }
    }
};
struct NullMedium final : Medium {
    NullMedium(Scene *s, const NodeDesc *d) : Medium{s, d, Tag::MEDIUM} {}
    bool is_null() const override { return true; }
};
struct VacuumMedium final : Medium {
    VacuumMedium(Scene *s, const NodeDesc *d) : Medium{s, d, Tag::MEDIUM} { priority = d->u("priority", 0u); }
    bool is_vacuum() const override { return true; }
};

}// namespace
LRH_PLUGIN("phasefunction-henyeygreenstein", HenyeyGreenstein)
LRH_PLUGIN("medium-homogeneous", HomogeneousMedium)
LRH_PLUGIN("medium-null", NullMedium)
LRH_PLUGIN("medium-vacuum", VacuumMedium)

// ---------------------------------------------------------------- environments

Environment::Environment(Scene *scene, const NodeDesc *desc) : SceneNode{scene, desc, Tag::ENVIRONMENT} {
    transform = scene->load_transform(desc->node("transform"));
}

namespace {

struct SphericalEnvironment final : Environment {
    SphericalEnvironment(Scene *s, const NodeDesc *d) : Environment{s, d} {
        emission = s->load_texture(d->required_node("emission"));
        if (!emission->is_constant() && !emission->is_image())
            throw Error("Only constant and image textures are supported ('emission'). [" + d->location() + "]");
        scale = std::max(d->f("scale", 1.0f), 0.0f);
        compensate_mis = d->b("compensate_mis", true);
    }
    bool is_black() const override { return scale == 0.0f || emission->is_black(); }
};

struct NullEnvironment final : Environment {
    NullEnvironment(Scene *s, const NodeDesc *d) : Environment{s, d} {}
    bool is_null() const override { return true; }
    bool is_black() const override { return true; }
};

}// namespace
LRH_PLUGIN("environment-spherical", SphericalEnvironment)
LRH_PLUGIN("environment-null", NullEnvironment)

// ---------------------------------------------------------------- surfaces / lights

Surface::Surface(Scene *scene, const NodeDesc *desc, Tag tag) : SceneNode{scene, desc, tag} {
    if (desc == nullptr || desc->impl_type() == "null") return;
    auto load = [&](const char *name) -> const Texture * {
        auto t = scene->load_texture(desc->node(name));
        if (t && !t->is_constant() && !t->is_image())
            throw Error("Only constant and image textures are supported ('" + std::string{name} + "'). [" + desc->location() + "]");
        return t;
    };
    opacity = load(desc->has_property("alpha") ? "alpha" : "opacity");// surface.h:200-206
    normal_map = load("normal_map");
    normal_map_strength = desc->f("normal_map_strength", 1.f);
}

bool Surface::maybe_non_opaque() const {
    if (opacity == nullptr) return false;
    return opacity->is_image() ? true : opacity->value().x < 1.f;// evaluate_static().value_or(0).x < 1
}

void Surface::flatten_wrappers(lrk_surface &out, TextureTable &textures) const {
    out.opacity = 1.f;
    out.normal_strength = 1.f;
    if (maybe_non_opaque()) {
        out.flags |= LRK_SURFACE_MAYBE_NON_OPAQUE;
        if (opacity->is_image()) out.opacity_tex = textures.slot(opacity);
        else out.opacity = opacity->value().x;
    }
    if (normal_map != nullptr) {
        out.flags |= LRK_SURFACE_HAS_NORMAL_MAP;
        out.normal_strength = normal_map_strength;
        if (normal_map->is_image()) {
            out.normal_tex = textures.slot(normal_map);
        } else {
            auto v = normal_map->value();
            out.normal_value[0] = v.x;
            out.normal_value[1] = v.y;
            out.normal_value[2] = v.z;
        }
    }
}

namespace {

struct MatteSurface final : Surface {
    // src/surfaces/matte.cpp:22-27,119-134
    const Texture *kd;
    const Texture *sigma;
    MatteSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        kd = surface_texture(s, d, "Kd");
        sigma = surface_texture(s, d, "sigma");
    }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_MATTE;
        if (kd && kd->is_image()) {
            out.tex[0] = textures.slot(kd);
        } else {
            auto c = decode_albedo(kd, nullptr);
            out.p[0] = c.x; out.p[1] = c.y; out.p[2] = c.z;
        }
        if (sigma && !sigma->is_black()) {
            if (sigma->is_image()) out.tex[3] = textures.slot(sigma);
            else out.p[3] = saturate(sigma->value().x) * 90.f;
        }
        if (out.tex[0] || out.tex[3]) out.flags |= LRK_SURFACE_HAS_TEXTURES;
        flatten_wrappers(out, textures);
        return out;
    }
};

struct DisneySurface final : Surface {
    // src/surfaces/disney.cpp:36-58,932-998
    const Texture *color, *metallic, *eta, *roughness, *specular_tint, *anisotropic, *sheen, *sheen_tint,
        *clearcoat, *clearcoat_gloss, *specular_trans, *flatness, *diffuse_trans;
    bool thin, remap_roughness;
    DisneySurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        color = surface_texture(s, d, d->has_property("color") ? "color" : "Kd");
        thin = d->b("thin", false);
        remap_roughness = d->b("remap_roughness", true);
        metallic = surface_texture(s, d, "metallic");
        eta = surface_texture(s, d, "eta");
        roughness = surface_texture(s, d, "roughness");
        specular_tint = surface_texture(s, d, "specular_tint");
        anisotropic = surface_texture(s, d, "anisotropic");
        sheen = surface_texture(s, d, "sheen");
        sheen_tint = surface_texture(s, d, "sheen_tint");
        clearcoat = surface_texture(s, d, "clearcoat");
        clearcoat_gloss = surface_texture(s, d, "clearcoat_gloss");
        specular_trans = surface_texture(s, d, "specular_trans");
        flatness = surface_texture(s, d, "flatness");
        diffuse_trans = surface_texture(s, d, "diffuse_trans");
    }
    // src/surfaces/disney.cpp:61-75: the closure classes "disney_thin" (a thin node with either transmission) and "disney_trans"
    // (a specular-transmission lobe exists); a thin node without transmission is an ordinary opaque one
    bool is_thin() const {
        return thin && ((specular_trans != nullptr && !specular_trans->is_black()) || (diffuse_trans != nullptr && !diffuse_trans->is_black()));
    }
    bool is_transmissive() const { return !thin && specular_trans != nullptr && !specular_trans->is_black(); }
    uint32_t lobes() const {
        // src/surfaces/disney.cpp:966-990
        uint32_t l = 0u;
        if (!color || !color->is_black()) {
            l |= LRK_DISNEY_LOBE_DIFFUSE | LRK_DISNEY_LOBE_RETRO;
            if (sheen && !sheen->is_black()) l |= LRK_DISNEY_LOBE_SHEEN;
            if (flatness && !flatness->is_black()) l |= LRK_DISNEY_LOBE_FAKE_SS;
        }
        l |= LRK_DISNEY_LOBE_SPECULAR;
        if (clearcoat && !clearcoat->is_black()) l |= LRK_DISNEY_LOBE_CLEARCOAT;
        if (specular_trans && !specular_trans->is_black()) l |= LRK_DISNEY_LOBE_SPEC_TRANS;
        if (diffuse_trans && !diffuse_trans->is_black()) l |= LRK_DISNEY_LOBE_DIFF_TRANS;// the opaque closures never look at these two bits
        return l;
    }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_DISNEY;
        out.lobes = lobes();
        if (remap_roughness) out.flags |= LRK_SURFACE_REMAP_ROUGHNESS;
        if (is_transmissive()) out.flags |= LRK_SURFACE_DISNEY_TRANSMISSIVE;
        if (is_thin()) out.flags |= LRK_SURFACE_DISNEY_THIN;
        // a parameter is either the node's constant (p[k]) or an image texture evaluated per hit (tex[k])
        auto x = [&](const Texture *t, float dflt, uint32_t slot) {
            if (t && t->is_image()) {
                out.tex[slot] = textures.slot(t);
                out.flags |= LRK_SURFACE_HAS_TEXTURES;
                return dflt;
            }
            return t ? t->value().x : dflt;
        };
        if (color && color->is_image()) {
            out.tex[0] = textures.slot(color);
            out.flags |= LRK_SURFACE_HAS_TEXTURES;
        } else {
            float lum;
            auto c = decode_albedo(color, &lum);
            out.p[0] = c.x; out.p[1] = c.y; out.p[2] = c.z; out.p[3] = lum;
        }
        out.p[4] = x(metallic, 0.f, 4);
        out.p[5] = x(eta, 1.5f, 5);
        auto r = x(roughness, .5f, 6);
        if (remap_roughness) r = std::max(r * r, 1e-4f);// roughness_to_alpha, src/util/scattering.cpp:137-139
        out.p[6] = r;
        out.p[7] = x(specular_tint, 0.f, 7);
        out.p[8] = x(anisotropic, 0.f, 8);
        out.p[9] = x(sheen, 0.f, 9);
        out.p[10] = x(sheen_tint, 0.f, 10);
        out.p[11] = x(clearcoat, 0.f, 11);
        out.p[12] = x(clearcoat_gloss, 1.f, 12);
        out.p[13] = x(specular_trans, 0.f, 13);
        out.p[14] = x(flatness, 0.f, 14);
        out.p[15] = thin ? x(diffuse_trans, 0.f, 15) : 0.f;// diffuse_trans is only built for thin surfaces (src/surfaces/disney.cpp:1024)
        flatten_wrappers(out, textures);
        return out;
    }
};

// roughness -> alpha of Mirror / Glass / Plastic / Metal (e.g. mirror.cpp:144-155): one channel is used for both axes,
// remap_roughness applies TrowbridgeReitzDistribution::roughness_to_alpha = max(r^2, 1e-4) (scattering.cpp:129-135)
void flatten_alpha(const Texture *roughness, bool remap, float dflt, float out[2]) {
    out[0] = out[1] = dflt;
    if (roughness == nullptr) return;
    auto r = roughness->value();
    auto r2a = [](float x) { return std::max(x * x, 1e-4f); };
    if (roughness->channels() == 1u) {
        out[0] = out[1] = remap ? r2a(r.x) : r.x;
    } else {
        out[0] = remap ? r2a(r.x) : r.x;
        out[1] = remap ? r2a(r.y) : r.y;
    }
}

// Mirror / Glass / Plastic / Metal with an image-textured parameter: the record carries the node's RAW
// parameters (LRK_SURFACE_RAW_PARAMS) and the closure context is derived per hit, on the device, from the raw values and the
// textures (include/lrk.h gives the layouts).  These helpers fill one raw slot.
bool any_image(std::initializer_list<const Texture *> list) {
    for (auto t : list) if (t != nullptr && t->is_image()) return true;
    return false;
}
void raw_colour(lrk_surface &out, TextureTable &textures, const Texture *t, uint32_t slot, float3 dflt) {
    if (t != nullptr && t->is_image()) {
        out.tex[slot] = textures.slot(t);
    } else {
        auto c = t != nullptr ? decode_albedo(t, nullptr) : dflt;
        out.p[slot] = c.x; out.p[slot + 1u] = c.y; out.p[slot + 2u] = c.z;
    }
}
// constant / absent roughness: the final alpha (as flatten_alpha); image roughness: remapped on the device when the flag is set
void raw_alpha(lrk_surface &out, TextureTable &textures, const Texture *roughness, bool remap, float dflt, uint32_t slot) {
    if (roughness != nullptr && roughness->is_image()) {
        out.tex[slot] = textures.slot(roughness);
        if (remap) out.flags |= LRK_SURFACE_REMAP_ROUGHNESS;
    } else {
        flatten_alpha(roughness, remap, dflt, &out.p[slot]);
    }
}
float raw_scalar(lrk_surface &out, TextureTable &textures, const Texture *t, uint32_t slot, float dflt) {
    if (t != nullptr && t->is_image()) {
        out.tex[slot] = textures.slot(t);
        return dflt;
    }
    return t != nullptr ? t->value().x : dflt;
}

const Texture *constant_surface_texture(Scene *s, const NodeDesc *d, const char *name) {
    auto t = s->load_texture(d->node(name));
    if (t && !t->is_constant())
        throw Error("Only constant textures are supported for '" + std::string{name} + "' of " + std::string{d->impl_type()} +
                    " surfaces. [" + d->location() + "]");
    return t;
}

struct MirrorSurface final : Surface {
    // src/surfaces/mirror.cpp:13-36,142-162
    const Texture *color, *roughness;
    bool remap_roughness;
    MirrorSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        color = surface_texture(s, d, d->has_property("color") ? "color" : "Kd");
        roughness = surface_texture(s, d, "roughness");
        remap_roughness = d->b("remap_roughness", true);
    }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_MIRROR;
        if (any_image({color, roughness})) {// raw layout: p[0..2] colour, p[3..4] alpha / roughness
            out.flags |= LRK_SURFACE_HAS_TEXTURES | LRK_SURFACE_RAW_PARAMS;
            raw_colour(out, textures, color, 0u, float3{1.f, 1.f, 1.f});
            raw_alpha(out, textures, roughness, remap_roughness, 0.f, 3u);
            flatten_wrappers(out, textures);
            return out;
        }
        auto c = decode_albedo(color, nullptr);
        out.p[0] = c.x; out.p[1] = c.y; out.p[2] = c.z;
        flatten_alpha(roughness, remap_roughness, 0.f, &out.p[3]);
        flatten_wrappers(out, textures);
        return out;
    }
};

struct GlassSurface final : Surface {
    // src/surfaces/glass.cpp:54-91,222-279
    const Texture *kr, *kt, *roughness, *eta{};
    float named_eta{0.f};
    bool remap_roughness;
    GlassSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        kr = surface_texture(s, d, "Kr");
        kt = surface_texture(s, d, "Kt");
        roughness = surface_texture(s, d, "roughness");
        remap_roughness = d->b("remap_roughness", true);
        if (auto name = d->s("eta", ""); !name.empty()) {
            // built-in glasses (glass.cpp:28-41): refractive index at the Fraunhofer C line (656.27 nm) - with the fixed sRGB
            // spectrum the closure reads the first component of the (C, d, F) triple only (glass.cpp:262-264)
            static const std::pair<const char *, float> known[] = {
                {"bk7", 1.5140814565098806f}, {"baf10", 1.665552211440938f}, {"fk51a", 1.4846524304153899f},
                {"lasf9", 1.8422161861952726f}, {"sf5", 1.6663001504164476f}, {"sf10", 1.720557014155419f},
                {"sf11", 1.7754589288508518f}, {"diamond", 2.410486117067883f}, {"ice", 1.3077084260466776f},
                {"quartz", 1.4562471554155727f}, {"salt", 1.5404463273409252f}, {"sapphire", 1.764706495252994f}};
            for (auto &c : name) c = static_cast<char>(std::tolower(c));
            for (auto &&[k, v] : known) if (name == k) named_eta = v;
            // unknown names: the reference warns and falls back to 1.5 (glass.cpp:65-71)
        } else {
            eta = constant_surface_texture(s, d, "eta");
            if (eta && (eta->channels() == 2u || eta->channels() == 4u))
                throw Error("Invalid channel count for GlassSurface::eta. [" + d->location() + "]");
        }
    }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_GLASS;
        if (any_image({kr, kt, roughness})) {// raw layout: p[0..2] Kr, p[3..5] Kt, p[6] eta, p[7..8] alpha / roughness; p[9] derived per hit
            out.flags |= LRK_SURFACE_HAS_TEXTURES | LRK_SURFACE_RAW_PARAMS;
            raw_colour(out, textures, kr, 0u, float3{1.f, 1.f, 1.f});
            raw_colour(out, textures, kt, 3u, float3{1.f, 1.f, 1.f});
            out.p[6] = named_eta != 0.f ? named_eta : (eta ? eta->value().x : 1.5f);
            raw_alpha(out, textures, roughness, remap_roughness, 0.f, 7u);
            flatten_wrappers(out, textures);
            return out;
        }
        float kr_lum = 1.f, kt_lum = 1.f;
        auto r = kr ? decode_albedo(kr, &kr_lum) : float3{1.f, 1.f, 1.f};
        auto t = kt ? decode_albedo(kt, &kt_lum) : float3{1.f, 1.f, 1.f};
        out.p[0] = r.x; out.p[1] = r.y; out.p[2] = r.z;
        out.p[3] = t.x; out.p[4] = t.y; out.p[5] = t.z;
        out.p[6] = named_eta != 0.f ? named_eta : (eta ? eta->value().x : 1.5f);
        flatten_alpha(roughness, remap_roughness, 0.f, &out.p[7]);
        out.p[9] = kr_lum == 0.f ? 0.f : kr_lum / (kr_lum + kt_lum);
        flatten_wrappers(out, textures);
        return out;
    }
};

struct PlasticSurface final : Surface {
    // src/surfaces/plastic.cpp:42-66,252-291
    const Texture *kd, *roughness, *sigma_a, *eta, *thickness;
    bool remap_roughness;
    PlasticSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        kd = surface_texture(s, d, "Kd");
        roughness = surface_texture(s, d, "roughness");
        sigma_a = surface_texture(s, d, "sigma_a");
        eta = constant_surface_texture(s, d, "eta");
        thickness = surface_texture(s, d, "thickness");
        remap_roughness = d->b("remap_roughness", true);
    }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_PLASTIC;
        auto e = (eta ? eta->value().x : 1.5f) / 1.f;// eta_i = 1
        if (any_image({kd, roughness, sigma_a, thickness})) {
            // raw layout: p[0..2] Kd, p[4..6] sigma_a, p[7] eta, p[8..9] alpha / roughness, p[10] thickness; p[0..3] derived per hit
            out.flags |= LRK_SURFACE_HAS_TEXTURES | LRK_SURFACE_RAW_PARAMS;
            raw_colour(out, textures, kd, 0u, float3{1.f, 1.f, 1.f});
            raw_colour(out, textures, sigma_a, 4u, float3{0.f, 0.f, 0.f});
            out.p[7] = e;
            raw_alpha(out, textures, roughness, remap_roughness, 0.f, 8u);
            out.p[10] = raw_scalar(out, textures, thickness, 10u, 1.f);
            flatten_wrappers(out, textures);
            return out;
        }
        float kd_lum = 1.f, sa_lum = 0.f;
        auto c = kd ? decode_albedo(kd, &kd_lum) : float3{1.f, 1.f, 1.f};
        auto sa = sigma_a ? decode_albedo(sigma_a, &sa_lum) : float3{0.f, 0.f, 0.f};
        auto th = thickness ? thickness->value().x : 1.f;
        auto average_transmittance = std::exp(-2.f * sa_lum * th);
        // fresnel_dielectric_integral (scattering.cpp:98-108): the fitted polynomials, Horner from the last coefficient
        auto saturate1 = [](float x) { return std::fmin(std::fmax(x, 0.f), 1.f); };
        float fdr;
        if (e == 1.f) fdr = 0.f;
        else if (e < 1.f) fdr = e * (e * (e * -0.90663979f + 2.23559031f) + -2.09069066f) + 0.75985009f;
        else { auto x = 1.f / e; fdr = x * (x * -1.18995376f + 0.21762732f) + 0.97945724f; }
        fdr = saturate1(fdr);
        out.p[0] = c.x / (1.f - c.x * fdr);
        out.p[1] = c.y / (1.f - c.y * fdr);
        out.p[2] = c.z / (1.f - c.z * fdr);
        out.p[3] = kd_lum * average_transmittance;
        out.p[4] = sa.x; out.p[5] = sa.y; out.p[6] = sa.z;
        out.p[7] = e;
        flatten_alpha(roughness, remap_roughness, 0.f, &out.p[8]);
        flatten_wrappers(out, textures);
        return out;
    }
};

struct MetalSurface final : Surface {
    // src/surfaces/metal.cpp:52-151,273-310.  The complex index comes from an `eta` list of (wavelength, n, k) triples or from
    // the name of one of the reference's eleven built-in metals (measured spectra: luisarender_b200/data/metal_ior.bin).
    const Texture *roughness, *kd;
    bool remap_roughness;
    float n[3], k[3];
    MetalSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        roughness = surface_texture(s, d, "roughness");
        kd = surface_texture(s, d, "Kd");
        remap_roughness = d->b("remap_roughness", true);
        constexpr uint32_t lut_size = (830u - 360u) / 5u + 1u;
        std::vector<float> lut_n(lut_size), lut_k(lut_size);
        if (auto name = d->s("eta", ""); !name.empty()) {
            // a named metal (:73-104): measured (n, k) on the same 5 nm grid, luisarender_b200/data/metal_ior.bin
            for (auto &c : name) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
            static const std::pair<const char *, const char *> names[] = {
                {"ag", "Ag"}, {"silver", "Ag"}, {"al", "Al"}, {"aluminium", "Al"}, {"au", "Au"}, {"gold", "Au"}, {"cu", "Cu"}, {"copper", "Cu"},
                {"cuzn", "CuZn"}, {"cu-zn", "CuZn"}, {"brass", "CuZn"}, {"fe", "Fe"}, {"iron", "Fe"}, {"ti", "Ti"}, {"titanium", "Ti"},
                {"v", "V"}, {"vanadium", "V"}, {"vn", "VN"}, {"li", "Li"}, {"lithium", "Li"}, {"cr", "Cr"}, {"chromium", "Cr"}};
            std::string table = "Al";// "Unknown metal ... Fallback to Aluminium" (:98-103)
            bool known = false;
            for (auto &[alias, t] : names) if (name == alias) { table = t; known = true; }
            if (!known) std::fprintf(stderr, "[warning] Unknown metal '%s'. Fallback to Aluminium. [%s]\n", name.c_str(), d->location().c_str());
            const auto &ior = metal_ior_table(table);
            std::copy(ior.n.begin(), ior.n.end(), lut_n.begin());
            std::copy(ior.k.begin(), ior.k.end(), lut_k.begin());
        } else {
            auto eta = d->float_list("eta");
            if (eta.empty() || eta.size() % 3u != 0u) throw Error("Invalid eta list size. [" + d->location() + "]");
            auto count = eta.size() / 3u;
            std::vector<float> lambda(count), nn(count), kk(count);
            for (size_t i = 0; i < count; i++) { lambda[i] = eta[i * 3u]; nn[i] = eta[i * 3u + 1u]; kk[i] = eta[i * 3u + 2u]; }
            if (!std::is_sorted(lambda.begin(), lambda.end())) throw Error("Unsorted wavelengths in eta list. [" + d->location() + "]");
            if (lambda.front() > 360.f || lambda.back() < 830.f) throw Error("Invalid wavelength range in eta list. [" + d->location() + "]");
            if (count < 2u) throw Error("Invalid eta list size. [" + d->location() + "]");
            // the 5 nm look-up table over [360, 830] nm (:132-146) ...
            for (uint32_t i = 0; i < lut_size; i++) {
                auto wavelength = static_cast<float>(i * 5u + 360u);
                auto lb = std::lower_bound(lambda.begin(), lambda.end(), wavelength);
                auto index = std::clamp(static_cast<size_t>(std::distance(lambda.begin(), lb)), size_t{1u}, lambda.size() - 1u);
                auto t = (wavelength - lambda[index - 1u]) / (lambda[index] - lambda[index - 1u]);
                // std::lerp (C++20) as libstdc++ / libc++ implement it: exact at the end points, monotonic
                auto std_lerp = [](float a, float b, float tt) {
                    if ((a <= 0.f && b >= 0.f) || (a >= 0.f && b <= 0.f)) return tt * b + (1.f - tt) * a;
                    if (tt == 1.f) return b;
                    auto x = a + tt * (b - a);
                    return (tt > 1.f) == (b > a) ? (b < x ? x : b) : (x < b ? x : b);
                };
                lut_n[i] = std_lerp(nn[index - 1u], nn[index], t);
                lut_k[i] = std_lerp(kk[index - 1u], kk[index], t);
            }
        }
        // ... sampled at the sRGB spectrum's three wavelengths (srgb.cpp:27-33, spec.h:22-23) by SPD::sample (spd.cpp:91-99)
        const float peaks[3] = {602.785f, 539.285f, 445.772f};
        for (int c = 0; c < 3; c++) {
            auto t = (std::fmin(std::fmax(peaks[c], 360.f), 830.f) - 360.f) / 5.f;
            auto i = static_cast<uint32_t>(std::fmin(t, static_cast<float>(lut_size - 2u)));
            auto fr = t - std::floor(t);
            n[c] = fr * (lut_n[i + 1u] - lut_n[i]) + lut_n[i];// lerp(a, b, t) = t * (b - a) + a
            k[c] = fr * (lut_k[i + 1u] - lut_k[i]) + lut_k[i];
        }
    }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_METAL;
        for (int c = 0; c < 3; c++) { out.p[c] = n[c]; out.p[3 + c] = k[c]; }
        if (any_image({kd, roughness})) {// raw layout = the context layout: p[6..8] Kd, p[9..10] alpha / roughness
            out.flags |= LRK_SURFACE_HAS_TEXTURES | LRK_SURFACE_RAW_PARAMS;
            raw_colour(out, textures, kd, 6u, float3{1.f, 1.f, 1.f});
            raw_alpha(out, textures, roughness, remap_roughness, .5f, 9u);
            flatten_wrappers(out, textures);
            return out;
        }
        auto r = kd ? decode_albedo(kd, nullptr) : float3{1.f, 1.f, 1.f};
        out.p[6] = r.x; out.p[7] = r.y; out.p[8] = r.z;
        flatten_alpha(roughness, remap_roughness, .5f, &out.p[9]);
        flatten_wrappers(out, textures);
        return out;
    }
};

struct MixSurface final : Surface {
    // src/surfaces/mix.cpp:22-31,195-211: ratio = clamp(ratio.x, 0, 1), default 0.5
    const Surface *a, *b;
    const Texture *ratio;
    MixSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        a = s->load_surface(d->required_node("a"));
        b = s->load_surface(d->required_node("b"));
        ratio = constant_surface_texture(s, d, "ratio");
        if (a == nullptr || b == nullptr || a->is_null() || b->is_null()) throw Error("MixSurface: Both surfaces must be valid. [" + d->location() + "]");
        for (auto c : {a, b}) {
            if (c->mix_children().first != nullptr) throw Error("Nested Mix surfaces are not supported. [" + d->location() + "]");
            if (c->opacity != nullptr || c->normal_map != nullptr)
                throw Error("Mix: surfaces with opacity / normal maps cannot be mixed. [" + d->location() + "]");
        }
    }
    std::pair<const Surface *, const Surface *> mix_children() const override { return {a, b}; }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_MIX;
        out.p[0] = ratio ? std::fmin(std::fmax(ratio->value().x, 0.f), 1.f) : .5f;
        flatten_wrappers(out, textures);
        return out;// mix_a / mix_b are filled by flatten_scene, which owns the record array
    }
};

struct LayeredSurface final : Surface {
    // src/surfaces/layered.cpp:110-141,478-503: two interfaces around a scattering slab; thickness = max(t.x, FLT_MIN) (default 1e-2),
    // g (default 0), albedo (default 1), max_depth (10), samples (1).  The interfaces become two extra records, like a Mix's.
    const Surface *top, *bottom;
    const Texture *thickness, *g, *albedo;
    uint32_t max_depth, samples;
    LayeredSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {
        top = s->load_surface(d->required_node("top"));
        bottom = s->load_surface(d->required_node("bottom"));
        thickness = constant_surface_texture(s, d, "thickness");
        g = constant_surface_texture(s, d, "g");
        albedo = constant_surface_texture(s, d, "albedo");
        max_depth = d->u("max_depth", 10u);
        samples = d->u("samples", 1u);
        if (top == nullptr || bottom == nullptr || top->is_null() || bottom->is_null())
            throw Error("Creating closure for null LayeredSurface. [" + d->location() + "]");
        if (max_depth > 0xffffu || samples == 0u || samples > 0xffffu) throw Error("Layered: max_depth / samples out of range. [" + d->location() + "]");
        for (auto c : {top, bottom}) {
            if (c->mix_children().first != nullptr) throw Error("Layered: Mix / Layered interfaces are not supported. [" + d->location() + "]");
            if (c->opacity != nullptr || c->normal_map != nullptr)
                throw Error("Layered: interfaces with opacity / normal maps are not supported. [" + d->location() + "]");
        }
    }
    std::pair<const Surface *, const Surface *> mix_children() const override { return {top, bottom}; }
    lrk_surface flatten(TextureTable &textures) const override {
        lrk_surface out{};
        out.type = LRK_SURFACE_LAYERED;
        out.p[0] = thickness ? std::max(thickness->value().x, std::numeric_limits<float>::min()) : 1e-2f;
        out.p[1] = g ? g->value().x : 0.f;
        auto a = albedo ? decode_albedo(albedo, nullptr) : float3{1.f, 1.f, 1.f};
        out.p[2] = a.x; out.p[3] = a.y; out.p[4] = a.z;
        out.lobes = max_depth | (samples << 16u);
        flatten_wrappers(out, textures);
        return out;// mix_a (top) / mix_b (bottom) are filled by flatten_scene
    }
};

struct NullSurface final : Surface {
    NullSurface(Scene *s, const NodeDesc *d) : Surface{s, d, Tag::SURFACE} {}
    bool is_null() const override { return true; }
    lrk_surface flatten(TextureTable &) const override { throw Error("NullSurface cannot be instantiated."); }
};

struct DiffuseLight final : Light {
    // src/lights/diffuse.cpp:21-30
    const Texture *emission;
    float scale;
    bool two_sided;
    DiffuseLight(Scene *s, const NodeDesc *d) : Light{s, d, Tag::LIGHT} {
        auto e = d->node("emission");
        if (!e) e = s->shared_default(Tag::TEXTURE, "Constant");
        emission = s->load_texture(e);
        if (!emission->is_constant() && !emission->is_image()) throw Error("Only constant and image emission textures are supported.");
        scale = std::max(d->f("scale", 1.0f), 0.0f);
        two_sided = d->b("two_sided", false);
    }
    bool is_null() const override { return scale == 0.0f || emission->is_black(); }
    lrk_light flatten(TextureTable &textures) const override {
        lrk_light out{};
        if (emission->is_image()) {
            out.emission_tex = textures.slot(emission);// evaluated per point (texture.cpp:47-57)
        } else {
            auto c = extend_color_to_rgb(emission->value(), emission->channels());
            out.emission[0] = std::max(c.x, 0.f);
            out.emission[1] = std::max(c.y, 0.f);
            out.emission[2] = std::max(c.z, 0.f);
        }
        out.scale = scale;
        out.two_sided = two_sided ? 1u : 0u;
        return out;
    }
};

struct NullLight final : Light {
    NullLight(Scene *s, const NodeDesc *d) : Light{s, d, Tag::LIGHT} {}
    bool is_null() const override { return true; }
    lrk_light flatten(TextureTable &) const override { throw Error("NullLight cannot be instantiated."); }
};

}// namespace
LRH_PLUGIN("surface-matte", MatteSurface)
LRH_PLUGIN("surface-disney", DisneySurface)
LRH_PLUGIN("surface-mirror", MirrorSurface)
LRH_PLUGIN("surface-glass", GlassSurface)
LRH_PLUGIN("surface-plastic", PlasticSurface)
LRH_PLUGIN("surface-metal", MetalSurface)
LRH_PLUGIN("surface-layered", LayeredSurface)
LRH_PLUGIN("surface-mix", MixSurface)
LRH_PLUGIN("surface-null", NullSurface)
LRH_PLUGIN("light-diffuse", DiffuseLight)
LRH_PLUGIN("light-null", NullLight)

// ---------------------------------------------------------------- cameras

namespace {

struct PinholeCamera final : Camera {
    // src/cameras/pinhole.cpp:33-37
    float fov;
    PinholeCamera(Scene *s, const NodeDesc *d) : Camera{s, d} {
        fov = radians(std::min(std::max(d->f("fov", 35.0f), 1e-3f), 180.f - 1e-3f));
    }
    float tan_half_fov() const override { return std::tan(fov * 0.5f); }
};

}// namespace
LRH_PLUGIN("camera-pinhole", PinholeCamera)

// ---------------------------------------------------------------- shapes

namespace {

// VisibilityShapeWrapper / ShadowTerminatorShapeWrapper / IntersectionOffsetShapeWrapper: src/base/shape.h:66-115
void read_mesh_wrappers(Shape *shape, Scene *s, const NodeDesc *d) {
    shape->visible = d->b("visible", true);
    shape->shadow_terminator = std::min(std::max(d->f("shadow_terminator", s->shadow_terminator_factor()), 0.f), 1.f);
    shape->intersection_offset = std::min(std::max(d->f("intersection_offset", s->intersection_offset_factor()), 0.f), 1.f);
}

struct InlineMesh final : Shape {
    // src/shapes/inline_mesh.cpp:21-71
    std::vector<lrk_vertex> verts;
    std::vector<lrk_triangle> tris;
    uint32_t props{0};
    InlineMesh(Scene *s, const NodeDesc *d) : Shape{s, d} {
        read_mesh_wrappers(this, s, d);
        if (!d->has_property("indices") || !d->has_property("positions"))
            throw Error("No valid values given for property 'indices'/'positions' in scene description node '" +
                        d->identifier() + "'. [" + d->location() + "]");
        auto indices = d->uint_list("indices");
        auto positions = d->float_list("positions");
        auto normals = d->float_list("normals");
        auto uvs = d->float_list("uvs");
        if (indices.size() % 3u != 0u || positions.size() % 3u != 0u || normals.size() % 3u != 0u || uvs.size() % 2u != 0u ||
            (!normals.empty() && normals.size() != positions.size()) ||
            (!uvs.empty() && uvs.size() / 2u != positions.size() / 3u)) {
            throw Error("Invalid vertex or triangle count. [" + d->location() + "]");
        }
        props = (!uvs.empty() ? LRK_SHAPE_HAS_VERTEX_UV : 0u) | (!normals.empty() ? LRK_SHAPE_HAS_VERTEX_NORMAL : 0u);
        auto nv = positions.size() / 3u;
        tris.resize(indices.size() / 3u);
        for (size_t i = 0; i < tris.size(); i++) {
            tris[i] = {indices[i * 3], indices[i * 3 + 1], indices[i * 3 + 2]};
            if (tris[i].i0 >= nv || tris[i].i1 >= nv || tris[i].i2 >= nv)
                throw Error("Triangle index out of range. [" + d->location() + "]");
        }
        verts.resize(nv);
        for (size_t i = 0; i < nv; i++) {
            auto &v = verts[i];
            v.p[0] = positions[i * 3]; v.p[1] = positions[i * 3 + 1]; v.p[2] = positions[i * 3 + 2];
            if (normals.empty()) { v.n[0] = 0.f; v.n[1] = 0.f; v.n[2] = 1.f; }
            else { v.n[0] = normals[i * 3]; v.n[1] = normals[i * 3 + 1]; v.n[2] = normals[i * 3 + 2]; }
            if (uvs.empty()) { v.uv[0] = v.uv[1] = 0.f; }
            else { v.uv[0] = uvs[i * 2]; v.uv[1] = uvs[i * 2 + 1]; }
        }
        if (verts.empty() || tris.empty()) throw Error("Empty mesh. [" + d->location() + "]");
    }
    bool is_mesh() const override { return true; }
    uint32_t vertex_properties() const override { return props; }
    const std::vector<lrk_vertex> &vertices() const override { return verts; }
    const std::vector<lrk_triangle> &triangles() const override { return tris; }
};

struct MeshShape final : Shape {
    // src/shapes/mesh.cpp:148-162 ; files are cached per (path, options) like the reference's lru cache (:32-39)
    std::shared_ptr<const MeshData> data;
    MeshShape(Scene *s, const NodeDesc *d) : Shape{s, d} {
        read_mesh_wrappers(this, s, d);
        auto path = d->path("file");
        if (d->u("subdivision", 0u) != 0u)
            throw Error("Mesh subdivision (Catmull-Clark) is not supported. [" + d->location() + "]");
        bool flip_uv = d->b("flip_uv", false), drop_normal = d->b("drop_normal", false), drop_uv = d->b("drop_uv", false);
        std::error_code ec;
        auto canonical = std::filesystem::weakly_canonical(path, ec);
        auto key = (ec ? path : canonical).string() + (flip_uv ? "|f" : "|-") + (drop_normal ? "n" : "-") + (drop_uv ? "u" : "-");
        static std::mutex mutex;
        static std::unordered_map<std::string, std::shared_ptr<const MeshData>> cache;
        std::scoped_lock lock{mutex};
        auto it = cache.find(key);
        if (it == cache.end()) {
            try {
                it = cache.emplace(key, std::make_shared<const MeshData>(load_mesh(path, flip_uv, drop_normal, drop_uv))).first;
            } catch (const std::exception &e) {
                throw Error(std::string{e.what()} + " [" + d->location() + "]");
            }
        }
        data = it->second;
    }
    bool is_mesh() const override { return true; }
    uint32_t vertex_properties() const override { return data->properties; }
    const std::vector<lrk_vertex> &vertices() const override { return data->vertices; }
    const std::vector<lrk_triangle> &triangles() const override { return data->triangles; }
};

struct SphereShape final : Shape {
    // src/shapes/sphere.cpp:103-118 ; geometry cached per subdivision level like the reference (:88-100)
    struct Geometry {
        std::vector<lrk_vertex> v;
        std::vector<lrk_triangle> t;
    };
    const Geometry *geom;
    SphereShape(Scene *s, const NodeDesc *d) : Shape{s, d} {
        read_mesh_wrappers(this, s, d);
        auto level = std::min(d->u("subdivision", 0u), 8u);
        static std::mutex mutex;
        static std::unique_ptr<Geometry> cache[9];
        std::scoped_lock lock{mutex};
        if (!cache[level]) {
            auto g = std::make_unique<Geometry>();
            make_sphere(level, g->v, g->t);
            cache[level] = std::move(g);
        }
        geom = cache[level].get();
    }
    bool is_mesh() const override { return true; }
    uint32_t vertex_properties() const override { return LRK_SHAPE_HAS_VERTEX_NORMAL | LRK_SHAPE_HAS_VERTEX_UV; }
    const std::vector<lrk_vertex> &vertices() const override { return geom->v; }
    const std::vector<lrk_triangle> &triangles() const override { return geom->t; }
};

struct LoopSubdivShape final : Shape {
    // src/shapes/loop_subdiv.cpp:14-62: Loop subdivision (limit surface, limit normals, uvs dropped) of another mesh shape;
    // only the base shape's geometry is used - surface, light, transform and the wrappers are this node's own
    const Shape *base;
    std::vector<lrk_vertex> verts;
    std::vector<lrk_triangle> tris;
    bool subdivided{false};
    LoopSubdivShape(Scene *s, const NodeDesc *d) : Shape{s, d} {
        read_mesh_wrappers(this, s, d);
        auto n = d->node("mesh");
        if (!n) n = d->node("shape");
        if (!n) n = d->required_node("base");
        base = s->load_shape(n);
        if (!base->is_mesh()) throw Error("LoopSubdiv only supports mesh shapes. [" + d->location() + "]");
        auto level = std::min(d->u("level", 1u), 10u);
        if (level != 0u) {
            loop_subdivide_mesh(base->vertices(), base->triangles(), level, verts, tris);
            subdivided = true;
        }
    }
    bool is_mesh() const override { return true; }
    uint32_t vertex_properties() const override { return subdivided ? LRK_SHAPE_HAS_VERTEX_NORMAL : base->vertex_properties(); }
    const std::vector<lrk_vertex> &vertices() const override { return subdivided ? verts : base->vertices(); }
    const std::vector<lrk_triangle> &triangles() const override { return subdivided ? tris : base->triangles(); }
};

struct InstanceShape final : Shape {
    const Shape *child;
    InstanceShape(Scene *s, const NodeDesc *d) : Shape{s, d} {
        visible = d->b("visible", true);
        child = s->load_shape(d->required_node("shape"));
    }
    std::vector<const Shape *> children() const override { return {child}; }
};

struct GroupShape final : Shape {
    std::vector<const Shape *> kids;
    GroupShape(Scene *s, const NodeDesc *d) : Shape{s, d} {
        visible = d->b("visible", true);
        for (auto c : d->required_nodes("shapes")) kids.push_back(s->load_shape(c));
    }
    std::vector<const Shape *> children() const override { return kids; }
};

}// namespace
LRH_PLUGIN("shape-inlinemesh", InlineMesh)
LRH_PLUGIN("shape-mesh", MeshShape)
LRH_PLUGIN("shape-sphere", SphereShape)
LRH_PLUGIN("shape-loopsubdiv", LoopSubdivShape)
LRH_PLUGIN("shape-instance", InstanceShape)
LRH_PLUGIN("shape-group", GroupShape)

}// namespace lrh
