// Scene description layer: the SceneDesc / SceneNodeDesc tree plus the text (.luisa) and JSON
// parsers.  Grammar, tag aliases, property typing and base-node inheritance follow the reference
// (src/sdl/scene_parser.cpp:72-451, scene_parser_json.cpp:22-196, scene_node_tag.cpp:15-46,
// scene_node_desc.h:212-361) so the same scene files load unchanged.  Errors throw lrh::Error
// (the reference logs and aborts; the C-ABI/CLI turn the exception into that behaviour).
#pragma once
#include <filesystem>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <variant>
#include <vector>

namespace lrh {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

enum class Tag : uint32_t {
    ROOT, INTERNAL, DECLARATION,
    CAMERA, SHAPE, SURFACE, LIGHT, TRANSFORM, FILM, FILTER, SAMPLER, INTEGRATOR, LIGHT_SAMPLER,
    ENVIRONMENT, TEXTURE, TEXTURE_MAPPING, SPECTRUM, MEDIUM, PHASE_FUNCTION
};

// lower-case category name used in plugin keys "luisa-render-<tag>-<impl>" (src/base/scene.cpp:67)
std::string_view tag_description(Tag tag);
// parse a tag or one of its aliases, case-insensitively; ROOT when unknown
Tag parse_tag(std::string_view s);

class NodeDesc {
public:
    using BoolList = std::vector<bool>;
    using NumberList = std::vector<double>;
    using StringList = std::vector<std::string>;
    using NodeList = std::vector<const NodeDesc *>;
    using Value = std::variant<BoolList, NumberList, StringList, NodeList>;

    NodeDesc(std::string identifier, Tag tag) : _identifier{std::move(identifier)}, _tag{tag} {}
    NodeDesc(const NodeDesc &) = delete;
    NodeDesc &operator=(const NodeDesc &) = delete;

    const std::string &identifier() const { return _identifier; }
    Tag tag() const { return _tag; }
    const std::string &impl_type() const { return _impl; }// lower-cased
    const std::string &location() const { return _location; }
    const std::filesystem::path &source_dir() const { return _source_dir; }
    bool is_defined() const { return _tag != Tag::DECLARATION && !_impl.empty(); }
    bool is_internal() const { return _tag == Tag::INTERNAL; }

    void define(Tag tag, std::string_view impl, std::string location, std::filesystem::path dir,
                const NodeDesc *base = nullptr);
    NodeDesc *define_internal(std::string_view impl, std::string location, std::filesystem::path dir,
                              const NodeDesc *base = nullptr);
    void add_property(std::string_view name, Value v);
    bool has_property(std::string_view name) const;

    // typed getters ("..._or" return the default when absent / wrong kind / too few values)
    std::optional<double> number(std::string_view name) const;
    std::optional<std::vector<double>> numbers(std::string_view name) const;
    std::optional<bool> boolean(std::string_view name) const;
    std::optional<std::string> string(std::string_view name) const;
    const NodeDesc *node(std::string_view name) const;// nullptr when absent
    std::vector<const NodeDesc *> nodes(std::string_view name) const;

    float f(std::string_view name, float dflt) const;
    uint32_t u(std::string_view name, uint32_t dflt) const;
    bool b(std::string_view name, bool dflt) const;
    std::string s(std::string_view name, const std::string &dflt) const;
    // a required file path, relative paths resolved against the directory of the defining file (property_path)
    std::filesystem::path path(std::string_view name) const;
    // N floats (N<=4); returns false if absent or fewer than N values
    bool fN(std::string_view name, int n, float *out) const;
    std::vector<float> float_list(std::string_view name) const;
    std::vector<uint32_t> uint_list(std::string_view name) const;
    const NodeDesc *required_node(std::string_view name) const;
    std::vector<const NodeDesc *> required_nodes(std::string_view name) const;

private:
    const Value *find(std::string_view name) const;
    std::string _identifier;
    std::string _impl;
    std::string _location;
    std::filesystem::path _source_dir;
    const NodeDesc *_base{nullptr};
    Tag _tag;
    std::vector<std::unique_ptr<NodeDesc>> _internal;
    std::map<std::string, Value, std::less<>> _props;
};

class SceneDesc {
public:
    static constexpr std::string_view root_identifier = "render";
    SceneDesc() : _root{std::string{root_identifier}, Tag::ROOT} {}
    const NodeDesc *root() const { return &_root; }
    const NodeDesc *reference(std::string_view identifier);
    NodeDesc *define(std::string_view identifier, Tag tag, std::string_view impl, std::string location,
                     std::filesystem::path dir, const NodeDesc *base);
    NodeDesc *define_root(std::string location, std::filesystem::path dir);
    const std::map<std::string, std::unique_ptr<NodeDesc>, std::less<>> &nodes() const { return _nodes; }

private:
    NodeDesc _root;
    bool _root_defined{false};
    std::map<std::string, std::unique_ptr<NodeDesc>, std::less<>> _nodes;
};

using MacroMap = std::map<std::string, std::string, std::less<>>;

// parse a .luisa text file or a .json file (by extension), following imports
std::unique_ptr<SceneDesc> parse_scene_file(const std::filesystem::path &path, const MacroMap &cli_macros);
// parse text source directly (used by tests); `dir` resolves relative imports / paths
std::unique_ptr<SceneDesc> parse_scene_source(std::string_view source, const std::filesystem::path &dir,
                                              const MacroMap &cli_macros, bool json = false);

}// namespace lrh
