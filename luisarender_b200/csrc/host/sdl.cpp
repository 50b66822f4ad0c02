#include "sdl.h"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace lrh {

namespace {
std::string lower(std::string_view s) {
    std::string r{s};
    for (auto &c : r) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    return r;
}
}// namespace

std::string_view tag_description(Tag tag) {
    switch (tag) {
        case Tag::ROOT: return "__root__";
        case Tag::INTERNAL: return "__internal__";
        case Tag::DECLARATION: return "__declaration__";
        case Tag::CAMERA: return "camera";
        case Tag::SHAPE: return "shape";
        case Tag::SURFACE: return "surface";
        case Tag::LIGHT: return "light";
        case Tag::TRANSFORM: return "transform";
        case Tag::FILM: return "film";
        case Tag::FILTER: return "filter";
        case Tag::SAMPLER: return "sampler";
        case Tag::INTEGRATOR: return "integrator";
        case Tag::LIGHT_SAMPLER: return "lightsampler";
        case Tag::ENVIRONMENT: return "environment";
        case Tag::TEXTURE: return "texture";
        case Tag::TEXTURE_MAPPING: return "texturemapping";
        case Tag::SPECTRUM: return "spectrum";
        case Tag::MEDIUM: return "medium";
        case Tag::PHASE_FUNCTION: return "phasefunction";
    }
    return "__invalid__";
}

Tag parse_tag(std::string_view s) {
    // aliases: src/sdl/scene_node_tag.cpp:15-46
    static const std::unordered_map<std::string, Tag> table{
        {"camera", Tag::CAMERA}, {"cam", Tag::CAMERA},
        {"shape", Tag::SHAPE}, {"object", Tag::SHAPE}, {"obj", Tag::SHAPE},
        {"surface", Tag::SURFACE}, {"surf", Tag::SURFACE},
        {"lightsource", Tag::LIGHT}, {"light", Tag::LIGHT}, {"illuminant", Tag::LIGHT}, {"illum", Tag::LIGHT},
        {"transform", Tag::TRANSFORM}, {"xform", Tag::TRANSFORM},
        {"film", Tag::FILM}, {"filter", Tag::FILTER}, {"sampler", Tag::SAMPLER},
        {"integrator", Tag::INTEGRATOR}, {"lightsampler", Tag::LIGHT_SAMPLER},
        {"environment", Tag::ENVIRONMENT}, {"env", Tag::ENVIRONMENT},
        {"texture", Tag::TEXTURE}, {"tex", Tag::TEXTURE},
        {"texturemapping", Tag::TEXTURE_MAPPING}, {"texmapping", Tag::TEXTURE_MAPPING},
        {"spectrum", Tag::SPECTRUM}, {"spec", Tag::SPECTRUM},
        {"generic", Tag::DECLARATION}, {"template", Tag::DECLARATION},
        {"medium", Tag::MEDIUM}, {"phasefunction", Tag::PHASE_FUNCTION}};
    auto it = table.find(lower(s));
    return it == table.end() ? Tag::ROOT : it->second;
}

// ---------------------------------------------------------------- NodeDesc

void NodeDesc::define(Tag tag, std::string_view impl, std::string location, std::filesystem::path dir,
                      const NodeDesc *base) {
    _tag = tag;
    _impl = lower(impl);
    _location = std::move(location);
    _source_dir = std::move(dir);
    _base = base;
}

NodeDesc *NodeDesc::define_internal(std::string_view impl, std::string location, std::filesystem::path dir,
                                    const NodeDesc *base) {
    auto id = _identifier + ".$internal" + std::to_string(_internal.size());
    auto &n = _internal.emplace_back(std::make_unique<NodeDesc>(std::move(id), Tag::INTERNAL));
    n->define(Tag::INTERNAL, impl, std::move(location), std::move(dir), base);
    return n.get();
}

void NodeDesc::add_property(std::string_view name, Value v) {
    if (!_props.emplace(std::string{name}, std::move(v)).second) {
        throw Error("Redefinition of property '" + std::string{name} + "' in scene description node '" +
                    _identifier + "'. [" + _location + "]");
    }
}

const NodeDesc::Value *NodeDesc::find(std::string_view name) const {
    if (auto it = _props.find(name); it != _props.end()) return &it->second;
    return _base ? _base->find(name) : nullptr;
}

bool NodeDesc::has_property(std::string_view name) const { return find(name) != nullptr; }

std::optional<std::vector<double>> NodeDesc::numbers(std::string_view name) const {
    auto v = find(name);
    if (!v) return std::nullopt;
    if (auto p = std::get_if<NumberList>(v)) return *p;
    return std::nullopt;
}
std::optional<double> NodeDesc::number(std::string_view name) const {
    auto v = numbers(name);
    if (!v || v->empty()) return std::nullopt;
    return v->front();
}
std::optional<bool> NodeDesc::boolean(std::string_view name) const {
    auto v = find(name);
    if (!v) return std::nullopt;
    if (auto p = std::get_if<BoolList>(v); p && !p->empty()) return (*p)[0];
    return std::nullopt;
}
std::optional<std::string> NodeDesc::string(std::string_view name) const {
    auto v = find(name);
    if (!v) return std::nullopt;
    if (auto p = std::get_if<StringList>(v); p && !p->empty()) return (*p)[0];
    return std::nullopt;
}
const NodeDesc *NodeDesc::node(std::string_view name) const {
    auto v = find(name);
    if (!v) return nullptr;
    if (auto p = std::get_if<NodeList>(v); p && !p->empty()) return (*p)[0];
    return nullptr;
}
std::vector<const NodeDesc *> NodeDesc::nodes(std::string_view name) const {
    auto v = find(name);
    if (!v) return {};
    if (auto p = std::get_if<NodeList>(v)) return *p;
    return {};
}
float NodeDesc::f(std::string_view name, float dflt) const {
    auto v = number(name);
    return v ? static_cast<float>(*v) : dflt;
}
uint32_t NodeDesc::u(std::string_view name, uint32_t dflt) const {
    auto v = number(name);
    if (!v) return dflt;
    auto r = static_cast<uint32_t>(*v);
    if (static_cast<double>(r) != *v) {
        throw Error("Cannot convert property '" + std::string{name} + "' to integer in scene description node '" +
                    _identifier + "'. [" + _location + "]");
    }
    return r;
}
bool NodeDesc::b(std::string_view name, bool dflt) const {
    auto v = boolean(name);
    return v ? *v : dflt;
}

std::string NodeDesc::s(std::string_view name, const std::string &dflt) const {
    auto v = string(name);
    return v ? *v : dflt;
}

std::filesystem::path NodeDesc::path(std::string_view name) const {
    auto v = string(name);
    if (!v) throw Error("No valid values given for property '" + std::string{name} + "' in scene description node '" + _identifier + "'. [" + _location + "]");
    std::filesystem::path p{*v};
    return p.is_absolute() ? p : _source_dir / p;
}
bool NodeDesc::fN(std::string_view name, int n, float *out) const {
    auto v = numbers(name);
    if (!v || static_cast<int>(v->size()) < n) return false;
    for (int i = 0; i < n; i++) out[i] = static_cast<float>((*v)[i]);
    return true;
}
std::vector<float> NodeDesc::float_list(std::string_view name) const {
    std::vector<float> r;
    if (auto v = numbers(name)) {
        r.reserve(v->size());
        for (auto x : *v) r.push_back(static_cast<float>(x));
    }
    return r;
}
std::vector<uint32_t> NodeDesc::uint_list(std::string_view name) const {
    std::vector<uint32_t> r;
    if (auto v = numbers(name)) {
        r.reserve(v->size());
        for (auto x : *v) {
            auto q = static_cast<uint32_t>(x);
            if (static_cast<double>(q) != x) {
                throw Error("Cannot convert property '" + std::string{name} + "' to integer in node '" + _identifier + "'.");
            }
            r.push_back(q);
        }
    }
    return r;
}
const NodeDesc *NodeDesc::required_node(std::string_view name) const {
    if (auto n = node(name)) return n;
    throw Error("No valid values given for property '" + std::string{name} + "' in scene description node '" +
                _identifier + "'. [" + _location + "]");
}
std::vector<const NodeDesc *> NodeDesc::required_nodes(std::string_view name) const {
    auto v = find(name);
    if (v) {
        if (auto p = std::get_if<NodeList>(v)) return *p;
    }
    throw Error("No valid values given for property '" + std::string{name} + "' in scene description node '" +
                _identifier + "'. [" + _location + "]");
}

// ---------------------------------------------------------------- SceneDesc

const NodeDesc *SceneDesc::reference(std::string_view identifier) {
    if (identifier == root_identifier) throw Error("Invalid reference to root node.");
    auto it = _nodes.find(identifier);
    if (it == _nodes.end()) {
        it = _nodes.emplace(std::string{identifier},
                            std::make_unique<NodeDesc>(std::string{identifier}, Tag::DECLARATION)).first;
    }
    return it->second.get();
}

NodeDesc *SceneDesc::define(std::string_view identifier, Tag tag, std::string_view impl, std::string location,
                            std::filesystem::path dir, const NodeDesc *base) {
    if (identifier == root_identifier || tag == Tag::ROOT)
        throw Error("Defining root node as a normal global node is not allowed. [" + location + "]");
    if (tag == Tag::INTERNAL || tag == Tag::DECLARATION)
        throw Error("Defining internal or declaration node as a global node is not allowed. [" + location + "]");
    auto it = _nodes.find(identifier);
    if (it == _nodes.end()) {
        it = _nodes.emplace(std::string{identifier}, std::make_unique<NodeDesc>(std::string{identifier}, tag)).first;
    }
    auto node = it->second.get();
    if (node->is_defined()) {
        throw Error("Redefinition of node '" + node->identifier() + "' in scene description. [" + location + "]");
    }
    node->define(tag, impl, std::move(location), std::move(dir), base);
    return node;
}

NodeDesc *SceneDesc::define_root(std::string location, std::filesystem::path dir) {
    if (_root_defined) throw Error("Redefinition of root node in scene description. [" + location + "]");
    _root_defined = true;
    _root.define(Tag::ROOT, root_identifier, std::move(location), std::move(dir));
    return &_root;
}

// ---------------------------------------------------------------- text parser

namespace {

void parse_any_file(SceneDesc &desc, const std::filesystem::path &path, const MacroMap &cli);

class TextParser {
public:
    TextParser(SceneDesc &desc, std::string source, std::string file, std::filesystem::path dir, const MacroMap &cli)
        : _desc{desc}, _src{std::move(source)}, _file{std::move(file)}, _dir{std::move(dir)}, _cli{cli} {}

    void run() {
        skip_blanks();
        while (!eof()) {
            auto loc = location();
            auto token = read_identifier();
            if (token == "import") {
                skip_blanks();
                std::filesystem::path p{read_string()};
                if (!p.is_absolute()) p = _dir / p;
                parse_any_file(_desc, p, _cli);
            } else if (token == "define") {
                parse_define();
            } else if (token == SceneDesc::root_identifier) {
                parse_node_body(_desc.define_root(loc, _dir));
            } else {
                parse_global_node(loc, token);
            }
            skip_blanks();
        }
    }

private:
    SceneDesc &_desc;
    std::string _src;
    std::string _file;
    std::filesystem::path _dir;
    const MacroMap &_cli;
    MacroMap _local;
    std::vector<std::string> _macro_stack;// pending macro expansions, innermost last
    size_t _cursor{0};
    uint32_t _line{0}, _col{0};

    std::string location() const { return _file + ":" + std::to_string(_line + 1) + ":" + std::to_string(_col); }
    [[noreturn]] void fail(const std::string &msg) const { throw Error(msg + " [" + location() + "]"); }
    bool eof() const { return _macro_stack.empty() && _cursor >= _src.size(); }

    char raw_peek() {
        if (!_macro_stack.empty()) return _macro_stack.back().front();
        if (_cursor >= _src.size()) fail("Premature EOF.");
        auto c = _src[_cursor];
        if (c == '\r') return '\n';
        return c;
    }
    char raw_get() {
        if (!_macro_stack.empty()) {
            auto &m = _macro_stack.back();
            auto c = m.front();
            if (m.size() > 1) m.erase(0, 1);
            else _macro_stack.pop_back();
            return c;
        }
        if (_cursor >= _src.size()) fail("Premature EOF.");
        auto c = _src[_cursor++];
        if (c == '\r') {
            if (_cursor < _src.size() && _src[_cursor] == '\n') _cursor++;
            c = '\n';
        }
        if (c == '\n') { _line++; _col = 0; } else { _col++; }
        return c;
    }
    // `escape_macro` = do not expand '#'
    char peek(bool escape_macro = false) {
        auto c = raw_peek();
        if (!escape_macro) {
            while (c == '#') {
                raw_get();
                parse_macro();
                c = raw_peek();
            }
        }
        return c;
    }
    char get(bool escape_macro = false) {
        auto c = raw_get();
        if (!escape_macro) {
            while (c == '#') {
                parse_macro();
                c = raw_get();
            }
        }
        return c;
    }
    void match(char c) {
        auto got = get();
        if (got != c) fail(std::string{"Invalid character '"} + got + "' (expected '" + c + "').");
    }
    void skip_blanks() {
        while (!eof()) {
            auto c = peek(true);
            if (c == ' ' || c == '\t' || c == '\n') {
                raw_get();
            } else if (c == '/') {
                raw_get();
                if (raw_get() != '/') fail("Invalid character (expected '/').");
                while (!eof() && raw_get() != '\n') {}
            } else {
                break;
            }
        }
    }
    std::string read_identifier(bool escape_macro = false) {
        std::string id;
        auto c = get(escape_macro);
        if (c != '$' && c != '_' && !std::isalpha(static_cast<unsigned char>(c)))
            fail(std::string{"Invalid character '"} + c + "' in identifier.");
        id.push_back(c);
        auto body = [](char ch) {
            return std::isalnum(static_cast<unsigned char>(ch)) || ch == '_' || ch == '$' || ch == '-';
        };
        while (!eof() && body(peek(escape_macro))) id.push_back(get(escape_macro));
        return id;
    }
    double read_number() {
        std::string s;
        if (auto c = peek(); c == '+') {
            get();
            skip_blanks();
        } else if (c == '-') {
            s.push_back(get());
            skip_blanks();
        }
        auto is_num = [](char ch) {
            return std::isdigit(static_cast<unsigned char>(ch)) || ch == '.' || ch == 'e' || ch == '-' || ch == '+';
        };
        while (!eof() && is_num(peek())) s.push_back(get());
        double value = 0.0;
        auto res = std::from_chars(s.data(), s.data() + s.size(), value);
        if (res.ec != std::errc{} || s.empty()) fail("Invalid number string '" + s.substr(0, 4) + "...'.");
        return value;
    }
    bool read_bool() {
        if (peek() == 't') {
            for (char x : std::string_view{"true"}) match(x);
            return true;
        }
        for (char x : std::string_view{"false"}) match(x);
        return false;
    }
    std::string read_string() {
        auto quote = get();
        if (quote != '"' && quote != '\'') fail(std::string{"Expected string but got "} + quote + ".");
        std::string s;
        for (auto c = get(); c != quote; c = get()) {
            if (!std::isprint(static_cast<unsigned char>(c))) fail("Unexpected non-printable character.");
            if (c == '\\') {
                auto esc = get(true);
                switch (esc) {
                    case 'b': c = '\b'; break;
                    case 'f': c = '\f'; break;
                    case 'n': c = '\n'; break;
                    case 'r': c = '\r'; break;
                    case 't': c = '\t'; break;
                    case '\\': c = '\\'; break;
                    case '\'': c = '\''; break;
                    case '"': c = '"'; break;
                    case '#': c = '#'; break;
                    default: fail(std::string{"Invalid escaped character '"} + esc + "'.");
                }
            }
            s.push_back(c);
        }
        return s;
    }
    void parse_macro() {
        skip_blanks();
        auto key = read_identifier(true);
        if (auto it = _cli.find(key); it != _cli.end()) {
            if (!it->second.empty()) _macro_stack.push_back(it->second);
        } else if (auto lt = _local.find(key); lt != _local.end()) {
            if (!lt->second.empty()) _macro_stack.push_back(lt->second);
        } else {
            fail("Undefined macro '" + key + "'.");
        }
    }
    void parse_define() {
        skip_blanks();
        auto key = read_identifier(true);
        skip_blanks();
        std::string value;
        while (!eof() && peek(true) != '\n' && peek(true) != '/') value.push_back(get(true));
        if (!_cli.count(key)) _local[key] = value;
    }
    const NodeDesc *parse_base() {
        match('(');
        skip_blanks();
        match('@');
        skip_blanks();
        auto base = _desc.reference(read_identifier());
        skip_blanks();
        match(')');
        return base;
    }
    void parse_global_node(const std::string &loc, const std::string &tag_desc) {
        auto tag = parse_tag(tag_desc);
        if (tag == Tag::ROOT) fail("Invalid scene node type '" + tag_desc + "'.");
        skip_blanks();
        auto name = read_identifier();
        skip_blanks();
        const NodeDesc *base = nullptr;
        std::string impl;
        if (peek() == ':') {
            match(':');
            skip_blanks();
            impl = read_identifier();
            skip_blanks();
            if (peek() == '(') base = parse_base();
            skip_blanks();
        }
        parse_node_body(_desc.define(name, tag, impl, loc, _dir, base));
    }
    void parse_node_body(NodeDesc *node) {
        skip_blanks();
        match('{');
        skip_blanks();
        while (peek() != '}') {
            auto prop = read_identifier();
            skip_blanks();
            if (peek() == ':') {// inline node
                get();
                skip_blanks();
                auto loc = location();
                auto impl = read_identifier();
                const NodeDesc *base = nullptr;
                if (peek() == '(') base = parse_base();
                auto internal = node->define_internal(impl, loc, _dir, base);
                parse_node_body(internal);
                node->add_property(prop, NodeDesc::NodeList{internal});
            } else {
                node->add_property(prop, parse_value_list(node));
            }
            skip_blanks();
        }
        match('}');
    }
    template<typename T, typename F>
    std::vector<T> parse_list(F &&read_one) {
        std::vector<T> list;
        list.emplace_back(read_one());
        skip_blanks();
        while (peek() != '}') {
            match(',');
            skip_blanks();
            list.emplace_back(read_one());
            skip_blanks();
        }
        return list;
    }
    NodeDesc::Value parse_value_list(NodeDesc *node) {
        match('{');
        skip_blanks();
        NodeDesc::Value value;
        auto c = peek();
        if (c == '}') fail("Empty value list.");
        if (c == '@' || std::isupper(static_cast<unsigned char>(c))) {
            value = parse_list<const NodeDesc *>([&]() -> const NodeDesc * {
                if (peek() == '@') {
                    get();
                    skip_blanks();
                    return _desc.reference(read_identifier());
                }
                auto loc = location();
                auto impl = read_identifier();
                const NodeDesc *base = nullptr;
                if (peek() == '(') base = parse_base();
                auto internal = node->define_internal(impl, loc, _dir, base);
                parse_node_body(internal);
                return internal;
            });
        } else if (c == '"' || c == '\'') {
            value = parse_list<std::string>([&] { return read_string(); });
        } else if (c == 't' || c == 'f') {
            value = parse_list<bool>([&] { return read_bool(); });
        } else {
            value = parse_list<double>([&] { return read_number(); });
        }
        skip_blanks();
        match('}');
        return value;
    }
};

// ---------------------------------------------------------------- tiny JSON reader (comments allowed)

struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind{Null};
    bool b{};
    double n{};
    std::string s;
    std::vector<Json> a;
    std::vector<std::pair<std::string, Json>> o;// insertion order kept
    const Json *find(std::string_view k) const {
        for (auto &kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

class JsonReader {
public:
    explicit JsonReader(std::string_view src) : _s{src} {}
    Json parse() {
        auto v = value();
        ws();
        if (_p != _s.size()) fail("trailing characters");
        return v;
    }

private:
    std::string_view _s;
    size_t _p{0};
    [[noreturn]] void fail(const std::string &m) const { throw Error("JSON parse error at offset " + std::to_string(_p) + ": " + m); }
    void ws() {
        for (;;) {
            while (_p < _s.size() && std::isspace(static_cast<unsigned char>(_s[_p]))) _p++;
            if (_p + 1 < _s.size() && _s[_p] == '/' && _s[_p + 1] == '/') {
                while (_p < _s.size() && _s[_p] != '\n') _p++;
            } else if (_p + 1 < _s.size() && _s[_p] == '/' && _s[_p + 1] == '*') {
                _p += 2;
                while (_p + 1 < _s.size() && !(_s[_p] == '*' && _s[_p + 1] == '/')) _p++;
                _p += 2;
            } else {
                break;
            }
        }
    }
    Json value() {
        ws();
        if (_p >= _s.size()) fail("unexpected end");
        Json j;
        char c = _s[_p];
        if (c == '{') {
            _p++;
            j.kind = Json::Object;
            ws();
            if (_p < _s.size() && _s[_p] == '}') { _p++; return j; }
            for (;;) {
                ws();
                auto k = str();
                ws();
                if (_p >= _s.size() || _s[_p] != ':') fail("expected ':'");
                _p++;
                j.o.emplace_back(std::move(k), value());
                ws();
                if (_p < _s.size() && _s[_p] == ',') { _p++; continue; }
                if (_p < _s.size() && _s[_p] == '}') { _p++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            _p++;
            j.kind = Json::Array;
            ws();
            if (_p < _s.size() && _s[_p] == ']') { _p++; return j; }
            for (;;) {
                j.a.emplace_back(value());
                ws();
                if (_p < _s.size() && _s[_p] == ',') { _p++; continue; }
                if (_p < _s.size() && _s[_p] == ']') { _p++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            j.kind = Json::String;
            j.s = str();
        } else if (_s.compare(_p, 4, "true") == 0) {
            j.kind = Json::Bool; j.b = true; _p += 4;
        } else if (_s.compare(_p, 5, "false") == 0) {
            j.kind = Json::Bool; j.b = false; _p += 5;
        } else if (_s.compare(_p, 4, "null") == 0) {
            j.kind = Json::Null; _p += 4;
        } else {
            j.kind = Json::Number;
            auto begin = _p;
            while (_p < _s.size() && (std::isdigit(static_cast<unsigned char>(_s[_p])) || _s[_p] == '-' || _s[_p] == '+' ||
                                      _s[_p] == '.' || _s[_p] == 'e' || _s[_p] == 'E')) _p++;
            auto res = std::from_chars(_s.data() + begin, _s.data() + _p, j.n);
            if (res.ec != std::errc{} || begin == _p) fail("invalid number");
        }
        return j;
    }
    std::string str() {
        if (_p >= _s.size() || _s[_p] != '"') fail("expected string");
        _p++;
        std::string r;
        while (_p < _s.size() && _s[_p] != '"') {
            char c = _s[_p++];
            if (c == '\\') {
                if (_p >= _s.size()) fail("bad escape");
                char e = _s[_p++];
                switch (e) {
                    case 'n': r.push_back('\n'); break;
                    case 't': r.push_back('\t'); break;
                    case 'r': r.push_back('\r'); break;
                    case 'b': r.push_back('\b'); break;
                    case 'f': r.push_back('\f'); break;
                    case 'u': {// keep BMP code points as UTF-8
                        if (_p + 4 > _s.size()) fail("bad \\u escape");
                        unsigned cp = std::strtoul(std::string{_s.substr(_p, 4)}.c_str(), nullptr, 16);
                        _p += 4;
                        if (cp < 0x80) r.push_back(static_cast<char>(cp));
                        else if (cp < 0x800) { r.push_back(static_cast<char>(0xc0 | (cp >> 6))); r.push_back(static_cast<char>(0x80 | (cp & 0x3f))); }
                        else { r.push_back(static_cast<char>(0xe0 | (cp >> 12))); r.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3f))); r.push_back(static_cast<char>(0x80 | (cp & 0x3f))); }
                        break;
                    }
                    default: r.push_back(e);
                }
            } else {
                r.push_back(c);
            }
        }
        if (_p >= _s.size()) fail("unterminated string");
        _p++;
        return r;
    }
};

class JsonSceneParser {
public:
    JsonSceneParser(SceneDesc &desc, std::string file, std::filesystem::path dir, const MacroMap &cli)
        : _desc{desc}, _file{std::move(file)}, _dir{std::move(dir)}, _cli{cli} {}

    void run(std::string_view src) {
        auto root = JsonReader{src}.parse();
        if (root.kind != Json::Object) throw Error("Invalid JSON scene (top level must be an object). [" + _file + "]");
        if (auto imp = root.find("import")) {
            auto one = [&](const Json &j) {
                if (j.kind != Json::String) throw Error("Invalid import node. [" + _file + "]");
                std::filesystem::path p{j.s};
                if (!p.is_absolute()) p = _dir / p;
                parse_any_file(_desc, p, _cli);
            };
            if (imp->kind == Json::Array) for (auto &j : imp->a) one(j);
            else one(*imp);
        }
        for (auto &[key, val] : root.o) {
            if (key == "import") continue;
            if (val.kind != Json::Object) throw Error("Invalid global node '" + key + "'. [" + _file + "]");
            if (key == SceneDesc::root_identifier) {
                parse_node(*_desc.define_root(_file, _dir), val);
                continue;
            }
            check_keys(key, val);
            auto type = val.find("type");
            if (!type || type->kind != Json::String) throw Error("Missing node type in global node '" + key + "'. [" + _file + "]");
            auto tag = parse_tag(type->s);
            if (tag == Tag::ROOT) throw Error("Unknown scene node type: " + type->s + " [" + _file + "]");
            auto impl = val.find("impl");
            if (!impl || impl->kind != Json::String) throw Error("Missing impl in global node '" + key + "'. [" + _file + "]");
            const NodeDesc *base = nullptr;
            if (auto b = val.find("base")) base = ref(b->s);
            auto node = _desc.define(key, tag, impl->s, _file, _dir, base);
            if (auto prop = val.find("prop")) parse_node(*node, *prop);
        }
    }

private:
    SceneDesc &_desc;
    std::string _file;
    std::filesystem::path _dir;
    const MacroMap &_cli;

    void check_keys(const std::string &name, const Json &n) const {
        for (auto &kv : n.o) {
            if (kv.first != "type" && kv.first != "impl" && kv.first != "base" && kv.first != "prop")
                throw Error("Invalid node property '" + name + "." + kv.first + "'. [" + _file + "]");
        }
    }
    const NodeDesc *ref(const std::string &name) const {
        if (name.empty() || name[0] != '@') throw Error("Invalid reference name '" + name + "'. [" + _file + "]");
        return _desc.reference(std::string_view{name}.substr(1));
    }
    const NodeDesc *internal(NodeDesc &desc, const std::string &key, const Json &n) {
        if (n.kind != Json::Object) throw Error("Invalid node reference in '" + desc.identifier() + "'.'" + key + "'. [" + _file + "]");
        check_keys(key, n);
        auto impl = n.find("impl");
        if (!impl || impl->kind != Json::String) throw Error("Missing impl in internal node '" + key + "'. [" + _file + "]");
        const NodeDesc *base = nullptr;
        if (auto b = n.find("base")) base = ref(b->s);
        auto node = desc.define_internal(impl->s, _file, _dir, base);
        if (auto prop = n.find("prop")) parse_node(*node, *prop);
        return node;
    }
    void parse_node(NodeDesc &desc, const Json &node) {
        if (node.kind != Json::Object) throw Error("Invalid property object in '" + desc.identifier() + "'. [" + _file + "]");
        for (auto &[key, v] : node.o) {
            switch (v.kind) {
                case Json::String:
                    if (!v.s.empty() && v.s[0] == '@') desc.add_property(key, NodeDesc::NodeList{ref(v.s)});
                    else desc.add_property(key, NodeDesc::StringList{v.s});
                    break;
                case Json::Number: desc.add_property(key, NodeDesc::NumberList{v.n}); break;
                case Json::Bool: desc.add_property(key, NodeDesc::BoolList{v.b}); break;
                case Json::Array: {
                    if (v.a.empty()) throw Error("Empty array is not allowed in '" + desc.identifier() + "'.'" + key + "'. [" + _file + "]");
                    auto &first = v.a[0];
                    if (first.kind == Json::String && !(first.s.size() && first.s[0] == '@')) {
                        NodeDesc::StringList l;
                        for (auto &e : v.a) l.push_back(e.s);
                        desc.add_property(key, std::move(l));
                    } else if (first.kind == Json::Number) {
                        NodeDesc::NumberList l;
                        for (auto &e : v.a) l.push_back(e.n);
                        desc.add_property(key, std::move(l));
                    } else if (first.kind == Json::Bool) {
                        NodeDesc::BoolList l;
                        for (auto &e : v.a) l.push_back(e.b);
                        desc.add_property(key, std::move(l));
                    } else {
                        NodeDesc::NodeList l;
                        for (auto &e : v.a) l.push_back(e.kind == Json::String ? ref(e.s) : internal(desc, key, e));
                        desc.add_property(key, std::move(l));
                    }
                    break;
                }
                case Json::Object: desc.add_property(key, NodeDesc::NodeList{internal(desc, key, v)}); break;
                case Json::Null: break;
            }
        }
    }
};

std::string read_file(const std::filesystem::path &path) {
    std::ifstream f{path, std::ios::binary};
    if (!f) throw Error("Failed to open file '" + path.string() + "'.");
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

void parse_any_file(SceneDesc &desc, const std::filesystem::path &path_in, const MacroMap &cli) {
    std::error_code ec;
    auto path = std::filesystem::canonical(path_in, ec);
    if (ec) throw Error("Failed to open file '" + path_in.string() + "'.");
    auto src = read_file(path);
    if (lower(path.extension().string()) == ".json") {
        JsonSceneParser{desc, path.string(), path.parent_path(), cli}.run(src);
    } else {
        TextParser{desc, std::move(src), path.string(), path.parent_path(), cli}.run();
    }
}

}// namespace

std::unique_ptr<SceneDesc> parse_scene_file(const std::filesystem::path &path, const MacroMap &cli_macros) {
    auto desc = std::make_unique<SceneDesc>();
    parse_any_file(*desc, path, cli_macros);
    return desc;
}

std::unique_ptr<SceneDesc> parse_scene_source(std::string_view source, const std::filesystem::path &dir,
                                              const MacroMap &cli_macros, bool json) {
    auto desc = std::make_unique<SceneDesc>();
    if (json) JsonSceneParser{*desc, "<memory>", dir, cli_macros}.run(source);
    else TextParser{*desc, std::string{source}, "<memory>", dir, cli_macros}.run();
    return desc;
}

}// namespace lrh
