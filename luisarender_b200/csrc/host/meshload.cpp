#include "meshload.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace lrh {

namespace {

[[noreturn]] void fail(const std::filesystem::path &p, const std::string &why) {
    throw std::runtime_error("Failed to load mesh '" + p.string() + "': " + why + ".");
}

struct P3 {
    float x, y, z;
};
struct P2 {
    float x, y;
};
inline P3 sub(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 cross(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline P3 normalized(P3 a) {
    float l = std::sqrt(dot(a, a));
    return l > 0.f ? P3{a.x / l, a.y / l, a.z / l} : P3{0.f, 0.f, 1.f};
}

// the raw content of a file: attribute arrays + polygon corners that index them (-1 = absent)
struct Corner {
    int32_t p, t, n;
};
struct RawMesh {
    std::vector<P3> positions, normals;
    std::vector<P2> uvs;
    std::vector<Corner> corners;       // 3 per triangle (already fan-triangulated)
};

void add_polygon(RawMesh &m, const std::vector<Corner> &poly) {
    for (size_t i = 1; i + 1 < poly.size(); i++) {
        m.corners.push_back(poly[0]);
        m.corners.push_back(poly[i]);
        m.corners.push_back(poly[i + 1]);
    }
}

// ---- Wavefront OBJ ---------------------------------------------------------------------------------------------------
RawMesh read_obj(const std::filesystem::path &path) {
    std::ifstream f{path};
    if (!f) fail(path, "cannot open file");
    RawMesh m;
    std::string line;
    std::vector<Corner> poly;
    size_t line_no = 0;
    while (std::getline(f, line)) {
        line_no++;
        while (!line.empty() && line.back() == '\\') {// continuation
            line.pop_back();
            std::string more;
            if (!std::getline(f, more)) break;
            line += more;
        }
        std::istringstream ss{line};
        std::string tag;
        if (!(ss >> tag) || tag[0] == '#') continue;
        if (tag == "v") {
            P3 p{};
            ss >> p.x >> p.y >> p.z;
            m.positions.push_back(p);
        } else if (tag == "vt") {
            P2 t{};
            ss >> t.x >> t.y;
            m.uvs.push_back(t);
        } else if (tag == "vn") {
            P3 n{};
            ss >> n.x >> n.y >> n.z;
            m.normals.push_back(n);
        } else if (tag == "f") {
            poly.clear();
            std::string item;
            while (ss >> item) {
                Corner c{-1, -1, -1};
                int32_t *slots[3] = {&c.p, &c.t, &c.n};
                size_t start = 0;
                for (int k = 0; k < 3 && start <= item.size(); k++) {
                    size_t slash = item.find('/', start);
                    std::string tok = item.substr(start, slash == std::string::npos ? std::string::npos : slash - start);
                    if (!tok.empty()) {
                        long idx = std::strtol(tok.c_str(), nullptr, 10);
                        long count = k == 0 ? static_cast<long>(m.positions.size()) : k == 1 ? static_cast<long>(m.uvs.size()) : static_cast<long>(m.normals.size());
                        long resolved = idx > 0 ? idx - 1 : count + idx;// negative = relative to the end
                        if (idx == 0 || resolved < 0 || resolved >= count) fail(path, "index out of range on line " + std::to_string(line_no));
                        *slots[k] = static_cast<int32_t>(resolved);
                    }
                    if (slash == std::string::npos) break;
                    start = slash + 1;
                }
                if (c.p < 0) fail(path, "face corner without a position on line " + std::to_string(line_no));
                poly.push_back(c);
            }
            if (poly.size() < 3) continue;// points / lines are removed (AI_CONFIG_PP_SBP_REMOVE, mesh.cpp:45-46)
            add_polygon(m, poly);
        }
        // o / g / s / usemtl / mtllib: all geometry is merged into one mesh (materials are stripped by the reference, :54)
    }
    return m;
}

// ---- Stanford PLY ------------------------------------------------------------------------------------------------------
size_t ply_type_size(const std::string &t, const std::filesystem::path &path) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    fail(path, "unknown PLY type '" + t + "'");
}
double ply_read_scalar(const uint8_t *p, const std::string &t, bool swap) {
    uint8_t b[8];
    size_t n = t == "double" || t == "float64" ? 8 : (t == "char" || t == "uchar" || t == "int8" || t == "uint8") ? 1 : (t == "short" || t == "ushort" || t == "int16" || t == "uint16") ? 2 : 4;
    std::memcpy(b, p, n);
    if (swap) std::reverse(b, b + n);
    if (t == "char" || t == "int8") return static_cast<int8_t>(b[0]);
    if (t == "uchar" || t == "uint8") return b[0];
    if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, b, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, b, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, b, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; std::memcpy(&v, b, 4); return v; }
    if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); return v; }
    double v;
    std::memcpy(&v, b, 8);
    return v;
}

RawMesh read_ply(const std::filesystem::path &path) {
    std::ifstream f{path, std::ios::binary};
    if (!f) fail(path, "cannot open file");
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    auto next_line = [&]() {
        std::string l;
        while (pos < d.size() && d[pos] != '\n') l.push_back(static_cast<char>(d[pos++]));
        pos++;
        if (!l.empty() && l.back() == '\r') l.pop_back();
        return l;
    };
    if (next_line() != "ply") fail(path, "not a PLY file");
    struct Prop {
        std::string name, type, count_type;// count_type non-empty = list
    };
    struct Element {
        std::string name;
        size_t count;
        std::vector<Prop> props;
    };
    std::vector<Element> elements;
    std::string format;
    for (;;) {
        if (pos >= d.size()) fail(path, "truncated PLY header");
        std::istringstream ss{next_line()};
        std::string tok;
        if (!(ss >> tok)) continue;
        if (tok == "end_header") break;
        if (tok == "format") ss >> format;
        else if (tok == "element") {
            Element e;
            ss >> e.name >> e.count;
            elements.push_back(e);
        } else if (tok == "property") {
            if (elements.empty()) fail(path, "PLY property outside an element");
            Prop p;
            std::string t;
            ss >> t;
            if (t == "list") ss >> p.count_type >> p.type >> p.name;
            else {
                p.type = t;
                ss >> p.name;
            }
            elements.back().props.push_back(p);
        }
    }
    const bool ascii = format == "ascii", swap = format == "binary_big_endian";
    if (!ascii && format != "binary_little_endian" && !swap) fail(path, "unknown PLY format '" + format + "'");
    RawMesh m;
    bool have_n = false, have_t = false;
    std::istringstream text;
    if (ascii) text.str(std::string(reinterpret_cast<const char *>(d.data() + pos), d.size() - pos));
    auto scalar = [&](const std::string &type) -> double {
        if (ascii) {
            double v;
            if (!(text >> v)) fail(path, "truncated PLY data");
            return v;
        }
        size_t n = ply_type_size(type, path);
        if (pos + n > d.size()) fail(path, "truncated PLY data");
        double v = ply_read_scalar(&d[pos], type, swap);
        pos += n;
        return v;
    };
    for (auto &e : elements) {
        if (e.name == "vertex") {
            for (auto &p : e.props) {
                if (p.name == "nx") have_n = true;
                if (p.name == "s" || p.name == "u" || p.name == "texture_u") have_t = true;
            }
            for (size_t i = 0; i < e.count; i++) {
                P3 p{}, n{};
                P2 t{};
                for (auto &pr : e.props) {
                    if (!pr.count_type.empty()) {
                        size_t c = static_cast<size_t>(scalar(pr.count_type));
                        for (size_t k = 0; k < c; k++) scalar(pr.type);
                        continue;
                    }
                    float v = static_cast<float>(scalar(pr.type));
                    if (pr.name == "x") p.x = v;
                    else if (pr.name == "y") p.y = v;
                    else if (pr.name == "z") p.z = v;
                    else if (pr.name == "nx") n.x = v;
                    else if (pr.name == "ny") n.y = v;
                    else if (pr.name == "nz") n.z = v;
                    else if (pr.name == "s" || pr.name == "u" || pr.name == "texture_u") t.x = v;
                    else if (pr.name == "t" || pr.name == "v" || pr.name == "texture_v") t.y = v;
                }
                m.positions.push_back(p);
                if (have_n) m.normals.push_back(n);
                if (have_t) m.uvs.push_back(t);
            }
        } else if (e.name == "face") {
            std::vector<Corner> poly;
            for (size_t i = 0; i < e.count; i++) {
                poly.clear();
                for (auto &pr : e.props) {
                    if (pr.count_type.empty()) {
                        scalar(pr.type);
                        continue;
                    }
                    size_t c = static_cast<size_t>(scalar(pr.count_type));
                    bool is_index = pr.name == "vertex_indices" || pr.name == "vertex_index";
                    for (size_t k = 0; k < c; k++) {
                        auto idx = static_cast<int64_t>(scalar(pr.type));
                        if (!is_index) continue;
                        if (idx < 0 || static_cast<size_t>(idx) >= m.positions.size()) fail(path, "PLY face index out of range");
                        auto ii = static_cast<int32_t>(idx);
                        poly.push_back({ii, have_t ? ii : -1, have_n ? ii : -1});
                    }
                }
                if (poly.size() >= 3) add_polygon(m, poly);
            }
        } else {// skip unknown elements
            for (size_t i = 0; i < e.count; i++)
                for (auto &pr : e.props) {
                    if (pr.count_type.empty()) scalar(pr.type);
                    else {
                        size_t c = static_cast<size_t>(scalar(pr.count_type));
                        for (size_t k = 0; k < c; k++) scalar(pr.type);
                    }
                }
        }
    }
    return m;
}

// smooth normals with a crease angle (assimp GenVertexNormalsProcess semantics: per face corner, the normalised sum of the
// unit face normals of all faces that share the POSITION and whose normal is within `max_angle` of this face's normal)
std::vector<P3> smooth_corner_normals(const RawMesh &m, float max_angle_degrees) {
    const size_t nt = m.corners.size() / 3u;
    std::vector<P3> face_n(nt);
    for (size_t t = 0; t < nt; t++) {
        P3 a = m.positions[m.corners[t * 3].p], b = m.positions[m.corners[t * 3 + 1].p], c = m.positions[m.corners[t * 3 + 2].p];
        face_n[t] = normalized(cross(sub(b, a), sub(c, a)));
    }
    // faces around each position; identical coordinates count as one position (assimp's spatial sort)
    std::map<std::tuple<float, float, float>, uint32_t> by_coord;
    std::vector<uint32_t> pos_class(m.positions.size());
    for (size_t i = 0; i < m.positions.size(); i++) {
        auto key = std::make_tuple(m.positions[i].x, m.positions[i].y, m.positions[i].z);
        pos_class[i] = by_coord.emplace(key, static_cast<uint32_t>(by_coord.size())).first->second;
    }
    std::vector<std::vector<uint32_t>> faces_of(by_coord.size());
    for (size_t t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++) {
            auto &v = faces_of[pos_class[m.corners[t * 3 + k].p]];
            if (v.empty() || v.back() != t) v.push_back(static_cast<uint32_t>(t));
        }
    const float limit = std::cos(max_angle_degrees * 3.14159265358979323846f / 180.f);
    std::vector<P3> out(m.corners.size());
    for (size_t t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++) {
            P3 sum{0.f, 0.f, 0.f};
            for (auto other : faces_of[pos_class[m.corners[t * 3 + k].p]]) {
                if (dot(face_n[other], face_n[t]) < limit) continue;
                sum.x += face_n[other].x;
                sum.y += face_n[other].y;
                sum.z += face_n[other].z;
            }
            out[t * 3 + k] = normalized(sum);
        }
    return out;
}

}// namespace

MeshData load_mesh(const std::filesystem::path &path, bool flip_uv, bool drop_normal, bool drop_uv) {
    auto ext = path.extension().string();
    for (auto &c : ext) c = static_cast<char>(std::tolower(c));
    RawMesh raw;
    if (ext == ".obj") raw = read_obj(path);
    else if (ext == ".ply") raw = read_ply(path);
    else fail(path, "unsupported mesh format '" + ext + "' (supported: .obj .ply)");
    if (raw.corners.empty()) fail(path, "no triangles");
    bool use_uv = !drop_uv && !raw.uvs.empty();
    for (auto &c : raw.corners)
        if (use_uv && c.t < 0) use_uv = false;// mixed files: assimp drops the channel
    bool file_normals = !drop_normal && !raw.normals.empty();
    for (auto &c : raw.corners)
        if (file_normals && c.n < 0) file_normals = false;
    std::vector<P3> generated;
    if (!drop_normal && !file_normals) generated = smooth_corner_normals(raw, 45.f);
    const bool use_n = !drop_normal;
    MeshData out;
    out.properties = (use_n ? LRK_SHAPE_HAS_VERTEX_NORMAL : 0u) | (use_uv ? LRK_SHAPE_HAS_VERTEX_UV : 0u);
    // join identical vertices: bit-identical (position, normal, uv) triples share one vertex
    struct Key {
        float v[8];
        bool operator==(const Key &o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
    };
    struct Hash {
        size_t operator()(const Key &k) const {
            uint64_t h = 1469598103934665603ull;
            const auto *b = reinterpret_cast<const uint8_t *>(k.v);
            for (size_t i = 0; i < sizeof(k.v); i++) h = (h ^ b[i]) * 1099511628211ull;
            return static_cast<size_t>(h);
        }
    };
    std::unordered_map<Key, uint32_t, Hash> lookup;
    out.triangles.resize(raw.corners.size() / 3u);
    for (size_t i = 0; i < raw.corners.size(); i++) {
        const Corner &c = raw.corners[i];
        P3 p = raw.positions[c.p];
        P3 n = !use_n ? P3{0.f, 0.f, 1.f} : file_normals ? normalized(raw.normals[c.n]) : generated[i];
        P2 t = use_uv ? raw.uvs[c.t] : P2{0.f, 0.f};
        if (use_uv && !flip_uv) t.y = 1.f - t.y;
        Key k{{p.x, p.y, p.z, n.x, n.y, n.z, t.x, t.y}};
        auto [it, inserted] = lookup.emplace(k, static_cast<uint32_t>(out.vertices.size()));
        if (inserted) {
            lrk_vertex v{};
            v.p[0] = p.x, v.p[1] = p.y, v.p[2] = p.z;
            v.n[0] = n.x, v.n[1] = n.y, v.n[2] = n.z;
            v.uv[0] = t.x, v.uv[1] = t.y;
            out.vertices.push_back(v);
        }
        uint32_t *slot = i % 3u == 0u ? &out.triangles[i / 3u].i0 : i % 3u == 1u ? &out.triangles[i / 3u].i1 : &out.triangles[i / 3u].i2;
        *slot = it->second;
    }
    return out;
}

}// namespace lrh
