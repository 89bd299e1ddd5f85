// host/project.cpp — see project.hpp. File:line citations are relative to /root/reference/bisemutum/.
#include "project.hpp"
#include <cerrno>
#include "gltf.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace bi::project {

// =====================================================================================================================
// TOML subset
// =====================================================================================================================
auto Toml::find(std::string_view key) const -> Toml const* {
    if (kind != Kind::table) return nullptr;
    for (auto& kv : tab) if (kv.first == key) return &kv.second;
    return nullptr;
}
auto Toml::at(std::string_view path) const -> Toml const* {
    Toml const* cur = this;
    while (cur && !path.empty()) {
        auto dot = path.find('.');
        cur = cur->find(path.substr(0, dot));
        path = dot == std::string_view::npos ? std::string_view{} : path.substr(dot + 1);
    }
    return cur;
}
auto Toml::number_or(std::string_view path, double dflt) const -> double { auto v = at(path); return v && v->kind == Kind::number ? v->num : dflt; }
auto Toml::id_or(std::string_view path, uint64_t dflt) const -> uint64_t {
    auto v = at(path);
    if (!v || v->kind != Kind::number) return dflt;
    if (v->is_int) return v->u64;
    return (v->num >= 0.0 && v->num < 18446744073709551616.0) ? (uint64_t)v->num : dflt;      // (no cast of a negative / out-of-range double: undefined)
}
auto Toml::string_or(std::string_view path, std::string dflt) const -> std::string { auto v = at(path); return v && v->kind == Kind::string ? v->str : dflt; }
auto Toml::bool_or(std::string_view path, bool dflt) const -> bool { auto v = at(path); return v && v->kind == Kind::boolean ? v->b : dflt; }

namespace {
struct TomlParser {
    std::string const& s;
    size_t i = 0;
    std::string err;
    explicit TomlParser(std::string const& text) : s(text) {}

    auto fail(std::string msg) -> bool {
        size_t line = 1;
        for (size_t k = 0; k < i && k < s.size(); k++) line += s[k] == '\n';
        err = "toml line " + std::to_string(line) + ": " + msg;
        return false;
    }
    auto skip_ws(bool newlines) -> void {
        for (;;) {
            while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r' || (newlines && s[i] == '\n'))) i++;
            if (i < s.size() && s[i] == '#') { while (i < s.size() && s[i] != '\n') i++; continue; }
            return;
        }
    }
    auto parse_key(std::string& key) -> bool {
        skip_ws(false);
        key.clear();
        if (i < s.size() && (s[i] == '"' || s[i] == '\'')) {
            char q = s[i++];
            while (i < s.size() && s[i] != q) key += s[i++];
            if (i >= s.size()) return fail("unterminated quoted key");
            i++;
            return true;
        }
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-')) key += s[i++];
        return key.empty() ? fail("expected a key") : true;
    }
    auto parse_key_path(std::vector<std::string>& path) -> bool {
        path.clear();
        for (;;) {
            std::string k;
            if (!parse_key(k)) return false;
            path.push_back(k);
            skip_ws(false);
            if (i < s.size() && s[i] == '.') { i++; continue; }
            return true;
        }
    }
    auto parse_string(Toml& v) -> bool {
        v.kind = Toml::Kind::string;
        v.str.clear();
        if (s.compare(i, 3, "'''") == 0 || s.compare(i, 3, "\"\"\"") == 0) {
            const std::string q = s.substr(i, 3);
            i += 3;
            if (i < s.size() && s[i] == '\n') i++;                      // a newline right after the opening quotes is trimmed
            auto end = s.find(q, i);
            if (end == std::string::npos) return fail("unterminated multi-line string");
            v.str = s.substr(i, end - i);
            i = end + 3;
            return true;
        }
        char q = s[i++];
        while (i < s.size() && s[i] != q && s[i] != '\n') {
            if (q == '"' && s[i] == '\\' && i + 1 < s.size()) {
                char c = s[i + 1];
                v.str += c == 'n' ? '\n' : (c == 't' ? '\t' : c);
                i += 2;
            } else v.str += s[i++];
        }
        if (i >= s.size() || s[i] != q) return fail("unterminated string");
        i++;
        return true;
    }
    int depth = 0;                                   // nesting of arrays / inline tables (a hostile file must not overflow the stack)
    struct DepthGuard { int& d; explicit DepthGuard(int& x) : d(x) { ++d; } ~DepthGuard() { --d; } };
    auto parse_value(Toml& v) -> bool {
        DepthGuard guard(depth);
        if (depth > 256) return fail("arrays / inline tables nested deeper than 256");
        skip_ws(false);
        if (i >= s.size()) return fail("expected a value");
        char c = s[i];
        if (c == '"' || c == '\'') return parse_string(v);
        if (c == '[') {
            i++;
            v.kind = Toml::Kind::array;
            for (;;) {
                skip_ws(true);
                if (i < s.size() && s[i] == ']') { i++; return true; }
                Toml e;
                if (!parse_value(e)) return false;
                v.arr.push_back(std::move(e));
                skip_ws(true);
                if (i < s.size() && s[i] == ',') { i++; continue; }
                skip_ws(true);
                if (i < s.size() && s[i] == ']') { i++; return true; }
                return fail("expected ',' or ']' in array");
            }
        }
        if (c == '{') {
            i++;
            v.kind = Toml::Kind::table;
            for (;;) {
                skip_ws(false);
                if (i < s.size() && s[i] == '}') { i++; return true; }
                std::vector<std::string> path;
                if (!parse_key_path(path)) return false;
                skip_ws(false);
                if (i >= s.size() || s[i] != '=') return fail("expected '=' in inline table");
                i++;
                Toml e;
                if (!parse_value(e)) return false;
                Toml* t = &v;
                for (size_t k = 0; k + 1 < path.size(); k++) t = &child_table(*t, path[k]);
                t->tab.emplace_back(path.back(), std::move(e));
                skip_ws(false);
                if (i < s.size() && s[i] == ',') { i++; continue; }
            }
        }
        if (s.compare(i, 4, "true") == 0) { v.kind = Toml::Kind::boolean; v.b = true; i += 4; return true; }
        if (s.compare(i, 5, "false") == 0) { v.kind = Toml::Kind::boolean; v.b = false; i += 5; return true; }
        size_t j = i;
        while (j < s.size() && (std::isalnum((unsigned char)s[j]) || s[j] == '+' || s[j] == '-' || s[j] == '.' || s[j] == '_')) j++;
        std::string tok = s.substr(i, j - i);
        tok.erase(std::remove(tok.begin(), tok.end(), '_'), tok.end());
        if (tok.empty()) return fail("expected a value");
        char* end = nullptr;
        double d = (tok == "inf" || tok == "+inf") ? INFINITY : (tok == "-inf" ? -INFINITY : std::strtod(tok.c_str(), &end));
        if (end && *end != 0) return fail("bad number '" + tok + "'");
        v.kind = Toml::Kind::number; v.num = d;
        {   // integer token: keep the exact 64-bit value beside the double
            size_t k = (tok[0] == '+' || tok[0] == '-') ? 1 : 0;
            bool digits = k < tok.size();
            for (size_t q = k; q < tok.size(); q++) digits = digits && std::isdigit((unsigned char)tok[q]);
            if (digits) {
                errno = 0;
                char* e2 = nullptr;
                if (tok[0] == '-') { long long sv = std::strtoll(tok.c_str(), &e2, 10); v.u64 = (uint64_t)sv; v.is_int = errno == 0 && sv >= 0; }
                else { v.u64 = std::strtoull(tok.c_str() + (tok[0] == '+' ? 1 : 0), &e2, 10); v.is_int = errno == 0; }
            }
        }
        i = j;
        return true;
    }
    static auto child_table(Toml& t, std::string const& key) -> Toml& {
        if (t.kind == Toml::Kind::nil) t.kind = Toml::Kind::table;
        for (auto& kv : t.tab)
            if (kv.first == key) {
                if (kv.second.kind == Toml::Kind::array && !kv.second.arr.empty()) return kv.second.arr.back();   // array of tables: its last element
                return kv.second;
            }
        t.tab.emplace_back(key, Toml{});
        t.tab.back().second.kind = Toml::Kind::table;
        return t.tab.back().second;
    }
    auto parse(Toml& root) -> bool {
        root = Toml{};
        root.kind = Toml::Kind::table;
        Toml* cur = &root;
        for (;;) {
            skip_ws(true);
            if (i >= s.size()) return true;
            if (s[i] == '[') {
                bool aot = i + 1 < s.size() && s[i + 1] == '[';
                i += aot ? 2 : 1;
                std::vector<std::string> path;
                if (!parse_key_path(path)) return false;
                skip_ws(false);
                if (s.compare(i, aot ? 2 : 1, aot ? "]]" : "]") != 0) return fail("expected closing bracket of a table header");
                i += aot ? 2 : 1;
                Toml* t = &root;
                for (size_t k = 0; k + 1 < path.size(); k++) t = &child_table(*t, path[k]);
                if (aot) {
                    Toml* arr = nullptr;
                    for (auto& kv : t->tab) if (kv.first == path.back()) arr = &kv.second;
                    if (!arr) { t->tab.emplace_back(path.back(), Toml{}); arr = &t->tab.back().second; arr->kind = Toml::Kind::array; }
                    if (arr->kind != Toml::Kind::array) return fail("'" + path.back() + "' is not an array of tables");
                    arr->arr.emplace_back();
                    arr->arr.back().kind = Toml::Kind::table;
                    cur = &arr->arr.back();
                } else {
                    cur = &child_table(*t, path.back());
                }
                continue;
            }
            std::vector<std::string> path;
            if (!parse_key_path(path)) return false;
            skip_ws(false);
            if (i >= s.size() || s[i] != '=') return fail("expected '='");
            i++;
            Toml v;
            if (!parse_value(v)) return false;
            Toml* t = cur;
            for (size_t k = 0; k + 1 < path.size(); k++) t = &child_table(*t, path[k]);
            t->tab.emplace_back(path.back(), std::move(v));
        }
    }
};

auto read_file(std::string const& path, std::string& out) -> bool {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}
} // namespace

auto parse_toml(std::string const& text, Toml& out, std::string& err) -> bool {
    TomlParser p(text);
    if (p.parse(out)) return true;
    err = p.err;
    return false;
}

// =====================================================================================================================
// .biasset
// =====================================================================================================================
namespace {
struct ByteReader {                                                      // ReadByteStream, byte_stream.hpp:31-44 / byte_stream.cpp:20-45
    const uint8_t* p = nullptr; size_t n = 0, o = 0; bool ok = true;
    template <class T> auto pod(T& v) -> ByteReader& {
        if (!ok || sizeof(T) > n - o) { ok = false; return *this; }
        std::memcpy(&v, p + o, sizeof(T)); o += sizeof(T);
        return *this;
    }
    auto str(std::string& s) -> ByteReader& {
        uint64_t len = 0; pod(len);
        if (!ok || len > n - o) { ok = false; return *this; }                      // (compared without o + len: a hostile length must not wrap)
        s.assign(reinterpret_cast<const char*>(p + o), len); o += len;
        return *this;
    }
    template <class T> auto vec(std::vector<T>& v, size_t elems_per_item = 1) -> ByteReader& {      // u64 count of ITEMS, raw elements
        uint64_t cnt = 0; pod(cnt);
        if (!ok || cnt > (n - o) / (elems_per_item * sizeof(T))) { ok = false; return *this; }
        size_t bytes = cnt * elems_per_item * sizeof(T);
        v.resize(cnt * elems_per_item);
        if (bytes) std::memcpy(v.data(), p + o, bytes);
        o += bytes;
        return *this;
    }
    auto compressed_part(std::vector<uint8_t>& raw) -> bool {                                        // byte_stream.cpp:28-45,77-97
        uint64_t ulen = 0, clen = 0; pod(ulen).pod(clen);
        if (!ok || clen > n - o || ulen > clen * 1032 + 1024) return ok = false;                     // deflate expands at most ~1032x: nothing larger is allocated
        raw.resize(ulen);
        uLongf dst = (uLongf)ulen;
        if (uncompress(raw.data(), &dst, p + o, (uLong)clen) != Z_OK || dst != ulen) return ok = false;
        o += clen;
        return true;
    }
};
auto open_asset(std::string const& path, const char* type, std::string& data, ByteReader& r, uint32_t& version, std::string& err) -> bool {
    if (!read_file(path, data)) { err = "cannot read " + path; return false; }
    r = ByteReader{reinterpret_cast<const uint8_t*>(data.data()), data.size()};
    uint32_t magic = 0; std::string type_name;
    r.pod(magic).str(type_name).pod(version);                           // static_mesh.cpp:14-20, texture.cpp:87-93
    if (!r.ok || magic != 0x0b1a55e7u || type_name != type) { err = path + ": not a " + type + " .biasset"; return false; }
    return true;
}
} // namespace

auto load_static_mesh(std::string const& path, StaticMeshData& m, std::string& err) -> bool {
    std::string data; ByteReader r; uint32_t version = 0;
    if (!open_asset(path, "StaticMesh", data, r, version, err)) return false;
    std::vector<uint8_t> raw;
    ByteReader body = r;
    if (version >= 2) {                                                 // static_mesh.cpp:24-31
        if (!r.compressed_part(raw)) { err = path + ": bad compressed part"; return false; }
        body = ByteReader{raw.data(), raw.size()};
    }
    std::vector<uint8_t> sub;                                           // mesh.cpp:137-146
    body.vec(m.positions, 3).vec(m.normals, 3).vec(m.tangents, 4).vec(m.colors, 3).vec(m.texcoords, 2).vec(m.texcoords2, 2).vec(m.indices).vec(sub, 16);
    if (!body.ok) { err = path + ": truncated mesh data"; return false; }
    m.submeshes.resize(sub.size() / 16);
    for (size_t k = 0; k < m.submeshes.size(); k++) {                   // SubmeshDesc, mesh.hpp:16-22 (16 B)
        std::memcpy(&m.submeshes[k].base_vertex, &sub[16 * k], 4); std::memcpy(&m.submeshes[k].index_offset, &sub[16 * k + 4], 4);
        std::memcpy(&m.submeshes[k].num_indices, &sub[16 * k + 8], 4); m.submeshes[k].topology = sub[16 * k + 12];
    }
    return true;
}

auto load_texture(std::string const& path, TextureData& t, std::string& err) -> bool {
    std::string data; ByteReader r; uint32_t version = 0;
    if (!open_asset(path, "Texture", data, r, version, err)) return false;
    uint8_t sampler[28] = {0}, desc[20] = {0};                                      // rhi::SamplerDesc / rhi::TextureDesc raw (sampler.hpp:37-51, resource.hpp:91-99)
    r.pod(sampler).pod(desc);
    if (!r.ok) { err = path + ": truncated header"; return false; }
    t.mag_filter = sampler[0]; t.min_filter = sampler[1]; t.address_u = sampler[3]; t.address_v = sampler[4];
    std::memcpy(&t.width, desc, 4); std::memcpy(&t.height, desc + 4, 4); std::memcpy(&t.depth, desc + 8, 4); std::memcpy(&t.levels, desc + 12, 4);
    t.format = desc[16]; t.dim = desc[17];
    if (version == 1) {                                                 // texture.cpp:105-132
        uint32_t storage = 0; r.pod(storage);
        if (storage == 1) {                                             // one PNG per layer, stbi's native channel count, bytes_per_layer copied
            const uint32_t texel = (t.format >= 9 && t.format <= 15) ? 1u : (t.format >= 16 && t.format <= 22) ? 2u : (t.format >= 37 && t.format <= 57) ? 4u : 0u;   // rhi/defines.hpp:44-100
            if (!texel) { err = path + ": PNG-per-layer storage with a format that is not 8 bits per channel"; return false; }
            const size_t layer_bytes = (size_t)t.width * t.height * texel;
            for (uint32_t layer = 0; layer < t.depth; layer++) {
                std::vector<uint8_t> png, px; uint32_t w = 0, h = 0, ch = 0; std::string e;
                r.vec(png);
                if (!r.ok) { err = path + ": truncated PNG layer"; return false; }
                if (!decode_png(std::string(png.begin(), png.end()), w, h, ch, px, false, e)) { err = path + ": layer " + std::to_string(layer) + ": " + e; return false; }
                if (w != t.width || h != t.height || ch != texel) { err = path + ": layer " + std::to_string(layer) + ": the PNG does not match the texture description"; return false; }
                t.texels.insert(t.texels.end(), px.begin(), px.begin() + (std::ptrdiff_t)layer_bytes);
            }
        } else if (storage != 0) { err = path + ": unknown storage type " + std::to_string(storage); return false; }
        else r.vec(t.texels);
    } else {
        std::vector<uint8_t> raw;
        if (!r.compressed_part(raw)) { err = path + ": bad compressed part"; return false; }
        ByteReader body{raw.data(), raw.size()};
        body.vec(t.texels);
        r.ok = body.ok;
    }
    if (!r.ok) { err = path + ": truncated texel data"; return false; }
    // untrusted header: level 0 of the payload must hold width x height texels of the declared format before anything reads it
    // (upload_project hands texels.data() to bpt_scene_upload_materials, which copies / decodes width*height*bytes of it)
    if (t.width == 0 || t.height == 0 || t.width > 65536 || t.height > 65536) { err = path + ": texture extent " + std::to_string(t.width) + " x " + std::to_string(t.height) + " out of range"; return false; }
    const size_t texel_bytes = (t.format >= 9 && t.format <= 15) ? 1u : (t.format >= 16 && t.format <= 22) ? 2u : (t.format >= 37 && t.format <= 57) ? 4u
                             : (t.format == 109 ? 16u : 0u);                                                                    // rhi/defines.hpp:44-100
    if (texel_bytes && t.texels.size() < (size_t)t.width * t.height * texel_bytes) {
        err = path + ": texel payload (" + std::to_string(t.texels.size()) + " bytes) is smaller than " + std::to_string(t.width) + " x " + std::to_string(t.height) + " texels of " + std::to_string(texel_bytes) + " bytes";
        return false;
    }
    return true;
}

// =====================================================================================================================
// project
// =====================================================================================================================
namespace {
auto vec3_of(Toml const* v, double d0, double d1, double d2, double out[3]) -> void {
    out[0] = d0; out[1] = d1; out[2] = d2;
    if (v && v->kind == Toml::Kind::array)
        for (size_t k = 0; k < 3 && k < v->arr.size(); k++) if (v->arr[k].kind == Toml::Kind::number) out[k] = v->arr[k].num;
}
// Transform: rotation = euler degrees, R = Rz * Rx * Ry, M = T * R * S (src/math/transform.cpp:98-123)
struct Xf { double r[9]; double s[3]; double t[3]; };
auto transform_of(Toml const& value) -> Xf {
    Xf x{};
    double rot[3];
    vec3_of(value.find("rotation"), 0, 0, 0, rot); vec3_of(value.find("scaling"), 1, 1, 1, x.s); vec3_of(value.find("translation"), 0, 0, 0, x.t);
    const double d2r = 3.14159265358979323846 / 180.0;
    double cx = std::cos(rot[0] * d2r), sx = std::sin(rot[0] * d2r), cy = std::cos(rot[1] * d2r), sy = std::sin(rot[1] * d2r), cz = std::cos(rot[2] * d2r), sz = std::sin(rot[2] * d2r);
    double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1}, Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    auto mul = [](const double* a, const double* b, double* o) {
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
    };
    double zx[9];
    mul(Rz, Rx, zx); mul(zx, Ry, x.r);
    return x;
}
auto light_transform(Xf const& x) -> LightTransform {
    LightTransform lt;
    lt.translation = {(float)x.t[0], (float)x.t[1], (float)x.t[2]};
    for (int k = 0; k < 9; k++) lt.rotation[k] = (float)x.r[k];
    return lt;
}
auto squeeze(std::string s) -> std::string { s.erase(std::remove_if(s.begin(), s.end(), [](unsigned char c) { return std::isspace(c); }), s.end()); return s; }

struct MaterialFile { bpt_material m{}; std::vector<std::pair<int*, uint64_t>> tex_refs; };     // (field, asset id of the texture)
}  // namespace

auto load_project(std::string const& dir, Project& out, std::string& err) -> bool {
    out = Project{};
    auto resolve = [&](std::string const& p) { return p.rfind("/project/", 0) == 0 ? dir + p.substr(8) : p; };     // the "/project" mount of the VFS
    std::string text; Toml proj, meta, scene;
    if (!read_file(dir + "/project.toml", text) || !parse_toml(text, proj, err)) { if (err.empty()) err = "cannot read " + dir + "/project.toml"; return false; }
    const std::string meta_path = resolve(proj.string_or("asset_metadata_file", "/project/asset_metadata.toml"));
    const std::string scene_path = resolve(proj.string_or("scene_file", "/project/scene.toml"));
    if (!read_file(meta_path, text) || !parse_toml(text, meta, err)) { if (err.empty()) err = "cannot read " + meta_path; return false; }
    if (!read_file(scene_path, text) || !parse_toml(text, scene, err)) { if (err.empty()) err = "cannot read " + scene_path; return false; }

    struct AssetRef { std::string path, type; };
    std::map<uint64_t, AssetRef> assets;
    if (auto a = meta.find("assets"); a && a->kind == Toml::Kind::array)
        for (auto& e : a->arr) assets[e.id_or("id", ~0ull)] = AssetRef{resolve(e.string_or("path", "")), e.string_or("type", "")};

    std::map<uint64_t, uint32_t> mesh_first_blas, texture_index, material_index;     // asset id -> index, loaded on first use
    std::vector<std::vector<uint32_t>> mesh_blas;                                    // per loaded mesh: BLAS index per submesh
    std::vector<std::vector<StaticMeshData::Submesh>> mesh_submeshes;
    std::vector<uint32_t> mesh_vbase, mesh_ibase;

    auto need_texture = [&](uint64_t id, int& index) -> bool {
        if (auto it = texture_index.find(id); it != texture_index.end()) { index = (int)it->second; return true; }
        auto a = assets.find(id);
        if (a == assets.end() || a->second.type != "Texture") { err = "asset " + std::to_string(id) + " is not a texture"; return false; }
        TextureData t;
        if (!load_texture(a->second.path, t, err)) return false;
        if (t.dim != 1 || (t.format != 37 && t.format != 43 && t.format != 109)) { err = a->second.path + ": only 2-D rgba8_unorm / rgba8_srgb / rgba32_sfloat textures are supported"; return false; }
        index = (int)out.textures.size();
        texture_index[id] = (uint32_t)index;
        out.textures.push_back(std::move(t));
        return true;
    };
    auto need_material = [&](uint64_t id, uint32_t& index) -> bool {
        if (auto it = material_index.find(id); it != material_index.end()) { index = it->second; return true; }
        auto a = assets.find(id);
        if (a == assets.end() || a->second.type != "Material") { err = "asset " + std::to_string(id) + " is not a material"; return false; }
        std::string mt; Toml mf;
        if (!read_file(a->second.path, mt) || !parse_toml(mt, mf, err)) { if (err.empty()) err = "cannot read " + a->second.path; return false; }
        // material.cpp:12-90: surface_model, blend_mode, material_function, params (scalar / array / {asset_id})
        std::map<std::string, Toml const*> params;
        if (auto ps = mf.find("params"); ps && ps->kind == Toml::Kind::array)
            for (auto& p : ps->arr) params[p.string_or("name", "")] = p.find("value");
        auto pnum = [&](const char* n, double d) { auto it = params.find(n); return it != params.end() && it->second && it->second->kind == Toml::Kind::number ? it->second->num : d; };
        auto pvec = [&](const char* n, double d, double o[3]) { auto it = params.find(n); vec3_of(it != params.end() ? it->second : nullptr, d, d, d, o); };
        auto ptex = [&](const char* n, int& idx) -> bool {
            auto it = params.find(n);
            if (it == params.end() || !it->second || it->second->kind != Toml::Kind::table) { err = a->second.path + ": texture parameter '" + n + "' missing"; return false; }
            return need_texture(it->second->id_or("asset_id", ~0ull), idx);
        };
        bpt_material m{};
        m.base_color[0] = m.base_color[1] = m.base_color[2] = 0.5f; m.base_color[3] = 1.0f;
        m.roughness = 0.5f; m.normal_map_scale = 1.0f; m.occlusion_strength = 1.0f;
        m.base_color_tex = m.metallic_roughness_tex = m.normal_map_tex = m.occlusion_tex = -1;
        const std::string fn = squeeze(mf.string_or("material_function", ""));
        const std::string blend_s = mf.string_or("blend_mode", "opaque"), model_s = mf.string_or("surface_model", "lit");
        uint32_t blend = blend_s == "opaque" ? BPT_BLEND_OPAQUE : (blend_s == "alpha_test" ? BPT_BLEND_ALPHA_TEST : BPT_BLEND_TRANSLUCENT);
        uint32_t model = model_s == "lit" ? BPT_SURFACE_MODEL_LIT : BPT_SURFACE_MODEL_UNLIT;
        uint32_t kind = 0, two_sided = fn.find("surface.two_sided=true;") != std::string::npos ? 1u : 0u;
        double v[3];
        // the closed set of snippets (examples/scene_basic/materials/*.toml), matched on their whitespace-free text
        if (fn == "surface.base_color=PARAM_base_color;") {
            kind = BPT_MATERIAL_KIND_CONSTANT_COLOR; pvec("base_color", 0.5, v);
            m.base_color[0] = (float)v[0]; m.base_color[1] = (float)v[1]; m.base_color[2] = (float)v[2];
        } else if (fn == "intgrid=int(floor(vertex.position_world.x))^int(floor(vertex.position_world.z));surface.base_color=(grid&1)==1?PARAM_base_color_0:PARAM_base_color_1;"
                         "surface.roughness=(grid&1)==1?PARAM_roughness_0:PARAM_roughness_1;") {
            kind = BPT_MATERIAL_KIND_CHECKERBOARD;
            pvec("base_color_0", 0.5, v); m.base_color[0] = (float)v[0]; m.base_color[1] = (float)v[1]; m.base_color[2] = (float)v[2];
            m.base_color[3] = (float)pnum("roughness_0", 0.5);
            pvec("base_color_1", 0.5, v); m.emission[0] = (float)v[0]; m.emission[1] = (float)v[1]; m.emission[2] = (float)v[2];
            m.roughness = (float)pnum("roughness_1", 0.5);
        } else if (fn == "surface.base_color=PARAM_base_color_tex.Sample(PARAM_base_color_tex_sampler,vertex.texcoord).xyz;"
                         "surface.normal_map_value=PARAM_normal_map.Sample(PARAM_normal_map_sampler,vertex.texcoord).xyz;surface.roughness=PARAM_roughness;") {
            kind = BPT_MATERIAL_KIND_TEXTURED; m.roughness = (float)pnum("roughness", 0.5);
            if (!ptex("base_color_tex", m.base_color_tex) || !ptex("normal_map", m.normal_map_tex)) return false;
        } else if (fn == "surface.base_color=PARAM_base_color;surface.opacity=PARAM_opacity;surface.two_sided=true;") {
            kind = BPT_MATERIAL_KIND_TRANSPARENT; pvec("base_color", 0.5, v);
            m.base_color[0] = (float)v[0]; m.base_color[1] = (float)v[1]; m.base_color[2] = (float)v[2]; m.base_color[3] = (float)pnum("opacity", 1.0);
        } else if (fn == "float4value=PARAM_cage_tex.Sample(PARAM_cage_tex_sampler,vertex.texcoord);surface.base_color=value.xyz;surface.f0_color=value.xyz;"
                         "surface.opacity=value.w<0.5?0.0:1.0;surface.two_sided=true;") {
            kind = BPT_MATERIAL_KIND_CAGE;
            if (!ptex("cage_tex", m.base_color_tex)) return false;
        } else if (fn.empty()) {
            kind = BPT_MATERIAL_KIND_DEFAULT;
        } else {
            err = a->second.path + ": material_function is not one of the snippets the CUDA kernels restate (bpt.h BPT_MATERIAL_KIND_*)";
            return false;
        }
        m.flags = two_sided | (kind << BPT_MATERIAL_KIND_SHIFT) | (blend << BPT_MATERIAL_BLEND_SHIFT) | (model << BPT_MATERIAL_MODEL_SHIFT);
        index = (uint32_t)out.materials.size();
        material_index[id] = index;
        out.materials.push_back(m);
        return true;
    };
    auto need_mesh = [&](uint64_t id, uint32_t& mesh) -> bool {
        if (auto it = mesh_first_blas.find(id); it != mesh_first_blas.end()) { mesh = it->second; return true; }
        auto a = assets.find(id);
        if (a == assets.end() || a->second.type != "StaticMesh") { err = "asset " + std::to_string(id) + " is not a static mesh"; return false; }
        StaticMeshData sm;
        if (!load_static_mesh(a->second.path, sm, err)) return false;
        const size_t nv = sm.positions.size() / 3;
        if (sm.normals.size() != nv * 3 || sm.tangents.size() != nv * 4 || sm.texcoords.size() != nv * 2) { err = a->second.path + ": missing normals / tangents / texcoords"; return false; }
        const uint32_t vbase = (uint32_t)(out.positions.size() / 3), ibase = (uint32_t)out.indices.size();
        out.positions.insert(out.positions.end(), sm.positions.begin(), sm.positions.end());
        out.normals.insert(out.normals.end(), sm.normals.begin(), sm.normals.end());
        out.tangents.insert(out.tangents.end(), sm.tangents.begin(), sm.tangents.end());
        out.texcoords.insert(out.texcoords.end(), sm.texcoords.begin(), sm.texcoords.end());
        out.indices.insert(out.indices.end(), sm.indices.begin(), sm.indices.end());
        std::vector<uint32_t> per_sub;
        for (auto& sub : sm.submeshes) {                                   // BLAS per (mesh, submesh): graphics_manager.cpp:616-654
            uint32_t avail = (uint32_t)sm.indices.size() - std::min<uint32_t>(sub.index_offset, (uint32_t)sm.indices.size());
            uint32_t num = std::min<uint32_t>(sub.num_indices, avail);     // num_indices = ~0u means "to the end" (mesh.hpp:16-22)
            // a file is untrusted input: every index must address a vertex of this mesh (the kernels fetch positions / attributes unchecked)
            if (sub.base_vertex > nv) { err = a->second.path + ": submesh base_vertex beyond the vertex count"; return false; }
            const size_t first = std::min<size_t>(sub.index_offset, sm.indices.size());
            for (size_t k = first; k < first + num / 3 * 3; k++)
                if (sm.indices[k] >= nv - sub.base_vertex) { err = a->second.path + ": index " + std::to_string(sm.indices[k]) + " beyond the vertex count"; return false; }
            bpt_blas_desc bd{};
            bd.position_offset = (vbase + sub.base_vertex) * 3; bd.index_offset = ibase + sub.index_offset; bd.num_triangles = num / 3;
            per_sub.push_back((uint32_t)out.blas.size());
            out.blas.push_back(bd);
        }
        mesh = (uint32_t)mesh_blas.size();
        mesh_first_blas[id] = mesh;
        mesh_blas.push_back(per_sub); mesh_submeshes.push_back(sm.submeshes); mesh_vbase.push_back(vbase); mesh_ibase.push_back(ibase);
        return true;
    };

    auto objects = scene.find("objects");
    if (!objects || objects->kind != Toml::Kind::array) { err = scene_path + ": no [[objects]]"; return false; }
    bool have_camera = false;
    for (auto& obj : objects->arr) {
        auto comps = obj.find("components");
        if (!comps || comps->kind != Toml::Kind::array) continue;
        Xf xf = transform_of(Toml{});
        Toml const* mesh_c = nullptr; Toml const* renderer_c = nullptr;
        for (auto& c : comps->arr) {                                      // the Transform first: the other components read it
            if (c.string_or("type", "") == "Transform") if (auto v = c.find("value")) xf = transform_of(*v);
        }
        for (auto& c : comps->arr) {
            const std::string type = c.string_or("type", "");
            Toml empty; empty.kind = Toml::Kind::table;
            Toml const& v = c.find("value") ? *c.find("value") : empty;
            if (type == "CameraComponent" && !have_camera) {              // scene_basic/camera.hpp:13-25, camera_system.cpp:17-29
                have_camera = true;
                for (int k = 0; k < 3; k++) {
                    out.cam_position[k] = (float)xf.t[k];
                    out.cam_front[k] = (float)(-xf.r[3 * k + 2]);          // R * (0, 0, -1)
                    out.cam_up[k] = (float)xf.r[3 * k + 1];                // R * (0, 1, 0)
                }
                out.yfov = (float)v.number_or("yfov", 30.0); out.near_z = (float)v.number_or("near_z", 0.001); out.far_z = (float)v.number_or("far_z", 1e5);
                out.orthographic = v.string_or("projection_type", "perspective") == "orthographic";
                if (auto sz = v.find("render_target_size"); sz && sz->kind == Toml::Kind::array && sz->arr.size() == 2) {
                    out.target_width = (uint32_t)sz->arr[0].num; out.target_height = (uint32_t)sz->arr[1].num;
                }
            } else if (type == "DirectionalLightComponent") {             // scene_basic/light.hpp, lights.cpp:52-63
                DirectionalLightComponent l; double col[3];
                vec3_of(v.find("color"), 1, 1, 1, col);
                l.color = {(float)col[0], (float)col[1], (float)col[2]}; l.strength = (float)v.number_or("strength", 1.0); l.cast_shadow = v.bool_or("cast_shadow", false);
                out.lights.add(l, light_transform(xf));
            } else if (type == "PointLightComponent") {                   // lights.cpp:125-143
                PointLightComponent l; double col[3];
                vec3_of(v.find("color"), 1, 1, 1, col);
                l.color = {(float)col[0], (float)col[1], (float)col[2]}; l.strength = (float)v.number_or("strength", 1.0); l.range = (float)v.number_or("range", 30.0);
                l.spot = v.bool_or("spot", false); l.spot_inner_angle = (float)v.number_or("spot_inner_angle", 30.0); l.spot_outer_angle = (float)v.number_or("spot_outer_angle", 60.0);
                out.lights.add(l, light_transform(xf));
            } else if (type == "RectLightComponent") {                    // lights.cpp:208-229
                RectLightComponent l; double col[3];
                vec3_of(v.find("color"), 1, 1, 1, col);
                l.color = {(float)col[0], (float)col[1], (float)col[2]}; l.strength = (float)v.number_or("strength", 1.0);
                l.width = (float)v.number_or("width", 1.0); l.height = (float)v.number_or("height", 1.0); l.two_sided = v.bool_or("two_sided", false);
                out.lights.add(l, light_transform(xf));
            } else if (type == "StaticMeshComponent") mesh_c = &v;
            else if (type == "MeshRendererComponent") renderer_c = &v;
            else if (type == "BasicRendererOverrideVolume") {             // renderer/basic.hpp:40-50,76-81
                if (auto pt = v.at("settings.path_tracing")) {
                    out.path_tracing.ray_length = (float)pt->number_or("ray_length", 100.0); out.path_tracing.max_bounces = (uint32_t)pt->number_or("max_bounces", 3);
                    out.path_tracing.accumulate = pt->bool_or("accumulate", true); out.path_tracing.denoise = pt->bool_or("denoise", true);
                }
                if (auto ao = v.at("settings.ambient_occlusion")) {
                    out.ambient_occlusion.range = (float)ao->number_or("range", 0.5); out.ambient_occlusion.strength = (float)ao->number_or("strength", 0.5);
                    out.ambient_occlusion.half_resolution = ao->bool_or("half_resolution", true) ? 1u : 0u;
                }
            }
        }
        if (!mesh_c || !renderer_c) continue;
        // one drawable per submesh from submesh_start_index on, material k for submesh start + k (static_mesh_render_system.cpp)
        uint32_t mesh = 0;
        if (!need_mesh(mesh_c->id_or("static_mesh.asset_id", ~0ull), mesh)) return false;
        auto mats = renderer_c->find("materials");
        const uint32_t start = (uint32_t)renderer_c->number_or("submesh_start_index", 0);
        for (size_t k = 0; mats && mats->kind == Toml::Kind::array && k < mats->arr.size(); k++) {
            const uint32_t sub = start + (uint32_t)k;
            if (sub >= mesh_blas[mesh].size()) break;
            uint32_t mat = 0;
            if (!need_material(mats->arr[k].id_or("asset_id", ~0ull), mat)) return false;
            const uint32_t vb = mesh_vbase[mesh] + mesh_submeshes[mesh][sub].base_vertex, io = mesh_ibase[mesh] + mesh_submeshes[mesh][sub].index_offset;
            const uint32_t index = (uint32_t)out.drawables.size();
            bpt_drawable_sbt_data dr{};                                    // drawable_stb_data.hpp:7-17, written at graphics_manager.cpp:1318-1329
            dr.drawable_index = index; dr.position_offset = vb * 3; dr.normal_offset = vb * 3; dr.tangent_offset = vb * 4; dr.color_offset = 0;
            dr.texcoord_offset = vb * 2; dr.texcoord2_offset = 0; dr.index_offset = io; dr.material_offset = mat * (uint32_t)sizeof(bpt_material);
            out.drawables.push_back(dr);
            out.drawable_va.push_back(BPT_VA_POSITION | BPT_VA_NORMAL | BPT_VA_TANGENT | BPT_VA_TEXCOORD);
            bpt_instance_desc in{};                                        // accel.cpp:104-132
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) in.transform[r][c] = (float)(xf.r[3 * r + c] * xf.s[c]);
                in.transform[r][3] = (float)xf.t[r];
            }
            const uint32_t blend = (out.materials[mat].flags >> BPT_MATERIAL_BLEND_SHIFT) & 0xffu;
            in.instance_id_and_mask = index | (0xffu << 24);
            in.sbt_offset_and_flags = index | ((blend == BPT_BLEND_OPAQUE ? BPT_INSTANCE_FORCE_OPAQUE : BPT_INSTANCE_FORCE_NON_OPAQUE) << 24);
            in.blas = mesh_blas[mesh][sub];
            out.instances.push_back(in);
            out.object_names.push_back(obj.string_or("name", ""));
        }
    }
    if (out.drawables.empty()) { err = scene_path + ": no renderable object"; return false; }
    return true;
}

auto upload_project(Project const& p, bpt_context* ctx, uint32_t accel_mode, std::string& err) -> bpt_status {
    auto fail = [&](bpt_status s) { err = bpt_last_error(ctx); return s; };
    bpt_geometry_streams g{};
    g.positions = p.positions.data(); g.num_position_floats = p.positions.size();
    g.normals = p.normals.data(); g.num_normal_floats = p.normals.size();
    g.tangents = p.tangents.data(); g.num_tangent_floats = p.tangents.size();
    g.texcoords = p.texcoords.data(); g.num_texcoord_floats = p.texcoords.size();
    g.indices = p.indices.data(); g.num_indices = p.indices.size();
    bpt_status s;
    if ((s = bpt_scene_upload_geometry(ctx, &g, p.drawables.data(), p.drawable_va.data(), (uint32_t)p.drawables.size(), p.blas.data(), (uint32_t)p.blas.size()))) return fail(s);
    std::vector<bpt_texture_desc> td(p.textures.size());
    for (size_t k = 0; k < p.textures.size(); k++) {
        auto& t = p.textures[k];
        td[k].texels = t.texels.data(); td[k].width = t.width; td[k].height = t.height;
        td[k].format = t.format == 43 ? BPT_TEXTURE_RGBA8_SRGB : (t.format == 109 ? BPT_TEXTURE_RGBA32_FLOAT : BPT_TEXTURE_RGBA8_UNORM);   // rhi/defines.hpp:44-138
        td[k].address_mode_u = t.address_u == 0 ? BPT_ADDRESS_REPEAT : BPT_ADDRESS_CLAMP; td[k].address_mode_v = t.address_v == 0 ? BPT_ADDRESS_REPEAT : BPT_ADDRESS_CLAMP;
        td[k].filter_linear = t.mag_filter == 1 ? 1u : 0u;
    }
    if ((s = bpt_scene_upload_materials(ctx, p.materials.data(), (uint32_t)p.materials.size(), td.data(), (uint32_t)td.size()))) return fail(s);
    if ((s = bpt_scene_upload_instances(ctx, p.instances.data(), (uint32_t)p.instances.size()))) return fail(s);
    if (!p.lights.rect_lights.empty() && !p.lights.ltc_luts.matrix_lut0) { err = "rect lights need LightsContext::set_ltc_luts"; return BPT_ERR_INVALID; }
    if ((s = bpt_scene_upload_lights(ctx, p.lights.dir_lights.data(), (uint32_t)p.lights.dir_lights.size(), p.lights.point_lights.data(), (uint32_t)p.lights.point_lights.size(),
                                     p.lights.rect_lights.data(), (uint32_t)p.lights.rect_lights.size(), &p.lights.ltc_luts))) return fail(s);
    const float ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, white[3] = {1, 1, 1};
    if ((s = bpt_scene_upload_sky(ctx, nullptr, 0, ident, white))) return fail(s);          // scene_basic has no skybox component: black 1x1 default (skybox.cpp:43-73)
    if ((s = bpt_build_accel(ctx, accel_mode))) return fail(s);
    return BPT_OK;
}

} // namespace bi::project
