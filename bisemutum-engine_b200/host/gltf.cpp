// host/gltf.cpp — see gltf.hpp. Reference: bisemutum/src/scene_basic/menu_actions/import_model.cpp:27-430,
// bisemutum/src/scene_basic/static_mesh.cpp:93-152, bisemutum/src/math/transform.cpp:11-52, bisemutum/src/runtime/scene_object.cpp:31-52.
#include "gltf.hpp"
#include <zlib.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

namespace bi::project {
namespace {

// ---------------------------------------------------------------------------------------------------------------------------------
// JSON (RFC 8259) — the part of tinygltf's parser the importer needs
// ---------------------------------------------------------------------------------------------------------------------------------
struct Json {
    enum class Kind { nil, boolean, number, string, array, object } kind = Kind::nil;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    auto get(std::string_view key) const -> Json const* {
        if (kind != Kind::object) return nullptr;
        for (auto& [k, v] : obj) if (k == key) return &v;
        return nullptr;
    }
    auto number_or(std::string_view key, double d) const -> double { auto v = get(key); return v && v->kind == Kind::number ? v->num : d; }
    auto int_or(std::string_view key, int d) const -> int { auto v = get(key); return v && v->kind == Kind::number && v->num > -2e9 && v->num < 2e9 ? (int)v->num : d; }
    // byte offsets / counts / indices: a negative, fractional-huge or NaN number becomes a value every bounds check rejects
    static auto to_size(double x) -> size_t { return x >= 0.0 && x < 9e15 ? (size_t)x : (size_t)-1 / 4; }
    auto size_or(std::string_view key, size_t d) const -> size_t { auto v = get(key); return v && v->kind == Kind::number ? to_size(v->num) : d; }
    auto string_or(std::string_view key, std::string d) const -> std::string { auto v = get(key); return v && v->kind == Kind::string ? v->str : d; }
    auto bool_or(std::string_view key, bool d) const -> bool { auto v = get(key); return v && v->kind == Kind::boolean ? v->b : d; }
    auto array_of(std::string_view key) const -> std::vector<Json> const& {
        static const std::vector<Json> empty;
        auto v = get(key);
        return v && v->kind == Kind::array ? v->arr : empty;
    }
};

struct JsonParser {
    const char* p; const char* end; std::string err; int depth = 0;
    auto ws() -> void { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    auto fail(const char* what) -> bool { if (err.empty()) err = what; return false; }
    auto hex4(unsigned& out) -> bool {
        if (end - p < 4) return fail("json: short \\u escape");
        out = 0;
        for (int k = 0; k < 4; k++) {
            char c = *p++; out <<= 4;
            if (c >= '0' && c <= '9') out |= (unsigned)(c - '0'); else if (c >= 'a' && c <= 'f') out |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') out |= (unsigned)(c - 'A' + 10); else return fail("json: bad \\u escape");
        }
        return true;
    }
    auto utf8(std::string& s, unsigned cp) -> void {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
        else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    }
    auto string(std::string& s) -> bool {
        if (p >= end || *p != '"') return fail("json: expected string");
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                if (++p >= end) break;
                char c = *p++;
                switch (c) {
                    case '"': s += '"'; break; case '\\': s += '\\'; break; case '/': s += '/'; break; case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break; case 'n': s += '\n'; break; case 'r': s += '\r'; break; case 't': s += '\t'; break;
                    case 'u': {
                        unsigned cp, lo;
                        if (!hex4(cp)) return false;
                        if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') { p += 2; if (!hex4(lo)) return false; cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); }
                        utf8(s, cp); break;
                    }
                    default: return fail("json: bad escape");
                }
            } else s += *p++;
        }
        if (p >= end) return fail("json: unterminated string");
        p++;
        return true;
    }
    auto value(Json& v) -> bool {
        if (++depth > 256) return fail("json: nesting too deep");
        ws();
        if (p >= end) return fail("json: unexpected end");
        bool ok = true;
        if (*p == '{') {
            v.kind = Json::Kind::object; p++; ws();
            if (p < end && *p == '}') p++;
            else for (;;) {
                ws();
                std::string k;
                if (!string(k)) { ok = false; break; }
                ws();
                if (p >= end || *p != ':') { ok = fail("json: expected ':'"); break; }
                p++;
                v.obj.emplace_back(std::move(k), Json{});
                if (!value(v.obj.back().second)) { ok = false; break; }
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = fail("json: expected ',' or '}'"); break;
            }
        } else if (*p == '[') {
            v.kind = Json::Kind::array; p++; ws();
            if (p < end && *p == ']') p++;
            else for (;;) {
                v.arr.emplace_back();
                if (!value(v.arr.back())) { ok = false; break; }
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = fail("json: expected ',' or ']'"); break;
            }
        } else if (*p == '"') { v.kind = Json::Kind::string; ok = string(v.str); }
        else if (end - p >= 4 && !std::strncmp(p, "true", 4)) { v.kind = Json::Kind::boolean; v.b = true; p += 4; }
        else if (end - p >= 5 && !std::strncmp(p, "false", 5)) { v.kind = Json::Kind::boolean; v.b = false; p += 5; }
        else if (end - p >= 4 && !std::strncmp(p, "null", 4)) { v.kind = Json::Kind::nil; p += 4; }
        else {
            const char* q = p;
            if (q < end && (*q == '-' || *q == '+')) q++;
            while (q < end && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '-' || *q == '+')) q++;
            if (q == p) ok = fail("json: unexpected character");
            else { v.kind = Json::Kind::number; v.num = std::strtod(std::string(p, q).c_str(), nullptr); p = q; }
        }
        depth--;
        return ok;
    }
};

auto read_binary(std::string const& path, std::string& out) -> bool {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss; ss << f.rdbuf(); out = ss.str();
    return true;
}

auto base64_decode(std::string_view in, std::string& out) -> bool {
    int val = 0, bits = -8;
    for (unsigned char c : in) {
        int d;
        if (c >= 'A' && c <= 'Z') d = c - 'A'; else if (c >= 'a' && c <= 'z') d = c - 'a' + 26; else if (c >= '0' && c <= '9') d = c - '0' + 52;
        else if (c == '+' || c == '-') d = 62; else if (c == '/' || c == '_') d = 63; else if (c == '=') break; else if (c == '\n' || c == '\r') continue; else return false;
        val = (val << 6) | d; bits += 6;
        if (bits >= 0) { out += (char)((val >> bits) & 0xFF); bits -= 8; }
    }
    return true;
}

// a buffer or image `uri`: "data:<mime>;base64,<payload>" or a path relative to the .gltf file (tinygltf's LoadExternalFile / DecodeDataURI)
auto load_uri(std::string const& uri, std::string const& base_dir, std::string& out, std::string& err) -> bool {
    if (uri.rfind("data:", 0) == 0) {
        auto comma = uri.find(',');
        if (comma == std::string::npos || uri.substr(0, comma).find(";base64") == std::string::npos) { err = "unsupported data URI"; return false; }
        if (!base64_decode(std::string_view(uri).substr(comma + 1), out)) { err = "bad base64 payload"; return false; }
        return true;
    }
    std::string decoded;                                                  // percent-decoding of file URIs
    for (size_t i = 0; i < uri.size(); i++) {
        if (uri[i] == '%' && i + 2 < uri.size() && std::isxdigit((unsigned char)uri[i + 1]) && std::isxdigit((unsigned char)uri[i + 2])) {
            decoded += (char)std::stoi(uri.substr(i + 1, 2), nullptr, 16); i += 2;
        } else decoded += uri[i];
    }
    if (!read_binary(base_dir + decoded, out)) { err = "cannot read '" + base_dir + decoded + "'"; return false; }
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// PNG (8-bit, non-interlaced; grey / grey+alpha / RGB / RGBA / palette) -> RGBA8, the conversion stb_image applies for tinygltf's
// default `req_comp = 4` (grey -> (g, g, g, 255), grey+alpha -> (g, g, g, a), RGB -> (r, g, b, 255)).
// ---------------------------------------------------------------------------------------------------------------------------------
auto be32(const unsigned char* p) -> uint32_t { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

} // namespace

auto decode_png(std::string const& file, uint32_t& w, uint32_t& h, uint32_t& channels, std::vector<uint8_t>& rgba, bool force_rgba, std::string& err) -> bool {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto d = reinterpret_cast<const unsigned char*>(file.data());
    if (file.size() < 8 || std::memcmp(d, sig, 8)) { err = "image is not a PNG (only PNG images are decoded by the headless importer)"; return false; }
    size_t o = 8;
    unsigned depth = 0, ctype = 0, interlace = 0;
    std::string idat;
    std::vector<uint8_t> palette, trns;
    bool have_ihdr = false;
    while (o + 12 <= file.size()) {
        uint32_t len = be32(d + o);
        if (o + 12 + (size_t)len > file.size()) { err = "PNG: truncated chunk"; return false; }
        const char* type = file.data() + o + 4; const unsigned char* body = d + o + 8;
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) { w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; have_ihdr = true; }
        else if (!std::memcmp(type, "PLTE", 4)) palette.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.append(reinterpret_cast<const char*>(body), len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        o += 12 + (size_t)len;
    }
    if (!have_ihdr || w == 0 || h == 0) { err = "PNG: no IHDR"; return false; }
    if (depth != 8 || interlace != 0) { err = "PNG: only 8-bit non-interlaced images are supported"; return false; }
    unsigned ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) { err = "PNG: bad colour type"; return false; }
    const size_t stride = (size_t)w * ch;
    // a header that promises more than the IDAT stream can inflate to (deflate expands at most ~1032x) is rejected before anything is allocated
    if (w > 65536 || h > 65536 || (stride + 1) * h > idat.size() * 1032 + 1024) { err = "PNG: image size does not match its data"; return false; }
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf dst = (uLongf)raw.size();
    if (uncompress(raw.data(), &dst, reinterpret_cast<const Bytef*>(idat.data()), (uLong)idat.size()) != Z_OK || dst != raw.size()) { err = "PNG: bad IDAT stream"; return false; }
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; y++) {                                    // PNG filters 0..4 (None, Sub, Up, Average, Paeth)
        const uint8_t f = raw[(stride + 1) * y];
        const uint8_t* src = &raw[(stride + 1) * y + 1];
        uint8_t* cur = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= ch ? cur[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= ch) ? up[x - ch] : 0;
            int pred = 0;
            switch (f) {
                case 0: pred = 0; break; case 1: pred = a; break; case 2: pred = b; break; case 3: pred = (a + b) >> 1; break;
                case 4: { int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: err = "PNG: bad filter type"; return false;
            }
            cur[x] = (uint8_t)(src[x] + pred);
        }
    }
    if (!force_rgba) {                                                     // stb_image with req_comp = 0: the file's own channel count, palettes expanded
        channels = ctype == 3 ? (trns.empty() ? 3u : 4u) : ch;
        if (ctype != 3) { rgba = std::move(img); return true; }
        rgba.resize((size_t)w * h * channels);
        for (size_t i = 0; i < (size_t)w * h; i++) {
            if ((size_t)img[i] * 3 + 2 >= palette.size()) { err = "PNG: palette index out of range"; return false; }
            for (uint32_t c = 0; c < 3; c++) rgba[channels * i + c] = palette[img[i] * 3 + c];
            if (channels == 4) rgba[4 * i + 3] = img[i] < trns.size() ? trns[img[i]] : 255;
        }
        return true;
    }
    channels = 4;
    rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t* q = &rgba[4 * i]; const uint8_t* s = &img[(size_t)ch * i];
        switch (ctype) {
            case 0: q[0] = q[1] = q[2] = s[0]; q[3] = 255; break;
            case 2: q[0] = s[0]; q[1] = s[1]; q[2] = s[2]; q[3] = 255; break;
            case 3: {
                if ((size_t)s[0] * 3 + 2 >= palette.size()) { err = "PNG: palette index out of range"; return false; }
                q[0] = palette[s[0] * 3]; q[1] = palette[s[0] * 3 + 1]; q[2] = palette[s[0] * 3 + 2]; q[3] = s[0] < trns.size() ? trns[s[0]] : 255; break;
            }
            case 4: q[0] = q[1] = q[2] = s[0]; q[3] = s[1]; break;
            default: q[0] = s[0]; q[1] = s[1]; q[2] = s[2]; q[3] = s[3]; break;
        }
    }
    return true;
}

namespace {

// ---------------------------------------------------------------------------------------------------------------------------------
// FP32 vector helpers in MikkTSpace's operation order
// ---------------------------------------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline auto vsub(V3 a, V3 b) -> V3 { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline auto vadd(V3 a, V3 b) -> V3 { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline auto vscale(float s, V3 v) -> V3 { return {s * v.x, s * v.y, s * v.z}; }
inline auto vdot(V3 a, V3 b) -> float { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline auto vlen(V3 v) -> float { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
inline auto not_zero(float f) -> bool { return std::fabs(f) > 1.17549435e-38f; }                 // FLT_MIN
inline auto vnot_zero(V3 v) -> bool { return not_zero(v.x) || not_zero(v.y) || not_zero(v.z); }
inline auto vnormalize(V3 v) -> V3 { return vscale(1.0f / vlen(v), v); }
inline auto project_normalize(V3 v, V3 n) -> V3 { V3 r = vsub(v, vscale(vdot(n, v), n)); return vnot_zero(r) ? vnormalize(r) : r; }

struct VertexKey {
    uint32_t w[8];
    auto operator==(VertexKey const& o) const -> bool { return !std::memcmp(w, o.w, sizeof(w)); }
};
struct VertexKeyHash {
    auto operator()(VertexKey const& k) const -> size_t { uint64_t h = 1469598103934665603ull; for (uint32_t v : k.w) { h ^= v; h *= 1099511628211ull; } return (size_t)h; }
};

} // namespace

// MikkTSpace, restated from the published algorithm (mikktspace.c, Morten S. Mikkelsen) for triangle lists:
//   1. weld: corners with identical (position, normal, texcoord) share one vertex id (GenerateSharedVerticesIndexList);
//   2. per triangle (InitTriInfo): d1 = p1 - p0, d2 = p2 - p0, t21 = uv1 - uv0, t31 = uv2 - uv0, area2 = t21.x t31.y - t21.y t31.x,
//      vOs = t31.y d1 - t21.y d2 normalised and signed by the orientation (area2 > 0 = orientation preserving); a triangle whose uv area
//      or |vOs| / |vOt| magnitudes vanish "groups with any"; triangles with two equal positions are degenerate;
//   3. neighbours across edges of equal welded ids; per welded vertex the incident triangles form groups = fans connected through those
//      edges with equal orientation (Build4RuleGroups);
//   4. per corner (GenerateTSpaces / EvalTspace): over the members of its group whose projected vOs / vOt are within the angular
//      threshold (cos > -1) of this corner's, sum (angle at the vertex between the two projected edges) x (vOs projected into the plane of
//      the vertex normal, normalised), then normalise; sign = +1 if orientation preserving else -1;
//   5. degenerate triangles copy the tangent of a non-degenerate triangle at the same welded vertex, else keep the default (1, 0, 0), -1.
auto mikk_tangents(const float* positions, const float* normals, const float* texcoords, float* tangents,
                   const uint32_t* indices, size_t num_indices, uint32_t base_vertex) -> void {
    const size_t nf = num_indices / 3;
    auto P = [&](uint32_t v) { const float* p = positions + 3 * (size_t)(base_vertex + v); return V3{p[0], p[1], p[2]}; };
    auto N = [&](uint32_t v) { const float* p = normals + 3 * (size_t)(base_vertex + v); return V3{p[0], p[1], p[2]}; };
    // 1. weld (float ==: -0 equals +0)
    std::vector<uint32_t> weld(nf * 3);
    {
        std::unordered_map<VertexKey, uint32_t, VertexKeyHash> ids;
        for (size_t c = 0; c < nf * 3; c++) {
            const size_t v = (size_t)base_vertex + indices[c];
            float f[8] = {positions[3 * v], positions[3 * v + 1], positions[3 * v + 2], normals[3 * v], normals[3 * v + 1], normals[3 * v + 2], texcoords[2 * v], texcoords[2 * v + 1]};
            VertexKey k;
            for (int i = 0; i < 8; i++) { if (f[i] == 0.0f) f[i] = 0.0f; std::memcpy(&k.w[i], &f[i], 4); }
            weld[c] = ids.try_emplace(k, (uint32_t)ids.size()).first->second;
        }
    }
    // 2. per-triangle tangent, orientation, flags
    struct Tri { V3 os, ot; bool orient = false, any = true, degenerate = false; int nb[3] = {-1, -1, -1}; };
    std::vector<Tri> tri(nf);
    for (size_t f = 0; f < nf; f++) {
        const uint32_t i0 = indices[3 * f], i1 = indices[3 * f + 1], i2 = indices[3 * f + 2];
        const V3 p0 = P(i0), p1 = P(i1), p2 = P(i2);
        auto same = [](V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
        Tri& t = tri[f];
        t.degenerate = same(p0, p1) || same(p0, p2) || same(p1, p2);
        const float* u0 = texcoords + 2 * (size_t)(base_vertex + i0); const float* u1 = texcoords + 2 * (size_t)(base_vertex + i1); const float* u2 = texcoords + 2 * (size_t)(base_vertex + i2);
        const float t21x = u1[0] - u0[0], t21y = u1[1] - u0[1], t31x = u2[0] - u0[0], t31y = u2[1] - u0[1];
        const V3 d1 = vsub(p1, p0), d2 = vsub(p2, p0);
        const float area2 = t21x * t31y - t21y * t31x;
        t.os = vsub(vscale(t31y, d1), vscale(t21y, d2));
        t.ot = vadd(vscale(-t31x, d1), vscale(t21x, d2));
        t.orient = area2 > 0;
        if (not_zero(area2)) {
            const float abs_area = std::fabs(area2), len_os = vlen(t.os), len_ot = vlen(t.ot), s = t.orient ? 1.0f : -1.0f;
            if (not_zero(len_os)) t.os = vscale(s / len_os, t.os);
            if (not_zero(len_ot)) t.ot = vscale(s / len_ot, t.ot);
            if (not_zero(len_os / abs_area) && not_zero(len_ot / abs_area)) t.any = false;
        }
    }
    // 3. neighbours: an edge (unordered pair of welded ids) links the first two non-degenerate triangles that own it
    {
        std::map<std::pair<uint32_t, uint32_t>, std::pair<int, int>> open;    // edge -> (face, edge slot) still waiting for a partner
        for (size_t f = 0; f < nf; f++) {
            if (tri[f].degenerate) continue;
            for (int e = 0; e < 3; e++) {
                uint32_t a = weld[3 * f + e], b = weld[3 * f + (e + 1) % 3];
                if (a > b) std::swap(a, b);
                auto it = open.find({a, b});
                if (it == open.end()) open[{a, b}] = {(int)f, e};
                else if (it->second.first >= 0) { tri[f].nb[e] = it->second.first; tri[(size_t)it->second.first].nb[it->second.second] = (int)f; it->second.first = -1; }
            }
        }
    }
    // groups: union-find over corners; corner c = 3 f + i joins the corner of the neighbouring face that holds the same welded vertex
    std::vector<uint32_t> parent(nf * 3);
    for (size_t c = 0; c < parent.size(); c++) parent[c] = (uint32_t)c;
    std::function<uint32_t(uint32_t)> find = [&](uint32_t c) { while (parent[c] != c) { parent[c] = parent[parent[c]]; c = parent[c]; } return c; };
    for (size_t f = 0; f < nf; f++) {
        if (tri[f].degenerate) continue;
        for (int i = 0; i < 3; i++) {
            const int edges[2] = {i, (i + 2) % 3};                          // the two edges of face f that meet at corner i
            for (int e : edges) {
                const int g = tri[f].nb[e];
                if (g < 0) continue;
                if (!(tri[f].any || tri[(size_t)g].any || tri[f].orient == tri[(size_t)g].orient)) continue;
                for (int j = 0; j < 3; j++) if (weld[3 * (size_t)g + j] == weld[3 * f + i]) {
                    const uint32_t a = find((uint32_t)(3 * f + i)), b = find((uint32_t)(3 * (size_t)g + j));
                    if (a != b) parent[std::max(a, b)] = std::min(a, b);
                }
            }
        }
    }
    std::unordered_map<uint32_t, std::vector<uint32_t>> members;             // group root -> corners, ascending face order
    for (size_t f = 0; f < nf; f++) if (!tri[f].degenerate) for (int i = 0; i < 3; i++) members[find((uint32_t)(3 * f + i))].push_back((uint32_t)(3 * f + i));
    // 4. per corner
    struct TSpace { V3 os{1.0f, 0.0f, 0.0f}; bool orient = false; bool set = false; };
    std::vector<TSpace> corner(nf * 3);
    std::unordered_map<uint32_t, TSpace> by_weld;                           // for the degenerate epilogue
    for (size_t f = 0; f < nf; f++) {
        if (tri[f].degenerate) continue;
        for (int i = 0; i < 3; i++) {
            const uint32_t c = (uint32_t)(3 * f + i);
            auto& grp = members[find(c)];
            const V3 n = N(indices[c]);
            const V3 os_f = project_normalize(tri[f].os, n), ot_f = project_normalize(tri[f].ot, n);
            V3 sum{0.0f, 0.0f, 0.0f};
            bool group_orient = tri[f].orient;
            for (uint32_t mc : grp) if (!tri[mc / 3].any) { group_orient = tri[mc / 3].orient; break; }
            for (uint32_t mc : grp) {
                const size_t t = mc / 3; const int ti = (int)(mc % 3);
                const V3 os_t = project_normalize(tri[t].os, n), ot_t = project_normalize(tri[t].ot, n);
                const bool any = tri[f].any || tri[t].any;
                if (!(any || t == f || (vdot(os_f, os_t) > -1.0f && vdot(ot_f, ot_t) > -1.0f))) continue;
                if (tri[t].any) continue;                                   // EvalTspace skips members without a usable tangent
                const V3 nt = N(indices[mc]);
                const V3 os = project_normalize(tri[t].os, nt);
                const V3 pa = P(indices[3 * t + (size_t)(ti > 0 ? ti - 1 : 2)]), pb = P(indices[mc]), pc = P(indices[3 * t + (size_t)(ti < 2 ? ti + 1 : 0)]);
                const V3 v1 = project_normalize(vsub(pa, pb), nt), v2 = project_normalize(vsub(pc, pb), nt);
                float cosine = vdot(v1, v2);
                cosine = cosine > 1.0f ? 1.0f : (cosine < -1.0f ? -1.0f : cosine);
                const float angle = (float)std::acos((double)cosine);
                sum = vadd(sum, vscale(angle, os));
            }
            if (vnot_zero(sum)) sum = vnormalize(sum);
            corner[c].os = sum; corner[c].orient = group_orient; corner[c].set = true;
            by_weld.try_emplace(weld[c], corner[c]);
        }
    }
    // 5. degenerate epilogue + output in face order (last corner written to a mesh vertex wins)
    for (size_t f = 0; f < nf; f++) for (int i = 0; i < 3; i++) {
        const uint32_t c = (uint32_t)(3 * f + i);
        TSpace ts = corner[c];
        if (!ts.set) if (auto it = by_weld.find(weld[c]); it != by_weld.end()) ts = it->second;
        float* o = tangents + 4 * (size_t)(base_vertex + indices[c]);
        o[0] = ts.os.x; o[1] = ts.os.y; o[2] = ts.os.z; o[3] = ts.orient ? 1.0f : -1.0f;
    }
}

namespace {

// Transform (math/transform.hpp): rotation is a column-major float3x3, matrix() = [R0 s.x | R1 s.y | R2 s.z | t]
struct Xform { float r[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}; float s[3] = {1, 1, 1}; float t[3] = {0, 0, 0}; };   // r[col][row]
struct Mat4 { float m[4][4]; };                                            // m[col][row]
auto matrix_of(Xform const& x) -> Mat4 {
    Mat4 o{};
    for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) o.m[c][r] = x.r[c][r] * x.s[c];
    for (int r = 0; r < 3; r++) o.m[3][r] = x.t[r];
    o.m[3][3] = 1.0f;
    return o;
}
auto mat_mul(Mat4 const& a, Mat4 const& b) -> Mat4 {                        // glm: column c = a0 b[c].x + a1 b[c].y + a2 b[c].z + a3 b[c].w
    Mat4 o{};
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) o.m[c][r] = ((a.m[0][r] * b.m[c][0] + a.m[1][r] * b.m[c][1]) + a.m[2][r] * b.m[c][2]) + a.m[3][r] * b.m[c][3];
    return o;
}
auto from_matrix(Mat4 const& m) -> Xform {                                  // transform.cpp:41-52
    Xform x;
    for (int r = 0; r < 3; r++) x.t[r] = m.m[3][r];
    for (int c = 0; c < 3; c++) {
        for (int r = 0; r < 3; r++) x.r[c][r] = m.m[c][r];
        x.s[c] = std::sqrt((x.r[c][0] * x.r[c][0] + x.r[c][1] * x.r[c][1]) + x.r[c][2] * x.r[c][2]);
        if (x.s[c] != 0.0f) for (int r = 0; r < 3; r++) x.r[c][r] /= x.s[c];
    }
    return x;
}
auto set_rotation_with_quaternion(Xform& x, const float q[4]) -> void {     // transform.cpp:11-18 (q = x, y, z, w)
    float m[4][4];
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) m[a][b] = (q[b] * q[a]) * 2.0f;
    x.r[0][0] = 1.0f - (m[1][1] + m[2][2]); x.r[0][1] = m[0][1] + m[3][2]; x.r[0][2] = m[0][2] - m[3][1];
    x.r[1][0] = m[0][1] - m[3][2]; x.r[1][1] = 1.0f - (m[0][0] + m[2][2]); x.r[1][2] = m[1][2] + m[3][0];
    x.r[2][0] = m[0][2] + m[3][1]; x.r[2][1] = m[1][2] - m[3][0]; x.r[2][2] = 1.0f - (m[0][0] + m[1][1]);
}

constexpr int GLTF_FLOAT = 5126, GLTF_U8 = 5121, GLTF_U16 = 5123, GLTF_U32 = 5125, GLTF_TRIANGLES = 4;

struct Model {
    Json root;
    std::vector<std::string> buffers;
    std::string base_dir;
    // raw bytes of an accessor: pointer to its first element, element stride in bytes (0 = tightly packed), count
    auto accessor(int index, const unsigned char*& data, size_t& stride, size_t& count, int& component_type, std::string& type, size_t elem_bytes_hint, std::string& err) const -> bool {
        auto& accs = root.array_of("accessors");
        if (index < 0 || (size_t)index >= accs.size()) { err = "accessor index out of range"; return false; }
        auto& a = accs[(size_t)index];
        const int view = a.int_or("bufferView", -1);
        auto& views = root.array_of("bufferViews");
        if (view < 0 || (size_t)view >= views.size()) { err = "accessor without a bufferView (sparse / zero-filled accessors are not supported)"; return false; }
        auto& bv = views[(size_t)view];
        const int buf = bv.int_or("buffer", -1);
        if (buf < 0 || (size_t)buf >= buffers.size()) { err = "bufferView.buffer out of range"; return false; }
        const size_t off = a.size_or("byteOffset", 0) + bv.size_or("byteOffset", 0);
        stride = bv.size_or("byteStride", 0);
        count = a.size_or("count", 0);
        component_type = a.int_or("componentType", 0);
        type = a.string_or("type", "");
        const size_t step = stride ? stride : elem_bytes_hint, have = buffers[(size_t)buf].size();
        if (stride > 4096 || off > have || count > have) { err = "accessor reads past the end of its buffer"; return false; }      // keeps the products below from wrapping
        if (count && off + (count - 1) * step + elem_bytes_hint > have) { err = "accessor reads past the end of its buffer"; return false; }
        data = reinterpret_cast<const unsigned char*>(buffers[(size_t)buf].data()) + off;
        return true;
    }
};

auto unique_name(std::string name, const char* fallback, size_t i, std::unordered_set<std::string>& used) -> std::string {
    if (name.empty() || used.count(name)) name = std::string(fallback) + std::to_string(i);   // import_model.cpp:72-76 and siblings
    used.insert(name);
    return name;
}

} // namespace

auto import_gltf(std::string const& path, Project& out, std::string& err) -> bool {
    Model model;
    std::string file;
    if (!read_binary(path, file)) { err = "cannot read '" + path + "'"; return false; }
    auto slash = path.find_last_of("/\\");
    model.base_dir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
    std::string json_text, glb_bin;
    bool is_glb = file.size() >= 12 && !std::memcmp(file.data(), "glTF", 4);
    if (is_glb) {                                                          // GLB container: header, JSON chunk, optional BIN chunk
        auto u32 = [&](size_t o) { uint32_t v; std::memcpy(&v, file.data() + o, 4); return v; };
        if (u32(4) != 2) { err = path + ": GLB version " + std::to_string(u32(4)) + " is not supported"; return false; }
        size_t o = 12;
        while (o + 8 <= file.size()) {
            const uint32_t len = u32(o), type = u32(o + 4);
            if (o + 8 + (size_t)len > file.size()) { err = path + ": truncated GLB chunk"; return false; }
            if (type == 0x4E4F534Au) json_text.assign(file, o + 8, len); else if (type == 0x004E4942u && glb_bin.empty()) glb_bin.assign(file, o + 8, len);
            o += 8 + (size_t)len;
        }
    } else json_text = std::move(file);
    JsonParser jp{json_text.data(), json_text.data() + json_text.size(), {}, 0};
    if (!jp.value(model.root) || model.root.kind != Json::Kind::object) { err = path + ": " + (jp.err.empty() ? "not a JSON object" : jp.err); return false; }
    const Json& root = model.root;

    for (size_t i = 0; auto& b : root.array_of("buffers")) {
        model.buffers.emplace_back();
        auto uri = b.get("uri");
        if (!uri) { if (is_glb && i == 0) model.buffers.back() = glb_bin; else { err = path + ": buffer " + std::to_string(i) + " has no uri"; return false; } }
        else if (std::string e; !load_uri(uri->str, model.base_dir, model.buffers.back(), e)) { err = path + ": buffer " + std::to_string(i) + ": " + e; return false; }
        if (model.buffers.back().size() < b.size_or("byteLength", 0)) { err = path + ": buffer " + std::to_string(i) + " is shorter than its byteLength"; return false; }
        ++i;
    }

    // ---- textures (import_model.cpp:60-142)
    const int32_t tex_base = (int32_t)out.textures.size();
    auto& images = root.array_of("images");
    auto& samplers = root.array_of("samplers");
    for (size_t i = 0; auto& gt : root.array_of("textures")) {
        TextureData t;
        t.depth = 1; t.levels = 1; t.format = 37; t.dim = 1;               // rgba8_unorm 2D
        t.mag_filter = 1; t.min_filter = 1; t.address_u = 0; t.address_v = 0;   // linear / repeat defaults (import_model.cpp:110-117)
        const int src = gt.int_or("source", -1);
        if (src >= 0 && (size_t)src < images.size()) {
            auto& img = images[(size_t)src];
            std::string bytes, e;
            if (auto uri = img.get("uri")) { if (!load_uri(uri->str, model.base_dir, bytes, e)) { err = path + ": image " + std::to_string(src) + ": " + e; return false; } }
            else {
                const int view = img.int_or("bufferView", -1);
                auto& views = root.array_of("bufferViews");
                if (view < 0 || (size_t)view >= views.size()) { err = path + ": image " + std::to_string(src) + " has neither uri nor bufferView"; return false; }
                auto& bv = views[(size_t)view];
                const int buf = bv.int_or("buffer", -1);
                const size_t off = bv.size_or("byteOffset", 0), len = bv.size_or("byteLength", 0);
                if (buf < 0 || (size_t)buf >= model.buffers.size() || off + len > model.buffers[(size_t)buf].size()) { err = path + ": image " + std::to_string(src) + ": bad bufferView"; return false; }
                bytes.assign(model.buffers[(size_t)buf], off, len);
            }
            if (uint32_t ch = 0; !decode_png(bytes, t.width, t.height, ch, t.texels, true, e)) { err = path + ": image " + std::to_string(src) + ": " + e; return false; }
        } else { t.width = t.height = 1; t.texels = {255, 255, 255, 255}; } // the importer's 1x1 white default image (import_model.cpp:63-69)
        const int smp = gt.int_or("sampler", -1);
        if (smp >= 0 && (size_t)smp < samplers.size()) {                   // import_model.cpp:118-139
            auto& s = samplers[(size_t)smp];
            if (s.int_or("magFilter", -1) == 9728) t.mag_filter = 0;
            const int minf = s.int_or("minFilter", -1);
            if (minf == 9728 || minf == 9984) t.min_filter = 0;
            auto wrap = [&](const char* key, uint8_t& mode) {
                const int w = s.int_or(key, 10497);
                if (w == 33071) mode = 2; else if (w == 33648) mode = 1;    // clamp_to_edge / mirror_repeat (rhi/sampler.hpp:20-26)
            };
            wrap("wrapS", t.address_u); wrap("wrapT", t.address_v);
            if (t.address_u == 1 || t.address_v == 1) { err = path + ": texture " + std::to_string(i) + ": mirrored-repeat addressing is not offered by the CUDA material sampler"; return false; }
        }
        out.textures.push_back(std::move(t));
        ++i;
    }

    // ---- materials (import_model.cpp:144-232): the metallic-roughness template, blend mode opaque
    const uint32_t mat_base = (uint32_t)out.materials.size();
    const size_t num_textures = root.array_of("textures").size();
    for (size_t i = 0; auto& gm : root.array_of("materials")) {
        bpt_material m{};
        Json none; none.kind = Json::Kind::object;
        const Json& pbr = gm.get("pbrMetallicRoughness") ? *gm.get("pbrMetallicRoughness") : none;
        auto factor = [](Json const& o, const char* key, float* dst, int n, float dflt) {
            auto& a = o.array_of(key);
            for (int k = 0; k < n; k++) dst[k] = (size_t)k < a.size() && a[(size_t)k].kind == Json::Kind::number ? (float)a[(size_t)k].num : dflt;
        };
        bool bad_tex = false;
        auto tex_index = [&](Json const& o, const char* key) -> int32_t {
            auto t = o.get(key);
            const int idx = t ? t->int_or("index", -1) : -1;
            if (idx >= 0 && (size_t)idx >= num_textures) bad_tex = true;
            return idx >= 0 ? tex_base + idx : -1;                          // -1 = white1x1 / normal1x1 (import_model.cpp:168-185)
        };
        factor(pbr, "baseColorFactor", m.base_color, 4, 1.0f);
        factor(gm, "emissiveFactor", m.emission, 3, 0.0f);
        m.roughness = (float)pbr.number_or("roughnessFactor", 1.0);
        m.metallic = (float)pbr.number_or("metallicFactor", 1.0);
        m.normal_map_scale = gm.get("normalTexture") ? (float)gm.get("normalTexture")->number_or("scale", 1.0) : 1.0f;
        m.occlusion_strength = gm.get("occlusionTexture") ? (float)gm.get("occlusionTexture")->number_or("strength", 1.0) : 1.0f;
        m.base_color_tex = tex_index(pbr, "baseColorTexture");
        m.metallic_roughness_tex = tex_index(pbr, "metallicRoughnessTexture");
        m.normal_map_tex = tex_index(gm, "normalTexture");
        m.occlusion_tex = tex_index(gm, "occlusionTexture");
        if (bad_tex) { err = path + ": material " + std::to_string(i) + " references a texture that does not exist"; return false; }
        m.flags = (gm.bool_or("doubleSided", false) ? BPT_MATERIAL_FLAG_TWO_SIDED : 0u) | ((uint32_t)BPT_MATERIAL_KIND_GLTF_PBR << BPT_MATERIAL_KIND_SHIFT)
                | ((uint32_t)BPT_BLEND_OPAQUE << BPT_MATERIAL_BLEND_SHIFT) | ((uint32_t)BPT_SURFACE_MODEL_LIT << BPT_MATERIAL_MODEL_SHIFT);
        out.materials.push_back(m);
        ++i;
    }
    const size_t num_materials = root.array_of("materials").size();

    // ---- meshes (import_model.cpp:234-358)
    struct MeshEntry { std::vector<uint32_t> blas; std::vector<int> material; std::vector<uint32_t> vertex_base, index_base; bool present = false; };
    std::vector<MeshEntry> meshes(root.array_of("meshes").size());
    for (size_t mi = 0; auto& gmesh : root.array_of("meshes")) {
        MeshEntry& me = meshes[mi];
        const std::string where = path + ": mesh " + std::to_string(mi);
        const size_t vbase = out.positions.size() / 3, ibase = out.indices.size();
        size_t num_vertices = 0, num_indices = 0;
        struct Sub { uint32_t base_vertex, index_offset, num_indices; };
        std::vector<Sub> subs;
        for (auto& prim : gmesh.array_of("primitives")) {
            if (prim.int_or("mode", GLTF_TRIANGLES) != GLTF_TRIANGLES) continue;
            Json none; none.kind = Json::Kind::object;
            const Json& attrs = prim.get("attributes") ? *prim.get("attributes") : none;
            Sub sub{(uint32_t)num_vertices, (uint32_t)num_indices, 0};
            const unsigned char* data; size_t stride, count; int ctype; std::string type, e;
            // indices: the reference dereferences accessors[prim.indices] unconditionally, a non-indexed primitive is undefined there
            const int iacc = prim.int_or("indices", -1);
            if (iacc < 0) { err = where + ": non-indexed primitives are not importable (import_model.cpp:277 reads accessors[indices])"; return false; }
            if (!model.accessor(iacc, data, stride, count, ctype, type, 1, e)) { err = where + ": indices: " + e; return false; }
            const size_t isz = ctype == GLTF_U32 ? 4 : ctype == GLTF_U16 ? 2 : ctype == GLTF_U8 ? 1 : 0;
            if (!isz) { err = where + ": index component type " + std::to_string(ctype) + " is not an unsigned integer"; return false; }
            if (!model.accessor(iacc, data, stride, count, ctype, type, isz, e)) { err = where + ": indices: " + e; return false; }
            // glTF 2.0 forbids byteStride on index bufferViews; the bounds check above used `stride` as the step, so the reads below use
            // that same step (a hostile byteStride smaller than the element would otherwise pass the check and over-read the buffer)
            if (stride != 0 && stride < isz) { err = where + ": indices: bufferView.byteStride smaller than the index size"; return false; }
            const size_t istep = stride ? stride : isz;
            for (size_t k = 0; k < count; k++) {
                uint32_t v = 0;
                const unsigned char* q = data + istep * k;
                if (isz == 4) std::memcpy(&v, q, 4); else if (isz == 2) { uint16_t h; std::memcpy(&h, q, 2); v = h; } else v = q[0];
                out.indices.push_back(v);
            }
            sub.num_indices = (uint32_t)count; num_indices += count;
            // attributes: float only, a missing one is zero-filled (import_model.cpp:323-346)
            size_t sub_vertices = 0;
            auto add_attribute = [&](const char* name, std::vector<float>& dst, size_t comps, bool defines_count) -> bool {
                const int acc = attrs.int_or(name, -1);
                if (acc < 0) {
                    if (defines_count) { err = where + ": primitive without POSITION"; return false; }
                    dst.resize(dst.size() + sub_vertices * comps, 0.0f);
                    return true;
                }
                if (!model.accessor(acc, data, stride, count, ctype, type, comps * 4, e)) { err = where + ": " + name + ": " + e; return false; }
                if (ctype != GLTF_FLOAT) { err = where + ": " + name + " must be FLOAT (import_model.cpp:337 asserts it)"; return false; }
                if (!defines_count && count != sub_vertices) { err = where + ": " + name + " has " + std::to_string(count) + " elements, POSITION has " + std::to_string(sub_vertices); return false; }
                const size_t step = stride ? stride : comps * 4;
                for (size_t k = 0; k < count; k++) { float v[4]; std::memcpy(v, data + k * step, comps * 4); dst.insert(dst.end(), v, v + comps); }
                if (defines_count) sub_vertices = count;
                return true;
            };
            if (!add_attribute("POSITION", out.positions, 3, true)) return false;
            num_vertices += sub_vertices;
            if (!add_attribute("NORMAL", out.normals, 3, false)) return false;
            if (!add_attribute("TEXCOORD_0", out.texcoords, 2, false)) return false;
            for (size_t k = out.indices.size() - sub.num_indices; k < out.indices.size(); k++)
                if (out.indices[k] >= sub_vertices) { err = where + ": index " + std::to_string(out.indices[k]) + " out of range"; return false; }
            const int mat = prim.int_or("material", -1);
            if (mat < 0 || (size_t)mat >= num_materials) { err = where + ": primitive without a valid material (import_model.cpp:402 reads mat_ids[material])"; return false; }
            me.material.push_back(mat);
            subs.push_back(sub);
        }
        if (num_vertices != 0) {
            me.present = true;
            out.tangents.resize(out.positions.size() / 3 * 4, 0.0f);
            for (auto& sub : subs) {                                        // StaticMesh::calculate_tspace, one MikkTSpace run per submesh
                mikk_tangents(out.positions.data() + 3 * vbase, out.normals.data() + 3 * vbase, out.texcoords.data() + 2 * vbase, out.tangents.data() + 4 * vbase,
                              out.indices.data() + ibase + sub.index_offset, sub.num_indices, sub.base_vertex);
                bpt_blas_desc bd{};                                         // graphics_manager.cpp:616-654: one BLAS per (mesh, submesh)
                bd.position_offset = (uint32_t)(vbase + sub.base_vertex) * 3; bd.index_offset = (uint32_t)(ibase + sub.index_offset); bd.num_triangles = sub.num_indices / 3;
                me.blas.push_back((uint32_t)out.blas.size());
                me.vertex_base.push_back((uint32_t)(vbase + sub.base_vertex)); me.index_base.push_back((uint32_t)(ibase + sub.index_offset));
                out.blas.push_back(bd);
            }
        }
        ++mi;
    }

    // ---- nodes (import_model.cpp:360-421) under an identity base object
    auto& nodes = root.array_of("nodes");
    std::unordered_set<std::string> used_names;
    size_t num_nodes = 0;
    bool ok = true;
    std::function<void(Xform const&, size_t, int)> process_node = [&](Xform const& parent_world, size_t ni, int depth) {
        if (!ok) return;
        if (ni >= nodes.size() || depth > 256 || num_nodes > (1u << 20)) { err = path + ": bad node hierarchy"; ok = false; return; }   // cycles / DAG blow-up
        const Json& node = nodes[ni];
        const std::string name = unique_name(node.string_or("name", ""), "node", num_nodes, used_names);
        ++num_nodes;
        Xform local;
        auto vec = [&](const char* key, float* dst, size_t n) { auto& a = node.array_of(key); if (a.size() >= n) for (size_t k = 0; k < n; k++) dst[k] = (float)a[k].num; return a.size() >= n; };
        vec("translation", local.t, 3);
        float q[4];
        if (vec("rotation", q, 4)) set_rotation_with_quaternion(local, q);
        vec("scale", local.s, 3);
        if (auto& a = node.array_of("matrix"); a.size() == 16) {
            Mat4 m{};
            for (size_t k = 0; k < 16; k++) m.m[k / 4][k % 4] = (float)a[k].num;
            local = from_matrix(m);
        }
        const Xform world = from_matrix(mat_mul(matrix_of(parent_world), matrix_of(local)));    // scene_object.cpp:49, transform.cpp:100-102
        const int mesh = node.int_or("mesh", -1);
        if (mesh >= 0) {
            if ((size_t)mesh >= meshes.size()) { err = path + ": node " + std::to_string(ni) + " references a mesh that does not exist"; ok = false; return; }
            const MeshEntry& me = meshes[(size_t)mesh];
            for (size_t k = 0; me.present && k < me.blas.size(); k++) {
                const uint32_t index = (uint32_t)out.drawables.size(), vb = me.vertex_base[k];
                bpt_drawable_sbt_data dr{};                                 // drawable_stb_data.hpp:7-17
                dr.drawable_index = index; dr.position_offset = vb * 3; dr.normal_offset = vb * 3; dr.tangent_offset = vb * 4; dr.texcoord_offset = vb * 2;
                dr.index_offset = me.index_base[k]; dr.material_offset = (mat_base + (uint32_t)me.material[k]) * (uint32_t)sizeof(bpt_material);
                out.drawables.push_back(dr);
                out.drawable_va.push_back(BPT_VA_POSITION | BPT_VA_NORMAL | BPT_VA_TANGENT | BPT_VA_TEXCOORD);
                bpt_instance_desc in{};                                     // accel.cpp:104-132
                for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) in.transform[r][c] = world.r[c][r] * world.s[c]; in.transform[r][3] = world.t[r]; }
                in.instance_id_and_mask = index | (0xffu << 24);
                in.sbt_offset_and_flags = index | ((uint32_t)BPT_INSTANCE_FORCE_OPAQUE << 24);
                in.blas = me.blas[k];
                out.instances.push_back(in);
                out.object_names.push_back(name);
            }
        }
        for (auto& ch : node.array_of("children")) if (ch.kind == Json::Kind::number) process_node(world, Json::to_size(ch.num), depth + 1);
    };
    auto& scenes = root.array_of("scenes");
    if (scenes.empty()) { err = path + ": no scenes"; return false; }
    const size_t scene_index = (size_t)std::max(0, root.int_or("scene", -1));   // scenes[max(0, defaultScene)]
    if (scene_index >= scenes.size()) { err = path + ": default scene out of range"; return false; }
    const Xform base_world = from_matrix(mat_mul(matrix_of(Xform{}), matrix_of(Xform{})));
    for (auto& n : scenes[scene_index].array_of("nodes")) if (n.kind == Json::Kind::number) process_node(base_world, Json::to_size(n.num), 0);
    if (!ok) return false;
    if (out.drawables.empty()) { err = path + ": no renderable primitive"; return false; }
    return true;
}

} // namespace bi::project
