// oracle_ltc.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
// LTC rect lights: restatement of shaders/renderer/lights.hlsl:164-513 and the caller loop
// shaders/renderer/raytracing/deferred_lighting_secondary.hlsl:72-96, including rect-light textures
// (lights.hlsl:425-447,495-511) on the mip chain of shaders/core/mipmap.hlsl (built in oracle_render.cpp).
#pragma once
#include "oracle_scene.hpp"

namespace orc {

struct m33 { f3 r0, r1, r2; };   // rows
static inline f3 mul(const m33& m, f3 v) { return mk3(dot(m.r0, v), dot(m.r1, v), dot(m.r2, v)); }
static inline m33 mul(const m33& a, const m33& b) {
    f3 c0 = mk3(b.r0.x, b.r1.x, b.r2.x), c1 = mk3(b.r0.y, b.r1.y, b.r2.y), c2 = mk3(b.r0.z, b.r1.z, b.r2.z);
    return m33{mk3(dot(a.r0, c0), dot(a.r0, c1), dot(a.r0, c2)), mk3(dot(a.r1, c0), dot(a.r1, c1), dot(a.r1, c2)),
               mk3(dot(a.r2, c0), dot(a.r2, c1), dot(a.r2, c2))};
}
static inline m33 lerp_m(const m33& a, const m33& b, float w) { return m33{lerp3(a.r0, b.r0, w), lerp3(a.r1, b.r1, w), lerp3(a.r2, b.r2, w)}; }
// core/utils/math.hlsl:33-46
static inline m33 inverse(const m33& m) {
    float m00 = m.r0.x, m01 = m.r0.y, m02 = m.r0.z, m10 = m.r1.x, m11 = m.r1.y, m12 = m.r1.z, m20 = m.r2.x, m21 = m.r2.y, m22 = m.r2.z;
    float det = (m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20)) + m02 * (m10 * m21 - m11 * m20);
    float inv_det = 1.0f / det;
    return m33{mk3((m11 * m22 - m21 * m12) * inv_det, (m21 * m02 - m01 * m22) * inv_det, (m01 * m12 - m11 * m02) * inv_det),
               mk3((m20 * m12 - m10 * m22) * inv_det, (m00 * m22 - m20 * m02) * inv_det, (m10 * m02 - m00 * m12) * inv_det),
               mk3((m10 * m21 - m20 * m11) * inv_det, (m20 * m01 - m00 * m21) * inv_det, (m00 * m11 - m10 * m01) * inv_det)};
}

// 8x8x64 3-D LUT, linear filter, clamp-to-edge, explicit FP32 trilinear.
static inline void lut_trilinear(const float* lut, int channels, f3 u, float out[4]) {
    const int SX = 8, SY = 8, SZ = 64;
    float x = u.x * (float)SX - 0.5f, y = u.y * (float)SY - 0.5f, z = u.z * (float)SZ - 0.5f;
    float x0f = floorf(x), y0f = floorf(y), z0f = floorf(z);
    float fx = x - x0f, fy = y - y0f, fz = z - z0f;
    auto cl = [](int c, int n) { return c < 0 ? 0 : (c >= n ? n - 1 : c); };
    int x0 = cl((int)x0f, SX), x1 = cl((int)x0f + 1, SX), y0 = cl((int)y0f, SY), y1 = cl((int)y0f + 1, SY), z0 = cl((int)z0f, SZ), z1 = cl((int)z0f + 1, SZ);
    for (int c = 0; c < channels; c++) {
        auto at = [&](int xx, int yy, int zz) { return lut[(((size_t)zz * SY + yy) * SX + xx) * channels + c]; };
        float c00 = lerpf(at(x0, y0, z0), at(x1, y0, z0), fx), c10 = lerpf(at(x0, y1, z0), at(x1, y1, z0), fx);
        float c01 = lerpf(at(x0, y0, z1), at(x1, y0, z1), fx), c11 = lerpf(at(x0, y1, z1), at(x1, y1, z1), fx);
        out[c] = lerpf(lerpf(c00, c10, fy), lerpf(c01, c11, fy), fz);
    }
}
// lights.hlsl:164-178
static inline void get_ltc_tex3d_coord(f4 u, f3& u1, f3& u2, float& w) {
    float ws = u.w * 7.0f;
    float ws_f = floorf(ws);
    float ws_c = fmin_(floorf(ws + 1.0f), 7.0f);
    w = ws - floorf(ws);
    float x = (u.x * 7.0f + 0.5f) / 8.0f;
    float y = (u.y * 7.0f + 0.5f) / 8.0f;
    float z1 = ((u.z * 7.0f + 8.0f * ws_f) + 0.5f) / 64.0f;
    float z2 = ((u.z * 7.0f + 8.0f * ws_c) + 0.5f) / 64.0f;
    u1 = mk3(x, y, z1); u2 = mk3(x, y, z2);
}
static inline m33 fetch_fetch_ltc_matrix(const Scene& sc, f3 u) {   // lights.hlsl:179-184
    float a[4], b[4], c[4];
    lut_trilinear(sc.ltc_m0.data(), 4, u, a); lut_trilinear(sc.ltc_m1.data(), 4, u, b); lut_trilinear(sc.ltc_m2.data(), 4, u, c);
    return m33{mk3(a[0], a[1], a[2]), mk3(b[0], b[1], b[2]), mk3(c[0], c[1], c[2])};
}
static inline m33 fetch_ltc_matrix(const Scene& sc, f4 u) {          // lights.hlsl:185-193
    f3 u1, u2; float w;
    get_ltc_tex3d_coord(u, u1, u2, w);
    return lerp_m(fetch_fetch_ltc_matrix(sc, u1), fetch_fetch_ltc_matrix(sc, u2), w);
}
static inline f2 fetch_ltc_brdf(const Scene& sc, f4 u) {             // lights.hlsl:194-201
    f3 u1, u2; float w;
    get_ltc_tex3d_coord(u, u1, u2, w);
    float a[4], b[4];
    lut_trilinear(sc.ltc_norm.data(), 2, u1, a); lut_trilinear(sc.ltc_norm.data(), 2, u2, b);
    return f2{lerpf(a[0], b[0], w), lerpf(a[1], b[1], w)};
}
static inline void wind(f3 L[4]) { f3 t0 = L[0], t1 = L[1]; L[0] = L[3]; L[1] = L[2]; L[2] = t1; L[3] = t0; }  // mul(winding, L)

// lights.hlsl:203-273
static inline void get_ltc_matrix_and_brdf(const Scene& sc, f3 local_v, float rx, float ry, f3 L[4], m33& ltc_matrix, f2& ltc_brdf) {
    float theta_wi = acos_(local_v.z);
    bool flip_roughness = ry > rx;
    float phi_wi = atan2_(local_v.y, local_v.x);
    phi_wi = flip_roughness ? (PI / 2.0f - phi_wi) : phi_wi;
    phi_wi = phi_wi >= 0.0f ? phi_wi : phi_wi + 2.0f * PI;
    float u0 = fmax_((flip_roughness ? ry : rx) - 0.001f, 0.0f) / (1.0f - 0.001f);
    float u1 = flip_roughness ? rx / ry : ry / rx;
    float u2 = theta_wi / (PI * 0.5f);
    if (phi_wi < PI * 0.5f) {
        float u3 = phi_wi / (PI * 0.5f);
        f4 u = f4{u3, u2, u1, u0};
        ltc_matrix = fetch_ltc_matrix(sc, u);
        ltc_brdf = fetch_ltc_brdf(sc, u);
    } else if (phi_wi < PI) {
        float u3 = (PI - phi_wi) / (PI * 0.5f);
        f4 u = f4{u3, u2, u1, u0};
        m33 flip{mk3(-1, 0, 0), mk3(0, 1, 0), mk3(0, 0, 1)};
        wind(L);
        ltc_matrix = mul(flip, fetch_ltc_matrix(sc, u));
        ltc_brdf = fetch_ltc_brdf(sc, u);
    } else if (phi_wi < 1.5f * PI) {
        float u3 = (phi_wi - PI) / (PI * 0.5f);
        f4 u = f4{u3, u2, u1, u0};
        m33 flip{mk3(-1, 0, 0), mk3(0, -1, 0), mk3(0, 0, 1)};
        ltc_matrix = mul(flip, fetch_ltc_matrix(sc, u));
        ltc_brdf = fetch_ltc_brdf(sc, u);
    } else {
        float u3 = (2.0f * PI - phi_wi) / (PI * 0.5f);
        f4 u = f4{u3, u2, u1, u0};
        m33 flip{mk3(1, 0, 0), mk3(0, -1, 0), mk3(0, 0, 1)};
        wind(L);
        ltc_matrix = mul(flip, fetch_ltc_matrix(sc, u));
        ltc_brdf = fetch_ltc_brdf(sc, u);
    }
    if (flip_roughness) {
        m33 flip{mk3(0, 1, 0), mk3(1, 0, 0), mk3(0, 0, 1)};
        wind(L);
        ltc_matrix = mul(flip, ltc_matrix);
    }
}

// lights.hlsl:275-365
static inline void ltc_clip_quad(f3 L[5], int& n) {
    int config = 0;
    if (L[0].z > 0.0f) config += 1;
    if (L[1].z > 0.0f) config += 2;
    if (L[2].z > 0.0f) config += 4;
    if (L[3].z > 0.0f) config += 8;
    n = 0;
    switch (config) {
    case 0: break;
    case 1: n = 3; L[1] = -L[1].z * L[0] + L[0].z * L[1]; L[2] = -L[3].z * L[0] + L[0].z * L[3]; break;
    case 2: n = 3; L[0] = -L[0].z * L[1] + L[1].z * L[0]; L[2] = -L[2].z * L[1] + L[1].z * L[2]; break;
    case 3: n = 4; L[2] = -L[2].z * L[1] + L[1].z * L[2]; L[3] = -L[3].z * L[0] + L[0].z * L[3]; break;
    case 4: n = 3; L[0] = -L[3].z * L[2] + L[2].z * L[3]; L[1] = -L[1].z * L[2] + L[2].z * L[1]; break;
    case 5: n = 0; break;
    case 6: n = 4; L[0] = -L[0].z * L[1] + L[1].z * L[0]; L[3] = -L[3].z * L[2] + L[2].z * L[3]; break;
    case 7: n = 5; L[4] = -L[3].z * L[0] + L[0].z * L[3]; L[3] = -L[3].z * L[2] + L[2].z * L[3]; break;
    case 8: n = 3; L[0] = -L[0].z * L[3] + L[3].z * L[0]; L[1] = -L[2].z * L[3] + L[3].z * L[2]; L[2] = L[3]; break;
    case 9: n = 4; L[1] = -L[1].z * L[0] + L[0].z * L[1]; L[2] = -L[2].z * L[3] + L[3].z * L[2]; break;
    case 10: n = 0; break;
    case 11: n = 5; L[4] = L[3]; L[3] = -L[2].z * L[3] + L[3].z * L[2]; L[2] = -L[2].z * L[1] + L[1].z * L[2]; break;
    case 12: n = 4; L[1] = -L[1].z * L[2] + L[2].z * L[1]; L[0] = -L[0].z * L[3] + L[3].z * L[0]; break;
    case 13: n = 5; L[4] = L[3]; L[3] = L[2]; L[2] = -L[1].z * L[2] + L[2].z * L[1]; L[1] = -L[1].z * L[0] + L[0].z * L[1]; break;
    case 14: n = 5; L[4] = -L[0].z * L[3] + L[3].z * L[0]; L[0] = -L[0].z * L[1] + L[1].z * L[0]; break;
    case 15: n = 4; break;
    }
    if (n == 3) L[3] = L[0];
    if (n == 4) L[4] = L[0];
}
// lights.hlsl:366-382 — returns (cross.xyz, cross.z) * theta/sin(theta)
static inline f4 ltc_integrate_edge(f3 v1, f3 v2) {
    float x = dot(v1, v2);
    float y = fabsf(x);
    float a = 5.42031f + (3.12829f + 0.0902326f * y) * y;
    float b = 3.45068f + (4.18814f + y) * y;
    float tdst = a / b;
    if (x < 0.0f) tdst = PI * (1.0f / sqrtf(1.0f - x * x)) - tdst;
    f3 c = cross(v1, v2);
    return f4{c.x * tdst, c.y * tdst, c.z * tdst, c.z * tdst};
}
// lights.hlsl:383-423; `ltc_matrix` (nullptr = identity) only enters the most representative point
static inline float ltc_integrate(f3 P, f3 N, f3 T, f3 B, const m33& ltc_matrix_inv, const f3 L[4], bool two_sided, f3* mrp = nullptr, const m33* ltc_matrix = nullptr) {
    m33 TBN{T, B, N};
    f3 LP[5];
    for (int k = 0; k < 4; k++) LP[k] = mul(ltc_matrix_inv, mul(TBN, L[k] - P));
    LP[4] = splat3(0.0f);
    int n;
    ltc_clip_quad(LP, n);
    if (n == 0) return 0.0f;
    for (int k = 0; k < 5; k++) LP[k] = normalize(LP[k]);
    f4 sum = ltc_integrate_edge(LP[0], LP[1]);
    auto acc = [&](f4 e) { sum.x += e.x; sum.y += e.y; sum.z += e.z; sum.w += e.w; };
    acc(ltc_integrate_edge(LP[1], LP[2]));
    acc(ltc_integrate_edge(LP[2], LP[3]));
    if (n >= 4) acc(ltc_integrate_edge(LP[3], LP[4]));
    if (n == 5) acc(ltc_integrate_edge(LP[4], LP[0]));
    float integral = two_sided ? fabsf(sum.w) : fmax_(0.0f, sum.w);
    if (!std::isfinite(integral)) integral = 0.0f;
    // lights.hlsl:420: mrp = normalize(mul(mul(ltc_matrix, sum.xyz), TBN)); a row vector times TBN is x*T + y*B + z*N
    if (mrp) {
        f3 v = mk3(sum.x, sum.y, sum.z);
        if (ltc_matrix) v = mul(*ltc_matrix, v);
        *mrp = normalize((v.x * T + v.y * B) + v.z * N);
    }
    return integral;
}

// Texture2D.SampleLevel on the chain: bilinear (or nearest) inside a level, linear or nearest between levels (rhi::SamplerDesc).
static inline f3 light_texture_fetch(const LightTexture& t, int level, float u, float v) {
    const int w = (int)std::max(t.w >> level, 1u), h = (int)std::max(t.h >> level, 1u);
    const float* img = t.level[(size_t)level].data();
    auto wrap = [](int c, int n, uint32_t mode) { if (mode == BPT_ADDRESS_REPEAT) { int m = c % n; return m < 0 ? m + n : m; } return c < 0 ? 0 : (c >= n ? n - 1 : c); };
    auto px = [&](int x, int y) { const float* p = img + ((size_t)y * w + x) * 4; return mk3(p[0], p[1], p[2]); };
    if (!t.linear) return px(wrap((int)floorf(u * (float)w), w, t.addr_u), wrap((int)floorf(v * (float)h), h, t.addr_v));
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap((int)x0f, w, t.addr_u), x1 = wrap((int)x0f + 1, w, t.addr_u), y0 = wrap((int)y0f, h, t.addr_v), y1 = wrap((int)y0f + 1, h, t.addr_v);
    return lerp3(lerp3(px(x0, y0), px(x1, y0), fx), lerp3(px(x0, y1), px(x1, y1), fx), fy);
}
static inline f3 light_texture_sample_level(const LightTexture& t, float u, float v, float level) {
    const int last = (int)t.level.size() - 1;
    level = fmin_(fmax_(level, 0.0f), (float)last);
    if (!t.mip_linear) {
        int l = (int)ceilf(level + 0.5f) - 1;                     // Vulkan: nearest level
        return light_texture_fetch(t, std::min(std::max(l, 0), last), u, v);
    }
    const float base = floorf(level);
    const int l0 = (int)base, l1 = std::min(l0 + 1, last);
    f3 a = light_texture_fetch(t, l0, u, v);
    if (l1 == l0) return a;
    return lerp3(a, light_texture_fetch(t, l1, u, v), level - base);
}
// lights.hlsl:425-447
static inline f3 rect_light_sample_texture(const LightTexture& tex, const bpt_rect_light_data& light, f3 direction, f3 P, float roughness) {
    const f3 n = mk3(light.normal[0], light.normal[1], light.normal[2]);
    const f3 p1 = mk3(light.position1[0], light.position1[1], light.position1[2]), p2 = mk3(light.position2[0], light.position2[1], light.position2[2]),
             p3 = mk3(light.position3[0], light.position3[1], light.position3[2]);
    float step = fabsf(dot(direction, n));
    if (step < 0.0001f) return splat3(0.0f);
    float dist = fabsf(dot(P - p2, n));
    float t = dist / step;
    f3 rect_pos = (P + direction * t) - p2;
    float u = saturate(dot(rect_pos, p3 - p2) * light.inv_width_sqr);
    float v = saturate(dot(rect_pos, p1 - p2) * light.inv_height_sqr);
    float num_texels = (t * roughness) * light.inv_texel_size;
    float level = log2_(fmax_(1.0f, num_texels));
    return light_texture_sample_level(tex, u, v, level);
}

// rect_light_eval_ltc (lights.hlsl:449-513) + surface_eval_lut (deferred_lighting_secondary.hlsl:80-96)
static inline f3 ltc_rect_light(const Scene& sc, const bpt_rect_light_data& light, f3 P, f3 N, f3 T, f3 B, f3 V,
                                const SurfaceData& surf, uint32_t surface_model, f3* diff_mrp = nullptr) {
    float rx, ry;
    get_anisotropic_roughness(surf.roughness, surf.anisotropy, rx, ry);
    f3 local_v = mk3(dot(V, T), dot(V, B), dot(V, N));
    f3 ltc_spec = splat3(0.0f), ltc_diff = splat3(0.0f);
    f2 ltc_brdf{0.0f, 0.0f};
    if (local_v.z > 0.0f) {
        f3 L[4] = {mk3(light.position3[0], light.position3[1], light.position3[2]), mk3(light.position2[0], light.position2[1], light.position2[2]),
                   mk3(light.position1[0], light.position1[1], light.position1[2]), mk3(light.position0[0], light.position0[1], light.position0[2])};
        f3 emission = mk3(light.emission[0], light.emission[1], light.emission[2]);
        m33 identity{mk3(1, 0, 0), mk3(0, 1, 0), mk3(0, 0, 1)};
        f3 mrp_d = splat3(0.0f), mrp_s = splat3(0.0f);
        float integral_diff = ltc_integrate(P, N, T, B, identity, L, light.two_sided != 0, &mrp_d);
        if (diff_mrp) *diff_mrp = mrp_d;
        ltc_diff = emission * integral_diff;
        m33 ltc_matrix;
        get_ltc_matrix_and_brdf(sc, local_v, rx, ry, L, ltc_matrix, ltc_brdf);
        m33 ltc_matrix_inv = inverse(ltc_matrix);
        float integral_spec = ltc_integrate(P, N, T, B, ltc_matrix_inv, L, light.two_sided != 0, &mrp_s, &ltc_matrix);
        ltc_spec = emission * integral_spec;
        if (light.texture_index >= 0 && (size_t)light.texture_index < sc.light_textures.size()) {       // lights.hlsl:495-511
            // (with a zero integral the reference leaves `mrp` unwritten and multiplies 0 by whatever the lookup returns; the term stays 0 here)
            const LightTexture& tex = sc.light_textures[(size_t)light.texture_index];
            ltc_diff = integral_diff != 0.0f ? ltc_diff * rect_light_sample_texture(tex, light, mrp_d, P, 1.0f) : splat3(0.0f);
            ltc_spec = integral_spec != 0.0f ? ltc_spec * rect_light_sample_texture(tex, light, mrp_s, P, sqrtf(rx * ry)) : splat3(0.0f);
        }
    }
    return surface_eval_lut(N, V, surf, ltc_diff, ltc_spec, ltc_brdf, surface_model);
}

} // namespace orc
