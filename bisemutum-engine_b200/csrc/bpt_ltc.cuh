// bpt_ltc.cuh — LTC rect-light evaluation (device).
// Follows bisemutum/shaders/renderer/lights.hlsl:164-513 (LUT addressing, quadrant flips and
// winding, horizon clipping, edge integrals) and the caller in
// shaders/renderer/raytracing/deferred_lighting_secondary.hlsl:72-96. The 8x8x64 LUTs are read
// with explicit FP32 trilinear interpolation from plain global memory (no texture unit: its 8-bit
// filter weights would exceed the 1e-4 parity budget). Light textures (lights.hlsl:425-447,495-511): bpt_scene.cuh: light_texture_sample.
#pragma once
#include "bpt_scene.cuh"

namespace bptd {

struct Mat3 { float3 r0, r1, r2; };
BPT_HD float3 mul_mv(const Mat3& m, float3 v) { return v3(dot3(m.r0, v), dot3(m.r1, v), dot3(m.r2, v)); }
BPT_HD Mat3 mul_mm(const Mat3& a, const Mat3& b) {
    float3 c0 = v3(b.r0.x, b.r1.x, b.r2.x), c1 = v3(b.r0.y, b.r1.y, b.r2.y), c2 = v3(b.r0.z, b.r1.z, b.r2.z);
    Mat3 r;
    r.r0 = v3(dot3(a.r0, c0), dot3(a.r0, c1), dot3(a.r0, c2));
    r.r1 = v3(dot3(a.r1, c0), dot3(a.r1, c1), dot3(a.r1, c2));
    r.r2 = v3(dot3(a.r2, c0), dot3(a.r2, c1), dot3(a.r2, c2));
    return r;
}
BPT_HD Mat3 mat3_inverse(const Mat3& m) {                         // core/utils/math.hlsl:33-46
    float m00 = m.r0.x, m01 = m.r0.y, m02 = m.r0.z, m10 = m.r1.x, m11 = m.r1.y, m12 = m.r1.z, m20 = m.r2.x, m21 = m.r2.y, m22 = m.r2.z;
    float det = (m00 * (m11 * m22 - m12 * m21) - m01 * (m10 * m22 - m12 * m20)) + m02 * (m10 * m21 - m11 * m20);
    float id = 1.0f / det;
    Mat3 r;
    r.r0 = v3((m11 * m22 - m21 * m12) * id, (m21 * m02 - m01 * m22) * id, (m01 * m12 - m11 * m02) * id);
    r.r1 = v3((m20 * m12 - m10 * m22) * id, (m00 * m22 - m20 * m02) * id, (m10 * m02 - m00 * m12) * id);
    r.r2 = v3((m10 * m21 - m20 * m11) * id, (m20 * m01 - m00 * m21) * id, (m00 * m11 - m10 * m01) * id);
    return r;
}

template <int CH>
BPT_HD void lut_fetch(const float* lut, float3 u, float* out) {
    const int SX = 8, SY = 8, SZ = 64;
    float x = u.x * (float)SX - 0.5f, y = u.y * (float)SY - 0.5f, z = u.z * (float)SZ - 0.5f;
    float x0f = floorf(x), y0f = floorf(y), z0f = floorf(z);
    float fx = x - x0f, fy = y - y0f, fz = z - z0f;
    int x0 = wrap_tc((int)x0f, SX, BPT_ADDRESS_CLAMP), x1 = wrap_tc((int)x0f + 1, SX, BPT_ADDRESS_CLAMP);
    int y0 = wrap_tc((int)y0f, SY, BPT_ADDRESS_CLAMP), y1 = wrap_tc((int)y0f + 1, SY, BPT_ADDRESS_CLAMP);
    int z0 = wrap_tc((int)z0f, SZ, BPT_ADDRESS_CLAMP), z1 = wrap_tc((int)z0f + 1, SZ, BPT_ADDRESS_CLAMP);
#pragma unroll
    for (int c = 0; c < (CH == 4 ? 3 : 2); c++) {     // only xyz of the rgba32f matrix LUTs are used
        float a000 = BPT_LDG(lut + (((size_t)z0 * SY + y0) * SX + x0) * CH + c), a100 = BPT_LDG(lut + (((size_t)z0 * SY + y0) * SX + x1) * CH + c);
        float a010 = BPT_LDG(lut + (((size_t)z0 * SY + y1) * SX + x0) * CH + c), a110 = BPT_LDG(lut + (((size_t)z0 * SY + y1) * SX + x1) * CH + c);
        float a001 = BPT_LDG(lut + (((size_t)z1 * SY + y0) * SX + x0) * CH + c), a101 = BPT_LDG(lut + (((size_t)z1 * SY + y0) * SX + x1) * CH + c);
        float a011 = BPT_LDG(lut + (((size_t)z1 * SY + y1) * SX + x0) * CH + c), a111 = BPT_LDG(lut + (((size_t)z1 * SY + y1) * SX + x1) * CH + c);
        float c00 = mix1(a000, a100, fx), c10 = mix1(a010, a110, fx), c01 = mix1(a001, a101, fx), c11 = mix1(a011, a111, fx);
        out[c] = mix1(mix1(c00, c10, fy), mix1(c01, c11, fy), fz);
    }
}
BPT_HD void ltc_coords(float4 u, float3& u1, float3& u2, float& w) {          // lights.hlsl:164-178
    float ws = u.w * 7.0f;
    float ws_f = floorf(ws);
    float ws_c = tmin_(floorf(ws + 1.0f), 7.0f);
    w = ws - floorf(ws);
    float x = (u.x * 7.0f + 0.5f) / 8.0f;
    float y = (u.y * 7.0f + 0.5f) / 8.0f;
    float z1 = ((u.z * 7.0f + 8.0f * ws_f) + 0.5f) / 64.0f;
    float z2 = ((u.z * 7.0f + 8.0f * ws_c) + 0.5f) / 64.0f;
    u1 = v3(x, y, z1); u2 = v3(x, y, z2);
}
BPT_HD Mat3 ltc_matrix_at(const DScene& sc, float3 u) {                        // lights.hlsl:179-184
    float a[3], b[3], c[3];
    lut_fetch<4>(sc.ltc_m0, u, a); lut_fetch<4>(sc.ltc_m1, u, b); lut_fetch<4>(sc.ltc_m2, u, c);
    Mat3 m; m.r0 = v3(a[0], a[1], a[2]); m.r1 = v3(b[0], b[1], b[2]); m.r2 = v3(c[0], c[1], c[2]);
    return m;
}
BPT_HD Mat3 ltc_matrix_lerp(const DScene& sc, float4 u) {                      // lights.hlsl:185-193
    float3 u1, u2; float w;
    ltc_coords(u, u1, u2, w);
    Mat3 a = ltc_matrix_at(sc, u1), b = ltc_matrix_at(sc, u2), r;
    r.r0 = mix3(a.r0, b.r0, w); r.r1 = mix3(a.r1, b.r1, w); r.r2 = mix3(a.r2, b.r2, w);
    return r;
}
BPT_HD float2 ltc_brdf_lerp(const DScene& sc, float4 u) {                      // lights.hlsl:194-201
    float3 u1, u2; float w;
    ltc_coords(u, u1, u2, w);
    float a[2], b[2];
    lut_fetch<2>(sc.ltc_norm, u1, a); lut_fetch<2>(sc.ltc_norm, u2, b);
    return make_float2(mix1(a[0], b[0], w), mix1(a[1], b[1], w));
}
BPT_HD void rewind(float3* L) { float3 t0 = L[0], t1 = L[1]; L[0] = L[3]; L[1] = L[2]; L[2] = t1; L[3] = t0; }

// What of rect_light_eval_ltc depends on the SURFACE only (view direction in the tangent frame, roughness): LUT coordinates, the
// two LUT fetches + lerps, quadrant flips, the matrix inverse (lights.hlsl:203-273 + :494). It is computed once per path vertex and
// shared by all rect lights (16 on BASELINE configs[2]); per light only the corner winding (`rewind`) and the two integrals remain.
// Same operations in the same order as the per-light form, so the result is unchanged bit for bit.
struct LtcSetup { Mat3 Minv; Mat3 M; float2 brdf; bool wind, flip_roughness; };
BPT_HD LtcSetup ltc_setup(const DScene& sc, float3 lv, float rx, float ry) {
    LtcSetup ls;
    float theta_wi = acos_(lv.z);
    bool flip_roughness = ry > rx;
    float phi_wi = atan2_(lv.y, lv.x);
    phi_wi = flip_roughness ? (kPi / 2.0f - phi_wi) : phi_wi;
    phi_wi = phi_wi >= 0.0f ? phi_wi : phi_wi + 2.0f * kPi;
    float u0 = tmax_((flip_roughness ? ry : rx) - 0.001f, 0.0f) / (1.0f - 0.001f);
    float u1 = flip_roughness ? rx / ry : ry / rx;
    float u2 = theta_wi / (kPi * 0.5f);
    Mat3 flip; flip.r0 = v3(1, 0, 0); flip.r1 = v3(0, 1, 0); flip.r2 = v3(0, 0, 1);
    float u3; bool do_flip = true, do_wind = false;
    if (phi_wi < kPi * 0.5f) { u3 = phi_wi / (kPi * 0.5f); do_flip = false; }
    else if (phi_wi < kPi) { u3 = (kPi - phi_wi) / (kPi * 0.5f); flip.r0 = v3(-1, 0, 0); do_wind = true; }
    else if (phi_wi < 1.5f * kPi) { u3 = (phi_wi - kPi) / (kPi * 0.5f); flip.r0 = v3(-1, 0, 0); flip.r1 = v3(0, -1, 0); }
    else { u3 = (2.0f * kPi - phi_wi) / (kPi * 0.5f); flip.r1 = v3(0, -1, 0); do_wind = true; }
    float4 u = make_float4(u3, u2, u1, u0);
    Mat3 M = ltc_matrix_lerp(sc, u);
    if (do_flip) M = mul_mm(flip, M);
    ls.brdf = ltc_brdf_lerp(sc, u);
    if (flip_roughness) {
        Mat3 sw; sw.r0 = v3(0, 1, 0); sw.r1 = v3(1, 0, 0); sw.r2 = v3(0, 0, 1);
        M = mul_mm(sw, M);
    }
    ls.Minv = mat3_inverse(M);
    ls.M = M;
    ls.wind = do_wind; ls.flip_roughness = flip_roughness;
    return ls;
}

BPT_HD float3 clip_mix(float3 a, float3 b) { return -a.z * b + b.z * a; }      // -A.z * B + B.z * A
BPT_HD void ltc_clip(float3* L, int& n) {                                      // lights.hlsl:275-365
    int config = 0;
    if (L[0].z > 0.0f) config += 1;
    if (L[1].z > 0.0f) config += 2;
    if (L[2].z > 0.0f) config += 4;
    if (L[3].z > 0.0f) config += 8;
    n = 0;
    switch (config) {
    case 1: n = 3; L[1] = clip_mix(L[1], L[0]); L[2] = clip_mix(L[3], L[0]); break;
    case 2: n = 3; L[0] = clip_mix(L[0], L[1]); L[2] = clip_mix(L[2], L[1]); break;
    case 3: n = 4; L[2] = clip_mix(L[2], L[1]); L[3] = clip_mix(L[3], L[0]); break;
    case 4: n = 3; L[0] = clip_mix(L[3], L[2]); L[1] = clip_mix(L[1], L[2]); break;
    case 6: n = 4; L[0] = clip_mix(L[0], L[1]); L[3] = clip_mix(L[3], L[2]); break;
    case 7: n = 5; L[4] = clip_mix(L[3], L[0]); L[3] = clip_mix(L[3], L[2]); break;
    case 8: n = 3; L[0] = clip_mix(L[0], L[3]); L[1] = clip_mix(L[2], L[3]); L[2] = L[3]; break;
    case 9: n = 4; L[1] = clip_mix(L[1], L[0]); L[2] = clip_mix(L[2], L[3]); break;
    case 11: n = 5; L[4] = L[3]; L[3] = clip_mix(L[2], L[3]); L[2] = clip_mix(L[2], L[1]); break;
    case 12: n = 4; L[1] = clip_mix(L[1], L[2]); L[0] = clip_mix(L[0], L[3]); break;
    case 13: n = 5; L[4] = L[3]; L[3] = L[2]; L[2] = clip_mix(L[1], L[2]); L[1] = clip_mix(L[1], L[0]); break;
    case 14: n = 5; L[4] = clip_mix(L[0], L[3]); L[0] = clip_mix(L[0], L[1]); break;
    case 15: n = 4; break;
    default: break;   // 0, 5, 10: nothing visible
    }
    if (n == 3) L[3] = L[0];
    if (n == 4) L[4] = L[0];
}
BPT_HD float4 ltc_edge(float3 v1, float3 v2) {                                  // lights.hlsl:366-382
    float x = dot3(v1, v2);
    float y = fabsf(x);
    float a = 5.42031f + (3.12829f + 0.0902326f * y) * y;
    float b = 3.45068f + (4.18814f + y) * y;
    float k = a / b;
    if (x < 0.0f) k = kPi * (1.0f / sqrtf(1.0f - x * x)) - k;
    float3 c = cross3(v1, v2);
    return make_float4(c.x * k, c.y * k, c.z * k, c.z * k);
}
// `M` = the lobe's matrix for the most representative point (nullptr: identity, the diffuse lobe)
BPT_HD float ltc_integrate(float3 P, float3 N, float3 T, float3 B, const Mat3& Minv, const float3* L, bool two_sided, float3* mrp = nullptr, const Mat3* M = nullptr) {   // lights.hlsl:383-423
    Mat3 TBN; TBN.r0 = T; TBN.r1 = B; TBN.r2 = N;
    float3 LP[5];
#pragma unroll
    for (int k = 0; k < 4; k++) LP[k] = mul_mv(Minv, mul_mv(TBN, L[k] - P));
    LP[4] = v3s(0.0f);
    int n;
    ltc_clip(LP, n);
    if (n == 0) return 0.0f;
#pragma unroll
    for (int k = 0; k < 5; k++) LP[k] = normalize3(LP[k]);
    float4 sum = ltc_edge(LP[0], LP[1]);
    float4 e = ltc_edge(LP[1], LP[2]); sum.x += e.x; sum.y += e.y; sum.z += e.z; sum.w += e.w;
    e = ltc_edge(LP[2], LP[3]); sum.x += e.x; sum.y += e.y; sum.z += e.z; sum.w += e.w;
    if (n >= 4) { e = ltc_edge(LP[3], LP[4]); sum.x += e.x; sum.y += e.y; sum.z += e.z; sum.w += e.w; }
    if (n == 5) { e = ltc_edge(LP[4], LP[0]); sum.x += e.x; sum.y += e.y; sum.z += e.z; sum.w += e.w; }
    float integral = two_sided ? fabsf(sum.w) : tmax_(0.0f, sum.w);
    if (!is_finite1(integral)) integral = 0.0f;
    // most representative point direction, lights.hlsl:420: mul(mul(ltc_matrix, sum.xyz), TBN) = x*T + y*B + z*N of the lobe-space vector
    if (mrp) {
        float3 sv = v3(sum.x, sum.y, sum.z);
        if (M) sv = mul_mv(*M, sv);
        *mrp = normalize3((sv.x * T + sv.y * B) + sv.z * N);
    }
    return integral;
}

// rect_light_sample_texture (lights.hlsl:425-447): where the direction `dir` from P meets the light's plane, in the rectangle's (u, v),
// at the level whose texels are as large as the footprint (distance x roughness x texels per metre).
BPT_HD float3 rect_light_texture(const DLightTexture& tex, const bpt_rect_light_data& light, float3 dir, float3 P, float roughness) {
    const float3 ln = v3(light.normal[0], light.normal[1], light.normal[2]);
    const float3 p1 = v3(light.position1[0], light.position1[1], light.position1[2]), p2 = v3(light.position2[0], light.position2[1], light.position2[2]),
                 p3 = v3(light.position3[0], light.position3[1], light.position3[2]);
    float step = fabsf(dot3(dir, ln));
    if (step < 0.0001f) return v3s(0.0f);
    float dist = fabsf(dot3(P - p2, ln));
    float t = dist / step;
    float3 rect_pos = (P + dir * t) - p2;
    float u = sat(dot3(rect_pos, p3 - p2) * light.inv_width_sqr);
    float v = sat(dot3(rect_pos, p1 - p2) * light.inv_height_sqr);
    float num_texels = (t * roughness) * light.inv_texel_size;
    float level = log2_(tmax_(1.0f, num_texels));
    return light_texture_sample(tex, u, v, level);
}

// rect_light_eval_ltc (lights.hlsl:449-513) + surface_eval_lut. `setup` = ltc_setup of this vertex when lv.z > 0 (else unused).
BPT_HD float3 eval_rect_light(const DScene& sc, const bpt_rect_light_data& light, float3 P, float3 N, float3 T, float3 B, float3 V,
                              const Surface& s, uint32_t surface_model, float3 lv, const LtcSetup& setup, float3* diff_mrp = nullptr) {
    float3 spec = v3s(0.0f), diff = v3s(0.0f);
    float2 brdf = make_float2(0.0f, 0.0f);
    if (lv.z > 0.0f) {
        float3 L[4] = {v3(light.position3[0], light.position3[1], light.position3[2]), v3(light.position2[0], light.position2[1], light.position2[2]),
                       v3(light.position1[0], light.position1[1], light.position1[2]), v3(light.position0[0], light.position0[1], light.position0[2])};
        float3 emission = v3(light.emission[0], light.emission[1], light.emission[2]);
        Mat3 I; I.r0 = v3(1, 0, 0); I.r1 = v3(0, 1, 0); I.r2 = v3(0, 0, 1);
        const bool textured = light.texture_index >= 0 && (uint32_t)light.texture_index < sc.num_light_textures;
        float3 dmrp = v3s(0.0f), smrp = v3s(0.0f);
        const float i_diff = ltc_integrate(P, N, T, B, I, L, light.two_sided != 0, (diff_mrp || textured) ? &dmrp : nullptr);
        if (diff_mrp) *diff_mrp = dmrp;
        diff = emission * i_diff;
        if (setup.wind) rewind(L);                                  // lights.hlsl:203-273: the quadrant's winding, then the roughness swap's
        if (setup.flip_roughness) rewind(L);
        brdf = setup.brdf;
        const float i_spec = ltc_integrate(P, N, T, B, setup.Minv, L, light.two_sided != 0, textured ? &smrp : nullptr, &setup.M);
        spec = emission * i_spec;
        if (textured) {                                             // lights.hlsl:495-511 (a zero integral leaves `mrp` unwritten there: the product stays 0 here)
            const DLightTexture& tex = sc.light_textures[light.texture_index];
            float rx, ry;
            aniso_roughness(s.roughness, s.anisotropy, rx, ry);
            diff = i_diff != 0.0f ? diff * rect_light_texture(tex, light, dmrp, P, 1.0f) : v3s(0.0f);
            spec = i_spec != 0.0f ? spec * rect_light_texture(tex, light, smrp, P, sqrtf(rx * ry)) : v3s(0.0f);
        }
    }
    return bsdf_eval_lut(N, V, s, diff, spec, brdf, surface_model);
}

} // namespace bptd
