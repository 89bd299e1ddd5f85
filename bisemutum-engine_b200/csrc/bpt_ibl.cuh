// bpt_ibl.cuh — image-based lighting of the skybox: the three precompute shaders of SkyboxPrecomputePass
// (bisemutum/src/renderer/pass/skybox_precompute.cpp:66-162) and the IBL block of the secondary lighting shader
// (shaders/renderer/raytracing/deferred_lighting_secondary.hlsl:98-108), which ray-traced reflections evaluate at every hit.
//   ibl_brdf_lut_cs               shaders/renderer/skybox/ibl_brdf_lut.hlsl:8-32              -> rg8_unorm, 128 x 128
//   skybox_precompute_diffuse_cs  shaders/renderer/skybox/skybox_precompute_diffuse.hlsl:12-41 -> rgba16_sfloat cube, 256
//   skybox_precompute_specular_cs shaders/renderer/skybox/skybox_precompute_specular.hlsl:12-41 -> rgba16_sfloat cube, 256, 5 mips
//   (sizes and formats: src/renderer/context/skybox.cpp:11-29; they are parameters here)
// One thread per texel runs the shader's own sample loop in its own order, so the result does not depend on the launch shape;
// sin / cos of an angle a are sincos_2pi(a / 2pi) (the numeric contract's fixed-order form). The cubemap is read through
// sample_cube (bilinear inside a face, as sample_sky). Stores round to the target format (q_half / q_unorm).
#pragma once
#include "bpt_scene.cuh"

namespace bptd {

constexpr float kTwoPi = 2.0f * 3.14159265359f, kInvTwoPi = 0.15915494309189533577f;      // (kPi: bpt_math.cuh = core/utils/math.hlsl:3)
constexpr uint32_t kIblSamples = 1024;                   // ibl_brdf_lut.hlsl:6, skybox_precompute_specular.hlsl:9
constexpr float kIblClampLum = 12.0f;                    // skybox_precompute_*.hlsl: clamp_lum

BPT_HD float radical_inverse_vdc(uint32_t bits) {        // core/utils/low_discrepancy.hlsl:3-10
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}
BPT_HD float3 cubemap_direction(float u, float v, uint32_t layer) {      // core/utils/cubemap.hlsl:3-21
    u = u * 2.0f - 1.0f; v = v * 2.0f - 1.0f;
    float3 d;
    if (layer == 0) d = v3(1.0f, -v, -u);
    else if (layer == 1) d = v3(-1.0f, -v, u);
    else if (layer == 2) d = v3(u, 1.0f, v);
    else if (layer == 3) d = v3(u, -1.0f, -v);
    else if (layer == 4) d = v3(u, -v, 1.0f);
    else d = v3(-u, -v, -1.0f);
    return normalize3(d);
}
BPT_HD float luminance3(float3 c) { return (c.x * 0.212671f + c.y * 0.715160f) + c.z * 0.072169f; }   // core/utils/color.hlsl:3-5

// ibl_brdf_lut_cs: texel (x, y) of a res x res LUT; the value an rg8_unorm target holds
BPT_HD float2 ibl_brdf_lut_texel(uint32_t x, uint32_t y, uint32_t res) {
    const float inv = 1.0f / (float)res;
    float ndotv = ((float)x + 0.5f) * inv, roughness = ((float)y + 0.5f) * inv;
    float3 wi = v3(sqrtf(1.0f - ndotv * ndotv), 0.0f, ndotv);
    float vx = 0.0f, vy = 0.0f;
    for (uint32_t i = 0; i < kIblSamples; i++) {
        float3 wh = ggx_vndf_sample(wi, roughness, roughness, (float)i / (float)kIblSamples, radical_inverse_vdc(i));
        float3 wo = reflect3(-wi, wh);
        if (wo.z <= 0.0f) continue;
        float weight = ggx_g1(wo, roughness, roughness);                 // ggx_vndf_sample_weight_sep (utils.hlsl:123-125)
        float f = pow5f(1.0f - dot3(wi, wh));
        vx = vx + (1.0f - f) * weight; vy = vy + f * weight;
    }
    return make_float2(q_unorm(vx / (float)kIblSamples, 255.0f), q_unorm(vy / (float)kIblSamples, 255.0f));
}
// skybox_precompute_diffuse_cs: texel (x, y) of face `layer` of a size^2 cube; 128 x 32 (phi, theta) cells of the hemisphere
BPT_HD float3 ibl_diffuse_texel(const float4* sky, uint32_t sky_size, uint32_t x, uint32_t y, uint32_t layer, uint32_t size) {
    const float inv = 1.0f / (float)size;
    float3 dir = cubemap_direction(((float)x + 0.5f) * inv, ((float)y + 0.5f) * inv, layer);
    Frame3 frame = frame_from_normal(dir);
    float3 irradiance = v3s(0.0f);
    const float delta = kPi / 64.0f;
    for (float phi = delta * 0.5f; phi < kTwoPi; phi += delta) {
        float sin_phi, cos_phi;
        sincos_2pi(phi * kInvTwoPi, sin_phi, cos_phi);
        for (float theta = delta * 0.5f; theta < 0.5f * kPi; theta += delta) {
            float sin_theta, cos_theta;
            sincos_2pi(theta * kInvTwoPi, sin_theta, cos_theta);
            float3 v = to_world(frame, v3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
            float3 color = sample_cube(sky, sky_size, v);
            float scale = kIblClampLum / tmax_(luminance3(color), kIblClampLum);
            irradiance = irradiance + ((color * scale) * cos_theta) * sin_theta;
        }
    }
    return q_half3((irradiance * kPi) / 4096.0f);
}
// skybox_precompute_specular_cs: texel of mip `level` (size = base >> level, roughness = level / (levels - 1))
BPT_HD float3 ibl_specular_texel(const float4* sky, uint32_t sky_size, uint32_t x, uint32_t y, uint32_t layer, uint32_t size, float roughness) {
    const float inv = 1.0f / (float)size;
    float3 dir = cubemap_direction(((float)x + 0.5f) * inv, ((float)y + 0.5f) * inv, layer);
    Frame3 frame = frame_from_normal(dir);
    const float3 wi = v3(0.0f, 0.0f, 1.0f);
    float3 filtered = v3s(0.0f);
    float weight_sum = 0.0f;
    for (uint32_t i = 0; i < kIblSamples; i++) {
        float3 wh = ggx_vndf_sample(wi, roughness, roughness, (float)i / (float)kIblSamples, radical_inverse_vdc(i));
        float3 wo = reflect3(-wi, wh);
        if (wo.z <= 0.0f) continue;
        float3 color = sample_cube(sky, sky_size, to_world(frame, wo));
        float scale = kIblClampLum / tmax_(luminance3(color), kIblClampLum);
        filtered = filtered + (color * scale) * wo.z;
        weight_sum = weight_sum + wo.z;
    }
    return q_half3(filtered / weight_sum);
}

// The IBL block of deferred_lighting_secondary.hlsl:98-108 at a hit. skybox_sampler is linear / clamp with the default
// NEAREST mip mode (skybox.cpp:31-37, rhi/sampler.hpp:40): the specular level is ceil(lod + 0.5) - 1.
BPT_HD float3 ibl_lighting(const DScene& sc, float3 N, float3 V, const Surface& surf, uint32_t surface_model) {
    const float* m = sc.sky_transform;
    auto xf = [&](float3 d) { return v3((m[0] * d.x + m[1] * d.y) + m[2] * d.z, (m[3] * d.x + m[4] * d.y) + m[5] * d.z, (m[6] * d.x + m[7] * d.y) + m[8] * d.z); };
    float3 ibl_diffuse = sample_cube(sc.ibl_diffuse, sc.ibl_diffuse_size, xf(N)) * v3(sc.ibl_diffuse_color[0], sc.ibl_diffuse_color[1], sc.ibl_diffuse_color[2]);
    float lod = surf.roughness * (float)(sc.ibl_specular_levels - 1u);
    int level = (int)ceilf(lod + 0.5f) - 1;
    level = level < 0 ? 0 : (level >= (int)sc.ibl_specular_levels ? (int)sc.ibl_specular_levels - 1 : level);
    size_t offset = 0;
    for (int l = 0; l < level; l++) { size_t s = sc.ibl_specular_size >> l; offset += 6 * s * s; }
    float3 ibl_specular = sample_cube(sc.ibl_specular + offset, sc.ibl_specular_size >> level, xf(reflect3(-V, N))) *
                          v3(sc.ibl_specular_color[0], sc.ibl_specular_color[1], sc.ibl_specular_color[2]);
    // brdf_lut.SampleLevel(sampler, (dot(N, V), roughness), 0).xy: bilinear, clamp to edge
    const int n = (int)sc.ibl_brdf_size;
    float x = dot3(N, V) * (float)n - 0.5f, y = surf.roughness * (float)n - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap_tc((int)x0f, n, BPT_ADDRESS_CLAMP), x1 = wrap_tc((int)x0f + 1, n, BPT_ADDRESS_CLAMP);
    int y0 = wrap_tc((int)y0f, n, BPT_ADDRESS_CLAMP), y1 = wrap_tc((int)y0f + 1, n, BPT_ADDRESS_CLAMP);
    float2 a = BPT_LDG(sc.ibl_brdf + (size_t)y0 * n + x0), b = BPT_LDG(sc.ibl_brdf + (size_t)y0 * n + x1);
    float2 c = BPT_LDG(sc.ibl_brdf + (size_t)y1 * n + x0), d = BPT_LDG(sc.ibl_brdf + (size_t)y1 * n + x1);
    float2 top = make_float2(mix1(a.x, b.x, fx), mix1(a.y, b.y, fx)), bot = make_float2(mix1(c.x, d.x, fx), mix1(c.y, d.y, fx));
    float2 brdf = make_float2(mix1(top.x, bot.x, fy), mix1(top.y, bot.y, fy));
    return bsdf_eval_lut(N, V, surf, ibl_diffuse, ibl_specular, brdf, surface_model);
}

} // namespace bptd
