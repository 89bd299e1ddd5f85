// bpt_shade.cuh — one path vertex: closest-hit material + direct lighting + next direction.
//
// This fuses what the reference runs as three full-screen passes per bounce
//   closest hit      shaders/renderer/raytracing/hits/rt_gbuffer_hit.hlsl:6-18  (G-buffer write)
//   lighting         shaders/renderer/raytracing/deferred_lighting_secondary.hlsl:11-111
//   next direction   shaders/renderer/raytracing/direction_sample/sample_secondary_ray.hlsl:11-69
// into one per-path function: the surface never leaves registers, so the 28 B/px G-buffer round
// trip (and its fp16 / rg11b10 / unorm8 quantisation) disappears (state_precision = fp32).
// Visibility of directional / point / spot lights is a shadow ray handed to `sink.shadow()` (NEW:
// the reference samples rasterised shadow maps, lights.hlsl:27-159).
#pragma once
#include "bpt_ltc.cuh"
#include "bpt_ddgi.cuh"
#include "bpt_ibl.cuh"
#include "bpt_trace.cuh"

namespace bptd {

struct ShadeParams {
    uint32_t width, height;
    uint32_t max_bounces;      // clamped to [2,16] (path_tracing.cpp:290)
    uint32_t nee_mode;
    float ray_length;
    uint32_t russian_roulette; // NEW switch, default off
    uint32_t rect_shadow;      // NEW switch: 1 = one shadow ray per rect light towards its most representative point
    uint32_t diffuse_only;     // probe tracing: light every vertex as surface_data_diffuse(base_color) (ddgi/deferred_lighting.hlsl:44)
    uint32_t state_precision;  // BPT_STATE_REFERENCE_FP16: the reference's texture formats are applied to the state (see below)
    uint32_t ibl;              // ray-traced reflections: add the IBL block of deferred_lighting_secondary.hlsl:98-108 (the path tracer compiles it out)
};

// state_precision = reference_fp16 (SURVEY §8a "storage quantisation"; formats at path_tracing.cpp:248-288 and
// pass/gbuffer.hpp:14-17). Every value the reference moves between passes through a texture takes that texture's
// format here too: ray directions and throughput are halves (also the camera ray's), the surface goes through
// pack_surface_to_gbuffer / unpack_gbuffer_to_surface, and a bounce's colour is the half of
// (sum of its light terms) * throughput (deferred_lighting_secondary.hlsl:110), which bounce 1 writes and later
// bounces add to the colour texture with a half result (the additive blit, path_tracing.cpp:441-459). The sink
// then receives UNWEIGHTED light terms and the caller commits them with commit_bounce_fp16.
BPT_HD float3 commit_bounce_fp16(float3 C, float3 bounce_sum, float3 Wt, uint32_t bounce) {
    float3 c = q_half3(bounce_sum * Wt);
    return bounce == 1 ? c : q_half3(C + c);
}
// pt_accumulate.hlsl:3-11 on an rgba16_sfloat target: image = n == 1 ? C : half(lerp(image, C, 1/n))
BPT_HD float3 accumulate_fp16(float3 image, float3 C, uint32_t n) {
    if (n <= 1) return C;
    return q_half3(mix3(image, C, 1.0f / (float)n));
}

// Sink concept:
//   void add(float3 c)                                         — unshadowed radiance for this pixel
//   void shadow(float3 P, float3 L, float tmax, float3 c, uint32_t light) — NEE candidate
// Returns true when the path continues; (nO, nD, nW) is then the next extend ray.
// IBL (compile time): the reflection pass's kernel evaluates the IBL block; the path tracer's kernel is compiled without it, like
// the reference compiles its shader with DEFERRED_LIGHTING_NO_IBL (path_tracing.cpp:152).
// RECT / GENERAL (compile time): the path tracer's common case — no rect lights; FP32 state, no probe mode, no Russian roulette — gets a
// kernel without the LTC evaluation and without the branches of those modes (fewer live registers: 722 -> ~130 bytes of spills at 64).
template <class Sink, bool IBL = false, bool RECT = true, bool GENERAL = true>
BPT_HD bool shade_vertex(const DScene& sc, const ShadeParams& sp, uint32_t frame_index, uint32_t bounce, uint32_t pixel,
                         float3 O, float3 D, float3 Wt, const TraceResult& hit, Sink& sink, float3& nO, float3& nD, float3& nW) {
    const bool fp16 = GENERAL && sp.state_precision == BPT_STATE_REFERENCE_FP16;
    const float3 Wl = fp16 ? v3s(1.0f) : Wt;                            // weight of the light terms handed to the sink
    if (!hit.hit) {                                                     // deferred_lighting_secondary.hlsl:24-29
        const float* m = sc.sky_transform;
        float3 dir = v3((m[0] * D.x + m[1] * D.y) + m[2] * D.z, (m[3] * D.x + m[4] * D.y) + m[5] * D.z, (m[6] * D.x + m[7] * D.y) + m[8] * D.z);
        float3 color = sample_sky(sc, dir) * v3(sc.sky_color[0], sc.sky_color[1], sc.sky_color[2]);
        sink.add(fp16 ? color : color * Wt);
        return false;
    }
    const DInstance& in = sc.instances[hit.slot];
    const bpt_drawable_sbt_data& dr = sc.drawables[in.instance_id];
    const bpt_material& mat = sc.materials[dr.material_offset / (uint32_t)sizeof(bpt_material)];
    uint32_t surface_model = (mat.flags >> BPT_MATERIAL_MODEL_SHIFT) & 0xffu;
    float3 P = O + D * hit.t;                                           // rt_gbuffer.hlsl:32
    HitVertex hv = fetch_hit_vertex(sc, in, hit.prim, hit.u, hit.v, material_needs_position(mat), material_needs_color(mat));
    Surface surf = eval_material(sc, mat, hv.texcoord, hv.position_world, hv.color);
    float3 nts = surf.normal_map_value * 2.0f - v3s(1.0f);              // rt_gbuffer_hit.hlsl:10-14
    float3 N = normalize3((nts.x * hv.tangent_world + nts.y * hv.bitangent_world) + nts.z * hv.normal_world);
    if (surf.two_sided && dot3(D, N) > 0.0f) N = -N;
    float3 T;
    if (fp16) {                                                         // gbuffer.hlsl:18-45 through the texture formats
        float3 Nq;
        frame_through_gbuffer(N, hv.tangent_world, Nq, T);
        N = Nq;
        surface_through_gbuffer(surf, surface_model);
    } else {
        T = tangent_after_gbuffer(N, hv.tangent_world);                 // gbuffer.hlsl:27,41
    }
    float3 B = cross3(N, T);
    surf.opacity = 1.0f;                                                // gbuffer.hlsl:44
    if (GENERAL && sp.diffuse_only) {                                   // ddgi/deferred_lighting.hlsl:44-45
        Surface d = surface_default();
        d.base_color = surf.base_color;
        surf = d;
        surface_model = 1u;                                             // MATERIAL_SURFACE_MODEL_LIT (:45)
    }
    float3 V = normalize3(O - P);                                       // deferred_lighting_secondary.hlsl:45

    LtcSetup ltc{};                                                     // the LUT side of the LTC evaluation: once per vertex, not per light
    float3 ltc_lv = v3s(0.0f);
    if (RECT && sc.num_rect) {
        float rx, ry;
        aniso_roughness(surf.roughness, surf.anisotropy, rx, ry);
        ltc_lv = v3(dot3(V, T), dot3(V, B), dot3(V, N));
        if (ltc_lv.z > 0.0f) ltc = ltc_setup(sc, ltc_lv, rx, ry);
    }
    for (uint32_t l = 0; RECT && l < sc.num_rect; l++) {                // :72-96 (unshadowed, as the reference, unless rect_shadow)
        const bpt_rect_light_data& rl = sc.rect_lights[l];
        float3 mrp = v3s(0.0f);
        float3 c = eval_rect_light(sc, rl, P, N, T, B, V, surf, surface_model, ltc_lv, ltc, sp.rect_shadow ? &mrp : nullptr) * Wl;
        if (!sp.rect_shadow) { sink.add(c); continue; }
        if (!(max3c(c) > 0.0f)) continue;
        // distance to the light's plane along mrp (as rect_light_sample_texture, lights.hlsl:425-438)
        float3 ln = v3(rl.normal[0], rl.normal[1], rl.normal[2]);
        float step = fabsf(dot3(mrp, ln));
        if (!(step >= 0.0001f)) { sink.add(c); continue; }
        float dist = fabsf(dot3(P - v3(rl.position2[0], rl.position2[1], rl.position2[2]), ln));
        sink.shadow(P, mrp, (dist / step) * 0.999f, c, sc.num_dir + sc.num_point + l);
    }
    if (IBL && sp.ibl && sc.ibl_enabled) sink.add(ibl_lighting(sc, N, V, surf, surface_model) * Wl);     // deferred_lighting_secondary.hlsl:98-108
    // Probe paths only: the previous DDGI update lights the path's last vertex (ddgi/deferred_lighting.hlsl:102-115:
    // color += ddgi.xyz / ddgi.a * base_color / pi). The reference traces one bounce, so every probe-ray hit gets it; with
    // more bounces (BASELINE configs[4]) it closes the path instead of being added at every vertex.
    if (GENERAL && sp.diffuse_only && sc.ddgi_enabled && bounce + 1 >= sp.max_bounces) {
        float4 g = ddgi_volume_lighting(sc.ddgi_volume, sc.ddgi_irr_size, sc.ddgi_vis_size, sc.ddgi_irradiance, sc.ddgi_visibility, P, N, V);
        if (g.w > 0.0f) sink.add(((v3(g.x / g.w, g.y / g.w, g.z / g.w) * surf.base_color) * kInvPi) * Wl);
    }
    for (uint32_t l = 0; l < sc.num_dir; l++) {                         // :51-60
        const bpt_dir_light_data& li = sc.dir_lights[l];
        float3 L = v3(li.direction[0], li.direction[1], li.direction[2]);
        float3 c = (v3(li.emission[0], li.emission[1], li.emission[2]) * bsdf_eval(N, T, B, V, L, surf, surface_model)) * Wl;
        if (max3c(c) > 0.0f) sink.shadow(P, L, sp.ray_length, c, l);
    }
    for (uint32_t l = 0; l < sc.num_point; l++) {                       // :61-70
        float3 L; float dist;
        float3 le = eval_point_light(sc.point_lights[l], P, L, dist);
        if (le.x == 0.0f && le.y == 0.0f && le.z == 0.0f) continue;     // out of range / outside the cone: 0 * bsdf can only be 0 or NaN, neither emits a ray
        float3 c = (le * bsdf_eval(N, T, B, V, L, surf, surface_model)) * Wl;
        if (max3c(c) > 0.0f) sink.shadow(P, L, dist * 0.999f, c, sc.num_dir + l);
    }

    // next direction — sample_secondary_ray.hlsl:11-69 with bounce_index = bounce
    if (bounce + 1 >= sp.max_bounces) return false;
    if (max3c(Wt) < 0.001f) return false;                               // :23-28
    Frame3 frame = frame_from_nt(N, T);                                 // :42
    float3 V_local = to_local(frame, V);
    float rx, ry;
    aniso_roughness(surf.roughness, surf.anisotropy, rx, ry);
    uint32_t seed = rng_tea(pixel, frame_index + bounce * 3u);          // :52 (pixel = y * width + x)
    float u1 = rng_next(seed);
    float u2 = rng_next(seed);
    float3 half_dir = ggx_vndf_sample(V_local, rx, ry, u1, u2);
    float3 out_local = reflect3(-V_local, half_dir);
    float pdf_wh = ggx_vndf_pdf(half_dir, V_local, rx, ry);
    float pdf = pdf_wh / (4.0f * fabsf(dot3(half_dir, V_local)));
    float3 out_dir = to_world(frame, out_local);
    float3 bsdf = bsdf_eval(N, T, B, V, out_dir, surf, surface_model);
    float3 weight = bsdf / pdf;
    if (!is_finite3(weight)) weight = v3s(0.0f);                        // :62-64
    float3 w2 = weight * Wt;
    if (fp16) { w2 = q_half3(w2); out_dir = q_half3(out_dir); }         // ray_weights / ray_directions are rgba16_sfloat (:66-68)
    // a zero-weight path can never contribute again (deferred_lighting_secondary.hlsl:17-21): drop it
    if (w2.x == 0.0f && w2.y == 0.0f && w2.z == 0.0f) return false;
    if (GENERAL && sp.russian_roulette && bounce >= 2) {                           // third draw of this bounce's stream
        float q = clampf_(max3c(w2), 0.05f, 1.0f);
        float u3 = rng_next(seed);
        if (!(u3 < q)) return false;
        w2 = w2 / q;
        if (fp16) w2 = q_half3(w2);
    }
    nO = P; nD = out_dir; nW = w2;
    return true;
}

} // namespace bptd
