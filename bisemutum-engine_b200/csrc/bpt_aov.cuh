// bpt_aov.cuh — per-thread functions of (1) the primary-hit outputs that PathTracingPass::render returns next to the colour
// (OutputData{depth, gbuffer}, path_tracing.cpp:482-487) and (2) the ray-traced ambient-occlusion pass that consumes such
// a depth + normal G-buffer.
//   closest hit + pack   shaders/renderer/raytracing/hits/rt_gbuffer_hit.hlsl:6-18, shaders/renderer/gbuffer.hlsl:18-33
//   texture formats      src/renderer/pass/gbuffer.hpp:14-17 (rgba16_sfloat, rgba16_sfloat, rgba16_unorm, rgba8_unorm)
//   depth                shaders/renderer/raytracing/pt_depth.hlsl:7-16 (d32_sfloat, reverse-Z: 0 = background)
//   RTAO                 shaders/renderer/ambient_occlusion/ambient_occlusion_rt.hlsl:14-66
#pragma once
#include "bpt_trace.cuh"

namespace bptd {

// What the four G-buffer textures hold for one pixel after the trace pass (values as a later load reads them).
BPT_HD void primary_outputs(const DScene& sc, const bpt_camera& cam, float3 O, float3 D, const TraceResult& hit, float& depth, bpt_gbuffer_texel& g) {
    if (!hit.hit) {                                                     // rt_gbuffer.hlsl:33-35 leaves the texels untouched: defined as 0 here
        depth = 0.0f;                                                   // DEVICE_Z_FARTHEST
        for (int k = 0; k < 4; k++) { g.base_color[k] = 0.0f; g.normal_roughness[k] = 0.0f; g.fresnel[k] = 0.0f; g.material_0[k] = 0.0f; }
        return;
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
    float3 pf = frame_to_gbuffer(N, hv.tangent_world);                  // gbuffer.hlsl:27-28
    float2 f0 = fresnel_to_gbuffer(surf.f0_color), f90 = fresnel_to_gbuffer(surf.f90_color);
    g.base_color[0] = q_half(surf.base_color.x); g.base_color[1] = q_half(surf.base_color.y); g.base_color[2] = q_half(surf.base_color.z); g.base_color[3] = 1.0f;
    g.normal_roughness[0] = pf.x; g.normal_roughness[1] = pf.y; g.normal_roughness[2] = pf.z; g.normal_roughness[3] = q_half(surf.roughness);
    g.fresnel[0] = f0.x; g.fresnel[1] = f0.y; g.fresnel[2] = f90.x; g.fresnel[3] = f90.y;
    g.material_0[0] = q_unorm(surf.anisotropy, 255.0f); g.material_0[1] = q_unorm(1.0f / surf.ior, 255.0f); g.material_0[2] = 0.0f;
    g.material_0[3] = q_unorm((float)surface_model / 256.0f, 255.0f);
    const float* m = cam.matrix_proj_view;                              // pt_depth.hlsl:13-15, glm column-major storage: M[c*4 + r]
    float z = ((m[2] * P.x + m[6] * P.y) + m[10] * P.z) + m[14];
    float w = ((m[3] * P.x + m[7] * P.y) + m[11] * P.z) + m[15];
    depth = z / w;
}

// One AO pixel: its four cosine-hemisphere ray directions and the shared origin. Returns false for background pixels
// (ao = (1, 0)). `depth` / `normal_roughness` are full-resolution W x H images; half resolution reads the texel that the
// frame's sub-pixel offset selects (ambient_occlusion_rt.hlsl:18-26; exact texel centres when W and H are even).
BPT_HD bool ao_pixel_rays(const bpt_camera& cam, uint32_t px, uint32_t py, uint32_t aw, uint32_t ah, uint32_t W, uint32_t H, uint32_t frame_index,
                          uint32_t half_res, const float* depth_img, const float4* normal_roughness, float3& origin, float3 dirs[4]) {
    float sx = 0.5f, sy = 0.5f;
    uint32_t tx = px, ty = py;
    if (half_res) {
        sx = (frame_index & 1u) ? 0.75f : 0.25f; sy = (frame_index & 2u) ? 0.75f : 0.25f;
        tx = 2u * px + ((frame_index & 1u) ? 1u : 0u); ty = 2u * py + ((frame_index & 2u) ? 1u : 0u);
        tx = tx < W ? tx : W - 1u; ty = ty < H ? ty : H - 1u;
    }
    float uvx = ((float)px + sx) / (float)aw, uvy = ((float)py + sy) / (float)ah;
    float depth = depth_img[(size_t)ty * W + tx];
    if (depth == 0.0f) return false;                                    // is_depth_background
    float4 nr = normal_roughness[(size_t)ty * W + tx];
    float3 N, T;
    frame_from_gbuffer(v3(nr.x, nr.y, nr.z), N, T);
    Frame3 frame = frame_from_nt(N, T);
    // position_view_from_depth (projection.hlsl:5-10) then inv_view (ambient_occlusion_rt.hlsl:41-42)
    const float* ip = cam.matrix_inv_proj; const float* iv = cam.matrix_inv_view;
    float nx = uvx * 2.0f - 1.0f, ny = 1.0f - uvy * 2.0f;
    float vx = ((ip[0] * nx + ip[4] * ny) + ip[8] * depth) + ip[12];
    float vy = ((ip[1] * nx + ip[5] * ny) + ip[9] * depth) + ip[13];
    float vz = ((ip[2] * nx + ip[6] * ny) + ip[10] * depth) + ip[14];
    float vw = ((ip[3] * nx + ip[7] * ny) + ip[11] * depth) + ip[15];
    vx = vx / vw; vy = vy / vw; vz = vz / vw;
    float3 Pw = v3(((iv[0] * vx + iv[4] * vy) + iv[8] * vz) + iv[12], ((iv[1] * vx + iv[5] * vy) + iv[9] * vz) + iv[13],
                   ((iv[2] * vx + iv[6] * vy) + iv[10] * vz) + iv[14]);
    origin = Pw + N * 0.001f;                                           // :56
    uint32_t seed = rng_tea(py * aw + px, frame_index);                 // :44
    for (int i = 0; i < 4; i++) {
        float r0 = rng_next(seed);
        float r1 = rng_next(seed);
        dirs[i] = to_world(frame, cos_hemisphere_sample(r0, r1));
    }
    return true;
}
BPT_HD float ao_value(uint32_t occluded_rays, float strength) { return 1.0f - ((float)occluded_rays * strength) / 4.0f; }   // :64

// ---- ray-traced reflections: "RTR Sample Direction" (direction_sample/specular_sample.hlsl:14-83) ----
// unpack_gbuffer_to_surface (gbuffer.hlsl:35-45) on the texel values of the four G-buffer textures
BPT_HD void surface_from_gbuffer(const bpt_gbuffer_texel& g, float3& N, float3& T, Surface& s, uint32_t& surface_model) {
    s = surface_default();
    surface_model = ftou(g.material_0[3] * 255.5f);
    s.base_color = v3(g.base_color[0], g.base_color[1], g.base_color[2]);
    s.f0_color = fresnel_from_gbuffer(make_float2(g.fresnel[0], g.fresnel[1]));
    s.f90_color = fresnel_from_gbuffer(make_float2(g.fresnel[2], g.fresnel[3]));
    s.roughness = g.normal_roughness[3];
    frame_from_gbuffer(v3(g.normal_roughness[0], g.normal_roughness[1], g.normal_roughness[2]), N, T);
    s.anisotropy = g.material_0[0];
    s.ior = 1.0f / g.material_0[1];
    s.opacity = 1.0f;
}
// One reflection pixel: false = no ray (background, or rougher than max_roughness); otherwise the ray (origin = the pixel's
// world position, VNDF-sampled direction) and its weight = specular BSDF * fade / pdf (NaN / Inf -> 0). Texel selection as RTAO.
BPT_HD bool rtr_pixel_ray(const bpt_camera& cam, uint32_t px, uint32_t py, uint32_t rw, uint32_t rh, uint32_t W, uint32_t H, uint32_t frame_index,
                          uint32_t half_res, const float* depth_img, const bpt_gbuffer_texel* gbuffer, float max_roughness, float fade_roughness,
                          float3& origin, float3& dir, float3& weight) {
    float sx = 0.5f, sy = 0.5f;
    uint32_t tx = px, ty = py;
    if (half_res) {
        sx = (frame_index & 1u) ? 0.75f : 0.25f; sy = (frame_index & 2u) ? 0.75f : 0.25f;
        tx = 2u * px + ((frame_index & 1u) ? 1u : 0u); ty = 2u * py + ((frame_index & 2u) ? 1u : 0u);
        tx = tx < W ? tx : W - 1u; ty = ty < H ? ty : H - 1u;
    }
    float uvx = ((float)px + sx) / (float)rw, uvy = ((float)py + sy) / (float)rh;
    float depth = depth_img[(size_t)ty * W + tx];
    if (depth == 0.0f) return false;                                    // :28-33
    float3 N, T; Surface surf; uint32_t surface_model;
    surface_from_gbuffer(gbuffer[(size_t)ty * W + tx], N, T, surf, surface_model);
    if (surf.roughness > max_roughness) return false;                   // :46-50
    float3 B = cross3(N, T);
    Frame3 frame = frame_from_nt(N, T);
    const float* ip = cam.matrix_inv_proj; const float* iv = cam.matrix_inv_view;
    float nx = uvx * 2.0f - 1.0f, ny = 1.0f - uvy * 2.0f;
    float vx = ((ip[0] * nx + ip[4] * ny) + ip[8] * depth) + ip[12];
    float vy = ((ip[1] * nx + ip[5] * ny) + ip[9] * depth) + ip[13];
    float vz = ((ip[2] * nx + ip[6] * ny) + ip[10] * depth) + ip[14];
    float vw = ((ip[3] * nx + ip[7] * ny) + ip[11] * depth) + ip[15];
    vx = vx / vw; vy = vy / vw; vz = vz / vw;
    float3 Pw = v3(((iv[0] * vx + iv[4] * vy) + iv[8] * vz) + iv[12], ((iv[1] * vx + iv[5] * vy) + iv[9] * vz) + iv[13],
                   ((iv[2] * vx + iv[6] * vy) + iv[10] * vz) + iv[14]);
    float3 V = normalize3(v3(iv[12], iv[13], iv[14]) - Pw);             // :57, camera_position_world (camera.hlsl:7-9)
    float3 V_local = to_local(frame, V);
    float rx, ry;
    aniso_roughness(surf.roughness, surf.anisotropy, rx, ry);
    uint32_t seed = rng_tea(py * rw + px, frame_index);                 // :63
    float u1 = rng_next(seed);
    float u2 = rng_next(seed);
    float3 half_dir = ggx_vndf_sample(V_local, rx, ry, u1, u2);
    float3 out_local = reflect3(-V_local, half_dir);
    float pdf_wh = ggx_vndf_pdf(half_dir, V_local, rx, ry);
    float pdf = pdf_wh / (4.0f * fabsf(dot3(half_dir, V_local)));
    float3 out_dir = to_world(frame, out_local);
    float3 spec = bsdf_eval_specular(N, T, B, V, out_dir, surf, surface_model);
    float fade = 1.0f - tmax_(surf.roughness - fade_roughness, 0.0f) / tmax_(max_roughness - fade_roughness, 0.0001f);
    weight = (spec * fade) / pdf;
    if (!is_finite3(weight)) weight = v3s(0.0f);                        // :76-78
    origin = Pw; dir = out_dir;
    return true;
}

// ---- simple_upscale_cs (shaders/renderer/simple_upscale.hlsl:47-97): one full-resolution pixel ----
BPT_HD float linear_01_depth(float depth, const bpt_camera& cam) {      // core/utils/depth.hlsl:31-35; inv_proj[3] is ROW 3: (m[3], m[7], m[11], m[15])
    const float a = cam.matrix_inv_proj[11], b = cam.matrix_inv_proj[15];
    return ((1.0f - depth) * b) / (a * depth + b);
}
BPT_HD float4 upscale_pixel(const bpt_camera& cam, uint32_t x, uint32_t y, uint32_t W, uint32_t H, uint32_t frame_index, const float* depth_img,
                            const float4* normal_roughness, const float4* in_half) {
    const int rw = (int)((W + 1) / 2), rh = (int)((H + 1) / 2);
    auto depth_at = [&](int px, int py) { return (px < (int)W && py < (int)H) ? depth_img[(size_t)py * W + px] : 0.0f; };        // Texture.Load: 0 out of range
    auto normal_at = [&](int px, int py) {
        float4 nr = (px < (int)W && py < (int)H) ? normal_roughness[(size_t)py * W + px] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        return oct_decode(make_float2(nr.x, nr.y));
    };
    const float center_depth = linear_01_depth(depth_at((int)x, (int)y), cam);
    const float3 center_normal = normal_at((int)x, (int)y);
    const float cx = (x & 1u) ? 0.75f : 0.25f, cy = (y & 1u) ? 0.75f : 0.25f;                           // :66-73
    const float ix = (frame_index & 1u) ? 0.75f : 0.25f, iy = (frame_index & 2u) ? 0.75f : 0.25f;
    const int sx = (int)(frame_index & 1u), sy = (int)((frame_index >> 1) & 1u);
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float sum_weight = 0.0f;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            float ox = ((float)dx + ix) - cx, oy = ((float)dy + iy) - cy;
            float r = sqrtf(sqrtf(ox * ox + oy * oy));
            float temp = r / 0.2f;
            float w = exp_neg(-(temp * temp));                                                              // gaussian(r, 0.2)
            int hx = (int)(x / 2) + dx, hy = (int)(y / 2) + dy;                                            // fetch_shared_data, :25-35
            hx = hx < 0 ? 0 : (hx > rw ? rw : hx); hy = hy < 0 ? 0 : (hy > rh ? rh : hy);               // clamp(.., 0, (tex_size + 1) / 2)
            float4 v = (hx < rw && hy < rh) ? in_half[(size_t)hy * rw + hx] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float tap_depth = linear_01_depth(depth_at(hx * 2 + sx, hy * 2 + sy), cam);
            float3 tap_normal = normal_at(hx * 2 + sx, hy * 2 + sy);
            w = w * tmax_(dot3(tap_normal, center_normal), 0.0f);
            w = w * tmax_(0.0f, 1.0f - fabsf(tap_depth - center_depth));
            sum = make_float4(sum.x + v.x * w, sum.y + v.y * w, sum.z + v.z * w, sum.w + v.w * w);
            sum_weight = sum_weight + w;
        }
    if (sum_weight == 0.0f) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    return make_float4(sum.x / sum_weight, sum.y / sum_weight, sum.z / sum_weight, sum.w / sum_weight);
}

} // namespace bptd
