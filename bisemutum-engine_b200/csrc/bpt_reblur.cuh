// bpt_reblur.cuh — ReBLUR, the denoiser of the ray-traced reflections, one pixel of each pass.
//
// Replaces ReblurPass::render (bisemutum/src/renderer/pass/reblur.cpp:273-588: eight compute passes) and the shaders under
// shaders/renderer/reblur/: pre_blur, temporal_accumulate, fetch_linear_depth, gen_depth_mip, fix_history, blur, temporal_stabilize,
// post_blur (+ utils.hlsl, filter.hlsl). Called from ReflectionPass::render (reflection.cpp:529) with the reflection colour
// (`noised_tex`) and the hit positions of bpt_trace_reflection.
//
// Storage follows the reference's textures: lighting_dist_0 / _1 and the result are rgba16_sfloat, the accumulation speed r16_sfloat
// (every store rounds to half, q_half), linear depth r32_sfloat; lighting_dist_0 and linear depth carry 4 mip levels. A Texture.Load
// outside the texture returns 0; SampleLevel uses the pass's linear / clamp-to-edge sampler as explicit FP32 bilinear arithmetic.
// Transcendentals (atan, log, pow, exp, exp2 in utils.hlsl / filter.hlsl) are the fixed-order forms of bpt_math.cuh, so the host
// compiler and nvcc agree bit for bit.
#pragma once
#include "bpt_math.cuh"
#include "../../include/bpt/bpt.h"

namespace bptd {

constexpr float kReblurMaxAccum = 32.0f;        // REBLUR_NUM_MAX_ACCUM_FRAME
constexpr float kReblurMipLevels = 4.0f;        // REBLUR_NUM_MIP_LEVEL
constexpr float kReblurMaxFixFrames = 4.0f;     // REBLUR_NUM_MAX_FRAME_WITH_HISTORY_FIX

struct ReblurView {
    uint32_t w, h;             // denoiser extent (= noised_tex; the shaders' tex_size AND gbuffer_tex_size, reblur.cpp:334-335)
    uint32_t gw, gh;           // extent of depth / normal_roughness / velocity (the camera target)
    uint32_t half_res, frame_index, has_history, virtual_history;
    float blur_radius, anti_flicker;
    bpt_camera cam, hist_cam;
    float4 rot_pre, rot_blur, rot_post;
    const float* depth; const float4* normal_roughness; const float2* velocity; const uint8_t* validation;
    const float4* hit_positions; const float4* noised;
    const float* hist_depth; const float4* hist_normal_roughness;            // previous frame, gw x gh
    const float4* hist_ld0; const float4* hist_ld1; const float* hist_accum;  // previous frame, w x h
    float4* ld0; float4* ld1; float* accum; float* lin_depth; float4* denoised;   // ld0 / lin_depth: 4 levels, level after level
};

BPT_HD size_t reblur_mip_offset(uint32_t w, uint32_t h, int level) {
    size_t o = 0;
    for (int l = 0; l < level; l++) o += (size_t)(w >> l) * (h >> l);
    return o;
}
BPT_HD float length3(float3 v) { return sqrtf(dot3(v, v)); }      // HLSL length
BPT_HD float4 rb_q4(float4 v) { return make_float4(q_half(v.x), q_half(v.y), q_half(v.z), q_half(v.w)); }
BPT_HD float4 rb_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
BPT_HD float4 rb_mul(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
BPT_HD float4 rb_div(float4 a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
BPT_HD float4 rb_mix(float4 a, float4 b, float t) { return make_float4(mix1(a.x, b.x, t), mix1(a.y, b.y, t), mix1(a.z, b.z, t), mix1(a.w, b.w, t)); }
// Texture2D.Load: zero outside
template <class T> BPT_HD T rb_zero();
template <> BPT_HD float rb_zero<float>() { return 0.0f; }
template <> BPT_HD float4 rb_zero<float4>() { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
template <> BPT_HD float2 rb_zero<float2>() { return make_float2(0.0f, 0.0f); }
template <class T> BPT_HD T rb_load(const T* tex, uint32_t w, uint32_t h, int x, int y) {
    return (tex && x >= 0 && y >= 0 && x < (int)w && y < (int)h) ? tex[(size_t)y * w + x] : rb_zero<T>();
}
// SampleLevel(input_sampler, uv, 0): linear, clamp to edge (reblur.cpp:242-248); a null history texture samples as 0
BPT_HD float4 rb_sample4(const float4* tex, uint32_t w, uint32_t h, float u, float v) {
    if (!tex) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y), fx = x - x0f, fy = y - y0f;
    auto cl = [](float c, int n) { int i = (int)c; return i < 0 ? 0 : (i >= n ? n - 1 : i); };
    // (a huge or NaN coordinate is clamped in float before the cast, whose result would otherwise depend on the compiler; the weights
    // fx, fy keep the NaN, so such a sample is NaN on every implementation)
    x0f = !(x0f >= -2.0f) ? -2.0f : (x0f > (float)w ? (float)w : x0f); y0f = !(y0f >= -2.0f) ? -2.0f : (y0f > (float)h ? (float)h : y0f);
    int x0 = cl(x0f, (int)w), x1 = cl(x0f + 1.0f, (int)w), y0 = cl(y0f, (int)h), y1 = cl(y0f + 1.0f, (int)h);
    float4 a = tex[(size_t)y0 * w + x0], b = tex[(size_t)y0 * w + x1], c = tex[(size_t)y1 * w + x0], e = tex[(size_t)y1 * w + x1];
    return rb_mix(rb_mix(a, b, fx), rb_mix(c, e, fx), fy);
}
BPT_HD float rb_sample1(const float* tex, uint32_t w, uint32_t h, float u, float v) {
    if (!tex) return 0.0f;
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y), fx = x - x0f, fy = y - y0f;
    auto cl = [](float c, int n) { int i = (int)c; return i < 0 ? 0 : (i >= n ? n - 1 : i); };
    x0f = !(x0f >= -2.0f) ? -2.0f : (x0f > (float)w ? (float)w : x0f); y0f = !(y0f >= -2.0f) ? -2.0f : (y0f > (float)h ? (float)h : y0f);
    int x0 = cl(x0f, (int)w), x1 = cl(x0f + 1.0f, (int)w), y0 = cl(y0f, (int)h), y1 = cl(y0f + 1.0f, (int)h);
    return mix1(mix1(tex[(size_t)y0 * w + x0], tex[(size_t)y0 * w + x1], fx), mix1(tex[(size_t)y1 * w + x0], tex[(size_t)y1 * w + x1], fx), fy);
}
// HLSL int(x): truncation towards zero, saturating
BPT_HD int rb_trunc(float f) { return f != f ? 0 : (f >= 2147483520.0f ? 2147483647 : (f <= -2147483648.0f ? (int)0x80000000 : (int)f)); }

// ---- fixed-order transcendentals for utils.hlsl ----
BPT_HD float rb_pow(float x, float y) {                  // pow(x, y) for 0 <= x <= 1, y > 0 (the only uses): x^y = exp(y ln x)
    if (!(x > 0.0f)) return 0.0f;
    if (x >= 1.0f) return 1.0f;
    return exp_neg((y * log2_(x)) * 0.693147181f);
}
BPT_HD float rb_atan_pos(float x) { return atan2_(x, 1.0f); }

// ---- utils.hlsl ----
BPT_HD float2 rb_rotate(float4 r, float2 v) { return make_float2(v.x * r.x + v.y * r.z, v.x * r.y + v.y * r.w); }      // :12-14
BPT_HD float rb_lobe_half_angle(float roughness, float percentage) {                                                        // :16-19
    float m = roughness * roughness;
    return rb_atan_pos((m * percentage) / (1.0f - percentage));
}
BPT_HD float rb_magic_curve2(float roughness, float percentage) {                                                           // :21-25
    return sat(rb_lobe_half_angle(roughness, percentage) / rb_lobe_half_angle(1.0f, percentage));
}
BPT_HD float rb_dominant_factor(float ndotv, float roughness) {                                                             // :27-31
    float a = 0.298475f * (log2_(39.4115f - 39.0029f * roughness) * 0.693147181f);
    float f = rb_pow(sat(1.0f - ndotv), 10.8649f) * (1.0f - a) + a;
    return sat(f);
}
BPT_HD float4 rb_dominant_direction(float3 N, float3 V, float roughness) {                                                  // :33-44
    float ndotv = fabsf(dot3(N, V));
    float f = rb_dominant_factor(ndotv, roughness);
    float3 R = reflect3(-V, N);
    float3 D = normalize3(mix3(N, R, f));
    return make_float4(D.x, D.y, D.z, f);
}
BPT_HD void rb_kernel_basis(float3 V, float3 N, float roughness, float3& T, float3& B) {                                     // :46-66
    Frame3 basis = frame_from_normal(N);
    T = basis.x; B = basis.y;
    float4 dd = rb_dominant_direction(N, V, roughness);
    float3 D = v3(dd.x, dd.y, dd.z);
    float ndotd = fabsf(dot3(N, D));
    if (ndotd < 0.999f && roughness != 1.0f) {
        float3 dr = reflect3(-D, N);
        T = normalize3(cross3(N, dr));
        B = cross3(dr, T);
        float ndotv = fabsf(dot3(N, V));
        float acos01sq = sat(1.0f - ndotv);
        float skew = mix1(1.0f, roughness, sqrtf(acos01sq));
        T = T * skew;
    }
}
BPT_HD float rb_parallax(float3 cur, float3 prev) {                                                                          // :68-71
    float cosa = sat(dot3(cur, prev));
    return (sqrtf(1.0f - cosa * cosa) / tmax_(cosa, 0.00001f)) * 60.0f;
}
BPT_HD float rb_accum_speed(float roughness, float ndotv, float parallax) {                                                  // :73-86 (SPEC_ACCUM_CURVE 0.5, BASE_POWER 1)
    float acos01sq = 1.0f - ndotv;
    float a = sqrtf(sat(acos01sq));                         // pow(x, 0.5)
    float b = 1.1f + roughness * roughness;
    float sens = (b + a) / (b - a);
    float power_scale = 1.0f + parallax * sens;
    float f = 1.0f - exp_neg(((-200.0f * roughness) * roughness) * 0.693147181f);       // exp2
    f = f * rb_pow(sat(roughness), 1.0f * power_scale);
    return kReblurMaxAccum * f;
}
BPT_HD float rb_hit_dist_atten(float roughness, float camera_dist, float hit_dist) {                                         // :88-91
    float f = hit_dist < 0.0f ? 1.0f : hit_dist / (hit_dist + camera_dist);
    return mix1(0.5f * roughness, 1.0f, f);
}

// ---- depth.hlsl / projection.hlsl with this camera's matrices (column-major storage: row 3 of inv_proj = m[3], m[7], m[11], m[15]) ----
BPT_HD float rb_linear_01(float depth, const bpt_camera& cam) { const float a = cam.matrix_inv_proj[11], b = cam.matrix_inv_proj[15]; return ((1.0f - depth) * b) / (a * depth + b); }
BPT_HD float3 rb_position_view(float u, float v, float depth, const bpt_camera& cam) {                                       // projection.hlsl:5-10
    const float* ip = cam.matrix_inv_proj;
    float nx = u * 2.0f - 1.0f, ny = 1.0f - v * 2.0f;
    float vx = ((ip[0] * nx + ip[4] * ny) + ip[8] * depth) + ip[12], vy = ((ip[1] * nx + ip[5] * ny) + ip[9] * depth) + ip[13];
    float vz = ((ip[2] * nx + ip[6] * ny) + ip[10] * depth) + ip[14], vw = ((ip[3] * nx + ip[7] * ny) + ip[11] * depth) + ip[15];
    return v3(vx / vw, vy / vw, vz / vw);
}
BPT_HD float3 rb_to_world(float3 p, const bpt_camera& cam) {
    const float* iv = cam.matrix_inv_view;
    return v3(((iv[0] * p.x + iv[4] * p.y) + iv[8] * p.z) + iv[12], ((iv[1] * p.x + iv[5] * p.y) + iv[9] * p.z) + iv[13], ((iv[2] * p.x + iv[6] * p.y) + iv[10] * p.z) + iv[14]);
}
BPT_HD float3 rb_position_world(float u, float v, float depth, const bpt_camera& cam) { return rb_to_world(rb_position_view(u, v, depth, cam), cam); }   // :22-26
BPT_HD float3 rb_camera_position(const bpt_camera& cam) { return v3(cam.matrix_inv_view[12], cam.matrix_inv_view[13], cam.matrix_inv_view[14]); }
// mul(matrix_proj_view, float4(p, 1)) / w -> (uv, ndc z)
BPT_HD float3 rb_project(float3 p, const bpt_camera& cam) {
    const float* m = cam.matrix_proj_view;
    float x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12], y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
    float z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14], w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15];
    x = x / w; y = y / w; z = z / w;
    return v3(x * 0.5f + 0.5f, 0.5f - y * 0.5f, z);
}
BPT_HD void rb_gbuffer_coord(const ReblurView& rv, int x, int y, int& gx, int& gy) {
    if (rv.half_res) { gx = x * 2 + (int)(rv.frame_index & 1u); gy = y * 2 + (int)((rv.frame_index >> 1) & 1u); }
    else { gx = x; gy = y; }
}

// ---- filter.hlsl ----
struct RbTap { float3 position, normal; float z01, roughness; };
BPT_HD RbTap rb_tap(const ReblurView& rv, int gx, int gy, float u, float v) {                                               // :40-61
    RbTap t;
    float depth = rb_load(rv.depth, rv.gw, rv.gh, gx, gy);
    t.z01 = rb_linear_01(depth, rv.cam);
    t.position = rb_position_world(u, v, depth, rv.cam);
    float4 nr = rb_load(rv.normal_roughness, rv.gw, rv.gh, gx, gy);
    t.normal = oct_decode(make_float2(nr.x, nr.y));
    t.roughness = nr.w;
    return t;
}
BPT_HD float rb_bilateral_weight(const RbTap& c, const RbTap& t) {                                                           // :63-77
    float w_depth = tmax_(0.0f, 1.0f - fabsf(t.z01 - c.z01));
    float nc = tmax_(0.0f, dot3(t.normal, c.normal));
    nc = nc * nc; nc = nc * nc;
    float w_normal = tmax_(0.0f, 1.0f - (1.0f - nc));
    float3 dq = c.position - t.position;
    float d2 = dot3(dq, dq);
    float plane_error = tmax_(fabsf(dot3(dq, t.normal)), fabsf(dot3(dq, c.normal)));
    float wp = tmax_(0.0f, 1.0f - (2.0f * plane_error) / sqrtf(d2));
    float w_plane = d2 < 0.0001f ? 1.0f : wp * wp;
    float w_rough = tmax_(0.0f, 1.0f - fabsf(t.roughness - c.roughness));
    return ((w_depth * w_normal) * w_plane) * w_rough;
}
BPT_HD float rb_gauss(float r) { return exp_neg((-0.66f * r) * r); }                                                          // :22-24
BPT_HD float rb_blur_radius(float roughness, float max_radius) { return max_radius * rb_magic_curve2(roughness, 0.75f); }   // :26-29
BPT_HD float3 rb_poisson(int i) {                                                                                             // :11-20
    const float d = 0.25f * 1.41421356237309504880f;
    switch (i) {
        case 0: return v3(-1.0f, 0.0f, 1.0f);
        case 1: return v3(0.0f, 1.0f, 1.0f);
        case 2: return v3(1.0f, 0.0f, 1.0f);
        case 3: return v3(0.0f, -1.0f, 1.0f);
        case 4: return v3(-d, d, 0.5f);
        case 5: return v3(d, d, 0.5f);
        case 6: return v3(d, -d, 0.5f);
        default: return v3(-d, -d, 0.5f);
    }
}
BPT_HD float3 rb_clamped_lighting(float4 c) {                                                                                 // pre_blur.hlsl:67-72
    float3 l = v3(c.x > 0.0f ? c.x : 0.0f, c.y > 0.0f ? c.y : 0.0f, c.z > 0.0f ? c.z : 0.0f);        // max(x, 0): NaN -> 0
    if (!is_finite3(l)) l = v3s(0.0f);
    float lum = (l.x * 0.212671f + l.y * 0.715160f) + l.z * 0.072169f;
    float scale = tmin_(lum, 1.5f) / tmax_(lum, 0.0001f);
    return l * scale;
}

// ---- pass 1: pre_blur.hlsl -> lighting_dist_1 ----
BPT_HD void reblur_pre_blur(const ReblurView& rv, int x, int y) {
    const float tsx = 1.0f / (float)rv.w, tsy = 1.0f / (float)rv.h;
    const float cu = ((float)x + 0.5f) * tsx, cv = ((float)y + 0.5f) * tsy;
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    float4* out = rv.ld1 + (size_t)y * rv.w + x;
    RbTap center = rb_tap(rv, gx, gy, ((float)gx + 0.5f) * tsx, ((float)gy + 0.5f) * tsy);
    if (center.z01 > 0.999f) { *out = make_float4(0.0f, 0.0f, 0.0f, -1.0f); return; }
    float center_dist = rb_load(rv.hit_positions, rv.w, rv.h, x, y).w;
    float radius = rb_blur_radius(center.roughness, 20.0f) * rv.blur_radius;
    float camera_dist = length3(center.position - rb_camera_position(rv.cam));
    radius = radius * rb_hit_dist_atten(center.roughness, camera_dist, center_dist);
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float sum_w = 0.0f;
    for (int i = 0; i < 8; i++) {
        float3 o = rb_poisson(i);
        float2 r = rb_rotate(rv.rot_pre, make_float2(o.x, o.y));
        float u = cu + (r.x * tsx) * radius, v = cv + (r.y * tsy) * radius;
        int tx = rb_trunc(u * (float)rv.w), ty = rb_trunc(v * (float)rv.h);
        tx = tx < 0 ? 0 : (tx > (int)rv.w - 1 ? (int)rv.w - 1 : tx); ty = ty < 0 ? 0 : (ty > (int)rv.h - 1 ? (int)rv.h - 1 : ty);
        int tgx, tgy; rb_gbuffer_coord(rv, tx, ty, tgx, tgy);
        RbTap tap = rb_tap(rv, tgx, tgy, ((float)tgx + 0.5f) * tsx, ((float)tgy + 0.5f) * tsy);
        float w = ((rb_gauss(o.z) * rb_bilateral_weight(center, tap)) * (tap.z01 < 0.999f ? 1.0f : 0.0f)) * ((sat(u) == u && sat(v) == v) ? 1.0f : 0.0f);
        float3 l = rb_clamped_lighting(rb_load(rv.noised, rv.w, rv.h, tx, ty));
        float dist = rb_load(rv.hit_positions, rv.w, rv.h, tx, ty).w;
        sum = rb_add(sum, rb_mul(make_float4(l.x, l.y, l.z, dist), w));
        sum_w = sum_w + w;
    }
    float4 blurred;
    if (sum_w != 0.0f) blurred = rb_div(sum, sum_w);
    else { float3 l = rb_clamped_lighting(rb_load(rv.noised, rv.w, rv.h, x, y)); blurred = make_float4(l.x, l.y, l.z, center_dist); }
    blurred.w = clampf_(blurred.w, 0.0f, 16.0f);
    *out = rb_q4(blurred);
}

// ---- pass 2: temporal_accumulate.hlsl -> lighting_dist_0 (level 0), accumulation ----
BPT_HD void reblur_temporal_accumulate(const ReblurView& rv, int x, int y) {
    const float tsx = 1.0f / (float)rv.w, tsy = 1.0f / (float)rv.h;
    const float cu = ((float)x + 0.5f) * tsx, cv = ((float)y + 0.5f) * tsy;
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    const size_t at = (size_t)y * rv.w + x;
    float depth = rb_load(rv.depth, rv.gw, rv.gh, gx, gy);
    uint32_t mask = rv.validation ? rv.validation[at] : 0u;
    float4 ld = rv.ld1[at];
    if (depth == 0.0f || mask != 0u) { rv.ld0[at] = ld; rv.accum[at] = 0.0f; return; }
    float4 nr = rb_load(rv.normal_roughness, rv.gw, rv.gh, x, y);                  // (pixel_coord, not the G-buffer coordinate: :34)
    float3 normal = oct_decode(make_float2(nr.x, nr.y));
    float roughness = nr.w;
    float3 position = rb_position_world(cu, cv, depth, rv.cam);
    float3 view = normalize3(rb_camera_position(rv.cam) - position);
    float ndotv = dot3(normal, view);
    float2 vel = rb_load(rv.velocity, rv.gw, rv.gh, gx, gy);
    float pu = cu - vel.x, pv = cv - vel.y;
    int pgx = rb_trunc(pu * (float)rv.w), pgy = rb_trunc(pv * (float)rv.h);       // prev_uv * gbuffer_tex_size (= tex_size)
    float hist_depth = rv.has_history ? rb_load(rv.hist_depth, rv.gw, rv.gh, pgx, pgy) : 0.0f;
    float3 hist_position = rb_position_world(pu, pv, hist_depth, rv.hist_cam);
    float3 hist_view = normalize3(rb_camera_position(rv.hist_cam) - hist_position);
    float parallax = rb_parallax(view, hist_view);
    const float4* h0 = rv.has_history ? rv.hist_ld0 : nullptr;
    float4 hist_ld = rb_sample4(h0, rv.w, rv.h, pu, pv);
    float accum_hist = rb_sample1(rv.has_history ? rv.hist_accum : nullptr, rv.w, rv.h, pu, pv);
    float accum = rb_accum_speed(roughness, ndotv, parallax);
    accum = tmin_(tmin_(accum, accum_hist), kReblurMaxAccum);
    float4 lerped = rb_mix(hist_ld, ld, 1.0f / (1.0f + accum));
    if (rv.virtual_history != 0u) {
        float f = rb_dominant_factor(ndotv, roughness);                            // get_virtual_position, utils.hlsl:93-96
        float3 vp = position - (view * lerped.w) * f;
        float3 c = rb_project(vp, rv.hist_cam);
        bool depth_ok = c.z >= -1.0f && c.z <= 1.0f;
        if (sat(c.x) == c.x && sat(c.y) == c.y && depth_ok) {
            int vgx = rb_trunc(c.x * (float)rv.w), vgy = rb_trunc(c.y * (float)rv.h);
            float hvd = rv.has_history ? rb_load(rv.hist_depth, rv.gw, rv.gh, vgx, vgy) : 0.0f;
            float4 hv = rb_sample4(h0, rv.w, rv.h, c.x, c.y);
            float amount = rb_dominant_factor(ndotv, roughness);
            const float confidence = 1.0f;
            float lin = rb_linear_01(depth, rv.cam), vlin = rb_linear_01(hvd, rv.hist_cam);
            amount = amount * (fabsf(lin - vlin) < lin * 0.1f ? 1.0f : 0.0f);
            float4 hnr = rv.has_history ? rb_load(rv.hist_normal_roughness, rv.gw, rv.gh, vgx, vgy) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float3 hn = oct_decode(make_float2(hnr.x, hnr.y));
            amount = amount * (dot3(normal, hn) > 0.9f ? 1.0f : 0.0f);
            float a_virtual = rb_accum_speed(roughness, ndotv, 0.0f);
            float a_min = tmin_(a_virtual, kReblurMipLevels * sqrtf(roughness));
            float a = mix1(1.0f / (1.0f + a_min), 1.0f / (1.0f + a_virtual), confidence);
            a_virtual = 1.0f / a - 1.0f;
            float a_hit = tmin_(a_virtual, kReblurMaxAccum);
            float wl = 1.0f / (1.0f + a_virtual), wd = 1.0f / (1.0f + a_hit);
            hv = make_float4(mix1(hv.x, ld.x, wl), mix1(hv.y, ld.y, wl), mix1(hv.z, ld.z, wl), mix1(hv.w, ld.w, wd));
            float4 result = rb_mix(lerped, hv, amount);
            a = mix1(1.0f / (1.0f + accum), 1.0f / (1.0f + a_virtual), amount);
            accum = 1.0f / a - 1.0f;
            lerped = result;
        }
    }
    rv.ld0[at] = rb_q4(lerped);
    rv.accum[at] = q_half(accum);
}

// ---- pass 3: fetch_linear_depth.hlsl -> linear depth level 0 ----
BPT_HD void reblur_fetch_linear_depth(const ReblurView& rv, int x, int y) {
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    rv.lin_depth[(size_t)y * rv.w + x] = rb_linear_01(rb_load(rv.depth, rv.gw, rv.gh, gx, gy), rv.cam);
}

// ---- pass 4: gen_depth_mip.hlsl: one 16 x 16 tile -> levels 1..3 of lighting_dist_0 and of linear depth. The coarser levels are
//      built from the UNROUNDED level below (the shader keeps them in groupshared memory), only the stored copies are halves. ----
BPT_HD float rb_mip_weight(float d, float dm, float4 v) { return (fabsf(d - dm) < dm * 0.1f && v.w > 0.0f) ? 1.0f : 0.0f; }
BPT_HD void rb_mip_reduce(const float4 v[4], const float d[4], float4& out_v, float& out_d) {
    float dm = tmin_(tmin_(d[0], d[1]), tmin_(d[2], d[3]));
    float w0 = rb_mip_weight(d[0], dm, v[0]), w1 = rb_mip_weight(d[1], dm, v[1]), w2 = rb_mip_weight(d[2], dm, v[2]), w3 = rb_mip_weight(d[3], dm, v[3]);
    float ws = ((w0 + w1) + w2) + w3;
    out_d = dm;
    if (ws == 0.0f) out_v = rb_mul(rb_add(rb_add(rb_add(v[0], v[1]), v[2]), v[3]), 0.25f);
    else out_v = rb_div(rb_add(rb_add(rb_add(rb_mul(v[0], w0), rb_mul(v[1], w1)), rb_mul(v[2], w2)), rb_mul(v[3], w3)), ws);
}
// The 2x2 source block of thread `local` (0..63) of tile (tile_x, tile_y): pixel = tile * 16 + 2 * (even bits, odd bits of local)
BPT_HD void rb_mip_level1(const ReblurView& rv, int tile_x, int tile_y, uint32_t local, float4& v, float& d, int& px, int& py) {
    auto even = [](uint32_t b) { b &= 0x55555555u; b = (b | (b >> 1)) & 0x33333333u; b = (b | (b >> 2)) & 0x0f0f0f0fu; b = (b | (b >> 4)) & 0x00ff00ffu; b = (b | (b >> 8)) & 0x0000ffffu; return b; };
    px = tile_x * 16 + (int)even(local) * 2; py = tile_y * 16 + (int)even(local >> 1) * 2;
    float4 vv[4]; float dd[4];
    for (int k = 0; k < 4; k++) {
        int sx = px + (k & 1), sy = py + (k >> 1);
        sx = sx < 0 ? 0 : (sx > (int)rv.w - 1 ? (int)rv.w - 1 : sx); sy = sy < 0 ? 0 : (sy > (int)rv.h - 1 ? (int)rv.h - 1 : sy);      // safe_texel_fetch
        vv[k] = rv.ld0[(size_t)sy * rv.w + sx]; dd[k] = rv.lin_depth[(size_t)sy * rv.w + sx];
    }
    rb_mip_reduce(vv, dd, v, d);
}
BPT_HD void rb_mip_store(const ReblurView& rv, int level, int x, int y, float4 v, float d) {
    const uint32_t lw = rv.w >> level, lh = rv.h >> level;
    if (x < (int)lw && y < (int)lh) {
        const size_t o = reblur_mip_offset(rv.w, rv.h, level) + (size_t)y * lw + x;
        rv.ld0[o] = rb_q4(v); rv.lin_depth[o] = d;
    }
}

// ---- pass 5: fix_history.hlsl -> lighting_dist_1 (pixels it does not write keep the pre-blur value) ----
BPT_HD void reblur_fix_history(const ReblurView& rv, int x, int y) {
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    const size_t at = (size_t)y * rv.w + x;
    float lin = rv.lin_depth[at];
    if (lin > 0.999f) return;
    float4 ld = rv.ld0[at];
    float accum = rv.accum[at];
    float norm_accum = sat(accum / kReblurMaxFixFrames);
    if (norm_accum == 1.0f) { rv.ld1[at] = ld; return; }
    float roughness = rb_load(rv.normal_roughness, rv.gw, rv.gh, gx, gy).w;
    int mip = rb_trunc((kReblurMipLevels * (1.0f - norm_accum)) * roughness);
    mip = mip < 3 ? mip : 3;
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float sum_w = 0.0f;
    while (mip >= 0) {
        const int mx = x >> mip, my = y >> mip;
        const uint32_t lw = rv.w >> mip, lh = rv.h >> mip;
        const size_t off = reblur_mip_offset(rv.w, rv.h, mip);
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                float td = rb_load(rv.lin_depth + off, lw, lh, mx + dx, my + dy);
                float4 tv = rb_load((const float4*)(rv.ld0 + off), lw, lh, mx + dx, my + dy);
                if (fabsf(lin - td) < lin * 0.1f && tv.w > 0.0f) {
                    float r = sqrtf((float)(dx * dx + dy * dy));                     // length(float2(dx, dy)); gaussian(r, 1) = exp(-r^2)
                    float w = exp_neg(-(r * r));
                    sum = rb_add(sum, rb_mul(tv, w));
                    sum_w = sum_w + w;
                }
            }
        if (sum_w > 3.5f) break;
        --mip;
    }
    rv.ld1[at] = rb_q4((sum_w == 0.0f || mip < 0) ? ld : rb_div(sum, sum_w));
}

// ---- passes 6 and 8: blur.hlsl (world-space kernel) -> lighting_dist_0 level 0; post_blur.hlsl (screen-space kernel) -> denoised ----
BPT_HD void reblur_blur(const ReblurView& rv, int x, int y) {
    const float tsx = 1.0f / (float)rv.w, tsy = 1.0f / (float)rv.h;
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    const size_t at = (size_t)y * rv.w + x;
    RbTap center = rb_tap(rv, gx, gy, ((float)gx + 0.5f) * tsx, ((float)gy + 0.5f) * tsy);
    if (center.z01 > 0.999f) { rv.ld0[at] = make_float4(0.0f, 0.0f, 0.0f, -1.0f); return; }
    float4 cl = rv.ld1[at];
    float accum = rv.accum[at];
    float3 view_vec = rb_camera_position(rv.cam) - center.position;
    float camera_dist = length3(view_vec);
    float3 view = view_vec / camera_dist;
    float4 dominant = rb_dominant_direction(center.normal, view, center.roughness);
    float radius = (((rb_blur_radius(center.roughness, 0.04f) * rv.blur_radius) * (1.0f - sat(accum / kReblurMaxAccum))) *
                    rb_hit_dist_atten(center.roughness, camera_dist, cl.w)) * sat((camera_dist - 0.03f) / 0.05f);
    float3 T, B;
    rb_kernel_basis(v3(dominant.x, dominant.y, dominant.z), center.normal, center.roughness, T, B);     // (V := the dominant direction, blur.hlsl:51)
    T = T * radius; B = B * radius;
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float sum_w = 0.0f;
    for (int i = 0; i < 8; i++) {
        float3 o = rb_poisson(i);
        float2 r = rb_rotate(rv.rot_blur, make_float2(o.x, o.y));
        float3 pi = (center.position + T * r.x) + B * r.y;
        float3 c = rb_project(pi, rv.cam);
        int tx = rb_trunc(c.x * (float)rv.w), ty = rb_trunc(c.y * (float)rv.h);
        tx = tx < 0 ? 0 : (tx > (int)rv.w - 1 ? (int)rv.w - 1 : tx); ty = ty < 0 ? 0 : (ty > (int)rv.h - 1 ? (int)rv.h - 1 : ty);
        int tgx, tgy; rb_gbuffer_coord(rv, tx, ty, tgx, tgy);
        RbTap tap = rb_tap(rv, tgx, tgy, c.x, c.y);
        float w = ((rb_gauss(o.z) * rb_bilateral_weight(center, tap)) * (tap.z01 < 0.999f ? 1.0f : 0.0f)) * ((sat(c.x) == c.x && sat(c.y) == c.y) ? 1.0f : 0.0f);
        sum = rb_add(sum, rb_mul(rv.ld1[(size_t)ty * rv.w + tx], w));
        sum_w = sum_w + w;
    }
    rv.ld0[at] = rb_q4(sum_w == 0.0f ? cl : rb_div(sum, sum_w));
}
BPT_HD void reblur_post_blur(const ReblurView& rv, int x, int y) {
    const float tsx = 1.0f / (float)rv.w, tsy = 1.0f / (float)rv.h;
    const float cu = ((float)x + 0.5f) * tsx, cv = ((float)y + 0.5f) * tsy;
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    const size_t at = (size_t)y * rv.w + x;
    RbTap center = rb_tap(rv, gx, gy, ((float)gx + 0.5f) * tsx, ((float)gy + 0.5f) * tsy);
    if (center.z01 > 0.999f) { rv.denoised[at] = make_float4(0.0f, 0.0f, 0.0f, -1.0f); return; }
    float4 cl = rv.ld1[at];
    float accum = rv.accum[at];
    float camera_dist = length3(center.position - rb_camera_position(rv.cam));
    float radius = ((rb_blur_radius(center.roughness, 15.0f) * rv.blur_radius) * (1.0f - sat(accum / kReblurMaxAccum))) *
                   rb_hit_dist_atten(center.roughness, camera_dist, cl.w);
    float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float sum_w = 0.0f;
    for (int i = 0; i < 8; i++) {
        float3 o = rb_poisson(i);
        float2 r = rb_rotate(rv.rot_post, make_float2(o.x, o.y));
        float u = cu + (r.x * tsx) * radius, v = cv + (r.y * tsy) * radius;
        int tx = rb_trunc(u * (float)rv.w), ty = rb_trunc(v * (float)rv.h);
        tx = tx < 0 ? 0 : (tx > (int)rv.w - 1 ? (int)rv.w - 1 : tx); ty = ty < 0 ? 0 : (ty > (int)rv.h - 1 ? (int)rv.h - 1 : ty);
        int tgx, tgy; rb_gbuffer_coord(rv, tx, ty, tgx, tgy);
        RbTap tap = rb_tap(rv, tgx, tgy, ((float)tgx + 0.5f) * tsx, ((float)tgy + 0.5f) * tsy);
        float w = ((rb_gauss(o.z) * rb_bilateral_weight(center, tap)) * (tap.z01 < 0.999f ? 1.0f : 0.0f)) * ((sat(u) == u && sat(v) == v) ? 1.0f : 0.0f);
        sum = rb_add(sum, rb_mul(rv.ld1[(size_t)ty * rv.w + tx], w));
        sum_w = sum_w + w;
    }
    rv.denoised[at] = rb_q4(sum_w == 0.0f ? cl : rb_div(sum, sum_w));
}

// ---- pass 7: temporal_stabilize.hlsl: lighting_dist_0 (blurred) + last frame's stabilised image -> lighting_dist_1 ----
BPT_HD float3 rb_rgb_to_ycocg(float3 c) { return v3((0.25f * c.x + 0.5f * c.y) + 0.25f * c.z, 0.5f * c.x - 0.5f * c.z, (-0.25f * c.x + 0.5f * c.y) - 0.25f * c.z); }   // color.hlsl:23-29
BPT_HD float3 rb_ycocg_to_rgb(float3 c) { return v3((c.x + c.y) - c.z, c.x + c.z, (c.x - c.y) - c.z); }                                                                 // color.hlsl:31-37
BPT_HD float4 rb_clip_aabb(float4 inside, float4 p, float4 lo, float4 hi) {                                                                                              // math.hlsl:106-114
    const float eps = 1.0f / 65536.0f;
    float d[4] = {p.x - inside.x, p.y - inside.y, p.z - inside.z, p.w - inside.w};
    float in_[4] = {inside.x, inside.y, inside.z, inside.w}, l[4] = {lo.x, lo.y, lo.z, lo.w}, h[4] = {hi.x, hi.y, hi.z, hi.w};
    float inter[3];
    for (int k = 0; k < 3; k++) {
        float inv = 1.0f / d[k];
        float dir_inv = mix1(inv, eps, fabsf(d[k]) < eps ? 1.0f : 0.0f);     // lerp(1 / dir, 1 / 65536, |dir| < 1 / 65536) as written: inv + (eps - inv) * t
        float imax = (h[k] - in_[k]) * dir_inv, imin = (l[k] - in_[k]) * dir_inv;
        inter[k] = imax > imin ? imax : (imin > imax ? imin : (imax == imax ? imax : imin));      // max(a, b): the non-NaN operand
    }
    auto min_nan = [](float a, float b) { return a != a ? b : (b != b ? a : (a < b ? a : b)); };     // HLSL min: the non-NaN operand
    float m = min_nan(inter[0], min_nan(inter[1], inter[2]));
    float t = m != m ? 0.0f : sat(m);                                          // saturate(NaN) = 0
    return rb_mix(inside, p, t);
}
// the 3x3 neighbourhood texel (x + dx, y + dy) as fill_shared_data reads it: clamped to the image, colour as YCoCg
BPT_HD void rb_stab_tap(const ReblurView& rv, int x, int y, float4& ld, float& depth, float& lin) {
    x = x < 0 ? 0 : (x > (int)rv.w - 1 ? (int)rv.w - 1 : x); y = y < 0 ? 0 : (y > (int)rv.h - 1 ? (int)rv.h - 1 : y);
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    float4 v = rv.ld0[(size_t)y * rv.w + x];
    float3 c = rb_rgb_to_ycocg(v3(v.x, v.y, v.z));
    ld = make_float4(c.x, c.y, c.z, v.w);
    depth = rb_load(rv.depth, rv.gw, rv.gh, gx, gy);
    lin = rb_linear_01(depth, rv.cam);
}
BPT_HD void reblur_temporal_stabilize(const ReblurView& rv, int x, int y) {
    const float tsx = 1.0f / (float)rv.w, tsy = 1.0f / (float)rv.h;
    const float cu = ((float)x + 0.5f) * tsx, cv = ((float)y + 0.5f) * tsy;
    int gx, gy; rb_gbuffer_coord(rv, x, y, gx, gy);
    const size_t at = (size_t)y * rv.w + x;
    float4 n[9]; float nd[9], nl[9];
    // shared_data_index clamps the LINEAR index of the 10 x 10 tile, which only matters outside the tile: never for offsets in [-1, 1]
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) rb_stab_tap(rv, x + dx, y + dy, n[(dy + 1) * 3 + dx + 1], nd[(dy + 1) * 3 + dx + 1], nl[(dy + 1) * 3 + dx + 1]);
    const float center_depth = nd[4];
    if (center_depth == 0.0f) { rv.ld1[at] = make_float4(0.0f, 0.0f, 0.0f, -1.0f); return; }
    uint32_t mask = rv.validation ? rv.validation[at] : 0u;
    if (mask != 0u) { float3 c = rb_ycocg_to_rgb(v3(n[4].x, n[4].y, n[4].z)); rv.ld1[at] = rb_q4(make_float4(c.x, c.y, c.z, n[4].w)); return; }
    float roughness = rb_load(rv.normal_roughness, rv.gw, rv.gh, gx, gy).w;
    float3 pv = rb_position_view(cu, cv, center_depth, rv.cam);
    float camera_dist = length3(pv);
    float2 vel = rb_load(rv.velocity, rv.gw, rv.gh, gx, gy);
    float pu = ((float)x + 0.5f) * tsx - vel.x, pvv = ((float)y + 0.5f) * tsy - vel.y;
    float4 prev = rb_sample4(rv.has_history ? rv.hist_ld1 : nullptr, rv.w, rv.h, pu, pvv);
    { float3 c = rb_rgb_to_ycocg(v3(prev.x, prev.y, prev.z)); prev = make_float4(c.x, c.y, c.z, prev.w); }
    const float cl = nl[4];
    for (int k = 0; k < 9; k++)
        if (k != 4 && (nd[k] == 0.0f || fabsf(nl[k] - cl) > cl * 0.2f)) n[k] = n[4];
    float vlen = sqrtf((vel.x * (float)rv.w) * (vel.x * (float)rv.w) + (vel.y * (float)rv.h) * (vel.y * (float)rv.h));
    float atten = rb_hit_dist_atten(roughness, camera_dist, n[4].w);
    float params = rv.anti_flicker * atten;
    float4 m1 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), m2 = m1;
    for (int k = 0; k < 9; k++) {
        m1 = rb_add(m1, n[k]);
        m2 = rb_add(m2, make_float4(n[k].x * n[k].x, n[k].y * n[k].y, n[k].z * n[k].z, n[k].w * n[k].w));
    }
    m1 = rb_div(m1, 9.0f); m2 = rb_div(m2, 9.0f);
    float4 sd = make_float4(sqrtf(fabsf(m2.x - m1.x * m1.x)), sqrtf(fabsf(m2.y - m1.y * m1.y)), sqrtf(fabsf(m2.z - m1.z * m1.z)), sqrtf(fabsf(m2.w - m1.w * m1.w)));
    float localized = mix1(params * 0.8f, params * 2.25f, sat(1.0f - 2.0f * vlen));
    float mult = 1.5f + localized;
    mult = mix1(mult, 0.75f, sat(vlen / 50.0f));
    float4 lo = make_float4(m1.x - sd.x * mult, m1.y - sd.y * mult, m1.z - sd.z * mult, m1.w - sd.w * mult);
    float4 hi = make_float4(m1.x + sd.x * mult, m1.y + sd.y * mult, m1.z + sd.z * mult, m1.w + sd.w * mult);
    prev = rb_clip_aabb(n[4], prev, lo, hi);
    prev = rb_mix(prev, n[4], 0.05f);
    float3 c = rb_ycocg_to_rgb(v3(prev.x, prev.y, prev.z));
    rv.ld1[at] = rb_q4(make_float4(c.x, c.y, c.z, prev.w));
}

} // namespace bptd
