// oracle_reblur.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked, imported or executed by the product.
//
// ReBLUR, the denoiser of the ray-traced reflections: a pass-by-pass restatement of ReblurPass::render
// (bisemutum/src/renderer/pass/reblur.cpp:273-588) and of shaders/renderer/reblur/*.hlsl. Each pass below is a loop over the image that
// reads and writes whole textures, in the order the render graph runs them; every texture has the reference's format (a store to an
// rgba16_sfloat / r16_sfloat target rounds to half). PARITY UNPINNED: the reference has no test or golden image for this pass and cannot
// be run here (HLSL + Vulkan); tests/test_reblur.py pins single functions against float64 numpy and the pass against its invariants.
//
// HLSL intrinsics without a bit-defined result (atan, log, pow, exp, exp2) use the fixed-order forms of oracle_math.hpp so that
// two compilers agree; lerp(a, b, t) is a + (b - a) * t; min / max / saturate drop a NaN operand as HLSL does.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "oracle.h"
#include "oracle_scene.hpp"

using namespace orc;

namespace {

struct Img4 {                     // an RGBA texture level; Load outside returns 0
    uint32_t w = 0, h = 0; std::vector<f4> px;
    void resize(uint32_t ww, uint32_t hh) { w = ww; h = hh; px.assign((size_t)ww * hh, f4{0, 0, 0, 0}); }
    f4 load(int x, int y) const { return (x >= 0 && y >= 0 && x < (int)w && y < (int)h) ? px[(size_t)y * w + x] : f4{0, 0, 0, 0}; }
    f4& at(int x, int y) { return px[(size_t)y * w + x]; }
    bool empty() const { return px.empty(); }
};
struct Img1 {
    uint32_t w = 0, h = 0; std::vector<float> px;
    void resize(uint32_t ww, uint32_t hh) { w = ww; h = hh; px.assign((size_t)ww * hh, 0.0f); }
    float load(int x, int y) const { return (x >= 0 && y >= 0 && x < (int)w && y < (int)h) ? px[(size_t)y * w + x] : 0.0f; }
    float& at(int x, int y) { return px[(size_t)y * w + x]; }
    bool empty() const { return px.empty(); }
};
f4 add4(f4 a, f4 b) { return f4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
f4 scale4(f4 a, float s) { return f4{a.x * s, a.y * s, a.z * s, a.w * s}; }
f4 over4(f4 a, float s) { return f4{a.x / s, a.y / s, a.z / s, a.w / s}; }
f4 lerp4(f4 a, f4 b, float t) { return f4{lerpf(a.x, b.x, t), lerpf(a.y, b.y, t), lerpf(a.z, b.z, t), lerpf(a.w, b.w, t)}; }
f4 half4(f4 v) { return f4{store_half(v.x), store_half(v.y), store_half(v.z), store_half(v.w)}; }
float len3(f3 v) { return sqrtf(dot(v, v)); }
int to_int(float f) { return f != f ? 0 : (f >= 2147483520.0f ? 2147483647 : (f <= -2147483648.0f ? (int)0x80000000 : (int)f)); }      // HLSL int(x)
int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// linear / clamp-to-edge SampleLevel(uv, 0) of a w x h texture (reblur.cpp:242-248); an unbound history texture reads 0
template <class Img, class T, class Lerp>
T bilinear(const Img& t, float u, float v, T zero, Lerp lerp) {
    if (t.empty()) return zero;
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float xf = floorf(x), yf = floorf(y);
    const float fx = x - xf, fy = y - yf;
    xf = !(xf >= -2.0f) ? -2.0f : (xf > (float)t.w ? (float)t.w : xf); yf = !(yf >= -2.0f) ? -2.0f : (yf > (float)t.h ? (float)t.h : yf);
    const int x0 = clampi((int)xf, 0, (int)t.w - 1), x1 = clampi((int)(xf + 1.0f), 0, (int)t.w - 1);
    const int y0 = clampi((int)yf, 0, (int)t.h - 1), y1 = clampi((int)(yf + 1.0f), 0, (int)t.h - 1);
    return lerp(lerp(t.px[(size_t)y0 * t.w + x0], t.px[(size_t)y0 * t.w + x1], fx), lerp(t.px[(size_t)y1 * t.w + x0], t.px[(size_t)y1 * t.w + x1], fx), fy);
}
f4 sample4(const Img4& t, float u, float v) { return bilinear(t, u, v, f4{0, 0, 0, 0}, lerp4); }
float sample1(const Img1& t, float u, float v) { return bilinear(t, u, v, 0.0f, lerpf); }

// ---- utils.hlsl -------------------------------------------------------------------------------------------------
float pow_01(float x, float y) {            // pow(x, y), 0 <= x <= 1, y > 0
    if (!(x > 0.0f)) return 0.0f;
    if (x >= 1.0f) return 1.0f;
    return exp_neg((y * log2_(x)) * 0.693147181f);
}
f2 rotate_vector(f4 rotator, f2 v) { return f2{v.x * rotator.x + v.y * rotator.z, v.x * rotator.y + v.y * rotator.w}; }                  // :12-14
float get_specular_lobe_half_angle(float roughness, float percentage) { float m = roughness * roughness; return atan2_((m * percentage) / (1.0f - percentage), 1.0f); }   // :16-19
float get_specular_magic_curve2(float roughness, float percentage) {                                                                    // :21-25
    return saturate(get_specular_lobe_half_angle(roughness, percentage) / get_specular_lobe_half_angle(1.0f, percentage));
}
float get_specular_dominant_factor(float ndotv, float roughness) {                                                                      // :27-31
    float a = 0.298475f * (log2_(39.4115f - 39.0029f * roughness) * 0.693147181f);
    return saturate(pow_01(saturate(1.0f - ndotv), 10.8649f) * (1.0f - a) + a);
}
f4 get_specular_dominant_direction(f3 N, f3 V, float roughness) {                                                                       // :33-44
    float factor = get_specular_dominant_factor(fabsf(dot(N, V)), roughness);
    f3 D = normalize(lerp3(N, reflect(-V, N), factor));
    return f4{D.x, D.y, D.z, factor};
}
void get_kernel_basis(f3 V, f3 N, float roughness, f3& T, f3& B) {                                                                       // :46-66
    Frame basis = create_frame(N);
    T = basis.x; B = basis.y;
    f4 d4 = get_specular_dominant_direction(N, V, roughness);
    f3 D = mk3(d4.x, d4.y, d4.z);
    if (fabsf(dot(N, D)) < 0.999f && roughness != 1.0f) {
        f3 d_reflected = reflect(-D, N);
        T = normalize(cross(N, d_reflected));
        B = cross(d_reflected, T);
        float acos01sq = saturate(1.0f - fabsf(dot(N, V)));
        T = T * lerpf(1.0f, roughness, sqrtf(acos01sq));
    }
}
float calc_parallax(f3 curr_view, f3 prev_view) { float c = saturate(dot(curr_view, prev_view)); return (sqrtf(1.0f - c * c) / fmax_(c, 0.00001f)) * 60.0f; }   // :68-71
float get_specular_accum_speed(float roughness, float ndotv, float parallax) {                                                         // :73-86
    float a = sqrtf(saturate(1.0f - ndotv));                         // pow(., SPEC_ACCUM_CURVE = 0.5)
    float b = 1.1f + roughness * roughness;
    float power_scale = 1.0f + parallax * ((b + a) / (b - a));
    float f = 1.0f - exp_neg(((-200.0f * roughness) * roughness) * 0.693147181f);
    f = f * pow_01(saturate(roughness), 1.0f * power_scale);
    return 32.0f * f;
}
float hit_distance_attenuation(float roughness, float camera_dist, float hit_dist) {                                                   // :88-91
    float f = hit_dist < 0.0f ? 1.0f : hit_dist / (hit_dist + camera_dist);
    return lerpf(0.5f * roughness, 1.0f, f);
}

// ---- camera: depth.hlsl:31-35, projection.hlsl:5-26, camera.hlsl:7-13 (glm column-major storage) ----------------------------------------
struct Cam { bpt_camera m; };
float linear_01(const bpt_camera& c, float depth) { return ((1.0f - depth) * c.matrix_inv_proj[15]) / (c.matrix_inv_proj[11] * depth + c.matrix_inv_proj[15]); }
f3 position_view(const bpt_camera& c, float u, float v, float depth) {
    const float* ip = c.matrix_inv_proj;
    const float nx = u * 2.0f - 1.0f, ny = 1.0f - v * 2.0f;
    const float x = ((ip[0] * nx + ip[4] * ny) + ip[8] * depth) + ip[12], y = ((ip[1] * nx + ip[5] * ny) + ip[9] * depth) + ip[13];
    const float z = ((ip[2] * nx + ip[6] * ny) + ip[10] * depth) + ip[14], w = ((ip[3] * nx + ip[7] * ny) + ip[11] * depth) + ip[15];
    return mk3(x / w, y / w, z / w);
}
f3 position_world(const bpt_camera& c, float u, float v, float depth) {
    const float* iv = c.matrix_inv_view;
    f3 p = position_view(c, u, v, depth);
    return mk3(((iv[0] * p.x + iv[4] * p.y) + iv[8] * p.z) + iv[12], ((iv[1] * p.x + iv[5] * p.y) + iv[9] * p.z) + iv[13], ((iv[2] * p.x + iv[6] * p.y) + iv[10] * p.z) + iv[14]);
}
f3 camera_position(const bpt_camera& c) { return mk3(c.matrix_inv_view[12], c.matrix_inv_view[13], c.matrix_inv_view[14]); }
f3 project_uv(const bpt_camera& c, f3 p) {                    // mul(matrix_proj_view, float4(p, 1)) / w -> (u, v, ndc z)
    const float* m = c.matrix_proj_view;
    float x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12], y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
    float z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14], w = ((m[3] * p.x + m[7] * p.y) + m[11] * p.z) + m[15];
    x = x / w; y = y / w; z = z / w;
    return mk3(x * 0.5f + 0.5f, 0.5f - y * 0.5f, z);
}

struct State {                   // what survives a frame (the camera's history textures + matrices)
    bool valid = false; uint64_t last_frame = 0; bpt_camera cam{};
    uint32_t w = 0, h = 0, gw = 0, gh = 0;
    Img1 depth; Img4 normal_roughness; Img4 ld0, ld1; Img1 accum;
    // debug copies of the last frame's working textures
    std::vector<Img4> dbg_ld0_mips; std::vector<Img1> dbg_depth_mips; Img4 dbg_ld1; Img1 dbg_accum;
};

struct Frame_ {                  // one frame's inputs and constants
    uint32_t w, h, gw, gh; bool half; uint32_t frame_index; bool has_history; uint32_t virtual_history; float blur_radius, anti_flicker;
    bpt_camera cam, hist_cam; f4 rot_pre, rot_blur, rot_post;
    Img1 depth; Img4 normal_roughness; std::vector<f2> velocity; std::vector<uint8_t> validation; Img4 hit, noised;
    void gcoord(int x, int y, int& gx, int& gy) const {
        if (half) { gx = x * 2 + (int)(frame_index & 1u); gy = y * 2 + (int)((frame_index >> 1) & 1u); } else { gx = x; gy = y; }
    }
    f2 vel(int gx, int gy) const { return (!velocity.empty() && gx >= 0 && gy >= 0 && gx < (int)gw && gy < (int)gh) ? velocity[(size_t)gy * gw + gx] : f2{0.0f, 0.0f}; }
    uint32_t mask(int x, int y) const { return validation.empty() ? 0u : validation[(size_t)y * w + x]; }
};

// ---- filter.hlsl ---------------------------------------------------------------------------------------------------
struct BilateralData { f3 position, normal; float z_01, roughness; };
BilateralData tap_bilateral_data(const Frame_& f, int gx, int gy, float u, float v) {                                                  // :40-61
    BilateralData d;
    const float depth = f.depth.load(gx, gy);
    d.z_01 = linear_01(f.cam, depth);
    d.position = position_world(f.cam, u, v, depth);
    const f4 nr = f.normal_roughness.load(gx, gy);
    d.normal = oct_decode(f2{nr.x, nr.y});
    d.roughness = nr.w;
    return d;
}
float calc_bilateral_weight(const BilateralData& c, const BilateralData& t) {                                                           // :63-77
    const float w_depth = fmax_(0.0f, 1.0f - fabsf(t.z_01 - c.z_01));
    float closeness = fmax_(0.0f, dot(t.normal, c.normal));
    closeness = closeness * closeness; closeness = closeness * closeness;
    const float w_normal = fmax_(0.0f, 1.0f - (1.0f - closeness));
    const f3 dq = c.position - t.position;
    const float dist_sqr = dot(dq, dq);
    const float plane_error = fmax_(fabsf(dot(dq, t.normal)), fabsf(dot(dq, c.normal)));
    const float p = fmax_(0.0f, 1.0f - (2.0f * plane_error) / sqrtf(dist_sqr));
    const float w_plane = dist_sqr < 0.0001f ? 1.0f : p * p;
    const float w_roughness = fmax_(0.0f, 1.0f - fabsf(t.roughness - c.roughness));
    return ((w_depth * w_normal) * w_plane) * w_roughness;
}
float get_gaussian_weight(float r) { return exp_neg((-0.66f * r) * r); }                                                                 // :22-24
float calc_blur_radius(float roughness, float max_radius) { return max_radius * get_specular_magic_curve2(roughness, 0.75f); }          // :26-29
const float kD = 0.25f * 1.41421356237309504880f;
const f3 poisson_disk_samples[8] = {{-1.0f, 0.0f, 1.0f}, {0.0f, 1.0f, 1.0f}, {1.0f, 0.0f, 1.0f}, {0.0f, -1.0f, 1.0f}, {-kD, kD, 0.5f}, {kD, kD, 0.5f}, {kD, -kD, 0.5f}, {-kD, -kD, 0.5f}};
f3 clamp_lighting(f4 c) {                                                                                                                // pre_blur.hlsl:67-72
    f3 l = mk3(c.x > 0.0f ? c.x : 0.0f, c.y > 0.0f ? c.y : 0.0f, c.z > 0.0f ? c.z : 0.0f);
    if (!finite3(l)) l = splat3(0.0f);
    const float lum = (l.x * 0.212671f + l.y * 0.715160f) + l.z * 0.072169f;
    return l * (fmin_(lum, 1.5f) / fmax_(lum, 0.0001f));
}
float on_screen(float u, float v) { return (saturate(u) == u && saturate(v) == v) ? 1.0f : 0.0f; }

// screen-space Poisson blur shared by "Pre Blur" (source: the noised colour + hit distance) and "Post Blur" (source: lighting_dist_1)
template <class Source>
f4 screen_blur(const Frame_& f, int x, int y, const BilateralData& center, float blur_radius, f4 rotator, Source source, float& sum_weight) {
    const float tsx = 1.0f / (float)f.w, tsy = 1.0f / (float)f.h;
    const float cu = ((float)x + 0.5f) * tsx, cv = ((float)y + 0.5f) * tsy;
    f4 sum{0, 0, 0, 0};
    sum_weight = 0.0f;
    for (int i = 0; i < 8; i++) {
        const f3 offset = poisson_disk_samples[i];
        const f2 r = rotate_vector(rotator, f2{offset.x, offset.y});
        const float u = cu + (r.x * tsx) * blur_radius, v = cv + (r.y * tsy) * blur_radius;
        const int tx = clampi(to_int(u * (float)f.w), 0, (int)f.w - 1), ty = clampi(to_int(v * (float)f.h), 0, (int)f.h - 1);
        int tgx, tgy; f.gcoord(tx, ty, tgx, tgy);
        const BilateralData tap = tap_bilateral_data(f, tgx, tgy, ((float)tgx + 0.5f) * tsx, ((float)tgy + 0.5f) * tsy);
        const float w = ((get_gaussian_weight(offset.z) * calc_bilateral_weight(center, tap)) * (tap.z_01 < 0.999f ? 1.0f : 0.0f)) * on_screen(u, v);
        sum = add4(sum, scale4(source(tx, ty), w));
        sum_weight = sum_weight + w;
    }
    return sum;
}

// ---- the eight passes ---------------------------------------------------------------------------------------------
void pre_blur(const Frame_& f, Img4& out) {                                                                                              // pre_blur.hlsl
    const float tsx = 1.0f / (float)f.w, tsy = 1.0f / (float)f.h;
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) {
            int gx, gy; f.gcoord(x, y, gx, gy);
            const BilateralData center = tap_bilateral_data(f, gx, gy, ((float)gx + 0.5f) * tsx, ((float)gy + 0.5f) * tsy);
            if (center.z_01 > 0.999f) { out.at(x, y) = f4{0.0f, 0.0f, 0.0f, -1.0f}; continue; }
            const float center_dist = f.hit.load(x, y).w;
            float blur_radius = calc_blur_radius(center.roughness, 20.0f) * f.blur_radius;
            const float camera_dist = len3(center.position - camera_position(f.cam));
            blur_radius = blur_radius * hit_distance_attenuation(center.roughness, camera_dist, center_dist);
            float sum_weight;
            const f4 sum = screen_blur(f, x, y, center, blur_radius, f.rot_pre,
                                       [&](int tx, int ty) { f3 l = clamp_lighting(f.noised.load(tx, ty)); return f4{l.x, l.y, l.z, f.hit.load(tx, ty).w}; }, sum_weight);
            f4 blurred;
            if (sum_weight != 0.0f) blurred = over4(sum, sum_weight);
            else { f3 l = clamp_lighting(f.noised.load(x, y)); blurred = f4{l.x, l.y, l.z, center_dist}; }
            blurred.w = clampf(blurred.w, 0.0f, 16.0f);
            out.at(x, y) = half4(blurred);
        }
}

void temporal_accumulate(const Frame_& f, const State& hist, const Img4& in, Img4& out, Img1& accumulation) {                            // temporal_accumulate.hlsl
    const float tsx = 1.0f / (float)f.w, tsy = 1.0f / (float)f.h;
    const Img4 none4; const Img1 none1;
    const Img4& hist_ld = f.has_history ? hist.ld0 : none4;
    const Img1& hist_acc = f.has_history ? hist.accum : none1;
    const Img1& hist_depth = f.has_history ? hist.depth : none1;
    const Img4& hist_nr = f.has_history ? hist.normal_roughness : none4;
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) {
            const float cu = ((float)x + 0.5f) * tsx, cv = ((float)y + 0.5f) * tsy;
            int gx, gy; f.gcoord(x, y, gx, gy);
            const float depth = f.depth.load(gx, gy);
            const f4 lighting_dist = in.load(x, y);
            if (depth == 0.0f || f.mask(x, y) != 0u) { out.at(x, y) = lighting_dist; accumulation.at(x, y) = 0.0f; continue; }
            const f4 nr = f.normal_roughness.load(x, y);                          // Load(pixel_coord): :34
            const f3 normal = oct_decode(f2{nr.x, nr.y});
            const float roughness = nr.w;
            const f3 position = position_world(f.cam, cu, cv, depth);
            const f3 view = normalize(camera_position(f.cam) - position);
            const float ndotv = dot(normal, view);
            const f2 velocity = f.vel(gx, gy);
            const float pu = cu - velocity.x, pv = cv - velocity.y;
            const float hd = hist_depth.load(to_int(pu * (float)f.w), to_int(pv * (float)f.h));
            const f3 hist_position = position_world(f.hist_cam, pu, pv, hd);
            const f3 hist_view = normalize(camera_position(f.hist_cam) - hist_position);
            const float parallax = calc_parallax(view, hist_view);
            const f4 hist_lighting_dist = sample4(hist_ld, pu, pv);
            const float accum_hist = sample1(hist_acc, pu, pv);
            float accum_factor = get_specular_accum_speed(roughness, ndotv, parallax);
            accum_factor = fmin_(fmin_(accum_factor, accum_hist), 32.0f);
            f4 lerped = lerp4(hist_lighting_dist, lighting_dist, 1.0f / (1.0f + accum_factor));
            if (f.virtual_history != 0u) {
                const float dominant = get_specular_dominant_factor(ndotv, roughness);
                const f3 virtual_position = position - (view * lerped.w) * dominant;                   // utils.hlsl:93-96
                const f3 vc = project_uv(f.hist_cam, virtual_position);
                if (on_screen(vc.x, vc.y) != 0.0f && vc.z >= -1.0f && vc.z <= 1.0f) {
                    const int vx = to_int(vc.x * (float)f.w), vy = to_int(vc.y * (float)f.h);
                    const float hist_virtual_depth = hist_depth.load(vx, vy);
                    f4 hv = sample4(hist_ld, vc.x, vc.y);
                    float amount = get_specular_dominant_factor(ndotv, roughness);
                    const float confidence = 1.0f;
                    const float linear_depth = linear_01(f.cam, depth), virtual_linear_depth = linear_01(f.hist_cam, hist_virtual_depth);
                    amount = amount * (fabsf(linear_depth - virtual_linear_depth) < linear_depth * 0.1f ? 1.0f : 0.0f);
                    const f4 hnr = hist_nr.load(vx, vy);
                    amount = amount * (dot(normal, oct_decode(f2{hnr.x, hnr.y})) > 0.9f ? 1.0f : 0.0f);
                    float a_virtual = get_specular_accum_speed(roughness, ndotv, 0.0f);
                    const float a_min = fmin_(a_virtual, 4.0f * sqrtf(roughness));
                    float a = lerpf(1.0f / (1.0f + a_min), 1.0f / (1.0f + a_virtual), confidence);
                    a_virtual = 1.0f / a - 1.0f;
                    const float a_hit_dist = fmin_(a_virtual, 32.0f);
                    const float wc = 1.0f / (1.0f + a_virtual), wd = 1.0f / (1.0f + a_hit_dist);
                    hv = f4{lerpf(hv.x, lighting_dist.x, wc), lerpf(hv.y, lighting_dist.y, wc), lerpf(hv.z, lighting_dist.z, wc), lerpf(hv.w, lighting_dist.w, wd)};
                    const f4 result = lerp4(lerped, hv, amount);
                    a = lerpf(1.0f / (1.0f + accum_factor), 1.0f / (1.0f + a_virtual), amount);
                    accum_factor = 1.0f / a - 1.0f;
                    lerped = result;
                }
            }
            out.at(x, y) = half4(lerped);
            accumulation.at(x, y) = store_half(accum_factor);
        }
}

void fetch_linear_depth(const Frame_& f, Img1& out) {                                                                                   // fetch_linear_depth.hlsl
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) { int gx, gy; f.gcoord(x, y, gx, gy); out.at(x, y) = linear_01(f.cam, f.depth.load(gx, gy)); }
}

// gen_depth_mip.hlsl: per 16 x 16 group, 64 threads; level k + 1 is reduced from the group-shared (unrounded) level k
uint32_t extract_even_bits(uint32_t x) { x &= 0x55555555u; x = (x | (x >> 1)) & 0x33333333u; x = (x | (x >> 2)) & 0x0f0f0f0fu; x = (x | (x >> 4)) & 0x00ff00ffu; x = (x | (x >> 8)) & 0x0000ffffu; return x; }
void reduce4(const f4 v[4], const float d[4], f4& out_v, float& out_d) {
    const float dm = fmin_(fmin_(d[0], d[1]), fmin_(d[2], d[3]));
    float w[4];
    for (int k = 0; k < 4; k++) w[k] = (fabsf(d[k] - dm) < dm * 0.1f && v[k].w > 0.0f) ? 1.0f : 0.0f;
    const float ws = ((w[0] + w[1]) + w[2]) + w[3];
    out_d = dm;
    out_v = ws == 0.0f ? scale4(add4(add4(add4(v[0], v[1]), v[2]), v[3]), 0.25f)
                       : over4(add4(add4(add4(scale4(v[0], w[0]), scale4(v[1], w[1])), scale4(v[2], w[2])), scale4(v[3], w[3])), ws);
}
void gen_depth_mip(const Frame_& f, std::vector<Img4>& color, std::vector<Img1>& depth) {
    for (int l = 1; l < 4; l++) { color[l].resize(f.w >> l, f.h >> l); depth[l].resize(f.w >> l, f.h >> l); }
    auto store = [&](int l, int x, int y, f4 v, float d) { if (x < (int)(f.w >> l) && y < (int)(f.h >> l)) { color[l].at(x, y) = half4(v); depth[l].at(x, y) = d; } };
    for (uint32_t gy = 0; gy < (f.h + 15) / 16; gy++)
        for (uint32_t gx = 0; gx < (f.w + 15) / 16; gx++) {
            f4 s_data[64]; float s_depth[64]; int px[64], py[64];
            for (uint32_t li = 0; li < 64; li++) {
                px[li] = (int)gx * 16 + (int)extract_even_bits(li) * 2; py[li] = (int)gy * 16 + (int)extract_even_bits(li >> 1) * 2;
                f4 v[4]; float d[4];
                for (int k = 0; k < 4; k++) {
                    const int sx = clampi(px[li] + (k & 1), 0, (int)f.w - 1), sy = clampi(py[li] + (k >> 1), 0, (int)f.h - 1);          // safe_texel_fetch
                    v[k] = color[0].at(sx, sy); d[k] = depth[0].at(sx, sy);
                }
                reduce4(v, d, s_data[li], s_depth[li]);
                store(1, px[li] >> 1, py[li] >> 1, s_data[li], s_depth[li]);
            }
            for (uint32_t li = 0; li < 64; li += 4) {
                const f4 v[4] = {s_data[li], s_data[li + 1], s_data[li + 2], s_data[li + 3]};
                const float d[4] = {s_depth[li], s_depth[li + 1], s_depth[li + 2], s_depth[li + 3]};
                reduce4(v, d, s_data[li], s_depth[li]);
                store(2, px[li] >> 2, py[li] >> 2, s_data[li], s_depth[li]);
            }
            for (uint32_t li = 0; li < 64; li += 16) {
                const f4 v[4] = {s_data[li], s_data[li + 4], s_data[li + 8], s_data[li + 12]};
                const float d[4] = {s_depth[li], s_depth[li + 4], s_depth[li + 8], s_depth[li + 12]};
                f4 rv; float rd;
                reduce4(v, d, rv, rd);
                store(3, px[li] >> 3, py[li] >> 3, rv, rd);
            }
        }
}

void fix_history(const Frame_& f, const std::vector<Img4>& color, const std::vector<Img1>& depth, const Img1& accumulation, Img4& out) {    // fix_history.hlsl
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) {
            int gx, gy; f.gcoord(x, y, gx, gy);
            const float linear_depth = depth[0].load(x, y);
            if (linear_depth > 0.999f) continue;                                      // the target keeps what "Pre Blur" wrote
            const f4 lighting_dist = color[0].load(x, y);
            const float norm_accum = saturate(accumulation.load(x, y) / 4.0f);
            if (norm_accum == 1.0f) { out.at(x, y) = lighting_dist; continue; }
            const float roughness = f.normal_roughness.load(gx, gy).w;
            int mip_level = std::min(to_int((4.0f * (1.0f - norm_accum)) * roughness), 3);
            f4 sum{0, 0, 0, 0};
            float sum_weight = 0.0f;
            while (mip_level >= 0) {
                const int mx = x >> mip_level, my = y >> mip_level;
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        const float tap_linear_depth = depth[(size_t)mip_level].load(mx + dx, my + dy);
                        const f4 tap = color[(size_t)mip_level].load(mx + dx, my + dy);
                        if (fabsf(linear_depth - tap_linear_depth) < linear_depth * 0.1f && tap.w > 0.0f) {
                            const float r = sqrtf((float)(dx * dx + dy * dy));
                            const float w = exp_neg(-(r * r));                        // gaussian(length(dx, dy), 1)
                            sum = add4(sum, scale4(tap, w));
                            sum_weight = sum_weight + w;
                        }
                    }
                if (sum_weight > 3.5f) break;
                --mip_level;
            }
            out.at(x, y) = half4((sum_weight == 0.0f || mip_level < 0) ? lighting_dist : over4(sum, sum_weight));
        }
}

void blur(const Frame_& f, const Img4& in, const Img1& accumulation, Img4& out) {                                                        // blur.hlsl
    const float tsx = 1.0f / (float)f.w, tsy = 1.0f / (float)f.h;
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) {
            int gx, gy; f.gcoord(x, y, gx, gy);
            const BilateralData center = tap_bilateral_data(f, gx, gy, ((float)gx + 0.5f) * tsx, ((float)gy + 0.5f) * tsy);
            if (center.z_01 > 0.999f) { out.at(x, y) = f4{0.0f, 0.0f, 0.0f, -1.0f}; continue; }
            const f4 center_ld = in.load(x, y);
            const float accum = accumulation.load(x, y);
            const f3 view_vec = camera_position(f.cam) - center.position;
            const float camera_dist = len3(view_vec);
            const f3 view = view_vec / camera_dist;
            const f4 dominant = get_specular_dominant_direction(center.normal, view, center.roughness);
            const float blur_radius = (((calc_blur_radius(center.roughness, 0.04f) * f.blur_radius) * (1.0f - saturate(accum / 32.0f))) *
                                       hit_distance_attenuation(center.roughness, camera_dist, center_ld.w)) * saturate((camera_dist - 0.03f) / 0.05f);
            f3 Tv, Bv;
            get_kernel_basis(mk3(dominant.x, dominant.y, dominant.z), center.normal, center.roughness, Tv, Bv);
            Tv = Tv * blur_radius; Bv = Bv * blur_radius;
            f4 sum{0, 0, 0, 0};
            float sum_weight = 0.0f;
            for (int i = 0; i < 8; i++) {
                const f3 offset = poisson_disk_samples[i];
                const f2 r = rotate_vector(f.rot_blur, f2{offset.x, offset.y});
                const f3 position_i = (center.position + Tv * r.x) + Bv * r.y;
                const f3 c = project_uv(f.cam, position_i);
                const int tx = clampi(to_int(c.x * (float)f.w), 0, (int)f.w - 1), ty = clampi(to_int(c.y * (float)f.h), 0, (int)f.h - 1);
                int tgx, tgy; f.gcoord(tx, ty, tgx, tgy);
                const BilateralData tap = tap_bilateral_data(f, tgx, tgy, c.x, c.y);
                const float w = ((get_gaussian_weight(offset.z) * calc_bilateral_weight(center, tap)) * (tap.z_01 < 0.999f ? 1.0f : 0.0f)) * on_screen(c.x, c.y);
                sum = add4(sum, scale4(in.load(tx, ty), w));
                sum_weight = sum_weight + w;
            }
            out.at(x, y) = half4(sum_weight == 0.0f ? center_ld : over4(sum, sum_weight));
        }
}

f3 rgb_to_ycocg(f3 c) { return mk3((0.25f * c.x + 0.5f * c.y) + 0.25f * c.z, 0.5f * c.x - 0.5f * c.z, (-0.25f * c.x + 0.5f * c.y) - 0.25f * c.z); }      // color.hlsl:23-29
f3 ycocg_to_rgb(f3 c) { return mk3((c.x + c.y) - c.z, c.x + c.z, (c.x - c.y) - c.z); }                                                                    // color.hlsl:31-37
float hlsl_min(float a, float b) { return a != a ? b : (b != b ? a : (a < b ? a : b)); }
float hlsl_max(float a, float b) { return a != a ? b : (b != b ? a : (a > b ? a : b)); }
f4 clip_aabb_4d(f4 p_inside, f4 p, f4 p_min, f4 p_max) {                                                                                                   // math.hlsl:106-114
    const float in_[3] = {p_inside.x, p_inside.y, p_inside.z}, pp[3] = {p.x, p.y, p.z}, lo[3] = {p_min.x, p_min.y, p_min.z}, hi[3] = {p_max.x, p_max.y, p_max.z};
    float inter[3];
    for (int k = 0; k < 3; k++) {
        const float dir = pp[k] - in_[k];
        const float dir_inv = lerpf(1.0f / dir, 1.0f / 65536.0f, fabsf(dir) < 1.0f / 65536.0f ? 1.0f : 0.0f);
        inter[k] = hlsl_max((hi[k] - in_[k]) * dir_inv, (lo[k] - in_[k]) * dir_inv);
    }
    const float m = hlsl_min(inter[0], hlsl_min(inter[1], inter[2]));
    return lerp4(p_inside, p, m != m ? 0.0f : saturate(m));
}
void temporal_stabilize(const Frame_& f, const State& hist, const Img4& in, Img4& out) {                                                                  // temporal_stabilize.hlsl
    const float tsx = 1.0f / (float)f.w, tsy = 1.0f / (float)f.h;
    const Img4 none4;
    const Img4& hist_ld = f.has_history ? hist.ld1 : none4;
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) {
            int gx, gy; f.gcoord(x, y, gx, gy);
            f4 neighbors[9]; float nd[9], nl[9];
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {                                 // fill_shared_data: clamped to the image, colour as YCoCg
                    const int k = (dy + 1) * 3 + dx + 1;
                    const int sx = clampi(x + dx, 0, (int)f.w - 1), sy = clampi(y + dy, 0, (int)f.h - 1);
                    int sgx, sgy; f.gcoord(sx, sy, sgx, sgy);
                    const f4 v = in.load(sx, sy);
                    const f3 c = rgb_to_ycocg(mk3(v.x, v.y, v.z));
                    neighbors[k] = f4{c.x, c.y, c.z, v.w};
                    nd[k] = f.depth.load(sgx, sgy);
                    nl[k] = linear_01(f.cam, nd[k]);
                }
            if (nd[4] == 0.0f) { out.at(x, y) = f4{0.0f, 0.0f, 0.0f, -1.0f}; continue; }
            if (f.mask(x, y) != 0u) { const f3 c = ycocg_to_rgb(mk3(neighbors[4].x, neighbors[4].y, neighbors[4].z)); out.at(x, y) = half4(f4{c.x, c.y, c.z, neighbors[4].w}); continue; }
            const float roughness = f.normal_roughness.load(gx, gy).w;
            const float camera_dist = len3(position_view(f.cam, ((float)x + 0.5f) * tsx, ((float)y + 0.5f) * tsy, nd[4]));
            const f2 velocity = f.vel(gx, gy);
            const float pu = ((float)x + 0.5f) * tsx - velocity.x, pv = ((float)y + 0.5f) * tsy - velocity.y;
            f4 prev = sample4(hist_ld, pu, pv);
            { const f3 c = rgb_to_ycocg(mk3(prev.x, prev.y, prev.z)); prev = f4{c.x, c.y, c.z, prev.w}; }
            for (int k = 0; k < 9; k++)
                if (k != 4 && (nd[k] == 0.0f || fabsf(nl[k] - nl[4]) > nl[4] * 0.2f)) neighbors[k] = neighbors[4];
            const float vx = velocity.x * (float)f.w, vy = velocity.y * (float)f.h;
            const float velocity_length = sqrtf(vx * vx + vy * vy);
            const float anti_flickering_params = f.anti_flicker * hit_distance_attenuation(roughness, camera_dist, neighbors[4].w);
            f4 moment1{0, 0, 0, 0}, moment2{0, 0, 0, 0};
            for (int k = 0; k < 9; k++) {
                moment1 = add4(moment1, neighbors[k]);
                moment2 = add4(moment2, f4{neighbors[k].x * neighbors[k].x, neighbors[k].y * neighbors[k].y, neighbors[k].z * neighbors[k].z, neighbors[k].w * neighbors[k].w});
            }
            moment1 = over4(moment1, 9.0f); moment2 = over4(moment2, 9.0f);
            const f4 std_dev{sqrtf(fabsf(moment2.x - moment1.x * moment1.x)), sqrtf(fabsf(moment2.y - moment1.y * moment1.y)),
                             sqrtf(fabsf(moment2.z - moment1.z * moment1.z)), sqrtf(fabsf(moment2.w - moment1.w * moment1.w))};
            const float localized = lerpf(anti_flickering_params * 0.8f, anti_flickering_params * 2.25f, saturate(1.0f - 2.0f * velocity_length));
            float mult = 1.5f + localized;
            mult = lerpf(mult, 0.75f, saturate(velocity_length / 50.0f));
            const f4 lo{moment1.x - std_dev.x * mult, moment1.y - std_dev.y * mult, moment1.z - std_dev.z * mult, moment1.w - std_dev.w * mult};
            const f4 hi{moment1.x + std_dev.x * mult, moment1.y + std_dev.y * mult, moment1.z + std_dev.z * mult, moment1.w + std_dev.w * mult};
            prev = clip_aabb_4d(neighbors[4], prev, lo, hi);
            prev = lerp4(prev, neighbors[4], 0.05f);
            const f3 c = ycocg_to_rgb(mk3(prev.x, prev.y, prev.z));
            out.at(x, y) = half4(f4{c.x, c.y, c.z, prev.w});
        }
}

void post_blur(const Frame_& f, const Img4& in, const Img1& accumulation, Img4& out) {                                                   // post_blur.hlsl
    const float tsx = 1.0f / (float)f.w, tsy = 1.0f / (float)f.h;
    for (int y = 0; y < (int)f.h; y++)
        for (int x = 0; x < (int)f.w; x++) {
            int gx, gy; f.gcoord(x, y, gx, gy);
            const BilateralData center = tap_bilateral_data(f, gx, gy, ((float)gx + 0.5f) * tsx, ((float)gy + 0.5f) * tsy);
            if (center.z_01 > 0.999f) { out.at(x, y) = f4{0.0f, 0.0f, 0.0f, -1.0f}; continue; }
            const f4 center_ld = in.load(x, y);
            const float camera_dist = len3(center.position - camera_position(f.cam));
            const float blur_radius = ((calc_blur_radius(center.roughness, 15.0f) * f.blur_radius) * (1.0f - saturate(accumulation.load(x, y) / 32.0f))) *
                                      hit_distance_attenuation(center.roughness, camera_dist, center_ld.w);
            float sum_weight;
            const f4 sum = screen_blur(f, x, y, center, blur_radius, f.rot_post, [&](int tx, int ty) { return in.load(tx, ty); }, sum_weight);
            out.at(x, y) = half4(sum_weight == 0.0f ? center_ld : over4(sum, sum_weight));
        }
}

// reblur.cpp:174-197
const float pre_blur_rotator_rands[32] = {0.840188f, 0.394383f, 0.783099f, 0.79844f, 0.911647f, 0.197551f, 0.335223f, 0.76823f, 0.277775f, 0.55397f, 0.477397f, 0.628871f,
    0.364784f, 0.513401f, 0.95223f, 0.916195f, 0.635712f, 0.717297f, 0.141603f, 0.606969f, 0.0163006f, 0.242887f, 0.137232f, 0.804177f, 0.156679f, 0.400944f, 0.12979f, 0.108809f,
    0.998924f, 0.218257f, 0.512932f, 0.839112f};
const float blur_rotator_rands[32] = {0.61264f, 0.296032f, 0.637552f, 0.524287f, 0.493583f, 0.972775f, 0.292517f, 0.771358f, 0.526745f, 0.769914f, 0.400229f, 0.891529f,
    0.283315f, 0.352458f, 0.807725f, 0.919026f, 0.0697553f, 0.949327f, 0.525995f, 0.0860558f, 0.192214f, 0.663227f, 0.890233f, 0.348893f, 0.0641713f, 0.020023f, 0.457702f,
    0.0630958f, 0.23828f, 0.970634f, 0.902208f, 0.85092f};
const float post_blur_rotator_rands[32] = {0.266666f, 0.53976f, 0.375207f, 0.760249f, 0.512535f, 0.667724f, 0.531606f, 0.0392803f, 0.437638f, 0.931835f, 0.93081f, 0.720952f,
    0.284293f, 0.738534f, 0.639979f, 0.354049f, 0.687861f, 0.165974f, 0.440105f, 0.880075f, 0.829201f, 0.330337f, 0.228968f, 0.893372f, 0.35036f, 0.68667f, 0.956468f, 0.58864f,
    0.657304f, 0.858676f, 0.43956f, 0.92397f};
f4 get_rotator(float angle) { const float ca = std::cos(angle), sa = std::sin(angle); return f4{ca, sa, -sa, ca}; }

} // namespace

struct obpt_reblur_state { State s; };
static State& reblur_state(obpt_context* c) {
    if (!c->reblur) c->reblur = new obpt_reblur_state();
    return c->reblur->s;
}
void obpt_reblur_free(obpt_context* c) { delete c->reblur; c->reblur = nullptr; }

extern "C" {

bpt_status obpt_reblur_reset(obpt_context* c) { if (!c) return BPT_ERR_INVALID; reblur_state(c).valid = false; return BPT_OK; }

bpt_status obpt_denoise_reblur(obpt_context* c, const bpt_camera* cam, uint64_t frame_count, const bpt_reblur_settings* st, const bpt_reblur_inputs* in, float* out) {
    if (!c || !cam || !st || !in || !out) return BPT_ERR_INVALID;
    State& hist = reblur_state(c);
    Frame_ f;
    f.w = in->width; f.h = in->height; f.gw = c->width; f.gh = c->height;
    f.half = f.w != f.gw;                                                                  // reblur.cpp:280
    if (f.w < 8 || f.h < 8 || (f.half && (f.w != (f.gw + 1) / 2 || f.h != (f.gh + 1) / 2)) || (!f.half && f.h != f.gh)) { c->err = "reblur: bad extent"; return BPT_ERR_INVALID; }
    if (!in->noised || !in->hit_positions || !in->depth || !in->normal_roughness) { c->err = "reblur: null input"; return BPT_ERR_INVALID; }
    if (hist.w != f.w || hist.h != f.h || hist.gw != f.gw || hist.gh != f.gh) hist.valid = false;
    f.frame_index = (uint32_t)frame_count;
    f.has_history = hist.valid && hist.last_frame + 1 == frame_count;                      // reblur.cpp:282-285,357-361
    f.virtual_history = st->virtual_history; f.blur_radius = st->blur_radius; f.anti_flicker = st->anti_flickering_strength;
    f.cam = *cam; f.hist_cam = f.has_history ? hist.cam : *cam;
    const uint32_t ri = (uint32_t)(frame_count % 32);
    f.rot_pre = get_rotator(pre_blur_rotator_rands[ri]); f.rot_blur = get_rotator(blur_rotator_rands[ri]); f.rot_post = get_rotator(post_blur_rotator_rands[ri]);
    const size_t n = (size_t)f.w * f.h, gn = (size_t)f.gw * f.gh;
    f.depth.resize(f.gw, f.gh); std::memcpy(f.depth.px.data(), in->depth, gn * 4);
    f.normal_roughness.resize(f.gw, f.gh); std::memcpy(f.normal_roughness.px.data(), in->normal_roughness, gn * 16);
    f.noised.resize(f.w, f.h); std::memcpy(f.noised.px.data(), in->noised, n * 16);
    f.hit.resize(f.w, f.h); std::memcpy(f.hit.px.data(), in->hit_positions, n * 16);
    if (in->velocity) { f.velocity.resize(gn); std::memcpy(f.velocity.data(), in->velocity, gn * 8); }
    if (in->history_validation) f.validation.assign(in->history_validation, in->history_validation + n);

    std::vector<Img4> lighting_dist_0(4); std::vector<Img1> linear_depth(4);
    Img4 lighting_dist_1, denoised; Img1 accumulation;
    lighting_dist_0[0].resize(f.w, f.h); linear_depth[0].resize(f.w, f.h);
    lighting_dist_1.resize(f.w, f.h); denoised.resize(f.w, f.h); accumulation.resize(f.w, f.h);

    pre_blur(f, lighting_dist_1);                                                          // "ReBLUR Pre Blur"
    temporal_accumulate(f, hist, lighting_dist_1, lighting_dist_0[0], accumulation);       // "ReBLUR Temporal Accumulate"
    fetch_linear_depth(f, linear_depth[0]);                                                // "ReBLUR Fetch Linear Depth"
    gen_depth_mip(f, lighting_dist_0, linear_depth);                                       // "ReBLUR Gen Depth Mip"
    fix_history(f, lighting_dist_0, linear_depth, accumulation, lighting_dist_1);          // "ReBLUR Fix History"
    hist.dbg_ld0_mips = lighting_dist_0; hist.dbg_depth_mips = linear_depth;
    Img4 blurred; blurred.resize(f.w, f.h);
    blur(f, lighting_dist_1, accumulation, blurred);                                       // "ReBLUR Blur" (writes level 0 of lighting_dist_0)
    hist.dbg_ld0_mips[0] = blurred;
    temporal_stabilize(f, hist, blurred, lighting_dist_1);                                 // "ReBLUR Temporal Stabilize" (reads LAST frame's stabilised image)
    post_blur(f, lighting_dist_1, accumulation, denoised);                                 // "ReBLUR Post Blur"
    std::memcpy(out, denoised.px.data(), n * 16);

    hist.ld0 = std::move(blurred); hist.accum = accumulation; hist.ld1 = lighting_dist_1;  // reblur.cpp:512-513,554
    hist.dbg_ld1 = hist.ld1; hist.dbg_accum = hist.accum;
    hist.depth = std::move(f.depth); hist.normal_roughness = std::move(f.normal_roughness);
    hist.cam = *cam; hist.last_frame = frame_count; hist.valid = true; hist.w = f.w; hist.h = f.h; hist.gw = f.gw; hist.gh = f.gh;
    return BPT_OK;
}

// scalar functions of utils.hlsl / filter.hlsl for the float64 pins: {dominant factor, accumulation speed, magic curve (0.75), gaussian weight(x),
// hit-distance attenuation(roughness, camera distance = ndotv + 1e-3, hit distance = x)} with x = `parallax`
void obpt_unit_reblur_scalars(float ndotv, float roughness, float parallax, float out[6]) {
    out[0] = get_specular_dominant_factor(ndotv, roughness);
    out[1] = get_specular_accum_speed(roughness, ndotv, parallax);
    out[2] = get_specular_magic_curve2(roughness, 0.75f);
    out[3] = get_gaussian_weight(parallax);
    out[4] = hit_distance_attenuation(roughness, ndotv + 1e-3f, parallax);
    out[5] = 0.0f;
}

bpt_status obpt_debug_read_reblur(obpt_context* c, uint32_t which, float* out, uint64_t cap) {
    if (!c || !out) return BPT_ERR_INVALID;
    State& s = reblur_state(c);
    std::vector<float> flat;
    if (which == 0) for (auto& l : s.dbg_ld0_mips) for (auto& p : l.px) { flat.push_back(p.x); flat.push_back(p.y); flat.push_back(p.z); flat.push_back(p.w); }
    else if (which == 1) for (auto& p : s.dbg_ld1.px) { flat.push_back(p.x); flat.push_back(p.y); flat.push_back(p.z); flat.push_back(p.w); }
    else if (which == 2) flat = s.dbg_accum.px;
    else if (which == 3) for (auto& l : s.dbg_depth_mips) flat.insert(flat.end(), l.px.begin(), l.px.end());
    else { c->err = "reblur_debug_read: which must be 0..3"; return BPT_ERR_INVALID; }
    if (flat.empty() || cap < flat.size()) { c->err = "reblur_debug_read: nothing rendered yet or capacity too small"; return BPT_ERR_INVALID; }
    std::memcpy(out, flat.data(), flat.size() * 4);
    return BPT_OK;
}

} // extern "C"
