// oracle_math.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
//
// Scalar FP32 restatement of the reference's shader math for the wavefront path tracer.
// Every function cites the reference HLSL it follows (paths relative to
// /root/reference/bisemutum/shaders unless noted).
//
// Numeric contract shared with the CUDA kernels (DESIGN.md "Numerics"):
//   * IEEE-754 binary32, round-to-nearest, no contraction (compiled -ffp-contract=off),
//     fused multiply-add ONLY where `fmaf` is written explicitly.
//   * dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z ; normalize(v) = v * (1/sqrt(dot(v,v))).
//   * min/max are the ternary forms (a<b?a:b / a>b?a:b).
//   * sin/cos/acos/atan2 are the polynomial forms defined here (libm differs by ulps
//     between glibc and CUDA, which would break bit-parity through a chaotic process).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

static inline f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
static inline f3 splat3(float v) { return f3{v, v, v}; }
static inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline f3 operator/(f3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
static inline f3 operator/(f3 a, f3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
static inline float fmin_(float a, float b) { return a < b ? a : b; }
static inline float fmax_(float a, float b) { return a > b ? a : b; }
static inline f3 min3(f3 a, f3 b) { return {fmin_(a.x, b.x), fmin_(a.y, b.y), fmin_(a.z, b.z)}; }
static inline f3 max3(f3 a, f3 b) { return {fmax_(a.x, b.x), fmax_(a.y, b.y), fmax_(a.z, b.z)}; }
static inline float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline f3 cross(f3 a, f3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline f3 normalize(f3 v) {
    float inv = 1.0f / sqrtf(dot(v, v));
    return v * inv;
}
static inline float saturate(float x) { return fmin_(fmax_(x, 0.0f), 1.0f); }
static inline float clampf(float x, float lo, float hi) { return fmin_(fmax_(x, lo), hi); }
static inline float lerpf(float a, float b, float t) { return a + (b - a) * t; }
static inline f3 lerp3(f3 a, f3 b, float t) { return a + (b - a) * t; }
static inline f3 lerp3v(f3 a, f3 b, f3 t) { return a + (b - a) * t; }
// HLSL reflect(i, n) = i - 2 * n * dot(i, n)
static inline f3 reflect(f3 i, f3 n) { return i - n * (2.0f * dot(n, i)); }
static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline bool finite3(f3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

// core/utils/math.hlsl:3-8
static constexpr float PI = 3.14159265359f;
static constexpr float TWO_PI = 2.0f * 3.14159265359f;
static constexpr float INV_PI = 1.0f / 3.14159265359f;
static constexpr float HALF_PI = 1.57079632679f;

// core/utils/math.hlsl:12-23
static inline float pow2(float x) { return x * x; }
static inline float pow4(float x) { return pow2(pow2(x)); }
static inline float pow5(float x) { return pow4(x) * x; }

// ---- RNG: core/utils/random.hlsl:3-26 ---------------------------------------------------
static inline uint32_t rng_tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (uint32_t n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
static inline uint32_t rng_lcg(uint32_t& prev) {
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00ffffffu;
}
static inline float rng_next(uint32_t& prev) { return (float)rng_lcg(prev) / (float)0x01000000; }

// ---- trig with a fixed evaluation order (see header) ------------------------------------
// sincos of 2*pi*u for u in [0,1]: quadrant reduction + Taylor polynomials on [-pi/4, pi/4].
static inline void sincos_2pi(float u, float& s, float& c) {
    float x = u * 4.0f;
    float q = floorf(x + 0.5f);
    float r = x - q;
    float a = r * 1.57079637f;
    float a2 = a * a;
    float sp = -1.98412698e-4f + a2 * 2.75573192e-6f;
    sp = 8.33333333e-3f + a2 * sp;
    sp = -1.66666667e-1f + a2 * sp;
    float sr = a + (a * a2) * sp;
    float cp = 2.48015873e-5f + a2 * -2.75573192e-7f;
    cp = -1.38888889e-3f + a2 * cp;
    cp = 4.16666667e-2f + a2 * cp;
    cp = -0.5f + a2 * cp;
    float cr = 1.0f + a2 * cp;
    int qi = ((int)q) & 3;
    if (qi == 0) { s = sr; c = cr; }
    else if (qi == 1) { s = cr; c = -sr; }
    else if (qi == 2) { s = -sr; c = -cr; }
    else { s = -cr; c = sr; }
}
// atan on [0,1] (odd minimax polynomial, |err| < 1e-6), then octant fix-up → atan2.
static inline float atan_unit(float z) {
    float z2 = z * z;
    float p = 0.00282363896f;
    p = -0.0159569028f + z2 * p;
    p = 0.0425049886f + z2 * p;
    p = -0.0748900772f + z2 * p;
    p = 0.106347933f + z2 * p;
    p = -0.142027363f + z2 * p;
    p = 0.199926957f + z2 * p;
    p = -0.333331018f + z2 * p;
    return z + (z * z2) * p;
}
static inline float atan2_(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmax_(ax, ay), mn = fmin_(ax, ay);
    if (mx == 0.0f) return 0.0f;
    float a = atan_unit(mn / mx);
    if (ay > ax) a = HALF_PI - a;
    if (x < 0.0f) a = PI - a;
    return y < 0.0f ? -a : a;
}
// acos(x) = 2*atan2(sqrt(1-x), sqrt(1+x)) for x in [-1,1]
static inline float acos_(float x) {
    x = clampf(x, -1.0f, 1.0f);
    return 2.0f * atan2_(sqrtf(1.0f - x), sqrtf(1.0f + x));
}

// log2(x), x >= 1 finite, fixed operation order (lights.hlsl:445 needs a log2 that two compilers agree on bit for bit):
// x = m * 2^e, m in [sqrt(1/2), sqrt(2)); log2(m) = 2/ln2 * atanh(s) with s = (m - 1)/(m + 1), series through s^9.
static inline float log2_(float x) {
    uint32_t bits; std::memcpy(&bits, &x, 4);
    int e = (int)((bits >> 23) & 0xffu) - 127;
    uint32_t mb = (bits & 0x007fffffu) | 0x3f800000u;
    float m; std::memcpy(&m, &mb, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    const float s = (m - 1.0f) / (m + 1.0f), s2 = s * s;
    float p = 1.0f / 9.0f;
    p = p * s2 + 1.0f / 7.0f;
    p = p * s2 + 1.0f / 5.0f;
    p = p * s2 + 1.0f / 3.0f;
    p = p * s2 + 1.0f;
    return (float)e + (2.88539008f * s) * p;
}

// ---- frames: core/utils/frame.hlsl:9-34 --------------------------------------------------
struct Frame { f3 x, y, z; };
static inline Frame create_frame(f3 n) {             // frame.hlsl:9-18
    Frame f; f.z = n;
    float sign = n.z > 0.0f ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    f.x = mk3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    f.y = mk3(b, sign + n.y * n.y * a, -n.y);
    return f;
}
static inline Frame create_frame(f3 n, f3 t) {       // frame.hlsl:20-26
    Frame f; f.z = n;
    f.y = normalize(cross(n, t));
    f.x = cross(f.y, n);
    return f;
}
static inline f3 frame_to_local(const Frame& f, f3 v) { return mk3(dot(v, f.x), dot(v, f.y), dot(v, f.z)); }
static inline f3 frame_to_world(const Frame& f, f3 v) { return v.x * f.x + v.y * f.y + v.z * f.z; }

// The reference stores (N, T) in the G-buffer with pack_normal_and_tangent and reads them back
// with unpack_normal_and_tangent (core/utils/pack.hlsl:112-129). In fp32 state mode the
// quantisation is dropped but the *re-orthogonalisation of T against N* that the round trip
// performs is kept: T' = normalize(frame(N).x * px + frame(N).y * py), (px,py) the L1-normalised
// projection of T. N itself passes through unchanged.
static inline f3 gbuffer_roundtrip_tangent(f3 N, f3 T) {
    Frame fr = create_frame(N);
    float px = dot(T, fr.x), py = dot(T, fr.y);
    float lnorm = fabsf(px) + fabsf(py);
    px = px / lnorm; py = py / lnorm;
    float packed_z = px * 0.5f + 0.5f;                   // pack.hlsl:118
    float sign = py < 0.0f ? -1.0f : 1.0f;               // pack.hlsl:119,124 (sign carried by packed.z)
    float projected_x = packed_z * 2.0f - 1.0f;          // pack.hlsl:125 (sign*sign*packed_z)
    float projected_y = sign * (1.0f - fabsf(projected_x));
    return normalize(fr.x * projected_x + fr.y * projected_y);
}

// ---- texture formats of state_precision = reference_fp16 -------------------------------------
// The reference's wavefront state lives in rgba16_sfloat / rgba16_unorm / rgba8_unorm textures
// (src/renderer/pass/path_tracing.cpp:248-288, src/renderer/pass/gbuffer.hpp:14-17). A store is modelled as the
// round-to-nearest-even conversion of the FP32 value; these return what a later load reads.
static inline float store_half(float f) {
    if (!std::isfinite(f) || f == 0.0f) return f;
    float a = fabsf(f);
    if (a >= 65520.0f) return copysignf(INFINITY, f);           // half max 65504, next step would be 65536
    int e;
    (void)frexpf(a, &e);                                        // a = m * 2^e, m in [0.5, 1)
    int ulp_exp = (e - 1 < -14 ? -14 : e - 1) - 10;             // spacing of halves around a (denormals below 2^-14)
    float r = ldexpf(nearbyintf(ldexpf(a, -ulp_exp)), ulp_exp); // exact scaling; nearbyint = ties-to-even (default mode)
    return copysignf(r, f);
}
static inline f3 store_half3(f3 v) { return mk3(store_half(v.x), store_half(v.y), store_half(v.z)); }
static inline float store_unorm(float f, int bits) {            // NaN and negatives -> 0, >1 -> 1
    float scale = (float)((1u << bits) - 1u);
    float c = f > 0.0f ? (f < 1.0f ? f : 1.0f) : 0.0f;
    return nearbyintf(c * scale) / scale;
}
static inline uint32_t to_uint(float f) {                       // HLSL uint(x): truncate; saturating outside [0, 2^32)
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}
// core/utils/pack.hlsl:6-23
static inline f2 pack_u32_to_unorm16x2(uint32_t x) { return f2{(float)(x & 0xffffu) / 65536.0f, (float)(x >> 16) / 65536.0f}; }
static inline uint32_t unpack_u32_from_unorm16x2(f2 x) { return to_uint(x.x * 65535.5f) | (to_uint(x.y * 65535.5f) << 16); }
static inline uint32_t pack_color_rg11b10(f3 c) { return to_uint(c.x * 2047.0f) | (to_uint(c.y * 2047.0f) << 11) | (to_uint(c.z * 1023.0f) << 22); }
static inline f3 unpack_color_rg11b10(uint32_t p) {
    return mk3((float)(p & 0x7ffu) / 2047.0f, (float)((p >> 11) & 0x7ffu) / 2047.0f, (float)((p >> 22) & 0x3ffu) / 1023.0f);
}
// core/utils/pack.hlsl:85-104
static inline f2 oct_wrap(f2 v) { return f2{(1.0f - fabsf(v.y)) * (v.x >= 0.0f ? 1.0f : -1.0f), (1.0f - fabsf(v.x)) * (v.y >= 0.0f ? 1.0f : -1.0f)}; }
static inline f2 oct_encode(f3 n) {
    n = n / ((fabsf(n.x) + fabsf(n.y)) + fabsf(n.z));
    return n.z >= 0.0f ? f2{n.x, n.y} : oct_wrap(f2{n.x, n.y});
}
static inline f3 oct_decode(f2 f) {
    f3 n = mk3(f.x, f.y, (1.0f - fabsf(f.x)) - fabsf(f.y));
    float t = clampf(-n.z, 0.0f, 1.0f);
    float ax = n.x >= 0.0f ? -t : t, ay = n.y >= 0.0f ? -t : t;
    n.x = n.x + ax; n.y = n.y + ay;
    return normalize(n);
}
// core/utils/pack.hlsl:112-129
static inline f3 pack_normal_and_tangent(f3 N, f3 T) {
    f2 oct = oct_encode(N);
    Frame fr = create_frame(N);
    float px = dot(T, fr.x), py = dot(T, fr.y);
    float lnorm = fabsf(px) + fabsf(py);
    px = px / lnorm; py = py / lnorm;
    float projected_x = px * 0.5f + 0.5f;
    return mk3(oct.x, oct.y, py < 0.0f ? -projected_x : projected_x);
}
static inline void unpack_normal_and_tangent(f3 packed, f3& N, f3& T) {
    N = oct_decode(f2{packed.x, packed.y});
    float sign = packed.z < 0.0f ? -1.0f : 1.0f;
    float projected_x = (sign * packed.z) * 2.0f - 1.0f;
    float projected_y = sign * (1.0f - fabsf(projected_x));
    Frame fr = create_frame(N);
    T = normalize(fr.x * projected_x + fr.y * projected_y);
}

// ---- surface + BSDF: core/material/utils.hlsl, core/material/lit.hlsl -------------------
struct SurfaceData {                                  // material/utils.hlsl:5-16
    f3 emission, base_color, f0_color, f90_color, normal_map_value;
    float roughness, anisotropy, ior, opacity;
    bool two_sided;
};
static inline SurfaceData surface_data_default() {   // material/utils.hlsl:18-31
    SurfaceData s;
    s.emission = splat3(0.0f); s.base_color = splat3(0.5f); s.f0_color = splat3(0.04f);
    s.f90_color = splat3(1.0f); s.normal_map_value = mk3(0.5f, 0.5f, 1.0f);
    s.roughness = 0.5f; s.anisotropy = 0.0f; s.ior = 1.5f; s.opacity = 1.0f; s.two_sided = false;
    return s;
}
static inline SurfaceData surface_data_diffuse(f3 base_color) {   // material/utils.hlsl:32-36
    SurfaceData s = surface_data_default();
    s.base_color = base_color;
    return s;
}
// renderer/gbuffer.hlsl:18-45: pack_surface_to_gbuffer -> the four G-buffer textures (rgba16_sfloat base colour,
// rgba16_sfloat normal+roughness, rgba16_unorm Fresnel, rgba8_unorm material_0) -> unpack_gbuffer_to_surface.
struct GBuffer { f4 base_color, normal_roughness, fresnel, material_0; };
static inline GBuffer pack_surface_to_gbuffer(f3 N, f3 T, const SurfaceData& s, uint32_t surface_model) {
    GBuffer g;
    g.base_color = f4{s.base_color.x, s.base_color.y, s.base_color.z, 1.0f};
    f2 a = pack_u32_to_unorm16x2(pack_color_rg11b10(s.f0_color)), b = pack_u32_to_unorm16x2(pack_color_rg11b10(s.f90_color));
    g.fresnel = f4{a.x, a.y, b.x, b.y};
    f3 pf = pack_normal_and_tangent(N, T);
    g.normal_roughness = f4{pf.x, pf.y, pf.z, s.roughness};
    g.material_0 = f4{s.anisotropy, 1.0f / s.ior, 0.0f, (float)surface_model / 256.0f};
    return g;
}
static inline GBuffer store_gbuffer(const GBuffer& g) {           // the texture formats of pass/gbuffer.hpp:14-17
    GBuffer o;
    o.base_color = f4{store_half(g.base_color.x), store_half(g.base_color.y), store_half(g.base_color.z), store_half(g.base_color.w)};
    o.normal_roughness = f4{store_half(g.normal_roughness.x), store_half(g.normal_roughness.y), store_half(g.normal_roughness.z), store_half(g.normal_roughness.w)};
    o.fresnel = f4{store_unorm(g.fresnel.x, 16), store_unorm(g.fresnel.y, 16), store_unorm(g.fresnel.z, 16), store_unorm(g.fresnel.w, 16)};
    o.material_0 = f4{store_unorm(g.material_0.x, 8), store_unorm(g.material_0.y, 8), store_unorm(g.material_0.z, 8), store_unorm(g.material_0.w, 8)};
    return o;
}
static inline void unpack_gbuffer_to_surface(const GBuffer& g, f3& N, f3& T, SurfaceData& s, uint32_t& surface_model) {
    surface_model = to_uint(g.material_0.w * 255.5f);
    s.base_color = mk3(g.base_color.x, g.base_color.y, g.base_color.z);
    s.f0_color = unpack_color_rg11b10(unpack_u32_from_unorm16x2(f2{g.fresnel.x, g.fresnel.y}));
    s.f90_color = unpack_color_rg11b10(unpack_u32_from_unorm16x2(f2{g.fresnel.z, g.fresnel.w}));
    s.roughness = g.normal_roughness.w;
    unpack_normal_and_tangent(mk3(g.normal_roughness.x, g.normal_roughness.y, g.normal_roughness.z), N, T);
    s.anisotropy = g.material_0.x;
    s.ior = 1.0f / g.material_0.y;
    s.opacity = 1.0f;
}
static inline f3 schlick_mix(f3 f0, f3 f90, float cos_theta) {    // utils.hlsl:40-42
    return lerp3(f0, f90, pow5(1.0f - cos_theta));
}
static inline f3 schlick_fresnel(f3 f0, f3 f90, float cos_theta, float ior) { // utils.hlsl:44-51
    if (cos_theta < 0.0f) {
        float eta = 1.0f / ior;
        float sin_theta_sqr = eta * eta * (1.0f - cos_theta * cos_theta);
        cos_theta = sqrtf(fmax_(1.0f - sin_theta_sqr, 0.0f));
    }
    return schlick_mix(f0, f90, cos_theta);
}
static inline void get_anisotropic_roughness(float roughness, float anisotropy, float& rx, float& ry) { // utils.hlsl:54-59
    float aniso = sqrtf(1.0f - anisotropy * 0.9f);
    float r2 = roughness * roughness;
    rx = fmax_(r2 / aniso, 0.001f);
    ry = fmax_(r2 * aniso, 0.001f);
}
static inline float ggx_ndf(f3 h, float rx, float ry) {           // utils.hlsl:62-65
    float a = (pow2(h.x / rx) + pow2(h.y / ry)) + pow2(h.z);
    return INV_PI / (rx * ry * a * a);
}
static inline float ggx_g1(f3 v, float rx, float ry) {            // utils.hlsl:67-70
    float a = (pow2(rx * v.x) + pow2(ry * v.y)) / fmax_(v.z * v.z, 0.0001f);
    return 2.0f / (1.0f + sqrtf(1.0f + a));
}
static inline float ggx_visible_hc(f3 v, f3 l, float rx, float ry) { // utils.hlsl:92-96
    float vv = l.z * sqrtf((pow2(rx * v.x) + pow2(ry * v.y)) + pow2(v.z));
    float ll = v.z * sqrtf((pow2(rx * l.x) + pow2(ry * l.y)) + pow2(l.z));
    return 0.5f / fmax_(vv + ll, 0.0001f);
}
static inline f3 ggx_vndf_sample(f3 v, float rx, float ry, float rand_x, float rand_y) { // utils.hlsl:98-116
    if (v.z < 0.0f) v = -v;
    f3 vh = normalize(mk3(rx * v.x, ry * v.y, v.z));
    float len_sqr = vh.x * vh.x + vh.y * vh.y;
    f3 t1v = len_sqr > 0.0f ? mk3(-vh.y, vh.x, 0.0f) / sqrtf(len_sqr) : mk3(1.0f, 0.0f, 0.0f);
    f3 t2v = cross(vh, t1v);
    float r = sqrtf(rand_x);
    float sn, cs;
    sincos_2pi(rand_y, sn, cs);                         // phi = TWO_PI * rand.y
    float t1 = r * cs;
    float t2 = r * sn;
    float s = 0.5f * (1.0f + vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    f3 nh = (t1 * t1v + t2 * t2v) + sqrtf(fmax_((1.0f - t1 * t1) - t2 * t2, 0.0f)) * vh;
    return normalize(mk3(rx * nh.x, ry * nh.y, fmax_(nh.z, 0.0f)));
}
static inline float ggx_vndf_sample_pdf(f3 h, f3 v, float rx, float ry) { // utils.hlsl:118-121
    return ggx_g1(v, rx, ry) * ggx_ndf(h, rx, ry) * fmax_(dot(h, v), 0.0f) / fmax_(v.z, 0.0001f);
}

// core/material/lit.hlsl:5-35 (+ unlit → 0, material.hlsl:92-103). Returns diffuse + specular.
static inline f3 surface_eval(f3 N, f3 T, f3 B, f3 V, f3 L, const SurfaceData& s, uint32_t surface_model) {
    if (surface_model != 1u) return splat3(0.0f);
    f3 H = normalize(V + L);
    f3 lh = mk3(dot(H, T), dot(H, B), dot(H, N));
    f3 lv = mk3(dot(V, T), dot(V, B), dot(V, N));
    f3 ll = mk3(dot(L, T), dot(L, B), dot(L, N));
    if (lv.z <= 0.0f || ll.z <= 0.0f) return splat3(0.0f);
    f3 fr = schlick_fresnel(s.f0_color, s.f90_color, fmax_(dot(V, H), 0.0f), s.ior);
    f3 diffuse = (splat3(1.0f) - fr) * s.base_color * INV_PI * fmax_(ll.z, 0.0f);
    float rx, ry;
    get_anisotropic_roughness(s.roughness, s.anisotropy, rx, ry);
    float ndf = ggx_ndf(lh, rx, ry);
    float vis = ggx_visible_hc(lv, ll, rx, ry);
    f3 specular = fr * ndf * vis * fmax_(ll.z, 0.0f);
    return diffuse + specular;
}
// exp(x) for x <= 0 in fixed order (shared definition with csrc/bpt_math.cuh: exp_neg): ln 2 range reduction + degree-6 polynomial
static inline float exp_neg(float x) {
    if (x < -87.0f) return 0.0f;
    float n = floorf(x * 1.44269504f + 0.5f);
    float r = (x - n * 0.693145752f) - n * 1.42860677e-6f;
    float p = 0.00833333333f + r * 0.00138888889f;
    p = 0.0416666667f + r * p;
    p = 0.166666667f + r * p;
    p = 0.5f + r * p;
    p = 1.0f + r * p;
    p = 1.0f + r * p;
    return p * u2f((uint32_t)((int)n + 127) << 23);
}
// the specular term alone (the `bsdf_specular` out-parameter of surface_eval, material.hlsl:81-118 / lit.hlsl:5-35)
static inline f3 surface_eval_specular(f3 N, f3 T, f3 B, f3 V, f3 L, const SurfaceData& s, uint32_t surface_model) {
    if (surface_model != 1u) return splat3(0.0f);
    f3 H = normalize(V + L);
    f3 lh = mk3(dot(H, T), dot(H, B), dot(H, N));
    f3 lv = mk3(dot(V, T), dot(V, B), dot(V, N));
    f3 ll = mk3(dot(L, T), dot(L, B), dot(L, N));
    if (lv.z <= 0.0f || ll.z <= 0.0f) return splat3(0.0f);
    f3 fr = schlick_fresnel(s.f0_color, s.f90_color, fmax_(dot(V, H), 0.0f), s.ior);
    float rx, ry;
    get_anisotropic_roughness(s.roughness, s.anisotropy, rx, ry);
    float ndf = ggx_ndf(lh, rx, ry);
    float vis = ggx_visible_hc(lv, ll, rx, ry);
    return fr * ndf * vis * fmax_(ll.z, 0.0f);
}
// core/material/lit.hlsl:37-58
static inline f3 surface_eval_lut(f3 N, f3 V, const SurfaceData& s, f3 int_diffuse, f3 int_specular, f2 int_brdf, uint32_t surface_model) {
    if (surface_model != 1u) return splat3(0.0f);
    float ndotv = dot(N, V);
    if (ndotv <= 0.0f) return splat3(0.0f);
    f3 fr = schlick_fresnel(s.f0_color, s.f90_color, ndotv, s.ior);
    f3 diffuse = (splat3(1.0f) - fr) * s.base_color * INV_PI;
    f3 specular = s.f0_color * int_brdf.x + s.f90_color * int_brdf.y;
    return diffuse * int_diffuse + specular * int_specular;
}

// ---- sampling: core/utils/sampling.hlsl:14-19 -------------------------------------------
// core/utils/sampling.hlsl:24-28
static inline f3 cos_hemisphere_sample(float rand_x, float rand_y) {
    float sn, cs;
    sincos_2pi(rand_x, sn, cs);                         // phi = rand.x * TWO_PI
    float r = sqrtf(rand_y);
    return mk3(cs * r, sn * r, sqrtf(fmax_(1.0f - rand_y, 0.0f)));
}
static inline f3 uniform_sphere_sample(float rand_x, float rand_y) {
    float sn, cs;
    sincos_2pi(rand_x, sn, cs);
    float z = rand_y * 2.0f - 1.0f;
    float r = sqrtf(fmax_(1.0f - z * z, 0.0f));
    return mk3(cs * r, sn * r, z);
}

} // namespace orc
