// bpt_math.cuh — device math of the wavefront path tracer (sm_100a).
//
// FP32 shading math that the kernels share: RNG, frames, GGX/VNDF, Schlick, the lit BSDF.
// Reference formulas (paths under bisemutum/shaders/):
//   core/utils/random.hlsl:3-26, core/utils/frame.hlsl:9-34, core/utils/pack.hlsl:112-129,
//   core/material/utils.hlsl:18-129, core/material/lit.hlsl:5-58, core/utils/sampling.hlsl:14-19
//
// Numeric contract (DESIGN.md "Numerics"): compiled with -fmad=false so a*b+c is never fused
// implicitly; fmaf() appears only where the contract calls for a fused op (ray/box slabs).
// dot(a,b) = (ax*bx + ay*by) + az*bz; normalize(v) = v * (1/sqrt(dot(v,v))); trigonometry is the
// fixed-order polynomial below, not libdevice, so results do not depend on the math library.
//
// Everything is BPT_HD so the same source can be compiled for the host by tests/hostcheck
// (a test-only harness; libbpt.so itself contains no host execution path).
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define BPT_HD __host__ __device__ __forceinline__
#else
#define BPT_HD inline
#endif

namespace bptd {

#if defined(__CUDA_ARCH__)
BPT_HD uint32_t f2u(float f) { return __float_as_uint(f); }
BPT_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
BPT_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
BPT_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

BPT_HD float3 v3(float x, float y, float z) { return make_float3(x, y, z); }
BPT_HD float3 v3s(float s) { return make_float3(s, s, s); }
BPT_HD float3 operator+(float3 a, float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
BPT_HD float3 operator-(float3 a, float3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
BPT_HD float3 operator*(float3 a, float3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
BPT_HD float3 operator*(float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
BPT_HD float3 operator*(float s, float3 a) { return v3(s * a.x, s * a.y, s * a.z); }
BPT_HD float3 operator/(float3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
BPT_HD float3 operator/(float3 a, float3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
BPT_HD float3 operator-(float3 a) { return v3(-a.x, -a.y, -a.z); }
// ternary min/max: identical semantics on host and device (no NaN/-0 library differences)
BPT_HD float tmin_(float a, float b) { return a < b ? a : b; }
BPT_HD float tmax_(float a, float b) { return a > b ? a : b; }
BPT_HD float3 vmin(float3 a, float3 b) { return v3(tmin_(a.x, b.x), tmin_(a.y, b.y), tmin_(a.z, b.z)); }
BPT_HD float3 vmax(float3 a, float3 b) { return v3(tmax_(a.x, b.x), tmax_(a.y, b.y), tmax_(a.z, b.z)); }
BPT_HD float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
BPT_HD float3 cross3(float3 a, float3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
BPT_HD float3 normalize3(float3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return v * inv; }
BPT_HD float sat(float x) { return tmin_(tmax_(x, 0.0f), 1.0f); }
BPT_HD float clampf_(float x, float lo, float hi) { return tmin_(tmax_(x, lo), hi); }
BPT_HD float mix1(float a, float b, float t) { return a + (b - a) * t; }
BPT_HD float3 mix3(float3 a, float3 b, float t) { return a + (b - a) * t; }
BPT_HD float3 reflect3(float3 i, float3 n) { return i - n * (2.0f * dot3(n, i)); }   // HLSL reflect
BPT_HD float max3c(float3 c) { return tmax_(c.x, tmax_(c.y, c.z)); }
BPT_HD bool is_finite1(float x) { return (f2u(x) & 0x7f800000u) != 0x7f800000u; }
BPT_HD bool is_finite3(float3 v) { return is_finite1(v.x) && is_finite1(v.y) && is_finite1(v.z); }

constexpr float kPi = 3.14159265359f;           // core/utils/math.hlsl:3
constexpr float kInvPi = 1.0f / 3.14159265359f;
constexpr float kHalfPi = 1.57079632679f;

BPT_HD float sq(float x) { return x * x; }
BPT_HD float pow5f(float x) { float x2 = x * x; return (x2 * x2) * x; }     // math.hlsl:12-23

// ---- RNG (random.hlsl:3-26) ---------------------------------------------------------------
BPT_HD uint32_t rng_tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 16; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
BPT_HD float rng_next(uint32_t& state) {
    state = 1664525u * state + 1013904223u;
    return (float)(state & 0x00ffffffu) / 16777216.0f;
}

// ---- fixed-order trigonometry ---------------------------------------------------------------
BPT_HD void sincos_2pi(float u, float& s, float& c) {
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
    s = qi == 0 ? sr : (qi == 1 ? cr : (qi == 2 ? -sr : -cr));
    c = qi == 0 ? cr : (qi == 1 ? -sr : (qi == 2 ? -cr : sr));
}
// exp(x) for x <= 0 in fixed order (range reduction by ln 2, degree-6 polynomial): ~2e-7 relative; 0 below -87
BPT_HD float exp_neg(float x) {
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
BPT_HD float atan_unit(float z) {
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
BPT_HD float atan2_(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = tmax_(ax, ay), mn = tmin_(ax, ay);
    if (mx == 0.0f) return 0.0f;
    float a = atan_unit(mn / mx);
    if (ay > ax) a = kHalfPi - a;
    if (x < 0.0f) a = kPi - a;
    return y < 0.0f ? -a : a;
}
// log2 of a finite x >= 1 in fixed-order FP32 arithmetic (libm and libdevice differ by ulps; lights.hlsl:445 takes log2 of a texel
// count): x = m * 2^e with m in [sqrt(1/2), sqrt(2)), log2(m) = (2 / ln 2) * atanh(s), s = (m - 1) / (m + 1), |s| < 0.1716, odd series to s^9
// (truncation < 4e-10); max error 2e-7 against libm over [1, 2^24].
BPT_HD float log2_(float x) {
    uint32_t b = f2u(x);
    int e = (int)((b >> 23) & 0xffu) - 127;
    float m = u2f((b & 0x007fffffu) | 0x3f800000u);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    float s = (m - 1.0f) / (m + 1.0f);
    float s2 = s * s;
    float p = 1.0f / 9.0f;
    p = p * s2 + 1.0f / 7.0f;
    p = p * s2 + 1.0f / 5.0f;
    p = p * s2 + 1.0f / 3.0f;
    p = p * s2 + 1.0f;
    return (float)e + (2.88539008f * s) * p;
}
BPT_HD float acos_(float x) {
    x = clampf_(x, -1.0f, 1.0f);
    return 2.0f * atan2_(sqrtf(1.0f - x), sqrtf(1.0f + x));
}

// ---- frames (frame.hlsl) -----------------------------------------------------------------
struct Frame3 { float3 x, y, z; };
BPT_HD Frame3 frame_from_normal(float3 n) {                       // frame.hlsl:9-18
    Frame3 f; f.z = n;
    float sign = n.z > 0.0f ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    f.x = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    f.y = v3(b, sign + n.y * n.y * a, -n.y);
    return f;
}
BPT_HD Frame3 frame_from_nt(float3 n, float3 t) {                 // frame.hlsl:20-26
    Frame3 f; f.z = n;
    f.y = normalize3(cross3(n, t));
    f.x = cross3(f.y, n);
    return f;
}
BPT_HD float3 to_local(const Frame3& f, float3 v) { return v3(dot3(v, f.x), dot3(v, f.y), dot3(v, f.z)); }
BPT_HD float3 to_world(const Frame3& f, float3 v) { return v.x * f.x + v.y * f.y + v.z * f.z; }

// T after the G-buffer pack/unpack round trip (pack.hlsl:112-129) without the storage quantisation.
BPT_HD float3 tangent_after_gbuffer(float3 N, float3 T) {
    Frame3 fr = frame_from_normal(N);
    float px = dot3(T, fr.x), py = dot3(T, fr.y);
    float lnorm = fabsf(px) + fabsf(py);
    px = px / lnorm; py = py / lnorm;
    float packed_z = px * 0.5f + 0.5f;
    float sign = py < 0.0f ? -1.0f : 1.0f;
    float projected_x = packed_z * 2.0f - 1.0f;
    float projected_y = sign * (1.0f - fabsf(projected_x));
    return normalize3(fr.x * projected_x + fr.y * projected_y);
}

// ---- storage formats of state_precision = reference_fp16 (SURVEY §8a storage quantisation note) ---------
// The reference keeps its wavefront state in textures (path_tracing.cpp:248-288, pass/gbuffer.hpp:14-17):
// rgba16_sfloat (directions, weights, colours, base colour, normal/roughness), rgba16_unorm (Fresnel) and
// rgba8_unorm (material_0). A store to those formats is modelled as round-to-nearest-even of the FP32 value;
// q_*() return the value a later load sees.
BPT_HD float q_half(float f) {
#if defined(__CUDA_ARCH__)
    return __half2float(__float2half_rn(f));           // IEEE RNE, identical to the host form below
#else
    uint32_t x = f2u(f), sign = x & 0x80000000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return f;                     // inf / NaN pass through
    if (x >= 0x477ff000u) return u2f(sign | 0x7f800000u);   // >= 65520 rounds to inf
    if (x <= 0x33000000u) return u2f(sign);             // <= 2^-25 rounds to (signed) zero
    uint32_t h;
    if (x < 0x38800000u) {                              // below 2^-14: half denormal, unit 2^-24
        uint32_t e = x >> 23, m = (x & 0x7fffffu) | 0x800000u, shift = 126u - e;
        h = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1u);
        if (rem > halfway || (rem == halfway && (h & 1u))) h++;
        return u2f(sign | f2u((float)h * 5.9604644775390625e-8f));
    }
    h = (x - 0x38000000u) >> 13;
    uint32_t rem = x & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return u2f(sign | ((h << 13) + 0x38000000u));      // (h = 0x7c00 cannot happen: x < 65520)
#endif
}
BPT_HD float3 q_half3(float3 v) { return v3(q_half(v.x), q_half(v.y), q_half(v.z)); }
BPT_HD float q_unorm(float f, float scale) {            // float -> unorm -> float; NaN -> 0
    float c = f > 0.0f ? (f < 1.0f ? f : 1.0f) : 0.0f;
    return rintf(c * scale) / scale;
}
BPT_HD uint32_t ftou(float f) {                         // HLSL uint(x): truncation; out of range saturates (as cvt.rzi.u32.f32)
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}
BPT_HD uint32_t pack_color_rg11b10(float3 c) {          // pack.hlsl:14-16
    return ftou(c.x * 2047.0f) | (ftou(c.y * 2047.0f) << 11) | (ftou(c.z * 1023.0f) << 22);
}
BPT_HD float3 unpack_color_rg11b10(uint32_t p) {        // pack.hlsl:17-23
    return v3((float)(p & 0x7ffu) / 2047.0f, (float)((p >> 11) & 0x7ffu) / 2047.0f, (float)((p >> 22) & 0x3ffu) / 1023.0f);
}
// Fresnel colour into gbuffer.fresnel: rg11b10 -> two unorm16 channels (value k/65536, pack.hlsl:6-8) -> rgba16_unorm texel;
// and back: uint(x * 65535.5) (pack.hlsl:9-12) -> rg11b10. (For k > 32768 the texel rounds to k-1: a reference quirk.)
BPT_HD float2 fresnel_to_gbuffer(float3 c) {
    uint32_t p = pack_color_rg11b10(c);
    return make_float2(q_unorm((float)(p & 0xffffu) / 65536.0f, 65535.0f), q_unorm((float)(p >> 16) / 65536.0f, 65535.0f));
}
BPT_HD float3 fresnel_from_gbuffer(float2 t) { return unpack_color_rg11b10(ftou(t.x * 65535.5f) | (ftou(t.y * 65535.5f) << 16)); }
BPT_HD float3 fresnel_through_gbuffer(float3 c) { return fresnel_from_gbuffer(fresnel_to_gbuffer(c)); }
BPT_HD float2 oct_encode(float3 n) {                    // pack.hlsl:88-95
    float l1 = (fabsf(n.x) + fabsf(n.y)) + fabsf(n.z);
    n = n / l1;
    if (n.z >= 0.0f) return make_float2(n.x, n.y);
    return make_float2((1.0f - fabsf(n.y)) * (n.x >= 0.0f ? 1.0f : -1.0f), (1.0f - fabsf(n.x)) * (n.y >= 0.0f ? 1.0f : -1.0f));
}
BPT_HD float3 oct_decode(float2 f) {                    // pack.hlsl:99-104
    float3 n = v3(f.x, f.y, (1.0f - fabsf(f.x)) - fabsf(f.y));
    float t = clampf_(-n.z, 0.0f, 1.0f);
    n.x = n.x + (n.x >= 0.0f ? -t : t);
    n.y = n.y + (n.y >= 0.0f ? -t : t);
    return normalize3(n);
}
// (N, T) into gbuffer.normal_roughness.xyz (pack_normal_and_tangent, pack.hlsl:112-120) stored as three halves, and back
// (unpack_normal_and_tangent, pack.hlsl:121-129).
BPT_HD float3 frame_to_gbuffer(float3 N, float3 T) {
    float2 oct = oct_encode(N);
    Frame3 fr = frame_from_normal(N);
    float px = dot3(T, fr.x), py = dot3(T, fr.y);
    float lnorm = fabsf(px) + fabsf(py);
    px = px / lnorm; py = py / lnorm;
    float packed_x = px * 0.5f + 0.5f;
    return v3(q_half(oct.x), q_half(oct.y), q_half(py < 0.0f ? -packed_x : packed_x));
}
BPT_HD void frame_from_gbuffer(float3 packed, float3& No, float3& To) {
    No = oct_decode(make_float2(packed.x, packed.y));
    float sign = packed.z < 0.0f ? -1.0f : 1.0f;
    float projected_x = (sign * packed.z) * 2.0f - 1.0f;
    float projected_y = sign * (1.0f - fabsf(projected_x));
    Frame3 fo = frame_from_normal(No);
    To = normalize3(fo.x * projected_x + fo.y * projected_y);
}
BPT_HD void frame_through_gbuffer(float3 N, float3 T, float3& No, float3& To) { frame_from_gbuffer(frame_to_gbuffer(N, T), No, To); }

// ---- surface / BSDF (material/utils.hlsl, material/lit.hlsl) ----------------------------------
struct Surface {
    float3 base_color, f0_color, f90_color, normal_map_value;
    float roughness, anisotropy, ior, opacity;
    bool two_sided;
};
BPT_HD Surface surface_default() {                               // utils.hlsl:18-31
    Surface s;
    s.base_color = v3s(0.5f); s.f0_color = v3s(0.04f); s.f90_color = v3s(1.0f);
    s.normal_map_value = v3(0.5f, 0.5f, 1.0f);
    s.roughness = 0.5f; s.anisotropy = 0.0f; s.ior = 1.5f; s.opacity = 1.0f; s.two_sided = false;
    return s;
}
// pack_surface_to_gbuffer -> texture formats -> unpack_gbuffer_to_surface (gbuffer.hlsl:18-45), state_precision =
// reference_fp16: base colour and roughness as halves, Fresnel colours as rg11b10 in unorm16 pairs, anisotropy /
// 1/ior / surface model as unorm8; emission and two_sided are not stored, opacity reads back as 1.
BPT_HD void surface_through_gbuffer(Surface& s, uint32_t& surface_model) {
    s.base_color = q_half3(s.base_color);
    s.f0_color = fresnel_through_gbuffer(s.f0_color);
    s.f90_color = fresnel_through_gbuffer(s.f90_color);
    s.roughness = q_half(s.roughness);
    s.anisotropy = q_unorm(s.anisotropy, 255.0f);
    s.ior = 1.0f / q_unorm(1.0f / s.ior, 255.0f);
    surface_model = ftou(q_unorm((float)surface_model / 256.0f, 255.0f) * 255.5f);
    s.opacity = 1.0f;
}
BPT_HD float3 schlick_fresnel(float3 f0, float3 f90, float cos_theta, float ior) {   // utils.hlsl:40-51
    if (cos_theta < 0.0f) {
        float eta = 1.0f / ior;
        float sin_theta_sqr = eta * eta * (1.0f - cos_theta * cos_theta);
        cos_theta = sqrtf(tmax_(1.0f - sin_theta_sqr, 0.0f));
    }
    return mix3(f0, f90, pow5f(1.0f - cos_theta));
}
BPT_HD void aniso_roughness(float roughness, float anisotropy, float& rx, float& ry) {   // utils.hlsl:54-59
    float aniso = sqrtf(1.0f - anisotropy * 0.9f);
    float r2 = roughness * roughness;
    rx = tmax_(r2 / aniso, 0.001f);
    ry = tmax_(r2 * aniso, 0.001f);
}
BPT_HD float ggx_ndf(float3 h, float rx, float ry) {                                      // utils.hlsl:62-65
    float a = (sq(h.x / rx) + sq(h.y / ry)) + sq(h.z);
    return kInvPi / (rx * ry * a * a);
}
BPT_HD float ggx_g1(float3 v, float rx, float ry) {                                       // utils.hlsl:67-70
    float a = (sq(rx * v.x) + sq(ry * v.y)) / tmax_(v.z * v.z, 0.0001f);
    return 2.0f / (1.0f + sqrtf(1.0f + a));
}
BPT_HD float ggx_visible_hc(float3 v, float3 l, float rx, float ry) {                     // utils.hlsl:92-96
    float vv = l.z * sqrtf((sq(rx * v.x) + sq(ry * v.y)) + sq(v.z));
    float ll = v.z * sqrtf((sq(rx * l.x) + sq(ry * l.y)) + sq(l.z));
    return 0.5f / tmax_(vv + ll, 0.0001f);
}
BPT_HD float3 ggx_vndf_sample(float3 v, float rx, float ry, float rand_x, float rand_y) { // utils.hlsl:98-116
    if (v.z < 0.0f) v = -v;
    float3 vh = normalize3(v3(rx * v.x, ry * v.y, v.z));
    float len_sqr = vh.x * vh.x + vh.y * vh.y;
    float3 t1v = len_sqr > 0.0f ? v3(-vh.y, vh.x, 0.0f) / sqrtf(len_sqr) : v3(1.0f, 0.0f, 0.0f);
    float3 t2v = cross3(vh, t1v);
    float r = sqrtf(rand_x);
    float sn, cs;
    sincos_2pi(rand_y, sn, cs);
    float t1 = r * cs;
    float t2 = r * sn;
    float s = 0.5f * (1.0f + vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    float3 nh = (t1 * t1v + t2 * t2v) + sqrtf(tmax_((1.0f - t1 * t1) - t2 * t2, 0.0f)) * vh;
    return normalize3(v3(rx * nh.x, ry * nh.y, tmax_(nh.z, 0.0f)));
}
BPT_HD float ggx_vndf_pdf(float3 h, float3 v, float rx, float ry) {                        // utils.hlsl:118-121
    return ggx_g1(v, rx, ry) * ggx_ndf(h, rx, ry) * tmax_(dot3(h, v), 0.0f) / tmax_(v.z, 0.0001f);
}
// surface_eval (material.hlsl:81-118 → lit.hlsl:5-35): diffuse + specular; unlit/none → 0
BPT_HD float3 bsdf_eval(float3 N, float3 T, float3 B, float3 V, float3 L, const Surface& s, uint32_t surface_model) {
    if (surface_model != 1u) return v3s(0.0f);
    float3 H = normalize3(V + L);
    float3 lh = v3(dot3(H, T), dot3(H, B), dot3(H, N));
    float3 lv = v3(dot3(V, T), dot3(V, B), dot3(V, N));
    float3 ll = v3(dot3(L, T), dot3(L, B), dot3(L, N));
    if (lv.z <= 0.0f || ll.z <= 0.0f) return v3s(0.0f);
    float3 fr = schlick_fresnel(s.f0_color, s.f90_color, tmax_(dot3(V, H), 0.0f), s.ior);
    float3 diffuse = (v3s(1.0f) - fr) * s.base_color * kInvPi * tmax_(ll.z, 0.0f);
    float rx, ry;
    aniso_roughness(s.roughness, s.anisotropy, rx, ry);
    float ndf = ggx_ndf(lh, rx, ry);
    float vis = ggx_visible_hc(lv, ll, rx, ry);
    float3 specular = fr * ndf * vis * tmax_(ll.z, 0.0f);
    return diffuse + specular;
}
// the specular term alone: the `bsdf_specular` out-parameter of surface_eval (material.hlsl:81-118 / lit.hlsl:5-35)
BPT_HD float3 bsdf_eval_specular(float3 N, float3 T, float3 B, float3 V, float3 L, const Surface& s, uint32_t surface_model) {
    if (surface_model != 1u) return v3s(0.0f);
    float3 H = normalize3(V + L);
    float3 lh = v3(dot3(H, T), dot3(H, B), dot3(H, N));
    float3 lv = v3(dot3(V, T), dot3(V, B), dot3(V, N));
    float3 ll = v3(dot3(L, T), dot3(L, B), dot3(L, N));
    if (lv.z <= 0.0f || ll.z <= 0.0f) return v3s(0.0f);
    float3 fr = schlick_fresnel(s.f0_color, s.f90_color, tmax_(dot3(V, H), 0.0f), s.ior);
    float rx, ry;
    aniso_roughness(s.roughness, s.anisotropy, rx, ry);
    float ndf = ggx_ndf(lh, rx, ry);
    float vis = ggx_visible_hc(lv, ll, rx, ry);
    return fr * ndf * vis * tmax_(ll.z, 0.0f);
}
// surface_eval_lut (lit.hlsl:37-58)
BPT_HD float3 bsdf_eval_lut(float3 N, float3 V, const Surface& s, float3 int_diffuse, float3 int_specular, float2 int_brdf, uint32_t surface_model) {
    if (surface_model != 1u) return v3s(0.0f);
    float ndotv = dot3(N, V);
    if (ndotv <= 0.0f) return v3s(0.0f);
    float3 fr = schlick_fresnel(s.f0_color, s.f90_color, ndotv, s.ior);
    float3 diffuse = (v3s(1.0f) - fr) * s.base_color * kInvPi;
    float3 specular = s.f0_color * int_brdf.x + s.f90_color * int_brdf.y;
    return diffuse * int_diffuse + specular * int_specular;
}
BPT_HD float3 cos_hemisphere_sample(float rand_x, float rand_y) {                           // sampling.hlsl:24-28
    float sn, cs;
    sincos_2pi(rand_x, sn, cs);
    float r = sqrtf(rand_y);
    return v3(cs * r, sn * r, sqrtf(tmax_(1.0f - rand_y, 0.0f)));
}
BPT_HD float3 uniform_sphere_sample(float rand_x, float rand_y) {                           // sampling.hlsl:14-19
    float sn, cs;
    sincos_2pi(rand_x, sn, cs);
    float z = rand_y * 2.0f - 1.0f;
    float r = sqrtf(tmax_(1.0f - z * z, 0.0f));
    return v3(cs * r, sn * r, z);
}

} // namespace bptd
