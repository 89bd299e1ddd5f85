// host/math.hpp — minimal float3 / float4x4 (column-major, glm conventions) for the host mirror.
// The reference uses glm through aliases (bisemutum/include/bisemutum/math/math.hpp:11-42) with
// GLM_FORCE_DEPTH_ZERO_TO_ONE (:3); only the handful of functions the path-tracing path touches
// are provided here.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>

namespace bi {

struct float3 {
    float x = 0, y = 0, z = 0;
    float3() = default;
    float3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    explicit float3(float s) : x(s), y(s), z(s) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    bool operator==(float3 const& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct float4 {
    float x = 0, y = 0, z = 0, w = 0;
    float4() = default;
    float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    float4(float3 const& v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};
inline float3 operator+(float3 a, float3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(float3 a, float3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator-(float3 a) { return {-a.x, -a.y, -a.z}; }
inline float3 operator*(float3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float3 a, float3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }

// column-major 4x4: m[c][r]
struct float4x4 {
    std::array<std::array<float, 4>, 4> m{};
    float4x4() = default;
    explicit float4x4(float d) { for (int i = 0; i < 4; i++) m[i][i] = d; }
    std::array<float, 4>& operator[](int c) { return m[c]; }
    std::array<float, 4> const& operator[](int c) const { return m[c]; }
    bool operator==(float4x4 const& o) const { return m == o.m; }
    float const* data() const { return &m[0][0]; }
};

namespace math {
inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float3 cross(float3 a, float3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float3 normalize(float3 v) { float inv = 1.0f / std::sqrt(dot(v, v)); return v * inv; }
inline float3 abs(float3 v) { return {std::fabs(v.x), std::fabs(v.y), std::fabs(v.z)}; }
inline float3 min(float3 a, float3 b) { return {a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z}; }
inline float3 max(float3 a, float3 b) { return {a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z}; }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

float4x4 operator_mul(float4x4 const& a, float4x4 const& b);
float4x4 lookAt(float3 eye, float3 center, float3 up);                                       // glm::lookAtRH
float4x4 perspective_reverse_z(float fovy, float aspect, float near, float far);             // src/math/math.cpp:5-11
float4x4 ortho_reverse_z(float left, float right, float bottom, float top, float near, float far);   // src/math/math.cpp:13-19
float4x4 inverse(float4x4 const& m);
float3 rotate_direction(float angle, float3 axis, float3 v);                                 // (glm::rotate(mat4(1), angle, axis) * float4(v, 0)).xyz
} // namespace math

inline float4x4 operator*(float4x4 const& a, float4x4 const& b) { return math::operator_mul(a, b); }

// bisemutum/include/bisemutum/math/bbox.hpp + src/math/bbox.cpp
struct BoundingBox {
    float3 p_min{3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float3 p_max{-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    float3 center() const { return (p_min + p_max) * 0.5f; }
    float3 extent() const { return p_max - p_min; }
    BoundingBox& add(BoundingBox const& b) { p_min = math::min(p_min, b.p_min); p_max = math::max(p_max, b.p_max); return *this; }
    bool test_with_planes(float4 const* planes, size_t count) const;                         // bbox.cpp:43-56
};
BoundingBox transform_bounding_box(float const m3x4[12], BoundingBox const& b);              // transform.cpp:59-78

} // namespace bi
