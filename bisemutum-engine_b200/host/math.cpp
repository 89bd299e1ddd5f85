// host/math.cpp — see math.hpp. glm's published formulas (lookAtRH, perspectiveRH_ZO, orthoRH_ZO,
// rotate) restated; glm itself is an un-vendored, unpinned dependency of the reference.
#include "math.hpp"

namespace bi {
namespace math {

float4x4 operator_mul(float4x4 const& a, float4x4 const& b) {
    float4x4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a[k][row] * b[c][k];
            r[c][row] = s;
        }
    return r;
}

float4x4 lookAt(float3 eye, float3 center, float3 up) {
    float3 f = normalize(center - eye);
    float3 s = normalize(cross(f, up));
    float3 u = cross(s, f);
    float4x4 r(1.0f);
    r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
    r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
    r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
    r[3][0] = -dot(s, eye); r[3][1] = -dot(u, eye); r[3][2] = dot(f, eye);
    return r;
}

float4x4 perspective_reverse_z(float fovy, float aspect, float near, float far) {
    float const tan_half = std::tan(fovy / 2.0f);
    float4x4 mat;
    mat[0][0] = 1.0f / (aspect * tan_half);
    mat[1][1] = 1.0f / tan_half;
    mat[2][2] = far / (near - far);
    mat[2][3] = -1.0f;
    mat[3][2] = -(far * near) / (far - near);
    auto inv = 1.0f / (far - near);
    mat[2][2] = near * inv;
    mat[3][2] = near * far * inv;
    return mat;
}

float4x4 ortho_reverse_z(float left, float right, float bottom, float top, float near, float far) {
    float4x4 mat(1.0f);
    mat[0][0] = 2.0f / (right - left);
    mat[1][1] = 2.0f / (top - bottom);
    mat[2][2] = -1.0f / (far - near);
    mat[3][0] = -(right + left) / (right - left);
    mat[3][1] = -(top + bottom) / (top - bottom);
    mat[3][2] = -near / (far - near);
    auto inv = 1.0f / (far - near);
    mat[2][2] = inv;
    mat[3][2] = far * inv;
    return mat;
}

// cofactor expansion on the flat column-major array
float4x4 inverse(float4x4 const& a) {
    float const* m = a.data();
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float idet = 1.0f / det;
    float4x4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) r[c][row] = inv[c * 4 + row] * idet;
    return r;
}

float3 rotate_direction(float angle, float3 axis_in, float3 v) {
    float const c = std::cos(angle), s = std::sin(angle);
    float3 axis = normalize(axis_in);
    float3 temp = axis * (1.0f - c);
    float r00 = c + temp.x * axis.x, r01 = temp.x * axis.y + s * axis.z, r02 = temp.x * axis.z - s * axis.y;
    float r10 = temp.y * axis.x - s * axis.z, r11 = c + temp.y * axis.y, r12 = temp.y * axis.z + s * axis.x;
    float r20 = temp.z * axis.x + s * axis.y, r21 = temp.z * axis.y - s * axis.x, r22 = c + temp.z * axis.z;
    return {r00 * v.x + r10 * v.y + r20 * v.z, r01 * v.x + r11 * v.y + r21 * v.z, r02 * v.x + r12 * v.y + r22 * v.z};
}

} // namespace math

bool BoundingBox::test_with_planes(float4 const* planes, size_t count) const {
    auto extent = this->extent() * 0.5f;
    auto center = this->center();
    for (size_t i = 0; i < count; i++) {
        auto const& plane = planes[i];
        float3 n{plane.x, plane.y, plane.z};
        auto box_radius = math::dot(math::abs(n), extent);
        auto plane_dist = math::dot(n, center) + plane.w;
        if (plane_dist <= -box_radius) return false;
    }
    return true;
}

BoundingBox transform_bounding_box(float const m[12], BoundingBox const& b) {
    BoundingBox r;
    for (int c = 0; c < 8; c++) {
        float3 p{(c & 4) ? b.p_max.x : b.p_min.x, (c & 2) ? b.p_max.y : b.p_min.y, (c & 1) ? b.p_max.z : b.p_min.z};
        float3 w{((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
                 ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]};
        r.p_min = math::min(r.p_min, w);
        r.p_max = math::max(r.p_max, w);
    }
    return r;
}

} // namespace bi
