// oracle_host.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
// Host-side restatements: camera matrices, frustum planes, AABB-vs-planes culling.
//   Camera::update_shader_params   bisemutum/src/graphics/camera.cpp:73-118
//   perspective_reverse_z          bisemutum/src/math/math.cpp:5-21
//   Camera::get_frustum_planes     bisemutum/src/graphics/camera.cpp:126-174
//   BoundingBox::test_with_planes  bisemutum/src/math/bbox.cpp:43-56
//   Transform::transform_bounding_box  bisemutum/src/math/transform.cpp:59-78
// glm (un-vendored, unpinned xmake package) supplies lookAt / perspective / inverse / rotate in
// the reference; their published formulas (glm/ext/matrix_transform.inl, matrix_clip_space.inl,
// RH + GLM_FORCE_DEPTH_ZERO_TO_ONE as bisemutum/include/bisemutum/math/math.hpp:3 sets) are
// restated here. Matrices are column-major: m[c*4 + r].
#include <cmath>
#include "oracle.h"
#include "oracle_math.hpp"

using namespace orc;

namespace {
struct M4 { float m[16]; float& at(int c, int r) { return m[c * 4 + r]; } float at(int c, int r) const { return m[c * 4 + r]; } };

M4 identity() { M4 r{}; for (int i = 0; i < 4; i++) r.at(i, i) = 1.0f; return r; }
M4 mul(const M4& a, const M4& b) {
    M4 r{};
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a.at(k, row) * b.at(c, k);
            r.at(c, row) = s;
        }
    return r;
}
// glm::lookAtRH
M4 look_at(f3 eye, f3 center, f3 up) {
    f3 f = normalize(center - eye);
    f3 s = normalize(cross(f, up));
    f3 u = cross(s, f);
    M4 r = identity();
    r.at(0, 0) = s.x; r.at(1, 0) = s.y; r.at(2, 0) = s.z;
    r.at(0, 1) = u.x; r.at(1, 1) = u.y; r.at(2, 1) = u.z;
    r.at(0, 2) = -f.x; r.at(1, 2) = -f.y; r.at(2, 2) = -f.z;
    r.at(3, 0) = -dot(s, eye); r.at(3, 1) = -dot(u, eye); r.at(3, 2) = dot(f, eye);
    return r;
}
// glm::perspectiveRH_ZO then math.cpp:5-11
M4 perspective_reverse_z(float fovy, float aspect, float zn, float zf) {
    float th = std::tan(fovy / 2.0f);
    M4 r{};
    r.at(0, 0) = 1.0f / (aspect * th);
    r.at(1, 1) = 1.0f / th;
    r.at(2, 2) = zf / (zn - zf);
    r.at(2, 3) = -1.0f;
    r.at(3, 2) = -(zf * zn) / (zf - zn);
    float inv = 1.0f / (zf - zn);
    r.at(2, 2) = zn * inv;
    r.at(3, 2) = zn * zf * inv;
    return r;
}
// glm::orthoRH_ZO then math.cpp:13-19
M4 ortho_reverse_z(float l, float rgt, float b, float t, float zn, float zf) {
    M4 r = identity();
    r.at(0, 0) = 2.0f / (rgt - l);
    r.at(1, 1) = 2.0f / (t - b);
    r.at(2, 2) = -1.0f / (zf - zn);
    r.at(3, 0) = -(rgt + l) / (rgt - l);
    r.at(3, 1) = -(t + b) / (t - b);
    r.at(3, 2) = -zn / (zf - zn);
    float inv = 1.0f / (zf - zn);
    r.at(2, 2) = inv;
    r.at(3, 2) = zf * inv;
    return r;
}
// general 4x4 inverse by cofactors (double accumulation is NOT used: FP32 like glm)
M4 inverse(const M4& a) {
    const float* m = a.m;
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
    M4 r{};
    for (int i = 0; i < 16; i++) r.m[i] = inv[i] * idet;
    return r;
}
// glm::rotate(mat4(1), angle, axis) applied to (v, 0): Rodrigues in glm's matrix form
f3 rotate_dir(float angle, f3 axis_in, f3 v) {
    float c = std::cos(angle), s = std::sin(angle);
    f3 axis = normalize(axis_in);
    f3 temp = axis * (1.0f - c);
    float r00 = c + temp.x * axis.x, r01 = temp.x * axis.y + s * axis.z, r02 = temp.x * axis.z - s * axis.y;
    float r10 = temp.y * axis.x - s * axis.z, r11 = c + temp.y * axis.y, r12 = temp.y * axis.z + s * axis.x;
    float r20 = temp.z * axis.x + s * axis.y, r21 = temp.z * axis.y - s * axis.x, r22 = c + temp.z * axis.z;
    // column-major Rotate[c][r]; result = Rotate * v
    return mk3(r00 * v.x + r10 * v.y + r20 * v.z, r01 * v.x + r11 * v.y + r21 * v.z, r02 * v.x + r12 * v.y + r22 * v.z);
}
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
} // namespace

extern "C" {

void obpt_camera_matrices(const obpt_camera_desc* cam, float view[16], float proj[16], bpt_camera* out) {
    f3 pos = mk3(cam->position[0], cam->position[1], cam->position[2]);
    f3 front = mk3(cam->front_dir[0], cam->front_dir[1], cam->front_dir[2]);
    f3 up = mk3(cam->up_dir[0], cam->up_dir[1], cam->up_dir[2]);
    M4 v = look_at(pos, pos + front, up);                                                // camera.cpp:96
    M4 p;
    if (!cam->orthographic) p = perspective_reverse_z(radians(cam->yfov), cam->aspect, cam->near_z, cam->far_z);   // camera.cpp:98
    else {
        float oh = std::tan(radians(cam->yfov * 0.5f));                                   // camera.cpp:100-104
        float ow = oh * cam->aspect;
        p = ortho_reverse_z(-ow, ow, -oh, oh, cam->near_z, cam->far_z);
    }
    M4 iv = inverse(v), ip = inverse(p), pv = mul(p, v);                                  // camera.cpp:106-108
    if (view) for (int i = 0; i < 16; i++) view[i] = v.m[i];
    if (proj) for (int i = 0; i < 16; i++) proj[i] = p.m[i];
    if (out) for (int i = 0; i < 16; i++) { out->matrix_inv_view[i] = iv.m[i]; out->matrix_inv_proj[i] = ip.m[i]; out->matrix_proj_view[i] = pv.m[i]; }
}

void obpt_frustum_planes(const obpt_camera_desc* cam, float planes[24]) {              // camera.cpp:126-174
    f3 pos = mk3(cam->position[0], cam->position[1], cam->position[2]);
    f3 front = normalize(mk3(cam->front_dir[0], cam->front_dir[1], cam->front_dir[2]));
    f3 up_dir = mk3(cam->up_dir[0], cam->up_dir[1], cam->up_dir[2]);
    f3 right = normalize(cross(front, up_dir));
    f3 up = cross(right, front);
    float xfov = cam->yfov * cam->aspect;                                                 // camera.cpp:138 (sic)
    auto set = [&](int i, f3 n, float w) { planes[i * 4] = n.x; planes[i * 4 + 1] = n.y; planes[i * 4 + 2] = n.z; planes[i * 4 + 3] = w; };
    float pos_dot_front = dot(pos, front);
    set(0, front, -cam->near_z - pos_dot_front);
    set(1, -front, pos_dot_front + cam->far_z);
    if (!cam->orthographic) {
        float vert_angle = radians(90.0f - cam->yfov * 0.5f);
        f3 n2 = rotate_dir(-vert_angle, right, front); set(2, n2, -dot(pos, n2));
        f3 n3 = rotate_dir(vert_angle, right, front); set(3, n3, -dot(pos, n3));
        float hori_angle = radians(90.0f - xfov * 0.5f);
        f3 n4 = rotate_dir(-hori_angle, up, front); set(4, n4, -dot(pos, n4));
        f3 n5 = rotate_dir(hori_angle, up, front); set(5, n5, -dot(pos, n5));
    } else {
        float oh = std::tan(radians(cam->yfov * 0.5f));
        float ow = oh * cam->aspect;
        float pos_dot_up = dot(pos, up);
        set(2, -up, oh + pos_dot_up);
        set(3, up, oh - pos_dot_up);
        set(4, right, ow - pos_dot_up);                                                   // camera.cpp:169-170 (sic: pos_dot_up)
        set(5, -right, ow + pos_dot_up);
    }
}

void obpt_cull_aabbs(const float planes[24], const float* mm, uint32_t n, uint8_t* visible) {   // bbox.cpp:43-56
    for (uint32_t i = 0; i < n; i++) {
        f3 lo = mk3(mm[i * 6], mm[i * 6 + 1], mm[i * 6 + 2]), hi = mk3(mm[i * 6 + 3], mm[i * 6 + 4], mm[i * 6 + 5]);
        f3 extent = (hi - lo) * 0.5f;
        f3 center = (lo + hi) * 0.5f;
        bool vis = true;
        for (int p = 0; p < 6; p++) {
            f3 pn = mk3(planes[p * 4], planes[p * 4 + 1], planes[p * 4 + 2]);
            float box_radius = dot(mk3(fabsf(pn.x), fabsf(pn.y), fabsf(pn.z)), extent);
            float plane_dist = dot(pn, center) + planes[p * 4 + 3];
            if (plane_dist <= -box_radius) { vis = false; break; }
        }
        visible[i] = vis ? 1 : 0;
    }
}

void obpt_transform_aabb(const float m[12], const float in[6], float out[6]) {          // transform.cpp:59-78
    f3 mn = splat3(3.402823466e+38f), mx = splat3(-3.402823466e+38f);
    for (int c = 0; c < 8; c++) {
        f3 p = mk3((c & 4) ? in[3] : in[0], (c & 2) ? in[4] : in[1], (c & 1) ? in[5] : in[2]);
        f3 w = mk3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
                   ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
        mn = min3(mn, w); mx = max3(mx, w);
    }
    out[0] = mn.x; out[1] = mn.y; out[2] = mn.z; out[3] = mx.x; out[4] = mx.y; out[5] = mx.z;
}

} // extern "C"
