// oracle_render.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Per-pixel restatement of the reference's wavefront path tracer:
//   host loop        bisemutum/src/renderer/pass/path_tracing.cpp:224-488
//   camera ray       shaders/renderer/raytracing/direction_sample/generate_camera_ray.hlsl:4-16
//   trace            shaders/renderer/raytracing/rt_gbuffer.hlsl:7-36
//   closest hit      shaders/renderer/raytracing/hits/rt_gbuffer_hit.hlsl:6-18
//                    shaders/core/raytracing/hit.hlsl:27-173
//   lighting         shaders/renderer/raytracing/deferred_lighting_secondary.hlsl:11-111
//                    shaders/renderer/lights.hlsl:10-25
//   next direction   shaders/renderer/raytracing/direction_sample/sample_secondary_ray.hlsl:11-69
//   accumulate       shaders/renderer/raytracing/pt_accumulate.hlsl:3-11
// NEE visibility is a shadow ray instead of the reference's rasterised shadow maps (NEW).
// state_precision = fp32 (default): state and colours stay FP32, accumulation is an FP32 sum / N.
// state_precision = reference_fp16: every value the reference passes between its passes through a texture takes that
// texture's format (path_tracing.cpp:248-288, pass/gbuffer.hpp:14-17): half ray directions / throughput / colours,
// the packed G-buffer, the half additive blit per bounce, and the running half lerp of pt_accumulate.
#include <algorithm>
#include <cfloat>
#include <thread>
#include <type_traits>
#include "oracle_scene.hpp"
#include "oracle_ltc.hpp"

namespace orc {

// ---- textures: explicit FP32 bilinear (no 8-bit HW filter weights; SURVEY §7 hard parts) -----
static inline f4 fetch_texel(const Texture& t, int x, int y) {
    size_t i = (size_t)y * t.w + x;
    if (t.format == BPT_TEXTURE_RGBA8_UNORM) {
        const uint8_t* p = &t.texels[i * 4];
        return f4{(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f};
    }
    const float* p = reinterpret_cast<const float*>(t.texels.data()) + i * 4;
    return f4{p[0], p[1], p[2], p[3]};
}
static inline int wrap_coord(int c, int n, uint32_t mode) {
    if (mode == BPT_ADDRESS_REPEAT) { int m = c % n; return m < 0 ? m + n : m; }
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
static inline f4 lerp4(f4 a, f4 b, float t) {
    return f4{a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t};
}
static inline f4 sample_texture(const Texture& t, float u, float v) {
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    if (!t.linear) {
        int xi = wrap_coord((int)floorf(u * (float)t.w), (int)t.w, t.addr_u);
        int yi = wrap_coord((int)floorf(v * (float)t.h), (int)t.h, t.addr_v);
        return fetch_texel(t, xi, yi);
    }
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap_coord((int)x0f, (int)t.w, t.addr_u), x1 = wrap_coord((int)x0f + 1, (int)t.w, t.addr_u);
    int y0 = wrap_coord((int)y0f, (int)t.h, t.addr_v), y1 = wrap_coord((int)y0f + 1, (int)t.h, t.addr_v);
    f4 top = lerp4(fetch_texel(t, x0, y0), fetch_texel(t, x1, y0), fx);
    f4 bot = lerp4(fetch_texel(t, x0, y1), fetch_texel(t, x1, y1), fx);
    return lerp4(top, bot, fy);
}
static inline f4 sample_or_default(const Scene& sc, int32_t tex, f2 uv, f4 dflt) {
    if (tex < 0 || (size_t)tex >= sc.textures.size()) return dflt;
    return sample_texture(sc.textures[tex], uv.x, uv.y);
}

// ---- sky: deferred_lighting_secondary.hlsl:24-29. TextureCube.SampleLevel with the linear skybox sampler: face / (s, t) selection
//      is the Vulkan cube rule (the inverse of core/utils/cubemap.hlsl:3-21); bilinear, SEAMLESS across edges as Vulkan and D3D12
//      filter cube maps: taps beyond an edge come from the adjacent face, the missing fourth tap at a corner is the mean of the three.
struct CubeCoord { int face; float s, t; bool ok; };
static inline CubeCoord cube_coord(f3 d) {
    CubeCoord c{0, 0.0f, 0.0f, false};
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z), sc_, tc, ma;
    if (ax >= ay && ax >= az) { c.face = d.x >= 0.0f ? 0 : 1; ma = ax; sc_ = d.x >= 0.0f ? -d.z : d.z; tc = -d.y; }
    else if (ay >= az) { c.face = d.y >= 0.0f ? 2 : 3; ma = ay; sc_ = d.x; tc = d.y >= 0.0f ? d.z : -d.z; }
    else { c.face = d.z >= 0.0f ? 4 : 5; ma = az; sc_ = d.z >= 0.0f ? d.x : -d.x; tc = -d.y; }
    if (!(ma > 0.0f)) return c;
    c.s = sc_ / ma; c.t = tc / ma; c.ok = true;
    return c;
}
static inline f3 cube_point(int face, float s, float t) {           // cubemap.hlsl:3-21 before the normalize; (s, t) may leave [-1, 1]
    static const float sx[6][3] = {{0, 0, -1}, {0, 0, 1}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {-1, 0, 0}};     // d = major + s * sx + t * tx
    static const float tx[6][3] = {{0, -1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}, {0, -1, 0}, {0, -1, 0}};
    static const float mj[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    // one non-zero term per component: no rounding is involved, the sums below are exact selections
    return mk3(mj[face][0] + (sx[face][0] * s + tx[face][0] * t), mj[face][1] + (sx[face][1] * s + tx[face][1] * t), mj[face][2] + (sx[face][2] * s + tx[face][2] * t));
}
static inline f3 cube_texel(const float* faces, int n, int face, int x, int y) {
    if (x < 0 || x >= n || y < 0 || y >= n) {                        // across an edge: centre of the virtual texel, re-projected
        const float inv = 1.0f / (float)n;
        CubeCoord c = cube_coord(cube_point(face, (2.0f * (float)x + 1.0f) * inv - 1.0f, (2.0f * (float)y + 1.0f) * inv - 1.0f));
        face = c.face;
        x = (int)floorf(0.5f * (c.s + 1.0f) * (float)n); y = (int)floorf(0.5f * (c.t + 1.0f) * (float)n);
        x = std::min(std::max(x, 0), n - 1); y = std::min(std::max(y, 0), n - 1);
    }
    const float* p = faces + (((size_t)face * n + y) * n + x) * 4;
    return mk3(p[0], p[1], p[2]);
}
static inline f3 sample_cube(const float* faces, uint32_t size, f3 d) {
    if (size == 0) return splat3(0.0f);
    CubeCoord cc = cube_coord(d);
    if (!cc.ok) return splat3(0.0f);
    float u = 0.5f * (cc.s + 1.0f), v = 0.5f * (cc.t + 1.0f);
    int n = (int)size;
    float x = u * (float)n - 0.5f, y = v * (float)n - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int xs[2] = {(int)x0f, (int)x0f + 1}, ys[2] = {(int)y0f, (int)y0f + 1};
    f3 tap[2][2];                                                    // [row][column]
    int bad_r = -1, bad_c = -1;
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 2; c++) {
            bool out_x = xs[c] < 0 || xs[c] >= n, out_y = ys[r] < 0 || ys[r] >= n;
            if (out_x && out_y) { bad_r = r; bad_c = c; continue; }  // beyond a cube corner: no such texel
            tap[r][c] = cube_texel(faces, n, cc.face, xs[c], ys[r]);
        }
    if (bad_r >= 0)                                                  // mean of the three that exist: (in-face + across x) + across y
        tap[bad_r][bad_c] = ((tap[1 - bad_r][1 - bad_c] + tap[1 - bad_r][bad_c]) + tap[bad_r][1 - bad_c]) * (1.0f / 3.0f);
    f3 top = lerp3(tap[0][0], tap[0][1], fx);
    f3 bot = lerp3(tap[1][0], tap[1][1], fx);
    return lerp3(top, bot, fy);
}
static inline f3 sample_sky(const Scene& sc, f3 d) { return sample_cube(sc.sky_faces.data(), sc.sky_size, d); }

// ---- image-based lighting: the IBL block of deferred_lighting_secondary.hlsl:98-108 (textures: obpt_precompute_sky_ibl below).
// skybox_sampler is linear / clamp_to_edge with the default NEAREST mip mode (src/renderer/context/skybox.cpp:31-37,
// include/bisemutum/rhi/sampler.hpp:40): the specular level is ceil(lod + 0.5) - 1 (the Vulkan rule for nearest mip selection).
static inline f3 ibl_lighting(const Scene& sc, f3 N, f3 V, const SurfaceData& surf, uint32_t surface_model) {
    const float* m = sc.sky_transform;
    auto xf = [&](f3 d) { return mk3((m[0] * d.x + m[1] * d.y) + m[2] * d.z, (m[3] * d.x + m[4] * d.y) + m[5] * d.z, (m[6] * d.x + m[7] * d.y) + m[8] * d.z); };
    const bpt_sky_ibl_desc& d = sc.ibl_desc;
    f3 diffuse_color = mk3(sc.sky_color[0] * d.diffuse_strength, sc.sky_color[1] * d.diffuse_strength, sc.sky_color[2] * d.diffuse_strength);      // skybox.cpp:42
    f3 specular_color = mk3(sc.sky_color[0] * d.specular_strength, sc.sky_color[1] * d.specular_strength, sc.sky_color[2] * d.specular_strength);  // skybox.cpp:43
    f3 ibl_diffuse = sample_cube(sc.ibl_diffuse.data(), d.diffuse_size, xf(N)) * diffuse_color;                 // :99-100
    float lod = surf.roughness * (float)(d.specular_levels - 1u);                                               // :102-104
    int level = (int)ceilf(lod + 0.5f) - 1;
    level = level < 0 ? 0 : (level >= (int)d.specular_levels ? (int)d.specular_levels - 1 : level);
    size_t offset = 0;
    for (int l = 0; l < level; l++) { size_t s = d.specular_size >> l; offset += 6 * s * s * 4; }
    f3 ibl_specular = sample_cube(sc.ibl_specular.data() + offset, d.specular_size >> level, xf(reflect(-V, N))) * specular_color;
    const int n = (int)d.brdf_lut_size;                                                                         // :105 bilinear, clamp
    float x = dot(N, V) * (float)n - 0.5f, y = surf.roughness * (float)n - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap_coord((int)x0f, n, BPT_ADDRESS_CLAMP), x1 = wrap_coord((int)x0f + 1, n, BPT_ADDRESS_CLAMP);
    int y0 = wrap_coord((int)y0f, n, BPT_ADDRESS_CLAMP), y1 = wrap_coord((int)y0f + 1, n, BPT_ADDRESS_CLAMP);
    auto tx = [&](int xx, int yy) { const float* p = &sc.ibl_brdf[((size_t)yy * n + xx) * 2]; return f2{p[0], p[1]}; };
    f2 a = tx(x0, y0), b = tx(x1, y0), c = tx(x0, y1), e = tx(x1, y1);
    f2 top = f2{lerpf(a.x, b.x, fx), lerpf(a.y, b.y, fx)}, bot = f2{lerpf(c.x, e.x, fx), lerpf(c.y, e.y, fx)};
    f2 brdf = f2{lerpf(top.x, bot.x, fy), lerpf(top.y, bot.y, fy)};
    return surface_eval_lut(N, V, surf, ibl_diffuse, ibl_specular, brdf, surface_model);                        // :106-107
}

// ---- vertex fetch: core/raytracing/hit.hlsl:27-164 ---------------------------------------------
struct Vertex { f3 normal_world, tangent_world, bitangent_world, position_world, color; f2 texcoord; };

static inline void fetch_indices(const Scene& sc, const bpt_drawable_sbt_data& dr, uint32_t prim, uint32_t idx[3]) {
    for (int c = 0; c < 3; c++) idx[c] = sc.indices[(size_t)dr.index_offset + 3ull * prim + c];   // hit.hlsl:28-32
}
static inline f2 fetch_texcoord(const Scene& sc, const bpt_drawable_sbt_data& dr, uint32_t va, const uint32_t idx[3], float bu, float bv) {
    if (!(va & BPT_VA_TEXCOORD)) return f2{0.0f, 0.0f};                                            // hit.hlsl:117-132
    const float* t0 = &sc.texcoords[(size_t)dr.texcoord_offset + 2ull * idx[0]];
    const float* t1 = &sc.texcoords[(size_t)dr.texcoord_offset + 2ull * idx[1]];
    const float* t2 = &sc.texcoords[(size_t)dr.texcoord_offset + 2ull * idx[2]];
    return f2{(t0[0] + (t1[0] - t0[0]) * bu) + (t2[0] - t0[0]) * bv, (t0[1] + (t1[1] - t0[1]) * bu) + (t2[1] - t0[1]) * bv};
}
static inline f3 interp3(const float* a0, const float* a1, const float* a2, float bu, float bv) {
    return mk3((a0[0] + (a1[0] - a0[0]) * bu) + (a2[0] - a0[0]) * bv,
               (a0[1] + (a1[1] - a0[1]) * bu) + (a2[1] - a0[1]) * bv,
               (a0[2] + (a1[2] - a0[2]) * bu) + (a2[2] - a0[2]) * bv);
}
static Vertex fetch_vertex_attributes(const Scene& sc, const InstanceXf& x, uint32_t prim, float bu, float bv) {
    const bpt_drawable_sbt_data& dr = sc.drawables[x.instance_id];
    uint32_t va = sc.drawable_va[x.instance_id];
    uint32_t idx[3];
    fetch_indices(sc, dr, prim, idx);
    f3 normal = mk3(0.0f, 0.0f, 1.0f);                                                             // hit.hlsl:54-72
    if (va & BPT_VA_NORMAL)
        normal = interp3(&sc.normals[(size_t)dr.normal_offset + 3ull * idx[0]], &sc.normals[(size_t)dr.normal_offset + 3ull * idx[1]],
                         &sc.normals[(size_t)dr.normal_offset + 3ull * idx[2]], bu, bv);
    f3 tangent = mk3(1.0f, 0.0f, 0.0f); float tangent_w = 1.0f;                                    // hit.hlsl:74-95
    if (va & BPT_VA_TANGENT) {
        const float* t0 = &sc.tangents[(size_t)dr.tangent_offset + 4ull * idx[0]];
        const float* t1 = &sc.tangents[(size_t)dr.tangent_offset + 4ull * idx[1]];
        const float* t2 = &sc.tangents[(size_t)dr.tangent_offset + 4ull * idx[2]];
        tangent = interp3(t0, t1, t2, bu, bv);
        tangent_w = (t0[3] + (t1[3] - t0[3]) * bu) + (t2[3] - t0[3]) * bv;
    }
    Vertex vt;
    vt.normal_world = normalize(xf_vector_transposed(x.w2o, normal));                              // hit.hlsl:157
    vt.tangent_world = normalize(xf_vector(x.o2w, tangent));                                       // hit.hlsl:158
    vt.bitangent_world = normalize(cross(vt.normal_world, vt.tangent_world)) * tangent_w;          // hit.hlsl:159
    vt.texcoord = fetch_texcoord(sc, dr, va, idx, bu, bv);
    vt.color = splat3(0.0f);                                                                        // hit.hlsl:96-113
    if ((va & BPT_VA_COLOR) && !sc.colors.empty()) {
        const float* base = &sc.colors[dr.color_offset];
        // upstream reads the third corner's colour at index.x (hit.hlsl:108-112 repeats `index.x`), so a triangle's colour varies along one
        // edge only; kept as is
        vt.color = interp3(base + 3ull * idx[0], base + 3ull * idx[1], base + 3ull * idx[0], bu, bv);
    }
    {                                                                                               // hit.hlsl:34-52,152
        const float* b = &sc.positions[dr.position_offset];
        f3 p = interp3(b + 3ull * idx[0], b + 3ull * idx[1], b + 3ull * idx[2], bu, bv);
        vt.position_world = xf_point(x.o2w, p);
    }
    return vt;
}

// ---- material_function: the closed set of snippets (hit.hlsl:166-173) --------------------------
static SurfaceData material_function(const Scene& sc, const bpt_material& m, f2 uv, f3 position_world, f3 vertex_color = f3{0.0f, 0.0f, 0.0f}) {
    SurfaceData s = surface_data_default();
    uint32_t kind = (m.flags >> BPT_MATERIAL_KIND_SHIFT) & 0xffu;
    if (kind == BPT_MATERIAL_KIND_GLTF_PBR) {                    // import_model.cpp:208-230
        f4 bt = sample_or_default(sc, m.base_color_tex, uv, f4{1, 1, 1, 1});
        f4 base = f4{bt.x * m.base_color[0], bt.y * m.base_color[1], bt.z * m.base_color[2], bt.w * m.base_color[3]};
        s.base_color = mk3(base.x, base.y, base.z);
        s.opacity = base.w;
        f4 nt = sample_or_default(sc, m.normal_map_tex, uv, f4{0.5f, 0.5f, 1.0f, 1.0f});
        f3 nm = mk3(nt.x * 2.0f - 1.0f, nt.y * 2.0f - 1.0f, nt.z * 2.0f - 1.0f);
        nm = normalize(nm * mk3(m.normal_map_scale, m.normal_map_scale, 1.0f));
        s.normal_map_value = nm * 0.5f + splat3(0.5f);
        f4 mr = sample_or_default(sc, m.metallic_roughness_tex, uv, f4{1, 1, 1, 1});
        s.roughness = m.roughness * mr.y;
        s.f0_color = lerp3(splat3(0.04f), s.base_color, m.metallic * mr.z);
        float occlusion = sample_or_default(sc, m.occlusion_tex, uv, f4{1, 1, 1, 1}).x * m.occlusion_strength;
        s.base_color = s.base_color * occlusion;
        s.f0_color = s.f0_color * occlusion;
        s.f90_color = s.f90_color * occlusion;
        s.emission = mk3(m.emission[0], m.emission[1], m.emission[2]);
        s.two_sided = (m.flags & BPT_MATERIAL_FLAG_TWO_SIDED) != 0;
    } else if (kind == BPT_MATERIAL_KIND_ASSIMP_DIFFUSE) {       // import_model.cpp:490-493
        s.base_color = mk3(m.base_color[0], m.base_color[1], m.base_color[2]);
        s.roughness = m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_CONSTANT_COLOR) {       // examples/scene_basic/materials/white.toml
        s.base_color = mk3(m.base_color[0], m.base_color[1], m.base_color[2]);
    } else if (kind == BPT_MATERIAL_KIND_CHECKERBOARD) {         // examples/scene_basic/materials/checkerboard.toml
        int grid = (int)floorf(position_world.x) ^ (int)floorf(position_world.z);
        bool odd = (grid & 1) == 1;
        s.base_color = odd ? mk3(m.base_color[0], m.base_color[1], m.base_color[2]) : mk3(m.emission[0], m.emission[1], m.emission[2]);
        s.roughness = odd ? m.base_color[3] : m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_TEXTURED) {             // examples/scene_basic/materials/textured.toml
        f4 bt = sample_or_default(sc, m.base_color_tex, uv, f4{1, 1, 1, 1});
        f4 nt = sample_or_default(sc, m.normal_map_tex, uv, f4{0.5f, 0.5f, 1.0f, 1.0f});
        s.base_color = mk3(bt.x, bt.y, bt.z);
        s.normal_map_value = mk3(nt.x, nt.y, nt.z);
        s.roughness = m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_TRANSPARENT) {          // examples/scene_basic/materials/transparent.toml
        s.base_color = mk3(m.base_color[0], m.base_color[1], m.base_color[2]);
        s.opacity = m.base_color[3];
        s.two_sided = true;
    } else if (kind == BPT_MATERIAL_KIND_VERTEX_COLOR) {         // `surface.base_color = vertex.color;`
        s.base_color = vertex_color;
        s.roughness = m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_CAGE) {                 // examples/scene_basic/materials/cage.toml
        f4 v = sample_or_default(sc, m.base_color_tex, uv, f4{1, 1, 1, 1});
        s.base_color = mk3(v.x, v.y, v.z);
        s.f0_color = mk3(v.x, v.y, v.z);
        s.opacity = v.w < 0.5f ? 0.0f : 1.0f;
        s.two_sided = true;
    }
    return s;
}

float eval_opacity(const Scene& sc, uint32_t instance_id, uint32_t prim, float u, float v) {
    const bpt_drawable_sbt_data& dr = sc.drawables[instance_id];
    const bpt_material& m = sc.materials[dr.material_offset / sizeof(bpt_material)];
    uint32_t idx[3];
    fetch_indices(sc, dr, prim, idx);
    f2 uv = fetch_texcoord(sc, dr, sc.drawable_va[instance_id], idx, u, v);
    return material_function(sc, m, uv, splat3(0.0f)).opacity;
}

// ---- lights: shaders/renderer/lights.hlsl:10-25 -------------------------------------------------
static inline f3 point_light_eval(const bpt_point_light_data& l, f3 P, f3& light_dir, float& dist) {
    f3 lv = mk3(l.position[0], l.position[1], l.position[2]) - P;
    float d2 = dot(lv, lv);
    dist = sqrtf(d2);
    light_dir = lv / dist;
    float att = saturate(1.0f - pow2(d2 * l.range_sqr_inv)) / fmax_(d2, 0.001f);
    if (l.cos_inner > l.cos_outer) {
        float ct = clampf(dot(light_dir, mk3(l.direction[0], l.direction[1], l.direction[2])), l.cos_outer, l.cos_inner);
        att = att * ((ct - l.cos_outer) / fmax_(l.cos_inner - l.cos_outer, 0.001f));
    }
    return mk3(l.emission[0], l.emission[1], l.emission[2]) * att;
}

// ---- camera ray: generate_camera_ray.hlsl:4-16, camera.hlsl:7-9 (glm column-major storage) ----
// pixel_jitter is a NEW switch (default off = the reference's fixed pixel centres, generate_camera_ray.hlsl:9):
// when on, the sub-pixel offset comes from the bounce-0 stream rng_tea(pixel, frame_index).
static inline void camera_ray(const bpt_camera& cam, uint32_t px, uint32_t py, uint32_t W, uint32_t H, uint32_t jitter, uint32_t frame_index, f3& O, f3& D) {
    const float* ip = cam.matrix_inv_proj;
    const float* iv = cam.matrix_inv_view;
    float jx = 0.5f, jy = 0.5f;
    if (jitter) { uint32_t seed = rng_tea(py * W + px, frame_index); jx = rng_next(seed); jy = rng_next(seed); }
    float uvx = ((float)px + jx) / (float)W, uvy = ((float)py + jy) / (float)H;
    float nx = uvx * 2.0f - 1.0f, ny = 1.0f - uvy * 2.0f;
    // mul(M, float4(nx, ny, 1, 1)).xyz, M column-major: M[c*4 + r]
    f3 dl = mk3(((ip[0] * nx + ip[4] * ny) + ip[8]) + ip[12],
                ((ip[1] * nx + ip[5] * ny) + ip[9]) + ip[13],
                ((ip[2] * nx + ip[6] * ny) + ip[10]) + ip[14]);
    dl = normalize(dl);
    f3 dw = mk3((iv[0] * dl.x + iv[4] * dl.y) + iv[8] * dl.z,
                (iv[1] * dl.x + iv[5] * dl.y) + iv[9] * dl.z,
                (iv[2] * dl.x + iv[6] * dl.y) + iv[10] * dl.z);
    D = normalize(dw);
    O = mk3(iv[12], iv[13], iv[14]);
}

// ---- calc_ddgi_volume_lighting: shaders/renderer/ddgi/ddgi_lighting.hlsl:7-83, one volume ---------------------------
// Generalised from DDGI_PROBES_SIZE = 8 / 6 / 14 (ddgi_struct.hlsl:3-12) to the bound volume's probe counts and atlas sizes.
// Texture2DArray.SampleLevel with the linear sampler (ddgi.cpp:14-17) = explicit FP32 bilinear inside the probe's tile.
static inline f2 oct_encode_01(f3 n) { f2 e = oct_encode(n); return f2{e.x * 0.5f + 0.5f, e.y * 0.5f + 0.5f}; }   // pack.hlsl:96-98
template <int CH>
static inline void atlas_bilinear(const float* atlas, size_t row_stride_texels, size_t tile_x0, size_t tile_y0, float cx, float cy, float out[CH]) {
    float tx = cx - 0.5f, ty = cy - 0.5f;
    float x0f = floorf(tx), y0f = floorf(ty);
    float fx = tx - x0f, fy = ty - y0f;
    size_t x0 = tile_x0 + (size_t)(int)x0f, y0 = tile_y0 + (size_t)(int)y0f;
    const float* p00 = atlas + (y0 * row_stride_texels + x0) * CH;
    const float* p10 = p00 + CH;
    const float* p01 = atlas + ((y0 + 1) * row_stride_texels + x0) * CH;
    const float* p11 = p01 + CH;
    for (int k = 0; k < CH; k++) out[k] = lerpf(lerpf(p00[k], p10[k], fx), lerpf(p01[k], p11[k], fx), fy);
}
static f4 calc_ddgi_volume_lighting(const obpt_context& ctx, f3 pos, f3 normal, f3 view) {
    const bpt_probe_volume& vol = ctx.ddgi_volume;
    const uint32_t IS = ctx.ddgi_irr_size, VS = ctx.ddgi_vis_size;
    pos = (pos + normal * 0.2f) + view * 0.8f;                                               // :16
    f3 base = mk3(vol.base_position[0], vol.base_position[1], vol.base_position[2]);
    f3 fx = mk3(vol.frame_x[0], vol.frame_x[1], vol.frame_x[2]), fy = mk3(vol.frame_y[0], vol.frame_y[1], vol.frame_y[2]), fz = mk3(vol.frame_z[0], vol.frame_z[1], vol.frame_z[2]);
    f3 vec = pos - base;
    float x = dot(vec, fx), y = dot(vec, fy), z = dot(vec, fz);                              // :18-21
    if (x < 0.0f || x > vol.extent[0] || y < 0.0f || y > vol.extent[1] || z < 0.0f || z > vol.extent[2]) return f4{0, 0, 0, 0};   // :22-28
    const uint32_t n[3] = {vol.probe_counts[0], vol.probe_counts[1], vol.probe_counts[2]};
    const float m1[3] = {(float)(n[0] > 1 ? n[0] - 1 : 1), (float)(n[1] > 1 ? n[1] - 1 : 1), (float)(n[2] > 1 ? n[2] - 1 : 1)};   // DDGI_PROBES_SIZE_M_1
    f2 oct_norm = oct_encode_01(normal);                                                     // :30
    // probe uv (oct * SIZE + 1) / (SIZE + 2) of a (SIZE + 2)-texel tile, in texels of that tile (:31-32)
    float icx = oct_norm.x * (float)IS + 1.0f, icy = oct_norm.y * (float)IS + 1.0f;
    float vcx = oct_norm.x * (float)VS + 1.0f, vcy = oct_norm.y * (float)VS + 1.0f;
    float idx_f[3] = {vol.extent[0] > 0.0f ? x * m1[0] / vol.extent[0] : 0.0f, vol.extent[1] > 0.0f ? y * m1[1] / vol.extent[1] : 0.0f,
                      vol.extent[2] > 0.0f ? z * m1[2] / vol.extent[2] : 0.0f};              // :34-38
    uint32_t idx[3];
    for (int a = 0; a < 3; a++) {                                                            // :39-40
        uint32_t cap = n[a] > 1 ? n[a] - 2 : 0;
        idx[a] = std::min(to_uint(idx_f[a]), cap);
        idx_f[a] = idx_f[a] - (float)idx[a];
    }
    const size_t istride = (size_t)n[0] * n[1] * (IS + 2), vstride = (size_t)n[0] * n[1] * (VS + 2);
    f3 sum = splat3(0.0f);
    float sum_weight = 0.0f;
    for (uint32_t i = 0; i < 8; i++) {                                                       // :44-79
        uint32_t d[3] = {i & 1u, (i >> 1) & 1u, i >> 2};
        float w_probe = (lerpf(1.0f - idx_f[0], idx_f[0], (float)d[0]) * lerpf(1.0f - idx_f[1], idx_f[1], (float)d[1])) * lerpf(1.0f - idx_f[2], idx_f[2], (float)d[2]);
        uint32_t pi[3];
        for (int a = 0; a < 3; a++) pi[a] = std::min(idx[a] + d[a], n[a] - 1);
        f3 probe_center = ((base + ((float)pi[0] * vol.extent[0] / m1[0]) * fx) + ((float)pi[1] * vol.extent[1] / m1[1]) * fy) + ((float)pi[2] * vol.extent[2] / m1[2]) * fz;
        f3 to_probe = probe_center - pos;
        f3 dir = normalize(to_probe);
        float w_dir = pow2((dot(dir, normal) + 1.0f) * 0.5f) + 0.2f;                         // :57
        size_t tile = (size_t)pi[1] * n[0] + pi[0];                                          // probe_start.x (:59)
        float visibility[2], irradiance[4];
        atlas_bilinear<2>(ctx.ddgi_visibility.data(), vstride, tile * (VS + 2), (size_t)pi[2] * (VS + 2), vcx, vcy, visibility);
        float sigma2 = visibility[1] - pow2(visibility[0]);                                  // :66
        float dist = sqrtf(dot(to_probe, to_probe));
        float w_vis = sigma2 / (sigma2 + pow2(fmax_(dist - visibility[0], 0.0f)));           // :68
        atlas_bilinear<4>(ctx.ddgi_irradiance.data(), istride, tile * (IS + 2), (size_t)pi[2] * (IS + 2), icx, icy, irradiance);
        float w = (w_probe * w_dir) * ((w_vis * w_vis) * w_vis);                             // :75
        sum = sum + mk3(irradiance[0], irradiance[1], irradiance[2]) * w;
        sum_weight += w;
    }
    f4 r = sum_weight == 0.0f ? f4{0, 0, 0, 1} : f4{sum.x / sum_weight, sum.y / sum_weight, sum.z / sum_weight, 1.0f};   // :81
    if (!finite3(mk3(r.x, r.y, r.z))) r = f4{0, 0, 0, 0};                                    // :82
    return r;
}

struct ThreadOut {
    TraceStats ext, shd;
    uint64_t ext_per_bounce[16] = {0}, shd_per_bounce[16] = {0};
    uint64_t shaded = 0, missed = 0, pixel_samples = 0, ext_wide_rays = 0, shd_wide_rays = 0;
    struct CapE { uint32_t bounce, pixel; bpt_hit hit; };
    struct CapS { uint32_t bounce, pixel, light; };
    std::vector<CapE> cap_e;
    std::vector<f3> pending;
    std::vector<CapS> cap_s;
};

// One path from (O, D): the loop of SURVEY Appendix A. `pixel` keys the RNG stream (py * W + px for camera
// paths, probe * rays + ray for probe paths); `first_t` (optional) receives the first hit distance or -1.
template <class Acc>
static void trace_path(const obpt_context& ctx, const bpt_settings& st, bool diffuse_only, uint32_t frame_index, uint32_t pixel,
                       f3 O, f3 D, Acc* a, float* first_t, ThreadOut& out, uint32_t fp16_n = 0, const f3* W0 = nullptr, bool ibl = false) {
    const Scene& sc = ctx.scene;
    const uint32_t B = std::min(std::max(st.max_bounces, 2u), 16u);               // path_tracing.cpp:290
    const bool fp16 = st.state_precision == BPT_STATE_REFERENCE_FP16;
    f3 Wt = W0 ? *W0 : splat3(1.0f);              // (ray-traced reflections start with the specular sample's weight)
    // Per-sample colour C_s starts at 0, receives this sample's contributions in order, and is added to the
    // FP32 sum buffer when the sample ends (the reference's color texture + accumulate pass,
    // path_tracing.cpp:421-480): sum += C_s, samples in ascending frame order.
    Acc C[3] = {0, 0, 0};
    // reference_fp16: C is the rgba16_sfloat colour texture; a bounce's light terms are summed in FP32 (bsum), multiplied
    // by the throughput, stored as halves (deferred_lighting_secondary.hlsl:110) and written (bounce 1) or blended
    // additively with a half result (path_tracing.cpp:441-459). The sample is folded into the image by the running
    // lerp of pt_accumulate.hlsl:9 with weight 1/n, n = `fp16_n` (path_tracing.cpp:473).
    struct Flush {
        Acc* a; Acc* C; uint32_t n;
        ~Flush() {
            if (n == 0) { a[0] += C[0]; a[1] += C[1]; a[2] += C[2]; return; }
            for (int k = 0; k < 3; k++) a[k] = n == 1 ? C[k] : (Acc)store_half(lerpf((float)a[k], (float)C[k], 1.0f / (float)n));
        }
    } flush{a, C, fp16_n};
    f3 bsum = splat3(0.0f);
    auto add = [&](f3 c) { if (fp16) bsum = bsum + c; else { C[0] += c.x; C[1] += c.y; C[2] += c.z; } };
    auto commit = [&](uint32_t i) {
        if (!fp16) return;
        f3 c = store_half3(bsum * Wt);
        f3 prev = mk3((float)C[0], (float)C[1], (float)C[2]);
        f3 r = i == 1 ? c : store_half3(prev + c);
        C[0] = r.x; C[1] = r.y; C[2] = r.z;
        bsum = splat3(0.0f);
    };
    for (uint32_t i = 1; i < B; i++) {
        const f3 Wl = fp16 ? splat3(1.0f) : Wt;  // fp16: the light terms are summed unweighted, `commit` applies the throughput
        out.ext_per_bounce[i]++;
        const bool wide = sc.wide_from_bounce != 0 && i >= sc.wide_from_bounce;           // which tree the CUDA kernels walk at this bounce
        if (wide) { out.ext_wide_rays++; }
        HitRec h = trace_closest(sc, O, D, 0.001f, st.ray_length, frame_index, out.ext, wide);   // rt_gbuffer.hlsl:17-25
        if (ctx.capture) out.cap_e.push_back({i, pixel, bpt_hit{h.t, h.u, h.v, h.hit ? h.instance_id : 0xffffffffu, h.hit ? h.prim : 0xffffffffu}});
        if (i == 1 && first_t) *first_t = h.hit ? h.t : -1.0f;
        if (!h.hit) {                                                                       // deferred_lighting_secondary.hlsl:24-29
            const float* m = sc.sky_transform;
            f3 dir = mk3((m[0] * D.x + m[1] * D.y) + m[2] * D.z, (m[3] * D.x + m[4] * D.y) + m[5] * D.z, (m[6] * D.x + m[7] * D.y) + m[8] * D.z);
            f3 color = sample_sky(sc, dir) * mk3(sc.sky_color[0], sc.sky_color[1], sc.sky_color[2]);
            add(fp16 ? color : color * Wt);
            commit(i);
            out.missed++;
            return;
        }
        out.shaded++;
        const InstanceXf& x = sc.xf[h.inst_slot];
        const bpt_drawable_sbt_data& dr = sc.drawables[x.instance_id];
        const bpt_material& mat = sc.materials[dr.material_offset / sizeof(bpt_material)];
        uint32_t surface_model = (mat.flags >> BPT_MATERIAL_MODEL_SHIFT) & 0xffu;
        f3 P = O + D * h.t;                                                                 // rt_gbuffer.hlsl:32
        Vertex vt = fetch_vertex_attributes(sc, x, h.prim, h.u, h.v);
        SurfaceData surf = material_function(sc, mat, vt.texcoord, vt.position_world, vt.color);
        f3 nts = surf.normal_map_value * 2.0f - splat3(1.0f);                               // rt_gbuffer_hit.hlsl:10-14
        f3 N = normalize((nts.x * vt.tangent_world + nts.y * vt.bitangent_world) + nts.z * vt.normal_world);
        if (surf.two_sided && dot(D, N) > 0.0f) N = -N;
        f3 T;
        if (fp16) {                                                                         // rt_gbuffer_hit.hlsl:15, rt_gbuffer.hlsl:27-31
            GBuffer g = store_gbuffer(pack_surface_to_gbuffer(N, vt.tangent_world, surf, surface_model));
            unpack_gbuffer_to_surface(g, N, T, surf, surface_model);                        // deferred_lighting_secondary.hlsl:37-40
        } else {
            T = gbuffer_roundtrip_tangent(N, vt.tangent_world);                             // gbuffer.hlsl:27,41
        }
        f3 Bv = cross(N, T);                                                                // deferred_lighting_secondary.hlsl:41
        surf.opacity = 1.0f;                                                                // gbuffer.hlsl:44
        if (diffuse_only) {                                                                 // ddgi/deferred_lighting.hlsl:44-45
            surf = surface_data_diffuse(surf.base_color);
            surface_model = 1u;
        }
        f3 V = normalize(O - P);                                                            // deferred_lighting_secondary.hlsl:45

        // Contributions of this vertex. The GPU adds the unshadowed (immediate) terms in the shade
        // kernel and the shadow-ray terms in the connect kernel that follows it, so the oracle keeps
        // that order: rect lights (LTC, unshadowed) first, then the NEE terms in light order.
        out.pending.clear();
        auto direct = [&](f3 le, f3 L, float tmax, uint32_t light_index) {
            f3 c = (le * surface_eval(N, T, Bv, V, L, surf, surface_model)) * Wl;
            if (!(fmax_(c.x, fmax_(c.y, c.z)) > 0.0f)) return;            // zero contribution: no ray
            if (st.nee_mode == BPT_NEE_NONE) { out.pending.push_back(c); return; }
            out.shd_per_bounce[i]++;
            if (ctx.capture) out.cap_s.push_back({i, pixel, light_index});
            if (wide) out.shd_wide_rays++;
            if (!trace_any(sc, P, L, 0.001f, tmax, frame_index, out.shd, false, wide)) out.pending.push_back(c);
        };
        for (size_t l = 0; l < sc.rect_lights.size(); l++) {                                // deferred_lighting_secondary.hlsl:72-96
            const bpt_rect_light_data& rl = sc.rect_lights[l];
            f3 mrp = splat3(0.0f);
            f3 c = ltc_rect_light(sc, rl, P, N, T, Bv, V, surf, surface_model, st.rect_shadow ? &mrp : nullptr) * Wl;
            if (!st.rect_shadow) { add(c); continue; }                                      // reference: rect lights are unshadowed
            // NEW switch rect_shadow = mrp_ray: one shadow ray towards the most representative point of the diffuse lobe;
            // distance to the light's plane as in rect_light_sample_texture (lights.hlsl:425-438)
            if (!(fmax_(c.x, fmax_(c.y, c.z)) > 0.0f)) continue;
            f3 ln = mk3(rl.normal[0], rl.normal[1], rl.normal[2]);
            float step = fabsf(dot(mrp, ln));
            if (!(step >= 0.0001f)) { add(c); continue; }
            float dist = fabsf(dot(P - mk3(rl.position2[0], rl.position2[1], rl.position2[2]), ln));
            out.shd_per_bounce[i]++;
            uint32_t light_index = (uint32_t)(sc.dir_lights.size() + sc.point_lights.size() + l);
            if (ctx.capture) out.cap_s.push_back({i, pixel, light_index});
            if (wide) out.shd_wide_rays++;
            if (!trace_any(sc, P, mrp, 0.001f, (dist / step) * 0.999f, frame_index, out.shd, false, wide)) out.pending.push_back(c);
        }
        if (ibl && sc.ibl_valid) add(ibl_lighting(sc, N, V, surf, surface_model) * Wl);       // deferred_lighting_secondary.hlsl:98-108 (RTR only)
        // Probe paths: the previous DDGI update at the path's last vertex (ddgi/deferred_lighting.hlsl:102-115). The reference
        // traces one bounce, so every probe-ray hit receives it; with more bounces it closes the path.
        if (diffuse_only && ctx.ddgi_enabled && i + 1 >= B) {
            f4 g = calc_ddgi_volume_lighting(ctx, P, N, V);
            if (g.w > 0.0f) add(((mk3(g.x / g.w, g.y / g.w, g.z / g.w) * surf.base_color) * INV_PI) * Wl);
        }
        for (size_t l = 0; l < sc.dir_lights.size(); l++) {                                 // deferred_lighting_secondary.hlsl:51-60
            const bpt_dir_light_data& li = sc.dir_lights[l];
            direct(mk3(li.emission[0], li.emission[1], li.emission[2]), mk3(li.direction[0], li.direction[1], li.direction[2]), st.ray_length, (uint32_t)l);
        }
        for (size_t l = 0; l < sc.point_lights.size(); l++) {                               // deferred_lighting_secondary.hlsl:61-70
            f3 L; float dist;
            f3 le = point_light_eval(sc.point_lights[l], P, L, dist);
            direct(le, L, dist * 0.999f, (uint32_t)(sc.dir_lights.size() + l));
        }
        for (const f3& c : out.pending) add(c);
        commit(i);

        // next direction: sample_secondary_ray.hlsl:11-69 (bounce_index = i)
        if (i + 1 >= B) return;
        if (fmax_(Wt.x, fmax_(Wt.y, Wt.z)) < 0.001f) return;                                // :23-28
        Frame frame = create_frame(N, T);                                                   // :42
        f3 V_local = frame_to_local(frame, V);
        float rx, ry;
        get_anisotropic_roughness(surf.roughness, surf.anisotropy, rx, ry);
        uint32_t seed = rng_tea(pixel, frame_index + i * 3u);                               // :52 (pixel = py * W + px)
        float u1 = rng_next(seed);
        float u2 = rng_next(seed);
        f3 half_dir = ggx_vndf_sample(V_local, rx, ry, u1, u2);
        f3 out_local = reflect(-V_local, half_dir);
        float pdf_wh = ggx_vndf_sample_pdf(half_dir, V_local, rx, ry);
        float pdf = pdf_wh / (4.0f * fabsf(dot(half_dir, V_local)));
        f3 out_dir = frame_to_world(frame, out_local);
        f3 bsdf = surface_eval(N, T, Bv, V, out_dir, surf, surface_model);
        f3 weight = bsdf / pdf;
        if (!finite3(weight)) weight = splat3(0.0f);                                        // :62-64
        f3 newW = weight * Wt;
        if (fp16) { newW = store_half3(newW); out_dir = store_half3(out_dir); }             // rgba16_sfloat ray_weights / ray_directions (:66-68)
        // A zero-weight path can never contribute again (deferred_lighting_secondary.hlsl:17-21
        // writes 0 and the next sample pass kills it), so it is dropped here instead of traced.
        if (newW.x == 0.0f && newW.y == 0.0f && newW.z == 0.0f) return;
        if (st.russian_roulette && i >= 2) {   // NEW switch: third draw of this bounce's stream, survive with q
            float q = clampf(fmax_(newW.x, fmax_(newW.y, newW.z)), 0.05f, 1.0f);
            float u3 = rng_next(seed);
            if (!(u3 < q)) return;
            newW = newW / q;
            if (fp16) newW = store_half3(newW);
        }
        O = P; D = out_dir; Wt = newW;
    }
}

template <class Acc>
static void render_pixel(const obpt_context& ctx, const bpt_camera& cam, const bpt_settings& st, uint32_t frame_index,
                         uint32_t px, uint32_t py, Acc* a, ThreadOut& out, uint32_t fp16_n) {
    f3 O, D;
    camera_ray(cam, px, py, ctx.width, ctx.height, st.pixel_jitter, frame_index, O, D);                                  // generate_camera_ray.hlsl:4-16
    if (st.state_precision == BPT_STATE_REFERENCE_FP16) D = store_half3(D);                                              // ray_directions is rgba16_sfloat (path_tracing.cpp:276-280)
    trace_path(ctx, st, false, frame_index, py * ctx.width + px, O, D, a, (float*)nullptr, out, fp16_n);
}

template <class Acc>
static void render_impl(obpt_context& ctx, const bpt_camera& cam, uint32_t first, uint32_t ns, const bpt_settings& st, Acc* accum, bool count, uint32_t fp16_done = 0) {
    const bool fp16_lerp = st.state_precision == BPT_STATE_REFERENCE_FP16 && std::is_same<Acc, float>::value;   // (the double-precision "converged" path sums)
    const uint32_t W = ctx.width, H = ctx.height;
    uint32_t nt = ctx.threads ? ctx.threads : std::max(1u, std::thread::hardware_concurrency());
    const uint32_t TILE = 16;
    const uint32_t tx = (W + TILE - 1) / TILE, ty = (H + TILE - 1) / TILE;
    std::atomic<uint32_t> next{0};
    std::vector<ThreadOut> outs(nt);
    auto work = [&](uint32_t tid) {
        ThreadOut& out = outs[tid];
        for (;;) {
            uint32_t t = next.fetch_add(1);
            if (t >= tx * ty) break;
            if (t % ctx.tile_stride != ctx.tile_offset) continue;
            uint32_t x0 = (t % tx) * TILE, y0 = (t / tx) * TILE;
            for (uint32_t y = y0; y < std::min(y0 + TILE, H); y++)
                for (uint32_t x = x0; x < std::min(x0 + TILE, W); x++)
                    for (uint32_t s = 0; s < ns; s++, out.pixel_samples++)
                        render_pixel(ctx, cam, st, first + s, x, y, accum + 4ull * (y * W + x), out, fp16_lerp ? fp16_done + s + 1 : 0u);
        }
    };
    std::vector<std::thread> th;
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    if (!count) return;
    TraceStats ext, shd;
    for (auto& o : outs) {
        ext.add(o.ext); shd.add(o.shd);
        for (int b = 0; b < 16; b++) {
            ctx.counters.extend_rays_per_bounce[b] += o.ext_per_bounce[b];
            ctx.counters.shadow_rays_per_bounce[b] += o.shd_per_bounce[b];
        }
        ctx.stats.shaded_vertices += o.shaded;
        ctx.stats.miss_vertices += o.missed;
        ctx.counters.samples += o.pixel_samples;
        ctx.stats.samples += o.pixel_samples;
    }
    ctx.counters.extend_rays += ext.rays; ctx.counters.shadow_rays += shd.rays;
    ctx.stats.extend_rays += ext.rays; ctx.stats.extend_nodes += ext.nodes; ctx.stats.extend_tris += ext.tris; ctx.stats.extend_instances += ext.instances;
    ctx.stats.shadow_rays += shd.rays; ctx.stats.shadow_nodes += shd.nodes; ctx.stats.shadow_tris += shd.tris; ctx.stats.shadow_instances += shd.instances;
    ctx.stats.extend_wide_nodes += ext.wide_nodes; ctx.stats.extend_leaf_boxes += ext.leaf_boxes;
    ctx.stats.shadow_wide_nodes += shd.wide_nodes; ctx.stats.shadow_leaf_boxes += shd.leaf_boxes;
    for (auto& o : outs) { ctx.stats.extend_wide_rays += o.ext_wide_rays; ctx.stats.shadow_wide_rays += o.shd_wide_rays; }
    if (ctx.capture) {
        uint32_t B = std::min(std::max(st.max_bounces, 2u), 16u);
        ctx.cap_extend_pixels.assign(B, {}); ctx.cap_extend_hits.assign(B, {});
        ctx.cap_shadow_pixels.assign(B, {}); ctx.cap_shadow_lights.assign(B, {});
        std::vector<ThreadOut::CapE> e; std::vector<ThreadOut::CapS> s;
        for (auto& o : outs) { e.insert(e.end(), o.cap_e.begin(), o.cap_e.end()); s.insert(s.end(), o.cap_s.begin(), o.cap_s.end()); }
        std::sort(e.begin(), e.end(), [](auto& a, auto& b) { return a.bounce != b.bounce ? a.bounce < b.bounce : a.pixel < b.pixel; });
        std::sort(s.begin(), s.end(), [](auto& a, auto& b) { return a.bounce != b.bounce ? a.bounce < b.bounce : (a.pixel != b.pixel ? a.pixel < b.pixel : a.light < b.light); });
        for (auto& r : e) { ctx.cap_extend_pixels[r.bounce].push_back(r.pixel); ctx.cap_extend_hits[r.bounce].push_back(r.hit); }
        for (auto& r : s) { ctx.cap_shadow_pixels[r.bounce].push_back(r.pixel); ctx.cap_shadow_lights[r.bounce].push_back(r.light); }
    }
}

} // namespace orc

using namespace orc;

#define CHECK_CTX(c) do { if (!(c)) return BPT_ERR_INVALID; } while (0)
static bpt_status fail(obpt_context* c, bpt_status s, const char* msg) { c->err = msg; return s; }

extern "C" {

bpt_status obpt_create(const bpt_config* cfg, obpt_context** out) {
    if (!cfg || !out || cfg->width == 0 || cfg->height == 0) return BPT_ERR_INVALID;
    auto* c = new obpt_context();
    c->width = cfg->width; c->height = cfg->height;
    c->accum.assign((size_t)cfg->width * cfg->height * 4, 0.0f);
    *out = c;
    return BPT_OK;
}
bpt_status obpt_destroy(obpt_context* c) { if (c) obpt_reblur_free(c); delete c; return BPT_OK; }
const char* obpt_last_error(const obpt_context* c) { return c ? c->err.c_str() : "null context"; }
bpt_status obpt_set_threads(obpt_context* c, uint32_t n) { CHECK_CTX(c); c->threads = n; return BPT_OK; }
uint32_t obpt_get_threads(const obpt_context* c) { return c->threads ? c->threads : std::max(1u, std::thread::hardware_concurrency()); }
bpt_status obpt_resize(obpt_context* c, uint32_t w, uint32_t h) {
    CHECK_CTX(c); if (!w || !h) return BPT_ERR_INVALID;
    c->width = w; c->height = h; c->accum.assign((size_t)w * h * 4, 0.0f);
    c->accum_used = false; c->accum_fp16 = false; c->accum_count = 0;
    return BPT_OK;
}

bpt_status obpt_scene_upload_geometry(obpt_context* c, const bpt_geometry_streams* s, const bpt_drawable_sbt_data* dr, const uint32_t* va,
                                      uint32_t nd, const bpt_blas_desc* blas, uint32_t nb) {
    CHECK_CTX(c);
    if (!s || !dr || !blas || !s->positions || !s->indices) return fail(c, BPT_ERR_INVALID, "geometry: null stream");
    Scene& sc = c->scene;
    auto cp = [](std::vector<float>& v, const float* p, uint64_t n) { v.assign(p ? p : nullptr, p ? p + n : nullptr); };
    cp(sc.positions, s->positions, s->num_position_floats); cp(sc.normals, s->normals, s->num_normal_floats);
    cp(sc.tangents, s->tangents, s->num_tangent_floats); cp(sc.colors, s->colors, s->num_color_floats);
    cp(sc.texcoords, s->texcoords, s->num_texcoord_floats); cp(sc.texcoords2, s->texcoords2, s->num_texcoord2_floats);
    sc.indices.assign(s->indices, s->indices + s->num_indices);
    sc.drawables.assign(dr, dr + nd);
    sc.drawable_va.resize(nd);
    for (uint32_t i = 0; i < nd; i++) {
        uint32_t m = va ? va[i] : (BPT_VA_POSITION | BPT_VA_NORMAL | BPT_VA_TANGENT | BPT_VA_TEXCOORD | (s->colors ? (uint32_t)BPT_VA_COLOR : 0u));
        if (!s->colors) m &= ~BPT_VA_COLOR;
        if (!s->normals) m &= ~BPT_VA_NORMAL;
        if (!s->tangents) m &= ~BPT_VA_TANGENT;
        if (!s->texcoords) m &= ~BPT_VA_TEXCOORD;
        sc.drawable_va[i] = m;
    }
    sc.blas_descs.assign(blas, blas + nb);
    sc.accel_built = false;
    return BPT_OK;
}
bpt_status obpt_scene_upload_instances(obpt_context* c, const bpt_instance_desc* inst, uint32_t n) {
    CHECK_CTX(c); if (!inst && n) return BPT_ERR_INVALID;
    c->scene.instances.assign(inst, inst + n);
    return BPT_OK;
}
bpt_status obpt_scene_upload_materials(obpt_context* c, const bpt_material* m, uint32_t n, const bpt_texture_desc* t, uint32_t nt) {
    CHECK_CTX(c); if (!m || !n) return fail(c, BPT_ERR_INVALID, "materials: empty");
    c->scene.materials.assign(m, m + n);
    c->scene.textures.clear();
    for (uint32_t i = 0; i < nt; i++) {
        Texture tx; tx.w = t[i].width; tx.h = t[i].height; tx.format = t[i].format; tx.addr_u = t[i].address_mode_u; tx.addr_v = t[i].address_mode_v; tx.linear = t[i].filter_linear;
        if (tx.format == BPT_TEXTURE_RGBA8_SRGB) {      // decode once to linear FP32 texels (filtering happens after the decode)
            float lut[256];
            for (int k = 0; k < 256; k++) { double v = k / 255.0; lut[k] = (float)(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4)); }
            const uint8_t* src = (const uint8_t*)t[i].texels;
            std::vector<float> lin((size_t)tx.w * tx.h * 4);
            for (size_t k = 0; k < lin.size(); k += 4) { lin[k] = lut[src[k]]; lin[k + 1] = lut[src[k + 1]]; lin[k + 2] = lut[src[k + 2]]; lin[k + 3] = (float)src[k + 3] / 255.0f; }
            tx.format = BPT_TEXTURE_RGBA32_FLOAT;
            tx.texels.assign((const uint8_t*)lin.data(), (const uint8_t*)lin.data() + lin.size() * 4);
        } else {
            size_t bytes = (size_t)tx.w * tx.h * (tx.format == BPT_TEXTURE_RGBA8_UNORM ? 4 : 16);
            tx.texels.assign((const uint8_t*)t[i].texels, (const uint8_t*)t[i].texels + bytes);
        }
        c->scene.textures.push_back(std::move(tx));
    }
    return BPT_OK;
}
bpt_status obpt_scene_upload_lights(obpt_context* c, const bpt_dir_light_data* d, uint32_t nd, const bpt_point_light_data* p, uint32_t np,
                                    const bpt_rect_light_data* r, uint32_t nr, const bpt_ltc_luts* luts) {
    CHECK_CTX(c);
    Scene& sc = c->scene;
    sc.dir_lights.assign(d, d + nd); sc.point_lights.assign(p, p + np); sc.rect_lights.assign(r, r + nr);
    if (nr) {
        if (!luts || !luts->matrix_lut0 || !luts->matrix_lut1 || !luts->matrix_lut2 || !luts->norm_lut) return fail(c, BPT_ERR_INVALID, "rect lights need the LTC LUTs");
        sc.ltc_m0.assign(luts->matrix_lut0, luts->matrix_lut0 + 8 * 8 * 64 * 4);
        sc.ltc_m1.assign(luts->matrix_lut1, luts->matrix_lut1 + 8 * 8 * 64 * 4);
        sc.ltc_m2.assign(luts->matrix_lut2, luts->matrix_lut2 + 8 * 8 * 64 * 4);
        sc.ltc_norm.assign(luts->norm_lut, luts->norm_lut + 8 * 8 * 64 * 2);
    }
    return BPT_OK;
}
// Rect-light textures. Level 0 is decoded to what a sampler returns; the chain follows shaders/core/mipmap.hlsl:46-93 as
// CommandHelpers::generate_mipmaps_2d dispatches it (command_helpers.cpp:66-215: tex_size = the DESTINATION extent, so the "odd" taps are
// taken when the destination is odd; reads past the source level return 0), every level stored in the texture's format.
namespace {
double srgb_to_linear(double v) { return v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4); }
float store_in_format(float x, uint32_t format, int channel) {
    if (format == BPT_TEXTURE_RGBA32_FLOAT) return x;
    if (format == BPT_TEXTURE_RGBA8_UNORM || channel == 3) return store_unorm(x, 8);
    if (!(x > 0.0f)) return 0.0f;
    // nearest 8-bit sRGB code in the encoded domain = the code c with EOTF((c - 0.5)/255) <= x < EOTF((c + 0.5)/255); then decoded again
    int code = 0;
    for (int c = 1; c < 256; c++) if ((float)srgb_to_linear((c - 0.5) / 255.0) <= x) code = c; else break;
    return (float)srgb_to_linear(code / 255.0);
}
}
bpt_status obpt_scene_upload_light_textures(obpt_context* c, const bpt_light_texture_desc* t, uint32_t nt) {
    CHECK_CTX(c);
    if ((nt && !t) || nt > BPT_MAX_RECT_LIGHT_TEXTURES) return fail(c, BPT_ERR_INVALID, "light textures: null array or more than 16");
    std::vector<LightTexture> out(nt);
    for (uint32_t i = 0; i < nt; i++) {
        const bpt_light_texture_desc& d = t[i];
        if (!d.texels || !d.width || !d.height || d.width > 16384 || d.height > 16384 || d.format > BPT_TEXTURE_RGBA8_SRGB || d.levels == 0 ||
            d.address_mode_u > BPT_ADDRESS_CLAMP || d.address_mode_v > BPT_ADDRESS_CLAMP) return fail(c, BPT_ERR_INVALID, "bad light texture desc");
        LightTexture& lt = out[i];
        lt.w = d.width; lt.h = d.height; lt.addr_u = d.address_mode_u; lt.addr_v = d.address_mode_v; lt.linear = d.filter_linear != 0; lt.mip_linear = d.mip_linear != 0;
        uint32_t levels = 1;
        while ((std::max(d.width, d.height) >> levels) != 0) levels++;
        levels = std::min(levels, d.levels);
        lt.level.resize(levels);
        std::vector<float>& l0 = lt.level[0];
        l0.resize((size_t)d.width * d.height * 4);
        if (d.format == BPT_TEXTURE_RGBA32_FLOAT) std::memcpy(l0.data(), d.texels, l0.size() * 4);
        else {
            const uint8_t* src = static_cast<const uint8_t*>(d.texels);
            for (size_t k = 0; k < l0.size(); k++)
                l0[k] = (d.format == BPT_TEXTURE_RGBA8_SRGB && (k & 3) != 3) ? (float)srgb_to_linear(src[k] / 255.0) : (float)src[k] / 255.0f;
        }
        for (uint32_t l = 1; l < levels; l++) {
            const uint32_t sw = std::max(d.width >> (l - 1), 1u), sh = std::max(d.height >> (l - 1), 1u), dw = std::max(sw / 2, 1u), dh = std::max(sh / 2, 1u);
            const std::vector<float>& src = lt.level[l - 1];
            std::vector<float>& dst = lt.level[l];
            dst.resize((size_t)dw * dh * 4);
            const bool odd_x = (dw & 1u) != 0, odd_y = (dh & 1u) != 0;
            for (uint32_t y = 0; y < dh; y++)
                for (uint32_t x = 0; x < dw; x++)
                    for (int ch = 0; ch < 4; ch++) {
                        auto at = [&](uint32_t xx, uint32_t yy) { return (xx < sw && yy < sh) ? src[((size_t)yy * sw + xx) * 4 + ch] : 0.0f; };
                        const uint32_t sx = 2 * x, sy = 2 * y;
                        float r = (at(sx, sy) + at(sx, sy + 1)) + (at(sx + 1, sy) + at(sx + 1, sy + 1));
                        uint32_t num = 4;
                        if (odd_x) { r = r + (at(sx + 2, sy) + at(sx + 2, sy + 1)); num += 2; }
                        if (odd_y) { r = r + (at(sx, sy + 2) + at(sx + 1, sy + 2)); num += 2; }
                        if (odd_x && odd_y) { r = r + at(sx + 2, sy + 2); num += 1; }
                        dst[((size_t)y * dw + x) * 4 + ch] = store_in_format(r / (float)num, d.format, ch);
                    }
        }
    }
    c->scene.light_textures = std::move(out);
    return BPT_OK;
}
bpt_status obpt_debug_read_light_texture(obpt_context* c, uint32_t index, float* out, uint64_t cap, uint64_t* out_texels) {
    CHECK_CTX(c);
    if (index >= c->scene.light_textures.size()) return fail(c, BPT_ERR_INVALID, "light texture index out of range");
    uint64_t total = 0;
    for (auto& l : c->scene.light_textures[index].level) total += l.size() / 4;
    if (out_texels) *out_texels = total;
    if (!out) return BPT_OK;
    if (cap < total) return fail(c, BPT_ERR_INVALID, "capacity too small");
    for (auto& l : c->scene.light_textures[index].level) { std::memcpy(out, l.data(), l.size() * 4); out += l.size(); }
    return BPT_OK;
}

bpt_status obpt_scene_upload_sky(obpt_context* c, const float* faces, uint32_t size, const float xf[9], const float col[3]) {
    CHECK_CTX(c);
    Scene& sc = c->scene;
    sc.ibl_valid = false;
    if (faces && size) { sc.sky_faces.assign(faces, faces + (size_t)6 * size * size * 4); sc.sky_size = size; }
    else { sc.sky_faces.clear(); sc.sky_size = 0; }
    if (xf) std::memcpy(sc.sky_transform, xf, sizeof(float) * 9);
    if (col) std::memcpy(sc.sky_color, col, sizeof(float) * 3);
    return BPT_OK;
}

bpt_status obpt_scene_update_sky_params(obpt_context* c, const float xf[9], const float col[3]) {
    CHECK_CTX(c);
    if (xf) std::memcpy(c->scene.sky_transform, xf, sizeof(float) * 9);
    if (col) std::memcpy(c->scene.sky_color, col, sizeof(float) * 3);
    return BPT_OK;
}

bpt_status obpt_build_accel(obpt_context* c, uint32_t mode) {
    CHECK_CTX(c);
    if (c->scene.materials.empty()) return fail(c, BPT_ERR_STATE, "upload materials before build_accel");
    for (auto& d : c->scene.drawables)
        if (d.material_offset % sizeof(bpt_material) || d.material_offset / sizeof(bpt_material) >= c->scene.materials.size())
            return fail(c, BPT_ERR_INVALID, "drawable material_offset out of range");
    std::string err;
    if (!build_accel(c->scene, mode, err)) { c->err = err; return BPT_ERR_INVALID; }
    return BPT_OK;
}
bpt_status obpt_update_tlas(obpt_context* c) {
    CHECK_CTX(c);
    if (!c->scene.accel_built || c->scene.accel_mode != BPT_ACCEL_TWO_LEVEL) return fail(c, BPT_ERR_STATE, "update_tlas needs a built two-level accel");
    std::string err;
    if (!build_tlas(c->scene, err)) { c->err = err; return BPT_ERR_INVALID; }
    return BPT_OK;
}
bpt_status obpt_debug_read_bvh(obpt_context* c, uint32_t which, uint32_t* np, uint64_t* morton, uint32_t* prims, bpt_bvh_node* nodes, uint32_t cap, int32_t* root) {
    CHECK_CTX(c);
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "accel not built");
    const Bvh* b;
    if (which == BPT_BVH_TLAS) { if (c->scene.accel_mode != BPT_ACCEL_TWO_LEVEL) return fail(c, BPT_ERR_INVALID, "no TLAS in merged mode"); b = &c->scene.tlas; }
    else { if (which >= c->scene.blas.size()) return fail(c, BPT_ERR_INVALID, "blas index out of range"); b = &c->scene.blas[which]; }
    if (np) *np = b->n;
    if (root) *root = b->root;
    if ((morton || prims || nodes) && cap < b->n) return fail(c, BPT_ERR_INVALID, "capacity too small");
    if (morton) std::copy(b->morton.begin(), b->morton.end(), morton);
    if (prims) std::copy(b->prims.begin(), b->prims.end(), prims);
    if (nodes) std::copy(b->nodes.begin(), b->nodes.end(), nodes);
    return BPT_OK;
}

bpt_status obpt_clear_accum(obpt_context* c) {
    CHECK_CTX(c);
    std::fill(c->accum.begin(), c->accum.end(), 0.0f);
    c->accum_used = false; c->accum_fp16 = false; c->accum_count = 0;
    return BPT_OK;
}
bpt_status obpt_render(obpt_context* c, const bpt_camera* cam, uint32_t first, uint32_t ns, const bpt_settings* st) {
    CHECK_CTX(c); if (!cam || !st) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "render before build_accel");
    if (st->state_precision != BPT_STATE_FP32 && st->state_precision != BPT_STATE_REFERENCE_FP16) return fail(c, BPT_ERR_INVALID, "state_precision: unknown value");
    if ((st->nee_mode != BPT_NEE_SHADOW_RAY && st->nee_mode != BPT_NEE_NONE) || st->rect_shadow > 1 || st->russian_roulette > 1 || st->pixel_jitter > 1)
        return fail(c, BPT_ERR_INVALID, "settings: unknown value of a mode switch (nee_mode, rect_shadow, russian_roulette, pixel_jitter)");
    const bool fp16 = st->state_precision == BPT_STATE_REFERENCE_FP16;
    if (c->accum_used && fp16 != c->accum_fp16) return fail(c, BPT_ERR_STATE, "state_precision changed without bpt_clear_accum");
    c->accum_used = true; c->accum_fp16 = fp16;
    render_impl<float>(*c, *cam, first, ns, *st, c->accum.data(), true, c->accum_count);
    if (fp16) c->accum_count += ns;
    return BPT_OK;
}
bpt_status obpt_render_converged(obpt_context* c, const bpt_camera* cam, uint32_t first, uint32_t ns, const bpt_settings* st, float* out) {
    CHECK_CTX(c); if (!cam || !st || !out || !ns) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "render before build_accel");
    std::vector<double> acc((size_t)c->width * c->height * 4, 0.0);
    bool cap = c->capture; c->capture = false;
    render_impl<double>(*c, *cam, first, ns, *st, acc.data(), false);
    c->capture = cap;
    for (size_t p = 0; p < (size_t)c->width * c->height; p++) {
        for (int k = 0; k < 3; k++) out[p * 4 + k] = (float)(acc[p * 4 + k] / (double)ns);
        out[p * 4 + 3] = 1.0f;
    }
    return BPT_OK;
}
bpt_status obpt_resolve(obpt_context* c, uint32_t total, float* out) {
    CHECK_CTX(c); if (!out || !total) return BPT_ERR_INVALID;
    float inv = c->accum_fp16 ? 1.0f : 1.0f / (float)total;       // reference_fp16: the buffer already is the running average
    for (size_t p = 0; p < (size_t)c->width * c->height; p++) {
        for (int k = 0; k < 3; k++) out[p * 4 + k] = c->accum[p * 4 + k] * inv;
        out[p * 4 + 3] = 1.0f;
    }
    return BPT_OK;
}
bpt_status obpt_get_counters(obpt_context* c, bpt_counters* o) { CHECK_CTX(c); *o = c->counters; return BPT_OK; }
bpt_status obpt_reset_counters(obpt_context* c) { CHECK_CTX(c); c->counters = bpt_counters{}; c->stats = obpt_stats{}; return BPT_OK; }
bpt_status obpt_get_stats(obpt_context* c, obpt_stats* o) { CHECK_CTX(c); *o = c->stats; return BPT_OK; }
bpt_status obpt_set_tile_sample(obpt_context* c, uint32_t stride, uint32_t offset) {
    CHECK_CTX(c); if (!stride || offset >= stride) return BPT_ERR_INVALID;
    c->tile_stride = stride; c->tile_offset = offset; return BPT_OK;
}

bpt_status obpt_trace_rays(obpt_context* c, const bpt_ray* rays, uint64_t n, uint32_t frame, bpt_hit* out) {
    CHECK_CTX(c); if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "accel not built");
    uint32_t nt = obpt_get_threads(c);
    std::vector<std::thread> th; std::vector<TraceStats> sts(nt);
    auto work = [&](uint32_t tid) {
        for (uint64_t i = tid; i < n; i += nt) {
            const bpt_ray& r = rays[i];
            HitRec h = trace_closest(c->scene, mk3(r.origin[0], r.origin[1], r.origin[2]), mk3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, frame, sts[tid]);
            out[i] = bpt_hit{h.t, h.u, h.v, h.hit ? h.instance_id : 0xffffffffu, h.hit ? h.prim : 0xffffffffu};
        }
    };
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    for (auto& s : sts) { c->stats.extend_rays += s.rays; c->stats.extend_nodes += s.nodes; c->stats.extend_tris += s.tris; c->stats.extend_instances += s.instances; }
    return BPT_OK;
}
bpt_status obpt_trace_shadow_rays(obpt_context* c, const bpt_ray* rays, uint64_t n, uint32_t frame, uint8_t* vis) {
    CHECK_CTX(c); if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "accel not built");
    uint32_t nt = obpt_get_threads(c);
    std::vector<std::thread> th; std::vector<TraceStats> sts(nt);
    auto work = [&](uint32_t tid) {
        for (uint64_t i = tid; i < n; i += nt) {
            const bpt_ray& r = rays[i];
            vis[i] = trace_any(c->scene, mk3(r.origin[0], r.origin[1], r.origin[2]), mk3(r.direction[0], r.direction[1], r.direction[2]), r.tmin, r.tmax, frame, sts[tid]) ? 0 : 1;
        }
    };
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    for (auto& s : sts) { c->stats.shadow_rays += s.rays; c->stats.shadow_nodes += s.nodes; c->stats.shadow_tris += s.tris; c->stats.shadow_instances += s.instances; }
    return BPT_OK;
}
bpt_status obpt_debug_capture(obpt_context* c, uint32_t en) { CHECK_CTX(c); c->capture = en != 0; return BPT_OK; }
bpt_status obpt_debug_read_queue(obpt_context* c, uint32_t bounce, uint32_t kind, uint32_t* pixels, uint32_t* lights, bpt_hit* hits, uint64_t cap, uint64_t* count) {
    CHECK_CTX(c);
    auto& px = kind == 0 ? c->cap_extend_pixels : c->cap_shadow_pixels;
    if (bounce >= px.size()) { if (count) *count = 0; return BPT_OK; }
    uint64_t n = px[bounce].size();
    if (count) *count = n;
    if (!pixels && !lights && !hits) return BPT_OK;
    if (cap < n) return fail(c, BPT_ERR_INVALID, "capacity too small");
    if (pixels) std::copy(px[bounce].begin(), px[bounce].end(), pixels);
    if (kind == 0 && hits) std::copy(c->cap_extend_hits[bounce].begin(), c->cap_extend_hits[bounce].end(), hits);
    if (kind == 1 && lights) std::copy(c->cap_shadow_lights[bounce].begin(), c->cap_shadow_lights[bounce].end(), lights);
    return BPT_OK;
}
bpt_status obpt_set_wide_from_bounce(obpt_context* c, uint32_t bounce) {
    CHECK_CTX(c);
    if (bounce && (!c->scene.accel_built || c->scene.accel_mode != BPT_ACCEL_MERGED)) return fail(c, BPT_ERR_STATE, "set_wide_from_bounce needs a built merged accel");
    if (bounce && c->scene.blas[0].wide.empty()) build_wide4(c->scene.blas[0]);
    c->scene.wide_from_bounce = bounce;
    return BPT_OK;
}
bpt_status obpt_set_ddgi_volume(obpt_context* c, const bpt_probe_volume* vol, const bpt_probe_blend* bl, const float* irr, const float* vis) {
    CHECK_CTX(c);
    if (!vol || !bl || !irr || !vis) { c->ddgi_enabled = false; return BPT_OK; }
    const uint64_t nx = vol->probe_counts[0], ny = vol->probe_counts[1], nz = vol->probe_counts[2];
    if (!nx || !ny || !nz || bl->irradiance_size < 2 || bl->visibility_size < 2 || bl->irradiance_size > 30 || bl->visibility_size > 30)
        return fail(c, BPT_ERR_INVALID, "set_ddgi_volume: bad sizes");
    c->ddgi_irradiance.assign(irr, irr + nx * ny * (bl->irradiance_size + 2) * nz * (bl->irradiance_size + 2) * 4);
    c->ddgi_visibility.assign(vis, vis + nx * ny * (bl->visibility_size + 2) * nz * (bl->visibility_size + 2) * 2);
    c->ddgi_volume = *vol; c->ddgi_irr_size = bl->irradiance_size; c->ddgi_vis_size = bl->visibility_size; c->ddgi_enabled = true;
    return BPT_OK;
}
bpt_status obpt_ddgi_lighting(obpt_context* c, uint64_t n, const float* pos, const float* normal, const float* view, float* out) {
    CHECK_CTX(c); if (!pos || !normal || !view || !out) return BPT_ERR_INVALID;
    if (!c->ddgi_enabled) return fail(c, BPT_ERR_STATE, "ddgi_lighting: no volume bound (obpt_set_ddgi_volume)");
    for (uint64_t i = 0; i < n; i++) {
        f4 r = calc_ddgi_volume_lighting(*c, mk3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), mk3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]),
                                         mk3(view[3 * i], view[3 * i + 1], view[3 * i + 2]));
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
    return BPT_OK;
}

// Primary-hit outputs of the pass: the first trace pass's G-buffer (rt_gbuffer_hit.hlsl:6-18 packed by gbuffer.hlsl:18-33 into the
// formats of pass/gbuffer.hpp:14-17) and the depth pass (pt_depth.hlsl:7-16). Texels of missing rays: 0 (the reference leaves them untouched).
bpt_status obpt_render_primary(obpt_context* c, const bpt_camera* cam, uint32_t frame_index, const bpt_settings* st, float* out_depth, bpt_gbuffer_texel* out_g) {
    CHECK_CTX(c); if (!cam || !st) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "render_primary before build_accel");
    const Scene& sc = c->scene;
    const uint32_t W = c->width, H = c->height;
    uint32_t nt = obpt_get_threads(c);
    std::vector<std::thread> th; std::vector<TraceStats> sts(nt);
    auto work = [&](uint32_t tid) {
        for (uint32_t p = tid; p < W * H; p += nt) {
            f3 O, D;
            camera_ray(*cam, p % W, p / W, W, H, st->pixel_jitter, frame_index, O, D);
            if (st->state_precision == BPT_STATE_REFERENCE_FP16) D = store_half3(D);
            HitRec h = trace_closest(sc, O, D, 0.001f, st->ray_length, frame_index, sts[tid]);
            float depth = 0.0f;                                                          // DEVICE_Z_FARTHEST
            bpt_gbuffer_texel g{};
            if (h.hit) {
                const InstanceXf& x = sc.xf[h.inst_slot];
                const bpt_material& mat = sc.materials[sc.drawables[x.instance_id].material_offset / sizeof(bpt_material)];
                uint32_t surface_model = (mat.flags >> BPT_MATERIAL_MODEL_SHIFT) & 0xffu;
                f3 P = O + D * h.t;
                Vertex vt = fetch_vertex_attributes(sc, x, h.prim, h.u, h.v);
                SurfaceData surf = material_function(sc, mat, vt.texcoord, vt.position_world, vt.color);
                f3 nts = surf.normal_map_value * 2.0f - splat3(1.0f);
                f3 N = normalize((nts.x * vt.tangent_world + nts.y * vt.bitangent_world) + nts.z * vt.normal_world);
                if (surf.two_sided && dot(D, N) > 0.0f) N = -N;
                GBuffer gb = store_gbuffer(pack_surface_to_gbuffer(N, vt.tangent_world, surf, surface_model));
                std::memcpy(g.base_color, &gb.base_color, 16); std::memcpy(g.normal_roughness, &gb.normal_roughness, 16);
                std::memcpy(g.fresnel, &gb.fresnel, 16); std::memcpy(g.material_0, &gb.material_0, 16);
                const float* m = cam->matrix_proj_view;                                  // pt_depth.hlsl:13-15
                float z = ((m[2] * P.x + m[6] * P.y) + m[10] * P.z) + m[14];
                float w = ((m[3] * P.x + m[7] * P.y) + m[11] * P.z) + m[15];
                depth = z / w;
            }
            if (out_depth) out_depth[p] = depth;
            if (out_g) out_g[p] = g;
        }
    };
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    for (auto& s : sts) { c->stats.extend_rays += s.rays; c->stats.extend_nodes += s.nodes; c->stats.extend_tris += s.tris; c->counters.extend_rays += s.rays; c->counters.extend_rays_per_bounce[1] += s.rays; }
    return BPT_OK;
}

// ambient_occlusion_rt.hlsl:14-66 (AmbientOcclusionPass::render_raytraced, ambient_occlusion.cpp:217-262).
bpt_status obpt_trace_ao(obpt_context* c, const bpt_camera* cam, uint32_t frame_index, const bpt_ao_settings* ao, const float* depth_img, const float* nr_img, float* out) {
    CHECK_CTX(c); if (!cam || !ao || !depth_img || !nr_img || !out) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "trace_ao before build_accel");
    const Scene& sc = c->scene;
    const uint32_t W = c->width, H = c->height;
    if (ao->half_resolution && ((W | H) & 1u)) return fail(c, BPT_ERR_UNSUPPORTED, "trace_ao: half resolution needs even width and height (texel-centre reads)");
    const uint32_t aw = ao->half_resolution ? W / 2 : W, ah = ao->half_resolution ? H / 2 : H;   // ambient_occlusion.cpp:217-218
    const float range = ao->range > 0.05f ? ao->range : 0.05f;                                   // :241
    uint32_t nt = obpt_get_threads(c);
    std::vector<std::thread> th; std::vector<TraceStats> sts(nt);
    auto work = [&](uint32_t tid) {
        for (uint32_t p = tid; p < aw * ah; p += nt) {
            uint32_t px = p % aw, py = p / aw;
            float sx = 0.5f, sy = 0.5f;                                                          // :18-26
            uint32_t tx = px, ty = py;
            if (ao->half_resolution) {
                sx = (frame_index & 1u) ? 0.75f : 0.25f; sy = (frame_index & 2u) ? 0.75f : 0.25f;
                tx = std::min(2u * px + ((frame_index & 1u) ? 1u : 0u), W - 1u); ty = std::min(2u * py + ((frame_index & 2u) ? 1u : 0u), H - 1u);
            }
            float uvx = ((float)px + sx) / (float)aw, uvy = ((float)py + sy) / (float)ah;
            float depth = depth_img[(size_t)ty * W + tx];                                        // linear sampler at a texel centre = that texel
            if (depth == 0.0f) { out[2 * p] = 1.0f; out[2 * p + 1] = 0.0f; continue; }           // :29-32
            const float* q = nr_img + 4 * ((size_t)ty * W + tx);
            f3 normal, tangent;
            unpack_normal_and_tangent(mk3(q[0], q[1], q[2]), normal, tangent);                   // :35-38
            Frame frame = create_frame(normal, tangent);
            const float* ip = cam->matrix_inv_proj; const float* iv = cam->matrix_inv_view;      // projection.hlsl:5-10
            float nx = uvx * 2.0f - 1.0f, ny = 1.0f - uvy * 2.0f;
            float vx = ((ip[0] * nx + ip[4] * ny) + ip[8] * depth) + ip[12];
            float vy = ((ip[1] * nx + ip[5] * ny) + ip[9] * depth) + ip[13];
            float vz = ((ip[2] * nx + ip[6] * ny) + ip[10] * depth) + ip[14];
            float vw = ((ip[3] * nx + ip[7] * ny) + ip[11] * depth) + ip[15];
            vx = vx / vw; vy = vy / vw; vz = vz / vw;
            f3 Pw = mk3(((iv[0] * vx + iv[4] * vy) + iv[8] * vz) + iv[12], ((iv[1] * vx + iv[5] * vy) + iv[9] * vz) + iv[13],
                        ((iv[2] * vx + iv[6] * vy) + iv[10] * vz) + iv[14]);                     // :41-42
            f3 origin = Pw + normal * 0.001f;                                                    // :56
            uint32_t seed = rng_tea(py * aw + px, frame_index);                                  // :44
            uint32_t occluded = 0;
            for (int i = 0; i < 4; i++) {                                                        // :47-63
                float r0 = rng_next(seed);
                float r1 = rng_next(seed);
                f3 dir = frame_to_world(frame, cos_hemisphere_sample(r0, r1));
                if (trace_any(sc, origin, dir, 0.001f, range, frame_index, sts[tid], true)) occluded++;
            }
            out[2 * p] = store_half(1.0f - ((float)occluded * ao->strength) / 4.0f);             // :64, rg16_sfloat target (ambient_occlusion.cpp:14)
            out[2 * p + 1] = 1.0f;
        }
    };
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    for (auto& s : sts) { c->stats.shadow_rays += s.rays; c->stats.shadow_nodes += s.nodes; c->stats.shadow_tris += s.tris; c->counters.shadow_rays += s.rays; c->counters.shadow_rays_per_bounce[1] += s.rays; }
    return BPT_OK;
}

// ---- SkyboxPrecomputePass::render (src/renderer/pass/skybox_precompute.cpp:66-162): the three precompute shaders ----
namespace {
float radical_inverse_vdc(uint32_t bits) {                               // core/utils/low_discrepancy.hlsl:3-10
    bits = (bits << 16u) | (bits >> 16u);
    bits = ((bits & 0x55555555u) << 1u) | ((bits & 0xAAAAAAAAu) >> 1u);
    bits = ((bits & 0x33333333u) << 2u) | ((bits & 0xCCCCCCCCu) >> 2u);
    bits = ((bits & 0x0F0F0F0Fu) << 4u) | ((bits & 0xF0F0F0F0u) >> 4u);
    bits = ((bits & 0x00FF00FFu) << 8u) | ((bits & 0xFF00FF00u) >> 8u);
    return (float)bits * 2.3283064365386963e-10f;
}
f3 cubemap_direction_from_layered_uv(float u, float v, uint32_t layer) { // core/utils/cubemap.hlsl:3-21
    u = u * 2.0f - 1.0f; v = v * 2.0f - 1.0f;
    f3 d;
    if (layer == 0) d = mk3(1.0f, -v, -u);
    else if (layer == 1) d = mk3(-1.0f, -v, u);
    else if (layer == 2) d = mk3(u, 1.0f, v);
    else if (layer == 3) d = mk3(u, -1.0f, -v);
    else if (layer == 4) d = mk3(u, -v, 1.0f);
    else d = mk3(-u, -v, -1.0f);
    return normalize(d);
}
float luminance(f3 c) { return (c.x * 0.212671f + c.y * 0.715160f) + c.z * 0.072169f; }   // core/utils/color.hlsl:3-5
float store_unorm8(float v) { float c = v > 0.0f ? (v < 1.0f ? v : 1.0f) : 0.0f; return rintf(c * 255.0f) / 255.0f; }
const float INV_TWO_PI = 0.15915494309189533577f;      // sin / cos of an angle a are sincos_2pi(a / 2pi): the numeric contract's fixed-order form
const uint32_t IBL_SAMPLES = 1024;                     // ibl_brdf_lut.hlsl:6, skybox_precompute_specular.hlsl:9
const float CLAMP_LUM = 12.0f;                         // skybox_precompute_diffuse.hlsl:10, skybox_precompute_specular.hlsl:10
} // namespace

bpt_status obpt_precompute_sky_ibl(obpt_context* c, const bpt_sky_ibl_desc* d) {
    CHECK_CTX(c); if (!d) return BPT_ERR_INVALID;
    if (!d->diffuse_size || !d->specular_size || !d->brdf_lut_size || d->specular_levels < 2 || d->specular_levels > 12 ||
        (d->specular_size >> (d->specular_levels - 1)) == 0 || d->diffuse_size > 4096 || d->specular_size > 4096 || d->brdf_lut_size > 4096)
        return fail(c, BPT_ERR_INVALID, "sky ibl: sizes out of range (2 <= levels <= 12, last mip >= 1 texel)");
    Scene& sc = c->scene;
    sc.ibl_valid = false;
    const float* sky = sc.sky_faces.data(); const uint32_t sky_size = sc.sky_size;
    const uint32_t R = d->brdf_lut_size, DS = d->diffuse_size;
    sc.ibl_brdf.assign((size_t)R * R * 2, 0.0f);
    sc.ibl_diffuse.assign((size_t)6 * DS * DS * 4, 0.0f);
    size_t spec_floats = 0;
    for (uint32_t l = 0; l < d->specular_levels; l++) { size_t n = d->specular_size >> l; spec_floats += 6 * n * n * 4; }
    sc.ibl_specular.assign(spec_floats, 0.0f);
    uint32_t nt = obpt_get_threads(c);
    auto parallel = [&](size_t n, auto&& fn) {
        std::vector<std::thread> th;
        auto work = [&](uint32_t tid) { for (size_t i = tid; i < n; i += nt) fn(i); };
        for (uint32_t t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto& t : th) t.join();
    };
    // "IBL BRDF LUT": ibl_brdf_lut.hlsl:8-32, rg8_unorm (skybox.cpp:25-29)
    parallel((size_t)R * R, [&](size_t i) {
        uint32_t x = (uint32_t)(i % R), y = (uint32_t)(i / R);
        const float inv = 1.0f / (float)R;                                                       // skybox_precompute.cpp:87
        float ndotv = ((float)x + 0.5f) * inv, roughness = ((float)y + 0.5f) * inv;
        f3 wi = mk3(sqrtf(1.0f - ndotv * ndotv), 0.0f, ndotv);
        float vx = 0.0f, vy = 0.0f;
        for (uint32_t k = 0; k < IBL_SAMPLES; k++) {
            f3 wh = ggx_vndf_sample(wi, roughness, roughness, (float)k / (float)IBL_SAMPLES, radical_inverse_vdc(k));
            f3 wo = reflect(-wi, wh);
            if (wo.z <= 0.0f) continue;
            float weight = ggx_g1(wo, roughness, roughness);                                     // ggx_vndf_sample_weight_sep, utils.hlsl:123-125
            float hdotv = dot(wi, wh);
            float f = pow5(1.0f - hdotv);
            vx = vx + (1.0f - f) * weight; vy = vy + f * weight;
        }
        sc.ibl_brdf[2 * i] = store_unorm8(vx / (float)IBL_SAMPLES); sc.ibl_brdf[2 * i + 1] = store_unorm8(vy / (float)IBL_SAMPLES);
    });
    // "Skybox Precompute Diffuse": skybox_precompute_diffuse.hlsl:12-41, rgba16_sfloat cube
    parallel((size_t)6 * DS * DS, [&](size_t i) {
        uint32_t layer = (uint32_t)(i / ((size_t)DS * DS)), r = (uint32_t)(i % ((size_t)DS * DS)), x = r % DS, y = r / DS;
        const float inv = 1.0f / (float)DS;
        f3 dir = cubemap_direction_from_layered_uv(((float)x + 0.5f) * inv, ((float)y + 0.5f) * inv, layer);
        Frame frame = create_frame(dir);
        f3 irradiance = splat3(0.0f);
        const float delta = PI / 64.0f;                                                          // num_samples_sqrt = 64
        for (float phi = delta * 0.5f; phi < TWO_PI; phi += delta) {
            float sin_phi, cos_phi;
            sincos_2pi(phi * INV_TWO_PI, sin_phi, cos_phi);
            for (float theta = delta * 0.5f; theta < 0.5f * PI; theta += delta) {
                float sin_theta, cos_theta;
                sincos_2pi(theta * INV_TWO_PI, sin_theta, cos_theta);
                f3 v = frame_to_world(frame, mk3(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta));
                f3 color = sample_cube(sky, sky_size, v);
                float scale = CLAMP_LUM / fmax_(luminance(color), CLAMP_LUM);
                irradiance = irradiance + ((color * scale) * cos_theta) * sin_theta;
            }
        }
        f3 o = store_half3((irradiance * PI) / 4096.0f);
        sc.ibl_diffuse[4 * i] = o.x; sc.ibl_diffuse[4 * i + 1] = o.y; sc.ibl_diffuse[4 * i + 2] = o.z; sc.ibl_diffuse[4 * i + 3] = 1.0f;
    });
    // "Skybox Precompute Specular i": skybox_precompute_specular.hlsl:12-41, one pass per mip, roughness = i / (levels - 1)
    size_t offset = 0;
    for (uint32_t l = 0; l < d->specular_levels; l++) {
        const uint32_t S = d->specular_size >> l;
        const float roughness = (float)l / (float)(d->specular_levels - 1);                      // skybox_precompute.cpp:148
        float* out = sc.ibl_specular.data() + offset;
        parallel((size_t)6 * S * S, [&](size_t i) {
            uint32_t layer = (uint32_t)(i / ((size_t)S * S)), r = (uint32_t)(i % ((size_t)S * S)), x = r % S, y = r / S;
            const float inv = 1.0f / (float)S;
            f3 dir = cubemap_direction_from_layered_uv(((float)x + 0.5f) * inv, ((float)y + 0.5f) * inv, layer);
            Frame frame = create_frame(dir);
            const f3 wi = mk3(0.0f, 0.0f, 1.0f);
            f3 filtered = splat3(0.0f);
            float weight_sum = 0.0f;
            for (uint32_t k = 0; k < IBL_SAMPLES; k++) {
                f3 wh = ggx_vndf_sample(wi, roughness, roughness, (float)k / (float)IBL_SAMPLES, radical_inverse_vdc(k));
                f3 wo = reflect(-wi, wh);
                if (wo.z <= 0.0f) continue;
                f3 color = sample_cube(sky, sky_size, frame_to_world(frame, wo));
                float scale = CLAMP_LUM / fmax_(luminance(color), CLAMP_LUM);
                filtered = filtered + (color * scale) * wo.z;
                weight_sum = weight_sum + wo.z;
            }
            f3 o = store_half3(filtered / weight_sum);
            out[4 * i] = o.x; out[4 * i + 1] = o.y; out[4 * i + 2] = o.z; out[4 * i + 3] = 1.0f;
        });
        offset += (size_t)6 * S * S * 4;
    }
    sc.ibl_desc = *d; sc.ibl_valid = true;
    return BPT_OK;
}
bpt_status obpt_debug_read_sky_ibl(obpt_context* c, float* diffuse, float* specular, float* brdf) {
    CHECK_CTX(c);
    if (!c->scene.ibl_valid) return fail(c, BPT_ERR_STATE, "sky ibl not computed");
    if (diffuse) std::memcpy(diffuse, c->scene.ibl_diffuse.data(), c->scene.ibl_diffuse.size() * 4);
    if (specular) std::memcpy(specular, c->scene.ibl_specular.data(), c->scene.ibl_specular.size() * 4);
    if (brdf) std::memcpy(brdf, c->scene.ibl_brdf.data(), c->scene.ibl_brdf.size() * 4);
    return BPT_OK;
}

// Ray-traced reflections: ReflectionPass::render_raytraced (reflection.cpp:317-450).
//   specular_sample_cs              direction_sample/specular_sample.hlsl:14-83
//   trace + lighting                rt_gbuffer.hlsl:7-36 (ray_length = range) and deferred_lighting_secondary.hlsl:11-111
//                                   (lighting_strength = strength) = one bounce of trace_path starting with the sample's weight;
//                                   the IBL block (:98-108) is not evaluated (as with DEFERRED_LIGHTING_NO_IBL), see bpt.h.
bpt_status obpt_trace_reflection(obpt_context* c, const bpt_camera* cam, uint32_t frame_index, const bpt_reflection_settings* rs, const float* depth_img,
                                 const bpt_gbuffer_texel* gb, float* out_refl, float* out_hit) {
    CHECK_CTX(c); if (!cam || !rs || !depth_img || !gb || !out_refl || !out_hit) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "trace_reflection before build_accel");
    const uint32_t W = c->width, H = c->height;
    if (rs->half_resolution && ((W | H) & 1u)) return fail(c, BPT_ERR_UNSUPPORTED, "trace_reflection: half resolution needs even width and height (texel-centre reads)");
    if (rs->ibl > 1 || (rs->ibl && !c->scene.ibl_valid)) return fail(c, rs->ibl > 1 ? BPT_ERR_INVALID : BPT_ERR_STATE, "trace_reflection: settings.ibl needs obpt_precompute_sky_ibl");
    const uint32_t rw = rs->half_resolution ? (W + 1) / 2 : W, rh = rs->half_resolution ? (H + 1) / 2 : H;      // reflection.cpp:324-325
    const float max_roughness = rs->max_roughness;
    const float fade_roughness = std::min(rs->fade_roughness, max_roughness - 0.0001f);                         // reflection.cpp:361-362
    bpt_settings st{};
    st.ray_length = rs->range; st.max_bounces = 2; st.nee_mode = BPT_NEE_SHADOW_RAY;
    uint32_t nt = obpt_get_threads(c);
    std::vector<ThreadOut> outs(nt);
    bool cap = c->capture; c->capture = false;
    auto work = [&](uint32_t tid) {
        for (uint32_t p = tid; p < rw * rh; p += nt) {
            float* refl = out_refl + 4 * (size_t)p; float* hitp = out_hit + 4 * (size_t)p;
            refl[0] = refl[1] = refl[2] = 0.0f; refl[3] = 1.0f;
            hitp[0] = hitp[1] = hitp[2] = 0.0f; hitp[3] = -1.0f;
            uint32_t px = p % rw, py = p / rw;
            float sx = 0.5f, sy = 0.5f;                                                          // specular_sample.hlsl:18-26
            uint32_t tx = px, ty = py;
            if (rs->half_resolution) {
                sx = (frame_index & 1u) ? 0.75f : 0.25f; sy = (frame_index & 2u) ? 0.75f : 0.25f;
                tx = std::min(2u * px + ((frame_index & 1u) ? 1u : 0u), W - 1u); ty = std::min(2u * py + ((frame_index & 2u) ? 1u : 0u), H - 1u);
            }
            float uvx = ((float)px + sx) / (float)rw, uvy = ((float)py + sy) / (float)rh;
            float depth = depth_img[(size_t)ty * W + tx];                                        // a sampler at a texel centre = that texel
            if (depth == 0.0f) continue;                                                         // :28-33 is_depth_background
            const bpt_gbuffer_texel& t = gb[(size_t)ty * W + tx];
            GBuffer g;
            g.base_color = f4{t.base_color[0], t.base_color[1], t.base_color[2], t.base_color[3]};
            g.normal_roughness = f4{t.normal_roughness[0], t.normal_roughness[1], t.normal_roughness[2], t.normal_roughness[3]};
            g.fresnel = f4{t.fresnel[0], t.fresnel[1], t.fresnel[2], t.fresnel[3]};
            g.material_0 = f4{t.material_0[0], t.material_0[1], t.material_0[2], t.material_0[3]};
            f3 N, T; SurfaceData surface; uint32_t surface_model;
            unpack_gbuffer_to_surface(g, N, T, surface, surface_model);                          // :41-44
            if (surface.roughness > max_roughness) continue;                                     // :46-50
            f3 Bv = cross(N, T);
            Frame frame = create_frame(N, T);
            const float* ip = cam->matrix_inv_proj; const float* iv = cam->matrix_inv_view;      // projection.hlsl:5-10
            float nx = uvx * 2.0f - 1.0f, ny = 1.0f - uvy * 2.0f;
            float vx = ((ip[0] * nx + ip[4] * ny) + ip[8] * depth) + ip[12];
            float vy = ((ip[1] * nx + ip[5] * ny) + ip[9] * depth) + ip[13];
            float vz = ((ip[2] * nx + ip[6] * ny) + ip[10] * depth) + ip[14];
            float vw = ((ip[3] * nx + ip[7] * ny) + ip[11] * depth) + ip[15];
            vx = vx / vw; vy = vy / vw; vz = vz / vw;
            f3 Pw = mk3(((iv[0] * vx + iv[4] * vy) + iv[8] * vz) + iv[12], ((iv[1] * vx + iv[5] * vy) + iv[9] * vz) + iv[13],
                        ((iv[2] * vx + iv[6] * vy) + iv[10] * vz) + iv[14]);                     // :56
            f3 V = normalize(mk3(iv[12], iv[13], iv[14]) - Pw);                                  // :57, camera.hlsl:7-9
            f3 V_local = frame_to_local(frame, V);
            float rx, ry;
            get_anisotropic_roughness(surface.roughness, surface.anisotropy, rx, ry);
            uint32_t seed = rng_tea(py * rw + px, frame_index);                                  // :63
            float u1 = rng_next(seed);
            float u2 = rng_next(seed);
            f3 half_dir = ggx_vndf_sample(V_local, rx, ry, u1, u2);
            f3 out_local = reflect(-V_local, half_dir);
            float pdf_wh = ggx_vndf_sample_pdf(half_dir, V_local, rx, ry);
            float pdf = pdf_wh / (4.0f * fabsf(dot(half_dir, V_local)));
            f3 out_dir = frame_to_world(frame, out_local);
            f3 spec = surface_eval_specular(N, T, Bv, V, out_dir, surface, surface_model);       // :70-72
            float fade = 1.0f - fmax_(surface.roughness - fade_roughness, 0.0f) / fmax_(max_roughness - fade_roughness, 0.0001f);
            f3 weight = (spec * fade) / pdf;
            if (!finite3(weight)) weight = splat3(0.0f);                                         // :76-78
            // rt_gbuffer.hlsl:13-15 skips a zero direction (it cannot happen for a sampled pixel: out_dir is a unit vector)
            f3 W0 = weight * rs->strength;                                                       // deferred_lighting_secondary.hlsl:17
            float rgb[3] = {0, 0, 0}, first_t = -1.0f;
            trace_path<float>(*c, st, false, frame_index, p, Pw, out_dir, rgb, &first_t, outs[tid], 0, &W0, rs->ibl != 0);
            refl[0] = rgb[0]; refl[1] = rgb[1]; refl[2] = rgb[2];
            if (first_t >= 0.0f) { f3 hp = Pw + out_dir * first_t; hitp[0] = hp.x; hitp[1] = hp.y; hitp[2] = hp.z; hitp[3] = first_t; }     // rt_gbuffer.hlsl:32
            else { hitp[0] = out_dir.x; hitp[1] = out_dir.y; hitp[2] = out_dir.z; hitp[3] = -1.0f; }                                          // :34
        }
    };
    std::vector<std::thread> th;
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    c->capture = cap;
    for (auto& o : outs) {
        c->counters.extend_rays += o.ext.rays; c->counters.shadow_rays += o.shd.rays;
        for (int b = 0; b < 16; b++) { c->counters.extend_rays_per_bounce[b] += o.ext_per_bounce[b]; c->counters.shadow_rays_per_bounce[b] += o.shd_per_bounce[b]; }
        c->stats.extend_rays += o.ext.rays; c->stats.extend_nodes += o.ext.nodes; c->stats.extend_tris += o.ext.tris;
        c->stats.shadow_rays += o.shd.rays; c->stats.shadow_nodes += o.shd.nodes; c->stats.shadow_tris += o.shd.tris;
    }
    return BPT_OK;
}

// "RTR Upscale Hit" / "RTR Upscale Color": simple_upscale_cs (shaders/renderer/simple_upscale.hlsl:47-97; reflection.cpp:452-530). The group-shared
// tile of the shader is only a cache: tap (dx, dy) of pixel (x, y) is half-res texel clamp((x / 2 + dx, y / 2 + dy), 0, (tex_size + 1) / 2) (:25-35).
bpt_status obpt_upscale_half_res(obpt_context* c, const bpt_camera* cam, uint32_t frame_index, const float* depth_img, const float* nr_img, const float* in_half, float* out) {
    CHECK_CTX(c); if (!cam || !depth_img || !nr_img || !in_half || !out) return BPT_ERR_INVALID;
    const int W = (int)c->width, H = (int)c->height, rw = (W + 1) / 2, rh = (H + 1) / 2;
    const float a = cam->matrix_inv_proj[11], b = cam->matrix_inv_proj[15];                      // inv_proj[3].z / .w: HLSL M[3] is row 3 (depth.hlsl:31-35)
    auto linear01 = [&](float d) { return ((1.0f - d) * b) / (a * d + b); };
    auto depth_at = [&](int px, int py) { return (px < W && py < H) ? depth_img[(size_t)py * W + px] : 0.0f; };      // Texture.Load out of range = 0
    auto normal_at = [&](int px, int py) {
        f2 e = (px < W && py < H) ? f2{nr_img[4 * ((size_t)py * W + px)], nr_img[4 * ((size_t)py * W + px) + 1]} : f2{0.0f, 0.0f};
        return oct_decode(e);
    };
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float center_depth = linear01(depth_at(x, y));
            const f3 center_normal = normal_at(x, y);
            const float cx = (x & 1) ? 0.75f : 0.25f, cy = (y & 1) ? 0.75f : 0.25f;             // :66-69
            const float ix = (frame_index & 1u) ? 0.75f : 0.25f, iy = (frame_index & 2u) ? 0.75f : 0.25f;     // :70-73
            const int sx = (int)(frame_index & 1u), sy = (int)((frame_index >> 1) & 1u);        // :32
            float sum[4] = {0, 0, 0, 0}, sum_weight = 0.0f;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    float ox = ((float)dx + ix) - cx, oy = ((float)dy + iy) - cy;               // :79
                    float r = sqrtf(sqrtf(ox * ox + oy * oy));                                  // :80 sqrt(length(offset))
                    float temp = r / 0.2f;
                    float w = exp_neg(-(temp * temp));                                          // gaussian(r, 0.2), :38-41
                    int hx = std::min(std::max(x / 2 + dx, 0), rw), hy = std::min(std::max(y / 2 + dy, 0), rh);
                    float v[4] = {0, 0, 0, 0};
                    if (hx < rw && hy < rh) for (int k = 0; k < 4; k++) v[k] = in_half[4 * ((size_t)hy * rw + hx) + k];
                    float tap_depth = linear01(depth_at(hx * 2 + sx, hy * 2 + sy));
                    f3 tap_normal = normal_at(hx * 2 + sx, hy * 2 + sy);
                    w = w * fmax_(dot(tap_normal, center_normal), 0.0f);                        // :87
                    w = w * fmax_(0.0f, 1.0f - fabsf(tap_depth - center_depth));                // :89
                    for (int k = 0; k < 4; k++) sum[k] = sum[k] + v[k] * w;
                    sum_weight = sum_weight + w;
                }
            float* o = out + 4 * ((size_t)y * W + x);
            for (int k = 0; k < 4; k++) o[k] = sum_weight == 0.0f ? 0.0f : sum[k] / sum_weight; // :96
        }
    return BPT_OK;
}

// DDGI-style probe tracing: ddgi/trace_gbuffer.hlsl:10-51 (probe centre, R2-table direction, TraceRay) +
// ddgi/deferred_lighting.hlsl:12-118 (diffuse-only surface, V = normalize(probe - P)); further bounces continue
// the path through the same trace/shade code (BASELINE configs[4]); previous-frame DDGI feedback is not modelled.
static bpt_status trace_probes_impl(obpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, uint32_t num_bounces, uint64_t first_path,
                                    uint64_t total, float* out);
bpt_status obpt_trace_probes(obpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, uint32_t num_bounces, float* out) {
    CHECK_CTX(c); if (!vol || !table || !out) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "trace_probes before build_accel");
    const uint64_t total = (uint64_t)vol->probe_counts[0] * vol->probe_counts[1] * vol->probe_counts[2] * vol->rays_per_probe;
    if (total == 0 || total > 0xffffffffull) return fail(c, BPT_ERR_INVALID, "trace_probes: bad volume");
    return trace_probes_impl(c, vol, table, frame, num_bounces, 0, total, out);
}
// Probes [first_probe, first_probe + num_probes): the sharding unit of the DDGI update. Keys are the GLOBAL probe / path
// indices, so a range reproduces the same rays of the full call bit for bit.
bpt_status obpt_trace_probes_range(obpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, uint32_t num_bounces, uint32_t first_probe,
                                   uint32_t num_probes, float* out) {
    CHECK_CTX(c); if (!vol || !table || (!out && num_probes) || num_probes == 0xffffffffu) return BPT_ERR_INVALID;
    if (!c->scene.accel_built) return fail(c, BPT_ERR_STATE, "trace_probes before build_accel");
    const uint64_t all_probes = (uint64_t)vol->probe_counts[0] * vol->probe_counts[1] * vol->probe_counts[2];
    if (all_probes == 0 || vol->rays_per_probe == 0 || all_probes * vol->rays_per_probe > 0xffffffffull) return fail(c, BPT_ERR_INVALID, "trace_probes: bad volume");
    if ((uint64_t)first_probe + num_probes > all_probes) return fail(c, BPT_ERR_INVALID, "trace_probes: probe range outside the volume");
    if (num_probes == 0) return BPT_OK;
    return trace_probes_impl(c, vol, table, frame, num_bounces, (uint64_t)first_probe * vol->rays_per_probe, (uint64_t)num_probes * vol->rays_per_probe, out);
}
static bpt_status trace_probes_impl(obpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, uint32_t num_bounces, uint64_t first_path,
                                    uint64_t total, float* out) {
    bpt_settings st{};
    st.ray_length = vol->ray_length; st.max_bounces = std::min(std::max(num_bounces, 1u), 15u) + 1; st.nee_mode = BPT_NEE_SHADOW_RAY;
    uint32_t nt = obpt_get_threads(c);
    std::vector<ThreadOut> outs(nt);
    std::atomic<uint64_t> next{0};
    bool cap = c->capture; c->capture = false;
    auto work = [&](uint32_t tid) {
        const uint64_t CH = 1024;
        for (;;) {
            uint64_t b0 = next.fetch_add(CH);
            if (b0 >= total) break;
            for (uint64_t local = b0; local < std::min(b0 + CH, total); local++) {
                const uint64_t path = first_path + local;                                        // global path id = probe * rays_per_probe + ray
                uint32_t ray_index = (uint32_t)(path % vol->rays_per_probe), lin = (uint32_t)(path / vol->rays_per_probe);
                uint32_t ix = lin % vol->probe_counts[0], iy = (lin / vol->probe_counts[0]) % vol->probe_counts[1], iz = lin / vol->probe_counts[0] / vol->probe_counts[1];
                float mx = (float)(vol->probe_counts[0] > 1 ? vol->probe_counts[0] - 1 : 1), my = (float)(vol->probe_counts[1] > 1 ? vol->probe_counts[1] - 1 : 1),
                      mz = (float)(vol->probe_counts[2] > 1 ? vol->probe_counts[2] - 1 : 1);
                f3 fx = mk3(vol->frame_x[0], vol->frame_x[1], vol->frame_x[2]), fy = mk3(vol->frame_y[0], vol->frame_y[1], vol->frame_y[2]), fz = mk3(vol->frame_z[0], vol->frame_z[1], vol->frame_z[2]);
                f3 O = ((mk3(vol->base_position[0], vol->base_position[1], vol->base_position[2]) + ((float)ix * vol->extent[0] / mx) * fx) + ((float)iy * vol->extent[1] / my) * fy) +
                       ((float)iz * vol->extent[2] / mz) * fz;                                   // trace_gbuffer.hlsl:20-23
                uint32_t seed = rng_tea(lin, frame);                                             // :25
                uint32_t rand_index = ((uint32_t)(rng_next(seed) * 8192.0f) + ray_index) % 8192u;   // :26
                f3 D = uniform_sphere_sample(table[2 * rand_index], table[2 * rand_index + 1]);  // :27-29
                float rgb[3] = {0, 0, 0}, first_t = -1.0f;
                trace_path<float>(*c, st, true, frame, (uint32_t)path, O, D, rgb, &first_t, outs[tid]);
                out[4 * local] = rgb[0]; out[4 * local + 1] = rgb[1]; out[4 * local + 2] = rgb[2]; out[4 * local + 3] = first_t;
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t i = 1; i < nt; i++) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    c->capture = cap;
    for (auto& o : outs) {
        c->counters.extend_rays += o.ext.rays; c->counters.shadow_rays += o.shd.rays;
        for (int b = 0; b < 16; b++) { c->counters.extend_rays_per_bounce[b] += o.ext_per_bounce[b]; c->counters.shadow_rays_per_bounce[b] += o.shd_per_bounce[b]; }
        c->stats.extend_rays += o.ext.rays; c->stats.extend_nodes += o.ext.nodes; c->stats.extend_tris += o.ext.tris;
        c->stats.shadow_rays += o.shd.rays; c->stats.shadow_nodes += o.shd.nodes; c->stats.shadow_tris += o.shd.tris;
    }
    c->counters.samples += total;
    return BPT_OK;
}

// ---- DDGI probe blending: ddgi/probe_blend_irradiance.hlsl:11-80, probe_blend_visibility.hlsl:10-78 ----
namespace {
f3 oct_decode_01(float fx, float fy) {                                   // core/utils/pack.hlsl:94-102
    float x = fx * 2.0f - 1.0f, y = fy * 2.0f - 1.0f;
    f3 n = mk3(x, y, (1.0f - fabsf(x)) - fabsf(y));
    float t = clampf(-n.z, 0.0f, 1.0f);
    n.x = n.x + (n.x >= 0.0f ? -t : t);
    n.y = n.y + (n.y >= 0.0f ? -t : t);
    return normalize(n);
}
// pow(x, 1/5), pow(x, 5), pow(x, 50) in the fixed-order forms of the numeric contract
float root5(float x) {
    if (!(x > 0.0f)) return 0.0f;
    int32_t i = (int32_t)f2u(x);
    float y = u2f((uint32_t)((i - 0x3f800000) / 5 + 0x3f800000));
    for (int k = 0; k < 6; k++) { float y2 = y * y; float y4 = y2 * y2; y = (4.0f * y + x / y4) * 0.2f; }
    return y;
}
float pow5f(float x) { float x2 = x * x; return (x2 * x2) * x; }
float pow50f(float x) { float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8, x32 = x16 * x16; return (x32 * x16) * x2; }
float temporal(float cur, float hist, float alpha) { return pow5f(lerpf(root5(cur), root5(hist), alpha)); }   // gamma = 5
void border_coord(uint32_t cx, uint32_t cy, uint32_t size, uint32_t& bx, uint32_t& by) {                      // probe_blend_common.hlsl:3-26
    bx = cx; by = cy;
    if (cx == 1) { if (cy == 1) { bx = size + 1; by = size + 1; } else if (cy == size) { bx = size + 1; by = 0; } else { bx = 0; by = size + 1 - cy; } }
    else if (cx == size) { if (cy == 1) { bx = 0; by = size + 1; } else if (cy == size) { bx = 0; by = 0; } else { bx = size + 1; by = size + 1 - cy; } }
    else if (cy == 1) { bx = size + 1 - cx; by = 0; }
    else if (cy == size) { bx = size + 1 - cx; by = size + 1; }
}
bool corner_coords(uint32_t cx, uint32_t cy, uint32_t size, uint32_t c[4]) {                                   // probe_blend_common.hlsl:28-50
    if (cx == 1) { if (cy == 1) { c[0] = size; c[1] = 0; c[2] = 0; c[3] = size; return true; } if (cy == size) { c[0] = 0; c[1] = 1; c[2] = size; c[3] = size + 1; return true; } }
    else if (cx == size) { if (cy == 1) { c[0] = 1; c[1] = 0; c[2] = size + 1; c[3] = size; return true; } if (cy == size) { c[0] = size + 1; c[1] = 1; c[2] = 1; c[3] = size + 1; return true; } }
    return false;
}
} // namespace

bpt_status obpt_blend_probes(obpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, const float* rays,
                             const bpt_probe_blend* bl, float* irr, float* vis) {
    CHECK_CTX(c); if (!vol || !table || !rays || !bl || !irr || !vis) return BPT_ERR_INVALID;
    const uint32_t nx = vol->probe_counts[0], ny = vol->probe_counts[1], nz = vol->probe_counts[2], nr = vol->rays_per_probe;
    if (!nx || !ny || !nz || !nr || bl->irradiance_size < 2 || bl->visibility_size < 2) return fail(c, BPT_ERR_INVALID, "blend_probes: bad sizes");
    std::vector<f3> dirs(nr);
    for (uint32_t probe = 0; probe < nx * ny * nz; probe++) {
        uint32_t ix = probe % nx, iy = (probe / nx) % ny, iz = probe / nx / ny;
        float mx = (float)(nx > 1 ? nx - 1 : 1), my = (float)(ny > 1 ? ny - 1 : 1), mz = (float)(nz > 1 ? nz - 1 : 1);
        f3 fx = mk3(vol->frame_x[0], vol->frame_x[1], vol->frame_x[2]), fy = mk3(vol->frame_y[0], vol->frame_y[1], vol->frame_y[2]), fz = mk3(vol->frame_z[0], vol->frame_z[1], vol->frame_z[2]);
        f3 O = ((mk3(vol->base_position[0], vol->base_position[1], vol->base_position[2]) + ((float)ix * vol->extent[0] / mx) * fx) + ((float)iy * vol->extent[1] / my) * fy) +
               ((float)iz * vol->extent[2] / mz) * fz;
        const float* pr = rays + 4ull * probe * nr;
        for (uint32_t r = 0; r < nr; r++) {
            uint32_t seed = rng_tea(probe, frame);
            uint32_t rand_index = ((uint32_t)(rng_next(seed) * 8192.0f) + r) % 8192u;
            f3 D = uniform_sphere_sample(table[2 * rand_index], table[2 * rand_index + 1]);
            float t = pr[4 * r + 3];
            dirs[r] = t < 0.0f ? D : normalize((O + D * t) - O);                                 // probe_blend_irradiance.hlsl:49
        }
        for (int pass = 0; pass < 2; pass++) {
            const bool visp = pass == 1;
            const uint32_t size = visp ? bl->visibility_size : bl->irradiance_size, ch = visp ? 2 : 4;
            const uint32_t stride = nx * ny * (size + 2), sx = (iy * nx + ix) * (size + 2), sy = iz * (size + 2);
            float* atlas = visp ? vis : irr;
            auto at = [&](uint32_t x, uint32_t y) { return atlas + ((size_t)(sy + y) * stride + (sx + x)) * ch; };
            std::vector<f3> vals(size * size);
            for (uint32_t ty = 0; ty < size; ty++)
                for (uint32_t tx = 0; tx < size; tx++) {
                    f3 probe_dir = oct_decode_01(((float)tx + 0.5f) / (float)size, ((float)ty + 0.5f) / (float)size);
                    f3 sum = splat3(0.0f); float wsum = 0.0f;
                    for (uint32_t r = 0; r < nr; r++) {
                        float w = fmax_(dot(probe_dir, dirs[r]), 0.0f);
                        if (visp) {                                                               // probe_blend_visibility.hlsl:45-50
                            w = pow50f(w);
                            float dist = pr[4 * r + 3] < 0.0f ? 1e6f : pr[4 * r + 3];
                            sum.x += w * dist; sum.y += w * (dist * dist);
                        } else sum = sum + w * mk3(pr[4 * r], pr[4 * r + 1], pr[4 * r + 2]);       // probe_blend_irradiance.hlsl:50-52
                        wsum += w;
                    }
                    f3 v = wsum == 0.0f ? splat3(0.0f) : sum / wsum;
                    if (bl->history_valid) {
                        const float* h = at(tx + 1, ty + 1);
                        v.x = temporal(v.x, h[0], bl->alpha); v.y = temporal(v.y, h[1], bl->alpha);
                        if (!visp) v.z = temporal(v.z, h[2], bl->alpha);
                    }
                    vals[ty * size + tx] = v;
                }
            for (uint32_t ty = 0; ty < size; ty++)
                for (uint32_t tx = 0; tx < size; tx++) {
                    f3 v = vals[ty * size + tx];
                    auto put = [&](uint32_t x, uint32_t y) { float* o = at(x, y); o[0] = v.x; o[1] = v.y; if (!visp) { o[2] = v.z; o[3] = 1.0f; } };
                    uint32_t cx = tx + 1, cy = ty + 1, bx, by, cc[4];
                    put(cx, cy);
                    border_coord(cx, cy, size, bx, by); put(bx, by);
                    if (corner_coords(cx, cy, size, cc)) { put(cc[0], cc[1]); put(cc[2], cc[3]); }
                }
        }
    }
    return BPT_OK;
}

uint32_t obpt_rng_tea(uint32_t a, uint32_t b) { return rng_tea(a, b); }
uint32_t obpt_rng_lcg(uint32_t* s) { return rng_lcg(*s); }
void obpt_sincos_2pi(float u, float* s, float* c) { sincos_2pi(u, *s, *c); }
float obpt_atan2(float y, float x) { return atan2_(y, x); }
float obpt_acos(float x) { return acos_(x); }
void obpt_ggx_vndf_sample(const float v[3], float rx, float ry, float u1, float u2, float o[3]) {
    f3 h = ggx_vndf_sample(mk3(v[0], v[1], v[2]), rx, ry, u1, u2); o[0] = h.x; o[1] = h.y; o[2] = h.z;
}
void obpt_surface_eval_lit(const float N[3], const float T[3], const float V[3], const float L[3], const float base[3], const float f0[3], const float f90[3],
                           float roughness, float anisotropy, float out[3]) {
    SurfaceData s = surface_data_default();
    s.base_color = mk3(base[0], base[1], base[2]); s.f0_color = mk3(f0[0], f0[1], f0[2]); s.f90_color = mk3(f90[0], f90[1], f90[2]);
    s.roughness = roughness; s.anisotropy = anisotropy;
    f3 n = mk3(N[0], N[1], N[2]), t = mk3(T[0], T[1], T[2]);
    f3 r = surface_eval(n, t, cross(n, t), mk3(V[0], V[1], V[2]), mk3(L[0], L[1], L[2]), s, 1u);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// ---- unit entry points for the independent float64 pins of tests/test_oracle.py (test hooks, not part of the boundary) ----
// ltc_integrate (lights.hlsl:383-423) of the quad L[4] (row k = L[k]) seen from P in the frame (T, B, N); Minv = nullptr: identity
void obpt_unit_ltc_integrate(const float P[3], const float N[3], const float T[3], const float B[3], const float* Minv, const float L[12], uint32_t two_sided,
                             float* integral, float mrp[3]) {
    m33 I{mk3(1, 0, 0), mk3(0, 1, 0), mk3(0, 0, 1)};
    m33 mi = Minv ? m33{mk3(Minv[0], Minv[1], Minv[2]), mk3(Minv[3], Minv[4], Minv[5]), mk3(Minv[6], Minv[7], Minv[8])} : I;
    f3 Lq[4] = {mk3(L[0], L[1], L[2]), mk3(L[3], L[4], L[5]), mk3(L[6], L[7], L[8]), mk3(L[9], L[10], L[11])};
    f3 m = splat3(0.0f);
    *integral = ltc_integrate(mk3(P[0], P[1], P[2]), mk3(N[0], N[1], N[2]), mk3(T[0], T[1], T[2]), mk3(B[0], B[1], B[2]), mi, Lq, two_sided != 0, &m);
    mrp[0] = m.x; mrp[1] = m.y; mrp[2] = m.z;
}
// rect_light_eval_ltc + surface_eval_lut for one light with the context's LUTs and light textures (deferred_lighting_secondary.hlsl:80-96)
void obpt_unit_rect_light(obpt_context* c, const bpt_rect_light_data* light, const float P[3], const float N[3], const float T[3], const float B[3], const float V[3],
                          const float base[3], const float f0[3], const float f90[3], float roughness, float anisotropy, float out[3], float diff_mrp[3]) {
    SurfaceData s = surface_data_default();
    s.base_color = mk3(base[0], base[1], base[2]); s.f0_color = mk3(f0[0], f0[1], f0[2]); s.f90_color = mk3(f90[0], f90[1], f90[2]);
    s.roughness = roughness; s.anisotropy = anisotropy;
    f3 m = splat3(0.0f);
    f3 r = ltc_rect_light(c->scene, *light, mk3(P[0], P[1], P[2]), mk3(N[0], N[1], N[2]), mk3(T[0], T[1], T[2]), mk3(B[0], B[1], B[2]), mk3(V[0], V[1], V[2]), s, 1u, &m);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    if (diff_mrp) { diff_mrp[0] = m.x; diff_mrp[1] = m.y; diff_mrp[2] = m.z; }
}
void obpt_unit_point_light(const bpt_point_light_data* l, const float P[3], float radiance[3], float dir[3], float* dist) {     // lights.hlsl:14-25
    f3 d; float t;
    f3 e = point_light_eval(*l, mk3(P[0], P[1], P[2]), d, t);
    radiance[0] = e.x; radiance[1] = e.y; radiance[2] = e.z; dir[0] = d.x; dir[1] = d.y; dir[2] = d.z; *dist = t;
}
void obpt_unit_sample_sky(obpt_context* c, const float d[3], float out[3]) {                // skybox.SampleLevel(sampler, dir, 0)
    f3 r = sample_sky(c->scene, mk3(d[0], d[1], d[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// fetch_vertex_attributes (core/raytracing/hit.hlsl:27-164) for a hit on instance slot `slot`: normal, tangent, bitangent, position, texcoord
bpt_status obpt_unit_hit_vertex(obpt_context* c, uint32_t slot, uint32_t prim, float u, float v, float out[17]) {
    CHECK_CTX(c);
    if (!c->scene.accel_built || slot >= c->scene.xf.size()) return fail(c, BPT_ERR_INVALID, "unit_hit_vertex: bad slot or no accel");
    Vertex vt = fetch_vertex_attributes(c->scene, c->scene.xf[slot], prim, u, v);
    const f3 a[4] = {vt.normal_world, vt.tangent_world, vt.bitangent_world, vt.position_world};
    for (int k = 0; k < 4; k++) { out[3 * k] = a[k].x; out[3 * k + 1] = a[k].y; out[3 * k + 2] = a[k].z; }
    out[12] = vt.texcoord.x; out[13] = vt.texcoord.y;
    out[14] = vt.color.x; out[15] = vt.color.y; out[16] = vt.color.z;
    return BPT_OK;
}
float obpt_unit_log2(float x) { return log2_(x); }
float obpt_store_half(float f) { return store_half(f); }
void obpt_gbuffer_roundtrip(const float N[3], const float T[3], const float in[12], uint32_t model, float out[18], uint32_t* model_out) {
    SurfaceData s = surface_data_default();
    s.base_color = mk3(in[0], in[1], in[2]); s.f0_color = mk3(in[3], in[4], in[5]); s.f90_color = mk3(in[6], in[7], in[8]);
    s.roughness = in[9]; s.anisotropy = in[10]; s.ior = in[11];
    GBuffer g = store_gbuffer(pack_surface_to_gbuffer(mk3(N[0], N[1], N[2]), mk3(T[0], T[1], T[2]), s, model));
    f3 No, To;
    unpack_gbuffer_to_surface(g, No, To, s, model);
    const float o[18] = {No.x, No.y, No.z, To.x, To.y, To.z, s.base_color.x, s.base_color.y, s.base_color.z, s.f0_color.x, s.f0_color.y, s.f0_color.z,
                         s.f90_color.x, s.f90_color.y, s.f90_color.z, s.roughness, s.anisotropy, s.ior};
    std::memcpy(out, o, sizeof(o)); *model_out = model;
}
uint64_t obpt_morton63(const float c[3], const float lo[3], const float hi[3]) {
    return morton63(mk3(c[0], c[1], c[2]), mk3(lo[0], lo[1], lo[2]), mk3(hi[0], hi[1], hi[2]));
}

} // extern "C"
