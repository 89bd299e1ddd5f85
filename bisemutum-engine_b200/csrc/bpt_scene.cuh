// bpt_scene.cuh — device-side scene description, vertex fetch, material function, sky and light
// evaluation. Layout in HBM (DESIGN.md "Data layout"):
//   * geometry streams exactly as the reference's GpuSceneData (flat float / uint arrays,
//     bisemutum/src/graphics/gpu_scene_data.hpp:10-24) + 36-B DrawableSbtData records
//   * BVH nodes: 64 B = 4 x float4 (both child boxes + child ids) → four 16-B vector loads
//   * triangles: 48 B = 3 x float4 (v0|prim, e1|instance slot, e2|-) in BVH leaf order
//   * instances: 128-B records (object→world 3x4, world→object 3x4, ids) as float4[8]
#pragma once
#include "../../include/bpt/bpt.h"
#include "bpt_math.cuh"

#if defined(__CUDA_ARCH__)
#define BPT_LDG(p) __ldg(p)
#else
#define BPT_LDG(p) (*(p))
#endif
// 256-bit read-only loads (sm_100: LDG.E.256): a 64-B BVH node is two load instructions instead of four, an exact leaf box one instead of
// two. The traversal kernels of the incoherent bounces run the L1 at 75-88 % of its throughput with every lane fetching its own node, and
// the L1 handles one request per instruction per distinct line — fewer, wider instructions are fewer requests. `p` must be 32-byte aligned
// (nodes: 64-B records, leaf boxes: 32-B records, both in 256-B aligned allocations). BPT_LDG256=0: four / two 128-bit loads (A/B builds).
// Measured (profiles/r2y_variants.jsonl): configs[1] 1.525 -> 1.498 ms per sample (extend 0.924 -> 0.905, connect 0.215 -> 0.211); the
// two-level kernels keep the 128-bit loads (W256 = false): atrium two-level +0.8 %, but the DRAM-sized instanced scene -2 %.
#ifndef BPT_LDG256
#define BPT_LDG256 1
#endif

namespace bptd {

struct DBlas {
    const float4* nodes;   // 4 per node
    const float4* tris;    // 3 per triangle, leaf order
    int32_t root;
    uint32_t n;
    const float4* wide;    // 4 per node: 4-wide quantised form (bpt_wide.cuh); nullptr when not built
    const float4* leafbox; // 2 per leaf: exact leaf boxes for the wide traversal
};

struct DTexture {
    const void* texels;
    uint32_t w, h, format, addr_u, addr_v, linear;
};

// A rect-light texture: the whole mip chain as FP32 texels, level after level (level l is max(w >> l, 1) x max(h >> l, 1)).
struct DLightTexture {
    const float4* texels;
    uint32_t w, h, levels, addr_u, addr_v, linear, mip_linear;
};

struct DInstance {           // 128 B
    float o2w[12];
    float w2o[12];
    uint32_t instance_id, flags, blas;
    uint32_t anyhit;         // 1 if candidates of this instance must run the any-hit opacity rule
    uint32_t pad[4];
};

struct DScene {
    const float* positions; const float* normals; const float* tangents; const float* texcoords; const float* colors;
    const uint32_t* indices;
    const bpt_drawable_sbt_data* drawables;
    const uint32_t* drawable_va;
    const bpt_material* materials;
    const DTexture* textures; uint32_t num_textures;
    const DInstance* instances; uint32_t num_instances;
    // accel
    uint32_t accel_mode;
    const float4* tlas_nodes; const uint32_t* tlas_prims; int32_t tlas_root; uint32_t tlas_n;
    const float4* tlas_wide; const float4* tlas_leafbox;      // 4-wide form of the TLAS + exact instance boxes (two-level mode)
    const DBlas* blas;
    // lights
    const bpt_dir_light_data* dir_lights; uint32_t num_dir;
    const bpt_point_light_data* point_lights; uint32_t num_point;
    const bpt_rect_light_data* rect_lights; uint32_t num_rect;
    const float* ltc_m0; const float* ltc_m1; const float* ltc_m2; const float* ltc_norm;
    const DLightTexture* light_textures; uint32_t num_light_textures;
    // sky
    const float4* sky_faces; uint32_t sky_size;
    float sky_transform[9]; float sky_color[3];
    // image-based lighting of the sky (bpt_precompute_sky_ibl; ibl_enabled = 0: not computed). Layout: bpt_ibl.cuh
    uint32_t ibl_enabled, ibl_diffuse_size, ibl_specular_size, ibl_specular_levels, ibl_brdf_size;
    const float4* ibl_diffuse; const float4* ibl_specular; const float2* ibl_brdf;
    float ibl_diffuse_color[3]; float ibl_specular_color[3];
    // DDGI volume of the previous update (bpt_set_ddgi_volume; ddgi_enabled = 0: none): atlases in bpt_blend_probes' layout
    uint32_t ddgi_enabled, ddgi_irr_size, ddgi_vis_size;
    const float4* ddgi_irradiance; const float2* ddgi_visibility;
    bpt_probe_volume ddgi_volume;
};

template <bool W256 = true>
BPT_HD void ldg_32B(const float4* p, float4& a, float4& b) {
#if defined(__CUDA_ARCH__) && BPT_LDG256
    if (W256) {
        asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
        return;
    }
#endif
    a = BPT_LDG(p); b = BPT_LDG(p + 1);
}
template <bool W256 = true>
BPT_HD void ldg_64B(const float4* p, float4& a, float4& b, float4& c, float4& d) { ldg_32B<W256>(p, a, b); ldg_32B<W256>(p + 2, c, d); }

BPT_HD float3 xf_point(const float* m, float3 p) {
    return v3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
              ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
BPT_HD float3 xf_vector(const float* m, float3 v) {
    return v3((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z, (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}
BPT_HD float3 xf_vector_t(const float* m, float3 v) {   // transpose(upper 3x3) * v
    return v3((m[0] * v.x + m[4] * v.y) + m[8] * v.z, (m[1] * v.x + m[5] * v.y) + m[9] * v.z, (m[2] * v.x + m[6] * v.y) + m[10] * v.z);
}

// Inverse of a row-major 3x4 affine matrix: adjugate / determinant, then t' = -(A^-1 t).
BPT_HD void invert_3x4(const float* m, float* o) {
    float a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[4], a11 = m[5], a12 = m[6], a20 = m[8], a21 = m[9], a22 = m[10];
    float c00 = a11 * a22 - a12 * a21;
    float c01 = a12 * a20 - a10 * a22;
    float c02 = a10 * a21 - a11 * a20;
    float det = (a00 * c00 + a01 * c01) + a02 * c02;
    float inv_det = 1.0f / det;
    o[0] = c00 * inv_det; o[1] = (a02 * a21 - a01 * a22) * inv_det; o[2] = (a01 * a12 - a02 * a11) * inv_det;
    o[4] = c01 * inv_det; o[5] = (a00 * a22 - a02 * a20) * inv_det; o[6] = (a02 * a10 - a00 * a12) * inv_det;
    o[8] = c02 * inv_det; o[9] = (a01 * a20 - a00 * a21) * inv_det; o[10] = (a00 * a11 - a01 * a10) * inv_det;
    float tx = m[3], ty = m[7], tz = m[11];
    o[3] = -((o[0] * tx + o[1] * ty) + o[2] * tz);
    o[7] = -((o[4] * tx + o[5] * ty) + o[6] * tz);
    o[11] = -((o[8] * tx + o[9] * ty) + o[10] * tz);
}

// ---- textures: explicit FP32 bilinear (texture units filter with 8-bit weights → > 1e-4) --------
BPT_HD int wrap_tc(int c, int n, uint32_t mode) {
    if (mode == BPT_ADDRESS_REPEAT) { int m = c % n; return m < 0 ? m + n : m; }
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}
// wrap_tc(c) and wrap_tc(c + 1) with ONE modulo (an integer % by a run-time n is ~20 instructions; a bilinear fetch needed four of them),
// none at all for power-of-two sizes. Integer arithmetic: the same indices as two wrap_tc calls.
BPT_HD void wrap_tc2(int c, int n, uint32_t mode, int& a, int& b) {
    if (mode == BPT_ADDRESS_REPEAT) {
        int m;
        if ((n & (n - 1)) == 0) m = c & (n - 1);
        else { m = c % n; m = m < 0 ? m + n : m; }
        a = m; b = m + 1 == n ? 0 : m + 1;
    } else {
        a = c < 0 ? 0 : (c >= n ? n - 1 : c);
        const int d = c + 1;
        b = d < 0 ? 0 : (d >= n ? n - 1 : d);
    }
}
// k / 255 for k in 0..255, correctly rounded, without an IEEE division (16 of them per bilinear RGBA8 fetch were ~160 instructions):
// q = k * r, one Newton residual step with r = fl(1 / 255). Equal to (float)k / 255.0f for all 256 inputs (tests/test_hostcheck_parity.py
// checks every one against the division).
BPT_HD float unorm8_to_float(uint32_t k) {
    const float r = 1.0f / 255.0f, kf = (float)k;
    const float q = kf * r;
    return fmaf(fmaf(-q, 255.0f, kf), r, q);
}
BPT_HD float4 texel_at(const DTexture& t, int x, int y) {
    size_t i = (size_t)y * t.w + x;
    if (t.format == BPT_TEXTURE_RGBA8_UNORM) {
        uchar4 p = BPT_LDG(reinterpret_cast<const uchar4*>(t.texels) + i);
        return make_float4(unorm8_to_float(p.x), unorm8_to_float(p.y), unorm8_to_float(p.z), unorm8_to_float(p.w));
    }
    return BPT_LDG(reinterpret_cast<const float4*>(t.texels) + i);
}
BPT_HD float4 mix4(float4 a, float4 b, float t) {
    return make_float4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t);
}
BPT_HD float4 sample_tex(const DTexture& t, float u, float v) {
    if (!t.linear) {
        int xi = wrap_tc((int)floorf(u * (float)t.w), (int)t.w, t.addr_u);
        int yi = wrap_tc((int)floorf(v * (float)t.h), (int)t.h, t.addr_v);
        return texel_at(t, xi, yi);
    }
    float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0, x1, y0, y1;
    wrap_tc2((int)x0f, (int)t.w, t.addr_u, x0, x1);
    wrap_tc2((int)y0f, (int)t.h, t.addr_v, y0, y1);
    float4 top = mix4(texel_at(t, x0, y0), texel_at(t, x1, y0), fx);
    float4 bot = mix4(texel_at(t, x0, y1), texel_at(t, x1, y1), fx);
    return mix4(top, bot, fy);
}
BPT_HD float4 sample_or(const DScene& sc, int32_t tex, float2 uv, float4 dflt) {
    if (tex < 0 || (uint32_t)tex >= sc.num_textures) return dflt;
    return sample_tex(sc.textures[tex], uv.x, uv.y);
}

// ---- sky (deferred_lighting_secondary.hlsl:24-29; Vulkan cube face rule = inverse of
//      core/utils/cubemap.hlsl:3-21; bilinear inside the face, clamp to edge) --------------------
// ---- rect-light textures: Texture2D.SampleLevel(sampler, uv, level) on the generated chain (lights.hlsl:446) ----
BPT_HD float4 light_texel(const float4* base, int w, int x, int y) { return BPT_LDG(base + (size_t)y * w + x); }
BPT_HD float4 light_level_sample(const DLightTexture& t, uint32_t level, float u, float v) {
    size_t off = 0;
    for (uint32_t l = 0; l < level; l++) off += (size_t)((t.w >> l) ? (t.w >> l) : 1u) * ((t.h >> l) ? (t.h >> l) : 1u);
    const int w = (int)((t.w >> level) ? (t.w >> level) : 1u), h = (int)((t.h >> level) ? (t.h >> level) : 1u);
    const float4* base = t.texels + off;
    if (!t.linear) return light_texel(base, w, wrap_tc((int)floorf(u * (float)w), w, t.addr_u), wrap_tc((int)floorf(v * (float)h), h, t.addr_v));
    float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = wrap_tc((int)x0f, w, t.addr_u), x1 = wrap_tc((int)x0f + 1, w, t.addr_u);
    int y0 = wrap_tc((int)y0f, h, t.addr_v), y1 = wrap_tc((int)y0f + 1, h, t.addr_v);
    float4 top = mix4(light_texel(base, w, x0, y0), light_texel(base, w, x1, y0), fx);
    float4 bot = mix4(light_texel(base, w, x0, y1), light_texel(base, w, x1, y1), fx);
    return mix4(top, bot, fy);
}
BPT_HD float3 light_texture_sample(const DLightTexture& t, float u, float v, float level) {
    const float top = (float)(t.levels - 1u);
    level = level < 0.0f ? 0.0f : (level > top ? top : level);
    if (!t.mip_linear) {
        int l = (int)ceilf(level + 0.5f) - 1;
        l = l < 0 ? 0 : (l > (int)t.levels - 1 ? (int)t.levels - 1 : l);
        float4 c = light_level_sample(t, (uint32_t)l, u, v);
        return v3(c.x, c.y, c.z);
    }
    const float lf = floorf(level);
    const uint32_t l0 = (uint32_t)lf, l1 = l0 + 1u < t.levels ? l0 + 1u : l0;
    float4 a = light_level_sample(t, l0, u, v);
    if (l1 == l0) return v3(a.x, a.y, a.z);
    float4 b = light_level_sample(t, l1, u, v);
    return mix3(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), level - lf);
}

// Major-axis face selection of a direction (the Vulkan / D3D cube rule, the inverse of core/utils/cubemap.hlsl:3-21):
// face, and (s, t) in [-1, 1] on it.
BPT_HD bool cube_face_st(float3 d, int& face, float& s, float& t) {
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    float s_, t_, ma;
    if (ax >= ay && ax >= az) { face = d.x >= 0.0f ? 0 : 1; ma = ax; s_ = d.x >= 0.0f ? -d.z : d.z; t_ = -d.y; }
    else if (ay >= az) { face = d.y >= 0.0f ? 2 : 3; ma = ay; s_ = d.x; t_ = d.y >= 0.0f ? d.z : -d.z; }
    else { face = d.z >= 0.0f ? 4 : 5; ma = az; s_ = d.z >= 0.0f ? d.x : -d.x; t_ = -d.y; }
    if (!(ma > 0.0f)) return false;
    s = s_ / ma; t = t_ / ma;
    return true;
}
// cubemap_direction_from_layered_uv (core/utils/cubemap.hlsl:3-21) without the normalisation; (s, t) may lie outside [-1, 1]
BPT_HD float3 cube_dir_of(int face, float s, float t) {
    switch (face) {
        case 0: return v3(1.0f, -t, -s);
        case 1: return v3(-1.0f, -t, s);
        case 2: return v3(s, 1.0f, t);
        case 3: return v3(s, -1.0f, -t);
        case 4: return v3(s, -t, 1.0f);
        default: return v3(-s, -t, -1.0f);
    }
}
// One tap of a seamless bilinear footprint: texel (x, y) of `face`, where ONE of the coordinates may be -1 or n. Such a texel lies
// across an edge of the cube: its centre on the extended face plane is re-projected onto the cube and the nearest texel of the
// adjacent face is taken — the texel that touches the edge at the same position along it ((j + 1) n / (n + 1) floors to j).
BPT_HD float3 cube_tap(const float4* faces, int n, int face, int x, int y) {
    if (x < 0 || x >= n || y < 0 || y >= n) {
        const float inv = 1.0f / (float)n;
        float3 d = cube_dir_of(face, (2.0f * (float)x + 1.0f) * inv - 1.0f, (2.0f * (float)y + 1.0f) * inv - 1.0f);
        float s, t;
        cube_face_st(d, face, s, t);
        x = (int)floorf(0.5f * (s + 1.0f) * (float)n); y = (int)floorf(0.5f * (t + 1.0f) * (float)n);
        x = x < 0 ? 0 : (x >= n ? n - 1 : x); y = y < 0 ? 0 : (y >= n ? n - 1 : y);
    }
    float4 v = BPT_LDG(faces + ((size_t)face * n + y) * n + x);
    return v3(v.x, v.y, v.z);
}
// TextureCube.SampleLevel with a linear sampler (deferred_lighting_secondary.hlsl:26, skybox_precompute_*.hlsl): bilinear with
// SEAMLESS edges, as Vulkan / D3D12 filter cube maps — a footprint that crosses an edge takes the outside texels from the adjacent
// face; at a cube corner only three texels exist and the fourth is their average.
BPT_HD float3 sample_cube(const float4* faces, uint32_t size, float3 d) {
    if (size == 0) return v3s(0.0f);
    int face; float s_, t_;
    if (!cube_face_st(d, face, s_, t_)) return v3s(0.0f);
    float u = 0.5f * (s_ + 1.0f), v = 0.5f * (t_ + 1.0f);
    int n = (int)size;
    float x = u * (float)n - 0.5f, y = v * (float)n - 0.5f;
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const bool ox0 = x0 < 0, ox1 = x1 >= n, oy0 = y0 < 0, oy1 = y1 >= n;
    float3 a, b, c, e;
    if (!(ox0 || ox1 || oy0 || oy1)) {                      // the whole footprint inside the face (all but a 1-texel rim): four plain loads
        const float4* row0 = faces + ((size_t)face * n + y0) * n + x0;
        float4 va = BPT_LDG(row0), vb = BPT_LDG(row0 + 1), vc = BPT_LDG(row0 + n), ve = BPT_LDG(row0 + n + 1);
        a = v3(va.x, va.y, va.z); b = v3(vb.x, vb.y, vb.z); c = v3(vc.x, vc.y, vc.z); e = v3(ve.x, ve.y, ve.z);
    } else if ((ox0 || ox1) && (oy0 || oy1)) {                     // cube corner: the tap outside in both directions does not exist
        const int xi = ox0 ? x1 : x0, yi = oy0 ? y1 : y0;   // the in-face column / row
        const int xo = ox0 ? x0 : x1, yo = oy0 ? y0 : y1;   // the outside column / row
        float3 in_ = cube_tap(faces, n, face, xi, yi), ex = cube_tap(faces, n, face, xo, yi), ey = cube_tap(faces, n, face, xi, yo);
        float3 avg = ((in_ + ex) + ey) * (1.0f / 3.0f);
        auto pick = [&](int xx, int yy) { return xx == xi ? (yy == yi ? in_ : ey) : (yy == yi ? ex : avg); };
        a = pick(x0, y0); b = pick(x1, y0); c = pick(x0, y1); e = pick(x1, y1);
    } else {
        a = cube_tap(faces, n, face, x0, y0); b = cube_tap(faces, n, face, x1, y0);
        c = cube_tap(faces, n, face, x0, y1); e = cube_tap(faces, n, face, x1, y1);
    }
    float3 top = mix3(a, b, fx);
    float3 bot = mix3(c, e, fx);
    return mix3(top, bot, fy);
}
BPT_HD float3 sample_sky(const DScene& sc, float3 d) { return sample_cube(sc.sky_faces, sc.sky_size, d); }

// ---- vertex fetch (core/raytracing/hit.hlsl:27-164) ---------------------------------------------
struct HitVertex { float3 normal_world, tangent_world, bitangent_world, position_world, color; float2 texcoord; };

BPT_HD void load_tri_indices(const DScene& sc, const bpt_drawable_sbt_data& dr, uint32_t prim, uint32_t idx[3]) {
    const uint32_t* p = sc.indices + (size_t)dr.index_offset + 3ull * prim;
    idx[0] = BPT_LDG(p); idx[1] = BPT_LDG(p + 1); idx[2] = BPT_LDG(p + 2);
}
BPT_HD float bary_mix(float a0, float a1, float a2, float bu, float bv) { return (a0 + (a1 - a0) * bu) + (a2 - a0) * bv; }
BPT_HD float2 load_texcoord(const DScene& sc, const bpt_drawable_sbt_data& dr, uint32_t va, const uint32_t idx[3], float bu, float bv) {
    if (!(va & BPT_VA_TEXCOORD)) return make_float2(0.0f, 0.0f);
    const float* b = sc.texcoords + dr.texcoord_offset;
    float2 t0 = BPT_LDG(reinterpret_cast<const float2*>(b + 2ull * idx[0]));
    float2 t1 = BPT_LDG(reinterpret_cast<const float2*>(b + 2ull * idx[1]));
    float2 t2 = BPT_LDG(reinterpret_cast<const float2*>(b + 2ull * idx[2]));
    return make_float2(bary_mix(t0.x, t1.x, t2.x, bu, bv), bary_mix(t0.y, t1.y, t2.y, bu, bv));
}
BPT_HD float3 load3(const float* p) { return v3(BPT_LDG(p), BPT_LDG(p + 1), BPT_LDG(p + 2)); }
BPT_HD HitVertex fetch_hit_vertex(const DScene& sc, const DInstance& in, uint32_t prim, float bu, float bv, bool need_position, bool need_color = false) {
    const bpt_drawable_sbt_data& dr = sc.drawables[in.instance_id];
    uint32_t va = BPT_LDG(sc.drawable_va + in.instance_id);
    uint32_t idx[3];
    load_tri_indices(sc, dr, prim, idx);
    float3 normal = v3(0.0f, 0.0f, 1.0f);
    if (va & BPT_VA_NORMAL) {
        const float* b = sc.normals + dr.normal_offset;
        float3 n0 = load3(b + 3ull * idx[0]), n1 = load3(b + 3ull * idx[1]), n2 = load3(b + 3ull * idx[2]);
        normal = v3(bary_mix(n0.x, n1.x, n2.x, bu, bv), bary_mix(n0.y, n1.y, n2.y, bu, bv), bary_mix(n0.z, n1.z, n2.z, bu, bv));
    }
    float3 tangent = v3(1.0f, 0.0f, 0.0f); float tangent_w = 1.0f;
    if (va & BPT_VA_TANGENT) {
        const float* b = sc.tangents + dr.tangent_offset;
        float4 t0 = BPT_LDG(reinterpret_cast<const float4*>(b + 4ull * idx[0]));
        float4 t1 = BPT_LDG(reinterpret_cast<const float4*>(b + 4ull * idx[1]));
        float4 t2 = BPT_LDG(reinterpret_cast<const float4*>(b + 4ull * idx[2]));
        tangent = v3(bary_mix(t0.x, t1.x, t2.x, bu, bv), bary_mix(t0.y, t1.y, t2.y, bu, bv), bary_mix(t0.z, t1.z, t2.z, bu, bv));
        tangent_w = bary_mix(t0.w, t1.w, t2.w, bu, bv);
    }
    HitVertex hv;
    hv.normal_world = normalize3(xf_vector_t(in.w2o, normal));                       // hit.hlsl:157
    hv.tangent_world = normalize3(xf_vector(in.o2w, tangent));                       // hit.hlsl:158
    hv.bitangent_world = normalize3(cross3(hv.normal_world, hv.tangent_world)) * tangent_w;   // hit.hlsl:159
    hv.texcoord = load_texcoord(sc, dr, va, idx, bu, bv);
    hv.color = v3s(0.0f);
    if (need_color && (va & BPT_VA_COLOR)) {                                         // hit.hlsl:97-113: color2 is read at index.x (as upstream)
        const float* b = sc.colors + dr.color_offset;
        float3 c0 = load3(b + 3ull * idx[0]), c1 = load3(b + 3ull * idx[1]), c2 = load3(b + 3ull * idx[0]);
        hv.color = v3(bary_mix(c0.x, c1.x, c2.x, bu, bv), bary_mix(c0.y, c1.y, c2.y, bu, bv), bary_mix(c0.z, c1.z, c2.z, bu, bv));
    }
    hv.position_world = v3s(0.0f);
    if (need_position) {                                                             // hit.hlsl:34-52,152
        const float* b = sc.positions + dr.position_offset;
        float3 p0 = load3(b + 3ull * idx[0]), p1 = load3(b + 3ull * idx[1]), p2 = load3(b + 3ull * idx[2]);
        float3 p = v3(bary_mix(p0.x, p1.x, p2.x, bu, bv), bary_mix(p0.y, p1.y, p2.y, bu, bv), bary_mix(p0.z, p1.z, p2.z, bu, bv));
        hv.position_world = xf_point(in.o2w, p);
    }
    return hv;
}

// ---- material_function: closed set of the reference's HLSL snippets -----------------------------
BPT_HD bool material_needs_position(const bpt_material& m) { return ((m.flags >> BPT_MATERIAL_KIND_SHIFT) & 0xffu) == BPT_MATERIAL_KIND_CHECKERBOARD; }
BPT_HD bool material_needs_color(const bpt_material& m) { return ((m.flags >> BPT_MATERIAL_KIND_SHIFT) & 0xffu) == BPT_MATERIAL_KIND_VERTEX_COLOR; }
BPT_HD Surface eval_material(const DScene& sc, const bpt_material& m, float2 uv, float3 position_world, float3 vertex_color = v3s(0.0f)) {
    Surface s = surface_default();
    uint32_t kind = (m.flags >> BPT_MATERIAL_KIND_SHIFT) & 0xffu;
    if (kind == BPT_MATERIAL_KIND_GLTF_PBR) {                                        // import_model.cpp:208-230
        float4 bt = sample_or(sc, m.base_color_tex, uv, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
        s.base_color = v3(bt.x * m.base_color[0], bt.y * m.base_color[1], bt.z * m.base_color[2]);
        s.opacity = bt.w * m.base_color[3];
        float4 nt = sample_or(sc, m.normal_map_tex, uv, make_float4(0.5f, 0.5f, 1.0f, 1.0f));
        float3 nm = v3(nt.x * 2.0f - 1.0f, nt.y * 2.0f - 1.0f, nt.z * 2.0f - 1.0f);
        nm = normalize3(nm * v3(m.normal_map_scale, m.normal_map_scale, 1.0f));
        s.normal_map_value = nm * 0.5f + v3s(0.5f);
        float4 mr = sample_or(sc, m.metallic_roughness_tex, uv, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
        s.roughness = m.roughness * mr.y;
        s.f0_color = mix3(v3s(0.04f), s.base_color, m.metallic * mr.z);
        float occlusion = sample_or(sc, m.occlusion_tex, uv, make_float4(1.0f, 1.0f, 1.0f, 1.0f)).x * m.occlusion_strength;
        s.base_color = s.base_color * occlusion;
        s.f0_color = s.f0_color * occlusion;
        s.f90_color = s.f90_color * occlusion;
        s.two_sided = (m.flags & BPT_MATERIAL_FLAG_TWO_SIDED) != 0;
    } else if (kind == BPT_MATERIAL_KIND_ASSIMP_DIFFUSE) {                           // import_model.cpp:490-493
        s.base_color = v3(m.base_color[0], m.base_color[1], m.base_color[2]);
        s.roughness = m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_CONSTANT_COLOR) {                           // examples/scene_basic/materials/white.toml
        s.base_color = v3(m.base_color[0], m.base_color[1], m.base_color[2]);
    } else if (kind == BPT_MATERIAL_KIND_CHECKERBOARD) {                             // .../checkerboard.toml
        int grid = (int)floorf(position_world.x) ^ (int)floorf(position_world.z);
        bool odd = (grid & 1) == 1;
        s.base_color = odd ? v3(m.base_color[0], m.base_color[1], m.base_color[2]) : v3(m.emission[0], m.emission[1], m.emission[2]);
        s.roughness = odd ? m.base_color[3] : m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_TEXTURED) {                                 // .../textured.toml
        float4 bt = sample_or(sc, m.base_color_tex, uv, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
        float4 nt = sample_or(sc, m.normal_map_tex, uv, make_float4(0.5f, 0.5f, 1.0f, 1.0f));
        s.base_color = v3(bt.x, bt.y, bt.z);
        s.normal_map_value = v3(nt.x, nt.y, nt.z);
        s.roughness = m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_TRANSPARENT) {                              // .../transparent.toml
        s.base_color = v3(m.base_color[0], m.base_color[1], m.base_color[2]);
        s.opacity = m.base_color[3];
        s.two_sided = true;
    } else if (kind == BPT_MATERIAL_KIND_VERTEX_COLOR) {                             // surface.base_color = vertex.color
        s.base_color = vertex_color;
        s.roughness = m.roughness;
    } else if (kind == BPT_MATERIAL_KIND_CAGE) {                                     // .../cage.toml
        float4 v = sample_or(sc, m.base_color_tex, uv, make_float4(1.0f, 1.0f, 1.0f, 1.0f));
        s.base_color = v3(v.x, v.y, v.z);
        s.f0_color = v3(v.x, v.y, v.z);
        s.opacity = v.w < 0.5f ? 0.0f : 1.0f;
        s.two_sided = true;
    }
    return s;
}
BPT_HD float eval_hit_opacity(const DScene& sc, uint32_t instance_id, uint32_t prim, float bu, float bv) {
    const bpt_drawable_sbt_data& dr = sc.drawables[instance_id];
    const bpt_material& m = sc.materials[dr.material_offset / (uint32_t)sizeof(bpt_material)];
    uint32_t idx[3];
    load_tri_indices(sc, dr, prim, idx);
    float2 uv = load_texcoord(sc, dr, BPT_LDG(sc.drawable_va + instance_id), idx, bu, bv);
    return eval_material(sc, m, uv, v3s(0.0f)).opacity;
}

// ---- point / spot light (lights.hlsl:14-25) -----------------------------------------------------
BPT_HD float3 eval_point_light(const bpt_point_light_data& l, float3 P, float3& light_dir, float& dist) {
    float3 lv = v3(l.position[0], l.position[1], l.position[2]) - P;
    float d2 = dot3(lv, lv);
    dist = sqrtf(d2);
    light_dir = lv / dist;
    float att = sat(1.0f - sq(d2 * l.range_sqr_inv)) / tmax_(d2, 0.001f);
    if (l.cos_inner > l.cos_outer) {
        float ct = clampf_(dot3(light_dir, v3(l.direction[0], l.direction[1], l.direction[2])), l.cos_outer, l.cos_inner);
        att = att * ((ct - l.cos_outer) / tmax_(l.cos_inner - l.cos_outer, 0.001f));
    }
    return v3(l.emission[0], l.emission[1], l.emission[2]) * att;
}

// ---- camera ray (generate_camera_ray.hlsl:4-16, camera.hlsl:7-9) --------------------------------
// pixel_jitter (NEW switch, default off = the reference's fixed pixel centres): offsets from the bounce-0 RNG stream
BPT_HD void camera_ray(const bpt_camera& cam, uint32_t px, uint32_t py, uint32_t W, uint32_t H, uint32_t jitter, uint32_t frame_index, float3& O, float3& D) {
    const float* ip = cam.matrix_inv_proj;
    const float* iv = cam.matrix_inv_view;
    float jx = 0.5f, jy = 0.5f;
    if (jitter) { uint32_t seed = rng_tea(py * W + px, frame_index); jx = rng_next(seed); jy = rng_next(seed); }
    float uvx = ((float)px + jx) / (float)W, uvy = ((float)py + jy) / (float)H;
    float nx = uvx * 2.0f - 1.0f, ny = 1.0f - uvy * 2.0f;
    float3 dl = v3(((ip[0] * nx + ip[4] * ny) + ip[8]) + ip[12], ((ip[1] * nx + ip[5] * ny) + ip[9]) + ip[13],
                   ((ip[2] * nx + ip[6] * ny) + ip[10]) + ip[14]);
    dl = normalize3(dl);
    float3 dw = v3((iv[0] * dl.x + iv[4] * dl.y) + iv[8] * dl.z, (iv[1] * dl.x + iv[5] * dl.y) + iv[9] * dl.z,
                   (iv[2] * dl.x + iv[6] * dl.y) + iv[10] * dl.z);
    D = normalize3(dw);
    O = v3(iv[12], iv[13], iv[14]);
}

// ---- DDGI-style probe rays (ddgi/trace_gbuffer.hlsl:10-36) ---------------------------------------
BPT_HD void probe_ray(const bpt_probe_volume& vol, const float2* sample_table, uint32_t path, uint32_t frame_index, float3& O, float3& D) {
    uint32_t ray_index = path % vol.rays_per_probe, lin = path / vol.rays_per_probe;
    uint32_t ix = lin % vol.probe_counts[0], iy = (lin / vol.probe_counts[0]) % vol.probe_counts[1], iz = lin / vol.probe_counts[0] / vol.probe_counts[1];
    float mx = (float)(vol.probe_counts[0] > 1 ? vol.probe_counts[0] - 1 : 1), my = (float)(vol.probe_counts[1] > 1 ? vol.probe_counts[1] - 1 : 1),
          mz = (float)(vol.probe_counts[2] > 1 ? vol.probe_counts[2] - 1 : 1);
    float3 fx = v3(vol.frame_x[0], vol.frame_x[1], vol.frame_x[2]), fy = v3(vol.frame_y[0], vol.frame_y[1], vol.frame_y[2]), fz = v3(vol.frame_z[0], vol.frame_z[1], vol.frame_z[2]);
    O = ((v3(vol.base_position[0], vol.base_position[1], vol.base_position[2]) + ((float)ix * vol.extent[0] / mx) * fx) + ((float)iy * vol.extent[1] / my) * fy) +
        ((float)iz * vol.extent[2] / mz) * fz;
    uint32_t seed = rng_tea(lin, frame_index);
    uint32_t rand_index = ((uint32_t)(rng_next(seed) * 8192.0f) + ray_index) % 8192u;
    float2 r = BPT_LDG(sample_table + rand_index);
    D = uniform_sphere_sample(r.x, r.y);
}

} // namespace bptd
