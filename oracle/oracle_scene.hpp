// oracle_scene.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Scene storage + BVH structs.
#pragma once
#include <atomic>
#include <string>
#include <vector>
#include "oracle.h"
#include "oracle_math.hpp"

namespace orc {

struct Tri {            // 48 B record, same content as the GPU's AoS-of-float4 triangle
    f3 v0; uint32_t prim;
    f3 e1; uint32_t inst;   // inst only meaningful in merged mode
    f3 e2; uint32_t pad;
};

struct Bvh {
    uint32_t n = 0;                    // leaves
    std::vector<uint64_t> morton;      // sorted
    std::vector<uint32_t> prims;       // sorted primitive ids
    std::vector<bpt_bvh_node> nodes;   // n-1 internal nodes
    std::vector<int32_t> leaf_parent;
    int32_t root = 0;                  // 0, or ~0 when n == 1
    f3 lo{0, 0, 0}, hi{0, 0, 0};       // root bounds
    std::vector<Tri> tris;             // BLAS only, sorted order
    // 4-wide quantised form of this tree (oracle_wide.cpp; merged mode, built on demand): node i = subtree of binary node i,
    // boxes already decoded to the planes the traversal tests
    struct Wide4 { int nchild; int32_t child[4]; f3 clo[4], chi[4]; };
    std::vector<Wide4> wide;
};
void build_wide4(Bvh& b);

struct Texture {
    uint32_t w = 0, h = 0, format = 0, addr_u = 0, addr_v = 0, linear = 1;
    std::vector<uint8_t> texels;
};

struct InstanceXf {      // derived per instance at build time
    float o2w[12];       // row-major 3x4
    float w2o[12];       // row-major 3x4 inverse
    uint32_t instance_id, flags, blas;
};

struct TraceStats {
    uint64_t rays = 0, nodes = 0, tris = 0, instances = 0;
    uint64_t wide_nodes = 0, leaf_boxes = 0;      // steps through the 4-wide tree and exact leaf-box tests (rays traced with wide = true)
    void add(const TraceStats& o) { rays += o.rays; nodes += o.nodes; tris += o.tris; instances += o.instances; wide_nodes += o.wide_nodes; leaf_boxes += o.leaf_boxes; }
};

// Rect-light texture with its generated mip chain: one RGBA32F image per level (what the sampler returns after format decoding).
struct LightTexture {
    uint32_t w = 0, h = 0, addr_u = 0, addr_v = 0;
    bool linear = true, mip_linear = false;
    std::vector<std::vector<float>> level;          // level[l]: max(w >> l, 1) * max(h >> l, 1) * 4 floats
};

struct Scene {
    std::vector<float> positions, normals, tangents, colors, texcoords, texcoords2;
    std::vector<uint32_t> indices;
    std::vector<bpt_drawable_sbt_data> drawables;
    std::vector<uint32_t> drawable_va;
    std::vector<bpt_blas_desc> blas_descs;
    std::vector<bpt_instance_desc> instances;
    std::vector<bpt_material> materials;
    std::vector<Texture> textures;
    std::vector<bpt_dir_light_data> dir_lights;
    std::vector<bpt_point_light_data> point_lights;
    std::vector<bpt_rect_light_data> rect_lights;
    std::vector<float> ltc_m0, ltc_m1, ltc_m2, ltc_norm;
    std::vector<LightTexture> light_textures;
    std::vector<float> sky_faces;     // 6 * size * size * 4
    uint32_t sky_size = 0;
    float sky_transform[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float sky_color[3] = {1, 1, 1};
    // SkyboxPrecomputePass outputs (obpt_precompute_sky_ibl): rgba cube faces (diffuse; specular = all mips, level 0 first), rg LUT
    bool ibl_valid = false;
    bpt_sky_ibl_desc ibl_desc{};
    std::vector<float> ibl_diffuse, ibl_specular, ibl_brdf;

    // accel
    bool accel_built = false;
    uint32_t wide_from_bounce = 0;   // obpt_set_wide_from_bounce: bounces >= this walk the 4-wide tree (0 = never; merged mode only)
    uint32_t accel_mode = 0;
    std::vector<Bvh> blas;
    Bvh tlas;
    std::vector<InstanceXf> xf;
};

// oracle_bvh.cpp
void build_lbvh(Bvh& out, uint32_t n, const f3* lo, const f3* hi);
bool build_accel(Scene& sc, uint32_t mode, std::string& err);
bool build_tlas(Scene& sc, std::string& err);
uint64_t morton63(f3 c, f3 lo, f3 hi);
void invert_3x4(const float m[12], float out[12]);
static inline f3 xf_point(const float m[12], f3 p) {
    return mk3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
               ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
               ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
static inline f3 xf_vector(const float m[12], f3 v) {
    return mk3((m[0] * v.x + m[1] * v.y) + m[2] * v.z,
               (m[4] * v.x + m[5] * v.y) + m[6] * v.z,
               (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
}
// transpose(upper 3x3 of m) * v  — used for normals with m = world→object
static inline f3 xf_vector_transposed(const float m[12], f3 v) {
    return mk3((m[0] * v.x + m[4] * v.y) + m[8] * v.z,
               (m[1] * v.x + m[5] * v.y) + m[9] * v.z,
               (m[2] * v.x + m[6] * v.y) + m[10] * v.z);
}

struct HitRec { float t, u, v; uint32_t inst_slot, instance_id, prim; bool hit; };
// wide = true (merged mode with sc.blas[0].wide built): walk the 4-wide quantised tree like the CUDA kernels do from the second
// bounce on. Same hits by construction (oracle_wide.cpp header); only the work counters differ.
HitRec trace_closest(const Scene& sc, f3 O, f3 D, float tmin, float tmax, uint32_t frame_index, TraceStats& st, bool wide = false);
bool trace_any(const Scene& sc, f3 O, f3 D, float tmin, float tmax, uint32_t frame_index, TraceStats& st, bool cull_non_opaque = false, bool wide = false);

// oracle_render.cpp helpers used by traversal (any-hit opacity)
float eval_opacity(const Scene& sc, uint32_t instance_id, uint32_t prim, float u, float v);

} // namespace orc

struct obpt_reblur_state;            // oracle_reblur.cpp: the denoiser's per-camera history
void obpt_reblur_free(struct obpt_context* c);

struct obpt_context {
    orc::Scene scene;
    obpt_reblur_state* reblur = nullptr;
    uint32_t width = 0, height = 0;
    uint32_t threads = 0;
    uint32_t tile_stride = 1, tile_offset = 0;
    std::string err;
    // DDGI volume of the previous update (obpt_set_ddgi_volume): atlases in obpt_blend_probes' layout
    bool ddgi_enabled = false; uint32_t ddgi_irr_size = 0, ddgi_vis_size = 0; bpt_probe_volume ddgi_volume{};
    std::vector<float> ddgi_irradiance, ddgi_visibility;
    std::vector<float> accum;        // W*H*4 FP32 sums (reference_fp16: the running half-valued average)
    bool accum_used = false, accum_fp16 = false;   // accumulation rule in force since the last clear
    uint32_t accum_count = 0;        // reference_fp16: samples already folded in (pt_accumulate weight 1/(count+1))
    bpt_counters counters{};
    obpt_stats stats{};
    bool capture = false;
    // capture storage, per bounce (index bounce-1)
    std::vector<std::vector<uint32_t>> cap_extend_pixels;
    std::vector<std::vector<bpt_hit>> cap_extend_hits;
    std::vector<std::vector<uint32_t>> cap_shadow_pixels, cap_shadow_lights;
};
