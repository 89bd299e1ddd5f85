// tests/hostcheck/hostcheck.cpp — TEST HARNESS ONLY, never shipped, never linked into libbpt.so.
//
// Compiles the kernels' per-thread device functions (bpt_math/scene/trace/shade .cuh, all
// BPT_HD) for the HOST with g++ -ffp-contract=off so that `pytest -m "not gpu"` can compare the
// CUDA source's arithmetic against the oracle without a GPU: an expression-order mismatch shows
// up here in seconds instead of after a GPU round trip. The BVH is NOT built here (the build is
// __global__ kernels); the harness is handed the oracle's BVH arrays, which the GPU build must
// match bit-exactly anyway (tests/test_gpu_parity.py checks that on the device).
#include <algorithm>
#include <cstdio>
#include <vector>
#include "../../bisemutum-engine_b200/csrc/bpt_shade.cuh"
#include "../../bisemutum-engine_b200/csrc/bpt_ddgi.cuh"
#include "../../bisemutum-engine_b200/csrc/bpt_aov.cuh"
#include "../../bisemutum-engine_b200/csrc/bpt_wide.cuh"
#include "../../bisemutum-engine_b200/csrc/bpt_post.cuh"
#include "../../bisemutum-engine_b200/csrc/bpt_reblur.cuh"
#include <cmath>

using namespace bptd;

struct hc_bvh {
    uint32_t n; int32_t root;
    const bpt_bvh_node* nodes;     // n-1
    const uint32_t* prims;         // n sorted primitive ids
};
struct hc_scene {
    const float *positions, *normals, *tangents, *texcoords;
    const uint32_t* indices;
    const bpt_drawable_sbt_data* drawables; const uint32_t* drawable_va; uint32_t num_drawables;
    const bpt_blas_desc* blas_desc; uint32_t num_blas;
    const bpt_instance_desc* instances; uint32_t num_instances;
    const bpt_material* materials; uint32_t num_materials;
    const bpt_dir_light_data* dir; uint32_t num_dir;
    const bpt_point_light_data* point; uint32_t num_point;
    const bpt_rect_light_data* rect; uint32_t num_rect;
    const float *ltc_m0, *ltc_m1, *ltc_m2, *ltc_norm;
    const float* sky_faces; uint32_t sky_size; float sky_transform[9]; float sky_color[3];
    uint32_t accel_mode;
    const hc_bvh* blas_bvh;        // num_blas entries (two-level) or 1 (merged)
    hc_bvh tlas;
    const bpt_texture_desc* textures; uint32_t num_textures;
    // rect-light textures: the generated chain as RGBA32F (read back from the oracle: the generator itself is a GPU kernel, checked on the GPU),
    // so that the SAMPLING and LTC arithmetic of the device headers runs on the host
    struct light_tex { const float* chain; uint32_t w, h, levels, addr_u, addr_v, linear, mip_linear; };
    const light_tex* light_textures; uint32_t num_light_textures;
    const float* colors;           // vertex colours (may be null)
};

namespace {
// DDGI volume bound by hc_set_ddgi (the host-check counterpart of bpt_set_ddgi_volume); applied by build()
struct { bool enabled = false; bpt_probe_volume vol{}; uint32_t irr_size = 0, vis_size = 0; std::vector<float> irr, vis; } g_ddgi;
// Sky IBL textures bound by hc_precompute_sky_ibl (the host-check counterpart of bpt_precompute_sky_ibl); applied by build()
struct { bool enabled = false; bpt_sky_ibl_desc desc{}; std::vector<float4> diffuse, specular; std::vector<float2> brdf; } g_ibl;
struct Built {
    std::vector<DInstance> inst;
    std::vector<std::vector<float4>> tris;
    std::vector<DBlas> blas;
    std::vector<std::vector<float4>> wide, leafbox;   // two-level mode: 4-wide form of every BLAS ([0..nb)) and of the TLAS ([nb])
    std::vector<DTexture> textures;
    std::vector<DLightTexture> light_textures;
    std::vector<std::vector<float>> decoded;     // sRGB textures decoded to linear FP32 (as bpt_scene_upload_materials does)
    DScene sc{};
};

float4 f4(float x, float y, float z, uint32_t w) { return make_float4(x, y, z, u2f(w)); }

// collapse_node4 / leaf_boxes_of_node per binary node, as k_collapse4 runs them
void collapse_tree(const hc_bvh& hb, std::vector<float4>& wide, std::vector<float4>& leafbox) {
    wide.assign(hb.n >= 2 ? 4ull * (hb.n - 1) : 0, make_float4(0, 0, 0, 0));
    leafbox.assign(2ull * hb.n, make_float4(0, 0, 0, 0));
    const float4* nodes2 = reinterpret_cast<const float4*>(hb.nodes);
    for (uint32_t i = 0; i + 1 < hb.n; i++) {
        collapse_node4(nodes2, (int32_t)i, &wide[4ull * i]);
        leaf_boxes_of_node(nodes2, (int32_t)i, leafbox.data());
    }
}

void build(const hc_scene& h, Built& b) {
    b.inst.resize(h.num_instances);
    for (uint32_t i = 0; i < h.num_instances; i++) {
        DInstance& d = b.inst[i];
        memcpy(d.o2w, h.instances[i].transform, 48);
        invert_3x4(d.o2w, d.w2o);
        d.instance_id = h.instances[i].instance_id_and_mask & 0xffffffu;
        d.flags = h.instances[i].sbt_offset_and_flags >> 24;
        d.blas = (uint32_t)h.instances[i].blas;
        uint32_t blend = (h.materials[h.drawables[d.instance_id].material_offset / sizeof(bpt_material)].flags >> BPT_MATERIAL_BLEND_SHIFT) & 0xffu;
        d.anyhit = ((d.flags & BPT_INSTANCE_FORCE_NON_OPAQUE) && blend != BPT_BLEND_OPAQUE) ? 1u : 0u;
    }
    auto vert = [&](const bpt_blas_desc& bd, uint32_t k, int c) {
        uint32_t idx = h.indices[(size_t)bd.index_offset + 3ull * k + c];
        const float* p = h.positions + (size_t)bd.position_offset + 3ull * idx;
        return v3(p[0], p[1], p[2]);
    };
    uint32_t nb = h.accel_mode == BPT_ACCEL_TWO_LEVEL ? h.num_blas : 1;
    b.tris.resize(nb); b.blas.resize(nb);
    if (h.accel_mode == BPT_ACCEL_TWO_LEVEL) {
        for (uint32_t bi = 0; bi < nb; bi++) {
            const hc_bvh& hb = h.blas_bvh[bi];
            b.tris[bi].resize(3ull * hb.n);
            for (uint32_t j = 0; j < hb.n; j++) {
                uint32_t k = hb.prims[j];
                float3 v0 = vert(h.blas_desc[bi], k, 0), v1 = vert(h.blas_desc[bi], k, 1), v2 = vert(h.blas_desc[bi], k, 2);
                float3 e1 = v1 - v0, e2 = v2 - v0;
                b.tris[bi][3 * j] = f4(v0.x, v0.y, v0.z, k); b.tris[bi][3 * j + 1] = f4(e1.x, e1.y, e1.z, 0); b.tris[bi][3 * j + 2] = f4(e2.x, e2.y, e2.z, 0);
            }
        }
    } else {
        const hc_bvh& hb = h.blas_bvh[0];
        std::vector<uint32_t> slot_of(hb.n), prim_of(hb.n);
        uint32_t g = 0;
        for (uint32_t s = 0; s < h.num_instances; s++)
            for (uint32_t k = 0; k < h.blas_desc[b.inst[s].blas].num_triangles; k++, g++) { slot_of[g] = s; prim_of[g] = k; }
        b.tris[0].resize(3ull * hb.n);
        for (uint32_t j = 0; j < hb.n; j++) {
            uint32_t p = hb.prims[j], s = slot_of[p], k = prim_of[p];
            const bpt_blas_desc& bd = h.blas_desc[b.inst[s].blas];
            float3 v0 = xf_point(b.inst[s].o2w, vert(bd, k, 0)), v1 = xf_point(b.inst[s].o2w, vert(bd, k, 1)), v2 = xf_point(b.inst[s].o2w, vert(bd, k, 2));
            float3 e1 = v1 - v0, e2 = v2 - v0;
            b.tris[0][3 * j] = f4(v0.x, v0.y, v0.z, k); b.tris[0][3 * j + 1] = f4(e1.x, e1.y, e1.z, s); b.tris[0][3 * j + 2] = f4(e2.x, e2.y, e2.z, b.inst[s].anyhit);
        }
    }
    const bool two_level = h.accel_mode == BPT_ACCEL_TWO_LEVEL;
    b.wide.resize(nb + 1); b.leafbox.resize(nb + 1);
    for (uint32_t bi = 0; bi < nb; bi++) {
        if (two_level) collapse_tree(h.blas_bvh[bi], b.wide[bi], b.leafbox[bi]);
        b.blas[bi] = DBlas{reinterpret_cast<const float4*>(h.blas_bvh[bi].nodes), b.tris[bi].data(), h.blas_bvh[bi].root, h.blas_bvh[bi].n,
                           two_level && h.blas_bvh[bi].n >= 2 ? b.wide[bi].data() : nullptr, two_level && h.blas_bvh[bi].n >= 2 ? b.leafbox[bi].data() : nullptr};
    }
    if (two_level) collapse_tree(h.tlas, b.wide[nb], b.leafbox[nb]);
    DScene& s = b.sc;
    s.positions = h.positions; s.normals = h.normals; s.tangents = h.tangents; s.texcoords = h.texcoords; s.indices = h.indices;
    s.drawables = h.drawables; s.drawable_va = h.drawable_va; s.materials = h.materials;
    b.textures.resize(h.num_textures); b.decoded.resize(h.num_textures);
    for (uint32_t i = 0; i < h.num_textures; i++) {
        const bpt_texture_desc& t = h.textures[i];
        const void* texels = t.texels; uint32_t fmt = t.format;
        if (fmt == BPT_TEXTURE_RGBA8_SRGB) {
            float lut[256];
            for (int k = 0; k < 256; k++) { double v = k / 255.0; lut[k] = (float)(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4)); }
            const uint8_t* src = static_cast<const uint8_t*>(t.texels);
            std::vector<float>& lin = b.decoded[i];
            lin.resize((size_t)t.width * t.height * 4);
            for (size_t k = 0; k < lin.size(); k += 4) { lin[k] = lut[src[k]]; lin[k + 1] = lut[src[k + 1]]; lin[k + 2] = lut[src[k + 2]]; lin[k + 3] = (float)src[k + 3] / 255.0f; }
            texels = lin.data(); fmt = BPT_TEXTURE_RGBA32_FLOAT;
        }
        b.textures[i] = DTexture{texels, t.width, t.height, fmt, t.address_mode_u, t.address_mode_v, t.filter_linear};
    }
    s.textures = b.textures.data(); s.num_textures = h.num_textures;
    b.light_textures.resize(h.num_light_textures);
    for (uint32_t i = 0; i < h.num_light_textures; i++) {
        const hc_scene::light_tex& t = h.light_textures[i];
        b.light_textures[i] = DLightTexture{reinterpret_cast<const float4*>(t.chain), t.w, t.h, t.levels, t.addr_u, t.addr_v, t.linear, t.mip_linear};
    }
    s.light_textures = b.light_textures.data(); s.num_light_textures = h.num_light_textures;
    s.colors = h.colors;
    s.instances = b.inst.data(); s.num_instances = h.num_instances;
    s.accel_mode = h.accel_mode;
    s.tlas_nodes = reinterpret_cast<const float4*>(h.tlas.nodes); s.tlas_prims = h.tlas.prims; s.tlas_root = h.tlas.root; s.tlas_n = h.tlas.n;
    s.tlas_wide = two_level && h.tlas.n >= 2 ? b.wide[nb].data() : nullptr; s.tlas_leafbox = two_level && h.tlas.n >= 2 ? b.leafbox[nb].data() : nullptr;
    s.blas = b.blas.data();
    s.dir_lights = h.dir; s.num_dir = h.num_dir; s.point_lights = h.point; s.num_point = h.num_point; s.rect_lights = h.rect; s.num_rect = h.num_rect;
    s.ltc_m0 = h.ltc_m0; s.ltc_m1 = h.ltc_m1; s.ltc_m2 = h.ltc_m2; s.ltc_norm = h.ltc_norm;
    s.sky_faces = reinterpret_cast<const float4*>(h.sky_faces); s.sky_size = h.sky_size;
    memcpy(s.sky_transform, h.sky_transform, 36); memcpy(s.sky_color, h.sky_color, 12);
    s.ibl_enabled = g_ibl.enabled ? 1u : 0u; s.ibl_diffuse_size = g_ibl.desc.diffuse_size; s.ibl_specular_size = g_ibl.desc.specular_size;
    s.ibl_specular_levels = g_ibl.desc.specular_levels; s.ibl_brdf_size = g_ibl.desc.brdf_lut_size;
    s.ibl_diffuse = g_ibl.diffuse.data(); s.ibl_specular = g_ibl.specular.data(); s.ibl_brdf = g_ibl.brdf.data();
    for (int k = 0; k < 3; k++) { s.ibl_diffuse_color[k] = h.sky_color[k] * g_ibl.desc.diffuse_strength; s.ibl_specular_color[k] = h.sky_color[k] * g_ibl.desc.specular_strength; }
    s.ddgi_enabled = g_ddgi.enabled ? 1u : 0u; s.ddgi_irr_size = g_ddgi.irr_size; s.ddgi_vis_size = g_ddgi.vis_size; s.ddgi_volume = g_ddgi.vol;
    s.ddgi_irradiance = reinterpret_cast<const float4*>(g_ddgi.irr.data()); s.ddgi_visibility = reinterpret_cast<const float2*>(g_ddgi.vis.data());
}

struct HostSink {
    const DScene& sc; uint32_t nee_mode; uint32_t frame_index; float* px;
    std::vector<float3> pending;
    void add(float3 c) { px[0] += c.x; px[1] += c.y; px[2] += c.z; }
    void shadow(float3 P, float3 L, float tmax, float3 c, uint32_t) {
        if (nee_mode == BPT_NEE_NONE) { add(c); return; }
        // the connect kernel runs after the shade kernel: its contributions land after the immediate ones
        TraceResult r = trace_ray<true>(sc, P, L, 0.001f, tmax, frame_index);
        if (!r.hit) pending.push_back(c);
    }
};
} // namespace

extern "C" {

// Runs the wavefront logic (raygen → extend → shade → connect) for every pixel, one path at a time.
__attribute__((visibility("default")))
int hc_render(const hc_scene* h, const bpt_camera* cam, uint32_t width, uint32_t height, uint32_t frame_first, uint32_t nsamples,
              const bpt_settings* st, float* accum_rgba) {
    Built b; build(*h, b);
    ShadeParams sp; sp.width = width; sp.height = height; sp.max_bounces = std::min(std::max(st->max_bounces, 2u), 16u); sp.nee_mode = st->nee_mode; sp.ray_length = st->ray_length; sp.diffuse_only = 0; sp.russian_roulette = st->russian_roulette; sp.rect_shadow = st->rect_shadow;
    sp.state_precision = st->state_precision; sp.ibl = 0;
    const bool fp16 = st->state_precision == BPT_STATE_REFERENCE_FP16;      // driven like k_raygen / k_commit_bounce / k_accumulate_fp16 drive it
    for (uint32_t s = 0; s < nsamples; s++)
        for (uint32_t p = 0; p < width * height; p++) {
            float3 O, D, W = v3s(1.0f);
            float color[4] = {0, 0, 0, 0};           // per-sample colour, added to the sum when the sample ends
            camera_ray(*cam, p % width, p / width, width, height, st->pixel_jitter, frame_first + s, O, D);
            if (fp16) D = q_half3(D);
            for (uint32_t i = 1; i < sp.max_bounces; i++) {
                TraceResult r = trace_ray<false>(b.sc, O, D, 0.001f, sp.ray_length, frame_first + s);
                float bsum[4] = {0, 0, 0, 0};
                HostSink sink{b.sc, st->nee_mode, frame_first + s, fp16 ? bsum : color, {}};
                float3 nO, nD, nW;
                bool cont = shade_vertex(b.sc, sp, frame_first + s, i, p, O, D, W, r, sink, nO, nD, nW);
                for (auto& c : sink.pending) sink.add(c);
                if (fp16) {
                    float3 c = commit_bounce_fp16(v3(color[0], color[1], color[2]), v3(bsum[0], bsum[1], bsum[2]), W, i);
                    color[0] = c.x; color[1] = c.y; color[2] = c.z;
                }
                if (!cont) break;
                O = nO; D = nD; W = nW;
            }
            float* px = accum_rgba + 4ull * p;
            if (fp16) {
                float3 img = accumulate_fp16(v3(px[0], px[1], px[2]), v3(color[0], color[1], color[2]), s + 1);
                px[0] = img.x; px[1] = img.y; px[2] = img.z;
            } else for (int k = 0; k < 3; k++) px[k] += color[k];
        }
    return 0;
}

// Primary-hit outputs and RTAO: the per-thread functions of bpt_aov.cuh driven like k_primary_aov / k_ao_raygen + connect + k_ao_finish.
__attribute__((visibility("default")))
int hc_render_primary(const hc_scene* h, const bpt_camera* cam, uint32_t width, uint32_t height, uint32_t frame_index, const bpt_settings* st, float* depth, bpt_gbuffer_texel* g) {
    Built b; build(*h, b);
    for (uint32_t p = 0; p < width * height; p++) {
        float3 O, D;
        camera_ray(*cam, p % width, p / width, width, height, st->pixel_jitter, frame_index, O, D);
        if (st->state_precision == BPT_STATE_REFERENCE_FP16) D = q_half3(D);
        TraceResult r = trace_ray<false>(b.sc, O, D, 0.001f, st->ray_length, frame_index);
        primary_outputs(b.sc, *cam, O, D, r, depth[p], g[p]);
    }
    return 0;
}
__attribute__((visibility("default")))
int hc_trace_ao(const hc_scene* h, const bpt_camera* cam, uint32_t width, uint32_t height, uint32_t frame_index, const bpt_ao_settings* ao, const float* depth,
                const float* normal_roughness, float* out) {
    Built b; build(*h, b);
    const uint32_t aw = ao->half_resolution ? width / 2 : width, ah = ao->half_resolution ? height / 2 : height;
    const float range = ao->range > 0.05f ? ao->range : 0.05f;
    for (uint32_t p = 0; p < aw * ah; p++) {
        float3 origin, dirs[4];
        if (!ao_pixel_rays(*cam, p % aw, p / aw, aw, ah, width, height, frame_index, ao->half_resolution, depth, reinterpret_cast<const float4*>(normal_roughness), origin, dirs)) {
            out[2 * p] = 1.0f; out[2 * p + 1] = 0.0f; continue;
        }
        uint32_t unoccluded = 0;
        for (int i = 0; i < 4; i++) unoccluded += trace_ray<true>(b.sc, origin, dirs[i], 0.001f, range, frame_index, true).hit ? 0u : 1u;
        out[2 * p] = q_half(ao_value(4u - unoccluded, ao->strength)); out[2 * p + 1] = 1.0f;
    }
    return 0;
}

// Ray-traced reflections (bpt_trace_reflection): rtr_pixel_ray, then one bounce of the same trace / shade functions.
__attribute__((visibility("default")))
int hc_trace_reflection(const hc_scene* h, const bpt_camera* cam, uint32_t width, uint32_t height, uint32_t frame_index, const bpt_reflection_settings* rs,
                        const float* depth, const bpt_gbuffer_texel* gbuffer, float* out_refl, float* out_hit) {
    Built b; build(*h, b);
    const uint32_t rw = rs->half_resolution ? (width + 1) / 2 : width, rh = rs->half_resolution ? (height + 1) / 2 : height;
    const float max_roughness = rs->max_roughness, fade_roughness = std::min(rs->fade_roughness, max_roughness - 0.0001f);
    ShadeParams sp; sp.width = width; sp.height = height; sp.max_bounces = 2; sp.nee_mode = BPT_NEE_SHADOW_RAY;
    sp.ray_length = rs->range; sp.diffuse_only = 0; sp.russian_roulette = 0; sp.rect_shadow = 0; sp.state_precision = BPT_STATE_FP32; sp.ibl = rs->ibl;
    for (uint32_t p = 0; p < rw * rh; p++) {
        float color[4] = {0, 0, 0, -1.0f};
        float hp[4] = {0, 0, 0, -1.0f};
        float3 O, D, W;
        if (rtr_pixel_ray(*cam, p % rw, p / rw, rw, rh, width, height, frame_index, rs->half_resolution, depth, gbuffer, max_roughness, fade_roughness, O, D, W)) {
            W = W * rs->strength;
            TraceResult r = trace_ray<false>(b.sc, O, D, 0.001f, sp.ray_length, frame_index);
            HostSink sink{b.sc, BPT_NEE_SHADOW_RAY, frame_index, color, {}};
            float3 nO, nD, nW;
            shade_vertex<HostSink, true>(b.sc, sp, frame_index, 1u, p, O, D, W, r, sink, nO, nD, nW);
            for (auto& c : sink.pending) sink.add(c);
            if (r.hit) { float3 P = O + D * r.t; hp[0] = P.x; hp[1] = P.y; hp[2] = P.z; hp[3] = r.t; }
            else { hp[0] = D.x; hp[1] = D.y; hp[2] = D.z; hp[3] = -1.0f; }
        }
        out_refl[4 * p] = color[0]; out_refl[4 * p + 1] = color[1]; out_refl[4 * p + 2] = color[2]; out_refl[4 * p + 3] = 1.0f;
        for (int k = 0; k < 4; k++) out_hit[4 * p + k] = hp[k];
    }
    return 0;
}

// Probe paths (bpt_trace_probes): same loop, rays start at probe centres, diffuse-only surface.
__attribute__((visibility("default")))
int hc_trace_probes(const hc_scene* h, const bpt_probe_volume* vol, const float* table, uint32_t frame_index, uint32_t num_bounces, float* out) {
    Built b; build(*h, b);
    ShadeParams sp; sp.width = 0; sp.height = 0; sp.max_bounces = std::min(std::max(num_bounces, 1u), 15u) + 1; sp.nee_mode = BPT_NEE_SHADOW_RAY;
    sp.ray_length = vol->ray_length; sp.diffuse_only = 1; sp.russian_roulette = 0; sp.rect_shadow = 0; sp.state_precision = BPT_STATE_FP32; sp.ibl = 0;
    uint64_t total = (uint64_t)vol->probe_counts[0] * vol->probe_counts[1] * vol->probe_counts[2] * vol->rays_per_probe;
    for (uint64_t p = 0; p < total; p++) {
        float3 O, D, W = v3s(1.0f);
        probe_ray(*vol, reinterpret_cast<const float2*>(table), (uint32_t)p, frame_index, O, D);
        float color[4] = {0, 0, 0, -1.0f};
        for (uint32_t i = 1; i < sp.max_bounces; i++) {
            TraceResult r = trace_ray<false>(b.sc, O, D, 0.001f, sp.ray_length, frame_index);
            if (i == 1) color[3] = r.t;
            HostSink sink{b.sc, BPT_NEE_SHADOW_RAY, frame_index, color, {}};
            float3 nO, nD, nW;
            bool cont = shade_vertex(b.sc, sp, frame_index, i, (uint32_t)p, O, D, W, r, sink, nO, nD, nW);
            for (auto& c : sink.pending) sink.add(c);
            if (!cont) break;
            O = nO; D = nD; W = nW;
        }
        for (int k = 0; k < 4; k++) out[4 * p + k] = color[k];
    }
    return 0;
}

__attribute__((visibility("default")))
int hc_trace(const hc_scene* h, const bpt_ray* rays, uint64_t n, uint32_t frame_index, bpt_hit* hits, uint8_t* visible) {
    Built b; build(*h, b);
    for (uint64_t i = 0; i < n; i++) {
        float3 O = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), D = v3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        if (hits) {
            TraceResult t = trace_ray<false>(b.sc, O, D, rays[i].tmin, rays[i].tmax, frame_index);
            hits[i] = bpt_hit{t.t, t.u, t.v, t.hit ? b.inst[t.slot].instance_id : 0xffffffffu, t.hit ? t.prim : 0xffffffffu};
        }
        if (visible) visible[i] = trace_ray<true>(b.sc, O, D, rays[i].tmin, rays[i].tmax, frame_index).hit ? 0 : 1;
    }
    return 0;
}

// 4-wide quantised tree of the merged BVH: collapse_node4 / leaf_boxes_of_node per binary node like k_collapse4, then the
// run-to-completion wide traversal (trace_ray_wide) that the persistent kernels interleave.
static void build_wide(const hc_scene& h, std::vector<float4>& wide, std::vector<float4>& leafbox) { collapse_tree(h.blas_bvh[0], wide, leafbox); }
__attribute__((visibility("default")))
int hc_read_wide(const hc_scene* h, float* wide_out, float* leafbox_out) {
    if (h->accel_mode != BPT_ACCEL_MERGED) return 1;
    std::vector<float4> wide, leafbox;
    build_wide(*h, wide, leafbox);
    if (wide_out) memcpy(wide_out, wide.data(), wide.size() * 16);
    if (leafbox_out) memcpy(leafbox_out, leafbox.data(), leafbox.size() * 16);
    return 0;
}
__attribute__((visibility("default")))
int hc_trace_wide(const hc_scene* h, const bpt_ray* rays, uint64_t n, uint32_t frame_index, bpt_hit* hits, uint8_t* visible) {
    Built b; build(*h, b);
    if (h->accel_mode == BPT_ACCEL_TWO_LEVEL) {          // wide TLAS + wide BLASes (build() collapsed them)
        for (uint64_t i = 0; i < n; i++) {
            float3 O = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), D = v3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
            if (hits) {
                TraceResult t = trace_ray_wide_two_level<false>(b.sc, O, D, rays[i].tmin, rays[i].tmax, frame_index);
                hits[i] = bpt_hit{t.t, t.u, t.v, t.hit ? b.inst[t.slot].instance_id : 0xffffffffu, t.hit ? t.prim : 0xffffffffu};
            }
            if (visible) visible[i] = trace_ray_wide_two_level<true>(b.sc, O, D, rays[i].tmin, rays[i].tmax, frame_index).hit ? 0 : 1;
        }
        return 0;
    }
    std::vector<float4> wide, leafbox;
    build_wide(*h, wide, leafbox);
    for (uint64_t i = 0; i < n; i++) {
        float3 O = v3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), D = v3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        if (hits) {
            TraceResult t = trace_ray_wide<false>(b.sc, wide.data(), leafbox.data(), O, D, rays[i].tmin, rays[i].tmax, frame_index);
            hits[i] = bpt_hit{t.t, t.u, t.v, t.hit ? b.inst[t.slot].instance_id : 0xffffffffu, t.hit ? t.prim : 0xffffffffu};
        }
        if (visible) visible[i] = trace_ray_wide<true>(b.sc, wide.data(), leafbox.data(), O, D, rays[i].tmin, rays[i].tmax, frame_index).hit ? 0 : 1;
    }
    return 0;
}

// Probe blending: the per-texel device functions of bpt_ddgi.cuh driven like k_probe_blend drives them.
__attribute__((visibility("default")))
int hc_blend_probes(const bpt_probe_volume* vol, const float* table, uint32_t frame_index, const float* rays, const bpt_probe_blend* bl, float* irr, float* vis) {
    const uint32_t nx = vol->probe_counts[0], ny = vol->probe_counts[1], nz = vol->probe_counts[2], nr = vol->rays_per_probe;
    std::vector<float3> dirs(nr); std::vector<float4> rad(nr);
    for (uint32_t probe = 0; probe < nx * ny * nz; probe++) {
        for (uint32_t r = 0; r < nr; r++) {
            float3 O, D;
            probe_ray(*vol, reinterpret_cast<const float2*>(table), probe * nr + r, frame_index, O, D);
            const float* p = rays + 4ull * (probe * nr + r);
            rad[r] = make_float4(p[0], p[1], p[2], p[3]);
            dirs[r] = blend_trace_dir(O, D, p[3]);
        }
        const uint32_t ix = probe % nx, iy = (probe / nx) % ny, iz = probe / nx / ny;
        for (int pass = 0; pass < 2; pass++) {
            const bool visp = pass == 1;
            const uint32_t size = visp ? bl->visibility_size : bl->irradiance_size, ch = visp ? 2 : 4;
            const uint32_t stride = nx * ny * (size + 2), sx = (iy * nx + ix) * (size + 2), sy = iz * (size + 2);
            float* atlas = visp ? vis : irr;
            auto at = [&](uint32_t x, uint32_t y) { return atlas + ((size_t)(sy + y) * stride + (sx + x)) * ch; };
            std::vector<float3> vals(size * size);
            for (uint32_t t = 0; t < size * size; t++) {
                uint32_t tx = t % size, ty = t / size;
                float3 v = visp ? blend_texel<true>(tx, ty, size, dirs.data(), rad.data(), nr) : blend_texel<false>(tx, ty, size, dirs.data(), rad.data(), nr);
                if (bl->history_valid) {
                    const float* h = at(tx + 1, ty + 1);
                    v.x = temporal_blend(v.x, h[0], bl->alpha); v.y = temporal_blend(v.y, h[1], bl->alpha);
                    if (!visp) v.z = temporal_blend(v.z, h[2], bl->alpha);
                }
                vals[t] = v;
            }
            for (uint32_t t = 0; t < size * size; t++) {
                uint32_t cx = t % size + 1, cy = t / size + 1, bx, by, c[4];
                float3 v = vals[t];
                auto put = [&](uint32_t x, uint32_t y) { float* o = at(x, y); o[0] = v.x; o[1] = v.y; if (!visp) { o[2] = v.z; o[3] = 1.0f; } };
                put(cx, cy); border_coord(cx, cy, size, bx, by); put(bx, by);
                if (corner_coords(cx, cy, size, c)) { put(c[0], c[1]); put(c[2], c[3]); }
            }
        }
    }
    return 0;
}

__attribute__((visibility("default")))
void hc_set_ddgi(const bpt_probe_volume* vol, const bpt_probe_blend* bl, const float* irr, const float* vis) {
    g_ddgi.enabled = vol && bl && irr && vis;
    if (!g_ddgi.enabled) return;
    const size_t nx = vol->probe_counts[0], ny = vol->probe_counts[1], nz = vol->probe_counts[2];
    g_ddgi.vol = *vol; g_ddgi.irr_size = bl->irradiance_size; g_ddgi.vis_size = bl->visibility_size;
    g_ddgi.irr.assign(irr, irr + nx * ny * (bl->irradiance_size + 2) * nz * (bl->irradiance_size + 2) * 4);
    g_ddgi.vis.assign(vis, vis + nx * ny * (bl->visibility_size + 2) * nz * (bl->visibility_size + 2) * 2);
}
__attribute__((visibility("default")))
void hc_ddgi_lighting(uint64_t n, const float* pos, const float* normal, const float* view, float* out) {
    for (uint64_t i = 0; i < n; i++) {
        float4 r = ddgi_volume_lighting(g_ddgi.vol, g_ddgi.irr_size, g_ddgi.vis_size, reinterpret_cast<const float4*>(g_ddgi.irr.data()),
                                        reinterpret_cast<const float2*>(g_ddgi.vis.data()), v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]),
                                        v3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]), v3(view[3 * i], view[3 * i + 1], view[3 * i + 2]));
        out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
    }
}
__attribute__((visibility("default"))) float hc_q_half(float f) { return q_half(f); }
__attribute__((visibility("default"))) void hc_surface_through_gbuffer(const float N[3], const float T[3], const float in[12], uint32_t model, float out[18], uint32_t* model_out) {
    Surface s = surface_default();
    s.base_color = v3(in[0], in[1], in[2]); s.f0_color = v3(in[3], in[4], in[5]); s.f90_color = v3(in[6], in[7], in[8]);
    s.roughness = in[9]; s.anisotropy = in[10]; s.ior = in[11];
    float3 No, To;
    frame_through_gbuffer(v3(N[0], N[1], N[2]), v3(T[0], T[1], T[2]), No, To);
    surface_through_gbuffer(s, model);
    const float o[18] = {No.x, No.y, No.z, To.x, To.y, To.z, s.base_color.x, s.base_color.y, s.base_color.z, s.f0_color.x, s.f0_color.y, s.f0_color.z,
                         s.f90_color.x, s.f90_color.y, s.f90_color.z, s.roughness, s.anisotropy, s.ior};
    memcpy(out, o, sizeof(o)); *model_out = model;
}
__attribute__((visibility("default"))) uint32_t hc_rng_tea(uint32_t a, uint32_t b) { return rng_tea(a, b); }
__attribute__((visibility("default"))) void hc_sincos_2pi(float u, float* s, float* c) { sincos_2pi(u, *s, *c); }
__attribute__((visibility("default"))) float hc_atan2(float y, float x) { return atan2_(y, x); }
__attribute__((visibility("default"))) float hc_acos(float x) { return acos_(x); }
__attribute__((visibility("default"))) void hc_ggx_vndf_sample(const float v[3], float rx, float ry, float u1, float u2, float o[3]) {
    float3 h = ggx_vndf_sample(v3(v[0], v[1], v[2]), rx, ry, u1, u2); o[0] = h.x; o[1] = h.y; o[2] = h.z;
}
__attribute__((visibility("default"))) void hc_surface_eval_lit(const float N[3], const float T[3], const float V[3], const float L[3], const float base[3],
                                                                 const float f0[3], const float f90[3], float roughness, float anisotropy, float out[3]) {
    Surface s = surface_default();
    s.base_color = v3(base[0], base[1], base[2]); s.f0_color = v3(f0[0], f0[1], f0[2]); s.f90_color = v3(f90[0], f90[1], f90[2]);
    s.roughness = roughness; s.anisotropy = anisotropy;
    float3 n = v3(N[0], N[1], N[2]), t = v3(T[0], T[1], T[2]);
    float3 r = bsdf_eval(n, t, cross3(n, t), v3(V[0], V[1], V[2]), v3(L[0], L[1], L[2]), s, 1u);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

} // extern "C"

// Post-process (bloom + output): the per-pixel device functions of bpt_post.cuh driven like post.cu's kernels drive them
// (level kernel = horizontal pass into a tile, vertical pass out of it; here the "tile" is the whole target).
namespace {
struct HostTex {
    const std::vector<float3>* px; int w;
    float3 at(int x, int y) const { return (*px)[(size_t)y * w + x]; }
};
struct HostPre {
    const float* in; int w; float inv; BloomWeights bw;
    float3 at(int x, int y) const { const float* s = in + ((size_t)y * w + x) * 4; return bloom_pre(v3(s[0] * inv, s[1] * inv, s[2] * inv), bw); }
};
template <class Src>
std::vector<float3> bloom_level(const Src& src, int sw, int sh, int dw, int dh) {
    std::vector<float3> hpass((size_t)dw * dh), out((size_t)dw * dh);
    for (int y = 0; y < dh; y++) for (int x = 0; x < dw; x++) hpass[(size_t)y * dw + x] = bloom_horizontal(src, sw, sh, x, y, dw, dh);
    HostTex ht{&hpass, dw};
    for (int y = 0; y < dh; y++) for (int x = 0; x < dw; x++) out[(size_t)y * dw + x] = bloom_vertical(ht, x, y, dw, dh);
    return out;
}
}
extern "C" __attribute__((visibility("default")))
int hc_post_process(const float* sums_rgba32f, uint32_t width, uint32_t height, uint32_t total_samples, const bpt_post_settings* st, float* out) {
    const int W = (int)width, H = (int)height;
    const float inv = 1.0f / (float)total_samples;
    std::vector<float3> bloom; int bw_ = 1, bh_ = 1;
    if (st->bloom) {
        int lw[3], lh[3];
        for (int i = 0; i < 3; i++) { lw[i] = std::max(W >> (i + 1), 1); lh[i] = std::max(H >> (i + 1), 1); }
        auto v1 = bloom_level(HostPre{sums_rgba32f, W, inv, bloom_weights(st->bloom_threshold, st->bloom_threshold_softness)}, W, H, lw[0], lh[0]);
        auto v2 = bloom_level(HostTex{&v1, lw[0]}, lw[0], lh[0], lw[1], lh[1]);
        auto v3_ = bloom_level(HostTex{&v2, lw[1]}, lw[1], lh[1], lw[2], lh[2]);
        std::vector<float3> c2((size_t)lw[1] * lh[1]), c1((size_t)lw[0] * lh[0]);
        for (int y = 0; y < lh[1]; y++) for (int x = 0; x < lw[1]; x++)
            c2[(size_t)y * lw[1] + x] = bloom_combine(v2[(size_t)y * lw[1] + x], HostTex{&v3_, lw[2]}, lw[2], lh[2], x, y, lw[1], lh[1]);
        for (int y = 0; y < lh[0]; y++) for (int x = 0; x < lw[0]; x++)
            c1[(size_t)y * lw[0] + x] = bloom_combine(v1[(size_t)y * lw[0] + x], HostTex{&c2, lw[1]}, lw[1], lh[1], x, y, lw[0], lh[0]);
        bloom = std::move(c1); bw_ = lw[0]; bh_ = lh[0];
    }
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const float* s = sums_rgba32f + ((size_t)y * W + x) * 4;
        float3 c = v3(s[0] * inv, s[1] * inv, s[2] * inv);
        if (st->bloom) c = bloom_combine(c, HostTex{&bloom, bw_}, bw_, bh_, x, y, W, H);
        float* o = out + ((size_t)y * W + x) * 4;
        o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = 1.0f;
    }
    return 0;
}

// Sky IBL precompute: the per-texel functions of bpt_ibl.cuh driven like ibl.cu's kernels (one call per texel). desc == NULL unbinds.
extern "C" __attribute__((visibility("default")))
int hc_precompute_sky_ibl(const float* sky_faces, uint32_t sky_size, const bpt_sky_ibl_desc* d, float* diffuse, float* specular, float* brdf) {
    g_ibl.enabled = false;
    if (!d) return 0;
    const float4* sky = reinterpret_cast<const float4*>(sky_faces);
    g_ibl.desc = *d;
    g_ibl.brdf.resize((size_t)d->brdf_lut_size * d->brdf_lut_size);
    for (uint32_t i = 0; i < d->brdf_lut_size * d->brdf_lut_size; i++) g_ibl.brdf[i] = ibl_brdf_lut_texel(i % d->brdf_lut_size, i / d->brdf_lut_size, d->brdf_lut_size);
    const uint32_t DS = d->diffuse_size;
    g_ibl.diffuse.resize((size_t)6 * DS * DS);
    for (uint32_t i = 0; i < 6 * DS * DS; i++) {
        uint32_t r = i % (DS * DS);
        float3 c = ibl_diffuse_texel(sky, sky_size, r % DS, r / DS, i / (DS * DS), DS);
        g_ibl.diffuse[i] = make_float4(c.x, c.y, c.z, 1.0f);
    }
    g_ibl.specular.clear();
    for (uint32_t l = 0; l < d->specular_levels; l++) {
        const uint32_t S = d->specular_size >> l;
        for (uint32_t i = 0; i < 6 * S * S; i++) {
            uint32_t r = i % (S * S);
            float3 c = ibl_specular_texel(sky, sky_size, r % S, r / S, i / (S * S), S, (float)l / (float)(d->specular_levels - 1));
            g_ibl.specular.push_back(make_float4(c.x, c.y, c.z, 1.0f));
        }
    }
    if (diffuse) memcpy(diffuse, g_ibl.diffuse.data(), g_ibl.diffuse.size() * 16);
    if (specular) memcpy(specular, g_ibl.specular.data(), g_ibl.specular.size() * 16);
    if (brdf) memcpy(brdf, g_ibl.brdf.data(), g_ibl.brdf.size() * 8);
    g_ibl.enabled = true;
    return 0;
}

extern "C" __attribute__((visibility("default")))
int hc_upscale_half_res(const bpt_camera* cam, uint32_t W, uint32_t H, uint32_t frame_index, const float* depth, const float* nr, const float* in_half, float* out) {
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            float4 v = upscale_pixel(*cam, x, y, W, H, frame_index, depth, reinterpret_cast<const float4*>(nr), reinterpret_cast<const float4*>(in_half));
            float* o = out + 4 * ((size_t)y * W + x);
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        }
    return 0;
}


// ---- ReBLUR: the per-pixel pass functions of bpt_reblur.cuh run on the host in the order launch_reblur (reblur.cu) launches them ----
namespace {
struct HcReblur {
    bool has_history = false; uint64_t last_frame = 0; bpt_camera last_cam{}; uint32_t w = 0, h = 0, gw = 0, gh = 0;
    std::vector<float4> ld0, ld1, denoised, hist_ld0, hist_ld1, nr[2], noised, hit;
    std::vector<float> accum, lin_depth, hist_accum, depth[2];
    std::vector<float2> velocity; std::vector<uint8_t> validation;
    int cur = 0;
} g_rb;
const float kHcPre[32] = {0.840188f, 0.394383f, 0.783099f, 0.79844f, 0.911647f, 0.197551f, 0.335223f, 0.76823f, 0.277775f, 0.55397f, 0.477397f, 0.628871f, 0.364784f, 0.513401f,
                          0.95223f, 0.916195f, 0.635712f, 0.717297f, 0.141603f, 0.606969f, 0.0163006f, 0.242887f, 0.137232f, 0.804177f, 0.156679f, 0.400944f, 0.12979f, 0.108809f,
                          0.998924f, 0.218257f, 0.512932f, 0.839112f};
const float kHcBlur[32] = {0.61264f, 0.296032f, 0.637552f, 0.524287f, 0.493583f, 0.972775f, 0.292517f, 0.771358f, 0.526745f, 0.769914f, 0.400229f, 0.891529f, 0.283315f, 0.352458f,
                           0.807725f, 0.919026f, 0.0697553f, 0.949327f, 0.525995f, 0.0860558f, 0.192214f, 0.663227f, 0.890233f, 0.348893f, 0.0641713f, 0.020023f, 0.457702f,
                           0.0630958f, 0.23828f, 0.970634f, 0.902208f, 0.85092f};
const float kHcPost[32] = {0.266666f, 0.53976f, 0.375207f, 0.760249f, 0.512535f, 0.667724f, 0.531606f, 0.0392803f, 0.437638f, 0.931835f, 0.93081f, 0.720952f, 0.284293f, 0.738534f,
                           0.639979f, 0.354049f, 0.687861f, 0.165974f, 0.440105f, 0.880075f, 0.829201f, 0.330337f, 0.228968f, 0.893372f, 0.35036f, 0.68667f, 0.956468f, 0.58864f,
                           0.657304f, 0.858676f, 0.43956f, 0.92397f};
float4 hc_rotator(float a) { float ca = std::cos(a), sa = std::sin(a); return make_float4(ca, sa, -sa, ca); }
} // namespace

extern "C" __attribute__((visibility("default"))) void hc_reblur_reset() { g_rb.has_history = false; }

extern "C" __attribute__((visibility("default")))
int hc_reblur(const bpt_camera* cam, uint64_t frame_count, const bpt_reblur_settings* st, const bpt_reblur_inputs* in, uint32_t gw, uint32_t gh, float* out) {
    HcReblur& r = g_rb;
    const uint32_t w = in->width, h = in->height;
    const size_t n = (size_t)w * h, gn = (size_t)gw * gh, chain = reblur_mip_offset(w, h, 4);
    if (r.w != w || r.h != h || r.gw != gw || r.gh != gh) { r.has_history = false; r.w = w; r.h = h; r.gw = gw; r.gh = gh; }
    r.ld0.resize(chain); r.ld1.resize(n); r.denoised.resize(n); r.hist_ld0.resize(n); r.hist_ld1.resize(n); r.accum.resize(n); r.lin_depth.resize(chain); r.hist_accum.resize(n);
    for (int k = 0; k < 2; k++) { r.depth[k].resize(gn); r.nr[k].resize(gn); }
    const bool has_history = r.has_history && r.last_frame + 1 == frame_count;
    const int cur = r.cur ^ 1;
    memcpy(r.depth[cur].data(), in->depth, gn * 4); memcpy(r.nr[cur].data(), in->normal_roughness, gn * 16);
    r.noised.assign(reinterpret_cast<const float4*>(in->noised), reinterpret_cast<const float4*>(in->noised) + n);
    r.hit.assign(reinterpret_cast<const float4*>(in->hit_positions), reinterpret_cast<const float4*>(in->hit_positions) + n);
    if (in->velocity) r.velocity.assign(reinterpret_cast<const float2*>(in->velocity), reinterpret_cast<const float2*>(in->velocity) + gn);
    if (in->history_validation) r.validation.assign(in->history_validation, in->history_validation + n);
    ReblurView rv{};
    rv.w = w; rv.h = h; rv.gw = gw; rv.gh = gh; rv.half_res = w != gw ? 1u : 0u; rv.frame_index = (uint32_t)frame_count;
    rv.has_history = has_history ? 1u : 0u; rv.virtual_history = st->virtual_history; rv.blur_radius = st->blur_radius; rv.anti_flicker = st->anti_flickering_strength;
    rv.cam = *cam; rv.hist_cam = has_history ? r.last_cam : *cam;
    const uint32_t ri = (uint32_t)(frame_count % 32);
    rv.rot_pre = hc_rotator(kHcPre[ri]); rv.rot_blur = hc_rotator(kHcBlur[ri]); rv.rot_post = hc_rotator(kHcPost[ri]);
    rv.depth = r.depth[cur].data(); rv.normal_roughness = r.nr[cur].data();
    rv.velocity = in->velocity ? r.velocity.data() : nullptr; rv.validation = in->history_validation ? r.validation.data() : nullptr;
    rv.hit_positions = r.hit.data(); rv.noised = r.noised.data();
    rv.hist_depth = r.depth[cur ^ 1].data(); rv.hist_normal_roughness = r.nr[cur ^ 1].data();
    rv.hist_ld0 = r.hist_ld0.data(); rv.hist_ld1 = r.hist_ld1.data(); rv.hist_accum = r.hist_accum.data();
    rv.ld0 = r.ld0.data(); rv.ld1 = r.ld1.data(); rv.accum = r.accum.data(); rv.lin_depth = r.lin_depth.data(); rv.denoised = r.denoised.data();
    auto each = [&](auto fn) { for (int y = 0; y < (int)h; y++) for (int x = 0; x < (int)w; x++) fn(rv, x, y); };
    each(reblur_pre_blur);
    each(reblur_temporal_accumulate);
    each(reblur_fetch_linear_depth);
    for (int ty = 0; ty < (int)((h + 15) / 16); ty++)                       // k_reblur_gen_depth_mip: one 64-thread group per tile
        for (int tx = 0; tx < (int)((w + 15) / 16); tx++) {
            float4 sv[64]; float sd[64]; int px[64], py[64];
            for (uint32_t l = 0; l < 64; l++) { rb_mip_level1(rv, tx, ty, l, sv[l], sd[l], px[l], py[l]); rb_mip_store(rv, 1, px[l] >> 1, py[l] >> 1, sv[l], sd[l]); }
            for (uint32_t l = 0; l < 64; l += 4) {
                float4 vv[4] = {sv[l], sv[l + 1], sv[l + 2], sv[l + 3]}; float dd[4] = {sd[l], sd[l + 1], sd[l + 2], sd[l + 3]};
                rb_mip_reduce(vv, dd, sv[l], sd[l]);
                rb_mip_store(rv, 2, px[l] >> 2, py[l] >> 2, sv[l], sd[l]);
            }
            for (uint32_t l = 0; l < 64; l += 16) {
                float4 vv[4] = {sv[l], sv[l + 4], sv[l + 8], sv[l + 12]}; float dd[4] = {sd[l], sd[l + 4], sd[l + 8], sd[l + 12]};
                float4 v; float d;
                rb_mip_reduce(vv, dd, v, d);
                rb_mip_store(rv, 3, px[l] >> 3, py[l] >> 3, v, d);
            }
        }
    each(reblur_fix_history);
    each(reblur_blur);
    memcpy(r.hist_ld0.data(), r.ld0.data(), n * 16); r.hist_accum = r.accum;
    each(reblur_temporal_stabilize);
    r.hist_ld1 = r.ld1;
    each(reblur_post_blur);
    memcpy(out, r.denoised.data(), n * 16);
    r.has_history = true; r.last_frame = frame_count; r.last_cam = *cam; r.cur = cur;
    return 0;
}
extern "C" __attribute__((visibility("default")))
int hc_reblur_read(uint32_t which, float* out) {
    HcReblur& r = g_rb;
    const size_t n = (size_t)r.w * r.h;
    if (which == 0) memcpy(out, r.ld0.data(), r.ld0.size() * 16);
    else if (which == 1) memcpy(out, r.ld1.data(), n * 16);
    else if (which == 2) memcpy(out, r.accum.data(), n * 4);
    else if (which == 3) memcpy(out, r.lin_depth.data(), r.lin_depth.size() * 4);
    else return -1;
    return 0;
}

// exhaustive check hooks for the exact-arithmetic shortcuts of bpt_scene.cuh
extern "C" __attribute__((visibility("default"))) float hc_unorm8_to_float(uint32_t k) { return unorm8_to_float(k); }
extern "C" __attribute__((visibility("default"))) void hc_wrap_tc2(int c, int n, uint32_t mode, int* a, int* b) { wrap_tc2(c, n, mode, *a, *b); }
extern "C" __attribute__((visibility("default"))) int hc_wrap_tc(int c, int n, uint32_t mode) { return wrap_tc(c, n, mode); }
