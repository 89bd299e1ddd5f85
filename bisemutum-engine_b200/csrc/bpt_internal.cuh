// bpt_internal.cuh — host-side context of libbpt.so (not part of the ABI).
#pragma once
#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>
#include <string>
#include <utility>
#include <vector>
#include "bpt_scene.cuh"

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One built BVH (BLAS or TLAS) in device memory.
struct DevBvh {
    uint32_t n = 0;
    int32_t root = 0;
    DevBuf nodes;      // float4[4*(n-1)]
    DevBuf tris;       // float4[3*n] (BLAS only)
    DevBuf wide;       // float4[4*(n-1)]: 4-wide quantised nodes (merged BLAS only, bpt_wide.cuh); node i = subtree of binary node i
    DevBuf leafbox;    // float4[2*n]: exact box of every leaf (merged BLAS only)
    DevBuf morton;     // uint64[n] sorted
    DevBuf prims;      // uint32[n] sorted primitive ids
};

struct WavefrontState {     // per-path SoA, two ray buffers (ping-pong through compaction)
    uint64_t capacity = 0;  // paths in flight = slots * width * height
    uint32_t npx = 0, slots = 1;
    DevBuf color;           // float4 per path: per-sample colour C_s
    uint32_t ahead_slots = 0, ahead_cursor = 0, ahead_frame_first = 0;   // prefetched samples still in `color`
    std::vector<std::pair<const void*, unsigned>> grids;   // resident grid size per persistent traversal kernel (render.cu: resident_grid)
    DevBuf ray_o[2], ray_d[2], ray_w[2];   // float4 each: (O|pixel), (D|-), (W|-)
    DevBuf hit;             // float4 (t,u,v,prim)
    DevBuf hit_slot;        // uint32
    uint64_t shadow_capacity = 0;
    DevBuf sh_o, sh_d, sh_c; // float4 each: (P|pixel), (L|tmax), (c|light)
    DevBuf accum;           // float4 per pixel: FP32 sums (reference_fp16: the running half-valued average)
    DevBuf bcol;            // float4 per path, reference_fp16 only: light terms of the current bounce before the half store
    uint32_t accum_count = 0;   // reference_fp16: samples already folded into `accum` (pt_accumulate weight = 1 / (count + 1))
    bool accum_fp16 = false;    // which of the two accumulation rules `accum` currently follows (fixed until bpt_clear_accum)
    DevBuf qcount;          // uint32[128]: extend / shadow queue lengths per bounce + work cursors of the persistent kernels
    DevBuf totals;          // uint64[40]: extend per bounce [0..15], shadow per bounce [16..31], samples [32]
};

// ReBLUR (reblur.cu): working textures, and the history the reference keeps on the camera between frames
struct ReblurState {
    uint32_t w = 0, h = 0, gw = 0, gh = 0;
    bool has_history = false; uint64_t last_frame = 0; bpt_camera last_cam{}; int cur = 0;
    DevBuf ld0, ld1, accum, lin_depth, denoised, hist_ld0, hist_ld1, hist_accum, depth[2], nr[2], velocity, validation, noised, hit;
};

struct bpt_context {
    int device = 0;
    uint32_t width = 0, height = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    void* nccl_comm = nullptr; bool nccl_owned = false;   // bpt_comm_init / bpt_comm_attach (NCCL is dlopen'ed, see bpt_api.cu)
    uint64_t wave_paths_budget = 0;   // bpt_set_wave_budget: paths in flight per wave (0 = default 2^26)
    bool accum_used = false;    // a render has added to wf.accum since the last bpt_clear_accum

    // host copies needed for validation / rebuilds
    std::vector<bpt_drawable_sbt_data> h_drawables;
    std::vector<bpt_blas_desc> h_blas_desc;
    std::vector<bpt_instance_desc> h_instances;
    std::vector<bpt_material> h_materials;
    uint64_t num_position_floats = 0, num_indices = 0;
    uint32_t num_dir = 0, num_point = 0, num_rect = 0;
    std::vector<uint8_t> h_dir_bytes, h_point_bytes, h_rect_bytes;   // the light arrays as last uploaded (prefetch validity, bpt_scene_upload_lights)
    uint64_t scene_generation = 0;                                    // bumped by every call that changes what a sample would see
    bool has_normals = false, has_tangents = false, has_texcoords = false, has_colors = false;

    // device scene
    DevBuf d_colors;
    DevBuf d_positions, d_normals, d_tangents, d_texcoords, d_indices, d_drawables, d_drawable_va, d_materials;
    DevBuf d_textures; std::vector<DevBuf> d_texels;
    DevBuf d_instances;          // DInstance[]
    DevBuf d_dir, d_point, d_rect, d_ltc[4];
    std::vector<DevBuf> d_light_texels; DevBuf d_light_textures, d_srgb_tables;     // rect-light textures (lighttex.cu)
    std::vector<bptd::DLightTexture> h_light_textures;
    DevBuf d_sky; uint32_t sky_size = 0;
    DevBuf d_ibl_diffuse, d_ibl_specular, d_ibl_brdf;      // bpt_precompute_sky_ibl (ibl.cu)
    bool ibl_valid = false; bpt_sky_ibl_desc ibl_desc{};
    DevBuf d_ddgi_irr, d_ddgi_vis;      // atlases of the bound DDGI volume (bpt_set_ddgi_volume)
    bool ddgi_enabled = false; uint32_t ddgi_irr_size = 0, ddgi_vis_size = 0; bpt_probe_volume ddgi_volume{};
    float sky_transform[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float sky_color[3] = {1, 1, 1};

    // accel
    uint64_t instanced_triangles = 0;      // sum over the instances of their BLAS's triangles (upload_instance_table): what a camera ray can see
    bool scene_has_anyhit = true;          // some instance needs the any-hit opacity rule (upload_instance_table); false selects the kernels without it
    bool accel_built = false;
    bool bloom_attr_set = false;           // the same opt-in for the bloom level kernels (post.cu)
    bool lbvh_small_attr_set = false;      // cudaFuncSetAttribute(k_lbvh_small, MaxDynamicSharedMemorySize) done on this context's device
    uint32_t accel_mode = 0;
    std::vector<DevBvh> blas;
    DevBvh tlas;
    DevBuf d_blas_table;         // DBlas[]
    DevBuf d_inst_aabb;          // scratch for TLAS build
    DevBuf d_blas_bounds;        // float[6] per BLAS (lo, hi), written by the BLAS builds, read by the TLAS build
    // build scratch: grow-only chunks, bump-allocated with stack discipline (bvh_build.cu: Scratch), so a per-frame rebuild
    // (the reference rebuilds its TLAS every frame, accel.cpp:134-159) makes no cudaMalloc / cudaFree calls
    std::vector<DevBuf> arena_chunks; size_t arena_chunk = 0, arena_offset = 0;
    bool arena_hold = false; size_t arena_held_bytes = 0;       // build_all_blas_two_level: scratch of concurrent builds is released together

    WavefrontState wf;
    ReblurState reblur;
    DevBuf d_post;               // rgba16_sfloat targets of the bloom chain (post.cu)
    DevBuf d_post_out;           // staging of bpt_post_process's host read-back
    bool profile = false;
    struct ProfEvent { cudaEvent_t a, b; int cls; };
    std::vector<ProfEvent> prof_events;
    bool capture = false;
    uint32_t cap_bounces = 0;
    std::vector<std::vector<uint32_t>> cap_extend_pixels, cap_shadow_pixels, cap_shadow_lights;
    std::vector<std::vector<bpt_hit>> cap_extend_hits;

    bptd::DScene scene_view() const;
};

#define BPT_CUDA_TRY(ctx, expr)                                                                      \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                        \
            return e__ == cudaErrorMemoryAllocation ? BPT_ERR_OOM : BPT_ERR_CUDA;                    \
        }                                                                                            \
    } while (0)

bpt_status dev_alloc(bpt_context* ctx, DevBuf& b, size_t bytes);
void dev_free(DevBuf& b);
// keeps the allocation when it already holds `bytes` without being more than twice too large (rebuilds of the same scene)
bpt_status dev_reserve(bpt_context* ctx, DevBuf& b, size_t bytes);
bpt_status dev_upload(bpt_context* ctx, DevBuf& b, const void* src, size_t bytes);

// bvh_build.cu
// Builds an LBVH over `n` primitives whose AABBs are in d_lo/d_hi (float4 each, device).
bpt_status lbvh_build(bpt_context* ctx, DevBvh& out, uint32_t n, const float4* d_lo, const float4* d_hi, float* d_bounds6);
bpt_status build_blas_two_level(bpt_context* ctx, uint32_t blas_index);
bpt_status build_all_blas_two_level(bpt_context* ctx);
bpt_status build_blas_merged(bpt_context* ctx);
bpt_status build_tlas(bpt_context* ctx);
bpt_status upload_instance_table(bpt_context* ctx);

// render.cu
bpt_status wavefront_alloc(bpt_context* ctx);
bpt_status wavefront_render(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_first, uint32_t nsamples, const bpt_settings& st, bool keep_ahead = false);
bpt_status wavefront_accumulate_ahead(bpt_context* ctx, uint32_t count);
bpt_status wavefront_accumulate_ahead_rgba16f(bpt_context* ctx, uint32_t total_samples, void* d_out);
bpt_status wavefront_render_primary(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const bpt_settings& st, float* h_depth, bpt_gbuffer_texel* h_gbuffer);
bpt_status wavefront_trace_ao(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const bpt_ao_settings& ao, const float* h_depth,
                              const float* h_normal_roughness, float* h_out);
bpt_status wavefront_trace_reflection(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const bpt_reflection_settings& rs, const float* h_depth,
                                      const bpt_gbuffer_texel* h_gbuffer, float* h_refl, float* h_hit);
bpt_status launch_upscale_half_res(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const float* h_depth, const float* h_nr, const float* h_in, float* h_out);
bpt_status launch_ddgi_lighting(bpt_context* ctx, uint64_t n, const float* h_pos, const float* h_normal, const float* h_view, float* h_out);
bpt_status wavefront_trace_probes(bpt_context* ctx, const bpt_probe_volume& vol, const float* h_table, uint32_t frame_index, uint32_t num_bounces, float* h_out,
                                  uint32_t first_probe = 0, uint32_t num_probes = 0xffffffffu);
bpt_status launch_blend_probes(bpt_context* ctx, const bpt_probe_volume& vol, const float* h_table, uint32_t frame_index, const float* h_rays,
                               const bpt_probe_blend& bl, float* h_irr, float* h_vis);
bpt_status launch_resolve(bpt_context* ctx, uint32_t total_samples, float* d_out);
bpt_status launch_resolve_rgba16f(bpt_context* ctx, uint32_t total_samples, void* d_out);
// lighttex.cu
bpt_status upload_light_textures(bpt_context* ctx, const bpt_light_texture_desc* textures, uint32_t num_textures);
bpt_status read_light_texture(bpt_context* ctx, uint32_t index, float* out, uint64_t capacity_texels, uint64_t* out_texels);
// reblur.cu
bpt_status launch_reblur(bpt_context* ctx, const bpt_camera& cam, uint64_t frame_count, const bpt_reblur_settings& st, const bpt_reblur_inputs& in, float* h_out);
bpt_status reblur_reset(bpt_context* ctx);
bpt_status reblur_debug_read(bpt_context* ctx, uint32_t which, float* out, uint64_t capacity_floats);
// ibl.cu
bpt_status launch_precompute_sky_ibl(bpt_context* ctx, const bpt_sky_ibl_desc& desc);
// post.cu
bpt_status launch_post_process(bpt_context* ctx, const bpt_post_settings& st, uint32_t total_samples, float* d_out);
bpt_status launch_trace_batch(bpt_context* ctx, const bpt_ray* h_rays, uint64_t n, uint32_t frame_index, bpt_hit* h_hits, uint8_t* h_visible);
