/*
 * oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU oracle: a multithreaded C++ restatement of the reference's wavefront
 * path-tracing kernels (PepcyCh/bisemutum-engine; see oracle/README.md for the file map) plus
 * the NEW stages the north star adds (LBVH build, software traversal, shadow-ray NEE), which
 * have no reference counterpart and are DEFINED by this oracle.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer values for
 * this path and cannot be built or run here (HLSL/DXC + Vulkan HW ray tracing + window; see
 * DESIGN.md).  The only external pins are the RNG known-answer values of SURVEY.md Appendix D.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libbpt.so) never links or calls it.
 *
 * The entry points deliberately mirror include/bpt/bpt.h one-to-one with the prefix `obpt_`
 * so a parity test drives both sides with the same calls and the same POD inputs.
 */
#ifndef ORACLE_H_
#define ORACLE_H_
#include "../include/bpt/bpt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct obpt_context obpt_context;

BPT_API bpt_status obpt_create(const bpt_config* cfg, obpt_context** out_ctx);
BPT_API bpt_status obpt_destroy(obpt_context* ctx);
BPT_API const char* obpt_last_error(const obpt_context* ctx);
BPT_API bpt_status obpt_set_threads(obpt_context* ctx, uint32_t num_threads); /* 0 = hardware_concurrency */
BPT_API uint32_t obpt_get_threads(const obpt_context* ctx);
BPT_API bpt_status obpt_resize(obpt_context* ctx, uint32_t width, uint32_t height);

BPT_API bpt_status obpt_scene_upload_geometry(
    obpt_context* ctx, const bpt_geometry_streams* streams,
    const bpt_drawable_sbt_data* drawables, const uint32_t* drawable_vertex_attributes, uint32_t num_drawables,
    const bpt_blas_desc* blas, uint32_t num_blas);
BPT_API bpt_status obpt_scene_upload_instances(obpt_context* ctx, const bpt_instance_desc* instances, uint32_t num_instances);
BPT_API bpt_status obpt_scene_upload_materials(
    obpt_context* ctx, const bpt_material* materials, uint32_t num_materials,
    const bpt_texture_desc* textures, uint32_t num_textures);
BPT_API bpt_status obpt_scene_upload_lights(
    obpt_context* ctx,
    const bpt_dir_light_data* dir_lights, uint32_t num_dir_lights,
    const bpt_point_light_data* point_lights, uint32_t num_point_lights,
    const bpt_rect_light_data* rect_lights, uint32_t num_rect_lights,
    const bpt_ltc_luts* ltc_luts);
BPT_API bpt_status obpt_scene_upload_sky(
    obpt_context* ctx, const float* faces_rgba32f, uint32_t face_size,
    const float skybox_transform[9], const float skybox_color[3]);

BPT_API bpt_status obpt_scene_update_sky_params(obpt_context* ctx, const float skybox_transform[9], const float skybox_color[3]);
BPT_API bpt_status obpt_build_accel(obpt_context* ctx, uint32_t mode);
BPT_API bpt_status obpt_update_tlas(obpt_context* ctx);
BPT_API bpt_status obpt_debug_read_bvh(
    obpt_context* ctx, uint32_t which, uint32_t* num_prims,
    uint64_t* sorted_morton, uint32_t* sorted_prims, bpt_bvh_node* nodes, uint32_t capacity, int32_t* root);

BPT_API bpt_status obpt_clear_accum(obpt_context* ctx);
BPT_API bpt_status obpt_render(
    obpt_context* ctx, const bpt_camera* camera, uint32_t frame_index_first, uint32_t num_samples,
    const bpt_settings* settings);
BPT_API bpt_status obpt_resolve(obpt_context* ctx, uint32_t total_samples, float* out_rgba32f);
BPT_API bpt_status obpt_get_counters(obpt_context* ctx, bpt_counters* out);
BPT_API bpt_status obpt_reset_counters(obpt_context* ctx);

BPT_API bpt_status obpt_trace_rays(obpt_context* ctx, const bpt_ray* rays, uint64_t num_rays, uint32_t frame_index, bpt_hit* out_hits);
BPT_API bpt_status obpt_trace_shadow_rays(obpt_context* ctx, const bpt_ray* rays, uint64_t num_rays, uint32_t frame_index, uint8_t* out_visible);
BPT_API bpt_status obpt_debug_capture(obpt_context* ctx, uint32_t enable);
BPT_API bpt_status obpt_debug_read_queue(
    obpt_context* ctx, uint32_t bounce, uint32_t kind,
    uint32_t* pixels, uint32_t* lights, bpt_hit* hits, uint64_t capacity, uint64_t* count);
/* Oracle-only: bounces >= `bounce` of obpt_render / obpt_trace_probes walk the 4-wide quantised tree (as the CUDA kernels do with
 * bounce = 2); 0 = binary tree everywhere (default). Hits are identical by construction; only obpt_stats changes. Merged mode. */
BPT_API bpt_status obpt_set_wide_from_bounce(obpt_context* ctx, uint32_t bounce);
/* The 4-wide quantised tree of merged mode (oracle_wide.cpp = the definition csrc/bpt_wide.cuh must reproduce) and work
 * statistics of wide traversals on a ray batch: counts = {rays, wide nodes, triangles, child boxes, exact leaf boxes}. */
BPT_API bpt_status obpt_denoise_reblur(obpt_context* ctx, const bpt_camera* camera, uint64_t frame_count, const bpt_reblur_settings* settings,
                                       const bpt_reblur_inputs* inputs, float* out_rgba32f);
BPT_API bpt_status obpt_reblur_reset(obpt_context* ctx);
BPT_API bpt_status obpt_debug_read_reblur(obpt_context* ctx, uint32_t which, float* out, uint64_t capacity_floats);
BPT_API bpt_status obpt_scene_upload_light_textures(obpt_context* ctx, const bpt_light_texture_desc* textures, uint32_t num_textures);
BPT_API bpt_status obpt_debug_read_light_texture(obpt_context* ctx, uint32_t index, float* out_rgba32f, uint64_t capacity_texels, uint64_t* out_texels);
BPT_API bpt_status obpt_debug_read_wide(obpt_context* ctx, float* wide_nodes, float* leaf_boxes, uint32_t capacity_leaves);
BPT_API bpt_status obpt_wide_stats(obpt_context* ctx, uint32_t width, uint32_t quantised, uint32_t order, const bpt_ray* rays, uint64_t n,
                                   float* t_out, uint32_t* prim_out, uint64_t counts[5]);
/* calc_ddgi_volume_lighting (ddgi/ddgi_lighting.hlsl:7-83) and the previous-update feedback of the probe lighting pass. */
BPT_API bpt_status obpt_set_ddgi_volume(obpt_context* ctx, const bpt_probe_volume* volume, const bpt_probe_blend* sizes,
                                        const float* irradiance_atlas, const float* visibility_atlas);
BPT_API bpt_status obpt_ddgi_lighting(obpt_context* ctx, uint64_t n, const float* position, const float* normal, const float* view, float* out_rgba);
/* OutputData{depth, gbuffer} of PathTracingPass::render (path_tracing.cpp:482-487) and AmbientOcclusionPass::render_raytraced. */
BPT_API bpt_status obpt_render_primary(obpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const bpt_settings* settings,
                                       float* out_depth, bpt_gbuffer_texel* out_gbuffer);
BPT_API bpt_status obpt_trace_ao(obpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const bpt_ao_settings* settings,
                                 const float* depth, const float* normal_roughness, float* out_ao);
BPT_API bpt_status obpt_precompute_sky_ibl(obpt_context* ctx, const bpt_sky_ibl_desc* desc);
BPT_API bpt_status obpt_debug_read_sky_ibl(obpt_context* ctx, float* diffuse_rgba32f, float* specular_rgba32f, float* brdf_lut_rg32f);
BPT_API bpt_status obpt_trace_reflection(obpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const bpt_reflection_settings* settings,
                                         const float* depth, const bpt_gbuffer_texel* gbuffer, float* out_reflection, float* out_hit_positions);
BPT_API bpt_status obpt_upscale_half_res(obpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const float* depth, const float* normal_roughness,
                                         const float* in_half_res, float* out_full_res);
BPT_API bpt_status obpt_trace_probes(
    obpt_context* ctx, const bpt_probe_volume* volume, const float* sample_table_r2,
    uint32_t frame_index, uint32_t num_bounces, float* out_radiance_dist);

BPT_API bpt_status obpt_trace_probes_range(
    obpt_context* ctx, const bpt_probe_volume* volume, const float* sample_table_r2,
    uint32_t frame_index, uint32_t num_bounces, uint32_t first_probe, uint32_t num_probes, float* out_radiance_dist);

BPT_API bpt_status obpt_blend_probes(
    obpt_context* ctx, const bpt_probe_volume* volume, const float* sample_table_r2, uint32_t frame_index,
    const float* ray_radiance_dist, const bpt_probe_blend* blend, float* irradiance_atlas_rgba32f, float* visibility_atlas_rg32f);

/* PostProcessPass::render (post_process.cpp:92-273): bloom chain + output pass on a W x H rgba32f image (oracle_post.cpp). */
BPT_API bpt_status obpt_post_process_image(const float* in_rgba32f, uint32_t width, uint32_t height, const bpt_post_settings* settings, float* out_rgba32f);

/* Oracle-only: traversal work counters, the source of the "algorithmic bytes" of SURVEY §8d
 * (64 B per node visit, 48 B per triangle test) on a BVH that is bit-identical to the GPU's. */
typedef struct obpt_stats {
    uint64_t extend_rays, extend_nodes, extend_tris, extend_instances;
    uint64_t shadow_rays, shadow_nodes, shadow_tris, shadow_instances;
    uint64_t shaded_vertices;  /* path vertices that ran material + lighting */
    uint64_t miss_vertices;    /* paths that ended on the sky               */
    uint64_t samples;
    /* rays of bounces >= obpt_set_wide_from_bounce walk the 4-wide quantised tree: steps and exact leaf-box tests */
    uint64_t extend_wide_nodes, extend_leaf_boxes, shadow_wide_nodes, shadow_leaf_boxes;
    uint64_t extend_wide_rays, shadow_wide_rays;
} obpt_stats;
BPT_API bpt_status obpt_get_stats(obpt_context* ctx, obpt_stats* out);
/* Oracle-only: render only the 16x16 tiles t with t % stride == offset (a bounded, spatially uniform
 * sample of the frame for the CPU-baseline timing). stride 1 = every tile (default). */
BPT_API bpt_status obpt_set_tile_sample(obpt_context* ctx, uint32_t stride, uint32_t offset);

/* Oracle-only: double-precision accumulation of a many-spp reference image (relMSE gate). */
BPT_API bpt_status obpt_render_converged(
    obpt_context* ctx, const bpt_camera* camera, uint32_t frame_index_first, uint32_t num_samples,
    const bpt_settings* settings, float* out_rgba32f);

/* ---- host-side restatements (camera / culling), reference: src/graphics/camera.cpp,
 *      src/math/{math,bbox,transform}.cpp, src/graphics/render_graph.cpp:391-461 ---------- */
typedef struct obpt_camera_desc {
    float position[3]; float front_dir[3]; float up_dir[3];
    float yfov; float near_z; float far_z; float aspect;
    uint32_t orthographic;
} obpt_camera_desc;
BPT_API void obpt_camera_matrices(const obpt_camera_desc* cam, float view[16], float proj[16], bpt_camera* out);
BPT_API void obpt_frustum_planes(const obpt_camera_desc* cam, float planes[24]);
/* visible[i] = BoundingBox::test_with_planes of the world AABB of drawable i (bbox.cpp:43-56). */
BPT_API void obpt_cull_aabbs(const float planes[24], const float* aabb_min_max /* n*6 */, uint32_t n, uint8_t* visible);
/* Transform::transform_bounding_box for a 3x4 row-major matrix (transform.cpp:59-78). */
BPT_API void obpt_transform_aabb(const float m[12], const float in_min_max[6], float out_min_max[6]);

/* Known-answer helpers for unit tests. */
BPT_API uint32_t obpt_rng_tea(uint32_t v0, uint32_t v1);
BPT_API uint32_t obpt_rng_lcg(uint32_t* state);
BPT_API void obpt_sincos_2pi(float u, float* s, float* c);
BPT_API float obpt_atan2(float y, float x);
BPT_API float obpt_acos(float x);
BPT_API void obpt_ggx_vndf_sample(const float v[3], float rx, float ry, float u1, float u2, float out_h[3]);
BPT_API void obpt_surface_eval_lit(const float N[3], const float T[3], const float V[3], const float L[3],
                                   const float base[3], const float f0[3], const float f90[3],
                                   float roughness, float anisotropy, float out_rgb[3]);
/* state_precision = reference_fp16: one value through an rgba16_sfloat store; (N, T, surface) through the four G-buffer
 * textures (gbuffer.hlsl:18-45). in12 = base[3] f0[3] f90[3] roughness anisotropy ior; out18 = N T base f0 f90 roughness anisotropy ior. */
/* Unit entry points for the independent float64 pins (tests/test_oracle.py): one reference function each. */
BPT_API void obpt_unit_ltc_integrate(const float P[3], const float N[3], const float T[3], const float B[3], const float* Minv9_or_null, const float L[12],
                                     uint32_t two_sided, float* integral, float mrp[3]);
BPT_API void obpt_unit_rect_light(obpt_context* ctx, const bpt_rect_light_data* light, const float P[3], const float N[3], const float T[3], const float B[3],
                                  const float V[3], const float base[3], const float f0[3], const float f90[3], float roughness, float anisotropy,
                                  float out_rgb[3], float diff_mrp_or_null[3]);
BPT_API void obpt_unit_point_light(const bpt_point_light_data* light, const float P[3], float radiance[3], float dir[3], float* dist);
BPT_API void obpt_unit_sample_sky(obpt_context* ctx, const float dir[3], float out_rgb[3]);
BPT_API bpt_status obpt_unit_hit_vertex(obpt_context* ctx, uint32_t instance_slot, uint32_t prim, float u, float v, float out17[17]);
BPT_API float obpt_unit_log2(float x);
BPT_API void obpt_unit_reblur_scalars(float ndotv, float roughness, float parallax, float out6[6]);
BPT_API float obpt_store_half(float f);
BPT_API void obpt_gbuffer_roundtrip(const float N[3], const float T[3], const float in12[12], uint32_t model, float out18[18], uint32_t* model_out);
BPT_API uint64_t obpt_morton63(const float c[3], const float lo[3], const float hi[3]);

#ifdef __cplusplus
}
#endif
#endif
