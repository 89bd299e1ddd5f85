/* bpt_host.h — C view of libbpt_host.so, the C++20 host mirror of the reference interfaces around the path tracing pass
 * (bisemutum-engine_b200/host/: gfx::Camera, frustum culling, LightsContext packing, PathTracingPass / PostProcessPass,
 * the IRenderer plugin, the headless project loader and the glTF importer). The engine itself links the C++ classes
 * (INTEGRATION.md); this plain-C surface is what non-C++ hosts and the Python test harness bind. The device boundary
 * proper is include/bpt/bpt.h (libbpt.so); nothing here computes on the CPU what the kernels compute.
 * Reference citations are relative to /root/reference (PepcyCh/bisemutum-engine). */
#ifndef BPT_HOST_H
#define BPT_HOST_H
#include "bpt.h"

#ifdef __cplusplus
extern "C" {
#endif
#define BPT_HOST_API __attribute__((visibility("default")))

/* gfx::Camera (include/bisemutum/graphics/camera.hpp:20-60): what CameraSystem writes per frame */
typedef struct bpt_host_camera_desc {
    float position[3]; float front_dir[3]; float up_dir[3];
    float yfov; float near_z; float far_z;
    uint32_t width; uint32_t height; uint32_t orthographic;
} bpt_host_camera_desc;

typedef struct bpt_host_pass bpt_host_pass;         /* a bi::PathTracingPass + its camera + the engine's frame counter */
typedef struct bpt_host_project bpt_host_project;   /* a loaded project directory or an imported glTF model */

typedef struct bpt_host_project_info {
    uint32_t num_drawables, num_blas, num_materials, num_textures, num_dir_lights, num_point_lights, num_rect_lights;
    uint32_t target_width, target_height;
    bpt_host_camera_desc camera;
    float ray_length; uint32_t max_bounces; uint32_t accumulate;   /* BasicRenderer::PathTracingSettings (renderer/basic.hpp:76-81) */
    bpt_ao_settings ambient_occlusion;
} bpt_host_project_info;

/* ---- camera, culling, light packing -------------------------------------------------------------------------------------- */
/* Camera::update_shader_params (src/graphics/camera.cpp:96-159): glm lookAt / reverse-Z perspective / inverse, column-major */
BPT_HOST_API void bpt_host_camera_matrices(const bpt_host_camera_desc* d, bpt_camera* out, float view[16], float proj[16]);
/* Camera::get_frustum_planes (camera.cpp:161-201): six planes (a, b, c, d) */
BPT_HOST_API void bpt_host_frustum_planes(const bpt_host_camera_desc* d, float planes[24]);
/* the CPU frustum culling of render_graph.cpp:391-461: visible[i] = box i (min xyz, max xyz) touches the frustum */
BPT_HOST_API void bpt_host_cull(const float planes[24], const float* aabb_min_max, uint32_t n, uint8_t* visible);
/* Transform::transform_bounding_box (src/math/transform.cpp:62-84) with a row-major 3x4 matrix */
BPT_HOST_API void bpt_host_transform_aabb(const float m[12], const float in[6], float out[6]);
/* LightsContext::collect_all_lights (src/renderer/context/lights.cpp:52-63, 125-143, 208-229); *emitted = 0 for a black light */
BPT_HOST_API void bpt_host_pack_point_light(const float color[3], float strength, float range, int spot, float inner, float outer,
                                            const float translation[3], const float rotation[9], bpt_point_light_data* out, int* emitted);
BPT_HOST_API void bpt_host_pack_rect_light(const float color[3], float strength, float width, float height, int two_sided,
                                           const float translation[3], const float rotation[9], bpt_rect_light_data* out, int* emitted);
BPT_HOST_API void bpt_host_pack_dir_light(const float color[3], float strength, const float rotation[9], bpt_dir_light_data* out, int* emitted);

/* ---- PathTracingPass / PostProcessPass (src/renderer/pass/path_tracing.cpp:224-488, post_process.cpp:92-273) ------------------ */
BPT_HOST_API bpt_host_pass* bpt_host_pass_create(bpt_context* ctx, const bpt_host_camera_desc* cam);
BPT_HOST_API void bpt_host_pass_destroy(bpt_host_pass* p);
BPT_HOST_API void bpt_host_pass_set_camera(bpt_host_pass* p, const bpt_host_camera_desc* cam);
BPT_HOST_API void bpt_host_pass_set_frame(bpt_host_pass* p, uint64_t frame);
BPT_HOST_API void bpt_host_pass_set_prefetch(bpt_host_pass* p, uint32_t frames);   /* samples traced ahead while the history stays valid */
/* device memory behind OutputData.color (rgba16_sfloat, W*H*8 bytes): every frame writes its accumulated colour there (NULL: unbound) */
BPT_HOST_API void bpt_host_pass_set_color_target(bpt_host_pass* p, void* device_rgba16f);
BPT_HOST_API void bpt_host_pass_reset_history(bpt_host_pass* p);                   /* every camera starts a new accumulation at its next frame (drops samples traced ahead) */
/* one engine frame: camera.update_shader_params -> pass.render (records) -> RenderGraph::execute; returns bpt_status */
BPT_HOST_API int bpt_host_pass_frame(bpt_host_pass* p, float ray_length, uint32_t max_bounces, int accumulate, uint64_t* accumulated_frames);
/* PathTracingPass::OutputData depth + G-buffer of the current camera (W*H floats, W*H texels); returns bpt_status */
BPT_HOST_API int bpt_host_pass_read_primary(bpt_host_pass* p, float ray_length, uint32_t max_bounces, float* depth, bpt_gbuffer_texel* gbuffer);
/* PostProcessPass::render on the accumulated colour into out_rgba32f (W*H*4 floats = the back buffer); returns bpt_status */
BPT_HOST_API int bpt_host_pass_post_process(bpt_host_pass* p, int bloom, float threshold, float softness, float* out_rgba32f);

/* ---- IRenderer plugin: register_renderer<CudaPathTracingRenderer>() + set_renderer(name) + `frames` x GraphicsManager::render_frame
 *      (src/graphics/graphics_manager.cpp:407-427). Returns bpt_status, -1 when `renderer_name` is not registered. ------------------ */
BPT_HOST_API int bpt_host_renderer_run(bpt_context* ctx, const char* renderer_name, const bpt_host_camera_desc* cam, uint32_t frames,
                                       const bpt_dir_light_data* dir, uint32_t num_dir, const float* sky_faces, uint32_t sky_size, const float* sky_transform,
                                       const float* sky_color, float ray_length, uint32_t max_bounces, int bloom, float bloom_threshold, float bloom_softness,
                                       float* back_buffer, uint32_t* passes_per_frame);

/* ---- headless ingestion: a project directory (asset manager + ECS deserialisation restated, host/project.hpp) or a glTF 2.0 model
 *      (menu_action_import_model_gltf, src/scene_basic/menu_actions/import_model.cpp:27-430, host/gltf.hpp). NULL + message in `err` on failure. */
BPT_HOST_API bpt_host_project* bpt_host_project_load(const char* dir, char* err, uint64_t err_len);
BPT_HOST_API bpt_host_project* bpt_host_project_import_gltf(const char* path, char* err, uint64_t err_len);
BPT_HOST_API void bpt_host_project_free(bpt_host_project* h);
BPT_HOST_API void bpt_host_project_get_info(const bpt_host_project* h, bpt_host_project_info* o);
/* which: 0 positions, 1 normals, 2 tangents, 3 texcoords, 4 indices, 5 bpt_blas_desc, 6 bpt_drawable_sbt_data, 7 bpt_instance_desc, 8 bpt_material,
 * 9 / 10 / 11 dir / point / rect lights, 16 + k texels of texture k (+ its size and rhi::ResourceFormat). Size in BYTES. */
BPT_HOST_API const void* bpt_host_project_array(const bpt_host_project* h, uint32_t which, uint64_t* bytes, uint32_t* width, uint32_t* height, uint32_t* format);
/* rhi::SamplerDesc enums of texture k: out = {mag_filter, min_filter, address_mode_u, address_mode_v}; -1 if k is out of range */
BPT_HOST_API int bpt_host_project_texture_sampler(const bpt_host_project* h, uint32_t k, uint32_t out[4]);
/* geometry + materials / textures + instances + lights + sky + acceleration structure into ctx; returns bpt_status */
BPT_HOST_API int bpt_host_project_upload(bpt_host_project* h, bpt_context* ctx, uint32_t accel_mode);
BPT_HOST_API const char* bpt_host_project_error(const bpt_host_project* h);
/* StaticMesh::calculate_tspace for one submesh (src/scene_basic/static_mesh.cpp:93-152 = MikkTSpace genTangSpaceDefault); tangents: 4 floats / vertex (out) */
BPT_HOST_API void bpt_host_mikk_tangents(const float* positions, const float* normals, const float* texcoords, float* tangents,
                                         const uint32_t* indices, uint64_t num_indices, uint32_t base_vertex);

#ifdef __cplusplus
}
#endif
#endif
