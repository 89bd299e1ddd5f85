// host/c_exports.cpp — plain-C view of the host mirror for the Python harness (tests / bench).
#include <cstring>
#include <exception>
#include <string>
#include "engine.hpp"
#include "project.hpp"
#include "gltf.hpp"
#include "../../include/bpt/bpt_host.h"     // the declarations: the compiler checks every definition below against them

using namespace bi;

extern "C" {
#define HOST_API __attribute__((visibility("default")))

static gfx::Camera make_camera(const bpt_host_camera_desc* d) {
    gfx::Camera c;
    c.position = {d->position[0], d->position[1], d->position[2]};
    c.front_dir = {d->front_dir[0], d->front_dir[1], d->front_dir[2]};
    c.up_dir = {d->up_dir[0], d->up_dir[1], d->up_dir[2]};
    c.yfov = d->yfov; c.near_z = d->near_z; c.far_z = d->far_z;
    c.projection_type = d->orthographic ? gfx::ProjectionType::orthographic : gfx::ProjectionType::perspective;
    c.set_target_extent(d->width, d->height);
    return c;
}

HOST_API void bpt_host_camera_matrices(const bpt_host_camera_desc* d, bpt_camera* out, float view[16], float proj[16]) {
    gfx::Camera c = make_camera(d);
    c.update_shader_params(0);
    std::memcpy(out->matrix_inv_view, c.matrix_inv_view().data(), 64);
    std::memcpy(out->matrix_inv_proj, c.matrix_inv_proj().data(), 64);
    std::memcpy(out->matrix_proj_view, c.matrix_proj_view().data(), 64);
    if (view) std::memcpy(view, c.matrix_view().data(), 64);
    if (proj) std::memcpy(proj, c.matrix_proj().data(), 64);
}

HOST_API void bpt_host_frustum_planes(const bpt_host_camera_desc* d, float planes[24]) {
    gfx::Camera c = make_camera(d);
    auto p = c.get_frustum_planes();
    for (int i = 0; i < 6; i++) { planes[4 * i] = p[i].x; planes[4 * i + 1] = p[i].y; planes[4 * i + 2] = p[i].z; planes[4 * i + 3] = p[i].w; }
}

// The culling half of RenderGraph::add_rendered_object_list (render_graph.cpp:391-461): visibility of
// each drawable's world AABB against the six planes. (The PT pass itself does not cull — SURVEY §0.)
HOST_API void bpt_host_cull(const float planes[24], const float* aabb_min_max, uint32_t n, uint8_t* visible) {
    float4 pl[6];
    for (int i = 0; i < 6; i++) pl[i] = float4(planes[4 * i], planes[4 * i + 1], planes[4 * i + 2], planes[4 * i + 3]);
    for (uint32_t i = 0; i < n; i++) {
        BoundingBox b;
        b.p_min = {aabb_min_max[6 * i], aabb_min_max[6 * i + 1], aabb_min_max[6 * i + 2]};
        b.p_max = {aabb_min_max[6 * i + 3], aabb_min_max[6 * i + 4], aabb_min_max[6 * i + 5]};
        visible[i] = b.test_with_planes(pl, 6) ? 1 : 0;
    }
}

HOST_API void bpt_host_transform_aabb(const float m[12], const float in[6], float out[6]) {
    BoundingBox b; b.p_min = {in[0], in[1], in[2]}; b.p_max = {in[3], in[4], in[5]};
    BoundingBox r = transform_bounding_box(m, b);
    out[0] = r.p_min.x; out[1] = r.p_min.y; out[2] = r.p_min.z; out[3] = r.p_max.x; out[4] = r.p_max.y; out[5] = r.p_max.z;
}

// LightsContext packing for the harness.
HOST_API void bpt_host_pack_point_light(const float color[3], float strength, float range, int spot, float inner, float outer,
                                        const float translation[3], const float rotation[9], bpt_point_light_data* out, int* emitted) {
    LightsContext lc; PointLightComponent l; LightTransform t;
    l.color = {color[0], color[1], color[2]}; l.strength = strength; l.range = range; l.spot = spot != 0; l.spot_inner_angle = inner; l.spot_outer_angle = outer;
    t.translation = {translation[0], translation[1], translation[2]}; std::memcpy(t.rotation, rotation, 36);
    lc.add(l, t);
    *emitted = (int)lc.point_lights.size();
    if (*emitted) *out = lc.point_lights[0];
}
HOST_API void bpt_host_pack_rect_light(const float color[3], float strength, float width, float height, int two_sided,
                                       const float translation[3], const float rotation[9], bpt_rect_light_data* out, int* emitted) {
    LightsContext lc; RectLightComponent l; LightTransform t;
    l.color = {color[0], color[1], color[2]}; l.strength = strength; l.width = width; l.height = height; l.two_sided = two_sided != 0;
    t.translation = {translation[0], translation[1], translation[2]}; std::memcpy(t.rotation, rotation, 36);
    lc.add(l, t);
    *emitted = (int)lc.rect_lights.size();
    if (*emitted) *out = lc.rect_lights[0];
}
HOST_API void bpt_host_pack_dir_light(const float color[3], float strength, const float rotation[9], bpt_dir_light_data* out, int* emitted) {
    LightsContext lc; DirectionalLightComponent l; LightTransform t;
    l.color = {color[0], color[1], color[2]}; l.strength = strength; std::memcpy(t.rotation, rotation, 36);
    lc.add(l, t);
    *emitted = (int)lc.dir_lights.size();
    if (*emitted) *out = lc.dir_lights[0];
}

// Drives N frames through PathTracingPass::{update_params, render} + RenderGraph::execute exactly as
// BasicRenderer::render_camera would (basic.cpp:46-48,157-166): used by tests and bench e2e.
struct bpt_host_pass { PathTracingPass* pass; gfx::Camera camera; uint64_t frame; };

HOST_API bpt_host_pass* bpt_host_pass_create(bpt_context* ctx, const bpt_host_camera_desc* cam) {
    auto* p = new bpt_host_pass{new PathTracingPass(ctx), make_camera(cam), 0};
    return p;
}
HOST_API void bpt_host_pass_destroy(bpt_host_pass* p) { if (p) { delete p->pass; delete p; } }
HOST_API void bpt_host_pass_set_camera(bpt_host_pass* p, const bpt_host_camera_desc* cam) {
    gfx::Camera c = make_camera(cam);
    p->camera.position = c.position; p->camera.front_dir = c.front_dir; p->camera.up_dir = c.up_dir;
    p->camera.yfov = c.yfov; p->camera.near_z = c.near_z; p->camera.far_z = c.far_z; p->camera.projection_type = c.projection_type;
    p->camera.set_target_extent(cam->width, cam->height);
}
HOST_API void bpt_host_pass_set_frame(bpt_host_pass* p, uint64_t frame) { p->frame = frame; }
// OutputData.depth / .gbuffer of the camera's current frame (PathTracingPass::read_primary_outputs). Returns bpt_status.
HOST_API int bpt_host_pass_read_primary(bpt_host_pass* p, float ray_length, uint32_t max_bounces, float* depth, bpt_gbuffer_texel* gbuffer) {
    BasicRenderer::PathTracingSettings s; s.ray_length = ray_length; s.max_bounces = max_bounces;
    p->camera.update_shader_params(p->frame);
    return (int)p->pass->read_primary_outputs(p->camera, s, depth, gbuffer);
}
// ---- headless project loading (host/project.hpp): the reference's project directory -> the C ABI's arrays ------------------------
struct bpt_host_project { project::Project p; std::string err; };
HOST_API bpt_host_project* bpt_host_project_load(const char* dir, char* err, uint64_t err_len) {
    auto* h = new bpt_host_project();
    bool ok = false;
    try { ok = project::load_project(dir, h->p, h->err); } catch (std::exception const& e) { h->err = std::string("exception: ") + e.what(); }   // nothing may unwind through the C boundary
    if (!ok) {
        if (err && err_len) { std::strncpy(err, h->err.c_str(), err_len - 1); err[err_len - 1] = 0; }
        delete h;
        return nullptr;
    }
    return h;
}
HOST_API void bpt_host_project_free(bpt_host_project* h) { delete h; }
// glTF 2.0 import (host/gltf.hpp = menu_action_import_model_gltf, import_model.cpp:27-430): the model only; camera / lights come from the caller.
HOST_API bpt_host_project* bpt_host_project_import_gltf(const char* path, char* err, uint64_t err_len) {
    auto* h = new bpt_host_project();
    bool ok = false;
    try { ok = project::import_gltf(path, h->p, h->err); } catch (std::exception const& e) { h->err = std::string("exception: ") + e.what(); }   // nothing may unwind through the C boundary
    if (!ok) {
        if (err && err_len) { std::strncpy(err, h->err.c_str(), err_len - 1); err[err_len - 1] = 0; }
        delete h;
        return nullptr;
    }
    return h;
}
// StaticMesh::calculate_tspace for one submesh (static_mesh.cpp:93-152); arrays are indexed from vertex 0, tangents: 4 floats per vertex (out).
HOST_API void bpt_host_mikk_tangents(const float* positions, const float* normals, const float* texcoords, float* tangents,
                                     const uint32_t* indices, uint64_t num_indices, uint32_t base_vertex) {
    project::mikk_tangents(positions, normals, texcoords, tangents, indices, (size_t)num_indices, base_vertex);
}
HOST_API void bpt_host_project_get_info(const bpt_host_project* h, bpt_host_project_info* o) {
    const project::Project& p = h->p;
    *o = bpt_host_project_info{};
    o->num_drawables = (uint32_t)p.drawables.size(); o->num_blas = (uint32_t)p.blas.size(); o->num_materials = (uint32_t)p.materials.size();
    o->num_textures = (uint32_t)p.textures.size(); o->num_dir_lights = (uint32_t)p.lights.dir_lights.size();
    o->num_point_lights = (uint32_t)p.lights.point_lights.size(); o->num_rect_lights = (uint32_t)p.lights.rect_lights.size();
    o->target_width = p.target_width; o->target_height = p.target_height;
    for (int k = 0; k < 3; k++) { o->camera.position[k] = p.cam_position[k]; o->camera.front_dir[k] = p.cam_front[k]; o->camera.up_dir[k] = p.cam_up[k]; }
    o->camera.yfov = p.yfov; o->camera.near_z = p.near_z; o->camera.far_z = p.far_z; o->camera.width = p.target_width; o->camera.height = p.target_height;
    o->camera.orthographic = p.orthographic ? 1u : 0u;
    o->ray_length = p.path_tracing.ray_length; o->max_bounces = p.path_tracing.max_bounces; o->accumulate = p.path_tracing.accumulate ? 1u : 0u;
    o->ambient_occlusion = p.ambient_occlusion;
}
// which: 0 positions, 1 normals, 2 tangents, 3 texcoords, 4 indices, 5 blas descs, 6 drawables, 7 instances, 8 materials, 9 dir lights,
// 10 point lights, 11 rect lights, 16 + k: texels of texture k. Returns the array and its size in BYTES (for tests and for hosts that upload themselves).
HOST_API const void* bpt_host_project_array(const bpt_host_project* h, uint32_t which, uint64_t* bytes, uint32_t* width, uint32_t* height, uint32_t* format) {
    const project::Project& p = h->p;
    auto ret = [&](const void* d, size_t n) { *bytes = n; return d; };
    switch (which) {
        case 0: return ret(p.positions.data(), p.positions.size() * 4);
        case 1: return ret(p.normals.data(), p.normals.size() * 4);
        case 2: return ret(p.tangents.data(), p.tangents.size() * 4);
        case 3: return ret(p.texcoords.data(), p.texcoords.size() * 4);
        case 4: return ret(p.indices.data(), p.indices.size() * 4);
        case 5: return ret(p.blas.data(), p.blas.size() * sizeof(bpt_blas_desc));
        case 6: return ret(p.drawables.data(), p.drawables.size() * sizeof(bpt_drawable_sbt_data));
        case 7: return ret(p.instances.data(), p.instances.size() * sizeof(bpt_instance_desc));
        case 8: return ret(p.materials.data(), p.materials.size() * sizeof(bpt_material));
        case 9: return ret(p.lights.dir_lights.data(), p.lights.dir_lights.size() * sizeof(bpt_dir_light_data));
        case 10: return ret(p.lights.point_lights.data(), p.lights.point_lights.size() * sizeof(bpt_point_light_data));
        case 11: return ret(p.lights.rect_lights.data(), p.lights.rect_lights.size() * sizeof(bpt_rect_light_data));
        default: break;
    }
    if (which >= 16 && which - 16 < p.textures.size()) {
        auto& t = p.textures[which - 16];
        if (width) *width = t.width;
        if (height) *height = t.height;
        if (format) *format = t.format;
        return ret(t.texels.data(), t.texels.size());
    }
    *bytes = 0;
    return nullptr;
}
// sampler of texture k as rhi::SamplerDesc enum values (rhi/sampler.hpp:10-26): out = {mag_filter, min_filter, address_mode_u, address_mode_v}
HOST_API int bpt_host_project_texture_sampler(const bpt_host_project* h, uint32_t k, uint32_t out[4]) {
    if (k >= h->p.textures.size()) return -1;
    auto& t = h->p.textures[k];
    out[0] = t.mag_filter; out[1] = t.min_filter; out[2] = t.address_u; out[3] = t.address_v;
    return 0;
}
// geometry + materials/textures + instances + lights + sky + acceleration structure into `ctx`. Returns bpt_status.
HOST_API int bpt_host_project_upload(bpt_host_project* h, bpt_context* ctx, uint32_t accel_mode) {
    return (int)project::upload_project(h->p, ctx, accel_mode, h->err);
}
HOST_API const char* bpt_host_project_error(const bpt_host_project* h) { return h->err.c_str(); }

// Samples traced ahead per wave while the history stays valid (throughput vs latency of the first frame; default 8).
HOST_API void bpt_host_pass_set_prefetch(bpt_host_pass* p, uint32_t frames) { p->pass->set_prefetch_frames(frames); }
HOST_API void bpt_host_pass_set_color_target(bpt_host_pass* p, void* device_rgba16f) { p->pass->set_color_target(device_rgba16f); }
HOST_API void bpt_host_pass_reset_history(bpt_host_pass* p) { p->pass->reset_history(); }
// One engine frame: camera.update_shader_params → pass.render (records) → rg.execute. Returns bpt_status.
HOST_API int bpt_host_pass_frame(bpt_host_pass* p, float ray_length, uint32_t max_bounces, int accumulate, uint64_t* accumulated_frames) {
    BasicRenderer::PathTracingSettings s; s.ray_length = ray_length; s.max_bounces = max_bounces; s.accumulate = accumulate != 0;
    p->camera.update_shader_params(p->frame);
    p->pass->set_frame_count(p->frame);
    gfx::RenderGraph rg;
    p->pass->render(p->camera, rg, {}, s);
    rg.execute();
    if (accumulated_frames) *accumulated_frames = p->pass->accumulated_frames(p->camera);
    p->frame++;
    return (int)p->pass->last_status();
}

// The step after the pass, as BasicRenderer::render_camera issues it (basic.cpp:228-231): PostProcessPass::render on the
// camera's accumulated colour, written to `out_rgba32f` (W*H*4 floats, stands in for the back buffer). Returns bpt_status.
HOST_API int bpt_host_pass_post_process(bpt_host_pass* p, int bloom, float threshold, float softness, float* out_rgba32f) {
    PostProcessPass post(p->pass->context());
    post.default_volume_.bloom = bloom != 0; post.default_volume_.bloom_threshold = threshold; post.default_volume_.bloom_threshold_softness = softness;
    post.set_output(out_rgba32f, p->pass->accumulated_frames(p->camera));
    gfx::RenderGraph rg;
    post.render(p->camera, rg, {});
    rg.execute();
    return (int)post.last_status();
}

// Renderer level: register_renderer<CudaPathTracingRenderer>() + set_renderer(name) + `frames` x (prepare_renderer_per_frame_data,
// render_camera, RenderGraph::execute), as GraphicsManager::render_frame drives a renderer (graphics_manager.cpp:407-427). Lights are the
// packed arrays a LightsContext would hold. Writes the back buffer of the last frame; *passes_per_frame = render-graph passes recorded.
// Returns bpt_status; -1 when `renderer_name` is not registered.
HOST_API int bpt_host_renderer_run(bpt_context* ctx, const char* renderer_name, const bpt_host_camera_desc* cam, uint32_t frames,
                                   const bpt_dir_light_data* dir, uint32_t num_dir, const float* sky_faces, uint32_t sky_size, const float* sky_transform,
                                   const float* sky_color, float ray_length, uint32_t max_bounces, int bloom, float bloom_threshold, float bloom_softness,
                                   float* back_buffer, uint32_t* passes_per_frame) {
    gfx::GraphicsManager mgr(ctx);
    mgr.register_renderer<CudaPathTracingRenderer>();
    if (!mgr.set_renderer(renderer_name)) return -1;
    auto* r = static_cast<CudaPathTracingRenderer*>(mgr.renderer());
    r->lights_ctx.dir_lights.assign(dir, dir + num_dir);
    r->skybox_ctx.faces_rgba32f = sky_faces; r->skybox_ctx.face_size = sky_size;
    if (sky_transform) std::memcpy(r->skybox_ctx.skybox_transform, sky_transform, 36);
    if (sky_color) r->skybox_ctx.color = float3{sky_color[0], sky_color[1], sky_color[2]};
    r->settings.path_tracing.ray_length = ray_length; r->settings.path_tracing.max_bounces = max_bounces;
    r->post_process.bloom = bloom != 0; r->post_process.bloom_threshold = bloom_threshold; r->post_process.bloom_threshold_softness = bloom_softness;
    r->back_buffer = back_buffer;
    gfx::Camera camera = make_camera(cam);
    for (uint32_t f = 0; f < frames; f++) {
        camera.update_shader_params(f);
        r->path_tracing_pass.set_frame_count(f);
        r->prepare_renderer_per_frame_data();
        r->prepare_renderer_per_camera_data(camera);
        gfx::RenderGraph rg;
        r->render_camera(camera, rg);
        rg.execute();
        if (passes_per_frame) *passes_per_frame = (uint32_t)rg.executed_pass_names().size();
        if (r->last_status() != BPT_OK) break;
    }
    return (int)r->last_status();
}

} // extern "C"
