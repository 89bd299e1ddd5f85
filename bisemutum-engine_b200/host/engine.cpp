// host/engine.cpp — see engine.hpp for the reference citations of every function.
#include "engine.hpp"

#include <algorithm>
#include <cstring>

namespace bi {
namespace gfx {

auto Camera::update_shader_params(uint64_t frame_count) -> void {
    frame_index_ = static_cast<uint32_t>(frame_count);                                       // camera.cpp:80
    auto aspect = 1.0f;
    if (width_ != 0 && height_ != 0) aspect = static_cast<float>(width_) / height_;          // camera.cpp:83-87
    matrix_view_ = math::lookAt(position, position + front_dir, up_dir);                     // camera.cpp:96
    if (projection_type == ProjectionType::perspective) {
        matrix_proj_ = math::perspective_reverse_z(math::radians(yfov), aspect, near_z, far_z);
    } else {
        auto ortho_height = std::tan(math::radians(yfov * 0.5f));
        auto ortho_width = ortho_height * aspect;
        matrix_proj_ = math::ortho_reverse_z(-ortho_width, ortho_width, -ortho_height, ortho_height, near_z, far_z);
    }
    matrix_inv_view_ = math::inverse(matrix_view_);
    matrix_inv_proj_ = math::inverse(matrix_proj_);
    matrix_proj_view_ = matrix_proj_ * matrix_view_;
}

auto Camera::get_frustum_planes() const -> std::array<float4, 6> {
    std::array<float4, 6> planes;
    auto front = math::normalize(front_dir);
    auto right = math::normalize(math::cross(front, up_dir));
    auto up = math::cross(right, front);
    auto aspect = 1.0f;
    if (width_ != 0 && height_ != 0) aspect = static_cast<float>(width_) / height_;
    auto xfov = yfov * aspect;                                                               // camera.cpp:138 (kept as is)
    auto pos_dot_front = math::dot(position, front);
    planes[0] = float4(front, -near_z - pos_dot_front);
    planes[1] = float4(-front, pos_dot_front + far_z);
    auto with_w = [&](float3 n) { return float4(n, -math::dot(position, n)); };
    if (projection_type == ProjectionType::perspective) {
        auto vert_angle = math::radians(90.0f - yfov * 0.5f);
        planes[2] = with_w(math::rotate_direction(-vert_angle, right, front));
        planes[3] = with_w(math::rotate_direction(vert_angle, right, front));
        auto hori_angle = math::radians(90.0f - xfov * 0.5f);
        planes[4] = with_w(math::rotate_direction(-hori_angle, up, front));
        planes[5] = with_w(math::rotate_direction(hori_angle, up, front));
    } else {
        auto ortho_height = std::tan(math::radians(yfov * 0.5f));
        auto ortho_width = ortho_height * aspect;
        auto pos_dot_up = math::dot(position, up);
        planes[2] = float4(-up, ortho_height + pos_dot_up);
        planes[3] = float4(up, ortho_height - pos_dot_up);
        planes[4] = float4(right, ortho_width - pos_dot_up);                                  // camera.cpp:169-170 (kept as is)
        planes[5] = float4(-right, ortho_width + pos_dot_up);
    }
    return planes;
}

auto RenderGraph::add_texture(uint32_t, uint32_t, uint32_t) -> TextureHandle { return TextureHandle{next_texture_++}; }

auto RenderGraph::execute() -> void {
    ComputePassContext ctx{this};
    for (auto& p : passes_) {
        if (p->builder.execute_) p->builder.execute_(&p->data, ctx);
        executed_.push_back(p->name);
    }
    passes_.clear();
}

} // namespace gfx

auto LightsContext::add(DirectionalLightComponent const& light, LightTransform const& transform) -> void {
    bpt_dir_light_data data{};
    float3 emission = light.color * light.strength;
    if (emission == float3(0.0f)) return;
    data.emission[0] = emission.x; data.emission[1] = emission.y; data.emission[2] = emission.z;
    float3 d = transform.transform_direction_without_scaling({0.0f, 1.0f, 0.0f});
    data.direction[0] = d.x; data.direction[1] = d.y; data.direction[2] = d.z;
    data.sm_index = -1;     // shadow maps are not produced: visibility is a shadow ray
    dir_lights.push_back(data);
}
auto LightsContext::add(PointLightComponent const& light, LightTransform const& transform) -> void {
    bpt_point_light_data data{};
    float3 emission = light.color * light.strength;
    if (emission == float3(0.0f)) return;
    data.emission[0] = emission.x; data.emission[1] = emission.y; data.emission[2] = emission.z;
    data.position[0] = transform.translation.x; data.position[1] = transform.translation.y; data.position[2] = transform.translation.z;
    float3 d = transform.transform_direction_without_scaling({0.0f, 1.0f, 0.0f});
    data.direction[0] = d.x; data.direction[1] = d.y; data.direction[2] = d.z;
    if (light.spot) {
        data.cos_outer = std::cos(math::radians(light.spot_outer_angle));
        data.cos_inner = std::cos(math::radians(light.spot_inner_angle));
    } else { data.cos_outer = 0.0f; data.cos_inner = 0.0f; }
    data.range_sqr_inv = 1.0f / (light.range * light.range);
    data.sm_index = -1;
    point_lights.push_back(data);
}
auto LightsContext::add(RectLightComponent const& light, LightTransform const& transform) -> void {
    bpt_rect_light_data data{};
    float3 emission = light.color * light.strength;
    if (emission == float3(0.0f)) return;
    data.emission[0] = emission.x; data.emission[1] = emission.y; data.emission[2] = emission.z;
    float3 c = transform.translation;
    auto w = 0.5f * light.width, h = 0.5f * light.height;
    auto put = [](float* dst, float3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    put(data.center_position, c);
    put(data.position0, c + transform.transform_direction_without_scaling({w, h, 0.0f}));
    put(data.position1, c + transform.transform_direction_without_scaling({-w, h, 0.0f}));
    put(data.position2, c + transform.transform_direction_without_scaling({-w, -h, 0.0f}));
    put(data.position3, c + transform.transform_direction_without_scaling({w, -h, 0.0f}));
    put(data.normal, transform.transform_direction_without_scaling({0.0f, 0.0f, 1.0f}));
    data.inv_width_sqr = 1.0f / (light.width * light.width);
    data.inv_height_sqr = 1.0f / (light.height * light.height);
    data.two_sided = light.two_sided;
    data.texture_index = -1;
    rect_lights.push_back(data);
}

auto PathTracingPass::update_params(LightsContext& lights_ctx, SkyboxContext& skybox_ctx, BasicRenderer::PathTracingSettings const&) -> void {
    // The reference re-binds lights/sky into per-bounce parameter blocks here (path_tracing.cpp:183-222);
    // the CUDA context keeps one copy, re-uploaded once per frame like the reference's uniform update.
    status_ = bpt_scene_upload_lights(ctx_, lights_ctx.dir_lights.data(), (uint32_t)lights_ctx.dir_lights.size(),
                                      lights_ctx.point_lights.data(), (uint32_t)lights_ctx.point_lights.size(),
                                      lights_ctx.rect_lights.data(), (uint32_t)lights_ctx.rect_lights.size(), &lights_ctx.ltc_luts);
    if (status_ != BPT_OK) return;
    float col[3] = {skybox_ctx.color.x, skybox_ctx.color.y, skybox_ctx.color.z};
    // The skybox texture is handed over only when it changes (the reference's `last_precomputed_skybox_ != current_skybox.tex`,
    // skybox_precompute.cpp:97); transform and colour are per-frame uniforms (skybox.cpp:40-50).
    if (!sky_uploaded_ || skybox_ctx.faces_rgba32f != last_sky_faces_ || skybox_ctx.face_size != last_sky_size_) {
        status_ = bpt_scene_upload_sky(ctx_, skybox_ctx.faces_rgba32f, skybox_ctx.face_size, skybox_ctx.skybox_transform, col);
        sky_uploaded_ = status_ == BPT_OK; last_sky_faces_ = skybox_ctx.faces_rgba32f; last_sky_size_ = skybox_ctx.face_size;
    } else {
        status_ = bpt_scene_update_sky_params(ctx_, skybox_ctx.skybox_transform, col);
    }
}

auto PathTracingPass::render(gfx::Camera const& camera, gfx::RenderGraph& rg, InputData const&,
                             BasicRenderer::PathTracingSettings const& settings) -> OutputData {
    auto width = camera.target_width();
    auto height = camera.target_height();
    auto frame_count = frame_counter_;
    // history validity: path_tracing.cpp:231-246
    auto [hist_camera_it, is_new_camera] = camera_history_infos_.try_emplace(&camera);
    auto has_history = !is_new_camera
        && hist_camera_it->second.last_frame + 1 == frame_count
        && hist_camera_it->second.width == width
        && hist_camera_it->second.height == height
        && hist_camera_it->second.proj_view == camera.matrix_proj_view();
    hist_camera_it->second.last_frame = frame_count;
    hist_camera_it->second.width = width;
    hist_camera_it->second.height = height;
    hist_camera_it->second.proj_view = camera.matrix_proj_view();
    if (has_history && settings.accumulate) ++hist_camera_it->second.frame_count;
    else hist_camera_it->second.frame_count = 1;
    bool reset = hist_camera_it->second.frame_count == 1;

    OutputData out;
    out.color = rg.add_texture(width, height, 16);
    out.depth = rg.add_texture(width, height, 4);                                  // d32_sfloat (path_tracing.cpp:258-262)
    out.velocity = gfx::TextureHandle{};                                           // invalid, as the reference (:485)
    out.gbuffer.base_color = rg.add_texture(width, height, 8);                     // gbuffer.hpp:14-17
    out.gbuffer.normal_roughness = rg.add_texture(width, height, 8);
    out.gbuffer.fresnel = rg.add_texture(width, height, 8);
    out.gbuffer.material_0 = rg.add_texture(width, height, 4);

    struct PassData { gfx::TextureHandle color; };
    auto [builder, pass_data] = rg.add_compute_pass<PassData>("PT CUDA wavefront");
    pass_data->color = builder.write(out.color);
    builder.set_execution_function<PassData>(
        [this, &camera, settings, reset, frames = hist_camera_it->second.frame_count](CRef<PassData>, gfx::ComputePassContext const&) {
            bpt_camera cam;
            std::memcpy(cam.matrix_inv_view, camera.matrix_inv_view().data(), 64);
            std::memcpy(cam.matrix_inv_proj, camera.matrix_inv_proj().data(), 64);
            std::memcpy(cam.matrix_proj_view, camera.matrix_proj_view().data(), 64);
            bpt_settings st{};
            st.ray_length = settings.ray_length;
            st.max_bounces = settings.max_bounces;
            st.accumulate = settings.accumulate;
            if (reset) status_ = bpt_clear_accum(ctx_);                       // also drops prefetched samples
            if (status_ != BPT_OK) return;
            // Sample prefetch: while the history is valid the next frames will ask for frame_index+1, +2, ...
            // with the same camera, so a whole wave of samples is traced at once and handed out one per frame.
            // The accumulated image is bit-identical to tracing one sample per frame.
            uint32_t pending = 0, next = 0;
            bpt_pending_ahead(ctx_, &pending, &next);
            bool same_settings = std::memcmp(&st, &ahead_settings_, sizeof(st)) == 0;
            if (pending == 0 || next != camera.frame_index() || !same_settings) {
                uint32_t depth = settings.accumulate ? prefetch_frames_ : 1u;
                status_ = bpt_render_ahead(ctx_, &cam, camera.frame_index(), depth, &st, nullptr);
                ahead_settings_ = st;
            }
            if (status_ != BPT_OK) return;
            // "PT Accumulate" (path_tracing.cpp:461-480): the frame's sample joins the history; with a colour target bound the
            // accumulated colour is written in the same launch
            status_ = color_target_ ? bpt_accumulate_ahead_rgba16f(ctx_, (uint32_t)frames, color_target_) : bpt_accumulate_ahead(ctx_, 1);
        });
    return out;
}

auto PathTracingPass::read_primary_outputs(gfx::Camera const& camera, BasicRenderer::PathTracingSettings const& settings, float* depth,
                                           bpt_gbuffer_texel* gbuffer) -> bpt_status {
    bpt_camera cam;
    std::memcpy(cam.matrix_inv_view, camera.matrix_inv_view().data(), 64);
    std::memcpy(cam.matrix_inv_proj, camera.matrix_inv_proj().data(), 64);
    std::memcpy(cam.matrix_proj_view, camera.matrix_proj_view().data(), 64);
    bpt_settings st{};
    st.ray_length = settings.ray_length;
    st.max_bounces = settings.max_bounces;
    status_ = bpt_render_primary(ctx_, &cam, camera.frame_index(), &st, depth, gbuffer);
    return status_;
}

auto PathTracingPass::accumulated_frames(gfx::Camera const& camera) const -> uint64_t {
    auto it = camera_history_infos_.find(&camera);
    return it == camera_history_infos_.end() ? 0 : it->second.frame_count;
}

// post_process.cpp:92-273: with bloom the reference records "Bloom Pre Pass", 3 x ("Bloom Horizontal Pass #i", "Bloom Vertical
// Pass #i"), "Bloom Combine Pass #2", "#1", "Bloom Final Combine Pass" and "Post Process Pass"; here they are one pass.
auto PostProcessPass::render(gfx::Camera const& camera, gfx::RenderGraph& rg, InputData const& input) -> void {
    auto const& volume = default_volume_;
    struct PassData { gfx::TextureHandle input_color; gfx::TextureHandle input_depth; gfx::TextureHandle output; };
    auto [builder, pass_data] = rg.add_compute_pass<PassData>("Post Process CUDA");
    pass_data->input_color = builder.read(input.color);
    pass_data->input_depth = builder.read(input.depth);
    pass_data->output = builder.write(rg.add_texture(camera.target_width(), camera.target_height(), 16));
    builder.set_execution_function<PassData>(
        [this, volume](CRef<PassData>, gfx::ComputePassContext const&) {
            bpt_post_settings st{};
            st.bloom = volume.bloom ? 1u : 0u;
            st.bloom_threshold = volume.bloom_threshold;
            st.bloom_threshold_softness = volume.bloom_threshold_softness;
            status_ = out_ ? bpt_post_process(ctx_, &st, (uint32_t)frames_, out_) : BPT_ERR_INVALID;
        });
}

// reblur.cpp:273-588. The history rule (consecutive frame counts per camera, :282-285) and the history textures the reference parks
// on the camera live inside the library's context; blur_radius / anti_flickering_strength / virtual_history are the constants of :318-319,383.
auto ReblurPass::render(gfx::Camera const& camera, gfx::RenderGraph& rg, InputData const& input) -> gfx::TextureHandle {
    auto denoised = rg.add_texture(textures_.width, textures_.height, 8);                  // rgba16_sfloat (:287-291)
    struct PassData { gfx::TextureHandle noised; gfx::TextureHandle denoised; };
    auto [builder, pass_data] = rg.add_compute_pass<PassData>("ReBLUR CUDA");
    pass_data->noised = builder.read(input.noised_tex);
    pass_data->denoised = builder.write(denoised);
    builder.set_execution_function<PassData>(
        [this, &camera](CRef<PassData>, gfx::ComputePassContext const&) {
            bpt_camera cam;
            std::memcpy(cam.matrix_inv_view, camera.matrix_inv_view().data(), 64);
            std::memcpy(cam.matrix_inv_proj, camera.matrix_inv_proj().data(), 64);
            std::memcpy(cam.matrix_proj_view, camera.matrix_proj_view().data(), 64);
            bpt_reblur_settings st{};
            st.virtual_history = 1; st.blur_radius = 0.9f; st.anti_flickering_strength = 3.5f;
            status_ = out_ ? bpt_denoise_reblur(ctx_, &cam, frame_counter_, &st, &textures_, out_) : BPT_ERR_INVALID;
        });
    return denoised;
}

// ---- renderer level ------------------------------------------------------------------------------------------------
auto CudaPathTracingRenderer::prepare_renderer_per_frame_data() -> void {                 // basic.cpp:31-48
    path_tracing_pass.update_params(lights_ctx, skybox_ctx, settings.path_tracing);
}
auto CudaPathTracingRenderer::prepare_renderer_per_camera_data(gfx::Camera const&) -> void {}   // basic.cpp:73-76: nothing for this pipeline
auto CudaPathTracingRenderer::render_camera(gfx::Camera const& camera, gfx::RenderGraph& rg) -> void {
    auto pt_output = path_tracing_pass.render(camera, rg, {scene_accel}, settings.path_tracing);          // basic.cpp:157-166
    post_process_pass.default_volume_ = post_process;
    post_process_pass.set_output(back_buffer, std::max<uint64_t>(path_tracing_pass.accumulated_frames(camera), 1));
    post_process_pass.render(camera, rg, {pt_output.color, pt_output.depth});                             // basic.cpp:228-231
}
auto CudaPathTracingRenderer::last_status() const -> bpt_status {
    return path_tracing_pass.last_status() != BPT_OK ? path_tracing_pass.last_status() : post_process_pass.last_status();
}

auto gfx::GraphicsManager::set_renderer(std::string_view name) -> bool {                  // graphics_manager.hpp:82-87, engine.cpp:147-148
    auto it = renderer_creators_.find(std::string(name));
    if (it == renderer_creators_.end()) return false;
    renderer_ = it->second(ctx_);
    return true;
}

} // namespace bi
