// host/engine.hpp — host-side mirror of the reference interfaces that sit either side of the
// path-tracing pass, so that `PathTracingPass` below reads like (and drops in for) the reference's
// bisemutum/src/renderer/pass/path_tracing.{hpp,cpp}. Only what this path touches is mirrored:
//   gfx::Camera            include/bisemutum/graphics/camera.hpp:30-75, src/graphics/camera.cpp:73-174
//   gfx::RenderGraph       include/bisemutum/graphics/render_graph.hpp:22-123 (add_compute_pass /
//                          builder.read/write / set_execution_function / execute) — a SHIM: passes run
//                          in submission order, there is no resource aliasing or barrier logic because the
//                          CUDA pass owns its device buffers (the reference's RHI exposes no CUDA interop,
//                          include/bisemutum/graphics/resource.hpp:49-185)
//   light components       include/bisemutum/scene_basic/light.hpp:9-95
//   LightsContext          src/renderer/context/lights.{hpp,cpp} (collect_all_lights packing)
//   SkyboxContext          src/renderer/context/skybox.hpp:9-35 + src/scene_basic/skybox.cpp:43-73
//   BasicRenderer::PathTracingSettings   include/bisemutum/renderer/basic.hpp:76-81
//   IRenderer              include/bisemutum/graphics/renderer.hpp:10-23 (as an abstract class; the
//                          reference type-erases with AnyAny)
#pragma once
#include <algorithm>
#include <any>
#include <functional>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/bpt/bpt.h"
#include "math.hpp"

namespace bi {

template <typename T> using Ref = T*;
template <typename T> using CRef = T const*;

namespace gfx {

enum class ProjectionType : uint8_t { perspective, orthographic };

struct TextureHandle { uint32_t id = ~0u; bool valid() const { return id != ~0u; } };
struct AccelerationStructureHandle { uint32_t id = ~0u; };

struct Camera final {
    auto update_shader_params(uint64_t frame_count) -> void;           // camera.cpp:73-118
    auto matrix_proj() const -> float4x4 const& { return matrix_proj_; }
    auto matrix_view() const -> float4x4 const& { return matrix_view_; }
    auto matrix_proj_view() const -> float4x4 const& { return matrix_proj_view_; }
    auto matrix_inv_view() const -> float4x4 const& { return matrix_inv_view_; }
    auto matrix_inv_proj() const -> float4x4 const& { return matrix_inv_proj_; }
    auto frame_index() const -> uint32_t { return frame_index_; }
    // Order: front, back, top, down, left, right; normals point inwards (camera.cpp:126-174)
    auto get_frustum_planes() const -> std::array<float4, 6>;
    auto set_target_extent(uint32_t width, uint32_t height) -> void { width_ = width; height_ = height; }
    auto target_width() const -> uint32_t { return width_; }
    auto target_height() const -> uint32_t { return height_; }

    float3 position = float3{0.0f, 0.0f, -1.0f};
    float3 front_dir = float3{0.0f, 0.0f, 1.0f};
    float3 up_dir = float3{0.0f, 1.0f, 0.0f};
    float yfov = 30.0f;
    float near_z = 0.01f;
    float far_z = 10000.0f;
    ProjectionType projection_type = ProjectionType::perspective;
    bool enabled = true;

private:
    uint32_t width_ = 0, height_ = 0;
    uint32_t frame_index_ = 0;
    float4x4 matrix_view_{1.0f}, matrix_proj_{1.0f}, matrix_inv_view_{1.0f}, matrix_inv_proj_{1.0f}, matrix_proj_view_{1.0f};
};

struct RenderGraph;
struct ComputePassContext final { RenderGraph* rg = nullptr; };

struct ComputePassBuilder final {
    auto read(TextureHandle h) -> TextureHandle { return h; }
    auto write(TextureHandle h) -> TextureHandle { return h; }
    template <typename PassData>
    auto set_execution_function(std::function<auto(CRef<PassData>, ComputePassContext const&) -> void> func) -> void {
        execute_ = [func = std::move(func)](std::any const* data, ComputePassContext const& ctx) { func(std::any_cast<PassData>(data), ctx); };
    }
    std::function<void(std::any const*, ComputePassContext const&)> execute_;
};

struct RenderGraph final {
    auto add_texture(uint32_t width, uint32_t height, uint32_t bytes_per_texel) -> TextureHandle;
    template <typename PassData>
    auto add_compute_pass(std::string_view name) -> std::pair<ComputePassBuilder&, Ref<PassData>> {
        passes_.push_back(std::make_unique<Pass>());
        Pass& p = *passes_.back();
        p.name = std::string(name);
        p.data = PassData{};
        return {p.builder, std::any_cast<PassData>(&p.data)};
    }
    auto execute() -> void;                                              // render_graph.cpp:567-587 (in order)
    auto executed_pass_names() const -> std::vector<std::string> const& { return executed_; }

private:
    struct Pass { std::string name; std::any data; ComputePassBuilder builder; };
    std::vector<std::unique_ptr<Pass>> passes_;
    std::vector<std::string> executed_;
    uint32_t next_texture_ = 0;
};

// IRenderer trait (include/bisemutum/graphics/renderer.hpp:10-23)
struct IRenderer {
    virtual ~IRenderer() = default;
    virtual auto override_volume_component_name() const -> std::string_view = 0;
    virtual auto prepare_renderer_per_frame_data() -> void = 0;
    virtual auto prepare_renderer_per_camera_data(Camera const& camera) -> void = 0;
    virtual auto render_camera(Camera const& camera, RenderGraph& rg) -> void = 0;
};

} // namespace gfx

// ---- scene_basic light components (include/bisemutum/scene_basic/light.hpp) ---------------------
struct LightTransform {                   // what collect_all_lights reads from object->world_transform()
    float3 translation{0.0f, 0.0f, 0.0f};
    float rotation[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // row-major 3x3, no scaling
    auto transform_direction_without_scaling(float3 d) const -> float3 {
        return {(rotation[0] * d.x + rotation[1] * d.y) + rotation[2] * d.z, (rotation[3] * d.x + rotation[4] * d.y) + rotation[5] * d.z,
                (rotation[6] * d.x + rotation[7] * d.y) + rotation[8] * d.z};
    }
};
struct DirectionalLightComponent final { float3 color = float3{1.0f}; float strength = 1.0f; bool cast_shadow = false; };
struct PointLightComponent final {
    float3 color = float3{1.0f}; float strength = 1.0f; float range = 30.0f;
    bool spot = false; float spot_inner_angle = 30.0f; float spot_outer_angle = 60.0f; bool cast_shadow = false;
};
struct RectLightComponent final { float3 color = float3{1.0f}; float strength = 1.0f; float width = 1.0f; float height = 1.0f; bool two_sided = false; };

// ---- LightsContext (src/renderer/context/lights.{hpp,cpp}) -------------------------------------
struct LightsContext final {
    auto clear() -> void { dir_lights.clear(); point_lights.clear(); rect_lights.clear(); }
    auto add(DirectionalLightComponent const& light, LightTransform const& transform) -> void;   // lights.cpp:52-63
    auto add(PointLightComponent const& light, LightTransform const& transform) -> void;         // lights.cpp:125-143
    auto add(RectLightComponent const& light, LightTransform const& transform) -> void;          // lights.cpp:208-229
    auto set_ltc_luts(bpt_ltc_luts const& luts) -> void { ltc_luts = luts; }
    std::vector<bpt_dir_light_data> dir_lights;
    std::vector<bpt_point_light_data> point_lights;
    std::vector<bpt_rect_light_data> rect_lights;
    bpt_ltc_luts ltc_luts{};
};

// ---- SkyboxContext + current skybox (src/renderer/context/skybox.hpp, src/scene_basic/skybox.cpp) --
struct SkyboxContext final {
    float skybox_transform[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float3 color{1.0f, 1.0f, 1.0f};
    float const* faces_rgba32f = nullptr;   // 6 faces, Vulkan order; nullptr = black 1x1 (skybox.cpp default)
    uint32_t face_size = 0;
};

struct BasicRenderer {
    struct PathTracingSettings final {     // include/bisemutum/renderer/basic.hpp:76-81
        float ray_length = 100.0f;
        uint32_t max_bounces = 3;
        bool denoise = true;                // declared, never read
        bool accumulate = true;
    };
};

// ---- the pass ------------------------------------------------------------------------------------
// Drop-in for bi::PathTracingPass (src/renderer/pass/path_tracing.hpp:12-69): same two methods.
struct PathTracingPass final {
    struct InputData final { gfx::AccelerationStructureHandle scene_accel; };
    struct GBufferTextures final { gfx::TextureHandle base_color, normal_roughness, fresnel, material_0; };   // src/renderer/pass/gbuffer.hpp:9-17
    struct OutputData final { gfx::TextureHandle color; gfx::TextureHandle depth; gfx::TextureHandle velocity; GBufferTextures gbuffer; };

    explicit PathTracingPass(bpt_context* ctx) : ctx_(ctx) {}

    auto update_params(LightsContext& lights_ctx, SkyboxContext& skybox_ctx, BasicRenderer::PathTracingSettings const& settings) -> void;
    auto render(gfx::Camera const& camera, gfx::RenderGraph& rg, InputData const& input,
                BasicRenderer::PathTracingSettings const& settings) -> OutputData;

    // frames accumulated for `camera` so far (the `frame_count` of path_tracing.cpp:239-246,473)
    auto accumulated_frames(gfx::Camera const& camera) const -> uint64_t;
    auto last_status() const -> bpt_status { return status_; }
    auto context() const -> bpt_context* { return ctx_; }
    // Content of OutputData.depth / .gbuffer for `camera`'s current frame (the reference writes them in its first trace
    // pass and "PT Depth", path_tracing.cpp:351-385,421-435; here they are produced on demand by bpt_render_primary).
    auto read_primary_outputs(gfx::Camera const& camera, BasicRenderer::PathTracingSettings const& settings, float* depth, bpt_gbuffer_texel* gbuffer) -> bpt_status;

private:
    struct CameraHistoryInfo final { uint64_t last_frame; uint64_t frame_count; float4x4 proj_view; uint32_t width; uint32_t height; };
    std::unordered_map<gfx::Camera const*, CameraHistoryInfo> camera_history_infos_;
    bpt_context* ctx_;
    bpt_status status_ = BPT_OK;
    bool sky_uploaded_ = false; float const* last_sky_faces_ = nullptr; uint32_t last_sky_size_ = 0;   // see update_params
    uint64_t frame_counter_ = 0;      // stands in for g_engine->window()->frame_count()
    uint32_t prefetch_frames_ = 8;    // samples traced per wave when the history is valid (clamped by the library)
    bpt_settings ahead_settings_{};
    void* color_target_ = nullptr;
public:
    auto set_frame_count(uint64_t f) -> void { frame_counter_ = f; }
    auto set_prefetch_frames(uint32_t n) -> void { prefetch_frames_ = n ? n : 1; }
    // The device memory behind OutputData.color (rgba16_sfloat, W x H x 8 bytes) — in the real engine the CUDA mapping of the render-graph
    // texture. When set, render() writes the frame's accumulated colour there in the same launch that folds the
    // frame's sample into the history; nullptr: the caller resolves separately (bpt_resolve_device*).
    auto set_color_target(void* device_rgba16f) -> void { color_target_ = device_rgba16f; }
    // Forgets every camera's history (what destroying and re-creating the reference's pass does): the next render() of any camera
    // starts a new accumulation (frame_count = 1) and clears the image, which also drops samples traced ahead.
    auto reset_history() -> void { camera_history_infos_.clear(); }
};

// ---- the step after it ------------------------------------------------------------------------------
// PostProcessVolume (include/bisemutum/renderer/post_process_volume.hpp:12-16): the fields the pass reads.
struct PostProcessVolume final {
    bool bloom = false;
    float bloom_threshold = 1.5f;
    float bloom_threshold_softness = 0.5f;
};
// Drop-in for bi::PostProcessPass (src/renderer/pass/post_process.hpp; render: post_process.cpp:92-273): the same render()
// signature; records ONE render-graph pass whose lambda calls bpt_post_process (the reference records 2 + 9 passes with bloom).
struct PostProcessPass final {
    struct InputData final { gfx::TextureHandle color; gfx::TextureHandle depth; };

    explicit PostProcessPass(bpt_context* ctx) : ctx_(ctx) {}
    auto render(gfx::Camera const& camera, gfx::RenderGraph& rg, InputData const& input) -> void;

    PostProcessVolume default_volume_;                 // rt::find_volume_component_for(camera.position, default_volume_)
    auto set_output(float* back_buffer_rgba32f, uint64_t accumulated_frames) -> void { out_ = back_buffer_rgba32f; frames_ = accumulated_frames; }
    auto last_status() const -> bpt_status { return status_; }

private:
    bpt_context* ctx_;
    float* out_ = nullptr;            // stands in for rg.import_back_buffer()
    uint64_t frames_ = 1;             // samples in the accumulation buffer (PathTracingPass::accumulated_frames)
    bpt_status status_ = BPT_OK;
};

// Drop-in for bi::ReblurPass (src/renderer/pass/reblur.hpp:9-47; render: reblur.cpp:273-588), the denoiser ReflectionPass::render calls
// on the ray-traced reflection (reflection.cpp:529): the same InputData and render() signature; records ONE render-graph pass whose
// lambda calls bpt_denoise_reblur (the reference records eight). The textures the handles name are the host arrays given to
// bind_textures() — the engine-side binding hands over its own depth / G-buffer / velocity / validation / reflection images there.
struct ReblurPass final {
    struct InputData final {
        gfx::TextureHandle velocity;
        gfx::TextureHandle depth;
        PathTracingPass::GBufferTextures gbuffer;
        gfx::TextureHandle history_validation;
        gfx::TextureHandle hit_positions_tex;
        gfx::TextureHandle noised_tex;
    };
    explicit ReblurPass(bpt_context* ctx) : ctx_(ctx) {}
    auto render(gfx::Camera const& camera, gfx::RenderGraph& rg, InputData const& input) -> gfx::TextureHandle;

    // host images behind the handles of InputData (`width` x `height` = extent of noised_tex: the camera's, or half of it) and the result
    auto bind_textures(bpt_reblur_inputs const& textures, float* denoised_rgba32f) -> void { textures_ = textures; out_ = denoised_rgba32f; }
    auto set_frame_count(uint64_t f) -> void { frame_counter_ = f; }      // stands in for g_engine->window()->frame_count()
    auto last_status() const -> bpt_status { return status_; }

private:
    bpt_context* ctx_;
    bpt_reblur_inputs textures_{};
    float* out_ = nullptr;
    uint64_t frame_counter_ = 0;
    bpt_status status_ = BPT_OK;
};

// ---- renderer level: the plugin the engine selects with `renderer = "..."` in project.toml -----------------------------
// An IRenderer (include/bisemutum/graphics/renderer.hpp:10-23) that runs BasicRenderer's path-tracing pipeline
// (src/renderer/basic.cpp:31-48 per frame; :157-166 and :228-231 per camera) through the two CUDA passes above. Registered and
// selected like the reference's renderers: GraphicsManager::register_renderer<T>() needs T::renderer_type_name
// (graphics_manager.hpp:82-87, src/engine/register_renderer.cpp:9-11), set_renderer(name) instantiates it (engine.cpp:147-148).
struct CudaPathTracingRenderer final : gfx::IRenderer {
    static constexpr std::string_view renderer_type_name = "CudaPathTracingRenderer";
    explicit CudaPathTracingRenderer(bpt_context* ctx) : path_tracing_pass(ctx), post_process_pass(ctx) {}

    auto override_volume_component_name() const -> std::string_view override { return "BasicRendererOverrideVolume"; }   // basic.cpp:268-270
    auto prepare_renderer_per_frame_data() -> void override;
    auto prepare_renderer_per_camera_data(gfx::Camera const& camera) -> void override;
    auto render_camera(gfx::Camera const& camera, gfx::RenderGraph& rg) -> void override;
    auto last_status() const -> bpt_status;

    struct Settings final { BasicRenderer::PathTracingSettings path_tracing; } settings;   // BasicRendererOverrideVolume::settings
    PostProcessVolume post_process;
    LightsContext lights_ctx;
    SkyboxContext skybox_ctx;
    gfx::AccelerationStructureHandle scene_accel;
    float* back_buffer = nullptr;                     // W*H rgba32f, stands in for rg.import_back_buffer()
    PathTracingPass path_tracing_pass;
    PostProcessPass post_process_pass;
};

namespace gfx {
struct GraphicsManager final {
    explicit GraphicsManager(bpt_context* ctx) : ctx_(ctx) {}
    template <typename Renderer>
    auto register_renderer() -> void {
        renderer_creators_[std::string(Renderer::renderer_type_name)] = [](bpt_context* c) -> std::unique_ptr<IRenderer> { return std::make_unique<Renderer>(c); };
    }
    auto set_renderer(std::string_view name) -> bool;
    auto renderer() -> IRenderer* { return renderer_.get(); }
private:
    bpt_context* ctx_;
    std::unordered_map<std::string, std::function<std::unique_ptr<IRenderer>(bpt_context*)>> renderer_creators_;
    std::unique_ptr<IRenderer> renderer_;
};
} // namespace gfx

} // namespace bi
