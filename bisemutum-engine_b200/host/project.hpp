// host/project.hpp — headless loader for a bisemutum project directory: what the reference reads through its
// asset manager and ECS deserialisation, restated without the engine runtime so the reference's own example project can be
// handed to the CUDA path tracer:
//   .biasset containers     src/runtime/byte_stream.{hpp,cpp} (length-prefixed vectors, zlib "compressed part"),
//                           src/graphics/mesh.cpp:127-146 (StaticMesh v1/v2), src/scene_basic/texture.cpp:83-137 (Texture v1 raw / v2)
//   asset_metadata.toml     [[assets]] id / path / type (asset ids resolve through it)
//   materials/*.toml        src/scene_basic/material.cpp:12-90: surface_model, blend_mode, material_function (HLSL text), [[params]]
//   scene.toml              [[objects]] with [[objects.components]]: Transform, CameraComponent, *LightComponent, StaticMeshComponent,
//                           MeshRendererComponent, BasicRendererOverrideVolume (settings.path_tracing / ambient_occlusion)
// Materials are HLSL snippets in the reference; the loader recognises the closed set of snippets the CUDA kernels restate
// (include/bpt/bpt.h: BPT_MATERIAL_KIND_*) and fails loudly on anything else. glTF import (import_model.cpp:27-430) is host/gltf.hpp; assimp import is not built.
#pragma once
#include <string>
#include <utility>
#include <vector>
#include "engine.hpp"

namespace bi::project {

// ---- a TOML subset: tables, arrays of tables, dotted headers, strings ('..', "..", '''..'''), numbers, booleans, arrays, inline tables
struct Toml {
    enum class Kind { nil, boolean, number, string, array, table } kind = Kind::nil;
    bool b = false;
    double num = 0.0;
    bool is_int = false; uint64_t u64 = 0;               // an integer token keeps all 64 bits (asset ids are 2^62 + k: a double cannot hold them)
    std::string str;
    std::vector<Toml> arr;
    std::vector<std::pair<std::string, Toml>> tab;

    auto find(std::string_view key) const -> Toml const*;
    auto at(std::string_view dotted_path) const -> Toml const*;       // "a.b.c" through tables
    auto number_or(std::string_view path, double dflt) const -> double;
    auto id_or(std::string_view path, uint64_t dflt) const -> uint64_t;   // exact for integer tokens; a non-negative whole double otherwise
    auto string_or(std::string_view path, std::string dflt) const -> std::string;
    auto bool_or(std::string_view path, bool dflt) const -> bool;
};
auto parse_toml(std::string const& text, Toml& out, std::string& err) -> bool;

// ---- .biasset
struct StaticMeshData {
    std::vector<float> positions, normals, tangents, colors, texcoords, texcoords2;   // 3 / 3 / 4 / 3 / 2 / 2 floats per vertex
    std::vector<uint32_t> indices;
    struct Submesh { uint32_t base_vertex, index_offset, num_indices; uint8_t topology; };
    std::vector<Submesh> submeshes;
};
struct TextureData {
    uint32_t width = 0, height = 0, depth = 0, levels = 0;
    uint8_t format = 0, dim = 0;                         // rhi::ResourceFormat value (37 rgba8_unorm, 43 rgba8_srgb, 109 rgba32_sfloat, ...)
    uint8_t mag_filter = 0, min_filter = 0, address_u = 0, address_v = 0;   // rhi::SamplerDesc enums (sampler.hpp:10-26)
    std::vector<uint8_t> texels;                         // level 0 ... as stored
};
auto load_static_mesh(std::string const& path, StaticMeshData& out, std::string& err) -> bool;
auto load_texture(std::string const& path, TextureData& out, std::string& err) -> bool;

// ---- a loaded project, already in the layout the C ABI takes
struct Project {
    // geometry streams (gpu_scene_data.hpp:10-24), one BLAS per (mesh, submesh), one drawable + instance per rendered submesh
    std::vector<float> positions, normals, tangents, texcoords;
    std::vector<uint32_t> indices;
    std::vector<bpt_blas_desc> blas;
    std::vector<bpt_drawable_sbt_data> drawables;
    std::vector<uint32_t> drawable_va;
    std::vector<bpt_instance_desc> instances;
    std::vector<bpt_material> materials;
    std::vector<TextureData> textures;                   // index = texture index used by the materials
    LightsContext lights;
    // camera (CameraComponent + Transform; looks down local -Z, up +Y: camera_system.cpp:17-29)
    float cam_position[3] = {0, 0, 0}, cam_front[3] = {0, 0, -1}, cam_up[3] = {0, 1, 0};
    float yfov = 30.0f, near_z = 0.001f, far_z = 1e5f;
    uint32_t target_width = 0, target_height = 0;
    bool orthographic = false;
    // BasicRendererOverrideVolume settings
    BasicRenderer::PathTracingSettings path_tracing{};
    bpt_ao_settings ambient_occlusion{0.5f, 0.5f, 1u};
    std::vector<std::string> object_names;               // of the drawables, in order
};
auto load_project(std::string const& dir, Project& out, std::string& err) -> bool;
// geometry, materials + textures, instances, lights, (black) sky, acceleration structure
auto upload_project(Project const& p, bpt_context* ctx, uint32_t accel_mode, std::string& err) -> bpt_status;

} // namespace bi::project
