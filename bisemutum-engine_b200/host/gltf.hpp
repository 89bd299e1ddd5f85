// host/gltf.hpp — headless glTF 2.0 import (BASELINE configs[0] "Cornell-box glTF"): what the reference's editor
// action `menu_action_import_model_gltf` (src/scene_basic/menu_actions/import_model.cpp:27-430) does through tinygltf + the asset manager +
// the ECS, restated without them so a .gltf / .glb file becomes the C ABI's arrays:
//   textures   import_model.cpp:60-142   RGBA8 unorm texels (1-3 channel images widened as the reference does), sampler filter / wrap modes
//   materials  import_model.cpp:144-232  the glTF metallic-roughness template = BPT_MATERIAL_KIND_GLTF_PBR, blend mode opaque ("TODO - alpha mode")
//   meshes     import_model.cpp:234-358  one StaticMesh per glTF mesh, one submesh (= one BLAS) per TRIANGLES primitive; POSITION / NORMAL /
//                                        TEXCOORD_0 must be float, a missing attribute is zero-filled; TANGENT in the file is ignored
//   tangents   static_mesh.cpp:93-152    StaticMesh::calculate_tspace = MikkTSpace `genTangSpaceDefault` per submesh (see mikk_tangents below)
//   nodes      import_model.cpp:360-421  TRS / matrix -> Transform (math/transform.cpp:11-52), world = parent * local through
//                                        Transform::from_matrix in FP32 (runtime/scene_object.cpp:31-52); one drawable + instance per primitive
// Third-party code the reference calls here and that is NOT under /root/reference: tinygltf (JSON + buffers + stb_image decoding; unpinned xmake
// package) — replaced by the small JSON reader and the PNG (8-bit, non-interlaced) decoder in gltf.cpp, any other image encoding fails loudly;
// MikkTSpace (unpinned) — restated from its published algorithm, parity unpinned (no reference test or fixture holds tangents of an import).
#pragma once
#include <string>
#include <vector>
#include "project.hpp"

namespace bi::project {

// MikkTSpace `genTangSpaceDefault` (angular threshold 180 degrees) for one indexed triangle list, as driven by StaticMesh::calculate_tspace:
// the callbacks read positions / normals / texcoords of mesh vertex `base_vertex + indices[..]` and `m_setTSpaceBasic` writes
// (tangent.xyz, sign) to that same mesh vertex, so for a vertex shared by several corners the LAST corner in face order wins.
// positions / normals: 3 floats per vertex, texcoords: 2, tangents (out): 4; all indexed from `base_vertex`.
auto mikk_tangents(const float* positions, const float* normals, const float* texcoords, float* tangents,
                   const uint32_t* indices, size_t num_indices, uint32_t base_vertex) -> void;

// PNG, 8 bits per channel, non-interlaced (the subset stb_image is asked for here). force_rgba = true: tinygltf's default `req_comp = 4`
// (grey -> (g, g, g, 255), grey+alpha -> (g, g, g, a), RGB -> (r, g, b, 255)); false: stbi_load_from_memory(..., req_comp = 0), the
// file's own channel count (palettes expanded to 3 or 4), as TextureAsset::load uses it for PNG-per-layer storage (texture.cpp:110-131).
auto decode_png(std::string const& file, uint32_t& width, uint32_t& height, uint32_t& channels, std::vector<uint8_t>& pixels, bool force_rgba, std::string& err) -> bool;

// Appends the model to `out` (geometry streams, BLAS descs, drawables + instances, materials, textures, object names).
// Camera, lights and renderer settings are not part of a glTF import in the reference either (the importer ignores glTF cameras / lights).
auto import_gltf(std::string const& path, Project& out, std::string& err) -> bool;

} // namespace bi::project
