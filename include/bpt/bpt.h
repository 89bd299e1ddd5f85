/*
 * bpt.h — C ABI of the B200-native wavefront path tracer ("bpt").
 *
 * This is the drop-in boundary for ONE hot path of PepcyCh/bisemutum-engine: the wavefront
 * path-tracing pass (`PathTracingPass`, bisemutum/src/renderer/pass/path_tracing.{hpp,cpp}).
 * A `CudaPathTracingPass` with the same two methods (`update_params`, `render`) calls only the
 * functions below; see INTEGRATION.md for the reference-side binding.
 *
 * Conventions
 *  - Every function returns a bpt_status (0 = ok). No exceptions or longjmp cross this boundary.
 *  - One bpt_context per GPU. A context is NOT thread-safe; all work is ordered on one CUDA
 *    stream (bpt_set_stream) and is asynchronous until bpt_sync / a host-reading call.
 *  - All pointers are HOST pointers unless the parameter name ends in `_device`.
 *  - Matrices are column-major float[16] exactly as glm uploads them (reference:
 *    bisemutum/src/graphics/camera.cpp:96-108). 3x4 instance transforms are row-major
 *    (reference: bisemutum/include/bisemutum/rhi/accel.hpp:48-55).
 *  - Struct byte layouts marked "REF layout" are byte-identical to the reference structs so
 *    that the engine can hand over its CPU-side vectors without repacking.
 *
 * There is no CPU fallback: every entry point that computes needs a CUDA device and fails
 * with BPT_ERR_NO_DEVICE / BPT_ERR_CUDA otherwise.
 */
#ifndef BPT_H_
#define BPT_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define BPT_API __declspec(dllexport)
#else
#define BPT_API __attribute__((visibility("default")))
#endif

typedef struct bpt_context bpt_context;

typedef enum bpt_status {
    BPT_OK = 0,
    BPT_ERR_INVALID = 1,   /* bad argument / inconsistent scene description            */
    BPT_ERR_CUDA = 2,      /* a CUDA runtime call failed (see bpt_last_error)           */
    BPT_ERR_OOM = 3,       /* device allocation failed                                  */
    BPT_ERR_STATE = 4,     /* call order violated (e.g. render before build_accel)      */
    BPT_ERR_NO_DEVICE = 5, /* no CUDA device visible — there is no CPU path             */
    BPT_ERR_UNSUPPORTED = 6
} bpt_status;

/* ---------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------- */
/* Bulk data pointers (images, ray batches, per-ray results, atlases, G-buffers: every `const float*` / `float*` / texel-array
 * parameter of the pass-level entry points below) may be HOST or DEVICE pointers: copies use unified addressing. A host that
 * already holds its depth / G-buffer / results on the GPU (external memory, another CUDA library, an NCCL collective) therefore
 * pays no PCIe round trip; scene-description uploads (geometry, materials, lights, instances) are host arrays. */
typedef struct bpt_config {
    int32_t device;      /* CUDA ordinal */
    uint32_t width;      /* camera target extent; reference: path_tracing.cpp:228-229 */
    uint32_t height;
    uint32_t max_lights_per_vertex; /* sizing hint for the shadow-ray queue; 0 = derive at upload_lights */
} bpt_config;

BPT_API bpt_status bpt_create(const bpt_config* cfg, bpt_context** out_ctx);
BPT_API bpt_status bpt_destroy(bpt_context* ctx);
/* Footprint control for a pass that shares the GPU with an engine. A render call keeps `samples per wave` = min(64, max_paths / pixels,
 * 8 GiB / (pixels * lights * 48 B)) samples of every pixel in flight; the wavefront state is ~132 B per path plus 48 B per path and light
 * for the shadow-ray queue (default max_paths = 2^26: 32 samples and ~9 GB at 1080p; 2^23 = 4 samples, ~1.1 GB, about 15 % slower —
 * DESIGN.md section 5). 0 restores the default. Re-allocates the wave buffers and drops samples traced ahead. */
BPT_API bpt_status bpt_set_wave_budget(bpt_context* ctx, uint64_t max_paths_in_flight);
BPT_API const char* bpt_last_error(const bpt_context* ctx);
BPT_API const char* bpt_version(void);
/* All later work is enqueued on `cuda_stream` (a cudaStream_t; NULL = legacy default stream). */
BPT_API bpt_status bpt_set_stream(bpt_context* ctx, void* cuda_stream);
BPT_API bpt_status bpt_sync(bpt_context* ctx);
/* Re-allocates the per-path wavefront state for a new target extent (history is dropped,
 * as the reference does when width/height change: path_tracing.cpp:231-246). */
BPT_API bpt_status bpt_resize(bpt_context* ctx, uint32_t width, uint32_t height);

/* ---------------------------------------------------------------------------------------
 * Scene: geometry.  Replaces GpuSceneSystem / mesh upload / SBT fill:
 *   bisemutum/src/graphics/gpu_scene_data.hpp:10-24 (flat float/uint streams)
 *   bisemutum/src/graphics/drawable_stb_data.hpp:7-17 (DrawableSbtData, 36 B)
 *   bisemutum/src/graphics/graphics_manager.cpp:1314-1330 (offsets are ELEMENT offsets into
 *   the streams; material_offset is a BYTE offset into the material-params buffer)
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_geometry_streams {
    const float* positions;   uint64_t num_position_floats;   /* 3 per vertex */
    const float* normals;     uint64_t num_normal_floats;     /* 3 per vertex */
    const float* tangents;    uint64_t num_tangent_floats;    /* 4 per vertex */
    const float* colors;      uint64_t num_color_floats;      /* 3 per vertex, may be NULL */
    const float* texcoords;   uint64_t num_texcoord_floats;   /* 2 per vertex */
    const float* texcoords2;  uint64_t num_texcoord2_floats;  /* 2 per vertex, may be NULL */
    const uint32_t* indices;  uint64_t num_indices;
} bpt_geometry_streams;

/* REF layout: DrawableSbtData, bisemutum/shaders/core/raytracing/hit.hlsl:7-17. */
typedef struct bpt_drawable_sbt_data {
    uint32_t drawable_index;
    uint32_t position_offset;
    uint32_t normal_offset;
    uint32_t tangent_offset;
    uint32_t color_offset;
    uint32_t texcoord_offset;
    uint32_t texcoord2_offset;
    uint32_t index_offset;
    uint32_t material_offset; /* bytes; multiple of sizeof(bpt_material) */
} bpt_drawable_sbt_data;

/* Vertex-attribute presence mask of a drawable (VERTEX_ATTRIBUTES_IN, hit.hlsl:35-149). */
enum {
    BPT_VA_POSITION = 1, BPT_VA_NORMAL = 2, BPT_VA_TANGENT = 4,
    BPT_VA_COLOR = 8, BPT_VA_TEXCOORD = 16, BPT_VA_TEXCOORD2 = 32
};

/* One bottom-level structure per (mesh id, submesh): graphics_manager.cpp:616-654.
 * Offsets are element offsets into the positions (floats) / indices streams. */
typedef struct bpt_blas_desc {
    uint32_t position_offset; /* floats; = mesh positions offset + base_vertex*3 */
    uint32_t index_offset;    /* uints  */
    uint32_t num_triangles;   /* min(submesh.num_indices, mesh.num_indices)/3 */
    uint32_t reserved;
} bpt_blas_desc;

BPT_API bpt_status bpt_scene_upload_geometry(
    bpt_context* ctx, const bpt_geometry_streams* streams,
    const bpt_drawable_sbt_data* drawables, const uint32_t* drawable_vertex_attributes /* may be NULL = pos|normal|tangent|texcoord */,
    uint32_t num_drawables,
    const bpt_blas_desc* blas, uint32_t num_blas);

/* REF layout (64 B): AccelerationStructureInstanceDesc, rhi/accel.hpp:48-55, filled as
 * bisemutum/src/graphics/accel.cpp:104-132: instance_id = sbt_offset = continuous drawable
 * index, mask 0xff, flags = force_opaque (4) iff blend_mode == opaque else force_non_opaque (8).
 * `blas` holds the index into the bpt_blas_desc array instead of a GPU address. */
typedef struct bpt_instance_desc {
    float transform[3][4];
    uint32_t instance_id_and_mask;   /* instance_id : 24 | mask : 8  */
    uint32_t sbt_offset_and_flags;   /* sbt_offset  : 24 | flags : 8 */
    uint64_t blas;
} bpt_instance_desc;
enum { BPT_INSTANCE_FORCE_OPAQUE = 4, BPT_INSTANCE_FORCE_NON_OPAQUE = 8 };

BPT_API bpt_status bpt_scene_upload_instances(bpt_context* ctx, const bpt_instance_desc* instances, uint32_t num_instances);

/* ---------------------------------------------------------------------------------------
 * Scene: materials.  The reference's materials are HLSL snippets spliced at $MATERIAL_FUNCTION
 * (hit.hlsl:166-173).  CUDA cannot run them, so a closed set of "material kinds" is offered,
 * each the restatement of one snippet the reference ships.
 * ------------------------------------------------------------------------------------- */
enum {
    BPT_MATERIAL_KIND_GLTF_PBR = 0,        /* import_model.cpp:208-230 */
    BPT_MATERIAL_KIND_ASSIMP_DIFFUSE = 1,  /* import_model.cpp:490-493 (base_color, roughness) */
    BPT_MATERIAL_KIND_DEFAULT = 2,         /* surface_data_default, material/utils.hlsl:18-31 */
    /* the five materials of the reference's example project (the .toml files under examples/scene_basic/materials) */
    BPT_MATERIAL_KIND_CONSTANT_COLOR = 3,  /* white.toml: base_color = PARAM_base_color */
    BPT_MATERIAL_KIND_CHECKERBOARD = 4,    /* checkerboard.toml: world-space xz checker. base_color.rgb = base_color_0,
                                              base_color.a = roughness_0, emission.rgb = base_color_1, roughness = roughness_1 */
    BPT_MATERIAL_KIND_TEXTURED = 5,        /* textured.toml: base_color_tex, normal_map_tex (raw texel), roughness */
    BPT_MATERIAL_KIND_TRANSPARENT = 6,     /* transparent.toml: base_color.rgb, opacity = base_color.a, two-sided */
    BPT_MATERIAL_KIND_CAGE = 7,            /* cage.toml: base = f0 = tex.rgb, opacity = tex.a < 0.5 ? 0 : 1, two-sided */
    /* `surface.base_color = vertex.color;` — the one consumer of the vertex-colour stream (VA_TYPE_COLOR, core/raytracing/hit.hlsl:97-113).
     * The hit shader interpolates it with its THIRD corner read at index.x instead of index.z (:108-112); reproduced. roughness = material's. */
    BPT_MATERIAL_KIND_VERTEX_COLOR = 8
};
enum { BPT_SURFACE_MODEL_UNLIT = 0, BPT_SURFACE_MODEL_LIT = 1 };         /* material.hlsl:3-5 */
enum { BPT_BLEND_OPAQUE = 0, BPT_BLEND_ALPHA_TEST = 1, BPT_BLEND_TRANSLUCENT = 2 };

#define BPT_MATERIAL_FLAG_TWO_SIDED 1u
#define BPT_MATERIAL_KIND_SHIFT 8
#define BPT_MATERIAL_BLEND_SHIFT 16
#define BPT_MATERIAL_MODEL_SHIFT 24

/* 64-byte, 16-byte-aligned parameter record (the reference aligns records to 16 B in a byte
 * buffer: graphics_manager.cpp:221-256). Texture indices: -1 = the importer's 1x1 default
 * (white1x1 / normal1x1, import_model.cpp:186-203). */
typedef struct bpt_material {
    float base_color[4];
    float emission[3];   /* carried for API fidelity; the PT G-buffer drops it (gbuffer.hlsl:18-33) */
    float roughness;
    float metallic;
    float normal_map_scale;
    float occlusion_strength;
    uint32_t flags;      /* two_sided | kind<<8 | blend_mode<<16 | surface_model<<24 */
    int32_t base_color_tex;
    int32_t metallic_roughness_tex;
    int32_t normal_map_tex;
    int32_t occlusion_tex;
} bpt_material;

/* RGBA8_SRGB (rhi format rgba8_srgb) is decoded to linear FP32 texels at upload (256-entry table), then filtered. */
enum { BPT_TEXTURE_RGBA8_UNORM = 0, BPT_TEXTURE_RGBA32_FLOAT = 1, BPT_TEXTURE_RGBA8_SRGB = 2 };
enum { BPT_ADDRESS_REPEAT = 0, BPT_ADDRESS_CLAMP = 1 };
typedef struct bpt_texture_desc {
    const void* texels;   /* level 0 only (hit shaders have no derivatives → level 0) */
    uint32_t width, height;
    uint32_t format;
    uint32_t address_mode_u, address_mode_v;
    uint32_t filter_linear; /* 0 = nearest, 1 = bilinear (explicit FP32 lerp, no HW filtering) */
} bpt_texture_desc;

BPT_API bpt_status bpt_scene_upload_materials(
    bpt_context* ctx, const bpt_material* materials, uint32_t num_materials,
    const bpt_texture_desc* textures, uint32_t num_textures);

/* ---------------------------------------------------------------------------------------
 * Scene: lights.  REF layouts of bisemutum/src/renderer/context/lights.hpp:12-53
 * (= shaders/renderer/lights_struct.hlsl:3-53), as packed by LightsContext::collect_all_lights
 * (lights.cpp:48-244).
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_dir_light_data {      /* 64 B */
    float emission[3]; int32_t sm_index;
    float direction[3]; float shadow_strength;   /* direction points TO the light */
    float cascade_shadow_radius_sqr[4];
    float shadow_depth_bias; float shadow_normal_bias; float _pad1[2];
} bpt_dir_light_data;

typedef struct bpt_point_light_data {    /* 64 B */
    float emission[3]; float range_sqr_inv;
    float position[3]; float cos_inner;
    float direction[3]; float cos_outer;
    int32_t sm_index; float shadow_strength; float shadow_depth_bias; float shadow_normal_bias;
} bpt_point_light_data;

typedef struct bpt_rect_light_data {     /* 112 B */
    float emission[3]; int32_t texture_index;
    float center_position[3]; uint32_t two_sided;
    float position0[3]; float inv_width_sqr;
    float position1[3]; float inv_height_sqr;
    float position2[3]; float inv_texel_size;
    float position3[3]; float _pad1;
    float normal[3]; float _pad2;
} bpt_rect_light_data;

/* LTC look-up tables, bisemutum/assets/textures/ltc_*.biasset: 8x8x64 texels each,
 * matrix luts rgba32f (65536 B each), norm lut rg32f (32768 B). May be NULL iff num_rect == 0. */
typedef struct bpt_ltc_luts {
    const float* matrix_lut0; const float* matrix_lut1; const float* matrix_lut2; /* 8*8*64*4 floats */
    const float* norm_lut;                                                          /* 8*8*64*2 floats */
} bpt_ltc_luts;

/* Rect-light textures (RectLightComponent::texture, bisemutum/src/renderer/context/lights.cpp:229-239; sampled by
 * rect_light_sample_texture, shaders/renderer/lights.hlsl:425-447, at the two call sites :495-511). Texture k is the one
 * bpt_rect_light_data::texture_index == k refers to (the reference binds up to 16, lights.hpp:65); a light whose index has no
 * texture here is evaluated untextured. `texels` is level 0; with levels > 1 the chain is generated as the engine generates it
 * when a TextureAsset is uploaded (scene_basic/texture.cpp:176 -> GraphicsManager::generate_mipmaps_2d ->
 * shaders/core/mipmap.hlsl, MIPMAP_MODE_AVG, each level stored in the texture's own format). SampleLevel(level) filters inside a level
 * as `filter_linear` says and between levels as `mip_linear` says (rhi::SamplerDesc::mipmap_mode; nearest = the Vulkan rule
 * ceil(level + 0.5) - 1), all in explicit FP32 arithmetic. Call before or after bpt_scene_upload_lights; 0 textures clears them. */
typedef struct bpt_light_texture_desc {
    const void* texels;
    uint32_t width, height;
    uint32_t format;        /* BPT_TEXTURE_* */
    uint32_t levels;        /* >= 1; clamped to floor(log2(max(width, height))) + 1 as generate_mipmaps_2d does */
    uint32_t address_mode_u, address_mode_v;
    uint32_t filter_linear; /* 0 = nearest, 1 = bilinear inside a level */
    uint32_t mip_linear;    /* 0 = nearest level, 1 = linear between levels */
} bpt_light_texture_desc;
#define BPT_MAX_RECT_LIGHT_TEXTURES 16u
BPT_API bpt_status bpt_scene_upload_light_textures(bpt_context* ctx, const bpt_light_texture_desc* textures, uint32_t num_textures);
/* Debug: the generated chain of light texture `index` as RGBA32F texels, level after level (W*H + (W/2)*(H/2) + ... float4). */
BPT_API bpt_status bpt_debug_read_light_texture(bpt_context* ctx, uint32_t index, float* out_rgba32f, uint64_t capacity_texels, uint64_t* out_texels);

BPT_API bpt_status bpt_scene_upload_lights(
    bpt_context* ctx,
    const bpt_dir_light_data* dir_lights, uint32_t num_dir_lights,
    const bpt_point_light_data* point_lights, uint32_t num_point_lights,
    const bpt_rect_light_data* rect_lights, uint32_t num_rect_lights,
    const bpt_ltc_luts* ltc_luts);

/* Sky: six square rgba32f faces in Vulkan cube order (+X,-X,+Y,-Y,+Z,-Z), skybox_transform is
 * the upper-left 3x3 (row-major here) applied to the miss direction, skybox_color multiplies the
 * lookup (deferred_lighting_secondary.hlsl:24-29; skybox.hpp:9-17). faces == NULL → black 1x1. */
BPT_API bpt_status bpt_scene_upload_sky(
    bpt_context* ctx, const float* faces_rgba32f, uint32_t face_size,
    const float skybox_transform[9], const float skybox_color[3]);
/* Same, transform and colour only: the faces (and the textures bpt_precompute_sky_ibl derived from them) stay. What the
 * reference's per-frame SkyboxContext::update_shader_params changes when the skybox texture itself did not (skybox.cpp:40-50). */
BPT_API bpt_status bpt_scene_update_sky_params(bpt_context* ctx, const float skybox_transform[9], const float skybox_color[3]);

/* ---------------------------------------------------------------------------------------
 * Acceleration structure (replaces AccelerationStructure ctor, accel.cpp:11-159; rebuilt when
 * instances change).  LBVH: 63-bit Morton of AABB centroids in scene bounds → stable LSD radix
 * sort (code, primitive) → Karras-2012 hierarchy → bottom-up refit.
 * ------------------------------------------------------------------------------------- */
enum {
    BPT_ACCEL_TWO_LEVEL = 0, /* one BLAS per bpt_blas_desc (object space) + TLAS over instances */
    BPT_ACCEL_MERGED = 1     /* instances pre-transformed into one world-space BLAS              */
};
BPT_API bpt_status bpt_build_accel(bpt_context* ctx, uint32_t mode);
/* Rebuilds only the TLAS after bpt_scene_upload_instances (the reference rebuilds its TLAS
 * every frame: render_graph.cpp:818-825). Two-level mode only. */
BPT_API bpt_status bpt_update_tlas(bpt_context* ctx);

/* 64-byte node: both children's boxes live in the parent so one node fetch = four 16-B loads.
 * child >= 0: internal node index; child < 0: leaf, ~child = sorted primitive slot. */
typedef struct bpt_bvh_node {
    float c0_lo_x, c0_hi_x, c0_lo_y, c0_hi_y;
    float c1_lo_x, c1_hi_x, c1_lo_y, c1_hi_y;
    float c0_lo_z, c0_hi_z, c1_lo_z, c1_hi_z;
    int32_t child0, child1;
    int32_t parent;     /* -1 for the root */
    uint32_t reserved;
} bpt_bvh_node;

/* which: 0..num_blas-1 = that BLAS (in merged mode only 0 exists), 0xffffffff = the TLAS.
 * Any out pointer may be NULL. `num_prims` receives the leaf count N; arrays need N
 * (morton/prims) and max(N-1,0) (nodes) entries. */
#define BPT_BVH_TLAS 0xffffffffu
BPT_API bpt_status bpt_debug_read_bvh(
    bpt_context* ctx, uint32_t which, uint32_t* num_prims,
    uint64_t* sorted_morton, uint32_t* sorted_prims, bpt_bvh_node* nodes, uint32_t capacity,
    int32_t* root);

/* ---------------------------------------------------------------------------------------
 * Render (replaces the GPU work recorded by PathTracingPass::render, path_tracing.cpp:224-488)
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_camera {
    float matrix_inv_view[16];
    float matrix_inv_proj[16];
    float matrix_proj_view[16]; /* only compared by the host pass for history validity */
} bpt_camera;

enum { BPT_NEE_SHADOW_RAY = 0, BPT_NEE_NONE = 1 };
enum { BPT_RECT_SHADOW_OFF = 0, BPT_RECT_SHADOW_MRP_RAY = 1 };
enum { BPT_STATE_FP32 = 0, BPT_STATE_REFERENCE_FP16 = 1 };

/* BasicRenderer::PathTracingSettings (renderer/basic.hpp:76-81) + the mode switches of
 * SURVEY.md §0. Defaults (all-zero switches) are the parity configuration. All are implemented.
 * state_precision = reference_fp16 applies the reference's texture formats to every value it passes between passes
 * (half ray directions / throughput / colours, the packed G-buffer, the half additive blit, the running half lerp of
 * pt_accumulate.hlsl); the accumulation buffer then follows that rule until bpt_clear_accum (mixing: BPT_ERR_STATE). */
typedef struct bpt_settings {
    float ray_length;        /* 100 */
    uint32_t max_bounces;    /* clamped to [2,16] as path_tracing.cpp:187,290 */
    uint32_t accumulate;     /* honoured by the host pass (history reset); kept for fidelity */
    uint32_t nee_mode;
    uint32_t rect_shadow;
    uint32_t russian_roulette;
    uint32_t pixel_jitter;
    uint32_t state_precision;
} bpt_settings;

/* Zeroes the FP32 accumulation buffer (history invalid). */
BPT_API bpt_status bpt_clear_accum(bpt_context* ctx);
/* Adds `num_samples` samples (frame_index = frame_index_first + s, one reference "frame" each)
 * for every pixel to the FP32 sum buffer. Asynchronous. */
BPT_API bpt_status bpt_render(
    bpt_context* ctx, const bpt_camera* camera, uint32_t frame_index_first, uint32_t num_samples,
    const bpt_settings* settings);
/* Sample prefetch for frame-at-a-time callers (the reference renders ONE sample per engine frame and
 * accumulates while the camera is static, path_tracing.cpp:231-246): bpt_render_ahead traces up to
 * `max_samples` consecutive samples (frame_index_first, +1, ...) in one wave and keeps their per-sample
 * colours; bpt_accumulate_ahead then adds the next `count` of them to the sum buffer, in frame order.
 * The image after k accumulated frames is bit-identical to k calls of bpt_render(.., 1, ..). Any other
 * render / clear / resize call drops the pending samples. `*out_samples` = samples actually traced. */
BPT_API bpt_status bpt_render_ahead(
    bpt_context* ctx, const bpt_camera* camera, uint32_t frame_index_first, uint32_t max_samples,
    const bpt_settings* settings, uint32_t* out_samples);
BPT_API bpt_status bpt_accumulate_ahead(bpt_context* ctx, uint32_t count);
BPT_API bpt_status bpt_pending_ahead(bpt_context* ctx, uint32_t* out_pending, uint32_t* out_next_frame_index);
/* One engine frame in one launch: bpt_accumulate_ahead(ctx, 1) followed by bpt_resolve_device_rgba16f(ctx, total_samples, out) — the next
 * prefetched sample joins the sum and OutputData.color (rgba16_sfloat, path_tracing.cpp:248-252,461-480 "PT Accumulate") is written from it.
 * `total_samples` = frames in the history including this one. Same bits as the two calls. */
BPT_API bpt_status bpt_accumulate_ahead_rgba16f(bpt_context* ctx, uint32_t total_samples, void* out_rgba16f_device);
/* out[p] = (sum[p].rgb * (1/total_samples), 1). Host destination (synchronises). */
BPT_API bpt_status bpt_resolve(bpt_context* ctx, uint32_t total_samples, float* out_rgba32f);
/* Same, written to device memory (no synchronisation). */
BPT_API bpt_status bpt_resolve_device(bpt_context* ctx, uint32_t total_samples, float* out_rgba32f_device);
/* Same image in the format of the reference's OutputData.color (rgba16_sfloat, bisemutum/src/renderer/pass/path_tracing.cpp:248-252):
 * W*H x 4 IEEE halves (8 bytes per pixel), round-to-nearest-even of the FP32 mean, alpha = 1; device memory, no synchronisation. */
BPT_API bpt_status bpt_resolve_device_rgba16f(bpt_context* ctx, uint32_t total_samples, void* out_rgba16f_device);
/* Raw FP32 sum buffer (device pointer, W*H float4) for the multi-GPU reduce. */
BPT_API bpt_status bpt_accum_device_ptr(bpt_context* ctx, float** out_device_ptr);
/* Replaces the sum buffer contents from the host (checkpoint/resume of the history). */
BPT_API bpt_status bpt_upload_accum(bpt_context* ctx, const float* sum_rgba32f);

/* ---------------------------------------------------------------------------------------
 * ReBLUR, the denoiser of the ray-traced reflections: ReblurPass::render
 * (bisemutum/src/renderer/pass/reblur.cpp:273-588, shaders/renderer/reblur/\*.hlsl), called by ReflectionPass::render with the
 * output of the reflection trace (reflection.cpp:529). Eight passes: pre blur, temporal accumulate (surface + virtual-position
 * history), fetch linear depth, gen depth mip, fix history, blur, temporal stabilize, post blur. `noised` / `hit_positions` are what
 * bpt_trace_reflection returns (camera extent, or half of it: the denoiser then runs at half resolution like the reference,
 * reblur.cpp:280); depth (d32: 0 = background), normal_roughness (the G-buffer texture: xy = octahedral normal, w = roughness) and
 * velocity (uv units, NULL = static) have the context's extent; history_validation (denoiser extent, NULL = all valid) is the mask of
 * validate_history.hlsl. The history the reference keeps on the camera (last frame's depth / normal_roughness, the blurred and the
 * stabilised image, the accumulation speed) and last frame's camera matrices are kept by the context; it is used when `frame_count`
 * is last call's + 1 (reblur.cpp:282-285). Every texture the reference stores as rgba16_sfloat / r16_sfloat is rounded to half at
 * the same points. out_rgba32f: the denoised image, denoiser extent. */
typedef struct bpt_reblur_settings {
    uint32_t virtual_history;          /* reblur.cpp:383: 1 */
    float blur_radius;                 /* reblur.cpp:318: 0.9 */
    float anti_flickering_strength;    /* reblur.cpp:319: 3.5 */
    uint32_t _pad;
} bpt_reblur_settings;
typedef struct bpt_reblur_inputs {
    uint32_t width, height;            /* of noised / hit_positions / history_validation / the result */
    const float* noised;               /* rgba32f */
    const float* hit_positions;        /* rgba32f: w = hit distance, -1 = miss / no ray */
    const float* depth;                /* r32f, context extent */
    const float* normal_roughness;     /* rgba32f, context extent */
    const float* velocity;             /* rg32f, context extent, or NULL */
    const uint8_t* history_validation; /* r8_uint, or NULL */
} bpt_reblur_inputs;
BPT_API bpt_status bpt_denoise_reblur(bpt_context* ctx, const bpt_camera* camera, uint64_t frame_count, const bpt_reblur_settings* settings,
                                      const bpt_reblur_inputs* inputs, float* out_rgba32f);
BPT_API bpt_status bpt_reblur_reset(bpt_context* ctx);      /* forgets the history (a new camera, reblur.cpp:283) */
/* Debug: 0 = lighting_dist_0 with its 4 mip levels (level 0 = the blurred image), 1 = lighting_dist_1 (the stabilised image),
 * 2 = accumulation, 3 = linear depth with its 4 mip levels; as float32, after the last bpt_denoise_reblur. */
BPT_API bpt_status bpt_debug_read_reblur(bpt_context* ctx, uint32_t which, float* out, uint64_t capacity_floats);

/* ---------------------------------------------------------------------------------------
 * The step after the path: PostProcessPass::render, called with the path tracer's colour
 * (bisemutum/src/renderer/basic.cpp:228-231; bisemutum/src/renderer/pass/post_process.cpp:92-273) — bloom
 * (bloom_pre.hlsl, 3 iterations of bloom_filter.hlsl at W>>1, W>>2, W>>3, bloom_combine.hlsl chain) and the output pass
 * (post_process.hlsl: (colour.xyz, 1)). Settings = the bloom fields of PostProcessVolume
 * (include/bisemutum/renderer/post_process_volume.hpp:13-15). Input = the resolved accumulation image
 * (sum * 1/total_samples, as bpt_resolve); every intermediate target is rgba16_sfloat as in the reference
 * (round-to-nearest-even half stores); sampling contract in oracle/oracle_post.cpp. bloom == 0: out = (colour.xyz, 1).
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_post_settings {
    uint32_t bloom;                    /* PostProcessVolume::bloom, default 0 */
    float bloom_threshold;             /* default 1.5 */
    float bloom_threshold_softness;    /* default 0.5, in [0, 1] */
    uint32_t _pad;
} bpt_post_settings;
/* Host destination (W*H rgba32f; synchronises). */
BPT_API bpt_status bpt_post_process(bpt_context* ctx, const bpt_post_settings* settings, uint32_t total_samples, float* out_rgba32f);
/* Device destination (no synchronisation). */
BPT_API bpt_status bpt_post_process_device(bpt_context* ctx, const bpt_post_settings* settings, uint32_t total_samples, float* out_rgba32f_device);

typedef struct bpt_counters {
    uint64_t extend_rays;        /* rays actually traversed by the extend kernel          */
    uint64_t shadow_rays;        /* rays actually traversed by the connect kernel         */
    uint64_t samples;            /* pixel-samples started                                  */
    uint64_t kernel_launches;    /* kernels launched by this library since last reset      */
    uint64_t extend_rays_per_bounce[16];
    uint64_t shadow_rays_per_bounce[16];
} bpt_counters;
BPT_API bpt_status bpt_get_counters(bpt_context* ctx, bpt_counters* out);   /* synchronises */
BPT_API bpt_status bpt_reset_counters(bpt_context* ctx);

/* Per-kernel device timing (CUDA events on the launching stream around every kernel of bpt_render).
 * Used by bench.py for the roofline of the dominant kernel; off by default (no events recorded). */
typedef struct bpt_kernel_times {
    double raygen_ms, extend_ms, shade_ms, connect_ms, other_ms;
    uint64_t raygen_launches, extend_launches, shade_launches, connect_launches, other_launches;
} bpt_kernel_times;
BPT_API bpt_status bpt_profile_enable(bpt_context* ctx, uint32_t enable);
BPT_API bpt_status bpt_profile_read(bpt_context* ctx, bpt_kernel_times* out); /* synchronises, then resets */

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: samples shard by frame_index, scene + BVH are replicated, and ONE reduce of the
 * W x H x 4 FP32 sum buffers per batch of frames is the only exchange. One process (and one bpt_context) per GPU.
 * NCCL is loaded at run time (dlopen "libnccl.so.2"); nothing here is needed for single-GPU use.
 *   rank 0: bpt_comm_unique_id -> send the 128 bytes to every rank by any means (MPI, torch.distributed, a file);
 *   every rank: bpt_comm_init(ctx, id, rank, world), or bpt_comm_attach with an ncclComm_t the host already owns;
 *   per batch: bpt_render(...) with this rank's frames, then bpt_reduce(ctx, root): in-place ncclReduce(sum) of the
 *   accumulation buffer to `root`, stream-ordered on the context's stream; root then calls bpt_resolve(total samples).
 * state_precision = reference_fp16 cannot be reduced (a running lerp is order-dependent): BPT_ERR_STATE.
 * ------------------------------------------------------------------------------------- */
#define BPT_COMM_UNIQUE_ID_BYTES 128
BPT_API bpt_status bpt_comm_unique_id(uint8_t out_id[BPT_COMM_UNIQUE_ID_BYTES]);
BPT_API bpt_status bpt_comm_init(bpt_context* ctx, const uint8_t id[BPT_COMM_UNIQUE_ID_BYTES], int rank, int world_size);
BPT_API bpt_status bpt_comm_attach(bpt_context* ctx, void* nccl_comm /* ncclComm_t, stays owned by the caller */);
BPT_API bpt_status bpt_comm_destroy(bpt_context* ctx);
BPT_API bpt_status bpt_reduce(bpt_context* ctx, int root);

/* ---------------------------------------------------------------------------------------
 * Debug / parity hooks
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_ray {       /* 32 B */
    float origin[3]; float tmin;
    float direction[3]; float tmax;
} bpt_ray;
typedef struct bpt_hit {       /* 16 B + ids */
    float t;                   /* < 0 → miss */
    float u, v;                /* barycentrics of vertex 1 and 2 */
    uint32_t instance;         /* instance_id (continuous drawable index) */
    uint32_t primitive;        /* triangle index inside its BLAS          */
} bpt_hit;
/* Closest-hit trace of an arbitrary host ray batch through the extend kernel
 * (rt_gbuffer.hlsl:17-25 semantics; opacity rule of hits/rt_gbuffer_hit.hlsl:20-35 with
 * `frame_index` as the opacity seed). */
BPT_API bpt_status bpt_trace_rays(bpt_context* ctx, const bpt_ray* rays, uint64_t num_rays, uint32_t frame_index, bpt_hit* out_hits);
/* Any-hit (occlusion) trace through the connect kernel; out_visible[i] = 1 if unoccluded. */
BPT_API bpt_status bpt_trace_shadow_rays(bpt_context* ctx, const bpt_ray* rays, uint64_t num_rays, uint32_t frame_index, uint8_t* out_visible);

/* Merged mode: the 4-wide quantised tree the traversal kernels walk (csrc/bpt_wide.cuh; derived from the binary LBVH by a
 * deterministic collapse) and the exact leaf boxes: wide_nodes = (n-1) x 16 floats (64-B nodes, raw bits), leaf_boxes =
 * n x 8 floats. Either may be NULL. For bit-exact structure checks against the oracle. */
BPT_API bpt_status bpt_debug_read_wide(bpt_context* ctx, float* wide_nodes, float* leaf_boxes, uint32_t capacity_leaves);

/* When enabled, bpt_render(num_samples == 1) keeps, per bounce, the extend queue (pixel index
 * per live path), the hit records and the shadow-ray queue (pixel, light) for read-back. */
BPT_API bpt_status bpt_debug_capture(bpt_context* ctx, uint32_t enable);
/* kind 0: extend queue → pixels[i], hits[i] (hits may be NULL).
 * kind 1: shadow queue → pixels[i], lights[i] (index: dir lights first, then point lights). */
BPT_API bpt_status bpt_debug_read_queue(
    bpt_context* ctx, uint32_t bounce /* 1..max_bounces-1 */, uint32_t kind,
    uint32_t* pixels, uint32_t* lights, bpt_hit* hits, uint64_t capacity, uint64_t* count);

/* ---------------------------------------------------------------------------------------
 * Primary-hit outputs: PathTracingPass::render returns OutputData{color, depth, velocity = invalid, gbuffer}
 * (path_tracing.cpp:482-487; consumed by basic.cpp:162-165 and the post-process pass). bpt_render_primary produces
 * `depth` (pt_depth.hlsl:7-16: reverse-Z device depth of the primary hit, 0 = background) and the primary G-buffer as
 * the trace pass packs it (hits/rt_gbuffer_hit.hlsl:6-18, gbuffer.hlsl:18-33) — every channel holds the value a later
 * pass LOADS from the reference's texture format (gbuffer.hpp:14-17: rgba16_sfloat, rgba16_sfloat, rgba16_unorm,
 * rgba8_unorm). Texels of rays that miss are zero. Synchronous; either output may be NULL.
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_gbuffer_texel {   /* 64 B */
    float base_color[4];
    float normal_roughness[4];
    float fresnel[4];
    float material_0[4];
} bpt_gbuffer_texel;
BPT_API bpt_status bpt_render_primary(
    bpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const bpt_settings* settings,
    float* out_depth /* W*H */, bpt_gbuffer_texel* out_gbuffer /* W*H */);

/* Ray-traced ambient occlusion: AmbientOcclusionPass::render_raytraced (ambient_occlusion.cpp:217-262)
 * = ambient_occlusion_rt.hlsl:14-66 through the connect kernel: per pixel 4 cosine-hemisphere any-hit rays from the
 * position reconstructed from `depth`, RAY_FLAG_CULL_NON_OPAQUE | ACCEPT_FIRST_HIT, t in [0.001, max(range, 0.05)].
 * Inputs are full-resolution W x H images in the layout bpt_render_primary writes (depth; normal_roughness = float4 per
 * pixel). out_ao: (W or W/2) x (H or H/2) float2 = (ao, valid) as stored in the reference's rg16_sfloat target.
 * half_resolution follows the shader's per-frame sub-pixel choice and needs even W and H. Synchronous. */
typedef struct bpt_ao_settings {     /* BasicRenderer::AmbientOcclusionSettings (renderer/basic.hpp:40-50) */
    float range;                     /* 0.5 */
    float strength;                  /* 0.5 */
    uint32_t half_resolution;        /* reference default: 1 */
} bpt_ao_settings;
BPT_API bpt_status bpt_trace_ao(
    bpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const bpt_ao_settings* settings,
    const float* depth /* W*H */, const float* normal_roughness /* W*H*4 */, float* out_ao /* aw*ah*2 */);

/* Image-based lighting of the sky: SkyboxPrecomputePass::render (src/renderer/pass/skybox_precompute.cpp:66-162) =
 *   "IBL BRDF LUT"                 shaders/renderer/skybox/ibl_brdf_lut.hlsl:8-32               rg8_unorm, brdf_lut_size^2 (reference 128)
 *   "Skybox Precompute Diffuse"    shaders/renderer/skybox/skybox_precompute_diffuse.hlsl:12-41  rgba16_sfloat cube, diffuse_size (256)
 *   "Skybox Precompute Specular i" shaders/renderer/skybox/skybox_precompute_specular.hlsl:12-41 rgba16_sfloat cube, specular_size (256),
 *                                                                                                specular_levels mips (5), roughness = i / (levels - 1)
 * computed from the faces given to bpt_scene_upload_sky (texture sizes / formats: src/renderer/context/skybox.cpp:11-29). The
 * results stay on the device and feed the IBL block of the secondary lighting shader (bpt_trace_reflection with settings.ibl);
 * diffuse_strength / specular_strength are Skybox::diffuse_strength / specular_strength (scene_basic/skybox.hpp:19-20), multiplied
 * with the sky colour as SkyboxContext::update_shader_params does (skybox.cpp:40-43). A later bpt_scene_upload_sky invalidates them.
 * bpt_debug_read_sky_ibl copies them out (any pointer may be NULL): diffuse 6*d*d*4 floats, specular sum over levels of 6*s_i*s_i*4
 * floats (level 0 first), brdf lut n*n*2 floats. */
typedef struct bpt_sky_ibl_desc {
    uint32_t diffuse_size, specular_size, specular_levels, brdf_lut_size;
    float diffuse_strength, specular_strength;
} bpt_sky_ibl_desc;
BPT_API bpt_status bpt_precompute_sky_ibl(bpt_context* ctx, const bpt_sky_ibl_desc* desc);
BPT_API bpt_status bpt_debug_read_sky_ibl(bpt_context* ctx, float* diffuse_rgba32f, float* specular_rgba32f, float* brdf_lut_rg32f);

/* Ray-traced reflections: ReflectionPass::render_raytraced (src/renderer/pass/reflection.cpp:317-450) =
 *   "RTR Sample Direction"  shaders/renderer/raytracing/direction_sample/specular_sample.hlsl:14-83 (VNDF sample of the specular
 *                           lobe from the camera's depth + G-buffer; pixels rougher than max_roughness are skipped, weights fade
 *                           between fade_roughness and max_roughness),
 *   "RTR Trace GBuffer"     rt_gbuffer.hlsl:7-36 with ray_length = range (the extend kernel),
 *   "RTR Lighting"          deferred_lighting_secondary.hlsl:11-111 with lighting_strength = strength (the shade + connect kernels;
 *                           light visibility by shadow ray as in the path tracer; with settings.ibl the IBL block :98-108 is
 *                           evaluated from the textures of bpt_precompute_sky_ibl, otherwise the pass behaves as with
 *                           DEFERRED_LIGHTING_NO_IBL).
 * Inputs: full-resolution depth and G-buffer as bpt_render_primary writes them. Outputs: (W or W/2) x (H or H/2) rgba32f
 * reflection colour (rgb, 1) and hit positions (hit: (P, t); miss: (ray direction, -1); pixel without a ray: (0, 0, 0, -1)).
 * half_resolution follows the shader's per-frame sub-pixel and needs even W and H. The upscale and denoise passes that follow
 * in the reference (reflection.cpp:452-571, ReBLUR) are not part of this call. Synchronous. */
typedef struct bpt_reflection_settings {     /* BasicRenderer::ReflectionSettings (renderer/basic.hpp:52-65), mode = raytraced */
    float range;                     /* 16 */
    float strength;                  /* 1 */
    float max_roughness;             /* 0.3 */
    float fade_roughness;            /* 0.1 */
    uint32_t half_resolution;        /* reference default: 1 */
    uint32_t ibl;                    /* 1: add the IBL block of the lighting shader (needs bpt_precompute_sky_ibl); 0: lights + sky only */
} bpt_reflection_settings;
BPT_API bpt_status bpt_trace_reflection(
    bpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const bpt_reflection_settings* settings,
    const float* depth /* W*H */, const bpt_gbuffer_texel* gbuffer /* W*H */,
    float* out_reflection /* rw*rh*4 */, float* out_hit_positions /* rw*rh*4 */);
/* "RTR Upscale Hit" / "RTR Upscale Color" (ReflectionPass::render_upscale, reflection.cpp:452-530; shaders/renderer/simple_upscale.hlsl):
 * a half-resolution image ((W+1)/2 x (H+1)/2 float4, traced with the frame's sub-pixel) to full resolution with a 3 x 3 half-res
 * neighbourhood weighted by a Gaussian of the sub-pixel distance, the normal agreement and the linear-depth difference. `depth`
 * and `normal_roughness` are the full-resolution images of bpt_render_primary; out-of-range taps read 0 like Texture.Load.
 * exp() is the fixed-order form of the numeric contract. Synchronous. */
BPT_API bpt_status bpt_upscale_half_res(
    bpt_context* ctx, const bpt_camera* camera, uint32_t frame_index, const float* depth /* W*H */, const float* normal_roughness /* W*H*4 */,
    const float* in_half_res /* rw*rh*4 */, float* out_full_res /* W*H*4 */);

/* ---------------------------------------------------------------------------------------
 * DDGI-style probe tracing through the same extend/shade kernels
 * (shaders/renderer/ddgi/trace_gbuffer.hlsl:10-51, ddgi/deferred_lighting.hlsl:12-118).
 * ------------------------------------------------------------------------------------- */
typedef struct bpt_probe_volume {
    float base_position[3]; float _pad0;
    float frame_x[3]; float _pad1;
    float frame_y[3]; float _pad2;
    float frame_z[3]; float _pad3;
    float extent[3]; float ray_length;
    uint32_t probe_counts[3];
    uint32_t rays_per_probe;
} bpt_probe_volume;
/* out_radiance_dist: num_probes*rays_per_probe float4 = (radiance rgb, hit distance or -1). */
BPT_API bpt_status bpt_trace_probes(
    bpt_context* ctx, const bpt_probe_volume* volume, const float* sample_table_r2 /* 8192 float2 */,
    uint32_t frame_index, uint32_t num_bounces, float* out_radiance_dist);
/* The same for the probes [first_probe, first_probe + num_probes) only (linear probe index = (iz * ny + iy) * nx + ix):
 * out_radiance_dist holds num_probes*rays_per_probe float4. Every random choice is keyed by the probe's GLOBAL index
 * (ddgi/trace_gbuffer.hlsl:25-29), so the rays of a range are bit-identical to the same rays of a full bpt_trace_probes — the unit
 * of the multi-GPU sharding of the DDGI update (shard by probe index, all-gather the per-ray results, sharding.py).
 * out_radiance_dist of bpt_trace_probes[_range] and ray_radiance_dist of bpt_blend_probes may be HOST or DEVICE pointers (unified
 * addressing): a sharded update keeps the per-ray results on the device between the trace, the NCCL all-gather and the blend. */
BPT_API bpt_status bpt_trace_probes_range(
    bpt_context* ctx, const bpt_probe_volume* volume, const float* sample_table_r2 /* 8192 float2 */,
    uint32_t frame_index, uint32_t num_bounces, uint32_t first_probe, uint32_t num_probes, float* out_radiance_dist);

/* DDGI probe blending (ddgi/probe_blend_irradiance.hlsl, probe_blend_visibility.hlsl):
 * gathers the per-ray output of bpt_trace_probes into octahedral atlases with a 1-texel border,
 *   irradiance: rgba32f, width = nx*ny*(irradiance_size+2), height = nz*(irradiance_size+2)   (reference: 6)
 *   visibility: rg32f,   width = nx*ny*(visibility_size+2), height = nz*(visibility_size+2)   (reference: 14)
 * probe (ix,iy,iz) starts at ((iy*nx + ix), iz) * (size+2). When history_valid != 0 the atlases are read as the
 * previous frame and blended with alpha (reference 0.97) in gamma-5 space; they are overwritten with the result. */
typedef struct bpt_probe_blend {
    uint32_t irradiance_size, visibility_size;
    float alpha;
    uint32_t history_valid;
} bpt_probe_blend;
BPT_API bpt_status bpt_blend_probes(
    bpt_context* ctx, const bpt_probe_volume* volume, const float* sample_table_r2, uint32_t frame_index,
    const float* ray_radiance_dist /* output of bpt_trace_probes */, const bpt_probe_blend* blend,
    float* irradiance_atlas_rgba32f, float* visibility_atlas_rg32f);

/* The consumer of the atlases: calc_ddgi_volume_lighting (ddgi/ddgi_lighting.hlsl:7-83) — 8-probe trilinear weights x
 * wrap-shading weight x Chebyshev visibility^3, generalised from the reference's 8x8x8 probes to the volume's own counts.
 * bpt_set_ddgi_volume binds one volume and its atlases (layout of bpt_blend_probes; copied to the device; all-NULL
 * unbinds). While bound, bpt_trace_probes adds the previous update's irradiance at the last vertex of every probe path
 * (ddgi/deferred_lighting.hlsl:102-115), which is how the reference's one-bounce probes converge to multi-bounce GI. */
BPT_API bpt_status bpt_set_ddgi_volume(
    bpt_context* ctx, const bpt_probe_volume* volume, const bpt_probe_blend* sizes,
    const float* irradiance_atlas_rgba32f, const float* visibility_atlas_rg32f);
/* out[i] = calc_ddgi_volume_lighting(position[i], normal[i], view[i]) for the bound volume: (irradiance rgb, 1) or 0
 * outside the volume. position/normal/view: n x float3. Synchronous. */
BPT_API bpt_status bpt_ddgi_lighting(
    bpt_context* ctx, uint64_t n, const float* position, const float* normal, const float* view, float* out_rgba);

#ifdef __cplusplus
}
#endif
#endif /* BPT_H_ */
