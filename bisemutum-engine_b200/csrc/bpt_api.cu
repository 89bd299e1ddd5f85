// bpt_api.cu — the C ABI of include/bpt/bpt.h: context, scene upload, accel build, render, debug.
// Each entry point cites the reference interface it replaces in the header.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include "bpt_internal.cuh"

using namespace bptd;

bpt_status dev_alloc(bpt_context* ctx, DevBuf& b, size_t bytes) {
    b.p = nullptr; b.bytes = 0;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        ctx->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        b.p = nullptr;
        (void)cudaGetLastError();      // clear the sticky last-error so a later launch check does not report it
        return e == cudaErrorMemoryAllocation ? BPT_ERR_OOM : BPT_ERR_CUDA;
    }
    b.bytes = bytes;
    return BPT_OK;
}
void dev_free(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.bytes = 0;
}
bpt_status dev_reserve(bpt_context* ctx, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.p && b.bytes >= bytes && b.bytes <= 2 * bytes + (1u << 20)) return BPT_OK;
    dev_free(b);
    return dev_alloc(ctx, b, bytes);
}
bpt_status dev_upload(bpt_context* ctx, DevBuf& b, const void* src, size_t bytes) {
    bpt_status s = dev_reserve(ctx, b, bytes);
    if (s) return s;
    if (bytes && src) BPT_CUDA_TRY(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyDefault, ctx->stream));    // (unified addressing: `src` may be a device pointer)
    return BPT_OK;
}

DScene bpt_context::scene_view() const {
    DScene s{};
    s.positions = d_positions.as<float>(); s.normals = d_normals.as<float>(); s.tangents = d_tangents.as<float>();
    s.texcoords = d_texcoords.as<float>(); s.colors = d_colors.as<float>(); s.indices = d_indices.as<uint32_t>();
    s.drawables = d_drawables.as<bpt_drawable_sbt_data>(); s.drawable_va = d_drawable_va.as<uint32_t>();
    s.materials = d_materials.as<bpt_material>();
    s.textures = d_textures.as<DTexture>(); s.num_textures = (uint32_t)d_texels.size();
    s.instances = d_instances.as<DInstance>(); s.num_instances = (uint32_t)h_instances.size();
    s.accel_mode = accel_mode;
    s.tlas_nodes = tlas.nodes.as<float4>(); s.tlas_prims = tlas.prims.as<uint32_t>(); s.tlas_root = tlas.root; s.tlas_n = tlas.n;
    s.tlas_wide = tlas.wide.as<float4>(); s.tlas_leafbox = tlas.leafbox.as<float4>();
    s.ibl_enabled = ibl_valid ? 1u : 0u; s.ibl_diffuse_size = ibl_desc.diffuse_size; s.ibl_specular_size = ibl_desc.specular_size;
    s.ibl_specular_levels = ibl_desc.specular_levels; s.ibl_brdf_size = ibl_desc.brdf_lut_size;
    s.ibl_diffuse = d_ibl_diffuse.as<float4>(); s.ibl_specular = d_ibl_specular.as<float4>(); s.ibl_brdf = d_ibl_brdf.as<float2>();
    for (int k = 0; k < 3; k++) { s.ibl_diffuse_color[k] = sky_color[k] * ibl_desc.diffuse_strength; s.ibl_specular_color[k] = sky_color[k] * ibl_desc.specular_strength; }
    s.blas = d_blas_table.as<DBlas>();
    s.dir_lights = d_dir.as<bpt_dir_light_data>(); s.num_dir = num_dir;
    s.point_lights = d_point.as<bpt_point_light_data>(); s.num_point = num_point;
    s.rect_lights = d_rect.as<bpt_rect_light_data>(); s.num_rect = num_rect;
    s.light_textures = d_light_textures.as<DLightTexture>(); s.num_light_textures = (uint32_t)h_light_textures.size();
    s.ltc_m0 = d_ltc[0].as<float>(); s.ltc_m1 = d_ltc[1].as<float>(); s.ltc_m2 = d_ltc[2].as<float>(); s.ltc_norm = d_ltc[3].as<float>();
    s.sky_faces = d_sky.as<float4>(); s.sky_size = sky_size;
    memcpy(s.sky_transform, sky_transform, sizeof(sky_transform));
    memcpy(s.sky_color, sky_color, sizeof(sky_color));
    s.ddgi_enabled = ddgi_enabled ? 1u : 0u; s.ddgi_irr_size = ddgi_irr_size; s.ddgi_vis_size = ddgi_vis_size;
    s.ddgi_irradiance = d_ddgi_irr.as<float4>(); s.ddgi_visibility = d_ddgi_vis.as<float2>(); s.ddgi_volume = ddgi_volume;
    return s;
}

static bpt_status fail(bpt_context* c, bpt_status s, const char* msg) { c->err = msg; return s; }
// Samples traced ahead (bpt_render_ahead) were shaded with the scene as it was: every entry point that changes what a sample would
// see drops them, so bpt_pending_ahead reports 0 and the pass traces the frame again with the current scene.
static void invalidate_ahead(bpt_context* c) { c->wf.ahead_slots = c->wf.ahead_cursor = 0; c->scene_generation++; }
#define NEED(c) do { if (!(c)) return BPT_ERR_INVALID; cudaSetDevice((c)->device); } while (0)

// ---- NCCL, loaded at run time so that libbpt.so has no link-time dependency on it (single-GPU hosts need none) -------------
#include <dlfcn.h>
namespace {
struct NcclId { char internal[128]; };        // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128), passed BY VALUE to ncclCommInitRank
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi& nccl() {
    static NcclApi api = [] {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!a.lib) return a;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
        a.Reduce = reinterpret_cast<decltype(a.Reduce)>(dlsym(a.lib, "ncclReduce"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Reduce && a.GetErrorString;
        return a;
    }();
    return api;
}
constexpr int kNcclFloat32 = 7, kNcclSum = 0;      // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)
} // namespace

extern "C" {

const char* bpt_version(void) { return "bpt 0.1 (sm_100a)"; }

bpt_status bpt_create(const bpt_config* cfg, bpt_context** out) {
    if (!cfg || !out || cfg->width == 0 || cfg->height == 0) return BPT_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return BPT_ERR_NO_DEVICE;   // no CPU path exists
    if (cfg->device < 0 || cfg->device >= count) return BPT_ERR_INVALID;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return BPT_ERR_CUDA;
    bpt_context* c = new (std::nothrow) bpt_context();
    if (!c) return BPT_ERR_OOM;
    c->device = cfg->device; c->width = cfg->width; c->height = cfg->height;
    bpt_status s = wavefront_alloc(c);
    if (s) { delete c; return s; }
    *out = c;
    return BPT_OK;
}

bpt_status bpt_destroy(bpt_context* c) {
    if (!c) return BPT_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->nccl_comm && c->nccl_owned) nccl().CommDestroy(c->nccl_comm);
    DevBuf* bufs[] = {&c->d_colors, &c->d_positions, &c->d_normals, &c->d_tangents, &c->d_texcoords, &c->d_indices, &c->d_drawables, &c->d_drawable_va,
                      &c->d_materials, &c->d_textures, &c->d_instances, &c->d_dir, &c->d_point, &c->d_rect, &c->d_ltc[0], &c->d_ltc[1],
                      &c->d_ltc[2], &c->d_ltc[3], &c->d_light_textures, &c->d_srgb_tables, &c->d_sky, &c->d_ddgi_irr, &c->d_ddgi_vis, &c->d_blas_table, &c->d_inst_aabb, &c->d_blas_bounds, &c->d_post, &c->d_post_out, &c->d_ibl_diffuse, &c->d_ibl_specular, &c->d_ibl_brdf, &c->wf.hit, &c->wf.hit_slot, &c->wf.sh_o,
                      &c->wf.sh_d, &c->wf.sh_c, &c->wf.accum, &c->wf.color, &c->wf.bcol, &c->wf.qcount, &c->wf.totals, &c->tlas.nodes, &c->tlas.tris, &c->tlas.morton, &c->tlas.prims,
                      &c->tlas.wide, &c->tlas.leafbox};
    for (DevBuf* b : bufs) dev_free(*b);
    for (int k = 0; k < 2; k++) { dev_free(c->wf.ray_o[k]); dev_free(c->wf.ray_d[k]); dev_free(c->wf.ray_w[k]); }
    for (auto& t : c->d_texels) dev_free(t);
    for (auto& t : c->d_light_texels) dev_free(t);
    {
        ReblurState& r = c->reblur;
        DevBuf* rb[] = {&r.ld0, &r.ld1, &r.accum, &r.lin_depth, &r.denoised, &r.hist_ld0, &r.hist_ld1, &r.hist_accum, &r.depth[0], &r.depth[1], &r.nr[0], &r.nr[1],
                        &r.velocity, &r.validation, &r.noised, &r.hit};
        for (DevBuf* b : rb) dev_free(*b);
    }
    for (auto& b : c->blas) { dev_free(b.nodes); dev_free(b.tris); dev_free(b.morton); dev_free(b.prims); dev_free(b.wide); dev_free(b.leafbox); }
    for (auto& a : c->arena_chunks) dev_free(a);
    delete c;
    return BPT_OK;
}

const char* bpt_last_error(const bpt_context* c) { return c ? c->err.c_str() : "null context"; }

bpt_status bpt_set_stream(bpt_context* c, void* stream) {
    NEED(c);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));      // build scratch and uploads are ordered on the stream they were issued on
    c->stream = (cudaStream_t)stream;
    return BPT_OK;
}
bpt_status bpt_sync(bpt_context* c) { NEED(c); BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream)); return BPT_OK; }

bpt_status bpt_resize(bpt_context* c, uint32_t w, uint32_t h) {
    NEED(c);
    if (!w || !h) return BPT_ERR_INVALID;
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->width = w; c->height = h;
    invalidate_ahead(c);
    return wavefront_alloc(c);
}

bpt_status bpt_set_wave_budget(bpt_context* c, uint64_t max_paths) {
    NEED(c);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->wave_paths_budget = max_paths;
    invalidate_ahead(c);
    return wavefront_alloc(c);
}

bpt_status bpt_scene_upload_geometry(bpt_context* c, const bpt_geometry_streams* g, const bpt_drawable_sbt_data* dr, const uint32_t* va,
                                     uint32_t nd, const bpt_blas_desc* blas, uint32_t nb) {
    NEED(c);
    if (!g || !dr || !blas || !nd || !nb || !g->positions || !g->indices) return fail(c, BPT_ERR_INVALID, "geometry: null stream or empty scene");
    invalidate_ahead(c);
    bpt_status s;
    if ((s = dev_upload(c, c->d_positions, g->positions, g->num_position_floats * 4))) return s;
    if ((s = dev_upload(c, c->d_normals, g->normals, g->normals ? g->num_normal_floats * 4 : 0))) return s;
    if ((s = dev_upload(c, c->d_tangents, g->tangents, g->tangents ? g->num_tangent_floats * 4 : 0))) return s;
    if ((s = dev_upload(c, c->d_texcoords, g->texcoords, g->texcoords ? g->num_texcoord_floats * 4 : 0))) return s;
    if ((s = dev_upload(c, c->d_colors, g->colors, g->colors ? g->num_color_floats * 4 : 0))) return s;
    if ((s = dev_upload(c, c->d_indices, g->indices, g->num_indices * 4))) return s;
    c->has_normals = g->normals != nullptr; c->has_tangents = g->tangents != nullptr; c->has_texcoords = g->texcoords != nullptr;
    c->has_colors = g->colors != nullptr;
    c->num_position_floats = g->num_position_floats; c->num_indices = g->num_indices;
    c->h_drawables.assign(dr, dr + nd);
    std::vector<uint32_t> vam(nd);
    for (uint32_t i = 0; i < nd; i++) {
        uint32_t m = va ? va[i] : (BPT_VA_POSITION | BPT_VA_NORMAL | BPT_VA_TANGENT | BPT_VA_TEXCOORD | (g->colors ? (uint32_t)BPT_VA_COLOR : 0u));
        if (!c->has_colors) m &= ~BPT_VA_COLOR;
        if (!c->has_normals) m &= ~BPT_VA_NORMAL;
        if (!c->has_tangents) m &= ~BPT_VA_TANGENT;
        if (!c->has_texcoords) m &= ~BPT_VA_TEXCOORD;
        vam[i] = m;
    }
    if ((s = dev_upload(c, c->d_drawables, dr, (size_t)nd * sizeof(bpt_drawable_sbt_data)))) return s;
    if ((s = dev_upload(c, c->d_drawable_va, vam.data(), (size_t)nd * 4))) return s;
    c->h_blas_desc.assign(blas, blas + nb);
    for (auto& bd : c->h_blas_desc)
        if (bd.num_triangles == 0 || (uint64_t)bd.index_offset + 3ull * bd.num_triangles > g->num_indices)
            return fail(c, BPT_ERR_INVALID, "BLAS index range out of bounds or empty");
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // host buffers may be released by the caller
    c->accel_built = false;
    return BPT_OK;
}

bpt_status bpt_scene_upload_instances(bpt_context* c, const bpt_instance_desc* inst, uint32_t n) {
    NEED(c);
    if (!inst || !n) return fail(c, BPT_ERR_INVALID, "instances: empty");
    c->h_instances.assign(inst, inst + n);       // (takes effect with the next bpt_build_accel / bpt_update_tlas, which drop prefetched samples)
    return BPT_OK;
}

bpt_status bpt_scene_upload_materials(bpt_context* c, const bpt_material* m, uint32_t n, const bpt_texture_desc* t, uint32_t nt) {
    NEED(c);
    if (!m || !n) return fail(c, BPT_ERR_INVALID, "materials: empty");
    invalidate_ahead(c);
    c->h_materials.assign(m, m + n);
    bpt_status s;
    if ((s = dev_upload(c, c->d_materials, m, (size_t)n * sizeof(bpt_material)))) return s;
    for (auto& b : c->d_texels) dev_free(b);
    c->d_texels.assign(nt, DevBuf{});
    std::vector<DTexture> table(std::max(nt, 1u));
    for (uint32_t i = 0; i < nt; i++) {
        if (!t[i].texels || !t[i].width || !t[i].height || t[i].format > BPT_TEXTURE_RGBA8_SRGB) return fail(c, BPT_ERR_INVALID, "bad texture desc");
        uint32_t fmt = t[i].format;
        if (fmt == BPT_TEXTURE_RGBA8_SRGB) {
            // decode to linear FP32 texels once (the sampler decodes before filtering): 256-entry sRGB EOTF table
            float lut[256];
            for (int k = 0; k < 256; k++) { double v = k / 255.0; lut[k] = (float)(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4)); }
            const uint8_t* src = static_cast<const uint8_t*>(t[i].texels);
            std::vector<float> lin((size_t)t[i].width * t[i].height * 4);
            for (size_t k = 0; k < lin.size(); k += 4) { lin[k] = lut[src[k]]; lin[k + 1] = lut[src[k + 1]]; lin[k + 2] = lut[src[k + 2]]; lin[k + 3] = (float)src[k + 3] / 255.0f; }
            if ((s = dev_upload(c, c->d_texels[i], lin.data(), lin.size() * 4))) return s;
            BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            fmt = BPT_TEXTURE_RGBA32_FLOAT;
        } else {
            size_t bytes = (size_t)t[i].width * t[i].height * (fmt == BPT_TEXTURE_RGBA8_UNORM ? 4 : 16);
            if ((s = dev_upload(c, c->d_texels[i], t[i].texels, bytes))) return s;
        }
        table[i] = DTexture{c->d_texels[i].p, t[i].width, t[i].height, fmt, t[i].address_mode_u, t[i].address_mode_v, t[i].filter_linear};
    }
    if ((s = dev_upload(c, c->d_textures, table.data(), table.size() * sizeof(DTexture)))) return s;
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

bpt_status bpt_scene_upload_lights(bpt_context* c, const bpt_dir_light_data* d, uint32_t nd, const bpt_point_light_data* p, uint32_t np,
                                   const bpt_rect_light_data* r, uint32_t nr, const bpt_ltc_luts* luts) {
    NEED(c);
    if ((nd && !d) || (np && !p) || (nr && !r)) return fail(c, BPT_ERR_INVALID, "lights: null array");
    bpt_status s;
    // prefetched samples stay valid only if the lights are byte-for-byte what they were shaded with (the per-frame refresh of an
    // unchanged LightsContext); the LTC tables are constants of the reference (a new set arrives with a count change only)
    auto same = [](std::vector<uint8_t>& keep, const void* src, size_t bytes) {
        bool eq = keep.size() == bytes && (bytes == 0 || memcmp(keep.data(), src, bytes) == 0);
        if (!eq) keep.assign(static_cast<const uint8_t*>(src), static_cast<const uint8_t*>(src) + bytes);
        return eq;
    };
    const bool eq_d = same(c->h_dir_bytes, d, (size_t)nd * sizeof(*d)), eq_p = same(c->h_point_bytes, p, (size_t)np * sizeof(*p)),
               eq_r = same(c->h_rect_bytes, r, (size_t)nr * sizeof(*r));
    if (!(eq_d && eq_p && eq_r)) invalidate_ahead(c);
    if (nd == c->num_dir && np == c->num_point && nr == c->num_rect && c->d_dir.p && (!nr || c->d_ltc[0].p)) {
        // per-frame refresh (PathTracingPass::update_params): same counts → overwrite in place, stream-ordered
        if (nd) BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_dir.p, d, (size_t)nd * sizeof(*d), cudaMemcpyHostToDevice, c->stream));
        if (np) BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_point.p, p, (size_t)np * sizeof(*p), cudaMemcpyHostToDevice, c->stream));
        if (nr) BPT_CUDA_TRY(c, cudaMemcpyAsync(c->d_rect.p, r, (size_t)nr * sizeof(*r), cudaMemcpyHostToDevice, c->stream));
        return BPT_OK;
    }
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if ((s = dev_upload(c, c->d_dir, d, (size_t)nd * sizeof(*d)))) return s;
    if ((s = dev_upload(c, c->d_point, p, (size_t)np * sizeof(*p)))) return s;
    if ((s = dev_upload(c, c->d_rect, r, (size_t)nr * sizeof(*r)))) return s;
    if (nr) {
        if (!luts || !luts->matrix_lut0 || !luts->matrix_lut1 || !luts->matrix_lut2 || !luts->norm_lut) return fail(c, BPT_ERR_INVALID, "rect lights need the LTC LUTs");
        const float* src[4] = {luts->matrix_lut0, luts->matrix_lut1, luts->matrix_lut2, luts->norm_lut};
        for (int k = 0; k < 4; k++)
            if ((s = dev_upload(c, c->d_ltc[k], src[k], (size_t)8 * 8 * 64 * (k < 3 ? 4 : 2) * 4))) return s;
    }
    c->num_dir = nd; c->num_point = np; c->num_rect = nr;
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return wavefront_alloc(c);
}

bpt_status bpt_scene_upload_light_textures(bpt_context* c, const bpt_light_texture_desc* t, uint32_t nt) {
    NEED(c);
    if ((nt && !t) || nt > BPT_MAX_RECT_LIGHT_TEXTURES) return fail(c, BPT_ERR_INVALID, "light textures: null array or more than 16");
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    invalidate_ahead(c);
    return upload_light_textures(c, t, nt);
}
bpt_status bpt_debug_read_light_texture(bpt_context* c, uint32_t index, float* out, uint64_t cap, uint64_t* out_texels) {
    NEED(c);
    return read_light_texture(c, index, out, cap, out_texels);
}

bpt_status bpt_scene_upload_sky(bpt_context* c, const float* faces, uint32_t size, const float xf[9], const float col[3]) {
    NEED(c);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    invalidate_ahead(c);
    c->ibl_valid = false;                                   // derived from the faces: bpt_precompute_sky_ibl again
    if (faces && size) {
        bpt_status s = dev_upload(c, c->d_sky, faces, (size_t)6 * size * size * 16);
        if (s) return s;
        c->sky_size = size;
        BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    } else { dev_free(c->d_sky); c->sky_size = 0; }
    if (xf) memcpy(c->sky_transform, xf, sizeof(float) * 9);
    if (col) memcpy(c->sky_color, col, sizeof(float) * 3);
    return BPT_OK;
}

bpt_status bpt_scene_update_sky_params(bpt_context* c, const float xf[9], const float col[3]) {
    NEED(c);
    if ((xf && memcmp(c->sky_transform, xf, sizeof(float) * 9) != 0) || (col && memcmp(c->sky_color, col, sizeof(float) * 3) != 0)) invalidate_ahead(c);
    if (xf) memcpy(c->sky_transform, xf, sizeof(float) * 9);
    if (col) memcpy(c->sky_color, col, sizeof(float) * 3);
    return BPT_OK;
}

static bpt_status validate_scene(bpt_context* c) {
    if (c->h_blas_desc.empty() || c->h_instances.empty()) return fail(c, BPT_ERR_STATE, "build_accel: upload geometry and instances first");
    if (c->h_materials.empty()) return fail(c, BPT_ERR_STATE, "build_accel: upload materials first");
    for (auto& d : c->h_drawables)
        if (d.material_offset % sizeof(bpt_material) || d.material_offset / sizeof(bpt_material) >= c->h_materials.size())
            return fail(c, BPT_ERR_INVALID, "drawable material_offset out of range");
    for (auto& in : c->h_instances) {
        if (in.blas >= c->h_blas_desc.size()) return fail(c, BPT_ERR_INVALID, "instance references a BLAS out of range");
        if ((in.instance_id_and_mask & 0xffffffu) >= c->h_drawables.size()) return fail(c, BPT_ERR_INVALID, "instance_id out of drawable range");
    }
    return BPT_OK;
}

static bpt_status upload_blas_table(bpt_context* c) {
    std::vector<DBlas> t(c->blas.size());
    for (size_t i = 0; i < c->blas.size(); i++) t[i] = DBlas{c->blas[i].nodes.as<float4>(), c->blas[i].tris.as<float4>(), c->blas[i].root, c->blas[i].n,
                                                               c->blas[i].wide.as<float4>(), c->blas[i].leafbox.as<float4>()};
    bpt_status s = dev_upload(c, c->d_blas_table, t.data(), t.size() * sizeof(DBlas));
    if (s) return s;
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

bpt_status bpt_build_accel(bpt_context* c, uint32_t mode) {
    NEED(c);
    bpt_status s;
    if ((s = validate_scene(c))) return s;
    if (mode != BPT_ACCEL_TWO_LEVEL && mode != BPT_ACCEL_MERGED) return fail(c, BPT_ERR_INVALID, "unknown accel mode");
    invalidate_ahead(c);
    c->accel_built = false;
    c->accel_mode = mode;
    if ((s = upload_instance_table(c))) return s;
    // buffers of a previous build are kept when the BLAS count is unchanged (dev_reserve re-uses them)
    size_t want_blas = mode == BPT_ACCEL_TWO_LEVEL ? c->h_blas_desc.size() : 1;
    if (c->blas.size() != want_blas) {
        for (auto& b : c->blas) { dev_free(b.nodes); dev_free(b.tris); dev_free(b.morton); dev_free(b.prims); dev_free(b.wide); dev_free(b.leafbox); }
        c->blas.assign(want_blas, DevBvh{});
    }
    if ((s = dev_reserve(c, c->d_blas_bounds, want_blas * 6 * sizeof(float)))) return s;
    if (mode == BPT_ACCEL_TWO_LEVEL) {
        if ((s = build_all_blas_two_level(c))) return s;
        if ((s = build_tlas(c))) return s;
    } else {
        if ((s = build_blas_merged(c))) return s;
        dev_free(c->tlas.nodes); dev_free(c->tlas.morton); dev_free(c->tlas.prims); dev_free(c->tlas.wide); dev_free(c->tlas.leafbox);
        c->tlas = DevBvh{};
    }
    if ((s = upload_blas_table(c))) return s;
    c->accel_built = true;
    return BPT_OK;
}

bpt_status bpt_update_tlas(bpt_context* c) {
    NEED(c);
    if (!c->accel_built || c->accel_mode != BPT_ACCEL_TWO_LEVEL) return fail(c, BPT_ERR_STATE, "update_tlas needs a built two-level accel");
    bpt_status s;
    if ((s = validate_scene(c))) return s;
    invalidate_ahead(c);
    if ((s = upload_instance_table(c))) return s;
    return build_tlas(c);
}

bpt_status bpt_debug_read_bvh(bpt_context* c, uint32_t which, uint32_t* np, uint64_t* morton, uint32_t* prims, bpt_bvh_node* nodes, uint32_t cap, int32_t* root) {
    NEED(c);
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "accel not built");
    const DevBvh* b;
    if (which == BPT_BVH_TLAS) { if (c->accel_mode != BPT_ACCEL_TWO_LEVEL) return fail(c, BPT_ERR_INVALID, "no TLAS in merged mode"); b = &c->tlas; }
    else { if (which >= c->blas.size()) return fail(c, BPT_ERR_INVALID, "blas index out of range"); b = &c->blas[which]; }
    if (np) *np = b->n;
    if (root) *root = b->root;
    if ((morton || prims || nodes) && cap < b->n) return fail(c, BPT_ERR_INVALID, "capacity too small");
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (morton) BPT_CUDA_TRY(c, cudaMemcpy(morton, b->morton.p, (size_t)b->n * 8, cudaMemcpyDeviceToHost));
    if (prims) BPT_CUDA_TRY(c, cudaMemcpy(prims, b->prims.p, (size_t)b->n * 4, cudaMemcpyDeviceToHost));
    if (nodes && b->n >= 2) BPT_CUDA_TRY(c, cudaMemcpy(nodes, b->nodes.p, (size_t)(b->n - 1) * 64, cudaMemcpyDeviceToHost));
    return BPT_OK;
}

bpt_status bpt_comm_unique_id(uint8_t out_id[BPT_COMM_UNIQUE_ID_BYTES]) {
    if (!out_id) return BPT_ERR_INVALID;
    if (!nccl().ok) return BPT_ERR_UNSUPPORTED;
    NcclId id;
    if (nccl().GetUniqueId(&id) != 0) return BPT_ERR_CUDA;
    memcpy(out_id, id.internal, BPT_COMM_UNIQUE_ID_BYTES);
    return BPT_OK;
}
bpt_status bpt_comm_destroy(bpt_context* c) {
    NEED(c);
    if (c->nccl_comm && c->nccl_owned) { cudaStreamSynchronize(c->stream); nccl().CommDestroy(c->nccl_comm); }
    c->nccl_comm = nullptr; c->nccl_owned = false;
    return BPT_OK;
}
bpt_status bpt_comm_init(bpt_context* c, const uint8_t id[BPT_COMM_UNIQUE_ID_BYTES], int rank, int world) {
    NEED(c);
    if (!id || world < 1 || rank < 0 || rank >= world) return BPT_ERR_INVALID;
    if (!nccl().ok) return fail(c, BPT_ERR_UNSUPPORTED, "comm_init: libnccl.so.2 could not be loaded");
    bpt_comm_destroy(c);
    NcclId uid;
    memcpy(uid.internal, id, 128);
    void* comm = nullptr;
    int r = nccl().CommInitRank(&comm, world, uid, rank);
    if (r != 0) { c->err = std::string("ncclCommInitRank: ") + nccl().GetErrorString(r); return BPT_ERR_CUDA; }
    c->nccl_comm = comm; c->nccl_owned = true;
    return BPT_OK;
}
bpt_status bpt_comm_attach(bpt_context* c, void* comm) {
    NEED(c);
    if (!comm) return BPT_ERR_INVALID;
    if (!nccl().ok) return fail(c, BPT_ERR_UNSUPPORTED, "comm_attach: libnccl.so.2 could not be loaded");
    bpt_comm_destroy(c);
    c->nccl_comm = comm; c->nccl_owned = false;
    return BPT_OK;
}
bpt_status bpt_reduce(bpt_context* c, int root) {
    NEED(c);
    if (!c->nccl_comm) return fail(c, BPT_ERR_STATE, "reduce: no communicator (bpt_comm_init / bpt_comm_attach)");
    if (c->wf.accum_fp16) return fail(c, BPT_ERR_STATE, "reduce: the reference_fp16 running average cannot be summed across GPUs");
    const size_t count = (size_t)c->width * c->height * 4;
    int r = nccl().Reduce(c->wf.accum.p, c->wf.accum.p, count, kNcclFloat32, kNcclSum, root, c->nccl_comm, c->stream);
    if (r != 0) { c->err = std::string("ncclReduce: ") + nccl().GetErrorString(r); return BPT_ERR_CUDA; }
    return BPT_OK;
}

bpt_status bpt_debug_read_wide(bpt_context* c, float* wide, float* leafbox, uint32_t cap) {
    NEED(c);
    if (!c->accel_built || c->accel_mode != BPT_ACCEL_MERGED) return fail(c, BPT_ERR_STATE, "debug_read_wide: needs a built merged accel");
    const DevBvh& b = c->blas[0];
    if (cap < b.n) return fail(c, BPT_ERR_INVALID, "capacity too small");
    if (b.n < 2) return BPT_OK;
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (wide) BPT_CUDA_TRY(c, cudaMemcpy(wide, b.wide.p, (size_t)(b.n - 1) * 64, cudaMemcpyDeviceToHost));
    if (leafbox) BPT_CUDA_TRY(c, cudaMemcpy(leafbox, b.leafbox.p, (size_t)b.n * 32, cudaMemcpyDeviceToHost));
    return BPT_OK;
}

bpt_status bpt_clear_accum(bpt_context* c) {
    NEED(c);
    c->wf.ahead_slots = c->wf.ahead_cursor = 0;
    c->wf.accum_count = 0; c->wf.accum_fp16 = false; c->accum_used = false;
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->wf.accum.p, 0, (size_t)c->width * c->height * 16, c->stream));
    return BPT_OK;
}

// The accumulation buffer follows ONE rule between two bpt_clear_accum calls: FP32 sum (fp32 state) or the reference's
// running fp16 lerp (reference_fp16, pt_accumulate.hlsl:3-11); mixing them would make bpt_resolve meaningless.
static bpt_status check_precision(bpt_context* c, const bpt_settings* st) {
    if (st->state_precision != BPT_STATE_FP32 && st->state_precision != BPT_STATE_REFERENCE_FP16)
        return fail(c, BPT_ERR_INVALID, "state_precision: unknown value");
    if ((st->nee_mode != BPT_NEE_SHADOW_RAY && st->nee_mode != BPT_NEE_NONE) || st->rect_shadow > 1 || st->russian_roulette > 1 || st->pixel_jitter > 1)
        return fail(c, BPT_ERR_INVALID, "settings: unknown value of a mode switch (nee_mode, rect_shadow, russian_roulette, pixel_jitter)");
    const bool fp16 = st->state_precision == BPT_STATE_REFERENCE_FP16;
    if (c->accum_used && fp16 != c->wf.accum_fp16)
        return fail(c, BPT_ERR_STATE, "state_precision changed without bpt_clear_accum");
    c->accum_used = true; c->wf.accum_fp16 = fp16;
    return BPT_OK;
}

bpt_status bpt_render(bpt_context* c, const bpt_camera* cam, uint32_t first, uint32_t ns, const bpt_settings* st) {
    NEED(c);
    if (!cam || !st) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "render before build_accel");
    if (bpt_status ps = check_precision(c, st)) return ps;
    return wavefront_render(c, *cam, first, ns, *st);
}

bpt_status bpt_render_ahead(bpt_context* c, const bpt_camera* cam, uint32_t first, uint32_t max_samples, const bpt_settings* st, uint32_t* out_samples) {
    NEED(c);
    if (!cam || !st || !max_samples) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "render before build_accel");
    if (bpt_status ps = check_precision(c, st)) return ps;
    bpt_status s = wavefront_render(c, *cam, first, max_samples, *st, true);
    if (out_samples) *out_samples = c->wf.ahead_slots;
    return s;
}
bpt_status bpt_accumulate_ahead(bpt_context* c, uint32_t count) { NEED(c); return wavefront_accumulate_ahead(c, count); }
bpt_status bpt_accumulate_ahead_rgba16f(bpt_context* c, uint32_t total_samples, void* d_out) {
    NEED(c);
    if (!total_samples || !d_out) return BPT_ERR_INVALID;
    return wavefront_accumulate_ahead_rgba16f(c, total_samples, d_out);
}
bpt_status bpt_pending_ahead(bpt_context* c, uint32_t* pending, uint32_t* next_frame) {
    NEED(c);
    if (pending) *pending = c->wf.ahead_slots - c->wf.ahead_cursor;
    if (next_frame) *next_frame = c->wf.ahead_frame_first + c->wf.ahead_cursor;
    return BPT_OK;
}

bpt_status bpt_render_primary(bpt_context* c, const bpt_camera* cam, uint32_t frame_index, const bpt_settings* st, float* out_depth, bpt_gbuffer_texel* out_gbuffer) {
    NEED(c);
    if (!cam || !st) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "render_primary before build_accel");
    if (st->state_precision != BPT_STATE_FP32 && st->state_precision != BPT_STATE_REFERENCE_FP16) return fail(c, BPT_ERR_INVALID, "state_precision: unknown value");
    return wavefront_render_primary(c, *cam, frame_index, *st, out_depth, out_gbuffer);
}
bpt_status bpt_trace_ao(bpt_context* c, const bpt_camera* cam, uint32_t frame_index, const bpt_ao_settings* ao, const float* depth, const float* normal_roughness, float* out_ao) {
    NEED(c);
    if (!cam || !ao || !depth || !normal_roughness || !out_ao) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "trace_ao before build_accel");
    return wavefront_trace_ao(c, *cam, frame_index, *ao, depth, normal_roughness, out_ao);
}

bpt_status bpt_upscale_half_res(bpt_context* c, const bpt_camera* cam, uint32_t frame_index, const float* depth, const float* nr, const float* in_half, float* out) {
    NEED(c);
    if (!cam || !depth || !nr || !in_half || !out) return BPT_ERR_INVALID;
    return launch_upscale_half_res(c, *cam, frame_index, depth, nr, in_half, out);
}

bpt_status bpt_precompute_sky_ibl(bpt_context* c, const bpt_sky_ibl_desc* d) {
    NEED(c);
    if (!d) return BPT_ERR_INVALID;
    if (!d->diffuse_size || !d->specular_size || !d->brdf_lut_size || d->specular_levels < 2 || d->specular_levels > 12 ||
        (d->specular_size >> (d->specular_levels - 1)) == 0 || d->diffuse_size > 4096 || d->specular_size > 4096 || d->brdf_lut_size > 4096)
        return fail(c, BPT_ERR_INVALID, "sky ibl: sizes out of range (2 <= levels <= 12, last mip >= 1 texel)");
    c->ibl_valid = false;
    invalidate_ahead(c);
    bpt_status s = launch_precompute_sky_ibl(c, *d);
    if (s) return s;
    c->ibl_desc = *d; c->ibl_valid = true;
    return BPT_OK;
}

bpt_status bpt_debug_read_sky_ibl(bpt_context* c, float* diffuse, float* specular, float* brdf) {
    NEED(c);
    if (!c->ibl_valid) return fail(c, BPT_ERR_STATE, "sky ibl not computed");
    const bpt_sky_ibl_desc& d = c->ibl_desc;
    size_t spec = 0;
    for (uint32_t l = 0; l < d.specular_levels; l++) { size_t n = d.specular_size >> l; spec += 6 * n * n * 16; }
    if (diffuse) BPT_CUDA_TRY(c, cudaMemcpyAsync(diffuse, c->d_ibl_diffuse.p, (size_t)6 * d.diffuse_size * d.diffuse_size * 16, cudaMemcpyDeviceToHost, c->stream));
    if (specular) BPT_CUDA_TRY(c, cudaMemcpyAsync(specular, c->d_ibl_specular.p, spec, cudaMemcpyDeviceToHost, c->stream));
    if (brdf) BPT_CUDA_TRY(c, cudaMemcpyAsync(brdf, c->d_ibl_brdf.p, (size_t)d.brdf_lut_size * d.brdf_lut_size * 8, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

bpt_status bpt_trace_reflection(bpt_context* c, const bpt_camera* cam, uint32_t frame_index, const bpt_reflection_settings* rs, const float* depth,
                                const bpt_gbuffer_texel* gbuffer, float* out_reflection, float* out_hit_positions) {
    NEED(c);
    if (!cam || !rs || !depth || !gbuffer || !out_reflection || !out_hit_positions) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "trace_reflection before build_accel");
    if (rs->ibl > 1 || (rs->ibl && !c->ibl_valid)) return fail(c, rs->ibl > 1 ? BPT_ERR_INVALID : BPT_ERR_STATE, "trace_reflection: settings.ibl needs bpt_precompute_sky_ibl");
    return wavefront_trace_reflection(c, *cam, frame_index, *rs, depth, gbuffer, out_reflection, out_hit_positions);
}

bpt_status bpt_set_ddgi_volume(bpt_context* c, const bpt_probe_volume* vol, const bpt_probe_blend* bl, const float* irr, const float* vis) {
    NEED(c);
    invalidate_ahead(c);
    if (!vol || !bl || !irr || !vis) { c->ddgi_enabled = false; return BPT_OK; }        // unbind
    const uint64_t nx = vol->probe_counts[0], ny = vol->probe_counts[1], nz = vol->probe_counts[2];
    if (!nx || !ny || !nz || bl->irradiance_size < 2 || bl->visibility_size < 2 || bl->irradiance_size > 30 || bl->visibility_size > 30)
        return fail(c, BPT_ERR_INVALID, "set_ddgi_volume: bad sizes");
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    bpt_status s;
    if ((s = dev_upload(c, c->d_ddgi_irr, irr, nx * ny * (bl->irradiance_size + 2) * nz * (bl->irradiance_size + 2) * 16))) return s;
    if ((s = dev_upload(c, c->d_ddgi_vis, vis, nx * ny * (bl->visibility_size + 2) * nz * (bl->visibility_size + 2) * 8))) return s;
    c->ddgi_volume = *vol; c->ddgi_irr_size = bl->irradiance_size; c->ddgi_vis_size = bl->visibility_size; c->ddgi_enabled = true;
    return BPT_OK;
}
bpt_status bpt_ddgi_lighting(bpt_context* c, uint64_t n, const float* pos, const float* normal, const float* view, float* out) {
    NEED(c);
    if (!pos || !normal || !view || !out) return BPT_ERR_INVALID;
    if (!c->ddgi_enabled) return fail(c, BPT_ERR_STATE, "ddgi_lighting: no volume bound (bpt_set_ddgi_volume)");
    return launch_ddgi_lighting(c, n, pos, normal, view, out);
}

bpt_status bpt_resolve_device(bpt_context* c, uint32_t total, float* d_out) {
    NEED(c);
    if (!total || !d_out) return BPT_ERR_INVALID;
    return launch_resolve(c, total, d_out);
}

bpt_status bpt_resolve_device_rgba16f(bpt_context* c, uint32_t total, void* d_out) {
    NEED(c);
    if (!total || !d_out) return BPT_ERR_INVALID;
    return launch_resolve_rgba16f(c, total, d_out);
}

bpt_status bpt_denoise_reblur(bpt_context* c, const bpt_camera* cam, uint64_t frame_count, const bpt_reblur_settings* st, const bpt_reblur_inputs* in, float* out) {
    NEED(c);
    if (!cam || !st || !in || !out) return BPT_ERR_INVALID;
    return launch_reblur(c, *cam, frame_count, *st, *in, out);
}
bpt_status bpt_reblur_reset(bpt_context* c) { NEED(c); return reblur_reset(c); }
bpt_status bpt_debug_read_reblur(bpt_context* c, uint32_t which, float* out, uint64_t cap) { NEED(c); if (!out) return BPT_ERR_INVALID; return reblur_debug_read(c, which, out, cap); }

bpt_status bpt_post_process_device(bpt_context* c, const bpt_post_settings* st, uint32_t total, float* d_out) {
    NEED(c);
    if (!st || !total || !d_out) return BPT_ERR_INVALID;
    if (st->bloom > 1 || !(st->bloom_threshold_softness >= 0.0f && st->bloom_threshold_softness <= 1.0f) || !(st->bloom_threshold == st->bloom_threshold))
        return fail(c, BPT_ERR_INVALID, "post settings: bloom must be 0/1, softness in [0, 1]");
    return launch_post_process(c, *st, total, d_out);
}

bpt_status bpt_post_process(bpt_context* c, const bpt_post_settings* st, uint32_t total, float* out) {
    NEED(c);
    if (!out) return BPT_ERR_INVALID;
    size_t bytes = (size_t)c->width * c->height * 16;
    bpt_status s = dev_reserve(c, c->d_post_out, bytes);
    if (s) return s;
    if ((s = bpt_post_process_device(c, st, total, c->d_post_out.as<float>()))) return s;
    BPT_CUDA_TRY(c, cudaMemcpyAsync(out, c->d_post_out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

bpt_status bpt_resolve(bpt_context* c, uint32_t total, float* out) {
    NEED(c);
    if (!total || !out) return BPT_ERR_INVALID;
    DevBuf tmp;
    size_t bytes = (size_t)c->width * c->height * 16;
    bpt_status s = dev_alloc(c, tmp, bytes);
    if (s) return s;
    s = launch_resolve(c, total, tmp.as<float>());
    cudaError_t e = cudaSuccess;
    if (!s) e = cudaMemcpyAsync(out, tmp.p, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (!s && e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    dev_free(tmp);
    if (s) return s;
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

bpt_status bpt_accum_device_ptr(bpt_context* c, float** out) { NEED(c); if (!out) return BPT_ERR_INVALID; *out = c->wf.accum.as<float>(); return BPT_OK; }

bpt_status bpt_upload_accum(bpt_context* c, const float* sums) {
    NEED(c);
    if (!sums) return BPT_ERR_INVALID;
    BPT_CUDA_TRY(c, cudaMemcpyAsync(c->wf.accum.p, sums, (size_t)c->width * c->height * 16, cudaMemcpyDefault, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BPT_OK;
}

bpt_status bpt_get_counters(bpt_context* c, bpt_counters* out) {
    NEED(c);
    if (!out) return BPT_ERR_INVALID;
    uint64_t t[40];
    BPT_CUDA_TRY(c, cudaMemcpyAsync(t, c->wf.totals.p, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    memset(out, 0, sizeof(*out));
    for (int b = 0; b < 16; b++) {
        out->extend_rays_per_bounce[b] = t[b]; out->shadow_rays_per_bounce[b] = t[16 + b];
        out->extend_rays += t[b]; out->shadow_rays += t[16 + b];
    }
    out->samples = t[32];
    out->kernel_launches = c->launches;
    return BPT_OK;
}

bpt_status bpt_reset_counters(bpt_context* c) {
    NEED(c);
    BPT_CUDA_TRY(c, cudaMemsetAsync(c->wf.totals.p, 0, 40 * sizeof(uint64_t), c->stream));
    c->launches = 0;
    return BPT_OK;
}

bpt_status bpt_profile_enable(bpt_context* c, uint32_t en) {
    NEED(c);
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (auto& e : c->prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    c->prof_events.clear();
    c->profile = en != 0;
    return BPT_OK;
}
bpt_status bpt_profile_read(bpt_context* c, bpt_kernel_times* out) {
    NEED(c);
    if (!out) return BPT_ERR_INVALID;
    BPT_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    double ms[5] = {0, 0, 0, 0, 0}; uint64_t n[5] = {0, 0, 0, 0, 0};
    for (auto& e : c->prof_events) {
        float t = 0.0f;
        BPT_CUDA_TRY(c, cudaEventElapsedTime(&t, e.a, e.b));
        ms[e.cls] += t; n[e.cls]++;
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    c->prof_events.clear();
    *out = bpt_kernel_times{ms[0], ms[1], ms[2], ms[3], ms[4], n[0], n[1], n[2], n[3], n[4]};
    return BPT_OK;
}

bpt_status bpt_trace_rays(bpt_context* c, const bpt_ray* rays, uint64_t n, uint32_t frame, bpt_hit* out) {
    NEED(c);
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "accel not built");
    if (n && (!rays || !out)) return BPT_ERR_INVALID;
    return launch_trace_batch(c, rays, n, frame, out, nullptr);
}
bpt_status bpt_trace_shadow_rays(bpt_context* c, const bpt_ray* rays, uint64_t n, uint32_t frame, uint8_t* vis) {
    NEED(c);
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "accel not built");
    if (n && (!rays || !vis)) return BPT_ERR_INVALID;
    return launch_trace_batch(c, rays, n, frame, nullptr, vis);
}

bpt_status bpt_debug_capture(bpt_context* c, uint32_t en) { NEED(c); c->capture = en != 0; return BPT_OK; }

bpt_status bpt_debug_read_queue(bpt_context* c, uint32_t bounce, uint32_t kind, uint32_t* pixels, uint32_t* lights, bpt_hit* hits, uint64_t cap, uint64_t* count) {
    NEED(c);
    auto& px = kind == 0 ? c->cap_extend_pixels : c->cap_shadow_pixels;
    if (bounce >= px.size()) { if (count) *count = 0; return BPT_OK; }
    uint64_t n = px[bounce].size();
    if (count) *count = n;
    if (!pixels && !lights && !hits) return BPT_OK;
    if (cap < n) return fail(c, BPT_ERR_INVALID, "capacity too small");
    if (pixels) std::copy(px[bounce].begin(), px[bounce].end(), pixels);
    if (kind == 0 && hits) std::copy(c->cap_extend_hits[bounce].begin(), c->cap_extend_hits[bounce].end(), hits);
    if (kind == 1 && lights) std::copy(c->cap_shadow_lights[bounce].begin(), c->cap_shadow_lights[bounce].end(), lights);
    return BPT_OK;
}

bpt_status bpt_trace_probes(bpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, uint32_t bounces, float* out) {
    NEED(c);
    if (!vol || !table || !out) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "trace_probes before build_accel");
    return wavefront_trace_probes(c, *vol, table, frame, bounces, out);
}

bpt_status bpt_trace_probes_range(bpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, uint32_t bounces, uint32_t first_probe,
                                  uint32_t num_probes, float* out) {
    NEED(c);
    if (!vol || !table || (!out && num_probes) || num_probes == 0xffffffffu) return BPT_ERR_INVALID;
    if (!c->accel_built) return fail(c, BPT_ERR_STATE, "trace_probes before build_accel");
    return wavefront_trace_probes(c, *vol, table, frame, bounces, out, first_probe, num_probes);
}

bpt_status bpt_blend_probes(bpt_context* c, const bpt_probe_volume* vol, const float* table, uint32_t frame, const float* rays,
                            const bpt_probe_blend* blend, float* irr, float* vis) {
    NEED(c);
    if (!vol || !table || !rays || !blend || !irr || !vis) return BPT_ERR_INVALID;
    return launch_blend_probes(c, *vol, table, frame, rays, *blend, irr, vis);
}

} // extern "C"
