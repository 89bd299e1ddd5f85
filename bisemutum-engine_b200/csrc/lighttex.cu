// lighttex.cu — rect-light textures: level 0 decode + the engine's mip chain, on the device.
//
// Replaces what happens to a TextureAsset with levels > 1 at upload (bisemutum/src/scene_basic/texture.cpp:155-189):
// copy_buffer_to_texture of level 0, then GraphicsManager::generate_mipmaps_2d (src/graphics/command_helpers.cpp:66-215), which runs
// shaders/core/mipmap.hlsl (MIPMAP_MODE_AVG) once per level: a destination texel is the mean of the 2x2 source texels at twice its
// coordinate, plus a third column / row / corner texel when the DESTINATION extent is odd in that direction (the shader tests
// `push_c.tex_size & 1` on the size it was dispatched with, the destination's), divided by the number of texels taken. A source
// coordinate past the level returns 0 (robust image access). Every level is stored in the texture's own format: unorm8 values are
// rounded to n / 255, an sRGB target encodes and the sampler decodes again, FP32 is kept.
//
// The chain is kept as FP32 texels (what a sampler hands to the shader), level after level, and filtered in explicit FP32 arithmetic
// (bpt_scene.cuh: light_texture_sample).
#include "bpt_internal.cuh"

using namespace bptd;

namespace {

// sRGB <-> linear on 8-bit codes. decode[c] = EOTF(c / 255); encode(x) = the code whose interval [thr[c], thr[c + 1]) holds x, with
// thr[c] = EOTF((c - 0.5) / 255): rounding to the nearest code in the encoded domain, expressed as comparisons in the linear domain so
// that no pow() runs on the device. Both tables are computed on the host in double precision.
struct SrgbTables { float decode[256]; float thr[256]; };

__device__ __forceinline__ float store_value(float x, uint32_t format, int channel, const SrgbTables* tb) {
    if (format == BPT_TEXTURE_RGBA32_FLOAT) return x;
    if (format == BPT_TEXTURE_RGBA8_UNORM || channel == 3) return q_unorm(x, 255.0f);
    // rgba8_srgb colour channel: encode, then what the sampler decodes
    if (!(x > 0.0f)) return 0.0f;
    int lo = 0, hi = 255;                       // largest c with thr[c] <= x (thr[0] < 0)
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (tb->thr[mid] <= x) lo = mid; else hi = mid - 1; }
    return tb->decode[lo];
}

__global__ void k_light_tex_level0(const void* __restrict__ src, uint32_t n, uint32_t format, const SrgbTables* __restrict__ tb, float4* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (format == BPT_TEXTURE_RGBA32_FLOAT) { dst[i] = reinterpret_cast<const float4*>(src)[i]; return; }
    uchar4 p = reinterpret_cast<const uchar4*>(src)[i];
    if (format == BPT_TEXTURE_RGBA8_UNORM) dst[i] = make_float4((float)p.x / 255.0f, (float)p.y / 255.0f, (float)p.z / 255.0f, (float)p.w / 255.0f);
    else dst[i] = make_float4(tb->decode[p.x], tb->decode[p.y], tb->decode[p.z], (float)p.w / 255.0f);
}

// shaders/core/mipmap.hlsl:46-93 (compute path; the graphics path used for sRGB targets runs the same body per fragment)
__global__ void k_mip_downsample(const float4* __restrict__ src, uint32_t sw, uint32_t sh, float4* __restrict__ dst, uint32_t dw, uint32_t dh,
                                 uint32_t format, const SrgbTables* __restrict__ tb) {
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    auto at = [&](uint32_t xx, uint32_t yy) { return (xx < sw && yy < sh) ? src[(size_t)yy * sw + xx] : make_float4(0.0f, 0.0f, 0.0f, 0.0f); };
    auto add = [](float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); };
    const uint32_t sx = 2 * x, sy = 2 * y;
    float4 r = add(add(at(sx, sy), at(sx, sy + 1)), add(at(sx + 1, sy), at(sx + 1, sy + 1)));
    uint32_t num = 4;
    const bool odd_x = (dw & 1u) != 0, odd_y = (dh & 1u) != 0;
    if (odd_x) { r = add(r, add(at(sx + 2, sy), at(sx + 2, sy + 1))); num += 2; }
    if (odd_y) { r = add(r, add(at(sx, sy + 2), at(sx + 1, sy + 2))); num += 2; }
    if (odd_x && odd_y) { r = add(r, at(sx + 2, sy + 2)); num += 1; }
    const float d = (float)num;
    dst[(size_t)y * dw + x] = make_float4(store_value(r.x / d, format, 0, tb), store_value(r.y / d, format, 1, tb), store_value(r.z / d, format, 2, tb),
                                          store_value(r.w / d, format, 3, tb));
}

} // namespace

static uint64_t chain_texels(uint32_t w, uint32_t h, uint32_t levels) {
    uint64_t n = 0;
    for (uint32_t l = 0; l < levels; l++) n += (uint64_t)std::max(w >> l, 1u) * std::max(h >> l, 1u);
    return n;
}

bpt_status upload_light_textures(bpt_context* ctx, const bpt_light_texture_desc* t, uint32_t nt) {
    for (auto& b : ctx->d_light_texels) dev_free(b);
    ctx->d_light_texels.assign(nt, DevBuf{});
    ctx->h_light_textures.assign(nt, DLightTexture{});
    bpt_status s;
    if (nt == 0) return BPT_OK;
    if (!ctx->d_srgb_tables.p) {
        SrgbTables tb;
        auto eotf = [](double v) { return v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4); };
        for (int c = 0; c < 256; c++) { tb.decode[c] = (float)eotf(c / 255.0); tb.thr[c] = c == 0 ? -1.0f : (float)eotf((c - 0.5) / 255.0); }
        if ((s = dev_upload(ctx, ctx->d_srgb_tables, &tb, sizeof(tb)))) return s;
        BPT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));         // `tb` is a stack object
    }
    const SrgbTables* tb = ctx->d_srgb_tables.as<SrgbTables>();
    for (uint32_t i = 0; i < nt; i++) {
        const bpt_light_texture_desc& d = t[i];
        if (!d.texels || !d.width || !d.height || d.width > 16384 || d.height > 16384 || d.format > BPT_TEXTURE_RGBA8_SRGB || d.levels == 0 ||
            d.address_mode_u > BPT_ADDRESS_CLAMP || d.address_mode_v > BPT_ADDRESS_CLAMP) { ctx->err = "bad light texture desc"; return BPT_ERR_INVALID; }
        uint32_t full = 1; while ((std::max(d.width, d.height) >> full) != 0) full++;       // floor(log2(max)) + 1 (command_helpers.cpp:179)
        const uint32_t levels = std::min(d.levels, full);
        const uint64_t total = chain_texels(d.width, d.height, levels);
        if ((s = dev_alloc(ctx, ctx->d_light_texels[i], total * 16))) return s;
        float4* chain = ctx->d_light_texels[i].as<float4>();
        const uint32_t n0 = d.width * d.height;
        DevBuf staging;
        if ((s = dev_upload(ctx, staging, d.texels, (size_t)n0 * (d.format == BPT_TEXTURE_RGBA32_FLOAT ? 16 : 4)))) return s;
        k_light_tex_level0<<<(n0 + 255) / 256, 256, 0, ctx->stream>>>(staging.p, n0, d.format, tb, chain);
        ctx->launches++;
        uint64_t off = 0;
        for (uint32_t l = 0; l + 1 < levels; l++) {
            const uint32_t sw = std::max(d.width >> l, 1u), sh = std::max(d.height >> l, 1u), dw = std::max(sw / 2, 1u), dh = std::max(sh / 2, 1u);
            const float4* src = chain + off;
            off += (uint64_t)sw * sh;
            k_mip_downsample<<<dim3((dw + 15) / 16, (dh + 15) / 16), dim3(16, 16), 0, ctx->stream>>>(src, sw, sh, chain + off, dw, dh, d.format, tb);
            ctx->launches++;
        }
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // the caller's texels and the staging copy are released now
        dev_free(staging);
        if (e != cudaSuccess) { ctx->err = std::string("light texture upload: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
        ctx->h_light_textures[i] = DLightTexture{chain, d.width, d.height, levels, d.address_mode_u, d.address_mode_v, d.filter_linear ? 1u : 0u, d.mip_linear ? 1u : 0u};
    }
    if ((s = dev_upload(ctx, ctx->d_light_textures, ctx->h_light_textures.data(), (size_t)nt * sizeof(DLightTexture)))) return s;
    BPT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return BPT_OK;
}

bpt_status read_light_texture(bpt_context* ctx, uint32_t index, float* out, uint64_t cap, uint64_t* out_texels) {
    if (index >= ctx->h_light_textures.size()) { ctx->err = "light texture index out of range"; return BPT_ERR_INVALID; }
    const DLightTexture& t = ctx->h_light_textures[index];
    const uint64_t total = chain_texels(t.w, t.h, t.levels);
    if (out_texels) *out_texels = total;
    if (!out) return BPT_OK;
    if (cap < total) { ctx->err = "capacity too small"; return BPT_ERR_INVALID; }
    BPT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    BPT_CUDA_TRY(ctx, cudaMemcpy(out, t.texels, total * 16, cudaMemcpyDeviceToHost));
    return BPT_OK;
}
