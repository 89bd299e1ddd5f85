// ibl.cu — SkyboxPrecomputePass (src/renderer/pass/skybox_precompute.cpp:66-162) on the GPU: BRDF LUT, diffuse irradiance cube,
// prefiltered specular cube (5 mips). Per-texel functions and the numeric contract: bpt_ibl.cuh. One thread per texel runs the
// shader's whole sample loop (1024 / 4096 samples), so the sums have the shader's order; the sky cubemap (6 MB at 256^2) is
// L2-resident and the kernels are FP32 / L1 bound. The reference runs these once per skybox change, not per frame.
#include "bpt_internal.cuh"
#include "bpt_ibl.cuh"

using namespace bptd;

namespace {

__global__ void __launch_bounds__(128) k_ibl_brdf_lut(float2* __restrict__ lut, uint32_t res) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= res * res) return;
    lut[i] = ibl_brdf_lut_texel(i % res, i / res, res);
}
__global__ void __launch_bounds__(128) k_ibl_diffuse(const float4* __restrict__ sky, uint32_t sky_size, float4* __restrict__ out, uint32_t size) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6u * size * size) return;
    uint32_t layer = i / (size * size), r = i % (size * size);
    float3 c = ibl_diffuse_texel(sky, sky_size, r % size, r / size, layer, size);
    out[i] = make_float4(c.x, c.y, c.z, 1.0f);
}
__global__ void __launch_bounds__(128) k_ibl_specular(const float4* __restrict__ sky, uint32_t sky_size, float4* __restrict__ out, uint32_t size, float roughness) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6u * size * size) return;
    uint32_t layer = i / (size * size), r = i % (size * size);
    float3 c = ibl_specular_texel(sky, sky_size, r % size, r / size, layer, size, roughness);
    out[i] = make_float4(c.x, c.y, c.z, 1.0f);
}

} // namespace

bpt_status launch_precompute_sky_ibl(bpt_context* ctx, const bpt_sky_ibl_desc& d) {
    bpt_status s;
    size_t spec_texels = 0;
    for (uint32_t l = 0; l < d.specular_levels; l++) { size_t n = d.specular_size >> l; spec_texels += 6 * n * n; }
    if ((s = dev_reserve(ctx, ctx->d_ibl_brdf, (size_t)d.brdf_lut_size * d.brdf_lut_size * 8))) return s;
    if ((s = dev_reserve(ctx, ctx->d_ibl_diffuse, (size_t)6 * d.diffuse_size * d.diffuse_size * 16))) return s;
    if ((s = dev_reserve(ctx, ctx->d_ibl_specular, spec_texels * 16))) return s;
    const float4* sky = ctx->d_sky.as<float4>();
    const uint32_t sky_size = ctx->sky_size;                     // 0: no skybox -> black (scene_basic/skybox.cpp:43-53)
    auto check = [&](const char* what) -> bpt_status {
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { ctx->err = std::string("launch of ") + what + ": " + cudaGetErrorString(e); return BPT_ERR_CUDA; }
        return BPT_OK;
    };
    uint32_t n = d.brdf_lut_size * d.brdf_lut_size;
    k_ibl_brdf_lut<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_ibl_brdf.as<float2>(), d.brdf_lut_size);
    if ((s = check("k_ibl_brdf_lut"))) return s;
    n = 6u * d.diffuse_size * d.diffuse_size;
    k_ibl_diffuse<<<(n + 127) / 128, 128, 0, ctx->stream>>>(sky, sky_size, ctx->d_ibl_diffuse.as<float4>(), d.diffuse_size);
    if ((s = check("k_ibl_diffuse"))) return s;
    size_t offset = 0;
    for (uint32_t l = 0; l < d.specular_levels; l++) {
        const uint32_t size = d.specular_size >> l;
        n = 6u * size * size;
        k_ibl_specular<<<(n + 127) / 128, 128, 0, ctx->stream>>>(sky, sky_size, ctx->d_ibl_specular.as<float4>() + offset, size,
                                                                 (float)l / (float)(d.specular_levels - 1));      // skybox_precompute.cpp:148
        if ((s = check("k_ibl_specular"))) return s;
        offset += n;
    }
    return BPT_OK;
}
