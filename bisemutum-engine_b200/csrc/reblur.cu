// reblur.cu — bpt_denoise_reblur: the eight passes of ReblurPass::render (bisemutum/src/renderer/pass/reblur.cpp:273-588) as kernels.
// One thread per denoiser pixel (the reference dispatches 8 x 8 groups, reblur.cpp:344-347); "Gen Depth Mip" keeps the reference's
// 64-thread group per 16 x 16 tile because its coarser levels are reduced from the unrounded level below in shared memory
// (gen_depth_mip.hlsl:34-118). The history the reference parks on the camera (Camera::add_history_texture: "reblur_lighting_dist_0",
// "reblur_lighting_dist_1", "reblur_accumulation", plus the G-buffer's "depth" and "gbuffer_normal_roughness") lives in the context.
#include <cmath>
#include <nvtx3/nvToolsExt.h>
#include "bpt_internal.cuh"
#include "bpt_reblur.cuh"

using namespace bptd;

namespace {

constexpr int kTile = 8;

template <int PASS>
__global__ void __launch_bounds__(kTile * kTile) k_reblur_pass(const __grid_constant__ ReblurView rv) {
    const int x = blockIdx.x * kTile + threadIdx.x, y = blockIdx.y * kTile + threadIdx.y;
    if (x >= (int)rv.w || y >= (int)rv.h) return;
    if (PASS == 1) reblur_pre_blur(rv, x, y);
    else if (PASS == 2) reblur_temporal_accumulate(rv, x, y);
    else if (PASS == 3) reblur_fetch_linear_depth(rv, x, y);
    else if (PASS == 5) reblur_fix_history(rv, x, y);
    else if (PASS == 6) reblur_blur(rv, x, y);
    else if (PASS == 7) reblur_temporal_stabilize(rv, x, y);
    else if (PASS == 8) reblur_post_blur(rv, x, y);
}

__global__ void __launch_bounds__(64) k_reblur_gen_depth_mip(const __grid_constant__ ReblurView rv) {
    __shared__ float4 s_v[64];
    __shared__ float s_d[64];
    const uint32_t local = threadIdx.x;
    float4 v; float d; int px, py;
    rb_mip_level1(rv, (int)blockIdx.x, (int)blockIdx.y, local, v, d, px, py);
    s_v[local] = v; s_d[local] = d;
    rb_mip_store(rv, 1, px >> 1, py >> 1, v, d);
    __syncthreads();
    if ((local & 3u) == 0u) {
        float4 vv[4] = {s_v[local], s_v[local + 1], s_v[local + 2], s_v[local + 3]};
        float dd[4] = {s_d[local], s_d[local + 1], s_d[local + 2], s_d[local + 3]};
        rb_mip_reduce(vv, dd, v, d);
    }
    __syncthreads();                                    // every read of level 1 happens before a level-2 value replaces it
    if ((local & 3u) == 0u) {
        s_v[local] = v; s_d[local] = d;
        rb_mip_store(rv, 2, px >> 2, py >> 2, v, d);
    }
    __syncthreads();
    if ((local & 15u) == 0u) {
        float4 vv[4] = {s_v[local], s_v[local + 4], s_v[local + 8], s_v[local + 12]};
        float dd[4] = {s_d[local], s_d[local + 4], s_d[local + 8], s_d[local + 12]};
        rb_mip_reduce(vv, dd, v, d);
        rb_mip_store(rv, 3, px >> 3, py >> 3, v, d);
    }
}

// reblur.cpp:174-197: the per-frame Poisson rotators
const float kPreRot[32] = {0.840188f, 0.394383f, 0.783099f, 0.79844f, 0.911647f, 0.197551f, 0.335223f, 0.76823f, 0.277775f, 0.55397f, 0.477397f, 0.628871f, 0.364784f, 0.513401f,
                           0.95223f, 0.916195f, 0.635712f, 0.717297f, 0.141603f, 0.606969f, 0.0163006f, 0.242887f, 0.137232f, 0.804177f, 0.156679f, 0.400944f, 0.12979f, 0.108809f,
                           0.998924f, 0.218257f, 0.512932f, 0.839112f};
const float kBlurRot[32] = {0.61264f, 0.296032f, 0.637552f, 0.524287f, 0.493583f, 0.972775f, 0.292517f, 0.771358f, 0.526745f, 0.769914f, 0.400229f, 0.891529f, 0.283315f, 0.352458f,
                            0.807725f, 0.919026f, 0.0697553f, 0.949327f, 0.525995f, 0.0860558f, 0.192214f, 0.663227f, 0.890233f, 0.348893f, 0.0641713f, 0.020023f, 0.457702f,
                            0.0630958f, 0.23828f, 0.970634f, 0.902208f, 0.85092f};
const float kPostRot[32] = {0.266666f, 0.53976f, 0.375207f, 0.760249f, 0.512535f, 0.667724f, 0.531606f, 0.0392803f, 0.437638f, 0.931835f, 0.93081f, 0.720952f, 0.284293f, 0.738534f,
                            0.639979f, 0.354049f, 0.687861f, 0.165974f, 0.440105f, 0.880075f, 0.829201f, 0.330337f, 0.228968f, 0.893372f, 0.35036f, 0.68667f, 0.956468f, 0.58864f,
                            0.657304f, 0.858676f, 0.43956f, 0.92397f};
float4 rotator(float angle) { float ca = std::cos(angle), sa = std::sin(angle); return make_float4(ca, sa, -sa, ca); }

} // namespace

bpt_status reblur_reset(bpt_context* ctx) { ctx->reblur.has_history = false; return BPT_OK; }

bpt_status launch_reblur(bpt_context* ctx, const bpt_camera& cam, uint64_t frame_count, const bpt_reblur_settings& st, const bpt_reblur_inputs& in, float* h_out) {
    ReblurState& rs = ctx->reblur;
    const uint32_t w = in.width, h = in.height, gw = ctx->width, gh = ctx->height;
    const bool half = w != gw;                                                          // reblur.cpp:280
    if (w < 8 || h < 8 || (half && (w != (gw + 1) / 2 || h != (gh + 1) / 2)) || (!half && h != gh)) { ctx->err = "reblur: the noised image must have the camera's extent or half of it (>= 8 x 8)"; return BPT_ERR_INVALID; }
    if (!in.noised || !in.hit_positions || !in.depth || !in.normal_roughness) { ctx->err = "reblur: null input"; return BPT_ERR_INVALID; }
    const size_t n = (size_t)w * h, gn = (size_t)gw * gh;
    const size_t chain = reblur_mip_offset(w, h, 4);
    bpt_status s;
    if (rs.w != w || rs.h != h || rs.gw != gw || rs.gh != gh) { rs.has_history = false; rs.w = w; rs.h = h; rs.gw = gw; rs.gh = gh; }
    if ((s = dev_reserve(ctx, rs.ld0, chain * 16)) || (s = dev_reserve(ctx, rs.ld1, n * 16)) || (s = dev_reserve(ctx, rs.accum, n * 4)) ||
        (s = dev_reserve(ctx, rs.lin_depth, chain * 4)) || (s = dev_reserve(ctx, rs.denoised, n * 16)) || (s = dev_reserve(ctx, rs.hist_ld0, n * 16)) ||
        (s = dev_reserve(ctx, rs.hist_ld1, n * 16)) || (s = dev_reserve(ctx, rs.hist_accum, n * 4)) || (s = dev_reserve(ctx, rs.depth[0], gn * 4)) ||
        (s = dev_reserve(ctx, rs.depth[1], gn * 4)) || (s = dev_reserve(ctx, rs.nr[0], gn * 16)) || (s = dev_reserve(ctx, rs.nr[1], gn * 16)) ||
        (s = dev_reserve(ctx, rs.velocity, gn * 8)) || (s = dev_reserve(ctx, rs.validation, n)) || (s = dev_reserve(ctx, rs.noised, n * 16)) ||
        (s = dev_reserve(ctx, rs.hit, n * 16))) return s;
    const bool has_history = rs.has_history && rs.last_frame + 1 == frame_count;          // reblur.cpp:282-285,357-361
    const int cur = rs.cur ^ 1;                                                            // G-buffer ping-pong: [cur] = this frame, [cur ^ 1] = the previous one
    cudaStream_t q = ctx->stream;
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.depth[cur].p, in.depth, gn * 4, cudaMemcpyDefault, q));
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.nr[cur].p, in.normal_roughness, gn * 16, cudaMemcpyDefault, q));
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.noised.p, in.noised, n * 16, cudaMemcpyDefault, q));
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.hit.p, in.hit_positions, n * 16, cudaMemcpyDefault, q));
    if (in.velocity) BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.velocity.p, in.velocity, gn * 8, cudaMemcpyDefault, q));
    if (in.history_validation) BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.validation.p, in.history_validation, n, cudaMemcpyDefault, q));

    ReblurView rv{};
    rv.w = w; rv.h = h; rv.gw = gw; rv.gh = gh; rv.half_res = half ? 1u : 0u; rv.frame_index = (uint32_t)frame_count;
    rv.has_history = has_history ? 1u : 0u; rv.virtual_history = st.virtual_history;
    rv.blur_radius = st.blur_radius; rv.anti_flicker = st.anti_flickering_strength;
    rv.cam = cam; rv.hist_cam = has_history ? rs.last_cam : cam;
    const uint32_t ri = (uint32_t)(frame_count % 32);                                      // reblur.cpp:322
    rv.rot_pre = rotator(kPreRot[ri]); rv.rot_blur = rotator(kBlurRot[ri]); rv.rot_post = rotator(kPostRot[ri]);
    rv.depth = rs.depth[cur].as<float>(); rv.normal_roughness = rs.nr[cur].as<float4>();
    rv.velocity = in.velocity ? rs.velocity.as<float2>() : nullptr; rv.validation = in.history_validation ? rs.validation.as<uint8_t>() : nullptr;
    rv.hit_positions = rs.hit.as<float4>(); rv.noised = rs.noised.as<float4>();
    rv.hist_depth = rs.depth[cur ^ 1].as<float>(); rv.hist_normal_roughness = rs.nr[cur ^ 1].as<float4>();
    rv.hist_ld0 = rs.hist_ld0.as<float4>(); rv.hist_ld1 = rs.hist_ld1.as<float4>(); rv.hist_accum = rs.hist_accum.as<float>();
    rv.ld0 = rs.ld0.as<float4>(); rv.ld1 = rs.ld1.as<float4>(); rv.accum = rs.accum.as<float>(); rv.lin_depth = rs.lin_depth.as<float>();
    rv.denoised = rs.denoised.as<float4>();

    const dim3 grid((w + kTile - 1) / kTile, (h + kTile - 1) / kTile), block(kTile, kTile);
    auto range = [](const char* name) { nvtxRangePushA(name); };
    range("ReBLUR Pre Blur"); k_reblur_pass<1><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    range("ReBLUR Temporal Accumulate"); k_reblur_pass<2><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    range("ReBLUR Fetch Linear Depth"); k_reblur_pass<3><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    range("ReBLUR Gen Depth Mip"); k_reblur_gen_depth_mip<<<dim3((w + 15) / 16, (h + 15) / 16), 64, 0, q>>>(rv); nvtxRangePop();
    range("ReBLUR Fix History"); k_reblur_pass<5><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    range("ReBLUR Blur"); k_reblur_pass<6><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    // the blurred image and the accumulation speed are next frame's history (reblur.cpp:512-513); the stabilise pass below still reads
    // LAST frame's stabilised image, so that one is replaced after it ran (:554)
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.hist_ld0.p, rs.ld0.p, n * 16, cudaMemcpyDeviceToDevice, q));
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.hist_accum.p, rs.accum.p, n * 4, cudaMemcpyDeviceToDevice, q));
    range("ReBLUR Temporal Stabilize"); k_reblur_pass<7><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rs.hist_ld1.p, rs.ld1.p, n * 16, cudaMemcpyDeviceToDevice, q));
    range("ReBLUR Post Blur"); k_reblur_pass<8><<<grid, block, 0, q>>>(rv); nvtxRangePop();
    ctx->launches += 8;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && h_out) e = cudaMemcpyAsync(h_out, rs.denoised.p, n * 16, cudaMemcpyDefault, q);
    if (e == cudaSuccess) e = cudaStreamSynchronize(q);
    if (e != cudaSuccess) { ctx->err = std::string("reblur: ") + cudaGetErrorString(e); rs.has_history = false; return BPT_ERR_CUDA; }
    rs.has_history = true; rs.last_frame = frame_count; rs.last_cam = cam; rs.cur = cur;
    return BPT_OK;
}

bpt_status reblur_debug_read(bpt_context* ctx, uint32_t which, float* out, uint64_t capacity_floats) {
    ReblurState& rs = ctx->reblur;
    const size_t n = (size_t)rs.w * rs.h, chain = rs.w ? reblur_mip_offset(rs.w, rs.h, 4) : 0;
    const DevBuf* src = nullptr; size_t floats = 0;
    switch (which) {
        case 0: src = &rs.ld0; floats = chain * 4; break;       // lighting_dist_0 with its mips (level 0 = the blurred image)
        case 1: src = &rs.ld1; floats = n * 4; break;           // lighting_dist_1 (the stabilised image)
        case 2: src = &rs.accum; floats = n; break;
        case 3: src = &rs.lin_depth; floats = chain; break;
        default: ctx->err = "reblur_debug_read: which must be 0..3"; return BPT_ERR_INVALID;
    }
    if (!src->p || capacity_floats < floats) { ctx->err = "reblur_debug_read: nothing rendered yet or capacity too small"; return BPT_ERR_INVALID; }
    BPT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    BPT_CUDA_TRY(ctx, cudaMemcpy(out, src->p, floats * 4, cudaMemcpyDeviceToHost));
    return BPT_OK;
}
