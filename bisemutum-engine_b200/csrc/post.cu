// post.cu — bloom + output pass after the path tracer (PostProcessPass::render, post_process.cpp:92-273) as 6 launches
// instead of the reference's 11 full-screen passes; see bpt_post.cuh for the per-pixel functions and the fusion plan.
// HBM-bound image work: intermediates are stored as real halves (8 B / texel), the level-1 kernel reads the FP32 sum buffer
// once (plus halo) and never materialises the resolved image or the bloom_pre target.
#include <algorithm>
#include <cuda_fp16.h>
#include "bpt_internal.cuh"
#include "bpt_post.cuh"

using namespace bptd;

namespace {

constexpr int kTX = 32, kTY = 16, kHalo = 5, kRows = kTY + 2 * kHalo, kPostThreads = 256;
// Source footprint of one tile (taps reach 3.24 destination texels = ~6.5 source texels to either side, the source is ~2x the
// destination): at most ceil((32 + 6.5) * 2.1) + 3 = 84 columns x ceil(26 * 2.1) + 3 = 58 rows for every size >= 10 texels.
constexpr int kSrcCap = 88 * 60;
// Both shared-memory images hold values that are half-representable by construction (an rgba16_sfloat target's content), so
// they are kept as packed halves: 8 B per texel, one LDS.64 per fetch, 47 KB per block -> 4 blocks per SM.
constexpr size_t kLevelSmem = (size_t)(kSrcCap + kRows * kTX) * sizeof(uint2);

// rgba16_sfloat texel <-> float3 (values are already half-representable: the conversions are exact)
__device__ __forceinline__ float3 unpack_h4(uint2 p) {
    __half2 a = *reinterpret_cast<__half2*>(&p.x), b = *reinterpret_cast<__half2*>(&p.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return v3(fa.x, fa.y, fb.x);
}
__device__ __forceinline__ uint2 pack_h4(float3 c) {
    __half2 a = __floats2half2_rn(c.x, c.y), b = __floats2half2_rn(c.z, 1.0f);
    uint2 p; p.x = *reinterpret_cast<uint32_t*>(&a); p.y = *reinterpret_cast<uint32_t*>(&b);
    return p;
}
// (raw() / conv() are split so that a batch of loads can be in flight before the first value is used)
struct TexHalf {                      // an rgba16_sfloat target of an earlier pass
    typedef uint2 Raw;
    const uint2* p; int w;
    __device__ Raw raw(int x, int y) const { return __ldg(p + (size_t)y * w + x); }
    __device__ float3 conv(Raw r) const { return unpack_h4(r); }
    __device__ float3 at(int x, int y) const { return conv(raw(x, y)); }
};
struct TexPre {                       // bloom_pre of the resolved colour, evaluated on fetch (never stored)
    typedef float4 Raw;
    const float4* accum; int w; float inv; BloomWeights bw;
    __device__ Raw raw(int x, int y) const { return __ldg(accum + (size_t)y * w + x); }
    __device__ float3 conv(Raw s) const { return bloom_pre(v3(s.x * inv, s.y * inv, s.z * inv), bw); }
    __device__ float3 at(int x, int y) const { return conv(raw(x, y)); }
};
struct TexStaged {                    // the tile's source footprint, fetched (and bloom_pre'd) once into shared memory
    const uint2* p; int x_lo, y_lo, fw;
    __device__ float3 at(int x, int y) const { return unpack_h4(p[(y - y_lo) * fw + (x - x_lo)]); }
};
struct TexTile {                      // horizontal-pass output of this block's tile (+ halo rows) in shared memory
    const uint2* p; int x0, ybase;
    __device__ float3 at(int x, int y) const { return unpack_h4(p[(y - ybase) * kTX + (x - x0)]); }
};

// One bloom level = bloom_horizontal_fs + bloom_vertical_fs of the reference in one launch. Per 32 x 16 tile:
//   1. the source texels the tile's horizontal pass can touch go to shared memory ONCE (for level 1 that is where the resolve
//      scale and bloom_pre run: one evaluation per source texel instead of one per bilinear fetch),
//   2. the horizontal pass of the tile's rows + 5 halo rows goes to shared memory (values as the rgba16_sfloat target holds them),
//   3. the vertical pass reads it from there and writes the level's only global output (8 B / texel).
// The footprint is computed with the very coordinate function the taps use (monotone in x and in the offset), so it is exact;
// a footprint that does not fit (degenerate aspect ratios only) falls back to fetching from global memory.
template <class Src>
__global__ void __launch_bounds__(kPostThreads, 4) k_bloom_level(Src src, int sw, int sh, uint2* __restrict__ dst, int dw, int dh) {
    extern __shared__ uint2 smem[];
    uint2* s_src = smem;                                   // kSrcCap texels
    uint2* s_h = smem + kSrcCap;                           // kRows * kTX texels
    const int x0 = blockIdx.x * kTX, y0 = blockIdx.y * kTY, ybase = y0 - kHalo;
    const int x_last = min(x0 + kTX, dw) - 1, row_lo = max(ybase, 0), row_hi = min(y0 + kTY + kHalo, dh) - 1;
    const float offsets[5] = BPT_BLOOM_OFFSETS;
    const float tx = 1.0f / (float)dw;
    const int sx_lo = lin_coord(((float)x0 + 0.5f) / (float)dw + offsets[0] * tx, sw).i0;
    const int sx_hi = lin_coord(((float)x_last + 0.5f) / (float)dw + offsets[4] * tx, sw).i1;
    const int sy_lo = sh == dh ? row_lo : lin_coord(((float)row_lo + 0.5f) / (float)dh, sh).i0;
    const int sy_hi = sh == dh ? row_hi : lin_coord(((float)row_hi + 0.5f) / (float)dh, sh).i1;
    const int fw = sx_hi - sx_lo + 1, fh = sy_hi - sy_lo + 1;
    const bool staged = fw * fh <= kSrcCap;                // block-uniform
    if (staged) {
        constexpr int kBatch = 8;                          // loads in flight per thread (the first version waited for each one: latency-bound)
        const int n = fw * fh;
        const float inv_fw = 1.0f / (float)fw;             // i / fw for i < 2^13: (i + 0.5) * (1 / fw) is never within 1e-3 of an integer
        for (int base = threadIdx.x; base < n; base += kPostThreads * kBatch) {
            typename Src::Raw raw[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; k++) {
                int i = base + k * kPostThreads;
                if (i < n) { int y = (int)(((float)i + 0.5f) * inv_fw); raw[k] = src.raw(sx_lo + (i - y * fw), sy_lo + y); }
            }
#pragma unroll
            for (int k = 0; k < kBatch; k++) {
                int i = base + k * kPostThreads;
                if (i < n) s_src[i] = pack_h4(src.conv(raw[k]));
            }
        }
        __syncthreads();
    }
    const TexStaged st{s_src, sx_lo, sy_lo, fw};
    for (int i = threadIdx.x; i < kRows * kTX; i += kPostThreads) {
        int x = x0 + (i % kTX), row = ybase + i / kTX;
        float3 h = v3(0.0f, 0.0f, 0.0f);
        if (x < dw && row >= row_lo && row <= row_hi)      // (rows outside the target are never read: the vertical pass clamps)
            h = staged ? bloom_horizontal(st, sw, sh, x, row, dw, dh) : bloom_horizontal(src, sw, sh, x, row, dw, dh);
        s_h[i] = pack_h4(h);
    }
    __syncthreads();
    const TexTile tile{s_h, x0, ybase};
    for (int i = threadIdx.x; i < kTY * kTX; i += kPostThreads) {
        int x = x0 + (i % kTX), y = y0 + i / kTX;
        if (x < dw && y < dh) dst[(size_t)y * dw + x] = pack_h4(bloom_vertical(tile, x, y, dw, dh));
    }
}

// bloom_combine_fs between two bloom levels
__global__ void k_bloom_combine(const uint2* __restrict__ in1, TexHalf in2, int h2, uint2* __restrict__ dst, int dw, int dh) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    size_t p = (size_t)y * dw + x;
    dst[p] = pack_h4(bloom_combine(unpack_h4(__ldg(in1 + p)), in2, in2.w, h2, x, y, dw, dh));
}

// "Bloom Final Combine Pass" + "Post Process Pass": colour + upsampled bloom -> rgba16_sfloat -> (xyz, 1); bloom off: (colour.xyz, 1)
__global__ void k_post_output(const float4* __restrict__ accum, float inv, TexHalf bloom, int hb, int bloom_on, float4* __restrict__ out, int w, int h) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    size_t p = (size_t)y * w + x;
    float4 s = accum[p];
    float3 c = v3(s.x * inv, s.y * inv, s.z * inv);
    if (bloom_on) c = bloom_combine(c, bloom, bloom.w, hb, x, y, w, h);
    out[p] = make_float4(c.x, c.y, c.z, 1.0f);
}

} // namespace

#define LAUNCH2(ctx, kernel, grid, block, smem_bytes, ...)                                  \
    do {                                                                        \
        kernel<<<(grid), (block), (smem_bytes), (ctx)->stream>>>(__VA_ARGS__);  \
        (ctx)->launches++;                                                      \
        cudaError_t le__ = cudaGetLastError();                                  \
        if (le__ != cudaSuccess) { (ctx)->err = std::string("launch of " #kernel ": ") + cudaGetErrorString(le__); return BPT_ERR_CUDA; } \
    } while (0)

bpt_status launch_post_process(bpt_context* ctx, const bpt_post_settings& st, uint32_t total_samples, float* d_out) {
    const int W = (int)ctx->width, H = (int)ctx->height;
    const float inv = ctx->wf.accum_fp16 ? 1.0f : 1.0f / (float)total_samples;       // as launch_resolve
    const float4* accum = ctx->wf.accum.as<float4>();
    const dim3 blk2(32, 8);
    auto grid2 = [&](int w, int h) { return dim3((unsigned)((w + 31) / 32), (unsigned)((h + 7) / 8)); };
    TexHalf bloom{nullptr, 1}; int bloom_h = 1;
    if (st.bloom) {
        int lw[3], lh[3];
        size_t off[5], total = 0;                                                   // V1, V2, V3, C2 (W>>2), C1 (W>>1)
        for (int i = 0; i < 3; i++) { lw[i] = std::max(W >> (i + 1), 1); lh[i] = std::max(H >> (i + 1), 1); }
        const int idx_w[5] = {lw[0], lw[1], lw[2], lw[1], lw[0]}, idx_h[5] = {lh[0], lh[1], lh[2], lh[1], lh[0]};
        for (int i = 0; i < 5; i++) { off[i] = total; total += ((size_t)idx_w[i] * idx_h[i] * 8 + 255) & ~(size_t)255; }
        bpt_status s = dev_reserve(ctx, ctx->d_post, total);
        if (s) return s;
        uint2* t[5];
        for (int i = 0; i < 5; i++) t[i] = reinterpret_cast<uint2*>(static_cast<char*>(ctx->d_post.p) + off[i]);
        auto gridt = [&](int w, int h) { return dim3((unsigned)((w + kTX - 1) / kTX), (unsigned)((h + kTY - 1) / kTY)); };
        const BloomWeights bw = bloom_weights(st.bloom_threshold, st.bloom_threshold_softness);
        if (!ctx->bloom_attr_set) {      // per context (= per device): the opt-in is a property of the function on that device
            BPT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_bloom_level<TexPre>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLevelSmem));
            BPT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_bloom_level<TexHalf>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLevelSmem));
            ctx->bloom_attr_set = true;
        }
        LAUNCH2(ctx, k_bloom_level<TexPre>, gridt(lw[0], lh[0]), kPostThreads, kLevelSmem, TexPre{accum, W, inv, bw}, W, H, t[0], lw[0], lh[0]);
        LAUNCH2(ctx, k_bloom_level<TexHalf>, gridt(lw[1], lh[1]), kPostThreads, kLevelSmem, TexHalf{t[0], lw[0]}, lw[0], lh[0], t[1], lw[1], lh[1]);
        LAUNCH2(ctx, k_bloom_level<TexHalf>, gridt(lw[2], lh[2]), kPostThreads, kLevelSmem, TexHalf{t[1], lw[1]}, lw[1], lh[1], t[2], lw[2], lh[2]);
        LAUNCH2(ctx, k_bloom_combine, grid2(lw[1], lh[1]), blk2, 0, t[1], TexHalf{t[2], lw[2]}, lh[2], t[3], lw[1], lh[1]);      // "Bloom Combine Pass #2"
        LAUNCH2(ctx, k_bloom_combine, grid2(lw[0], lh[0]), blk2, 0, t[0], TexHalf{t[3], lw[1]}, lh[1], t[4], lw[0], lh[0]);      // "Bloom Combine Pass #1"
        bloom = TexHalf{t[4], lw[0]}; bloom_h = lh[0];
    }
    LAUNCH2(ctx, k_post_output, grid2(W, H), blk2, 0, accum, inv, bloom, bloom_h, st.bloom ? 1 : 0, reinterpret_cast<float4*>(d_out), W, H);
    return BPT_OK;
}
