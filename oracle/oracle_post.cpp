// oracle_post.cpp — TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the step after the path tracer: PostProcessPass::render
// (bisemutum/src/renderer/pass/post_process.cpp:92-273; called from basic.cpp:228-231), pass by pass:
//   bloom_pre_fs            shaders/renderer/post_process/bloom_pre.hlsl:7-14     (weights: post_process.cpp:118-124)
//   bloom_horizontal_fs /   shaders/renderer/post_process/bloom_filter.hlsl:14-34 (3 iterations at W>>1, W>>2, W>>3;
//   bloom_vertical_fs                                                              texel_size = 1/dst size, post_process.cpp:147-194)
//   bloom_combine_fs        shaders/renderer/post_process/bloom_combine.hlsl:5-9  (chain post_process.cpp:203-250)
//   post_process_pass_fs    shaders/renderer/post_process/post_process.hlsl:7-10  (xyz, 1)
// Every intermediate render target is rgba16_sfloat (post_process.cpp:97-100,128-131,155-158,176-179,212-215): a store is
// modelled as IEEE round-to-nearest-even to half (store_half), like state_precision = reference_fp16 of the path tracer.
//
// Sampling contract (shared with csrc/bpt_post.cuh, which must reproduce these images bit for bit):
//   * a full-screen pass has texcoord = (pixel + 0.5) / destination size (screen_triangle.hlsl:3-13), FP32 division;
//   * the sampler is linear / clamp_to_edge (post_process.cpp:80-86); filtering is the explicit FP32 form
//     x = u * W - 0.5, i = floor(x), f = x - i, a + (b - a) * f, indices clamped — the texture unit's 8-bit weights are not modelled;
//   * along an axis on which source and destination have the SAME size the pass reads texel centres, where the filter
//     weight is 0 up to the rounding of (x + 0.5) / n * n - 0.5: that axis is fetched at the texel (no interpolation).
//     This applies to bloom_pre, to the vertical filter's x axis, to input_color1 of every combine and to the output pass.
//   * no FMA contraction; vector operations are per component in x, y, z order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>
#include "oracle.h"
#include "oracle_math.hpp"

namespace {

struct Image {
    uint32_t w = 0, h = 0;
    std::vector<float> px;       // w*h*3 (alpha is 1 in every pass)
    Image() = default;
    Image(uint32_t w_, uint32_t h_) : w(w_), h(h_), px((size_t)w_ * h_ * 3, 0.0f) {}
    const float* at(int x, int y) const { return &px[((size_t)y * w + x) * 3]; }
    float* at(int x, int y) { return &px[((size_t)y * w + x) * 3]; }
};

inline int clampi(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }
inline float store_half(float f) { return obpt_store_half(f); }

// SampleLevel(linear, clamp) at (u, v); `exact_x` / `exact_y` >= 0: that axis is fetched at the given texel (same-size axis)
void sample(const Image& t, float u, float v, int exact_x, int exact_y, float out[3]) {
    int x0, x1, y0, y1; float fx = 0.0f, fy = 0.0f;
    if (exact_x >= 0) { x0 = x1 = exact_x; }
    else { float x = u * (float)t.w - 0.5f, xf = std::floor(x); fx = x - xf; x0 = clampi((int)xf, (int)t.w); x1 = clampi((int)xf + 1, (int)t.w); }
    if (exact_y >= 0) { y0 = y1 = exact_y; }
    else { float y = v * (float)t.h - 0.5f, yf = std::floor(y); fy = y - yf; y0 = clampi((int)yf, (int)t.h); y1 = clampi((int)yf + 1, (int)t.h); }
    const float *a = t.at(x0, y0), *b = t.at(x1, y0), *c = t.at(x0, y1), *d = t.at(x1, y1);
    for (int k = 0; k < 3; k++) {
        float top = exact_x >= 0 ? a[k] : a[k] + (b[k] - a[k]) * fx;
        float bot = exact_x >= 0 ? c[k] : c[k] + (d[k] - c[k]) * fx;
        out[k] = exact_y >= 0 ? top : top + (bot - top) * fy;
    }
}

const float kOffsets[5] = {-3.23076923f, -1.38461538f, 0.0f, 1.38461538f, 3.23076923f};      // bloom_filter.hlsl:6-8
const float kWeights[5] = {0.07027027f, 0.31621622f, 0.22702703f, 0.31621622f, 0.07027027f};   // bloom_filter.hlsl:10-12

// bloom_horizontal_fs / bloom_vertical_fs into a dst_w x dst_h target
Image filter_pass(const Image& src, uint32_t dw, uint32_t dh, bool vertical) {
    Image dst(dw, dh);
    const float tx = 1.0f / (float)dw, ty = 1.0f / (float)dh;              // post_process.cpp:168,189
    for (uint32_t y = 0; y < dh; y++)
        for (uint32_t x = 0; x < dw; x++) {
            const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh;
            float sum[3] = {0.0f, 0.0f, 0.0f};
            for (int i = 0; i < 5; i++) {
                float c[3];
                if (!vertical) { float off = kOffsets[i] * tx; sample(src, u + off, v, -1, src.h == dh ? (int)y : -1, c); }
                else { float off = kOffsets[i] * ty; sample(src, u, v + off, src.w == dw ? (int)x : -1, -1, c); }
                for (int k = 0; k < 3; k++) sum[k] = sum[k] + c[k] * kWeights[i];
            }
            bool bad = false;
            for (int k = 0; k < 3; k++) bad = bad || std::isnan(sum[k]) || std::isinf(sum[k]);
            float* o = dst.at((int)x, (int)y);
            for (int k = 0; k < 3; k++) o[k] = store_half(bad ? 0.0f : sum[k]);
        }
    return dst;
}

// bloom_combine_fs: input_color1 has the destination's size (texel fetch), input_color2 is sampled
Image combine_pass(const Image& c1, const Image& c2) {
    Image dst(c1.w, c1.h);
    for (uint32_t y = 0; y < dst.h; y++)
        for (uint32_t x = 0; x < dst.w; x++) {
            const float u = ((float)x + 0.5f) / (float)dst.w, v = ((float)y + 0.5f) / (float)dst.h;
            float b[3];
            sample(c2, u, v, c2.w == dst.w ? (int)x : -1, c2.h == dst.h ? (int)y : -1, b);
            const float* a = c1.at((int)x, (int)y);
            float* o = dst.at((int)x, (int)y);
            for (int k = 0; k < 3; k++) o[k] = store_half(a[k] + b[k]);
        }
    return dst;
}

} // namespace

extern "C" bpt_status obpt_post_process_image(const float* in_rgba32f, uint32_t width, uint32_t height, const bpt_post_settings* st, float* out_rgba32f) {
    if (!in_rgba32f || !out_rgba32f || !st || !width || !height) return BPT_ERR_INVALID;
    const size_t npx = (size_t)width * height;
    Image color(width, height);
    for (size_t p = 0; p < npx; p++) for (int k = 0; k < 3; k++) color.px[p * 3 + k] = in_rgba32f[p * 4 + k];
    const Image* result = &color;
    Image after;
    if (st->bloom) {
        // post_process.cpp:118-124
        const float soft_threshold = st->bloom_threshold_softness * (st->bloom_threshold * 0.9f + 0.1f);
        float bw[4];
        bw[0] = st->bloom_threshold;
        bw[1] = st->bloom_threshold * soft_threshold;
        bw[2] = 2.0f * bw[1];
        bw[3] = 0.25f / (bw[1] + 0.00001f);
        bw[1] -= st->bloom_threshold;
        Image pre(width, height);                                           // bloom_pre.hlsl:7-14
        for (size_t p = 0; p < npx; p++) {
            const float* c = &color.px[p * 3];
            float lum = (c[0] * 0.212671f + c[1] * 0.715160f) + c[2] * 0.072169f;      // core/utils/color.hlsl:3-5
            float soft = lum + bw[1];
            soft = std::fmin(std::fmax(soft, 0.0f), bw[0]);
            soft = soft * soft * bw[3];
            float weight = std::fmax(soft, lum - bw[0]) / std::fmax(lum, 0.0001f);
            for (int k = 0; k < 3; k++) pre.px[p * 3 + k] = store_half(c[k] * weight);
        }
        std::vector<Image> temp;                                            // post_process.cpp:133-197
        for (uint32_t i = 0; i < 3; i++) {
            uint32_t dw = std::max(width >> (i + 1), 1u), dh = std::max(height >> (i + 1), 1u);
            temp.push_back(filter_pass(i == 0 ? pre : temp[2 * i - 1], dw, dh, false));
            temp.push_back(filter_pass(temp[2 * i], dw, dh, true));
        }
        for (uint32_t i = 2; i > 0; i--) {                                  // post_process.cpp:203-232 (bloom_num_iterations = 3)
            const Image& c1 = temp[2 * i - 1];
            const Image& c2 = temp[i == 2 ? 2 * i + 1 : 3 * 3 - i - 2];
            temp.push_back(combine_pass(c1, c2));                           // lands at index 3*3 - i - 1
        }
        after = combine_pass(color, temp.back());                           // "Bloom Final Combine Pass", post_process.cpp:234-250
        result = &after;
    }
    for (size_t p = 0; p < npx; p++) {                                      // post_process.hlsl:7-10
        for (int k = 0; k < 3; k++) out_rgba32f[p * 4 + k] = result->px[p * 3 + k];
        out_rgba32f[p * 4 + 3] = 1.0f;
    }
    return BPT_OK;
}
