// bpt_post.cuh — the step after the path: bloom + output of PostProcessPass::render
// (bisemutum/src/renderer/pass/post_process.cpp:92-273, shaders/renderer/post_process/*.hlsl), per-pixel device functions.
//
// The reference runs 11 full-screen fragment passes over rgba16_sfloat targets (pre, 3 x (horizontal, vertical), 2 combines,
// final combine, output). Here (post.cu) they are 6 launches: one fused kernel per bloom level (the horizontal pass of a tile plus
// its halo goes to shared memory, the vertical pass reads it from there; level 1 also applies the resolve scale and bloom_pre
// while fetching), two combines, and one kernel for final combine + output pass. Every value the reference would store in an
// rgba16_sfloat target is rounded to half exactly there (q_half), so the image equals the pass-by-pass restatement of
// oracle/oracle_post.cpp bit for bit; the sampling contract (explicit FP32 bilinear, texel fetch along same-size axes) is
// stated in that file's header. Texel accessors are template parameters so tests/hostcheck can run the same functions on the host.
#pragma once
#include "bpt_math.cuh"

namespace bptd {

struct BloomWeights { float x, y, z, w; };
// post_process.cpp:118-124
BPT_HD BloomWeights bloom_weights(float threshold, float softness) {
    float soft_threshold = softness * (threshold * 0.9f + 0.1f);
    BloomWeights b;
    b.x = threshold;
    b.y = threshold * soft_threshold;
    b.z = 2.0f * b.y;
    b.w = 0.25f / (b.y + 0.00001f);
    b.y -= threshold;
    return b;
}
// bloom_pre.hlsl:7-14 on one texel; the result as an rgba16_sfloat target holds it
BPT_HD float3 bloom_pre(float3 c, const BloomWeights& bw) {
    float lum = (c.x * 0.212671f + c.y * 0.715160f) + c.z * 0.072169f;          // core/utils/color.hlsl:3-5
    float soft = lum + bw.y;
    soft = fminf(fmaxf(soft, 0.0f), bw.x);
    soft = soft * soft * bw.w;
    float weight = fmaxf(soft, lum - bw.x) / fmaxf(lum, 0.0001f);
    return q_half3(v3(c.x * weight, c.y * weight, c.z * weight));
}

struct LinCoord { int i0, i1; float f; bool exact; };
// linear / clamp_to_edge along one axis of n texels
BPT_HD LinCoord lin_coord(float u, int n) {
    float x = u * (float)n - 0.5f, xf = floorf(x);
    LinCoord c; c.f = x - xf; c.exact = false;
    int i = (int)xf;
    c.i0 = i < 0 ? 0 : (i >= n ? n - 1 : i);
    c.i1 = i + 1 < 0 ? 0 : (i + 1 >= n ? n - 1 : i + 1);
    return c;
}
BPT_HD LinCoord exact_coord(int i) { LinCoord c; c.i0 = i; c.i1 = i; c.f = 0.0f; c.exact = true; return c; }   // same-size axis: no interpolation
// (mix3(a, b, f) = a + (b - a) * f per component: bpt_math.cuh)
template <class Tex>
BPT_HD float3 sample_lin(const Tex& t, const LinCoord& cx, const LinCoord& cy) {
    float3 top = cx.exact ? t.at(cx.i0, cy.i0) : mix3(t.at(cx.i0, cy.i0), t.at(cx.i1, cy.i0), cx.f);
    if (cy.exact) return top;
    float3 bot = cx.exact ? t.at(cx.i0, cy.i1) : mix3(t.at(cx.i0, cy.i1), t.at(cx.i1, cy.i1), cx.f);
    return mix3(top, bot, cy.f);
}

#define BPT_BLOOM_OFFSETS {-3.23076923f, -1.38461538f, 0.0f, 1.38461538f, 3.23076923f}        /* bloom_filter.hlsl:6-8 */
#define BPT_BLOOM_WEIGHTS {0.07027027f, 0.31621622f, 0.22702703f, 0.31621622f, 0.07027027f}    /* bloom_filter.hlsl:10-12 */

BPT_HD float3 bloom_finish(float3 sum) {              // NaN / Inf -> 0 (bloom_filter.hlsl:21,32), then the rgba16_sfloat store
    bool bad = !(fabsf(sum.x) <= 3.402823466e38f) || !(fabsf(sum.y) <= 3.402823466e38f) || !(fabsf(sum.z) <= 3.402823466e38f);
    return bad ? v3(0.0f, 0.0f, 0.0f) : q_half3(sum);
}
// bloom_horizontal_fs for destination texel (x, y) of a dw x dh target; `src` has sw x sh texels
template <class Tex>
BPT_HD float3 bloom_horizontal(const Tex& src, int sw, int sh, int x, int y, int dw, int dh) {
    const float offsets[5] = BPT_BLOOM_OFFSETS, weights[5] = BPT_BLOOM_WEIGHTS;
    const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh, tx = 1.0f / (float)dw;
    const LinCoord cy = sh == dh ? exact_coord(y) : lin_coord(v, sh);
    float3 sum = v3(0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        float off = offsets[i] * tx;
        float3 c = sample_lin(src, lin_coord(u + off, sw), cy);
        sum = v3(sum.x + c.x * weights[i], sum.y + c.y * weights[i], sum.z + c.z * weights[i]);
    }
    return bloom_finish(sum);
}
// bloom_vertical_fs for destination texel (x, y); `src` is the horizontal pass's output (same size: x is fetched at the texel).
// src.at(x, row) is only asked for rows within 5 of y (|offset| <= 3.24 texels + the bilinear neighbour).
template <class Tex>
BPT_HD float3 bloom_vertical(const Tex& src, int x, int y, int dw, int dh) {
    const float offsets[5] = BPT_BLOOM_OFFSETS, weights[5] = BPT_BLOOM_WEIGHTS;
    const float v = ((float)y + 0.5f) / (float)dh, ty = 1.0f / (float)dh;
    float3 sum = v3(0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < 5; i++) {
        float off = offsets[i] * ty;
        LinCoord cy = lin_coord(v + off, dh);
        float3 c = mix3(src.at(x, cy.i0), src.at(x, cy.i1), cy.f);
        sum = v3(sum.x + c.x * weights[i], sum.y + c.y * weights[i], sum.z + c.z * weights[i]);
    }
    return bloom_finish(sum);
}
// bloom_combine_fs for destination texel (x, y): c1 = input_color1 at the texel (same size), input_color2 (w2 x h2) sampled
template <class Tex>
BPT_HD float3 bloom_combine(float3 c1, const Tex& in2, int w2, int h2, int x, int y, int dw, int dh) {
    const float u = ((float)x + 0.5f) / (float)dw, v = ((float)y + 0.5f) / (float)dh;
    float3 c2 = sample_lin(in2, w2 == dw ? exact_coord(x) : lin_coord(u, w2), h2 == dh ? exact_coord(y) : lin_coord(v, h2));
    return q_half3(v3(c1.x + c2.x, c1.y + c2.y, c1.z + c2.z));
}

} // namespace bptd
