// bpt_ddgi.cuh — DDGI probe blending: cosine-weighted irradiance and cos^50-weighted
// visibility gathers of a probe's rays into octahedral texels, temporal blend in gamma-5 space, border copy.
// Reference: shaders/renderer/ddgi/probe_blend_irradiance.hlsl:11-80, probe_blend_visibility.hlsl:10-78,
// probe_blend_common.hlsl:3-50, core/utils/pack.hlsl:94-102 (oct_decode). `pow` is replaced by fixed-order
// forms (x^5, x^50 by multiplications, x^(1/5) by a fixed number of Newton steps) — numeric contract, DESIGN §3.
#pragma once
#include "bpt_scene.cuh"

namespace bptd {

BPT_HD float3 oct_decode_01(float fx, float fy) {                     // pack.hlsl:94-102
    float x = fx * 2.0f - 1.0f, y = fy * 2.0f - 1.0f;
    float3 n = v3(x, y, (1.0f - fabsf(x)) - fabsf(y));
    float t = clampf_(-n.z, 0.0f, 1.0f);
    n.x = n.x + (n.x >= 0.0f ? -t : t);
    n.y = n.y + (n.y >= 0.0f ? -t : t);
    return normalize3(n);
}
// x^(1/5) for x >= 0: exponent-scaled initial guess, then 6 Newton steps y <- (4y + x / y^4) / 5
BPT_HD float root5(float x) {
    if (!(x > 0.0f)) return 0.0f;
    int32_t i = (int32_t)f2u(x);
    float y = u2f((uint32_t)((i - 0x3f800000) / 5 + 0x3f800000));
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float y2 = y * y;
        float y4 = y2 * y2;
        y = (4.0f * y + x / y4) * 0.2f;
    }
    return y;
}
BPT_HD float pow5_(float x) { float x2 = x * x; return (x2 * x2) * x; }
BPT_HD float pow50_(float x) { float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8, x32 = x16 * x16; return (x32 * x16) * x2; }
BPT_HD float temporal_blend(float cur, float hist, float alpha) {     // probe_blend_irradiance.hlsl:56-66 (gamma 5)
    return pow5_(mix1(root5(cur), root5(hist), alpha));
}
// direction the blend kernels use for ray `r` of a probe: D on a miss, normalize(P - centre) on a hit (:49)
BPT_HD float3 blend_trace_dir(float3 O, float3 D, float t) {
    if (t < 0.0f) return D;
    float3 P = O + D * t;
    return normalize3(P - O);
}
BPT_HD void border_coord(uint32_t cx, uint32_t cy, uint32_t size, uint32_t& bx, uint32_t& by) {   // probe_blend_common.hlsl:3-26
    bx = cx; by = cy;
    if (cx == 1) {
        if (cy == 1) { bx = size + 1; by = size + 1; }
        else if (cy == size) { bx = size + 1; by = 0; }
        else { bx = 0; by = size + 1 - cy; }
    } else if (cx == size) {
        if (cy == 1) { bx = 0; by = size + 1; }
        else if (cy == size) { bx = 0; by = 0; }
        else { bx = size + 1; by = size + 1 - cy; }
    } else if (cy == 1) { bx = size + 1 - cx; by = 0; }
    else if (cy == size) { bx = size + 1 - cx; by = size + 1; }
}
BPT_HD bool corner_coords(uint32_t cx, uint32_t cy, uint32_t size, uint32_t c[4]) {               // probe_blend_common.hlsl:28-50
    if (cx == 1) {
        if (cy == 1) { c[0] = size; c[1] = 0; c[2] = 0; c[3] = size; return true; }
        if (cy == size) { c[0] = 0; c[1] = 1; c[2] = size; c[3] = size + 1; return true; }
    } else if (cx == size) {
        if (cy == 1) { c[0] = 1; c[1] = 0; c[2] = size + 1; c[3] = size; return true; }
        if (cy == size) { c[0] = size + 1; c[1] = 1; c[2] = 1; c[3] = size + 1; return true; }
    }
    return false;
}

// One octahedral texel (tx, ty) of one probe. `dirs`/`rad` hold the probe's rays (direction, radiance rgb + hit distance).
// VIS = false: irradiance rgb; VIS = true: (mean distance, mean squared distance) with cos^50 weights, misses count as 1e6.
template <bool VIS>
BPT_HD float3 blend_texel(uint32_t tx, uint32_t ty, uint32_t size, const float3* dirs, const float4* rad, uint32_t nrays) {
    float3 probe_dir = oct_decode_01(((float)tx + 0.5f) / (float)size, ((float)ty + 0.5f) / (float)size);
    float3 sum = v3s(0.0f);
    float weight_sum = 0.0f;
    for (uint32_t i = 0; i < nrays; i++) {
        float w = tmax_(dot3(probe_dir, dirs[i]), 0.0f);
        float4 r = rad[i];
        if (VIS) {
            w = pow50_(w);
            float dist = r.w < 0.0f ? 1e6f : r.w;
            sum.x += w * dist; sum.y += w * (dist * dist);
        } else {
            sum = sum + w * v3(r.x, r.y, r.z);
        }
        weight_sum += w;
    }
    return weight_sum == 0.0f ? v3s(0.0f) : sum / weight_sum;
}

} // namespace bptd
