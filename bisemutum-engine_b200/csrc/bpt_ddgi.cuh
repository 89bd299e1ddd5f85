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

// ---- the consumer: calc_ddgi_volume_lighting (ddgi/ddgi_lighting.hlsl:7-83) for one volume --------------------------
// Generalised from the reference's fixed 8x8x8 probes / 6 / 14 texels to the volume's own counts and atlas sizes.
// Atlas fetches are explicit FP32 bilinear (the reference's linear sampler; the 1-texel border keeps both taps in the tile).
// Returns (irradiance rgb, 1), or 0 outside the volume.
BPT_HD float2 ddgi_oct_encode_01(float3 n) {                           // pack.hlsl:88-98
    float l1 = (fabsf(n.x) + fabsf(n.y)) + fabsf(n.z);
    float x = n.x / l1, y = n.y / l1, z = n.z / l1;
    if (!(z >= 0.0f)) {
        float wx = (1.0f - fabsf(y)) * (x >= 0.0f ? 1.0f : -1.0f), wy = (1.0f - fabsf(x)) * (y >= 0.0f ? 1.0f : -1.0f);
        x = wx; y = wy;
    }
    return make_float2(x * 0.5f + 0.5f, y * 0.5f + 0.5f);
}
// bilinear tap pair along one axis of a probe tile: texel-space coordinate c (0.5 = centre of texel 0)
BPT_HD void ddgi_taps(float c, uint32_t& i0, float& f) {
    float t = c - 0.5f;
    float fl = floorf(t);
    i0 = (uint32_t)(int32_t)fl;
    f = t - fl;
}
BPT_HD float4 ddgi_volume_lighting(const bpt_probe_volume& vol, uint32_t irr_size, uint32_t vis_size, const float4* irradiance, const float2* visibility,
                                   float3 pos, float3 normal, float3 view) {
    pos = (pos + normal * 0.2f) + view * 0.8f;                        // :16
    const float3 base = v3(vol.base_position[0], vol.base_position[1], vol.base_position[2]);
    const float3 fx = v3(vol.frame_x[0], vol.frame_x[1], vol.frame_x[2]), fy = v3(vol.frame_y[0], vol.frame_y[1], vol.frame_y[2]),
                 fz = v3(vol.frame_z[0], vol.frame_z[1], vol.frame_z[2]);
    float3 vec = pos - base;
    float x = dot3(vec, fx), y = dot3(vec, fy), z = dot3(vec, fz);
    if (x < 0.0f || x > vol.extent[0] || y < 0.0f || y > vol.extent[1] || z < 0.0f || z > vol.extent[2]) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const uint32_t nx = vol.probe_counts[0], ny = vol.probe_counts[1], nz = vol.probe_counts[2];
    const float mx = (float)(nx > 1 ? nx - 1 : 1), my = (float)(ny > 1 ? ny - 1 : 1), mz = (float)(nz > 1 ? nz - 1 : 1);
    float2 oct = ddgi_oct_encode_01(normal);
    // (oct * SIZE + 1) / (SIZE + 2) of the tile = texel coordinate oct * SIZE + 1 inside the (SIZE + 2)-texel tile (:32-33)
    float icx = oct.x * (float)irr_size + 1.0f, icy = oct.y * (float)irr_size + 1.0f;
    float vcx = oct.x * (float)vis_size + 1.0f, vcy = oct.y * (float)vis_size + 1.0f;
    float ifx_ = x * mx / vol.extent[0], ify_ = y * my / vol.extent[1], ifz_ = z * mz / vol.extent[2];     // :35-39
    if (!(vol.extent[0] > 0.0f)) ifx_ = 0.0f;
    if (!(vol.extent[1] > 0.0f)) ify_ = 0.0f;
    if (!(vol.extent[2] > 0.0f)) ifz_ = 0.0f;
    uint32_t ix = ftou(ifx_), iy = ftou(ify_), iz = ftou(ifz_);                                             // :40
    ix = ix < (nx > 1 ? nx - 2 : 0) ? ix : (nx > 1 ? nx - 2 : 0);
    iy = iy < (ny > 1 ? ny - 2 : 0) ? iy : (ny > 1 ? ny - 2 : 0);
    iz = iz < (nz > 1 ? nz - 2 : 0) ? iz : (nz > 1 ? nz - 2 : 0);
    ifx_ = ifx_ - (float)ix; ify_ = ify_ - (float)iy; ifz_ = ifz_ - (float)iz;
    const uint32_t istride = nx * ny * (irr_size + 2), vstride = nx * ny * (vis_size + 2);
    uint32_t ii0x, ii0y, vi0x, vi0y; float iffx, iffy, vffx, vffy;
    ddgi_taps(icx, ii0x, iffx); ddgi_taps(icy, ii0y, iffy); ddgi_taps(vcx, vi0x, vffx); ddgi_taps(vcy, vi0y, vffy);
    float3 sum = v3s(0.0f);
    float sum_weight = 0.0f;
    for (uint32_t i = 0; i < 8; i++) {                                                                       // :44-79
        uint32_t dx = i & 1u, dy = (i >> 1) & 1u, dz = i >> 2;
        float wx = mix1(1.0f - ifx_, ifx_, (float)dx), wy = mix1(1.0f - ify_, ify_, (float)dy), wz = mix1(1.0f - ifz_, ifz_, (float)dz);
        float w_probe = wx * wy * wz;
        uint32_t px = ix + dx, py = iy + dy, pz = iz + dz;
        px = px < nx ? px : nx - 1; py = py < ny ? py : ny - 1; pz = pz < nz ? pz : nz - 1;                  // (degenerate 1-probe axes)
        float3 probe_center = ((base + ((float)px * vol.extent[0] / mx) * fx) + ((float)py * vol.extent[1] / my) * fy) + ((float)pz * vol.extent[2] / mz) * fz;
        float3 to_probe = probe_center - pos;
        float3 dir = normalize3(to_probe);
        float w_dir = sq((dot3(dir, normal) + 1.0f) * 0.5f) + 0.2f;
        const uint32_t tile = py * nx + px;
        const float2* vt = visibility + ((size_t)pz * (vis_size + 2)) * vstride + (size_t)tile * (vis_size + 2);
        float2 v00 = vt[(size_t)vi0y * vstride + vi0x], v10 = vt[(size_t)vi0y * vstride + vi0x + 1];
        float2 v01 = vt[(size_t)(vi0y + 1) * vstride + vi0x], v11 = vt[(size_t)(vi0y + 1) * vstride + vi0x + 1];
        float vis_x = mix1(mix1(v00.x, v10.x, vffx), mix1(v01.x, v11.x, vffx), vffy);
        float vis_y = mix1(mix1(v00.y, v10.y, vffx), mix1(v01.y, v11.y, vffx), vffy);
        float sigma2 = vis_y - sq(vis_x);
        float dist = sqrtf(dot3(to_probe, to_probe));
        float w_vis = sigma2 / (sigma2 + sq(tmax_(dist - vis_x, 0.0f)));
        const float4* it = irradiance + ((size_t)pz * (irr_size + 2)) * istride + (size_t)tile * (irr_size + 2);
        float4 a00 = it[(size_t)ii0y * istride + ii0x], a10 = it[(size_t)ii0y * istride + ii0x + 1];
        float4 a01 = it[(size_t)(ii0y + 1) * istride + ii0x], a11 = it[(size_t)(ii0y + 1) * istride + ii0x + 1];
        float3 irr = mix3(mix3(v3(a00.x, a00.y, a00.z), v3(a10.x, a10.y, a10.z), iffx), mix3(v3(a01.x, a01.y, a01.z), v3(a11.x, a11.y, a11.z), iffx), iffy);
        float w = (w_probe * w_dir) * ((w_vis * w_vis) * w_vis);
        sum = sum + irr * w;
        sum_weight += w;
    }
    float4 r = sum_weight == 0.0f ? make_float4(0.0f, 0.0f, 0.0f, 1.0f) : make_float4(sum.x / sum_weight, sum.y / sum_weight, sum.z / sum_weight, 1.0f);
    if (!(is_finite1(r.x) && is_finite1(r.y) && is_finite1(r.z))) r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);   // :82
    return r;
}

} // namespace bptd
