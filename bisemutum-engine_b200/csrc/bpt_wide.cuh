// bpt_wide.cuh — 4-wide BVH with 8-bit child boxes in 64-byte nodes, derived from the binary LBVH (merged mode).
//
// Why: from the second bounce on the traversal kernels are bound by the L1 pipeline (85-94 % of its peak,
// profiles/r1_v7_kernels.md): every lane fetches its own 64-B binary node (four 16-B loads) ~36 times per ray. A 4-wide
// node of the same 64 B replaces ~1.9 binary steps (oracle/oracle_wide.cpp measured 19.2 vs 36.1 nodes per diffuse ray),
// so node bytes, tag look-ups and stack traffic per ray roughly halve; the instruction count stays about the same.
//
// Results do not change: the wide tree only PROPOSES leaves. A proposed triangle is tested only if the ray also passes
// the EXACT box of that leaf (the box the binary tree tests in the leaf's parent), with the same slab formula. Slab
// values are monotone in the plane coordinates, so passing the exact leaf box implies passing every exact ancestor box
// of the binary tree — the set of triangles that reach the triangle test is exactly the binary traversal's (the
// quantised boxes are supersets by construction, see collapse_node4), and the result rule of bpt_trace.cuh makes the
// answer independent of the visiting order. tests/: hits bit-equal to the binary traversal of the CPU restatement.
//
// Node i of the wide tree describes the subtree of binary node i (same index: no allocation, every binary node gets a
// wide node, only those reachable from the root are ever read). Layout (4 x 16 B):
//   q0 = (origin.x, origin.y, origin.z, bits: ex | ey << 8 | ez << 16)     scale_a = 2^(e_a - 127)
//   q1 = child[4]      >= 0: binary/wide node index, < 0: ~leaf (position in the sorted triangle array), kNoChild: empty
//   q2 = (lo.x[4], lo.y[4], lo.z[4], hi.x[4])   one byte per child, child k in bits 8k..8k+7
//   q3 = (hi.y[4], hi.z[4], -, -)
// decoded plane = fmaf(float(q), scale, origin): the build checks conservativeness with this very expression.
#pragma once
#include "bpt_trace.cuh"

namespace bptd {

constexpr int32_t kNoChild = 0x7ffffffe;

struct Box3 { float3 lo, hi; };
BPT_HD float box_area(const Box3& b) { float3 d = b.hi - b.lo; return (d.x * d.y + d.y * d.z) + d.z * d.x; }
BPT_HD void binary_children(const float4* nodes2, int32_t i, Box3 b[2], int32_t c[2]) {
    const float4* n = nodes2 + 4 * (size_t)i;
    float4 n0 = n[0], n1 = n[1], n2 = n[2], n3 = n[3];
    b[0].lo = v3(n0.x, n0.z, n2.x); b[0].hi = v3(n0.y, n0.w, n2.y);
    b[1].lo = v3(n1.x, n1.z, n2.z); b[1].hi = v3(n1.y, n1.w, n2.w);
    c[0] = (int32_t)f2u(n3.x); c[1] = (int32_t)f2u(n3.y);
}
BPT_HD float axis(const float3& v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// Smallest e with 255 * 2^e >= ext (e >= -126), from the exponent field of ext.
BPT_HD int quant_exponent(float ext) {
    uint32_t eb = (f2u(ext) >> 23) & 0xffu;
    int E = eb == 0 ? -126 : (int)eb - 127;
    int e = E - 7;
    if (e < -126) e = -126;
    if (255.0f * u2f((uint32_t)(e + 127) << 23) < ext) e += 1;
    return e;
}

// Collapse rule (same in oracle/oracle_wide.cpp): start with the two children of binary node i; while fewer than 4,
// replace the INTERNAL child with the largest box area (ties: lowest slot) by its two children (child0 in place,
// child1 appended). Then quantise each axis on a power-of-two grid anchored at the node's lo corner.
BPT_HD void collapse_node4(const float4* nodes2, int32_t i, float4 out[4]) {
    Box3 b[4]; int32_t c[4];
    int n = 2;
    binary_children(nodes2, i, b, c);
    while (n < 4) {
        int best = -1; float ba = -1.0f;
        for (int k = 0; k < n; k++)
            if (c[k] >= 0) { float a = box_area(b[k]); if (a > ba) { ba = a; best = k; } }
        if (best < 0) break;
        Box3 cb[2]; int32_t cc[2];
        binary_children(nodes2, c[best], cb, cc);
        c[best] = cc[0]; b[best] = cb[0];
        c[n] = cc[1]; b[n] = cb[1];
        n++;
    }
    float3 lo = b[0].lo, hi = b[0].hi;
    for (int k = 1; k < n; k++) { lo = vmin(lo, b[k].lo); hi = vmax(hi, b[k].hi); }
    uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0}, ebits = 0;
    for (int a = 0; a < 3; a++) {
        const float o = axis(lo, a);
        int e = quant_exponent(axis(hi, a) - o);
        for (;;) {                                         // (a second pass only if rounding of hi - lo left the far plane short)
            const float scale = u2f((uint32_t)(e + 127) << 23);
            bool ok = true;
            uint32_t pl = 0, ph = 0;
            for (int k = 0; k < 4; k++) {
                float ql = 255.0f, qh = 0.0f;              // empty slot: inverted box, never hit
                if (k < n) {
                    const float cl = axis(b[k].lo, a), ch = axis(b[k].hi, a);
                    ql = floorf((cl - o) / scale); qh = ceilf((ch - o) / scale);
                    ql = tmin_(tmax_(ql, 0.0f), 255.0f); qh = tmin_(tmax_(qh, 0.0f), 255.0f);
                    while (ql > 0.0f && fmaf(ql, scale, o) > cl) ql -= 1.0f;
                    while (qh < 255.0f && fmaf(qh, scale, o) < ch) qh += 1.0f;
                    if (fmaf(ql, scale, o) > cl || fmaf(qh, scale, o) < ch) ok = false;
                }
                pl |= (uint32_t)ql << (8 * k); ph |= (uint32_t)qh << (8 * k);
            }
            if (ok || e >= 127) { qlo[a] = pl; qhi[a] = ph; break; }
            e += 1;
        }
        ebits |= (uint32_t)(e + 127) << (8 * a);
    }
    out[0] = make_float4(lo.x, lo.y, lo.z, u2f(ebits));
    out[1] = make_float4(u2f((uint32_t)c[0]), u2f((uint32_t)c[1]), u2f((uint32_t)(n > 2 ? c[2] : kNoChild)), u2f((uint32_t)(n > 3 ? c[3] : kNoChild)));
    out[2] = make_float4(u2f(qlo[0]), u2f(qlo[1]), u2f(qlo[2]), u2f(qhi[0]));
    out[3] = make_float4(u2f(qhi[1]), u2f(qhi[2]), 0.0f, 0.0f);
}

// Exact box of every leaf, in sorted-triangle order: leafbox[2j] = (lo | -), leafbox[2j+1] = (hi | -); taken from the
// binary parent's child-box fields (bit-identical to what the binary traversal tests).
BPT_HD void leaf_boxes_of_node(const float4* nodes2, int32_t i, float4* leafbox) {
    Box3 b[2]; int32_t c[2];
    binary_children(nodes2, i, b, c);
    for (int k = 0; k < 2; k++)
        if (c[k] < 0) {
            size_t j = (size_t)(uint32_t)~c[k];
            leafbox[2 * j] = make_float4(b[k].lo.x, b[k].lo.y, b[k].lo.z, 0.0f);
            leafbox[2 * j + 1] = make_float4(b[k].hi.x, b[k].hi.y, b[k].hi.z, 0.0f);
        }
}
BPT_HD bool leaf_box_hit_rec(float4 lo, float4 hi, const RaySpace& r, float tmin, float tcull);
BPT_HD bool leaf_box_hit(const float4* leafbox, uint32_t j, const RaySpace& r, float tmin, float tcull) {
    float4 lo, hi;
    ldg_32B(leafbox + 2 * (size_t)j, lo, hi);
    return leaf_box_hit_rec(lo, hi, r, tmin, tcull);
}
BPT_HD bool leaf_box_hit_rec(float4 lo, float4 hi, const RaySpace& r, float tmin, float tcull) {
    float lx = fmaf(lo.x, r.idir.x, -r.ood.x), hx = fmaf(hi.x, r.idir.x, -r.ood.x);
    float ly = fmaf(lo.y, r.idir.y, -r.ood.y), hy = fmaf(hi.y, r.idir.y, -r.ood.y);
    float lz = fmaf(lo.z, r.idir.z, -r.ood.z), hz = fmaf(hi.z, r.idir.z, -r.ood.z);
    float tn = fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fmaxf(fminf(lz, hz), tmin));
    float tf = fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fminf(fmaxf(lz, hz), tcull));
    return tn <= tf;
}

// byte K of `w` as a float: one I2F.U8 with a byte selector. (Measured against PRMT(2^23 | q) - 2^23 on the full-rate pipes:
// the plain cast is 5 % faster on configs[1].)
template <int K>
BPT_HD float byte_to_float(uint32_t w) { return (float)((w >> (8 * K)) & 0xffu); }

// One wide step: the four child references, a bit mask of the children whose decoded box the ray enters, and the slot
// of the nearest of them (ties: lowest slot; -1 when none is hit).
template <bool W256 = true>
BPT_HD void node_test4q(const float4* wide, int32_t cur, const RaySpace& r, float tmin, float tcull, int32_t ch[4], uint32_t& hitmask, int& best) {
    const float4* n = wide + 4 * (size_t)cur;
    float4 q0, q1, q2, q3;
    ldg_64B<W256>(n, q0, q1, q2, q3);
    const uint32_t eb = f2u(q0.w);
    const float sx = u2f((eb & 0xffu) << 23), sy = u2f(((eb >> 8) & 0xffu) << 23), sz = u2f(((eb >> 16) & 0xffu) << 23);
    const uint32_t lox = f2u(q2.x), loy = f2u(q2.y), loz = f2u(q2.z), hix = f2u(q2.w), hiy = f2u(q3.x), hiz = f2u(q3.y);
    ch[0] = (int32_t)f2u(q1.x); ch[1] = (int32_t)f2u(q1.y); ch[2] = (int32_t)f2u(q1.z); ch[3] = (int32_t)f2u(q1.w);
    // Near / far plane words by ray octant: for a valid box (lo <= hi) the slab values are monotone in the plane, so
    // min(t(lo), t(hi)) = t(near plane) bit for bit — one select per axis per NODE instead of two min/max per axis per child.
    const bool nx = r.idir.x < 0.0f, ny = r.idir.y < 0.0f, nz = r.idir.z < 0.0f;
    const uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
    const uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
    const uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
    hitmask = 0; best = -1;
    float tb = 0.0f;
#define BPT_WIDE_CHILD(K)                                                                                                   \
    {                                                                                                                       \
        float pnx = fmaf(byte_to_float<K>(nearx), sx, q0.x), pfx = fmaf(byte_to_float<K>(farx), sx, q0.x);                  \
        float pny = fmaf(byte_to_float<K>(neary), sy, q0.y), pfy = fmaf(byte_to_float<K>(fary), sy, q0.y);                  \
        float pnz = fmaf(byte_to_float<K>(nearz), sz, q0.z), pfz = fmaf(byte_to_float<K>(farz), sz, q0.z);                  \
        float tnx = fmaf(pnx, r.idir.x, -r.ood.x), tfx = fmaf(pfx, r.idir.x, -r.ood.x);                                     \
        float tny = fmaf(pny, r.idir.y, -r.ood.y), tfy = fmaf(pfy, r.idir.y, -r.ood.y);                                     \
        float tnz = fmaf(pnz, r.idir.z, -r.ood.z), tfz = fmaf(pfz, r.idir.z, -r.ood.z);                                     \
        float t0 = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                                                \
        float t1 = fminf(fminf(tfx, tfy), fminf(tfz, tcull));                                                               \
        const bool h = t0 <= t1 && ch[K] != kNoChild;   /* an empty slot decodes to an inverted box; the child id rejects it */ \
        const bool nearer = h && (best < 0 || t0 < tb);                                                                     \
        hitmask |= h ? (1u << K) : 0u;                                                                                      \
        best = nearer ? K : best;                                                                                           \
        tb = nearer ? t0 : tb;                                                                                              \
    }
    BPT_WIDE_CHILD(0) BPT_WIDE_CHILD(1) BPT_WIDE_CHILD(2) BPT_WIDE_CHILD(3)
#undef BPT_WIDE_CHILD
}
// The same step in "next node + nodes to push" form (run-to-completion traversal): the others in slot order.
BPT_HD int32_t node_step4q(const float4* wide, int32_t cur, const RaySpace& r, float tmin, float tcull, int32_t push[3], int& npush) {
    int32_t ch[4]; uint32_t hitmask; int best;
    node_test4q(wide, cur, r, tmin, tcull, ch, hitmask, best);
    npush = 0;
    if (best < 0) return BPT_POP;
    for (int k = 0; k < 4; k++)
        if (((hitmask >> k) & 1u) && k != best) push[npush++] = ch[k];
    return ch[best];
}

// Run-to-completion wide traversal of the merged BVH (ray batches, host-check): the same steps the persistent kernel interleaves.
template <bool ANY>
BPT_HD TraceResult trace_ray_wide(const DScene& sc, const float4* wide, const float4* leafbox, float3 O, float3 D, float tmin, float tmax,
                                  uint32_t frame_index, bool cull_non_opaque = false) {
    RayState rs;
    rs.O = O; rs.D = D; rs.tmin = tmin; rs.tbest = tmax; rs.tcull = tmax * 1.00001f; rs.best_slot = 0xffffffffu; rs.best_prim = 0xffffffffu;
    rs.bu = 0.0f; rs.bv = 0.0f; rs.frame_index = frame_index; rs.opacity_u = 0.0f; rs.have_u = false; rs.found = false; rs.cull_non_opaque = cull_non_opaque;
    const DBlas& bl = sc.blas[0];
    RaySpace sp_ = make_space(O, D);
    int32_t stack[kStackSize];
    int sp = 0;
    TraceResult res;
    if (bl.n != 0) {
        int32_t node = bl.root;                              // ~0 when the tree is a single leaf
        for (;;) {
            if (node >= 0) {
                int32_t push[3]; int np;
                int32_t next = node_step4q(wide, node, sp_, rs.tmin, rs.tcull, push, np);
                for (int k = 0; k < np; k++) stack[sp++] = push[k];
                if (next != BPT_POP) { node = next; continue; }
            } else {
                uint32_t j = (uint32_t)~node;
                if (bl.n == 1 || leaf_box_hit(leafbox, j, sp_, rs.tmin, rs.tcull)) {
                    bool accepted = test_triangle<ANY>(sc, rs, bl.tris + 3 * (size_t)j, sp_.O, sp_.D, 0xffffffffu, 0u);
                    if (ANY && accepted) break;
                }
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    res.hit = rs.found; res.t = rs.found ? rs.tbest : -1.0f; res.u = rs.bu; res.v = rs.bv; res.slot = rs.best_slot; res.prim = rs.best_prim;
    return res;
}

// Two-level counterpart (ray batches, host-check): wide TLAS whose proposed instances must pass their exact world box, then the
// instance's wide BLAS with the object-space ray. Same candidates as the binary two-level traversal of bpt_trace.cuh, hence the same hits.
template <bool ANY>
BPT_HD TraceResult trace_ray_wide_two_level(const DScene& sc, float3 O, float3 D, float tmin, float tmax, uint32_t frame_index, bool cull_non_opaque = false) {
    RayState rs;
    rs.O = O; rs.D = D; rs.tmin = tmin; rs.tbest = tmax; rs.tcull = tmax * 1.00001f; rs.best_slot = 0xffffffffu; rs.best_prim = 0xffffffffu;
    rs.bu = 0.0f; rs.bv = 0.0f; rs.frame_index = frame_index; rs.opacity_u = 0.0f; rs.have_u = false; rs.found = false; rs.cull_non_opaque = cull_non_opaque;
    const RaySpace world = make_space(O, D);
    RaySpace cur = world;
    const float4* nodes = sc.tlas_wide; const float4* tris = nullptr; const float4* leafbox = nullptr;
    uint32_t slot = 0xffffffffu, inst_anyhit = 0;
    bool in_blas = false;
    int32_t stack[2 * kStackSize];
    int sp = 0;
    TraceResult res;
    if (sc.tlas_n != 0) {
        int32_t node = sc.tlas_root;                         // ~0 when the TLAS is a single instance
        for (;;) {
            if (node == kSentinel) {
                cur = world; nodes = sc.tlas_wide; tris = nullptr; in_blas = false;
            } else if (node >= 0) {
                int32_t push[3]; int np;
                int32_t next = node_step4q(nodes, node, cur, rs.tmin, rs.tcull, push, np);
                for (int k = 0; k < np; k++) stack[sp++] = push[k];
                if (next != BPT_POP) { node = next; continue; }
            } else if (!in_blas) {
                uint32_t j = (uint32_t)~node;
                if (sc.tlas_n == 1 || leaf_box_hit(sc.tlas_leafbox, j, cur, rs.tmin, rs.tcull)) {
                    slot = BPT_LDG(sc.tlas_prims + j);
                    const DInstance& in = sc.instances[slot];
                    const DBlas& bl = sc.blas[in.blas];
                    inst_anyhit = in.anyhit;
                    cur = make_space(xf_point(in.w2o, rs.O), xf_vector(in.w2o, rs.D));
                    nodes = bl.wide; tris = bl.tris; leafbox = bl.n == 1 ? nullptr : bl.leafbox; in_blas = true;
                    stack[sp++] = kSentinel;
                    node = bl.root;
                    continue;
                }
            } else {
                uint32_t j = (uint32_t)~node;
                if (!leafbox || leaf_box_hit(leafbox, j, cur, rs.tmin, rs.tcull)) {
                    bool accepted = test_triangle<ANY>(sc, rs, tris + 3 * (size_t)j, cur.O, cur.D, slot, inst_anyhit);
                    if (ANY && accepted) break;
                }
            }
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
    res.hit = rs.found; res.t = rs.found ? rs.tbest : -1.0f; res.u = rs.bu; res.v = rs.bv; res.slot = rs.best_slot; res.prim = rs.best_prim;
    return res;
}

} // namespace bptd
