// bpt_trace.cuh — stack-based software BVH traversal (B200 has no RT cores).
//
// Semantics kept from the reference's hardware path: closest hit in (TMin, TMax] with
// RAY_FLAG_NONE and no face culling (shaders/renderer/raytracing/rt_gbuffer.hlsl:17-25), the
// any-hit opacity rule for non-opaque instances (hits/rt_gbuffer_hit.hlsl:20-35) with ONE random
// number per ray seeded from the ray's bit pattern (hits/rt_gbuffer.hlsl:14-19).
//
// Determinism: a hit replaces the current best iff t < best || (t == best && id < best_id) with
// id = (instance slot << 32 | primitive), and boxes are culled against tcull = best * (1 + 1e-5) (the slab
// and triangle tests round differently by a few ulps; the margin guarantees that every candidate that can
// still win is visited), so the result does not depend on traversal order — kernels may reorder work.
//
// Node fetch = four 16-B loads (64 B); triangle fetch = three 16-B loads (48 B).
#pragma once
#include "bpt_scene.cuh"

namespace bptd {

// Karras depth bound: 63 code bits + 32 index bits = 95 levels. The binary traversal holds at most one entry per level; a step of the
// 4-wide tree pushes up to three nodes and may descend a single binary level, hence 3 x 95 + 1 (only touched entries cost anything).
constexpr int kStackSize = 288;
constexpr int32_t kSentinel = (int32_t)0x80000000;

struct TraceResult {
    float t, u, v;
    uint32_t slot, prim;
    bool hit;
};

struct RayState {
    float3 O, D;          // world-space ray (opacity seed)
    float tmin, tbest;
    float tcull;          // tbest * (1 + 1e-5): what boxes are culled against (order-independence, see header)
    uint32_t best_slot, best_prim;
    float bu, bv;
    uint32_t frame_index;
    float opacity_u;
    bool have_u, found;
    bool cull_non_opaque; // RAY_FLAG_CULL_NON_OPAQUE (RTAO rays): triangles that would run the any-hit rule are skipped instead
};

BPT_HD float ray_opacity_random(RayState& rs) {                 // hits/rt_gbuffer.hlsl:14-19
    if (!rs.have_u) {
        uint32_t seed = f2u(rs.O.x) ^ f2u(rs.O.y) ^ f2u(rs.O.z) ^ f2u(rs.D.x) ^ f2u(rs.D.y) ^ f2u(rs.D.z);
        uint32_t st = rng_tea(seed, rs.frame_index);
        rs.opacity_u = rng_next(st);
        rs.have_u = true;
    }
    return rs.opacity_u;
}

// true = keep the candidate (hits/rt_gbuffer_hit.hlsl:20-35)
// Only reached for instances whose DInstance::anyhit flag is set (non-opaque instance AND non-opaque blend).
BPT_HD bool anyhit_keep(const DScene& sc, RayState& rs, uint32_t slot, uint32_t prim, float u, float v) {
    const DInstance& in = sc.instances[slot];
    const bpt_drawable_sbt_data& dr = sc.drawables[in.instance_id];
    const bpt_material& m = sc.materials[dr.material_offset / (uint32_t)sizeof(bpt_material)];
    uint32_t blend = (m.flags >> BPT_MATERIAL_BLEND_SHIFT) & 0xffu;
    float opacity = eval_hit_opacity(sc, in.instance_id, prim, u, v);
    if (blend == BPT_BLEND_ALPHA_TEST) return !(opacity < 0.01f);
    return !(ray_opacity_random(rs) < 1.0f - opacity);
}

// Möller–Trumbore on the (v0, e1, e2) record; unfused arithmetic (see bpt_math.cuh header).
// AH = false: the scene has no instance that needs the any-hit rule (every material opaque or every instance force-opaque), so the
// per-triangle flag is not even looked at and the opacity evaluation (material + texture fetches) is not compiled into the kernel.
template <bool ANY, bool AH = true>
BPT_HD bool test_triangle_rec(const DScene& sc, RayState& rs, float4 a, float4 b, float4 c, float3 O, float3 D, uint32_t slot_or_none, uint32_t instance_anyhit);
template <bool ANY>
BPT_HD bool test_triangle(const DScene& sc, RayState& rs, const float4* tri, float3 O, float3 D, uint32_t slot_or_none, uint32_t instance_anyhit) {
    float4 a = BPT_LDG(tri), b = BPT_LDG(tri + 1), c = BPT_LDG(tri + 2);
    return test_triangle_rec<ANY>(sc, rs, a, b, c, O, D, slot_or_none, instance_anyhit);
}
// the same test on an already fetched (v0 | prim), (e1 | slot), (e2 | any-hit flag) record
template <bool ANY, bool AH>
BPT_HD bool test_triangle_rec(const DScene& sc, RayState& rs, float4 a, float4 b, float4 c, float3 O, float3 D, uint32_t slot_or_none, uint32_t instance_anyhit) {
    float3 v0 = v3(a.x, a.y, a.z), e1 = v3(b.x, b.y, b.z), e2 = v3(c.x, c.y, c.z);
    float3 pvec = cross3(D, e2);
    float det = dot3(e1, pvec);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float3 tvec = O - v0;
    float u = dot3(tvec, pvec) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    float3 qvec = cross3(tvec, e1);
    float v = dot3(D, qvec) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    float t = dot3(e2, qvec) * inv;
    if (!(t > rs.tmin)) return false;
    uint32_t prim = f2u(a.w);
    uint32_t slot = slot_or_none == 0xffffffffu ? f2u(b.w) : slot_or_none;
    bool better = t < rs.tbest || (t == rs.tbest && (!rs.found || slot < rs.best_slot || (slot == rs.best_slot && prim < rs.best_prim)));
    if (!better) return false;
    // any-hit flag: per triangle (merged: e2.w) or per entered instance (two-level)
    if (AH) {
        uint32_t need_anyhit = slot_or_none == 0xffffffffu ? f2u(c.w) : instance_anyhit;
        if (need_anyhit && (rs.cull_non_opaque || !anyhit_keep(sc, rs, slot, prim, u, v))) return false;
    }
    rs.tbest = t; rs.tcull = t * 1.00001f; rs.bu = u; rs.bv = v; rs.best_slot = slot; rs.best_prim = prim; rs.found = true;
    return true;
}

struct RaySpace { float3 O, D, idir, ood; };
BPT_HD RaySpace make_space(float3 O, float3 D) {
    const float ooeps = 8.27180613e-25f;   // 2^-80
    RaySpace r; r.O = O; r.D = D;
    r.idir.x = 1.0f / (fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    r.idir.y = 1.0f / (fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    r.idir.z = 1.0f / (fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    r.ood = O * r.idir;
    return r;
}

// One node step: returns the next node to visit (or BPT_POP) and, when both children are hit, the farther one in `far`
// (BPT_POP otherwise) for the caller to push.
#define BPT_POP ((int32_t)0x80000001)
template <bool W256 = true>
BPT_HD int32_t node_step2(const float4* nodes, int32_t cur, const RaySpace& r, float tmin, float tcull, int32_t& far) {
    const float4* n = nodes + 4 * (size_t)cur;
    float4 n0, n1, n2, n3;
    ldg_64B<W256>(n, n0, n1, n2, n3);
    float c0lox = fmaf(n0.x, r.idir.x, -r.ood.x), c0hix = fmaf(n0.y, r.idir.x, -r.ood.x);
    float c0loy = fmaf(n0.z, r.idir.y, -r.ood.y), c0hiy = fmaf(n0.w, r.idir.y, -r.ood.y);
    float c1lox = fmaf(n1.x, r.idir.x, -r.ood.x), c1hix = fmaf(n1.y, r.idir.x, -r.ood.x);
    float c1loy = fmaf(n1.z, r.idir.y, -r.ood.y), c1hiy = fmaf(n1.w, r.idir.y, -r.ood.y);
    float c0loz = fmaf(n2.x, r.idir.z, -r.ood.z), c0hiz = fmaf(n2.y, r.idir.z, -r.ood.z);
    float c1loz = fmaf(n2.z, r.idir.z, -r.ood.z), c1hiz = fmaf(n2.w, r.idir.z, -r.ood.z);
    float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin));
    float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tcull));
    float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin));
    float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tcull));
    bool h0 = t0n <= t0f, h1 = t1n <= t1f;
    int32_t ch0 = (int32_t)f2u(n3.x), ch1 = (int32_t)f2u(n3.y);
    far = BPT_POP;
    if (h0 && h1) {
        bool c0_near = t0n <= t1n;
        far = c0_near ? ch1 : ch0;
        return c0_near ? ch0 : ch1;
    }
    if (h0) return ch0;
    if (h1) return ch1;
    return BPT_POP;
}
BPT_HD int32_t node_step(const float4* nodes, int32_t cur, const RaySpace& r, float tmin, float tcull, int32_t* stack, int& sp) {
    int32_t far;
    int32_t next = node_step2(nodes, cur, r, tmin, tcull, far);
    if (far != BPT_POP) stack[sp++] = far;
    return next;
}

// ---- resumable traversal state -------------------------------------------------------------------
// One ray's traversal as a small state machine so that the same steps serve (a) the simple
// run-to-completion loop (trace_ray: ray batches, host-check) and (b) the persistent while-while
// kernels of render.cu, which interleave steps of 32 rays per warp and refill finished lanes.
struct Trav {
    RayState rs;
    RaySpace world, cur;
    const float4* nodes;
    const float4* tris;
    uint32_t slot;          // instance slot of the BLAS being traversed; 0xffffffff = read it from the triangle (merged)
    uint32_t inst_anyhit;   // DInstance::anyhit of that instance (two-level)
    int32_t node;           // >= 0 internal node, < 0 leaf (~index)
    int sp;
    bool in_blas, done;
};

BPT_HD void trav_begin(const DScene& sc, Trav& t, float3 O, float3 D, float tmin, float tmax, uint32_t frame_index) {
    t.rs.O = O; t.rs.D = D; t.rs.tmin = tmin; t.rs.tbest = tmax; t.rs.tcull = tmax * 1.00001f; t.rs.best_slot = 0xffffffffu; t.rs.best_prim = 0xffffffffu;
    t.rs.bu = 0.0f; t.rs.bv = 0.0f; t.rs.frame_index = frame_index; t.rs.opacity_u = 0.0f; t.rs.have_u = false; t.rs.found = false; t.rs.cull_non_opaque = false;
    const bool two_level = sc.accel_mode == BPT_ACCEL_TWO_LEVEL;
    t.world = make_space(O, D);
    t.cur = t.world;
    t.nodes = two_level ? sc.tlas_nodes : sc.blas[0].nodes;
    t.tris = two_level ? nullptr : sc.blas[0].tris;
    t.slot = 0xffffffffu;
    t.inst_anyhit = 0;
    t.in_blas = !two_level;
    t.node = two_level ? sc.tlas_root : sc.blas[0].root;
    t.sp = 0;
    t.done = (two_level ? sc.tlas_n : sc.blas[0].n) == 0;
}
BPT_HD void trav_pop(const DScene& sc, Trav& t, int32_t* stack) {
    if (t.sp == 0) { t.done = true; return; }
    t.node = stack[--t.sp];
    if (t.node == kSentinel) {          // leaving an instance: back to the world-space ray and the TLAS
        t.cur = t.world; t.nodes = sc.tlas_nodes; t.tris = nullptr; t.in_blas = false;
        if (t.sp == 0) { t.done = true; return; }
        t.node = stack[--t.sp];
    }
}
BPT_HD void trav_node(const DScene& sc, Trav& t, int32_t* stack) {     // requires t.node >= 0
    int32_t next = node_step(t.nodes, t.node, t.cur, t.rs.tmin, t.rs.tcull, stack, t.sp);
    if (next != BPT_POP) t.node = next; else trav_pop(sc, t, stack);
}
template <bool ANY>
BPT_HD void trav_leaf(const DScene& sc, Trav& t, int32_t* stack) {     // requires t.node < 0
    if (t.in_blas) {
        bool accepted = test_triangle<ANY>(sc, t.rs, t.tris + 3 * (size_t)(uint32_t)~t.node, t.cur.O, t.cur.D, t.slot, t.inst_anyhit);
        if (ANY && accepted) { t.done = true; return; }
        trav_pop(sc, t, stack);
    } else {
        // TLAS leaf: enter the instance (object-space ray, direction NOT renormalised so t is shared)
        t.slot = BPT_LDG(sc.tlas_prims + (uint32_t)~t.node);
        const DInstance& in = sc.instances[t.slot];
        const DBlas& bl = sc.blas[in.blas];
        t.inst_anyhit = in.anyhit;
        t.cur = make_space(xf_point(in.w2o, t.rs.O), xf_vector(in.w2o, t.rs.D));
        t.nodes = bl.nodes; t.tris = bl.tris; t.in_blas = true;
        stack[t.sp++] = kSentinel;
        t.node = bl.root;
    }
}
BPT_HD TraceResult trav_result(const Trav& t) {
    TraceResult res;
    res.hit = t.rs.found; res.t = t.rs.found ? t.rs.tbest : -1.0f; res.u = t.rs.bu; res.v = t.rs.bv; res.slot = t.rs.best_slot; res.prim = t.rs.best_prim;
    return res;
}

template <bool ANY>
BPT_HD TraceResult trace_ray(const DScene& sc, float3 O, float3 D, float tmin, float tmax, uint32_t frame_index, bool cull_non_opaque = false) {
    Trav t;
    int32_t stack[kStackSize];
    trav_begin(sc, t, O, D, tmin, tmax, frame_index);
    t.rs.cull_non_opaque = cull_non_opaque;
    while (!t.done) {
        if (t.node >= 0) trav_node(sc, t, stack);
        else trav_leaf<ANY>(sc, t, stack);
    }
    return trav_result(t);
}

} // namespace bptd
