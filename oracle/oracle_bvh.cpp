// oracle_bvh.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// LBVH build + stack traversal. NEW code: the reference builds its BLAS/TLAS with the Vulkan /
// D3D12 driver (bisemutum/src/graphics/accel.cpp:11-159) and traverses with RT hardware
// (shaders/renderer/raytracing/rt_gbuffer.hlsl:17-25), so there is nothing to transliterate.
// What IS kept from the reference: one BLAS per (mesh, submesh) described exactly like
// graphics_manager.cpp:616-654, one TLAS instance per drawable with the 3x4 transform / id /
// opaque-flag semantics of accel.cpp:104-132, TMin/TMax/any-hit rules of rt_gbuffer.hlsl and
// hits/rt_gbuffer_hit.hlsl:20-35.
//
// Build definition (the GPU must reproduce it bit-exactly):
//   1. primitive AABB = min/max of its vertices; centroid c = (lo + hi) * 0.5f
//   2. bounds = min/max over primitive AABBs
//   3. per axis: t = ext > 0 ? (c - lo) / ext : 0 ; q = uint(min(max(t * 2^21, 0), 2^21 - 1));
//      code = interleave(qx, qy, qz) (x highest), 63 bits
//   4. stable sort of (code, primitive) by code (ties keep ascending primitive id)
//   5. Karras 2012 with delta(i,j) = code_i != code_j ? clz64(code_i ^ code_j) : 64 + clz32(i ^ j)
//   6. bottom-up union of child boxes (exact min/max)
#include <algorithm>
#include <cfloat>
#include <numeric>
#include "oracle_scene.hpp"

namespace orc {

static inline uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
static inline uint32_t quant21(float c, float lo, float hi) {
    float ext = hi - lo;
    float t = ext > 0.0f ? (c - lo) / ext : 0.0f;
    float s = t * 2097152.0f;
    s = fmax_(s, 0.0f);
    s = fmin_(s, 2097151.0f);
    return (uint32_t)s;
}
uint64_t morton63(f3 c, f3 lo, f3 hi) {
    return (expand21(quant21(c.x, lo.x, hi.x)) << 2) | (expand21(quant21(c.y, lo.y, hi.y)) << 1) |
           expand21(quant21(c.z, lo.z, hi.z));
}

static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }

void build_lbvh(Bvh& b, uint32_t n, const f3* plo, const f3* phi) {
    b = Bvh{};
    b.n = n;
    if (n == 0) return;
    f3 lo = plo[0], hi = phi[0];
    for (uint32_t i = 1; i < n; i++) { lo = min3(lo, plo[i]); hi = max3(hi, phi[i]); }
    b.lo = lo; b.hi = hi;
    std::vector<uint64_t> codes(n);
    for (uint32_t i = 0; i < n; i++) {
        f3 c = (plo[i] + phi[i]) * 0.5f;
        codes[i] = morton63(c, lo, hi);
    }
    b.prims.resize(n);
    std::iota(b.prims.begin(), b.prims.end(), 0u);
    std::stable_sort(b.prims.begin(), b.prims.end(), [&](uint32_t a, uint32_t c) { return codes[a] < codes[c]; });
    b.morton.resize(n);
    for (uint32_t i = 0; i < n; i++) b.morton[i] = codes[b.prims[i]];
    b.leaf_parent.assign(n, -1);
    if (n == 1) { b.root = ~0; return; }
    b.root = 0;
    b.nodes.assign(n - 1, bpt_bvh_node{});
    const uint64_t* k = b.morton.data();
    auto delta = [&](int64_t i, int64_t j) -> int {
        if (j < 0 || j >= (int64_t)n) return -1;
        uint64_t a = k[i], c = k[j];
        return a != c ? clz64(a ^ c) : 64 + clz32((uint32_t)i ^ (uint32_t)j);
    };
    b.nodes[0].parent = -1;
    for (int64_t i = 0; i < (int64_t)n - 1; i++) {
        int d = delta(i, i + 1) > delta(i, i - 1) ? 1 : -1;
        int dmin = delta(i, i - d);
        int64_t lmax = 2;
        while (delta(i, i + lmax * d) > dmin) lmax *= 2;
        int64_t l = 0;
        for (int64_t t = lmax / 2; t >= 1; t /= 2)
            if (delta(i, i + (l + t) * d) > dmin) l += t;
        int64_t j = i + l * d;
        int dnode = delta(i, j);
        int64_t s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        int64_t gamma = i + s * d + std::min(d, 0);
        int64_t lo_i = std::min(i, j), hi_i = std::max(i, j);
        int32_t left = (lo_i == gamma) ? ~(int32_t)gamma : (int32_t)gamma;
        int32_t right = (hi_i == gamma + 1) ? ~(int32_t)(gamma + 1) : (int32_t)(gamma + 1);
        b.nodes[i].child0 = left;
        b.nodes[i].child1 = right;
        if (left >= 0) b.nodes[left].parent = (int32_t)i; else b.leaf_parent[~left] = (int32_t)i;
        if (right >= 0) b.nodes[right].parent = (int32_t)i; else b.leaf_parent[~right] = (int32_t)i;
    }
    // refit: post-order with an explicit stack
    std::vector<f3> nlo(n - 1), nhi(n - 1);
    std::vector<uint8_t> state(n - 1, 0);
    std::vector<int32_t> stack;
    stack.push_back(0);
    auto child_box = [&](int32_t c, f3& clo, f3& chi) {
        if (c < 0) { uint32_t p = b.prims[~c]; clo = plo[p]; chi = phi[p]; }
        else { clo = nlo[c]; chi = nhi[c]; }
    };
    while (!stack.empty()) {
        int32_t i = stack.back();
        bpt_bvh_node& nd = b.nodes[i];
        if (state[i] == 0) {
            state[i] = 1;
            if (nd.child0 >= 0) stack.push_back(nd.child0);
            if (nd.child1 >= 0) stack.push_back(nd.child1);
        } else {
            stack.pop_back();
            f3 l0, h0, l1, h1;
            child_box(nd.child0, l0, h0);
            child_box(nd.child1, l1, h1);
            nd.c0_lo_x = l0.x; nd.c0_hi_x = h0.x; nd.c0_lo_y = l0.y; nd.c0_hi_y = h0.y;
            nd.c1_lo_x = l1.x; nd.c1_hi_x = h1.x; nd.c1_lo_y = l1.y; nd.c1_hi_y = h1.y;
            nd.c0_lo_z = l0.z; nd.c0_hi_z = h0.z; nd.c1_lo_z = l1.z; nd.c1_hi_z = h1.z;
            nd.reserved = 0;
            nlo[i] = min3(l0, l1);
            nhi[i] = max3(h0, h1);
        }
    }
}

// Inverse of a row-major 3x4 affine matrix: adjugate / determinant, then t' = -(A^-1 t).
void invert_3x4(const float m[12], float o[12]) {
    float a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[4], a11 = m[5], a12 = m[6], a20 = m[8], a21 = m[9], a22 = m[10];
    float c00 = a11 * a22 - a12 * a21;
    float c01 = a12 * a20 - a10 * a22;
    float c02 = a10 * a21 - a11 * a20;
    float det = (a00 * c00 + a01 * c01) + a02 * c02;
    float inv_det = 1.0f / det;
    o[0] = c00 * inv_det;
    o[1] = (a02 * a21 - a01 * a22) * inv_det;
    o[2] = (a01 * a12 - a02 * a11) * inv_det;
    o[4] = c01 * inv_det;
    o[5] = (a00 * a22 - a02 * a20) * inv_det;
    o[6] = (a02 * a10 - a00 * a12) * inv_det;
    o[8] = c02 * inv_det;
    o[9] = (a01 * a20 - a00 * a21) * inv_det;
    o[10] = (a00 * a11 - a01 * a10) * inv_det;
    float tx = m[3], ty = m[7], tz = m[11];
    o[3] = -((o[0] * tx + o[1] * ty) + o[2] * tz);
    o[7] = -((o[4] * tx + o[5] * ty) + o[6] * tz);
    o[11] = -((o[8] * tx + o[9] * ty) + o[10] * tz);
}

static void fetch_tri(const Scene& sc, const bpt_blas_desc& bd, uint32_t k, f3 v[3]) {
    for (int c = 0; c < 3; c++) {
        uint32_t idx = sc.indices[(size_t)bd.index_offset + 3ull * k + c];
        const float* p = &sc.positions[(size_t)bd.position_offset + 3ull * idx];
        v[c] = mk3(p[0], p[1], p[2]);
    }
}

static bool derive_instances(Scene& sc, std::string& err) {
    sc.xf.resize(sc.instances.size());
    for (size_t i = 0; i < sc.instances.size(); i++) {
        const bpt_instance_desc& d = sc.instances[i];
        InstanceXf& x = sc.xf[i];
        std::memcpy(x.o2w, d.transform, sizeof(float) * 12);
        invert_3x4(x.o2w, x.w2o);
        x.instance_id = d.instance_id_and_mask & 0xffffffu;
        x.flags = d.sbt_offset_and_flags >> 24;
        x.blas = (uint32_t)d.blas;
        if (x.blas >= sc.blas_descs.size()) { err = "instance references a BLAS out of range"; return false; }
        if (x.instance_id >= sc.drawables.size()) { err = "instance_id out of drawable range"; return false; }
    }
    return true;
}

bool build_tlas(Scene& sc, std::string& err) {
    if (!derive_instances(sc, err)) return false;
    size_t n = sc.xf.size();
    std::vector<f3> lo(n), hi(n);
    for (size_t i = 0; i < n; i++) {
        const Bvh& b = sc.blas[sc.xf[i].blas];
        // Transform::transform_bounding_box order (src/math/transform.cpp:59-78): 8 corners, min/max
        f3 mn = splat3(FLT_MAX), mx = splat3(-FLT_MAX);
        for (int c = 0; c < 8; c++) {
            f3 p = mk3((c & 4) ? b.hi.x : b.lo.x, (c & 2) ? b.hi.y : b.lo.y, (c & 1) ? b.hi.z : b.lo.z);
            f3 w = xf_point(sc.xf[i].o2w, p);
            mn = min3(mn, w); mx = max3(mx, w);
        }
        lo[i] = mn; hi[i] = mx;
    }
    build_lbvh(sc.tlas, (uint32_t)n, lo.data(), hi.data());
    return true;
}

bool build_accel(Scene& sc, uint32_t mode, std::string& err) {
    sc.accel_built = false;
    sc.wide_from_bounce = 0;
    if (sc.blas_descs.empty() || sc.instances.empty()) { err = "build_accel: no geometry or no instances"; return false; }
    for (auto& bd : sc.blas_descs) {
        if (bd.num_triangles == 0) { err = "empty BLAS"; return false; }
        if ((size_t)bd.index_offset + 3ull * bd.num_triangles > sc.indices.size()) { err = "BLAS index range out of bounds"; return false; }
    }
    sc.accel_mode = mode;
    if (mode == BPT_ACCEL_TWO_LEVEL) {
        sc.blas.assign(sc.blas_descs.size(), Bvh{});
        for (size_t bi = 0; bi < sc.blas_descs.size(); bi++) {
            const bpt_blas_desc& bd = sc.blas_descs[bi];
            uint32_t n = bd.num_triangles;
            std::vector<f3> lo(n), hi(n);
            for (uint32_t k = 0; k < n; k++) {
                f3 v[3]; fetch_tri(sc, bd, k, v);
                lo[k] = min3(min3(v[0], v[1]), v[2]);
                hi[k] = max3(max3(v[0], v[1]), v[2]);
            }
            Bvh& b = sc.blas[bi];
            build_lbvh(b, n, lo.data(), hi.data());
            b.tris.resize(n);
            for (uint32_t j = 0; j < n; j++) {
                f3 v[3]; fetch_tri(sc, bd, b.prims[j], v);
                b.tris[j] = Tri{v[0], b.prims[j], v[1] - v[0], 0u, v[2] - v[0], 0u};
            }
        }
        if (!build_tlas(sc, err)) return false;
    } else if (mode == BPT_ACCEL_MERGED) {
        if (!derive_instances(sc, err)) return false;
        size_t total = 0;
        for (auto& x : sc.xf) total += sc.blas_descs[x.blas].num_triangles;
        if (total > 0x7fffffffull) { err = "merged accel too large"; return false; }
        std::vector<f3> lo(total), hi(total), wv(total * 3);
        std::vector<uint32_t> tprim(total), tinst(total);
        size_t g = 0;
        for (size_t s = 0; s < sc.xf.size(); s++) {
            const bpt_blas_desc& bd = sc.blas_descs[sc.xf[s].blas];
            for (uint32_t k = 0; k < bd.num_triangles; k++, g++) {
                f3 v[3]; fetch_tri(sc, bd, k, v);
                for (int c = 0; c < 3; c++) { v[c] = xf_point(sc.xf[s].o2w, v[c]); wv[g * 3 + c] = v[c]; }
                lo[g] = min3(min3(v[0], v[1]), v[2]);
                hi[g] = max3(max3(v[0], v[1]), v[2]);
                tprim[g] = k; tinst[g] = (uint32_t)s;
            }
        }
        sc.blas.assign(1, Bvh{});
        Bvh& b = sc.blas[0];
        build_lbvh(b, (uint32_t)total, lo.data(), hi.data());
        b.tris.resize(total);
        for (size_t j = 0; j < total; j++) {
            size_t p = b.prims[j];
            b.tris[j] = Tri{wv[p * 3], tprim[p], wv[p * 3 + 1] - wv[p * 3], tinst[p], wv[p * 3 + 2] - wv[p * 3], 0u};
        }
        sc.tlas = Bvh{};
    } else {
        err = "unknown accel mode";
        return false;
    }
    sc.accel_built = true;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Traversal
//
// Result rule (what makes the answer independent of traversal ORDER, so the GPU may reorder work):
//   * a candidate replaces the best iff t < tbest || (t == tbest && id < best_id), id = (slot << 32 | prim);
//   * a box is entered iff t_near <= min(t_far, tcull) with tcull = tbest * (1 + 1e-5). The slab test and the
//     triangle test round differently (a few ulps), so culling against tbest itself could skip the box of an
//     equal-t (shared edge) or few-ulp-closer candidate depending on which was found first; the 1e-5 margin
//     (~84 ulps) guarantees every candidate that can still win is visited in any order.
// ---------------------------------------------------------------------------------------------
struct RayCtx {
    f3 O, D;             // world ray (opacity seed, rt_gbuffer.hlsl:24)
    float tmin;
    float tbest;
    float tcull;         // tbest * (1 + 1e-5): boxes are culled against this, see trace rule below
    uint64_t best_id;    // (instance slot << 32 | prim), tie-break on equal t
    float bu, bv;
    uint32_t best_slot, best_prim;
    bool any_mode, terminated;
    uint32_t frame_index;
    bool have_u; float opacity_u;
    bool cull_non_opaque;    // RAY_FLAG_CULL_NON_OPAQUE (ambient_occlusion_rt.hlsl:59)
};

static inline float opacity_random(RayCtx& rc) {        // hits/rt_gbuffer.hlsl:14-19
    if (!rc.have_u) {
        uint32_t seed = f2u(rc.O.x) ^ f2u(rc.O.y) ^ f2u(rc.O.z) ^ f2u(rc.D.x) ^ f2u(rc.D.y) ^ f2u(rc.D.z);
        uint32_t st = rng_tea(seed, rc.frame_index);
        rc.opacity_u = rng_next(st);
        rc.have_u = true;
    }
    return rc.opacity_u;
}

// any-hit filter, hits/rt_gbuffer_hit.hlsl:20-35. true = keep the hit.
static inline bool anyhit_accept(const Scene& sc, RayCtx& rc, uint32_t slot, uint32_t prim, float u, float v) {
    const InstanceXf& x = sc.xf[slot];
    if (!(x.flags & BPT_INSTANCE_FORCE_NON_OPAQUE)) return true;
    const bpt_drawable_sbt_data& dr = sc.drawables[x.instance_id];
    const bpt_material& m = sc.materials[dr.material_offset / sizeof(bpt_material)];
    uint32_t blend = (m.flags >> BPT_MATERIAL_BLEND_SHIFT) & 0xffu;
    if (blend == BPT_BLEND_OPAQUE) return true;  // MATERIAL_BLEND_MODE_OPAQUE: any-hit body compiled out
    // RAY_FLAG_CULL_NON_OPAQUE drops non-opaque geometry. The reference marks an instance non-opaque iff its material's
    // blend mode is not opaque (accel.cpp:121-125), i.e. exactly the triangles that reach this point.
    if (rc.cull_non_opaque) return false;
    float opacity = eval_opacity(sc, x.instance_id, prim, u, v);
    if (blend == BPT_BLEND_ALPHA_TEST) return !(opacity < 0.01f);
    return !(opacity_random(rc) < 1.0f - opacity);
}

static inline void tri_test(const Scene& sc, RayCtx& rc, const Tri& tr, f3 O, f3 D, uint32_t slot, TraceStats& st) {
    st.tris++;
    f3 pvec = cross(D, tr.e2);
    float det = dot(tr.e1, pvec);
    if (det == 0.0f) return;
    float inv = 1.0f / det;
    f3 tvec = O - tr.v0;
    float u = dot(tvec, pvec) * inv;
    if (u < 0.0f || u > 1.0f) return;
    f3 qvec = cross(tvec, tr.e1);
    float v = dot(D, qvec) * inv;
    if (v < 0.0f || u + v > 1.0f) return;
    float t = dot(tr.e2, qvec) * inv;
    if (!(t > rc.tmin)) return;
    uint64_t id = ((uint64_t)slot << 32) | tr.prim;
    if (!(t < rc.tbest || (t == rc.tbest && id < rc.best_id))) return;
    if (!anyhit_accept(sc, rc, slot, tr.prim, u, v)) return;
    rc.tbest = t; rc.tcull = t * 1.00001f; rc.best_id = id; rc.bu = u; rc.bv = v; rc.best_slot = slot; rc.best_prim = tr.prim;
    if (rc.any_mode) rc.terminated = true;
}

static inline void ray_prep(f3 D, f3 O, f3& idir, f3& ood) {
    const float ooeps = 8.27180613e-25f;  // 2^-80
    idir.x = 1.0f / (fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    idir.y = 1.0f / (fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    idir.z = 1.0f / (fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    ood = O * idir;
}

template <class LeafFn>
static inline void traverse(const Bvh& b, f3 O, f3 D, RayCtx& rc, TraceStats& st, LeafFn&& leaf) {
    if (b.n == 0) return;
    f3 idir, ood;
    ray_prep(D, O, idir, ood);
    int32_t stack[192];
    int sp = 0;
    int32_t cur = b.root;
    for (;;) {
        if (cur >= 0) {
            const bpt_bvh_node& n = b.nodes[cur];
            st.nodes++;
            float c0lox = fmaf(n.c0_lo_x, idir.x, -ood.x), c0hix = fmaf(n.c0_hi_x, idir.x, -ood.x);
            float c0loy = fmaf(n.c0_lo_y, idir.y, -ood.y), c0hiy = fmaf(n.c0_hi_y, idir.y, -ood.y);
            float c0loz = fmaf(n.c0_lo_z, idir.z, -ood.z), c0hiz = fmaf(n.c0_hi_z, idir.z, -ood.z);
            float c1lox = fmaf(n.c1_lo_x, idir.x, -ood.x), c1hix = fmaf(n.c1_hi_x, idir.x, -ood.x);
            float c1loy = fmaf(n.c1_lo_y, idir.y, -ood.y), c1hiy = fmaf(n.c1_hi_y, idir.y, -ood.y);
            float c1loz = fmaf(n.c1_lo_z, idir.z, -ood.z), c1hiz = fmaf(n.c1_hi_z, idir.z, -ood.z);
            float t0n = fmax_(fmax_(fmin_(c0lox, c0hix), fmin_(c0loy, c0hiy)), fmax_(fmin_(c0loz, c0hiz), rc.tmin));
            float t0f = fmin_(fmin_(fmax_(c0lox, c0hix), fmax_(c0loy, c0hiy)), fmin_(fmax_(c0loz, c0hiz), rc.tcull));
            float t1n = fmax_(fmax_(fmin_(c1lox, c1hix), fmin_(c1loy, c1hiy)), fmax_(fmin_(c1loz, c1hiz), rc.tmin));
            float t1f = fmin_(fmin_(fmax_(c1lox, c1hix), fmax_(c1loy, c1hiy)), fmin_(fmax_(c1loz, c1hiz), rc.tcull));
            bool h0 = t0n <= t0f, h1 = t1n <= t1f;
            if (h0 && h1) {
                bool c0_near = t0n <= t1n;
                stack[sp++] = c0_near ? n.child1 : n.child0;
                cur = c0_near ? n.child0 : n.child1;
                continue;
            } else if (h0) { cur = n.child0; continue; }
            else if (h1) { cur = n.child1; continue; }
        } else {
            leaf((uint32_t)~cur);
            if (rc.terminated) return;
        }
        if (sp == 0) return;
        cur = stack[--sp];
    }
}

// The 4-wide quantised tree (oracle_wide.cpp): nearest hit child first, the others pushed in slot order; a proposed leaf
// reaches the triangle test only if the ray passes its exact box (the child box held by its binary parent).
template <class LeafFn>
static inline void traverse_wide(const Bvh& b, f3 O, f3 D, RayCtx& rc, TraceStats& st, LeafFn&& leaf) {
    if (b.n == 0) return;
    f3 idir, ood;
    ray_prep(D, O, idir, ood);
    auto slab = [&](f3 lo, f3 hi, float& tn) {
        float lx = fmaf(lo.x, idir.x, -ood.x), hx = fmaf(hi.x, idir.x, -ood.x);
        float ly = fmaf(lo.y, idir.y, -ood.y), hy = fmaf(hi.y, idir.y, -ood.y);
        float lz = fmaf(lo.z, idir.z, -ood.z), hz = fmaf(hi.z, idir.z, -ood.z);
        tn = fmax_(fmax_(fmin_(lx, hx), fmin_(ly, hy)), fmax_(fmin_(lz, hz), rc.tmin));
        float tf = fmin_(fmin_(fmax_(lx, hx), fmax_(ly, hy)), fmin_(fmax_(lz, hz), rc.tcull));
        return tn <= tf;
    };
    int32_t stack[512];
    int sp = 0;
    int32_t cur = b.root;
    for (;;) {
        if (cur >= 0) {
            const Bvh::Wide4& n = b.wide[cur];
            st.wide_nodes++;
            int best = -1; float tb = 0.0f; bool hit[4] = {false, false, false, false};
            for (int k = 0; k < n.nchild; k++) {
                float tn;
                if (slab(n.clo[k], n.chi[k], tn)) { hit[k] = true; if (best < 0 || tn < tb) { best = k; tb = tn; } }
            }
            if (best >= 0) {
                for (int k = 0; k < n.nchild; k++) if (hit[k] && k != best) stack[sp++] = n.child[k];
                cur = n.child[best];
                continue;
            }
        } else {
            const uint32_t j = (uint32_t)~cur;
            bool candidate = true;
            if (b.n > 1) {
                const bpt_bvh_node& pn = b.nodes[b.leaf_parent[j]];
                const bool first = pn.child0 == cur;
                float tn;
                st.leaf_boxes++;
                candidate = first ? slab(mk3(pn.c0_lo_x, pn.c0_lo_y, pn.c0_lo_z), mk3(pn.c0_hi_x, pn.c0_hi_y, pn.c0_hi_z), tn)
                                  : slab(mk3(pn.c1_lo_x, pn.c1_lo_y, pn.c1_lo_z), mk3(pn.c1_hi_x, pn.c1_hi_y, pn.c1_hi_z), tn);
            }
            if (candidate) {
                leaf(j);
                if (rc.terminated) return;
            }
        }
        if (sp == 0) return;
        cur = stack[--sp];
    }
}

static void trace_generic(const Scene& sc, RayCtx& rc, TraceStats& st, bool wide = false) {
    st.rays++;
    if (sc.accel_mode == BPT_ACCEL_MERGED && wide && !sc.blas[0].wide.empty()) {
        const Bvh& b = sc.blas[0];
        traverse_wide(b, rc.O, rc.D, rc, st, [&](uint32_t j) {
            const Tri& tr = b.tris[j];
            tri_test(sc, rc, tr, rc.O, rc.D, tr.inst, st);
        });
    } else if (sc.accel_mode == BPT_ACCEL_MERGED) {
        const Bvh& b = sc.blas[0];
        traverse(b, rc.O, rc.D, rc, st, [&](uint32_t j) {
            const Tri& tr = b.tris[j];
            tri_test(sc, rc, tr, rc.O, rc.D, tr.inst, st);
        });
    } else {
        traverse(sc.tlas, rc.O, rc.D, rc, st, [&](uint32_t j) {
            uint32_t slot = sc.tlas.prims[j];
            st.instances++;
            const InstanceXf& x = sc.xf[slot];
            f3 Oo = xf_point(x.w2o, rc.O), Do = xf_vector(x.w2o, rc.D);
            const Bvh& b = sc.blas[x.blas];
            traverse(b, Oo, Do, rc, st, [&](uint32_t jj) { tri_test(sc, rc, b.tris[jj], Oo, Do, slot, st); });
        });
    }
}

HitRec trace_closest(const Scene& sc, f3 O, f3 D, float tmin, float tmax, uint32_t frame_index, TraceStats& st, bool wide) {
    RayCtx rc{};
    rc.O = O; rc.D = D; rc.tmin = tmin; rc.tbest = tmax; rc.tcull = tmax * 1.00001f; rc.best_id = ~0ull;
    rc.any_mode = false; rc.terminated = false; rc.frame_index = frame_index; rc.have_u = false;
    trace_generic(sc, rc, st, wide);
    HitRec h{};
    h.hit = rc.best_id != ~0ull;
    if (h.hit) {
        h.t = rc.tbest; h.u = rc.bu; h.v = rc.bv; h.inst_slot = rc.best_slot; h.prim = rc.best_prim;
        h.instance_id = sc.xf[rc.best_slot].instance_id;
    } else { h.t = -1.0f; }
    return h;
}

bool trace_any(const Scene& sc, f3 O, f3 D, float tmin, float tmax, uint32_t frame_index, TraceStats& st, bool cull_non_opaque, bool wide) {
    RayCtx rc{};
    rc.cull_non_opaque = cull_non_opaque;
    rc.O = O; rc.D = D; rc.tmin = tmin; rc.tbest = tmax; rc.tcull = tmax * 1.00001f; rc.best_id = ~0ull;
    rc.any_mode = true; rc.terminated = false; rc.frame_index = frame_index; rc.have_u = false;
    trace_generic(sc, rc, st, wide);
    return rc.terminated;  // true = occluded
}

} // namespace orc
