// oracle_wide.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// The 4-wide BVH with 8-bit child boxes that the CUDA traversal kernels walk in merged mode (csrc/bpt_wide.cuh), restated
// on the CPU: (1) as the DEFINITION the CUDA collapse must reproduce bit for bit (obpt_debug_read_wide), (2) to count the
// work of the wide traversal on the same rays (wide nodes / exact leaf boxes / triangles per ray = the "algorithmic
// bytes" of the extend and connect kernels), and (3) to check that it returns exactly the hits of the binary
// traversal of oracle_bvh.cpp. It also still answers the design question it was first written for: how many steps a 4- or
// 8-wide tree saves on this scene (obpt_wide_stats with other widths).
// NEW code: the reference traverses with RT hardware (rt_gbuffer.hlsl:17-25).
//
// Definition.
//   Collapse: wide node i describes the subtree of binary node i. Start with the two children of binary node i; while
//   there are fewer than W, replace the INTERNAL child with the largest box area ((dx*dy + dy*dz) + dz*dx; ties: lowest
//   slot) by its two children — child0 takes the slot, child1 is appended.
//   Quantisation, per axis: origin o = node lo; e = smallest exponent >= -126 with 255 * 2^e >= hi - lo, read off the
//   exponent field of (hi - lo) (E - 7 or E - 6); child lo -> floor((lo - o) / 2^e), child hi -> ceil((hi - o) / 2^e),
//   clamped to [0, 255], then stepped outwards while the DECODED plane fmaf(q, 2^e, o) still cuts into the child box; if
//   that cannot be satisfied (rounding of hi - lo), e + 1 is tried. Empty slots: lo = 255, hi = 0.
//   Traversal: a hit child is one whose decoded box passes the slab test of oracle_bvh.cpp; the nearest (ties: lowest
//   slot) is visited next, the others are pushed in slot order. A proposed leaf is tested only if the ray passes the
//   leaf's EXACT box (the child box stored in its binary parent) — slab values are monotone in the plane coordinates, so
//   that implies every exact ancestor box too, and the triangles that reach the triangle test are exactly those of the
//   binary traversal.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include "oracle_scene.hpp"

namespace orc {

struct WideNode {
    int nchild = 0;
    int32_t child[8];          // >= 0: binary/wide node index, < 0: ~leaf (sorted-triangle position)
    f3 clo[8], chi[8];         // child boxes as tested (decoded when quantised)
    f3 origin{0, 0, 0};
    int e[3] = {0, 0, 0};
    uint8_t qlo[3][8], qhi[3][8];
};

static inline float area(f3 lo, f3 hi) { f3 d = hi - lo; return (d.x * d.y + d.y * d.z) + d.z * d.x; }
static inline void child_boxes(const Bvh& b, int32_t bin, f3 lo[2], f3 hi[2], int32_t ch[2]) {
    const bpt_bvh_node& n = b.nodes[bin];
    lo[0] = mk3(n.c0_lo_x, n.c0_lo_y, n.c0_lo_z); hi[0] = mk3(n.c0_hi_x, n.c0_hi_y, n.c0_hi_z);
    lo[1] = mk3(n.c1_lo_x, n.c1_lo_y, n.c1_lo_z); hi[1] = mk3(n.c1_hi_x, n.c1_hi_y, n.c1_hi_z);
    ch[0] = n.child0; ch[1] = n.child1;
}

static int smallest_exponent(float ext) {
    uint32_t eb = (f2u(ext) >> 23) & 0xffu;
    int E = eb == 0 ? -126 : (int)eb - 127;
    int e = std::max(E - 7, -126);
    if (255.0f * ldexpf(1.0f, e) < ext) e += 1;
    return e;
}

static void quantise(WideNode& w, int W) {
    f3 lo = w.clo[0], hi = w.chi[0];
    for (int k = 1; k < w.nchild; k++) { lo = min3(lo, w.clo[k]); hi = max3(hi, w.chi[k]); }
    w.origin = lo;
    const float* o3 = &lo.x; const float* h3 = &hi.x;
    for (int a = 0; a < 3; a++) {
        const float o = o3[a];
        int e = smallest_exponent(h3[a] - o);
        for (;;) {
            const float scale = ldexpf(1.0f, e);
            bool ok = true;
            for (int k = 0; k < W; k++) {
                float ql = 255.0f, qh = 0.0f;
                if (k < w.nchild) {
                    const float cl = (&w.clo[k].x)[a], ch = (&w.chi[k].x)[a];
                    ql = fmin_(fmax_(floorf((cl - o) / scale), 0.0f), 255.0f);
                    qh = fmin_(fmax_(ceilf((ch - o) / scale), 0.0f), 255.0f);
                    while (ql > 0.0f && fmaf(ql, scale, o) > cl) ql -= 1.0f;
                    while (qh < 255.0f && fmaf(qh, scale, o) < ch) qh += 1.0f;
                    if (fmaf(ql, scale, o) > cl || fmaf(qh, scale, o) < ch) ok = false;
                }
                w.qlo[a][k] = (uint8_t)ql; w.qhi[a][k] = (uint8_t)qh;
            }
            if (ok || e >= 127) break;
            e += 1;
        }
        w.e[a] = e;
        const float scale = ldexpf(1.0f, e);
        for (int k = 0; k < w.nchild; k++) {
            (&w.clo[k].x)[a] = fmaf((float)w.qlo[a][k], scale, o);
            (&w.chi[k].x)[a] = fmaf((float)w.qhi[a][k], scale, o);
        }
    }
}

// wide node of binary node `bin`
static WideNode collapse_one(const Bvh& b, int32_t bin, int W, bool quant) {
    WideNode w;
    int32_t cb[8]; f3 lo[8], hi[8];
    int nc = 2;
    { f3 l[2], h[2]; int32_t c[2]; child_boxes(b, bin, l, h, c); for (int k = 0; k < 2; k++) { cb[k] = c[k]; lo[k] = l[k]; hi[k] = h[k]; } }
    while (nc < W) {
        int best = -1; float ba = -1.0f;
        for (int k = 0; k < nc; k++) if (cb[k] >= 0) { float a = area(lo[k], hi[k]); if (a > ba) { ba = a; best = k; } }
        if (best < 0) break;
        f3 l[2], h[2]; int32_t c[2]; child_boxes(b, cb[best], l, h, c);
        cb[best] = c[0]; lo[best] = l[0]; hi[best] = h[0];
        cb[nc] = c[1]; lo[nc] = l[1]; hi[nc] = h[1]; nc++;
    }
    w.nchild = nc;
    for (int k = 0; k < nc; k++) { w.child[k] = cb[k]; w.clo[k] = lo[k]; w.chi[k] = hi[k]; }
    if (quant) quantise(w, W);
    return w;
}

struct WideBvh { std::vector<WideNode> nodes; };
static void collapse_all(const Bvh& b, int W, bool quant, WideBvh& out) {
    out.nodes.clear();
    if (b.n < 2) return;
    out.nodes.resize(b.n - 1);
    for (uint32_t i = 0; i + 1 < b.n; i++) out.nodes[i] = collapse_one(b, (int32_t)i, W, quant);
}

} // namespace orc
// Fills b.wide (declared in oracle_scene.hpp) with the decoded 4-wide quantised nodes, one per binary node.
void orc::build_wide4(Bvh& b) {
    b.wide.clear();
    if (b.n < 2) return;
    b.wide.resize(b.n - 1);
    for (uint32_t i = 0; i + 1 < b.n; i++) {
        WideNode w = collapse_one(b, (int32_t)i, 4, true);
        Bvh::Wide4& o = b.wide[i];
        o.nchild = w.nchild;
        for (int k = 0; k < 4; k++) { o.child[k] = k < w.nchild ? w.child[k] : 0; o.clo[k] = w.clo[k]; o.chi[k] = w.chi[k]; }
    }
}
namespace orc {

struct WideCounts { uint64_t rays = 0, nodes = 0, tris = 0, boxes = 0, leaf_boxes = 0; };

static inline bool slab(f3 lo, f3 hi, f3 idir, f3 ood, float tmin, float tcull, float& tnear) {
    float lx = fmaf(lo.x, idir.x, -ood.x), hx = fmaf(hi.x, idir.x, -ood.x);
    float ly = fmaf(lo.y, idir.y, -ood.y), hy = fmaf(hi.y, idir.y, -ood.y);
    float lz = fmaf(lo.z, idir.z, -ood.z), hz = fmaf(hi.z, idir.z, -ood.z);
    float t0 = fmax_(fmax_(fmin_(lx, hx), fmin_(ly, hy)), fmax_(fmin_(lz, hz), tmin));
    float t1 = fmin_(fmin_(fmax_(lx, hx), fmax_(ly, hy)), fmin_(fmax_(lz, hz), tcull));
    tnear = t0;
    return t0 <= t1;
}

// order: 0 = nearest hit child first, the rest in slot order (what the kernels do); 1 = all hit children near to far.
// leaf_check: test the exact leaf box before the triangle (what the kernels do when the boxes are quantised).
static void wide_closest(const Bvh& b, const WideBvh& w, f3 O, f3 D, float tmin, float tmax, int order, bool leaf_check, WideCounts& cnt, float& tout, uint32_t& pout) {
    f3 idir, ood;
    const float ooeps = 8.27180613e-25f;
    idir.x = 1.0f / (fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    idir.y = 1.0f / (fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    idir.z = 1.0f / (fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    ood = O * idir;
    float tbest = tmax, tcull = tmax * 1.00001f; uint64_t best_id = ~0ull;
    cnt.rays++;
    tout = -1.0f; pout = ~0u;
    if (b.n == 0) return;
    std::vector<int32_t> stack; stack.reserve(256);
    int32_t cur = b.root;
    for (;;) {
        if (cur >= 0) {
            const WideNode& n = w.nodes[cur];
            cnt.nodes++; cnt.boxes += n.nchild;
            int hit[8]; float tn[8]; int nh = 0;
            for (int k = 0; k < n.nchild; k++) { float t0; if (slab(n.clo[k], n.chi[k], idir, ood, tmin, tcull, t0)) { hit[nh] = k; tn[nh] = t0; nh++; } }
            if (nh) {
                if (order == 1) {
                    for (int i = 1; i < nh; i++) { int h = hit[i]; float t = tn[i]; int j = i - 1; while (j >= 0 && tn[j] > t) { hit[j + 1] = hit[j]; tn[j + 1] = tn[j]; j--; } hit[j + 1] = h; tn[j + 1] = t; }
                    for (int i = nh - 1; i >= 1; i--) stack.push_back(n.child[hit[i]]);
                    cur = n.child[hit[0]];
                } else {
                    int m = 0; for (int i = 1; i < nh; i++) if (tn[i] < tn[m]) m = i;
                    for (int i = 0; i < nh; i++) if (i != m) stack.push_back(n.child[hit[i]]);
                    cur = n.child[hit[m]];
                }
                continue;
            }
        } else {
            const uint32_t j = (uint32_t)~cur;
            bool candidate = true;
            if (leaf_check && b.n > 1) {
                const bpt_bvh_node& pn = b.nodes[b.leaf_parent[j]];
                const bool first = pn.child0 == cur;
                f3 lo = first ? mk3(pn.c0_lo_x, pn.c0_lo_y, pn.c0_lo_z) : mk3(pn.c1_lo_x, pn.c1_lo_y, pn.c1_lo_z);
                f3 hi = first ? mk3(pn.c0_hi_x, pn.c0_hi_y, pn.c0_hi_z) : mk3(pn.c1_hi_x, pn.c1_hi_y, pn.c1_hi_z);
                float t0; cnt.leaf_boxes++;
                candidate = slab(lo, hi, idir, ood, tmin, tcull, t0);
            }
            if (candidate) {
                const Tri& tr = b.tris[j];
                cnt.tris++;
                f3 pvec = cross(D, tr.e2);
                float det = dot(tr.e1, pvec);
                if (det != 0.0f) {
                    float inv = 1.0f / det;
                    f3 tvec = O - tr.v0;
                    float u = dot(tvec, pvec) * inv;
                    if (!(u < 0.0f || u > 1.0f)) {
                        f3 qvec = cross(tvec, tr.e1);
                        float v = dot(D, qvec) * inv;
                        if (!(v < 0.0f || u + v > 1.0f)) {
                            float t = dot(tr.e2, qvec) * inv;
                            uint64_t id = ((uint64_t)tr.inst << 32) | tr.prim;
                            if (t > tmin && (t < tbest || (t == tbest && id < best_id))) { tbest = t; tcull = t * 1.00001f; best_id = id; }
                        }
                    }
                }
            }
        }
        if (stack.empty()) break;
        cur = stack.back(); stack.pop_back();
    }
    if (best_id != ~0ull) { tout = tbest; pout = (uint32_t)best_id; }
}

} // namespace orc

using namespace orc;

extern "C" {
// Statistics of a `width`-ary collapse of the merged BVH on a ray batch (closest hits; any-hit filters ignored: opaque scenes).
// counts = {rays, wide nodes visited, triangles tested, child boxes tested, exact leaf boxes tested}.
__attribute__((visibility("default")))
bpt_status obpt_wide_stats(obpt_context* c, uint32_t width, uint32_t quant, uint32_t order, const bpt_ray* rays, uint64_t n, float* t_out, uint32_t* prim_out, uint64_t counts[5]) {
    if (!c || !c->scene.accel_built || c->scene.accel_mode != BPT_ACCEL_MERGED || width < 2 || width > 8) return BPT_ERR_INVALID;
    const Bvh& b = c->scene.blas[0];
    WideBvh w;
    collapse_all(b, (int)width, quant != 0, w);
    WideCounts cnt;
    for (uint64_t i = 0; i < n; i++) {
        float t; uint32_t p;
        wide_closest(b, w, mk3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), mk3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]),
                     rays[i].tmin, rays[i].tmax, (int)order, quant != 0, cnt, t, p);
        if (t_out) t_out[i] = t;
        if (prim_out) prim_out[i] = p;
    }
    counts[0] = cnt.rays; counts[1] = cnt.nodes; counts[2] = cnt.tris; counts[3] = cnt.boxes; counts[4] = cnt.leaf_boxes;
    return BPT_OK;
}

// The 64-byte nodes and the exact leaf boxes exactly as bpt_debug_read_wide returns them (layout: csrc/bpt_wide.cuh).
__attribute__((visibility("default")))
bpt_status obpt_debug_read_wide(obpt_context* c, float* wide, float* leafbox, uint32_t cap) {
    if (!c || !c->scene.accel_built || c->scene.accel_mode != BPT_ACCEL_MERGED) return BPT_ERR_STATE;
    const Bvh& b = c->scene.blas[0];
    if (cap < b.n) return BPT_ERR_INVALID;
    if (b.n < 2) return BPT_OK;
    const int32_t kNoChild = 0x7ffffffe;
    for (uint32_t i = 0; i + 1 < b.n; i++) {
        if (wide) {
            WideNode w = collapse_one(b, (int32_t)i, 4, true);
            uint32_t out[16];
            out[0] = f2u(w.origin.x); out[1] = f2u(w.origin.y); out[2] = f2u(w.origin.z);
            out[3] = (uint32_t)(w.e[0] + 127) | ((uint32_t)(w.e[1] + 127) << 8) | ((uint32_t)(w.e[2] + 127) << 16);
            for (int k = 0; k < 4; k++) out[4 + k] = (uint32_t)(k < w.nchild ? w.child[k] : kNoChild);
            auto pack = [&](const uint8_t q[8]) { return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24); };
            out[8] = pack(w.qlo[0]); out[9] = pack(w.qlo[1]); out[10] = pack(w.qlo[2]); out[11] = pack(w.qhi[0]);
            out[12] = pack(w.qhi[1]); out[13] = pack(w.qhi[2]); out[14] = 0; out[15] = 0;
            std::memcpy(wide + 16ull * i, out, 64);
        }
        if (leafbox) {
            const bpt_bvh_node& n = b.nodes[i];
            if (n.child0 < 0) { float* o = leafbox + 8ull * (uint32_t)~n.child0; o[0] = n.c0_lo_x; o[1] = n.c0_lo_y; o[2] = n.c0_lo_z; o[3] = 0; o[4] = n.c0_hi_x; o[5] = n.c0_hi_y; o[6] = n.c0_hi_z; o[7] = 0; }
            if (n.child1 < 0) { float* o = leafbox + 8ull * (uint32_t)~n.child1; o[0] = n.c1_lo_x; o[1] = n.c1_lo_y; o[2] = n.c1_lo_z; o[3] = 0; o[4] = n.c1_hi_x; o[5] = n.c1_hi_y; o[6] = n.c1_hi_z; o[7] = 0; }
        }
    }
    return BPT_OK;
}
}
