// oracle_wide.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle).
//
// Wide (4- / 8-ary) BVH derived from the binary LBVH by collapsing, and its traversal, restated on the CPU:
//   * to count the work of the wide traversal (nodes / triangles per ray = the "algorithmic bytes" of the
//     CUDA extend / connect kernels) on the same rays, and
//   * as the definition the CUDA collapse must reproduce (same children, same quantised boxes).
// NEW code (the reference traverses with RT hardware, rt_gbuffer.hlsl:17-25). Hits do not depend on the
// acceleration structure: every structure is conservative and the result rule of oracle_bvh.cpp (tie-break +
// cull margin) makes the answer independent of traversal order, so the wide traversal must return exactly what
// the binary one returns — checked in tests/test_oracle.py.
//
// Collapse rule: a wide node starts as the two children of a binary node; while it has fewer than W children,
// the INTERNAL child with the largest box surface area (ties: lowest slot) is replaced in place by its two
// children (child0 takes the slot, child1 is appended). Quantisation (8-bit, per node): origin = node box lo,
// per-axis scale 2^e with e the smallest exponent such that (hi - lo) / 2^e <= 255; child lo rounds down, child
// hi rounds up, in units of 2^e from the origin.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include "oracle_scene.hpp"

namespace orc {

struct WideNode {
    f3 lo, hi;                 // this node's box (union of its children)
    int nchild = 0;
    int32_t child[8];          // >= 0: wide node index, < 0: ~leaf index (sorted-primitive order of the binary BVH)
    f3 clo[8], chi[8];         // child boxes as tested (after optional quantisation)
};

struct WideBvh { std::vector<WideNode> nodes; int32_t root = 0; };

static inline float area(f3 lo, f3 hi) { f3 d = hi - lo; return (d.x * d.y + d.y * d.z) + d.z * d.x; }

static void quantise(WideNode& w) {
    const float* lo = &w.lo.x; const float* hi = &w.hi.x;
    for (int a = 0; a < 3; a++) {
        float ext = hi[a] - lo[a];
        int e = -126;
        if (ext > 0.0f) { int ex; (void)frexpf(ext / 255.0f, &ex); e = ex; while (ldexpf(255.0f, e) < ext) e++; while (e > -126 && ldexpf(255.0f, e - 1) >= ext) e--; }
        float scale = ldexpf(1.0f, e);
        for (int c = 0; c < w.nchild; c++) {
            float* cl = &w.clo[c].x; float* ch = &w.chi[c].x;
            float ql = floorf((cl[a] - lo[a]) / scale), qh = ceilf((ch[a] - lo[a]) / scale);
            ql = fmin_(fmax_(ql, 0.0f), 255.0f); qh = fmin_(fmax_(qh, 0.0f), 255.0f);
            // the decoded planes must contain the exact ones whatever the rounding of the divisions did
            while (lo[a] + ql * scale > cl[a] && ql > 0.0f) ql -= 1.0f;
            while (lo[a] + qh * scale < ch[a] && qh < 255.0f) qh += 1.0f;
            cl[a] = lo[a] + ql * scale; ch[a] = lo[a] + qh * scale;
        }
    }
}

static void collapse(const Bvh& b, int W, bool quant, WideBvh& out) {
    out.nodes.clear();
    if (b.n < 2) { out.root = ~0; return; }
    struct Item { int32_t bin; int32_t wide; };
    std::vector<Item> todo;
    out.nodes.emplace_back();
    todo.push_back({b.root, 0});
    auto child_boxes = [&](int32_t bin, f3 lo[2], f3 hi[2], int32_t ch[2]) {
        const bpt_bvh_node& n = b.nodes[bin];
        lo[0] = mk3(n.c0_lo_x, n.c0_lo_y, n.c0_lo_z); hi[0] = mk3(n.c0_hi_x, n.c0_hi_y, n.c0_hi_z);
        lo[1] = mk3(n.c1_lo_x, n.c1_lo_y, n.c1_lo_z); hi[1] = mk3(n.c1_hi_x, n.c1_hi_y, n.c1_hi_z);
        ch[0] = n.child0; ch[1] = n.child1;
    };
    while (!todo.empty()) {
        Item it = todo.back(); todo.pop_back();
        int32_t cb[8]; f3 lo[8], hi[8];
        int nc = 2;
        { f3 l[2], h[2]; int32_t c[2]; child_boxes(it.bin, l, h, c); for (int k = 0; k < 2; k++) { cb[k] = c[k]; lo[k] = l[k]; hi[k] = h[k]; } }
        while (nc < W) {
            int best = -1; float ba = -1.0f;
            for (int k = 0; k < nc; k++) if (cb[k] >= 0) { float a = area(lo[k], hi[k]); if (a > ba) { ba = a; best = k; } }
            if (best < 0) break;
            f3 l[2], h[2]; int32_t c[2]; child_boxes(cb[best], l, h, c);
            cb[best] = c[0]; lo[best] = l[0]; hi[best] = h[0];
            cb[nc] = c[1]; lo[nc] = l[1]; hi[nc] = h[1]; nc++;
        }
        WideNode w;
        w.nchild = nc;
        w.lo = lo[0]; w.hi = hi[0];
        for (int k = 1; k < nc; k++) { w.lo = min3(w.lo, lo[k]); w.hi = max3(w.hi, hi[k]); }
        for (int k = 0; k < nc; k++) {
            w.clo[k] = lo[k]; w.chi[k] = hi[k];
            if (cb[k] < 0) w.child[k] = cb[k];
            else { w.child[k] = (int32_t)out.nodes.size(); out.nodes.emplace_back(); todo.push_back({cb[k], w.child[k]}); }
        }
        if (quant) quantise(w);
        out.nodes[it.wide] = w;
    }
    out.root = 0;
}

struct WideCounts { uint64_t rays = 0, nodes = 0, tris = 0, boxes = 0; };

// order: 0 = nearest hit child first, the rest pushed in slot order; 1 = all hit children visited near to far
static void wide_closest(const Bvh& b, const WideBvh& w, f3 O, f3 D, float tmin, float tmax, int order, WideCounts& cnt, float& tout, uint32_t& pout) {
    f3 idir, ood;
    const float ooeps = 8.27180613e-25f;
    idir.x = 1.0f / (fabsf(D.x) > ooeps ? D.x : copysignf(ooeps, D.x));
    idir.y = 1.0f / (fabsf(D.y) > ooeps ? D.y : copysignf(ooeps, D.y));
    idir.z = 1.0f / (fabsf(D.z) > ooeps ? D.z : copysignf(ooeps, D.z));
    ood = O * idir;
    float tbest = tmax, tcull = tmax * 1.00001f; uint64_t best_id = ~0ull;
    cnt.rays++;
    int32_t stack[512]; int sp = 0;
    int32_t cur = w.root;
    if (b.n < 2) { tout = -1.0f; pout = ~0u; return; }
    for (;;) {
        if (cur >= 0) {
            const WideNode& n = w.nodes[cur];
            cnt.nodes++; cnt.boxes += n.nchild;
            int hit[8]; float tn[8]; int nh = 0;
            for (int k = 0; k < n.nchild; k++) {
                float lx = fmaf(n.clo[k].x, idir.x, -ood.x), hx = fmaf(n.chi[k].x, idir.x, -ood.x);
                float ly = fmaf(n.clo[k].y, idir.y, -ood.y), hy = fmaf(n.chi[k].y, idir.y, -ood.y);
                float lz = fmaf(n.clo[k].z, idir.z, -ood.z), hz = fmaf(n.chi[k].z, idir.z, -ood.z);
                float t0 = fmax_(fmax_(fmin_(lx, hx), fmin_(ly, hy)), fmax_(fmin_(lz, hz), tmin));
                float t1 = fmin_(fmin_(fmax_(lx, hx), fmax_(ly, hy)), fmin_(fmax_(lz, hz), tcull));
                if (t0 <= t1) { hit[nh] = k; tn[nh] = t0; nh++; }
            }
            if (nh) {
                if (order == 1) {
                    for (int i = 1; i < nh; i++) { int h = hit[i]; float t = tn[i]; int j = i - 1; while (j >= 0 && tn[j] > t) { hit[j + 1] = hit[j]; tn[j + 1] = tn[j]; j--; } hit[j + 1] = h; tn[j + 1] = t; }
                } else {
                    int m = 0; for (int i = 1; i < nh; i++) if (tn[i] < tn[m]) m = i;
                    std::swap(hit[0], hit[m]); std::swap(tn[0], tn[m]);
                }
                for (int i = nh - 1; i >= 1; i--) stack[sp++] = n.child[hit[i]];
                cur = n.child[hit[0]];
                continue;
            }
        } else {
            const Tri& tr = b.tris[(uint32_t)~cur];
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
        if (sp == 0) break;
        cur = stack[--sp];
    }
    tout = best_id != ~0ull ? tbest : -1.0f; pout = (uint32_t)best_id;
}

} // namespace orc

using namespace orc;

extern "C" {
// Experimental statistics: merged-mode BVH collapsed to `width`-ary nodes; closest hits of `rays` (any-hit filters
// ignored: opaque scenes only). counts = {rays, wide nodes visited, triangles tested, child boxes tested, wide nodes total}.
__attribute__((visibility("default")))
bpt_status obpt_wide_stats(obpt_context* c, uint32_t width, uint32_t quant, uint32_t order, const bpt_ray* rays, uint64_t n, float* t_out, uint32_t* prim_out, uint64_t counts[5]) {
    if (!c || !c->scene.accel_built || c->scene.accel_mode != BPT_ACCEL_MERGED || (width != 4 && width != 8 && width != 2 && width != 6)) return BPT_ERR_INVALID;
    const Bvh& b = c->scene.blas[0];
    WideBvh w;
    collapse(b, (int)width, quant != 0, w);
    WideCounts cnt;
    for (uint64_t i = 0; i < n; i++) {
        float t; uint32_t p;
        wide_closest(b, w, mk3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]), mk3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]),
                     rays[i].tmin, rays[i].tmax, (int)order, cnt, t, p);
        if (t_out) t_out[i] = t;
        if (prim_out) prim_out[i] = p;
    }
    counts[0] = cnt.rays; counts[1] = cnt.nodes; counts[2] = cnt.tris; counts[3] = cnt.boxes; counts[4] = w.nodes.size();
    return BPT_OK;
}
}
