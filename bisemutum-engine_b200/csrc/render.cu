// render.cu — the wavefront: raygen → (extend → shade → connect)^(B-1) per sample, FP32 accumulation.
//
// Replaces the per-frame GPU work recorded by PathTracingPass::render
// (bisemutum/src/renderer/pass/path_tracing.cpp:290-480): the reference runs full-screen passes
// over screen-sized textures (one texel = one path, dead paths keep their thread). Here paths live
// in compacted queues: every kernel reads its queue length from device memory, so no host
// synchronisation happens inside a sample, and dead paths cost nothing.
//
// Queue compaction: warp ballot + popc, one atomicAdd per warp. Queue ORDER is
// therefore not deterministic, queue CONTENT is (every entry is keyed by its pixel index).
#include <algorithm>
#include <cstdio>
#include <type_traits>
#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: a no-op unless a profiler injects itself
#include "bpt_internal.cuh"
#include "bpt_shade.cuh"
#include "bpt_ddgi.cuh"
#include "bpt_aov.cuh"
#include "bpt_wide.cuh"
#pragma nv_diag_suppress 128   // "loop is not reachable": the two-level branch of the merged-mode instantiation

using namespace bptd;

namespace {

constexpr int kBlock = 128;
constexpr int QE = 0;      // qcount[QE + bounce]  : extend queue length of that bounce
constexpr int QS = 32;     // qcount[QS + bounce]  : shadow queue length of that bounce
constexpr int QWE = 64;    // qcount[QWE + bounce] : work-fetch cursor of the persistent extend kernel
constexpr int QWS = 96;    // qcount[QWS + bounce] : work-fetch cursor of the persistent connect kernel
constexpr int QN = 128;
// Tunables of the persistent traversal kernels (overridable at compile time for tuning runs, tools/build_variant.sh).
// Measured on configs[1] with the binary tree: refill threshold 2 -> 1.18, 8 -> 1.14, 14 -> 1.16, 20 -> 1.21, 26 -> 1.25 ms of
// extend per 1080p sample (refilling early mixes root-level rays into warps that are deep in the tree and costs more than the
// idle lanes). Re-tuned with the 4-wide tree from bounce 2 (whole frame): node-phase exit at 1 / 4 / 6 / 8 / 12 / 16 lanes ->
// 1.888 / 1.739 / 1.704 / 1.696 / 1.709 / 1.739 ms; refill 4 / 8 / 14 -> 1.755 / 1.709 / 1.707; (8, 12) -> 1.690 ms.
#ifndef BPT_MIN_NODE_LANES
#define BPT_MIN_NODE_LANES 8
#endif
#ifndef BPT_REFILL
#define BPT_REFILL 8
#endif
#ifndef BPT_SMEM_STACK
#define BPT_SMEM_STACK 0
#endif
// Two-level mode: 1 = a lane that reaches a TLAS leaf PARKS there until the node phase ends; all parked lanes then enter their instances together
// (instance record fetch, ray transform, three IEEE divisions of make_space: ~100 instructions that otherwise ran at ~2 active lanes)
#ifndef BPT_PARK_INSTANCES
#define BPT_PARK_INSTANCES 1
#endif
#ifndef BPT_SORT_OCTANT
#define BPT_SORT_OCTANT 0
#endif
constexpr int kSmemStack = BPT_SMEM_STACK;             // stack entries per lane kept in shared memory (0: the whole stack in local memory)
// Two-level mode has its own pair (instanced 2 M x 512 scene, ms per 4K sample: (4, 12) 43.7, (8, 12) 40.2, (12, 16) 38.9, (16, 20) 37.9 ...:
// leaving the node phase earlier keeps more lanes together through the instance entries; with the parked instance entries (BPT_PARK_INSTANCES):
// (8, 12) 38.6, (12, 16) 36.5, (16, 20) 35.3, (16, 24) 35.0, (20, 24) 35.0, (24, 28) 35.5)
#ifndef BPT_MIN_NODE_LANES_2L
#define BPT_MIN_NODE_LANES_2L 16
#endif
#ifndef BPT_REFILL_2L
#define BPT_REFILL_2L 24
#endif
constexpr int kMinNodeLanes1 = BPT_MIN_NODE_LANES;      // node phase ends when fewer lanes than this still have an internal node
constexpr int kRefillThreshold1 = BPT_REFILL;           // refill a warp's finished lanes when fewer rays than this are still in flight
// The merged-mode kernels that walk the 4-wide tree (the incoherent bounces) have their own pair. Round 2, whole frame on configs[1]
// (profiles/r2p..r2r_variants.jsonl; binary pair (8, 12)): wide (6, 12) 1.587, (8, 8) 1.580, (8, 12) 1.568, (8, 16) 1.569, (10, 20) 1.559,
// (12, 16) 1.562, (12, 20) 1.552, (12, 24) 1.552, (16, 20) 1.549, (16, 24) 1.548, (20, 24) 1.556 ms; then the binary pair under wide (16, 24):
// (8, 10) 1.545, (8, 8) 1.543, (6, 8) 1.541, (8, 6) 1.540 (all within the noise of each other): incoherent rays want to leave the node
// phase early and refill early, like the two-level kernels.
#ifndef BPT_MIN_NODE_LANES_W
#define BPT_MIN_NODE_LANES_W 16
#endif
#ifndef BPT_REFILL_W
#define BPT_REFILL_W 24
#endif
constexpr int kMinNodeLanesW = BPT_MIN_NODE_LANES_W, kRefillThresholdW = BPT_REFILL_W;
// Triangle phase: 0 = a lane whose postponed leaves are done keeps testing the triangles it pops one after the other (a run of leaf
// children of a wide node) while the other lanes wait; 1 = it only postpones the next (up to two) and goes back to the node phase with the warp.
// Measured (profiles/r2p_variants.jsonl): 1 is 4.5 % slower on the whole frame (1.640 vs 1.569 ms), a third postponed leaf (BPT_LEAF3) 1.4 %
// slower (1.591), both together 6.6 % (1.673): the short triangle runs are cheaper than another trip through the node phase.
#ifndef BPT_TRI_LOOP
#define BPT_TRI_LOOP 0
#endif
// Postponed leaves per lane: 2 (leaf, leaf2) or 3 (A/B builds)
#ifndef BPT_LEAF3
#define BPT_LEAF3 0
#endif
// Wide step, pushing the hit children that are not visited next: 1 = predicated stores (only the kept children, ~0.9 per step, cost an L1
// request), 0 = four unconditional stores above the top of which the kept ones stay (same instruction count). The wide kernels run the L1 at
// 77-88 % of its throughput (profiles/r2w_kernels.md), so requests matter: 1 is 1.3 % faster on configs[1], 3 % on the two-level scenes
// (profiles/r2x_variants.jsonl).
#ifndef BPT_WIDE_PUSH_PRED
#define BPT_WIDE_PUSH_PRED 1
#endif
// Ray loads / hit stores of the traversal kernels with the streaming (evict-first) policy: each is touched once, the L1 / L2 lines are
// worth more to the BVH (+0.4 % on configs[1], profiles/r2ab_variants.jsonl)
// Packet traversal (k_trace_packet): resident blocks per SM
#ifndef BPT_PACKET_MIN_BLOCKS
#define BPT_PACKET_MIN_BLOCKS 10
#endif
#ifndef BPT_STREAM_RAYS
#define BPT_STREAM_RAYS 1
#endif
constexpr int kMinNodeLanes2 = BPT_MIN_NODE_LANES_2L, kRefillThreshold2 = BPT_REFILL_2L;

struct RenderArgs {
    DScene sc;
    ShadeParams sp;
    bpt_camera cam;
    float4 *ray_o_in, *ray_d_in, *ray_w_in;
    float4 *ray_o_out, *ray_d_out, *ray_w_out;
    float4* hit; uint32_t* hit_slot;
    float4 *sh_o, *sh_d, *sh_c;
    float4* accum;
    float4* color;            // per-sample colour, [slot][pixel]
    float4* contrib;          // where light terms are added: `color` (fp32 state) or the per-bounce sum `bcol` (reference_fp16)
    uint32_t* qcount;
    uint64_t shadow_capacity;
    uint32_t npx, nslots;     // pixels, samples in this wave; a path's id is slot * npx + pixel
    uint32_t frame_base;      // frame_index of slot 0
    const float4* m_nodes; const float4* m_tris; int32_t m_root; uint32_t m_n;   // the single BVH of merged mode
    const float4* m_wide; const float4* m_leafbox;                                // its 4-wide quantised form (bpt_wide.cuh)
    uint32_t pixel_base;      // probe tracing in chunks: global path id = pixel_base + local id (RNG key)
    uint32_t pixel_jitter;
    uint32_t probe_mode;      // 1: paths start at probes; colour.w receives the first hit distance
    uint32_t cull_non_opaque; // connect kernel only: RAY_FLAG_CULL_NON_OPAQUE (the RTAO rays)
    uint32_t camera_paths;    // 1: the queue of bounce 1 was written by k_raygen (32 consecutive entries = the camera rays of one 8x4 pixel tile)
};

// warp-aggregated append: returns the slot for this lane (valid only if `emit`)
__device__ __forceinline__ uint32_t queue_push(uint32_t* counter, bool emit) {
    uint32_t active = __activemask();
    uint32_t ballot = __ballot_sync(active, emit);
    if (!emit) return 0;
    uint32_t lane = threadIdx.x & 31;
    uint32_t leader = __ffs(ballot) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(ballot));
    base = __shfl_sync(ballot, base, leader);
    return base + __popc(ballot & ((1u << lane) - 1u));
}

// ---- raygen (generate_camera_ray.hlsl:4-16): one thread per (sample slot, pixel), weight = 1, no jitter;
//      also zeroes the path's per-sample colour -----------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_raygen(const __grid_constant__ RenderArgs a) {
    // Thread → pixel through 8x4 tiles (a warp = one tile) so that the rays of a warp, and the hit points and
    // shadow rays they spawn, stay spatially coherent. The pixel's identity (RNG key, colour slot) is unchanged.
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t tiles_x = (a.sp.width + 7) / 8, tiles_y = (a.sp.height + 3) / 4;
    const uint32_t per_slot = tiles_x * tiles_y * 32;
    if (tid == 0) a.qcount[QE + 1] = a.npx * a.nslots;
    if (tid >= per_slot * a.nslots) return;
    uint32_t slot = tid / per_slot, r = tid % per_slot;
    uint32_t tile = r / 32, in_tile = r % 32;
    uint32_t px = (tile % tiles_x) * 8 + (in_tile & 7), py = (tile / tiles_x) * 4 + (in_tile >> 3);
    bool valid = px < a.sp.width && py < a.sp.height;
    // compact the valid pixels of the wave to consecutive queue slots (edge tiles may be partial)
    uint32_t p = py * a.sp.width + px;
    uint32_t id = slot * a.npx + p;                                          // path id = slot * npx + pixel
    // full tiles everywhere (e.g. 1920x1080): the slot is the thread id, no atomic; otherwise compact the
    // valid pixels with the warp-aggregated push (scratch counter, zeroed with the queue counters)
    const bool exact = (a.sp.width % 8 == 0) && (a.sp.height % 4 == 0);
    uint32_t qslot = exact ? tid : queue_push(&a.qcount[QN - 1], valid);
    if (!valid) return;
    float3 O, D;
    camera_ray(a.cam, px, py, a.sp.width, a.sp.height, a.pixel_jitter, a.frame_base + slot, O, D);
    if (a.sp.state_precision == BPT_STATE_REFERENCE_FP16) { D = q_half3(D); a.contrib[id] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); }   // ray_directions is rgba16_sfloat
    a.ray_o_out[qslot] = make_float4(O.x, O.y, O.z, __uint_as_float(id));
    a.ray_d_out[qslot] = make_float4(D.x, D.y, D.z, 0.0f);
    a.ray_w_out[qslot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    a.color[id] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- probe raygen (ddgi/trace_gbuffer.hlsl:10-36): one thread per (probe, ray) of the chunk -------
__global__ void __launch_bounds__(kBlock) k_probe_raygen(const __grid_constant__ RenderArgs a, const __grid_constant__ bpt_probe_volume vol,
                                                         const float2* __restrict__ sample_table) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id == 0) a.qcount[QE + 1] = a.npx;
    if (id >= a.npx) return;
    float3 O, D;
    probe_ray(vol, sample_table, a.pixel_base + id, a.frame_base, O, D);
    a.ray_o_out[id] = make_float4(O.x, O.y, O.z, __uint_as_float(id));
    a.ray_d_out[id] = make_float4(D.x, D.y, D.z, 0.0f);
    a.ray_w_out[id] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    a.color[id] = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
}

// ---- persistent speculative while-while traversal (extend: rt_gbuffer.hlsl:7-36; connect: NEW shadow rays) ---
// ncu on the one-thread-per-ray version showed 9.5/32 active lanes on incoherent bounces: rays of a warp finish at
// very different times and node/leaf steps diverge. So: one resident grid; every warp keeps 32 traversal state
// machines, finished lanes are written out and refilled from the queue (one atomicAdd per warp) once fewer than
// kRefillThreshold rays are in flight. Results do not depend on the order (tie-break + cull margin, bpt_trace.cuh).
// Speculative while-while (Aila & Laine 2009). Merged mode: one world-space BVH, the per-lane state is just
// (ray, node, postponed leaves, stack). A lane that reaches a leaf POSTPONES it and keeps descending; the warp switches to the triangle
// phase only when no lane is still searching, so both phases run with (nearly) all live lanes — the
// one-step-at-a-time loop above measured 9.7/32 active lanes on incoherent bounces.
constexpr int32_t kEmpty = 0x7fffffff;      // no internal node left for this lane
// TWO_LEVEL: TLAS leaves (instances) are entered at once — push kSentinel, switch to the object-space ray,
// continue at the BLAS root — and only triangles are postponed. A lane never leaves an instance (pops the
// sentinel) while it still holds postponed triangles of that instance: they need its object-space ray.
// Resident blocks per SM of the traversal kernels = the register budget: 8 -> 64 registers, 10 -> 48 (a few bytes of spills), 12 -> 40
// (1 KB of spills). The two-level kernels carry more state and stay at 8 (at 10: atrium two-level 1427 -> 1089, instanced 680 -> 536
// Mrays/s). Merged mode with the binary tree only: 8 was best (10: extend -2 %, connect +4 %). With the wide tree and the re-tuned phase
// thresholds, whole frame: (binary, wide) = (8, 8) 1.690, (8, 10) 1.680, (8, 12) 1.700, (9, 9) 1.666, (10, 10) 1.665, (10, 12) 1.685 ms.
#ifndef BPT_TRACE_MIN_BLOCKS
#define BPT_TRACE_MIN_BLOCKS 10
#endif
#ifndef BPT_TRACE_MIN_BLOCKS_2L
#define BPT_TRACE_MIN_BLOCKS_2L 8       // two-level kernels of scenes without any-hit instances (the general ones stay at 8: see above); final kernels on the
                                        // instanced scene: 7 blocks (72 registers) 876, 8 (64) 920, 9 (56, 64 B more spills) 867 Mrays/s (profiles/r2am_variants.jsonl)
#endif
#ifndef BPT_TRACE_MIN_BLOCKS_WIDE
#define BPT_TRACE_MIN_BLOCKS_WIDE 10
#endif
// WIDE: the node phase steps through the 4-wide quantised trees (merged: a.m_wide; two-level: the TLAS's and every BLAS's own) instead
// of the binary ones, and a proposed leaf — a triangle, or an instance of the TLAS — is taken only if the ray passes that leaf's exact
// box (bpt_wide.cuh: same hits, about half the node fetches).
// AH = false: no instance of the scene needs the any-hit opacity rule (bpt_trace.cuh) — the usual case; the kernel then carries no material code.
template <bool ANY, bool TWO_LEVEL, bool WIDE = false, bool AH = true>
__global__ void __launch_bounds__(kBlock, TWO_LEVEL ? (AH ? 8 : BPT_TRACE_MIN_BLOCKS_2L) : (WIDE ? BPT_TRACE_MIN_BLOCKS_WIDE : BPT_TRACE_MIN_BLOCKS)) k_trace_spec(const __grid_constant__ RenderArgs a, uint32_t bounce) {
    const uint32_t n = ANY ? (uint32_t)min((uint64_t)a.qcount[QS + bounce], a.shadow_capacity) : a.qcount[QE + bounce];
    uint32_t* cursor = &a.qcount[(ANY ? QWS : QWE) + bounce];
    const float4* __restrict__ qo = ANY ? a.sh_o : a.ray_o_in;
    const float4* __restrict__ qd = ANY ? a.sh_d : a.ray_d_in;
    const float4* const tlas_nodes = WIDE ? a.sc.tlas_wide : a.sc.tlas_nodes;
    const float4* __restrict__ nodes = TWO_LEVEL ? tlas_nodes : (WIDE ? a.m_wide : a.m_nodes);
    const float4* __restrict__ tris = TWO_LEVEL ? nullptr : a.m_tris;
    const float4* __restrict__ leafbox = nullptr;       // TWO_LEVEL && WIDE: exact leaf boxes of the BLAS being traversed (nullptr: single-leaf BLAS)
    const uint32_t lane = threadIdx.x & 31;
    constexpr int kMinNodeLanes = TWO_LEVEL ? kMinNodeLanes2 : (WIDE ? kMinNodeLanesW : kMinNodeLanes1);
    constexpr int kRefillThreshold = TWO_LEVEL ? kRefillThreshold2 : (WIDE ? kRefillThresholdW : kRefillThreshold1);
    // Short stack in shared memory (north star): the BOTTOM kSmemStack entries of every lane's stack live in shared memory as
    // [entry][thread] (one bank per lane: conflict-free whatever the lanes' depths are), deeper entries in local memory. A push or pop
    // at depth d touches exactly one of the two, so the local-memory traffic that is left is the accesses at depth >= kSmemStack.
    constexpr int S = kSmemStack;
    __shared__ int32_t s_stack[S > 0 ? S : 1][kBlock];
    int32_t stack[((TWO_LEVEL && WIDE) ? 2 * kStackSize : kStackSize) - S];     // two wide trees deep
    auto st_put = [&](int i, int32_t v) { if (S > 0 && i < S) s_stack[i][threadIdx.x] = v; else stack[i - S] = v; };
    auto st_get = [&](int i) -> int32_t { return (S > 0 && i < S) ? s_stack[i][threadIdx.x] : stack[i - S]; };
    RayState rs;
    RaySpace sp_;
    rs.found = false; rs.tbest = 0.0f; rs.tcull = 0.0f; rs.tmin = 0.001f; rs.cull_non_opaque = ANY && a.cull_non_opaque != 0;
    int32_t node = kEmpty, leaf = 0, leaf2 = 0, leaf3 = 0;
    constexpr bool kLeaf3 = BPT_LEAF3 != 0;
    auto leaf_room = [&]() { return kLeaf3 ? leaf3 == 0 : leaf2 == 0; };
    auto leaf_put = [&](int32_t v) { if (leaf == 0) leaf = v; else if (!kLeaf3 || leaf2 == 0) leaf2 = v; else leaf3 = v; };
    // The top of the stack lives in a register (`tos`, kEmpty = nothing left): a pop hands out `tos` at once and the load of
    // the entry below it overlaps the node fetch instead of preceding it (ncu: 9 % of the stall samples sat behind that load).
    int32_t tos = kEmpty;
    int sp = 0;
    // (the wide kernels push up to three nodes per step with predicated stores and keep a plain stack)
    // (A lazily spilled register top for the wide kernels — the last push stays in a register until another push follows — was measured:
    // fewer L1 requests, but 3-5 % slower, profiles/r2ab_variants.jsonl: the extra predicated moves cost more than the requests saved.)
    auto push = [&](int32_t v) { if (WIDE) { st_put(sp++, v); } else { st_put(sp++, tos); tos = v; } };
    auto pop = [&]() { if (WIDE) return sp ? st_get(--sp) : kEmpty; int32_t v = tos; tos = sp ? st_get(--sp) : kEmpty; return v; };
    uint32_t ray = 0xffffffffu, path = 0;
    uint32_t slot = 0xffffffffu, inst_anyhit = 0;     // TWO_LEVEL: the instance being traversed
    bool in_blas = !TWO_LEVEL;
    bool exhausted = false;

    constexpr bool PARK = TWO_LEVEL && (BPT_PARK_INSTANCES != 0);
    float3 widir = v3s(0.0f);                           // 1 / D of the world-space ray (PARK: leaving an instance restores the ray space without dividing again)
    // a TLAS leaf the lane stands on (two-level mode): an instance to enter
    auto at_instance = [&]() { return TWO_LEVEL && node < 0 && node != kSentinel && !in_blas; };
    // node is a triangle leaf of the tree being traversed
    auto at_triangle = [&]() { return node < 0 && node != kSentinel && (!TWO_LEVEL || in_blas); };
    // Enters the instance of the TLAS leaf `node` (or skips it when only a quantised box proposed it): the object-space ray, a sentinel on the stack.
    auto enter_instance = [&]() {
        if (WIDE && a.sc.tlas_n != 1 && !leaf_box_hit(a.sc.tlas_leafbox, (uint32_t)~node, sp_, rs.tmin, rs.tcull)) {
            node = pop();                              // the binary TLAS would not have reached it
            return;
        }
        slot = __ldg(a.sc.tlas_prims + (uint32_t)~node);
        const DInstance& in = a.sc.instances[slot];
        const DBlas bl = a.sc.blas[in.blas];
        inst_anyhit = in.anyhit;
        sp_ = make_space(xf_point(in.w2o, rs.O), xf_vector(in.w2o, rs.D));
        nodes = WIDE ? bl.wide : bl.nodes; tris = bl.tris; in_blas = true;
        if (WIDE) leafbox = bl.leafbox;
        push(kSentinel);
        node = bl.root;
    };
    // Brings `node` into a state the node phase understands: leaves instances when allowed; enters them at once (PARK = false) or leaves the
    // lane parked on the TLAS leaf for the entry phase (PARK = true).
    auto settle = [&]() {
        if (!TWO_LEVEL) return;
        for (;;) {
            if (node == kSentinel) {
                if (leaf != 0) return;                 // postponed triangles of this instance first
                if (PARK) { sp_.O = rs.O; sp_.D = rs.D; sp_.idir = widir; sp_.ood = rs.O * widir; }      // = make_space(rs.O, rs.D), bit for bit
                else sp_ = make_space(rs.O, rs.D);
                nodes = tlas_nodes; tris = nullptr; in_blas = false;
                node = pop();
                continue;
            }
            if (!PARK && at_instance()) { enter_instance(); continue; }
            return;
        }
    };

    for (;;) {
        if (ray != 0xffffffffu && node == kEmpty && leaf == 0) {          // finished: write out
            if (ANY) {
                if (!rs.found) {
                    float4 c = a.sh_c[ray];
                    float* px = reinterpret_cast<float*>(a.contrib + path);
                    atomicAdd(px + 0, c.x); atomicAdd(px + 1, c.y); atomicAdd(px + 2, c.z);
                }
            } else {
#if BPT_STREAM_RAYS
                __stcs(a.hit + ray, make_float4(rs.found ? rs.tbest : -1.0f, rs.bu, rs.bv, __uint_as_float(rs.best_prim)));
                __stcs(a.hit_slot + ray, rs.best_slot);
#else
                a.hit[ray] = make_float4(rs.found ? rs.tbest : -1.0f, rs.bu, rs.bv, __uint_as_float(rs.best_prim));
                a.hit_slot[ray] = rs.best_slot;
#endif
            }
            ray = 0xffffffffu;
        }
        if (!exhausted) {                                                  // refill idle lanes, one atomic per warp
            bool want = ray == 0xffffffffu;
            uint32_t mask = __ballot_sync(0xffffffffu, want);
            if (mask) {
                uint32_t base = 0;
                if (lane == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(cursor, (uint32_t)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
                if (want) {
                    uint32_t idx = base + __popc(mask & ((1u << lane) - 1u));
                    if (idx < n) {
#if BPT_STREAM_RAYS
                        float4 o = __ldcs(qo + idx), d = __ldcs(qd + idx);      // read once: do not displace BVH lines
#else
                        float4 o = qo[idx], d = qd[idx];
#endif
                        ray = idx;
                        path = __float_as_uint(o.w);
                        rs.O = v3(o.x, o.y, o.z); rs.D = v3(d.x, d.y, d.z);
                        rs.tmin = 0.001f; rs.tbest = ANY ? d.w : a.sp.ray_length; rs.tcull = rs.tbest * 1.00001f;
                        rs.best_slot = 0xffffffffu; rs.best_prim = 0xffffffffu; rs.bu = 0.0f; rs.bv = 0.0f;
                        rs.frame_index = a.frame_base + path / a.npx; rs.opacity_u = 0.0f; rs.have_u = false; rs.found = false;
                        sp_ = make_space(rs.O, rs.D);
                        widir = sp_.idir;
                        sp = 0; tos = kEmpty; leaf = 0; leaf2 = 0; leaf3 = 0;
                        if (TWO_LEVEL) {
                            nodes = tlas_nodes; tris = nullptr; in_blas = false; slot = 0xffffffffu;
                            node = a.sc.tlas_n == 0 ? kEmpty : a.sc.tlas_root;
                            settle();
                        } else {
                            node = a.m_n == 0 ? kEmpty : a.m_root;
                        }
                        if (at_triangle()) { leaf = node; node = pop(); settle(); }   // root is a leaf
                    }
                }
                if (base + (uint32_t)__popc(mask) >= n) exhausted = true;  // warp-uniform
            }
        }
        if (__ballot_sync(0xffffffffu, ray != 0xffffffffu) == 0) break;
        for (;;) {
            // ---- node phase: until no lane is still searching for its first leaf ----
            // A lane postpones up to two leaves (leaf, leaf2) and keeps descending; it only idles when a third shows up.
            for (;;) {
                if (node >= 0 && node != kEmpty) {
                    int32_t next;
                    if (WIDE) {
                        int32_t ch[4]; uint32_t hitmask; int best;
                        node_test4q<!TWO_LEVEL>(nodes, node, sp_, rs.tmin, rs.tcull, ch, hitmask, best);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const bool keep = ((hitmask >> k) & 1u) && k != best;
#if BPT_WIDE_PUSH_PRED
                            if (keep) st_put(sp, ch[k]);                    // predicated store: only the kept children cost an L1 transaction
#else
                            st_put(sp, ch[k]);                              // store above the top unconditionally, keep it only if wanted
#endif
                            sp += keep ? 1 : 0;
                        }
                        next = best < 0 ? BPT_POP : (best == 0 ? ch[0] : (best == 1 ? ch[1] : (best == 2 ? ch[2] : ch[3])));
                    } else {
                        int32_t far;
                        next = node_step2<!TWO_LEVEL>(nodes, node, sp_, rs.tmin, rs.tcull, far);
                        if (far != BPT_POP) push(far);
                    }
                    if (next == BPT_POP) next = pop();
                    node = next;
                    settle();
                    if (at_triangle() && leaf_room()) {                   // a triangle: postpone, keep descending
                        leaf_put(node);
#ifdef BPT_PREFETCH_TRI
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(tris + 3 * (size_t)(uint32_t)~node));
#endif
                        node = pop();
                        settle();
                    }
                }
                const bool internal = node >= 0 && node != kEmpty;
                const uint32_t can_work = __ballot_sync(0xffffffffu, internal);
                const uint32_t searching = __ballot_sync(0xffffffffu, internal && leaf == 0);
                // leave the node phase when nobody is still looking for a first leaf, or when too few lanes have node
                // work left (the others idle with postponed leaves or finished rays): test triangles / refill instead
                if (searching == 0 || __popc(can_work) < kMinNodeLanes) break;
            }
            // ---- instance-entry phase (two-level, PARK): every lane parked on a TLAS leaf enters its instance now, together ----
            if (PARK) {
                const bool parked = at_instance();
                if (__ballot_sync(0xffffffffu, parked) != 0u) {
                    if (parked) {
                        enter_instance();
                        settle();                                        // (a skipped instance may have popped a sentinel)
                        if (at_triangle() && leaf_room()) {              // a single-triangle BLAS, or a triangle popped after a skip
                            leaf_put(node);
                            node = pop();
                            settle();
                        }
                    }
                    continue;                                            // back to the node phase: the entered lanes are searching again
                }
            }
            // ---- triangle phase ----
            while (leaf != 0) {
                bool accepted;
                if (WIDE && TWO_LEVEL) {
                    const uint32_t j = (uint32_t)~leaf;
                    const float4* tp = tris + 3 * (size_t)j;
                    float4 ta = BPT_LDG(tp), tb = BPT_LDG(tp + 1), tc = BPT_LDG(tp + 2);
                    float4 blo = ta, bhi = ta;
                    if (leafbox) ldg_32B<false>(leafbox + 2 * (size_t)j, blo, bhi);
                    accepted = (!leafbox || leaf_box_hit_rec(blo, bhi, sp_, rs.tmin, rs.tcull)) &&
                               test_triangle_rec<ANY, AH>(a.sc, rs, ta, tb, tc, sp_.O, sp_.D, slot, inst_anyhit);
                } else if (WIDE) {                                           // leaf box and triangle fetched together: one latency, not two
                    const uint32_t j = (uint32_t)~leaf;
                    const float4* tp = tris + 3 * (size_t)j;
                    const float4* bp = a.m_leafbox + 2 * (size_t)(a.m_n == 1 ? 0u : j);
                    float4 ta = BPT_LDG(tp), tb = BPT_LDG(tp + 1), tc = BPT_LDG(tp + 2);
                    float4 blo = ta, bhi = ta;
                    if (a.m_n != 1) ldg_32B(bp, blo, bhi);
                    accepted = (a.m_n == 1 || leaf_box_hit_rec(blo, bhi, sp_, rs.tmin, rs.tcull)) &&
                               test_triangle_rec<ANY, AH>(a.sc, rs, ta, tb, tc, sp_.O, sp_.D, 0xffffffffu, 0u);
                } else {
                    const float4* tp = tris + 3 * (size_t)(uint32_t)~leaf;
                    float4 ta = BPT_LDG(tp), tb = BPT_LDG(tp + 1), tc = BPT_LDG(tp + 2);
                    accepted = test_triangle_rec<ANY, AH>(a.sc, rs, ta, tb, tc, sp_.O, sp_.D, TWO_LEVEL ? slot : 0xffffffffu, inst_anyhit);
                }
                leaf = leaf2; leaf2 = kLeaf3 ? leaf3 : 0; leaf3 = 0;
                if (ANY && accepted) { node = kEmpty; sp = 0; tos = kEmpty; leaf = 0; leaf2 = 0; break; }
                if (leaf == 0) {
                    settle();                                            // a sentinel that was waiting for the triangles
                    if (at_triangle()) {
                        leaf = node; node = pop(); settle();
#if BPT_TRI_LOOP == 1
                        if (at_triangle()) { leaf2 = node; node = pop(); settle(); }
                        break;
#endif
                    }
                }
            }
            uint32_t alive = __ballot_sync(0xffffffffu, node != kEmpty);
            if (alive == 0) break;
            if (!exhausted && __popc(alive) < kRefillThreshold) break;
        }
    }
}

// the instantiations (macro arguments cannot carry the template commas); *_o: scenes without any-hit instances
static const auto k_extend_merged = k_trace_spec<false, false>;
static const auto k_connect_merged = k_trace_spec<true, false>;
static const auto k_extend_wide = k_trace_spec<false, false, true>;
static const auto k_connect_wide = k_trace_spec<true, false, true>;
static const auto k_extend_two_level = k_trace_spec<false, true>;
static const auto k_connect_two_level = k_trace_spec<true, true>;
static const auto k_extend_two_level_wide = k_trace_spec<false, true, true>;
static const auto k_connect_two_level_wide = k_trace_spec<true, true, true>;
static const auto k_extend_merged_o = k_trace_spec<false, false, false, false>;
static const auto k_connect_merged_o = k_trace_spec<true, false, false, false>;
static const auto k_extend_wide_o = k_trace_spec<false, false, true, false>;
static const auto k_connect_wide_o = k_trace_spec<true, false, true, false>;
static const auto k_extend_two_level_o = k_trace_spec<false, true, false, false>;
static const auto k_connect_two_level_o = k_trace_spec<true, true, false, false>;
static const auto k_extend_two_level_wide_o = k_trace_spec<false, true, true, false>;
static const auto k_connect_two_level_wide_o = k_trace_spec<true, true, true, false>;

// ---- packet traversal of coherent rays (merged mode, binary tree): camera rays, and the shadow rays of the camera rays' hit points ----------
// The persistent kernel above gives every lane its own traversal; for the camera rays of an 8x4 pixel tile that is 32 walks through nearly
// the same nodes: measured on the binary BVH of configs[1] over 300 tiles, a ray visits 37.9 nodes and 2.3 leaves, the UNION over the tile is 44.6
// nodes and 5.5 leaves. Here one warp walks that union ONCE: 32 consecutive queue entries are a packet, the warp keeps one stack of
// (node, lane mask) in shared memory, every lane tests both child boxes of the same node (one broadcast fetch instead of 32), a ballot
// gives the lanes that enter each child, the child more lanes prefer is visited first and the other pushed with its mask; at a leaf
// the lanes of its mask test the triangle. A lane tests a triangle iff its own ray passed every ancestor's box with its own current
// cull distance — the candidates of its own traversal in another order, and results do not depend on the order (bpt_trace.cuh header):
// hits are bit-identical to k_trace_spec's. Correct for ANY 32 rays; fast when they are coherent, so it is launched for bounce 1 of
// camera paths only (RenderArgs::camera_paths). Measured on configs[1] (profiles/r2ad_variants.jsonl, r2ae_packet_source.md): the camera
// rays' extend launch 5.24 -> 4.54 ms per 16-sample wave (3.9 G instead of 4.9 G warp instructions at 29.8 instead of 24 lanes); what
// binds it now is the L1 data path (every lane still receives every 64-B node: 76 %) together with the ALU pipe (min / max: 74 %).
// The shadow rays of the camera rays' hit points as packets: slower (connect 0.209 -> 0.227 ms; mixed lights 9.3 -> 10.8), off.
// A FRUSTUM walk instead of per-lane box tests (twelve lanes test the twelve (child, plane) pairs of a node against the packet's four
// side planes and depth range, every lane runs the exact leaf-box + triangle test at the leaves — valid because a ray passes a leaf's
// exact box only if it passes every ancestor's) was built and is bit-exact too, but visits 3x the nodes (336 instead of 107 stack
// accesses per packet: its depth pruning is the packet's WORST cull distance) and is slower than k_trace_spec:
// profiles/r2ag_frustum_kernel.patch, r2ag_frustum_source.md, r2af_variants.jsonl.
constexpr int kPacketStack = 100;      // one entry per tree level at most (Karras depth bound 95, bpt_trace.cuh)
template <bool ANY, bool AH>
__global__ void __launch_bounds__(kBlock, BPT_PACKET_MIN_BLOCKS) k_trace_packet(const __grid_constant__ RenderArgs a, uint32_t bounce) {
    const uint32_t n = ANY ? (uint32_t)min((uint64_t)a.qcount[QS + bounce], a.shadow_capacity) : a.qcount[QE + bounce];
    uint32_t* cursor = &a.qcount[(ANY ? QWS : QWE) + bounce];
    const float4* __restrict__ qo = ANY ? a.sh_o : a.ray_o_in;
    const float4* __restrict__ qd = ANY ? a.sh_d : a.ray_d_in;
    const float4* __restrict__ nodes = a.m_nodes;
    const float4* __restrict__ tris = a.m_tris;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int32_t s_node_all[kBlock / 32][kPacketStack];
    __shared__ uint32_t s_mask_all[kBlock / 32][kPacketStack];
    int32_t* const s_node = s_node_all[warp];
    uint32_t* const s_mask = s_mask_all[warp];
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t idx = base + lane;
        const bool valid = idx < n;
        RayState rs;
        RaySpace sp_;
        uint32_t path = 0;
        {
            float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
            if (valid) { o = __ldcs(qo + idx); d = __ldcs(qd + idx); }
            path = __float_as_uint(o.w);
            rs.O = v3(o.x, o.y, o.z); rs.D = v3(d.x, d.y, d.z);
            rs.tmin = 0.001f; rs.tbest = ANY ? d.w : a.sp.ray_length; rs.tcull = rs.tbest * 1.00001f;
            rs.best_slot = 0xffffffffu; rs.best_prim = 0xffffffffu; rs.bu = 0.0f; rs.bv = 0.0f;
            rs.frame_index = a.frame_base + path / a.npx; rs.opacity_u = 0.0f; rs.have_u = false; rs.found = false;
            rs.cull_non_opaque = ANY && a.cull_non_opaque != 0;
            sp_ = make_space(rs.O, rs.D);
        }
        bool alive = valid;                                  // ANY: cleared by the first accepted hit
        int sp = 0;
        int32_t node = a.m_n == 0 ? kEmpty : a.m_root;
        uint32_t mask = __ballot_sync(0xffffffffu, alive);
        auto walk = [&](auto mode_tag) {
            constexpr int MODE = decltype(mode_tag)::value;
        while (node != kEmpty) {
            if (node >= 0) {
                const float4* np = nodes + 4 * (size_t)node;
                float4 n0, n1, n2, n3;
                ldg_64B(np, n0, n1, n2, n3);
                const RaySpace& r = sp_;
                float t0n, t0f, t1n, t1f;
                if (MODE == 8) {                                     // lanes disagree on a direction sign: node_step2's own expressions
                    float c0lox = fmaf(n0.x, r.idir.x, -r.ood.x), c0hix = fmaf(n0.y, r.idir.x, -r.ood.x);
                    float c0loy = fmaf(n0.z, r.idir.y, -r.ood.y), c0hiy = fmaf(n0.w, r.idir.y, -r.ood.y);
                    float c1lox = fmaf(n1.x, r.idir.x, -r.ood.x), c1hix = fmaf(n1.y, r.idir.x, -r.ood.x);
                    float c1loy = fmaf(n1.z, r.idir.y, -r.ood.y), c1hiy = fmaf(n1.w, r.idir.y, -r.ood.y);
                    float c0loz = fmaf(n2.x, r.idir.z, -r.ood.z), c0hiz = fmaf(n2.y, r.idir.z, -r.ood.z);
                    float c1loz = fmaf(n2.z, r.idir.z, -r.ood.z), c1hiz = fmaf(n2.w, r.idir.z, -r.ood.z);
                    t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), rs.tmin));
                    t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rs.tcull));
                    t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), rs.tmin));
                    t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rs.tcull));
                } else {
                    // The whole packet looks into ONE octant (camera rays of a tile nearly always do): the near plane of every slab is known
                    // at compile time, and for a valid box (lo <= hi) the slab value of the near plane IS min(t(lo), t(hi)) bit for bit (fmaf is
                    // monotone in the plane) — the twelve per-axis min / max of a node step disappear.
                    constexpr bool SX = (MODE & 1) != 0, SY = (MODE & 2) != 0, SZ = (MODE & 4) != 0;
                    const float a0nx = fmaf(SX ? n0.y : n0.x, r.idir.x, -r.ood.x), a0fx = fmaf(SX ? n0.x : n0.y, r.idir.x, -r.ood.x);
                    const float a0ny = fmaf(SY ? n0.w : n0.z, r.idir.y, -r.ood.y), a0fy = fmaf(SY ? n0.z : n0.w, r.idir.y, -r.ood.y);
                    const float a0nz = fmaf(SZ ? n2.y : n2.x, r.idir.z, -r.ood.z), a0fz = fmaf(SZ ? n2.x : n2.y, r.idir.z, -r.ood.z);
                    const float a1nx = fmaf(SX ? n1.y : n1.x, r.idir.x, -r.ood.x), a1fx = fmaf(SX ? n1.x : n1.y, r.idir.x, -r.ood.x);
                    const float a1ny = fmaf(SY ? n1.w : n1.z, r.idir.y, -r.ood.y), a1fy = fmaf(SY ? n1.z : n1.w, r.idir.y, -r.ood.y);
                    const float a1nz = fmaf(SZ ? n2.w : n2.z, r.idir.z, -r.ood.z), a1fz = fmaf(SZ ? n2.z : n2.w, r.idir.z, -r.ood.z);
                    t0n = fmaxf(fmaxf(a0nx, a0ny), fmaxf(a0nz, rs.tmin)); t0f = fminf(fminf(a0fx, a0fy), fminf(a0fz, rs.tcull));
                    t1n = fmaxf(fmaxf(a1nx, a1ny), fmaxf(a1nz, rs.tmin)); t1f = fminf(fminf(a1fx, a1fy), fminf(a1fz, rs.tcull));
                }
                const bool mine = alive && ((mask >> lane) & 1u);
                const bool h0 = mine && t0n <= t0f, h1 = mine && t1n <= t1f;
                const uint32_t m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                const int32_t ch0 = (int32_t)f2u(n3.x), ch1 = (int32_t)f2u(n3.y);
                if ((m0 | m1) == 0u) {
                    if (sp == 0) { node = kEmpty; continue; }
                    --sp; node = s_node[sp]; mask = s_mask[sp];
                } else if (m0 != 0u && m1 != 0u) {
                    // visit first the child more lanes reach first (a lane that enters one child only votes for that one)
                    const uint32_t pref0 = __ballot_sync(0xffffffffu, h0 && (!h1 || t0n <= t1n));
                    const bool first0 = 2 * __popc(pref0) >= __popc(m0 | m1);
                    s_node[sp] = first0 ? ch1 : ch0; s_mask[sp] = first0 ? m1 : m0; ++sp;
                    node = first0 ? ch0 : ch1; mask = first0 ? m0 : m1;
                } else {
                    node = m0 != 0u ? ch0 : ch1; mask = m0 | m1;
                }
            } else {
                if (alive && ((mask >> lane) & 1u)) {
                    const float4* tp = tris + 3 * (size_t)(uint32_t)~node;
                    float4 ta = BPT_LDG(tp), tb = BPT_LDG(tp + 1), tc = BPT_LDG(tp + 2);
                    const bool accepted = test_triangle_rec<ANY, AH>(a.sc, rs, ta, tb, tc, sp_.O, sp_.D, 0xffffffffu, 0u);
                    if (ANY && accepted) alive = false;
                }
                if (ANY && __ballot_sync(0xffffffffu, alive) == 0u) { node = kEmpty; continue; }
                if (sp == 0) { node = kEmpty; continue; }
                --sp; node = s_node[sp]; mask = s_mask[sp];
            }
        }
        };
        {
            // direction octant of the packet, if its lanes agree on it
            const uint32_t vm = __ballot_sync(0xffffffffu, valid);
            const uint32_t bx = __ballot_sync(0xffffffffu, valid && sp_.idir.x < 0.0f), by = __ballot_sync(0xffffffffu, valid && sp_.idir.y < 0.0f),
                           bz = __ballot_sync(0xffffffffu, valid && sp_.idir.z < 0.0f);
            const bool uniform = (bx == 0u || bx == vm) && (by == 0u || by == vm) && (bz == 0u || bz == vm);
            const int mode = uniform ? ((bx ? 1 : 0) | (by ? 2 : 0) | (bz ? 4 : 0)) : 8;
            switch (mode) {
                case 0: walk(std::integral_constant<int, 0>{}); break;
                case 1: walk(std::integral_constant<int, 1>{}); break;
                case 2: walk(std::integral_constant<int, 2>{}); break;
                case 3: walk(std::integral_constant<int, 3>{}); break;
                case 4: walk(std::integral_constant<int, 4>{}); break;
                case 5: walk(std::integral_constant<int, 5>{}); break;
                case 6: walk(std::integral_constant<int, 6>{}); break;
                case 7: walk(std::integral_constant<int, 7>{}); break;
                default: walk(std::integral_constant<int, 8>{}); break;
            }
        }
        if (valid) {
            if (ANY) {
                if (!rs.found) {
                    float4 c = a.sh_c[idx];
                    float* px = reinterpret_cast<float*>(a.contrib + path);
                    atomicAdd(px + 0, c.x); atomicAdd(px + 1, c.y); atomicAdd(px + 2, c.z);
                }
            } else {
                __stcs(a.hit + idx, make_float4(rs.found ? rs.tbest : -1.0f, rs.bu, rs.bv, __uint_as_float(rs.best_prim)));
                __stcs(a.hit_slot + idx, rs.best_slot);
            }
        }
    }
}

// ---- the same for two-level mode: the packet walks the TLAS with the world-space rays; at an instance the lanes of its mask move their
//      ray into object space TOGETHER (one instance record for the warp, the three IEEE divisions of make_space at full width instead of
//      the ~2 lanes an instance entry has in k_trace_spec) and the packet walks that BLAS; a sentinel on the stack brings everybody back
//      to world space. Binary trees, general box tests (the octant changes with every instance). Candidates per lane = those of its own
//      binary two-level traversal (= the wide one's, bpt_wide.cuh), hence the same hits. ---------------------------------------------------
template <bool AH>
__global__ void __launch_bounds__(kBlock, 8) k_extend_packet_2l(const __grid_constant__ RenderArgs a, uint32_t bounce) {
    const uint32_t n = a.qcount[QE + bounce];
    uint32_t* cursor = &a.qcount[QWE + bounce];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kDepth = 2 * kPacketStack;
    __shared__ int32_t s_node_all[kBlock / 32][kDepth];
    __shared__ uint32_t s_mask_all[kBlock / 32][kDepth];
    int32_t* const s_node = s_node_all[warp];
    uint32_t* const s_mask = s_mask_all[warp];
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t idx = base + lane;
        const bool valid = idx < n;
        RayState rs;
        RaySpace cur;
        float3 widir, wood;                                   // the world-space ray's 1 / D and O / D (restored when the packet leaves an instance)
        {
            float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f), d = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
            if (valid) { o = __ldcs(a.ray_o_in + idx); d = __ldcs(a.ray_d_in + idx); }
            const uint32_t path = __float_as_uint(o.w);
            rs.O = v3(o.x, o.y, o.z); rs.D = v3(d.x, d.y, d.z);
            rs.tmin = 0.001f; rs.tbest = a.sp.ray_length; rs.tcull = rs.tbest * 1.00001f;
            rs.best_slot = 0xffffffffu; rs.best_prim = 0xffffffffu; rs.bu = 0.0f; rs.bv = 0.0f;
            rs.frame_index = a.frame_base + path / a.npx; rs.opacity_u = 0.0f; rs.have_u = false; rs.found = false;
            rs.cull_non_opaque = false;
            cur = make_space(rs.O, rs.D);
            widir = cur.idir; wood = cur.ood;
        }
        const float4* nodes = a.sc.tlas_nodes;
        const float4* tris = nullptr;
        bool in_blas = false;                                  // warp-uniform
        uint32_t slot = 0xffffffffu, inst_anyhit = 0u;         // warp-uniform: the instance the packet is inside
        int sp = 0;
        int32_t node = a.sc.tlas_n == 0 ? kEmpty : a.sc.tlas_root;
        uint32_t mask = __ballot_sync(0xffffffffu, valid);
        auto pop = [&]() {
            if (sp == 0) { node = kEmpty; return; }
            --sp; node = s_node[sp]; mask = s_mask[sp];
        };
        while (node != kEmpty) {
            if (node == kSentinel) {                           // the BLAS is done: back to the world-space rays and the TLAS
                cur.O = rs.O; cur.D = rs.D; cur.idir = widir; cur.ood = wood;
                nodes = a.sc.tlas_nodes; tris = nullptr; in_blas = false; slot = 0xffffffffu;
                pop();
            } else if (node >= 0) {
                const float4* np = nodes + 4 * (size_t)node;
                float4 n0, n1, n2, n3;
                ldg_64B<false>(np, n0, n1, n2, n3);
                const RaySpace& r = cur;
                float c0lox = fmaf(n0.x, r.idir.x, -r.ood.x), c0hix = fmaf(n0.y, r.idir.x, -r.ood.x);
                float c0loy = fmaf(n0.z, r.idir.y, -r.ood.y), c0hiy = fmaf(n0.w, r.idir.y, -r.ood.y);
                float c1lox = fmaf(n1.x, r.idir.x, -r.ood.x), c1hix = fmaf(n1.y, r.idir.x, -r.ood.x);
                float c1loy = fmaf(n1.z, r.idir.y, -r.ood.y), c1hiy = fmaf(n1.w, r.idir.y, -r.ood.y);
                float c0loz = fmaf(n2.x, r.idir.z, -r.ood.z), c0hiz = fmaf(n2.y, r.idir.z, -r.ood.z);
                float c1loz = fmaf(n2.z, r.idir.z, -r.ood.z), c1hiz = fmaf(n2.w, r.idir.z, -r.ood.z);
                float t0n = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), rs.tmin));
                float t0f = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), rs.tcull));
                float t1n = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), rs.tmin));
                float t1f = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), rs.tcull));
                const bool mine = (mask >> lane) & 1u;
                const bool h0 = mine && t0n <= t0f, h1 = mine && t1n <= t1f;
                const uint32_t m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1);
                const int32_t ch0 = (int32_t)f2u(n3.x), ch1 = (int32_t)f2u(n3.y);
                if ((m0 | m1) == 0u) {
                    pop();
                } else if (m0 != 0u && m1 != 0u) {
                    const uint32_t pref0 = __ballot_sync(0xffffffffu, h0 && (!h1 || t0n <= t1n));
                    const bool first0 = 2 * __popc(pref0) >= __popc(m0 | m1);
                    s_node[sp] = first0 ? ch1 : ch0; s_mask[sp] = first0 ? m1 : m0; ++sp;
                    node = first0 ? ch0 : ch1; mask = first0 ? m0 : m1;
                } else {
                    node = m0 != 0u ? ch0 : ch1; mask = m0 | m1;
                }
            } else if (!in_blas) {                             // a TLAS leaf: the lanes of `mask` enter the instance
                slot = __ldg(a.sc.tlas_prims + (uint32_t)~node);
                const DInstance& in = a.sc.instances[slot];
                const DBlas bl = a.sc.blas[in.blas];
                inst_anyhit = in.anyhit;
                if ((mask >> lane) & 1u) cur = make_space(xf_point(in.w2o, rs.O), xf_vector(in.w2o, rs.D));
                s_node[sp] = kSentinel; s_mask[sp] = 0u; ++sp;
                nodes = bl.nodes; tris = bl.tris; in_blas = true;
                node = bl.n == 0 ? kSentinel : bl.root;        // (an empty BLAS: straight back out)
                if (bl.n == 0) --sp;
            } else {                                           // a triangle of the instance
                if ((mask >> lane) & 1u) {
                    const float4* tp = tris + 3 * (size_t)(uint32_t)~node;
                    float4 ta = BPT_LDG(tp), tb = BPT_LDG(tp + 1), tc = BPT_LDG(tp + 2);
                    test_triangle_rec<false, AH>(a.sc, rs, ta, tb, tc, cur.O, cur.D, slot, inst_anyhit);
                }
                pop();
            }
        }
        if (valid) {
            __stcs(a.hit + idx, make_float4(rs.found ? rs.tbest : -1.0f, rs.bu, rs.bv, __uint_as_float(rs.best_prim)));
            __stcs(a.hit_slot + idx, rs.best_slot);
        }
    }
}
static const auto k_extend_packet_2l_a = k_extend_packet_2l<true>;
static const auto k_extend_packet_2l_o = k_extend_packet_2l<false>;
static const auto k_extend_packet = k_trace_packet<false, true>;
static const auto k_extend_packet_o = k_trace_packet<false, false>;
static const auto k_connect_packet = k_trace_packet<true, true>;
static const auto k_connect_packet_o = k_trace_packet<true, false>;

// ---- shade: material + lighting + next direction, emits shadow rays and the next extend ray ---
struct KernelSink {
    const RenderArgs& a;
    uint32_t bounce, path;      // path = slot * npx + pixel
    __device__ void add(float3 c) {
        float4 v = a.contrib[path];
        v.x += c.x; v.y += c.y; v.z += c.z;
        a.contrib[path] = v;
    }
    __device__ void shadow(float3 P, float3 L, float tmax, float3 c, uint32_t light) {
        if (a.sp.nee_mode == BPT_NEE_NONE) { add(c); return; }
        uint32_t slot = queue_push(&a.qcount[QS + bounce], true);
        if (slot < a.shadow_capacity) {
            a.sh_o[slot] = make_float4(P.x, P.y, P.z, __uint_as_float(path));
            a.sh_d[slot] = make_float4(L.x, L.y, L.z, tmax);
            a.sh_c[slot] = make_float4(c.x, c.y, c.z, __uint_as_float(light));
        }
    }
};

#ifndef BPT_SHADE_MIN_BLOCKS
#define BPT_SHADE_MIN_BLOCKS 8      // 64 registers; 6 (80) and 5 (96) measured: see DESIGN section 5
#endif
// RECT = false: the scene has no rect lights — the LTC evaluation (LUT fetches, clipping, edge integrals, light textures) is compiled out
template <bool IBL, bool RECT = true, bool GENERAL = true>
__global__ void __launch_bounds__(kBlock, BPT_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ RenderArgs a, uint32_t bounce) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i < a.qcount[QE + bounce];
    bool cont = false;
    float3 nO = v3s(0.0f), nD = v3s(0.0f), nW = v3s(0.0f);
    uint32_t path = 0;
    if (live) {
        float4 o = a.ray_o_in[i], d = a.ray_d_in[i], w = a.ray_w_in[i], h = a.hit[i];
        path = __float_as_uint(o.w);
        uint32_t pixel = path % a.npx + a.pixel_base, frame = a.frame_base + path / a.npx;
        TraceResult r;
        r.t = h.x; r.u = h.y; r.v = h.z; r.prim = __float_as_uint(h.w); r.slot = a.hit_slot[i]; r.hit = h.x >= 0.0f;
        if (GENERAL && a.probe_mode && bounce == 1) a.color[path].w = r.t;          // hit distance of the probe ray (or -1)
        KernelSink sink{a, bounce, path};
        cont = shade_vertex<KernelSink, IBL, RECT, GENERAL>(a.sc, a.sp, frame, bounce, pixel, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), v3(w.x, w.y, w.z), r, sink, nO, nD, nW);
    }
    // next-ray queue: ballot per warp, ONE atomicAdd per block (all threads of the block reach this point)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#if BPT_SORT_OCTANT
    // The block's surviving rays are written grouped by direction octant (a counting sort over 8 keys inside the block): the warps of
    // the next extend launch then hold rays that agree on the near / far order of every node, which is what keeps their lanes together.
    // Queue CONTENT is unchanged (a set), only the order inside the block's slot range differs.
    __shared__ uint32_t s_cnt[kBlock / 32][8], s_off[kBlock / 32][8];
    if (threadIdx.x < (kBlock / 32) * 8) (&s_cnt[0][0])[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t key = cont ? ((nD.x < 0.0f ? 1u : 0u) | (nD.y < 0.0f ? 2u : 0u) | (nD.z < 0.0f ? 4u : 0u)) : 8u;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (cont && rank == 0) s_cnt[warp][key] = __popc(peers);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int k = 0; k < 8; k++)
            for (int w = 0; w < kBlock / 32; w++) { s_off[w][k] = total; total += s_cnt[w][k]; }
        uint32_t base = total ? atomicAdd(&a.qcount[QE + bounce + 1], total) : 0u;
        for (int k = 0; k < 8; k++)
            for (int w = 0; w < kBlock / 32; w++) s_off[w][k] += base;
    }
    __syncthreads();
    uint32_t slot = cont ? s_off[warp][key] + rank : 0u;
#else
    __shared__ uint32_t s_warp_cnt[kBlock / 32], s_warp_base[kBlock / 32];
    const uint32_t ballot = __ballot_sync(0xffffffffu, cont);
    if (lane == 0) s_warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) { s_warp_base[w] = total; total += s_warp_cnt[w]; }
        uint32_t base = total ? atomicAdd(&a.qcount[QE + bounce + 1], total) : 0u;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) s_warp_base[w] += base;
    }
    __syncthreads();
    uint32_t slot = s_warp_base[warp] + __popc(ballot & ((1u << lane) - 1u));
#endif
    if (cont) {
        a.ray_o_out[slot] = make_float4(nO.x, nO.y, nO.z, __uint_as_float(path));
        a.ray_d_out[slot] = make_float4(nD.x, nD.y, nD.z, 0.0f);
        a.ray_w_out[slot] = make_float4(nW.x, nW.y, nW.z, 0.0f);
    }
}

static const auto k_shade_ibl = k_shade<true, true>;
static const auto k_shade_rect = k_shade<false, true>;
static const auto k_shade_norect = k_shade<false, false>;
static const auto k_shade_plain = k_shade<false, false, false>;

// ---- reference_fp16 only: end of a bounce. colour = half(bounce sum * throughput), written (bounce 1) or added with a
//      half result (deferred_lighting_secondary.hlsl:110 + the additive blit, path_tracing.cpp:441-459) --------------
__global__ void __launch_bounds__(kBlock) k_commit_bounce(const __grid_constant__ RenderArgs a, uint32_t bounce) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.qcount[QE + bounce]) return;
    float4 w = a.ray_w_in[i];
    uint32_t path = __float_as_uint(a.ray_o_in[i].w);
    float4 b = a.contrib[path], c = a.color[path];
    float3 r = commit_bounce_fp16(v3(c.x, c.y, c.z), v3(b.x, b.y, b.z), v3(w.x, w.y, w.z), bounce);
    a.color[path] = make_float4(r.x, r.y, r.z, c.w);
    a.contrib[path] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ---- per-sample bookkeeping: queue lengths → 64-bit totals -------------------------------------
__global__ void k_tally(const uint32_t* __restrict__ qcount, uint64_t* __restrict__ totals, uint32_t npx) {
    uint32_t t = threadIdx.x;
    if (t < 16) totals[t] += qcount[QE + t];
    else if (t < 32) totals[t] += qcount[QS + (t - 16)];
    else if (t == 32) totals[32] += npx;
}

// ---- accumulate (pt_accumulate.hlsl:3-11 as FP32 sum): sum[p] += C_s[p], slots in ascending order ---
__global__ void k_accumulate(const float4* __restrict__ color, float4* __restrict__ accum, uint32_t npx, uint32_t slot_begin, uint32_t slot_end) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    float4 v = accum[p];
    for (uint32_t s = slot_begin; s < slot_end; s++) {
        float4 c = color[(size_t)s * npx + p];
        v.x += c.x; v.y += c.y; v.z += c.z;
    }
    accum[p] = v;
}

// One engine frame of a frame-at-a-time pass: fold sample slot `slot` into the sum AND write OutputData.color (rgba16_sfloat, as
// k_resolve_rgba16f) in the same pass over the pixels — 116 instead of 150 bytes per pixel and one launch instead of two per frame.
__global__ void k_accumulate_resolve_rgba16f(const float4* __restrict__ color, float4* __restrict__ accum, uint2* __restrict__ out, uint32_t npx, uint32_t slot, float inv) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    float4 v = accum[p];
    float4 c = color[(size_t)slot * npx + p];
    v.x += c.x; v.y += c.y; v.z += c.z;
    accum[p] = v;
    const uint32_t r = __half_as_ushort(__float2half_rn(v.x * inv)), g = __half_as_ushort(__float2half_rn(v.y * inv)),
                   b = __half_as_ushort(__float2half_rn(v.z * inv));
    out[p] = make_uint2(r | (g << 16), b | (0x3c00u << 16));
}

// reference_fp16: the running fp16 lerp of pt_accumulate.hlsl, samples in ascending frame order; `count` = samples already in the image
__global__ void k_accumulate_fp16(const float4* __restrict__ color, float4* __restrict__ accum, uint32_t npx, uint32_t slot_begin, uint32_t slot_end, uint32_t count) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    float4 v = accum[p];
    float3 img = v3(v.x, v.y, v.z);
    for (uint32_t s = slot_begin; s < slot_end; s++) {
        float4 c = color[(size_t)s * npx + p];
        img = accumulate_fp16(img, v3(c.x, c.y, c.z), ++count);
    }
    accum[p] = make_float4(img.x, img.y, img.z, v.w);
}

__global__ void k_resolve(const float4* __restrict__ accum, float4* __restrict__ out, uint32_t npx, float inv) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    float4 v = accum[p];
    out[p] = make_float4(v.x * inv, v.y * inv, v.z * inv, 1.0f);
}

// OutputData.color in the reference's own texture format (rgba16_sfloat, path_tracing.cpp:248-252): four IEEE halves per pixel,
// round-to-nearest-even of the FP32 mean (q_half's rounding), alpha 1.
__global__ void k_resolve_rgba16f(const float4* __restrict__ accum, uint2* __restrict__ out, uint32_t npx, float inv) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    float4 v = accum[p];
    const uint32_t r = __half_as_ushort(__float2half_rn(v.x * inv)), g = __half_as_ushort(__float2half_rn(v.y * inv)),
                   b = __half_as_ushort(__float2half_rn(v.z * inv));
    out[p] = make_uint2(r | (g << 16), b | (0x3c00u << 16));
}

// ---- arbitrary ray batches (bpt_trace_rays / bpt_trace_shadow_rays) ----------------------------
__global__ void __launch_bounds__(kBlock) k_trace_batch(const __grid_constant__ DScene sc, const bpt_ray* __restrict__ rays, uint64_t n, uint32_t frame_index,
                                                        bpt_hit* __restrict__ hits, uint8_t* __restrict__ visible, const DInstance* __restrict__ inst,
                                                        const float4* __restrict__ wide, const float4* __restrict__ leafbox, uint32_t wide2) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bpt_ray r = rays[i];
    float3 O = v3(r.origin[0], r.origin[1], r.origin[2]), D = v3(r.direction[0], r.direction[1], r.direction[2]);
    if (hits) {
        TraceResult t = wide2 ? trace_ray_wide_two_level<false>(sc, O, D, r.tmin, r.tmax, frame_index) : wide ? trace_ray_wide<false>(sc, wide, leafbox, O, D, r.tmin, r.tmax, frame_index) : trace_ray<false>(sc, O, D, r.tmin, r.tmax, frame_index);
        bpt_hit h;
        h.t = t.t; h.u = t.u; h.v = t.v;
        h.instance = t.hit ? inst[t.slot].instance_id : 0xffffffffu;
        h.primitive = t.hit ? t.prim : 0xffffffffu;
        hits[i] = h;
    } else {
        TraceResult t = wide2 ? trace_ray_wide_two_level<true>(sc, O, D, r.tmin, r.tmax, frame_index) : wide ? trace_ray_wide<true>(sc, wide, leafbox, O, D, r.tmin, r.tmax, frame_index) : trace_ray<true>(sc, O, D, r.tmin, r.tmax, frame_index);
        visible[i] = t.hit ? 0 : 1;
    }
}

// ---- primary-hit outputs (OutputData.depth / .gbuffer of PathTracingPass::render): one thread per camera ray ----
__global__ void __launch_bounds__(kBlock) k_primary_aov(const __grid_constant__ RenderArgs a, float* __restrict__ depth, bpt_gbuffer_texel* __restrict__ gbuffer) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.qcount[QE + 1]) return;
    float4 o = a.ray_o_in[i], d = a.ray_d_in[i], h = a.hit[i];
    uint32_t pixel = __float_as_uint(o.w);                             // one slot: path id = pixel
    TraceResult r;
    r.t = h.x; r.u = h.y; r.v = h.z; r.prim = __float_as_uint(h.w); r.slot = a.hit_slot[i]; r.hit = h.x >= 0.0f;
    float z; bpt_gbuffer_texel g;
    primary_outputs(a.sc, a.cam, v3(o.x, o.y, o.z), v3(d.x, d.y, d.z), r, z, g);
    depth[pixel] = z;
    gbuffer[pixel] = g;
}

// ---- ray-traced ambient occlusion (ambient_occlusion_rt.hlsl:14-66): 4 any-hit rays per pixel through the connect kernel ----
__global__ void __launch_bounds__(kBlock) k_ao_raygen(const __grid_constant__ RenderArgs a, uint32_t aw, uint32_t ah, uint32_t half_res, float range,
                                                      const float* __restrict__ depth, const float4* __restrict__ normal_roughness) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    float3 origin = v3s(0.0f), dirs[4];
    if (p < aw * ah) {
        valid = ao_pixel_rays(a.cam, p % aw, p / aw, aw, ah, a.sp.width, a.sp.height, a.frame_base, half_res, depth, normal_roughness, origin, dirs);
        a.color[p] = make_float4(0.0f, 0.0f, 0.0f, valid ? 1.0f : 0.0f);       // x counts the UNoccluded rays
    }
    for (int i = 0; i < 4; i++) {                                               // (all threads of the warp reach queue_push)
        uint32_t slot = queue_push(&a.qcount[QS + 1], valid);
        if (valid && slot < a.shadow_capacity) {
            a.sh_o[slot] = make_float4(origin.x, origin.y, origin.z, __uint_as_float(p));
            a.sh_d[slot] = make_float4(dirs[i].x, dirs[i].y, dirs[i].z, range);
            a.sh_c[slot] = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
        }
    }
}
// ---- ray-traced reflections (reflection.cpp:317-450): specular_sample -> extend -> shade + connect of ONE bounce ----
__global__ void __launch_bounds__(kBlock) k_rtr_raygen(const __grid_constant__ RenderArgs a, uint32_t rw, uint32_t rh, uint32_t half_res, float max_roughness,
                                                       float fade_roughness, float strength, const float* __restrict__ depth, const bpt_gbuffer_texel* __restrict__ gbuffer) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    float3 O = v3s(0.0f), D = v3s(0.0f), Wt = v3s(0.0f);
    if (p < rw * rh) {
        valid = rtr_pixel_ray(a.cam, p % rw, p / rw, rw, rh, a.sp.width, a.sp.height, a.frame_base, half_res, depth, gbuffer, max_roughness, fade_roughness, O, D, Wt);
        a.color[p] = make_float4(0.0f, 0.0f, 0.0f, -1.0f);             // w: hit distance of the reflection ray (probe_mode), -1 = miss / no ray
    }
    uint32_t slot = queue_push(&a.qcount[QE + 1], valid);             // (all threads of the warp reach queue_push)
    if (valid) {
        Wt = Wt * strength;                                             // deferred_lighting_secondary.hlsl:17
        a.ray_o_out[slot] = make_float4(O.x, O.y, O.z, __uint_as_float(p));
        a.ray_d_out[slot] = make_float4(D.x, D.y, D.z, 0.0f);
        a.ray_w_out[slot] = make_float4(Wt.x, Wt.y, Wt.z, 0.0f);
    }
}
// colour -> (rgb, 1); hit positions: (P, t) on a hit (rt_gbuffer.hlsl:32), (direction, -1) on a miss (:34), (0, 0, 0, -1) without a ray
__global__ void __launch_bounds__(kBlock) k_rtr_finish(const __grid_constant__ RenderArgs a, uint32_t rw, uint32_t rh, uint32_t half_res, float max_roughness,
                                                       float fade_roughness, const float* __restrict__ depth, const bpt_gbuffer_texel* __restrict__ gbuffer,
                                                       float4* __restrict__ out_refl, float4* __restrict__ out_hit) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= rw * rh) return;
    float4 c = a.color[p];
    out_refl[p] = make_float4(c.x, c.y, c.z, 1.0f);
    float3 O, D, Wt;
    float4 hp = make_float4(0.0f, 0.0f, 0.0f, -1.0f);
    if (rtr_pixel_ray(a.cam, p % rw, p / rw, rw, rh, a.sp.width, a.sp.height, a.frame_base, half_res, depth, gbuffer, max_roughness, fade_roughness, O, D, Wt)) {
        if (c.w >= 0.0f) { float3 P = O + D * c.w; hp = make_float4(P.x, P.y, P.z, c.w); }
        else hp = make_float4(D.x, D.y, D.z, -1.0f);
    }
    out_hit[p] = hp;
}

__global__ void k_upscale_half_res(const __grid_constant__ bpt_camera cam, uint32_t W, uint32_t H, uint32_t frame_index, const float* __restrict__ depth,
                                   const float4* __restrict__ normal_roughness, const float4* __restrict__ in_half, float4* __restrict__ out) {
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    out[(size_t)y * W + x] = upscale_pixel(cam, x, y, W, H, frame_index, depth, normal_roughness, in_half);
}

__global__ void k_ao_finish(const float4* __restrict__ color, uint32_t n, float strength, float2* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 c = color[p];
    // ao_tex is rg16_sfloat (ambient_occlusion.cpp:14): the value a later pass loads is the half
    out[p] = c.w != 0.0f ? make_float2(q_half(ao_value(4u - (uint32_t)c.x, strength)), 1.0f) : make_float2(1.0f, 0.0f);
}

// ---- DDGI consumer for arbitrary points (calc_ddgi_volume_lighting as the deferred lighting passes call it) ----
__global__ void k_ddgi_lighting(const __grid_constant__ DScene sc, uint64_t n, const float* __restrict__ pos, const float* __restrict__ normal,
                                const float* __restrict__ view, float4* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = ddgi_volume_lighting(sc.ddgi_volume, sc.ddgi_irr_size, sc.ddgi_vis_size, sc.ddgi_irradiance, sc.ddgi_visibility,
                                  v3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), v3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]),
                                  v3(view[3 * i], view[3 * i + 1], view[3 * i + 2]));
}

// ---- DDGI probe blending: one block per probe, one thread per octahedral texel --------------------
template <bool VIS>
__global__ void k_probe_blend(const __grid_constant__ bpt_probe_volume vol, const float2* __restrict__ table, uint32_t frame_index,
                              const float4* __restrict__ rays, uint32_t size, float alpha, uint32_t history_valid, float* __restrict__ atlas) {
    extern __shared__ float4 s_mem[];
    const uint32_t nrays = vol.rays_per_probe, probe = blockIdx.x;
    float4* s_rad = s_mem;
    float3* s_dir = reinterpret_cast<float3*>(s_mem + nrays);
    for (uint32_t r = threadIdx.x; r < nrays; r += blockDim.x) {
        float3 O, D;
        probe_ray(vol, table, probe * nrays + r, frame_index, O, D);
        float4 v = rays[(size_t)probe * nrays + r];
        s_rad[r] = v;
        s_dir[r] = blend_trace_dir(O, D, v.w);
    }
    __syncthreads();
    const bool active = threadIdx.x < size * size;       // (block size is rounded up to a warp multiple)
    const uint32_t tx = threadIdx.x % size, ty = active ? threadIdx.x / size : 0;
    float3 val = active ? blend_texel<VIS>(tx, ty, size, s_dir, s_rad, nrays) : v3s(0.0f);
    const uint32_t nx = vol.probe_counts[0], ny = vol.probe_counts[1];
    const uint32_t ix = probe % nx, iy = (probe / nx) % ny, iz = probe / nx / ny;
    const uint32_t stride = nx * ny * (size + 2), ch = VIS ? 2 : 4;
    const uint32_t sx = (iy * nx + ix) * (size + 2), sy = iz * (size + 2);
    const uint32_t cx = tx + 1, cy = ty + 1;
    auto at = [&](uint32_t x, uint32_t y) { return atlas + ((size_t)(sy + y) * stride + (sx + x)) * ch; };
    if (history_valid && active) {
        const float* h = at(cx, cy);
        val.x = temporal_blend(val.x, h[0], alpha); val.y = temporal_blend(val.y, h[1], alpha);
        if (!VIS) val.z = temporal_blend(val.z, h[2], alpha);
    }
    auto put = [&](uint32_t x, uint32_t y) { float* o = at(x, y); o[0] = val.x; o[1] = val.y; if (!VIS) { o[2] = val.z; o[3] = 1.0f; } };
    __syncthreads();     // every history texel of this probe has been read before borders are overwritten
    if (!active) return;
    put(cx, cy);
    uint32_t bx, by;
    border_coord(cx, cy, size, bx, by);
    put(bx, by);
    uint32_t c[4];
    if (corner_coords(cx, cy, size, c)) { put(c[0], c[1]); put(c[2], c[3]); }
}

} // namespace

// NVTX ranges carry the labels the reference gives its render-graph passes (path_tracing.cpp:293,313,347,385,429,442,465): a timeline
// of this library reads like a RenderDoc / Nsight capture of the engine. The fused kernels map as: k_raygen = "PT Generate Camera Ray",
// extend = "PT Trace GBuffer #i", k_shade = "PT Lighting #i" + "PT Sample Ray #i+1" (one kernel), connect = the shadow-map lookups of
// "PT Lighting #i" as rays, k_accumulate = "PT Blit Color" + "PT Accumulate".
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    NvtxRange(const char* fmt, unsigned i) { char b[64]; snprintf(b, sizeof(b), fmt, i); nvtxRangePushA(b); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

#define LAUNCH(ctx, kernel, grid, block, ...)                                   \
    do {                                                                        \
        kernel<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);             \
        (ctx)->launches++;                                                      \
        {                                                                       \
            cudaError_t le__ = cudaGetLastError();                              \
            if (le__ != cudaSuccess) {                                          \
                (ctx)->err = std::string("launch of " #kernel " (grid ") + std::to_string((unsigned long long)(grid)) + "): " + cudaGetErrorString(le__); \
                return BPT_ERR_CUDA;                                            \
            }                                                                   \
        }                                                                       \
    } while (0)

// LAUNCH with optional event bracketing (bpt_profile_enable): class 0 raygen, 1 extend, 2 shade, 3 connect, 4 other
#define LAUNCH_T(ctx, cls, kernel, grid, block, ...)                                              \
    do {                                                                                          \
        bpt_context::ProfEvent pe__{nullptr, nullptr, (cls)};                                     \
        if ((ctx)->profile) {                                                                     \
            BPT_CUDA_TRY(ctx, cudaEventCreate(&pe__.a)); BPT_CUDA_TRY(ctx, cudaEventCreate(&pe__.b)); \
            BPT_CUDA_TRY(ctx, cudaEventRecord(pe__.a, (ctx)->stream));                            \
        }                                                                                         \
        LAUNCH(ctx, kernel, grid, block, __VA_ARGS__);                                            \
        if ((ctx)->profile) { BPT_CUDA_TRY(ctx, cudaEventRecord(pe__.b, (ctx)->stream)); (ctx)->prof_events.push_back(pe__); } \
    } while (0)

// Samples per wave: several samples of every pixel are in flight together so that the late, nearly
// empty bounces (each kernel has a ~100 us latency floor: one warp's longest traversal) are amortised
// over more paths. Bounded by a path budget and by the shadow-queue footprint.
static uint32_t wave_slots(const bpt_context* ctx) {
    const uint64_t npx = (uint64_t)ctx->width * ctx->height;
    const uint64_t nl = std::max<uint64_t>((uint64_t)ctx->num_dir + ctx->num_point + ctx->num_rect, 1);
    // <= 67 M paths in flight (32 samples at 1080p; measured on configs[1]: 2^24 -> 1.95, 2^25 -> 1.87, 2^26 -> 1.82, 2^27 -> 1.80 ms
    // per sample: every late, nearly empty bounce has a ~150 us latency floor that more samples per wave amortise);
    // BPT_WAVE_PATHS_LOG2 overrides for tuning
    static const int log2_paths = [] { const char* e = getenv("BPT_WAVE_PATHS_LOG2"); int v = e ? atoi(e) : 26; return v < 16 ? 16 : (v > 30 ? 30 : v); }();
    const uint64_t budget = ctx->wave_paths_budget ? ctx->wave_paths_budget : (1ull << log2_paths);     // bpt_set_wave_budget: the host's cap on the footprint
    uint64_t by_paths = std::max<uint64_t>(1, budget / npx);
    uint64_t by_shadow = std::max<uint64_t>(1, (8ull << 30) / (npx * nl * 48));    // <= 8 GiB of shadow-ray records
    return (uint32_t)std::min<uint64_t>(std::min(by_paths, by_shadow), 64);
}

bpt_status wavefront_alloc(bpt_context* ctx) {
    WavefrontState& wf = ctx->wf;
    const uint32_t npx = ctx->width * ctx->height;
    const uint32_t slots = wave_slots(ctx);
    const uint64_t paths = (uint64_t)npx * slots;
    bpt_status s;
    if (wf.npx != npx) {
        dev_free(wf.accum);
        if ((s = dev_alloc(ctx, wf.accum, (size_t)npx * 16))) return s;
        BPT_CUDA_TRY(ctx, cudaMemsetAsync(wf.accum.p, 0, (size_t)npx * 16, ctx->stream));
        wf.npx = npx;
        wf.capacity = 0;
        wf.accum_count = 0; wf.accum_fp16 = false; ctx->accum_used = false;
    }
    if (wf.capacity != paths) {
        for (int k = 0; k < 2; k++) {
            dev_free(wf.ray_o[k]); dev_free(wf.ray_d[k]); dev_free(wf.ray_w[k]);
            if ((s = dev_alloc(ctx, wf.ray_o[k], paths * 16))) return s;
            if ((s = dev_alloc(ctx, wf.ray_d[k], paths * 16))) return s;
            if ((s = dev_alloc(ctx, wf.ray_w[k], paths * 16))) return s;
        }
        dev_free(wf.hit); dev_free(wf.hit_slot); dev_free(wf.color); dev_free(wf.bcol);
        if ((s = dev_alloc(ctx, wf.hit, paths * 16))) return s;
        if ((s = dev_alloc(ctx, wf.hit_slot, paths * 4))) return s;
        if ((s = dev_alloc(ctx, wf.color, paths * 16))) return s;
        wf.capacity = paths;
        wf.slots = slots;
        wf.shadow_capacity = 0;
        wf.ahead_slots = wf.ahead_cursor = 0;        // the per-sample colours of a prefetched wave went with the old buffer
    }
    if (!wf.qcount.p) {
        if ((s = dev_alloc(ctx, wf.qcount, QN * sizeof(uint32_t)))) return s;
        if ((s = dev_alloc(ctx, wf.totals, 40 * sizeof(uint64_t)))) return s;
        BPT_CUDA_TRY(ctx, cudaMemsetAsync(wf.totals.p, 0, 40 * sizeof(uint64_t), ctx->stream));
    }
    const uint64_t nl = std::max<uint64_t>((uint64_t)ctx->num_dir + ctx->num_point + ctx->num_rect, 1);   // rect lights may emit rays (rect_shadow)
    const uint64_t need = paths * nl;
    if (wf.shadow_capacity < need) {
        if (need * 48 > (64ull << 30)) { ctx->err = "shadow-ray queue would exceed 64 GiB; reduce lights or resolution"; return BPT_ERR_OOM; }
        dev_free(wf.sh_o); dev_free(wf.sh_d); dev_free(wf.sh_c);
        if ((s = dev_alloc(ctx, wf.sh_o, need * 16))) return s;
        if ((s = dev_alloc(ctx, wf.sh_d, need * 16))) return s;
        if ((s = dev_alloc(ctx, wf.sh_c, need * 16))) return s;
        wf.shadow_capacity = need;
    }
    return BPT_OK;
}

static bpt_status capture_bounce(bpt_context* ctx, uint32_t bounce, int in_buf) {
    WavefrontState& wf = ctx->wf;
    uint32_t qc[QN];
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(qc, wf.qcount.p, sizeof(qc), cudaMemcpyDeviceToHost, ctx->stream));
    BPT_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t ne = qc[QE + bounce], ns = qc[QS + bounce];
    std::vector<float4> o(ne), h(ne); std::vector<uint32_t> slot(ne);
    std::vector<DInstance> inst(ctx->h_instances.size());
    BPT_CUDA_TRY(ctx, cudaMemcpy(o.data(), wf.ray_o[in_buf].p, (size_t)ne * 16, cudaMemcpyDeviceToHost));
    BPT_CUDA_TRY(ctx, cudaMemcpy(h.data(), wf.hit.p, (size_t)ne * 16, cudaMemcpyDeviceToHost));
    BPT_CUDA_TRY(ctx, cudaMemcpy(slot.data(), wf.hit_slot.p, (size_t)ne * 4, cudaMemcpyDeviceToHost));
    BPT_CUDA_TRY(ctx, cudaMemcpy(inst.data(), ctx->d_instances.p, inst.size() * sizeof(DInstance), cudaMemcpyDeviceToHost));
    auto& ep = ctx->cap_extend_pixels[bounce]; auto& eh = ctx->cap_extend_hits[bounce];
    ep.resize(ne); eh.resize(ne);
    for (uint32_t i = 0; i < ne; i++) {
        uint32_t px, prim;
        memcpy(&px, &o[i].w, 4); memcpy(&prim, &h[i].w, 4);
        ep[i] = px;
        bool hit = h[i].x >= 0.0f;
        eh[i] = bpt_hit{h[i].x, h[i].y, h[i].z, hit ? inst[slot[i]].instance_id : 0xffffffffu, hit ? prim : 0xffffffffu};
    }
    std::vector<float4> so(ns), scv(ns);
    BPT_CUDA_TRY(ctx, cudaMemcpy(so.data(), wf.sh_o.p, (size_t)ns * 16, cudaMemcpyDeviceToHost));
    BPT_CUDA_TRY(ctx, cudaMemcpy(scv.data(), wf.sh_c.p, (size_t)ns * 16, cudaMemcpyDeviceToHost));
    auto& spx = ctx->cap_shadow_pixels[bounce]; auto& sl = ctx->cap_shadow_lights[bounce];
    spx.resize(ns); sl.resize(ns);
    for (uint32_t i = 0; i < ns; i++) { memcpy(&spx[i], &so[i].w, 4); memcpy(&sl[i], &scv[i].w, 4); }
    return BPT_OK;
}

// Fills the parts of RenderArgs that do not depend on the wave.
static bpt_status prepare_args(bpt_context* ctx, RenderArgs& a, const bpt_settings& st, uint32_t B) {
    WavefrontState& wf = ctx->wf;
    a.sc = ctx->scene_view();
    a.sp.width = ctx->width; a.sp.height = ctx->height; a.sp.max_bounces = B; a.sp.nee_mode = st.nee_mode; a.sp.ray_length = st.ray_length;
    a.sp.diffuse_only = 0; a.sp.russian_roulette = st.russian_roulette; a.sp.rect_shadow = st.rect_shadow; a.pixel_jitter = st.pixel_jitter;
    a.sp.state_precision = st.state_precision; a.sp.ibl = 0;
    if (st.state_precision == BPT_STATE_REFERENCE_FP16 && !wf.bcol.p) {          // per-bounce light sums, only this mode needs them
        bpt_status sb = dev_alloc(ctx, wf.bcol, wf.capacity * 16);
        if (sb) return sb;
    }
    a.hit = wf.hit.as<float4>(); a.hit_slot = wf.hit_slot.as<uint32_t>();
    a.sh_o = wf.sh_o.as<float4>(); a.sh_d = wf.sh_d.as<float4>(); a.sh_c = wf.sh_c.as<float4>();
    a.accum = wf.accum.as<float4>(); a.color = wf.color.as<float4>();
    a.contrib = st.state_precision == BPT_STATE_REFERENCE_FP16 ? wf.bcol.as<float4>() : a.color;
    a.qcount = wf.qcount.as<uint32_t>(); a.shadow_capacity = wf.shadow_capacity;
    a.pixel_base = 0; a.probe_mode = 0; a.cull_non_opaque = 0; a.camera_paths = 0;
    if (ctx->accel_mode == BPT_ACCEL_MERGED) {
        a.m_nodes = ctx->blas[0].nodes.as<float4>(); a.m_tris = ctx->blas[0].tris.as<float4>(); a.m_root = ctx->blas[0].root; a.m_n = ctx->blas[0].n;
        a.m_wide = ctx->blas[0].wide.as<float4>(); a.m_leafbox = ctx->blas[0].leafbox.as<float4>();
    } else { a.m_nodes = nullptr; a.m_tris = nullptr; a.m_root = 0; a.m_n = 0; a.m_wide = nullptr; a.m_leafbox = nullptr; }
    return BPT_OK;
}

// Which traversal kernel serves the current acceleration structure: merged mode uses the 4-wide quantised tree when it was
// built (always, unless BPT_WIDE=0 asks for the binary tree — kept for A/B measurements), two-level mode the binary trees.
// (both from bounce 2: camera rays and the shadow rays of camera-ray hits are coherent — a warp fetches 32 neighbouring pixels'
// rays towards the same light — and issue-bound, so the binary tree is faster for them. Measured: extend wide from bounce 1 / 2 /
// 3 -> 1.838 / 1.725 / 1.774 ms per configs[1] frame; connect wide from 1 vs 2: equal on configs[1], 22.65 vs 22.12 ms on configs[2].)
static bool use_wide(const bpt_context* ctx, uint32_t bounce = 99, bool connect = false) {
    static const bool enabled = [] { const char* e = getenv("BPT_WIDE"); return !e || atoi(e) != 0; }();
    static const uint32_t from_bounce = [] { const char* e = getenv("BPT_WIDE_FROM_BOUNCE"); return e ? (uint32_t)atoi(e) : 2u; }();
    static const uint32_t connect_from = [] { const char* e = getenv("BPT_WIDE_CONNECT_FROM_BOUNCE"); return e ? (uint32_t)atoi(e) : 2u; }();
    return enabled && bounce >= (connect ? connect_from : from_bounce) && ctx->accel_mode == BPT_ACCEL_MERGED && ctx->blas[0].wide.p != nullptr;
}
// Two-level mode: the wide TLAS + wide BLASes (BPT_WIDE2=0: binary trees; BPT_WIDE2_FROM_BOUNCE / BPT_WIDE2_CONNECT_FROM_BOUNCE: thresholds).
static bool use_wide2(const bpt_context* ctx, uint32_t bounce = 99, bool connect = false) {
    static const bool enabled = [] { const char* e = getenv("BPT_WIDE2"); return !e || atoi(e) != 0; }();
    static const uint32_t from_bounce = [] { const char* e = getenv("BPT_WIDE2_FROM_BOUNCE"); return e ? (uint32_t)atoi(e) : 1u; }();
    static const uint32_t connect_from = [] { const char* e = getenv("BPT_WIDE2_CONNECT_FROM_BOUNCE"); return e ? (uint32_t)atoi(e) : 1u; }();
    if (!enabled || bounce < (connect ? connect_from : from_bounce) || ctx->accel_mode != BPT_ACCEL_TWO_LEVEL) return false;
    return ctx->tlas.n <= 1 || ctx->tlas.wide.p != nullptr;      // (every BLAS with two or more triangles has its wide form, bvh_build.cu)
}
// Resident grid of a persistent traversal kernel: SMs x the blocks of that very instantiation that fit on one (queried once per kernel).
template <class K>
static unsigned resident_grid(bpt_context* ctx, K kernel) {
    const void* key = reinterpret_cast<const void*>(kernel);
    for (auto& e : ctx->wf.grids) if (e.first == key) return e.second;
    int dev = 0, sms = 148, blocks = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, kBlock, 0) != cudaSuccess) { blocks = 8; (void)cudaGetLastError(); }
    const unsigned g = (unsigned)(sms * std::max(blocks, 1));
    ctx->wf.grids.emplace_back(key, g);
    return g;
}
// BPT_SPECIALISE=0 launches the general kernels everywhere (A/B measurements)
static bool specialise_opaque() { static const bool on = [] { const char* e = getenv("BPT_SPECIALISE"); return !e || atoi(e) != 0; }(); return on; }
// Camera rays as packets (k_trace_packet in merged mode, k_extend_packet_2l in two-level mode) pay when a pixel tile's rays really share
// their nodes, i.e. when triangles are not much smaller than pixels: configs[1] (262 k triangles at 1080p: 0.13 per pixel) +6 %, the same
// atrium in two-level mode 1 816 -> 2 011 Mrays/s, but configs[3] (2 M x 512 instanced triangles at 4K: 129 per pixel) 924 -> 782
// (profiles/r2an_variants.jsonl). Default: packets iff the instanced triangle count is at most twice the pixel count.
// BPT_PACKET overrides (A/B runs): bit 0 = merged-mode camera rays, bit 1 = the shadow rays of their hit points (measured: slower),
// bit 4 = two-level camera rays; 0 = everything through k_trace_spec.
static uint32_t packet_mode(const bpt_context* ctx) {
    static const int forced = [] { const char* e = getenv("BPT_PACKET"); return e ? atoi(e) : -1; }();
    if (forced >= 0) return (uint32_t)forced;
    return ctx->instanced_triangles <= 2ull * ctx->width * ctx->height ? 17u : 0u;
}
static bpt_status launch_extend(bpt_context* ctx, const RenderArgs& a, uint32_t i) {
    const bool ah = ctx->scene_has_anyhit || !specialise_opaque();
    if (i == 1 && a.camera_paths && (packet_mode(ctx) & 16u) && ctx->accel_mode == BPT_ACCEL_TWO_LEVEL) {
        if (ah) LAUNCH_T(ctx, 1, k_extend_packet_2l_a, resident_grid(ctx, k_extend_packet_2l_a), kBlock, a, i); else LAUNCH_T(ctx, 1, k_extend_packet_2l_o, resident_grid(ctx, k_extend_packet_2l_o), kBlock, a, i);
        return BPT_OK;
    }
    if (i == 1 && a.camera_paths && (packet_mode(ctx) & 1u) && ctx->accel_mode == BPT_ACCEL_MERGED) {
        if (ah) LAUNCH_T(ctx, 1, k_extend_packet, resident_grid(ctx, k_extend_packet), kBlock, a, i); else LAUNCH_T(ctx, 1, k_extend_packet_o, resident_grid(ctx, k_extend_packet_o), kBlock, a, i);
        return BPT_OK;
    }
    if (use_wide2(ctx, i)) { if (ah) LAUNCH_T(ctx, 1, k_extend_two_level_wide, resident_grid(ctx, k_extend_two_level_wide), kBlock, a, i); else LAUNCH_T(ctx, 1, k_extend_two_level_wide_o, resident_grid(ctx, k_extend_two_level_wide_o), kBlock, a, i); }
    else if (use_wide(ctx, i)) { if (ah) LAUNCH_T(ctx, 1, k_extend_wide, resident_grid(ctx, k_extend_wide), kBlock, a, i); else LAUNCH_T(ctx, 1, k_extend_wide_o, resident_grid(ctx, k_extend_wide_o), kBlock, a, i); }
    else if (ctx->accel_mode == BPT_ACCEL_MERGED) { if (ah) LAUNCH_T(ctx, 1, k_extend_merged, resident_grid(ctx, k_extend_merged), kBlock, a, i); else LAUNCH_T(ctx, 1, k_extend_merged_o, resident_grid(ctx, k_extend_merged_o), kBlock, a, i); }
    else { if (ah) LAUNCH_T(ctx, 1, k_extend_two_level, resident_grid(ctx, k_extend_two_level), kBlock, a, i); else LAUNCH_T(ctx, 1, k_extend_two_level_o, resident_grid(ctx, k_extend_two_level_o), kBlock, a, i); }
    return BPT_OK;
}
static bpt_status launch_connect(bpt_context* ctx, const RenderArgs& a, uint32_t i) {
    const bool ah = ctx->scene_has_anyhit || !specialise_opaque();
    if (i == 1 && a.camera_paths && (packet_mode(ctx) & 2u) && ctx->accel_mode == BPT_ACCEL_MERGED) {
        if (ah) LAUNCH_T(ctx, 3, k_connect_packet, resident_grid(ctx, k_connect_packet), kBlock, a, i); else LAUNCH_T(ctx, 3, k_connect_packet_o, resident_grid(ctx, k_connect_packet_o), kBlock, a, i);
        return BPT_OK;
    }
    if (use_wide2(ctx, i, true)) { if (ah) LAUNCH_T(ctx, 3, k_connect_two_level_wide, resident_grid(ctx, k_connect_two_level_wide), kBlock, a, i); else LAUNCH_T(ctx, 3, k_connect_two_level_wide_o, resident_grid(ctx, k_connect_two_level_wide_o), kBlock, a, i); }
    else if (use_wide(ctx, i, true)) { if (ah) LAUNCH_T(ctx, 3, k_connect_wide, resident_grid(ctx, k_connect_wide), kBlock, a, i); else LAUNCH_T(ctx, 3, k_connect_wide_o, resident_grid(ctx, k_connect_wide_o), kBlock, a, i); }
    else if (ctx->accel_mode == BPT_ACCEL_MERGED) { if (ah) LAUNCH_T(ctx, 3, k_connect_merged, resident_grid(ctx, k_connect_merged), kBlock, a, i); else LAUNCH_T(ctx, 3, k_connect_merged_o, resident_grid(ctx, k_connect_merged_o), kBlock, a, i); }
    else { if (ah) LAUNCH_T(ctx, 3, k_connect_two_level, resident_grid(ctx, k_connect_two_level), kBlock, a, i); else LAUNCH_T(ctx, 3, k_connect_two_level_o, resident_grid(ctx, k_connect_two_level_o), kBlock, a, i); }
    return BPT_OK;
}

// (extend → shade → connect) for bounces 1..B-1 over the queue that raygen left in ray buffer 0.
static bpt_status run_bounces(bpt_context* ctx, RenderArgs& a, const bpt_settings& st, uint32_t B, uint64_t paths, bool capture) {
    WavefrontState& wf = ctx->wf;
    const unsigned grid_paths = (unsigned)((paths + kBlock - 1) / kBlock);
    bpt_status s;
    int cur = 0;
    for (uint32_t i = 1; i < B; i++) {
        a.ray_o_in = wf.ray_o[cur].as<float4>(); a.ray_d_in = wf.ray_d[cur].as<float4>(); a.ray_w_in = wf.ray_w[cur].as<float4>();
        a.ray_o_out = wf.ray_o[cur ^ 1].as<float4>(); a.ray_d_out = wf.ray_d[cur ^ 1].as<float4>(); a.ray_w_out = wf.ray_w[cur ^ 1].as<float4>();
        {
            NvtxRange r("PT Trace GBuffer #%u", i);
            if ((s = launch_extend(ctx, a, i))) return s;
        }
        {
            NvtxRange r("PT Lighting + Sample Ray #%u", i);
            if (a.sp.ibl) LAUNCH_T(ctx, 2, k_shade_ibl, grid_paths, kBlock, a, i);      // ray-traced reflections with settings.ibl
            else if (ctx->num_rect || !specialise_opaque()) LAUNCH_T(ctx, 2, k_shade_rect, grid_paths, kBlock, a, i);
            else if (a.probe_mode || a.sp.diffuse_only || a.sp.russian_roulette || a.sp.state_precision != BPT_STATE_FP32) LAUNCH_T(ctx, 2, k_shade_norect, grid_paths, kBlock, a, i);
            else LAUNCH_T(ctx, 2, k_shade_plain, grid_paths, kBlock, a, i);
        }
        if (st.nee_mode == BPT_NEE_SHADOW_RAY && (ctx->num_dir + ctx->num_point + (st.rect_shadow ? ctx->num_rect : 0)) > 0) {
            NvtxRange r("PT Lighting #%u (shadow rays)", i);
            if ((s = launch_connect(ctx, a, i))) return s;
        }
        if (st.state_precision == BPT_STATE_REFERENCE_FP16) LAUNCH_T(ctx, 4, k_commit_bounce, grid_paths, kBlock, a, i);
        if (capture && (s = capture_bounce(ctx, i, cur))) return s;
        cur ^= 1;
    }
    return BPT_OK;
}

bpt_status wavefront_render(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_first, uint32_t nsamples, const bpt_settings& st, bool keep_ahead) {
    bpt_status s;
    if ((s = wavefront_alloc(ctx))) return s;
    WavefrontState& wf = ctx->wf;
    wf.ahead_slots = wf.ahead_cursor = 0;         // the colour buffer is about to be overwritten
    if (keep_ahead) nsamples = std::min(nsamples, wf.slots);
    const uint32_t npx = ctx->width * ctx->height;
    const uint32_t B = std::min(std::max(st.max_bounces, 2u), 16u);           // path_tracing.cpp:187,290
    RenderArgs a;
    if ((s = prepare_args(ctx, a, st, B))) return s;
    a.cam = cam;
    a.npx = npx;
    a.camera_paths = 1;
    const bool capture = ctx->capture && nsamples == 1;
    if (capture) {
        ctx->cap_bounces = B;
        ctx->cap_extend_pixels.assign(B, {}); ctx->cap_extend_hits.assign(B, {});
        ctx->cap_shadow_pixels.assign(B, {}); ctx->cap_shadow_lights.assign(B, {});
    }
    for (uint32_t done = 0; done < nsamples;) {
        const uint32_t slots = std::min(nsamples - done, wf.slots);
        const uint64_t paths = (uint64_t)npx * slots;
        a.nslots = slots;
        a.frame_base = frame_first + done;
        BPT_CUDA_TRY(ctx, cudaMemsetAsync(wf.qcount.p, 0, QN * sizeof(uint32_t), ctx->stream));
        a.ray_o_out = wf.ray_o[0].as<float4>(); a.ray_d_out = wf.ray_d[0].as<float4>(); a.ray_w_out = wf.ray_w[0].as<float4>();
        {
            NvtxRange r("PT Generate Camera Ray");
            const uint64_t gen_threads = (uint64_t)((ctx->width + 7) / 8) * ((ctx->height + 3) / 4) * 32 * slots;
            LAUNCH_T(ctx, 0, k_raygen, (unsigned)((gen_threads + kBlock - 1) / kBlock), kBlock, a);
        }
        if ((s = run_bounces(ctx, a, st, B, paths, capture))) return s;
        NvtxRange racc("PT Accumulate");
        if (keep_ahead) { wf.ahead_slots = slots; wf.ahead_cursor = 0; wf.ahead_frame_first = frame_first; }
        else if (st.state_precision == BPT_STATE_REFERENCE_FP16) {
            LAUNCH_T(ctx, 4, k_accumulate_fp16, (npx + 255) / 256, 256, wf.color.as<float4>(), wf.accum.as<float4>(), npx, 0u, slots, wf.accum_count);
            wf.accum_count += slots;
        } else LAUNCH_T(ctx, 4, k_accumulate, (npx + 255) / 256, 256, wf.color.as<float4>(), wf.accum.as<float4>(), npx, 0u, slots);
        LAUNCH_T(ctx, 4, k_tally, 1, 64, wf.qcount.as<uint32_t>(), wf.totals.as<uint64_t>(), (uint32_t)paths);
        done += slots;
    }
    return BPT_OK;
}

// OutputData{depth, gbuffer} of PathTracingPass::render: camera rays of one frame, their closest hits, packed like the trace pass packs them.
bpt_status wavefront_render_primary(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const bpt_settings& st, float* h_depth, bpt_gbuffer_texel* h_gbuffer) {
    NvtxRange r("PT Trace GBuffer #1 + PT Depth");
    bpt_status s;
    if ((s = wavefront_alloc(ctx))) return s;
    WavefrontState& wf = ctx->wf;
    wf.ahead_slots = wf.ahead_cursor = 0;                                     // raygen overwrites the per-sample colours
    const uint32_t npx = ctx->width * ctx->height;
    RenderArgs a;
    if ((s = prepare_args(ctx, a, st, 2))) return s;
    a.cam = cam; a.npx = npx; a.nslots = 1; a.frame_base = frame_index; a.camera_paths = 1;
    DevBuf d_depth, d_g;
    auto cleanup = [&]() { dev_free(d_depth); dev_free(d_g); };
    if ((s = dev_alloc(ctx, d_depth, (size_t)npx * 4)) || (s = dev_alloc(ctx, d_g, (size_t)npx * sizeof(bpt_gbuffer_texel)))) { cleanup(); return s; }
    cudaError_t e = cudaMemsetAsync(wf.qcount.p, 0, QN * sizeof(uint32_t), ctx->stream);
    a.ray_o_out = wf.ray_o[0].as<float4>(); a.ray_d_out = wf.ray_d[0].as<float4>(); a.ray_w_out = wf.ray_w[0].as<float4>();
    a.ray_o_in = a.ray_o_out; a.ray_d_in = a.ray_d_out; a.ray_w_in = a.ray_w_out;
    const uint64_t gen_threads = (uint64_t)((ctx->width + 7) / 8) * ((ctx->height + 3) / 4) * 32;
    if (e == cudaSuccess) {
        k_raygen<<<(unsigned)((gen_threads + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(a);
        if (launch_extend(ctx, a, 1u) != BPT_OK) e = cudaErrorLaunchFailure;
        k_primary_aov<<<(npx + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(a, d_depth.as<float>(), d_g.as<bpt_gbuffer_texel>());
        k_tally<<<1, 64, 0, ctx->stream>>>(wf.qcount.as<uint32_t>(), wf.totals.as<uint64_t>(), 0u);
        ctx->launches += 4;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && h_depth) e = cudaMemcpyAsync(h_depth, d_depth.p, (size_t)npx * 4, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess && h_gbuffer) e = cudaMemcpyAsync(h_gbuffer, d_g.p, (size_t)npx * sizeof(bpt_gbuffer_texel), cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("render_primary: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

// AmbientOcclusionPass::render_raytraced (ambient_occlusion.cpp:217-262): depth + normal G-buffer in, rg16_sfloat (ao, valid) out.
bpt_status wavefront_trace_ao(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const bpt_ao_settings& ao, const float* h_depth,
                              const float* h_normal_roughness, float* h_out) {
    NvtxRange r("RTAO");                                                       // ambient_occlusion.cpp:217-262
    bpt_status s;
    if ((s = wavefront_alloc(ctx))) return s;
    WavefrontState& wf = ctx->wf;
    wf.ahead_slots = wf.ahead_cursor = 0;
    const uint32_t W = ctx->width, H = ctx->height;
    if (ao.half_resolution && ((W | H) & 1u)) { ctx->err = "trace_ao: half resolution needs even width and height (texel-centre reads)"; return BPT_ERR_UNSUPPORTED; }
    const uint32_t aw = ao.half_resolution ? W / 2 : W, ah = ao.half_resolution ? H / 2 : H, n = aw * ah;
    if ((uint64_t)n * 4 > wf.shadow_capacity || n > wf.capacity) { ctx->err = "trace_ao: frame too large for the shadow-ray queue"; return BPT_ERR_UNSUPPORTED; }
    bpt_settings st{};
    st.ray_length = ao.range; st.max_bounces = 2; st.nee_mode = BPT_NEE_SHADOW_RAY;
    RenderArgs a;
    if ((s = prepare_args(ctx, a, st, 2))) return s;
    a.cam = cam; a.npx = n; a.nslots = 1; a.frame_base = frame_index; a.cull_non_opaque = 1;
    DevBuf d_depth, d_nr, d_out;
    auto cleanup = [&]() { dev_free(d_depth); dev_free(d_nr); dev_free(d_out); };
    if ((s = dev_upload(ctx, d_depth, h_depth, (size_t)W * H * 4)) || (s = dev_upload(ctx, d_nr, h_normal_roughness, (size_t)W * H * 16)) ||
        (s = dev_alloc(ctx, d_out, (size_t)n * 8))) { cleanup(); return s; }
    cudaError_t e = cudaMemsetAsync(wf.qcount.p, 0, QN * sizeof(uint32_t), ctx->stream);
    if (e == cudaSuccess) {
        const float range = ao.range > 0.05f ? ao.range : 0.05f;              // ambient_occlusion.cpp:241
        k_ao_raygen<<<(n + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(a, aw, ah, ao.half_resolution ? 1u : 0u, range, d_depth.as<float>(), d_nr.as<float4>());
        if (launch_connect(ctx, a, 1u) != BPT_OK) e = cudaErrorLaunchFailure;
        k_ao_finish<<<(n + 255) / 256, 256, 0, ctx->stream>>>(wf.color.as<float4>(), n, ao.strength, d_out.as<float2>());
        k_tally<<<1, 64, 0, ctx->stream>>>(wf.qcount.as<uint32_t>(), wf.totals.as<uint64_t>(), 0u);
        ctx->launches += 4;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_out.p, (size_t)n * 8, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("trace_ao: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

// Ray-traced reflections: one wave of (specular sample -> extend -> shade -> connect) over the reflection image.
bpt_status wavefront_trace_reflection(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const bpt_reflection_settings& rs, const float* h_depth,
                                      const bpt_gbuffer_texel* h_gbuffer, float* h_refl, float* h_hit) {
    NvtxRange r("RTR Trace + Lighting");                                       // reflection.cpp:317-450
    bpt_status s;
    if ((s = wavefront_alloc(ctx))) return s;
    WavefrontState& wf = ctx->wf;
    wf.ahead_slots = wf.ahead_cursor = 0;
    const uint32_t W = ctx->width, H = ctx->height;
    if (rs.half_resolution && ((W | H) & 1u)) { ctx->err = "trace_reflection: half resolution needs even width and height (texel-centre reads)"; return BPT_ERR_UNSUPPORTED; }
    const uint32_t rw = rs.half_resolution ? (W + 1) / 2 : W, rh = rs.half_resolution ? (H + 1) / 2 : H, n = rw * rh;     // reflection.cpp:324-325
    const float max_roughness = rs.max_roughness, fade_roughness = std::min(rs.fade_roughness, max_roughness - 0.0001f);     // reflection.cpp:361-362
    bpt_settings st{};
    st.ray_length = rs.range; st.max_bounces = 2; st.nee_mode = BPT_NEE_SHADOW_RAY;
    RenderArgs a;
    if ((s = prepare_args(ctx, a, st, 2))) return s;
    a.cam = cam; a.npx = n; a.nslots = 1; a.frame_base = frame_index; a.probe_mode = 1; a.sp.ibl = rs.ibl;
    DevBuf d_depth, d_gb, d_refl, d_hit;
    auto cleanup = [&]() { dev_free(d_depth); dev_free(d_gb); dev_free(d_refl); dev_free(d_hit); };
    if ((s = dev_upload(ctx, d_depth, h_depth, (size_t)W * H * 4)) || (s = dev_upload(ctx, d_gb, h_gbuffer, (size_t)W * H * sizeof(bpt_gbuffer_texel))) ||
        (s = dev_alloc(ctx, d_refl, (size_t)n * 16)) || (s = dev_alloc(ctx, d_hit, (size_t)n * 16))) { cleanup(); return s; }
    cudaError_t e = cudaMemsetAsync(wf.qcount.p, 0, QN * sizeof(uint32_t), ctx->stream);
    if (e != cudaSuccess) { cleanup(); ctx->err = cudaGetErrorString(e); return BPT_ERR_CUDA; }
    a.ray_o_out = wf.ray_o[0].as<float4>(); a.ray_d_out = wf.ray_d[0].as<float4>(); a.ray_w_out = wf.ray_w[0].as<float4>();
    const uint32_t hr = rs.half_resolution ? 1u : 0u;
    k_rtr_raygen<<<(n + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(a, rw, rh, hr, max_roughness, fade_roughness, rs.strength, d_depth.as<float>(), d_gb.as<bpt_gbuffer_texel>());
    ctx->launches++;
    if ((s = run_bounces(ctx, a, st, 2, n, false))) { cleanup(); return s; }
    k_rtr_finish<<<(n + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(a, rw, rh, hr, max_roughness, fade_roughness, d_depth.as<float>(), d_gb.as<bpt_gbuffer_texel>(),
                                                                         d_refl.as<float4>(), d_hit.as<float4>());
    k_tally<<<1, 64, 0, ctx->stream>>>(wf.qcount.as<uint32_t>(), wf.totals.as<uint64_t>(), 0u);
    ctx->launches += 2;
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_refl, d_refl.p, (size_t)n * 16, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_hit, d_hit.p, (size_t)n * 16, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("trace_reflection: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

bpt_status launch_upscale_half_res(bpt_context* ctx, const bpt_camera& cam, uint32_t frame_index, const float* h_depth, const float* h_nr, const float* h_in, float* h_out) {
    const uint32_t W = ctx->width, H = ctx->height, rw = (W + 1) / 2, rh = (H + 1) / 2;
    DevBuf d_depth, d_nr, d_in, d_out;
    auto cleanup = [&]() { dev_free(d_depth); dev_free(d_nr); dev_free(d_in); dev_free(d_out); };
    bpt_status s;
    if ((s = dev_upload(ctx, d_depth, h_depth, (size_t)W * H * 4)) || (s = dev_upload(ctx, d_nr, h_nr, (size_t)W * H * 16)) ||
        (s = dev_upload(ctx, d_in, h_in, (size_t)rw * rh * 16)) || (s = dev_alloc(ctx, d_out, (size_t)W * H * 16))) { cleanup(); return s; }
    k_upscale_half_res<<<dim3((W + 31) / 32, (H + 7) / 8), dim3(32, 8), 0, ctx->stream>>>(cam, W, H, frame_index, d_depth.as<float>(), d_nr.as<float4>(),
                                                                                     d_in.as<float4>(), d_out.as<float4>());
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, d_out.p, (size_t)W * H * 16, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("upscale_half_res: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

// DDGI-style probe tracing through the same extend / shade / connect kernels (BASELINE configs[4]).
bpt_status wavefront_trace_probes(bpt_context* ctx, const bpt_probe_volume& vol, const float* h_table, uint32_t frame_index, uint32_t num_bounces, float* h_out,
                                  uint32_t first_probe, uint32_t num_probes) {
    NvtxRange r("DDGI Trace GBuffer + Lighting");                              // ddgi.cpp probe update
    bpt_status s;
    if ((s = wavefront_alloc(ctx))) return s;
    WavefrontState& wf = ctx->wf;
    wf.ahead_slots = wf.ahead_cursor = 0;
    const uint64_t all_probes = (uint64_t)vol.probe_counts[0] * vol.probe_counts[1] * vol.probe_counts[2];
    if (all_probes == 0 || all_probes * vol.rays_per_probe > 0xffffffffull || vol.rays_per_probe == 0) { ctx->err = "trace_probes: bad volume"; return BPT_ERR_INVALID; }
    if (num_probes == 0xffffffffu) { first_probe = 0; num_probes = (uint32_t)all_probes; }          // the whole volume
    if ((uint64_t)first_probe + num_probes > all_probes) { ctx->err = "trace_probes: probe range outside the volume"; return BPT_ERR_INVALID; }
    if (num_probes == 0) return BPT_OK;
    const uint64_t first_path = (uint64_t)first_probe * vol.rays_per_probe;                           // global path id = probe * rays_per_probe + ray
    const uint64_t total = (uint64_t)num_probes * vol.rays_per_probe;
    const uint32_t B = std::min(std::max(num_bounces, 1u), 15u) + 1;          // num_bounces extend passes
    bpt_settings st{};
    st.ray_length = vol.ray_length; st.max_bounces = B; st.nee_mode = BPT_NEE_SHADOW_RAY;
    RenderArgs a;
    if ((s = prepare_args(ctx, a, st, B))) return s;
    a.sp.diffuse_only = 1; a.probe_mode = 1; a.nslots = 1; a.frame_base = frame_index;
    memset(&a.cam, 0, sizeof(a.cam));
    DevBuf table;
    if ((s = dev_upload(ctx, table, h_table, 8192 * sizeof(float2)))) return s;
    for (uint64_t base = 0; base < total; base += wf.capacity) {
        const uint32_t count = (uint32_t)std::min<uint64_t>(wf.capacity, total - base);
        a.npx = count; a.pixel_base = (uint32_t)(first_path + base);
        cudaError_t e = cudaMemsetAsync(wf.qcount.p, 0, QN * sizeof(uint32_t), ctx->stream);
        if (e != cudaSuccess) { dev_free(table); ctx->err = cudaGetErrorString(e); return BPT_ERR_CUDA; }
        a.ray_o_out = wf.ray_o[0].as<float4>(); a.ray_d_out = wf.ray_d[0].as<float4>(); a.ray_w_out = wf.ray_w[0].as<float4>();
        k_probe_raygen<<<(count + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(a, vol, table.as<float2>());
        ctx->launches++;
        if ((s = run_bounces(ctx, a, st, B, count, false))) { dev_free(table); return s; }
        k_tally<<<1, 64, 0, ctx->stream>>>(wf.qcount.as<uint32_t>(), wf.totals.as<uint64_t>(), count);
        ctx->launches++;
        e = cudaMemcpyAsync(h_out + 4 * base, wf.color.p, (size_t)count * 16, cudaMemcpyDefault, ctx->stream);      // host or device destination
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { dev_free(table); ctx->err = cudaGetErrorString(e); return BPT_ERR_CUDA; }
    }
    dev_free(table);
    return BPT_OK;
}

bpt_status wavefront_accumulate_ahead(bpt_context* ctx, uint32_t count) {
    NvtxRange r("PT Accumulate");
    WavefrontState& wf = ctx->wf;
    if (count == 0 || wf.ahead_cursor + count > wf.ahead_slots) { ctx->err = "accumulate_ahead: not enough prefetched samples"; return BPT_ERR_STATE; }
    const uint32_t npx = ctx->width * ctx->height;
    if (wf.accum_fp16) {
        LAUNCH_T(ctx, 4, k_accumulate_fp16, (npx + 255) / 256, 256, wf.color.as<float4>(), wf.accum.as<float4>(), npx, wf.ahead_cursor, wf.ahead_cursor + count, wf.accum_count);
        wf.accum_count += count;
    } else LAUNCH_T(ctx, 4, k_accumulate, (npx + 255) / 256, 256, wf.color.as<float4>(), wf.accum.as<float4>(), npx, wf.ahead_cursor, wf.ahead_cursor + count);
    wf.ahead_cursor += count;
    return BPT_OK;
}

// bpt_accumulate_ahead(ctx, 1) + bpt_resolve_device_rgba16f(ctx, total_samples, d_out) in one launch (FP32 state; the literal fp16 state
// keeps its two kernels: its sum buffer already is the running average)
bpt_status wavefront_accumulate_ahead_rgba16f(bpt_context* ctx, uint32_t total_samples, void* d_out) {
    WavefrontState& wf = ctx->wf;
    if (wf.accum_fp16) {
        bpt_status s = wavefront_accumulate_ahead(ctx, 1);
        return s ? s : launch_resolve_rgba16f(ctx, total_samples, d_out);
    }
    NvtxRange r("PT Accumulate");
    if (wf.ahead_cursor + 1 > wf.ahead_slots) { ctx->err = "accumulate_ahead: not enough prefetched samples"; return BPT_ERR_STATE; }
    const uint32_t npx = ctx->width * ctx->height;
    LAUNCH_T(ctx, 4, k_accumulate_resolve_rgba16f, (npx + 255) / 256, 256, wf.color.as<float4>(), wf.accum.as<float4>(), reinterpret_cast<uint2*>(d_out), npx,
             wf.ahead_cursor, 1.0f / (float)total_samples);
    wf.ahead_cursor += 1;
    return BPT_OK;
}

bpt_status launch_resolve(bpt_context* ctx, uint32_t total_samples, float* d_out) {
    uint32_t npx = ctx->width * ctx->height;
    float inv = ctx->wf.accum_fp16 ? 1.0f : 1.0f / (float)total_samples;      // reference_fp16: the buffer already is the running average
    LAUNCH(ctx, k_resolve, (npx + 255) / 256, 256, ctx->wf.accum.as<float4>(), reinterpret_cast<float4*>(d_out), npx, inv);
    return BPT_OK;
}

bpt_status launch_resolve_rgba16f(bpt_context* ctx, uint32_t total_samples, void* d_out) {
    uint32_t npx = ctx->width * ctx->height;
    float inv = ctx->wf.accum_fp16 ? 1.0f : 1.0f / (float)total_samples;
    LAUNCH(ctx, k_resolve_rgba16f, (npx + 255) / 256, 256, ctx->wf.accum.as<float4>(), reinterpret_cast<uint2*>(d_out), npx, inv);
    return BPT_OK;
}

bpt_status launch_trace_batch(bpt_context* ctx, const bpt_ray* h_rays, uint64_t n, uint32_t frame_index, bpt_hit* h_hits, uint8_t* h_visible) {
    if (n == 0) return BPT_OK;
    DevBuf rays, out;
    bpt_status s;
    if ((s = dev_upload(ctx, rays, h_rays, n * sizeof(bpt_ray)))) return s;
    size_t out_bytes = h_hits ? n * sizeof(bpt_hit) : n;
    if ((s = dev_alloc(ctx, out, out_bytes))) { dev_free(rays); return s; }
    DScene sc = ctx->scene_view();
    k_trace_batch<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, ctx->stream>>>(sc, rays.as<bpt_ray>(), n, frame_index,
        h_hits ? out.as<bpt_hit>() : nullptr, h_hits ? nullptr : out.as<uint8_t>(), ctx->d_instances.as<DInstance>(),
        use_wide(ctx) ? ctx->blas[0].wide.as<float4>() : nullptr, use_wide(ctx) ? ctx->blas[0].leafbox.as<float4>() : nullptr,
        use_wide2(ctx) ? 1u : 0u);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_hits ? (void*)h_hits : (void*)h_visible, out.p, out_bytes, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(rays); dev_free(out);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

bpt_status launch_ddgi_lighting(bpt_context* ctx, uint64_t n, const float* h_pos, const float* h_normal, const float* h_view, float* h_out) {
    if (n == 0) return BPT_OK;
    DevBuf pos, nrm, view, out;
    auto cleanup = [&]() { dev_free(pos); dev_free(nrm); dev_free(view); dev_free(out); };
    bpt_status s;
    if ((s = dev_upload(ctx, pos, h_pos, n * 12)) || (s = dev_upload(ctx, nrm, h_normal, n * 12)) || (s = dev_upload(ctx, view, h_view, n * 12)) ||
        (s = dev_alloc(ctx, out, n * 16))) { cleanup(); return s; }
    k_ddgi_lighting<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->scene_view(), n, pos.as<float>(), nrm.as<float>(), view.as<float>(), out.as<float4>());
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out, out.p, n * 16, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("ddgi_lighting: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}

bpt_status launch_blend_probes(bpt_context* ctx, const bpt_probe_volume& vol, const float* h_table, uint32_t frame_index, const float* h_rays,
                               const bpt_probe_blend& bl, float* h_irr, float* h_vis) {
    const uint32_t nprobes = vol.probe_counts[0] * vol.probe_counts[1] * vol.probe_counts[2];
    const uint64_t nrays = (uint64_t)nprobes * vol.rays_per_probe;
    if (!nprobes || !vol.rays_per_probe || bl.irradiance_size < 2 || bl.visibility_size < 2 || bl.irradiance_size > 30 || bl.visibility_size > 30 ||
        vol.rays_per_probe > 1024) { ctx->err = "blend_probes: bad sizes"; return BPT_ERR_INVALID; }
    const size_t irr_bytes = (size_t)vol.probe_counts[0] * vol.probe_counts[1] * (bl.irradiance_size + 2) * vol.probe_counts[2] * (bl.irradiance_size + 2) * 16;
    const size_t vis_bytes = (size_t)vol.probe_counts[0] * vol.probe_counts[1] * (bl.visibility_size + 2) * vol.probe_counts[2] * (bl.visibility_size + 2) * 8;
    DevBuf table, rays, irr, vis;
    bpt_status s = BPT_OK;
    auto cleanup = [&]() { dev_free(table); dev_free(rays); dev_free(irr); dev_free(vis); };
    if ((s = dev_upload(ctx, table, h_table, 8192 * sizeof(float2))) || (s = dev_upload(ctx, rays, h_rays, nrays * 16)) ||
        (s = dev_upload(ctx, irr, h_irr, irr_bytes)) || (s = dev_upload(ctx, vis, h_vis, vis_bytes))) { cleanup(); return s; }
    const size_t smem = (size_t)vol.rays_per_probe * (16 + 12);
    auto threads = [](uint32_t size) { return ((size * size + 31) / 32) * 32; };
    k_probe_blend<false><<<nprobes, threads(bl.irradiance_size), smem, ctx->stream>>>(vol, table.as<float2>(), frame_index, rays.as<float4>(), bl.irradiance_size,
                                                                                     bl.alpha, bl.history_valid, irr.as<float>());
    k_probe_blend<true><<<nprobes, threads(bl.visibility_size), smem, ctx->stream>>>(vol, table.as<float2>(), frame_index, rays.as<float4>(), bl.visibility_size,
                                                                                    bl.alpha, bl.history_valid, vis.as<float>());
    ctx->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_irr, irr.p, irr_bytes, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_vis, vis.p, vis_bytes, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cleanup();
    if (e != cudaSuccess) { ctx->err = std::string("blend_probes: ") + cudaGetErrorString(e); return BPT_ERR_CUDA; }
    return BPT_OK;
}
