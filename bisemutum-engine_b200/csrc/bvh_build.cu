// bvh_build.cu — LBVH construction on the GPU (sm_100a), all kernels hand-written.
//
// Replaces the driver-side BLAS/TLAS build of the reference (AccelerationStructure ctor,
// bisemutum/src/graphics/accel.cpp:11-159; BLAS desc graphics_manager.cpp:616-654; instance
// records accel.cpp:104-132). NEW algorithm, defined by DESIGN.md "LBVH" and checked bit-exactly
// against oracle/oracle_bvh.cpp:
//   primitive AABBs → bounds → 63-bit Morton of AABB centroids → stable LSD radix sort (8 x 8 bit)
//   → Karras 2012 hierarchy (64-bit codes, index tie-break) → bottom-up refit (atomic arrival flags)
//   → 64-B nodes (both child boxes in the parent) + 48-B triangle records in leaf order.
// All passes are streaming and HBM-bound; build time is reported separately from render time.
#include <cfloat>
#include "bpt_internal.cuh"
#include "bpt_wide.cuh"

using namespace bptd;

namespace {

constexpr int kThreads = 256;
inline unsigned grid_for(uint64_t n, int threads = kThreads) { return (unsigned)((n + threads - 1) / threads); }

// ---- ordered-uint encoding of floats for atomic min/max ---------------------------------------
__device__ __forceinline__ uint32_t f_key(float f) { uint32_t k = __float_as_uint(f); return (k & 0x80000000u) ? ~k : (k | 0x80000000u); }
__device__ __forceinline__ float key_f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void k_init_bounds(uint32_t* keys) {
    if (threadIdx.x < 3) keys[threadIdx.x] = 0xffffffffu;       // min keys
    else if (threadIdx.x < 6) keys[threadIdx.x] = 0u;           // max keys
}

// Triangle gather: object space (inst == nullptr: one BLAS, `ranges` has one entry) or world space (merged mode: every
// instance's triangles through its o2w; `ranges[slot]` = first output record of instance `slot`, found by binary search).
// Writes unsorted 48-B records (v0|prim, v1|slot, v2|any-hit) and the AABB.
struct GatherRange { bpt_blas_desc bd; uint32_t base; };
__global__ void k_gather_tris(const float* __restrict__ positions, const uint32_t* __restrict__ indices, const GatherRange* __restrict__ ranges, GatherRange single,
                              uint32_t num_ranges, uint32_t total, const DInstance* __restrict__ inst,
                              float4* __restrict__ raw, float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint32_t slot = 0, end = num_ranges;
    while (end - slot > 1) { uint32_t m = (slot + end) >> 1; if (ranges[m].base <= g) slot = m; else end = m; }
    const bpt_blas_desc bd = ranges ? ranges[slot].bd : single.bd;       // (one BLAS: the range travels as a kernel argument)
    const uint32_t k = g - (ranges ? ranges[slot].base : 0u);
    float3 v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        uint32_t idx = indices[(size_t)bd.index_offset + 3ull * k + c];
        const float* p = positions + (size_t)bd.position_offset + 3ull * idx;
        v[c] = v3(p[0], p[1], p[2]);
        if (inst) v[c] = xf_point(inst[slot].o2w, v[c]);
    }
    raw[3 * (size_t)g + 0] = make_float4(v[0].x, v[0].y, v[0].z, __uint_as_float(k));
    raw[3 * (size_t)g + 1] = make_float4(v[1].x, v[1].y, v[1].z, __uint_as_float(slot));
    raw[3 * (size_t)g + 2] = make_float4(v[2].x, v[2].y, v[2].z, __uint_as_float(inst ? inst[slot].anyhit : 0u));
    float3 l = vmin(vmin(v[0], v[1]), v[2]), h = vmax(vmax(v[0], v[1]), v[2]);
    lo[g] = make_float4(l.x, l.y, l.z, 0.0f);
    hi[g] = make_float4(h.x, h.y, h.z, 0.0f);
}

__global__ void k_reduce_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, uint32_t* keys) {
    float3 l = v3s(FLT_MAX), h = v3s(-FLT_MAX);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = lo[i], b = hi[i];
        l = vmin(l, v3(a.x, a.y, a.z)); h = vmax(h, v3(b.x, b.y, b.z));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l.x = fminf(l.x, __shfl_xor_sync(0xffffffffu, l.x, o)); l.y = fminf(l.y, __shfl_xor_sync(0xffffffffu, l.y, o)); l.z = fminf(l.z, __shfl_xor_sync(0xffffffffu, l.z, o));
        h.x = fmaxf(h.x, __shfl_xor_sync(0xffffffffu, h.x, o)); h.y = fmaxf(h.y, __shfl_xor_sync(0xffffffffu, h.y, o)); h.z = fmaxf(h.z, __shfl_xor_sync(0xffffffffu, h.z, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&keys[0], f_key(l.x)); atomicMin(&keys[1], f_key(l.y)); atomicMin(&keys[2], f_key(l.z));
        atomicMax(&keys[3], f_key(h.x)); atomicMax(&keys[4], f_key(h.y)); atomicMax(&keys[5], f_key(h.z));
    }
}
__global__ void k_decode_bounds(const uint32_t* keys, float* out6) {
    if (threadIdx.x < 6) out6[threadIdx.x] = key_f(keys[threadIdx.x]);
}

__device__ __forceinline__ uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__device__ __forceinline__ uint32_t quant21(float c, float lo, float hi) {
    float ext = hi - lo;
    float t = ext > 0.0f ? (c - lo) / ext : 0.0f;
    float s = t * 2097152.0f;
    s = tmax_(s, 0.0f);
    s = tmin_(s, 2097151.0f);
    return (uint32_t)s;
}
__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, const float* __restrict__ b6,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = lo[i], b = hi[i];
    float3 c = (v3(a.x, a.y, a.z) + v3(b.x, b.y, b.z)) * 0.5f;
    keys[i] = (expand21(quant21(c.x, b6[0], b6[3])) << 2) | (expand21(quant21(c.y, b6[1], b6[4])) << 1) | expand21(quant21(c.z, b6[2], b6[5]));
    vals[i] = i;
}

// ---- stable LSD radix sort, 8 bits per pass ---------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;

__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortItems; r++) {
        uint32_t i = base + r * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];     // digit-major
}
// exclusive scan of 256 values held one per thread of a 256-thread block
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t* warp_sums /* [8] shared */, uint32_t* total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { uint32_t t = warp_sums[w]; if ((uint32_t)w < warp) before += t; all += t; }
    __syncthreads();
    if (total) *total = all;
    return x - v + before;
}
// Block d turns the per-block counts of digit d (hist is digit-major) into exclusive offsets WITHIN the digit and
// writes the digit's total; the scatter kernel adds the prefix over the 256 digit totals itself.
__global__ void __launch_bounds__(kSortThreads) k_sort_scan(uint32_t* __restrict__ hist, uint32_t nblocks, uint32_t* __restrict__ digit_total) {
    __shared__ uint32_t warp_sums[8];
    uint32_t* row = hist + (size_t)blockIdx.x * nblocks;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblocks; base += kSortThreads) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? row[i] : 0, tot;
        uint32_t e = block_excl_scan256(v, warp_sums, &tot);
        if (i < nblocks) row[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                               uint32_t n, int shift, const uint32_t* __restrict__ hist, uint32_t nblocks,
                                                               const uint32_t* __restrict__ digit_total) {
    __shared__ uint32_t digit_base[256];                       // global offset of this block's first key of each digit
    __shared__ uint32_t warp_count[kSortThreads / 32][256];
    __shared__ uint32_t warp_sums[8];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    digit_base[threadIdx.x] = block_excl_scan256(digit_total[threadIdx.x], warp_sums, nullptr) + hist[threadIdx.x * nblocks + blockIdx.x];
    uint32_t base = blockIdx.x * kSortTile;
    for (int r = 0; r < kSortItems; r++) {
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; w++) warp_count[w][threadIdx.x] = 0;
        __syncthreads();
        uint32_t i = base + r * kSortThreads + threadIdx.x;
        bool valid = i < n;
        uint64_t key = valid ? keys_in[i] : 0;
        uint32_t d = valid ? ((uint32_t)(key >> shift) & 0xffu) : (0x100u + lane);
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_count[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t off = digit_base[d] + rank;
            for (uint32_t w = 0; w < warp; w++) off += warp_count[w][d];
            keys_out[off] = key;
            vals_out[off] = vals_in[i];
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; w++) tot += warp_count[w][threadIdx.x];
        digit_base[threadIdx.x] += tot;
        __syncthreads();
    }
}

// ---- Karras 2012 -----------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ k, uint32_t n, int64_t i, int64_t j) {
    if (j < 0 || j >= (int64_t)n) return -1;
    uint64_t a = k[i], c = k[j];
    return a != c ? __clzll((long long)(a ^ c)) : 64 + __clz((int)((uint32_t)i ^ (uint32_t)j));
}
__device__ __forceinline__ void karras_node(const uint64_t* __restrict__ keys, uint32_t n, int64_t i, int32_t* __restrict__ child0, int32_t* __restrict__ child1,
                                            int32_t* __restrict__ node_parent, int32_t* __restrict__ leaf_parent) {
    if (i == 0) node_parent[0] = -1;
    int d = delta(keys, n, i, i + 1) > delta(keys, n, i, i - 1) ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int64_t j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int64_t s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int64_t gamma = i + s * d + (d < 0 ? d : 0);
    int64_t lo_i = i < j ? i : j, hi_i = i < j ? j : i;
    int32_t left = (lo_i == gamma) ? ~(int32_t)gamma : (int32_t)gamma;
    int32_t right = (hi_i == gamma + 1) ? ~(int32_t)(gamma + 1) : (int32_t)(gamma + 1);
    child0[i] = left; child1[i] = right;
    if (left >= 0) node_parent[left] = (int32_t)i; else leaf_parent[~left] = (int32_t)i;
    if (right >= 0) node_parent[right] = (int32_t)i; else leaf_parent[~right] = (int32_t)i;
}
__global__ void k_karras(const uint64_t* __restrict__ keys, uint32_t n, int32_t* __restrict__ child0, int32_t* __restrict__ child1,
                         int32_t* __restrict__ node_parent, int32_t* __restrict__ leaf_parent) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n - 1) return;
    karras_node(keys, n, i, child0, child1, node_parent, leaf_parent);
}

// ---- bottom-up refit ---------------------------------------------------------------------------
// Leaf j walks towards the root; at every node the SECOND arrival (atomic flag) merges the two child boxes and goes on.
__device__ __forceinline__ void refit_from_leaf(uint32_t j, const uint32_t* __restrict__ prims, const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                                                const int32_t* __restrict__ child0, const int32_t* __restrict__ child1, const int32_t* __restrict__ node_parent,
                                                const int32_t* __restrict__ leaf_parent, float4* node_lo, float4* node_hi, uint32_t* flags, float4* __restrict__ nodes) {
    int32_t node = leaf_parent[j];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&flags[node], 1u) == 0u) return;     // first arrival: the sibling subtree is not done yet
        __threadfence();
        int32_t c0 = child0[node], c1 = child1[node];
        float4 l0, h0, l1, h1;
        if (c0 < 0) { uint32_t p = prims[~c0]; l0 = prim_lo[p]; h0 = prim_hi[p]; } else { l0 = __ldcg(&node_lo[c0]); h0 = __ldcg(&node_hi[c0]); }
        if (c1 < 0) { uint32_t p = prims[~c1]; l1 = prim_lo[p]; h1 = prim_hi[p]; } else { l1 = __ldcg(&node_lo[c1]); h1 = __ldcg(&node_hi[c1]); }
        float4* out = nodes + 4 * (size_t)node;
        out[0] = make_float4(l0.x, h0.x, l0.y, h0.y);
        out[1] = make_float4(l1.x, h1.x, l1.y, h1.y);
        out[2] = make_float4(l0.z, h0.z, l1.z, h1.z);
        out[3] = make_float4(__int_as_float(c0), __int_as_float(c1), __int_as_float(node_parent[node]), 0.0f);
        float3 ul = vmin(v3(l0.x, l0.y, l0.z), v3(l1.x, l1.y, l1.z)), uh = vmax(v3(h0.x, h0.y, h0.z), v3(h1.x, h1.y, h1.z));
        __stcg(&node_lo[node], make_float4(ul.x, ul.y, ul.z, 0.0f));
        __stcg(&node_hi[node], make_float4(uh.x, uh.y, uh.z, 0.0f));
        node = node_parent[node];
    }
}
__global__ void k_refit(uint32_t n, const uint32_t* __restrict__ prims, const float4* __restrict__ prim_lo, const float4* __restrict__ prim_hi,
                        const int32_t* __restrict__ child0, const int32_t* __restrict__ child1, const int32_t* __restrict__ node_parent,
                        const int32_t* __restrict__ leaf_parent, float4* node_lo, float4* node_hi, uint32_t* flags, float4* __restrict__ nodes) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    refit_from_leaf(j, prims, prim_lo, prim_hi, child0, child1, node_parent, leaf_parent, node_lo, node_hi, flags, nodes);
}

// ---- small trees (TLAS, small BLASes): the whole build in ONE block ------------------------------------------------
// Same definition, same result: bounds (min/max are exact in any order), the same Morton codes, a bitonic sort of the
// composite (code, primitive index) — whose order is unique, i.e. the stable LSD sort's — in shared memory, then the Karras
// and refit steps above separated by block barriers. One launch instead of 33: the reference rebuilds its TLAS every frame
// (accel.cpp:134-159), and a 512-instance TLAS was 0.2 ms of launch latency.
constexpr uint32_t kSmallMax = 8192;      // 96 KB of shared memory at the upper end
constexpr int kSmallThreads = 1024;
__global__ void __launch_bounds__(kSmallThreads) k_lbvh_small(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, uint32_t np2,
                                                              float* __restrict__ bounds6, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ prims_out,
                                                              int32_t* __restrict__ child0, int32_t* __restrict__ child1, int32_t* __restrict__ node_parent,
                                                              int32_t* __restrict__ leaf_parent, float4* node_lo, float4* node_hi, uint32_t* flags,
                                                              float4* __restrict__ nodes) {
    extern __shared__ uint64_t s_keys[];                     // np2 codes, then np2 primitive indices
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_keys + np2);
    __shared__ uint32_t s_bkeys[6];
    __shared__ float s_b6[6];
    const uint32_t tid = threadIdx.x;
    if (tid < 3) s_bkeys[tid] = 0xffffffffu; else if (tid < 6) s_bkeys[tid] = 0u;
    __syncthreads();
    {
        float3 l = v3s(FLT_MAX), h = v3s(-FLT_MAX);
        for (uint32_t i = tid; i < n; i += kSmallThreads) {
            float4 a = lo[i], b = hi[i];
            l = vmin(l, v3(a.x, a.y, a.z)); h = vmax(h, v3(b.x, b.y, b.z));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l.x = fminf(l.x, __shfl_xor_sync(0xffffffffu, l.x, o)); l.y = fminf(l.y, __shfl_xor_sync(0xffffffffu, l.y, o)); l.z = fminf(l.z, __shfl_xor_sync(0xffffffffu, l.z, o));
            h.x = fmaxf(h.x, __shfl_xor_sync(0xffffffffu, h.x, o)); h.y = fmaxf(h.y, __shfl_xor_sync(0xffffffffu, h.y, o)); h.z = fmaxf(h.z, __shfl_xor_sync(0xffffffffu, h.z, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(&s_bkeys[0], f_key(l.x)); atomicMin(&s_bkeys[1], f_key(l.y)); atomicMin(&s_bkeys[2], f_key(l.z));
            atomicMax(&s_bkeys[3], f_key(h.x)); atomicMax(&s_bkeys[4], f_key(h.y)); atomicMax(&s_bkeys[5], f_key(h.z));
        }
    }
    __syncthreads();
    if (tid < 6) { float v = key_f(s_bkeys[tid]); s_b6[tid] = v; bounds6[tid] = v; }
    __syncthreads();
    for (uint32_t i = tid; i < np2; i += kSmallThreads) {
        uint64_t key = ~0ull; uint32_t idx = 0xffffffffu;       // padding sorts behind every 63-bit code
        if (i < n) {
            float4 a = lo[i], b = hi[i];
            float3 c = (v3(a.x, a.y, a.z) + v3(b.x, b.y, b.z)) * 0.5f;
            key = (expand21(quant21(c.x, s_b6[0], s_b6[3])) << 2) | (expand21(quant21(c.y, s_b6[1], s_b6[4])) << 1) | expand21(quant21(c.z, s_b6[2], s_b6[5]));
            idx = i;
        }
        s_keys[i] = key; s_idx[i] = idx;
    }
    __syncthreads();
    for (uint32_t k = 2; k <= np2; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < np2; i += kSmallThreads) {
                uint32_t x = i ^ j;
                if (x > i) {
                    uint64_t ka = s_keys[i], kb = s_keys[x]; uint32_t ia = s_idx[i], ib = s_idx[x];
                    bool a_after_b = ka > kb || (ka == kb && ia > ib);
                    bool up = (i & k) == 0;
                    if (a_after_b == up) { s_keys[i] = kb; s_keys[x] = ka; s_idx[i] = ib; s_idx[x] = ia; }
                }
            }
            __syncthreads();
        }
    for (uint32_t i = tid; i < n; i += kSmallThreads) { keys_out[i] = s_keys[i]; prims_out[i] = s_idx[i]; }
    if (n < 2) return;
    for (uint32_t i = tid; i < n - 1; i += kSmallThreads) flags[i] = 0u;
    for (uint32_t i = tid; i < n - 1; i += kSmallThreads) karras_node(s_keys, n, (int64_t)i, child0, child1, node_parent, leaf_parent);
    __syncthreads();                                         // (block barrier: global writes of this block are visible to it)
    for (uint32_t j = tid; j < n; j += kSmallThreads)
        refit_from_leaf(j, s_idx, lo, hi, child0, child1, node_parent, leaf_parent, node_lo, node_hi, flags, nodes);
}

// ---- 4-wide quantised nodes + exact leaf boxes from the refitted binary tree (bpt_wide.cuh): one thread per binary node ----
__global__ void k_collapse4(uint32_t n_internal, const float4* __restrict__ nodes2, float4* __restrict__ wide, float4* __restrict__ leafbox) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    float4 out[4];
    collapse_node4(nodes2, (int32_t)i, out);
    wide[4 * (size_t)i] = out[0]; wide[4 * (size_t)i + 1] = out[1]; wide[4 * (size_t)i + 2] = out[2]; wide[4 * (size_t)i + 3] = out[3];
    leaf_boxes_of_node(nodes2, (int32_t)i, leafbox);
}

__global__ void k_emit_tris(uint32_t n, const uint32_t* __restrict__ prims, const float4* __restrict__ raw, float4* __restrict__ tris) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    size_t p = prims[j];
    float4 a = raw[3 * p], b = raw[3 * p + 1], c = raw[3 * p + 2];
    float3 v0 = v3(a.x, a.y, a.z);
    float3 e1 = v3(b.x, b.y, b.z) - v0, e2 = v3(c.x, c.y, c.z) - v0;
    tris[3 * (size_t)j + 0] = a;                                            // v0 | prim
    tris[3 * (size_t)j + 1] = make_float4(e1.x, e1.y, e1.z, b.w);           // e1 | instance slot
    tris[3 * (size_t)j + 2] = make_float4(e2.x, e2.y, e2.z, c.w);           // e2 | any-hit flag (merged mode)
}

// ---- instances ---------------------------------------------------------------------------------
__global__ void k_make_instances(const bpt_instance_desc* __restrict__ desc, uint32_t n, const bpt_drawable_sbt_data* __restrict__ drawables,
                                 const bpt_material* __restrict__ materials, DInstance* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DInstance d;
    const float* t = &desc[i].transform[0][0];
#pragma unroll
    for (int k = 0; k < 12; k++) d.o2w[k] = t[k];
    invert_3x4(d.o2w, d.w2o);
    d.instance_id = desc[i].instance_id_and_mask & 0xffffffu;
    d.flags = desc[i].sbt_offset_and_flags >> 24;
    d.blas = (uint32_t)desc[i].blas;
    for (int k = 0; k < 4; k++) d.pad[k] = 0;
    // any-hit needed? (non-opaque instance AND a non-opaque blend mode: hits/rt_gbuffer_hit.hlsl:20-35, accel.cpp:112-116)
    uint32_t blend = (materials[drawables[d.instance_id].material_offset / (uint32_t)sizeof(bpt_material)].flags >> BPT_MATERIAL_BLEND_SHIFT) & 0xffu;
    d.anyhit = ((d.flags & BPT_INSTANCE_FORCE_NON_OPAQUE) && blend != BPT_BLEND_OPAQUE) ? 1u : 0u;
    out[i] = d;
}
// world AABB of an instance = min/max of the 8 transformed corners of its BLAS bounds
// (Transform::transform_bounding_box, bisemutum/src/math/transform.cpp:59-78)
__global__ void k_instance_bounds(const DInstance* __restrict__ inst, uint32_t n, const float* __restrict__ blas_bounds6, float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = blas_bounds6 + 6 * (size_t)inst[i].blas;
    float3 mn = v3s(FLT_MAX), mx = v3s(-FLT_MAX);
#pragma unroll
    for (int c = 0; c < 8; c++) {
        float3 p = v3((c & 4) ? b[3] : b[0], (c & 2) ? b[4] : b[1], (c & 1) ? b[5] : b[2]);
        float3 w = xf_point(inst[i].o2w, p);
        mn = vmin(mn, w); mx = vmax(mx, w);
    }
    lo[i] = make_float4(mn.x, mn.y, mn.z, 0.0f);
    hi[i] = make_float4(mx.x, mx.y, mx.z, 0.0f);
}

// Build scratch from the context's grow-only arena (bpt_internal.cuh): bump allocation, released in stack order when the
// Scratch goes out of scope. Everything runs on ctx->stream, so re-use by the next build is stream-ordered.
struct Scratch {
    bpt_context* ctx;
    size_t chunk0, offset0;
    explicit Scratch(bpt_context* c) : ctx(c), chunk0(c->arena_chunk), offset0(c->arena_offset) {}
    ~Scratch() { if (!ctx->arena_hold) { ctx->arena_chunk = chunk0; ctx->arena_offset = offset0; } }
    bpt_status get(bpt_context*, DevBuf& out, size_t bytes) {
        bytes = (std::max<size_t>(bytes, 16) + 255) & ~(size_t)255;
        auto& ch = ctx->arena_chunks;
        while (ctx->arena_chunk < ch.size() && ctx->arena_offset + bytes > ch[ctx->arena_chunk].bytes) { ctx->arena_chunk++; ctx->arena_offset = 0; }
        if (ctx->arena_chunk == ch.size()) {
            DevBuf b; bpt_status s = dev_alloc(ctx, b, std::max<size_t>(bytes, (size_t)32 << 20)); if (s) return s;
            ch.push_back(b); ctx->arena_offset = 0;
        }
        out.p = (char*)ch[ctx->arena_chunk].p + ctx->arena_offset; out.bytes = bytes;
        ctx->arena_offset += bytes;
        if (ctx->arena_hold) ctx->arena_held_bytes += bytes;
        return BPT_OK;
    }
};

} // namespace

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

// `d_bounds6` (device, 6 floats) receives the bounds of the primitives: they stay on the device (the TLAS build reads the
// BLAS bounds there), so a build never waits for the GPU.
bpt_status lbvh_build(bpt_context* ctx, DevBvh& out, uint32_t n, const float4* d_lo, const float4* d_hi, float* d_bounds6) {
    out.n = n;
    out.root = n == 1 ? ~0 : 0;
    Scratch sc(ctx);
    bpt_status s;
    if ((s = dev_reserve(ctx, out.morton, (size_t)n * 8))) return s;
    if ((s = dev_reserve(ctx, out.prims, (size_t)n * 4))) return s;
    DevBuf c0, c1, np, lp, nlo, nhi, flags;
    const uint32_t ni = n >= 2 ? n - 1 : 1;
    if (n >= 2 && (s = dev_reserve(ctx, out.nodes, (size_t)(n - 1) * 64))) return s;
    if ((s = sc.get(ctx, c0, (size_t)ni * 4))) return s;
    if ((s = sc.get(ctx, c1, (size_t)ni * 4))) return s;
    if ((s = sc.get(ctx, np, (size_t)ni * 4))) return s;
    if ((s = sc.get(ctx, lp, (size_t)n * 4))) return s;
    if ((s = sc.get(ctx, nlo, (size_t)ni * 16))) return s;
    if ((s = sc.get(ctx, nhi, (size_t)ni * 16))) return s;
    if ((s = sc.get(ctx, flags, (size_t)ni * 4))) return s;
    if (n <= kSmallMax) {                        // TLAS / small BLAS: one block does everything
        uint32_t np2 = 2; while (np2 < n) np2 <<= 1;
        const size_t smem = (size_t)np2 * 12;
        // the opt-in is per device (one process may hold contexts on several GPUs): remembered in the context, not in a process-wide flag
        if (!ctx->lbvh_small_attr_set) { BPT_CUDA_TRY(ctx, cudaFuncSetAttribute(k_lbvh_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmallMax * 12))); ctx->lbvh_small_attr_set = true; }
        k_lbvh_small<<<1, kSmallThreads, smem, ctx->stream>>>(d_lo, d_hi, n, np2, d_bounds6, out.morton.as<uint64_t>(), out.prims.as<uint32_t>(), c0.as<int32_t>(),
                                                              c1.as<int32_t>(), np.as<int32_t>(), lp.as<int32_t>(), nlo.as<float4>(), nhi.as<float4>(),
                                                              flags.as<uint32_t>(), out.nodes.as<float4>());
        ctx->launches++;
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) { ctx->err = std::string("launch of k_lbvh_small: ") + cudaGetErrorString(le); return BPT_ERR_CUDA; }
        return BPT_OK;
    }
    DevBuf bkeys;
    if ((s = sc.get(ctx, bkeys, 6 * sizeof(uint32_t)))) return s;
    LAUNCH(ctx, k_init_bounds, 1, 32, bkeys.as<uint32_t>());
    LAUNCH(ctx, k_reduce_bounds, std::min(grid_for(n), 1024u), kThreads, d_lo, d_hi, n, bkeys.as<uint32_t>());
    LAUNCH(ctx, k_decode_bounds, 1, 32, bkeys.as<uint32_t>(), d_bounds6);
    // sort
    DevBuf keys2, vals2;
    if ((s = sc.get(ctx, keys2, (size_t)n * 8))) return s;
    if ((s = sc.get(ctx, vals2, (size_t)n * 4))) return s;
    LAUNCH(ctx, k_morton, grid_for(n), kThreads, d_lo, d_hi, n, d_bounds6, out.morton.as<uint64_t>(), out.prims.as<uint32_t>());
    uint32_t nblocks = (n + kSortTile - 1) / kSortTile;
    DevBuf hist, dtot;
    if ((s = sc.get(ctx, hist, (size_t)256 * nblocks * 4))) return s;
    if ((s = sc.get(ctx, dtot, 256 * 4))) return s;
    uint64_t* ka = out.morton.as<uint64_t>(); uint64_t* kb = keys2.as<uint64_t>();
    uint32_t* va = out.prims.as<uint32_t>(); uint32_t* vb = vals2.as<uint32_t>();
    for (int pass = 0; pass < 8; pass++) {       // 63-bit keys: all 8 bytes
        int shift = pass * 8;
        LAUNCH(ctx, k_sort_hist, nblocks, kSortThreads, ka, n, shift, hist.as<uint32_t>(), nblocks);
        LAUNCH(ctx, k_sort_scan, 256, kSortThreads, hist.as<uint32_t>(), nblocks, dtot.as<uint32_t>());
        LAUNCH(ctx, k_sort_scatter, nblocks, kSortThreads, ka, va, kb, vb, n, shift, hist.as<uint32_t>(), nblocks, dtot.as<uint32_t>());
        std::swap(ka, kb); std::swap(va, vb);
    }   // even number of passes: result is back in out.morton / out.prims
    BPT_CUDA_TRY(ctx, cudaMemsetAsync(flags.p, 0, (size_t)(n - 1) * 4, ctx->stream));
    LAUNCH(ctx, k_karras, grid_for(n - 1), kThreads, out.morton.as<uint64_t>(), n, c0.as<int32_t>(), c1.as<int32_t>(), np.as<int32_t>(), lp.as<int32_t>());
    LAUNCH(ctx, k_refit, grid_for(n), kThreads, n, out.prims.as<uint32_t>(), d_lo, d_hi, c0.as<int32_t>(), c1.as<int32_t>(), np.as<int32_t>(),
           lp.as<int32_t>(), nlo.as<float4>(), nhi.as<float4>(), flags.as<uint32_t>(), out.nodes.as<float4>());
    return BPT_OK;
}

bpt_status upload_instance_table(bpt_context* ctx) {
    uint32_t n = (uint32_t)ctx->h_instances.size();
    // the rule of k_make_instances, on the host copies (validated by bpt_build_accel): does ANY instance need the any-hit opacity rule?
    ctx->instanced_triangles = 0;
    for (const bpt_instance_desc& in : ctx->h_instances)
        if (in.blas < ctx->h_blas_desc.size()) ctx->instanced_triangles += ctx->h_blas_desc[in.blas].num_triangles;
    ctx->scene_has_anyhit = false;
    for (const bpt_instance_desc& in : ctx->h_instances) {
        const uint32_t flags = in.sbt_offset_and_flags >> 24, id = in.instance_id_and_mask & 0xffffffu;
        const uint32_t blend = (ctx->h_materials[ctx->h_drawables[id].material_offset / (uint32_t)sizeof(bpt_material)].flags >> BPT_MATERIAL_BLEND_SHIFT) & 0xffu;
        if ((flags & BPT_INSTANCE_FORCE_NON_OPAQUE) && blend != BPT_BLEND_OPAQUE) { ctx->scene_has_anyhit = true; break; }
    }
    Scratch sc(ctx);
    DevBuf desc;
    bpt_status s;
    if ((s = sc.get(ctx, desc, (size_t)n * sizeof(bpt_instance_desc)))) return s;
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(desc.p, ctx->h_instances.data(), (size_t)n * sizeof(bpt_instance_desc), cudaMemcpyHostToDevice, ctx->stream));
    if ((s = dev_reserve(ctx, ctx->d_instances, (size_t)n * sizeof(DInstance)))) return s;
    LAUNCH(ctx, k_make_instances, grid_for(n), kThreads, desc.as<bpt_instance_desc>(), n, ctx->d_drawables.as<bpt_drawable_sbt_data>(),
           ctx->d_materials.as<bpt_material>(), ctx->d_instances.as<DInstance>());
    return BPT_OK;      // stream-ordered: the builds that follow run on the same stream
}

static bpt_status emit_tris(bpt_context* ctx, DevBvh& b, const float4* raw) {
    bpt_status s;
    if ((s = dev_reserve(ctx, b.tris, (size_t)b.n * 48))) return s;
    LAUNCH(ctx, k_emit_tris, grid_for(b.n), kThreads, b.n, b.prims.as<uint32_t>(), raw, b.tris.as<float4>());
    return BPT_OK;      // `raw` is arena scratch: its re-use by a later build is stream-ordered
}

// 4-wide quantised nodes + exact leaf boxes of a built binary tree (any BLAS, or the TLAS with instance boxes as leaves)
static bpt_status collapse_wide(bpt_context* ctx, DevBvh& b) {
    if (b.n < 2) { dev_free(b.wide); dev_free(b.leafbox); return BPT_OK; }     // a single leaf is entered without any box test
    bpt_status s;
    if ((s = dev_reserve(ctx, b.wide, (size_t)(b.n - 1) * 64))) return s;
    if ((s = dev_reserve(ctx, b.leafbox, (size_t)b.n * 32))) return s;
    LAUNCH(ctx, k_collapse4, grid_for(b.n - 1), kThreads, b.n - 1, b.nodes.as<float4>(), b.wide.as<float4>(), b.leafbox.as<float4>());
    return BPT_OK;
}

bpt_status build_blas_two_level(bpt_context* ctx, uint32_t bi) {
    const bpt_blas_desc& bd = ctx->h_blas_desc[bi];
    uint32_t n = bd.num_triangles;
    Scratch sc(ctx); DevBuf raw, lo, hi; bpt_status s;
    if ((s = sc.get(ctx, raw, (size_t)n * 48))) return s;
    if ((s = sc.get(ctx, lo, (size_t)n * 16))) return s;
    if ((s = sc.get(ctx, hi, (size_t)n * 16))) return s;
    LAUNCH(ctx, k_gather_tris, grid_for(n), kThreads, ctx->d_positions.as<float>(), ctx->d_indices.as<uint32_t>(), (const GatherRange*)nullptr, GatherRange{bd, 0u}, 1u, n,
           (const DInstance*)nullptr, raw.as<float4>(), lo.as<float4>(), hi.as<float4>());
    if ((s = lbvh_build(ctx, ctx->blas[bi], n, lo.as<float4>(), hi.as<float4>(), ctx->d_blas_bounds.as<float>() + 6 * (size_t)bi))) return s;
    if ((s = emit_tris(ctx, ctx->blas[bi], raw.as<float4>()))) return s;
    return collapse_wide(ctx, ctx->blas[bi]);
}

// Every BLAS of the scene. The builds are independent, and a small one is a handful of launches that occupy one SM (k_lbvh_small), so they
// are issued round-robin on a few streams and overlap; their scratch is held until all of them are done (the arena's stack discipline
// assumes one stream) and released in one step, with a flush whenever more than 1 GB is held. The TLAS build that follows on the
// context's stream waits for all of them through events.
bpt_status build_all_blas_two_level(bpt_context* ctx) {
    const uint32_t nb = (uint32_t)ctx->blas.size();
    if (nb <= 2) {
        for (uint32_t b = 0; b < nb; b++) if (bpt_status s = build_blas_two_level(ctx, b)) return s;
        return BPT_OK;
    }
    constexpr int kStreams = 8;
    cudaStream_t main_stream = ctx->stream, st[kStreams];
    cudaEvent_t ready, done[kStreams];
    BPT_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    BPT_CUDA_TRY(ctx, cudaEventRecord(ready, main_stream));                      // the instance table / earlier uploads on the context's stream
    for (int k = 0; k < kStreams; k++) {
        BPT_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking));
        BPT_CUDA_TRY(ctx, cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
        BPT_CUDA_TRY(ctx, cudaStreamWaitEvent(st[k], ready, 0));
    }
    const size_t chunk0 = ctx->arena_chunk, offset0 = ctx->arena_offset;
    ctx->arena_hold = true; ctx->arena_held_bytes = 0;
    bpt_status s = BPT_OK;
    auto flush = [&]() {                                                        // wait for every stream, then the held scratch is free again
        for (int k = 0; k < kStreams; k++) cudaStreamSynchronize(st[k]);
        ctx->arena_chunk = chunk0; ctx->arena_offset = offset0; ctx->arena_held_bytes = 0;
    };
    for (uint32_t b = 0; b < nb && s == BPT_OK; b++) {
        if (ctx->arena_held_bytes > ((size_t)1 << 30)) flush();
        ctx->stream = st[b % kStreams];
        s = build_blas_two_level(ctx, b);
    }
    ctx->stream = main_stream;
    ctx->arena_hold = false;
    for (int k = 0; k < kStreams; k++) {                                        // what follows on the context's stream (TLAS build) is ordered after all builds
        cudaEventRecord(done[k], st[k]);
        cudaStreamWaitEvent(main_stream, done[k], 0);
    }
    if (s != BPT_OK) flush();                                                   // error path: nothing may still be using the scratch
    ctx->arena_chunk = chunk0; ctx->arena_offset = offset0;                     // (success: re-use is ordered by the events above)
    for (int k = 0; k < kStreams; k++) { cudaStreamDestroy(st[k]); cudaEventDestroy(done[k]); }
    cudaEventDestroy(ready);
    return s;
}

bpt_status build_blas_merged(bpt_context* ctx) {
    uint64_t total = 0;
    for (auto& in : ctx->h_instances) total += ctx->h_blas_desc[(uint32_t)in.blas].num_triangles;
    if (total == 0 || total > 0x7fffffffull) { ctx->err = "merged accel: triangle count out of range"; return BPT_ERR_INVALID; }
    uint32_t n = (uint32_t)total;
    {   // refuse up front when the merged soup cannot fit (build scratch + result ≈ 300 B per triangle)
        size_t free_b = 0, total_b = 0;
        BPT_CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        for (auto& a : ctx->arena_chunks) free_b += a.bytes;       // scratch of an earlier build is re-used
        const DevBvh& old = ctx->blas[0];
        free_b += old.nodes.bytes + old.tris.bytes + old.morton.bytes + old.prims.bytes + old.wide.bytes + old.leafbox.bytes;
        if ((double)total * 300.0 > (double)free_b) { ctx->err = "merged accel: not enough device memory for the flattened scene; use BPT_ACCEL_TWO_LEVEL"; return BPT_ERR_OOM; }
    }
    Scratch sc(ctx); DevBuf raw, lo, hi, rng; bpt_status s;
    if ((s = sc.get(ctx, raw, (size_t)n * 48))) return s;
    if ((s = sc.get(ctx, lo, (size_t)n * 16))) return s;
    if ((s = sc.get(ctx, hi, (size_t)n * 16))) return s;
    std::vector<GatherRange> ranges(ctx->h_instances.size());
    uint32_t base = 0;
    for (uint32_t slot = 0; slot < ctx->h_instances.size(); slot++) {
        ranges[slot] = GatherRange{ctx->h_blas_desc[(uint32_t)ctx->h_instances[slot].blas], base};
        base += ranges[slot].bd.num_triangles;
    }
    if ((s = sc.get(ctx, rng, ranges.size() * sizeof(GatherRange)))) return s;
    BPT_CUDA_TRY(ctx, cudaMemcpyAsync(rng.p, ranges.data(), ranges.size() * sizeof(GatherRange), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_gather_tris, grid_for(n), kThreads, ctx->d_positions.as<float>(), ctx->d_indices.as<uint32_t>(), rng.as<GatherRange>(), GatherRange{},
           (uint32_t)ranges.size(), n, ctx->d_instances.as<DInstance>(), raw.as<float4>(), lo.as<float4>(), hi.as<float4>());
    if ((s = lbvh_build(ctx, ctx->blas[0], n, lo.as<float4>(), hi.as<float4>(), ctx->d_blas_bounds.as<float>()))) return s;
    if ((s = emit_tris(ctx, ctx->blas[0], raw.as<float4>()))) return s;
    return collapse_wide(ctx, ctx->blas[0]);
}

bpt_status build_tlas(bpt_context* ctx) {
    uint32_t n = (uint32_t)ctx->h_instances.size();
    Scratch sc(ctx); DevBuf lo, hi, tb; bpt_status s;
    if ((s = sc.get(ctx, lo, (size_t)n * 16))) return s;
    if ((s = sc.get(ctx, hi, (size_t)n * 16))) return s;
    if ((s = sc.get(ctx, tb, 6 * sizeof(float)))) return s;
    LAUNCH(ctx, k_instance_bounds, grid_for(n), kThreads, ctx->d_instances.as<DInstance>(), n, ctx->d_blas_bounds.as<float>(), lo.as<float4>(), hi.as<float4>());
    if ((s = lbvh_build(ctx, ctx->tlas, n, lo.as<float4>(), hi.as<float4>(), tb.as<float>()))) return s;
    return collapse_wide(ctx, ctx->tlas);
}
