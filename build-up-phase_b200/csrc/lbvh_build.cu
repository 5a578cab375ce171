// lbvh_build.cu — LBVH acceleration-structure build for sm_100a: what the reference hands to
// vkCmdBuildAccelerationStructuresKHR (BLAS: main.cpp:748-808, TLAS: main.cpp:870-932).
//
// BLAS pipeline (one set of launches builds a whole BATCH of BLASes, segmented by key prefix):
//   k_tri_setup   fetch indices+vertices, bake the per-geometry 3x4 transform (transformData /
//                 transformOffset, main.cpp:737,799,803), write a 48-B triangle record, reduce the
//                 per-BLAS bounds (register -> warp shuffle -> one atomic set per warp)
//   k_tri_morton  30-bit Morton code of the triangle-box centre inside its BLAS bounds;
//                 key = (blas << 30) | morton, value = triangle id
//   sort_pairs    onesweep radix sort (radix_sort.cu), only the significant key bytes
//   k_karras      Karras 2012: one thread per internal node finds its range and split with clz on
//                 key XOR (index-augmented for duplicate keys). Because the BLAS id is the key
//                 prefix, every BLAS is exactly one subtree of the global radix tree.
//   k_refit_tris  one thread per sorted leaf: emits the sorted triangle, then climbs; each node has
//                 an arrival counter, the second thread to arrive unions the two child halves and
//                 continues (atomic bottom-up refit). Subtrees of <= 4 triangles collapse into a
//                 leaf. The thread that completes a BLAS's subtree publishes its root/bounds/height.
// TLAS: same machinery over instance world boxes (k_inst_setup computes world->object in fp64).
//
// All kernels are HBM-bound streaming passes: coalesced 128-bit accesses, grids sized from the
// problem (multiples of 148 SMs x resident CTAs for the big ones), no tensor cores (nothing here
// is a contraction).
#include <float.h>

#include "rt_device.cuh"

namespace rt {

namespace {

constexpr int SETUP_THREADS = 256;
constexpr int SETUP_ITEMS = 8;                       // triangles per thread
constexpr int SETUP_WARP_SPAN = 32 * SETUP_ITEMS;    // contiguous triangles per warp
constexpr int SETUP_CHUNK = SETUP_THREADS * SETUP_ITEMS;

__device__ __forceinline__ void atomic_bounds(int* b, const float* lo, const float* hi) {
    atomicMin(b + 0, float_to_ordered(lo[0])); atomicMin(b + 1, float_to_ordered(lo[1])); atomicMin(b + 2, float_to_ordered(lo[2]));
    atomicMax(b + 3, float_to_ordered(hi[0])); atomicMax(b + 4, float_to_ordered(hi[1])); atomicMax(b + 5, float_to_ordered(hi[2]));
}

// index of the geometry containing global triangle t (prefix has n_geoms+1 entries)
__device__ __forceinline__ uint32_t find_geom(const uint32_t* __restrict__ prefix, uint32_t n_geoms, uint32_t t) {
    uint32_t lo = 0, hi = n_geoms;      // invariant: prefix[lo] <= t < prefix[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= t) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(SETUP_THREADS) k_tri_setup(const GeomDesc* __restrict__ geoms, const uint32_t* __restrict__ prefix,
                                                            uint32_t n_geoms, uint32_t n_tris, TriRec* __restrict__ out,
                                                            int* __restrict__ bounds) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wfirst = blockIdx.x * (uint32_t)SETUP_CHUNK + warp * (uint32_t)SETUP_WARP_SPAN;
    if (wfirst >= n_tris) return;
    const uint32_t wlast = min(wfirst + (uint32_t)SETUP_WARP_SPAN, n_tris) - 1u;
    const uint32_t g_first = find_geom(prefix, n_geoms, wfirst);
    const uint32_t g_last = find_geom(prefix, n_geoms, wlast);
    const uint32_t blas_first = geoms[g_first].blas, blas_last = geoms[g_last].blas;
    const bool uniform = blas_first == blas_last;      // warp-uniform by construction

    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t g = g_first;
    uint32_t gbeg = __ldg(prefix + g), gend = __ldg(prefix + g + 1);
#pragma unroll 2
    for (int i = 0; i < SETUP_ITEMS; ++i) {
        const uint32_t t = wfirst + i * 32 + lane;
        if (t >= n_tris) break;
        if (t < gbeg || t >= gend) { g = find_geom(prefix, n_geoms, t); gbeg = __ldg(prefix + g); gend = __ldg(prefix + g + 1); }
        const GeomDesc& G = geoms[g];
        const uint32_t p = t - gbeg;
        uint32_t i0, i1, i2;
        if (G.idx) { i0 = __ldg(G.idx + 3 * (size_t)p); i1 = __ldg(G.idx + 3 * (size_t)p + 1); i2 = __ldg(G.idx + 3 * (size_t)p + 2); }
        else { i0 = 3 * p; i1 = 3 * p + 1; i2 = 3 * p + 2; }
        const float* a = G.verts + (size_t)i0 * G.stride_f;
        const float* b = G.verts + (size_t)i1 * G.stride_f;
        const float* c = G.verts + (size_t)i2 * G.stride_f;
        V3 v0 = {__ldg(a), __ldg(a + 1), __ldg(a + 2)};
        V3 v1 = {__ldg(b), __ldg(b + 1), __ldg(b + 2)};
        V3 v2 = {__ldg(c), __ldg(c + 1), __ldg(c + 2)};
        if (G.has_xform) { v0 = xform_point(G.xform, v0); v1 = xform_point(G.xform, v1); v2 = xform_point(G.xform, v2); }
        float4* dst = reinterpret_cast<float4*>(out + t);
        dst[0] = make_float4(v0.x, v0.y, v0.z, v1.x);
        dst[1] = make_float4(v1.y, v1.z, v2.x, v2.y);
        dst[2] = make_float4(v2.z, __uint_as_float(G.geo_index), __uint_as_float(p), __uint_as_float(G.blas));
        float tlo[3] = {fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z)};
        float thi[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z)};
        if (uniform) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], tlo[k]); hi[k] = fmaxf(hi[k], thi[k]); }
        } else {
            atomic_bounds(bounds + 6 * (size_t)G.blas, tlo, thi);
        }
    }
    if (uniform) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
                hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
            }
        }
        if (lane == 0) atomic_bounds(bounds + 6 * (size_t)blas_first, lo, hi);
    }
}

// vb > 0: packed record `key << vb | triangle id` (vals unused)
__global__ void __launch_bounds__(256) k_tri_morton(const TriRec* __restrict__ tris, uint32_t n_tris, const int* __restrict__ bounds,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int vb) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    const float4* src = reinterpret_cast<const float4*>(tris + t);
    const float4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
    const uint32_t blas = __float_as_uint(q2.w);
    float plo[3] = {fminf(fminf(q0.x, q0.w), q1.z), fminf(fminf(q0.y, q1.x), q1.w), fminf(fminf(q0.z, q1.y), q2.x)};
    float phi[3] = {fmaxf(fmaxf(q0.x, q0.w), q1.z), fmaxf(fmaxf(q0.y, q1.x), q1.w), fmaxf(fmaxf(q0.z, q1.y), q2.x)};
    const int* b = bounds + 6 * (size_t)blas;
    float slo[3] = {ordered_to_float(__ldg(b)), ordered_to_float(__ldg(b + 1)), ordered_to_float(__ldg(b + 2))};
    float shi[3] = {ordered_to_float(__ldg(b + 3)), ordered_to_float(__ldg(b + 4)), ordered_to_float(__ldg(b + 5))};
    const uint64_t key = ((uint64_t)blas << MORTON_BITS) | (uint64_t)morton30(plo, phi, slo, shi);
    if (vb) keys[t] = (key << vb) | (uint64_t)t;
    else { keys[t] = key; vals[t] = t; }
}

// ---- Karras 2012 ------------------------------------------------------------------------------
__device__ __forceinline__ int karras_delta(const uint64_t* __restrict__ keys, int vb, int n, int i, uint64_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t kj = __ldg(keys + j) >> vb;
    if (ki == kj) return 64 + __clz(i ^ j);
    return __clzll((long long)(ki ^ kj));
}

__global__ void __launch_bounds__(256) k_karras(const uint64_t* __restrict__ keys, int vb, int n, int32_t* __restrict__ other_end,
                                               uint32_t* __restrict__ parent_node, uint32_t* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const uint64_t ki = __ldg(keys + i) >> vb;
    int d = (karras_delta(keys, vb, n, i, ki, i + 1) - karras_delta(keys, vb, n, i, ki, i - 1)) >= 0 ? 1 : -1;
    if (i == 0) d = 1;
    const int dmin = karras_delta(keys, vb, n, i, ki, i - d);
    int lmax = 2;
    while (karras_delta(keys, vb, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (karras_delta(keys, vb, n, i, ki, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = karras_delta(keys, vb, n, i, ki, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (karras_delta(keys, vb, n, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    other_end[i] = j;
    if (first == gamma) parent_leaf[gamma] = ((uint32_t)i << 1);
    else parent_node[gamma] = ((uint32_t)i << 1);
    if (last == gamma + 1) parent_leaf[gamma + 1] = ((uint32_t)i << 1) | 1u;
    else parent_node[gamma + 1] = ((uint32_t)i << 1) | 1u;
}

// ---- atomic bottom-up refit ---------------------------------------------------------------------
struct Box3 { float lo[3], hi[3]; };

__device__ __forceinline__ void store_half(BvhNode* nodes, uint32_t node, uint32_t side, const Box3& b, int32_t ref, uint32_t height) {
    float4* dst = reinterpret_cast<float4*>(&nodes[node].c[side]);
    dst[0] = make_float4(b.lo[0], b.lo[1], b.lo[2], b.hi[0]);
    dst[1] = make_float4(b.hi[1], b.hi[2], __int_as_float(ref), __uint_as_float(height));
}

// Climbs from a leaf whose box/ref are given. seg_first/seg_count delimit the subtree (BLAS segment or the
// whole TLAS) whose root terminates the climb; refs are made relative to seg_first.
// Returns true when THIS thread completed the segment root (outputs valid).
template <int LEAF_MAX>
__device__ __forceinline__ bool climb(BvhNode* __restrict__ nodes, const uint32_t* __restrict__ parent_leaf,
                                      const uint32_t* __restrict__ parent_node, const int32_t* __restrict__ other_end,
                                      uint32_t* __restrict__ arrived, uint32_t leaf, uint32_t seg_first, uint32_t seg_count,
                                      Box3& b, int32_t& ref, uint32_t& height) {
    ref = leaf_ref(leaf - seg_first, 1);
    height = 0;
    if (seg_count == 1) return true;
    BvhNode* seg_nodes = nodes + seg_first;             // node slot of global node g is seg_nodes[g - seg_first]
    uint32_t p = parent_leaf[leaf];
    for (;;) {
        const uint32_t node = p >> 1, side = p & 1u;
        store_half(seg_nodes, node - seg_first, side, b, ref, height);
        __threadfence();
        if (atomicAdd(arrived + node, 1u) == 0u) return false;      // first arrival: the sibling's thread finishes this node
        const float4* sib = reinterpret_cast<const float4*>(&seg_nodes[node - seg_first].c[side ^ 1u]);
        const float4 s0 = __ldcg(sib), s1 = __ldcg(sib + 1);
        b.lo[0] = fminf(b.lo[0], s0.x); b.lo[1] = fminf(b.lo[1], s0.y); b.lo[2] = fminf(b.lo[2], s0.z);
        b.hi[0] = fmaxf(b.hi[0], s0.w); b.hi[1] = fmaxf(b.hi[1], s1.x); b.hi[2] = fmaxf(b.hi[2], s1.y);
        const uint32_t sib_height = __float_as_uint(s1.w);
        const int32_t j = other_end[node];
        const uint32_t first = min(node, (uint32_t)j), last = max(node, (uint32_t)j);
        const uint32_t count = last - first + 1u;
        if (count <= (uint32_t)LEAF_MAX) { ref = leaf_ref(first - seg_first, count); height = 0; }
        else { ref = (int32_t)(node - seg_first); height = max(height, sib_height) + 1u; }
        if (count == seg_count) return true;                        // completed the segment's root
        p = parent_node[node];
    }
}

__global__ void __launch_bounds__(256) k_refit_tris(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int vb, uint32_t n,
                                                   const TriRec* __restrict__ unsorted, TriRec* __restrict__ sorted,
                                                   BvhNode* __restrict__ nodes, BlasRecord* __restrict__ records,
                                                   const uint32_t* __restrict__ parent_leaf, const uint32_t* __restrict__ parent_node,
                                                   const int32_t* __restrict__ other_end, uint32_t* __restrict__ arrived) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    const uint64_t rec = __ldg(keys + leaf);
    const uint32_t blas = (uint32_t)(rec >> (MORTON_BITS + vb));
    const uint32_t src_i = vb ? (uint32_t)(rec & ((1ull << vb) - 1ull)) : __ldg(vals + leaf);
    const float4* src = reinterpret_cast<const float4*>(unsorted + src_i);
    const float4 q0 = __ldg(src), q1 = __ldg(src + 1);
    float4 q2 = __ldg(src + 2);
    q2.w = 0.0f;                                          // pad (carried the BLAS id through the sort)
    float4* dst = reinterpret_cast<float4*>(sorted + leaf);
    dst[0] = q0; dst[1] = q1; dst[2] = q2;
    Box3 b;
    b.lo[0] = fminf(fminf(q0.x, q0.w), q1.z); b.lo[1] = fminf(fminf(q0.y, q1.x), q1.w); b.lo[2] = fminf(fminf(q0.z, q1.y), q2.x);
    b.hi[0] = fmaxf(fmaxf(q0.x, q0.w), q1.z); b.hi[1] = fmaxf(fmaxf(q0.y, q1.x), q1.w); b.hi[2] = fmaxf(fmaxf(q0.z, q1.y), q2.x);
    const uint32_t seg_first = records[blas].first, seg_count = records[blas].tri_count;
    int32_t ref; uint32_t height;
    if (climb<BLAS_LEAF_MAX>(nodes, parent_leaf, parent_node, other_end, arrived, leaf, seg_first, seg_count, b, ref, height)) {
        BlasRecord& R = records[blas];
        R.root = ref; R.height = height;
        R.lo[0] = b.lo[0]; R.lo[1] = b.lo[1]; R.lo[2] = b.lo[2];
        R.hi[0] = b.hi[0]; R.hi[1] = b.hi[1]; R.hi[2] = b.hi[2];
    }
}

// ---- TLAS ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_inst_setup(const rt_instance* __restrict__ inst, uint32_t n, InstanceRec* __restrict__ out,
                                                   float* __restrict__ boxes, int* __restrict__ bounds, int32_t* __restrict__ meta) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int sbt_plus_geo = 0, sbt = 0, geo = 0, bheight = 0;
    if (i < n) {
        const float4* src = reinterpret_cast<const float4*>(inst + i);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
        float o2w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        const uint32_t custom_mask = __float_as_uint(d.x), sbt_flags = __float_as_uint(d.y);
        const unsigned long long addr = (unsigned long long)__float_as_uint(d.z) | ((unsigned long long)__float_as_uint(d.w) << 32);
        InstanceRec R;
        bool ok = invert3x4(o2w, R.w2o);
        R.nodes = nullptr; R.tris = nullptr; R.root = REF_EMPTY;
        R.custom_mask = custom_mask; R.sbt_flags = sbt_flags; R.instance_id = i;
        R.absmax[0] = R.absmax[1] = R.absmax[2] = 0.0f;
        if (addr != 0ull) {
            const BlasRecord* B = reinterpret_cast<const BlasRecord*>(addr);
            R.nodes = B->nodes; R.tris = B->tris; R.root = B->root;
            ok = ok && B->tri_count > 0 && B->root != REF_EMPTY;
            if (ok) {
#pragma unroll
                for (int k = 0; k < 3; ++k) R.absmax[k] = fmaxf(fabsf(B->lo[k]), fabsf(B->hi[k]));
                for (int cnr = 0; cnr < 8; ++cnr) {
                    V3 p = {(cnr & 1) ? B->hi[0] : B->lo[0], (cnr & 2) ? B->hi[1] : B->lo[1], (cnr & 4) ? B->hi[2] : B->lo[2]};
                    V3 w = xform_point(o2w, p);
                    lo[0] = fminf(lo[0], w.x); lo[1] = fminf(lo[1], w.y); lo[2] = fminf(lo[2], w.z);
                    hi[0] = fmaxf(hi[0], w.x); hi[1] = fmaxf(hi[1], w.y); hi[2] = fmaxf(hi[2], w.z);
                }
            }
            const int ng = B->n_geoms ? (int)B->n_geoms : 1;
            sbt = (int)(sbt_flags & 0xFFFFFFu); geo = ng - 1; sbt_plus_geo = sbt + geo; bheight = (int)B->height;
        } else ok = false;
        R.active = ok ? 1u : 0u;
        if (!ok) R.root = REF_EMPTY;
        out[i] = R;
        float* bx = boxes + 6 * (size_t)i;
        bx[0] = lo[0]; bx[1] = lo[1]; bx[2] = lo[2]; bx[3] = hi[0]; bx[4] = hi[1]; bx[5] = hi[2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        sbt_plus_geo = max(sbt_plus_geo, __shfl_xor_sync(0xffffffffu, sbt_plus_geo, o));
        sbt = max(sbt, __shfl_xor_sync(0xffffffffu, sbt, o));
        geo = max(geo, __shfl_xor_sync(0xffffffffu, geo, o));
        bheight = max(bheight, __shfl_xor_sync(0xffffffffu, bheight, o));
    }
    if (lane == 0) {
        if (lo[0] <= hi[0]) atomic_bounds(bounds, lo, hi);
        atomicMax(meta + 2, sbt_plus_geo); atomicMax(meta + 3, sbt); atomicMax(meta + 4, geo); atomicMax(meta + 5, bheight);
    }
}

__global__ void __launch_bounds__(256) k_inst_morton(const InstanceRec* __restrict__ inst, const float* __restrict__ boxes, uint32_t n,
                                                    const int* __restrict__ bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int vb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* bx = boxes + 6 * (size_t)i;
    float plo[3] = {bx[0], bx[1], bx[2]}, phi[3] = {bx[3], bx[4], bx[5]};
    float slo[3] = {ordered_to_float(bounds[0]), ordered_to_float(bounds[1]), ordered_to_float(bounds[2])};
    float shi[3] = {ordered_to_float(bounds[3]), ordered_to_float(bounds[4]), ordered_to_float(bounds[5])};
    const uint64_t key = inst[i].active ? (uint64_t)morton30(plo, phi, slo, shi) : 0x3FFFFFFFull;
    if (vb) keys[i] = (key << vb) | (uint64_t)i;
    else { keys[i] = key; vals[i] = i; }
}

__global__ void __launch_bounds__(256) k_refit_inst(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int vb, uint32_t n, const InstanceRec* __restrict__ unsorted,
                                                   const float* __restrict__ boxes, InstanceRec* __restrict__ sorted,
                                                   BvhNode* __restrict__ nodes, int32_t* __restrict__ meta, float* __restrict__ bounds_out,
                                                   const uint32_t* __restrict__ parent_leaf, const uint32_t* __restrict__ parent_node,
                                                   const int32_t* __restrict__ other_end, uint32_t* __restrict__ arrived) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    const uint32_t src_i = vb ? (uint32_t)(__ldg(keys + leaf) & ((1ull << vb) - 1ull)) : __ldg(vals + leaf);
    const float4* src = reinterpret_cast<const float4*>(unsorted + src_i);
    float4* dst = reinterpret_cast<float4*>(sorted + leaf);
#pragma unroll
    for (int k = 0; k < 6; ++k) dst[k] = __ldg(src + k);
    const float* bx = boxes + 6 * (size_t)src_i;
    Box3 b = {{bx[0], bx[1], bx[2]}, {bx[3], bx[4], bx[5]}};
    int32_t ref; uint32_t height;
    if (climb<TLAS_LEAF_MAX>(nodes, parent_leaf, parent_node, other_end, arrived, leaf, 0u, n, b, ref, height)) {
        meta[0] = ref; meta[1] = (int32_t)height;
        bounds_out[0] = b.lo[0]; bounds_out[1] = b.lo[1]; bounds_out[2] = b.lo[2];
        bounds_out[3] = b.hi[0]; bounds_out[4] = b.hi[1]; bounds_out[5] = b.hi[2];
    }
}

inline int div_up(uint32_t a, uint32_t b) { return (int)((a + b - 1) / b); }

}  // namespace

int launch_blas_build(const BlasBuildArgs& a, cudaStream_t st, const BuildEvents* ev, bool* sorted_in_b) {
    int launches = 0;
    *sorted_in_b = false;
    if (a.n_tris == 0) return 0;
    if (ev) cudaEventRecord(ev->e[0], st);
    k_tri_setup<<<div_up(a.n_tris, SETUP_CHUNK), SETUP_THREADS, 0, st>>>(a.geoms, a.geom_tri_first, a.n_geoms, a.n_tris, a.tris_unsorted, a.bounds_ordered);
    ++launches;
    if (ev) cudaEventRecord(ev->e[1], st);
    const int vb = a.sort.packed_val_bits;
    k_tri_morton<<<div_up(a.n_tris, 256), 256, 0, st>>>(a.tris_unsorted, a.n_tris, a.bounds_ordered, a.s.keys_a, a.s.vals_a, vb);
    ++launches;
    if (ev) cudaEventRecord(ev->e[2], st);
    int sl = sort_pairs(a.sort, a.s.keys_a, a.s.keys_b, vb ? nullptr : a.s.vals_a, vb ? nullptr : a.s.vals_b, a.s.sort_scratch, a.s.error_flag, st, sorted_in_b);
    if (sl < 0) return -1;
    launches += sl;
    const uint64_t* keys = *sorted_in_b ? a.s.keys_b : a.s.keys_a;
    const uint32_t* vals = *sorted_in_b ? a.s.vals_b : a.s.vals_a;
    if (ev) cudaEventRecord(ev->e[3], st);
    if (a.n_tris > 1) {
        if (cudaMemsetAsync(a.s.arrived, 0, sizeof(uint32_t) * (size_t)a.n_tris, st) != cudaSuccess) return -1;
        k_karras<<<div_up(a.n_tris - 1, 256), 256, 0, st>>>(keys, vb, (int)a.n_tris, a.s.other_end, a.s.parent_node, a.s.parent_leaf);
        ++launches;
    }
    if (ev) cudaEventRecord(ev->e[4], st);
    k_refit_tris<<<div_up(a.n_tris, 256), 256, 0, st>>>(keys, vals, vb, a.n_tris, a.tris_unsorted, a.tris_sorted, a.nodes, a.records,
                                                       a.s.parent_leaf, a.s.parent_node, a.s.other_end, a.s.arrived);
    ++launches;
    if (ev) cudaEventRecord(ev->e[5], st);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

int launch_tlas_build(const TlasBuildArgs& a, cudaStream_t st) {
    int launches = 0;
    if (a.n == 0) return 0;
    k_inst_setup<<<div_up(a.n, 128), 128, 0, st>>>(a.instances, a.n, a.inst_unsorted, a.boxes_unsorted, a.bounds_ordered, a.root_out);
    const int vb = a.sort.packed_val_bits;
    k_inst_morton<<<div_up(a.n, 256), 256, 0, st>>>(a.inst_unsorted, a.boxes_unsorted, a.n, a.bounds_ordered, a.s.keys_a, a.s.vals_a, vb);
    launches += 2;
    bool in_b = false;
    int sl = sort_pairs(a.sort, a.s.keys_a, a.s.keys_b, vb ? nullptr : a.s.vals_a, vb ? nullptr : a.s.vals_b, a.s.sort_scratch, a.s.error_flag, st, &in_b);
    if (sl < 0) return -1;
    launches += sl;
    const uint64_t* keys = in_b ? a.s.keys_b : a.s.keys_a;
    const uint32_t* vals = in_b ? a.s.vals_b : a.s.vals_a;
    if (a.n > 1) {
        if (cudaMemsetAsync(a.s.arrived, 0, sizeof(uint32_t) * (size_t)a.n, st) != cudaSuccess) return -1;
        k_karras<<<div_up(a.n - 1, 256), 256, 0, st>>>(keys, vb, (int)a.n, a.s.other_end, a.s.parent_node, a.s.parent_leaf);
        ++launches;
    }
    k_refit_inst<<<div_up(a.n, 256), 256, 0, st>>>(keys, vals, vb, a.n, a.inst_unsorted, a.boxes_unsorted, a.inst_sorted, a.nodes, a.root_out,
                                                  a.bounds_out, a.s.parent_leaf, a.s.parent_node, a.s.other_end, a.s.arrived);
    ++launches;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace rt
