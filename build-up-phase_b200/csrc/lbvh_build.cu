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
//   k_seg2_setup_sort  batches of small BLASes (each <= 11,264 triangles): the three steps above for ONE BLAS in one CTA
//                 (light shared-memory sort, two CTAs per SM, the vertices of a one-geometry BLAS staged in shared memory)
//   k_refit_tris  hierarchy emission AND refit in one bottom-up pass over Karras' radix tree (clz on
//                 key XOR, index-augmented for duplicate keys; found bottom-up as in Apetrei 2014, see
//                 build_tree_tile): one thread per sorted leaf emits the sorted triangle, then climbs;
//                 the second child to arrive at a split unions the boxes and continues. Splits inside
//                 a CTA's 256-leaf tile meet in shared memory (after two merges the subtrees still climbing
//                 are regrouped into the first warps). Because the BLAS id is the key prefix, every BLAS of
//                 a batch is exactly one subtree of the global radix tree; subtrees of <= 2 triangles collapse
//                 into a leaf; the thread that completes a BLAS publishes its root/bounds/height.
//   k_tree_border the subtrees that touch a tile border finish through global memory: a job queue walked by
//                 strided lanes, one acquire-add per arrival, deposits that carry the deltas of their range.
// TLAS: same machinery over instance world boxes (k_inst_setup computes world->object in fp64).
//
// The setup / sort kernels stream their data with coalesced 128-bit accesses; the tree kernels are latency-chain
// bound (profiles/README.md r2_x - r2_zk). Grids are sized from the problem or from the resident CTAs of the
// device; no tensor cores (nothing here is a contraction).
#include <float.h>
#include <stdlib.h>

#include "rt_device.cuh"
#include "seg_sort.cuh"

namespace rt {

namespace {

constexpr int SETUP_THREADS = 256;
#ifndef RT_SETUP_BATCH
#define RT_SETUP_BATCH 1
#endif
constexpr int SETUP_ITEMS = 8;                       // triangles per thread
constexpr int SETUP_BATCH = RT_SETUP_BATCH;           // of which this many are in flight together
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
    // Batches of SETUP_BATCH triangles per lane: all index loads of a batch are issued before its first vertex load, all
    // vertex loads before the first use. Measured on B200 (inst10m, profiles/README.md r01q): 0.201 / 0.226 / 0.270 / 0.329 ms
    // for batches of 1 / 2 / 4 / 8 — the kernel is bound by LSU transactions of the 4-byte gathers (lg_throttle), not by
    // their latency, so wider batches only cost occupancy. Default 1.
#pragma unroll 1
    for (int i0 = 0; i0 < SETUP_ITEMS; i0 += SETUP_BATCH) {
        uint32_t gi[SETUP_BATCH], pi[SETUP_BATCH], ix[SETUP_BATCH][3];
        bool ok[SETUP_BATCH];
#pragma unroll
        for (int k = 0; k < SETUP_BATCH; ++k) {
            const uint32_t t = wfirst + (i0 + k) * 32 + lane;
            ok[k] = t < n_tris;
            if (ok[k] && (t < gbeg || t >= gend)) { g = find_geom(prefix, n_geoms, t); gbeg = __ldg(prefix + g); gend = __ldg(prefix + g + 1); }
            gi[k] = g; pi[k] = t - gbeg;
            const uint32_t* idx = geoms[g].idx;
            if (ok[k] && idx) { ix[k][0] = __ldg(idx + 3 * (size_t)pi[k]); ix[k][1] = __ldg(idx + 3 * (size_t)pi[k] + 1); ix[k][2] = __ldg(idx + 3 * (size_t)pi[k] + 2); }
            else { ix[k][0] = 3 * pi[k]; ix[k][1] = 3 * pi[k] + 1; ix[k][2] = 3 * pi[k] + 2; }
        }
        float vx[SETUP_BATCH][9];
#pragma unroll
        for (int k = 0; k < SETUP_BATCH; ++k) {
            if (!ok[k]) continue;
            const GeomDesc& G = geoms[gi[k]];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* a = G.verts + (size_t)ix[k][c] * G.stride_f;
                vx[k][3 * c] = __ldg(a); vx[k][3 * c + 1] = __ldg(a + 1); vx[k][3 * c + 2] = __ldg(a + 2);
            }
        }
#pragma unroll
        for (int k = 0; k < SETUP_BATCH; ++k) {
            if (!ok[k]) continue;
            const uint32_t t = wfirst + (i0 + k) * 32 + lane;
            const GeomDesc& G = geoms[gi[k]];
            V3 v0 = {vx[k][0], vx[k][1], vx[k][2]}, v1 = {vx[k][3], vx[k][4], vx[k][5]}, v2 = {vx[k][6], vx[k][7], vx[k][8]};
            if (G.has_xform) { v0 = xform_point(G.xform, v0); v1 = xform_point(G.xform, v1); v2 = xform_point(G.xform, v2); }
            float4* dst = reinterpret_cast<float4*>(out + t);
            dst[0] = make_float4(v0.x, v0.y, v0.z, v1.x);
            dst[1] = make_float4(v1.y, v1.z, v2.x, v2.y);
            dst[2] = make_float4(v2.z, __uint_as_float(G.geo_index), __uint_as_float(pi[k]), __uint_as_float(G.blas | (G.flags << 24)));
            float tlo[3] = {fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z)};
            float thi[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z)};
            if (uniform) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { lo[c] = fminf(lo[c], tlo[c]); hi[c] = fmaxf(hi[c], thi[c]); }
            } else {
                atomic_bounds(bounds + 6 * (size_t)G.blas, tlo, thi);
            }
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
    const uint32_t blas = __float_as_uint(q2.w) & 0xFFFFFFu;
    float plo[3] = {fminf(fminf(q0.x, q0.w), q1.z), fminf(fminf(q0.y, q1.x), q1.w), fminf(fminf(q0.z, q1.y), q2.x)};
    float phi[3] = {fmaxf(fmaxf(q0.x, q0.w), q1.z), fmaxf(fmaxf(q0.y, q1.x), q1.w), fmaxf(fmaxf(q0.z, q1.y), q2.x)};
    const int* b = bounds + 6 * (size_t)blas;
    float slo[3] = {ordered_to_float(__ldg(b)), ordered_to_float(__ldg(b + 1)), ordered_to_float(__ldg(b + 2))};
    float shi[3] = {ordered_to_float(__ldg(b + 3)), ordered_to_float(__ldg(b + 4)), ordered_to_float(__ldg(b + 5))};
    const uint64_t key = ((uint64_t)blas << MORTON_BITS) | (uint64_t)morton30(plo, phi, slo, shi);
    if (vb) keys[t] = (key << vb) | (uint64_t)t;
    else { keys[t] = key; vals[t] = t; }
}

#ifndef RT_SEG_EMIT_SORTED
#define RT_SEG_EMIT_SORTED 0     // 1: the fused per-BLAS kernel also writes the SORTED triangle records and k_refit_tris<true> only streams them. Measured
                                 // (r2_zg): the tile kernel gets 95 us faster (534 -> 439), the one-CTA-per-SM segment kernel 160 us slower (456 -> 615): off
#endif
#ifndef RT_SEG_EMIT_BATCH
#define RT_SEG_EMIT_BATCH 3
#endif
// ---- batches of small BLASes: setup + Morton + sort of one whole BLAS in ONE CTA ---------------------------------------
// When every BLAS of the batch fits one CTA's shared memory (SEG_SORT_CAPACITY triangles), CTA b does for BLAS b what
// k_tri_setup, k_tri_morton and the sort do for the general case: fetch + bake its triangles (48-B records out), reduce
// ITS bounds in the block (no global atomics), compute the Morton keys (each thread re-reads its own records: cache hits),
// and sort the packed records in shared memory (seg_sort.cuh). One launch, the records
// written once, the keys written once, already sorted.
__global__ void __launch_bounds__(SEG_THREADS, 1) k_seg_setup_sort(const GeomDesc* __restrict__ geoms, const uint32_t* __restrict__ prefix, uint32_t n_geoms,
                                                                  const BlasRecord* __restrict__ recs, TriRec* __restrict__ out,
                                                                  uint64_t* __restrict__ keys_out, int vb, TriRec* __restrict__ sorted_out) {
    extern __shared__ __align__(16) unsigned char seg_smem[];
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(seg_smem);
    __shared__ float s_red[6][SEG_WARPS];
    __shared__ float s_bounds[6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t blas = blockIdx.x, first = recs[blas].first, n = recs[blas].tri_count;
    if (n == 0) return;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t g = 0, gbeg = 1, gend = 0;                  // empty range: the first triangle looks its geometry up
#pragma unroll 1
    for (int i = 0; i < SEG_ITEMS; ++i) {
        const uint32_t t = (uint32_t)tid + (uint32_t)i * SEG_THREADS;
        if (t >= n) continue;
        const uint32_t T = first + t;
        if (T < gbeg || T >= gend) { g = find_geom(prefix, n_geoms, T); gbeg = __ldg(prefix + g); gend = __ldg(prefix + g + 1); }
        const GeomDesc& G = geoms[g];
        const uint32_t p = T - gbeg;
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
        float4* dst = reinterpret_cast<float4*>(out + T);
        dst[0] = make_float4(v0.x, v0.y, v0.z, v1.x);
        dst[1] = make_float4(v1.y, v1.z, v2.x, v2.y);
        dst[2] = make_float4(v2.z, __uint_as_float(G.geo_index), __uint_as_float(p), __uint_as_float(G.blas | (G.flags << 24)));
        const float tlo[3] = {fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z)};
        const float thi[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], tlo[k]); hi[k] = fmaxf(hi[k], thi[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if (lane == 0) { s_red[k][warp] = lo[k]; s_red[3 + k][warp] = hi[k]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float v = s_red[k][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const float w = __shfl_xor_sync(0xffffffffu, v, o); v = k < 3 ? fminf(v, w) : fmaxf(v, w); }
            if (lane == 0) s_bounds[k] = v;
        }
    }
    __syncthreads();
    const float slo[3] = {s_bounds[0], s_bounds[1], s_bounds[2]}, shi[3] = {s_bounds[3], s_bounds[4], s_bounds[5]};
#pragma unroll
    for (int i = 0; i < SEG_ITEMS; ++i) {
        const uint32_t t = (uint32_t)tid + (uint32_t)i * SEG_THREADS;
        uint64_t rec = ~0ull;                              // padding sorts last and stays last
        if (t < n) {
            // this thread's own record, written a moment ago (L1/L2 hit): cheaper than keeping 33 centre floats live across the reduction
            const float4* src = reinterpret_cast<const float4*>(out + first + t);
            const float4 q0 = src[0], q1 = src[1], q2 = src[2];
            const float plo[3] = {fminf(fminf(q0.x, q0.w), q1.z), fminf(fminf(q0.y, q1.x), q1.w), fminf(fminf(q0.z, q1.y), q2.x)};
            const float phi[3] = {fmaxf(fmaxf(q0.x, q0.w), q1.z), fmaxf(fmaxf(q0.y, q1.x), q1.w), fmaxf(fmaxf(q0.z, q1.y), q2.x)};
            rec = ((((uint64_t)blas << MORTON_BITS) | (uint64_t)morton30(plo, phi, slo, shi)) << vb) | (uint64_t)(first + t);
        }
        s_keys[t] = rec;
    }
    __syncthreads();
    seg_sort_passes(seg_smem, vb, (int)MORTON_BITS, n);
#if RT_SEG_EMIT_SORTED
    // The sorted triangle records too (sorted_out != nullptr): this CTA wrote the unsorted ones a moment ago, so the gather by sorted id is
    // served by L2 here, and the tree pass reads its leaves as one coalesced stream instead of a dependent key -> record gather
    // (34 % of k_refit_tris' stall samples, profiles/README.md r2_y). Three records in flight per thread.
    if (sorted_out) {
        const uint64_t idmask = (1ull << vb) - 1ull;
        constexpr int EB = RT_SEG_EMIT_BATCH;
        for (uint32_t p0 = (uint32_t)tid; p0 < n; p0 += EB * SEG_THREADS) {
            float4 q[EB][3];
#pragma unroll
            for (int k = 0; k < EB; ++k) {
                const uint32_t p = p0 + (uint32_t)k * SEG_THREADS;
                if (p < n) {
                    const float4* src = reinterpret_cast<const float4*>(out + (uint32_t)(s_keys[p] & idmask));
                    q[k][0] = src[0]; q[k][1] = src[1]; q[k][2] = src[2];
                }
            }
#pragma unroll
            for (int k = 0; k < EB; ++k) {
                const uint32_t p = p0 + (uint32_t)k * SEG_THREADS;
                if (p < n) {
                    float4* dst = reinterpret_cast<float4*>(sorted_out + first + p);
                    q[k][2].w = __uint_as_float(__float_as_uint(q[k][2].w) >> 24);   // geometry flags (the low 24 bits carried the BLAS id)
                    dst[0] = q[k][0]; dst[1] = q[k][1]; dst[2] = q[k][2];
                }
            }
        }
    }
#endif
    seg_store_bulk(keys_out + first, s_keys, n);          // the sorted segment leaves shared memory as one TMA bulk copy (UBLKCP)
}

// The same with the LIGHT segment sort (seg_sort.cuh): 512 threads x 22 triangles, {u32 Morton, u16 position} records, two CTAs per SM.
// Bit-identical output (same stable order); RT_SEG_LIGHT=0 selects the 1024-thread kernel above.
#ifndef RT_SEG_LIGHT
#define RT_SEG_LIGHT 1
#endif
#ifndef RT_SEG_STAGE_VERTS
#define RT_SEG_STAGE_VERTS 1
#endif
#ifndef RT_SEG_MORTON_BATCH
#define RT_SEG_MORTON_BATCH 4
#endif
__global__ void __launch_bounds__(SEG2_THREADS, 2) k_seg2_setup_sort(const GeomDesc* __restrict__ geoms, const uint32_t* __restrict__ prefix, uint32_t n_geoms,
                                                                    const BlasRecord* __restrict__ recs, TriRec* __restrict__ out,
                                                                    uint64_t* __restrict__ keys_out, int vb, TriRec* __restrict__ sorted_out) {
    extern __shared__ __align__(16) unsigned char seg_smem[];
    uint32_t* s_m = reinterpret_cast<uint32_t*>(seg_smem + SEG2_OFF_M);
    uint16_t* s_id = reinterpret_cast<uint16_t*>(seg_smem + SEG2_OFF_ID);
    __shared__ float s_red[6][SEG2_WARPS];
    __shared__ float s_bounds[6];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t blas = blockIdx.x, first = recs[blas].first, n = recs[blas].tri_count;
    if (n == 0) return;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t g = 0, gbeg = 1, gend = 0;                  // empty range: the first triangle looks its geometry up
#if RT_SEG_STAGE_VERTS
    // A BLAS that is ONE indexed geometry whose vertex array fits the (not yet used) sort arrays: copy the vertices into shared memory with
    // coalesced loads and gather from there. The gather of nine 4-byte values per triangle is what bounds the setup phase (L1TEX wavefronts:
    // every lane its own line), and neighbouring triangles share their vertices.
    const float* s_verts = reinterpret_cast<const float*>(seg_smem);
    bool staged = false;
    {
        const uint32_t g0 = find_geom(prefix, n_geoms, first);
        const GeomDesc& G0 = geoms[g0];
        const uint64_t words = (uint64_t)G0.vert_count * G0.stride_f;
        if (G0.idx && __ldg(prefix + g0) == first && __ldg(prefix + g0 + 1) >= first + n && words != 0 && words * 4ull <= (uint64_t)SEG2_SMEM_BYTES) {
            staged = true;                               // uniform for the CTA
            float* dstv = reinterpret_cast<float*>(seg_smem);
            for (uint32_t w = (uint32_t)tid; w < (uint32_t)words; w += SEG2_THREADS) dstv[w] = __ldg(G0.verts + w);
            g = g0; gbeg = first; gend = __ldg(prefix + g0 + 1);
        }
    }
    __syncthreads();
#else
    const float* s_verts = nullptr;
    const bool staged = false;
#endif
    if (staged) {
        // one indexed geometry, vertices in shared memory: the only global loads left are the three indices of a triangle, fetched TWO
        // iterations ahead (a thread's 22 triangles were 22 dependent index round trips: 10 % of the kernel's stall samples, r2_zj)
        const GeomDesc& G = geoms[g];
        const uint32_t* __restrict__ idx = G.idx;
        const uint32_t stride_f = G.stride_f, tagw = G.blas | (G.flags << 24), geo = G.geo_index;
        const bool xf = G.has_xform != 0;
        constexpr int AHEAD = 2;
        uint32_t pi[AHEAD][3];
#pragma unroll
        for (int k = 0; k < AHEAD; ++k) {
            const uint32_t t = (uint32_t)tid + (uint32_t)k * SEG2_THREADS;
            if (t < n) { pi[k][0] = __ldg(idx + 3 * (size_t)t); pi[k][1] = __ldg(idx + 3 * (size_t)t + 1); pi[k][2] = __ldg(idx + 3 * (size_t)t + 2); }
        }
#pragma unroll
        for (int i = 0; i < SEG2_ITEMS; ++i) {
            const uint32_t t = (uint32_t)tid + (uint32_t)i * SEG2_THREADS;
            const uint32_t i0 = pi[i % AHEAD][0], i1 = pi[i % AHEAD][1], i2 = pi[i % AHEAD][2];
            if (i + AHEAD < SEG2_ITEMS) {
                const uint32_t tn = t + (uint32_t)AHEAD * SEG2_THREADS;
                if (tn < n) { pi[i % AHEAD][0] = __ldg(idx + 3 * (size_t)tn); pi[i % AHEAD][1] = __ldg(idx + 3 * (size_t)tn + 1); pi[i % AHEAD][2] = __ldg(idx + 3 * (size_t)tn + 2); }
            }
            if (t < n) {
                const float* a = s_verts + i0 * stride_f;
                const float* b = s_verts + i1 * stride_f;
                const float* c = s_verts + i2 * stride_f;
                V3 v0 = {a[0], a[1], a[2]}, v1 = {b[0], b[1], b[2]}, v2 = {c[0], c[1], c[2]};
                if (xf) { v0 = xform_point(G.xform, v0); v1 = xform_point(G.xform, v1); v2 = xform_point(G.xform, v2); }
                float4* dst = reinterpret_cast<float4*>(out + first + t);
                dst[0] = make_float4(v0.x, v0.y, v0.z, v1.x);
                dst[1] = make_float4(v1.y, v1.z, v2.x, v2.y);
                dst[2] = make_float4(v2.z, __uint_as_float(geo), __uint_as_float(t), __uint_as_float(tagw));
                lo[0] = fminf(lo[0], fminf(fminf(v0.x, v1.x), v2.x)); lo[1] = fminf(lo[1], fminf(fminf(v0.y, v1.y), v2.y)); lo[2] = fminf(lo[2], fminf(fminf(v0.z, v1.z), v2.z));
                hi[0] = fmaxf(hi[0], fmaxf(fmaxf(v0.x, v1.x), v2.x)); hi[1] = fmaxf(hi[1], fmaxf(fmaxf(v0.y, v1.y), v2.y)); hi[2] = fmaxf(hi[2], fmaxf(fmaxf(v0.z, v1.z), v2.z));
            }
        }
    } else {
#pragma unroll 1
    for (int i = 0; i < SEG2_ITEMS; ++i) {
        const uint32_t t = (uint32_t)tid + (uint32_t)i * SEG2_THREADS;
        if (t >= n) continue;
        const uint32_t T = first + t;
        if (T < gbeg || T >= gend) { g = find_geom(prefix, n_geoms, T); gbeg = __ldg(prefix + g); gend = __ldg(prefix + g + 1); }
        const GeomDesc& G = geoms[g];
        const uint32_t p = T - gbeg;
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
        float4* dst = reinterpret_cast<float4*>(out + T);
        dst[0] = make_float4(v0.x, v0.y, v0.z, v1.x);
        dst[1] = make_float4(v1.y, v1.z, v2.x, v2.y);
        dst[2] = make_float4(v2.z, __uint_as_float(G.geo_index), __uint_as_float(p), __uint_as_float(G.blas | (G.flags << 24)));
        const float tlo[3] = {fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z)};
        const float thi[3] = {fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], tlo[k]); hi[k] = fmaxf(hi[k], thi[k]); }
    }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if (lane == 0) { s_red[k][warp] = lo[k]; s_red[3 + k][warp] = hi[k]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float v = lane < SEG2_WARPS ? s_red[k][lane] : (k < 3 ? FLT_MAX : -FLT_MAX);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const float w = __shfl_xor_sync(0xffffffffu, v, o); v = k < 3 ? fminf(v, w) : fmaxf(v, w); }
            if (lane == 0) s_bounds[k] = v;
        }
    }
    __syncthreads();
    const float slo[3] = {s_bounds[0], s_bounds[1], s_bounds[2]}, shi[3] = {s_bounds[3], s_bounds[4], s_bounds[5]};
    // Morton keys from this thread's own records, written a moment ago (L2 hits, ~1300 cycles under load): RT_SEG_MORTON_BATCH records' loads are
    // in flight together (two at a time were 11 dependent round trips: 14 % of the kernel's stall samples, r2_zj)
    constexpr int MB = RT_SEG_MORTON_BATCH;
#pragma unroll 1
    for (int i0 = 0; i0 < SEG2_ITEMS; i0 += MB) {
        float4 q[MB][3];
#pragma unroll
        for (int k = 0; k < MB; ++k) {
            const uint32_t t = (uint32_t)tid + (uint32_t)(i0 + k) * SEG2_THREADS;
            if (i0 + k < SEG2_ITEMS && t < n) {
                const float4* src = reinterpret_cast<const float4*>(out + first + t);
                q[k][0] = src[0]; q[k][1] = src[1]; q[k][2] = src[2];
            }
        }
#pragma unroll
        for (int k = 0; k < MB; ++k) {
            const uint32_t t = (uint32_t)tid + (uint32_t)(i0 + k) * SEG2_THREADS;
            if (i0 + k >= SEG2_ITEMS) continue;
            uint32_t mk = 0xFFFFFFFFu;                         // padding sorts last and stays last
            if (t < n) {
                const float4 q0 = q[k][0], q1 = q[k][1], q2 = q[k][2];
                const float plo[3] = {fminf(fminf(q0.x, q0.w), q1.z), fminf(fminf(q0.y, q1.x), q1.w), fminf(fminf(q0.z, q1.y), q2.x)};
                const float phi[3] = {fmaxf(fmaxf(q0.x, q0.w), q1.z), fmaxf(fmaxf(q0.y, q1.x), q1.w), fmaxf(fmaxf(q0.z, q1.y), q2.x)};
                mk = morton30(plo, phi, slo, shi);
            }
            s_m[seg2_m_at(t)] = mk;
            s_id[t] = (uint16_t)t;
        }
    }
    __syncthreads();
    seg2_sort_passes(seg_smem, n);
    const uint64_t hi_bits = (uint64_t)blas << MORTON_BITS;
    for (uint32_t p = (uint32_t)tid; p < n; p += SEG2_THREADS)
        keys_out[first + p] = ((hi_bits | (uint64_t)s_m[seg2_m_at(p)]) << vb) | (uint64_t)(first + (uint32_t)s_id[p]);
#if RT_SEG_EMIT_SORTED
    // the sorted triangle records too (see k_seg_setup_sort): gathered from L2 by the CTA that wrote them; k_refit_tris<true> then streams its leaves
    if (sorted_out) {
        constexpr int EB = RT_SEG_EMIT_BATCH;
        for (uint32_t p0 = (uint32_t)tid; p0 < n; p0 += EB * SEG2_THREADS) {
            float4 q[EB][3];
#pragma unroll
            for (int k = 0; k < EB; ++k) {
                const uint32_t p = p0 + (uint32_t)k * SEG2_THREADS;
                if (p < n) {
                    const float4* src = reinterpret_cast<const float4*>(out + first + (uint32_t)s_id[p]);
                    q[k][0] = src[0]; q[k][1] = src[1]; q[k][2] = src[2];
                }
            }
#pragma unroll
            for (int k = 0; k < EB; ++k) {
                const uint32_t p = p0 + (uint32_t)k * SEG2_THREADS;
                if (p < n) {
                    float4* dst = reinterpret_cast<float4*>(sorted_out + first + p);
                    q[k][2].w = __uint_as_float(__float_as_uint(q[k][2].w) >> 24);   // geometry flags (the low 24 bits carried the BLAS id)
                    dst[0] = q[k][0]; dst[1] = q[k][1]; dst[2] = q[k][2];
                }
            }
        }
    }
#else
    (void)sorted_out;
#endif
}

// ---- hierarchy emission + refit, one bottom-up pass ---------------------------------------------------
// The tree is Karras' binary radix tree over the (key, index) strings, but it is found BOTTOM-UP (Apetrei 2014):
// a finished subtree over sorted leaves [l, r] merges with its right neighbour when it shares the longer prefix
// with it (delta(r) > delta(l-1)), else with the left one; the two meet at the SPLIT position g = r resp. l-1, which is
// unique per node. The first child to arrive at a split deposits {box, ref, height, far end of its range} and retires;
// the second one unions and climbs on. No top-down Karras pass, no parent arrays.
// Node numbering is Karras' own: the left child of split g is stored in slot g, the right child in slot g + 1 (so a
// node's slot is the right end of its range when it will merge to the right, the left end otherwise; a root takes the
// first slot of its segment). Siblings therefore share one 128-byte line, which the traversal likes (measured: +1.3 %
// Mrays/s over numbering nodes by their own split position).
//
// A CTA owns TILE consecutive leaves. Every split whose two leaves g, g+1 lie inside the tile is resolved through
// SHARED memory (deposit slots + an atomicOr flag, CTA-scope fences only); live nodes are then written once, as a
// full 64-B record, by the thread that completes them, and the halves of subtrees that collapse into a <= LEAF_MAX
// leaf are never written. What is left after the CTA quiesces (subtrees that touch the tile border: ~log2(TILE)/TILE
// of the nodes) continues through global memory with the classic fence + arrival counter.
struct Box3 { float lo[3], hi[3]; };

#ifndef RT_TREE_TILE
#define RT_TREE_TILE 256
#endif
#ifndef RT_TREE_MIN_CTAS
#define RT_TREE_MIN_CTAS (1536 / RT_TREE_TILE)
#endif
constexpr int TREE_TILE = RT_TREE_TILE;            // <= 512: a tile-local range end is packed into 9 bits next to the height
static_assert(TREE_TILE <= 512 && (TREE_TILE & (TREE_TILE - 1)) == 0, "tile");

// CTA-scope acquire/release fence (cheaper than the sequentially consistent one __threadfence_block() emits)
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
// GPU-scope acquire/release fence: what the deposit -> arrival counter -> sibling read handshake needs; __threadfence() emits the
// sequentially consistent one (MEMBAR.SC.GPU), which measured 36 % of the border kernel's stall samples
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// common-prefix length of the augmented strings (key_i, i) and (key_{i+1}, i+1)
__device__ __forceinline__ int delta_adjacent(uint64_t ka, uint64_t kb, uint32_t i) {
    return ka == kb ? 64 + __clz((int)(i ^ (i + 1u))) : __clzll((long long)(ka ^ kb));
}

struct TreeJob {
    uint32_t l, r;              // sorted-leaf range of the finished subtree
    Box3 b;
    int32_t ref;                // how the parent refers to this subtree (relative to the segment)
    uint32_t height;
    uint32_t seg_first, seg_count;
    int dl, dr;                 // delta(l - 1), delta(r): only maintained on the border path (in a tile they are read from shared memory)
};

// Segment policies: seg_of(key) gives first/count of the segment (one BLAS of a batch, or the whole TLAS) a key belongs
// to; on_root(job) is called by the one thread that completes a segment's root.
struct TriSegments {
    BlasRecord* records; const uint64_t* keys; int vb;
    __device__ __forceinline__ void seg_of(uint64_t key, uint32_t& first, uint32_t& count) const {
        const BlasRecord& R = records[(uint32_t)(key >> MORTON_BITS)];
        first = R.first; count = R.tri_count;
    }
    __device__ __forceinline__ void on_root(const TreeJob& j) const {
        BlasRecord& R = records[(uint32_t)((__ldg(keys + j.l) >> vb) >> MORTON_BITS)];
        R.root = j.ref; R.height = j.height;
        R.lo[0] = j.b.lo[0]; R.lo[1] = j.b.lo[1]; R.lo[2] = j.b.lo[2];
        R.hi[0] = j.b.hi[0]; R.hi[1] = j.b.hi[1]; R.hi[2] = j.b.hi[2];
    }
};
struct InstSegment {
    uint32_t n; int32_t* meta; float* bounds_out;
    __device__ __forceinline__ void seg_of(uint64_t, uint32_t& first, uint32_t& count) const { first = 0u; count = n; }
    __device__ __forceinline__ void on_root(const TreeJob& j) const {
        meta[0] = j.ref; meta[1] = (int32_t)j.height;
        bounds_out[0] = j.b.lo[0]; bounds_out[1] = j.b.lo[1]; bounds_out[2] = j.b.lo[2];
        bounds_out[3] = j.b.hi[0]; bounds_out[4] = j.b.hi[1]; bounds_out[5] = j.b.hi[2];
    }
};

template <int LEAF_MAX, class Seg>
__device__ __forceinline__ void climb_global(TreeJob j, BvhNode* __restrict__ nodes,
                                             float4* __restrict__ xchg, uint32_t* __restrict__ far_end, uint32_t* __restrict__ arrived, const Seg& seg);
// RT_TREE_INLINE_BORDER=1: the threads that hold a tile's unfinished subtrees / orphans continue through global memory themselves, right
// after the tile quiesces, instead of handing them to a second kernel through the job queue (no k_tree_border launch).
#ifndef RT_TREE_INLINE_BORDER
#define RT_TREE_INLINE_BORDER 0
#endif
struct BorderMem { float4* xchg; uint32_t* far_end; uint32_t* arrived; };

// Phase 1 (inside the leaf kernels): everything a tile can finish on its own. Unfinished subtrees (at most 2 per tile:
// the ones whose next split lies outside it) and orphans (a child deposited in shared memory whose sibling straddles
// the tile border: at most one per straddling ancestor of the two border leaves) are appended to the border-job queue
// as 48-B records {lo.xyz hi.x | hi.yz ref height+deltas | l r seg_first seg_count}; one global atomic per CTA. A job carries its
// segment and the deltas at both ends of its range, so the border kernel never loads a key or a segment record.
//
// RT_TREE_REGROUP = R > 0: a thread climbs for at most R merges, then the CTA meets once and the subtrees that are still
// climbing are handed to the first threads of the CTA, which finish the tile. Every merge adds at least one leaf, so a survivor
// holds > R leaves and at most TILE / (R + 1) of them exist (three warps of eight for R = 2 and 256-leaf tiles): the long tail of the
// climb, where 1-3 lanes of EVERY warp keep their warp issuing, runs in the first warps only. Same deposits, same flags, same
// node records - only who carries a subtree changes.
#ifndef RT_TREE_REGROUP
#define RT_TREE_REGROUP 2
#endif
constexpr uint32_t TREE_REGROUP_CAP = RT_TREE_REGROUP > 0 ? (uint32_t)TREE_TILE / (RT_TREE_REGROUP + 1) : 1u;
// height (8 bits) | delta(l - 1) + 1 (8 bits) | delta(r) + 1 (8 bits): the fourth word of a border job's / a global deposit's second half
__device__ __forceinline__ uint32_t pack_hd(uint32_t height, int dl, int dr) { return height | ((uint32_t)(dl + 1) << 8) | ((uint32_t)(dr + 1) << 16); }
__device__ __forceinline__ void unpack_hd(uint32_t w, uint32_t& height, int& dl, int& dr) { height = w & 255u; dl = (int)((w >> 8) & 255u) - 1; dr = (int)((w >> 16) & 255u) - 1; }
__device__ __forceinline__ TreeJob unpack_job(const float4 q0, const float4 q1, const float4 q2) {
    TreeJob j;
    j.b.lo[0] = q0.x; j.b.lo[1] = q0.y; j.b.lo[2] = q0.z; j.b.hi[0] = q0.w; j.b.hi[1] = q1.x; j.b.hi[2] = q1.y;
    j.ref = __float_as_int(q1.z);
    unpack_hd(__float_as_uint(q1.w), j.height, j.dl, j.dr);
    j.l = __float_as_uint(q2.x); j.r = __float_as_uint(q2.y); j.seg_first = __float_as_uint(q2.z); j.seg_count = __float_as_uint(q2.w);
    return j;
}

template <int LEAF_MAX, class Seg>
__device__ __forceinline__ void build_tree_tile(const uint64_t* __restrict__ keys, int vb, uint32_t n, BvhNode* __restrict__ nodes,
                                                float4* __restrict__ jobs, uint32_t* __restrict__ job_count,
                                                const Box3& leaf_box, const Seg& seg, const BorderMem bm) {
    __shared__ int s_delta[TREE_TILE + 1];          // s_delta[k] = delta(L0 - 1 + k)
    __shared__ uint32_t s_flag[TREE_TILE];          // bit side: that child of split L0 + k has been deposited
    __shared__ float4 s_a[2 * TREE_TILE], s_b[2 * TREE_TILE];
    __shared__ float4 s_carry[2 * 3];               // the (at most two) subtrees whose next split lies outside the tile, as job records
    __shared__ float4 s_re[TREE_REGROUP_CAP * 3];   // subtrees handed over at the regroup point
    __shared__ uint32_t s_ncarry, s_norph, s_nre, s_base;
    const uint32_t tid = threadIdx.x, L0 = blockIdx.x * (uint32_t)TREE_TILE, leaf = L0 + tid;
    const uint64_t k0 = leaf < n ? __ldg(keys + leaf) >> vb : 0ull;
    s_delta[tid + 1] = leaf + 1u < n ? delta_adjacent(k0, __ldg(keys + leaf + 1) >> vb, leaf) : -1;
    if (tid == 0) { s_delta[0] = L0 > 0u ? delta_adjacent(__ldg(keys + L0 - 1) >> vb, k0, L0 - 1u) : -1; s_ncarry = 0u; s_norph = 0u; s_nre = 0u; }
    s_flag[tid] = 0u;
    __syncthreads();

    // In the tile the subtree is {lt, rt} (tile-local ends), box, ref, height; a deposit is two float4:
    // {lo.xyz hi.x} {hi.y hi.z ref height | far_end_local << 8}.
    TreeJob j = {};
    uint32_t lt = tid, rt = tid;
    // climbs until the subtree is a root, leaves the tile, is the first to arrive at its split (all: false) or has merged max_merges times (true)
    auto climb = [&](int max_merges) -> bool {
        const uint32_t rel = L0 - j.seg_first;                               // tile-local -> segment-relative (mod 2^32)
        for (int merges = 0;;) {
            if (rt - lt + 1u == j.seg_count) { j.l = L0 + lt; j.r = L0 + rt; seg.on_root(j); return false; }
            const bool right = s_delta[rt + 1u] > s_delta[lt];
            const uint32_t slot = right ? rt : lt - 1u;                      // split g = L0 + slot; lt - 1 wraps to 2^32-1 at the tile's left end
            if (slot >= (uint32_t)TREE_TILE - 1u) {                          // next split outside the tile -> border job
                float4* q = s_carry + 3 * atomicAdd(&s_ncarry, 1u);
                q[0] = make_float4(j.b.lo[0], j.b.lo[1], j.b.lo[2], j.b.hi[0]);
                q[1] = make_float4(j.b.hi[1], j.b.hi[2], __int_as_float(j.ref), __uint_as_float(pack_hd(j.height, s_delta[lt], s_delta[rt + 1u])));
                q[2] = make_float4(__uint_as_float(L0 + lt), __uint_as_float(L0 + rt), __uint_as_float(j.seg_first), __uint_as_float(j.seg_count));
                return false;
            }
            const uint32_t side = right ? 0u : 1u;
            const float4 m0 = make_float4(j.b.lo[0], j.b.lo[1], j.b.lo[2], j.b.hi[0]);
            const float4 m1 = make_float4(j.b.hi[1], j.b.hi[2], __int_as_float(j.ref), __uint_as_float(j.height | ((right ? lt : rt) << 8)));
            s_a[2 * slot + side] = m0; s_b[2 * slot + side] = m1;
            fence_cta();
            if (atomicOr(&s_flag[slot], 1u << side) == 0u) return false;     // first arrival: the sibling finishes this node
            fence_cta();
            const float4 s0 = s_a[2 * slot + (side ^ 1u)], s1 = s_b[2 * slot + (side ^ 1u)];
            const uint32_t sw = __float_as_uint(s1.w);
            if (right) rt = sw >> 8; else lt = sw >> 8;
            j.b.lo[0] = fminf(j.b.lo[0], s0.x); j.b.lo[1] = fminf(j.b.lo[1], s0.y); j.b.lo[2] = fminf(j.b.lo[2], s0.z);
            j.b.hi[0] = fmaxf(j.b.hi[0], s0.w); j.b.hi[1] = fmaxf(j.b.hi[1], s1.x); j.b.hi[2] = fmaxf(j.b.hi[2], s1.y);
            const uint32_t count = rt - lt + 1u;
            if (count <= (uint32_t)LEAF_MAX) { j.ref = leaf_ref(lt + rel, count); j.height = 0; }
            else {                                                           // live node: one thread writes the whole 64-B record
                // its slot: right end of the range if it is a left child (will merge to the right), else the left end
                const uint32_t ns = (count != j.seg_count && s_delta[rt + 1u] > s_delta[lt]) ? rt : lt;
                j.ref = (int32_t)(ns + rel); j.height = max(j.height, sw & 255u) + 1u;
                float4* dst = reinterpret_cast<float4*>(nodes + L0 + ns);
                dst[2 * side] = m0; dst[2 * side + 1] = make_float4(m1.x, m1.y, m1.z, __uint_as_float(__float_as_uint(m1.w) & 255u));
                dst[2 * (side ^ 1u)] = s0; dst[2 * (side ^ 1u) + 1] = make_float4(s1.x, s1.y, s1.z, __uint_as_float(sw & 255u));
            }
            if (max_merges > 0 && ++merges == max_merges) return true;
        }
    };
    bool climbing = false;
    if (leaf < n) {
        j.b = leaf_box; j.height = 0;
        seg.seg_of(k0, j.seg_first, j.seg_count);
        j.ref = leaf_ref(leaf - j.seg_first, 1);
        climbing = climb(RT_TREE_REGROUP);
    }
#if RT_TREE_REGROUP > 0
    if (climbing) {                                                           // hand the subtree over
        float4* q = s_re + 3 * atomicAdd(&s_nre, 1u);
        q[0] = make_float4(j.b.lo[0], j.b.lo[1], j.b.lo[2], j.b.hi[0]);
        q[1] = make_float4(j.b.hi[1], j.b.hi[2], __int_as_float(j.ref), __uint_as_float(j.height | (lt << 8) | (rt << 17)));
        q[2] = make_float4(__uint_as_float(j.seg_first), __uint_as_float(j.seg_count), 0.0f, 0.0f);
    }
    __syncthreads();
    if (tid < s_nre) {
        const float4 q0 = s_re[3 * tid], q1 = s_re[3 * tid + 1], q2 = s_re[3 * tid + 2];
        j.b.lo[0] = q0.x; j.b.lo[1] = q0.y; j.b.lo[2] = q0.z; j.b.hi[0] = q0.w; j.b.hi[1] = q1.x; j.b.hi[2] = q1.y;
        j.ref = __float_as_int(q1.z);
        const uint32_t w = __float_as_uint(q1.w);
        j.height = w & 255u; lt = (w >> 8) & 511u; rt = w >> 17;
        j.seg_first = __float_as_uint(q2.x); j.seg_count = __float_as_uint(q2.y);
        climb(0);
    }
#else
    (void)climbing;
#endif
    __syncthreads();
    const uint32_t f = s_flag[tid];
    const bool orphan = f == 1u || f == 2u;
#if RT_TREE_INLINE_BORDER
    (void)jobs; (void)job_count;
    if (tid < s_ncarry) climb_global<LEAF_MAX>(unpack_job(s_carry[3 * tid], s_carry[3 * tid + 1], s_carry[3 * tid + 2]), nodes, bm.xchg, bm.far_end, bm.arrived, seg);
    if (orphan) {
        const uint32_t side = f - 1u, g = L0 + tid;
        const float4 m0 = s_a[2 * tid + side], m1 = s_b[2 * tid + side];
        const uint32_t farl = __float_as_uint(m1.w) >> 8;
        TreeJob o;
        o.b.lo[0] = m0.x; o.b.lo[1] = m0.y; o.b.lo[2] = m0.z; o.b.hi[0] = m0.w; o.b.hi[1] = m1.x; o.b.hi[2] = m1.y;
        o.ref = __float_as_int(m1.z); o.height = __float_as_uint(m1.w) & 255u;
        if (side == 0u) { o.l = L0 + farl; o.r = g; o.dl = s_delta[farl]; o.dr = s_delta[tid + 1u]; }
        else { o.l = g + 1u; o.r = L0 + farl; o.dl = s_delta[tid + 1u]; o.dr = s_delta[farl + 1u]; }
        seg.seg_of(k0, o.seg_first, o.seg_count);
        climb_global<LEAF_MAX>(o, nodes, bm.xchg, bm.far_end, bm.arrived, seg);
    }
    return;
#endif
    (void)bm;
    const uint32_t ib = orphan ? atomicAdd(&s_norph, 1u) : 0u;
    __syncthreads();
    const uint32_t ncarry = s_ncarry;
    if (tid == 0 && ncarry + s_norph) s_base = atomicAdd(job_count, ncarry + s_norph);
    __syncthreads();
    if (tid < ncarry) {
        float4* q = jobs + 3 * (size_t)(s_base + tid);
        q[0] = s_carry[3 * tid]; q[1] = s_carry[3 * tid + 1]; q[2] = s_carry[3 * tid + 2];
    }
    if (orphan) {
        // an orphan deposited at split g = L0 + tid covers [far, g] (left child, merges to the right) or [g + 1, far] (right child); both lie
        // in the segment of leaf g: a right child that reaches back to a segment's first leaf would be that segment's root, not an orphan
        const uint32_t side = f - 1u, g = L0 + tid;
        float4* q = jobs + 3 * (size_t)(s_base + ncarry + ib);
        float4 m1 = s_b[2 * tid + side];
        const uint32_t farl = __float_as_uint(m1.w) >> 8, far = L0 + farl, h = __float_as_uint(m1.w) & 255u;
        m1.w = __uint_as_float(side == 0u ? pack_hd(h, s_delta[farl], s_delta[tid + 1u]) : pack_hd(h, s_delta[tid + 1u], s_delta[farl + 1u]));
        uint32_t sf, sc;
        seg.seg_of(k0, sf, sc);
        q[0] = s_a[2 * tid + side]; q[1] = m1;
        q[2] = side == 0u ? make_float4(__uint_as_float(far), __uint_as_float(g), __uint_as_float(sf), __uint_as_float(sc))
                          : make_float4(__uint_as_float(g + 1u), __uint_as_float(far), __uint_as_float(sf), __uint_as_float(sc));
    }
}
// capacity of the border-job queue: 2 unfinished subtrees + one orphan per straddling ancestor (tree height <= 64 key
// bits + 32 index bits) of each border leaf, and never more than one orphan per split of the tile
__host__ __device__ constexpr uint32_t tree_jobs_per_tile() { return 2u + 2u * 96u; }
inline uint64_t tree_job_capacity_host(uint32_t n) {
    const uint64_t tiles = ((uint64_t)n + TREE_TILE - 1) / TREE_TILE;
    const uint64_t per_tile = tree_jobs_per_tile() < (uint32_t)TREE_TILE + 2u ? tree_jobs_per_tile() : (uint32_t)TREE_TILE + 2u;
    const uint64_t cap = tiles * per_tile;
    return cap < 2ull * n + 2 ? cap : 2ull * n + 2;
}

// Phase 2: a border job climbs through global memory. Arrival at a split is ONE acquire-add on the split's counter:
//   0 -> this is the first child: it deposits {half, far end}, release-adds 1 more (counter 2 = "deposit complete") and retires;
//   2 -> the sibling's deposit is complete (and visible: the acquire-add read the value of its release-add): union, write the finished node
//        into its Karras slot, go on one level;
//   1 -> the sibling has announced itself but its deposit is still in flight: the job keeps its place and looks again (acquire load) on the
//        next turn of its lane's loop. The sibling never waits for anybody, so this resolves after one store latency.
// The second child - the one that carries the chain on - pays one atomic and one load round trip per level and no fence; the release fence
// is paid by the child that retires. (The deposit-first protocol, stores -> fence -> add -> fence -> loads on every arrival, spent 54 % of
// this kernel's stall samples in its two fences: profiles/README.md r2_y.)
// A subtree carries delta(l - 1) and delta(r) (8 bits each, +1 so that -1 fits, next to the height in the deposit): the union of two
// siblings inherits the outer delta of each, so the direction of the next merge needs no key loads (14 % of the stall samples and another
// dependent L2 round trip per level).
__device__ __forceinline__ uint32_t atom_add_acquire(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acquire.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_add_release(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one turn of a job: an arrival (waiting == false) or another look at a sibling whose deposit was in flight (waiting == true);
// false when the job retires (first arrival at its split, or it completed its segment's root)
template <int LEAF_MAX, class Seg>
__device__ __forceinline__ bool climb_level(TreeJob& j, bool& waiting, BvhNode* __restrict__ nodes,
                                            float4* __restrict__ xchg, uint32_t* __restrict__ far_end, uint32_t* __restrict__ arrived, const Seg& seg) {
    const bool right = j.dr > j.dl;
    const uint32_t g = right ? j.r : j.l - 1u, side = right ? 0u : 1u;
    const float4 m0 = make_float4(j.b.lo[0], j.b.lo[1], j.b.lo[2], j.b.hi[0]);
    float4* dep = xchg + 4 * (size_t)g;
    if (!waiting) {
        const uint32_t seen = atom_add_acquire(arrived + g, 1u);
        if (seen == 0u) {
            dep[2 * side] = m0;
            dep[2 * side + 1] = make_float4(j.b.hi[1], j.b.hi[2], __int_as_float(j.ref), __uint_as_float(pack_hd(j.height, j.dl, j.dr)));
            far_end[2 * (size_t)g + side] = right ? j.l : j.r;
            red_add_release(arrived + g, 1u);
            return false;
        }
        if (seen == 1u) { waiting = true; return true; }
    } else {
        if (ld_acquire(arrived + g) != 3u) return true;
        waiting = false;
    }
    const float4 s0 = __ldcg(dep + 2 * (side ^ 1u));
    float4 s1 = __ldcg(dep + 2 * (side ^ 1u) + 1);
    const uint32_t sfar = __ldcg(far_end + 2 * (size_t)g + (side ^ 1u));
    uint32_t sh; int sdl, sdr;
    unpack_hd(__float_as_uint(s1.w), sh, sdl, sdr);
    s1.w = __uint_as_float(sh);
    const float4 m1n = make_float4(j.b.hi[1], j.b.hi[2], __int_as_float(j.ref), __uint_as_float(j.height));
    if (right) { j.r = sfar; j.dr = sdr; } else { j.l = sfar; j.dl = sdl; }
    j.b.lo[0] = fminf(j.b.lo[0], s0.x); j.b.lo[1] = fminf(j.b.lo[1], s0.y); j.b.lo[2] = fminf(j.b.lo[2], s0.z);
    j.b.hi[0] = fmaxf(j.b.hi[0], s0.w); j.b.hi[1] = fmaxf(j.b.hi[1], s1.x); j.b.hi[2] = fmaxf(j.b.hi[2], s1.y);
    const uint32_t count = j.r - j.l + 1u;
    const bool root = count == j.seg_count;
    if (count <= (uint32_t)LEAF_MAX) { j.ref = leaf_ref(j.l - j.seg_first, count); j.height = 0; }
    else {
        const uint32_t ns = (!root && j.dr > j.dl) ? j.r : j.l;
        j.ref = (int32_t)(ns - j.seg_first); j.height = max(j.height, sh) + 1u;
        float4* dst = reinterpret_cast<float4*>(nodes + ns);
        dst[2 * side] = m0; dst[2 * side + 1] = m1n; dst[2 * (side ^ 1u)] = s0; dst[2 * (side ^ 1u) + 1] = s1;
    }
    if (root) { seg.on_root(j); return false; }
    return true;
}
template <int LEAF_MAX, class Seg>
__device__ __forceinline__ void climb_global(TreeJob j, BvhNode* __restrict__ nodes,
                                             float4* __restrict__ xchg, uint32_t* __restrict__ far_end, uint32_t* __restrict__ arrived, const Seg& seg) {
    if (j.r - j.l + 1u == j.seg_count) { seg.on_root(j); return; }
    bool waiting = false;
    while (climb_level<LEAF_MAX>(j, waiting, nodes, xchg, far_end, arrived, seg)) {}
}

// Every lane walks its own strided share of the job queue and fetches its next job the moment the current one retires: most jobs retire at
// their first deposit and a few climb many levels, so one job per thread left a warp running with 4.8 of 32 lanes for as long as its longest climb
// (profiles/README.md r2_y). The grid is what the device holds at once (RT_BORDER_CTAS_PER_SM), not the length of the queue.
#ifndef RT_BORDER_CTAS_PER_SM
#define RT_BORDER_CTAS_PER_SM 12
#endif
template <int LEAF_MAX, class Seg>
__global__ void __launch_bounds__(128, RT_BORDER_CTAS_PER_SM) k_tree_border(const uint64_t* __restrict__ keys, int vb, uint32_t n, BvhNode* __restrict__ nodes,
                                                    const float4* __restrict__ jobs, const uint32_t* __restrict__ job_count,
                                                    float4* __restrict__ xchg, uint32_t* __restrict__ far_end, uint32_t* __restrict__ arrived, const Seg seg) {
    (void)keys; (void)vb; (void)n;
    const uint32_t n_jobs = *job_count, stride = gridDim.x * blockDim.x;
    uint32_t next = blockIdx.x * blockDim.x + threadIdx.x;
    TreeJob j = {};
    bool have = false, waiting = false;
    for (;;) {
        if (!have) {
            if (next >= n_jobs) break;
            j = unpack_job(__ldg(jobs + 3 * (size_t)next), __ldg(jobs + 3 * (size_t)next + 1), __ldg(jobs + 3 * (size_t)next + 2));
            next += stride;
            if (j.r - j.l + 1u == j.seg_count) { seg.on_root(j); continue; }
        }
        have = climb_level<LEAF_MAX>(j, waiting, nodes, xchg, far_end, arrived, seg);
    }
}
// grid of the border kernel: the CTAs the device holds at once (the lanes loop over the queue), never more than one thread per possible job
inline uint32_t border_grid(uint32_t n) {
    static int per_sm = 0, sms[64] = {};
    if (per_sm == 0) {
        per_sm = RT_BORDER_CTAS_PER_SM;
        if (const char* v = getenv("RTCORE_BORDER_CTAS_PER_SM")) { const int k = atoi(v); if (k >= 1 && k <= 64) per_sm = k; }
    }
    int dev = 0;
    cudaGetDevice(&dev);
    int sm = 148;
    if (dev >= 0 && dev < 64) {
        if (sms[dev] == 0 && cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms[dev] = 148;
        sm = sms[dev];
    }
    const uint64_t by_jobs = (tree_job_capacity_host(n) + 127) / 128;
    const uint64_t resident = (uint64_t)sm * (uint64_t)per_sm;
    const uint64_t g = by_jobs < resident ? by_jobs : resident;
    return g ? (uint32_t)g : 1u;
}

// PRESORTED: the sorted triangle records are already in place (written by k_seg_setup_sort): stream them, nothing to gather or to write
template <bool PRESORTED>
__global__ void __launch_bounds__(TREE_TILE, RT_TREE_MIN_CTAS) k_refit_tris(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int vb, uint32_t n,
                                                         const TriRec* __restrict__ unsorted, TriRec* __restrict__ sorted,
                                                         BvhNode* __restrict__ nodes, const TriSegments seg,
                                                         float4* __restrict__ jobs, uint32_t* __restrict__ job_count, const BorderMem bm) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    Box3 b = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
    if (leaf < n) {
        float4 q0, q1, q2;
        if (PRESORTED) {
            const float4* src = reinterpret_cast<const float4*>(sorted + leaf);
            q0 = __ldcs(src); q1 = __ldcs(src + 1); q2 = __ldcs(src + 2);
        } else {
            const uint32_t src_i = vb ? (uint32_t)(__ldg(keys + leaf) & ((1ull << vb) - 1ull)) : __ldg(vals + leaf);
            const float4* src = reinterpret_cast<const float4*>(unsorted + src_i);
            q0 = __ldg(src); q1 = __ldg(src + 1); q2 = __ldg(src + 2);
            q2.w = __uint_as_float(__float_as_uint(q2.w) >> 24);  // geometry flags (the low 24 bits carried the BLAS id through the sort)
            float4* dst = reinterpret_cast<float4*>(sorted + leaf);
            dst[0] = q0; dst[1] = q1; dst[2] = q2;
        }
        b.lo[0] = fminf(fminf(q0.x, q0.w), q1.z); b.lo[1] = fminf(fminf(q0.y, q1.x), q1.w); b.lo[2] = fminf(fminf(q0.z, q1.y), q2.x);
        b.hi[0] = fmaxf(fmaxf(q0.x, q0.w), q1.z); b.hi[1] = fmaxf(fmaxf(q0.y, q1.x), q1.w); b.hi[2] = fmaxf(fmaxf(q0.z, q1.y), q2.x);
    }
    build_tree_tile<BLAS_LEAF_MAX>(keys, vb, n, nodes, jobs, job_count, b, seg, bm);
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
        R.active = (ok ? 1u : 0u) | ((uint32_t)geo << 1);      // bit 0: traversable; bits 1..: geometry count - 1 of its BLAS (exact SBT range check)
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
    const uint64_t key = (inst[i].active & 1u) ? (uint64_t)morton30(plo, phi, slo, shi) : 0x3FFFFFFFull;
    if (vb) keys[i] = (key << vb) | (uint64_t)i;
    else { keys[i] = key; vals[i] = i; }
}

__global__ void __launch_bounds__(TREE_TILE, RT_TREE_MIN_CTAS) k_refit_inst(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int vb, uint32_t n, const InstanceRec* __restrict__ unsorted,
                                                         const float* __restrict__ boxes, InstanceRec* __restrict__ sorted,
                                                         BvhNode* __restrict__ nodes, const InstSegment seg,
                                                         float4* __restrict__ jobs, uint32_t* __restrict__ job_count, const BorderMem bm) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    Box3 b = {{0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}};
    if (leaf < n) {
        const uint32_t src_i = vb ? (uint32_t)(__ldg(keys + leaf) & ((1ull << vb) - 1ull)) : __ldg(vals + leaf);
        const float4* src = reinterpret_cast<const float4*>(unsorted + src_i);
        float4* dst = reinterpret_cast<float4*>(sorted + leaf);
#pragma unroll
        for (int k = 0; k < 6; ++k) dst[k] = __ldg(src + k);
        const float* bx = boxes + 6 * (size_t)src_i;
        b.lo[0] = bx[0]; b.lo[1] = bx[1]; b.lo[2] = bx[2]; b.hi[0] = bx[3]; b.hi[1] = bx[4]; b.hi[2] = bx[5];
    }
    build_tree_tile<TLAS_LEAF_MAX>(keys, vb, n, nodes, jobs, job_count, b, seg, bm);
}

inline int div_up(uint32_t a, uint32_t b) { return (int)((a + b - 1) / b); }

// ---- compaction (rt_compact_blas; VK_COPY_ACCELERATION_STRUCTURE_MODE_COMPACT_KHR) ---------------------------------------------
// Slot i of a segment is Karras' internal node i: its leaf range starts (or ends) at leaf i and extends in direction d for as long as
// the common prefix with leaf i stays longer than delta(i, i - d). The build writes the node only when that range holds more than
// BLAS_LEAF_MAX leaves (smaller subtrees collapse into the parent's leaf reference), which can be decided from the sorted records alone:
// live(i) <=> delta(i, i + BLAS_LEAF_MAX * d) > delta(i, i - d), with delta = -1 outside the segment.
constexpr int CPT_THREADS = 256, CPT_ITEMS = 8, CPT_CHUNK = CPT_THREADS * CPT_ITEMS;

__device__ __forceinline__ int delta_seg(const uint64_t* __restrict__ keys, int vb, uint32_t i, int64_t j, uint32_t first, uint32_t count) {
    if (j < (int64_t)first || j >= (int64_t)first + count) return -1;
    const uint64_t ka = __ldg(keys + i) >> vb, kb = __ldg(keys + j) >> vb;
    return ka == kb ? 64 + __clz((int)(i ^ (uint32_t)j)) : __clzll((long long)(ka ^ kb));
}
__device__ __forceinline__ bool slot_is_live(const uint64_t* __restrict__ keys, int vb, uint32_t n, const BlasRecord* __restrict__ records, uint32_t i,
                                             uint32_t& first) {
    first = 0;
    if (i >= n) return false;
    const BlasRecord& R = records[(uint32_t)((__ldg(keys + i) >> vb) >> MORTON_BITS)];
    first = R.first;
    const uint32_t count = R.tri_count;
    if (count <= (uint32_t)BLAS_LEAF_MAX || i - first + 2u > count) return false;           // no internal nodes / not a node slot
    const int dr = delta_seg(keys, vb, i, (int64_t)i + 1, first, count), dl = delta_seg(keys, vb, i, (int64_t)i - 1, first, count);
    const int d = dr > dl ? 1 : -1;
    const int dmin = d > 0 ? dl : dr;
    return delta_seg(keys, vb, i, (int64_t)i + (int64_t)BLAS_LEAF_MAX * d, first, count) > dmin;
}

// pass 1a: live slots per chunk
__global__ void __launch_bounds__(CPT_THREADS) k_compact_count(const uint64_t* __restrict__ keys, int vb, uint32_t n, const BlasRecord* __restrict__ records,
                                                              uint32_t* __restrict__ chunk_sums) {
    __shared__ uint32_t s_w[CPT_THREADS / 32];
    uint32_t c = 0, first;
#pragma unroll
    for (int k = 0; k < CPT_ITEMS; ++k) c += slot_is_live(keys, vb, n, records, blockIdx.x * CPT_CHUNK + k * CPT_THREADS + threadIdx.x, first) ? 1u : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < CPT_THREADS / 32; ++w) t += s_w[w]; chunk_sums[blockIdx.x] = t; }
}
// pass 1b: exclusive scan of the chunk sums (one CTA, serial over 1024-wide steps: n / 2048 values)
__global__ void __launch_bounds__(1024) k_compact_scan(uint32_t* __restrict__ chunk_sums, uint32_t n_chunks) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_chunks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_chunks ? chunk_sums[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            s_w[lane] = w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + (warp ? s_w[warp - 1] : 0u) + inc - v;
        if (i < n_chunks) chunk_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}
// pass 1c: cidx[i] = number of live slots before i; cidx[n] = live total. Items are strided (slot = chunk base + k * CPT_THREADS + tid),
// so the in-chunk rank is: live slots of earlier strides (all threads) + live slots of this stride in earlier threads.
__global__ void __launch_bounds__(CPT_THREADS) k_compact_index(const uint64_t* __restrict__ keys, int vb, uint32_t n, const BlasRecord* __restrict__ records,
                                                              const uint32_t* __restrict__ chunk_sums, uint32_t* __restrict__ cidx) {
    __shared__ uint32_t s_w[CPT_THREADS / 32];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = chunk_sums[blockIdx.x];
    __syncthreads();
    for (int k = 0; k < CPT_ITEMS; ++k) {
        const uint32_t i = blockIdx.x * CPT_CHUNK + k * CPT_THREADS + threadIdx.x;
        uint32_t first;
        const bool live = slot_is_live(keys, vb, n, records, i, first);
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        uint32_t before = s_run + __popc(m & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) before += s_w[w];
        if (i <= n) cidx[i] = before;                                   // i == n: the total (no slot n exists, so live is false there)
        __syncthreads();
        if (threadIdx.x == CPT_THREADS - 1) s_run = before + (live ? 1u : 0u);
        __syncthreads();
    }
}
// pass 2: live node i -> dst[cidx[i]]; internal child refs (relative to the segment's first slot) are re-based on the compact numbering
__global__ void __launch_bounds__(256) k_compact_nodes(const BvhNode* __restrict__ src, BvhNode* __restrict__ dst, const uint32_t* __restrict__ cidx,
                                                      const uint64_t* __restrict__ keys, int vb, uint32_t n, const BlasRecord* __restrict__ records) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = __ldg(cidx + i);
    if (__ldg(cidx + i + 1) == c) return;                                // dead slot
    const uint32_t first = records[(uint32_t)((__ldg(keys + i) >> vb) >> MORTON_BITS)].first;
    const uint32_t cfirst = __ldg(cidx + first);
    const float4* s4 = reinterpret_cast<const float4*>(src + i);
    float4 q0 = __ldg(s4), q1 = __ldg(s4 + 1), q2 = __ldg(s4 + 2), q3 = __ldg(s4 + 3);
    const int32_t r0 = __float_as_int(q1.z), r1 = __float_as_int(q3.z);
    if (r0 >= 0 && r0 < REF_SENTINEL_MIN) q1.z = __int_as_float((int32_t)(__ldg(cidx + first + (uint32_t)r0) - cfirst));
    if (r1 >= 0 && r1 < REF_SENTINEL_MIN) q3.z = __int_as_float((int32_t)(__ldg(cidx + first + (uint32_t)r1) - cfirst));
    float4* d4 = reinterpret_cast<float4*>(dst + c);
    d4[0] = q0; d4[1] = q1; d4[2] = q2; d4[3] = q3;
}

// the records of the compacted storage: node base = new nodes + compact index of the BLAS's first slot
__global__ void __launch_bounds__(256) k_compact_records(const BlasRecord* __restrict__ src, BlasRecord* __restrict__ dst, uint32_t n_blas,
                                                        const uint32_t* __restrict__ cidx, const BvhNode* nodes, const TriRec* tris) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blas) return;
    BlasRecord R = src[b];
    const uint32_t c0 = __ldg(cidx + R.first), c1 = __ldg(cidx + R.first + R.tri_count);
    R.nodes = nodes + c0; R.tris = tris + R.first; R.node_slots = c1 - c0;
    dst[b] = R;
}

}  // namespace

int launch_compact_records(const BlasRecord* src, BlasRecord* dst, uint32_t n_blas, const uint32_t* cidx, const BvhNode* nodes, const TriRec* tris, cudaStream_t st) {
    k_compact_records<<<div_up(n_blas, 256), 256, 0, st>>>(src, dst, n_blas, cidx, nodes, tris);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

size_t compact_scratch_bytes(uint32_t n) { return 4ull * (div_up(n + 1u, CPT_CHUNK) + 1) + 256; }

int launch_compact_index(const uint64_t* keys, int vb, uint32_t n, const BlasRecord* records, uint32_t* cidx, void* scratch, cudaStream_t st) {
    const uint32_t chunks = (uint32_t)div_up(n + 1u, CPT_CHUNK);          // slot n included: it receives the total
    uint32_t* chunk_sums = reinterpret_cast<uint32_t*>(scratch);
    k_compact_count<<<chunks, CPT_THREADS, 0, st>>>(keys, vb, n, records, chunk_sums);
    k_compact_scan<<<1, 1024, 0, st>>>(chunk_sums, chunks);
    k_compact_index<<<chunks, CPT_THREADS, 0, st>>>(keys, vb, n, records, chunk_sums, cidx);
    return cudaGetLastError() == cudaSuccess ? 3 : -1;
}
int launch_compact_nodes(const BvhNode* src, BvhNode* dst, const uint32_t* cidx, const uint64_t* keys, int vb, uint32_t n,
                         const BlasRecord* records, cudaStream_t st) {
    if (n == 0) return 0;
    k_compact_nodes<<<div_up(n, 256), 256, 0, st>>>(src, dst, cidx, keys, vb, n, records);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

uint32_t tree_job_capacity(uint32_t n) { return (uint32_t)tree_job_capacity_host(n); }

int launch_blas_build(const BlasBuildArgs& a, cudaStream_t st, const BuildEvents* ev, bool* sorted_in_b) {
    int launches = 0;
    bool presorted = false;
    *sorted_in_b = false;
    if (a.n_tris == 0) return 0;
    if (ev) cudaEventRecord(ev->e[0], st);
    const int vb = a.sort.packed_val_bits;
    if (a.reuse_keys) {
        // refit-only update (RT_BUILD_MODE_REFIT): bake the new vertices, keep the sorted records of the last full build (= the topology)
        k_tri_setup<<<div_up(a.n_tris, SETUP_CHUNK), SETUP_THREADS, 0, st>>>(a.geoms, a.geom_tri_first, a.n_geoms, a.n_tris, a.tris_unsorted, a.bounds_ordered);
        ++launches;
        if (ev) { cudaEventRecord(ev->e[1], st); cudaEventRecord(ev->e[2], st); }
    } else if (a.sort.seg_records && vb > 0 && a.sort.seg_fused) {
        // every BLAS fits one CTA: setup + Morton + sort fused, one launch (its time is reported as sort_ms)
        // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: set it once per device
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            if (cudaFuncSetAttribute(k_seg_setup_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEG_SMEM_BYTES) != cudaSuccess) return -1;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        if (ev) { cudaEventRecord(ev->e[1], st); cudaEventRecord(ev->e[2], st); }
#if RT_SEG_LIGHT
        {
            static bool attr2_set[64] = {};
            if (dev < 0 || dev >= 64 || !attr2_set[dev]) {
                if (cudaFuncSetAttribute(k_seg2_setup_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEG2_SMEM_BYTES) != cudaSuccess) return -1;
                if (dev >= 0 && dev < 64) attr2_set[dev] = true;
            }
        }
        k_seg2_setup_sort<<<a.sort.n_segments, SEG2_THREADS, SEG2_SMEM_BYTES, st>>>(a.geoms, a.geom_tri_first, a.n_geoms, a.sort.seg_records, a.tris_unsorted, a.s.keys_b, vb,
                                                                                    RT_SEG_EMIT_SORTED ? a.tris_sorted : nullptr);
#else
        k_seg_setup_sort<<<a.sort.n_segments, SEG_THREADS, SEG_SMEM_BYTES, st>>>(a.geoms, a.geom_tri_first, a.n_geoms, a.sort.seg_records, a.tris_unsorted, a.s.keys_b, vb,
                                                                                 RT_SEG_EMIT_SORTED ? a.tris_sorted : nullptr);
#endif
        presorted = RT_SEG_EMIT_SORTED != 0;
        ++launches;
        *sorted_in_b = true;
    } else {
        k_tri_setup<<<div_up(a.n_tris, SETUP_CHUNK), SETUP_THREADS, 0, st>>>(a.geoms, a.geom_tri_first, a.n_geoms, a.n_tris, a.tris_unsorted, a.bounds_ordered);
        ++launches;
        if (ev) cudaEventRecord(ev->e[1], st);
        k_tri_morton<<<div_up(a.n_tris, 256), 256, 0, st>>>(a.tris_unsorted, a.n_tris, a.bounds_ordered, a.s.keys_a, a.s.vals_a, vb);
        ++launches;
        if (ev) cudaEventRecord(ev->e[2], st);
        int sl = sort_pairs(a.sort, a.s.keys_a, a.s.keys_b, vb ? nullptr : a.s.vals_a, vb ? nullptr : a.s.vals_b, a.s.sort_scratch, a.s.error_flag, st, sorted_in_b);
        if (sl < 0) return -1;
        launches += sl;
    }
    const uint64_t* keys = a.reuse_keys ? a.reuse_keys : (*sorted_in_b ? a.s.keys_b : a.s.keys_a);
    const uint32_t* vals = a.reuse_keys ? a.reuse_vals : (*sorted_in_b ? a.s.vals_b : a.s.vals_a);
    if (ev) cudaEventRecord(ev->e[3], st);
    if (cudaMemsetAsync(a.s.arrived, 0, sizeof(uint32_t) * ((size_t)a.n_tris + 1), st) != cudaSuccess) return -1;   // counters + job count
    if (ev) cudaEventRecord(ev->e[4], st);
    {
        const TriSegments seg{a.records, keys, vb};
        const uint32_t tiles = (uint32_t)div_up(a.n_tris, TREE_TILE);
        uint32_t* job_count = a.s.arrived + a.n_tris;
        const BorderMem bm{a.s.xchg, a.s.far_end, a.s.arrived};
        if (presorted) k_refit_tris<true><<<tiles, TREE_TILE, 0, st>>>(keys, vals, vb, a.n_tris, a.tris_unsorted, a.tris_sorted, a.nodes, seg, a.s.jobs, job_count, bm);
        else k_refit_tris<false><<<tiles, TREE_TILE, 0, st>>>(keys, vals, vb, a.n_tris, a.tris_unsorted, a.tris_sorted, a.nodes, seg, a.s.jobs, job_count, bm);
#if !RT_TREE_INLINE_BORDER
        k_tree_border<BLAS_LEAF_MAX, TriSegments><<<border_grid(a.n_tris), 128, 0, st>>>(keys, vb, a.n_tris, a.nodes, a.s.jobs, job_count,
                                                                                                          a.s.xchg, a.s.far_end, a.s.arrived, seg);
        ++launches;
#endif
    }
    launches += 1;
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
    if (cudaMemsetAsync(a.s.arrived, 0, sizeof(uint32_t) * ((size_t)a.n + 1), st) != cudaSuccess) return -1;
    {
        const InstSegment seg{a.n, a.root_out, a.bounds_out};
        uint32_t* job_count = a.s.arrived + a.n;
        const BorderMem bm{a.s.xchg, a.s.far_end, a.s.arrived};
        k_refit_inst<<<div_up(a.n, TREE_TILE), TREE_TILE, 0, st>>>(keys, vals, vb, a.n, a.inst_unsorted, a.boxes_unsorted, a.inst_sorted, a.nodes, seg, a.s.jobs, job_count, bm);
#if !RT_TREE_INLINE_BORDER
        k_tree_border<TLAS_LEAF_MAX, InstSegment><<<border_grid(a.n), 128, 0, st>>>(keys, vb, a.n, a.nodes, a.s.jobs, job_count,
                                                                                                     a.s.xchg, a.s.far_end, a.s.arrived, seg);
        ++launches;
#endif
    }
    launches += 1;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace rt
