// rt_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code). See rt_oracle.h.
//
// Build: g++ -O2 -std=c++17 -fopenmp -ffp-contract=off -fPIC -shared (oracle/Makefile).
// -ffp-contract=off is REQUIRED: every fp32 expression below is evaluated exactly as written
// (IEEE-754 binary32, round-to-nearest-even, no fused multiply-add), which is what the CUDA
// product reproduces with -fmad=false. All ids/keys are integers.
//
// PARITY UNPINNED by the reference (no tests / golden data upstream); pinned against the
// analytically derived known answers of the sample scene in tests/test_oracle_sample.py.
#include "rt_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

namespace {

// ------------------------------------------------------------------------------------------
// fp32 vector helpers with explicit evaluation order
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline float comp(const V3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
static inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross3(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// row-major 3x4 affine: Vulkan VkTransformMatrixKHR (main.cpp:684-695,835-846)
static inline V3 xform_point(const float* m, V3 p) {
    return {((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
            ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
            ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]};
}
static inline V3 xform_vec(const float* m, V3 d) {
    return {(m[0] * d.x + m[1] * d.y) + m[2] * d.z,
            (m[4] * d.x + m[5] * d.y) + m[6] * d.z,
            (m[8] * d.x + m[9] * d.y) + m[10] * d.z};
}
// n_world = (world->object)^T * n_object
static inline V3 xform_normal(const float* w2o, V3 n) {
    return {(w2o[0] * n.x + w2o[4] * n.y) + w2o[8] * n.z,
            (w2o[1] * n.x + w2o[5] * n.y) + w2o[9] * n.z,
            (w2o[2] * n.x + w2o[6] * n.y) + w2o[10] * n.z};
}
// world->object = inverse(object->world), computed in fp64 (no contraction) and rounded once to fp32.
static bool invert3x4(const float* m, float* out) {
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double tx = m[3], ty = m[7], tz = m[11];
    double c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
    double c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
    double c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
    double det = (a * c00 + b * c10) + c * c20;
    if (!(det != 0.0) || !std::isfinite(det)) { for (int k = 0; k < 12; ++k) out[k] = 0.0f; return false; }
    double inv = 1.0 / det;
    double r00 = c00 * inv, r01 = c01 * inv, r02 = c02 * inv;
    double r10 = c10 * inv, r11 = c11 * inv, r12 = c12 * inv;
    double r20 = c20 * inv, r21 = c21 * inv, r22 = c22 * inv;
    double r03 = -((r00 * tx + r01 * ty) + r02 * tz);
    double r13 = -((r10 * tx + r11 * ty) + r12 * tz);
    double r23 = -((r20 * tx + r21 * ty) + r22 * tz);
    out[0] = (float)r00; out[1] = (float)r01; out[2]  = (float)r02; out[3]  = (float)r03;
    out[4] = (float)r10; out[5] = (float)r11; out[6]  = (float)r12; out[7]  = (float)r13;
    out[8] = (float)r20; out[9] = (float)r21; out[10] = (float)r22; out[11] = (float)r23;
    return true;
}

static inline uint32_t pcg_hash(uint32_t v) {
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

// ------------------------------------------------------------------------------------------
// scene containers
// ------------------------------------------------------------------------------------------
struct Tri { V3 v0, v1, v2; uint32_t geo, prim, flags; };   // flags: VkGeometryFlagBitsKHR of its geometry

constexpr int32_t REF_DONE = 0x7FFFFFFF;
constexpr int32_t REF_EMPTY = 0x7FFFFFFD;
constexpr int LEAF_MAX = 2;      // triangles per BLAS leaf (the product's BLAS_LEAF_MAX; measured on B200: 1/2/3/4/6/8 -> 3124/3236/3198/3145/2981/2823 Mrays/s)
static inline int32_t leaf_ref(uint32_t first, uint32_t count) { return ~(int32_t)((first << 3) | (count - 1)); }
static inline bool ref_is_leaf(int32_t r) { return r < 0; }
static inline uint32_t leaf_first(int32_t r) { return ((uint32_t)~r) >> 3; }
static inline uint32_t leaf_count(int32_t r) { return (((uint32_t)~r) & 7u) + 1u; }

struct NodeHalf { float lo[3]; float hi[3]; int32_t ref; uint32_t height; };  // 32 B
struct Node { NodeHalf c[2]; };                                               // 64 B
static_assert(sizeof(Node) == 64, "node layout");

struct Box { float lo[3], hi[3]; };
static inline Box empty_box() { return {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; }
static inline void grow(Box& b, V3 p) {
    b.lo[0] = fminf(b.lo[0], p.x); b.lo[1] = fminf(b.lo[1], p.y); b.lo[2] = fminf(b.lo[2], p.z);
    b.hi[0] = fmaxf(b.hi[0], p.x); b.hi[1] = fmaxf(b.hi[1], p.y); b.hi[2] = fmaxf(b.hi[2], p.z);
}
static inline void grow(Box& b, const Box& o) {
    for (int k = 0; k < 3; ++k) { b.lo[k] = fminf(b.lo[k], o.lo[k]); b.hi[k] = fmaxf(b.hi[k], o.hi[k]); }
}
static inline Box tri_box(const Tri& t) { Box b = empty_box(); grow(b, t.v0); grow(b, t.v1); grow(b, t.v2); return b; }

static inline uint32_t expand10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// 30-bit Morton code of the centre c of a primitive box inside scene box [lo,hi].
static inline uint32_t quant10(float c, float lo, float hi) {
    float ext = hi - lo;
    float inv = ext > 0.0f ? 1024.0f / ext : 0.0f;
    float q = (c - lo) * inv;
    q = fminf(q, 1023.0f);
    if (!(q >= 0.0f)) q = 0.0f;
    return (uint32_t)q;
}
static inline uint32_t morton30(const Box& pb, const Box& scene) {
    float cx = (pb.lo[0] + pb.hi[0]) * 0.5f, cy = (pb.lo[1] + pb.hi[1]) * 0.5f, cz = (pb.lo[2] + pb.hi[2]) * 0.5f;
    uint32_t x = quant10(cx, scene.lo[0], scene.hi[0]);
    uint32_t y = quant10(cy, scene.lo[1], scene.hi[1]);
    uint32_t z = quant10(cz, scene.lo[2], scene.hi[2]);
    return (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
}

// Karras-2012 LBVH over N primitive boxes with keys; leaves collapsed to <= leaf_max primitives.
// Node numbering is the paper's (node i covers a range that starts or ends at sorted primitive i; the children of a
// node with split gamma are nodes gamma and gamma + 1). The ranges/splits are found top-down exactly as in the paper;
// the GPU product finds the same tree bottom-up, which makes the bit-for-bit comparison an independent check.
struct Lbvh {
    std::vector<Node> nodes;        // N-1 slots (Karras numbering), unreachable ones are garbage
    std::vector<uint64_t> keys;     // sorted
    std::vector<uint32_t> prims;    // sorted position -> original primitive
    int32_t root = REF_EMPTY;
    uint32_t max_depth = 0;
    Box bounds = empty_box();
};

static inline int clz64(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
static inline int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }

static void lbvh_build(Lbvh& bvh, const std::vector<Box>& pboxes, const std::vector<uint64_t>& unsorted_keys, int leaf_max) {
    const int64_t N = (int64_t)pboxes.size();
    bvh.keys.resize(N); bvh.prims.resize(N);
    bvh.nodes.clear(); bvh.root = REF_EMPTY; bvh.max_depth = 0;
    if (N == 0) return;
    std::vector<std::pair<uint64_t, uint32_t>> kv(N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) kv[i] = {unsorted_keys[i], (uint32_t)i};
    auto cmp = [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first < b.first; };
#ifdef _OPENMP
    __gnu_parallel::stable_sort(kv.begin(), kv.end(), cmp);
#else
    std::stable_sort(kv.begin(), kv.end(), cmp);
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) { bvh.keys[i] = kv[i].first; bvh.prims[i] = kv[i].second; }
    kv.clear(); kv.shrink_to_fit();

    if (N <= leaf_max) {
        bvh.root = leaf_ref(0, (uint32_t)N);
        return;
    }
    const std::vector<uint64_t>& K = bvh.keys;
    auto delta = [&](int64_t i, int64_t j) -> int {
        if (j < 0 || j >= N) return -1;
        uint64_t a = K[i], b = K[j];
        if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
        return clz64(a ^ b);
    };
    const int64_t NI = N - 1;
    bvh.nodes.resize(NI);
    std::vector<int32_t> other_end(NI);           // j of node i; range = [min(i,j), max(i,j)]
    std::vector<uint32_t> parent_of_node(NI, 0xFFFFFFFFu), parent_of_leaf(N);  // (parent << 1) | side
    std::vector<int32_t> left_child(NI), right_child(NI);                       // >=0 internal, ~idx leaf position
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < NI; ++i) {
        int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
        if (i == 0) d = 1;
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
        int64_t first = std::min(i, j), last = std::max(i, j);
        other_end[i] = (int32_t)j;
        if (first == gamma) { left_child[i] = ~(int32_t)gamma; parent_of_leaf[gamma] = ((uint32_t)i << 1) | 0u; }
        else { left_child[i] = (int32_t)gamma; parent_of_node[gamma] = ((uint32_t)i << 1) | 0u; }
        if (last == gamma + 1) { right_child[i] = ~(int32_t)(gamma + 1); parent_of_leaf[gamma + 1] = ((uint32_t)i << 1) | 1u; }
        else { right_child[i] = (int32_t)(gamma + 1); parent_of_node[gamma + 1] = ((uint32_t)i << 1) | 1u; }
    }
    // atomic bottom-up refit; the second thread to arrive at a node owns it and continues upward
    std::vector<std::atomic<int>> arrived(NI);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < NI; ++i) arrived[i].store(0, std::memory_order_relaxed);
    Box root_box = empty_box(); uint32_t root_height = 0; int32_t root_ref = REF_EMPTY;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t leaf = 0; leaf < N; ++leaf) {
        Box b = pboxes[bvh.prims[leaf]];
        int32_t ref = leaf_ref((uint32_t)leaf, 1);
        uint32_t height = 0;
        uint32_t p = parent_of_leaf[leaf];
        for (;;) {
            uint32_t node = p >> 1, side = p & 1u;
            NodeHalf& h = bvh.nodes[node].c[side];
            for (int k = 0; k < 3; ++k) { h.lo[k] = b.lo[k]; h.hi[k] = b.hi[k]; }
            h.ref = ref; h.height = height;
            if (arrived[node].fetch_add(1, std::memory_order_acq_rel) == 0) break;  // first: sibling will finish
            const NodeHalf& o = bvh.nodes[node].c[side ^ 1u];
            for (int k = 0; k < 3; ++k) { b.lo[k] = fminf(b.lo[k], o.lo[k]); b.hi[k] = fmaxf(b.hi[k], o.hi[k]); }
            int64_t j = other_end[node];
            int64_t first = std::min<int64_t>(node, j), last = std::max<int64_t>(node, j);
            uint32_t count = (uint32_t)(last - first + 1);
            if ((int)count <= leaf_max) { ref = leaf_ref((uint32_t)first, count); height = 0; }
            else { ref = (int32_t)node; height = std::max(height, o.height) + 1; }
            if (node == 0) { root_box = b; root_height = height; root_ref = ref; break; }
            p = parent_of_node[node];
        }
    }
    bvh.root = root_ref; bvh.max_depth = root_height; bvh.bounds = root_box;
}

}  // namespace

struct orc_blas {
    std::vector<Tri> tris;          // build order: geometry-major, primitive-minor
    std::vector<Tri> sorted_tris;   // Morton order (BVH mode)
    uint32_t n_geoms = 0;
    Box bounds = empty_box();
    bool has_bvh = false;
    Lbvh bvh;
    std::vector<uint64_t> build_keys;   // Morton keys (build order) of the last FULL build: a refit re-uses them, i.e. keeps the topology
};

struct OInst {
    float o2w[12], w2o[12];
    uint32_t custom, mask, sbt, flags;
    const orc_blas* blas;
    bool active;
    Box wbox;
};
struct orc_tlas {
    std::vector<OInst> inst;
    bool has_bvh = false;
    Lbvh bvh;
    Box bounds = empty_box();
};

namespace {

static void gather_tris(const orc_geometry* geoms, uint32_t n_geoms, std::vector<Tri>& tris, Box& bounds) {
    uint64_t total = 0;
    std::vector<uint64_t> offs(n_geoms + 1, 0);
    for (uint32_t g = 0; g < n_geoms; ++g) { offs[g] = total; total += geoms[g].triangle_count; }
    offs[n_geoms] = total;
    tris.resize(total);
    for (uint32_t g = 0; g < n_geoms; ++g) {
        const orc_geometry& G = geoms[g];
        const uint32_t stride = G.vertex_stride_bytes / 4u;
        const int64_t nt = G.triangle_count;
#pragma omp parallel for schedule(static) if (nt > 65536)
        for (int64_t p = 0; p < nt; ++p) {
            uint32_t idx[3];
            for (int k = 0; k < 3; ++k) idx[k] = G.indices ? G.indices[3 * p + k] : (uint32_t)(3 * p + k);
            V3 v[3];
            for (int k = 0; k < 3; ++k) {
                const float* src = G.vertices + (size_t)idx[k] * stride;
                v[k] = {src[0], src[1], src[2]};
                if (G.transform3x4) v[k] = xform_point(G.transform3x4, v[k]);   // baked at build time
            }
            tris[offs[g] + p] = {v[0], v[1], v[2], g, (uint32_t)p, G.flags & 0xFFu};
        }
    }
    Box b = empty_box();
    const int64_t N = (int64_t)total;
#pragma omp parallel
    {
        Box lb = empty_box();
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < N; ++i) { grow(lb, tris[i].v0); grow(lb, tris[i].v1); grow(lb, tris[i].v2); }
#pragma omp critical
        grow(b, lb);
    }
    bounds = b;
}

static void blas_build_bvh(orc_blas* B) {
    const int64_t N = (int64_t)B->tris.size();
    std::vector<Box> pb(N); std::vector<uint64_t> keys(N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) { pb[i] = tri_box(B->tris[i]); keys[i] = morton30(pb[i], B->bounds); }
    lbvh_build(B->bvh, pb, keys, LEAF_MAX);
    B->build_keys = keys;
    B->sorted_tris.resize(N);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) B->sorted_tris[i] = B->tris[B->bvh.prims[i]];
    B->has_bvh = true;
}

// ------------------------------------------------------------------------------------------
// watertight ray/triangle test (Woop, Benthin, Wald 2013), no culling
// ------------------------------------------------------------------------------------------
struct RayPre { V3 o, d; int kx, ky, kz; float Sx, Sy, Sz; };
static inline RayPre ray_pre(V3 o, V3 d) {
    RayPre r; r.o = o; r.d = d;
    int kz = 0; float m = fabsf(d.x);
    if (fabsf(d.y) > m) { kz = 1; m = fabsf(d.y); }
    if (fabsf(d.z) > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (comp(d, kz) < 0.0f) std::swap(kx, ky);
    r.kx = kx; r.ky = ky; r.kz = kz;
    r.Sx = comp(d, kx) / comp(d, kz);
    r.Sy = comp(d, ky) / comp(d, kz);
    r.Sz = 1.0f / comp(d, kz);
    return r;
}
// returns true when the ray's line crosses the triangle; t,u,v as Vulkan defines them
static inline bool woop(const RayPre& r, const Tri& tr, float& t, float& bu, float& bv, float& bw0) {
    V3 A = sub(tr.v0, r.o), B = sub(tr.v1, r.o), C = sub(tr.v2, r.o);
    float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    float Ax = comp(A, r.kx) - r.Sx * Akz, Ay = comp(A, r.ky) - r.Sy * Akz;
    float Bx = comp(B, r.kx) - r.Sx * Bkz, By = comp(B, r.ky) - r.Sy * Bkz;
    float Cx = comp(C, r.kx) - r.Sx * Ckz, Cy = comp(C, r.ky) - r.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = (U + V) + W;
    if (det == 0.0f) return false;
    float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    float T = (U * Az + V * Bz) + W * Cz;
    float rcp = 1.0f / det;
    // "+ 0.0f" canonicalises -0 to +0: the sign of an exactly-zero t/u/v depends on the kx/ky swap above,
    // which has no meaning without face culling (a faster implementation may omit the swap).
    t = T * rcp + 0.0f; bu = V * rcp + 0.0f; bv = W * rcp + 0.0f; bw0 = U * rcp + 0.0f;
    return true;
}

struct Best {
    float t; uint32_t inst, geo, prim; float u, v, w0; const OInst* I; const Tri* tri;
};
static inline bool id_less(uint32_t i0, uint32_t g0, uint32_t p0, uint32_t i1, uint32_t g1, uint32_t p1) {
    if (i0 != i1) return i0 < i1;
    if (g0 != g1) return g0 < g1;
    return p0 < p1;
}
// Ray flag semantics [spec: GL_EXT_ray_tracing / Vulkan "Ray Intersection Culling"], stated once for both sides:
//  * opacity of a candidate = geometry OPAQUE bit, overridden by the instance FORCE_OPAQUE / FORCE_NO_OPAQUE flags,
//    overridden by the ray Opaque / NoOpaque flags; CullOpaque / CullNoOpaque then drop it. There is no any-hit
//    shader in the sample (main.cpp:1199-1216 builds raygen + miss + closest-hit only), so without any-hit records a surviving
//    non-opaque candidate is accepted like an opaque one.
//  * facing is decided in object space: front = vertices clockwise as seen from the ray origin, i.e.
//    ((v1-v0) x (v2-v0)) . d > 0, inverted by the instance FLIP_FACING flag; CullBack/CullFront are ignored for
//    instances with TRIANGLE_FACING_CULL_DISABLE (what the sample sets, main.cpp:852).
//  * any-hit stage (rt_oracle.h, orc_anyhit_record): runs for non-opaque survivors that are not farther than the committed closest hit.
struct RayCfg { uint32_t ray_flags; const orc_anyhit_record* anyhit; uint32_t n_anyhit, sbt_stride, sbt_offset; };
// the alpha-test any-hit shader: true = accept, false = ignoreIntersectionEXT
static inline bool anyhit_accepts(const orc_anyhit_record& a, float u, float v) {
    if (a.kind != ORC_ANYHIT_ALPHA_MASK || !a.mask) return true;
    const uint32_t res = 1u << a.log2_res;
    uint32_t cu = (uint32_t)(int)(u * (float)res), cv = (uint32_t)(int)(v * (float)res);
    if (cu > res - 1u) cu = res - 1u;
    if (cv > res - 1u) cv = res - 1u;
    const uint32_t bit = cv * res + cu;
    return ((a.mask[bit >> 5] >> (bit & 31u)) & 1u) != 0u;
}
// Returns 0 when the candidate was rejected, 1 when it was ACCEPTED, 2 when it was accepted and ends the ray (TerminateOnFirstHit, or
// terminateRayEXT of an any-hit record).
static inline int consider(Best& best, const RayPre& r, const Tri& tr, uint32_t inst_id, const OInst* I,
                           float tmin, float tmax, const RayCfg& rc) {
    const uint32_t ray_flags = rc.ray_flags;
    float t, u, v, w0;
    if (!woop(r, tr, t, u, v, w0)) return 0;
    if (!(t > tmin && t < tmax)) return 0;      // tmin < t < tmax, both exclusive
    bool opaque = (tr.flags & ORC_GEOM_OPAQUE) != 0u;
    if (I->flags & ORC_INST_FORCE_OPAQUE) opaque = true;
    else if (I->flags & ORC_INST_FORCE_NO_OPAQUE) opaque = false;
    if (ray_flags & ORC_RAY_OPAQUE) opaque = true;
    else if (ray_flags & ORC_RAY_NO_OPAQUE) opaque = false;
    if (opaque ? (ray_flags & ORC_RAY_CULL_OPAQUE) : (ray_flags & ORC_RAY_CULL_NO_OPAQUE)) return 0;
    if ((ray_flags & (ORC_RAY_CULL_BACK | ORC_RAY_CULL_FRONT)) && !(I->flags & ORC_INST_FACING_CULL_DISABLE)) {
        V3 n = cross3(sub(tr.v1, tr.v0), sub(tr.v2, tr.v0));
        bool front = (dot3(n, r.d) > 0.0f) != ((I->flags & ORC_INST_FLIP_FACING) != 0u);
        if (front ? (ray_flags & ORC_RAY_CULL_FRONT) : (ray_flags & ORC_RAY_CULL_BACK)) return 0;
    }
    int accepted = (ray_flags & ORC_RAY_TERMINATE_ON_FIRST_HIT) ? 2 : 1;
    if (!opaque && rc.n_anyhit) {
        if (t > best.t) return 0;               // outside the current ray interval: the any-hit shader is not invoked
        const uint64_t rec = (uint64_t)I->sbt + (uint64_t)tr.geo * rc.sbt_stride + rc.sbt_offset;
        if (rec < rc.n_anyhit) {
            const orc_anyhit_record& a = rc.anyhit[rec];
            if (!anyhit_accepts(a, u, v)) return 0;
            if (a.flags & ORC_ANYHIT_TERMINATE) accepted = 2;
        }
    }
    bool better = t < best.t || (t == best.t && id_less(inst_id, tr.geo, tr.prim, best.inst, best.geo, best.prim));
    if (!better) return accepted;
    best.t = t; best.inst = inst_id; best.geo = tr.geo; best.prim = tr.prim; best.u = u; best.v = v; best.w0 = w0;
    best.I = I; best.tri = &tr;
    return accepted;
}

// conservative slab test state for one (ray, space)
struct Slab { float rd[3], on[3], of[3]; bool pos[3]; };
static inline Slab slab_pre(V3 o, V3 d, const Box& root) {
    Slab s;
    float M = 0.0f;
    for (int k = 0; k < 3; ++k) {
        float ok = comp(o, k);
        M = fmaxf(M, fmaxf(fabsf(root.lo[k]), fabsf(root.hi[k])) + fabsf(ok));
    }
    float e = M * 1.52587890625e-05f;  // 2^-16: generous spatial pad (the oracle need not be fast)
    for (int k = 0; k < 3; ++k) {
        float dk = comp(d, k), ok = comp(o, k);
        if (fabsf(dk) < 1e-20f) dk = copysignf(1e-20f, dk);
        s.pos[k] = dk > 0.0f;
        s.rd[k] = 1.0f / dk;
        s.on[k] = s.pos[k] ? ok + e : ok - e;
        s.of[k] = s.pos[k] ? ok - e : ok + e;
    }
    return s;
}
static inline bool slab_hit(const Slab& s, const float* lo, const float* hi, float tmin, float tbest, float& tn_out) {
    float tn = tmin, tf = tbest;
    for (int k = 0; k < 3; ++k) {
        float pn = s.pos[k] ? lo[k] : hi[k], pf = s.pos[k] ? hi[k] : lo[k];
        float a = (pn - s.on[k]) * s.rd[k];
        float b = (pf - s.of[k]) * s.rd[k];
        a = a > 0.0f ? a * 0.99999905f : a * 1.00000095f;
        b = b > 0.0f ? b * 1.00000095f : b * 0.99999905f;
        tn = fmaxf(tn, a); tf = fminf(tf, b);
    }
    tn_out = tn;
    return tn <= tf;
}

struct Counters { uint64_t nodes = 0, tris = 0, insts = 0; };

// returns true when the ray was terminated (TerminateOnFirstHit and a candidate accepted)
static bool traverse_blas(const orc_blas* B, const RayPre& r, uint32_t inst_id, const OInst* I, float tmin, float tmax,
                          const RayCfg& rc, Best& best, Counters& cnt) {
    const Lbvh& bvh = B->bvh;
    if (bvh.root == REF_EMPTY) return false;
    Slab s = slab_pre(r.o, r.d, B->bounds);
    int32_t stack[192]; int sp = 0;
    int32_t cur = bvh.root;
    for (;;) {
        if (ref_is_leaf(cur)) {
            uint32_t f = leaf_first(cur), c = leaf_count(cur);
            for (uint32_t i = 0; i < c; ++i) {
                ++cnt.tris;
                if (consider(best, r, B->sorted_tris[f + i], inst_id, I, tmin, tmax, rc) == 2) return true;
            }
            if (sp == 0) break;
            cur = stack[--sp];
            continue;
        }
        const Node& n = bvh.nodes[cur]; ++cnt.nodes;
        float t0, t1;
        bool h0 = slab_hit(s, n.c[0].lo, n.c[0].hi, tmin, best.t, t0);
        bool h1 = slab_hit(s, n.c[1].lo, n.c[1].hi, tmin, best.t, t1);
        if (h0 && h1) {
            int nearc = t1 < t0 ? 1 : 0;
            stack[sp++] = n.c[nearc ^ 1].ref;
            cur = n.c[nearc].ref;
        } else if (h0) cur = n.c[0].ref;
        else if (h1) cur = n.c[1].ref;
        else { if (sp == 0) break; cur = stack[--sp]; }
    }
    return false;
}

struct TraceCtx {
    const orc_tlas* T; orc_ray_params rp; const orc_shader_data* sd; int mode;
};

static inline bool enter_instance(const TraceCtx& c, uint32_t i, V3 o, V3 d, float tmin, float tmax, Best& best, Counters& cnt) {
    const OInst& I = c.T->inst[i];
    if (!I.active) return false;
    if ((I.mask & c.rp.cull_mask) == 0) return false;
    const RayCfg rc = {c.rp.ray_flags, c.sd->anyhit_records, c.sd->anyhit_records ? c.sd->anyhit_record_count : 0u, c.rp.sbt_record_stride, c.rp.sbt_record_offset};
    ++cnt.insts;
    V3 oo = xform_point(I.w2o, o), od = xform_vec(I.w2o, d);
    RayPre r = ray_pre(oo, od);
    if (c.mode == ORC_MODE_BRUTE || !I.blas->has_bvh) {
        for (const Tri& tr : I.blas->tris) {
            ++cnt.tris;
            if (consider(best, r, tr, i, &I, tmin, tmax, rc) == 2) return true;
        }
        return false;
    }
    return traverse_blas(I.blas, r, i, &I, tmin, tmax, rc, best, cnt);
}

static Best trace_ray(const TraceCtx& c, V3 o, V3 d, float tmin, float tmax, Counters& cnt) {
    Best best; best.t = tmax; best.inst = best.geo = best.prim = 0xFFFFFFFFu; best.u = best.v = best.w0 = 0.0f;
    best.I = nullptr; best.tri = nullptr;
    const orc_tlas* T = c.T;
    if (c.mode == ORC_MODE_BRUTE || !T->has_bvh) {
        for (uint32_t i = 0; i < (uint32_t)T->inst.size(); ++i)
            if (enter_instance(c, i, o, d, tmin, tmax, best, cnt)) break;
        return best;
    }
    const Lbvh& bvh = T->bvh;
    if (bvh.root == REF_EMPTY) return best;
    Slab s = slab_pre(o, d, T->bounds);
    int32_t stack[192]; int sp = 0;
    int32_t cur = bvh.root;
    for (;;) {
        if (ref_is_leaf(cur)) {
            uint32_t f = leaf_first(cur), n = leaf_count(cur);
            bool done = false;
            for (uint32_t k = 0; k < n && !done; ++k) done = enter_instance(c, bvh.prims[f + k], o, d, tmin, tmax, best, cnt);
            if (done || sp == 0) break;
            cur = stack[--sp];
            continue;
        }
        const Node& n = bvh.nodes[cur]; ++cnt.nodes;
        float t0, t1;
        bool h0 = slab_hit(s, n.c[0].lo, n.c[0].hi, tmin, best.t, t0);
        bool h1 = slab_hit(s, n.c[1].lo, n.c[1].hi, tmin, best.t, t1);
        if (h0 && h1) {
            int nearc = t1 < t0 ? 1 : 0;
            stack[sp++] = n.c[nearc ^ 1].ref;
            cur = n.c[nearc].ref;
        } else if (h0) cur = n.c[0].ref;
        else if (h1) cur = n.c[1].ref;
        else { if (sp == 0) break; cur = stack[--sp]; }
    }
    return best;
}

// closest-hit shader, main.cpp:1080-1091; SBT rule main.cpp:1260-1262
static inline V3 closest_hit(const TraceCtx& c, const Best& b) {
    if (b.prim == 1u && b.inst == 1u && b.I->custom == 100u && b.geo == 1u)
        return {1.0f - b.u - b.v, b.u, b.v};
    uint32_t rec = b.I->sbt + b.geo * c.rp.sbt_record_stride + c.rp.sbt_record_offset;
    if (rec >= c.sd->hit_record_count) return {0.0f, 0.0f, 0.0f};
    const float* col = c.sd->hit_records_rgb + 3 * (size_t)rec;
    return {col[0], col[1], col[2]};
}
static inline uint8_t unorm8(float c) {
    float v = c;
    if (!(v > 0.0f)) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return (uint8_t)(int)nearbyintf(v * 255.0f);   // RNE (default rounding mode)
}
static inline void write_hit(orc_hit* out, const Best& b, float tmax) {
    if (b.I) { out->instance_id = b.inst; out->geometry_index = b.geo; out->primitive_id = b.prim; out->custom_index = b.I->custom; out->t = b.t; out->u = b.u; out->v = b.v; }
    else { out->instance_id = out->geometry_index = out->primitive_id = out->custom_index = 0xFFFFFFFFu; out->t = tmax; out->u = 0.0f; out->v = 0.0f; }
}

// Deterministic diffuse bounce (OUR definition; the reference's recursion depth is 1, main.cpp:1168).
// Only + - * / sqrt so that CPU and GPU agree bit for bit.
static inline float u01(uint32_t h) { return (float)(h >> 8) * 5.9604644775390625e-08f; }  // 2^-24
static bool bounce_ray(const Best& b, V3 o, V3 d, uint32_t pixel, uint32_t seed, V3& o2, V3& d2) {
    V3 p = {o.x + b.t * d.x, o.y + b.t * d.y, o.z + b.t * d.z};
    V3 e1 = sub(b.tri->v1, b.tri->v0), e2 = sub(b.tri->v2, b.tri->v0);
    V3 n = xform_normal(b.I->w2o, cross3(e1, e2));
    float l2 = dot3(n, n);
    if (l2 > 0.0f && l2 < INFINITY) {
        float l = sqrtf(l2);
        n = {n.x / l, n.y / l, n.z / l};
    } else {
        float dl = sqrtf(dot3(d, d));
        n = {-d.x / dl, -d.y / dl, -d.z / dl};
    }
    if (dot3(n, d) > 0.0f) n = {-n.x, -n.y, -n.z};
    uint32_t h = pcg_hash(pixel + pcg_hash(seed + 0x9E3779B9u));
    V3 s = {0.0f, 0.0f, 0.0f};
    for (int tries = 0; tries < 8; ++tries) {
        uint32_t a = pcg_hash(h), bb = pcg_hash(a), cc = pcg_hash(bb);
        h = cc;
        V3 q = {u01(a) * 2.0f - 1.0f, u01(bb) * 2.0f - 1.0f, u01(cc) * 2.0f - 1.0f};
        float q2 = dot3(q, q);
        if (q2 <= 1.0f && q2 > 1e-8f) {
            float ql = sqrtf(q2);
            s = {q.x / ql, q.y / ql, q.z / ql};
            break;
        }
    }
    V3 dir = {n.x + s.x, n.y + s.y, n.z + s.z};
    float dl2 = dot3(dir, dir);
    if (dl2 < 1e-12f) dir = n;
    else { float dl = sqrtf(dl2); dir = {dir.x / dl, dir.y / dl, dir.z / dl}; }
    const float eps = 0.0009765625f;  // 2^-10
    o2 = {p.x + n.x * eps, p.y + n.y * eps, p.z + n.z * eps};
    d2 = dir;
    return true;
}

}  // namespace

extern "C" {

orc_blas* orc_build_blas(const orc_geometry* geoms, uint32_t n_geoms, int build_bvh, int /*key_bits*/) {
    orc_blas* B = new orc_blas();
    B->n_geoms = n_geoms;
    gather_tris(geoms, n_geoms, B->tris, B->bounds);
    if (build_bvh) blas_build_bvh(B);
    return B;
}
void orc_free_blas(orc_blas* b) { delete b; }

// VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR as the product's RT_BUILD_MODE_REFIT defines it: same geometry / triangle counts, new
// vertices; the sorted order (hence Karras' tree) of the last full build is kept, triangle records and every node box are recomputed.
int orc_refit_blas(orc_blas* B, const orc_geometry* geoms, uint32_t n_geoms) {
    if (!B || !B->has_bvh || n_geoms != B->n_geoms) return -1;
    std::vector<Tri> tris; Box bounds;
    gather_tris(geoms, n_geoms, tris, bounds);
    if (tris.size() != B->tris.size()) return -1;
    B->tris.swap(tris); B->bounds = bounds;
    const int64_t N = (int64_t)B->tris.size();
    std::vector<Box> pb(N);
    for (int64_t i = 0; i < N; ++i) pb[i] = tri_box(B->tris[i]);
    lbvh_build(B->bvh, pb, B->build_keys, LEAF_MAX);             // old keys -> old order and topology; new boxes
    for (int64_t i = 0; i < N; ++i) B->sorted_tris[i] = B->tris[B->bvh.prims[i]];
    return 0;
}

orc_tlas* orc_build_tlas(const orc_instance* inst, uint32_t n, int build_bvh) {
    orc_tlas* T = new orc_tlas();
    T->inst.resize(n);
    Box all = empty_box();
    for (uint32_t i = 0; i < n; ++i) {
        OInst& I = T->inst[i];
        memcpy(I.o2w, inst[i].transform, sizeof(I.o2w));
        I.custom = inst[i].custom_index_and_mask & 0xFFFFFFu; I.mask = inst[i].custom_index_and_mask >> 24;
        I.sbt = inst[i].sbt_offset_and_flags & 0xFFFFFFu; I.flags = inst[i].sbt_offset_and_flags >> 24;
        I.blas = inst[i].blas;
        bool ok = invert3x4(I.o2w, I.w2o);
        I.active = ok && I.blas && !I.blas->tris.empty();
        I.wbox = empty_box();
        if (I.active) {
            const Box& bb = I.blas->bounds;
            for (int c = 0; c < 8; ++c) {
                V3 p = {(c & 1) ? bb.hi[0] : bb.lo[0], (c & 2) ? bb.hi[1] : bb.lo[1], (c & 4) ? bb.hi[2] : bb.lo[2]};
                grow(I.wbox, xform_point(I.o2w, p));
            }
            grow(all, I.wbox);
        }
    }
    T->bounds = all;
    if (build_bvh) {
        std::vector<Box> pb(n); std::vector<uint64_t> keys(n);
        for (uint32_t i = 0; i < n; ++i) {
            pb[i] = T->inst[i].wbox;
            keys[i] = T->inst[i].active ? morton30(pb[i], all) : 0x3FFFFFFFu;
        }
        lbvh_build(T->bvh, pb, keys, 1);
        T->has_bvh = true;
    }
    return T;
}
void orc_free_tlas(orc_tlas* t) { delete t; }

float orc_aspect_y(float yfov_deg) { return tanf((yfov_deg * 0.017453292519943295f) * 0.5f); }
uint32_t orc_pcg_hash(uint32_t v) { return pcg_hash(v); }
uint32_t orc_morton30(float x, float y, float z, const float lo[3], const float hi[3]) {
    Box pb = {{x, y, z}, {x, y, z}}; Box sc = {{lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
    return morton30(pb, sc);
}
void orc_unorm8(const float rgb[3], uint8_t out[4]) { out[0] = unorm8(rgb[0]); out[1] = unorm8(rgb[1]); out[2] = unorm8(rgb[2]); out[3] = 0; }
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_trace(const orc_tlas* tlas, const orc_camera* cam, const orc_ray_params* rp_in, const orc_shader_data* sd,
              uint32_t width, uint32_t height, uint32_t bounces, int mode,
              uint32_t row_begin, uint32_t row_end, uint32_t row_step,
              uint8_t* rgba_out, orc_hit* primary_out, orc_hit* secondary_out, orc_stats* stats_out) {
    if (!tlas || !cam || !sd || width == 0 || height == 0 || row_step == 0) return -1;
    TraceCtx c; c.T = tlas; c.sd = sd; c.mode = mode;
    if (rp_in) c.rp = *rp_in; else c.rp = {0.0f, 100.0f, 0xffu, 0u, 1u, 1u, ORC_RAY_OPAQUE, 0u};
    // miss shader table: each miss shader, like the sample's (main.cpp:1063-1066), writes one constant colour
    const float* miss_rgb = sd->miss_rgb;
    if (sd->miss_records_rgb) { if (c.rp.miss_index >= sd->miss_record_count) return -5; miss_rgb = sd->miss_records_rgb + 3 * (size_t)c.rp.miss_index; }
    else if (c.rp.miss_index != 0u) return -5;
    const bool skip_chit = (c.rp.ray_flags & ORC_RAY_SKIP_CLOSEST_HIT) != 0u;   // payload stays (0,0,0) on a hit (main.cpp:1045)
    // SBT range pre-check (the product reports RT_ERROR_SBT_RANGE for the same condition)
    for (const OInst& I : tlas->inst) {
        if (!I.blas) continue;
        uint32_t ng = I.blas->n_geoms ? I.blas->n_geoms : 1u;
        uint64_t last = (uint64_t)I.sbt + (uint64_t)(ng - 1) * c.rp.sbt_record_stride + c.rp.sbt_record_offset;
        if (last >= sd->hit_record_count) return -5;
    }
    // raygen constants, main.cpp:1038-1039 (tan hoisted to the host on both sides)
    const float aspect_y = orc_aspect_y(cam->yfov_deg);
    const float aspect_x = aspect_y * (float)width / (float)height;
    const V3 cam_o = {cam->pos[0], cam->pos[1], cam->pos[2]};
    if (row_end > height) row_end = height;
    uint64_t s_nodes = 0, s_tris = 0, s_insts = 0, s_ph = 0, s_sh = 0, s_sec = 0, s_edge = 0, s_prim = 0;
    const int64_t nrows = row_end > row_begin ? (int64_t)((row_end - row_begin + row_step - 1) / row_step) : 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : s_nodes, s_tris, s_insts, s_ph, s_sh, s_sec, s_edge, s_prim)
    for (int64_t ri = 0; ri < nrows; ++ri) {
        const uint32_t py = row_begin + (uint32_t)ri * row_step;
        Counters cnt;
        for (uint32_t px = 0; px < width; ++px) {
            const uint32_t pixel = py * width + px;
            // raygen, main.cpp:1033-1055
            float scx = (float)px + 0.5f, scy = (float)py + 0.5f;
            float ndcx = scx / (float)width * 2.0f - 1.0f;
            float ndcy = scy / (float)height * 2.0f - 1.0f;
            float ax = ndcx * aspect_x, ay = ndcy * aspect_y;
            V3 d = {(ax * 1.0f + ay * 0.0f) + 0.0f, (ax * 0.0f + ay * -1.0f) + 0.0f, (ax * 0.0f + ay * 0.0f) + -1.0f};
            V3 hit_value = {0.0f, 0.0f, 0.0f};
            Best b = trace_ray(c, cam_o, d, c.rp.tmin, c.rp.tmax, cnt);
            ++s_prim;
            if (primary_out) write_hit(primary_out + pixel, b, c.rp.tmax);
            Best b2; b2.I = nullptr; b2.t = c.rp.tmax; b2.inst = b2.geo = b2.prim = 0xFFFFFFFFu; b2.u = b2.v = 0.0f;
            if (b.I) {
                ++s_ph;
                if (fminf(fminf(b.u, b.v), b.w0) < 9.5367431640625e-07f) ++s_edge;
                if (!skip_chit) hit_value = closest_hit(c, b);
                if (bounces > 0 && !skip_chit) {
                    V3 o2, d2;
                    bounce_ray(b, cam_o, d, pixel, c.rp.bounce_seed, o2, d2);
                    b2 = trace_ray(c, o2, d2, c.rp.tmin, c.rp.tmax, cnt);
                    ++s_sec;
                    V3 sc;
                    if (b2.I) { ++s_sh; sc = closest_hit(c, b2); }
                    else sc = {miss_rgb[0], miss_rgb[1], miss_rgb[2]};
                    hit_value = {0.5f * hit_value.x + 0.5f * sc.x, 0.5f * hit_value.y + 0.5f * sc.y, 0.5f * hit_value.z + 0.5f * sc.z};
                }
            } else {
                hit_value = {miss_rgb[0], miss_rgb[1], miss_rgb[2]};   // miss shader, main.cpp:1063-1066
            }
            if (secondary_out) write_hit(secondary_out + pixel, b2, c.rp.tmax);
            if (rgba_out) {   // imageStore(image, xy, vec4(hitValue, 0.0)) into rgba8, main.cpp:1054
                uint8_t* o = rgba_out + 4 * (size_t)pixel;
                o[0] = unorm8(hit_value.x); o[1] = unorm8(hit_value.y); o[2] = unorm8(hit_value.z); o[3] = 0;
            }
        }
        s_nodes += cnt.nodes; s_tris += cnt.tris; s_insts += cnt.insts;
    }
    if (stats_out) {
        stats_out->rays_primary = s_prim; stats_out->rays_secondary = s_sec; stats_out->nodes_visited = s_nodes;
        stats_out->triangles_tested = s_tris; stats_out->instances_entered = s_insts;
        stats_out->primary_hits = s_ph; stats_out->secondary_hits = s_sh; stats_out->near_edge_hits = s_edge;
    }
    return 0;
}

int orc_blas_get_info(const orc_blas* b, orc_blas_info* out) {
    if (!b || !out) return -1;
    out->triangle_count = (uint32_t)b->tris.size();
    out->node_count = (uint32_t)b->bvh.nodes.size();
    out->root_ref = b->has_bvh ? b->bvh.root : REF_EMPTY;
    out->max_depth = b->bvh.max_depth;
    for (int k = 0; k < 3; ++k) { out->bounds_lo[k] = b->bounds.lo[k]; out->bounds_hi[k] = b->bounds.hi[k]; }
    return 0;
}

int orc_blas_export(const orc_blas* b, void* nodes_out, void* tris_out, uint64_t* keys_out, uint32_t* prims_out) {
    if (!b || !b->has_bvh) return -1;
    if (nodes_out && !b->bvh.nodes.empty()) memcpy(nodes_out, b->bvh.nodes.data(), b->bvh.nodes.size() * sizeof(Node));
    if (tris_out) {
        // product triangle layout: 9 floats + geo + prim + pad = 48 B
        uint8_t* dst = (uint8_t*)tris_out;
        for (size_t i = 0; i < b->sorted_tris.size(); ++i) {
            const Tri& t = b->sorted_tris[i];
            float f[9] = {t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z};
            uint32_t ids[3] = {t.geo, t.prim, 0u};
            memcpy(dst + 48 * i, f, 36); memcpy(dst + 48 * i + 36, ids, 12);
        }
    }
    if (keys_out) memcpy(keys_out, b->bvh.keys.data(), b->bvh.keys.size() * sizeof(uint64_t));
    if (prims_out) memcpy(prims_out, b->bvh.prims.data(), b->bvh.prims.size() * sizeof(uint32_t));
    return 0;
}

double orc_time_blas_build(const orc_geometry* geoms, uint32_t n_geoms, int key_bits) {
    auto t0 = std::chrono::steady_clock::now();
    orc_blas* B = orc_build_blas(geoms, n_geoms, 1, key_bits);
    auto t1 = std::chrono::steady_clock::now();
    orc_free_blas(B);
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
