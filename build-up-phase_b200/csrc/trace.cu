// trace.cu — vkCmdTraceRaysKHR(W,H,1) for sm_100a (reference dispatch: main.cpp:1349-1355).
//
// One call = raygen prologue (main.cpp:1033-1052) -> two-level TLAS->BLAS while-while traversal
// (what traceRayEXT hands to the driver / RT cores; B200 has none, so this runs on the SMs) ->
// closest-hit / miss epilogue (main.cpp:1063-1066,1080-1091) -> rgba8 imageStore (main.cpp:1054).
// With one diffuse bounce the work is a two-stage wavefront: stage 0 traces the primary rays and
// appends a secondary ray per hit to a queue in HBM; stage 1 traces the queue and blends.
//
//  * Persistent warps: the grid is sized to the SMs (occupancy x 148), every warp pulls rays from a
//    global counter. When fewer than REFILL_THRESHOLD lanes of a warp are still traversing, the
//    warp leaves the traversal loop, finished lanes run their epilogue and the idle lanes are
//    refilled with one ballot + one atomicAdd + one shuffle (warp-level ray compaction), so SIMD
//    lanes stay busy although ray lengths differ by orders of magnitude.
//  * Rays are numbered tile-major (8x4-pixel tiles) so the 32 rays a warp fetches together are
//    neighbours and concurrently running warps work on neighbouring tiles (L1/L2 reuse of nodes).
//  * 64-byte nodes fetched as 4 x LDG.128 through the read-only path; 48-byte triangles as 3 x LDG.128.
//  * Box test: conservative slabs in FMA form (pad derived per ray/space from |origin| + |bounds|),
//    so a box is never culled when the exact-arithmetic triangle test could still report a hit.
//  * Triangle test: watertight Woop/Benthin/Wald 2013, plain IEEE mul/add (no contraction; the
//    file is compiled with -fmad=false), fp64 fallback on exact-zero edge functions. No culling:
//    the sample uses gl_RayFlagsOpaqueEXT only and TRIANGLE_FACING_CULL_DISABLE (main.cpp:852,1048).
//  * Closest hit = smallest t in (tmin, tmax); equal t resolved by lowest (instance, geometry,
//    primitive) so the result does not depend on BVH shape or traversal order.
//  * GENERAL kernel variant (selected only when the ray flags ask for more than Opaque/NoOpaque, so the sample's
//    path pays nothing): opacity resolution geometry -> instance FORCE_* -> ray flags with CULL_OPAQUE/CULL_NO_OPAQUE,
//    front/back-face culling in object space with the instance FLIP_FACING / FACING_CULL_DISABLE flags,
//    TERMINATE_ON_FIRST_HIT and SKIP_CLOSEST_HIT_SHADER (include/rtcore.h, RT_RAY_FLAG_*).
#include <float.h>

#include "rt_device.cuh"

namespace rt {

namespace {

#ifndef RT_TRACE_MIN_BLOCKS
#define RT_TRACE_MIN_BLOCKS 9
#endif
#ifndef RT_REFILL_THRESHOLD
#define RT_REFILL_THRESHOLD 12
#endif
constexpr int TRACE_THREADS = 128;
constexpr int TRACE_MIN_BLOCKS = RT_TRACE_MIN_BLOCKS;     // register cap 65536 / (128 * 9) = 56 (with the min/max slab test; 8 CTAs = 64 registers before it)
// BIG kernel variant: ONE 1024-thread CTA per SM (same 32 warps, same 64-register cap) whose dynamic shared memory holds a copy of the
// TLAS nodes, staged once per CTA with a TMA bulk copy (cp.async.bulk + mbarrier). Top-level node fetches - every ray starts there, and
// the rays that miss everything never leave it - then are LDS.128 instead of divergent L1 tag look-ups, and the TLAS stops competing
// with the BLAS nodes for L1 lines. Selected when the TLAS has RT_SMEM_TLAS_MIN_NODES..RT_SMEM_TLAS_MAX_NODES nodes.
constexpr int TRACE_THREADS_BIG = 1024;
// Measured on B200 (inst10m 4K + bounce, 1023 TLAS nodes = 64 KB; profiles/README.md r2_g, prof_trace_r2_g.json): 3281 Mrays/s WITH the
// staged TLAS vs 3499 without. ncu: l1tex throughput 75 % -> 81 % (stage 0) and 71 % -> 82 % (stage 1) - shared-memory reads go through
// the same L1TEX data pipeline, and the generic-address loads (LD.E.128 instead of LDG.E.128.CONSTANT) that let one loop walk both
// spaces cost 3.5 % more warp-instructions and the BLAS nodes their read-only path. Default 0; the variant stays as a build option.
#ifndef RT_SMEM_TLAS
#define RT_SMEM_TLAS 0
#endif
#ifndef RT_SMEM_TLAS_MAX_NODES
#define RT_SMEM_TLAS_MAX_NODES 1024      // 64 KB of the SM's 256 KB L1/shared array
#endif
#ifndef RT_SMEM_TLAS_MIN_NODES
#define RT_SMEM_TLAS_MIN_NODES 15
#endif
constexpr uint32_t SMEM_TLAS_HEADER = 128;    // mbarrier (8 B) + padding: the node copy is 128-byte aligned
constexpr int REFILL_THRESHOLD = RT_REFILL_THRESHOLD;     // leave the traversal loop when fewer lanes are active
#ifndef RT_NODE_CAP
#define RT_NODE_CAP 6
#endif
constexpr int NODE_CAP = RT_NODE_CAP;                     // leave the inner node loop when fewer lanes are still in it (0 = never);
                                                          // measured 0/2/4/6/8 -> 3028/3138/3247/3273/3265 Mrays/s (profiles r02b)
#ifndef RT_BOUNCE_ORDERED
#define RT_BOUNCE_ORDERED 1
#endif
// RT_BOUNCE_ORDERED=1: stage 0 stores the bounce ray of a pixel in that pixel's slot (tile-major numbering) and sets its
// bit in a per-tile mask; two small kernels turn the masks into a compact, TILE-ORDERED index list for stage 1, so the
// 32 secondary rays a warp fetches start on neighbouring surface points. (=0: the rays are appended to a queue in the
// order the persistent warps finish them, which scatters them over the ~40 pixel rows that are in flight at a time.)
// Measured and removed (profiles/README.md r01t): listing the rays of every 64-tile group by direction octant first
// (three index kernels, directions re-read) made the frame 3 % SLOWER (3019 vs 3108 Mrays/s).
#ifndef RT_FUSED_STAGES
#define RT_FUSED_STAGES 1
#endif
// RT_FUSED_STAGES=1: a SMALL launch with a bounce is ONE persistent kernel. Warps take primary rays while there are any,
// then bounce rays from an append queue that the finishing primary rays fill, so the tail of the primary stage (its
// longest rays running alone) overlaps the bulk of the secondary stage and the dependent launches in between disappear.
// That fixed cost (~0.35 ms per frame) is what limits multi-GPU scaling. Measured on B200 (profiles/README.md r01u), fused
// vs stage 0 + index kernels + stage 1: 0.49 vs 0.58 ms at 0.5 M pixels, 0.78 vs 0.83 at 1.0 M, 1.29 vs 1.27 at 2.1 M,
// 3.92 vs 3.60 at 8.3 M (the fused kernel carries more live state through the traversal loop and pays ~18 % per ray), so
// launches of up to RT_FUSED_MAX_PIXELS pixels are fused and larger ones are not. (=0: never fuse.)
#ifndef RT_FUSED_MAX_PIXELS
#define RT_FUSED_MAX_PIXELS 1500000u
#endif
#ifndef RT_LDG256
#define RT_LDG256 0
#endif
// Ray numbering / fetch counters: RT_REGIONS (rt_internal.h). A region = REGION_TW x REGION_TH tiles of 8x4 pixels; ray ids are
// region-major (regions row-major over the launch, tiles row-major inside a region, 32 pixels per tile), padded to whole regions.
constexpr uint32_t REGION_TW = RT_REGIONS ? RT_REGION_TW : 1u, REGION_TH = RT_REGIONS ? RT_REGION_TH : 1u;
constexpr uint32_t TPR = REGION_TW * REGION_TH, RPR = TPR * 32u;    // tiles / rays per region
// RT_SMEM_STACK = S > 0: the first S entries of every lane's traversal stack live in shared memory ([entry][thread]: conflict-free),
// deeper entries in local memory; 0: the whole stack in local memory (which goes through L1 like the node fetches do).
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 0
#endif
// the node loop's "too few lanes left" test every RT_CAP_EVERY-th iteration only (it costs a vote + popc + branch per node step)
#ifndef RT_CAP_EVERY
#define RT_CAP_EVERY 1
#endif
#ifndef RT_BRANCHLESS_NODE
#define RT_BRANCHLESS_NODE 0
#endif
#ifndef RT_PRIMARY_ORIGIN_CONST
#define RT_PRIMARY_ORIGIN_CONST 1
#endif
// RT_SLAB_MINMAX=1: the slab test picks near / far with min / max instead of per-axis selects on the direction sign (bit-identical values):
// 3 instructions less per node step (no predicates to rebuild from the packed sign bytes) and three registers less in the loop, which is
// what makes a ninth CTA per SM pay. Measured on B200 (inst10m, profiles/README.md r2_u / r2_v): select form at 8 / 9 CTAs per SM 3518 / 3479
// Mrays/s, min/max form at 8 / 9 / 10 CTAs 3522 / 3603 / 3594.
#ifndef RT_SLAB_MINMAX
#define RT_SLAB_MINMAX 1
#endif
// RT_STACK_TOS=1: top of the traversal stack in a register (see push/pop)
#ifndef RT_STACK_TOS
#define RT_STACK_TOS 0
#endif
// registers of the bounce stage: its own minimum CTAs per SM (it is the more latency-bound stage)
#ifndef RT_TRACE_MIN_BLOCKS_S1
#define RT_TRACE_MIN_BLOCKS_S1 RT_TRACE_MIN_BLOCKS
#endif
#ifndef RT_FAST_SLAB
#define RT_FAST_SLAB 1
#endif

// Node-half fetch. RT_LDG256=1 uses the sm_100a 256-bit load (LDG.E.ENL2.256): measured SLOWER than two LDG.128
// on this kernel (2652 vs 2978 Mrays/s, profiles/README.md r01g), so the default is 2 x LDG.128.
struct F8 { float4 a, b; };
// experiments (profiles/README.md r2_x): RT_TRI_NOALLOC = 1 fetches triangle records with L1::no_allocate (a 48-byte record is used once per ray;
// keeping it out of L1 leaves the lines to the nodes), = 2 with L1::evict_first; RT_L1_CARVEOUT >= 0 pins the shared-memory carve-out
#ifndef RT_TRI_NOALLOC
#define RT_TRI_NOALLOC 0
#endif
#ifndef RT_L1_CARVEOUT
#define RT_L1_CARVEOUT -1
#endif
#ifndef RT_PREFETCH_LEAF
#define RT_PREFETCH_LEAF 0      // 1: prefetch.global.L1 of the leaf's first record when a lane reaches a leaf; 2: also its last byte; 3: BLAS leaves only
#endif
__device__ __forceinline__ float4 ldg_tri(const float4* p) {
#if RT_TRI_NOALLOC == 1
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#elif RT_TRI_NOALLOC == 2
    float4 r;
    asm("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return __ldg(p);
#endif
}
// GENERIC: the node may live in shared memory (the staged TLAS) or in global memory (a BLAS): generic-address loads (LD.E.128), no branch
template <bool GENERIC>
__device__ __forceinline__ F8 ldg256(const void* p) {
    F8 r;
    if (GENERIC) { r.a = *reinterpret_cast<const float4*>(p); r.b = *(reinterpret_cast<const float4*>(p) + 1); return r; }
#if RT_LDG256
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
#else
    r.a = __ldg(reinterpret_cast<const float4*>(p)); r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
#endif
    return r;
}
constexpr uint32_t NO_HIT = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Slab {
    float rdx, rdy, rdz;     // 1/d (zero components replaced by +-1e-20)
    float cnx, cny, cnz;     // -(o +- e) * rd for the near planes
    float cfx, cfy, cfz;     // -(o -+ e) * rd for the far planes
    bool px, py, pz;         // d > 0
};

__device__ __forceinline__ void slab_setup(Slab& s, V3 o, V3 d, float ax, float ay, float az) {
    const float M = fmaxf(fmaxf(ax + fabsf(o.x), ay + fabsf(o.y)), az + fabsf(o.z));
    const float e = M * 1.9073486328125e-06f;   // 2^-19 relative spatial pad
    const float dx = fabsf(d.x) < 1e-20f ? copysignf(1e-20f, d.x) : d.x;
    const float dy = fabsf(d.y) < 1e-20f ? copysignf(1e-20f, d.y) : d.y;
    const float dz = fabsf(d.z) < 1e-20f ? copysignf(1e-20f, d.z) : d.z;
    s.px = dx > 0.0f; s.py = dy > 0.0f; s.pz = dz > 0.0f;
#if RT_FAST_SLAB
    // MUFU.RCP (<= 1 ulp): the box test only has to be conservative and the 2^-19 spatial pad dwarfs a 2^-23 scale error of t
    asm("rcp.approx.f32 %0, %1;" : "=f"(s.rdx) : "f"(dx));
    asm("rcp.approx.f32 %0, %1;" : "=f"(s.rdy) : "f"(dy));
    asm("rcp.approx.f32 %0, %1;" : "=f"(s.rdz) : "f"(dz));
#else
    s.rdx = 1.0f / dx; s.rdy = 1.0f / dy; s.rdz = 1.0f / dz;
#endif
    s.cnx = -((s.px ? o.x + e : o.x - e) * s.rdx); s.cfx = -((s.px ? o.x - e : o.x + e) * s.rdx);
    s.cny = -((s.py ? o.y + e : o.y - e) * s.rdy); s.cfy = -((s.py ? o.y - e : o.y + e) * s.rdy);
    s.cnz = -((s.pz ? o.z + e : o.z - e) * s.rdz); s.cfz = -((s.pz ? o.z - e : o.z + e) * s.rdz);
#if RT_SLAB_MINMAX
    // min/max form: cn* becomes the constant that goes with the box's LO plane, cf* the one that goes with its HI plane (swapped for a
    // negative direction), so that slab_test needs no per-axis select: near = min(t_lo, t_hi), far = max(t_lo, t_hi). The padded near value
    // can never exceed the padded far value ((hi - o + e) >= (lo - o - e)), so min/max pick exactly the values the select form picks.
    if (!s.px) { const float t = s.cnx; s.cnx = s.cfx; s.cfx = t; }
    if (!s.py) { const float t = s.cny; s.cny = s.cfy; s.cfy = t; }
    if (!s.pz) { const float t = s.cnz; s.cnz = s.cfz; s.cfz = t; }
#endif
}

// half = {lo.x lo.y lo.z hi.x} {hi.y hi.z ref height}
__device__ __forceinline__ bool slab_test(const Slab& s, const float4 h0, const float4 h1, float tmin, float tbest, float& tn) {
#if RT_SLAB_MINMAX
    const float lx = __fmaf_rn(h0.x, s.rdx, s.cnx), hx = __fmaf_rn(h0.w, s.rdx, s.cfx);
    const float ly = __fmaf_rn(h0.y, s.rdy, s.cny), hy = __fmaf_rn(h1.x, s.rdy, s.cfy);
    const float lz = __fmaf_rn(h0.z, s.rdz, s.cnz), hz = __fmaf_rn(h1.y, s.rdz, s.cfz);
    tn = fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fmaxf(fminf(lz, hz), tmin));
    const float tf_ = fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fminf(fmaxf(lz, hz), tbest));
    return tn <= tf_;
#endif
    const float nx = s.px ? h0.x : h0.w, fx = s.px ? h0.w : h0.x;
    const float ny = s.py ? h0.y : h1.x, fy = s.py ? h1.x : h0.y;
    const float nz = s.pz ? h0.z : h1.y, fz = s.pz ? h1.y : h0.z;
    const float tnx = __fmaf_rn(nx, s.rdx, s.cnx), tfx = __fmaf_rn(fx, s.rdx, s.cfx);
    const float tny = __fmaf_rn(ny, s.rdy, s.cny), tfy = __fmaf_rn(fy, s.rdy, s.cfy);
    const float tnz = __fmaf_rn(nz, s.rdz, s.cnz), tfz = __fmaf_rn(fz, s.rdz, s.cfz);
    tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
    const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tbest));
    return tn <= tf;
}

struct Woop {
    float okx, oky, okz;     // origin permuted to (kx, ky, kz)
    float Sx, Sy, Sz;
    bool z0, z1;             // kz == 0, kz == 1
};
// kz = dominant axis (lowest index wins ties), kx = kz+1, ky = kz+2 (mod 3). The kx/ky swap of the
// paper (for d[kz] < 0) only negates U, V, W and det together; without face culling that leaves the
// hit decision and t, u, v unchanged except for the sign of exact zeros, which is canonicalised.
__device__ __forceinline__ void woop_setup(Woop& w, V3 o, V3 d) {
    int kz = 0; float m = fabsf(d.x);
    if (fabsf(d.y) > m) { kz = 1; m = fabsf(d.y); }
    if (fabsf(d.z) > m) { kz = 2; }
    w.z0 = kz == 0; w.z1 = kz == 1;
    const float dkx = w.z0 ? d.y : (w.z1 ? d.z : d.x);
    const float dky = w.z0 ? d.z : (w.z1 ? d.x : d.y);
    const float dkz = w.z0 ? d.x : (w.z1 ? d.y : d.z);
    w.okx = w.z0 ? o.y : (w.z1 ? o.z : o.x);
    w.oky = w.z0 ? o.z : (w.z1 ? o.x : o.y);
    w.okz = w.z0 ? o.x : (w.z1 ? o.y : o.z);
    w.Sx = dkx / dkz; w.Sy = dky / dkz; w.Sz = 1.0f / dkz;
}

__device__ __forceinline__ bool woop_test(const Woop& w, const float4 q0, const float4 q1, const float4 q2,
                                          float& t, float& bu, float& bv, float& bw0) {
    // permute the vertices to (kx, ky, kz), then translate by the permuted origin
    const float Akx = (w.z0 ? q0.y : (w.z1 ? q0.z : q0.x)) - w.okx;
    const float Aky = (w.z0 ? q0.z : (w.z1 ? q0.x : q0.y)) - w.oky;
    const float Akz = (w.z0 ? q0.x : (w.z1 ? q0.y : q0.z)) - w.okz;
    const float Bkx = (w.z0 ? q1.x : (w.z1 ? q1.y : q0.w)) - w.okx;
    const float Bky = (w.z0 ? q1.y : (w.z1 ? q0.w : q1.x)) - w.oky;
    const float Bkz = (w.z0 ? q0.w : (w.z1 ? q1.x : q1.y)) - w.okz;
    const float Ckx = (w.z0 ? q1.w : (w.z1 ? q2.x : q1.z)) - w.okx;
    const float Cky = (w.z0 ? q2.x : (w.z1 ? q1.z : q1.w)) - w.oky;
    const float Ckz = (w.z0 ? q1.z : (w.z1 ? q1.w : q2.x)) - w.okz;
    const float Ax = Akx - w.Sx * Akz, Ay = Aky - w.Sy * Akz;
    const float Bx = Bkx - w.Sx * Bkz, By = Bky - w.Sy * Bkz;
    const float Cx = Ckx - w.Sx * Ckz, Cy = Cky - w.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = (U + V) + W;
    if (det == 0.0f) return false;
    const float Az = w.Sz * Akz, Bz = w.Sz * Bkz, Cz = w.Sz * Ckz;
    const float T = (U * Az + V * Bz) + W * Cz;
    const float rcp = 1.0f / det;
    // "+ 0.0f" turns -0 into +0: omitting the kx/ky swap flips the sign of exactly-zero results only
    t = T * rcp + 0.0f; bu = V * rcp + 0.0f; bv = W * rcp + 0.0f; bw0 = U * rcp + 0.0f;
    return true;
}

__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * 5.9604644775390625e-08f; }

__device__ __forceinline__ unsigned char unorm8(float c) {
    float v = c;
    if (!(v > 0.0f)) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return (unsigned char)__float2int_rn(v * 255.0f);
}

// rgba8 imageStore of vec4(hitValue, 0.0) (main.cpp:1054); bgra: the byte order of the sample's B8G8R8A8 swapchain (main.cpp:50)
__device__ __forceinline__ uchar4 store_pixel(const TraceParams& P, float r, float g, float b) {
    const unsigned char cr = unorm8(r), cg = unorm8(g), cb = unorm8(b);
    return P.bgra ? make_uchar4(cb, cg, cr, 0) : make_uchar4(cr, cg, cb, 0);
}

__device__ __forceinline__ void load_w2o(const InstanceRec* R, float* w2o) {
    const float4* m4 = reinterpret_cast<const float4*>(R);
    const float4 m0 = __ldg(m4), m1 = __ldg(m4 + 1), m2 = __ldg(m4 + 2);
    w2o[0] = m0.x; w2o[1] = m0.y; w2o[2] = m0.z; w2o[3] = m0.w;
    w2o[4] = m1.x; w2o[5] = m1.y; w2o[6] = m1.z; w2o[7] = m1.w;
    w2o[8] = m2.x; w2o[9] = m2.y; w2o[10] = m2.z; w2o[11] = m2.w;
}

__device__ __forceinline__ rt_hit miss_record(float tmax) {
    rt_hit r;
    r.instance_id = r.geometry_index = r.primitive_id = r.custom_index = NO_HIT;
    r.t = tmax; r.u = 0.0f; r.v = 0.0f;
    return r;
}

// ---- ray identity ---------------------------------------------------------------------------------------
struct RayId { bool in_buffer, valid; uint32_t lidx, pixel, tm, out; };   // tm: tile-major ray index of the launch; out: where the RGBA8 pixel goes

// Primary ray `idx` (tile-major numbering: 8x4-pixel tiles, 32 consecutive ids per tile). A lane keeps only `idx` while it
// traverses; the pixel identity is recomputed for the epilogue (a dozen integer instructions instead of five live registers).
__device__ __forceinline__ RayId primary_id(const TraceParams& P, uint32_t idx, uint32_t tiles_x, uint32_t& x, uint32_t& y) {
    RayId id;
    const uint32_t within = idx & 31u;
    const uint32_t tile_lin = idx >> 5, region = tile_lin / TPR, tin = tile_lin - region * TPR;
    const uint32_t ry = region / P.regions_x, rx = region - ry * P.regions_x;
    const uint32_t tile_x = rx * REGION_TW + (tin % REGION_TW), tile_y = ry * REGION_TH + (tin / REGION_TW);
    (void)tiles_x;
    x = tile_x * 8u + (within & 7u);
    const uint32_t lr = P.row0 + tile_y * 4u + (within >> 3);
    const uint32_t band = lr / P.block_rows;
    y = (band * P.part_count + P.part_index) * P.block_rows + (lr - band * P.block_rows);
    id.in_buffer = x < P.width && lr < P.row0 + P.local_rows;
    id.valid = id.in_buffer && y < P.height;
    id.lidx = lr * P.width + x;
    id.pixel = y * P.width + x;
    id.tm = idx;
    id.out = P.full_frame ? id.pixel : id.lidx;
    return id;
}
// raygen shader (main.cpp:1033-1046); aspect_x/aspect_y are computed once on the host (tanf). Returns id.valid.
__device__ __forceinline__ bool primary_ray(const TraceParams& P, uint32_t idx, uint32_t tiles_x, V3& o, V3& d) {
    uint32_t x, y;
    const RayId id = primary_id(P, idx, tiles_x, x, y);
    const float scx = (float)x + 0.5f, scy = (float)y + 0.5f;
    const float ndcx = scx / (float)P.width * 2.0f - 1.0f;
    const float ndcy = scy / (float)P.height * 2.0f - 1.0f;
    const float ax = ndcx * P.aspect_x, ay = ndcy * P.aspect_y;
    o = {P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]};
    d = {(ax * 1.0f + ay * 0.0f) + 0.0f, (ax * 0.0f + ay * -1.0f) + 0.0f, (ax * 0.0f + ay * 0.0f) + -1.0f};
    return id.valid;
}

struct Counters { unsigned long long nodes = 0, tris = 0, insts = 0, hits = 0, rays = 0, edge = 0; };

template <int STAGE>
__device__ __forceinline__ void flush_stats(const TraceParams& P, const Counters& c, int lane) {
    if (!P.stats) return;
    // rt_trace_stats order: rays_primary, rays_secondary, nodes, tris, insts, primary_hits, secondary_hits, near_edge
    unsigned long long v[8] = {STAGE == 0 ? c.rays : 0ull, STAGE == 1 ? c.rays : 0ull, c.nodes, c.tris, c.insts,
                               STAGE == 0 ? c.hits : 0ull, STAGE == 1 ? c.hits : 0ull, c.edge};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], ofs);
        if (lane == 0 && v[k]) atomicAdd(P.stats + k, v[k]);
    }
}

// ---- the shaders' epilogue: closest-hit (main.cpp:1080-1091, SBT rule main.cpp:1260-1262) / miss (main.cpp:1063-1066)
// / imageStore (main.cpp:1054), plus the generation of the diffuse bounce ray (stage 0) or the blend (stage 1).
template <int STAGE, bool STATS, bool GENERAL>
__device__ __forceinline__ void shade(const TraceParams& P, const uint32_t rid, const uint32_t tiles_x, const V3 o, const V3 d,
                                      float best_t, float best_u, float best_v, float best_w0, uint32_t best_slot, uint32_t best_tri,
                                      bool& enqueue, float4& e0, float4& e1, float4& e2, Counters& c) {
    // rid: tile-major ray index (stage 0) or ray-slot index of the bounce ray (stage 1: pixel ids and the primary colour are
    // re-read from the slot {o.xyz lidx | d.xyz out | r g b -})
    RayId id;
    float col0 = 0.0f, col1 = 0.0f, col2 = 0.0f;
    if (STAGE == 0) { uint32_t x, y; id = primary_id(P, rid, tiles_x, x, y); }
    else {
        const float4* q = P.queue + 3 * (size_t)rid;
        const float4 q2 = __ldcg(q + 2);
        id.lidx = __float_as_uint(__ldcg(&q[0].w)); id.out = __float_as_uint(__ldcg(&q[1].w));
        id.pixel = 0u; id.tm = rid; id.in_buffer = id.valid = true;
        col0 = q2.x; col1 = q2.y; col2 = q2.z;
    }
    const uint32_t lidx = id.lidx;
    if (!id.valid) {
        if (id.in_buffer) {
            if (!P.full_frame) reinterpret_cast<uchar4*>(P.rgba)[lidx] = make_uchar4(0, 0, 0, 0);
            if (P.primary_hits) P.primary_hits[lidx] = miss_record(P.tmax);
            if (P.secondary_hits) P.secondary_hits[lidx] = miss_record(P.tmax);
        }
        return;
    }
    if (STATS) ++c.rays;
    const bool hit = best_slot != NO_HIT;
    const InstanceRec* R = P.instances + (hit ? best_slot : 0u);
    float sc0, sc1, sc2;
    rt_hit rec = miss_record(P.tmax);
    float4 tq0 = make_float4(0.f, 0.f, 0.f, 0.f), tq1 = tq0, tq2 = tq0;
    if (hit) {
        const uint32_t cm = __ldg(&R->custom_mask), sf = __ldg(&R->sbt_flags), inst_id = __ldg(&R->instance_id);
        const float4* t4 = reinterpret_cast<const float4*>(R->tris + best_tri);
        tq0 = __ldg(t4); tq1 = __ldg(t4 + 1); tq2 = __ldg(t4 + 2);
        const uint32_t geo = __float_as_uint(tq2.y), prim = __float_as_uint(tq2.z), custom = cm & 0xFFFFFFu;
        if (prim == 1u && inst_id == 1u && custom == 100u && geo == 1u) {
            sc0 = 1.0f - best_u - best_v; sc1 = best_u; sc2 = best_v;
        } else {
            const uint32_t r = (sf & 0xFFFFFFu) + geo * P.sbt_stride + P.sbt_offset;
            if (r < P.n_records) { sc0 = __ldg(P.hit_records + 3 * r); sc1 = __ldg(P.hit_records + 3 * r + 1); sc2 = __ldg(P.hit_records + 3 * r + 2); }
            else { sc0 = sc1 = sc2 = 0.0f; }
        }
        rec.instance_id = inst_id; rec.geometry_index = geo; rec.primitive_id = prim; rec.custom_index = custom;
        rec.t = best_t; rec.u = best_u; rec.v = best_v;
        if (GENERAL && (P.ray_flags & RT_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER)) sc0 = sc1 = sc2 = 0.0f;   // payload keeps its initial value (main.cpp:1045)
        if (STATS) { ++c.hits; if (STAGE == 0 && fminf(fminf(best_u, best_v), best_w0) < 9.5367431640625e-07f) ++c.edge; }
    } else {
        sc0 = P.miss[0]; sc1 = P.miss[1]; sc2 = P.miss[2];
    }
    if (STAGE == 0) {
        if (P.primary_hits) P.primary_hits[lidx] = rec;
        if (hit && P.bounces > 0u && !(GENERAL && (P.ray_flags & RT_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER))) {
            // ---- deterministic diffuse bounce (our definition; the reference's recursion depth is 1) ----
            const V3 p = {o.x + best_t * d.x, o.y + best_t * d.y, o.z + best_t * d.z};
            float w2o[12];
            load_w2o(R, w2o);
            const V3 ed1 = {tq0.w - tq0.x, tq1.x - tq0.y, tq1.y - tq0.z};
            const V3 ed2 = {tq1.z - tq0.x, tq1.w - tq0.y, tq2.x - tq0.z};
            V3 n = xform_normal(w2o, cross3(ed1, ed2));
            const float l2 = dot3(n, n);
            if (l2 > 0.0f && l2 < INFINITY) { const float l = sqrtf(l2); n = {n.x / l, n.y / l, n.z / l}; }
            else { const float dl = sqrtf(dot3(d, d)); n = {-d.x / dl, -d.y / dl, -d.z / dl}; }
            if (dot3(n, d) > 0.0f) n = {-n.x, -n.y, -n.z};
            uint32_t h = pcg_hash(id.pixel + pcg_hash(P.bounce_seed + 0x9E3779B9u));
            V3 s = {0.0f, 0.0f, 0.0f};
            for (int tries = 0; tries < 8; ++tries) {
                const uint32_t ha = pcg_hash(h), hb = pcg_hash(ha), hc = pcg_hash(hb);
                h = hc;
                const V3 q = {u01(ha) * 2.0f - 1.0f, u01(hb) * 2.0f - 1.0f, u01(hc) * 2.0f - 1.0f};
                const float qq = dot3(q, q);
                if (qq <= 1.0f && qq > 1e-8f) { const float ql = sqrtf(qq); s = {q.x / ql, q.y / ql, q.z / ql}; break; }
            }
            V3 dir = {n.x + s.x, n.y + s.y, n.z + s.z};
            const float dl2 = dot3(dir, dir);
            if (dl2 < 1e-12f) dir = n;
            else { const float dl = sqrtf(dl2); dir = {dir.x / dl, dir.y / dl, dir.z / dl}; }
            const float eps = 0.0009765625f;
            enqueue = true;
            e0 = make_float4(p.x + n.x * eps, p.y + n.y * eps, p.z + n.z * eps, __uint_as_float(lidx));
            e1 = make_float4(dir.x, dir.y, dir.z, __uint_as_float(id.out));
            e2 = make_float4(sc0, sc1, sc2, 0.0f);
        } else {
            reinterpret_cast<uchar4*>(P.rgba)[id.out] = store_pixel(P, sc0, sc1, sc2);   // imageStore, main.cpp:1054
            if (P.secondary_hits) P.secondary_hits[lidx] = miss_record(P.tmax);
        }
    } else {
        if (P.secondary_hits) P.secondary_hits[lidx] = rec;
        const float f0 = 0.5f * col0 + 0.5f * sc0, f1 = 0.5f * col1 + 0.5f * sc1, f2 = 0.5f * col2 + 0.5f * sc2;
        reinterpret_cast<uchar4*>(P.rgba)[id.out] = store_pixel(P, f0, f1, f2);
    }
}

#if !RT_BOUNCE_ORDERED
// warp-aggregated append to the bounce queue (all 32 lanes must call)
__device__ __forceinline__ void enqueue_bounce(const TraceParams& P, bool enqueue, const float4& e0, const float4& e1, const float4& e2,
                                               int lane, uint32_t lt_mask) {
    const unsigned em = __ballot_sync(0xffffffffu, enqueue);
    if (em) {
        const int leader = __ffs(em) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(P.counters + 2, (uint32_t)__popc(em));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (enqueue) {
            float4* q = P.queue + 3 * (size_t)(base + __popc(em & lt_mask));
            __stcg(q, e0); __stcg(q + 1, e1); __stcg(q + 2, e2);
        }
    }
}
#endif

// STAGE 0: primary rays generated from pixel ids. STAGE 1: secondary rays read from the bounce queue.
// The shaders' epilogue runs inside this kernel for the lanes whose ray just finished. (Measured alternatives, both
// slower on B200 - profiles/README.md r01h: a separate warp-convergent shading kernel with refill thresholds 16..28,
// and speculative traversal with one postponed leaf.)
template <int STAGE, bool STATS, int STACK, bool GENERAL, bool BIG>
__global__ void __launch_bounds__(BIG ? TRACE_THREADS_BIG : TRACE_THREADS, BIG ? 1 : (STAGE == 1 ? RT_TRACE_MIN_BLOCKS_S1 : TRACE_MIN_BLOCKS)) k_trace(const TraceParams P) {
    const int lane = threadIdx.x & 31;
    const BvhNode* tlas_nodes = P.tlas_nodes;
    if (BIG) {
        // stage the TLAS nodes: one thread arms the mbarrier with the byte count and issues ONE bulk copy; everyone waits on phase 0
        extern __shared__ __align__(128) unsigned char s_dyn[];
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_dyn);
        const uint32_t dst = bar + SMEM_TLAS_HEADER;
        const uint32_t bytes = P.tlas_smem_nodes * (uint32_t)sizeof(BvhNode);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(P.tlas_nodes), "r"(bytes), "r"(bar) : "memory");
        }
        __syncthreads();
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar) : "memory");
        }
        tlas_nodes = reinterpret_cast<const BvhNode*>(s_dyn + SMEM_TLAS_HEADER);
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t tiles_x = (P.width + 7u) >> 3;
    const uint32_t tiles_y = (P.local_rows + 3u) >> 2;
    constexpr bool FUSED = STAGE == 2;       // both stages in one launch: a lane holds a primary OR a secondary ray (`sec`)
    (void)tiles_y;
    const uint32_t n_prim = P.n_regions * RPR;                             // ray ids are padded to whole regions; the padding is never `valid`
    const uint32_t total = STAGE == 1 ? P.counters[2] : n_prim;
    uint32_t* fetch_counter = P.counters + (STAGE == 1 ? 1 : 0);
    constexpr int THRESHOLD = REFILL_THRESHOLD;
    // RT_REGIONS == 2 (not in the fused launch, which keeps the one global counter): every region has its own fetch counter. Stage 0 hands
    // out the rays of an image region, stage 1 those of a contiguous chunk of the (region-ordered) bounce list. The warps of an SM start in
    // that SM's home region and move on to the next open one when theirs is exhausted, so they all work on the same patch of the image (the
    // same one or two instances) and the BLAS nodes one warp pulls into L1 are hits for the others.
    constexpr bool USE_REGIONS = RT_REGIONS == 2 && !FUSED;
    uint32_t my_region = 0, region_len = 1;
    uint32_t* region_next = nullptr;
    if (USE_REGIONS) {
        uint32_t smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        my_region = (uint32_t)(((uint64_t)(smid % P.n_sm) * P.n_regions) / P.n_sm);      // warp-uniform
        region_len = STAGE == 1 ? (total + P.n_regions - 1u) / P.n_regions : RPR;
        if (region_len == 0u) region_len = 1u;
        region_next = P.region_next + (STAGE == 1 ? P.n_regions : 0u);
    }
#if RT_SMEM_STACK > 0
    __shared__ int32_t s_stack[RT_SMEM_STACK][BIG ? TRACE_THREADS_BIG : TRACE_THREADS];
#endif

    Counters c, c1;                          // c1: FUSED only, the secondary rays' ray/hit counts

    // ---- per-lane ray state ----
    bool have_ray = false, exhausted = false;
    bool sec = STAGE == 1;                   // FUSED: this lane's ray is a bounce ray
    bool pending = false;                    // FUSED: holds a claim on bounce-queue entry rid that is not written yet
    bool prim_empty = false;                 // FUSED, warp-uniform: the primary pool is exhausted
    uint32_t idle_polls = 0;
    uint32_t fin_local = 0;                  // FUSED, warp-uniform: primary rays this warp finished but has not reported yet
    uint32_t rid = 0;                        // tile-major index of the primary ray / slot of the bounce ray this lane traces
    V3 o = {0.0f, 0.0f, 0.0f}, d = {0.0f, 0.0f, 1.0f};
    // primary-only launches: every ray starts at the camera, so the origin is read from the parameter block instead of living in three registers
#if RT_PRIMARY_ORIGIN_CONST
#define RAY_O (STAGE == 0 ? V3{P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]} : o)
#else
#define RAY_O o
#endif
    int32_t cur = REF_DONE;
    int sp = 0;
    bool in_blas = false;
    const BvhNode* nodes = tlas_nodes;
    const TriRec* tris = nullptr;
    Slab sl; Woop wp;
    sl.rdx = sl.rdy = sl.rdz = sl.cnx = sl.cny = sl.cnz = sl.cfx = sl.cfy = sl.cfz = 0.0f; sl.px = sl.py = sl.pz = false;
    wp.okx = wp.oky = wp.okz = wp.Sx = wp.Sy = wp.Sz = 0.0f; wp.z0 = wp.z1 = false;
    uint32_t cur_slot = 0;
    uint32_t cur_iflags = 0;                 // GENERAL: instance flags of the BLAS being traversed
    uint32_t cur_sbt = 0;                    // GENERAL: instanceShaderBindingTableRecordOffset of that instance (any-hit record lookup)
    V3 cur_od = {0.0f, 0.0f, 0.0f};          // GENERAL: object-space ray direction (facing test)
    float best_t = P.tmax, best_u = 0.0f, best_v = 0.0f, best_w0 = 0.0f;
    uint32_t best_slot = NO_HIT, best_tri = 0;
    int32_t stack[STACK - RT_SMEM_STACK > 0 ? STACK - RT_SMEM_STACK : 1];
#if RT_CAP_EVERY > 1
    uint32_t cap_ctr = 0;
#endif
#if RT_STACK_TOS
    // the top of the stack lives in a register: a push directly followed by a pop (both children hit, the near one turns out to be a dead end)
    // never touches local memory; a push onto an occupied top spills the old top first
    int32_t tos = REF_DONE; bool tos_valid = false;
    auto push = [&](int32_t v) {
        if (tos_valid) { stack[sp] = tos; ++sp; }
        tos = v; tos_valid = true;
    };
    auto pop = [&]() -> int32_t {
        if (tos_valid) { tos_valid = false; return tos; }
        --sp;
        return stack[sp];
    };
#else
    auto push = [&](int32_t v) {
#if RT_SMEM_STACK > 0
        if (sp < RT_SMEM_STACK) s_stack[sp][threadIdx.x] = v; else stack[sp - RT_SMEM_STACK] = v;
#else
        stack[sp] = v;
#endif
        ++sp;
    };
    auto pop = [&]() -> int32_t {
        --sp;
#if RT_SMEM_STACK > 0
        return sp < RT_SMEM_STACK ? s_stack[sp][threadIdx.x] : stack[sp - RT_SMEM_STACK];
#else
        return stack[sp];
#endif
    };
#endif

    for (;;) {
        // ================= refill idle lanes: ballot + one atomic + shuffle =================
        const unsigned need = __ballot_sync(0xffffffffu, !have_ray && !exhausted && !pending);
        if (need && !(FUSED && prim_empty)) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            uint32_t idx = 0xFFFFFFFFu;                                     // the ray this lane gets; 0xFFFFFFFF: none this round
            if (USE_REGIONS) {
                const uint32_t n = (uint32_t)__popc(need);
                bool pool_empty = false;
                for (;;) {
                    if (lane == leader) base = atomicAdd(region_next + my_region, n);
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (base < region_len) break;
                    // this region is exhausted: look for an open one, 32 candidates per step (plain loads; the atomic decides)
                    bool found = false;
                    for (uint32_t ofs = 1; ofs < P.n_regions; ofs += 32u) {
                        uint32_t r = my_region + ofs + (uint32_t)lane;
                        if (r >= P.n_regions) r -= P.n_regions;
                        const bool open = ofs + (uint32_t)lane < P.n_regions && ld_volatile_u32(region_next + r) < region_len;
                        const unsigned m = __ballot_sync(0xffffffffu, open);
                        if (m) { my_region = __shfl_sync(0xffffffffu, r, __ffs(m) - 1); found = true; break; }
                    }
                    if (!found) { pool_empty = true; break; }
                }
                if (!have_ray && !exhausted && !pending) {
                    if (pool_empty) exhausted = true;
                    else {
                        const uint32_t k = base + __popc(need & lt_mask);
                        if (k < region_len) { idx = my_region * region_len + k; if (idx >= total) idx = 0xFFFFFFFFu; }
                    }
                }
            } else {
                if (lane == leader) base = atomicAdd(fetch_counter, (uint32_t)__popc(need));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (FUSED && base + (uint32_t)__popc(need) >= total) prim_empty = true;
                if (!have_ray && !exhausted && !pending) {
                    idx = base + __popc(need & lt_mask);
                    if (idx >= total) { idx = 0xFFFFFFFFu; if (!FUSED) exhausted = true; }
                }
            }
            if (!have_ray && !exhausted && !pending) {
                if (idx != 0xFFFFFFFFu) {
                    have_ray = true;
                    bool valid = true;
                    if (STAGE != 1) {
                        sec = false;
                        rid = idx;
                        valid = primary_ray(P, idx, tiles_x, o, d);
                    } else {
#if RT_BOUNCE_ORDERED
                        rid = __ldg(P.bounce_index + idx);
#else
                        rid = idx;
#endif
                        const float4* q = P.queue + 3 * (size_t)rid;
                        const float4 q0 = __ldcg(q), q1 = __ldcg(q + 1);
                        o = {q0.x, q0.y, q0.z};
                        d = {q1.x, q1.y, q1.z};
                    }
                    // ---- traceRayEXT(topLevelAS, Opaque, cullMask, ..., o, tmin, d, tmax) (main.cpp:1047-1052) ----
                    best_t = P.tmax; best_u = best_v = best_w0 = 0.0f; best_slot = NO_HIT; best_tri = 0;
                    sp = 0;
#if RT_STACK_TOS
                    tos_valid = false;
#endif
                    push(REF_DONE);
                    in_blas = false;
                    nodes = tlas_nodes;
                    cur = valid ? P.tlas_root : REF_DONE;
                    if (cur == REF_EMPTY) cur = REF_DONE;
                    slab_setup(sl, o, d, P.tlas_absmax[0], P.tlas_absmax[1], P.tlas_absmax[2]);
                }
            }
        }
        if (FUSED) {
            if (prim_empty && fin_local) {            // report this warp's finished primary rays (after everything they published)
                __threadfence();
                if (lane == 0) atomicAdd(P.counters + 3, fin_local);
                fin_local = 0;
            }
            // Primary pool empty: idle lanes CLAIM bounce-queue entries (possibly ahead of the producers) ...
            const unsigned need2 = __ballot_sync(0xffffffffu, !have_ray && !exhausted && !pending);
            if (need2 && prim_empty) {
                const int leader = __ffs(need2) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(P.counters + 1, (uint32_t)__popc(need2));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!have_ray && !exhausted && !pending) { rid = base + __popc(need2 & lt_mask); pending = true; }
            }
            // ... and start the ray as soon as its entry has been published (flag == this launch's epoch). A claim beyond the
            // final queue length can only be recognised once every primary ray has finished (counters[3] == total).
            if (pending) {
                bool ready = false;
                if (rid < P.queue_capacity && ld_acquire_u32(P.queue_flags + rid) == P.epoch) ready = true;
                else if (ld_acquire_u32(P.counters + 3) == total) {
                    if (rid >= ld_volatile_u32(P.counters + 2)) { pending = false; exhausted = true; }
                    else if (ld_acquire_u32(P.queue_flags + rid) == P.epoch) ready = true;
                }
                if (ready) {
                    const float4* q = P.queue + 3 * (size_t)rid;
                    const float4 q0 = __ldcg(q), q1 = __ldcg(q + 1);
                    o = {q0.x, q0.y, q0.z};
                    d = {q1.x, q1.y, q1.z};
                    pending = false; have_ray = true; sec = true;
                    best_t = P.tmax; best_u = best_v = best_w0 = 0.0f; best_slot = NO_HIT; best_tri = 0;
                    sp = 0;
#if RT_STACK_TOS
                    tos_valid = false;
#endif
                    push(REF_DONE);
                    in_blas = false;
                    nodes = tlas_nodes;
                    cur = P.tlas_root;
                    if (cur == REF_EMPTY) cur = REF_DONE;
                    slab_setup(sl, o, d, P.tlas_absmax[0], P.tlas_absmax[1], P.tlas_absmax[2]);
                }
            }
            if (__ballot_sync(0xffffffffu, have_ray) == 0u) {
                if (__ballot_sync(0xffffffffu, !exhausted) == 0u) break;      // every lane is done
                if (__ballot_sync(0xffffffffu, pending) != 0u) {                  // only unpublished claims left: poll again
                    __nanosleep(200);
                    if (++idle_polls > (1u << 23)) {                              // ~2 s: protocol violated -> error flag instead of a hung GPU
                        if (lane == 0 && P.error_flag) atomicExch(P.error_flag, 2);
                        break;
                    }
                }
                continue;
            }
        }
        if (!FUSED && __ballot_sync(0xffffffffu, have_ray) == 0u) break;
        // nobody is waiting for work any more -> no reason to leave the traversal loops early
        const bool warp_exhausted = FUSED ? __ballot_sync(0xffffffffu, !have_ray && !exhausted) == 0u : __ballot_sync(0xffffffffu, exhausted) != 0u;

        // ================= two-level while-while traversal =================
        while (cur != REF_DONE) {
            while ((uint32_t)cur < (uint32_t)REF_SENTINEL_MIN) {               // internal node
                const F8 na = ldg256<BIG>(&nodes[cur].c[0]), nb = ldg256<BIG>(&nodes[cur].c[1]);
                const float4 a0 = na.a, a1 = na.b, b0 = nb.a, b1 = nb.b;
                if (STATS) ++c.nodes;
                float t0, t1;
                const bool hit0 = slab_test(sl, a0, a1, P.tmin, best_t, t0);
                const bool hit1 = slab_test(sl, b0, b1, P.tmin, best_t, t1);
                const int32_t r0 = __float_as_int(a1.z), r1 = __float_as_int(b1.z);
#if RT_BRANCHLESS_NODE
                {   // selects + one predicated store / load instead of a four-way branch (no BSSY/BSYNC pair around it)
                    const bool both = hit0 && hit1, any = hit0 || hit1, swap = t1 < t0;
                    const int32_t nearr = both ? (swap ? r1 : r0) : (hit0 ? r0 : r1);
                    if (both) push(swap ? r0 : r1);
                    int32_t nxt = nearr;
                    if (!any) nxt = pop();
                    cur = nxt;
                }
#else
                if (hit0 && hit1) {
                    const bool swap = t1 < t0;
                    push(swap ? r0 : r1);
                    cur = swap ? r1 : r0;
                } else if (hit0) cur = r0;
                else if (hit1) cur = r1;
                else cur = pop();
#endif
#if RT_PREFETCH_LEAF
                // experiment: a lane that reaches a leaf waits for the lanes still in the node loop; start the fetch of what it will read
                // (first triangle record / instance record) now, without holding registers for it
                if (cur < 0) {
                    const char* pa = in_blas ? reinterpret_cast<const char*>(tris + leaf_first(cur)) : reinterpret_cast<const char*>(P.instances + leaf_first(cur));
                    if (RT_PREFETCH_LEAF != 3 || in_blas) asm volatile("prefetch.global.L1 [%0];" ::"l"(pa));
                    if (RT_PREFETCH_LEAF == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(pa + 47));
                }
#endif
#if RT_CAP_EVERY > 1
                if (NODE_CAP > 0 && (++cap_ctr % RT_CAP_EVERY) == 0 && __popc(__activemask()) < NODE_CAP) break;
#else
                if (NODE_CAP > 0 && __popc(__activemask()) < NODE_CAP) break;   // do not idle the warp behind a few long node chains
#endif
            }
            if (cur < 0) {                                                       // leaf
                const uint32_t first = leaf_first(cur), count = leaf_count(cur);
                if (!in_blas) {
                    // TLAS leaf = one instance: cull mask (main.cpp:851,1048), then enter its BLAS in object space
                    const InstanceRec* R = P.instances + first;
                    const uint32_t cm = __ldg(&R->custom_mask);
                    const int32_t root = __ldg(&R->root);
                    if (((cm >> 24) & P.cull_mask) != 0u && root != REF_EMPTY) {
                        if (STATS) ++c.insts;
                        float w2o[12];
                        load_w2o(R, w2o);
                        const V3 oo = xform_point(w2o, RAY_O), od = xform_vec(w2o, d);
                        slab_setup(sl, oo, od, __ldg(&R->absmax[0]), __ldg(&R->absmax[1]), __ldg(&R->absmax[2]));
                        woop_setup(wp, oo, od);
                        nodes = R->nodes; tris = R->tris;
                        cur_slot = first;
                        if (GENERAL) { const uint32_t sf = __ldg(&R->sbt_flags); cur_iflags = sf >> 24; cur_sbt = sf & 0xFFFFFFu; cur_od = od; }
                        in_blas = true;
                        push(REF_POP_INSTANCE);
                        cur = root;
                    } else cur = pop();
                } else {
                    bool terminated = false;
                    for (uint32_t k = 0; k < count; ++k) {
                        const float4* t4 = reinterpret_cast<const float4*>(tris + first + k);
                        const float4 q0 = ldg_tri(t4), q1 = ldg_tri(t4 + 1), q2 = ldg_tri(t4 + 2);
                        if (STATS) ++c.tris;
                        float t, bu, bv, bw0;
                        bool anyhit_terminates = false;
                        if (woop_test(wp, q0, q1, q2, t, bu, bv, bw0) && t > P.tmin && t < P.tmax) {
                            if (GENERAL) {
                                // opacity: geometry flag -> instance FORCE_* -> ray flags; then the opacity culls
                                const uint32_t rf = P.ray_flags;
                                bool opaque = (__float_as_uint(q2.w) & RT_GEOMETRY_OPAQUE) != 0u;
                                if (cur_iflags & RT_INSTANCE_FORCE_OPAQUE) opaque = true;
                                else if (cur_iflags & RT_INSTANCE_FORCE_NO_OPAQUE) opaque = false;
                                if (rf & RT_RAY_FLAG_OPAQUE) opaque = true;
                                else if (rf & RT_RAY_FLAG_NO_OPAQUE) opaque = false;
                                if (opaque ? (rf & RT_RAY_FLAG_CULL_OPAQUE) : (rf & RT_RAY_FLAG_CULL_NO_OPAQUE)) continue;
                                if ((rf & (RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES | RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES)) &&
                                    !(cur_iflags & RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE)) {
                                    // front = clockwise seen from the ray origin = ((v1-v0) x (v2-v0)) . d > 0, object space
                                    const V3 ed1 = {q0.w - q0.x, q1.x - q0.y, q1.y - q0.z};
                                    const V3 ed2 = {q1.z - q0.x, q1.w - q0.y, q2.x - q0.z};
                                    const bool front = (dot3(cross3(ed1, ed2), cur_od) > 0.0f) != ((cur_iflags & RT_INSTANCE_TRIANGLE_FLIP_FACING) != 0u);
                                    if (front ? (rf & RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES) : (rf & RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES)) continue;
                                }
                                // any-hit stage of the hit group (include/rtcore.h, rt_anyhit_record): non-opaque candidates inside the
                                // current ray interval only; fixed-function alpha test on the candidate's barycentrics
                                if (!opaque && P.n_anyhit) {
                                    if (t > best_t) continue;
                                    const uint32_t ar = cur_sbt + __float_as_uint(q2.y) * P.sbt_stride + P.sbt_offset;
                                    if (ar < P.n_anyhit) {
                                        const uint4 a = __ldg(P.anyhit + ar);
                                        if (a.x == RT_ANYHIT_ALPHA_MASK) {
                                            const uint32_t res = 1u << a.y;
                                            uint32_t cu = (uint32_t)(int)(bu * (float)res), cv = (uint32_t)(int)(bv * (float)res);
                                            cu = min(cu, res - 1u); cv = min(cv, res - 1u);
                                            const uint32_t bit = cv * res + cu;
                                            const uint32_t word = __ldg(reinterpret_cast<const uint32_t*>(P.anyhit) + a.w + (bit >> 5));
                                            if (((word >> (bit & 31u)) & 1u) == 0u) continue;             // ignoreIntersectionEXT
                                        }
                                        if (a.z & RT_ANYHIT_TERMINATE_RAY) anyhit_terminates = true;      // terminateRayEXT after the commit below
                                    }
                                }
                            }
                            bool better = t < best_t;
                            if (t == best_t && best_slot != NO_HIT) {
                                // equal t: lowest (instance, geometry, primitive) wins; rare, so the ids of the
                                // current best are re-read from memory instead of living in registers
                                const InstanceRec* Rb = P.instances + best_slot;
                                const uint32_t bi = __ldg(&Rb->instance_id), ci = __ldg(&(P.instances + cur_slot)->instance_id);
                                const TriRec* bt = Rb->tris + best_tri;
                                const uint32_t bg = __ldg(&bt->geo), bp = __ldg(&bt->prim);
                                const uint32_t geo = __float_as_uint(q2.y), prim = __float_as_uint(q2.z);
                                better = ci != bi ? ci < bi : (geo != bg ? geo < bg : prim < bp);
                            }
                            if (better) { best_t = t; best_u = bu; best_v = bv; best_w0 = bw0; best_slot = cur_slot; best_tri = first + k; }
                            if (GENERAL && ((P.ray_flags & RT_RAY_FLAG_TERMINATE_ON_FIRST_HIT) || anyhit_terminates)) { terminated = true; break; }
                        }
                    }
                    if (GENERAL && terminated) { cur = REF_DONE; break; }
                    cur = pop();
                }
            } else if (cur == REF_POP_INSTANCE) {                                // back to world space
                in_blas = false;
                nodes = tlas_nodes;
                slab_setup(sl, RAY_O, d, P.tlas_absmax[0], P.tlas_absmax[1], P.tlas_absmax[2]);
                cur = pop();
            } else if (cur == REF_EMPTY) {
                cur = pop();
            }
            // warp-level compaction trigger: too few lanes still traversing -> go refill the idle ones
            if (!warp_exhausted && __popc(__activemask()) < THRESHOLD) break;
        }

        // ================= epilogue of the lanes whose ray just finished =================
        const bool finish = have_ray && cur == REF_DONE;
        bool enqueue = false;
        float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0, e2 = e0;
        if (finish) {
            have_ray = false;
            if (FUSED) {
                if (sec) shade<1, STATS, GENERAL>(P, rid, tiles_x, o, d, best_t, best_u, best_v, best_w0, best_slot, best_tri, enqueue, e0, e1, e2, c1);
                else shade<0, STATS, GENERAL>(P, rid, tiles_x, o, d, best_t, best_u, best_v, best_w0, best_slot, best_tri, enqueue, e0, e1, e2, c);
            } else {
                shade<STAGE == 2 ? 0 : STAGE, STATS, GENERAL>(P, rid, tiles_x, RAY_O, d, best_t, best_u, best_v, best_w0, best_slot, best_tri, enqueue, e0, e1, e2, c);
            }
        }
        if (FUSED) {
            // append the bounce rays (warp-aggregated slot allocation): entries, ONE fence for the warp, then the per-entry
            // publication flags (= this launch's epoch). Finished primary rays are counted per warp and added to counters[3]
            // only once the primary pool is empty (that is when consumers start to care); the count is released after the
            // slot allocations it covers, so "counters[3] == total" implies counters[2] is final.
            const unsigned em = __ballot_sync(0xffffffffu, enqueue);
            if (em) {
                const int leader = __ffs(em) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(P.counters + 2, (uint32_t)__popc(em));
                base = __shfl_sync(0xffffffffu, base, leader);
                const uint32_t pos = base + __popc(em & lt_mask);
                if (enqueue) {
                    float4* q = P.queue + 3 * (size_t)pos;
                    __stcg(q, e0); __stcg(q + 1, e1); __stcg(q + 2, e2);
                }
                __threadfence();
                if (enqueue) st_volatile_u32(P.queue_flags + pos, P.epoch);
            }
            fin_local += (uint32_t)__popc(__ballot_sync(0xffffffffu, finish && !sec));    // reported at the top of the loop
        } else {
#if RT_BOUNCE_ORDERED
        if (STAGE == 0 && enqueue) {
            float4* q = P.queue + 3 * (size_t)rid;
            __stcg(q, e0); __stcg(q + 1, e1); __stcg(q + 2, e2);
            atomicOr(P.tile_mask + (rid >> 5), 1u << (rid & 31u));
        }
#else
        if (STAGE == 0) enqueue_bounce(P, enqueue, e0, e1, e2, lane, lt_mask);
#endif
        }
    }
    if (STATS) { flush_stats<STAGE == 2 ? 0 : STAGE>(P, c, lane); if (FUSED) flush_stats<1>(P, c1, lane); }
}

// ---- bounce index: per-tile hit masks -> compact tile-ordered list of ray slots -----------------------------
constexpr int BIDX_THREADS = 256, BIDX_WORDS = 4, BIDX_CHUNK = BIDX_THREADS * BIDX_WORDS;   // mask words per block

// block_sums[b] = number of bounce rays in the b-th chunk of mask words; counters[2] += it
__global__ void __launch_bounds__(BIDX_THREADS) k_bounce_count(const uint32_t* __restrict__ tile_mask, uint32_t n_words,
                                                               uint32_t* __restrict__ block_sums, uint32_t* __restrict__ counters) {
    __shared__ uint32_t s_w[BIDX_THREADS / 32];
    const uint32_t w0 = blockIdx.x * BIDX_CHUNK + threadIdx.x * BIDX_WORDS;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < BIDX_WORDS; ++k) if (w0 + k < n_words) c += __popc(__ldg(tile_mask + w0 + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < BIDX_THREADS / 32; ++w) t += s_w[w];
        block_sums[blockIdx.x] = t;
        if (t) atomicAdd(counters + 2, t);
    }
}

// index[prefix + rank] = 32 * word + bit, in word/bit order (= tile-major pixel order)
__global__ void __launch_bounds__(BIDX_THREADS) k_bounce_index(const uint32_t* __restrict__ tile_mask, uint32_t n_words,
                                                               const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ index) {
    __shared__ uint32_t s_w[BIDX_THREADS / 32];
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // exclusive prefix of the preceding blocks' sums (a few hundred values)
    uint32_t pre = 0;
    for (uint32_t b = threadIdx.x; b < blockIdx.x; b += BIDX_THREADS) pre += __ldg(block_sums + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if (lane == 0) s_w[warp] = pre;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < BIDX_THREADS / 32; ++w) t += s_w[w]; s_base = t; }
    __syncthreads();
    const uint32_t w0 = blockIdx.x * BIDX_CHUNK + threadIdx.x * BIDX_WORDS;
    uint32_t m[BIDX_WORDS], c = 0;
#pragma unroll
    for (int k = 0; k < BIDX_WORDS; ++k) { m[k] = w0 + k < n_words ? __ldg(tile_mask + w0 + k) : 0u; c += __popc(m[k]); }
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t pos = s_base + inc - c;
    for (int w = 0; w < warp; ++w) pos += s_w[w];
#pragma unroll
    for (int k = 0; k < BIDX_WORDS; ++k) {
        uint32_t mm = m[k];
        while (mm) { const int b = __ffs(mm) - 1; mm &= mm - 1u; index[pos++] = (w0 + k) * 32u + (uint32_t)b; }
    }
}
// ---- exact static SBT range: max_i (instanceSbtOffset_i + (geometryCount_i - 1) * sbtRecordStride), rule main.cpp:1260-1262 ----
__global__ void __launch_bounds__(256) k_sbt_bound(const InstanceRec* __restrict__ inst, uint32_t n, uint32_t stride, unsigned long long* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0ull;
    if (i < n) v = (unsigned long long)(__ldg(&inst[i].sbt_flags) & 0xFFFFFFu) + (unsigned long long)(__ldg(&inst[i].active) >> 1) * stride;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}
// ---- stream-ordered flags in (peer) device memory: the multi-GPU frame handshake without a collective ----------------------
__global__ void k_flag_add(uint32_t* counter) {
    __threadfence_system();
    atomicAdd_system(counter, 1u);
}
__global__ void k_flag_wait_ge(const uint32_t* counter, uint32_t target, int* error_flag) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if ((int32_t)(v - target) >= 0) break;
        __nanosleep(spins < 64 ? 100 : 1000);
        if (++spins > (1u << 22)) { if (error_flag) atomicExch(error_flag, 3); break; }   // ~4 s: give up instead of hanging the GPU
    }
}

__global__ void __launch_bounds__(256) k_unpack_rows(const uchar4* __restrict__ packed_all, uint32_t width, uint32_t height,
                                                    uint32_t block_rows, uint32_t part_count, uint32_t rows_per_part, uchar4* __restrict__ out) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y;
    if (x >= width || y >= height) return;
    const uint32_t band = y / block_rows, part = band % part_count, local_band = band / part_count;
    const uint32_t lr = local_band * block_rows + (y - band * block_rows);
    out[(size_t)y * width + x] = packed_all[((size_t)part * rows_per_part + lr) * width + x];
}

template <int STAGE, bool STATS, int STACK, bool GENERAL, bool BIG>
int launch_stage_v(const TraceParams& p, int sm_count, cudaStream_t st) {
    constexpr int THREADS = BIG ? TRACE_THREADS_BIG : TRACE_THREADS;
    const size_t smem = BIG ? SMEM_TLAS_HEADER + (size_t)p.tlas_smem_nodes * sizeof(BvhNode) : 0;
    static int per_device[64] = {};     // occupancy of this instantiation, cached per device ordinal
    int dev = 0;
    cudaGetDevice(&dev);
    int blocks_per_sm = dev >= 0 && dev < 64 ? per_device[dev] : 0;
    if (blocks_per_sm == 0) {
        if (BIG && cudaFuncSetAttribute(k_trace<STAGE, STATS, STACK, GENERAL, BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(SMEM_TLAS_HEADER + (size_t)RT_SMEM_TLAS_MAX_NODES * sizeof(BvhNode))) != cudaSuccess) return -1;
#if RT_L1_CARVEOUT >= 0
        // experiment: ask for a fixed shared-memory carve-out (per cent of the L1/shared array); the kernel itself uses no shared memory
        if (!BIG) cudaFuncSetAttribute(k_trace<STAGE, STATS, STACK, GENERAL, BIG>, cudaFuncAttributePreferredSharedMemoryCarveout, RT_L1_CARVEOUT);
#endif
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_trace<STAGE, STATS, STACK, GENERAL, BIG>, THREADS, smem) != cudaSuccess || blocks_per_sm < 1)
            blocks_per_sm = 1;
        if (dev >= 0 && dev < 64) per_device[dev] = blocks_per_sm;
    }
    const uint32_t tiles = ((p.width + 7u) >> 3) * ((p.local_rows + 3u) >> 2);
    uint32_t blocks = (uint32_t)(sm_count * blocks_per_sm);            // persistent: a multiple of the SM count
    const uint32_t max_useful = (tiles + (THREADS / 32) - 1) / (THREADS / 32);
    if (STAGE != 1 && blocks > max_useful) blocks = max_useful;
    if (blocks == 0) return 0;
    k_trace<STAGE, STATS, STACK, GENERAL, BIG><<<blocks, THREADS, smem, st>>>(p);
    return 1;
}
template <int STAGE, bool STATS, int STACK, bool GENERAL>
int launch_stage(const TraceParams& p, int sm_count, cudaStream_t st) {
#if RT_SMEM_TLAS
    if (p.tlas_smem_nodes) return launch_stage_v<STAGE, STATS, STACK, GENERAL, true>(p, sm_count, st);
#endif
    return launch_stage_v<STAGE, STATS, STACK, GENERAL, false>(p, sm_count, st);
}

template <bool STATS, int STACK, bool GENERAL>
int launch_both(const TraceParams& p, int sm_count, cudaStream_t st) {
#if RT_FUSED_STAGES
    if (p.bounces > 0 && (uint64_t)p.width * p.local_rows <= (uint64_t)RT_FUSED_MAX_PIXELS) return launch_stage<2, STATS, STACK, GENERAL>(p, sm_count, st);
#endif
    int n = launch_stage<0, STATS, STACK, GENERAL>(p, sm_count, st);
    if (p.bounces > 0) {
#if RT_BOUNCE_ORDERED
        const uint32_t n_words = p.n_regions * TPR;            // one mask word per (padded) tile
        const uint32_t blocks = (n_words + BIDX_CHUNK - 1) / BIDX_CHUNK;
        uint32_t* block_sums = p.tile_mask + n_words;
        k_bounce_count<<<blocks, BIDX_THREADS, 0, st>>>(p.tile_mask, n_words, block_sums, p.counters);
        k_bounce_index<<<blocks, BIDX_THREADS, 0, st>>>(p.tile_mask, n_words, block_sums, p.bounce_index);
        n += 2;
#endif
        n += launch_stage<1, STATS, STACK, GENERAL>(p, sm_count, st);
    }
    return n;
}

template <int STACK>
int launch_stack(const TraceParams& p, bool stats, int sm_count, cudaStream_t st) {
    // the sample's flags (Opaque) and NoOpaque change nothing without any-hit records: fast variant
    const bool general = (p.ray_flags & ~(uint32_t)(RT_RAY_FLAG_OPAQUE | RT_RAY_FLAG_NO_OPAQUE)) != 0u ||
                         (p.n_anyhit != 0u && !(p.ray_flags & RT_RAY_FLAG_OPAQUE));        // any-hit records can only matter for non-opaque candidates
    if (general) return stats ? launch_both<true, STACK, true>(p, sm_count, st) : launch_both<false, STACK, true>(p, sm_count, st);
    return stats ? launch_both<true, STACK, false>(p, sm_count, st) : launch_both<false, STACK, false>(p, sm_count, st);
}

}  // namespace

int launch_trace(const TraceParams& p_in, bool stats, int stack_needed, int sm_count, cudaStream_t st) {
    TraceParams p = p_in;
    if (!RT_SMEM_TLAS || p.tlas_smem_nodes < RT_SMEM_TLAS_MIN_NODES || p.tlas_smem_nodes > RT_SMEM_TLAS_MAX_NODES) p.tlas_smem_nodes = 0;
    const uint32_t tiles = ((p.width + 7u) >> 3) * ((p.local_rows + 3u) >> 2);
    if (tiles == 0) return 0;
    // the 16 counters and (RT_REGIONS == 2) the 2 x n_regions region fetch counters behind them: one memset
    if (cudaMemsetAsync(p.counters, 0, 4 * (16 + (RT_REGIONS == 2 ? 2 * (size_t)p.n_regions : 0)), st) != cudaSuccess) return -1;
#if RT_BOUNCE_ORDERED
    const bool fused = RT_FUSED_STAGES && (uint64_t)p.width * p.local_rows <= (uint64_t)RT_FUSED_MAX_PIXELS;
    if (p.bounces > 0 && !fused && cudaMemsetAsync(p.tile_mask, 0, sizeof(uint32_t) * (size_t)p.n_regions * TPR, st) != cudaSuccess) return -1;
#endif
    int n;
    if (stack_needed <= 64) n = launch_stack<64>(p, stats, sm_count, st);
    else if (stack_needed <= 160) n = launch_stack<160>(p, stats, sm_count, st);
    else return -2;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return n;
}

int launch_sbt_bound(const InstanceRec* inst, uint32_t n, uint32_t stride, unsigned long long* out, cudaStream_t st) {
    if (n == 0) return 0;
    k_sbt_bound<<<(n + 255) / 256, 256, 0, st>>>(inst, n, stride, out);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_flag_add(uint32_t* counter, cudaStream_t st) {
    k_flag_add<<<1, 1, 0, st>>>(counter);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}
int launch_flag_wait_ge(const uint32_t* counter, uint32_t target, int* error_flag, cudaStream_t st) {
    k_flag_wait_ge<<<1, 1, 0, st>>>(counter, target, error_flag);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_unpack_rows(const uint8_t* packed_all, uint32_t width, uint32_t height, uint32_t block_rows,
                       uint32_t part_count, uint8_t* out, cudaStream_t st) {
    const uint32_t bands = (height + block_rows - 1) / block_rows;
    const uint32_t rows_per_part = ((bands + part_count - 1) / part_count) * block_rows;
    dim3 grid((width + 255) / 256, height);
    k_unpack_rows<<<grid, 256, 0, st>>>(reinterpret_cast<const uchar4*>(packed_all), width, height, block_rows, part_count,
                                        rows_per_part, reinterpret_cast<uchar4*>(out));
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 1;
}

}  // namespace rt
