// trace.cu — vkCmdTraceRaysKHR(W,H,1) for sm_100a (reference dispatch: main.cpp:1349-1355).
//
// One launch = raygen prologue (main.cpp:1033-1052) -> two-level TLAS->BLAS while-while traversal
// (what traceRayEXT hands to the driver / RT cores; B200 has none, so this runs on the SMs) ->
// closest-hit / miss epilogue (main.cpp:1063-1066,1080-1091) -> rgba8 imageStore (main.cpp:1054),
// optionally followed in the same thread by one deterministic diffuse bounce.
//
//  * 64-byte nodes fetched as 4 x LDG.128 through the read-only path; 48-byte triangles as 3 x LDG.128.
//  * Box test: conservative slabs in FMA form (pad derived per ray/space from |origin| + |bounds|),
//    so a box is never culled when the exact-arithmetic triangle test could still report a hit.
//  * Triangle test: watertight Woop/Benthin/Wald 2013, plain IEEE mul/add (no contraction; the
//    file is compiled with -fmad=false), fp64 fallback on exact-zero edge functions. No culling:
//    the sample uses gl_RayFlagsOpaqueEXT only and TRIANGLE_FACING_CULL_DISABLE (main.cpp:852,1048).
//  * Closest hit = smallest t in (tmin, tmax); equal t resolved by lowest (instance, geometry,
//    primitive) so the result does not depend on BVH shape or traversal order.
//  * Each warp owns an 8x4-pixel tile so primary rays of a warp stay coherent.
#include <float.h>

#include "rt_device.cuh"

namespace rt {

namespace {

constexpr int TRACE_THREADS = 256;

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
    s.rdx = 1.0f / dx; s.rdy = 1.0f / dy; s.rdz = 1.0f / dz;
    s.cnx = -((s.px ? o.x + e : o.x - e) * s.rdx); s.cfx = -((s.px ? o.x - e : o.x + e) * s.rdx);
    s.cny = -((s.py ? o.y + e : o.y - e) * s.rdy); s.cfy = -((s.py ? o.y - e : o.y + e) * s.rdy);
    s.cnz = -((s.pz ? o.z + e : o.z - e) * s.rdz); s.cfz = -((s.pz ? o.z - e : o.z + e) * s.rdz);
}

// half = {lo.x lo.y lo.z hi.x} {hi.y hi.z ref height}
__device__ __forceinline__ bool slab_test(const Slab& s, const float4 h0, const float4 h1, float tmin, float tbest, float& tn) {
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
// paper (for d[kz] < 0) only negates U, V, W and det together, which leaves the hit decision and
// t, u, v bit-identical when there is no face culling, so it is omitted.
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

struct Hit {
    float t, u, v, w0;
    uint32_t inst_id, geo, prim;     // tie-break ids (0xFFFFFFFF = miss)
    uint32_t slot, tri;              // TLAS slot (sorted) and triangle index inside its BLAS
};

__device__ __forceinline__ float u01(uint32_t h) { return (float)(h >> 8) * 5.9604644775390625e-08f; }

__device__ __forceinline__ unsigned char unorm8(float c) {
    float v = c;
    if (!(v > 0.0f)) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    return (unsigned char)__float2int_rn(v * 255.0f);
}

template <bool STATS, int STACK>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(const TraceParams P) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = blockIdx.x * (TRACE_THREADS / 32) + (threadIdx.x >> 5);
    const uint32_t tiles_x = (P.width + 7u) >> 3;
    const uint32_t tiles_y = (P.local_rows + 3u) >> 2;
    if (warp_global >= tiles_x * tiles_y) return;
    const uint32_t x = (warp_global % tiles_x) * 8u + (lane & 7);
    const uint32_t lr = (warp_global / tiles_x) * 4u + (lane >> 3);
    const uint32_t band = lr / P.block_rows;
    const uint32_t y = (band * P.part_count + P.part_index) * P.block_rows + (lr - band * P.block_rows);
    const bool in_buffer = x < P.width && lr < P.local_rows;
    const bool valid = in_buffer && y < P.height;

    unsigned long long c_nodes = 0, c_tris = 0, c_insts = 0, c_ph = 0, c_sh = 0, c_sec = 0, c_edge = 0, c_prim = 0;
    float col[3] = {0.0f, 0.0f, 0.0f};
    Hit h1, h2;
    h1.t = P.tmax; h1.u = h1.v = h1.w0 = 0.0f; h1.inst_id = h1.geo = h1.prim = 0xFFFFFFFFu; h1.slot = h1.tri = 0;
    h2 = h1;

    if (valid) {
        // ---- raygen (main.cpp:1033-1046); aspect_x/aspect_y are computed once on the host (tanf) ----
        const float scx = (float)x + 0.5f, scy = (float)y + 0.5f;
        const float ndcx = scx / (float)P.width * 2.0f - 1.0f;
        const float ndcy = scy / (float)P.height * 2.0f - 1.0f;
        const float ax = ndcx * P.aspect_x, ay = ndcy * P.aspect_y;
        V3 o = {P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]};
        V3 d = {(ax * 1.0f + ay * 0.0f) + 0.0f, (ax * 0.0f + ay * -1.0f) + 0.0f, (ax * 0.0f + ay * 0.0f) + -1.0f};
        const uint32_t pixel = y * P.width + x;
        V3 prim_o = o, prim_d = d;

#pragma unroll 1
        for (uint32_t stage = 0; stage <= P.bounces; ++stage) {
            Hit best;
            best.t = P.tmax; best.u = best.v = best.w0 = 0.0f; best.inst_id = best.geo = best.prim = 0xFFFFFFFFu; best.slot = best.tri = 0;
            // ---- traceRayEXT(topLevelAS, Opaque, cullMask, ..., o, tmin, d, tmax) (main.cpp:1047-1052) ----
            int32_t stack[STACK];
            int sp = 0;
            stack[sp++] = REF_DONE;
            const BvhNode* nodes = P.tlas_nodes;
            const TriRec* tris = nullptr;
            Slab sl; Woop wp;
            slab_setup(sl, o, d, P.tlas_absmax[0], P.tlas_absmax[1], P.tlas_absmax[2]);
            wp.okx = wp.oky = wp.okz = wp.Sx = wp.Sy = wp.Sz = 0.0f; wp.z0 = wp.z1 = false;
            bool in_blas = false;
            uint32_t cur_slot = 0, cur_inst_id = 0;
            int32_t cur = P.tlas_root;
            for (;;) {
                while ((uint32_t)cur < (uint32_t)REF_SENTINEL_MIN) {           // internal node
                    const float4* n4 = reinterpret_cast<const float4*>(nodes + cur);
                    const float4 a0 = __ldg(n4), a1 = __ldg(n4 + 1), b0 = __ldg(n4 + 2), b1 = __ldg(n4 + 3);
                    if (STATS) ++c_nodes;
                    float t0, t1;
                    const bool hit0 = slab_test(sl, a0, a1, P.tmin, best.t, t0);
                    const bool hit1 = slab_test(sl, b0, b1, P.tmin, best.t, t1);
                    const int32_t r0 = __float_as_int(a1.z), r1 = __float_as_int(b1.z);
                    if (hit0 && hit1) {
                        const bool swap = t1 < t0;
                        stack[sp++] = swap ? r0 : r1;
                        cur = swap ? r1 : r0;
                    } else if (hit0) cur = r0;
                    else if (hit1) cur = r1;
                    else cur = stack[--sp];
                }
                if (cur < 0) {                                                   // leaf
                    const uint32_t first = leaf_first(cur), count = leaf_count(cur);
                    if (!in_blas) {
                        // TLAS leaf: one instance. Cull mask (main.cpp:851,1048), then enter its BLAS in object space.
                        const InstanceRec* R = P.instances + first;
                        const uint32_t cm = __ldg(&R->custom_mask);
                        const int32_t root = __ldg(&R->root);
                        if (((cm >> 24) & P.cull_mask) != 0u && root != REF_EMPTY) {
                            if (STATS) ++c_insts;
                            const float4* m4 = reinterpret_cast<const float4*>(R);
                            const float4 m0 = __ldg(m4), m1 = __ldg(m4 + 1), m2 = __ldg(m4 + 2);
                            const float w2o[12] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
                            const V3 oo = xform_point(w2o, o), od = xform_vec(w2o, d);
                            slab_setup(sl, oo, od, __ldg(&R->absmax[0]), __ldg(&R->absmax[1]), __ldg(&R->absmax[2]));
                            woop_setup(wp, oo, od);
                            nodes = R->nodes; tris = R->tris;
                            cur_slot = first; cur_inst_id = __ldg(&R->instance_id);
                            in_blas = true;
                            stack[sp++] = REF_POP_INSTANCE;
                            cur = root;
                        } else cur = stack[--sp];
                    } else {
                        for (uint32_t k = 0; k < count; ++k) {
                            const float4* t4 = reinterpret_cast<const float4*>(tris + first + k);
                            const float4 q0 = __ldg(t4), q1 = __ldg(t4 + 1), q2 = __ldg(t4 + 2);
                            if (STATS) ++c_tris;
                            float t, bu, bv, bw0;
                            if (woop_test(wp, q0, q1, q2, t, bu, bv, bw0) && t > P.tmin && t < P.tmax) {
                                const uint32_t geo = __float_as_uint(q2.y), prim = __float_as_uint(q2.z);
                                bool better = t < best.t;
                                if (t == best.t) {
                                    better = cur_inst_id != best.inst_id ? cur_inst_id < best.inst_id
                                             : (geo != best.geo ? geo < best.geo : prim < best.prim);
                                }
                                if (better) {
                                    best.t = t; best.u = bu; best.v = bv; best.w0 = bw0;
                                    best.inst_id = cur_inst_id; best.geo = geo; best.prim = prim;
                                    best.slot = cur_slot; best.tri = first + k;
                                }
                            }
                        }
                        cur = stack[--sp];
                    }
                } else if (cur == REF_POP_INSTANCE) {                            // back to world space
                    in_blas = false;
                    nodes = P.tlas_nodes;
                    slab_setup(sl, o, d, P.tlas_absmax[0], P.tlas_absmax[1], P.tlas_absmax[2]);
                    cur = stack[--sp];
                } else if (cur == REF_EMPTY) {
                    cur = stack[--sp];
                } else break;                                                     // REF_DONE
            }

            if (stage == 0) { h1 = best; if (STATS) ++c_prim; } else { h2 = best; if (STATS) ++c_sec; }
            float sc[3];
            const bool hit = best.inst_id != 0xFFFFFFFFu;
            const InstanceRec* R = P.instances + best.slot;
            uint32_t custom = 0;
            if (hit) {
                // ---- closest-hit (main.cpp:1080-1091) with the SBT rule of main.cpp:1260-1262 ----
                const uint32_t cm = __ldg(&R->custom_mask), sf = __ldg(&R->sbt_flags);
                custom = cm & 0xFFFFFFu;
                if (best.prim == 1u && best.inst_id == 1u && custom == 100u && best.geo == 1u) {
                    sc[0] = 1.0f - best.u - best.v; sc[1] = best.u; sc[2] = best.v;
                } else {
                    const uint32_t rec = (sf & 0xFFFFFFu) + best.geo * P.sbt_stride + P.sbt_offset;
                    if (rec < P.n_records) { sc[0] = __ldg(P.hit_records + 3 * rec); sc[1] = __ldg(P.hit_records + 3 * rec + 1); sc[2] = __ldg(P.hit_records + 3 * rec + 2); }
                    else { sc[0] = sc[1] = sc[2] = 0.0f; }
                }
                if (STATS) { if (stage == 0) { ++c_ph; if (fminf(fminf(best.u, best.v), best.w0) < 9.5367431640625e-07f) ++c_edge; } else ++c_sh; }
            } else {
                sc[0] = P.miss[0]; sc[1] = P.miss[1]; sc[2] = P.miss[2];          // miss shader (main.cpp:1063-1066)
            }
            rt_hit* hout = stage == 0 ? P.primary_hits : P.secondary_hits;
            if (hout) {
                rt_hit r;
                r.instance_id = best.inst_id; r.geometry_index = best.geo; r.primitive_id = best.prim;
                r.custom_index = hit ? custom : 0xFFFFFFFFu; r.t = hit ? best.t : P.tmax; r.u = best.u; r.v = best.v;
                hout[(size_t)lr * P.width + x] = r;
            }
            if (stage == 0) { col[0] = sc[0]; col[1] = sc[1]; col[2] = sc[2]; }
            else { col[0] = 0.5f * col[0] + 0.5f * sc[0]; col[1] = 0.5f * col[1] + 0.5f * sc[1]; col[2] = 0.5f * col[2] + 0.5f * sc[2]; }
            if (!hit || stage == P.bounces) break;

            // ---- deterministic diffuse bounce (our definition; the reference's recursion depth is 1) ----
            {
                const V3 p = {prim_o.x + best.t * prim_d.x, prim_o.y + best.t * prim_d.y, prim_o.z + best.t * prim_d.z};
                const float4* m4 = reinterpret_cast<const float4*>(R);
                const float4 m0 = __ldg(m4), m1 = __ldg(m4 + 1), m2 = __ldg(m4 + 2);
                const float w2o[12] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
                const float4* t4 = reinterpret_cast<const float4*>(R->tris + best.tri);
                const float4 q0 = __ldg(t4), q1 = __ldg(t4 + 1), q2 = __ldg(t4 + 2);
                const V3 e1 = {q0.w - q0.x, q1.x - q0.y, q1.y - q0.z};
                const V3 e2 = {q1.z - q0.x, q1.w - q0.y, q2.x - q0.z};
                V3 n = xform_normal(w2o, cross3(e1, e2));
                const float l2 = dot3(n, n);
                if (l2 > 0.0f && l2 < INFINITY) { const float l = sqrtf(l2); n = {n.x / l, n.y / l, n.z / l}; }
                else { const float dl = sqrtf(dot3(prim_d, prim_d)); n = {-prim_d.x / dl, -prim_d.y / dl, -prim_d.z / dl}; }
                if (dot3(n, prim_d) > 0.0f) n = {-n.x, -n.y, -n.z};
                uint32_t h = pcg_hash(pixel + pcg_hash(P.bounce_seed + 0x9E3779B9u));
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
                o = {p.x + n.x * eps, p.y + n.y * eps, p.z + n.z * eps};
                d = dir;
            }
        }
    }

    if (in_buffer) {
        const size_t idx = (size_t)lr * P.width + x;
        uchar4 px = valid ? make_uchar4(unorm8(col[0]), unorm8(col[1]), unorm8(col[2]), 0) : make_uchar4(0, 0, 0, 0);
        reinterpret_cast<uchar4*>(P.rgba)[idx] = px;                              // imageStore(vec4(hitValue, 0.0)), main.cpp:1054
        if (!valid) {
            rt_hit r; r.instance_id = r.geometry_index = r.primitive_id = r.custom_index = 0xFFFFFFFFu; r.t = P.tmax; r.u = r.v = 0.0f;
            if (P.primary_hits) P.primary_hits[idx] = r;
            if (P.secondary_hits) P.secondary_hits[idx] = r;
        } else if (P.secondary_hits && h1.inst_id == 0xFFFFFFFFu) {
            rt_hit r; r.instance_id = r.geometry_index = r.primitive_id = r.custom_index = 0xFFFFFFFFu; r.t = P.tmax; r.u = r.v = 0.0f;
            P.secondary_hits[idx] = r;
        } else if (P.secondary_hits && P.bounces == 0) {
            rt_hit r; r.instance_id = r.geometry_index = r.primitive_id = r.custom_index = 0xFFFFFFFFu; r.t = P.tmax; r.u = r.v = 0.0f;
            P.secondary_hits[idx] = r;
        }
    }
    if (STATS && P.stats) {
        unsigned long long v[8] = {c_prim, c_sec, c_nodes, c_tris, c_insts, c_ph, c_sh, c_edge};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(P.stats + k, v[k]);
        }
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

}  // namespace

int launch_trace(const TraceParams& p, bool stats, int stack_needed, int /*sm_count*/, cudaStream_t st) {
    const uint32_t tiles = ((p.width + 7u) >> 3) * ((p.local_rows + 3u) >> 2);
    if (tiles == 0) return 0;
    const uint32_t blocks = (tiles + (TRACE_THREADS / 32) - 1) / (TRACE_THREADS / 32);
    if (stack_needed <= 64) {
        if (stats) k_trace<true, 64><<<blocks, TRACE_THREADS, 0, st>>>(p);
        else k_trace<false, 64><<<blocks, TRACE_THREADS, 0, st>>>(p);
    } else if (stack_needed <= 160) {
        if (stats) k_trace<true, 160><<<blocks, TRACE_THREADS, 0, st>>>(p);
        else k_trace<false, 160><<<blocks, TRACE_THREADS, 0, st>>>(p);
    } else return -2;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 1;
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
