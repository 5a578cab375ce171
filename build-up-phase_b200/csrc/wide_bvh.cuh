// wide_bvh.cuh — the 8-wide compressed BVH node the trace kernel walks, and the per-node step of the
// collapse that derives it from the binary LBVH (lbvh_build.cu: Karras hierarchy + atomic refit).
//
// Why a second structure: the binary LBVH needs ~29 DEPENDENT 64-byte node fetches per ray; on B200 the
// trace kernel is bound by exactly that chain (long-scoreboard stalls, L1 request rate), not by HBM.
// Collapsing three binary levels into one 80-byte node with 8-bit child boxes cuts the chain and the L1
// bytes per ray by ~2.5x. The box test only has to be CONSERVATIVE (hit results are defined by the exact
// watertight triangle test, see trace.cu), so quantised boxes cannot change any hit id.
//
// Layout (80 B, five 16-byte words; same idea as Ylitie/Karras/Laine 2017, own implementation):
//   w0  px py pz | ex ey ez imask      origin of the node's quantisation grid; biased fp32 exponents of the
//                                      grid steps (step_k = 2^(e_k-127)); bit s of imask = slot s is internal
//   w1  child_base prim_base           internal children are consecutive wide nodes child_base + rank(slot among internal);
//       prim_valid spare               leaf slot s owns bits 3s..3s+2 of prim_valid, one bit per primitive it holds (1..3);
//                                      the primitives of a node are consecutive: index = prim_base + rank(bit in prim_valid)
//   w2  qlo.x[8] qlo.y[8]              child boxes: lo = p + qlo*step (rounded down), hi = p + qhi*step (rounded up)
//   w3  qlo.z[8] qhi.x[8]
//   w4  qhi.y[8] qhi.z[8]
// Slots are assigned so that slot bit 2/1/0 set means the child lies towards +x/+y/+z of the node centre:
// visiting hit slots in descending (slot XOR octant) order is then approximately front to back for every ray
// octant, without sorting distances. Empty slots carry an inverted box (qlo 255, qhi 0) and no prim_valid bits.
//
// This header is also compiled as plain C++ by tests/wide_host.cpp (host emulation of the collapse and of
// the node test) so that the logic is checked on CPU against brute force without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

struct alignas(16) WNode {
    float px, py, pz;
    uint8_t ex, ey, ez, imask;
    uint32_t child_base, prim_base;
    uint32_t prim_valid, spare;
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(WNode) == 80, "WNode must be 80 B");

constexpr int WIDE = 8;
constexpr int WIDE_TRI_LEAF_MAX = 3;     // 8 slots x 3 = the 24 primitive bits of a hit mask
constexpr uint32_t WNODE_NONE = 0xFFFFFFFFu;

// Binary node half as produced by the refit: {lo.xyz, hi.xyz, ref, height}; refs: >= 0 internal (segment
// relative), < 0 leaf ~((first << 3) | (count - 1)).
struct WChild {
    float lo[3], hi[3];
    int32_t ref;
};

RT_HD float wide_half_area(const WChild& c) {
    const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
    if (!(dx >= 0.0f) || !(dy >= 0.0f) || !(dz >= 0.0f)) return -1.0f;     // inverted (inactive) box
    return (dx * dy + dy * dz) + dz * dx;
}

// smallest biased exponent E in [1, 254] with 2^(E-127) * 255 >= extent
RT_HD uint32_t wide_step_exponent(float extent) {
    if (!(extent > 0.0f)) return 1u;
    const double need = (double)extent / 255.0;
    int k;
    const double m = frexp(need, &k);            // need = m * 2^k, m in [0.5, 1)
    int e = (m == 0.5) ? k - 1 : k;              // 2^e >= need
    int E = e + 127;
    if (E < 1) E = 1;
    if (E > 254) E = 254;
    return (uint32_t)E;
}

RT_HD double wide_step_value(uint32_t E) { return ldexp(1.0, (int)E - 127); }

// Result of collapsing one wide node: which binary subtrees become its internal children (slot order) and
// which primitive ranges its leaf slots hold (slot order). The caller allocates child_base / prim_base.
struct WideEmit {
    int n_internal;
    int32_t internal_ref[WIDE];      // binary refs (segment relative) of the internal children, in slot order
    int n_leaf;
    uint32_t leaf_first[WIDE];       // segment-relative first primitive of each leaf slot, in slot order
    uint32_t leaf_count[WIDE];
    uint32_t n_prims;                // sum of leaf_count
};

// Expands binary node `src` (or wraps the single leaf `src` of a tiny segment) into up to 8 children by
// repeatedly opening the internal child with the largest surface area, assigns slots, quantises.
// fetch(ref, WChild out[2]) reads the two halves of binary internal node `ref`.
template <class Fetch>
RT_HD void widen_one(int32_t src, const float* src_lo, const float* src_hi, Fetch fetch, WNode& out, WideEmit& em) {
    WChild ch[WIDE];
    int n = 0;
    if (src < 0) {                                   // the whole segment is one leaf
        for (int k = 0; k < 3; ++k) { ch[0].lo[k] = src_lo[k]; ch[0].hi[k] = src_hi[k]; }
        ch[0].ref = src;
        n = 1;
    } else {
        fetch(src, ch);
        n = 2;
        while (n < WIDE) {
            int best = -1; float best_area = -2.0f;
            for (int i = 0; i < n; ++i) {
                if (ch[i].ref < 0) continue;
                const float a = wide_half_area(ch[i]);
                if (a > best_area) { best_area = a; best = i; }
            }
            if (best < 0) break;
            WChild two[2];
            fetch(ch[best].ref, two);
            ch[best] = two[0];
            ch[n++] = two[1];
        }
    }
    // ---- node box = union of the valid child boxes ----
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool valid[WIDE];
    for (int i = 0; i < n; ++i) {
        valid[i] = ch[i].lo[0] <= ch[i].hi[0] && ch[i].lo[1] <= ch[i].hi[1] && ch[i].lo[2] <= ch[i].hi[2];
        if (!valid[i]) continue;
        for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], ch[i].lo[k]); hi[k] = fmaxf(hi[k], ch[i].hi[k]); }
    }
    if (!(lo[0] <= hi[0])) { for (int k = 0; k < 3; ++k) { lo[k] = 0.0f; hi[k] = 0.0f; } }
    // ---- greedy slot assignment: slot bit set <=> child towards the positive side of that axis ----
    float cen[3] = {0.5f * lo[0] + 0.5f * hi[0], 0.5f * lo[1] + 0.5f * hi[1], 0.5f * lo[2] + 0.5f * hi[2]};
    float off[WIDE][3];
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) off[i][k] = valid[i] ? (0.5f * ch[i].lo[k] + 0.5f * ch[i].hi[k]) - cen[k] : 0.0f;
    int slot_of[WIDE], child_in[WIDE];
    for (int i = 0; i < WIDE; ++i) { slot_of[i] = -1; child_in[i] = -1; }
    for (int round = 0; round < n; ++round) {
        int bi = -1, bs = -1; float bc = -INFINITY;
        for (int i = 0; i < n; ++i) {
            if (slot_of[i] >= 0) continue;
            for (int s = 0; s < WIDE; ++s) {
                if (child_in[s] >= 0) continue;
                const float c = ((s & 4) ? off[i][0] : -off[i][0]) + ((s & 2) ? off[i][1] : -off[i][1]) + ((s & 1) ? off[i][2] : -off[i][2]);
                if (c > bc) { bc = c; bi = i; bs = s; }
            }
        }
        if (bi < 0) {                                 // only NaN costs left: fill the first free pair
            for (int i = 0; i < n && bi < 0; ++i) if (slot_of[i] < 0) bi = i;
            for (int s = 0; s < WIDE && bs < 0; ++s) if (child_in[s] < 0) bs = s;
        }
        slot_of[bi] = bs; child_in[bs] = bi;
    }
    // ---- header + quantisation ----
    const uint32_t E[3] = {wide_step_exponent(hi[0] - lo[0]), wide_step_exponent(hi[1] - lo[1]), wide_step_exponent(hi[2] - lo[2])};
    out.px = lo[0]; out.py = lo[1]; out.pz = lo[2];
    out.ex = (uint8_t)E[0]; out.ey = (uint8_t)E[1]; out.ez = (uint8_t)E[2];
    out.imask = 0; out.child_base = 0; out.prim_base = 0; out.prim_valid = 0; out.spare = 0;
    em.n_internal = 0; em.n_leaf = 0; em.n_prims = 0;
    uint8_t* qlo[3] = {out.qlox, out.qloy, out.qloz};
    uint8_t* qhi[3] = {out.qhix, out.qhiy, out.qhiz};
    for (int s = 0; s < WIDE; ++s) {
        const int i = child_in[s];
        if (i < 0 || !valid[i]) {                    // empty slot, or a subtree without any active primitive
            for (int k = 0; k < 3; ++k) { qlo[k][s] = 255; qhi[k][s] = 0; }
            continue;
        }
        for (int k = 0; k < 3; ++k) {
            const double inv = ldexp(1.0, 127 - (int)E[k]);                       // 1 / step, exact
            double a = floor(((double)ch[i].lo[k] - (double)lo[k]) * inv);
            double b = ceil(((double)ch[i].hi[k] - (double)lo[k]) * inv);
            if (a < 0.0) a = 0.0;
            if (a > 255.0) a = 255.0;
            if (b > 255.0) b = 255.0;
            if (b < 0.0) b = 0.0;
            qlo[k][s] = (uint8_t)a; qhi[k][s] = (uint8_t)b;
        }
        if (ch[i].ref >= 0) {
            out.imask |= (uint8_t)(1u << s);
            em.internal_ref[em.n_internal++] = ch[i].ref;
        } else {
            const uint32_t u = (uint32_t)~ch[i].ref;
            const uint32_t first = u >> 3, count = (u & 7u) + 1u;             // count <= 3 by construction of the binary tree
            out.prim_valid |= ((1u << count) - 1u) << (3 * s);
            em.leaf_first[em.n_leaf] = first; em.leaf_count[em.n_leaf] = count; ++em.n_leaf;
            em.n_prims += count;
        }
    }
}


// ---- the few intrinsics the node test needs, with host stand-ins for the CPU emulation in tests/ ----
#ifdef __CUDA_ARCH__
typedef uint4 WWord;
RT_HD float rt_u2f(uint32_t u) { return __uint_as_float(u); }
RT_HD uint32_t rt_prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
RT_HD float rt_fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
RT_HD float rt_fma_rd(float a, float b, float c) { return __fmaf_rd(a, b, c); }
RT_HD float rt_fma_ru(float a, float b, float c) { return __fmaf_ru(a, b, c); }
RT_HD float rt_rcp(float x) { float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
#else
#ifdef __CUDACC__
typedef uint4 WWord;
#else
struct WWord { uint32_t x, y, z, w; };
#endif
inline float rt_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t rt_prmt(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xFu;
        const uint32_t byte = (uint32_t)(v >> (8 * (sel & 7u))) & 0xFFu;     // like __byte_perm: selector bit 3 is ignored
        r |= byte << (8 * i);
    }
    return r;
}
inline float rt_rcp(float x) { return 1.0f / x; }
inline float rt_fma_rn(float a, float b, float c) { return (float)fma((double)a, (double)b, (double)c); }   // double-rounding is irrelevant for the emulation
inline float rt_fma_rd(float a, float b, float c) { const double e = fma((double)a, (double)b, (double)c); float f = (float)e; return (double)f > e ? nextafterf(f, -INFINITY) : f; }
inline float rt_fma_ru(float a, float b, float c) { const double e = fma((double)a, (double)b, (double)c); float f = (float)e; return (double)f < e ? nextafterf(f, INFINITY) : f; }
#endif

// Per-ray, per-space constants of the conservative box test.
struct RayBox {
    float idx, idy, idz;     // 1/d (zero components replaced by +-1e-20)
    float cnx, cny, cnz;     // -(o +- e) * id for the near planes
    float cfx, cfy, cfz;     // -(o -+ e) * id for the far planes
    uint32_t oct;            // bit 2/1/0 set <=> d[0]/d[1]/d[2] >= 0 ("octant inverse": XOR turns slot numbers into priorities)
};

RT_HD void raybox_setup(RayBox& s, const float* o, const float* d, float ax, float ay, float az) {
    const float M = fmaxf(fmaxf(ax + fabsf(o[0]), ay + fabsf(o[1])), az + fabsf(o[2]));
    const float e = M * 1.9073486328125e-06f;   // 2^-19 relative spatial pad
    const float dx = fabsf(d[0]) < 1e-20f ? copysignf(1e-20f, d[0]) : d[0];
    const float dy = fabsf(d[1]) < 1e-20f ? copysignf(1e-20f, d[1]) : d[1];
    const float dz = fabsf(d[2]) < 1e-20f ? copysignf(1e-20f, d[2]) : d[2];
    const bool px = dx > 0.0f, py = dy > 0.0f, pz = dz > 0.0f;
    s.oct = (px ? 4u : 0u) | (py ? 2u : 0u) | (pz ? 1u : 0u);
    // the box test only has to be conservative: a 1-ulp-class reciprocal (MUFU.RCP, rel. error 2^-22) is covered 4x by the
    // 2^-19 pad (|plane - o| <= 2M, so the induced error in t is below M*|id|*2^-21), and saves three IEEE divisions
    s.idx = rt_rcp(dx); s.idy = rt_rcp(dy); s.idz = rt_rcp(dz);
    s.cnx = -((px ? o[0] + e : o[0] - e) * s.idx); s.cfx = -((px ? o[0] - e : o[0] + e) * s.idx);
    s.cny = -((py ? o[1] + e : o[1] - e) * s.idy); s.cfy = -((py ? o[1] - e : o[1] + e) * s.idy);
    s.cnz = -((pz ? o[2] + e : o[2] - e) * s.idz); s.cfz = -((pz ? o[2] - e : o[2] + e) * s.idz);
}

// byte k of `word` -> float 32768 + byte (the byte lands in mantissa bits 8..15 of 32768.0f = 0x47000000)
template <int K>
RT_HD float q2f(uint32_t word) { return rt_u2f(rt_prmt(word, 0x47000000u, 0x7404u | (K << 4))); }

// Tests the 8 children of one wide node. inner: bits 31..24, bit 24 + (slot ^ oct) set for every hit internal slot
// (highest bit = nearest in octant order). prims: the prim_valid bits of the hit leaf slots.
RT_HD void wide_node_hits(const RayBox& rb, const WWord n0, const WWord n1, const WWord n2, const WWord n3, const WWord n4,
                          float tmin, float tbest, uint32_t& inner, uint32_t& prims) {
    const uint32_t ew = n0.w;
    const float ax = rt_u2f((ew & 0xFFu) << 23) * rb.idx;
    const float ay = rt_u2f(((ew >> 8) & 0xFFu) << 23) * rb.idy;
    const float az = rt_u2f(((ew >> 16) & 0xFFu) << 23) * rb.idz;
    const float px = rt_u2f(n0.x), py = rt_u2f(n0.y), pz = rt_u2f(n0.z);
    // t(q) = q*a + (p - o -+ e)*id; the "- 32768*a" folds the magic-number offset of q2f into the constant and is
    // rounded OUTWARDS (near planes down, far planes up) so the dequantisation can only enlarge the box.
    const float bnx = rt_fma_rd(-32768.0f, ax, rt_fma_rn(px, rb.idx, rb.cnx)), bfx = rt_fma_ru(-32768.0f, ax, rt_fma_rn(px, rb.idx, rb.cfx));
    const float bny = rt_fma_rd(-32768.0f, ay, rt_fma_rn(py, rb.idy, rb.cny)), bfy = rt_fma_ru(-32768.0f, ay, rt_fma_rn(py, rb.idy, rb.cfy));
    const float bnz = rt_fma_rd(-32768.0f, az, rt_fma_rn(pz, rb.idz, rb.cnz)), bfz = rt_fma_ru(-32768.0f, az, rt_fma_rn(pz, rb.idz, rb.cfz));
    const bool dpx = (rb.oct & 4u) != 0u, dpy = (rb.oct & 2u) != 0u, dpz = (rb.oct & 1u) != 0u;
    // n2 = {qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7]}  n3 = {qlo.z.., qlo.z.., qhi.x.., qhi.x..}  n4 = {qhi.y.., qhi.y.., qhi.z.., qhi.z..}
    uint32_t hit8 = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t lox = h ? n2.y : n2.x, loy = h ? n2.w : n2.z, loz = h ? n3.y : n3.x;
        const uint32_t hix = h ? n3.w : n3.z, hiy = h ? n4.y : n4.x, hiz = h ? n4.w : n4.z;
        const uint32_t nx = dpx ? lox : hix, fx = dpx ? hix : lox;
        const uint32_t ny = dpy ? loy : hiy, fy = dpy ? hiy : loy;
        const uint32_t nz = dpz ? loz : hiz, fz = dpz ? hiz : loz;
#define RT_WIDE_CHILD(J)                                                                                        \
        {                                                                                                       \
            const float tnx = rt_fma_rn(q2f<J>(nx), ax, bnx), tfx = rt_fma_rn(q2f<J>(fx), ax, bfx);             \
            const float tny = rt_fma_rn(q2f<J>(ny), ay, bny), tfy = rt_fma_rn(q2f<J>(fy), ay, bfy);             \
            const float tnz = rt_fma_rn(q2f<J>(nz), az, bnz), tfz = rt_fma_rn(q2f<J>(fz), az, bfz);             \
            const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                          \
            const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tbest));                                         \
            if (tn <= tf) hit8 |= 1u << (4 * h + J);                                                            \
        }
        RT_WIDE_CHILD(0) RT_WIDE_CHILD(1) RT_WIDE_CHILD(2) RT_WIDE_CHILD(3)
#undef RT_WIDE_CHILD
    }
    const uint32_t imask = ew >> 24;
    // internal slots: move bit s to bit s ^ oct (three conditional swaps), then up to bits 31..24
    uint32_t ih = hit8 & imask;
    { const uint32_t t = ((ih & 0x55u) << 1) | ((ih >> 1) & 0x55u); ih = (rb.oct & 1u) ? t : ih; }
    { const uint32_t t = ((ih & 0x33u) << 2) | ((ih >> 2) & 0x33u); ih = (rb.oct & 2u) ? t : ih; }
    { const uint32_t t = ((ih & 0x0Fu) << 4) | ((ih >> 4) & 0x0Fu); ih = (rb.oct & 4u) ? t : ih; }
    inner = ih << 24;
    // leaf slots: spread bit s to bit 3s, widen to the slot's three primitive bits, keep the primitives that exist
    uint32_t x = hit8 & ~imask;
    x = (x | (x << 8)) & 0x00F00Fu;
    x = (x | (x << 4)) & 0x0C30C3u;
    x = (x | (x << 2)) & 0x249249u;
    prims = (x * 7u) & n1.z;
}

}  // namespace rt
