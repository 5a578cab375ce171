// wide_host.cpp — CPU emulation of the product's binary->8-wide collapse and of its wide-node traversal,
// compiled from the SAME header the CUDA kernels use (build-up-phase_b200/csrc/wide_bvh.cuh). Test
// infrastructure only: lets `pytest -m "not gpu"` check the collapse / quantisation / bit tricks against
// brute force without a GPU. Nothing in the product links this.
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../build-up-phase_b200/csrc/wide_bvh.cuh"

using namespace rt;

struct HostFetch {
    const uint32_t* nodes;      // n x 16 words: two 32-B halves {lo.xyz, hi.xyz, ref, height}
    void operator()(int32_t ref, WChild* out) const {
        for (int h = 0; h < 2; ++h) {
            const uint32_t* w = nodes + 16 * (size_t)ref + 8 * h;
            for (int k = 0; k < 3; ++k) { memcpy(&out[h].lo[k], w + k, 4); memcpy(&out[h].hi[k], w + 3 + k, 4); }
            out[h].ref = (int32_t)w[6];
        }
    }
};

extern "C" {

// Level-synchronous collapse exactly like k_widen (sequential "threads"). perm_out[wide position] = sorted position.
int wide_host_build(const uint32_t* bnodes, int32_t root, const float* lo, const float* hi, uint32_t n_prims,
                    uint32_t* wnodes_out, uint32_t cap, uint32_t* perm_out, uint32_t* n_nodes_out, uint32_t* depth_out) {
    std::vector<int32_t> src(cap);
    WNode* wn = reinterpret_cast<WNode*>(wnodes_out);
    uint32_t begin = 0, end = 1, level = 0, cursor = 0;
    src[0] = root;
    HostFetch fetch{bnodes};
    while (begin < end) {
        uint32_t next = 0;
        for (uint32_t wi = begin; wi < end; ++wi) {
            WNode node; WideEmit em;
            widen_one(src[wi], lo, hi, fetch, node, em);
            if (em.n_internal) {
                const uint32_t base = end + next;
                next += em.n_internal;
                if (base + em.n_internal > cap) return -1;
                node.child_base = base;
                for (int k = 0; k < em.n_internal; ++k) src[base + k] = em.internal_ref[k];
            }
            if (em.n_prims) {
                node.prim_base = cursor;
                for (int l = 0; l < em.n_leaf; ++l)
                    for (uint32_t k = 0; k < em.leaf_count[l]; ++k) perm_out[cursor++] = em.leaf_first[l] + k;
            }
            wn[wi] = node;
        }
        begin = end; end += next; ++level;
    }
    if (cursor != n_prims) return -2;
    *n_nodes_out = end; *depth_out = level;
    return 0;
}

// Walks the wide BVH for each ray like the trace kernel's per-lane loop, WITHOUT distance culling (tbest = tmax),
// and reports whether primitive expect[r] (wide position) gets tested. tested_out[r] = number of primitives tested.
// Returns the number of rays whose expected primitive was never reached.
int wide_host_reach(const uint32_t* wnodes, uint32_t root, const float* absmax, const float* rays /* n x 6: o, d */,
                    const int32_t* expect, uint32_t n_rays, float tmin, float tmax, uint32_t* tested_out, uint32_t* nodes_out) {
    const WNode* wn = reinterpret_cast<const WNode*>(wnodes);
    int missing = 0;
    for (uint32_t r = 0; r < n_rays; ++r) {
        RayBox rb;
        raybox_setup(rb, rays + 6 * r, rays + 6 * r + 3, absmax[0], absmax[1], absmax[2]);
        struct G { uint32_t x, y; };
        std::vector<G> stack;
        G ng{root, 0x80000000u}, tg{0u, 0u};
        bool found = expect[r] < 0;
        uint32_t tested = 0, visited = 0;
        bool traversing = true;
        while (traversing) {
            if (ng.y > 0x00FFFFFFu) {
                const uint32_t bit = 31u - (uint32_t)__builtin_clz(ng.y);
                const uint32_t imask = ng.y & 0xFFu;
                ng.y &= ~(1u << bit);
                if (ng.y > 0x00FFFFFFu) stack.push_back(ng);
                const uint32_t slot = (bit - 24u) ^ rb.oct;
                const uint32_t child = ng.x + (uint32_t)__builtin_popcount(imask & ~(0xFFFFFFFFu << slot));
                const WWord* w = reinterpret_cast<const WWord*>(wn + child);
                ++visited;
                uint32_t inner, prims;
                wide_node_hits(rb, w[0], w[1], w[2], w[3], w[4], tmin, tmax, inner, prims);
                ng = G{w[1].x, inner | (w[0].w >> 24)};
                tg = G{child, prims};
            }
            if (tg.y) {
                const WNode& N = wn[tg.x];
                while (tg.y) {
                    const uint32_t bit = 31u - (uint32_t)__builtin_clz(tg.y);
                    tg.y &= ~(1u << bit);
                    ++tested;
                    const uint32_t pidx = N.prim_base + (uint32_t)__builtin_popcount(N.prim_valid & ~(0xFFFFFFFFu << bit));
                    if ((int32_t)pidx == expect[r]) found = true;
                }
            }
            if (ng.y <= 0x00FFFFFFu) {
                if (stack.empty()) traversing = false;
                else { const G e = stack.back(); stack.pop_back(); if (e.y > 0x00FFFFFFu) ng = e; else { tg = e; ng = G{0u, 0u}; } }
            }
        }
        if (!found) ++missing;
        tested_out[r] = tested; nodes_out[r] = visited;
    }
    return missing;
}

}  // extern "C"
