// rt_host.h — host-side object definitions shared by the translation units that implement the C ABI
// (rtcore_api.cu: contexts, builds, dispatch; rt_group.cu: the multi-GPU render group).
#pragma once
#include <float.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "rt_internal.h"

using namespace rt;   // internal header: only the two ABI translation units include it

struct rt_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr;                           // device->host copies of finished row chunks (overlaps the next chunk's trace)
    cudaStream_t aux_stream = nullptr;                            // second compute stream: odd row chunks of a frame (their CTAs fill the even chunks' kernel tails)
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    int trace_chunks = 1;                                         // row chunks of a DEVICE-output trace (RTCORE_TRACE_CHUNKS); measured 3.56/3.55/3.67/3.67/4.02/4.07 ms for
                                                                  // 1/2/3/4/6/8 chunks: tail filling only pays back the extra launches, so the default is one launch
    cudaEvent_t chunk_ev[8]{};
    cudaEvent_t host_ev = nullptr;                                // recorded on copy_stream behind the copies of an ASYNC host-output trace (rt_host_frame_wait)
    bool host_pending = false;                                    // such a trace is in flight: the staging buffer and the error word are not to be reused yet
    int* h_err_pinned = nullptr;                                  // pinned: the device watchdog word of that trace, copied behind its frame
    int e2e_chunks = 3;                                            // row chunks of a HOST-output trace (RTCORE_E2E_CHUNKS): chunk c is copied device->host
                                                                   // while chunk c + 1 is traced; the chunks alternate over two compute streams (the next
                                                                   // chunk's CTAs fill the previous chunk's kernel tails) and shrink towards the end of the
                                                                   // frame (3 : 2 : 1), since only the last copy is exposed. Measured on B200, inst10m 4K:
                                                                   // 1 chunk 4.19 ms; equal chunks 2/3/4: 3.88/3.90/3.85; shrinking 3/4/5/6: 3.59/3.72/3.81/3.85
    cudaDeviceProp prop{};
    std::string err;
    // shader data
    float* d_hit_records = nullptr; uint32_t n_records = 0;
    uint4* d_anyhit = nullptr; uint32_t n_anyhit = 0;            // any-hit records {kind, log2_res, flags, first mask word} followed by the mask words (one allocation)
    bool own_hit_records = true;                                  // false for the render group's internal second context (it borrows the user context's table)
    std::vector<float> miss = {0.0f, 0.0f, 0.2f};                 // miss records, 3 floats each; record 0 = main.cpp:1065
    rt_ray_params rp = {0.0f, 100.0f, 0xffu, 0u, 1u, 1u, RT_RAY_FLAG_OPAQUE, 0u};   // main.cpp:1047-1052
    // grow-only device buffers
    void* scratch = nullptr; size_t scratch_cap = 0;
    void* fb = nullptr; size_t fb_cap = 0;
    void* hits1 = nullptr; size_t hits1_cap = 0;
    void* hits2 = nullptr; size_t hits2_cap = 0;
    void* queue = nullptr; size_t queue_cap = 0;       // bounce queue of the two-stage wavefront
    uint32_t* qflags = nullptr; size_t qflags_cap = 0; // per-entry publication flags of the fused launch (hold the epoch of the launch that wrote the entry)
    uint32_t trace_epoch = 0;
    void* d_counters = nullptr; size_t counters_cap = 0;   // per launch: 16 ray-fetch / queue counters + 2 x n_regions region fetch counters of the persistent trace kernels
    unsigned long long* d_stats = nullptr;
    int* d_error = nullptr;
    cudaEvent_t ev[8]{};
    rt_build_timing timing{};
    size_t last_scratch_need = 0;                      // scratch bytes the most recent build asked for (rt_last_build_scratch_bytes)
    float last_trace_ms = 0.0f;
    rt_trace_stats last_stats{};
    uint64_t launches = 0;
    // debug view of the last BLAS build's sorted keys (lives in scratch until the next build)
    const uint64_t* dbg_keys = nullptr; const uint32_t* dbg_vals = nullptr; uint32_t dbg_n = 0; int dbg_vb = 0;
};

struct BlasStorage {
    int refs = 0;
    void* dev = nullptr;            // nodes[N] | tris[N] | records[n_blas]
    size_t bytes = 0;
    BvhNode* nodes = nullptr; TriRec* tris = nullptr; BlasRecord* records = nullptr;
    uint32_t n_tris = 0, n_blas = 0;
    uint32_t n_node_slots = 0;      // node slots in `nodes` (n_tris after a build; the live-node count after rt_compact_blas)
    // retained by builds with RT_BUILD_ALLOW_UPDATE / RT_BUILD_ALLOW_COMPACTION: the sorted Morton records of the last full build
    // (packed `key << key_vb | id`, or keys + ids when key_vb == 0) - they ARE the tree topology
    uint64_t* keys = nullptr; uint32_t* vals = nullptr; int key_vb = 0;
    uint32_t build_flags = 0;
    bool compacted = false;
    std::vector<BlasRecord> host_recs;   // host copies of records[]
};
struct rt_blas {
    BlasStorage* st = nullptr;
    uint32_t index = 0;
    BlasRecord& r() const { return st->host_recs[index]; }     // host copy of this BLAS's record (shared by all handles of a batch: compaction updates it once)
};
struct rt_tlas {
    void* dev = nullptr; size_t bytes = 0;
    InstanceRec* inst = nullptr; BvhNode* nodes = nullptr;
    uint32_t n = 0;
    int32_t root = REF_EMPTY; uint32_t height = 0;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int32_t max_sbt_plus_geo = 0, max_sbt = 0, max_geo = 0, max_blas_height = 0;
    uint32_t bound_stride = 1; uint64_t bound = 0;     // cached max_i(sbt_i + (n_geoms_i - 1) * bound_stride); stride 1 comes with the build
};

inline int rt_fail(rt_context* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
        ctx->err = buf;
    }
    return code;
}
#define RT_CUDA(ctx, call)                                                                                     \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess)                                                                                \
            return rt_fail(ctx, e__ == cudaErrorMemoryAllocation ? RT_ERROR_OUT_OF_MEMORY : RT_ERROR_CUDA,        \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__);         \
    } while (0)

inline int rt_ensure(rt_context* ctx, void** p, size_t* cap, size_t need) {
    if (*cap >= need && *p) return RT_SUCCESS;
    if (*p) { cudaFree(*p); *p = nullptr; *cap = 0; }
    size_t want = need + need / 8 + 256;
    RT_CUDA(ctx, cudaMalloc(p, want));
    *cap = want;
    return RT_SUCCESS;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }


// internal entry points of rtcore_api.cu used by rt_group.cu
int rt_context_mirror_shader_state(rt_context* dst, const rt_context* src);
