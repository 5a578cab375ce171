// rtcore_api.cu — implementation of the C ABI in include/rtcore.h (host runtime of librtcore).
//
// Mirrors the reference's host-side sequence: createBLAS (main.cpp:674-831) -> rt_build_blas,
// createTLAS (main.cpp:833-949) -> rt_build_tlas, createUniformBuffer/createShaderBindingTable
// (main.cpp:1001-1017,1264-1320) -> rt_camera / rt_set_hit_records, render (main.cpp:1322-1423)
// -> rt_trace. Builds are synchronous on return like the reference's vkQueueWaitIdle
// (main.cpp:820,942); inputs are consumed before return (the reference frees them at :823-830).
// There is no CPU path: without a CUDA device every call fails with RT_ERROR_CUDA.
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "rt_device.cuh"
#include "rt_host.h"

using namespace rt;

namespace {

#define fail rt_fail
#define ensure rt_ensure

struct Carver {
    uint8_t* base; size_t off = 0;
    explicit Carver(void* b) : base((uint8_t*)b) {}
    template <typename T> T* take(size_t count) { off = align_up(off, 256); T* p = (T*)(base + off); off += sizeof(T) * count; return p; }
};

size_t build_scratch_bytes(uint32_t n, const SortPlan& sp, bool tris, uint32_t n_geoms, uint32_t n_blas) {
    size_t b = 0;
    auto add = [&](size_t bytes) { b = align_up(b, 256) + bytes; };
    add(8ull * n); add(8ull * n); add(4ull * n); add(4ull * n);          // keys a/b, vals a/b
    add(8ull * n); add(4ull * n + 4); add(48ull * tree_job_capacity(n)); add(64ull * n);   // far_end, arrived (+ job count), border jobs, deposits
    add(sp.scratch_bytes);
    if (tris) { add(48ull * n); add(sizeof(GeomDesc) * (size_t)n_geoms); add(4ull * (n_geoms + 1)); add(24ull * n_blas); }
    else { add(96ull * n); add(24ull * n); add(64ull * n); add(64); add(64); }
    return b + 4096;
}

void carve_common(Carver& c, uint32_t n, const SortPlan& sp, BuildScratch& s) {
    s.keys_a = c.take<uint64_t>(n); s.keys_b = c.take<uint64_t>(n);
    s.vals_a = c.take<uint32_t>(n); s.vals_b = c.take<uint32_t>(n);
    s.far_end = c.take<uint32_t>(2 * (size_t)n); s.arrived = c.take<uint32_t>((size_t)n + 1);
    s.jobs = c.take<float4>(3 * (size_t)tree_job_capacity(n));
    s.xchg = c.take<float4>(4 * (size_t)n);
    s.sort_scratch = c.take<uint8_t>(sp.scratch_bytes);
}

// What the builds allocate; rt_blas_build_sizes / rt_tlas_build_sizes (vkGetAccelerationStructureBuildSizesKHR, main.cpp:756-762,892-898)
// report these very numbers, so they are upper bounds by construction.
size_t blas_storage_bytes(uint32_t n_tris, uint32_t n_blas) {
    return align_up(sizeof(BvhNode) * (size_t)n_tris, 256) + align_up(sizeof(TriRec) * (size_t)n_tris, 256) + sizeof(BlasRecord) * (size_t)n_blas + 256;
}
size_t blas_scratch_bytes(uint32_t n_tris, const SortPlan& sp, uint32_t n_geoms, uint32_t n_blas, size_t stage_bytes) {
    return build_scratch_bytes(n_tris ? n_tris : 1, sp, true, n_geoms, n_blas) + align_up(stage_bytes, 256) + 4096;
}
size_t tlas_storage_bytes(uint32_t n) {
    return align_up(sizeof(InstanceRec) * (size_t)(n ? n : 1), 256) + align_up(sizeof(BvhNode) * (size_t)(n ? n : 1), 256) + 256;
}
size_t tlas_scratch_bytes(uint32_t n, const SortPlan& sp) { return build_scratch_bytes(n ? n : 1, sp, false, 0, 0) + 64ull * (n ? n : 1) + 4096; }

uint32_t ceil_log2(uint32_t v) { uint32_t b = 0; while ((1ull << b) < v) ++b; return b; }

// Packed sort records `key << vb | id` whenever key and id fit one 64-bit word (8 B instead of 12 B moved per element and pass).
void choose_record_format(SortPlan& sp, uint32_t n, uint32_t key_bits, uint32_t build_flags) {
    const uint32_t vb = ceil_log2(n) ? ceil_log2(n) : 1u;
    sp.packed_val_bits = (!(build_flags & RT_BUILD_NO_PACKED_SORT) && key_bits + vb <= 64u) ? (int)vb : 0;
}

}  // namespace

extern "C" {

const char* rt_version(void) { return "rtcore-b200 0.1 (sm_100a)"; }

void rt_destroy(rt_context* ctx);

int rt_create(int device_ordinal, rt_context** out) {
    if (!out) return RT_ERROR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_ordinal < 0 || device_ordinal >= count) return RT_ERROR_CUDA;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return RT_ERROR_CUDA;
    rt_context* ctx = new rt_context();
    ctx->device = device_ordinal;
    bool ok = cudaGetDeviceProperties(&ctx->prop, device_ordinal) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ctx->own_stream = ok;
    ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : ctx->ev) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    for (auto& e : ctx->chunk_ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->host_ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&ctx->h_err_pinned, 64) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->join_ev, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_stats, 8 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_error, 64) == cudaSuccess;
    ok = ok && cudaMemset(ctx->d_error, 0, 64) == cudaSuccess;
    if (!ok) { rt_destroy(ctx); return RT_ERROR_CUDA; }          // releases whatever was created so far
    if (const char* v = getenv("RTCORE_E2E_CHUNKS")) { int k = atoi(v); if (k >= 1 && k <= 8) ctx->e2e_chunks = k; }
    if (const char* v = getenv("RTCORE_TRACE_CHUNKS")) { int k = atoi(v); if (k >= 1 && k <= 8) ctx->trace_chunks = k; }
    *out = ctx;
    return RT_SUCCESS;
}

void rt_destroy(rt_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);      // a host frame enqueued with RT_TRACE_ASYNC may still be copying
    if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
    if (ctx->own_hit_records) { cudaFree(ctx->d_hit_records); cudaFree(ctx->d_anyhit); }
    cudaFree(ctx->scratch); cudaFree(ctx->fb); cudaFree(ctx->hits1); cudaFree(ctx->hits2);
    cudaFree(ctx->d_stats); cudaFree(ctx->d_error); cudaFree(ctx->queue); cudaFree(ctx->qflags); cudaFree(ctx->d_counters);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->chunk_ev) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->host_ev) cudaEventDestroy(ctx->host_ev);
    if (ctx->h_err_pinned) cudaFreeHost(ctx->h_err_pinned);
    if (ctx->join_ev) cudaEventDestroy(ctx->join_ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rt_release_scratch(rt_context* ctx) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_cap = 0;
    cudaFree(ctx->fb); ctx->fb = nullptr; ctx->fb_cap = 0;
    cudaFree(ctx->hits1); ctx->hits1 = nullptr; ctx->hits1_cap = 0;
    cudaFree(ctx->hits2); ctx->hits2 = nullptr; ctx->hits2_cap = 0;
    cudaFree(ctx->queue); ctx->queue = nullptr; ctx->queue_cap = 0;
    cudaFree(ctx->qflags); ctx->qflags = nullptr; ctx->qflags_cap = 0;
    ctx->dbg_keys = nullptr; ctx->dbg_vals = nullptr; ctx->dbg_n = 0;     // the debug view of the last sort lived in the scratch
    return RT_SUCCESS;
}

const char* rt_last_error(const rt_context* ctx) { return ctx ? ctx->err.c_str() : "no context (CUDA device unavailable?)"; }

int rt_device_info(const rt_context* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (total_mem) *total_mem = ctx->prop.totalGlobalMem;
    return RT_SUCCESS;
}

int rt_set_stream(rt_context* ctx, void* cuda_stream) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) RT_CUDA(ctx, cudaStreamDestroy(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return RT_SUCCESS;
}

int rt_sync(rt_context* ctx) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    if (ctx->host_pending) { const int rc = rt_host_frame_wait(ctx); if (rc != RT_SUCCESS) return rc; }
    int h_err = 0;
    RT_CUDA(ctx, cudaMemcpyAsync(&h_err, ctx->d_error, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_err) { cudaMemsetAsync(ctx->d_error, 0, 4, ctx->stream); return fail(ctx, RT_ERROR_INTERNAL, "device watchdog fired (code %d: 2 = bounce queue, 3 = flag wait timed out)", h_err); }
    return RT_SUCCESS;
}

int rt_blas_build_sizes(rt_context* ctx, const uint32_t* max_triangle_counts, uint32_t n_geoms, rt_build_sizes* out) {
    if (!ctx || !out || (!max_triangle_counts && n_geoms)) return RT_ERROR_INVALID_ARG;
    uint64_t n = 0;
    for (uint32_t g = 0; g < n_geoms; ++g) n += max_triangle_counts[g];
    if (n > MAX_PRIMS) return fail(ctx, RT_ERROR_INVALID_ARG, "too many triangles (%llu > %u)", (unsigned long long)n, MAX_PRIMS);
    SortPlan sp = sort_plan((uint32_t)n, MORTON_BITS);
    out->acceleration_structure_size = blas_storage_bytes((uint32_t)n, 1);
    // staging of HOST inputs: an upper bound for 12-byte vertices with at most 3 vertices per triangle (non-indexed lists are exactly
    // that) plus 12 B of indices per triangle and the 16-byte alignment of every array; device-pointer inputs are not staged
    const size_t stage = (size_t)n * 48 + 32ull * n_geoms;
    out->build_scratch_size = blas_scratch_bytes((uint32_t)n, sp, n_geoms, 1, stage);
    return RT_SUCCESS;
}

int rt_tlas_build_sizes(rt_context* ctx, uint32_t max_instances, rt_build_sizes* out) {
    if (!ctx || !out) return RT_ERROR_INVALID_ARG;
    SortPlan sp = sort_plan(max_instances, MORTON_BITS);
    out->acceleration_structure_size = tlas_storage_bytes(max_instances);
    out->build_scratch_size = tlas_scratch_bytes(max_instances, sp);
    return RT_SUCCESS;
}

static void storage_release(BlasStorage* st) {
    if (!st) return;
    if (--st->refs <= 0) { cudaFree(st->dev); cudaFree(st->keys); cudaFree(st->vals); delete st; }
}

// update != nullptr: rebuild that (single, unshared) BLAS inside its existing device allocation, so that its handle
// and the device address TLAS instances refer to stay valid (VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR semantics:
// same geometry/primitive counts, new vertex data).
static int build_blas_batch_impl(rt_context* ctx, const rt_geometry* geoms, const uint32_t* geom_counts, uint32_t n_blas,
                                 uint32_t build_flags, rt_blas** out_array, rt_blas* update) {
    if (!ctx || (!out_array && !update) || n_blas == 0 || !geom_counts) return RT_ERROR_INVALID_ARG;
    if (n_blas > (1u << 24)) return fail(ctx, RT_ERROR_INVALID_ARG, "at most 2^24 BLASes per batch");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (out_array) for (uint32_t b = 0; b < n_blas; ++b) out_array[b] = nullptr;
    uint32_t n_geoms = 0;
    for (uint32_t b = 0; b < n_blas; ++b) n_geoms += geom_counts[b];
    if (n_geoms && !geoms) return RT_ERROR_INVALID_ARG;

    // ---- host-side layout: geometry prefix, per-BLAS segments ----
    std::vector<GeomDesc> descs(n_geoms);
    std::vector<uint32_t> prefix(n_geoms + 1, 0);
    std::vector<BlasRecord> recs(n_blas);
    std::vector<int> bounds(6 * (size_t)n_blas);
    uint64_t total = 0; size_t stage_bytes = 0;
    {
        uint32_t g = 0;
        for (uint32_t b = 0; b < n_blas; ++b) {
            memset(&recs[b], 0, sizeof(BlasRecord));
            recs[b].first = (uint32_t)total; recs[b].n_geoms = geom_counts[b]; recs[b].root = REF_EMPTY;
            for (int k = 0; k < 3; ++k) { recs[b].lo[k] = FLT_MAX; recs[b].hi[k] = -FLT_MAX; bounds[6 * b + k] = float_to_ordered(FLT_MAX); bounds[6 * b + 3 + k] = float_to_ordered(-FLT_MAX); }
            uint64_t bt = 0;
            for (uint32_t k = 0; k < geom_counts[b]; ++k, ++g) {
                const rt_geometry& G = geoms[g];
                if (G.triangle_count && (!G.vertices || G.vertex_stride_bytes < 12 || (G.vertex_stride_bytes & 3)))
                    return fail(ctx, RT_ERROR_INVALID_ARG, "geometry %u: bad vertex buffer/stride", g);
                if (!G.indices && G.triangle_count && (uint64_t)G.triangle_count * 3 > G.vertex_count)
                    return fail(ctx, RT_ERROR_INVALID_ARG, "geometry %u: non-indexed list needs 3*triangle_count vertices", g);
                prefix[g] = (uint32_t)total;
                GeomDesc& D = descs[g];
                memset(&D, 0, sizeof(D));
                D.stride_f = G.vertex_stride_bytes / 4; D.tri_first = (uint32_t)total; D.tri_count = G.triangle_count;
                D.blas = b; D.geo_index = k; D.flags = G.flags & 0xFFu; D.vert_count = G.vertex_count;
                total += G.triangle_count; bt += G.triangle_count;
                if (!(G.flags & RT_GEOMETRY_DEVICE_POINTERS)) {
                    stage_bytes = align_up(stage_bytes, 16) + (size_t)G.vertex_count * G.vertex_stride_bytes;
                    if (G.indices) stage_bytes = align_up(stage_bytes, 16) + 12ull * G.triangle_count;
                    if (G.transform3x4) memcpy(D.xform, G.transform3x4, 48);
                    D.has_xform = G.transform3x4 ? 1u : 0u;
                }
            }
            recs[b].tri_count = (uint32_t)bt;
        }
        prefix[n_geoms] = (uint32_t)total;
    }
    if (total > MAX_PRIMS) return fail(ctx, RT_ERROR_INVALID_ARG, "too many triangles in one build (%llu > %u)", (unsigned long long)total, MAX_PRIMS);
    const uint32_t N = (uint32_t)total;
    const uint32_t seg_bits = ceil_log2(n_blas);
    if (seg_bits + MORTON_BITS > 64) return RT_ERROR_INVALID_ARG;
    SortPlan sp = sort_plan(N, (int)(MORTON_BITS + seg_bits));
    choose_record_format(sp, N, MORTON_BITS + seg_bits, build_flags);

    // ---- output storage: nodes | tris | records ----
    BlasStorage* st;
    if (update) {
        st = update->st;
        if (st->n_blas != 1 || n_blas != 1) return fail(ctx, RT_ERROR_INVALID_ARG, "rt_update_blas needs a BLAS that was built on its own (not part of a batch)");
        if (st->n_tris != N || update->r().n_geoms != geom_counts[0])
            return fail(ctx, RT_ERROR_INVALID_ARG, "rt_update_blas: geometry/triangle counts differ from the original build (%u/%u vs %u/%u)",
                        geom_counts[0], N, update->r().n_geoms, st->n_tris);
        if (st->compacted) return fail(ctx, RT_ERROR_INVALID_ARG, "rt_update_blas: a compacted BLAS cannot be updated (Vulkan: compaction copies drop ALLOW_UPDATE)");
        ++st->refs;    // the guard below drops it again
    } else {
        st = new BlasStorage();
        st->n_tris = N; st->n_blas = n_blas; st->n_node_slots = N;
        const size_t nodes_b = align_up(sizeof(BvhNode) * (size_t)N, 256), tris_b = align_up(sizeof(TriRec) * (size_t)N, 256);
        st->bytes = blas_storage_bytes(N, n_blas);
        cudaError_t ce = cudaMalloc(&st->dev, st->bytes);
        if (ce != cudaSuccess) { const size_t want = st->bytes; delete st; return fail(ctx, RT_ERROR_OUT_OF_MEMORY, "cudaMalloc(%zu) for BLAS storage failed: %s", want, cudaGetErrorString(ce)); }
        st->nodes = (BvhNode*)st->dev; st->tris = (TriRec*)((uint8_t*)st->dev + nodes_b); st->records = (BlasRecord*)((uint8_t*)st->dev + nodes_b + tris_b);
        st->refs = 1;   // held by this function until handles exist
    }
    for (uint32_t b = 0; b < n_blas; ++b) { recs[b].nodes = st->nodes + recs[b].first; recs[b].tris = st->tris + recs[b].first; recs[b].node_slots = recs[b].tri_count; }
    // refit-only update: needs the sorted records of the last full build, in the record format this build would choose again
    const bool refit = update && (build_flags & RT_BUILD_MODE_REFIT) != 0;
    if (refit && (!st->keys || st->key_vb != sp.packed_val_bits || (st->key_vb == 0 && !st->vals)))
        return fail(ctx, RT_ERROR_INVALID_ARG, "RT_BUILD_MODE_REFIT needs a BLAS whose last full build had RT_BUILD_ALLOW_UPDATE");
    struct Guard { BlasStorage* s; ~Guard() { if (s) storage_release(s); } } guard{st};

    // ---- scratch ----
    const size_t need = blas_scratch_bytes(N, sp, n_geoms, n_blas, stage_bytes);
    ctx->last_scratch_need = need;
    int rc = ensure(ctx, &ctx->scratch, &ctx->scratch_cap, need);
    if (rc != RT_SUCCESS) return rc;
    Carver c(ctx->scratch);
    BlasBuildArgs a{};
    carve_common(c, N ? N : 1, sp, a.s);
    a.s.error_flag = ctx->d_error;
    a.tris_unsorted = c.take<TriRec>(N ? N : 1);
    GeomDesc* d_descs = c.take<GeomDesc>(n_geoms ? n_geoms : 1);
    uint32_t* d_prefix = c.take<uint32_t>(n_geoms + 1);
    int* d_bounds = c.take<int>(6 * (size_t)n_blas);
    uint8_t* d_stage = c.take<uint8_t>(stage_bytes + 16);

    // ---- stage host inputs (H2D, timed separately) ----
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
    {
        std::vector<uint8_t> pack;           // small arrays are packed into one copy; large ones go directly
        const size_t DIRECT = 1u << 20;
        size_t off = 0;
        struct Direct { size_t off; const void* src; size_t bytes; };
        std::vector<Direct> direct;
        pack.resize(stage_bytes + 16);
        uint32_t g = 0;
        for (uint32_t b = 0; b < n_blas; ++b)
            for (uint32_t k = 0; k < geom_counts[b]; ++k, ++g) {
                const rt_geometry& G = geoms[g];
                GeomDesc& D = descs[g];
                if (G.flags & RT_GEOMETRY_DEVICE_POINTERS) {
                    D.verts = G.vertices; D.idx = G.indices;
                    if (G.transform3x4) {
                        RT_CUDA(ctx, cudaMemcpyAsync(D.xform, G.transform3x4, 48, cudaMemcpyDeviceToHost, ctx->stream));
                        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                        D.has_xform = 1u;
                    }
                    continue;
                }
                const size_t vb = (size_t)G.vertex_count * G.vertex_stride_bytes;
                off = align_up(off, 16);
                D.verts = (const float*)(d_stage + off);
                if (vb >= DIRECT) direct.push_back({off, G.vertices, vb}); else if (vb) memcpy(pack.data() + off, G.vertices, vb);
                off += vb;
                if (G.indices) {
                    const size_t ib = 12ull * G.triangle_count;
                    off = align_up(off, 16);
                    D.idx = (const uint32_t*)(d_stage + off);
                    if (ib >= DIRECT) direct.push_back({off, G.indices, ib}); else if (ib) memcpy(pack.data() + off, G.indices, ib);
                    off += ib;
                }
            }
        if (direct.empty()) {
            if (off) RT_CUDA(ctx, cudaMemcpyAsync(d_stage, pack.data(), off, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            // copy packed small arrays piecewise around the direct ones
            size_t cur = 0;
            for (const Direct& dct : direct) {
                if (dct.off > cur) RT_CUDA(ctx, cudaMemcpyAsync(d_stage + cur, pack.data() + cur, dct.off - cur, cudaMemcpyHostToDevice, ctx->stream));
                RT_CUDA(ctx, cudaMemcpyAsync(d_stage + dct.off, dct.src, dct.bytes, cudaMemcpyHostToDevice, ctx->stream));
                cur = dct.off + dct.bytes;
            }
            if (off > cur) RT_CUDA(ctx, cudaMemcpyAsync(d_stage + cur, pack.data() + cur, off - cur, cudaMemcpyHostToDevice, ctx->stream));
        }
        if (n_geoms) RT_CUDA(ctx, cudaMemcpyAsync(d_descs, descs.data(), sizeof(GeomDesc) * n_geoms, cudaMemcpyHostToDevice, ctx->stream));
        RT_CUDA(ctx, cudaMemcpyAsync(d_prefix, prefix.data(), 4ull * (n_geoms + 1), cudaMemcpyHostToDevice, ctx->stream));
        RT_CUDA(ctx, cudaMemcpyAsync(d_bounds, bounds.data(), 24ull * n_blas, cudaMemcpyHostToDevice, ctx->stream));
        RT_CUDA(ctx, cudaMemcpyAsync(st->records, recs.data(), sizeof(BlasRecord) * (size_t)n_blas, cudaMemcpyHostToDevice, ctx->stream));
        RT_CUDA(ctx, cudaMemsetAsync(ctx->d_error, 0, 4, ctx->stream));
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // pack/descs are host temporaries
    }
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));

    // batches of BLASes that each fit one CTA's shared memory: segmented sort (one kernel) instead of 5 global passes
    {
        uint32_t max_seg = 0;
        for (uint32_t b = 0; b < n_blas; ++b) max_seg = recs[b].tri_count > max_seg ? recs[b].tri_count : max_seg;
        if (sp.packed_val_bits > 0 && max_seg <= SEG_SORT_CAPACITY && !(build_flags & RT_BUILD_NO_SEGMENTED_SORT)) {
            sp.seg_records = st->records; sp.n_segments = n_blas; sp.seg_key_bits = (int)MORTON_BITS;
            sp.seg_fused = !(build_flags & RT_BUILD_NO_FUSED_SETUP);
        }
    }
    // ---- device build ----
    a.geoms = d_descs; a.n_geoms = n_geoms; a.geom_tri_first = d_prefix; a.n_tris = N; a.n_blas = n_blas; a.seg_bits = seg_bits;
    a.tris_sorted = st->tris; a.nodes = st->nodes; a.records = st->records; a.bounds_ordered = d_bounds; a.sort = sp;
    if (refit) { a.reuse_keys = st->keys; a.reuse_vals = st->vals; }
    BuildEvents be; for (int k = 0; k < 6; ++k) be.e[k] = ctx->ev[k];
    bool in_b = false;
    int launches = 0;
    if (N > 0) {
        launches = launch_blas_build(a, ctx->stream, &be, &in_b);
        if (launches < 0) return fail(ctx, RT_ERROR_CUDA, "BLAS build launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        ctx->launches += (uint64_t)launches;
    }
    int h_err = 0;
    RT_CUDA(ctx, cudaMemcpyAsync(&h_err, ctx->d_error, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaMemcpyAsync(recs.data(), st->records, sizeof(BlasRecord) * (size_t)n_blas, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_err) return fail(ctx, RT_ERROR_INTERNAL, "radix-sort look-back watchdog fired");
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    ctx->timing.primitives = N;
    cudaEventElapsedTime(&ctx->timing.h2d_ms, ctx->ev[6], ctx->ev[7]);
    if (N > 0) {
        cudaEventElapsedTime(&ctx->timing.setup_ms, be.e[0], be.e[1]);
        cudaEventElapsedTime(&ctx->timing.morton_ms, be.e[1], be.e[2]);
        cudaEventElapsedTime(&ctx->timing.sort_ms, be.e[2], be.e[3]);
        cudaEventElapsedTime(&ctx->timing.hierarchy_ms, be.e[3], be.e[4]);
        cudaEventElapsedTime(&ctx->timing.refit_ms, be.e[4], be.e[5]);
        cudaEventElapsedTime(&ctx->timing.total_ms, be.e[0], be.e[5]);
    }
    if (!refit) { ctx->dbg_keys = in_b ? a.s.keys_b : a.s.keys_a; ctx->dbg_vals = in_b ? a.s.vals_b : a.s.vals_a; ctx->dbg_n = N; ctx->dbg_vb = sp.packed_val_bits; }
    // ALLOW_UPDATE / ALLOW_COMPACTION: the BLAS keeps the sorted records of this (full) build: they are its topology
    if (!refit) {
        // (a refit needs a BLAS that was built on its own; a compaction also works on a whole batch)
        const uint32_t keep = (build_flags & RT_BUILD_ALLOW_COMPACTION) | (n_blas == 1 ? (build_flags & RT_BUILD_ALLOW_UPDATE) : 0u);
        if (!keep || N == 0) { cudaFree(st->keys); cudaFree(st->vals); st->keys = nullptr; st->vals = nullptr; }
        else {
            if (!st->keys) { cudaError_t ce = cudaMalloc((void**)&st->keys, 8ull * N); if (ce != cudaSuccess) return fail(ctx, RT_ERROR_OUT_OF_MEMORY, "cudaMalloc for the retained sort records failed"); }
            RT_CUDA(ctx, cudaMemcpyAsync(st->keys, ctx->dbg_keys, 8ull * N, cudaMemcpyDeviceToDevice, ctx->stream));
            if (sp.packed_val_bits == 0) {
                if (!st->vals) { cudaError_t ce = cudaMalloc((void**)&st->vals, 4ull * N); if (ce != cudaSuccess) return fail(ctx, RT_ERROR_OUT_OF_MEMORY, "cudaMalloc for the retained sort records failed"); }
                RT_CUDA(ctx, cudaMemcpyAsync(st->vals, ctx->dbg_vals, 4ull * N, cudaMemcpyDeviceToDevice, ctx->stream));
            } else { cudaFree(st->vals); st->vals = nullptr; }
            RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            st->key_vb = sp.packed_val_bits;
        }
        st->build_flags = build_flags;
    }
    st->host_recs = recs;

    if (update) return RT_SUCCESS;
    for (uint32_t b = 0; b < n_blas; ++b) {
        rt_blas* h = new rt_blas();
        h->st = st; h->index = b;
        ++st->refs;
        out_array[b] = h;
    }
    return RT_SUCCESS;   // guard drops the construction reference
}

int rt_build_blas_batch(rt_context* ctx, const rt_geometry* geoms, const uint32_t* geom_counts, uint32_t n_blas,
                        uint32_t build_flags, rt_blas** out_array) {
    if (!out_array) return RT_ERROR_INVALID_ARG;
    return build_blas_batch_impl(ctx, geoms, geom_counts, n_blas, build_flags, out_array, nullptr);
}

int rt_update_blas(rt_context* ctx, rt_blas* blas, const rt_geometry* geoms, uint32_t n_geoms, uint32_t build_flags) {
    if (!blas) return RT_ERROR_INVALID_ARG;
    uint32_t counts[1] = {n_geoms};
    return build_blas_batch_impl(ctx, geoms, counts, 1, build_flags, nullptr, blas);
}

int rt_build_blas(rt_context* ctx, const rt_geometry* geoms, uint32_t n_geoms, uint32_t build_flags, rt_blas** out) {
    if (!out) return RT_ERROR_INVALID_ARG;
    uint32_t counts[1] = {n_geoms};
    return rt_build_blas_batch(ctx, geoms, counts, 1, build_flags, out);
}

int rt_compact_blas(rt_context* ctx, rt_blas* blas, uint64_t* bytes_before, uint64_t* bytes_after) {
    if (!ctx || !blas || !blas->st) return RT_ERROR_INVALID_ARG;
    BlasStorage* st = blas->st;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bytes_before) *bytes_before = st->bytes;
    if (bytes_after) *bytes_after = st->bytes;
    if (st->compacted) return RT_SUCCESS;
    const uint32_t N = st->n_tris, n_blas = st->n_blas;
    if (N == 0) { st->compacted = true; return RT_SUCCESS; }
    if (!st->keys || !(st->build_flags & RT_BUILD_ALLOW_COMPACTION) || (st->key_vb == 0 && !st->vals))
        return fail(ctx, RT_ERROR_INVALID_ARG, "rt_compact_blas needs a BLAS built with RT_BUILD_ALLOW_COMPACTION");
    // pass 1: compact index of every Karras slot (scratch: cidx[N + 1] + chunk sums)
    const size_t need = align_up(4ull * ((size_t)N + 1), 256) + compact_scratch_bytes(N);
    int rc = ensure(ctx, &ctx->scratch, &ctx->scratch_cap, need);
    if (rc != RT_SUCCESS) return rc;
    ctx->dbg_n = 0;                                                  // the scratch no longer holds the last build's sorted keys
    uint32_t* cidx = (uint32_t*)ctx->scratch;
    void* aux = (uint8_t*)ctx->scratch + align_up(4ull * ((size_t)N + 1), 256);
    int l = launch_compact_index(st->keys, st->key_vb, N, st->records, cidx, aux, ctx->stream);
    if (l < 0) return fail(ctx, RT_ERROR_CUDA, "compaction launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches += (uint64_t)l;
    uint32_t live = 0;
    RT_CUDA(ctx, cudaMemcpyAsync(&live, cidx + N, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // new right-sized storage: nodes[live] | tris[N] | records[n_blas]
    const size_t nodes_b = align_up(sizeof(BvhNode) * (size_t)(live ? live : 1), 256), tris_b = align_up(sizeof(TriRec) * (size_t)N, 256);
    const size_t bytes = nodes_b + tris_b + sizeof(BlasRecord) * (size_t)n_blas + 256;
    void* dev = nullptr;
    cudaError_t ce = cudaMalloc(&dev, bytes);
    if (ce != cudaSuccess) return fail(ctx, RT_ERROR_OUT_OF_MEMORY, "cudaMalloc(%zu) for the compacted BLAS failed: %s", bytes, cudaGetErrorString(ce));
    BvhNode* nnodes = (BvhNode*)dev; TriRec* ntris = (TriRec*)((uint8_t*)dev + nodes_b); BlasRecord* nrecs = (BlasRecord*)((uint8_t*)dev + nodes_b + tris_b);
    l = launch_compact_nodes(st->nodes, nnodes, cidx, st->keys, st->key_vb, N, st->records, ctx->stream);
    if (l < 0) { cudaFree(dev); return fail(ctx, RT_ERROR_CUDA, "compaction launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
    ctx->launches += (uint64_t)l;
    cudaError_t e1 = cudaMemcpyAsync(ntris, st->tris, sizeof(TriRec) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream);
    const int l2 = launch_compact_records(st->records, nrecs, n_blas, cidx, nnodes, ntris, ctx->stream);
    if (l2 < 0) { cudaFree(dev); return fail(ctx, RT_ERROR_CUDA, "compaction launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
    ctx->launches += (uint64_t)l2;
    cudaError_t e2 = cudaMemcpyAsync(st->host_recs.data(), nrecs, sizeof(BlasRecord) * (size_t)n_blas, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cudaFree(dev); return fail(ctx, RT_ERROR_CUDA, "compaction copy failed"); }
    cudaFree(st->dev); cudaFree(st->keys); cudaFree(st->vals);
    st->keys = nullptr; st->vals = nullptr;
    st->dev = dev; st->bytes = bytes; st->nodes = nnodes; st->tris = ntris; st->records = nrecs; st->n_node_slots = live; st->compacted = true;
    if (bytes_after) *bytes_after = bytes;
    return RT_SUCCESS;
}

void rt_free_blas(rt_context* ctx, rt_blas* blas) {
    if (!blas) return;
    if (ctx) cudaSetDevice(ctx->device);
    storage_release(blas->st);
    delete blas;
}

int rt_last_build_timing(const rt_context* ctx, rt_build_timing* out) {
    if (!ctx || !out) return RT_ERROR_INVALID_ARG;
    *out = ctx->timing;
    return RT_SUCCESS;
}
float rt_last_build_ms(const rt_context* ctx) { return ctx ? ctx->timing.total_ms : 0.0f; }
uint64_t rt_last_build_scratch_bytes(const rt_context* ctx) { return ctx ? (uint64_t)ctx->last_scratch_need : 0; }

uint64_t rt_blas_device_reference(const rt_context* ctx, const rt_blas* blas) {
    if (!ctx || !blas || !blas->st) return 0;
    return (uint64_t)(uintptr_t)(blas->st->records + blas->index);
}
uint64_t rt_tlas_storage_bytes(const rt_context* ctx, const rt_tlas* tlas) { return (ctx && tlas) ? (uint64_t)tlas->bytes : 0; }

int rt_blas_get_info(rt_context* ctx, const rt_blas* blas, rt_blas_info* out) {
    if (!ctx || !blas || !out) return RT_ERROR_INVALID_ARG;
    out->triangle_count = blas->r().tri_count;
    out->node_count = blas->r().node_slots;    // Karras slots [0, n) after a build (slot indices are BLAS-relative); the live nodes after compaction
    out->root_ref = blas->r().root;
    out->max_depth = blas->r().height;
    for (int k = 0; k < 3; ++k) { out->bounds_lo[k] = blas->r().lo[k]; out->bounds_hi[k] = blas->r().hi[k]; }
    out->storage_bytes = blas->st->n_blas == 1 ? (uint64_t)blas->st->bytes : 0;
    out->device_storage = blas->st->n_blas == 1 ? blas->st->dev : nullptr;
    return RT_SUCCESS;
}

int rt_blas_export(rt_context* ctx, const rt_blas* blas, void* nodes_out, void* tris_out) {
    if (!ctx || !blas) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t n = blas->r().tri_count, nn = blas->r().node_slots;
    if (nodes_out && nn) RT_CUDA(ctx, cudaMemcpy(nodes_out, blas->r().nodes, sizeof(BvhNode) * nn, cudaMemcpyDeviceToHost));
    if (tris_out && n) RT_CUDA(ctx, cudaMemcpy(tris_out, blas->r().tris, sizeof(TriRec) * n, cudaMemcpyDeviceToHost));
    return RT_SUCCESS;
}

int rt_debug_last_sorted_keys(rt_context* ctx, uint64_t* keys_out, uint32_t* prim_out, uint32_t capacity, uint32_t* n_out) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    if (n_out) *n_out = ctx->dbg_n;
    const uint32_t n = ctx->dbg_n < capacity ? ctx->dbg_n : capacity;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->dbg_vb) {                                   // packed records: split `key << vb | id` on the host
        std::vector<uint64_t> rec(n);
        if (n) RT_CUDA(ctx, cudaMemcpy(rec.data(), ctx->dbg_keys, 8ull * n, cudaMemcpyDeviceToHost));
        const uint64_t mask = (1ull << ctx->dbg_vb) - 1ull;
        for (uint32_t i = 0; i < n; ++i) {
            if (keys_out) keys_out[i] = rec[i] >> ctx->dbg_vb;
            if (prim_out) prim_out[i] = (uint32_t)(rec[i] & mask);
        }
        return RT_SUCCESS;
    }
    if (n && keys_out) RT_CUDA(ctx, cudaMemcpy(keys_out, ctx->dbg_keys, 8ull * n, cudaMemcpyDeviceToHost));
    if (n && prim_out) RT_CUDA(ctx, cudaMemcpy(prim_out, ctx->dbg_vals, 4ull * n, cudaMemcpyDeviceToHost));
    return RT_SUCCESS;
}

int rt_blas_import(rt_context* ctx, const rt_blas_info* info, const void* device_blob, rt_blas** out) {
    if (!ctx || !info || !device_blob || !out) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t N = info->triangle_count, NN = info->node_count;
    BlasStorage* st = new BlasStorage();
    st->n_tris = N; st->n_blas = 1; st->n_node_slots = NN; st->compacted = NN != N;
    const size_t nodes_b = align_up(sizeof(BvhNode) * (size_t)NN, 256), tris_b = align_up(sizeof(TriRec) * (size_t)N, 256);
    st->bytes = nodes_b + tris_b + sizeof(BlasRecord) + 256;
    if (info->storage_bytes != st->bytes) { delete st; return fail(ctx, RT_ERROR_INVALID_ARG, "blob size mismatch"); }
    cudaError_t ce = cudaMalloc(&st->dev, st->bytes);
    if (ce != cudaSuccess) { delete st; return fail(ctx, RT_ERROR_OUT_OF_MEMORY, "cudaMalloc failed"); }
    st->nodes = (BvhNode*)st->dev; st->tris = (TriRec*)((uint8_t*)st->dev + nodes_b); st->records = (BlasRecord*)((uint8_t*)st->dev + nodes_b + tris_b);
    st->refs = 1;
    rt_blas* h = new rt_blas();
    h->st = st; h->index = 0;
    st->host_recs.assign(1, BlasRecord{});
    BlasRecord& R = h->r();
    memset(&R, 0, sizeof(R));
    R.nodes = st->nodes; R.tris = st->tris; R.root = info->root_ref; R.height = info->max_depth; R.tri_count = N; R.n_geoms = 0; R.first = 0; R.node_slots = NN;
    for (int k = 0; k < 3; ++k) { R.lo[k] = info->bounds_lo[k]; R.hi[k] = info->bounds_hi[k]; }
    cudaError_t e1 = cudaMemcpyAsync(st->dev, device_blob, nodes_b + tris_b, cudaMemcpyDeviceToDevice, ctx->stream);
    // n_geoms travels in the source record
    BlasRecord src{};
    cudaError_t e2 = cudaMemcpyAsync(&src, (const uint8_t*)device_blob + nodes_b + tris_b, sizeof(BlasRecord), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e3 = cudaStreamSynchronize(ctx->stream);
    R.n_geoms = src.n_geoms;
    cudaError_t e4 = cudaMemcpy(st->records, &R, sizeof(BlasRecord), cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) { rt_free_blas(ctx, h); return fail(ctx, RT_ERROR_CUDA, "blob import copy failed"); }
    *out = h;
    return RT_SUCCESS;
}

// ---- TLAS -------------------------------------------------------------------------------------------
static int tlas_build_into(rt_context* ctx, rt_tlas* T, const rt_instance* instances, uint32_t n, uint32_t build_flags) {
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    SortPlan sp = sort_plan(n, MORTON_BITS);
    choose_record_format(sp, n, MORTON_BITS, build_flags);
    if (sp.packed_val_bits > 0 && n <= SEG_SORT_CAPACITY && !(build_flags & RT_BUILD_NO_SEGMENTED_SORT)) {
        sp.seg_single = true; sp.seg_key_bits = (int)MORTON_BITS;     // a TLAS of up to 11,264 instances: one shared-memory sort kernel instead of six launches
    }
    const size_t inst_b = align_up(sizeof(InstanceRec) * (size_t)(n ? n : 1), 256), nodes_b = align_up(sizeof(BvhNode) * (size_t)(n ? n : 1), 256);
    const size_t bytes = tlas_storage_bytes(n);
    if (T->bytes < bytes) {
        if (T->dev) cudaFree(T->dev);
        T->dev = nullptr; T->bytes = 0;
        RT_CUDA(ctx, cudaMalloc(&T->dev, bytes));
        T->bytes = bytes;
    }
    T->inst = (InstanceRec*)T->dev; T->nodes = (BvhNode*)((uint8_t*)T->dev + inst_b);
    T->n = n; T->root = REF_EMPTY; T->height = 0;
    for (int k = 0; k < 3; ++k) { T->lo[k] = FLT_MAX; T->hi[k] = -FLT_MAX; }
    T->max_sbt_plus_geo = T->max_sbt = T->max_geo = T->max_blas_height = 0;
    T->bound_stride = 1; T->bound = 0;
    if (n == 0) return RT_SUCCESS;

    const size_t need = tlas_scratch_bytes(n, sp);
    ctx->last_scratch_need = need;
    int rc = ensure(ctx, &ctx->scratch, &ctx->scratch_cap, need);
    if (rc != RT_SUCCESS) return rc;
    Carver c(ctx->scratch);
    TlasBuildArgs a{};
    carve_common(c, n, sp, a.s);
    a.s.error_flag = ctx->d_error;
    a.inst_unsorted = c.take<InstanceRec>(n);
    a.boxes_unsorted = c.take<float>(6 * (size_t)n);
    rt_instance* d_inst = c.take<rt_instance>(n);
    a.bounds_ordered = c.take<int>(8);
    a.root_out = c.take<int32_t>(8);
    a.bounds_out = c.take<float>(8);

    if (build_flags & RT_BUILD_INSTANCES_ON_DEVICE) {
        a.instances = instances;
    } else {
        std::vector<rt_instance> tmp(instances, instances + n);
        for (uint32_t i = 0; i < n; ++i) {
            const rt_blas* b = instances[i].blas;
            uint64_t addr = b ? (uint64_t)(uintptr_t)(b->st->records + b->index) : 0ull;
            memcpy(&tmp[i].blas, &addr, 8);
        }
        RT_CUDA(ctx, cudaMemcpyAsync(d_inst, tmp.data(), sizeof(rt_instance) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        a.instances = d_inst;
    }
    int hb[8]; for (int k = 0; k < 3; ++k) { hb[k] = float_to_ordered(FLT_MAX); hb[3 + k] = float_to_ordered(-FLT_MAX); }
    int32_t hmeta[8] = {REF_EMPTY, 0, 0, 0, 0, 0, 0, 0};
    float hbo[8] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, 0, 0};
    RT_CUDA(ctx, cudaMemcpyAsync(a.bounds_ordered, hb, 24, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaMemcpyAsync(a.root_out, hmeta, 32, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaMemcpyAsync(a.bounds_out, hbo, 32, cudaMemcpyHostToDevice, ctx->stream));
    RT_CUDA(ctx, cudaMemsetAsync(ctx->d_error, 0, 4, ctx->stream));
    a.n = n; a.inst_sorted = T->inst; a.nodes = T->nodes; a.sort = sp;
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    int launches = launch_tlas_build(a, ctx->stream);
    if (launches < 0) return fail(ctx, RT_ERROR_CUDA, "TLAS build launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->launches += (uint64_t)launches;
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
    int h_err = 0;
    RT_CUDA(ctx, cudaMemcpyAsync(&h_err, ctx->d_error, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaMemcpyAsync(hmeta, a.root_out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaMemcpyAsync(hbo, a.bounds_out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_err) return fail(ctx, RT_ERROR_INTERNAL, "radix-sort look-back watchdog fired");
    T->root = hmeta[0]; T->height = (uint32_t)hmeta[1];
    T->max_sbt_plus_geo = hmeta[2]; T->max_sbt = hmeta[3]; T->max_geo = hmeta[4]; T->max_blas_height = hmeta[5];
    T->bound_stride = 1; T->bound = (uint64_t)hmeta[2];
    for (int k = 0; k < 3; ++k) { T->lo[k] = hbo[k]; T->hi[k] = hbo[3 + k]; }
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    ctx->timing.primitives = n;
    cudaEventElapsedTime(&ctx->timing.total_ms, ctx->ev[0], ctx->ev[5]);
    ctx->dbg_n = 0;
    return RT_SUCCESS;
}

int rt_build_tlas(rt_context* ctx, const rt_instance* instances, uint32_t n_instances, uint32_t build_flags, rt_tlas** out) {
    if (!ctx || !out || (n_instances && !instances)) return RT_ERROR_INVALID_ARG;
    *out = nullptr;
    if (n_instances > MAX_PRIMS) return RT_ERROR_INVALID_ARG;
    rt_tlas* T = new rt_tlas();
    int rc = tlas_build_into(ctx, T, instances, n_instances, build_flags);
    if (rc != RT_SUCCESS) { if (T->dev) cudaFree(T->dev); delete T; return rc; }
    *out = T;
    return RT_SUCCESS;
}

int rt_update_tlas(rt_context* ctx, rt_tlas* tlas, const rt_instance* instances, uint32_t n_instances, uint32_t build_flags) {
    if (!ctx || !tlas || (n_instances && !instances)) return RT_ERROR_INVALID_ARG;
    return tlas_build_into(ctx, tlas, instances, n_instances, build_flags);
}

void rt_free_tlas(rt_context* ctx, rt_tlas* tlas) {
    if (!tlas) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (tlas->dev) cudaFree(tlas->dev);
    delete tlas;
}

int rt_tlas_get_info(rt_context* ctx, const rt_tlas* tlas, rt_tlas_info* out) {
    if (!ctx || !tlas || !out) return RT_ERROR_INVALID_ARG;
    out->instance_count = tlas->n; out->node_count = tlas->n; out->root_ref = tlas->root; out->max_depth = tlas->height;
    for (int k = 0; k < 3; ++k) { out->bounds_lo[k] = tlas->lo[k]; out->bounds_hi[k] = tlas->hi[k]; }
    return RT_SUCCESS;
}

// ---- shader data ---------------------------------------------------------------------------------------
int rt_set_hit_records(rt_context* ctx, const float* rgb, uint32_t count) {
    if (!ctx || (count && !rgb)) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_hit_records) { cudaFree(ctx->d_hit_records); ctx->d_hit_records = nullptr; }
    ctx->n_records = 0;
    if (count) {
        RT_CUDA(ctx, cudaMalloc(&ctx->d_hit_records, 12ull * count));
        RT_CUDA(ctx, cudaMemcpy(ctx->d_hit_records, rgb, 12ull * count, cudaMemcpyHostToDevice));
        ctx->n_records = count;
    }
    return RT_SUCCESS;
}

int rt_set_miss_color(rt_context* ctx, const float rgb[3]) {
    if (!ctx || !rgb) return RT_ERROR_INVALID_ARG;
    ctx->miss[0] = rgb[0]; ctx->miss[1] = rgb[1]; ctx->miss[2] = rgb[2];
    return RT_SUCCESS;
}

int rt_set_miss_records(rt_context* ctx, const float* rgb, uint32_t count) {
    if (!ctx || !rgb || count == 0) return RT_ERROR_INVALID_ARG;
    ctx->miss.assign(rgb, rgb + 3 * (size_t)count);
    return RT_SUCCESS;
}

int rt_set_anyhit_records(rt_context* ctx, const rt_anyhit_record* records, uint32_t count) {
    if (!ctx || (count && !records)) return RT_ERROR_INVALID_ARG;
    if (!ctx->own_hit_records) return fail(ctx, RT_ERROR_INVALID_ARG, "shader data of a group's internal context follows the user context");
    // one device array: count x {kind, log2_res, flags, first mask word} (16 B each), then the masks back to back
    std::vector<uint32_t> blob(4 * (size_t)count);
    for (uint32_t i = 0; i < count; ++i) {
        const rt_anyhit_record& r = records[i];
        if (r.kind > RT_ANYHIT_ALPHA_MASK || r.log2_res > 10u || (r.kind == RT_ANYHIT_ALPHA_MASK && !r.mask))
            return fail(ctx, RT_ERROR_INVALID_ARG, "any-hit record %u: kind <= ALPHA_MASK, log2_res <= 10, mask set", i);
        blob[4 * i] = r.kind; blob[4 * i + 1] = r.log2_res; blob[4 * i + 2] = r.flags; blob[4 * i + 3] = 0;
        if (r.kind == RT_ANYHIT_ALPHA_MASK) {
            const size_t words = (((size_t)1 << (2 * r.log2_res)) + 31) / 32;
            blob[4 * i + 3] = (uint32_t)blob.size();
            blob.insert(blob.end(), r.mask, r.mask + words);
        }
    }
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->d_anyhit) { cudaFree(ctx->d_anyhit); ctx->d_anyhit = nullptr; }
    ctx->n_anyhit = 0;
    if (count) {
        RT_CUDA(ctx, cudaMalloc((void**)&ctx->d_anyhit, blob.size() * 4));
        RT_CUDA(ctx, cudaMemcpy(ctx->d_anyhit, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
        ctx->n_anyhit = count;
    }
    return RT_SUCCESS;
}

int rt_set_ray_params(rt_context* ctx, const rt_ray_params* params) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    if (params) ctx->rp = *params; else ctx->rp = {0.0f, 100.0f, 0xffu, 0u, 1u, 1u, RT_RAY_FLAG_OPAQUE, 0u};
    return RT_SUCCESS;
}

}  // extern "C"
// the render group's second context (odd frames of a pipelined sequence) follows the user's context
int rt_context_mirror_shader_state(rt_context* dst, const rt_context* src) {
    if (!dst || !src || dst->own_hit_records) return RT_ERROR_INVALID_ARG;
    dst->d_hit_records = src->d_hit_records; dst->n_records = src->n_records;
    dst->d_anyhit = src->d_anyhit; dst->n_anyhit = src->n_anyhit;
    dst->miss = src->miss; dst->rp = src->rp;
    return RT_SUCCESS;
}
extern "C" {

// ---- dispatch ------------------------------------------------------------------------------------------
uint64_t rt_rows_packed_pixels(uint32_t width, uint32_t height, uint32_t block_rows, uint32_t part_count) {
    if (!block_rows || !part_count) return 0;
    const uint64_t bands = (height + block_rows - 1) / block_rows;
    return ((bands + part_count - 1) / part_count) * block_rows * (uint64_t)width;
}

int rt_host_frame_wait(rt_context* ctx) {
    if (!ctx) return RT_ERROR_INVALID_ARG;
    if (!ctx->host_pending) return RT_SUCCESS;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaEventSynchronize(ctx->host_ev));
    ctx->host_pending = false;
    if (*ctx->h_err_pinned) {
        *ctx->h_err_pinned = 0;
        cudaMemsetAsync(ctx->d_error, 0, 4, ctx->stream);
        return fail(ctx, RT_ERROR_INTERNAL, "trace watchdog fired (bounce-queue entry never published)");
    }
    return RT_SUCCESS;
}

// range_rows == 0: all packed rows of this part; else only the packed rows [range_first, range_first + range_rows) (device output)
static int trace_rows_impl(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam, uint32_t width, uint32_t height, uint32_t bounces,
                           uint32_t flags, uint32_t block_rows, uint32_t part_index, uint32_t part_count, uint32_t range_first, uint32_t range_rows,
                           uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out) {
    if (!ctx || !tlas || !cam || !width || !height || !rgba_out) return RT_ERROR_INVALID_ARG;
    if (!block_rows || (block_rows & 3u) || !part_count || part_index >= part_count) return fail(ctx, RT_ERROR_INVALID_ARG, "block_rows must be a positive multiple of 4 and part_index < part_count");
    if ((uint64_t)width * height > 0xFFFFFFFFull) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bounces > 1) bounces = 1;
    // static SBT range check (Vulkan leaves out-of-range records undefined; we refuse)
    // (exact: max over the instances of sbt_offset + (geometryCount - 1) * stride, + offset; no hit record is read at all when the
    // closest-hit shader is skipped)
    if (tlas->n && !(ctx->rp.ray_flags & RT_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER)) {
        rt_tlas* T = const_cast<rt_tlas*>(tlas);
        if (T->bound_stride != ctx->rp.sbt_record_stride) {
            unsigned long long h = 0ull;
            RT_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, 8, ctx->stream));
            if (launch_sbt_bound(T->inst, T->n, ctx->rp.sbt_record_stride, ctx->d_stats, ctx->stream) < 0) return fail(ctx, RT_ERROR_CUDA, "SBT bound launch failed");
            ctx->launches += 1;
            RT_CUDA(ctx, cudaMemcpyAsync(&h, ctx->d_stats, 8, cudaMemcpyDeviceToHost, ctx->stream));
            RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            T->bound_stride = ctx->rp.sbt_record_stride; T->bound = h;
        }
        const uint64_t bound = T->bound + ctx->rp.sbt_record_offset;
        if (bound >= ctx->n_records) return fail(ctx, RT_ERROR_SBT_RANGE, "hit record %llu addressed but only %u set", (unsigned long long)bound, ctx->n_records);
    }
    if ((size_t)ctx->rp.miss_index * 3 + 3 > ctx->miss.size())
        return fail(ctx, RT_ERROR_SBT_RANGE, "miss record %u addressed but only %u set", ctx->rp.miss_index, (unsigned)(ctx->miss.size() / 3));
    const int stack_needed = (int)tlas->height + tlas->max_blas_height + 4;
    if (stack_needed > 160) return fail(ctx, RT_ERROR_STACK_DEPTH, "BVH depth %d exceeds the traversal stack", stack_needed);

    // one part = the plain width x height image; several parts = equal-sized packed band buffers
    const uint64_t pixels = part_count == 1 ? (uint64_t)width * height : rt_rows_packed_pixels(width, height, block_rows, part_count);
    const bool dev_out = (flags & RT_TRACE_OUT_DEVICE) != 0;
    const bool full_frame = (flags & RT_TRACE_OUT_FULL_FRAME) != 0;
    // host output + FULL_FRAME: the kernels write this part's packed bands into the context's staging buffer and every finished row
    // chunk is scattered band by band (one 2-D copy) to its final rows of the caller's whole-frame HOST buffer — with one process
    // per GPU and a frame in shared pinned host memory, every GPU moves its share over its own PCIe link
    const bool host_scatter = full_frame && !dev_out && part_count > 1;
    if (host_scatter && (primary_hits_out || secondary_hits_out)) return fail(ctx, RT_ERROR_INVALID_ARG, "hit buffers are not scattered: use packed output for them");
    TraceParams P{};
    P.full_frame = (full_frame && dev_out) ? 1u : 0u;
    P.bgra = (flags & RT_TRACE_OUT_BGRA) ? 1u : 0u;
    P.tlas_nodes = tlas->nodes; P.instances = tlas->inst; P.tlas_root = tlas->root;
    P.tlas_smem_nodes = tlas->n > 1 ? tlas->n - 1 : 0;     // Karras numbering: the n - 1 internal nodes of n leaves are slots 0..n-2
    for (int k = 0; k < 3; ++k) { P.tlas_absmax[k] = tlas->n && tlas->lo[k] <= tlas->hi[k] ? fmaxf(fabsf(tlas->lo[k]), fabsf(tlas->hi[k])) : 0.0f; P.cam_pos[k] = cam->pos[k]; P.miss[k] = ctx->miss[3 * (size_t)ctx->rp.miss_index + k]; }
    // raygen constants of main.cpp:1038-1039; tan is evaluated once on the host in fp32
    P.aspect_y = tanf((cam->yfov_deg * 0.017453292519943295f) * 0.5f);
    P.aspect_x = P.aspect_y * (float)width / (float)height;
    P.width = width; P.height = height; P.block_rows = block_rows; P.part_index = part_index; P.part_count = part_count;
    P.local_rows = (uint32_t)(pixels / width);
    P.tmin = ctx->rp.tmin; P.tmax = ctx->rp.tmax; P.cull_mask = ctx->rp.cull_mask; P.sbt_offset = ctx->rp.sbt_record_offset;
    P.sbt_stride = ctx->rp.sbt_record_stride; P.bounce_seed = ctx->rp.bounce_seed; P.bounces = bounces; P.ray_flags = ctx->rp.ray_flags;
    P.hit_records = ctx->d_hit_records; P.n_records = ctx->n_records;
    P.anyhit = ctx->d_anyhit; P.n_anyhit = ctx->n_anyhit;
    int rc;
    if (dev_out) { P.rgba = rgba_out; P.primary_hits = primary_hits_out; P.secondary_hits = secondary_hits_out; }
    else {
        if ((rc = ensure(ctx, &ctx->fb, &ctx->fb_cap, pixels * 4)) != RT_SUCCESS) return rc;
        P.rgba = (uint8_t*)ctx->fb;
        if (primary_hits_out) { if ((rc = ensure(ctx, &ctx->hits1, &ctx->hits1_cap, pixels * sizeof(rt_hit))) != RT_SUCCESS) return rc; P.primary_hits = (rt_hit*)ctx->hits1; }
        if (secondary_hits_out) { if ((rc = ensure(ctx, &ctx->hits2, &ctx->hits2_cap, pixels * sizeof(rt_hit))) != RT_SUCCESS) return rc; P.secondary_hits = (rt_hit*)ctx->hits2; }
    }
    const bool stats = (flags & RT_TRACE_STATS) != 0;
    // RT_TRACE_ASYNC with a HOST framebuffer: trace + copies are enqueued, the call returns, rt_host_frame_wait() completes the frame.
    // The staging buffer (and the watchdog word) still belong to the previous such frame until its copies are done.
    const bool async_host = !dev_out && (flags & RT_TRACE_ASYNC) != 0 && !stats && !primary_hits_out && !secondary_hits_out;
    if (!dev_out && ctx->host_pending) {
        if (async_host) RT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->host_ev, 0));
        else if ((rc = rt_host_frame_wait(ctx)) != RT_SUCCESS) return rc;
    }
    if (stats) { RT_CUDA(ctx, cudaMemsetAsync(ctx->d_stats, 0, 64, ctx->stream)); P.stats = ctx->d_stats; }
    // Row chunks (multiples of 8 rows). Host output: every finished chunk is copied device->host on the copy stream while the next
    // one is traced. Several chunks alternate over TWO compute streams: the persistent kernels of chunk c + 1 are queued behind
    // those of chunk c on the block scheduler, so their CTAs start as the CTAs of c retire and fill its kernel tails.
    const uint32_t total_rows = P.local_rows;
    uint32_t chunks = 1;
    if (pixels >= (1u << 20)) chunks = (uint32_t)(dev_out ? ctx->trace_chunks : ctx->e2e_chunks);
    if (async_host) chunks = 1;        // the copy of this frame overlaps the NEXT frame's trace: no reason to pay a kernel tail per chunk
    // chunk c covers rows [row_begin[c], row_begin[c + 1]), multiples of 8. Host output: the chunks SHRINK towards the end of the frame
    // (weights 4 : 3 : 2 : 1 for four chunks), because only the copy of the last chunk is not hidden behind a trace
    uint32_t row_begin[10] = {0};
    {
        if (chunks > 8) chunks = 8;
        const uint32_t wsum = dev_out ? chunks : chunks * (chunks + 1) / 2;
        uint32_t acc = 0, n = 0;
        for (uint32_t c = 0; c < chunks; ++c) {
            acc += dev_out ? 1u : chunks - c;
            uint32_t end = (uint32_t)(((uint64_t)total_rows * acc / wsum + 7u) & ~7ull);
            if (end > total_rows || c + 1 == chunks) end = total_rows;
            if (end > row_begin[n]) row_begin[++n] = end;
        }
        chunks = n ? n : 1;
        row_begin[chunks] = total_rows;
    }
    if (host_scatter && chunks > 1) {                   // chunk boundaries on whole bands (and multiples of 8 rows)
        const uint32_t unit = (block_rows % 8u == 0u) ? block_rows : 2u * block_rows;
        uint32_t n = 0;
        for (uint32_t c = 1; c < chunks; ++c) {
            const uint32_t e = row_begin[c] / unit * unit;
            if (e > row_begin[n] && e < total_rows) row_begin[++n] = e;
        }
        chunks = n + 1;
        row_begin[chunks] = total_rows;
    }
    if (range_rows) {                                  // an explicit row range: exactly that, one launch
        // whole bands only: packed rows [a, b) of all parts together are the image rows [a * part_count, b * part_count) only then
        const uint32_t unit = part_count == 1 ? 8u : block_rows;
        if (!dev_out || (range_first % unit) || range_first >= total_rows) return fail(ctx, RT_ERROR_INVALID_ARG, "row range: device output, first row a multiple of block_rows inside the part");
        chunks = 1;
        row_begin[0] = range_first;
        row_begin[1] = range_first + range_rows < total_rows ? range_first + range_rows : total_rows;
        if ((row_begin[1] % unit) && row_begin[1] != total_rows) return fail(ctx, RT_ERROR_INVALID_ARG, "row range: row count must be a multiple of block_rows unless it ends the part");
    }
    const bool two_streams = chunks > 1;
    // per-chunk scratch: ray slots (tile-major over whole 8x4 tiles), index list, tile masks + block sums, publication flags
    size_t ray_off[9] = {0}, idx_off[9] = {0}, mask_off[9] = {0}, slot_off[9] = {0}, ctr_off[9] = {0};
    {
        size_t words = 0;                                 // per chunk: 16 counters + 2 x n_regions region fetch counters
        for (uint32_t c = 0; c < chunks; ++c) {
            ctr_off[c] = words;
            words += 16 + 2 * (size_t)trace_regions_x(width) * trace_regions_y(row_begin[c + 1] - row_begin[c]);
            words = (words + 15) & ~(size_t)15;
        }
        ctr_off[chunks] = words;
        if ((rc = ensure(ctx, &ctx->d_counters, &ctx->counters_cap, words * 4)) != RT_SUCCESS) return rc;
    }
    if (bounces > 0) {
        size_t bytes = 0, slots_total = 0;
        for (uint32_t c = 0; c < chunks; ++c) {
            const uint32_t rows = row_begin[c + 1] - row_begin[c];
            const uint64_t tiles = trace_tiles_padded(width, rows), slots = tiles * 32u;
            ray_off[c] = bytes; bytes += align_up(slots * TRACE_QUEUE_ENTRY_BYTES, 256);
            idx_off[c] = bytes; bytes += align_up(slots * 4, 256);
            mask_off[c] = bytes; bytes += align_up((tiles + tiles / 1024 + 16) * 4, 256);
            slot_off[c] = slots_total; slots_total += slots;
        }
        slot_off[chunks] = slots_total;
        if ((rc = ensure(ctx, &ctx->queue, &ctx->queue_cap, bytes)) != RT_SUCCESS) return rc;
        if (ctx->qflags_cap < slots_total) {               // grown: new flags start at 0, an epoch no launch ever uses
            if (ctx->qflags) { RT_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->qflags); ctx->qflags = nullptr; ctx->qflags_cap = 0; }
            RT_CUDA(ctx, cudaMalloc((void**)&ctx->qflags, (slots_total + slots_total / 8) * 4));
            ctx->qflags_cap = slots_total + slots_total / 8;
            RT_CUDA(ctx, cudaMemsetAsync(ctx->qflags, 0, ctx->qflags_cap * 4, ctx->stream));
        }
        P.error_flag = ctx->d_error;
    }
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    if (two_streams) {
        RT_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
        RT_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->fork_ev, 0));
    }
    for (uint32_t c = 0; c < chunks; ++c) {
        TraceParams Pc = P;
        if (++ctx->trace_epoch == 0u) ctx->trace_epoch = 1u;
        Pc.epoch = ctx->trace_epoch;
        Pc.row0 = row_begin[c];
        Pc.local_rows = row_begin[c + 1] - row_begin[c];
        Pc.counters = (uint32_t*)ctx->d_counters + ctr_off[c];
        Pc.region_next = Pc.counters + 16;
        Pc.regions_x = trace_regions_x(width);
        Pc.n_regions = Pc.regions_x * trace_regions_y(Pc.local_rows);
        Pc.n_sm = (uint32_t)ctx->prop.multiProcessorCount;
        if (bounces > 0) {
            Pc.queue = (float4*)((uint8_t*)ctx->queue + ray_off[c]);
            Pc.bounce_index = (uint32_t*)((uint8_t*)ctx->queue + idx_off[c]);
            Pc.tile_mask = (uint32_t*)((uint8_t*)ctx->queue + mask_off[c]);
            Pc.queue_flags = ctx->qflags + slot_off[c];
            Pc.queue_capacity = (uint32_t)(slot_off[c + 1] - slot_off[c]);
        }
        cudaStream_t cs = (two_streams && (c & 1u)) ? ctx->aux_stream : ctx->stream;
        int l = launch_trace(Pc, stats, stack_needed, ctx->prop.multiProcessorCount, cs);
        if (l < 0) return fail(ctx, RT_ERROR_CUDA, "trace launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        ctx->launches += (uint64_t)l;
        if (!dev_out) RT_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[c], cs));
    }
    if (two_streams) {
        RT_CUDA(ctx, cudaEventRecord(ctx->join_ev, ctx->aux_stream));
        RT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_ev, 0));
    }
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (!dev_out) {
        for (uint32_t c = 0; c < chunks; ++c) {
            const size_t p0 = (size_t)row_begin[c] * width;
            const size_t np = (size_t)(row_begin[c + 1] - row_begin[c]) * width;
            RT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[c], 0));
            if (host_scatter) {
                // packed band lb of this part = image rows [(lb * part_count + part_index) * block_rows, + block_rows), clipped to the image
                const uint32_t lb0 = row_begin[c] / block_rows, lb1 = (row_begin[c + 1] + block_rows - 1) / block_rows;
                const size_t band_bytes = (size_t)block_rows * width * 4;
                uint32_t full = 0, tail_rows = 0;
                for (uint32_t lb = lb0; lb < lb1; ++lb) {
                    const uint64_t y0 = ((uint64_t)lb * part_count + part_index) * block_rows;
                    if (y0 + block_rows <= height) ++full; else { if (y0 < height) tail_rows = (uint32_t)(height - y0); break; }
                }
                const uint8_t* src = P.rgba + (size_t)lb0 * band_bytes;
                uint8_t* dst = rgba_out + ((size_t)lb0 * part_count + part_index) * band_bytes;
                if (full) RT_CUDA(ctx, cudaMemcpy2DAsync(dst, band_bytes * part_count, src, band_bytes, band_bytes, full, cudaMemcpyDeviceToHost, ctx->copy_stream));
                if (tail_rows) RT_CUDA(ctx, cudaMemcpyAsync(dst + (size_t)full * band_bytes * part_count, src + (size_t)full * band_bytes, (size_t)tail_rows * width * 4,
                                                            cudaMemcpyDeviceToHost, ctx->copy_stream));
                continue;
            }
            RT_CUDA(ctx, cudaMemcpyAsync(rgba_out + 4 * p0, P.rgba + 4 * p0, np * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
            if (primary_hits_out) RT_CUDA(ctx, cudaMemcpyAsync(primary_hits_out + p0, P.primary_hits + p0, np * sizeof(rt_hit), cudaMemcpyDeviceToHost, ctx->copy_stream));
            if (secondary_hits_out) RT_CUDA(ctx, cudaMemcpyAsync(secondary_hits_out + p0, P.secondary_hits + p0, np * sizeof(rt_hit), cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
        if (async_host) {
            // the watchdog word travels behind the frame; rt_host_frame_wait() looks at it
            if (bounces > 0) {
                RT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[1], 0));
                RT_CUDA(ctx, cudaMemcpyAsync(ctx->h_err_pinned, ctx->d_error, 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
            } else *ctx->h_err_pinned = 0;
            RT_CUDA(ctx, cudaEventRecord(ctx->host_ev, ctx->copy_stream));
            ctx->host_pending = true;
            return RT_SUCCESS;
        }
        RT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    }
    if ((flags & RT_TRACE_ASYNC) && dev_out && !stats) return RT_SUCCESS;
    if (stats) RT_CUDA(ctx, cudaMemcpyAsync(&ctx->last_stats, ctx->d_stats, 64, cudaMemcpyDeviceToHost, ctx->stream));
    int h_err = 0;
    if (bounces > 0) RT_CUDA(ctx, cudaMemcpyAsync(&h_err, ctx->d_error, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_err) { cudaMemsetAsync(ctx->d_error, 0, 4, ctx->stream); return fail(ctx, RT_ERROR_INTERNAL, "trace watchdog fired (bounce-queue entry never published)"); }
    cudaEventElapsedTime(&ctx->last_trace_ms, ctx->ev[0], ctx->ev[1]);
    if (!stats) memset(&ctx->last_stats, 0, sizeof(ctx->last_stats));
    return RT_SUCCESS;
}

int rt_trace_rows(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam, uint32_t width, uint32_t height, uint32_t bounces,
                  uint32_t flags, uint32_t block_rows, uint32_t part_index, uint32_t part_count,
                  uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out) {
    return trace_rows_impl(ctx, tlas, cam, width, height, bounces, flags, block_rows, part_index, part_count, 0, 0, rgba_out, primary_hits_out, secondary_hits_out);
}

int rt_trace_rows_range(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam, uint32_t width, uint32_t height, uint32_t bounces,
                        uint32_t flags, uint32_t block_rows, uint32_t part_index, uint32_t part_count, uint32_t first_row, uint32_t n_rows,
                        uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out) {
    if (!n_rows) return RT_ERROR_INVALID_ARG;
    return trace_rows_impl(ctx, tlas, cam, width, height, bounces, flags, block_rows, part_index, part_count, first_row, n_rows, rgba_out, primary_hits_out, secondary_hits_out);
}

int rt_trace(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam, uint32_t width, uint32_t height, uint32_t bounces,
             uint32_t flags, uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out) {
    if (!height) return RT_ERROR_INVALID_ARG;
    const uint32_t block_rows = (height + 3u) & ~3u;     // one band = the whole image
    return rt_trace_rows(ctx, tlas, cam, width, height, bounces, flags, block_rows, 0, 1, rgba_out, primary_hits_out, secondary_hits_out);
}

int rt_unpack_rows(rt_context* ctx, const uint8_t* packed_all, uint32_t width, uint32_t height, uint32_t block_rows,
                   uint32_t part_count, uint8_t* rgba_out_device) {
    if (!ctx || !packed_all || !rgba_out_device || !width || !height || !block_rows || !part_count) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    int l = launch_unpack_rows(packed_all, width, height, block_rows, part_count, rgba_out_device, ctx->stream);
    if (l < 0) return fail(ctx, RT_ERROR_CUDA, "unpack launch failed");
    ctx->launches += (uint64_t)l;
    RT_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_trace_ms, ctx->ev[0], ctx->ev[1]);
    return RT_SUCCESS;
}

int rt_frame_share_create(rt_context* ctx, uint64_t bytes, void** device_ptr_out, uint8_t handle_out[64]) {
    if (!ctx || !bytes || !device_ptr_out || !handle_out) return RT_ERROR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    RT_CUDA(ctx, cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(ctx, RT_ERROR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
    memcpy(handle_out, &h, 64);
    RT_CUDA(ctx, cudaMemset(p, 0, bytes));
    *device_ptr_out = p;
    return RT_SUCCESS;
}
int rt_frame_share_open(rt_context* ctx, const uint8_t handle[64], void** device_ptr_out) {
    if (!ctx || !handle || !device_ptr_out) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    RT_CUDA(ctx, cudaIpcOpenMemHandle(device_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_SUCCESS;
}
int rt_frame_share_close(rt_context* ctx, void* mapped_device_ptr) {
    if (!ctx || !mapped_device_ptr) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RT_CUDA(ctx, cudaIpcCloseMemHandle(mapped_device_ptr));
    return RT_SUCCESS;
}
int rt_frame_share_free(rt_context* ctx, void* device_ptr) {
    if (!ctx || !device_ptr) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaFree(device_ptr));
    return RT_SUCCESS;
}

int rt_flag_add(rt_context* ctx, uint32_t* counter) {
    if (!ctx || !counter) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (launch_flag_add(counter, ctx->stream) < 0) return fail(ctx, RT_ERROR_CUDA, "flag launch failed");
    ctx->launches += 1;
    return RT_SUCCESS;
}
int rt_flag_wait_ge(rt_context* ctx, const uint32_t* counter, uint32_t target) {
    if (!ctx || !counter) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (launch_flag_wait_ge(counter, target, ctx->d_error, ctx->stream) < 0) return fail(ctx, RT_ERROR_CUDA, "flag launch failed");
    ctx->launches += 1;
    return RT_SUCCESS;
}

int rt_copy_to_host(rt_context* ctx, void* dst_host, const void* src_device, uint64_t bytes) {
    if (!ctx || (bytes && (!dst_host || !src_device))) return RT_ERROR_INVALID_ARG;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bytes) RT_CUDA(ctx, cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_SUCCESS;
}

int rt_last_trace_stats(const rt_context* ctx, rt_trace_stats* out) {
    if (!ctx || !out) return RT_ERROR_INVALID_ARG;
    *out = ctx->last_stats;
    return RT_SUCCESS;
}
float rt_last_trace_ms(const rt_context* ctx) { return ctx ? ctx->last_trace_ms : 0.0f; }
uint64_t rt_kernel_launch_count(const rt_context* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
