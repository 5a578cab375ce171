// rt_group.cu — the render group: the GPUs of one box, one process per GPU, rendering ONE frame together
// (SURVEY 8(b) "rt_create_group", 8(e); the dispatch it scales is vkCmdTraceRaysKHR, reference main.cpp:1349-1355).
//
// The reference is single-GPU (one device, one queue, main.cpp:294-388). Here `world` processes — each with its own rt_context
// and a replica of the scene — cut the image into 8-scanline bands (band b -> rank b % world) and assemble the frame
//   * in rank 0's DEVICE memory: every rank's trace kernel stores its pixels at their final position of rank 0's frame through
//     a CUDA-IPC mapping, i.e. over NVLink / NVSwitch; completion is two stream-ordered counters in the frame's tail
//     (rt_flag_add / rt_flag_wait_ge): no collective, no host round trip, two frames alternate; or
//   * in a pinned HOST frame in POSIX shared memory that every rank registered with CUDA: every GPU copies its bands over its
//     OWN PCIe link (the single 33-MB D2H copy from rank 0 was as long as the 8-GPU trace itself: VERDICT r01, weak 6).
// The ranks meet in a shared-memory control block (no MPI / NCCL / torch in the library). BLAS blobs built per GPU (cfg5) are
// PULLED by the other ranks from the owner's memory over NVLink (IPC mapping + cudaMemcpyAsync on a copy stream), overlapping
// the puller's own builds.
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <new>

#include "rt_host.h"

namespace {

constexpr uint32_t GROUP_MAGIC = 0x52544731u;      // "RTG1"
constexpr uint32_t GROUP_SLOTS = 64;
constexpr double   GROUP_TIMEOUT_S = 60.0;
constexpr uint32_t BAND_ROWS = 8;
constexpr size_t   FRAME_TAIL = 256;               // done / free counters behind the pixels of a device frame

struct SharedBlasSlot {
    std::atomic<uint32_t> seq;
    uint8_t  handle[64];
    uint32_t triangle_count, n_geoms; int32_t root_ref; uint32_t max_depth;
    float    lo[3], hi[3];
    uint64_t storage_bytes;
};

struct GroupShm {
    uint32_t magic, world, max_w, max_h;
    uint64_t created_ns, frame_bytes, host_frame_bytes;
    std::atomic<uint32_t> ready, joined, left, abort;
    std::atomic<uint32_t> bar_count, bar_gen;
    uint8_t frame_handle[2][64];
    std::atomic<uint32_t> host_done[2];             // cumulative: shares of host frame k that have landed
    std::atomic<uint32_t> host_enter[2];            // how often rank 0 has entered a use of host frame k
    SharedBlasSlot slots[GROUP_SLOTS];
};
static_assert(std::atomic<uint32_t>::is_always_lock_free, "lock-free 32-bit atomics are address-free: usable across processes");

double now_s() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
uint64_t wall_ns() { timespec t; clock_gettime(CLOCK_REALTIME, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; }

// spins (tight at first: the frame handshakes are latency critical), then yields; false on timeout or group abort
template <class Pred>
bool wait_until(const GroupShm* shm, Pred pred, double timeout_s = GROUP_TIMEOUT_S) {
    const double t0 = now_s();
    for (uint64_t it = 0;; ++it) {
        if (pred()) return true;
        if (shm && shm->abort.load(std::memory_order_relaxed)) return false;
        if (it < 20000) { __builtin_ia32_pause(); continue; }
        if ((it & 63) == 0 && now_s() - t0 > timeout_s) return false;
        if (it < 200000) sched_yield(); else { timespec ts{0, 50000}; nanosleep(&ts, nullptr); }
    }
}

std::string shm_name_of(const char* name, const char* suffix) {
    std::string s = "/rtcore.";
    for (const char* p = name; *p && s.size() < 200; ++p) s += (isalnum((unsigned char)*p) || *p == '-' || *p == '_' || *p == '.') ? *p : '_';
    return s + suffix;
}

}  // namespace

struct rt_group {
    rt_context* ctx = nullptr;          // the caller's context (may be null: host-only group, no CUDA — barrier + shared host frame)
    rt_context* ctx_b = nullptr;        // second context on the same device: odd frames of a pipelined sequence
    // pipelined host output: the frame that has been enqueued but not completed yet
    bool hp_pending = false; uint32_t hp_k = 0, hp_use = 0; rt_context* hp_ctx = nullptr;
    int rank = 0, world = 1;
    uint32_t max_w = 0, max_h = 0;
    GroupShm* shm = nullptr; size_t shm_bytes = 0;
    uint8_t* host_frames = nullptr; size_t host_frame_bytes = 0; bool host_registered = false;
    void* dev_frame[2] = {nullptr, nullptr};
    uint64_t dev_no = 0, host_no = 0;
    bool host_open = false; uint32_t host_k = 0, host_use = 0;
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
    bool fork_pending = false, forked = false;
    // BLAS exchange
    cudaStream_t pull_stream = nullptr; cudaEvent_t pull_e0 = nullptr, pull_e1 = nullptr;
    bool pull_started = false; float last_share_ms = 0.0f;
    double share_wait_s = 0.0, share_open_s = 0.0, share_alloc_s = 0.0;   // host time of the last finished exchange: waiting for owners, IPC opens, allocations
    double cur_wait_s = 0.0, cur_open_s = 0.0, cur_alloc_s = 0.0;
    uint64_t b_launches_seen = 0;
    uint32_t slot_seq[GROUP_SLOTS] = {};
    std::vector<void*> peer_maps;
    std::vector<BlasRecord*> staged_records;
    std::string err;
};

namespace {

int gfail(rt_group* g, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (g) { g->err = buf; if (g->ctx) g->ctx->err = buf; if (g->shm && code != RT_SUCCESS) g->shm->abort.store(1); }
    return code;
}
#define G_CUDA(g, call)                                                                                        \
    do {                                                                                                       \
        cudaError_t e__ = (call);                                                                              \
        if (e__ != cudaSuccess) return gfail(g, RT_ERROR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int host_barrier(rt_group* g) {
    GroupShm* s = g->shm;
    if (g->world == 1) return RT_SUCCESS;
    const uint32_t gen = s->bar_gen.load(std::memory_order_acquire);
    if (s->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)g->world) {
        s->bar_count.store(0, std::memory_order_relaxed);
        s->bar_gen.store(gen + 1, std::memory_order_release);
        return RT_SUCCESS;
    }
    if (!wait_until(s, [&] { return s->bar_gen.load(std::memory_order_acquire) != gen; }))
        return gfail(g, RT_ERROR_INTERNAL, "render group barrier timed out (rank %d of %d)", g->rank, g->world);
    return RT_SUCCESS;
}

void* map_shm(const std::string& name, size_t bytes, bool create, int* err_out) {
    int fd;
    if (create) {
        shm_unlink(name.c_str());                    // a stale object of a crashed run
        fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) { *err_out = errno; if (fd >= 0) close(fd); return nullptr; }
    } else {
        fd = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd < 0) { *err_out = errno; return nullptr; }
        struct stat st;
        if (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes) { *err_out = EAGAIN; close(fd); return nullptr; }   // not sized yet
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { *err_out = errno; return nullptr; }
    return p;
}

uint32_t* done_counter(rt_group* g, int k) { return (uint32_t*)((uint8_t*)g->dev_frame[k] + (size_t)g->max_w * g->max_h * 4); }
uint32_t* free_counter(rt_group* g, int k) { return done_counter(g, k) + 1; }

int join_streams(rt_group* g) {
    if (g->forked && g->ctx && g->ctx_b) {
        G_CUDA(g, cudaEventRecord(g->join_ev, g->ctx_b->stream));
        G_CUDA(g, cudaStreamWaitEvent(g->ctx->stream, g->join_ev, 0));
    }
    g->forked = false; g->fork_pending = false;
    return RT_SUCCESS;
}

}  // namespace

extern "C" {

int rt_group_rank(const rt_group* g) { return g ? g->rank : -1; }
int rt_group_world(const rt_group* g) { return g ? g->world : 0; }
float rt_group_last_share_ms(const rt_group* g) { return g ? g->last_share_ms : 0.0f; }
int rt_group_last_share_host_ms(const rt_group* g, float out[3]) {
    if (!g || !out) return RT_ERROR_INVALID_ARG;
    out[0] = (float)(g->share_wait_s * 1e3); out[1] = (float)(g->share_open_s * 1e3); out[2] = (float)(g->share_alloc_s * 1e3);
    return RT_SUCCESS;
}
const char* rt_group_last_error(const rt_group* g) { return g ? g->err.c_str() : "no group"; }

void rt_group_destroy(rt_group* g) {
    if (!g) return;
    if (g->ctx) {
        cudaSetDevice(g->ctx->device);
        join_streams(g);
        cudaStreamSynchronize(g->ctx->stream);
        if (g->ctx_b) cudaStreamSynchronize(g->ctx_b->stream);
        if (g->pull_stream) { cudaStreamSynchronize(g->pull_stream); }
        for (void* p : g->peer_maps) cudaIpcCloseMemHandle(p);
        for (BlasRecord* r : g->staged_records) cudaFreeHost(r);
        if (g->rank != 0) { for (int k = 0; k < 2; ++k) if (g->dev_frame[k]) cudaIpcCloseMemHandle(g->dev_frame[k]); }
    }
    if (g->shm) {
        if (g->rank != 0) g->shm->left.fetch_add(1, std::memory_order_acq_rel);
        else if (g->world > 1) wait_until(nullptr, [&] { return g->shm->left.load(std::memory_order_acquire) >= (uint32_t)g->world - 1u; }, 10.0);   // peers unmap first, then the owner frees
    }
    if (g->ctx) {
        if (g->rank == 0) for (int k = 0; k < 2; ++k) if (g->dev_frame[k]) cudaFree(g->dev_frame[k]);
        if (g->host_registered) cudaHostUnregister(g->host_frames);
        if (g->pull_stream) cudaStreamDestroy(g->pull_stream);
        if (g->pull_e0) cudaEventDestroy(g->pull_e0);
        if (g->pull_e1) cudaEventDestroy(g->pull_e1);
        if (g->fork_ev) cudaEventDestroy(g->fork_ev);
        if (g->join_ev) cudaEventDestroy(g->join_ev);
        if (g->ctx_b) rt_destroy(g->ctx_b);
    }
    if (g->host_frames) munmap(g->host_frames, 2 * g->host_frame_bytes);
    if (g->shm) munmap(g->shm, g->shm_bytes);
    delete g;
}

int rt_group_create(rt_context* ctx, const char* name, int rank, int world, uint32_t max_width, uint32_t max_height, rt_group** out) {
    if (!out) return RT_ERROR_INVALID_ARG;
    *out = nullptr;
    if (!name || !*name || world < 1 || world > 64 || rank < 0 || rank >= world || !max_width || !max_height) return rt_fail(ctx, RT_ERROR_INVALID_ARG, "rt_group_create: bad arguments");
    rt_group* g = new rt_group();
    g->ctx = ctx; g->rank = rank; g->world = world; g->max_w = max_width; g->max_h = max_height;
    g->shm_bytes = (sizeof(GroupShm) + 4095) / 4096 * 4096;
    g->host_frame_bytes = ((size_t)max_width * max_height * 4 + 4095) / 4096 * 4096;
    const std::string ctl = shm_name_of(name, ".ctl"), frm = shm_name_of(name, ".frame");
    int e = 0;
    auto bail = [&](int code, const char* what) { int rc = gfail(g, code, "rt_group_create (rank %d of %d, '%s'): %s (errno %d)", rank, world, name, what, e); if (ctx) ctx->err = g->err; rt_group_destroy(g); return rc; };
    if (ctx && cudaSetDevice(ctx->device) != cudaSuccess) return bail(RT_ERROR_CUDA, "cudaSetDevice");
    const size_t frame_bytes = (size_t)max_width * max_height * 4 + FRAME_TAIL;
    if (rank == 0) {
        g->shm = (GroupShm*)map_shm(ctl, g->shm_bytes, true, &e);
        if (!g->shm) return bail(RT_ERROR_INTERNAL, "cannot create the control block in /dev/shm");
        memset((void*)g->shm, 0, g->shm_bytes);
        new (g->shm) GroupShm();
        g->shm->world = (uint32_t)world; g->shm->max_w = max_width; g->shm->max_h = max_height;
        g->shm->frame_bytes = frame_bytes; g->shm->host_frame_bytes = g->host_frame_bytes; g->shm->created_ns = wall_ns();
        g->host_frames = (uint8_t*)map_shm(frm, 2 * g->host_frame_bytes, true, &e);
        if (!g->host_frames) return bail(RT_ERROR_INTERNAL, "cannot create the shared host frame in /dev/shm");
        if (ctx) {
            for (int k = 0; k < 2; ++k) {
                if (cudaMalloc(&g->dev_frame[k], frame_bytes) != cudaSuccess || cudaMemset(g->dev_frame[k], 0, frame_bytes) != cudaSuccess) return bail(RT_ERROR_OUT_OF_MEMORY, "cudaMalloc of the device frame");
                if (world > 1) {
                    cudaIpcMemHandle_t h;
                    if (cudaIpcGetMemHandle(&h, g->dev_frame[k]) != cudaSuccess) return bail(RT_ERROR_CUDA, "cudaIpcGetMemHandle");
                    memcpy(g->shm->frame_handle[k], &h, 64);
                }
            }
        }
        g->shm->magic = GROUP_MAGIC;
        g->shm->ready.store(1, std::memory_order_release);
    } else {
        const double t0 = now_s();
        for (;;) {
            g->shm = (GroupShm*)map_shm(ctl, g->shm_bytes, false, &e);
            if (g->shm) {
                // rank 0 of THIS run: ready, right shape, created recently (a stale block of a crashed run is skipped until rank 0 replaces it)
                const bool ok = wait_until(nullptr, [&] { return g->shm->ready.load(std::memory_order_acquire) == 1u; }, 0.05) && g->shm->magic == GROUP_MAGIC &&
                                g->shm->world == (uint32_t)world && g->shm->max_w == max_width && g->shm->max_h == max_height &&
                                wall_ns() - g->shm->created_ns < 600ull * 1000000000ull && g->shm->joined.load() < (uint32_t)world;
                if (ok) break;
                munmap(g->shm, g->shm_bytes); g->shm = nullptr;
            }
            if (now_s() - t0 > GROUP_TIMEOUT_S) return bail(RT_ERROR_INTERNAL, "rank 0's control block did not appear within 60 s");
            timespec ts{0, 2000000}; nanosleep(&ts, nullptr);
        }
        g->host_frames = (uint8_t*)map_shm(frm, 2 * g->host_frame_bytes, false, &e);
        if (!g->host_frames) return bail(RT_ERROR_INTERNAL, "cannot map the shared host frame");
        if (ctx) {
            for (int k = 0; k < 2; ++k) {
                cudaIpcMemHandle_t h;
                memcpy(&h, g->shm->frame_handle[k], 64);
                if (cudaIpcOpenMemHandle(&g->dev_frame[k], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return bail(RT_ERROR_CUDA, "cudaIpcOpenMemHandle of rank 0's frame (is CUDA IPC / P2P available between the GPUs?)"); }
            }
        }
    }
    if (ctx) {
        // the shared host frame becomes pinned memory of THIS process's CUDA context: every rank copies over its own PCIe link
        if (cudaHostRegister(g->host_frames, 2 * g->host_frame_bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return bail(RT_ERROR_CUDA, "cudaHostRegister of the shared host frame"); }
        g->host_registered = true;
        bool ok = cudaEventCreateWithFlags(&g->fork_ev, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&g->join_ev, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&g->pull_stream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&g->pull_e0) == cudaSuccess && cudaEventCreate(&g->pull_e1) == cudaSuccess;
        if (!ok) return bail(RT_ERROR_CUDA, "stream / event creation");
        if (rt_create(ctx->device, &g->ctx_b) != RT_SUCCESS) return bail(RT_ERROR_CUDA, "second context");
        g->ctx_b->own_hit_records = false;
    }
    g->shm->joined.fetch_add(1, std::memory_order_acq_rel);
    if (!wait_until(g->shm, [&] { return g->shm->joined.load(std::memory_order_acquire) >= (uint32_t)world; })) return bail(RT_ERROR_INTERNAL, "not all ranks joined within 60 s");
    if (rank == 0) { shm_unlink(ctl.c_str()); shm_unlink(frm.c_str()); }     // the mappings keep the objects alive; nothing is left behind in /dev/shm
    *out = g;
    return RT_SUCCESS;
}

int rt_group_barrier(rt_group* g) { return g ? host_barrier(g) : RT_ERROR_INVALID_ARG; }

int rt_group_join(rt_group* g) {
    if (!g) return RT_ERROR_INVALID_ARG;
    if (g->ctx) G_CUDA(g, cudaSetDevice(g->ctx->device));
    return join_streams(g);
}

int rt_group_sync(rt_group* g) {
    if (!g || !g->ctx) return RT_ERROR_INVALID_ARG;
    int rc = rt_group_flush_host(g, nullptr);
    if (rc != RT_SUCCESS) return rc;
    rc = rt_group_join(g);
    if (rc != RT_SUCCESS) return rc;
    if (g->ctx_b && (rc = rt_sync(g->ctx_b)) != RT_SUCCESS) { g->ctx->err = g->ctx_b->err; return rc; }
    return rt_sync(g->ctx);
}

// ---- host frame handshake (also usable on its own: a caller may fill its bands of the shared host frame by other means) ------------
int rt_group_host_frame_begin(rt_group* g, uint8_t** frame_out) {
    if (!g || g->host_open) return RT_ERROR_INVALID_ARG;
    if (g->hp_pending) { const int rc = rt_group_flush_host(g, nullptr); if (rc != RT_SUCCESS) return rc; }
    const uint32_t k = (uint32_t)(g->host_no & 1u), use = (uint32_t)(g->host_no >> 1);
    ++g->host_no;
    // buffer k still holds the frame of `use - 1`, which rank 0's caller may read until rank 0 enters this use
    if (g->rank == 0) g->shm->host_enter[k].store(use + 1u, std::memory_order_release);
    else if (!wait_until(g->shm, [&] { return g->shm->host_enter[k].load(std::memory_order_acquire) >= use + 1u; }))
        return gfail(g, RT_ERROR_INTERNAL, "rank %d: rank 0 never entered frame %llu", g->rank, (unsigned long long)(g->host_no - 1));
    g->host_open = true; g->host_k = k; g->host_use = use;
    if (frame_out) *frame_out = g->host_frames + (size_t)k * g->host_frame_bytes;
    return RT_SUCCESS;
}

int rt_group_host_frame_end(rt_group* g, const uint8_t** frame_out) {
    if (!g || !g->host_open) return RT_ERROR_INVALID_ARG;
    g->host_open = false;
    const uint32_t k = g->host_k, use = g->host_use;
    g->shm->host_done[k].fetch_add(1, std::memory_order_acq_rel);              // this rank's bands are in the shared frame
    if (g->rank == 0) {
        if (!wait_until(g->shm, [&] { return g->shm->host_done[k].load(std::memory_order_acquire) >= (uint32_t)g->world * (use + 1u); }))
            return gfail(g, RT_ERROR_INTERNAL, "host frame %u incomplete after 60 s: %u of %u shares", k, g->shm->host_done[k].load(), (uint32_t)g->world * (use + 1u));
        if (frame_out) *frame_out = g->host_frames + (size_t)k * g->host_frame_bytes;
    } else if (frame_out) *frame_out = nullptr;
    return RT_SUCCESS;
}

// completes the pending frame of a pipelined host-output sequence: this rank's copies have landed (event), its share is announced,
// rank 0 waits for everybody's
int rt_group_flush_host(rt_group* g, const uint8_t** frame_out) {
    if (!g) return RT_ERROR_INVALID_ARG;
    if (frame_out) *frame_out = nullptr;
    if (!g->hp_pending) return RT_SUCCESS;
    g->hp_pending = false;
    int rc = rt_host_frame_wait(g->hp_ctx);
    if (rc != RT_SUCCESS) { g->shm->abort.store(1); g->err = g->hp_ctx->err; g->ctx->err = g->hp_ctx->err; return rc; }
    const uint32_t k = g->hp_k, use = g->hp_use;
    g->shm->host_done[k].fetch_add(1, std::memory_order_acq_rel);
    if (g->rank == 0) {
        if (!wait_until(g->shm, [&] { return g->shm->host_done[k].load(std::memory_order_acquire) >= (uint32_t)g->world * (use + 1u); }))
            return gfail(g, RT_ERROR_INTERNAL, "host frame %u incomplete after 60 s: %u of %u shares", k, g->shm->host_done[k].load(), (uint32_t)g->world * (use + 1u));
        if (frame_out) *frame_out = g->host_frames + (size_t)k * g->host_frame_bytes;
    }
    return RT_SUCCESS;
}

int rt_group_trace(rt_group* g, const rt_tlas* tlas, const rt_camera* cam, uint32_t width, uint32_t height, uint32_t bounces,
                   uint32_t flags, const uint8_t** frame_out) {
    if (!g || !g->ctx || !tlas || !cam) return RT_ERROR_INVALID_ARG;
    if (frame_out) *frame_out = nullptr;
    if (width > g->max_w || height > g->max_h || (uint64_t)width * height > (uint64_t)g->max_w * g->max_h)
        return gfail(g, RT_ERROR_INVALID_ARG, "frame %ux%u exceeds the group's %ux%u", width, height, g->max_w, g->max_h);
    const bool to_host = (flags & RT_GROUP_OUT_HOST) != 0;
    if (to_host == ((flags & RT_GROUP_OUT_DEVICE) != 0)) return gfail(g, RT_ERROR_INVALID_ARG, "exactly one of RT_GROUP_OUT_DEVICE / RT_GROUP_OUT_HOST");
    G_CUDA(g, cudaSetDevice(g->ctx->device));
    int rc;
    if (to_host && (flags & RT_GROUP_PIPELINE)) {
        // ---- two host frames in flight: enqueue frame k, then complete frame k - 1 ----
        if (g->host_open) return gfail(g, RT_ERROR_INVALID_ARG, "a host frame opened with rt_group_host_frame_begin is still open");
        const uint32_t k = (uint32_t)(g->host_no & 1u), use = (uint32_t)(g->host_no >> 1);
        ++g->host_no;
        // buffer k holds frame `k - 2` of the sequence, which rank 0's caller may read until rank 0 enters this use
        if (g->rank == 0) g->shm->host_enter[k].store(use + 1u, std::memory_order_release);
        else if (!wait_until(g->shm, [&] { return g->shm->host_enter[k].load(std::memory_order_acquire) >= use + 1u; }))
            return gfail(g, RT_ERROR_INTERNAL, "rank %d: rank 0 never entered frame %llu", g->rank, (unsigned long long)(g->host_no - 1));
        rt_context* c = (k == 1u && g->ctx_b) ? g->ctx_b : g->ctx;
        if (c != g->ctx && (rc = rt_context_mirror_shader_state(c, g->ctx)) != RT_SUCCESS) return rc;
        uint8_t* frame = g->host_frames + (size_t)k * g->host_frame_bytes;
        rc = rt_trace_rows(c, tlas, cam, width, height, bounces, RT_TRACE_OUT_FULL_FRAME | RT_TRACE_ASYNC, BAND_ROWS, (uint32_t)g->rank, (uint32_t)g->world, frame, nullptr, nullptr);
        if (c != g->ctx) { g->ctx->launches += c->launches - g->b_launches_seen; g->b_launches_seen = c->launches; }
        if (rc != RT_SUCCESS) { g->shm->abort.store(1); g->err = c->err; g->ctx->err = c->err; return rc; }
        const uint8_t* done_frame = nullptr;
        if ((rc = rt_group_flush_host(g, &done_frame)) != RT_SUCCESS) return rc;        // frame k - 1
        if (frame_out) *frame_out = done_frame;
        g->hp_pending = true; g->hp_k = k; g->hp_use = use; g->hp_ctx = c;
        return RT_SUCCESS;
    }
    if (g->hp_pending && (rc = rt_group_flush_host(g, nullptr)) != RT_SUCCESS) return rc;
    if (to_host) {
        if ((rc = join_streams(g)) != RT_SUCCESS) return rc;
        uint8_t* frame = nullptr;
        if ((rc = rt_group_host_frame_begin(g, &frame)) != RT_SUCCESS) return rc;
        // synchronous: returns when this rank's bands are in the shared host frame (row chunks copied while the next one is traced)
        rc = rt_trace_rows(g->ctx, tlas, cam, width, height, bounces, RT_TRACE_OUT_FULL_FRAME, BAND_ROWS, (uint32_t)g->rank, (uint32_t)g->world, frame, nullptr, nullptr);
        if (rc != RT_SUCCESS) { g->host_open = false; g->shm->abort.store(1); g->err = g->ctx->err; return rc; }
        return rt_group_host_frame_end(g, frame_out);
    }
    // ---- device frame on rank 0, written by every rank's kernel over NVLink ----
    const bool pipeline = (flags & RT_GROUP_PIPELINE) != 0;
    const int k = (int)(g->dev_no & 1u);
    const uint32_t use = (uint32_t)(g->dev_no >> 1);
    ++g->dev_no;
    rt_context* c = g->ctx;
    if (pipeline && k == 1) {
        c = g->ctx_b;
        if ((rc = rt_context_mirror_shader_state(c, g->ctx)) != RT_SUCCESS) return rc;
        if (!g->forked) {
            if (!g->fork_pending) G_CUDA(g, cudaEventRecord(g->fork_ev, g->ctx->stream));     // (no even frame before it: nothing to overlap with yet)
            G_CUDA(g, cudaStreamWaitEvent(c->stream, g->fork_ev, 0));
            g->forked = true; g->fork_pending = false;
        }
    } else if (pipeline) {
        if (!g->forked && !g->fork_pending) { G_CUDA(g, cudaEventRecord(g->fork_ev, g->ctx->stream)); g->fork_pending = true; }   // BEFORE this frame: the odd frame may overlap it
    } else if ((rc = join_streams(g)) != RT_SUCCESS) return rc;
    uint32_t* done = done_counter(g, k);
    uint32_t* free_ = free_counter(g, k);
    // the frame this buffer held (use - 1) was the caller's until now: rank 0 hands the buffer back, everybody waits for that
    if (g->rank == 0 && use > 0 && (rc = rt_flag_add(c, free_)) != RT_SUCCESS) { g->err = c->err; return rc; }
    if ((rc = rt_flag_wait_ge(c, free_, use)) != RT_SUCCESS) { g->err = c->err; return rc; }
    rc = rt_trace_rows(c, tlas, cam, width, height, bounces, RT_TRACE_OUT_DEVICE | RT_TRACE_OUT_FULL_FRAME | RT_TRACE_ASYNC, BAND_ROWS,
                       (uint32_t)g->rank, (uint32_t)g->world, (uint8_t*)g->dev_frame[k], nullptr, nullptr);
    if (rc != RT_SUCCESS) { g->err = c->err; g->ctx->err = c->err; g->shm->abort.store(1); return rc; }
    if ((rc = rt_flag_add(c, done)) != RT_SUCCESS) { g->err = c->err; return rc; }          // this rank's pixels have landed in rank 0's memory
    if (c != g->ctx) { g->ctx->launches += c->launches - g->b_launches_seen; g->b_launches_seen = c->launches; }   // rt_kernel_launch_count(ctx) covers the group's second stream
    if (g->rank == 0) {
        if ((rc = rt_flag_wait_ge(c, done, (uint32_t)g->world * (use + 1u))) != RT_SUCCESS) { g->err = c->err; return rc; }   // ... and everybody else's
        if (frame_out) *frame_out = (const uint8_t*)g->dev_frame[k];
    }
    if (!(flags & (RT_GROUP_ASYNC | RT_GROUP_PIPELINE))) return rt_group_sync(g);
    return RT_SUCCESS;
}

// ---- BLAS exchange: pull over NVLink -----------------------------------------------------------------------------------------------
int rt_group_share_blas(rt_group* g, uint32_t slot, int owner_rank, const rt_blas* mine, rt_blas** out) {
    if (!g || !g->ctx || !out || slot >= GROUP_SLOTS || owner_rank < 0 || owner_rank >= g->world) return RT_ERROR_INVALID_ARG;
    *out = nullptr;
    G_CUDA(g, cudaSetDevice(g->ctx->device));
    SharedBlasSlot& S = g->shm->slots[slot];
    const uint32_t seq = ++g->slot_seq[slot];
    if (g->rank == owner_rank) {
        if (!mine || !mine->st || mine->st->n_blas != 1) return gfail(g, RT_ERROR_INVALID_ARG, "rt_group_share_blas: the owner passes a BLAS that was built on its own (not part of a batch)");
        if (g->world > 1) {
            cudaIpcMemHandle_t h;
            G_CUDA(g, cudaIpcGetMemHandle(&h, mine->st->dev));
            memcpy(S.handle, &h, 64);
            S.triangle_count = mine->r().tri_count; S.n_geoms = mine->r().n_geoms; S.root_ref = mine->r().root; S.max_depth = mine->r().height;
            for (int k = 0; k < 3; ++k) { S.lo[k] = mine->r().lo[k]; S.hi[k] = mine->r().hi[k]; }
            S.storage_bytes = mine->st->bytes;
            S.seq.store(seq, std::memory_order_release);
        }
        *out = const_cast<rt_blas*>(mine);
        return RT_SUCCESS;
    }
    double t0 = now_s();
    if (!wait_until(g->shm, [&] { return S.seq.load(std::memory_order_acquire) >= seq; }))
        return gfail(g, RT_ERROR_INTERNAL, "rank %d: rank %d never published BLAS slot %u", g->rank, owner_rank, slot);
    double t1 = now_s();
    g->cur_wait_s += t1 - t0;
    cudaIpcMemHandle_t h;
    memcpy(&h, S.handle, 64);
    void* peer = nullptr;
    G_CUDA(g, cudaIpcOpenMemHandle(&peer, h, cudaIpcMemLazyEnablePeerAccess));
    g->peer_maps.push_back(peer);
    t0 = now_s();
    g->cur_open_s += t0 - t1;
    const uint32_t N = S.triangle_count;
    BlasStorage* st = new BlasStorage();
    st->n_tris = N; st->n_blas = 1;
    const size_t nodes_b = align_up(sizeof(BvhNode) * (size_t)N, 256), tris_b = align_up(sizeof(TriRec) * (size_t)N, 256);
    st->bytes = nodes_b + tris_b + sizeof(BlasRecord) + 256;
    if (st->bytes != S.storage_bytes) { delete st; return gfail(g, RT_ERROR_INTERNAL, "BLAS blob size mismatch"); }
    if (cudaMalloc(&st->dev, st->bytes) != cudaSuccess) { delete st; cudaGetLastError(); return gfail(g, RT_ERROR_OUT_OF_MEMORY, "cudaMalloc(%zu) for the pulled BLAS failed", (size_t)S.storage_bytes); }
    st->nodes = (BvhNode*)st->dev; st->tris = (TriRec*)((uint8_t*)st->dev + nodes_b); st->records = (BlasRecord*)((uint8_t*)st->dev + nodes_b + tris_b);
    st->refs = 1;
    g->cur_alloc_s += now_s() - t0;
    rt_blas* hnd = new rt_blas();
    hnd->st = st; hnd->index = 0;
    hnd->st->host_recs.assign(1, BlasRecord{});
    BlasRecord& R = hnd->r();
    memset(&R, 0, sizeof(R));
    R.nodes = st->nodes; R.tris = st->tris; R.root = S.root_ref; R.height = S.max_depth; R.tri_count = N; R.n_geoms = S.n_geoms; R.first = 0; R.node_slots = N;
    for (int k = 0; k < 3; ++k) { R.lo[k] = S.lo[k]; R.hi[k] = S.hi[k]; }
    BlasRecord* staged = nullptr;                      // pinned: the record copy is asynchronous like the blob's
    if (cudaMallocHost((void**)&staged, sizeof(BlasRecord)) != cudaSuccess) { rt_free_blas(g->ctx, hnd); return gfail(g, RT_ERROR_OUT_OF_MEMORY, "cudaMallocHost"); }
    *staged = R;
    g->staged_records.push_back(staged);
    if (!g->pull_started) { G_CUDA(g, cudaEventRecord(g->pull_e0, g->pull_stream)); g->pull_started = true; }
    cudaError_t e1 = cudaMemcpyAsync(st->dev, peer, nodes_b + tris_b, cudaMemcpyDeviceToDevice, g->pull_stream);      // nodes | triangles, over NVLink
    cudaError_t e2 = cudaMemcpyAsync(st->records, staged, sizeof(BlasRecord), cudaMemcpyHostToDevice, g->pull_stream);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { rt_free_blas(g->ctx, hnd); return gfail(g, RT_ERROR_CUDA, "BLAS pull failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)); }
    *out = hnd;
    return RT_SUCCESS;
}

int rt_group_share_finish(rt_group* g) {
    if (!g || !g->ctx) return RT_ERROR_INVALID_ARG;
    G_CUDA(g, cudaSetDevice(g->ctx->device));
    g->last_share_ms = 0.0f;
    if (g->pull_started) {
        G_CUDA(g, cudaEventRecord(g->pull_e1, g->pull_stream));
        G_CUDA(g, cudaStreamSynchronize(g->pull_stream));
        cudaEventElapsedTime(&g->last_share_ms, g->pull_e0, g->pull_e1);
        g->pull_started = false;
    }
    for (void* p : g->peer_maps) cudaIpcCloseMemHandle(p);
    g->peer_maps.clear();
    for (BlasRecord* r : g->staged_records) cudaFreeHost(r);
    g->staged_records.clear();
    g->share_wait_s = g->cur_wait_s; g->share_open_s = g->cur_open_s; g->share_alloc_s = g->cur_alloc_s;
    g->cur_wait_s = g->cur_open_s = g->cur_alloc_s = 0.0;
    return host_barrier(g);                                // after it owners may free / update what they shared
}

}  // extern "C"
