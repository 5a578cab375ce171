// sample_scene_mgpu.cpp — the reference's main() (vulkan-raytracing-basic/main.cpp:1425-1454) on the GPUs of one box:
// one PROCESS per GPU (this program forks them itself), the scene replicated, one frame rendered together through the
// render-group part of the C ABI (rt_group_*, include/rtcore.h). Nothing but librtcore and libc is involved: no MPI, no NCCL,
// no Python. Rank 0 writes the frame as PPM/PNG and prints its CRC-32, which must not depend on the GPU count.
//
//   sample_scene_mgpu out.ppm [--gpus N] [--size W H] [--scene file.rtscene] [--frames K] [--device-frame]
//
//   --scene   a scene file written by build_up_phase_b200/scenes.py:save_scene (magic "RTSCENE1": the BLAS geometries, the 64-byte
//             instance records, hit records, camera — exactly the host arrays the sample's createBLAS/createTLAS/
//             createUniformBuffer/createShaderBindingTable fill, main.cpp:674-949,1001-1017,1264-1320). Without it: the sample's scene.
//   --frames  render K frames back to back and report the frame rate (the sample's render loop, main.cpp:1444-1448)
//   --device-frame  assemble in rank 0's device memory over NVLink (RT_GROUP_OUT_DEVICE) and copy the finished frame to the host
//             once; default: every GPU copies its bands over its own PCIe link into the shared pinned host frame (RT_GROUP_OUT_HOST)
//   --split-build  cfg5 style: BLAS b is built by rank b % N only and pulled by the others over NVLink (rt_group_share_blas)
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "rtcore.h"
#include "rtcore_io.h"

struct HostGeometry { std::vector<float> verts; std::vector<uint32_t> idx; std::vector<float> xform; uint32_t flags = RT_GEOMETRY_OPAQUE; };
struct HostScene {
    std::vector<std::vector<HostGeometry>> blases;
    std::vector<rt_instance> instances;              // .blas holds the BLAS index until the handles exist
    std::vector<uint64_t> instance_blas;
    std::vector<float> records;
    rt_camera camera{{0, 0, 10}, 60};
    float miss[3] = {0.0f, 0.0f, 0.2f};
    uint32_t width = 1200, height = 800, bounces = 0;
};

static void sampleScene(HostScene& s) {              // the literals of main.cpp:676-695, 835-858, 1310-1317, 1015
    HostGeometry g;
    g.verts = {-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, 1, 0};
    g.idx = {0, 1, 3, 1, 2, 3};
    HostGeometry g0 = g, g1 = g;
    g0.xform = {1, 0, 0, -2, 0, 1, 0, 0, 0, 0, 1, 0};
    g1.xform = {1, 0, 0, 2, 0, 1, 0, 0, 0, 0, 1, 0};
    s.blases.push_back({g0, g1});
    rt_instance i0{};
    const float t0[12] = {1, 0, 0, 0, 0, 1, 0, 2, 0, 0, 1, 0}, t1[12] = {1, 0, 0, 0, 0, 1, 0, -2, 0, 0, 1, 0};
    i0.custom_index = 100; i0.mask = 0xFF; i0.sbt_offset = 0; i0.flags = RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE;
    rt_instance i1 = i0;
    memcpy(i0.transform, t0, 48); memcpy(i1.transform, t1, 48);
    i1.sbt_offset = 2;
    s.instances = {i0, i1};
    s.instance_blas = {0, 0};
    s.records = {0.6f, 0.1f, 0.2f, 0.1f, 0.8f, 0.4f, 0.9f, 0.7f, 0.1f, 0.3f, 0.6f, 0.9f};
}

template <class T> static void rd(FILE* f, T* p, size_t n) { if (n && fread(p, sizeof(T), n, f) != n) throw std::runtime_error("scene file truncated"); }

static void loadScene(const char* path, HostScene& s) {
    FILE* f = fopen(path, "rb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    char magic[8]; rd(f, magic, 8);
    if (memcmp(magic, "RTSCENE1", 8) != 0) { fclose(f); throw std::runtime_error("not an RTSCENE1 file"); }
    uint32_t h[6]; rd(f, h, 6);                      // n_blas n_instances n_records width height bounces
    float c[7]; rd(f, c, 7);                         // camera xyz, fov, miss rgb
    s.width = h[3]; s.height = h[4]; s.bounces = h[5];
    s.camera = {{c[0], c[1], c[2]}, c[3]};
    s.miss[0] = c[4]; s.miss[1] = c[5]; s.miss[2] = c[6];
    s.blases.resize(h[0]);
    for (auto& b : s.blases) {
        uint32_t ng; rd(f, &ng, 1);
        b.resize(ng);
        for (auto& g : b) {
            uint32_t m[4]; rd(f, m, 4);              // n_verts n_indexed_tris has_xform flags
            g.verts.resize(3ull * m[0]); rd(f, g.verts.data(), g.verts.size());
            g.idx.resize(3ull * m[1]); rd(f, g.idx.data(), g.idx.size());
            if (m[2]) { g.xform.resize(12); rd(f, g.xform.data(), 12); }
            g.flags = m[3];
        }
    }
    s.instances.resize(h[1]); s.instance_blas.resize(h[1]);
    for (uint32_t i = 0; i < h[1]; ++i) {
        uint8_t rec[64]; rd(f, rec, 64);
        memcpy(&s.instances[i], rec, 56);
        memcpy(&s.instance_blas[i], rec + 56, 8);
        s.instances[i].blas = nullptr;
    }
    s.records.resize(3ull * h[2]); rd(f, s.records.data(), s.records.size());
    fclose(f);
}

static uint32_t crc32_of(const uint8_t* p, size_t n) {
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

static int run_rank(int rank, int world, const HostScene& s, const std::string& out, const std::string& group_name, int frames, bool device_frame,
                    bool split_build) {
    rt_context* ctx = nullptr;
    rt_group* group = nullptr;
    auto check = [&](int rc, const char* what) { if (rc != RT_SUCCESS) throw std::runtime_error(std::string(what) + ": " + (group ? rt_group_last_error(group) : "") + " / " + rt_last_error(ctx)); };
    try {
        int n_dev = 1;
        {   // device of this rank = rank modulo the device count (a one-GPU box runs every rank on device 0)
            rt_context* probe = nullptr;
            for (n_dev = 0; n_dev < 64 && rt_create(n_dev, &probe) == RT_SUCCESS; ++n_dev) { rt_destroy(probe); probe = nullptr; }
            if (n_dev == 0) throw std::runtime_error("rt_create failed: no CUDA device (there is no CPU fallback)");
        }
        if (rt_create(rank % n_dev, &ctx) != RT_SUCCESS) throw std::runtime_error("rt_create failed");
        check(rt_group_create(ctx, group_name.c_str(), rank, world, s.width, s.height, &group), "rt_group_create");
        // createBLAS (main.cpp:674-831), once per BLAS; replicated on every rank unless --split-build
        std::vector<rt_blas*> blases(s.blases.size(), nullptr);
        float build_ms = 0.0f;
        auto describe = [&](size_t b, std::vector<rt_geometry>& geoms) {
            for (const HostGeometry& g : s.blases[b]) {
                rt_geometry G{};
                G.vertices = g.verts.data(); G.vertex_count = (uint32_t)(g.verts.size() / 3); G.vertex_stride_bytes = 12;
                G.indices = g.idx.empty() ? nullptr : g.idx.data();
                G.triangle_count = g.idx.empty() ? G.vertex_count / 3 : (uint32_t)(g.idx.size() / 3);
                G.transform3x4 = g.xform.empty() ? nullptr : g.xform.data();
                G.flags = g.flags;
                geoms.push_back(G);
            }
        };
        if (!split_build) {           // every rank builds the whole scene: ONE batched build (one set of launches for all BLASes)
            std::vector<rt_geometry> geoms;
            std::vector<uint32_t> counts;
            for (size_t b = 0; b < s.blases.size(); ++b) { describe(b, geoms); counts.push_back((uint32_t)s.blases[b].size()); }
            check(rt_build_blas_batch(ctx, geoms.data(), counts.data(), (uint32_t)counts.size(), RT_BUILD_PREFER_FAST_TRACE, blases.data()), "rt_build_blas_batch");
            build_ms = rt_last_build_ms(ctx);
        } else {
            for (size_t b = 0; b < s.blases.size(); ++b) {
                const int owner = (int)(b % (size_t)world);
                rt_blas* mine = nullptr;
                if (owner == rank) {
                    std::vector<rt_geometry> geoms;
                    describe(b, geoms);
                    check(rt_build_blas(ctx, geoms.data(), (uint32_t)geoms.size(), RT_BUILD_PREFER_FAST_TRACE, &mine), "rt_build_blas");
                    build_ms += rt_last_build_ms(ctx);
                }
                check(rt_group_share_blas(group, (uint32_t)(b % 64), owner, mine, &blases[b]), "rt_group_share_blas");
                if (b % 64 == 63) check(rt_group_share_finish(group), "rt_group_share_finish");
            }
        }
        if (split_build) check(rt_group_share_finish(group), "rt_group_share_finish");
        // createTLAS (main.cpp:833-949)
        std::vector<rt_instance> inst = s.instances;
        for (size_t i = 0; i < inst.size(); ++i) inst[i].blas = blases[s.instance_blas[i]];
        rt_tlas* tlas = nullptr;
        check(rt_build_tlas(ctx, inst.data(), (uint32_t)inst.size(), RT_BUILD_PREFER_FAST_TRACE, &tlas), "rt_build_tlas");
        check(rt_set_hit_records(ctx, s.records.data(), (uint32_t)(s.records.size() / 3)), "rt_set_hit_records");      // createShaderBindingTable
        check(rt_set_miss_color(ctx, s.miss), "rt_set_miss_color");
        // render loop (main.cpp:1444-1448)
        const uint8_t* frame = nullptr;
        std::vector<uint8_t> host_copy;
        check(rt_group_barrier(group), "rt_group_barrier");
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f) {
            if (device_frame) {
                check(rt_group_trace(group, tlas, &s.camera, s.width, s.height, s.bounces, RT_GROUP_OUT_DEVICE | RT_GROUP_PIPELINE, &frame), "rt_group_trace");
            } else {
                check(rt_group_trace(group, tlas, &s.camera, s.width, s.height, s.bounces, RT_GROUP_OUT_HOST, &frame), "rt_group_trace");
            }
        }
        if (device_frame) check(rt_group_sync(group), "rt_group_sync");
        check(rt_group_barrier(group), "rt_group_barrier");
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / frames;
        if (rank == 0) {
            const size_t bytes = (size_t)s.width * s.height * 4;
            if (device_frame) {       // this program links librtcore only, not the CUDA runtime: the ABI fetches the frame
                host_copy.resize(bytes);
                check(rt_copy_to_host(ctx, host_copy.data(), frame, bytes), "rt_copy_to_host");
                frame = host_copy.data();
            }
            const bool png = out.size() > 4 && out.compare(out.size() - 4, 4, ".png") == 0;
            if ((png ? rt_write_png(out.c_str(), frame, s.width, s.height, 0) : rt_write_ppm(out.c_str(), frame, s.width, s.height, 0)) != RT_SUCCESS)
                throw std::runtime_error(std::string("writing the frame failed: ") + rt_obj_last_error());
            size_t tris = 0;
            for (const auto& b : s.blases) for (const auto& g : b) tris += g.idx.empty() ? g.verts.size() / 9 : g.idx.size() / 3;
            printf("%s: %ux%u, %d GPU process(es), %zu triangles in %zu BLAS, %zu instances, %u bounce(s), crc32 %08x, %.3f ms/frame over %d frame(s) (%s), "
                   "BLAS build %.3f ms on rank 0\n", out.c_str(), s.width, s.height, world, tris, s.blases.size(), s.instances.size(), s.bounces,
                   crc32_of(frame, bytes), ms, frames, device_frame ? "device frame over NVLink" : "shared pinned host frame", build_ms);
        }
        check(rt_group_barrier(group), "rt_group_barrier");
        rt_free_tlas(ctx, tlas);
        for (rt_blas* b : blases) rt_free_blas(ctx, b);
        rt_group_destroy(group);
        rt_destroy(ctx);
    } catch (const std::exception& e) {
        fprintf(stderr, "rank %d: error: %s\n", rank, e.what());
        if (group) rt_group_destroy(group);
        if (ctx) rt_destroy(ctx);
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    std::string out = "sample_scene_mgpu.ppm";
    const char* scene_path = nullptr;
    int gpus = 1, frames = 1;
    uint32_t w = 0, h = 0;
    bool device_frame = false, split_build = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--frames" && i + 1 < argc) frames = atoi(argv[++i]);
        else if (a == "--scene" && i + 1 < argc) scene_path = argv[++i];
        else if (a == "--size" && i + 2 < argc) { w = (uint32_t)atoi(argv[++i]); h = (uint32_t)atoi(argv[++i]); }
        else if (a == "--device-frame") device_frame = true;
        else if (a == "--split-build") split_build = true;
        else out = a;
    }
    if (gpus < 1 || gpus > 64 || frames < 1) { fprintf(stderr, "usage: sample_scene_mgpu out.ppm [--gpus N] [--size W H] [--scene file] [--frames K] [--device-frame] [--split-build]\n"); return 2; }
    HostScene scene;
    try {
        if (scene_path) loadScene(scene_path, scene); else sampleScene(scene);
    } catch (const std::exception& e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
    if (w && h) { scene.width = w; scene.height = h; }
    const std::string group_name = "mgpu-" + std::to_string((long)getpid());
    // one process per GPU; fork BEFORE anything touches CUDA (a CUDA context does not survive fork)
    std::vector<pid_t> kids;
    for (int r = 1; r < gpus; ++r) {
        pid_t p = fork();
        if (p < 0) { perror("fork"); return 1; }
        if (p == 0) _exit(run_rank(r, gpus, scene, out, group_name, frames, device_frame, split_build));
        kids.push_back(p);
    }
    int rc = run_rank(0, gpus, scene, out, group_name, frames, device_frame, split_build);
    for (pid_t p : kids) { int st = 0; waitpid(p, &st, 0); if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = rc ? rc : 1; }
    return rc;
}
