// sample_scene.cpp — headless C++ host program above the C ABI: the same call sequence as the
// reference's main() (vulkan-raytracing-basic/main.cpp:1425-1454) with the Vulkan nouns removed.
//
//   createBLAS()                  main.cpp:674-831   -> rt_build_blas
//   createTLAS()                  main.cpp:833-949   -> rt_build_tlas
//   createUniformBuffer()         main.cpp:1001-1017 -> rt_camera
//   createShaderBindingTable()    main.cpp:1244-1320 -> rt_set_hit_records
//   render()                      main.cpp:1322-1423 -> rt_trace
//
// No GLFW window, swapchain, pipeline, descriptor sets or SPIR-V: the shaders are compiled into
// librtcore's trace kernel. The frame is written as a binary PPM instead of being presented.
//
// Build: g++ -std=c++17 sample_scene.cpp -I../../include -L.. -lrtcore -Wl,-rpath,'$ORIGIN/..' -o sample_scene
// Usage: sample_scene [out.ppm | out.png] [width height] [--obj mesh.obj] [--srgb]
//   --obj: the reference's next assignment (vulkan-raytracing-basic/README.md:225-226): "load an obj file -> build the
//          acceleration structure" — the mesh replaces the two quads (one geometry per o/g group, one instance, fitted
//          into the view); everything else is the sample's pipeline.
//   --srgb: encode like the sample's B8G8R8A8_SRGB swapchain does on store (main.cpp:50).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "rtcore.h"
#include "rtcore_io.h"

static const uint32_t WIDTH = 1200;    // main.cpp:13
static const uint32_t HEIGHT = 800;    // main.cpp:14

struct Global {
    rt_context* ctx = nullptr;
    rt_blas* blas = nullptr;
    rt_tlas* tlas = nullptr;
    rt_camera camera{};
    ~Global() {
        rt_free_tlas(ctx, tlas);
        rt_free_blas(ctx, blas);
        rt_destroy(ctx);
    }
} vk;

static void check(int rc, const char* what) {
    if (rc != RT_SUCCESS) throw std::runtime_error(std::string(what) + ": " + rt_last_error(vk.ctx));
}

void createBLAS() {
    float vertices[][3] = {
        {-1.0f, -1.0f, 0.0f},
        {1.0f, -1.0f, 0.0f},
        {1.0f, 1.0f, 0.0f},
        {-1.0f, 1.0f, 0.0f},
    };
    uint32_t indices[] = {0, 1, 3, 1, 2, 3};
    float geoTransforms[2][12] = {
        {1.0f, 0.0f, 0.0f, -2.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f},
        {1.0f, 0.0f, 0.0f, 2.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f},
    };
    rt_geometry geometry0{};
    geometry0.vertices = &vertices[0][0];
    geometry0.vertex_count = sizeof(vertices) / sizeof(vertices[0]);       // maxVertex + 1
    geometry0.vertex_stride_bytes = sizeof(vertices[0]);
    geometry0.indices = indices;
    geometry0.triangle_count = sizeof(indices) / (sizeof(indices[0]) * 3);
    geometry0.flags = RT_GEOMETRY_OPAQUE;
    rt_geometry geometries[] = {geometry0, geometry0};                     // same buffers twice, main.cpp:743
    geometries[0].transform3x4 = geoTransforms[0];                         // transformOffset 0
    geometries[1].transform3x4 = geoTransforms[1];                         // transformOffset sizeof(geoTransforms[0])

    uint32_t triangleCounts[] = {geometry0.triangle_count, geometry0.triangle_count};
    rt_build_sizes requiredSize{};
    check(rt_blas_build_sizes(vk.ctx, triangleCounts, 2, &requiredSize), "rt_blas_build_sizes");
    check(rt_build_blas(vk.ctx, geometries, 2, RT_BUILD_PREFER_FAST_TRACE, &vk.blas), "rt_build_blas");
    // vertices / indices / transforms go out of scope here, exactly like the reference frees its
    // input buffers right after the build (main.cpp:823-830): the BLAS is self-contained.
}

// --obj: one BLAS with one geometry per group of the file; returns the scale/offset that fits it into a 4x4x4 box at the origin
void createBLASFromObj(const char* path, float fit[4]) {
    rt_obj_mesh* mesh = nullptr;
    if (rt_obj_load(path, &mesh) != RT_SUCCESS) throw std::runtime_error(std::string("rt_obj_load: ") + rt_obj_last_error());
    const uint32_t groups = rt_obj_group_count(mesh);
    if (groups == 0) { rt_obj_free(mesh); throw std::runtime_error("the .obj file has no faces"); }
    std::vector<rt_geometry> geometries(groups);
    for (uint32_t g = 0; g < groups; ++g) rt_obj_geometry(mesh, g, &geometries[g]);
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    const float* v = rt_obj_vertices(mesh);
    for (uint32_t i = 0; i < rt_obj_vertex_count(mesh); ++i)
        for (int k = 0; k < 3; ++k) { lo[k] = v[3 * i + k] < lo[k] ? v[3 * i + k] : lo[k]; hi[k] = v[3 * i + k] > hi[k] ? v[3 * i + k] : hi[k]; }
    float ext = 0.0f;
    for (int k = 0; k < 3; ++k) ext = hi[k] - lo[k] > ext ? hi[k] - lo[k] : ext;
    fit[0] = ext > 0.0f ? 4.0f / ext : 1.0f;
    for (int k = 0; k < 3; ++k) fit[1 + k] = -0.5f * (lo[k] + hi[k]) * fit[0];
    check(rt_build_blas(vk.ctx, geometries.data(), groups, RT_BUILD_PREFER_FAST_TRACE, &vk.blas), "rt_build_blas");
    printf("%s: %u vertices, %u triangles, %u group(s)\n", path, rt_obj_vertex_count(mesh), rt_obj_triangle_count(mesh), groups);
    rt_obj_free(mesh);      // the BLAS is self-contained (main.cpp:823-830)
}

void createTLASForObj(const float fit[4]) {
    rt_instance instance0{};
    const float t[12] = {fit[0], 0, 0, fit[1], 0, fit[0], 0, fit[2], 0, 0, fit[0], fit[3]};
    for (int k = 0; k < 12; ++k) instance0.transform[k] = t[k];
    instance0.custom_index = 0; instance0.mask = 0xFF; instance0.sbt_offset = 0;
    instance0.flags = RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE;
    instance0.blas = vk.blas;
    check(rt_build_tlas(vk.ctx, &instance0, 1, RT_BUILD_PREFER_FAST_TRACE, &vk.tlas), "rt_build_tlas");
}

void createTLAS() {
    float insTransforms[2][12] = {
        {1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 2.0f, 0.0f, 0.0f, 1.0f, 0.0f},
        {1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, -2.0f, 0.0f, 0.0f, 1.0f, 0.0f},
    };
    rt_instance instance0{};
    instance0.custom_index = 100;
    instance0.mask = 0xFF;
    instance0.sbt_offset = 0;
    instance0.flags = RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE;
    instance0.blas = vk.blas;
    rt_instance instanceData[] = {instance0, instance0};
    for (int k = 0; k < 12; ++k) { instanceData[0].transform[k] = insTransforms[0][k]; instanceData[1].transform[k] = insTransforms[1][k]; }
    instanceData[1].sbt_offset = 2;   // 2 geometry (in instance0) + 2 geometry (in instance1)
    check(rt_build_tlas(vk.ctx, instanceData, 2, RT_BUILD_PREFER_FAST_TRACE, &vk.tlas), "rt_build_tlas");
}

void createUniformBuffer() { vk.camera = {{0, 0, 10}, 60}; }

void createShaderBindingTable() {
    // only the payloads of the four hit-group records survive (handles are driver tokens)
    const float hitgCustomData[4][3] = {
        {0.6f, 0.1f, 0.2f},   // Deep Red Wine
        {0.1f, 0.8f, 0.4f},   // Emerald Green
        {0.9f, 0.7f, 0.1f},   // Golden Yellow
        {0.3f, 0.6f, 0.9f},   // Dawn Sky Blue
    };
    check(rt_set_hit_records(vk.ctx, &hitgCustomData[0][0], 4), "rt_set_hit_records");
    const float miss[3] = {0.0f, 0.0f, 0.2f};
    check(rt_set_miss_color(vk.ctx, miss), "rt_set_miss_color");
}

void render(std::vector<uint8_t>& frame, uint32_t width, uint32_t height) {
    frame.resize((size_t)width * height * 4);
    check(rt_trace(vk.ctx, vk.tlas, &vk.camera, width, height, 0, 0, frame.data(), nullptr, nullptr), "rt_trace");
}

int main(int argc, char** argv) {
    const char* obj = nullptr;
    uint32_t image_flags = 0;
    std::vector<const char*> pos;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--obj" && i + 1 < argc) obj = argv[++i];
        else if (a == "--srgb") image_flags |= RT_IMAGE_SRGB_ENCODE;
        else pos.push_back(argv[i]);
    }
    const char* out = pos.size() > 0 ? pos[0] : "sample_scene.ppm";
    const uint32_t width = pos.size() > 2 ? (uint32_t)atoi(pos[1]) : WIDTH;
    const uint32_t height = pos.size() > 2 ? (uint32_t)atoi(pos[2]) : HEIGHT;
    try {
        if (rt_create(0, &vk.ctx) != RT_SUCCESS) throw std::runtime_error("rt_create failed: no CUDA device (there is no CPU fallback)");
        if (obj) {
            float fit[4];
            createBLASFromObj(obj, fit);
            createTLASForObj(fit);
        } else {
            createBLAS();
            createTLAS();
        }
        createUniformBuffer();
        createShaderBindingTable();
        if (obj) {      // geometry g of the single instance uses hit record g: repeat the sample's four colours
            std::vector<float> records;
            const float c[4][3] = {{0.6f, 0.1f, 0.2f}, {0.1f, 0.8f, 0.4f}, {0.9f, 0.7f, 0.1f}, {0.3f, 0.6f, 0.9f}};
            for (uint32_t g = 0; g < 4096; ++g) for (int k = 0; k < 3; ++k) records.push_back(c[g & 3][k]);
            check(rt_set_hit_records(vk.ctx, records.data(), 4096), "rt_set_hit_records");
        }
        std::vector<uint8_t> frame;
        render(frame, width, height);
        size_t hits = 0;
        for (size_t p = 0; p < (size_t)width * height; ++p)
            if (!(frame[4 * p] == 0 && frame[4 * p + 1] == 0 && frame[4 * p + 2] == 51)) ++hits;
        const std::string outs = out;
        const bool png = outs.size() > 4 && outs.compare(outs.size() - 4, 4, ".png") == 0;
        if ((png ? rt_write_png(out, frame.data(), width, height, image_flags) : rt_write_ppm(out, frame.data(), width, height, image_flags)) != RT_SUCCESS)
            throw std::runtime_error(std::string("writing the frame failed: ") + rt_obj_last_error());
        printf("%s: %ux%u, %zu non-miss pixels, trace kernel %.3f ms, build %.3f ms, %llu kernel launches\n", out, width, height, hits,
               rt_last_trace_ms(vk.ctx), rt_last_build_ms(vk.ctx), (unsigned long long)rt_kernel_launch_count(vk.ctx));
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
