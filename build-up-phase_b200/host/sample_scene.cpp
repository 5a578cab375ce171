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
// Usage: sample_scene [out.ppm] [width height]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "rtcore.h"

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
    const char* out = argc > 1 ? argv[1] : "sample_scene.ppm";
    const uint32_t width = argc > 3 ? (uint32_t)atoi(argv[2]) : WIDTH;
    const uint32_t height = argc > 3 ? (uint32_t)atoi(argv[3]) : HEIGHT;
    try {
        if (rt_create(0, &vk.ctx) != RT_SUCCESS) throw std::runtime_error("rt_create failed: no CUDA device (there is no CPU fallback)");
        createBLAS();
        createTLAS();
        createUniformBuffer();
        createShaderBindingTable();
        std::vector<uint8_t> frame;
        render(frame, width, height);
        size_t hits = 0;
        for (size_t p = 0; p < (size_t)width * height; ++p)
            if (!(frame[4 * p] == 0 && frame[4 * p + 1] == 0 && frame[4 * p + 2] == 51)) ++hits;
        FILE* f = fopen(out, "wb");
        if (!f) throw std::runtime_error("cannot open output file");
        fprintf(f, "P6\n%u %u\n255\n", width, height);
        for (size_t p = 0; p < (size_t)width * height; ++p) fwrite(&frame[4 * p], 1, 3, f);
        fclose(f);
        printf("%s: %ux%u, %zu non-miss pixels, trace kernel %.3f ms, build %.3f ms, %llu kernel launches\n", out, width, height, hits,
               rt_last_trace_ms(vk.ctx), rt_last_build_ms(vk.ctx), (unsigned long long)rt_kernel_launch_count(vk.ctx));
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
