/*
 * rt_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar C++ restatement of the reference's ray-tracing path:
 *   raygen / miss / closest-hit GLSL      vulkan-raytracing-basic/main.cpp:1019-1091
 *   SBT hit-record indexing rule          main.cpp:1260-1262, records :1310-1317
 *   BLAS inputs + per-geometry transform  main.cpp:676-695,726-746,795-804
 *   TLAS instance records                 main.cpp:835-858
 *   camera UBO                            main.cpp:1003-1006,1015
 *   launch size / output format           main.cpp:13-14,1355 / :953,1024,1054
 * Everything the GLSL leaves to the driver (instance transform semantics, built-in ids,
 * barycentric convention, tmin/tmax exclusivity, UNORM8 conversion) follows the Vulkan /
 * GL_EXT_ray_tracing specifications.
 *
 * PARITY UNPINNED BY THE REFERENCE: the path's arithmetic lives in a closed vendor Vulkan
 * driver (no version pinned), the reference ships no tests, golden images or expected values,
 * and no Vulkan toolchain/lavapipe exists in this environment. The oracle is pinned instead
 * against the analytically derived known answers of the sample scene (tests/test_oracle_sample.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product (librtcore.so) never links or calls it.
 *
 * The structs deliberately share the memory layout of include/rtcore.h so one set of numpy
 * arrays feeds both sides.
 */
#ifndef RT_ORACLE_H_
#define RT_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_blas orc_blas;
typedef struct orc_tlas orc_tlas;

typedef struct orc_geometry {
    const float*    vertices;
    uint32_t        vertex_count;
    uint32_t        vertex_stride_bytes;
    const uint32_t* indices;
    uint32_t        triangle_count;
    const float*    transform3x4;
    uint32_t        flags;
} orc_geometry;

typedef struct orc_instance {
    float    transform[12];
    uint32_t custom_index_and_mask;   /* custom_index : 24 (low) | mask : 8 (high) */
    uint32_t sbt_offset_and_flags;    /* sbt_offset : 24 (low) | flags : 8 (high) */
    const orc_blas* blas;
} orc_instance;

typedef struct orc_camera { float pos[3]; float yfov_deg; } orc_camera;

typedef struct orc_ray_params {
    float tmin, tmax; uint32_t cull_mask, sbt_record_offset, sbt_record_stride, bounce_seed;
    uint32_t ray_flags;    /* gl_RayFlags*EXT bits (GL_EXT_ray_tracing); the sample passes Opaque = 0x1 (main.cpp:1048) */
    uint32_t miss_index;   /* missIndex argument of traceRayEXT (main.cpp:1051 passes 0) */
} orc_ray_params;

/* gl_RayFlags*EXT / VkGeometryInstanceFlagBitsKHR / VkGeometryFlagBitsKHR values [spec] */
enum {
    ORC_RAY_OPAQUE = 0x1, ORC_RAY_NO_OPAQUE = 0x2, ORC_RAY_TERMINATE_ON_FIRST_HIT = 0x4, ORC_RAY_SKIP_CLOSEST_HIT = 0x8,
    ORC_RAY_CULL_BACK = 0x10, ORC_RAY_CULL_FRONT = 0x20, ORC_RAY_CULL_OPAQUE = 0x40, ORC_RAY_CULL_NO_OPAQUE = 0x80,
    ORC_INST_FACING_CULL_DISABLE = 0x1, ORC_INST_FLIP_FACING = 0x2, ORC_INST_FORCE_OPAQUE = 0x4, ORC_INST_FORCE_NO_OPAQUE = 0x8,
    ORC_GEOM_OPAQUE = 0x1
};

typedef struct orc_hit {
    uint32_t instance_id, geometry_index, primitive_id, custom_index;
    float t, u, v;
} orc_hit;

typedef struct orc_stats {
    uint64_t rays_primary, rays_secondary, nodes_visited, triangles_tested,
             instances_entered, primary_hits, secondary_hits, near_edge_hits;
} orc_stats;

/* Any-hit shader of a hit group [spec: GL_EXT_ray_tracing any-hit stage; the reference maps the stage at shader_module.h:90 but creates
 * no any-hit shader, main.cpp:1199-1216]. A C ABI cannot carry shader code, so the stage is a fixed-function alpha test described by the
 * record: the candidate's barycentrics (u -> vertex 1, v -> vertex 2) select one cell of a res x res bit mask over [0,1)^2,
 * cell = (min(int(u * res), res - 1), min(int(v * res), res - 1)), bit index = cell_v * res + cell_u (LSB first); bit 0 =
 * ignoreIntersectionEXT, bit 1 = accept; ORC_ANYHIT_TERMINATE additionally ends the ray on an accepted candidate (terminateRayEXT).
 * Runs for NON-opaque candidates only, after the opacity and facing culls, and only for candidates that are not farther than the
 * closest hit committed so far. Indexed like the hit records (main.cpp:1260-1262); an index past the table = no any-hit shader. */
typedef struct orc_anyhit_record { uint32_t kind, log2_res, flags, reserved; const uint32_t* mask; } orc_anyhit_record;
enum { ORC_ANYHIT_ACCEPT = 0, ORC_ANYHIT_ALPHA_MASK = 1, ORC_ANYHIT_TERMINATE = 0x1 };

typedef struct orc_shader_data {
    const float* hit_records_rgb; uint32_t hit_record_count; float miss_rgb[3];
    const float* miss_records_rgb; uint32_t miss_record_count;   /* optional table of constant-colour miss shaders; NULL = {miss_rgb} */
    const orc_anyhit_record* anyhit_records; uint32_t anyhit_record_count;   /* optional; NULL/0 = no any-hit shaders (the sample) */
} orc_shader_data;

enum { ORC_MODE_BRUTE = 0, ORC_MODE_BVH = 1 };

orc_blas* orc_build_blas(const orc_geometry* geoms, uint32_t n_geoms, int build_bvh, int key_bits);
void      orc_free_blas(orc_blas*);
/* refit-only update (the product's RT_BUILD_MODE_REFIT): new vertices, the topology of the last full build; 0 on success */
int       orc_refit_blas(orc_blas*, const orc_geometry* geoms, uint32_t n_geoms);
orc_tlas* orc_build_tlas(const orc_instance* inst, uint32_t n, int build_bvh);
void      orc_free_tlas(orc_tlas*);

/* Traces rows row_begin, row_begin+row_step, ... < row_end of a width x height launch.
 * Outputs are full-size images; only the traced rows are written. Any output may be NULL. */
int orc_trace(const orc_tlas* tlas, const orc_camera* cam, const orc_ray_params* rp,
              const orc_shader_data* sd, uint32_t width, uint32_t height, uint32_t bounces, int mode,
              uint32_t row_begin, uint32_t row_end, uint32_t row_step,
              uint8_t* rgba_out, orc_hit* primary_out, orc_hit* secondary_out, orc_stats* stats_out);

/* LBVH introspection (for bit-exact comparison with the GPU build). */
typedef struct orc_blas_info {
    uint32_t triangle_count, node_count; int32_t root_ref; uint32_t max_depth;
    float bounds_lo[3], bounds_hi[3];
} orc_blas_info;
int orc_blas_get_info(const orc_blas*, orc_blas_info* out);
/* nodes: node_count x 64 B in the product's node layout; tris: triangle_count x 48 B sorted triangles;
 * keys/prims: the sorted Morton keys and original primitive order. Any may be NULL. */
int orc_blas_export(const orc_blas*, void* nodes_out, void* tris_out, uint64_t* keys_out, uint32_t* prims_out);

/* Timed CPU LBVH build of the same input (for the cpu_baseline build leg): returns seconds. */
double orc_time_blas_build(const orc_geometry* geoms, uint32_t n_geoms, int key_bits);

/* Scalar helpers exposed for known-answer tests. */
float    orc_aspect_y(float yfov_deg);
uint32_t orc_pcg_hash(uint32_t v);
uint32_t orc_morton30(float x, float y, float z, const float lo[3], const float hi[3]);
void     orc_unorm8(const float rgb[3], uint8_t out[4]);
int      orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
