/*
 * rtcore.h — C ABI of the B200-native ray-tracing core.
 *
 * This is the drop-in boundary for what the reference samples
 * (vulkan-raytracing-basic/main.cpp == vulkan-raytraced-triangle/main.cpp) hand to the
 * Vulkan driver through the KHR entry-point table they load at main.cpp:25-33,135-145:
 *
 *   vkGetAccelerationStructureBuildSizesKHR   main.cpp:757,893   -> rt_blas_build_sizes / rt_tlas_build_sizes
 *   vkCreateAccelerationStructureKHR +
 *   vkCmdBuildAccelerationStructuresKHR (BLAS) main.cpp:775-820  -> rt_build_blas / rt_build_blas_batch
 *   vkCmdBuildAccelerationStructuresKHR (TLAS) main.cpp:911-942  -> rt_build_tlas
 *   UBO {cameraPos, yFov_degree}              main.cpp:1001-1017 -> rt_camera
 *   SBT hit-record payloads / miss constant   main.cpp:1310-1317,1065 -> rt_set_hit_records / rt_set_miss_color
 *   traceRayEXT arguments                     main.cpp:1047-1052 -> rt_set_ray_params
 *   vkCmdTraceRaysKHR(W,H,1) + fence          main.cpp:1333,1349-1355,1411 -> rt_trace / rt_trace_rows
 *   vkDestroyAccelerationStructureKHR         main.cpp:89-96     -> rt_free_blas / rt_free_tlas
 *
 * Plain C: pointers and sizes only, no C++/torch types. All calls are synchronous on return
 * (the reference does vkQueueWaitIdle after every build, main.cpp:820,942) unless the name
 * ends in _async. A context is bound to one CUDA device and is not thread-safe.
 * No CPU fallback exists: every entry point fails with RT_ERROR_CUDA when no device is usable.
 */
#ifndef RTCORE_H_
#define RTCORE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API __attribute__((visibility("default")))

/* ---- status codes -------------------------------------------------------------------- */
enum {
    RT_SUCCESS            =  0,
    RT_ERROR_INVALID_ARG  = -1,
    RT_ERROR_CUDA         = -2,   /* a CUDA runtime call failed; see rt_last_error() */
    RT_ERROR_OUT_OF_MEMORY= -3,
    RT_ERROR_STACK_DEPTH  = -4,   /* BVH deeper than the traversal stack (never silently wrong) */
    RT_ERROR_SBT_RANGE    = -5,   /* an instance/geometry addresses a hit record that was not set */
    RT_ERROR_INTERNAL     = -6    /* device-side watchdog / invariant violation */
};

/* ---- geometry flags (rt_geometry.flags) ------------------------------------------------ */
#define RT_GEOMETRY_OPAQUE          0x1u  /* VK_GEOMETRY_OPAQUE_BIT_KHR, main.cpp:741 */
#define RT_GEOMETRY_NO_DUPLICATE_ANY_HIT 0x2u /* VK_GEOMETRY_NO_DUPLICATE_ANY_HIT_INVOCATION_BIT_KHR: accepted; the any-hit records below are pure functions of the candidate */
#define RT_GEOMETRY_DEVICE_POINTERS 0x100u /* vertices/indices/transform are CUDA device pointers */

/* ---- build flags ---------------------------------------------------------------------- */
#define RT_BUILD_ALLOW_UPDATE       0x1u  /* VK_BUILD_ACCELERATION_STRUCTURE_ALLOW_UPDATE_BIT_KHR: the BLAS keeps its sorted Morton records (8 B per triangle)
                                            so that rt_update_blas(RT_BUILD_MODE_REFIT) can re-fit the boxes without re-sorting */
#define RT_BUILD_ALLOW_COMPACTION   0x2u  /* VK_BUILD_ACCELERATION_STRUCTURE_ALLOW_COMPACTION_BIT_KHR: the BLAS (or batch) keeps its sorted Morton records
                                            until rt_compact_blas() has packed its live nodes */
#define RT_BUILD_PREFER_FAST_TRACE  0x4u  /* VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_TRACE_BIT_KHR, main.cpp:751,887: what the reference asks for and what
                                            every build here does by default (two-triangle leaves: measured 1/2/3/4/6/8 per leaf -> 3124/3236/3198/3145/2981/2823 Mrays/s) */
#define RT_BUILD_PREFER_FAST_BUILD  0x8u  /* VK_BUILD_ACCELERATION_STRUCTURE_PREFER_FAST_BUILD_BIT_KHR: accepted; the LBVH build is the fast build already */
#define RT_BUILD_MODE_REFIT         0x1000u /* rt_update_blas only: VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR proper - keep the tree TOPOLOGY of the
                                              last full build (needs RT_BUILD_ALLOW_UPDATE at that build) and only re-fit the boxes to the new vertices */
#define RT_BUILD_INSTANCES_ON_DEVICE 0x100u /* rt_build_tlas: the rt_instance array is DEVICE memory (what the reference's instance buffer is, main.cpp:860-868);
                                              its `blas` fields must then hold rt_blas_device_reference() values, not host handles */
#define RT_BUILD_NO_PACKED_SORT     0x200u /* keep (key, id) pairs in separate arrays during the sort (the general path; test hook) */
#define RT_BUILD_NO_FUSED_SETUP     0x800u /* segmented sort, but triangle setup and Morton keys as separate kernels (test hook) */
#define RT_BUILD_NO_SEGMENTED_SORT  0x400u /* always use the global onesweep sort, also for BLASes that fit one CTA's shared memory (test hook) */

/* ---- instance flags (rt_instance.flags, low 8 bits like VkGeometryInstanceFlagsKHR) ---- */
#define RT_INSTANCE_TRIANGLE_FACING_CULL_DISABLE 0x1u /* main.cpp:852 */
#define RT_INSTANCE_TRIANGLE_FLIP_FACING         0x2u /* VK_GEOMETRY_INSTANCE_TRIANGLE_FLIP_FACING_BIT_KHR (= ..._FRONT_COUNTERCLOCKWISE) */
#define RT_INSTANCE_FORCE_OPAQUE                 0x4u /* VK_GEOMETRY_INSTANCE_FORCE_OPAQUE_BIT_KHR */
#define RT_INSTANCE_FORCE_NO_OPAQUE              0x8u /* VK_GEOMETRY_INSTANCE_FORCE_NO_OPAQUE_BIT_KHR */

/* ---- ray flags (rt_ray_params.ray_flags): the gl_RayFlags*EXT argument of traceRayEXT (main.cpp:1048 passes Opaque) ----
 * Opacity of a candidate: geometry RT_GEOMETRY_OPAQUE, overridden by the instance FORCE_* flags, overridden by the ray
 * OPAQUE / NO_OPAQUE flags. Opacity feeds CULL_OPAQUE / CULL_NO_OPAQUE, and a surviving NON-opaque candidate runs the any-hit
 * record of its hit group (rt_set_anyhit_records; without records - the sample registers none - it is accepted like an opaque one).
 * Facing (object space, so the baked geometry transform counts and the instance transform does not): a triangle is
 * FRONT facing when its vertices appear clockwise from the ray origin, i.e. ((v1-v0) x (v2-v0)) . dir > 0, inverted by
 * RT_INSTANCE_TRIANGLE_FLIP_FACING; facing culls are ignored for instances with TRIANGLE_FACING_CULL_DISABLE.
 * TERMINATE_ON_FIRST_HIT accepts the first candidate the traversal meets: hit/miss is defined, WHICH hit is not. */
#define RT_RAY_FLAG_OPAQUE                      0x01u
#define RT_RAY_FLAG_NO_OPAQUE                   0x02u
#define RT_RAY_FLAG_TERMINATE_ON_FIRST_HIT      0x04u
#define RT_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER     0x08u /* a hit leaves the payload at its initial value (0,0,0), main.cpp:1045 */
#define RT_RAY_FLAG_CULL_BACK_FACING_TRIANGLES  0x10u
#define RT_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES 0x20u
#define RT_RAY_FLAG_CULL_OPAQUE                 0x40u
#define RT_RAY_FLAG_CULL_NO_OPAQUE              0x80u

/* ---- rt_trace output flags -------------------------------------------------------------- */
#define RT_TRACE_OUT_DEVICE 0x1u  /* rgba_out / hit buffers are device pointers */
#define RT_TRACE_STATS      0x2u  /* also accumulate traversal counters (slower kernel variant) */
#define RT_TRACE_ASYNC      0x4u  /* with RT_TRACE_OUT_DEVICE: enqueue on the context stream and return (rt_sync() to wait).
                                    With a HOST framebuffer (pinned memory, no hit buffers): trace and device->host copies are enqueued and the
                                    call returns; rt_host_frame_wait() returns when the frame is complete in the host buffer. One such frame
                                    may be in flight per context (the next host-output trace waits for it first). */
#define RT_TRACE_OUT_BGRA   0x10u /* store B,G,R,A byte order — the sample copies its image raw into a B8G8R8A8 swapchain (main.cpp:50,971-974,1371-1375);
                                    the default is the logical R,G,B,A of imageStore(vec4(hitValue, 0.0)) */
#define RT_TRACE_OUT_FULL_FRAME 0x8u /* rt_trace_rows: rgba_out is the WHOLE width*height*4 frame and this part's pixels land at their final
                                        position (no packed buffer to gather, no unpack step).
                                        With RT_TRACE_OUT_DEVICE the trace kernel stores them there itself; the frame may be peer memory of
                                        another GPU (rt_frame_share_open): the kernel then writes over NVLink straight into rank 0's framebuffer.
                                        With a HOST rgba_out (e.g. a frame in shared pinned host memory that every rank's process maps) the
                                        part's bands are copied to their final rows, chunk by chunk while the next chunk is traced: every GPU
                                        moves its share of the image over its own PCIe link. */

typedef struct rt_context rt_context;
typedef struct rt_blas    rt_blas;
typedef struct rt_tlas    rt_tlas;

/* One triangle geometry of a BLAS: mirrors VkAccelerationStructureGeometryTrianglesDataKHR
 * (R32G32B32_SFLOAT vertices, UINT32 indices, optional 3x4 row-major VkTransformMatrixKHR)
 * plus the primitiveCount of its VkAccelerationStructureBuildRangeInfoKHR (main.cpp:726-746,795-804). */
typedef struct rt_geometry {
    const float*    vertices;            /* vertex_count x (vertex_stride_bytes/4) floats        */
    uint32_t        vertex_count;        /* maxVertex + 1                                        */
    uint32_t        vertex_stride_bytes; /* >= 12, multiple of 4                                 */
    const uint32_t* indices;             /* 3*triangle_count indices, or NULL = non-indexed list  */
    uint32_t        triangle_count;      /* primitiveCount                                       */
    const float*    transform3x4;        /* row-major 3x4 applied at build time, or NULL         */
    uint32_t        flags;               /* RT_GEOMETRY_*                                        */
} rt_geometry;

/* 64-byte TLAS instance record, same layout as VkAccelerationStructureInstanceKHR (main.cpp:848-858)
 * with the BLAS handle in place of accelerationStructureReference. */
typedef struct rt_instance {
    float    transform[12];              /* object->world, row-major 3x4 */
    uint32_t custom_index : 24;          /* instanceCustomIndex */
    uint32_t mask         : 8;
    uint32_t sbt_offset   : 24;          /* instanceShaderBindingTableRecordOffset */
    uint32_t flags        : 8;
    const rt_blas* blas;                 /* accelerationStructureReference */
} rt_instance;

/* UBO of the raygen shader (main.cpp:1003-1006,1015): {0,0,10, 60} in the sample. */
typedef struct rt_camera {
    float pos[3];
    float yfov_deg;
} rt_camera;

/* traceRayEXT parameters (main.cpp:1047-1052). Defaults are the sample's. */
typedef struct rt_ray_params {
    float    tmin;              /* 0.0   */
    float    tmax;              /* 100.0 */
    uint32_t cull_mask;         /* 0xff  */
    uint32_t sbt_record_offset; /* 0     */
    uint32_t sbt_record_stride; /* 1     */
    uint32_t bounce_seed;       /* seed of the deterministic diffuse bounce (ours; default 1) */
    uint32_t ray_flags;         /* RT_RAY_FLAG_*; default RT_RAY_FLAG_OPAQUE (main.cpp:1048) */
    uint32_t miss_index;        /* which miss record runs on a miss (main.cpp:1051 passes 0) */
} rt_ray_params;

/* Per-ray result: the built-ins the closest-hit shader sees (main.cpp:1080-1086).
 * A miss has all four ids == 0xFFFFFFFF and t == tmax. */
typedef struct rt_hit {
    uint32_t instance_id;     /* gl_InstanceID (slot in the rt_instance array) */
    uint32_t geometry_index;  /* gl_GeometryIndexEXT */
    uint32_t primitive_id;    /* gl_PrimitiveID (index within its geometry) */
    uint32_t custom_index;    /* gl_InstanceCustomIndexEXT */
    float    t;               /* along the un-normalised ray direction */
    float    u, v;            /* hitAttributeEXT barycentrics: weights of vertex 1 and vertex 2 */
} rt_hit;

typedef struct rt_build_sizes {           /* VkAccelerationStructureBuildSizesInfoKHR */
    uint64_t acceleration_structure_size;
    uint64_t build_scratch_size;
} rt_build_sizes;

/* Device-side traversal counters (RT_TRACE_STATS); summed over all rays of the call. */
typedef struct rt_trace_stats {
    uint64_t rays_primary;
    uint64_t rays_secondary;
    uint64_t nodes_visited;       /* 64-byte BVH nodes fetched (TLAS + BLAS) */
    uint64_t triangles_tested;
    uint64_t instances_entered;
    uint64_t primary_hits;
    uint64_t secondary_hits;
    uint64_t near_edge_hits;      /* hits with min barycentric < 2^-20: the "within epsilon of a shared edge" count */
} rt_trace_stats;

/* Phase timings of the most recent build on this context, CUDA-event milliseconds. */
typedef struct rt_build_timing {
    float total_ms;        /* first setup kernel .. last refit kernel (no H2D) */
    float setup_ms;        /* triangle fetch + transform + bounds */
    float morton_ms;
    float sort_ms;
    float hierarchy_ms;
    float refit_ms;        /* leaf emit + atomic bottom-up refit */
    float h2d_ms;          /* staging copies of host inputs (outside total_ms) */
    uint64_t primitives;
} rt_build_timing;

/* ---- context ---------------------------------------------------------------------------- */
RT_API int  rt_create(int device_ordinal, rt_context** out);
RT_API void rt_destroy(rt_context* ctx);
RT_API const char* rt_last_error(const rt_context* ctx);       /* never NULL */
RT_API int  rt_device_info(const rt_context* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* Run subsequent work of this context on an existing CUDA stream (cudaStream_t as void*), e.g. torch's. */
RT_API int  rt_set_stream(rt_context* ctx, void* cuda_stream);
RT_API int  rt_sync(rt_context* ctx);
/* The context keeps its build scratch, framebuffer staging and bounce-ray buffers between calls (they only grow: ~250 B per triangle
 * of the largest build, ~60 B per pixel of the largest frame). rt_release_scratch frees them; the next call re-allocates. */
RT_API int  rt_release_scratch(rt_context* ctx);

/* ---- acceleration structures ------------------------------------------------------------ */
/* vkGetAccelerationStructureBuildSizesKHR (main.cpp:756-762, 892-898): upper bounds of what a build of these primitive counts
 * allocates, computed by the helpers the builds themselves use. acceleration_structure_size == rt_blas_get_info().storage_bytes
 * (resp. rt_tlas_storage_bytes) of the built structure; build_scratch_size >= rt_last_build_scratch_bytes() for host inputs of
 * 12-byte vertices with at most 3 vertices per triangle (device-pointer inputs need less: they are not staged). */
RT_API int  rt_blas_build_sizes(rt_context* ctx, const uint32_t* max_triangle_counts, uint32_t n_geoms, rt_build_sizes* out);
RT_API int  rt_tlas_build_sizes(rt_context* ctx, uint32_t max_instances, rt_build_sizes* out);
/* Scratch bytes the most recent build on this context needed (the context's grow-only scratch buffer is at least that large). */
RT_API uint64_t rt_last_build_scratch_bytes(const rt_context* ctx);
RT_API uint64_t rt_tlas_storage_bytes(const rt_context* ctx, const rt_tlas* tlas);

/* Builds one BLAS over n_geoms geometries. Inputs are copied/consumed before return: the caller
 * may free them immediately (the reference does, main.cpp:823-830). */
RT_API int  rt_build_blas(rt_context* ctx, const rt_geometry* geoms, uint32_t n_geoms, uint32_t build_flags, rt_blas** out);

/* Builds n_blas independent BLASes in ONE set of kernel launches (segmented LBVH).
 * geoms holds all geometries back to back; geom_counts[b] of them belong to BLAS b. */
RT_API int  rt_build_blas_batch(rt_context* ctx, const rt_geometry* geoms, const uint32_t* geom_counts,
                                uint32_t n_blas, uint32_t build_flags, rt_blas** out_array);

/* VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR for a BLAS built by rt_build_blas: same geometry and triangle counts,
 * new vertex data / transforms (the per-frame loop of an animated mesh, main.cpp:1444-1448), inside the existing device
 * allocation; the handle stays valid. As in Vulkan, a TLAS that references it must be rebuilt or updated afterwards.
 *   default:              a full LBVH re-build (new Morton order, new topology): the tree quality of a fresh build.
 *   RT_BUILD_MODE_REFIT:  the topology of the last full build is kept (its sorted Morton records were retained because that build
 *                         had RT_BUILD_ALLOW_UPDATE); only the triangle records and every node box are recomputed, by the same
 *                         bottom-up pass the build uses. Skips the Morton and sort phases; tree quality degrades as the mesh
 *                         moves away from the pose it was sorted for (measured: profiles/README.md r2_i). */
RT_API int  rt_update_blas(rt_context* ctx, rt_blas* blas, const rt_geometry* geoms, uint32_t n_geoms, uint32_t build_flags);

RT_API int  rt_build_tlas(rt_context* ctx, const rt_instance* instances, uint32_t n_instances, uint32_t build_flags, rt_tlas** out);
/* The per-frame path: re-BUILDS the TLAS over new instance records inside the existing handle (and, when the count does not grow,
 * inside its existing device allocation). A full rebuild of 1024 instances costs ~0.07 ms, so no refit-only shortcut is taken. */
RT_API int  rt_update_tlas(rt_context* ctx, rt_tlas* tlas, const rt_instance* instances, uint32_t n_instances, uint32_t build_flags);

/* accelerationStructureReference of a BLAS (vkGetAccelerationStructureDeviceAddressKHR, main.cpp:785 `vk.blasAddress`): the 64-bit device
 * address a DEVICE-resident rt_instance array (RT_BUILD_INSTANCES_ON_DEVICE) must carry in its `blas` field. 0 on error. */
RT_API uint64_t rt_blas_device_reference(const rt_context* ctx, const rt_blas* blas);

/* VK_COPY_ACCELERATION_STRUCTURE_MODE_COMPACT_KHR. The build leaves the node of every collapsed subtree (<= 2 triangles: about a
 * third of the Karras slots) unused; compaction packs the live nodes (order-preserving, so siblings stay adjacent), moves the BLAS -
 * or the whole batch it was built in: every handle of the batch stays valid - into a right-sized allocation (64 B x live nodes +
 * 48 B x triangles) and releases the old one. Needs RT_BUILD_ALLOW_COMPACTION at the build. Device addresses change: TLASes that
 * reference the BLAS(es) must be rebuilt afterwards, as in Vulkan. *bytes_before / *bytes_after (may be NULL) report the storage. */
RT_API int  rt_compact_blas(rt_context* ctx, rt_blas* blas, uint64_t* bytes_before, uint64_t* bytes_after);

RT_API void rt_free_blas(rt_context* ctx, rt_blas* blas);
RT_API void rt_free_tlas(rt_context* ctx, rt_tlas* tlas);

RT_API int  rt_last_build_timing(const rt_context* ctx, rt_build_timing* out);
RT_API float rt_last_build_ms(const rt_context* ctx);

/* Introspection used by the parity tests (build invariants) and by multi-GPU BLAS broadcast. */
typedef struct rt_blas_info {
    uint32_t triangle_count;
    uint32_t node_count;        /* 64-byte node slots stored (= triangle_count after a build, = live nodes after rt_compact_blas) */
    int32_t  root_ref;          /* >=0 internal node, <0 leaf, RT_REF_EMPTY for an empty BLAS */
    uint32_t max_depth;
    float    bounds_lo[3], bounds_hi[3];
    uint64_t storage_bytes;     /* relocatable device blob: nodes then triangles */
    void*    device_storage;    /* device pointer of that blob */
} rt_blas_info;
#define RT_REF_EMPTY 0x7FFFFFFD
RT_API int  rt_blas_get_info(rt_context* ctx, const rt_blas* blas, rt_blas_info* out);
/* Copies nodes (node_count x 64 B) and triangles (triangle_count x 48 B) to host buffers; either may be NULL. */
RT_API int  rt_blas_export(rt_context* ctx, const rt_blas* blas, void* nodes_out, void* tris_out);
/* Debug/parity export of the sorted Morton keys+primitive ids of the most recent single-BLAS build. */
RT_API int  rt_debug_last_sorted_keys(rt_context* ctx, uint64_t* keys_out, uint32_t* prim_out, uint32_t capacity, uint32_t* n_out);
/* Wraps a device blob produced by another context/GPU's rt_blas_get_info().device_storage
 * (after ncclBroadcast) as a BLAS on this context. The blob is copied. */
RT_API int  rt_blas_import(rt_context* ctx, const rt_blas_info* info, const void* device_blob, rt_blas** out);

typedef struct rt_tlas_info {
    uint32_t instance_count;
    uint32_t node_count;
    int32_t  root_ref;
    uint32_t max_depth;
    float    bounds_lo[3], bounds_hi[3];
} rt_tlas_info;
RT_API int  rt_tlas_get_info(rt_context* ctx, const rt_tlas* tlas, rt_tlas_info* out);

/* ---- shader data ------------------------------------------------------------------------ */
/* Payloads of the hit-group records: count x {r,g,b} (main.cpp:1310-1317). */
RT_API int  rt_set_hit_records(rt_context* ctx, const float* rgb, uint32_t count);
RT_API int  rt_set_miss_color(rt_context* ctx, const float rgb[3]);          /* miss record 0; default (0,0,0.2), main.cpp:1065 */
/* Several miss shaders (each one, like the sample's, writes a constant colour): count x {r,g,b}; rt_ray_params.miss_index selects. */
RT_API int  rt_set_miss_records(rt_context* ctx, const float* rgb, uint32_t count);
RT_API int  rt_set_ray_params(rt_context* ctx, const rt_ray_params* params); /* NULL restores the defaults */

/* The ANY_HIT stage of the hit groups (the reference maps the stage at shader_module.h:90 and creates no any-hit shader,
 * main.cpp:1199-1216). A C ABI cannot carry shader code, so the stage is a fixed-function alpha test described by the record of the
 * hit group, indexed like the hit records (instanceSbtOffset + geometryIndex * sbtRecordStride + sbtRecordOffset, main.cpp:1260-1262;
 * an index past the table = no any-hit shader = accept):
 *   RT_ANYHIT_ACCEPT      no any-hit shader in this hit group.
 *   RT_ANYHIT_ALPHA_MASK  the candidate's barycentrics (u -> vertex 1, v -> vertex 2) select one cell of a res x res bit mask over
 *                         [0,1)^2, res = 1 << log2_res <= 1024: cell = (min(int(u * res), res - 1), min(int(v * res), res - 1)),
 *                         bit index = cell_v * res + cell_u (LSB first). Bit 0 = ignoreIntersectionEXT, bit 1 = accept.
 *   flags RT_ANYHIT_TERMINATE_RAY: an accepted candidate also ends the ray (terminateRayEXT): hit/miss stays defined, WHICH hit does not.
 * It runs for NON-opaque candidates only, after the opacity and facing culls, and only for candidates not farther than the closest
 * hit committed so far (the current ray interval). Masks are host memory, copied by the call. count 0 removes the table. */
#define RT_ANYHIT_ACCEPT        0u
#define RT_ANYHIT_ALPHA_MASK    1u
#define RT_ANYHIT_TERMINATE_RAY 0x1u
typedef struct rt_anyhit_record {
    uint32_t kind;             /* RT_ANYHIT_* */
    uint32_t log2_res;         /* 0..10 */
    uint32_t flags;            /* RT_ANYHIT_TERMINATE_RAY */
    uint32_t reserved;
    const uint32_t* mask;      /* (res * res + 31) / 32 words, host memory; NULL with RT_ANYHIT_ACCEPT */
} rt_anyhit_record;
RT_API int  rt_set_anyhit_records(rt_context* ctx, const rt_anyhit_record* records, uint32_t count);

/* ---- dispatch --------------------------------------------------------------------------- */
/* vkCmdTraceRaysKHR(width, height, 1): raygen + traversal + closest-hit/miss + imageStore.
 * rgba_out: width*height*4 bytes, logical R,G,B,A with A = 0 (imageStore(vec4(hitValue,0.0)), main.cpp:1054).
 * bounces: 0 = the reference's pipeline (recursion depth 1); 1 = plus one deterministic diffuse bounce.
 * primary_hits_out / secondary_hits_out: width*height rt_hit each, or NULL. */
RT_API int  rt_trace(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam,
                     uint32_t width, uint32_t height, uint32_t bounces, uint32_t flags,
                     uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out);

/* Image-space partition for multi-GPU: traces only the row blocks b (of block_rows scanlines) with
 * b % part_count == part_index, writing them PACKED (block after block, each block_rows*width pixels;
 * the last block of the image may be short and is zero-padded) so that equal-sized rank buffers can
 * be gathered. Output capacity: rt_rows_packed_pixels(). Ray generation still uses the full
 * width x height launch size. */
RT_API int  rt_trace_rows(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam,
                          uint32_t width, uint32_t height, uint32_t bounces, uint32_t flags,
                          uint32_t block_rows, uint32_t part_index, uint32_t part_count,
                          uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out);
/* The same for only the packed rows [first_row, first_row + n_rows) of this part (device output; first_row and n_rows multiples of
 * block_rows, i.e. whole bands, unless the range ends the part): lets a caller pipeline a frame in row chunks, e.g. rank 0 copying
 * the finished rows of all ranks to the host while the next chunk is traced.
 * With interleaved bands, packed rows [a, b) of every part together are then the image rows [a * part_count, b * part_count). */
RT_API int  rt_trace_rows_range(rt_context* ctx, const rt_tlas* tlas, const rt_camera* cam,
                                uint32_t width, uint32_t height, uint32_t bounces, uint32_t flags,
                                uint32_t block_rows, uint32_t part_index, uint32_t part_count, uint32_t first_row, uint32_t n_rows,
                                uint8_t* rgba_out, rt_hit* primary_hits_out, rt_hit* secondary_hits_out);
/* Completes the host-output trace that was enqueued with RT_TRACE_ASYNC (no-op when none is in flight). */
RT_API int  rt_host_frame_wait(rt_context* ctx);
RT_API uint64_t rt_rows_packed_pixels(uint32_t width, uint32_t height, uint32_t block_rows, uint32_t part_count);
/* Rank 0 after the gather: scatter part_count packed buffers (contiguous, each
 * rt_rows_packed_pixels()*4 bytes, device memory) into the final width*height*4 framebuffer (device). */
RT_API int  rt_unpack_rows(rt_context* ctx, const uint8_t* packed_all, uint32_t width, uint32_t height,
                           uint32_t block_rows, uint32_t part_count, uint8_t* rgba_out_device);

/* A framebuffer that other processes (one per GPU) can map: cudaMalloc + cudaIpcGetMemHandle on the owner, cudaIpcOpenMemHandle on
 * the peers (the 64-byte handle travels through any host channel, e.g. a torch.distributed broadcast). Peers pass the mapped pointer
 * to rt_trace_rows(..., RT_TRACE_OUT_DEVICE | RT_TRACE_OUT_FULL_FRAME, ...). Completion is the caller's job (a barrier after the trace). */
RT_API int  rt_frame_share_create(rt_context* ctx, uint64_t bytes, void** device_ptr_out, uint8_t handle_out[64]);
RT_API int  rt_frame_share_open(rt_context* ctx, const uint8_t handle[64], void** device_ptr_out);
RT_API int  rt_frame_share_close(rt_context* ctx, void* mapped_device_ptr);     /* a pointer from rt_frame_share_open */
/* Stream-ordered 32-bit flags in (peer-)device memory, e.g. in the tail of a shared frame: the frame-complete / buffer-free
 * handshake between the ranks without a collective. rt_flag_add enqueues "counter += 1" (system scope, after everything enqueued
 * before it on the context stream, including stores to peer memory); rt_flag_wait_ge makes the context stream wait until
 * *counter >= target (a one-thread kernel polling with back-off; after ~4 s it gives up, sets the context's error state —
 * reported by the next synchronising call as RT_ERROR_INTERNAL — and lets the stream continue rather than hanging the GPU). */
RT_API int  rt_flag_add(rt_context* ctx, uint32_t* counter_device_ptr);
RT_API int  rt_flag_wait_ge(rt_context* ctx, const uint32_t* counter_device_ptr, uint32_t target);
RT_API int  rt_frame_share_free(rt_context* ctx, void* device_ptr);             /* a pointer from rt_frame_share_create */

/* ---- render group: the 8 GPUs of one box, one process per GPU (SURVEY 8(b) "rt_create_group", 8(e)) ------------------------------
 * The reference is single-GPU (one device, one queue: main.cpp:294-388); a group makes `world` processes — each with its own
 * rt_context on its own GPU and the scene replicated — render ONE frame: the image is cut into 8-scanline bands, band b belongs to
 * rank b % world. No MPI / NCCL / torch is needed: the ranks meet in a POSIX shared-memory object named after `name`
 * (every rank passes the same name, e.g. derived from the job id; rank 0 creates it, the others attach; 60-s timeouts everywhere).
 *   RT_GROUP_OUT_DEVICE: every rank's trace kernel stores its pixels straight into rank 0's device frame over NVLink / NVSwitch
 *     (CUDA-IPC mapping, two alternating frames, stream-ordered completion counters in the frame's tail: no collective, no host
 *     round trip). On rank 0 *frame_out is the device pointer of the finished frame, valid until the second-next group trace.
 *   RT_GROUP_OUT_HOST: every rank copies its bands over its OWN PCIe link into the group's pinned host frame (shared memory that
 *     all ranks registered with CUDA); on rank 0 *frame_out is the host pointer of the finished frame (valid until the second-next
 *     group trace). The call returns on rank 0 when the whole frame is there, on the other ranks when their share is.
 *   RT_GROUP_ASYNC (device output): enqueue and return; rt_group_sync() waits for everything enqueued so far.
 *   RT_GROUP_PIPELINE (device output, implies ASYNC): consecutive frames alternate over two internal streams, so frame k + 1's
 *     first rays fill the tail of frame k's persistent kernels (the per-frame fixed cost that limits scaling at 8 GPUs).
 *   RT_GROUP_PIPELINE with RT_GROUP_OUT_HOST: two frames in flight. Call k enqueues frame k (trace + this rank's PCIe copies, on
 *     alternating internal contexts) and then COMPLETES frame k - 1: on rank 0 *frame_out is the finished frame k - 1 (NULL for the
 *     first call), valid until the next group trace. rt_group_flush_host() completes the last frame. Frame k's copies and the
 *     ranks' handshake thus overlap frame k + 1's trace.
 * Every rank must make the same sequence of group calls. A group of world = 1 is valid (no peer traffic). */
typedef struct rt_group rt_group;
/* Completes the pending frame of a pipelined host-output sequence (no-op without one); on rank 0 *frame_out (may be NULL) is that frame. */
RT_API int  rt_group_flush_host(rt_group* group, const uint8_t** frame_out);
#define RT_GROUP_OUT_DEVICE 0x1u
#define RT_GROUP_OUT_HOST   0x2u
#define RT_GROUP_ASYNC      0x4u
#define RT_GROUP_PIPELINE   0x8u
RT_API int  rt_group_create(rt_context* ctx, const char* name, int rank, int world, uint32_t max_width, uint32_t max_height, rt_group** out);
RT_API void rt_group_destroy(rt_group* group);
RT_API int  rt_group_trace(rt_group* group, const rt_tlas* tlas, const rt_camera* cam, uint32_t width, uint32_t height, uint32_t bounces,
                           uint32_t flags, const uint8_t** frame_out);
RT_API int  rt_group_sync(rt_group* group);       /* all frames enqueued by this rank are complete (and, on rank 0, assembled) */
RT_API int  rt_group_join(rt_group* group);       /* stream-ordered: the context stream waits for the group's internal streams (no host wait) */
RT_API int  rt_group_barrier(rt_group* group);    /* host barrier of all ranks */
/* The host-frame handshake on its own (rt_group_trace with RT_GROUP_OUT_HOST = begin + rt_trace_rows into the frame + end): begin gives
 * every rank the shared pinned host frame of this turn once rank 0's caller has let go of its previous content; end publishes this rank's
 * share and, on rank 0, returns when all `world` shares are there. Works without CUDA too (ctx == NULL at rt_group_create). */
RT_API int  rt_group_host_frame_begin(rt_group* group, uint8_t** frame_out);
RT_API int  rt_group_host_frame_end(rt_group* group, const uint8_t** frame_out_rank0);
RT_API const char* rt_group_last_error(const rt_group* group);
RT_API int  rt_group_rank(const rt_group* group);
RT_API int  rt_group_world(const rt_group* group);
/* cfg5: "BLASes are built per GPU and broadcast". Collective over the group: rank `owner_rank` passes its BLAS (built on
 * its own, rt_build_blas), the others NULL; on return *out is that BLAS on the owner and, on every other rank, a handle to a local copy
 * that is being PULLED from the owner's memory over NVLink on a copy stream (CUDA-IPC mapping + cudaMemcpyAsync: the copy overlaps
 * whatever the rank builds next). `slot` < 64 names the exchange; rt_group_share_finish() waits for this rank's pulls and is a barrier:
 * after it every handle is usable and owners may free or update their BLASes. */
RT_API int  rt_group_share_blas(rt_group* group, uint32_t slot, int owner_rank, const rt_blas* mine, rt_blas** out);
RT_API int  rt_group_share_finish(rt_group* group);
RT_API float rt_group_last_share_ms(const rt_group* group);   /* device time of this rank's pulls between the first share and share_finish */
/* Host milliseconds of the last finished exchange on this rank: {waiting for owners to publish, cudaIpcOpenMemHandle, cudaMalloc of the copies}. */
RT_API int  rt_group_last_share_host_ms(const rt_group* group, float out[3]);

/* Copies device memory of this context's GPU (e.g. the frame rt_group_trace returned with RT_GROUP_OUT_DEVICE) to host memory, ordered
 * after the work enqueued on the context stream; returns when the bytes are there. For callers that do not link the CUDA runtime. */
RT_API int  rt_copy_to_host(rt_context* ctx, void* dst_host, const void* src_device, uint64_t bytes);

RT_API int  rt_last_trace_stats(const rt_context* ctx, rt_trace_stats* out);
/* CUDA-event milliseconds of the kernels of the most recent rt_trace / rt_trace_rows / rt_unpack_rows call (no copies). */
RT_API float rt_last_trace_ms(const rt_context* ctx);
/* Number of kernels this library has launched on the context since creation. */
RT_API uint64_t rt_kernel_launch_count(const rt_context* ctx);

RT_API const char* rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTCORE_H_ */
