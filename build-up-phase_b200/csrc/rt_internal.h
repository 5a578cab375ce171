// rt_internal.h — device data layout + internal host interfaces of librtcore (sm_100a).
//
// HBM layout (all little-endian, 16-byte aligned so every fetch is an LDG.128):
//   BvhNode     64 B  = two 32-B halves {lo.xyz, hi.xyz, ref, height}, one per child. A half is
//                       exactly one 32-B DRAM sector and is written by the ONE thread that climbs
//                       through that child during the atomic bottom-up refit.
//   TriRec      48 B  = v0.xyz v1.xyz v2.xyz geometryIndex primitiveIndex geometryFlags  (3 x LDG.128);
//                       BLAS-space vertices with the per-geometry transform already baked in
//                       (vkCmdBuildAccelerationStructuresKHR semantics, reference main.cpp:795-808).
//   InstanceRec 96 B  = world->object 3x4, BLAS node/triangle base pointers, root ref, packed
//                       customIndex|mask and sbtOffset|flags (VkAccelerationStructureInstanceKHR,
//                       reference main.cpp:848-858), TLAS slot id, |BLAS bounds| for the slab pad.
// Child refs: >= 0 internal node index (relative to the AS's node base); < 0 leaf = ~((first<<3)|(count-1))
// into the AS's primitive array; large positive values are traversal sentinels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtcore.h"

namespace rt {

constexpr int32_t REF_DONE  = 0x7FFFFFFF;
constexpr int32_t REF_POP_INSTANCE = 0x7FFFFFFE;
constexpr int32_t REF_EMPTY = RT_REF_EMPTY;      // 0x7FFFFFFD
constexpr int32_t REF_SENTINEL_MIN = 0x7FFFFFF0;
#ifndef RT_BLAS_LEAF_MAX
#define RT_BLAS_LEAF_MAX 2
#endif
constexpr int BLAS_LEAF_MAX = RT_BLAS_LEAF_MAX;   // <= 8 (leaf refs pack count - 1 into 3 bits); the oracle's LEAF_MAX must match for the bit-for-bit build test
constexpr int TLAS_LEAF_MAX = 1;
constexpr uint32_t MORTON_BITS = 30;
constexpr uint32_t MAX_PRIMS = 1u << 28;         // leaf ref packs (first << 3) into 31 bits

struct __align__(32) BvhNodeHalf { float lo[3]; float hi[3]; int32_t ref; uint32_t height; };
struct __align__(64) BvhNode { BvhNodeHalf c[2]; };
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 B");

struct __align__(16) TriRec { float v[9]; uint32_t geo, prim, flags; };   // flags: RT_GEOMETRY_* of its geometry (during the build: flags << 24 | BLAS id)
static_assert(sizeof(TriRec) == 48, "TriRec must be 48 B");

struct __align__(16) InstanceRec {
    float w2o[12];              // 48
    const BvhNode* nodes;       // 56
    const TriRec*  tris;        // 64
    int32_t  root;              // 68
    uint32_t custom_mask;       // 72   custom_index:24 | mask:8
    uint32_t sbt_flags;         // 76   sbt_offset:24 | flags:8
    uint32_t instance_id;       // 80   slot in the caller's rt_instance array (gl_InstanceID)
    float    absmax[3];         // 92   max(|blas.lo|, |blas.hi|) per axis
    uint32_t active;            // 96   bit 0: traversable; bits 1..31: (geometry count of the BLAS) - 1
};
static_assert(sizeof(InstanceRec) == 96, "InstanceRec must be 96 B");

// What accelerationStructureReference points at: one record per BLAS in device memory.
struct __align__(16) BlasRecord {
    const BvhNode* nodes;
    const TriRec*  tris;
    int32_t  root;
    uint32_t height;
    float    lo[3], hi[3];
    uint32_t tri_count;
    uint32_t n_geoms;
    uint32_t first;             // first primitive of this BLAS inside a batched build
    uint32_t node_slots;        // node slots of this BLAS behind `nodes` (tri_count after a build, its live-node count after compaction)
};
static_assert(sizeof(BlasRecord) == 64, "BlasRecord must be 64 B");

// One geometry of a (possibly batched) BLAS build, device-resident inputs.
struct GeomDesc {
    const float*    verts;
    const uint32_t* idx;        // may be null
    uint32_t stride_f;          // vertex stride in floats
    uint32_t tri_first;         // global index of this geometry's first triangle
    uint32_t tri_count;
    uint32_t blas;              // which BLAS of the batch
    uint32_t geo_index;         // gl_GeometryIndexEXT inside that BLAS
    uint32_t has_xform;
    uint32_t flags;             // RT_GEOMETRY_* (low 8 bits are kept with every triangle)
    float    xform[12];
    uint32_t vert_count;        // rt_geometry.vertex_count (how much of `verts` the indices may address)
};
static_assert(sizeof(GeomDesc) == 96, "GeomDesc layout");

// ---- radix sort (csrc/radix_sort.cu) ---------------------------------------------------------
struct BlasRecord;
struct SortPlan {
    uint32_t n = 0;
    int passes = 0;             // 8-bit digits
    uint32_t tiles = 0;
    size_t scratch_bytes = 0;   // histogram + look-back state
    int packed_val_bits = 0;    // > 0: records are 64-bit words `key << packed_val_bits | value` (vals arrays unused)
    // segmented path (batches of small BLASes, packed records): the input is already grouped by segment, every segment
    // fits one CTA's shared memory and is sorted there by its low seg_key_bits key bits in one kernel (no global passes)
    const struct BlasRecord* seg_records = nullptr; uint32_t n_segments = 0; int seg_key_bits = 0;
    bool seg_fused = false;     // also fuse triangle setup + Morton into the segment's CTA (k_seg_setup_sort)
    bool seg_single = false;    // no records: the whole input (n <= SEG_SORT_CAPACITY) is one segment (the TLAS build)
};
SortPlan sort_plan(uint32_t n, int key_bits);
constexpr uint32_t SEG_SORT_CAPACITY = 11264;   // records one CTA sorts in shared memory (1024 threads x 11)
// Sorts (keys, vals) by the low key_bits of the key, stable. vals_a == nullptr selects the packed format. Result ends in keys_a/vals_a or keys_b/vals_b;
// *result_in_b tells which. Returns the number of kernels launched, or <0 on a launch error.
int sort_pairs(const SortPlan& plan, uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b,
               void* scratch, int* device_error_flag, cudaStream_t stream, bool* result_in_b);

// ---- LBVH build (csrc/lbvh_build.cu) -----------------------------------------------------------
struct BuildScratch {           // all device pointers, sized for n primitives
    uint64_t* keys_a; uint64_t* keys_b;
    uint32_t* vals_a; uint32_t* vals_b;
    uint32_t* far_end;          // 2n: far end of the range of the child deposited at (split, side), border subtrees only
    uint32_t* arrived;          // n + 1: arrival counters of the splits resolved through global memory, then the border-job count
    float4*   jobs;             // 3 x float4 per border job, tree_job_capacity(n) of them
    float4*   xchg;             // 4n: the two deposited child halves of the splits resolved through global memory (sparsely touched)
    void*     sort_scratch;
    int*      error_flag;       // 1 int
};

struct BlasBuildArgs {
    const GeomDesc* geoms; uint32_t n_geoms;
    const uint32_t* geom_tri_first;   // n_geoms+1 prefix array (device) for the binary search
    uint32_t n_tris; uint32_t n_blas;
    uint32_t seg_bits;                // ceil(log2(n_blas))
    TriRec*   tris_unsorted;          // scratch, n_tris
    TriRec*   tris_sorted;            // output
    BvhNode*  nodes;                  // output, n_tris slots
    BlasRecord* records;              // n_blas (device), nodes/tris/first/tri_count/n_geoms pre-filled by the host
    int*      bounds_ordered;         // n_blas * 6 ints (ordered-int encoded floats), pre-initialised
    BuildScratch s;
    SortPlan  sort;
    // refit-only update: the sorted records of the last full build (device; packed when sort.packed_val_bits > 0). The setup kernel
    // bakes the new vertices, Morton + sort are skipped and the tree pass runs over these records: same topology, new boxes.
    const uint64_t* reuse_keys = nullptr; const uint32_t* reuse_vals = nullptr;
};
struct BuildEvents { cudaEvent_t e[6]; };  // setup | morton | sort | hierarchy | refit | end
uint32_t tree_job_capacity(uint32_t n);   // upper bound of the border-job queue length for n leaves
int launch_blas_build(const BlasBuildArgs& a, cudaStream_t st, const BuildEvents* ev, bool* sorted_in_b);
// Compaction (rt_compact_blas). Pass 1: live flags of the n Karras slots from the sorted records (a slot is dead iff its node's leaf
// range holds <= BLAS_LEAF_MAX leaves), exclusive prefix -> cidx[n + 1] (cidx[n] = live total). Pass 2: live node i -> dst[cidx[i]] with
// its internal child refs re-based. scratch: (n + 1) uint32 for cidx + compact_scratch_bytes(n).
size_t compact_scratch_bytes(uint32_t n);
int launch_compact_index(const uint64_t* keys, int vb, uint32_t n, const BlasRecord* records, uint32_t* cidx, void* scratch, cudaStream_t st);
int launch_compact_nodes(const BvhNode* src, BvhNode* dst, const uint32_t* cidx, const uint64_t* keys, int vb, uint32_t n,
                         const BlasRecord* records, cudaStream_t st);
int launch_compact_records(const BlasRecord* src, BlasRecord* dst, uint32_t n_blas, const uint32_t* cidx, const BvhNode* nodes, const TriRec* tris, cudaStream_t st);

struct TlasBuildArgs {
    const rt_instance* instances;     // device copy of the caller's 64-byte records (blas field = BlasRecord device address)
    uint32_t n;
    InstanceRec* inst_unsorted;       // scratch
    InstanceRec* inst_sorted;         // output
    float*       boxes_unsorted;      // scratch n*6
    BvhNode*     nodes;               // output
    int*         bounds_ordered;      // 6 ints
    int32_t*     root_out;            // {root, height, max_sbt_plus_geo, max_sbt, max_geo, max_blas_height}  (6 ints, device)
    float*       bounds_out;          // 6 floats (device)
    BuildScratch s;
    SortPlan     sort;
};
int launch_tlas_build(const TlasBuildArgs& a, cudaStream_t st);

// ---- trace (csrc/trace.cu) -----------------------------------------------------------------------
struct TraceParams {
    const BvhNode* tlas_nodes;
    uint32_t tlas_smem_nodes;   // TLAS node slots 0..this-1 hold every node of the TLAS (launch_trace stages them in shared memory when they fit; 0 = do not)
    const InstanceRec* instances;
    int32_t tlas_root;
    float tlas_absmax[3];
    float cam_pos[3];
    float aspect_x, aspect_y;
    uint32_t width, height;
    uint32_t block_rows, part_index, part_count;   // image-space partition (1 part = whole image)
    uint32_t local_rows;                           // rows this launch covers (packed)
    uint32_t row0;                                 // first (packed) row of this launch: a frame may be traced in row chunks so that the
                                                   // device->host copy of one chunk overlaps the trace of the next
    float tmin, tmax;
    uint32_t cull_mask, sbt_offset, sbt_stride, bounce_seed;
    uint32_t ray_flags;         // RT_RAY_FLAG_*; anything beyond OPAQUE/NO_OPAQUE selects the GENERAL kernel variant
    uint32_t bounces;
    const float* hit_records; uint32_t n_records;
    const uint4* anyhit; uint32_t n_anyhit;   // any-hit records {kind, log2_res, flags, first mask word (index into this array, in 32-bit words)}, GENERAL variant only
    float miss[3];
    uint8_t* rgba;              // packed local_rows x width x 4, or (full_frame) the whole height x width x 4 image, possibly peer memory
    uint32_t bgra;              // 1: store B,G,R,A byte order (the sample's swapchain format) instead of R,G,B,A
    uint32_t full_frame;        // 1: pixels are stored at their final position y * width + x (multi-GPU: straight into rank 0's frame over NVLink)
    rt_hit* primary_hits;       // may be null
    rt_hit* secondary_hits;     // may be null
    unsigned long long* stats;  // 8 counters (rt_trace_stats order), may be null
    uint32_t* counters;         // [0] primary ray fetch counter, [1] secondary fetch counter, [2] bounce-queue length (zeroed per launch)
    float4* queue;              // bounce rays: 3 x float4 per secondary ray {pixel, o.xyz | d.xyz, r | g, b, -, -}, one slot per tile-major ray index
    uint32_t* tile_mask;        // one word per 8x4-pixel tile: which of its pixels spawned a bounce ray; followed by the block sums of the index build
    uint32_t* bounce_index;     // compact tile-ordered list of the occupied queue slots (stage 1 reads rays through it)
    // fused launch (both stages in one persistent kernel): the queue is an append queue again, entry i is published by
    // queue_flags[i] = epoch (a per-launch number, so the flags never need clearing); counters[3] counts finished primary rays
    uint32_t* queue_flags; uint32_t epoch; uint32_t queue_capacity;
    int* error_flag;            // device int: set to 2 by the fused kernel's watchdog (a claim that is never published)
    // region-major ray numbering (RT_REGIONS >= 1): a region = RT_REGION_TW x RT_REGION_TH tiles; ray ids are padded to whole regions
    uint32_t regions_x, n_regions;
    uint32_t n_sm;              // RT_REGIONS == 2: SMs of the device (home region of a warp = smid * n_regions / n_sm)
    uint32_t* region_next;      // RT_REGIONS == 2: 2 * n_regions fetch counters (stage 0 | stage 1), zeroed per launch together with `counters`
};
// RT_REGIONS: 0 = rays numbered tile-major along image rows; 1 = region-major numbering (the rays in flight at any time cover a compact
// 2-D patch of the image, i.e. a handful of instances, instead of a full-width strip), one global fetch counter; 2 = region-major numbering
// AND one fetch counter per region with every SM starting in its own home regions (all warps of an SM work on the same patch: L1 reuse).
// Measured on B200, inst10m 4K + bounce (profiles/README.md r2_f): 0 / 1 / 2 -> 3518 / 3511 / 3355 Mrays/s (regions of 8x8, 8x16, 32x16 tiles:
// 3169 / 3259 / 3393): the locality of the row-major tile order is already enough for L2, and SM-affine regions cost more at the
// refill and in the kernel tail than they return in L1 hits. Default 0; the other modes stay as build options.
#ifndef RT_REGIONS
#define RT_REGIONS 0
#endif
#ifndef RT_REGION_TW
#define RT_REGION_TW 16
#endif
#ifndef RT_REGION_TH
#define RT_REGION_TH 16
#endif
constexpr uint32_t TRACE_REGION_TILES = RT_REGIONS ? RT_REGION_TW * RT_REGION_TH : 1u;
// ray slots (= tile-major ray ids, whole tiles, padded to whole regions) of a launch of `rows` packed rows
inline uint32_t trace_regions_x(uint32_t width) { return RT_REGIONS ? (((width + 7u) >> 3) + RT_REGION_TW - 1u) / RT_REGION_TW : ((width + 7u) >> 3); }
inline uint32_t trace_regions_y(uint32_t rows) { return RT_REGIONS ? (((rows + 3u) >> 2) + RT_REGION_TH - 1u) / RT_REGION_TH : ((rows + 3u) >> 2); }
inline uint64_t trace_tiles_padded(uint32_t width, uint32_t rows) { return (uint64_t)trace_regions_x(width) * trace_regions_y(rows) * TRACE_REGION_TILES; }
constexpr size_t TRACE_QUEUE_ENTRY_BYTES = 48;
// per ray slot: the ray record + its index entry + its share of the tile mask / block sums (rounded up)
constexpr size_t TRACE_BOUNCE_AUX_BYTES_PER_SLOT = 4 + 1;
int launch_trace(const TraceParams& p, bool stats, int stack_needed, int sm_count, cudaStream_t st);
// max over the instances of sbt_offset + (n_geoms - 1) * stride: the highest hit record a trace with this stride can address (before
// sbtRecordOffset); atomicMax into *out (device, pre-zeroed)
int launch_sbt_bound(const InstanceRec* inst, uint32_t n, uint32_t stride, unsigned long long* out, cudaStream_t st);
int launch_flag_add(uint32_t* counter, cudaStream_t st);
int launch_flag_wait_ge(const uint32_t* counter, uint32_t target, int* error_flag, cudaStream_t st);
int launch_unpack_rows(const uint8_t* packed_all, uint32_t width, uint32_t height, uint32_t block_rows,
                       uint32_t part_count, uint8_t* out, cudaStream_t st);

}  // namespace rt
