/*
 * rtcore_io.h — host-side data formats either side of the ray-tracing path.
 *
 *   Wavefront .obj -> rt_geometry   the reference's own next assignment: "load an obj file -> build the acceleration
 *                                   structure" (vulkan-raytracing-basic/README.md:225-226; tinyobj is named in
 *                                   vulkan-basic-triangle/README.md:21). Gives real meshes to the same build ABI.
 *   RGBA8 framebuffer -> image file what the sample does with the traced image: vkCmdCopyImage into a
 *                                   B8G8R8A8_SRGB swapchain (main.cpp:50,1371-1375); headless, that is a file.
 *
 * Pure host code (file parsing / writing); nothing here touches the GPU or the hot path. Plain C ABI.
 */
#ifndef RTCORE_IO_H_
#define RTCORE_IO_H_

#include "rtcore.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RT_ERROR_IO     = -16,   /* file could not be opened / written */
    RT_ERROR_PARSE  = -17    /* malformed input; rt_obj_last_error() names the line */
};

/* ---- Wavefront .obj ------------------------------------------------------------------------------------------
 * Supported: `v x y z [w]` (w ignored), `f` with v, v/vt, v/vt/vn, v//vn references, positive (1-based) and negative
 * (relative) indices, polygons (fan-triangulated around their first vertex: (0,i,i+1)), `o`/`g` statements (each
 * starts a new group = one geometry = one gl_GeometryIndexEXT), comments, CRLF line ends, `\` line continuation.
 * vt/vn/usemtl/mtllib/s/l/p are skipped. Coordinates are parsed as IEEE doubles (strtod) and rounded once to fp32.
 * All groups share one vertex array (as in the file); a group owns a contiguous range of the index array. Groups
 * without faces are dropped; faces before the first o/g form an unnamed group "". */
typedef struct rt_obj_mesh rt_obj_mesh;

RT_API int  rt_obj_load(const char* path, rt_obj_mesh** out);
RT_API int  rt_obj_parse(const char* text, size_t length, rt_obj_mesh** out);     /* same, from memory */
RT_API void rt_obj_free(rt_obj_mesh* mesh);
RT_API const char* rt_obj_last_error(void);                                       /* thread-local, never NULL */

RT_API uint32_t rt_obj_vertex_count(const rt_obj_mesh* mesh);
RT_API uint32_t rt_obj_triangle_count(const rt_obj_mesh* mesh);                   /* all groups */
RT_API uint32_t rt_obj_group_count(const rt_obj_mesh* mesh);
RT_API const float*    rt_obj_vertices(const rt_obj_mesh* mesh);                  /* vertex_count x 3 floats */
RT_API const uint32_t* rt_obj_indices(const rt_obj_mesh* mesh);                   /* triangle_count x 3, 0-based */
RT_API const char* rt_obj_group_name(const rt_obj_mesh* mesh, uint32_t group);
RT_API uint32_t rt_obj_group_first_triangle(const rt_obj_mesh* mesh, uint32_t group);
RT_API uint32_t rt_obj_group_triangle_count(const rt_obj_mesh* mesh, uint32_t group);
/* Fills *out with the rt_geometry of one group (host pointers into the mesh, which must outlive the build call;
 * transform3x4 = NULL, flags = RT_GEOMETRY_OPAQUE). rt_build_blas(ctx, geoms, rt_obj_group_count(), ...) then
 * builds one BLAS with one geometry per group. */
RT_API int  rt_obj_geometry(const rt_obj_mesh* mesh, uint32_t group, rt_geometry* out);

/* ---- framebuffer export ------------------------------------------------------------------------------------------ */
#define RT_IMAGE_SRGB_ENCODE 0x1u   /* treat the stored values as linear and apply the sRGB OETF (what a *_SRGB swapchain does on store) */
#define RT_IMAGE_FLIP_Y      0x2u
/* Binary PPM (P6, maxval 255) of the R,G,B channels of a host RGBA8 image (alpha is dropped: the sample stores 0). */
RT_API int  rt_write_ppm(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height, uint32_t flags);
/* PNG (8-bit RGB, colour type 2, filter 0, zlib stream of STORED deflate blocks: no compression library needed, any decoder reads
 * it). Same flags as rt_write_ppm. */
RT_API int  rt_write_png(const char* path, const uint8_t* rgba, uint32_t width, uint32_t height, uint32_t flags);
/* The 8-bit sRGB encode table used by RT_IMAGE_SRGB_ENCODE: out[i] = round(255 * oetf(i / 255)). */
RT_API void rt_srgb8_table(uint8_t out[256]);

#ifdef __cplusplus
}
#endif
#endif /* RTCORE_IO_H_ */
