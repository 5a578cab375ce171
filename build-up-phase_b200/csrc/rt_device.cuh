// rt_device.cuh — device-side arithmetic shared by the build and trace kernels.
//
// Compiled with -fmad=false: every fp32 expression is evaluated exactly as written (IEEE-754
// binary32 RNE, IEEE division/sqrt), FMAs only where __fmaf_rn is spelled out (the conservative
// box test, which does not need to be bit-reproducible — only conservative). The result-defining
// arithmetic (vertex transform, ray generation, instance transform, watertight triangle test,
// shading, UNORM8 store) is therefore bit-identical to a strict-IEEE scalar CPU evaluation.
#pragma once
#include "rt_internal.h"

namespace rt {

struct V3 { float x, y, z; };

__device__ __forceinline__ float comp3(const V3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// row-major 3x4 (VkTransformMatrixKHR)
__device__ __forceinline__ V3 xform_point(const float* m, V3 p) {
    return {((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3],
            ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
            ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]};
}
__device__ __forceinline__ V3 xform_vec(const float* m, V3 d) {
    return {(m[0] * d.x + m[1] * d.y) + m[2] * d.z,
            (m[4] * d.x + m[5] * d.y) + m[6] * d.z,
            (m[8] * d.x + m[9] * d.y) + m[10] * d.z};
}
__device__ __forceinline__ V3 xform_normal(const float* w2o, V3 n) {   // (w2o)^T * n
    return {(w2o[0] * n.x + w2o[4] * n.y) + w2o[8] * n.z,
            (w2o[1] * n.x + w2o[5] * n.y) + w2o[9] * n.z,
            (w2o[2] * n.x + w2o[6] * n.y) + w2o[10] * n.z};
}

// inverse of a 3x4 affine in fp64, rounded once to fp32; returns false when singular
__device__ inline bool invert3x4(const float* m, float* out) {
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double tx = m[3], ty = m[7], tz = m[11];
    double c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
    double c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
    double c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
    double det = (a * c00 + b * c10) + c * c20;
    if (!(det != 0.0) || isinf(det) || isnan(det)) { for (int k = 0; k < 12; ++k) out[k] = 0.0f; return false; }
    double inv = 1.0 / det;
    double r00 = c00 * inv, r01 = c01 * inv, r02 = c02 * inv;
    double r10 = c10 * inv, r11 = c11 * inv, r12 = c12 * inv;
    double r20 = c20 * inv, r21 = c21 * inv, r22 = c22 * inv;
    double r03 = -((r00 * tx + r01 * ty) + r02 * tz);
    double r13 = -((r10 * tx + r11 * ty) + r12 * tz);
    double r23 = -((r20 * tx + r21 * ty) + r22 * tz);
    out[0] = (float)r00; out[1] = (float)r01; out[2]  = (float)r02; out[3]  = (float)r03;
    out[4] = (float)r10; out[5] = (float)r11; out[6]  = (float)r12; out[7]  = (float)r13;
    out[8] = (float)r20; out[9] = (float)r21; out[10] = (float)r22; out[11] = (float)r23;
    return true;
}

__host__ __device__ __forceinline__ uint32_t pcg_hash(uint32_t v) {
    uint32_t state = v * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

// ---- ordered-int encoding of floats for atomicMin/atomicMax ----------------------------------
__host__ __device__ __forceinline__ int float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    int i = __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__host__ __device__ __forceinline__ float ordered_to_float(int i) {
    int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
#ifdef __CUDA_ARCH__
    return __int_as_float(j);
#else
    float f; memcpy(&f, &j, 4); return f;
#endif
}

// ---- Morton ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t quant10(float c, float lo, float hi) {
    float ext = hi - lo;
    float inv = ext > 0.0f ? 1024.0f / ext : 0.0f;
    float q = (c - lo) * inv;
    q = fminf(q, 1023.0f);
    if (!(q >= 0.0f)) q = 0.0f;
    return (uint32_t)q;
}
// 30-bit Morton code of a box centre inside scene box [slo,shi]
__device__ __forceinline__ uint32_t morton30_centre(float cx, float cy, float cz, const float* slo, const float* shi) {
    uint32_t x = quant10(cx, slo[0], shi[0]);
    uint32_t y = quant10(cy, slo[1], shi[1]);
    uint32_t z = quant10(cz, slo[2], shi[2]);
    return (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
}
// 30-bit Morton code of the centre of primitive box [plo,phi] inside scene box [slo,shi]
__device__ __forceinline__ uint32_t morton30(const float* plo, const float* phi, const float* slo, const float* shi) {
    float cx = (plo[0] + phi[0]) * 0.5f, cy = (plo[1] + phi[1]) * 0.5f, cz = (plo[2] + phi[2]) * 0.5f;
    uint32_t x = quant10(cx, slo[0], shi[0]);
    uint32_t y = quant10(cy, slo[1], shi[1]);
    uint32_t z = quant10(cz, slo[2], shi[2]);
    return (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
}

// ---- leaf refs --------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int32_t leaf_ref(uint32_t first, uint32_t count) { return ~(int32_t)((first << 3) | (count - 1u)); }
__host__ __device__ __forceinline__ uint32_t leaf_first(int32_t r) { return ((uint32_t)~r) >> 3; }
__host__ __device__ __forceinline__ uint32_t leaf_count(int32_t r) { return (((uint32_t)~r) & 7u) + 1u; }

}  // namespace rt
