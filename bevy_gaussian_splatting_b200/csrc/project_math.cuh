// project_math.cuh -- the bit-exact part of the per-gaussian maths (key-gen + projection).
//
// Translation units including this header are compiled with -fmad=false: every a*b+c is a
// rounded multiply followed by a rounded add, in the order written, exactly as the CPU oracle
// evaluates it (-ffp-contract=off).  Division and sqrt are IEEE (nvcc defaults
// -prec-div=true -prec-sqrt=true); no fast-math.  This is what makes sort keys, projected
// records, pixel bounding boxes and therefore tile ranges bit-identical to the oracle.
#pragma once
#include "common.cuh"

namespace bgs {

// M * (x,y,z,1), summed ((m0*x + m1*y) + m2*z) + m3.  Reference: WGSL mat4x4 * vec4
// (radix.wgsl:88, transform.wgsl:6, helpers.wgsl:18).
__device__ __forceinline__ void mat4_point(const float* m, float x, float y, float z, float out[4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = ((m[0 + r] * x + m[4 + r] * y) + m[8 + r] * z) + m[12 + r];
}
__device__ __forceinline__ void mat4_dir(const float* m, float x, float y, float z, float out[4]) {
#pragma unroll
    for (int r = 0; r < 4; ++r) out[r] = (m[0 + r] * x + m[4 + r] * y) + m[8 + r] * z;
}

struct KeyOut {
    uint32_t key;
    bool visible;
    float pw[3];
    float ndc[2];
    float d2;
};

// radix.wgsl:86-101 (key) + transform.wgsl:5-14 (world_to_clip, in_frustum).
__device__ __forceinline__ KeyOut key_of(const FrameConsts& c, float px, float py, float pz) {
    KeyOut k;
    float pw[4];
    mat4_point(c.model, px, py, pz, pw);
    k.pw[0] = pw[0]; k.pw[1] = pw[1]; k.pw[2] = pw[2];
    float cl[4];
    mat4_point(c.clip_from_world, pw[0], pw[1], pw[2], cl);
    const float den = cl[3] + 0.000000001f;
    const float nx = cl[0] / den, ny = cl[1] / den, nz = cl[2] / den;
    k.ndc[0] = nx; k.ndc[1] = ny;
    k.visible = fabsf(nx) < 1.1f && fabsf(ny) < 1.1f && fabsf(nz - 0.5f) < 0.5f;
    const float dx = pw[0] - c.cam[0], dy = pw[1] - c.cam[1], dz = pw[2] - c.cam[2];
    k.d2 = (dx * dx + dy * dy) + dz * dz;
    uint32_t key = 0xFFFFFFFFu;
    if (k.visible) key = 0xFFFFFFFFu - __float_as_uint(k.d2);
    k.key = key >> c.key_shift;
    return k;
}

// Key-gen's version of key_of: the SAME key and the SAME visibility decision, cheaper.
//  * identity model matrix (the common case): M*(p,1) == p bit for bit when p is finite (x*1 + y*0 + z*0 + 0), so
//    the multiply is skipped for finite positions;
//  * the three IEEE divisions of world_to_clip only feed the frustum test here: one approximate reciprocal
//    (|error| < 4e-7 relative) decides every gaussian whose ndc is not within 1e-4 of a frustum bound; the few that
//    are fall back to the exact divisions.  Denominators outside [1e-30, 1e30] (rcp.approx flushes) also fall back.
__device__ __forceinline__ uint32_t key_of_fast(const FrameConsts& c, float px, float py, float pz, bool& visible) {
    float pw[4];
    if (c.model_identity && (fabsf(px) + fabsf(py)) + fabsf(pz) < __uint_as_float(0x7F800000u)) {
        pw[0] = px; pw[1] = py; pw[2] = pz;
    } else {
        mat4_point(c.model, px, py, pz, pw);
    }
    float cl[4];
    mat4_point(c.clip_from_world, pw[0], pw[1], pw[2], cl);
    const float den = cl[3] + 0.000000001f;
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));
    const float ax = fabsf(cl[0] * rc), ay = fabsf(cl[1] * rc), z = cl[2] * rc;
    const float ad = fabsf(den);
    const bool sure_in = ax < 1.0999f && ay < 1.0999f && z > 1e-6f && z < 0.9999f;
    const bool sure_out = ax > 1.1001f || ay > 1.1001f || z < -1e-6f || z > 1.0001f;
    const bool den_ok = ad > 1e-30f && ad < 1e30f;
    bool vis;
    if (den_ok && (sure_in || sure_out)) {
        vis = sure_in;
    } else {
        const float nx = cl[0] / den, ny = cl[1] / den, nz = cl[2] / den;
        vis = fabsf(nx) < 1.1f && fabsf(ny) < 1.1f && fabsf(nz - 0.5f) < 0.5f;
    }
    visible = vis;
    const float dx = pw[0] - c.cam[0], dy = pw[1] - c.cam[1], dz = pw[2] - c.cam[2];
    const float d2 = (dx * dx + dy * dy) + dz * dz;
    uint32_t key = 0xFFFFFFFFu;
    if (vis) key = 0xFFFFFFFFu - __float_as_uint(d2);
    return key >> c.key_shift;
}

// The visibility decision of key_of_fast alone (key-gen phase 1: one bit per gaussian, no key).
__device__ __forceinline__ bool visible_fast(const FrameConsts& c, float px, float py, float pz) {
    float pw[4];
    if (c.model_identity && (fabsf(px) + fabsf(py)) + fabsf(pz) < __uint_as_float(0x7F800000u)) {
        pw[0] = px; pw[1] = py; pw[2] = pz;
    } else {
        mat4_point(c.model, px, py, pz, pw);
    }
    float cl[4];
    mat4_point(c.clip_from_world, pw[0], pw[1], pw[2], cl);
    const float den = cl[3] + 0.000000001f;
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(den));
    const float ax = fabsf(cl[0] * rc), ay = fabsf(cl[1] * rc), z = cl[2] * rc;
    const float ad = fabsf(den);
    const bool sure_in = ax < 1.0999f && ay < 1.0999f && z > 1e-6f && z < 0.9999f;
    const bool sure_out = ax > 1.1001f || ay > 1.1001f || z < -1e-6f || z > 1.0001f;
    const bool den_ok = ad > 1e-30f && ad < 1e30f;
    if (den_ok && (sure_in || sure_out)) return sure_in;
    const float nx = cl[0] / den, ny = cl[1] / den, nz = cl[2] / den;
    return fabsf(nx) < 1.1f && fabsf(ny) < 1.1f && fabsf(nz - 0.5f) < 0.5f;
}

// The key of a gaussian already known to be visible (key-gen phase 2): same pw, same d2 as key_of.
__device__ __forceinline__ uint32_t key_only(const FrameConsts& c, float px, float py, float pz) {
    float pw[4];
    if (c.model_identity && (fabsf(px) + fabsf(py)) + fabsf(pz) < __uint_as_float(0x7F800000u)) {
        pw[0] = px; pw[1] = py; pw[2] = pz;
    } else {
        mat4_point(c.model, px, py, pz, pw);
    }
    const float dx = pw[0] - c.cam[0], dy = pw[1] - c.cam[1], dz = pw[2] - c.cam[2];
    const float d2 = (dx * dx + dy * dy) + dz * dz;
    return (0xFFFFFFFFu - __float_as_uint(d2)) >> c.key_shift;
}

// Fixed-series natural log in f64 (the policy replacement for WGSL log(), gaussian.wgsl:229):
// x = m 2^e, m in [sqrt(1/2), sqrt 2); s = (m-1)/(m+1); ln x = e ln2 + 2 s P(s^2), rounded to f32.
__device__ __forceinline__ float det_ln(float xf) {
    if (xf != xf) return xf;
    if (xf < 0.0f) return __uint_as_float(0x7FC00000u);
    if (xf == 0.0f) return __uint_as_float(0xFF800000u);
    if (xf == __uint_as_float(0x7F800000u)) return xf;
    const double x = (double)xf;
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7FFull) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m = __longlong_as_double((long long)bits);
    if (m > 1.4142135623730951) { m = __dmul_rn(m, 0.5); e += 1; }
    const double s = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
    const double z = __dmul_rn(s, s);
    double p = 1.0 / 23.0;
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 21.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 19.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 17.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 15.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 13.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 11.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 9.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 7.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 5.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0 / 3.0);
    p = __dadd_rn(__dmul_rn(p, z), 1.0);
    const double r = __dadd_rn(__dmul_rn((double)e, 0.6931471805599453), __dmul_rn(2.0, __dmul_rn(s, p)));
    return (float)r;   // cvt.rn.f32.f64
}

}  // namespace bgs
