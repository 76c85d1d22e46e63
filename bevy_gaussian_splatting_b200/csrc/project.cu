// project.cu -- stage 3: per-gaussian 3D->2D covariance projection + SH-degree-3 colour, run
// ONCE per visible gaussian in front-to-back rank order (the reference runs it 4x per gaussian in
// its vertex stage: src/render/gaussian.wgsl:185-436).
//
// rank r (0 = nearest) -> gaussian id = sorted_ids[n_vis-1-r] (the sort is far->near like the
// reference's, src/sort/radix.wgsl) -> gather the planar attributes (f32 240 B or f16 128 B per
// gaussian) -> write one 48 B SplatRec at recs[r] (coalesced), which is all the tile stages read.
//
// Geometry (centre, OBB uv rows, pixel bbox) is bit-exact vs the oracle: compiled -fmad=false.
#include <cuda_fp16.h>

#include <cstdlib>

#include "project_math.cuh"

namespace bgs {

__constant__ float c_shc[16] = {   // src/material/spherical_harmonics.wgsl:3-20
    0.28209479177387814f, -0.4886025119029199f, 0.4886025119029199f, -0.4886025119029199f,
    1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
    0.5462742152960396f, -0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
    0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

// colour path only (compared under tolerance): one rsqrt.approx (2 ulp) instead of an IEEE sqrt + three divisions
__device__ __forceinline__ void normalize3(const float a[3], float out[3]) {
    const float inv = rsqrtf((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
    out[0] = a[0] * inv; out[1] = a[1] * inv; out[2] = a[2] * inv;
}
__device__ __forceinline__ float dot3(const float a[3], const float b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
// spherical_harmonics.wgsl:22-32; no clamp on either side
// (colour is compared under tolerance, not bit-exactly: fast pow = ex2(2.4 * lg2 x), ~1e-6 relative)
__device__ __forceinline__ float srgb_to_linear(float v) {
    if (v <= 0.04045f) return v * (1.0f / 12.92f);
    return __powf((v + 0.055f) * (1.0f / 1.055f), 2.4f);
}
__device__ __forceinline__ uint32_t pack_bbox(float lo, float hi) {
    return (uint32_t)(int)lo | ((uint32_t)(int)hi << 16);
}
constexpr uint32_t BBOX_EMPTY = 1u;   // lo = 1, hi = 0

// conservative pixel bbox of |fx - cx| <= hx, |fy - cy| <= hy (this repo's a7 rule; see oracle)
__device__ __forceinline__ void make_bbox(float cx, float cy, float hx, float hy, int Wi, int Hi,
                                          uint32_t& bx, uint32_t& by) {
    bx = BBOX_EMPTY; by = BBOX_EMPTY;
    if (!(hx >= 0.0f) || !(hy >= 0.0f) || !(cx == cx) || !(cy == cy)) return;
    const float sx = hx * 1.0e-3f + 1.0e-2f, sy = hy * 1.0e-3f + 1.0e-2f;
    float x0 = ceilf((cx - hx) - (0.5f + sx));
    float x1 = floorf((cx + hx) - (0.5f - sx));
    float y0 = ceilf((cy - hy) - (0.5f + sy));
    float y1 = floorf((cy + hy) - (0.5f - sy));
    if (!(x0 <= x1) || !(y0 <= y1)) return;
    x0 = fmaxf(x0, 0.0f); y0 = fmaxf(y0, 0.0f);
    x1 = fminf(x1, (float)(Wi - 1)); y1 = fminf(y1, (float)(Hi - 1));
    if (!(x0 <= x1) || !(y0 <= y1)) return;
    bx = pack_bbox(x0, x1); by = pack_bbox(y0, y1);
}

// Attribute fetch.  PLANAR: the reference's SoA planes as uploaded.  BLOCKED (default): a library-owned
// gaussian-major copy made once at upload -- f16: 128 B = one cache line per gaussian (pos | rot/scale/opacity |
// sh x 6), f32: 256 B (pos | rot | scale_opacity | sh x 12 | pad) -- so the random gather of a visible splat
// touches exactly its own line(s) instead of 3-4 partially used ones (the planes stay planar for key-gen).
template <bool F16, bool BLOCKED>
struct Attr;
template <>
struct Attr<false, true> {
    __device__ static float4 load(const float4*, const void* blocks, const void*, const void*, uint32_t id, float* sh,
                                  float q[4], float so[4], bool need_sh, uint32_t* = nullptr) {
        const float4* b = reinterpret_cast<const float4*>(blocks) + (size_t)id * 16;
        const float4 p = __ldg(b), r = __ldg(b + 1), s = __ldg(b + 2);
        q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
        so[0] = s.x; so[1] = s.y; so[2] = s.z; so[3] = s.w;
        if (need_sh) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float4 v = __ldg(b + 3 + i);
                sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w;
            }
        }
        return p;
    }
};
template <>
struct Attr<true, true> {
    __device__ static float lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
    __device__ static float hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
    __device__ static float4 load(const float4*, const void* blocks, const void*, const void*, uint32_t id, float* sh,
                                  float q[4], float so[4], bool need_sh, uint32_t* op_bits = nullptr) {
        const uint4* b = reinterpret_cast<const uint4*>(blocks) + (size_t)id * 8;
        const uint4 pw = __ldg(b), w = __ldg(b + 1);
        if (op_bits) *op_bits = w.w & 0xFFFFu;
        q[0] = hi(w.x); q[1] = lo(w.x); q[2] = hi(w.y); q[3] = lo(w.y);
        so[0] = hi(w.z); so[1] = lo(w.z); so[2] = hi(w.w); so[3] = lo(w.w);
        if (need_sh) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const uint4 v = __ldg(b + 2 + i);
                sh[8 * i] = lo(v.x); sh[8 * i + 1] = hi(v.x); sh[8 * i + 2] = lo(v.y); sh[8 * i + 3] = hi(v.y);
                sh[8 * i + 4] = lo(v.z); sh[8 * i + 5] = hi(v.z); sh[8 * i + 6] = lo(v.w); sh[8 * i + 7] = hi(v.w);
            }
        }
        return make_float4(__uint_as_float(pw.x), __uint_as_float(pw.y), __uint_as_float(pw.z), __uint_as_float(pw.w));
    }
};
template <>
struct Attr<false, false> {   // f32 planes: src/gaussian/f32.rs:53-175, planar.wgsl:334-364
    __device__ static float4 load(const float4* pos, const void* sh_p, const void* rot_p, const void* so_p, uint32_t id,
                                  float* sh, float q[4], float so[4], bool need_sh, uint32_t* = nullptr) {
        const float4 p = __ldg(pos + id);
        const float4 r = __ldg(reinterpret_cast<const float4*>(rot_p) + id);
        const float4 s = __ldg(reinterpret_cast<const float4*>(so_p) + id);
        q[0] = r.x; q[1] = r.y; q[2] = r.z; q[3] = r.w;
        so[0] = s.x; so[1] = s.y; so[2] = s.z; so[3] = s.w;
        if (need_sh) {
            const float4* p = reinterpret_cast<const float4*>(sh_p) + (size_t)id * 12;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float4 v = __ldg(p + i);
                sh[4 * i] = v.x; sh[4 * i + 1] = v.y; sh[4 * i + 2] = v.z; sh[4 * i + 3] = v.w;
            }
        }
        return p;
    }
};
template <>
struct Attr<true, false> {    // f16 planes: src/gaussian/f16.rs:30-56,244-263; planar.wgsl:117-176
    __device__ static float lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
    __device__ static float hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
    __device__ static float4 load(const float4* pos, const void* sh_p, const void* rso_p, const void*, uint32_t id, float* sh,
                                  float q[4], float so[4], bool need_sh, uint32_t* op_bits = nullptr) {
        const float4 p = __ldg(pos + id);
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(rso_p) + id);
        if (op_bits) *op_bits = w.w & 0xFFFFu;
        q[0] = hi(w.x); q[1] = lo(w.x); q[2] = hi(w.y); q[3] = lo(w.y);
        so[0] = hi(w.z); so[1] = lo(w.z); so[2] = hi(w.w); so[3] = lo(w.w);
        if (need_sh) {
            const uint4* p = reinterpret_cast<const uint4*>(sh_p) + (size_t)id * 6;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const uint4 v = __ldg(p + i);
                sh[8 * i] = lo(v.x); sh[8 * i + 1] = hi(v.x); sh[8 * i + 2] = lo(v.y); sh[8 * i + 3] = hi(v.y);
                sh[8 * i + 4] = lo(v.z); sh[8 * i + 5] = hi(v.z); sh[8 * i + 6] = lo(v.w); sh[8 * i + 7] = hi(v.w);
            }
        }
        return p;
    }
};

// upload-time repack of the planes into gaussian-major blocks (one thread per 16 B chunk; coalesced both ways)
template <bool F16>
__global__ void repack_kernel(const uint4* __restrict__ pos, const uint4* __restrict__ sh, const uint4* __restrict__ rot,
                              const uint4* __restrict__ so, uint32_t n, uint4* __restrict__ blocks) {
    constexpr uint32_t CH = F16 ? 8u : 16u;          // 16 B chunks per block
    constexpr uint32_t SHC = F16 ? 6u : 12u;         // sh chunks per gaussian
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * CH) return;
    const uint32_t id = (uint32_t)(i / CH), c = (uint32_t)(i % CH);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (c == 0) v = pos[id];
    else if (c == 1) v = rot[id];                    // f16: packed rotation + scale + opacity
    else if (!F16 && c == 2) v = so[id];
    else {
        const uint32_t k = c - (F16 ? 2u : 3u);
        if (k < SHC) v = sh[(size_t)id * SHC + k];
    }
    blocks[i] = v;
}
void launch_repack(bool f16, const void* pos, const void* sh, const void* rot, const void* so, uint32_t n, void* blocks,
                   cudaStream_t stream) {
    const size_t total = (size_t)n * (f16 ? 8 : 16);
    const uint32_t grid = (uint32_t)((total + 255) / 256);
    if (f16) repack_kernel<true><<<grid, 256, 0, stream>>>((const uint4*)pos, (const uint4*)sh, (const uint4*)rot, (const uint4*)so, n, (uint4*)blocks);
    else repack_kernel<false><<<grid, 256, 0, stream>>>((const uint4*)pos, (const uint4*)sh, (const uint4*)rot, (const uint4*)so, n, (uint4*)blocks);
}

// RasterizeMode::Depth (gaussian.wgsl:329-349): min distance from sorted[N-1], max from sorted[1] of the
// reference's FULL sorted buffer (culled entries keyed all-ones sit at its end, in index order) -- literal,
// including the [1] (not [0]) and the fact that the "nearest" entry is a culled gaussian whenever one exists.
__global__ void depth_range_kernel(const float4* __restrict__ pos, uint32_t n, const uint32_t* __restrict__ sorted_payload,
                                   const uint32_t* __restrict__ slot_ids /* null: payload is the gaussian index */,
                                   FrameCounters* __restrict__ ctr, FrameConsts fc) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || n < 2u) return;
    const uint32_t n_vis = ctr->n_vis, n_sorted = ctr->n_sort;
    auto id_at = [&](uint32_t i) { const uint32_t p = sorted_payload[i]; return slot_ids ? slot_ids[p] : p; };
    uint32_t first, last;
    if (n_sorted == n) {                       // SORT_ALL: the full buffer is materialised
        first = id_at(1u); last = id_at(n - 1u);
    } else {
        const uint32_t cmin = 0xFFFFFFFFu - ctr->culled_min_inv, cmax = ctr->culled_max_p1 - 1u;
        first = n_vis >= 2u ? id_at(1u) : cmin;            // n_vis == 0 draws nothing: value irrelevant
        last = (n - n_vis) >= 1u ? cmax : id_at(n - 1u);
    }
    auto dist = [&](uint32_t id) {
        const float4 p = pos[id];
        float pw[4];
        mat4_point(fc.model, p.x, p.y, p.z, pw);
        const float d[3] = {pw[0] - fc.cam[0], pw[1] - fc.cam[1], pw[2] - fc.cam[2]};
        return sqrtf(dot3(d, d));
    };
    ctr->depth_min = dist(last);
    ctr->depth_max = dist(first);
}

// gaussian.wgsl:228-232: cutoff = sqrt(max(9 + 2 ln(opacity), 1e-6)) (fixed-series ln, see project_math.cuh)
__device__ __forceinline__ float adaptive_cutoff(float opacity) {
    const float a = 9.0f + 2.0f * det_ln(opacity);
    return sqrtf(a > 0.000001f ? a : 0.000001f);
}
// f16 clouds: the adaptive cutoff depends on the 16-bit opacity alone -> a 65536-entry table built once per context
// by this very function (bit-identical to evaluating it in place, ~70 f64 instructions cheaper per gaussian)
__global__ void cutoff_table_kernel(float* __restrict__ tab) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < 65536u) tab[h] = adaptive_cutoff(__half2float(__ushort_as_half((unsigned short)h)));
}
void launch_cutoff_table(float* tab, cudaStream_t stream) { cutoff_table_kernel<<<256, 256, 0, stream>>>(tab); }

// One visible gaussian: projection + colour -> the 48 B record at recs[r] (+ the 2DGS extra record).
// `load_sh(float sh[48])` fetches the SH coefficients (only called for RasterizeMode::Color); `cutoff_pre` is the
// tabulated adaptive cutoff, or NaN to evaluate it here.
template <class ShLoader>
__device__ __forceinline__ void project_one(const FrameConsts& fc, const FrameCounters* __restrict__ ctr, uint32_t r, float4 p4,
                                            const float q[4], const float so[4], float cutoff_pre, ShLoader load_sh,
                                            SplatRec* __restrict__ recs, float4* __restrict__ extra, float4* __restrict__ aux) {
    SplatRec rec;
    rec.ux = 0.f; rec.uy = 0.f; rec.vx = 0.f; rec.vy = 0.f;
    rec.bx = BBOX_EMPTY; rec.by = BBOX_EMPTY;
    rec.r = 0.f; rec.g = 0.f; rec.b = 0.f;

    const KeyOut k = key_of(fc, p4.x, p4.y, p4.z);
    const float W = fc.W, H = fc.H;
    const float hw = 0.5f * W, hh = 0.5f * H;
    const float cx = k.ndc[0] * hw + hw;
    const float cy = hh - k.ndc[1] * hh;
    rec.cx = cx; rec.cy = cy;
    const float opacity = so[3];
    rec.op = opacity * fc.global_opacity;
    const bool drawn = k.visible && !(fc.draw_mode == BGS_DRAW_SELECTED && p4.w < 0.5f);

    float A[3][3];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) A[rr][cc] = fc.model[cc * 4 + rr];
    // helpers.wgsl:137-158 as (row, col); the quaternion is NOT normalised
    float Rm[3][3];
    {
        const float qr = q[0], x = q[1], y = q[2], z = q[3];
        Rm[0][0] = 1.0f - 2.0f * (y * y + z * z);
        Rm[1][0] = 2.0f * (x * y - qr * z);
        Rm[2][0] = 2.0f * (x * z + qr * y);
        Rm[0][1] = 2.0f * (x * y + qr * z);
        Rm[1][1] = 1.0f - 2.0f * (x * x + z * z);
        Rm[2][1] = 2.0f * (y * z - qr * x);
        Rm[0][2] = 2.0f * (x * z - qr * y);
        Rm[1][2] = 2.0f * (y * z + qr * x);
        Rm[2][2] = 1.0f - 2.0f * (x * x + y * y);
    }
    const float sc[3] = {so[0] * fc.global_scale, so[1] * fc.global_scale, so[2] * fc.global_scale};

    if (drawn) {
        float cutoff = 3.0f;
        if (fc.adaptive) {   // gaussian.wgsl:228-232
            cutoff = cutoff_pre == cutoff_pre ? cutoff_pre : adaptive_cutoff(opacity);
        }
        if (fc.gaussian_mode == BGS_GAUSSIAN_3D) {
        // gaussian_3d.wgsl:49-72
        float M[3][3], Sg[3][3], X[3][3], TS[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) M[i][j] = sc[i] * Rm[i][j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Sg[i][j] = (M[0][i] * M[0][j] + M[1][i] * M[1][j]) + M[2][i] * M[2][j];
        // identity model: T Sigma T^t == Sigma bit for bit as long as every entry of Sigma is finite and NON-ZERO
        // (x*1 + y*0 + z*0 then only ever adds +-0 to a non-zero value); zero entries (axis-aligned splats) keep
        // the multiply so that even the sign of a zero matches the oracle
        const float mn = fminf(fminf(fminf(fabsf(Sg[0][0]), fabsf(Sg[0][1])), fminf(fabsf(Sg[0][2]), fabsf(Sg[1][1]))),
                               fminf(fabsf(Sg[1][2]), fabsf(Sg[2][2])));
        const float mx = fmaxf(fmaxf(fmaxf(fabsf(Sg[0][0]), fabsf(Sg[0][1])), fmaxf(fabsf(Sg[0][2]), fabsf(Sg[1][1]))),
                               fmaxf(fabsf(Sg[1][2]), fabsf(Sg[2][2])));
        if (fc.model_identity && mn > 0.0f && mx < __uint_as_float(0x7F800000u) && Sg[0][0] == Sg[0][0] &&
            Sg[0][1] == Sg[0][1] && Sg[0][2] == Sg[0][2] && Sg[1][1] == Sg[1][1] && Sg[1][2] == Sg[1][2] && Sg[2][2] == Sg[2][2]) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) TS[i][j] = Sg[i][j];
        } else {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) X[i][j] = (A[i][0] * Sg[0][j] + A[i][1] * Sg[1][j]) + A[i][2] * Sg[2][j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) TS[i][j] = (X[i][0] * A[j][0] + X[i][1] * A[j][1]) + X[i][2] * A[j][2];
        }
        float c3[6] = {TS[0][0], TS[1][0], TS[2][0], TS[1][1], TS[2][1], TS[2][2]};
        if (fc.cov_pre) {
            // PRECOMPUTE_COVARIANCE_3D (gaussian_3d.wgsl:78-79, planar.wgsl:133-152): the decoded record goes straight into
            // cov2d -- no global_scale, no model 3x3.  It arrives in the plane slots it occupies: q = c0..c3, so = c4, c5,
            // opacity, opacity
            c3[0] = q[0]; c3[1] = q[1]; c3[2] = q[2]; c3[3] = q[3]; c3[4] = so[0]; c3[5] = so[1];
        }
        const float Vrk[3][3] = {{c3[0], c3[1], c3[2]}, {c3[1], c3[3], c3[4]}, {c3[2], c3[4], c3[5]}};
        // helpers.wgsl:8-47
        float tv[4];
        mat4_point(fc.view_from_world, k.pw[0], k.pw[1], k.pw[2], tv);
        const float fx = fc.p00 * W, fy = fc.p11 * H;
        const float sz = 1.0f / (tv[2] * tv[2]);
        const float J00 = fx / tv[2], J20 = -(fx * tv[0]) * sz;
        const float J11 = -fy / tv[2], J21 = (fy * tv[1]) * sz;
        float Tm[3][2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float v0 = fc.view_from_world[i * 4 + 0], v1 = fc.view_from_world[i * 4 + 1],
                        v2 = fc.view_from_world[i * 4 + 2];
            Tm[i][0] = v0 * J00 + v2 * J20;
            Tm[i][1] = v1 * J11 + v2 * J21;
        }
        float Y[3][2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int b = 0; b < 2; ++b) Y[i][b] = (Vrk[i][0] * Tm[0][b] + Vrk[i][1] * Tm[1][b]) + Vrk[i][2] * Tm[2][b];
        const float a = ((Tm[0][0] * Y[0][0] + Tm[1][0] * Y[1][0]) + Tm[2][0] * Y[2][0]) + 0.3f;
        const float b = (Tm[0][1] * Y[0][0] + Tm[1][1] * Y[1][0]) + Tm[2][1] * Y[2][0];
        const float c = ((Tm[0][1] * Y[0][1] + Tm[1][1] * Y[1][1]) + Tm[2][1] * Y[2][1]) + 0.3f;
        // helpers.wgsl:49-67
        const float det = a * c - b * b;
        const float mid = 0.5f * (a + c);
        const float disc = fmaxf(0.0f, mid * mid - det);
        const float term = sqrtf(disc);
        const float l1 = mid + term;
        if (fc.aabb) {
            // helpers.wgsl:69-79 + gaussian.wgsl:299-309: square of half-side cutoff*sqrt(l1), conic
            // record (USE_AABB): ux,uy,vx = conic.x,.y,.z; vy = quad half-side (half-pixels)
            const float l2 = fmaxf(mid - term, 0.0f);
            const float Rq = cutoff * fmaxf(sqrtf(l1), sqrtf(l2));
            const float dinv = 1.0f / det;
            rec.ux = c * dinv; rec.uy = -b * dinv; rec.vx = a * dinv; rec.vy = Rq;
            const float h = 0.5f * Rq;
            make_bbox(cx, cy, h, h, fc.Wi, fc.Hi, rec.bx, rec.by);
        } else {
        // helpers.wgsl:81-119 (USE_OBB)
        const float aa = (a - c) * (a - c);
        const float bb = sqrtf(aa + (4.0f * b) * b);
        const float major = sqrtf(((a + c) + bb) * 0.5f);
        const float minor = sqrtf(((a + c) - bb) * 0.5f);
        const float Bx = cutoff * major, By = cutoff * minor;
        const float evx = -b, evy = l1 - a;
        const float el = sqrtf(evx * evx + evy * evy);
        const float e1x = evx / el, e1y = evy / el;
        const float e2x = e1y, e2y = -e1x;
        rec.ux = (2.0f * e1x) / Bx; rec.uy = (-2.0f * e1y) / Bx;
        rec.vx = (2.0f * e2x) / By; rec.vy = (-2.0f * e2y) / By;
        const float hx = 0.5f * (fabsf(e1x) * Bx + fabsf(e2x) * By);
        const float hy = 0.5f * (fabsf(e1y) * Bx + fabsf(e2y) * By);
        if (rec.ux == rec.ux && rec.uy == rec.uy && rec.vx == rec.vx && rec.vy == rec.vy)
            make_bbox(cx, cy, hx, hy, fc.Wi, fc.Hi, rec.bx, rec.by);
        }
        } else {
            // ---- 2DGS surfel: gaussian_2d.wgsl:77-132 (homography) + :49-75 (quad)
            float L[3][2];   // first two columns of A * R_std * S, R_std = transpose(Rm)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float rc[3] = {Rm[j][0] * sc[j], Rm[j][1] * sc[j], Rm[j][2] * sc[j]};
#pragma unroll
                for (int i = 0; i < 3; ++i) L[i][j] = (A[i][0] * rc[0] + A[i][1] * rc[1]) + A[i][2] * rc[2];
            }
            float G[3][4];
            mat4_dir(fc.clip_from_world, L[0][0], L[1][0], L[2][0], G[0]);
            mat4_dir(fc.clip_from_world, L[0][1], L[1][1], L[2][1], G[1]);
            mat4_point(fc.clip_from_world, k.pw[0], k.pw[1], k.pw[2], G[2]);
            const float fxk = fc.p00 * W / 2.0f, fyk = fc.p11 * H / 2.0f;     // helpers.wgsl:122-135
            const float cxk = (W - 1.0f) / 2.0f, cyk = (H - 1.0f) / 2.0f;
            float T0[3], T1[3], T2[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                T0[j] = fxk * G[j][0] + cxk * G[j][3];
                T1[j] = fyk * G[j][1] + cyk * G[j][3];
                T2[j] = G[j][3];
            }
            const float c2 = cutoff * cutoff;
            const float test[3] = {c2, c2, -1.0f};
            const float tt[3] = {test[0] * T2[0], test[1] * T2[1], test[2] * T2[2]};
            const float d = dot3(tt, T2);
            float Rq = 0.0f, mean0 = 0.0f, mean1 = 0.0f;
            bool ok = !(fabsf(d) < 1.0e-4f);
            if (ok) {
                const float inv = 1.0f / d;
                const float f[3] = {inv * test[0], inv * test[1], inv * test[2]};
                const float t02[3] = {T0[0] * T2[0], T0[1] * T2[1], T0[2] * T2[2]};
                const float t12[3] = {T1[0] * T2[0], T1[1] * T2[1], T1[2] * T2[2]};
                mean0 = dot3(f, t02); mean1 = dot3(f, t12);
                const float f0[3] = {f[0] * T0[0], f[1] * T0[1], f[2] * T0[2]};
                const float f1[3] = {f[0] * T1[0], f[1] * T1[1], f[2] * T1[2]};
                const float ex = mean0 * mean0 - dot3(f0, T0);
                const float ey = mean1 * mean1 - dot3(f1, T1);
                if (ex < 1.0e-4f || ey < 1.0e-4f) ok = false;
                else Rq = fmaxf(fmaxf(sqrtf(ex), sqrtf(ey)), cutoff * 0.707106f);
            }
            if (ok) {
                rec.ux = 2.0f / Rq; rec.uy = 0.0f; rec.vx = 0.0f; rec.vy = -2.0f / Rq;   // OBB branch: e1=(1,0), e2=(0,1)
                const float h = 0.5f * Rq;
                if (Rq == Rq) make_bbox(cx, cy, h, h, fc.Wi, fc.Hi, rec.bx, rec.by);
                if (fc.aabb && extra != nullptr) {
                    float4* e = extra + (size_t)r * 4;
                    e[0] = make_float4(Rq, mean0, mean1, W / H);
                    e[1] = make_float4(T0[0], T0[1], T0[2], 0.0f);
                    e[2] = make_float4(T1[0], T1[1], T1[2], 0.0f);
                    e[3] = make_float4(T2[0], T2[1], T2[2], 0.0f);
                }
            }
        }

        // colour source
        float rgb[3] = {0.f, 0.f, 0.f};
        if (fc.rasterize_mode == BGS_RASTERIZE_COLOR) {
            float sh[48];
            load_sh(sh);
            // gaussian.wgsl:166-183,406-416
            const float dlt[3] = {k.pw[0] - fc.cam[0], k.pw[1] - fc.cam[1], k.pw[2] - fc.cam[2]};
            float dw[3], loc[3], dl[3];
            normalize3(dlt, dw);
            if (fc.model_identity) {     // the normalised model columns are the unit axes
                loc[0] = dw[0]; loc[1] = dw[1]; loc[2] = dw[2];
            } else {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) {
                    const float col[3] = {A[0][cc], A[1][cc], A[2][cc]};
                    float bn[3];
                    normalize3(col, bn);
                    loc[cc] = dot3(bn, dw);
                }
            }
            normalize3(loc, dl);
            // spherical_harmonics.wgsl:34-68
            const float x = dl[0], y = dl[1], z = dl[2];
            const float xx = x * x, yy = y * y, zz = z * z;
            float basis[16];   // (the SH constants are folded in below: 16 multiplies instead of 48)
            basis[0] = 1.0f;
            basis[1] = y; basis[2] = z; basis[3] = x;
            basis[4] = x * y; basis[5] = y * z; basis[6] = (2.0f * zz - xx) - yy;
            basis[7] = x * z; basis[8] = xx - yy;
            basis[9] = y * (3.0f * xx - yy);
            basis[10] = (x * y) * z;
            basis[11] = y * ((4.0f * zz - xx) - yy);
            basis[12] = z * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
            basis[13] = x * ((4.0f * zz - xx) - yy);
            basis[14] = z * (xx - yy);
            basis[15] = x * (xx - 3.0f * yy);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) basis[kk] *= c_shc[kk];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                float acc = 0.5f;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) acc = fmaf(sh[3 * kk + cc], basis[kk], acc);   // colour: FMA is fine
                rgb[cc] = acc;
            }
            if (fc.color_space == 0u) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) rgb[cc] = srgb_to_linear(rgb[cc]);
            }
        }
        // aux outputs (bgs_render_aux, config C4): the Depth and Normal colour sources ride along with the main one, so one
        // pass yields what three single-mode frames would (the geometry and alpha of a splat do not depend on the mode)
        float drgb[3] = {0.f, 0.f, 0.f}, nrgb[3] = {0.f, 0.f, 0.f};
        if (fc.rasterize_mode == BGS_RASTERIZE_DEPTH || fc.aux) {
            float* rgb = drgb;
            // material/depth.wgsl:3-11
            const float dlt[3] = {k.pw[0] - fc.cam[0], k.pw[1] - fc.cam[1], k.pw[2] - fc.cam[2]};
            const float depth = sqrtf(dot3(dlt, dlt));
            const float dmin = ctr->depth_min, dmax = ctr->depth_max;
            if (fc.n_cloud >= 2u) {   // the reference reads sorted[1]: undefined for a 1-gaussian cloud (oracle: black)
            float nd = (depth - dmin) / (dmax - dmin);
            nd = fminf(fmaxf(nd, 0.0f), 1.0f);   // fmin/fmax ignore a NaN operand, like the oracle's
            float t1 = (nd - 0.5f) / (1.0f - 0.5f); t1 = fminf(fmaxf(t1, 0.0f), 1.0f);
            float t2 = (nd - 0.0f) / (0.5f - 0.0f); t2 = fminf(fmaxf(t2, 0.0f), 1.0f);
            rgb[0] = t1 * t1 * (3.0f - 2.0f * t1);
            rgb[1] = 1.0f - fabsf(nd - 0.5f) * 2.0f;
            rgb[2] = 1.0f - t2 * t2 * (3.0f - 2.0f * t2);
            }
        }
        if (fc.rasterize_mode == BGS_RASTERIZE_POSITION) {
            // gaussian.wgsl:375-376: (transformed_position - min) / (max - min)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) rgb[cc] = (k.pw[cc] - fc.aabb_min[cc]) / (fc.aabb_max[cc] - fc.aabb_min[cc]);
        }
        if (fc.rasterize_mode == BGS_RASTERIZE_NORMAL || fc.aux) {
            float* rgb = nrgb;
            // gaussian.wgsl:350-368
            float SR[3], Ln[3], wn[4];
#pragma unroll
            for (int i = 0; i < 3; ++i) SR[i] = sc[i] * Rm[i][2];
#pragma unroll
            for (int i = 0; i < 3; ++i) Ln[i] = (A[i][0] * SR[0] + A[i][1] * SR[1]) + A[i][2] * SR[2];
            mat4_dir(fc.view_from_world, Ln[0], Ln[1], Ln[2], wn);
            const float l = sqrtf(((wn[0] * wn[0] + wn[1] * wn[1]) + wn[2] * wn[2]) + wn[3] * wn[3]);
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) rgb[cc] = 0.5f * (wn[cc] / l + 1.0f);
        }
        if (fc.rasterize_mode == BGS_RASTERIZE_DEPTH) { rgb[0] = drgb[0]; rgb[1] = drgb[1]; rgb[2] = drgb[2]; }
        if (fc.rasterize_mode == BGS_RASTERIZE_NORMAL) { rgb[0] = nrgb[0]; rgb[1] = nrgb[1]; rgb[2] = nrgb[2]; }
        rec.r = rgb[0]; rec.g = rgb[1]; rec.b = rgb[2];
        if (fc.draw_mode == BGS_DRAW_HIGHLIGHT_SELECTED && p4.w > 0.5f) {   // gaussian.wgsl:423-427
            rec.r = 0.3f; rec.g = 1.0f; rec.b = 0.1f; rec.op = 1.0f;
            drgb[0] = nrgb[0] = 0.3f; drgb[1] = nrgb[1] = 1.0f; drgb[2] = nrgb[2] = 0.1f;
        }
        if (fc.aux && aux != nullptr) {
            aux[(size_t)r * 2] = make_float4(drgb[0], drgb[1], drgb[2], 0.0f);
            aux[(size_t)r * 2 + 1] = make_float4(nrgb[0], nrgb[1], nrgb[2], 0.0f);
        }
    }
    // 48 B record, three 16 B stores
    float4* out = reinterpret_cast<float4*>(recs + r);
    out[0] = make_float4(rec.cx, rec.cy, rec.ux, rec.uy);
    out[1] = make_float4(rec.vx, rec.vy, __uint_as_float(rec.bx), __uint_as_float(rec.by));
    out[2] = make_float4(rec.r, rec.g, rec.b, rec.op);
}

#ifndef PROJ_MIN_CTAS
#define PROJ_MIN_CTAS 6
#endif
template <bool F16, bool BLOCKED>
__global__ void __launch_bounds__(128, PROJ_MIN_CTAS)
project_kernel(const float4* __restrict__ pos, const void* __restrict__ sh_p, const void* __restrict__ rot_p,
               const void* __restrict__ so_p, const uint32_t* __restrict__ index_list, int by_slot,
               const FrameCounters* __restrict__ ctr, FrameConsts fc, SplatRec* __restrict__ recs,
               float4* __restrict__ extra /* 4 x float4 per record, 2DGS + USE_AABB only */, const float* __restrict__ cutoff_tab,
               float4* __restrict__ aux /* 2 x float4 per record (depth rgb, normal rgb), bgs_render_aux only */) {
    const uint32_t n_vis = ctr->n_vis;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_vis; r += gridDim.x * blockDim.x) {
        // by_slot: r is a compact slot (ascending gaussian index; runs concurrently with the depth sort)
        // else   : r is a front-to-back rank, the list is the far->near sorted index list
        const uint32_t id = by_slot ? __ldg(index_list + r) : __ldg(index_list + (n_vis - 1u - r));
        float sh[48], q[4], so[4];
        const bool need_sh = fc.rasterize_mode == BGS_RASTERIZE_COLOR;
        uint32_t op_bits = 0u;
        const float4 p4 = Attr<F16, BLOCKED>::load(pos, sh_p, rot_p, so_p, id, sh, q, so, need_sh, &op_bits);
        // f16 clouds: the adaptive cutoff of this 16-bit opacity comes from the per-context table (bit-identical)
        const float cutoff_pre = (F16 && fc.adaptive) ? __ldg(cutoff_tab + op_bits) : __uint_as_float(0x7FC00000u);
        project_one(fc, ctr, r, p4, q, so, cutoff_pre,
                    [&](float* out) {
#pragma unroll
                        for (int i = 0; i < 48; ++i) out[i] = sh[i];
                    },
                    recs, extra, aux);
    }
}

// ---- TMA ring variant (gaussian-major blocks): each warp streams batches of 32 gaussians through a ring of
// shared-memory stages.  Every lane issues ONE bulk async copy (cp.async.bulk, SASS UBLKCP) of its gaussian's whole
// block -- 128 B (f16) / 256 B (f32) -- tracked by the stage's mbarrier (expect_tx = bytes of the batch); the copies
// of the next STAGES - 1 batches are in flight while the warp computes the current one from shared memory, so the
// gather's latency hides behind the ~1000 instructions per gaussian regardless of the register budget.
// Rows are padded by 16 B (144 / 272 B stride): lane l reading 16 B chunk c touches bank group (l + c) mod 8, so the
// per-lane LDS.128 of a batch are conflict-free.
constexpr int PR_THREADS = 128;
constexpr int PR_WARPS = PR_THREADS / 32;
template <bool F16> struct Ring {
    static constexpr int ROW = F16 ? 128 : 256;
    static constexpr int ROWP = ROW + 16;
    static constexpr int STAGES = F16 ? 3 : 2;
    static constexpr int WARP_BYTES = STAGES * 32 * ROWP;
    static constexpr int SMEM = PR_WARPS * WARP_BYTES;
};
__device__ __forceinline__ void pr_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void pr_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pr_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void pr_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ float h_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float h_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

template <bool F16>
__global__ void __launch_bounds__(PR_THREADS, F16 ? 4 : 3)
project_ring_kernel(const void* __restrict__ blocks, const uint32_t* __restrict__ index_list, int by_slot,
                    const FrameCounters* __restrict__ ctr, FrameConsts fc, SplatRec* __restrict__ recs,
                    float4* __restrict__ extra, const float* __restrict__ cutoff_tab) {
    using R = Ring<F16>;
    extern __shared__ __align__(128) unsigned char s_ring[];
    __shared__ __align__(8) unsigned long long s_bar[PR_WARPS][R::STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_vis = ctr->n_vis;
    const uint32_t nb = (n_vis + 31u) >> 5;
    const uint32_t wg = blockIdx.x * PR_WARPS + warp, nw = gridDim.x * PR_WARPS;
    const uint32_t a_ring = (uint32_t)__cvta_generic_to_shared(s_ring) + (uint32_t)warp * R::WARP_BYTES;
    const uint32_t a_bar = (uint32_t)__cvta_generic_to_shared(&s_bar[warp][0]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < R::STAGES; ++s) pr_mbar_init(a_bar + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](uint32_t k) {
        const uint32_t batch = wg + k * nw;
        if (batch >= nb) return;                       // (warp-uniform)
        const int s = (int)(k % R::STAGES);
        const uint32_t r = batch * 32u + lane;
        const bool valid = r < n_vis;
        uint32_t id = 0u;
        if (valid) id = by_slot ? __ldg(index_list + r) : __ldg(index_list + (n_vis - 1u - r));
        const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, valid));
        if (lane == 0) pr_mbar_expect_tx(a_bar + 8u * s, cnt * (uint32_t)R::ROW);
        __syncwarp();
        if (valid)
            pr_bulk_g2s(a_ring + (uint32_t)(s * 32 + lane) * R::ROWP, reinterpret_cast<const char*>(blocks) + (size_t)id * R::ROW,
                        (uint32_t)R::ROW, a_bar + 8u * s);
    };
#pragma unroll
    for (int k = 0; k < R::STAGES; ++k) issue((uint32_t)k);

    for (uint32_t k = 0;; ++k) {
        const uint32_t batch = wg + k * nw;
        if (batch >= nb) break;
        const int s = (int)(k % R::STAGES);
        pr_mbar_wait(a_bar + 8u * s, (k / R::STAGES) & 1u);
        const uint32_t r = batch * 32u + lane;
        if (r < n_vis) {
            const uint4* row = reinterpret_cast<const uint4*>(s_ring + (size_t)warp * R::WARP_BYTES + (size_t)(s * 32 + lane) * R::ROWP);
            float q[4], so[4];
            float cutoff_pre = __uint_as_float(0x7FC00000u);
            const uint4 pw = row[0];
            const float4 p4 = make_float4(__uint_as_float(pw.x), __uint_as_float(pw.y), __uint_as_float(pw.z), __uint_as_float(pw.w));
            if (F16) {
                const uint4 w = row[1];
                q[0] = h_hi(w.x); q[1] = h_lo(w.x); q[2] = h_hi(w.y); q[3] = h_lo(w.y);
                so[0] = h_hi(w.z); so[1] = h_lo(w.z); so[2] = h_hi(w.w); so[3] = h_lo(w.w);
                if (fc.adaptive) cutoff_pre = __ldg(cutoff_tab + (w.w & 0xFFFFu));
            } else {
                const uint4 a = row[1], b = row[2];
                q[0] = __uint_as_float(a.x); q[1] = __uint_as_float(a.y); q[2] = __uint_as_float(a.z); q[3] = __uint_as_float(a.w);
                so[0] = __uint_as_float(b.x); so[1] = __uint_as_float(b.y); so[2] = __uint_as_float(b.z); so[3] = __uint_as_float(b.w);
            }
            project_one(fc, ctr, r, p4, q, so, cutoff_pre,
                        [&](float* sh) {
                            if (F16) {
#pragma unroll
                                for (int i = 0; i < 6; ++i) {
                                    const uint4 v = row[2 + i];
                                    sh[8 * i] = h_lo(v.x); sh[8 * i + 1] = h_hi(v.x); sh[8 * i + 2] = h_lo(v.y); sh[8 * i + 3] = h_hi(v.y);
                                    sh[8 * i + 4] = h_lo(v.z); sh[8 * i + 5] = h_hi(v.z); sh[8 * i + 6] = h_lo(v.w); sh[8 * i + 7] = h_hi(v.w);
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 12; ++i) {
                                    const uint4 v = row[3 + i];
                                    sh[4 * i] = __uint_as_float(v.x); sh[4 * i + 1] = __uint_as_float(v.y);
                                    sh[4 * i + 2] = __uint_as_float(v.z); sh[4 * i + 3] = __uint_as_float(v.w);
                                }
                            }
                        },
                        recs, extra, nullptr);
        }
        __syncwarp();                                  // every lane is done with the stage before it is refilled
        issue(k + (uint32_t)R::STAGES);
    }
}

void launch_depth_range(const float4* pos, uint32_t n, const uint32_t* sorted_payload, const uint32_t* slot_ids,
                        FrameCounters* ctr, const FrameConsts& fc, cudaStream_t stream) {
    depth_range_kernel<<<1, 32, 0, stream>>>(pos, n, sorted_payload, slot_ids, ctr, fc);
}

void launch_project(bool f16, bool blocked, const float4* pos, const void* sh, const void* rot, const void* so,
                    const uint32_t* index_list, int by_slot, const FrameCounters* ctr, const FrameConsts& fc,
                    SplatRec* recs, float4* extra, uint32_t n_hint, int sm_count, int ctas_per_sm, const float* cutoff_tab,
                    float4* aux, cudaStream_t stream) {
    static int use_ring = -1;
    if (use_ring < 0) { const char* e = getenv("BGS_PROJECT_RING"); use_ring = (e && atoi(e) > 0) ? 1 : 0; }
    if (blocked && use_ring && aux == nullptr) {
        // (measured slower on B200 -- one UBLKCP per 128 B row sustains ~1 copy / 20 cycles / SM: 48 us vs 36 us for
        // the per-thread gather at C3 -- kept behind BGS_PROJECT_RING=1 as the evidence, profiles/r2_experiments.md)
        // TMA ring over the gaussian-major blocks (`sh` carries the block array): a persistent grid of at most
        // ctas_per_sm CTAs per SM (2 when a depth sort shares the SMs, else the occupancy limit)
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaFuncSetAttribute(project_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<true>::SMEM);
            cudaFuncSetAttribute(project_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ring<false>::SMEM);
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        const int occ = f16 ? 4 : 3;
        if (ctas_per_sm <= 0 || ctas_per_sm > occ) ctas_per_sm = occ;
        uint32_t blocks = (n_hint + PR_THREADS - 1) / PR_THREADS;
        const uint32_t cap = (uint32_t)(sm_count * ctas_per_sm);
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        if (f16) project_ring_kernel<true><<<blocks, PR_THREADS, Ring<true>::SMEM, stream>>>(sh, index_list, by_slot, ctr, fc, recs, extra, cutoff_tab);
        else project_ring_kernel<false><<<blocks, PR_THREADS, Ring<false>::SMEM, stream>>>(sh, index_list, by_slot, ctr, fc, recs, extra, cutoff_tab);
        return;
    }
    // per-thread gather (gaussian-major blocks by default, the reference's planes with BGS_LAYOUT=planar).  Grid sized from a hint (last frame's visible count +
    // head-room); the grid-stride loop keeps any n_vis correct.
    uint32_t blocks = (n_hint + 127) / 128;
    if (blocks > 65535u * 8u) blocks = 65535u * 8u;
    if (blocks < 148u) blocks = 148u;
    {   // tuning knob: cap the grid at BGS_PROJECT_CTAS CTAs per SM (grid-stride loop; smaller footprint beside other frames)
        static int cap = -1;
        if (cap < 0) { const char* e = getenv("BGS_PROJECT_CTAS"); cap = e ? atoi(e) : 0; }
        if (cap > 0 && blocks > (uint32_t)(cap * sm_count)) blocks = (uint32_t)(cap * sm_count);
    }
    // blocked layout: `sh` carries the block array
    if (f16 && blocked) project_kernel<true, true><<<blocks, 128, 0, stream>>>(pos, sh, rot, so, index_list, by_slot, ctr, fc, recs, extra, cutoff_tab, aux);
    else if (f16) project_kernel<true, false><<<blocks, 128, 0, stream>>>(pos, sh, rot, so, index_list, by_slot, ctr, fc, recs, extra, cutoff_tab, aux);
    else if (blocked) project_kernel<false, true><<<blocks, 128, 0, stream>>>(pos, sh, rot, so, index_list, by_slot, ctr, fc, recs, extra, cutoff_tab, aux);
    else project_kernel<false, false><<<blocks, 128, 0, stream>>>(pos, sh, rot, so, index_list, by_slot, ctr, fc, recs, extra, cutoff_tab, aux);
}

}  // namespace bgs
