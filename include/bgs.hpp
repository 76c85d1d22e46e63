// bgs.hpp -- C++ host mirror of the reference's plugin surface for the forward splat path, header-only,
// layered strictly above the C ABI of bgs.h.  (The reference host is Rust; no Rust toolchain exists in the
// build image, so the compiled-language host side is C++.  INTEGRATION.md has the Rust binding.)
//
// Names and defaults follow mosure/bevy_gaussian_splatting:
//   CloudSettings + enums            src/gaussian/settings.rs:6-133
//   PlanarGaussian3d                 src/gaussian/formats/planar_3d.rs:28-54 (struct of four planes)
//   random_gaussians_3d_seeded       src/gaussian/formats/planar_3d.rs:120-191 (distributions + field order)
//   GaussianCamera                   src/camera.rs:6-9
//   GaussianSplattingPlugin          src/lib.rs:48-80 -- here: owns the bgs_context of one GPU; render_view() is
//                                    run_radix_sort (src/sort/radix.rs:616-756) + DrawGaussians
//                                    (src/render/mod.rs:986-992) for one view
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "bgs.h"

namespace bgs {

enum class DrawMode : uint32_t { All = 0, Selected = 1, HighlightSelected = 2 };
enum class GaussianMode : uint32_t { Gaussian2d = 0, Gaussian3d = 1 };
enum class RasterizeMode : uint32_t { Color = 0, Depth = 1, Normal = 2, Position = 3 };
enum class RadixSortDepthBits : uint32_t { Bits16 = 16, Bits24 = 24, Bits32 = 32 };
enum class GaussianColorSpace : uint32_t { SrgbRec709Display = 0, LinRec709Display = 1 };

struct CloudSettings {   // src/gaussian/settings.rs:110-133 (defaults)
    bool aabb = false;
    float global_opacity = 1.0f;
    float global_scale = 1.0f;
    bool opacity_adaptive_radius = true;
    RadixSortDepthBits radix_sort_depth_bits = RadixSortDepthBits::Bits32;
    DrawMode draw_mode = DrawMode::All;
    GaussianMode gaussian_mode = GaussianMode::Gaussian3d;
    RasterizeMode rasterize_mode = RasterizeMode::Color;
    GaussianColorSpace color_space = GaussianColorSpace::SrgbRec709Display;
    float time = 0.0f;
    // this repo's extension: front-to-back binning rounds (BGS_FLAG_CHUNKS): -1 = the library's choice from the last
    // frame's footprint statistics, 1 = always, 0 = never (the tile debug hooks need a one-round frame)
    int binning_rounds = -1;

    bgs_settings to_abi(uint32_t flags = 0) const {
        if (binning_rounds >= 0) flags |= binning_rounds ? BGS_FLAG_CHUNKS : BGS_FLAG_NO_CHUNKS;
        bgs_settings s{};
        s.gaussian_mode = (uint32_t)gaussian_mode; s.rasterize_mode = (uint32_t)rasterize_mode;
        s.aabb = aabb; s.opacity_adaptive_radius = opacity_adaptive_radius; s.draw_mode = (uint32_t)draw_mode;
        s.radix_sort_depth_bits = (uint32_t)radix_sort_depth_bits; s.flags = flags;
        return s;
    }
};

struct ShaderDefines {   // src/render/mod.rs:698-760: the radix pass plan
    uint32_t radix_bits_per_digit, radix_digit_places, radix_key_shift, radix_base;
    static ShaderDefines for_radix_depth_bits(RadixSortDepthBits b) {
        const uint32_t bits = (uint32_t)b;
        return {8u, bits / 8u, 32u - bits, 256u};
    }
    uint32_t radix_initial_parity() const { return radix_digit_places % 2u; }
};

struct PlanarGaussian3d {   // four planes, binding order (planar_3d.rs:45-54)
    std::vector<float> position_visibility;   // n*4
    std::vector<float> spherical_harmonic;    // n*48, sh[3k + c]
    std::vector<float> rotation;              // n*4, (w, x, y, z)
    std::vector<float> scale_opacity;         // n*4
    size_t len() const { return position_visibility.size() / 4; }
    // the entity Aabb's min()/max(): compute_aabb (interface.rs:22-66, positions +- 0.1) -> Aabb {center, half_extents}
    // (cloud.rs:45-62) -> center -+ half_extents (render/mod.rs:1070-1071), all in f32
    void compute_aabb(float mn[3], float mx[3]) const {
        for (int k = 0; k < 3; ++k) {
            float lo = INFINITY, hi = -INFINITY;
            for (size_t i = 0; i < len(); ++i) {
                const float p = position_visibility[4 * i + k];
                lo = std::fmin(lo, p - 0.1f); hi = std::fmax(hi, p + 0.1f);
            }
            const float center = (lo + hi) / 2.0f, half = (hi - lo) / 2.0f;
            mn[k] = center - half; mx[k] = center + half;
        }
    }
};

// splitmix64-based counter PRNG: this repo's generator for the C++ host (the reference's ChaCha stream is not
// reproduced).  Same distributions and field order as planar_3d.rs:120-168.
inline PlanarGaussian3d random_gaussians_3d_seeded(size_t n, uint64_t seed) {
    PlanarGaussian3d c;
    c.position_visibility.resize(n * 4); c.spherical_harmonic.resize(n * 48); c.rotation.resize(n * 4); c.scale_opacity.resize(n * 4);
    uint64_t ctr = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto uni = [&ctr]() {
        uint64_t z = (ctr += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        return (float)(z >> 40) * (1.0f / 16777216.0f);   // [0, 1)
    };
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 4; ++k) c.rotation[4 * i + k] = uni() * 2.0f - 1.0f;
        for (int k = 0; k < 3; ++k) c.position_visibility[4 * i + k] = uni() * 40.0f - 20.0f;
        c.position_visibility[4 * i + 3] = 1.0f;
        for (int k = 0; k < 3; ++k) c.scale_opacity[4 * i + k] = uni();
        c.scale_opacity[4 * i + 3] = uni() * 0.8f;
        for (int k = 0; k < 48; ++k) c.spherical_harmonic[48 * i + k] = uni() * 2.0f - 1.0f;
    }
    return c;
}

struct GaussianCamera { bool warmup = false; };   // src/camera.rs:6-9

// Column-major 4x4 helpers (Bevy/glam conventions).
struct Mat4 {
    float m[16];
    static Mat4 identity() { Mat4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }
    Mat4 operator*(const Mat4& b) const {   // f32 arithmetic like glam
        Mat4 r{};
        for (int c = 0; c < 4; ++c) for (int rr = 0; rr < 4; ++rr) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += m[k * 4 + rr] * b.m[c * 4 + k];
            r.m[c * 4 + rr] = s;
        }
        return r;
    }
};
inline Mat4 perspective_infinite_reverse_rh(float fov_y, float aspect, float z_near) {
    const float f = 1.0f / std::tan(0.5f * fov_y);
    Mat4 r{};
    r.m[0] = f / aspect; r.m[5] = f; r.m[11] = -1.0f; r.m[14] = z_near;
    return r;
}
inline Mat4 view_from_translation(float x, float y, float z) {   // camera with identity rotation at (x, y, z)
    Mat4 r = Mat4::identity();
    r.m[12] = -x; r.m[13] = -y; r.m[14] = -z;
    return r;
}
inline bgs_view make_view(const Mat4& view_from_world, const Mat4& clip_from_view, const float eye[3], int w, int h) {
    bgs_view v{};
    std::memcpy(v.view_from_world, view_from_world.m, 64);
    std::memcpy(v.clip_from_view, clip_from_view.m, 64);
    const Mat4 cw = clip_from_view * view_from_world;
    std::memcpy(v.clip_from_world, cw.m, 64);
    std::memcpy(v.world_position, eye, 12);
    v.viewport[0] = 0; v.viewport[1] = 0; v.viewport[2] = (float)w; v.viewport[3] = (float)h;
    return v;
}
// examples/headless.rs:177-184: Camera3d at (0, 1.5, 5), identity rotation, default perspective
inline bgs_view headless_view(int w = 1920, int h = 1080) {
    const float eye[3] = {0.0f, 1.5f, 5.0f};
    return make_view(view_from_translation(eye[0], eye[1], eye[2]),
                     perspective_infinite_reverse_rh(3.14159265358979323846f / 4.0f, (float)w / (float)h, 0.1f), eye, w, h);
}

class Error : public std::runtime_error {
public:
    Error(bgs_status st, const std::string& msg) : std::runtime_error(msg), status(st) {}
    bgs_status status;
};

class GaussianSplattingPlugin;

class PlanarGaussian3dHandle {   // a cloud resident in HBM
public:
    PlanarGaussian3dHandle() = default;
    PlanarGaussian3dHandle(const PlanarGaussian3dHandle&) = delete;
    PlanarGaussian3dHandle& operator=(const PlanarGaussian3dHandle&) = delete;
    PlanarGaussian3dHandle(PlanarGaussian3dHandle&& o) noexcept : h_(o.h_), n_(o.n_) {
        std::memcpy(aabb_min_, o.aabb_min_, 12); std::memcpy(aabb_max_, o.aabb_max_, 12); o.h_ = nullptr;
    }
    const float* aabb_min() const { return aabb_min_; }
    const float* aabb_max() const { return aabb_max_; }
    ~PlanarGaussian3dHandle() { if (h_) bgs_cloud_destroy(h_); }
    bgs_cloud* get() const { return h_; }
    uint32_t len() const { return n_; }
private:
    friend class GaussianSplattingPlugin;
    bgs_cloud* h_ = nullptr;
    uint32_t n_ = 0;
    float aabb_min_[3] = {0, 0, 0}, aabb_max_[3] = {1, 1, 1};
};

class GaussianSplattingPlugin {
public:
    explicit GaussianSplattingPlugin(int cuda_device = 0) {
        const bgs_status st = bgs_context_create(cuda_device, &ctx_);
        if (st != BGS_OK) throw Error(st, "bgs_context_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    GaussianSplattingPlugin(const GaussianSplattingPlugin&) = delete;
    GaussianSplattingPlugin& operator=(const GaussianSplattingPlugin&) = delete;
    ~GaussianSplattingPlugin() { bgs_context_destroy(ctx_); }

    PlanarGaussian3dHandle add_cloud(const PlanarGaussian3d& c) {   // asset prepare
        PlanarGaussian3dHandle h;
        h.n_ = (uint32_t)c.len();
        c.compute_aabb(h.aabb_min_, h.aabb_max_);
        check(bgs_cloud_upload_f32(ctx_, h.n_, c.position_visibility.data(), c.spherical_harmonic.data(), c.rotation.data(),
                                   c.scale_opacity.data(), &h.h_));
        return h;
    }
    static bgs_cloud_uniform cloud_uniform(const CloudSettings& s, const Mat4& transform = Mat4::identity()) {
        bgs_cloud_uniform u{};
        std::memcpy(u.transform, transform.m, 64);
        u.global_opacity = s.global_opacity; u.global_scale = s.global_scale;
        u.color_space = (uint32_t)s.color_space; u.time = s.time;
        u.aabb_min[3] = u.aabb_max[3] = 1.0f;
        for (int k = 0; k < 3; ++k) u.aabb_max[k] = 1.0f;
        return u;
    }
    // One view of one cloud.  Returns false when the frame is skipped (warm-up camera / not ready), like the
    // reference's silent skip (src/render/mod.rs:361-371, src/sort/radix.rs:645-658).
    // `extra_flags`: BGS_FLAG_BLEND_OVER_TARGET (blend over what the target holds: one call per cloud, far cloud first, as
    // the reference's Transparent3d items do, render/mod.rs:398-452, :944-948), BGS_FLAG_PREMULTIPLIED_OUT (the layer alone).
    bool render_view(const PlanarGaussian3dHandle& cloud, const CloudSettings& settings, const bgs_view& view, void* out_rgba,
                     uint32_t format = BGS_FORMAT_RGBA8_SRGB, const GaussianCamera& camera = {}, bool out_is_device = false,
                     uint32_t extra_flags = 0) {
        if (camera.warmup) return false;
        bgs_cloud_uniform u = cloud_uniform(settings);
        std::memcpy(u.aabb_min, cloud.aabb_min(), 12); std::memcpy(u.aabb_max, cloud.aabb_max(), 12);
        const bgs_settings s = settings.to_abi(extra_flags);
        const bgs_status st = bgs_render(ctx_, cloud.get(), &view, &u, &s, out_rgba, format, out_is_device ? 1 : 0);
        if (st == BGS_NOT_READY) return false;
        check(st);
        return true;
    }
    // Colour + depth + normal frames of one view in one pass (BASELINE.json config 4; bgs_render_aux).
    bool render_view_aux(const PlanarGaussian3dHandle& cloud, const CloudSettings& settings, const bgs_view& view, void* out_rgba,
                         void* out_depth, void* out_normal, uint32_t format = BGS_FORMAT_RGBA8_SRGB, bool out_is_device = false) {
        bgs_cloud_uniform u = cloud_uniform(settings);
        std::memcpy(u.aabb_min, cloud.aabb_min(), 12); std::memcpy(u.aabb_max, cloud.aabb_max(), 12);
        const bgs_settings s = settings.to_abi();
        const bgs_status st = bgs_render_aux(ctx_, cloud.get(), &view, &u, &s, out_rgba, out_depth, out_normal, format, out_is_device ? 1 : 0);
        if (st == BGS_NOT_READY) return false;
        check(st);
        return true;
    }
    bgs_frame_stats frame_stats() { bgs_frame_stats fs{}; check(bgs_frame_stats_get(ctx_, &fs)); return fs; }
    bgs_context* context() const { return ctx_; }

private:
    void check(bgs_status st) { if (st != BGS_OK) throw Error(st, bgs_last_error(ctx_)); }
    bgs_context* ctx_ = nullptr;
};

}  // namespace bgs
