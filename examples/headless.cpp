// headless.cpp -- C++ host example over include/bgs.hpp, the counterpart of the reference's
// examples/headless.rs (offscreen target, camera at (0, 1.5, 5), random cloud, one frame written to disk).
//
//   headless [count=100000] [width=1920] [height=1080] [global_scale=1.0] [out=headless_output/0.ppm]
//            [--dump-cloud file] [--raw file]
// --dump-cloud writes the four f32 planes (n, then pos_vis, sh, rot, scale_opacity) so another host can
// render the identical cloud; --raw writes the RGBA8 frame bytes.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <sys/stat.h>

#include "../include/bgs.hpp"

int main(int argc, char** argv) {
    size_t count = 100000; int W = 1920, H = 1080; float scale = 1.0f;
    std::string out = "headless_output/0.ppm", dump, raw;
    int pos = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--dump-cloud" && i + 1 < argc) { dump = argv[++i]; continue; }
        if (a == "--raw" && i + 1 < argc) { raw = argv[++i]; continue; }
        switch (pos++) {
            case 0: count = std::strtoull(a.c_str(), nullptr, 10); break;
            case 1: W = std::atoi(a.c_str()); break;
            case 2: H = std::atoi(a.c_str()); break;
            case 3: scale = (float)std::atof(a.c_str()); break;
            case 4: out = a; break;
        }
    }
    try {
        bgs::GaussianSplattingPlugin plugin(0);
        const bgs::PlanarGaussian3d cloud = bgs::random_gaussians_3d_seeded(count, 0);
        bgs::PlanarGaussian3dHandle handle = plugin.add_cloud(cloud);
        bgs::CloudSettings settings;
        settings.global_scale = scale;
        const bgs_view view = bgs::headless_view(W, H);
        std::vector<unsigned char> frame((size_t)W * H * 4);
        bool drawn = false;
        for (int f = 0; f < 3; ++f) drawn = plugin.render_view(handle, settings, view, frame.data(), BGS_FORMAT_RGBA8_SRGB);
        const bgs_frame_stats fs = plugin.frame_stats();
        float us[6];
        bgs_stage_times_us(plugin.context(), us);
        std::printf("rendered=%d n=%u visible=%u pairs=%llu frame=%.1f us\n", (int)drawn, fs.n, fs.n_visible,
                    (unsigned long long)fs.n_pairs, us[5]);
        if (!dump.empty()) {
            std::ofstream f(dump, std::ios::binary);
            const uint64_t n = cloud.len();
            f.write((const char*)&n, 8);
            f.write((const char*)cloud.position_visibility.data(), n * 16);
            f.write((const char*)cloud.spherical_harmonic.data(), n * 192);
            f.write((const char*)cloud.rotation.data(), n * 16);
            f.write((const char*)cloud.scale_opacity.data(), n * 16);
        }
        if (!raw.empty()) { std::ofstream f(raw, std::ios::binary); f.write((const char*)frame.data(), frame.size()); }
        const size_t slash = out.find_last_of('/');
        if (slash != std::string::npos) mkdir(out.substr(0, slash).c_str(), 0755);
        std::ofstream f(out, std::ios::binary);
        f << "P6\n" << W << " " << H << "\n255\n";
        for (size_t i = 0; i < (size_t)W * H; ++i) f.write((const char*)&frame[4 * i], 3);
        std::printf("wrote %s\n", out.c_str());
    } catch (const bgs::Error& e) {
        std::fprintf(stderr, "headless: %s (status %d)\n", e.what(), (int)e.status);
        return 2;
    }
    return 0;
}
