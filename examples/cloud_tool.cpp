// cloud_tool.cpp -- CPU-only helper over include/bgs_io.hpp: load a cloud file the way the reference's asset loader
// does (src/io/loader.rs:38-66, `.ply` branch) and dump the four f32 planes in the format headless.cpp's
// --dump-cloud uses (u64 n, then pos_vis n*4, sh n*48, rot n*4, scale_opacity n*4), so another host (or a test) can
// check the planes or render the identical cloud.
//
//   cloud_tool <in.ply> <out.bin>
#include <cstdio>
#include <fstream>

#include "../include/bgs_io.hpp"

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: cloud_tool <in.ply> <out.bin>\n"); return 1; }
    const std::string path = argv[1];
    if (path.size() < 4 || path.substr(path.size() - 4) != ".ply") {
        std::fprintf(stderr, "cloud_tool: only .ply is read here (.gcloud: the Python host mirror)\n");
        return 1;
    }
    try {
        std::ifstream in(path, std::ios::binary);
        if (!in) { std::fprintf(stderr, "cloud_tool: cannot open %s\n", path.c_str()); return 1; }
        const bgs::PlanarGaussian3d cloud = bgs::io::parse_ply_3d(in);
        std::ofstream f(argv[2], std::ios::binary);
        const uint64_t n = cloud.len();
        f.write((const char*)&n, 8);
        f.write((const char*)cloud.position_visibility.data(), n * 16);
        f.write((const char*)cloud.spherical_harmonic.data(), n * 192);
        f.write((const char*)cloud.rotation.data(), n * 16);
        f.write((const char*)cloud.scale_opacity.data(), n * 16);
        std::printf("%llu gaussians\n", (unsigned long long)n);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "cloud_tool: %s\n", e.what());
        return 2;
    }
    return 0;
}
