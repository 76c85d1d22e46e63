// cloud_tool.cpp -- CPU-only helper over include/bgs_io.hpp: load a cloud file the way the reference's asset loader
// does (src/io/loader.rs:38-66: `.ply` and `.gcloud`) and dump the four f32 planes in the format headless.cpp's
// --dump-cloud uses (u64 n, then pos_vis n*4, sh n*48, rot n*4, scale_opacity n*4), so another host (or a test) can
// check the planes or render the identical cloud.
//
//   cloud_tool <in.ply|in.gcloud> <out.bin>
//   cloud_tool <in.ply|in.gcloud> <out.gcloud> --gcloud
#include <cstdio>
#include <fstream>
#include <string>

#include "../include/bgs_io.hpp"

int main(int argc, char** argv) {
    if (argc != 3 && !(argc == 4 && std::string(argv[3]) == "--gcloud")) {
        std::fprintf(stderr, "usage: cloud_tool <in.ply|in.gcloud> <out.bin>            (dump the four f32 planes)\n"
                             "       cloud_tool <in.ply|in.gcloud> <out.gcloud> --gcloud  (re-encode as .gcloud)\n");
        return 1;
    }
    try {
        const bgs::PlanarGaussian3d cloud = bgs::io::load_cloud(argv[1]);      // src/io/loader.rs:38-66
        std::ofstream f(argv[2], std::ios::binary);
        const uint64_t n = cloud.len();
        if (argc == 4) {
            const std::vector<unsigned char> bytes = bgs::io::encode_gcloud(cloud);
            f.write((const char*)bytes.data(), (std::streamsize)bytes.size());
        } else {
            f.write((const char*)&n, 8);
            f.write((const char*)cloud.position_visibility.data(), n * 16);
            f.write((const char*)cloud.spherical_harmonic.data(), n * 192);
            f.write((const char*)cloud.rotation.data(), n * 16);
            f.write((const char*)cloud.scale_opacity.data(), n * 16);
        }
        std::printf("%llu gaussians\n", (unsigned long long)n);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "cloud_tool: %s\n", e.what());
        return 2;
    }
    return 0;
}
