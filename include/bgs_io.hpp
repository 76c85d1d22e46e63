// bgs_io.hpp -- C++ host mirror of the reference's cloud loader for the splat path's input row,
// header-only, no dependency on the C ABI: it produces the four planes `bgs::PlanarGaussian3d` (bgs.hpp) uploads.
//
//   parse_ply_3d     src/io/ply.rs:23-132   INRIA 3DGS `.ply` -> PlanarGaussian3d, with the reference's quirks:
//     * only `float` properties of the `vertex` element are consumed (ply.rs:33-90: `Property::Float`)
//     * opacity = sigmoid(raw)                                          ply.rs:40-42
//     * f_rest_i -> channel i / 16 (not / 15), coefficient (i % 15) + 1, interleaved index coefficient * 3 + channel,
//       dropped when >= 48; later properties overwrite earlier ones       ply.rs:49-69
//     * scale_i = exp(clamp(raw_i, mean(raw) -+ 4))                      ply.rs:103-116
//     * rotation normalised                                              ply.rs:118-124
//     * padded with default gaussians by 32 - (n % 32) entries           ply.rs:127-129
//   (`.gcloud`, the FlexBuffers serde form, is read by the Python host mirror: bevy_gaussian_splatting_b200/gcloud.py.)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <istream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "bgs.hpp"

namespace bgs {
namespace io {

namespace detail {
struct PlyProp { std::string name; int size; bool is_f32; bool is_f64; int list_count_size = 0; };   // list: size = item size
struct PlyElement { std::string name; size_t count = 0; std::vector<PlyProp> props; };

inline int ply_type_size(const std::string& t, bool& f32, bool& f64) {
    f32 = (t == "float" || t == "float32"); f64 = (t == "double" || t == "float64");
    if (f32) return 4;
    if (f64) return 8;
    if (t == "uchar" || t == "uint8" || t == "char" || t == "int8") return 1;
    if (t == "short" || t == "int16" || t == "ushort" || t == "uint16") return 2;
    if (t == "int" || t == "int32" || t == "uint" || t == "uint32") return 4;
    throw std::runtime_error("ply: unknown property type " + t);
}
inline float load_f32(const unsigned char* p, bool big_endian) {
    unsigned char b[4];
    if (big_endian) { b[0] = p[3]; b[1] = p[2]; b[2] = p[1]; b[3] = p[0]; } else std::memcpy(b, p, 4);
    float v; std::memcpy(&v, b, 4); return v;
}
}  // namespace detail

inline PlanarGaussian3d parse_ply_3d(std::istream& in) {
    using namespace detail;
    std::string line;
    if (!std::getline(in, line) || line.substr(0, 3) != "ply") throw std::runtime_error("not a PLY file");
    std::string format;
    std::vector<PlyElement> elements;
    bool ended = false;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok; ls >> tok;
        if (tok == "format") ls >> format;
        else if (tok == "element") { PlyElement e; ls >> e.name >> e.count; elements.push_back(e); }
        else if (tok == "property") {
            std::string ty, name; ls >> ty;
            if (elements.empty()) throw std::runtime_error("ply: property before element");
            PlyProp p;
            if (ty == "list") {   // fine in other elements (faces), which are skipped; not in the vertex element
                if (elements.back().name == "vertex") throw std::runtime_error("ply: list properties are not supported in the vertex element");
                std::string cty, ity; ls >> cty >> ity >> name;
                bool a, b;
                p.list_count_size = ply_type_size(cty, a, b);
                p.size = ply_type_size(ity, a, b); p.is_f32 = p.is_f64 = false;
            } else {
                ls >> name;
                p.size = ply_type_size(ty, p.is_f32, p.is_f64);
            }
            p.name = name;
            elements.back().props.push_back(p);
        } else if (tok == "end_header") { ended = true; break; }
    }
    if (!ended) throw std::runtime_error("unterminated PLY header");
    const bool ascii = format == "ascii", big = format == "binary_big_endian";
    if (!ascii && !big && format != "binary_little_endian") throw std::runtime_error("ply: unknown format " + format);

    // column index of every consumed property inside the vertex rows (header order = the order set_property is called in)
    size_t n = 0;
    std::vector<std::vector<float>> cols;       // one vector per float property of the vertex element
    std::vector<std::string> names;
    bool have_vertex = false;
    for (const PlyElement& el : elements) {
        size_t stride = 0;
        for (const PlyProp& p : el.props) stride += (size_t)p.size;
        if (el.name != "vertex") {   // skip
            bool lists = false;
            for (const PlyProp& p : el.props) lists = lists || p.list_count_size != 0;
            if (ascii) {
                in >> std::ws;
                for (size_t i = 0; i < el.count; ++i) std::getline(in, line);
            } else if (!lists) {
                in.ignore((std::streamsize)(stride * el.count));
            } else {
                for (size_t i = 0; i < el.count; ++i)
                    for (const PlyProp& p : el.props) {
                        if (p.list_count_size == 0) { in.ignore(p.size); continue; }
                        unsigned char cb[8] = {0};
                        in.read((char*)cb, p.list_count_size);
                        uint64_t cnt = 0;
                        for (int k = 0; k < p.list_count_size; ++k)
                            cnt |= (uint64_t)cb[big ? p.list_count_size - 1 - k : k] << (8 * k);
                        in.ignore((std::streamsize)(cnt * (uint64_t)p.size));
                    }
            }
            continue;
        }
        static const char* required[] = {"x", "y", "z", "f_dc_0", "f_dc_1", "f_dc_2", "scale_0", "scale_1", "opacity",
                                         "rot_0", "rot_1", "rot_2", "rot_3"};
        for (const char* r : required) {
            bool found = false;
            for (const PlyProp& p : el.props) found = found || p.name == r;
            if (!found) throw std::runtime_error("missing required properties");   // ply.rs:92-97
        }
        have_vertex = true; n = el.count;
        for (const PlyProp& p : el.props)
            if (ascii || p.is_f32) { names.push_back(p.name); cols.emplace_back(n); }
        if (ascii) {
            for (size_t i = 0; i < n; ++i) {
                size_t c = 0;
                for (size_t k = 0; k < el.props.size(); ++k) { double v; in >> v; cols[c++][i] = (float)v; }
            }
        } else {
            std::vector<unsigned char> row(stride);
            for (size_t i = 0; i < n; ++i) {
                in.read((char*)row.data(), (std::streamsize)stride);
                if (!in) throw std::runtime_error("ply: truncated vertex data");
                size_t off = 0, c = 0;
                for (const PlyProp& p : el.props) {
                    if (p.is_f32) cols[c++][i] = load_f32(row.data() + off, big);
                    off += (size_t)p.size;
                }
            }
        }
    }
    PlanarGaussian3d out;
    if (!have_vertex) return out;
    const size_t pad = 32 - (n % 32), total = n + pad;
    out.position_visibility.assign(total * 4, 0.0f);
    out.spherical_harmonic.assign(total * 48, 0.0f);
    out.rotation.assign(total * 4, 0.0f);
    out.scale_opacity.assign(total * 4, 0.0f);
    for (size_t i = 0; i < total; ++i) out.position_visibility[4 * i + 3] = 1.0f;   // PositionVisibility::default
    for (size_t c = 0; c < names.size(); ++c) {
        const std::string& key = names[c];
        const std::vector<float>& v = cols[c];
        auto put = [&](std::vector<float>& plane, size_t stride, size_t at) { for (size_t i = 0; i < n; ++i) plane[stride * i + at] = v[i]; };
        if (key == "x") put(out.position_visibility, 4, 0);
        else if (key == "y") put(out.position_visibility, 4, 1);
        else if (key == "z") put(out.position_visibility, 4, 2);
        else if (key == "visibility") put(out.position_visibility, 4, 3);
        else if (key == "f_dc_0") put(out.spherical_harmonic, 48, 0);
        else if (key == "f_dc_1") put(out.spherical_harmonic, 48, 1);
        else if (key == "f_dc_2") put(out.spherical_harmonic, 48, 2);
        else if (key == "scale_0") put(out.scale_opacity, 4, 0);
        else if (key == "scale_1") put(out.scale_opacity, 4, 1);
        else if (key == "scale_2") put(out.scale_opacity, 4, 2);
        else if (key == "opacity") { for (size_t i = 0; i < n; ++i) out.scale_opacity[4 * i + 3] = 1.0f / (1.0f + std::exp(-v[i])); }
        else if (key == "rot_0") put(out.rotation, 4, 0);
        else if (key == "rot_1") put(out.rotation, 4, 1);
        else if (key == "rot_2") put(out.rotation, 4, 2);
        else if (key == "rot_3") put(out.rotation, 4, 3);
        else if (key.compare(0, 7, "f_rest_") == 0) {
            // the reference does `parse::<usize>().unwrap()` (io/ply.rs): anything but decimal digits is a hard error
            const char* dig = key.c_str() + 7;
            if (*dig == 0 || std::strlen(dig) > 6) throw std::runtime_error("ply: bad property name " + key);
            for (const char* q = dig; *q; ++q)
                if (*q < '0' || *q > '9') throw std::runtime_error("ply: bad property name " + key);
            const unsigned long i = std::strtoul(dig, nullptr, 10);
            const unsigned long channel = i / 16, coefficient = (i % 15) + 1, idx = coefficient * 3 + channel;
            if (idx < 48) put(out.spherical_harmonic, 48, (size_t)idx);
        }
    }
    for (size_t i = 0; i < n; ++i) {
        float* so = &out.scale_opacity[4 * i];
        const float mean = ((so[0] + so[1]) + so[2]) / 3.0f;
        for (int k = 0; k < 3; ++k) so[k] = std::exp(std::fmin(std::fmax(so[k], mean - 4.0f), mean + 4.0f));
        float* q = &out.rotation[4 * i];
        const float norm = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
        for (int k = 0; k < 4; ++k) q[k] = q[k] / norm;
    }
    return out;
}

}  // namespace io
}  // namespace bgs
