// The three hard-coded reference scenes (scene/scene.cu:73-257), the synthetic soup of
// BASELINE.json configs[3], and a minimal OBJ reader.
#include "srt_host.hpp"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace srt {
namespace {
constexpr float kPi = 3.1415926535897932385f;
inline float radians(float deg) { return deg * kPi / 180.0f; }

// refraction/sellmeier.cuh:6-13
const float kBK7_b[3] = {1.03961212f, 0.231792344f, 1.01046945f};
const float kBK7_c[3] = {6.00069867e-3f, 2.00179144e-2f, 1.03560653e2f};
const float kSilica_b[3] = {0.6961663f, 0.4079426f, 0.8974794f};  // no reference scene uses fused silica; kept for srt_glass_coefficients
const float kSilica_c[3] = {0.0684043f, 0.1162414f, 9.896161f};
const float kFlint_b[3] = {1.34533359f, 0.209073176f, 0.937357162f};
const float kFlint_c[3] = {0.00997743871f, 0.0470450767f, 111.886764f};

HostMaterial lambertian(float r, float g, float b) {
    srt_material_desc d{};
    d.type = SRT_MAT_LAMBERTIAN; d.color[0] = r; d.color[1] = g; d.color[2] = b; d.fuzz = 1.0f;
    return HostMaterial::from_desc(d, true);
}
HostMaterial metallic(float r, float g, float b, float fuzz) {
    srt_material_desc d{};
    d.type = SRT_MAT_METALLIC; d.color[0] = r; d.color[1] = g; d.color[2] = b; d.fuzz = fuzz;
    return HostMaterial::from_desc(d, true);
}
HostMaterial emissive(float r, float g, float b, float power) {
    srt_material_desc d{};
    d.type = SRT_MAT_EMISSIVE; d.color[0] = r; d.color[1] = g; d.color[2] = b; d.fuzz = 1.0f; d.emission_power = power;
    return HostMaterial::from_desc(d, true);
}
HostMaterial dielectric(const float b[3], const float c[3], bool ref_compat) {
    srt_material_desc d{};
    d.type = SRT_MAT_DIELECTRIC; d.color[0] = d.color[1] = d.color[2] = 1.0f; d.fuzz = 1.0f;
    for (int i = 0; i < 3; i++) { d.sellmeier_b[i] = b[i]; d.sellmeier_c[i] = c[i]; }
    return HostMaterial::from_desc(d, ref_compat);
}

}  // namespace
bool glass_coefficients(int which, float b[3], float c[3]) {  // refraction/sellmeier.cuh:6-13
    const float *sb, *sc;
    switch (which) {
        case SRT_GLASS_BK7: sb = kBK7_b; sc = kBK7_c; break;
        case SRT_GLASS_FUSED_SILICA: sb = kSilica_b; sc = kSilica_c; break;
        case SRT_GLASS_FLINT: sb = kFlint_b; sc = kFlint_c; break;
        default: return false;
    }
    for (int i = 0; i < 3; i++) { b[i] = sb[i]; c[i] = sc[i]; }
    return true;
}
namespace {

// five walls + ceiling light, written to fixed triangle slots 0..11 (scene.cu:83-102)
void room(TriangleSoup& g, const uint32_t wall[5], uint32_t light) {
    g.add_quad(vec3f(0, 0, 0), vec3f(0, 0, 555), vec3f(555, 0, 0), wall[0]);          // floor   -> 0,1
    g.add_quad(vec3f(555, 555, 555), vec3f(-555, 0, 0), vec3f(0, 0, -555), wall[2]);  // ceiling -> 2,3
    g.add_quad(vec3f(0, 0, 555.f), vec3f(0, 555, 0), vec3f(555, 0, 0), wall[1]);      // back    -> 4,5
    g.add_quad(vec3f(555, 0, 0), vec3f(0, 0, 555), vec3f(0, 555, 0), wall[3]);        // x = 555 -> 6,7
    g.add_quad(vec3f(0, 0, 0), vec3f(0, 555, 0), vec3f(0, 0, 555), wall[4]);          // x = 0   -> 8,9
    const vec3f c(555.f / 2.f, 554.f, 555.f / 2.f);
    const float width = 100.f, depth = 100.f;
    g.add_quad(vec3f(c.x + width / 2.f, c.y, c.z + depth / 2.f), vec3f(-width, 0, 0), vec3f(0, 0, -depth), light);  // 10,11
}
// two boxes and the glass pyramid (scene.cu:114-128)
void furniture(TriangleSoup& g, const uint32_t box1[6], const uint32_t box2[6], uint32_t pyramid_mat) {
    size_t b = g.add_box(vec3f(0.f, 0.f, 0.f), vec3f(165.f, 330.f, 165.f), box1);
    g.rotate_y_about(b, 12, g.box_center(b), radians(25.f));
    g.translate(b, 12, vec3f(265.f, 0.f, 295.f), true);
    b = g.add_box(vec3f(0.f, 0.f, 0.f), vec3f(165.f, 165.f, 165.f), box2);
    g.rotate_y_about(b, 12, g.box_center(b), radians(-18.f));
    g.translate(b, 12, vec3f(130.f, 0.f, 65.f), true);
    const size_t p = g.add_pyramid(vec3f(165.f, 166.f, 0.f), vec3f(-165.f, 0.f, 0.f), vec3f(0.f, 0.f, 165.f), vec3f(0.f, 165.f, 0.f), pyramid_mat);
    g.rotate_y_about(p, 6, g.quad_center(p), radians(-18.f));
    g.translate(p, 6, vec3f(130.f, 0.f, 65.f), true);
}
CameraBuilder reference_camera() {  // scene.cu:259-320 (identical for the three scenes)
    CameraBuilder cb;
    cb.vfov = 40.0f;
    cb.lookfrom = vec3f(278, 278, -800);
    cb.lookat = vec3f(278, 278, 0);
    cb.vup = vec3f(0, 1, 0);
    cb.defocus_angle = 0.0f;
    cb.focus_dist = 10.0f;
    cb.background = vec3f(0.0f, 0.0f, 0.0f);
    return cb;
}
}  // namespace

SceneDescription make_reference_scene(unsigned id, bool ref_compat) {
    SceneDescription s;
    TriangleSoup g;
    s.camera = reference_camera();
    if (id == 1) {  // Prism World, scene.cu:132-173
        s.mats = {lambertian(.73f, .73f, .73f), emissive(1, 1, 1, 5), dielectric(kFlint_b, kFlint_c, ref_compat)};
        const uint32_t wall[5] = {0, 0, 0, 0, 0};
        room(g, wall, 1);
        const vec3f c(555.f / 2.f, 554.f, 555.f / 2.f);
        const float width = 100.f, prism_width = 165.f, prism_height = 200.f;
        const size_t p = g.add_prism(vec3f(c.x - width / 2.f, c.y - 1.f, c.z - prism_height / 2.f), vec3f(0.f, -prism_width, 0.f),
                                     vec3f((prism_width * std::sqrt(3.f)) / 2.f, -prism_width / 2.f, 0.f), vec3f(0.f, 0.f, 200.f), 2);
        g.rotate_y_about(p, 8, g.prism_centroid(p), radians(10.f));
        g.rederive(p, 8);
    } else if (id == 2) {  // Different Materials, scene.cu:175-226
        s.mats = {lambertian(.65f, .05f, .05f), lambertian(.12f, .45f, .15f), dielectric(kFlint_b, kFlint_c, ref_compat),
                  lambertian(.73f, .73f, .73f), emissive(1.f, 1.f, 1.f, 5.f), metallic(.5f, .5f, .5f, 0.3f),
                  lambertian(.12f, .15f, .45f), dielectric(kBK7_b, kBK7_c, ref_compat), metallic(.7f, .7f, .7f, 0.8f)};
        const uint32_t wall[5] = {6, 1, 2, 8, 5};
        room(g, wall, 4);
        const uint32_t b1[6] = {3, 8, 0, 1, 2, 3}, b2[6] = {7, 6, 8, 7, 1, 2};
        furniture(g, b1, b2, 2);
    } else {  // Cornell Box, scene.cu:73-130 (also the reference's default: branch)
        s.mats = {lambertian(.65f, .05f, .05f), lambertian(.12f, .45f, .15f), dielectric(kFlint_b, kFlint_c, ref_compat),
                  lambertian(.73f, .73f, .73f), emissive(1.f, 1.f, 1.f, 5.f), metallic(.5f, .5f, .5f, 0.3f),
                  lambertian(.12f, .15f, .45f)};
        const uint32_t wall[5] = {3, 3, 3, 1, 6};
        room(g, wall, 4);
        const uint32_t b1[6] = {5, 5, 5, 5, 5, 5}, b2[6] = {0, 0, 0, 0, 0, 0};
        furniture(g, b1, b2, 2);
    }
    s.tris = std::move(g.tris);
    return s;
}

namespace {
inline uint64_t splitmix64(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline float uniform24(uint64_t& x) { return (float)(splitmix64(x) >> 40) * (1.0f / 16777216.0f); }
}  // namespace

SceneDescription make_soup_scene(uint32_t n, uint64_t seed) {
    // centre ~ U[0,1)^3 * 555, vertices = centre + U(-s,s)^3 with s = 555 n^(-1/3);
    // grey lambertian soup + one emissive 100x100 quad under the ceiling (last two triangles).
    SceneDescription s;
    s.camera = reference_camera();
    s.mats = {lambertian(.73f, .73f, .73f), emissive(1.f, 1.f, 1.f, 5.f)};
    TriangleSoup g;
    g.tris.reserve(n);
    const float sz = 555.0f * std::pow((float)n, -1.0f / 3.0f);
    uint64_t st = seed;
    const uint32_t nsoup = n >= 3 ? n - 2 : n;  // tiny soups: no light quad
    for (uint32_t i = 0; i < nsoup; i++) {
        float c[3], p[9];
        for (int k = 0; k < 3; k++) c[k] = uniform24(st) * 555.0f;
        for (int k = 0; k < 9; k++) p[k] = c[k % 3] + (uniform24(st) * 2.0f - 1.0f) * sz;
        g.add_tri(vec3f(p[0], p[1], p[2]), vec3f(p[3], p[4], p[5]), vec3f(p[6], p[7], p[8]), 0, false);
    }
    if (n >= 3) {
        const vec3f c(555.f / 2.f, 554.f, 555.f / 2.f);
        g.add_quad(vec3f(c.x + 50.f, c.y, c.z + 50.f), vec3f(-100.f, 0, 0), vec3f(0, 0, -100.f), 1);
    }
    s.tris = std::move(g.tris);
    return s;
}

bool load_obj(const char* path, std::vector<float>& verts9) {
    std::ifstream in(path);
    if (!in) { set_error(std::string("cannot open OBJ file ") + path); return false; }
    std::vector<float> pos;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        if (tag == "v") {
            float x, y, z;
            if (ss >> x >> y >> z) { pos.push_back(x); pos.push_back(y); pos.push_back(z); }
        } else if (tag == "f") {
            std::vector<long> idx;
            std::string tok;
            while (ss >> tok) {
                long v = std::strtol(tok.c_str(), nullptr, 10);  // "v", "v/vt", "v//vn", "v/vt/vn"
                const long nv = (long)(pos.size() / 3);
                if (v < 0) v = nv + v + 1;
                if (v < 1 || v > nv) { set_error("OBJ face references a missing vertex"); return false; }
                idx.push_back(v - 1);
            }
            for (size_t k = 2; k < idx.size(); k++) {
                const long tri[3] = {idx[0], idx[k - 1], idx[k]};
                for (long v : tri)
                    for (int c = 0; c < 3; c++) verts9.push_back(pos[3 * v + c]);
            }
        }
    }
    return true;
}

// Stanford PLY: ascii or binary_little_endian; element vertex with x y z (any scalar type, extra
// properties skipped), element face with one list property (fan triangulation); other elements skipped
// when they come after the faces, refused when they come before (their size may be unknown).
namespace {
int ply_type_size(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
double ply_read_scalar(std::istream& in, const std::string& t, bool ascii) {
    if (ascii) { double v = 0; in >> v; return v; }
    unsigned char b[8] = {0};
    in.read(reinterpret_cast<char*>(b), ply_type_size(t));
    if (t == "char" || t == "int8") return (double)(signed char)b[0];
    if (t == "uchar" || t == "uint8") return (double)b[0];
    if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, b, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, b, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, b, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; std::memcpy(&v, b, 4); return v; }
    if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); return v; }
    double v; std::memcpy(&v, b, 8); return v;
}
}  // namespace

bool load_ply(const char* path, std::vector<float>& verts9) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { set_error(std::string("cannot open PLY file ") + path); return false; }
    std::string line;
    if (!std::getline(in, line) || line.compare(0, 3, "ply") != 0) { set_error("not a PLY file"); return false; }
    struct Prop { std::string name, type, count_type; bool is_list; };
    struct Elem { std::string name; size_t count; std::vector<Prop> props; };
    std::vector<Elem> elems;
    bool ascii = true, header_done = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        if (tag == "format") {
            std::string f;
            ss >> f;
            if (f == "ascii") ascii = true;
            else if (f == "binary_little_endian") ascii = false;
            else { set_error("PLY format " + f + " is not supported"); return false; }
        } else if (tag == "element") {
            Elem e;
            ss >> e.name >> e.count;
            elems.push_back(e);
        } else if (tag == "property" && !elems.empty()) {
            Prop pr;
            std::string t;
            ss >> t;
            pr.is_list = t == "list";
            if (pr.is_list) ss >> pr.count_type >> pr.type >> pr.name;
            else { pr.type = t; ss >> pr.name; }
            if (!ply_type_size(pr.type) || (pr.is_list && !ply_type_size(pr.count_type))) { set_error("PLY property type not understood"); return false; }
            elems.back().props.push_back(pr);
        } else if (tag == "end_header") { header_done = true; break; }
    }
    if (!header_done) { set_error("PLY header has no end_header"); return false; }
    std::vector<float> pos;
    bool have_faces = false;
    for (const Elem& e : elems) {
        if (e.name == "vertex") {
            int ix = -1, iy = -1, iz = -1;
            for (size_t k = 0; k < e.props.size(); k++) {
                if (e.props[k].is_list) { set_error("PLY vertex list properties are not supported"); return false; }
                if (e.props[k].name == "x") ix = (int)k;
                if (e.props[k].name == "y") iy = (int)k;
                if (e.props[k].name == "z") iz = (int)k;
            }
            if (ix < 0 || iy < 0 || iz < 0) { set_error("PLY vertex element lacks x/y/z"); return false; }
            pos.resize(3 * e.count);
            for (size_t v = 0; v < e.count; v++)
                for (size_t k = 0; k < e.props.size(); k++) {
                    const double val = ply_read_scalar(in, e.props[k].type, ascii);
                    if ((int)k == ix) pos[3 * v] = (float)val;
                    if ((int)k == iy) pos[3 * v + 1] = (float)val;
                    if ((int)k == iz) pos[3 * v + 2] = (float)val;
                }
        } else if (e.name == "face") {
            const long nv = (long)(pos.size() / 3);
            for (size_t f = 0; f < e.count; f++)
                for (const Prop& pr : e.props) {
                    if (!pr.is_list) { ply_read_scalar(in, pr.type, ascii); continue; }
                    const long cnt = (long)ply_read_scalar(in, pr.count_type, ascii);
                    std::vector<long> idx;
                    for (long k = 0; k < cnt; k++) idx.push_back((long)ply_read_scalar(in, pr.type, ascii));
                    if (pr.name != "vertex_indices" && pr.name != "vertex_index") continue;
                    for (long v : idx)
                        if (v < 0 || v >= nv) { set_error("PLY face references a missing vertex"); return false; }
                    for (size_t k = 2; k < idx.size(); k++) {
                        const long tri[3] = {idx[0], idx[k - 1], idx[k]};
                        for (long v : tri)
                            for (int c = 0; c < 3; c++) verts9.push_back(pos[3 * v + c]);
                    }
                }
            have_faces = true;
        } else if (!have_faces) {
            set_error("PLY element '" + e.name + "' before the faces is not supported");
            return false;
        }
        if (!in) { set_error("PLY file is truncated"); return false; }
    }
    if (!have_faces) { set_error("PLY file has no face element"); return false; }
    return true;
}

}  // namespace srt
