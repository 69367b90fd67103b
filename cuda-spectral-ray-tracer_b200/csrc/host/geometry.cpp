// Scene -> triangle expansion on the HOST (the reference does this in a <<<1,1>>> kernel with
// device `new`, scene/scene.cu:22-54).  Every derived quantity follows the float operation order
// of primitives/tri.cu:47-84 so the triangles are bit-identical to the host oracle's; the result
// is one flat array that uploads with a single copy.
#include "srt_host.hpp"
#include <cmath>

namespace srt {
namespace {
inline vec3f add(vec3f a, vec3f b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3f sub(vec3f a, vec3f b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3f neg(vec3f a) { return {-a.x, -a.y, -a.z}; }
inline vec3f mul(float t, vec3f v) { return {t * v.x, t * v.y, t * v.z}; }
inline vec3f divs(vec3f v, float t) { return mul(1 / t, v); }
inline float dot(vec3f u, vec3f v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
inline vec3f cross(vec3f u, vec3f v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline float len(vec3f v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }

void projection_axes(int aa_plane, int& w, int& h) {  // tri.cu:161-179
    if (aa_plane == AA_YZ) { w = 1; h = 2; }
    else if (aa_plane == AA_XZ) { w = 0; h = 2; }
    else { w = 0; h = 1; }
}
float signed_area2(int aa_plane, vec3f a, vec3f b, vec3f c) {  // tri.cu:153-181
    int w, h;
    projection_axes(aa_plane, w, h);
    return (a[w] - c[w]) * (b[h] - c[h]) - (b[w] - c[w]) * (a[h] - c[h]);
}
}  // namespace

void HostTri::derive() {  // tri::init, tri.cu:47-84
    const vec3f n = cross(sub(v[1], v[0]), sub(v[2], v[0]));
    normal = divs(n, len(n));
    const bool perp_x = std::fabs(dot(normal, vec3f(1.f, 0.f, 0.f))) < 1e-8f;
    const bool perp_y = std::fabs(dot(normal, vec3f(0.f, 1.f, 0.f))) < 1e-8f;
    const bool perp_z = std::fabs(dot(normal, vec3f(0.f, 0.f, 1.f))) < 1e-8f;
    if (perp_y && perp_z) aa_plane = AA_YZ;
    else if (perp_x && perp_z) aa_plane = AA_XZ;
    else if (perp_x && perp_y) aa_plane = AA_XY;
    // otherwise the previous value sticks (tri.cuh:30 never initialises it; fresh = NONE)
    D = dot(normal, v[0]);
    clockwise = signed_area2(aa_plane, v[0], v[1], v[2]) >= 0;
    for (int a = 0; a < 3; a++) {  // aabb(v0,v1,v2).pad(), bvh/aabb.cuh:49-57,93-102
        float lo = std::fmin(v[0][a], std::fmin(v[1][a], v[2][a]));
        float hi = std::fmax(v[0][a], std::fmax(v[1][a], v[2][a]));
        const float delta = 0.0001f;
        if (!((hi - lo) >= delta)) {
            const float padding = delta / 2;
            lo = lo - padding;
            hi = hi + padding;
        }
        bbox[2 * a] = lo;
        bbox[2 * a + 1] = hi;
    }
}

SrtTri HostTri::pack(uint32_t mat_type, uint32_t prio) const {
    int w, h;
    projection_axes(aa_plane, w, h);
    SrtTri t{};
    t.nx = normal.x; t.ny = normal.y; t.nz = normal.z; t.D = D;
    t.w0 = v[0][w]; t.h0 = v[0][h];
    t.w1 = v[1][w]; t.h1 = v[1][h];
    t.w2 = v[2][w]; t.h2 = v[2][h];
    t.bits = (mat & 0xFFFFu) | ((uint32_t)(clockwise ? 1 : 0) << 16) | ((uint32_t)w << 17) | ((uint32_t)h << 19) |
             ((mat_type & 7u) << 21);
    t.prio = prio;
    return t;
}

size_t TriangleSoup::add_tri(vec3f a, vec3f b, vec3f c, uint32_t mat, bool as_vectors) {  // tri.cuh:28-48
    HostTri t;
    t.mat = mat;
    t.v[0] = a;
    t.v[1] = as_vectors ? add(a, b) : b;
    t.v[2] = as_vectors ? add(a, c) : c;
    t.derive();
    tris.push_back(t);
    return tris.size() - 1;
}
size_t TriangleSoup::add_quad(vec3f Q, vec3f u, vec3f v, uint32_t mat) {  // tri_quad.cuh:13-20
    const size_t first = add_tri(Q, u, v, mat, true);
    add_tri(add(add(Q, u), v), neg(u), neg(v), mat, true);
    return first;
}
size_t TriangleSoup::add_box(vec3f a, vec3f b, const uint32_t m[6]) {  // tri_box.cuh:10-44, faces front/right/back/left/top/bottom
    const vec3f lo(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z));
    const vec3f hi(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z));
    const vec3f dx(hi.x - lo.x, 0.f, 0.f), dy(0, hi.y - lo.y, 0.f), dz(0, 0, hi.z - lo.z);
    const size_t first = add_quad(vec3f(lo.x, lo.y, hi.z), dx, dy, m[0]);
    add_quad(vec3f(hi.x, lo.y, hi.z), neg(dz), dy, m[1]);
    add_quad(vec3f(hi.x, lo.y, lo.z), neg(dx), dy, m[2]);
    add_quad(vec3f(lo.x, lo.y, lo.z), dz, dy, m[3]);
    add_quad(vec3f(lo.x, hi.y, hi.z), dx, neg(dz), m[4]);
    add_quad(vec3f(lo.x, lo.y, lo.z), dx, dz, m[5]);
    return first;
}
vec3f TriangleSoup::quad_center(size_t q) const {  // tri_quad.cuh:44-46
    const vec3f u = sub(tris[q].v[1], tris[q].v[0]), v = sub(tris[q].v[2], tris[q].v[0]);
    return add(divs(add(u, v), 2.0f), tris[q].v[0]);
}
vec3f TriangleSoup::box_center(size_t bx) const {  // tri_box.cuh:100-137: from the bottom (+10) and left (+6) quads
    const HostTri& bottom = tris[bx + 10];
    const HostTri& left = tris[bx + 6];
    const vec3f lo = bottom.v[0];
    const vec3f wv = sub(bottom.v[1], bottom.v[0]), hv = sub(left.v[2], left.v[0]), dv = sub(bottom.v[2], bottom.v[0]);
    const vec3f hi = add(add(add(lo, wv), hv), dv);
    return add(divs(sub(hi, lo), 2.0f), lo);
}
size_t TriangleSoup::add_pyramid(vec3f Q, vec3f u, vec3f v, vec3f w, uint32_t mat) {  // pyramid.cuh:29-47
    const size_t first = add_quad(Q, u, v, mat);
    const vec3f top = add(quad_center(first), w);
    const vec3f v1 = add(Q, u), v2 = add(Q, v), v3 = add(v2, u);
    add_tri(Q, top, v2, mat, false);
    add_tri(v1, top, Q, mat, false);
    add_tri(v2, top, v3, mat, false);
    add_tri(v3, top, v1, mat, false);
    return first;
}
size_t TriangleSoup::add_prism(vec3f Q, vec3f u, vec3f v, vec3f w, uint32_t mat) {  // prism.cuh:22-32
    const size_t first = add_tri(Q, v, u, mat, true);
    add_tri(add(Q, w), u, v, mat, true);
    add_quad(Q, u, w, mat);
    add_quad(Q, w, v, mat);
    add_quad(add(Q, u), sub(v, u), w, mat);
    return first;
}
vec3f TriangleSoup::prism_centroid(size_t p) const {  // prism.cuh:44-54
    vec3f s = add(tris[p].v[0], tris[p].v[1]);
    s = add(s, tris[p].v[2]);
    s = add(s, tris[p + 1].v[0]);
    s = add(s, tris[p + 1].v[1]);
    s = add(s, tris[p + 1].v[2]);
    return divs(s, 6.f);
}
void TriangleSoup::translate(size_t first, size_t count, vec3f d, bool rederive_after) {  // tri.cu:86-94
    for (size_t i = first; i < first + count; i++) {
        for (auto& p : tris[i].v) p = add(p, d);
        if (rederive_after) tris[i].derive();
    }
}
void TriangleSoup::rederive(size_t first, size_t count) {
    for (size_t i = first; i < first + count; i++) tris[i].derive();
}
void TriangleSoup::rotate_y_about(size_t first, size_t count, vec3f pivot, float theta) {
    // local rotation = translate(-pivot), R_y(theta), translate(+pivot), no re-derivation
    // (tri_box.cu:14-35, pyramid.cu:14-36, prism.cu:14-35; matrix from transform.cu:18-23)
    const float c = std::cos(theta), s = std::sin(theta);
    const float m[9] = {c, 0.f, s, 0.f, 1.0f, 0.f, -s, 0.f, c};
    translate(first, count, neg(pivot), false);
    for (size_t i = first; i < first + count; i++)
        for (auto& p : tris[i].v)
            p = vec3f((m[0] * p.x) + (m[1] * p.y) + (m[2] * p.z), (m[3] * p.x) + (m[4] * p.y) + (m[5] * p.z),
                      (m[6] * p.x) + (m[7] * p.y) + (m[8] * p.z));
    translate(first, count, pivot, false);
}

}  // namespace srt
