// Wide-leaf records for scenes small enough to be ONE leaf (<= 32 pre-test units).
//
// A unit is a triangle or a PARALLELOGRAM PAIR: two triangles (a,b,c) and (d,c',b') with
// d = b + c - a -- exactly what the reference's tri_quad emits (primitives/tri_quad.cuh:13-20), i.e.
// every wall, box face and prism side.  Both halves share one plane and one affine frame
// (alpha, beta): the point is in the first half iff alpha,beta >= 0 and alpha+beta <= 1, in the second
// iff alpha,beta <= 1 and alpha+beta >= 1 -- so one conservative pre-test serves two triangles.
// Records are computed in double precision; the error budgets are folded into the stored constants.
#include "srt_host.hpp"
#include <cmath>
#include <algorithm>

namespace srt {
namespace {
struct D3 { double x, y, z; };
inline D3 sub(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 dv(const vec3f& v) { return {v.x, v.y, v.z}; }
struct Frame { D3 A, B; double aw, bw; bool ok; };
Frame frame_of(const HostTri& t) {
    const D3 a = dv(t.v[0]), e1 = sub(dv(t.v[1]), a), e2 = sub(dv(t.v[2]), a), N = cross(e1, e2);
    const double nn = dot(N, N);
    Frame f{};
    f.ok = nn > 0;
    if (!f.ok) return f;
    const D3 c1 = cross(e2, N), c2 = cross(N, e1);
    f.A = {c1.x / nn, c1.y / nn, c1.z / nn};
    f.B = {c2.x / nn, c2.y / nn, c2.z / nn};
    f.aw = -dot(f.A, a);
    f.bw = -dot(f.B, a);
    return f;
}
}  // namespace

bool build_flat_leaf(const std::vector<HostTri>& tris, const std::vector<HostMaterial>& mats, const std::vector<uint32_t>& prio,
                     double origin_l1_bound, FlatLeaf& out) {
    const int n = (int)tris.size();
    out = FlatLeaf();
    if (n == 0 || n > 64) return false;
    double radius = 0;
    for (const HostTri& t : tris)
        for (float b : t.bbox) radius = std::max(radius, (double)std::fabs(b));
    std::vector<int> partner(n, -1);
    for (int i = 0; i < n; i++) {
        if (partner[i] >= 0) continue;
        const Frame f = frame_of(tris[i]);
        if (!f.ok) continue;
        for (int j = i + 1; j < n; j++) {
            if (partner[j] >= 0) continue;
            // tri j must sit at (1,1), (0,1), (1,0) of tri i's frame (any vertex order), in tri i's plane
            bool seen[3] = {false, false, false};
            bool good = true;
            for (int k = 0; k < 3 && good; k++) {
                const D3 p = dv(tris[j].v[k]);
                const double al = dot(f.A, p) + f.aw, be = dot(f.B, p) + f.bw;
                const double dist = std::fabs(tris[i].normal.x * p.x + tris[i].normal.y * p.y + tris[i].normal.z * p.z - tris[i].D);
                int which = -1;
                if (std::fabs(al - 1) < 1e-5 && std::fabs(be - 1) < 1e-5) which = 0;
                else if (std::fabs(al) < 1e-5 && std::fabs(be - 1) < 1e-5) which = 1;
                else if (std::fabs(al - 1) < 1e-5 && std::fabs(be) < 1e-5) which = 2;
                if (which < 0 || seen[which] || dist > 1e-3 + 1e-6 * radius) good = false;
                else seen[which] = true;
            }
            if (good) { partner[i] = j; partner[j] = i; break; }
        }
    }
    int units = 0;
    for (int i = 0; i < n; i++)
        if (partner[i] < 0 || partner[i] > i) units++;
    if (units > 32) return false;
    const double O1 = origin_l1_bound;
    // Unit order = the order phase 2 runs the exact tests in.  Small things first (light, box faces), big walls last:
    // a ray that pierces an inner object and the wall behind it then finds the near hit first, and the wall's exact test
    // ends at its `t <= closest` check instead of running to the end.  The result does not depend on the order.
    std::vector<int> heads;
    for (int i = 0; i < n; i++)
        if (partner[i] < 0 || partner[i] > i) heads.push_back(i);
    auto area2 = [&](int i) {
        const D3 a = dv(tris[i].v[0]), c = cross(sub(dv(tris[i].v[1]), a), sub(dv(tris[i].v[2]), a));
        return dot(c, c);
    };
    std::stable_sort(heads.begin(), heads.end(), [&](int a, int b) { return area2(a) < area2(b); });
    for (int i : heads) {
        const HostTri& T = tris[i];
        const Frame f = frame_of(T);
        SrtFlatUnit u{};
        u.nx = T.normal.x; u.ny = T.normal.y; u.nz = T.normal.z; u.D = T.D;
        const double l1 = std::fabs(f.A.x) + std::fabs(f.A.y) + std::fabs(f.A.z) + std::fabs(f.B.x) + std::fabs(f.B.y) + std::fabs(f.B.z);
        const double n1 = std::fabs(T.normal.x) + std::fabs(T.normal.y) + std::fabs(T.normal.z);
        // |p_approx - p| <= 2e-6 (3 R + 2 |o|_1) for true hits; 16x safety; + the a_w/b_w rounding; + pair mismatch
        const double eps = 1e-4 + 16.0 * 2e-6 * (3.0 * radius + 2.0 * O1) * l1 + 4e-6 * (std::fabs(f.aw) + std::fabs(f.bw));
        const double tol = 8e-6 * (std::fabs((double)T.D) + 1e-3) + 8e-6 * n1 * O1 + (partner[i] >= 0 ? 2e-3 : 0.0);
        u.ax = (float)f.A.x; u.ay = (float)f.A.y; u.az = (float)f.A.z; u.aw = (float)(f.aw + eps);
        u.bx = (float)f.B.x; u.by = (float)f.B.y; u.bz = (float)f.B.z; u.bw = (float)(f.bw + eps);
        u.c1 = (float)(1.0 + 3.0 * eps);                          // first half:  alpha' + beta' <= c1
        u.c2 = partner[i] >= 0 ? (float)(1.0 + 2.0 * eps) : -1.f;  // second half: alpha', beta' <= c2 (never true for singles)
        u.c3 = (float)(1.0 + eps);                                // second half: alpha' + beta' >= c3
        u.tol = (float)tol;
        if (!f.ok) { u.ax = u.ay = u.az = u.bx = u.by = u.bz = 0.f; u.aw = u.bw = 0.25f; u.c1 = 1.f; }  // degenerate: always a candidate
        out.units.push_back(u);
        const int pair[2] = {i, partner[i]};
        for (int k = 0; k < 2; k++) {
            if (pair[k] < 0) {  // pad: the second slot of a single triangle repeats the first (its bit is never set)
                out.tris.push_back(out.tris.back());
                out.to_orig.push_back(out.to_orig.back());
                continue;
            }
            const HostTri& S = tris[pair[k]];
            const uint32_t mt = S.mat < mats.size() ? mats[S.mat].type : SRT_LAMBERTIAN;
            out.tris.push_back(S.pack(mt, prio[pair[k]]));
            out.to_orig.push_back((uint32_t)pair[k]);
        }
    }
    return true;
}

}  // namespace srt
