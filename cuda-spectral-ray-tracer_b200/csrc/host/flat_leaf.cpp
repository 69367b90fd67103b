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
    // Two triangles become one unit only when they are coplanar TO ROUNDING: every partner vertex within a few float
    // ulps of the scene radius of the head's plane, and the two stored normals parallel.  (The pre-test intersects
    // the ray with the head's plane only; a partner that is really tilted against it would be tested at a point that
    // drifts by distance x tan(incidence) and must stay a unit of its own.)  What distance remains is measured here
    // and charged to the unit's error budget for incidence angles up to atan(1000).
    const double coplanar_tol = 4.0 * 1.1920929e-7 * std::max(radius, 1.0);
    std::vector<int> partner(n, -1);
    std::vector<double> pair_dist(n, 0.0);
    for (int i = 0; i < n; i++) {
        if (partner[i] >= 0) continue;
        const Frame f = frame_of(tris[i]);
        if (!f.ok) continue;
        for (int j = i + 1; j < n; j++) {
            if (partner[j] >= 0) continue;
            // tri j must sit at (1,1), (0,1), (1,0) of tri i's frame (any vertex order), in tri i's plane
            bool seen[3] = {false, false, false};
            const double nn = (double)tris[i].normal.x * tris[j].normal.x + (double)tris[i].normal.y * tris[j].normal.y + (double)tris[i].normal.z * tris[j].normal.z;
            bool good = std::fabs(std::fabs(nn) - 1.0) < 1e-6;
            double worst = 0;
            for (int k = 0; k < 3 && good; k++) {
                const D3 p = dv(tris[j].v[k]);
                const double al = dot(f.A, p) + f.aw, be = dot(f.B, p) + f.bw;
                const double dist = std::fabs(tris[i].normal.x * p.x + tris[i].normal.y * p.y + tris[i].normal.z * p.z - tris[i].D);
                int which = -1;
                if (std::fabs(al - 1) < 1e-5 && std::fabs(be - 1) < 1e-5) which = 0;
                else if (std::fabs(al) < 1e-5 && std::fabs(be - 1) < 1e-5) which = 1;
                else if (std::fabs(al - 1) < 1e-5 && std::fabs(be) < 1e-5) which = 2;
                if (which < 0 || seen[which] || dist > coplanar_tol) good = false;
                else { seen[which] = true; worst = std::max(worst, dist); }
            }
            if (good) { partner[i] = j; partner[j] = i; pair_dist[i] = pair_dist[j] = worst; break; }
        }
    }
    int units = 0;
    for (int i = 0; i < n; i++)
        if (partner[i] < 0 || partner[i] > i) units++;
    if (units > 32) return false;
    const double O1 = origin_l1_bound;
    // Error budget of the pre-test (u = 2^-24, R2 = sqrt(3) R, true hits lie within |o| + R2 of the origin).
    //  * Head triangle of a unit, single triangles, and partners whose stored plane (n, D) equals the head's bit for bit
    //    (every axis-aligned quad): the kernel evaluates numerator and denominator of t with the very expressions of the
    //    exact test, so the two t differ by the approximate reciprocal only, 2^-22 relative: the tested point is off by
    //    <= 2^-21 (O1 + R2) + a few ulps of the coordinates, at any incidence angle.  Slack: eps_exact world units.
    //  * Partners whose plane differs in the last bits (faces of a rotated box: each triangle derived its own normal):
    //    their exact test rounds D' - n'.o and n'.d on its own, so the point moves by
    //        [ 12u (R2 + O1) + |D - D'| + |n - n'|_inf O1 + |n - n'|_2 (O1 + R2) ] / cos(incidence).
    //    Slack eps_pair; a ray flatter than cos(incidence) < guard = 1.5 * that bracket / eps_pair can only hit inside the
    //    scene when it starts within near = guard (O1 + R2) + bracket of the plane -- then the partner stays a candidate
    //    without a verdict (trace_impl.cuh, flat_unit_test).  Units with identical planes carry near = 0: never unsure.
    const double u24 = 5.9604645e-8, R2 = std::sqrt(3.0) * radius;
    const double eps_exact = 16.0 * 2e-6 * (3.0 * radius + 2.0 * O1);
    const double eps_pair = 4.0 * eps_exact;
    std::vector<double> mismatch(n, 0.0);  // per head: the bracket's plane-difference terms, 0 for bit-identical planes
    double worst_bracket = 0;
    for (int i = 0; i < n; i++) {
        const int j = partner[i];
        if (j < i) continue;  // singles (-1) and second halves
        const HostTri &H = tris[i], &Pn = tris[j];
        const float sgn = ((double)H.normal.x * Pn.normal.x + (double)H.normal.y * Pn.normal.y + (double)H.normal.z * Pn.normal.z) < 0 ? -1.0f : 1.0f;
        const bool same = H.normal.x == sgn * Pn.normal.x && H.normal.y == sgn * Pn.normal.y && H.normal.z == sgn * Pn.normal.z && H.D == sgn * Pn.D;
        if (same) continue;
        const double dx = std::fabs((double)H.normal.x - sgn * (double)Pn.normal.x), dy = std::fabs((double)H.normal.y - sgn * (double)Pn.normal.y),
                     dz = std::fabs((double)H.normal.z - sgn * (double)Pn.normal.z);
        const double bracket = 12.0 * u24 * (R2 + O1) + std::fabs((double)H.D - sgn * (double)Pn.D) + std::max({dx, dy, dz}) * O1 +
                               std::sqrt(dx * dx + dy * dy + dz * dz) * (O1 + R2);
        mismatch[i] = bracket;
        worst_bracket = std::max(worst_bracket, bracket);
    }
    out.guard = (float)(1.5 * worst_bracket / eps_pair);
    const double near_dist = 1.5 * worst_bracket / eps_pair * (O1 + R2) + worst_bracket;
    out.tol = 0.f;
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
        // slack in frame units: eps_world through the frame's row sums; + the a_w/b_w rounding; + pair mismatch
        const double eps = 1e-4 + (mismatch[i] > 0 ? eps_pair : eps_exact) * l1 + 4e-6 * (std::fabs(f.aw) + std::fabs(f.bw)) +
                           (partner[i] >= 0 ? 1000.0 * pair_dist[i] * l1 : 0.0);
        const double tol = 8e-6 * (std::fabs((double)T.D) + 1e-3) + 8e-6 * n1 * O1 + (partner[i] >= 0 ? 2e-3 : 0.0);
        u.ax = (float)f.A.x; u.ay = (float)f.A.y; u.az = (float)f.A.z; u.aw = (float)(f.aw + eps);
        u.bx = (float)f.B.x; u.by = (float)f.B.y; u.bz = (float)f.B.z; u.bw = (float)(f.bw + eps);
        u.c1 = (float)(1.0 + 3.0 * eps);                          // first half:  alpha' + beta' <= c1
        u.c2 = partner[i] >= 0 ? (float)(1.0 + 2.0 * eps) : -1.f;  // second half: alpha', beta' <= c2 (never true for singles)
        u.c3 = (float)(1.0 + eps);                                // second half: alpha' + beta' >= c3
        u.near = mismatch[i] > 0 ? (float)near_dist : 0.f;
        out.tol = std::max(out.tol, (float)tol);  // `behind` needs |D - n.o| above the rounding of that difference: one bound for all units
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
