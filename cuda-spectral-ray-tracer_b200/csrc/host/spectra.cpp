// Spectral upsampling and per-material spectra, all on the host (the reference bakes them in
// its <<<1,1>>> create_world_kernel, scene/scene.cu:48-50).
//
// The reference looks coefficients up in a 9.4 MB table (utils/srgb_to_spectrum.cu) that is not
// part of its repository snapshot.  That table is the output of the Jakob-Hanika optimiser
// (pbrt-v4 rgb2spec_opt: Gauss-Newton fit of a sigmoid-of-quadratic spectrum in CIELAB, sRGB
// gamut, D65).  Instead of shipping the table we evaluate the optimiser ON DEMAND for exactly the
// cells a colour needs and cache them: a cell costs a few dozen 3x3 solves.
#include "srt_host.hpp"
#include "../common/cie_tables.h"
#include <cmath>
#include <cstring>
#include <map>
#include <array>
#include <mutex>

namespace srt {

// ------------------------------------------------------------------ CIE tables as float
namespace {
struct CieF {
    float t[4][SRT_NS];
    CieF() {
        for (int i = 0; i < SRT_NS; i++) {
            t[0][i] = (float)srt_cie_rows[i][0];
            t[1][i] = (float)srt_cie_rows[i][1];
            t[2][i] = (float)srt_cie_rows[i][2];
            t[3][i] = (float)(srt_cie_rows[i][3] / SRT_D65_NORM);  // utils/cie_const.cu:83-101
        }
    }
};
const CieF& cie() {
    static CieF c;
    return c;
}
}  // namespace
const float* cie_table(int which) { return cie().t[which]; }

float spectrum_interp_host(const float* s, float lambda) {  // spectrum/spectrum.cu:11-22
    lambda -= 360.0f;
    lambda *= (float(SRT_NS) - 1) / (830.0f - 360.0f);
    int offset = (int)lambda;
    if (offset < 0) offset = 0;
    if (offset > SRT_NS - 2) offset = SRT_NS - 2;
    const float w = lambda - float(offset);
    return (1.0f - w) * s[offset] + w * s[offset + 1];
}

// ------------------------------------------------------------------ Jakob-Hanika optimiser
namespace rgb2spec {
namespace {
constexpr int kRes = 64;
constexpr int kFine = (SRT_NS - 1) * 3 + 1;  // 283 quadrature nodes
constexpr double kLo = 360.0, kHi = 830.0;

struct Quadrature {
    double lambda[kFine];
    double rgb[3][kFine];  // D65-weighted sRGB response per node
    double white[3];
    static double sample(int col, double x) {
        x -= kLo;
        x *= (SRT_NS - 1) / (kHi - kLo);
        int o = (int)x;
        if (o < 0) o = 0;
        if (o > SRT_NS - 2) o = SRT_NS - 2;
        const double w = x - o;
        double a = srt_cie_rows[o][col], b = srt_cie_rows[o + 1][col];
        if (col == 3) { a /= SRT_D65_NORM; b /= SRT_D65_NORM; }
        return (1.0 - w) * a + w * b;
    }
    Quadrature() {
        static const double xyz_to_srgb[3][3] = {{3.240479, -1.537150, -0.498535}, {-0.969256, 1.875991, 0.041556}, {0.055648, -0.204043, 1.057311}};
        std::memset(rgb, 0, sizeof rgb);
        white[0] = white[1] = white[2] = 0.0;
        const double h = (kHi - kLo) / (kFine - 1);
        for (int i = 0; i < kFine; ++i) {
            const double l = kLo + i * h;
            const double xyz[3] = {sample(0, l), sample(1, l), sample(2, l)};
            const double I = sample(3, l);
            double wgt = 3.0 / 8.0 * h;  // Simpson 3/8 composite weights
            if (i == 0 || i == kFine - 1) {}
            else if ((i - 1) % 3 == 2) wgt *= 2.0;
            else wgt *= 3.0;
            lambda[i] = l;
            for (int k = 0; k < 3; ++k)
                for (int j = 0; j < 3; ++j) rgb[k][i] += xyz_to_srgb[k][j] * xyz[j] * I * wgt;
            for (int k = 0; k < 3; ++k) white[k] += xyz[k] * I * wgt;
        }
    }
};
const Quadrature& quad() {
    static Quadrature q;
    return q;
}
inline double sigmoid(double x) { return 0.5 * x / std::sqrt(1.0 + x * x) + 0.5; }
inline double smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }
inline double lab_f(double t) {
    const double d = 6.0 / 29.0;
    return t > d * d * d ? std::cbrt(t) : t / (d * d * 3.0) + (4.0 / 29.0);
}
void to_lab(double* p) {
    static const double srgb_to_xyz[3][3] = {{0.412453, 0.357580, 0.180423}, {0.212671, 0.715160, 0.072169}, {0.019334, 0.119193, 0.950227}};
    const Quadrature& q = quad();
    double X = 0.0, Y = 0.0, Z = 0.0;
    for (int j = 0; j < 3; ++j) {
        X += p[j] * srgb_to_xyz[0][j];
        Y += p[j] * srgb_to_xyz[1][j];
        Z += p[j] * srgb_to_xyz[2][j];
    }
    const double fx = lab_f(X / q.white[0]), fy = lab_f(Y / q.white[1]), fz = lab_f(Z / q.white[2]);
    p[0] = 116.0 * fy - 16.0;
    p[1] = 500.0 * (fx - fy);
    p[2] = 200.0 * (fy - fz);
}
void residual(const double* c, const double* target, double* res) {
    const Quadrature& q = quad();
    double out[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < kFine; ++i) {
        const double t = (q.lambda[i] - kLo) / (kHi - kLo);
        double x = 0.0;
        for (int k = 0; k < 3; ++k) x = x * t + c[k];
        const double s = sigmoid(x);
        for (int j = 0; j < 3; ++j) out[j] += q.rgb[j][i] * s;
    }
    to_lab(out);
    std::memcpy(res, target, sizeof(double) * 3);
    to_lab(res);
    for (int j = 0; j < 3; ++j) res[j] -= out[j];
}
bool solve3(double m[3][3], const double* b, double* x) {  // LU with partial pivoting on row pointers
    double* A[3] = {m[0], m[1], m[2]};
    int P[4] = {0, 1, 2, 3};
    for (int i = 0; i < 3; ++i) {
        double best = 0.0;
        int imax = i;
        for (int k = i; k < 3; ++k) {
            const double a = std::fabs(A[k][i]);
            if (a > best) { best = a; imax = k; }
        }
        if (best < 1e-15) return false;
        if (imax != i) {
            std::swap(P[i], P[imax]);
            std::swap(A[i], A[imax]);
            P[3]++;
        }
        for (int j = i + 1; j < 3; ++j) {
            A[j][i] /= A[i][i];
            for (int k = i + 1; k < 3; ++k) A[j][k] -= A[j][i] * A[i][k];
        }
    }
    for (int i = 0; i < 3; ++i) {
        x[i] = b[P[i]];
        for (int k = 0; k < i; ++k) x[i] -= A[i][k] * x[k];
    }
    for (int i = 2; i >= 0; --i) {
        for (int k = i + 1; k < 3; ++k) x[i] -= A[i][k] * x[k];
        x[i] /= A[i][i];
    }
    return true;
}
bool gauss_newton(const double target[3], double c[3]) {
    constexpr double eps = 1e-4;
    for (int it = 0; it < 15; ++it) {
        double res[3], J[3][3], step[3];
        residual(c, target, res);
        for (int i = 0; i < 3; ++i) {  // central differences
            double r0[3], r1[3], tmp[3];
            std::memcpy(tmp, c, sizeof tmp);
            tmp[i] -= eps;
            residual(tmp, target, r0);
            std::memcpy(tmp, c, sizeof tmp);
            tmp[i] += eps;
            residual(tmp, target, r1);
            for (int j = 0; j < 3; ++j) J[j][i] = (r1[j] - r0[j]) * 1.0 / (2 * eps);
        }
        if (!solve3(J, res, step)) return false;
        double r = 0.0;
        for (int j = 0; j < 3; ++j) {
            c[j] -= step[j];
            r += res[j] * res[j];
        }
        const double mx = std::fmax(std::fmax(c[0], c[1]), c[2]);
        if (mx > 200) for (int j = 0; j < 3; ++j) c[j] *= 200 / mx;
        if (r < 1e-6) break;
    }
    return true;
}
std::mutex g_mu;
std::map<std::array<int, 4>, std::array<float, 3>> g_cache;
}  // namespace

float scale(int k) { return (float)smoothstep(smoothstep(k / double(kRes - 1))); }

bool cell(int l, int k, int j, int i, float out[3]) {
    std::lock_guard<std::mutex> lock(g_mu);
    const std::array<int, 4> key{l, k, j, i};
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { std::memcpy(out, it->second.data(), sizeof(float) * 3); return true; }
    // replay the optimiser's warm-started brightness sweep for this (l, j, i) column: it starts
    // at k0 = res/5 from zero coefficients and walks towards k, caching every cell on the way
    const double y = j / double(kRes - 1), x = i / double(kRes - 1);
    const int start = kRes / 5, dir = k >= start ? 1 : -1;
    double c[3] = {0.0, 0.0, 0.0}, rgb[3];
    for (int kk = start;; kk += dir) {
        const double b = (double)scale(kk);
        rgb[l] = b;
        rgb[(l + 1) % 3] = x * b;
        rgb[(l + 2) % 3] = y * b;
        if (!gauss_newton(rgb, c)) return false;
        const double c0 = 360.0, c1 = 1.0 / (830.0 - 360.0);
        const double A = c[0], B = c[1], C = c[2];
        std::array<float, 3> o{float(A * (c1 * c1)), float(B * c1 - 2 * A * c0 * (c1 * c1)), float(C - B * c0 * c1 + A * ((c0 * c1) * (c0 * c1)))};
        g_cache[{l, kk, j, i}] = o;
        if (kk == k) { std::memcpy(out, o.data(), sizeof(float) * 3); return true; }
    }
}

namespace {
struct CellAddr { int maxc, zi, yi, xi; float dx, dy, dz; };
int find_scale_interval(float z) {  // largest i in [0, 62] with Scale[i] < z (binary search form of color_to_spectrum.cuh:49-61)
    long size = kRes - 2, first = 1;
    while (size > 0) {
        const long half = size >> 1, middle = first + half;
        const bool pr = scale((int)middle) < z;
        first = pr ? middle + 1 : first;
        size = pr ? size - (half + 1) : half;
    }
    long r = first - 1;
    return (int)(r < 0 ? 0 : (r > kRes - 2 ? kRes - 2 : r));
}
CellAddr address(vec3f c) {  // color_to_spectrum.cuh:122-133
    const float rgb[3] = {c.x, c.y, c.z};
    CellAddr a;
    a.maxc = (c.x > c.y) ? ((c.x > c.z) ? 0 : 2) : ((c.y > c.z) ? 1 : 2);
    const float z = rgb[a.maxc];
    const float x = rgb[(a.maxc + 1) % 3] * (kRes - 1) / z;
    const float y = rgb[(a.maxc + 2) % 3] * (kRes - 1) / z;
    a.xi = (int)x < kRes - 2 ? (int)x : kRes - 2;
    a.yi = (int)y < kRes - 2 ? (int)y : kRes - 2;
    a.zi = find_scale_interval(z);
    a.dx = x - a.xi;
    a.dy = y - a.yi;
    a.dz = (z - scale(a.zi)) / (scale(a.zi + 1) - scale(a.zi));
    return a;
}
inline float lerp(float t, float a, float b) { return (1 - t) * a + t * b; }
}  // namespace

void coeffs_nearest(vec3f c, float out[3]) {  // dev_get_sigmoid_coeffs, color_to_spectrum.cuh:109-151
    if (c.x == c.y && c.y == c.z) { out[0] = 0.0f; out[1] = 0.0f; out[2] = (c.x - .5f) / std::sqrt(c.x * (1 - c.x)); return; }
    const CellAddr a = address(c);
    float v3[3], r[3];
    cell(a.maxc, a.zi + (int)a.dz, a.yi + (int)a.dy, a.xi + (int)a.dx, v3);
    for (int i = 0; i < 3; i++) {  // the reference still runs its trilinear Lerp chain on the single repeated value
        const float v = v3[i];
        r[i] = lerp(a.dz, lerp(a.dy, lerp(a.dx, v, v), lerp(a.dx, v, v)), lerp(a.dy, lerp(a.dx, v, v), lerp(a.dx, v, v)));
    }
    out[0] = r[2]; out[1] = r[1]; out[2] = r[0];
}
void coeffs_trilinear(vec3f c, float out[3]) {  // get_sigmoid_coeffs (host path), color_to_spectrum.cuh:69-107
    if (c.x == c.y && c.y == c.z) { out[0] = 0.0f; out[1] = 0.0f; out[2] = (c.x - .5f) / std::sqrt(c.x * (1 - c.x)); return; }
    const CellAddr a = address(c);
    float v[2][2][2][3], r[3];
    for (int dz = 0; dz < 2; dz++)
        for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++) cell(a.maxc, a.zi + dz, a.yi + dy, a.xi + dx, v[dz][dy][dx]);
    for (int i = 0; i < 3; i++)
        r[i] = lerp(a.dz, lerp(a.dy, lerp(a.dx, v[0][0][0][i], v[0][0][1][i]), lerp(a.dx, v[0][1][0][i], v[0][1][1][i])),
                    lerp(a.dy, lerp(a.dx, v[1][0][0][i], v[1][0][1][i]), lerp(a.dx, v[1][1][0][i], v[1][1][1][i])));
    out[0] = r[2]; out[1] = r[1]; out[2] = r[0];
}
}  // namespace rgb2spec

// ------------------------------------------------------------------ spectra
namespace {
inline float sigmoid_checked(float x) {  // color_to_spectrum.cuh:37-40
    if (std::isinf(x)) return x > 0 ? 1 : 0;
    return 0.5f * x / std::sqrt(1.0f + x * x) + 0.5f;
}
// Samples sigmoid(poly) at 360 + i * (470/95) nm -- the reference's 4.947 nm write spacing (Q8) --
// with polynomial(lambda, c.z, c.y, c.x) exactly as color_to_spectrum.cuh:153-156,212-213 orders it.
void sample_sigmoid(const float c[3], float gain, const float* illuminant, float* out) {
    const float step = (830.0f - 360.0f) / SRT_NS;
    float lambda = 360.0f;
    for (int i = 0; i < SRT_NS; i++) {
        const float x = lambda * lambda * c[2] + lambda * c[1] + c[0];
        const float s = sigmoid_checked(x);
        out[i] = illuminant ? gain * s * spectrum_interp_host(illuminant, lambda) : s;
        lambda += step;
    }
}
}  // namespace

void HostMaterial::bake_spectrum() {  // material::compute_spectral_distr, materials/material.cuh:71-84
    float c[3];
    switch (type) {
        case SRT_EMISSIVE:
            rgb2spec::coeffs_nearest(color, c);
            sample_sigmoid(c, std::pow(power, 2.0f), cie_table(3), spec);
            break;
        case SRT_DIELECTRIC:
            for (float& v : spec) v = 1.0f;
            break;
        default:
            rgb2spec::coeffs_nearest(color, c);
            sample_sigmoid(c, 1.0f, nullptr, spec);
            break;
    }
}

HostMaterial HostMaterial::from_desc(const srt_material_desc& d, bool ref_compat) {
    HostMaterial m;
    m.type = d.type;
    m.color = vec3f(d.color[0], d.color[1], d.color[2]);
    m.fuzz = d.fuzz;
    m.power = d.emission_power;
    if (d.type == SRT_DIELECTRIC) {
        for (int i = 0; i < 3; i++) {
            m.B[i] = d.sellmeier_b[i];
            m.C[i] = ref_compat ? d.sellmeier_b[i] : d.sellmeier_c[i];  // material.cuh:67 stores b[] into sellmeier_C (F4)
        }
    } else {
        m.B[0] = 1.0f;  // the 5-argument constructor parks `ir` in sellmeier_B[0] (material.cuh:50-61)
    }
    m.bake_spectrum();
    return m;
}

void background_spectrum(vec3f rgb, float out[SRT_NS]) {  // srgb_to_illuminance_spectrum (host), color_to_spectrum.cuh:159-171
    float c[3];
    rgb2spec::coeffs_trilinear(rgb, c);
    sample_sigmoid(c, std::pow(1.0f, 2.0f), cie_table(3), out);
}

}  // namespace srt
