/* TEST INFRASTRUCTURE ONLY (oracle/) -- see srt_oracle.h.
 *
 * Plain-C restatement of the reference's hot path, one function per reference function, same
 * float32 operation order everywhere (compile with -O2 -ffp-contract=off, as oracle/Makefile
 * does).  All file:line citations are relative to /root/reference.
 */
#include "srt_oracle.h"
#include "rgb2spec.h"
#include "cie_tables.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define N_WL 7
#define NS 95
#define EPSILON 0.0001f          /* materials/material.cuh:14 */
#define PI_F 3.1415926535897932385f /* utils/utility.h:12 */

/* ------------------------------------------------------------------ vec3 (math/vec3.cuh) */
static inline ov3 V(float x, float y, float z) { ov3 r = {x, y, z}; return r; }
static inline ov3 vadd(ov3 a, ov3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }          /* :119-121 */
static inline ov3 vsub(ov3 a, ov3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }          /* :124-126 */
static inline ov3 vneg(ov3 a) { return V(-a.x, -a.y, -a.z); }                                /* :36 */
static inline ov3 vscale(float t, ov3 v) { return V(t * v.x, t * v.y, t * v.z); }            /* :134-136 */
static inline ov3 vdiv(ov3 v, float t) { return vscale(1 / t, v); }                          /* :144-147 */
static inline float vdot(ov3 u, ov3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }         /* :149-153 */
static inline ov3 vcross(ov3 u, ov3 v) {                                                     /* :155-159 */
    return V(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
static inline float vlen2(ov3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }               /* :66-71 */
static inline float vlen(ov3 v) { return sqrtf(vlen2(v)); }                                  /* :75-77 */
static inline ov3 vunit(ov3 v) { return vdiv(v, vlen(v)); }                                  /* :161-163 */
static inline float vget(ov3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
static inline ov3 vmatmul(ov3 v, const float* m) {                                           /* :80-91 */
    return V((m[0] * v.x) + (m[1] * v.y) + (m[2] * v.z), (m[3] * v.x) + (m[4] * v.y) + (m[5] * v.z),
             (m[6] * v.x) + (m[7] * v.y) + (m[8] * v.z));
}

/* ------------------------------------------------------------------ XORWOW (curand_kernel.h) */
void srt_oracle_rng_init(uint32_t seed, orng* s) {
    uint32_t s0 = seed ^ 0xaad26b49u, s1 = 0u ^ 0xf7dcefddu; /* seed>>32 == 0: seeds are 32-bit (rendering.cu:137) */
    uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}
uint32_t srt_oracle_rng_next(orng* s) {
    uint32_t t = s->v[0] ^ (s->v[0] >> 2);
    s->v[0] = s->v[1]; s->v[1] = s->v[2]; s->v[2] = s->v[3]; s->v[3] = s->v[4];
    s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
    s->d += 362437u;
    return s->v[4] + s->d;
}
typedef struct { orng r; ocounters* c; } rngc;
static inline float rnd(rngc* s) { /* cuda_random_float, utils/cuda_utility.cu:19-25 */
    if (s->c) s->c->rng_draws++;
    return (float)srt_oracle_rng_next(&s->r) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}
float srt_oracle_rng_uniform(orng* s) { return (float)srt_oracle_rng_next(s) * 2.3283064e-10f + (2.3283064e-10f / 2.0f); }
static inline float rnd_range(float mn, float mx, rngc* s) { /* cuda_utility.cu:27-41 */
    float range_width = mx - mn;
    float random = rnd(s);
    return random * range_width + mn;
}
static inline int rnd_int(int mn, int mx, rngc* s) { /* cuda_utility.cu:43-48 */
    float f = rnd_range((float)(mn - 1), (float)(mx - 1), s);
    return (int)ceilf(f);
}

/* ------------------------------------------------------------------ spectrum (spectrum/spectrum.cu) */
static float cie_f[4][NS]; /* x y z normalised-D65 as float */
static int cie_ready = 0;
static void cie_init(void) {
    if (cie_ready) return;
    for (int i = 0; i < NS; ++i) {
        cie_f[0][i] = (float)srt_cie_rows[i][0];
        cie_f[1][i] = (float)srt_cie_rows[i][1];
        cie_f[2][i] = (float)srt_cie_rows[i][2];
        cie_f[3][i] = (float)(srt_cie_rows[i][3] / SRT_D65_NORM); /* utils/cie_const.cu:83-101 */
    }
    cie_ready = 1;
}
static inline float interp95(const float* spectrum, float lambda) { /* spectrum.cu:11-22 */
    lambda -= 360.0f;
    lambda *= ((float)NS - 1) / (830.0f - 360.0f);
    int offset = (int)lambda;
    if (offset < 0) offset = 0;
    if (offset > NS - 2) offset = NS - 2;
    float weight = lambda - (float)offset;
    return (1.0f - weight) * spectrum[offset] + weight * spectrum[offset + 1];
}
float srt_oracle_spectrum_interp(const float* t, float lambda) { return interp95(t, lambda); }

static void init_hero(float* wl, rngc* s) { /* spectrum.cu:31-46 */
    float step = (830.0f - 360.0f) / (float)N_WL;
    float hero = rnd_range(360.0f, 830.0f, s);
    wl[0] = hero;
    float lambda = hero;
    for (int i = 1; i < N_WL; i++) {
        lambda += step;
        if (lambda > 830.0f) {
            float remainder = lambda - 830.0f;
            lambda = 360.0f + remainder;
        }
        wl[i] = lambda;
    }
}

/* ------------------------------------------------------------------ colour -> spectrum (color/color_to_spectrum.cuh) */
static float sigmoid_inf_check(float x) { /* :37-40 */
    if (isinf(x)) return x > 0 ? 1 : 0;
    return 0.5f * x / sqrtf(1.0f + x * x) + 0.5f;
}
static inline float lerpf(float x, float a, float b) { return (1 - x) * a + x * b; } /* :64-66 */
static inline float polynomial(float x, float c2, float c1, float c0) { return x * x * c2 + x * c1 + c0; } /* :153-156 */

static float scale64[64];
static int scale_ready = 0;
static int find_interval_scale(float z) { /* FindInterval, :49-61, pred = Scale[i] < z */
    long size = 64 - 2, first = 1;
    while (size > 0) {
        long half = size >> 1, middle = first + half;
        int pr = scale64[middle] < z;
        first = pr ? middle + 1 : first;
        size = pr ? size - (half + 1) : half;
    }
    long r = first - 1;
    if (r < 0) r = 0;
    if (r > 62) r = 62;
    return (int)r;
}
/* dev_get_sigmoid_coeffs (:109-151): nearest-cell lookup (the interpolation lambda ignores its
 * arguments, :139-142), the Lerp chain is still evaluated on the repeated value. */
static ov3 dev_get_sigmoid_coeffs(ov3 col) {
    float r = col.x, g = col.y, b = col.z;
    float rgb[3] = {r, g, b};
    if (r == g && g == b) return V(0.0f, 0.0f, (r - .5f) / sqrtf(r * (1 - r)));
    if (!scale_ready) {
        for (int k = 0; k < 64; ++k) scale64[k] = srt_oracle_rgb2spec_scale(k, 64);
        scale_ready = 1;
    }
    int maxc = (r > g) ? ((r > b) ? 0 : 2) : ((g > b) ? 1 : 2);
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * (64 - 1) / z;
    float y = rgb[(maxc + 2) % 3] * (64 - 1) / z;
    int xi = (int)x < 62 ? (int)x : 62, yi = (int)y < 62 ? (int)y : 62;
    int zi = find_interval_scale(z);
    float dx = x - xi, dy = y - yi, dz = (z - scale64[zi]) / (scale64[zi + 1] - scale64[zi]);
    float c[3], cell[3];
    srt_oracle_rgb2spec_cell(maxc, zi + (int)dz, yi + (int)dy, xi + (int)dx, 64, cell);
    for (int i = 0; i < 3; ++i) {
        float v = cell[i];
        c[i] = lerpf(dz, lerpf(dy, lerpf(dx, v, v), lerpf(dx, v, v)), lerpf(dy, lerpf(dx, v, v), lerpf(dx, v, v)));
    }
    return V(c[2], c[1], c[0]);
}
static void dev_srgb_to_spectrum(ov3 col, float* out) { /* :204-219 */
    float step = (830.0f - 360.0f) / NS;
    ov3 co = dev_get_sigmoid_coeffs(col);
    float lambda = 360.0f;
    for (int i = 0; i < NS; i++) {
        float x = polynomial(lambda, co.z, co.y, co.x);
        out[i] = sigmoid_inf_check(x);
        lambda += step;
    }
}
static void dev_srgb_to_illuminance_spectrum(ov3 col, float* out, float power) { /* :173-186 */
    cie_init();
    float step = (830.0f - 360.0f) / NS;
    ov3 co = dev_get_sigmoid_coeffs(col);
    float lambda = 360.0f;
    for (int i = 0; i < NS; i++) {
        float x = polynomial(lambda, co.z, co.y, co.x);
        out[i] = powf(power, 2.0f) * sigmoid_inf_check(x) * interp95(cie_f[3], lambda);
        lambda += step;
    }
}
void srt_oracle_color_spectrum(float r, float g, float b, int emissive, float power, float out95[95]) {
    if (emissive) dev_srgb_to_illuminance_spectrum(V(r, g, b), out95, power);
    else dev_srgb_to_spectrum(V(r, g, b), out95);
}

/* ------------------------------------------------------------------ materials (materials/material.cuh) */
static omat mat_make(ov3 col, float fuzz, float ir, float power, int type) { /* :50-61 */
    omat m;
    memset(&m, 0, sizeof m);
    m.col = col; m.fuzz = fuzz; m.type = type; m.power = power;
    m.B[0] = ir;
    return m;
}
static omat mat_lambertian(ov3 c) { return mat_make(c, 1.0f, 1.0f, 0.0f, O_LAMBERTIAN); }       /* :111-113 */
static omat mat_metallic(ov3 c, float fuzz) { return mat_make(c, fuzz, 1.0f, 0.0f, O_METALLIC); } /* :116-118 */
static omat mat_emissive(ov3 c, float p) { return mat_make(c, 1.0f, 1.0f, p, O_EMISSIVE); }      /* :106-108 */
/* 0 (default): the reference as shipped, sellmeier_C[i] = b[i] (material.cuh:67, sic).  1: the same constructor with the
 * one-token fix sellmeier_C[i] = c[i] -- the physically meant Sellmeier equation.  Pinned against oracle/_ref built with
 * exactly that token patched (build_ref.sh --physical, tests/golden/ref_physical.npz). */
static int g_physical_sellmeier = 0;
void srt_oracle_set_physical_sellmeier(int on) { g_physical_sellmeier = on != 0; }
static omat mat_dielectric(const float b[3], const float c[3]) { /* :63-69 */
    omat m;
    memset(&m, 0, sizeof m);
    m.col = V(1.0f, 1.0f, 1.0f); m.fuzz = 1.0f; m.type = O_DIELECTRIC; m.power = 0.0f;
    for (int i = 0; i < 3; i++) { m.B[i] = b[i]; m.C[i] = g_physical_sellmeier ? c[i] : b[i]; }
    return m;
}
static void mat_compute_spectral_distr(omat* m) { /* :71-84 */
    switch (m->type) {
    case O_EMISSIVE: dev_srgb_to_illuminance_spectrum(m->col, m->spec, m->power); break;
    case O_DIELECTRIC: for (int i = 0; i < NS; i++) m->spec[i] = 1.0f; break;
    default: dev_srgb_to_spectrum(m->col, m->spec); break;
    }
}
/* refraction/sellmeier.cuh:6-13 */
static const float BK7_b[3] = {1.03961212f, 0.231792344f, 1.01046945f};
static const float BK7_c[3] = {6.00069867e-3f, 2.00179144e-2f, 1.03560653e2f};
static const float fused_silica_b[3] = {0.6961663f, 0.4079426f, 0.8974794f};
static const float fused_silica_c[3] = {0.0684043f, 0.1162414f, 9.896161f};
static const float flint_glass_b[3] = {1.34533359f, 0.209073176f, 0.937357162f};
static const float flint_glass_c[3] = {0.00997743871f, 0.0470450767f, 111.886764f};

/* the three coefficient tables of refraction/sellmeier.cuh:6-13: 0 BK7, 1 fused silica, 2 flint glass */
void srt_oracle_glass(int which, float b[3], float c[3]) {
    const float* sb = which == 0 ? BK7_b : (which == 1 ? fused_silica_b : flint_glass_b);
    const float* sc = which == 0 ? BK7_c : (which == 1 ? fused_silica_c : flint_glass_c);
    for (int i = 0; i < 3; i++) { b[i] = sb[i]; c[i] = sc[i]; }
}
float srt_oracle_sellmeier(const float b[3], const float c[3], float lambda) { /* sellmeier.cu:11-23 */
    lambda *= 1e-3f;
    float l2 = lambda * lambda;
    float index = 1.0f + (b[0] * l2) / (l2 - c[0]) + (b[1] * l2) / (l2 - c[1]) + (b[2] * l2) / (l2 - c[2]);
    return sqrtf(index);
}

/* ------------------------------------------------------------------ triangles (primitives/tri.cu) */
static float dsa2d(const otri* t, ov3 v1, ov3 v2, ov3 v3) { /* tri.cu:153-181 */
    int w, h;
    switch (t->aa_plane) {
    case O_AA_YZ: w = 1; h = 2; break;
    case O_AA_XZ: w = 0; h = 2; break;
    default: w = 0; h = 1;
    }
    return (vget(v1, w) - vget(v3, w)) * (vget(v2, h) - vget(v3, h)) - (vget(v2, w) - vget(v3, w)) * (vget(v1, h) - vget(v3, h));
}
static void tri_init(otri* t) { /* tri.cu:47-84 */
    ov3 n = vcross(vsub(t->v[1], t->v[0]), vsub(t->v[2], t->v[0]));
    t->n = vunit(n);
    int perp_x = fabsf(vdot(t->n, V(1.f, 0.f, 0.f))) < 1e-8f;
    int perp_y = fabsf(vdot(t->n, V(0.f, 1.f, 0.f))) < 1e-8f;
    int perp_z = fabsf(vdot(t->n, V(0.f, 0.f, 1.f))) < 1e-8f;
    if (perp_y && perp_z) t->aa_plane = O_AA_YZ;
    else if (perp_x && perp_z) t->aa_plane = O_AA_XZ;
    else if (perp_x && perp_y) t->aa_plane = O_AA_XY;
    /* else: sticky -- keeps whatever it had (fresh triangles: 0 = NONE, Q6) */
    t->D = vdot(t->n, t->v[0]);
    t->clockwise = dsa2d(t, t->v[0], t->v[1], t->v[2]) >= 0;
    /* aabb(v0,v1,v2).pad(): bvh/aabb.cuh:49-57, 93-102; interval.cuh:58-62 */
    for (int a = 0; a < 3; a++) {
        float lo = fminf(vget(t->v[0], a), fminf(vget(t->v[1], a), vget(t->v[2], a)));
        float hi = fmaxf(vget(t->v[0], a), fmaxf(vget(t->v[1], a), vget(t->v[2], a)));
        float delta = 0.0001f;
        if (!((hi - lo) >= delta)) {
            float padding = delta / 2;
            lo = lo - padding;
            hi = hi + padding;
        }
        t->bb[2 * a] = lo;
        t->bb[2 * a + 1] = hi;
    }
}
static otri tri_new(ov3 v1, ov3 v2, ov3 v3, uint32_t mat, int vectors) { /* tri.cuh:28-48 */
    otri t;
    memset(&t, 0, sizeof t); /* fresh heap memory assumed zero: aa_plane = NONE (Q6) */
    t.mat = mat;
    t.v[0] = v1;
    if (vectors) { t.v[1] = vadd(v1, v2); t.v[2] = vadd(v1, v3); }
    else { t.v[1] = v2; t.v[2] = v3; }
    tri_init(&t);
    return t;
}
static void tri_translate(otri* t, ov3 dir, int reinit) { /* tri.cu:86-94 */
    for (int k = 0; k < 3; k++) t->v[k] = vadd(t->v[k], dir);
    if (reinit) tri_init(t);
}
static void rot_matrix_y(float theta, float* m) { /* primitives/transform.cu:3-34, AXIS::Y */
    float c = cosf(theta), s = sinf(theta);
    m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1;
    m[0] = c; m[2] = s; m[6] = -s; m[8] = c;
}
static void tri_rotate_y_nolocal(otri* t, float theta) { /* tri.cu:96-119 with local=false, reinit=false */
    float m[9];
    rot_matrix_y(theta, m);
    for (int k = 0; k < 3; k++) t->v[k] = vmatmul(t->v[k], m);
}
/* tri_quad (primitives/tri_quad.cuh:13-20) -> two triangles at dst[0..1] */
static void quad_new(otri* dst, ov3 Q, ov3 u, ov3 v, uint32_t mat) {
    dst[0] = tri_new(Q, u, v, mat, 1);
    dst[1] = tri_new(vadd(vadd(Q, u), v), vneg(u), vneg(v), mat, 1);
}
static ov3 quad_u(const otri* q) { return vsub(q[0].v[1], q[0].v[0]); }
static ov3 quad_v(const otri* q) { return vsub(q[0].v[2], q[0].v[0]); }
static ov3 quad_Q(const otri* q) { return q[0].v[0]; }
static ov3 quad_center(const otri* q) { return vadd(vdiv(vadd(quad_u(q), quad_v(q)), 2.0f), quad_Q(q)); } /* :44-46 */

/* tri_box (primitives/tri_box.cuh:29-44): 6 quads front/right/back/left/top/bottom at dst[0..11] */
static void box_new(otri* dst, ov3 a, ov3 b, const uint32_t mats[6]) {
    ov3 mn = V(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z));
    ov3 mx = V(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z));
    ov3 dx = V(mx.x - mn.x, 0.f, 0.f), dy = V(0, mx.y - mn.y, 0.f), dz = V(0, 0, mx.z - mn.z);
    quad_new(dst + 0, V(mn.x, mn.y, mx.z), dx, dy, mats[0]);
    quad_new(dst + 2, V(mx.x, mn.y, mx.z), vneg(dz), dy, mats[1]);
    quad_new(dst + 4, V(mx.x, mn.y, mn.z), vneg(dx), dy, mats[2]);
    quad_new(dst + 6, V(mn.x, mn.y, mn.z), dz, dy, mats[3]);
    quad_new(dst + 8, V(mn.x, mx.y, mx.z), dx, vneg(dz), mats[4]);
    quad_new(dst + 10, V(mn.x, mn.y, mn.z), dx, dz, mats[5]);
}
static ov3 box_center(const otri* bx) { /* tri_box.cuh:117-122 */
    ov3 mn = quad_Q(bx + 10);
    ov3 mx = vadd(vadd(vadd(mn, quad_u(bx + 10)), quad_v(bx + 6)), quad_v(bx + 10));
    return vadd(vdiv(vsub(mx, mn), 2.0f), mn);
}
static void tris_rotate_y_local(otri* t, int n, ov3 center, float theta) { /* tri_box.cu:14-35 etc. */
    for (int i = 0; i < n; i++) tri_translate(t + i, vneg(center), 0);
    for (int i = 0; i < n; i++) tri_rotate_y_nolocal(t + i, theta);
    for (int i = 0; i < n; i++) tri_translate(t + i, center, 0);
}
/* pyramid (primitives/pyramid.cuh:29-47): base quad dst[0..1], sides dst[2..5] */
static void pyramid_new(otri* dst, ov3 Q, ov3 u, ov3 v, ov3 w, uint32_t mat) {
    quad_new(dst, Q, u, v, mat);
    ov3 top = vadd(quad_center(dst), w);
    ov3 v1 = vadd(Q, u), v2 = vadd(Q, v), v3 = vadd(v2, u);
    dst[2] = tri_new(Q, top, v2, mat, 0);
    dst[3] = tri_new(v1, top, Q, mat, 0);
    dst[4] = tri_new(v2, top, v3, mat, 0);
    dst[5] = tri_new(v3, top, v1, mat, 0);
}
/* prism (primitives/prism.cuh:22-32): caps dst[0..1], side quads dst[2..7] */
static void prism_new(otri* dst, ov3 Q, ov3 u, ov3 v, ov3 w, uint32_t mat) {
    dst[0] = tri_new(Q, v, u, mat, 1);
    dst[1] = tri_new(vadd(Q, w), u, v, mat, 1);
    quad_new(dst + 2, Q, u, w, mat);
    quad_new(dst + 4, Q, w, v, mat);
    quad_new(dst + 6, vadd(Q, u), vsub(v, u), w, mat);
}
static ov3 prism_centroid(const otri* p) { /* prism.cuh:44-54 */
    ov3 s = vadd(vadd(vadd(vadd(vadd(p[0].v[0], p[0].v[1]), p[0].v[2]), p[1].v[0]), p[1].v[1]), p[1].v[2]);
    return vdiv(s, 6.f);
}
static inline float deg2rad(float d) { return d * PI_F / 180.0f; } /* utils/cuda_utility.cuh:41-43 */

/* ------------------------------------------------------------------ reference BVH (bvh/bvh.cu) */
typedef struct { int left, right, is_leaf, prim; float bb[6]; } onode;
struct oscene {
    int ntris, nmats;
    otri* tris;
    omat* mats;
    int* order; /* the tri* array after the reference's in-place sorts */
    onode* nodes;
    int nnodes, root;
    int valid;
};

static int box_compare(const oscene* s, int a, int b, int axis) { /* bvh.cuh:179-184 */
    return s->tris[a].bb[2 * axis] < s->tris[b].bb[2 * axis];
}
static int partition_(const oscene* s, int* o, int l, int h, int axis) { /* bvh.cu:14-31 */
    if (l == h) return l;
    int x = o[h];
    int i = l - 1;
    for (int j = l; j < h; j++) {
        if (box_compare(s, o[j], x, axis)) {
            i++;
            int t = o[i]; o[i] = o[j]; o[j] = t;
        }
    }
    int t = o[i + 1]; o[i + 1] = o[h]; o[h] = t;
    return i + 1;
}
static void quicksort_(const oscene* s, int* o, int start, int end, int axis) { /* bvh.cu:33-71 */
    int* stack = (int*)malloc(sizeof(int) * (size_t)(end - start + 2) * 2);
    int top = -1;
    stack[++top] = start;
    stack[++top] = end;
    while (top >= 0) {
        end = stack[top--];
        start = stack[top--];
        int p = partition_(s, o, start, end, axis);
        if (p - 1 > start) { stack[++top] = start; stack[++top] = p - 1; }
        if (p + 1 < end) { stack[++top] = p + 1; stack[++top] = end; }
    }
    free(stack);
}
static int node_alloc(oscene* s, int is_leaf) {
    onode* n = &s->nodes[s->nnodes];
    n->left = n->right = -1;
    n->is_leaf = is_leaf;
    n->prim = -1;
    n->bb[0] = n->bb[2] = n->bb[4] = FLT_MAX;
    n->bb[1] = n->bb[3] = n->bb[5] = -FLT_MAX;
    return s->nnodes++;
}
static void node_box(const oscene* s, int n, float* bb) { /* bvh_node::bounding_box, bvh.cuh:56-58 */
    if (s->nodes[n].is_leaf) memcpy(bb, s->tris[s->nodes[n].prim].bb, sizeof(float) * 6);
    else memcpy(bb, s->nodes[n].bb, sizeof(float) * 6);
}
static void postorder_boxes(oscene* s, int n) { /* build_nodes_bboxes, bvh.cu:311-345 (same unions, recursive form) */
    if (s->nodes[n].is_leaf) return;
    postorder_boxes(s, s->nodes[n].left);
    postorder_boxes(s, s->nodes[n].right);
    float a[6], b[6];
    node_box(s, s->nodes[n].left, a);
    node_box(s, s->nodes[n].right, b);
    for (int k = 0; k < 3; k++) {
        s->nodes[n].bb[2 * k] = fminf(a[2 * k], b[2 * k]);
        s->nodes[n].bb[2 * k + 1] = fmaxf(a[2 * k + 1], b[2 * k + 1]);
    }
}
#define REF_MAX_DEPTH 64
static int build_ref_bvh(oscene* s) { /* bvh.cu:206-309, create_bvh_kernel scene.cu:9-20 (seed 1984) */
    rngc rs;
    rs.c = NULL;
    srt_oracle_rng_init(1984u, &rs.r);
    int n = s->ntris;
    s->nodes = (onode*)malloc(sizeof(onode) * (size_t)(2 * n + 2));
    s->order = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) s->order[i] = i;
    s->nnodes = 0;
    s->valid = 0;
    if (n <= 0) return 0;
    size_t st_start[REF_MAX_DEPTH], st_end[REF_MAX_DEPTH];
    int st_node[REF_MAX_DEPTH];
    int tos = -1;
    s->root = node_alloc(s, 0);
    tos++;
    st_start[tos] = 0; st_end[tos] = (size_t)n; st_node[tos] = s->root;
    while (tos >= 0) {
        size_t cs = st_start[tos], ce = st_end[tos];
        int node = st_node[tos];
        tos--;
        size_t span = ce - cs;
        if (span > 0) {
            if (span == 1) {
                s->nodes[node].is_leaf = 1;
                s->nodes[node].left = s->nodes[node].right = -1;
                s->nodes[node].prim = s->order[cs];
            } else {
                int axis = rnd_int(0, 2, &rs);
                if (span == 2) {
                    int a = s->order[cs], b = s->order[cs + 1];
                    int l = node_alloc(s, 1), r = node_alloc(s, 1);
                    s->nodes[node].left = l; s->nodes[node].right = r;
                    if (box_compare(s, a, b, axis)) { s->nodes[l].prim = a; s->nodes[r].prim = b; }
                    else { s->nodes[l].prim = b; s->nodes[r].prim = a; }
                } else {
                    quicksort_(s, s->order, (int)cs, (int)(ce - 1), axis);
                    int l = node_alloc(s, 0), r = node_alloc(s, 0);
                    s->nodes[node].left = l; s->nodes[node].right = r;
                    size_t mid = cs + span / 2;
                    tos++;
                    if (tos >= REF_MAX_DEPTH) return 0;
                    st_start[tos] = cs; st_end[tos] = mid; st_node[tos] = l;
                    tos++;
                    if (tos >= REF_MAX_DEPTH) return 0;
                    st_start[tos] = mid; st_end[tos] = ce; st_node[tos] = r;
                }
            }
        }
    }
    postorder_boxes(s, s->root);
    s->valid = 1;
    return 1;
}
static void preorder_(const oscene* s, int n, int* out, int* k) {
    if (n < 0) return;
    out[(*k)++] = s->nodes[n].is_leaf ? s->nodes[n].prim : -1;
    preorder_(s, s->nodes[n].left, out, k);
    preorder_(s, s->nodes[n].right, out, k);
}
int srt_oracle_scene_refbvh_preorder(const oscene* s, int* out) {
    int k = 0;
    if (s->valid) preorder_(s, s->root, out, &k);
    return k;
}

/* ------------------------------------------------------------------ intersection */
typedef struct { ov3 p, n; float t; int front; uint32_t mat; } ohit;
typedef struct { ov3 o, d; uint32_t valid; float wl[N_WL], pw[N_WL]; } oray;

static int aabb_hit(const float* bb, const oray* r, float mn, float mx, ocounters* c) { /* bvh/aabb.cu:7-39 */
    if (c) c->box_tests++;
    for (int a = 0; a < 3; a++) {
        float inv = 1 / vget(r->d, a);
        float orig = vget(r->o, a);
        float t0, t1;
        if (inv >= 0) { t0 = (bb[2 * a] - orig) * inv; t1 = (bb[2 * a + 1] - orig) * inv; }
        else { t1 = (bb[2 * a] - orig) * inv; t0 = (bb[2 * a + 1] - orig) * inv; }
        if (t0 > mn) mn = t0;
        if (t1 < mx) mx = t1;
        if (mx <= mn) return 0;
    }
    return 1;
}
static int tri_hit(const otri* t, const oray* r, float mn, float mx, ohit* rec, ocounters* c) { /* tri.cu:3-45 */
    if (c) c->tri_tests++;
    float denom = vdot(t->n, r->d);
    if (fabsf(denom) < 1e-8f) return 0;
    float tt = (t->D - vdot(t->n, r->o)) / denom;
    if (!(mn <= tt && tt <= mx)) return 0; /* interval::contains, math/interval.cuh:43-46 */
    ov3 p = vadd(r->o, vscale(tt, r->d));  /* ray::at, ray/ray.cuh:44-47 */
    float a1 = dsa2d(t, p, t->v[0], t->v[1]); /* is_interior_faster, tri.cu:121-128 */
    float a2 = dsa2d(t, p, t->v[1], t->v[2]);
    float a3 = dsa2d(t, p, t->v[2], t->v[0]);
    int inside = t->clockwise ? (a1 >= 0.f && a2 >= 0.f && a3 >= 0.f) : (a1 <= 0.f && a2 <= 0.f && a3 <= 0.f);
    if (!inside) return 0;
    rec->t = tt;
    rec->p = p;
    rec->mat = t->mat;
    rec->front = vdot(r->d, t->n) < 0; /* set_face_normal, primitives/hit_record.cuh:30-43 */
    rec->n = rec->front ? t->n : vneg(t->n);
    return 1;
}
static int node_hit(const oscene* s, int n, const oray* r, float mn, float mx, ohit* rec, ocounters* c) { /* bvh.cu:73-76 */
    return s->nodes[n].is_leaf ? tri_hit(&s->tris[s->nodes[n].prim], r, mn, mx, rec, c) : aabb_hit(s->nodes[n].bb, r, mn, mx, c);
}
static int g_debug_brute = 0; /* srt_oracle_debug_pixel only: closest hit over ALL triangles instead of the reference's pruned walk */
static int bvh_hit(const oscene* s, const oray* r, float mn, float mx, ohit* rec, ocounters* c) { /* bvh.cu:98-166 */
    if (!s->valid) return 0;
    if (g_debug_brute) {
        int h = 0;
        float cl = mx;
        ohit tmp;
        for (int i = 0; i < s->ntris; i++)
            if (tri_hit(&s->tris[i], r, mn, cl, &tmp, NULL)) { h = 1; cl = tmp.t; *rec = tmp; }
        return h;
    }
    if (c) c->rays++;
    int hit_anything = 0;
    float closest = mx;
    int stack[64];
    int sp = 0;
    stack[sp++] = -1;
    int node = s->root;
    if (s->nodes[node].is_leaf) {
        if (node_hit(s, node, r, mn, closest, rec, c)) { hit_anything = 1; closest = rec->t; }
    } else do {
        int cl = s->nodes[node].left, cr = s->nodes[node].right;
        ohit tmp;
        int hl = cl >= 0 && node_hit(s, cl, r, mn, closest, &tmp, c);
        if (hl && s->nodes[cl].is_leaf) { hit_anything = 1; closest = tmp.t; *rec = tmp; }
        int hr = cr >= 0 && node_hit(s, cr, r, mn, closest, &tmp, c);
        if (hr && s->nodes[cr].is_leaf) { hit_anything = 1; closest = tmp.t; *rec = tmp; }
        int tl = cl >= 0 && (hl && !s->nodes[cl].is_leaf);
        int tr = cr >= 0 && (hr && !s->nodes[cr].is_leaf);
        if (!tl && !tr) node = stack[--sp];
        else {
            node = tl ? cl : cr;
            if (tl && tr) stack[sp++] = cr;
        }
    } while (node >= 0);
    return hit_anything;
}

/* ------------------------------------------------------------------ scatter (materials/material.cu) */
static ov3 random_in_unit_sphere(rngc* s) { /* vec3.cuh:209-218; draw order x,y,z (device order, Q13) */
    for (;;) {
        if (s->c) s->c->rejection_iters++;
        float a = rnd_range(-1, 1, s);
        float b = rnd_range(-1, 1, s);
        float c = rnd_range(-1, 1, s);
        ov3 p = V(a, b, c);
        if (vlen2(p) < 1.0f) return p;
    }
}
static ov3 random_unit_vector(rngc* s) { return vunit(random_in_unit_sphere(s)); } /* :220-227 */
static ov3 reflect_(ov3 v, ov3 n) { return vsub(v, vscale(2 * vdot(v, n), n)); }     /* :179-183 */
static ov3 refract_(ov3 uv, ov3 n, float eta) {                                       /* :198-205 */
    float cos_theta = fminf(vdot(vneg(uv), n), 1.0f);
    ov3 perp = vscale(eta, vadd(uv, vscale(cos_theta, n)));
    ov3 par = vscale(-sqrtf(fabsf(1.0f - vlen2(perp))), n);
    return vadd(perp, par);
}
static float reflectance_(float cosine, float ref_idx) { /* material.cu:39-53 */
    float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0f - r0) * powf(1.0f - cosine, 5.0f);
}
static void mul_spectrum(oray* r, const float* spec, ocounters* c) { /* ray/ray.cuh:60-69 */
    for (uint32_t i = 0; i < r->valid; i++) {
        if (c) c->interps++;
        r->pw[i] *= interp95(spec, r->wl[i]);
    }
}
static int scatter(const omat* m, oray* r, const ohit* rec, rngc* s) { /* material.cu:55-100 */
    ov3 dir = V(0, 0, 0);
    float eps_sign = 1.0f;
    int did = 1;
    ov3 unit_in = vunit(r->d);
    switch (m->type) {
    case O_METALLIC: { /* reflection_scatter :22-37 */
        if (s->c) s->c->scatter_metal++;
        ov3 reflected = reflect_(unit_in, rec->n);
        dir = vadd(reflected, vscale(m->fuzz, random_unit_vector(s)));
        did = vdot(dir, rec->n) > 0;
        if (!did) r->valid = 0;
        break;
    }
    case O_DIELECTRIC: { /* refraction_scatter :103-135 */
        if (s->c) s->c->scatter_dielectric++;
        float ir = srt_oracle_sellmeier(m->B, m->C, r->wl[0]);
        float ratio = rec->front ? (1.0f / ir) : ir;
        float cos_theta = fminf(vdot(vneg(unit_in), rec->n), 1.0f);
        float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        int cannot = ratio * sin_theta > 1.0f || reflectance_(cos_theta, ratio) > rnd(s);
        if (cannot) dir = reflect_(unit_in, rec->n);
        else { dir = refract_(unit_in, rec->n, ratio); eps_sign = -1.0f; }
        if (!cannot) r->valid = 1;
        break;
    }
    case O_EMISSIVE:
        did = 0;
        break;
    default: { /* lambertian_scatter :9-19 */
        if (s->c) s->c->scatter_lambert++;
        dir = vadd(rec->n, random_unit_vector(s));
        float e = 1e-8f;
        if ((fabsf(dir.x) < e) && (fabsf(dir.y) < e) && (fabsf(dir.z) < e)) dir = rec->n;
        break;
    }
    }
    mul_spectrum(r, m->spec, s->c);
    r->o = vadd(rec->p, vscale(eps_sign * EPSILON, rec->n));
    r->d = dir;
    return did;
}

/* ------------------------------------------------------------------ camera (rendering/camera.cu:7-58) */
void srt_oracle_camera_make(int w, int h, float vfov, ov3 lookfrom, ov3 lookat, ov3 vup, float defocus_angle,
                            float focus_dist, ov3 background, ocam* c) {
    c->w = w; c->h = h;
    c->center = lookfrom;
    float theta = deg2rad(vfov);
    float hh = tanf(theta / 2.0f) * focus_dist;
    float viewport_height = 2.0f * hh;
    float viewport_width = viewport_height * ((float)w / (float)h);
    ov3 ww = vunit(vsub(lookfrom, lookat));
    ov3 u = vunit(vcross(vup, ww));
    ov3 v = vcross(ww, u);
    ov3 viewport_u = vscale(viewport_width, u);
    ov3 viewport_v = vscale(viewport_height, vneg(v));
    c->du = vdiv(viewport_u, (float)w);
    c->dv = vdiv(viewport_v, (float)h);
    ov3 ul = vsub(vsub(vsub(c->center, vscale(focus_dist, ww)), vdiv(viewport_u, 2)), vdiv(viewport_v, 2));
    c->p00 = vadd(ul, vscale(0.5f, vadd(c->du, c->dv)));
    float defocus_radius = focus_dist * tanf(deg2rad(defocus_angle / 2));
    c->disk_u = vscale(defocus_radius, u);
    c->disk_v = vscale(defocus_radius, v);
    c->defocus_angle = defocus_angle;
    c->background = background;
}
void srt_oracle_camera_default(int w, int h, ocam* out) { /* scene/scene.cu:259-320: identical for all three scenes */
    srt_oracle_camera_make(w, h, 40.0f, V(278, 278, -800), V(278, 278, 0), V(0, 1, 0), 0.0f, 10.0f, V(0, 0, 0), out);
}
int srt_oracle_yres(int xres, float ar) { /* io/params.h:176-180 */
    unsigned y = (unsigned)((unsigned)xres / ar);
    return y < 1 ? 1 : (int)y;
}

/* host-side background spectrum: srgb_to_illuminance_spectrum with the TRILINEAR host lookup
 * (color/color_to_spectrum.cuh:69-107,159-171).  Grey/black backgrounds take the closed form. */
static void background_spectrum(ov3 col, float* out) {
    cie_init();
    float r = col.x, g = col.y, b = col.z;
    ov3 co;
    if (r == g && g == b) co = V(0.0f, 0.0f, (r - .5f) / sqrtf(r * (1 - r)));
    else {
        float rgb[3] = {r, g, b};
        if (!scale_ready) { for (int k = 0; k < 64; ++k) scale64[k] = srt_oracle_rgb2spec_scale(k, 64); scale_ready = 1; }
        int maxc = (r > g) ? ((r > b) ? 0 : 2) : ((g > b) ? 1 : 2);
        float z = rgb[maxc];
        float x = rgb[(maxc + 1) % 3] * (64 - 1) / z, y = rgb[(maxc + 2) % 3] * (64 - 1) / z;
        int xi = (int)x < 62 ? (int)x : 62, yi = (int)y < 62 ? (int)y : 62, zi = find_interval_scale(z);
        float dx = x - xi, dy = y - yi, dz = (z - scale64[zi]) / (scale64[zi + 1] - scale64[zi]);
        float cc[3], cell[2][2][2][3];
        for (int a = 0; a < 2; a++) for (int bb = 0; bb < 2; bb++) for (int c = 0; c < 2; c++)
            srt_oracle_rgb2spec_cell(maxc, zi + a, yi + bb, xi + c, 64, cell[a][bb][c]);
        for (int i = 0; i < 3; i++)
            cc[i] = lerpf(dz, lerpf(dy, lerpf(dx, cell[0][0][0][i], cell[0][0][1][i]), lerpf(dx, cell[0][1][0][i], cell[0][1][1][i])),
                          lerpf(dy, lerpf(dx, cell[1][0][0][i], cell[1][0][1][i]), lerpf(dx, cell[1][1][0][i], cell[1][1][1][i])));
        co = V(cc[2], cc[1], cc[0]);
    }
    float step = (830.0f - 360.0f) / NS, lambda = 360.0f;
    for (int i = 0; i < NS; i++) {
        float x = polynomial(lambda, co.z, co.y, co.x);
        out[i] = powf(1.0f, 2.0f) * sigmoid_inf_check(x) * interp95(cie_f[3], lambda);
        lambda += step;
    }
}

/* ------------------------------------------------------------------ scenes (scene/scene.cu:73-257) */
static oscene* scene_alloc(int ntris, int nmats) {
    oscene* s = (oscene*)calloc(1, sizeof(oscene));
    s->ntris = ntris; s->nmats = nmats;
    s->tris = (otri*)calloc((size_t)(ntris > 0 ? ntris : 1), sizeof(otri));
    s->mats = (omat*)calloc((size_t)(nmats > 0 ? nmats : 1), sizeof(omat));
    return s;
}
static void walls_and_light(otri* d, const uint32_t wall_mats[5], uint32_t light_mat) { /* scene.cu:83-102 */
    quad_new(d + 0, V(0, 0, 0), V(0, 0, 555), V(555, 0, 0), wall_mats[0]);       /* bottom */
    quad_new(d + 4, V(0, 0, 555.f), V(0, 555, 0), V(555, 0, 0), wall_mats[1]);   /* back */
    quad_new(d + 2, V(555, 555, 555), V(-555, 0, 0), V(0, 0, -555), wall_mats[2]); /* top */
    quad_new(d + 6, V(555, 0, 0), V(0, 0, 555), V(0, 555, 0), wall_mats[3]);     /* left */
    quad_new(d + 8, V(0, 0, 0), V(0, 555, 0), V(0, 0, 555), wall_mats[4]);       /* right */
    ov3 center = V(555.f / 2.f, 554.f, 555.f / 2.f);
    float width = 100.f, depth = 100.f;
    ov3 Q = V((center.x + width / 2.f), center.y, (center.z + depth / 2.f));
    quad_new(d + 10, Q, V(-width, 0, 0), V(0, 0, -depth), light_mat);
}
static void cornell_objects(otri* d, const uint32_t box1[6], const uint32_t box2[6], uint32_t pyr_mat) { /* scene.cu:114-128 */
    box_new(d + 12, V(0.f, 0.f, 0.f), V(165.f, 330.f, 165.f), box1);
    tris_rotate_y_local(d + 12, 12, box_center(d + 12), deg2rad(25.f));
    for (int i = 0; i < 12; i++) tri_translate(d + 12 + i, V(265.f, 0.f, 295.f), 1);
    box_new(d + 24, V(0.f, 0.f, 0.f), V(165.f, 165.f, 165.f), box2);
    tris_rotate_y_local(d + 24, 12, box_center(d + 24), deg2rad(-18.f));
    for (int i = 0; i < 12; i++) tri_translate(d + 24 + i, V(130.f, 0.f, 65.f), 1);
    pyramid_new(d + 36, V(165.f, 166.f, 0.f), V(-165.f, 0.f, 0.f), V(0.f, 0.f, 165.f), V(0.f, 165.f, 0.f), pyr_mat);
    tris_rotate_y_local(d + 36, 6, quad_center(d + 36), deg2rad(-18.f));
    for (int i = 0; i < 6; i++) tri_translate(d + 36 + i, V(130.f, 0.f, 65.f), 1);
}
oscene* srt_oracle_scene_create(int id) {
    oscene* s;
    if (id == 1) { /* device_prism_test, scene.cu:132-173 */
        s = scene_alloc(20, 3);
        s->mats[0] = mat_lambertian(V(.73f, .73f, .73f));
        s->mats[1] = mat_emissive(V(1, 1, 1), 5);
        s->mats[2] = mat_dielectric(flint_glass_b, flint_glass_c);
        const uint32_t wm[5] = {0, 0, 0, 0, 0};
        walls_and_light(s->tris, wm, 1);
        ov3 center = V(555.f / 2.f, 554.f, 555.f / 2.f);
        float width = 100.f, prism_width = 165.f, prism_height = 200.f;
        prism_new(s->tris + 12, V(center.x - width / 2.f, center.y - 1.f, center.z - prism_height / 2.f), V(0.f, -prism_width, 0.f),
                  V((prism_width * sqrtf(3.f)) / 2.f, -prism_width / 2.f, 0.f), V(0.f, 0.f, 200.f), 2);
        tris_rotate_y_local(s->tris + 12, 8, prism_centroid(s->tris + 12), deg2rad(10.f));
        for (int i = 0; i < 8; i++) tri_init(s->tris + 12 + i); /* rotate(..., reinit = true) */
    } else if (id == 2) { /* device_different_mats_world, scene.cu:175-226 */
        s = scene_alloc(42, 9);
        s->mats[0] = mat_lambertian(V(.65f, .05f, .05f));
        s->mats[1] = mat_lambertian(V(.12f, .45f, .15f));
        s->mats[2] = mat_dielectric(flint_glass_b, flint_glass_c);
        s->mats[3] = mat_lambertian(V(.73f, .73f, .73f));
        s->mats[4] = mat_emissive(V(1.f, 1.f, 1.f), 5.f);
        s->mats[5] = mat_metallic(V(.5f, .5f, .5f), 0.3f);
        s->mats[6] = mat_lambertian(V(.12f, .15f, .45f));
        s->mats[7] = mat_dielectric(BK7_b, BK7_c);
        s->mats[8] = mat_metallic(V(.7f, .7f, .7f), 0.8f);
        const uint32_t wm[5] = {6, 1, 2, 8, 5};
        walls_and_light(s->tris, wm, 4);
        const uint32_t b1[6] = {3, 8, 0, 1, 2, 3}, b2[6] = {7, 6, 8, 7, 1, 2};
        cornell_objects(s->tris, b1, b2, 2);
    } else { /* device_cornell_box, scene.cu:73-130 */
        s = scene_alloc(42, 7);
        s->mats[0] = mat_lambertian(V(.65f, .05f, .05f));
        s->mats[1] = mat_lambertian(V(.12f, .45f, .15f));
        s->mats[2] = mat_dielectric(flint_glass_b, flint_glass_c);
        s->mats[3] = mat_lambertian(V(.73f, .73f, .73f));
        s->mats[4] = mat_emissive(V(1.f, 1.f, 1.f), 5.f);
        s->mats[5] = mat_metallic(V(.5f, .5f, .5f), 0.3f);
        s->mats[6] = mat_lambertian(V(.12f, .15f, .45f));
        const uint32_t wm[5] = {3, 3, 3, 1, 6};
        walls_and_light(s->tris, wm, 4);
        const uint32_t b1[6] = {5, 5, 5, 5, 5, 5}, b2[6] = {0, 0, 0, 0, 0, 0};
        cornell_objects(s->tris, b1, b2, 2);
    }
    for (int i = 0; i < s->nmats; i++) mat_compute_spectral_distr(&s->mats[i]); /* scene.cu:48-50 */
    build_ref_bvh(s);
    return s;
}

/* SplitMix64 -> 24-bit uniform in [0,1); shared spec with the product's soup generator */
static inline uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float sm_uniform(uint64_t* x) { return (float)(splitmix64(x) >> 40) * (1.0f / 16777216.0f); }
oscene* srt_oracle_scene_soup(int n, uint64_t seed) {
    /* centre ~ U[0,1)^3 * 555, vertices = centre + U(-s,s)^3, s = 555 * n^(-1/3);
     * material 0 = lambertian .73 grey, material 1 = emissive white power 5 on the last 2 triangles
     * (a 100x100 light quad under the ceiling, as in the Cornell scene). */
    oscene* s = scene_alloc(n, 2);
    s->mats[0] = mat_lambertian(V(.73f, .73f, .73f));
    s->mats[1] = mat_emissive(V(1.f, 1.f, 1.f), 5.f);
    float sz = 555.0f * powf((float)n, -1.0f / 3.0f);
    uint64_t st = seed;
    int nsoup = n >= 3 ? n - 2 : n; /* tiny soups: no light quad */
    for (int i = 0; i < nsoup; i++) {
        float c[3], p[9];
        for (int k = 0; k < 3; k++) c[k] = sm_uniform(&st) * 555.0f;
        for (int k = 0; k < 9; k++) p[k] = c[k % 3] + (sm_uniform(&st) * 2.0f - 1.0f) * sz;
        s->tris[i] = tri_new(V(p[0], p[1], p[2]), V(p[3], p[4], p[5]), V(p[6], p[7], p[8]), 0, 0);
    }
    if (n >= 3) {
        ov3 center = V(555.f / 2.f, 554.f, 555.f / 2.f);
        quad_new(s->tris + nsoup, V(center.x + 50.f, center.y, center.z + 50.f), V(-100.f, 0, 0), V(0, 0, -100.f), 1);
    }
    for (int i = 0; i < s->nmats; i++) mat_compute_spectral_distr(&s->mats[i]);
    if (n <= 4096) build_ref_bvh(s); /* the reference's serial builder is O(n log^2 n) with depth-64 stack; small soups only */
    return s;
}
void srt_oracle_scene_destroy(oscene* s) {
    if (!s) return;
    free(s->tris); free(s->mats); free(s->order); free(s->nodes); free(s);
}
int srt_oracle_scene_ntris(const oscene* s) { return s->ntris; }
int srt_oracle_scene_nmats(const oscene* s) { return s->nmats; }
const otri* srt_oracle_scene_tris(const oscene* s) { return s->tris; }
const omat* srt_oracle_scene_mats(const oscene* s) { return s->mats; }

/* ------------------------------------------------------------------ film (color/color.cu) */
static const float XYZ2SRGB[9] = {3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f, 0.0415560f,
                                  0.0556434f, -0.2040259f, 1.0572252f}; /* utils/color_const.cu:17-19 */
static float correct_channel(float v) { /* color.cu:15-22 */
    return v < 0.0f ? 0.0f : (v < 0.0031308f ? 12.92f * v : (v < 1.0f ? ((1.055f * powf(v, 0.416666f)) - 0.055f) : 1.0f));
}
void srt_oracle_tonemap(const float m[3], float out[3]) { /* XYZ_to_sRGB :35-41 + expand_sRGB :43-49 */
    ov3 s = vmatmul(V(m[0], m[1], m[2]), XYZ2SRGB);
    out[0] = (float)(int)(correct_channel(s.x) * 255.99f);
    out[1] = (float)(int)(correct_channel(s.y) * 255.99f);
    out[2] = (float)(int)(correct_channel(s.z) * 255.99f);
}
static ov3 spectrum_to_xyz(const float* wl, const float* pw, uint32_t nvalid, ocounters* c) { /* color.cu:88-104 */
    cie_init();
    float delta = (830.0f - 360.0f) / (float)N_WL;
    float x = 0.0f, y = 0.0f, z = 0.0f;
    for (uint32_t i = 0; i < nvalid; i++) {
        if (c) c->interps += 3;
        x += interp95(cie_f[0], wl[i]) * pw[i] * delta;
        y += interp95(cie_f[1], wl[i]) * pw[i] * delta;
        z += interp95(cie_f[2], wl[i]) * pw[i] * delta;
    }
    return V(x, y, z);
}
void srt_oracle_spectrum_to_xyz(const float wl[7], const float pw[7], int nvalid, float xyz[3]) {
    ov3 r = spectrum_to_xyz(wl, pw, (uint32_t)nvalid, NULL);
    xyz[0] = r.x; xyz[1] = r.y; xyz[2] = r.z;
}

/* ------------------------------------------------------------------ render (rendering/rendering.cu) */
/* strat_n = 0: get_ray :66-87 (pixel_sample_square :49-56); strat_n > 0: get_ray_stratified_sample :89-118
 * (pixel_stratified_sample_square :58-64) for sub-cell (sx, sy) of a strat_n x strat_n grid */
static oray get_ray_s(const ocam* c, uint32_t i, uint32_t j, uint32_t sx, uint32_t sy, float recip_sqrt_spp, int strat, rngc* s) { /* ray.cuh:27-58 */
    ov3 pixel_center = vadd(vadd(c->p00, vscale((float)i, c->du)), vscale((float)j, c->dv));
    float px, py;
    if (strat) {
        px = -0.5f + recip_sqrt_spp * ((float)sx + rnd(s));
        py = -0.5f + recip_sqrt_spp * ((float)sy + rnd(s));
    } else {
        px = -0.5f + rnd(s);
        py = -0.5f + rnd(s);
    }
    ov3 pixel_sample = vadd(pixel_center, vadd(vscale(px, c->du), vscale(py, c->dv)));
    ov3 origin = c->center;
    if (!(c->defocus_angle <= 0.0f)) { /* defocus_disk_sample :42-47, random_in_unit_disk vec3.cuh:240-246 */
        ov3 p;
        for (;;) {
            float a = rnd_range(-1, 1, s);
            float b = rnd_range(-1, 1, s);
            p = V(a, b, 0);
            if (vlen2(p) < 1.0f) break;
        }
        origin = vadd(vadd(c->center, vscale(p.x, c->disk_u)), vscale(p.y, c->disk_v));
    }
    oray r;
    r.o = origin;
    r.d = vsub(pixel_sample, origin);
    init_hero(r.wl, s);
    for (int k = 0; k < N_WL; k++) r.pw[k] = 1.0f;
    r.valid = N_WL;
    return r;
}
static oray get_ray(const ocam* c, uint32_t i, uint32_t j, rngc* s) { return get_ray_s(c, i, j, 0, 0, 0.0f, 0, s); }
static void ray_bounce(const oscene* sc, const float* bg, oray* r, int bounce_limit, rngc* s) { /* :12-40 */
    ohit rec;
    for (int n = 0; n < bounce_limit; n++) {
        if (s->c && (isnan(r->d.x) || isnan(r->d.y) || isnan(r->d.z))) s->c->nan_rays++;
        if (!bvh_hit(sc, r, 0.0f, FLT_MAX, &rec, s->c)) {
            mul_spectrum(r, bg, s->c);
            if (s->c) s->c->end_miss++;
            return;
        }
        const omat* m = &sc->mats[rec.mat];
        if (!scatter(m, r, &rec, s)) {
            if (s->c) { if (m->type == O_EMISSIVE) s->c->end_emissive++; else s->c->end_absorbed++; }
            return;
        }
    }
    if (s->c) s->c->end_limit++;
    r->valid = 0;
}

static void counters_add(ocounters* a, const ocounters* b) {
    uint64_t* pa = (uint64_t*)a;
    const uint64_t* pb = (const uint64_t*)b;
    for (size_t i = 0; i < sizeof(ocounters) / sizeof(uint64_t); i++) pa[i] += pb[i];
}

int srt_oracle_render(const oscene* sc, const ocam* cam, int spp, int bounce_limit, int chunk_w, int chunk_h, float* rgb,
                      float* xyz, ocounters* counters, int nthreads) {
    return srt_oracle_render_opts(sc, cam, spp, bounce_limit, chunk_w, chunk_h, 0, rgb, xyz, counters, nthreads);
}
int srt_oracle_render_opts(const oscene* sc, const ocam* cam, int spp, int bounce_limit, int chunk_w, int chunk_h, int stratified,
                           float* rgb, float* xyz, ocounters* counters, int nthreads) {
    const int W = cam->w, H = cam->h;
    spp = (int)(unsigned short)spp;                   /* short_uint kernel parameters, rendering.cu:154 (Q14) */
    /* opt-in stratified pixel sampling: the reference carries the sampler (rendering.cu:58-64, 89-118) but its kernel
     * never calls it.  Sample k of a pixel takes sub-cell (k % n, k / n) of an n x n grid, n*n = spp (required). */
    int strat_n = 0;
    if (stratified) {
        while ((strat_n + 1) * (strat_n + 1) <= spp) strat_n++;
        if (strat_n * strat_n != spp) return -2;
    }
    const float recip_sqrt_spp = strat_n ? 1.0f / (float)strat_n : 0.0f;
    bounce_limit = (int)(unsigned short)bounce_limit;
    if (chunk_w <= 0 && chunk_h > 0) chunk_w = chunk_h; /* io/params.h:53-63 */
    if (chunk_h <= 0 && chunk_w > 0) chunk_h = chunk_w;
    if (chunk_w <= 0) chunk_w = W;
    if (chunk_h <= 0) chunk_h = H;
    const uint32_t tx = 28, ty = 16; /* render_manager.cu:93-96 */
    const uint32_t gx = (uint32_t)chunk_w / tx + 1, gy = (uint32_t)chunk_h / ty + 1;
    const size_t nslots = (size_t)gx * gy * tx * ty;
    orng* states = (orng*)malloc(sizeof(orng) * nslots);
    for (size_t k = 0; k < nslots; k++) srt_oracle_rng_init(1984u + (uint32_t)k, &states[k]); /* rendering.cu:120-138 */
    float bg[NS];
    background_spectrum(cam->background, bg); /* rendering.cu:324 */
    if (counters) memset(counters, 0, sizeof *counters);
    const int x_chunks = (int)ceilf((float)W / (float)chunk_w), y_chunks = (int)ceilf((float)H / (float)chunk_h);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    for (int ci = 0; ci < x_chunks * y_chunks; ci++) { /* render_manager::step, render_manager.cu:3-66 */
        const int off_x = (ci % x_chunks) * chunk_w, off_y = (ci / x_chunks) * chunk_h;
        const int cw = off_x + chunk_w > W ? W - off_x : chunk_w, ch = off_y + chunk_h > H ? H - off_y : chunk_h;
#pragma omp parallel
        {
            ocounters local;
            memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 4)
            for (int j = 0; j < ch; j++) {
                for (int i = 0; i < cw; i++) {
                    const uint32_t idx = ((uint32_t)j % ty) * tx + ((uint32_t)i % tx) + tx * ty * (((uint32_t)j / ty) * gx + (uint32_t)i / tx);
                    rngc s;
                    s.r = states[idx];
                    s.c = counters ? &local : NULL;
                    ov3 acc = V(0, 0, 0);
                    if (sc->valid) {
                        for (int k = 0; k < (int)(unsigned short)spp; k++) {
                            oray r = get_ray_s(cam, (uint32_t)(off_x + i), (uint32_t)(off_y + j), strat_n ? (uint32_t)(k % strat_n) : 0u,
                                               strat_n ? (uint32_t)(k / strat_n) : 0u, recip_sqrt_spp, strat_n != 0, &s);
                            ray_bounce(sc, bg, &r, (int)(unsigned short)bounce_limit, &s);
                            ov3 c3 = spectrum_to_xyz(r.wl, r.pw, r.valid, s.c);
                            acc = vadd(acc, c3);
                            if (s.c) s.c->samples++;
                        }
                    }
                    states[idx] = s.r;
                    ov3 mean = vdiv(acc, (float)spp); /* save_to_fb :140-149 */
                    const size_t p = (size_t)(off_y + j) * W + (size_t)(off_x + i);
                    float m3[3] = {mean.x, mean.y, mean.z}, o3[3];
                    srt_oracle_tonemap(m3, o3);
                    for (int c = 0; c < 3; c++) {
                        rgb[(size_t)c * W * H + p] = o3[c];
                        if (xyz) xyz[(size_t)c * W * H + p] = m3[c];
                    }
                }
            }
            if (counters) {
#pragma omp critical
                counters_add(counters, &local);
            }
        }
    }
    free(states);
    return 0;
}

int srt_oracle_render_tiles(const oscene* sc, const ocam* cam, int spp, int bounce_limit, int tile_w, int tile_h, int rank,
                            int world, float* xyz_sum) {
    const int W = cam->w, H = cam->h;
    const uint32_t tx = 28, ty = 16;
    const uint32_t gx = (uint32_t)W / tx + 1;
    const int tiles_x = (W + tile_w - 1) / tile_w;
    float bg[NS];
    background_spectrum(cam->background, bg);
    memset(xyz_sum, 0, sizeof(float) * 3 * (size_t)W * H);
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < H; j++) {
        for (int i = 0; i < W; i++) {
            (void)tiles_x;
            if (((i / tile_w) + 5 * (j / tile_h)) % world != rank) continue; /* diagonal tile interleave */
            const uint32_t idx = ((uint32_t)j % ty) * tx + ((uint32_t)i % tx) + tx * ty * (((uint32_t)j / ty) * gx + (uint32_t)i / tx);
            rngc s;
            srt_oracle_rng_init(1984u + idx, &s.r);
            s.c = NULL;
            ov3 acc = V(0, 0, 0);
            for (int k = 0; k < spp; k++) {
                oray r = get_ray(cam, (uint32_t)i, (uint32_t)j, &s);
                ray_bounce(sc, bg, &r, bounce_limit, &s);
                acc = vadd(acc, spectrum_to_xyz(r.wl, r.pw, r.valid, NULL));
            }
            const size_t p = (size_t)j * W + i;
            xyz_sum[p] = acc.x; xyz_sum[(size_t)W * H + p] = acc.y; xyz_sum[2 * (size_t)W * H + p] = acc.z;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ KAT entry points */
static void hit_out(int h, const ohit* rec, float* out) {
    out[0] = h ? 1.f : 0.f;
    if (h) {
        out[1] = rec->t;
        out[2] = rec->p.x; out[3] = rec->p.y; out[4] = rec->p.z;
        out[5] = rec->n.x; out[6] = rec->n.y; out[7] = rec->n.z;
        out[8] = rec->front ? 1.f : 0.f;
        out[9] = (float)rec->mat;
    }
}
static oray mkray(const float o[3], const float d[3]) {
    oray r;
    memset(&r, 0, sizeof r);
    r.o = V(o[0], o[1], o[2]);
    r.d = V(d[0], d[1], d[2]);
    r.valid = N_WL;
    return r;
}
int srt_oracle_tri_hit(const otri* t, const float o[3], const float d[3], float tmin, float tmax, float out[10]) {
    oray r = mkray(o, d);
    ohit rec;
    int h = tri_hit(t, &r, tmin, tmax, &rec, NULL);
    hit_out(h, &rec, out);
    return h;
}
int srt_oracle_aabb_hit(const float box6[6], const float o[3], const float d[3], float tmin, float tmax) {
    oray r = mkray(o, d);
    return aabb_hit(box6, &r, tmin, tmax, NULL);
}
int srt_oracle_bvh_hit(const oscene* s, const float o[3], const float d[3], float out[10]) {
    oray r = mkray(o, d);
    ohit rec;
    int h = bvh_hit(s, &r, 0.0f, FLT_MAX, &rec, NULL);
    hit_out(h, &rec, out);
    return h;
}
int srt_oracle_brute_hit(const oscene* s, const float o[3], const float d[3], float out[10], int* tri_index) {
    oray r = mkray(o, d);
    ohit rec, tmp;
    int h = 0, best = -1;
    float closest = FLT_MAX;
    for (int i = 0; i < s->ntris; i++) {
        if (tri_hit(&s->tris[i], &r, 0.0f, closest, &tmp, NULL)) { h = 1; closest = tmp.t; rec = tmp; best = i; }
    }
    hit_out(h, &rec, out);
    if (tri_index) *tri_index = best;
    return h;
}
int srt_oracle_scatter(const oscene* s, int mat, float ray_io[21], const float rec_in[8], uint32_t rng[6]) {
    oray r;
    r.o = V(ray_io[0], ray_io[1], ray_io[2]);
    r.d = V(ray_io[3], ray_io[4], ray_io[5]);
    r.valid = (uint32_t)ray_io[6];
    for (int k = 0; k < 7; k++) { r.wl[k] = ray_io[7 + k]; r.pw[k] = ray_io[14 + k]; }
    ohit rec;
    rec.p = V(rec_in[0], rec_in[1], rec_in[2]);
    rec.n = V(rec_in[3], rec_in[4], rec_in[5]);
    rec.t = rec_in[6];
    rec.front = rec_in[7] != 0.f;
    rec.mat = (uint32_t)mat;
    rngc st;
    st.c = NULL;
    st.r.d = rng[0];
    for (int k = 0; k < 5; k++) st.r.v[k] = rng[1 + k];
    int did = scatter(&s->mats[mat], &r, &rec, &st);
    rng[0] = st.r.d;
    for (int k = 0; k < 5; k++) rng[1 + k] = st.r.v[k];
    ray_io[0] = r.o.x; ray_io[1] = r.o.y; ray_io[2] = r.o.z;
    ray_io[3] = r.d.x; ray_io[4] = r.d.y; ray_io[5] = r.d.z;
    ray_io[6] = (float)r.valid;
    for (int k = 0; k < 7; k++) { ray_io[7 + k] = r.wl[k]; ray_io[14 + k] = r.pw[k]; }
    return did;
}
void srt_oracle_get_ray(const ocam* cam, uint32_t i, uint32_t j, uint32_t rng[6], float out[13]) {
    rngc st;
    st.c = NULL;
    st.r.d = rng[0];
    for (int k = 0; k < 5; k++) st.r.v[k] = rng[1 + k];
    oray r = get_ray(cam, i, j, &st);
    rng[0] = st.r.d;
    for (int k = 0; k < 5; k++) rng[1 + k] = st.r.v[k];
    out[0] = r.o.x; out[1] = r.o.y; out[2] = r.o.z;
    out[3] = r.d.x; out[4] = r.d.y; out[5] = r.d.z;
    for (int k = 0; k < 7; k++) out[6 + k] = r.wl[k];
}
void srt_oracle_get_ray_stratified(const ocam* cam, uint32_t i, uint32_t j, uint32_t sx, uint32_t sy, float recip_sqrt_spp, uint32_t rng[6],
                                   float out[13]) {
    rngc st;
    st.c = NULL;
    st.r.d = rng[0];
    for (int k = 0; k < 5; k++) st.r.v[k] = rng[1 + k];
    oray r = get_ray_s(cam, i, j, sx, sy, recip_sqrt_spp, 1, &st);
    rng[0] = st.r.d;
    for (int k = 0; k < 5; k++) rng[1 + k] = st.r.v[k];
    out[0] = r.o.x; out[1] = r.o.y; out[2] = r.o.z;
    out[3] = r.d.x; out[4] = r.d.y; out[5] = r.d.z;
    for (int k = 0; k < 7; k++) out[6 + k] = r.wl[k];
}
/* debugging aid for parity investigations: XYZ of every sample of ONE pixel of a single-chunk render, either through the
 * reference's BVH walk (brute = 0) or with a closest hit over all triangles (brute = 1; not thread safe) */
int srt_oracle_debug_pixel(const oscene* sc, const ocam* cam, int spp, int bounce_limit, int i, int j, int brute, float* xyz_per_sample) {
    const uint32_t tx = 28, ty = 16, gx = (uint32_t)cam->w / tx + 1;
    const uint32_t idx = ((uint32_t)j % ty) * tx + ((uint32_t)i % tx) + tx * ty * (((uint32_t)j / ty) * gx + (uint32_t)i / tx);
    rngc s;
    s.c = NULL;
    srt_oracle_rng_init(1984u + idx, &s.r);
    float bg[NS];
    background_spectrum(cam->background, bg);
    g_debug_brute = brute;
    for (int k = 0; k < spp; k++) {
        oray r = get_ray(cam, (uint32_t)i, (uint32_t)j, &s);
        ray_bounce(sc, bg, &r, bounce_limit, &s);
        ov3 c3 = spectrum_to_xyz(r.wl, r.pw, r.valid, NULL);
        xyz_per_sample[3 * k] = c3.x; xyz_per_sample[3 * k + 1] = c3.y; xyz_per_sample[3 * k + 2] = c3.z;
    }
    g_debug_brute = 0;
    return 0;
}
/* the reference sorts its tri* array in place while building (bvh.cu:262); order[k] = original
 * index of the triangle that ends up at position k of that array */
void srt_oracle_scene_reforder(const oscene* s, int* out) { memcpy(out, s->order, sizeof(int) * (size_t)s->ntris); }
