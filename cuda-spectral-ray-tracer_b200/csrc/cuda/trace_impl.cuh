// Device code of the spectral path tracer.  Included twice -- by trace_fast.cu (default nvcc
// flags: FMA contraction, as the reference's own nvcc build gets) and by trace_strict.cu
// (-fmad=false: rounds exactly like the host oracle) -- inside namespace SRT_FP_NS.
//
// Pipelines
//   wavefront (default): two kernels per iteration over queues of pixel slots
//       k_generate : slots that need a new sample: camera ray -> traverse -> classify
//       k_shade    : slots sorted by the material type they hit (lambertian | metallic |
//                    dielectric segments, warp-uniform): scatter -> traverse -> classify
//     "classify" finishes the sample on miss / emitter / absorption / bounce limit (XYZ added to
//     the film) and pushes the slot to the regenerate queue, or pushes it to the queue of the
//     material it hit.  Queue pushes are warp-aggregated (ballot + one atomicAdd per warp).
//     Exactly one sample is in flight per pixel, so every pixel consumes its XORWOW stream in the
//     reference's order (rendering/rendering.cu:215-228).
//   megakernel: one thread per pixel looping over samples and bounces with the same device
//     functions (used as a cross-check: it must produce bit-identical films).
//
// Reference semantics (file:line relative to the reference): ray generation rendering.cu:66-87,
// hero wavelengths spectrum.cu:31-46, closest hit bvh.cu:98-166 + tri.cu:3-45, scatter
// material.cu:55-135, spectrum lookup spectrum.cu:11-22, XYZ color.cu:88-104, film rendering.cu:140-149.
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include "../common/srt_types.h"
#include "trace_params.h"

namespace srt {
namespace SRT_FP_NS {

// ------------------------------------------------------------------------------ small math
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(float t, V3 v) { return mk(t * v.x, t * v.y, t * v.z); }
__device__ __forceinline__ float dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }
__device__ __forceinline__ float len2(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
__device__ __forceinline__ V3 unit(V3 v) { return (1 / sqrtf(len2(v))) * v; }  // v / |v| = (1/|v|) * v, math/vec3.cuh:144-163
__device__ __forceinline__ float sel3(float x, float y, float z, uint32_t a) { return a == 0 ? x : (a == 1 ? y : z); }

// ------------------------------------------------------------------------------ XORWOW
struct Rng { uint32_t d, v0, v1, v2, v3, v4; };
__device__ __forceinline__ Rng rng_seed(uint32_t seed) {  // curand_init(seed, 0, 0), curand_kernel.h:807-823
    Rng s;
    const uint32_t s0 = seed ^ 0xaad26b49u, s1 = 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s.d = 6615241u + t1 + t0;
    s.v0 = 123456789u + t0;
    s.v1 = 362436069u ^ t0;
    s.v2 = 521288629u + t1;
    s.v3 = 88675123u ^ t1;
    s.v4 = 5783321u + t0;
    return s;
}
__device__ __forceinline__ float rng_uniform(Rng& s) {  // curand_uniform: (0, 1]
    const uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return (float)(s.v4 + s.d) * 2.3283064e-10f + 1.16415320e-10f;
}
__device__ __forceinline__ float rng_range(Rng& s, float lo, float hi) {  // utils/cuda_utility.cu:27-41
    const float width = hi - lo;
    return rng_uniform(s) * width + lo;
}

// ------------------------------------------------------------------------------ scene access
struct SceneRef {
    const SrtNode* nodes;
    const SrtTri* tris;
    const SrtMaterial* mats;
    const float* cie;  // x[95] y[95] z[95]
    const float* bg;   // [95]
    int n_tris;
};

__device__ __forceinline__ float interp95(const float* __restrict__ s, float lambda) {  // spectrum.cu:11-22
    lambda -= 360.0f;
    lambda *= (95.0f - 1) / (830.0f - 360.0f);
    int o = (int)lambda;
    o = o < 0 ? 0 : o;
    o = o > SRT_NS - 2 ? SRT_NS - 2 : o;
    const float w = lambda - (float)o;
    return (1.0f - w) * s[o] + w * s[o + 1];
}

// the 6 rotations of the hero wavelength (spectrum.cu:31-46); recomputed, never stored
__device__ __forceinline__ void hero_rotations(float hero, float wl[SRT_N_WL]) {
    const float step = (830.0f - 360.0f) / 7.0f;
    wl[0] = hero;
    float l = hero;
#pragma unroll
    for (int i = 1; i < SRT_N_WL; i++) {
        l += step;
        if (l > 830.0f) {
            const float rem = l - 830.0f;
            l = 360.0f + rem;
        }
        wl[i] = l;
    }
}

// ------------------------------------------------------------------------------ intersection
// tri::hit (primitives/tri.cu:3-45) on the packed 48-B triangle; returns t through t_out.
__device__ __forceinline__ bool tri_test(const SrtTri* __restrict__ tp, V3 o, V3 d, float closest, float& t_out) {
    const float4 q0 = *reinterpret_cast<const float4*>(tp);
    const float denom = q0.x * d.x + q0.y * d.y + q0.z * d.z;
    if (fabsf(denom) < 1e-8f) return false;
    const float t = (q0.w - (q0.x * o.x + q0.y * o.y + q0.z * o.z)) / denom;
    if (!(0.0f <= t && t <= closest)) return false;
    const float4 q1 = *(reinterpret_cast<const float4*>(tp) + 1);
    const float4 q2 = *(reinterpret_cast<const float4*>(tp) + 2);
    const uint32_t bits = __float_as_uint(q2.z);
    const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
    const float pw = sel3(px, py, pz, SRT_TRI_WAX(bits)), ph = sel3(px, py, pz, SRT_TRI_HAX(bits));
    // double_signed_area_2D (tri.cu:153-181) for (p,v0,v1), (p,v1,v2), (p,v2,v0)
    const float a1 = (pw - q1.z) * (q1.y - q1.w) - (q1.x - q1.z) * (ph - q1.w);
    const float a2 = (pw - q2.x) * (q1.w - q2.y) - (q1.z - q2.x) * (ph - q2.y);
    const float a3 = (pw - q1.x) * (q2.y - q1.y) - (q2.x - q1.x) * (ph - q1.y);
    const bool inside = SRT_TRI_CW(bits) ? (a1 >= 0.f && a2 >= 0.f && a3 >= 0.f) : (a1 <= 0.f && a2 <= 0.f && a3 <= 0.f);
    if (!inside) return false;
    t_out = t;
    return true;
}

// closest hit over the LBVH: both child boxes live in the parent node (4 x 16-B loads),
// near child first, far child pushed.  Returns leaf-order triangle index or -1.
__device__ __forceinline__ int closest_hit(const SceneRef& sc, V3 o, V3 d, float& t_hit) {
    float closest = FLT_MAX;
    int best = -1;
    if (sc.n_tris <= 0) return -1;
    // A NaN ray (buggy-Sellmeier refraction, Q1) can never hit: every tri::hit computes a NaN t.
    // Answer "miss" up front instead of walking the whole tree.
    if (!(d.x == d.x && d.y == d.y && d.z == d.z && o.x == o.x && o.y == o.y && o.z == o.z)) return -1;
    if (sc.n_tris == 1) {
        float t;
        if (tri_test(sc.tris, o, d, closest, t)) { t_hit = t; return 0; }
        return -1;
    }
    const float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    int stack[64];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
        const float4 b0 = np[0], b1 = np[1], b2 = np[2];
        const int4 ch = *reinterpret_cast<const int4*>(np + 3);
        // slabs; a NaN direction makes every comparison below false -> both children are visited,
        // every triangle test then fails (NaN t), i.e. the path misses exactly like the reference (Q1)
        float t0x = (b0.x - o.x) * ix, t1x = (b0.y - o.x) * ix;
        float t0y = (b0.z - o.y) * iy, t1y = (b0.w - o.y) * iy;
        float t0z = (b2.x - o.z) * iz, t1z = (b2.y - o.z) * iz;
        float n0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
        float f0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
        t0x = (b1.x - o.x) * ix; t1x = (b1.y - o.x) * ix;
        t0y = (b1.z - o.y) * iy; t1y = (b1.w - o.y) * iy;
        t0z = (b2.z - o.z) * iz; t1z = (b2.w - o.z) * iz;
        float n1 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
        float f1 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
        // conservative: one-sided slack on both ends so rounding can only add candidates
        bool h0 = !(n0 * 0.9999995f > fminf(f0 * 1.0000005f, closest));
        bool h1 = !(n1 * 0.9999995f > fminf(f1 * 1.0000005f, closest));
        int c0 = ch.x, c1 = ch.y;
        if (h0 && h1 && n1 < n0) {  // visit the nearer child first
            const int tc = c0; c0 = c1; c1 = tc;
        } else if (!h0) {
            c0 = c1; h0 = h1; h1 = false;
        }
        // c0 = first child to process (if h0), c1 = second (if h1)
        int next = -1;
        if (h0) {
            if (c0 < 0) {
                float t;
                if (tri_test(sc.tris + (~c0), o, d, closest, t)) { closest = t; best = ~c0; }
            } else next = c0;
        }
        if (h1) {
            if (c1 < 0) {
                float t;
                if (tri_test(sc.tris + (~c1), o, d, closest, t)) { closest = t; best = ~c1; }
            } else if (next < 0) next = c1;
            else stack[sp++] = c1;
        }
        if (next >= 0) { node = next; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
    t_hit = closest;
    return best;
}

// ------------------------------------------------------------------------------ path state
struct Path {
    V3 o, d;       // ray about to be traced / incoming direction at the hit
    float hero;
    float pw[SRT_N_WL];
    uint32_t valid;   // valid wavelengths (7, 1 after a refraction, 0 = dead)
    uint32_t bounce;  // scatter events so far in this sample
};

// renderer::get_ray (rendering.cu:66-87) + ray ctor / init_spectrum (ray/ray.cuh:27-58)
__device__ __forceinline__ void camera_ray(const SrtCamera& c, uint32_t i, uint32_t j, Rng& rng, Path& p) {
    const V3 du = mk(c.du[0], c.du[1], c.du[2]), dv = mk(c.dv[0], c.dv[1], c.dv[2]);
    const V3 center = mk(c.center[0], c.center[1], c.center[2]);
    const V3 pixel_center = (mk(c.p00[0], c.p00[1], c.p00[2]) + ((float)i * du)) + ((float)j * dv);
    const float px = -0.5f + rng_uniform(rng);
    const float py = -0.5f + rng_uniform(rng);
    const V3 pixel_sample = pixel_center + ((px * du) + (py * dv));
    V3 origin = center;
    if (!(c.defocus_angle <= 0.0f)) {  // defocus_disk_sample :42-47, random_in_unit_disk vec3.cuh:240-246
        float a, b;
        do {
            a = rng_range(rng, -1, 1);
            b = rng_range(rng, -1, 1);
        } while (!((a * a + b * b + 0.0f * 0.0f) < 1.0f));
        origin = (center + (a * mk(c.disk_u[0], c.disk_u[1], c.disk_u[2]))) + (b * mk(c.disk_v[0], c.disk_v[1], c.disk_v[2]));
    }
    p.o = origin;
    p.d = pixel_sample - origin;
    p.hero = rng_range(rng, 360.0f, 830.0f);
#pragma unroll
    for (int k = 0; k < SRT_N_WL; k++) p.pw[k] = 1.0f;
    p.valid = SRT_N_WL;
    p.bounce = 0;
}

__device__ __forceinline__ void mul_spectrum(Path& p, const float* __restrict__ spec) {  // ray/ray.cuh:60-69
    float wl[SRT_N_WL];
    hero_rotations(p.hero, wl);
#pragma unroll
    for (int k = 0; k < SRT_N_WL; k++)
        if ((uint32_t)k < p.valid) p.pw[k] *= interp95(spec, wl[k]);
}

// dev_spectrum_to_XYZ (color/color.cu:88-104) added into the film accumulator of one pixel
__device__ __forceinline__ void film_add(const SceneRef& sc, const Path& p, float* __restrict__ acc, size_t plane, size_t pix) {
    if (p.valid == 0) return;  // contributes (0,0,0): x + 0 leaves the sum unchanged
    float wl[SRT_N_WL];
    hero_rotations(p.hero, wl);
    const float delta = (830.0f - 360.0f) / 7.0f;
    float x = 0.0f, y = 0.0f, z = 0.0f;
#pragma unroll
    for (int k = 0; k < SRT_N_WL; k++)
        if ((uint32_t)k < p.valid) {
            x += interp95(sc.cie, wl[k]) * p.pw[k] * delta;
            y += interp95(sc.cie + SRT_NS, wl[k]) * p.pw[k] * delta;
            z += interp95(sc.cie + 2 * SRT_NS, wl[k]) * p.pw[k] * delta;
        }
    acc[pix] += x;
    acc[plane + pix] += y;
    acc[2 * plane + pix] += z;
}

__device__ __forceinline__ V3 random_unit_vector(Rng& rng) {  // vec3.cuh:209-227; draws x, y, z in that order (Q13)
    float a, b, c;
    do {
        a = rng_range(rng, -1, 1);
        b = rng_range(rng, -1, 1);
        c = rng_range(rng, -1, 1);
    } while (!((a * a + b * b + c * c) < 1.0f));
    return unit(mk(a, b, c));
}
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return v - ((2 * dot(v, n)) * n); }  // vec3.cuh:179-183

// sellmeier_index (refraction/sellmeier.cu:11-23)
__device__ __forceinline__ float sellmeier(const SrtMaterial* __restrict__ m, float lambda) {
    lambda *= 1e-3f;
    const float l2 = lambda * lambda;
    const float idx = 1.0f + (m->sellB[0] * l2) / (l2 - m->sellC[0]) + (m->sellB[1] * l2) / (l2 - m->sellC[1]) +
                      (m->sellB[2] * l2) / (l2 - m->sellC[2]);
    return sqrtf(idx);
}

// material::scatter (materials/material.cu:55-100) for a non-emissive hit.
// In: p.o = hit point, p.d = incoming direction, tri = the triangle hit.  Out: new ray in p.
// Returns false when the path ends here (metal absorbed the ray).
template <uint32_t MTYPE>
__device__ __forceinline__ bool scatter(const SceneRef& sc, const SrtTri* __restrict__ tri, Path& p, Rng& rng) {
    const float4 q0 = *reinterpret_cast<const float4*>(tri);
    const uint32_t bits = __float_as_uint((reinterpret_cast<const float4*>(tri) + 2)->z);
    const SrtMaterial* m = sc.mats + SRT_TRI_MAT(bits);
    V3 n = mk(q0.x, q0.y, q0.z);
    const bool front = dot(p.d, n) < 0;  // hit_record::set_face_normal, primitives/hit_record.cuh:30-43
    if (!front) n = -n;
    const V3 uin = unit(p.d);
    V3 out;
    float eps_sign = 1.0f;
    bool alive = true;
    if (MTYPE == SRT_METALLIC) {  // reflection_scatter :22-37
        const V3 refl = reflect(uin, n);
        out = refl + (m->fuzz * random_unit_vector(rng));
        alive = dot(out, n) > 0;
        if (!alive) p.valid = 0;
    } else if (MTYPE == SRT_DIELECTRIC) {  // refraction_scatter :103-135, evaluated at the hero wavelength only
        const float ir = sellmeier(m, p.hero);
        const float ratio = front ? (1.0f / ir) : ir;
        const float cos_theta = fminf(dot(-uin, n), 1.0f);
        const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
        bool cannot = ratio * sin_theta > 1.0f;
        if (!cannot) {  // `||` short-circuits the RNG draw (Q12)
            float r0 = (1.0f - ratio) / (1.0f + ratio);  // reflectance (Schlick) :39-53
            r0 = r0 * r0;
            const float refl = r0 + (1.0f - r0) * powf(1.0f - cos_theta, 5.0f);
            cannot = refl > rng_uniform(rng);
        }
        if (cannot) out = reflect(uin, n);
        else {  // refract, vec3.cuh:198-205
            const float ct = fminf(dot(-uin, n), 1.0f);
            const V3 perp = ratio * (uin + (ct * n));
            const V3 par = (-sqrtf(fabsf(1.0f - len2(perp)))) * n;
            out = perp + par;
            eps_sign = -1.0f;
            p.valid = 1;  // only the hero wavelength survives a refraction (Q5)
        }
    } else {  // lambertian_scatter :9-19
        out = n + random_unit_vector(rng);
        const float s = 1e-8f;
        if ((fabsf(out.x) < s) && (fabsf(out.y) < s) && (fabsf(out.z) < s)) out = n;
    }
    mul_spectrum(p, m->spec);
    p.o = p.o + ((eps_sign * SRT_EPSILON) * n);
    p.d = out;
    return alive;
}

// What happened to the ray that was just traced
enum : int { EV_DONE = -1 };  // sample finished; otherwise the value is the queue index 1..3 (lambert, metal, dielectric)

// trace p's ray and either finish the sample or leave p at the hit (p.o = hit point).
// Returns EV_DONE or the material queue (1 lambertian, 2 metallic, 3 dielectric); tri_out = leaf-order index.
__device__ __forceinline__ int extend(const SceneRef& sc, const WaveParams& P, Path& p, int& tri_out, float* acc, size_t pix) {
    float t = 0.f;
    const int tri = closest_hit(sc, p.o, p.d, t);
    if (tri < 0) {  // miss: ray_bounce, rendering.cu:24-27
        if (!P.bg_is_zero) {
            mul_spectrum(p, sc.bg);
            film_add(sc, p, acc, P.plane, pix);
        }
        return EV_DONE;
    }
    const uint32_t bits = __float_as_uint((reinterpret_cast<const float4*>(sc.tris + tri) + 2)->z);
    const uint32_t mtype = SRT_TRI_MTYPE(bits);
    if (mtype == SRT_EMISSIVE) {  // scatter() returns false after multiplying by the emission spectrum
        mul_spectrum(p, sc.mats[SRT_TRI_MAT(bits)].spec);
        film_add(sc, p, acc, P.plane, pix);
        return EV_DONE;
    }
    p.o = mk(p.o.x + t * p.d.x, p.o.y + t * p.d.y, p.o.z + t * p.d.z);  // ray::at, ray/ray.cuh:44-47
    tri_out = tri;
    return mtype == SRT_METALLIC ? 2 : (mtype == SRT_DIELECTRIC ? 3 : 1);
}

// ------------------------------------------------------------------------------ shared-memory scene
// Small scenes (all three reference scenes) are staged once per block: nodes | tris | mats | cie | bg
template <bool SMEM>
__device__ __forceinline__ SceneRef load_scene(const WaveParams& P, unsigned char* smem) {
    SceneRef sc;
    sc.n_tris = P.n_tris;
    if (!SMEM) {
        sc.nodes = P.nodes; sc.tris = P.tris; sc.mats = P.mats; sc.cie = P.cie; sc.bg = P.bg;
        return sc;
    }
    const int n_nodes = P.n_tris > 1 ? P.n_tris - 1 : 0;
    float4* dst = reinterpret_cast<float4*>(smem);
    const int v_nodes = n_nodes * (int)(sizeof(SrtNode) / 16), v_tris = P.n_tris * (int)(sizeof(SrtTri) / 16),
              v_mats = P.n_mats * (int)(sizeof(SrtMaterial) / 16);
    for (int i = threadIdx.x; i < v_nodes; i += blockDim.x) dst[i] = reinterpret_cast<const float4*>(P.nodes)[i];
    for (int i = threadIdx.x; i < v_tris; i += blockDim.x) dst[v_nodes + i] = reinterpret_cast<const float4*>(P.tris)[i];
    for (int i = threadIdx.x; i < v_mats; i += blockDim.x) dst[v_nodes + v_tris + i] = reinterpret_cast<const float4*>(P.mats)[i];
    float* f = reinterpret_cast<float*>(dst + v_nodes + v_tris + v_mats);
    for (int i = threadIdx.x; i < 3 * SRT_NS; i += blockDim.x) f[i] = P.cie[i];
    for (int i = threadIdx.x; i < SRT_NS; i += blockDim.x) f[3 * SRT_NS + i] = P.bg[i];
    __syncthreads();
    sc.nodes = reinterpret_cast<const SrtNode*>(dst);
    sc.tris = reinterpret_cast<const SrtTri*>(dst + v_nodes);
    sc.mats = reinterpret_cast<const SrtMaterial*>(dst + v_nodes + v_tris);
    sc.cie = f;
    sc.bg = f + 3 * SRT_NS;
    return sc;
}

// ------------------------------------------------------------------------------ slot <-> pixel
__device__ __forceinline__ void slot_pixel(const WaveParams& P, uint32_t slot, uint32_t& ci, uint32_t& cj) {
    cj = slot / P.slot_w;
    ci = slot - cj * P.slot_w;
}
// the reference's per-thread seed index (rendering.cu:125-137, 28x16 blocks, grid from the nominal chunk size)
__device__ __forceinline__ uint32_t ref_thread_index(const WaveParams& P, uint32_t ci, uint32_t cj) {
    return (cj % 16u) * 28u + (ci % 28u) + 448u * ((cj / 16u) * P.ref_grid_x + ci / 28u);
}

// warp-aggregated queue push: one atomicAdd per warp per queue
__device__ __forceinline__ void queue_push(uint32_t* __restrict__ q, uint32_t* counter, bool pred, uint32_t slot) {
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == __ffs(mask) - 1) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
    if (pred) q[base + __popc(mask & ((1u << lane) - 1))] = slot;
}

__device__ __forceinline__ void store_hit_state(const WaveParams& P, uint32_t slot, const Path& p, int tri) {
    P.R0[slot] = make_float4(p.o.x, p.o.y, p.o.z, __int_as_float(tri));
    P.R1[slot] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(p.valid | (p.bounce << 3)));
    P.P0[slot] = make_float4(p.pw[0], p.pw[1], p.pw[2], p.pw[3]);
    P.P1[slot] = make_float4(p.pw[4], p.pw[5], p.pw[6], p.hero);
}
__device__ __forceinline__ void store_rng(const WaveParams& P, uint32_t slot, const Rng& r) {
    P.G0[slot] = make_uint4(r.d, r.v0, r.v1, r.v2);
    P.G1[slot] = make_uint2(r.v3, r.v4);
}
__device__ __forceinline__ Rng load_rng(const WaveParams& P, uint32_t slot) {
    const uint4 a = P.G0[slot];
    const uint2 b = P.G1[slot];
    Rng r;
    r.d = a.x; r.v0 = a.y; r.v1 = a.z; r.v2 = a.w; r.v3 = b.x; r.v4 = b.y;
    return r;
}

// ------------------------------------------------------------------------------ kernels
__global__ void k_init_slots(WaveParams P) {  // init_random_states (rendering.cu:120-138), one state per reference thread slot
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.nslots) return;
    uint32_t ci, cj;
    slot_pixel(P, slot, ci, cj);
    store_rng(P, slot, rng_seed(1984u + ref_thread_index(P, ci, cj)));
}

// first regenerate queue of a chunk: every pixel of the chunk this rank owns, samples reset
__global__ void k_begin_chunk(WaveParams P) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool mine = false;
    if (slot < P.nslots) {
        uint32_t ci, cj;
        slot_pixel(P, slot, ci, cj);
        if (ci < P.cw && cj < P.ch) {
            const uint32_t x = P.off_x + ci, y = P.off_y + cj;
            const uint32_t tile = (x / P.tile_w) + (y / P.tile_h) * P.tiles_x;
            mine = (tile % P.world) == P.rank;
            if (mine) P.sidx[slot] = 0;
        }
    }
    queue_push(P.qr_out, P.cnt_out + 0, mine, slot);
}

template <bool SMEM>
__global__ void __launch_bounds__(SRT_BLOCK) k_generate(WaveParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneRef sc = load_scene<SMEM>(P, smem);
    const uint32_t n = P.cnt_in[0];
    unsigned long long rays = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t idx = base + threadIdx.x;
        const bool active = idx < n;
        int ev = EV_DONE;
        bool requeue = false;
        uint32_t slot = 0;
        if (active) {
            slot = P.qr_in[idx];
            uint32_t ci, cj;
            slot_pixel(P, slot, ci, cj);
            const size_t pix = (size_t)(P.off_y + cj) * P.img_w + (P.off_x + ci);
            Rng rng = load_rng(P, slot);
            uint32_t s = P.sidx[slot];
            Path p;
            int tri = -1;
            int loops = 0;
            while (s < P.spp) {
                if (loops == P.regen_loop) { requeue = true; break; }
                loops++;
                camera_ray(P.cam, P.off_x + ci, P.off_y + cj, rng, p);
                s++;
                if (P.bounce_limit == 0) { continue; }  // loop body of ray_bounce never runs: valid = 0
                rays++;
                ev = extend(sc, P, p, tri, P.acc, pix);
                if (ev != EV_DONE) break;
            }
            P.sidx[slot] = s;
            store_rng(P, slot, rng);
            if (ev != EV_DONE) store_hit_state(P, slot, p, tri);
        }
        queue_push(P.qr_out, P.cnt_out + 0, requeue, slot);
        queue_push(P.qm_out, P.cnt_out + 1, ev == 1, slot);
        queue_push(P.qm_out + P.nslots, P.cnt_out + 2, ev == 2, slot);
        queue_push(P.qm_out + 2 * (size_t)P.nslots, P.cnt_out + 3, ev == 3, slot);
    }
    if (P.ray_counter && rays) atomicAdd(P.ray_counter, rays);
}

template <uint32_t MTYPE>
__device__ __forceinline__ int shade_one(const SceneRef& sc, const WaveParams& P, uint32_t slot, unsigned long long& rays) {
    uint32_t ci, cj;
    slot_pixel(P, slot, ci, cj);
    const size_t pix = (size_t)(P.off_y + cj) * P.img_w + (P.off_x + ci);
    const float4 r0 = P.R0[slot], r1 = P.R1[slot], p0 = P.P0[slot], p1 = P.P1[slot];
    Path p;
    p.o = mk(r0.x, r0.y, r0.z);
    p.d = mk(r1.x, r1.y, r1.z);
    const uint32_t meta = __float_as_uint(r1.w);
    p.valid = meta & 7u;
    p.bounce = meta >> 3;
    p.pw[0] = p0.x; p.pw[1] = p0.y; p.pw[2] = p0.z; p.pw[3] = p0.w;
    p.pw[4] = p1.x; p.pw[5] = p1.y; p.pw[6] = p1.z;
    p.hero = p1.w;
    int tri = __float_as_int(r0.w);
    Rng rng = load_rng(P, slot);
    const bool alive = scatter<MTYPE>(sc, sc.tris + tri, p, rng);
    store_rng(P, slot, rng);
    p.bounce++;
    if (!alive || p.bounce >= P.bounce_limit) return EV_DONE;  // absorbed, or bounce limit: valid = 0 (rendering.cu:38)
    rays++;
    const int ev = extend(sc, P, p, tri, P.acc, pix);
    if (ev != EV_DONE) store_hit_state(P, slot, p, tri);
    return ev;
}

// one launch covers the three material segments; each segment is padded to a warp multiple so a
// warp never mixes materials
template <bool SMEM>
__global__ void __launch_bounds__(SRT_BLOCK) k_shade(WaveParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneRef sc = load_scene<SMEM>(P, smem);
    const uint32_t nL = P.cnt_in[1], nM = P.cnt_in[2], nD = P.cnt_in[3];
    const uint32_t eL = (nL + 31u) & ~31u, eM = eL + ((nM + 31u) & ~31u), eD = eM + ((nD + 31u) & ~31u);
    unsigned long long rays = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < eD; base += gridDim.x * blockDim.x) {
        const uint32_t v = base + threadIdx.x;
        int ev = EV_DONE;
        bool done = false;
        uint32_t slot = 0;
        if (v < eL) {
            if (v < nL) { slot = P.qm_in[v]; ev = shade_one<SRT_LAMBERTIAN>(sc, P, slot, rays); done = ev == EV_DONE; }
        } else if (v < eM) {
            const uint32_t k = v - eL;
            if (k < nM) { slot = P.qm_in[P.nslots + k]; ev = shade_one<SRT_METALLIC>(sc, P, slot, rays); done = ev == EV_DONE; }
        } else if (v < eD) {
            const uint32_t k = v - eM;
            if (k < nD) { slot = P.qm_in[2 * (size_t)P.nslots + k]; ev = shade_one<SRT_DIELECTRIC>(sc, P, slot, rays); done = ev == EV_DONE; }
        }
        queue_push(P.qr_out, P.cnt_out + 0, done, slot);
        queue_push(P.qm_out, P.cnt_out + 1, ev == 1, slot);
        queue_push(P.qm_out + P.nslots, P.cnt_out + 2, ev == 2, slot);
        queue_push(P.qm_out + 2 * (size_t)P.nslots, P.cnt_out + 3, ev == 3, slot);
    }
    if (P.ray_counter && rays) atomicAdd(P.ray_counter, rays);
}

// per-pixel persistent kernel: same device functions, no queues (cross-check / comparison)
template <bool SMEM>
__global__ void __launch_bounds__(SRT_BLOCK) k_megakernel(WaveParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneRef sc = load_scene<SMEM>(P, smem);
    unsigned long long rays = 0;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < P.nslots; slot += gridDim.x * blockDim.x) {
        uint32_t ci, cj;
        slot_pixel(P, slot, ci, cj);
        if (ci >= P.cw || cj >= P.ch) continue;
        const uint32_t x = P.off_x + ci, y = P.off_y + cj;
        if (((x / P.tile_w) + (y / P.tile_h) * P.tiles_x) % P.world != P.rank) continue;
        const size_t pix = (size_t)y * P.img_w + x;
        Rng rng = load_rng(P, slot);
        for (uint32_t s = 0; s < P.spp; s++) {
            Path p;
            camera_ray(P.cam, x, y, rng, p);
            if (P.bounce_limit == 0) continue;
            int tri = -1;
            rays++;
            int ev = extend(sc, P, p, tri, P.acc, pix);
            while (ev != EV_DONE) {
                bool alive;
                if (ev == 2) alive = scatter<SRT_METALLIC>(sc, sc.tris + tri, p, rng);
                else if (ev == 3) alive = scatter<SRT_DIELECTRIC>(sc, sc.tris + tri, p, rng);
                else alive = scatter<SRT_LAMBERTIAN>(sc, sc.tris + tri, p, rng);
                p.bounce++;
                if (!alive || p.bounce >= P.bounce_limit) break;
                rays++;
                ev = extend(sc, P, p, tri, P.acc, pix);
            }
        }
        store_rng(P, slot, rng);
    }
    if (P.ray_counter && rays) atomicAdd(P.ray_counter, rays);
}

// film: XYZ sum / spp -> sRGB 0..255 (save_to_fb rendering.cu:140-149, color.cu:15-49), raster order
__global__ void k_resolve(const float* __restrict__ acc, size_t plane, uint32_t img_w, uint32_t off_x, uint32_t off_y, uint32_t w, uint32_t h,
                          uint32_t spp, float* __restrict__ out_rgb, float* __restrict__ out_xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const uint32_t cy = i / w, cx = i - cy * w;
    const size_t pix = (size_t)(off_y + cy) * img_w + (off_x + cx);
    const float inv = 1 / (float)spp;  // pixel_color / float(spp) = (1/spp) * v
    const float X = inv * acc[pix], Y = inv * acc[plane + pix], Z = inv * acc[2 * plane + pix];
    const float m[9] = {3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f, 0.0415560f, 0.0556434f, -0.2040259f, 1.0572252f};
    const float lin[3] = {(m[0] * X) + (m[1] * Y) + (m[2] * Z), (m[3] * X) + (m[4] * Y) + (m[5] * Z), (m[6] * X) + (m[7] * Y) + (m[8] * Z)};
    const size_t n = (size_t)w * h;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float v = lin[c];
        const float g = v < 0.0f ? 0.0f : (v < 0.0031308f ? 12.92f * v : (v < 1.0f ? ((1.055f * powf(v, 0.416666f)) - 0.055f) : 1.0f));
        out_rgb[c * n + i] = (float)(int)(g * 255.99f);
    }
    out_xyz[i] = X;
    out_xyz[n + i] = Y;
    out_xyz[2 * n + i] = Z;
}

// standalone closest-hit queries (BASELINE.json configs[3]); always global-memory scene
__global__ void __launch_bounds__(SRT_BLOCK) k_trace_rays(WaveParams P, uint32_t n, const float* __restrict__ o, const float* __restrict__ d,
                                                          const uint32_t* __restrict__ sorted_idx, float* __restrict__ t_out,
                                                          int32_t* __restrict__ tri_out) {
    SceneRef sc;
    sc.nodes = P.nodes; sc.tris = P.tris; sc.mats = P.mats; sc.cie = nullptr; sc.bg = nullptr; sc.n_tris = P.n_tris;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float t = 0.f;
        const int tri = closest_hit(sc, mk(o[3ull * i], o[3ull * i + 1], o[3ull * i + 2]), mk(d[3ull * i], d[3ull * i + 1], d[3ull * i + 2]), t);
        t_out[i] = tri >= 0 ? t : -1.0f;
        tri_out[i] = tri >= 0 ? (int32_t)sorted_idx[tri] : -1;
    }
}

// ------------------------------------------------------------------------------ launchers
static size_t scene_smem_bytes(const WaveParams& P) {
    const size_t n_nodes = P.n_tris > 1 ? P.n_tris - 1 : 0;
    return n_nodes * sizeof(SrtNode) + (size_t)P.n_tris * sizeof(SrtTri) + (size_t)P.n_mats * sizeof(SrtMaterial) + 4 * SRT_NS * sizeof(float);
}

LaunchTable make_launch_table() {
    LaunchTable t;
    t.smem_bytes = [](const WaveParams& P) { return scene_smem_bytes(P); };
    t.configure = [](size_t bytes) {
        cudaError_t e = cudaSuccess;
        if (bytes > 48 * 1024) {
            e = cudaFuncSetAttribute(k_generate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_shade<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_megakernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        }
        return e;
    };
    t.init_slots = [](const WaveParams& P, cudaStream_t st) { k_init_slots<<<(P.nslots + 255) / 256, 256, 0, st>>>(P); };
    t.begin_chunk = [](const WaveParams& P, cudaStream_t st) { k_begin_chunk<<<(P.nslots + 255) / 256, 256, 0, st>>>(P); };
    t.generate = [](const WaveParams& P, int grid, size_t smem, cudaStream_t st) {
        if (smem) k_generate<true><<<grid, SRT_BLOCK, smem, st>>>(P);
        else k_generate<false><<<grid, SRT_BLOCK, 0, st>>>(P);
    };
    t.shade = [](const WaveParams& P, int grid, size_t smem, cudaStream_t st) {
        if (smem) k_shade<true><<<grid, SRT_BLOCK, smem, st>>>(P);
        else k_shade<false><<<grid, SRT_BLOCK, 0, st>>>(P);
    };
    t.megakernel = [](const WaveParams& P, int grid, size_t smem, cudaStream_t st) {
        if (smem) k_megakernel<true><<<grid, SRT_BLOCK, smem, st>>>(P);
        else k_megakernel<false><<<grid, SRT_BLOCK, 0, st>>>(P);
    };
    t.resolve = [](const float* acc, size_t plane, uint32_t img_w, uint32_t ox, uint32_t oy, uint32_t w, uint32_t h, uint32_t spp, float* rgb,
                   float* xyz, cudaStream_t st) {
        k_resolve<<<(w * h + 255) / 256, 256, 0, st>>>(acc, plane, img_w, ox, oy, w, h, spp, rgb, xyz);
    };
    t.trace_rays = [](const WaveParams& P, uint32_t n, const float* o, const float* d, const uint32_t* sorted_idx, float* t_out, int32_t* tri_out,
                      int grid, cudaStream_t st) { k_trace_rays<<<grid, SRT_BLOCK, 0, st>>>(P, n, o, d, sorted_idx, t_out, tri_out); };
    return t;
}

}  // namespace SRT_FP_NS
}  // namespace srt
