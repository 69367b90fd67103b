// Device code of the spectral path tracer.  Included twice -- by trace_fast.cu (default nvcc
// flags: FMA contraction, as the reference's own nvcc build gets) and by trace_strict.cu
// (-fmad=false: rounds exactly like the host oracle) -- inside namespace SRT_FP_NS.
//
// Pipelines
//   wavefront (default): k_wavefront, ONE persistent-block kernel per round of samples.  Every block
//     keeps block_slots paths in flight in four shared-memory queues (regenerate | lambertian |
//     metallic | dielectric, warp-uniform work) and loops over passes: regenerate = camera ray ->
//     closest hit -> classify; material queues = scatter -> closest hit -> classify.  "classify"
//     finishes the sample on miss / emitter / absorption / bounce limit (XYZ added to the film) and
//     pushes the slot to the regenerate queue, or to the queue of the material it hit.  Queue pushes
//     are warp-aggregated (ballot + one atomicAdd per warp).  Exactly one sample is in flight per
//     pixel, so every pixel consumes its XORWOW stream in the reference's order
//     (rendering/rendering.cu:215-228).
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

#ifdef SRT_PHASE_CLOCKS
// probe build: cycles of warp 0 / block 0 between fine-grained points, summed over the launch (g_fine[2k] = cycles, [2k+1] = visits)
__device__ unsigned long long g_fine[32];
#define SRT_FINE_BEGIN unsigned long long fine_t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(fine_t_))
#define SRT_FINE(k) do { unsigned long long n_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(n_)); \
        if (blockIdx.x == 0 && threadIdx.x == 0) { g_fine[2 * (k)] += n_ - fine_t_; g_fine[2 * (k) + 1] += 1; } fine_t_ = n_; } while (0)
#else
#define SRT_FINE_BEGIN do { } while (0)
#define SRT_FINE(k) do { } while (0)
#endif

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
    const SrtWide* nodes;
    const float4* grid;   // quantisation grid of the node boxes (global memory): lo.xyz, 1 / cell
    bool nodes_global;    // nodes in global memory (one LDG.256 per node) or staged in shared memory (two 16-byte loads)
    const SrtTri* tris;
    const SrtFlatUnit* units;  // wide leaf only
    int n_units;
    float flat_guard, flat_tol;
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
// The plane part of tri::hit (primitives/tri.cu:10-16): denom = n.d and the numerator D - n.o of t.  ONE definition for the
// exact test and for the wide-leaf pre-test, so that both see the same bits whatever the compiler contracts into FMAs.
__device__ __forceinline__ float plane_denom(float4 q, V3 d) { return q.x * d.x + q.y * d.y + q.z * d.z; }
__device__ __forceinline__ float plane_num(float4 q, V3 o) { return q.w - (q.x * o.x + q.y * o.y + q.z * o.z); }
// tri::hit (primitives/tri.cu:3-45) on the packed 48-B triangle; returns t through t_out.
__device__ __forceinline__ bool tri_test(const SrtTri* __restrict__ tp, V3 o, V3 d, float closest, float& t_out) {
    const float4 q0 = *reinterpret_cast<const float4*>(tp);
    const float denom = plane_denom(q0, d);
    if (fabsf(denom) < 1e-8f) return false;
    const float t = plane_num(q0, o) / denom;
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

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// equal-t ties go to the triangle the reference would have tested last (SrtTri::prio, host/ref_order.cpp)
__device__ __forceinline__ void consider_hit(const SrtTri* __restrict__ tris, int i, V3 o, V3 d, float& closest, int& best, uint32_t& best_prio) {
    float t;
    if (!tri_test(tris + i, o, d, closest, t)) return;
    const uint32_t prio = tris[i].prio;
    if (t < closest || best < 0 || prio > best_prio) { closest = t; best = i; best_prio = prio; }
}

// Wide-leaf closest hit for scenes that collapse into <= 32 pre-test units (all three reference
// scenes): the whole LBVH is ONE leaf, so there is no tree walk, no stack and no divergent descent.
//   phase 1: every lane runs the same branch-free loop over the units (unit data is a warp-uniform
//            shared-memory broadcast).  A unit is a triangle or a parallelogram pair (two coplanar
//            triangles sharing a diagonal, i.e. every quad the reference builds): plane distance with
//            an approximate reciprocal, the point in the unit's affine frame, then "certainly outside"
//            tests with the error budget folded into the constants; survivors set bits 2u / 2u+1;
//   phase 2: the few survivors (the triangles the ray really pierces, ~1-4) get the exact
//            reference arithmetic (tri_test), nearest wins, ties by reference test order.
// Phase 1 only ever rejects pairs the exact test rejects too, so the result equals testing every
// triangle exactly -- which is what "closest hit" means in the reference (bvh.cu:98-166).
template <uint32_t BIT>
__device__ __forceinline__ void flat_unit_test(const float4* __restrict__ up, int u, V3 o, V3 d, float gd, float tol, uint32_t& mask) {
    const float4 pl = up[4 * u], A = up[4 * u + 1], B = up[4 * u + 2], C = up[4 * u + 3];
    // the SAME two expressions the exact test evaluates (plane_denom / plane_num): for the unit's head triangle -- and for
    // a partner whose plane is bit-identical -- numerator and denominator carry the reference's own rounding, and the
    // approximate t below differs from the reference's quotient by the reciprocal's 2^-22 only, at ANY incidence angle
    const float denom = plane_denom(pl, d);
    const float num = plane_num(pl, o);
    const float t = num * rcp_approx(denom);
    const float px = __fmaf_rn(t, d.x, o.x), py = __fmaf_rn(t, d.y, o.y), pz = __fmaf_rn(t, d.z, o.z);
    const float al = __fmaf_rn(A.z, pz, __fmaf_rn(A.y, py, __fmaf_rn(A.x, px, A.w)));
    const float be = __fmaf_rn(B.z, pz, __fmaf_rn(B.y, py, __fmaf_rn(B.x, px, B.w)));
    const float s = al + be;
    // "Certainly outside" per half as one chain of compare-and-combine predicate instructions:
    //   behind = t < 0 & |num| > tol
    //   out_i  = behind | alpha' < 0 | beta' < 0 | alpha'+beta' > c1       (first half)
    //   out_j  = behind | alpha' > c2 | beta' > c2 | alpha'+beta' < c3     (second half)
    //   unsure = |num| < near & |denom| < gd                              (second half kept without a verdict)
    // `unsure` only exists for pairs whose second triangle stores a plane that differs from the head's in the last bits
    // (faces of a rotated box; near = 0 everywhere else): the partner's own numerator / denominator round differently,
    // which moves its hit point by error / cos(incidence).  A grazing ray (|denom| < gd = guard |d|) can only hit inside
    // the scene when it starts next to the plane (|num| < near); host/flat_leaf.cpp derives both constants.
    // Every comparison is false for a NaN operand, so NaN / inf can only ever keep a candidate (the exact test
    // decides).  Written in PTX because the compiler otherwise builds the two mask bits through chains of selects
    // (twice the instructions).  ONE loop body for all units: a second, guard-free copy of it costs more in
    // instruction-cache misses (measured: +7 %) than the three extra instructions do (+4 %).
    asm("{\n\t"
        ".reg .pred pb, pi, pj, pg;\n\t"
        ".reg .f32 an, ad;\n\t"
        "abs.f32 an, %2;\n\t"
        "abs.f32 ad, %12;\n\t"
        "setp.lt.f32 pb, %1, 0f00000000;\n\t"
        "setp.gt.and.f32 pb, an, %3, pb;\n\t"
        "setp.lt.or.f32 pi, %4, 0f00000000, pb;\n\t"
        "setp.gt.or.f32 pj, %4, %8, pb;\n\t"
        "setp.lt.or.f32 pi, %5, 0f00000000, pi;\n\t"
        "setp.gt.or.f32 pj, %5, %8, pj;\n\t"
        "setp.gt.or.f32 pi, %6, %7, pi;\n\t"
        "setp.lt.or.f32 pj, %6, %9, pj;\n\t"
        "setp.ge.f32 pg, an, %14;\n\t"
        "setp.ge.or.f32 pg, ad, %13, pg;\n\t"
        "@!pi or.b32 %0, %0, %10;\n\t"
        "@!pj or.b32 %0, %0, %11;\n\t"
        "@!pg or.b32 %0, %0, %11;\n\t"
        "}"
        : "+r"(mask)
        : "f"(t), "f"(num), "f"(tol), "f"(al), "f"(be), "f"(s), "f"(C.x), "f"(C.y), "f"(C.z), "n"(BIT), "n"(BIT << 1), "f"(denom), "f"(gd), "f"(C.w));
}
// phase 1: the candidate mask of one ray (bit 2u / 2u+1: first / second triangle of unit u)
__device__ __forceinline__ unsigned long long flat_candidates(const SceneRef& sc, V3 o, V3 d) {
    const float4* __restrict__ up = reinterpret_cast<const float4*>(sc.units);
    // groups of four units: inside a group the candidate bits are compile-time constants (predicated ORs
    // of immediates), one shift per group places them in the 64-bit mask; the <= 3 left-over units follow
    unsigned long long m = 0;
    const int n = sc.n_units, groups = n >> 2;
    float gd;  // grazing threshold of this ray: guard * |d|, |d| to a relative 2^-22 and rounded up
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(gd) : "f"(len2(d)));
    gd *= sc.flat_guard * 1.0001f;
    const float tol = sc.flat_tol;
    for (int g = 0; g < groups; g++) {
        uint32_t loc = 0;
        flat_unit_test<1u>(up, 4 * g, o, d, gd, tol, loc);
        flat_unit_test<4u>(up, 4 * g + 1, o, d, gd, tol, loc);
        flat_unit_test<16u>(up, 4 * g + 2, o, d, gd, tol, loc);
        flat_unit_test<64u>(up, 4 * g + 3, o, d, gd, tol, loc);
        m |= (unsigned long long)loc << (8 * g);
    }
#pragma unroll 1
    for (int u = 4 * groups; u < n; u++) {
        uint32_t loc = 0;
        flat_unit_test<1u>(up, u, o, d, gd, tol, loc);
        m |= (unsigned long long)loc << (2 * u);
    }
    return m;
}
// phase 2, one ray per lane on its own: the megakernel, whose lanes sit in divergent loops
__device__ __forceinline__ int closest_hit_flat(const SceneRef& sc, V3 o, V3 d, float& t_hit) {
    SRT_FINE_BEGIN;
    const unsigned long long m = flat_candidates(sc, o, d);
    uint32_t m0 = (uint32_t)m, m1 = (uint32_t)(m >> 32);
    SRT_FINE(0);
    float closest = FLT_MAX;
    int best = -1;
    uint32_t best_prio = 0;
    while (m0 | m1) {
        int i;
        if (m0) { i = __ffs(m0) - 1; m0 &= m0 - 1; }
        else { i = 32 + __ffs(m1) - 1; m1 &= m1 - 1; }
        consider_hit(sc.tris, i, o, d, closest, best, best_prio);
    }
    SRT_FINE(1);
    t_hit = closest;
    return best;
}

// tri::hit on a triangle whose three vectors are already in registers (the walk below loads them next to the node)
__device__ __forceinline__ bool tri_test_q(float4 q0, float4 q1, float4 q2, V3 o, V3 d, float closest, float& t_out) {
    const float denom = plane_denom(q0, d);
    if (fabsf(denom) < 1e-8f) return false;
    const float t = plane_num(q0, o) / denom;
    if (!(0.0f <= t && t <= closest)) return false;
    const uint32_t bits = __float_as_uint(q2.z);
    const float px = o.x + t * d.x, py = o.y + t * d.y, pz = o.z + t * d.z;
    const float pw = sel3(px, py, pz, SRT_TRI_WAX(bits)), ph = sel3(px, py, pz, SRT_TRI_HAX(bits));
    const float a1 = (pw - q1.z) * (q1.y - q1.w) - (q1.x - q1.z) * (ph - q1.w);
    const float a2 = (pw - q2.x) * (q1.w - q2.y) - (q1.z - q2.x) * (ph - q2.y);
    const float a3 = (pw - q1.x) * (q2.y - q1.y) - (q2.x - q1.x) * (ph - q1.y);
    const bool inside = SRT_TRI_CW(bits) ? (a1 >= 0.f && a2 >= 0.f && a3 >= 0.f) : (a1 <= 0.f && a2 <= 0.f && a3 <= 0.f);
    if (!inside) return false;
    t_out = t;
    return true;
}

// ---- the LBVH walk: 32-byte grid nodes, deferred and warp-batched leaf tests, stack in shared memory ----
// A lane carries an internal node to open (`node`, -1 = none), one leaf whose triangle is still to be tested (`pend`,
// -1 = none) and a stack of further entries (>= 0 internal node, < 0 leaf = ~triangle) whose top lives in a register:
// a pop hands out the register and issues the load of the entry below, which nobody waits for before the next pop.
// One step = (1) a lane without a node takes the top of its stack; (2) lanes with a node load its 32 B (both child
// boxes live in the parent: one LDG.256), test the two boxes and route the hit children -- the nearer internal one is
// the next node, the farther goes on the stack, a leaf becomes the pending triangle (or goes on the stack when one is
// pending already); (3) pending triangles are tested TOGETHER: only when SRT_LEAF_BATCH lanes of the warp have one, or
// when a lane has nothing else left to do -- tested on the spot, 2.5 of 32 lanes ran the exact triangle arithmetic (ncu).
// The closest hit does not depend on the order of the tests (equal t: the highest `prio` wins), so the answer equals
// the plain near-first walk's; deferring only tests boxes against a `closest` that is a few steps old (more visits,
// never fewer).  Returns false when the lane's walk is over.  `lanes` = the lanes that call this step together.
#define SRT_STACK_EMPTY 0x7fffffff
// a 4-wide step stacks up to three entries and descends two binary levels; the binary tree over 30-bit codes + index bits is at
// most ~54 deep: 3 * 27 + 1 entries
#define SRT_STACK_MAX 96
#ifndef SRT_STACK_SMEM
#define SRT_STACK_SMEM 24  // stack entries per thread in shared memory (k_trace_rays: 24 KB per block of 256)
#endif
#ifndef SRT_LEAF_BATCH
#define SRT_LEAF_BATCH 8
#endif
struct Walk {
    int node, pend, top, sp;
    float closest;
    int best;
    uint32_t best_prio;
};
// The ray in the grid space of the node boxes (srt_types.h): origin g(o) + 2^23, snapped to the integer lattice by that very
// addition (error <= half a cell, inside the boxes' three-cell margin), and 1 / (d / cell).  2^23 + q is the float whose low
// mantissa bits are the 16-bit coordinate q, so a box plane becomes a float with ONE byte permute (no int -> float
// conversions) and `plane - origin` is an exact integer difference; the slab parameters are those of the world-space ray.
// Origins far outside the grid make the snap a relative error (<= 1.2e-7 of t), which the slack on the slab interval covers.
struct GridRay { V3 om, inv; };
__device__ __forceinline__ GridRay grid_ray(const SceneRef& sc, V3 o, V3 d) {
    const float4 G = __ldg(sc.grid);
    GridRay g;
    const float shift = SRT_GRID_OFFSET + 8388608.0f;
    g.om = mk((o.x - G.x) * G.w + shift, (o.y - G.y) * G.w + shift, (o.z - G.z) * G.w + shift);
    g.inv = mk(1.0f / (d.x * G.w), 1.0f / (d.y * G.w), 1.0f / (d.z * G.w));
    return g;
}
__device__ __forceinline__ float grid_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)); }  // 2^23 + (w & 0xffff)
__device__ __forceinline__ float grid_hi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); }  // 2^23 + (w >> 16)
__device__ __forceinline__ void walk_reset(Walk& w, bool active) {
    w.node = active ? 0 : -1; w.pend = -1; w.top = SRT_STACK_EMPTY; w.sp = 0;
    w.closest = FLT_MAX; w.best = -1; w.best_prio = 0;
}
// The entries below the cached top: the first SRT_STACK_SMEM of a lane live in shared memory, entry k of thread t at
// word k * blockDim + t -- every lane in its own bank whatever its depth, so a push or pop of a whole warp is one
// wavefront (the same stack in local memory: one L1 wavefront per distinct depth, next to the node fetches) -- deeper
// entries (trees over heavily duplicated Morton codes) spill to local memory.  sm == null: local memory only.
struct StackRef {
    int* sm;        // this thread's column of the shared-memory stack (null: local memory only)
    int* lm;        // local-memory overflow
    uint32_t stride;
};
__device__ __forceinline__ void walk_push(Walk& w, const StackRef& st, int e) {
    if (w.top != SRT_STACK_EMPTY) {
        if (st.sm && w.sp < SRT_STACK_SMEM) st.sm[w.sp * st.stride] = w.top;
        else st.lm[st.sm ? w.sp - SRT_STACK_SMEM : w.sp] = w.top;
        w.sp++;
    }
    w.top = e;
}
__device__ __forceinline__ int walk_pop(Walk& w, const StackRef& st) {
    const int e = w.top;
    w.top = SRT_STACK_EMPTY;
    if (w.sp > 0) {
        --w.sp;
        if (st.sm && w.sp < SRT_STACK_SMEM) w.top = st.sm[w.sp * st.stride];
        else w.top = st.lm[st.sm ? w.sp - SRT_STACK_SMEM : w.sp];
    }
    return e;
}
// Open a 4-wide node (srt_types.h SrtWide): 64 B as two LDG.256 (two 32-byte sectors of one line), four slab tests.
// key[k] = entry parameter of slot k's box, +inf when the slot is unused or the box is missed; sorted ascending on return
// together with the slots' refs (5 compare-exchanges), so that the caller can stack the farther children first.
// Conservative: one-sided slack on both ends of the slab interval, so rounding can only add candidates (the relative
// slack also covers origins far outside the grid, whose lattice snap is relative, not half a cell).  A NaN direction never
// gets here (closest_hit / k_trace_rays answer NaN rays up front); an infinite 1/d leaves +-inf or NaN slab parameters,
// which fminf / fmaxf order or drop.
__device__ __forceinline__ void wide_cx(float& ka, int& ra, float& kb, int& rb) {
    const bool sw = kb < ka;
    const float k0 = fminf(ka, kb), k1 = fmaxf(ka, kb);
    const int r0 = sw ? rb : ra, r1 = sw ? ra : rb;
    ka = k0; kb = k1; ra = r0; rb = r1;
}
__device__ __forceinline__ void wide_open(const SceneRef& sc, int node, const GridRay& g, float closest, float key[4], int ref[4]) {
    uint32_t bx[12];
    if (sc.nodes_global) {
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(bx[0]), "=r"(bx[1]), "=r"(bx[2]), "=r"(bx[3]), "=r"(bx[4]), "=r"(bx[5]), "=r"(bx[6]), "=r"(bx[7]) : "l"(sc.nodes + node));
        asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=r"(bx[8]), "=r"(bx[9]), "=r"(bx[10]), "=r"(bx[11]), "=r"(ref[0]), "=r"(ref[1]), "=r"(ref[2]), "=r"(ref[3])
            : "l"(reinterpret_cast<const char*>(sc.nodes + node) + 32));
    } else {
        const uint4* np = reinterpret_cast<const uint4*>(sc.nodes + node);
        const uint4 a = np[0], b = np[1], c = np[2], e = np[3];
        bx[0] = a.x; bx[1] = a.y; bx[2] = a.z; bx[3] = a.w; bx[4] = b.x; bx[5] = b.y; bx[6] = b.z; bx[7] = b.w;
        bx[8] = c.x; bx[9] = c.y; bx[10] = c.z; bx[11] = c.w;
        ref[0] = (int)e.x; ref[1] = (int)e.y; ref[2] = (int)e.z; ref[3] = (int)e.w;
    }
    const float ox = g.om.x, oy = g.om.y, oz = g.om.z, ix = g.inv.x, iy = g.inv.y, iz = g.inv.z;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float t0x = (grid_lo(bx[3 * k]) - ox) * ix, t1x = (grid_hi(bx[3 * k]) - ox) * ix;
        const float t0y = (grid_lo(bx[3 * k + 1]) - oy) * iy, t1y = (grid_hi(bx[3 * k + 1]) - oy) * iy;
        const float t0z = (grid_lo(bx[3 * k + 2]) - oz) * iz, t1z = (grid_hi(bx[3 * k + 2]) - oz) * iz;
        const float nn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
        const float ff = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
        const bool hit = ref[k] != SRT_WIDE_EMPTY && !(nn * 0.999999f > fminf(ff * 1.000001f, closest));
        key[k] = hit ? nn : INFINITY;
    }
    wide_cx(key[0], ref[0], key[1], ref[1]);
    wide_cx(key[2], ref[2], key[3], ref[3]);
    wide_cx(key[0], ref[0], key[2], ref[2]);
    wide_cx(key[1], ref[1], key[3], ref[3]);
    wide_cx(key[1], ref[1], key[2], ref[2]);
}
__device__ __forceinline__ void walk_test_leaf(const SceneRef& sc, V3 o, V3 d, Walk& w, float4 q0, float4 q1, float4 q2) {
    float t;
    if (tri_test_q(q0, q1, q2, o, d, w.closest, t)) {
        const uint32_t prio = __float_as_uint(q2.w);
        if (t < w.closest || w.best < 0 || prio > w.best_prio) { w.closest = t; w.best = w.pend; w.best_prio = prio; }
    }
}
__device__ __forceinline__ bool lbvh_step(const SceneRef& sc, V3 o, V3 d, const GridRay& g, Walk& w, const StackRef& stack, uint32_t lanes, uint32_t* visits) {
    if (w.node < 0 && w.top != SRT_STACK_EMPTY) {
        if (w.top >= 0) w.node = walk_pop(w, stack);
        else if (w.pend < 0) w.pend = ~walk_pop(w, stack);
    }
    if (w.node >= 0) {
        float key[4];
        int ref[4];
        wide_open(sc, w.node, g, w.closest, key, ref);
        if (visits) visits[0]++;
        // the farther children go on the stack first (the nearest of them pops first); the nearest child is the next node
#pragma unroll
        for (int k = 3; k >= 1; k--) {
            if (key[k] < INFINITY) {
                if (ref[k] < 0 && w.pend < 0) w.pend = ~ref[k];
                else walk_push(w, stack, ref[k]);
            }
        }
        int next = -1;
        if (key[0] < INFINITY) {
            if (ref[0] >= 0) next = ref[0];
            else if (w.pend < 0) w.pend = ~ref[0];
            else walk_push(w, stack, ref[0]);
        }
        w.node = next;
    }
    const bool has = w.pend >= 0;
    const bool stuck = has && w.node < 0 && (w.top == SRT_STACK_EMPTY || w.top < 0);  // only a triangle test gets this lane any further
    const uint32_t waiting = __ballot_sync(lanes, has);
    // (measured and dropped: letting up to 3 / 6 stuck lanes wait for company, and testing the leaves stacked right below the
    // pending one in the same batch: 1.37 -> 1.26-1.31 G incoherent rays/s)
    if (__popc(waiting) >= SRT_LEAF_BATCH || __any_sync(lanes, stuck)) {
        if (has) {
            if (visits) visits[1]++;
            const float4* tp = reinterpret_cast<const float4*>(sc.tris + w.pend);
            const float4 q0 = tp[0], q1 = tp[1], q2 = tp[2];
            walk_test_leaf(sc, o, d, w, q0, q1, q2);
            w.pend = -1;
        }
    }
    return (w.node >= 0) | (w.pend >= 0) | (w.top != SRT_STACK_EMPTY);
}

// The same walk for ONE lane on its own (the renderer's closest_hit: the lanes of a wavefront task each walk their own ray
// to the end, there is no warp to batch leaf tests with, and what counts is the latency of a step): the pending triangle's
// plane is loaded next to the node, so the two latencies overlap instead of adding up; the triangle is tested first,
// then the boxes against the updated `closest`.
__device__ __forceinline__ bool lbvh_step_single(const SceneRef& sc, V3 o, V3 d, const GridRay& g, Walk& w, const StackRef& stack, uint32_t* visits) {
    const bool has_node = w.node >= 0, has_leaf = w.pend >= 0;
    // unconditional load (triangle 0 stands in for "none": hot in L1), issued before the node's
    const float4* tp = reinterpret_cast<const float4*>(sc.tris + (has_leaf ? w.pend : 0));
    const float4 q0 = tp[0];
    float key[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    int ref[4] = {SRT_WIDE_EMPTY, SRT_WIDE_EMPTY, SRT_WIDE_EMPTY, SRT_WIDE_EMPTY};
    float closest_for_boxes = w.closest;
    if (has_leaf) {
        if (visits) visits[1]++;
        const float denom = plane_denom(q0, d);
        const float tq = plane_num(q0, o) / denom;
        if (fabsf(denom) >= 1e-8f && 0.0f <= tq && tq <= w.closest) {  // the plane part of tri::hit passes: now the other 32 B
            const float4 q1 = tp[1], q2 = tp[2];
            walk_test_leaf(sc, o, d, w, q0, q1, q2);
        }
        closest_for_boxes = w.closest;
    }
    w.pend = -1;
    int next = -1;
    if (has_node) {
        if (visits) visits[0]++;
        wide_open(sc, w.node, g, closest_for_boxes, key, ref);
#pragma unroll
        for (int k = 3; k >= 1; k--) {
            if (key[k] < INFINITY) {
                if (ref[k] < 0 && w.pend < 0) w.pend = ~ref[k];
                else walk_push(w, stack, ref[k]);
            }
        }
        if (key[0] < INFINITY) {
            if (ref[0] >= 0) next = ref[0];
            else if (w.pend < 0) w.pend = ~ref[0];
            else walk_push(w, stack, ref[0]);
        }
    }
    if (next < 0 && w.top != SRT_STACK_EMPTY) {
        if (w.top >= 0) next = walk_pop(w, stack);
        else if (w.pend < 0) w.pend = ~walk_pop(w, stack);
    }
    w.node = next;
    return (next >= 0) | (w.pend >= 0) | (w.top != SRT_STACK_EMPTY);
}

// closest hit over the LBVH (4-wide traversal nodes), nearest child first, the others stacked far to near.  Returns leaf-order triangle index or -1.
template <bool FLAT>
__device__ __forceinline__ int closest_hit(const SceneRef& sc, V3 o, V3 d, float& t_hit, uint32_t* visits = nullptr) {
    float closest = FLT_MAX;
    int best = -1;
    uint32_t best_prio = 0;
    if (sc.n_tris <= 0) return -1;
    // A NaN ray (buggy-Sellmeier refraction, Q1) can never hit: every tri::hit computes a NaN t.
    // Answer "miss" up front instead of walking the whole tree.
    if (!(d.x == d.x && d.y == d.y && d.z == d.z && o.x == o.x && o.y == o.y && o.z == o.z)) return -1;
    if (FLAT) return closest_hit_flat(sc, o, d, t_hit);
    if (sc.n_tris == 1) {
        float t;
        if (tri_test(sc.tris, o, d, closest, t)) { t_hit = t; return 0; }
        return -1;
    }
    const GridRay g = grid_ray(sc, o, d);
    int local_stack[SRT_STACK_MAX];
    StackRef stack;
    stack.sm = nullptr; stack.lm = local_stack; stack.stride = 0;
    Walk w;
    walk_reset(w, true);
    while (lbvh_step_single(sc, o, d, g, w, stack, visits)) {}
    t_hit = w.closest;
    return w.best;
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
// sample = index of this sample inside its pixel; only the opt-in stratified sampler looks at it
// (pixel_stratified_sample_square rendering.cu:58-64: sub-cell (sample % n, sample / n) of an n x n grid)
__device__ __forceinline__ void camera_ray(const WaveParams& P, uint32_t i, uint32_t j, uint32_t sample, Rng& rng, Path& p) {
    const SrtCamera& c = P.cam;
    const V3 du = mk(c.du[0], c.du[1], c.du[2]), dv = mk(c.dv[0], c.dv[1], c.dv[2]);
    const V3 center = mk(c.center[0], c.center[1], c.center[2]);
    const V3 pixel_center = (mk(c.p00[0], c.p00[1], c.p00[2]) + ((float)i * du)) + ((float)j * dv);
    float px, py;
    if (P.strat_n) {
        const uint32_t sy = sample / P.strat_n, sx = sample - sy * P.strat_n;
        px = -0.5f + P.strat_recip * ((float)sx + rng_uniform(rng));
        py = -0.5f + P.strat_recip * ((float)sy + rng_uniform(rng));
    } else {
        px = -0.5f + rng_uniform(rng);
        py = -0.5f + rng_uniform(rng);
    }
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

// One out-of-line copy for its two callers (scatter, and the miss / emitter end of a sample): the hot loop of k_wavefront
// sits at the edge of the 32 KB instruction cache, where every inlined copy of the seven interpolations costs more in
// fetch stalls than a call does.  Scalars in, scalars out, so the path state stays in registers.
struct Pw7 { float w0, w1, w2, w3, w4, w5, w6; };
__device__ __noinline__ Pw7 mul_spectrum_scalars(const float* __restrict__ spec, float hero, uint32_t valid, float w0, float w1, float w2, float w3, float w4,
                                                float w5, float w6) {
    float wl[SRT_N_WL];
    hero_rotations(hero, wl);
    float pw[SRT_N_WL] = {w0, w1, w2, w3, w4, w5, w6};
#pragma unroll
    for (int k = 0; k < SRT_N_WL; k++)
        if ((uint32_t)k < valid) pw[k] *= interp95(spec, wl[k]);
    Pw7 r;
    r.w0 = pw[0]; r.w1 = pw[1]; r.w2 = pw[2]; r.w3 = pw[3]; r.w4 = pw[4]; r.w5 = pw[5]; r.w6 = pw[6];
    return r;
}
__device__ __forceinline__ void mul_spectrum(Path& p, const float* __restrict__ spec) {  // ray/ray.cuh:60-69
    const Pw7 r = mul_spectrum_scalars(spec, p.hero, p.valid, p.pw[0], p.pw[1], p.pw[2], p.pw[3], p.pw[4], p.pw[5], p.pw[6]);
    p.pw[0] = r.w0; p.pw[1] = r.w1; p.pw[2] = r.w2; p.pw[3] = r.w3; p.pw[4] = r.w4; p.pw[5] = r.w5; p.pw[6] = r.w6;
}

// dev_spectrum_to_XYZ (color/color.cu:88-104) added into the film accumulator of one pixel.
// Rare (about 1 % of the rays reach an emitter), so it is an out-of-line call; all arguments are
// scalars so that the caller's path state stays in registers.
__device__ __noinline__ void film_add_scalars(const float* __restrict__ cie, float* __restrict__ acc, size_t plane, size_t pix, float hero,
                                              uint32_t valid, float w0, float w1, float w2, float w3, float w4, float w5, float w6) {
    float wl[SRT_N_WL];
    hero_rotations(hero, wl);
    const float pw[SRT_N_WL] = {w0, w1, w2, w3, w4, w5, w6};
    const float delta = (830.0f - 360.0f) / 7.0f;
    float x = 0.0f, y = 0.0f, z = 0.0f;
#pragma unroll
    for (int k = 0; k < SRT_N_WL; k++)
        if ((uint32_t)k < valid) {
            x += interp95(cie, wl[k]) * pw[k] * delta;
            y += interp95(cie + SRT_NS, wl[k]) * pw[k] * delta;
            z += interp95(cie + 2 * SRT_NS, wl[k]) * pw[k] * delta;
        }
    acc[pix] += x;
    acc[plane + pix] += y;
    acc[2 * plane + pix] += z;
}
__device__ __forceinline__ void film_add(const SceneRef& sc, const Path& p, float* __restrict__ acc, size_t plane, size_t pix) {
    if (p.valid == 0) return;  // contributes (0,0,0): x + 0 leaves the sum unchanged
    film_add_scalars(sc.cie, acc, plane, pix, p.hero, p.valid, p.pw[0], p.pw[1], p.pw[2], p.pw[3], p.pw[4], p.pw[5], p.pw[6]);
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
// Returns false when the path ends here (metal absorbed the ray).  `mtype` is warp-uniform in the
// wavefront (queues are sorted by material), so the branches below do not diverge there.
__device__ __forceinline__ bool scatter(const SceneRef& sc, const SrtTri* __restrict__ tri, uint32_t mtype, Path& p, Rng& rng) {
    SRT_FINE_BEGIN;
    const float4 q0 = *reinterpret_cast<const float4*>(tri);
    const uint32_t bits = __float_as_uint((reinterpret_cast<const float4*>(tri) + 2)->z);
    const SrtMaterial* m = sc.mats + SRT_TRI_MAT(bits);
    V3 n = mk(q0.x, q0.y, q0.z);
    const bool front = dot(p.d, n) < 0;  // hit_record::set_face_normal, primitives/hit_record.cuh:30-43
    if (!front) n = -n;
    const V3 uin = unit(p.d);
    SRT_FINE(6);
    V3 out;
    float eps_sign = 1.0f;
    bool alive = true;
    if (mtype == SRT_DIELECTRIC) {  // refraction_scatter :103-135, evaluated at the hero wavelength only
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
    } else {
        const V3 ruv = random_unit_vector(rng);  // both remaining materials draw one unit vector first
        if (mtype == SRT_METALLIC) {  // reflection_scatter :22-37
            out = reflect(uin, n) + (m->fuzz * ruv);
            alive = dot(out, n) > 0;
            if (!alive) p.valid = 0;
        } else {  // lambertian_scatter :9-19
            out = n + ruv;
            const float s = 1e-8f;
            if ((fabsf(out.x) < s) && (fabsf(out.y) < s) && (fabsf(out.z) < s)) out = n;
        }
    }
    SRT_FINE(7);
    mul_spectrum(p, m->spec);
    SRT_FINE(8);
    p.o = p.o + ((eps_sign * SRT_EPSILON) * n);
    p.d = out;
    return alive;
}

// What happened to the ray that was just traced
enum : int { EV_DONE = -1 };  // sample finished; otherwise the value is the queue index 1..3 (lambert, metal, dielectric)

// trace p's ray and either finish the sample or leave p at the hit (p.o = hit point).
// Returns EV_DONE or the material queue (1 lambertian, 2 metallic, 3 dielectric); tri_out = leaf-order index.
// ray vs the scene's bounding box, conservative (approximate reciprocals, 1e-4 relative slack)
__device__ __forceinline__ bool misses_scene_box(const WaveParams& P, V3 o, V3 d) {
    const float ix = rcp_approx(d.x), iy = rcp_approx(d.y), iz = rcp_approx(d.z);
    const float ax = (P.scene_lo[0] - o.x) * ix, bx = (P.scene_hi[0] - o.x) * ix;
    const float ay = (P.scene_lo[1] - o.y) * iy, by = (P.scene_hi[1] - o.y) * iy;
    const float az = (P.scene_lo[2] - o.z) * iz, bz = (P.scene_hi[2] - o.z) * iz;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    return tn > tf * 1.0002f + 1e-3f;  // NaN anywhere -> false -> not a certain miss
}

template <bool FLAT>
__device__ __forceinline__ int extend(const SceneRef& sc, const WaveParams& P, Path& p, int& tri_out, float* acc, size_t pix, bool primary = false) {
    float t = 0.f;
    int tri = -1;
    SRT_FINE_BEGIN;
    // camera rays of a whole warp often all pass beside the scene (46 % of the 16:9 frame lies outside
    // the box): one cheap slab test spares the warp the whole wide-leaf loop.
    bool skip = false;
    if (FLAT && primary) skip = __all_sync(__activemask(), misses_scene_box(P, p.o, p.d));
    SRT_FINE(2);
    if (!skip) tri = closest_hit<FLAT>(sc, p.o, p.d, t);
    SRT_FINE(3);
    uint32_t bits = 0, mtype = SRT_LAMBERTIAN;
    if (tri >= 0) {
        bits = __float_as_uint((reinterpret_cast<const float4*>(sc.tris + tri) + 2)->z);
        mtype = SRT_TRI_MTYPE(bits);
    }
    if (tri < 0 || mtype == SRT_EMISSIVE) {
        // the sample ends here: a miss multiplies by the background spectrum (ray_bounce, rendering.cu:24-27),
        // an emitter by its emission spectrum (scatter() returns false, material.cu:83-86,95)
        if (tri >= 0 || !P.bg_is_zero) {
            mul_spectrum(p, tri < 0 ? sc.bg : sc.mats[SRT_TRI_MAT(bits)].spec);
            film_add(sc, p, acc, P.plane, pix);
        }
        SRT_FINE(4);
        return EV_DONE;
    }
    SRT_FINE(5);
    p.o = mk(p.o.x + t * p.d.x, p.o.y + t * p.d.y, p.o.z + t * p.d.z);  // ray::at, ray/ray.cuh:44-47
    tri_out = tri;
    return mtype == SRT_METALLIC ? 2 : (mtype == SRT_DIELECTRIC ? 3 : 1);
}

// ------------------------------------------------------------------------------ shared-memory scene
// Small scenes (all three reference scenes) are staged once per block: nodes | tris | mats | cie | bg
template <bool SMEM, bool FLAT>
__device__ __forceinline__ SceneRef load_scene(const WaveParams& P, unsigned char* smem) {
    SceneRef sc;
    sc.n_tris = P.n_tris;
    if (!SMEM) {
        sc.nodes = P.nodes; sc.grid = P.grid; sc.nodes_global = true; sc.tris = P.tris; sc.units = nullptr; sc.n_units = 0; sc.flat_guard = 0.f; sc.flat_tol = 0.f; sc.mats = P.mats; sc.cie = P.cie; sc.bg = P.bg;
        return sc;
    }
    // layout: [nodes | pre-test records (flat scenes)] | tris | mats | cie | bg
    const int n_nodes = P.n_tris > 1 ? P.n_tris - 1 : 0;
    float4* dst = reinterpret_cast<float4*>(smem);
    const int v_head = FLAT ? P.n_units * (int)(sizeof(SrtFlatUnit) / 16) : n_nodes * (int)(sizeof(SrtWide) / 16);
    const float4* head = FLAT ? reinterpret_cast<const float4*>(P.flat_units) : reinterpret_cast<const float4*>(P.nodes);
    const int n_stage_tris = FLAT ? 2 * P.n_units : P.n_tris;  // flat order: two triangle slots per unit
    const float4* tri_src = FLAT ? reinterpret_cast<const float4*>(P.flat_tris) : reinterpret_cast<const float4*>(P.tris);
    const int v_tris = n_stage_tris * (int)(sizeof(SrtTri) / 16), v_mats = P.n_mats * (int)(sizeof(SrtMaterial) / 16);
    for (int i = threadIdx.x; i < v_head; i += blockDim.x) dst[i] = head[i];
    for (int i = threadIdx.x; i < v_tris; i += blockDim.x) dst[v_head + i] = tri_src[i];
    for (int i = threadIdx.x; i < v_mats; i += blockDim.x) dst[v_head + v_tris + i] = reinterpret_cast<const float4*>(P.mats)[i];
    float* f = reinterpret_cast<float*>(dst + v_head + v_tris + v_mats);
    for (int i = threadIdx.x; i < 3 * SRT_NS; i += blockDim.x) f[i] = P.cie[i];
    for (int i = threadIdx.x; i < SRT_NS; i += blockDim.x) f[3 * SRT_NS + i] = P.bg[i];
    __syncthreads();
    sc.nodes = FLAT ? nullptr : reinterpret_cast<const SrtWide*>(dst);
    sc.grid = P.grid; sc.nodes_global = false;
    sc.units = FLAT ? reinterpret_cast<const SrtFlatUnit*>(dst) : nullptr;
    sc.n_units = FLAT ? P.n_units : 0;
    sc.flat_guard = P.flat_guard; sc.flat_tol = P.flat_tol;
    sc.tris = reinterpret_cast<const SrtTri*>(dst + v_head);
    sc.mats = reinterpret_cast<const SrtMaterial*>(dst + v_head + v_tris);
    sc.cie = f;
    sc.bg = f + 3 * SRT_NS;
    return sc;
}

// ------------------------------------------------------------------------------ slot <-> pixel
// Pixel slots are grouped by image tile: slot = k * (tile_w*tile_h) + pixel-in-tile, where k indexes
// the list of chunk-local tiles THIS rank owns.  Consecutive slots are one row of a tile, so a warp
// that fetches 32 fresh slots renders a compact patch, and a rank only numbers its own pixels.
__device__ __forceinline__ void slot_pixel(const WaveParams& P, uint32_t slot, uint32_t& ci, uint32_t& cj) {
    // tile_w and tile_w*tile_h are powers of two (checked on the host): shifts, not divisions
    const uint32_t k = slot >> P.tile_slots_log2, l = slot & ((1u << P.tile_slots_log2) - 1u);
    const uint32_t tile = P.tiles[k];
    const uint32_t ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
    const uint32_t ly = l >> P.tile_w_log2, lx = l & (P.tile_w - 1u);
    ci = tx * P.tile_w + lx;
    cj = ty * P.tile_h + ly;
}
// the reference's per-thread seed index (rendering.cu:125-137, 28x16 blocks, grid from the nominal chunk size)
__device__ __forceinline__ uint32_t ref_thread_index(const WaveParams& P, uint32_t ci, uint32_t cj) {
    return (cj % 16u) * 28u + (ci % 28u) + 448u * ((cj / 16u) * P.ref_grid_x + ci / 28u);
}

// warp-aggregated push into the block's four shared-memory queues of local slot ids, all four at once: lanes that go to
// the same queue find each other with one __match_any_sync, the first lane of every group reserves the group's room with
// one shared-memory atomicAdd (up to four addresses in one instruction), lanes write at base + rank-in-group.
// which = 0..3 (regenerate, lambertian, metallic, dielectric), or 4 = this lane pushes nothing.
__device__ __forceinline__ void queue_push_all(uint16_t* __restrict__ q, uint32_t S, int* counters, uint32_t which, uint32_t local_slot) {
    const uint32_t peers = __match_any_sync(0xffffffffu, which);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader && which < 4u) base = atomicAdd(counters + which, __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (which < 4u) q[which * S + base + __popc(peers & ((1u << lane) - 1))] = (uint16_t)local_slot;
}

__device__ __forceinline__ void store_hit_state(const WaveParams& P, uint32_t slot, const Path& p, int tri) {
    P.R0[slot] = make_float4(p.o.x, p.o.y, p.o.z, __int_as_float(tri));
    P.R1[slot] = make_float4(p.d.x, p.d.y, p.d.z, __uint_as_float(p.valid | (p.bounce << 3)));
    P.P0[slot] = make_float4(p.pw[0], p.pw[1], p.pw[2], p.pw[3]);
    P.P1[slot] = make_float4(p.pw[4], p.pw[5], p.pw[6], p.hero);
}
__device__ __forceinline__ void load_hit_state(const WaveParams& P, uint32_t slot, Path& p, int& tri) {
    const float4 r0 = P.R0[slot], r1 = P.R1[slot], p0 = P.P0[slot], p1 = P.P1[slot];
    p.o = mk(r0.x, r0.y, r0.z);
    p.d = mk(r1.x, r1.y, r1.z);
    const uint32_t meta = __float_as_uint(r1.w);
    p.valid = meta & 7u;
    p.bounce = meta >> 3;
    p.pw[0] = p0.x; p.pw[1] = p0.y; p.pw[2] = p0.z; p.pw[3] = p0.w;
    p.pw[4] = p1.x; p.pw[5] = p1.y; p.pw[6] = p1.z;
    p.hero = p1.w;
    tri = __float_as_int(r0.w);
}
__device__ __forceinline__ void store_rng(uint4* __restrict__ g0, uint2* __restrict__ g1, uint32_t slot, const Rng& r) {
    g0[slot] = make_uint4(r.d, r.v0, r.v1, r.v2);
    g1[slot] = make_uint2(r.v3, r.v4);
}
__device__ __forceinline__ Rng load_rng(const uint4* __restrict__ g0, const uint2* __restrict__ g1, uint32_t slot) {
    const uint4 a = g0[slot];
    const uint2 b = g1[slot];
    Rng r;
    r.d = a.x; r.v0 = a.y; r.v1 = a.z; r.v2 = a.w; r.v3 = b.x; r.v4 = b.y;
    return r;
}
__device__ __forceinline__ bool slot_owned(const WaveParams& P, uint32_t slot, uint32_t& ci, uint32_t& cj) {
    if (slot >= P.nslots) return false;
    slot_pixel(P, slot, ci, cj);
    return ci < P.cw && cj < P.ch;  // edge tiles stick out of the (clipped) chunk
}

// ------------------------------------------------------------------------------ kernels
__global__ void k_init_slots(WaveParams P) {  // init_random_states (rendering.cu:120-138), one state per reference thread slot
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.nslots) return;
    uint32_t ci, cj;
    slot_pixel(P, slot, ci, cj);
    store_rng(P.G0, P.G1, slot, rng_seed(1984u + ref_thread_index(P, ci, cj)));
}

// First guess of a pixel's cost before any of it is rendered: the un-jittered ray through the pixel centre misses
// everything or ends on an emitter (1 pass per sample), or starts a bounce chain (counted as 6).  Only used to hand
// out the expensive pixels first in the first round of a chunk; later rounds use the passes really spent.
__global__ void __launch_bounds__(256) k_prior_cost(WaveParams P) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.nslots) return;
    uint32_t ci, cj, c = 0;
    if (slot_owned(P, slot, ci, cj)) {
        const SrtCamera& cam = P.cam;
        const V3 du = mk(cam.du[0], cam.du[1], cam.du[2]), dv = mk(cam.dv[0], cam.dv[1], cam.dv[2]);
        const V3 o = mk(cam.center[0], cam.center[1], cam.center[2]);
        const V3 d = ((mk(cam.p00[0], cam.p00[1], cam.p00[2]) + ((float)(P.off_x + ci) * du)) + ((float)(P.off_y + cj) * dv)) - o;
        SceneRef sc;
        sc.nodes = P.nodes; sc.grid = P.grid; sc.nodes_global = true; sc.tris = P.tris; sc.units = nullptr; sc.n_units = 0; sc.flat_guard = 0.f; sc.flat_tol = 0.f; sc.mats = P.mats; sc.cie = nullptr; sc.bg = nullptr; sc.n_tris = P.n_tris;
        float t;
        const int tri = closest_hit<false>(sc, o, d, t);
        c = 1;
        if (tri >= 0 && SRT_TRI_MTYPE(__float_as_uint((reinterpret_cast<const float4*>(sc.tris + tri) + 2)->z)) != SRT_EMISSIVE) c = 6;
    }
    P.cost[slot] = c;
}

// Persistent-block wavefront with pixel streaming.  A block keeps P.block_slots paths in flight and
// runs its own bounce loop: four queues of local slot ids in shared memory (regenerate | lambertian |
// metallic | dielectric), double buffered.  Every pass cuts the four queues into 32-item tasks -- the
// full warps of each queue (one kind of work per warp), then the four remainders pooled into mixed
// warps -- and warps pull tasks from a shared counter (cheap items do not leave a warp idle); results
// are pushed into the other buffer with warp-aggregated shared-memory atomics.  A local slot renders one pixel at a time, all
// its samples in order (the pixel's XORWOW stream is serial); when the pixel is finished the slot
// fetches the next unrendered pixel slot from a global counter (one atomic per warp), so every block
// stays full until the whole chunk runs out of pixels -- no per-tile tail, no wave quantisation, and
// the state of the paths in flight (grid x block_slots records) stays L2 resident.  One launch per
// chunk, no host round trips.  The ray trace (extend) has ONE call site so the hot loop stays inside
// the instruction cache.
#define SRT_NO_SLOT 0xFFFFFFFFu
// probe builds only (make EXTRA=-DSRT_PHASE_CLOCKS, tools/latency_probe.py): warp 0 of blocks 0..3 records the SM cycles its first
// task of every pass spends in {task fetch + state addresses, regenerate / scatter, closest hit, state store + queue push}
// into the pass-log rows of blocks 4..7
#ifdef SRT_PHASE_CLOCKS
#define SRT_CLK(v) do { asm volatile("mov.u64 %0, %%clock64;" : "=l"(v)); } while (0)
#else
#define SRT_CLK(v) do { } while (0)
#endif
// MINB = resident blocks per SM the register allocation allows: 4 (64 registers, 32 warps per SM: the most work in flight, best when a
// rank has more pixels than paths in flight) or 3 (80 registers, 24 warps: shorter dependent-issue chains per task, best when a rank's
// render is bound by the serial chain of its longest pixels -- the per-rank share of an 8-GPU split)
template <bool SMEM, bool FLAT, int MINB>
__global__ void __launch_bounds__(SRT_WAVE_BLOCK, MINB) k_wavefront(WaveParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t S = P.block_slots;
    uint16_t* qbuf = reinterpret_cast<uint16_t*>(smem);                        // [2][4][S] queues of local slot ids
    uint32_t* pslot = reinterpret_cast<uint32_t*>(smem + 16u * S);             // [S] pixel slot a local slot renders
    uint16_t* started = reinterpret_cast<uint16_t*>(smem + 20u * S);           // [S] samples started of that pixel
    uint16_t* passes = reinterpret_cast<uint16_t*>(smem + 22u * S);            // [S] passes spent on that pixel in this launch
    uint32_t* pxy = reinterpret_cast<uint32_t*>(smem + 24u * S);               // [S] that pixel's chunk coordinates, y << 16 | x
    // ONE block barrier per pass: queue lengths are triple buffered (pass k reads cnt[k % 3], counts its output in cnt[(k + 1) % 3] and
    // clears cnt[(k + 2) % 3], which nobody has touched since the barrier before), the task ticket is double buffered likewise
    __shared__ int cnt[3][4];
    __shared__ int next_task[2];
    const SceneRef sc = load_scene<SMEM, FLAT>(P, smem + P.queue_bytes);
    const uint32_t first = blockIdx.x * S;  // this block's records in the in-flight state arrays
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 12) (&cnt[0][0])[threadIdx.x] = 0;
    if (threadIdx.x < 2) next_task[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) cnt[0][0] = (int)S;
    // pass 0 input: every local slot asks for a pixel
    for (uint32_t l = threadIdx.x; l < S; l += blockDim.x) {
        pslot[l] = SRT_NO_SLOT;
        started[l] = 0;
        passes[l] = 0;
        qbuf[l] = (uint16_t)l;
    }
    __syncthreads();
    unsigned long long rays = 0;
    int cur = 0, cset = 0;  // queue buffer (pass % 2) and counter set (pass % 3) this pass reads
    uint32_t npass = 0;
    while (true) {
        const int nR = cnt[cset][0], nL = cnt[cset][1], nM = cnt[cset][2], nD = cnt[cset][3];
        if ((nR | nL | nM | nD) == 0) break;
        const int co_i = cset == 2 ? 0 : cset + 1, cz_i = co_i == 2 ? 0 : co_i + 1;
        if (threadIdx.x < 4) cnt[cz_i][threadIdx.x] = 0;  // the output counters of the NEXT pass
        if (threadIdx.x == 0) next_task[cur ^ 1] = 0;
#ifdef SRT_PHASE_CLOCKS
        if (P.pass_log && blockIdx.x < 4 && threadIdx.x == 0 && npass < SRT_PASS_LOG_PASSES) {
#else
        if (P.pass_log && blockIdx.x < SRT_PASS_LOG_BLOCKS && threadIdx.x == 0 && npass < SRT_PASS_LOG_PASSES) {
#endif
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            P.pass_log[blockIdx.x * SRT_PASS_LOG_PASSES + npass] = make_uint4((uint32_t)now, (uint32_t)nR, (uint32_t)nL, (uint32_t)nM | ((uint32_t)nD << 16));
        }
        // tasks of this pass (32 items each): the full warps of every queue first -- one kind of work per warp --, then the
        // queues' remainders (< 32 items each) pooled into mixed warps, so that S items never need more than S / 32 tasks
        const int fR = nR >> 5, fL = nL >> 5, fM = nM >> 5, fD = nD >> 5;
        const int rR = nR & 31, rL = nL & 31, rM = nM & 31, rD = nD & 31;
        const int tL = fR + fL, tM = tL + fM, tF = tM + fD, rem = rR + rL + rM + rD;
        const bool pooled = !(P.sched_flags & 4u);
        const int n_tasks = tF + (pooled ? ((rem + 31) >> 5) : ((rR > 0) + (rL > 0) + (rM > 0) + (rD > 0)));
        const uint16_t* qi = qbuf + (size_t)cur * 4 * S;
        uint16_t* qo = qbuf + (size_t)(cur ^ 1) * 4 * S;
        int* co = cnt[co_i];
        int* ticket = &next_task[cur];
#ifdef SRT_PHASE_CLOCKS
        bool first_task = true;
#endif
        while (true) {
            unsigned long long ck0 = 0, ck1 = 0, ck2 = 0, ck3 = 0, ck4 = 0;
            SRT_CLK(ck0);
            int task = 0;
            if (lane == 0) task = atomicAdd(ticket, 1);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= n_tasks) break;
            int kind, k;
            bool have = true;
            if (task < tF) {  // warp-uniform kind
                if (task < fR) { kind = 0; k = task * 32 + lane; }
                else if (task < tL) { kind = 1; k = (task - fR) * 32 + lane; }
                else if (task < tM) { kind = 2; k = (task - tL) * 32 + lane; }
                else { kind = 3; k = (task - tM) * 32 + lane; }
            } else if (!pooled) {  // one partial warp per non-empty remainder
                int j = task - tF;
                if (rR > 0 && j-- == 0) { kind = 0; k = fR * 32 + lane; have = lane < rR; }
                else if (rL > 0 && j-- == 0) { kind = 1; k = fL * 32 + lane; have = lane < rL; }
                else if (rM > 0 && j-- == 0) { kind = 2; k = fM * 32 + lane; have = lane < rM; }
                else { kind = 3; k = fD * 32 + lane; have = lane < rD; }
            } else {
                const int v = (task - tF) * 32 + lane;
                have = v < rem;
                if (v < rR) { kind = 0; k = fR * 32 + v; }
                else if (v < rR + rL) { kind = 1; k = fL * 32 + (v - rR); }
                else if (v < rR + rL + rM) { kind = 2; k = fM * 32 + (v - rR - rL); }
                else { kind = 3; k = fD * 32 + (v - rR - rL - rM); }
            }
            const uint32_t l = have ? qi[kind * S + k] : 0u;
            const uint32_t rec = first + l;
            uint32_t slot = have ? pslot[l] : SRT_NO_SLOT;
            uint32_t s = started[l];
            uint32_t np = passes[l] + 1u;
            uint32_t ci = 0, cj = 0;
            size_t pix = 0;
            Path p;
            Rng rng;
            int tri = -1, ev = EV_DONE;
            bool trace = false, retired = false;
            // local slots without a pixel take the next entries of the hand-out order: one global atomic per warp
            const bool fetch = have && kind == 0 && slot == SRT_NO_SLOT;
            const uint32_t fm = __ballot_sync(0xffffffffu, fetch);
            if (fm) {
                uint32_t fbase = 0;
                if (lane == __ffs(fm) - 1) fbase = atomicAdd(P.next_slot, (uint32_t)__popc(fm));
                fbase = __shfl_sync(0xffffffffu, fbase, __ffs(fm) - 1);
                const uint32_t idx = fbase + __popc(fm & ((1u << lane) - 1));
                if (fetch) {
                    if (idx >= P.n_order) retired = true;  // the chunk has no pixels left: this local slot is done
                    else {
                        const uint32_t cand = P.order ? P.order[idx] : idx;
                        if (slot_owned(P, cand, ci, cj)) {  // edge tiles stick out of the chunk: such a slot asks again next pass
                            slot = cand;
                            pslot[l] = slot;
                            pxy[l] = (cj << 16) | ci;  // tile look-up and divisions once per pixel, not once per pass
                            s = P.s_begin;
                            np = 1;
                            rng = load_rng(P.G0, P.G1, slot);
                        }
                    }
                }
                if (P.drain_clock && __any_sync(0xffffffffu, retired) && lane == 0) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    atomicMin(P.drain_clock, now);
                }
            }
            SRT_CLK(ck1);
            if (have && slot != SRT_NO_SLOT) {
                if (kind == 0) {
                    if (!fetch) {
                        const uint32_t xy = pxy[l];
                        ci = xy & 0xFFFFu; cj = xy >> 16;
                        rng = load_rng(P.L0, P.L1, rec);
                    }
                    pix = (size_t)(P.off_y + cj) * P.img_w + (P.off_x + ci);
                    if (s < P.s_end) {  // next sample of this pixel (rendering.cu:215-228)
                        camera_ray(P, P.off_x + ci, P.off_y + cj, s, rng, p);
                        s++;
                        started[l] = (uint16_t)s;
                        trace = P.bounce_limit != 0;  // limit 0: the bounce loop never runs, valid = 0
                    }
                } else {  // scatter at the stored hit
                    const uint32_t xy = pxy[l];
                    ci = xy & 0xFFFFu; cj = xy >> 16;
                    pix = (size_t)(P.off_y + cj) * P.img_w + (P.off_x + ci);
                    rng = load_rng(P.L0, P.L1, rec);
                    load_hit_state(P, rec, p, tri);
                    const uint32_t mtype = kind == 2 ? SRT_METALLIC : (kind == 3 ? SRT_DIELECTRIC : SRT_LAMBERTIAN);
                    const bool alive = scatter(sc, sc.tris + tri, mtype, p, rng);
                    p.bounce++;
                    trace = alive && p.bounce < P.bounce_limit;  // absorbed, or bounce limit: valid = 0 (rendering.cu:38)
                }
            }
            SRT_CLK(ck2);
            if (trace) {
                rays++;
                ev = extend<FLAT>(sc, P, p, tri, P.acc, pix, kind == 0);
            }
            SRT_CLK(ck3);
            if (have && slot != SRT_NO_SLOT) {
                if (ev == EV_DONE && s >= P.s_end) {  // pixel finished: its RNG state goes back to the pixel (carried into the next round / chunk)
                    store_rng(P.G0, P.G1, slot, rng);
                    if (P.cost) P.cost[slot] += np;
                    pslot[l] = SRT_NO_SLOT;
                } else {
                    store_rng(P.L0, P.L1, rec, rng);
                    passes[l] = (uint16_t)np;
                    if (ev != EV_DONE) store_hit_state(P, rec, p, tri);
                }
            }
            // next sample or next pixel (queue 0), or the queue of the material that was hit; retired slots and empty lanes push nothing
            queue_push_all(qo, S, co, ev != EV_DONE ? (uint32_t)ev : ((have && !retired) ? 0u : 4u), l);
            SRT_CLK(ck4);
#ifdef SRT_PHASE_CLOCKS
            if (first_task && P.pass_log && blockIdx.x < 4 && threadIdx.x == 0 && npass < SRT_PASS_LOG_PASSES)
                P.pass_log[(blockIdx.x + 4) * SRT_PASS_LOG_PASSES + npass] = make_uint4((uint32_t)(ck1 - ck0), (uint32_t)(ck2 - ck1), (uint32_t)(ck3 - ck2), (uint32_t)(ck4 - ck3) | ((uint32_t)kind << 28));
            first_task = false;
#endif
            (void)ck0; (void)ck1; (void)ck2; (void)ck3; (void)ck4;
        }
        __syncthreads();
        cur ^= 1;
        cset = co_i;
        npass++;
    }
#ifdef SRT_PHASE_CLOCKS
    if (P.pass_log && blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < 16; k++) {
            P.pass_log[7 * SRT_PASS_LOG_PASSES + k] = make_uint4((uint32_t)g_fine[2 * k], (uint32_t)(g_fine[2 * k] >> 32), (uint32_t)g_fine[2 * k + 1], 0u);
            g_fine[2 * k] = 0; g_fine[2 * k + 1] = 0;
        }
#endif
    if (P.ray_counter && rays) atomicAdd(P.ray_counter, rays);
    if (P.drain_clock && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(P.drain_clock + 1, now);
    }
}

// per-pixel persistent kernel: same device functions, no queues (cross-check / comparison)
template <bool SMEM, bool FLAT>
__global__ void __launch_bounds__(SRT_BLOCK) k_megakernel(WaveParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneRef sc = load_scene<SMEM, FLAT>(P, smem);
    unsigned long long rays = 0;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < P.nslots; slot += gridDim.x * blockDim.x) {
        uint32_t ci, cj;
        if (!slot_owned(P, slot, ci, cj)) continue;
        const uint32_t x = P.off_x + ci, y = P.off_y + cj;
        const size_t pix = (size_t)y * P.img_w + x;
        Rng rng = load_rng(P.G0, P.G1, slot);
        for (uint32_t s = P.s_begin; s < P.s_end; s++) {
            Path p;
            camera_ray(P, x, y, s, rng, p);
            if (P.bounce_limit == 0) continue;
            int tri = -1;
            rays++;
            int ev = extend<FLAT>(sc, P, p, tri, P.acc, pix);
            while (ev != EV_DONE) {
                const bool alive = scatter(sc, sc.tris + tri, ev == 2 ? SRT_METALLIC : (ev == 3 ? SRT_DIELECTRIC : SRT_LAMBERTIAN), p, rng);
                p.bounce++;
                if (!alive || p.bounce >= P.bounce_limit) break;
                rays++;
                ev = extend<FLAT>(sc, P, p, tri, P.acc, pix);
            }
        }
        store_rng(P.G0, P.G1, slot, rng);
    }
    if (P.ray_counter && rays) atomicAdd(P.ray_counter, rays);
}

// film: XYZ sum / spp -> sRGB 0..255 (save_to_fb rendering.cu:140-149, color.cu:15-49)
__device__ __forceinline__ void tonemap_pixel(const float* __restrict__ acc, size_t plane, size_t pix, uint32_t spp, unsigned char* __restrict__ out_rgb,
                                              size_t out_plane, size_t out_i) {
    const float inv = 1 / (float)spp;  // pixel_color / float(spp) = (1/spp) * v
    const float X = inv * acc[pix], Y = inv * acc[plane + pix], Z = inv * acc[2 * plane + pix];
    const float m[9] = {3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f, 0.0415560f, 0.0556434f, -0.2040259f, 1.0572252f};
    const float lin[3] = {(m[0] * X) + (m[1] * Y) + (m[2] * Z), (m[3] * X) + (m[4] * Y) + (m[5] * Z), (m[6] * X) + (m[7] * Y) + (m[8] * Z)};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float v = lin[c];
        const float g = v < 0.0f ? 0.0f : (v < 0.0031308f ? 12.92f * v : (v < 1.0f ? ((1.055f * powf(v, 0.416666f)) - 0.055f) : 1.0f));
        out_rgb[c * out_plane + out_i] = (unsigned char)(int)(g * 255.99f);  // 0..255: one byte per channel crosses PCIe, the host widens it
    }
}
// a rectangle of the film, raster order (one chunk, or the whole image)
__global__ void k_resolve(const float* __restrict__ acc, size_t plane, uint32_t img_w, uint32_t off_x, uint32_t off_y, uint32_t w, uint32_t h,
                          uint32_t spp, unsigned char* __restrict__ out_rgb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const uint32_t cy = i / w, cx = i - cy * w;
    tonemap_pixel(acc, plane, (size_t)(off_y + cy) * img_w + (off_x + cx), spp, out_rgb, (size_t)w * h, i);
}
// pixels [first, first + count) of the raster: the slice a rank owns after the film reduce-scatter (renderer.cu)
__global__ void k_resolve_slice(const float* __restrict__ acc, size_t plane, size_t first, uint32_t count, uint32_t out_plane, uint32_t spp,
                                unsigned char* __restrict__ out_rgb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    tonemap_pixel(acc, plane, first + i, spp, out_rgb, out_plane, i);
}

// Standalone closest-hit queries (BASELINE.json configs[3]); the scene stays in global memory.
// Persistent warps with ray refill: rays of one warp end after very different numbers of node visits (ncu on 1M incoherent
// rays: 7.8 of 32 lanes active with one ray per thread), so lanes that are done fetch the next unprocessed rays from a
// global counter instead of idling until the warp's longest ray ends.  Every lane takes one step per loop iteration.
#define SRT_NO_RAY 0xFFFFFFFFu
template <bool COUNT>
__global__ void __launch_bounds__(SRT_BLOCK, SRT_TRACE_MIN_BLOCKS) k_trace_rays(WaveParams P, uint32_t n, const float* __restrict__ o, const float* __restrict__ d,
                                                          const uint32_t* __restrict__ sorted_idx, float* __restrict__ t_out,
                                                          int32_t* __restrict__ tri_out, unsigned long long* counters, uint32_t* next_ray) {
    SceneRef sc;
    sc.nodes = P.nodes; sc.grid = P.grid; sc.nodes_global = true; sc.tris = P.tris; sc.units = nullptr; sc.n_units = 0; sc.flat_guard = 0.f; sc.flat_tol = 0.f; sc.mats = P.mats; sc.cie = nullptr; sc.bg = nullptr; sc.n_tris = P.n_tris;
    const uint32_t lane = threadIdx.x & 31;
    __shared__ int shared_stack[SRT_STACK_SMEM * SRT_BLOCK];
    int local_stack[SRT_STACK_MAX - SRT_STACK_SMEM];
    StackRef stack;
    stack.sm = shared_stack + threadIdx.x; stack.lm = local_stack; stack.stride = SRT_BLOCK;
    Walk w;
    walk_reset(w, false);
    uint32_t ray = SRT_NO_RAY;
    V3 ro = mk(0, 0, 0), rd = mk(0, 0, 0);
    GridRay g;
    g.om = g.inv = mk(0, 0, 0);
    uint32_t visits[2] = {0, 0};
    bool exhausted = false;
    while (true) {
        const uint32_t idle = __ballot_sync(0xffffffffu, ray == SRT_NO_RAY);
        if (!exhausted && (__popc(idle) >= SRT_REFILL_LANES || idle == 0xffffffffu)) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(next_ray, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + (uint32_t)__popc(idle) >= n) exhausted = true;
            const uint32_t mine = base + __popc(idle & ((1u << lane) - 1));
            if (ray == SRT_NO_RAY && base < n && mine < n) {
                ray = mine;
                ro = mk(o[3ull * mine], o[3ull * mine + 1], o[3ull * mine + 2]);
                rd = mk(d[3ull * mine], d[3ull * mine + 1], d[3ull * mine + 2]);
                g = grid_ray(sc, ro, rd);
                walk_reset(w, true);
                // same early answers as closest_hit: empty scene, NaN ray (Q1), single triangle
                bool done = sc.n_tris <= 0 || !(rd.x == rd.x && rd.y == rd.y && rd.z == rd.z && ro.x == ro.x && ro.y == ro.y && ro.z == ro.z);
                if (!done && sc.n_tris == 1) {
                    float t;
                    if (tri_test(sc.tris, ro, rd, w.closest, t)) { w.closest = t; w.best = 0; }
                    done = true;
                }
                if (done) {
                    t_out[mine] = w.best >= 0 ? w.closest : -1.0f;
                    tri_out[mine] = w.best >= 0 ? (int32_t)sorted_idx[w.best] : -1;
                    ray = SRT_NO_RAY;
                    walk_reset(w, false);
                }
            }
        }
        if (__all_sync(0xffffffffu, ray == SRT_NO_RAY)) {
            if (exhausted) break;
            continue;
        }
        // every lane takes the step (the leaf batches are a warp's decision); a lane without a ray has an empty walk
        if (!lbvh_step(sc, ro, rd, g, w, stack, 0xffffffffu, COUNT ? visits : nullptr) && ray != SRT_NO_RAY) {
            t_out[ray] = w.best >= 0 ? w.closest : -1.0f;
            tri_out[ray] = w.best >= 0 ? (int32_t)sorted_idx[w.best] : -1;
            ray = SRT_NO_RAY;
        }
    }
    if (COUNT) { atomicAdd(counters, (unsigned long long)visits[0]); atomicAdd(counters + 1, (unsigned long long)visits[1]); }
}

// The same queries through the render path's wide-leaf closest hit (scenes of <= 32 units): one ray per thread, the scene
// staged in shared memory exactly as k_wavefront stages it.  Exists so that tests can compare the conservative pre-test +
// exact re-test with the plain LBVH walk ray by ray (tests/test_gpu_parity.py).
__global__ void __launch_bounds__(SRT_BLOCK) k_trace_rays_flat(WaveParams P, uint32_t n, const float* __restrict__ o, const float* __restrict__ d,
                                                               const uint32_t* __restrict__ flat_to_orig, float* __restrict__ t_out, int32_t* __restrict__ tri_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SceneRef sc = load_scene<true, true>(P, smem);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float t = 0.f;
        const int tri = closest_hit<true>(sc, mk(o[3ull * i], o[3ull * i + 1], o[3ull * i + 2]), mk(d[3ull * i], d[3ull * i + 1], d[3ull * i + 2]), t);
        t_out[i] = tri >= 0 ? t : -1.0f;
        tri_out[i] = tri >= 0 ? (int32_t)flat_to_orig[tri] : -1;
    }
}

// ------------------------------------------------------------------------------ launchers
// mode 0: scene in global memory, LBVH walk; 1: scene staged in shared memory, LBVH walk;
// 2: scene staged in shared memory, <= 64 triangles, wide-leaf closest hit (no tree walk)
static size_t scene_smem_bytes(const WaveParams& P, int mode) {
    const size_t n_nodes = P.n_tris > 1 ? P.n_tris - 1 : 0;
    const size_t head = mode == 2 ? (size_t)P.n_units * sizeof(SrtFlatUnit) : n_nodes * sizeof(SrtWide);
    return head + (size_t)(mode == 2 ? 2 * P.n_units : P.n_tris) * sizeof(SrtTri) + (size_t)P.n_mats * sizeof(SrtMaterial) + 4 * SRT_NS * sizeof(float);
}
#define SRT_DISPATCH(KERNEL, MODE, GRID, SMEMB, ST, ...)                                  \
    do {                                                                                  \
        if ((MODE) == 2) KERNEL<true, true><<<GRID, SRT_BLOCK, SMEMB, ST>>>(__VA_ARGS__); \
        else if ((MODE) == 1) KERNEL<true, false><<<GRID, SRT_BLOCK, SMEMB, ST>>>(__VA_ARGS__); \
        else KERNEL<false, false><<<GRID, SRT_BLOCK, 0, ST>>>(__VA_ARGS__);               \
    } while (0)

LaunchTable make_launch_table() {
    LaunchTable t;
    t.smem_bytes = [](const WaveParams& P, int mode) { return scene_smem_bytes(P, mode); };
    t.configure = [](size_t bytes) {
        cudaError_t e = cudaSuccess;
        {
            const int b = (int)std::max<size_t>(bytes, 48 * 1024);
            const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
            e = cudaFuncSetAttribute(k_wavefront<true, false, 4>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wavefront<true, true, 4>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wavefront<false, false, 4>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wavefront<true, false, 3>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wavefront<true, true, 3>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_wavefront<false, false, 3>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_megakernel<true, false>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_megakernel<true, true>, attr, b);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_trace_rays_flat, attr, b);
        }
        return e;
    };
    t.init_slots = [](const WaveParams& P, cudaStream_t st) { k_init_slots<<<(P.nslots + 255) / 256, 256, 0, st>>>(P); };
    t.prior_cost = [](const WaveParams& P, cudaStream_t st) { k_prior_cost<<<(P.nslots + 255) / 256, 256, 0, st>>>(P); };
    t.wavefront = [](const WaveParams& P, int mode, int grid, size_t smem, cudaStream_t st) {
        // the queues always live in shared memory, also when the scene does not
        const int threads = (int)P.block_threads;  // <= SRT_WAVE_BLOCK, the launch bound the registers were allocated for
        if (P.min_blocks == 3) {
            if (mode == 2) k_wavefront<true, true, 3><<<grid, threads, smem, st>>>(P);
            else if (mode == 1) k_wavefront<true, false, 3><<<grid, threads, smem, st>>>(P);
            else k_wavefront<false, false, 3><<<grid, threads, smem, st>>>(P);
        } else if (mode == 2) k_wavefront<true, true, 4><<<grid, threads, smem, st>>>(P);
        else if (mode == 1) k_wavefront<true, false, 4><<<grid, threads, smem, st>>>(P);
        else k_wavefront<false, false, 4><<<grid, threads, smem, st>>>(P);
    };
    t.wavefront_blocks_per_sm = [](int mode, int min_blocks, int threads, size_t smem) {
        int n = 0;
        cudaError_t e;
        if (min_blocks == 3) {
            if (mode == 2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront<true, true, 3>, threads, smem);
            else if (mode == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront<true, false, 3>, threads, smem);
            else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront<false, false, 3>, threads, smem);
        } else if (mode == 2) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront<true, true, 4>, threads, smem);
        else if (mode == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront<true, false, 4>, threads, smem);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_wavefront<false, false, 4>, threads, smem);
        return e == cudaSuccess ? n : 0;
    };
    t.megakernel = [](const WaveParams& P, int mode, int grid, size_t smem, cudaStream_t st) { SRT_DISPATCH(k_megakernel, mode, grid, smem, st, P); };
    t.resolve = [](const float* acc, size_t plane, uint32_t img_w, uint32_t ox, uint32_t oy, uint32_t w, uint32_t h, uint32_t spp, unsigned char* rgb,
                   cudaStream_t st) { k_resolve<<<(w * h + 255) / 256, 256, 0, st>>>(acc, plane, img_w, ox, oy, w, h, spp, rgb); };
    t.resolve_slice = [](const float* acc, size_t plane, size_t first, uint32_t count, uint32_t out_plane, uint32_t spp, unsigned char* rgb, cudaStream_t st) {
        if (count) k_resolve_slice<<<(count + 255) / 256, 256, 0, st>>>(acc, plane, first, count, out_plane, spp, rgb);
    };
    t.trace_rays_flat = [](const WaveParams& P, uint32_t n, const float* o, const float* d, const uint32_t* flat_to_orig, float* t_out, int32_t* tri_out,
                           int grid, size_t smem, cudaStream_t st) {
        k_trace_rays_flat<<<grid, SRT_BLOCK, smem, st>>>(P, n, o, d, flat_to_orig, t_out, tri_out);
    };
    t.trace_rays = [](const WaveParams& P, uint32_t n, const float* o, const float* d, const uint32_t* sorted_idx, float* t_out, int32_t* tri_out,
                      unsigned long long* counters, uint32_t* next_ray, int grid, cudaStream_t st) {
        cudaMemsetAsync(next_ray, 0, sizeof(uint32_t), st);
        if (counters) k_trace_rays<true><<<grid, SRT_BLOCK, 0, st>>>(P, n, o, d, sorted_idx, t_out, tri_out, counters, next_ray);
        else k_trace_rays<false><<<grid, SRT_BLOCK, 0, st>>>(P, n, o, d, sorted_idx, t_out, tri_out, nullptr, next_ray);
    };
    return t;
}

}  // namespace SRT_FP_NS
}  // namespace srt
