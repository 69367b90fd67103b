// Shared host/device plain-old-data layouts of the B200 path tracer (see DESIGN.md "HBM layout").
#ifndef SRT_TYPES_H
#define SRT_TYPES_H
#include <stdint.h>

#define SRT_N_WL 7           // hero + 6 rotations (reference ray/ray.cuh:12)
#define SRT_NS 95            // spectrum samples, 360..830 nm @ 5 nm (utils/cie_const.cuh:8)
#define SRT_EPSILON 0.0001f  // self-intersection offset (materials/material.cuh:14)

// material type tags (reference materials/material.cuh:16-22)
#define SRT_LAMBERTIAN 0u
#define SRT_METALLIC 1u
#define SRT_DIELECTRIC 2u
#define SRT_EMISSIVE 4u

// ---- device triangle, 48 B = 3 x 16-B vectors ------------------------------------------
// q0 = plane (nx, ny, nz, D)
// q1 = (w0, h0, w1, h1)   projected vertices on the triangle's 2-D test plane
// q2 = (w2, h2, bits, prio)  prio = rank in the reference's test order (equal-t tie-break); bits: [15:0] material index, [16] clockwise, [18:17] w axis,
//                         [20:19] h axis, [23:21] material type
struct alignas(16) SrtTri {
    float nx, ny, nz, D;
    float w0, h0, w1, h1;
    float w2, h2;
    uint32_t bits;
    uint32_t prio;
};
#define SRT_TRI_MAT(bits) ((bits) & 0xFFFFu)
#define SRT_TRI_CW(bits) (((bits) >> 16) & 1u)
#define SRT_TRI_WAX(bits) (((bits) >> 17) & 3u)
#define SRT_TRI_HAX(bits) (((bits) >> 19) & 3u)
#define SRT_TRI_MTYPE(bits) (((bits) >> 21) & 7u)

#define SRT_FLAT_MAX_TRIS 64  // scenes up to this size are one wide leaf: no tree walk at all

// ---- wide-leaf pre-test unit, 64 B: one triangle or one parallelogram pair (host/flat_leaf.cpp) ----
// q0 = plane; q1 = (A.xyz, a_w + eps); q2 = (B.xyz, b_w + eps); q3 = (c1, c2, c3, near)
// near: 0, or -- pairs whose two triangles store planes that differ in the last bits -- how close to the plane a grazing
// ray must start for the second triangle to stay a candidate without a verdict (host/flat_leaf.cpp)
// unit u covers flat triangle positions 2u (first half) and 2u+1 (second half, if c2 >= 0)
struct alignas(16) SrtFlatUnit {
    float nx, ny, nz, D;
    float ax, ay, az, aw;
    float bx, by, bz, bw;
    float c1, c2, c3, near;
};
#define SRT_FLAT_MAX_UNITS 32

// ---- 4-wide traversal node, 64 B = two 32-byte sectors, two LDG.256: the boxes and refs of a binary node's (up to) four grandchildren ----
// Divergent node fetches cost by the sector and by the request (tools/micro/gather_probe.cu: 64-B records as 4 x LDG.128 gather
// at 96 G records/s on a B200, as 2 x LDG.256 at 119 G, 32-B records as one LDG.256 at 250 G), and a walk pays its per-step
// overhead (stack, leaf batches, ray refill) once per node it opens.  So the walk does not read binary nodes with float boxes
// (64 B for two children): one streaming pass after the bottom-up build (lbvh.cu k_collapse4) writes, under the SAME index as the
// binary node it collapses, slots = for each child of node i, the child itself when it is a leaf, else the child's two children;
// a step of the walk tests four boxes and descends two levels of the binary tree.
//   w[3k .. 3k+2] = slot k: (xmin | xmax << 16), (ymin | ymax << 16), (zmin | zmax << 16)
//   w[12 + k]     = slot k's ref: >= 0 wide node index, < 0 leaf with triangle index ~ref, SRT_WIDE_EMPTY = unused slot
// Boxes live on a uniform 16-bit grid over the scene box (cell = largest extent / 65529, grid coordinate g(x) = (x - lo) / cell + 3),
// min = floor(g) - 3, max = ceil(g) + 3: three cells of margin on every side.  The walk's roundings (csrc/cuda/trace_impl.cuh
// grid_ray: the origin snaps to the grid's integer lattice, <= 0.5 cell; build and transform, < 0.05) stay inside that margin,
// so a stored box always contains, for the walk's arithmetic, the exact float box the reference's closest hit is defined on.
// The exact boxes (Karras artefacts, bit-exact vs the oracle) live in DeviceScene::node_box and never enter the walk.
struct alignas(64) SrtWide {
    uint32_t box[12];
    int32_t ref[4];
};
#define SRT_WIDE_EMPTY 0x7fffffff
#define SRT_GRID_MARGIN 3
#define SRT_GRID_CELLS 65529.0f   // 65535 - 2 * margin
#define SRT_GRID_OFFSET 3.0f     // = margin: the scene box starts at grid coordinate 3, so min - margin >= 0

// ---- device material, 400 B: 95-sample spectrum + parameters ------------------------------
struct alignas(16) SrtMaterial {
    float spec[SRT_NS];
    float fuzz;
    float sellB[3];
    float sellC[3];
    uint32_t type;
};

struct SrtCamera {
    uint32_t width, height;
    float du[3], dv[3], p00[3];
    float defocus_angle;
    float center[3], disk_u[3], disk_v[3];
};

#endif
