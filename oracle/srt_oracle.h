/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's hot path
 * (PieSil/CUDA-spectral-ray-tracer).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product (libsrt.so) never does.
 *
 * Pinned against the real reference host-compiled in oracle/_ref (tests/test_oracle_pin.py):
 * triangles, materials, camera, BVH topology, per-ray KATs and whole images are bit-exact.
 * PARITY UNPINNED parts: (1) the three sRGB->spectrum table cells the Cornell scenes read
 * (the reference's table file is missing from the mount; see rgb2spec.c); (2) the LBVH
 * (lbvh_oracle.c) -- the reference has no LBVH at all, so there the oracle is the spec.
 */
#ifndef SRT_ORACLE_H
#define SRT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z; } ov3;

enum { O_LAMBERTIAN = 0, O_METALLIC = 1, O_DIELECTRIC = 2, O_EMISSIVE = 4, O_NO_MAT = 6 };
enum { O_AA_NONE = 0, O_AA_XY = 1, O_AA_YZ = 2, O_AA_XZ = 3 };

typedef struct {
    ov3 v[3];
    int clockwise, aa_plane;
    uint32_t mat;
    float bb[6]; /* xmin xmax ymin ymax zmin zmax (padded) */
    ov3 n;
    float D;
} otri;

typedef struct {
    int type;
    ov3 col;
    float fuzz, power;
    float B[3], C[3];
    float spec[95];
} omat;

typedef struct {
    int w, h;
    ov3 du, dv, p00;
    float defocus_angle;
    ov3 center, disk_u, disk_v;
    ov3 background;
} ocam;

typedef struct { uint32_t d, v[5]; } orng;

typedef struct {
    uint64_t samples, rays, box_tests, tri_tests, rng_draws, interps;
    uint64_t scatter_lambert, scatter_metal, scatter_dielectric, rejection_iters;
    uint64_t end_miss, end_limit, end_emissive, end_absorbed, nan_rays;
} ocounters;

typedef struct oscene oscene;

/* scene ids of the reference: 0 Cornell, 1 Prism, 2 Different Materials (io/params.h:15-17) */
oscene* srt_oracle_scene_create(int scene_id);
/* seeded random-triangle soup (BASELINE.json configs[3]); NOT a reference scene */
oscene* srt_oracle_scene_soup(int n_tris, uint64_t seed);
void srt_oracle_scene_destroy(oscene*);
int srt_oracle_scene_ntris(const oscene*);
int srt_oracle_scene_nmats(const oscene*);
const otri* srt_oracle_scene_tris(const oscene*);
const omat* srt_oracle_scene_mats(const oscene*);
/* the reference's own (serial, random-axis median split) BVH in preorder: leaf -> tri index, internal -> -1 */
int srt_oracle_scene_refbvh_preorder(const oscene*, int* out);
void srt_oracle_scene_reforder(const oscene*, int* out);

/* camera of the three reference scenes (scene/scene.cu:259-320) at the given resolution */
void srt_oracle_camera_default(int w, int h, ocam* out);
void srt_oracle_camera_make(int w, int h, float vfov, ov3 lookfrom, ov3 lookat, ov3 vup,
                            float defocus_angle, float focus_dist, ov3 background, ocam* out);
/* yres = uint(xres / ar) as io/params.h:176-180 */
int srt_oracle_yres(int xres, float ar);

/* XORWOW (cuRAND device API) */
void srt_oracle_rng_init(uint32_t seed, orng* s);
uint32_t srt_oracle_rng_next(orng* s);
float srt_oracle_rng_uniform(orng* s);

/* single-function KATs */
int srt_oracle_tri_hit(const otri* t, const float o[3], const float d[3], float tmin, float tmax, float out[10]);
int srt_oracle_aabb_hit(const float box6[6], const float o[3], const float d[3], float tmin, float tmax);
int srt_oracle_bvh_hit(const oscene*, const float o[3], const float d[3], float out[10]);
/* closest hit by brute force over all triangles in index order (topology-free definition) */
int srt_oracle_brute_hit(const oscene*, const float o[3], const float d[3], float out[10], int* tri_index);
int srt_oracle_scatter(const oscene*, int mat, float ray_io[21], const float rec_in[8], uint32_t rng[6]);
float srt_oracle_sellmeier(const float b[3], const float c[3], float lambda);
void srt_oracle_glass(int which, float b[3], float c[3]);       /* sellmeier.cuh:6-13: 0 BK7, 1 fused silica, 2 flint */
void srt_oracle_set_physical_sellmeier(int on);                 /* scenes created afterwards: material.cuh:67 as shipped (0) or fixed (1) */
float srt_oracle_spectrum_interp(const float* table95, float lambda);
void srt_oracle_spectrum_to_xyz(const float wl[7], const float pw[7], int nvalid, float xyz[3]);
void srt_oracle_tonemap(const float xyz_mean[3], float rgb255[3]);
void srt_oracle_get_ray(const ocam* cam, uint32_t i, uint32_t j, uint32_t rng[6], float out[13]);
void srt_oracle_color_spectrum(float r, float g, float b, int emissive, float power, float out95[95]);

/* whole render (rendering/rendering.cu:151-235 + render_manager.cu:3-66 chunk loop).
 * chunk_w/chunk_h = 0 -> single full-image chunk.  rgb: 3 raster planes 0..255;
 * xyz (optional): 3 raster planes of XYZ/spp.  counters optional.  nthreads<=0: all cores. */
int srt_oracle_render(const oscene*, const ocam*, int spp, int bounce_limit, int chunk_w, int chunk_h,
                      float* rgb, float* xyz, ocounters* counters, int nthreads);

/* same, with the opt-in stratified pixel sampler (rendering.cu:58-64, 89-118; dormant in the reference):
 * sample k -> sub-cell (k % n, k / n), n*n == spp required (returns -2 otherwise) */
int srt_oracle_render_opts(const oscene*, const ocam*, int spp, int bounce_limit, int chunk_w, int chunk_h, int stratified,
                           float* rgb, float* xyz, ocounters* counters, int nthreads);
void srt_oracle_get_ray_stratified(const ocam* cam, uint32_t i, uint32_t j, uint32_t sx, uint32_t sy, float recip_sqrt_spp,
                                   uint32_t rng[6], float out[13]);

/* parity debugging: per-sample XYZ (3*spp floats) of one pixel of a single-chunk render; brute = 1 replaces the reference's
 * pruned BVH walk by a closest hit over all triangles */
int srt_oracle_debug_pixel(const oscene*, const ocam*, int spp, int bounce_limit, int i, int j, int brute, float* xyz_per_sample);

/* render only the pixels whose (x/tile_w + 5*(y/tile_h)) % world == rank (multi-GPU
 * tile ownership test); other pixels are left zero.  Single full-image chunk. */
int srt_oracle_render_tiles(const oscene*, const ocam*, int spp, int bounce_limit, int tile_w, int tile_h,
                            int rank, int world, float* xyz_sum);

#ifdef __cplusplus
}
#endif
#endif
