/* srt.h -- C-ABI of libsrt.so, the B200-native (sm_100a) replacement for the hot path of
 * PieSil/CUDA-spectral-ray-tracer: LBVH build -> BVH traversal + ray/triangle -> hero-wavelength
 * scatter -> CIE XYZ -> sRGB film.
 *
 * The reference has no FFI layer; its seam is the C++ object API used by main.cpp:74-133.
 * Every entry point below names the reference interface it replaces (paths relative to the
 * reference repository).  Plain pointers and sizes only; all buffers are caller-owned HOST
 * memory unless the name says "device".  Functions return 0 (SRT_OK) on success unless noted;
 * on failure srt_last_error() describes why (the reference prints and exit(99)s instead,
 * utils/cuda_utility.cu:7-17 -- a library must not).
 */
#ifndef SRT_H
#define SRT_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SRT_OK 0
#define SRT_ERR_ARG 1
#define SRT_ERR_CUDA 2
#define SRT_ERR_STATE 3
#define SRT_ERR_NO_DEVICE 4

const char* srt_last_error(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
uint64_t srt_kernel_launch_count(void);
int srt_device_count(void);
/* select the CUDA device for subsequently created objects (one process per GPU: LOCAL_RANK) */
int srt_set_device(int device);

/* ------------------------------------------------------------------ params
 * replaces io/params.h:21-223 (class parameters) and :226-315 (param_manager singleton). */
typedef struct srt_params srt_params;
/* param_manager::getInstance(): process-wide instance, created on first use (params.h:228-234) */
srt_params* srt_params_instance(void);
/* destroys the singleton so the next srt_params_instance() starts from defaults (test helper) */
void srt_params_reset(void);
/* param_manager::parseArgs (params.h:236-304): same flags, same defaults, bad values keep the
 * previous value and print to stderr, unknown flags print "Unkown argument name" and continue. */
void srt_params_parse(srt_params*, int argc, char** argv);
unsigned srt_params_scene_id(const srt_params*);      /* getSceneId   :37 */
unsigned srt_params_xres(const srt_params*);          /* getXres      :41 */
unsigned srt_params_yres(const srt_params*);          /* getYres      :45 */
float srt_params_ar(const srt_params*);               /* getAR        :49 */
unsigned srt_params_xcsize(const srt_params*);        /* getXcsize    :53 */
unsigned srt_params_ycsize(const srt_params*);        /* getYcsize    :59 */
unsigned srt_params_nsamples(const srt_params*);      /* getNSamples  :65 */
unsigned srt_params_bounce_limit(const srt_params*);  /* getBounceLimit :69 */
int srt_params_log_active(const srt_params*);         /* logActive    :73 */
int srt_params_do_save(const srt_params*);            /* doSaveImage  :77 */
int srt_params_show_render(const srt_params*);        /* showRender   :81 */
const char* srt_params_img_title(const srt_params*);  /* getImgTitle  :28 */
const char* srt_params_log_subdir(const srt_params*); /* getLogSubdir :33 */

/* ------------------------------------------------------------------ camera
 * replaces rendering/camera_builder.cuh:13-71 and rendering/camera.cuh / camera.cu:7-58.
 * srt_camera carries exactly the fields of camera_data (rendering/rendering.cuh:19-37)
 * plus the background colour (camera.cuh:73). */
typedef struct { float x, y, z; } srt_vec3;
typedef struct {
    uint32_t width, height;
    srt_vec3 pixel_delta_u, pixel_delta_v, pixel00_loc;
    float defocus_angle;
    srt_vec3 camera_center, defocus_disk_u, defocus_disk_v;
    srt_vec3 background;
} srt_camera;
typedef struct srt_camera_builder srt_camera_builder;
srt_camera_builder* srt_camera_builder_create(void); /* defaults of camera_builder.cuh:64-70 */
void srt_camera_builder_destroy(srt_camera_builder*);
void srt_camera_builder_set_vfov(srt_camera_builder*, float);
void srt_camera_builder_set_lookfrom(srt_camera_builder*, float, float, float);
void srt_camera_builder_set_lookat(srt_camera_builder*, float, float, float);
void srt_camera_builder_set_vup(srt_camera_builder*, float, float, float);
void srt_camera_builder_set_defocus_angle(srt_camera_builder*, float);
void srt_camera_builder_set_focus_dist(srt_camera_builder*, float);
void srt_camera_builder_set_background(srt_camera_builder*, float, float, float);
/* getCamera(): resolution and aspect ratio come from the params singleton (camera_builder.cuh:57-61) */
int srt_camera_builder_get_camera(const srt_camera_builder*, srt_camera* out);
/* same with an explicit resolution (no reference counterpart; avoids the global) */
int srt_camera_builder_get_camera_res(const srt_camera_builder*, uint32_t w, uint32_t h, srt_camera* out);

/* ------------------------------------------------------------------ scene
 * replaces scene/scene.cuh:103-176 (scene_manager) and the <<<1,1>>> world/BVH kernels of
 * scene/scene.cu:9-71.  Triangles and material spectra are built on the host with the
 * reference's float operation order, uploaded, and the LBVH is built on the device. */
typedef struct srt_scene srt_scene;
#define SRT_MAT_LAMBERTIAN 0
#define SRT_MAT_METALLIC 1
#define SRT_MAT_DIELECTRIC 2
#define SRT_MAT_EMISSIVE 4
typedef struct {
    uint32_t type;            /* SRT_MAT_* (materials/material.cuh:16-22) */
    float color[3];           /* sRGB, lambertian / metallic / emissive */
    float fuzz;               /* metallic */
    float emission_power;     /* emissive: spectrum is scaled by power^2 (color_to_spectrum.cuh:181) */
    float sellmeier_b[3];     /* dielectric */
    float sellmeier_c[3];     /* dielectric; ignored in reference-compatible mode (material.cuh:67 stores B twice) */
} srt_material_desc;
/* scene_manager(): scene id 0 Cornell, 1 Prism, 2 Different Materials (io/params.h:15-19);
 * unknown ids build Cornell like the reference's default: branch (scene.cu:40-43). */
srt_scene* srt_scene_create(unsigned scene_id);
/* seeded random-triangle soup for BASELINE.json configs[3] (no reference counterpart) */
srt_scene* srt_scene_create_soup(uint32_t n_tris, uint64_t seed);
/* arbitrary triangle list: verts = n_tris*9 floats (v0 v1 v2), one material index per triangle */
srt_scene* srt_scene_create_mesh(const float* verts, const uint32_t* mat_index, uint32_t n_tris,
                                 const srt_material_desc* mats, uint32_t n_mats);
/* Wavefront OBJ (v / f records, fan triangulation), all faces get material 0 of `mats` */
srt_scene* srt_scene_create_obj(const char* path, const srt_material_desc* mats, uint32_t n_mats);
/* Stanford PLY (ascii or binary_little_endian; vertex x y z + face vertex_indices lists, fan triangulation) */
srt_scene* srt_scene_create_ply(const char* path, const srt_material_desc* mats, uint32_t n_mats);
void srt_scene_destroy(srt_scene*);
/* getResult(): returns 1 when the world was created, *msg = "World created" or the error */
int srt_scene_result(const srt_scene*, const char** msg);
/* getCamPtr(): the scene's camera (scene.cu:259-320) at the params-singleton resolution */
int srt_scene_camera(const srt_scene*, srt_camera* out);
int srt_scene_camera_res(const srt_scene*, uint32_t w, uint32_t h, srt_camera* out);
uint32_t srt_scene_num_tris(const srt_scene*);
uint32_t srt_scene_num_materials(const srt_scene*);
/* pre-test units of the wide-leaf closest hit (a triangle, or a coplanar parallelogram pair such as every tri_quad of the
 * reference emits); 0 when the scene is too large for it (> 64 triangles or > 32 units) and is traversed through the LBVH */
uint32_t srt_scene_num_units(const srt_scene*);
/* 1 = dielectrics reproduce material.cuh:67 (C := B), the default; 0 = physical Sellmeier.
 * Must be set before srt_scene_create*. */
void srt_set_ref_compat(int on);
/* the Sellmeier coefficient tables the reference carries (refraction/sellmeier.cuh:6-13) for srt_material_desc */
#define SRT_GLASS_BK7 0
#define SRT_GLASS_FUSED_SILICA 1
#define SRT_GLASS_FLINT 2
int srt_glass_coefficients(int which, float b[3], float c[3]);
/* parity dumps.  tris_f: n*22 floats (v0 v1 v2 normal D bbox[xmin xmax ymin ymax zmin zmax] pad3),
 * tris_i: n*3 ints (clockwise, aa_plane, material).  mats_f: m*108 floats (col3 fuzz power B3 C3
 * spectrum95 pad2), mats_i: m ints (type). */
int srt_scene_get_tris(const srt_scene*, float* tris_f, int32_t* tris_i);
int srt_scene_get_materials(const srt_scene*, float* mats_f, int32_t* mats_i);
/* device-built LBVH, downloaded: codes[n] (original triangle order), sorted_idx[n], left[n-1],
 * right[n-1], parent[2n-1], node_boxes[(2n-1)*6], scene_box[6]; node ids as in DESIGN.md
 * (internal 0..n-2, leaf k = n-1+k).  Any pointer may be NULL. */
int srt_scene_get_lbvh(const srt_scene*, uint32_t* codes, uint32_t* sorted_idx, int32_t* left, int32_t* right,
                       int32_t* parent, float* node_boxes, float* scene_box);
/* rebuilds the LBVH `repeats` times on device-resident triangles and returns the per-stage
 * CUDA-event times of the LAST build in ms: [total, bounds+morton, sort, tree = hierarchy+refit+binary nodes, collapse = the 4-wide traversal nodes] */
int srt_scene_rebuild_lbvh(srt_scene*, int repeats, float ms_out[5]);
/* closest-hit queries through the device BVH (BASELINE.json configs[3] traversal rays/s):
 * o,d: n*3 floats; t_out[n] (-1 on miss... see tri_out), tri_out[n] = triangle index or -1.
 * ms_out (optional) = kernel time from CUDA events. */
int srt_scene_trace_rays(const srt_scene*, uint32_t n, const float* o, const float* d, float* t_out,
                         int32_t* tri_out, float* ms_out);

/* FP mode of the kernels behind the srt_scene_trace_rays* queries: 1 (default) = strict (-fmad=false), 0 = fast (FMA contraction) */
void srt_set_query_fp_mode(int strict);
/* the same queries answered by the wide-leaf closest hit the renderer uses for scenes of <= 32 units (conservative
 * pre-test over all units, exact reference arithmetic on the survivors); fails when srt_scene_num_units() is 0.
 * Must agree with srt_scene_trace_rays ray by ray. */
int srt_scene_trace_rays_flat(const srt_scene*, uint32_t n, const float* o, const float* d, float* t_out, int32_t* tri_out);

/* same, plus visits_out = {BVH nodes visited, leaf triangles tested} summed over all rays (counted in an
 * extra untimed pass): the algorithmic traffic of the walk is 32 B per node visit + 48 B per triangle test */
int srt_scene_trace_rays_counted(const srt_scene*, uint32_t n, const float* o, const float* d, float* t_out,
                                 int32_t* tri_out, float* ms_out, uint64_t visits_out[2]);

/* ------------------------------------------------------------------ render manager
 * replaces rendering/render_manager.cuh:37-173 + render_manager.cu:3-132 and the renderer
 * behind it (rendering/rendering.cuh:39-155, rendering.cu:151-357). */
typedef struct srt_render_manager srt_render_manager;
/* render_manager(bvh**, material*, camera*, frame_buffer*): borrows the scene and the three
 * caller-owned planar float channels (frame_buffer.cuh:6-44: raster order y*W+x, values 0..255). */
srt_render_manager* srt_render_manager_create(srt_scene*, const srt_camera*, float* fb_r, float* fb_g, float* fb_b);
void srt_render_manager_destroy(srt_render_manager*);
int srt_rm_init_renderer(srt_render_manager*, unsigned bounce_limit, unsigned samples_per_pixel); /* :121-132 */
/* init_device_params(chunk_w, chunk_h) (:91-102); (0,0) = init_device_params() full image (:104-119) */
int srt_rm_init_device_params(srt_render_manager*, unsigned chunk_w, unsigned chunk_h);
int srt_rm_is_ready_to_render(const srt_render_manager*); /* isReadyToRender */
int srt_rm_is_done(const srt_render_manager*);            /* isDone */
unsigned srt_rm_im_width(const srt_render_manager*);      /* getImWidth */
unsigned srt_rm_im_height(const srt_render_manager*);     /* getImHeight */
/* step(): renders the next chunk on the device and hands it to the consumer slot; returns 1
 * while more chunks remain, 0 after the last, <0 on error (render_manager.cu:3-66) */
int srt_rm_step(srt_render_manager*);
/* update_fb(): copies the oldest finished chunk into the caller's frame buffer; returns 1 while
 * more chunks will follow, 0 after the last (render_manager.cuh:68-142) */
int srt_rm_update_fb(srt_render_manager*);
/* render_cycle(): starts the worker thread that calls step() until done (:160-167) */
int srt_rm_render_cycle(srt_render_manager*);
int srt_rm_end_render(srt_render_manager*); /* joins the worker (:169-174) */
/* convenience: render_cycle + update_fb loop + end_render (what main.cpp:29-43 does) */
int srt_rm_render_all(srt_render_manager*);

/* additions needed for grading / multi-GPU (no reference counterpart) */
#define SRT_OPT_FP_MODE 1      /* 0 = fast (FMA contraction, like the reference's nvcc build), 1 = strict (-fmad=false, matches the host oracle) */
#define SRT_OPT_PIPELINE 2     /* 0 = wavefront (default), 1 = per-pixel persistent megakernel */
#define SRT_OPT_TILE_W 3       /* image tile: the unit of multi-GPU ownership and of pixel numbering (powers of two; default 0 = automatic: 32x32, 32x16 or 16x16 by pixels per rank) */
#define SRT_OPT_TILE_H 4
#define SRT_OPT_RANK 5         /* ... this process renders the tiles with (tile_x + 5 tile_y) % world == rank */
#define SRT_OPT_WORLD 6
#define SRT_OPT_KERNEL_TIMING 8 /* 1 = bracket every kernel launch with CUDA events (per-kernel totals in srt_stats) */
#define SRT_OPT_BLOCK_SLOTS 11   /* paths in flight per wavefront block: power of two in [32, 4096], 0 = automatic */
#define SRT_OPT_BLOCK_THREADS 12 /* threads per wavefront block: 0 = 256 */
#define SRT_OPT_STRATIFIED 13    /* 1 = stratified pixel sampler (renderer::get_ray_stratified_sample, rendering/rendering.cu:58-64,89-118,
                                    which the reference carries but never calls): sample k takes sub-cell (k % n, k / n) of an n x n
                                    grid, spp must be n*n; 0 (default) = the reference's sampler */
#define SRT_OPT_ROUNDS 14        /* launches per chunk: the samples of a pixel are rendered in K rounds [0,b1) [b1,b2) .. [b,spp), each 8x
                                    longer than the one before; the pixel's XORWOW state waits in HBM between rounds like it does between
                                    chunks (rendering/rendering.cu:209,232), so the film does not depend on K.  From the second round on
                                    pixels are handed out most-passes-per-sample first, which is what ends all blocks together.
                                    0 (default) = automatic (1 below 16 spp, 2, 3 from 512 spp), 1 = a single launch */
#define SRT_OPT_L2_PERSIST 15    /* 1 (default) = keep the state of the paths in flight in a persisting-L2 window (a process-wide
                                    cudaDeviceSetLimit while a renderer exists), 0 = leave the device's L2 configuration alone */
#define SRT_OPT_PASS_LOG 16      /* debug: 1 = the first 8 blocks of the LAST k_wavefront launch record a time stamp and their queue lengths
                                    every pass; read back with srt_rm_get_pass_log */
#define SRT_OPT_SCHED_FLAGS 17   /* debug / ablation, bit mask: 1 = first round in slot order (no first-guess cost order),
                                    4 = queue remainders are not pooled into mixed warps,
                                    32 / 64 = force the 64-register (4 blocks per SM) / 80-register (3 blocks per SM) build of the
                                    wavefront kernel (default: the second one only when the rank's pixels do not fill the
                                    resident blocks with more than 256 paths each) */
#define SRT_OPT_TRAVERSAL 10     /* 0 auto (wide-leaf closest hit when the scene has <= 64 triangles), 1 force the LBVH walk
                                    (scene in shared memory), 3 force the LBVH walk with the scene in global memory */
int srt_rm_set_option(srt_render_manager*, int option, int value);
/* pre-tonemap film of the whole image: 3 raster planes of XYZ (mean over spp); downloaded on demand.
 * After srt_rm_exchange_film the call is collective (the reduced slices are all-gathered first). */
int srt_rm_get_xyz(srt_render_manager*, float* xyz);
/* device pointer to the XYZ SUM film (3 planes, W*H floats each, zeros outside owned tiles);
 * a caller that owns an NCCL communicator may reduce it in place and then call srt_rm_resolve_film */
float* srt_rm_device_film(srt_render_manager*);
/* tonemaps the (possibly reduced) device film into the caller's frame buffer and XYZ planes */
int srt_rm_resolve_film(srt_render_manager*);
/* rewinds the manager to its first chunk with an empty film and freshly seeded RNG slots, so the
 * same image can be rendered again (the reference's render_manager is one-shot) */
int srt_rm_restart(srt_render_manager*);
typedef struct {
    uint64_t samples, rays, kernel_launches;
    uint64_t wavefront_launches; /* k_wavefront launches so far (rounds x chunks) */
    uint64_t rounds;             /* launches per chunk (SRT_OPT_ROUNDS after the automatic choice) */
    double render_ms;   /* CUDA-event time of all step() work on the device */
    double lbvh_ms;     /* last LBVH build of the scene */
    /* filled when SRT_OPT_KERNEL_TIMING is on: summed CUDA-event durations per kernel family ... */
    double wavefront_ms, megakernel_ms, order_ms, other_ms; /* k_wavefront | k_megakernel | pixel-order sort | RNG seeding */
    /* ... and the end-of-launch drain of k_wavefront: last block exit minus the moment the first path slot
     * found no pixel left to fetch (device globaltimer), summed over launches */
    double drain_ms;
    double exchange_ms; /* srt_rm_exchange_film, CUDA events: the collective (reduce-scatter of the XYZ planes, incl. waiting for the slowest rank) */
    double film_out_ms; /* srt_rm_exchange_film, CUDA events: slice tonemap + byte gather to rank 0 + device-to-host copy */
} srt_stats;
int srt_rm_get_stats(const srt_render_manager*, srt_stats* out);
/* SRT_OPT_PASS_LOG read-out: out = 8 blocks x 8192 passes x 4 uint32 {globaltimer ns (low 32 bits), regenerate, lambertian,
 * metallic | dielectric << 16}; unused entries are zero */
int srt_rm_get_pass_log(srt_render_manager*, uint32_t* out);

/* ------------------------------------------------------------------ multi-GPU film exchange
 * No reference counterpart (the reference is single-GPU); BASELINE.json north_star: "rendering shards by image
 * tiles across the GPUs of one box, with per-GPU film buffers combined by an NCCL reduce/gather over NVLink".  One process per
 * GPU; libsrt owns the NCCL communicator (libnccl.so.2 is opened on first use).  Rank 0 creates the id and hands its 128 bytes to
 * the other ranks by any out-of-band channel (file, socket, MPI, torch.distributed ...). */
#define SRT_NCCL_UNIQUE_ID_BYTES 128
typedef struct srt_comm srt_comm;
int srt_comm_get_unique_id(unsigned char id[SRT_NCCL_UNIQUE_ID_BYTES]);
/* ncclCommInitRank on the current device (srt_set_device first); collective over all `world` ranks */
srt_comm* srt_comm_create(const unsigned char id[SRT_NCCL_UNIQUE_ID_BYTES], int rank, int world);
void srt_comm_destroy(srt_comm*);
int srt_comm_rank(const srt_comm*);
int srt_comm_world(const srt_comm*);
/* *v = max over ranks of *v (device-timed milliseconds in bench.py); collective, doubles as a barrier */
int srt_comm_max_double(srt_comm*, double* v);
int srt_nccl_version(void); /* e.g. 22703; 0 when libnccl.so.2 cannot be loaded */
/* attach before srt_rm_init_device_params: tile ownership (SRT_OPT_RANK / SRT_OPT_WORLD) follows the communicator and
 * the film planes are laid out for the in-place reduce-scatter */
int srt_rm_set_comm(srt_render_manager*, srt_comm*);
/* collective, after the last srt_rm_step: reduce-scatter (sum) of the XYZ films, every rank tonemaps its 1/world of the
 * pixels, the byte planes are gathered on rank 0 and land in rank 0's frame buffer.  Replaces srt_rm_update_fb /
 * srt_rm_resolve_film in a multi-GPU render. */
int srt_rm_exchange_film(srt_render_manager*);
/* order-independent 64-bit checksum of the XYZ-sum film bits.  Without a communicator: of this manager's film.  With one
 * (after srt_rm_exchange_film, collective): of the reduced film, the same value on every rank -- and equal to the
 * single-GPU value exactly when the reduced film equals the single-GPU film bit for bit. */
int srt_rm_film_checksum(srt_render_manager*, uint64_t* out);
/* returns every cached, currently unused device / pinned block to the driver (the library keeps freed buffers for reuse) */
void srt_trim_caches(void);

/* measured FP32 FMA issue peak of the current device in TFLOP/s (dependent-free FFMA chains on every
 * SM, CUDA-event timed): the roofline denominator for the instruction-bound render kernels */
double srt_measure_fp32_tflops(void);
/* measured device-to-device copy bandwidth in GB/s (read + write bytes) over `mbytes` MiB */
double srt_measure_copy_gbs(uint32_t mbytes);
/* measured L2 read bandwidth in GB/s (all SMs stream a 32 MiB buffer that stays in L2): the second roofline of the LBVH walk */
double srt_measure_l2_read_gbs(void);

/* ------------------------------------------------------------------ output (io/save_image.cpp:8-13, io/io.cuh:10-23) */
int srt_write_ppm(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h);
int srt_write_bmp(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h);

#ifdef __cplusplus
}
#endif
#endif
