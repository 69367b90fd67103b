// Host side of libsrt.so: the mirror of the reference's L3-L5 objects (params, camera_builder,
// scene_manager, render_manager) re-designed around flat arrays that upload in one copy.
#pragma once
#include <cstdint>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>

#include "../common/srt_types.h"
#include "../../../include/srt.h"

typedef struct CUstream_st* cudaStream_t;  // as in driver_types.h: host files need not include the CUDA headers

namespace srt {

void set_error(const std::string& msg);
const std::string& last_error();

struct vec3f {
    float x = 0, y = 0, z = 0;
    vec3f() = default;
    vec3f(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

// ---------------------------------------------------------------- params (io/params.h)
struct Params {
    std::string image_title, log_subdir;
    unsigned scene = 0, xres = 600, yres = 600;
    float ar = 1.0f;
    unsigned xcsize = 0, ycsize = 0, n_samples = 500, bounce_limit = 10;
    bool do_log = false, show_render = true, do_save = false;
    mutable std::string title_cache;

    Params() { reset_yres(); }
    void reset_yres();
    unsigned get_xcsize() const;
    unsigned get_ycsize() const;
    const std::string& get_title() const;
    void parse(int argc, char** argv);
    static Params& instance();
    static void reset_instance();
};

// ---------------------------------------------------------------- camera
struct CameraBuilder {
    float vfov = 90.0f;
    vec3f lookfrom{0, 0, -1}, lookat{0, 0, 0}, vup{0, 1, 0};
    float defocus_angle = 0, focus_dist = 10;
    vec3f background{0, 0, 0};
    srt_camera build(uint32_t w, uint32_t h) const;
};

// ---------------------------------------------------------------- geometry (host build, bit-faithful)
enum AAPlane : int { AA_NONE = 0, AA_XY = 1, AA_YZ = 2, AA_XZ = 3 };
struct HostTri {
    vec3f v[3];
    vec3f normal;
    float D = 0;
    int clockwise = 0;
    int aa_plane = AA_NONE;
    uint32_t mat = 0;
    float bbox[6] = {0, 0, 0, 0, 0, 0};
    void derive();  // normal, plane constant, winding, projection plane, padded box
    SrtTri pack(uint32_t mat_type, uint32_t prio) const;
};
struct HostMaterial {
    uint32_t type = SRT_LAMBERTIAN;
    vec3f color;
    float fuzz = 1.0f, power = 0.0f;
    float B[3] = {0, 0, 0}, C[3] = {0, 0, 0};
    float spec[SRT_NS];
    void bake_spectrum();
    static HostMaterial from_desc(const srt_material_desc& d, bool ref_compat);
};

class TriangleSoup {  // growing triangle list with the reference's composite builders
public:
    std::vector<HostTri> tris;
    size_t add_tri(vec3f a, vec3f b, vec3f c, uint32_t mat, bool as_vectors);
    size_t add_quad(vec3f Q, vec3f u, vec3f v, uint32_t mat);
    size_t add_box(vec3f a, vec3f b, const uint32_t mats[6]);
    size_t add_pyramid(vec3f Q, vec3f u, vec3f v, vec3f w, uint32_t mat);
    size_t add_prism(vec3f Q, vec3f u, vec3f v, vec3f w, uint32_t mat);
    vec3f quad_center(size_t first) const;
    vec3f box_center(size_t first) const;
    vec3f prism_centroid(size_t first) const;
    void rotate_y_about(size_t first, size_t count, vec3f pivot, float theta);
    void translate(size_t first, size_t count, vec3f d, bool rederive);
    void rederive(size_t first, size_t count);
};

struct SceneDescription {
    std::vector<HostTri> tris;
    std::vector<HostMaterial> mats;
    CameraBuilder camera;
};
SceneDescription make_reference_scene(unsigned scene_id, bool ref_compat);
SceneDescription make_soup_scene(uint32_t n, uint64_t seed);
bool glass_coefficients(int which, float b[3], float c[3]);
bool load_obj(const char* path, std::vector<float>& verts9);
bool load_ply(const char* path, std::vector<float>& verts9);

// sRGB -> sigmoid-polynomial coefficients, computed on demand (replaces the 9.4 MB table)
namespace rgb2spec {
float scale(int k);                                          // Scale[k], res 64
bool cell(int l, int k, int j, int i, float out[3]);         // Data[l][k][j][i][0..2], cached
void coeffs_nearest(vec3f rgb, float out_c[3]);              // device-path lookup (nearest cell)
void coeffs_trilinear(vec3f rgb, float out_c[3]);            // host-path lookup (background)
}  // namespace rgb2spec
const float* cie_table(int which);  // 0 x, 1 y, 2 z, 3 normalised D65 (95 floats each)
float spectrum_interp_host(const float* table95, float lambda);
void background_spectrum(vec3f rgb, float out95[SRT_NS]);

std::vector<uint32_t> reference_test_order(const std::vector<HostTri>& tris);
struct FlatLeaf {
    std::vector<SrtFlatUnit> units;   // <= 32
    std::vector<SrtTri> tris;         // 2 per unit
    std::vector<uint32_t> to_orig;    // flat position -> original triangle index
    float guard = 0.f;                // grazing threshold on |n.d| / |d| for pairs whose partner plane differs from the head's in the last bits
    float tol = 0.f;                  // `behind` tolerance on |D - n.o| (the largest over the units)
    uint32_t ymask = 0;               // bit g: units 4g..4g+3 are all y-aligned (normal +-y, frame rows without a y term) and stored in the short form
};
bool build_flat_leaf(const std::vector<HostTri>& tris, const std::vector<HostMaterial>& mats, const std::vector<uint32_t>& prio,
                     double origin_l1_bound, FlatLeaf& out);

// ---------------------------------------------------------------- device objects (defined in cuda/*.cu)
struct Comm;            // NCCL communicator owned by libsrt (host/nccl_comm.cpp)
bool comm_unique_id(unsigned char id[SRT_NCCL_UNIQUE_ID_BYTES]);
Comm* comm_create(const unsigned char id[SRT_NCCL_UNIQUE_ID_BYTES], int rank, int world);
void comm_destroy(Comm*);
int comm_rank(const Comm*);
int comm_world(const Comm*);
int comm_device(const Comm*);
int comm_nccl_version();
bool comm_film_reduce_scatter(Comm*, float* acc, size_t plane_stride, size_t cnt, cudaStream_t);
bool comm_film_all_gather(Comm*, float* acc, size_t plane_stride, size_t cnt, cudaStream_t);
bool comm_gather_bytes(Comm*, const unsigned char* send, unsigned char* recv_all, size_t bytes, cudaStream_t);
bool comm_all_reduce_u64_sum(Comm*, unsigned long long* dev, size_t n, cudaStream_t);
bool comm_max_double(Comm*, double* v);
void trim_caches();

struct DeviceScene;     // triangles, materials, LBVH
struct DeviceRenderer;  // per-pixel state, queues, film

struct LbvhDump {
    std::vector<uint32_t> codes, sorted_idx;
    std::vector<int32_t> left, right, parent;
    std::vector<float> node_boxes;
    float scene_box[6];
};

DeviceScene* device_scene_create(const std::vector<HostTri>& tris, const std::vector<HostMaterial>& mats, double origin_l1_bound);
void device_scene_destroy(DeviceScene*);
bool device_scene_build_lbvh(DeviceScene*, int repeats, float ms_out[5]);
bool device_scene_download_lbvh(const DeviceScene*, LbvhDump& out);
bool device_scene_trace(const DeviceScene*, uint32_t n, const float* o, const float* d, float* t, int32_t* tri, float* ms, uint64_t* visits);
void set_query_fp_mode(int strict);
bool device_scene_trace_flat(const DeviceScene*, uint32_t n, const float* o, const float* d, float* t, int32_t* tri);
double device_scene_lbvh_ms(const DeviceScene*);
uint32_t device_scene_num_units(const DeviceScene*);

struct RenderConfig {
    srt_camera cam;
    unsigned spp = 1, bounce_limit = 10;
    unsigned chunk_w = 0, chunk_h = 0;   // nominal chunk geometry (seeds depend on it, reference Q15)
    int fp_strict = 0, pipeline = 0, kernel_timing = 0;
    int stratified = 0;      // opt-in stratified pixel sampler (dormant in the reference, rendering.cu:58-64,89-118)
    int block_slots = 0;     // paths in flight per wavefront block (power of two), 0 = automatic
    int block_threads = 0;   // threads per wavefront block, 0 = automatic
    int sched_flags = 0;     // debug: switches single scheduling features off (SRT_OPT_SCHED_FLAGS)
    int pass_log = 0;        // debug: per-pass log of the first blocks (SRT_OPT_PASS_LOG)
    int rounds = 0;          // wavefront launches per chunk, 0 = automatic (SRT_OPT_ROUNDS)
    int l2_persist = 1;      // persisting-L2 window over the in-flight path state (SRT_OPT_L2_PERSIST)
    int traversal = 0;  // 0 auto (wide leaf when <= 64 triangles), 1 force LBVH walk in shared memory, 3 force LBVH walk in global memory
    int tile_w = 0, tile_h = 0, rank = 0, world = 1;  // tile 0x0 = pick automatically
    Comm* comm = nullptr;    // multi-GPU film exchange (borrowed); rank / world then come from it
    float bg_spectrum[SRT_NS];
    int bg_is_zero = 1;
};
DeviceRenderer* device_renderer_create(const DeviceScene*, const RenderConfig&);
void device_renderer_destroy(DeviceRenderer*);
// renders one chunk (offset/size in pixels) into the device film (XYZ sums, full-image raster)
bool device_renderer_render_chunk(DeviceRenderer*, unsigned off_x, unsigned off_y, unsigned w, unsigned h);
// tonemaps film region -> host planes (pinned staging inside); any pointer may be null
bool device_renderer_resolve(DeviceRenderer*, unsigned off_x, unsigned off_y, unsigned w, unsigned h, float* r, float* g,
                             float* b, unsigned img_w, unsigned img_h);
bool device_renderer_download_xyz(DeviceRenderer*, float* xyz);
bool device_renderer_exchange_film(DeviceRenderer*, float* r, float* g, float* b, unsigned img_w, unsigned img_h);
bool device_renderer_film_checksum(DeviceRenderer*, uint64_t* out);
float* device_renderer_film(DeviceRenderer*);
bool device_renderer_reset(DeviceRenderer*);
bool device_renderer_pass_log(DeviceRenderer*, uint32_t* out);
void device_renderer_stats(const DeviceRenderer*, srt_stats* s);

uint64_t kernel_launches();
double measure_fp32_tflops();
double measure_copy_gbs(uint32_t mbytes);
void widen_u8_to_f32(const unsigned char* src, float* dst, size_t n);  // image_io.cpp
double measure_l2_read_gbs();
bool cuda_select_device(int dev);
int cuda_device_count();

// ---------------------------------------------------------------- scene + render manager
struct Scene {
    SceneDescription desc;
    DeviceScene* dev = nullptr;
    bool ok = false;
    std::string msg;
    ~Scene();
};

class RenderManager {  // rendering/render_manager.cuh:37-225
public:
    RenderManager(Scene* scene, const srt_camera& cam, float* r, float* g, float* b);
    ~RenderManager();
    int init_renderer(unsigned bounce_limit, unsigned spp);
    int init_device_params(unsigned cw, unsigned ch);
    bool ready() const { return device_inited_ && i_ < n_iterations_; }
    bool done() const { return done_; }
    int step();
    int update_fb();
    int render_cycle();
    int end_render();
    int set_option(int opt, int value);
    int set_comm(Comm* comm);
    int exchange_film();
    int film_checksum(uint64_t* out);
    int get_xyz(float* xyz);
    int get_pass_log(uint32_t* out) { return dev_ && device_renderer_pass_log(dev_, out) ? SRT_OK : SRT_ERR_STATE; }
    float* device_film();
    int resolve_film();
    int restart();
    int stats(srt_stats* s) const;
    unsigned width() const { return cam_.width; }
    unsigned height() const { return cam_.height; }

private:
    struct Slot {  // render_step_data (render_manager.cuh:9-35) without host staging copies
        unsigned off_x = 0, off_y = 0, w = 0, h = 0;
        bool is_last = false;
        bool full = false;
    };
    Scene* scene_;
    srt_camera cam_;
    float *fb_r_, *fb_g_, *fb_b_;
    bool scene_inited_ = false, renderer_inited_ = false, device_inited_ = false, done_ = true;
    RenderConfig cfg_;
    DeviceRenderer* dev_ = nullptr;
    unsigned i_ = 0, n_iterations_ = 0, x_chunks_ = 0, chunk_w_ = 0, chunk_h_ = 0, off_x_ = 0, off_y_ = 0;
    Slot slots_[2];
    size_t next_write_ = 0, next_read_ = 0;
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread worker_;
    bool worker_started_ = false;
    int worker_rc_ = 0;
    std::string worker_error_;  // the worker's srt_last_error() text (the error string is per thread)
};

bool write_ppm(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h);
bool write_bmp(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h);

}  // namespace srt
