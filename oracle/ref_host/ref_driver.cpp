// TEST INFRASTRUCTURE ONLY (oracle/): C entry points over the UNMODIFIED-IN-MEANING reference
// (PieSil/CUDA-spectral-ray-tracer), host-compiled through shim/cuda_shim.h.  Built by
// build_ref.sh into oracle/_ref/libsrt_ref.so; used by tests/ (to pin the C restatement in
// oracle/srt_oracle.c and to generate tests/golden/*), by bench.py's reference arm and by
// nothing else.  The product library never links or loads this.
//
// The driver plays the role of the reference's main.cpp:74-167: parse the params, upload
// the constant tables, build scene_manager + render_manager, run step()/update_fb().
#include "scene.cuh"
#include "device_init.cuh"
#include "render_manager.cuh"
#include "log_context.h"
#include "params.h"
#include "../rgb2spec.h"

#include <vector>
#include <string>
#include <memory>
#include <mutex>
#include <new>

using namespace scene;

// ------------------------------------------------------------------ shim globals
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local char array[256 * 1024];
int srt_ref_rng_draws = 0;

// Fresh device-heap memory: `new tri` leaves tri::aa_plane uninitialised for non-axis-aligned
// triangles (primitives/tri.cuh:30, tri.cu:60-77).  Zero-fill every allocation so that the
// host oracle deterministically reads AAPlane::NONE (= 0) there, the documented assumption.
void* operator new(std::size_t n) { void* p = std::calloc(n ? n : 1, 1); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { void* p = std::calloc(n ? n : 1, 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

// ------------------------------------------------------------------ the missing table
// (defined writable in ref_table.cpp, filled on demand; see there)
extern "C" void srt_ref_table_init_scale();
extern "C" void srt_ref_table_fill_for_color(float r, float g, float b);
extern "C" float* srt_ref_table_data();

// ------------------------------------------------------------------ state
struct ref_session {
    std::unique_ptr<scene_manager> sm;
    std::unique_ptr<frame_buffer> fb;
    std::unique_ptr<render_manager> rm;
    std::vector<float> xyz_chunk;  // pre-tonemap XYZ SUM per coalesced index of the running chunk
    std::vector<float> xyz;        // raster, 3 planes, XYZ/spp
    uint spp = 1;
};
static ref_session* g = nullptr;
static bool g_symbols = false;

void srt_ref_xyz_hook(const float* xyz_sum, unsigned int idx) {
    if (!g || g->xyz_chunk.empty()) return;
    g->xyz_chunk[3 * (size_t)idx + 0] = xyz_sum[0];
    g->xyz_chunk[3 * (size_t)idx + 1] = xyz_sum[1];
    g->xyz_chunk[3 * (size_t)idx + 2] = xyz_sum[2];
}

static const float k_scene_colors[][3] = {
    {.65f, .05f, .05f}, {.12f, .45f, .15f}, {.73f, .73f, .73f}, {1.f, 1.f, 1.f},
    {.5f, .5f, .5f},    {.12f, .15f, .45f}, {.7f, .7f, .7f},    {0.f, 0.f, 0.f}};

extern "C" {

void srt_ref_close() {
    delete g;
    g = nullptr;
}

// argv-style open, exactly what the reference binary would be started with
// (io/params.h:236-304), e.g. {"-s","0","-xr","400","-ar","16/9","-ns","8","-bl","10"}.
int srt_ref_open(int argc, const char** argv) {
    srt_ref_close();
    // fresh parameter singleton (io/params.h:226-233 keeps a process-wide instance)
    param_manager::instance.reset();
    std::vector<char*> av;
    std::string prog = "srt_ref";
    av.push_back(prog.data());
    std::vector<std::string> store(argv, argv + argc);
    for (auto& s : store) av.push_back(s.data());
    param_manager::getInstance()->parseArgs((int)av.size(), av.data());

    if (!g_symbols) {
        srt_ref_table_init_scale();
        for (auto& c : k_scene_colors) srt_ref_table_fill_for_color(c[0], c[1], c[2]);
        init_device_symbols();  // utils/device_init.cuh:14-46
        g_symbols = true;
    }
    g = new ref_session();
    g->sm.reset(new scene_manager());
    if (!g->sm->getResult().success) { srt_ref_close(); return 1; }
    uint w = g->sm->img_width(), h = g->sm->img_height();
    g->fb.reset(new frame_buffer((size_t)w * h));
    g->rm.reset(new render_manager(g->sm->getWorld(), g->sm->getMaterials(), g->sm->getCamPtr(), g->fb.get()));
    auto pm = param_manager::getInstance();
    g->spp = pm->getParams().getNSamples();
    g->rm->init_renderer(pm->getParams().getBounceLimit(), pm->getParams().getNSamples());
    g->rm->init_device_params(pm->getParams().getXcsize(), pm->getParams().getYcsize());
    return g->rm->isReadyToRender() ? 0 : 2;
}

// extra colour cells for KATs on arbitrary colours (must be called before srt_ref_open's
// first use to reach the "device" copy the scene kernels read).
void srt_ref_prepare_color(float r, float gg, float b) {
    srt_ref_table_fill_for_color(r, gg, b);
}

int srt_ref_width() { return g ? (int)g->sm->img_width() : 0; }
int srt_ref_height() { return g ? (int)g->sm->img_height() : 0; }
int srt_ref_num_tris() { return g ? *g->sm->h_world_size_ptr : 0; }
int srt_ref_num_materials() { return g ? *g->sm->h_n_materials_ptr : 0; }

// the main.cpp:44-55 single-threaded loop: step() then update_fb() per chunk.
// r,g,b: W*H raster planes with values 0..255; xyz (optional): 3 raster planes XYZ/spp.
int srt_ref_render(float* r, float* gch, float* b, float* xyz) {
    if (!g || !g->rm->isReadyToRender()) return 1;
    render_manager& rm = *g->rm;
    const size_t W = rm.getImWidth(), H = rm.getImHeight();
    const size_t per_chunk = (size_t)rm.threads.x * rm.threads.y * rm.blocks.x * rm.blocks.y;
    g->xyz_chunk.assign(3 * per_chunk, 0.f);
    g->xyz.assign(3 * W * H, 0.f);
    bool has_data = true;
    do {
        has_data = rm.step();
        rm.update_fb();
        // same block-linear -> raster mapping as render_manager::update_fb
        // (rendering/render_manager.cuh:88-133), applied to the hooked XYZ sums
        const uint bs = rm.threads.x * rm.threads.y;
        for (size_t idx = 0; idx < per_chunk; ++idx) {
            uint blk = idx / bs, t = idx % bs;
            uint fx = rm.threads.x * (blk % rm.blocks.x) + t % rm.threads.x;
            uint fy = rm.threads.y * (blk / rm.blocks.x) + t / rm.threads.x;
            if (fx < rm.last_chunk_width && fy < rm.last_chunk_height) {
                size_t p = (size_t)(fy + rm.last_offset_y) * W + (fx + rm.last_offset_x);
                for (int c = 0; c < 3; ++c) g->xyz[c * W * H + p] = g->xyz_chunk[3 * idx + c] / float(g->spp);
            }
        }
    } while (has_data);
    std::memcpy(r, g->fb->r, W * H * sizeof(float));
    std::memcpy(gch, g->fb->g, W * H * sizeof(float));
    std::memcpy(b, g->fb->b, W * H * sizeof(float));
    if (xyz) std::memcpy(xyz, g->xyz.data(), 3 * W * H * sizeof(float));
    return 0;
}

// per triangle: f[22] = v0 v1 v2 normal D bbox(xmin,xmax,ymin,ymax,zmin,zmax); i[3] = clockwise, aa_plane, mat
void srt_ref_get_tris(float* f, int* iv) {
    int n = srt_ref_num_tris();
    for (int t = 0; t < n; ++t) {
        const tri* T = g->sm->dev_world[t];
        float* o = f + 22 * t;
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) o[3 * k + c] = T->v[k][c];
        for (int c = 0; c < 3; ++c) o[9 + c] = T->normal[c];
        o[12] = T->D;
        o[13] = T->bbox.x.min; o[14] = T->bbox.x.max;
        o[15] = T->bbox.y.min; o[16] = T->bbox.y.max;
        o[17] = T->bbox.z.min; o[18] = T->bbox.z.max;
        o[19] = o[20] = o[21] = 0.f;
        iv[3 * t + 0] = T->clockwise ? 1 : 0;
        iv[3 * t + 1] = (int)T->aa_plane;
        iv[3 * t + 2] = (int)T->mat_index;
    }
}

// per material: f[107] = col(3) fuzz power B(3) C(3) spectrum(95) pad; type in iv
void srt_ref_get_materials(float* f, int* iv) {
    int n = srt_ref_num_materials();
    for (int m = 0; m < n; ++m) {
        const material& M = g->sm->dev_mat_list[m];
        float* o = f + 108 * m;
        for (int c = 0; c < 3; ++c) o[c] = M.col[c];
        o[3] = M.reflection_fuzz;
        o[4] = M.emission_power;
        for (int c = 0; c < 3; ++c) { o[5 + c] = M.sellmeier_B[c]; o[8 + c] = M.sellmeier_C[c]; }
        for (int c = 0; c < N_CIE_SAMPLES; ++c) o[11 + c] = M.spectral_distribution[c];
        o[106] = o[107] = 0.f;
        iv[m] = (int)M.material_type;
    }
}

// camera_data (rendering/rendering.cuh:19-37): out[22] = w h du(3) dv(3) p00(3) defocus_angle center(3) disk_u(3) disk_v(3)
void srt_ref_get_camera(float* out) {
    camera* c = g->sm->getCamPtr();
    int k = 0;
    out[k++] = (float)c->getImageWidth();
    out[k++] = (float)c->getImageHeight();
    for (int a = 0; a < 3; ++a) out[k++] = c->getPixelDeltaU()[a];
    for (int a = 0; a < 3; ++a) out[k++] = c->getPixelDeltaV()[a];
    for (int a = 0; a < 3; ++a) out[k++] = c->getPixel00Loc()[a];
    out[k++] = c->getDefocusAngle();
    for (int a = 0; a < 3; ++a) out[k++] = c->getCenter()[a];
    for (int a = 0; a < 3; ++a) out[k++] = c->getDefocusDiskU()[a];
    for (int a = 0; a < 3; ++a) out[k++] = c->getDefocusDiskV()[a];
}

// reference BVH in preorder: leaf -> triangle index, internal node -> -1.  Returns count.
static int tri_index_of(const tri* p) {
    int n = srt_ref_num_tris();
    for (int t = 0; t < n; ++t) if (g->sm->dev_world[t] == p) return t;
    return -2;
}
static void preorder(const bvh_node* nd, int* out, int& n) {
    if (!nd) return;
    out[n++] = nd->is_leaf ? tri_index_of(nd->primitive) : -1;
    preorder(nd->left, out, n);
    preorder(nd->right, out, n);
}
int srt_ref_bvh_preorder(int* out) {
    int n = 0;
    preorder((*g->sm->getWorld())->getRoot(), out, n);
    return n;
}

// ------------------------------------------------------------------ known-answer hooks
void srt_ref_xorwow(unsigned int seed, int n, unsigned int* raw, float* uni) {
    curandState a, b;
    curand_init(seed, 0, 0, &a);
    b = a;
    for (int i = 0; i < n; ++i) { raw[i] = curand(&a); uni[i] = cuda_random_float(&b); }
}

// bvh::hit on the global tree (bvh/bvh.cu:88-166).  out[9] = hit t p(3) n(3) front mat ; returns hit
int srt_ref_bvh_hit(const float* o, const float* d, float* out) {
    ray r(point3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]));
    hit_record rec;
    bool h = (*g->sm->getWorld())->hit(r, 0.0f, FLT_MAX, rec);
    out[0] = h ? 1.f : 0.f;
    if (h) {
        out[1] = rec.t;
        for (int c = 0; c < 3; ++c) { out[2 + c] = rec.p[c]; out[5 + c] = rec.normal[c]; }
        out[8] = rec.front_face ? 1.f : 0.f;
        out[9] = (float)rec.mat_index;
    }
    return h;
}

int srt_ref_tri_hit(int t, const float* o, const float* d, float tmin, float tmax, float* out) {
    ray r(point3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]));
    hit_record rec;
    bool h = g->sm->dev_world[t]->hit(r, tmin, tmax, rec);
    out[0] = h ? 1.f : 0.f;
    if (h) {
        out[1] = rec.t;
        for (int c = 0; c < 3; ++c) { out[2 + c] = rec.p[c]; out[5 + c] = rec.normal[c]; }
        out[8] = rec.front_face ? 1.f : 0.f;
        out[9] = (float)rec.mat_index;
    }
    return h;
}

int srt_ref_aabb_hit(const float* box6, const float* o, const float* d, float tmin, float tmax) {
    aabb bx(numeric_interval(box6[0], box6[1]), numeric_interval(box6[2], box6[3]), numeric_interval(box6[4], box6[5]));
    ray r(point3(o[0], o[1], o[2]), vec3(d[0], d[1], d[2]));
    return bx.hit(r, tmin, tmax) ? 1 : 0;
}

// material::scatter (materials/material.cu:55-100).
// ray_io[21] = orig(3) dir(3) valid wavelengths(7) power(7); rec_in[8] = p(3) n(3) t front;
// rng[6] = d v0..v4 (in/out).  Returns did_scatter.
int srt_ref_scatter(int mat, float* ray_io, const float* rec_in, unsigned int* rng) {
    ray r(point3(ray_io[0], ray_io[1], ray_io[2]), vec3(ray_io[3], ray_io[4], ray_io[5]));
    r.valid_wavelengths = (uint)ray_io[6];
    for (int k = 0; k < 7; ++k) { r.wavelengths[k] = ray_io[7 + k]; r.power_distr[k] = ray_io[14 + k]; }
    hit_record rec;
    rec.p = point3(rec_in[0], rec_in[1], rec_in[2]);
    rec.normal = vec3(rec_in[3], rec_in[4], rec_in[5]);
    rec.t = rec_in[6];
    rec.front_face = rec_in[7] != 0.f;
    rec.mat_index = mat;
    curandState s{};
    s.d = rng[0];
    for (int k = 0; k < 5; ++k) s.v[k] = rng[1 + k];
    bool did = g->sm->dev_mat_list[mat].scatter(r, rec, &s);
    rng[0] = s.d;
    for (int k = 0; k < 5; ++k) rng[1 + k] = s.v[k];
    for (int c = 0; c < 3; ++c) { ray_io[c] = r.orig[c]; ray_io[3 + c] = r.dir[c]; }
    ray_io[6] = (float)r.valid_wavelengths;
    for (int k = 0; k < 7; ++k) { ray_io[7 + k] = r.wavelengths[k]; ray_io[14 + k] = r.power_distr[k]; }
    return did ? 1 : 0;
}

float srt_ref_sellmeier(const float* b, const float* c, float lambda) { return sellmeier_index(b, c, lambda); }
float srt_ref_spectrum_interp(const float* table95, float lambda) { return spectrum_interp(table95, lambda, N_CIE_SAMPLES); }
void srt_ref_spectrum_to_xyz(const float* wl, const float* pw, int nvalid, float* xyz) {
    color c = dev_spectrum_to_XYZ(wl, pw, N_RAY_WAVELENGTHS, nvalid);
    xyz[0] = c[0]; xyz[1] = c[1]; xyz[2] = c[2];
}
void srt_ref_tonemap(const float* xyz_mean, float* rgb255) {
    color c = expand_sRGB(XYZ_to_sRGB(color(xyz_mean[0], xyz_mean[1], xyz_mean[2]), reinterpret_cast<const float*>(dev_d65_XYZ_to_sRGB)));
    rgb255[0] = c[0]; rgb255[1] = c[1]; rgb255[2] = c[2];
}
// renderer::get_ray (rendering/rendering.cu:66-87) for pixel (i,j): out[13] = orig dir wavelengths(7)
void srt_ref_get_ray(unsigned int i, unsigned int j, unsigned int* rng, float* out) {
    camera* c = g->sm->getCamPtr();
    curandState s{};
    s.d = rng[0];
    for (int k = 0; k < 5; ++k) s.v[k] = rng[1 + k];
    ray r = renderer::get_ray(i, j, c->getPixel00Loc(), c->getPixelDeltaU(), c->getPixelDeltaV(), c->getCenter(),
                              c->getDefocusDiskU(), c->getDefocusDiskV(), c->getDefocusAngle(), &s);
    rng[0] = s.d;
    for (int k = 0; k < 5; ++k) rng[1 + k] = s.v[k];
    for (int a = 0; a < 3; ++a) { out[a] = r.orig[a]; out[3 + a] = r.dir[a]; }
    for (int k = 0; k < 7; ++k) out[6 + k] = r.wavelengths[k];
}
// renderer::get_ray_stratified_sample (rendering/rendering.cu:89-118; never called by the reference's own kernel)
void srt_ref_get_ray_stratified(unsigned int i, unsigned int j, unsigned int sx, unsigned int sy, float recip_sqrt_spp, unsigned int* rng, float* out) {
    camera* c = g->sm->getCamPtr();
    curandState s{};
    s.d = rng[0];
    for (int k = 0; k < 5; ++k) s.v[k] = rng[1 + k];
    ray r = renderer::get_ray_stratified_sample(i, j, c->getPixel00Loc(), c->getPixelDeltaU(), c->getPixelDeltaV(), sx, sy, recip_sqrt_spp,
                                                c->getCenter(), c->getDefocusAngle(), c->getDefocusDiskU(), c->getDefocusDiskV(), &s);
    rng[0] = s.d;
    for (int k = 0; k < 5; ++k) rng[1 + k] = s.v[k];
    for (int a = 0; a < 3; ++a) { out[a] = r.orig[a]; out[3 + a] = r.dir[a]; }
    for (int k = 0; k < 7; ++k) out[6 + k] = r.wavelengths[k];
}
// the sRGB -> reflectance table sampling of material::compute_spectral_distr for one colour
void srt_ref_color_spectrum(float r, float gg, float b, int emissive, float power, float* out95) {
    material m = emissive ? material::emissive(color(r, gg, b), power) : material::lambertian(color(r, gg, b));
    m.compute_spectral_distr(srt_ref_table_data());
    std::memcpy(out95, m.spectral_distribution, sizeof(float) * N_CIE_SAMPLES);
}

}  // extern "C"
