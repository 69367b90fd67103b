// extern "C" surface of libsrt.so (include/srt.h).  Thin: argument checks + forwarding.
#include "srt_host.hpp"
#include <cstring>
#include <cmath>
#include <algorithm>

using namespace srt;

struct srt_params { Params* p; };
struct srt_camera_builder { CameraBuilder b; };
struct srt_scene { Scene s; };
struct srt_render_manager { RenderManager* rm; };
struct srt_comm { Comm* c; };

static srt_params g_params_handle{nullptr};
static bool g_ref_compat = true;

extern "C" {

const char* srt_last_error(void) { return last_error().c_str(); }
uint64_t srt_kernel_launch_count(void) { return kernel_launches(); }
int srt_device_count(void) { return cuda_device_count(); }
int srt_set_device(int device) { return cuda_select_device(device) ? SRT_OK : SRT_ERR_CUDA; }

srt_params* srt_params_instance(void) { g_params_handle.p = &Params::instance(); return &g_params_handle; }
void srt_params_reset(void) { Params::reset_instance(); g_params_handle.p = nullptr; }
void srt_params_parse(srt_params* h, int argc, char** argv) { if (h && h->p) h->p->parse(argc, argv); }
unsigned srt_params_scene_id(const srt_params* h) { return h->p->scene; }
unsigned srt_params_xres(const srt_params* h) { return h->p->xres; }
unsigned srt_params_yres(const srt_params* h) { return h->p->yres; }
float srt_params_ar(const srt_params* h) { return h->p->ar; }
unsigned srt_params_xcsize(const srt_params* h) { return h->p->get_xcsize(); }
unsigned srt_params_ycsize(const srt_params* h) { return h->p->get_ycsize(); }
unsigned srt_params_nsamples(const srt_params* h) { return h->p->n_samples; }
unsigned srt_params_bounce_limit(const srt_params* h) { return h->p->bounce_limit; }
int srt_params_log_active(const srt_params* h) { return h->p->do_log; }
int srt_params_do_save(const srt_params* h) { return h->p->do_save; }
int srt_params_show_render(const srt_params* h) { return h->p->show_render; }
const char* srt_params_img_title(const srt_params* h) { return h->p->get_title().c_str(); }
const char* srt_params_log_subdir(const srt_params* h) { return h->p->log_subdir.c_str(); }

srt_camera_builder* srt_camera_builder_create(void) { return new srt_camera_builder(); }
void srt_camera_builder_destroy(srt_camera_builder* b) { delete b; }
void srt_camera_builder_set_vfov(srt_camera_builder* b, float v) { b->b.vfov = v; }
void srt_camera_builder_set_lookfrom(srt_camera_builder* b, float x, float y, float z) { b->b.lookfrom = vec3f(x, y, z); }
void srt_camera_builder_set_lookat(srt_camera_builder* b, float x, float y, float z) { b->b.lookat = vec3f(x, y, z); }
void srt_camera_builder_set_vup(srt_camera_builder* b, float x, float y, float z) { b->b.vup = vec3f(x, y, z); }
void srt_camera_builder_set_defocus_angle(srt_camera_builder* b, float v) { b->b.defocus_angle = v; }
void srt_camera_builder_set_focus_dist(srt_camera_builder* b, float v) { b->b.focus_dist = v; }
void srt_camera_builder_set_background(srt_camera_builder* b, float x, float y, float z) { b->b.background = vec3f(x, y, z); }
int srt_camera_builder_get_camera(const srt_camera_builder* b, srt_camera* out) {
    if (!b || !out) { set_error("null argument"); return SRT_ERR_ARG; }
    const Params& p = Params::instance();
    *out = b->b.build(p.xres, p.yres);
    return SRT_OK;
}
int srt_camera_builder_get_camera_res(const srt_camera_builder* b, uint32_t w, uint32_t h, srt_camera* out) {
    if (!b || !out || !w || !h) { set_error("bad argument"); return SRT_ERR_ARG; }
    *out = b->b.build(w, h);
    return SRT_OK;
}

void srt_set_ref_compat(int on) { g_ref_compat = on != 0; }
int srt_glass_coefficients(int which, float b[3], float c[3]) {
    if (!b || !c || !glass_coefficients(which, b, c)) { set_error("srt_glass_coefficients: unknown glass"); return SRT_ERR_ARG; }
    return SRT_OK;
}

static srt_scene* finish_scene(srt_scene* h) {
    Scene& s = h->s;
    // ray origins are the camera (lens) or points inside the scene box: bound |o|_1 for the wide-leaf error budgets
    double radius = 0;
    for (const HostTri& t : s.desc.tris)
        for (float b : t.bbox) radius = std::max(radius, (double)std::fabs(b));
    const vec3f c = s.desc.camera.lookfrom;
    const double bound = std::max(3.0 * radius, 1.25 * (std::fabs(c.x) + std::fabs(c.y) + std::fabs(c.z)) + 16.0);
    s.dev = device_scene_create(s.desc.tris, s.desc.mats, bound);
    s.ok = s.dev != nullptr;
    s.msg = s.ok ? "World created" : last_error();  // scene.cu:427 / the result{false,...} returns of init_world
    return h;
}
srt_scene* srt_scene_create(unsigned scene_id) {
    auto* h = new srt_scene();
    h->s.desc = make_reference_scene(scene_id, g_ref_compat);
    return finish_scene(h);
}
srt_scene* srt_scene_create_soup(uint32_t n, uint64_t seed) {
    auto* h = new srt_scene();
    h->s.desc = make_soup_scene(n, seed);
    return finish_scene(h);
}
srt_scene* srt_scene_create_mesh(const float* verts, const uint32_t* mat_index, uint32_t n, const srt_material_desc* mats, uint32_t n_mats) {
    if ((n && !verts) || (n_mats && !mats)) { set_error("null argument"); return nullptr; }
    auto* h = new srt_scene();
    TriangleSoup g;
    g.tris.reserve(n);
    for (uint32_t i = 0; i < n; i++) {
        const float* v = verts + 9ull * i;
        g.add_tri(vec3f(v[0], v[1], v[2]), vec3f(v[3], v[4], v[5]), vec3f(v[6], v[7], v[8]), mat_index ? mat_index[i] : 0, false);
    }
    h->s.desc.tris = std::move(g.tris);
    for (uint32_t m = 0; m < n_mats; m++) h->s.desc.mats.push_back(HostMaterial::from_desc(mats[m], g_ref_compat));
    h->s.desc.camera = make_reference_scene(1, true).camera;
    for (const HostTri& t : h->s.desc.tris)
        if (t.mat >= n_mats) { h->s.ok = false; h->s.msg = "triangle references a missing material"; set_error(h->s.msg); return h; }
    return finish_scene(h);
}
srt_scene* srt_scene_create_obj(const char* path, const srt_material_desc* mats, uint32_t n_mats) {
    std::vector<float> v;
    if (!path || !load_obj(path, v)) return nullptr;
    return srt_scene_create_mesh(v.data(), nullptr, (uint32_t)(v.size() / 9), mats, n_mats);
}
srt_scene* srt_scene_create_ply(const char* path, const srt_material_desc* mats, uint32_t n_mats) {
    std::vector<float> v;
    if (!path || !load_ply(path, v)) return nullptr;
    return srt_scene_create_mesh(v.data(), nullptr, (uint32_t)(v.size() / 9), mats, n_mats);
}
void srt_scene_destroy(srt_scene* s) { delete s; }
int srt_scene_result(const srt_scene* s, const char** msg) {
    if (!s) { if (msg) *msg = "null scene"; return 0; }
    if (msg) *msg = s->s.msg.c_str();
    return s->s.ok ? 1 : 0;
}
int srt_scene_camera_res(const srt_scene* s, uint32_t w, uint32_t h, srt_camera* out) {
    if (!s || !out || !w || !h) { set_error("bad argument"); return SRT_ERR_ARG; }
    *out = s->s.desc.camera.build(w, h);
    return SRT_OK;
}
int srt_scene_camera(const srt_scene* s, srt_camera* out) {
    const Params& p = Params::instance();
    return srt_scene_camera_res(s, p.xres, p.yres, out);
}
uint32_t srt_scene_num_tris(const srt_scene* s) { return s ? (uint32_t)s->s.desc.tris.size() : 0; }
uint32_t srt_scene_num_units(const srt_scene* s) { return s && s->s.dev ? device_scene_num_units(s->s.dev) : 0; }
uint32_t srt_scene_num_materials(const srt_scene* s) { return s ? (uint32_t)s->s.desc.mats.size() : 0; }
int srt_scene_get_tris(const srt_scene* s, float* f, int32_t* iv) {
    if (!s || !f || !iv) { set_error("null argument"); return SRT_ERR_ARG; }
    size_t t = 0;
    for (const HostTri& T : s->s.desc.tris) {
        float* o = f + 22 * t;
        for (int k = 0; k < 3; k++) { o[3 * k] = T.v[k].x; o[3 * k + 1] = T.v[k].y; o[3 * k + 2] = T.v[k].z; }
        o[9] = T.normal.x; o[10] = T.normal.y; o[11] = T.normal.z; o[12] = T.D;
        for (int k = 0; k < 6; k++) o[13 + k] = T.bbox[k];
        o[19] = o[20] = o[21] = 0.f;
        iv[3 * t] = T.clockwise; iv[3 * t + 1] = T.aa_plane; iv[3 * t + 2] = (int32_t)T.mat;
        t++;
    }
    return SRT_OK;
}
int srt_scene_get_materials(const srt_scene* s, float* f, int32_t* iv) {
    if (!s || !f || !iv) { set_error("null argument"); return SRT_ERR_ARG; }
    size_t m = 0;
    for (const HostMaterial& M : s->s.desc.mats) {
        float* o = f + 108 * m;
        o[0] = M.color.x; o[1] = M.color.y; o[2] = M.color.z; o[3] = M.fuzz; o[4] = M.power;
        for (int k = 0; k < 3; k++) { o[5 + k] = M.B[k]; o[8 + k] = M.C[k]; }
        std::memcpy(o + 11, M.spec, sizeof M.spec);
        o[106] = o[107] = 0.f;
        iv[m++] = (int32_t)M.type;
    }
    return SRT_OK;
}
int srt_scene_get_lbvh(const srt_scene* s, uint32_t* codes, uint32_t* sorted_idx, int32_t* left, int32_t* right, int32_t* parent, float* node_boxes,
                       float* scene_box) {
    if (!s || !s->s.dev) { set_error("scene has no device BVH"); return SRT_ERR_STATE; }
    LbvhDump d;
    if (!device_scene_download_lbvh(s->s.dev, d)) return SRT_ERR_CUDA;
    auto cp = [](auto* dst, const auto& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(codes, d.codes); cp(sorted_idx, d.sorted_idx); cp(left, d.left); cp(right, d.right); cp(parent, d.parent); cp(node_boxes, d.node_boxes);
    if (scene_box) std::memcpy(scene_box, d.scene_box, sizeof d.scene_box);
    return SRT_OK;
}
int srt_scene_rebuild_lbvh(srt_scene* s, int repeats, float ms_out[5]) {
    if (!s || !s->s.dev || repeats < 1 || !ms_out) { set_error("bad argument"); return SRT_ERR_ARG; }
    return device_scene_build_lbvh(s->s.dev, repeats, ms_out) ? SRT_OK : SRT_ERR_CUDA;
}
int srt_scene_trace_rays(const srt_scene* s, uint32_t n, const float* o, const float* d, float* t_out, int32_t* tri_out, float* ms_out) {
    if (!s || !s->s.dev || !o || !d || !t_out || !tri_out) { set_error("bad argument"); return SRT_ERR_ARG; }
    if (n == 0) return SRT_OK;
    return device_scene_trace(s->s.dev, n, o, d, t_out, tri_out, ms_out, nullptr) ? SRT_OK : SRT_ERR_CUDA;
}
void srt_set_query_fp_mode(int strict) { set_query_fp_mode(strict); }
int srt_scene_trace_rays_flat(const srt_scene* s, uint32_t n, const float* o, const float* d, float* t_out, int32_t* tri_out) {
    if (!s || !s->s.dev || !o || !d || !t_out || !tri_out) { set_error("bad argument"); return SRT_ERR_ARG; }
    if (n == 0) return SRT_OK;
    return device_scene_trace_flat(s->s.dev, n, o, d, t_out, tri_out) ? SRT_OK : SRT_ERR_STATE;
}
int srt_scene_trace_rays_counted(const srt_scene* s, uint32_t n, const float* o, const float* d, float* t_out, int32_t* tri_out, float* ms_out,
                                 uint64_t visits_out[2]) {
    if (!s || !s->s.dev || !o || !d || !t_out || !tri_out || !visits_out) { set_error("bad argument"); return SRT_ERR_ARG; }
    visits_out[0] = visits_out[1] = 0;
    if (n == 0) return SRT_OK;
    return device_scene_trace(s->s.dev, n, o, d, t_out, tri_out, ms_out, visits_out) ? SRT_OK : SRT_ERR_CUDA;
}

srt_render_manager* srt_render_manager_create(srt_scene* s, const srt_camera* cam, float* r, float* g, float* b) {
    if (!cam) { set_error("null camera"); return nullptr; }
    auto* h = new srt_render_manager();
    h->rm = new RenderManager(s ? &s->s : nullptr, *cam, r, g, b);
    return h;
}
void srt_render_manager_destroy(srt_render_manager* h) { if (h) { delete h->rm; delete h; } }
int srt_rm_init_renderer(srt_render_manager* h, unsigned bl, unsigned spp) { return h->rm->init_renderer(bl, spp); }
int srt_rm_init_device_params(srt_render_manager* h, unsigned cw, unsigned ch) { return h->rm->init_device_params(cw, ch); }
int srt_rm_is_ready_to_render(const srt_render_manager* h) { return h->rm->ready() ? 1 : 0; }
int srt_rm_is_done(const srt_render_manager* h) { return h->rm->done() ? 1 : 0; }
unsigned srt_rm_im_width(const srt_render_manager* h) { return h->rm->width(); }
unsigned srt_rm_im_height(const srt_render_manager* h) { return h->rm->height(); }
int srt_rm_step(srt_render_manager* h) { return h->rm->step(); }
int srt_rm_update_fb(srt_render_manager* h) { return h->rm->update_fb(); }
int srt_rm_render_cycle(srt_render_manager* h) { return h->rm->render_cycle(); }
int srt_rm_end_render(srt_render_manager* h) { return h->rm->end_render(); }
int srt_rm_render_all(srt_render_manager* h) {
    int rc = h->rm->render_cycle();
    if (rc != SRT_OK) return rc;
    int more;
    do { more = h->rm->update_fb(); } while (more > 0);
    h->rm->end_render();
    return more < 0 ? -more : SRT_OK;
}
int srt_rm_set_option(srt_render_manager* h, int opt, int v) { return h->rm->set_option(opt, v); }
int srt_rm_get_xyz(srt_render_manager* h, float* xyz) { return xyz ? h->rm->get_xyz(xyz) : SRT_ERR_ARG; }
float* srt_rm_device_film(srt_render_manager* h) { return h->rm->device_film(); }
int srt_rm_resolve_film(srt_render_manager* h) { return h->rm->resolve_film(); }
int srt_rm_restart(srt_render_manager* h) { return h->rm->restart(); }
int srt_rm_get_pass_log(srt_render_manager* h, uint32_t* out) { return out ? h->rm->get_pass_log(out) : SRT_ERR_ARG; }
int srt_rm_get_stats(const srt_render_manager* h, srt_stats* out) { return out ? h->rm->stats(out) : SRT_ERR_ARG; }

int srt_comm_get_unique_id(unsigned char* id) { return id && comm_unique_id(id) ? SRT_OK : SRT_ERR_CUDA; }
srt_comm* srt_comm_create(const unsigned char* id, int rank, int world) {
    if (!id) { set_error("null unique id"); return nullptr; }
    Comm* c = comm_create(id, rank, world);
    if (!c) return nullptr;
    auto* h = new srt_comm();
    h->c = c;
    return h;
}
void srt_comm_destroy(srt_comm* h) { if (h) { comm_destroy(h->c); delete h; } }
int srt_comm_rank(const srt_comm* h) { return comm_rank(h->c); }
int srt_comm_world(const srt_comm* h) { return comm_world(h->c); }
int srt_comm_max_double(srt_comm* h, double* v) { return h && v && comm_max_double(h->c, v) ? SRT_OK : SRT_ERR_CUDA; }
int srt_nccl_version(void) { return comm_nccl_version(); }
int srt_rm_set_comm(srt_render_manager* h, srt_comm* c) { return h->rm->set_comm(c ? c->c : nullptr); }
int srt_rm_exchange_film(srt_render_manager* h) { return h->rm->exchange_film(); }
int srt_rm_film_checksum(srt_render_manager* h, uint64_t* out) { return out ? h->rm->film_checksum(out) : SRT_ERR_ARG; }
void srt_trim_caches(void) { trim_caches(); }

double srt_measure_fp32_tflops(void) { return measure_fp32_tflops(); }
double srt_measure_copy_gbs(uint32_t mbytes) { return measure_copy_gbs(mbytes); }
double srt_measure_l2_read_gbs(void) { return measure_l2_read_gbs(); }

int srt_write_ppm(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h) { return write_ppm(path, r, g, b, w, h) ? SRT_OK : SRT_ERR_ARG; }
int srt_write_bmp(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h) { return write_bmp(path, r, g, b, w, h) ? SRT_OK : SRT_ERR_ARG; }

}  // extern "C"
