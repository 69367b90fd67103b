/* srt_shim.h -- source-compatible stand-ins for the four reference classes main.cpp talks to
 * (PieSil/CUDA-spectral-ray-tracer main.cpp:16-72,74-133), implemented on the C-ABI of libsrt.so.
 * A maintainer who keeps main.cpp drops rendering/, bvh/, materials/, spectrum/, color/, scene/ from the
 * build, includes this header instead of their headers and links with -lsrt.  Compiled and run by
 * tests/shim/shim_main.cpp (tests/test_host_cpu.py builds it, tests/test_gpu_parity.py runs it). */
#ifndef SRT_SHIM_H
#define SRT_SHIM_H
#include <cstddef>
#include <string>
#include "srt.h"

typedef unsigned int uint;

struct frame_buffer {                       /* rendering/frame_buffer.cuh:6-44 */
    size_t channel_size;
    float *r, *g, *b;
    explicit frame_buffer(size_t n) : channel_size(n), r(new float[n]()), g(new float[n]()), b(new float[n]()) {}
    ~frame_buffer() { delete[] r; delete[] g; delete[] b; }
    frame_buffer(const frame_buffer&) = delete;
    frame_buffer& operator=(const frame_buffer&) = delete;
};
struct image_channels {                     /* rendering/frame_buffer.cuh:46-71: uchar copies of the three planes */
    size_t n;
    unsigned char *r, *g, *b;
    explicit image_channels(const frame_buffer& fb) : n(fb.channel_size), r(new unsigned char[n]), g(new unsigned char[n]), b(new unsigned char[n]) { *this = fb; }
    ~image_channels() { delete[] r; delete[] g; delete[] b; }
    image_channels& operator=(const frame_buffer& fb) {
        for (size_t i = 0; i < n; i++) { r[i] = (unsigned char)fb.r[i]; g[i] = (unsigned char)fb.g[i]; b[i] = (unsigned char)fb.b[i]; }
        return *this;
    }
};

class parameters {                          /* io/params.h:21-223, getters only */
public:
    uint getSceneId() const { return srt_params_scene_id(srt_params_instance()); }
    uint getXres() const { return srt_params_xres(srt_params_instance()); }
    uint getYres() const { return srt_params_yres(srt_params_instance()); }
    uint getNSamples() const { return srt_params_nsamples(srt_params_instance()); }
    uint getBounceLimit() const { return srt_params_bounce_limit(srt_params_instance()); }
    uint getXcsize() const { return srt_params_xcsize(srt_params_instance()); }
    uint getYcsize() const { return srt_params_ycsize(srt_params_instance()); }
    bool logActive() const { return srt_params_log_active(srt_params_instance()) != 0; }
    bool doSaveImage() const { return srt_params_do_save(srt_params_instance()) != 0; }
    bool showRender() const { return srt_params_show_render(srt_params_instance()) != 0; }
    std::string getImgTitle() const { return srt_params_img_title(srt_params_instance()); }
};
class param_manager {                       /* io/params.h:226-315: the singleton itself lives inside libsrt */
    parameters p;
public:
    static param_manager* getInstance() { static param_manager pm; return &pm; }
    void parseArgs(int argc, char* argv[]) { srt_params_parse(srt_params_instance(), argc, argv); }
    const parameters& getParams() const { return p; }
};

namespace scene {
struct result { bool success; std::string msg; };
class scene_manager {                       /* scene/scene.cuh:103-176 */
    srt_scene* s;
    srt_camera cam;
public:
    scene_manager() : s(srt_scene_create(srt_params_scene_id(srt_params_instance()))), cam() { if (s) srt_scene_camera(s, &cam); }
    ~scene_manager() { srt_scene_destroy(s); }
    scene_manager(const scene_manager&) = delete;
    scene_manager& operator=(const scene_manager&) = delete;
    result getResult() const {
        const char* m = "";
        const bool ok = s && srt_scene_result(s, &m);
        return {ok, s ? m : srt_last_error()};
    }
    uint img_width() const { return cam.width; }
    uint img_height() const { return cam.height; }
    srt_scene* getWorld() { return s; }      /* was bvh** */
    srt_scene* getMaterials() { return s; }  /* was material* */
    srt_camera* getCamPtr() { return &cam; } /* was camera* */
};
}  // namespace scene

class render_manager {                      /* rendering/render_manager.cuh:37-173 */
    srt_render_manager* rm;
public:
    render_manager(srt_scene* world, srt_scene* /*materials*/, srt_camera* cam, frame_buffer* fb)
        : rm(srt_render_manager_create(world, cam, fb->r, fb->g, fb->b)) {}
    ~render_manager() { srt_render_manager_destroy(rm); }
    render_manager(const render_manager&) = delete;
    render_manager& operator=(const render_manager&) = delete;
    void init_renderer(uint bounce_limit, uint samples_per_pixel) { srt_rm_init_renderer(rm, bounce_limit, samples_per_pixel); }
    void init_device_params(uint chunk_width, uint chunk_height) { srt_rm_init_device_params(rm, chunk_width, chunk_height); }
    void init_device_params() { srt_rm_init_device_params(rm, 0, 0); }
    bool isReadyToRender() const { return srt_rm_is_ready_to_render(rm) != 0; }
    bool isDone() const { return srt_rm_is_done(rm) != 0; }
    bool step() { return srt_rm_step(rm) > 0; }
    bool update_fb() { return srt_rm_update_fb(rm) > 0; }
    void render_cycle() { srt_rm_render_cycle(rm); }   /* a worker thread calls step(); the caller's thread calls update_fb() */
    void end_render() { srt_rm_end_render(rm); }
    uint getImWidth() const { return srt_rm_im_width(rm); }
    uint getImHeight() const { return srt_rm_im_height(rm); }
    srt_render_manager* handle() { return rm; }        /* for the options the reference does not have (srt_rm_set_option) */
};

#endif
