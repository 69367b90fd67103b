// srt_cli: drop-in for the reference executable (main.cpp:135-167) on top of the C-ABI.
// Same flags (io/params.h:236-304); --no-show is implied (there is no display window), --save
// writes renders/<title>.bmp like main.cpp:113-118, plus a .ppm next to it.
// Extensions the reference does not have (stripped before the reference's parser sees the line):
//   --stratified        stratified pixel sampler (spp must be a square number)
//   --strict-fp         kernels built without FMA contraction (bit-identical to the reference's host build)
//   --physical          dielectrics with the physically meant Sellmeier coefficients (materials/material.cuh:67 fixed);
//   --ref-compat        ... or exactly as the reference ships them, C := B (the default)
//   --gpus <N>          render on N GPUs of this machine: one host thread and one NCCL rank per device, image tiles
//                       interleaved over the ranks, films combined by srt_rm_exchange_film into rank 0's frame buffer
//   --mesh <file>       render a .obj / .ply mesh (grey lambertian) with the scene's camera instead of scene <id>
//   --xyz <file>        dump the film as raw float32 XYZ planes (X plane, Y plane, Z plane; mean per pixel)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include "srt.h"

namespace {
struct Options {
    bool stratified = false, strict_fp = false;
    int gpus = 1;
    std::string mesh_path, xyz_path;
};
srt_scene* make_scene(const Options& o, srt_params* pm) {
    if (o.mesh_path.empty()) return srt_scene_create(srt_params_scene_id(pm));
    srt_material_desc grey{};
    grey.type = SRT_MAT_LAMBERTIAN;
    grey.color[0] = grey.color[1] = grey.color[2] = 0.73f;
    grey.fuzz = 1.0f;
    const bool ply = o.mesh_path.size() > 4 && o.mesh_path.compare(o.mesh_path.size() - 4, 4, ".ply") == 0;
    return ply ? srt_scene_create_ply(o.mesh_path.c_str(), &grey, 1) : srt_scene_create_obj(o.mesh_path.c_str(), &grey, 1);
}
// everything one GPU does; rank 0 owns the caller's frame buffer and reports
struct RankResult { int rc = 0; std::string error; srt_stats st{}; uint32_t ntris = 0, nmats = 0; };
void render_rank(const Options& o, srt_params* pm, int rank, const unsigned char* uid, const srt_camera* cam_in, float* r, float* g, float* b,
                 std::vector<float>* xyz, RankResult* out) {
    auto fail = [&](const char* what) { out->rc = 1; out->error = std::string(what) + ": " + srt_last_error(); };
    if (o.gpus > 1 && srt_set_device(rank) != SRT_OK) return fail("srt_set_device");
    srt_comm* comm = nullptr;
    if (o.gpus > 1 && !(comm = srt_comm_create(uid, rank, o.gpus))) return fail("srt_comm_create");
    srt_scene* scene = make_scene(o, pm);
    const char* msg = nullptr;
    if (!scene || !srt_scene_result(scene, &msg)) { out->rc = 1; out->error = scene ? msg : srt_last_error(); return; }
    if (rank == 0) std::printf("%s\n", msg);
    out->ntris = srt_scene_num_tris(scene); out->nmats = srt_scene_num_materials(scene);
    srt_camera cam = *cam_in;
    srt_render_manager* rm = srt_render_manager_create(scene, &cam, r, g, b);
    if (o.stratified) srt_rm_set_option(rm, SRT_OPT_STRATIFIED, 1);
    if (o.strict_fp) srt_rm_set_option(rm, SRT_OPT_FP_MODE, 1);
    if (comm && srt_rm_set_comm(rm, comm) != SRT_OK) return fail("srt_rm_set_comm");
    if (srt_rm_init_renderer(rm, srt_params_bounce_limit(pm), srt_params_nsamples(pm)) != SRT_OK ||
        srt_rm_init_device_params(rm, srt_params_xcsize(pm), srt_params_ycsize(pm)) != SRT_OK)
        return fail("render manager set-up");
    if (comm) {
        int more;
        while ((more = srt_rm_step(rm)) > 0) {}
        if (more < 0 || srt_rm_exchange_film(rm) != SRT_OK) return fail("multi-GPU render");
    } else if (srt_rm_render_all(rm) != SRT_OK) return fail("render");
    srt_rm_get_stats(rm, &out->st);
    if (xyz) {  // after an exchange the read-out is collective: the other ranks read into a scratch copy
        std::vector<float> scratch;
        if (rank != 0) scratch.resize(xyz->size());
        if (srt_rm_get_xyz(rm, rank == 0 ? xyz->data() : scratch.data()) != SRT_OK) return fail("srt_rm_get_xyz");
    }
    srt_render_manager_destroy(rm);
    srt_scene_destroy(scene);
    srt_comm_destroy(comm);
}
}  // namespace

int main(int argc, char** argv) {
    Options o;
    std::vector<char*> ref_args;
    for (int i = 0; i < argc; i++) {
        const std::string a(argv[i]);
        if (i > 0 && a == "--stratified") o.stratified = true;
        else if (i > 0 && a == "--strict-fp") o.strict_fp = true;
        else if (i > 0 && a == "--physical") srt_set_ref_compat(0);
        else if (i > 0 && a == "--ref-compat") srt_set_ref_compat(1);
        else if (i > 0 && a == "--gpus" && i + 1 < argc) o.gpus = std::max(1, atoi(argv[++i]));
        else if (i > 0 && a == "--mesh" && i + 1 < argc) o.mesh_path = argv[++i];
        else if (i > 0 && a == "--xyz" && i + 1 < argc) o.xyz_path = argv[++i];
        else ref_args.push_back(argv[i]);
    }
    argc = (int)ref_args.size();
    argv = ref_args.data();
    srt_params* pm = srt_params_instance();
    srt_params_parse(pm, argc, argv);
    std::printf("Image Title: %s\nScene ID: %u\nX res: %u\nY res: %u\nAR: %g\nX chunk size: %u\nY chunk size: %u\n# samples: %u\n# max bounces: %u\n",
                srt_params_img_title(pm), srt_params_scene_id(pm), srt_params_xres(pm), srt_params_yres(pm), srt_params_ar(pm),
                srt_params_xcsize(pm), srt_params_ycsize(pm), srt_params_nsamples(pm), srt_params_bounce_limit(pm));
    if (o.gpus > srt_device_count()) { std::fprintf(stderr, "--gpus %d: only %d CUDA devices visible\n", o.gpus, srt_device_count()); return 1; }
    // the camera only depends on the scene id and the resolution: build it once from a host-side look at the scene
    srt_camera cam;
    {
        srt_scene* probe = make_scene(o, pm);
        const char* msg = nullptr;
        if (!probe || !srt_scene_result(probe, &msg)) { std::fprintf(stderr, "%s\n", probe ? msg : srt_last_error()); return 1; }
        srt_scene_camera(probe, &cam);
        srt_scene_destroy(probe);
    }
    const size_t n = (size_t)cam.width * cam.height;
    std::vector<float> r(n), g(n), b(n), xyz;
    if (!o.xyz_path.empty()) xyz.resize(3 * n);
    unsigned char uid[SRT_NCCL_UNIQUE_ID_BYTES] = {0};
    if (o.gpus > 1 && srt_comm_get_unique_id(uid) != SRT_OK) { std::fprintf(stderr, "%s\n", srt_last_error()); return 1; }
    std::vector<RankResult> res(o.gpus);
    std::fprintf(stderr, "Rendering... ");
    const auto t0 = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> ranks;
        for (int k = 1; k < o.gpus; k++)
            ranks.emplace_back(render_rank, std::cref(o), pm, k, uid, &cam, r.data(), g.data(), b.data(), xyz.empty() ? nullptr : &xyz, &res[k]);
        render_rank(o, pm, 0, uid, &cam, r.data(), g.data(), b.data(), xyz.empty() ? nullptr : &xyz, &res[0]);
        for (std::thread& t : ranks) t.join();
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int k = 0; k < o.gpus; k++)
        if (res[k].rc) { std::fprintf(stderr, "rank %d: %s\n", k, res[k].error.c_str()); return 1; }
    srt_stats st = res[0].st;
    for (int k = 1; k < o.gpus; k++) {  // whole-job numbers: samples and rays add up, the kernel time is the slowest rank's
        st.samples += res[k].st.samples; st.rays += res[k].st.rays;
        st.render_ms = std::max(st.render_ms, res[k].st.render_ms);
    }
    const uint32_t ntris = res[0].ntris, nmats = res[0].nmats;
    std::fprintf(stderr, "done, took %g seconds.\n", sec);
    std::printf("total rendering time (seconds): %g\nkernel time (ms): %g\nsamples/s: %g\nrays: %llu\n", sec, st.render_ms,
                st.samples / (st.render_ms * 1e-3), (unsigned long long)st.rays);
    if (o.gpus > 1) std::printf("gpus: %d\nfilm exchange (ms): %g\n", o.gpus, st.exchange_ms + st.film_out_ms);
    if (srt_params_log_active(pm)) {  // same "key: value" lines as the reference's log_context (_log_/log_context.cpp:5-65)
        mkdir("logs", 0755);
        std::string dir = "logs";
        if (srt_params_log_subdir(pm)[0]) { dir += std::string("/") + srt_params_log_subdir(pm); mkdir(dir.c_str(), 0755); }
        std::string title = srt_params_img_title(pm);
        for (char& c : title) c = c == ' ' ? '_' : (char)std::tolower((unsigned char)c);
        const long long ts = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
        const std::string path = dir + "/" + std::to_string(ts) + "_" + title + "_log.txt";
        if (FILE* f = std::fopen(path.c_str(), "w")) {
            std::fprintf(f, "image width: %u\nimage height: %u\nscene type: %s\n# primitives: %u\n# materials: %u\nsamples per pixel: %u\nbounce limit: %u\n",
                         cam.width, cam.height, srt_params_scene_id(pm) == 1 ? "Prism World" : (srt_params_scene_id(pm) == 2 ? "Different Materials" : "Cornell Box"),
                         ntris, nmats, srt_params_nsamples(pm), srt_params_bounce_limit(pm));
            std::fprintf(f, "chunk width: %u\nchunk height: %u\ntotal rendering time (seconds): %g\nkernel time (ms): %g\nsamples per second: %g\nrays: %llu\nlbvh build (ms): %g\n",
                         srt_params_xcsize(pm), srt_params_ycsize(pm), sec, st.render_ms, st.samples / (st.render_ms * 1e-3), (unsigned long long)st.rays, st.lbvh_ms);
            std::fclose(f);
            std::printf("log written to %s\n", path.c_str());
        }
    }
    if (srt_params_do_save(pm)) {
        std::string name = srt_params_img_title(pm);
        for (char& c : name) c = c == ' ' ? '_' : (char)std::tolower((unsigned char)c);  // utils/utility.h:30-39
        mkdir("renders", 0755);
        srt_write_bmp(("renders/" + name + ".bmp").c_str(), r.data(), g.data(), b.data(), cam.width, cam.height);
        srt_write_ppm(("renders/" + name + ".ppm").c_str(), r.data(), g.data(), b.data(), cam.width, cam.height);
    }
    if (!o.xyz_path.empty()) {
        if (FILE* f = std::fopen(o.xyz_path.c_str(), "wb")) {
            std::fwrite(xyz.data(), sizeof(float), xyz.size(), f);
            std::fclose(f);
        } else { std::fprintf(stderr, "cannot write %s\n", o.xyz_path.c_str()); return 1; }
    }
    return 0;
}
