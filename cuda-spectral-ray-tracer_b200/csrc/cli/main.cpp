// srt_cli: drop-in for the reference executable (main.cpp:135-167) on top of the C-ABI.
// Same flags (io/params.h:236-304); --no-show is implied (there is no display window), --save
// writes renders/<title>.bmp like main.cpp:113-118, plus a .ppm next to it.
// Extensions the reference does not have (stripped before the reference's parser sees the line):
//   --stratified        stratified pixel sampler (spp must be a square number)
//   --strict-fp         kernels built without FMA contraction (bit-identical to the reference's host build)
//   --mesh <file>       render a .obj / .ply mesh (grey lambertian) with the scene's camera instead of scene <id>
//   --xyz <file>        dump the film as raw float32 XYZ planes (X plane, Y plane, Z plane; mean per pixel)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <string>
#include <vector>
#include <sys/stat.h>
#include "srt.h"

int main(int argc, char** argv) {
    bool stratified = false, strict_fp = false;
    std::string mesh_path, xyz_path;
    std::vector<char*> ref_args;
    for (int i = 0; i < argc; i++) {
        const std::string a(argv[i]);
        if (i > 0 && a == "--stratified") stratified = true;
        else if (i > 0 && a == "--strict-fp") strict_fp = true;
        else if (i > 0 && a == "--mesh" && i + 1 < argc) mesh_path = argv[++i];
        else if (i > 0 && a == "--xyz" && i + 1 < argc) xyz_path = argv[++i];
        else ref_args.push_back(argv[i]);
    }
    argc = (int)ref_args.size();
    argv = ref_args.data();
    srt_params* pm = srt_params_instance();
    srt_params_parse(pm, argc, argv);
    std::printf("Image Title: %s\nScene ID: %u\nX res: %u\nY res: %u\nAR: %g\nX chunk size: %u\nY chunk size: %u\n# samples: %u\n# max bounces: %u\n",
                srt_params_img_title(pm), srt_params_scene_id(pm), srt_params_xres(pm), srt_params_yres(pm), srt_params_ar(pm),
                srt_params_xcsize(pm), srt_params_ycsize(pm), srt_params_nsamples(pm), srt_params_bounce_limit(pm));
    srt_scene* scene = nullptr;
    if (mesh_path.empty()) scene = srt_scene_create(srt_params_scene_id(pm));
    else {
        srt_material_desc grey{};
        grey.type = SRT_MAT_LAMBERTIAN;
        grey.color[0] = grey.color[1] = grey.color[2] = 0.73f;
        grey.fuzz = 1.0f;
        const bool ply = mesh_path.size() > 4 && mesh_path.compare(mesh_path.size() - 4, 4, ".ply") == 0;
        scene = ply ? srt_scene_create_ply(mesh_path.c_str(), &grey, 1) : srt_scene_create_obj(mesh_path.c_str(), &grey, 1);
        if (!scene) { std::fprintf(stderr, "%s\n", srt_last_error()); return 1; }
    }
    const char* msg = nullptr;
    if (!srt_scene_result(scene, &msg)) { std::fprintf(stderr, "%s\n", msg); return 1; }
    std::printf("%s\n", msg);
    srt_camera cam;
    srt_scene_camera(scene, &cam);
    const size_t n = (size_t)cam.width * cam.height;
    std::vector<float> r(n), g(n), b(n);
    srt_render_manager* rm = srt_render_manager_create(scene, &cam, r.data(), g.data(), b.data());
    if (stratified) srt_rm_set_option(rm, SRT_OPT_STRATIFIED, 1);
    if (strict_fp) srt_rm_set_option(rm, SRT_OPT_FP_MODE, 1);
    if (srt_rm_init_renderer(rm, srt_params_bounce_limit(pm), srt_params_nsamples(pm)) != SRT_OK ||
        srt_rm_init_device_params(rm, srt_params_xcsize(pm), srt_params_ycsize(pm)) != SRT_OK) {
        std::fprintf(stderr, "%s\n", srt_last_error());
        return 1;
    }
    std::fprintf(stderr, "Rendering... ");
    const auto t0 = std::chrono::steady_clock::now();
    if (srt_rm_render_all(rm) != SRT_OK) { std::fprintf(stderr, "%s\n", srt_last_error()); return 1; }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    srt_stats st;
    srt_rm_get_stats(rm, &st);
    std::fprintf(stderr, "done, took %g seconds.\n", sec);
    std::printf("total rendering time (seconds): %g\nkernel time (ms): %g\nsamples/s: %g\nrays: %llu\n", sec, st.render_ms,
                st.samples / (st.render_ms * 1e-3), (unsigned long long)st.rays);
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
                         srt_scene_num_tris(scene), srt_scene_num_materials(scene), srt_params_nsamples(pm), srt_params_bounce_limit(pm));
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
    if (!xyz_path.empty()) {
        std::vector<float> xyz(3 * n);
        if (srt_rm_get_xyz(rm, xyz.data()) != SRT_OK) { std::fprintf(stderr, "%s\n", srt_last_error()); return 1; }
        if (FILE* f = std::fopen(xyz_path.c_str(), "wb")) {
            std::fwrite(xyz.data(), sizeof(float), xyz.size(), f);
            std::fclose(f);
        } else { std::fprintf(stderr, "cannot write %s\n", xyz_path.c_str()); return 1; }
    }
    srt_render_manager_destroy(rm);
    srt_scene_destroy(scene);
    return 0;
}
