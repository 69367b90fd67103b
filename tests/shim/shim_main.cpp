// The reference's main.cpp:16-72 (render_cycle) and :74-133 (render) re-typed against include/srt_shim.h: the same call
// sequence on the same class names, without the CImg window (display/UI is out of scope) and with a PPM / raw-XYZ write
// in place of save_img.  Test fixture: tests/test_host_cpu.py compiles and links it, tests/test_gpu_parity.py runs it
// and compares the picture with the library's own render of the same arguments.
//   shim_main [reference flags] [--strict-fp] --out file.ppm
#include <cstdio>
#include <cstring>
#include <ctime>
#include <iostream>
#include <string>
#include <vector>
#include "srt_shim.h"

using namespace scene;
using namespace std;

static bool g_strict = false;

bool render_cycle(render_manager& rm, frame_buffer& fb, image_channels& ch, bool multithread) {  // main.cpp:16-72
    bool completed = false;
    if (rm.isReadyToRender()) {
        clog << "Rendering... ";
        clock_t start = clock();
        if (multithread) {
            rm.render_cycle();             // main.cpp:29
            bool has_data;
            do {                           // main.cpp:31-41
                has_data = rm.update_fb();
                ch = fb;
            } while (has_data);
        } else {
            bool more;
            do {                           // main.cpp:45-55
                more = rm.step();
                rm.update_fb();
                ch = fb;
            } while (more);
        }
        clock_t stop = clock();
        clog << "done, took " << ((double)(stop - start)) / CLOCKS_PER_SEC << " seconds.\n";
        rm.end_render();                   // main.cpp:64
        completed = true;
    } else {
        cerr << "Device parameters not yet initialized";
    }
    return completed;
}

int render(bool multithread, const string& out) {  // main.cpp:74-133
    scene_manager sm;
    result res = sm.getResult();
    if (!res.success) { cerr << res.msg << endl; return 1; }
    clog << res.msg << endl;
    uint width = sm.img_width();
    uint height = sm.img_height();
    frame_buffer fb((size_t)width * height);
    image_channels ch(fb);
    render_manager rm(sm.getWorld(), sm.getMaterials(), sm.getCamPtr(), &fb);
    auto pm = param_manager::getInstance();
    if (g_strict) srt_rm_set_option(rm.handle(), SRT_OPT_FP_MODE, 1);
    rm.init_renderer(pm->getParams().getBounceLimit(), pm->getParams().getNSamples());
    rm.init_device_params(pm->getParams().getXcsize(), pm->getParams().getYcsize());
    if (!render_cycle(rm, fb, ch, multithread)) return 1;
    if (!out.empty() && srt_write_ppm(out.c_str(), fb.r, fb.g, fb.b, width, height) != SRT_OK) { cerr << "cannot write " << out << endl; return 1; }
    return 0;
}

int main(int argc, char* argv[]) {  // main.cpp:135-167
    string out;
    bool multithread = true;
    vector<char*> args;
    for (int i = 0; i < argc; i++) {
        if (i > 0 && !strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (i > 0 && !strcmp(argv[i], "--strict-fp")) g_strict = true;
        else if (i > 0 && !strcmp(argv[i], "--single-thread")) multithread = false;
        else args.push_back(argv[i]);
    }
    param_manager::getInstance()->parseArgs((int)args.size(), args.data());
    return render(multithread, out);
}
