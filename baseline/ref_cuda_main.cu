// BASELINE INFRASTRUCTURE (not product code): driver for the UNMODIFIED reference CUDA renderer
// (PieSil/CUDA-spectral-ray-tracer) compiled for sm_100a from /root/reference by
// baseline/build_ref_cuda.sh.  Plays the role of the reference's main.cpp:74-167 without the
// CImg window: parse the reference's own flags, build scene_manager + render_manager, render with
// step()/update_fb(), and time ONLY the render kernel (renderer::render = launch + sync,
// rendering.cu:244-277) with CUDA events, as BASELINE.md section 2 (B1) prescribes.
//
//   ref_cuda_render [reference flags] [--repeat N] [--dump file.f32]
// prints one JSON line: {"impl":"reference-cuda","w":..,"h":..,"spp":..,"kernel_ms":[..],"samples_per_s":..}
#define private public  // reach render_manager::r / renderer::render for kernel-only timing
#include "scene.cuh"
#include "render_manager.cuh"
#undef private
#include "device_init.cuh"
#include "log_context.h"
#include "params.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

using namespace scene;

int main(int argc, char** argv) {
    int repeat = 3;
    std::string dump;
    std::vector<char*> av;
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "--repeat") && i + 1 < argc) { repeat = atoi(argv[++i]); continue; }
        if (!strcmp(argv[i], "--dump") && i + 1 < argc) { dump = argv[++i]; continue; }
        av.push_back(argv[i]);
    }
    auto pm = param_manager::getInstance();
    pm->parseArgs((int)av.size(), av.data());
    init_device_symbols();
    std::vector<double> times;
    uint W = 0, H = 0;
    const uint spp = pm->getParams().getNSamples();
    for (int rep = 0; rep < repeat; rep++) {
        scene_manager sm;
        result res = sm.getResult();
        if (!res.success) { fprintf(stderr, "%s\n", res.msg.c_str()); return 1; }
        W = sm.img_width(); H = sm.img_height();
        frame_buffer fb((size_t)W * H);
        render_manager rm(sm.getWorld(), sm.getMaterials(), sm.getCamPtr(), &fb);
        rm.init_renderer(pm->getParams().getBounceLimit(), spp);
        rm.init_device_params(pm->getParams().getXcsize(), pm->getParams().getYcsize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        double kernel_ms = 0;
        // same chunk walk as render_manager::step (render_manager.cu:3-66), kernel bracketed by events
        bool more = true;
        while (more) {
            uint endx = rm.chunk_width + rm.offset_x, endy = rm.chunk_height + rm.offset_y;
            uint cw = endx > rm.image_width ? rm.chunk_width - (endx - rm.image_width) : rm.chunk_width;
            uint ch = endy > rm.image_height ? rm.chunk_height - (endy - rm.image_height) : rm.chunk_height;
            cudaEventRecord(e0);
            rm.r.render(cw, ch, rm.offset_x, rm.offset_y);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            kernel_ms += ms;
            // hand the chunk to update_fb exactly like step() does after the kernel
            render_step_data* rd = &rm.render_data_container[rm.next_write_render_data_index];
            rm.next_write_render_data_index = (rm.next_write_render_data_index + 1) % 2;
            rd->empty.acquire();
            rd->chunk_width = cw; rd->chunk_height = ch; rd->starting_offset_x = rm.offset_x; rd->starting_offset_y = rm.offset_y;
            size_t size = rm.threads.x * rm.blocks.x * rm.threads.y * rm.blocks.y * sizeof(float);
            cudaMemcpy(rd->fb_r, rm.r.getDevFBr(), size, cudaMemcpyDeviceToHost);
            cudaMemcpy(rd->fb_g, rm.r.getDevFBg(), size, cudaMemcpyDeviceToHost);
            cudaMemcpy(rd->fb_b, rm.r.getDevFBb(), size, cudaMemcpyDeviceToHost);
            rm.i++;
            more = rm.i != rm.n_iterations;
            rd->is_last = !more;
            rd->full.release();
            rm.offset_x = (rm.i % rm.x_chunks) * rm.chunk_width;
            rm.offset_y = (rm.i / rm.x_chunks) * rm.chunk_height;
            rm.update_fb();
        }
        times.push_back(kernel_ms);
        if (rep == repeat - 1 && !dump.empty()) {
            FILE* f = fopen(dump.c_str(), "wb");
            if (f) {
                fwrite(fb.r, sizeof(float), (size_t)W * H, f);
                fwrite(fb.g, sizeof(float), (size_t)W * H, f);
                fwrite(fb.b, sizeof(float), (size_t)W * H, f);
                fclose(f);
            }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    std::vector<double> sorted = times;
    std::sort(sorted.begin(), sorted.end());
    const double med = sorted[sorted.size() / 2];
    printf("{\"impl\": \"reference-cuda\", \"w\": %u, \"h\": %u, \"spp\": %u, \"kernel_ms\": [", W, H, spp);
    for (size_t k = 0; k < times.size(); k++) printf("%s%.3f", k ? ", " : "", times[k]);
    printf("], \"median_ms\": %.3f, \"samples_per_s\": %.6g}\n", med, (double)W * H * spp / (med * 1e-3));
    return 0;
}
