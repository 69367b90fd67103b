// Host side of the device renderer: per-slot state and queues in HBM, the wavefront iteration
// loop, film resolve, and the standalone ray-query entry.  Replaces the reference's renderer
// class (rendering/rendering.cuh:39-155, rendering.cu:244-357).
#include "cuda_common.cuh"
#include "trace_params.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace srt {

const SrtNode* device_scene_nodes(const DeviceScene* s);
const SrtTri* device_scene_tris(const DeviceScene* s);
const SrtMaterial* device_scene_mats(const DeviceScene* s);
uint32_t device_scene_ntris(const DeviceScene* s);
uint32_t device_scene_nmats(const DeviceScene* s);
const uint32_t* device_scene_sorted_idx(const DeviceScene* s);

namespace {
const LaunchTable& table(int strict) {
    static LaunchTable fast = fastfp::make_launch_table();
    static LaunchTable strict_t = strictfp_::make_launch_table();
    return strict ? strict_t : fast;
}
int sm_count() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}
constexpr size_t kSmemSceneLimit = 96 * 1024;  // scenes whose nodes+tris+materials fit are staged in shared memory
}  // namespace

struct DeviceRenderer {
    const DeviceScene* scene = nullptr;
    RenderConfig cfg;
    WaveParams P{};
    size_t smem = 0;
    int grid = 0;
    // owned device memory
    float* d_cie = nullptr;
    float* d_bg = nullptr;
    uint32_t* d_queues = nullptr;   // 2 x (regen[nslots] + 3*nslots)
    uint32_t* d_counters = nullptr; // 2 x 4
    unsigned long long* d_rays = nullptr;
    float* d_rgb = nullptr;         // resolve staging (device), 3 planes of max chunk
    float* d_xyz = nullptr;
    float* h_stage = nullptr;       // pinned, 6 planes of max chunk
    uint32_t* h_counters = nullptr; // pinned
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // stats
    uint64_t launches = 0, iterations = 0, samples = 0, rays = 0;
    double render_ms = 0;
    bool slots_inited = false;
};

void device_renderer_destroy(DeviceRenderer* r) {
    if (!r) return;
    cudaFree(r->d_cie); cudaFree(r->d_bg); cudaFree(r->d_queues); cudaFree(r->d_counters); cudaFree(r->d_rays); cudaFree(r->d_rgb); cudaFree(r->d_xyz);
    cudaFree(r->P.R0); cudaFree(r->P.R1); cudaFree(r->P.P0); cudaFree(r->P.P1); cudaFree(r->P.G0); cudaFree(r->P.G1); cudaFree(r->P.sidx); cudaFree(r->P.acc);
    if (r->h_stage) cudaFreeHost(r->h_stage);
    if (r->h_counters) cudaFreeHost(r->h_counters);
    if (r->stream) cudaStreamDestroy(r->stream);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    delete r;
}

static bool renderer_setup(DeviceRenderer* r) {
    const RenderConfig& c = r->cfg;
    WaveParams& P = r->P;
    P.nodes = device_scene_nodes(r->scene);
    P.tris = device_scene_tris(r->scene);
    P.mats = device_scene_mats(r->scene);
    P.n_tris = (int)device_scene_ntris(r->scene);
    P.n_mats = (int)device_scene_nmats(r->scene);
    P.bg_is_zero = c.bg_is_zero;
    P.cam.width = c.cam.width; P.cam.height = c.cam.height;
    memcpy(P.cam.du, &c.cam.pixel_delta_u, 12); memcpy(P.cam.dv, &c.cam.pixel_delta_v, 12); memcpy(P.cam.p00, &c.cam.pixel00_loc, 12);
    P.cam.defocus_angle = c.cam.defocus_angle;
    memcpy(P.cam.center, &c.cam.camera_center, 12); memcpy(P.cam.disk_u, &c.cam.defocus_disk_u, 12); memcpy(P.cam.disk_v, &c.cam.defocus_disk_v, 12);
    P.img_w = c.cam.width; P.img_h = c.cam.height;
    P.slot_w = c.chunk_w;
    P.nslots = c.chunk_w * c.chunk_h;
    P.ref_grid_x = c.chunk_w / 28u + 1u;  // render_manager.cu:93-96
    P.spp = c.spp & 0xFFFFu;              // short_uint kernel parameters (rendering.cu:154, Q14)
    P.bounce_limit = c.bounce_limit & 0xFFFFu;
    P.regen_loop = c.regen_loop < 1 ? 1 : c.regen_loop;
    P.tile_w = (uint32_t)std::max(1, c.tile_w); P.tile_h = (uint32_t)std::max(1, c.tile_h);
    P.tiles_x = (P.img_w + P.tile_w - 1) / P.tile_w;
    P.rank = (uint32_t)c.rank; P.world = (uint32_t)std::max(1, c.world);
    P.plane = (size_t)P.img_w * P.img_h;
    const size_t ns = P.nslots;
    SRT_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    SRT_CUDA(cudaEventCreate(&r->ev0));
    SRT_CUDA(cudaEventCreate(&r->ev1));
    SRT_CUDA(cudaMalloc((void**)&P.R0, ns * sizeof(float4)));
    SRT_CUDA(cudaMalloc((void**)&P.R1, ns * sizeof(float4)));
    SRT_CUDA(cudaMalloc((void**)&P.P0, ns * sizeof(float4)));
    SRT_CUDA(cudaMalloc((void**)&P.P1, ns * sizeof(float4)));
    SRT_CUDA(cudaMalloc((void**)&P.G0, ns * sizeof(uint4)));
    SRT_CUDA(cudaMalloc((void**)&P.G1, ns * sizeof(uint2)));
    SRT_CUDA(cudaMalloc((void**)&P.sidx, ns * sizeof(uint32_t)));
    SRT_CUDA(cudaMalloc((void**)&P.acc, 3 * P.plane * sizeof(float)));
    SRT_CUDA(cudaMemsetAsync(P.acc, 0, 3 * P.plane * sizeof(float), r->stream));
    SRT_CUDA(cudaMalloc((void**)&r->d_queues, 2 * 4 * ns * sizeof(uint32_t)));
    SRT_CUDA(cudaMalloc((void**)&r->d_counters, 8 * sizeof(uint32_t)));
    SRT_CUDA(cudaMalloc((void**)&r->d_rays, sizeof(unsigned long long)));
    SRT_CUDA(cudaMemsetAsync(r->d_rays, 0, sizeof(unsigned long long), r->stream));
    SRT_CUDA(cudaMalloc((void**)&r->d_rgb, 3 * ns * sizeof(float)));
    SRT_CUDA(cudaMalloc((void**)&r->d_xyz, 3 * ns * sizeof(float)));
    SRT_CUDA(cudaMallocHost((void**)&r->h_stage, 6 * ns * sizeof(float)));
    SRT_CUDA(cudaMallocHost((void**)&r->h_counters, 8 * sizeof(uint32_t)));
    std::vector<float> cie(3 * SRT_NS);
    for (int k = 0; k < 3; k++) memcpy(cie.data() + k * SRT_NS, cie_table(k), SRT_NS * sizeof(float));
    SRT_CUDA(cudaMalloc((void**)&r->d_cie, cie.size() * sizeof(float)));
    SRT_CUDA(cudaMalloc((void**)&r->d_bg, SRT_NS * sizeof(float)));
    SRT_CUDA(cudaMemcpy(r->d_cie, cie.data(), cie.size() * sizeof(float), cudaMemcpyHostToDevice));
    SRT_CUDA(cudaMemcpy(r->d_bg, c.bg_spectrum, SRT_NS * sizeof(float), cudaMemcpyHostToDevice));
    P.cie = r->d_cie;
    P.bg = r->d_bg;
    P.ray_counter = r->d_rays;
    const LaunchTable& T = table(c.fp_strict);
    const size_t need = T.smem_bytes(P);
    r->smem = need <= kSmemSceneLimit ? need : 0;
    if (r->smem) SRT_CUDA(T.configure(r->smem));
    // persistent grid: a multiple of the SM count, enough blocks to fill every SM's thread slots
    const int sms = sm_count();
    const int per_sm = r->smem ? std::max(1, std::min(8, (int)(200 * 1024 / (r->smem + 1024)))) : 8;
    r->grid = sms * per_sm;
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    return true;
}

DeviceRenderer* device_renderer_create(const DeviceScene* scene, const RenderConfig& cfg) {
    auto* r = new DeviceRenderer();
    r->scene = scene;
    r->cfg = cfg;
    if (!renderer_setup(r)) { device_renderer_destroy(r); return nullptr; }
    return r;
}

bool device_renderer_render_chunk(DeviceRenderer* r, unsigned off_x, unsigned off_y, unsigned w, unsigned h) {
    WaveParams P = r->P;
    const LaunchTable& T = table(r->cfg.fp_strict);
    cudaStream_t st = r->stream;
    const size_t ns = P.nslots;
    P.off_x = off_x; P.off_y = off_y; P.cw = w; P.ch = h;
    SRT_CUDA(cudaEventRecord(r->ev0, st));
    if (!r->slots_inited) {  // RNG states are seeded once and carried across chunks (rendering.cu:209,232)
        T.init_slots(P, st);
        r->launches++; count_launch();
        r->slots_inited = true;
    }
    uint64_t owned = 0;
    if (r->cfg.pipeline == 1) {
        T.megakernel(P, r->grid, r->smem, st);
        r->launches++; count_launch();
        SRT_CUDA_LAST();
    } else {
        uint32_t* q[2] = {r->d_queues, r->d_queues + 4 * ns};
        uint32_t* cnt[2] = {r->d_counters, r->d_counters + 4};
        SRT_CUDA(cudaMemsetAsync(r->d_counters, 0, 8 * sizeof(uint32_t), st));
        P.qr_out = q[0]; P.qm_out = q[0] + ns; P.cnt_out = cnt[0];
        T.begin_chunk(P, st);
        r->launches++; count_launch();
        int cur = 0;
        // upper bound on iterations: every sample needs <= bounce_limit + 1 passes
        const uint64_t max_iters = (uint64_t)P.spp * (P.bounce_limit + 2ull) + 2;
        const int check_every = 8;
        bool first = true;
        for (uint64_t it = 0; it < max_iters;) {
            for (int k = 0; k < check_every; k++, it++) {
                const int nxt = cur ^ 1;
                P.qr_in = q[cur]; P.qm_in = q[cur] + ns; P.cnt_in = cnt[cur];
                P.qr_out = q[nxt]; P.qm_out = q[nxt] + ns; P.cnt_out = cnt[nxt];
                SRT_CUDA(cudaMemsetAsync(cnt[nxt], 0, 4 * sizeof(uint32_t), st));
                T.generate(P, r->grid, r->smem, st);
                T.shade(P, r->grid, r->smem, st);
                r->launches += 2; count_launch(2);
                r->iterations++;
                cur = nxt;
            }
            SRT_CUDA_LAST();
            SRT_CUDA(cudaMemcpyAsync(r->h_counters, cnt[cur], 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            SRT_CUDA(cudaStreamSynchronize(st));
            (void)first;
            if (r->h_counters[0] + r->h_counters[1] + r->h_counters[2] + r->h_counters[3] == 0) break;
        }
    }
    SRT_CUDA(cudaEventRecord(r->ev1, st));
    SRT_CUDA(cudaEventSynchronize(r->ev1));
    float ms = 0;
    SRT_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
    r->render_ms += ms;
    // owned pixels of this chunk (tile ownership) for the sample count
    for (unsigned y = off_y; y < off_y + h; y += 1) {
        const unsigned ty = y / P.tile_h;
        for (unsigned tx = off_x / P.tile_w; tx <= (off_x + w - 1) / P.tile_w; tx++) {
            if ((tx + ty * P.tiles_x) % P.world != P.rank) continue;
            const unsigned x0 = std::max(off_x, tx * P.tile_w), x1 = std::min(off_x + w, (tx + 1) * P.tile_w);
            owned += x1 - x0;
        }
    }
    r->samples += owned * P.spp;
    unsigned long long rays = 0;
    SRT_CUDA(cudaMemcpy(&rays, r->d_rays, sizeof rays, cudaMemcpyDeviceToHost));
    r->rays = rays;
    return true;
}

bool device_renderer_resolve(DeviceRenderer* r, unsigned off_x, unsigned off_y, unsigned w, unsigned h, float* fr, float* fg, float* fb, float* xyz,
                             unsigned img_w, unsigned img_h) {
    const LaunchTable& T = table(r->cfg.fp_strict);
    const size_t n = (size_t)w * h;
    if (n > (size_t)r->P.nslots) {  // whole-image resolve after a chunked / reduced render: go band by band
        const unsigned rows = std::max(1u, (unsigned)(r->P.nslots / w));
        for (unsigned y = 0; y < h; y += rows)
            if (!device_renderer_resolve(r, off_x, off_y + y, w, std::min(rows, h - y), fr, fg, fb, xyz, img_w, img_h)) return false;
        return true;
    }
    T.resolve(r->P.acc, r->P.plane, r->P.img_w, off_x, off_y, w, h, r->P.spp, r->d_rgb, r->d_xyz, r->stream);
    r->launches++; count_launch();
    SRT_CUDA_LAST();
    SRT_CUDA(cudaMemcpyAsync(r->h_stage, r->d_rgb, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    SRT_CUDA(cudaMemcpyAsync(r->h_stage + 3 * n, r->d_xyz, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    float* dst[3] = {fr, fg, fb};
    const size_t img_plane = (size_t)img_w * img_h;
    for (int c = 0; c < 3; c++)
        for (unsigned y = 0; y < h; y++) {
            if (dst[c]) memcpy(dst[c] + (size_t)(off_y + y) * img_w + off_x, r->h_stage + c * n + (size_t)y * w, w * sizeof(float));
            if (xyz) memcpy(xyz + c * img_plane + (size_t)(off_y + y) * img_w + off_x, r->h_stage + 3 * n + c * n + (size_t)y * w, w * sizeof(float));
        }
    return true;
}

float* device_renderer_film(DeviceRenderer* r) { return r->P.acc; }

void device_renderer_stats(const DeviceRenderer* r, srt_stats* s) {
    s->samples = r->samples;
    s->rays = r->rays;
    s->kernel_launches = r->launches;
    s->wavefront_iterations = r->iterations;
    s->render_ms = r->render_ms;
    s->lbvh_ms = device_scene_lbvh_ms(r->scene);
}

bool device_scene_trace(const DeviceScene* s, uint32_t n, const float* o, const float* d, float* t, int32_t* tri, float* ms) {
    WaveParams P{};
    P.nodes = device_scene_nodes(s); P.tris = device_scene_tris(s); P.mats = device_scene_mats(s); P.n_tris = (int)device_scene_ntris(s);
    float *d_o = nullptr, *d_d = nullptr, *d_t = nullptr;
    int32_t* d_tri = nullptr;
    cudaEvent_t e0, e1;
    SRT_CUDA(cudaEventCreate(&e0));
    SRT_CUDA(cudaEventCreate(&e1));
    SRT_CUDA(cudaMalloc((void**)&d_o, 3ull * n * sizeof(float)));
    SRT_CUDA(cudaMalloc((void**)&d_d, 3ull * n * sizeof(float)));
    SRT_CUDA(cudaMalloc((void**)&d_t, (size_t)n * sizeof(float)));
    SRT_CUDA(cudaMalloc((void**)&d_tri, (size_t)n * sizeof(int32_t)));
    SRT_CUDA(cudaMemcpy(d_o, o, 3ull * n * sizeof(float), cudaMemcpyHostToDevice));
    SRT_CUDA(cudaMemcpy(d_d, d, 3ull * n * sizeof(float), cudaMemcpyHostToDevice));
    const int grid = (int)std::min<uint64_t>((n + SRT_BLOCK - 1) / SRT_BLOCK, (uint64_t)sm_count() * 8);
    const LaunchTable& T = table(1);
    T.trace_rays(P, n, d_o, d_d, device_scene_sorted_idx(s), d_t, d_tri, grid, nullptr);  // warm-up
    SRT_CUDA(cudaEventRecord(e0));
    T.trace_rays(P, n, d_o, d_d, device_scene_sorted_idx(s), d_t, d_tri, grid, nullptr);
    SRT_CUDA(cudaEventRecord(e1));
    count_launch(2);
    SRT_CUDA(cudaEventSynchronize(e1));
    SRT_CUDA_LAST();
    float el = 0;
    SRT_CUDA(cudaEventElapsedTime(&el, e0, e1));
    if (ms) *ms = el;
    SRT_CUDA(cudaMemcpy(t, d_t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    SRT_CUDA(cudaMemcpy(tri, d_tri, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    cudaFree(d_o); cudaFree(d_d); cudaFree(d_t); cudaFree(d_tri);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return true;
}

}  // namespace srt
