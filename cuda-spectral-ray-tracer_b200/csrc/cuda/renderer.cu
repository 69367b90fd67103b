// Host side of the device renderer: per-pixel RNG state, the state of the paths in flight and the
// film in HBM; the rounds of k_wavefront launches that render a chunk (with the cost-ordered pixel
// lists between them); film resolve; and the standalone ray-query entry.  Replaces the reference's renderer
// class (rendering/rendering.cuh:39-155, rendering.cu:244-357).
#include "cuda_common.cuh"
#include "trace_params.h"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <cmath>
#include <vector>
#include <mutex>
#include <thread>

namespace srt {

const SrtWide* device_scene_nodes(const DeviceScene* s);
const float4* device_scene_grid(const DeviceScene* s);
const SrtTri* device_scene_tris(const DeviceScene* s);
const SrtFlatUnit* device_scene_flat_units(const DeviceScene* s);
const SrtTri* device_scene_flat_tris(const DeviceScene* s);
uint32_t device_scene_n_units(const DeviceScene* s);
double device_scene_origin_bound(const DeviceScene* s);
void device_scene_bounds(const DeviceScene* s, float lo[3], float hi[3]);
const SrtMaterial* device_scene_mats(const DeviceScene* s);
uint32_t device_scene_ntris(const DeviceScene* s);
uint32_t device_scene_nmats(const DeviceScene* s);
const uint32_t* device_scene_sorted_idx(const DeviceScene* s);
const uint32_t* device_scene_flat_to_orig(const DeviceScene* s);
void device_scene_flat_guard(const DeviceScene* s, float* guard, float* tol);

namespace {
constexpr int kMaxRounds = 8;
constexpr size_t kMinOrderedSlots = 8192;  // chunks with fewer pixel slots are rendered in slot order (the sort would cost more than it saves)
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

// Process-wide caches of large device buffers and of pinned host staging.  cudaMalloc/cudaFree of the ~250 MB of
// per-pixel state cost tens of milliseconds per render manager (more than the render itself at 1080p), page-locking
// the staging buffer costs milliseconds; a caller that renders frame after frame gets the previous frame's blocks back.
// Blocks are tagged with the CUDA device they live on; srt_trim_caches() returns everything unused to the driver.
namespace {
struct PoolBlock { void* ptr; size_t bytes; bool used; int device; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool;      // device memory
std::vector<PoolBlock> g_pinned;    // page-locked host memory (device = -1)
int current_device() { int d = 0; cudaGetDevice(&d); return d; }
void release_unused(std::vector<PoolBlock>& pool, bool pinned) {
    for (size_t i = 0; i < pool.size();) {
        if (!pool[i].used) {
            if (pinned) cudaFreeHost(pool[i].ptr); else cudaFree(pool[i].ptr);
            pool.erase(pool.begin() + i);
        } else i++;
    }
}
}  // namespace
bool device_pool_alloc(void** out, size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(g_pool_mu);
    PoolBlock* best = nullptr;
    for (PoolBlock& b : g_pool)
        if (!b.used && b.device == dev && b.bytes >= bytes && b.bytes <= bytes + bytes / 4 && (!best || b.bytes < best->bytes)) best = &b;
    if (best) { best->used = true; *out = best->ptr; return true; }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {  // out of memory: drop the cache and retry once
        cudaGetLastError();
        release_unused(g_pool, false);
        if (!cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc(pool)", __FILE__, __LINE__)) return false;
    }
    g_pool.push_back({p, bytes, true, dev});
    *out = p;
    return true;
}
void device_pool_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (PoolBlock& b : g_pool)
        if (b.ptr == p) { b.used = false; return; }
    cudaFree(p);
}
void* pinned_pool_alloc(size_t bytes) {
    bytes = (bytes + 4095) & ~(size_t)4095;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    PoolBlock* best = nullptr;
    for (PoolBlock& b : g_pinned)
        if (!b.used && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
    if (best) { best->used = true; return best->ptr; }
    void* p = nullptr;
    if (!cuda_ok(cudaMallocHost(&p, bytes), "cudaMallocHost(staging)", __FILE__, __LINE__)) return nullptr;
    g_pinned.push_back({p, bytes, true, -1});
    return p;
}
void pinned_pool_free(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (PoolBlock& b : g_pinned)
        if (b.ptr == p) { b.used = false; return; }
    cudaFreeHost(p);
}
void trim_caches() {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    release_unused(g_pool, false);
    release_unused(g_pinned, true);
}

static std::atomic<int> g_l2_windows{0};

struct DeviceRenderer {
    const DeviceScene* scene = nullptr;
    RenderConfig cfg;
    WaveParams P{};
    size_t smem = 0;
    int grid = 0, wave_grid = 0, mode = 0;  // 0 global LBVH, 1 shared-memory LBVH, 2 shared-memory wide leaf
    // owned device memory
    float* d_cie = nullptr;
    float* d_bg = nullptr;
    uint32_t* d_tiles = nullptr;
    std::vector<uint32_t> h_tiles;  // chunk-local tiles this rank owns
    unsigned long long* d_rays = nullptr;
    uint32_t* d_cost = nullptr;          // per pixel slot: passes spent so far in the current chunk (rounds > 1)
    PixelOrder* order = nullptr;         // sort scratch + the ordered slot list of the current round
    uint4* d_pass_log = nullptr;
    unsigned long long* d_drain = nullptr;  // kMaxRounds x {first slot retired, last block exit} globaltimer stamps
    std::vector<uint32_t> round_end;     // sample index where each round of a chunk ends (last = spp)
    bool l2_window = false;         // this renderer holds a persisting-L2 window (set-aside released with the last one)
    void* d_state = nullptr;        // in-flight path records (R0 R1 P0 P1 L0 L1 carved out of one block)
    size_t state_bytes = 0;
    unsigned char* d_rgb = nullptr; // resolve staging (device): 3 byte planes of the largest region resolved so far
    size_t rgb_cap = 0;
    unsigned char* h_stage = nullptr;  // pinned (owned by this renderer): 3 byte planes of a region, or 3 float planes of the whole film (get_xyz)
    size_t stage_cap = 0;
    std::mutex stage_mu;            // resolve / get_xyz may come from the consumer thread while another caller reads the film
    int device = 0;                 // CUDA device everything above lives on; every entry point binds its calling thread to it
    // multi-GPU film exchange (cfg.comm): slice of the raster this rank tonemaps after the reduce-scatter
    size_t slice_cnt = 0;
    unsigned char* d_gather = nullptr;   // rank 0: world x 3 x slice_cnt bytes
    unsigned long long* d_sum = nullptr; // film checksum accumulator
    bool exchanged = false;
    cudaEvent_t ex0 = nullptr, ex1 = nullptr, ex2 = nullptr;
    double exchange_ms = 0, film_out_ms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // stats
    std::atomic<uint64_t> launches{0};
    uint64_t iterations = 0, samples = 0, rays = 0;
    double render_ms = 0, drain_ms = 0;
    bool slots_inited = false;
    size_t chunk_px = 0;
    // optional per-kernel timing: event pairs tagged with a category (0 pixel order, 1 k_wavefront, 2 k_megakernel, 3 other)
    std::vector<cudaEvent_t> ev_pool;
    std::vector<cudaEvent_t> band_ev;  // one per row band of a pipelined whole-frame resolve
    std::vector<int> ev_tag;
    size_t ev_used = 0;
    double cat_ms[4] = {0, 0, 0, 0};
    uint64_t cat_launches[4] = {0, 0, 0, 0};
};

namespace {
struct KernelTimer {  // brackets one launch with two events when per-kernel timing is on
    DeviceRenderer* r;
    int tag;
    KernelTimer(DeviceRenderer* rr, int t) : r(rr), tag(t) {
        r->cat_launches[tag]++;
        if (!r->cfg.kernel_timing) return;
        if (r->ev_used + 2 > r->ev_pool.size()) {
            for (int k = 0; k < 2; k++) { cudaEvent_t e; cudaEventCreate(&e); r->ev_pool.push_back(e); }
            r->ev_tag.push_back(tag);
        } else r->ev_tag[r->ev_used / 2] = tag;
        cudaEventRecord(r->ev_pool[r->ev_used], r->stream);
    }
    ~KernelTimer() {
        if (!r->cfg.kernel_timing) return;
        cudaEventRecord(r->ev_pool[r->ev_used + 1], r->stream);
        r->ev_used += 2;
    }
};
void collect_kernel_times(DeviceRenderer* r) {
    for (size_t k = 0; k + 1 < r->ev_used; k += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r->ev_pool[k], r->ev_pool[k + 1]) == cudaSuccess) r->cat_ms[r->ev_tag[k / 2]] += ms;
    }
    r->ev_used = 0;
}
}  // namespace

void device_renderer_destroy(DeviceRenderer* r) {
    if (!r) return;
    device_pool_free(r->d_cie); device_pool_free(r->d_bg); device_pool_free(r->d_tiles); device_pool_free(r->d_rays); device_pool_free(r->d_rgb);
    cudaSetDevice(r->device);
    device_pool_free(r->d_pass_log); device_pool_free(r->d_gather); device_pool_free(r->d_sum); pinned_pool_free(r->h_stage);
    if (r->ex0) cudaEventDestroy(r->ex0);
    if (r->ex1) cudaEventDestroy(r->ex1);
    if (r->ex2) cudaEventDestroy(r->ex2);
    device_pool_free(r->d_cost); device_pool_free(r->d_drain); pixel_order_destroy(r->order);
    device_pool_free(r->d_state); device_pool_free(r->P.G0); device_pool_free(r->P.G1); device_pool_free(r->P.next_slot); device_pool_free(r->P.acc);
    if (r->l2_window) {  // hand the pinned L2 lines back; the last renderer also returns the set-aside to the normal cache
        cudaCtxResetPersistingL2Cache();
        if (--g_l2_windows == 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        cudaGetLastError();
    }
    if (r->stream) cudaStreamDestroy(r->stream);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    for (cudaEvent_t e : r->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : r->band_ev) cudaEventDestroy(e);
    delete r;
}

// grows the device / pinned staging of the tonemapped bytes to `bytes` (3 planes of a region)
static bool ensure_staging(DeviceRenderer* r, size_t dev_bytes, size_t host_bytes) {
    if (dev_bytes > r->rgb_cap) {
        device_pool_free(r->d_rgb);
        r->d_rgb = nullptr; r->rgb_cap = 0;
        if (!device_pool_alloc((void**)&r->d_rgb, dev_bytes)) return false;
        r->rgb_cap = dev_bytes;
    }
    if (host_bytes > r->stage_cap) {
        pinned_pool_free(r->h_stage);
        r->h_stage = nullptr; r->stage_cap = 0;
        r->h_stage = (unsigned char*)pinned_pool_alloc(host_bytes);
        if (!r->h_stage) return false;
        r->stage_cap = host_bytes;
    }
    return true;
}

// shared memory of a wavefront block ahead of the staged scene: queues [2][4][S] u16 | pixel slot [S] u32 | samples
// started [S] u16 | passes [S] u16 | pixel xy [S] u32
static void set_queue_layout(WaveParams& P) { P.queue_bytes = (28u * P.block_slots + 15u) & ~15u; }

static bool renderer_setup(DeviceRenderer* r) {
    const RenderConfig& c = r->cfg;
    WaveParams& P = r->P;
    PhaseTrace tr("renderer_setup");
    SRT_CUDA(cudaGetDevice(&r->device));
    if (c.comm && comm_device(c.comm) != r->device) { set_error("the communicator was created on another CUDA device"); return false; }
    P.nodes = device_scene_nodes(r->scene);
    P.grid = device_scene_grid(r->scene);
    P.tris = device_scene_tris(r->scene);
    P.flat_units = device_scene_flat_units(r->scene);
    P.flat_tris = device_scene_flat_tris(r->scene);
    P.n_units = (int)device_scene_n_units(r->scene);
    device_scene_flat_guard(r->scene, &P.flat_guard, &P.flat_tol);
    P.mats = device_scene_mats(r->scene);
    P.n_tris = (int)device_scene_ntris(r->scene);
    P.n_mats = (int)device_scene_nmats(r->scene);
    P.bg_is_zero = c.bg_is_zero;
    P.cam.width = c.cam.width; P.cam.height = c.cam.height;
    memcpy(P.cam.du, &c.cam.pixel_delta_u, 12); memcpy(P.cam.dv, &c.cam.pixel_delta_v, 12); memcpy(P.cam.p00, &c.cam.pixel00_loc, 12);
    P.cam.defocus_angle = c.cam.defocus_angle;
    memcpy(P.cam.center, &c.cam.camera_center, 12); memcpy(P.cam.disk_u, &c.cam.defocus_disk_u, 12); memcpy(P.cam.disk_v, &c.cam.defocus_disk_v, 12);
    P.img_w = c.cam.width; P.img_h = c.cam.height;
    P.tile_w = (uint32_t)std::max(1, c.tile_w); P.tile_h = (uint32_t)std::max(1, c.tile_h);
    const uint32_t tile_slots = P.tile_w * P.tile_h;
    if (tile_slots > 65536 || tile_slots < 32 || (P.tile_w & (P.tile_w - 1)) || (P.tile_h & (P.tile_h - 1))) {
        set_error("tile width and height must be powers of two with an area in [32, 65536]");
        return false;
    }
    if (c.chunk_w > 65535u || c.chunk_h > 65535u) {  // a pixel's chunk coordinates travel as y << 16 | x
        set_error("chunks larger than 65535 pixels in one dimension are not supported");
        return false;
    }
    for (P.tile_slots_log2 = 0; (1u << P.tile_slots_log2) < tile_slots; P.tile_slots_log2++) {}
    for (P.tile_w_log2 = 0; (1u << P.tile_w_log2) < P.tile_w; P.tile_w_log2++) {}
    device_scene_bounds(r->scene, P.scene_lo, P.scene_hi);
    P.block_threads = c.block_threads > 0 ? (uint32_t)std::min(256, std::max(32, c.block_threads & ~31)) : 256u;
    P.tiles_x = (c.chunk_w + P.tile_w - 1) / P.tile_w;
    const uint32_t tiles_y = (c.chunk_h + P.tile_h - 1) / P.tile_h;
    P.rank = (uint32_t)c.rank; P.world = (uint32_t)std::max(1, c.world);
    // owner(tile) = (tx + 5 ty) mod world: diagonal interleave, so every rank samples every image
    // column and row (vertical stripes would hand one rank the glass pyramid and another the margin)
    for (uint32_t t = 0; t < P.tiles_x * tiles_y; t++)
        if (((t % P.tiles_x) + 5u * (t / P.tiles_x)) % P.world == P.rank) r->h_tiles.push_back(t);
    P.n_tiles = (uint32_t)r->h_tiles.size();
    P.nslots = P.n_tiles * tile_slots;
    P.ref_grid_x = c.chunk_w / 28u + 1u;  // render_manager.cu:93-96
    P.spp = c.spp & 0xFFFFu;              // short_uint kernel parameters (rendering.cu:154, Q14)
    P.bounce_limit = c.bounce_limit & 0xFFFFu;
    P.strat_n = 0; P.strat_recip = 0.0f;
    if (c.stratified) {
        while ((P.strat_n + 1) * (P.strat_n + 1) <= P.spp) P.strat_n++;
        if (P.strat_n * P.strat_n != P.spp) {
            set_error("stratified sampling needs a square number of samples per pixel");
            return false;
        }
        P.strat_recip = 1.0f / (float)P.strat_n;
    }
    P.s_begin = 0; P.s_end = P.spp; P.order = nullptr; P.n_order = 0; P.cost = nullptr; P.drain_clock = nullptr;
    // film planes: W*H floats each; with a communicator the stride is padded to world equal slices (16-byte aligned)
    // so that the in-place reduce-scatter hands rank r the pixels [r * slice_cnt, (r+1) * slice_cnt) of every plane
    P.plane = (size_t)P.img_w * P.img_h;
    if (c.comm) {
        const size_t world = (size_t)comm_world(c.comm);
        r->slice_cnt = (((P.plane + world - 1) / world) + 3) & ~(size_t)3;
        P.plane = r->slice_cnt * world;
        SRT_CUDA(cudaEventCreate(&r->ex0));
        SRT_CUDA(cudaEventCreate(&r->ex1));
        SRT_CUDA(cudaEventCreate(&r->ex2));
        if (!device_pool_alloc((void**)&r->d_gather, comm_rank(c.comm) == 0 ? world * 3 * r->slice_cnt : 3 * r->slice_cnt)) return false;
    }
    if (!device_pool_alloc((void**)&r->d_sum, sizeof(unsigned long long))) return false;
    const size_t ns = std::max<size_t>(P.nslots, 1);
    P.pass_log = nullptr;
    if (c.pass_log) {
        if (!device_pool_alloc((void**)&r->d_pass_log, (size_t)SRT_PASS_LOG_BLOCKS * SRT_PASS_LOG_PASSES * sizeof(uint4))) return false;
        P.pass_log = r->d_pass_log;
    }
    if (c.kernel_timing) {
        if (!device_pool_alloc((void**)&r->d_drain, 2 * kMaxRounds * sizeof(unsigned long long))) return false;
    }
    if (!device_pool_alloc((void**)&r->d_tiles, std::max<size_t>(1, r->h_tiles.size()) * sizeof(uint32_t))) return false;
    if (!r->h_tiles.empty()) SRT_CUDA(cudaMemcpy(r->d_tiles, r->h_tiles.data(), r->h_tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    P.tiles = r->d_tiles;
    tr.mark("tiles");
    SRT_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    SRT_CUDA(cudaEventCreate(&r->ev0));
    SRT_CUDA(cudaEventCreate(&r->ev1));
    if (!device_pool_alloc((void**)&P.G0, ns * sizeof(uint4))) return false;
    if (!device_pool_alloc((void**)&P.G1, ns * sizeof(uint2))) return false;
    if (!device_pool_alloc((void**)&P.next_slot, sizeof(uint32_t))) return false;
    if (!device_pool_alloc((void**)&P.acc, 3 * P.plane * sizeof(float))) return false;
    SRT_CUDA(cudaMemsetAsync(P.acc, 0, 3 * P.plane * sizeof(float), r->stream));
    if (!device_pool_alloc((void**)&r->d_rays, sizeof(unsigned long long))) return false;
    SRT_CUDA(cudaMemsetAsync(r->d_rays, 0, sizeof(unsigned long long), r->stream));
    const size_t chunk_px = std::min((size_t)c.chunk_w * c.chunk_h, (size_t)P.img_w * P.img_h);
    r->chunk_px = chunk_px;
    if (!ensure_staging(r, 3 * chunk_px, 3 * chunk_px)) return false;
    std::vector<float> cie(3 * SRT_NS);
    for (int k = 0; k < 3; k++) memcpy(cie.data() + k * SRT_NS, cie_table(k), SRT_NS * sizeof(float));
    if (!device_pool_alloc((void**)&r->d_cie, cie.size() * sizeof(float))) return false;
    if (!device_pool_alloc((void**)&r->d_bg, SRT_NS * sizeof(float))) return false;
    SRT_CUDA(cudaMemcpy(r->d_cie, cie.data(), cie.size() * sizeof(float), cudaMemcpyHostToDevice));
    SRT_CUDA(cudaMemcpy(r->d_bg, c.bg_spectrum, SRT_NS * sizeof(float), cudaMemcpyHostToDevice));
    P.cie = r->d_cie;
    P.bg = r->d_bg;
    P.ray_counter = r->d_rays;
    tr.mark("buffers + constants");
    const LaunchTable& T = table(c.fp_strict);
    // wide leaf only when the scene collapsed into <= 32 units AND this camera's origins respect the
    // origin bound the units' error budgets were computed for
    const double cam_l1 = std::fabs(c.cam.camera_center.x) + std::fabs(c.cam.camera_center.y) + std::fabs(c.cam.camera_center.z) +
                          std::fabs(c.cam.defocus_disk_u.x) + std::fabs(c.cam.defocus_disk_u.y) + std::fabs(c.cam.defocus_disk_u.z) +
                          std::fabs(c.cam.defocus_disk_v.x) + std::fabs(c.cam.defocus_disk_v.y) + std::fabs(c.cam.defocus_disk_v.z);
    const bool flat_ok = P.n_units > 0 && cam_l1 <= device_scene_origin_bound(r->scene);
    r->mode = flat_ok && c.traversal == 0 ? 2 : 1;
    size_t need = T.smem_bytes(P, r->mode);
    if (need > kSmemSceneLimit || c.traversal == 3) { r->mode = 0; need = 0; }
    r->smem = need;
    // megakernel grid: a multiple of the SM count, enough blocks to fill every SM's thread slots
    const int sms = sm_count();
    const int per_sm = r->smem ? std::max(1, std::min(8, (int)(200 * 1024 / (r->smem + 1024)))) : 8;
    r->grid = sms * per_sm;
    // wavefront grid: one resident wave of persistent blocks, each with up to 1024 paths in flight
    // (4 per thread); a rank with few pixels takes fewer paths per block so that every SM still gets work
    uint32_t resident = 0;
    auto fit_block_slots = [&]() -> bool {
        for (P.block_slots = 1024;; P.block_slots >>= 1) {
            set_queue_layout(P);
            SRT_CUDA(T.configure(r->smem + P.queue_bytes));
            resident = (uint32_t)sms * (uint32_t)std::max(1, T.wavefront_blocks_per_sm(r->mode, (int)P.min_blocks, (int)P.block_threads, r->smem + P.queue_bytes));
            if (P.block_slots <= P.block_threads || (uint64_t)resident * P.block_slots <= (uint64_t)P.nslots) break;
        }
        return true;
    };
    P.min_blocks = 4;
    if (!fit_block_slots()) return false;
    // a rank whose pixels do not even fill the resident blocks once is bound by the serial chain of its longest pixels, not by
    // throughput: the 80-register build of the kernel (3 blocks per SM) runs a task's dependent chain faster (sched flag 32 / 64 force 4 / 3)
    const bool chain_bound = P.block_slots <= 256 && r->mode == 2;  // measured on the per-rank shares of C2: 1/8 of the frame 6.67 -> 6.14 ms, 1/4 (512 paths per block) 9.88 -> 10.42 ms
    if (((c.sched_flags & 64) || (chain_bound && !(c.sched_flags & 32))) && c.pipeline != 1) {
        P.min_blocks = 3;
        if (!fit_block_slots()) return false;
    }
    if (c.block_slots >= 32 && c.block_slots <= 4096 && !(c.block_slots & (c.block_slots - 1))) {
        P.block_slots = (uint32_t)c.block_slots;
        set_queue_layout(P);
        SRT_CUDA(T.configure(r->smem + P.queue_bytes));
        resident = (uint32_t)sms * (uint32_t)std::max(1, T.wavefront_blocks_per_sm(r->mode, (int)P.min_blocks, (int)P.block_threads, r->smem + P.queue_bytes));
    }
    tr.mark("launch configuration");
    r->wave_grid = (int)std::min<uint64_t>(resident, ((uint64_t)P.nslots + P.block_slots - 1) / P.block_slots);
    // rounds of samples: K launches per chunk that render samples [0,b1) [b1,b2) ... [b_{K-1},spp), each 8x longer than
    // the one before.  The first round hands the pixels out by a first guess of their cost (k_prior_cost), the later
    // ones by the passes they really took.  Automatic choice (measured, tools/drain_probe.py): a second round pays when
    // the per-pixel chains are long (>= 128 spp) or when a path slot renders several pixels in a row.
    {
        int K = c.rounds;
        if (K <= 0) {
            const double generations = (double)P.nslots / std::max<double>(1.0, (double)r->wave_grid * P.block_slots);
            K = (P.spp >= 128 || (P.spp >= 16 && generations >= 2.5)) ? 2 : 1;
        }
        K = std::min(K, kMaxRounds);
        if (c.pipeline == 1) K = 1;
        for (int i = 1; i < K; i++) {
            const uint32_t b = P.spp >> (3 * (K - i));
            if (b > 0 && (r->round_end.empty() || b > r->round_end.back())) r->round_end.push_back(b);
        }
        r->round_end.push_back(P.spp);
    }
    if (c.pipeline != 1 && (r->round_end.size() > 1 || ns >= kMinOrderedSlots)) {
        if (!device_pool_alloc((void**)&r->d_cost, ns * sizeof(uint32_t))) return false;
        r->order = pixel_order_create((uint32_t)ns);
        if (!r->order) return false;
    }
    // state of the paths in flight: ONE allocation (88 B per record, six arrays) so that a single L2 access-policy
    // window can keep it resident -- it is rewritten every pass for the whole launch and never needs to reach DRAM
    const size_t nrec = (std::max<size_t>(1, (size_t)std::max(1, r->wave_grid) * P.block_slots) + 15) & ~(size_t)15;
    r->state_bytes = nrec * 88;
    if (!device_pool_alloc((void**)&r->d_state, r->state_bytes)) return false;
    unsigned char* sb = (unsigned char*)r->d_state;
    P.R0 = (float4*)sb; P.R1 = (float4*)(sb + nrec * 16); P.P0 = (float4*)(sb + nrec * 32); P.P1 = (float4*)(sb + nrec * 48);
    P.L0 = (uint4*)(sb + nrec * 64); P.L1 = (uint2*)(sb + nrec * 80);
    tr.mark("order + path state");
    {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        if (c.l2_persist && max_persist > 0 && max_window > 0) {
            const size_t persist = std::min<size_t>((size_t)max_persist, r->state_bytes);
            cudaStreamAttrValue av{};
            av.accessPolicyWindow.base_ptr = r->d_state;
            av.accessPolicyWindow.num_bytes = std::min<size_t>((size_t)max_window, r->state_bytes);
            av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)av.accessPolicyWindow.num_bytes);
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) != cudaSuccess ||
                cudaStreamSetAttribute(r->stream, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess)
                cudaGetLastError();  // a hint only: rendering does not depend on it
            else { r->l2_window = true; ++g_l2_windows; }
        }
    }
    tr.mark("L2 window");
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    tr.mark("sync");
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
    SRT_CUDA(cudaSetDevice(r->device));  // the CUDA device is per host thread: the render_cycle() worker starts on device 0
    WaveParams P = r->P;
    r->exchanged = false;
    const LaunchTable& T = table(r->cfg.fp_strict);
    cudaStream_t st = r->stream;
    P.off_x = off_x; P.off_y = off_y; P.cw = w; P.ch = h;
    SRT_CUDA(cudaEventRecord(r->ev0, st));
    if (!r->slots_inited && P.n_tiles) {  // RNG states are seeded once and carried across chunks (rendering.cu:209,232)
        KernelTimer kt(r, 3);
        T.init_slots(P, st);
        r->launches++; count_launch();
        r->slots_inited = true;
    }
    // owned pixels of this chunk: the owned tiles clipped to the chunk
    uint64_t owned = 0;
    for (uint32_t tile : r->h_tiles) {
        const uint32_t tx = tile % P.tiles_x, ty = tile / P.tiles_x;
        const uint32_t x0 = tx * P.tile_w, y0 = ty * P.tile_h;
        if (x0 >= w || y0 >= h) continue;
        owned += (uint64_t)(std::min(w, x0 + P.tile_w) - x0) * (std::min(h, y0 + P.tile_h) - y0);
    }
    size_t timed_rounds = 0;
    if (P.n_tiles == 0 || owned == 0) {
        // this rank owns no pixel of the chunk: nothing to launch
    } else if (r->cfg.pipeline == 1) {
        KernelTimer kt(r, 2);
        T.megakernel(P, r->mode, r->grid, r->smem, st);
        r->launches++; count_launch();
    } else {  // one persistent-block launch per round of samples renders the whole chunk
        const size_t K = r->round_end.size();
        P.sched_flags = (uint32_t)r->cfg.sched_flags;
        const uint32_t* order0 = nullptr;
        if (r->order && P.nslots >= kMinOrderedSlots && !(P.sched_flags & 1u)) {  // first round: expensive-looking pixels first (k_prior_cost)
            KernelTimer kt(r, 0);
            P.cost = r->d_cost;
            T.prior_cost(P, st);
            order0 = pixel_order_build(r->order, r->d_cost, P.nslots, 1, st);
            if (!order0) { set_error("pixel order build failed"); return false; }
            r->launches += 4; count_launch();
        }
        if (K > 1) SRT_CUDA(cudaMemsetAsync(r->d_cost, 0, (size_t)P.nslots * sizeof(uint32_t), st));
        if (r->d_drain) SRT_CUDA(cudaMemsetAsync(r->d_drain, 0xFF, 2 * kMaxRounds * sizeof(unsigned long long), st));  // min slots; max slots are zeroed below
        uint32_t s0 = 0;
        for (size_t k = 0; k < K; k++) {
            P.s_begin = s0; P.s_end = r->round_end[k];
            P.cost = K > 1 ? r->d_cost : nullptr;
            P.order = order0; P.n_order = order0 ? (uint32_t)owned : P.nslots;
            if (k > 0) {  // most passes per sample first
                KernelTimer kt(r, 0);
                P.order = pixel_order_build(r->order, r->d_cost, P.nslots, s0, st);
                if (!P.order) { set_error("pixel order build failed"); return false; }
                P.n_order = (uint32_t)owned;
                r->launches += 3;
            }
            if (r->d_drain) {
                P.drain_clock = r->d_drain + 2 * k;
                SRT_CUDA(cudaMemsetAsync(r->d_drain + 2 * k + 1, 0, sizeof(unsigned long long), st));
            }
            SRT_CUDA(cudaMemsetAsync(P.next_slot, 0, sizeof(uint32_t), st));
            if (P.pass_log) SRT_CUDA(cudaMemsetAsync(P.pass_log, 0, (size_t)SRT_PASS_LOG_BLOCKS * SRT_PASS_LOG_PASSES * sizeof(uint4), st));
            {
                KernelTimer kt(r, 1);
                T.wavefront(P, r->mode, r->wave_grid, r->smem + P.queue_bytes, st);
            }
            r->launches++; count_launch();
            r->iterations++;
            s0 = P.s_end;
        }
        timed_rounds = K;
    }
    SRT_CUDA_LAST();
    SRT_CUDA(cudaEventRecord(r->ev1, st));
    SRT_CUDA(cudaEventSynchronize(r->ev1));
    float ms = 0;
    SRT_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
    r->render_ms += ms;
    collect_kernel_times(r);
    if (r->d_drain && timed_rounds) {  // per launch: last block exit - first local slot that found no pixel left
        unsigned long long stamps[2 * kMaxRounds];
        SRT_CUDA(cudaMemcpy(stamps, r->d_drain, sizeof stamps, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < timed_rounds; k++)
            if (stamps[2 * k] != ~0ull && stamps[2 * k + 1] > stamps[2 * k]) r->drain_ms += (double)(stamps[2 * k + 1] - stamps[2 * k]) * 1e-6;
    }
    r->samples += owned * P.spp;
    unsigned long long rays = 0;
    SRT_CUDA(cudaMemcpy(&rays, r->d_rays, sizeof rays, cudaMemcpyDeviceToHost));
    r->rays = rays;
    return true;
}

// frame_buffer holds 0..255 as float (frame_buffer.cuh:6-44): widen `count` bytes per channel into the caller's planes
static void widen_rows(const unsigned char* stage, size_t stage_plane, float* const dst[3], unsigned off_x, unsigned off_y, unsigned w, unsigned h, unsigned img_w,
                       bool one_thread = false) {
    auto widen = [&](unsigned y0, unsigned y1) {
        for (int c = 0; c < 3; c++) {
            if (!dst[c]) continue;
            for (unsigned y = y0; y < y1; y++) {
                const unsigned char* src = stage + c * stage_plane + (size_t)y * w;
                float* o = dst[c] + (size_t)(off_y + y) * img_w + off_x;
                widen_u8_to_f32(src, o, w);
            }
        }
    };
    if (!one_thread && (size_t)w * h >= (1u << 18)) {  // big regions: the 4x wider float planes are written by several host threads, one band of rows each
        const unsigned nt = std::max(1u, std::min({8u, std::thread::hardware_concurrency(), h}));
        std::vector<std::thread> pool;
        for (unsigned k = 1; k < nt; k++) pool.emplace_back(widen, (unsigned)((uint64_t)h * k / nt), (unsigned)((uint64_t)h * (k + 1) / nt));
        widen(0, h / nt);
        for (std::thread& th : pool) th.join();
    } else {
        widen(0, h);
    }
}

bool device_renderer_resolve(DeviceRenderer* r, unsigned off_x, unsigned off_y, unsigned w, unsigned h, float* fr, float* fg, float* fb,
                             unsigned img_w, unsigned img_h) {
    SRT_CUDA(cudaSetDevice(r->device));
    const LaunchTable& T = table(r->cfg.fp_strict);
    const size_t n = (size_t)w * h;
    if (n == 0) return true;
    std::lock_guard<std::mutex> lock(r->stage_mu);
    if (!ensure_staging(r, 3 * n, 3 * n)) return false;  // a whole-image resolve after a chunked / reduced render grows the buffers once
    float* const dst[3] = {fr, fg, fb};
    // Big regions (a whole frame) go out as a pipeline of row bands: tonemap + D2H of band k+1 run while a host thread widens band k
    // into the caller's float planes, so the frame costs one band's copy plus one band's widening after the last tonemap instead
    // of the whole copy followed by the whole widening.
    const unsigned nb = n >= (1u << 18) ? std::max(1u, std::min({8u, std::thread::hardware_concurrency(), h})) : 1u;
    if (nb > 1) {
        while (r->band_ev.size() < nb) {
            cudaEvent_t e = nullptr;
            SRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            r->band_ev.push_back(e);
        }
        auto row0 = [&](unsigned b) { return (unsigned)((uint64_t)h * b / nb); };
        for (unsigned b = 0; b < nb; b++) {  // band b: rows [y0, y1), three byte planes of w * (y1 - y0) at byte 3 * w * y0 of both stagings
            const unsigned y0 = row0(b), y1 = row0(b + 1);
            const size_t at = (size_t)3 * w * y0, bytes = (size_t)3 * w * (y1 - y0);
            T.resolve(r->P.acc, r->P.plane, r->P.img_w, off_x, off_y + y0, w, y1 - y0, r->P.spp, r->d_rgb + at, r->stream);
            r->launches++; count_launch();
            SRT_CUDA_LAST();
            SRT_CUDA(cudaMemcpyAsync(r->h_stage + at, r->d_rgb + at, bytes, cudaMemcpyDeviceToHost, r->stream));
            SRT_CUDA(cudaEventRecord(r->band_ev[b], r->stream));
        }
        std::atomic<int> failed{0};
        auto band = [&](unsigned b) {
            const unsigned y0 = row0(b), y1 = row0(b + 1);
            if (cudaSetDevice(r->device) != cudaSuccess || cudaEventSynchronize(r->band_ev[b]) != cudaSuccess) { failed = 1; return; }
            widen_rows(r->h_stage + (size_t)3 * w * y0, (size_t)w * (y1 - y0), dst, off_x, off_y + y0, w, y1 - y0, img_w, true);
        };
        std::vector<std::thread> pool;  // (a persistent worker pool instead of eight threads per frame: measured, no gain)
        for (unsigned b = 1; b < nb; b++) pool.emplace_back(band, b);
        band(0);
        for (std::thread& th : pool) th.join();
        if (failed) { cudaGetLastError(); set_error("film read-back failed (band event)"); return false; }
        return true;
    }
    T.resolve(r->P.acc, r->P.plane, r->P.img_w, off_x, off_y, w, h, r->P.spp, r->d_rgb, r->stream);
    r->launches++; count_launch();
    SRT_CUDA_LAST();
    SRT_CUDA(cudaMemcpyAsync(r->h_stage, r->d_rgb, 3 * n, cudaMemcpyDeviceToHost, r->stream));
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    widen_rows(r->h_stage, n, dst, off_x, off_y, w, h, img_w);
    return true;
}

// Multi-GPU film exchange (north_star: "per-GPU film buffers combined by an NCCL reduce/gather over NVLink"):
//   1. in-place reduce-scatter (sum) of the three XYZ planes: rank r then owns pixels [r*cnt, (r+1)*cnt) of the raster
//      (every pixel was rendered by exactly one rank and is +0 elsewhere, so the sums are the single-GPU bits);
//   2. every rank tonemaps its slice (1/world of the pixels) to bytes;
//   3. the byte slices are gathered on rank 0 (send/recv), copied to the host and widened into the caller's planes.
// Rank 0 therefore touches 1/world of the floats and W*H*3 bytes; nothing but bytes crosses PCIe.
bool device_renderer_exchange_film(DeviceRenderer* r, float* fr, float* fg, float* fb, unsigned img_w, unsigned img_h) {
    SRT_CUDA(cudaSetDevice(r->device));
    Comm* c = r->cfg.comm;
    if (!c) { set_error("no communicator attached (srt_rm_set_comm)"); return false; }
    const LaunchTable& T = table(r->cfg.fp_strict);
    const size_t world = (size_t)comm_world(c), rank = (size_t)comm_rank(c), cnt = r->slice_cnt, npx = (size_t)r->P.img_w * r->P.img_h;
    const size_t first = rank * cnt;
    const uint32_t count = first < npx ? (uint32_t)std::min(cnt, npx - first) : 0u;
    std::lock_guard<std::mutex> lock(r->stage_mu);
    if (!ensure_staging(r, 3 * cnt, rank == 0 ? world * 3 * cnt : 0)) return false;
    cudaStream_t st = r->stream;
    PhaseTrace tr("exchange_film");
    SRT_CUDA(cudaEventRecord(r->ex0, st));
    if (!comm_film_reduce_scatter(c, r->P.acc, r->P.plane, cnt, st)) return false;
    SRT_CUDA(cudaEventRecord(r->ex1, st));
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("reduce-scatter"); }
    T.resolve_slice(r->P.acc, r->P.plane, first, count, (uint32_t)cnt, r->P.spp, r->d_rgb, st);
    if (count) { r->launches++; count_launch(); }
    SRT_CUDA_LAST();
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("slice tonemap"); }
    if (!comm_gather_bytes(c, r->d_rgb, r->d_gather, 3 * cnt, st)) return false;
    if (tr.on) { cudaStreamSynchronize(st); tr.mark("gather"); }
    // rank 0 copies the gathered bytes to the host slice by slice; the host thread that widens slice k waits for that copy only
    const bool threaded = rank == 0 && npx >= (1u << 18) && world > 1;
    if (rank == 0) {
        while (threaded && r->band_ev.size() < world) {
            cudaEvent_t e = nullptr;
            SRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            r->band_ev.push_back(e);
        }
        for (size_t k = 0; k < world; k++) {
            SRT_CUDA(cudaMemcpyAsync(r->h_stage + k * 3 * cnt, r->d_gather + k * 3 * cnt, 3 * cnt, cudaMemcpyDeviceToHost, st));
            if (threaded) SRT_CUDA(cudaEventRecord(r->band_ev[k], st));
        }
    }
    SRT_CUDA(cudaEventRecord(r->ex2, st));
    if (!threaded) SRT_CUDA(cudaEventSynchronize(r->ex2));
    r->exchanged = true;
    std::atomic<int> failed{0};
    if (rank == 0) {  // slices are runs of the raster: widen them as one-row regions
        float* const planes[3] = {fr, fg, fb};
        // a slice is widened by `parts` host threads (about eight in all, whatever the world size), each waiting for its slice's copy
        const size_t parts = threaded ? std::max<size_t>(1, 8 / world) : 1;
        auto widen = [&](size_t job) {
            const size_t k = job / parts, part = job % parts;
            if (threaded && (cudaSetDevice(r->device) != cudaSuccess || cudaEventSynchronize(r->band_ev[k]) != cudaSuccess)) { failed = 1; return; }
            const size_t f = k * cnt;
            if (f >= npx) return;
            const size_t n = std::min(cnt, npx - f), a = n * part / parts, b = n * (part + 1) / parts;
            for (int ch = 0; ch < 3; ch++) {
                if (!planes[ch]) continue;
                const unsigned char* src = r->h_stage + (k * 3 + ch) * cnt;
                float* o = planes[ch] + f;
                widen_u8_to_f32(src + a, o + a, b - a);
            }
        };
        if (threaded) {
            std::vector<std::thread> pool;
            for (size_t j = 1; j < world * parts; j++) pool.emplace_back(widen, j);
            widen(0);
            for (std::thread& th : pool) th.join();
        } else {
            for (size_t k = 0; k < world; k++) widen(k);
        }
    }
    if (threaded) SRT_CUDA(cudaEventSynchronize(r->ex2));
    if (failed) { cudaGetLastError(); set_error("film read-back failed (slice event)"); return false; }
    float ms = 0;
    SRT_CUDA(cudaEventElapsedTime(&ms, r->ex0, r->ex1));
    r->exchange_ms += ms;
    SRT_CUDA(cudaEventElapsedTime(&ms, r->ex1, r->ex2));
    r->film_out_ms += ms;
    tr.mark("d2h + widen");
    return true;
}

// Order-independent 64-bit checksum of the XYZ-sum film: sum over (plane, pixel) of a mix of the index and the float's
// bits.  After an exchange every rank sums its own slice and one all-reduce makes the value common to all ranks; it
// equals the value a single GPU computes over its whole film exactly when the films agree bit for bit.
__global__ void __launch_bounds__(256) k_film_checksum(const float* __restrict__ acc, size_t plane, size_t npx, size_t first, size_t count,
                                                       unsigned long long* __restrict__ out) {
    unsigned long long sum = 0;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < 3 * count; j += (size_t)gridDim.x * blockDim.x) {
        const size_t c = j / count, i = first + (j - c * count);
        unsigned long long z = ((unsigned long long)(c * npx + i) << 32) | __float_as_uint(acc[c * plane + i]);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;  // SplitMix64 finaliser
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        sum += z ^ (z >> 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(out, sum);
}
bool device_renderer_film_checksum(DeviceRenderer* r, uint64_t* out) {
    SRT_CUDA(cudaSetDevice(r->device));
    Comm* c = r->cfg.comm;
    const size_t npx = (size_t)r->P.img_w * r->P.img_h;
    size_t first = 0, count = npx;
    if (c) {
        if (!r->exchanged) { set_error("film checksum with a communicator attached needs srt_rm_exchange_film first"); return false; }
        first = (size_t)comm_rank(c) * r->slice_cnt;
        count = first < npx ? std::min(r->slice_cnt, npx - first) : 0;
    }
    SRT_CUDA(cudaMemsetAsync(r->d_sum, 0, sizeof(unsigned long long), r->stream));
    if (count) {
        k_film_checksum<<<(unsigned)std::min<size_t>((3 * count + 255) / 256, 148 * 8), 256, 0, r->stream>>>(r->P.acc, r->P.plane, npx, first, count, r->d_sum);
        r->launches++; count_launch();
        SRT_CUDA_LAST();
    }
    if (c && !comm_all_reduce_u64_sum(c, r->d_sum, 1, r->stream)) return false;
    unsigned long long h = 0;
    SRT_CUDA(cudaMemcpyAsync(&h, r->d_sum, sizeof h, cudaMemcpyDeviceToHost, r->stream));
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    *out = h;
    return true;
}

// pre-tonemap film: XYZ sums -> mean, with the same float operation as the kernels ((1/spp) * v).  After an exchange
// the call is collective: the slices are all-gathered first so that every rank reads the whole reduced film.
bool device_renderer_download_xyz(DeviceRenderer* r, float* xyz) {
    SRT_CUDA(cudaSetDevice(r->device));
    const size_t npx = (size_t)r->P.img_w * r->P.img_h;
    std::lock_guard<std::mutex> lock(r->stage_mu);
    if (!ensure_staging(r, 0, 3 * npx * sizeof(float))) return false;
    if (r->cfg.comm && r->exchanged && !comm_film_all_gather(r->cfg.comm, r->P.acc, r->P.plane, r->slice_cnt, r->stream)) return false;
    float* stage = (float*)r->h_stage;
    for (int c = 0; c < 3; c++)
        SRT_CUDA(cudaMemcpyAsync(stage + c * npx, r->P.acc + c * r->P.plane, npx * sizeof(float), cudaMemcpyDeviceToHost, r->stream));
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    const float inv = 1 / (float)r->P.spp;
    for (size_t i = 0; i < 3 * npx; i++) xyz[i] = inv * stage[i];
    return true;
}

float* device_renderer_film(DeviceRenderer* r) { return r->P.acc; }
bool device_renderer_pass_log(DeviceRenderer* r, uint32_t* out) {
    SRT_CUDA(cudaSetDevice(r->device));
    if (!r->d_pass_log) { set_error("SRT_OPT_PASS_LOG was not set"); return false; }
    SRT_CUDA(cudaMemcpy(out, r->d_pass_log, (size_t)SRT_PASS_LOG_BLOCKS * SRT_PASS_LOG_PASSES * sizeof(uint4), cudaMemcpyDeviceToHost));
    return true;
}

bool device_renderer_reset(DeviceRenderer* r) {  // back to the state right after creation: empty film, unseeded RNG slots
    SRT_CUDA(cudaSetDevice(r->device));
    r->exchanged = false; r->exchange_ms = 0; r->film_out_ms = 0;
    SRT_CUDA(cudaMemsetAsync(r->P.acc, 0, 3 * r->P.plane * sizeof(float), r->stream));
    SRT_CUDA(cudaMemsetAsync(r->d_rays, 0, sizeof(unsigned long long), r->stream));
    SRT_CUDA(cudaStreamSynchronize(r->stream));
    r->slots_inited = false;
    r->samples = r->rays = 0;
    r->render_ms = 0; r->drain_ms = 0;
    for (int k = 0; k < 4; k++) { r->cat_ms[k] = 0; r->cat_launches[k] = 0; }
    return true;
}

void device_renderer_stats(const DeviceRenderer* r, srt_stats* s) {
    s->samples = r->samples;
    s->rays = r->rays;
    s->kernel_launches = r->launches;
    s->wavefront_launches = r->iterations;
    s->rounds = r->round_end.size();
    s->drain_ms = r->drain_ms;
    s->exchange_ms = r->exchange_ms;
    s->film_out_ms = r->film_out_ms;
    s->render_ms = r->render_ms;
    s->lbvh_ms = device_scene_lbvh_ms(r->scene);
    s->order_ms = r->cat_ms[0]; s->wavefront_ms = r->cat_ms[1]; s->megakernel_ms = r->cat_ms[2]; s->other_ms = r->cat_ms[3];
}

static std::atomic<int> g_query_strict{1};
void set_query_fp_mode(int strict) { g_query_strict = strict ? 1 : 0; }

namespace {
struct QueryBuffers {  // device buffers and events of one closest-hit query call, released on every exit path
    float *o = nullptr, *d = nullptr, *t = nullptr;
    int32_t* tri = nullptr;
    uint32_t* next = nullptr;
    unsigned long long* cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool init(uint32_t n, const float* ho, const float* hd) {
        SRT_CUDA(cudaEventCreate(&e0));
        SRT_CUDA(cudaEventCreate(&e1));
        if (!device_pool_alloc((void**)&o, 3ull * n * sizeof(float)) || !device_pool_alloc((void**)&d, 3ull * n * sizeof(float)) ||
            !device_pool_alloc((void**)&t, (size_t)n * sizeof(float)) || !device_pool_alloc((void**)&tri, (size_t)n * sizeof(int32_t)) ||
            !device_pool_alloc((void**)&next, sizeof(uint32_t)) || !device_pool_alloc((void**)&cnt, 2 * sizeof(unsigned long long)))
            return false;
        SRT_CUDA(cudaMemcpy(o, ho, 3ull * n * sizeof(float), cudaMemcpyHostToDevice));
        SRT_CUDA(cudaMemcpy(d, hd, 3ull * n * sizeof(float), cudaMemcpyHostToDevice));
        return true;
    }
    bool read_back(uint32_t n, float* ht, int32_t* htri) {
        SRT_CUDA(cudaMemcpy(ht, t, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
        SRT_CUDA(cudaMemcpy(htri, tri, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        return true;
    }
    ~QueryBuffers() {
        device_pool_free(o); device_pool_free(d); device_pool_free(t); device_pool_free(tri); device_pool_free(next); device_pool_free(cnt);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};
}  // namespace

bool device_scene_trace(const DeviceScene* s, uint32_t n, const float* o, const float* d, float* t, int32_t* tri, float* ms, uint64_t* visits) {
    WaveParams P{};
    P.nodes = device_scene_nodes(s); P.grid = device_scene_grid(s); P.tris = device_scene_tris(s); P.mats = device_scene_mats(s); P.n_tris = (int)device_scene_ntris(s);
    QueryBuffers q;
    if (!q.init(n, o, d)) return false;
    // persistent warps that refill themselves from a ray counter: one resident wave is the whole grid
    const int grid = (int)std::min<uint64_t>((n + SRT_BLOCK - 1) / SRT_BLOCK, (uint64_t)sm_count() * SRT_TRACE_MIN_BLOCKS);
    const LaunchTable& T = table(g_query_strict);
    T.trace_rays(P, n, q.o, q.d, device_scene_sorted_idx(s), q.t, q.tri, nullptr, q.next, grid, nullptr);  // warm-up
    SRT_CUDA(cudaEventRecord(q.e0));
    T.trace_rays(P, n, q.o, q.d, device_scene_sorted_idx(s), q.t, q.tri, nullptr, q.next, grid, nullptr);
    SRT_CUDA(cudaEventRecord(q.e1));
    count_launch(2);
    if (visits) {  // untimed third pass that counts node visits and leaf tests (algorithmic bytes of the walk)
        SRT_CUDA(cudaMemset(q.cnt, 0, 2 * sizeof(unsigned long long)));
        T.trace_rays(P, n, q.o, q.d, device_scene_sorted_idx(s), q.t, q.tri, q.cnt, q.next, grid, nullptr);
        count_launch();
        unsigned long long h[2] = {0, 0};
        SRT_CUDA(cudaMemcpy(h, q.cnt, sizeof h, cudaMemcpyDeviceToHost));
        visits[0] = h[0]; visits[1] = h[1];
    }
    SRT_CUDA(cudaEventSynchronize(q.e1));
    SRT_CUDA_LAST();
    float el = 0;
    SRT_CUDA(cudaEventElapsedTime(&el, q.e0, q.e1));
    if (ms) *ms = el;
    return q.read_back(n, t, tri);
}

// the same queries through the wide-leaf closest hit the renderer uses for scenes of <= 32 units
bool device_scene_trace_flat(const DeviceScene* s, uint32_t n, const float* o, const float* d, float* t, int32_t* tri) {
    if (device_scene_n_units(s) == 0) { set_error("the scene has no wide leaf (more than 64 triangles or 32 units)"); return false; }
    WaveParams P{};
    P.flat_units = device_scene_flat_units(s); P.flat_tris = device_scene_flat_tris(s); P.n_units = (int)device_scene_n_units(s);
    device_scene_flat_guard(s, &P.flat_guard, &P.flat_tol);
    P.mats = device_scene_mats(s); P.n_tris = (int)device_scene_ntris(s); P.n_mats = (int)device_scene_nmats(s);
    QueryBuffers q;
    if (!q.init(n, o, d)) return false;
    float zeros[4 * SRT_NS] = {0};  // load_scene also stages the CIE / background tables; the queries never read them
    float* d_tables = nullptr;
    if (!device_pool_alloc((void**)&d_tables, sizeof zeros)) return false;
    struct Free { float* p; ~Free() { device_pool_free(p); } } free_tables{d_tables};
    SRT_CUDA(cudaMemcpy(d_tables, zeros, sizeof zeros, cudaMemcpyHostToDevice));
    P.cie = d_tables; P.bg = d_tables + 3 * SRT_NS;
    const LaunchTable& T = table(g_query_strict);
    const size_t smem = T.smem_bytes(P, 2);
    SRT_CUDA(T.configure(smem));
    const int grid = (int)std::min<uint64_t>((n + SRT_BLOCK - 1) / SRT_BLOCK, (uint64_t)sm_count() * 4);
    T.trace_rays_flat(P, n, q.o, q.d, device_scene_flat_to_orig(s), q.t, q.tri, grid, smem, nullptr);
    count_launch();
    SRT_CUDA_LAST();
    SRT_CUDA(cudaDeviceSynchronize());
    return q.read_back(n, t, tri);
}

}  // namespace srt
