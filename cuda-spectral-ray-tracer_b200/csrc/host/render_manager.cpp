// Chunk scheduler and producer/consumer hand-off: the mirror of the reference's
// rendering/render_manager.{cuh,cu}.  step() (producer, usually a worker thread) renders the next
// chunk on the device; update_fb() (consumer, the caller's thread) resolves the oldest finished
// chunk into the caller's frame buffer.  Two slots, like render_step_data[2] in the reference,
// guarded by one mutex + condition variable instead of four binary semaphores.
#include "srt_host.hpp"
#include <cmath>
#include <iostream>

namespace srt {

Scene::~Scene() { device_scene_destroy(dev); }

RenderManager::RenderManager(Scene* scene, const srt_camera& cam, float* r, float* g, float* b)
    : scene_(scene), cam_(cam), fb_r_(r), fb_g_(g), fb_b_(b) {
    // render_manager.cuh:39-52: everything must be non-null for the manager to be usable
    scene_inited_ = scene && scene->ok && scene->dev && r && g && b && cam.width > 0 && cam.height > 0;
}

RenderManager::~RenderManager() {
    end_render();
    device_renderer_destroy(dev_);
}

int RenderManager::init_renderer(unsigned bounce_limit, unsigned spp) {  // render_manager.cu:121-132
    if (!scene_inited_) {
        std::cerr << "Scene not yet initialized" << std::endl;
        set_error("render manager: scene not initialised");
        return SRT_ERR_STATE;
    }
    cfg_.cam = cam_;
    cfg_.spp = spp;
    cfg_.bounce_limit = bounce_limit;
    background_spectrum(vec3f(cam_.background.x, cam_.background.y, cam_.background.z), cfg_.bg_spectrum);  // rendering.cu:324
    cfg_.bg_is_zero = 1;
    for (float v : cfg_.bg_spectrum) if (v != 0.0f) cfg_.bg_is_zero = 0;
    renderer_inited_ = true;
    return SRT_OK;
}

int RenderManager::set_option(int opt, int value) {
    if (device_inited_) { set_error("options must be set before init_device_params"); return SRT_ERR_STATE; }
    switch (opt) {
        case SRT_OPT_FP_MODE: cfg_.fp_strict = value ? 1 : 0; break;
        case SRT_OPT_PIPELINE: cfg_.pipeline = value ? 1 : 0; break;
        case SRT_OPT_TILE_W: cfg_.tile_w = value; break;
        case SRT_OPT_TILE_H: cfg_.tile_h = value; break;
        case SRT_OPT_RANK: cfg_.rank = value; break;
        case SRT_OPT_WORLD: cfg_.world = value; break;
        case SRT_OPT_KERNEL_TIMING: cfg_.kernel_timing = value ? 1 : 0; break;
        case SRT_OPT_SCHED_FLAGS: cfg_.sched_flags = value; break;
        case SRT_OPT_PASS_LOG: cfg_.pass_log = value ? 1 : 0; break;
        case SRT_OPT_ROUNDS: cfg_.rounds = value; break;
        case SRT_OPT_L2_PERSIST: cfg_.l2_persist = value ? 1 : 0; break;
        case SRT_OPT_STRATIFIED: cfg_.stratified = value ? 1 : 0; break;
        case SRT_OPT_TRAVERSAL: cfg_.traversal = value; break;
        case SRT_OPT_BLOCK_SLOTS: cfg_.block_slots = value; break;
        case SRT_OPT_BLOCK_THREADS: cfg_.block_threads = value; break;
        default: set_error("unknown option"); return SRT_ERR_ARG;
    }
    return SRT_OK;
}

int RenderManager::set_comm(Comm* comm) {
    if (device_inited_) { set_error("the communicator must be attached before init_device_params"); return SRT_ERR_STATE; }
    cfg_.comm = comm;
    cfg_.rank = comm ? comm_rank(comm) : 0;
    cfg_.world = comm ? comm_world(comm) : 1;
    return SRT_OK;
}
int RenderManager::exchange_film() {
    end_render();
    if (!dev_) { set_error("render manager: not initialised"); return SRT_ERR_STATE; }
    return device_renderer_exchange_film(dev_, fb_r_, fb_g_, fb_b_, cam_.width, cam_.height) ? SRT_OK : SRT_ERR_CUDA;
}
int RenderManager::film_checksum(uint64_t* out) {
    if (!dev_) { set_error("render manager: not initialised"); return SRT_ERR_STATE; }
    return device_renderer_film_checksum(dev_, out) ? SRT_OK : SRT_ERR_CUDA;
}

int RenderManager::init_device_params(unsigned cw, unsigned ch) {  // render_manager.cu:68-119
    if (!renderer_inited_) {
        std::cerr << "Init renderer before assigning device parameters" << std::endl;
        set_error("render manager: init_renderer must come first");
        return SRT_ERR_STATE;
    }
    if (cw == 0 && ch == 0) { cw = cam_.width; ch = cam_.height; }  // init_device_params() overload
    if (cw == 0) cw = ch;  // io/params.h:53-63 defaults
    if (ch == 0) ch = cw;
    if (cfg_.tile_w <= 0 || cfg_.tile_h <= 0) {
        // auto tile size: a wavefront block walks its tile's bounce chain serially, so a rank wants at
        // least ~3 waves of tiles over its block slots (148 SMs x 4); fewer pixels per rank => smaller tiles.
        // Deterministic in (chunk size, world), so every rank picks the same partition.
        const double per_rank = double(cw) * ch / std::max(1, cfg_.world);
        const double target = per_rank / (148.0 * 4.0 * 3.0);
        if (target >= 2048) { cfg_.tile_w = 32; cfg_.tile_h = 32; }
        else if (target >= 512) { cfg_.tile_w = 32; cfg_.tile_h = 16; }
        else { cfg_.tile_w = 16; cfg_.tile_h = 16; }
    }
    if (cfg_.rank < 0 || cfg_.rank >= cfg_.world) { set_error("bad tile ownership options"); return SRT_ERR_ARG; }
    chunk_w_ = cw;
    chunk_h_ = ch;
    x_chunks_ = (unsigned)std::ceil(float(cam_.width) / float(cw));
    const unsigned y_chunks = (unsigned)std::ceil(float(cam_.height) / float(ch));
    n_iterations_ = x_chunks_ * y_chunks;
    cfg_.chunk_w = cw;
    cfg_.chunk_h = ch;
    device_renderer_destroy(dev_);
    dev_ = device_renderer_create(scene_->dev, cfg_);
    if (!dev_) return SRT_ERR_CUDA;
    i_ = 0;
    off_x_ = off_y_ = 0;
    next_write_ = next_read_ = 0;
    slots_[0] = Slot();
    slots_[1] = Slot();
    device_inited_ = true;
    return SRT_OK;
}

int RenderManager::step() {  // render_manager.cu:3-66
    if (!device_inited_) {
        std::cerr << "Device parameters were not initialized, render aborted" << std::endl;
        set_error("render manager: init_device_params must come first");
        return -SRT_ERR_STATE;
    }
    if (i_ >= n_iterations_) return 0;
    const unsigned end_x = chunk_w_ + off_x_, end_y = chunk_h_ + off_y_;
    const unsigned w = end_x > cam_.width ? chunk_w_ - (end_x - cam_.width) : chunk_w_;
    const unsigned h = end_y > cam_.height ? chunk_h_ - (end_y - cam_.height) : chunk_h_;
    Slot* slot = &slots_[next_write_];
    next_write_ = (next_write_ + 1) % 2;
    {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return !slot->full; });  // empty.acquire()
    }
    slot->off_x = off_x_; slot->off_y = off_y_; slot->w = w; slot->h = h;
    if (!device_renderer_render_chunk(dev_, off_x_, off_y_, w, h)) return -SRT_ERR_CUDA;
    i_++;
    const bool last = i_ == n_iterations_;
    slot->is_last = last;
    {
        std::lock_guard<std::mutex> lock(mu_);
        slot->full = true;  // full.release()
    }
    cv_.notify_all();
    off_x_ = (i_ % x_chunks_) * chunk_w_;
    off_y_ = (i_ / x_chunks_) * chunk_h_;
    if (last) done_ = true;
    return last ? 0 : 1;
}

int RenderManager::update_fb() {  // render_manager.cuh:68-142
    if (!device_inited_) { set_error("render manager: not initialised"); return -SRT_ERR_STATE; }
    Slot* slot = &slots_[next_read_];
    next_read_ = (next_read_ + 1) % 2;
    {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&] { return slot->full || worker_rc_ < 0; });  // full.acquire()
        if (!slot->full) { set_error(worker_error_); return worker_rc_; }
    }
    // the film already is in raster order: the reference's block-linear un-swizzle (:88-133) has no counterpart
    const bool ok = device_renderer_resolve(dev_, slot->off_x, slot->off_y, slot->w, slot->h, fb_r_, fb_g_, fb_b_, cam_.width, cam_.height);
    const bool last = slot->is_last;
    {
        std::lock_guard<std::mutex> lock(mu_);
        slot->full = false;  // empty.release()
    }
    cv_.notify_all();
    if (!ok) return -SRT_ERR_CUDA;
    return last ? 0 : 1;
}

int RenderManager::render_cycle() {  // render_manager.cuh:160-167
    end_render();
    if (!ready()) { set_error("render manager: not ready to render"); return SRT_ERR_STATE; }
    done_ = false;
    worker_rc_ = 0;
    worker_ = std::thread([this] {
        int rc;
        while ((rc = step()) > 0) {}
        if (rc < 0) {
            { std::lock_guard<std::mutex> lock(mu_); worker_rc_ = rc; worker_error_ = last_error(); }
            cv_.notify_all();
        }
    });
    worker_started_ = true;
    return SRT_OK;
}

int RenderManager::end_render() {  // render_manager.cuh:169-174
    if (worker_started_) {
        worker_.join();
        worker_started_ = false;
    }
    return SRT_OK;
}

int RenderManager::get_xyz(float* xyz) {
    if (!device_inited_) { set_error("render manager: not initialised"); return SRT_ERR_STATE; }
    return device_renderer_download_xyz(dev_, xyz) ? SRT_OK : SRT_ERR_CUDA;
}
float* RenderManager::device_film() { return dev_ ? device_renderer_film(dev_) : nullptr; }
int RenderManager::resolve_film() {
    if (!dev_) { set_error("render manager: not initialised"); return SRT_ERR_STATE; }
    return device_renderer_resolve(dev_, 0, 0, cam_.width, cam_.height, fb_r_, fb_g_, fb_b_, cam_.width, cam_.height) ? SRT_OK : SRT_ERR_CUDA;
}
int RenderManager::restart() {  // make the manager renderable again from its first chunk (same seeds, empty film)
    end_render();
    if (!device_inited_ || !dev_) { set_error("render manager: not initialised"); return SRT_ERR_STATE; }
    if (!device_renderer_reset(dev_)) return SRT_ERR_CUDA;
    i_ = 0;
    off_x_ = off_y_ = 0;
    next_write_ = next_read_ = 0;
    slots_[0] = Slot();
    slots_[1] = Slot();
    done_ = true;
    return SRT_OK;
}
int RenderManager::stats(srt_stats* s) const {
    if (!dev_) { set_error("render manager: not initialised"); return SRT_ERR_STATE; }
    device_renderer_stats(dev_, s);
    return SRT_OK;
}

}  // namespace srt
