// Small CUDA host helpers shared by the .cu files of libsrt.so.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include "../host/srt_host.hpp"

namespace srt {
extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }
bool cuda_ok(cudaError_t e, const char* what, const char* file, int line);
// process-wide cache of device buffers (renderer.cu): a freed block is handed to the next request of a similar size
bool device_pool_alloc(void** out, size_t bytes);
void device_pool_free(void* p);
// cost-ordered pixel slot lists for the wavefront renderer's rounds (lbvh.cu, next to the radix sort it reuses)
struct PixelOrder;
PixelOrder* pixel_order_create(uint32_t max_slots);
void pixel_order_destroy(PixelOrder*);
const uint32_t* pixel_order_build(PixelOrder*, const uint32_t* cost, uint32_t n, uint32_t samples, cudaStream_t st);
}  // namespace srt

// SRT_TRACE=1 in the environment: host wall-clock of the set-up / exchange phases on stderr (a debugging aid)
struct PhaseTrace {
    const char* what;
    std::chrono::steady_clock::time_point t0;
    bool on;
    explicit PhaseTrace(const char* w) : what(w), t0(std::chrono::steady_clock::now()) {
        static const bool enabled = getenv("SRT_TRACE") != nullptr;
        on = enabled;
    }
    void mark(const char* phase) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[srt trace] %s: %s %.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

#define SRT_CUDA(call)                                                   \
    do {                                                                 \
        if (!::srt::cuda_ok((call), #call, __FILE__, __LINE__)) return false; \
    } while (0)
#define SRT_CUDA_LAST() SRT_CUDA(cudaGetLastError())
