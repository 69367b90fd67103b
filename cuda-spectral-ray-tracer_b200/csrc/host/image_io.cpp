// Image writers for the frame buffer (planar float 0..255): binary PPM and 24-bit BMP.
// Replaces io/save_image.cpp:8-13 (CImg .bmp) and io/io.cuh:10-23 (ASCII PPM to stdout).
#include "srt_host.hpp"
#include <cstdio>
#include <vector>

namespace srt {
static unsigned char to_byte(float v) { return (unsigned char)(v < 0.f ? 0.f : (v > 255.f ? 255.f : v)); }  // frame_buffer.cuh:31-37

bool write_ppm(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h) {
    FILE* f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("cannot open ") + path); return false; }
    std::fprintf(f, "P6\n%u %u\n255\n", w, h);
    std::vector<unsigned char> row(3ull * w);
    for (uint32_t y = 0; y < h; y++) {
        for (uint32_t x = 0; x < w; x++) {
            const size_t p = (size_t)y * w + x;
            row[3 * x] = to_byte(r[p]); row[3 * x + 1] = to_byte(g[p]); row[3 * x + 2] = to_byte(b[p]);
        }
        std::fwrite(row.data(), 1, row.size(), f);
    }
    std::fclose(f);
    return true;
}

bool write_bmp(const char* path, const float* r, const float* g, const float* b, uint32_t w, uint32_t h) {
    FILE* f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("cannot open ") + path); return false; }
    const uint32_t stride = (3 * w + 3) & ~3u, data = stride * h, off = 54, size = off + data;
    unsigned char hdr[54] = {'B', 'M'};
    auto put32 = [&](int at, uint32_t v) { hdr[at] = v & 255; hdr[at + 1] = (v >> 8) & 255; hdr[at + 2] = (v >> 16) & 255; hdr[at + 3] = (v >> 24) & 255; };
    put32(2, size); put32(10, off); put32(14, 40); put32(18, w); put32(22, h);
    hdr[26] = 1; hdr[28] = 24;
    put32(34, data); put32(38, 2835); put32(42, 2835);
    std::fwrite(hdr, 1, 54, f);
    std::vector<unsigned char> row(stride, 0);
    for (uint32_t y = 0; y < h; y++) {  // bottom-up, BGR
        const uint32_t sy = h - 1 - y;
        for (uint32_t x = 0; x < w; x++) {
            const size_t p = (size_t)sy * w + x;
            row[3 * x] = to_byte(b[p]); row[3 * x + 1] = to_byte(g[p]); row[3 * x + 2] = to_byte(r[p]);
        }
        std::fwrite(row.data(), 1, stride, f);
    }
    std::fclose(f);
    return true;
}
// frame_buffer holds 0..255 as float (frame_buffer.cuh:6-44) while the film crosses PCIe as one byte per channel: the widening loop,
// compiled once per instruction set the host may have (the loader picks the clone), vectorised by the compiler
__attribute__((target_clones("avx2", "default"), optimize("O3"))) void widen_u8_to_f32(const unsigned char* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = 0; i < n; i++) dst[i] = (float)src[i];
}

}  // namespace srt
