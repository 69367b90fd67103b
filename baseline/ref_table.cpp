// BASELINE INFRASTRUCTURE: stands in for the reference's missing utils/srgb_to_spectrum.cu
// (.MISSING_LARGE_BLOBS:1).  Defines the three symbols of utils/srgb_to_spectrum.cuh:17-19 (writable
// here; this TU does not include that header) and fills, before main(), the table cells the three
// shipped scenes read, using the oracle's restatement of rgb2spec_opt (oracle/rgb2spec.c).
#include "rgb2spec.h"
#include <algorithm>
int sRGBToSpectrumTable_Res = 64;
float sRGBToSpectrumTable_Scale[64];
float sRGBToSpectrumTable_Data[3][64][64][64][3];
namespace {
void fill(float r, float g, float b) {
    if (r == g && g == b) return;
    float rgb[3] = {r, g, b};
    int maxc = (r > g) ? ((r > b) ? 0 : 2) : ((g > b) ? 1 : 2);
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * 63 / z, y = rgb[(maxc + 2) % 3] * 63 / z;
    int xi = std::min((int)x, 62), yi = std::min((int)y, 62), zi = 0;
    while (zi < 62 && sRGBToSpectrumTable_Scale[zi + 1] < z) ++zi;
    for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 2; ++dy) for (int dx = 0; dx < 2; ++dx) {
        float c[3];
        if (zi + dz > 63 || yi + dy > 63 || xi + dx > 63) continue;
        if (!srt_oracle_rgb2spec_cell(maxc, zi + dz, yi + dy, xi + dx, 64, c)) continue;
        for (int q = 0; q < 3; ++q) sRGBToSpectrumTable_Data[maxc][zi + dz][yi + dy][xi + dx][q] = c[q];
    }
}
struct Init {
    Init() {
        for (int k = 0; k < 64; ++k) sRGBToSpectrumTable_Scale[k] = srt_oracle_rgb2spec_scale(k, 64);
        fill(.65f, .05f, .05f); fill(.12f, .45f, .15f); fill(.12f, .15f, .45f);
    }
} init_table;
}  // namespace
