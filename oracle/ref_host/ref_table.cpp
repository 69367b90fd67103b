// TEST INFRASTRUCTURE ONLY (oracle/): definitions for the reference's MISSING translation unit
// utils/srgb_to_spectrum.cu (.MISSING_LARGE_BLOBS:1).  The header utils/srgb_to_spectrum.cuh:17-19
// declares these three symbols `extern const`; this TU deliberately does not include that header
// and defines them writable (variable names are not type-mangled), so individual cells can be
// filled on demand from the oracle's restatement of rgb2spec_opt (oracle/rgb2spec.c) instead
// of shipping 9.4 MB of numbers.  PARITY UNPINNED: no reference artefact pins these values.
#include "../rgb2spec.h"
#include <algorithm>

int sRGBToSpectrumTable_Res = 64;
float sRGBToSpectrumTable_Scale[64];
float sRGBToSpectrumTable_Data[3][64][64][64][3];
static unsigned char cell_done[3][64][64][64];
static bool scale_done = false;

extern "C" void srt_ref_table_init_scale() {
    if (scale_done) return;
    for (int k = 0; k < 64; ++k) sRGBToSpectrumTable_Scale[k] = srt_oracle_rgb2spec_scale(k, 64);
    scale_done = true;
}

static void fill_cell(int l, int k, int j, int i) {
    if (l < 0 || l > 2 || k < 0 || k > 63 || j < 0 || j > 63 || i < 0 || i > 63) return;
    if (cell_done[l][k][j][i]) return;
    float c[3];
    if (!srt_oracle_rgb2spec_cell(l, k, j, i, 64, c)) c[0] = c[1] = c[2] = 0.f;
    for (int q = 0; q < 3; ++q) sRGBToSpectrumTable_Data[l][k][j][i][q] = c[q];
    cell_done[l][k][j][i] = 1;
}

// Fill every cell get_sigmoid_coeffs / dev_get_sigmoid_coeffs (color/color_to_spectrum.cuh:69-151)
// can touch for this colour: the 2x2x2 trilinear neighbourhood of the host path, which contains
// the single nearest cell the device path reads.
extern "C" void srt_ref_table_fill_for_color(float r, float g, float b) {
    srt_ref_table_init_scale();
    if (r == g && g == b) return;
    float rgb[3] = {r, g, b};
    int maxc = (r > g) ? ((r > b) ? 0 : 2) : ((g > b) ? 1 : 2);
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * 63 / z;
    float y = rgb[(maxc + 2) % 3] * 63 / z;
    int xi = std::min((int)x, 62), yi = std::min((int)y, 62);
    // zi = last index whose Scale is < z, clamped to [0, 62]
    int zi = 0;
    while (zi < 62 && sRGBToSpectrumTable_Scale[zi + 1] < z) ++zi;
    for (int dz = 0; dz < 2; ++dz)
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) fill_cell(maxc, zi + dz, yi + dy, xi + dx);
}

extern "C" float* srt_ref_table_data() { return &sRGBToSpectrumTable_Data[0][0][0][0][0]; }
