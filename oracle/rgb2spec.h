/* TEST INFRASTRUCTURE ONLY (oracle/).  See rgb2spec.c. */
#ifndef SRT_ORACLE_RGB2SPEC_H
#define SRT_ORACLE_RGB2SPEC_H
#ifdef __cplusplus
extern "C" {
#endif
/* Scale[k] of the missing table: float(smoothstep(smoothstep(k/(res-1)))). */
float srt_oracle_rgb2spec_scale(int k, int res);
/* Data[l][k][j][i][0..2] of the missing table (utils/srgb_to_spectrum.cuh:19), computed by
 * replaying the optimiser's warm-started sweep for that (l, j, i) column up/down to k.
 * Returns 0 if the Gauss-Newton system became singular. */
int srt_oracle_rgb2spec_cell(int l, int k, int j, int i, int res, float out[3]);
void srt_oracle_rgb2spec_eval(const float c[3], double rgb_out[3]);
#ifdef __cplusplus
}
#endif
#endif
