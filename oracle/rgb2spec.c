/* TEST INFRASTRUCTURE ONLY (oracle/) -- never linked into the product library.
 *
 * Regenerates cells of the reference's MISSING table utils/srgb_to_spectrum.cu
 * (/root/reference/.MISSING_LARGE_BLOBS:1; declared in utils/srgb_to_spectrum.cuh:17-19 as
 * float[3][64][64][64][3] + Scale[64]).  The table is the output of pbrt-v4's
 * src/pbrt/cmd/rgb2spec_opt.cpp @ 39e01e61 (Jakob & Hanika 2019, "A Low-Dimensional Function
 * Space for Efficient Spectral Upsampling"), a third-party program that is NOT vendored in
 * /root/reference.  This file restates its published algorithm for the sRGB gamut:
 *   - 3/8-Simpson quadrature over 283 fine samples of x_bar,y_bar,z_bar * D65
 *   - residual in CIELAB between the target rgb and the integrated sigmoid(poly) spectrum
 *   - Gauss-Newton (<= 15 iterations), central-difference Jacobian, 3x3 LU with pivoting
 *   - per (max-channel l, y=j, x=i): warm-started sweep over the brightness index k, starting
 *     at k0 = res/5 going up, then (restarted from 0) going down
 *   - coefficients remapped from the normalised [0,1] wavelength to nanometres.
 * PARITY UNPINNED: the reference ships no copy of these numbers and no test that pins them;
 * only self-consistency (round trip rgb -> spectrum -> rgb) is checked in tests/.
 */
#include "rgb2spec.h"
#include "cie_tables.h"
#include <math.h>
#include <string.h>

#define FINE_SAMPLES ((SRT_CIE_SAMPLES - 1) * 3 + 1) /* 283 */
#define R2S_EPS 1e-4
#define L_MIN 360.0
#define L_MAX 830.0

static const double xyz_to_srgb[3][3] = {{3.240479, -1.537150, -0.498535},
                                         {-0.969256, 1.875991, 0.041556},
                                         {0.055648, -0.204043, 1.057311}};
static const double srgb_to_xyz[3][3] = {{0.412453, 0.357580, 0.180423},
                                         {0.212671, 0.715160, 0.072169},
                                         {0.019334, 0.119193, 0.950227}};

static double lambda_tbl[FINE_SAMPLES];
static double rgb_tbl[3][FINE_SAMPLES];
static double xyz_white[3];
static int tables_ready = 0;

static double tbl_interp(int column, double x) {
    x -= L_MIN;
    x *= (SRT_CIE_SAMPLES - 1) / (L_MAX - L_MIN);
    int offset = (int)x;
    if (offset < 0) offset = 0;
    if (offset > SRT_CIE_SAMPLES - 2) offset = SRT_CIE_SAMPLES - 2;
    double w = x - offset;
    double a = srt_cie_rows[offset][column], b = srt_cie_rows[offset + 1][column];
    if (column == 3) { a /= SRT_D65_NORM; b /= SRT_D65_NORM; }
    return (1.0 - w) * a + w * b;
}

static void init_tables(void) {
    if (tables_ready) return;
    memset(rgb_tbl, 0, sizeof rgb_tbl);
    memset(xyz_white, 0, sizeof xyz_white);
    const double h = (L_MAX - L_MIN) / (FINE_SAMPLES - 1);
    for (int i = 0; i < FINE_SAMPLES; ++i) {
        double lambda = L_MIN + i * h;
        double xyz[3] = {tbl_interp(0, lambda), tbl_interp(1, lambda), tbl_interp(2, lambda)};
        double I = tbl_interp(3, lambda);
        double weight = 3.0 / 8.0 * h;
        if (i == 0 || i == FINE_SAMPLES - 1) {
        } else if ((i - 1) % 3 == 2) weight *= 2.0;
        else weight *= 3.0;
        lambda_tbl[i] = lambda;
        for (int k = 0; k < 3; ++k)
            for (int j = 0; j < 3; ++j) rgb_tbl[k][i] += xyz_to_srgb[k][j] * xyz[j] * I * weight;
        for (int k = 0; k < 3; ++k) xyz_white[k] += xyz[k] * I * weight;
    }
    tables_ready = 1;
}

static double sigmoid(double x) { return 0.5 * x / sqrt(1.0 + x * x) + 0.5; }
static double smoothstep(double x) { return x * x * (3.0 - 2.0 * x); }

static double lab_f(double t) {
    const double delta = 6.0 / 29.0;
    if (t > delta * delta * delta) return cbrt(t);
    return t / (delta * delta * 3.0) + (4.0 / 29.0);
}

static void cie_lab(double* p) {
    double X = 0.0, Y = 0.0, Z = 0.0;
    for (int j = 0; j < 3; ++j) {
        X += p[j] * srgb_to_xyz[0][j];
        Y += p[j] * srgb_to_xyz[1][j];
        Z += p[j] * srgb_to_xyz[2][j];
    }
    double fx = lab_f(X / xyz_white[0]), fy = lab_f(Y / xyz_white[1]), fz = lab_f(Z / xyz_white[2]);
    p[0] = 116.0 * fy - 16.0;
    p[1] = 500.0 * (fx - fy);
    p[2] = 200.0 * (fy - fz);
}

static void eval_residual(const double* coeffs, const double* rgb, double* residual) {
    double out[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < FINE_SAMPLES; ++i) {
        double lambda = (lambda_tbl[i] - L_MIN) / (L_MAX - L_MIN);
        double x = 0.0;
        for (int c = 0; c < 3; ++c) x = x * lambda + coeffs[c];
        double s = sigmoid(x);
        for (int j = 0; j < 3; ++j) out[j] += rgb_tbl[j][i] * s;
    }
    cie_lab(out);
    memcpy(residual, rgb, sizeof(double) * 3);
    cie_lab(residual);
    for (int j = 0; j < 3; ++j) residual[j] -= out[j];
}

static void eval_jacobian(const double* coeffs, const double* rgb, double jac[3][3]) {
    double r0[3], r1[3], tmp[3];
    for (int i = 0; i < 3; ++i) {
        memcpy(tmp, coeffs, sizeof tmp);
        tmp[i] -= R2S_EPS;
        eval_residual(tmp, rgb, r0);
        memcpy(tmp, coeffs, sizeof tmp);
        tmp[i] += R2S_EPS;
        eval_residual(tmp, rgb, r1);
        for (int j = 0; j < 3; ++j) jac[j][i] = (r1[j] - r0[j]) * 1.0 / (2 * R2S_EPS);
    }
}

/* LU decomposition with partial pivoting on row pointers (the classic textbook routine the
 * optimiser uses); returns 0 when the matrix is numerically singular. */
static int lup_decompose(double* A[3], int P[4]) {
    for (int i = 0; i <= 3; ++i) P[i] = i;
    for (int i = 0; i < 3; ++i) {
        double maxA = 0.0;
        int imax = i;
        for (int k = i; k < 3; ++k) {
            double a = fabs(A[k][i]);
            if (a > maxA) { maxA = a; imax = k; }
        }
        if (maxA < 1e-15) return 0;
        if (imax != i) {
            int j = P[i]; P[i] = P[imax]; P[imax] = j;
            double* ptr = A[i]; A[i] = A[imax]; A[imax] = ptr;
            P[3]++;
        }
        for (int j = i + 1; j < 3; ++j) {
            A[j][i] /= A[i][i];
            for (int k = i + 1; k < 3; ++k) A[j][k] -= A[j][i] * A[i][k];
        }
    }
    return 1;
}

static void lup_solve(double* A[3], const int P[4], const double* b, double* x) {
    for (int i = 0; i < 3; ++i) {
        x[i] = b[P[i]];
        for (int k = 0; k < i; ++k) x[i] -= A[i][k] * x[k];
    }
    for (int i = 2; i >= 0; --i) {
        for (int k = i + 1; k < 3; ++k) x[i] -= A[i][k] * x[k];
        x[i] /= A[i][i];
    }
}

static int gauss_newton(const double rgb[3], double coeffs[3]) {
    for (int it = 0; it < 15; ++it) {
        double jac[3][3], residual[3], x[3];
        double* J[3] = {jac[0], jac[1], jac[2]};
        int P[4];
        eval_residual(coeffs, rgb, residual);
        eval_jacobian(coeffs, rgb, jac);
        if (!lup_decompose(J, P)) return 0;
        lup_solve(J, P, residual, x);
        double r = 0.0;
        for (int j = 0; j < 3; ++j) {
            coeffs[j] -= x[j];
            r += residual[j] * residual[j];
        }
        double mx = fmax(fmax(coeffs[0], coeffs[1]), coeffs[2]);
        if (mx > 200) {
            for (int j = 0; j < 3; ++j) coeffs[j] *= 200 / mx;
        }
        if (r < 1e-6) break;
    }
    return 1;
}

float srt_oracle_rgb2spec_scale(int k, int res) {
    return (float)smoothstep(smoothstep(k / (double)(res - 1)));
}

int srt_oracle_rgb2spec_cell(int l, int k, int j, int i, int res, float out[3]) {
    init_tables();
    const double y = j / (double)(res - 1), x = i / (double)(res - 1);
    const int start = res / 5;
    double coeffs[3] = {0.0, 0.0, 0.0}, rgb[3];
    const int step = (k >= start) ? 1 : -1;
    for (int kk = start;; kk += step) {
        double b = (double)srt_oracle_rgb2spec_scale(kk, res);
        rgb[l] = b;
        rgb[(l + 1) % 3] = x * b;
        rgb[(l + 2) % 3] = y * b;
        if (!gauss_newton(rgb, coeffs)) return 0;
        if (kk == k) break;
    }
    const double c0 = 360.0, c1 = 1.0 / (830.0 - 360.0);
    const double A = coeffs[0], B = coeffs[1], C = coeffs[2];
    out[0] = (float)(A * (c1 * c1));
    out[1] = (float)(B * c1 - 2 * A * c0 * (c1 * c1));
    out[2] = (float)(C - B * c0 * c1 + A * ((c0 * c1) * (c0 * c1)));
    return 1;
}

/* forward model used by the self-consistency test: integrates sigmoid(c0 l^2 + c1 l + c2)
 * against the D65-weighted sRGB matching curves; returns linear sRGB. */
void srt_oracle_rgb2spec_eval(const float c[3], double rgb_out[3]) {
    init_tables();
    rgb_out[0] = rgb_out[1] = rgb_out[2] = 0.0;
    for (int i = 0; i < FINE_SAMPLES; ++i) {
        double l = lambda_tbl[i];
        double s = sigmoid(((double)c[0] * l + (double)c[1]) * l + (double)c[2]);
        for (int j = 0; j < 3; ++j) rgb_out[j] += rgb_tbl[j][i] * s;
    }
}
