/* TEST INFRASTRUCTURE ONLY (oracle/) -- see lbvh_oracle.h. */
#include "lbvh_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static uint32_t expand_bits(uint32_t v) { /* 10 bits -> every third bit */
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static uint32_t quant10(float c, float lo, float hi) {
    float x = (c - lo) / (hi - lo);
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    return (uint32_t)x;
}
uint32_t srt_oracle_morton30(float cx, float cy, float cz, const float sb[6]) {
    uint32_t xx = expand_bits(quant10(cx, sb[0], sb[1]));
    uint32_t yy = expand_bits(quant10(cy, sb[2], sb[3]));
    uint32_t zz = expand_bits(quant10(cz, sb[4], sb[5]));
    return xx * 4 + yy * 2 + zz;
}
void srt_oracle_tri_centroid(const float v[9], float c[3]) {
    float inv = 1 / 3.f; /* vec3 operator/ multiplies by the reciprocal, math/vec3.cuh:144-147 */
    for (int a = 0; a < 3; a++) c[a] = inv * ((v[a] + v[3 + a]) + v[6 + a]);
}

static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
typedef struct { const uint32_t* key; int n; } keyset;
static inline int delta(const keyset* k, int i, int j) {
    if (j < 0 || j >= k->n) return -1;
    uint32_t a = k->key[i], b = k->key[j];
    if (a == b) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz32(a ^ b);
}

void srt_oracle_lbvh_build(int n, const float* lb, const float* cen, float* sb, uint32_t* codes, uint32_t* sorted_idx,
                           int32_t* left, int32_t* right, int32_t* parent, float* nb) {
    if (n <= 0) return;
    sb[0] = sb[2] = sb[4] = INFINITY;
    sb[1] = sb[3] = sb[5] = -INFINITY;
    for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) {
            sb[2 * a] = fminf(sb[2 * a], lb[6 * i + 2 * a]);
            sb[2 * a + 1] = fmaxf(sb[2 * a + 1], lb[6 * i + 2 * a + 1]);
        }
    for (int i = 0; i < n; i++) codes[i] = srt_oracle_morton30(cen[3 * i], cen[3 * i + 1], cen[3 * i + 2], sb);

    /* stable LSD radix sort of (code, index), 4 passes of 8 bits (30-bit keys) */
    uint32_t* ka = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* kb = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* ia = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* ib = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    for (int i = 0; i < n; i++) { ka[i] = codes[i]; ia[i] = (uint32_t)i; }
    for (int pass = 0; pass < 4; pass++) {
        size_t hist[257];
        memset(hist, 0, sizeof hist);
        int sh = 8 * pass;
        for (int i = 0; i < n; i++) hist[((ka[i] >> sh) & 255u) + 1]++;
        for (int b = 0; b < 256; b++) hist[b + 1] += hist[b];
        for (int i = 0; i < n; i++) {
            size_t p = hist[(ka[i] >> sh) & 255u]++;
            kb[p] = ka[i];
            ib[p] = ia[i];
        }
        uint32_t* t = ka; ka = kb; kb = t;
        t = ia; ia = ib; ib = t;
    }
    memcpy(sorted_idx, ia, sizeof(uint32_t) * (size_t)n);

    /* leaves */
    for (int k = 0; k < n; k++) memcpy(nb + 6 * (size_t)(n - 1 + k), lb + 6 * (size_t)ia[k], sizeof(float) * 6);
    for (int i = 0; i < 2 * n - 1; i++) parent[i] = -1;
    if (n == 1) { free(ka); free(kb); free(ia); free(ib); return; }

    keyset ks = {ka, n};
    for (int i = 0; i < n - 1; i++) {
        int d = (delta(&ks, i, i + 1) - delta(&ks, i, i - 1)) >= 0 ? 1 : -1;
        /* delta(i,i+1) == delta(i,i-1) cannot happen for distinct augmented keys except at i = 0 (-1 on the left) */
        int dmin = delta(&ks, i, i - d);
        int lmax = 2;
        while (delta(&ks, i, i + lmax * d) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2)
            if (delta(&ks, i, i + (l + t) * d) > dmin) l += t;
        int j = i + l * d;
        int dnode = delta(&ks, i, j);
        int s = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(&ks, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        int gamma = i + s * d + (d < 0 ? d : 0);
        int lo = i < j ? i : j, hi = i < j ? j : i;
        int L = (lo == gamma) ? (n - 1 + gamma) : gamma;
        int R = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
        left[i] = L; right[i] = R;
        parent[L] = i; parent[R] = i;
    }
    /* bottom-up refit: every internal node once both children are done (order-independent unions) */
    int* visits = (int*)calloc((size_t)(n - 1), sizeof(int));
    for (int k = 0; k < n; k++) {
        int node = parent[n - 1 + k];
        while (node >= 0) {
            if (visits[node]++ == 0) break; /* first arrival waits for the sibling */
            const float* a = nb + 6 * (size_t)left[node];
            const float* b = nb + 6 * (size_t)right[node];
            float* o = nb + 6 * (size_t)node;
            for (int c = 0; c < 3; c++) { o[2 * c] = fminf(a[2 * c], b[2 * c]); o[2 * c + 1] = fmaxf(a[2 * c + 1], b[2 * c + 1]); }
            node = parent[node];
        }
    }
    free(visits); free(ka); free(kb); free(ia); free(ib);
}
