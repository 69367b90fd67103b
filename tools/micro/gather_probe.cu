// Gather micro-benchmark behind the LBVH walk's node layout: every lane of a warp follows its own chain of
// random 64-B (or 32-B) records -- the next index comes out of the record just read, like a child pointer -- fetched as
// 4 x LDG.128, 2 x LDG.256 (sm_100: ld.global.v8.f32), 2 x LDG.128 or 1 x LDG.256.  Prints ns per record per lane-chain
// and records/s: the L1 wavefront cost of a divergent node fetch.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a gather_probe.cu -o gather_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld256(const void* p) {
    F8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_chase(const float4* __restrict__ rec, uint32_t mask, int steps, float* out) {
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask;
    float acc = 0.f;
    for (int s = 0; s < steps; s++) {
        if (MODE == 0) {  // 64 B as 4 x 16 B
            const float4* p = rec + 4ull * i;
            const float4 a = p[0], b = p[1], c = p[2], d = p[3];
            acc += a.x + b.y + c.z;
            i = __float_as_uint(d.x) & mask;
        } else if (MODE == 1) {  // 64 B as 2 x 32 B
            const F8 a = ld256(rec + 4ull * i), b = ld256(rec + 4ull * i + 2);
            acc += a.v[0] + a.v[5] + b.v[2];
            i = __float_as_uint(b.v[4]) & mask;
        } else if (MODE == 2) {  // 32 B as 2 x 16 B
            const float4* p = rec + 2ull * i;
            const float4 a = p[0], d = p[1];
            acc += a.x + d.y;
            i = __float_as_uint(d.x) & mask;
        } else {  // 32 B as 1 x 32 B
            const F8 a = ld256(rec + 2ull * i);
            acc += a.v[0] + a.v[5];
            i = __float_as_uint(a.v[4]) & mask;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + i;
}
int main(int argc, char** argv) {
    const int log2n = argc > 1 ? atoi(argv[1]) : 20;   // records
    const int steps = 64;
    const uint32_t n = 1u << log2n, mask = n - 1;
    std::vector<float> h(16ull * n);
    uint32_t x = 12345;
    for (size_t r = 0; r < n; r++)
        for (int k = 0; k < 16; k++) { x = x * 1664525u + 1013904223u; uint32_t v = (x >> 8) & mask; h[16 * r + k] = *reinterpret_cast<float*>(&v); }
    float4* d; float* out;
    cudaMalloc(&d, 64ull * n); cudaMalloc(&out, 4ull << 20);
    cudaMemcpy(d, h.data(), 64ull * n, cudaMemcpyHostToDevice);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[4] = {"64 B = 4 x LDG.128", "64 B = 2 x LDG.256", "32 B = 2 x LDG.128", "32 B = 1 x LDG.256"};
    for (int bps = 2; bps <= 8; bps += 3)
        for (int mode = 0; mode < 4; mode++) {
            const int grid = sms * bps;
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k_chase<0><<<grid, 256>>>(d, mask, steps, out);
                if (mode == 1) k_chase<1><<<grid, 256>>>(d, mask, steps, out);
                if (mode == 2) k_chase<2><<<grid, 256>>>(d, mask >> 0, steps, out);
                if (mode == 3) k_chase<3><<<grid, 256>>>(d, mask >> 0, steps, out);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            const double recs = (double)grid * 256 * steps;
            printf("records 2^%d, %d blocks/SM, %-20s: %.3f ms, %.2f G records/s, %.1f cycles per warp-step per SM-resident warp set\n", log2n, bps, names[mode], best,
                   recs / best / 1e6, best * 1e-3 * 1.965e9 / (steps * bps * 8.0));
        }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
