// Known-answer probe for SURVEY quirk Q13: in which order does nvcc DEVICE code evaluate the
// three side-effecting arguments of `vec3(f(), f(), f())` (reference math/vec3.cuh:102-109)?
// C++ leaves it unspecified; g++ goes right-to-left.  Prints "ltr" or "rtl" (or "other").
#include <cstdio>
#include <cuda_runtime.h>
struct v3 {
    float e[3];
    __host__ __device__ v3(float a, float b, float c) : e{a, b, c} {}
};
__device__ __noinline__ float draw(int* counter) { return (float)((*counter)++); }
__device__ __noinline__ float draw_range(float lo, float hi, int* counter) { return draw(counter) * (hi - lo) + lo; }
__global__ void probe(float* out) {
    int c = 0;
    v3 a = v3(draw(&c), draw(&c), draw(&c));
    int d = 0;
    v3 b = v3(draw_range(0, 1, &d), draw_range(0, 1, &d), draw_range(0, 1, &d));
    for (int k = 0; k < 3; k++) { out[k] = a.e[k]; out[3 + k] = b.e[k]; }
}
int main() {
    float* d;
    float h[6];
    if (cudaMalloc(&d, sizeof h) != cudaSuccess) { printf("no device\n"); return 1; }
    probe<<<1, 1>>>(d);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const bool ltr = h[0] == 0 && h[1] == 1 && h[2] == 2 && h[3] == 0 && h[4] == 1 && h[5] == 2;
    const bool rtl = h[0] == 2 && h[1] == 1 && h[2] == 0 && h[3] == 2 && h[4] == 1 && h[5] == 0;
    printf("q13 device argument evaluation order: %s (%g %g %g | %g %g %g)\n", ltr ? "ltr" : (rtl ? "rtl" : "other"), h[0], h[1], h[2], h[3], h[4], h[5]);
    return 0;
}
