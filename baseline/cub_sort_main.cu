// BASELINE INFRASTRUCTURE (timing context only): CUB's DeviceRadixSort::SortPairs on the same problem
// as the LBVH's Morton sort -- n (30-bit key, 32-bit index) pairs -- so that the hand-written onesweep in
// csrc/cuda/lbvh.cu has a library number next to it.  Not linked into libsrt.so, not on any product path.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cub/device/device_radix_sort.cuh>

static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : (1u << 20);
    const int reps = argc > 2 ? atoi(argv[2]) : 20;
    std::vector<uint32_t> hk(n), hv(n);
    uint64_t s = 1984;
    for (size_t i = 0; i < n; i++) { hk[i] = (uint32_t)(splitmix(s) & 0x3FFFFFFFu); hv[i] = (uint32_t)i; }
    uint32_t *k0, *k1, *v0, *v1;
    cudaMalloc(&k0, n * 4); cudaMalloc(&k1, n * 4); cudaMalloc(&v0, n * 4); cudaMalloc(&v1, n * 4);
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k0, k1, v0, v1, (int)n, 0, 30);
    cudaMalloc(&tmp, tmp_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ms30, ms32;
    for (int bits = 30; bits <= 32; bits += 2)
        for (int r = 0; r < reps + 3; r++) {
            cudaMemcpy(k0, hk.data(), n * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(v0, hv.data(), n * 4, cudaMemcpyHostToDevice);
            cudaEventRecord(e0);
            cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k0, k1, v0, v1, (int)n, 0, bits);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 3) (bits == 30 ? ms30 : ms32).push_back(ms);
        }
    if (cudaGetLastError() != cudaSuccess) { printf("{\"error\": \"cuda\"}\n"); return 1; }
    std::sort(ms30.begin(), ms30.end()); std::sort(ms32.begin(), ms32.end());
    printf("{\"n\": %zu, \"cub_sort_pairs_30bit_ms\": %.5f, \"cub_sort_pairs_32bit_ms\": %.5f, \"reps\": %d}\n", n, ms30[ms30.size() / 2], ms32[ms32.size() / 2], reps);
    return 0;
}
