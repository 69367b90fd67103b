// Kernel parameter block and the per-FP-mode launch table of the path tracer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include "../common/srt_types.h"

#define SRT_BLOCK 256
#ifndef SRT_REFILL_LANES
#define SRT_REFILL_LANES 8  // k_trace_rays fetches new rays once this many lanes of a warp are idle
#endif
#ifndef SRT_TRACE_MIN_BLOCKS
#define SRT_TRACE_MIN_BLOCKS 5   // resident blocks per SM of k_trace_rays (48 registers; 4 and 6 blocks measured: -1 % / -6 %)
#endif
#ifndef SRT_WAVE_BLOCK
#define SRT_WAVE_BLOCK 256      // threads of a persistent wavefront block
#endif

namespace srt {

struct WaveParams {
    // scene (global memory; small scenes are staged into shared memory by every block)
    const SrtWide* nodes;  // four-wide traversal nodes
    const float4* grid;  // quantisation grid of the node boxes: lo.xyz, 1 / cell (written by the build)
    const SrtTri* tris;  // leaf order
    const SrtFlatUnit* flat_units;  // wide-leaf pre-test units (mode 2), triangles in flat order follow
    const SrtTri* flat_tris;
    int n_units;
    float flat_guard, flat_tol;     // wide leaf: grazing threshold on |n.d| / |d| (pairs with differing plane bits), and the `behind` tolerance on |D - n.o|
    const SrtMaterial* mats;
    const float* cie;    // x[95] y[95] z[95]
    const float* bg;     // background spectrum [95]
    int n_tris, n_mats, bg_is_zero;
    // camera and chunk geometry
    SrtCamera cam;
    uint32_t off_x, off_y, cw, ch;   // current chunk (pixels)
    uint32_t img_w, img_h;
    uint32_t nslots;                 // pixel slots of this rank: n_tiles * tile_w * tile_h
    const uint32_t* tiles;           // chunk-local tile ids owned by this rank ((tx + 5 ty) % world == rank)
    uint32_t n_tiles;
    uint32_t ref_grid_x;             // nominal_chunk_w / 28 + 1 (reference launch geometry -> seeds)
    uint32_t spp, bounce_limit;
    // one wavefront launch renders samples [s_begin, s_end) of every pixel it is handed (a "round"); the pixel's
    // XORWOW state waits in G0/G1 between rounds exactly as it does between chunks (rendering.cu:209,232)
    uint32_t s_begin, s_end;
    const uint32_t* order;   // pixel slots in the order they are handed out (most passes per sample first); null = 0..nslots-1
    uint32_t n_order;        // entries of `order` (all owned), or nslots
    uint32_t sched_flags;    // debug (SRT_OPT_SCHED_FLAGS): 1 no first-guess order, 4 no pooled remainders
    uint32_t* cost;          // per pixel slot: wavefront passes spent on it so far in this chunk (null = not recorded)
    unsigned long long* drain_clock;  // [0] globaltimer when the first local slot found no pixel left (min), [1] last block exit (max); null = off
    uint32_t strat_n;      // 0: the reference's sampler; n: stratified n x n sub-cells per pixel (n*n == spp)
    float strat_recip;     // 1 / n
    uint32_t tile_w, tile_h, tiles_x, rank, world;  // tiles_x: tiles per row of the nominal chunk
    // path state of the paths in flight, one record per (wavefront block, local slot), 16-byte vectors
    float4* R0;    // hit point xyz | triangle (leaf order)
    float4* R1;    // incoming direction xyz | valid[2:0] + bounce
    float4* P0;    // power 0..3
    float4* P1;    // power 4..6 | hero wavelength
    uint4* L0;     // XORWOW state of the pixel a local slot is rendering: d v0 v1 v2
    uint2* L1;     //                                                    v3 v4
    // per pixel slot: XORWOW state between pixels / chunks (seeded once, carried across chunks)
    uint4* G0;
    uint2* G1;
    uint32_t* next_slot;  // next pixel slot nobody renders yet (zeroed before every wavefront launch)
    float* acc;      // film: XYZ sums, 3 planes of `plane` floats, full-image raster
    size_t plane;
    // persistent-block wavefront: paths in flight per block and the shared-memory bytes its queues take
    uint32_t block_slots, queue_bytes;
    uint32_t tile_slots_log2, tile_w_log2;  // log2(tile_w * tile_h), log2(tile_w)
    uint32_t block_threads;  // threads of a wavefront block (128 or 256)
    uint32_t min_blocks;     // k_wavefront variant: 4 = 64 registers, 4 blocks per SM; 3 = 80 registers, 3 blocks per SM
    float scene_lo[3], scene_hi[3];  // bounding box of all triangles (host)
    unsigned long long* ray_counter;
    uint4* pass_log;  // debug (SRT_OPT_PASS_LOG): blocks 0..7 record {globaltimer ns (low 32 bits), regenerate, lambertian, metallic | dielectric << 16} per pass
};
#define SRT_PASS_LOG_BLOCKS 8
#define SRT_PASS_LOG_PASSES 8192

struct LaunchTable {
    size_t (*smem_bytes)(const WaveParams&, int mode);
    cudaError_t (*configure)(size_t smem_bytes);
    void (*init_slots)(const WaveParams&, cudaStream_t);
    void (*prior_cost)(const WaveParams&, cudaStream_t);
    void (*wavefront)(const WaveParams&, int mode, int grid, size_t smem, cudaStream_t);
    int (*wavefront_blocks_per_sm)(int mode, int min_blocks, int threads, size_t smem);  // resident blocks the hardware grants
    void (*megakernel)(const WaveParams&, int mode, int grid, size_t smem, cudaStream_t);
    void (*resolve)(const float* acc, size_t plane, uint32_t img_w, uint32_t ox, uint32_t oy, uint32_t w, uint32_t h, uint32_t spp, unsigned char* rgb,
                    cudaStream_t);
    void (*resolve_slice)(const float* acc, size_t plane, size_t first, uint32_t count, uint32_t out_plane, uint32_t spp, unsigned char* rgb, cudaStream_t);
    void (*trace_rays_flat)(const WaveParams&, uint32_t n, const float* o, const float* d, const uint32_t* flat_to_orig, float* t_out, int32_t* tri_out,
                            int grid, size_t smem, cudaStream_t);
    void (*trace_rays)(const WaveParams&, uint32_t n, const float* o, const float* d, const uint32_t* sorted_idx, float* t_out, int32_t* tri_out,
                       unsigned long long* counters, uint32_t* next_ray, int grid, cudaStream_t);
};

namespace fastfp { LaunchTable make_launch_table(); }
namespace strictfp_ { LaunchTable make_launch_table(); }

}  // namespace srt
