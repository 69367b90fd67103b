// LBVH construction on the device (Karras 2012) + the standalone closest-hit query kernel.
//
//   bounds   : leaf box = padded triangle box, centroid, one partial scene box per block; also clears what the later
//              phases count in (refit arrival flags, look-back status words, digit histograms, tile tickets)
//   morton   : scene box = union of the partial boxes; 30-bit code per triangle (10 bits per axis, x highest)
//   sort     : hand-written ONESWEEP least-significant-digit radix sort, 4 passes of 8 bits,
//              one global histogram pre-pass, decoupled look-back, warp match/ballot ranking,
//              stable => equal codes stay in triangle-index order
//   tree     : hierarchy AND refit in one bottom-up kernel, one thread per leaf: a finished subtree [a, b] picks its parent by
//              comparing the key differences at its two ends (the split with the longer common prefix is the nearer ancestor),
//              the second subtree to arrive at a split unions both child boxes and records the node's children and box under
//              the node's Karras index (no fences: published boxes carry the build's epoch).  There is no separate top-down
//              split search; left / right / parent arrays are derived from the children only when a caller dumps them
//   collapse : the 4-wide traversal nodes the walk reads (srt_types.h): one streaming pass, same indices, child boxes quantised
//              to the 16-bit scene grid with all lanes at work
//   permute  : triangles to leaf order -- a pure gather, on a second stream next to the tree kernel
//
// Specification and bit-exactness oracle: oracle/lbvh_oracle.c.  The reference
// itself builds its BVH serially in one thread (bvh/bvh.cu:206-345, scene/scene.cu:9-20); this
// file replaces that step.  Compiled with -fmad=false: everything here is bandwidth bound and
// the Morton quantisation must round exactly like the CPU specification.
#include "cuda_common.cuh"
#include <cfloat>
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cmath>

namespace srt {

std::atomic<uint64_t> g_kernel_launches{0};
static std::atomic<uint32_t> g_refit_epoch{0};  // marks the node boxes a build has published (k_refit_emit)
uint64_t kernel_launches() { return g_kernel_launches.load(); }

bool cuda_ok(cudaError_t e, const char* what, const char* file, int line) {
    if (e == cudaSuccess) return true;
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d '%s'", (int)e, cudaGetErrorString(e), file, line, what);
    set_error(buf);
    return false;
}
int cuda_device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
bool cuda_select_device(int dev) { SRT_CUDA(cudaSetDevice(dev)); return true; }

// ------------------------------------------------------------------------------------------
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_PASSES = 4;
constexpr int SORT_THREADS = 256;  // tiles of 4096 keys; small sorts use 512 threads and 8192-key tiles (k_onesweep<512>, sort_threads_for)
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per tile
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr size_t sort_smem(int threads) { return 2ull * threads * SORT_ITEMS * sizeof(uint32_t); }  // dynamic shared memory of k_onesweep (the sorted tile)
#ifndef SRT_SORT_BALLOT
#define SRT_SORT_BALLOT 1
#endif
#ifndef SRT_SORT_LOOKBACK
#define SRT_SORT_LOOKBACK 4
#endif
#ifndef SRT_SORT_LOOKBACK_512
#define SRT_SORT_LOOKBACK_512 4  // the one-tile-per-SM kernel of small sorts: all tiles run at once, every walk is long
#endif
constexpr int SORT_LOOKBACK = SRT_SORT_LOOKBACK;  // predecessor status words read per round of the decoupled look-back

struct DeviceScene {
    uint32_t n = 0, n_mats = 0;
    // inputs
    float* verts = nullptr;       // n x 9
    SrtTri* tris_in = nullptr;    // n, original order
    SrtMaterial* mats = nullptr;
    // build products
    float4* leaf_boxes = nullptr;  // n x (lo.xyz -, hi.xyz -), original order: one 32-byte sector per leaf for the refit's gather
    float* centroids = nullptr;   // n x 3
    float* scene_box = nullptr;   // 6 floats + at [8..11] the grid of the traversal nodes: lo.xyz, 1 / cell
    uint32_t* codes = nullptr;    // n, original order
    uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
    uint32_t* hist = nullptr;     // SORT_PASSES x RADIX global digit counts -> exclusive bases
    uint32_t* lookback = nullptr; // SORT_PASSES x tiles x RADIX status words
    uint32_t* tile_counter = nullptr;  // SORT_PASSES dynamic tile ids
    int32_t *left = nullptr, *right = nullptr, *parent = nullptr;
    float* block_boxes = nullptr;  // kBoundsBlocks x 6 partial scene boxes
    float4* node_box = nullptr;   // 2 x (2n-1): node i at [2i] = (xmin,ymin,zmin,epoch), [2i+1] = (xmax,ymax,zmax,epoch): one 32-byte sector per box
    uint32_t* visit = nullptr;    // n-1 refit arrival flags
    int2* children = nullptr;     // n-1 (left, right) per Karras node number: >= 0 internal number, < 0 leaf ~k
    SrtWide* wide = nullptr;      // n-1 four-wide traversal nodes, same indices
    SrtTri* tris = nullptr;       // n, LEAF order
    // wide leaf (scenes of <= 32 pre-test units): flat-order triangles + units, built on the host
    SrtFlatUnit* flat_units = nullptr;
    SrtTri* flat_tris = nullptr;
    uint32_t* flat_to_orig = nullptr;  // flat triangle position -> original triangle index
    uint32_t n_units = 0;
    float flat_guard = 0.f, flat_tol = 0.f;
    double origin_l1_bound = 0;
    float host_lo[3] = {0, 0, 0}, host_hi[3] = {0, 0, 0};  // union of the triangle boxes (host)
    uint32_t tiles = 0;
    cudaEvent_t ev[6];
    double last_build_ms = 0;
    cudaStream_t stream = nullptr;   // the caller's stream (null = default)
    cudaStream_t side = nullptr;     // triangle permutation next to hierarchy + refit
    cudaEvent_t ev_sorted = nullptr, ev_side = nullptr;
};

// ---- float <-> order-preserving uint for atomic min/max ----
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

constexpr int kBoundsBlocks = 592;  // 4 per SM (2 per SM left the loads without enough warps to hide their latency); every Morton block re-reduces the partial boxes

// leaf box (bvh/aabb.cuh:49-57 + pad :93-102), centroid (primitives/tri.cuh:73-77), one partial scene box per block
// (min / max are exact, so any reduction order gives the bits of the serial union); and the housekeeping of the build:
// refit arrival flags, look-back status words, digit histograms and tile tickets start at zero.
__global__ void __launch_bounds__(256) k_bounds(const float* __restrict__ verts, uint32_t n, float4* __restrict__ leaf_boxes,
                                                float* __restrict__ centroids, float* __restrict__ block_boxes, uint32_t* __restrict__ visit,
                                                uint32_t* __restrict__ lookback, uint32_t n_lookback, uint32_t* __restrict__ hist, uint32_t n_hist,
                                                uint32_t* __restrict__ tickets, uint32_t n_tickets) {
    __shared__ float sbox[8][6];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
    for (uint32_t i = gtid; i < n; i += gsize) visit[i] = 0xFFFFFFFFu;  // k_build_tree: no subtree has arrived at this split yet
    for (uint32_t i = gtid; i < n_lookback; i += gsize) lookback[i] = 0;
    if (gtid < n_hist) hist[gtid] = 0;
    if (gtid < n_tickets) tickets[gtid] = 0;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = gtid; i < n; i += gsize) {
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; k++) v[k] = verts[9ull * i + k];
        const float third = 1 / 3.f;
        float bl[3], bh[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float mn = fminf(v[a], fminf(v[3 + a], v[6 + a]));
            float mx = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
            if (!((mx - mn) >= 0.0001f)) {
                const float padding = 0.0001f / 2;
                mn = mn - padding;
                mx = mx + padding;
            }
            bl[a] = mn;
            bh[a] = mx;
            centroids[3ull * i + a] = third * ((v[a] + v[3 + a]) + v[6 + a]);
            lo[a] = fminf(lo[a], mn);
            hi[a] = fmaxf(hi[a], mx);
        }
        leaf_boxes[2ull * i] = make_float4(bl[0], bl[1], bl[2], 0.f);
        leaf_boxes[2ull * i + 1] = make_float4(bh[0], bh[1], bh[2], 0.f);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0) { sbox[threadIdx.x >> 5][2 * a] = lo[a]; sbox[threadIdx.x >> 5][2 * a + 1] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = sbox[0][threadIdx.x];
        for (int w = 1; w < 8; w++) v = (threadIdx.x & 1) ? fmaxf(v, sbox[w][threadIdx.x]) : fminf(v, sbox[w][threadIdx.x]);
        block_boxes[6 * blockIdx.x + threadIdx.x] = v;
    }
}

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t quant10(float c, float lo, float hi) {
    float x = (c - lo) / (hi - lo);
    x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
    return (uint32_t)x;
}
// Morton code per triangle + identity payload + the global digit histogram of all 4 passes
__global__ void __launch_bounds__(256) k_morton_hist(const float* __restrict__ centroids, const float* __restrict__ block_boxes, int n_block_boxes,
                                                     float* __restrict__ scene_box, uint32_t n, uint32_t* __restrict__ codes, uint32_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[SORT_PASSES * RADIX];
    __shared__ float sbox[8][6], box[6];
    for (int i = threadIdx.x; i < SORT_PASSES * RADIX; i += blockDim.x) sh[i] = 0;
    {  // scene box = union of k_bounds' partial boxes (a few KB, L2 resident); block 0 also publishes it
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = threadIdx.x; i < n_block_boxes; i += blockDim.x)
#pragma unroll
            for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], block_boxes[6 * i + 2 * a]); hi[a] = fmaxf(hi[a], block_boxes[6 * i + 2 * a + 1]); }
#pragma unroll
        for (int a = 0; a < 3; a++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            }
            if ((threadIdx.x & 31) == 0) { sbox[threadIdx.x >> 5][2 * a] = lo[a]; sbox[threadIdx.x >> 5][2 * a + 1] = hi[a]; }
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            float v = sbox[0][threadIdx.x];
            for (int w = 1; w < 8; w++) v = (threadIdx.x & 1) ? fmaxf(v, sbox[w][threadIdx.x]) : fminf(v, sbox[w][threadIdx.x]);
            box[threadIdx.x] = v;
            if (blockIdx.x == 0) scene_box[threadIdx.x] = v;
        }
        __syncthreads();
        // the quantisation grid of the traversal nodes (srt_types.h): origin and 1 / cell, cell = largest extent / SRT_GRID_CELLS
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const float extent = fmaxf(fmaxf(box[1] - box[0], box[3] - box[2]), fmaxf(box[5] - box[4], 1e-30f));
            scene_box[8] = box[0]; scene_box[9] = box[2]; scene_box[10] = box[4]; scene_box[11] = SRT_GRID_CELLS / extent;
        }
    }
    const float b0 = box[0], b1 = box[1], b2 = box[2], b3 = box[3], b4 = box[4], b5 = box[5];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t xx = expand_bits(quant10(centroids[3ull * i], b0, b1));
        const uint32_t yy = expand_bits(quant10(centroids[3ull * i + 1], b2, b3));
        const uint32_t zz = expand_bits(quant10(centroids[3ull * i + 2], b4, b5));
        const uint32_t code = xx * 4 + yy * 2 + zz;
        codes[i] = code;
        keys[i] = code;
        vals[i] = i;
#pragma unroll
        for (int p = 0; p < SORT_PASSES; p++) atomicAdd(&sh[p * RADIX + ((code >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SORT_PASSES * RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// ---- one onesweep pass --------------------------------------------------------------------
// status word: [31:30] 0 = empty, 1 = tile-local count, 2 = inclusive prefix; [29:0] value
constexpr uint32_t LB_LOCAL = 1u << 30, LB_INCL = 2u << 30, LB_MASK = (1u << 30) - 1;

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 4 : 1) k_onesweep(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                           const uint32_t* __restrict__ digit_count, volatile uint32_t* lookback,
                                                           uint32_t* tile_counter, int rounds) {
    // rounds = keys per thread actually used (<= SORT_ITEMS): the host cuts a small sort into one equal tile per SM
    constexpr int WARPS = THREADS / 32, TILE = THREADS * SORT_ITEMS;  // THREADS >= RADIX: the per-digit steps are done by the first RADIX threads
    __shared__ uint32_t s_warp_hist[WARPS][RADIX];  // per-warp digit counts -> per-warp exclusive offsets
    __shared__ uint32_t s_tile_off[RADIX];               // exclusive offset of each digit inside the sorted tile
    __shared__ uint32_t s_glob_off[RADIX];               // global position of the tile's first key of each digit
    extern __shared__ uint32_t s_sorted[];               // [2][TILE]: the tile's keys and values in sorted order
    uint32_t* s_keys = s_sorted;
    uint32_t* s_vals = s_sorted + TILE;
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_dtot[RADIX / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);  // dynamic tile id: earlier tiles are always already running (first: its round trip overlaps the digit scan)
    // global base of every digit = exclusive scan of the pass' 256 digit counts: every block does the tiny scan itself
    uint32_t digit_base = 0;
    if (tid < RADIX) {
        const uint32_t c = digit_count[tid];
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_dtot[warp] = inc;
        digit_base = inc - c;  // + the totals of the warps before, added after the first barrier below
    }
    for (int i = tid; i < WARPS * RADIX; i += THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    if (tid < RADIX)
        for (int w = 0; w < warp; w++) digit_base += s_dtot[w];
    const uint32_t tile = s_tile;
    const uint32_t tile_keys = (uint32_t)(THREADS * rounds), tile_base = tile * tile_keys;
    // warp-striped load: warp w owns rounds * 32 consecutive keys, item i of lane l = w * rounds * 32 + i * 32 + l
    uint32_t key[SORT_ITEMS], val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t g = tile_base + warp * (rounds * 32) + i * 32 + lane;
        const bool valid = i < rounds && g < n;
        key[i] = valid ? keys_in[g] : 0xFFFFFFFFu;
        val[i] = valid ? vals_in[g] : 0u;
    }
    // stable rank of every key among equal digits of its warp (items in order, lanes in order)
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        if (i >= rounds) break;  // (uniform)
        const uint32_t g = tile_base + warp * (rounds * 32) + i * 32 + lane;
        const bool valid = g < n;
        const uint32_t digit = (key[i] >> shift) & (RADIX - 1);
#if SRT_SORT_BALLOT
        // lanes with the same digit, from one ballot per digit bit: constant time, where MATCH.ANY takes one step per distinct value
        // in the warp (~30 of 32 for the digits of Morton codes)
        uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < RADIX_BITS; b++) {
            const bool bit = (digit >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        if (!valid) peers = 1u << lane;
#else
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (RADIX + lane));
#endif
        const uint32_t before = __popc(peers & ((1u << lane) - 1));
        uint32_t base = 0;
        if (valid && before == 0) {  // leader of the peer group bumps the warp counter
            base = s_warp_hist[warp][digit];
            s_warp_hist[warp][digit] = base + __popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
        rank[i] = base + before;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan over warps (stability across warps) and the tile-local count
    uint32_t local_count = 0;
    if (tid < RADIX) {
        const int d = tid;  // one digit per thread
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            const uint32_t c = s_warp_hist[w][d];
            s_warp_hist[w][d] = run;
            run += c;
        }
        local_count = run;
        // publish the local count, then look back over earlier tiles for the exclusive prefix
        volatile uint32_t* lb = lookback + (size_t)tile * RADIX + d;
        if (tile > 0) *lb = LB_LOCAL | local_count;
        uint32_t excl = 0;
        if (tile > 0) {
            // the status words of SORT_LOOKBACK predecessors are read together (independent loads, one L2 round trip) and then
            // consumed in order: with all tiles of a pass resident at once the walk is many tiles long, and one load per
            // iteration made it a chain of L2 latencies
            constexpr int LB = THREADS == 512 ? SRT_SORT_LOOKBACK_512 : SORT_LOOKBACK;
            int t = (int)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t sw[LB];
#pragma unroll
                for (int j = 0; j < LB; j++) sw[j] = lookback[(size_t)max(t - j, 0) * RADIX + d];
#pragma unroll
                for (int j = 0; j < LB; j++) {
                    if (done || t < 0) break;   // (t < 0 cannot happen: tile 0 always publishes an inclusive prefix)
                    const uint32_t flag = sw[j] & ~LB_MASK;
                    if (flag == LB_INCL) { excl += sw[j] & LB_MASK; done = true; }
                    else if (flag == LB_LOCAL) { excl += sw[j] & LB_MASK; t--; }
                    else break;  // not published yet: poll again from this tile
                }
            }
        }
        __threadfence();
        *lb = LB_INCL | (excl + local_count);
        s_glob_off[d] = digit_base + excl;
    }
    // exclusive scan of the tile-local digit counts -> position of each digit inside the sorted tile
    {
        uint32_t inc = local_count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __shared__ uint32_t s_wtot[RADIX / 32];
        if (lane == 31 && tid < RADIX) s_wtot[warp] = inc;
        __syncthreads();
        if (tid < RADIX) {
            uint32_t base = 0;
            for (int w = 0; w < warp; w++) base += s_wtot[w];
            s_tile_off[tid] = base + inc - local_count;
        }
    }
    __syncthreads();
    // scatter into shared memory in sorted order
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t g = tile_base + warp * (rounds * 32) + i * 32 + lane;
        if (i < rounds && g < n) {
            const uint32_t digit = (key[i] >> shift) & (RADIX - 1);
            const uint32_t pos = s_tile_off[digit] + s_warp_hist[warp][digit] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
    // coalesced write-out: consecutive threads write consecutive addresses inside each digit run
    const uint32_t count = min(tile_keys, n - tile_base);
    for (uint32_t j = tid; j < count; j += THREADS) {
        const uint32_t k = s_keys[j];
        const uint32_t digit = (k >> shift) & (RADIX - 1);
        const uint32_t dst = s_glob_off[digit] + (j - s_tile_off[digit]);
        keys_out[dst] = k;
        vals_out[dst] = s_vals[j];
    }
}

// the sorted tile lives in dynamic shared memory: above 48 KB a kernel has to opt in (once per device)
static bool sort_configure() {
    static std::atomic<uint64_t> done_mask{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    const uint64_t bit = 1ull << (dev & 63);
    if (done_mask.load() & bit) return true;
    if (cudaFuncSetAttribute(k_onesweep<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem(512)) != cudaSuccess) return false;
    done_mask.fetch_or(bit);
    return true;
}
static int sm_count();
// Sorts whose 8192-key tiles all fit the machine at once (one 512-thread block per SM: up to 1.2 M keys on 148 SMs) use them: half
// the tiles, half the look-back, one block's fixed costs per SM -- 1M keys: 24.5 -> 22.5 us per pass; 10M keys are faster on the
// 4096-key tiles (0.51 vs 0.60 ms: three resident blocks per SM instead of one).
static int sort_threads_for(uint32_t n) { return (n + 512 * SORT_ITEMS - 1) / (512 * SORT_ITEMS) <= (uint32_t)sm_count() ? 512 : 256; }
// keys per thread: SORT_ITEMS, or -- small sorts -- as few as give every SM one tile (1M keys: 14 x 512 = 7168 keys, 147 tiles)
static int sort_rounds_for(uint32_t n) {
    if (sort_threads_for(n) != 512) return SORT_ITEMS;
    const uint32_t per_sm = (n + (uint32_t)sm_count() - 1) / (uint32_t)sm_count();
    return (int)std::min<uint32_t>(SORT_ITEMS, std::max<uint32_t>(1u, (per_sm + 511u) / 512u));
}
static uint32_t sort_tiles_for(uint32_t n) { const uint32_t tile = (uint32_t)(sort_threads_for(n) * sort_rounds_for(n)); return (n + tile - 1) / tile; }
static void launch_onesweep(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, uint32_t n, int shift,
                            const uint32_t* digit_count, uint32_t* lookback, uint32_t* tile_counter, cudaStream_t st) {
    const int rounds = sort_rounds_for(n);
    if (sort_threads_for(n) == 512) k_onesweep<512><<<sort_tiles_for(n), 512, sort_smem(512), st>>>(keys_in, vals_in, keys_out, vals_out, n, shift, digit_count, lookback, tile_counter, rounds);
    else k_onesweep<256><<<sort_tiles_for(n), 256, sort_smem(256), st>>>(keys_in, vals_in, keys_out, vals_out, n, shift, digit_count, lookback, tile_counter, rounds);
}

// ---- pixel order of the wavefront renderer (renderer.cu) -------------------------------------
// Between two rounds of samples the renderer re-orders the pixel slots of a chunk so that the pixels
// with the most passes per sample so far are handed out first (longest processing time first): the
// end of the launch then consists of the cheapest pixels and the persistent blocks drain together.
// One stable onesweep pass over an 8-bit key: 254 - min(254, 23 * passes / samples) for rendered
// slots (1 pass per sample -> 231, 11 -> 1), 255 for slots nobody rendered (tile overhang; they sort
// behind the n_order owned slots and are never read).  Stable, so equal-cost neighbours stay neighbours.
__global__ void __launch_bounds__(256) k_cost_keys(const uint32_t* __restrict__ cost, uint32_t n, uint32_t samples, uint32_t* __restrict__ keys,
                                                   uint32_t* __restrict__ vals, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RADIX];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = cost[i];
        const uint32_t key = c ? 254u - min(254u, (uint32_t)(((unsigned long long)c * 23u) / samples)) : 255u;
        keys[i] = key;
        vals[i] = i;
        atomicAdd(&sh[key], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

struct PixelOrder {
    uint32_t cap = 0, tiles = 0;
    uint32_t *keys[2] = {nullptr, nullptr}, *vals[2] = {nullptr, nullptr};
    uint32_t* small = nullptr;  // hist[RADIX] | tile ticket | lookback[tiles * RADIX]
};
PixelOrder* pixel_order_create(uint32_t max_slots) {
    auto* o = new PixelOrder();
    o->cap = max_slots ? max_slots : 1;
    o->tiles = std::max<uint32_t>((o->cap + SORT_TILE - 1) / SORT_TILE, (uint32_t)sm_count() + 1u);
    bool ok = true;
    for (int k = 0; k < 2; k++) ok = ok && device_pool_alloc((void**)&o->keys[k], o->cap * sizeof(uint32_t)) && device_pool_alloc((void**)&o->vals[k], o->cap * sizeof(uint32_t));
    ok = ok && device_pool_alloc((void**)&o->small, ((size_t)RADIX + 64 + (size_t)o->tiles * RADIX) * sizeof(uint32_t));
    if (!ok) { pixel_order_destroy(o); return nullptr; }
    return o;
}
void pixel_order_destroy(PixelOrder* o) {
    if (!o) return;
    for (int k = 0; k < 2; k++) { device_pool_free(o->keys[k]); device_pool_free(o->vals[k]); }
    device_pool_free(o->small);
    delete o;
}
// enqueues the sort on `st`; the returned device array (n entries, owned slots first) is valid until the next call
const uint32_t* pixel_order_build(PixelOrder* o, const uint32_t* cost, uint32_t n, uint32_t samples, cudaStream_t st) {
    if (!o || n == 0 || n > o->cap || samples == 0) return nullptr;
    const uint32_t tiles = std::max<uint32_t>((n + SORT_TILE - 1) / SORT_TILE, sort_tiles_for(n));  // status words to clear
    uint32_t* hist = o->small;
    uint32_t* ticket = o->small + RADIX;
    uint32_t* lookback = o->small + RADIX + 64;
    if (cudaMemsetAsync(o->small, 0, ((size_t)RADIX + 64 + (size_t)tiles * RADIX) * sizeof(uint32_t), st) != cudaSuccess) return nullptr;
    const int grid = (int)min((uint32_t)(148 * 8), (n + 255) / 256);
    k_cost_keys<<<grid, 256, 0, st>>>(cost, n, samples, o->keys[0], o->vals[0], hist);
    if (!sort_configure()) return nullptr;
    launch_onesweep(o->keys[0], o->vals[0], o->keys[1], o->vals[1], n, 0, hist, lookback, ticket, st);
    count_launch(2);
    return o->vals[1];
}

// ---- hierarchy + refit + emit, bottom-up ------------------------------------------------------------
// The binary radix tree over the sorted (code, index) keys is unique, so it can be grown from the leaves (Apetrei 2014) instead of
// searched top-down per internal node (Karras 2012, what oracle/lbvh_oracle.c restates): a finished subtree covering leaves [a, b] is
// the LEFT child of the split after b when the keys differ less across that split than across the one before a (longer common prefix =
// nearer ancestor), else the RIGHT child of the split before a.  The first subtree to arrive at a split leaves its far end there and
// stops; the second one takes it, owns both child boxes and goes on as the merged subtree.  Node NUMBERS are Karras': a left child is
// numbered by the last leaf of its range, a right child by the first, the root is 0 -- so a node learns its number at the moment it
// learns which child it is, and its children are `split` (or leaf n-1+split) and `split+1` (or leaf n-1+split+1), exactly the oracle's
// left[] / right[].  Boxes travel as two 16-byte vectors (lo.xyz | epoch, hi.xyz | epoch) stored under the node's number.
// a child box on the 16-bit scene grid, SRT_GRID_MARGIN cells of margin (srt_types.h); min in the low half, max in the high half
__device__ __forceinline__ uint32_t grid_pack(float mn, float mx, float lo, float inv_cell) {
    const float gl = (mn - lo) * inv_cell + SRT_GRID_OFFSET, gh = (mx - lo) * inv_cell + SRT_GRID_OFFSET;
    const int ql = max(__float2int_rd(gl) - SRT_GRID_MARGIN, 0), qh = min(__float2int_ru(gh) + SRT_GRID_MARGIN, 65535);  // floor / ceil in the conversion
    return (uint32_t)ql | ((uint32_t)qh << 16);
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// how much the augmented keys (code, position) differ across the split between sorted positions i and i + 1: smaller = longer common prefix
__device__ __forceinline__ unsigned long long split_delta(const uint32_t* __restrict__ keys, int i) {
    return ((unsigned long long)(__ldg(keys + i) ^ __ldg(keys + i + 1)) << 32) | (uint32_t)(i ^ (i + 1));
}
__global__ void __launch_bounds__(256) k_build_tree(int n, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ sorted_idx,
                                                    const float4* __restrict__ leaf_boxes, int32_t* other, float4* node_box, int2* __restrict__ children,
                                                    uint32_t epoch_bits) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t src = sorted_idx[k];
    float4 lo = leaf_boxes[2ull * src], hi = leaf_boxes[2ull * src + 1];
    const float epoch = __uint_as_float(epoch_bits);  // only ever compared as bits
    lo.w = hi.w = epoch;
    node_box[2 * (n - 1 + k)] = lo;
    node_box[2 * (n - 1 + k) + 1] = hi;
    int a = k, b = k, split = -1;          // the subtree this thread carries: leaves [a, b], split between its children (-1: a leaf)
    while (true) {
        const bool root = a == 0 && b == n - 1;
        const bool is_left = root ? false : (a == 0 ? true : (b == n - 1 ? false : split_delta(keys, b) < split_delta(keys, a - 1)));
        if (split >= 0) {  // an internal node: now that its number is known, record its children and publish its box
            const int self = root ? 0 : (is_left ? b : a);
            const int L = a == split ? n - 1 + split : split, R = b == split + 1 ? n - 1 + split + 1 : split + 1;
            children[self] = make_int2(L >= n - 1 ? ~(L - (n - 1)) : L, R >= n - 1 ? ~(R - (n - 1)) : R);
            node_box[2 * self] = lo;
            node_box[2 * self + 1] = hi;
        }
        if (root) return;
        const int p = is_left ? b : a - 1;  // the parent's split
        const int far = atomicExch(other + p, is_left ? a : b);
        if (far < 0) return;  // the sibling subtree finishes this node
        // No fence anywhere: a published box carries this build's epoch in the fourth lane of both of its 16-byte halves (a 16-byte
        // store lands as one piece), so the finisher simply re-reads the sibling's halves until both show the epoch -- its exchange
        // may have overtaken the sibling's stores, which were issued before the sibling's own.
        const int sib = is_left ? (far == b + 1 ? n - 1 + far : b + 1) : (far == a - 1 ? n - 1 + far : a - 1);
        float4 olo, ohi;
        do {
            olo = ld_volatile_f4(node_box + 2 * sib);
            ohi = ld_volatile_f4(node_box + 2 * sib + 1);
        } while (__float_as_uint(olo.w) != epoch_bits || __float_as_uint(ohi.w) != epoch_bits);
        if (is_left) b = far;
        else a = far;
        split = p;
        lo = make_float4(fminf(lo.x, olo.x), fminf(lo.y, olo.y), fminf(lo.z, olo.z), epoch);
        hi = make_float4(fmaxf(hi.x, ohi.x), fmaxf(hi.y, ohi.y), fmaxf(hi.z, ohi.z), epoch);
    }
}
// The 4-wide traversal nodes (srt_types.h), one thread per node: the slots of node i are its leaf children and the children of its
// internal children.  The grid quantisation happens HERE, with all 32 lanes at work, not in the finisher of k_build_tree, where 3 - 6
// lanes of a warp are left (the tree kernel: 0.111 -> 0.09 ms at 1M triangles).  A left child is numbered `split`, a right child
// `split + 1`, so the two children of a node -- refs and boxes -- are neighbours in memory.
__device__ __forceinline__ uint4 wide_slot(const float4* __restrict__ node_box, int n, int ref, float4 G) {
    const int number = ref >= 0 ? ref : n - 1 + ~ref;
    const float4 lo = __ldg(node_box + 2 * number), hi = __ldg(node_box + 2 * number + 1);
    return make_uint4(grid_pack(lo.x, hi.x, G.x, G.w), grid_pack(lo.y, hi.y, G.y, G.w), grid_pack(lo.z, hi.z, G.z, G.w), (uint32_t)ref);
}
__global__ void __launch_bounds__(256) k_collapse4(int n, const int2* __restrict__ children, const float4* __restrict__ node_box,
                                                   const float4* __restrict__ grid, SrtWide* __restrict__ wide) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const float4 G = __ldg(grid);
    const int2 c = __ldg(children + i);
    // the children's own children (node 0 stands in for a leaf child: no dependent branch before the loads are out)
    const int2 g0 = __ldg(children + (c.x >= 0 ? c.x : 0)), g1 = __ldg(children + (c.y >= 0 ? c.y : 0));
    const bool in0 = c.x >= 0, in1 = c.y >= 0;
    const int r0 = in0 ? g0.x : c.x, r1 = in0 ? g0.y : (in1 ? g1.x : c.y), r2 = in0 ? (in1 ? g1.x : c.y) : (in1 ? g1.y : SRT_WIDE_EMPTY),
              r3 = (in0 && in1) ? g1.y : SRT_WIDE_EMPTY;
    const uint4 none = make_uint4(0u, 0u, 0u, (uint32_t)SRT_WIDE_EMPTY);
    const uint4 s0 = wide_slot(node_box, n, r0, G), s1 = wide_slot(node_box, n, r1, G);
    const uint4 s2 = r2 != SRT_WIDE_EMPTY ? wide_slot(node_box, n, r2, G) : none, s3 = r3 != SRT_WIDE_EMPTY ? wide_slot(node_box, n, r3, G) : none;
    uint4* wp = reinterpret_cast<uint4*>(wide + i);
    wp[0] = make_uint4(s0.x, s0.y, s0.z, s1.x);
    wp[1] = make_uint4(s1.y, s1.z, s2.x, s2.y);
    wp[2] = make_uint4(s2.z, s3.x, s3.y, s3.z);
    wp[3] = make_uint4(s0.w, s1.w, s2.w, s3.w);
}
// left / right / parent in the oracle's numbering (internal i, leaf n-1+k), read off the emitted nodes: only dumps need them
__global__ void __launch_bounds__(256) k_topology(int n, const int2* __restrict__ children, int32_t* __restrict__ left, int32_t* __restrict__ right,
                                                  int32_t* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) parent[0] = -1;
    if (i >= n - 1) return;
    const int c0 = children[i].x, c1 = children[i].y;
    const int L = c0 < 0 ? n - 1 + ~c0 : c0, R = c1 < 0 ? n - 1 + ~c1 : c1;
    left[i] = L;
    right[i] = R;
    parent[L] = i;
    parent[R] = i;
}
// triangles to leaf order: three lanes move one 48-byte triangle, 16 bytes each, so the stores are fully coalesced.
// A pure gather that needs nothing but the sorted order: it runs on a second stream next to hierarchy + refit.
__global__ void __launch_bounds__(256) k_permute_tris(uint32_t n, const uint32_t* __restrict__ sorted_idx, const float4* __restrict__ tris_in,
                                                      float4* __restrict__ tris) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3u * n) return;
    const uint32_t k = t / 3u, part = t - 3u * k;
    tris[t] = __ldg(tris_in + 3ull * sorted_idx[k] + part);
}

// ------------------------------------------------------------------------------------------ host
template <class T> static bool dalloc(T*& p, size_t count) {  // cached: creating a scene per frame must not pay ~25 cudaMallocs
    return device_pool_alloc((void**)&p, (count ? count : 1) * sizeof(T));
}
static void dfree(void* p) { device_pool_free(p); }

DeviceScene* device_scene_create(const std::vector<HostTri>& tris, const std::vector<HostMaterial>& mats, double origin_l1_bound) {
    if (cuda_device_count() == 0) { set_error("libsrt: no CUDA device available (the product has no CPU fallback)"); return nullptr; }
    auto* s = new DeviceScene();
    s->n = (uint32_t)tris.size();
    s->n_mats = (uint32_t)mats.size();
    const uint32_t n = s->n;
    std::vector<float> verts(9ull * n);
    std::vector<SrtTri> packed(n);
    std::vector<SrtMaterial> dm(mats.size());
    for (size_t m = 0; m < mats.size(); m++) {
        memcpy(dm[m].spec, mats[m].spec, sizeof dm[m].spec);
        dm[m].fuzz = mats[m].fuzz;
        for (int k = 0; k < 3; k++) { dm[m].sellB[k] = mats[m].B[k]; dm[m].sellC[k] = mats[m].C[k]; }
        dm[m].type = mats[m].type;
    }
    const std::vector<uint32_t> prio = reference_test_order(tris);
    for (int a = 0; a < 3; a++) { s->host_lo[a] = n ? INFINITY : 0.f; s->host_hi[a] = n ? -INFINITY : 0.f; }
    for (const HostTri& t : tris)
        for (int a = 0; a < 3; a++) { s->host_lo[a] = std::min(s->host_lo[a], t.bbox[2 * a]); s->host_hi[a] = std::max(s->host_hi[a], t.bbox[2 * a + 1]); }
    for (uint32_t i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) { verts[9ull * i + 3 * k] = tris[i].v[k].x; verts[9ull * i + 3 * k + 1] = tris[i].v[k].y; verts[9ull * i + 3 * k + 2] = tris[i].v[k].z; }
        const uint32_t mt = tris[i].mat < mats.size() ? mats[tris[i].mat].type : SRT_LAMBERTIAN;
        packed[i] = tris[i].pack(mt, prio[i]);
    }
    s->tiles = std::max<uint32_t>((n + SORT_TILE - 1) / SORT_TILE, (uint32_t)sm_count() + 1u);  // look-back words: small sorts run one (smaller) tile per SM
    bool ok = dalloc(s->verts, 9ull * n) && dalloc(s->tris_in, n) && dalloc(s->mats, mats.size()) && dalloc(s->leaf_boxes, 2ull * n) &&
              dalloc(s->centroids, 3ull * n) && dalloc(s->scene_box, 12) && dalloc(s->codes, n) && dalloc(s->keys[0], n) && dalloc(s->keys[1], n) &&
              dalloc(s->vals[0], n) && dalloc(s->vals[1], n) && dalloc(s->hist, SORT_PASSES * RADIX) &&
              dalloc(s->lookback, (size_t)SORT_PASSES * (s->tiles ? s->tiles : 1) * RADIX) && dalloc(s->tile_counter, SORT_PASSES) &&
              dalloc(s->left, n) && dalloc(s->right, n) && dalloc(s->parent, 2ull * n) && dalloc(s->block_boxes, 6 * kBoundsBlocks) && dalloc(s->node_box, 4ull * n) && dalloc(s->visit, n) &&
              dalloc(s->children, n) && dalloc(s->wide, n) && dalloc(s->tris, n);
    for (auto& e : s->ev) ok = ok && cuda_ok(cudaEventCreate(&e), "cudaEventCreate", __FILE__, __LINE__);
    ok = ok && cuda_ok(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking), "cudaStreamCreate", __FILE__, __LINE__) &&
         cuda_ok(cudaEventCreateWithFlags(&s->ev_sorted, cudaEventDisableTiming), "cudaEventCreate", __FILE__, __LINE__) &&
         cuda_ok(cudaEventCreateWithFlags(&s->ev_side, cudaEventDisableTiming), "cudaEventCreate", __FILE__, __LINE__);
    // pooled blocks carry whatever their last owner wrote: the refit tells published boxes by an epoch in their fourth lane
    if (ok) ok = cuda_ok(cudaMemset(s->node_box, 0, 4ull * std::max(n, 1u) * sizeof(float4)), "clear node boxes", __FILE__, __LINE__);
    if (ok && n) {
        ok = cuda_ok(cudaMemcpy(s->verts, verts.data(), verts.size() * sizeof(float), cudaMemcpyHostToDevice), "upload verts", __FILE__, __LINE__) &&
             cuda_ok(cudaMemcpy(s->tris_in, packed.data(), packed.size() * sizeof(SrtTri), cudaMemcpyHostToDevice), "upload tris", __FILE__, __LINE__);
    }
    if (ok && !mats.empty())
        ok = cuda_ok(cudaMemcpy(s->mats, dm.data(), dm.size() * sizeof(SrtMaterial), cudaMemcpyHostToDevice), "upload mats", __FILE__, __LINE__);
    FlatLeaf flat;
    s->origin_l1_bound = origin_l1_bound;
    if (ok && build_flat_leaf(tris, mats, prio, origin_l1_bound, flat)) {
        s->n_units = (uint32_t)flat.units.size();
        s->flat_guard = flat.guard; s->flat_tol = flat.tol;
        ok = dalloc(s->flat_units, flat.units.size()) && dalloc(s->flat_tris, flat.tris.size()) && dalloc(s->flat_to_orig, flat.to_orig.size()) &&
             cuda_ok(cudaMemcpy(s->flat_to_orig, flat.to_orig.data(), flat.to_orig.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "upload flat order", __FILE__, __LINE__) &&
             cuda_ok(cudaMemcpy(s->flat_units, flat.units.data(), flat.units.size() * sizeof(SrtFlatUnit), cudaMemcpyHostToDevice), "upload units", __FILE__, __LINE__) &&
             cuda_ok(cudaMemcpy(s->flat_tris, flat.tris.data(), flat.tris.size() * sizeof(SrtTri), cudaMemcpyHostToDevice), "upload flat tris", __FILE__, __LINE__);
    }
    float ms[5];
    if (ok) ok = device_scene_build_lbvh(s, 1, ms);
    if (!ok) { device_scene_destroy(s); return nullptr; }
    return s;
}

void device_scene_destroy(DeviceScene* s) {
    if (!s) return;
    dfree(s->verts); dfree(s->tris_in); dfree(s->mats); dfree(s->leaf_boxes); dfree(s->centroids); dfree(s->scene_box);
    dfree(s->codes); dfree(s->keys[0]); dfree(s->keys[1]); dfree(s->vals[0]); dfree(s->vals[1]); dfree(s->hist);
    dfree(s->block_boxes);
    if (s->side) cudaStreamDestroy(s->side);
    if (s->ev_sorted) cudaEventDestroy(s->ev_sorted);
    if (s->ev_side) cudaEventDestroy(s->ev_side);
    dfree(s->lookback); dfree(s->tile_counter); dfree(s->left); dfree(s->right); dfree(s->parent); dfree(s->node_box);
    dfree(s->visit); dfree(s->children); dfree(s->wide); dfree(s->tris); dfree(s->flat_units); dfree(s->flat_tris); dfree(s->flat_to_orig);
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    delete s;
}

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

bool device_scene_build_lbvh(DeviceScene* s, int repeats, float ms_out[5]) {
    const uint32_t n = s->n;
    for (int k = 0; k < 5; k++) ms_out[k] = 0.f;
    if (n == 0) return true;
    const int sms = sm_count();
    // every block of k_morton_hist flushes up to 1024 histogram counters with global atomics: 4 blocks per SM, not 8 (-6 us at 1M keys; 2 lose at 10M)
#ifndef SRT_MORTON_BLOCKS_PER_SM
#define SRT_MORTON_BLOCKS_PER_SM 4
#endif
    const int grid_stride = (int)min((uint32_t)(sms * SRT_MORTON_BLOCKS_PER_SM), (n + 255) / 256);
    const int grid_n = (int)((n + 255) / 256);
    cudaStream_t st = s->stream;
    // every launch of one build.  (Replaying the sequence as a CUDA graph was measured: 0.362 -> 0.359 ms at 1M triangles,
    // the launches already queue back to back, so the plain stream stays.)
    auto enqueue = [&]() -> bool {
        const uint32_t n_lookback = (uint32_t)((size_t)SORT_PASSES * s->tiles * RADIX);
        k_bounds<<<kBoundsBlocks, 256, 0, st>>>(s->verts, n, s->leaf_boxes, s->centroids, s->block_boxes, s->visit, s->lookback, n_lookback, s->hist,
                                                SORT_PASSES * RADIX, s->tile_counter, SORT_PASSES);
        k_morton_hist<<<grid_stride, 256, 0, st>>>(s->centroids, s->block_boxes, kBoundsBlocks, s->scene_box, n, s->codes, s->keys[0], s->vals[0], s->hist);
        SRT_CUDA(cudaEventRecord(s->ev[1], st));
        if (!sort_configure()) { set_error("k_onesweep: shared memory opt-in failed"); return false; }
        for (int p = 0; p < SORT_PASSES; p++) {
            launch_onesweep(s->keys[p & 1], s->vals[p & 1], s->keys[(p + 1) & 1], s->vals[(p + 1) & 1], n, p * RADIX_BITS, s->hist + p * RADIX,
                            s->lookback + (size_t)p * s->tiles * RADIX, s->tile_counter + p, st);
        }
        SRT_CUDA(cudaEventRecord(s->ev[2], st));
        // the triangle permutation only needs the sorted order: second stream, next to hierarchy + refit
        SRT_CUDA(cudaEventRecord(s->ev_sorted, st));
        SRT_CUDA(cudaStreamWaitEvent(s->side, s->ev_sorted, 0));
        k_permute_tris<<<(3 * n + 255) / 256, 256, 0, s->side>>>(n, s->vals[0], reinterpret_cast<const float4*>(s->tris_in), reinterpret_cast<float4*>(s->tris));
        SRT_CUDA(cudaEventRecord(s->ev_side, s->side));
        // process-wide counter: an epoch is never used twice, and the box arrays were zeroed when the scene was created (epoch 0 is never handed out)
        uint32_t epoch = g_refit_epoch.fetch_add(1u) + 1u;
        if (epoch == 0) epoch = g_refit_epoch.fetch_add(1u) + 1u;
        k_build_tree<<<grid_n, 256, 0, st>>>((int)n, s->keys[0], s->vals[0], s->leaf_boxes, reinterpret_cast<int32_t*>(s->visit), s->node_box, s->children, epoch);
        SRT_CUDA(cudaEventRecord(s->ev[3], st));  // ms_out[3] = hierarchy + refit, ms_out[4] = the 4-wide traversal nodes (+ what is left of the permutation)
        if (n > 1) k_collapse4<<<(n - 1 + 255) / 256, 256, 0, st>>>((int)n, s->children, s->node_box, reinterpret_cast<const float4*>(s->scene_box + 8), s->wide);
        SRT_CUDA(cudaStreamWaitEvent(st, s->ev_side, 0));
        return true;
    };
    for (int rep = 0; rep < repeats; rep++) {
        SRT_CUDA(cudaEventRecord(s->ev[0], st));
        if (!enqueue()) return false;
        SRT_CUDA(cudaEventRecord(s->ev[4], st));
        count_launch(5 + SORT_PASSES);  // bounds, morton, passes, permute, tree, collapse
        SRT_CUDA_LAST();
    }
    SRT_CUDA(cudaEventSynchronize(s->ev[4]));
    float t;
    SRT_CUDA(cudaEventElapsedTime(&t, s->ev[0], s->ev[4])); ms_out[0] = t;
    SRT_CUDA(cudaEventElapsedTime(&t, s->ev[0], s->ev[1])); ms_out[1] = t;
    SRT_CUDA(cudaEventElapsedTime(&t, s->ev[1], s->ev[2])); ms_out[2] = t;
    SRT_CUDA(cudaEventElapsedTime(&t, s->ev[2], s->ev[3])); ms_out[3] = t;
    SRT_CUDA(cudaEventElapsedTime(&t, s->ev[3], s->ev[4])); ms_out[4] = t;
    s->last_build_ms = ms_out[0];
    return true;
}
double device_scene_lbvh_ms(const DeviceScene* s) { return s->last_build_ms; }

bool device_scene_download_lbvh(const DeviceScene* s, LbvhDump& o) {
    const uint32_t n = s->n;
    o.codes.resize(n); o.sorted_idx.resize(n); o.left.resize(n > 0 ? n - 1 : 0); o.right.resize(n > 0 ? n - 1 : 0);
    o.parent.resize(n ? 2 * n - 1 : 0); o.node_boxes.resize(n ? 6ull * (2 * n - 1) : 0);
    if (!n) return true;
    if (n > 1) k_topology<<<(n + 255) / 256, 256, 0, s->stream>>>((int)n, s->children, s->left, s->right, s->parent);
    else SRT_CUDA(cudaMemsetAsync(s->parent, 0xFF, sizeof(int32_t), s->stream));
    count_launch();
    SRT_CUDA(cudaStreamSynchronize(s->stream));
    SRT_CUDA(cudaMemcpy(o.codes.data(), s->codes, n * 4, cudaMemcpyDeviceToHost));
    SRT_CUDA(cudaMemcpy(o.sorted_idx.data(), s->vals[0], n * 4, cudaMemcpyDeviceToHost));
    if (n > 1) {
        SRT_CUDA(cudaMemcpy(o.left.data(), s->left, (n - 1) * 4, cudaMemcpyDeviceToHost));
        SRT_CUDA(cudaMemcpy(o.right.data(), s->right, (n - 1) * 4, cudaMemcpyDeviceToHost));
    }
    SRT_CUDA(cudaMemcpy(o.parent.data(), s->parent, (2 * n - 1) * 4, cudaMemcpyDeviceToHost));
    {
        std::vector<float4> box(2 * (2 * (size_t)n - 1));
        SRT_CUDA(cudaMemcpy(box.data(), s->node_box, box.size() * sizeof(float4), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < box.size() / 2; i++) {
            float* b = o.node_boxes.data() + 6 * i;
            const float4 lo = box[2 * i], hi = box[2 * i + 1];
            b[0] = lo.x; b[1] = hi.x; b[2] = lo.y; b[3] = hi.y; b[4] = lo.z; b[5] = hi.z;
        }
    }
    SRT_CUDA(cudaMemcpy(o.scene_box, s->scene_box, 24, cudaMemcpyDeviceToHost));
    return true;
}

// accessors for the renderer translation units
const SrtWide* device_scene_nodes(const DeviceScene* s) { return s->wide; }
const float4* device_scene_grid(const DeviceScene* s) { return reinterpret_cast<const float4*>(s->scene_box + 8); }
const SrtTri* device_scene_tris(const DeviceScene* s) { return s->tris; }
const SrtFlatUnit* device_scene_flat_units(const DeviceScene* s) { return s->flat_units; }
const SrtTri* device_scene_flat_tris(const DeviceScene* s) { return s->flat_tris; }
uint32_t device_scene_n_units(const DeviceScene* s) { return s->n_units; }
uint32_t device_scene_num_units(const DeviceScene* s) { return s->n_units; }
const uint32_t* device_scene_flat_to_orig(const DeviceScene* s) { return s->flat_to_orig; }
void device_scene_flat_guard(const DeviceScene* s, float* guard, float* tol) { *guard = s->flat_guard; *tol = s->flat_tol; }
double device_scene_origin_bound(const DeviceScene* s) { return s->origin_l1_bound; }
void device_scene_bounds(const DeviceScene* s, float lo[3], float hi[3]) { for (int a = 0; a < 3; a++) { lo[a] = s->host_lo[a]; hi[a] = s->host_hi[a]; } }
const SrtMaterial* device_scene_mats(const DeviceScene* s) { return s->mats; }
uint32_t device_scene_ntris(const DeviceScene* s) { return s->n; }
uint32_t device_scene_nmats(const DeviceScene* s) { return s->n_mats; }
const uint32_t* device_scene_sorted_idx(const DeviceScene* s) { return s->vals[0]; }

}  // namespace srt
