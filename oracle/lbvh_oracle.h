/* TEST INFRASTRUCTURE ONLY (oracle/).
 * CPU restatement of the LBVH specified in SURVEY.md 8a-L (Karras 2012, "Maximizing
 * Parallelism in the Construction of BVHs, Octrees, and k-d Trees" / "Thinking Parallel III",
 * which the reference only cites: README.md:15).  PARITY UNPINNED against the reference: its
 * own BVH is a serial random-axis median split (bvh/bvh.cu:206-309) -- there is no LBVH to
 * compare with, so this file IS the specification the CUDA build must match bit for bit.
 *
 * Node ids: internal nodes 0..n-2 (0 = root), leaf k (k-th entry of the sorted order) = n-1+k.
 */
#ifndef SRT_LBVH_ORACLE_H
#define SRT_LBVH_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* leaf_boxes: n x 6 floats (xmin xmax ymin ymax zmin zmax, already padded), centroids: n x 3.
 * outputs (caller allocated): scene_box[6], codes[n] (in ORIGINAL triangle order),
 * sorted_idx[n], left[n-1], right[n-1], parent[2n-1], node_boxes[(2n-1)*6]. */
void srt_oracle_lbvh_build(int n, const float* leaf_boxes, const float* centroids, float* scene_box,
                           uint32_t* codes, uint32_t* sorted_idx, int32_t* left, int32_t* right,
                           int32_t* parent, float* node_boxes);
uint32_t srt_oracle_morton30(float cx, float cy, float cz, const float scene_box[6]);
/* centroid of a triangle as the reference computes it: (v0+v1+v2)/3.f, primitives/tri.cuh:73-77 */
void srt_oracle_tri_centroid(const float v9[9], float c[3]);
#ifdef __cplusplus
}
#endif
#endif
