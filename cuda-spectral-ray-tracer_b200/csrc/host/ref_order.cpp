// Tie-break priority of triangles: the order in which the REFERENCE's traversal tests them.
//
// bvh::hit (bvh/bvh.cu:98-166) accepts a hit when t <= closest_so_far, so among triangles hit at
// exactly the same t (coplanar faces: a box standing on the floor) the one it tests LAST wins, and
// its test order is fixed by its tree: internal nodes in pre-order (left subtree first, right child
// stacked), at every node the left child then the right child, leaves tested on the spot.  To give
// the same answer we rebuild that tree's shape on the host -- the serial top-down median split of
// bvh/bvh.cu:206-309 with its XORWOW(1984) axis stream (scene/scene.cu:9-20, cuda_utility.cu:43-48)
// and its Lomuto quicksort (bvh.cu:14-71) -- and hand every triangle its rank in that order.
// Only the ORDER is used; no box of that tree is ever tested.
#include "srt_host.hpp"
#include <cmath>
#include <numeric>

namespace srt {
namespace {
struct Xorwow {
    uint32_t d, v[5];
    explicit Xorwow(uint32_t seed) {
        const uint32_t s0 = seed ^ 0xaad26b49u, s1 = 0xf7dcefddu;
        const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
        d = 6615241u + t1 + t0;
        v[0] = 123456789u + t0; v[1] = 362436069u ^ t0; v[2] = 521288629u + t1; v[3] = 88675123u ^ t1; v[4] = 5783321u + t0;
    }
    float uniform() {
        const uint32_t t = v[0] ^ (v[0] >> 2);
        v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
        v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
        d += 362437u;
        return (float)(v[4] + d) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
    }
    int axis() {  // cuda_random_int(0, 2): ceil(u * 2 - 1) in {0, 1}
        const float f = uniform() * (1.0f - (-1.0f)) + (-1.0f);
        return (int)std::ceil(f);
    }
};
struct Shape { int left = -1, right = -1, prim = -1; bool leaf = false; };
}  // namespace

std::vector<uint32_t> reference_test_order(const std::vector<HostTri>& tris) {
    const int n = (int)tris.size();
    std::vector<uint32_t> prio(n);
    std::iota(prio.begin(), prio.end(), 0u);
    if (n < 2 || n > 4096) return prio;  // the reference cannot build larger scenes (serial build, depth-64 stack)
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::vector<Shape> nodes;
    nodes.reserve(2 * n);
    auto less = [&](int a, int b, int axis) { return tris[a].bbox[2 * axis] < tris[b].bbox[2 * axis]; };
    auto quicksort = [&](int start, int end, int axis) {  // bvh.cu:33-71 (iterative, Lomuto partition, last element as pivot)
        std::vector<int> st;
        st.push_back(start); st.push_back(end);
        while (!st.empty()) {
            const int h = st.back(); st.pop_back();
            const int l = st.back(); st.pop_back();
            int p = l;
            if (l != h) {
                const int x = order[h];
                int i = l - 1;
                for (int j = l; j < h; j++)
                    if (less(order[j], x, axis)) { i++; std::swap(order[i], order[j]); }
                std::swap(order[i + 1], order[h]);
                p = i + 1;
            }
            if (p - 1 > l) { st.push_back(l); st.push_back(p - 1); }
            if (p + 1 < h) { st.push_back(p + 1); st.push_back(h); }
        }
    };
    Xorwow rng(1984u);
    struct Item { size_t s, e; int node; };
    std::vector<Item> stack;
    nodes.emplace_back();
    stack.push_back({0, (size_t)n, 0});
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const size_t span = it.e - it.s;
        if (span == 0) continue;
        if (span == 1) { nodes[it.node].leaf = true; nodes[it.node].prim = order[it.s]; continue; }
        const int axis = rng.axis();
        if (span == 2) {
            const int a = order[it.s], b = order[it.s + 1];
            const bool ab = less(a, b, axis);
            Shape l, r;
            l.leaf = r.leaf = true;
            l.prim = ab ? a : b;
            r.prim = ab ? b : a;
            nodes[it.node].left = (int)nodes.size(); nodes.push_back(l);
            nodes[it.node].right = (int)nodes.size(); nodes.push_back(r);
        } else {
            quicksort((int)it.s, (int)it.e - 1, axis);
            const size_t mid = it.s + span / 2;
            const int l = (int)nodes.size(); nodes.emplace_back();
            const int r = (int)nodes.size(); nodes.emplace_back();
            nodes[it.node].left = l; nodes[it.node].right = r;
            if (stack.size() + 2 > 64) return prio;  // the reference would have failed here (MAX_DEPTH)
            stack.push_back({it.s, mid, l});
            stack.push_back({mid, it.e, r});
        }
    }
    // test order: internal nodes in pre-order, leaf children tested at their parent (left, then right)
    uint32_t rank = 0;
    std::vector<int> walk{0};
    while (!walk.empty()) {
        const int nd = walk.back();
        walk.pop_back();
        const Shape& s = nodes[nd];
        if (s.leaf) { prio[s.prim] = rank++; continue; }  // only a single-triangle root
        if (nodes[s.left].leaf) prio[nodes[s.left].prim] = rank++;
        if (nodes[s.right].leaf) prio[nodes[s.right].prim] = rank++;
        if (!nodes[s.right].leaf) walk.push_back(s.right);
        if (!nodes[s.left].leaf) walk.push_back(s.left);
    }
    return prio;
}

}  // namespace srt
