// Edge discretisation order of the reference, computed in closed form.
//
// generate_binary_search_indices(N) (P/problems/planning_env.py:34-51) is the breadth-first
// order of the midpoint tree over [0, N-1]: an interval holding n indices splits into a left
// part of (n-1)/2 and a right part of n-1-(n-1)/2.  Levels 0..H-1 (H = floor(log2 N)) are
// full; the last level holds the remaining N-(2^H-1) nodes left to right.  The p-th visited
// index is found by one root-to-node descent that counts how many depth-h nodes the left
// subtree holds -- no table, O(log N).
#pragma once
#include <cuda_runtime.h>

namespace mrb {

__host__ __device__ __forceinline__ int nodes_at_depth(int n, int h) {
    int full = (1 << h) - 1;
    int c = n - full;
    c = c < 0 ? 0 : c;
    return c < (1 << h) ? c : (1 << h);
}

__host__ __device__ __forceinline__ int floor_log2(unsigned x) {
#ifdef __CUDA_ARCH__
    return 31 - __clz(x);
#else
    int r = -1;
    while (x) { x >>= 1; r++; }
    return r;
#endif
}

__host__ __device__ __forceinline__ int binary_order_index(int N, int p) {
    int H = floor_log2((unsigned)N);
    int t = floor_log2((unsigned)(p + 1));
    t = t < H ? t : H;
    int r = p - ((1 << t) - 1);
    int s = 0, n = N, h = t;
    while (h > 0) {
        int nl = (n - 1) >> 1;
        int cl = nodes_at_depth(nl, h - 1);
        if (r < cl) {
            n = nl;
        } else {
            r -= cl;
            s += nl + 1;
            n = n - 1 - nl;
        }
        h--;
    }
    return s + ((n - 1) >> 1);
}

}  // namespace mrb
