// Shared pieces of the neighbour-search kernels: metric evaluation in the reference's operand
// order and a per-thread bounded max-heap living in shared memory.
//
// Reference semantics (P/ = src/multi_robot_multi_goal_planning/ in the reference):
//   batch_config_dist / NpConfiguration._batch_dist      P/problems/core/configuration.py:303-349
//   compute_sliced_euclidean_dists (sequential sum of squares, then sqrt)            :101-126
//   compute_sum_reduction / compute_max_reduction / compute_abs_max_reduction        :129-219
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mrb {

enum Metric : int { METRIC_EUCLIDEAN = 0, METRIC_SUM_EUCLIDEAN = 1, METRIC_MAX_EUCLIDEAN = 2, METRIC_MAX = 3 };

constexpr int KNN_MAX_D = 64;
constexpr int KNN_MAX_R = 16;

struct Slices {
    int R;
    int start[KNN_MAX_R];
    int end[KNN_MAX_R];
};

// operand plan of the tensor-core candidate generator (knn_tc_kernels.cu)
struct TcPlan {
    int n_acc;               // accumulators = robot slices (1 for euclidean)
    int start[KNN_MAX_R];    // slice of each accumulator
    int dim[KNN_MAX_R];
    int kstep0[KNN_MAX_R];   // first K step (of 8) of each accumulator
    int ksteps[KNN_MAX_R];
    int KS;                  // total K steps
    int tn;                  // corpus columns per accumulator tile (UMMA N)
};

// distance between q (registers / local) and p (any memory), fp64, reference operand order:
// diff = q - p; per-slice sum of squares left to right; sqrt; reduce over slices
template <int DMAX>
__device__ __forceinline__ double metric_dist(const double* q, const double* p, int D, const Slices& sl, int metric) {
    if (metric == METRIC_MAX) {
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < DMAX; k++)
            if (k < D) m = fmax(m, fabs(__dsub_rn(q[k], p[k])));
        return m;
    }
    if (metric == METRIC_EUCLIDEAN) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DMAX; k++)
            if (k < D) {
                const double d = __dsub_rn(q[k], p[k]);
                s = __dadd_rn(s, __dmul_rn(d, d));
            }
        return __dsqrt_rn(s);
    }
    double acc = 0.0;
    for (int r = 0; r < sl.R; r++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DMAX; k++)
            if (k >= sl.start[r] && k < sl.end[r]) {
                const double d = __dsub_rn(q[k], p[k]);
                s = __dadd_rn(s, __dmul_rn(d, d));
            }
        const double dr = __dsqrt_rn(s);
        if (metric == METRIC_SUM_EUCLIDEAN) acc = r == 0 ? dr : __dadd_rn(acc, dr);
        else acc = r == 0 ? dr : fmax(acc, dr);
    }
    return acc;
}

// Bounded max-heap of (key, index) pairs owned by one thread, stored column-wise in shared
// memory (entry e of thread t at [e * stride + t]) so that lanes never bank-conflict.
// Order: lexicographic (key, index) -- ties go to the smaller index, the deterministic choice
// among the results np.argpartition may return.
template <typename KeyT>
struct ThreadHeap {
    KeyT* key;
    int* idx;
    int stride, cap, n;

    __device__ __forceinline__ bool less(KeyT ka, int ia, KeyT kb, int ib) const { return ka < kb || (ka == kb && ia < ib); }
    __device__ __forceinline__ KeyT top_key() const { return key[0]; }
    __device__ __forceinline__ int top_idx() const { return idx[0]; }
    __device__ __forceinline__ bool full() const { return n == cap; }
    // does (k, i) belong in the heap?
    __device__ __forceinline__ bool accepts(KeyT k, int i) const { return n < cap || less(k, i, key[0], idx[0]); }

    __device__ __forceinline__ void push(KeyT k, int i) {
        if (n < cap) {  // sift up
            int c = n++;
            while (c > 0) {
                const int p = (c - 1) >> 1;
                const KeyT pk = key[p * stride];
                const int pi = idx[p * stride];
                if (!less(pk, pi, k, i)) break;
                key[c * stride] = pk;
                idx[c * stride] = pi;
                c = p;
            }
            key[c * stride] = k;
            idx[c * stride] = i;
        } else {  // replace the maximum, sift down
            int p = 0;
            for (;;) {
                int c = 2 * p + 1;
                if (c >= n) break;
                KeyT ck = key[c * stride];
                int ci = idx[c * stride];
                if (c + 1 < n) {
                    const KeyT ck2 = key[(c + 1) * stride];
                    const int ci2 = idx[(c + 1) * stride];
                    if (less(ck, ci, ck2, ci2)) { c++; ck = ck2; ci = ci2; }
                }
                if (!less(k, i, ck, ci)) break;
                key[p * stride] = ck;
                idx[p * stride] = ci;
                p = c;
            }
            key[p * stride] = k;
            idx[p * stride] = i;
        }
    }
    // remove and return the maximum
    __device__ __forceinline__ void pop(KeyT* k, int* i) {
        *k = key[0];
        *i = idx[0];
        n--;
        if (n > 0) {
            const KeyT lk = key[n * stride];
            const int li = idx[n * stride];
            int p = 0;
            for (;;) {
                int c = 2 * p + 1;
                if (c >= n) break;
                KeyT ck = key[c * stride];
                int ci = idx[c * stride];
                if (c + 1 < n) {
                    const KeyT ck2 = key[(c + 1) * stride];
                    const int ci2 = idx[(c + 1) * stride];
                    if (less(ck, ci, ck2, ci2)) { c++; ck = ck2; ci = ci2; }
                }
                if (!less(lk, li, ck, ci)) break;
                key[p * stride] = ck;
                idx[p * stride] = ci;
                p = c;
            }
            key[p * stride] = lk;
            idx[p * stride] = li;
        }
    }
};

}  // namespace mrb
