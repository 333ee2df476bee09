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
    int stages;              // shared-memory stages of corpus tiles (2..4)
    int sc;                  // columns an epilogue warp reads per tcgen05.ld batch (0: generic 16-column epilogue, > 4 accumulators)
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
    int stride;   // element stride of the key array
    int cap, n;
    int istride;  // element stride of the index array (keys and indices may live in different memories)

    __device__ __forceinline__ ThreadHeap(KeyT* k, int* i, int stride_, int cap_, int n_ = 0, int istride_ = -1)
        : key(k), idx(i), stride(stride_), cap(cap_), n(n_), istride(istride_ < 0 ? stride_ : istride_) {}

    // (ka, idx[ia_pos]) < (kb, ib): indices are only looked at on exact key ties
    __device__ __forceinline__ bool less_ki(KeyT ka, int ia, KeyT kb, int ib) const { return ka < kb || (ka == kb && ia < ib); }
    __device__ __forceinline__ KeyT top_key() const { return key[0]; }
    __device__ __forceinline__ int top_idx() const { return idx[0]; }
    __device__ __forceinline__ bool full() const { return n == cap; }
    __device__ __forceinline__ bool accepts(KeyT k, int i) const {
        if (n < cap) return true;
        const KeyT t = key[0];
        return k < t || (k == t && i < idx[0]);
    }
    // entry at heap position a  <  (k, i) ?
    __device__ __forceinline__ bool pos_less(int a, KeyT ka, KeyT k, int i) const { return ka < k || (ka == k && idx[a * istride] < i); }

    __device__ __forceinline__ void push(KeyT k, int i) {
        if (n < cap) {  // sift up
            int c = n++;
            while (c > 0) {
                const int p = (c - 1) >> 1;
                const KeyT pk = key[p * stride];
                if (!pos_less(p, pk, k, i)) break;
                key[c * stride] = pk;
                idx[c * istride] = idx[p * istride];
                c = p;
            }
            key[c * stride] = k;
            idx[c * istride] = i;
        } else {  // replace the maximum, sift down
            int p = 0;
            for (;;) {
                int c = 2 * p + 1;
                if (c >= n) break;
                KeyT ck = key[c * stride];
                if (c + 1 < n) {
                    const KeyT ck2 = key[(c + 1) * stride];
                    if (ck < ck2 || (ck == ck2 && idx[c * istride] < idx[(c + 1) * istride])) { c++; ck = ck2; }
                }
                // stop when (k, i) >= child
                if (!(k < ck || (k == ck && i < idx[c * istride]))) break;
                key[p * stride] = ck;
                idx[p * istride] = idx[c * istride];
                p = c;
            }
            key[p * stride] = k;
            idx[p * istride] = i;
        }
    }
    // remove and return the maximum
    __device__ __forceinline__ void pop(KeyT* k, int* i) {
        *k = key[0];
        *i = idx[0];
        n--;
        if (n > 0) {
            const KeyT lk = key[n * stride];
            const int li = idx[n * istride];
            int p = 0;
            for (;;) {
                int c = 2 * p + 1;
                if (c >= n) break;
                KeyT ck = key[c * stride];
                if (c + 1 < n) {
                    const KeyT ck2 = key[(c + 1) * stride];
                    if (ck < ck2 || (ck == ck2 && idx[c * istride] < idx[(c + 1) * istride])) { c++; ck = ck2; }
                }
                if (!(lk < ck || (lk == ck && li < idx[c * istride]))) break;
                key[p * stride] = ck;
                idx[p * istride] = idx[c * istride];
                p = c;
            }
            key[p * stride] = lk;
            idx[p * istride] = li;
        }
    }
};

}  // namespace mrb
