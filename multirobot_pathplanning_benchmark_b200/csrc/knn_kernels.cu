// Distance / nearest-neighbour kernels (exact fp64 path on the CUDA cores).
//
// Replaces (P/ = src/multi_robot_multi_goal_planning/ in the reference):
//   batch_config_dist one-to-many                         P/problems/core/configuration.py:303-349
//   PRM k-nearest: argpartition + argsort                 P/planners/prm/prm_graph.py:440-447
//   PRM r-disc: dists < r in index order                  P/planners/prm/prm_graph.py:488-500
//   RRT* / IT* radius: dists <= r + 1e-10                 P/planners/rrtstar_base.py:439-453,
//                                                         P/planners/itstar_base.py:1388-1527
// Each thread owns one query row and streams the corpus, tile by tile, from shared memory
// (tiles arrive by 1-D bulk async copies, double buffered); k-NN keeps a bounded max-heap per
// thread, radius search counts then fills so that indices come out in ascending order.
// The tensor-core (tcgen05) candidate generator for euclidean / max_euclidean lives in
// knn_tc_kernels.cu and is re-ranked with the same fp64 arithmetic as this file.
#include <cuda_runtime.h>
#include <stdint.h>

#include "async_copy.cuh"
#include "kernels.h"
#include "knn_common.cuh"

namespace mrb {

constexpr int KT = 128;  // threads per CTA = query rows per CTA
constexpr int TN = 64;   // corpus points per shared-memory tile
constexpr int RADIUS_ROWS_MAX_Q = 512;   // r-disc searches with at most this many rows take radius_rows_kernel

// ---------------------------------------------------------------------------------------------
// one-to-many distances (A10)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) batch_dist_kernel(const double* __restrict__ q, const double* __restrict__ pts, int64_t N,
                                                         int D, const __grid_constant__ Slices sl, int metric,
                                                         double* __restrict__ out) {
    __shared__ double qs[KNN_MAX_D];
    if (threadIdx.x < D) qs[threadIdx.x] = q[threadIdx.x];
    __syncthreads();
    double ql[KNN_MAX_D];
#pragma unroll
    for (int k = 0; k < KNN_MAX_D; k++) ql[k] = k < D ? qs[k] : 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = metric_dist<KNN_MAX_D>(ql, pts + i * D, D, sl, metric);
}

// ---------------------------------------------------------------------------------------------
// batch_config_cost (A11): per-robot euclidean or max-abs distance, reduced by sum or by
// max + w * sum (P/problems/core/configuration.py:156-171, 491-510).  a: one row (stride 0) or N rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) batch_cost_kernel(const double* __restrict__ a, int64_t a_stride, const double* __restrict__ b,
                                                         int64_t N, int D, const __grid_constant__ Slices sl, int per_robot_max,
                                                         int reduction_sum, double w, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double* pa = a + i * a_stride;
        const double* pb = b + i * D;
        double mx = 0.0, sum = 0.0;
        for (int r = 0; r < sl.R; r++) {
            double d = 0.0;
            if (per_robot_max) {
                for (int k = sl.start[r]; k < sl.end[r]; k++) d = fmax(d, fabs(__dsub_rn(pa[k], pb[k])));
            } else {
                double s = 0.0;
                for (int k = sl.start[r]; k < sl.end[r]; k++) {
                    const double x = __dsub_rn(pa[k], pb[k]);
                    s = __dadd_rn(s, __dmul_rn(x, x));
                }
                d = __dsqrt_rn(s);
            }
            mx = r == 0 ? d : fmax(mx, d);
            sum = r == 0 ? d : __dadd_rn(sum, d);
        }
        out[i] = reduction_sum ? sum : __dadd_rn(mx, __dmul_rn(w, sum));
    }
}

// ---------------------------------------------------------------------------------------------
// one relaxation step of the goal lower bound (P/planners/prm/prm_graph.py:143-220 runs it node by node):
//   out[i] = min_j ( cost(a_i, b_j) + lb_b[j] ),   arg[i] = the minimising j (smallest j on ties)
// a = the exit configurations of one mode, b = those of the modes it leads into with their bounds.  Warp per row of a,
// lanes stride over b, the cost arithmetic is that of batch_cost_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) minplus_cost_kernel(const double* __restrict__ a, int64_t T1, const double* __restrict__ b,
                                                           const double* __restrict__ lb_b, int64_t T2, int D,
                                                           const __grid_constant__ Slices sl, int per_robot_max, int reduction_sum,
                                                           double w, double* __restrict__ out, int32_t* __restrict__ arg) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < T1; i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const double* pa = a + i * D;
        double best = __longlong_as_double(0x7ff0000000000000LL);
        int bj = 0x7fffffff;
        for (int64_t j = lane; j < T2; j += 32) {
            const double* pb = b + j * D;
            double mx = 0.0, sum = 0.0;
            for (int r = 0; r < sl.R; r++) {
                double d = 0.0;
                if (per_robot_max) {
                    for (int k = sl.start[r]; k < sl.end[r]; k++) d = fmax(d, fabs(__dsub_rn(pa[k], pb[k])));
                } else {
                    double s2 = 0.0;
                    for (int k = sl.start[r]; k < sl.end[r]; k++) {
                        const double x = __dsub_rn(pa[k], pb[k]);
                        s2 = __dadd_rn(s2, __dmul_rn(x, x));
                    }
                    d = __dsqrt_rn(s2);
                }
                mx = r == 0 ? d : fmax(mx, d);
                sum = r == 0 ? d : __dadd_rn(sum, d);
            }
            const double c = __dadd_rn(reduction_sum ? sum : __dadd_rn(mx, __dmul_rn(w, sum)), lb_b[j]);
            if (c < best) { best = c; bj = (int)j; }     // (j ascending within a lane: ties keep the smaller j)
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
            if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (lane == 0) {
            out[i] = best;
            if (arg) arg[i] = bj == 0x7fffffff ? -1 : bj;
        }
    }
}

cudaError_t launch_minplus_cost(const double* a, int64_t T1, const double* b, const double* lb_b, int64_t T2, int D, const Slices& sl,
                                int per_robot_max, int reduction_sum, double w, double* out, int32_t* arg, cudaStream_t st) {
    if (T1 <= 0) return cudaSuccess;
    int64_t blocks = (T1 * 32 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    minplus_cost_kernel<<<(int)blocks, 256, 0, st>>>(a, T1, b, lb_b, T2, D, sl, per_robot_max, reduction_sum, w, out, arg);
    return cudaGetLastError();
}

cudaError_t launch_batch_cost(const double* a, int64_t a_stride, const double* b, int64_t N, int D, const Slices& sl, int per_robot_max,
                              int reduction_sum, double w, double* out, cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    int64_t blocks = (N + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    batch_cost_kernel<<<(int)blocks, 256, 0, st>>>(a, a_stride, b, N, D, sl, per_robot_max, reduction_sum, w, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// corpus streaming: calls sink(dist, index) for every corpus point in [n0, n1), ascending
// ---------------------------------------------------------------------------------------------
template <int DMAX, typename Sink>
__device__ __forceinline__ void stream_corpus(const double* __restrict__ corpus, int64_t n0, int64_t n1, int D, const double* qrow,
                                              bool row_valid, const Slices& sl, int metric, double* tiles, uint64_t* bars,
                                              bool bulk_ok, Sink&& sink) {
    const int64_t n_tiles = (n1 - n0 + TN - 1) / TN;
    uint32_t phase[2] = {0, 0};
    auto fetch = [&](int64_t t, int buf) {
        const int64_t first = n0 + t * TN;
        const int cnt = (int)min((int64_t)TN, n1 - first);
        const uint32_t bytes = (uint32_t)cnt * D * 8;
        if (bulk_ok && (bytes & 15) == 0) {
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(&bars[buf], bytes);
                bulk_g2s(tiles + (size_t)buf * TN * D, corpus + first * D, bytes, &bars[buf]);
            }
        } else {
            for (int i = threadIdx.x; i < cnt * D; i += KT) tiles[(size_t)buf * TN * D + i] = corpus[first * D + i];
        }
    };
    if (n_tiles > 0) fetch(0, 0);
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int buf = (int)(t & 1);
        if (t + 1 < n_tiles) fetch(t + 1, buf ^ 1);
        const int64_t first = n0 + t * TN;
        const int cnt = (int)min((int64_t)TN, n1 - first);
        const uint32_t bytes = (uint32_t)cnt * D * 8;
        if (bulk_ok && (bytes & 15) == 0) {
            mbar_wait(&bars[buf], phase[buf]);
            phase[buf] ^= 1;
        } else {
            __syncthreads();
        }
        if (row_valid) {
            const double* tp = tiles + (size_t)buf * TN * D;
            for (int j = 0; j < cnt; ++j) sink(metric_dist<DMAX>(qrow, tp + j * D, D, sl, metric), (int)(first + j));
        }
        __syncthreads();  // tile consumed: its buffer may be refilled
    }
}

// ---------------------------------------------------------------------------------------------
// exact k-NN: grid (query tiles, corpus splits); partial lists [split][Q][k], unsorted heaps
// ---------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(KT) knn_exact_kernel(const double* __restrict__ queries, const double* __restrict__ corpus, int64_t Q,
                                                       int64_t N, int D, const __grid_constant__ Slices sl, int metric, int k,
                                                       int64_t split_len, double* __restrict__ part_d, int* __restrict__ part_i,
                                                       int bulk_ok, const uint8_t* __restrict__ skip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);                  // [2][TN][D]
    double* hk = tiles + (size_t)2 * TN * D;                                // [k][KT]
    int* hi = reinterpret_cast<int*>(hk + (size_t)k * KT);                  // [k][KT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(hi + (size_t)k * KT);      // [2]
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t row = blockIdx.x * (int64_t)KT + threadIdx.x;
    // rows flagged in `skip` (already answered and certified by the tensor-core path) are left alone
    const bool valid = row < Q && !(skip && skip[row]);
    if (!__syncthreads_or(valid)) return;
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = (valid && d < D) ? queries[row * D + d] : 0.0;
    ThreadHeap<double> heap(hk + threadIdx.x, hi + threadIdx.x, KT, k, 0);
    const int64_t n0 = blockIdx.y * split_len, n1 = min(N, n0 + split_len);
    stream_corpus<DMAX>(corpus, n0, n1, D, q, valid, sl, metric, tiles, bars, bulk_ok != 0, [&](double d, int idx) {
        if (heap.accepts(d, idx)) heap.push(d, idx);
    });
    if (valid) {
        double* od = part_d + ((size_t)blockIdx.y * Q + row) * k;
        int* oi = part_i + ((size_t)blockIdx.y * Q + row) * k;
        for (int e = 0; e < k; e++) {
            od[e] = e < heap.n ? heap.key[e * KT] : __longlong_as_double(0x7ff0000000000000LL);
            oi[e] = e < heap.n ? heap.idx[e * KT] : -1;
        }
    }
}

// merge the partial lists of one row and emit the k nearest in ascending (distance, index) order
__global__ void __launch_bounds__(KT) knn_merge_kernel(const double* __restrict__ part_d, const int* __restrict__ part_i, int64_t Q, int k,
                                                       int splits, int32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                                       const uint8_t* __restrict__ skip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* hk = reinterpret_cast<double*>(smem_raw);
    int* hi = reinterpret_cast<int*>(hk + (size_t)k * KT);
    const int64_t row = blockIdx.x * (int64_t)KT + threadIdx.x;
    if (row >= Q || (skip && skip[row])) return;
    ThreadHeap<double> heap(hk + threadIdx.x, hi + threadIdx.x, KT, k, 0);
    for (int s = 0; s < splits; s++) {
        const double* pd = part_d + ((size_t)s * Q + row) * k;
        const int* pi = part_i + ((size_t)s * Q + row) * k;
        for (int e = 0; e < k; e++) {
            const int idx = pi[e];
            if (idx >= 0 && heap.accepts(pd[e], idx)) heap.push(pd[e], idx);
        }
    }
    const int found = heap.n;
    for (int e = found - 1; e >= 0; e--) {  // popping the maximum fills the output back to front
        double d;
        int i;
        heap.pop(&d, &i);
        out_idx[row * k + e] = i;
        if (out_dist) out_dist[row * k + e] = d;
    }
    for (int e = found; e < k; e++) {
        out_idx[row * k + e] = -1;
        if (out_dist) out_dist[row * k + e] = __longlong_as_double(0x7ff0000000000000LL);
    }
}

// ---------------------------------------------------------------------------------------------
// radius search: counts[row][split], then fill at the scanned offsets (ascending index order)
// ---------------------------------------------------------------------------------------------
template <int DMAX, bool FILL>
__global__ void __launch_bounds__(KT) radius_kernel(const double* __restrict__ queries, const double* __restrict__ corpus, int64_t Q,
                                                    int64_t N, int D, const __grid_constant__ Slices sl, int metric,
                                                    const double* __restrict__ radii, double radius, int inclusive, int64_t split_len,
                                                    int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                                    int32_t* __restrict__ out_idx, double* __restrict__ out_dist, int bulk_ok) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tiles = reinterpret_cast<double*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)2 * TN * D);
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t row = blockIdx.x * (int64_t)KT + threadIdx.x;
    const bool valid = row < Q;
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = (valid && d < D) ? queries[row * D + d] : 0.0;
    // PRM: d < r (prm_graph.py:500); RRT*/IT*: d <= r + 1e-10 (rrtstar_base.py:439-453)
    double r = valid ? (radii ? radii[row] : radius) : 0.0;
    if (inclusive) r = __dadd_rn(r, 1e-10);
    const int64_t n0 = blockIdx.y * split_len, n1 = min(N, n0 + split_len);
    const int64_t slot = valid ? row * gridDim.y + blockIdx.y : 0;
    int64_t cnt = 0;
    const int64_t base = (FILL && valid) ? offsets[slot] : 0;
    stream_corpus<DMAX>(corpus, n0, n1, D, q, valid, sl, metric, tiles, bars, bulk_ok != 0, [&](double d, int idx) {
        const bool in = inclusive ? (d <= r) : (d < r);
        if (in) {
            if (FILL) {
                out_idx[base + cnt] = idx;
                if (out_dist) out_dist[base + cnt] = d;
            }
            cnt++;
        }
    });
    if (!FILL && valid) counts[slot] = cnt;
}

// Few query rows (single RRT* / IT* `near` queries, the rows that overflow the tensor path's candidate buffers): one CTA per
// (row, corpus split), thread = corpus point, matches written in index order through a block-wide ballot scan.  Same slot
// layout as radius_kernel (counts / offsets [row][split]).  The thread-per-row kernel walks a split with ONE thread per
// row: 2.2 ms per pass for 72 rows against 100 000 points, whatever the row count.
constexpr int RR_THREADS = 256;
template <int DMAX, bool FILL>
__global__ void __launch_bounds__(RR_THREADS) radius_rows_kernel(const double* __restrict__ queries, const double* __restrict__ corpus, int64_t Q,
                                                                 int64_t N, int D, const __grid_constant__ Slices sl, int metric,
                                                                 const double* __restrict__ radii, double radius, int inclusive,
                                                                 int64_t split_len, int64_t* __restrict__ counts,
                                                                 const int64_t* __restrict__ offsets, int32_t* __restrict__ out_idx,
                                                                 double* __restrict__ out_dist) {
    __shared__ int wsum[RR_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t row = blockIdx.x;
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = d < D ? queries[row * D + d] : 0.0;
    double r = radii ? radii[row] : radius;
    if (inclusive) r = __dadd_rn(r, 1e-10);
    const int64_t n0 = blockIdx.y * split_len, n1 = min(N, n0 + split_len);
    const int64_t slot = row * gridDim.y + blockIdx.y;
    const int64_t base = FILL ? offsets[slot] : 0;
    int64_t cnt = 0;
    for (int64_t b0 = n0; b0 < n1; b0 += RR_THREADS) {
        const int64_t i = b0 + tid;
        bool in = false;
        double d = 0.0;
        if (i < n1) {
            d = metric_dist<DMAX>(q, corpus + (size_t)i * D, D, sl, metric);
            in = inclusive ? (d <= r) : (d < r);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RR_THREADS / 32; w++) {
            const int c = wsum[w];
            before += w < warp ? c : 0;
            total += c;
        }
        if (FILL && in) {
            const int64_t pos = base + cnt + before + __popc(bal & ((1u << lane) - 1u));
            out_idx[pos] = (int32_t)i;
            if (out_dist) out_dist[pos] = d;
        }
        cnt += total;
        __syncthreads();
    }
    if (!FILL && tid == 0) counts[slot] = cnt;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static int sm_count_knn() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

int knn_pick_splits(int64_t Q, int64_t N) {
    const int64_t qt = (Q + KT - 1) / KT;
    const int64_t want = 2 * (int64_t)sm_count_knn();
    int64_t s = (want + qt - 1) / qt;
    const int64_t max_s = (N + 4 * TN - 1) / (4 * TN);  // at least four tiles per split
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

cudaError_t launch_batch_dist(const double* q, const double* pts, int64_t N, int D, const Slices& sl, int metric, double* out,
                              cudaStream_t st) {
    if (N <= 0) return cudaSuccess;
    int64_t blocks = (N + 255) / 256;
    const int64_t cap = (int64_t)sm_count_knn() * 8;
    if (blocks > cap) blocks = cap;
    batch_dist_kernel<<<(int)blocks, 256, 0, st>>>(q, pts, N, D, sl, metric, out);
    return cudaGetLastError();
}

template <int DMAX>
static cudaError_t knn_exact_t(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                               int k, int splits, double* part_d, int* part_i, int bulk_ok, const uint8_t* skip, cudaStream_t st) {
    const size_t smem = (size_t)2 * TN * D * 8 + (size_t)k * KT * 12 + 16;
    cudaError_t e = cudaFuncSetAttribute(knn_exact_kernel<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t split_len = ((N + splits - 1) / splits + TN - 1) / TN * TN;
    dim3 grid((unsigned)((Q + KT - 1) / KT), (unsigned)splits);
    knn_exact_kernel<DMAX><<<grid, KT, smem, st>>>(queries, corpus, Q, N, D, sl, metric, k, split_len, part_d, part_i, bulk_ok, skip);
    return cudaGetLastError();
}

cudaError_t launch_knn_exact(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                             int k, int splits, double* part_d, int* part_i, int32_t* out_idx, double* out_dist, const uint8_t* skip,
                             cudaStream_t st) {
    if (Q <= 0) return cudaSuccess;
    const int bulk_ok = (((uintptr_t)corpus) & 15) == 0;
    cudaError_t e;
    if (D <= 8) e = knn_exact_t<8>(queries, corpus, Q, N, D, sl, metric, k, splits, part_d, part_i, bulk_ok, skip, st);
    else if (D <= 16) e = knn_exact_t<16>(queries, corpus, Q, N, D, sl, metric, k, splits, part_d, part_i, bulk_ok, skip, st);
    else if (D <= 24) e = knn_exact_t<24>(queries, corpus, Q, N, D, sl, metric, k, splits, part_d, part_i, bulk_ok, skip, st);
    else if (D <= 32) e = knn_exact_t<32>(queries, corpus, Q, N, D, sl, metric, k, splits, part_d, part_i, bulk_ok, skip, st);
    else e = knn_exact_t<64>(queries, corpus, Q, N, D, sl, metric, k, splits, part_d, part_i, bulk_ok, skip, st);
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)k * KT * 12;
    e = cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    knn_merge_kernel<<<(unsigned)((Q + KT - 1) / KT), KT, smem, st>>>(part_d, part_i, Q, k, splits, out_idx, out_dist, skip);
    return cudaGetLastError();
}

template <int DMAX, bool FILL>
static cudaError_t radius_t(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                            const double* radii, double radius, int inclusive, int splits, int64_t* counts, const int64_t* offsets,
                            int32_t* out_idx, double* out_dist, int bulk_ok, cudaStream_t st) {
    const size_t smem = (size_t)2 * TN * D * 8 + 16;
    const int64_t split_len = ((N + splits - 1) / splits + TN - 1) / TN * TN;
    if (Q <= RADIUS_ROWS_MAX_Q) {   // few rows: a CTA per (row, split)
        radius_rows_kernel<DMAX, FILL><<<dim3((unsigned)Q, (unsigned)splits), RR_THREADS, 0, st>>>(queries, corpus, Q, N, D, sl, metric, radii, radius,
                                                                                                   inclusive, split_len, counts, offsets, out_idx,
                                                                                                   out_dist);
        return cudaGetLastError();
    }
    dim3 grid((unsigned)((Q + KT - 1) / KT), (unsigned)splits);
    if (smem > 48 * 1024) {   // D >= 48: above the default dynamic shared-memory limit
        cudaError_t e = cudaFuncSetAttribute(radius_kernel<DMAX, FILL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    radius_kernel<DMAX, FILL><<<grid, KT, smem, st>>>(queries, corpus, Q, N, D, sl, metric, radii, radius, inclusive, split_len, counts,
                                                      offsets, out_idx, out_dist, bulk_ok);
    return cudaGetLastError();
}

cudaError_t launch_radius(bool fill, const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl,
                          int metric, const double* radii, double radius, int inclusive, int splits, int64_t* counts,
                          const int64_t* offsets, int32_t* out_idx, double* out_dist, cudaStream_t st) {
    if (Q <= 0) return cudaSuccess;
    const int bulk_ok = (((uintptr_t)corpus) & 15) == 0;
#define MRB_RADIUS(DM)                                                                                                              \
    (fill ? radius_t<DM, true>(queries, corpus, Q, N, D, sl, metric, radii, radius, inclusive, splits, counts, offsets, out_idx,    \
                               out_dist, bulk_ok, st)                                                                               \
          : radius_t<DM, false>(queries, corpus, Q, N, D, sl, metric, radii, radius, inclusive, splits, counts, offsets, out_idx,   \
                                out_dist, bulk_ok, st))
    if (D <= 8) return MRB_RADIUS(8);
    if (D <= 16) return MRB_RADIUS(16);
    if (D <= 24) return MRB_RADIUS(24);
    if (D <= 32) return MRB_RADIUS(32);
    return MRB_RADIUS(64);
#undef MRB_RADIUS
}

}  // namespace mrb
