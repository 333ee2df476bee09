// Tensor-core candidate generator for k-NN (euclidean / max_euclidean) + exact fp64 re-rank.
//
// Replaces the reference's per-query loop  dists = batch_dist_fun(q, corpus); argpartition; argsort
// (P/planners/prm/prm_graph.py:407-447, distances P/problems/core/configuration.py:303-329) for
// whole query batches.
//
// Per robot slice the squared distance is a dense contraction of augmented vectors, rounded ONCE to TF32:
//   A-row (query):  [  q_1 ..  q_d | |q|^2 | 1     | 0.. ]
//   B-row (corpus): [-2c_1 .. -2c_d | 1     | |c|^2 | 0.. ]        (K = d + 2 -> one K step of 8 for a 6-dof arm)
// The coarse value carries an absolute error of at most (2^-9 + 2^-10) * max squared slice norm (input rounding;
// products of TF32 numbers are exact in the FP32 accumulator), which the certification below accounts for -- round 1
// used a 3xTF32 split (three times the MMAs) to make the coarse value nearly exact, which the re-rank never needed.
// knn_tc_prep_kernel writes the corpus operand in the UMMA canonical K-major, no-swizzle layout (8-row x 16-byte core
// matrices; LBO = 128 B between the two K halves of an instruction, SBO = 256 B between 8-row groups), tile after tile,
// so a whole operand tile is ONE contiguous bulk copy; the query operand is plain row-major and lives in TENSOR MEMORY
// for the life of the CTA (tcgen05.st once, then every MMA reads A from TMEM: shared memory only feeds the small B tile).
// knn_tc_kernel: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (M = 128 query rows, N = corpus columns per
// accumulator, K = 8 per instruction, kind::tf32, FP32 accumulators in TMEM, one accumulator per robot, double
// buffered), warp 2 = TMEM allocator, warps 4-11 = epilogue: each warp owns a lane quarter and a contiguous column part of
// every tile, reads its SC columns of every robot's accumulator with one batch of tcgen05.ld, RELEASES the accumulator
// buffer as soon as the loads have landed (the MMAs of the tile after next overlap the selection work below; a warp that
// is busy compacting a list no longer holds tensor memory), then: max over robots, a min tree + warp vote that skips
// batches without a candidate, else THRESHOLD-FIRST selection: values below the thread's
// current threshold are appended to its list in shared memory (no ordering); when a list fills up, the warp bisects
// for the value that keeps about `m` entries, drops the rest and lowers the threshold.  knn_rerank_kernel recomputes the
// candidates' distances in fp64 with the reference's operand order, sorts by (distance, index) and certifies each
// row: if the exact k-th squared distance is not below (smallest discard threshold - error bound) the row is flagged
// and recomputed exactly by knn_rows_kernel.
//
// Measured and rejected (round 2, profiles/r2_knn_notes.md): one more accumulator holding the SUM over the robots as a
// prefilter of the max (max_r d_r^2 >= S / R, so S >= R * threshold discards a point from one accumulator load instead
// of R).  The per-list thresholds only approach their final value late in the corpus sweep, so about two thirds of the
// 16-column chunks still pass the prefilter for some lane of the warp, and the extra load + MMAs made the kernel slower
// (12.1 -> 14.5 ms at 100k x 100k, four arms).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "async_copy.cuh"
#include "kernels.h"
#include "knn_common.cuh"

namespace mrb {

constexpr int TC_TM = 128;        // query rows per CTA (UMMA M)
constexpr int TC_STAGES = 4;      // shared-memory stages of corpus tiles, at most (TcPlan::stages)
// Measured (round 2, gpurun_out/r2g_knn_variants.log; four arms 100k x 100k / euclidean 100k x 100k / two arms 60k x 60k):
// 2 epilogue warps per lane quarter 13.03 / 8.20 / 4.35 ms, 3 warps 12.75 / 9.03 / 4.18 ms -- a third warp hides latency but
// every list then sees a third of the columns, keeps a looser threshold and appends more; 2 stays.
#ifndef MRB_TC_PARTS
#define MRB_TC_PARTS 2
#endif
#ifndef MRB_TC_NBUF
#define MRB_TC_NBUF 2         // accumulator buffers in tensor memory (tile t uses buffer t % NBUF)
#endif
#ifndef MRB_TC_SLEEP
#define MRB_TC_SLEEP 200      // suspend-time hint (ns) of the producer's and the MMA thread's mbarrier waits: -0.1 ms
#endif
constexpr int TC_NBUF = MRB_TC_NBUF;
constexpr int TC_MMA_WARPS = 2;   // warps 1 and 3 issue alternate tiles
#ifndef MRB_TC_BOOT
#define MRB_TC_BOOT 1
#endif
constexpr int TC_BOOT_STEPS = 64;  // batch minima a list collects before its first threshold (<= TC_LIST - 16)
constexpr int TC_PARTS = MRB_TC_PARTS;              // epilogue warps per TMEM lane quarter: each takes every TC_PARTS-th 16-column chunk
constexpr int TC_THREADS = 32 * (4 + 4 * TC_PARTS); // warps 0-3: producer / MMA / TMEM allocator / spare, then the epilogue warps
constexpr int TC_LIST = TC_PARTS == 2 ? 80 : 64;    // candidate list entries per epilogue thread (row x column part)
constexpr int TC_TRIG = TC_LIST - 16;   // a 16-column chunk can add 16 entries: compact above this fill
constexpr int TC_SLACK = 8;       // the final compaction of a list keeps between m and m + TC_SLACK entries (the output slots)
#ifndef MRB_TC_SLACK_RUN
#define MRB_TC_SLACK_RUN 12
#endif
constexpr int TC_SLACK_RUN = MRB_TC_SLACK_RUN;   // compactions during the sweep accept up to m + TC_SLACK_RUN: more one-round searches
constexpr float TC_BIG = 1.0e30f;      // squared norm of the padded corpus rows
constexpr float TC_THR_MAX = 1.0e29f;  // no append threshold exceeds this: padded columns are never candidates
constexpr int TC_MAX_KS = 40;     // knn_tc_make_plan accepts plans up to this many K steps
constexpr unsigned TC_FULL = 0xffffffffu;
constexpr int TC_MAX_SMS = 192;   // the tail decomposition never assumes more SMs than this (workspace bound)
#ifndef MRB_TC_FAST_EPILOGUE
#define MRB_TC_FAST_EPILOGUE 1  // 0: experiment build with the generic epilogue for every plan
#endif

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T: A = 128 lanes x 8 TF32 columns of tensor memory, B K-major in shared memory, FP32 out
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u));
}
// one lane of a converged warp (ptxas predicates a tensor instruction on this without a loop over the active lanes)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// mbarrier wait with a suspend-time hint (ns): the producer and the MMA thread have stages of slack, and a hot try_wait
// loop competes for issue slots with the epilogue warps on the same scheduler
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_H:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE_H;\n"
        "bra WAIT_LOOP_H;\n"
        "WAIT_DONE_H:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(ns)
        : "memory");
}
#if MRB_TC_SLEEP > 0
#define MRB_TC_WAIT(bar, parity) mbar_wait_hint(bar, parity, MRB_TC_SLEEP)
#else
#define MRB_TC_WAIT(bar, parity) mbar_wait(bar, parity)
#endif
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive FP32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 8 consecutive FP32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// SC (a multiple of 8) consecutive columns as x16 pieces and at most one x8 piece
template <int SC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
#pragma unroll
    for (int c = 0; c + 16 <= SC; c += 16) tmem_ld16(taddr + (uint32_t)c, v + c);
    if constexpr (SC % 16 == 8) tmem_ld8(taddr + (uint32_t)(SC - 8), v + (SC - 8));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, no swizzle, LBO = 128 B, SBO = 256 B, version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(const void* smem) {
    const uint64_t addr = (uint64_t)((smem_u32(smem) >> 4) & 0x3fffu);
    return addr | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__host__ __device__ inline uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// monotone map float -> unsigned (finite floats): bisection over candidate keys works on these
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// ---------------------------------------------------------------------------------------------
// operand layout
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t tc_elem_offset(int rows_per_tile, int KS, int64_t row, int kglob) {
    // float index of element (row, kglob) of the corpus operand, stored tile after tile, K step after K step
    const int64_t tile = row / rows_per_tile;
    const int m = (int)(row - tile * rows_per_tile);
    const int ks = kglob >> 3, k = kglob & 7;
    return (tile * KS + ks) * (int64_t)rows_per_tile * 8 + (m >> 3) * 64 + (k >> 2) * 32 + (m & 7) * 4 + (k & 3);
}

// one thread per (padded) row; side 0 = queries (A, row-major [n_pad][8 KS]), 1 = corpus (B, UMMA tile layout)
__global__ void __launch_bounds__(128) knn_tc_prep_kernel(const double* __restrict__ X, int64_t n, int64_t n_pad, int D,
                                                          const __grid_constant__ TcPlan plan, int side, int rows_per_tile,
                                                          float* __restrict__ out, unsigned* __restrict__ max_norm_bits) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_pad) return;
    const bool real = row < n;
    float worst = 0.f;
    const int KW = plan.KS * 8;
    auto put = [&](int kglob, float v) {
        if (side == 0) out[row * KW + kglob] = v;
        else out[tc_elem_offset(rows_per_tile, plan.KS, row, kglob)] = v;
    };
    for (int a = 0; a < plan.n_acc; a++) {
        const int d = plan.dim[a], k0 = plan.kstep0[a] * 8, Ka = plan.ksteps[a] * 8;
        double nrm = 0.0;
        for (int j = 0; j < d; j++) {
            const double x = real ? X[row * D + plan.start[a] + j] : 0.0;
            nrm += x * x;
            const float t = to_tf32((float)x);
            put(k0 + j, side == 0 ? t : -2.f * t);
        }
        float nh = to_tf32((float)nrm);
        if (!real) nh = side == 1 ? TC_BIG : 0.f;   // padded corpus rows are infinitely far away
        worst = fmaxf(worst, (float)nrm);
        put(k0 + d, side == 0 ? nh : 1.f);
        put(k0 + d + 1, side == 0 ? 1.f : nh);
        for (int j = d + 2; j < Ka; j++) put(k0 + j, 0.f);
    }
    if (real) atomicMax(max_norm_bits, __float_as_uint(worst));  // non-negative floats order like unsigned ints
}

// ---------------------------------------------------------------------------------------------
// main kernel: grid (query tiles, corpus splits)
// ---------------------------------------------------------------------------------------------
struct TcParams {
    const float* A;      // prepared queries  [Q padded to 128][8 KS], row-major
    const float* B;      // prepared corpus   [c tiles][KS][tn x 8], UMMA layout
    int64_t Q, N;
    // work decomposition (1-D grid): CTAs [0, full_qtiles) sweep the whole corpus for one query tile each; the remaining
    // query tiles (the last, partial wave of the grid) are split tail_splits ways along the corpus so that they end
    // together with ~1 / tail_splits of a full sweep.  Small query sets: full_qtiles = 0, every tile is split.
    int64_t full_qtiles;
    int tail_splits;
    int64_t n_ctiles;
    int m;                    // entries a compaction keeps (at least)
    int kc;                   // output slots per row and list = m + TC_SLACK
    float* part_key;          // [list][kc], list = tc_list_id(row, split, part)
    int* part_idx;
    float* part_tau;          // [list]: every point of the list's corpus part NOT in the list has a coarse value >= tau
    // radius mode (r-disc search): fixed per-row thresholds, candidates appended to per-row buffers in global memory
    int radius_mode;
    const double* radii;      // [Q] or null
    double radius;            // used when radii is null
    double radius_pad;        // 1e-10 for the inclusive rule (d <= r + 1e-10), else 0
    const unsigned* max_norm_bits;   // [0] bit pattern of the largest squared slice norm (error bound of the coarse values)
    int* cand_idx;            // [Q][cap]
    int* cand_cnt;            // [Q] candidates found (may exceed cap: the row has overflowed)
    int cap;
    TcPlan plan;
    long long* trace;         // experiment builds (MRB_TC_TRACE): clock64 stamps of one CTA's pipeline events
};
#ifdef MRB_TC_TRACE
constexpr int TC_TRACE_T0 = 1000, TC_TRACE_N = 48, TC_TRACE_EV = 20;
#define MRB_TRACE(ev, t)                                                                                           \
    do {                                                                                                           \
        if (p.trace && blockIdx.x == 200 && (t) >= TC_TRACE_T0 && (t) < TC_TRACE_T0 + TC_TRACE_N)                  \
            p.trace[(ev) * TC_TRACE_N + ((t) - TC_TRACE_T0)] = clock64();                                          \
    } while (0)
#else
#define MRB_TRACE(ev, t)
#endif

// candidate lists of a row are consecutive: TC_PARTS per corpus split; rows of the full query tiles have one split
__host__ __device__ inline int64_t tc_list_id(int64_t row, int split, int part, int64_t full_rows, int tail_splits) {
    if (row < full_rows) return row * TC_PARTS + part;
    return full_rows * TC_PARTS + ((row - full_rows) * tail_splits + split) * TC_PARTS + part;
}

// Warp-cooperative compaction of a candidate list (n entries, all below `bound` = the list's append threshold): find a
// threshold T that keeps between m and m + slack entries, move the entries below T to the front, return (kept count,
// T) to every lane.  The search probes SEVEN pivots per round (an eight-way split of the value range, counted with
// ballots) -- one round isolates the cut among ~70 entries about every other time, a second round splits the one bucket
// that was too full.  (Fifteen pivots with packed per-lane counts and one warp reduction per word always finish in one
// round but the round takes as long as two of these: 9.8 against 9.6 ms.)  (Round 2 first bisected
// the ordered bit patterns one pivot at a time: 230 cycles per probe, 2-13 probes, 1000-3000 cycles per compaction, and
// every compaction stalls the CTA's two-buffer pipeline -- clock64 trace.)  If ties make it impossible to get under
// `trig` entries, reports keep = -1: the caller closes the list (the row is then recomputed by the exact kernel).
__device__ __forceinline__ void tc_compact(float* keys, int* idx, int n, int m, int slack, int trig, float bound, int lane, int* kept_out,
                                           float* thr_out, long long* dbg = nullptr) {
    constexpr int PER = (TC_LIST + 31) / 32;
    constexpr int NP = 7;   // pivots per round: an eight-way split of [lo, hi)
    const long long c0 = dbg ? clock64() : 0;
    float k[PER];
    uint32_t lo_u = 0xffffffffu, hi_u = 0u;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int e = lane + 32 * j;
        k[j] = e < n ? keys[e] : 3.0e38f;     // absent entries are never below a pivot
        if (e < n) { lo_u = min(lo_u, f2ord(k[j])); hi_u = max(hi_u, f2ord(k[j])); }
    }
    // invariant: count(k < lo) < m <= count(k < hi)
    float lo = ord2f(__reduce_min_sync(TC_FULL, lo_u));
    float hi = bound;
    if (bound >= TC_THR_MAX) hi = ord2f(__reduce_max_sync(TC_FULL, hi_u) + 1u);   // first compaction of a list: no threshold yet
    const long long c1 = dbg ? clock64() : 0;
    int iters = 0, cnt = n;
    float T = hi;
    while (true) {
        iters++;
        const float w = (hi - lo) * (1.f / (NP + 1));
        float pv[NP];
        int c[NP];
#pragma unroll
        for (int i = 0; i < NP; i++) {
            pv[i] = lo + w * (float)(i + 1);
            c[i] = 0;
        }
#pragma unroll
        for (int j = 0; j < PER; j++)
#pragma unroll
            for (int i = 0; i < NP; i++) c[i] += __popc(__ballot_sync(TC_FULL, k[j] < pv[i]));
        // the first pivot that keeps at least m entries (else the upper end itself); the last one that keeps fewer
        float nlo = lo, nT = hi;
        int ncnt = cnt;
#pragma unroll
        for (int i = NP - 1; i >= 0; i--) {
            if (c[i] >= m) { nT = pv[i]; ncnt = c[i]; }
        }
#pragma unroll
        for (int i = 0; i < NP; i++) {
            if (c[i] < m) nlo = pv[i];
        }
        const bool stuck = nlo == lo && nT == hi;   // the pivots no longer separate anything: ties
        T = nT;
        cnt = ncnt;
        if (cnt <= m + slack || stuck) break;
        lo = nlo;
        hi = nT;
    }
    const long long c2 = dbg ? clock64() : 0;
    if (cnt > m + slack && cnt > trig) {   // more ties than the list can ever shed
        *kept_out = -1;
        *thr_out = -3.0e38f;
        return;
    }
    bool keep[PER];
    unsigned bal[PER];
    int id[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        keep[j] = k[j] < T;
        bal[j] = __ballot_sync(TC_FULL, keep[j]);
        id[j] = keep[j] ? idx[lane + 32 * j] : 0;
    }
    __syncwarp();   // every slot has been read before any is overwritten
    int pos = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        if (keep[j]) {
            const int o = pos + __popc(bal[j] & ((1u << lane) - 1u));
            keys[o] = k[j];
            idx[o] = id[j];
        }
        pos += __popc(bal[j]);
    }
    __syncwarp();
    *kept_out = pos;
    *thr_out = T;
    if (dbg && lane == 0) {
        dbg[0] = c1 - c0;
        dbg[1] = c2 - c1;
        dbg[2] = clock64() - c2;
        dbg[3] = iters;
        dbg[4] = n;
    }
}

// NACC / SC > 0: fast epilogue for exactly NACC accumulators, SC columns per warp and load batch (plan.sc); NACC = 0:
// generic epilogue (any number of accumulators, 16-column chunks alternating between the parts, buffer held until the
// tile's selection work is done)
template <bool RMODE, int NACC, int SC>
__global__ void __launch_bounds__(TC_THREADS, 1) knn_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const TcPlan& plan = p.plan;
    const int KS = plan.KS, tn = plan.tn, n_acc = plan.n_acc, n_stages = plan.stages;
    const uint32_t b_bytes = (uint32_t)KS * tn * 32;
    unsigned char* sB = smem_raw;
    constexpr int LSTR = TC_LIST + 1;   // odd stride: the 32 lanes of a warp append to 32 different banks
    float* lk = reinterpret_cast<float*>(sB + (size_t)n_stages * b_bytes);              // [PARTS * 128][LSTR] candidate keys
    int* li = reinterpret_cast<int*>(lk + (size_t)TC_PARTS * TC_TM * LSTR);              // [PARTS * 128][LSTR] candidate indices
    uint64_t* bars = reinterpret_cast<uint64_t*>(li + (size_t)TC_PARTS * TC_TM * LSTR);  // (PARTS * 128 * LSTR ints: a multiple of 8 bytes)
    uint64_t* full_b = bars;                    // [STAGES] corpus tile landed
    uint64_t* empty_b = bars + TC_STAGES;       // [STAGES] corpus tile consumed by the MMAs
    uint64_t* a_full = bars + 2 * TC_STAGES;    // query tile written to tensor memory (4 warps arrive)
    uint64_t* tm_full = a_full + 1;             // [NBUF] accumulators ready
    uint64_t* tm_empty = tm_full + TC_NBUF;     // [NBUF] accumulators drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + TC_NBUF);
    // issue table of the MMA thread, one entry per K step: B descriptor of stage 0, A column, accumulator column
    // offset | accumulate flag << 31 (descriptor arithmetic inside the issue loop costs the single issuing thread more
    // cycles per instruction than the tensor pipe needs to run it)
    uint64_t* iss_b = reinterpret_cast<uint64_t*>(tmem_slot + 2);    // [KS]
    uint32_t* iss_a = reinterpret_cast<uint32_t*>(iss_b + TC_MAX_KS); // [KS]
    uint32_t* iss_d = iss_a + TC_MAX_KS;                              // [KS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int buf_cols = n_acc * tn;            // TMEM columns per accumulator buffer
    const uint32_t a_col0 = (uint32_t)(TC_NBUF * buf_cols);   // the query operand sits behind the accumulator buffers
    uint32_t alloc_cols = 32;
    while (alloc_cols < a_col0 + (uint32_t)(8 * KS)) alloc_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        mbar_init(a_full, 4);
        for (int b = 0; b < TC_NBUF; b++) { mbar_init(&tm_full[b], 1); mbar_init(&tm_empty[b], 4 * TC_PARTS); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, alloc_cols);
    if (warp == 1) {
        for (int a = 0; a < n_acc; a++)
            for (int ks = lane; ks < plan.ksteps[a]; ks += 32) {
                const int kg = plan.kstep0[a] + ks;
                iss_b[kg] = umma_desc(sB + (size_t)kg * tn * 32);
                iss_a[kg] = a_col0 + (uint32_t)(8 * kg);
                iss_d[kg] = (uint32_t)(a * tn) | (ks > 0 ? 0x80000000u : 0u);
            }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    int64_t qtile = blockIdx.x;
    int split = 0, n_split = 1;
    if (qtile >= p.full_qtiles) {
        const int64_t r = qtile - p.full_qtiles;
        qtile = p.full_qtiles + r / p.tail_splits;
        split = (int)(r % p.tail_splits);
        n_split = p.tail_splits;
    }
    const int64_t tiles_per_split = (p.n_ctiles + n_split - 1) / n_split;
    const int64_t t0 = split * tiles_per_split;
    const int64_t t1 = min(p.n_ctiles, t0 + tiles_per_split);
    const int64_t n_tiles = max((int64_t)0, t1 - t0);
    // Threshold bootstrap (fast epilogue, k-NN mode): the sweep is preceded by n_boot steps over its own first tiles in which
    // every epilogue thread only collects the MINIMUM of each batch it reads; the m-th smallest of these minima -- values of
    // m distinct corpus points -- becomes the list's first threshold, and the real sweep starts over at tile 0 with it.  A
    // list that starts from +inf compacts five times while ~1800 columns go by, all 32 lists of a warp at the same moments;
    // this is one compaction for the price of n_boot extra tiles of read-back (3 % of a full sweep).  Measured (100k x 100k):
    // euclidean 6.02 -> 5.77 ms, two arms 3.52 -> 3.43 ms, four arms 9.42 -> 9.56 ms (its 48-column tiles make the extra steps
    // cost what the saved compactions return): on for one and two accumulators.
    const int64_t n_boot = (NACC > 0 && NACC <= 2 && !RMODE && MRB_TC_BOOT && n_tiles >= 8 * TC_BOOT_STEPS) ? TC_BOOT_STEPS / ((tn / TC_PARTS) / (SC > 0 ? SC : 1)) : 0;
    const int64_t n_steps = n_tiles + n_boot;

    if (warp == 0) {
        // ===== producer: one bulk copy per corpus tile =====
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int64_t st = 0; st < n_steps; ++st) {
                const int64_t t = st < n_boot ? st : st - n_boot;
                MRB_TC_WAIT(&empty_b[s], ph ^ 1);
                MRB_TRACE(7, t);
                mbar_arrive_expect_tx(&full_b[s], b_bytes);
                bulk_g2s(sB + (size_t)s * b_bytes, p.B + (size_t)(t0 + t) * KS * tn * 8, b_bytes, &full_b[s]);
                if (++s == n_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1 || warp == 3) {
        // ===== MMA issuers: warp 1 takes the even tiles, warp 3 the odd ones (with two accumulator buffers each warp owns
        // one); the warps stay converged, one elected lane issues.  One issuer needs ~700 cycles per four-robot tile
        // (barrier polls, fence, uniform-register moves, four tcgen05.mma, two commits -- a serial chain on a scheduler it
        // shares with two epilogue warps), far more than the ~100 cycles the tensor pipe works on the tile, and the
        // epilogue warps were waiting for accumulators a third of their time (clock64 trace, MRB_TC_TRACE). =====
        // (Measured, round 2: a loop that read its issue table from shared memory in front of every tcgen05.mma, inside an
        // `if (lane == 0)` branch, spent ~170 cycles per instruction -- LDS -> R2UR chains serialised behind the previous
        // MMA by the asm memory clobber, plus the ELECT / BRA.U.ANY wrapper ptxas puts around a tensor instruction in
        // divergent code -- 714 cycles per four-robot tile against ~100 of tensor-pipe time, and the epilogue warps waited
        // on it for a third of their time.  Plans of up to 8 K steps keep the whole table in registers.)
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_tf32(tn);
        constexpr int REG_KS = 8;
        uint64_t rb_[REG_KS];
        uint32_t ra_[REG_KS], rd_[REG_KS], rp_[REG_KS];
#pragma unroll
        for (int kg = 0; kg < REG_KS; kg++) {
            const int kk = kg < KS ? kg : 0;
            rb_[kg] = iss_b[kk];
            ra_[kg] = tmem_base + iss_a[kk];
            rd_[kg] = tmem_base + (iss_d[kk] & 0x7fffffffu);
            rp_[kg] = iss_d[kk] >> 31;
        }
        mbar_wait(a_full, 0);
        tc_fence_after();
        for (int64_t t = (warp == 3 ? 1 : 0); t < n_steps; t += TC_MMA_WARPS) {   // (t counts steps: bootstrap steps, then tiles)
            const int buf = (int)(t % TC_NBUF);
            const uint32_t use = (uint32_t)(t / TC_NBUF);      // how often this buffer was used before
            const int s = (int)(t % n_stages);
            const uint32_t ph = (uint32_t)(t / n_stages) & 1u;
            if (lane == 0) MRB_TRACE(0, t);
            MRB_TC_WAIT(&tm_empty[buf], (use & 1) ^ 1);         // first use passes immediately
            if (lane == 0) MRB_TRACE(1, t);
            MRB_TC_WAIT(&full_b[s], ph);
            if (lane == 0) MRB_TRACE(2, t);
            tc_fence_after();
            const uint64_t stage_off = (uint64_t)(((uint32_t)s * b_bytes) >> 4);   // added to the 14-bit address field
            const uint32_t d_off = (uint32_t)(buf * buf_cols);
            if (leader) {
                if (KS <= REG_KS) {
#pragma unroll
                    for (int kg = 0; kg < REG_KS; kg++)
                        if (kg < KS) {
                            umma_tf32_ts(rd_[kg] + d_off, ra_[kg], rb_[kg] + stage_off, idesc, rp_[kg]);
                            MRB_TRACE(10 + kg, t);
                        }
                } else {
#pragma unroll 4
                    for (int kg = 0; kg < KS; kg++) {
                        const uint32_t d = iss_d[kg];
                        umma_tf32_ts(tmem_base + d_off + (d & 0x7fffffffu), tmem_base + iss_a[kg], iss_b[kg] + stage_off, idesc, d >> 31);
                    }
                }
                umma_commit(&empty_b[s]);     // the stage may be refilled once these MMAs have read it
                MRB_TRACE(14, t);
                umma_commit(&tm_full[buf]);   // accumulators complete
                MRB_TRACE(15, t);
            }
            __syncwarp();
            if (lane == 0) MRB_TRACE(3, t);
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; thread = (query row, half); 16-column chunks alternate between the halves =====
        const int quarter = warp & 3, half = (warp - 4) >> 2;   // `half` = this warp's column part, 0 .. TC_PARTS - 1
        const int r_in_tile = quarter * 32 + lane;
        const int64_t row = qtile * TC_TM + r_in_tile;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (half == 0) {
            // the query rows of this lane quarter -> tensor memory (A operand of every MMA of this CTA)
            const float4* arow = reinterpret_cast<const float4*>(p.A + (size_t)row * (size_t)(KS * 8));
            for (int ks = 0; ks < KS; ks++) {
                float v[8];
                const float4 x0 = arow[2 * ks], x1 = arow[2 * ks + 1];
                v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                tmem_st8(lane_addr + a_col0 + (uint32_t)(8 * ks), v);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);
        }
        const int hslot = half * TC_TM + r_in_tile;            // this thread's list
        float* mk = lk + (size_t)hslot * LSTR;
        int* mi = li + (size_t)hslot * LSTR;
        float* wk = lk + (size_t)(hslot - lane) * LSTR;        // list of lane 0 of this warp (lists of a warp are consecutive)
        int* wi = li + (size_t)(hslot - lane) * LSTR;
        int n = 0;                                             // entries in the list
        float thr = TC_THR_MAX;                                // append what is below; everything dropped so far was >= tau
        float tau = 3.0e38f;
        const int m = p.m;
        constexpr bool rmode = RMODE;
        if (rmode) {
            // every point within the radius has a coarse value below (r + pad)^2 + error bound: a fixed threshold
            thr = -3.0e38f;
            if (row < p.Q) {
                const double r = (p.radii ? p.radii[row] : p.radius) + p.radius_pad;
                const float eps = 2.95e-3f * __uint_as_float(p.max_norm_bits[0]) + 1e-4f;
                thr = r >= 0.0 ? fminf(__double2float_ru(r * r) * (1.f + 1e-6f) + eps, TC_THR_MAX) : -3.0e38f;
            }
        }
        // a list is compacted (spilled, radius mode) once it could not take the appends of another step: 16 columns per
        // chunk in the generic epilogue, a group of SC / 4 columns in the fast one
        constexpr int trig = TC_LIST - (NACC > 0 ? SC / 4 : 16);
        // radius mode: a full list goes to the row's global buffer (both halves of a row share it through its counter)
        auto spill = [&]() {
            if (n > 0) {
                const int pos = atomicAdd(&p.cand_cnt[row], n);
                if (pos + n <= p.cap)
                    for (int e = 0; e < n; e++) p.cand_idx[(size_t)row * p.cap + pos + e] = mi[e];
                n = 0;
            }
        };
        // one list at a time, the whole warp compacts the lists that could not take another 16 columns
        auto compact_full_lists = [&]() {
            unsigned need = __ballot_sync(TC_FULL, n > trig);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1u;
                const int cnt = __shfl_sync(TC_FULL, n, src);
                __syncwarp();
                int kept;
                float nt;
#ifdef MRB_TC_TRACE
                tc_compact(wk + (size_t)src * LSTR, wi + (size_t)src * LSTR, cnt, m, TC_SLACK_RUN, trig, __shfl_sync(TC_FULL, thr, src), lane, &kept, &nt,
                           (p.trace && blockIdx.x == 200 && warp == 4) ? p.trace + 18 * TC_TRACE_N : nullptr);
                if (p.trace && blockIdx.x == 200 && warp == 4 && lane == 0) {
                    long long* d = p.trace + 18 * TC_TRACE_N;
                    const int slot = (int)(d[47]++ % 8);
                    for (int q = 0; q < 5; q++) d[5 + slot * 5 + q] = d[q];
                }
#else
                tc_compact(wk + (size_t)src * LSTR, wi + (size_t)src * LSTR, cnt, m, TC_SLACK_RUN, trig, __shfl_sync(TC_FULL, thr, src), lane, &kept, &nt);
#endif
                if (lane == src) {
                    if (kept < 0) { n = 0; thr = -3.0e38f; tau = -3.0e38f; }   // closed: exact fallback for this row
                    else { n = kept; thr = nt; tau = nt; }
                }
                __syncwarp();
            }
        };
        if constexpr (NACC > 0) {
            // ===== fast epilogue: this warp owns columns [half * w, (half + 1) * w) of every tile, w = n_sub * SC =====
            const int w = tn / TC_PARTS, n_sub = w / SC;
            for (int64_t st = 0; st < n_steps; ++st) {
                const int64_t t = st < n_boot ? st : st - n_boot;
                const bool boot = st < n_boot;
                if (st == n_boot && n_boot > 0) {
                    // end of the bootstrap: every list holds n_boot * n_sub batch minima; cut each at its m-th smallest
                    for (int src = 0; src < 32; src++) {
                        const int cnt = __shfl_sync(TC_FULL, n, src);
                        int kept;
                        float nt;
                        tc_compact(wk + (size_t)src * LSTR, wi + (size_t)src * LSTR, cnt, m, TC_SLACK_RUN, TC_LIST, TC_THR_MAX, lane, &kept, &nt);
                        if (lane == src) {
                            if (kept >= 0) { thr = nt; tau = nt; }   // (ties: keep the open threshold, the sweep sorts it out)
                            n = 0;                                    // the minima are found again by the sweep
                        }
                        __syncwarp();
                    }
                }
                const int buf = (int)(st % TC_NBUF);
                const uint32_t use = (uint32_t)(st / TC_NBUF);
                if (lane == 0 && warp == 4) MRB_TRACE(9, t);
                mbar_wait(&tm_full[buf], use & 1);
                if (lane == 0 && warp == 4) MRB_TRACE(4, t);
                tc_fence_after();
                const uint32_t tbase = lane_addr + (uint32_t)(buf * buf_cols + half * w);
                const int64_t colw = (t0 + t) * tn + half * w;
                for (int sub = 0; sub < n_sub; ++sub) {
                    float va[NACC][SC];
#pragma unroll
                    for (int a = 0; a < NACC; a++) tmem_ld_cols<SC>(tbase + (uint32_t)(a * tn + sub * SC), va[a]);
                    tmem_ld_wait();
                    if (lane == 0 && warp == 4) MRB_TRACE(5, t);
                    if (lane == 0 && warp == 11) MRB_TRACE(8, t);
                    if (sub == n_sub - 1) {
                        // everything this warp needs from the buffer is in registers: hand it back to the MMA thread BEFORE
                        // the selection work (a list compaction takes as long as several tiles)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tm_empty[buf]);
                    }
                    float* v = va[0];
                    if constexpr (NACC == 2) {
#pragma unroll
                        for (int j = 0; j < SC; j++) v[j] = fmaxf(v[j], va[1][j]);
                    } else if constexpr (NACC == 3) {
#pragma unroll
                        for (int j = 0; j < SC; j++) v[j] = fmaxf(fmaxf(v[j], va[1][j]), va[2][j]);
                    } else if constexpr (NACC == 4) {
#pragma unroll
                        for (int j = 0; j < SC; j++) v[j] = fmaxf(fmaxf(v[j], va[1][j]), fmaxf(va[2][j], va[3][j]));
                    }
                    // minima per group g = column mod 4, then of the batch
                    float g4[4];
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        float x = v[g];
#pragma unroll
                        for (int j = g + 4; j < SC; j += 4) x = fminf(x, v[j]);
                        g4[g] = x;
                    }
                    const float lo = fminf(fminf(g4[0], g4[1]), fminf(g4[2], g4[3]));
#if defined(MRB_TC_ABL) && MRB_TC_ABL == 1   // experiment build: no candidate path at all (timing only)
                    if (lo < -1.0e30f) mk[0] = lo;
                    continue;
#endif
                    if (boot) {   // bootstrap step: only the batch minimum is kept
                        mk[n] = lo;
                        mi[n] = 0;
                        n++;
                        continue;
                    }
                    if (lane == 0 && warp == 4) MRB_TRACE(6, t);
                    if (!__any_sync(TC_FULL, lo < thr)) continue;
#ifdef MRB_TC_TRACE
                    int tr_groups = 0, tr_comp = 0, tr_lanes = __popc(__ballot_sync(TC_FULL, lo < thr));
#endif
                    // Some lane holds a candidate (two batches in three, mid-sweep: 768 values per vote).  Only warp-uniform
                    // branches from here on -- a group of SC / 4 columns is visited if any lane has a candidate in it, its
                    // appends are predicated: nested per-lane branches cost ~700 cycles per batch for one append (clock64
                    // trace).  Append-only: no ordering, no search; n <= TC_TRIG on entry leaves room for a group (<= 16).
                    // Columns past the end of the corpus carry TC_BIG and thresholds never exceed TC_THR_MAX < TC_BIG.
                    const int cbase = (int)(colw + sub * SC);
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if (__any_sync(TC_FULL, g4[g] < thr)) {
#pragma unroll
                            for (int j = g; j < SC; j += 4) {
                                if (v[j] < thr) {
                                    mk[n] = v[j];
                                    mi[n] = cbase + j;
                                    n++;
                                }
                            }
#ifdef MRB_TC_TRACE
                            tr_groups++;
                            tr_comp += __popc(__ballot_sync(TC_FULL, n > trig));
#endif
                            if (rmode) {
                                if (n > trig) spill();
                            } else {
                                compact_full_lists();
                            }
                        }
                    }
#ifdef MRB_TC_TRACE
                    if (lane == 0 && warp == 4) {
                        MRB_TRACE(16, t);
                        if (p.trace && blockIdx.x == 200 && t >= TC_TRACE_T0 && t < TC_TRACE_T0 + TC_TRACE_N)
                            p.trace[17 * TC_TRACE_N + (t - TC_TRACE_T0)] = tr_lanes | (tr_groups << 8) | (tr_comp << 16);
                    }
#endif
                }
            }
        } else {
            const int chunks_per_tile = tn >> 4;   // 16-column chunks: the loads of up to four robots' accumulators fly together
            for (int64_t t = 0; t < n_tiles; ++t) {
                const int buf = (int)(t % TC_NBUF);
                const uint32_t use = (uint32_t)(t / TC_NBUF);
                mbar_wait(&tm_full[buf], use & 1);
                tc_fence_after();
                const int64_t col0 = (t0 + t) * tn;
                // the two halves take alternate 16-column chunks of every tile; the parity flips from tile to tile so that an
                // odd number of chunks per tile is shared evenly.  (Measured alternative: each half owning every other TILE --
                // half as many barrier hand-offs per warp, but 13.6 instead of 12.1 ms: a tile then occupies its accumulator
                // buffer twice as long and a list compaction stalls the whole tile.)
                for (int ci = (half + (int)(t % TC_PARTS)) % TC_PARTS; ci < chunks_per_tile; ci += TC_PARTS) {
                    const int c = ci << 4;
                    float v[16], b1[16], b2[16], b3[16];
                    const uint32_t taddr = lane_addr + (uint32_t)(buf * buf_cols + c);
                    // issue every accumulator's load before the one wait (n_acc is uniform across the CTA)
                    tmem_ld16(taddr, v);
    #if defined(MRB_TC_ABL) && MRB_TC_ABL == 2   // experiment build: one accumulator read instead of n_acc (timing only)
    #pragma unroll
                    for (int j = 0; j < 16; j++) { b1[j] = v[j]; b2[j] = v[j]; b3[j] = v[j]; }
    #else
                    if (n_acc > 1) tmem_ld16(taddr + (uint32_t)tn, b1);
                    if (n_acc > 2) tmem_ld16(taddr + (uint32_t)(2 * tn), b2);
                    if (n_acc > 3) tmem_ld16(taddr + (uint32_t)(3 * tn), b3);
    #endif
                    tmem_ld_wait();
                    if (n_acc == 2) {
    #pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], b1[j]);
                    } else if (n_acc == 3) {
    #pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = fmaxf(fmaxf(v[j], b1[j]), b2[j]);   // three-input max
                    } else if (n_acc >= 4) {
    #pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = fmaxf(fmaxf(v[j], b1[j]), fmaxf(b2[j], b3[j]));
                    }
                    for (int a = 4; a < n_acc; a += 2) {  // more than four robots: two more at a time
                        tmem_ld16(taddr + (uint32_t)(a * tn), b1);
                        if (a + 1 < n_acc) tmem_ld16(taddr + (uint32_t)((a + 1) * tn), b2);
                        tmem_ld_wait();
                        if (a + 1 < n_acc) {
    #pragma unroll
                            for (int j = 0; j < 16; j++) v[j] = fmaxf(fmaxf(v[j], b1[j]), b2[j]);
                        } else {
    #pragma unroll
                            for (int j = 0; j < 16; j++) v[j] = fmaxf(v[j], b1[j]);
                        }
                    }
                    // after warm-up hardly any chunk holds a candidate: one min tree + one compare per lane, and the
                    // per-column work only when some lane of the warp needs it
                    float lo8[8], lo4[4];
    #pragma unroll
                    for (int j = 0; j < 8; j++) lo8[j] = fminf(v[j], v[8 + j]);
    #pragma unroll
                    for (int j = 0; j < 4; j++) lo4[j] = fminf(lo8[j], lo8[4 + j]);
                    const float lo = fminf(fminf(lo4[0], lo4[1]), fminf(lo4[2], lo4[3]));
    #if defined(MRB_TC_ABL) && MRB_TC_ABL == 1   // experiment build: no candidate path at all (timing only)
                    if (lo < -1.0e30f) mk[0] = lo;
                    continue;
    #endif
                    if (!__any_sync(TC_FULL, lo < thr)) continue;
                    if (lo < thr) {
                        const int64_t cbase = col0 + c;
                        const int lim = (int)min((int64_t)16, p.N - cbase);   // columns of this chunk that exist (last tile)
                        // append-only: no ordering, no search (n <= TC_TRIG here: room for all 16).  lo4[g] is the minimum
                        // of columns g, g + 4, g + 8, g + 12: only groups that hold a candidate look at their members
    #pragma unroll
                        for (int g = 0; g < 4; g++) {
                            if (lo4[g] < thr) {
    #pragma unroll
                                for (int jj = 0; jj < 4; jj++) {
                                    const int j = g + 4 * jj;
                                    if (v[j] < thr && j < lim) {
                                        mk[n] = v[j];
                                        mi[n] = (int)cbase + j;
                                        n++;
                                    }
                                }
                            }
                        }
                    }
                    if (rmode) {
                        if (n > trig) spill();
                        continue;
                    }
                    compact_full_lists();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tm_empty[buf]);
            }
        }
        if (rmode) spill();
        // final compaction to the output size (every list longer than kc), then write-out
        if (!rmode) {
            unsigned need = __ballot_sync(TC_FULL, n > p.kc);
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1u;
                const int cnt = __shfl_sync(TC_FULL, n, src);
                __syncwarp();
                int kept;
                float nt;
                tc_compact(wk + (size_t)src * LSTR, wi + (size_t)src * LSTR, cnt, m, TC_SLACK, trig, __shfl_sync(TC_FULL, thr, src), lane, &kept, &nt);
                if (lane == src) {
                    if (kept < 0 || kept > p.kc) { n = 0; tau = -3.0e38f; }
                    else { n = kept; tau = nt; }
                }
                __syncwarp();
            }
        }
        if (row < p.Q && !rmode) {
            const size_t lst = (size_t)tc_list_id(row, split, half, p.full_qtiles * TC_TM, p.tail_splits);
            float* ok = p.part_key + lst * p.kc;
            int* oi = p.part_idx + lst * p.kc;
            for (int e = 0; e < p.kc; e++) {
                ok[e] = e < n ? mk[e] : 3.0e38f;
                oi[e] = e < n ? mi[e] : -1;
            }
            p.part_tau[lst] = tau;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, alloc_cols);
}

// ---------------------------------------------------------------------------------------------
// exact re-rank + certification: thread = query row
// ---------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(128) knn_rerank_kernel(const double* __restrict__ queries, const double* __restrict__ corpus, int64_t Q,
                                                         int D, const __grid_constant__ Slices sl, int metric, int k, int kc,
                                                         int64_t full_rows, int tail_splits, const float* __restrict__ part_key,
                                                         const int* __restrict__ part_idx,
                                                         const float* __restrict__ part_tau,
                                                         unsigned* __restrict__ max_norm_bits, int* __restrict__ redo_rows,
                                                         int32_t* __restrict__ out_idx,
                                                         double* __restrict__ out_dist, uint8_t* __restrict__ certified) {
    extern __shared__ __align__(16) unsigned char smem_rr[];
    double* hk = reinterpret_cast<double*>(smem_rr);
    int* hi = reinterpret_cast<int*>(hk + (size_t)k * 128);
    const int64_t row = blockIdx.x * (int64_t)128 + threadIdx.x;
    if (row >= Q) return;
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = d < D ? queries[row * D + d] : 0.0;
    ThreadHeap<double> heap(hk + threadIdx.x, hi + threadIdx.x, 128, k, 0);
    float tau = 3.0e38f;  // smallest coarse value any discarded point can have
    const size_t lst0 = (size_t)tc_list_id(row, 0, 0, full_rows, tail_splits);
    const int n_lists = TC_PARTS * (row < full_rows ? 1 : tail_splits);
    for (int s = 0; s < n_lists; s++) {
        const int* pi = part_idx + (lst0 + s) * kc;
        for (int e = 0; e < kc; e++) {
            const int idx = pi[e];
            if (idx < 0) continue;
            const double d = metric_dist<DMAX>(q, corpus + (size_t)idx * D, D, sl, metric);
            if (heap.accepts(d, idx)) heap.push(d, idx);
        }
        tau = fminf(tau, part_tau[lst0 + s]);   // every point of list s's corpus part outside the list has a coarse value >= its tau
    }
    // error bound of a coarse value (single TF32 rounding of every operand entry, exact products, FP32 accumulation):
    // (2^-9 + 2^-10) * M for the cross and norm terms, M = largest squared slice norm of any query / corpus row
    const float eps = 2.95e-3f * __uint_as_float(max_norm_bits[0]) + 1e-4f;
    bool ok = true;
    if (heap.n == k) {
        const double dk = heap.top_key();
        ok = (float)(dk * dk) < tau - eps;
    } else {
        ok = tau > 1.0e38f;  // fewer than k found: fine only if nothing was discarded anywhere
    }
    certified[row] = ok ? 1 : 0;
    if (!ok) redo_rows[atomicAdd(&max_norm_bits[1], 1u)] = (int)row;   // rows handed to the exact per-row kernel below
    const int found = heap.n;
    for (int e = found - 1; e >= 0; e--) {
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
// exact k-NN of the few rows the re-rank could not certify: one CTA per row, all 256 threads share the corpus
// (thread-local bounded heaps of the k best, then a k-round block-wide merge by (distance, index)).  The streaming
// exact kernel (knn_kernels.cu) maps one THREAD to a row: a single uncertified row would crawl through the whole
// corpus on one thread (150 ms at N = 100 000) while its CTA neighbours idle.
// ---------------------------------------------------------------------------------------------
constexpr int TC_REDO_PARTS = 32;     // CTAs sharing one uncertified row
constexpr int TC_REDO_FAST = 1024;    // rows that get the multi-CTA treatment; any further rows take one CTA each

// parts > 1: CTA (x, y) handles rows x, x + gridDim.x, ... < min(total, TC_REDO_FAST) and the y-th part of the corpus; its k
// best go to part_d / part_i [row slot][part][k] for knn_rows_merge_kernel.  parts == 1: rows TC_REDO_FAST + x, ... and the
// whole corpus, written straight to the output.
template <int DMAX>
__global__ void __launch_bounds__(256) knn_rows_kernel(const double* __restrict__ queries, const double* __restrict__ corpus, int64_t N,
                                                       int D, const __grid_constant__ Slices sl, int metric, int k,
                                                       const int* __restrict__ rows, const unsigned* __restrict__ n_rows,
                                                       int parts, double* __restrict__ part_d, int* __restrict__ part_i,
                                                       int32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                                       uint8_t* __restrict__ certified) {
    extern __shared__ __align__(16) unsigned char smem_kr[];
    double* hk = reinterpret_cast<double*>(smem_kr);                 // [k][256]
    int* hi = reinterpret_cast<int*>(hk + (size_t)k * 256);          // [k][256]
    double* red_d = reinterpret_cast<double*>(hi + (size_t)k * 256); // [8]
    int* red_i = reinterpret_cast<int*>(red_d + 8);                  // [8] index, [8] owner thread
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned total = parts > 1 ? min(*n_rows, (unsigned)TC_REDO_FAST) : *n_rows;
    const int64_t plen = (N + parts - 1) / parts;
    const int64_t p0 = (int64_t)blockIdx.y * plen, p1 = min(N, p0 + plen);
    for (unsigned r = (parts > 1 ? 0u : (unsigned)TC_REDO_FAST) + blockIdx.x; r < total; r += gridDim.x) {
        const int64_t row = rows[r];
        double q[DMAX];
#pragma unroll
        for (int d = 0; d < DMAX; d++) q[d] = d < D ? queries[row * D + d] : 0.0;
        ThreadHeap<double> heap(hk + tid, hi + tid, 256, k, 0);
        for (int64_t i = p0 + tid; i < p1; i += 256) {
            const double d = metric_dist<DMAX>(q, corpus + (size_t)i * D, D, sl, metric);
            if (heap.accepts(d, (int)i)) heap.push(d, (int)i);
        }
        // ascending order in place: pop the maximum into the last free slot
        const int mine = heap.n;
        for (int e = mine - 1; e >= 0; e--) {
            double d;
            int i;
            heap.pop(&d, &i);
            hk[(size_t)e * 256 + tid] = d;
            hi[(size_t)e * 256 + tid] = i;
        }
        int head = 0;
        __syncthreads();
        for (int e = 0; e < k; e++) {
            double bd = head < mine ? hk[(size_t)head * 256 + tid] : __longlong_as_double(0x7ff0000000000000LL);
            int bi = head < mine ? hi[(size_t)head * 256 + tid] : 0x7fffffff;
            int bt = tid;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o), ot = __shfl_xor_sync(0xffffffffu, bt, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bt = ot; }
            }
            if (lane == 0) { red_d[warp] = bd; red_i[warp] = bi; red_i[8 + warp] = bt; }
            __syncthreads();
            bd = red_d[0]; bi = red_i[0]; bt = red_i[8];
#pragma unroll
            for (int w = 1; w < 8; w++) {
                const double od = red_d[w];
                const int oi = red_i[w];
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bt = red_i[8 + w]; }
            }
            if (tid == 0) {
                const bool found = bi != 0x7fffffff;
                if (parts > 1) {
                    const size_t o = ((size_t)r * parts + blockIdx.y) * k + e;
                    part_d[o] = found ? bd : __longlong_as_double(0x7ff0000000000000LL);
                    part_i[o] = found ? bi : 0x7fffffff;
                } else {
                    out_idx[row * k + e] = found ? bi : -1;
                    if (out_dist) out_dist[row * k + e] = found ? bd : __longlong_as_double(0x7ff0000000000000LL);
                }
            }
            if (tid == bt) head++;
            __syncthreads();
        }
        if (tid == 0 && parts == 1) certified[row] = 1;
    }
}

// k-way merge of the TC_REDO_PARTS sorted partial lists of a row: one warp per row, lane = part
__global__ void __launch_bounds__(32) knn_rows_merge_kernel(const int* __restrict__ rows, const unsigned* __restrict__ n_rows, int k,
                                                            const double* __restrict__ part_d, const int* __restrict__ part_i,
                                                            int32_t* __restrict__ out_idx, double* __restrict__ out_dist,
                                                            uint8_t* __restrict__ certified) {
    const unsigned total = min(*n_rows, (unsigned)TC_REDO_FAST);
    const int lane = threadIdx.x;
    for (unsigned r = blockIdx.x; r < total; r += gridDim.x) {
        const int64_t row = rows[r];
        const double* pd = part_d + ((size_t)r * TC_REDO_PARTS + lane) * k;
        const int* pi = part_i + ((size_t)r * TC_REDO_PARTS + lane) * k;
        int head = 0;
        for (int e = 0; e < k; e++) {
            double bd = head < k ? pd[head] : __longlong_as_double(0x7ff0000000000000LL);
            int bi = head < k ? pi[head] : 0x7fffffff;
            int bl = lane;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bl = ol; }
            }
            if (lane == 0) {
                const bool found = bi != 0x7fffffff;
                out_idx[row * k + e] = found ? bi : -1;
                if (out_dist) out_dist[row * k + e] = found ? bd : __longlong_as_double(0x7ff0000000000000LL);
            }
            if (lane == bl) head++;
        }
        if (lane == 0) certified[row] = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// r-disc search on the candidate generator: exact fp64 filter of every candidate (thread = query row), kept indices
// compacted to the front of the row's buffer and sorted ascending (the reference's np.where order), then a fill pass
// ---------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(128) radius_tc_filter_kernel(const double* __restrict__ queries, const double* __restrict__ corpus,
                                                               int64_t Q, int D, const __grid_constant__ Slices sl, int metric,
                                                               const double* __restrict__ radii, double radius, int inclusive, int cap,
                                                               int* __restrict__ cand_idx, int* __restrict__ cand_cnt,
                                                               int64_t* __restrict__ counts) {
    const int64_t row = blockIdx.x * (int64_t)128 + threadIdx.x;
    if (row >= Q) return;
    const int cnt = cand_cnt[row];
    if (cnt > cap) {          // more candidates than the buffer holds: the caller answers this row with the exact kernels
        counts[row] = -1;
        return;
    }
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = d < D ? queries[row * D + d] : 0.0;
    const double r = radii ? radii[row] : radius;
    const double lim = inclusive ? r + 1e-10 : r;
    int* buf = cand_idx + (size_t)row * cap;
    int kept = 0;
    for (int e = 0; e < cnt; e++) {
        const int idx = buf[e];
        const double d = metric_dist<DMAX>(q, corpus + (size_t)idx * D, D, sl, metric);
        const bool in = inclusive ? d <= lim : d < lim;
        if (in) {   // insertion into the sorted prefix (candidates arrive nearly sorted: one list per column half)
            int j = kept++;
            while (j > 0 && buf[j - 1] > idx) { buf[j] = buf[j - 1]; j--; }
            buf[j] = idx;
        }
    }
    cand_cnt[row] = kept;
    counts[row] = kept;
}

// The same filter with a WARP per query row (caps up to RF_CAP): lane = candidate for the fp64 predicate, survivors
// compacted in arrival order into shared memory (ballot prefix), a warp-synchronous bitonic sort by index, write-back to
// the front of the row's buffer.  The thread-per-row kernel above insertion-sorts in global memory with one lane of 32
// doing anything: 5.0-5.6 ms for 100 000 rows of ~50 candidates, as long as the candidate generator itself.
constexpr int RF_CAP = 1024, RF_WARPS = 8;
template <int DMAX>
__global__ void __launch_bounds__(32 * RF_WARPS) radius_tc_filter_warp_kernel(const double* __restrict__ queries, const double* __restrict__ corpus,
                                                                              int64_t Q, int D, const __grid_constant__ Slices sl, int metric,
                                                                              const double* __restrict__ radii, double radius, int inclusive,
                                                                              int cap, int* __restrict__ cand_idx, int* __restrict__ cand_cnt,
                                                                              int64_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smem_rf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* sidx = reinterpret_cast<int*>(smem_rf) + (size_t)warp * cap;
    const int64_t row = blockIdx.x * (int64_t)RF_WARPS + warp;
    if (row >= Q) return;
    const int cnt = cand_cnt[row];
    if (cnt > cap) {          // more candidates than the buffer holds: the caller answers this row with the exact kernels
        if (lane == 0) counts[row] = -1;
        return;
    }
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = d < D ? queries[row * D + d] : 0.0;
    const double r = radii ? radii[row] : radius;
    const double lim = inclusive ? r + 1e-10 : r;
    int* buf = cand_idx + (size_t)row * cap;
    int kept = 0;
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        bool in = false;
        int idx = 0;
        if (e < cnt) {
            idx = buf[e];
            const double d = metric_dist<DMAX>(q, corpus + (size_t)idx * D, D, sl, metric);
            in = inclusive ? d <= lim : d < lim;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (in) sidx[kept + __popc(bal & ((1u << lane) - 1u))] = idx;
        kept += __popc(bal);
    }
    int n2 = 32;
    while (n2 < kept) n2 <<= 1;
    for (int e = kept + lane; e < n2; e += 32) sidx[e] = 0x7fffffff;
    __syncwarp();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (n2 >> 1); t += 32) {
                const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));   // lower index of the compare pair
                const int j = i | stride;
                const bool up = (i & size) == 0;
                const int a = sidx[i], b = sidx[j];
                if ((a > b) == up) { sidx[i] = b; sidx[j] = a; }
            }
            __syncwarp();
        }
    }
    for (int e = lane; e < kept; e += 32) buf[e] = sidx[e];
    if (lane == 0) {
        cand_cnt[row] = kept;
        counts[row] = kept;
    }
}

template <int DMAX>
__global__ void __launch_bounds__(128) radius_tc_fill_kernel(const double* __restrict__ queries, const double* __restrict__ corpus, int64_t Q,
                                                             int D, const __grid_constant__ Slices sl, int metric, int cap,
                                                             const int* __restrict__ cand_idx, const int* __restrict__ cand_cnt,
                                                             const int64_t* __restrict__ offsets, int32_t* __restrict__ out_idx,
                                                             double* __restrict__ out_dist) {
    const int64_t row = blockIdx.x * (int64_t)128 + threadIdx.x;
    if (row >= Q) return;
    const int cnt = cand_cnt[row];
    if (cnt > cap) return;    // overflowed row: filled by the caller
    double q[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; d++) q[d] = d < D ? queries[row * D + d] : 0.0;
    const int* buf = cand_idx + (size_t)row * cap;
    const int64_t o = offsets[row];
    for (int e = 0; e < cnt; e++) {
        const int idx = buf[e];
        out_idx[o + e] = idx;
        if (out_dist) out_dist[o + e] = metric_dist<DMAX>(q, corpus + (size_t)idx * D, D, sl, metric);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int pad8(int x) { return (x + 7) / 8 * 8; }

bool knn_tc_make_plan(int D, const Slices& sl, int metric, TcPlan* plan) {
    plan->n_acc = 0;
    plan->KS = 0;
    if (metric == METRIC_EUCLIDEAN) {
        plan->n_acc = 1;
        plan->start[0] = 0;
        plan->dim[0] = D;
    } else if (metric == METRIC_MAX_EUCLIDEAN) {
        if (sl.R > 8) return false;
        plan->n_acc = sl.R;
        for (int r = 0; r < sl.R; r++) { plan->start[r] = sl.start[r]; plan->dim[r] = sl.end[r] - sl.start[r]; }
    } else {
        return false;
    }
    for (int a = 0; a < plan->n_acc; a++) {
        plan->kstep0[a] = plan->KS;
        plan->ksteps[a] = pad8(plan->dim[a] + 2) / 8;
        plan->KS += plan->ksteps[a];
    }
    if (plan->KS > TC_MAX_KS) return false;
    // tensor memory: TC_NBUF accumulator buffers of n_acc x tn columns and the query operand (8 columns per K step) in 512
    // columns; the UMMA N of a 128-row instruction is a multiple of 16
    const int tn_max = (512 - 8 * plan->KS) / (TC_NBUF * plan->n_acc) / 16 * 16;
    // shared memory: the candidate lists take TC_PARTS * 128 * (TC_LIST + 1) * 8 bytes; the corpus stages share the rest
    const long avail = 224L * 1024 - (long)TC_PARTS * TC_TM * (TC_LIST + 1) * 8 - 4096;
    // fast epilogue (<= 4 accumulators): every warp reads sc columns of each accumulator at once (n_acc * sc <= 96 registers);
    // tn = a multiple of TC_PARTS * sc
    static const int sc_of[5] = {0, 48, 48, 32, 24};
    plan->sc = 0;
    if (plan->n_acc <= 4 && MRB_TC_FAST_EPILOGUE) {
        const int sc = sc_of[plan->n_acc], unit = TC_PARTS * sc;
        int tn = tn_max / unit * unit;
        if (tn > 256) tn = 256 / unit * unit;
        while (tn >= unit && (tn % 16 != 0 || avail / ((long)plan->KS * tn * 32) < 2)) tn -= unit;
        if (tn >= unit) {
            const int stages = (int)(avail / ((long)plan->KS * tn * 32));
            plan->tn = tn;
            plan->sc = sc;
            plan->stages = stages > TC_STAGES ? TC_STAGES : stages;
            return true;
        }
    }
    int tn = tn_max > 256 ? 256 : tn_max;
    int stages = 0;
    for (; tn >= 16; tn -= 16) {
        stages = (int)(avail / ((long)plan->KS * tn * 32));
        if (stages >= 2) break;
    }
    if (tn < 16) return false;
    plan->tn = tn;
    plan->stages = stages > TC_STAGES ? TC_STAGES : stages;
    return true;
}

// entries a compaction keeps per list: the k nearest spread over n_lists lists binomially -- mean + 4 sigma + 2
int knn_tc_keep(int k, int n_lists) {
    const double pr = 1.0 / n_lists;
    int m = (int)(k * pr + 4.0 * sqrt(k * pr * (1.0 - pr)) + 2.999);
    if (m > k + 8) m = k + 8;
    if (m > TC_TRIG - 16) m = TC_TRIG - 16;
    return m < 1 ? 1 : m;
}
int knn_tc_slots(int k, int n_lists) { return knn_tc_keep(k, n_lists) + TC_SLACK; }
int knn_tc_max_k() {   // largest k whose per-list keep count (TC_PARTS lists at least) still leaves room for a chunk
    int k = 1;
    while (k < 128) {
        const double pr = 1.0 / TC_PARTS;
        const int m = (int)((k + 1) * pr + 4.0 * sqrt((k + 1) * pr * (1.0 - pr)) + 2.999);
        if ((m > k + 9 ? k + 9 : m) > TC_TRIG - 16) break;
        k++;
    }
    return k;
}
int knn_tc_parts() { return TC_PARTS; }

size_t knn_tc_smem_bytes(const TcPlan& plan, int /*kc*/) {
    return (size_t)plan.stages * plan.KS * plan.tn * 32 + (size_t)TC_PARTS * TC_TM * (TC_LIST + 1) * 8 + 16 + 16 * 8 + (size_t)TC_MAX_KS * 16 + 1024;
}

// Work decomposition of a launch (TcParams::full_qtiles / tail_splits).  A CTA sweeps the corpus for one query tile; the
// grid runs in waves of one CTA per SM, so the query tiles of the last, partial wave are split along the corpus into as
// many pieces as there are SMs left over (782 query tiles on 148 SMs: 740 full sweeps + 42 tiles x 3 splits -- 5.33
// sweep times instead of 6).  Query sets with fewer tiles than SMs are split altogether.
void knn_tc_shape(int64_t Q, int64_t n_ctiles, int64_t* full_qtiles, int* tail_splits) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms > TC_MAX_SMS) sms = TC_MAX_SMS;
    const int64_t qt = (Q + TC_TM - 1) / TC_TM;
    const int64_t rem = qt % sms;
    int64_t s = rem ? sms / rem : 1;
    const int64_t max_s = (n_ctiles + 15) / 16;      // at least 16 corpus tiles per split
    if (s > max_s) s = max_s;
    if (s > 32) s = 32;
    if (s < 1) s = 1;
    *tail_splits = (int)s;
    *full_qtiles = s > 1 ? qt - rem : qt;
}
// candidate lists a launch writes, at most (workspace sizing)
int64_t knn_tc_max_lists(int64_t Q) { return (int64_t)TC_PARTS * (Q + (int64_t)TC_TM * (TC_MAX_SMS + 1)); }
// lists per row, at least: what the per-list keep count is derived from
int knn_tc_min_lists(int64_t Q, int64_t n_ctiles) {
    int64_t full;
    int s;
    knn_tc_shape(Q, n_ctiles, &full, &s);
    return TC_PARTS * (full > 0 ? 1 : s);
}

size_t knn_tc_workspace_bytes(int64_t Q, int64_t N, const TcPlan& plan, int kc) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    return 256 + (size_t)qt * plan.KS * TC_TM * 32 + (size_t)ct * plan.KS * plan.tn * 32 + (size_t)knn_tc_max_lists(Q) * (kc * 8 + 4) + (size_t)Q * 4 + 16 + (size_t)TC_REDO_FAST * TC_REDO_PARTS * (kc + 8) * 12 + 1024;
}

// the candidate generator for the plan's epilogue shape
template <bool RMODE>
static cudaError_t launch_tc_generator(const TcParams& p, unsigned grid, size_t smem, cudaStream_t st) {
#define MRB_TC_GEN(NA, SC)                                                                                                  \
    do {                                                                                                                    \
        cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<RMODE, NA, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                                     \
        knn_tc_kernel<RMODE, NA, SC><<<grid, TC_THREADS, smem, st>>>(p);                                                    \
        return cudaGetLastError();                                                                                          \
    } while (0)
    if (p.plan.sc == 0) MRB_TC_GEN(0, 16);
    switch (p.plan.n_acc) {
        case 1: MRB_TC_GEN(1, 48);
        case 2: MRB_TC_GEN(2, 48);
        case 3: MRB_TC_GEN(3, 32);
        case 4: MRB_TC_GEN(4, 24);
    }
#undef MRB_TC_GEN
    return cudaErrorInvalidValue;
}

cudaError_t launch_knn_tc(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                          int k, int kc, const TcPlan& plan, void* workspace, int32_t* out_idx, double* out_dist,
                          uint8_t* certified, cudaStream_t st) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    unsigned char* w = (unsigned char*)workspace;
    unsigned* max_norm = (unsigned*)w;
    w += 256;
    float* A = (float*)w;
    w += (size_t)qt * plan.KS * TC_TM * 32;
    float* B = (float*)w;
    w += (size_t)ct * plan.KS * plan.tn * 32;
    const size_t max_lists = (size_t)knn_tc_max_lists(Q);
    float* part_key = (float*)w;
    w += max_lists * kc * 4;
    int* part_idx = (int*)w;
    w += max_lists * kc * 4;
    float* part_tau = (float*)w;
    w += max_lists * 4;
    int* redo_rows = (int*)w;
    w += ((size_t)Q * 4 + 15) / 16 * 16;
    double* redo_d = (double*)w;
    w += (size_t)TC_REDO_FAST * TC_REDO_PARTS * k * 8;
    int* redo_i = (int*)w;
    cudaError_t e = cudaMemsetAsync(max_norm, 0, 8, st);   // [0] largest squared slice norm, [1] rows handed to the exact kernel
    if (e != cudaSuccess) return e;
    knn_tc_prep_kernel<<<(unsigned)((qt * TC_TM + 127) / 128), 128, 0, st>>>(queries, Q, qt * TC_TM, D, plan, 0, TC_TM, A, max_norm);
    knn_tc_prep_kernel<<<(unsigned)((ct * plan.tn + 127) / 128), 128, 0, st>>>(corpus, N, ct * plan.tn, D, plan, 1, plan.tn, B, max_norm);
    TcParams p{};
    p.A = A;
    p.B = B;
    p.Q = Q;
    p.N = N;
    p.n_ctiles = ct;
    knn_tc_shape(Q, ct, &p.full_qtiles, &p.tail_splits);
    p.m = kc - TC_SLACK;
    p.kc = kc;
    p.part_key = part_key;
    p.part_idx = part_idx;
    p.part_tau = part_tau;
    p.plan = plan;
    const unsigned grid = (unsigned)(p.full_qtiles + (qt - p.full_qtiles) * p.tail_splits);
#ifdef MRB_TC_TRACE
    static long long* trace_dev = nullptr;
    if (!trace_dev) cudaMalloc(&trace_dev, sizeof(long long) * TC_TRACE_EV * TC_TRACE_N);
    cudaMemsetAsync(trace_dev, 0, sizeof(long long) * TC_TRACE_EV * TC_TRACE_N, st);
    p.trace = trace_dev;
#endif
    e = launch_tc_generator<false>(p, grid, knn_tc_smem_bytes(plan, kc), st);
    if (e != cudaSuccess) return e;
#ifdef MRB_TC_TRACE
    {
        static long long h[TC_TRACE_EV * TC_TRACE_N];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
        const long long base = h[0];
        if (base) {
            fprintf(stderr, "TRACE tile: mma_start mma_gotempty mma_gotB mma_committed | epi_waitfull epi_gotfull epi_loaded(w4) epi_loaded(w11) | prod_gotempty\n");
            for (int i = 0; i < TC_TRACE_N; i++)
                fprintf(stderr, "TRACE %4d: %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld | %7lld\n", TC_TRACE_T0 + i, h[0 * TC_TRACE_N + i] - base,
                        h[1 * TC_TRACE_N + i] - base, h[2 * TC_TRACE_N + i] - base, h[3 * TC_TRACE_N + i] - base, h[9 * TC_TRACE_N + i] - base,
                        h[4 * TC_TRACE_N + i] - base, h[5 * TC_TRACE_N + i] - base, h[8 * TC_TRACE_N + i] - base, h[7 * TC_TRACE_N + i] - base);
            fprintf(stderr, "TRACE3 tile: warp 4: gotfull->loaded, loaded->voted, voted->appended (0 = no hit), lanes, groups, compactions, tile period\n");
            for (int i = 0; i + 1 < TC_TRACE_N; i++) {
                const long long v = h[17 * TC_TRACE_N + i];
                fprintf(stderr, "TRACE3 %4d: %5lld %5lld %5lld  lanes %2lld groups %lld comp %lld  period %5lld\n", TC_TRACE_T0 + i,
                        h[5 * TC_TRACE_N + i] - h[4 * TC_TRACE_N + i], h[6 * TC_TRACE_N + i] - h[5 * TC_TRACE_N + i],
                        h[16 * TC_TRACE_N + i] ? h[16 * TC_TRACE_N + i] - h[6 * TC_TRACE_N + i] : 0, v & 255, (v >> 8) & 255, v >> 16,
                        h[4 * TC_TRACE_N + i + 1] - h[4 * TC_TRACE_N + i]);
            }
            fprintf(stderr, "TRACE4 last compactions of warp 4: load+minmax, search, move (cycles), probes, entries\n");
            for (int i = 0; i < 8; i++) {
                const long long* d = h + 18 * TC_TRACE_N + 5 + i * 5;
                fprintf(stderr, "TRACE4 %lld %lld %lld probes %lld n %lld\n", d[0], d[1], d[2], d[3], d[4]);
            }
            fprintf(stderr, "TRACE2 tile: gotB mma0 mma1 mma2 mma3 commit0 commit1 (relative to gotB)\n");
            for (int i = 0; i < TC_TRACE_N; i++) {
                const long long b2 = h[2 * TC_TRACE_N + i];
                fprintf(stderr, "TRACE2 %4d: %5lld %5lld %5lld %5lld %5lld %5lld\n", TC_TRACE_T0 + i, h[10 * TC_TRACE_N + i] - b2, h[11 * TC_TRACE_N + i] - b2,
                        h[12 * TC_TRACE_N + i] - b2, h[13 * TC_TRACE_N + i] - b2, h[14 * TC_TRACE_N + i] - b2, h[15 * TC_TRACE_N + i] - b2);
            }
        }
    }
#endif
    const size_t rsmem = (size_t)k * 128 * 12;
#define MRB_RERANK(DM)                                                                                                            \
    do {                                                                                                                          \
        e = cudaFuncSetAttribute(knn_rerank_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);                 \
        if (e != cudaSuccess) return e;                                                                                           \
        knn_rerank_kernel<DM><<<(unsigned)((Q + 127) / 128), 128, rsmem, st>>>(queries, corpus, Q, D, sl, metric, k, kc, p.full_qtiles * TC_TM, p.tail_splits,  \
                                                                               part_key, part_idx, part_tau, max_norm, redo_rows, out_idx, out_dist,  \
                                                                               certified);                                        \
    } while (0)
    // rows without a certificate: exact answer from the per-row kernel (a fixed small grid; CTAs beyond the list exit)
    const size_t ksmem = (size_t)k * 256 * 12 + 8 * 8 + 16 * 4;
#define MRB_ROWS(DM)                                                                                                              \
    do {                                                                                                                          \
        e = cudaFuncSetAttribute(knn_rows_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksmem);                   \
        if (e != cudaSuccess) return e;                                                                                           \
        knn_rows_kernel<DM><<<dim3(64, TC_REDO_PARTS), 256, ksmem, st>>>(queries, corpus, N, D, sl, metric, k, redo_rows,        \
                                                                         max_norm + 1, TC_REDO_PARTS, redo_d, redo_i, out_idx,    \
                                                                         out_dist, certified);                                    \
        knn_rows_merge_kernel<<<64, 32, 0, st>>>(redo_rows, max_norm + 1, k, redo_d, redo_i, out_idx, out_dist, certified);       \
        if (Q > TC_REDO_FAST)                                                                                                     \
            knn_rows_kernel<DM><<<dim3(296, 1), 256, ksmem, st>>>(queries, corpus, N, D, sl, metric, k, redo_rows, max_norm + 1,  \
                                                                  1, nullptr, nullptr, out_idx, out_dist, certified);             \
    } while (0)
    if (D <= 8) { MRB_RERANK(8); MRB_ROWS(8); }
    else if (D <= 16) { MRB_RERANK(16); MRB_ROWS(16); }
    else if (D <= 24) { MRB_RERANK(24); MRB_ROWS(24); }
    else if (D <= 32) { MRB_RERANK(32); MRB_ROWS(32); }
    else { MRB_RERANK(64); MRB_ROWS(64); }
#undef MRB_RERANK
#undef MRB_ROWS
    return cudaGetLastError();
}

size_t radius_tc_workspace_bytes(int64_t Q, int64_t N, const TcPlan& plan, int cap) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    return 256 + (size_t)qt * plan.KS * TC_TM * 32 + (size_t)ct * plan.KS * plan.tn * 32 + (size_t)Q * 4 + 16 + (size_t)Q * cap * 4 + 1024;
}

// operands, candidate generator in radius mode, exact filter -> counts[Q] (-1: the row overflowed its candidate buffer)
cudaError_t launch_radius_tc_count(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                                   const double* radii, double radius, int inclusive, int cap, const TcPlan& plan, void* workspace,
                                   int64_t* counts, cudaStream_t st) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    unsigned char* w = (unsigned char*)workspace;
    unsigned* max_norm = (unsigned*)w;
    w += 256;
    float* A = (float*)w;
    w += (size_t)qt * plan.KS * TC_TM * 32;
    float* B = (float*)w;
    w += (size_t)ct * plan.KS * plan.tn * 32;
    int* cand_cnt = (int*)w;
    w += ((size_t)Q * 4 + 15) / 16 * 16;
    int* cand_idx = (int*)w;
    cudaError_t e = cudaMemsetAsync(max_norm, 0, 8, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(cand_cnt, 0, (size_t)Q * 4, st);
    if (e != cudaSuccess) return e;
    knn_tc_prep_kernel<<<(unsigned)((qt * TC_TM + 127) / 128), 128, 0, st>>>(queries, Q, qt * TC_TM, D, plan, 0, TC_TM, A, max_norm);
    knn_tc_prep_kernel<<<(unsigned)((ct * plan.tn + 127) / 128), 128, 0, st>>>(corpus, N, ct * plan.tn, D, plan, 1, plan.tn, B, max_norm);
    TcParams p{};
    p.A = A;
    p.B = B;
    p.Q = Q;
    p.N = N;
    p.n_ctiles = ct;
    knn_tc_shape(Q, ct, &p.full_qtiles, &p.tail_splits);
    p.m = 1;
    p.kc = 1;
    p.radius_mode = 1;
    p.radii = radii;
    p.radius = radius;
    p.radius_pad = inclusive ? 1e-10 : 0.0;
    p.max_norm_bits = max_norm;
    p.cand_idx = cand_idx;
    p.cand_cnt = cand_cnt;
    p.cap = cap;
    p.plan = plan;
    const unsigned grid = (unsigned)(p.full_qtiles + (qt - p.full_qtiles) * p.tail_splits);
    e = launch_tc_generator<true>(p, grid, knn_tc_smem_bytes(plan, 0), st);
    if (e != cudaSuccess) return e;
#define MRB_RFILTER(DM)                                                                                                             \
    radius_tc_filter_kernel<DM><<<(unsigned)((Q + 127) / 128), 128, 0, st>>>(queries, corpus, Q, D, sl, metric, radii, radius, inclusive, \
                                                                             cap, cand_idx, cand_cnt, counts)
#define MRB_RFILTER_W(DM)                                                                                                           \
    do {                                                                                                                            \
        const size_t fsmem = (size_t)RF_WARPS * cap * 4;                                                                            \
        e = cudaFuncSetAttribute(radius_tc_filter_warp_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);        \
        if (e != cudaSuccess) return e;                                                                                             \
        radius_tc_filter_warp_kernel<DM><<<(unsigned)((Q + RF_WARPS - 1) / RF_WARPS), 32 * RF_WARPS, fsmem, st>>>(                  \
            queries, corpus, Q, D, sl, metric, radii, radius, inclusive, cap, cand_idx, cand_cnt, counts);                          \
    } while (0)
    if (cap <= RF_CAP) {
        if (D <= 8) MRB_RFILTER_W(8);
        else if (D <= 16) MRB_RFILTER_W(16);
        else if (D <= 24) MRB_RFILTER_W(24);
        else if (D <= 32) MRB_RFILTER_W(32);
        else MRB_RFILTER_W(64);
    } else if (D <= 8) MRB_RFILTER(8);
    else if (D <= 16) MRB_RFILTER(16);
    else if (D <= 24) MRB_RFILTER(24);
    else if (D <= 32) MRB_RFILTER(32);
    else MRB_RFILTER(64);
#undef MRB_RFILTER
#undef MRB_RFILTER_W
    return cudaGetLastError();
}

cudaError_t launch_radius_tc_fill(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                                  int cap, const TcPlan& plan, void* workspace, const int64_t* offsets, int32_t* out_idx, double* out_dist,
                                  cudaStream_t st) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    unsigned char* w = (unsigned char*)workspace + 256 + (size_t)qt * plan.KS * TC_TM * 32 + (size_t)ct * plan.KS * plan.tn * 32;
    const int* cand_cnt = (const int*)w;
    const int* cand_idx = (const int*)(w + ((size_t)Q * 4 + 15) / 16 * 16);
#define MRB_RFILL(DM)                                                                                                              \
    radius_tc_fill_kernel<DM><<<(unsigned)((Q + 127) / 128), 128, 0, st>>>(queries, corpus, Q, D, sl, metric, cap, cand_idx, cand_cnt,   \
                                                                           offsets, out_idx, out_dist)
    if (D <= 8) MRB_RFILL(8);
    else if (D <= 16) MRB_RFILL(16);
    else if (D <= 24) MRB_RFILL(24);
    else if (D <= 32) MRB_RFILL(32);
    else MRB_RFILL(64);
#undef MRB_RFILL
    return cudaGetLastError();
}

}  // namespace mrb
