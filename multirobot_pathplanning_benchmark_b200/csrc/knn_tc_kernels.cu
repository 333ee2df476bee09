// Tensor-core candidate generator for k-NN (euclidean / max_euclidean) + exact fp64 re-rank.
//
// Replaces the reference's per-query loop  dists = batch_dist_fun(q, corpus); argpartition; argsort
// (P/planners/prm/prm_graph.py:407-447, distances P/problems/core/configuration.py:303-329) for
// whole query batches.
//
// Per robot slice the squared distance is a dense contraction of augmented vectors
//   A-row (query):  [ q_hi  q_hi  q_lo | |q|^2_hi |q|^2_lo 1 1 | 0.. ]
//   B-row (corpus): [-2c_hi -2c_lo -2c_hi | 1 1 |c|^2_hi |c|^2_lo | 0.. ]
// where x_hi = tf32(x), x_lo = tf32(x - x_hi) (3xTF32 split, every entry exactly representable in
// TF32).  knn_tc_prep_kernel writes both operands in the UMMA canonical K-major, no-swizzle
// layout (8-row x 16-byte core matrices; LBO = 128 B between the two K halves of an instruction,
// SBO = 256 B between 8-row groups), tile after tile, so a whole operand tile is ONE contiguous
// bulk copy.  knn_tc_kernel: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (M = 128 query
// rows, N = corpus columns per accumulator, K = 8 per instruction, kind::tf32, FP32 accumulators
// in TMEM, one accumulator per robot, double buffered), warp 2 = TMEM allocator, warps 4-7 =
// epilogue: tcgen05.ld of 32 columns per robot, max over robots, per-row bounded heap of
// k + slack candidates in shared memory.  knn_rerank_kernel recomputes the candidates' distances
// in fp64 with the reference's operand order, sorts by (distance, index) and certifies each row:
// if the exact k-th squared distance is not below (smallest discarded coarse value - error bound)
// the row is flagged and the caller recomputes it with the exact kernel (knn_kernels.cu).
#include <cuda_runtime.h>
#include <stdint.h>

#include "async_copy.cuh"
#include "kernels.h"
#include "knn_common.cuh"

namespace mrb {

constexpr int TC_TM = 128;        // query rows per CTA (UMMA M)
constexpr int TC_STAGES = 3;      // shared-memory stages of corpus tiles
constexpr int TC_THREADS = 384;   // warps 0-3: producer / MMA / TMEM allocator / spare, warps 4-11: epilogue
constexpr int TC_STG = 8;         // staged candidates per epilogue thread between batched heap updates
constexpr float TC_BIG = 1.0e30f;
constexpr int TC_MAX_KS = 40;     // knn_tc_make_plan accepts plans up to this many K steps

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

// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, TF32 in, FP32 out
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
// arrive on an mbarrier once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive FP32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

// ---------------------------------------------------------------------------------------------
// operand layout
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t tc_elem_offset(int rows_per_tile, int KS, int64_t row, int kglob) {
    // float index of element (row, kglob) of an operand stored tile after tile, K step after K step
    const int64_t tile = row / rows_per_tile;
    const int m = (int)(row - tile * rows_per_tile);
    const int ks = kglob >> 3, k = kglob & 7;
    return (tile * KS + ks) * (int64_t)rows_per_tile * 8 + (m >> 3) * 64 + (k >> 2) * 32 + (m & 7) * 4 + (k & 3);
}

// one thread per (padded) row; side 0 = queries (A), 1 = corpus (B)
__global__ void __launch_bounds__(128) knn_tc_prep_kernel(const double* __restrict__ X, int64_t n, int64_t n_pad, int D,
                                                          const __grid_constant__ TcPlan plan, int side, int rows_per_tile,
                                                          float* __restrict__ out, unsigned* __restrict__ max_norm_bits) {
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= n_pad) return;
    const bool real = row < n;
    float worst = 0.f;
    for (int a = 0; a < plan.n_acc; a++) {
        const int d = plan.dim[a], k0 = plan.kstep0[a] * 8, Ka = plan.ksteps[a] * 8;
        double nrm = 0.0;
        for (int j = 0; j < d; j++) {
            const double x = real ? X[row * D + plan.start[a] + j] : 0.0;
            nrm += x * x;
            const float hi = to_tf32((float)x);
            const float lo = to_tf32((float)(x - (double)hi));
            float e0, e1, e2;
            if (side == 0) { e0 = hi; e1 = hi; e2 = lo; }
            else { e0 = -2.f * hi; e1 = -2.f * lo; e2 = -2.f * hi; }
            out[tc_elem_offset(rows_per_tile, plan.KS, row, k0 + j)] = e0;
            out[tc_elem_offset(rows_per_tile, plan.KS, row, k0 + d + j)] = e1;
            out[tc_elem_offset(rows_per_tile, plan.KS, row, k0 + 2 * d + j)] = e2;
        }
        float nh = to_tf32((float)nrm);
        float nl = to_tf32((float)(nrm - (double)nh));
        if (!real) { nh = side == 1 ? TC_BIG : 0.f; nl = 0.f; }  // padded corpus rows are infinitely far away
        worst = fmaxf(worst, (float)nrm);
        const float tail[4] = {side == 0 ? nh : 1.f, side == 0 ? nl : 1.f, side == 0 ? 1.f : nh, side == 0 ? 1.f : nl};
        for (int j = 0; j < 4; j++) out[tc_elem_offset(rows_per_tile, plan.KS, row, k0 + 3 * d + j)] = tail[j];
        for (int j = 3 * d + 4; j < Ka; j++) out[tc_elem_offset(rows_per_tile, plan.KS, row, k0 + j)] = 0.f;
    }
    if (real) atomicMax(max_norm_bits, __float_as_uint(worst));  // non-negative floats order like unsigned ints
}

// ---------------------------------------------------------------------------------------------
// main kernel: grid (query tiles, corpus splits)
// ---------------------------------------------------------------------------------------------
struct TcParams {
    const float* A;      // prepared queries  [q tiles][KS][128 x 8]
    const float* B;      // prepared corpus   [c tiles][KS][tn x 8]
    int64_t Q, N;
    int64_t tiles_per_split;  // corpus tiles per split
    int64_t n_ctiles;
    int kc;                   // candidates kept per row and split
    float* part_key;          // [split][half][Q][kc]
    int* part_idx;
    TcPlan plan;
};

__global__ void __launch_bounds__(TC_THREADS, 1) knn_tc_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const TcPlan& plan = p.plan;
    const int KS = plan.KS, tn = plan.tn, n_acc = plan.n_acc, kc = p.kc;
    const uint32_t a_bytes = (uint32_t)KS * TC_TM * 32;
    const uint32_t b_bytes = (uint32_t)KS * tn * 32;
    unsigned char* sA = smem_raw;
    unsigned char* sB = sA + a_bytes;
    float* hk = reinterpret_cast<float*>(sB + (size_t)TC_STAGES * b_bytes);   // [kc][2 * 128] heaps (two column halves per row)
    int* hi = reinterpret_cast<int*>(hk + (size_t)kc * 2 * TC_TM);             // [kc][2 * 128]
    float* stg_k = reinterpret_cast<float*>(hi + (size_t)kc * 2 * TC_TM);      // [TC_STG][2 * 128] staged candidates
    int* stg_i = reinterpret_cast<int*>(stg_k + (size_t)TC_STG * 2 * TC_TM);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg_i + (size_t)TC_STG * 2 * TC_TM);
    uint64_t* full_b = bars;                    // [STAGES] corpus tile landed
    uint64_t* empty_b = bars + TC_STAGES;       // [STAGES] corpus tile consumed by the MMAs
    uint64_t* a_full = bars + 2 * TC_STAGES;    // query tile landed
    uint64_t* tm_full = a_full + 1;             // [2] accumulators ready
    uint64_t* tm_empty = tm_full + 2;           // [2] accumulators drained by the epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tm_empty + 2);
    // issue table of the MMA thread, one entry per K step: A descriptor, B descriptor of stage 0, accumulator column
    // offset | accumulate flag << 31.  Building descriptors inside the issue loop (address conversion, 64-bit shifts,
    // plan look-ups) cost the single issuing thread more cycles per instruction than the tensor pipe needs to run it.
    uint64_t* iss_a = reinterpret_cast<uint64_t*>(tmem_slot + 2);   // [KS]
    uint64_t* iss_b = iss_a + TC_MAX_KS;                              // [KS]
    uint32_t* iss_d = reinterpret_cast<uint32_t*>(iss_b + TC_MAX_KS); // [KS]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int buf_cols = n_acc * tn;            // TMEM columns per accumulator buffer
    uint32_t alloc_cols = 32;
    while ((int)alloc_cols < 2 * buf_cols) alloc_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        mbar_init(a_full, 1);
        for (int b = 0; b < 2; b++) { mbar_init(&tm_full[b], 1); mbar_init(&tm_empty[b], 8); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, alloc_cols);
    if (warp == 1) {
        for (int a = 0; a < n_acc; a++)
            for (int ks = lane; ks < plan.ksteps[a]; ks += 32) {
                const int kg = plan.kstep0[a] + ks;
                iss_a[kg] = umma_desc(sA + (size_t)kg * TC_TM * 32);
                iss_b[kg] = umma_desc(sB + (size_t)kg * tn * 32);
                iss_d[kg] = (uint32_t)(a * tn) | (ks > 0 ? 0x80000000u : 0u);
            }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t qtile = blockIdx.x;
    const int64_t t0 = blockIdx.y * p.tiles_per_split;
    const int64_t t1 = min(p.n_ctiles, t0 + p.tiles_per_split);
    const int64_t n_tiles = max((int64_t)0, t1 - t0);

    if (warp == 0) {
        // ===== producer: one bulk copy per operand tile =====
        if (lane == 0) {
            mbar_arrive_expect_tx(a_full, a_bytes);
            bulk_g2s(sA, p.A + (size_t)qtile * KS * TC_TM * 8, a_bytes, a_full);
            int s = 0;
            uint32_t ph = 0;
            for (int64_t t = 0; t < n_tiles; ++t) {
                mbar_wait(&empty_b[s], ph ^ 1);
                mbar_arrive_expect_tx(&full_b[s], b_bytes);
                bulk_g2s(sB + (size_t)s * b_bytes, p.B + (size_t)(t0 + t) * KS * tn * 8, b_bytes, &full_b[s]);
                if (++s == TC_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(tn);
            mbar_wait(a_full, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int64_t t = 0; t < n_tiles; ++t) {
                const int buf = (int)(t & 1);
                const uint32_t use = (uint32_t)(t >> 1);           // how often this buffer was used before
                mbar_wait(&tm_empty[buf], (use & 1) ^ 1);           // first use passes immediately
                mbar_wait(&full_b[s], ph);
                tc_fence_after();
                const uint64_t stage_off = (uint64_t)(((uint32_t)s * b_bytes) >> 4);   // added to the 14-bit address field
                const uint32_t d_base = tmem_base + (uint32_t)(buf * buf_cols);
#pragma unroll 4
                for (int kg = 0; kg < KS; kg++) {
                    const uint32_t d = iss_d[kg];
                    umma_tf32(d_base + (d & 0x7fffffffu), iss_a[kg], iss_b[kg] + stage_off, idesc, d >> 31);
                }
                umma_commit(&empty_b[s]);     // the stage may be refilled once these MMAs have read it
                umma_commit(&tm_full[buf]);   // accumulators complete
                if (++s == TC_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; thread = (query row, half); 16-column chunks alternate between halves =====
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        const int r_in_tile = quarter * 32 + lane;
        const int hslot = half * TC_TM + r_in_tile;            // this thread's heap / staging column
        const int64_t row = qtile * TC_TM + r_in_tile;
        ThreadHeap<float> heap(hk + hslot, hi + hslot, 2 * TC_TM, kc, 0);
        float* sk = stg_k + hslot;                             // staging [TC_STG][2 * TC_TM]
        int* si = stg_i + hslot;
        int staged = 0;
        float thr = 3.0e38f;
        const int chunks_per_tile = tn >> 4;   // 16-column chunks: the loads of up to four robots' accumulators fly together
        // every lane pushes its staged candidates into its own heap at the same time: the sift loops of the 32
        // lanes run side by side instead of one lane at a time
        auto flush = [&]() {
            for (int e = 0; e < staged; e++) {
                const float key = sk[e * 2 * TC_TM];
                const int idx = si[e * 2 * TC_TM];
                if (heap.accepts(key, idx)) heap.push(key, idx);
            }
            staged = 0;
            if (heap.full()) thr = heap.top_key();
        };
        for (int64_t t = 0; t < n_tiles; ++t) {
            const int buf = (int)(t & 1);
            const uint32_t use = (uint32_t)(t >> 1);
            mbar_wait(&tm_full[buf], use & 1);
            tc_fence_after();
            const int64_t col0 = (t0 + t) * tn;
            // the two halves take alternate chunks (chunks_per_tile is even, so the parity of a chunk is that of ci)
            for (int ci = half; ci < chunks_per_tile; ci += 2) {
                const int c = ci << 4;
                float v[16], b1[16], b2[16], b3[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * buf_cols + c);
                // issue every accumulator's load before the one wait (n_acc is uniform across the CTA)
                tmem_ld16(taddr, v);
                if (n_acc > 1) tmem_ld16(taddr + (uint32_t)tn, b1);
                if (n_acc > 2) tmem_ld16(taddr + (uint32_t)(2 * tn), b2);
                if (n_acc > 3) tmem_ld16(taddr + (uint32_t)(3 * tn), b3);
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
                // per-column mask only when some lane of the warp needs it
                float lo8[8], lo4[4];
#pragma unroll
                for (int j = 0; j < 8; j++) lo8[j] = fminf(v[j], v[8 + j]);
#pragma unroll
                for (int j = 0; j < 4; j++) lo4[j] = fminf(lo8[j], lo8[4 + j]);
                const float lo = fminf(fminf(lo4[0], lo4[1]), fminf(lo4[2], lo4[3]));
                if (!__any_sync(0xffffffffu, lo < thr)) continue;
                uint32_t mask = 0u;
#pragma unroll
                for (int j = 0; j < 16; j++) mask |= (v[j] < thr ? 1u : 0u) << j;
                while (mask) {  // a handful of candidates per chunk and warp
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1u;
                    // v[j] with a run-time j: four levels of selects keep v in registers
                    float s8[8], s4[4], s2[2];
#pragma unroll
                    for (int i = 0; i < 8; i++) s8[i] = (j & 8) ? v[8 + i] : v[i];
#pragma unroll
                    for (int i = 0; i < 4; i++) s4[i] = (j & 4) ? s8[4 + i] : s8[i];
#pragma unroll
                    for (int i = 0; i < 2; i++) s2[i] = (j & 2) ? s4[2 + i] : s4[i];
                    const float val = (j & 1) ? s2[1] : s2[0];
                    const int64_t col = col0 + c + j;
                    if (col < p.N) {
                        if (staged == TC_STG) flush();  // only while the heap is still filling up
                        sk[staged * 2 * TC_TM] = val;
                        si[staged * 2 * TC_TM] = (int)col;
                        staged++;
                    }
                }
                if (__any_sync(0xffffffffu, staged >= TC_STG - 2)) flush();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tm_empty[buf]);
        }
        flush();
        if (row < p.Q) {
            float* ok = p.part_key + (((size_t)blockIdx.y * 2 + half) * p.Q + row) * kc;
            int* oi = p.part_idx + (((size_t)blockIdx.y * 2 + half) * p.Q + row) * kc;
            for (int e = 0; e < kc; e++) {
                ok[e] = e < heap.n ? heap.key[e * 2 * TC_TM] : 3.0e38f;
                oi[e] = e < heap.n ? heap.idx[e * 2 * TC_TM] : -1;
            }
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
                                                         int splits, const float* __restrict__ part_key, const int* __restrict__ part_idx,
                                                         const unsigned* __restrict__ max_norm_bits, int32_t* __restrict__ out_idx,
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
    for (int s = 0; s < splits; s++) {
        const float* pk = part_key + ((size_t)s * Q + row) * kc;
        const int* pi = part_idx + ((size_t)s * Q + row) * kc;
        float worst = -3.0e38f;
        bool full = true;
        for (int e = 0; e < kc; e++) {
            const int idx = pi[e];
            if (idx < 0) { full = false; continue; }
            worst = fmaxf(worst, pk[e]);
            const double d = metric_dist<DMAX>(q, corpus + (size_t)idx * D, D, sl, metric);
            if (heap.accepts(d, idx)) heap.push(d, idx);
        }
        if (full) tau = fminf(tau, worst);  // a split whose list is not full kept all of its points
    }
    // coarse values carry an absolute error of a few 2^-22 of the largest squared norms involved
    const float eps = 8e-6f * (2.f * __uint_as_float(*max_norm_bits) + 1.f);
    bool ok = true;
    if (heap.n == k) {
        const double dk = heap.top_key();
        ok = (float)(dk * dk) < tau - eps;
    } else {
        ok = tau > 1.0e38f;  // fewer than k found: fine only if nothing was discarded anywhere
    }
    certified[row] = ok ? 1 : 0;
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
        plan->ksteps[a] = pad8(3 * plan->dim[a] + 4) / 8;
        plan->KS += plan->ksteps[a];
    }
    int tn = 256 / plan->n_acc;
    plan->tn = tn >= 256 ? 256 : tn >= 128 ? 128 : tn >= 64 ? 64 : 32;
    return plan->KS <= TC_MAX_KS;
}

size_t knn_tc_smem_bytes(const TcPlan& plan, int kc) {
    return (size_t)plan.KS * TC_TM * 32 + (size_t)TC_STAGES * plan.KS * plan.tn * 32 + (size_t)(kc + TC_STG) * 2 * TC_TM * 8 + 16 * 8 + (size_t)TC_MAX_KS * 20 + 1024;
}

int knn_tc_splits(int64_t Q, int64_t n_ctiles) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t qt = (Q + TC_TM - 1) / TC_TM;
    int64_t s = (sms + qt - 1) / qt;                 // at least one CTA per SM
    const int64_t max_s = (n_ctiles + 15) / 16;      // at least 16 corpus tiles per split
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 32) s = 32;
    return (int)s;
}

size_t knn_tc_workspace_bytes(int64_t Q, int64_t N, const TcPlan& plan, int kc, int splits) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    return (size_t)qt * plan.KS * TC_TM * 32 + (size_t)ct * plan.KS * plan.tn * 32 + (size_t)splits * 2 * Q * kc * 8 + (size_t)Q + 1024;
}

cudaError_t launch_knn_tc(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                          int k, int kc, const TcPlan& plan, int splits, void* workspace, int32_t* out_idx, double* out_dist,
                          uint8_t* certified, cudaStream_t st) {
    const int64_t qt = (Q + TC_TM - 1) / TC_TM, ct = (N + plan.tn - 1) / plan.tn;
    unsigned char* w = (unsigned char*)workspace;
    unsigned* max_norm = (unsigned*)w;
    w += 256;
    float* A = (float*)w;
    w += (size_t)qt * plan.KS * TC_TM * 32;
    float* B = (float*)w;
    w += (size_t)ct * plan.KS * plan.tn * 32;
    float* part_key = (float*)w;
    w += (size_t)splits * 2 * Q * kc * 4;
    int* part_idx = (int*)w;
    cudaError_t e = cudaMemsetAsync(max_norm, 0, 4, st);
    if (e != cudaSuccess) return e;
    knn_tc_prep_kernel<<<(unsigned)((qt * TC_TM + 127) / 128), 128, 0, st>>>(queries, Q, qt * TC_TM, D, plan, 0, TC_TM, A, max_norm);
    knn_tc_prep_kernel<<<(unsigned)((ct * plan.tn + 127) / 128), 128, 0, st>>>(corpus, N, ct * plan.tn, D, plan, 1, plan.tn, B, max_norm);
    TcParams p;
    p.A = A;
    p.B = B;
    p.Q = Q;
    p.N = N;
    p.n_ctiles = ct;
    p.tiles_per_split = (ct + splits - 1) / splits;
    p.kc = kc;
    p.part_key = part_key;
    p.part_idx = part_idx;
    p.plan = plan;
    const size_t smem = knn_tc_smem_bytes(plan, kc);
    e = cudaFuncSetAttribute(knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    knn_tc_kernel<<<dim3((unsigned)qt, (unsigned)splits), TC_THREADS, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t rsmem = (size_t)k * 128 * 12;
#define MRB_RERANK(DM)                                                                                                            \
    do {                                                                                                                          \
        e = cudaFuncSetAttribute(knn_rerank_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);                 \
        if (e != cudaSuccess) return e;                                                                                           \
        knn_rerank_kernel<DM><<<(unsigned)((Q + 127) / 128), 128, rsmem, st>>>(queries, corpus, Q, D, sl, metric, k, kc, 2 * splits,  \
                                                                               part_key, part_idx, max_norm, out_idx, out_dist,  \
                                                                               certified);                                        \
    } while (0)
    if (D <= 8) MRB_RERANK(8);
    else if (D <= 16) MRB_RERANK(16);
    else if (D <= 24) MRB_RERANK(24);
    else if (D <= 32) MRB_RERANK(32);
    else MRB_RERANK(64);
#undef MRB_RERANK
    return cudaGetLastError();
}

}  // namespace mrb
