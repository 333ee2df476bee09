// FK + primitive narrowphase kernels for rai-style primitive scenes (sm_100a).
//
// Replaces, for a whole batch at once, the reference's per-configuration query
//   rai_env.is_collision_free / is_collision_free_np   P/problems/rai_base_env.py:442-513
//   rai_env.is_collision_free_for_robot                P/problems/rai_base_env.py:515-615
// and its per-edge loop
//   rai_env.is_edge_collision_free                     P/problems/rai_base_env.py:618-676
//   generate_binary_search_indices                     P/problems/planning_env.py:34-51
// (paths relative to /root/reference, P/ = src/multi_robot_multi_goal_planning/).
//
// Mapping: one CTA = 4 warps works on a tile of 32 configurations; lane = configuration.
//   phase 1 (FK):    warp w walks kinematic chains w, w+4, ... with the running link transform
//                    in registers and writes world-space shape data to shared memory, laid out
//                    [word][lane] so every later access is bank-conflict free.
//   phase 2 (pairs): the typed pair lists are split four ways; all lanes of a warp evaluate the
//                    same pair (uniform control flow, scene data broadcast from shared memory).
//   phase 3:         per-configuration penetration = static + 4 partial sums, added in a fixed
//                    order (deterministic); flag = !(penetration > tol).
// The scene blob (a few KB) and each tile of configurations arrive in shared memory through
// 1-D bulk async copies (TMA, cp.async.bulk + mbarrier); configuration tiles are double buffered.
#include <cuda_runtime.h>
#include <stdint.h>

#include "async_copy.cuh"
#include "binary_order.cuh"
#include "kernels.h"
#include "narrowphase.cuh"
#include "scene_blob.h"

namespace mrb {

constexpr int TILE = 32;
constexpr int WARPS = 4;
constexpr int THREADS = TILE * WARPS;
constexpr unsigned FULL = 0xffffffffu;

struct Smem {
    uint32_t* blob;  // staged scene blob
    float* q[2];     // configuration tiles [TILE][D]
    float* W;        // world shape data [world_words][TILE]
    float* pen;      // [WARPS][TILE]
    unsigned* relb;  // [WARPS] ballots of "relevant pair penetrates"
    uint8_t* sflag;  // [n_shapes] bit0 = relevant, bit1 = other robot (A6 rule)
    uint64_t* bar;   // [3] mbarriers: blob, q0, q1
    int* misc;       // [40] small broadcast scratch (edge kernel)
    double* ed;      // [2*D] edge start and step (fp64)
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t smem_layout(int blob_words, int D, int world_words, int n_shapes, size_t* off) {
    size_t o = 0;
    off[0] = o; o = align16(o + size_t(blob_words) * 4);
    off[1] = o; o = align16(o + size_t(TILE) * D * 4);
    off[2] = o; o = align16(o + size_t(TILE) * D * 4);
    off[3] = o; o = align16(o + size_t(world_words) * TILE * 4);
    off[4] = o; o = align16(o + size_t(WARPS) * TILE * 4);
    off[5] = o; o = align16(o + WARPS * 4);
    off[6] = o; o = align16(o + size_t(n_shapes));
    off[7] = o; o = align16(o + 3 * 8);
    off[8] = o; o = align16(o + 40 * 4);
    off[9] = o; o = align16(o + size_t(2) * D * 8);
    return o;
}

__device__ __forceinline__ Smem carve(unsigned char* base, int blob_words, int D, int world_words, int n_shapes) {
    size_t off[10];
    smem_layout(blob_words, D, world_words, n_shapes, off);
    Smem s;
    s.blob = (uint32_t*)(base + off[0]);
    s.q[0] = (float*)(base + off[1]);
    s.q[1] = (float*)(base + off[2]);
    s.W = (float*)(base + off[3]);
    s.pen = (float*)(base + off[4]);
    s.relb = (unsigned*)(base + off[5]);
    s.sflag = (uint8_t*)(base + off[6]);
    s.bar = (uint64_t*)(base + off[7]);
    s.misc = (int*)(base + off[8]);
    s.ed = (double*)(base + off[9]);
    return s;
}

size_t scene_smem_bytes(int blob_words, int D, int world_words, int n_shapes) {
    size_t off[10];
    return smem_layout(blob_words, D, world_words, n_shapes, off);
}

// ------------------------------------------------------------------------------------------
// phase 1: forward kinematics, link transform in registers, one configuration per lane
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rot_cols(float* R, int i, int j, float c, float s) {
    // columns i, j of R <- (c*ci + s*cj, -s*ci + c*cj)
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float a = R[r * 3 + i], b = R[r * 3 + j];
        R[r * 3 + i] = fmaf(c, a, s * b);
        R[r * 3 + j] = fmaf(c, b, -s * a);
    }
}

__device__ __forceinline__ void xform_point(const float* R, const float* t, const float* l, float* w, int stride) {
#pragma unroll
    for (int r = 0; r < 3; r++) w[r * stride] = fmaf(R[r * 3], l[0], fmaf(R[r * 3 + 1], l[1], fmaf(R[r * 3 + 2], l[2], t[r])));
}

__device__ __forceinline__ void fk_phase(const uint32_t* bi, const float* q, float* W, int warp, int lane) {
    const float* bf = reinterpret_cast<const float*>(bi);
    const int n_chains = bi[MRB_H_NCHAINS];
    const int offF = bi[MRB_H_OFF_FRAMES], offS = bi[MRB_H_OFF_SHAPES], offC = bi[MRB_H_OFF_CHAINS];
    for (int c = warp; c < n_chains; c += WARPS) {
        const int f0 = bi[offC + 2 * c], f1 = bi[offC + 2 * c + 1];
        float R[9], t[3];
        for (int f = f0; f < f1; ++f) {
            const int row = offF + f * MRB_FRAME_WORDS;
            const float* A = bf + row + 4;
            if (f == f0) {
#pragma unroll
                for (int k = 0; k < 9; k++) R[k] = A[k];
#pragma unroll
                for (int k = 0; k < 3; k++) t[k] = A[9 + k];
            } else {
                float nR[9], nt[3];
#pragma unroll
                for (int r = 0; r < 3; r++) {
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        nR[r * 3 + k] = fmaf(R[r * 3], A[k], fmaf(R[r * 3 + 1], A[3 + k], R[r * 3 + 2] * A[6 + k]));
                    nt[r] = fmaf(R[r * 3], A[9], fmaf(R[r * 3 + 1], A[10], fmaf(R[r * 3 + 2], A[11], t[r])));
                }
#pragma unroll
                for (int k = 0; k < 9; k++) R[k] = nR[k];
#pragma unroll
                for (int k = 0; k < 3; k++) t[k] = nt[k];
            }
            const int code = bi[row + 1], qi = bi[row + 2];
            float s, co;
            switch (code) {  // warp-uniform
                case MRB_J_HINGE_X: sincosf(q[qi], &s, &co); rot_cols(R, 1, 2, co, s); break;
                case MRB_J_HINGE_Y: sincosf(q[qi], &s, &co); rot_cols(R, 2, 0, co, s); break;
                case MRB_J_HINGE_Z: sincosf(q[qi], &s, &co); rot_cols(R, 0, 1, co, s); break;
                case MRB_J_TRANS_XY_PHI: {
                    float x = q[qi], y = q[qi + 1];
#pragma unroll
                    for (int r = 0; r < 3; r++) t[r] = fmaf(R[r * 3], x, fmaf(R[r * 3 + 1], y, t[r]));
                    sincosf(q[qi + 2], &s, &co);
                    rot_cols(R, 0, 1, co, s);
                } break;
                case MRB_J_TRANS_X:
                case MRB_J_TRANS_Y:
                case MRB_J_TRANS_Z: {
                    float x = q[qi];
                    const int col = code - MRB_J_TRANS_X;
#pragma unroll
                    for (int r = 0; r < 3; r++) t[r] = fmaf(R[r * 3 + col], x, t[r]);
                } break;
                default: break;
            }
            // world data of the shapes riding on this frame
            const uint32_t sh = bi[row + 3];
            const int s0 = sh & 0xffff, sn = sh >> 16;
            for (int si = s0; si < s0 + sn; ++si) {
                const int srow = offS + si * MRB_SHAPE_WORDS;
                const int core = bi[srow];
                const float* L = bf + srow + 4;
                float* w = W + bi[srow + 2] * TILE + lane;
                xform_point(R, t, L, w, TILE);
                if (core == MRB_CORE_SEG) {
                    xform_point(R, t, L + 3, w + 3 * TILE, TILE);
                } else if (core == MRB_CORE_BOX) {
#pragma unroll
                    for (int r = 0; r < 3; r++)
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            w[(3 + r * 3 + k) * TILE] = fmaf(R[r * 3], L[3 + k], fmaf(R[r * 3 + 1], L[6 + k], R[r * 3 + 2] * L[9 + k]));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// phase 2: pair loops
// ------------------------------------------------------------------------------------------
struct PairCtx {
    const uint32_t* bi;
    const float* bf;
    const float* W;
    const uint8_t* sflag;
    int offS, nmov, lane;
    bool rule;

    template <int NW>
    __device__ __forceinline__ void load(int s, float* out) const {
        const int row = offS + s * MRB_SHAPE_WORDS;
        if (s < nmov) {  // warp-uniform
            const float* p = W + bi[row + 2] * TILE + lane;
#pragma unroll
            for (int k = 0; k < NW; k++) out[k] = p[k * TILE];
        } else {
            const float* p = bf + row + 4;
#pragma unroll
            for (int k = 0; k < NW; k++) out[k] = p[k];
        }
    }
    __device__ __forceinline__ float radius(int s) const { return bf[offS + s * MRB_SHAPE_WORDS + 3]; }
    __device__ __forceinline__ const float* rowdata(int s) const { return bf + offS + s * MRB_SHAPE_WORDS + 4; }
    __device__ __forceinline__ bool relevant(int a, int b) const {
        const unsigned f = sflag[a] | sflag[b];
        return (f & 1u) && !(f & 2u);
    }
};

#define MRB_PAIR_LOOP(TYPE, BODY)                                                            \
    {                                                                                        \
        const int n_ = bi[MRB_H_N_PAIRS + (TYPE)], off_ = bi[MRB_H_OFF_PAIRS + (TYPE)];      \
        const int lo_ = (n_ * warp) / WARPS, hi_ = (n_ * (warp + 1)) / WARPS;                \
        int prev_a = -1;                                                                     \
        (void)prev_a;                                                                        \
        for (int i_ = lo_; i_ < hi_; ++i_) {                                                 \
            const uint32_t pk_ = bi[off_ + i_];                                              \
            const int a = pk_ & 0xffff, b = pk_ >> 16;                                       \
            float d;                                                                         \
            BODY;                                                                            \
            if (d < 0.f) {                                                                   \
                pen -= d;                                                                    \
                if (ctx.rule && ctx.relevant(a, b)) relpen = true;                           \
            }                                                                                \
            if (early && ((i_ & 7) == 7) && __all_sync(FULL, pen > tol)) goto pairs_done;    \
        }                                                                                    \
    }

// All THREADS threads call this.  On return warp 0 holds, per lane, the configuration's total
// penetration (return value) and whether a relevant pair penetrates (*relpen_out).
__device__ __forceinline__ float process_tile(const Smem& sm, const float* q_tile, int D, float tol, bool early, bool rule,
                                              bool* relpen_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t* bi = sm.blob;
    fk_phase(bi, q_tile + lane * D, sm.W, warp, lane);
    __syncthreads();

    PairCtx ctx{bi, reinterpret_cast<const float*>(bi), sm.W, sm.sflag, (int)bi[MRB_H_OFF_SHAPES], (int)bi[MRB_H_NMOV], lane, rule};
    float pen = 0.f;
    bool relpen = false;
    float A[12];
    float ra = 0.f;

    MRB_PAIR_LOOP(MRB_PT_SEG_SEG, {
        if (a != prev_a) { ctx.load<6>(a, A); ra = ctx.radius(a); prev_a = a; }
        float Bv[6];
        ctx.load<6>(b, Bv);
        d = d_seg_seg(A, Bv, ra + ctx.radius(b));
    })
    MRB_PAIR_LOOP(MRB_PT_SEG_BOX, {
        if (a != prev_a) { ctx.load<6>(a, A); ra = ctx.radius(a); prev_a = a; }
        float Bv[12];
        ctx.load<12>(b, Bv);
        d = d_seg_box(A, Bv, Bv + 3, ctx.rowdata(b) + 12, ra + ctx.radius(b));
    })
    MRB_PAIR_LOOP(MRB_PT_POINT_POINT, {
        float Bv[3];
        ctx.load<3>(a, A);
        ctx.load<3>(b, Bv);
        d = d_point_point(A, Bv, ctx.radius(a) + ctx.radius(b));
    })
    MRB_PAIR_LOOP(MRB_PT_POINT_SEG, {
        float Bv[6];
        ctx.load<3>(a, A);
        ctx.load<6>(b, Bv);
        d = d_point_seg(A, Bv, ctx.radius(a) + ctx.radius(b));
    })
    MRB_PAIR_LOOP(MRB_PT_POINT_BOX, {
        float Bv[12];
        ctx.load<3>(a, A);
        ctx.load<12>(b, Bv);
        d = d_point_box(A, Bv, Bv + 3, ctx.rowdata(b) + 12, ctx.radius(a) + ctx.radius(b));
    })
    MRB_PAIR_LOOP(MRB_PT_BOX_BOX, {
        float Bv[12];
        ctx.load<12>(a, A);
        ctx.load<12>(b, Bv);
        d = d_box_box(A, A + 3, ctx.rowdata(a) + 12, Bv, Bv + 3, ctx.rowdata(b) + 12, ctx.radius(a) + ctx.radius(b));
    })
    MRB_PAIR_LOOP(MRB_PT_CYLZ_CYLZ, {
        float Bv[3];
        ctx.load<3>(a, A);
        ctx.load<3>(b, Bv);
        d = d_cylz_cylz(A, ctx.rowdata(a)[3], ctx.rowdata(a)[4], Bv, ctx.rowdata(b)[3], ctx.rowdata(b)[4]);
    })
    MRB_PAIR_LOOP(MRB_PT_BOX_CYLZ, {
        float Bv[3];
        ctx.load<12>(a, A);
        ctx.load<3>(b, Bv);
        d = d_box_cylz(A, A + 3, ctx.rowdata(a) + 12, Bv, ctx.rowdata(b)[3], ctx.rowdata(b)[4]);
    })
pairs_done:
    sm.pen[warp * TILE + lane] = pen;
    const unsigned rb = __ballot_sync(FULL, relpen);
    if (lane == 0) sm.relb[warp] = rb;
    __syncthreads();
    float total = 0.f;
    if (warp == 0) {
        total = reinterpret_cast<const float*>(bi)[MRB_H_STATIC_PEN];
#pragma unroll
        for (int w = 0; w < WARPS; w++) total += sm.pen[w * TILE + lane];
        unsigned r = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) r |= sm.relb[w];
        *relpen_out = (r >> lane) & 1u;
    }
    return total;
}

// common prologue: barriers, blob staging through TMA, A6 shape flags
__device__ __forceinline__ void stage_scene(const Smem& sm, const uint32_t* blob, int blob_words, int n_shapes,
                                            const RobotRule& rr) {
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_init(&sm.bar[2], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&sm.bar[0], blob_words * 4);
        bulk_g2s(sm.blob, blob, blob_words * 4, &sm.bar[0]);
    }
    for (int s = threadIdx.x; s < n_shapes; s += THREADS) {
        unsigned f = 0;
        if (rr.enabled) f = ((rr.rel[s >> 6] >> (s & 63)) & 1u) | (((rr.oth[s >> 6] >> (s & 63)) & 1u) << 1);
        sm.sflag[s] = (uint8_t)f;
    }
    mbar_wait(&sm.bar[0], 0);
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// configuration batch kernel (A5 / A6 batch variant)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) check_configs_kernel(ConfigParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Smem sm = carve(smem_raw, p.blob_words, p.D, p.world_words, p.n_shapes);
    stage_scene(sm, p.blob, p.blob_words, p.n_shapes, p.rule);

    const int D = p.D;
    const float tol = p.tol < 0.f ? reinterpret_cast<const float*>(sm.blob)[MRB_H_TOL] : p.tol;
    const bool early = !p.full_eval && !p.rule.enabled;
    const int64_t n_tiles = (p.B + TILE - 1) / TILE;
    const uint32_t tile_bytes = TILE * D * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto fetch = [&](int64_t tile, int buf) {
        // full tiles: one bulk copy; the ragged tail tile (and unaligned inputs): plain loads
        const int64_t first = tile * TILE;
        const int nvalid = (int)min((int64_t)TILE, p.B - first);
        if (nvalid == TILE && p.bulk_ok) {
            if (threadIdx.x == 0) {
                mbar_arrive_expect_tx(&sm.bar[1 + buf], tile_bytes);
                bulk_g2s(sm.q[buf], p.q + first * D, tile_bytes, &sm.bar[1 + buf]);
            }
        } else {
            for (int i = threadIdx.x; i < TILE * D; i += THREADS) {
                const int c = i / D;
                sm.q[buf][i] = p.q[(first + (c < nvalid ? c : 0)) * D + (i - c * D)];
            }
        }
    };

    uint32_t phase[2] = {0, 0};
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) fetch(tile, 0);
    for (int it = 0; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t next = tile + gridDim.x;
        if (next < n_tiles) fetch(next, buf ^ 1);  // buffer buf^1 was released by the barrier closing iteration it-1
        const int64_t first = tile * TILE;
        const int nvalid = (int)min((int64_t)TILE, p.B - first);
        if (nvalid == TILE && p.bulk_ok) {
            mbar_wait(&sm.bar[1 + buf], phase[buf]);
            phase[buf] ^= 1;
        } else {
            __syncthreads();
        }
        bool relpen = false;
        const float total = process_tile(sm, sm.q[buf], D, tol, early, p.rule.enabled, &relpen);
        if (warp == 0 && lane < nvalid) {
            const bool coll = p.rule.enabled ? (total > tol && relpen) : (total > tol);
            p.flags[first + lane] = coll ? 0 : 1;
            if (p.pen_out) p.pen_out[first + lane] = total;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// edge batch kernel (A8 batch variant): one CTA per edge at a time, 32 interpolation points
// per step in the reference's binary order, early exit on the first colliding step
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) check_edges_kernel(EdgeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Smem sm = carve(smem_raw, p.blob_words, p.D, p.world_words, p.n_shapes);
    RobotRule none{};
    stage_scene(sm, p.blob, p.blob_words, p.n_shapes, none);

    const int D = p.D;
    const float tol = p.tol < 0.f ? reinterpret_cast<const float*>(sm.blob)[MRB_H_TOL] : p.tol;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* s_edge = sm.misc;       // [0] edge id, [1] N, [2] first colliding position
    int* s_idx = sm.misc + 8;    // [32] interpolation index per lane
    double* e_start = sm.ed;     // q1 (fp64)
    double* e_step = sm.ed + D;  // (q2 - q1) / (N - 1)

    for (;;) {
        if (threadIdx.x == 0) s_edge[0] = atomicAdd(p.counter, 1);
        __syncthreads();
        const int64_t e = s_edge[0];
        if (e >= p.E) break;
        // endpoints in fp64, N exactly as the reference: max(2, int(|dq|_inf / resolution) + 1)
        if (warp == 0) {
            double m = 0.0;
            for (int k = lane; k < D; k += 32) {
                const double a = (double)p.q1[e * D + k], b = (double)p.q2[e * D + k];
                e_start[k] = a;
                e_step[k] = __dsub_rn(b, a);
                m = fmax(m, fabs(__dsub_rn(a, b)));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
            int N = p.N ? p.N[e] : max(2, (int)__ddiv_rn(m, p.resolution) + 1);
            if (lane == 0) { s_edge[1] = N; s_edge[2] = -1; }
            const double inv = (double)(N - 1);
            for (int k = lane; k < D; k += 32) e_step[k] = __ddiv_rn(e_step[k], inv);
        }
        __syncthreads();
        const int N = s_edge[1];
        const int nmax = (p.n_max < 0 || p.n_max > N) ? N : p.n_max;
        for (int base = p.n_start; base < nmax; base += TILE) {
            if (warp == 0) {
                const int pos = base + lane;
                int i = -1;
                if (pos < nmax) {
                    i = binary_order_index(N, pos);
                    if (!p.include_endpoints && (i == 0 || i == N - 1)) i = -1;
                }
                s_idx[lane] = i;
            }
            __syncthreads();
            // q = q1 + step * i in fp64 (reference order of operations), rounded once to fp32
            for (int t = threadIdx.x; t < TILE * D; t += THREADS) {
                const int c = t / D, k = t - c * D;
                const int i = s_idx[c];
                sm.q[0][t] = (float)__dadd_rn(e_start[k], __dmul_rn(e_step[k], (double)(i < 0 ? 0 : i)));
            }
            __syncthreads();
            bool relpen;
            const float total = process_tile(sm, sm.q[0], D, tol, false, false, &relpen);
            if (warp == 0) {
                const unsigned hit = __ballot_sync(FULL, s_idx[lane] >= 0 && total > tol);
                if (lane == 0 && hit) s_edge[2] = base + __ffs(hit) - 1;
            }
            __syncthreads();
            if (s_edge[2] >= 0) break;
        }
        if (threadIdx.x == 0) {
            p.flags[e] = s_edge[2] < 0 ? 1 : 0;
            if (p.first_pos) p.first_pos[e] = s_edge[2];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// mode setup: penetration of static-static pairs (constant for the mode), written into the blob
// ------------------------------------------------------------------------------------------
__global__ void static_penetration_kernel(uint32_t* blob) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const uint32_t* bi = blob;
    const float* bf = reinterpret_cast<const float*>(blob);
    const int n = bi[MRB_H_N_STATIC_PAIRS], off = bi[MRB_H_OFF_STATIC_PAIRS], offS = bi[MRB_H_OFF_SHAPES];
    float pen = 0.f;
    for (int i = 0; i < n; i++) {
        const int t = bi[off + 3 * i], a = bi[off + 3 * i + 1], b = bi[off + 3 * i + 2];
        const float* A = bf + offS + a * MRB_SHAPE_WORDS + 4;
        const float* Bv = bf + offS + b * MRB_SHAPE_WORDS + 4;
        const float rs = bf[offS + a * MRB_SHAPE_WORDS + 3] + bf[offS + b * MRB_SHAPE_WORDS + 3];
        float d;
        switch (t) {
            case MRB_PT_POINT_POINT: d = d_point_point(A, Bv, rs); break;
            case MRB_PT_POINT_SEG: d = d_point_seg(A, Bv, rs); break;
            case MRB_PT_SEG_SEG: d = d_seg_seg(A, Bv, rs); break;
            case MRB_PT_POINT_BOX: d = d_point_box(A, Bv, Bv + 3, Bv + 12, rs); break;
            case MRB_PT_SEG_BOX: d = d_seg_box(A, Bv, Bv + 3, Bv + 12, rs); break;
            case MRB_PT_CYLZ_CYLZ: d = d_cylz_cylz(A, A[3], A[4], Bv, Bv[3], Bv[4]); break;
            case MRB_PT_BOX_CYLZ: d = d_box_cylz(A, A + 3, A + 12, Bv, Bv[3], Bv[4]); break;
            default: d = d_box_box(A, A + 3, A + 12, Bv, Bv + 3, Bv + 12, rs); break;
        }
        if (d < 0.f) pen -= d;
    }
    reinterpret_cast<float*>(blob)[MRB_H_STATIC_PEN] = pen;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <typename K>
static int grid_for(K kernel, size_t smem, int* blocks_per_sm) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, smem);
    if (occ < 1) occ = 1;
    *blocks_per_sm = occ;
    return occ * num_sms();
}

cudaError_t launch_static_penetration(uint32_t* blob, cudaStream_t st) {
    static_penetration_kernel<<<1, 32, 0, st>>>(blob);
    return cudaGetLastError();
}

cudaError_t launch_check_configs(const ConfigParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    const size_t smem = scene_smem_bytes(p.blob_words, p.D, p.world_words, p.n_shapes);
    int occ;
    int grid = grid_for(check_configs_kernel, smem, &occ);
    const int64_t n_tiles = (p.B + TILE - 1) / TILE;
    if (grid > n_tiles) grid = (int)n_tiles;
    check_configs_kernel<<<grid, THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_check_edges(const EdgeParams& p, cudaStream_t st) {
    if (p.E <= 0) return cudaSuccess;
    const size_t smem = scene_smem_bytes(p.blob_words, p.D, p.world_words, p.n_shapes);
    int occ;
    int grid = grid_for(check_edges_kernel, smem, &occ);
    if (grid > p.E) grid = (int)p.E;
    cudaError_t err = cudaMemsetAsync(p.counter, 0, sizeof(int), st);
    if (err != cudaSuccess) return err;
    check_edges_kernel<<<grid, THREADS, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace mrb
