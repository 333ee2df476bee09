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
constexpr int MAX_WARPS = 4;  // warps cooperating on one tile of 32 configurations: 2 or 4 (template parameter)
constexpr int LAT_WARPS = 8;  // ... and on the ONE tile of a single query (the reference planners' call pattern): the broadphase
                              // records and the kinematic chains are split eight ways -- latency, not throughput
constexpr unsigned FULL = 0xffffffffu;

// The one dynamic shared-memory window of every kernel in this file.  It is declared at file scope so
// that out-of-line device functions can rebuild their pointers from it (see TileArgs): a pointer
// passed through a call boundary is generic (LD.E / ATOM.E with 64-bit address arithmetic), a pointer
// derived from this symbol is known to be shared (LDS / STS / ATOMS, 32-bit addresses).
extern __shared__ __align__(128) unsigned char smem_raw[];

// (shared-memory pointers only: a global pointer stored next to them makes nvcc address the whole struct's
// pointers through global stores -- STG.E to a shared-window address faults)
// Narrowphase batches drained back to back once DRAIN_BATCH x 32 survivors are queued (experiment switch MRB_DRAIN_BATCH):
// a second pass through the same drain routine finds its instructions in the instruction caches.
#ifndef MRB_DRAIN_BATCH
#define MRB_DRAIN_BATCH 1
#endif
constexpr int DRAIN_BATCH = MRB_DRAIN_BATCH;
constexpr int QCAP_WORDS = 32 * (DRAIN_BATCH + 1);   // queue entries per warp

struct Smem {
    uint32_t* blob;  // staged prefix of the scene blob
    float* q[2];     // configuration tiles [TILE][D]
    float* W;        // world shape data [world_words][TILE]
    unsigned* pen_fx;  // [TILE] fixed-point penetration per configuration
    unsigned* relf;    // [TILE] "a relevant pair penetrates" (A6 rule)
    uint32_t* queue;   // [WARPS][QCAP] surviving (configuration, pair) items
    uint8_t* sflag;  // [n_shapes] bit0 = relevant, bit1 = other robot (A6 rule)
    uint64_t* bar;   // [3] mbarriers: blob, q0, q1
    int* misc;       // edge kernel: [EDGE_MISC] bookkeeping (active edge slots, lane assignment);
                     // configuration kernel: [CFG_MISC] survivor pool bookkeeping
    double* ed;      // edge kernel: [EDGE_SLOTS][2*D] start and step of the active edges (fp64);
                     // configuration kernel: [TILE][D] floats, the pooled survivors' configurations
};

constexpr int EDGE_SLOTS = 8;    // edges a CTA keeps in flight (refill maps 4 lanes to a slot: 8 x 4 = one warp): short edges share one tile of 32 interpolation points
constexpr int EDGE_MISC = 8 * EDGE_SLOTS + 2 * 32 + 8;
// configuration kernel, two-phase tiles: [0..63] pool slot -> configuration index (int64 x 32), [64..95] lane -> pool
// slot of the current tile, [96..103] control words
constexpr int CFG_MISC = 64 + 32 + 8;

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// kind: 0 = configuration kernel (single pass), 1 = edge kernel, 2 = configuration kernel with two-phase tiles,
// 3 = configuration kernel (single pass) with LAT_WARPS warps on one tile (single queries), 4 = the same for the edge kernel
__host__ __device__ inline size_t smem_layout(int blob_words, int D, int world_words, int n_shapes, int kind, size_t* off) {
    const bool edges = kind == 1 || kind == 4, pool = kind == 2;
    const int queue_warps = kind >= 3 ? LAT_WARPS : MAX_WARPS;
    size_t o = 0;
    off[0] = o; o = align16(o + size_t(blob_words) * 4);
    off[1] = o; o = align16(o + size_t(TILE) * D * 4);
    off[2] = o; o = align16(o + (edges ? 0 : size_t(TILE) * D * 4));  // second configuration buffer: configuration kernel only
    off[3] = o; o = align16(o + size_t(world_words) * TILE * 4);
    off[4] = o; o = align16(o + size_t(2) * TILE * 4);
    off[5] = o; o = align16(o + size_t(queue_warps) * QCAP_WORDS * 4);
    off[6] = o; o = align16(o + size_t(n_shapes));
    off[7] = o; o = align16(o + 3 * 8);
    off[8] = o; o = align16(o + (edges ? EDGE_MISC * 4 : pool ? CFG_MISC * 4 : 0));
    off[9] = o; o = align16(o + (edges ? size_t(EDGE_SLOTS) * 2 * D * 8 : pool ? size_t(TILE) * D * 4 : 0));
    return o;
}

__device__ __forceinline__ Smem carve(unsigned char* base, int blob_words, int D, int world_words, int n_shapes, int kind) {
    size_t off[10];
    smem_layout(blob_words, D, world_words, n_shapes, kind, off);
    Smem s;
    s.blob = (uint32_t*)(base + off[0]);
    s.q[0] = (float*)(base + off[1]);
    s.q[1] = (float*)(base + off[2]);
    s.W = (float*)(base + off[3]);
    s.pen_fx = (unsigned*)(base + off[4]);
    s.relf = s.pen_fx + TILE;
    s.queue = (uint32_t*)(base + off[5]);
    s.sflag = (uint8_t*)(base + off[6]);
    s.bar = (uint64_t*)(base + off[7]);
    s.misc = (int*)(base + off[8]);
    s.ed = (double*)(base + off[9]);
    return s;
}

size_t scene_smem_bytes(int blob_words, int D, int world_words, int n_shapes, int kind) {
    size_t off[10];
    return smem_layout(blob_words, D, world_words, n_shapes, kind, off);
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

template <int WARPS>
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
            float s = 0.f, co = 1.f;
            // one copy of sincosf for the four rotating joint types (each inlined copy carries its own large-argument
            // slow path: three copies less of it in the instruction stream)
            if (code <= MRB_J_TRANS_XY_PHI) sincosf(q[code == MRB_J_TRANS_XY_PHI ? qi + 2 : qi], &s, &co);
            switch (code) {  // warp-uniform
                case MRB_J_HINGE_X: rot_cols(R, 1, 2, co, s); break;
                case MRB_J_HINGE_Y: rot_cols(R, 2, 0, co, s); break;
                case MRB_J_HINGE_Z: rot_cols(R, 0, 1, co, s); break;
                case MRB_J_TRANS_XY_PHI: {
                    float x = q[qi], y = q[qi + 1];
#pragma unroll
                    for (int r = 0; r < 3; r++) t[r] = fmaf(R[r * 3], x, fmaf(R[r * 3 + 1], y, t[r]));
                    rot_cols(R, 0, 1, co, s);
                } break;
                // (constant column per case: a run-time column index would push R into local memory)
                case MRB_J_TRANS_X: {
                    const float x = q[qi];
#pragma unroll
                    for (int r = 0; r < 3; r++) t[r] = fmaf(R[r * 3], x, t[r]);
                } break;
                case MRB_J_TRANS_Y: {
                    const float x = q[qi];
#pragma unroll
                    for (int r = 0; r < 3; r++) t[r] = fmaf(R[r * 3 + 1], x, t[r]);
                } break;
                case MRB_J_TRANS_Z: {
                    const float x = q[qi];
#pragma unroll
                    for (int r = 0; r < 3; r++) t[r] = fmaf(R[r * 3 + 2], x, t[r]);
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
                if (core == MRB_CORE_SEG) {  // stored as midpoint + half vector
                    float pa[3], pb[3];
                    xform_point(R, t, L, pa, 1);
                    xform_point(R, t, L + 3, pb, 1);
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        w[k * TILE] = 0.5f * (pa[k] + pb[k]);
                        w[(3 + k) * TILE] = 0.5f * (pb[k] - pa[k]);
                    }
                    continue;
                }
                xform_point(R, t, L, w, TILE);
                if (core == MRB_CORE_BOX) {
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

// FK as an out-of-line call (two-phase kernel: phase A and the pooled single-pass tiles share ONE copy of the FK code;
// with two inlined copies the kernel ran out of instruction cache -- ncu: 2.8 no-instruction stall cycles per issue
// on the dual-arm scene).  Shared-memory byte offsets instead of pointers, see TileArgs below.
template <int WARPS>
__device__ __noinline__ void fk_call(uint32_t o_blob, uint32_t o_q, uint32_t o_W, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fk_phase<WARPS>(reinterpret_cast<const uint32_t*>(smem_raw + o_blob), reinterpret_cast<const float*>(smem_raw + o_q) + lane * D,
                    reinterpret_cast<float*>(smem_raw + o_W), warp, lane);
}

// ------------------------------------------------------------------------------------------
// phase 2: broadphase (lane = configuration, uniform over the pair list) -> per-warp queue of
// surviving (configuration, pair) items -> exact narrowphase, 32 queued items at a time
// ------------------------------------------------------------------------------------------
constexpr int QCAP = QCAP_WORDS;          // queue entries per warp
constexpr float PEN_SCALE = 67108864.f;   // penetration accumulates in units of 2^-26 m (order independent)
constexpr float CULL_SLACK = 1e-3f;       // bounding-volume tests keep everything closer than 1 mm

struct TileCtx {
    const uint32_t* bi;
    const float* bf;
    const float* W;
    const uint8_t* sflag;
    unsigned* pen_fx;  // [TILE]
    unsigned* relf;    // [TILE]
    uint32_t* queue;   // [QCAP], this warp's
    int offS, nmov, lane;
    bool rule;

    __device__ __forceinline__ int row(int s) const { return offS + s * MRB_SHAPE_WORDS; }
    __device__ __forceinline__ float radius(int s) const { return bf[row(s) + 3]; }
    __device__ __forceinline__ float bound_r(int s) const { return bf[row(s) + 4 + 15]; }
    __device__ __forceinline__ const float* rowdata(int s) const { return bf + row(s) + 4; }
    // world data of shape s for configuration cfg (moving: [word][cfg] in W; static: blob row)
    template <int NW>
    __device__ __forceinline__ void load(int s, int cfg, float* out) const {
        const int r = row(s);
        const bool mov = s < nmov;
        const float* p = mov ? W + bi[r + 2] * TILE + cfg : bf + r + 4;
        const int st = mov ? TILE : 1;
#pragma unroll
        for (int k = 0; k < NW; k++) out[k] = p[k * st];
    }
    // same, for a shape index that is uniform across the warp (lane = configuration): real
    // branches and plain LDS instead of per-lane pointer selects
    template <int NW>
    __device__ __forceinline__ void load_u(int s, float* out) const {
        const int r = row(s);
        if (s < nmov) {
            const float* p = W + bi[r + 2] * TILE + lane;
#pragma unroll
            for (int k = 0; k < NW; k++) out[k] = p[k * TILE];
        } else {
            const float* p = bf + r + 4;
#pragma unroll
            for (int k = 0; k < NW; k++) out[k] = p[k];
        }
    }
    __device__ __forceinline__ void add_pen(int cfg, int a, int b, float d) const {
        if (d < 0.f) {
            atomicAdd(&pen_fx[cfg], (unsigned)fmaf(-d, PEN_SCALE, 0.5f));
            if (rule) {
                const unsigned f = sflag[a] | sflag[b];
                if ((f & 1u) && !(f & 2u)) atomicOr(&relf[cfg], 1u);
            }
        }
    }
};

// What crosses an out-of-line call: byte offsets into smem_raw instead of pointers.
struct TileArgs {
    uint32_t o_blob, o_W, o_sflag, o_pen, o_queue;
    bool rule;
};
__device__ __forceinline__ TileCtx make_ctx(const TileArgs& a) {
    const uint32_t* bi = reinterpret_cast<const uint32_t*>(smem_raw + a.o_blob);
    unsigned* pen = reinterpret_cast<unsigned*>(smem_raw + a.o_pen);
    return TileCtx{bi, reinterpret_cast<const float*>(bi), reinterpret_cast<const float*>(smem_raw + a.o_W),
                   smem_raw + a.o_sflag, pen, pen + TILE, reinterpret_cast<uint32_t*>(smem_raw + a.o_queue),
                   (int)bi[MRB_H_OFF_SHAPES], (int)bi[MRB_H_NMOV], (int)(threadIdx.x & 31), a.rule};
}

// segments live in W as (midpoint, half vector); static blob rows hold (a, b)
__device__ __forceinline__ void load_seg(const TileCtx& c, int s, int cfg, float* ab) {
    float v[6];
    c.load<6>(s, cfg, v);
    if (s < c.nmov) {
#pragma unroll
        for (int k = 0; k < 3; k++) { ab[k] = v[k] - v[3 + k]; ab[3 + k] = v[k] + v[3 + k]; }
    } else {
#pragma unroll
        for (int k = 0; k < 6; k++) ab[k] = v[k];
    }
}
// bounding-sphere centre (+ half vector of a segment) of a warp-uniform shape index
template <int CORE, bool WANT_HALF>
__device__ __forceinline__ void load_centre_u(const TileCtx& c, int s, float* ctr, float* hv) {
    if constexpr (CORE == MRB_CORE_SEG) {
        if (s < c.nmov) {  // (midpoint, half vector) in W
            if constexpr (WANT_HALF) {
                float v[6];
                c.load_u<6>(s, v);
#pragma unroll
                for (int k = 0; k < 3; k++) { ctr[k] = v[k]; hv[k] = v[3 + k]; }
            } else {
                c.load_u<3>(s, ctr);
            }
        } else {  // static rows hold the end points
            float v[6];
            c.load_u<6>(s, v);
#pragma unroll
            for (int k = 0; k < 3; k++) { ctr[k] = 0.5f * (v[k] + v[3 + k]); hv[k] = 0.5f * (v[3 + k] - v[k]); }
        }
    } else {
        c.load_u<3>(s, ctr);
    }
}

template <int T>
__device__ __forceinline__ float narrow_pair(const TileCtx& c, int a, int b, int cfg) {
    const float rs = c.radius(a) + c.radius(b);
    if constexpr (T == MRB_PT_SEG_SEG) {
        float A[6], B[6];
        load_seg(c, a, cfg, A);
        load_seg(c, b, cfg, B);
        return d_seg_seg(A, B, rs);
    } else if constexpr (T == MRB_PT_SEG_BOX) {
        float A[6], B[12];
        load_seg(c, a, cfg, A);
        c.load<12>(b, cfg, B);
        return d_seg_box(A, B, B + 3, c.rowdata(b) + 12, rs);
    } else if constexpr (T == MRB_PT_POINT_POINT) {
        float A[3], B[3];
        c.load<3>(a, cfg, A);
        c.load<3>(b, cfg, B);
        return d_point_point(A, B, rs);
    } else if constexpr (T == MRB_PT_POINT_SEG) {
        float A[3], B[6];
        c.load<3>(a, cfg, A);
        load_seg(c, b, cfg, B);
        return d_point_seg(A, B, rs);
    } else if constexpr (T == MRB_PT_POINT_BOX) {
        float A[3], B[12];
        c.load<3>(a, cfg, A);
        c.load<12>(b, cfg, B);
        return d_point_box(A, B, B + 3, c.rowdata(b) + 12, rs);
    } else if constexpr (T == MRB_PT_BOX_BOX) {
        float A[12], B[12];
        c.load<12>(a, cfg, A);
        c.load<12>(b, cfg, B);
        return d_box_box(A, A + 3, c.rowdata(a) + 12, B, B + 3, c.rowdata(b) + 12, rs);
    } else if constexpr (T == MRB_PT_CYLZ_CYLZ) {
        float A[3], B[3];
        c.load<3>(a, cfg, A);
        c.load<3>(b, cfg, B);
        return d_cylz_cylz(A, c.rowdata(a)[3], c.rowdata(a)[4], B, c.rowdata(b)[3], c.rowdata(b)[4]);
    } else {
        float A[12], B[3];
        c.load<12>(a, cfg, A);
        c.load<3>(b, cfg, B);
        return d_box_cylz(A, A + 3, c.rowdata(a) + 12, B, c.rowdata(b)[3], c.rowdata(b)[4]);
    }
}

// queue entry: bits 0..4 configuration, bits 5..28 record index inside its sublist, bits 29..30 sublist
template <int T>
__device__ __noinline__ void drain(const TileArgs args, uint32_t entry, bool valid) {
    if (valid) {
        const TileCtx c = make_ctx(args);
        const int cfg = entry & 31;
        // record r of the blob has its packed pair ids at ids[r] (blob tail, global memory)
        // the pair ids of the entry's record: staged with the blob prefix on small scenes, else read from the blob's
        // tail in global memory (its address is parked in the staged header; keeping the pointer in TileArgs instead
        // costs two registers across the broadphase loops and 1-2 % on the small scenes)
        const uint32_t w = c.bi[MRB_H_BP_IDS + T * MRB_BP_SUBLISTS + (int)(entry >> 29)] + ((entry >> 5) & 0xffffffu);
        uint32_t pk;
        if (c.bi[MRB_H_IDS_STAGED]) pk = c.bi[w];   // warp-uniform
        else pk = (*reinterpret_cast<const uint32_t* const*>(c.bi + MRB_H_GPTR))[w];
        const int a = pk & 0xffff, b = (pk >> 16) & 0xfff;
        c.add_pen(cfg, a, b, narrow_pair<T>(c, a, b, cfg));
    }
}

__device__ __noinline__ void drain_any(const TileArgs c, int type, uint32_t entry, bool valid) {
    switch (type) {  // warp-uniform
        case MRB_PT_SEG_SEG: drain<MRB_PT_SEG_SEG>(c, entry, valid); break;
        case MRB_PT_SEG_BOX: drain<MRB_PT_SEG_BOX>(c, entry, valid); break;
        case MRB_PT_POINT_POINT: drain<MRB_PT_POINT_POINT>(c, entry, valid); break;
        case MRB_PT_POINT_SEG: drain<MRB_PT_POINT_SEG>(c, entry, valid); break;
        case MRB_PT_POINT_BOX: drain<MRB_PT_POINT_BOX>(c, entry, valid); break;
        default: drain<MRB_PT_BOX_BOX>(c, entry, valid); break;
    }
}

// Per-warp survivor queue.  The broadphase marks, per lane (= configuration), which of the last
// <= 32 records passed in a bit mask; flush() turns the masks into queue entries, one round
// per set bit, and drains 32 entries at a time through the exact narrowphase.
struct Survivors {
    const TileCtx c;  // a copy: a reference would force the context into local memory
    const TileArgs args;
    int type, qn;
    __device__ __forceinline__ void flush(uint32_t mask, int first_record, int sub) {
        const unsigned lt = (1u << c.lane) - 1u;
        for (;;) {
            const unsigned m = __ballot_sync(FULL, mask != 0u);
            if (!m) break;
            if (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1u;
                c.queue[qn + __popc(m & lt)] = ((uint32_t)sub << 29) | ((uint32_t)(first_record + j) << 5) | (uint32_t)c.lane;
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= DRAIN_BATCH * TILE) {
                uint32_t entry[DRAIN_BATCH];
#pragma unroll
                for (int b = 0; b < DRAIN_BATCH; b++) entry[b] = c.queue[b * TILE + c.lane];
                const uint32_t tail = c.queue[DRAIN_BATCH * TILE + c.lane];
                __syncwarp();
                qn -= DRAIN_BATCH * TILE;
                if (c.lane < qn) c.queue[c.lane] = tail;
                __syncwarp();
#pragma unroll
                for (int b = 0; b < DRAIN_BATCH; b++) drain_any(args, type, entry[b], true);
            }
        }
    }
    __device__ __forceinline__ void finish() {
#pragma unroll 1
        for (int b = 0; b * TILE < qn; b++) {
            const uint32_t entry = c.queue[b * TILE + c.lane];
            drain_any(args, type, entry, b * TILE + c.lane < qn);
        }
        qn = 0;
        __syncwarp();
    }
};

// One routine for every queued pair type: the broadphase only needs the records.
template <int WARPS>
__device__ __noinline__ void run_queued_types(const TileArgs args, int warp, bool skip_decided, unsigned tol_fx, bool boxes_first) {
    const TileCtx c = make_ctx(args);
    const uint32_t* bi = c.bi;
    const float* bf = c.bf;
    const int lane = c.lane;
    const char* Wl = reinterpret_cast<const char*>(c.W + lane);  // this lane's column of W
    const float4* scentre = reinterpret_cast<const float4*>(bf + bi[MRB_H_OFF_SCENTRE]);
    for (int ti = 0; ti <= MRB_PT_BOX_BOX; ++ti) {
        // boxes_first: pairs against boxes (table, floor, objects) before the robot-robot types -- they decide most
        // colliding configurations, and a decided configuration queues nothing for the types that follow
        // (single-pass flag queries on uniform samples: +3.5 % dual-arm, +2 % mobile; the survivors of the two-phase
        // tiles were NOT decided by the table, there the plain order is 2 % faster)
        const int type = !boxes_first ? ti : ti < 3 ? (ti == 0 ? MRB_PT_SEG_BOX : ti == 1 ? MRB_PT_POINT_BOX : MRB_PT_BOX_BOX)
                                                    : MRB_PT_BOX_BOX - ti;
        Survivors sv{c, args, type, 0};  // one queue per pair type: partial batches only at the end of a type
        for (int sub = 0; sub < MRB_BP_SUBLISTS; ++sub) {
            const int n = bi[MRB_H_BP + (type * MRB_BP_SUBLISTS + sub) * 2 + 1];
            if (n == 0) continue;
            const int off = bi[MRB_H_BP + (type * MRB_BP_SUBLISTS + sub) * 2];
            const int lo = (n * warp) / WARPS, hi = (n * (warp + 1)) / WARPS;
            const uint2* rec = reinterpret_cast<const uint2*>(bi + off);
            const bool seg_x = type == MRB_PT_SEG_BOX;
            unsigned prev_x = 0xffffffffu;
            float cx = 0.f, cy = 0.f, cz = 0.f, hx = 0.f, hy = 0.f, hz = 0.f;
            for (int base = lo; base < hi; base += 32) {
                const int cnt = min(32, hi - base);
                uint32_t mask = 0u;
                if (sub == 0) {  // partner moving: bounding spheres (branch-free body: independent iterations overlap)
#pragma unroll 4
                    for (int j = 0; j < cnt; ++j) {
                        const uint2 r = rec[base + j];
                        const float* px = reinterpret_cast<const float*>(Wl + (r.x & 0xffffu));
                        const float* py = reinterpret_cast<const float*>(Wl + (r.x >> 16));
                        const float x = px[0] - py[0], y = px[TILE] - py[TILE], z = px[2 * TILE] - py[2 * TILE];
                        mask |= (dot3(x, y, z, x, y, z) < __uint_as_float(r.y) ? 1u : 0u) << j;
                    }
                } else if (sub == 1) {  // partner static: bounding spheres
#pragma unroll 4
                    for (int j = 0; j < cnt; ++j) {
                        const uint2 r = rec[base + j];
                        const float* px = reinterpret_cast<const float*>(Wl + (r.x & 0xffffu));
                        const float4 cb = scentre[r.x >> 16];
                        const float x = px[0] - cb.x, y = px[TILE] - cb.y, z = px[2 * TILE] - cb.z;
                        mask |= (dot3(x, y, z, x, y, z) < __uint_as_float(r.y) ? 1u : 0u) << j;
                    }
                } else {  // partner is a large static box: separating-axis bound along its face normals
                    for (int j = 0; j < cnt; ++j) {
                        const uint2 r = rec[base + j];
                        const unsigned xo = r.x & 0xffffu;
                        if (xo != prev_x) {
                            const float* px = reinterpret_cast<const float*>(Wl + xo);
                            cx = px[0]; cy = px[TILE]; cz = px[2 * TILE];
                            if (seg_x) { hx = px[3 * TILE]; hy = px[4 * TILE]; hz = px[5 * TILE]; }
                            prev_x = xo;
                        }
                        const float* B = bf + c.row(c.nmov + (r.x >> 16)) + 4;  // c[3], R[9], half[3]
                        const float* R = B + 3;
                        const float x = cx - B[0], y = cy - B[1], z = cz - B[2];
                        const float l0 = fabsf(dot3(R[0], R[3], R[6], x, y, z)) - B[12] - fabsf(dot3(R[0], R[3], R[6], hx, hy, hz));
                        const float l1 = fabsf(dot3(R[1], R[4], R[7], x, y, z)) - B[13] - fabsf(dot3(R[1], R[4], R[7], hx, hy, hz));
                        const float l2 = fabsf(dot3(R[2], R[5], R[8], x, y, z)) - B[14] - fabsf(dot3(R[2], R[5], R[8], hx, hy, hz));
                        mask |= (fmaxf(fmaxf(l0, l1), l2) < __uint_as_float(r.y) ? 1u : 0u) << j;
                    }
                }
                if (skip_decided && c.pen_fx[lane] > tol_fx) mask = 0u;
                sv.flush(mask, base, sub);
            }
        }
        sv.finish();
    }
}

// cheap planar pair types are evaluated directly (no queue): lane = configuration
template <int T, int WARPS>
__device__ __forceinline__ void run_type_direct(const TileCtx& c, int warp) {
    const uint32_t* bi = c.bi;
    const int n = bi[MRB_H_N_PAIRS + T], off = bi[MRB_H_OFF_PAIRS + T];
    const int lo = (n * warp) / WARPS, hi = (n * (warp + 1)) / WARPS;
    float pen = 0.f;
    bool rel = false;
    for (int i = lo; i < hi; ++i) {
        const uint32_t pk = bi[off + i];
        const int a = pk & 0xffff, b = (pk >> 16) & 0xfff;
        const float d = narrow_pair<T>(c, a, b, c.lane);
        if (d < 0.f) {
            pen -= d;
            if (c.rule) {
                const unsigned f = c.sflag[a] | c.sflag[b];
                rel = rel || ((f & 1u) && !(f & 2u));
            }
        }
    }
    if (pen > 0.f) atomicAdd(&c.pen_fx[c.lane], (unsigned)fmaf(pen, PEN_SCALE, 0.5f));
    if (rel) atomicOr(&c.relf[c.lane], 1u);
}

// All THREADS threads call this.  On return warp 0 holds, per lane, the configuration's total
// penetration (return value) and whether a relevant pair penetrates (*relpen_out).
template <int WARPS, bool FK_CALL = false>
__device__ __forceinline__ float process_tile(const Smem& sm, const float* q_tile, int D, float tol, bool early, bool rule,
                                              bool* relpen_out, bool boxes_first = false) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t* bi = sm.blob;
    const float static_pen = reinterpret_cast<const float*>(bi)[MRB_H_STATIC_PEN];
    if (warp == WARPS - 1) {
        sm.pen_fx[lane] = 0u;
        sm.relf[lane] = 0u;
    }
    if constexpr (FK_CALL)
        fk_call<WARPS>((uint32_t)((const unsigned char*)sm.blob - smem_raw), (uint32_t)((const unsigned char*)q_tile - smem_raw),
                       (uint32_t)((const unsigned char*)sm.W - smem_raw), D);
    else
        fk_phase<WARPS>(bi, q_tile + lane * D, sm.W, warp, lane);
    __syncthreads();

    const TileCtx ctx{bi, reinterpret_cast<const float*>(bi), sm.W, sm.sflag, sm.pen_fx, sm.relf, sm.queue + warp * QCAP,
                      (int)bi[MRB_H_OFF_SHAPES], (int)bi[MRB_H_NMOV], lane, rule};
    // a configuration is decided once its accumulated penetration exceeds tol - static part
    const float budget = fmaxf(tol - static_pen, 0.f);
    const unsigned tol_fx = (unsigned)fminf(budget * PEN_SCALE, 4.0e9f);
    const TileArgs args{(uint32_t)((const unsigned char*)sm.blob - smem_raw), (uint32_t)((const unsigned char*)sm.W - smem_raw),
                        (uint32_t)((const unsigned char*)sm.sflag - smem_raw), (uint32_t)((const unsigned char*)sm.pen_fx - smem_raw),
                        (uint32_t)((const unsigned char*)(sm.queue + warp * QCAP) - smem_raw), rule};
    run_queued_types<WARPS>(args, warp, early, tol_fx, boxes_first);
    run_type_direct<MRB_PT_CYLZ_CYLZ, WARPS>(ctx, warp);
    run_type_direct<MRB_PT_BOX_CYLZ, WARPS>(ctx, warp);
    __syncthreads();
    float total = 0.f;
    if (warp == 0) {
        total = static_pen + (float)sm.pen_fx[lane] * (1.f / PEN_SCALE);
        *relpen_out = sm.relf[lane] != 0u;
    }
    return total;
}

// Phase A of the two-phase tiles: FK, then a LOWER bound of the penetration against the large static boxes
// (table, floor: broadphase sublist 2) from a few points of each moving core -- a point's centre, five
// points of a segment, a box's centre.  For any point P of core X, d(X, Y) <= dist(P, core Y) - r_X - r_Y
// (and every primitive routine returns at most -(r_X + r_Y) once the cores touch), so
// sum max(0, r_X + r_Y - dist(P, core Y)) never exceeds the penetration the exact routines report.  No queue,
// no narrowphase, no barrier but the two around the shared accumulator.  Warp 0 returns the bound per lane.
template <int WARPS>
__device__ __forceinline__ float table_phase(const Smem& sm, const float* q_tile, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t* bi = sm.blob;
    const float* bf = reinterpret_cast<const float*>(bi);
    if (warp == WARPS - 1) sm.pen_fx[lane] = 0u;
    fk_call<WARPS>((uint32_t)((const unsigned char*)sm.blob - smem_raw), (uint32_t)((const unsigned char*)q_tile - smem_raw),
                   (uint32_t)((const unsigned char*)sm.W - smem_raw), D);
    __syncthreads();
    const char* Wl = reinterpret_cast<const char*>(sm.W + lane);
    const int offS = bi[MRB_H_OFF_SHAPES], nmov = bi[MRB_H_NMOV];
    float pen = 0.f;
    for (int type = MRB_PT_POINT_BOX; type <= MRB_PT_BOX_BOX; ++type) {
        const int n = bi[MRB_H_BP + (type * MRB_BP_SUBLISTS + 2) * 2 + 1];
        if (n == 0) continue;
        const int off = bi[MRB_H_BP + (type * MRB_BP_SUBLISTS + 2) * 2];
        const uint2* rec = reinterpret_cast<const uint2*>(bi + off);
        const int lo = (n * warp) / WARPS, hi = (n * (warp + 1)) / WARPS;
        for (int i = lo; i < hi; ++i) {
            const uint2 rc = rec[i];
            const uint32_t rx = rc.x;
            // r_X + r_Y: the record's threshold minus the cull slack for points and segments; a moving box's threshold
            // is built from its bounding radius, so its radii come from the pair ids in the blob tail
            float rsum = __uint_as_float(rc.y) - CULL_SLACK;
            if (type == MRB_PT_BOX_BOX) {
                const uint32_t pk = (*reinterpret_cast<const uint32_t* const*>(bi + MRB_H_GPTR))[bi[MRB_H_IDS_BASE] + ((off - (int)bi[MRB_H_REC_BASE]) >> 1) + i];
                rsum = bf[offS + (int)(pk & 0xffff) * MRB_SHAPE_WORDS + 3] + bf[offS + (int)((pk >> 16) & 0xfff) * MRB_SHAPE_WORDS + 3];
            }
            const float* px = reinterpret_cast<const float*>(Wl + (rx & 0xffffu));
            const float* B = bf + offS + (nmov + (int)(rx >> 16)) * MRB_SHAPE_WORDS + 4;  // c[3], R[9], half[3]
            const float* R = B + 3;
            const float x = px[0] - B[0], y = px[TILE] - B[1], z = px[2 * TILE] - B[2];
            const float l0 = dot3(R[0], R[3], R[6], x, y, z), l1 = dot3(R[1], R[4], R[7], x, y, z), l2 = dot3(R[2], R[5], R[8], x, y, z);
            float d2;
            if (type == MRB_PT_SEG_BOX) {  // five points of the segment: midpoint + u * half vector
                const float hx = px[3 * TILE], hy = px[4 * TILE], hz = px[5 * TILE];
                const float h0 = dot3(R[0], R[3], R[6], hx, hy, hz), h1 = dot3(R[1], R[4], R[7], hx, hy, hz), h2 = dot3(R[2], R[5], R[8], hx, hy, hz);
                d2 = 3.0e38f;
#pragma unroll
                for (int k = -2; k <= 2; ++k) {
                    const float u = 0.5f * (float)k;
                    const float a0 = fmaxf(fabsf(fmaf(u, h0, l0)) - B[12], 0.f), a1 = fmaxf(fabsf(fmaf(u, h1, l1)) - B[13], 0.f),
                                a2 = fmaxf(fabsf(fmaf(u, h2, l2)) - B[14], 0.f);
                    d2 = fminf(d2, dot3(a0, a1, a2, a0, a1, a2));
                }
            } else {
                const float a0 = fmaxf(fabsf(l0) - B[12], 0.f), a1 = fmaxf(fabsf(l1) - B[13], 0.f), a2 = fmaxf(fabsf(l2) - B[14], 0.f);
                d2 = dot3(a0, a1, a2, a0, a1, a2);
            }
            pen += fmaxf(rsum - sqrtf(d2), 0.f);
        }
    }
    if (pen > 0.f) atomicAdd(&sm.pen_fx[lane], (unsigned)(fminf(pen, 60.f) * PEN_SCALE));  // rounded down
    __syncthreads();
    float total = 0.f;
    if (warp == 0) total = reinterpret_cast<const float*>(bi)[MRB_H_STATIC_PEN] + (float)sm.pen_fx[lane] * (1.f / PEN_SCALE);
    return total;
}

// common prologue: barriers, blob staging through TMA, A6 shape flags
__device__ __forceinline__ void stage_scene(const Smem& sm, const uint32_t* blob, int blob_words, int n_shapes,
                                            const RobotRule& rr, int THREADS) {
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
    if (threadIdx.x == 0) *reinterpret_cast<const uint32_t**>(sm.blob + MRB_H_GPTR) = blob;  // for readers of the blob's tail
    __syncthreads();
}

// One complete single-pass tile as an out-of-line call (two-phase kernel only: pooled survivors, the final flush and
// the single-pass fallback share ONE copy of FK + pair code instead of three inlined ones; the shared-memory
// pointers are rebuilt from the layout parameters so that they stay shared-space pointers, see TileArgs).
template <int WARPS>
__device__ __noinline__ float full_tile_call(int blob_words, int D, int world_words, int n_shapes, uint32_t q_off, float tol) {
    const Smem sm = carve(smem_raw, blob_words, D, world_words, n_shapes, 2);
    bool relpen;
    return process_tile<WARPS, true>(sm, reinterpret_cast<const float*>(smem_raw + q_off), D, tol, true, false, &relpen);
}

// ------------------------------------------------------------------------------------------
// configuration batch kernel (A5 / A6 batch variant)
// ------------------------------------------------------------------------------------------
// MINB: CTAs per SM the register allocation has to allow.  16 resident warps by default; scenes whose shared-memory
// footprint admits only three 4-warp CTAs anyway (four-arm scene: 65 KB) get the variant compiled for three, i.e. up
// to 168 registers per thread instead of 128.
template <int WARPS, bool TWO_PHASE, int MINB = 16 / WARPS>
__global__ void __launch_bounds__(TILE * WARPS, MINB) check_configs_kernel(ConfigParams p) {
    constexpr int THREADS = TILE * WARPS;
    const Smem sm = carve(smem_raw, p.blob_words, p.D, p.world_words, p.n_shapes, TWO_PHASE ? 2 : WARPS > MAX_WARPS ? 3 : 0);
    stage_scene(sm, p.blob, p.blob_words, p.n_shapes, p.rule, THREADS);

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

    if constexpr (!TWO_PHASE) {
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
            const float total = process_tile<WARPS>(sm, sm.q[buf], D, tol, early, p.rule.enabled, &relpen, early);
            if (warp == 0 && lane < nvalid) {
                const bool coll = p.rule.enabled ? (total > tol && relpen) : (total > tol);
                p.flags[first + lane] = coll ? 0 : 1;
                if (p.pen_out) p.pen_out[first + lane] = total;
            }
            __syncthreads();
        }
        return;
    }
    // Two-phase tiles (plain flag queries on scenes with a table / floor; the launcher picks this variant): phase A
    // runs FK and a cheap lower bound of the penetration against the large static boxes (table_phase), which
    // decides most colliding configurations; the undecided configurations of successive tiles are pooled (q and
    // index) and complete single-pass tiles run on 32 survivors at a time, so their flags are computed exactly as
    // by the single-pass kernel.  A CTA that sees few decisions in phase A falls back to single-pass tiles.
    int64_t* pool_idx = reinterpret_cast<int64_t*>(sm.misc);          // [TILE]
    int* s_map = sm.misc + 64;                                        // [TILE] lane -> pool slot (>= TILE: after the flush)
    int* s_ctl = sm.misc + 96;    // [0] pool fill, [1] configurations seen, [2] decided in phase A, [3] single-pass from now on
    float* pool_q = reinterpret_cast<float*>(sm.ed);                  // [TILE][D]
    if (threadIdx.x < 4) s_ctl[threadIdx.x] = 0;
    int n_large = 0;
    for (int t = MRB_PT_POINT_BOX; t <= MRB_PT_BOX_BOX; ++t) n_large += (int)sm.blob[MRB_H_BP + (t * MRB_BP_SUBLISTS + 2) * 2 + 1];
    const float retire_slack = 2e-6f + 2.5e-7f * (float)n_large;
    __syncthreads();

    // single-pass tile on the pooled survivors (n of them; idle lanes recompute slot 0)
    auto run_pool = [&](int n) {
        if (n < TILE) {
            for (int i = threadIdx.x; i < (TILE - n) * D; i += THREADS) pool_q[n * D + i] = pool_q[i % D];
            __syncthreads();
        }
        const float total = full_tile_call<WARPS>(p.blob_words, D, p.world_words, p.n_shapes,
                                                  (uint32_t)((const unsigned char*)pool_q - smem_raw), tol);
        if (warp == 0 && lane < n) p.flags[pool_idx[lane]] = total > tol ? 0 : 1;
        __syncthreads();
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
        if (s_ctl[3]) {
            const float total = full_tile_call<WARPS>(p.blob_words, D, p.world_words, p.n_shapes,
                                                      (uint32_t)((const unsigned char*)sm.q[buf] - smem_raw), tol);
            if (warp == 0 && lane < nvalid) p.flags[first + lane] = total > tol ? 0 : 1;
            __syncthreads();
            continue;
        }
        // ---- phase A ----
        const float bound_a = table_phase<WARPS>(sm, sm.q[buf], D);
        if (warp == 0) {
            int slot = -1;
            const bool valid = lane < nvalid;
            // the bound is exact arithmetic's lower bound; the slack covers the fp32 rounding of both evaluations
            // (a few 1e-7 per penetrating table pair)
            const bool coll = bound_a - retire_slack > tol;
            if (valid && coll) p.flags[first + lane] = 0;
            const unsigned surv = __ballot_sync(FULL, valid && !coll);
            const int fill = s_ctl[0], room = TILE - fill;
            if ((surv >> lane) & 1u) {
                const int rank = __popc(surv & ((1u << lane) - 1u));
                slot = rank < room ? fill + rank : TILE + rank - room;
                if (slot < TILE) pool_idx[slot] = first + lane;
            }
            s_map[lane] = slot;
            __syncwarp();
            if (lane == 0) {
                const int seen = s_ctl[1] + nvalid, decided = s_ctl[2] + nvalid - __popc(surv);
                s_ctl[0] = fill + __popc(surv);   // > TILE: the pool is flushed once in between
                s_ctl[1] = seen;
                s_ctl[2] = decided;
                if (seen >= 16 * TILE && decided * 4 < seen) s_ctl[3] = 1;   // phase A decides too little to pay for a second FK
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < TILE * D; t += THREADS) {
            const int c = t / D, sl = s_map[c];
            if (sl >= 0 && sl < TILE) pool_q[sl * D + (t - c * D)] = sm.q[buf][t];
        }
        __syncthreads();
        const int fill = s_ctl[0];
        if (fill >= TILE) {
            run_pool(TILE);
            if (warp == 0) {
                const int slot = s_map[lane];
                if (slot >= TILE) pool_idx[slot - TILE] = first + lane;
                if (lane == 0) s_ctl[0] = fill - TILE;
            }
            for (int t = threadIdx.x; t < TILE * D; t += THREADS) {
                const int c = t / D, sl = s_map[c];
                if (sl >= TILE) pool_q[(sl - TILE) * D + (t - c * D)] = sm.q[buf][t];
            }
        }
        __syncthreads();
    }
    const int fill = s_ctl[0];
    if (fill > 0) run_pool(fill);
    if (p.stats && threadIdx.x == 0) {  // feedback for the launcher: how much phase A decides on this mode
        atomicAdd(&p.stats[0], s_ctl[1]);
        atomicAdd(&p.stats[1], s_ctl[2]);
    }
}

// ------------------------------------------------------------------------------------------
// edge batch kernel (A8 batch variant).  A CTA keeps up to EDGE_SLOTS edges in flight and fills every tile
// of 32 interpolation points from them in slot order, so short edges (and short N_start / N_max windows)
// share a tile instead of leaving most lanes idle; a long edge fills whole tiles on its own.  Positions
// follow the reference's binary order; an edge retires at its first colliding position (early exit) or
// when its window is exhausted, and its slot is refilled from the dynamic edge counter.
// ------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(TILE * WARPS, 16 / WARPS) check_edges_kernel(EdgeParams p) {
    constexpr int THREADS = TILE * WARPS;
    constexpr int K = EDGE_SLOTS;
    const Smem sm = carve(smem_raw, p.blob_words, p.D, p.world_words, p.n_shapes, WARPS > MAX_WARPS ? 4 : 1);
    RobotRule none{};
    stage_scene(sm, p.blob, p.blob_words, p.n_shapes, none, THREADS);

    const int D = p.D;
    const float tol = p.tol < 0.f ? reinterpret_cast<const float*>(sm.blob)[MRB_H_TOL] : p.tol;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per slot: edge id (-1 = empty), N, end of the window, next position, lanes taken in this tile, first lane
    int* s_edge = sm.misc;
    int* s_N = sm.misc + K;
    int* s_nmax = sm.misc + 2 * K;
    int* s_cur = sm.misc + 3 * K;
    int* s_take = sm.misc + 4 * K;
    int* s_pref = sm.misc + 5 * K;
    int* s_idx = sm.misc + 8 * K;        // [32] interpolation index of the lane's sample (-1: idle lane)
    int* s_slot = s_idx + 32;            // [32] slot the lane's sample belongs to
    int* s_ctl = s_slot + 32;            // [0] samples in this tile, [1] edge counter exhausted, [2] samples per edge, last claim
    double* e_start = sm.ed;             // [K][D] q1 (fp64)
    double* e_step = sm.ed + K * D;      // [K][D] (q2 - q1) / (N - 1)

    if (threadIdx.x < K) s_edge[threadIdx.x] = -1;
    if (threadIdx.x == 0) { s_ctl[1] = 0; s_ctl[2] = TILE; }
    __syncthreads();

    for (;;) {
        if (warp == 0) {
            // ---- refill: claim only as many edges as it takes to fill the tile's 32 lanes (estimated from the
            // length of the edges claimed last), so long edges stay one per CTA -- edges parked in slots could not
            // be picked up by idle CTAs at the end of the batch -- while short ones are claimed by the handful.
            // Four lanes set up one slot (32 lanes = EDGE_SLOTS x 4).
            for (;;) {
                const int have = __reduce_add_sync(FULL, (lane < K && s_edge[lane] >= 0) ? s_nmax[lane] - s_cur[lane] : 0);
                if (have >= TILE) break;
                unsigned nm = __ballot_sync(FULL, lane < K && s_edge[lane] < 0);
                if (!nm || s_ctl[1]) break;
                const int est = max(s_ctl[2], 1);
                int want = min((TILE - have + est - 1) / est, __popc(nm));
                while (__popc(nm) > want) nm &= ~(0x80000000u >> __clz(nm));  // keep the lowest `want` empty slots
                int base = 0;
                if (lane == 0) base = atomicAdd(p.counter, __popc(nm));
                base = __shfl_sync(FULL, base, 0);
                const int g = lane >> 2, j = lane & 3;
                const int64_t e = (int64_t)base + __popc(nm & ((1u << g) - 1u));
                const bool mine = ((nm >> g) & 1u) && e < p.E;
                // endpoints in fp64, N exactly as the reference: max(2, int(|dq|_inf / resolution) + 1)
                double m = 0.0;
                if (mine) {
                    for (int k = j; k < D; k += 4) {
                        const double a = (double)p.q1[e * D + k], b = (double)p.q2[e * D + k];
                        e_start[g * D + k] = a;
                        e_step[g * D + k] = __dsub_rn(b, a);
                        m = fmax(m, fabs(__dsub_rn(a, b)));
                    }
                }
                m = fmax(m, __shfl_xor_sync(FULL, m, 1));
                m = fmax(m, __shfl_xor_sync(FULL, m, 2));
                int span = 0;
                if (mine) {
                    const int N = p.N ? p.N[e] : max(2, (int)__ddiv_rn(m, p.resolution) + 1);
                    const double inv = (double)(N - 1);
                    for (int k = j; k < D; k += 4) e_step[g * D + k] = __ddiv_rn(e_step[g * D + k], inv);
                    const int nmax = (p.n_max < 0 || p.n_max > N) ? N : p.n_max;
                    if (j == 0) {
                        if (p.n_start < nmax) {
                            s_edge[g] = (int)e;
                            s_N[g] = N;
                            s_nmax[g] = nmax;
                            s_cur[g] = p.n_start;
                            span = nmax - p.n_start;
                        } else {  // empty window: free, nothing to check
                            p.flags[e] = 1;
                            if (p.first_pos) p.first_pos[e] = -1;
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) span += __shfl_xor_sync(FULL, span, o);
                __syncwarp();  // every lane has read s_ctl[1], s_ctl[2] of this round
                if (lane == 0) {
                    s_ctl[2] = span / __popc(nm);  // samples per edge of this claim
                    if ((int64_t)base + __popc(nm) >= p.E) s_ctl[1] = 1;
                }
                __syncwarp();
            }
            // ---- hand the tile's 32 lanes out in slot order ----
            const int rem = (lane < K && s_edge[lane] >= 0) ? s_nmax[lane] - s_cur[lane] : 0;
            int incl = rem;
#pragma unroll
            for (int o = 1; o < K; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const int pref = min(incl - rem, TILE);
            const int take = min(rem, TILE - pref);
            if (lane < K) { s_pref[lane] = pref; s_take[lane] = take; }
            const int total = min(__shfl_sync(FULL, incl, K - 1), TILE);
            // lane -> slot without a search: bit pref_k marks the first lane of every slot that got lanes; a lane's
            // slot is the n-th such slot, n = number of marks at or below the lane
            const bool got = lane < K && take > 0;
            const unsigned starts = __reduce_or_sync(FULL, got ? 1u << pref : 0u);
            const unsigned taken = __ballot_sync(FULL, got);
            int slot = 0, i = -1;
            if (lane < total) {
                const unsigned below = starts & (0xffffffffu >> (31 - lane));  // marks at or below this lane (never empty)
                slot = (int)__fns(taken, 0, __popc(below));
                const int pos = s_cur[slot] + lane - (31 - __clz(below));
                const int N = s_N[slot];
                i = binary_order_index(N, pos);
                if (!p.include_endpoints && (i == 0 || i == N - 1)) i = -1;
            }
            // idle lanes recompute the start point of the last sample's edge (a valid configuration; result unused)
            const int last_slot = __shfl_sync(FULL, slot, total > 0 ? total - 1 : 0);
            s_idx[lane] = i;
            s_slot[lane] = lane < total ? slot : last_slot;
            if (lane == 0) s_ctl[0] = total;
        }
        __syncthreads();
        if (s_ctl[0] == 0) break;  // nothing in flight and the counter is exhausted
        // q = q1 + step * i in fp64 (reference order of operations), rounded once to fp32
        for (int t = threadIdx.x; t < TILE * D; t += THREADS) {
            const int c = t / D, k = t - c * D;
            const int i = s_idx[c];
            const int so = s_slot[c] * D + k;
            sm.q[0][t] = (float)__dadd_rn(e_start[so], __dmul_rn(e_step[so], (double)(i < 0 ? 0 : i)));
        }
        __syncthreads();
        bool relpen;
        // full evaluation of every sample: stopping the decided samples early (as the configuration kernel does) gains
        // 2-13 % on long uniform edges and loses 1-3 % on the short edges planners actually check (measured)
        const float total_pen = process_tile<WARPS>(sm, sm.q[0], D, tol, false, false, &relpen);
        if (warp == 0) {
            const unsigned hit = __ballot_sync(FULL, s_idx[lane] >= 0 && total_pen > tol);
            if (lane < K && s_take[lane] > 0) {
                const int take = s_take[lane], pref = s_pref[lane];
                const unsigned range = (take >= 32 ? FULL : ((1u << take) - 1u)) << pref;
                const unsigned m = hit & range;
                const int e = s_edge[lane];
                if (m) {  // first colliding position of this edge: earlier tiles were clean
                    p.flags[e] = 0;
                    if (p.first_pos) p.first_pos[e] = s_cur[lane] + (__ffs(m) - 1 - pref);
                    s_edge[lane] = -1;
                } else {
                    const int cur = s_cur[lane] + take;
                    s_cur[lane] = cur;
                    if (cur >= s_nmax[lane]) {
                        p.flags[e] = 1;
                        if (p.first_pos) p.first_pos[e] = -1;
                        s_edge[lane] = -1;
                    }
                }
            }
            __syncwarp();  // warp 0 goes straight on to refill and hand out the next tile; the others wait at the
                           // barrier that follows it (no CTA-wide barrier needed here: only warp 0 touches the slots)
        }
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
static int grid_for(K kernel, int threads, size_t smem) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (occ < 1) occ = 1;
    return occ * num_sms();
}

// warps per tile: small scenes run two warps per tile (balanced FK for two-robot scenes, cheap
// barriers); scenes with a large per-configuration footprint share a tile among four warps to
// keep enough warps resident
static int warps_per_tile(int world_words) { return world_words >= 128 ? 4 : 2; }

cudaError_t launch_static_penetration(uint32_t* blob, cudaStream_t st) {
    static_penetration_kernel<<<1, 32, 0, st>>>(blob);
    return cudaGetLastError();
}

cudaError_t launch_check_configs(const ConfigParams& p, cudaStream_t st) {
    if (p.B <= 0) return cudaSuccess;
    const bool two = p.two_phase && !p.full_eval && !p.rule.enabled && !p.pen_out;
    const int64_t n_tiles = (p.B + TILE - 1) / TILE;
    static const bool lat_off = getenv("MRB200_NO_LATENCY_KERNEL") != nullptr;   // measurement aid
    const bool latency = n_tiles == 1 && !two && !lat_off;
    const size_t smem = scene_smem_bytes(p.blob_words, p.D, p.world_words, p.n_shapes, two ? 2 : latency ? 3 : 0);
#define MRB_LAUNCH_CONFIGS(W, ...)                                                      \
    do {                                                                                \
        int grid = grid_for(check_configs_kernel<W, __VA_ARGS__>, 32 * W, smem);        \
        if (grid > n_tiles) grid = (int)n_tiles;                                        \
        check_configs_kernel<W, __VA_ARGS__><<<grid, 32 * W, smem, st>>>(p);            \
    } while (0)
    if (latency) {
        MRB_LAUNCH_CONFIGS(LAT_WARPS, false);
    } else if (warps_per_tile(p.world_words) == 4) {
        const bool three = 4 * (smem + 1024) > 228 * 1024;   // shared memory admits three CTAs per SM at most
        if (two && three) MRB_LAUNCH_CONFIGS(4, true, 3);
        else if (two) MRB_LAUNCH_CONFIGS(4, true);
        else if (three) MRB_LAUNCH_CONFIGS(4, false, 3);
        else MRB_LAUNCH_CONFIGS(4, false);
    } else {
        if (two) MRB_LAUNCH_CONFIGS(2, true);
        else MRB_LAUNCH_CONFIGS(2, false);
    }
#undef MRB_LAUNCH_CONFIGS
    return cudaGetLastError();
}

cudaError_t launch_check_edges(const EdgeParams& p, cudaStream_t st) {
    if (p.E <= 0) return cudaSuccess;
    static const bool lat_off = getenv("MRB200_NO_LATENCY_KERNEL") != nullptr;   // measurement aid
    const bool latency = p.E <= 2 && !lat_off;    // single edge queries of the reference planners: eight warps per tile
    const size_t smem = scene_smem_bytes(p.blob_words, p.D, p.world_words, p.n_shapes, latency ? 4 : 1);
    cudaError_t err = cudaMemsetAsync(p.counter, 0, sizeof(int), st);
    if (err != cudaSuccess) return err;
    if (latency) {
        int grid = grid_for(check_edges_kernel<LAT_WARPS>, 32 * LAT_WARPS, smem);
        if (grid > p.E) grid = (int)p.E;
        check_edges_kernel<LAT_WARPS><<<grid, 32 * LAT_WARPS, smem, st>>>(p);
    } else if (warps_per_tile(p.world_words) == 4) {
        int grid = grid_for(check_edges_kernel<4>, 128, smem);
        if (grid > p.E) grid = (int)p.E;
        check_edges_kernel<4><<<grid, 128, smem, st>>>(p);
    } else {
        int grid = grid_for(check_edges_kernel<2>, 64, smem);
        if (grid > p.E) grid = (int)p.E;
        check_edges_kernel<2><<<grid, 64, smem, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace mrb
