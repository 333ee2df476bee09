// extern "C" entry points declared in include/mrb200.h.  No torch types, no allocation on the
// data path: the caller owns every device buffer and the stream.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/mrb200.h"
#include "kernels.h"
#include "knn_common.cuh"
#include "scene_blob.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? MRB200_ERR_NO_DEVICE : MRB200_ERR_CUDA,
                "%s: %s", what, cudaGetErrorString(e));
}

struct ModeSlot {
    uint32_t* blob = nullptr;  // device
    int words = 0, staged = 0, D = 0, world_words = 0, n_shapes = 0, n_pairs = 0;   // staged: prefix kept in shared memory
    int n_large = 0;           // pairs against large static boxes (table, floor): phase A of the two-phase tiles
    // two-phase tiles pay off when phase A decides most configurations; that depends on the caller's inputs, so the
    // first large batches measure it (device counters copied to pinned memory behind the launch, read without
    // a synchronisation by a later call): 0 = measuring, 1 = keep, 2 = single-pass kernel from now on
    mutable int two_phase_state = 0;
    bool two_phase_forced = false;   // set through mrb200_scene_set_two_phase: no measuring
};

}  // namespace

// The edge scheduler claims edges from a device counter.  Calls on one handle may come from several streams / host
// threads, so every call takes its own counter from a ring (a counter is reused only after EDGE_COUNTERS further
// calls on the handle; each use is reset on the caller's stream right before the launch).
constexpr int EDGE_COUNTERS = 1024;

struct mrb200_scene {
    std::vector<ModeSlot> slots;
    int* counter = nullptr;  // device scratch for the edge scheduler: ring of EDGE_COUNTERS ints
    std::atomic<unsigned> counter_next{0};
    // asynchronous host-buffer edge batches (mrb200_submit_edges_host / mrb200_collect_edges_host): a ring of tickets,
    // each with its own pinned + device staging and an event; guarded by `mu`
    struct EdgeTicket {
        cudaEvent_t ev = nullptr;
        unsigned char* dev = nullptr;
        unsigned char* pin = nullptr;
        size_t bytes = 0, out_off = 0, nb = 0;
        int64_t E = 0, seq = -1;
        bool pending = false;
    };
    static constexpr int N_TICKETS = 64;
    EdgeTicket tickets[N_TICKETS];
    int64_t ticket_seq = 0;
    int* stats_dev = nullptr;   // [2 * slots] (configurations seen, decided in phase A) per mode slot
    int* stats_pin = nullptr;   // pinned copy, refreshed behind two-phase launches while a slot is still measuring
    // staging of the host-buffer query entry points (mrb200_query_*_host): one device and one pinned host
    // buffer, grown on demand, guarded by `mu`
    std::mutex mu;
    unsigned char* stage_dev = nullptr;
    unsigned char* stage_pin = nullptr;      // pinned + mapped host memory
    unsigned char* stage_pin_dev = nullptr;  // its device-side alias: tiny queries are read / written in place
    size_t stage_bytes = 0;
    // chunk pipeline of mrb200_check_configs_host: HOST_PIPE side streams, each with device staging for one chunk of
    // configurations and flags and (for pageable callers) a pinned bounce buffer; guarded by `mu`
#ifndef MRB_HOST_PIPE
#define MRB_HOST_PIPE 3
#endif
    static constexpr int HOST_PIPE = MRB_HOST_PIPE;
    cudaStream_t pipe_stream[HOST_PIPE] = {};
    cudaEvent_t pipe_start = nullptr, pipe_done[HOST_PIPE] = {};
    unsigned char* pipe_dev[HOST_PIPE] = {};
    unsigned char* pipe_pin[HOST_PIPE] = {};
    size_t pipe_bytes = 0, pipe_pin_bytes = 0;
};

struct mrb200_abstract {
    mrb::AbstractSceneData data;
    int* counter = nullptr;   // ring of EDGE_COUNTERS ints, one per call (see mrb200_scene)
    std::atomic<unsigned> counter_next{0};
    // mapped pinned staging of the host-buffer queries (mrb200_abstract_query_*_host), guarded by `mu`
    std::mutex mu;
    unsigned char* stage_pin = nullptr;
    unsigned char* stage_pin_dev = nullptr;
    size_t stage_bytes = 0;
};

extern "C" {

int mrb200_version(void) { return 100 + MRB_BLOB_VERSION; }
const char* mrb200_last_error(void) { return g_err; }
int64_t mrb200_launch_count(void) { return g_launches.load(); }

int mrb200_fp32_probe(int iters, float* out_dev, int32_t* n_threads, mrb200_stream_t stream) {
    int n = 0;
    cudaError_t e = mrb::launch_fp32_probe(iters, out_dev, &n, (cudaStream_t)stream);
    if (n_threads) *n_threads = n;
    if (e != cudaSuccess) return cuda_fail(e, "fp32_probe");
    if (out_dev) g_launches++;
    return MRB200_OK;
}

// ------------------------------------------------------------------ abstract environment
int mrb200_abstract_create(int n_agents, int dim, const double* radii, int n_sph, const double* spheres, int n_rect,
                           const double* rects, mrb200_abstract_t** out) {
    if (!out || !radii || n_agents < 1 || n_agents > mrb::ABS_MAX_AGENTS || dim < 1 || dim > mrb::ABS_MAX_DIM ||
        n_sph < 0 || n_sph > mrb::ABS_MAX_OBS || n_rect < 0 || n_rect > mrb::ABS_MAX_OBS || (n_sph && !spheres) ||
        (n_rect && !rects))
        return fail(MRB200_ERR_ARG, "abstract_create: unsupported sizes (agents<=%d dim<=%d obstacles<=%d per kind)",
                    mrb::ABS_MAX_AGENTS, mrb::ABS_MAX_DIM, mrb::ABS_MAX_OBS);
    auto* env = new mrb200_abstract();
    memset(&env->data, 0, sizeof(env->data));
    env->data.n_agents = n_agents;
    env->data.dim = dim;
    env->data.n_sph = n_sph;
    env->data.n_rect = n_rect;
    for (int i = 0; i < n_agents; i++) env->data.radii[i] = radii[i];
    for (int o = 0; o < n_sph; o++) {
        for (int k = 0; k < dim; k++) env->data.sph_c[o][k] = spheres[o * (dim + 1) + k];
        env->data.sph_r[o] = spheres[o * (dim + 1) + dim];
    }
    for (int o = 0; o < n_rect; o++)
        for (int k = 0; k < dim; k++) {
            env->data.rect_min[o][k] = rects[o * 2 * dim + k];
            env->data.rect_max[o][k] = rects[o * 2 * dim + dim + k];
        }
    cudaError_t e = cudaMalloc(&env->counter, sizeof(int) * EDGE_COUNTERS);
    if (e != cudaSuccess) {
        delete env;
        return cuda_fail(e, "abstract_create");
    }
    *out = env;
    return MRB200_OK;
}

int mrb200_abstract_destroy(mrb200_abstract_t* env) {
    if (!env) return MRB200_OK;
    cudaFree(env->counter);
    cudaFreeHost(env->stage_pin);
    delete env;
    return MRB200_OK;
}

int mrb200_abstract_check_configs(const mrb200_abstract_t* env, const double* q, int64_t B, uint8_t* free_dev,
                                  mrb200_stream_t stream) {
    if (!env || B < 0 || (B && (!q || !free_dev))) return fail(MRB200_ERR_ARG, "abstract_check_configs: bad argument");
    if (B == 0) return MRB200_OK;
    cudaError_t e = mrb::launch_abstract_configs(env->data, q, B, free_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "abstract_check_configs");
    g_launches++;
    return MRB200_OK;
}

int mrb200_abstract_check_edges(const mrb200_abstract_t* env_c, const double* q1, const double* q2, int64_t E,
                                double resolution, const int32_t* N_dev, int32_t n_start, int32_t n_max,
                                int include_endpoints, uint8_t* free_dev, int32_t* first_pos_dev, mrb200_stream_t stream) {
    mrb200_abstract_t* env = const_cast<mrb200_abstract_t*>(env_c);   // (only the counter ring index is touched)
    if (!env || E < 0 || (E && (!q1 || !q2 || !free_dev)) || n_start < 0 || (!N_dev && !(resolution > 0)))
        return fail(MRB200_ERR_ARG, "abstract_check_edges: bad argument");
    if (E == 0) return MRB200_OK;
    int* counter = env->counter + env->counter_next.fetch_add(1, std::memory_order_relaxed) % EDGE_COUNTERS;
    cudaError_t e = mrb::launch_abstract_edges(env->data, q1, q2, E, resolution, N_dev, n_start, n_max, include_endpoints,
                                               free_dev, first_pos_dev, counter, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "abstract_check_edges");
    g_launches++;
    return MRB200_OK;
}

static int abstract_stage(mrb200_abstract_t* env, size_t bytes, cudaStream_t st) {
    if (bytes <= env->stage_bytes) return MRB200_OK;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "abstract query: sync");
    cudaFreeHost(env->stage_pin);
    env->stage_pin = env->stage_pin_dev = nullptr;
    env->stage_bytes = 0;
    size_t cap = 4096;
    while (cap < bytes) cap *= 2;
    e = cudaHostAlloc(&env->stage_pin, cap, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&env->stage_pin_dev, env->stage_pin, 0);
    if (e != cudaSuccess) return cuda_fail(e, "abstract query: cudaHostAlloc");
    env->stage_bytes = cap;
    return MRB200_OK;
}

int mrb200_abstract_query_configs_host(mrb200_abstract_t* env, const double* q_host, int64_t B, uint8_t* free_host,
                                       mrb200_stream_t stream) {
    if (!env || B < 0 || (B && (!q_host || !free_host))) return fail(MRB200_ERR_ARG, "abstract_query_configs_host: bad argument");
    if (B == 0) return MRB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(env->mu);
    const size_t D = (size_t)env->data.n_agents * env->data.dim;
    const size_t qb = ((size_t)B * D * 8 + 15) & ~size_t(15);
    int rc = abstract_stage(env, qb + (size_t)B, st);
    if (rc) return rc;
    memcpy(env->stage_pin, q_host, (size_t)B * D * 8);
    rc = mrb200_abstract_check_configs(env, (const double*)env->stage_pin_dev, B, env->stage_pin_dev + qb, stream);
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "abstract_query_configs_host");
    memcpy(free_host, env->stage_pin + qb, (size_t)B);
    return MRB200_OK;
}

int mrb200_abstract_query_edges_host(mrb200_abstract_t* env, const double* q1_host, const double* q2_host, int64_t E,
                                     double resolution, const int32_t* N_host, int32_t n_start, int32_t n_max,
                                     int include_endpoints, uint8_t* free_host, int32_t* first_pos_host, mrb200_stream_t stream) {
    if (!env || E < 0 || (E && (!q1_host || !q2_host || !free_host)))
        return fail(MRB200_ERR_ARG, "abstract_query_edges_host: bad argument");
    if (E == 0) return MRB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(env->mu);
    const size_t D = (size_t)env->data.n_agents * env->data.dim;
    const size_t qb = ((size_t)E * D * 8 + 15) & ~size_t(15), nb = ((size_t)E * 4 + 15) & ~size_t(15);
    int rc = abstract_stage(env, 2 * qb + 2 * nb + (size_t)E, st);   // q1 | q2 | N | first | free
    if (rc) return rc;
    memcpy(env->stage_pin, q1_host, (size_t)E * D * 8);
    memcpy(env->stage_pin + qb, q2_host, (size_t)E * D * 8);
    if (N_host) memcpy(env->stage_pin + 2 * qb, N_host, (size_t)E * 4);
    unsigned char* d = env->stage_pin_dev;
    rc = mrb200_abstract_check_edges(env, (const double*)d, (const double*)(d + qb), E, resolution,
                                     N_host ? (const int32_t*)(d + 2 * qb) : nullptr, n_start, n_max, include_endpoints,
                                     d + 2 * qb + 2 * nb, (int32_t*)(d + 2 * qb + nb), stream);
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "abstract_query_edges_host");
    if (first_pos_host) memcpy(first_pos_host, env->stage_pin + 2 * qb + nb, (size_t)E * 4);
    memcpy(free_host, env->stage_pin + 2 * qb + 2 * nb, (size_t)E);
    return MRB200_OK;
}

// ------------------------------------------------------------------ primitive scenes
int mrb200_scene_create(int max_modes, mrb200_scene_t** out) {
    if (!out || max_modes < 1 || max_modes > 65536) return fail(MRB200_ERR_ARG, "scene_create: bad argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(MRB200_ERR_NO_DEVICE, "scene_create: no CUDA device (%s); there is no CPU fallback",
                                                cudaGetErrorString(e));
    auto* sc = new mrb200_scene();
    sc->slots.resize(max_modes);
    e = cudaMalloc(&sc->counter, sizeof(int) * EDGE_COUNTERS);
    if (e == cudaSuccess) e = cudaMalloc(&sc->stats_dev, 2 * sizeof(int) * max_modes);
    if (e == cudaSuccess) e = cudaMemset(sc->stats_dev, 0, 2 * sizeof(int) * max_modes);
    if (e == cudaSuccess) e = cudaHostAlloc(&sc->stats_pin, 2 * sizeof(int) * max_modes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaFree(sc->counter);
        cudaFree(sc->stats_dev);
        delete sc;
        return cuda_fail(e, "scene_create");
    }
    memset(sc->stats_pin, 0, 2 * sizeof(int) * max_modes);
    *out = sc;
    return MRB200_OK;
}

int mrb200_scene_destroy(mrb200_scene_t* sc) {
    if (!sc) return MRB200_OK;
    for (auto& s : sc->slots) cudaFree(s.blob);
    cudaFree(sc->counter);
    cudaFree(sc->stats_dev);
    cudaFreeHost(sc->stats_pin);
    cudaFree(sc->stage_dev);
    cudaFreeHost(sc->stage_pin);
    for (int i = 0; i < mrb200_scene::HOST_PIPE; i++) {
        if (sc->pipe_stream[i]) cudaStreamDestroy(sc->pipe_stream[i]);
        if (sc->pipe_done[i]) cudaEventDestroy(sc->pipe_done[i]);
        cudaFree(sc->pipe_dev[i]);
        cudaFreeHost(sc->pipe_pin[i]);
    }
    if (sc->pipe_start) cudaEventDestroy(sc->pipe_start);
    for (auto& t : sc->tickets) {
        if (t.ev) cudaEventDestroy(t.ev);
        cudaFree(t.dev);
        cudaFreeHost(t.pin);
    }
    delete sc;
    return MRB200_OK;
}

int mrb200_scene_set_mode(mrb200_scene_t* sc, int slot, const void* blob_host, size_t nbytes, mrb200_stream_t stream) {
    if (!sc || slot < 0 || slot >= (int)sc->slots.size() || !blob_host) return fail(MRB200_ERR_ARG, "scene_set_mode: bad argument");
    const uint32_t* h = (const uint32_t*)blob_host;
    if (nbytes < MRB_HDR_WORDS * 4 || h[MRB_H_MAGIC] != MRB_BLOB_MAGIC || h[MRB_H_VERSION] != MRB_BLOB_VERSION ||
        (size_t)h[MRB_H_TOTAL_WORDS] * 4 != nbytes || (nbytes & 15))
        return fail(MRB200_ERR_BLOB, "scene_set_mode: not a version-%d scene blob of %zu bytes", MRB_BLOB_VERSION, nbytes);
    if (h[MRB_H_STAGED_WORDS] < MRB_HDR_WORDS || h[MRB_H_STAGED_WORDS] > h[MRB_H_TOTAL_WORDS] || (h[MRB_H_STAGED_WORDS] & 3) ||
        h[MRB_H_REC_BASE] > h[MRB_H_STAGED_WORDS] || h[MRB_H_IDS_BASE] < h[MRB_H_REC_BASE] || h[MRB_H_IDS_BASE] > h[MRB_H_TOTAL_WORDS])
        return fail(MRB200_ERR_BLOB, "scene_set_mode: inconsistent staged / tail split in the blob header");
    const int n_shapes = (int)(h[MRB_H_NMOV] + h[MRB_H_NSTA]);
    if (n_shapes > 256) return fail(MRB200_ERR_ARG, "scene_set_mode: more than 256 collision shapes");
    // everything that can reject the blob is checked BEFORE the slot is touched
    const int dof = (int)h[MRB_H_DOF];
    if (dof < 1 || dof > 64) return fail(MRB200_ERR_ARG, "scene_set_mode: %d degrees of freedom (supported: 1..64)", dof);
    if (mrb::scene_smem_bytes((int)h[MRB_H_STAGED_WORDS], dof, (int)h[MRB_H_WORLD_WORDS], n_shapes, 1) > 227 * 1024 ||
        mrb::scene_smem_bytes((int)h[MRB_H_STAGED_WORDS], dof, (int)h[MRB_H_WORLD_WORDS], n_shapes, 2) > 227 * 1024)
        return fail(MRB200_ERR_ARG, "scene_set_mode: scene needs more than 227 KB of shared memory per CTA");
    ModeSlot& s = sc->slots[slot];
    cudaStream_t st = (cudaStream_t)stream;
    // Replacing a live slot: kernels launched earlier on OTHER streams (check_configs_host's side streams, other host
    // threads) may still read the old blob, so wait for the whole device, not only for `stream`.
    if (s.blob) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return cuda_fail(e, "scene_set_mode");
    }
    if (s.words != (int)h[MRB_H_TOTAL_WORDS]) {
        cudaFree(s.blob);
        s.blob = nullptr;
        s.words = 0;          // the slot is empty until the new blob is fully in place
        cudaError_t e = cudaMalloc(&s.blob, nbytes);
        if (e != cudaSuccess) {
            s.blob = nullptr;
            return cuda_fail(e, "scene_set_mode: cudaMalloc");
        }
    }
    auto reset_slot = [&]() {
        cudaFree(s.blob);
        s.blob = nullptr;
        s.words = 0;
    };
    cudaError_t e = cudaMemcpyAsync(s.blob, blob_host, nbytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // blob_host may be pageable and freed by the caller
    if (e != cudaSuccess) {
        reset_slot();
        return cuda_fail(e, "scene_set_mode: copy");
    }
    s.words = (int)h[MRB_H_TOTAL_WORDS];
    s.staged = (int)h[MRB_H_STAGED_WORDS];
    s.D = dof;
    s.world_words = (int)h[MRB_H_WORLD_WORDS];
    s.n_shapes = n_shapes;
    s.n_pairs = 0;
    for (int t = 0; t < MRB_NUM_PAIR_TYPES; t++) s.n_pairs += (int)h[MRB_H_N_PAIRS + t];
    s.n_large = 0;
    for (int t = 0; t < MRB_NUM_PAIR_TYPES; t++) s.n_large += (int)h[MRB_H_BP + (t * MRB_BP_SUBLISTS + 2) * 2 + 1];
    s.two_phase_state = 0;
    s.two_phase_forced = false;
    sc->stats_pin[2 * slot] = sc->stats_pin[2 * slot + 1] = 0;
    e = cudaMemsetAsync(sc->stats_dev + 2 * slot, 0, 2 * sizeof(int), st);
    if (e == cudaSuccess) e = mrb::launch_static_penetration(s.blob, st);
    if (e != cudaSuccess) {
        reset_slot();
        return cuda_fail(e, "scene_set_mode: static pairs");
    }
    g_launches++;
    return MRB200_OK;
}

static const ModeSlot* get_slot(const mrb200_scene_t* sc, int slot) {
    if (!sc || slot < 0 || slot >= (int)sc->slots.size() || !sc->slots[slot].blob) return nullptr;
    return &sc->slots[slot];
}

int mrb200_scene_info(const mrb200_scene_t* sc, int slot, int32_t* out4) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s || !out4) return fail(MRB200_ERR_ARG, "scene_info: empty slot");
    out4[0] = s->D;
    out4[1] = s->n_shapes;
    out4[2] = s->n_pairs;
    out4[3] = (int32_t)mrb::scene_smem_bytes(s->staged, s->D, s->world_words, s->n_shapes);
    return MRB200_OK;
}

int mrb200_scene_set_two_phase(mrb200_scene_t* sc, int slot, int policy) {
    if (!get_slot(sc, slot) || policy < 0 || policy > 2) return fail(MRB200_ERR_ARG, "scene_set_two_phase: bad argument");
    ModeSlot& s = sc->slots[slot];
    s.two_phase_state = policy;
    s.two_phase_forced = policy != 0;
    if (policy == 0) {
        cudaError_t e = cudaDeviceSynchronize();  // a counter copy of an earlier launch may still be in flight
        if (e == cudaSuccess) e = cudaMemset(sc->stats_dev + 2 * slot, 0, 2 * sizeof(int));
        if (e != cudaSuccess) return cuda_fail(e, "scene_set_two_phase");
        sc->stats_pin[2 * slot] = sc->stats_pin[2 * slot + 1] = 0;
    }
    return MRB200_OK;
}

int mrb200_scene_get_two_phase(const mrb200_scene_t* sc, int slot, int32_t* out3) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s || !out3) return fail(MRB200_ERR_ARG, "scene_get_two_phase: empty slot");
    out3[0] = s->two_phase_state;
    out3[1] = sc->stats_pin[2 * slot];
    out3[2] = sc->stats_pin[2 * slot + 1];
    return MRB200_OK;
}

constexpr int64_t MRB200_ZERO_COPY_MAX = 64;  // host-buffer queries up to this size run on mapped host memory in place

static int check_configs_impl(const mrb200_scene_t* sc, int slot, const float* q, int64_t B, float tol, uint8_t* free_dev,
                              float* pen_dev, int full_eval, const mrb::RobotRule& rule, mrb200_stream_t stream,
                              bool allow_bulk = true) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "check_configs: empty mode slot %d", slot);
    if (B < 0 || (B && (!q || !free_dev))) return fail(MRB200_ERR_ARG, "check_configs: bad argument");
    if (B == 0) return MRB200_OK;
    mrb::ConfigParams p{};
    p.blob = s->blob;
    p.blob_words = s->staged;
    p.D = s->D;
    p.world_words = s->world_words;
    p.n_shapes = s->n_shapes;
    p.q = q;
    p.B = B;
    p.tol = tol;
    p.flags = free_dev;
    p.pen_out = pen_dev;
    p.full_eval = full_eval;
    p.bulk_ok = allow_bulk && (((uintptr_t)q) & 15) == 0 && ((s->D * 4 * 32) % 16 == 0);
    p.rule = rule;
    // Two-phase tiles: small batches leave nothing to pool; otherwise the first large batches measure what phase A
    // decides (state 0) and the slot settles on one kernel.  MRB200_TWO_PHASE=0 / 1 forces the choice (measurement aid).
    static const int forced = [] { const char* v = getenv("MRB200_TWO_PHASE"); return v ? (v[0] == '1' ? 1 : 0) : -1; }();
    const bool eligible = s->n_large > 0 && B >= 4096 && !pen_dev && !full_eval && !rule.enabled;
    bool measuring = false;
    if (eligible && forced < 0) {
        // (several host threads may query one handle: the state is a relaxed atomic; a lost race only repeats the decision)
        int state = __atomic_load_n(&s->two_phase_state, __ATOMIC_RELAXED);
        if (state == 0 && !s->two_phase_forced) {
            // (seen, decided) land together as one 8-byte copy behind every measuring launch; read them with ONE 64-bit
            // load: two 4-byte reads could pair the `seen` of an older copy with the `decided` of a newer one and settle
            // the slot on the wrong kernel (seen in round 2: the dual-arm headline ran two-phase tiles, 3.0 instead of 2.7 ms)
            const uint64_t both = __atomic_load_n(reinterpret_cast<const uint64_t*>(sc->stats_pin + 2 * slot), __ATOMIC_RELAXED);
            const int seen = (int)(uint32_t)both, decided = (int)(uint32_t)(both >> 32);
            if (seen >= 4096) {
                state = (int64_t)decided * 20 >= (int64_t)seen * 11 ? 1 : 2;
                __atomic_store_n(&s->two_phase_state, state, __ATOMIC_RELAXED);
            }
        }
        p.two_phase = state != 2;
        measuring = state == 0 && !s->two_phase_forced;
    } else {
        p.two_phase = eligible && forced == 1;
    }
    p.stats = measuring ? sc->stats_dev + 2 * slot : nullptr;
    cudaError_t e = mrb::launch_check_configs(p, (cudaStream_t)stream);
    if (e == cudaSuccess && measuring)
        e = cudaMemcpyAsync(sc->stats_pin + 2 * slot, sc->stats_dev + 2 * slot, 2 * sizeof(int), cudaMemcpyDeviceToHost,
                            (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "check_configs");
    g_launches++;
    return MRB200_OK;
}

int mrb200_check_configs(const mrb200_scene_t* sc, int slot, const float* q, int64_t B, float tol, uint8_t* free_dev,
                         float* pen_dev, int full_eval, mrb200_stream_t stream) {
    mrb::RobotRule none{};
    return check_configs_impl(sc, slot, q, B, tol, free_dev, pen_dev, full_eval, none, stream);
}

int mrb200_check_configs_for_robot(const mrb200_scene_t* sc, int slot, const float* q, int64_t B, float tol,
                                   const uint8_t* relevant_host, const uint8_t* other_host, int n_shapes, uint8_t* free_dev,
                                   mrb200_stream_t stream) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "check_configs_for_robot: empty mode slot %d", slot);
    if (!relevant_host || !other_host || n_shapes != s->n_shapes)
        return fail(MRB200_ERR_ARG, "check_configs_for_robot: need %d shape flags", s->n_shapes);
    mrb::RobotRule rule{};
    rule.enabled = 1;
    for (int i = 0; i < n_shapes; i++) {
        if (relevant_host[i]) rule.rel[i >> 6] |= 1ull << (i & 63);
        if (other_host[i]) rule.oth[i >> 6] |= 1ull << (i & 63);
    }
    return check_configs_impl(sc, slot, q, B, tol, free_dev, nullptr, 1, rule, stream);
}

int mrb200_check_edges(const mrb200_scene_t* sc, int slot, const float* q1, const float* q2, int64_t E, double resolution,
                       const int32_t* N_dev, int32_t n_start, int32_t n_max, int include_endpoints, float tol,
                       uint8_t* free_dev, int32_t* first_pos_dev, mrb200_stream_t stream) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "check_edges: empty mode slot %d", slot);
    if (E < 0 || (E && (!q1 || !q2 || !free_dev)) || n_start < 0 || (!N_dev && !(resolution > 0)))
        return fail(MRB200_ERR_ARG, "check_edges: bad argument");
    if (E == 0) return MRB200_OK;
    mrb::EdgeParams p{};
    p.blob = s->blob;
    p.blob_words = s->staged;
    p.D = s->D;
    p.world_words = s->world_words;
    p.n_shapes = s->n_shapes;
    p.q1 = q1;
    p.q2 = q2;
    p.E = E;
    p.resolution = resolution;
    p.N = N_dev;
    p.n_start = n_start;
    p.n_max = n_max;
    p.include_endpoints = include_endpoints;
    p.tol = tol;
    p.flags = free_dev;
    p.first_pos = first_pos_dev;
    p.counter = sc->counter + const_cast<mrb200_scene_t*>(sc)->counter_next.fetch_add(1, std::memory_order_relaxed) % EDGE_COUNTERS;
    cudaError_t e = mrb::launch_check_edges(p, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "check_edges");
    g_launches++;
    return MRB200_OK;
}

// host-buffer queries: stage -> H2D -> kernel -> D2H -> synchronise, all inside the call
static int stage_reserve(mrb200_scene_t* sc, size_t bytes, cudaStream_t st) {
    if (bytes <= sc->stage_bytes) return MRB200_OK;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "query: sync");
    cudaFree(sc->stage_dev);
    cudaFreeHost(sc->stage_pin);
    sc->stage_dev = sc->stage_pin = sc->stage_pin_dev = nullptr;
    sc->stage_bytes = 0;
    size_t cap = 4096;
    while (cap < bytes) cap *= 2;
    e = cudaMalloc(&sc->stage_dev, cap);
    if (e != cudaSuccess) return cuda_fail(e, "query: cudaMalloc");
    e = cudaHostAlloc(&sc->stage_pin, cap, cudaHostAllocMapped);
    if (e != cudaSuccess) return cuda_fail(e, "query: cudaHostAlloc");
    e = cudaHostGetDevicePointer((void**)&sc->stage_pin_dev, sc->stage_pin, 0);
    if (e != cudaSuccess) return cuda_fail(e, "query: cudaHostGetDevicePointer");
    sc->stage_bytes = cap;
    return MRB200_OK;
}

static size_t up16(size_t x) { return (x + 15) & ~size_t(15); }

int mrb200_query_configs_host(mrb200_scene_t* sc, int slot, const float* q_host, int64_t B, float tol,
                              const uint8_t* relevant_host, const uint8_t* other_host, int n_shapes, uint8_t* free_host,
                              mrb200_stream_t stream) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "query_configs_host: empty mode slot %d", slot);
    if (B < 0 || (B && (!q_host || !free_host)) || ((relevant_host != nullptr) != (other_host != nullptr)))
        return fail(MRB200_ERR_ARG, "query_configs_host: bad argument");
    if (B == 0) return MRB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(sc->mu);
    const size_t qb = up16((size_t)B * s->D * 4), fb = up16((size_t)B);
    int rc = stage_reserve(sc, qb + fb, st);
    if (rc) return rc;
    memcpy(sc->stage_pin, q_host, (size_t)B * s->D * 4);
    // a handful of queries: the kernel reads the configurations from, and writes the flags to, the mapped host
    // buffer directly -- two copy calls less per query; larger batches go through device staging
    const bool in_place = B <= MRB200_ZERO_COPY_MAX;
    cudaError_t e = cudaSuccess;
    if (!in_place) e = cudaMemcpyAsync(sc->stage_dev, sc->stage_pin, (size_t)B * s->D * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "query_configs_host: H2D");
    unsigned char* base = in_place ? sc->stage_pin_dev : sc->stage_dev;
    const float* qd = (const float*)base;
    uint8_t* fd = base + qb;
    mrb::RobotRule rule{};
    if (relevant_host) {
        if (n_shapes != s->n_shapes) return fail(MRB200_ERR_ARG, "query_configs_host: need %d shape flags", s->n_shapes);
        rule.enabled = 1;
        for (int i = 0; i < n_shapes; i++) {
            if (relevant_host[i]) rule.rel[i >> 6] |= 1ull << (i & 63);
            if (other_host[i]) rule.oth[i >> 6] |= 1ull << (i & 63);
        }
    }
    rc = check_configs_impl(sc, slot, qd, B, tol, fd, nullptr, relevant_host ? 1 : 0, rule, stream, /*allow_bulk=*/!in_place);
    if (rc) return rc;
    if (!in_place) e = cudaMemcpyAsync(sc->stage_pin + qb, fd, (size_t)B, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "query_configs_host: D2H");
    memcpy(free_host, sc->stage_pin + qb, (size_t)B);
    return MRB200_OK;
}

// Large host batches: the batch is cut into chunks that travel H2D, through the configuration kernel and back D2H on
// HOST_PIPE side streams, so the copy of one chunk overlaps the kernel and the read-back of its neighbours.  With pinned
// caller buffers the copies go straight from / to them; pageable buffers are bounced through pinned staging (then the
// host memcpy is the limit).  Returns after every flag has landed in free_host.  (Round 2 ran this loop in Python over
// torch streams: ~45 us of interpreter time per chunk limit the chunk size to 256k configurations and the pipeline's tail
// to a 0.17 ms kernel; here a chunk costs three asynchronous driver calls.)
int mrb200_check_configs_host(mrb200_scene_t* sc, int slot, const float* q_host, int64_t B, float tol, uint8_t* free_host,
                              int64_t chunk, mrb200_stream_t stream) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "check_configs_host: empty mode slot %d", slot);
    if (B < 0 || (B && (!q_host || !free_host))) return fail(MRB200_ERR_ARG, "check_configs_host: bad argument");
    if (B == 0) return MRB200_OK;
    if (chunk <= 0) chunk = 1 << 18;
    chunk = (chunk + 31) / 32 * 32;
    if (chunk > B) chunk = (B + 31) / 32 * 32;
    constexpr int NP = mrb200_scene::HOST_PIPE;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(sc->mu);
    cudaError_t e = cudaSuccess;
    if (!sc->pipe_start) {
        e = cudaEventCreateWithFlags(&sc->pipe_start, cudaEventDisableTiming);
        for (int i = 0; i < NP && e == cudaSuccess; i++) {
            e = cudaStreamCreateWithFlags(&sc->pipe_stream[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sc->pipe_done[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: streams");
    }
    const size_t qb = up16((size_t)chunk * s->D * 4), need = qb + up16((size_t)chunk);
    if (need > sc->pipe_bytes) {
        for (int i = 0; i < NP; i++) {
            cudaStreamSynchronize(sc->pipe_stream[i]);
            cudaFree(sc->pipe_dev[i]);
            sc->pipe_dev[i] = nullptr;
        }
        sc->pipe_bytes = 0;
        for (int i = 0; i < NP; i++) {
            e = cudaMalloc(&sc->pipe_dev[i], need);
            if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: cudaMalloc");
        }
        sc->pipe_bytes = need;
    }
    // pinned (or registered) caller memory can be copied asynchronously; anything else goes through a pinned bounce buffer
    auto pinned = [](const void* p) {
        cudaPointerAttributes a{};
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    const bool direct = pinned(q_host) && pinned(free_host);
    if (!direct && need > sc->pipe_pin_bytes) {
        for (int i = 0; i < NP; i++) {
            cudaStreamSynchronize(sc->pipe_stream[i]);
            cudaFreeHost(sc->pipe_pin[i]);
            sc->pipe_pin[i] = nullptr;
        }
        sc->pipe_pin_bytes = 0;
        for (int i = 0; i < NP; i++) {
            e = cudaHostAlloc(&sc->pipe_pin[i], need, cudaHostAllocDefault);
            if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: cudaHostAlloc");
        }
        sc->pipe_pin_bytes = need;
    }
    // the side streams start behind the caller's stream
    e = cudaEventRecord(sc->pipe_start, st);
    for (int i = 0; i < NP && e == cudaSuccess; i++) e = cudaStreamWaitEvent(sc->pipe_stream[i], sc->pipe_start, 0);
    if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: start");
    mrb::RobotRule rule{};
    int64_t done_upto[NP];   // pageable callers: first row / rows of the chunk whose flags wait in pipe_pin[k]
    int64_t done_rows[NP];
    for (int i = 0; i < NP; i++) done_upto[i] = done_rows[i] = 0;
    int k = 0;
    for (int64_t first = 0; first < B; first += chunk, k = (k + 1) % NP) {
        const int64_t n = std::min<int64_t>(chunk, B - first);
        cudaStream_t ps = sc->pipe_stream[k];
        float* qd = (float*)sc->pipe_dev[k];
        uint8_t* fd = sc->pipe_dev[k] + qb;
        if (direct) {
            e = cudaMemcpyAsync(qd, q_host + first * s->D, (size_t)n * s->D * 4, cudaMemcpyHostToDevice, ps);
        } else {
            // the bounce buffer of this stream is free once its previous chunk has come back
            e = cudaStreamSynchronize(ps);
            if (e == cudaSuccess && done_rows[k]) memcpy(free_host + done_upto[k], sc->pipe_pin[k] + qb, (size_t)done_rows[k]);
            done_rows[k] = 0;
            memcpy(sc->pipe_pin[k], q_host + first * s->D, (size_t)n * s->D * 4);
            if (e == cudaSuccess) e = cudaMemcpyAsync(qd, sc->pipe_pin[k], (size_t)n * s->D * 4, cudaMemcpyHostToDevice, ps);
        }
        if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: H2D");
        const int rc = check_configs_impl(sc, slot, qd, n, tol, fd, nullptr, 0, rule, (mrb200_stream_t)ps, /*allow_bulk=*/true);
        if (rc) return rc;
        if (direct) {
            e = cudaMemcpyAsync(free_host + first, fd, (size_t)n, cudaMemcpyDeviceToHost, ps);
        } else {
            e = cudaMemcpyAsync(sc->pipe_pin[k] + qb, fd, (size_t)n, cudaMemcpyDeviceToHost, ps);
            done_upto[k] = first;
            done_rows[k] = n;
        }
        if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: D2H");
    }
    for (int i = 0; i < NP; i++) {
        e = cudaEventRecord(sc->pipe_done[i], sc->pipe_stream[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, sc->pipe_done[i], 0);   // later work on the caller's stream is ordered behind
        if (e == cudaSuccess) e = cudaStreamSynchronize(sc->pipe_stream[i]);
        if (e != cudaSuccess) return cuda_fail(e, "check_configs_host: finish");
        if (!direct && done_rows[i]) memcpy(free_host + done_upto[i], sc->pipe_pin[i] + qb, (size_t)done_rows[i]);
    }
    return MRB200_OK;
}

int mrb200_query_edges_host(mrb200_scene_t* sc, int slot, const float* q1_host, const float* q2_host, int64_t E,
                            double resolution, const int32_t* N_host, int32_t n_start, int32_t n_max, int include_endpoints,
                            float tol, uint8_t* free_host, int32_t* first_pos_host, mrb200_stream_t stream) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "query_edges_host: empty mode slot %d", slot);
    if (E < 0 || (E && (!q1_host || !q2_host || !free_host))) return fail(MRB200_ERR_ARG, "query_edges_host: bad argument");
    if (E == 0) return MRB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(sc->mu);
    const size_t qb = up16((size_t)E * s->D * 4), nb = up16((size_t)E * 4), fb = up16((size_t)E);
    // layout: q1 | q2 | N | first | free   (the inputs are one contiguous H2D copy, the outputs one D2H copy)
    int rc = stage_reserve(sc, 2 * qb + 2 * nb + fb, st);
    if (rc) return rc;
    memcpy(sc->stage_pin, q1_host, (size_t)E * s->D * 4);
    memcpy(sc->stage_pin + qb, q2_host, (size_t)E * s->D * 4);
    if (N_host) memcpy(sc->stage_pin + 2 * qb, N_host, (size_t)E * 4);
    const bool in_place = E <= MRB200_ZERO_COPY_MAX;
    cudaError_t e = cudaSuccess;
    if (!in_place) e = cudaMemcpyAsync(sc->stage_dev, sc->stage_pin, 2 * qb + (N_host ? nb : 0), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "query_edges_host: H2D");
    unsigned char* d = in_place ? sc->stage_pin_dev : sc->stage_dev;
    rc = mrb200_check_edges(sc, slot, (const float*)d, (const float*)(d + qb), E, resolution, N_host ? (const int32_t*)(d + 2 * qb) : nullptr,
                            n_start, n_max, include_endpoints, tol, d + 2 * qb + 2 * nb, (int32_t*)(d + 2 * qb + nb), stream);
    if (rc) return rc;
    if (!in_place) e = cudaMemcpyAsync(sc->stage_pin + 2 * qb + nb, d + 2 * qb + nb, nb + fb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "query_edges_host: D2H");
    if (first_pos_host) memcpy(first_pos_host, sc->stage_pin + 2 * qb + nb, (size_t)E * 4);
    memcpy(free_host, sc->stage_pin + 2 * qb + 2 * nb, (size_t)E);
    return MRB200_OK;
}

// Asynchronous edge batches with host buffers: submit copies the inputs into a library-owned pinned buffer, queues
// H2D copy + kernel + D2H copy on `stream`, records an event and returns at once with a ticket; collect waits for the
// event and hands the results out.  The caller (env.py's speculation of a PRM node's candidate edges) overlaps the
// device work with the planner's own host work.  A ticket stays valid until N_TICKETS further submits on the handle.
int mrb200_submit_edges_host(mrb200_scene_t* sc, int slot, const float* q1_host, int q1_rows, const float* q2_host, int64_t E,
                             double resolution, const int32_t* N_host, int include_endpoints, float tol, int64_t* ticket_out,
                             mrb200_stream_t stream) {
    const ModeSlot* s = get_slot(sc, slot);
    if (!s) return fail(MRB200_ERR_ARG, "submit_edges_host: empty mode slot %d", slot);
    if (E < 1 || !q1_host || !q2_host || !ticket_out || (q1_rows != 1 && q1_rows != E))
        return fail(MRB200_ERR_ARG, "submit_edges_host: bad argument (q1 is one row or one row per edge)");
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(sc->mu);
    const int64_t seq = sc->ticket_seq++;
    auto& t = sc->tickets[seq % mrb200_scene::N_TICKETS];
    cudaError_t e = cudaSuccess;
    if (!t.ev) e = cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming);
    if (e == cudaSuccess && t.pending) e = cudaEventSynchronize(t.ev);   // an uncollected older batch: its results are dropped
    if (e != cudaSuccess) return cuda_fail(e, "submit_edges_host: event");
    t.pending = false;
    const size_t row = (size_t)s->D * 4;
    const size_t qb = up16((size_t)E * row), nb = up16((size_t)E * 4), fb = up16((size_t)E);
    const size_t need = 2 * qb + 2 * nb + fb;     // q1 | q2 | N | first | free
    if (need > t.bytes) {
        cudaFree(t.dev);
        cudaFreeHost(t.pin);
        t.dev = t.pin = nullptr;
        t.bytes = 0;
        size_t cap = 1 << 16;
        while (cap < need) cap *= 2;
        e = cudaMalloc(&t.dev, cap);
        if (e == cudaSuccess) e = cudaHostAlloc(&t.pin, cap, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaFree(t.dev);
            t.dev = nullptr;
            return cuda_fail(e, "submit_edges_host: staging");
        }
        t.bytes = cap;
    }
    if (q1_rows == 1) for (int64_t i = 0; i < E; i++) memcpy(t.pin + (size_t)i * row, q1_host, row);
    else memcpy(t.pin, q1_host, (size_t)E * row);
    memcpy(t.pin + qb, q2_host, (size_t)E * row);
    if (N_host) memcpy(t.pin + 2 * qb, N_host, (size_t)E * 4);
    e = cudaMemcpyAsync(t.dev, t.pin, 2 * qb + (N_host ? nb : 0), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "submit_edges_host: H2D");
    int rc = mrb200_check_edges(sc, slot, (const float*)t.dev, (const float*)(t.dev + qb), E, resolution,
                                N_host ? (const int32_t*)(t.dev + 2 * qb) : nullptr, 0, -1, include_endpoints, tol,
                                t.dev + 2 * qb + 2 * nb, (int32_t*)(t.dev + 2 * qb + nb), stream);
    if (rc) return rc;
    e = cudaMemcpyAsync(t.pin + 2 * qb + nb, t.dev + 2 * qb + nb, nb + fb, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(t.ev, st);
    if (e != cudaSuccess) return cuda_fail(e, "submit_edges_host: D2H");
    t.E = E;
    t.seq = seq;
    t.out_off = 2 * qb + nb;
    t.nb = nb;
    t.pending = true;
    *ticket_out = seq;
    return MRB200_OK;
}

int mrb200_collect_edges_host(mrb200_scene_t* sc, int64_t ticket, int64_t E, uint8_t* free_host, int32_t* first_pos_host) {
    if (!sc || ticket < 0 || !free_host) return fail(MRB200_ERR_ARG, "collect_edges_host: bad argument");
    std::lock_guard<std::mutex> lock(sc->mu);
    auto& t = sc->tickets[ticket % mrb200_scene::N_TICKETS];
    if (t.seq != ticket || !t.pending || t.E != E)
        return fail(MRB200_ERR_ARG, "collect_edges_host: ticket %lld expired or already collected", (long long)ticket);
    cudaError_t e = cudaEventSynchronize(t.ev);
    if (e != cudaSuccess) return cuda_fail(e, "collect_edges_host");
    if (first_pos_host) memcpy(first_pos_host, t.pin + t.out_off, (size_t)E * 4);
    memcpy(free_host, t.pin + t.out_off + t.nb, (size_t)E);
    t.pending = false;
    return MRB200_OK;
}

// ------------------------------------------------------------------ distances / neighbours
static int make_slices(const int32_t* slices_host, int R, int D, int metric, mrb::Slices* out) {
    if (D < 1 || D > mrb::KNN_MAX_D) return fail(MRB200_ERR_ARG, "D must be in [1, %d]", mrb::KNN_MAX_D);
    if (metric < 0 || metric > 3) return fail(MRB200_ERR_ARG, "unknown metric %d", metric);
    memset(out, 0, sizeof(*out));
    if (metric == MRB200_METRIC_SUM_EUCLIDEAN || metric == MRB200_METRIC_MAX_EUCLIDEAN) {
        if (!slices_host || R < 1 || R > mrb::KNN_MAX_R) return fail(MRB200_ERR_ARG, "need 1..%d robot slices", mrb::KNN_MAX_R);
        out->R = R;
        for (int r = 0; r < R; r++) {
            out->start[r] = slices_host[2 * r];
            out->end[r] = slices_host[2 * r + 1];
            if (out->start[r] < 0 || out->end[r] > D || out->start[r] >= out->end[r]) return fail(MRB200_ERR_ARG, "bad slice %d", r);
        }
    }
    return MRB200_OK;
}

int mrb200_batch_dist(const double* q, const double* pts, int64_t N, int D, const int32_t* slices_host, int R, int metric,
                      double* out_dev, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, metric, &sl)) return rc;
    if (N < 0 || (N && (!q || !pts || !out_dev))) return fail(MRB200_ERR_ARG, "batch_dist: bad argument");
    if (N == 0) return MRB200_OK;
    cudaError_t e = mrb::launch_batch_dist(q, pts, N, D, sl, metric, out_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "batch_dist");
    g_launches++;
    return MRB200_OK;
}

static bool tc_usable(int64_t Q, int64_t N, int D, const mrb::Slices& sl, int metric, int k, mrb::TcPlan* plan) {
    if (k > mrb::knn_tc_max_k() || N < 1024 || Q < 1) return false;
    if (!mrb::knn_tc_make_plan(D, sl, metric, plan)) return false;
    return mrb::knn_tc_smem_bytes(*plan, 0) <= 224 * 1024;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

int mrb200_batch_cost(const double* a, int a_is_single, const double* b, int64_t N, int D, const int32_t* slices_host, int R,
                      int per_robot_max, int reduction_sum, double w, double* out_dev, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, MRB200_METRIC_MAX_EUCLIDEAN, &sl)) return rc;
    if (N < 0 || (N && (!a || !b || !out_dev))) return fail(MRB200_ERR_ARG, "batch_cost: bad argument");
    if (N == 0) return MRB200_OK;
    cudaError_t e = mrb::launch_batch_cost(a, a_is_single ? 0 : D, b, N, D, sl, per_robot_max, reduction_sum, w, out_dev,
                                           (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "batch_cost");
    g_launches++;
    return MRB200_OK;
}

int mrb200_minplus_cost(const double* a, int64_t T1, const double* b, const double* lb_b, int64_t T2, int D, const int32_t* slices_host,
                        int R, int per_robot_max, int reduction_sum, double w, double* out_dev, int32_t* arg_dev, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, MRB200_METRIC_MAX_EUCLIDEAN, &sl)) return rc;
    if (T1 < 0 || T2 < 0 || (T1 && (!a || !out_dev)) || (T2 && (!b || !lb_b))) return fail(MRB200_ERR_ARG, "minplus_cost: bad argument");
    if (T1 == 0) return MRB200_OK;
    cudaError_t e = mrb::launch_minplus_cost(a, T1, b, lb_b, T2, D, sl, per_robot_max, reduction_sum, w, out_dev, arg_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "minplus_cost");
    g_launches++;
    return MRB200_OK;
}

static size_t tc_workspace_bytes(int64_t Q, int64_t N, int D, int k) {
    // sized for the worst plan: K steps of the euclidean plan + one per robot (<= 8 accumulators), the narrowest tile
    const int ks = (D + 2 + 7) / 8 + 8;
    const size_t lists = (size_t)mrb::knn_tc_max_lists(Q);
    return 256 + align256((size_t)((Q + 127) / 128) * ks * 128 * 32) + align256((size_t)(N + 256) * ks * 32) +
           align256(lists * ((size_t)(k + 16) * 8 + 4)) + align256((size_t)Q * 4 + 16) + align256((size_t)1024 * 32 * (size_t)k * 12) + align256((size_t)Q) + 4096;
}

size_t mrb200_knn_workspace_bytes(int64_t Q, int64_t N, int D, int k) {
    if (Q <= 0 || k <= 0) return 256;
    const int splits = mrb::knn_pick_splits(Q, N);
    const size_t exact = align256((size_t)splits * (size_t)Q * (size_t)k * 12 + 256);
    return exact + (k <= mrb::knn_tc_max_k() ? tc_workspace_bytes(Q, N, D, k) : 0);
}

size_t mrb200_knn_stats_offset(int64_t Q, int64_t N, int D, int k) {
    if (Q <= 0 || k <= 0) return 0;
    const int splits = mrb::knn_pick_splits(Q, N);
    return align256((size_t)splits * (size_t)Q * (size_t)k * 12 + 256) + align256((size_t)Q);
}

int mrb200_knn(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const int32_t* slices_host, int R, int metric,
               int k, int32_t* out_idx, double* out_dist, void* workspace, size_t workspace_bytes, int mode, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, metric, &sl)) return rc;
    if (Q < 0 || N < 0 || k < 1 || k > 128 || (Q && (!queries || !out_idx)) || (N && !corpus))
        return fail(MRB200_ERR_ARG, "knn: bad argument (k must be in [1, 128])");
    if (Q == 0) return MRB200_OK;
    if (workspace_bytes < mrb200_knn_workspace_bytes(Q, N, D, k) || !workspace) return fail(MRB200_ERR_ARG, "knn: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = mrb::knn_pick_splits(Q, N);
    double* part_d = (double*)workspace;
    int* part_i = (int*)(part_d + (size_t)splits * Q * k);
    unsigned char* tc_ws = (unsigned char*)workspace + align256((size_t)splits * (size_t)Q * (size_t)k * 12 + 256);
    mrb::TcPlan plan;
    const bool tc_ok = tc_usable(Q, N, D, sl, metric, k, &plan);
    if (mode == 2 && !tc_ok)
        return fail(MRB200_ERR_ARG, "knn: the tensor-core path needs metric euclidean / max_euclidean, k <= %d, N >= 1024", mrb::knn_tc_max_k());
    const bool use_tc = mode == 2 || (mode == 0 && tc_ok && Q >= 256 && N >= 4096);
    const uint8_t* skip = nullptr;
    if (use_tc) {
        const int64_t ct = (N + plan.tn - 1) / plan.tn;
        const int kc = mrb::knn_tc_slots(k, mrb::knn_tc_min_lists(Q, ct));   // output slots per row and list
        uint8_t* certified = tc_ws;
        cudaError_t e = mrb::launch_knn_tc(queries, corpus, Q, N, D, sl, metric, k, kc, plan, tc_ws + align256((size_t)Q), out_idx,
                                           out_dist, certified, st);
        if (e != cudaSuccess) return cuda_fail(e, "knn (tensor-core path)");
        g_launches += 6;   // 2 x operand preparation, candidate generator, re-rank, exact rows without a certificate + their merge
        (void)skip;
        return MRB200_OK;
    }
    cudaError_t e = mrb::launch_knn_exact(queries, corpus, Q, N, D, sl, metric, k, splits, part_d, part_i, out_idx, out_dist, nullptr, st);
    if (e != cudaSuccess) return cuda_fail(e, "knn");
    g_launches += 2;
    return MRB200_OK;
}

// ---- r-disc search on the tensor-core candidate generator ----
static bool radius_tc_plan(int64_t Q, int64_t N, int D, const mrb::Slices& sl, int metric, mrb::TcPlan* plan) {
    if (N < 1024 || Q < 1) return false;
    if (!mrb::knn_tc_make_plan(D, sl, metric, plan)) return false;
    return mrb::knn_tc_smem_bytes(*plan, 0) <= 224 * 1024;
}

size_t mrb200_radius_tc_workspace_bytes(int64_t Q, int64_t N, int D, int cap) {
    if (Q <= 0 || cap <= 0) return 256;
    const int ks = (D + 2 + 7) / 8 + 8;   // worst plan, as for mrb200_knn_workspace_bytes
    return 256 + align256((size_t)((Q + 127) / 128) * ks * 128 * 32) + align256((size_t)(N + 256) * ks * 32) + align256((size_t)Q * 4 + 16) +
           align256((size_t)Q * (size_t)cap * 4) + 4096;
}

int mrb200_radius_tc_count(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const int32_t* slices_host, int R,
                           int metric, const double* radii, double radius, int inclusive, int cap, void* workspace, size_t workspace_bytes,
                           int64_t* counts, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, metric, &sl)) return rc;
    if (Q < 0 || N < 0 || cap < 1 || (Q && (!queries || !counts)) || (N && !corpus)) return fail(MRB200_ERR_ARG, "radius_tc_count: bad argument");
    if (Q == 0) return MRB200_OK;
    mrb::TcPlan plan;
    if (!radius_tc_plan(Q, N, D, sl, metric, &plan))
        return fail(MRB200_ERR_ARG, "radius_tc_count: the tensor-core path needs metric euclidean / max_euclidean and N >= 1024");
    if (!workspace || workspace_bytes < mrb200_radius_tc_workspace_bytes(Q, N, D, cap)) return fail(MRB200_ERR_ARG, "radius_tc_count: workspace too small");
    cudaError_t e = mrb::launch_radius_tc_count(queries, corpus, Q, N, D, sl, metric, radii, radius, inclusive, cap, plan, workspace, counts,
                                                (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "radius_tc_count");
    g_launches += 4;
    return MRB200_OK;
}

int mrb200_radius_tc_fill(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const int32_t* slices_host, int R,
                          int metric, int cap, void* workspace, size_t workspace_bytes, const int64_t* offsets, int32_t* out_idx,
                          double* out_dist, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, metric, &sl)) return rc;
    if (Q < 0 || (Q && (!queries || !offsets || !out_idx || !workspace))) return fail(MRB200_ERR_ARG, "radius_tc_fill: bad argument");
    if (Q == 0) return MRB200_OK;
    mrb::TcPlan plan;
    if (!radius_tc_plan(Q, N, D, sl, metric, &plan)) return fail(MRB200_ERR_ARG, "radius_tc_fill: unsupported metric / size");
    if (workspace_bytes < mrb200_radius_tc_workspace_bytes(Q, N, D, cap)) return fail(MRB200_ERR_ARG, "radius_tc_fill: workspace too small");
    cudaError_t e = mrb::launch_radius_tc_fill(queries, corpus, Q, N, D, sl, metric, cap, plan, workspace, offsets, out_idx, out_dist,
                                               (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "radius_tc_fill");
    g_launches++;
    return MRB200_OK;
}

int mrb200_radius_splits(int64_t Q, int64_t N) { return mrb::knn_pick_splits(Q, N); }

static int radius_impl(bool fill, const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const int32_t* slices_host,
                       int R, int metric, const double* radii, double radius, int inclusive, int splits, int64_t* counts,
                       const int64_t* offsets, int32_t* out_idx, double* out_dist, mrb200_stream_t stream) {
    mrb::Slices sl;
    if (int rc = make_slices(slices_host, R, D, metric, &sl)) return rc;
    if (Q < 0 || N < 0 || splits < 1 || (Q && !queries) || (N && !corpus) || (Q && !fill && !counts) || (Q && fill && !offsets))
        return fail(MRB200_ERR_ARG, "radius: bad argument");
    if (Q == 0) return MRB200_OK;
    cudaError_t e = mrb::launch_radius(fill, queries, corpus, Q, N, D, sl, metric, radii, radius, inclusive, splits, counts, offsets,
                                       out_idx, out_dist, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "radius");
    g_launches++;
    return MRB200_OK;
}

int mrb200_radius_count(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const int32_t* slices_host, int R,
                        int metric, const double* radii, double radius, int inclusive, int splits, int64_t* counts,
                        mrb200_stream_t stream) {
    return radius_impl(false, queries, corpus, Q, N, D, slices_host, R, metric, radii, radius, inclusive, splits, counts, nullptr,
                       nullptr, nullptr, stream);
}

int mrb200_radius_fill(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const int32_t* slices_host, int R,
                       int metric, const double* radii, double radius, int inclusive, int splits, const int64_t* offsets,
                       int32_t* out_idx, double* out_dist, mrb200_stream_t stream) {
    if (Q > 0 && !out_idx) return fail(MRB200_ERR_ARG, "radius_fill: out_idx is null");
    return radius_impl(true, queries, corpus, Q, N, D, slices_host, R, metric, radii, radius, inclusive, splits, nullptr, offsets,
                       out_idx, out_dist, stream);
}

}  // extern "C"
