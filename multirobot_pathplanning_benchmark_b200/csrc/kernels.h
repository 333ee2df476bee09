// Internal launch interface between the C-ABI layer (capi.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mrb {

// A6 (is_collision_free_for_robot) shape classification, one bit per shape (<= 256 shapes)
struct RobotRule {
    unsigned long long rel[4];  // shape belongs to a queried robot or is a frame of its active task
    unsigned long long oth[4];  // shape belongs to another robot
    int enabled;
};

struct ConfigParams {
    const uint32_t* blob;  // device, compiled scene for the mode (all of it)
    int blob_words;        // MRB_H_STAGED_WORDS: the prefix the kernel stages in shared memory
    int D, world_words, n_shapes;
    const float* q;  // device [B, D]
    int64_t B;
    float tol;        // < 0: use the blob's
    uint8_t* flags;   // device [B], 1 = collision free
    float* pen_out;   // device [B] or null
    int full_eval;    // 1: no early exit (pen_out is then the complete sum)
    int bulk_ok;      // q is 16-byte aligned: tiles may be fetched with cp.async.bulk
    int two_phase;    // 1: two-phase tiles (pairs against the table / floor first, pooled survivors for the rest)
    int* stats;       // device [2] or null: two-phase kernels add (configurations seen, decided in phase A)
    RobotRule rule;
};

struct EdgeParams {
    const uint32_t* blob;
    int blob_words, D, world_words, n_shapes;
    const float* q1;  // device [E, D]
    const float* q2;  // device [E, D]
    int64_t E;
    double resolution;
    const int32_t* N;  // device [E] or null
    int n_start, n_max, include_endpoints;
    float tol;
    uint8_t* flags;       // device [E], 1 = edge collision free
    int32_t* first_pos;   // device [E] or null: position (in binary order) of the first colliding sample
    int* counter;         // device scratch: dynamic edge scheduler
};

size_t scene_smem_bytes(int blob_words, int D, int world_words, int n_shapes, int kind = 1);  // 0 configs, 1 edges, 2 two-phase configs
cudaError_t launch_static_penetration(uint32_t* blob, cudaStream_t st);
cudaError_t launch_check_configs(const ConfigParams& p, cudaStream_t st);
cudaError_t launch_check_edges(const EdgeParams& p, cudaStream_t st);

// ---- abstract sphere-agent environment (fp64, bit-exact with the reference) ----
constexpr int ABS_MAX_AGENTS = 8;
constexpr int ABS_MAX_DIM = 12;
constexpr int ABS_MAX_OBS = 6;
struct AbstractSceneData {
    int n_agents, dim, n_sph, n_rect;
    double radii[ABS_MAX_AGENTS];
    double sph_c[ABS_MAX_OBS][ABS_MAX_DIM];
    double sph_r[ABS_MAX_OBS];
    double rect_min[ABS_MAX_OBS][ABS_MAX_DIM];
    double rect_max[ABS_MAX_OBS][ABS_MAX_DIM];
};
cudaError_t launch_abstract_configs(const AbstractSceneData& sc, const double* q, int64_t B, uint8_t* flags, cudaStream_t st);
cudaError_t launch_abstract_edges(const AbstractSceneData& sc, const double* q1, const double* q2, int64_t E,
                                  double resolution, const int32_t* N, int n_start, int n_max, int include_endpoints,
                                  uint8_t* flags, int32_t* first_pos, int* counter, cudaStream_t st);

// ---- distances and neighbour search (knn_kernels.cu) ----
struct Slices;
int knn_pick_splits(int64_t Q, int64_t N);
cudaError_t launch_batch_dist(const double* q, const double* pts, int64_t N, int D, const Slices& sl, int metric, double* out,
                              cudaStream_t st);
cudaError_t launch_knn_exact(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                             int k, int splits, double* part_d, int* part_i, int32_t* out_idx, double* out_dist, const uint8_t* skip,
                             cudaStream_t st);
cudaError_t launch_batch_cost(const double* a, int64_t a_stride, const double* b, int64_t N, int D, const Slices& sl, int per_robot_max,
                              int reduction_sum, double w, double* out, cudaStream_t st);
cudaError_t launch_minplus_cost(const double* a, int64_t T1, const double* b, const double* lb_b, int64_t T2, int D, const Slices& sl,
                                int per_robot_max, int reduction_sum, double w, double* out, int32_t* arg, cudaStream_t st);
// tensor-core candidate generator + exact re-rank (knn_tc_kernels.cu)
struct TcPlan;
bool knn_tc_make_plan(int D, const Slices& sl, int metric, TcPlan* plan);
size_t knn_tc_smem_bytes(const TcPlan& plan, int kc);
void knn_tc_shape(int64_t Q, int64_t n_ctiles, int64_t* full_qtiles, int* tail_splits);   // work decomposition of a launch
int64_t knn_tc_max_lists(int64_t Q);                    // candidate lists a launch writes, at most
int knn_tc_min_lists(int64_t Q, int64_t n_ctiles);      // lists per row, at least
int knn_tc_keep(int k, int n_lists);    // candidates a list keeps at least
int knn_tc_slots(int k, int n_lists);   // output slots per row and list
int knn_tc_max_k();
int knn_tc_parts();                    // candidate lists per query row and corpus split
size_t knn_tc_workspace_bytes(int64_t Q, int64_t N, const TcPlan& plan, int kc);
cudaError_t launch_knn_tc(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                          int k, int kc, const TcPlan& plan, void* workspace, int32_t* out_idx, double* out_dist,
                          uint8_t* certified, cudaStream_t st);
size_t radius_tc_workspace_bytes(int64_t Q, int64_t N, const TcPlan& plan, int cap);
cudaError_t launch_radius_tc_count(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                                   const double* radii, double radius, int inclusive, int cap, const TcPlan& plan, void* workspace,
                                   int64_t* counts, cudaStream_t st);
cudaError_t launch_radius_tc_fill(const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl, int metric,
                                  int cap, const TcPlan& plan, void* workspace, const int64_t* offsets, int32_t* out_idx, double* out_dist,
                                  cudaStream_t st);
cudaError_t launch_radius(bool fill, const double* queries, const double* corpus, int64_t Q, int64_t N, int D, const Slices& sl,
                          int metric, const double* radii, double radius, int inclusive, int splits, int64_t* counts,
                          const int64_t* offsets, int32_t* out_idx, double* out_dist, cudaStream_t st);

// out: device [n_threads] floats; call with out == nullptr to query n_threads
cudaError_t launch_fp32_probe(int iters, float* out, int* n_threads, cudaStream_t st);

}  // namespace mrb
