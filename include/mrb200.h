/* mrb200 -- C ABI of the B200-native collision / proximity backend.
 *
 * This is the drop-in boundary for the hot path of vhartman/multirobot-pathplanning-benchmark
 * (per-configuration and per-edge collision queries, their batch variants, and the
 * distance / nearest-neighbour calls of the planners).  Plain pointers and sizes only; every
 * data pointer marked `dev` is a CUDA device pointer owned by the caller, `stream` is a
 * cudaStream_t passed as void*.  All functions return 0 on success and a negative code on
 * error (mrb200_last_error() gives the text); nothing is ever reported "free" on failure.
 *
 * Reference interfaces replaced (paths relative to the reference repository,
 * P/ = src/multi_robot_multi_goal_planning/):
 *   BaseProblem.is_collision_free            P/problems/planning_env.py:1724-1734
 *   BaseProblem.is_collision_free_for_robot  P/problems/planning_env.py:1736-1744
 *   BaseProblem.is_edge_collision_free       P/problems/planning_env.py:1746-1763
 *   batch_config_dist / batch_config_cost    P/problems/core/configuration.py:342-349, 437-510
 *   MultimodalGraph.get_neighbors            P/planners/prm/prm_graph.py:389-549
 */
#ifndef MRB200_H
#define MRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRB200_OK 0
#define MRB200_ERR_ARG (-1)         /* bad argument / unsupported shape */
#define MRB200_ERR_BLOB (-2)        /* scene blob magic / version / size mismatch */
#define MRB200_ERR_CUDA (-3)        /* CUDA runtime error (see mrb200_last_error) */
#define MRB200_ERR_NO_DEVICE (-4)   /* no CUDA device: there is no CPU fallback */

typedef struct mrb200_scene mrb200_scene_t;       /* primitive scene with per-mode slots */
typedef struct mrb200_abstract mrb200_abstract_t; /* sphere-agent environment */
typedef void* mrb200_stream_t;                    /* cudaStream_t */

int mrb200_version(void);
const char* mrb200_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t mrb200_launch_count(void);

/* FP32 FMA roofline probe (measurement aid for bench.py): runs n_threads threads x iters x 16
 * FMAs (16 independent chains per thread, iters rounded up to a multiple of 16); out_dev holds n_threads floats (pass NULL to only query n_threads). */
int mrb200_fp32_probe(int iters, float* out_dev, int32_t* n_threads, mrb200_stream_t stream);

/* ---- sphere-agent environment: AbstractEnvironment, P/problems/abstract_env.py:112-354 ----
 * radii[n_agents]; spheres: n_sph x (dim centre + radius); rects: n_rect x (dim min + dim max).
 * fp64 throughout, flags bit-identical to the reference on the same inputs. */
int mrb200_abstract_create(int n_agents, int dim, const double* radii, int n_sph, const double* spheres,
                           int n_rect, const double* rects, mrb200_abstract_t** out);
int mrb200_abstract_destroy(mrb200_abstract_t* env);
/* is_collision_free for a batch (abstract_env.py:255-276): free[i] = 1 iff q[i] is collision free */
int mrb200_abstract_check_configs(const mrb200_abstract_t* env, const double* q_dev /*[B, n_agents*dim]*/,
                                  int64_t B, uint8_t* free_dev /*[B]*/, mrb200_stream_t stream);
/* is_edge_collision_free for a batch (abstract_env.py:301-354).  N_dev nullable (then
 * N = max(2, int(|q2-q1|_inf / resolution) + 1)); n_max < 0 means N.  first_pos_dev nullable:
 * position in the reference's binary order of the first colliding sample, -1 if none. */
int mrb200_abstract_check_edges(const mrb200_abstract_t* env, const double* q1_dev, const double* q2_dev,
                                int64_t E, double resolution, const int32_t* N_dev, int32_t n_start,
                                int32_t n_max, int include_endpoints, uint8_t* free_dev,
                                int32_t* first_pos_dev, mrb200_stream_t stream);

/* the same two queries with HOST buffers (one call = staging + launch + synchronisation; mapped pinned staging owned
 * by the handle, calls serialised per handle): the planners' one-query-at-a-time seam, see mrb200_query_configs_host */
int mrb200_abstract_query_configs_host(mrb200_abstract_t* env, const double* q_host, int64_t B, uint8_t* free_host,
                                       mrb200_stream_t stream);
int mrb200_abstract_query_edges_host(mrb200_abstract_t* env, const double* q1_host, const double* q2_host, int64_t E,
                                     double resolution, const int32_t* N_host, int32_t n_start, int32_t n_max,
                                     int include_endpoints, uint8_t* free_host, int32_t* first_pos_host,
                                     mrb200_stream_t stream);

/* ---- primitive scenes (rai-style): rai_env, P/problems/rai_base_env.py:442-836 ----
 * A scene holds `max_modes` slots; each slot is a compiled scene blob (layout:
 * multirobot_pathplanning_benchmark_b200/csrc/scene_blob.h) for one mode's kinematic tree
 * (rai_base_env.py:704-836 set_to_mode). */
int mrb200_scene_create(int max_modes, mrb200_scene_t** out);
int mrb200_scene_destroy(mrb200_scene_t* scene);
/* upload a host blob into a slot and evaluate its static-static pairs on the device */
int mrb200_scene_set_mode(mrb200_scene_t* scene, int slot, const void* blob_host, size_t nbytes,
                          mrb200_stream_t stream);
/* is_collision_free / is_collision_free_np for a batch (rai_base_env.py:442-513):
 * free[i] = 1 iff total penetration <= tol.  tol < 0 uses the blob's.  pen_dev nullable: total
 * penetration per configuration (complete only when full_eval != 0, otherwise the kernel may
 * stop early once a configuration is decided). */
int mrb200_check_configs(const mrb200_scene_t* scene, int slot, const float* q_dev /*[B, D]*/, int64_t B,
                         float tol, uint8_t* free_dev, float* pen_dev, int full_eval,
                         mrb200_stream_t stream);
/* is_collision_free_for_robot for a batch (rai_base_env.py:515-615): free[i] = 0 iff total
 * penetration > tol and some penetrating pair involves a `relevant` shape and no `other`
 * shape.  relevant / other: host arrays of n_shapes bytes (0/1). */
int mrb200_check_configs_for_robot(const mrb200_scene_t* scene, int slot, const float* q_dev, int64_t B,
                                   float tol, const uint8_t* relevant_host, const uint8_t* other_host,
                                   int n_shapes, uint8_t* free_dev, mrb200_stream_t stream);
/* is_edge_collision_free for a batch (rai_base_env.py:618-676); arguments as for the abstract
 * variant, endpoints are fp32 and interpolation runs in fp64 like the reference. */
int mrb200_check_edges(const mrb200_scene_t* scene, int slot, const float* q1_dev, const float* q2_dev,
                       int64_t E, double resolution, const int32_t* N_dev, int32_t n_start, int32_t n_max,
                       int include_endpoints, float tol, uint8_t* free_dev, int32_t* first_pos_dev,
                       mrb200_stream_t stream);
/* The planners' one-query-at-a-time seam with HOST buffers (is_collision_free / is_collision_free_for_robot /
 * is_edge_collision_free called per configuration or edge, e.g. P/planners/collision_free_sampler.py:106-143,
 * P/planners/rrtstar_base.py:626-629): copy in, launch, copy out and synchronise inside one call, through
 * staging buffers owned by the scene handle (grown on demand; calls on one handle are serialised).  Host
 * pointers may be pageable.  relevant_host / other_host: null for the plain rule, else as in
 * mrb200_check_configs_for_robot.  N_host: null = from the resolution.  first_pos_host nullable. */
int mrb200_query_configs_host(mrb200_scene_t* scene, int slot, const float* q_host /*[B, D]*/, int64_t B, float tol,
                              const uint8_t* relevant_host, const uint8_t* other_host, int n_shapes,
                              uint8_t* free_host, mrb200_stream_t stream);
int mrb200_query_edges_host(mrb200_scene_t* scene, int slot, const float* q1_host, const float* q2_host, int64_t E,
                            double resolution, const int32_t* N_host, int32_t n_start, int32_t n_max,
                            int include_endpoints, float tol, uint8_t* free_host, int32_t* first_pos_host,
                            mrb200_stream_t stream);
/* Whole sample batches with HOST buffers (BaseProblem.is_collision_free over an array of samples, e.g. the validation
 * of a PRM sample batch, P/planners/prm/prm_graph.py / collision_free_sampler.py:99-143): the batch is cut into chunks of
 * `chunk` configurations (<= 0: 262144) that overlap H2D copy, kernel and D2H read-back on side streams owned by the
 * handle.  Pinned (or cudaHostRegister-ed) buffers are copied from / to directly; pageable ones are bounced through
 * pinned staging.  Returns after every flag has landed in free_host; later work on `stream` is ordered behind it. */
int mrb200_check_configs_host(mrb200_scene_t* scene, int slot, const float* q_host /*[B, D]*/, int64_t B, float tol,
                              uint8_t* free_host /*[B]*/, int64_t chunk, mrb200_stream_t stream);
/* Asynchronous edge batches with HOST buffers -- the seam behind env.py's speculation of a PRM / EIT* node's candidate
 * edges (P/planners/prm/prm_graph.py:675 checks them lazily, one call each, right after get_neighbors :389-549 has
 * named them all).  submit: inputs are copied into a pinned buffer owned by the handle, H2D + kernel + D2H are queued on
 * `stream`, an event is recorded and the call returns with a ticket.  q1_host holds q1_rows = 1 (one start for all edges)
 * or E rows.  Whole edges only (no window).  collect: waits for the ticket's event, writes free_host[E] and
 * first_pos_host[E] (nullable).  A ticket expires after 64 further submits on the handle
 * (collect then fails with MRB200_ERR_ARG and the caller re-submits). */
int mrb200_submit_edges_host(mrb200_scene_t* scene, int slot, const float* q1_host, int q1_rows, const float* q2_host,
                             int64_t E, double resolution, const int32_t* N_host, int include_endpoints, float tol,
                             int64_t* ticket_out, mrb200_stream_t stream);
int mrb200_collect_edges_host(mrb200_scene_t* scene, int64_t ticket, int64_t E, uint8_t* free_host, int32_t* first_pos_host);
/* introspection of a slot: D, n_shapes, n_pairs (dynamic), shared memory bytes per CTA */
int mrb200_scene_info(const mrb200_scene_t* scene, int slot, int32_t* out4);
/* Large plain batches (mrb200_check_configs, B >= 4096, no penetration output) can run as two-phase tiles: a cheap
 * lower bound of the penetration against the large static boxes (table, floor) retires most colliding
 * configurations after FK; the undecided ones are pooled and evaluated exactly as by the single-pass kernel, so
 * flags are the same either way.  It pays when the bound decides more than about half of the batch, which depends
 * on the caller's inputs: by default the first large batches on a slot measure it and the slot settles on one
 * kernel.  policy: 0 = measure and settle (default, also resets the measurement), 1 = always two-phase,
 * 2 = always single pass.  mrb200_scene_get_two_phase: out3 = {policy / settled state (0 measuring, 1 two-phase,
 * 2 single pass), configurations seen while measuring, of those decided by the bound}. */
int mrb200_scene_set_two_phase(mrb200_scene_t* scene, int slot, int policy);
int mrb200_scene_get_two_phase(const mrb200_scene_t* scene, int slot, int32_t* out3);

/* ---- distances and neighbour search: batch_config_dist (P/problems/core/configuration.py:303-349),
 * PRM get_neighbors (P/planners/prm/prm_graph.py:389-549), RRT* near (P/planners/rrtstar_base.py:
 * 439-453, 1296-1337), IT* get_neighbors (P/planners/itstar_base.py:1388-1527) ----
 * All coordinates are fp64 like the reference's arrays.  slices_host: R x [start, end) per robot.
 * metric: 0 euclidean, 1 sum_euclidean, 2 max_euclidean, 3 max (infinity norm). */
#define MRB200_METRIC_EUCLIDEAN 0
#define MRB200_METRIC_SUM_EUCLIDEAN 1
#define MRB200_METRIC_MAX_EUCLIDEAN 2
#define MRB200_METRIC_MAX 3
/* one-to-many distance: out[n] = dist(q, pts[n]) */
int mrb200_batch_dist(const double* q_dev /*[D]*/, const double* pts_dev /*[N, D]*/, int64_t N, int D,
                      const int32_t* slices_host, int R, int metric, double* out_dev /*[N]*/,
                      mrb200_stream_t stream);
/* batch_config_cost (P/problems/core/configuration.py:437-510): out[n] = cost(a[n] or the single row a, b[n]);
 * per-robot metric euclidean (per_robot_max = 0) or max-abs (1); reduction max + w * sum (reduction_sum = 0,
 * the reference's "max" with w = 0.01) or sum (1). */
int mrb200_batch_cost(const double* a_dev, int a_is_single, const double* b_dev /*[N, D]*/, int64_t N, int D,
                      const int32_t* slices_host, int R, int per_robot_max, int reduction_sum, double w,
                      double* out_dev /*[N]*/, mrb200_stream_t stream);
/* One relaxation step of the lower bound on the cost to the goal (compute_lower_bound_to_goal,
 * P/planners/prm/prm_graph.py:143-220, relaxes transition nodes one at a time with batch_config_cost): for the exit
 * configurations a[T1, D] of one mode and those of the modes it leads into, b[T2, D] with their bounds lb_b[T2],
 *     out[i] = min_j ( cost(a_i, b_j) + lb_b[j] ),   arg[i] = the minimising j (smallest on ties; nullable; -1 if T2 = 0)
 * with the cost of mrb200_batch_cost (per-robot euclidean or max-abs, reduced by sum or by max + w * sum). */
int mrb200_minplus_cost(const double* a_dev, int64_t T1, const double* b_dev, const double* lb_b_dev, int64_t T2, int D,
                        const int32_t* slices_host, int R, int per_robot_max, int reduction_sum, double w, double* out_dev,
                        int32_t* arg_dev, mrb200_stream_t stream);
/* k nearest neighbours of every query row, ascending (distance, index); rows with fewer than k
 * corpus points are padded with index -1 / distance +inf.  out_dist_dev nullable.  The workspace
 * (device, mrb200_knn_workspace_bytes) is scratch for partial results.
 * mode: 0 = automatic, 1 = exact fp64 CUDA-core path, 2 = tensor-core candidates + fp64 re-rank
 * (euclidean / max_euclidean only); both return the same indices. */
size_t mrb200_knn_workspace_bytes(int64_t Q, int64_t N, int D, int k);
/* byte offset, inside the workspace of the last tensor-core mrb200_knn call with these sizes, of two uint32 statistics:
 * [0] bit pattern of the largest squared slice norm (the error bound's scale), [1] rows the re-rank could not certify
 * (recomputed by the exact kernel inside the same call).  Read after synchronising the stream. */
size_t mrb200_knn_stats_offset(int64_t Q, int64_t N, int D, int k);
int mrb200_knn(const double* queries_dev /*[Q, D]*/, const double* corpus_dev /*[N, D]*/, int64_t Q, int64_t N,
               int D, const int32_t* slices_host, int R, int metric, int k, int32_t* out_idx_dev /*[Q, k]*/,
               double* out_dist_dev /*[Q, k]*/, void* workspace_dev, size_t workspace_bytes, int mode,
               mrb200_stream_t stream);
/* r-disc search on the tensor-core candidate generator (euclidean / max_euclidean, N >= 1024): coarse TF32 distances
 * select, per query row, every point that can lie within the radius (at most `cap` of them); an exact fp64 pass in the
 * reference's operand order keeps those that do (d < r, or d <= r + 1e-10 when inclusive) in ascending index order.
 * count: counts_dev[Q] = neighbours of the row, or -1 if its candidates exceeded `cap` (answer that row with
 * mrb200_radius_count / _fill).  The caller scans the counts into offsets (rows with -1 take their exact count), then
 * fill writes out_idx / out_dist (nullable) for the rows that did not overflow.  Same workspace for both calls. */
size_t mrb200_radius_tc_workspace_bytes(int64_t Q, int64_t N, int D, int cap);
int mrb200_radius_tc_count(const double* queries_dev, const double* corpus_dev, int64_t Q, int64_t N, int D,
                           const int32_t* slices_host, int R, int metric, const double* radii_dev, double radius, int inclusive,
                           int cap, void* workspace_dev, size_t workspace_bytes, int64_t* counts_dev, mrb200_stream_t stream);
int mrb200_radius_tc_fill(const double* queries_dev, const double* corpus_dev, int64_t Q, int64_t N, int D,
                          const int32_t* slices_host, int R, int metric, int cap, void* workspace_dev, size_t workspace_bytes,
                          const int64_t* offsets_dev, int32_t* out_idx_dev, double* out_dist_dev, mrb200_stream_t stream);
/* radius search in two launches around a caller-side exclusive scan.  splits =
 * mrb200_radius_splits(Q, N) corpus ranges are searched independently; counts_dev is [Q * splits]
 * (row major).  radii_dev (per query) nullable, then `radius` applies to all.  inclusive = 0:
 * d < r (PRM); 1: d <= r + 1e-10 (RRT* / IT*).  fill writes the neighbour indices of row i, in
 * ascending index order, starting at offsets_dev[i * splits] (offsets = exclusive scan of counts). */
int mrb200_radius_splits(int64_t Q, int64_t N);
int mrb200_radius_count(const double* queries_dev, const double* corpus_dev, int64_t Q, int64_t N, int D,
                        const int32_t* slices_host, int R, int metric, const double* radii_dev, double radius,
                        int inclusive, int splits, int64_t* counts_dev, mrb200_stream_t stream);
int mrb200_radius_fill(const double* queries_dev, const double* corpus_dev, int64_t Q, int64_t N, int D,
                       const int32_t* slices_host, int R, int metric, const double* radii_dev, double radius,
                       int inclusive, int splits, const int64_t* offsets_dev, int32_t* out_idx_dev,
                       double* out_dist_dev, mrb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
