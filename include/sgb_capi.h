/*
 * sgb_capi.h -- C ABI of the B200-native pose-graph optimisation backend for sparse-gslam.
 *
 * This is the drop-in boundary of SURVEY.md section 8b: what a g2o plugin
 * (OptimizationAlgorithm / BlockSolver / LinearSolver) that replaces the three
 * nested objects built in the reference's
 *     src/sparse_gslam/src/graphs.cpp:9-15  setup_lm_opt   (LM,  BlockSolver<-1,2>, LinearSolverEigen)
 *     src/sparse_gslam/src/graphs.cpp:17-23 setup_pose_opt (GN,  BlockSolver<3,3>,  LinearSolverEigen)
 * binds to. Plain pointers and sizes only; no C++/torch types; no exceptions cross it.
 * The g2o-facing C++ shim that calls these entry points is
 * sparse-gslam_b200/adapter/sgb_g2o_adapter.h; INTEGRATION.md shows the two lines of
 * graphs.cpp that change.
 *
 * Every function returns an sgb_status (0 = OK) unless stated otherwise;
 * sgb_last_error() gives the message of the last failure on that handle.
 * A handle is NOT re-entrant, but distinct handles may be used concurrently from
 * different threads (reference: the landmark graph and the pose graph are optimised from
 * two threads in realtime mode, log_runner.cpp:217-224); each handle owns its CUDA
 * stream and never touches the default stream.
 * There is no CPU fallback: every compute entry point fails with SGB_ERR_NO_DEVICE
 * when no CUDA device is usable.
 */
#ifndef SGB_CAPI_H
#define SGB_CAPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sgb_status {
  SGB_OK = 0,
  SGB_ERR_INVALID = 1,      /* bad argument / malformed graph */
  SGB_ERR_NO_DEVICE = 2,    /* no usable CUDA device (there is no CPU path) */
  SGB_ERR_CUDA = 3,         /* CUDA runtime error, see sgb_last_error */
  SGB_ERR_NOT_INITIALIZED = 4, /* g2o: "0 vertices to optimize, maybe forgot to call initializeOptimization()" */
  SGB_ERR_UNSUPPORTED = 5,
  SGB_ERR_SOLVE_FAILED = 6, /* g2o SolverResult::Fail (non-SPD system / PCG breakdown) */
  SGB_ERR_COMM = 7          /* multi-GPU exchange failure */
} sgb_status;

/* g2o::OptimizationAlgorithm::SolverResult */
enum { SGB_RESULT_TERMINATE = 2, SGB_RESULT_OK = 1, SGB_RESULT_FAIL = -1 };
/* which g2o algorithm object is being replaced (graphs.cpp:12 vs graphs.cpp:20) */
enum { SGB_ALGO_LM = 0, SGB_ALGO_GN = 1 };
/* Jacobian of EdgeSE2RhoTheta: the reference inherits g2o's central differences
 * (include/g2o_bindings/edge_se2_rhotheta.h:8-17 has no linearizeOplus); analytic = SURVEY Appendix B */
enum { SGB_JAC_G2O_NUMERIC = 0, SGB_JAC_ANALYTIC = 1 };

typedef struct sgb_options {
  int32_t device;           /* CUDA device ordinal; -1 = current */
  int32_t jacobian_mode;    /* SGB_JAC_* (default SGB_JAC_G2O_NUMERIC = reference behaviour) */
  double pcg_tolerance;     /* stop when sqrt(r.z) <= tol * sqrt(r0.z0); <=0 -> 1e-10 */
  int32_t pcg_max_iters;    /* <=0 -> 4 * (3 * free poses) */
  int32_t verbose;          /* g2o setVerbose; the reference keeps it off (graphs.cpp:13,21) */
  double lm_tau;            /* lambda0 = tau * max|diag H|; <=0 -> 1e-5 (g2o default) */
  double lm_user_lambda;    /* >0 overrides lambda0 (g2o userLambdaInit) */
  int32_t lm_max_trials;    /* <=0 -> 10 (g2o maxTrialsAfterFailure) */
  int32_t incremental;      /* != 0: the handle keeps what sgb_update_graph needs (an index mirror of the graph on the host
                             * and the raw edge values on the device); 0 (default): sgb_set_graph keeps nothing of the caller's */
  int32_t coarse_nodes;     /* two-level preconditioner of the resident solve (small graphs on one GPU; no reference
                             * counterpart, LinearSolverEigen factorises exactly): > 0 = at most this many coarse nodes
                             * (<= 40), < 0 = off, 0 = the library default: on with 40 nodes unless the environment says
                             * SGB_COARSE=0 (off) or SGB_COARSE_NODES=n */
} sgb_options;

/* Host SoA graph. Edges reference vertices by ARRAY INDEX; ids only define g2o's vertex
 * order (active vertices sorted by id; the reference numbers poses 0,1,2.. drone.cpp:64,121 and
 * landmarks from 10 000 000, include/drone.h:22). All pose ids must be smaller than all
 * landmark ids. The caller owns every pointer and may free them when the call returns
 * (reference ownership convention: graphs.h:22-24, README.md:22-23). */
typedef struct sgb_graph_soa {
  int32_t n_poses;
  const int32_t* pose_id;    /* NULL: id = index */
  const double* pose_est;    /* [3*n_poses] x, y, theta (VertexSE2::estimate) */
  const uint8_t* pose_fixed; /* NULL: none */
  int32_t n_landmarks;
  const int32_t* lm_id;      /* NULL: id = 10000000 + index */
  const double* lm_est;      /* [2*n_landmarks] rho, theta (VertexRhoTheta::estimate) */
  const uint8_t* lm_fixed;   /* NULL: none */
  int32_t n_pp;              /* g2o::EdgeSE2 */
  const int32_t* pp_i;       /* vertex 0 (from) */
  const int32_t* pp_j;       /* vertex 1 (to) */
  const double* pp_z;        /* [3*n_pp] measurement dx, dy, dtheta */
  const double* pp_info;     /* [6*n_pp] information, upper triangle 11,12,13,22,23,33 */
  const double* pp_phi;      /* RobustKernelDCS delta per edge; <=0 or NULL = no robust kernel */
  const int64_t* pp_seq;     /* insertion rank (g2o internalId); NULL = array order */
  int32_t n_pl;              /* g2o::EdgeSE2RhoTheta */
  const int32_t* pl_pose;    /* vertex 0 */
  const int32_t* pl_lm;      /* vertex 1 */
  const double* pl_z;        /* [2*n_pl] rho, theta */
  const double* pl_info;     /* [3*n_pl] 11,12,22 */
  const int64_t* pl_seq;     /* NULL = after all pose-pose edges, array order */
} sgb_graph_soa;

typedef struct sgb_iter_stat {
  int32_t iteration;
  int32_t trials;           /* g2o levenbergIterations (1 for GN) */
  int32_t result;           /* SGB_RESULT_* */
  int32_t pcg_iters;        /* PCG iterations summed over the trials of this iteration */
  double chi2;              /* currentChi after the iteration (robustified) */
  double lambda;            /* LM lambda after the iteration */
  double rho;               /* last gain ratio */
  double chi2_before;       /* activeRobustChi2 at iteration start */
  double pcg_residual;      /* last relative preconditioned residual sqrt(r.z / r0.z0) */
} sgb_iter_stat;

/* sizes of the Hessian block structure after BlockSolver::buildStructure (no Schur;
 * the reference never marginalises) */
typedef struct sgb_structure_info {
  int32_t n_free;           /* vertices with a Hessian index */
  int32_t n_free_poses;
  int32_t n_free_landmarks;
  int32_t n_blocks;         /* stored upper-triangular blocks */
  int32_t scalar_dim;       /* 3*free poses + 2*free landmarks */
  int32_t n_active_pp;
  int32_t n_active_pl;
  int32_t coarse_nodes;     /* nodes of the two-level preconditioner's coarse space planned for this graph (0: block-Jacobi only) */
  int64_t block_values;     /* doubles in the concatenated block value array */
} sgb_structure_info;

/* per-phase device time of the last sgb_optimize / sgb_step (CUDA events on the handle's stream), ms */
typedef struct sgb_timings {
  double linearize_ms;      /* linearise + assemble (per-edge Jacobians, H, b, chi2) */
  double setup_ms;          /* damping, 2x2 landmark inverses, Schur diagonal + reduced rhs */
  double pcg_ms;            /* reduced-system PCG */
  double update_ms;         /* back-substitution, oplus, chi2 of the trial, LM control */
  double total_ms;
  int64_t pcg_iters;        /* total PCG iterations */
  int64_t trials;           /* total LM trials (GN: iterations) */
  int64_t linearizations;
  int64_t kernel_launches;  /* kernels launched by this library */
  /* wall time inside the persistent PCG kernel, split by phase of an iteration (barrier waits included):
   * [0] landmark-major pass t = W Hlp^T z, [1] pose-major pass w = S z and z.w, [2] the fused vector recurrences
   * (d, s, x, r, z) and r.z, [3] start-up (x = 0, r = b, z = M^-1 r) */
  double pcg_phase_ms[4];
} sgb_timings;

/* this rank's share of a row-block partitioned graph */
typedef struct sgb_partition_info {
  int32_t world, rank;
  int32_t n_poses;          /* free pose rows owned */
  int32_t n_landmarks;      /* free landmarks owned */
  int32_t n_pp, n_pl;       /* local edges (incident to an owned row) */
  int32_t n_pp_owned, n_pl_owned; /* edges whose chi2 is accounted on this rank */
  int64_t halo_pose_gathers;     /* off-rank pose-vector entries gathered per PCG iteration */
  int64_t halo_landmark_gathers; /* off-rank landmark-vector entries gathered per PCG iteration */
  int64_t hpp_entries, hpl_entries, hlp_entries; /* SELL entries including padding */
  int64_t hpp_blocks, hpl_blocks;                /* stored blocks */
} sgb_partition_info;

typedef struct sgb_handle sgb_handle;

const char* sgb_version(void);
/* number of usable CUDA devices (0 = none); never fails */
int32_t sgb_device_count(void);
void sgb_default_options(sgb_options* opt);

/* The first sgb_create of a process also raises glibc's mmap / trim thresholds (mallopt) so that the large host-side
 * symbolic arrays of sgb_set_graph are recycled between calls instead of being unmapped and page-faulted in again;
 * performance-only, process-wide, disabled by SGB_KEEP_HOST_MEMORY=0 in the environment. */
sgb_status sgb_create(const sgb_options* opt /* NULL = defaults */, sgb_handle** out);
void sgb_destroy(sgb_handle* h);
const char* sgb_last_error(const sgb_handle* h);

/* addVertex/addEdge of the whole graph + SparseOptimizer::initializeOptimization():
 * builds g2o's index mapping (active vertices sorted by id, fixed -> -1) and block
 * structure on the host, the symbolic scatter map, and uploads everything. */
sgb_status sgb_set_graph(sgb_handle* h, const sgb_graph_soa* g);
/* SparseOptimizer::updateInitialization(vset, eset) + OptimizationAlgorithm::updateStructure (the online path the
 * reference takes once per accepted key-frame, drone.cpp:152-156: `updateInitialization(new_vset, new_eset); push();
 * optimize(15, true)`): extends the graph of the last sgb_set_graph / sgb_update_graph by NEW vertices and NEW edges.
 * Only the delta crosses the boundary: the estimates of the existing vertices stay what they are on the device (the
 * result of the previous optimize()), the values of the existing edges are already there, the new ones are appended
 * in insertion order; the symbolic maps are re-derived on the host from the handle's index mirror and the values are
 * gathered on the device. New vertices are numbered after the existing ones (array index = old count + position);
 * edge endpoints index the extended arrays. Hessian indices follow the reference's batch rule (active vertices by id)
 * -- g2o itself appends new vertices in pointer order, which no two runs reproduce. Needs sgb_options.incremental.
 * Any backup of sgb_push is dropped (the reference pushes after updateInitialization). Single GPU. */
typedef struct sgb_graph_delta {
  int32_t n_new_poses;
  const int32_t* pose_id;    /* NULL: id = array index */
  const double* pose_est;    /* [3*n_new_poses] */
  const uint8_t* pose_fixed; /* NULL: none */
  int32_t n_new_landmarks;
  const int32_t* lm_id;      /* NULL: id = 10000000 + array index */
  const double* lm_est;      /* [2*n_new_landmarks] */
  const uint8_t* lm_fixed;
  int32_t n_new_pp;
  const int32_t* pp_i;       /* array indices into the EXTENDED vertex arrays */
  const int32_t* pp_j;
  const double* pp_z;
  const double* pp_info;
  const double* pp_phi;      /* NULL = no robust kernel */
  const int64_t* pp_seq;     /* insertion rank; NULL = after every existing edge, pose-pose before pose-line */
  int32_t n_new_pl;
  const int32_t* pl_pose;
  const int32_t* pl_lm;
  const double* pl_z;
  const double* pl_info;
  const int64_t* pl_seq;
} sgb_graph_delta;
sgb_status sgb_update_graph(sgb_handle* h, const sgb_graph_delta* d);
/* Multi-GPU, one process per GPU: every rank passes the SAME whole graph; rank r keeps the free-pose rows
 * [r*ceil(Pf/world), ...) of the reduced system and the landmarks first observed from them. Afterwards the ranks
 * exchange sgb_comm_get_handle() blobs (64 bytes each, e.g. with torch.distributed.all_gather) and call
 * sgb_comm_connect() with all of them in rank order; the PCG then gathers halo entries and reduces its dot
 * products directly through NVLink peer memory. Every rank must issue the same sequence of compute calls.
 * The reference has no counterpart (single process). world <= 8. */
sgb_status sgb_set_graph_partitioned(sgb_handle* h, const sgb_graph_soa* g, int32_t world, int32_t rank);
sgb_status sgb_comm_get_handle(sgb_handle* h, void* out64);
sgb_status sgb_comm_connect(sgb_handle* h, const void* handles /* world * 64 bytes, rank order */, int32_t world);
sgb_status sgb_get_partition_info(const sgb_handle* h, sgb_partition_info* out);
sgb_status sgb_get_structure_info(const sgb_handle* h, sgb_structure_info* out);
/* Hessian order: hessian index -> (kind 0 pose / 1 landmark, array index, scalar offset);
 * block list in column-major, row-ascending order (the order g2o's SparseBlockMatrix iterates);
 * hessian index per vertex (-1 = fixed or inactive). Any pointer may be NULL. */
sgb_status sgb_get_structure(const sgb_handle* h, int32_t* kind, int32_t* index, int32_t* offset,
                             int32_t* blk_row, int32_t* blk_col, int32_t* blk_nrows, int32_t* blk_ncols,
                             int32_t* pose_hidx, int32_t* lm_hidx);

/* Test hook = SparseOptimizer::computeActiveErrors + BlockSolver::buildSystem at the current
 * estimates: b [scalar_dim], block values in sgb_get_structure order (each block column-major,
 * like Eigen), chi2[0] = activeChi2, chi2[1] = activeRobustChi2. Any pointer may be NULL. */
sgb_status sgb_linearize(sgb_handle* h, double* b, double* Hblocks, double* chi2);
/* Test hook: one damped solve (H + lambda I) x = b at the current estimates (re-linearises).
 * x [scalar_dim] in Hessian order; pcg_iters / rel_residual may be NULL. */
sgb_status sgb_solve_once(sgb_handle* h, double lambda, double* x, int32_t* pcg_iters, double* rel_residual);

/* ---- LinearSolver level (the narrower drop-in) ------------------------------------------------------
 * What g2o::LinearSolver<MatrixType>::solve(const SparseBlockMatrix<MatrixType>& A, number_t* x, number_t* b) needs
 * (the object the reference builds as LinearSolverEigen at graphs.cpp:11,19): the caller -- g2o's own BlockSolver --
 * keeps linearisation, damping and the LM logic; only (A, b) -> x runs here (Schur elimination of the 2x2 blocks,
 * block-Jacobi PCG on the reduced 3x3-block system). A is the upper triangle in g2o's SparseBlockMatrix order: block
 * columns, rows ascending within a column, every block column-major (Eigen); 3x3 pose blocks first, then 2x2 landmark
 * blocks, off-diagonal blocks pose-pose or pose-landmark (what BlockSolver<-1,2> / BlockSolver<3,3> produce for this
 * repo's graphs). sgb_linear_set_pattern = LinearSolver::init() + the pattern analysis of the first solve (once per
 * optimize()); sgb_linear_solve = solve(): values in pattern order, b and x [3*n3 + 2*n2]; returns
 * SGB_ERR_SOLVE_FAILED where LinearSolverEigen returns false (matrix not positive definite). The handle's graph, if
 * any, is replaced. */
typedef struct sgb_block_matrix {
  int32_t n_block_cols;      /* = block rows */
  const int32_t* block_dim;  /* [n] 3 or 2, all 3s first */
  const int32_t* col_ptr;    /* [n+1] */
  const int32_t* row_idx;    /* [col_ptr[n]] block row <= block column, ascending within a column */
} sgb_block_matrix;
sgb_status sgb_linear_set_pattern(sgb_handle* h, const sgb_block_matrix* A);
sgb_status sgb_linear_solve(sgb_handle* h, const double* values, const double* b, double* x, int32_t* pcg_iters,
                            double* rel_residual);

/* g2o::OptimizationAlgorithm::computeMarginals(SparseBlockMatrix<MatrixX>& spinv, const std::vector<std::pair<int,int>>&
 * blockIndices) -- pure virtual in libg2o 2020.5.29 (g2o/core/optimization_algorithm.h), reached by
 * SparseOptimizer::computeMarginals; the reference's algorithms implement it through BlockSolver::computeMarginals ->
 * LinearSolver::solvePattern. Block (block_row[k], block_col[k]) (Hessian indices of sgb_get_structure) of the INVERSE
 * of the un-damped Hessian at the current estimates, k = 0..n_blocks-1; out_values receives the blocks one after the
 * other in request order, each column-major (dim(row) x dim(col), dim = 3 for poses, 2 for landmarks). Computed
 * column by column: H x = e_j solved on the device with the Schur + PCG path (lambda = 0), one solve per scalar
 * column of every distinct block column requested. SGB_ERR_SOLVE_FAILED when H is not positive definite (g2o returns
 * false). Single GPU. */
sgb_status sgb_compute_marginals(sgb_handle* h, int32_t n_blocks, const int32_t* block_row, const int32_t* block_col,
                                 double* out_values);

/* SparseOptimizer::optimize(iters, online): returns through *iters_done what g2o returns
 * (iterations run; 0 on Fail; -1 if not initialised). stats may be NULL or hold max_iters entries.
 * Estimates stay resident on the device and are also copied back (sgb_get_estimates). */
sgb_status sgb_optimize(sgb_handle* h, int32_t algo, int32_t max_iters, int32_t online,
                        int32_t* iters_done, sgb_iter_stat* stats);
/* One OptimizationAlgorithm::solve(iteration, online) -- the per-iteration plugin path. */
sgb_status sgb_step(sgb_handle* h, int32_t algo, int32_t iteration, int32_t online, int32_t* result,
                    sgb_iter_stat* stat);

/* estimates of ALL vertices, in the array order of sgb_set_graph */
sgb_status sgb_get_estimates(sgb_handle* h, double* pose_est, double* lm_est);
sgb_status sgb_set_estimates(sgb_handle* h, const double* pose_est, const double* lm_est);
/* SparseOptimizer::push / pop / discardTop on the device-resident estimates (drone.cpp:149,180,184) */
sgb_status sgb_push(sgb_handle* h);
sgb_status sgb_pop(sgb_handle* h);
sgb_status sgb_discard_top(sgb_handle* h);
/* computeActiveErrors(); chi2[0] = activeChi2() (what the caller's gate uses, drone.cpp:164-165),
 * chi2[1] = activeRobustChi2() */
sgb_status sgb_chi2(sgb_handle* h, double* chi2);
sgb_status sgb_get_timings(const sgb_handle* h, sgb_timings* out);

/* device-resident variant for benchmarking: the same as sgb_optimize but the estimates are first
 * reset on the device from the copy uploaded by sgb_set_graph, and nothing is copied back */
sgb_status sgb_optimize_resident(sgb_handle* h, int32_t algo, int32_t max_iters, int32_t* iters_done,
                                 sgb_iter_stat* stats);

/* Batched SparseOptimizer::optimize for n INDEPENDENT optimisers (n handles, each with its own graph set by
 * sgb_set_graph, all on the same device): one kernel launch, one thread block per graph, the complete LM / GN loop
 * on the device (no host round trip per trial). This is the B200 form of the reference's sliding-window use -- one
 * small landmark graph re-optimised per key-frame (drone.cpp:146-156, submap_loop_closer.cpp:256-262 keeps the
 * window small) -- when many such windows (robots, sessions, replays) are optimised at once. Results per handle are
 * what n separate sgb_optimize calls give (same arithmetic per row; reductions are block-local). The graphs should
 * be small enough for one thread block each (a few thousand vertices); larger ones still run, just slowly.
 * iters_done[n]: g2o's optimize() return value per graph; last_stats[n] (may be NULL): the last iteration's stats.
 * Estimates stay on the device: read them with sgb_get_estimates per handle. */
sgb_status sgb_optimize_batch(sgb_handle* const* handles, int32_t n, int32_t algo, int32_t max_iters,
                              int32_t* iters_done, sgb_iter_stat* last_stats);
/* the same with the estimates of every graph first reset on the device to the ones given to sgb_set_graph
 * (benchmarking: repeated runs from the same start, nothing copied back) */
sgb_status sgb_optimize_batch_resident(sgb_handle* const* handles, int32_t n, int32_t algo, int32_t max_iters,
                                       int32_t* iters_done, sgb_iter_stat* last_stats);

/* ---- graph fixture / wire format: g2o text files with the types of this path ----
 * VERTEX_SE2, EDGE_SE2, FIX as g2o writes them; VERTEX_RHOTHETA id rho theta and
 * EDGE_SE2_RHOTHETA pose line rho theta I11 I12 I22 for the reference's custom types (tags registered at
 * vertex_rhotheta.cpp:43 / edge_se2_rhotheta.cpp:24, whose read/write are empty in the reference);
 * ROBUST_KERNEL_DCS k delta for the DCS kernel of the k-th EDGE_SE2 line. Host only: works without a GPU.
 * sgb_g2o_load: *out owns the arrays; sgb_g2o_view fills an sgb_graph_soa with pointers into it (valid until
 * sgb_g2o_free). errbuf (may be NULL) receives "file:line: message" on failure. */
typedef struct sgb_graph_file sgb_graph_file;
sgb_status sgb_g2o_load(const char* path, sgb_graph_file** out, char* errbuf, int32_t errlen);
void sgb_g2o_view(const sgb_graph_file* f, sgb_graph_soa* out);
void sgb_g2o_free(sgb_graph_file* f);
sgb_status sgb_g2o_save(const char* path, const sgb_graph_soa* g);

/* ---- device-resident graph values -------------------------------------------------------------
 * sgb_set_graph with the VALUES (estimates, measurements, information matrices) already in device memory: the host
 * symbolic phase only needs the index arrays of `indices` (n_*, ids, fixed flags, pp_i/pp_j, pl_pose/pl_lm, *_seq;
 * its value pointers are ignored and may be NULL); the values are gathered on the device. Edge k of the index arrays
 * reads its values from device slot pp_slot[k] / pl_slot[k] (HOST int arrays; NULL = slot k), so a store that keeps
 * removed edges in place can hand over the active ones without compacting its value arrays. Single GPU. */
typedef struct sgb_device_values {
  const double* pose_est;   /* DEVICE [3*n_poses] */
  const double* lm_est;     /* DEVICE [2*n_landmarks] */
  const double* pp_z;       /* DEVICE [3*slots] */
  const double* pp_info;    /* DEVICE [6*slots] */
  const double* pp_phi;     /* DEVICE [slots] or NULL */
  const int32_t* pp_slot;   /* HOST [n_pp] or NULL */
  const double* pl_z;       /* DEVICE [2*slots] */
  const double* pl_info;    /* DEVICE [3*slots] */
  const int32_t* pl_slot;   /* HOST [n_pl] or NULL */
  int32_t has_robust;       /* any pp_phi > 0 */
  int32_t reserved;
} sgb_device_values;
sgb_status sgb_set_graph_device(sgb_handle* h, const sgb_graph_soa* indices, const sgb_device_values* dv);

/* ---- the pose graph and its edits, on the device ----------------------------------------------
 * Device-resident counterpart of the reference's PoseGraph (include/graphs.h:27-40: deque of pose + odometry edge,
 * all_closures) with the three edits the reference makes around setup_pose_opt's optimiser, so that a loop closure
 * costs an upload of ONE edge instead of the whole pose graph:
 *   sgb_pg_append_from_lm   submap_loop_closer.cpp:206-223  copy the newly optimised landmark-graph poses: odometry
 *                           edge measurement = (previous lm estimate)^-1 * (this lm estimate), information copied,
 *                           new vertex estimate = previous pose-graph estimate * measurement (chained)
 *   sgb_pg_add_closure      submap_loop_closer.cpp:272-285  EdgeSE2 between two existing vertices, DCS kernel
 *   sgb_pg_optimize         submap_loop_closer.cpp:286-288 / log_runner.cpp:203-204  initializeOptimization +
 *                           optimize(20): host symbolic phase on the store's index mirror, values gathered on the device
 *   sgb_pg_prune_closures   log_runner.cpp:182-190  computeError + chi2() of every closure (no robust kernel);
 *                           chi2 > threshold (11.345 in the reference) removes the edge
 * Vertices are addressed by their position in the store (0 = the fixed first pose, drone.cpp:68-75). */
typedef struct sgb_pose_graph sgb_pose_graph;
sgb_status sgb_pg_create(int32_t device /* -1 = current */, sgb_pose_graph** out);
void sgb_pg_destroy(sgb_pose_graph* pg);
const char* sgb_pg_last_error(const sgb_pose_graph* pg);
/* empties the store and adds the fixed first vertex */
sgb_status sgb_pg_reset(sgb_pose_graph* pg, int32_t first_id, const double first_est[3]);
/* lm: a handle whose graph is the landmark graph (estimates on the device); lm_first = array index of the first pose
 * to copy (>= 1: its predecessor supplies the relative measurement), count poses are copied; ids (HOST, NULL =
 * consecutive after the last id); info HOST [6*count] = upper triangles of the landmark graph's odometry edges */
sgb_status sgb_pg_append_from_lm(sgb_pose_graph* pg, sgb_handle* lm, int32_t lm_first, int32_t count,
                                 const int32_t* ids, const double* info);
/* the same edit from HOST estimates ([3*(count+1)]: predecessor first) when the landmark graph lives elsewhere */
sgb_status sgb_pg_append_from_host(sgb_pose_graph* pg, const double* lm_est, int32_t count, const int32_t* ids,
                                   const double* info);
sgb_status sgb_pg_add_closure(sgb_pose_graph* pg, int32_t from, int32_t to, const double z[3], const double info[6],
                              double dcs_phi, int32_t* closure_index /* may be NULL */);
/* solver: any handle on the same device (its previous graph is replaced); the optimised estimates are written back
 * into the store on the device. iters_done / stats as in sgb_optimize. */
sgb_status sgb_pg_optimize(sgb_pose_graph* pg, sgb_handle* solver, int32_t algo, int32_t max_iters,
                           int32_t* iters_done, sgb_iter_stat* stats);
/* chi2_out [n_closures] (may be NULL): chi2 of every closure ever added (removed ones: their chi2 when removed);
 * active_out [n_closures] (may be NULL): 1 = still in the graph */
sgb_status sgb_pg_prune_closures(sgb_pose_graph* pg, double threshold, int32_t* n_removed, double* chi2_out,
                                 uint8_t* active_out);
typedef struct sgb_pg_info {
  int32_t n_poses, n_edges /* odometry + closures, removed ones included */, n_closures, n_active_closures;
  double last_edit_ms;      /* device time of the kernels of the last edit (CUDA events) */
  int64_t kernel_launches;  /* kernels launched by the store so far */
} sgb_pg_info;
sgb_status sgb_pg_get_info(const sgb_pose_graph* pg, sgb_pg_info* out);
/* copies the store to the host; any pointer may be NULL. pp_* cover all n_edges slots in insertion order. */
sgb_status sgb_pg_download(sgb_pose_graph* pg, int32_t* pose_id, double* pose_est, int32_t* pp_i, int32_t* pp_j,
                           double* pp_z, double* pp_info, double* pp_phi, uint8_t* pp_active, uint8_t* pp_is_closure);

/* ---- information-matrix producers as batch kernels --------------------------------------------
 * Stateless: HOST buffers in and out, `device` = CUDA ordinal (-1 = current); kernel_ms (may be NULL) receives the
 * device time of the kernels alone (CUDA events), the call itself includes the copies.
 *
 * sgb_odom_information: OdomErrorPropagator<double> (include/odom_error_propagator.h:18-46) run over n_seg key-frame
 * intervals -- reset(), step() for every odometry delta of the interval (drone.cpp:84,143) -- giving per interval the
 * odometry edge's measurement (accumulated pose, drone.cpp:127), the covariance and its inverse, the edge's
 * information (drone.cpp:128). deltas [3*seg_ptr[n_seg]] (dx, dy, dtheta in the previous frame), seg_ptr [n_seg+1]. */
sgb_status sgb_odom_information(int32_t device, const double* deltas, const int32_t* seg_ptr, int32_t n_seg,
                                double std_x, double std_y, double std_w, double* z_out /* [3*n_seg] */,
                                double* cov_out /* [9*n_seg] row-major, may be NULL */,
                                double* info_out /* [6*n_seg] upper triangle */, double* kernel_ms);
/* sgb_scan_point_covariances: the covariance of every scan point of a multi-scan window in the frame of the newest
 * scan (src/multicloud2.cpp:56-83, single precision like the reference): n_windows windows of n_scans scans of
 * scan_size beams. deltas [n_windows][n_scans-1][3] = odometry between consecutive scans (scan i is propagated through
 * deltas i..n_scans-2); beam_cos_sin [2*scan_size]; pts [n_windows][n_scans][scan_size][2] (non-finite = no return).
 * Out per point: cov [4] row-major, rhotheta [2], valid (0 for non-finite points, whose outputs are zero). */
sgb_status sgb_scan_point_covariances(int32_t device, const double* deltas, int32_t n_windows, int32_t n_scans,
                                      int32_t scan_size, const float* beam_cos_sin, const float* pts, float std_x,
                                      float std_y, float std_w, float var_r, float* cov_out, float* rhotheta_out,
                                      uint8_t* valid_out, double* kernel_ms);
/* sgb_line_fit_information: LineSegment::leastSqFit (src/ls_extractor/src/impl/smc.cpp:30-68, single precision) for
 * n_seg segments: rho/theta, its 2x2 covariance from the per-point covariances, and the pose-line edge information
 * cov.cast<double>().inverse() (drone.cpp:203). pts [2*seg_ptr[n_seg]], pcov [4*...] row-major, seg_ptr [n_seg+1]. */
sgb_status sgb_line_fit_information(int32_t device, const float* pts, const float* pcov, const int32_t* seg_ptr,
                                    int32_t n_seg, float* rhotheta_out /* [2*n_seg] */, float* cov_out /* [4*n_seg] */,
                                    double* info_out /* [3*n_seg] 11,12,22 */, double* kernel_ms);
/* message of the last failure of a stateless call on this thread */
const char* sgb_frontend_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SGB_CAPI_H */
