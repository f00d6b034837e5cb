/*
 * sgo_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A dependency-free restatement of what sparse-gslam executes when it calls
 * g2o::SparseOptimizer::optimize() on its two graphs
 * (reference: src/sparse_gslam/src/graphs.cpp:9-23, callers drone.cpp:146-156,
 * submap_loop_closer.cpp:286-288, log_runner.cpp:203-204).
 *
 * PARITY UNPINNED: the arithmetic of this path lives in g2o (libg2o-release
 * 2020.5.29) + Eigen 3.3 SimplicialLDLT, which are NOT vendored under
 * /root/reference and cannot be built here; the reference has no tests, golden
 * vectors or stored graphs for this path. This oracle
 * restates the published g2o algorithm and the in-repo
 * custom types (g2o_bindings/edge_se2_rhotheta.cpp:9-16,
 * g2o_bindings/vertex_rhotheta.cpp:28-34, ls_extractor/utils.h:23-45). It is
 * cross-checked against an independent numpy/scipy restatement
 * (oracle/py_oracle.py) and sympy-derived known answers, not against g2o.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef SGO_ORACLE_H
#define SGO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same SoA convention as include/sgb_capi.h (kept separate on purpose: the
 * oracle must not include product headers). Edges reference vertices by ARRAY
 * INDEX; ids only define g2o's vertex ordering (sorted by id). */
typedef struct sgo_graph {
  int32_t n_poses;
  const int32_t* pose_id;    /* may be NULL: id = index */
  const double* pose_est;    /* [3*n_poses] x,y,theta */
  const uint8_t* pose_fixed; /* may be NULL: none fixed */
  int32_t n_landmarks;
  const int32_t* lm_id;      /* may be NULL: id = 10000000 + index (drone.h:22) */
  const double* lm_est;      /* [2*n_landmarks] rho,theta */
  const uint8_t* lm_fixed;   /* may be NULL */
  int32_t n_pp;              /* EdgeSE2 */
  const int32_t* pp_i;
  const int32_t* pp_j;
  const double* pp_z;        /* [3*n_pp] dx,dy,dtheta */
  const double* pp_info;     /* [6*n_pp] upper triangle 11,12,13,22,23,33 */
  const double* pp_phi;      /* DCS delta per edge, <=0 = no robust kernel; may be NULL */
  const int64_t* pp_seq;     /* insertion rank (g2o internalId); may be NULL */
  int32_t n_pl;              /* EdgeSE2RhoTheta */
  const int32_t* pl_pose;
  const int32_t* pl_lm;
  const double* pl_z;        /* [2*n_pl] rho,theta */
  const double* pl_info;     /* [3*n_pl] 11,12,22 */
  const int64_t* pl_seq;     /* may be NULL: all pp edges first, then pl */
} sgo_graph;

enum { SGO_ALGO_LM = 0, SGO_ALGO_GN = 1 };
enum { SGO_JAC_G2O_NUMERIC = 0, SGO_JAC_ANALYTIC = 1 };

typedef struct sgo_iter_stat {
  int32_t iteration;
  int32_t trials;      /* levenbergIterations (LM) / 1 (GN) */
  int32_t result;      /* 1 OK, 2 Terminate, -1 Fail */
  int32_t pad;
  double chi2;         /* currentChi after the iteration (robustified) */
  double lambda;       /* lambda after the iteration */
  double rho;          /* last rho */
  double chi2_before;  /* activeRobustChi2 at iteration start */
} sgo_iter_stat;

typedef struct sgo_handle sgo_handle;

sgo_handle* sgo_create(void);
void sgo_destroy(sgo_handle*);
/* copies everything; caller may free after return */
int sgo_set_graph(sgo_handle*, const sgo_graph*);
/* SparseOptimizer::initializeOptimization(): returns 1 ok, 0 failure (empty) */
int sgo_initialize(sgo_handle*);

/* Hessian block structure after BlockSolver::buildStructure (no Schur).
 * n_free vertices in hessian order; blocks in column-major, row-ascending order. */
int sgo_num_free(sgo_handle*);
int sgo_num_blocks(sgo_handle*);
int sgo_scalar_dim(sgo_handle*);
/* hessian index -> (kind 0 pose/1 landmark, array index), scalar offset */
void sgo_get_order(sgo_handle*, int32_t* kind, int32_t* index, int32_t* offset);
/* per block: row, col, nrows, ncols */
void sgo_get_blocks(sgo_handle*, int32_t* row, int32_t* col, int32_t* nr, int32_t* nc);
/* hessian index per vertex (-1 fixed / inactive) */
void sgo_get_hessian_index(sgo_handle*, int32_t* pose_hidx, int32_t* lm_hidx);

/* computeActiveErrors + buildSystem at the current estimates.
 * Outputs (any may be NULL): per-edge errors/Jacobians in SoA edge order,
 * dense b [scalar_dim], block values concatenated in sgo_get_blocks order
 * (each block column-major like Eigen), chi2[0]=activeChi2, chi2[1]=activeRobustChi2 */
int sgo_linearize(sgo_handle*, int jac_mode,
                  double* pp_err /*3*/, double* pp_A /*9 row-major*/, double* pp_B /*9*/,
                  double* pl_err /*2*/, double* pl_A /*6 row-major 2x3*/, double* pl_B /*4*/,
                  double* b, double* Hblocks, double* chi2);
int64_t sgo_block_values_size(sgo_handle*);

/* SparseOptimizer::optimize(iters, online=false). Returns iterations run, 0 on Fail, -1 if not initialised */
int sgo_optimize(sgo_handle*, int algo, int iters, int jac_mode, sgo_iter_stat* stats /*[iters] or NULL*/);
void sgo_get_estimates(sgo_handle*, double* pose_est, double* lm_est);
void sgo_set_estimates(sgo_handle*, const double* pose_est, const double* lm_est);
/* computeActiveErrors(); chi2[0]=activeChi2 (un-robustified), chi2[1]=activeRobustChi2 */
void sgo_chi2(sgo_handle*, double* chi2);
/* one damped solve with the exact LDLt at the current linearisation: (H + lambda I) x = b */
int sgo_solve_once(sgo_handle*, int jac_mode, double lambda, double* x);
/* timing / diagnostics of the last optimize: factor nnz, seconds in linearise / factor+solve */
void sgo_last_profile(sgo_handle*, double* out /*[4]: nnzL, t_lin, t_solve, t_total*/);

/* ---- rows either side of the optimiser (oracle/sgo_frontend.cpp) ---- */
/* submap_loop_closer.cpp:206-223; lm_est [3*(count+1)] predecessor first; z_out / est_out [3*count] */
void sgo_pg_append(const double* prev_pg_est, const double* lm_est, int32_t count, double* z_out, double* est_out);
/* log_runner.cpp:182-184; per edge k: vertices ei[k], ej[k], z [3*n], info6 [6*n] */
void sgo_closure_chi2(const double* est, const int32_t* ei, const int32_t* ej, const double* z, const double* info6,
                      int32_t n, double* chi_out);
/* odom_error_propagator.h + drone.cpp:84,127-128,143 */
void sgo_odom_information(const double* deltas, const int32_t* seg_ptr, int32_t n_seg, double std_x, double std_y,
                          double std_w, double* z_out, double* cov_out /* may be NULL */, double* info_out);
/* multicloud2.cpp:56-83 */
void sgo_scan_point_covariances(const double* deltas, int32_t n_windows, int32_t n_scans, int32_t scan_size,
                                const float* beam_cos_sin, const float* pts, float std_x, float std_y, float std_w,
                                float var_r, float* cov_out, float* rhotheta_out, uint8_t* valid_out);
/* smc.cpp:30-68 + drone.cpp:203 */
void sgo_line_fit_information(const float* pts, const float* pcov, const int32_t* seg_ptr, int32_t n_seg,
                              float* rhotheta_out, float* cov_out, double* info_out);

#ifdef __cplusplus
}
#endif
#endif
